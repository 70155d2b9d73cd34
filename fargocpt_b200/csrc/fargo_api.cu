// fargo_api.cu — the C ABI of include/fargo_b200.h on top of the sm_100a kernels.
//
// One fargo_ctx == one radial slab == one GPU (the reference's MPI rank).  All device work of a context
// runs on its own stream; the only per-step host<->device traffic is the CFL scalar (8 bytes back) and the
// per-ring damping factors / body positions (a few hundred bytes forth), exactly the scalars the reference's
// host code exchanges with its loop nests.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <string>
#include <vector>
#include <algorithm>

#include "fargo_dev.h"
#include "../../host/fargo_pvte.h" // the PVTE lookup tables are built on the host (pvte::initializeLookupTables)
#include "kernels_ring.cuh"
#include "kernels_ringsum.cuh"
#include "kernels_source.cuh"
#include "kernels_transport.cuh"
#include "kernels_azimuthal.cuh"
#include "kernels_fused.cuh"
#include "kernels_diag.cuh"
#include "fargo_selftest.cuh"

// ---------------------------------------------------------------------------------------------
// error handling: the reference die()s; we return non-zero and keep the message (thread-local)
static thread_local std::string g_err;
static int fail(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}
#define CUDA_OK(expr)                                                                                \
    do {                                                                                             \
	cudaError_t _e = (expr);                                                                     \
	if (_e != cudaSuccess)                                                                       \
	    return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// NCCL through dlopen: libfargo_b200.so must load on a box without NCCL (single GPU) and must share the NCCL
// already loaded by the host process (torch bundles one) instead of pulling in a second copy.
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclFloat64 = 8, ncclMin = 3 }; // nccl.h: ncclDataType_t / ncclRedOp_t values (stable since NCCL 2.0)
struct NcclApi {
    int (*GetUniqueId)(ncclUniqueId *);
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    int (*CommDestroy)(ncclComm_t);
    int (*GroupStart)();
    int (*GroupEnd)();
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(int);
    bool ok = false;
};
static NcclApi g_nccl;
static int load_nccl()
{
    if (g_nccl.ok)
	return 0;
    void *h = RTLD_DEFAULT;
    if (!dlsym(RTLD_DEFAULT, "ncclCommInitRank")) {
	const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
	h = nullptr;
	for (int k = 0; names[k] && !h; ++k)
	    h = dlopen(names[k], RTLD_NOW | RTLD_GLOBAL);
	if (!h)
	    return fail("NCCL not found: %s", dlerror());
    }
#define SYM(field, name)                                   \
    *(void **)(&g_nccl.field) = dlsym(h, name);            \
    if (!g_nccl.field)                                     \
	return fail("NCCL symbol %s missing", name);
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(AllReduce, "ncclAllReduce")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.ok = true;
    return 0;
}
#define NCCL_OK(expr)                                                                          \
    do {                                                                                       \
	int _r = (expr);                                                                       \
	if (_r != 0)                                                                           \
	    return fail("%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?"); \
    } while (0)

// ---------------------------------------------------------------------------------------------
struct fargo_ctx {
    DevView v;	   // host copy of what the kernels receive by value
    int device;
    cudaStream_t stream;
    ncclComm_t comm = nullptr;
    long long launches = 0;
    // host geometry (local view incl. 2 extra entries) for host-side ring factors
    std::vector<double> h_radii, h_rinf, h_rsup, h_rmed, h_cs_iso, h_inv_omega_k;
    std::vector<double *> dev_allocs;
    // state.  energy, v_rad and v_azi are double-buffered: the fused kernels read one buffer and write the other
    // (their column windows overlap on reads), and Transport cannot write v_azi where it still reads the residual
    // velocity.  ecur / vcur name the buffer holding the state between stages; the staged source kernels keep
    // their mid-step velocities in the other v buffer (v_mid).
    double *sigma, *eb[2], *vrb[2], *vpb[2];
    int ecur = 0, vcur = 0;
    bool v_mid = false;
    double *sigma0, *energy0, *vr0, *vp0;
    double *qplus, *qminus;
    // stage scratch
    double *pot, *qr, *qphi, *nu, *divv, *trr, *tpp, *trp, *nusig, *nusig_rp, *cf_r, *cf_phi;
    double *t_sigma, *t_rmp, *t_rmm, *t_amp, *t_amm, *t_e; // after the radial sweep
    double *vmean, *vconst, *expf_s, *expf_v, *d_dt, *scratch, *force4;
    // two-pass CFL reduction (kernels_ring.cuh:k_cfl_screen / k_cfl_candidates): per-block maxima of the screen, d_cfl_l = their max
    double *cfl_bmax = nullptr, *d_cfl_l = nullptr;
    double *mon_rings = nullptr; // fargo_monitor_disk: MD_N per-ring sums of the whole mesh
    double *massflow = nullptr;	 // fargo_track_massflow: MASSFLOW grid [nr + 1][ns]
    double *dmass = nullptr;	 // fargo_track_damping_mass: [4][ns] inner creation / removal, outer creation / removal, per column
    double *bflow = nullptr;	 // fargo_track_boundary_flow: [4][ns] inner inflow / outflow, outer inflow / outflow, per column
    bool track_massflow = false;
    bool keep_pot = false;	 // fargo_keep_potential: fused kicks also store the POTENTIAL grid
    int pot_state = 0;		 // the POTENTIAL grid: 0 zeros (no kick yet, like the reference's), 1 of the last kick, 2 older
    // FARGO_B200_FUSE_ARTVISC=1: the artificial-viscosity stage runs inside k_fused_sources<.., AV = true> instead of as its own
    // kernel.  Bit-identical, 56 bytes per cell less traffic — and slower (5.22 against 2.76 + 2.14 ms at 8192 x 16384: 64 bytes
    // of spills at 128 registers, and these kernels wait on FP64 chains, not on DRAM; profiles/r02_v12_*), so not the default.
    bool fuse_artvisc = false;
    int cfl_mode = 1; // FARGO_B200_CFL=full: 0 (k_cfl over every cell), screen: 1 (default), check: 2 (both, and they must agree)
    // damping zones folded into the azimuthal transport kernel's epilogue (fargo_step of an Euler step; AzSegs::dmask)
    int *d_dmask = nullptr;	     // per ring: 2 bits per field
    std::vector<int> h_dmask;	     // what d_dmask holds
    bool fold_damping = true;	     // FARGO_B200_FOLD_DAMPING=0 keeps k_damping as its own pass
    bool fold_armed = false;	     // the next launch_transport applies the damping of this step (factors uploaded)
    bool damp_folded = false;	     // ... and has done so: the final boundary call skips k_damping
    double *partials; // per-block partial sums of the accretion / monitor / disk-on-body reductions (sized from their launch grids)
    size_t partials_n;
    double *hstale = nullptr; // leapfrog: scale height of the first kick's viscosity stage (see fargo_kick)
    bool h_stale = false;
    // pre-accretion Sigma / e of rings [pre_lo, pre_hi) (fargo_dev.h:PreState): written by fargo_accrete_kley, consumed by
    // the next source-term stage, dropped when the derived quantities are recalculated (fargo_finish_step)
    double *sig_pre = nullptr, *e_pre = nullptr;
    int pre_lo = 0, pre_hi = 0;
    int *nshift;
    double *h_pin; // pinned host staging for the CFL scalar + ring factors
    bool visc_const_filled = false;
    cudaEvent_t ev_pin = nullptr; // completion of the last H2D copy out of h_pin
    int az_slots, rad_chunk, fs_R, n_sm;
    bool ringsum_scan = false;  // ring sums by k_ring_sum_scan (a warp per ring) instead of the one-thread-per-ring chain
    int rm_chunk = 0;	       // columns per TMA chunk of k_ring_mean (0: generic kernel)
    int rm_side_chunk = 0;     // ... of the launch that runs beside the radial sweep (0: rm_chunk there too; see fargo_ctx_create)
    size_t rm_smem = 0;
    cudaStream_t stream2 = nullptr; // side stream: the transport ring means run beside the radial sweep
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // Ghost-ring exchange over peer memory (fargo_stage_halo): every rank owns an inbox its neighbours write into
    // (double-buffered by step parity) and two arrival counters; the neighbours' inboxes are mapped with CUDA IPC.
    struct Halo {
	bool p2p = false;		   // peer path set up on every rank (else: ncclSend / ncclRecv)
	double *inbox = nullptr;	   // [parity 2][side 2: from prev, from next][field 4][CPUOVERLAP * ns], then 2 counters
	unsigned long long *flags = nullptr; // inside the inbox allocation: steps received from prev / from next
	void *peer_base[2] = {nullptr, nullptr}; // prev's / next's inbox allocation as mapped here
	size_t field_len = 0;		   // CPUOVERLAP * ns doubles
	unsigned long long seq = 0;	   // halo exchanges started
	bool pushed = false;		   // this step's edge rings are on their way (launch_transport), not yet received
	unsigned int *done = nullptr;	   // edge warps of the running transport launch that have finished (device)
    } halo;
    // asynchronous snapshots (fargo_snapshot_async): device-side copies of the four state fields and a copy stream
    double *snap[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t stream_snap = nullptr;
    cudaEvent_t ev_snap_ready = nullptr, ev_snap_done = nullptr;
    bool force_staged = false;
    cudaEvent_t ev_user[4] = {nullptr, nullptr, nullptr, nullptr}; // fargo_event_record slots
    // optional per-kernel device timing (bench.py roofline): CUDA events on the launching stream
    bool profiling = false;
    // host time between the CFL result reaching the host and the step's first kernel launch (the GPU idles meanwhile)
    double t_cfl_done = 0.0, host_turnaround_ms = 0.0;
    long long host_turnaround_n = 0;
    struct KStat { std::string name; double ms = 0; long long n = 0; };
    std::vector<KStat> kstats;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> pending;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
};

// "A" = the buffer holding the velocities between stages, "B" = the other one (mid-step velocities of the staged path)
#define VRA(c) ((c)->vrb[(c)->vcur])
#define VPA(c) ((c)->vpb[(c)->vcur])
#define VRB(c) ((c)->vrb[1 - (c)->vcur])
#define VPB(c) ((c)->vpb[1 - (c)->vcur])
#define EN(c) ((c)->eb[(c)->ecur])

static int kstat_id(fargo_ctx *c, const char *name)
{
    for (size_t k = 0; k < c->kstats.size(); ++k)
	if (c->kstats[k].name == name)
	    return (int)k;
    fargo_ctx::KStat s;
    s.name = name;
    c->kstats.push_back(s);
    return (int)c->kstats.size() - 1;
}
static void prof_begin(fargo_ctx *c, const char *name, cudaStream_t strm)
{
    std::pair<cudaEvent_t, cudaEvent_t> ev;
    if (!c->ev_pool.empty()) {
	ev = c->ev_pool.back();
	c->ev_pool.pop_back();
    } else {
	cudaEventCreate(&ev.first);
	cudaEventCreate(&ev.second);
    }
    cudaEventRecord(ev.first, strm);
    c->pending.push_back({kstat_id(c, name), ev});
}
static void prof_end(fargo_ctx *c, cudaStream_t strm) { cudaEventRecord(c->pending.back().second.second, strm); }
static void prof_collect(fargo_ctx *c)
{
    cudaStreamSynchronize(c->stream);
    if (c->stream2)
	cudaStreamSynchronize(c->stream2);
    for (auto &p : c->pending) {
	float ms = 0;
	cudaEventElapsedTime(&ms, p.second.first, p.second.second);
	c->kstats[p.first].ms += ms;
	c->kstats[p.first].n += 1;
	c->ev_pool.push_back(p.second);
    }
    c->pending.clear();
}

static int dalloc(fargo_ctx *c, double **p, size_t n)
{
    CUDA_OK(cudaMalloc((void **)p, (n ? n : 1) * sizeof(double)));
    CUDA_OK(cudaMemsetAsync(*p, 0, (n ? n : 1) * sizeof(double), c->stream));
    c->dev_allocs.push_back(*p);
    return 0;
}
static int upload_vec(fargo_ctx *c, const double **dst, const std::vector<double> &h)
{
    double *d;
    if (dalloc(c, &d, h.size()))
	return 1;
    CUDA_OK(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    *dst = d;
    return 0;
}

// Rings per march of a ring-marching kernel.  A CTA marches R rings after `warm` warm-up rings; the grid has
// ctas_x * ceil(nr / R) CTAs for `slots` resident CTA slots.  Model: one wave costs (R + warm); several waves are
// scheduled dynamically, so they cost their fractional count plus half a CTA of tail.
static int rings_per_march(int nr, int ctas_x, int slots, int warm)
{
    int best_R = nr;
    double best = 1e300;
    for (int R = (nr < 4 ? nr : 4); R <= nr; ++R) {
	const int y = (nr + R - 1) / R;
	if (R > 4 && (nr + R - 2) / (R - 1) == y)
	    continue; // same number of bands as R-1: the smaller R is the balanced choice
	const double ctas = (double)ctas_x * y;
	const double waves = ctas / slots;
	// A few waves of equal CTAs finish in whole waves, and SMs do not all run at the same speed, which whole waves cannot
	// absorb (measured on the azimuthal kernel: 1 CTA per slot 5-10 % slower than 8 per slot, 4 waves 2 % slower than 26);
	// with many waves the block scheduler evens the SMs out to about half a CTA of tail.
	const double cost = (R + warm) * (waves <= 4.0 ? ceil(waves) * (1.0 + 0.10 / sqrt(ceil(waves))) : waves + 0.5);
	if (cost < best * 0.999) {
	    best = cost;
	    best_R = R;
	}
    }
    return best_R < 1 ? 1 : best_R;
}

static inline unsigned cells_grid(long long n, int block = 256) { return (unsigned)((n + block - 1) / block); }

// split.cpp:38-87
// SplitDomain (split.cpp:38-87): contiguous rings per rank plus CPUOVERLAP ghost rings per interior side, and the loop
// bounds of a slab.  The reference gives every rank the same number of rings; here the cut points balance COST: a ring
// inside a damping zone costs more (~1.85 % of a ring's step per damped field in the azimuthal kernel's epilogue, measured on
// B200), and the damping zones sit on the first and last ranks — at 8 GPUs their step was 6 % longer than everyone
// else's.  Results do not depend on where the cuts are (constants.h:17; tests/test_gpu_multi.py holds N ranks to 1 rank
// bit for bit).  FARGO_B200_SPLIT=equal restores the reference's cut points.
// cut[r] = first ring owned by rank r (cut[np] = nrad); pure host arithmetic, no device needed
extern "C" int fargo_split_cuts(const fargo_params *params, const double *radii, int np, int *cut /* np + 1 */)
{
    const fargo_params &p = *params;
    const int nrad = p.nrad;
    if (np < 1)
	return fail("bad number of ranks %d", np);
    const int size_low = nrad / np, size_high = size_low + 1, rem = nrad % np;
    if (np > 1 && size_low < 2 * FARGO_CPUOVERLAP)
	return fail("The number of processes is too large or the mesh is radially too narrow.");
    for (int r = 0; r <= np; ++r) // the reference's equal split (split.cpp:38-55)
	cut[r] = r < rem ? size_high * r : size_high * rem + (r - rem) * size_low;
    const char *env = getenv("FARGO_B200_SPLIT");
    if (np > 1 && p.damping && radii && !(env && strcmp(env, "equal") == 0)) {
	int nf = 0;
	for (int k = 0; k < 2; ++k)
	    nf += (p.damp_vrad[k] != FARGO_DAMP_NONE) + (p.damp_vazi[k] != FARGO_DAMP_NONE) + (p.damp_sigma[k] != FARGO_DAMP_NONE) +
		  (p.adiabatic && p.damp_energy[k] != FARGO_DAMP_NONE);
	// (0.025 per field when k_damping was a pass of its own; folded into the azimuthal kernel's epilogue with its initial-field
	// columns staged through shared memory a damped ring costs 7.4 % more than an undamped one with all four fields damped:
	// 8-GPU kernel times of the edge rank against the interior ranks, profiles/r02_m8_bench_scaling.jsonl)
	const double wd = 0.0185 * 0.5 * nf; // both sides configured: nf counts every damped field twice
	std::vector<double> cum(nrad + 1, 0.0);
	for (int i = 0; i < nrad; ++i) {
	    const double r = 0.5 * (radii[i] + radii[i + 1]);
	    const bool damped = (p.damping_inner_limit > 1.0 && r < p.rmin * p.damping_inner_limit) ||
				(p.damping_outer_limit < 1.0 && r > p.rmax * p.damping_outer_limit);
	    cum[i + 1] = cum[i] + 1.0 + (damped ? wd : 0.0);
	}
	std::vector<int> w(np + 1, 0);
	w[np] = nrad;
	bool ok = true;
	for (int r = 1; r < np; ++r) {
	    const double target = cum[nrad] * r / np;
	    int i = w[r - 1];
	    while (i < nrad && cum[i] < target)
		++i;
	    w[r] = i;
	}
	for (int r = 0; r < np; ++r)
	    ok = ok && (w[r + 1] - w[r] >= 2 * FARGO_CPUOVERLAP);
	if (ok)
	    for (int r = 0; r <= np; ++r)
		cut[r] = w[r];
    }
    return 0;
}

static int split_domain(DevView &v, int nrad, int rank, int np, int *imax_out, const double *radii, const fargo_params &p)
{
    (void)nrad;
    std::vector<int> cut(np + 1, 0);
    if (fargo_split_cuts(&p, radii, np, cut.data()))
	return 1;
    int imin = cut[rank], imax = cut[rank + 1] - 1;
    if (rank > 0)
	imin -= FARGO_CPUOVERLAP;
    if (rank < np - 1)
	imax += FARGO_CPUOVERLAP;
    v.imin = imin;
    v.nr = imax - imin + 1;
    const bool first = rank == 0, last = rank == np - 1;
    v.zero_no_ghost = first ? 1 : 0;
    v.one_no_ghost_vr = first ? 2 : 1;
    v.max_no_ghost = v.nr - (last ? 1 : 0);
    v.maxmo_no_ghost_vr = v.nr + 1 - (last ? 2 : 1);
    v.first_active = first ? FARGO_GHOSTCELLS_B : FARGO_CPUOVERLAP;
    v.active_size = v.nr - (last ? FARGO_GHOSTCELLS_B : FARGO_CPUOVERLAP);
    *imax_out = imax;
    return 0;
}

static double omega_kepler_host(const fargo_params &p, double r) { return sqrt(p.G * p.hydro_center_mass / (r * r * r)); }

// init_radialarrays (init.cpp:169-225) for the local slab + per-ring constants the kernels use
static int init_geometry(fargo_ctx *c, const double *radii)
{
    DevView &v = c->v;
    const fargo_params &p = v.p;
    const int gn = p.nrad, nl = v.nr + 2;
    c->h_radii.assign(radii, radii + gn + 1);
    for (int k = 0; k < 4; ++k) // the reference fills a 15-ring search buffer beyond the grid; only entry nr(+1) is read
	c->h_radii.push_back(c->h_radii.back() * (c->h_radii[gn] / c->h_radii[gn - 1]));
    v.dphi = 2.0 * M_PI / (double)v.ns;
    v.invdphi = (double)v.ns / (2.0 * M_PI);
    v.sqrt_gamma = sqrt(p.gamma);
    std::vector<double> rinf(nl), rsup(nl), rmed(nl), surf(nl), invrmed(nl), invsurf(nl), invdiffrsup(nl), invdiffrsuprb(nl),
	twodiffrasq(nl), fourthird(nl), invrinf(nl), invdiffrmed(nl, 0.0), omega_k(nl), inv_omega_k(nl), cs_iso(nl), supp(nl),
	beta_e0(nl), idxt_mid(nl), idxt(nl);
    for (int n = 0; n < nl; ++n) {
	const double ri = c->h_radii[n + v.imin], rs = c->h_radii[n + v.imin + 1];
	rinf[n] = ri;
	rsup[n] = rs;
	double rm = 2.0 / 3.0 * (pow(rs, 3) - pow(ri, 3));
	rm = rm / (pow(rs, 2) - pow(ri, 2));
	rmed[n] = rm;
	surf[n] = M_PI * (pow(rs, 2) - pow(ri, 2)) / (double)v.ns;
	invrmed[n] = 1.0 / rm;
	invsurf[n] = 1.0 / surf[n];
	invdiffrsup[n] = 1.0 / (rs - ri);
	invdiffrsuprb[n] = 1.0 / ((rs - ri) * rm);
	twodiffrasq[n] = 2.0 / (rs * rs - ri * ri);
	fourthird[n] = 4.0 / 3.0 / rm * v.invdphi * v.invdphi;
	invrinf[n] = 1.0 / ri;
	idxt_mid[n] = 2.0 / (v.dphi * (rs + ri));
	{
	    const double dxtheta = v.dphi * rm;
	    idxt[n] = 1.0 / dxtheta;
	}
	omega_k[n] = omega_kepler_host(p, rm);
	inv_omega_k[n] = 1.0 / omega_k[n];
	{ // SourceEuler.cpp:984-991
	    const double vK = sqrt(p.G * p.hydro_center_mass / rm);
	    const double h = p.aspectratio_ref * pow(rm, p.flaring_index);
	    cs_iso[n] = h * vK;
	}
	supp[n] = (p.imposed_disk_drift != 0.0) ? p.imposed_disk_drift * 0.5 * pow(rm, -2.5 + p.sigma_slope) : 0.0;
	// SourceEuler.cpp:668-672: 1/(gamma-1) * h^2 * pow(R, 2 beta - 1) * G * M  (then * sigma per cell)
	beta_e0[n] = 1.0 / (p.gamma - 1.0) * (p.aspectratio_ref * p.aspectratio_ref) * pow(rm, 2.0 * p.flaring_index - 1.0) * p.G *
		     p.hydro_center_mass;
    }
    c->v.limiter_geo_ok = 1;
    for (int n = 1; n < nl; ++n) {
	invdiffrmed[n] = 1.0 / (rmed[n] - rmed[n - 1]);
	if (!(fabs(invdiffrmed[n]) >= 0x1p-30 && fabs(invdiffrmed[n]) <= 0x1p30))
	    c->v.limiter_geo_ok = 0; // fargo_dev.h:limiter_nb
    }
    std::vector<double> cosphi(v.ns), sinphi(v.ns);
    for (int j = 0; j < v.ns; ++j) { // SideEuler.cpp:56-65
	cosphi[j] = cos(v.dphi * (double)j);
	sinphi[j] = sin(v.dphi * (double)j);
    }
    c->h_rinf = rinf;
    c->h_rsup = rsup;
    c->h_rmed = rmed;
    c->h_cs_iso = cs_iso;
    c->h_inv_omega_k = inv_omega_k;
    Geo &g = v.g;
    if (upload_vec(c, &g.rinf, rinf) || upload_vec(c, &g.rsup, rsup) || upload_vec(c, &g.rmed, rmed) ||
	upload_vec(c, &g.surf, surf) || upload_vec(c, &g.invrmed, invrmed) || upload_vec(c, &g.invsurf, invsurf) ||
	upload_vec(c, &g.invdiffrsup, invdiffrsup) || upload_vec(c, &g.invdiffrsuprb, invdiffrsuprb) ||
	upload_vec(c, &g.twodiffrasq, twodiffrasq) || upload_vec(c, &g.fourthird, fourthird) ||
	upload_vec(c, &g.invrinf, invrinf) || upload_vec(c, &g.invdiffrmed, invdiffrmed) || upload_vec(c, &g.cosphi, cosphi) ||
	upload_vec(c, &g.sinphi, sinphi) || upload_vec(c, &g.omega_k, omega_k) || upload_vec(c, &g.inv_omega_k, inv_omega_k) ||
	upload_vec(c, &g.cs_iso, cs_iso) || upload_vec(c, &g.supp_torque, supp) || upload_vec(c, &g.beta_model_e0, beta_e0) ||
	upload_vec(c, &g.invdxtheta_mid, idxt_mid) || upload_vec(c, &g.invdxtheta, idxt))
	return 1;
    return 0;
}

extern "C" const char *fargo_last_error(void) { return g_err.c_str(); }

extern "C" int fargo_get_unique_id(void *out128)
{
    if (load_nccl())
	return 1;
    ncclUniqueId id;
    NCCL_OK(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return 0;
}

extern "C" void fargo_ctx_destroy(fargo_ctx *c)
{
    if (!c)
	return;
    cudaSetDevice(c->device);
    if (c->stream)
	cudaStreamSynchronize(c->stream);
    if (c->comm && g_nccl.ok)
	g_nccl.CommDestroy(c->comm);
    for (double *p : c->dev_allocs)
	cudaFree(p);
    if (c->nshift)
	cudaFree(c->nshift);
    if (c->d_dmask)
	cudaFree(c->d_dmask);
    if (c->h_pin)
	cudaFreeHost(c->h_pin);
    if (c->ev_pin)
	cudaEventDestroy(c->ev_pin);
    for (int k = 0; k < 4; ++k)
	if (c->ev_user[k])
	    cudaEventDestroy(c->ev_user[k]);
    for (int k = 0; k < 2; ++k)
	if (c->halo.peer_base[k])
	    cudaIpcCloseMemHandle(c->halo.peer_base[k]);
    for (int k = 0; k < 4; ++k)
	if (c->snap[k])
	    cudaFree(c->snap[k]);
    if (c->ev_snap_ready)
	cudaEventDestroy(c->ev_snap_ready);
    if (c->ev_snap_done)
	cudaEventDestroy(c->ev_snap_done);
    if (c->stream_snap)
	cudaStreamDestroy(c->stream_snap);
    if (c->halo.inbox)
	cudaFree(c->halo.inbox);
    if (c->halo.done)
	cudaFree(c->halo.done);
    if (c->ev_fork)
	cudaEventDestroy(c->ev_fork);
    if (c->ev_join)
	cudaEventDestroy(c->ev_join);
    if (c->stream2)
	cudaStreamDestroy(c->stream2);
    if (c->stream)
	cudaStreamDestroy(c->stream);
    delete c;
}

// Peer-memory halo path: allocate the inbox, swap CUDA IPC handles with the two neighbours over NCCL, map theirs.
// All ranks must agree (one allreduce); FARGO_B200_HALO=nccl forces the ncclSend / ncclRecv path.
static int halo_setup(fargo_ctx *c)
{
    fargo_ctx::Halo &h = c->halo;
    const DevView &v = c->v;
    h.field_len = (size_t)FARGO_CPUOVERLAP * v.ns;
    const char *env = getenv("FARGO_B200_HALO");
    bool want = !(env && strcmp(env, "nccl") == 0) && v.nr >= 4 * FARGO_CPUOVERLAP;
    const size_t ndbl = 2 * 2 * 4 * h.field_len;
    cudaIpcMemHandle_t mine, theirs[2];
    memset(&mine, 0, sizeof(mine));
    memset(theirs, 0, sizeof(theirs));
    if (want) {
	if (cudaMalloc((void **)&h.inbox, ndbl * sizeof(double) + 2 * sizeof(unsigned long long)) != cudaSuccess ||
	    cudaMemset(h.inbox, 0, ndbl * sizeof(double) + 2 * sizeof(unsigned long long)) != cudaSuccess ||
	    cudaIpcGetMemHandle(&mine, h.inbox) != cudaSuccess) {
	    cudaGetLastError();
	    want = false;
	}
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    // handles travel as 8 doubles each through the scratch field (device memory NCCL can address)
    double *d_mine = c->scratch, *d_prev = c->scratch + 8, *d_next = c->scratch + 16;
    CUDA_OK(cudaMemcpyAsync(d_mine, &mine, 64, cudaMemcpyHostToDevice, c->stream));
    NCCL_OK(g_nccl.GroupStart());
    if (v.rank > 0) {
	NCCL_OK(g_nccl.Send(d_mine, 8, ncclFloat64, v.rank - 1, c->comm, c->stream));
	NCCL_OK(g_nccl.Recv(d_prev, 8, ncclFloat64, v.rank - 1, c->comm, c->stream));
    }
    if (v.rank < v.nranks - 1) {
	NCCL_OK(g_nccl.Send(d_mine, 8, ncclFloat64, v.rank + 1, c->comm, c->stream));
	NCCL_OK(g_nccl.Recv(d_next, 8, ncclFloat64, v.rank + 1, c->comm, c->stream));
    }
    NCCL_OK(g_nccl.GroupEnd());
    CUDA_OK(cudaMemcpyAsync(&theirs[0], d_prev, 64, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaMemcpyAsync(&theirs[1], d_next, 64, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    if (want) {
	for (int k = 0; k < 2 && want; ++k) {
	    const bool have = k == 0 ? v.rank > 0 : v.rank < v.nranks - 1;
	    if (!have)
		continue;
	    cudaIpcMemHandle_t zero;
	    memset(&zero, 0, sizeof(zero));
	    if (memcmp(&theirs[k], &zero, sizeof(zero)) == 0 ||
		cudaIpcOpenMemHandle(&h.peer_base[k], theirs[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
		cudaGetLastError();
		h.peer_base[k] = nullptr;
		want = false;
	    }
	}
    }
    // agreement: min over ranks of "my side is ready"
    c->h_pin[0] = want ? 1.0 : 0.0;
    CUDA_OK(cudaMemcpyAsync(c->d_dt, c->h_pin, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    NCCL_OK(g_nccl.AllReduce(c->d_dt, c->d_dt, 1, ncclFloat64, ncclMin, c->comm, c->stream));
    CUDA_OK(cudaMemcpyAsync(c->h_pin + 1, c->d_dt, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    h.p2p = c->h_pin[1] == 1.0;
    if (h.p2p) {
	h.flags = (unsigned long long *)(h.inbox + ndbl);
	CUDA_OK(cudaMalloc((void **)&h.done, sizeof(unsigned int)));
	CUDA_OK(cudaMemset(h.done, 0, sizeof(unsigned int)));
    }
    return 0;
}
extern "C" int fargo_halo_mode(const fargo_ctx *c) { return c->v.nranks == 1 ? 0 : (c->halo.p2p ? 2 : 1); }

extern "C" int fargo_ctx_create(fargo_ctx **out, const fargo_params *params, const double *radii, int rank, int nranks,
				 const void *nccl_unique_id, int device)
{
    *out = nullptr;
    if (!params || params->abi_version != FARGO_ABI_VERSION)
	return fail("fargo_params.abi_version mismatch (got %d, library %d)", params ? params->abi_version : -1, FARGO_ABI_VERSION);
    if (params->nrad < 5 || params->naz < 1)
	return fail("grid too small: %d x %d", params->nrad, params->naz);
    if (nranks < 1 || rank < 0 || rank >= nranks)
	return fail("bad rank %d / %d", rank, nranks);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
	return fail("no CUDA device available: libfargo_b200 has no CPU fallback");
    if (device < 0 || device >= ndev)
	return fail("device %d out of range (%d devices)", device, ndev);
    CUDA_OK(cudaSetDevice(device));
    fargo_ctx *c = new fargo_ctx();
    memset(&c->v, 0, sizeof(c->v));
    c->device = device;
    c->h_pin = nullptr;
    c->ev_pin = nullptr;
    c->nshift = nullptr;
    c->stream = nullptr;
    c->v.p = *params;
    c->v.ns = params->naz;
    c->v.rank = rank;
    c->v.nranks = nranks;
    int imax;
    if (split_domain(c->v, params->nrad, rank, nranks, &imax, radii, *params)) {
	delete c;
	return 1;
    }
    c->v.b.n = 1;
    c->v.b.mass[0] = params->hydro_center_mass;
#define TRY(expr)                \
    if (expr) {                  \
	fargo_ctx_destroy(c);    \
	return 1;                \
    }
    {
	cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
	if (e != cudaSuccess) {
	    fail("cudaStreamCreate: %s", cudaGetErrorString(e));
	    fargo_ctx_destroy(c);
	    return 1;
	}
    }
    {
	// The side stream carries the transport's ring sums beside the radial sweep.  At default priority its blocks get onto
	// the SMs only as the sweep's CTAs retire (the sweep is enqueued first in practice) and the azimuthal kernel waits
	// for them: on a thin slab that wait is 3 % of the step, so there the side stream gets the highest priority (ring
	// sums done in 0.16 instead of 0.53 ms on 1038 rings).  On a thick slab the two then run one after the other and
	// the step is 0.6 % slower than with the default, so there it keeps the default.
	int prio_lo = 0, prio_hi = 0;
	cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
	// Also measured (profiles/r02_v11_*): small chunks for the side-stream launch (FARGO_B200_RM_SIDE_CHUNK=32: 35 KB of shared
	// memory per warp instead of 100 KB, so that two of its warps fit beside three CTAs of the sweep) at the highest priority.
	// The sums then do finish in the sweep's shadow (0.9 instead of 3.2 ms after the fork), but the sweep takes 0.28 ms longer —
	// the 1 GB the sums read costs the same DRAM time wherever it is placed — and the step is unchanged (18.50 vs 18.48 ms).
	// Not the default.
	const char *sc = getenv("FARGO_B200_RM_SIDE_CHUNK");
	c->rm_side_chunk = sc ? atoi(sc) & ~15 : 0;
	cudaError_t e = cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking,
						     (c->v.nr <= 2560 || c->rm_side_chunk > 0) ? prio_hi : prio_lo);
	if (e == cudaSuccess)
	    e = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
	if (e == cudaSuccess)
	    e = cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
	if (e != cudaSuccess) {
	    fail("side stream: %s", cudaGetErrorString(e));
	    fargo_ctx_destroy(c);
	    return 1;
	}
    }
    { // Ring sums: the chain kernel (a thread per ring) has nr / 32 warps and is bound by the latency of one dependent
      // DADD chain per ring until there are enough rings to fill the GPU's memory pipes; the scan kernel (a warp per ring)
      // costs ~60x the instructions but spreads them.  Measured at Ns = 16384 (ms, chain / scan): 1038 rings 0.166 / 0.100,
      // 2048 rings 0.166 / 0.142, 8192 rings 0.258 / 0.447.  FARGO_B200_RINGSUM=chain|scan overrides.
	const char *fold = getenv("FARGO_B200_FOLD_DAMPING");
	c->fold_damping = !(fold && strcmp(fold, "0") == 0);
	const char *fa = getenv("FARGO_B200_FUSE_ARTVISC");
	c->fuse_artvisc = fa && strcmp(fa, "1") == 0;
	const char *cm = getenv("FARGO_B200_CFL");
	c->cfl_mode = (cm && strcmp(cm, "full") == 0) ? 0 : (cm && strcmp(cm, "check") == 0) ? 2 : 1;
	const char *env = getenv("FARGO_B200_RINGSUM");
	c->ringsum_scan = c->v.nr <= 2560;
	if (env && strcmp(env, "chain") == 0)
	    c->ringsum_scan = false;
	if (env && strcmp(env, "scan") == 0)
	    c->ringsum_scan = true;
    }
    TRY(init_geometry(c, radii));
    const size_t ns = (size_t)c->v.nr * c->v.ns, nv = (size_t)(c->v.nr + 1) * c->v.ns;
    TRY(dalloc(c, &c->sigma, ns) || dalloc(c, &c->eb[0], ns) || dalloc(c, &c->vrb[0], nv) || dalloc(c, &c->vpb[0], ns) ||
	dalloc(c, &c->vrb[1], nv) || dalloc(c, &c->vpb[1], ns) || dalloc(c, &c->eb[1], params->adiabatic ? ns : 1));
    TRY(dalloc(c, &c->sigma0, ns) || dalloc(c, &c->energy0, ns) || dalloc(c, &c->vr0, nv) || dalloc(c, &c->vp0, ns));
    TRY(dalloc(c, &c->qplus, ns) || dalloc(c, &c->qminus, ns));
    if (params->bc_vrad[1] == FARGO_BC_KEPLERIAN || params->bc_vrad[1] == FARGO_BC_VISCOUS ||
	(params->bc_vrad[0] == FARGO_BC_VISCOUS && params->adiabatic && params->viscous_alpha > 0)) {
	// the outer variants address rings past their grids in the reference; the inner viscous outflow reads the VISCOSITY grid as
	// last stored, which the fused kernels do not keep
	fail("v_rad boundary: 'keplerian' / 'viscous' are offered on the inner side only, 'viscous' with a viscosity that does not depend on the state");
	fargo_ctx_destroy(c);
	return 1;
    }
    if (params->cooling_scurve != 0 &&
	(!params->adiabatic || params->heating_star || params->cooling_scurve < 0 || params->cooling_scurve > 2 || !(params->energy_flux_cgs > 0))) {
	// with an irradiating body the reference's irradiation would read the TAU_EFF scurve_cooling stored one call earlier
	fail("SurfaceCooling: scurve needs the energy equation, no irradiating body, ScurveType 1 | 2 and the cgs unit factors");
	fargo_ctx_destroy(c);
	return 1;
    }
    if (!params->body_force_from_potential) { // SourceEuler.cpp:348-353, 406-413: the kicks would read ACCEL_RADIAL / ACCEL_AZIMUTHAL
	fail("BodyForceFromPotential: no (body forces from the acceleration grids) is not implemented");
	fargo_ctx_destroy(c);
	return 1;
    }
    if (params->alpha_mode != 0) { // viscosity::get_alpha (viscosity.cpp:31-49)
	if (params->alpha_mode != 1 || !params->adiabatic || !(params->viscous_alpha > 0)) {
	    fail("AlphaMode %d: only the S-curve (1) with the energy equation and ViscousAlpha > 0 is implemented", params->alpha_mode);
	    fargo_ctx_destroy(c);
	    return 1;
	}
	TRY(dalloc(c, &c->v.t_alpha, ns));
    }
    TRY(dalloc(c, &c->pot, ns) || dalloc(c, &c->qr, ns) || dalloc(c, &c->qphi, ns) || dalloc(c, &c->nu, ns) ||
	dalloc(c, &c->divv, ns) || dalloc(c, &c->trr, ns) || dalloc(c, &c->tpp, ns) || dalloc(c, &c->trp, nv));
    if (params->stabilize_viscosity) {
	TRY(dalloc(c, &c->nusig, ns) || dalloc(c, &c->nusig_rp, nv) || dalloc(c, &c->cf_r, ns) || dalloc(c, &c->cf_phi, ns));
    } else {
	c->nusig = c->nusig_rp = c->cf_r = c->cf_phi = nullptr;
    }
    TRY(dalloc(c, &c->t_sigma, ns) || dalloc(c, &c->t_rmp, ns) || dalloc(c, &c->t_rmm, ns) || dalloc(c, &c->t_amp, ns) ||
	dalloc(c, &c->t_amm, ns) || dalloc(c, &c->t_e, params->adiabatic ? ns : 1));
    TRY(dalloc(c, &c->vmean, c->v.nr + 2) || dalloc(c, &c->vconst, c->v.nr + 2) || dalloc(c, &c->expf_s, 4 * (c->v.nr + 2)) ||
	dalloc(c, &c->expf_v, 1) || dalloc(c, &c->d_dt, 2) || dalloc(c, &c->scratch, ns) || dalloc(c, &c->force4, 4));
    TRY(dalloc(c, &c->mon_rings, (size_t)MD_N * params->nrad));
    TRY(dalloc(c, &c->cfl_bmax, (size_t)((c->v.ns + 511) / 512) * c->v.nr + 1) || dalloc(c, &c->d_cfl_l, 2));
    { // the reductions' partials have their own buffer: Nr x Nphi scratch is too small for them on grids with a few sectors
	const size_t nr_ = (size_t)c->v.nr;
	const size_t n_acc = (size_t)((c->v.ns + ACC_THREADS - 1) / ACC_THREADS) * nr_ * 3;
	const size_t n_mq = ((size_t)((c->v.ns + 4 * MQ_THREADS - 1) / (4 * MQ_THREADS)) * nr_ + 1) * MQ_N;
	const size_t n_dob = (size_t)((c->v.ns + 4 * DOB_THREADS - 1) / (4 * DOB_THREADS)) * nr_ * 4;
	const size_t n_md = (size_t)((c->v.ns + 4 * MQ_THREADS - 1) / (4 * MQ_THREADS)) * nr_ * MD_N; // fargo_monitor_disk
	c->partials_n = std::max(std::max(n_acc, n_md), std::max(n_mq, n_dob));
	TRY(dalloc(c, &c->partials, c->partials_n));
    }
    if (params->leapfrog)
	TRY(dalloc(c, &c->hstale, ns));
    {
	cudaError_t e = cudaMalloc((void **)&c->nshift, (c->v.nr + 2) * sizeof(int));
	if (e == cudaSuccess)
	    e = cudaMemsetAsync(c->nshift, 0, (c->v.nr + 2) * sizeof(int), c->stream);
	if (e == cudaSuccess)
	    e = cudaMallocHost((void **)&c->h_pin, (4 * (c->v.nr + 2) + 8) * sizeof(double));
	if (e == cudaSuccess)
	    e = cudaEventCreateWithFlags(&c->ev_pin, cudaEventDisableTiming);
	if (e != cudaSuccess) {
	    fail("allocation failed: %s", cudaGetErrorString(e));
	    fargo_ctx_destroy(c);
	    return 1;
	}
    }
    // launch geometry of the marching kernels: rings per march chosen so that the CTAs fill the GPU evenly
    {
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
	c->n_sm = sms;
	const int nwin_az = (c->v.ns + AZ_OUT - 1) / AZ_OUT, nwin_fs = (c->v.ns + FS_OUT - 1) / FS_OUT;
	// resident CTAs per SM as the driver reports them for the kernels this configuration launches
	const bool adi = c->v.p.adiabatic != 0, mc = c->v.p.flux_limiter == FARGO_LIMITER_MC;
	auto occ = [](const void *k) {
	    int n = 0;
	    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, 128, 0) != cudaSuccess || n < 1)
		n = 1;
	    return n;
	};
	const int occ_fs = adi ? std::min(occ((const void *)k_fused_sources<true, false, true>),
					  std::min(occ((const void *)k_fused_artvisc<true>), occ((const void *)k_fused_viscosity<true, false>)))
			       : std::min(occ((const void *)k_fused_sources<false, false, true>),
					  std::min(occ((const void *)k_fused_artvisc<false>), occ((const void *)k_fused_viscosity<false, false>)));
	const int occ_az = mc ? (adi ? occ((const void *)k_transport_azimuthal<FARGO_LIMITER_MC, true, false>)
				     : occ((const void *)k_transport_azimuthal<FARGO_LIMITER_MC, false, false>))
			      : (adi ? occ((const void *)k_transport_azimuthal<FARGO_LIMITER_VANLEER, true, false>)
				     : occ((const void *)k_transport_azimuthal<FARGO_LIMITER_VANLEER, false, false>));
	const int occ_rad = mc ? (adi ? occ((const void *)k_transport_radial<FARGO_LIMITER_MC, true>)
				      : occ((const void *)k_transport_radial<FARGO_LIMITER_MC, false>))
			       : (adi ? occ((const void *)k_transport_radial<FARGO_LIMITER_VANLEER, true>)
				      : occ((const void *)k_transport_radial<FARGO_LIMITER_VANLEER, false>));
	c->az_slots = occ_az * sms; // resident CTAs of the azimuthal kernel (launch_transport sizes its ring bands with it)
	c->fs_R = rings_per_march(c->v.nr, (nwin_fs + 3) / 4, occ_fs * sms, 1);
	c->rad_chunk = rings_per_march(c->v.nr, (c->v.ns + 127) / 128, occ_rad * sms, 2);
	// ring means: one warp per 32 rings; give each resident warp as much of the SM's shared memory as its share allows
	if ((c->v.ns & 1) == 0) {
	    const int nblocks = (c->v.nr + 31) / 32;
	    const int per_sm = (nblocks + sms - 1) / sms;
	    const size_t budget = (size_t)200 * 1024 / (per_sm > 4 ? 4 : per_sm);
	    int chunk = (int)(budget / (RM_STAGES * 32 * sizeof(double))) - 2;
	    chunk &= ~15;
	    if (chunk > 256)
		chunk = 256;
	    if (chunk > ((c->v.ns + 15) & ~15))
		chunk = (c->v.ns + 15) & ~15;
	    if (chunk >= 16) {
		c->rm_chunk = chunk;
		c->rm_smem = (size_t)RM_STAGES * 32 * (chunk + 2) * sizeof(double) + RM_STAGES * sizeof(unsigned long long);
		cudaError_t e = cudaFuncSetAttribute(k_ring_mean, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->rm_smem);
		if (e != cudaSuccess) {
		    fail("k_ring_mean shared memory opt-in: %s", cudaGetErrorString(e));
		    fargo_ctx_destroy(c);
		    return 1;
		}
	    }
	}
    }
    if (nranks > 1) {
	if (!nccl_unique_id) {
	    fail("nranks > 1 needs an ncclUniqueId");
	    fargo_ctx_destroy(c);
	    return 1;
	}
	TRY(load_nccl());
	ncclUniqueId id;
	memcpy(&id, nccl_unique_id, sizeof(id));
	int r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
	if (r != 0) {
	    fail("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
	    fargo_ctx_destroy(c);
	    return 1;
	}
	TRY(halo_setup(c));
    }
#undef TRY
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) {
	fail("context initialisation failed: %s", cudaGetErrorString(cudaGetLastError()));
	fargo_ctx_destroy(c);
	return 1;
    }
    *out = c;
    return 0;
}

extern "C" int fargo_local_nrad(const fargo_ctx *c) { return c->v.nr; }
extern "C" int fargo_local_imin(const fargo_ctx *c) { return c->v.imin; }
extern "C" long long fargo_launch_count(const fargo_ctx *c) { return c->launches; }
extern "C" int fargo_sync(fargo_ctx *c)
{
    CUDA_OK(cudaSetDevice(c->device));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

#define LAUNCH(c, kernel, grid, block, smem, ...) LAUNCH_ON(c, (c)->stream, kernel, grid, block, smem, __VA_ARGS__)
#define LAUNCH_ON(c, strm, kernel, grid, block, smem, ...) LAUNCH_NAMED(c, strm, #kernel, kernel, grid, block, smem, __VA_ARGS__)
static inline double host_now_ms()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
#define LAUNCH_NAMED(c, strm, label, kernel, grid, block, smem, ...)     \
    do {                                                                 \
	if ((c)->t_cfl_done != 0.0) {                                    \
	    (c)->host_turnaround_ms += host_now_ms() - (c)->t_cfl_done;  \
	    (c)->host_turnaround_n++;                                    \
	    (c)->t_cfl_done = 0.0;                                       \
	}                                                                \
	if ((c)->profiling)                                              \
	    prof_begin(c, label, strm);                                  \
	kernel<<<grid, block, smem, strm>>>(__VA_ARGS__);                \
	if ((c)->profiling)                                              \
	    prof_end(c, strm);                                           \
	(c)->launches++;                                                 \
	cudaError_t _e = cudaGetLastError();                             \
	if (_e != cudaSuccess)                                           \
	    return fail("launch %s: %s", #kernel, cudaGetErrorString(_e)); \
    } while (0)

static double *state_ptr(fargo_ctx *c, int f, int *rings)
{
    *rings = c->v.nr;
    switch (f) {
    case FARGO_SIGMA: return c->sigma;
    case FARGO_VRAD: *rings = c->v.nr + 1; return c->vrb[c->v_mid ? 1 - c->vcur : c->vcur];
    case FARGO_VAZI: return c->vpb[c->v_mid ? 1 - c->vcur : c->vcur];
    case FARGO_ENERGY: return c->eb[c->ecur];
    case FARGO_SIGMA0: return c->sigma0;
    case FARGO_VRAD0: *rings = c->v.nr + 1; return c->vr0;
    case FARGO_VAZI0: return c->vp0;
    case FARGO_ENERGY0: return c->energy0;
    case FARGO_QPLUS: return c->qplus;
    case FARGO_QMINUS: return c->qminus;
    case FARGO_POTENTIAL: return c->pot;
    case FARGO_GAMMAEFF: return c->v.pv.geff;
    case FARGO_MU: return c->v.pv.mu;
    case FARGO_GAMMA1: return c->v.pv.g1;
    case FARGO_TEMPERATURE: return c->v.t_alpha; // AlphaMode 1 keeps the TEMPERATURE grid (nullptr otherwise: evaluated on download)
    case FARGO_MASSFLOW: *rings = c->v.nr + 1; return c->massflow; // nullptr unless fargo_track_massflow
    case FARGO_SCALE_HEIGHT: return c->v.pv.H; // PVTE keeps the SCALE_HEIGHT grid (nullptr otherwise: evaluated on download)
    }
    return nullptr;
}

struct VBuf { double *vr, *vp; };
static inline VBuf cur_v(fargo_ctx *c, bool mid) { return mid ? VBuf{VRB(c), VPB(c)} : VBuf{VRA(c), VPA(c)}; }

// derived fields are never stored; they are evaluated into `scratch` when somebody asks for them
static int materialize(fargo_ctx *c, int f, double **ptr, int *rings)
{
    *ptr = state_ptr(c, f, rings);
    if (*ptr)
	return 0;
    if (f == FARGO_TEMPERATURE || f == FARGO_PRESSURE || f == FARGO_SOUNDSPEED || f == FARGO_SCALE_HEIGHT || f == FARGO_VISCOSITY) {
	*rings = c->v.nr;
	LAUNCH(c, k_derived_field, cells_grid((long long)c->v.nr * c->v.ns), 256, 0, c->v, c->sigma, EN(c), c->scratch, f);
	*ptr = c->scratch;
	return 0;
    }
    if (f == FARGO_T_REYNOLDS) { // stress::calculate_Reynolds_stress, evaluated when an output asks for it
	*rings = c->v.nr;
	VBuf vb = cur_v(c, c->v_mid);
	LAUNCH(c, k_reynolds_means, (unsigned)((c->v.nr + 63) / 64), 64, 0, c->v, vb.vr, vb.vp, c->vconst, c->vmean);
	LAUNCH(c, k_reynolds_cells, cells_grid((long long)c->v.nr * c->v.ns), 256, 0, c->v, c->sigma, vb.vr, vb.vp, c->vconst, c->vmean,
	       c->scratch);
	*ptr = c->scratch;
	return 0;
    }
    return fail("unknown field id %d", f);
}

extern "C" int fargo_upload_field(fargo_ctx *c, int f, const double *host_global)
{
    CUDA_OK(cudaSetDevice(c->device));
    int rings;
    double *d = state_ptr(c, f, &rings);
    if (!d)
	return fail("field %d cannot be uploaded", f);
    CUDA_OK(cudaMemcpyAsync(d, host_global + (size_t)c->v.imin * c->v.ns, (size_t)rings * c->v.ns * sizeof(double),
			    cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int fargo_download_field(fargo_ctx *c, int f, double *host_global)
{
    CUDA_OK(cudaSetDevice(c->device));
    int rings;
    double *d;
    if (materialize(c, f, &d, &rings))
	return 1;
    // write2D (polargrid.cpp:150-176): rings Zero_or_active .. Max_or_active (+1 for vector grids on the last rank)
    const bool first = c->v.rank == 0, last = c->v.rank == c->v.nranks - 1;
    const int lo = first ? 0 : FARGO_CPUOVERLAP;
    int count = (c->v.nr - (last ? 0 : FARGO_CPUOVERLAP)) - lo;
    if (rings == c->v.nr + 1 && last)
	count += 1;
    CUDA_OK(cudaMemcpyAsync(host_global + (size_t)(c->v.imin + lo) * c->v.ns, d + (size_t)lo * c->v.ns,
			    (size_t)count * c->v.ns * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int fargo_download_slab(fargo_ctx *c, int f, double *host_slab)
{
    CUDA_OK(cudaSetDevice(c->device));
    int rings;
    double *d;
    if (materialize(c, f, &d, &rings))
	return 1;
    CUDA_OK(cudaMemcpyAsync(host_slab, d, (size_t)rings * c->v.ns * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

// Asynchronous snapshot (the output half of sim::handle_outputs / write_full_output, simulation.cpp:50-98, output.cpp:249:
// the reference stops the time loop while it writes).  The four state fields as they are NOW are copied device-to-device
// on the compute stream (1.4 ms at 8192 x 16384) and from there to the caller's host arrays on a copy stream, so the next
// steps run while the snapshot crosses PCIe.  Host arrays: global layout as fargo_download_field (only the owned rings are
// written), page-locked for the copy to be asynchronous; `energy` may be NULL.  fargo_snapshot_wait blocks until the
// data is on the host; a second fargo_snapshot_async waits for the first one's copies by itself.
extern "C" int fargo_snapshot_async(fargo_ctx *c, double *sigma, double *vrad, double *vazi, double *energy)
{
    CUDA_OK(cudaSetDevice(c->device));
    if (c->v_mid)
	return fail("fargo_snapshot_async called mid-step");
    const size_t ns = (size_t)c->v.nr * c->v.ns, nv = (size_t)(c->v.nr + 1) * c->v.ns;
    if (!c->stream_snap) {
	CUDA_OK(cudaStreamCreateWithFlags(&c->stream_snap, cudaStreamNonBlocking));
	CUDA_OK(cudaEventCreateWithFlags(&c->ev_snap_ready, cudaEventDisableTiming));
	CUDA_OK(cudaEventCreateWithFlags(&c->ev_snap_done, cudaEventDisableTiming));
	CUDA_OK(cudaEventRecord(c->ev_snap_done, c->stream_snap));
	const size_t len[3] = {ns, nv, ns};
	for (int k = 0; k < 3; ++k)
	    CUDA_OK(cudaMalloc((void **)&c->snap[k], len[k] * sizeof(double)));
    }
    // the energy grid also travels for an isothermal run whose caller asks for it (the reference writes it out too: it holds
    // the energy ring of CircumBinaryRing, zeros otherwise)
    if (energy && !c->snap[3])
	CUDA_OK(cudaMalloc((void **)&c->snap[3], ns * sizeof(double)));
    double *src[4] = {c->sigma, VRA(c), VPA(c), EN(c)};
    double *host[4] = {sigma, vrad, vazi, energy};
    const bool first = c->v.rank == 0, last = c->v.rank == c->v.nranks - 1;
    const int lo = first ? 0 : FARGO_CPUOVERLAP;
    CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_snap_done, 0)); // the previous snapshot has left the device buffers
    size_t off[4], cnt[4];
    for (int k = 0; k < 4; ++k) {
	int count = (c->v.nr - (last ? 0 : FARGO_CPUOVERLAP)) - lo; // write2D (polargrid.cpp:150-176)
	if (k == 1 && last)
	    count += 1;
	off[k] = (size_t)lo * c->v.ns;
	cnt[k] = (size_t)count * c->v.ns;
	if (!host[k] || !c->snap[k])
	    continue;
	CUDA_OK(cudaMemcpyAsync(c->snap[k] + off[k], src[k] + off[k], cnt[k] * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    }
    CUDA_OK(cudaEventRecord(c->ev_snap_ready, c->stream));
    CUDA_OK(cudaStreamWaitEvent(c->stream_snap, c->ev_snap_ready, 0));
    for (int k = 0; k < 4; ++k) {
	if (!host[k] || !c->snap[k])
	    continue;
	CUDA_OK(cudaMemcpyAsync(host[k] + (size_t)c->v.imin * c->v.ns + off[k], c->snap[k] + off[k], cnt[k] * sizeof(double),
				cudaMemcpyDeviceToHost, c->stream_snap));
    }
    CUDA_OK(cudaEventRecord(c->ev_snap_done, c->stream_snap));
    return 0;
}
// page-locked host memory for callers that do not link the CUDA runtime themselves (host/fargo_host.cpp)
extern "C" void *fargo_pinned_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
	cudaGetLastError();
	fail("fargo_pinned_alloc(%zu) failed", bytes);
	return nullptr;
    }
    return p;
}
extern "C" void fargo_pinned_free(void *p)
{
    if (p)
	cudaFreeHost(p);
}
extern "C" int fargo_snapshot_wait(fargo_ctx *c)
{
    CUDA_OK(cudaSetDevice(c->device));
    if (c->ev_snap_done)
	CUDA_OK(cudaEventSynchronize(c->ev_snap_done));
    return 0;
}

extern "C" int fargo_copy_initial_values(fargo_ctx *c)
{
    CUDA_OK(cudaSetDevice(c->device));
    const size_t ns = (size_t)c->v.nr * c->v.ns * sizeof(double), nv = (size_t)(c->v.nr + 1) * c->v.ns * sizeof(double);
    CUDA_OK(cudaMemcpyAsync(c->vr0, VRA(c), nv, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(c->vp0, VPA(c), ns, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(c->sigma0, c->sigma, ns, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_OK(cudaMemcpyAsync(c->energy0, EN(c), ns, cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}

extern "C" int fargo_set_bodies(fargo_ctx *c, const fargo_bodies *b)
{
    if (b->n < 0 || b->n > FARGO_MAX_BODIES)
	return fail("too many bodies: %d", b->n);
    c->v.b = *b;
    return 0;
}
// init_eos_arrays (init.cpp:1190-1206): lookup tables (built on the host once per process and set of constants), the
// GAMMAEFF / GAMMA1 / MU grids filled with the constant gamma / mu, the SCALE_HEIGHT grid the lookups read
extern "C" int fargo_set_pvte(fargo_ctx *c, const fargo_pvte_consts *k)
{
    CUDA_OK(cudaSetDevice(c->device));
    if (!c->v.p.pvte || !c->v.p.adiabatic)
	return fail("fargo_set_pvte: the context was not created with params.pvte");
    static fargo_pvte_tables *cached = nullptr;
    if (!cached || memcmp(&cached->k, k, sizeof(*k)) != 0) {
	fargo_pvte_free(cached);
	cached = fargo_pvte_build(k);
	if (!cached)
	    return fail("fargo_set_pvte: out of memory building the lookup tables");
    }
    DevView::Pvte &pv = c->v.pv;
    const size_t n = (size_t)c->v.nr * c->v.ns, nt = (size_t)FARGO_PVTE_NI * FARGO_PVTE_NJ;
    if (!pv.geff) {
	if (dalloc(c, &pv.geff, n) || dalloc(c, &pv.mu, n) || dalloc(c, &pv.g1, n) || dalloc(c, &pv.H, n))
	    return 1;
	auto up = [&](const double **dst, const double *src, size_t cnt) {
	    return upload_vec(c, dst, std::vector<double>(src, src + cnt));
	};
	if (up(&pv.t_rho, cached->rho, FARGO_PVTE_NI) || up(&pv.t_e, cached->e, FARGO_PVTE_NJ) || up(&pv.t_mu, cached->mu, nt) ||
	    up(&pv.t_geff, cached->geff, nt) || up(&pv.t_g1, cached->g1, nt))
	    return 1;
	pv.dlogrho = cached->dlogrho, pv.dloge = cached->dloge;
    }
    LAUNCH(c, k_pvte_fill, cells_grid((long long)n), 256, 0, c->v, n);
    return 0;
}

extern "C" int fargo_set_time(fargo_ctx *c, double t)
{
    c->v.time = t;
    return 0;
}

// what fargo_accrete_kley kept of the state before the accretion (fargo_dev.h:PreState)
static PreState pre_state(const fargo_ctx *c)
{
    PreState q;
    q.sigma = c->sig_pre, q.energy = c->e_pre, q.lo = c->pre_lo, q.hi = c->pre_hi;
    return q;
}

// ---------------------------------------------------------------------------------------------
// stages
extern "C" int fargo_stage_potential(fargo_ctx *c)
{
    CUDA_OK(cudaSetDevice(c->device));
    LAUNCH(c, k_potential, cells_grid((long long)c->v.nr * c->v.ns), 256, 0, c->v, c->sigma, EN(c), c->pot,
	   (const double *)(c->h_stale ? c->hstale : nullptr), pre_state(c));
    c->pot_state = 1;
    return 0;
}

// NOTE on buffers (staged path): between steps v lives in buffer A (= vcur).  stage_sources reads A and writes B;
// artvisc / viscosity update B in place; Transport reads B and writes A.  When stages are called one by one (tests) the same
// protocol holds because every per-stage entry point leaves the "current" v where the next stage expects it:
// after sources..substep3 the current v is B, so the boundary stage must know which buffer is current.
extern "C" int fargo_stage_sources(fargo_ctx *c, double dt)
{
    CUDA_OK(cudaSetDevice(c->device));
    if (c->v_mid)
	return fail("stage_sources called while the velocities are mid-step (call stage_transport first)");
    LAUNCH(c, k_sources_velocity, cells_grid((long long)(c->v.nr + 1) * c->v.ns), 256, 0, c->v, c->sigma, EN(c), c->pot,
	   VRA(c), VPA(c), VRB(c), VPB(c), dt, pre_state(c));
    c->pre_lo = c->pre_hi = 0; // the stored PRESSURE has had its last reader (SourceEuler.cpp:325-428)
    c->v_mid = true;
    if (c->v.p.adiabatic)
	LAUNCH(c, k_compression_heating, cells_grid((long long)(c->v.nr - 1) * c->v.ns), 256, 0, c->v, VRB(c), VPB(c), EN(c), dt);
    return 0;
}

// make sure the current velocities are in the B buffers (stages that expect mid-step state, when called
// stand-alone from a between-steps state)
static int ensure_v_mid(fargo_ctx *c)
{
    if (!c->v_mid) {
	CUDA_OK(cudaMemcpyAsync(VRB(c), VRA(c), (size_t)(c->v.nr + 1) * c->v.ns * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
	CUDA_OK(cudaMemcpyAsync(VPB(c), VPA(c), (size_t)c->v.nr * c->v.ns * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
	c->v_mid = true;
    }
    return 0;
}

extern "C" int fargo_stage_artvisc(fargo_ctx *c, double dt)
{
    CUDA_OK(cudaSetDevice(c->device));
    if (ensure_v_mid(c))
	return 1;
    const fargo_params &p = c->v.p;
    const bool diss = p.adiabatic && p.artificial_viscosity_dissipation;
    if (p.artificial_viscosity == FARGO_ARTVISC_NONE && !diss)
	return 0;
    const unsigned g = cells_grid((long long)c->v.nr * c->v.ns);
    LAUNCH(c, k_artvisc_q, g, 256, 0, c->v, c->sigma, VRB(c), VPB(c), EN(c), c->qr, c->qphi, dt);
    if (p.artificial_viscosity != FARGO_ARTVISC_NONE)
	LAUNCH(c, k_artvisc_v, g, 256, 0, c->v, c->sigma, c->qr, c->qphi, VRB(c), VPB(c), dt);
    return 0;
}

static int launch_stress(fargo_ctx *c)
{
    const unsigned g = cells_grid((long long)c->v.nr * c->v.ns);
    const fargo_params &p = c->v.p;
    if (p.viscous_alpha > 0 || !c->visc_const_filled || p.leapfrog) {
	LAUNCH(c, k_viscosity_nu, g, 256, 0, c->v, c->sigma, EN(c), c->nu, (double *)(p.leapfrog ? c->hstale : nullptr));
	c->visc_const_filled = true;
	c->h_stale = p.leapfrog != 0;
    }
    VBuf v = cur_v(c, c->v_mid);
    LAUNCH(c, k_stress, g, 256, 0, c->v, c->sigma, c->nu, v.vr, v.vp, c->divv, c->trr, c->tpp, c->trp, c->nusig, c->nusig_rp);
    if (p.stabilize_viscosity)
	LAUNCH(c, k_stress_correction, g, 256, 0, c->v, c->sigma, c->nusig, c->nusig_rp, c->cf_r, c->cf_phi);
    return 0;
}

extern "C" int fargo_stage_viscosity(fargo_ctx *c, double dt)
{
    CUDA_OK(cudaSetDevice(c->device));
    if (ensure_v_mid(c))
	return 1;
    if (c->v.pv.geff) // recalculate_viscosity (SourceEuler.cpp:214-219): gamma_eff, mu, Gamma_1 from the STORED scale height
	LAUNCH(c, k_pvte_refresh, cells_grid((long long)c->v.nr * c->v.ns), 256, 0, c->v, c->sigma, EN(c), 0);
    if (launch_stress(c))
	return 1;
    LAUNCH(c, k_viscosity_v, cells_grid((long long)c->v.nr * c->v.ns), 256, 0, c->v, c->sigma, c->trr, c->tpp, c->trp, c->cf_r,
	   c->cf_phi, VRB(c), VPB(c), dt);
    return 0;
}

// beta_inv of thermal_relaxation (SourceEuler.cpp:645-653), host side (exp with glibc)
static double beta_inv_host(const fargo_ctx *c)
{
    const fargo_params &p = c->v.p;
    double beta_inv = 1 / p.cooling_beta_value;
    if (p.cooling_beta_ramp_up > 0.0) {
	const double a = 2 * c->v.time / p.cooling_beta_ramp_up;
	const double ramp_factor = 1 - exp(-(a * a));
	beta_inv = beta_inv * ramp_factor;
    }
    return beta_inv;
}

extern "C" int fargo_stage_substep3(fargo_ctx *c, double dt)
{
    CUDA_OK(cudaSetDevice(c->device));
    if (!c->v.p.adiabatic)
	return 0;
    LAUNCH(c, k_substep3, cells_grid((long long)c->v.nr * c->v.ns), 256, 0, c->v, c->sigma, c->nu, c->divv, c->trr, c->tpp,
	   c->trp, c->sigma0, c->energy0, EN(c), c->qplus, c->qminus, dt, beta_inv_host(c), 1);
    return 0;
}

// find_cell_id.cpp:218-291 (host, glibc log/pow like the reference)
static int rmed_id(const fargo_ctx *c, double r)
{
    const fargo_params &p = c->v.p;
    if (p.radial_spacing == FARGO_SPACING_LOG) {
	const double gf = pow(p.rmax / p.rmin, 1.0 / ((double)p.nrad - 2.0));
	const double optimization_const = 3.0 / 2.0 / p.rmin * (1 - pow(gf, 2.0)) / (1 - pow(gf, 3.0));
	const double inv_log_gf = 1.0 / log(gf);
	const double did = log(r * optimization_const) * inv_log_gf;
	return (int)floor(did) - c->v.imin + 1;
    }
    int id = 0;
    while (id < c->v.nr + 1 && c->h_rmed[id] < r)
	id++;
    return id - 1;
}
static int rinf_id(const fargo_ctx *c, double r)
{
    const fargo_params &p = c->v.p;
    if (p.radial_spacing == FARGO_SPACING_LOG) {
	const double gf = pow(p.rmax / p.rmin, 1.0 / ((double)p.nrad - 2.0));
	const double inv_log_gf = 1.0 / log(gf);
	const double did = log(r / p.rmin) * inv_log_gf;
	return (int)floor(did) - c->v.imin + 1;
    }
    int id = 0;
    while (id < c->v.nr + 1 && c->h_rinf[id] < r)
	id++;
    return id - 1;
}
static int clamp_id(const fargo_ctx *c, int id, bool is_vector)
{
    const int mx = c->v.nr - (is_vector ? 0 : 1);
    return id < 0 ? 0 : (id > mx ? mx : id);
}

// damping of one field (damping.cpp:311-752): host computes the ring range and exp factors, device applies
static int damp_field(fargo_ctx *c, double *x, const double *x0, bool is_vector, bool is_density, const int type[2], double dt,
		      double *d_expf, double *h_expf, DampJobs &jobs, int &rows)
{
    const fargo_params &p = c->v.p;
    const int rings = c->v.nr + (is_vector ? 1 : 0);
    const std::vector<double> &radius = is_vector ? c->h_rinf : c->h_rmed;
    const double RMIN = p.rmin, RMAX = p.rmax;
    const double x0_const = is_density ? p.sigma_floor * p.sigma0 : 0.0;
    struct Zone { int lo, hi, type, outer; } zones[2];
    int nz = 0;
    for (int n = 0; n < rings; ++n)
	h_expf[n] = 1.0;
    if (type[0] != FARGO_DAMP_NONE && (p.damping_inner_limit > 1.0) && (radius[0] < RMIN * p.damping_inner_limit)) {
	const int limit = is_vector ? clamp_id(c, rinf_id(c, RMIN * p.damping_inner_limit), true)
				    : clamp_id(c, rmed_id(c, RMIN * p.damping_inner_limit), false);
	const double tau = p.damping_time_factor * 2.0 * M_PI / omega_kepler_host(p, RMIN);
	for (int n = 0; n <= limit; ++n) {
	    const double q = (radius[n] - RMIN * p.damping_inner_limit) / (RMIN - RMIN * p.damping_inner_limit);
	    const double factor = q * q;
	    h_expf[n] = exp(-dt * factor / tau);
	}
	zones[nz++] = {0, limit + 1, type[0], 0};
    }
    if (type[1] != FARGO_DAMP_NONE && (p.damping_outer_limit < 1.0) && (radius[rings - 1] > RMAX * p.damping_outer_limit)) {
	const int limit = is_vector ? clamp_id(c, rinf_id(c, RMAX * p.damping_outer_limit) + 1, true)
				    : clamp_id(c, rmed_id(c, RMAX * p.damping_outer_limit) + 1, false);
	const double tau = p.damping_time_factor * 2.0 * M_PI / omega_kepler_host(p, p.damping_time_radius_outer);
	for (int n = limit; n < rings; ++n) {
	    const double q = (radius[n] - RMAX * p.damping_outer_limit) / (RMAX - RMAX * p.damping_outer_limit);
	    const double factor = q * q;
	    h_expf[n] = exp(-dt * factor / tau);
	}
	zones[nz++] = {limit, rings, type[1], 1};
    }
    for (int z = 0; z < nz; ++z) {
	const int nrings = zones[z].hi - zones[z].lo;
	if (nrings <= 0)
	    continue;
	DampJob &J = jobs.j[jobs.n++];
	J.x = x, J.x0 = x0, J.expf = d_expf;
	J.ring_lo = zones[z].lo, J.ring_hi = zones[z].hi, J.type = zones[z].type, J.row0 = rows, J.outer = zones[z].outer;
	J.x0_const = x0_const;
	J.dmass = (is_density && c->dmass) ? c->scratch : nullptr; // the scratch grid is free between Transport and the CFL
	rows += nrings;
    }
    return 0;
}

// the damping jobs of one step (order: vrad, vazi, sigma, energy, damping.cpp:204-270; the fields are independent) with this
// step's factors exp(-dt * f / tau) uploaded into c->expf_s (4 tables of stride nr + 2)
static int prepare_damping(fargo_ctx *c, const VBuf &v, double dt, DampJobs &jobs, int &rows)
{
    const fargo_params &p = c->v.p;
    const int st = c->v.nr + 2;
    double *h = c->h_pin + 8, *d = c->expf_s;
    CUDA_OK(cudaEventSynchronize(c->ev_pin)); // the previous step's copy out of h_pin is done
    jobs.n = 0;
    rows = 0;
    if (damp_field(c, v.vr, c->vr0, true, false, p.damp_vrad, dt, d, h, jobs, rows) ||
	damp_field(c, v.vp, c->vp0, false, false, p.damp_vazi, dt, d + st, h + st, jobs, rows) ||
	damp_field(c, c->sigma, c->sigma0, false, true, p.damp_sigma, dt, d + 2 * st, h + 2 * st, jobs, rows))
	return 1;
    if (p.adiabatic && damp_field(c, EN(c), c->energy0, false, false, p.damp_energy, dt, d + 3 * st, h + 3 * st, jobs, rows))
	return 1;
    if (jobs.n > 0) {
	CUDA_OK(cudaMemcpyAsync(d, h, 4 * (size_t)st * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	CUDA_OK(cudaEventRecord(c->ev_pin, c->stream));
    }
    return 0;
}

// Arms the fold: this step's damping (over `dt`) will be applied by the next launch_transport in the azimuthal kernel's
// epilogue instead of by k_damping.  Only for zones that damp towards the initial field or a constant: the ring-mean type needs
// the finished ring's index-ordered sum first (damping.cpp:578-585) and keeps its own pass.
static int arm_damping_fold(fargo_ctx *c, double dt)
{
    c->fold_armed = false;
    const fargo_params &p = c->v.p;
    if (!p.damping || !c->fold_damping)
	return 0;
    VBuf v = cur_v(c, c->v_mid);
    DampJobs jobs;
    int rows = 0;
    if (prepare_damping(c, v, dt, jobs, rows))
	return 1;
    if (jobs.n == 0)
	return 0;
    std::vector<int> mask((size_t)c->v.nr + 1, 0);
    for (int q = 0; q < jobs.n; ++q) {
	const DampJob &J = jobs.j[q];
	if (J.type != FARGO_DAMP_INITIAL && J.type != FARGO_DAMP_ZERO)
	    return 0; // ring-mean damping somewhere: everything stays with k_damping
	const int f = J.x == v.vr ? 0 : J.x == v.vp ? 1 : J.x == c->sigma ? 2 : 3;
	if (f == 2 && c->dmass)
	    continue; // MassDelta's wave-damping terms need Sigma before and after: its zones keep their own pass (fargo_stage_boundary)
	for (int i = J.ring_lo; i < J.ring_hi; ++i)
	    mask[i] |= (J.type == FARGO_DAMP_INITIAL ? 1 : 2) << (2 * f);
    }
    if (mask != c->h_dmask) { // the zones do not move: uploaded once
	if (!c->d_dmask)
	    CUDA_OK(cudaMalloc((void **)&c->d_dmask, mask.size() * sizeof(int)));
	CUDA_OK(cudaMemcpyAsync(c->d_dmask, mask.data(), mask.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
	CUDA_OK(cudaStreamSynchronize(c->stream)); // `mask` is pageable memory of this frame
	c->h_dmask = mask;
    }
    c->fold_armed = true;
    return 0;
}

extern "C" int fargo_stage_boundary(fargo_ctx *c, double dt, int final_call)
{
    CUDA_OK(cudaSetDevice(c->device));
    const fargo_params &p = c->v.p;
    VBuf v = cur_v(c, c->v_mid);
    if (final_call && p.damping && c->damp_folded && !c->dmass) {
	c->damp_folded = false; // this step's damping was applied in the azimuthal transport kernel's epilogue
    } else if (final_call && p.damping) {
	DampJobs jobs;
	int rows = 0;
	if (prepare_damping(c, v, dt, jobs, rows))
	    return 1;
	if (c->damp_folded) { // only Sigma's zones were left out of the fold (mass tracking): keep those jobs
	    DampJobs only;
	    only.n = 0;
	    rows = 0;
	    for (int q = 0; q < jobs.n; ++q)
		if (jobs.j[q].x == c->sigma) {
		    only.j[only.n] = jobs.j[q];
		    only.j[only.n].row0 = rows;
		    rows += jobs.j[q].ring_hi - jobs.j[q].ring_lo;
		    only.n++;
		}
	    jobs = only;
	    c->damp_folded = false;
	}
	if (jobs.n > 0) {
	    dim3 grid((unsigned)((c->v.ns + 1023) / 1024), (unsigned)rows);
	    LAUNCH(c, k_damping, grid, 256, 0, c->v, jobs);
	    for (int q = 0; q < jobs.n; ++q) // MassDelta: inner and outer zone apart (damping.cpp:311-357 / 359-420)
		if (jobs.j[q].dmass)
		    LAUNCH(c, k_dmass_accumulate, (unsigned)((c->v.ns + 127) / 128), 128, 0, c->v, (const double *)jobs.j[q].dmass,
			   jobs.j[q].ring_lo, jobs.j[q].ring_hi, c->dmass + (jobs.j[q].outer ? 2 : 0) * (size_t)c->v.ns);
	}
    }
    // keplerian_azimuthal.cpp:29-38, :51-59 (host: sqrt with glibc == IEEE, value is per call)
    const int Irad = c->v.nr - 1;
    double vk_in = p.keplerian_azimuthal_factor[0] * sqrt(p.G * p.hydro_center_mass / c->h_rmed[0]) - c->h_rmed[0] * c->v.b.omega_frame;
    double vk_out =
	p.keplerian_azimuthal_factor[1] * sqrt(p.G * p.hydro_center_mass / c->h_rmed[Irad]) - c->h_rmed[Irad] * c->v.b.omega_frame;
    if (p.bc_vazi[0] == FARGO_BC_BALANCED) { // balanced.cpp:62-65: sqrt(v_sq), then the frame rotation
	vk_in = sqrt(p.balanced_vazi_sq[0]);
	vk_in -= c->h_rmed[0] * c->v.b.omega_frame;
    }
    if (p.bc_vazi[1] == FARGO_BC_BALANCED) {
	vk_out = sqrt(p.balanced_vazi_sq[1]);
	vk_out -= c->h_rmed[Irad] * c->v.b.omega_frame;
    }
    BcVrad bv = {{0.0, 0.0}};
    if (p.bc_vrad[0] == FARGO_BC_KEPLERIAN) { // keplerian_radial.cpp:29-38
	for (int k = 0; k <= 1; ++k)
	    bv.in[k] = p.keplerian_radial_factor[0] * sqrt(p.G * p.hydro_center_mass / c->h_rmed[k]);
    } else if (p.bc_vrad[0] == FARGO_BC_VISCOUS) { // viscous.cpp:29-45, nu of rings 0 and 1 from the geometry alone
	double nu01[2];
	for (int k = 0; k <= 1; ++k) {
	    if (p.viscous_alpha > 0) { // update_viscosity (viscosity.cpp:98-137) of a locally isothermal disk: alpha H c_s
		const double cs = c->h_cs_iso[k];
		nu01[k] = p.viscous_alpha * (cs * c->h_inv_omega_k[k]) * cs;
	    } else {
		nu01[k] = p.constant_viscosity;
	    }
	}
	const double Nu = 0.5 * (nu01[0] + nu01[1]);
	bv.in[1] = -1.5 * p.viscous_outflow_speed / c->h_rinf[1] * Nu;
	bv.in[0] = -1.5 * p.viscous_outflow_speed / c->h_rinf[0] * Nu;
    }
    LAUNCH(c, k_boundary, (unsigned)((c->v.ns + 255) / 256), 256, 0, c->v, c->sigma, EN(c), v.vr, v.vp, c->sigma0, c->energy0,
	   c->vr0, c->vp0, vk_in, vk_out, bv);
    return 0;
}

// ring means of v_azi (+ Nshift and the constant residual for the transport) on stream `strm`
static int launch_ring_mean(fargo_ctx *c, cudaStream_t strm, const double *vp, double dt, int mode)
{
    const DevView &v = c->v;
    const unsigned nb = (unsigned)((v.nr + 31) / 32);
    const char *label = mode == 1 ? "k_ring_mean[transport,side-stream]" : "k_ring_mean[cfl]";
    if (c->ringsum_scan) { // a warp per ring, scans instead of the dependent chain (kernels_ringsum.cuh)
	LAUNCH_NAMED(c, strm, label, k_ring_sum_scan, (unsigned)((v.nr + 3) / 4), 128, 0, v, vp, c->vmean, c->nshift, c->vconst, dt, mode,
		     v.nr, v.ns);
	return 0;
    }
    if (c->rm_chunk > 0 && mode == 1 && c->rm_side_chunk > 0 && c->rm_side_chunk < c->rm_chunk) { // beside the radial sweep
	const size_t smem = (size_t)RM_STAGES * 32 * (c->rm_side_chunk + 2) * sizeof(double) + RM_STAGES * sizeof(unsigned long long);
	LAUNCH_NAMED(c, strm, label, k_ring_mean, nb, 32, smem, v, vp, c->vmean, c->nshift, c->vconst, dt, mode, c->rm_side_chunk);
    } else if (c->rm_chunk > 0)
	LAUNCH_NAMED(c, strm, label, k_ring_mean, nb, 32, c->rm_smem, v, vp, c->vmean, c->nshift, c->vconst, dt, mode, c->rm_chunk);
    else
	LAUNCH_NAMED(c, strm, label, k_ring_mean_generic, nb, 32, 0, v, vp, c->vmean, c->nshift, c->vconst, dt, mode);
    return 0;
}

// Transport: reads (Sigma, e, v) from `in`, writes Sigma and e in place and the new velocities to `out`
// (out != in: the azimuthal kernel still reads the old v_azi of neighbouring columns while it stores).
template <int LIM>
static int launch_transport(fargo_ctx *c, double dt, const double *vr_in, const double *vp_in, double *vr_out, double *vp_out)
{
    const DevView &v = c->v;
    // ring means of the pre-transport v_azi, Nshift, constant residual: latency-bound (one dependent add chain per
    // ring) and not needed by the radial sweep, so they run beside it on the side stream
    CUDA_OK(cudaEventRecord(c->ev_fork, c->stream));
    CUDA_OK(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
    if (launch_ring_mean(c, c->stream2, vp_in, dt, 1))
	return 1;
    CUDA_OK(cudaEventRecord(c->ev_join, c->stream2));
    {
	dim3 grid((unsigned)((v.ns + 127) / 128), (unsigned)((v.nr + c->rad_chunk - 1) / c->rad_chunk));
	if (c->track_massflow || c->bflow) { // WriteMassFlow / the boundary mass flows of Quantities.dat: the same sweep also accumulates them
#define RADIAL_MF(ADI_, MODE_, label)                                                                                                       \
    LAUNCH_NAMED(c, c->stream, label, (k_transport_radial<LIM, ADI_, MODE_>), grid, 128, 0, v, c->sigma, vr_in, vp_in, EN(c), c->t_sigma,     \
		 c->t_rmp, c->t_rmm, c->t_amp, c->t_amm, c->t_e, dt, c->rad_chunk, c->massflow, c->bflow)
	    if (c->track_massflow && !c->bflow && fargo_track_boundary_flow(c, 1)) // MF == 2 writes both
		return 1;
	    if (c->track_massflow) {
		if (v.p.adiabatic)
		    RADIAL_MF(true, 2, "(k_transport_radial<LIM, true>)[+massflow]");
		else
		    RADIAL_MF(false, 2, "(k_transport_radial<LIM, false>)[+massflow]");
	    } else {
		if (v.p.adiabatic)
		    RADIAL_MF(true, 1, "(k_transport_radial<LIM, true>)[+boundary flow]");
		else
		    RADIAL_MF(false, 1, "(k_transport_radial<LIM, false>)[+boundary flow]");
	    }
#undef RADIAL_MF
	} else if (v.p.adiabatic)
	    LAUNCH(c, (k_transport_radial<LIM, true>), grid, 128, 0, v, c->sigma, vr_in, vp_in, EN(c), c->t_sigma, c->t_rmp, c->t_rmm,
		   c->t_amp, c->t_amm, c->t_e, dt, c->rad_chunk);
	else
	    LAUNCH(c, (k_transport_radial<LIM, false>), grid, 128, 0, v, c->sigma, vr_in, vp_in, EN(c), c->t_sigma, c->t_rmp, c->t_rmm,
		   c->t_amp, c->t_amm, c->t_e, dt, c->rad_chunk);
    }
    const int nwin = (v.ns + AZ_OUT - 1) / AZ_OUT;
    const int gx = (nwin + 3) / 4;
    fargo_ctx::Halo &h = c->halo;
    AzSegs segs;
    memset(&segs, 0, sizeof(segs));
    segs.nwin = nwin;
    // segment 2 in bands of rings sized so that its (groups x bands) CTAs fill the GPU in whole waves
    int nbands = 0;
    auto cut_bands = [&](void) {
	const int nb = segs.hi[2] - segs.lo[2];
	segs.band = nb > 0 ? rings_per_march(nb, gx, c->az_slots, 1) : 1;
	nbands = nb > 0 ? (nb + segs.band - 1) / segs.band : 0;
    };
#define AZ_LAUNCH(strm, label, PUSH)                                                                                        \
    do {                                                                                                                    \
	const unsigned nblk = (unsigned)((segs.n_edge + nbands) * gx);                                        \
	if (v.p.adiabatic)                                                                                                  \
	    LAUNCH_NAMED(c, strm, label "<ADI>", (k_transport_azimuthal<LIM, true, PUSH>), nblk, 128, 0, v, c->t_sigma, c->t_rmp, c->t_rmm, \
			 c->t_amp, c->t_amm, c->t_e, vp_in, vr_in, c->vmean, c->nshift, c->vconst, c->sigma, vr_out, vp_out, EN(c), \
			 dt, segs);                                                                                         \
	else                                                                                                                \
	    LAUNCH_NAMED(c, strm, label "<ISO>", (k_transport_azimuthal<LIM, false, PUSH>), nblk, 128, 0, v, c->t_sigma, c->t_rmp, c->t_rmm, \
			 c->t_amp, c->t_amm, c->t_e, vp_in, vr_in, c->vmean, c->nshift, c->vconst, c->sigma, vr_out, vp_out, EN(c), \
			 dt, segs);                                                                                         \
    } while (0)
    if (c->fold_armed) { // fargo_step armed the fold: the factors of this step are in expf_s (stream order)
	segs.dmask = c->d_dmask;
	segs.dexpf = c->expf_s;
	segs.dstride = v.nr + 2;
	segs.dx0[0] = c->vr0, segs.dx0[1] = c->vp0, segs.dx0[2] = c->sigma0, segs.dx0[3] = c->energy0;
	segs.dx0c[0] = segs.dx0c[1] = segs.dx0c[3] = 0.0;
	segs.dx0c[2] = v.p.sigma_floor * v.p.sigma0; // damping.cpp: the density is damped towards the floor, not towards 0
	c->fold_armed = false;
	c->damp_folded = true;
    }
    CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    if (!h.p2p) {
	segs.hi[2] = v.nr;
	cut_bands();
	AZ_LAUNCH(c->stream, "k_transport_azimuthal", false);
	return 0;
    }
    // Peer-memory halo path: the 2 x CPUOVERLAP rings at either interior edge of the slab are the first CTAs of the
    // launch; their epilogue stores the rings the neighbours need straight into their inboxes (NVLink peer stores) and
    // the last edge warp bumps the neighbours' arrival counters.  The interior pieces follow in the same launch, so the
    // exchange costs no time of its own (fargo_stage_halo only waits for the counters and unpacks).
    {
	const int E = 2 * FARGO_CPUOVERLAP;
	const bool has_prev = v.rank > 0, has_next = v.rank < v.nranks - 1;
	const unsigned long long parity = h.seq & 1ull;
	const size_t ndbl = 2 * 2 * 4 * h.field_len;
	if (has_prev) { // rings [7, 14) -> prev's inbox, side "from next"
	    segs.edge_seg[segs.n_edge++] = 0;
	    segs.lo[0] = 0, segs.hi[0] = E, segs.push_lo[0] = FARGO_CPUOVERLAP;
	    for (int f = 0; f < 4; ++f)
		segs.push[0][f] = (double *)h.peer_base[0] + ((parity * 2 + 1) * 4 + f) * h.field_len;
	    segs.peer_flag[0] = (unsigned long long *)((double *)h.peer_base[0] + ndbl) + 1;
	}
	if (has_next) { // rings [nr-14, nr-7) -> next's inbox, side "from prev"
	    segs.edge_seg[segs.n_edge++] = 1;
	    segs.lo[1] = v.nr - E, segs.hi[1] = v.nr, segs.push_lo[1] = v.nr - E;
	    for (int f = 0; f < 4; ++f)
		segs.push[1][f] = (double *)h.peer_base[1] + ((parity * 2 + 0) * 4 + f) * h.field_len;
	    segs.peer_flag[1] = (unsigned long long *)((double *)h.peer_base[1] + ndbl) + 0;
	}
	segs.lo[2] = has_prev ? E : 0;
	segs.hi[2] = has_next ? v.nr - E : v.nr;
	cut_bands();
	segs.seq = h.seq + 1;
	segs.done = h.done;
	segs.expected = (unsigned)(nwin * segs.n_edge);
	AZ_LAUNCH(c->stream, "k_transport_azimuthal[+halo push]", true);
	h.pushed = true;
    }
#undef AZ_LAUNCH
    return 0;
}

extern "C" int fargo_stage_transport(fargo_ctx *c, double dt)
{
    CUDA_OK(cudaSetDevice(c->device));
    if (ensure_v_mid(c))
	return 1;
    int rc = c->v.p.flux_limiter == FARGO_LIMITER_MC ? launch_transport<FARGO_LIMITER_MC>(c, dt, VRB(c), VPB(c), VRA(c), VPA(c))
						      : launch_transport<FARGO_LIMITER_VANLEER>(c, dt, VRB(c), VPB(c), VRA(c), VPA(c));
    if (rc)
	return rc;
    c->v_mid = false;
    return 0;
}

// CommunicateBoundaries (commbound.cpp:98-182): rings are contiguous, so the CPUOVERLAP rings of each field go
// straight from / to field memory — no pack / unpack.
extern "C" int fargo_stage_halo(fargo_ctx *c)
{
    CUDA_OK(cudaSetDevice(c->device));
    if (c->v.nranks == 1)
	return 0;
    if (c->v_mid)
	return fail("stage_halo called mid-step");
    const size_t l = (size_t)FARGO_CPUOVERLAP * c->v.ns;
    const size_t oo = (size_t)(c->v.nr - FARGO_CPUOVERLAP) * c->v.ns;
    if (c->halo.pushed) { // peer path: the transport kernel has already stored our edge rings into the neighbours' inboxes
	fargo_ctx::Halo &h = c->halo;
	const bool has_prev = c->v.rank > 0, has_next = c->v.rank < c->v.nranks - 1;
	const unsigned long long parity = h.seq & 1ull;
	HaloUnpack u;
	memset(&u, 0, sizeof(u));
	double *fld[4] = {c->sigma, VRA(c), VPA(c), EN(c)};
	const int nfld = c->v.p.adiabatic ? 4 : 3;
	for (int side = 0; side < 2; ++side) {
	    if (side == 0 ? !has_prev : !has_next)
		continue;
	    for (int f = 0; f < nfld; ++f) {
		u.src[u.n] = h.inbox + ((parity * 2 + side) * 4 + f) * h.field_len;
		u.dst[u.n] = fld[f] + (side == 0 ? 0 : oo);
		++u.n;
	    }
	}
	u.len = l;
	u.flag[0] = has_prev ? h.flags + 0 : nullptr;
	u.flag[1] = has_next ? h.flags + 1 : nullptr;
	u.want = h.seq + 1;
	u.timed_out = c->d_dt + 1;
	dim3 grid((unsigned)((l + 4 * 256 - 1) / (4 * 256)), (unsigned)u.n);
	LAUNCH(c, k_halo_unpack, grid, 256, 0, u);
	h.seq++;
	h.pushed = false;
	return 0;
    }
    const size_t o = (size_t)(c->v.nr - 2 * FARGO_CPUOVERLAP) * c->v.ns;
    double *fields[4] = {c->sigma, VRA(c), VPA(c), EN(c)};
    const int nf = c->v.p.adiabatic ? 4 : 3;
    const int prev = c->v.rank - 1, next = c->v.rank + 1;
    NCCL_OK(g_nccl.GroupStart());
    for (int f = 0; f < nf; ++f) {
	if (c->v.rank > 0) {
	    NCCL_OK(g_nccl.Send(fields[f] + l, l, ncclFloat64, prev, c->comm, c->stream));
	    NCCL_OK(g_nccl.Recv(fields[f], l, ncclFloat64, prev, c->comm, c->stream));
	}
	if (c->v.rank < c->v.nranks - 1) {
	    NCCL_OK(g_nccl.Send(fields[f] + o, l, ncclFloat64, next, c->comm, c->stream));
	    NCCL_OK(g_nccl.Recv(fields[f] + oo, l, ncclFloat64, next, c->comm, c->stream));
	}
    }
    NCCL_OK(g_nccl.GroupEnd());
    c->launches++;
    return 0;
}

// recalculate_derived_disk_quantities (SourceEuler.cpp:225-249): T, c_s, H, P, nu are functions of (Sigma, e)
// that every consumer here re-evaluates in registers, so there is nothing to store.
// AlphaMode 1: the TEMPERATURE grid get_alpha will read during the next recalculate_viscosity (compute_temperature of
// recalculate_derived_disk_quantities / init_euler)
static int store_alpha_temperature(fargo_ctx *c)
{
    if (c->v.t_alpha)
	LAUNCH(c, k_derived_field, cells_grid((long long)c->v.nr * c->v.ns), 256, 0, c->v, c->sigma, EN(c), c->v.t_alpha, (int)FARGO_TEMPERATURE);
    return 0;
}

extern "C" int fargo_stage_derived(fargo_ctx *c)
{
    if (c->v.pv.geff) { // PVTE: scale height after Transport (step_Euler only: simulation.cpp:256-262), then the lookup and the new
			// scale height; step_LeapFrog goes straight to recalculate_derived_disk_quantities (:455), whose lookup
			// reads the scale height its second kick stored
	CUDA_OK(cudaSetDevice(c->device));
	LAUNCH(c, k_pvte_refresh, cells_grid((long long)c->v.nr * c->v.ns), 256, 0, c->v, c->sigma, EN(c), c->v.p.leapfrog ? 0 : 1);
    }
    if (c->v.t_alpha) {
	CUDA_OK(cudaSetDevice(c->device));
	if (store_alpha_temperature(c))
	    return 1;
    }
    c->h_stale = false; // the scale height is the current state's again (only a leapfrog mid-step keeps an older one)
    return 0;
}

// init_euler's compute_heating_cooling_for_CFL (SourceEuler.cpp:1410-1450)
extern "C" int fargo_init_derived(fargo_ctx *c)
{
    CUDA_OK(cudaSetDevice(c->device));
    if (!c->v.p.adiabatic)
	return 0;
    if (c->v.p.pvte) {
	if (!c->v.pv.geff)
	    return fail("EquationOfState: PVTE needs fargo_set_pvte before fargo_init_derived");
	// init_euler (SourceEuler.cpp:272-276): c_s and H from the constant gamma, the first lookup, c_s and H again
	LAUNCH(c, k_pvte_refresh, cells_grid((long long)c->v.nr * c->v.ns), 256, 0, c->v, c->sigma, EN(c), 1);
    }
    if (store_alpha_temperature(c))
	return 1;
    if (launch_stress(c))
	return 1;
    LAUNCH(c, k_substep3, cells_grid((long long)c->v.nr * c->v.ns), 256, 0, c->v, c->sigma, c->nu, c->divv, c->trr, c->tpp,
	   c->trp, c->sigma0, c->energy0, EN(c), c->qplus, c->qminus, 0.0, beta_inv_host(c), 0);
    c->h_stale = false;
    return 0;
}

extern "C" int fargo_condition_cfl(fargo_ctx *c, double *out)
{
    CUDA_OK(cudaSetDevice(c->device));
    const DevView &v = c->v;
    if (c->v_mid)
	return fail("condition_cfl called mid-step");
    c->h_pin[0] = 1.7976931348623157e308;
    CUDA_OK(cudaMemcpyAsync(c->d_dt, c->h_pin, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (launch_ring_mean(c, c->stream, VPA(c), 0.0, 0))
	return 1;
    const int nact = v.active_size - v.first_active;
    if (nact > 0 && (v.pv.geff || v.p.alpha_mode != 0)) { // PVTE: per-cell gamma_eff / Gamma_1; AlphaMode: per-cell alpha
	dim3 grid((unsigned)((v.ns + 127) / 128), (unsigned)nact);
	LAUNCH_NAMED(c, c->stream, "k_cfl", k_cfl_cells, grid, 128, 0, v, c->sigma, EN(c), VRA(c), VPA(c), c->qplus, c->qminus, c->cf_r,
		     c->cf_phi, c->vmean, c->d_dt);
    } else if (nact > 0) {
	dim3 grid((unsigned)((v.ns + 511) / 512), (unsigned)nact);
	const bool screen = c->cfl_mode != 0 && v.p.stabilize_viscosity != 2;
	if (screen) { // bound A of every cell cheaply, evaluate the criterion exactly only where the maximum can be
	    const int nrec = (int)(grid.x * grid.y);
	    CUDA_OK(cudaMemsetAsync(c->d_cfl_l, 0, 2 * sizeof(double), c->stream));
	    double *dt_dst = c->d_dt;
	    if (c->cfl_mode == 2) { // check: the screened result goes to d_cfl_l[1], k_cfl's to d_dt
		CUDA_OK(cudaMemcpyAsync(c->d_cfl_l + 1, c->h_pin, sizeof(double), cudaMemcpyHostToDevice, c->stream));
		dt_dst = c->d_cfl_l + 1;
	    }
	    LAUNCH_NAMED(c, c->stream, "k_cfl", k_cfl_screen, grid, 128, 0, v, c->sigma, EN(c), VRA(c), VPA(c), c->qplus, c->qminus, c->vmean,
			 dt_dst, c->cfl_bmax, c->d_cfl_l);
	    LAUNCH_NAMED(c, c->stream, "k_cfl_candidates", k_cfl_candidates, (unsigned)((nrec + 127) / 128), 128, 0, v, c->sigma, EN(c), VRA(c),
			 VPA(c), c->qplus, c->qminus, c->vmean, c->cfl_bmax, c->d_cfl_l, nrec, (int)grid.x, dt_dst);
	}
	if (!screen || c->cfl_mode == 2)
	    LAUNCH_NAMED(c, c->stream, screen ? "k_cfl[check]" : "k_cfl", k_cfl, grid, 128, 0, v, c->sigma, EN(c), VRA(c), VPA(c), c->qplus,
			 c->qminus, c->cf_r, c->cf_phi, c->vmean, c->d_dt);
	if (screen && c->cfl_mode == 2) {
	    double both[2];
	    CUDA_OK(cudaMemcpyAsync(&both[0], c->d_dt, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	    CUDA_OK(cudaMemcpyAsync(&both[1], c->d_cfl_l + 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	    CUDA_OK(cudaStreamSynchronize(c->stream));
	    if (memcmp(&both[0], &both[1], sizeof(double)) != 0)
		return fail("FARGO_B200_CFL=check: screened reduction %.17g != full reduction %.17g", both[1], both[0]);
	}
    }
    if (v.nranks > 1) { // MPI_Allreduce(MIN), cfl.cpp:379
	NCCL_OK(g_nccl.AllReduce(c->d_dt, c->d_dt, 1, ncclFloat64, ncclMin, c->comm, c->stream));
	c->launches++;
    }
    CUDA_OK(cudaMemcpyAsync(c->h_pin + 1, c->d_dt, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    *out = c->h_pin[1];
    if (c->h_pin[2] != 0.0) // k_halo_unpack gave up waiting (kernels_ring.cuh)
	return fail("ghost-ring exchange timed out: a neighbouring rank did not deliver its rings");
    if (c->profiling)
	c->t_cfl_done = host_now_ms();
    return 0;
}

// sim::CalculateTimeStep (simulation.cpp:100-118)
extern "C" int fargo_cfl(fargo_ctx *c, double *last_dt, double *dt_out)
{
    double cfl_dt;
    if (fargo_condition_cfl(c, &cfl_dt))
	return 1;
    const double lim = c->v.p.cfl_max_var * *last_dt;
    const double rv = (cfl_dt < lim) ? cfl_dt : lim; // std::min(CFL_max_var * last_dt, cfl_dt)
    *last_dt = rv;
    *dt_out = rv;
    return 0;
}

// The fused source-term kernels (kernels_fused.cuh): potential + sources + compression | artificial viscosity |
// viscosity + SubStep3.  Every kernel reads the current (e, v) buffers and writes the other ones.
template <bool ADI> static int launch_fused_sources(fargo_ctx *c, double dt)
{
    const DevView &v = c->v;
    const fargo_params &p = v.p;
    const int nwin = (v.ns + FS_OUT - 1) / FS_OUT;
    dim3 grid((unsigned)((nwin + 3) / 4), (unsigned)((v.nr + c->fs_R - 1) / c->fs_R));
    const int eo = ADI ? 1 - c->ecur : c->ecur;
    const bool diss = ADI && p.artificial_viscosity_dissipation;
    const bool artvisc = p.artificial_viscosity != FARGO_ARTVISC_NONE || diss;
    const bool av_in_sources = artvisc && c->fuse_artvisc; // the artificial-viscosity stage inside the source-term kernel
    const bool pre = c->pre_hi > c->pre_lo; // the step began with an accretion call: P and H of the touched rings from the state before it
#define FS_SRC_LAUNCH(PRE_, AV_, label)                                                                                                  \
    LAUNCH_NAMED(c, c->stream, label, (k_fused_sources<ADI, PRE_, AV_>), grid, 128, 0, v, c->sigma, EN(c), VRA(c), VPA(c),                  \
		 (const double *)(c->h_stale ? c->hstale : nullptr), VRB(c), VPB(c), c->eb[eo], dt, c->fs_R, pre_state(c))
    if (av_in_sources) {
	if (pre)
	    FS_SRC_LAUNCH(true, true, "(k_fused_sources+artvisc<ADI, true>)");
	else
	    FS_SRC_LAUNCH(false, true, "(k_fused_sources+artvisc<ADI, false>)");
    } else {
	if (pre)
	    FS_SRC_LAUNCH(true, false, "(k_fused_sources<ADI, true>)");
	else
	    FS_SRC_LAUNCH(false, false, "(k_fused_sources<ADI, false>)");
    }
#undef FS_SRC_LAUNCH
    if (pre)
	c->pre_lo = c->pre_hi = 0;
    c->vcur = 1 - c->vcur;
    c->ecur = eo;
    if (artvisc && !av_in_sources) {
	const int eo2 = ADI ? 1 - c->ecur : c->ecur;
	LAUNCH(c, k_fused_artvisc<ADI>, grid, 128, 0, v, c->sigma, EN(c), VRA(c), VPA(c), VRB(c), VPB(c), c->eb[eo2], dt, c->fs_R);
	c->vcur = 1 - c->vcur;
	c->ecur = eo2;
    }
    {
	const int eo3 = ADI ? 1 - c->ecur : c->ecur;
	if (ADI && (p.cooling_surface || p.heating_star)) // with thermal_cooling / irradiation (kernels_rad.cuh)
	    LAUNCH_NAMED(c, c->stream, "k_fused_viscosity<ADI>", (k_fused_viscosity<ADI, ADI>), grid, 128, 0, v, c->sigma, EN(c), VRA(c), VPA(c),
			 c->sigma0, c->energy0, VRB(c), VPB(c), c->eb[eo3], c->qplus, c->qminus,
			 (double *)(p.leapfrog ? c->hstale : nullptr), dt, ADI ? beta_inv_host(c) : 0.0, c->fs_R);
	else
	    LAUNCH_NAMED(c, c->stream, "k_fused_viscosity<ADI>", (k_fused_viscosity<ADI, false>), grid, 128, 0, v,
			 c->sigma, EN(c), VRA(c), VPA(c), c->sigma0, c->energy0, VRB(c), VPB(c), c->eb[eo3], c->qplus, c->qminus,
			 (double *)(p.leapfrog ? c->hstale : nullptr), dt, ADI ? beta_inv_host(c) : 0.0, c->fs_R);
	c->h_stale = p.leapfrog != 0;
	c->vcur = 1 - c->vcur;
	c->ecur = eo3;
    }
    return 0;
}

// The gas update in three pieces, so that a host can arrange them like step_Euler (kick dt, drift dt) or like
// step_LeapFrog (kick dt/2, drift dt, bodies moved to mid-step, kick dt/2): simulation.cpp:148-267 / 276-459.
//
// fargo_kick: CalculateNbodyPotential, update_with_sourceterms, artificial viscosity, recalculate_viscosity,
// viscous stress + velocity update, SubStep3 (simulation.cpp:167-175, 187-203) over a time step `dt`
extern "C" int fargo_kick(fargo_ctx *c, double dt)
{
    CUDA_OK(cudaSetDevice(c->device));
    const fargo_params &p = c->v.p;
    if (p.pvte && !c->v.pv.geff)
	return fail("EquationOfState: PVTE needs fargo_set_pvte before the first step");
    // PVTE, AlphaMode and the S-curve cooling: per-cell gamma / alpha / the cgs fit live in the staged kernels
    const bool fused = p.stabilize_viscosity == 0 && !c->force_staged && !p.pvte && p.alpha_mode == 0 && p.cooling_scurve == 0;
    if (fused) {
	if (c->v_mid)
	    return fail("fargo_kick (fused) called mid-step after a per-stage call; finish with fargo_stage_transport first");
	if (c->keep_pot) { // before the fused kernels consume the pre-accretion state (fargo_dev.h:PreState)
	    if (fargo_stage_potential(c))
		return 1;
	} else {
	    c->pot_state = 2; // whatever the grid holds, it is not this kick's
	}
	return p.adiabatic ? launch_fused_sources<true>(c, dt) : launch_fused_sources<false>(c, dt);
    }
    if (c->v_mid)
	return fail("fargo_kick (staged) called mid-step");
    const bool second_kick = c->h_stale; // a leapfrog step's second kick: the first one left its scale height behind
    if (fargo_stage_potential(c))
	return 1;
    if (second_kick && c->v.pv.geff) // PVTE (simulation.cpp:368-376): after the potential, which still read the stored scale height —
	LAUNCH(c, k_pvte_refresh, cells_grid((long long)c->v.nr * c->v.ns), 256, 0, c->v, c->sigma, EN(c), 1); // c_s, H, lookup, c_s, H
    if (fargo_stage_sources(c, dt) || fargo_stage_artvisc(c, dt) || fargo_stage_viscosity(c, dt))
	return 1;
    if (p.adiabatic && fargo_stage_substep3(c, dt))
	return 1;
    return 0;
}

// fargo_drift: apply_boundary_condition(final = false) + Transport over `dt` (simulation.cpp:213-215)
extern "C" int fargo_drift(fargo_ctx *c, double dt)
{
    CUDA_OK(cudaSetDevice(c->device));
    const fargo_params &p = c->v.p;
    if (fargo_stage_boundary(c, 0.0, 0))
	return 1;
    if (c->v_mid) // staged kernels left the velocities in the B buffers
	return fargo_stage_transport(c, dt);
    // Transport out of the current velocity buffer into the other one
    int rc = p.flux_limiter == FARGO_LIMITER_MC ? launch_transport<FARGO_LIMITER_MC>(c, dt, VRA(c), VPA(c), VRB(c), VPB(c))
						 : launch_transport<FARGO_LIMITER_VANLEER>(c, dt, VRA(c), VPA(c), VRB(c), VPB(c));
    if (rc)
	return rc;
    c->vcur = 1 - c->vcur;
    return 0;
}

// fargo_finish_step: CommunicateBoundaries, apply_boundary_condition(final = true) with the damping zones over the
// whole step `dt`, derived quantities (simulation.cpp:230-266)
extern "C" int fargo_finish_step(fargo_ctx *c, double dt)
{
    CUDA_OK(cudaSetDevice(c->device));
    if (c->v_mid) { // staged second kick of a leapfrog step: the mid-step buffers hold the final velocities
	c->vcur = 1 - c->vcur;
	c->v_mid = false;
    }
    if (fargo_stage_halo(c) || fargo_stage_boundary(c, dt, 1) || fargo_stage_derived(c))
	return 1;
    c->h_stale = false; // recalculate_derived_disk_quantities: H is the current state's again
    c->pre_lo = c->pre_hi = 0; // ... and so is the pressure (an accretion call after the last kick of a leapfrog step)
    return 0;
}

// gas part of step_Euler (simulation.cpp:167-175, 187-218, 230-266)
extern "C" int fargo_step(fargo_ctx *c, double dt)
{
    CUDA_OK(cudaSetDevice(c->device));
    if (c->v_mid)
	return fail("fargo_step called mid-step (after a per-stage call); finish the step with fargo_stage_transport first");
    if (fargo_kick(c, dt))
	return 1;
    // an Euler step: Transport is the last stage before the final boundary call, so the damping zones of that call
    // (boundary_conditions.cpp:65-114) are applied to the rings while the transport kernel still holds them
    if (arm_damping_fold(c, dt) || fargo_drift(c, dt))
	return 1;
    c->v.time += dt;
    return fargo_finish_step(c, dt);
}

// correct_v_azimuthal (SideEuler.cpp:79-95) for refframe::handle_corotation
extern "C" int fargo_correct_vazi(fargo_ctx *c, double domega)
{
    CUDA_OK(cudaSetDevice(c->device));
    if (c->v_mid)
	return fail("fargo_correct_vazi called mid-step");
    LAUNCH(c, k_correct_vazi, cells_grid((long long)c->v.nr * c->v.ns), 256, 0, c->v, VPA(c), domega);
    return 0;
}

// test hook: run fargo_step through the staged kernels (one per reference loop nest) instead of the fused ones
extern "C" int fargo_set_staged(fargo_ctx *c, int on)
{
    c->force_staged = on != 0;
    return 0;
}

extern "C" int fargo_get_nshift(fargo_ctx *c, int *out)
{
    CUDA_OK(cudaSetDevice(c->device));
    CUDA_OK(cudaMemcpyAsync(out, c->nshift, (size_t)c->v.nr * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

// accretion::AccreteOntoSinglePlanet (accretion.cpp:84-221) and SinkHoleSinglePlanet (:223-333): gas within frac1 * r_hill
// of the body loses the fraction facc1, within frac2 * r_hill another facc2 (frac2 = 0: no second zone)
static int accrete_zones(fargo_ctx *c, double x, double y, double r_hill, double facc1, double facc2, double frac1, double frac2,
			 double out3[3], bool viscous = false)
{
    CUDA_OK(cudaSetDevice(c->device));
    if (c->v_mid)
	return fail("accretion called mid-step");
    const DevView &v = c->v;
    const double frac = frac1;
    AccretionIn a;
    a.x = x, a.y = y, a.r_hill = r_hill;
    a.facc1 = facc1, a.facc2 = facc2;
    a.frac1 = frac1, a.frac2 = frac2;
    a.density_floor = v.p.sigma_floor * v.p.sigma0;
    // rings whose centres can lie within the accretion radius of the planet (with a ring of slack on either side)
    const double rp = sqrt(x * x + y * y), reach = frac * r_hill;
    int lo = 0, hi = v.nr;
    while (lo < v.nr && c->h_rmed[lo] < rp - reach)
	++lo;
    while (hi > lo && c->h_rmed[hi - 1] > rp + reach)
	--hi;
    a.ring_lo = lo > 0 ? lo - 1 : 0;
    a.ring_hi = hi < v.nr ? hi + 1 : v.nr;
    const int nrings = a.ring_hi - a.ring_lo;
    double *d_out = c->force4;
    if (nrings > 0) {
	// The reference's stored PRESSURE / SCALE_HEIGHT keep describing the state before the accretion until the end of the
	// step (fargo_dev.h:PreState): keep the rows about to change.  Several bodies may accrete before one step: the kept
	// band grows to the union, and rows outside the band kept so far have not been touched yet.
	const size_t rowb = (size_t)v.ns * sizeof(double);
	if (!c->sig_pre && dalloc(c, &c->sig_pre, (size_t)v.nr * v.ns))
	    return 1;
	if (v.p.adiabatic && !c->e_pre && dalloc(c, &c->e_pre, (size_t)v.nr * v.ns))
	    return 1;
	auto keep = [&](int lo_, int hi_) -> cudaError_t {
	    if (hi_ <= lo_)
		return cudaSuccess;
	    const size_t off = (size_t)lo_ * v.ns, len = (size_t)(hi_ - lo_) * rowb;
	    cudaError_t e = cudaMemcpyAsync(c->sig_pre + off, c->sigma + off, len, cudaMemcpyDeviceToDevice, c->stream);
	    if (e == cudaSuccess && v.p.adiabatic)
		e = cudaMemcpyAsync(c->e_pre + off, EN(c) + off, len, cudaMemcpyDeviceToDevice, c->stream);
	    return e;
	};
	if (c->pre_hi <= c->pre_lo) {
	    CUDA_OK(keep(a.ring_lo, a.ring_hi));
	    c->pre_lo = a.ring_lo, c->pre_hi = a.ring_hi;
	} else {
	    CUDA_OK(keep(a.ring_lo, std::min(a.ring_hi, c->pre_lo)));
	    CUDA_OK(keep(std::max(a.ring_lo, c->pre_hi), a.ring_hi));
	    if (a.ring_lo > c->pre_hi) // rows between two disjoint bands: untouched so far
		CUDA_OK(keep(c->pre_hi, a.ring_lo));
	    if (a.ring_hi < c->pre_lo)
		CUDA_OK(keep(a.ring_hi, c->pre_lo));
	    c->pre_lo = std::min(c->pre_lo, a.ring_lo), c->pre_hi = std::max(c->pre_hi, a.ring_hi);
	}
	const unsigned gx = (unsigned)((v.ns + ACC_THREADS - 1) / ACC_THREADS);
	const int nblocks = (int)gx * nrings;
	if ((size_t)nblocks * 3 > c->partials_n)
	    return fail("partials buffer too small for the accretion");
	dim3 grid(gx, (unsigned)nrings);
	if (viscous)
	    LAUNCH(c, k_accrete_viscous, grid, ACC_THREADS, 0, v, c->sigma, EN(c), VRA(c), VPA(c), a, pre_state(c), c->partials);
	else
	    LAUNCH(c, k_accrete_kley, grid, ACC_THREADS, 0, v, c->sigma, EN(c), VRA(c), VPA(c), a, c->partials);
	LAUNCH(c, k_accrete_final, 1, 96, 0, c->partials, nblocks, d_out);
    } else {
	CUDA_OK(cudaMemsetAsync(d_out, 0, 3 * sizeof(double), c->stream));
    }
    if (v.nranks > 1) { // MPI_Allreduce(SUM), accretion.cpp:199-213
	NCCL_OK(g_nccl.AllReduce(d_out, d_out, 3, ncclFloat64, 0 /* ncclSum */, c->comm, c->stream));
	c->launches++;
    }
    CUDA_OK(cudaMemcpyAsync(out3, d_out, 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int fargo_accrete_kley(fargo_ctx *c, double x, double y, double r_hill, double facc, double frac, double out3[3])
{
    return accrete_zones(c, x, y, r_hill, 1.0 / 3.0 * facc, 2.0 / 3.0 * facc, frac, 0.5 * frac, out3);
}
extern "C" int fargo_accrete_sinkhole(fargo_ctx *c, double x, double y, double r_hill, double facc, double frac, double out3[3])
{
    return accrete_zones(c, x, y, r_hill, facc, 0.0, frac, 0.0, out3); // distance < 0 * r_hill never holds: one zone
}

// AccreteOntoSinglePlanetViscous (accretion.cpp:335-417): facc = dt * 3 pi * accretion efficiency; the removed fraction of a cell
// is facc * nu(cell) * 3 / (pi d_max^2) * (1 - distance / d_max), d_max = frac * r_hill
extern "C" int fargo_accrete_viscous(fargo_ctx *c, double x, double y, double r_hill, double facc, double frac, double out3[3])
{
    const double dist_max = r_hill * frac;		    // accretion.cpp:367
    const double f_const = 3.0 / M_PI / pow(dist_max, 2); // :368
    // one zone of radius frac * r_hill; the second pair of slots carries f_const and d_max (k_accrete_viscous)
    return accrete_zones(c, x, y, r_hill, facc, f_const, frac, dist_max, out3, true);
}

// ComputeCircumPlanetaryMasses (circumplanetary_mass.cpp:11-51): the "mdcp" column of monitor/nbodyK.dat
extern "C" int fargo_circumplanetary_mass(fargo_ctx *c, double x, double y, double roche_radius, double *out)
{
    CUDA_OK(cudaSetDevice(c->device));
    const DevView &v = c->v;
    const double rp = sqrt(x * x + y * y);
    int lo = v.first_active, hi = v.active_size;
    while (lo < hi && c->h_rmed[lo] < rp - roche_radius)
	++lo;
    while (hi > lo && c->h_rmed[hi - 1] > rp + roche_radius)
	--hi;
    double *d_out = c->force4; // 4 doubles of device scratch
    const int nrings = hi - lo;
    if (nrings > 0) {
	const unsigned gx = (unsigned)((v.ns + 127) / 128);
	const int nblocks = (int)gx * nrings;
	if ((size_t)nblocks * 3 > c->partials_n)
	    return fail("partials buffer too small for the circumplanetary mass");
	dim3 grid(gx, (unsigned)nrings);
	LAUNCH(c, k_circumplanetary_mass, grid, 128, 0, v, c->sigma, x, y, roche_radius, lo, c->partials);
	LAUNCH(c, k_accrete_final, 1, 96, 0, c->partials, nblocks, d_out);
    } else {
	CUDA_OK(cudaMemsetAsync(d_out, 0, 3 * sizeof(double), c->stream));
    }
    if (v.nranks > 1) { // MPI_Allreduce(SUM), circumplanetary_mass.cpp:46
	NCCL_OK(g_nccl.AllReduce(d_out, d_out, 1, ncclFloat64, 0 /* ncclSum */, c->comm, c->stream));
	c->launches++;
    }
    CUDA_OK(cudaMemcpyAsync(out, d_out, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

// monitor/Quantities.dat sums (quantities.cpp:51-480 through output::write_quantities, output.cpp:326-520)
extern "C" int fargo_monitor_quantities(fargo_ctx *c, double radius_limit, double out8[8])
{
    CUDA_OK(cudaSetDevice(c->device));
    if (c->v_mid)
	return fail("fargo_monitor_quantities called mid-step");
    const int nact = c->v.active_size - c->v.first_active;
    const unsigned gx = (unsigned)((c->v.ns + 4 * MQ_THREADS - 1) / (4 * MQ_THREADS));
    const int nblocks = (int)gx * (nact > 0 ? nact : 0);
    if ((size_t)(nblocks + 1) * MQ_N > c->partials_n)
	return fail("partials buffer too small for the monitor sums");
    double *d_out = c->partials + (size_t)nblocks * MQ_N; // behind the partials
    if (nblocks > 0) {
	dim3 grid(gx, (unsigned)nact);
	LAUNCH(c, k_monitor_quantities, grid, MQ_THREADS, 0, c->v, c->sigma, EN(c), VRA(c), VPA(c), c->qplus, c->qminus, radius_limit,
	       c->partials);
	LAUNCH(c, k_monitor_final, 1, 32 * MQ_N, 0, c->partials, nblocks, d_out);
    } else {
	CUDA_OK(cudaMemsetAsync(d_out, 0, MQ_N * sizeof(double), c->stream));
    }
    if (c->v.nranks > 1) { // MPI_Allreduce / MPI_Reduce(SUM), quantities.cpp:73, 274, ...
	NCCL_OK(g_nccl.AllReduce(d_out, d_out, MQ_N, ncclFloat64, 0 /* ncclSum */, c->comm, c->stream));
	c->launches++;
    }
    CUDA_OK(cudaMemcpyAsync(out8, d_out, MQ_N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}

// WriteMassFlow (parameters.cpp:334-335, TransportEuler.cpp:610-616): see fargo_b200.h
extern "C" int fargo_track_massflow(fargo_ctx *c, int on)
{
    CUDA_OK(cudaSetDevice(c->device));
    if (on && !c->massflow && dalloc(c, &c->massflow, (size_t)(c->v.nr + 1) * c->v.ns)) // zeroed
	return 1;
    c->track_massflow = on != 0;
    return 0;
}
extern "C" int fargo_clear_massflow(fargo_ctx *c)
{
    CUDA_OK(cudaSetDevice(c->device));
    if (c->massflow)
	CUDA_OK(cudaMemsetAsync(c->massflow, 0, (size_t)(c->v.nr + 1) * c->v.ns * sizeof(double), c->stream));
    return 0;
}

// MassDelta.Inner / OuterWaveDampingMassCreation / Removal (damping.cpp:335-357, 394-420 and the _zero / _mean siblings; columns
// 21-24 of monitor/Quantities.dat).  While tracked, the zones of Sigma are damped by k_damping in its own pass (the other fields stay
// folded into the transport kernel): the mass change of a cell needs Sigma before and after.
extern "C" int fargo_track_damping_mass(fargo_ctx *c, int on)
{
    CUDA_OK(cudaSetDevice(c->device));
    if (on && !c->dmass && dalloc(c, &c->dmass, (size_t)4 * c->v.ns)) // zeroed
	return 1;
    if (!on)
	c->dmass = nullptr; // stays allocated until the context goes (dev_allocs)
    return 0;
}
extern "C" int fargo_damping_mass(fargo_ctx *c, double out4[4], int reset)
{
    CUDA_OK(cudaSetDevice(c->device));
    if (!c->dmass)
	return fail("fargo_damping_mass: not tracked (fargo_track_damping_mass)");
    const int ns = c->v.ns;
    std::vector<double> h((size_t)4 * ns);
    CUDA_OK(cudaMemcpyAsync(h.data(), c->dmass, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    double sums[4] = {0.0, 0.0, 0.0, 0.0};
    for (int q = 0; q < 4; ++q)
	for (int j = 0; j < ns; ++j)
	    sums[q] += h[(size_t)q * ns + j];
    if (c->v.nranks > 1) { // MPI_Reduce(SUM), output.cpp:446-453
	CUDA_OK(cudaMemcpyAsync(c->force4, sums, 4 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	NCCL_OK(g_nccl.AllReduce(c->force4, c->force4, 4, ncclFloat64, 0 /* ncclSum */, c->comm, c->stream));
	c->launches++;
	CUDA_OK(cudaMemcpyAsync(sums, c->force4, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	CUDA_OK(cudaStreamSynchronize(c->stream));
    }
    for (int q = 0; q < 4; ++q)
	out4[q] = sums[q];
    if (reset)
	CUDA_OK(cudaMemsetAsync(c->dmass, 0, h.size() * sizeof(double), c->stream));
    return 0;
}

// MassDelta.Inner / OuterBoundaryInflow / Outflow (TransportEuler.cpp:578-608; columns 17-20 of monitor/Quantities.dat, reset after
// every row: output.cpp:493)
extern "C" int fargo_track_boundary_flow(fargo_ctx *c, int on)
{
    CUDA_OK(cudaSetDevice(c->device));
    if (on && !c->bflow && dalloc(c, &c->bflow, (size_t)4 * c->v.ns)) // zeroed
	return 1;
    if (!on && c->bflow) {
	if (c->track_massflow)
	    return fail("fargo_track_boundary_flow(0) while the mass-flow grid is tracked");
	c->bflow = nullptr; // stays allocated until the context goes (dev_allocs)
    }
    return 0;
}
extern "C" int fargo_boundary_flow(fargo_ctx *c, double out4[4], int reset)
{
    CUDA_OK(cudaSetDevice(c->device));
    if (!c->bflow)
	return fail("fargo_boundary_flow: not tracked (fargo_track_boundary_flow)");
    const int ns = c->v.ns;
    std::vector<double> h((size_t)4 * ns);
    CUDA_OK(cudaMemcpyAsync(h.data(), c->bflow, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    double sums[4] = {0.0, 0.0, 0.0, 0.0};
    for (int q = 0; q < 4; ++q)
	for (int j = 0; j < ns; ++j)
	    sums[q] += h[(size_t)q * ns + j];
    if (c->v.nranks > 1) { // MPI_Reduce(SUM), output.cpp:438-445: only the first and the last rank hold something
	CUDA_OK(cudaMemcpyAsync(c->force4, sums, 4 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	NCCL_OK(g_nccl.AllReduce(c->force4, c->force4, 4, ncclFloat64, 0 /* ncclSum */, c->comm, c->stream));
	c->launches++;
	CUDA_OK(cudaMemcpyAsync(sums, c->force4, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	CUDA_OK(cudaStreamSynchronize(c->stream));
    }
    for (int q = 0; q < 4; ++q)
	out4[q] = sums[q];
    if (reset)
	CUDA_OK(cudaMemsetAsync(c->bflow, 0, h.size() * sizeof(double), c->stream));
    return 0;
}

// The fused source-term kernel evaluates the potential of a cell in registers and stores nothing.  With `on`, every following
// fargo_kick also runs k_potential into the POTENTIAL grid (the staged kernels always do), so that fargo_monitor_disk can form the
// reference's "potential energy" and "gravitational torque" columns, which read the grid of the LAST kick's start (output.cpp:413,
// gas_torques.cpp:122-153).  A host switches it on for the step that ends on a monitor time (one extra pass, 0.3 ms at
// 8192 x 16384) and off again.
extern "C" int fargo_keep_potential(fargo_ctx *c, int on)
{
    c->keep_pot = on != 0;
    return 0;
}

// The mass-weighted columns of monitor/Quantities.dat (output.cpp:373-423): see k_monitor_disk.  Per-ring sums on the device,
// summed over the ranks ring by ring (every rank contributes the rings it owns), then walked in ring order on the host like
// the reference's root (quantities.cpp:213-233).
extern "C" int fargo_monitor_disk(fargo_ctx *c, double radius_limit, double mass_fraction, double frame_angle, double out9[9])
{
    CUDA_OK(cudaSetDevice(c->device));
    if (c->v_mid)
	return fail("fargo_monitor_disk called mid-step");
    const int nrg = c->v.p.nrad;
    const unsigned gx = (unsigned)((c->v.ns + 4 * MQ_THREADS - 1) / (4 * MQ_THREADS));
    const size_t npart = (size_t)gx * c->v.nr * MD_N, nrings = (size_t)MD_N * nrg;
    if (npart > c->partials_n)
	return fail("partials buffer too small for the per-ring monitor sums");
    double *d_rings = c->mon_rings;
    CUDA_OK(cudaMemsetAsync(d_rings, 0, nrings * sizeof(double), c->stream));
    dim3 grid(gx, (unsigned)c->v.nr);
    LAUNCH(c, k_monitor_disk, grid, MQ_THREADS, 0, c->v, c->sigma, EN(c), VRA(c), VPA(c), c->pot, radius_limit, cos(frame_angle),
	   sin(frame_angle), c->partials);
    LAUNCH(c, k_monitor_disk_rings, (unsigned)((c->v.nr + 127) / 128), 128, 0, c->v, c->partials, (int)gx, nrg, d_rings);
    if (c->v.nranks > 1) {
	NCCL_OK(g_nccl.AllReduce(d_rings, d_rings, nrings, ncclFloat64, 0 /* ncclSum */, c->comm, c->stream));
	c->launches++;
    }
    std::vector<double> h(nrings);
    CUDA_OK(cudaMemcpyAsync(h.data(), d_rings, nrings * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    double sums[MD_N] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int q = 1; q < MD_N; ++q)
	for (int i = 0; i < nrg; ++i)
	    sums[q] += h[(size_t)q * nrg + i];
    double radius = 0.0, current = 0.0;
    for (int i = 1; i < nrg - 1; ++i) { // the root walks the active rings of every rank, counter from 1 (split.cpp:339-343)
	current += h[i];
	if (current > mass_fraction * sums[1]) {
	    const double ri = c->h_radii[i], rs = c->h_radii[i + 1];
	    double rm = 2.0 / 3.0 * (pow(rs, 3) - pow(ri, 3));
	    radius = rm / (pow(rs, 2) - pow(ri, 2)); // GlobalRmed[i] (init.cpp:178-179)
	    break;
	}
    }
    out9[0] = radius;
    out9[1] = sums[1] > 0.0 ? sums[2] / sums[1] : 0.0;
    out9[2] = sums[1] > 0.0 ? sums[3] / sums[1] : 0.0;
    out9[3] = sums[1] > 0.0 ? sums[4] / sums[1] : 0.0;
    out9[4] = sums[1];
    out9[5] = sums[5];
    out9[6] = sums[6];
    // the reference reads its POTENTIAL grid as the last kick left it (zeros before the first step).  The fused kernels keep the
    // potential in registers: the grid is only that of the last kick if the kick was told to store it (fargo_keep_potential)
    const bool pot_ok = c->v.p.body_force_from_potential && c->pot_state != 2;
    out9[7] = pot_ok ? -(sums[1] > 0.0 ? sums[7] / sums[1] : 0.0) : std::nan("");
    out9[8] = pot_ok ? sums[8] : std::nan("");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// per-kernel device timing for bench.py (CUDA events on the context's stream)
extern "C" int fargo_profile_enable(fargo_ctx *c, int on)
{
    CUDA_OK(cudaSetDevice(c->device));
    prof_collect(c);
    c->profiling = on != 0;
    if (on) {
	for (auto &k : c->kstats) {
	    k.ms = 0;
	    k.n = 0;
	}
	c->host_turnaround_ms = 0.0;
	c->host_turnaround_n = 0;
    }
    return 0;
}
// writes "name ms count\n" lines into buf; returns the number of bytes needed
extern "C" int fargo_profile_report(fargo_ctx *c, char *buf, int buflen)
{
    cudaSetDevice(c->device);
    prof_collect(c);
    std::string out;
    char line[256];
    for (auto &k : c->kstats) {
	snprintf(line, sizeof(line), "%s %.6f %lld\n", k.name.c_str(), k.ms, k.n);
	out += line;
    }
    if (c->host_turnaround_n > 0) {
	snprintf(line, sizeof(line), "host:cfl-result->first-launch %.6f %lld\n", c->host_turnaround_ms, c->host_turnaround_n);
	out += line;
    }
    if (buf && buflen > 0) {
	strncpy(buf, out.c_str(), buflen - 1);
	buf[buflen - 1] = 0;
    }
    return (int)out.size() + 1;
}

// device-side wall clock for callers that cannot see the context's stream (bench.py): 4 event slots
extern "C" int fargo_event_record(fargo_ctx *c, int slot)
{
    if (slot < 0 || slot >= 4)
	return fail("event slot %d out of range", slot);
    CUDA_OK(cudaSetDevice(c->device));
    if (!c->ev_user[slot])
	CUDA_OK(cudaEventCreate(&c->ev_user[slot]));
    CUDA_OK(cudaEventRecord(c->ev_user[slot], c->stream));
    return 0;
}
extern "C" int fargo_event_elapsed_ms(fargo_ctx *c, int slot_a, int slot_b, double *ms_out)
{
    CUDA_OK(cudaSetDevice(c->device));
    if (slot_a < 0 || slot_a >= 4 || slot_b < 0 || slot_b >= 4 || !c->ev_user[slot_a] || !c->ev_user[slot_b])
	return fail("event slots not recorded");
    CUDA_OK(cudaEventSynchronize(c->ev_user[slot_b]));
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, c->ev_user[slot_a], c->ev_user[slot_b]));
    *ms_out = ms;
    return 0;
}

// device self-test of the branch-free arithmetic (fargo_math.h) against the plain operators; counts[4] =
// {division mismatches, sqrt mismatches, exp mismatches, divisions that took the fast path}
extern "C" int fargo_selftest_math(fargo_ctx *c, unsigned long long seed, int blocks, int per_thread, int wide,
				    unsigned long long *counts4)
{
    CUDA_OK(cudaSetDevice(c->device));
    unsigned long long *d = nullptr;
    CUDA_OK(cudaMalloc((void **)&d, 4 * sizeof(unsigned long long)));
    CUDA_OK(cudaMemsetAsync(d, 0, 4 * sizeof(unsigned long long), c->stream));
    k_selftest_math<<<blocks, 256, 0, c->stream>>>(seed, per_thread, wide, d);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(counts4, d, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaFree(d));
    return 0;
}

// exp_ref on caller-provided arguments, for the host-side comparison with libm's exp
// the scan ring sums (kernels_ringsum.cuh) and the one-thread-per-ring chain on caller-provided rows: sums_scan[r] and
// sums_chain[r] must both equal the sequential sum  s = 0; for j: s += x[r][j]  bit for bit
extern "C" int fargo_selftest_ringsum(fargo_ctx *c, int nrows, int ns, const double *x_host, double *sums_scan, double *sums_chain)
{
    CUDA_OK(cudaSetDevice(c->device));
    if (nrows < 1 || ns < 1)
	return fail("selftest_ringsum: empty input");
    double *d_x = nullptr, *d_s = nullptr;
    CUDA_OK(cudaMalloc((void **)&d_x, (size_t)nrows * ns * sizeof(double)));
    CUDA_OK(cudaMalloc((void **)&d_s, (size_t)2 * nrows * sizeof(double)));
    CUDA_OK(cudaMemcpyAsync(d_x, x_host, (size_t)nrows * ns * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    DevView v = c->v;
    v.nr = nrows;
    v.ns = ns;
    LAUNCH(c, k_ring_sum_scan, (unsigned)((nrows + 3) / 4), 128, 0, v, d_x, d_s, nullptr, nullptr, 0.0, 2, nrows, ns);
    LAUNCH(c, k_ring_sum_chain, (unsigned)((nrows + 127) / 128), 128, 0, d_x, d_s + nrows, nrows, ns);
    CUDA_OK(cudaMemcpyAsync(sums_scan, d_s, (size_t)nrows * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaMemcpyAsync(sums_chain, d_s + nrows, (size_t)nrows * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    cudaFree(d_x);
    cudaFree(d_s);
    return 0;
}

extern "C" int fargo_selftest_exp(fargo_ctx *c, int n, const double *x_host, double *y_host)
{
    CUDA_OK(cudaSetDevice(c->device));
    double *d = nullptr;
    CUDA_OK(cudaMalloc((void **)&d, 2 * (size_t)n * sizeof(double)));
    CUDA_OK(cudaMemcpyAsync(d, x_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    k_selftest_exp<<<(n + 255) / 256, 256, 0, c->stream>>>(n, d, d + n);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(y_host, d + n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    CUDA_OK(cudaFree(d));
    return 0;
}

// ComputeDiskOnPlanetAccel (Force.cpp:23-122) + MPI_Allreduce(SUM, 4 doubles) (:115)
extern "C" int fargo_disk_on_body_accel(fargo_ctx *c, int body, double klahr_factor, double out4[4])
{
    CUDA_OK(cudaSetDevice(c->device));
    if (body < 0 || body >= c->v.b.n)
	return fail("body %d out of range (%d bodies set)", body, c->v.b.n);
    if (c->v_mid)
	return fail("fargo_disk_on_body_accel called mid-step");
    BodyForceIn B;
    B.x = c->v.b.x[body], B.y = c->v.b.y[body];
    B.a = sqrt(B.x * B.x + B.y * B.y); // t_planet::get_r
    B.klahr_factor = klahr_factor;
    B.r_sm = c->v.b.cubic_smoothing_radius[body];
    const int nact = c->v.active_size - c->v.first_active;
    const unsigned gx = (unsigned)((c->v.ns + 4 * DOB_THREADS - 1) / (4 * DOB_THREADS));
    const int nblocks = (int)gx * (nact > 0 ? nact : 0);
    if ((size_t)nblocks * 4 > c->partials_n)
	return fail("partials buffer too small for the force sums");
    double *d_out = c->force4;
    if (nblocks > 0) {
	dim3 grid(gx, (unsigned)nact);
	if (c->v.p.correct_disk_selfgravity) // the ring means land in vmean (rewritten by the next CFL / transport before use)
	    LAUNCH(c, k_sigma_ring_mean, (unsigned)((c->v.nr + 63) / 64), 64, 0, c->v, c->sigma, c->vmean);
	LAUNCH(c, k_disk_on_body, grid, DOB_THREADS, 0, c->v, c->sigma, EN(c), c->vmean, B, c->partials);
	LAUNCH(c, k_disk_on_body_final, 1, 128, 0, c->partials, nblocks, d_out);
    } else {
	CUDA_OK(cudaMemsetAsync(d_out, 0, 4 * sizeof(double), c->stream));
    }
    if (c->v.nranks > 1) {
	NCCL_OK(g_nccl.AllReduce(d_out, d_out, 4, ncclFloat64, 0 /* ncclSum */, c->comm, c->stream));
	c->launches++;
    }
    CUDA_OK(cudaMemcpyAsync(out4, d_out, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    return 0;
}
