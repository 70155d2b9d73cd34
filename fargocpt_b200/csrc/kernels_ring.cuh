// kernels_ring.cuh — ring-local work: boundary conditions, damping zones, CFL reduction.
#pragma once
#include "fargo_dev.h"
#include "kernels_source.cuh"

// ---------------------------------------------------------------------------------------------
// apply_boundary_condition (boundary_conditions.cpp:65-114): Sigma, e, v_rad, v_azi ghost rings, in the
// reference's order.  One thread per azimuthal column.
__device__ __forceinline__ void bc_scalar_col(const DevView &c, double *x, const double *x0, const int bc[2], int j)
{
    const int Irad = c.nr - 1;
    if (c.rank == 0) {
	if (bc[0] == FARGO_BC_ZEROGRADIENT)
	    AT(x, 0, j) = AT(x, 1, j); // zero_gradient.cpp:17-27
	else if (bc[0] == FARGO_BC_REFERENCE)
	    AT(x, 0, j) = AT(x0, 0, j); // reference.cpp:16-25
    }
    if (c.rank == c.nranks - 1) {
	if (bc[1] == FARGO_BC_ZEROGRADIENT)
	    AT(x, Irad, j) = AT(x, Irad - 1, j); // zero_gradient.cpp:56-67
	else if (bc[1] == FARGO_BC_REFERENCE)
	    AT(x, Irad, j) = AT(x0, Irad, j);
    }
}

// host-formed ghost values of the inner v_rad boundaries that do not depend on the state: keplerian_radial.cpp:18-39 and the
// viscous outflow of viscous.cpp:18-46 for a state-independent viscosity (interfaces 0 and 1)
struct BcVrad {
    double in[2];
};
__global__ void __launch_bounds__(256)
    k_boundary(const DevView c, double *__restrict__ sigma, double *__restrict__ energy, double *__restrict__ vr,
	       double *__restrict__ vp, const double *__restrict__ sigma0, const double *__restrict__ energy0,
	       const double *__restrict__ vr0, const double *__restrict__ vp0, const double vkep_inner,
	       const double vkep_outer, const BcVrad bv)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= c.ns)
	return;
    const bool first = (c.rank == 0), last = (c.rank == c.nranks - 1);
    bc_scalar_col(c, sigma, sigma0, c.p.bc_sigma, j);
    bc_scalar_col(c, energy, energy0, c.p.bc_energy, j);
    { // v_rad (vector grid: max_radial = nr)
	const int Irad = c.nr;
	switch (c.p.bc_vrad[0]) {
	case FARGO_BC_ZEROGRADIENT: // zero_gradient.cpp:29-40
	    if (first) {
		AT(vr, 0, j) = AT(vr, 2, j);
		AT(vr, 1, j) = AT(vr, 2, j);
	    }
	    break;
	case FARGO_BC_OUTFLOW: // outflow.cpp:16-35
	    if (first) {
		if (AT(vr, 2, j) > 0.0) {
		    AT(vr, 1, j) = 0.0;
		    AT(vr, 0, j) = 0.0;
		} else {
		    AT(vr, 1, j) = AT(vr, 2, j);
		    AT(vr, 0, j) = AT(vr, 2, j);
		}
	    }
	    break;
	case FARGO_BC_REFLECTING: // reflecting.cpp:15-26 — no rank guard in the reference
	    AT(vr, 0, j) = -AT(vr, 2, j);
	    AT(vr, 1, j) = 0;
	    break;
	case FARGO_BC_REFERENCE: // reference.cpp:27-38
	    if (first) {
		AT(vr, 0, j) = AT(vr0, 0, j);
		AT(vr, 1, j) = AT(vr0, 1, j);
	    }
	    break;
	case FARGO_BC_KEPLERIAN: // keplerian_radial.cpp:18-39
	case FARGO_BC_VISCOUS:	 // viscous.cpp:18-46
	    if (first) {
		AT(vr, 0, j) = bv.in[0];
		AT(vr, 1, j) = bv.in[1];
	    }
	    break;
	default:
	    break;
	}
	switch (c.p.bc_vrad[1]) {
	case FARGO_BC_ZEROGRADIENT: // zero_gradient.cpp:69-81
	    if (last) {
		AT(vr, Irad, j) = AT(vr, Irad - 2, j);
		AT(vr, Irad - 1, j) = AT(vr, Irad - 2, j);
	    }
	    break;
	case FARGO_BC_OUTFLOW: // outflow.cpp:37-57
	    if (last) {
		if (AT(vr, Irad - 2, j) < 0.0) {
		    AT(vr, Irad - 1, j) = 0.0;
		    AT(vr, Irad, j) = 0.0;
		} else {
		    AT(vr, Irad - 1, j) = AT(vr, Irad - 2, j);
		    AT(vr, Irad, j) = AT(vr, Irad - 2, j);
		}
	    }
	    break;
	case FARGO_BC_REFLECTING: // reflecting.cpp:28-40 — no rank guard
	    AT(vr, Irad, j) = -AT(vr, Irad - 2, j);
	    AT(vr, Irad - 1, j) = 0;
	    break;
	case FARGO_BC_REFERENCE:
	    if (last) {
		AT(vr, Irad, j) = AT(vr0, Irad, j);
		AT(vr, Irad - 1, j) = AT(vr0, Irad - 1, j);
	    }
	    break;

	default:
	    break;
	}
    }
    { // v_azi
	const int Irad = c.nr - 1;
	if (first) {
	    if (c.p.bc_vazi[0] == FARGO_BC_KEPLERIAN) // keplerian_azimuthal.cpp:19-39 (value computed on the host)
		AT(vp, 0, j) = vkep_inner;
	    else if (c.p.bc_vazi[0] == FARGO_BC_ZEROGRADIENT)
		AT(vp, 0, j) = AT(vp, 1, j);
	    else if (c.p.bc_vazi[0] == FARGO_BC_REFERENCE)
		AT(vp, 0, j) = AT(vp0, 0, j);
	    else if (c.p.bc_vazi[0] == FARGO_BC_BALANCED) // balanced.cpp:23-75 (value computed on the host)
		AT(vp, 0, j) = vkep_inner;
	    else if (c.p.bc_vazi[0] == FARGO_BC_ZEROSHEAR) { // zero_shear.cpp:20-36
		const double Omega_active = AT(vp, 1, j) / c.g.rmed[1];
		AT(vp, 0, j) = c.g.rmed[0] * Omega_active;
	    }
	}
	if (last) {
	    if (c.p.bc_vazi[1] == FARGO_BC_KEPLERIAN) // keplerian_azimuthal.cpp:41-60
		AT(vp, Irad, j) = vkep_outer;
	    else if (c.p.bc_vazi[1] == FARGO_BC_ZEROGRADIENT)
		AT(vp, Irad, j) = AT(vp, Irad - 1, j);
	    else if (c.p.bc_vazi[1] == FARGO_BC_REFERENCE)
		AT(vp, Irad, j) = AT(vp0, Irad, j);
	    else if (c.p.bc_vazi[1] == FARGO_BC_BALANCED)
		AT(vp, Irad, j) = vkep_outer;
	    else if (c.p.bc_vazi[1] == FARGO_BC_ZEROSHEAR) { // zero_shear.cpp:38-54
		const double Omega_active = AT(vp, Irad - 1, j) / c.g.rmed[Irad - 1];
		AT(vp, Irad, j) = c.g.rmed[Irad] * Omega_active;
	    }
	}
    }
}

// ---------------------------------------------------------------------------------------------
// damping zones (damping.cpp:311-752).  The per-ring factors exp(-dt*f/tau) are computed on the host with
// glibc (bit-identical to the reference) and uploaded; `ring_lo..ring_hi` is the ring range of one zone.
// type: FARGO_DAMP_INITIAL (X0 = x0 field), _ZERO (X0 = x0_const), _MEAN (X0 = ring mean, index-ordered sum).
// All zones of all fields go through ONE launch: blockIdx.y runs over the concatenated ring ranges of the jobs.
#define FARGO_MAX_DAMP_JOBS 8
struct DampJob {
    double *x;
    const double *x0, *expf;
    int ring_lo, ring_hi, type, row0; // row0 = first blockIdx.y of this job
    int outer;			       // 0: the inner damping zone, 1: the outer one
    double x0_const;
    double *dmass; // Sigma jobs of a run that tracks MassDelta's wave-damping terms (damping.cpp:335-357 and its siblings): the mass
		   // (Xnew - X) Surf a cell gained, stored at the cell's place in this scratch grid; nullptr otherwise
};
struct DampJobs {
    int n;
    DampJob j[FARGO_MAX_DAMP_JOBS];
};
__global__ void __launch_bounds__(256) k_damping(const DevView c, const DampJobs jobs)
{
    int q = 0;
#pragma unroll
    for (int k = 1; k < FARGO_MAX_DAMP_JOBS; ++k)
	if (k < jobs.n && (int)blockIdx.y >= jobs.j[k].row0)
	    q = k;
    const DampJob &J = jobs.j[q];
    const int ring = J.ring_lo + ((int)blockIdx.y - J.row0);
    if (ring >= J.ring_hi)
	return;
    double *__restrict__ x = J.x;
    const double *__restrict__ x0 = J.x0;
    const int type = J.type;
    __shared__ double mean_sh;
    int jbeg = blockIdx.x * blockDim.x + threadIdx.x, jstride = gridDim.x * blockDim.x;
    if (type == FARGO_DAMP_MEAN) {
	// ONE block handles the whole ring: thread 0 does the index-ordered sum (damping.cpp:578-585) before anyone
	// modifies the ring
	if (blockIdx.x != 0)
	    return;
	jbeg = threadIdx.x, jstride = blockDim.x;
	if (threadIdx.x == 0) {
	    double s = 0.0;
	    for (int j = 0; j < c.ns; ++j)
		s += AT(x, ring, j);
	    mean_sh = s / c.ns;
	    // the reference keeps the mean IN quantity0(n_radial, 0) (damping.cpp:578-585, 706-713): the zone's initial-value grid
	    // carries it from then on (reference boundaries, beta cooling towards the reference state).  Nobody reads x0 in this job.
	    const_cast<double *>(J.x0)[(size_t)ring * c.ns] = mean_sh;
	}
	__syncthreads();
    }
    const double mean = (type == FARGO_DAMP_MEAN) ? mean_sh : 0.0;
    const double ef = J.expf[ring];
    for (int j = jbeg; j < c.ns; j += jstride) {
	const double X = AT(x, ring, j);
	const double X0 = (type == FARGO_DAMP_INITIAL) ? AT(x0, ring, j) : (type == FARGO_DAMP_MEAN ? mean : J.x0_const);
	const double Xnew = (X - X0) * ef + X0;
	AT(x, ring, j) = Xnew;
	if (J.dmass) {
	    const double delta = Xnew - X;
	    AT(J.dmass, ring, j) = delta * c.g.surf[ring];
	}
    }
}

// ---------------------------------------------------------------------------------------------
// Ring means.  The reference sums v_azi(i, 0..Ns-1) strictly in index order (cfl.cpp:199-204,
// TransportEuler.cpp:179-187); the result feeds the CFL dt and the integer shifts, which must be bit-exact, so
// the order is kept: every LANE owns one ring and adds its ring's values one after the other.  A warp therefore
// works on 32 rings at once; 32x32 tiles are brought in with coalesced 8-byte cp.async copies through a 4-stage
// shared-memory ring (row stride 33 doubles: the transposed reads are conflict-free), so ~6 MB of loads are in
// flight chip-wide while the dependent DADD chains run.
// mode 0: only vmean (CFL).  mode 1: also Nshift[i] and the constant residual velocity (transport).
// k_ring_mean_generic: any Ns (8-byte cp.async).  k_ring_mean (below): even Ns, TMA bulk copies.
#define RM_STAGES 4

__global__ void __launch_bounds__(32)
    k_ring_mean_generic(const DevView c, const double *__restrict__ vp, double *__restrict__ vmean, int *__restrict__ nshift,
		double *__restrict__ vconst, const double dt, const int mode)
{
    __shared__ double tile[RM_STAGES][32][33];
    const int lane = threadIdx.x;
    const int ring0 = blockIdx.x * 32;
    const int nrows = min(32, c.nr - ring0);
    const int ns = c.ns;
    const int ntiles = (ns + 31) >> 5;
    auto issue = [&](const int t) {
	if (t < ntiles) {
	    const int j = (t << 5) + lane;
	    if (j < ns) {
		double(*dst)[33] = tile[t % RM_STAGES];
		const double *src = vp + (size_t)ring0 * ns + j;
		for (int r = 0; r < nrows; ++r)
		    cp_async_8(&dst[r][lane], src + (size_t)r * ns);
	    }
	}
	cp_async_commit();
    };
#pragma unroll
    for (int t = 0; t < RM_STAGES - 1; ++t)
	issue(t);
    double s = 0.0;
    for (int t = 0; t < ntiles; ++t) {
	cp_async_wait<RM_STAGES - 2>();
	__syncwarp();
	const double *rowp = tile[t % RM_STAGES][lane];
	const int kmax = min(32, ns - (t << 5));
	if (kmax == 32) {
#pragma unroll
	    for (int k = 0; k < 32; ++k)
		s += rowp[k];
	} else {
	    for (int k = 0; k < kmax; ++k)
		s += rowp[k];
	}
	__syncwarp();
	issue(t + RM_STAGES - 1);
    }
    if (lane < nrows) {
	const int i = ring0 + lane;
	const double mean = s / (double)ns;
	vmean[i] = mean;
	if (mode == 1) {
	    const double invdt = 1.0 / dt;
	    const double Ntilde = mean * c.g.invrmed[i] * dt * c.invdphi;
	    const double Nround = floor(Ntilde + 0.5);
	    nshift[i] = (int)Nround;
	    vconst[i] = (Ntilde - Nround) * c.g.rmed[i] * invdt * c.dphi;
	}
    }
}

// ring_mean_finish: mean, and for the transport also Nshift and the constant residual (TransportEuler.cpp:215-235)
__device__ __forceinline__ void ring_mean_finish(const DevView &c, const int i, const double s, double *__restrict__ vmean,
						  int *__restrict__ nshift, double *__restrict__ vconst, const double dt, const int mode)
{
    const double mean = s / (double)c.ns;
    vmean[i] = mean;
    if (mode == 1) {
	const double invdt = 1.0 / dt;
	const double Ntilde = mean * c.g.invrmed[i] * dt * c.invdphi;
	const double Nround = floor(Ntilde + 0.5);
	nshift[i] = (int)Nround;
	vconst[i] = (Ntilde - Nround) * c.g.rmed[i] * invdt * c.dphi;
    }
}

// TMA version: the kernel is bound by the latency of ONE dependent DADD chain per ring (Ns adds, ~10 clk each), so the
// only job of the memory side is to never let a chain wait.  Every lane fetches ITS ring's next `chunk` columns with
// one cp.async.bulk (1-D TMA) into its own shared-memory row, RM_STAGES chunks deep, completion through one mbarrier per
// stage; the row stride (chunk + 2 doubles) keeps the 16-byte LDS of the 32 lanes bank-conflict free.
// Dynamic shared memory: RM_STAGES * 32 * (chunk + 2) doubles + RM_STAGES mbarriers.  Requires even Ns and chunk.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(32)
    k_ring_mean(const DevView c, const double *__restrict__ vp, double *__restrict__ vmean, int *__restrict__ nshift,
		double *__restrict__ vconst, const double dt, const int mode, const int chunk)
{
    extern __shared__ __align__(128) unsigned char rm_smem[];
    const int lane = threadIdx.x;
    const int ring0 = blockIdx.x * 32;
    const int nrows = min(32, c.nr - ring0);
    const int ns = c.ns;
    const int stride = chunk + 2;
    double *tiles = reinterpret_cast<double *>(rm_smem);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(tiles + (size_t)RM_STAGES * 32 * stride);
    const int nchunks = (ns + chunk - 1) / chunk;
    if (lane == 0) {
#pragma unroll
	for (int st = 0; st < RM_STAGES; ++st)
	    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[st])), "r"(1));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const double *src_row = vp + (size_t)(ring0 + (lane < nrows ? lane : 0)) * ns;
    auto issue = [&](const int t) {
	if (t >= nchunks)
	    return;
	const int st = t % RM_STAGES;
	const int cols = min(chunk, ns - t * chunk);
	const unsigned bytes = (unsigned)cols * 8u;
	if (lane == 0)
	    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[st])), "r"(bytes * (unsigned)nrows)
			 : "memory");
	__syncwarp();
	if (lane < nrows)
	    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
			     smem_u32(tiles + ((size_t)st * 32 + lane) * stride)),
			 "l"(src_row + (size_t)t * chunk), "r"(bytes), "r"(smem_u32(&bars[st]))
			 : "memory");
    };
#pragma unroll
    for (int t = 0; t < RM_STAGES - 1; ++t)
	issue(t);
    double s = 0.0;
    for (int t = 0; t < nchunks; ++t) {
	__syncwarp(); // every lane is done reading the stage chunk t + RM_STAGES - 1 goes into
	issue(t + RM_STAGES - 1);
	const int st = t % RM_STAGES;
	const unsigned parity = (unsigned)(t / RM_STAGES) & 1u;
	asm volatile("{\n"
		     ".reg .pred p;\n"
		     "RM_WAIT_%=:\n"
		     "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		     "@p bra RM_DONE_%=;\n"
		     "bra RM_WAIT_%=;\n"
		     "RM_DONE_%=:\n"
		     "}" ::"r"(smem_u32(&bars[st])),
		     "r"(parity)
		     : "memory");
	const double2 *rowp = reinterpret_cast<const double2 *>(tiles + ((size_t)st * 32 + lane) * stride);
	const int n2 = min(chunk, ns - t * chunk) >> 1;
	int k = 0;
	for (; k + 8 <= n2; k += 8) {
	    double2 v[8];
#pragma unroll
	    for (int u = 0; u < 8; ++u)
		v[u] = rowp[k + u];
#pragma unroll
	    for (int u = 0; u < 8; ++u) {
		s += v[u].x;
		s += v[u].y;
	    }
	}
	for (; k < n2; ++k) {
	    const double2 v = rowp[k];
	    s += v.x;
	    s += v.y;
	}
    }
    if (lane < nrows)
	ring_mean_finish(c, ring0 + lane, s, vmean, nshift, vconst, dt, mode);
}

// MassDelta.Inner / OuterWaveDampingMassCreation / Removal: the per-cell mass changes k_damping left in `cellmass`, added per
// column in ring order over the ACTIVE rings of a zone (sum_without_ghost_cells) to acc[2][ns] = creation, removal
__global__ void __launch_bounds__(128)
    k_dmass_accumulate(const DevView c, const double *__restrict__ cellmass, const int ring_lo, const int ring_hi, double *__restrict__ acc)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= c.ns)
	return;
    double created = acc[j], removed = acc[c.ns + j];
    const int lo = max(ring_lo, c.first_active), hi = min(ring_hi, c.active_size);
    for (int i = lo; i < hi; ++i) {
	const double m = AT(cellmass, i, j);
	if (m > 0)
	    created += m;
	else
	    removed += -m;
    }
    acc[j] = created;
    acc[c.ns + j] = removed;
}

// ---------------------------------------------------------------------------------------------
// cfl::condition_cfl (cfl.cpp:185-382).  dt_cell = CFL / sqrt(A) with A the sum of the six squared inverse time
// scales; sqrt and the division are monotone under round-to-nearest, so min(dt_cell) == CFL / sqrt(max A) and the
// kernel reduces A exactly as the reference forms it (same operations, same order) and pays for one sqrt + one
// division per BLOCK.  NaN limits are ignored like the reference's `dt_cell < dt` does.  A thread owns 4
// consecutive columns of one ring, so the divisions by per-ring constants (cell sizes, sqrt(gamma)) go through
// shared exact reciprocals; what remains per cell is 2 IEEE divisions (by Sigma and by e) and one sqrt, and the
// kernel is bound by its 6 array reads.  Block-max via warp shuffles, then ONE atomicMin per block on the bit
// pattern of the (non-negative) double: a single-pass grid reduction.
__device__ __forceinline__ void atomic_min_pos_double(double *addr, double v)
{
    // for non-negative IEEE doubles the unsigned bit pattern is monotone in the value
    atomicMin(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

// per-ring constants of the CFL criterion: cell sizes with the reciprocals the divisions use
struct CflRing {
    double cell_size, dxr, dxa, cell2, sqg;	  // denominators
    double ycell, ydxr, ydxa, ycell2, ysqg;	  // fm_rcp_raw of them
    double ids, irb, iok, inv_limit, lf, C, vm;
};
// the six inverse time scales of one cell (cfl.cpp:240-328) against the arithmetic policy M; returns A = sum of squares
template <class M>
__device__ __forceinline__ double cfl_cell(const DevView &c, const CflRing &g, const int i, const bool adiabatic, const double s,
					    const double e, const double vr0, const double vr1, const double vp0, const double vp1,
					    const double qp, const double qm, FmAcc &A)
{
    const double vres = c.p.fast_transport ? vp0 - g.vm : vp0;
    const double cs = eos_cs_m<M>(c, i, s, e, A);
    const double invdt1 = M::div_y(cs, g.cell_size, g.ycell, A);
    const double invdt2 = M::div_y(vr0, g.dxr, g.ydxr, A);
    const double invdt3 = M::div_y(vres, g.dxa, g.ydxa, A);
    double invdt4;
    if (c.p.artificial_viscosity == FARGO_ARTVISC_SN) {
	double dvRadial = vr1 - vr0;
	double dvAzimuthal = vp1 - vp0;
	dvRadial = (dvRadial > 0.0) ? 0.0 : -dvRadial;
	dvAzimuthal = (dvAzimuthal > 0.0) ? 0.0 : -dvAzimuthal;
	invdt4 = 4.0 * (g.C * g.C) * stdmax(dvRadial / g.dxr, dvAzimuthal / g.dxa) * g.lf; // SN is not a throughput config
    } else { // TW form, also for ArtificialViscosity: None
	const double eps_rr = (vr1 - vr0) * g.ids;
	const double eps_pp = g.irb * ((vp1 - vp0) * c.invdphi + 0.5 * (vr1 + vr0));
	const double mdiv_V = -stdmin(eps_rr + eps_pp, 0.0);
	invdt4 = 4.0 * (g.C * g.C) * mdiv_V * g.lf;
    }
    double nu;
    if (c.p.viscous_alpha > 0) { // eos_nu with the division by sqrt(gamma) shared
	const double H = adiabatic ? M::div_y(cs, g.sqg, g.ysqg, A) * g.iok : cs * g.iok;
	nu = c.p.viscous_alpha * H * cs;
    } else {
	nu = c.p.constant_viscosity;
    }
    const double invdt5 = M::div_y(4.0 * nu, g.cell2, g.ycell2, A) * g.lf;
    double invdt6 = 0.0;
    if (adiabatic)
	invdt6 = g.inv_limit * fabs(M::div(qp - qm, e, A)) * g.lf;
    return invdt1 * invdt1 + invdt2 * invdt2 + invdt3 * invdt3 + invdt4 * invdt4 + invdt5 * invdt5 + invdt6 * invdt6;
}

// the thread's 4 consecutive columns j0 .. j0 + 3 of ring i (and the right-hand neighbour of v_azi), row-major arrays
struct CflIn {
    double S[4], E[4], V0[4], V1[4], P[5], QP[4], QM[4];
};
__device__ __forceinline__ void cfl_load4(CflIn &N, const int ns, const int i, const int j0, const bool adiabatic,
					   const double *__restrict__ sigma, const double *__restrict__ energy, const double *__restrict__ vr,
					   const double *__restrict__ vp, const double *__restrict__ qplus, const double *__restrict__ qminus)
{
    const bool vec = ((ns & 3) == 0);
    const size_t row = (size_t)i * ns;
    if (vec) {
	auto ld4 = [&](const double *base, double *x) {
	    const double2 a = *reinterpret_cast<const double2 *>(base + j0);
	    const double2 b = *reinterpret_cast<const double2 *>(base + j0 + 2);
	    x[0] = a.x, x[1] = a.y, x[2] = b.x, x[3] = b.y;
	};
	ld4(sigma + row, N.S);
	ld4(vr + row, N.V0);
	ld4(vr + row + ns, N.V1);
	ld4(vp + row, N.P);
	N.P[4] = vp[row + ((j0 + 4 == ns) ? 0 : j0 + 4)];
	if (adiabatic) {
	    ld4(energy + row, N.E);
	    ld4(qplus + row, N.QP);
	    ld4(qminus + row, N.QM);
	}
    } else {
#pragma unroll
	for (int k = 0; k < 5; ++k) {
	    const int j = j0 + k;
	    const int jj = (j < ns) ? j : j - ns; // only the neighbour column may wrap; surplus columns are masked by the caller
	    N.P[k] = vp[row + (jj < ns ? jj : 0)];
	    if (k < 4) {
		const int js = (j < ns) ? j : 0;
		N.S[k] = sigma[row + js];
		N.V0[k] = vr[row + js];
		N.V1[k] = vr[row + ns + js];
		if (adiabatic) {
		    N.E[k] = energy[row + js];
		    N.QP[k] = qplus[row + js];
		    N.QM[k] = qminus[row + js];
		}
	    }
	}
    }
    if (!adiabatic) {
#pragma unroll
	for (int k = 0; k < 4; ++k)
	    N.E[k] = N.QP[k] = N.QM[k] = 0.0;
    }
}

// the exact criterion of the thread's 4 columns j0 .. j0 + 3 (j0 < ns) of ring i, folded into Amax (max of A) and `best` (limits
// that are not of the CFL / sqrt(A) form)
__device__ __forceinline__ void cfl_exact4(const DevView &c, const int i, const int j0, const double vm,
					    const double *__restrict__ sigma, const double *__restrict__ energy, const double *__restrict__ vr,
					    const double *__restrict__ vp, const double *__restrict__ qplus, const double *__restrict__ qminus,
					    const double *__restrict__ cf_r, const double *__restrict__ cf_phi, double &Amax, double &best)
{
    const int ns = c.ns;
    const double CFL = c.p.cfl;
    const bool adiabatic = c.p.adiabatic != 0;
    CflRing g;
    g.lf = c.p.leapfrog ? 0.6 : 1.0;
    g.C = c.p.artificial_viscosity_factor;
    g.dxr = c.g.rsup[i] - c.g.rinf[i];
    g.dxa = c.g.rmed[i] * c.dphi;
    g.cell_size = stdmin(g.dxr, g.dxa);
    g.cell2 = g.cell_size * g.cell_size;
    g.sqg = c.sqrt_gamma;
    g.ycell = fm_rcp_raw_x(g.cell_size), g.ydxr = fm_rcp_raw_x(g.dxr), g.ydxa = fm_rcp_raw_x(g.dxa);
    g.ycell2 = fm_rcp_raw_x(g.cell2), g.ysqg = fm_rcp_raw_x(g.sqg);
    g.inv_limit = 1.0 / c.p.heating_cooling_cfl_limit;
    g.ids = c.g.invdiffrsup[i], g.irb = c.g.invrmed[i], g.iok = c.g.inv_omega_k[i];
    g.vm = vm;
    FmAcc A0; // validity of the five shared denominators
    fm_acc_nrm_x(A0, g.cell_size), fm_acc_nrm_x(A0, g.dxr), fm_acc_nrm_x(A0, g.dxa), fm_acc_nrm_x(A0, g.cell2), fm_acc_nrm_x(A0, g.sqg);
    const bool vec = ((ns & 3) == 0);
    const size_t row = (size_t)i * ns;
    CflIn N;
    cfl_load4(N, ns, i, j0, adiabatic, sigma, energy, vr, vp, qplus, qminus);
    // the four cells straight-line on the fast arithmetic, one validity test; cold redo with the plain operators
    double Ac[4];
    FmAcc A = A0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
	Ac[k] = cfl_cell<MathX>(c, g, i, adiabatic, N.S[k], N.E[k], N.V0[k], N.V1[k], N.P[k], N.P[k + 1], N.QP[k], N.QM[k], A);
    if (!fm_acc_ok_x(A)) {
#pragma unroll
	for (int k = 0; k < 4; ++k)
	    Ac[k] = cfl_cell<MathP<false>>(c, g, i, adiabatic, N.S[k], N.E[k], N.V0[k], N.V1[k], N.P[k], N.P[k + 1], N.QP[k], N.QM[k], A);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
	if (!vec && j0 + k >= ns)
	    continue;
	const double Ak = Ac[k];
	if (c.p.stabilize_viscosity == 2) { // per-cell min(dt_cell, -CFL / min(c_phi, c_r)), cfl.cpp:330-338
	    double dt_cell = CFL / sqrt(Ak);
	    const double cc = stdmin(cf_phi[row + j0 + k], cf_r[row + j0 + k]);
	    if (cc != 0.0)
		dt_cell = stdmin(dt_cell, -CFL / cc);
	    if (dt_cell < best)
		best = dt_cell;
	} else if (Ak > Amax) {
	    Amax = Ak;
	}
    }
}

// FARGO shear criterion of ring i (:207-220); the (0, 1) pair is the reference's initial dt_core
__device__ __forceinline__ double cfl_shear(const DevView &c, const int i, const double *__restrict__ vmean)
{
    const double CFL = c.p.cfl;
    const double denom = fabs(vmean[i] * c.g.invrmed[i] - vmean[i + 1] * c.g.invrmed[i + 1]) + 1.0e-100;
    double best = CFL * c.dphi / denom;
    if (i == c.first_active) {
	const double denom0 = fabs(vmean[0] * c.g.invrmed[0] - vmean[1] * c.g.invrmed[1]) + 1.0e-100;
	const double d0 = CFL * c.dphi / denom0;
	if (d0 < best)
	    best = d0;
    }
    return best;
}

// grid: x over azimuth (512 columns per block of 128 threads), y over the active rings
__global__ void __launch_bounds__(128, 4)
    k_cfl(const DevView c, const double *__restrict__ sigma, const double *__restrict__ energy,
	  const double *__restrict__ vr, const double *__restrict__ vp, const double *__restrict__ qplus,
	  const double *__restrict__ qminus, const double *__restrict__ cf_r, const double *__restrict__ cf_phi,
	  const double *__restrict__ vmean, double *__restrict__ dt_out)
{
    const int i = c.first_active + blockIdx.y;
    const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const double DMAX = 1.7976931348623157e308;
    double best = DMAX; // limits that are not of the CFL / sqrt(A) form
    double Amax = -1.0;
    if (blockIdx.x == 0 && threadIdx.x == 0)
	best = cfl_shear(c, i, vmean);
    if (j0 < c.ns)
	cfl_exact4(c, i, j0, vmean[i], sigma, energy, vr, vp, qplus, qminus, cf_r, cf_phi, Amax, best);
    // block reduction: max of A, min of the other limits
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
	const double oa = __shfl_xor_sync(0xffffffffu, Amax, o);
	const double ob = __shfl_xor_sync(0xffffffffu, best, o);
	if (oa > Amax)
	    Amax = oa;
	if (ob < best)
	    best = ob;
    }
    __shared__ double wa[4], wb[4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) {
	wa[w] = Amax;
	wb[w] = best;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
	for (int k = 1; k < 4; ++k) {
	    if (wa[k] > Amax)
		Amax = wa[k];
	    if (wb[k] < best)
		best = wb[k];
	}
	if (Amax >= 0.0) {
	    const double dt_cell = c.p.cfl / sqrt(Amax);
	    if (dt_cell < best)
		best = dt_cell;
	}
	if (best < DMAX)
	    atomic_min_pos_double(dt_out, best);
    }
}

// ---------------------------------------------------------------------------------------------
// The same reduction in two passes (round 2): a SCREEN that bounds A of every cell without an IEEE division or a square
// root, and the exact criterion above evaluated only where the maximum can be.
//
// k_cfl is bound by its FP64 work (2 IEEE divisions, a square root and five Markstein steps per cell: 84 FP64 instructions
// for 48 bytes), not by its six array reads.  The screen forms A~ of a cell from the same inputs with reciprocals that are
// good to 2^-60 (hardware seed + one cubic Newton step) and c_s^2 instead of c_s: every term of A~ is the reference's term to
// a relative 1e-14 (differences such as v_azi - <v_azi>, Q+ - Q-, the divergence of invdt4 are formed by the reference's own
// operations, so cancellation cannot amplify the deviation), hence |A~ - A| <= eps A with eps = 1e-10 to spare.  With
// L = max over all cells of A~, the cell that owns the true maximum of A satisfies A~ >= L (1 - 2 eps), so only blocks whose
// own maximum of A~ reaches L (1 - 1e-9) can own it: k_cfl_candidates evaluates those (typically one or two of 2.6e5) with
// cfl_exact4 and the result is the reference's dt bit for bit.  A block with a cell whose A~ is NaN / huge, or whose c_s^2 is
// negative (the reference's NaN limits are ignored: such a cell must not raise L), is always a candidate and does not count
// towards L; if L is tiny (< 1e-250: terms may have underflowed) every block is one.
__device__ __forceinline__ double rcp_screen(const double b)
{
    const double y0 = fm_rcp_seed(b);
    double e = fma(-b, y0, 1.0);
    e = fma(e, e, e);
    return fma(y0, e, y0);
}
__device__ __forceinline__ void atomic_max_pos_double(double *addr, double v)
{
    atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

// grid as k_cfl; bmax[blockIdx.y * gridDim.x + blockIdx.x] = the block's max of A~ (+inf: always a candidate), *lmax = L
#ifndef CFL_SCREEN_MINB
#define CFL_SCREEN_MINB 4
#endif
__global__ void __launch_bounds__(128, CFL_SCREEN_MINB)
    k_cfl_screen(const DevView c, const double *__restrict__ sigma, const double *__restrict__ energy, const double *__restrict__ vr,
		 const double *__restrict__ vp, const double *__restrict__ qplus, const double *__restrict__ qminus,
		 const double *__restrict__ vmean, double *__restrict__ dt_out, double *__restrict__ bmax, double *__restrict__ lmax)
{
    const int i = c.first_active + blockIdx.y;
    const int ns = c.ns;
    const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const bool adiabatic = c.p.adiabatic != 0;
    double At = -1.0;
    bool force = false;
    if (blockIdx.x == 0 && threadIdx.x == 0) // the shear criterion is exact as it is
	atomic_min_pos_double(dt_out, cfl_shear(c, i, vmean));
    if (j0 < ns) {
	const double vm = vmean[i];
	const double lf = c.p.leapfrog ? 0.6 : 1.0;
	const double C = c.p.artificial_viscosity_factor;
	const double dxr = c.g.rsup[i] - c.g.rinf[i];
	const double dxa = c.g.rmed[i] * c.dphi;
	const double idxr = fm_rcp_raw_x(dxr), idxa = fm_rcp_raw_x(dxa);
	const double icell = (dxr < dxa) ? idxr : idxa;
	const double icell2 = icell * icell;
	const double ids = c.g.invdiffrsup[i], irb = c.g.invrmed[i], iok = c.g.inv_omega_k[i];
	const double K4 = 4.0 * (C * C) * lf;
	const double kcs = c.p.gamma * (c.p.gamma - 1.0);
	const double cs_iso = c.g.cs_iso[i];
	double K5 = 0.0, i5c = 0.0; // invdt5 = K5 c_s^2 (alpha viscosity) or a constant of the ring
	if (c.p.viscous_alpha > 0)
	    K5 = 4.0 * c.p.viscous_alpha * iok * icell2 * lf * (adiabatic ? fm_rcp_raw_x(c.sqrt_gamma) : 1.0);
	else
	    i5c = 4.0 * c.p.constant_viscosity * icell2 * lf;
	const double K6 = (1.0 / c.p.heating_cooling_cfl_limit) * lf;
	const bool sn = c.p.artificial_viscosity == FARGO_ARTVISC_SN;
	const bool vec = ((ns & 3) == 0);
	CflIn N;
	cfl_load4(N, ns, i, j0, adiabatic, sigma, energy, vr, vp, qplus, qminus);
#pragma unroll
	for (int k = 0; k < 4; ++k) {
	    if (!vec && j0 + k >= ns)
		continue;
	    const double vr0 = N.V0[k], vr1 = N.V1[k], vp0 = N.P[k], vp1 = N.P[k + 1];
	    const double cs2 = adiabatic ? kcs * N.E[k] * rcp_screen(N.S[k]) : cs_iso * cs_iso;
	    const double t2 = vr0 * idxr;
	    const double vres = c.p.fast_transport ? vp0 - vm : vp0;
	    const double t3 = vres * idxa;
	    double t4;
	    if (sn) {
		double dvRadial = vr1 - vr0;
		double dvAzimuthal = vp1 - vp0;
		dvRadial = (dvRadial > 0.0) ? 0.0 : -dvRadial;
		dvAzimuthal = (dvAzimuthal > 0.0) ? 0.0 : -dvAzimuthal;
		t4 = K4 * stdmax(dvRadial * idxr, dvAzimuthal * idxa);
	    } else {
		const double eps_rr = (vr1 - vr0) * ids;
		const double eps_pp = irb * ((vp1 - vp0) * c.invdphi + 0.5 * (vr1 + vr0));
		t4 = K4 * -stdmin(eps_rr + eps_pp, 0.0);
	    }
	    const double t5 = (c.p.viscous_alpha > 0) ? K5 * cs2 : i5c;
	    double t6 = 0.0;
	    if (adiabatic)
		t6 = K6 * fabs((N.QP[k] - N.QM[k]) * rcp_screen(N.E[k]));
	    const double Ak = cs2 * icell2 + t2 * t2 + t3 * t3 + t4 * t4 + t5 * t5 + t6 * t6;
	    if (!(Ak < 1.0e300) || cs2 < 0.0)
		force = true;
	    else if (Ak > At)
		At = Ak;
	}
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
	const double oa = __shfl_xor_sync(0xffffffffu, At, o);
	if (oa > At)
	    At = oa;
    }
    force = __any_sync(0xffffffffu, force);
    __shared__ double wa[4];
    __shared__ int wf[4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) {
	wa[w] = At;
	wf[w] = force ? 1 : 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
	for (int k = 1; k < 4; ++k) {
	    if (wa[k] > At)
		At = wa[k];
	    force = force || (wf[k] != 0);
	}
	if (At > 0.0)
	    atomic_max_pos_double(lmax, At);
	bmax[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = force ? __longlong_as_double(0x7ff0000000000000LL) : At;
    }
}

// one thread per record of bmax; the CTA evaluates each of its candidate blocks (512 columns of one ring) exactly, a warp per
// 128 columns, like a block of k_cfl (the candidates are few: what counts is the latency of one of them)
__global__ void __launch_bounds__(128)
    k_cfl_candidates(const DevView c, const double *__restrict__ sigma, const double *__restrict__ energy, const double *__restrict__ vr,
		     const double *__restrict__ vp, const double *__restrict__ qplus, const double *__restrict__ qminus,
		     const double *__restrict__ vmean, const double *__restrict__ bmax, const double *__restrict__ lmax, const int nrec,
		     const int gx, double *__restrict__ dt_out)
{
    const double L = *lmax;
    const double thr = (L >= 1.0e-250) ? L * (1.0 - 1.0e-9) : -2.0;
    const int rec = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const double v = (rec < nrec) ? bmax[rec] : -3.0;
    __shared__ unsigned cand[4];
    const unsigned mine = __ballot_sync(0xffffffffu, v >= thr);
    if (lane == 0)
	cand[w] = mine;
    __syncthreads();
    if ((cand[0] | cand[1] | cand[2] | cand[3]) == 0u)
	return; // whole CTA
    double Amax = -1.0, best = 1.7976931348623157e308;
    for (int q = 0; q < 4; ++q) {
	unsigned mask = cand[q];
	while (mask) {
	    const int b = blockIdx.x * blockDim.x + q * 32 + (__ffs(mask) - 1);
	    mask &= mask - 1;
	    const int i = c.first_active + b / gx;
	    const int bx = b - (b / gx) * gx;
	    const int j0 = (bx * 128 + w * 32 + lane) * 4;
	    if (j0 < c.ns)
		cfl_exact4(c, i, j0, vmean[i], sigma, energy, vr, vp, qplus, qminus, nullptr, nullptr, Amax, best);
	}
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
	const double oa = __shfl_xor_sync(0xffffffffu, Amax, o);
	if (oa > Amax)
	    Amax = oa;
    }
    if (lane == 0 && Amax >= 0.0)
	atomic_min_pos_double(dt_out, c.p.cfl / sqrt(Amax));
}

// ---------------------------------------------------------------------------------------------
// EquationOfState: PVTE.  pvte::compute_gamma_mu (pvte_law.cpp:497-541) with the sound speed / scale height refreshes the
// reference wraps around it, per cell:
//   mode & 1: first recompute the scale height from the CURRENT grids and state (simulation.cpp:256-262 after Transport;
//             init_euler SourceEuler.cpp:272-275) — otherwise the lookup reads the STORED one (recalculate_viscosity :214-217,
//             whose scale height dates from the end of the previous step)
//   then the lookup at rho = Sigma / (density_factor H) and e / Sigma [cgs], and the scale height of the new grids is stored
//   (compute_sound_speed + compute_scale_height, :218-219 / :245-246).
__global__ void __launch_bounds__(256) k_pvte_refresh(const DevView c, const double *__restrict__ sigma,
						       const double *__restrict__ energy, const int mode)
{
    CELL_INDEX(c.nr);
    const size_t cell = (size_t)i * c.ns + j;
    const double s = AT(sigma, i, j), e = AT(energy, i, j);
    double H = c.pv.H[cell];
    if (mode & 1)
	H = eos_H_at(c, i, cell, eos_cs_at(c, i, cell, s, e));
    const double densityCGS = s / (c.p.density_factor * H) * c.p.density_cgs;
    const double energyCGS = e * c.p.energy_density_cgs / (s * c.p.surface_density_cgs);
    double geff, mu, g1;
    pv_lookup(c, densityCGS, energyCGS, geff, mu, g1);
    c.pv.geff[cell] = geff, c.pv.mu[cell] = mu, c.pv.g1[cell] = g1;
    const double cs = sqrt(g1 * (geff - 1.0) * e / s);
    c.pv.H[cell] = cs / (sqrt(g1)) * c.g.inv_omega_k[i];
}
__global__ void k_pvte_fill(const DevView c, const size_t n)
{ // init_eos_arrays (init.cpp:1197-1205)
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n)
	c.pv.geff[k] = c.p.gamma, c.pv.g1[k] = c.p.gamma, c.pv.mu[k] = c.p.mu, c.pv.H[k] = 0.0;
}
// The CFL criterion with per-cell gamma_eff / Gamma_1 (PVTE): one thread per cell on the plain operators (IEEE: the same dt
// as the keyed fast path would give), cfl.cpp:240-376.
__global__ void __launch_bounds__(128)
    k_cfl_cells(const DevView c, const double *__restrict__ sigma, const double *__restrict__ energy,
		const double *__restrict__ vr, const double *__restrict__ vp, const double *__restrict__ qplus,
		const double *__restrict__ qminus, const double *__restrict__ cf_r, const double *__restrict__ cf_phi,
		const double *__restrict__ vmean, double *__restrict__ dt_out)
{
    const int i = c.first_active + blockIdx.y;
    const int ns = c.ns;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const double DMAX = 1.7976931348623157e308;
    const double CFL = c.p.cfl;
    double best = DMAX;
    const double vm = vmean[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) { // FARGO shear criterion (:207-220), as in k_cfl
	const double denom = fabs(vm * c.g.invrmed[i] - vmean[i + 1] * c.g.invrmed[i + 1]) + 1.0e-100;
	best = CFL * c.dphi / denom;
	if (i == c.first_active) {
	    const double denom0 = fabs(vmean[0] * c.g.invrmed[0] - vmean[1] * c.g.invrmed[1]) + 1.0e-100;
	    const double d0 = CFL * c.dphi / denom0;
	    if (d0 < best)
		best = d0;
	}
    }
    if (j < ns) {
	const size_t cell = (size_t)i * ns + j;
	const int jn = (j + 1 == ns) ? 0 : j + 1;
	const double s = sigma[cell], e = c.p.adiabatic ? energy[cell] : 0.0;
	const double vr0 = vr[cell], vr1 = vr[cell + ns], vp0 = vp[cell], vp1 = vp[(size_t)i * ns + jn];
	const double dxr = c.g.rsup[i] - c.g.rinf[i], dxa = c.g.rmed[i] * c.dphi;
	const double cell_size = stdmin(dxr, dxa);
	const double lf = c.p.leapfrog ? 0.6 : 1.0, C = c.p.artificial_viscosity_factor;
	const double vres = c.p.fast_transport ? vp0 - vm : vp0;
	const double cs = eos_cs_at(c, i, cell, s, e);
	const double invdt1 = cs / cell_size, invdt2 = vr0 / dxr, invdt3 = vres / dxa;
	double invdt4;
	if (c.p.artificial_viscosity == FARGO_ARTVISC_SN) {
	    double dvR = vr1 - vr0, dvA = vp1 - vp0;
	    dvR = (dvR > 0.0) ? 0.0 : -dvR;
	    dvA = (dvA > 0.0) ? 0.0 : -dvA;
	    invdt4 = 4.0 * (C * C) * stdmax(dvR / dxr, dvA / dxa) * lf;
	} else {
	    const double eps_rr = (vr1 - vr0) * c.g.invdiffrsup[i];
	    const double eps_pp = c.g.invrmed[i] * ((vp1 - vp0) * c.invdphi + 0.5 * (vr1 + vr0));
	    const double mdiv_V = -stdmin(eps_rr + eps_pp, 0.0);
	    invdt4 = 4.0 * (C * C) * mdiv_V * lf;
	}
	const double nu = eos_nu_at(c, i, cell, s, e);
	const double invdt5 = 4.0 * nu / (cell_size * cell_size) * lf;
	double invdt6 = 0.0;
	if (c.p.adiabatic)
	    invdt6 = (1.0 / c.p.heating_cooling_cfl_limit) * fabs((qplus[cell] - qminus[cell]) / e) * lf;
	const double A = invdt1 * invdt1 + invdt2 * invdt2 + invdt3 * invdt3 + invdt4 * invdt4 + invdt5 * invdt5 + invdt6 * invdt6;
	double dt_cell = CFL / sqrt(A);
	if (c.p.stabilize_viscosity == 2) {
	    const double cc = stdmin(cf_phi[cell], cf_r[cell]);
	    if (cc != 0.0)
		dt_cell = stdmin(dt_cell, -CFL / cc);
	}
	if (dt_cell < best)
	    best = dt_cell;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
	const double ob = __shfl_xor_sync(0xffffffffu, best, o);
	if (ob < best)
	    best = ob;
    }
    if ((threadIdx.x & 31) == 0 && best < DMAX)
	atomic_min_pos_double(dt_out, best);
}

// derived fields on demand (downloads only): T, P, c_s, H, nu  (SourceEuler.cpp:957-1408, viscosity.cpp:98)
__global__ void __launch_bounds__(256) k_derived_field(const DevView c, const double *__restrict__ sigma,
							const double *__restrict__ energy, double *__restrict__ out,
							const int which)
{
    CELL_INDEX(c.nr);
    const double s = AT(sigma, i, j), e = AT(energy, i, j);
    const size_t cell = (size_t)i * c.ns + j;
    double v = 0.0;
    switch (which) {
    case FARGO_TEMPERATURE:
	if (c.p.adiabatic)
	    v = pv_mu(c, cell) / c.p.Rgas * (pv_geff(c, cell) - 1.0) * e / s;
	else
	    v = c.p.mu / c.p.Rgas * eos_P(c, i, s, e) / s;
	break;
    case FARGO_PRESSURE:
	v = eos_P_at(c, i, cell, s, e);
	break;
    case FARGO_SOUNDSPEED:
	v = eos_cs_at(c, i, cell, s, e);
	break;
    case FARGO_SCALE_HEIGHT: // PVTE: the SCALE_HEIGHT grid as stored
	v = c.pv.H ? c.pv.H[cell] : eos_H(c, i, eos_cs(c, i, s, e));
	break;
    case FARGO_VISCOSITY:
	v = eos_nu_at(c, i, cell, s, e);
	break;
    }
    AT(out, i, j) = v;
}

// ---------------------------------------------------------------------------------------------
// Ghost-ring exchange over peer memory (CommunicateBoundaries, commbound.cpp:98-182), receiving side.  The transport
// kernel's edge marches store the rings a neighbour needs into that neighbour's inbox and publish the step number in
// its arrival counters (kernels_azimuthal.cuh: AzSegs).  k_halo_unpack holds the receiver's stream until both counters
// have reached the step, then moves the inbox into the ghost rings.
struct HaloUnpack {
    const double *src[8];
    double *dst[8];
    int n;
    size_t len;
    const unsigned long long *flag[2];
    unsigned long long want;
    double *timed_out; // set to 1.0 if a neighbour's rings never arrived (read back with the next CFL result)
};
#define HALO_SPIN_LIMIT (1u << 26) // x >= 200 ns: gives up after roughly 15-30 s instead of hanging the GPU for ever
__global__ void __launch_bounds__(256) k_halo_unpack(const HaloUnpack u)
{
    if (threadIdx.x == 0) {
	for (int k = 0; k < 2; ++k) {
	    if (!u.flag[k])
		continue;
	    unsigned spins = 0;
	    while (*(const volatile unsigned long long *)u.flag[k] < u.want) {
		__nanosleep(200);
		if (++spins == HALO_SPIN_LIMIT) { // a neighbouring rank has died or fallen out of step: do not hang, report
		    *u.timed_out = 1.0;
		    break;
		}
	    }
	}
	__threadfence_system();
    }
    __syncthreads();
    const double *__restrict__ s = u.src[blockIdx.y];
    double *__restrict__ d = u.dst[blockIdx.y];
    const size_t i0 = ((size_t)blockIdx.x * 256 + threadIdx.x) * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k)
	if (i0 + k < u.len)
	    d[i0 + k] = __ldcv(s + i0 + k); // written by another GPU: never through a stale cache line
}
