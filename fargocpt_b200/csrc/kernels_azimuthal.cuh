// kernels_azimuthal.cuh — azimuthal half of Transport (TransportEuler.cpp:270-304, 416-466, 630-664), the integer
// FARGO shift (AdvectSHIFT :238-268), velocities from momenta (:498-535) and the floors (:123-131) in ONE kernel
// that keeps a ring segment in registers.
//
// Layout: a warp owns a window of 32 * AZ_NC consecutive OUTPUT columns of one ring, each lane AZ_NC consecutive
// columns (azimuthal neighbours are the thread's own registers or one warp shuffle away).  The segment is loaded already
// rotated by the ring's integer shift (output column j comes from pre-shift column j - Nshift[i]); the
// residual-velocity pass and the uniform pass then run on it in registers, quantity by quantity, with Sigma* and the
// upwind selectors shared by all six quantities.  Each van Leer pass invalidates 2 columns at either end of the window
// and v_azi needs one more on the left, so a window of 64 columns yields 54 finished ones ([6, 60), 16-byte aligned
// for vector stores); warps are independent — no shared memory, no block barrier.  The warp then marches outward
// ring by ring, carrying the previous ring's transported Sigma and rm+ (v_rad couples rings i-1 and i at the same
// output column).
// AZ_NC = 2 at 4-5 CTAs / SM (128 / 96 registers) beats 4 columns per lane at 3 CTAs / SM (168 registers, spills)
// although it recomputes 16 % instead of 9 % of the columns: 4.97 (4.77) vs 5.53 ms at 8192x16384.
//
// Arithmetic is the reference's, operation for operation (-fmad=false); the only algebraic liberties are exact
// ones: x - c*d == x + (-c)*d, dx + ksi == dx - |ksi| for ksi <= 0, 0.5 * (2ab / (a+b)) == ab / (a+b) inside the
// limiter's operand range (fargo_dev.h:limiter_nb), and the branch-free division of fargo_math.h (the compiler's own
// IEEE sequence, emitted straight-line so the Newton chains of columns and quantities interleave).  The hot path of a
// ring has NO branch: validity keys are accumulated and tested once per ring (az_ring).
#pragma once
#include "fargo_dev.h"
#include "fargo_math.h"

#ifndef AZ_NC
#define AZ_NC 2 // columns per lane
#endif
#define AZ_WIN (32 * AZ_NC) // columns per warp window
#define AZ_HL (AZ_NC == 2 ? 6 : 8) // invalid columns at the left end (5 needed, rounded up for aligned vector stores)
#ifndef AZ_MINB
#define AZ_MINB 5 // 96 registers (a few spills) at 20 warps / SM: 4.77 vs 4.96 ms with 128 registers at 16 warps / SM
#endif
#define AZ_HR 4	   // invalid columns at the right end
#define AZ_OUT (AZ_WIN - AZ_HL - AZ_HR)

__device__ __forceinline__ double shfl_from_left(const double x) { return __shfl_up_sync(0xffffffffu, x, 1); }
__device__ __forceinline__ double shfl_from_right(const double x) { return __shfl_down_sync(0xffffffffu, x, 1); }

struct AzRing {
    double dxtheta, invdxtheta, dxrad, invsurf;
};

// ComputeStarTheta (:416-466) for one base quantity B on the thread's columns: limited slopes, then the
// upwinded interface values.  pos[c]: ksi > 0 at interface c (between columns c-1 and c); cf[c] = +-(dxtheta -+ ksi).
// FAST: branch-free arithmetic, validity accumulated in A (checked once per ring by the kernel); !FAST: plain operators.
template <int LIM, bool FAST>
__device__ __forceinline__ void az_star(const double (&B)[AZ_NC], const AzRing &g, const bool (&pos)[AZ_NC], const double (&cf)[AZ_NC],
					 double (&star)[AZ_NC])
{
    const double Bm = shfl_from_left(B[AZ_NC - 1]);
    const double Bp = shfl_from_right(B[0]);
    double dq[AZ_NC + 1];
    dq[0] = B[0] - Bm;
#pragma unroll
    for (int c = 1; c < AZ_NC; ++c)
	dq[c] = B[c] - B[c - 1];
    dq[AZ_NC] = Bp - B[AZ_NC - 1];
    double D[AZ_NC];
#pragma unroll
    for (int c = 0; c < AZ_NC; ++c) {
	if (FAST)
	    D[c] = limiter_nb<LIM, true>(dq[c + 1], dq[c]) * g.invdxtheta; // key-free: the caller has keyed B
	else
	    D[c] = 0.5 * flux_limiter<LIM>(dq[c + 1], dq[c]) * g.invdxtheta;
    }
    const double Dm = shfl_from_left(D[AZ_NC - 1]);
    star[0] = fm_madd(cf[0], pos[0] ? Dm : D[0], pos[0] ? Bm : B[0]);
#pragma unroll
    for (int c = 1; c < AZ_NC; ++c)
	star[c] = fm_madd(cf[c], pos[c] ? D[c - 1] : D[c], pos[c] ? B[c - 1] : B[c]);
}

// VanLeerTheta (:630-664) conservative update of one quantity from its interface fluxes
__device__ __forceinline__ void az_update(double (&Q)[AZ_NC], const double (&G)[AZ_NC], const AzRing &g)
{
    const double Gp = shfl_from_right(G[0]);
#pragma unroll
    for (int c = 0; c < AZ_NC; ++c) {
	double varq = G[c];
	varq -= (c == AZ_NC - 1) ? Gp : G[(c + 1) % AZ_NC];
	Q[c] = fm_madd(varq, g.invsurf, Q[c]);
    }
}

// QuantitiesAdvection (:292-304): Sigma* from the current Sigma, Sigma_int = copy, then rm+, rm-, am+, am-, (e), Sigma.
// Q index: 0 rm+, 1 rm-, 2 am+, 3 am-, 4 e, 5 Sigma.
template <int LIM, bool ADI, bool FAST>
__device__ __forceinline__ void az_pass(double (&Q)[6][AZ_NC], const double (&u)[AZ_NC], const AzRing &g, const double dt, FmAcc &A)
{
    bool pos[AZ_NC];
    double cf[AZ_NC];
#pragma unroll
    for (int c = 0; c < AZ_NC; ++c) {
	const double ksi = u[c] * dt;
	pos[c] = ksi > 0.0;
	const double coef = g.dxtheta - fabs(ksi); // (dxtheta - ksi) for ksi > 0, (dxtheta + ksi) otherwise
	cf[c] = pos[c] ? coef : -coef;
    }
    double starS[AZ_NC];
    az_star<LIM, FAST>(Q[5], g, pos, cf, starS);
    double yS[AZ_NC];
#pragma unroll
    for (int c = 0; c < AZ_NC; ++c) {
	if (FAST) {
	    yS[c] = fm_rcp_raw(Q[5][c]);
	    fm_acc_nrm(A, Q[5][c]);
	}
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) {
	if (q == 4 && !ADI)
	    continue;
	double W[AZ_NC], st[AZ_NC], G[AZ_NC];
#pragma unroll
	for (int c = 0; c < AZ_NC; ++c) { // divise_polargrid (SideEuler.cpp:27-43)
	    if (FAST) {
		W[c] = fm_div_raw(Q[q][c], Q[5][c], yS[c]);
		fm_acc_nrm(A, W[c]); // exact zeros (v_rad == 0 at the boundaries) fail the key: that ring is redone
	    } else {
		W[c] = Q[q][c] / Q[5][c];
	    }
	}
	az_star<LIM, FAST>(W, g, pos, cf, st);
#pragma unroll
	for (int c = 0; c < AZ_NC; ++c)
	    G[c] = g.dxrad * st[c] * starS[c] * u[c];
	az_update(Q[q], G, g);
    }
    { // Sigma itself: QRStar == 1 (Sigma / Sigma_int with zero slopes)
	double G[AZ_NC];
#pragma unroll
	for (int c = 0; c < AZ_NC; ++c)
	    G[c] = g.dxrad * starS[c] * u[c];
	az_update(Q[5], G, g);
    }
}

// One ring segment of the seven input arrays, loaded already rotated by the ring's integer shift: output column j
// comes from pre-shift column j - Nshift[i] (AdvectSHIFT :238-268).
struct AzIn {
    double Q[6][AZ_NC], VP[AZ_NC];
};
template <bool ADI>
__device__ __forceinline__ void az_fetch(AzIn &N, const int ns, const int jout, const int nsh, const size_t row,
					  const double *__restrict__ t_rmp, const double *__restrict__ t_rmm,
					  const double *__restrict__ t_amp, const double *__restrict__ t_amm,
					  const double *__restrict__ t_e, const double *__restrict__ t_sigma,
					  const double *__restrict__ vp_old)
{
    int col = (jout - nsh) % ns; // pre-shift column of c = 0
    if (col < 0)
	col += ns;
#pragma unroll
    for (int k = 0; k < AZ_NC; ++k) {
	const size_t a = row + (size_t)col;
	N.Q[0][k] = t_rmp[a];
	N.Q[1][k] = t_rmm[a];
	N.Q[2][k] = t_amp[a];
	N.Q[3][k] = t_amm[a];
	N.Q[4][k] = ADI ? t_e[a] : 0.0;
	N.Q[5][k] = t_sigma[a];
	N.VP[k] = vp_old[a];
	col = (col + 1 == ns) ? 0 : col + 1;
    }
}
// the first and last of the thread's columns cover the sectors of its (rotated) segment
template <bool ADI>
__device__ __forceinline__ void az_prefetch(const int ns, const int jout, const int nsh, const size_t row,
					     const double *__restrict__ t_rmp, const double *__restrict__ t_rmm,
					     const double *__restrict__ t_amp, const double *__restrict__ t_amm,
					     const double *__restrict__ t_e, const double *__restrict__ t_sigma,
					     const double *__restrict__ vp_old)
{
    int cn = (jout - nsh) % ns;
    if (cn < 0)
	cn += ns;
    const int cl = (cn + AZ_NC - 1 >= ns) ? cn + AZ_NC - 1 - ns : cn + AZ_NC - 1;
    const size_t a0 = row + (size_t)cn, a1 = row + (size_t)cl;
    pf_global(t_rmp + a0), pf_global(t_rmp + a1);
    pf_global(t_rmm + a0), pf_global(t_rmm + a1);
    pf_global(t_amp + a0), pf_global(t_amp + a1);
    pf_global(t_amm + a0), pf_global(t_amm + a1);
    pf_global(t_sigma + a0), pf_global(t_sigma + a1);
    pf_global(vp_old + a0), pf_global(vp_old + a1);
    if (ADI)
	pf_global(t_e + a0), pf_global(t_e + a1);
}

// The same segment through shared memory (AZ_DEPTH > 0): the copies of ring i + 1 are issued while ring i is computed
// (fargo_dev.h:cp_async_16), 16 bytes per quantity when the ring's shift keeps the thread's pair of columns aligned, 8 bytes
// per column otherwise (warp-uniform: every lane starts at an even output column).  Slot layout [quantity][thread][column].
#ifndef AZ_DEPTH
#define AZ_DEPTH 1
#endif
typedef double AzSlot[7][128][AZ_NC];
template <bool ADI>
__device__ __forceinline__ void az_stage(AzSlot &dst, const int ns, const int jout, const int nsh, const size_t row,
					  const double *__restrict__ t_rmp, const double *__restrict__ t_rmm,
					  const double *__restrict__ t_amp, const double *__restrict__ t_amm,
					  const double *__restrict__ t_e, const double *__restrict__ t_sigma,
					  const double *__restrict__ vp_old)
{
    int col = (jout - nsh) % ns; // pre-shift column of c = 0
    if (col < 0)
	col += ns;
    const int t = threadIdx.x;
    if (AZ_NC == 2 && ((col | ns) & 1) == 0) {
	const size_t a = row + (size_t)col;
	cp_async_16(dst[0][t], t_rmp + a);
	cp_async_16(dst[1][t], t_rmm + a);
	cp_async_16(dst[2][t], t_amp + a);
	cp_async_16(dst[3][t], t_amm + a);
	if (ADI)
	    cp_async_16(dst[4][t], t_e + a);
	cp_async_16(dst[5][t], t_sigma + a);
	cp_async_16(dst[6][t], vp_old + a);
    } else {
#pragma unroll
	for (int k = 0; k < AZ_NC; ++k) {
	    const size_t a = row + (size_t)col;
	    cp_async_8(&dst[0][t][k], t_rmp + a);
	    cp_async_8(&dst[1][t][k], t_rmm + a);
	    cp_async_8(&dst[2][t][k], t_amp + a);
	    cp_async_8(&dst[3][t][k], t_amm + a);
	    if (ADI)
		cp_async_8(&dst[4][t][k], t_e + a);
	    cp_async_8(&dst[5][t][k], t_sigma + a);
	    cp_async_8(&dst[6][t][k], vp_old + a);
	    col = (col + 1 == ns) ? 0 : col + 1;
	}
    }
}
template <bool ADI> __device__ __forceinline__ void az_take(const AzSlot &src, AzIn &N)
{
    const int t = threadIdx.x;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
	if (q == 4 && !ADI) {
#pragma unroll
	    for (int k = 0; k < AZ_NC; ++k)
		N.Q[q][k] = 0.0;
	    continue;
	}
	if (AZ_NC == 2) {
	    const double2 a = *reinterpret_cast<const double2 *>(src[q][t]);
	    N.Q[q][0] = a.x, N.Q[q][AZ_NC - 1] = a.y;
	} else {
#pragma unroll
	    for (int k = 0; k < AZ_NC; ++k)
		N.Q[q][k] = src[q][t][k];
	}
    }
    if (AZ_NC == 2) {
	const double2 a = *reinterpret_cast<const double2 *>(src[6][t]);
	N.VP[0] = a.x, N.VP[AZ_NC - 1] = a.y;
    } else {
#pragma unroll
	for (int k = 0; k < AZ_NC; ++k)
	    N.VP[k] = src[6][t][k];
    }
}

// Everything the kernel does with one ring segment once it is in registers: the residual-velocity pass, the uniform
// pass, velocities from momenta (:498-535) and the floors (:123-131).  FAST = true is the hot path: straight-line,
// branch-free arithmetic whose validity is accumulated in A and tested ONCE per ring by the caller; FAST = false is the
// same code on the plain operators (the caller reruns the ring through it, warp-wide, when any lane's test fails —
// bit-identical for the lanes that were valid, so the result does not depend on who triggered the rerun).
template <int LIM, bool ADI, bool FAST>
__device__ __forceinline__ void az_ring(const DevView &c, const TempClampNB &tc, const AzIn &IN, const AzRing &g, const double dt,
					 const double vm, const double vc, const bool fargo, const int i, const bool lane_out,
					 const double (&PS)[AZ_NC], const double (&PR)[AZ_NC], double (&Q)[6][AZ_NC],
					 double (&vrn)[AZ_NC], double (&vpn)[AZ_NC], double (&sf)[AZ_NC], double (&en)[AZ_NC], FmAcc &A)
{
    double U[AZ_NC];
#pragma unroll
    for (int k = 0; k < AZ_NC; ++k) {
#pragma unroll
	for (int q = 0; q < 6; ++q)
	    Q[q][k] = IN.Q[q][k];
	double u = IN.VP[k] - vm; // compute_residual_velocity :194-205
	if (!fargo)
	    u = vc + u; // ComputeConstantResidual :225-231
	U[k] = u;
    }
    // pass 1: residual velocity; pass 2: constant residual velocity (skipped for standard transport, :646)
    // (two copies of the pass code: with the cold paths gone they fit the instruction cache, and the compiler
    // schedules across the pass boundary: 5.10 -> 4.97 ms at 8192x16384)
    az_pass<LIM, ADI, FAST>(Q, U, g, dt, A);
    if (fargo) {
#pragma unroll
	for (int k = 0; k < AZ_NC; ++k)
	    U[k] = vc;
	az_pass<LIM, ADI, FAST>(Q, U, g, dt, A);
    }
    // velocities from momenta (:498-535), floors (:123-131)
    const double rmed = c.g.rmed[i], invrmed = c.g.invrmed[i];
    const double OmegaF = c.b.omega_frame;
    const double floorv = c.p.sigma_floor * c.p.sigma0;
    const double am_left = shfl_from_left(Q[2][AZ_NC - 1]);
    const double s_left = shfl_from_left(Q[5][AZ_NC - 1]);
    FmAcc B; // keys of this stage count only where the lane's columns are stored
#pragma unroll
    for (int k = 0; k < AZ_NC; ++k) {
	const double s = Q[5][k];
	const double sm = (k == 0) ? s_left : Q[5][(k + AZ_NC - 1) % AZ_NC];
	const double amp_m = (k == 0) ? am_left : Q[2][(k + AZ_NC - 1) % AZ_NC];
	const double nvr = PR[k] + Q[1][k];
	const double dvr = PS[k] + s;
	const double nvp = amp_m + Q[3][k];
	const double dvp = sm + s;
	double qr, qp;
	if (FAST) {
	    qr = fm_div_raw(nvr, dvr, fm_rcp_raw(dvr));
	    qp = fm_div_raw(nvp, dvp, fm_rcp_raw(dvp));
	    fm_acc_nrm_if(B, i != 0, dvr);
	    fm_acc_nrm_if(B, i != 0, qr);
	    fm_acc_nrm(B, dvp);
	    fm_acc_nrm(B, qp);
	} else {
	    qr = (i == 0) ? 0.0 : nvr / dvr;
	    qp = nvp / dvp;
	}
	vrn[k] = (i == 0) ? 0.0 : qr;
	vpn[k] = qp * invrmed - rmed * OmegaF;
	sf[k] = (s < floorv) ? floorv : s;
	if (ADI)
	    en[k] = FAST ? temperature_clamp_nb(tc, sf[k], Q[4][k], B) : temperature_clamp(c, sf[k], Q[4][k]);
	else
	    en[k] = 0.0;
    }
    if (FAST && lane_out) {
	A.m = max(A.m, B.m);
	A.ms = max(A.ms, B.ms);
    }
}

// Which (window, ring) work a launch covers.  One CTA = four neighbouring windows (one per warp, so their overlapping
// columns meet in L1) marched through one band of rings.
// Segment 2 is the bulk of the slab: ceil(nwin / 4) window groups x bands of `band` rings, band-major in the block index,
// so the CTAs running at any moment sweep whole ring rows together (DRAM sees long sequential reads) and the block
// scheduler balances the bands over the SMs (rings_per_march picks `band`).  (Cutting the groups' ring ranges into one
// equal run per resident CTA instead was measured: 10 % slower at 8192 rings — SMs do not all run at the same speed —
// and a loop over the two marches such a run can consist of cost the single-march kernel 4.5 %.)
// Segments 0 and 1 exist on the peer-memory halo path (PUSH): the 2 x CPUOVERLAP rings at the slab's inner / outer edge,
// one march per window, scheduled FIRST (lowest block indices).  Their epilogue stores the finished rings the neighbour
// needs not only into this GPU's fields but also straight into the neighbouring GPU's halo inbox (peer memory mapped
// over NVLink), and the last edge warp to finish publishes the step number in the neighbours' arrival counters — so the
// ghost-ring exchange (CommunicateBoundaries, commbound.cpp:98-182) is part of this kernel and is over long before the
// interior bands are.
struct AzSegs {
    int lo[3], hi[3];
    int nwin;	     // windows per ring
    int band;	     // rings per band of segment 2
    int n_edge;	     // edge segments in this launch (0, 1 or 2); their CTAs come first: n_edge * ceil(nwin / 4)
    int edge_seg[2]; // which segment the e-th group of edge CTAs runs
    int push_lo[2];  // rings [push_lo[s], push_lo[s] + CPUOVERLAP) of segment s are mirrored to push[s][field]
    double *push[2][4]; // Sigma, v_rad, v_azi, e
    unsigned long long *peer_flag[2]; // the neighbour's arrival counter for rings coming from this rank
    unsigned long long seq;	      // value to publish
    unsigned int *done;		      // edge warps finished (local device memory, returns to 0)
    unsigned int expected;	      // edge warps in this launch
    // Damping zones folded into the epilogue (damping.cpp:311-752; boundary_conditions.cpp:65-114 applies them right after
    // Transport + CommunicateBoundaries in an Euler step): dmask == nullptr: not folded (k_damping runs as its own pass).
    const int *dmask;	  // per ring (nr + 1 entries): 2 bits per field f = 0 v_rad, 1 v_azi, 2 Sigma, 3 e: 0 none, 1 towards the
			  // initial field dx0[f], 2 towards the constant dx0c[f]
    const double *dexpf;  // 4 tables (stride dstride) of this step's exp(-dt * factor / tau) per ring, formed on the host with glibc
    int dstride;
    const double *dx0[4];
    double dx0c[4];
};

// one field of one ring through the damping formula of damping.cpp (:336-343 and its siblings): X = (X - X0) * e + X0
// x0s: the thread's columns of the initial field, staged through shared memory by az_stage_damp (nullptr: plain loads)
__device__ __forceinline__ void az_damp_field(const AzSegs &segs, const int f, const int type, const int i, const size_t a, const bool ld,
					       const int ncols_in, double (&X)[AZ_NC], const double *x0s = nullptr)
{
    const double ef = segs.dexpf[(size_t)f * segs.dstride + i];
#pragma unroll
    for (int k = 0; k < AZ_NC; ++k) {
	double X0 = segs.dx0c[f];
	if (type == 1 && ld && k < ncols_in)
	    X0 = x0s ? x0s[k] : segs.dx0[f][a + k];
	X[k] = (X[k] - X0) * ef + X0;
    }
}
#if AZ_DEPTH > 0
// the initial-field columns the damping of ring i will read, into the slot whose inputs have just been taken (fields 0..3)
template <bool ADI>
__device__ __forceinline__ void az_stage_damp(AzSlot &dst, const AzSegs &segs, const int dm, const size_t a, const bool ld, const int ncols_in)
{
    if (!ld)
	return;
    const int t = threadIdx.x;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
	if (f == 3 && !ADI)
	    continue;
	if (((dm >> (2 * f)) & 3) != 1)
	    continue;
#pragma unroll
	for (int k = 0; k < AZ_NC; ++k)
	    if (k < ncols_in)
		cp_async_8(&dst[f][t][k], segs.dx0[f] + a + k);
    }
}
#endif

template <int LIM, bool ADI, bool PUSH>
__global__ void __launch_bounds__(128, AZ_MINB)
    k_transport_azimuthal(const DevView c, const double *__restrict__ t_sigma, const double *__restrict__ t_rmp,
			  const double *__restrict__ t_rmm, const double *__restrict__ t_amp,
			  const double *__restrict__ t_amm, const double *__restrict__ t_e, const double *__restrict__ vp_old,
			  const double *__restrict__ vr_old, const double *__restrict__ vmean, const int *__restrict__ nshift,
			  const double *__restrict__ vconst, double *__restrict__ o_sigma, double *__restrict__ o_vr,
			  double *__restrict__ o_vp, double *__restrict__ o_e, const double dt, const AzSegs segs)
{
    const int ns = c.ns, nr = c.nr;
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5; // warp in block; warps never synchronise with each other
    const int gx = (segs.nwin + 3) >> 2;
    const int t0 = AZ_NC * lane; // local column of c = 0
    const bool lane_out = (t0 >= AZ_HL) && (t0 < AZ_WIN - AZ_HR);
    // this block's work: an edge march (PUSH) or one (band, window group) cell of segment 2
    int seg = 2, grp, i_first, i_last;
    const int edge_ctas = PUSH ? segs.n_edge * gx : 0;
    if (PUSH && (int)blockIdx.x < edge_ctas) {
	seg = segs.edge_seg[blockIdx.x / gx];
	grp = (int)blockIdx.x % gx;
	i_first = segs.lo[seg];
	i_last = segs.hi[seg];
    } else {
	const int cell = (int)blockIdx.x - edge_ctas;
	const int b = cell / gx;
	grp = cell - b * gx;
	i_first = segs.lo[2] + b * segs.band;
	i_last = min(i_first + segs.band, segs.hi[2]);
    }
    const int win = grp * 4 + wib;
    if (win >= segs.nwin || i_first >= i_last)
	return; // whole warp (the last group of four may be incomplete)
    const int jout = win * AZ_OUT - AZ_HL + t0; // output column of c = 0 (negative / >= ns in the halo)
    const bool vec_ok = ((ns % AZ_NC) == 0);
    const bool fargo = c.p.fast_transport != 0;
    TempClampNB tc;
    if (ADI)
	tc = make_temp_clamp_nb(c);

    double PS[AZ_NC], PR[AZ_NC]; // previous ring: transported Sigma, rm+
#pragma unroll
    for (int k = 0; k < AZ_NC; ++k)
	PS[k] = PR[k] = 0.0;

#if AZ_DEPTH > 0
    __shared__ AzSlot az_stg[AZ_DEPTH + 1];
    {
	const int ib = max(i_first - 1, 0);
#pragma unroll
	for (int d = 0; d < AZ_DEPTH; ++d) {
	    if (ib + d < i_last)
		az_stage<ADI>(az_stg[d], ns, jout, nshift[ib + d], (size_t)(ib + d) * ns, t_rmp, t_rmm, t_amp, t_amm, t_e, t_sigma, vp_old);
	    cp_async_commit();
	}
    }
    int slot = 0, slot_in = AZ_DEPTH;
#endif
    for (int i = max(i_first - 1, 0); i < i_last; ++i) {
	const int nsh = nshift[i];
	const double vm = vmean[i], vc = vconst[i];
	AzRing g;
	g.dxtheta = c.dphi * c.g.rmed[i];
	g.invdxtheta = c.g.invdxtheta[i]; // 1.0 / dxtheta, formed on the host with the same IEEE division
	g.dxrad = (c.g.rsup[i] - c.g.rinf[i]) * dt;
	g.invsurf = c.g.invsurf[i];
	const size_t row = (size_t)i * ns;
#if AZ_DEPTH > 0
	if (i + AZ_DEPTH < i_last)
	    az_stage<ADI>(az_stg[slot_in], ns, jout, nshift[i + AZ_DEPTH], row + (size_t)AZ_DEPTH * ns, t_rmp, t_rmm, t_amp, t_amm, t_e,
			  t_sigma, vp_old);
	cp_async_commit();
	slot_in = (slot_in == AZ_DEPTH) ? 0 : slot_in + 1;
	// damping of this ring (warp-uniform), decided before the inputs are taken so that the initial-field columns it reads
	// can travel into the just-emptied slot while the ring is computed
	const int dm = (segs.dmask != nullptr && i >= i_first) ? segs.dmask[i] : 0;
	const bool dld = lane_out && jout >= 0 && jout < ns;
	const int dnin = ns - jout; // columns of this lane inside the ring (>= AZ_NC except at a ragged end)
	const size_t da = row + (size_t)(dld ? jout : 0);
#else
	if (i + 1 < i_last)
	    az_prefetch<ADI>(ns, jout, nshift[i + 1], row + (size_t)ns, t_rmp, t_rmm, t_amp, t_amm, t_e, t_sigma, vp_old);
#endif
	double Q[6][AZ_NC], vrn[AZ_NC], vpn[AZ_NC], sf[AZ_NC], en[AZ_NC];
	FmAcc A;
	{
	    AzIn IN;
#if AZ_DEPTH > 0
	    cp_async_wait<AZ_DEPTH>(); // ring i has landed
	    az_take<ADI>(az_stg[slot], IN);
	    if (dm != 0) {
		az_stage_damp<ADI>(az_stg[slot], segs, dm, da, dld, dnin);
		cp_async_commit();
	    }
#else
	    az_fetch<ADI>(IN, ns, jout, nsh, row, t_rmp, t_rmm, t_amp, t_amm, t_e, t_sigma, vp_old);
#endif
	    az_ring<LIM, ADI, true>(c, tc, IN, g, dt, vm, vc, fargo, i, lane_out, PS, PR, Q, vrn, vpn, sf, en, A);
	}
	if (__any_sync(0xffffffffu, !fm_acc_ok(A))) { // cold, warp-wide (the ring body shuffles): exact zeros, extreme exponents
	    AzIn IN;
	    az_fetch<ADI>(IN, ns, jout, nsh, row, t_rmp, t_rmm, t_amp, t_amm, t_e, t_sigma, vp_old);
	    az_ring<LIM, ADI, false>(c, tc, IN, g, dt, vm, vc, fargo, i, lane_out, PS, PR, Q, vrn, vpn, sf, en, A);
	}
	if (ADI && c.pv.geff != nullptr && lane_out) { // PVTE (uniform): the temperature floor / ceiling with the cell's mu, gamma_eff
#pragma unroll
	    for (int k = 0; k < AZ_NC; ++k)
		if (jout + k >= 0 && jout + k < ns)
		    en[k] = temperature_clamp_at(c, row + (size_t)(jout + k), sf[k], Q[4][k]);
	}
#if AZ_DEPTH > 0
	if (dm != 0) { // warp-uniform: this ring lies in a damping zone of some field
	    cp_async_wait<0>();
	    const double(*x0)[128][AZ_NC] = az_stg[slot];
	    const int t = threadIdx.x;
	    if (dm & 3)
		az_damp_field(segs, 0, dm & 3, i, da, dld, dnin, vrn, x0[0][t]);
	    if ((dm >> 2) & 3)
		az_damp_field(segs, 1, (dm >> 2) & 3, i, da, dld, dnin, vpn, x0[1][t]);
	    if ((dm >> 4) & 3)
		az_damp_field(segs, 2, (dm >> 4) & 3, i, da, dld, dnin, sf, x0[2][t]);
	    if (ADI && ((dm >> 6) & 3))
		az_damp_field(segs, 3, (dm >> 6) & 3, i, da, dld, dnin, en, x0[3][t]);
	}
	slot = (slot == AZ_DEPTH) ? 0 : slot + 1;
#else
	if (segs.dmask != nullptr && i >= i_first) { // warp-uniform: this ring lies in a damping zone of some field
	    const int dm = segs.dmask[i];
	    if (dm != 0) {
		const bool ld = lane_out && jout >= 0 && jout < ns;
		const int nin = ns - jout; // columns of this lane inside the ring (>= AZ_NC except at a ragged end)
		const size_t a = row + (size_t)(ld ? jout : 0);
		if (dm & 3)
		    az_damp_field(segs, 0, dm & 3, i, a, ld, nin, vrn);
		if ((dm >> 2) & 3)
		    az_damp_field(segs, 1, (dm >> 2) & 3, i, a, ld, nin, vpn);
		if ((dm >> 4) & 3)
		    az_damp_field(segs, 2, (dm >> 4) & 3, i, a, ld, nin, sf);
		if (ADI && ((dm >> 6) & 3))
		    az_damp_field(segs, 3, (dm >> 6) & 3, i, a, ld, nin, en);
	    }
	}
#endif
	if (PUSH && seg < 2 && i >= i_first && lane_out) { // edge ring of the slab: mirror it into the neighbour's halo inbox
	    const int pr = i - segs.push_lo[seg];
	    if (pr >= 0 && pr < FARGO_CPUOVERLAP) {
#pragma unroll
		for (int k = 0; k < AZ_NC; ++k) {
		    if (jout + k < ns) {
			const size_t a = (size_t)pr * ns + (size_t)(jout + k);
			segs.push[seg][0][a] = sf[k];
			segs.push[seg][1][a] = vrn[k];
			segs.push[seg][2][a] = vpn[k];
			if (ADI)
			    segs.push[seg][3][a] = en[k];
		    }
		}
	    }
	}
	if (i >= i_first && lane_out) {
	    if (vec_ok) {
		if (jout < ns) { // jout is a multiple of AZ_NC, so all the columns are inside
		    const size_t a = row + (size_t)jout;
#pragma unroll
		    for (int k = 0; k < AZ_NC; k += 2) {
			*reinterpret_cast<double2 *>(o_vr + a + k) = make_double2(vrn[k], vrn[k + 1]);
			*reinterpret_cast<double2 *>(o_vp + a + k) = make_double2(vpn[k], vpn[k + 1]);
			*reinterpret_cast<double2 *>(o_sigma + a + k) = make_double2(sf[k], sf[k + 1]);
			if (ADI)
			    *reinterpret_cast<double2 *>(o_e + a + k) = make_double2(en[k], en[k + 1]);
		    }
		}
	    } else {
#pragma unroll
		for (int k = 0; k < AZ_NC; ++k) {
		    if (jout + k < ns) {
			const size_t a = row + (size_t)(jout + k);
			o_vr[a] = vrn[k];
			o_vp[a] = vpn[k];
			o_sigma[a] = sf[k];
			if (ADI)
			    o_e[a] = en[k];
		    }
		}
	    }
	}
#pragma unroll
	for (int k = 0; k < AZ_NC; ++k) {
	    PS[k] = Q[5][k];
	    PR[k] = Q[0][k];
	}
    }
#if AZ_DEPTH > 0
    cp_async_wait<0>();
#endif
    // v_rad ring nr is not touched by compute_velocities_from_momenta (:502-507): carry it over
    if (i_last == nr && lane_out) {
	const int dmv = segs.dmask != nullptr ? (segs.dmask[nr] & 3) : 0; // v_rad's outermost interface is damped like any other
#pragma unroll
	for (int k = 0; k < AZ_NC; ++k)
	    if (jout + k < ns) {
		const size_t a = (size_t)nr * ns + jout + k;
		double X = vr_old[a];
		if (dmv != 0) {
		    const double X0 = dmv == 1 ? segs.dx0[0][a] : segs.dx0c[0];
		    X = (X - X0) * segs.dexpf[nr] + X0;
		}
		o_vr[a] = X;
	    }
    }
    if (PUSH && seg < 2) {
	// every lane orders its peer stores before what follows, system-wide; the warp's lane 0 then counts the warp in, and
	// the last edge warp of the launch publishes the step in the neighbours' arrival counters (k_halo_unpack spins on them)
	__threadfence_system();
	__syncwarp();
	if (lane == 0) {
	    const unsigned int before = atomicAdd(segs.done, 1u);
	    if (before == segs.expected - 1u) {
		*segs.done = 0u; // next launch starts from 0 (every other edge warp has already counted itself in)
		__threadfence_system();
		if (segs.peer_flag[0])
		    *(volatile unsigned long long *)segs.peer_flag[0] = segs.seq;
		if (segs.peer_flag[1])
		    *(volatile unsigned long long *)segs.peer_flag[1] = segs.seq;
		__threadfence_system();
	    }
	}
    }
}
