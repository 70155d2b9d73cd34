// kernels_fused.cuh — the source-term half of the hydro step as three ring-marching kernels.
//
//   k_fused_sources    CalculateNbodyPotential (Pframeforce.cpp:21-86) + momentum_update_radial / _azimuthal
//                      (SourceEuler.cpp:325-428) + compression_heating (:459-493)
//   k_fused_artvisc    update_with_artificial_viscosity TW / SN (viscosity/artificial_viscosity.cpp:11-250)
//   k_fused_viscosity  recalculate_viscosity (SourceEuler.cpp:205-223) + compute_viscous_stress_tensor +
//                      update_velocities_with_viscosity (viscosity/viscosity.cpp:139-426) + SubStep3
//                      (SourceEuler.cpp:496-954: viscous heating, beta cooling, energy update, T floor)
//
// Same execution shape as the azimuthal transport kernel: a warp owns a window of 32 * FS_NC columns, each lane FS_NC
// consecutive ones (azimuthal neighbours are the thread's own registers or one shuffle away) and marches outward in
// radius; the rings i-1 / i-2 a stage needs are the registers of the previous iterations, so every state array is
// read ONCE and every intermediate (Phi, P, Q_rr, Q_phiphi, nu, div v, tau_rr, tau_phiphi, tau_rphi) lives only in
// registers.  A stage that needs column j+-1 of the previous stage's output makes the outermost columns of the window
// invalid, so a window of 64 columns yields 60 finished ones ([2, 62)); warps do not communicate.  Outputs go to the
// OTHER buffer of each double-buffered field (windows overlap on reads).
// FS_NC = 2 at 4 CTAs / SM (128 registers) beats 4 columns per lane at 3 CTAs / SM (168 registers, spills) by 17 % on
// the viscosity and source kernels: the FP64 chains need warps, not wider threads, to hide their latency.
//
// Compared with one kernel per reference loop nest (the staged kernels in kernels_source.cuh, kept for the
// per-stage entry points) this reads/writes 184 B per cell instead of 472 B and executes ~3x fewer instructions.
#pragma once
#include "fargo_dev.h"
#include "kernels_azimuthal.cuh"
#include "kernels_source.cuh"

#ifndef FS_NC
#define FS_NC 2 // columns per lane
#endif
#define FS_WIN (32 * FS_NC)
// Invalid columns at either end of a window.  Every stage of the three kernels reads column j +- 1 of RAW inputs (valid in
// every lane) or of the previous stage's output, and no stage chains two such reads, so exactly the first and the last column
// of a window are wrong (lane 0's "left" and lane 31's "right" shuffle); rounded up to whole lanes (16-byte vector stores).
#ifndef FS_HL
#define FS_HL 2
#endif
#ifndef FS_HR
#define FS_HR 2
#endif
#define FS_OUT (FS_WIN - FS_HL - FS_HR)

// window bookkeeping shared by the three kernels
struct FsLane {
    int jout;	  // unwrapped output column of the thread's first column
    int col;	  // wrapped column (0 <= col < ns) of the first column
    bool vec;	  // the 4 columns are contiguous and 32-byte aligned (ns % 4 == 0)
    bool lane_out; // this lane's columns are inside the valid part of the window and inside the ring
};
__device__ __forceinline__ bool fs_setup(const DevView &c, FsLane &L)
{
    const int lane = threadIdx.x & 31;
    const int win = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if ((long long)win * FS_OUT >= c.ns)
	return false;
    const int t0 = FS_NC * lane;
    L.jout = win * FS_OUT - FS_HL + t0;
    int col = L.jout % c.ns;
    if (col < 0)
	col += c.ns;
    L.col = col;
    L.vec = ((c.ns % FS_NC) == 0);
    L.lane_out = (t0 >= FS_HL) && (t0 < FS_WIN - FS_HR) && (L.jout < c.ns);
    return true;
}
__device__ __forceinline__ void fs_load(const double *__restrict__ arr, const int ring, const DevView &c, const FsLane &L,
					 double (&x)[FS_NC])
{
    const double *row = arr + (size_t)ring * c.ns;
    if (L.vec) {
#pragma unroll
	for (int k = 0; k < FS_NC; k += 2) {
	    const double2 a = *reinterpret_cast<const double2 *>(row + L.col + k);
	    x[k] = a.x;
	    x[k + 1] = a.y;
	}
    } else {
	int cc = L.col;
#pragma unroll
	for (int k = 0; k < FS_NC; ++k) {
	    x[k] = row[cc];
	    cc = (cc + 1 == c.ns) ? 0 : cc + 1;
	}
    }
}
__device__ __forceinline__ void fs_store(double *__restrict__ arr, const int ring, const DevView &c, const FsLane &L,
					  const double (&x)[FS_NC])
{
    if (!L.lane_out)
	return;
    double *row = arr + (size_t)ring * c.ns;
    if (L.vec) {
#pragma unroll
	for (int k = 0; k < FS_NC; k += 2)
	    *reinterpret_cast<double2 *>(row + L.jout + k) = make_double2(x[k], x[k + 1]);
    } else {
#pragma unroll
	for (int k = 0; k < FS_NC; ++k)
	    if (L.jout + k < c.ns)
		row[L.jout + k] = x[k];
    }
}
#define FS_FOR4 _Pragma("unroll") for (int k = 0; k < FS_NC; ++k)
// software prefetch of the next ring's row segment (the marching loops touch every row exactly once, so the
// hardware sees no reuse to exploit; without this the first consumer of each row eats the full DRAM latency)
__device__ __forceinline__ void fs_prefetch(const double *__restrict__ arr, const int ring, const DevView &c, const FsLane &L)
{
    pf_global(arr + (size_t)ring * c.ns + L.col);
}
// One ring of the four state fields as the fused kernels consume it: v_rad has nr + 1 rings, the cell-centred fields
// nr (ring nr reads as Sigma = e = 1, v_azi = 0: harmless operands for the stages that run on it).
// (Fetching ring kr + 1 into registers while ring kr is computed was measured and is slower on B200: the extra live
// registers cost more than the load latency they hide; 25.6 vs 23.3 ms/step.  The compiler hoists the loads of ring kr
// to the top of the iteration and the L2 prefetch of ring kr + 1 covers the DRAM latency.)
struct FsRing {
    double S[FS_NC], E[FS_NC], VP[FS_NC], VR[FS_NC];
};
template <bool ADI>
__device__ __forceinline__ void fs_fetch_ring(const DevView &c, const FsLane &L, const double *__restrict__ sigma,
					       const double *__restrict__ energy, const double *__restrict__ vr,
					       const double *__restrict__ vp, const int kr, FsRing &N)
{
    fs_load(vr, kr, c, L, N.VR);
    if (kr < c.nr) {
	fs_load(sigma, kr, c, L, N.S);
	fs_load(vp, kr, c, L, N.VP);
	if (ADI)
	    fs_load(energy, kr, c, L, N.E);
	else
	    FS_FOR4 N.E[k] = 0.0;
    } else {
	FS_FOR4 { N.S[k] = 1.0, N.E[k] = 1.0, N.VP[k] = 0.0; }
    }
}
template <bool ADI>
__device__ __forceinline__ void fs_prefetch_ring(const DevView &c, const FsLane &L, const double *__restrict__ sigma,
						  const double *__restrict__ energy, const double *__restrict__ vr,
						  const double *__restrict__ vp, const int kr)
{
    fs_prefetch(vr, kr, c, L);
    if (kr < c.nr) {
	fs_prefetch(sigma, kr, c, L);
	fs_prefetch(vp, kr, c, L);
	if (ADI)
	    fs_prefetch(energy, kr, c, L);
    }
}
// top of a marching iteration: ring kr into R, L2 prefetch of ring kr + 1
template <bool ADI>
__device__ __forceinline__ void fs_next_ring(const DevView &c, const FsLane &L, const double *__restrict__ sigma,
					      const double *__restrict__ energy, const double *__restrict__ vr,
					      const double *__restrict__ vp, const int kr, const int i_last, FsRing &R)
{
    fs_fetch_ring<ADI>(c, L, sigma, energy, vr, vp, kr, R);
    if (kr < i_last)
	fs_prefetch_ring<ADI>(c, L, sigma, energy, vr, vp, kr + 1);
}
#ifndef FS_MINB_SRC
#define FS_MINB_SRC 4
#endif
#ifndef FS_MINB_AV
#define FS_MINB_AV 4
#endif
#ifndef FS_MINB_VISC
#define FS_MINB_VISC 4
#endif

// ---------------------------------------------------------------------------------------------
// Stage bodies.  Each is written once against the arithmetic policy M (fargo_math.h): MathP<true> emits the
// divisions / square roots / exponentials as straight-line fast paths and accumulates their validity in A,
// MathP<false> uses the plain operators.  A kernel runs stage<MathP<true>> and, only if the accumulator failed
// (exponents far outside anything physical, exact zeros as numerators), redoes the stage with MathP<false> from
// the same inputs — outputs never alias inputs.  The hot path has no branch inside a stage, so the instruction
// streams of the 4 columns (and of the bodies / quantities) interleave and hide the 12-cycle DFMA latency.
#define FS_RUN(stage_call_fast, stage_call_slow) \
    {                                            \
	FmAcc A_;                                \
	{                                        \
	    FmAcc &A = A_;                       \
	    stage_call_fast;                     \
	}                                        \
	if (!fm_acc_ok(A_)) {                    \
	    FmAcc &A = A_;                       \
	    stage_call_slow;                     \
	}                                        \
    }

// CalculateNbodyPotential (Pframeforce.cpp:44-85) + pressure, ring kr
template <class M, bool ADI>
__device__ __forceinline__ void st_potential(const DevView &c, const EosC &ec, const int kr, const double (&S0)[FS_NC],
					      const double (&E0)[FS_NC], const double (&cosj)[FS_NC], const double (&sinj)[FS_NC],
					      const bool have_h, const double (&Hin)[FS_NC], double (&P0)[FS_NC], double (&F0)[FS_NC], FmAcc &A)
{
    const double rmed = c.g.rmed[kr];
    double smooth[FS_NC], x[FS_NC], y[FS_NC], pot[FS_NC];
    FS_FOR4
    {
	P0[k] = eos_P(c, kr, S0[k], E0[k]);
	double H;
	if (have_h) { // leapfrog, second kick: the scale height stored by the first kick's viscosity stage
	    H = Hin[k];
	} else {
	    const double cs = eos_cs_m<M>(c, kr, S0[k], E0[k], A);
	    H = eos_H_m<M>(c, ec, kr, cs, A);
	}
	x[k] = rmed * cosj[k];
	y[k] = rmed * sinj[k];
	smooth[k] = c.p.thickness_smoothing * H;
	pot[k] = 0.0;
    }
    for (int b = 0; b < c.b.n; ++b) {
	const double bx = c.b.x[b], by = c.b.y[b], gm = -c.p.G * c.b.mass[b], r_sm = c.b.cubic_smoothing_radius[b];
	FS_FOR4
	{
	    const double dx = x[k] - bx;
	    const double dy = y[k] - by;
	    const double dist_2 = dx * dx + dy * dy;
	    const double d_smoothed = M::sqrt(dist_2 + smooth[k] * smooth[k], A);
	    double smooth_factor_klahr = 1.0;
	    if (r_sm > 0.0 && d_smoothed < r_sm) { // rare: inside a planet's cubic smoothing radius
		const double q = d_smoothed / r_sm;
		smooth_factor_klahr = (pow(q, 4.0) - 2.0 * pow(q, 3.0) + 2.0 * d_smoothed / r_sm);
	    }
	    pot[k] += M::div(gm, d_smoothed, A) * smooth_factor_klahr;
	}
    }
    FS_FOR4
    {
	pot[k] += -c.b.indirect_x * x[k] - c.b.indirect_y * y[k];
	F0[k] = pot[k];
    }
}

// momentum_update_radial (SourceEuler.cpp:325-372): interface kr between rings kr-1 and kr
template <class M>
__device__ __forceinline__ void st_vrad(const DevView &c, const int kr, const double dt, const double (&S0)[FS_NC],
					 const double (&S1)[FS_NC], const double (&P0)[FS_NC], const double (&P1)[FS_NC],
					 const double (&F0)[FS_NC], const double (&F1)[FS_NC], const double (&VP0)[FS_NC], const double VP0r,
					 const double (&VP1)[FS_NC], const double VP1r, const double (&VR0)[FS_NC], double (&VRn0)[FS_NC], FmAcc &A)
{
    const double idr = c.g.invdiffrmed[kr], rinf = c.g.rinf[kr], invrinf = c.g.invrinf[kr];
    const double OmegaF = c.b.omega_frame;
    FS_FOR4
    {
	double gradp = M::div(2.0, S0[k] + S1[k], A);
	gradp *= (P0[k] - P1[k]);
	gradp *= idr;
	const double gradphi = (F0[k] - F1[k]) * idr;
	const double vp0n = (k == FS_NC - 1) ? VP0r : VP0[(k + 1) % FS_NC];
	const double vp1n = (k == FS_NC - 1) ? VP1r : VP1[(k + 1) % FS_NC];
	const double vsum = VP0[k] + vp0n + VP1[k] + vp1n;
	const double vt = 0.25 * vsum + rinf * OmegaF;
	const double vt2 = vt * vt;
	const double centrifugal_accel = vt2 * invrinf;
	VRn0[k] = VR0[k] + dt * (-gradp - gradphi + centrifugal_accel);
    }
}

// momentum_update_azimuthal (:375-428), ring kr
template <class M>
__device__ __forceinline__ void st_vazi(const DevView &c, const int kr, const double dt, const bool drift, const double (&S0)[FS_NC],
					 const double Sl, const double (&P0)[FS_NC], const double Pl, const double (&F0)[FS_NC],
					 const double Fl, const double (&VP0)[FS_NC], double (&VPn0)[FS_NC], FmAcc &A)
{
    const double invdxtheta = c.g.invdxtheta_mid[kr];
    const double supp = drift ? c.g.supp_torque[kr] : 0.0;
    FS_FOR4
    {
	const double sp = (k == 0) ? Sl : S0[(k + FS_NC - 1) % FS_NC];
	const double Pp = (k == 0) ? Pl : P0[(k + FS_NC - 1) % FS_NC];
	const double Fp = (k == 0) ? Fl : F0[(k + FS_NC - 1) % FS_NC];
	const double gradp = M::div(2.0, S0[k] + sp, A) * (P0[k] - Pp) * invdxtheta;
	const double gradphi = (F0[k] - Fp) * invdxtheta;
	double vpn = VP0[k] + dt * (-gradp - gradphi);
	if (drift)
	    vpn += dt * supp;
	VPn0[k] = vpn;
    }
}

// compression_heating (:459-493), ring r, with the UPDATED velocities
template <class M>
__device__ __forceinline__ void st_compress(const DevView &c, const int r, const double dt, const double (&E1)[FS_NC],
					     const double (&VRn0)[FS_NC], const double (&VRn1)[FS_NC], const double (&VPn1)[FS_NC],
					     const double VPn1r, double (&En)[FS_NC], FmAcc &A)
{
    const double ra1 = c.g.rinf[r + 1], ra0 = c.g.rinf[r], idrb = c.g.invdiffrsuprb[r], irb = c.g.invrmed[r];
    FS_FOR4
    {
	const double vpn = (k == FS_NC - 1) ? VPn1r : VPn1[(k + 1) % FS_NC];
	const double DIV_V = (VRn0[k] * ra1 - VRn1[k] * ra0) * idrb + (vpn - VPn1[k]) * c.invdphi * irb;
	En[k] = E1[k] * M::exp(-(c.p.gamma - 1.0) * dt * DIV_V, A);
    }
}

// ---------------------------------------------------------------------------------------------
// artificial viscosity stage bodies (viscosity/artificial_viscosity.cpp)
struct AvIn {
    // ring r: Sigma, e, v_rad(r), v_rad(r+1), v_azi (+ right neighbour), and ring r-1: Sigma, Q
    double S1[FS_NC], E1[FS_NC], VR1[FS_NC], VR0[FS_NC], VP1[FS_NC], VP1r, S2[FS_NC], QR2[FS_NC], QP2[FS_NC];
};
// Q_rr / Q_phiphi of ring r and the dissipation into e: TW :49-88, SN :165-218 (no division: exact on any input)
template <bool ADI>
__device__ __forceinline__ void st_av_q(const DevView &c, const int r, const double dt, const int type, const bool diss,
					 const AvIn &I, double (&QR1)[FS_NC], double (&QP1)[FS_NC], double (&En)[FS_NC])
{
    const double C = c.p.artificial_viscosity_factor;
    FS_FOR4
    {
	En[k] = I.E1[k];
	QR1[k] = QP1[k] = 0.0;
    }
    if (type == FARGO_ARTVISC_TW) {
	const double ids = c.g.invdiffrsup[r], irb = c.g.invrmed[r];
	const double Dr = c.g.rinf[r + 1] - c.g.rinf[r];
	const double rDphi = c.g.rmed[r] * c.dphi;
	const double m = (c.ns <= 16) ? stdmin(Dr, rDphi) : stdmax(Dr, rDphi);
	const double dx_sq = m * m;
	const double l_sq = (C * C) * dx_sq;
	const bool heat = diss && r > c.zero_no_ghost && r < c.max_no_ghost;
	FS_FOR4
	{
	    const double vpn = (k == FS_NC - 1) ? I.VP1r : I.VP1[(k + 1) % FS_NC];
	    const double eps_rr = (I.VR0[k] - I.VR1[k]) * ids;
	    const double eps_pp = irb * ((vpn - I.VP1[k]) * c.invdphi + 0.5 * (I.VR0[k] + I.VR1[k]));
	    const double div_V = stdmin(eps_rr + eps_pp, 0.0);
	    QR1[k] = l_sq * I.S1[k] * -div_V * (eps_rr - 1.0 / 3.0 * div_V);
	    QP1[k] = l_sq * I.S1[k] * -div_V * (eps_pp - 1.0 / 3.0 * div_V);
	    if (heat) {
		const double Qplus = -l_sq * div_V * I.S1[k] * 1.0 / 3.0 *
				     (eps_rr * eps_rr + eps_pp * eps_pp + (eps_rr - eps_pp) * (eps_rr - eps_pp));
		En[k] += Qplus * dt;
	    }
	}
    } else if (type == FARGO_ARTVISC_SN) {
	const bool heat = diss && r >= c.zero_no_ghost && r < c.max_no_ghost;
	const double invdxtheta = c.g.invdxtheta[r];
	FS_FOR4
	{
	    const double vpn = (k == FS_NC - 1) ? I.VP1r : I.VP1[(k + 1) % FS_NC];
	    const double dv_r = I.VR0[k] - I.VR1[k];
	    QR1[k] = (dv_r < 0.0) ? (C * C) * I.S1[k] * (dv_r * dv_r) : 0.0;
	    const double dv_phi = vpn - I.VP1[k];
	    QP1[k] = (dv_phi < 0.0) ? (C * C) * I.S1[k] * (dv_phi * dv_phi) : 0.0;
	    if (heat)
		En[k] = En[k] - dt * QR1[k] * dv_r * c.g.invdiffrsup[r] - dt * QP1[k] * dv_phi * invdxtheta;
	}
    }
}
// velocity updates of ring r from Q(r), Q(r-1): TW :90-139, SN :221-248
template <class M>
__device__ __forceinline__ void st_av_v(const DevView &c, const int r, const double dt, const int type, const AvIn &I,
					 const double (&QR1)[FS_NC], const double (&QP1)[FS_NC], const double QP1l, const double S1l,
					 double (&VRn)[FS_NC], double (&VPn)[FS_NC], FmAcc &A)
{
    const int nr = c.nr;
    FS_FOR4
    {
	VRn[k] = I.VR1[k];
	VPn[k] = I.VP1[k];
    }
    if (type == FARGO_ARTVISC_TW) {
	if (r >= 1 && r < nr - 1) {
	    const double rs = c.g.rsup[r] + c.g.rinf[r];
	    FS_FOR4
	    {
		const double sp = (k == 0) ? S1l : I.S1[(k + FS_NC - 1) % FS_NC];
		const double qpp = (k == 0) ? QP1l : QP1[(k + FS_NC - 1) % FS_NC];
		const double sigma_phi_avg = 0.5 * (I.S1[k] + sp);
		const double dVp = M::div(2.0 * dt, rs * sigma_phi_avg, A) * (QP1[k] - qpp) * c.invdphi;
		VPn[k] = I.VP1[k] + dVp;
	    }
	}
	if (r >= c.one_no_ghost_vr && r < c.maxmo_no_ghost_vr) {
	    const double rm = c.g.rmed[r], rmm = c.g.rmed[r - 1];
	    const double dr2 = rm * rm - rmm * rmm;
	    const double ydr2 = M::rcp(dr2, A);
	    FS_FOR4
	    {
		const double sigma_r_avg = 0.5 * (I.S1[k] + I.S2[k]);
		const double dVr = M::div_y(M::div(c.p.radial_viscosity_factor * dt, sigma_r_avg, A) * 2.0, dr2, ydr2, A) *
				   ((QR1[k] * rm - I.QR2[k] * rmm) - 0.5 * (QP1[k] + I.QP2[k]) * (rm - rmm));
		VRn[k] = I.VR1[k] + dVr;
	    }
	}
    } else if (type == FARGO_ARTVISC_SN) {
	if (r >= c.one_no_ghost_vr && r < c.maxmo_no_ghost_vr) {
	    const double idr = c.g.invdiffrmed[r];
	    FS_FOR4 VRn[k] = I.VR1[k] - M::div(dt * 2.0, I.S1[k] + I.S2[k], A) * (QR1[k] - I.QR2[k]) * idr;
	}
	if (r >= c.zero_no_ghost && r < c.max_no_ghost) {
	    const double invdxtheta = c.g.invdxtheta[r];
	    FS_FOR4
	    {
		const double sp = (k == 0) ? S1l : I.S1[(k + FS_NC - 1) % FS_NC];
		const double qpp = (k == 0) ? QP1l : QP1[(k + FS_NC - 1) % FS_NC];
		VPn[k] = I.VP1[k] - M::div(dt * 2.0, I.S1[k] + sp, A) * (QP1[k] - qpp) * invdxtheta;
	    }
	}
    }
}

// ---------------------------------------------------------------------------------------------
// k_fused_sources.  Iteration kr loads ring kr, forms Phi(kr), P(kr), v_rad'(kr) (needs ring kr-1) and v_azi'(kr),
// then finishes ring r = kr-1: e'(r) needs v_rad'(r+1).  Output: v_rad', v_azi', e' of ring r.
// PRE: the step began with an accretion call (fargo_dev.h:PreState) — P and H of the rings it may have touched are those of
// the pre-accretion state, like the reference's stored PRESSURE / SCALE_HEIGHT; a separate instantiation, so the kernel
// of every other step is unchanged.
// AV: the artificial-viscosity stage (k_fused_artvisc's) runs on ring r in the same iteration, on the registers that hold the
// source terms' results — v_rad'(r + 1), the one value of the next ring it needs, is already formed — instead of as a second
// kernel that reads them back: 56 bytes per cell less traffic, one warm-up ring and one set of loads / prefetches / address
// arithmetic instead of two.  The window yields the same 60 columns (v_azi'' of column j needs v_azi' of columns j - 1 .. j + 1,
// which needs P of column j - 2: columns [2, 62)); the march starts two rings early instead of one (Q of ring i_first - 1
// needs v_rad' of that ring, which needs ring i_first - 2).
template <bool ADI, bool PRE, bool AV>
__global__ void __launch_bounds__(128, FS_MINB_SRC)
    k_fused_sources(const DevView c, const double *__restrict__ sigma, const double *__restrict__ energy,
		    const double *__restrict__ vr, const double *__restrict__ vp, const double *__restrict__ h_in,
		    double *__restrict__ o_vr, double *__restrict__ o_vp, double *__restrict__ o_e, const double dt, const int R,
		    const PreState pre)
{
    typedef MathP<true> MF;
    typedef MathP<false> MS;
    FsLane L;
    if (!fs_setup(c, L))
	return;
    const int nr = c.nr;
    const int i_first = blockIdx.y * R;
    if (i_first >= nr)
	return;
    const int i_last = min(i_first + R, nr);
    const bool drift = c.p.imposed_disk_drift != 0.0;
    const bool have_h = h_in != nullptr;
    const EosC ec = make_eos_c(c);
    // azimuth of the thread's columns (SideEuler.cpp:56-65)
    double cosj[FS_NC], sinj[FS_NC];
    {
	int cc = L.col;
	FS_FOR4
	{
	    cosj[k] = c.g.cosphi[cc];
	    sinj[k] = c.g.sinphi[cc];
	    cc = (cc + 1 == c.ns) ? 0 : cc + 1;
	}
    }
    double S1[FS_NC], P1[FS_NC], F1[FS_NC], VP1[FS_NC], E1[FS_NC], VRn1[FS_NC], VPn1[FS_NC];
    FS_FOR4
    {
	S1[k] = 1.0;
	P1[k] = F1[k] = VP1[k] = VRn1[k] = VPn1[k] = 0.0;
	E1[k] = 1.0;
    }
    // AV: ring r - 1 as the artificial-viscosity stage left it (Sigma, Q_rr, Q_phiphi)
    const int av_type = c.p.artificial_viscosity;
    const bool av_diss = ADI && c.p.artificial_viscosity_dissipation;
    TempClampNB tc;
    if (AV && ADI)
	tc = make_temp_clamp_nb(c);
    double S2[FS_NC], QR2[FS_NC], QP2[FS_NC];
    FS_FOR4
    {
	S2[k] = 1.0;
	QR2[k] = QP2[k] = 0.0;
    }
    // ring r = kr-1 is finished in iteration kr; its v_rad' needs ring r-1, so start one ring early (AV: two, see above)
    const int kbeg = max(i_first - (AV ? 2 : 1), 0);
    const int rbeg = AV ? max(i_first - 1, 0) : i_first; // first ring whose source terms must be complete
    for (int kr = kbeg; kr <= i_last; ++kr) {
	const bool has_cells = kr < nr;
	double P0[FS_NC], F0[FS_NC], VRn0[FS_NC], VPn0[FS_NC];
	FsRing R0;
	fs_next_ring<ADI>(c, L, sigma, energy, vr, vp, kr, i_last, R0);
	double(&S0)[FS_NC] = R0.S, (&E0)[FS_NC] = R0.E, (&VP0)[FS_NC] = R0.VP, (&VR0)[FS_NC] = R0.VR;
	if (has_cells) {
	    double Hin[FS_NC];
	    FS_FOR4 Hin[k] = 0.0;
	    if (have_h)
		fs_load(h_in, kr, c, L, Hin);
	    if (PRE && pre_has(pre, kr)) { // warp-uniform: a warp marches through whole rings
		double Sp[FS_NC], Ep[FS_NC];
		fs_load(pre.sigma, kr, c, L, Sp);
		if (ADI)
		    fs_load(pre.energy, kr, c, L, Ep);
		else
		    FS_FOR4 Ep[k] = 0.0;
		FS_RUN((st_potential<MF, ADI>(c, ec, kr, Sp, Ep, cosj, sinj, have_h, Hin, P0, F0, A)),
		       (st_potential<MS, ADI>(c, ec, kr, Sp, Ep, cosj, sinj, have_h, Hin, P0, F0, A)));
	    } else {
		FS_RUN((st_potential<MF, ADI>(c, ec, kr, S0, E0, cosj, sinj, have_h, Hin, P0, F0, A)),
		       (st_potential<MS, ADI>(c, ec, kr, S0, E0, cosj, sinj, have_h, Hin, P0, F0, A)));
	    }
	} else {
	    FS_FOR4 { P0[k] = F0[k] = 0.0; }
	}
	FS_FOR4
	{
	    VRn0[k] = VR0[k];
	    VPn0[k] = VP0[k];
	}
	if (kr >= c.one_no_ghost_vr && kr < c.maxmo_no_ghost_vr) {
	    const double VP0r = shfl_from_right(VP0[0]), VP1r = shfl_from_right(VP1[0]);
	    FS_RUN((st_vrad<MF>(c, kr, dt, S0, S1, P0, P1, F0, F1, VP0, VP0r, VP1, VP1r, VR0, VRn0, A)),
		   (st_vrad<MS>(c, kr, dt, S0, S1, P0, P1, F0, F1, VP0, VP0r, VP1, VP1r, VR0, VRn0, A)));
	}
	if (has_cells && kr >= c.zero_no_ghost && kr < c.max_no_ghost) {
	    const double Sl = shfl_from_left(S0[FS_NC - 1]), Pl = shfl_from_left(P0[FS_NC - 1]), Fl = shfl_from_left(F0[FS_NC - 1]);
	    FS_RUN((st_vazi<MF>(c, kr, dt, drift, S0, Sl, P0, Pl, F0, Fl, VP0, VPn0, A)),
		   (st_vazi<MS>(c, kr, dt, drift, S0, Sl, P0, Pl, F0, Fl, VP0, VPn0, A)));
	}
	// ring r = kr-1: compression heating, then store
	const int r = kr - 1;
	if (r >= rbeg) {
	    double En[FS_NC];
	    FS_FOR4 En[k] = E1[k];
	    const double VPn1r = shfl_from_right(VPn1[0]);
	    if (ADI && r < nr - 1) {
		FS_RUN((st_compress<MF>(c, r, dt, E1, VRn0, VRn1, VPn1, VPn1r, En, A)),
		       (st_compress<MS>(c, r, dt, E1, VRn0, VRn1, VPn1, VPn1r, En, A)));
	    }
	    if (AV) { // update_with_artificial_viscosity on ring r: Q(r) from v'(r), v_rad'(r + 1); v''(r) from Q(r), Q(r - 1)
		AvIn I;
		FS_FOR4
		{
		    I.S1[k] = S1[k], I.E1[k] = En[k], I.VR1[k] = VRn1[k], I.VR0[k] = VRn0[k], I.VP1[k] = VPn1[k];
		    I.S2[k] = S2[k], I.QR2[k] = QR2[k], I.QP2[k] = QP2[k];
		}
		I.VP1r = VPn1r;
		double QR1[FS_NC], QP1[FS_NC], Ea[FS_NC], VRa[FS_NC], VPa[FS_NC];
		st_av_q<ADI>(c, r, dt, av_type, av_diss, I, QR1, QP1, Ea);
		if (av_diss) { // :19-21
		    double Ec[FS_NC];
		    FS_RUN(FS_FOR4 Ec[k] = temperature_clamp_nb(tc, I.S1[k], Ea[k], A), FS_FOR4 Ec[k] = temperature_clamp(c, I.S1[k], Ea[k]));
		    FS_FOR4 Ea[k] = Ec[k];
		}
		const double QP1l = shfl_from_left(QP1[FS_NC - 1]), S1l = shfl_from_left(S1[FS_NC - 1]);
		FS_RUN((st_av_v<MF>(c, r, dt, av_type, I, QR1, QP1, QP1l, S1l, VRa, VPa, A)),
		       (st_av_v<MS>(c, r, dt, av_type, I, QR1, QP1, QP1l, S1l, VRa, VPa, A)));
		if (r >= i_first) {
		    fs_store(o_vr, r, c, L, VRa);
		    fs_store(o_vp, r, c, L, VPa);
		    if (ADI)
			fs_store(o_e, r, c, L, Ea);
		}
		FS_FOR4
		{
		    S2[k] = S1[k];
		    QR2[k] = QR1[k];
		    QP2[k] = QP1[k];
		}
	    } else {
		fs_store(o_vr, r, c, L, VRn1);
		fs_store(o_vp, r, c, L, VPn1);
		if (ADI)
		    fs_store(o_e, r, c, L, En);
	    }
	}
	FS_FOR4
	{
	    S1[k] = S0[k];
	    P1[k] = P0[k];
	    F1[k] = F0[k];
	    VP1[k] = VP0[k];
	    E1[k] = E0[k];
	    VRn1[k] = VRn0[k];
	    VPn1[k] = VPn0[k];
	}
    }
    if (i_last == nr) // v_rad ring nr (outermost interface) is outside every update range
	fs_store(o_vr, nr, c, L, VRn1);
}

// ---------------------------------------------------------------------------------------------
// k_fused_artvisc.  Iteration kr loads ring kr; Q(r) of ring r = kr-1 needs v_rad(r+1); v_rad''(r) needs Q(r), Q(r-1).
template <bool ADI>
__global__ void __launch_bounds__(128, FS_MINB_AV)
    k_fused_artvisc(const DevView c, const double *__restrict__ sigma, const double *__restrict__ energy,
		    const double *__restrict__ vr, const double *__restrict__ vp, double *__restrict__ o_vr,
		    double *__restrict__ o_vp, double *__restrict__ o_e, const double dt, const int R)
{
    typedef MathP<true> MF;
    typedef MathP<false> MS;
    FsLane L;
    if (!fs_setup(c, L))
	return;
    const int nr = c.nr;
    const int i_first = blockIdx.y * R;
    if (i_first >= nr)
	return;
    const int i_last = min(i_first + R, nr);
    const int type = c.p.artificial_viscosity;
    const bool diss = ADI && c.p.artificial_viscosity_dissipation;
    TempClampNB tc;
    if (ADI)
	tc = make_temp_clamp_nb(c);
    AvIn I;
    FS_FOR4
    {
	I.S1[k] = I.S2[k] = 1.0;
	I.E1[k] = 1.0;
	I.VR1[k] = I.VP1[k] = I.QR2[k] = I.QP2[k] = 0.0;
    }
    // ring r = kr-1 is finished in iteration kr and needs Q(r-1), i.e. rings r-1 and r: start at r-1 = i_first-1
    const int kbeg = max(i_first - 1, 0);
    for (int kr = kbeg; kr <= i_last; ++kr) {
	FsRing R0;
	fs_next_ring<ADI>(c, L, sigma, energy, vr, vp, kr, i_last, R0);
	double(&S0)[FS_NC] = R0.S, (&E0)[FS_NC] = R0.E, (&VP0)[FS_NC] = R0.VP;
	FS_FOR4 I.VR0[k] = R0.VR[k];
	const int r = kr - 1;
	if (r >= kbeg) { // ring r is complete: S1, E1, VR1 = v_rad(r), VR0 = v_rad(r+1), VP1
	    double QR1[FS_NC], QP1[FS_NC], En[FS_NC], VRn[FS_NC], VPn[FS_NC];
	    I.VP1r = shfl_from_right(I.VP1[0]);
	    st_av_q<ADI>(c, r, dt, type, diss, I, QR1, QP1, En);
	    if (diss) { // :19-21
		double Ec[FS_NC];
		FS_RUN(FS_FOR4 Ec[k] = temperature_clamp_nb(tc, I.S1[k], En[k], A), FS_FOR4 Ec[k] = temperature_clamp(c, I.S1[k], En[k]));
		FS_FOR4 En[k] = Ec[k];
	    }
	    const double QP1l = shfl_from_left(QP1[FS_NC - 1]), S1l = shfl_from_left(I.S1[FS_NC - 1]);
	    FS_RUN((st_av_v<MF>(c, r, dt, type, I, QR1, QP1, QP1l, S1l, VRn, VPn, A)),
		   (st_av_v<MS>(c, r, dt, type, I, QR1, QP1, QP1l, S1l, VRn, VPn, A)));
	    if (r >= i_first) {
		fs_store(o_vr, r, c, L, VRn);
		fs_store(o_vp, r, c, L, VPn);
		if (ADI)
		    fs_store(o_e, r, c, L, En);
	    }
	    FS_FOR4
	    {
		I.S2[k] = I.S1[k];
		I.QR2[k] = QR1[k];
		I.QP2[k] = QP1[k];
	    }
	}
	FS_FOR4
	{
	    I.S1[k] = S0[k];
	    I.E1[k] = E0[k];
	    I.VR1[k] = I.VR0[k];
	    I.VP1[k] = VP0[k];
	}
    }
    if (i_last == nr)
	fs_store(o_vr, nr, c, L, I.VR1);
}

// ---------------------------------------------------------------------------------------------
// viscosity + SubStep3 stage bodies
// recalculate_viscosity (SourceEuler.cpp:205-223): c_s, H, nu of ring kr
template <class M>
__device__ __forceinline__ void st_nu(const DevView &c, const EosC &ec, const int kr, const double (&S0)[FS_NC], const double (&E0)[FS_NC],
				       double (&N0)[FS_NC], double (&H0)[FS_NC], FmAcc &A)
{
    FS_FOR4
    {
	const double cs = eos_cs_m<M>(c, kr, S0[k], E0[k], A);
	const double H = eos_H_m<M>(c, ec, kr, cs, A);
	H0[k] = H;
	N0[k] = (c.p.viscous_alpha > 0) ? c.p.viscous_alpha * H * cs : c.p.constant_viscosity; // viscosity.cpp:98-137
    }
}
struct VsIn {
    // ring r: Sigma, v_rad(r), v_azi, tau_rphi(r) (+ right neighbour), tau_rphi(r+1) (+ right neighbour); ring r-1
    double S1[FS_NC], VR1[FS_NC], VP1[FS_NC], TRP1[FS_NC], TRP0[FS_NC], TRP1r, TRP0r, S2[FS_NC], TRR2[FS_NC], TPP2[FS_NC];
};
// update_velocities_with_viscosity (viscosity.cpp:355-426), ring r
template <class M>
__device__ __forceinline__ void st_visc_v(const DevView &c, const int r, const double dt, const VsIn &I, const double (&TRR1)[FS_NC],
					   const double (&TPP1)[FS_NC], const double TPP1l, const double S1l, double (&VRn)[FS_NC],
					   double (&VPn)[FS_NC], FmAcc &A)
{
    FS_FOR4
    {
	VRn[k] = I.VR1[k];
	VPn[k] = I.VP1[k];
    }
    if (r >= 1 && r < c.nr - 1) {
	const double ra = c.g.rinf[r], rap = c.g.rinf[r + 1];
	const double ra2 = ra * ra, rap2 = rap * rap;
	const double irb = c.g.invrmed[r];
	const double tdr = c.g.twodiffrasq[r]; // 2.0 / (Ra[r+1]^2 - Ra[r]^2), formed on the host with the same IEEE operations
	FS_FOR4
	{
	    const double sp = (k == 0) ? S1l : I.S1[(k + FS_NC - 1) % FS_NC];
	    const double tppl = (k == 0) ? TPP1l : TPP1[(k + FS_NC - 1) % FS_NC];
	    const double sigma_avg = 0.5 * (I.S1[k] + sp);
	    const double dVp = M::div(dt * irb, sigma_avg, A) * (tdr * (rap2 * I.TRP0[k] - ra2 * I.TRP1[k]) + (TPP1[k] - tppl) * c.invdphi);
	    VPn[k] = I.VP1[k] + dVp;
	}
    }
    if (r >= c.one_no_ghost_vr && r < c.maxmo_no_ghost_vr) {
	const double rb = c.g.rmed[r], rbm = c.g.rmed[r - 1], idr = c.g.invdiffrmed[r];
	const double rsum = rb + rbm;
	const double yrsum = M::rcp(rsum, A);
	FS_FOR4
	{
	    const double trpn = (k == FS_NC - 1) ? I.TRP1r : I.TRP1[(k + 1) % FS_NC];
	    const double sigma_avg = 0.5 * (I.S1[k] + I.S2[k]);
	    const double dVr = M::div_y(M::div(dt, sigma_avg, A) * c.p.radial_viscosity_factor * 2.0, rsum, yrsum, A) *
			       ((rb * TRR1[k] - rbm * I.TRR2[k]) * idr + (trpn - I.TRP1[k]) * c.invdphi - 0.5 * (TPP1[k] + I.TPP2[k]));
	    VRn[k] = I.VR1[k] + dVr;
	}
    }
}
// SubStep3 (SourceEuler.cpp:859-954) for ring r: Q+ (viscous_heating :496-536), Q- (thermal_relaxation :632-690),
// radiative alpha_r, the energy update and the temperature floor / ceiling
// RAD: also thermal_cooling / irradiation (kernels_rad.cuh; plain operators, a separate instantiation of the kernel);
// cosj / sinj: azimuth of the thread's columns (only read with RAD)
template <class M, bool RAD>
__device__ __forceinline__ void st_substep3(const DevView &c, const TempClampNB &tc, const int r, const double dt,
					     const double beta_inv, const VsIn &I, const double (&E1)[FS_NC], const double (&N1)[FS_NC],
					     const double (&H1)[FS_NC], const double (&DV1)[FS_NC], const double (&TRR1)[FS_NC],
					     const double (&TPP1)[FS_NC], const double (&s0)[FS_NC], const double (&e0)[FS_NC],
					     const double (&cosj)[FS_NC], const double (&sinj)[FS_NC], double (&Qp)[FS_NC],
					     double (&Qm)[FS_NC], double (&En)[FS_NC], FmAcc &A)
{
    const bool inner = r >= 1 && r < c.nr - 1;
    const fargo_params &p = c.p;
    double tau_eff[FS_NC];
    FS_FOR4 tau_eff[k] = 0.0; // TAU_EFF stays 0 as allocated unless kappa_eff runs
    FS_FOR4
    {
	double qp = 0.0, qm = 0.0;
	if (p.heating_viscous && inner) {
	    const double trpn1 = (k == FS_NC - 1) ? I.TRP1r : I.TRP1[(k + 1) % FS_NC];
	    const double trpn0 = (k == FS_NC - 1) ? I.TRP0r : I.TRP0[(k + 1) % FS_NC];
	    const double tau_r_phi = 0.25 * (I.TRP1[k] + I.TRP0[k] + trpn1 + trpn0);
	    // nu == 0 cells are skipped by the reference; evaluate with a harmless denominator and select
	    const bool on = N1[k] != 0.0;
	    const double den = on ? 2.0 * N1[k] * I.S1[k] : 1.0;
	    double qplus = M::div(1.0, den, A) * (TRR1[k] * TRR1[k] + 2 * (tau_r_phi * tau_r_phi) + TPP1[k] * TPP1[k]);
	    qplus += (2.0 / 9.0) * N1[k] * I.S1[k] * (DV1[k] * DV1[k]);
	    qplus *= p.heating_viscous_factor;
	    qp = on ? 0.0 + qplus : 0.0;
	}
	if (p.cooling_beta && inner) { // qminus_cell
	    double delta_E = E1[k];
	    if (p.cooling_beta_reference & FARGO_BETA_REF_REFERENCE)
		delta_E -= M::div(e0[k], s0[k], A) * I.S1[k];
	    if (p.cooling_beta_reference & FARGO_BETA_REF_MODEL)
		delta_E -= c.g.beta_model_e0[r] * I.S1[k];
	    if (p.cooling_beta_reference & FARGO_BETA_REF_FLOOR)
		delta_E -= M::div_y(M::div_y(p.minimum_temperature * I.S1[k], tc.mu, tc.ymu, A) * p.Rgas, tc.gm1, tc.ygm1, A);
	    qm = 0.0 + delta_E * c.g.omega_k[r] * beta_inv;
	}
	if (RAD && inner) {
	    const double T = rad_temperature(c, I.S1[k], E1[k], p.mu, p.gamma);
	    tau_eff[k] = rad_tau_eff(c, I.S1[k], H1[k], T);
	    if (p.cooling_surface)
		qm += rad_qminus(c, T, tau_eff[k]);
	    if (p.heating_star)
		rad_add_qplus(c, r, cosj[k], sinj[k], H1[k], tau_eff[k], qp);
	}
	double en = E1[k];
	if (inner) {
	    // alpha_r = 1 + 2 H 4 sigma_SB / c (mu (gamma-1) / (R Sigma))^4 e^3  (:921-924)
	    const double inv_pow4 = fm_pow4(M::div(p.mu * (p.gamma - 1.0), p.Rgas * I.S1[k], A));
	    const double alpha = 1.0 + 2.0 * H1[k] * 4.0 * p.sigma_sb / p.c_light * inv_pow4 * fm_pow3(E1[k]);
	    const double ya = M::rcp(alpha, A);
	    qp = M::div_y(qp, alpha, ya, A);
	    qm = M::div_y(qm, alpha, ya, A);
	    en = E1[k] + dt * (qp - qm);
	}
	Qp[k] = qp;
	Qm[k] = qm;
	En[k] = en;
    }
    const double SigmaFloor = 10.0 * p.sigma0 * p.sigma_floor;
    if (inner) {
	FS_FOR4
	{
	    if (I.S1[k] < SigmaFloor) { // rare: cells at the density floor
		const double e4 = Qp[k] * tau_eff[k] / (2.0 * p.sigma_sb);
		const double constant = (p.Rgas / p.mu * I.S1[k] / (p.gamma - 1.0));
		Qm[k] = Qp[k];
		En[k] = pow(e4, 1.0 / 4.0) * constant;
	    }
	}
    }
}

// k_fused_viscosity.  Iteration kr loads ring kr and forms nu(kr) and the corner stress tau_rphi(kr); ring r = kr-1 then
// has everything: div v, tau_rr, tau_phiphi (need v_rad(r+1)), the velocity updates (need tau_rphi(r+1), the
// centred stresses of r-1) and, for the energy equation, Q+ / Q- / the new energy.
// StabilizeViscosity != 0 is not handled here (the host falls back to the staged kernels).
template <bool ADI, bool RAD>
__global__ void __launch_bounds__(128, FS_MINB_VISC)
    k_fused_viscosity(const DevView c, const double *__restrict__ sigma, const double *__restrict__ energy,
		      const double *__restrict__ vr, const double *__restrict__ vp, const double *__restrict__ sigma0,
		      const double *__restrict__ energy0, double *__restrict__ o_vr, double *__restrict__ o_vp,
		      double *__restrict__ o_e, double *__restrict__ o_qplus, double *__restrict__ o_qminus,
		      double *__restrict__ o_h, const double dt, const double beta_inv, const int R)
{
    typedef MathP<true> MF;
    typedef MathP<false> MS;
    FsLane L;
    if (!fs_setup(c, L))
	return;
    const int nr = c.nr;
    const int i_first = blockIdx.y * R;
    if (i_first >= nr)
	return;
    const int i_last = min(i_first + R, nr);
    const bool need0 = ADI && c.p.cooling_beta && (c.p.cooling_beta_reference & FARGO_BETA_REF_REFERENCE);
    const EosC ec = make_eos_c(c);
    TempClampNB tc;
    if (ADI)
	tc = make_temp_clamp_nb(c);
    double cosj[FS_NC], sinj[FS_NC]; // azimuth of the thread's columns (irradiation only)
    FS_FOR4 cosj[k] = sinj[k] = 0.0;
    if (RAD) {
	int cc = L.col;
	FS_FOR4
	{
	    cosj[k] = c.g.cosphi[cc];
	    sinj[k] = c.g.sinphi[cc];
	    cc = (cc + 1 == c.ns) ? 0 : cc + 1;
	}
    }
    VsIn I;
    double E1[FS_NC], N1[FS_NC], H1[FS_NC];
    FS_FOR4
    {
	I.S1[k] = I.S2[k] = 1.0;
	E1[k] = 1.0;
	I.VR1[k] = I.VP1[k] = N1[k] = H1[k] = I.TRP1[k] = I.TRR2[k] = I.TPP2[k] = 0.0;
    }
    // ring r needs the centred stresses of r-1 (rings r-1, r) and tau_rphi(r) (rings r-1, r): start at kr = r-1
    const int kbeg = max(i_first - 1, 0);
    for (int kr = kbeg; kr <= i_last; ++kr) {
	double N0[FS_NC], H0[FS_NC];
	FsRing R0;
	fs_next_ring<ADI>(c, L, sigma, energy, vr, vp, kr, i_last, R0);
	double(&S0)[FS_NC] = R0.S, (&E0)[FS_NC] = R0.E, (&VP0)[FS_NC] = R0.VP, (&VR0)[FS_NC] = R0.VR;
	if (need0 && kr < i_last && kr >= i_first) { // ring r + 1 = kr of the reference fields, read one iteration from now
	    fs_prefetch(sigma0, kr, c, L);
	    fs_prefetch(energy0, kr, c, L);
	}
	if (kr < nr) {
	    FS_RUN((st_nu<MF>(c, ec, kr, S0, E0, N0, H0, A)), (st_nu<MS>(c, ec, kr, S0, E0, N0, H0, A)));
	} else {
	    FS_FOR4 { N0[k] = H0[k] = 0.0; }
	}
	// tau_rphi at the corner (kr, j) (viscosity.cpp:213-253); rings 0 and nr of the grid stay 0
	{
	    const double VR0l = shfl_from_left(VR0[FS_NC - 1]);
	    const double N0l = shfl_from_left(N0[FS_NC - 1]), N1l = shfl_from_left(N1[FS_NC - 1]);
	    const double S0l = shfl_from_left(S0[FS_NC - 1]), S1l = shfl_from_left(I.S1[FS_NC - 1]);
	    if (kr >= 1 && kr < nr) {
		const double irb = c.g.invrmed[kr], irbm = c.g.invrmed[kr - 1], idr = c.g.invdiffrmed[kr];
		const double ra = c.g.rinf[kr], ira = c.g.invrinf[kr];
		FS_FOR4
		{
		    const double vrl = (k == 0) ? VR0l : VR0[(k + FS_NC - 1) % FS_NC];
		    const double n0l = (k == 0) ? N0l : N0[(k + FS_NC - 1) % FS_NC];
		    const double n1l = (k == 0) ? N1l : N1[(k + FS_NC - 1) % FS_NC];
		    const double s0l = (k == 0) ? S0l : S0[(k + FS_NC - 1) % FS_NC];
		    const double s1l = (k == 0) ? S1l : I.S1[(k + FS_NC - 1) % FS_NC];
		    const double dvazirdr = (VP0[k] * irb - I.VP1[k] * irbm) * idr;
		    const double dvrdphi = (VR0[k] - vrl) * c.invdphi;
		    const double drp = ra * dvazirdr + dvrdphi * ira;
		    const double nua = 0.25 * (N0[k] + N1[k] + n0l + n1l);
		    const double sa = 0.25 * (S0[k] + I.S1[k] + s0l + s1l);
		    I.TRP0[k] = nua * sa * drp;
		}
	    } else {
		FS_FOR4 I.TRP0[k] = 0.0;
	    }
	}
	const int r = kr - 1;
	if (r >= kbeg) {
	    // centred stresses of ring r (viscosity.cpp:150-211)
	    double DV1[FS_NC], TRR1[FS_NC], TPP1[FS_NC], VRn[FS_NC], VPn[FS_NC];
	    const double VP1r = shfl_from_right(I.VP1[0]);
	    {
		const double ra1 = c.g.rinf[r + 1], ra0 = c.g.rinf[r], idrb = c.g.invdiffrsuprb[r], irb = c.g.invrmed[r];
		const double ids = c.g.invdiffrsup[r];
		FS_FOR4
		{
		    const double vpn = (k == FS_NC - 1) ? VP1r : I.VP1[(k + 1) % FS_NC];
		    const double dv = (VR0[k] * ra1 - I.VR1[k] * ra0) * idrb + (vpn - I.VP1[k]) * c.invdphi * irb;
		    DV1[k] = dv;
		    const double drr = (VR0[k] - I.VR1[k]) * ids;
		    TRR1[k] = 2.0 * N1[k] * I.S1[k] * (drr - 1.0 / 3.0 * dv);
		    const double dpp = (vpn - I.VP1[k]) * c.invdphi * irb + 0.5 * (VR0[k] + I.VR1[k]) * irb;
		    TPP1[k] = 2.0 * N1[k] * I.S1[k] * (dpp - 1.0 / 3.0 * dv);
		}
	    }
	    const double TPP1l = shfl_from_left(TPP1[FS_NC - 1]), S1l = shfl_from_left(I.S1[FS_NC - 1]);
	    I.TRP1r = shfl_from_right(I.TRP1[0]);
	    I.TRP0r = shfl_from_right(I.TRP0[0]);
	    FS_RUN((st_visc_v<MF>(c, r, dt, I, TRR1, TPP1, TPP1l, S1l, VRn, VPn, A)),
		   (st_visc_v<MS>(c, r, dt, I, TRR1, TPP1, TPP1l, S1l, VRn, VPn, A)));
	    if (r >= i_first) {
		fs_store(o_vr, r, c, L, VRn);
		fs_store(o_vp, r, c, L, VPn);
		if (o_h) // leapfrog: H as recalculate_viscosity sees it, for the second kick's potential smoothing
		    fs_store(o_h, r, c, L, H1);
	    }
	    if (ADI) {
		double Qp[FS_NC], Qm[FS_NC], En[FS_NC], Ec[FS_NC], s0[FS_NC], e0[FS_NC];
		if (need0 && r >= i_first) {
		    fs_load(sigma0, r, c, L, s0);
		    fs_load(energy0, r, c, L, e0);
		} else {
		    FS_FOR4 { s0[k] = 1.0, e0[k] = 1.0; }
		}
		FS_RUN((st_substep3<MF, RAD>(c, tc, r, dt, beta_inv, I, E1, N1, H1, DV1, TRR1, TPP1, s0, e0, cosj, sinj, Qp, Qm, En, A)),
		       (st_substep3<MS, RAD>(c, tc, r, dt, beta_inv, I, E1, N1, H1, DV1, TRR1, TPP1, s0, e0, cosj, sinj, Qp, Qm, En, A)));
		FS_RUN(FS_FOR4 Ec[k] = temperature_clamp_nb(tc, I.S1[k], En[k], A), FS_FOR4 Ec[k] = temperature_clamp(c, I.S1[k], En[k]));
		if (r >= i_first) {
		    fs_store(o_qplus, r, c, L, Qp);
		    fs_store(o_qminus, r, c, L, Qm);
		    fs_store(o_e, r, c, L, Ec);
		}
	    }
	    FS_FOR4
	    {
		I.S2[k] = I.S1[k];
		I.TRR2[k] = TRR1[k];
		I.TPP2[k] = TPP1[k];
	    }
	}
	FS_FOR4
	{
	    I.S1[k] = S0[k];
	    E1[k] = E0[k];
	    I.VR1[k] = VR0[k];
	    I.VP1[k] = VP0[k];
	    N1[k] = N0[k];
	    H1[k] = H0[k];
	    I.TRP1[k] = I.TRP0[k];
	}
    }
    if (i_last == nr)
	fs_store(o_vr, nr, c, L, I.VR1);
}
