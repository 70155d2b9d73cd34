// kernels_azimuthal.cuh — azimuthal half of Transport (TransportEuler.cpp:270-304, 416-466, 630-664), the integer
// FARGO shift (AdvectSHIFT :238-268), velocities from momenta (:498-535) and the floors (:123-131) in ONE kernel
// that keeps a ring segment in registers.
//
// Layout: a warp owns a window of 128 consecutive OUTPUT columns of one ring, each lane 4 consecutive columns
// (so 3 of the 4 azimuthal neighbours of a cell are the thread's own registers and the fourth is one warp
// shuffle away).  The segment is loaded already rotated by the ring's integer shift (output column j comes from
// pre-shift column j - Nshift[i]); the residual-velocity pass and the uniform pass then run on it in registers,
// quantity by quantity, with Sigma* and the upwind selectors shared by all six quantities.  Each van Leer pass
// invalidates 2 columns at either end of the window and v_azi needs one more on the left, so a window of 128
// columns yields 116 finished ones ([8, 124), 32-byte aligned for vector stores); warps are independent — no
// shared memory, no block barrier.  The warp then marches outward ring by ring, carrying the previous ring's
// transported Sigma and rm+ (v_rad couples rings i-1 and i at the same output column).
//
// Arithmetic is the reference's, operation for operation (-fmad=false); the only algebraic liberties are exact
// ones: x - c*d == x + (-c)*d, dx + ksi == dx - |ksi| for ksi <= 0, and the branch-free division of fargo_math.h
// (the compiler's own IEEE sequence, emitted straight-line so the Newton chains of the 4 columns interleave).
#pragma once
#include "fargo_dev.h"
#include "fargo_math.h"

#define AZ_WIN 128 // columns per warp window
#define AZ_HL 8	   // invalid columns at the left end (5 needed, rounded up for 32-byte aligned stores)
#define AZ_HR 4	   // invalid columns at the right end
#define AZ_OUT (AZ_WIN - AZ_HL - AZ_HR)

__device__ __forceinline__ double shfl_from_left(const double x) { return __shfl_up_sync(0xffffffffu, x, 1); }
__device__ __forceinline__ double shfl_from_right(const double x) { return __shfl_down_sync(0xffffffffu, x, 1); }

struct AzRing {
    double dxtheta, invdxtheta, dxrad, invsurf;
};

// ComputeStarTheta (:416-466) for one base quantity B on the thread's 4 columns: limited slopes, then the
// upwinded interface values.  pos[c]: ksi > 0 at interface c (between columns c-1 and c); cf[c] = +-(dxtheta -+ ksi).
template <int LIM>
__device__ __forceinline__ void az_star(const double (&B)[4], const AzRing &g, const bool (&pos)[4], const double (&cf)[4],
					 double (&star)[4])
{
    const double Bm = shfl_from_left(B[3]);
    const double Bp = shfl_from_right(B[0]);
    const double dq[5] = {B[0] - Bm, B[1] - B[0], B[2] - B[1], B[3] - B[2], Bp - B[3]};
    double D[4];
    FmAcc acc;
#pragma unroll
    for (int c = 0; c < 4; ++c)
	D[c] = 0.5 * limiter_nb<LIM>(dq[c + 1], dq[c], acc) * g.invdxtheta;
    if (!fm_acc_ok(acc)) { // cold: extreme exponents
#pragma unroll
	for (int c = 0; c < 4; ++c)
	    D[c] = 0.5 * flux_limiter<LIM>(dq[c + 1], dq[c]) * g.invdxtheta;
    }
    const double Dm = shfl_from_left(D[3]);
    star[0] = (pos[0] ? Bm : B[0]) + cf[0] * (pos[0] ? Dm : D[0]);
    star[1] = (pos[1] ? B[0] : B[1]) + cf[1] * (pos[1] ? D[0] : D[1]);
    star[2] = (pos[2] ? B[1] : B[2]) + cf[2] * (pos[2] ? D[1] : D[2]);
    star[3] = (pos[3] ? B[2] : B[3]) + cf[3] * (pos[3] ? D[2] : D[3]);
}

// VanLeerTheta (:630-664) conservative update of one quantity from its interface fluxes
__device__ __forceinline__ void az_update(double (&Q)[4], const double (&G)[4], const AzRing &g)
{
    const double Gp = shfl_from_right(G[0]);
    double varq;
    varq = G[0];
    varq -= G[1];
    Q[0] += varq * g.invsurf;
    varq = G[1];
    varq -= G[2];
    Q[1] += varq * g.invsurf;
    varq = G[2];
    varq -= G[3];
    Q[2] += varq * g.invsurf;
    varq = G[3];
    varq -= Gp;
    Q[3] += varq * g.invsurf;
}

// QuantitiesAdvection (:292-304): Sigma* from the current Sigma, Sigma_int = copy, then rm+, rm-, am+, am-, (e), Sigma.
// Q index: 0 rm+, 1 rm-, 2 am+, 3 am-, 4 e, 5 Sigma.
template <int LIM, bool ADI>
__device__ __forceinline__ void az_pass(double (&Q)[6][4], const double (&u)[4], const AzRing &g, const double dt)
{
    bool pos[4];
    double cf[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
	const double ksi = u[c] * dt;
	pos[c] = ksi > 0.0;
	const double coef = g.dxtheta - fabs(ksi); // (dxtheta - ksi) for ksi > 0, (dxtheta + ksi) otherwise
	cf[c] = pos[c] ? coef : -coef;
    }
    double starS[4];
    az_star<LIM>(Q[5], g, pos, cf, starS);
    double yS[4];
    FmAcc accS;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
	yS[c] = fm_rcp_raw(Q[5][c]);
	fm_acc_nrm(accS, Q[5][c]);
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) {
	if (q == 4 && !ADI)
	    continue;
	double W[4], st[4], G[4];
	FmAcc acc = accS;
#pragma unroll
	for (int c = 0; c < 4; ++c) {
	    W[c] = fm_div_raw(Q[q][c], Q[5][c], yS[c]); // divise_polargrid (SideEuler.cpp:27-43)
	    fm_acc_nrm(acc, W[c]);
	}
	if (!fm_acc_ok(acc)) { // cold: zero / tiny momenta
#pragma unroll
	    for (int c = 0; c < 4; ++c)
		W[c] = Q[q][c] / Q[5][c];
	}
	az_star<LIM>(W, g, pos, cf, st);
#pragma unroll
	for (int c = 0; c < 4; ++c)
	    G[c] = g.dxrad * st[c] * starS[c] * u[c];
	az_update(Q[q], G, g);
    }
    { // Sigma itself: QRStar == 1 (Sigma / Sigma_int with zero slopes)
	double G[4];
#pragma unroll
	for (int c = 0; c < 4; ++c)
	    G[c] = g.dxrad * starS[c] * u[c];
	az_update(Q[5], G, g);
    }
}

template <int LIM, bool ADI>
__global__ void __launch_bounds__(128, 3)
    k_transport_azimuthal(const DevView c, const double *__restrict__ t_sigma, const double *__restrict__ t_rmp,
			  const double *__restrict__ t_rmm, const double *__restrict__ t_amp,
			  const double *__restrict__ t_amm, const double *__restrict__ t_e, const double *__restrict__ vp_old,
			  const double *__restrict__ vr_old, const double *__restrict__ vmean, const int *__restrict__ nshift,
			  const double *__restrict__ vconst, double *__restrict__ o_sigma, double *__restrict__ o_vr,
			  double *__restrict__ o_vp, double *__restrict__ o_e, const double dt, const int R)
{
    const int ns = c.ns, nr = c.nr;
    const int lane = threadIdx.x & 31;
    const int win = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if ((long long)win * AZ_OUT >= ns)
	return; // whole warp; warps never synchronise with each other
    const int i_first = blockIdx.y * R;
    if (i_first >= nr)
	return;
    const int i_last = min(i_first + R, nr);
    const int t0 = 4 * lane;			    // local column of c = 0
    const int jout = win * AZ_OUT - AZ_HL + t0;	    // output column of c = 0 (negative / >= ns in the halo)
    const bool lane_out = (t0 >= AZ_HL) && (t0 < AZ_WIN - AZ_HR);
    const bool vec_ok = ((ns & 3) == 0);
    const double OmegaF = c.b.omega_frame;
    const double floorv = c.p.sigma_floor * c.p.sigma0;
    const bool fargo = c.p.fast_transport != 0;
    TempClampNB tc;
    if (ADI)
	tc = make_temp_clamp_nb(c);

    double PS[4] = {0.0, 0.0, 0.0, 0.0}, PR[4] = {0.0, 0.0, 0.0, 0.0}; // previous ring: transported Sigma, rm+

    for (int i = max(i_first - 1, 0); i < i_last; ++i) {
	const int nsh = nshift[i];
	const double vm = vmean[i], vc = vconst[i];
	const double rmed = c.g.rmed[i], invrmed = c.g.invrmed[i];
	AzRing g;
	g.dxtheta = c.dphi * rmed;
	g.invdxtheta = c.g.invdxtheta[i]; // 1.0 / dxtheta, formed on the host with the same IEEE division
	g.dxrad = (c.g.rsup[i] - c.g.rinf[i]) * dt;
	g.invsurf = c.g.invsurf[i];
	// pre-shift column of c = 0
	int col = (jout - nsh) % ns;
	if (col < 0)
	    col += ns;
	const size_t row = (size_t)i * ns;
	if (i + 1 < i_last) { // prefetch the next ring's (rotated) segment: first and last of the 4 columns cover its sectors
	    int cn = (jout - nshift[i + 1]) % ns;
	    if (cn < 0)
		cn += ns;
	    const int cl = (cn + 3 >= ns) ? cn + 3 - ns : cn + 3;
	    const size_t a0 = row + (size_t)ns + (size_t)cn, a1 = row + (size_t)ns + (size_t)cl;
	    pf_global(t_rmp + a0), pf_global(t_rmp + a1);
	    pf_global(t_rmm + a0), pf_global(t_rmm + a1);
	    pf_global(t_amp + a0), pf_global(t_amp + a1);
	    pf_global(t_amm + a0), pf_global(t_amm + a1);
	    pf_global(t_sigma + a0), pf_global(t_sigma + a1);
	    pf_global(vp_old + a0), pf_global(vp_old + a1);
	    if (ADI)
		pf_global(t_e + a0), pf_global(t_e + a1);
	}
	double Q[6][4], U[4];
#pragma unroll
	for (int k = 0; k < 4; ++k) {
	    const size_t a = row + (size_t)col;
	    Q[0][k] = t_rmp[a];
	    Q[1][k] = t_rmm[a];
	    Q[2][k] = t_amp[a];
	    Q[3][k] = t_amm[a];
	    Q[4][k] = ADI ? t_e[a] : 0.0;
	    Q[5][k] = t_sigma[a];
	    double u = vp_old[a] - vm; // compute_residual_velocity :194-205
	    if (!fargo)
		u = vc + u; // ComputeConstantResidual :225-231
	    U[k] = u;
	    col = (col + 1 == ns) ? 0 : col + 1;
	}
	// pass 1: residual velocity; pass 2: constant residual velocity (skipped for standard transport, :646).
	// One copy of the pass code (the loop is deliberately not unrolled: the kernel must stay inside the I-cache).
	const int npass = fargo ? 2 : 1;
#pragma unroll 1
	for (int pass = 0; pass < npass; ++pass) {
	    az_pass<LIM, ADI>(Q, U, g, dt);
#pragma unroll
	    for (int k = 0; k < 4; ++k)
		U[k] = vc;
	}
	// velocities from momenta (:498-535), floors (:123-131)
	const double am_left = shfl_from_left(Q[2][3]);
	const double s_left = shfl_from_left(Q[5][3]);
	if (i >= i_first && lane_out) {
	    double vrn[4], vpn[4], sf[4], en[4], nvr[4], dvr[4], nvp[4], dvp[4];
	    FmAcc acc;
#pragma unroll
	    for (int k = 0; k < 4; ++k) {
		const double s = Q[5][k];
		const double sm = (k == 0) ? s_left : Q[5][k - 1];
		const double amp_m = (k == 0) ? am_left : Q[2][k - 1];
		nvr[k] = PR[k] + Q[1][k];
		dvr[k] = PS[k] + s;
		nvp[k] = amp_m + Q[3][k];
		dvp[k] = sm + s;
		const double qr = fm_div_raw(nvr[k], dvr[k], fm_rcp_raw(dvr[k]));
		const double qp = fm_div_raw(nvp[k], dvp[k], fm_rcp_raw(dvp[k]));
		fm_acc_nrm_if(acc, i != 0, dvr[k]);
		fm_acc_nrm_if(acc, i != 0, qr);
		fm_acc_nrm(acc, dvp[k]);
		fm_acc_nrm(acc, qp);
		vrn[k] = (i == 0) ? 0.0 : qr;
		vpn[k] = qp * invrmed - rmed * OmegaF;
		sf[k] = (s < floorv) ? floorv : s;
		en[k] = ADI ? temperature_clamp_nb(tc, sf[k], Q[4][k], acc) : 0.0;
	    }
	    if (!fm_acc_ok(acc)) { // cold
#pragma unroll
		for (int k = 0; k < 4; ++k) {
		    vrn[k] = (i == 0) ? 0.0 : nvr[k] / dvr[k];
		    vpn[k] = nvp[k] / dvp[k] * invrmed - rmed * OmegaF;
		    if (ADI)
			en[k] = temperature_clamp(c, sf[k], Q[4][k]);
		}
	    }
	    if (vec_ok) {
		if (jout < ns) { // jout is a multiple of 4, so all four columns are inside
		    const size_t a = row + (size_t)jout;
		    *reinterpret_cast<double2 *>(o_vr + a) = make_double2(vrn[0], vrn[1]);
		    *reinterpret_cast<double2 *>(o_vr + a + 2) = make_double2(vrn[2], vrn[3]);
		    *reinterpret_cast<double2 *>(o_vp + a) = make_double2(vpn[0], vpn[1]);
		    *reinterpret_cast<double2 *>(o_vp + a + 2) = make_double2(vpn[2], vpn[3]);
		    *reinterpret_cast<double2 *>(o_sigma + a) = make_double2(sf[0], sf[1]);
		    *reinterpret_cast<double2 *>(o_sigma + a + 2) = make_double2(sf[2], sf[3]);
		    if (ADI) {
			*reinterpret_cast<double2 *>(o_e + a) = make_double2(en[0], en[1]);
			*reinterpret_cast<double2 *>(o_e + a + 2) = make_double2(en[2], en[3]);
		    }
		}
	    } else {
#pragma unroll
		for (int k = 0; k < 4; ++k) {
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
	for (int k = 0; k < 4; ++k) {
	    PS[k] = Q[5][k];
	    PR[k] = Q[0][k];
	}
    }
    // v_rad ring nr is not touched by compute_velocities_from_momenta (:502-507): carry it over
    if (i_last == nr && lane_out) {
#pragma unroll
	for (int k = 0; k < 4; ++k)
	    if (jout + k < ns)
		o_vr[(size_t)nr * ns + jout + k] = vr_old[(size_t)nr * ns + jout + k];
    }
}
