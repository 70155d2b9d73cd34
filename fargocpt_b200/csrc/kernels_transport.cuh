// kernels_transport.cuh — Transport (TransportEuler.cpp:112-664) as two fused kernels + a ring pre-pass.
//
//   k_ring_mean            sequential (index-ordered) ring sums of v_azi  -> <v>, Nshift, v_const
//                          (compute_average_azimuthal_velocity :174-189, ComputeConstantResidual :207-236)
//   k_transport_radial     momenta from velocities (:471-493) + the six VanLeerRadial sweeps (:138-167,
//                          :349-406, :545-620) in ONE pass: a thread owns one azimuthal column and marches
//                          outward in radius with a register sliding window; Sigma* and Sigma_int are shared by
//                          all transported quantities; every interface flux is evaluated once.
//   k_transport_azimuthal  residual-velocity pass + uniform pass (QuantitiesAdvection :292-304, ComputeStarTheta
//                          :416-466, VanLeerTheta :630-664), the integer FARGO shift (AdvectSHIFT :238-268, as an
//                          index rotation on load), velocities from momenta (:498-535) and the density /
//                          temperature floors (:123-131), all on ring segments staged in shared memory.
//
// The reference streams ~300 full-grid arrays through memory for this; here Transport reads 4+1+7 and
// writes 6+4 arrays.
#pragma once
#include "fargo_dev.h"

// ---------------------------------------------------------------------------------------------
// Ring means.  The reference sums v_azi(i, 0..Ns-1) strictly in index order (cfl.cpp:199-204,
// TransportEuler.cpp:179-187); the result feeds the CFL dt and the integer shifts, which must be bit-exact,
// so the order is kept: one warp per ring, coalesced 32-wide loads, and a shuffle chain that adds the 32
// values in lane order.
// mode 0: only vmean (CFL).  mode 1: also Nshift[i] and the constant residual velocity (transport).
__global__ void __launch_bounds__(256)
    k_ring_mean(const DevView c, const double *__restrict__ vp, double *__restrict__ vmean, int *__restrict__ nshift,
		double *__restrict__ vconst, const double dt, const int mode)
{
    const int warp = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (warp >= c.nr)
	return;
    const double *row = vp + (size_t)warp * c.ns;
    double s = 0.0;
    const int nfull = c.ns & ~31;
    int j0 = 0;
    // software-pipelined: keep 4 independent loads in flight
    for (; j0 + 128 <= nfull; j0 += 128) {
	const double a0 = row[j0 + lane], a1 = row[j0 + 32 + lane], a2 = row[j0 + 64 + lane], a3 = row[j0 + 96 + lane];
#pragma unroll
	for (int k = 0; k < 32; ++k)
	    s += __shfl_sync(0xffffffffu, a0, k);
#pragma unroll
	for (int k = 0; k < 32; ++k)
	    s += __shfl_sync(0xffffffffu, a1, k);
#pragma unroll
	for (int k = 0; k < 32; ++k)
	    s += __shfl_sync(0xffffffffu, a2, k);
#pragma unroll
	for (int k = 0; k < 32; ++k)
	    s += __shfl_sync(0xffffffffu, a3, k);
    }
    for (; j0 < nfull; j0 += 32) {
	const double a0 = row[j0 + lane];
#pragma unroll
	for (int k = 0; k < 32; ++k)
	    s += __shfl_sync(0xffffffffu, a0, k);
    }
    if (j0 < c.ns) {
	const double a0 = (j0 + lane < c.ns) ? row[j0 + lane] : 0.0;
	const int rem = c.ns - j0;
	for (int k = 0; k < rem; ++k)
	    s += __shfl_sync(0xffffffffu, a0, k);
    }
    if (lane == 0) {
	const double mean = s / (double)c.ns;
	vmean[warp] = mean;
	if (mode == 1) {
	    const double invdt = 1.0 / dt;
	    const double Ntilde = mean * c.g.invrmed[warp] * dt * c.invdphi;
	    const double Nround = floor(Ntilde + 0.5);
	    nshift[warp] = (int)Nround;
	    vconst[warp] = (Ntilde - Nround) * c.g.rmed[warp] * invdt * c.dphi;
	}
    }
}

// ---------------------------------------------------------------------------------------------
// Radial sweep.  Quantity order (the reference's, Sigma last): rm+, rm-, am+, am-, e, Sigma.
// State per transported base b (w = Q / Sigma_int for the five non-density quantities, Sigma itself for
// Sigma*): b(k-2), b(k-1), dq(k-2).  When ring k arrives: dq(k-1), star(k-1) -> interface flux F(k-1), then
// cell k-2 is finished:  Q(k-2) += (F(k-2) - F(k-1)) * InvSurf[k-2].
template <int LIM, bool ADIABATIC>
__global__ void __launch_bounds__(128)
    k_transport_radial(const DevView c, const double *__restrict__ sigma, const double *__restrict__ vr,
		       const double *__restrict__ vp, const double *__restrict__ energy, double *__restrict__ o_sigma,
		       double *__restrict__ o_rmp, double *__restrict__ o_rmm, double *__restrict__ o_amp,
		       double *__restrict__ o_amm, double *__restrict__ o_e, const double dt, const int chunk)
{
    constexpr int NB = ADIABATIC ? 6 : 5; // bases: Sigma, w_rm+, w_rm-, w_am+, w_am-, (w_e)
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= c.ns)
	return;
    const int jp = (j == c.ns - 1) ? 0 : j + 1;
    const int nr = c.nr;
    const int i0 = blockIdx.y * chunk;
    const int i1 = min(i0 + chunk, nr);
    if (i0 >= nr)
	return;
    const double dtdphi = dt * c.dphi;
    const double OmegaF = c.b.omega_frame;

    double b1[NB], b2[NB], dq2[NB]; // b(k-1), b(k-2), dq(k-2)
    double raw1[NB], raw2[NB];	    // raw transported quantities of rings k-1, k-2 (index 0 = Sigma)
    double F2[NB];		    // interface fluxes F(k-2) (index 0 = mass flux)
    double v1 = 0.0;		    // v_r(k-1)
#pragma unroll
    for (int q = 0; q < NB; ++q) {
	b1[q] = b2[q] = dq2[q] = raw1[q] = raw2[q] = F2[q] = 0.0;
    }

    const int kstart = max(i0 - 2, 0);
    for (int k = kstart; k <= i1 + 1; ++k) {
	double bk[NB], rawk[NB];
	double vk = 0.0;
	if (k < nr) {
	    const double s = AT(sigma, k, j);
	    vk = AT(vr, k, j);
	    const double vk1 = AT(vr, k + 1, j);
	    const double vpj = AT(vp, k, j), vpn = AT(vp, k, jp);
	    const double r = c.g.rmed[k];
	    rawk[0] = s;
	    rawk[1] = s * vk1;			 // radial_momentum_plus  (:484)
	    rawk[2] = s * vk;			 // radial_momentum_minus (:485)
	    rawk[3] = s * (vpn + r * OmegaF) * r; // angular_momentum_plus (:489)
	    rawk[4] = s * (vpj + r * OmegaF) * r; // angular_momentum_minus(:490)
	    if (ADIABATIC)
		rawk[5] = AT(energy, k, j);
	    bk[0] = s;
#pragma unroll
	    for (int q = 1; q < NB; ++q)
		bk[q] = rawk[q] / s; // divise_polargrid (SideEuler.cpp:27-43), Sigma_int == Sigma before the sweep
	} else {
#pragma unroll
	    for (int q = 0; q < NB; ++q)
		bk[q] = rawk[q] = 0.0;
	    if (k == nr)
		vk = AT(vr, nr, j);
	}
	// slopes of ring k-1 (compute_star_radial :356-371)
	const int m = k - 1;
	double dq1[NB];
	if (m >= 1 && m < nr - 1 && k < nr) {
	    const double idm = c.g.invdiffrmed[m], idp = c.g.invdiffrmed[m + 1];
#pragma unroll
	    for (int q = 0; q < NB; ++q) {
		const double dqm = (b1[q] - b2[q]) * idm;
		const double dqp = (bk[q] - b1[q]) * idp;
		dq1[q] = flux_limiter<LIM>(dqp, dqm);
	    }
	} else {
#pragma unroll
	    for (int q = 0; q < NB; ++q)
		dq1[q] = 0.0;
	}
	// star values and fluxes through interface m = k-1 (:376-392, :569-575)
	double F1[NB];
	if (m >= 1 && m < nr && m >= kstart + 1) {
	    const double geo = dtdphi * c.g.rinf[m];
	    double star[NB];
	    if (v1 > 0.0) {
		const double h = (c.g.rmed[m] - c.g.rmed[m - 1] - v1 * dt) * 0.5;
#pragma unroll
		for (int q = 0; q < NB; ++q)
		    star[q] = b2[q] + h * dq2[q];
	    } else {
		const double h = (c.g.rmed[m + 1] - c.g.rmed[m] + v1 * dt) * 0.5;
#pragma unroll
		for (int q = 0; q < NB; ++q)
		    star[q] = b1[q] - h * dq1[q];
	    }
	    F1[0] = geo * 1.0 * star[0] * v1; // Sigma: QRStar == 1 (Sigma/Sigma_int), DensityStar = star[0]
#pragma unroll
	    for (int q = 1; q < NB; ++q)
		F1[q] = geo * star[q] * star[0] * v1;
	} else {
#pragma unroll
	    for (int q = 0; q < NB; ++q)
		F1[q] = 0.0; // star(0,.) = 0 (:394-396); ring nr of the star grids stays 0 (:82-91)
	}
	// finish cell n = k-2 (:577)
	const int n = k - 2;
	if (n >= i0 && n < i1) {
	    const double is = c.g.invsurf[n];
	    AT(o_sigma, n, j) = raw2[0] + (F2[0] - F1[0]) * is;
	    AT(o_rmp, n, j) = raw2[1] + (F2[1] - F1[1]) * is;
	    AT(o_rmm, n, j) = raw2[2] + (F2[2] - F1[2]) * is;
	    AT(o_amp, n, j) = raw2[3] + (F2[3] - F1[3]) * is;
	    AT(o_amm, n, j) = raw2[4] + (F2[4] - F1[4]) * is;
	    if (ADIABATIC)
		AT(o_e, n, j) = raw2[5] + (F2[5] - F1[5]) * is;
	}
#pragma unroll
	for (int q = 0; q < NB; ++q) {
	    b2[q] = b1[q];
	    b1[q] = bk[q];
	    dq2[q] = dq1[q];
	    raw2[q] = raw1[q];
	    raw1[q] = rawk[q];
	    F2[q] = F1[q];
	}
	v1 = vk;
    }
}

// ---------------------------------------------------------------------------------------------
// Azimuthal sweep on ring segments.  Output column j of ring i comes from pre-shift column j - Nshift[i]
// (AdvectSHIFT), so the segment is loaded already rotated and the two van Leer passes run on it in shared
// memory.  Halo: pass 1 and pass 2 each consume 2 cells per side, v_azi needs one more on the left.
#define AZ_HL 5
#define AZ_HR 4
#define AZ_NQ 6 // rm+, rm-, am+, am-, e, Sigma  (e unused when isothermal)

template <int LIM>
__device__ __forceinline__ void az_pass(const DevView &c, double *__restrict__ Q /*[AZ_NQ][L]*/, const double *__restrict__ U,
					const bool uniform_u, const double u_ring, double *__restrict__ W,
					double *__restrict__ D, const int L, const int lo, const int hi, const int i,
					const double dt)
{
    // valid input range [lo, hi); produces updated Q on [lo+2, hi-2)
    const double dxtheta = c.dphi * c.g.rmed[i];
    const double invdxtheta = 1.0 / dxtheta;
    const double dxrad = (c.g.rsup[i] - c.g.rinf[i]) * dt;
    const double invsurf = c.g.invsurf[i];
    const int SG = AZ_NQ - 1; // index of Sigma
    // phase a: bases.  W[q] = Q[q]/Sigma for the momenta/energy, W[SG] = Sigma (base of Sigma*)
    for (int t = lo + threadIdx.x; t < hi; t += blockDim.x) {
	const double s = Q[SG * L + t];
#pragma unroll
	for (int q = 0; q < SG; ++q)
	    W[q * L + t] = Q[q * L + t] / s;
	W[SG * L + t] = s;
    }
    __syncthreads();
    // phase b: limited slopes (ComputeStarTheta :423-442)
    for (int t = lo + 1 + threadIdx.x; t < hi - 1; t += blockDim.x) {
#pragma unroll
	for (int q = 0; q < AZ_NQ; ++q) {
	    const double w0 = W[q * L + t];
	    const double dqm = (w0 - W[q * L + t - 1]);
	    const double dqp = (W[q * L + t + 1] - w0);
	    D[q * L + t] = 0.5 * flux_limiter<LIM>(dqp, dqm) * invdxtheta;
	}
    }
    __syncthreads();
    // phase c: star values and interface fluxes G(t) (ComputeStarTheta :444-465, VanLeerTheta :655-658).
    // A thread owns at most two cells (L <= 2*blockDim + 3 is checked on the host); results are held in
    // registers until every thread has finished reading W and D, then overwrite W.
    double GA[AZ_NQ], GB[AZ_NQ];
    const int tA = lo + 2 + threadIdx.x, tB = tA + blockDim.x;
    auto flux_at = [&](const int t, double(&G)[AZ_NQ]) {
	const double u = uniform_u ? u_ring : U[t];
	const double ksi = u * dt;
	double star[AZ_NQ];
	if (ksi > 0.0) {
#pragma unroll
	    for (int q = 0; q < AZ_NQ; ++q)
		star[q] = W[q * L + t - 1] + (dxtheta - ksi) * D[q * L + t - 1];
	} else {
#pragma unroll
	    for (int q = 0; q < AZ_NQ; ++q)
		star[q] = W[q * L + t] - (dxtheta + ksi) * D[q * L + t];
	}
#pragma unroll
	for (int q = 0; q < SG; ++q)
	    G[q] = dxrad * star[q] * star[SG] * u;
	G[SG] = dxrad * 1.0 * star[SG] * u;
    };
    if (tA < hi - 1)
	flux_at(tA, GA);
    if (tB < hi - 1)
	flux_at(tB, GB);
    __syncthreads(); // everyone is done reading W, D
    if (tA < hi - 1) {
#pragma unroll
	for (int q = 0; q < AZ_NQ; ++q)
	    W[q * L + tA] = GA[q];
    }
    if (tB < hi - 1) {
#pragma unroll
	for (int q = 0; q < AZ_NQ; ++q)
	    W[q * L + tB] = GB[q];
    }
    __syncthreads();
    // phase d: conservative update (VanLeerTheta :655-660)
    for (int t = lo + 2 + threadIdx.x; t < hi - 2; t += blockDim.x) {
#pragma unroll
	for (int q = 0; q < AZ_NQ; ++q) {
	    double varq = W[q * L + t];
	    varq -= W[q * L + t + 1];
	    Q[q * L + t] += varq * invsurf;
	}
    }
    __syncthreads();
}

template <int LIM>
__global__ void __launch_bounds__(256)
    k_transport_azimuthal(const DevView c, const double *__restrict__ t_sigma, const double *__restrict__ t_rmp,
			  const double *__restrict__ t_rmm, const double *__restrict__ t_amp,
			  const double *__restrict__ t_amm, const double *__restrict__ t_e, const double *__restrict__ vp_old,
			  const double *__restrict__ vr_old, const double *__restrict__ vmean, const int *__restrict__ nshift,
			  const double *__restrict__ vconst, double *__restrict__ o_sigma, double *__restrict__ o_vr,
			  double *__restrict__ o_vp, double *__restrict__ o_e, const double dt, const int S, const int R)
{
    extern __shared__ double smem[];
    const int L = S + AZ_HL + AZ_HR;
    double *Q = smem;		      // [AZ_NQ][L]
    double *U = Q + AZ_NQ * L;	      // [L] residual velocity
    double *W = U + L;		      // [AZ_NQ][L]
    double *D = W + AZ_NQ * L;	      // [AZ_NQ][L]
    double *PS = D + AZ_NQ * L;	      // [S] Sigma of the previous ring (transported, pre-floor)
    double *PR = PS + S;	      // [S] rm+ of the previous ring (transported)
    const int ns = c.ns, nr = c.nr;
    const int j0 = blockIdx.x * S;
    const int seg = min(S, ns - j0); // columns this block writes
    if (seg <= 0)
	return;
    const int i_first = blockIdx.y * R;
    const int i_last = min(i_first + R, nr);
    if (i_first >= nr)
	return;
    const bool adiabatic = c.p.adiabatic != 0;
    const double OmegaF = c.b.omega_frame;
    const double floorv = c.p.sigma_floor * c.p.sigma0;
    const int SG = AZ_NQ - 1;

    for (int i = max(i_first - 1, 0); i < i_last; ++i) {
	const bool store = (i >= i_first);
	const int nsh = nshift[i];
	const double vm = vmean[i];
	const double vc = vconst[i];
	// load the rotated segment: local t <-> output column j0 + t - AZ_HL <-> pre-shift column (that - Nshift)
	for (int t = threadIdx.x; t < L; t += blockDim.x) {
	    long long col = (long long)j0 + t - AZ_HL - nsh;
	    col %= ns;
	    if (col < 0)
		col += ns;
	    const size_t a = (size_t)i * ns + (size_t)col;
	    Q[0 * L + t] = t_rmp[a];
	    Q[1 * L + t] = t_rmm[a];
	    Q[2 * L + t] = t_amp[a];
	    Q[3 * L + t] = t_amm[a];
	    Q[4 * L + t] = adiabatic ? t_e[a] : 0.0;
	    Q[SG * L + t] = t_sigma[a];
	    double u = vp_old[a] - vm; // compute_residual_velocity :194-205
	    if (!c.p.fast_transport)
		u = vc + u; // ComputeConstantResidual :225-231
	    U[t] = u;
	}
	__syncthreads();
	// pass 1: residual velocity, valid [0, L) -> [2, L-2)
	az_pass<LIM>(c, Q, U, false, 0.0, W, D, L, 0, L, i, dt);
	// pass 2: constant residual velocity (skipped for standard transport, :646), valid [2, L-2) -> [4, L-4)
	if (c.p.fast_transport)
	    az_pass<LIM>(c, Q, U, true, vc, W, D, L, 2, L - 2, i, dt);
	// velocities from momenta (:498-535), floors (:123-131), store
	if (store) {
	    for (int t = threadIdx.x; t < seg; t += blockDim.x) {
		const int tl = t + AZ_HL;
		const size_t a = (size_t)i * ns + (size_t)(j0 + t);
		const double s = Q[SG * L + tl], sm = Q[SG * L + tl - 1];
		double vrn;
		if (i == 0)
		    vrn = 0.0;
		else
		    vrn = (PR[t] + Q[1 * L + tl]) / (PS[t] + s);
		const double vpn = (Q[2 * L + tl - 1] + Q[3 * L + tl]) / (sm + s) * c.g.invrmed[i] - c.g.rmed[i] * OmegaF;
		o_vr[a] = vrn;
		o_vp[a] = vpn;
		const double sf = (s < floorv) ? floorv : s;
		o_sigma[a] = sf;
		if (adiabatic)
		    o_e[a] = temperature_clamp(c, sf, Q[4 * L + tl]);
	    }
	}
	__syncthreads();
	// keep this ring's transported Sigma and rm+ for the next ring's v_rad
	for (int t = threadIdx.x; t < seg; t += blockDim.x) {
	    PS[t] = Q[SG * L + t + AZ_HL];
	    PR[t] = Q[0 * L + t + AZ_HL];
	}
	__syncthreads();
    }
    // v_rad ring nr is not touched by compute_velocities_from_momenta (:502-507): carry it over
    if (i_last == nr)
	for (int t = threadIdx.x; t < seg; t += blockDim.x)
	    o_vr[(size_t)nr * ns + j0 + t] = vr_old[(size_t)nr * ns + j0 + t];
}
