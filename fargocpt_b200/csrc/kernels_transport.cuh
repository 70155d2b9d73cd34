// kernels_transport.cuh — radial half of Transport (TransportEuler.cpp:112-167, 349-406, 471-493, 545-620).
//
//   k_transport_radial     momenta from velocities (:471-493) + the six VanLeerRadial sweeps (:138-167,
//                          :349-406, :545-620) in ONE pass: a thread owns one azimuthal column and marches
//                          outward in radius with a register sliding window; Sigma* and Sigma_int are shared by
//                          all transported quantities; every interface flux is evaluated once.
// The ring means live in kernels_ring.cuh, the azimuthal half in kernels_azimuthal.cuh.
#pragma once
#include "fargo_dev.h"

// ---------------------------------------------------------------------------------------------
// Radial sweep.  Quantity order (the reference's, Sigma last): rm+, rm-, am+, am-, e, Sigma.
// State per transported base b (w = Q / Sigma_int for the five non-density quantities, Sigma itself for
// Sigma*): b(k-2), b(k-1), dq(k-2).  When ring k arrives: dq(k-1), star(k-1) -> interface flux F(k-1), then
// cell k-2 is finished:  Q(k-2) += (F(k-2) - F(k-1)) * InvSurf[k-2].
// The 5 + 6 divisions of an iteration (Q / Sigma_int, van Leer slopes) run straight-line on the branch-free
// arithmetic of fargo_math.h with one validity test; upwinding is done by selects.
#ifndef TR_DEPTH
#define TR_DEPTH 2 // rings in flight through shared memory per thread (0: direct loads + L2 prefetch)
#endif
#ifdef TR_MINB
#define TR_BOUNDS __launch_bounds__(128, TR_MINB)
#else
#define TR_BOUNDS __launch_bounds__(128)
#endif
// MF >= 1 (write_disk_quantities, TransportEuler.cpp:578-608): the mass that crosses the mesh's inner boundary (interface 1) and
// outer boundary (interface nr - 1) in this step is added, column by column, to bflow[4][ns] = inner inflow, inner outflow, outer
// inflow, outer outflow.  MF == 2 (WriteMassFlow, :610-616): the mass through the inner interface of every cell n, F2[0], is also
// added to massflow[n].  Separate instantiations: the kernel of a run that asks for neither is unchanged.
template <int LIM, bool ADIABATIC, int MF = 0>
__global__ void TR_BOUNDS
    k_transport_radial(const DevView c, const double *__restrict__ sigma, const double *__restrict__ vr,
		       const double *__restrict__ vp, const double *__restrict__ energy, double *__restrict__ o_sigma,
		       double *__restrict__ o_rmp, double *__restrict__ o_rmm, double *__restrict__ o_amp,
		       double *__restrict__ o_amm, double *__restrict__ o_e, const double dt, const int chunk,
		       double *__restrict__ massflow = nullptr, double *__restrict__ bflow = nullptr)
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
    // rings k-1, k-2 had all their bases inside the validity window R (fargo_math.h).  The slopes that involve a ring
    // below kstart (never loaded) only feed warm-up interfaces whose fluxes reach no stored cell, so "true" is safe.
    bool ok1 = true, ok2 = true;
#pragma unroll
    for (int q = 0; q < NB; ++q) {
	b1[q] = b2[q] = dq2[q] = raw1[q] = raw2[q] = F2[q] = 0.0;
    }

    const int kstart = max(i0 - 2, 0);
    double vnext = AT(vr, kstart, j); // v_r(k), carried from the previous iteration's v_r(k+1)
#if TR_DEPTH > 0
    // The inputs of ring k + TR_DEPTH are on their way into this thread's shared-memory slots while ring k is computed
    // (fargo_dev.h:cp_async_8).  With only an L2 prefetch the first consumer of every ring (the reciprocal of Sigma) waited
    // out an L2 hit, 31 % of all stall samples of this kernel (profiles/r02_v1_ncu_full_c5_8192x16384.md).
    __shared__ double stg[TR_DEPTH + 1][5][128];
    const int tid = threadIdx.x;
    const int klast = min(i1 + 1, nr - 1); // last ring this march loads
    auto stage_ring = [&](const int kk, const int slot) {
	if (kk <= klast) {
	    cp_async_8(&stg[slot][0][tid], &AT(sigma, kk, j));
	    cp_async_8(&stg[slot][1][tid], &AT(vr, kk + 1, j));
	    cp_async_8(&stg[slot][2][tid], &AT(vp, kk, j));
	    cp_async_8(&stg[slot][3][tid], &AT(vp, kk, jp));
	    if (ADIABATIC)
		cp_async_8(&stg[slot][4][tid], &AT(energy, kk, j));
	}
	cp_async_commit(); // (an empty group past the last ring keeps the group count uniform)
    };
#pragma unroll
    for (int d = 0; d < TR_DEPTH; ++d)
	stage_ring(kstart + d, d);
    int slot = 0, slot_in = TR_DEPTH; // slot of ring k, slot ring k + TR_DEPTH goes into
#endif
    // (unrolling the march six-fold over the rotation phases of the window removes the register copies at the end of
    // the loop body but was not faster on B200: 3.36 vs 3.28 ms at 8192x16384)
    for (int k = kstart; k <= i1 + 1; ++k) {
	double bk[NB], rawk[NB], dq1[NB], F1[NB];
	const double vk = vnext;
#if TR_DEPTH > 0
	stage_ring(k + TR_DEPTH, slot_in);
	slot_in = (slot_in == TR_DEPTH) ? 0 : slot_in + 1;
	cp_async_wait<TR_DEPTH>(); // ring k has landed
#else
	if (k + 1 < nr && k < i1 + 1) { // prefetch the next ring
	    pf_global(&AT(sigma, k + 1, j));
	    pf_global(&AT(vr, k + 2, j));
	    pf_global(&AT(vp, k + 1, j));
	    if (ADIABATIC)
		pf_global(&AT(energy, k + 1, j));
	}
#endif
	const int m = k - 1;
	const bool slopes = (m >= 1 && m < nr - 1 && k < nr);
	if (k < nr) {
#if TR_DEPTH > 0
	    const double s = stg[slot][0][tid];
	    const double vk1 = stg[slot][1][tid];
	    vnext = vk1;
	    const double vpj = stg[slot][2][tid], vpn = stg[slot][3][tid];
#else
	    const double s = AT(sigma, k, j);
	    const double vk1 = AT(vr, k + 1, j);
	    vnext = vk1;
	    const double vpj = AT(vp, k, j), vpn = AT(vp, k, jp);
#endif
	    const double r = c.g.rmed[k];
	    rawk[0] = s;
	    rawk[1] = s * vk1;			 // radial_momentum_plus  (:484)
	    rawk[2] = s * vk;			 // radial_momentum_minus (:485)
	    rawk[3] = s * (vpn + r * OmegaF) * r; // angular_momentum_plus (:489)
	    rawk[4] = s * (vpj + r * OmegaF) * r; // angular_momentum_minus(:490)
	    if (ADIABATIC) {
#if TR_DEPTH > 0
		rawk[5] = stg[slot][4][tid];
#else
		rawk[5] = AT(energy, k, j);
#endif
	    }
	    // divise_polargrid (SideEuler.cpp:27-43; Sigma_int == Sigma before the sweep) and the limited slopes of ring
	    // k-1 (compute_star_radial :356-371): 5 + 6 divisions, straight-line with one validity test (fargo_math.h)
	    const double idm = slopes ? c.g.invdiffrmed[m] : 0.0, idp = slopes ? c.g.invdiffrmed[m + 1] : 0.0;
	    FmAcc acc;
	    if (!c.limiter_geo_ok)
		acc.m = 0xffffffffu; // radial spacing outside [2^-30, 2^30]: plain operators (fargo_dev.h:limiter_nb)
	    bk[0] = s;
	    {
		const double ys = fm_rcp_raw(s);
		fm_acc_nrm(acc, s);
#pragma unroll
		for (int q = 1; q < NB; ++q) {
		    bk[q] = fm_div_raw(rawk[q], s, ys);
		    fm_acc_nrm(acc, bk[q]);
		}
	    }
	    if (slopes) {
#pragma unroll
		for (int q = 0; q < NB; ++q) {
		    const double dqm = (b1[q] - b2[q]) * idm;
		    const double dqp = (bk[q] - b1[q]) * idp;
		    dq1[q] = limiter_nb<LIM>(dqp, dqm); // key-free: b1, b2, bk are keyed, the geometry is checked by the host
		}
	    } else {
#pragma unroll
		for (int q = 0; q < NB; ++q)
		    dq1[q] = 0.0;
	    }
	    // the key-free limiter needs the bases of rings k, k-1 and k-2 inside R: remember the last two verdicts
	    const bool okk = fm_acc_ok(acc);
	    const bool fast = okk && ok1 && ok2;
	    ok2 = ok1;
	    ok1 = okk;
	    if (!fast) { // cold: exact zeros (v_rad == 0 at the boundaries) or extreme exponents
#pragma unroll
		for (int q = 1; q < NB; ++q)
		    bk[q] = rawk[q] / s;
		if (slopes) {
#pragma unroll
		    for (int q = 0; q < NB; ++q) {
			const double dqm = (b1[q] - b2[q]) * idm;
			const double dqp = (bk[q] - b1[q]) * idp;
			dq1[q] = flux_limiter<LIM>(dqp, dqm);
		    }
		}
	    }
	} else {
#pragma unroll
	    for (int q = 0; q < NB; ++q)
		bk[q] = rawk[q] = dq1[q] = 0.0;
	    ok2 = ok1;
	    ok1 = false;
	    if (k == nr)
		vnext = 0.0;
	}
	// star values and fluxes through interface m = k-1 (:376-392, :569-575); upwinding by selects
	if (m >= 1 && m < nr && m >= kstart + 1) {
	    const double geo = dtdphi * c.g.rinf[m];
	    const bool pos = v1 > 0.0;
	    const double hp = (c.g.rmed[m] - c.g.rmed[m - 1] - v1 * dt) * 0.5;
	    const double hm = (c.g.rmed[m + 1] - c.g.rmed[m] + v1 * dt) * 0.5;
	    const double hh = pos ? hp : -hm; // b1 - hm * dq1 == b1 + (-hm) * dq1 exactly
	    double star[NB];
#pragma unroll
	    for (int q = 0; q < NB; ++q)
		star[q] = fm_madd(hh, pos ? dq2[q] : dq1[q], pos ? b2[q] : b1[q]);
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
	    AT(o_sigma, n, j) = fm_madd(F2[0] - F1[0], is, raw2[0]);
	    AT(o_rmp, n, j) = fm_madd(F2[1] - F1[1], is, raw2[1]);
	    AT(o_rmm, n, j) = fm_madd(F2[2] - F1[2], is, raw2[2]);
	    AT(o_amp, n, j) = fm_madd(F2[3] - F1[3], is, raw2[3]);
	    AT(o_amm, n, j) = fm_madd(F2[4] - F1[4], is, raw2[4]);
	    if (ADIABATIC)
		AT(o_e, n, j) = fm_madd(F2[5] - F1[5], is, raw2[5]);
	    if (MF == 2) // (+ varq_sup of the mesh's last ring, :613-615: the flux through interface nr, which is 0)
		AT(massflow, n, j) += F2[0];
	    if (MF >= 1) {
		if (n == 1 && c.rank == 0) { // varq_inf of ring 1 (:586-595)
		    if (F2[0] > 0)
			bflow[j] += F2[0];
		    else
			bflow[c.ns + j] += -F2[0];
		} else if (n == nr - 2 && c.rank == c.nranks - 1) { // varq_sup of ring nr - 2 (:596-606)
		    if (F1[0] > 0)
			bflow[3 * c.ns + j] += F1[0];
		    else
			bflow[2 * c.ns + j] += -F1[0];
		}
	    }
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
#if TR_DEPTH > 0
	slot = (slot == TR_DEPTH) ? 0 : slot + 1;
#endif
    }
#if TR_DEPTH > 0
    cp_async_wait<0>();
#endif
}
