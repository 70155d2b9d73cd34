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

__global__ void __launch_bounds__(256)
    k_boundary(const DevView c, double *__restrict__ sigma, double *__restrict__ energy, double *__restrict__ vr,
	       double *__restrict__ vp, const double *__restrict__ sigma0, const double *__restrict__ energy0,
	       const double *__restrict__ vr0, const double *__restrict__ vp0, const double vkep_inner,
	       const double vkep_outer)
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
	}
	if (last) {
	    if (c.p.bc_vazi[1] == FARGO_BC_KEPLERIAN) // keplerian_azimuthal.cpp:41-60
		AT(vp, Irad, j) = vkep_outer;
	    else if (c.p.bc_vazi[1] == FARGO_BC_ZEROGRADIENT)
		AT(vp, Irad, j) = AT(vp, Irad - 1, j);
	    else if (c.p.bc_vazi[1] == FARGO_BC_REFERENCE)
		AT(vp, Irad, j) = AT(vp0, Irad, j);
	}
    }
}

// ---------------------------------------------------------------------------------------------
// damping zones (damping.cpp:311-752).  The per-ring factors exp(-dt*f/tau) are computed on the host with
// glibc (bit-identical to the reference) and uploaded; `ring_lo..ring_hi` is the ring range of one zone.
// type: FARGO_DAMP_INITIAL (X0 = x0 field), _ZERO (X0 = x0_const), _MEAN (X0 = ring mean, index-ordered sum).
__global__ void __launch_bounds__(256)
    k_damping(const DevView c, double *__restrict__ x, const double *__restrict__ x0, const double *__restrict__ expf,
	      const int ring_lo, const int ring_hi, const int type, const double x0_const)
{
    const int ring = ring_lo + blockIdx.y;
    if (ring >= ring_hi)
	return;
    __shared__ double mean_sh;
    if (type == FARGO_DAMP_MEAN) {
	// launched with ONE block per ring: thread 0 does the index-ordered sum (damping.cpp:578-585) before
	// anyone modifies the ring
	if (threadIdx.x == 0) {
	    double s = 0.0;
	    for (int j = 0; j < c.ns; ++j)
		s += AT(x, ring, j);
	    mean_sh = s / c.ns;
	}
	__syncthreads();
    }
    const double mean = (type == FARGO_DAMP_MEAN) ? mean_sh : 0.0;
    const double ef = expf[ring];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < c.ns; j += gridDim.x * blockDim.x) {
	const double X = AT(x, ring, j);
	const double X0 = (type == FARGO_DAMP_INITIAL) ? AT(x0, ring, j) : (type == FARGO_DAMP_MEAN ? mean : x0_const);
	AT(x, ring, j) = (X - X0) * ef + X0;
    }
}

// ---------------------------------------------------------------------------------------------
// cfl::condition_cfl (cfl.cpp:185-382).  Per-cell limits are recomputed from the state (c_s, nu are never
// stored); block min via warp shuffles, then ONE atomicMin per block on the bit pattern of the (positive)
// double — a single-pass grid reduction.
__device__ __forceinline__ void atomic_min_pos_double(double *addr, double v)
{
    // for non-negative IEEE doubles the unsigned bit pattern is monotone in the value
    atomicMin(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

__global__ void __launch_bounds__(256)
    k_cfl(const DevView c, const double *__restrict__ sigma, const double *__restrict__ energy,
	  const double *__restrict__ vr, const double *__restrict__ vp, const double *__restrict__ qplus,
	  const double *__restrict__ qminus, const double *__restrict__ cf_r, const double *__restrict__ cf_phi,
	  const double *__restrict__ vmean, double *__restrict__ dt_out)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int nact = c.active_size - c.first_active;
    const double CFL = c.p.cfl;
    double best = 1.7976931348623157e308;
    if (gid < (long long)nact * c.ns) {
	const int i = c.first_active + (int)(gid / c.ns);
	const int j = (int)(gid - (long long)(i - c.first_active) * c.ns);
	const int jp = (j == c.ns - 1) ? 0 : j + 1;
	const double lf = c.p.leapfrog ? 0.6 : 1.0;
	const double C = c.p.artificial_viscosity_factor;
	if (j == 0) { // FARGO shear criterion (:207-220); the (0,1) pair is the reference's initial dt_core
	    const double denom = fabs(vmean[i] * c.g.invrmed[i] - vmean[i + 1] * c.g.invrmed[i + 1]) + 1.0e-100;
	    best = CFL * c.dphi / denom;
	    if (i == c.first_active) {
		const double denom0 = fabs(vmean[0] * c.g.invrmed[0] - vmean[1] * c.g.invrmed[1]) + 1.0e-100;
		const double d0 = CFL * c.dphi / denom0;
		if (d0 < best)
		    best = d0;
	    }
	}
	const double dxRadial = c.g.rsup[i] - c.g.rinf[i];
	const double dxAzimuthal = c.g.rmed[i] * c.dphi;
	const double cell_size = stdmin(dxRadial, dxAzimuthal);
	const double s = AT(sigma, i, j), e = AT(energy, i, j);
	const double vr0 = AT(vr, i, j), vr1 = AT(vr, i + 1, j), vp0 = AT(vp, i, j), vp1 = AT(vp, i, jp);
	const double vres = c.p.fast_transport ? vp0 - vmean[i] : vp0;
	const double invdt1 = eos_cs(c, i, s, e) / cell_size;
	const double invdt2 = vr0 / dxRadial;
	const double invdt3 = vres / dxAzimuthal;
	double invdt4;
	if (c.p.artificial_viscosity == FARGO_ARTVISC_SN) {
	    double dvRadial = vr1 - vr0;
	    double dvAzimuthal = vp1 - vp0;
	    dvRadial = (dvRadial > 0.0) ? 0.0 : -dvRadial;
	    dvAzimuthal = (dvAzimuthal > 0.0) ? 0.0 : -dvAzimuthal;
	    invdt4 = 4.0 * (C * C) * stdmax(dvRadial / dxRadial, dvAzimuthal / dxAzimuthal) * lf;
	} else { // TW form, also for ArtificialViscosity: None
	    const double eps_rr = (vr1 - vr0) * c.g.invdiffrsup[i];
	    const double eps_pp = c.g.invrmed[i] * ((vp1 - vp0) * c.invdphi + 0.5 * (vr1 + vr0));
	    const double mdiv_V = -stdmin(eps_rr + eps_pp, 0.0);
	    invdt4 = 4.0 * (C * C) * mdiv_V * lf;
	}
	const double invdt5 = 4.0 * eos_nu(c, i, s, e) / (cell_size * cell_size) * lf;
	double invdt6 = 0.0;
	if (c.p.adiabatic) {
	    const double inv_limit = 1.0 / c.p.heating_cooling_cfl_limit;
	    invdt6 = inv_limit * fabs((AT(qplus, i, j) - AT(qminus, i, j)) / e) * lf;
	}
	double dt_cell =
	    CFL / sqrt(invdt1 * invdt1 + invdt2 * invdt2 + invdt3 * invdt3 + invdt4 * invdt4 + invdt5 * invdt5 + invdt6 * invdt6);
	if (c.p.stabilize_viscosity == 2) {
	    const double cc = stdmin(AT(cf_phi, i, j), AT(cf_r, i, j));
	    if (cc != 0.0)
		dt_cell = stdmin(dt_cell, -CFL / cc);
	}
	if (dt_cell < best)
	    best = dt_cell;
    }
    // block reduction
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
	const double other = __shfl_xor_sync(0xffffffffu, best, o);
	if (other < best)
	    best = other;
    }
    __shared__ double wmin[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0)
	wmin[w] = best;
    __syncthreads();
    if (w == 0) {
	best = (lane < (blockDim.x >> 5)) ? wmin[lane] : 1.7976931348623157e308;
#pragma unroll
	for (int o = 4; o > 0; o >>= 1) {
	    const double other = __shfl_xor_sync(0xffffffffu, best, o);
	    if (other < best)
		best = other;
	}
	if (lane == 0)
	    atomic_min_pos_double(dt_out, best);
    }
}

// derived fields on demand (downloads only): T, P, c_s, H, nu  (SourceEuler.cpp:957-1408, viscosity.cpp:98)
__global__ void __launch_bounds__(256) k_derived_field(const DevView c, const double *__restrict__ sigma,
							const double *__restrict__ energy, double *__restrict__ out,
							const int which)
{
    CELL_INDEX(c.nr);
    const double s = AT(sigma, i, j), e = AT(energy, i, j);
    double v = 0.0;
    switch (which) {
    case FARGO_TEMPERATURE:
	if (c.p.adiabatic)
	    v = c.p.mu / c.p.Rgas * (c.p.gamma - 1.0) * e / s;
	else
	    v = c.p.mu / c.p.Rgas * eos_P(c, i, s, e) / s;
	break;
    case FARGO_PRESSURE:
	v = eos_P(c, i, s, e);
	break;
    case FARGO_SOUNDSPEED:
	v = eos_cs(c, i, s, e);
	break;
    case FARGO_SCALE_HEIGHT:
	v = eos_H(c, i, eos_cs(c, i, s, e));
	break;
    case FARGO_VISCOSITY:
	v = eos_nu(c, i, s, e);
	break;
    }
    AT(out, i, j) = v;
}
