// kernels_source.cuh — source-term stages of the hydro step (everything before Transport):
//   potential  : CalculateNbodyPotential            Pframeforce.cpp:21-86
//   sources    : update_with_sourceterms            SourceEuler.cpp:325-493
//   artvisc    : update_with_artificial_viscosity   viscosity/artificial_viscosity.cpp:11-250
//   viscosity  : recalculate_viscosity + compute_viscous_stress_tensor + update_velocities_with_viscosity
//                SourceEuler.cpp:205-223, viscosity/viscosity.cpp:98-426
//   substep3   : SubStep3 / calculate_qplus / calculate_qminus   SourceEuler.cpp:496-954
//
// One thread per cell, azimuth fastest (coalesced rows); stencil neighbours are re-read through
// L1/L2.  Fields that neighbours read while a stage updates them are double-buffered
// (v_rad / v_azi ping-pong between the A and B buffers), everything else is updated in place.
#pragma once
#include "fargo_dev.h"
#include "kernels_rad.cuh"

#define CELL_INDEX(total_rings)                                              \
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;  \
    if (gid >= (long long)(total_rings) * c.ns)                              \
	return;                                                              \
    const int i = (int)(gid / c.ns);                                         \
    const int j = (int)(gid - (long long)i * c.ns);                          \
    const int jp = (j == c.ns - 1) ? 0 : j + 1;                              \
    const int jm = (j == 0) ? c.ns - 1 : j - 1;                              \
    (void)jp;                                                                \
    (void)jm;

#define AT(arr, ii, jj) (arr)[(size_t)(ii) * c.ns + (jj)]

// ---------------------------------------------------------------------------------------------
// Pframeforce.cpp:44-85.  Smoothing = ThicknessSmoothing * H(cell) (Force.cpp:124-159), H from the
// current (Sigma, e) exactly as the end-of-step recalculate_derived_disk_quantities left it.
// h_in: the scale height recalculate_viscosity stored during the first kick of a leapfrog step (the reference does not
// recompute c_s / H before the second kick, simulation.cpp:364-381); nullptr: H of the current state
__device__ __forceinline__ double potential_at(const DevView &c, int i, int j, double sigma, double energy, const double *h_in = nullptr)
{
    const size_t cell = (size_t)i * c.ns + j;
    // PVTE: the SCALE_HEIGHT grid as stored (the next lookup's input too); else H of the given state
    const double H = h_in ? h_in[cell] : (c.pv.H ? c.pv.H[cell] : eos_H(c, i, eos_cs(c, i, sigma, energy)));
    const double x = c.g.rmed[i] * c.g.cosphi[j];
    const double y = c.g.rmed[i] * c.g.sinphi[j];
    const double smooth = c.p.thickness_smoothing * H;
    double pot = 0.0;
    for (int k = 0; k < c.b.n; ++k) {
	const double dx = x - c.b.x[k];
	const double dy = y - c.b.y[k];
	const double dist_2 = dx * dx + dy * dy;
	const double d_smoothed = sqrt(dist_2 + smooth * smooth);
	double smooth_factor_klahr = 1.0;
	const double r_sm = c.b.cubic_smoothing_radius[k];
	if (r_sm > 0.0 && d_smoothed < r_sm) {
	    const double q = d_smoothed / r_sm;
	    smooth_factor_klahr = (pow(q, 4.0) - 2.0 * pow(q, 3.0) + 2.0 * d_smoothed / r_sm);
	}
	pot += -c.p.G * c.b.mass[k] / d_smoothed * smooth_factor_klahr;
    }
    pot += -c.b.indirect_x * x - c.b.indirect_y * y;
    return pot;
}

// pre: the pre-accretion state the reference's stored SCALE_HEIGHT still describes (fargo_dev.h:PreState)
__global__ void __launch_bounds__(256) k_potential(const DevView c, const double *__restrict__ sigma,
						    const double *__restrict__ energy, double *__restrict__ pot,
						    const double *__restrict__ h_in, const PreState pre)
{
    CELL_INDEX(c.nr);
    const bool old = pre_has(pre, i);
    const double s = old ? AT(pre.sigma, i, j) : AT(sigma, i, j);
    const double e = (old && c.p.adiabatic) ? AT(pre.energy, i, j) : AT(energy, i, j);
    AT(pot, i, j) = potential_at(c, i, j, s, e, h_in);
}

// ---------------------------------------------------------------------------------------------
// momentum_update_radial / momentum_update_azimuthal (SourceEuler.cpp:325-428).
// Reads the A buffers of v, writes the B buffers (all rings, untouched rings are copied).
__global__ void __launch_bounds__(256)
    k_sources_velocity(const DevView c, const double *__restrict__ sigma, const double *__restrict__ energy,
		       const double *__restrict__ pot, const double *__restrict__ vr, const double *__restrict__ vp,
		       double *__restrict__ vr_out, double *__restrict__ vp_out, const double dt, const PreState pre)
{
    CELL_INDEX(c.nr + 1);
    // the stored PRESSURE of the reference: of the pre-accretion state where accretion has touched the ring (PreState);
    // the densities in the denominators are the current ones (SourceEuler.cpp:349-353, :401-405)
    auto pressure = [&](int ii, int jj) {
	const bool old = pre_has(pre, ii);
	const double ps = old ? AT(pre.sigma, ii, jj) : AT(sigma, ii, jj);
	const double pe = (old && c.p.adiabatic) ? AT(pre.energy, ii, jj) : AT(energy, ii, jj);
	return eos_P_at(c, ii, (size_t)ii * c.ns + jj, ps, pe);
    };
    double vr_new = AT(vr, i, j);
    if (i >= c.one_no_ghost_vr && i < c.maxmo_no_ghost_vr) {
	const double s = AT(sigma, i, j), sm = AT(sigma, i - 1, j);
	const double P = pressure(i, j);
	const double Pm = pressure(i - 1, j);
	double gradp = 2.0 / (s + sm);
	gradp *= (P - Pm);
	gradp *= c.g.invdiffrmed[i];
	const double gradphi = (AT(pot, i, j) - AT(pot, i - 1, j)) * c.g.invdiffrmed[i];
	const double vsum = AT(vp, i, j) + AT(vp, i, jp) + AT(vp, i - 1, j) + AT(vp, i - 1, jp);
	const double vt = 0.25 * vsum + c.g.rinf[i] * c.b.omega_frame;
	const double vt2 = vt * vt;
	const double centrifugal_accel = vt2 * c.g.invrinf[i];
	vr_new += dt * (-gradp - gradphi + centrifugal_accel);
    }
    AT(vr_out, i, j) = vr_new;
    if (i < c.nr) {
	double vp_new = AT(vp, i, j);
	if (i >= c.zero_no_ghost && i < c.max_no_ghost) {
	    const double invdxtheta = 2.0 / (c.dphi * (c.g.rsup[i] + c.g.rinf[i]));
	    const double s = AT(sigma, i, j), sp = AT(sigma, i, jm);
	    const double P = pressure(i, j);
	    const double Pp = pressure(i, jm);
	    const double gradp = 2.0 / (s + sp) * (P - Pp) * invdxtheta;
	    const double gradphi = (AT(pot, i, j) - AT(pot, i, jm)) * invdxtheta;
	    vp_new = vp_new + dt * (-gradp - gradphi);
	    if (c.p.imposed_disk_drift != 0.0)
		vp_new += dt * c.g.supp_torque[i];
	}
	AT(vp_out, i, j) = vp_new;
    }
}

// correct_v_azimuthal (SideEuler.cpp:79-95): the frame's angular velocity changed by dOmega
__global__ void __launch_bounds__(256) k_correct_vazi(const DevView c, double *__restrict__ vp, const double domega)
{
    CELL_INDEX(c.nr);
    AT(vp, i, j) -= domega * c.g.rmed[i];
}

// div(v) as used by compression_heating and the viscous stress tensor (SourceEuler.cpp:471-477,
// viscosity.cpp:154-159)
__device__ __forceinline__ double div_v(const DevView &c, const double *__restrict__ vr, const double *__restrict__ vp,
					 int i, int j, int jp)
{
    return (AT(vr, i + 1, j) * c.g.rinf[i + 1] - AT(vr, i, j) * c.g.rinf[i]) * c.g.invdiffrsuprb[i] +
	   (AT(vp, i, jp) - AT(vp, i, j)) * c.invdphi * c.g.invrmed[i];
}

// compression_heating (SourceEuler.cpp:459-493); uses the UPDATED velocities.
__global__ void __launch_bounds__(256) k_compression_heating(const DevView c, const double *__restrict__ vr,
							      const double *__restrict__ vp, double *__restrict__ energy,
							      const double dt)
{
    CELL_INDEX(c.nr - 1);
    const double DIV_V = div_v(c, vr, vp, i, j, jp);
    const double e_old = AT(energy, i, j);
    AT(energy, i, j) = e_old * exp_ref(-(pv_geff(c, (size_t)i * c.ns + j) - 1.0) * dt * DIV_V); // glibc-exact exp (fargo_math.h)
}

// ---------------------------------------------------------------------------------------------
// artificial viscosity, pass 1: Q_rr / Q_phiphi (+ dissipation into e, + temperature floor)
// TW: artificial_viscosity.cpp:49-88; SN: :165-218; floor: :19-21
__global__ void __launch_bounds__(256)
    k_artvisc_q(const DevView c, const double *__restrict__ sigma, const double *__restrict__ vr,
		const double *__restrict__ vp, double *__restrict__ energy, double *__restrict__ qr,
		double *__restrict__ qphi, const double dt)
{
    CELL_INDEX(c.nr);
    const double C = c.p.artificial_viscosity_factor;
    const double s = AT(sigma, i, j);
    const bool diss = c.p.adiabatic && c.p.artificial_viscosity_dissipation;
    double e = diss ? AT(energy, i, j) : 0.0;
    if (c.p.artificial_viscosity == FARGO_ARTVISC_TW) {
	const double vr0 = AT(vr, i, j), vr1 = AT(vr, i + 1, j);
	const double eps_rr = (vr1 - vr0) * c.g.invdiffrsup[i];
	const double eps_pp = c.g.invrmed[i] * ((AT(vp, i, jp) - AT(vp, i, j)) * c.invdphi + 0.5 * (vr1 + vr0));
	const double div_V = stdmin(eps_rr + eps_pp, 0.0);
	const double Dr = c.g.rinf[i + 1] - c.g.rinf[i];
	const double rDphi = c.g.rmed[i] * c.dphi;
	const double m = (c.ns <= 16) ? stdmin(Dr, rDphi) : stdmax(Dr, rDphi);
	const double dx_sq = m * m;
	const double l_sq = (C * C) * dx_sq;
	AT(qr, i, j) = l_sq * s * -div_V * (eps_rr - 1.0 / 3.0 * div_V);
	AT(qphi, i, j) = l_sq * s * -div_V * (eps_pp - 1.0 / 3.0 * div_V);
	if (diss && i > c.zero_no_ghost && i < c.max_no_ghost) {
	    const double Qplus =
		-l_sq * div_V * s * 1.0 / 3.0 * (eps_rr * eps_rr + eps_pp * eps_pp + (eps_rr - eps_pp) * (eps_rr - eps_pp));
	    e += Qplus * dt;
	}
    } else if (c.p.artificial_viscosity == FARGO_ARTVISC_SN) {
	const double dv_r = AT(vr, i + 1, j) - AT(vr, i, j);
	const double q_r = (dv_r < 0.0) ? (C * C) * s * (dv_r * dv_r) : 0.0;
	const double dv_phi = AT(vp, i, jp) - AT(vp, i, j);
	const double q_p = (dv_phi < 0.0) ? (C * C) * s * (dv_phi * dv_phi) : 0.0;
	AT(qr, i, j) = q_r;
	AT(qphi, i, j) = q_p;
	if (diss && i >= c.zero_no_ghost && i < c.max_no_ghost) {
	    const double dxtheta = c.dphi * c.g.rmed[i];
	    const double invdxtheta = 1.0 / dxtheta;
	    e = e - dt * q_r * dv_r * c.g.invdiffrsup[i] - dt * q_p * dv_phi * invdxtheta;
	}
    }
    if (diss)
	AT(energy, i, j) = temperature_clamp_at(c, (size_t)i * c.ns + j, s, e);
}

// artificial viscosity, pass 2: velocity update from Q (in place: every thread touches only its own v)
// TW: artificial_viscosity.cpp:90-139; SN: :221-248
__global__ void __launch_bounds__(256)
    k_artvisc_v(const DevView c, const double *__restrict__ sigma, const double *__restrict__ qr,
		const double *__restrict__ qphi, double *__restrict__ vr, double *__restrict__ vp, const double dt)
{
    CELL_INDEX(c.nr);
    const double s = AT(sigma, i, j);
    if (c.p.artificial_viscosity == FARGO_ARTVISC_TW) {
	if (i >= 1 && i < c.nr - 1) {
	    const double sigma_phi_avg = 0.5 * (s + AT(sigma, i, jm));
	    const double dVp =
		2.0 * dt / ((c.g.rsup[i] + c.g.rinf[i]) * sigma_phi_avg) * (AT(qphi, i, j) - AT(qphi, i, jm)) * c.invdphi;
	    AT(vp, i, j) += dVp;
	}
	if (i >= c.one_no_ghost_vr && i < c.maxmo_no_ghost_vr) {
	    const double sigma_r_avg = 0.5 * (s + AT(sigma, i - 1, j));
	    const double rm = c.g.rmed[i], rmm = c.g.rmed[i - 1];
	    const double dVr = c.p.radial_viscosity_factor * dt / sigma_r_avg * 2.0 / (rm * rm - rmm * rmm) *
			       ((AT(qr, i, j) * rm - AT(qr, i - 1, j) * rmm) -
				0.5 * (AT(qphi, i, j) + AT(qphi, i - 1, j)) * (rm - rmm));
	    AT(vr, i, j) += dVr;
	}
    } else {
	if (i >= c.one_no_ghost_vr && i < c.maxmo_no_ghost_vr) {
	    AT(vr, i, j) = AT(vr, i, j) - dt * 2.0 / (s + AT(sigma, i - 1, j)) * (AT(qr, i, j) - AT(qr, i - 1, j)) * c.g.invdiffrmed[i];
	}
	if (i >= c.zero_no_ghost && i < c.max_no_ghost) {
	    const double dxtheta = c.dphi * c.g.rmed[i];
	    const double invdxtheta = 1.0 / dxtheta;
	    AT(vp, i, j) = AT(vp, i, j) - dt * 2.0 / (s + AT(sigma, i, jm)) * (AT(qphi, i, j) - AT(qphi, i, jm)) * invdxtheta;
	}
    }
}

// ---------------------------------------------------------------------------------------------
// nu field (viscosity.cpp:98-137) from the current (Sigma, e)
__global__ void __launch_bounds__(256) k_viscosity_nu(const DevView c, const double *__restrict__ sigma,
						       const double *__restrict__ energy, double *__restrict__ nu,
						       double *__restrict__ o_h)
{
    CELL_INDEX(c.nr);
    const size_t cell = (size_t)i * c.ns + j;
    AT(nu, i, j) = eos_nu_at(c, i, cell, AT(sigma, i, j), AT(energy, i, j), /*stored_T=*/true);
    if (o_h) // leapfrog: keep the scale height of this moment for the second kick's potential smoothing
	AT(o_h, i, j) = eos_H_at(c, i, cell, eos_cs_at(c, i, cell, AT(sigma, i, j), AT(energy, i, j)));
}

// compute_viscous_stress_tensor (viscosity.cpp:139-254): div v, tau_rr, tau_phiphi (cell centred), tau_rphi (corner)
__global__ void __launch_bounds__(256)
    k_stress(const DevView c, const double *__restrict__ sigma, const double *__restrict__ nu,
	     const double *__restrict__ vr, const double *__restrict__ vp, double *__restrict__ divv,
	     double *__restrict__ trr, double *__restrict__ tpp, double *__restrict__ trp, double *__restrict__ nusig,
	     double *__restrict__ nusig_rp)
{
    CELL_INDEX(c.nr);
    const double vr0 = AT(vr, i, j), vr1 = AT(vr, i + 1, j);
    const double vp0 = AT(vp, i, j);
    const double s = AT(sigma, i, j), n = AT(nu, i, j);
    const double dv = (vr1 * c.g.rinf[i + 1] - vr0 * c.g.rinf[i]) * c.g.invdiffrsuprb[i] +
		      (AT(vp, i, jp) - vp0) * c.invdphi * c.g.invrmed[i];
    AT(divv, i, j) = dv;
    const double drr = (vr1 - vr0) * c.g.invdiffrsup[i];
    AT(trr, i, j) = 2.0 * n * s * (drr - 1.0 / 3.0 * dv);
    const double dpp = (AT(vp, i, jp) - vp0) * c.invdphi * c.g.invrmed[i] + 0.5 * (vr1 + vr0) * c.g.invrmed[i];
    AT(tpp, i, j) = 2.0 * n * s * (dpp - 1.0 / 3.0 * dv);
    if (c.p.stabilize_viscosity)
	AT(nusig, i, j) = n * s;
    if (i >= 1) {
	const double dvazirdr = (vp0 * c.g.invrmed[i] - AT(vp, i - 1, j) * c.g.invrmed[i - 1]) * c.g.invdiffrmed[i];
	const double dvrdphi = (vr0 - AT(vr, i, jm)) * c.invdphi;
	const double drp = c.g.rinf[i] * dvazirdr + dvrdphi * c.g.invrinf[i];
	const double nua = 0.25 * (n + AT(nu, i - 1, j) + AT(nu, i, jm) + AT(nu, i - 1, jm));
	const double sa = 0.25 * (s + AT(sigma, i - 1, j) + AT(sigma, i, jm) + AT(sigma, i - 1, jm));
	AT(trp, i, j) = nua * sa * drp;
	if (c.p.stabilize_viscosity)
	    AT(nusig_rp, i, j) = nua * sa;
    } else {
	AT(trp, i, j) = 0.0; /* ring 0 is never written by the reference and stays 0 from allocation */
    }
}

// stabilisation factors (viscosity.cpp:256-348), only StabilizeViscosity >= 1
__global__ void __launch_bounds__(256)
    k_stress_correction(const DevView c, const double *__restrict__ sigma, const double *__restrict__ nusig,
			const double *__restrict__ nusig_rp, double *__restrict__ cf_r, double *__restrict__ cf_phi)
{
    CELL_INDEX(c.nr);
    if (i < 1)
	return;
    const double ns_corner = AT(nusig_rp, i, j);
    const double ns_corner_out = AT(nusig_rp, i + 1, j); /* ring nr of the (vector) grid stays 0 */
    const double ns_corner_next = AT(nusig_rp, i, jp);
    const double ns_cell = AT(nusig, i, j);
    const double ns_left = AT(nusig, i, jm);
    const double ns_inner = AT(nusig, i - 1, j);
    const double flux_in = ns_corner * pow(c.g.rinf[i], 3.0) * c.g.invdiffrmed[i];
    const double flux_out = ns_corner_out * pow(c.g.rinf[i + 1], 3.0) * c.g.invdiffrmed[i + 1];
    const double kphi_shear = -c.g.invrmed[i] * c.g.twodiffrasq[i] * (flux_out + flux_in);
    const double kphi_normal = -c.g.fourthird[i] * (ns_cell + ns_left);
    const double s_phi = 0.5 * (AT(sigma, i, j) + AT(sigma, i, jm));
    AT(cf_phi, i, j) = (kphi_shear + kphi_normal) / (s_phi * c.g.rmed[i]);
    const double s_rad = 0.5 * (AT(sigma, i, j) + AT(sigma, i - 1, j));
    const double kr_shear = -(ns_corner_next + ns_corner) / (c.dphi * c.dphi * c.g.rinf[i]);
    const double kr_hoop_a = 2.0 * ns_cell * (0.5 * c.g.invrmed[i] + 1.0 / 3.0 * c.g.rinf[i] * c.g.invdiffrsuprb[i]);
    const double kr_hoop_b = 2.0 * ns_inner * (0.5 * c.g.invrmed[i - 1] - 1.0 / 3.0 * c.g.rinf[i] * c.g.invdiffrsuprb[i - 1]);
    const double kr_norm_a = c.g.rmed[i] * 2.0 * ns_cell * (-c.g.invdiffrsup[i] + 1.0 / 3.0 * c.g.rinf[i] * c.g.invdiffrsuprb[i]);
    const double kr_norm_b =
	-1.0 * c.g.rmed[i - 1] * 2.0 * ns_inner * (c.g.invdiffrsup[i - 1] - 1.0 / 3.0 * c.g.rinf[i] * c.g.invdiffrsuprb[i - 1]);
    const double kr_hoop = -0.5 * (kr_hoop_a + kr_hoop_b);
    const double kr_norm = c.g.invdiffrmed[i] * (kr_norm_a + kr_norm_b);
    const double r_iface = 0.5 * (c.g.rmed[i] + c.g.rmed[i - 1]);
    AT(cf_r, i, j) = c.p.radial_viscosity_factor * (kr_norm + kr_shear + kr_hoop) / (s_rad * r_iface);
}

// update_velocities_with_viscosity (viscosity.cpp:355-426), in place
__global__ void __launch_bounds__(256)
    k_viscosity_v(const DevView c, const double *__restrict__ sigma, const double *__restrict__ trr,
		  const double *__restrict__ tpp, const double *__restrict__ trp, const double *__restrict__ cf_r,
		  const double *__restrict__ cf_phi, double *__restrict__ vr, double *__restrict__ vp, const double dt)
{
    CELL_INDEX(c.nr);
    const double s = AT(sigma, i, j);
    if (i >= 1 && i < c.nr - 1) {
	const double sigma_avg = 0.5 * (s + AT(sigma, i, jm));
	const double ra2 = c.g.rinf[i] * c.g.rinf[i], rap2 = c.g.rinf[i + 1] * c.g.rinf[i + 1];
	double dVp = dt * c.g.invrmed[i] / (sigma_avg) *
		     ((2.0 / (rap2 - ra2)) * (rap2 * AT(trp, i + 1, j) - ra2 * AT(trp, i, j)) + (AT(tpp, i, j) - AT(tpp, i, jm)) * c.invdphi);
	if (c.p.stabilize_viscosity == 1) {
	    const double cphi = AT(cf_phi, i, j);
	    const double corr = 1.0 / (stdmax(1.0 + dt * cphi, 0.0) - dt * cphi);
	    dVp *= corr;
	}
	AT(vp, i, j) += dVp;
    }
    if (i >= c.one_no_ghost_vr && i < c.maxmo_no_ghost_vr) {
	const double sigma_avg = 0.5 * (s + AT(sigma, i - 1, j));
	double dVr = dt / (sigma_avg)*c.p.radial_viscosity_factor * 2.0 / (c.g.rmed[i] + c.g.rmed[i - 1]) *
		     ((c.g.rmed[i] * AT(trr, i, j) - c.g.rmed[i - 1] * AT(trr, i - 1, j)) * c.g.invdiffrmed[i] +
		      (AT(trp, i, jp) - AT(trp, i, j)) * c.invdphi - 0.5 * (AT(tpp, i, j) + AT(tpp, i - 1, j)));
	if (c.p.stabilize_viscosity == 1) {
	    const double cr = AT(cf_r, i, j);
	    const double corr = 1.0 / (stdmax(1.0 + dt * cr, 0.0) - dt * cr);
	    dVr *= corr;
	}
	AT(vr, i, j) += dVr;
    }
}

// ---------------------------------------------------------------------------------------------
// Q+ (viscous_heating, SourceEuler.cpp:496-536) and Q- (thermal_relaxation, :632-690) for one cell.
__device__ __forceinline__ double qplus_cell(const DevView &c, const double *__restrict__ sigma, const double *__restrict__ nu,
					      const double *__restrict__ divv, const double *__restrict__ trr,
					      const double *__restrict__ tpp, const double *__restrict__ trp, int i, int j, int jp)
{
    double q = 0.0;
    if (c.p.heating_viscous && i >= 1 && i < c.nr - 1) {
	const double n = AT(nu, i, j);
	if (n != 0.0) {
	    const double s = AT(sigma, i, j);
	    const double tau_r_phi = 0.25 * (AT(trp, i, j) + AT(trp, i + 1, j) + AT(trp, i, jp) + AT(trp, i + 1, jp));
	    const double t_rr = AT(trr, i, j), t_pp = AT(tpp, i, j), dv = AT(divv, i, j);
	    double qplus = 1.0 / (2.0 * n * s) * (t_rr * t_rr + 2 * (tau_r_phi * tau_r_phi) + t_pp * t_pp);
	    qplus += (2.0 / 9.0) * n * s * (dv * dv);
	    qplus *= c.p.heating_viscous_factor;
	    q += qplus;
	}
    }
    return q;
}

__device__ __forceinline__ double qminus_cell(const DevView &c, double beta_inv, double sigma, double energy, double sigma0,
					       double energy0, int i, size_t cell)
{
    double q = 0.0;
    if (c.p.cooling_beta && i >= 1 && i < c.nr - 1) {
	double delta_E = energy;
	if (c.p.cooling_beta_reference & FARGO_BETA_REF_REFERENCE)
	    delta_E -= energy0 / sigma0 * sigma;
	if (c.p.cooling_beta_reference & FARGO_BETA_REF_MODEL)
	    delta_E -= c.g.beta_model_e0[i] * sigma;
	if (c.p.cooling_beta_reference & FARGO_BETA_REF_FLOOR)
	    delta_E -= c.p.minimum_temperature * sigma / pv_mu(c, cell) * c.p.Rgas / (pv_geff(c, cell) - 1.0);
	q += delta_E * c.g.omega_k[i] * beta_inv;
    }
    return q;
}

// alpha_r of SubStep3 (SourceEuler.cpp:921-924); H from recalculate_viscosity, i.e. from the current (Sigma, e)
__device__ __forceinline__ double radiative_alpha(const DevView &c, int i, size_t cell, double sigma, double energy)
{
    const double cs = eos_cs_at(c, i, cell, sigma, energy);
    const double H = eos_H_at(c, i, cell, cs);
    const double inv_pow4 = pow(pv_mu(c, cell) * (pv_geff(c, cell) - 1.0) / (c.p.Rgas * sigma), 4.0);
    return 1.0 + 2.0 * H * 4.0 * c.p.sigma_sb / c.p.c_light * inv_pow4 * pow(energy, 3.0);
}

// SubStep3 (SourceEuler.cpp:859-954).  update_energy = 0 reproduces compute_heating_cooling_for_CFL (:1410-1450).
__global__ void __launch_bounds__(256)
    k_substep3(const DevView c, const double *__restrict__ sigma, const double *__restrict__ nu,
	       const double *__restrict__ divv, const double *__restrict__ trr, const double *__restrict__ tpp,
	       const double *__restrict__ trp, const double *__restrict__ sigma0, const double *__restrict__ energy0,
	       double *__restrict__ energy, double *__restrict__ qplus, double *__restrict__ qminus, const double dt,
	       const double beta_inv, const int update_energy)
{
    CELL_INDEX(c.nr);
    const double s = AT(sigma, i, j);
    double e = AT(energy, i, j);
    const bool need0 = c.p.cooling_beta && (c.p.cooling_beta_reference & FARGO_BETA_REF_REFERENCE);
    double Qp = qplus_cell(c, sigma, nu, divv, trr, tpp, trp, i, j, jp);
    const size_t cell = (size_t)i * c.ns + j;
    if (c.t_alpha && update_energy) // SubStep3 begins with compute_temperature (SourceEuler.cpp:861): the grid a leapfrog's second
	c.t_alpha[cell] = pv_mu(c, cell) / c.p.Rgas * (pv_geff(c, cell) - 1.0) * e / s; // recalculate_viscosity will read
    double Qm = qminus_cell(c, beta_inv, s, e, need0 ? AT(sigma0, i, j) : 1.0, need0 ? AT(energy0, i, j) : 0.0, i, cell);
    double tau_eff = 0.0; // TAU_EFF stays 0 as allocated unless kappa_eff runs
    if (rad_enabled(c) && i >= 1 && i < c.nr - 1) { // thermal_cooling / irradiation (kernels_rad.cuh)
	const double H = eos_H_at(c, i, cell, eos_cs_at(c, i, cell, s, e));
	const double T = rad_temperature(c, s, e, pv_mu(c, cell), pv_geff(c, cell));
	tau_eff = rad_tau_eff(c, s, H, T);
	if (c.p.cooling_surface)
	    Qm += rad_qminus(c, T, tau_eff);
	if (c.p.heating_star)
	    rad_add_qplus(c, i, c.g.cosphi[j], c.g.sinphi[j], H, tau_eff, Qp);
    }
    if (c.p.cooling_scurve && i >= 1 && i < c.nr - 1) // scurve_cooling (SourceEuler.cpp:726-831), after thermal_cooling in calculate_qminus
	Qm += rad_scurve_qminus(c, i, s, rad_temperature(c, s, e, pv_mu(c, cell), pv_geff(c, cell)), pv_mu(c, cell), tau_eff);
    if (i >= 1 && i < c.nr - 1) {
	const double alpha = radiative_alpha(c, i, cell, s, e);
	Qp /= alpha;
	Qm /= alpha;
	if (update_energy) {
	    double energy_new = e + dt * (Qp - Qm);
	    const double SigmaFloor = 10.0 * c.p.sigma0 * c.p.sigma_floor;
	    if (s < SigmaFloor) {
		const double e4 = Qp * tau_eff / (2.0 * c.p.sigma_sb);
		const double constant = (c.p.Rgas / pv_mu(c, cell) * s / (pv_geff(c, cell) - 1.0));
		const double eq_energy = pow(e4, 1.0 / 4.0) * constant;
		Qm = Qp;
		energy_new = eq_energy;
	    }
	    e = energy_new;
	}
    }
    AT(qplus, i, j) = Qp;
    AT(qminus, i, j) = Qm;
    if (update_energy)
	AT(energy, i, j) = temperature_clamp_at(c, cell, s, e);
}
