// kernels_rad.cuh — radiative terms of SubStep3: opacity-based surface cooling and stellar irradiation.
//
//   thermal_cooling     SourceEuler.cpp:693-723    Q- += f 2 sigma_SB (T^4 - Tmin^4) / tau_eff
//   irradiation_single  SourceEuler.cpp:538-596    Q+ += ramp 2 sigma_SB T_irr^4 / tau_eff   (Menou & Goodman 2004, D'Angelo & Marzari 2012)
//   midplane_density, kappa_eff  compute.cpp:17-88  rho = Sigma / (f_rho H), tau = f_tau kappa Sigma / f_rho, Hubeny's tau_eff
//   opacity::opacity    opacity.cpp:11-298          Constant, Simple (kappa T^2), Lin & Papaloizou (1985), Bell & Lin (1994)
//
// The reference evaluates these from stored grids (TEMPERATURE refreshed at the top of SubStep3, SCALE_HEIGHT by
// midplane_density, TAU_EFF by kappa_eff); here they are functions of the cell's (Sigma, e) evaluated where SubStep3 runs —
// the staged k_substep3 and, behind a template flag, the fused viscosity kernel (the instantiation the BASELINE configs
// launch does not contain them).  Plain IEEE operators in the reference's order: with Opacity: Constant the result is the
// reference's bit for bit (x^4 as a correctly rounded power like glibc's pow(x, 4)); the two opacity tables call pow() with
// fractional exponents, where CUDA's and glibc's results differ in the last bits (fields agree to ~1e-14).
#pragma once
#include "fargo_dev.h"

// the eight power laws of (rho, T) [cgs] of the two opacity tables joined by their smoothing functions
struct RadOpLaw {
    double power1, power2, power3, t234, t456, t678;
    double ak1, ak2, ak3, bk3, bk4, bk5, bk6, bk7, bk8;
};
template <bool BELL> __device__ __forceinline__ RadOpLaw rad_op_law()
{
    if (BELL)
	return RadOpLaw{2.8369e-2, 1.1464e-2, 2.2667e-1, 1.46e3, 4.51e3, 2.37e6, 2.e-4, 2.e16, 0.1, 10., 2.e-15, 1e4, 1e4, 1.5e10, 0.348};
    return RadOpLaw{4.44444444e-2, 2.381e-2, 2.267e-1, 1.6e3, 5.7e3, 2.28e6, 2.e-4, 2.e16, 5.e-3, 50., 2.e-2, 2.e4, 1.e4, 1.5e10, 0.348};
}
template <bool BELL> __device__ double rad_opacity_table(const double rho, double T)
{
    const RadOpLaw L = rad_op_law<BELL>();
    if (BELL && T < 1.0)
	T = 10.0;
    if (!(T > L.t234 * pow(rho, L.power1))) { // ice grains, their evaporation, metal grains: powers of T itself
	const double t2 = T * T, t4 = t2 * t2, t8 = t4 * t4, t10 = t8 * t2;
	const double o1 = L.ak1 * t2, o2 = L.ak2 * T / t8;
	const double o3 = BELL ? L.ak3 * sqrt(T) : L.ak3 * T;
	const double a1 = o1 * o1, a2 = o2 * o2;
	const double u = a1 * a2 / (a1 + a2), w = o3 / (1 + 1.e22 / t10);
	return pow(u * u + fm_pow4(w), 0.25);
    }
    const double ts4 = 1.e-4 * T;
    const double rho13 = pow(rho, 1.0 / 3.0), rho23 = rho13 * rho13;
    const double ts42 = ts4 * ts4, ts44 = ts42 * ts42, ts48 = ts44 * ts44;
    if (!(T > L.t456 * pow(rho, L.power2))) { // metal grains, their evaporation, molecules
	const double o3 = BELL ? L.bk3 * sqrt(ts4) : L.bk3 * ts4;
	const double o4 = BELL ? L.bk4 * rho / (ts48 * ts48 * ts48) : L.bk4 * rho23 / (ts48 * ts4);
	const double o5 = L.bk5 * rho23 * ts42 * ts4;
	const double a4 = fm_pow4(o4), a3 = fm_pow4(o3);
	const double damp = BELL ? (1 + 6.561e-5 / ts48 * 1e2 * rho23) : (1.0 + 6.561e-5 / ts48);
	return pow((a4 * a3 / (a4 + a3)) + fm_pow4(o5 / damp), 0.25);
    }
    const bool below_scattering = BELL ? ((T < L.t678 * pow(rho, L.power3)) || ((rho <= 1e10) && (T < 1e4)))
				       : ((T < L.t678 * pow(rho, L.power3)) || (rho <= 1e-10));
    const double o7 = L.bk7 * rho / (ts42 * sqrt(ts4)), a7 = o7 * o7;
    if (below_scattering) { // molecules, H-, Kramers
	const double o5 = L.bk5 * rho23 * ts42 * ts4;
	const double o6 = L.bk6 * rho13 * ts48 * ts42, a6 = o6 * o6;
	const double u = a6 * a7 / (a6 + a7);
	const double w = o5 / (1.0 + pow(ts4 / (1.1 * pow(rho, 0.04762)), 10.0));
	return pow(u * u + fm_pow4(w), 0.25);
    }
    const double a8 = L.bk8 * L.bk8; // Kramers, electron scattering
    return pow(a7 * a7 + a8 * a8, 0.25);
}
// opacity::opacity (opacity.cpp:11-44), code units in and out
__device__ __forceinline__ double rad_opacity(const DevView &c, const double rho, const double T)
{
    const double Tcgs = T * c.p.temperature_cgs, rhocgs = rho * c.p.density_cgs;
    double rv;
    if (c.p.opacity == FARGO_OPACITY_CONST)
	rv = c.p.kappa_const;
    else if (c.p.opacity == FARGO_OPACITY_SIMPLE)
	rv = c.p.kappa_const * (Tcgs * Tcgs);
    else if (c.p.opacity == FARGO_OPACITY_BELL)
	rv = rad_opacity_table<true>(rhocgs, Tcgs) * c.p.opacity_code;
    else
	rv = rad_opacity_table<false>(rhocgs, Tcgs) * c.p.opacity_code;
    return c.p.kappa_factor * rv;
}
// compute_temperature (SourceEuler.cpp:1378-1408), adiabatic
__device__ __forceinline__ double rad_temperature(const DevView &c, const double sigma, const double energy, const double mu,
						   const double gamma_eff)
{
    const double c_v_inv = mu / c.p.Rgas * (gamma_eff - 1.0);
    return c_v_inv * energy / sigma;
}
// kappa_eff (compute.cpp:41-88) of one cell
__device__ __forceinline__ double rad_tau_eff(const DevView &c, const double sigma, const double H, const double T)
{
    const double rho = sigma / (c.p.density_factor * H);
    const double kappa = rad_opacity(c, rho, T);
    const double tau = c.p.tau_factor * (1.0 / c.p.density_factor) * kappa * sigma;
    if (c.p.opacity == FARGO_OPACITY_SIMPLE)
	return 3.0 / 8.0 * tau;
    if (c.p.heating_star)
	return 3.0 / 8.0 * tau + 0.5 + 1.0 / (4.0 * tau + c.p.tau_min);
    return 3.0 / 8.0 * tau + sqrt(3.0) / 4.0 + 1.0 / (4.0 * tau + c.p.tau_min);
}
__device__ __forceinline__ bool rad_enabled(const DevView &c) { return c.p.cooling_surface || c.p.heating_star; }
// what thermal_cooling adds to Q- of a cell of ring 1 .. nr - 2
__device__ __forceinline__ double rad_qminus(const DevView &c, const double T, const double tau_eff)
{
    const double T4 = fm_pow4(T), Tmin4 = fm_pow4(c.p.minimum_temperature);
    return c.p.surface_cooling_factor * 2 * c.p.sigma_sb * (T4 - Tmin4) / tau_eff;
}
// irradiation (SourceEuler.cpp:599-610): Q+ of a cell of ring 1 .. nr - 2 collects every irradiating body in turn
__device__ __forceinline__ void rad_add_qplus(const DevView &c, const int i, const double cosj, const double sinj, const double H,
					       const double tau_eff, double &qplus)
{
    const double xc = c.g.rmed[i] * cosj, yc = c.g.rmed[i] * sinj;
    const double HoverR = H / c.g.rmed[i]; // ASPECTRATIO as compute_scale_height stores it (SourceEuler.cpp:1148-1151)
    for (int k = 0; k < c.b.n; ++k) {
	const double T_star = c.b.temperature[k];
	if (!(T_star > 0))
	    continue;
	const double x = c.b.x[k], y = c.b.y[k], R_star = c.b.radius[k];
	const double min_dist = (x * x + y * y > 1e-10) ? stdmax(R_star, c.b.cubic_smoothing_radius[k]) : R_star;
	const double dx = x - xc, dy = y - yc;
	const double distance = stdmax(sqrt(dx * dx + dy * dy), min_dist);
	const double roverd = distance < R_star ? 1.0 : R_star / distance;
	const double W_G = 0.4 * roverd + HoverR * (9.0 / 7.0 - 1.0);
	const double T_irrad_pow4 = (1.0 - 0.5) * fm_pow4(T_star) * (roverd * roverd) * W_G;
	const double q = 2.0 * c.p.sigma_sb * T_irrad_pow4 / tau_eff;
	qplus += c.b.irradiation_ramp[k] * q;
    }
}

// scurve_cooling (SourceEuler.cpp:726-831; SurfaceCooling: scurve): the cooling S-curve of a dwarf-nova disk after Ichikawa &
// Osaki (1992) / Kimura et al. (2020), a fit written in cgs with log10 / pow — CUDA's against glibc's differ in the last bits,
// so Q- agrees with the reference's to ~1e-15 relative, not to the bit.  Returns what the function adds to Q- of a cell of
// ring 1 .. nr - 2 and the TAU_EFF it stores (read by SubStep3's density-floor branch).  T: the cell's temperature, code units.
__device__ __forceinline__ double rad_scurve_qminus(const DevView &c, const int i, const double sigma, const double T, const double mu,
						     double &tau_eff)
{
    const double SigmaCGS_threshold = 2.0, temperatureCGS_threshold = 1200.0;
    const bool kimura = c.p.cooling_scurve == 2;
    const double F_hot_const = kimura ? 23.405 : 25.49, muExponent = kimura ? 0.31 : -0.31;
    const double SigmaCGS = sigma * c.p.surface_density_cgs;
    const double SigmaCGS_tmp = stdmax(SigmaCGS, SigmaCGS_threshold);
    const double temperatureCGS = T * c.p.temperature_cgs;
    const double temperatureCGS_tmp = stdmax(temperatureCGS, temperatureCGS_threshold);
    const double rCGS = c.g.rmed[i] * c.p.length_cgs;
    const double M = c.p.hydro_center_mass * c.p.mass_cgs;
    const double omega_keplerCGS = sqrt(c.p.G_cgs * M / (rCGS * rCGS * rCGS));
    const double sigma_sb_cgs = c.p.sigma_sb_cgs;
    const double logTA = -1.0 / 5.49 * (0.62 * log10(omega_keplerCGS) + 1.62 * log10(SigmaCGS_tmp) + muExponent * log10(mu) - 25.48 -
				      log10(sigma_sb_cgs));
    const double TA = pow(10.0, logTA);
    const double FA = sigma_sb_cgs * fm_pow4(TA);
    const double logFA = log10(FA);
    const double KCGS = 11.0 + 0.4 * log10(2.0e10 / rCGS);
    const double logFB = stdmax(KCGS, logFA);
    const double logTB_aux = log10(omega_keplerCGS) + 2.0 * log10(SigmaCGS_tmp) + 0.5 * log10(mu) + F_hot_const;
    const double logTB = (logFB + logTB_aux) / 8.0;
    const double TB = pow(10.0, logTB);
    double logFtot;
    if (temperatureCGS_tmp < TA)
	logFtot = 9.49 * log10(temperatureCGS_tmp) + 0.62 * log10(omega_keplerCGS) + 1.62 * log10(SigmaCGS_tmp) + muExponent * log10(mu) - 25.48;
    else if (temperatureCGS_tmp > TB)
	logFtot = 8.0 * log10(temperatureCGS_tmp) - log10(omega_keplerCGS) - 2.0 * log10(SigmaCGS_tmp) - 0.5 * log10(mu) - F_hot_const;
    else
	logFtot = (logFA - logFB) * log10(temperatureCGS_tmp / TB) / log10(TA / TB) + logFB;
    const double T4 = fm_pow4(T);
    const double factor = c.p.surface_cooling_factor;
    double F_tot = pow(10.0, logFtot) * (1.0 / c.p.energy_flux_cgs);
    F_tot *= sqrt(SigmaCGS / SigmaCGS_tmp); // pow(x, 0.5): correctly rounded in glibc, like sqrt
    const double tr = temperatureCGS / temperatureCGS_tmp;
    F_tot *= tr * tr;
    const double F_Blackbody = c.p.sigma_sb * T4;
    const double qminus_scurve = 2.0 * factor * stdmin(F_tot, F_Blackbody);
    tau_eff = factor * 2 * c.p.sigma_sb * T4 / qminus_scurve;
    return qminus_scurve;
}
