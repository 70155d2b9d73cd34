/* fargo_oracle.c — CPU restatement of FargoCPT's per-timestep hydro step.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path in
 * fargocpt_b200/csrc; it is imported only by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg.  The product path never calls it (there is no CPU fallback).
 *
 * It is a plain-C restatement (not a copy) of the reference algorithm, one function per
 * reference function, each citing the reference file:line it follows (paths relative to the
 * reference's src/).  Arithmetic is IEEE double with the reference's operation order; build with
 * -O2 -ffp-contract=off so no FMA contraction happens (the pinned reference build uses the same
 * flags, see oracle/Makefile.ref).
 *
 * PARITY PINNED: tests/test_oracle_vs_golden.py checks this file bit-for-bit against snapshots
 * produced by the unmodified reference built into oracle/_ref (fixtures in tests/golden/, made by
 * tests/golden/make_golden.py).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#include "../include/fargo_b200.h"
#include "../host/fargo_pvte.h" /* the PVTE table builder is product code; the oracle only calls it */

#define SEARCH_BUFFER 15 /* init.cpp:43 */

typedef struct fargo_oracle {
    fargo_params p;
    int rank, nranks;
    int nr, ns;	    /* local NRadial, NAzimuthal */
    int imin, imax; /* split.cpp:50-61 */
    /* split.cpp:66-78 */
    int zero_no_ghost, one_no_ghost_vr, max_no_ghost, maxmo_no_ghost_vr;
    int zero_or_active, max_or_active, first_active, active_size;
    double dphi, invdphi;
    /* global 1-D arrays (size nrad_global + SEARCH_BUFFER + 2) and local views (offset imin) */
    double *g_radii, *g_rmed;
    double *rinf, *rsup, *rmed, *surf, *invrmed, *invsurf, *invdiffrsup, *invdiffrsuprb, *twodiffrasq,
	*fourthirdinvrbinvdphisq, *invrinf, *invdiffrmed;
    double *cosphi, *sinphi;
    /* state */
    double *sigma, *vrad, *vazi, *energy;
    double *sigma0, *vrad0, *vazi0, *energy0;
    /* derived */
    double *temperature, *pressure, *soundspeed, *scale_height, *viscosity, *potential;
    /* EquationOfState: PVTE: the GAMMAEFF, MU, GAMMA1 grids (data.h) and the lookup tables */
    double *gamma_eff, *mu_cell, *gamma1;
    fargo_pvte_tables *pv;
    int kicks_this_step; /* fargo_oracle_kick calls since the last fargo_oracle_finish_step */
    double *qplus, *qminus, *divv, *trr, *tpp, *trp, *qr, *qphi, *nusig, *nusig_rp, *cf_r, *cf_phi, *tau_eff;
    double *massflow; /* MASSFLOW grid [nr + 1][ns], NULL unless fargo_oracle_track_massflow */
    int track_bflow;
    double bflow[4]; /* MassDelta: inner inflow / outflow, outer inflow / outflow */
    double dmass[4]; /* MassDelta: inner wave-damping mass creation / removal, outer creation / removal */
    /* transport scratch (TransportEuler.cpp:32-46) */
    double *rmp, *rmm, *amp, *amm, *vres, *work, *qrstar, *densstar, *densint, *tempshift, *dq, *vmean;
    int *nshift;
    int visc_calculated; /* viscosity.cpp:100 */
    fargo_bodies bodies;
    double time;
    /* multi-rank: halo staging for tests (the exchange itself is done by the caller) */
} fargo_oracle;


/* std::min / std::max semantics (returns the first argument on ties and NaNs), used everywhere the
 * reference calls them so that signed zeros and NaNs propagate identically */
static inline double stdmin(double a, double b) { return (b < a) ? b : a; }
static inline double stdmax(double a, double b) { return (a < b) ? b : a; }

#define IDX(o, i, j) ((size_t)(i) * (size_t)(o)->ns + (size_t)(j))
/* pvte::get_gamma_eff / get_mu / get_gamma1 (pvte_law.cpp:543-568) */
#define GEFF(o, c) ((o)->p.pvte ? (o)->gamma_eff[c] : (o)->p.gamma)
#define MUC(o, c) ((o)->p.pvte ? (o)->mu_cell[c] : (o)->p.mu)
#define GAM1(o, c) ((o)->p.pvte ? (o)->gamma1[c] : (o)->p.gamma)

static double *dalloc(size_t n)
{
    double *p = (double *)calloc(n ? n : 1, sizeof(double));
    if (!p) {
	fprintf(stderr, "fargo_oracle: out of memory\n");
	abort();
    }
    return p;
}

/* ------------------------------------------------------------------------------------------
 * geometry: init_radialarrays, init.cpp:169-225 */
static void init_geometry(fargo_oracle *o, const double *radii)
{
    const int gn = o->p.nrad;
    const int n1 = gn + SEARCH_BUFFER + 2;
    o->g_radii = dalloc(n1 + 1);
    for (int i = 0; i <= gn; ++i)
	o->g_radii[i] = radii[i];
    /* rings beyond the grid (search buffer) are never read by the hot path; extend geometrically */
    for (int i = gn + 1; i <= n1; ++i)
	o->g_radii[i] = o->g_radii[i - 1] * (o->g_radii[gn] / o->g_radii[gn - 1]);
    o->dphi = 2.0 * M_PI / (double)o->ns;    /* Interpret.cpp:230 */
    o->invdphi = (double)o->ns / (2.0 * M_PI); /* Interpret.cpp:231 */

    const int nl = o->nr + 2;
    o->rinf = dalloc(nl); o->rsup = dalloc(nl); o->rmed = dalloc(nl); o->surf = dalloc(nl);
    o->invrmed = dalloc(nl); o->invsurf = dalloc(nl); o->invdiffrsup = dalloc(nl);
    o->invdiffrsuprb = dalloc(nl); o->twodiffrasq = dalloc(nl); o->fourthirdinvrbinvdphisq = dalloc(nl);
    o->invrinf = dalloc(nl); o->invdiffrmed = dalloc(nl);
    for (int n = 0; n < nl; ++n) { /* init.cpp:188-216 */
	const double ri = o->g_radii[n + o->imin], rs = o->g_radii[n + o->imin + 1];
	o->rinf[n] = ri;
	o->rsup[n] = rs;
	double rm = 2.0 / 3.0 * (pow(rs, 3) - pow(ri, 3));
	rm = rm / (pow(rs, 2) - pow(ri, 2));
	o->rmed[n] = rm;
	o->surf[n] = M_PI * (pow(rs, 2) - pow(ri, 2)) / (double)o->ns;
	o->invrmed[n] = 1.0 / rm;
	o->invsurf[n] = 1.0 / o->surf[n];
	o->invdiffrsup[n] = 1.0 / (rs - ri);
	o->invdiffrsuprb[n] = 1.0 / ((rs - ri) * rm);
	o->twodiffrasq[n] = 2.0 / (rs * rs - ri * ri);
	o->fourthirdinvrbinvdphisq[n] = 4.0 / 3.0 / rm * o->invdphi * o->invdphi;
	o->invrinf[n] = 1.0 / ri;
    }
    for (int n = 1; n < nl; ++n) /* init.cpp:221-225 */
	o->invdiffrmed[n] = 1.0 / (o->rmed[n] - o->rmed[n - 1]);
    /* cell centres, SideEuler.cpp:56-65: x = Rmed*cos(dphi*j) */
    o->cosphi = dalloc(o->ns);
    o->sinphi = dalloc(o->ns);
    for (int j = 0; j < o->ns; ++j) {
	o->cosphi[j] = cos(o->dphi * (double)j);
	o->sinphi[j] = sin(o->dphi * (double)j);
    }
}

/* split.cpp:38-87 */
static int split_domain(fargo_oracle *o)
{
    const int N = o->p.nrad, np = o->nranks, r = o->rank;
    const int size_low = N / np, size_high = size_low + 1, rem = N % np;
    if (np > 1 && size_low < 2 * FARGO_CPUOVERLAP)
	return 1;
    if (r < rem) {
	o->imin = size_high * r;
	o->imax = o->imin + size_high - 1;
    } else {
	o->imin = size_high * rem + (r - rem) * size_low;
	o->imax = o->imin + size_low - 1;
    }
    if (r > 0)
	o->imin -= FARGO_CPUOVERLAP;
    if (r < np - 1)
	o->imax += FARGO_CPUOVERLAP;
    o->nr = o->imax - o->imin + 1;
    const int first = (r == 0), last = (r == np - 1);
    o->zero_no_ghost = first ? 1 : 0;
    o->one_no_ghost_vr = first ? 2 : 1;
    o->max_no_ghost = o->nr - (last ? 1 : 0);
    o->maxmo_no_ghost_vr = o->nr + 1 - (last ? 2 : 1);
    o->zero_or_active = first ? 0 : FARGO_CPUOVERLAP;
    o->first_active = first ? FARGO_GHOSTCELLS_B : FARGO_CPUOVERLAP;
    o->max_or_active = o->nr - (last ? 0 : FARGO_CPUOVERLAP);
    o->active_size = o->nr - (last ? FARGO_GHOSTCELLS_B : FARGO_CPUOVERLAP);
    return 0;
}

fargo_oracle *fargo_oracle_create(const fargo_params *params, const double *radii, int rank, int nranks)
{
    fargo_oracle *o = (fargo_oracle *)calloc(1, sizeof(*o));
    o->p = *params;
    o->rank = rank;
    o->nranks = nranks;
    o->ns = params->naz;
    if (split_domain(o)) {
	free(o);
	return NULL;
    }
    init_geometry(o, radii);
    const size_t ns = (size_t)(o->nr) * o->ns, nv = (size_t)(o->nr + 1) * o->ns;
    o->sigma = dalloc(ns); o->vrad = dalloc(nv); o->vazi = dalloc(ns); o->energy = dalloc(ns);
    o->sigma0 = dalloc(ns); o->vrad0 = dalloc(nv); o->vazi0 = dalloc(ns); o->energy0 = dalloc(ns);
    o->temperature = dalloc(ns); o->pressure = dalloc(ns); o->soundspeed = dalloc(ns);
    o->scale_height = dalloc(ns); o->viscosity = dalloc(ns); o->potential = dalloc(ns);
    o->qplus = dalloc(ns); o->qminus = dalloc(ns); o->divv = dalloc(ns); o->trr = dalloc(ns);
    o->tpp = dalloc(ns); o->trp = dalloc(nv); o->qr = dalloc(ns); o->qphi = dalloc(ns);
    o->nusig = dalloc(ns); o->nusig_rp = dalloc(nv + o->ns); o->cf_r = dalloc(ns); o->cf_phi = dalloc(ns);
    o->tau_eff = dalloc(ns);
    o->rmp = dalloc(ns); o->rmm = dalloc(ns); o->amp = dalloc(ns); o->amm = dalloc(ns);
    o->vres = dalloc(ns); o->work = dalloc(ns); o->qrstar = dalloc(nv); o->densstar = dalloc(nv);
    o->densint = dalloc(ns); o->tempshift = dalloc(ns); o->dq = dalloc(ns); o->vmean = dalloc(o->nr + 1);
    o->nshift = (int *)calloc(o->nr + 1, sizeof(int));
    o->bodies.n = 1;
    o->bodies.mass[0] = params->hydro_center_mass;
    return o;
}

void fargo_oracle_destroy(fargo_oracle *o)
{
    if (!o)
	return;
    double **all[] = {&o->g_radii, &o->rinf, &o->rsup, &o->rmed, &o->surf, &o->invrmed, &o->invsurf, &o->invdiffrsup,
		      &o->invdiffrsuprb, &o->twodiffrasq, &o->fourthirdinvrbinvdphisq, &o->invrinf, &o->invdiffrmed,
		      &o->cosphi, &o->sinphi, &o->sigma, &o->vrad, &o->vazi, &o->energy, &o->sigma0, &o->vrad0,
		      &o->vazi0, &o->energy0, &o->temperature, &o->pressure, &o->soundspeed, &o->scale_height,
		      &o->viscosity, &o->potential, &o->qplus, &o->qminus, &o->divv, &o->trr, &o->tpp, &o->trp,
		      &o->qr, &o->qphi, &o->nusig, &o->nusig_rp, &o->cf_r, &o->cf_phi, &o->tau_eff, &o->rmp, &o->rmm,
		      &o->amp, &o->amm, &o->vres, &o->work, &o->qrstar, &o->densstar, &o->densint, &o->tempshift,
		      &o->dq, &o->vmean, &o->gamma_eff, &o->mu_cell, &o->gamma1};
    for (size_t k = 0; k < sizeof(all) / sizeof(all[0]); ++k)
	free(*all[k]);
    fargo_pvte_free(o->pv);
    free(o->nshift);
    free(o);
}

int fargo_oracle_local_nrad(const fargo_oracle *o) { return o->nr; }
int fargo_oracle_local_imin(const fargo_oracle *o) { return o->imin; }

/* stress::calculate_Reynolds_stress (stress.cpp:34-70): ring means of the cell-centred velocities (serial sums in index
 * order), then Sigma * (v_r - <v_r>) * (v_phi - <v_phi>).  Written into the transport scratch grid `work`. */
static double *reynolds_stress(fargo_oracle *o)
{
    const int ns = o->ns;
    for (int nr = 0; nr < o->nr; ++nr) {
	double v_radial_mean = 0.0;
	double v_azimuthal_mean = 0.0;
	for (int naz = 0; naz < ns; ++naz) {
	    const int naz_next = (naz == ns - 1 ? 0 : naz + 1);
	    v_radial_mean += 0.5 * (o->vrad[IDX(o, nr, naz)] + o->vrad[IDX(o, nr + 1, naz)]);
	    v_azimuthal_mean += 0.5 * (o->vazi[IDX(o, nr, naz)] + o->vazi[IDX(o, nr, naz_next)]);
	}
	v_azimuthal_mean /= (double)ns;
	v_radial_mean /= (double)ns;
	for (int naz = 0; naz < ns; ++naz) {
	    const int naz_next = (naz == ns - 1 ? 0 : naz + 1);
	    o->work[IDX(o, nr, naz)] = o->sigma[IDX(o, nr, naz)] *
				       (0.5 * (o->vrad[IDX(o, nr, naz)] + o->vrad[IDX(o, nr + 1, naz)]) - v_radial_mean) *
				       (0.5 * (o->vazi[IDX(o, nr, naz)] + o->vazi[IDX(o, nr, naz_next)]) - v_azimuthal_mean);
	}
    }
    return o->work;
}

static double *field_ptr(fargo_oracle *o, int f, int *rings)
{
    *rings = o->nr;
    switch (f) {
    case FARGO_T_REYNOLDS: return reynolds_stress(o);
    case FARGO_SIGMA: return o->sigma;
    case FARGO_VRAD: *rings = o->nr + 1; return o->vrad;
    case FARGO_VAZI: return o->vazi;
    case FARGO_ENERGY: return o->energy;
    case FARGO_SIGMA0: return o->sigma0;
    case FARGO_VRAD0: *rings = o->nr + 1; return o->vrad0;
    case FARGO_VAZI0: return o->vazi0;
    case FARGO_ENERGY0: return o->energy0;
    case FARGO_QPLUS: return o->qplus;
    case FARGO_QMINUS: return o->qminus;
    case FARGO_TEMPERATURE: return o->temperature;
    case FARGO_PRESSURE: return o->pressure;
    case FARGO_SOUNDSPEED: return o->soundspeed;
    case FARGO_SCALE_HEIGHT: return o->scale_height;
    case FARGO_VISCOSITY: return o->viscosity;
    case FARGO_POTENTIAL: return o->potential;
    case FARGO_GAMMAEFF: return o->gamma_eff;
    case FARGO_MU: return o->mu_cell;
    case FARGO_GAMMA1: return o->gamma1;
    case FARGO_MASSFLOW: *rings = o->nr + 1; return o->massflow;
    }
    return NULL;
}

/* read2D slab semantics, polargrid.cpp:343-349: local ring n <- global ring n+IMIN */
int fargo_oracle_upload_field(fargo_oracle *o, int f, const double *host_global)
{
    int rings;
    double *d = field_ptr(o, f, &rings);
    if (!d)
	return 1;
    memcpy(d, host_global + (size_t)o->imin * o->ns, (size_t)rings * o->ns * sizeof(double));
    return 0;
}

/* write2D slab semantics, polargrid.cpp:150-176 */
int fargo_oracle_download_field(fargo_oracle *o, int f, double *host_global)
{
    int rings;
    double *d = field_ptr(o, f, &rings);
    if (!d)
	return 1;
    int first = o->zero_or_active, count = o->max_or_active - o->zero_or_active;
    if (rings == o->nr + 1 && o->rank == o->nranks - 1)
	count += 1;
    memcpy(host_global + (size_t)(o->imin + first) * o->ns, d + (size_t)first * o->ns,
	   (size_t)count * o->ns * sizeof(double));
    return 0;
}

int fargo_oracle_download_slab(fargo_oracle *o, int f, double *host_slab)
{
    int rings;
    double *d = field_ptr(o, f, &rings);
    if (!d)
	return 1;
    memcpy(host_slab, d, (size_t)rings * o->ns * sizeof(double));
    return 0;
}

/* raw slab overwrite (tests: halo exchange between oracle slabs) */
int fargo_oracle_upload_slab(fargo_oracle *o, int f, const double *host_slab)
{
    int rings;
    double *d = field_ptr(o, f, &rings);
    if (!d)
	return 1;
    memcpy(d, host_slab, (size_t)rings * o->ns * sizeof(double));
    return 0;
}

/* damping.cpp:287-296 */
/* correct_v_azimuthal, SideEuler.cpp:79-95 */
int fargo_oracle_correct_vazi(fargo_oracle *o, double domega)
{
    for (int i = 0; i < o->nr; ++i)
	for (int j = 0; j < o->ns; ++j)
	    o->vazi[IDX(o, i, j)] -= domega * o->rmed[i];
    return 0;
}

int fargo_oracle_copy_initial_values(fargo_oracle *o)
{
    const size_t ns = (size_t)o->nr * o->ns, nv = (size_t)(o->nr + 1) * o->ns;
    memcpy(o->vrad0, o->vrad, nv * sizeof(double));
    memcpy(o->vazi0, o->vazi, ns * sizeof(double));
    memcpy(o->sigma0, o->sigma, ns * sizeof(double));
    memcpy(o->energy0, o->energy, ns * sizeof(double));
    return 0;
}

int fargo_oracle_set_bodies(fargo_oracle *o, const fargo_bodies *b)
{
    o->bodies = *b;
    return 0;
}
int fargo_oracle_set_time(fargo_oracle *o, double t)
{
    o->time = t;
    return 0;
}

/* Theo.cpp:246-249 */
static double omega_kepler(const fargo_oracle *o, double r) { return sqrt(o->p.G * o->p.hydro_center_mass / (r * r * r)); }

/* ------------------------------------------------------------------------------------------
 * EOS-derived fields */

/* compute_sound_speed_normal, SourceEuler.cpp:957-995 */
static void compute_sound_speed(fargo_oracle *o)
{
    const int Nr = o->nr, Nphi = o->ns;
#pragma omp parallel for
    for (int nr = 0; nr < Nr; ++nr) {
	for (int naz = 0; naz < Nphi; ++naz) {
	    const size_t c = IDX(o, nr, naz);
	    if (o->p.adiabatic) {
		const double gamma_eff = GEFF(o, c), gamma1 = GAM1(o, c);
		o->soundspeed[c] = sqrt(gamma1 * (gamma_eff - 1.0) * o->energy[c] / o->sigma[c]);
	    } else {
		const double vK = sqrt(o->p.G * o->p.hydro_center_mass / o->rmed[nr]);
		const double h = o->p.aspectratio_ref * pow(o->rmed[nr], o->p.flaring_index);
		o->soundspeed[c] = h * vK;
	    }
	}
    }
}

/* compute_scale_height_old, SourceEuler.cpp:1121-1154 (ASPECTRATIO grid is output-only: skipped) */
static void compute_scale_height(fargo_oracle *o)
{
    const int Nr = o->nr, Nphi = o->ns;
#pragma omp parallel for
    for (int nr = 0; nr < Nr; ++nr) {
	const double inv_omega_kepler = 1.0 / omega_kepler(o, o->rmed[nr]);
	for (int naz = 0; naz < Nphi; ++naz) {
	    const size_t c = IDX(o, nr, naz);
	    if (o->p.adiabatic)
		o->scale_height[c] = o->soundspeed[c] / (sqrt(GAM1(o, c))) * inv_omega_kepler;
	    else
		o->scale_height[c] = o->soundspeed[c] * inv_omega_kepler;
	}
    }
}

/* compute_pressure, SourceEuler.cpp:1345-1376 */
static void compute_pressure(fargo_oracle *o)
{
    const size_t n = (size_t)o->nr * o->ns;
#pragma omp parallel for
    for (size_t c = 0; c < n; ++c) {
	if (o->p.adiabatic)
	    o->pressure[c] = (GEFF(o, c) - 1.0) * o->energy[c];
	else
	    o->pressure[c] = o->sigma[c] * (o->soundspeed[c] * o->soundspeed[c]);
    }
}

/* compute_temperature, SourceEuler.cpp:1378-1408 */
static void compute_temperature(fargo_oracle *o)
{
    const size_t n = (size_t)o->nr * o->ns;
    const double Rgas = o->p.Rgas;
#pragma omp parallel for
    for (size_t c = 0; c < n; ++c) {
	if (o->p.adiabatic) {
	    const double c_v_inv = MUC(o, c) / Rgas * (GEFF(o, c) - 1.0);
	    o->temperature[c] = c_v_inv * o->energy[c] / o->sigma[c];
	} else {
	    o->temperature[c] = o->p.mu / Rgas * o->pressure[c] / o->sigma[c];
	}
    }
}

/* viscosity::update_viscosity, viscosity.cpp:98-137 (AlphaMode CONST_ALPHA) */
/* viscosity::get_alpha, viscosity/viscosity.cpp:31-49 (AlphaMode 0 and 1) */
static double get_alpha(const fargo_oracle *o, int nr, int naz)
{
    if (o->p.alpha_mode == 1) { /* SCURVE_ALPHA: reads the TEMPERATURE grid as last computed */
	const double temperatureCGS = o->temperature[IDX(o, nr, naz)] * o->p.temperature_cgs;
	const double alpha_cool = o->p.alpha_cold * pow(o->rmed[nr] / 0.4, 0.3);
	const double alpha_hot = o->p.alpha_hot;
	return pow(10.0, 0.5 * (log10(alpha_hot) - log10(alpha_cool)) * (1.0 - tanh((4.0 - log10(temperatureCGS)) / 0.4)) + log10(alpha_cool));
    }
    return o->p.viscous_alpha;
}

static void update_viscosity(fargo_oracle *o)
{
    const size_t n = (size_t)o->nr * o->ns;
    if (o->p.viscous_alpha > 0) {
#pragma omp parallel for
	for (int nr = 0; nr < o->nr; ++nr)
	    for (int naz = 0; naz < o->ns; ++naz) {
		const size_t c = IDX(o, nr, naz);
		o->viscosity[c] = get_alpha(o, nr, naz) * o->scale_height[c] * o->soundspeed[c];
	    }
    } else {
	if (!o->visc_calculated)
	    for (size_t c = 0; c < n; ++c)
		o->viscosity[c] = o->p.constant_viscosity;
	o->visc_calculated = 1;
    }
}

/* assure_temperature_range, SourceEuler.cpp:136-202 */
static void assure_temperature_range(fargo_oracle *o)
{
    const size_t n = (size_t)o->nr * o->ns;
    const double Tmin = o->p.minimum_temperature, Tmax = o->p.maximum_temperature;
    const double R = o->p.Rgas;
#pragma omp parallel for
    for (size_t c = 0; c < n; ++c) {
	const double mu = MUC(o, c), g = GEFF(o, c);
	const double minimum_energy = Tmin * o->sigma[c] / mu * R / (g - 1.0);
	const double maximum_energy = Tmax * o->sigma[c] / mu * R / (g - 1.0);
	if (!(o->energy[c] > minimum_energy))
	    o->energy[c] = Tmin * o->sigma[c] / mu * R / (g - 1.0);
	if (!(o->energy[c] < maximum_energy))
	    o->energy[c] = Tmax * o->sigma[c] / mu * R / (g - 1.0);
    }
}

/* pvte::compute_gamma_mu (pvte_law.cpp:497-541): gamma_eff, mu, Gamma_1 of every cell from the lookup tables, at the midplane
 * density the STORED scale height gives and the cell's specific energy */
static void compute_gamma_mu(fargo_oracle *o)
{
    const size_t n = (size_t)o->nr * o->ns;
#pragma omp parallel for
    for (size_t c = 0; c < n; ++c) {
	const double sigma = o->sigma[c], H = o->scale_height[c];
	const double densityCGS = sigma / (o->p.density_factor * H) * o->p.density_cgs;
	const double energyCGS = o->energy[c] * o->p.energy_density_cgs / (sigma * o->p.surface_density_cgs);
	fargo_pvte_lookup(o->pv, densityCGS, energyCGS, &o->gamma_eff[c], &o->mu_cell[c], &o->gamma1[c]);
    }
}

/* recalculate_viscosity, SourceEuler.cpp:205-223 */
static void recalculate_viscosity(fargo_oracle *o)
{
    if (o->p.adiabatic) {
	if (o->p.pvte)
	    compute_gamma_mu(o);
	compute_sound_speed(o);
	compute_scale_height(o);
    }
    update_viscosity(o);
}

/* recalculate_derived_disk_quantities, SourceEuler.cpp:225-249 */
int fargo_oracle_stage_derived(fargo_oracle *o)
{
    if (!o->p.adiabatic) {
	compute_pressure(o);
    } else {
	if (o->p.pvte)
	    compute_gamma_mu(o);
	compute_temperature(o);
	compute_sound_speed(o);
	compute_scale_height(o);
	compute_pressure(o);
    }
    update_viscosity(o);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * CalculateNbodyPotential, Pframeforce.cpp:21-86; smoothing: Force.cpp:124-159 */
int fargo_oracle_stage_potential(fargo_oracle *o)
{
    const int Nr = o->nr, Nphi = o->ns;
    const fargo_bodies *b = &o->bodies;
#pragma omp parallel for
    for (int nr = 0; nr < Nr; ++nr) {
	for (int naz = 0; naz < Nphi; ++naz) {
	    const size_t c = IDX(o, nr, naz);
	    const double x = o->rmed[nr] * o->cosphi[naz];
	    const double y = o->rmed[nr] * o->sinphi[naz];
	    double pot = 0.0;
	    for (int k = 0; k < b->n; ++k) {
		const double smooth = o->p.thickness_smoothing * o->scale_height[c];
		const double dx = x - b->x[k];
		const double dy = y - b->y[k];
		const double dist_2 = dx * dx + dy * dy;
		const double d_smoothed = sqrt(dist_2 + smooth * smooth);
		double smooth_factor_klahr = 1.0;
		if (b->cubic_smoothing_radius[k] > 0.0) {
		    const double r_sm = b->cubic_smoothing_radius[k];
		    if (d_smoothed < r_sm)
			smooth_factor_klahr =
			    (pow(d_smoothed / r_sm, 4.0) - 2.0 * pow(d_smoothed / r_sm, 3.0) + 2.0 * d_smoothed / r_sm);
		}
		pot += -o->p.G * b->mass[k] / d_smoothed * smooth_factor_klahr;
	    }
	    pot += -b->indirect_x * x - b->indirect_y * y;
	    o->potential[c] = pot;
	}
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * update_with_sourceterms, SourceEuler.cpp:325-493 */
int fargo_oracle_stage_sources(fargo_oracle *o, double dt)
{
    const int Nphi = o->ns;
    const double OmegaF = o->bodies.omega_frame;
    /* momentum_update_radial :337-370 */
#pragma omp parallel for
    for (int nr = o->one_no_ghost_vr; nr < o->maxmo_no_ghost_vr; ++nr) {
	for (int naz = 0; naz < Nphi; ++naz) {
	    const size_t c = IDX(o, nr, naz), cm = IDX(o, nr - 1, naz);
	    double gradp = 2.0 / (o->sigma[c] + o->sigma[cm]);
	    gradp *= (o->pressure[c] - o->pressure[cm]);
	    gradp *= o->invdiffrmed[nr];
	    const double gradphi = (o->potential[c] - o->potential[cm]) * o->invdiffrmed[nr];
	    const int naz_next = (naz == Nphi - 1 ? 0 : naz + 1);
	    const double vsum = o->vazi[c] + o->vazi[IDX(o, nr, naz_next)] + o->vazi[cm] + o->vazi[IDX(o, nr - 1, naz_next)];
	    const double vt = 0.25 * vsum + o->rinf[nr] * OmegaF;
	    const double vt2 = vt * vt;
	    const double centrifugal_accel = vt2 * o->invrinf[nr];
	    o->vrad[c] += dt * (-gradp - gradphi + centrifugal_accel);
	}
    }
    /* momentum_update_azimuthal :382-427 */
#pragma omp parallel for
    for (int nr = o->zero_no_ghost; nr < o->max_no_ghost; ++nr) {
	double supp_torque = 0.0;
	if (o->p.imposed_disk_drift != 0.0)
	    supp_torque = o->p.imposed_disk_drift * 0.5 * pow(o->rmed[nr], -2.5 + o->p.sigma_slope);
	const double invdxtheta = 2.0 / (o->dphi * (o->rsup[nr] + o->rinf[nr]));
	for (int naz = 0; naz < Nphi; ++naz) {
	    const int naz_prev = (naz == 0 ? Nphi - 1 : naz - 1);
	    const size_t c = IDX(o, nr, naz), cp = IDX(o, nr, naz_prev);
	    const double gradp = 2.0 / (o->sigma[c] + o->sigma[cp]) * (o->pressure[c] - o->pressure[cp]) * invdxtheta;
	    const double gradphi = (o->potential[c] - o->potential[cp]) * invdxtheta;
	    o->vazi[c] = o->vazi[c] + dt * (-gradp - gradphi);
	    if (o->p.imposed_disk_drift != 0.0)
		o->vazi[c] += dt * supp_torque;
	}
    }
    /* compression_heating :459-493 */
    if (o->p.adiabatic) {
	const int Nr = o->nr - 1;
#pragma omp parallel for
	for (int nr = 0; nr < Nr; ++nr) {
	    for (int naz = 0; naz < Nphi; ++naz) {
		const int naz_next = (naz == Nphi - 1 ? 0 : naz + 1);
		const size_t c = IDX(o, nr, naz);
		const double DIV_V = (o->vrad[IDX(o, nr + 1, naz)] * o->rinf[nr + 1] - o->vrad[c] * o->rinf[nr]) * o->invdiffrsuprb[nr] +
				     (o->vazi[IDX(o, nr, naz_next)] - o->vazi[c]) * o->invdphi * o->invrmed[nr];
		const double energy_old = o->energy[c];
		o->energy[c] = energy_old * exp(-(GEFF(o, c) - 1.0) * dt * DIV_V);
	    }
	}
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * art_visc::update_with_artificial_viscosity, artificial_viscosity.cpp:11-250 */
static void artvisc_TW(fargo_oracle *o, double dt)
{
    const int Nr = o->nr, Nphi = o->ns;
    const double C = o->p.artificial_viscosity_factor;
#pragma omp parallel for
    for (int nr = 0; nr < Nr; ++nr) { /* :49-88 */
	for (int naz = 0; naz < Nphi; ++naz) {
	    const int naz_next = (naz == Nphi - 1 ? 0 : naz + 1);
	    const size_t c = IDX(o, nr, naz), cu = IDX(o, nr + 1, naz);
	    const double eps_rr = (o->vrad[cu] - o->vrad[c]) * o->invdiffrsup[nr];
	    const double eps_pp =
		o->invrmed[nr] * ((o->vazi[IDX(o, nr, naz_next)] - o->vazi[c]) * o->invdphi + 0.5 * (o->vrad[cu] + o->vrad[c]));
	    const double div_V = stdmin(eps_rr + eps_pp, 0.0);
	    const double Dr = o->rinf[nr + 1] - o->rinf[nr];
	    const double rDphi = o->rmed[nr] * o->dphi;
	    double dx_sq;
	    if (Nphi <= 16) {
		const double m = stdmin(Dr, rDphi);
		dx_sq = m * m;
	    } else {
		const double m = stdmax(Dr, rDphi);
		dx_sq = m * m;
	    }
	    const double l_sq = (C * C) * dx_sq;
	    const double q_rr = l_sq * o->sigma[c] * -div_V * (eps_rr - 1.0 / 3.0 * div_V);
	    const double q_pp = l_sq * o->sigma[c] * -div_V * (eps_pp - 1.0 / 3.0 * div_V);
	    o->qr[c] = q_rr;
	    o->qphi[c] = q_pp;
	    if (o->p.adiabatic && o->p.artificial_viscosity_dissipation) {
		if (nr > o->zero_no_ghost && nr < o->max_no_ghost) {
		    const double Qplus = -l_sq * div_V * o->sigma[c] * 1.0 / 3.0 *
					 (eps_rr * eps_rr + eps_pp * eps_pp + (eps_rr - eps_pp) * (eps_rr - eps_pp));
		    o->energy[c] += Qplus * dt;
		}
	    }
	}
    }
#pragma omp parallel for
    for (int nr = 1; nr < Nr - 1; ++nr) { /* :90-117 */
	for (int naz = 0; naz < Nphi; ++naz) {
	    const int naz_prev = (naz == 0 ? Nphi - 1 : naz - 1);
	    const size_t c = IDX(o, nr, naz), cp = IDX(o, nr, naz_prev);
	    const double sigma_phi_avg = 0.5 * (o->sigma[c] + o->sigma[cp]);
	    const double dVp = 2.0 * dt / ((o->rsup[nr] + o->rinf[nr]) * sigma_phi_avg) * (o->qphi[c] - o->qphi[cp]) * o->invdphi;
	    o->vazi[c] += dVp;
	}
    }
#pragma omp parallel for
    for (int nr = o->one_no_ghost_vr; nr < o->maxmo_no_ghost_vr; ++nr) { /* :119-139 */
	for (int naz = 0; naz < Nphi; ++naz) {
	    const size_t c = IDX(o, nr, naz), cm = IDX(o, nr - 1, naz);
	    const double sigma_r_avg = 0.5 * (o->sigma[c] + o->sigma[cm]);
	    const double rm = o->rmed[nr], rmm = o->rmed[nr - 1];
	    const double dVr = o->p.radial_viscosity_factor * dt / sigma_r_avg * 2.0 / (rm * rm - rmm * rmm) *
			       ((o->qr[c] * rm - o->qr[cm] * rmm) - 0.5 * (o->qphi[c] + o->qphi[cm]) * (rm - rmm));
	    o->vrad[c] += dVr;
	}
    }
}

static void artvisc_SN(fargo_oracle *o, double dt)
{
    if (o->p.artificial_viscosity != FARGO_ARTVISC_SN)
	return; /* :150-151 */
    const int Nr = o->nr, Nphi = o->ns;
    const double C = o->p.artificial_viscosity_factor;
#pragma omp parallel for
    for (int nr = 0; nr < Nr; ++nr) { /* :165-189 */
	for (int naz = 0; naz < Nphi; ++naz) {
	    const int naz_next = (naz == Nphi - 1 ? 0 : naz + 1);
	    const size_t c = IDX(o, nr, naz);
	    const double dv_r = o->vrad[IDX(o, nr + 1, naz)] - o->vrad[c];
	    o->qr[c] = (dv_r < 0.0) ? (C * C) * o->sigma[c] * (dv_r * dv_r) : 0.0;
	    const double dv_phi = o->vazi[IDX(o, nr, naz_next)] - o->vazi[c];
	    o->qphi[c] = (dv_phi < 0.0) ? (C * C) * o->sigma[c] * (dv_phi * dv_phi) : 0.0;
	}
    }
    if (o->p.adiabatic && o->p.artificial_viscosity_dissipation) { /* :194-218 */
#pragma omp parallel for
	for (int nr = o->zero_no_ghost; nr < o->max_no_ghost; ++nr) {
	    const double dxtheta = o->dphi * o->rmed[nr];
	    const double invdxtheta = 1.0 / dxtheta;
	    for (int naz = 0; naz < Nphi; ++naz) {
		const int naz_next = (naz == Nphi - 1 ? 0 : naz + 1);
		const size_t c = IDX(o, nr, naz);
		const double dv_r = o->vrad[IDX(o, nr + 1, naz)] - o->vrad[c];
		const double dv_phi = o->vazi[IDX(o, nr, naz_next)] - o->vazi[c];
		o->energy[c] = o->energy[c] - dt * o->qr[c] * dv_r * o->invdiffrsup[nr] - dt * o->qphi[c] * dv_phi * invdxtheta;
	    }
	}
    }
#pragma omp parallel for
    for (int nr = o->one_no_ghost_vr; nr < o->maxmo_no_ghost_vr; ++nr) { /* :221-230 */
	for (int naz = 0; naz < Nphi; ++naz) {
	    const size_t c = IDX(o, nr, naz), cm = IDX(o, nr - 1, naz);
	    o->vrad[c] = o->vrad[c] - dt * 2.0 / (o->sigma[c] + o->sigma[cm]) * (o->qr[c] - o->qr[cm]) * o->invdiffrmed[nr];
	}
    }
#pragma omp parallel for
    for (int nr = o->zero_no_ghost; nr < o->max_no_ghost; ++nr) { /* :233-248 */
	const double dxtheta = o->dphi * o->rmed[nr];
	const double invdxtheta = 1.0 / dxtheta;
	for (int naz = 0; naz < Nphi; ++naz) {
	    const int naz_prev = (naz == 0 ? Nphi - 1 : naz - 1);
	    const size_t c = IDX(o, nr, naz), cp = IDX(o, nr, naz_prev);
	    o->vazi[c] = o->vazi[c] - dt * 2.0 / (o->sigma[c] + o->sigma[cp]) * (o->qphi[c] - o->qphi[cp]) * invdxtheta;
	}
    }
}

int fargo_oracle_stage_artvisc(fargo_oracle *o, double dt)
{
    if (o->p.artificial_viscosity == FARGO_ARTVISC_TW)
	artvisc_TW(o, dt);
    else
	artvisc_SN(o, dt);
    if (o->p.adiabatic && o->p.artificial_viscosity_dissipation)
	assure_temperature_range(o);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * viscosity::compute_viscous_stress_tensor, viscosity.cpp:139-350 */
static void compute_viscous_stress_tensor(fargo_oracle *o)
{
    const int Nr = o->nr, Nphi = o->ns;
#pragma omp parallel for
    for (int nr = 0; nr < Nr; ++nr) {
	for (int naz = 0; naz < Nphi; ++naz) {
	    const int naz_next = (naz == Nphi - 1 ? 0 : naz + 1);
	    const size_t c = IDX(o, nr, naz), cu = IDX(o, nr + 1, naz);
	    o->divv[c] = (o->vrad[cu] * o->rinf[nr + 1] - o->vrad[c] * o->rinf[nr]) * o->invdiffrsuprb[nr] +
			 (o->vazi[IDX(o, nr, naz_next)] - o->vazi[c]) * o->invdphi * o->invrmed[nr];
	}
    }
#pragma omp parallel for
    for (int nr = 0; nr < Nr; ++nr) {
	for (int naz = 0; naz < Nphi; ++naz) {
	    const int naz_next = (naz == Nphi - 1 ? 0 : naz + 1);
	    const size_t c = IDX(o, nr, naz), cu = IDX(o, nr + 1, naz);
	    const double drr = (o->vrad[cu] - o->vrad[c]) * o->invdiffrsup[nr];
	    o->trr[c] = 2.0 * o->viscosity[c] * o->sigma[c] * (drr - 1.0 / 3.0 * o->divv[c]);
	    const double dpp = (o->vazi[IDX(o, nr, naz_next)] - o->vazi[c]) * o->invdphi * o->invrmed[nr] +
			       0.5 * (o->vrad[cu] + o->vrad[c]) * o->invrmed[nr];
	    const double nu = o->viscosity[c], sigma = o->sigma[c];
	    o->tpp[c] = 2.0 * nu * sigma * (dpp - 1.0 / 3.0 * o->divv[c]);
	    o->nusig[c] = nu * sigma;
	}
    }
#pragma omp parallel for
    for (int nr = 1; nr < Nr; ++nr) {
	for (int naz = 0; naz < Nphi; ++naz) {
	    const int naz_prev = (naz == 0 ? Nphi - 1 : naz - 1);
	    const size_t c = IDX(o, nr, naz), cm = IDX(o, nr - 1, naz), cp = IDX(o, nr, naz_prev), cmp = IDX(o, nr - 1, naz_prev);
	    const double dvazirdr = (o->vazi[c] * o->invrmed[nr] - o->vazi[cm] * o->invrmed[nr - 1]) * o->invdiffrmed[nr];
	    const double dvrdphi = (o->vrad[c] - o->vrad[cp]) * o->invdphi;
	    const double drp = o->rinf[nr] * dvazirdr + dvrdphi * o->invrinf[nr];
	    const double nu = 0.25 * (o->viscosity[c] + o->viscosity[cm] + o->viscosity[cp] + o->viscosity[cmp]);
	    const double sigma = 0.25 * (o->sigma[c] + o->sigma[cm] + o->sigma[cp] + o->sigma[cmp]);
	    o->trp[c] = nu * sigma * drp;
	    o->nusig_rp[c] = nu * sigma;
	}
    }
    if (o->p.stabilize_viscosity) { /* :256-348 */
#pragma omp parallel for
	for (int nr = 1; nr < Nr; ++nr) {
	    for (int naz = 0; naz < Nphi; ++naz) {
		const int naz_prev = (naz == 0 ? Nphi - 1 : naz - 1);
		const int naz_next = (naz == Nphi - 1 ? 0 : naz + 1);
		const size_t c = IDX(o, nr, naz);
		const double NuSig_rp = o->nusig_rp[c];
		const double NuSig_rp_ip = o->nusig_rp[IDX(o, nr + 1, naz)];
		const double NuSig_rp_jp = o->nusig_rp[IDX(o, nr, naz_next)];
		const double NuSigma = o->nusig[c];
		const double NuSigma_jm = o->nusig[IDX(o, nr, naz_prev)];
		const double NuSigma_im = o->nusig[IDX(o, nr - 1, naz)];
		const double Ra3a = NuSig_rp * pow(o->rinf[nr], 3) * o->invdiffrmed[nr];
		const double Ra3b = NuSig_rp_ip * pow(o->rinf[nr + 1], 3) * o->invdiffrmed[nr + 1];
		const double cphi_rp = -o->invrmed[nr] * o->twodiffrasq[nr] * (Ra3b + Ra3a);
		const double cphi_pp = -o->fourthirdinvrbinvdphisq[nr] * (NuSigma + NuSigma_jm);
		const double sigma_avg_phi = 0.5 * (o->sigma[c] + o->sigma[IDX(o, nr, naz_prev)]);
		o->cf_phi[c] = (cphi_rp + cphi_pp) / (sigma_avg_phi * o->rmed[nr]);
		const double sigma_avg_r = 0.5 * (o->sigma[c] + o->sigma[IDX(o, nr - 1, naz)]);
		const double cr_rp = -(NuSig_rp_jp + NuSig_rp) / (o->dphi * o->dphi * o->rinf[nr]);
		const double cr_pp_1 = 2.0 * NuSigma * (0.5 * o->invrmed[nr] + 1.0 / 3.0 * o->rinf[nr] * o->invdiffrsuprb[nr]);
		const double cr_pp_2 = 2.0 * NuSigma_im * (0.5 * o->invrmed[nr - 1] - 1.0 / 3.0 * o->rinf[nr] * o->invdiffrsuprb[nr - 1]);
		const double cr_rr_1 = o->rmed[nr] * 2.0 * NuSigma * (-o->invdiffrsup[nr] + 1.0 / 3.0 * o->rinf[nr] * o->invdiffrsuprb[nr]);
		const double cr_rr_2 =
		    -1.0 * o->rmed[nr - 1] * 2.0 * NuSigma_im * (o->invdiffrsup[nr - 1] - 1.0 / 3.0 * o->rinf[nr] * o->invdiffrsuprb[nr - 1]);
		const double cr_pp = -0.5 * (cr_pp_1 + cr_pp_2);
		const double cr_rr = o->invdiffrmed[nr] * (cr_rr_1 + cr_rr_2);
		const double Rmed_mid = 0.5 * (o->rmed[nr] + o->rmed[nr - 1]);
		o->cf_r[c] = o->p.radial_viscosity_factor * (cr_rr + cr_rp + cr_pp) / (sigma_avg_r * Rmed_mid);
	    }
	}
    }
}

/* viscosity::update_velocities_with_viscosity, viscosity.cpp:355-426 */
static void update_velocities_with_viscosity(fargo_oracle *o, double dt)
{
    const int Nr = o->nr, Nphi = o->ns;
#pragma omp parallel for
    for (int nr = 1; nr < Nr - 1; ++nr) {
	for (int naz = 0; naz < Nphi; ++naz) {
	    const int naz_prev = (naz == 0 ? Nphi - 1 : naz - 1);
	    const size_t c = IDX(o, nr, naz), cp = IDX(o, nr, naz_prev);
	    const double sigma_avg = 0.5 * (o->sigma[c] + o->sigma[cp]);
	    const double ra2 = o->rinf[nr] * o->rinf[nr], rap2 = o->rinf[nr + 1] * o->rinf[nr + 1];
	    double dVp = dt * o->invrmed[nr] / (sigma_avg) *
			 ((2.0 / (rap2 - ra2)) * (rap2 * o->trp[IDX(o, nr + 1, naz)] - ra2 * o->trp[c]) + (o->tpp[c] - o->tpp[cp]) * o->invdphi);
	    if (o->p.stabilize_viscosity == 1) {
		const double cphi = o->cf_phi[c];
		const double corr = 1.0 / (stdmax(1.0 + dt * cphi, 0.0) - dt * cphi);
		dVp *= corr;
	    }
	    o->vazi[c] += dVp;
	}
    }
#pragma omp parallel for
    for (int nr = o->one_no_ghost_vr; nr < o->maxmo_no_ghost_vr; ++nr) {
	for (int naz = 0; naz < Nphi; ++naz) {
	    const int naz_next = (naz == Nphi - 1 ? 0 : naz + 1);
	    const size_t c = IDX(o, nr, naz), cm = IDX(o, nr - 1, naz);
	    const double sigma_avg = 0.5 * (o->sigma[c] + o->sigma[cm]);
	    double dVr = dt / (sigma_avg)*o->p.radial_viscosity_factor * 2.0 / (o->rmed[nr] + o->rmed[nr - 1]) *
			 ((o->rmed[nr] * o->trr[c] - o->rmed[nr - 1] * o->trr[cm]) * o->invdiffrmed[nr] +
			  (o->trp[IDX(o, nr, naz_next)] - o->trp[c]) * o->invdphi - 0.5 * (o->tpp[c] + o->tpp[cm]));
	    if (o->p.stabilize_viscosity == 1) {
		const double cr = o->cf_r[c];
		const double corr = 1.0 / (stdmax(1.0 + dt * cr, 0.0) - dt * cr);
		dVr *= corr;
	    }
	    o->vrad[c] += dVr;
	}
    }
}

int fargo_oracle_stage_viscosity(fargo_oracle *o, double dt)
{
    recalculate_viscosity(o);
    compute_viscous_stress_tensor(o);
    update_velocities_with_viscosity(o, dt);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * opacities (opacity.cpp:11-298).  The Lin & Papaloizou (1985) and Bell & Lin (1994) tables are eight power laws of
 * (rho, T) in cgs units joined by smoothing functions; the two differ in their constants and in five expressions, so they
 * are ONE function over a coefficient set here.  Operation order as in the reference (the results feed pow()). */
typedef struct {
    double power1, power2, power3, t234, t456, t678;
    double ak1, ak2, ak3, bk3, bk4, bk5, bk6, bk7, bk8;
    int bell;
} op_law;
static const op_law OP_LIN = {4.44444444e-2, 2.381e-2, 2.267e-1, 1.6e3, 5.7e3, 2.28e6, 2.e-4, 2.e16, 5.e-3, 50., 2.e-2, 2.e4, 1.e4, 1.5e10, 0.348, 0};
static const op_law OP_BELL = {2.8369e-2, 1.1464e-2, 2.2667e-1, 1.46e3, 4.51e3, 2.37e6, 2.e-4, 2.e16, 0.1e0, 10., 2.e-15, 1e4, 1e4, 1.5e10, 0.348, 1};
static double opacity_table(const op_law *L, double density, double temperature)
{
    if (L->bell && temperature < 1.0)
	temperature = 10.0; /* opacity.cpp:204-206 */
    if (temperature > L->t234 * pow(density, L->power1)) {
	const double ts4 = 1.e-4 * temperature;
	const double density13 = pow(density, 1.0 / 3.0);
	const double density23 = density13 * density13;
	const double ts42 = ts4 * ts4;
	const double ts44 = ts42 * ts42;
	const double ts48 = ts44 * ts44;
	if (temperature > L->t456 * pow(density, L->power2)) {
	    const int mid = L->bell ? ((temperature < L->t678 * pow(density, L->power3)) || ((density <= 1e10) && (temperature < 1e4)))
				    : ((temperature < L->t678 * pow(density, L->power3)) || (density <= 1e-10));
	    if (mid) { /* laws 5, 6, 7 */
		const double o5 = L->bk5 * density23 * ts42 * ts4;
		const double o6 = L->bk6 * density13 * ts48 * ts42;
		const double o7 = L->bk7 * density / (ts42 * sqrt(ts4));
		const double o6an = o6 * o6, o7an = o7 * o7;
		return pow(pow(o6an * o7an / (o6an + o7an), 2) + pow(o5 / (1.0 + pow(ts4 / (1.1 * pow(density, 0.04762)), 10.0)), 4.0), 0.25);
	    }
	    { /* laws 7, 8 */
		const double o7 = L->bk7 * density / (ts42 * sqrt(ts4));
		const double o8 = L->bk8;
		const double o7an = o7 * o7, o8an = o8 * o8;
		return pow(o7an * o7an + o8an * o8an, 0.25);
	    }
	}
	{ /* laws 3, 4, 5 */
	    const double o3 = L->bell ? L->bk3 * sqrt(ts4) : L->bk3 * ts4;
	    const double o4 = L->bell ? L->bk4 * density / (ts48 * ts48 * ts48) : L->bk4 * density23 / (ts48 * ts4);
	    const double o5 = L->bk5 * density23 * ts42 * ts4;
	    const double o4an = pow(o4, 4), o3an = pow(o3, 4);
	    const double damp = L->bell ? (1 + 6.561e-5 / ts48 * 1e2 * density23) : (1.0 + 6.561e-5 / ts48);
	    return pow((o4an * o3an / (o4an + o3an)) + pow(o5 / damp, 4), 0.25);
	}
    }
    { /* laws 1, 2, 3: powers of the temperature itself */
	const double t2 = temperature * temperature;
	const double t4 = t2 * t2;
	const double t8 = t4 * t4;
	const double t10 = t8 * t2;
	const double o1 = L->ak1 * t2;
	const double o2 = L->ak2 * temperature / t8;
	const double o3 = L->bell ? L->ak3 * sqrt(temperature) : L->ak3 * temperature;
	const double o1an = o1 * o1, o2an = o2 * o2;
	return pow(pow(o1an * o2an / (o1an + o2an), 2) + pow(o3 / (1 + 1.e22 / t10), 4), 0.25);
    }
}
/* opacity::opacity (opacity.cpp:11-44), code units in and out */
static double opacity_code(const fargo_oracle *o, double density, double temperature)
{
    const double temperatureCGS = temperature * o->p.temperature_cgs;
    const double densityCGS = density * o->p.density_cgs;
    double rv;
    switch (o->p.opacity) {
    case FARGO_OPACITY_LIN: rv = opacity_table(&OP_LIN, densityCGS, temperatureCGS) * o->p.opacity_code; break;
    case FARGO_OPACITY_BELL: rv = opacity_table(&OP_BELL, densityCGS, temperatureCGS) * o->p.opacity_code; break;
    case FARGO_OPACITY_CONST: rv = o->p.kappa_const; break;
    default: rv = o->p.kappa_const * pow(temperatureCGS, 2); break; /* Simple */
    }
    return o->p.kappa_factor * rv;
}

/* compute::midplane_density + compute::kappa_eff (compute.cpp:17-88): tau_eff of every cell from the stored temperature and a
 * freshly recomputed scale height */
static void compute_tau_eff(fargo_oracle *o)
{
    compute_scale_height(o);
    const size_t n = (size_t)o->nr * o->ns;
#pragma omp parallel for
    for (size_t c = 0; c < n; ++c) {
	const double rho = o->sigma[c] / (o->p.density_factor * o->scale_height[c]);
	const double kappa = opacity_code(o, rho, o->temperature[c]);
	const double tau = o->p.tau_factor * (1.0 / o->p.density_factor) * kappa * o->sigma[c];
	if (o->p.heating_star)
	    o->tau_eff[c] = 3.0 / 8.0 * tau + 0.5 + 1.0 / (4.0 * tau + o->p.tau_min);
	else
	    o->tau_eff[c] = 3.0 / 8.0 * tau + sqrt(3.0) / 4.0 + 1.0 / (4.0 * tau + o->p.tau_min);
	if (o->p.opacity == FARGO_OPACITY_SIMPLE)
	    o->tau_eff[c] = 3.0 / 8.0 * tau;
    }
}

/* thermal_cooling, SourceEuler.cpp:693-723 */
static void thermal_cooling(fargo_oracle *o)
{
    const int Nr = o->nr - 1, Nphi = o->ns;
    compute_tau_eff(o);
#pragma omp parallel for
    for (int nr = 1; nr < Nr; ++nr) {
	for (int naz = 0; naz < Nphi; ++naz) {
	    const size_t c = IDX(o, nr, naz);
	    const double T4 = pow(o->temperature[c], 4);
	    const double Tmin4 = pow(o->p.minimum_temperature, 4);
	    o->qminus[c] += o->p.surface_cooling_factor * 2 * o->p.sigma_sb * (T4 - Tmin4) / o->tau_eff[c];
	}
    }
}

/* scurve_cooling, SourceEuler.cpp:726-831 */
static void scurve_cooling(fargo_oracle *o)
{
    const double SigmaCGS_threshold = 2.0;
    const double temperatureCGS_threshold = 1200.0;
    double muExponent, F_hot_const;
    if (o->p.cooling_scurve == 2) { /* Kimura et al. 2020 */
	F_hot_const = 23.405;
	muExponent = 0.31;
    } else { /* Ichikawa & Osaki 1992 */
	F_hot_const = 25.49;
	muExponent = -0.31;
    }
    const int Nr = o->nr - 1, Nphi = o->ns;
    const double energy_flux_cgs_to_code = 1.0 / o->p.energy_flux_cgs; /* t_unit::m_inverse_cgs_factor (units.cpp:213) */
#pragma omp parallel for
    for (int nr = 1; nr < Nr; ++nr) {
	for (int naz = 0; naz < Nphi; ++naz) {
	    const size_t c = IDX(o, nr, naz);
	    const double Sigma = o->sigma[c];
	    const double SigmaCGS = Sigma * o->p.surface_density_cgs;
	    const double SigmaCGS_tmp = stdmax(SigmaCGS, SigmaCGS_threshold);
	    const double temperatureCGS = o->temperature[c] * o->p.temperature_cgs;
	    const double temperatureCGS_tmp = stdmax(temperatureCGS, temperatureCGS_threshold);
	    const double rCGS = o->rmed[nr] * o->p.length_cgs;
	    const double mu = MUC(o, c);
	    const double M = o->p.hydro_center_mass * o->p.mass_cgs;
	    const double cgs_G = o->p.G_cgs;
	    const double omega_keplerCGS = sqrt(cgs_G * M / (rCGS * rCGS * rCGS));
	    const double sigma_sb_cgs = o->p.sigma_sb_cgs;
	    const double logTA = -1.0 / 5.49 * (0.62 * log10(omega_keplerCGS) + 1.62 * log10(SigmaCGS_tmp) + muExponent * log10(mu) - 25.48 -
					      log10(sigma_sb_cgs));
	    const double TA = pow(10.0, logTA);
	    const double FA = sigma_sb_cgs * pow(TA, 4);
	    const double logFA = log10(FA);
	    const double KCGS = 11.0 + 0.4 * log10(2.0e10 / rCGS);
	    const double logFB = stdmax(KCGS, logFA);
	    const double logTB_aux = log10(omega_keplerCGS) + 2.0 * log10(SigmaCGS_tmp) + 0.5 * log10(mu) + F_hot_const;
	    const double logTB = (logFB + logTB_aux) / 8.0;
	    const double TB = pow(10.0, logTB);
	    double logFtot;
	    if (temperatureCGS_tmp < TA) {
		logFtot = 9.49 * log10(temperatureCGS_tmp) + 0.62 * log10(omega_keplerCGS) + 1.62 * log10(SigmaCGS_tmp) +
			  muExponent * log10(mu) - 25.48;
	    } else if (temperatureCGS_tmp > TB) {
		logFtot = 8.0 * log10(temperatureCGS_tmp) - log10(omega_keplerCGS) - 2.0 * log10(SigmaCGS_tmp) - 0.5 * log10(mu) - F_hot_const;
	    } else {
		logFtot = (logFA - logFB) * log10(temperatureCGS_tmp / TB) / log10(TA / TB) + logFB;
	    }
	    const double T4 = pow(o->temperature[c], 4);
	    const double sigma_sb = o->p.sigma_sb;
	    const double factor = o->p.surface_cooling_factor;
	    double F_tot = pow(10.0, logFtot) * energy_flux_cgs_to_code;
	    F_tot *= pow((SigmaCGS / SigmaCGS_tmp), 0.5);
	    F_tot *= pow(temperatureCGS / temperatureCGS_tmp, 2);
	    const double F_Blackbody = sigma_sb * T4;
	    const double qminus_scurve = 2.0 * factor * stdmin(F_tot, F_Blackbody);
	    o->qminus[c] += qminus_scurve;
	    o->tau_eff[c] = factor * 2 * sigma_sb * T4 / qminus_scurve;
	}
    }
}

/* irradiation_single, SourceEuler.cpp:538-596, for body k */
static void irradiation_single(fargo_oracle *o, int k)
{
    const fargo_bodies *b = &o->bodies;
    const double ramping = b->irradiation_ramp[k];
    const double x = b->x[k], y = b->y[k], R_star = b->radius[k], T_star = b->temperature[k];
    const double min_dist = (x * x + y * y > 1e-10) ? stdmax(R_star, b->cubic_smoothing_radius[k]) : R_star;
    const int Nrad = o->nr - 1, Naz = o->ns - 1;
#pragma omp parallel for
    for (int nrad = 1; nrad < Nrad; ++nrad) {
	for (int naz = 0; naz <= Naz; ++naz) {
	    const size_t c = IDX(o, nrad, naz);
	    const double xc = o->rmed[nrad] * o->cosphi[naz], yc = o->rmed[nrad] * o->sinphi[naz];
	    const double distance_measured = sqrt(pow(x - xc, 2) + pow(y - yc, 2));
	    const double distance = stdmax(distance_measured, min_dist);
	    const double roverd = distance < R_star ? 1.0 : R_star / distance;
	    const double HoverR = o->scale_height[c] / o->rmed[nrad]; /* ASPECTRATIO of compute_scale_height, SourceEuler.cpp:1148-1151 */
	    const double eps = 0.5;
	    const double dlogH_dlogr = 9.0 / 7.0;
	    const double W_G = 0.4 * roverd + HoverR * (dlogH_dlogr - 1.0);
	    const double T_irrad_pow4 = (1.0 - eps) * pow(T_star, 4) * pow(roverd, 2) * W_G;
	    const double qplus = 2.0 * o->p.sigma_sb * T_irrad_pow4 / o->tau_eff[c];
	    o->qplus[c] += ramping * qplus;
	}
    }
}

/* ------------------------------------------------------------------------------------------
 * energy sources: calculate_qminus :834, calculate_qplus :614, SubStep3 :859 (SourceEuler.cpp) */
static void calculate_qminus(fargo_oracle *o)
{
    const int Nr = o->nr - 1, Nphi = o->ns;
    memset(o->qminus, 0, (size_t)o->nr * o->ns * sizeof(double));
    /* thermal_relaxation :632-690 */
#pragma omp parallel for
    for (int nr = 1; nr < (o->p.cooling_beta ? Nr : 0); ++nr) {
	for (int naz = 0; naz < Nphi; ++naz) {
	    const size_t c = IDX(o, nr, naz);
	    const double omega_k = omega_kepler(o, o->rmed[nr]);
	    const double E = o->energy[c];
	    const double t_ramp_up = o->p.cooling_beta_ramp_up;
	    double beta_inv = 1 / o->p.cooling_beta_value;
	    if (t_ramp_up > 0.0) {
		const double a = 2 * o->time / t_ramp_up;
		const double ramp_factor = 1 - exp(-(a * a));
		beta_inv = beta_inv * ramp_factor;
	    }
	    double delta_E = E;
	    if (o->p.cooling_beta_reference & FARGO_BETA_REF_REFERENCE)
		delta_E -= o->energy0[c] / o->sigma0[c] * o->sigma[c];
	    if (o->p.cooling_beta_reference & FARGO_BETA_REF_MODEL) {
		const double h = o->p.aspectratio_ref;
		const double E0 = 1.0 / (o->p.gamma - 1.0) * (h * h) * pow(o->rmed[nr], 2.0 * o->p.flaring_index - 1.0) * o->p.G *
				  o->p.hydro_center_mass * o->sigma[c];
		delta_E -= E0;
	    }
	    if (o->p.cooling_beta_reference & FARGO_BETA_REF_FLOOR) {
		const double minimum_energy = o->p.minimum_temperature * o->sigma[c] / MUC(o, c) * o->p.Rgas / (GEFF(o, c) - 1.0);
		delta_E -= minimum_energy;
	    }
	    o->qminus[c] += delta_E * omega_k * beta_inv;
	}
    }
    if (o->p.cooling_surface)
	thermal_cooling(o);
    if (o->p.cooling_scurve)
	scurve_cooling(o);
}

static void calculate_qplus(fargo_oracle *o)
{
    const int Nr_m1 = o->nr - 1, Nphi = o->ns;
    memset(o->qplus, 0, (size_t)o->nr * o->ns * sizeof(double));
    /* viscous_heating :496-536 */
#pragma omp parallel for
    for (int nr = 1; nr < (o->p.heating_viscous ? Nr_m1 : 0); ++nr) {
	for (int naz = 0; naz < Nphi; ++naz) {
	    const size_t c = IDX(o, nr, naz);
	    if (o->viscosity[c] != 0.0) {
		const int naz_next = (naz == Nphi - 1 ? 0 : naz + 1);
		const double tau_r_phi = 0.25 * (o->trp[c] + o->trp[IDX(o, nr + 1, naz)] + o->trp[IDX(o, nr, naz_next)] + o->trp[IDX(o, nr + 1, naz_next)]);
		double qplus = 1.0 / (2.0 * o->viscosity[c] * o->sigma[c]) * (o->trr[c] * o->trr[c] + 2 * (tau_r_phi * tau_r_phi) + o->tpp[c] * o->tpp[c]);
		qplus += (2.0 / 9.0) * o->viscosity[c] * o->sigma[c] * (o->divv[c] * o->divv[c]);
		qplus *= o->p.heating_viscous_factor;
		o->qplus[c] += qplus;
	    }
	}
    }
    if (o->p.heating_star) { /* calculate_qplus :621-627 */
	if (!(o->p.cooling_surface || o->p.cooling_scurve))
	    compute_tau_eff(o);
	for (int k = 0; k < o->bodies.n; ++k)
	    if (o->bodies.temperature[k] > 0)
		irradiation_single(o, k);
    }
}

/* the alpha_r division shared by SubStep3 :921-927 and compute_heating_cooling_for_CFL :1440-1446 */
static inline double radiative_alpha(const fargo_oracle *o, size_t c, double H, double sigma, double energy)
{
    const double inv_pow4 = pow(MUC(o, c) * (GEFF(o, c) - 1.0) / (o->p.Rgas * sigma), 4);
    return 1.0 + 2.0 * H * 4.0 * o->p.sigma_sb / o->p.c_light * inv_pow4 * pow(energy, 3);
}

int fargo_oracle_stage_substep3(fargo_oracle *o, double dt)
{
    if (!o->p.adiabatic)
	return 0;
    const int Nr = o->nr, Nphi = o->ns;
    compute_temperature(o);
    calculate_qminus(o);
    calculate_qplus(o);
#pragma omp parallel for
    for (int nr = 1; nr < Nr - 1; ++nr) {
	for (int naz = 0; naz < Nphi; ++naz) {
	    const size_t c = IDX(o, nr, naz);
	    const double sigma = o->sigma[c], energy = o->energy[c];
	    const double alpha = radiative_alpha(o, c, o->scale_height[c], sigma, energy);
	    o->qplus[c] /= alpha;
	    o->qminus[c] /= alpha;
	    const double Qplus = o->qplus[c], Qminus = o->qminus[c];
	    double energy_new = energy + dt * (Qplus - Qminus);
	    const double SigmaFloor = 10.0 * o->p.sigma0 * o->p.sigma_floor;
	    if (sigma < SigmaFloor) {
		const double e4 = Qplus * o->tau_eff[c] / (2.0 * o->p.sigma_sb);
		const double constant = (o->p.Rgas / MUC(o, c) * sigma / (GEFF(o, c) - 1.0));
		const double eq_energy = pow(e4, 1.0 / 4.0) * constant;
		o->qminus[c] = Qplus;
		energy_new = eq_energy;
	    }
	    o->energy[c] = energy_new;
	}
    }
    assure_temperature_range(o);
    return 0;
}

/* compute_heating_cooling_for_CFL, SourceEuler.cpp:1410-1450 */
static void compute_heating_cooling_for_CFL(fargo_oracle *o)
{
    if (!o->p.adiabatic)
	return;
    update_viscosity(o);
    compute_viscous_stress_tensor(o);
    calculate_qminus(o);
    calculate_qplus(o);
    const int Nr = o->nr - 1, Nphi = o->ns;
#pragma omp parallel for
    for (int nr = 1; nr < Nr; ++nr) {
	for (int naz = 0; naz < Nphi; ++naz) {
	    const size_t c = IDX(o, nr, naz);
	    const double alpha = radiative_alpha(o, c, o->scale_height[c], o->sigma[c], o->energy[c]);
	    o->qplus[c] /= alpha;
	    o->qminus[c] /= alpha;
	}
    }
}

/* init_euler's derived fields, SourceEuler.cpp:264-284 */
int fargo_oracle_init_derived(fargo_oracle *o)
{
    if (!o->p.adiabatic) {
	compute_sound_speed(o);
	compute_pressure(o);
	compute_temperature(o);
	compute_scale_height(o);
    } else {
	if (o->p.pvte) { /* init_euler, SourceEuler.cpp:272-276: the arrays still hold the constant gamma / mu of init_eos_arrays */
	    compute_sound_speed(o);
	    compute_scale_height(o);
	    compute_gamma_mu(o);
	}
	compute_temperature(o);
	compute_sound_speed(o);
	compute_scale_height(o);
	compute_pressure(o);
    }
    update_viscosity(o);
    compute_heating_cooling_for_CFL(o);
    return 0;
}

/* init_eos_arrays (init.cpp:1190-1206): build the lookup tables, fill GAMMAEFF / GAMMA1 / MU with the constant values.
 * Must be called before fargo_oracle_init_derived when params.pvte is set. */
int fargo_oracle_set_pvte(fargo_oracle *o, const fargo_pvte_consts *k)
{
    const size_t n = (size_t)o->nr * o->ns;
    if (!o->p.pvte)
	return 1;
    fargo_pvte_free(o->pv);
    o->pv = fargo_pvte_build(k);
    if (!o->pv)
	return 1;
    if (!o->gamma_eff) {
	o->gamma_eff = dalloc(n);
	o->mu_cell = dalloc(n);
	o->gamma1 = dalloc(n);
    }
    for (size_t c = 0; c < n; ++c) {
	o->gamma_eff[c] = o->p.gamma;
	o->gamma1[c] = o->p.gamma;
	o->mu_cell[c] = o->p.mu;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * boundary conditions, boundary_conditions.cpp:65-114 and the per-variable functions */
static int rmed_id(const fargo_oracle *o, double r)
{ /* find_cell_id.cpp:218-257, 382-412: largest local id with Rmed[id] <= r */
    if (o->p.radial_spacing == FARGO_SPACING_LOG) {
	const double gf = pow(o->p.rmax / o->p.rmin, 1.0 / ((double)o->p.nrad - 2.0)); /* init.cpp:94 */
	const double optimization_const = 3.0 / 2.0 / o->p.rmin * (1 - pow(gf, 2.0)) / (1 - pow(gf, 3.0));
	const double inv_log_gf = 1.0 / log(gf);
	const double did = log(r * optimization_const) * inv_log_gf;
	return (int)floor(did) - o->imin + 1;
    }
    int id = 0;
    while (id < o->nr + 1 && o->rmed[id] < r)
	id++;
    return id - 1;
}
static int rinf_id(const fargo_oracle *o, double r)
{ /* find_cell_id.cpp:259-291 */
    if (o->p.radial_spacing == FARGO_SPACING_LOG) {
	const double gf = pow(o->p.rmax / o->p.rmin, 1.0 / ((double)o->p.nrad - 2.0));
	const double inv_log_gf = 1.0 / log(gf);
	const double did = log(r / o->p.rmin) * inv_log_gf;
	return (int)floor(did) - o->imin + 1;
    }
    int id = 0;
    while (id < o->nr + 1 && o->rinf[id] < r)
	id++;
    return id - 1;
}
static int clamp_id(const fargo_oracle *o, int id, int is_vector)
{ /* find_cell_id.cpp:14-38 */
    const int mx = o->nr - (is_vector ? 0 : 1);
    return id < 0 ? 0 : (id > mx ? mx : id);
}

/* damping_single_{inner,outer}{,_zero,_mean}, damping.cpp:311-752 */
static void damp_field(fargo_oracle *o, double *x, double *x0, int is_vector, int is_density, const int type[2], double dt)
{
    const int rings = o->nr + (is_vector ? 1 : 0), Nphi = o->ns;
    const double *radius = is_vector ? o->rinf : o->rmed;
    const double RMIN = o->p.rmin, RMAX = o->p.rmax;
    if (type[0] != FARGO_DAMP_NONE && (o->p.damping_inner_limit > 1.0) && (radius[0] < RMIN * o->p.damping_inner_limit)) {
	const int limit = is_vector ? clamp_id(o, rinf_id(o, RMIN * o->p.damping_inner_limit), 1)
				    : clamp_id(o, rmed_id(o, RMIN * o->p.damping_inner_limit), 0);
	const double tau = o->p.damping_time_factor * 2.0 * M_PI / omega_kepler(o, RMIN);
	double created = 0.0, removed = 0.0; /* the privates of `reduction(+ : ...)` (damping.cpp:338), added to MassDelta at the end */
	for (int nr = 0; nr <= limit; ++nr) {
	    const double q = (radius[nr] - RMIN * o->p.damping_inner_limit) / (RMIN - RMIN * o->p.damping_inner_limit);
	    const double factor = q * q;
	    const double exp_factor = exp(-dt * factor / tau);
	    double mean = 0.0;
	    if (type[0] == FARGO_DAMP_MEAN) {
		for (int j = 0; j < Nphi; ++j)
		    mean += x[IDX(o, nr, j)];
		mean /= Nphi;
		x0[IDX(o, nr, 0)] = mean; /* the reference keeps the mean IN quantity0(n_radial, 0), damping.cpp:578-585 / 706-713: the
					     initial-value grid of this zone carries it from then on (reference boundaries, beta cooling) */
	    }
	    for (int j = 0; j < Nphi; ++j) {
		const size_t c = IDX(o, nr, j);
		const double X = x[c];
		double X0;
		if (type[0] == FARGO_DAMP_INITIAL)
		    X0 = x0[c];
		else if (type[0] == FARGO_DAMP_MEAN)
		    X0 = mean;
		else
		    X0 = is_density ? o->p.sigma_floor * o->p.sigma0 : 0.0;
		const double Xnew = (X - X0) * exp_factor + X0;
		x[c] = Xnew;
		const double delta = Xnew - X;
		if (is_density && o->first_active <= nr && nr < o->active_size) { /* sum_without_ghost_cells, damping.cpp:345-357 */
		    if (delta > 0)
			created += delta * o->surf[nr];
		    else
			removed += -delta * o->surf[nr];
		}
	    }
	}
	o->dmass[0] += created;
	o->dmass[1] += removed;
    }
    if (type[1] != FARGO_DAMP_NONE && (o->p.damping_outer_limit < 1.0) && (radius[rings - 1] > RMAX * o->p.damping_outer_limit)) {
	const int limit = is_vector ? clamp_id(o, rinf_id(o, RMAX * o->p.damping_outer_limit) + 1, 1)
				    : clamp_id(o, rmed_id(o, RMAX * o->p.damping_outer_limit) + 1, 0);
	const double tau = o->p.damping_time_factor * 2.0 * M_PI / omega_kepler(o, o->p.damping_time_radius_outer);
	double created = 0.0, removed = 0.0;
	for (int nr = limit; nr < rings; ++nr) {
	    const double q = (radius[nr] - RMAX * o->p.damping_outer_limit) / (RMAX - RMAX * o->p.damping_outer_limit);
	    const double factor = q * q;
	    const double exp_factor = exp(-dt * factor / tau);
	    double mean = 0.0;
	    if (type[1] == FARGO_DAMP_MEAN) {
		for (int j = 0; j < Nphi; ++j)
		    mean += x[IDX(o, nr, j)];
		mean /= Nphi;
		x0[IDX(o, nr, 0)] = mean; /* the reference keeps the mean IN quantity0(n_radial, 0), damping.cpp:578-585 / 706-713: the
					     initial-value grid of this zone carries it from then on (reference boundaries, beta cooling) */
	    }
	    for (int j = 0; j < Nphi; ++j) {
		const size_t c = IDX(o, nr, j);
		const double X = x[c];
		double X0;
		if (type[1] == FARGO_DAMP_INITIAL)
		    X0 = x0[c];
		else if (type[1] == FARGO_DAMP_MEAN)
		    X0 = mean;
		else
		    X0 = is_density ? o->p.sigma_floor * o->p.sigma0 : 0.0;
		const double Xnew = (X - X0) * exp_factor + X0;
		x[c] = Xnew;
		const double delta = Xnew - X;
		if (is_density && o->first_active <= nr && nr < o->active_size) { /* sum_without_ghost_cells, damping.cpp:345-357 */
		    if (delta > 0)
			created += delta * o->surf[nr];
		    else
			removed += -delta * o->surf[nr];
		}
	    }
	}
	o->dmass[2] += created;
	o->dmass[3] += removed;
    }
}

static void bc_scalar(fargo_oracle *o, double *x, const double *x0, const int bc[2])
{
    const int Nphi = o->ns, Irad = o->nr - 1;
    if (o->rank == 0) {
	if (bc[0] == FARGO_BC_ZEROGRADIENT) /* zero_gradient.cpp:17-27 */
	    for (int j = 0; j < Nphi; ++j)
		x[IDX(o, 0, j)] = x[IDX(o, 1, j)];
	else if (bc[0] == FARGO_BC_REFERENCE) /* reference.cpp:16-25 */
	    for (int j = 0; j < Nphi; ++j)
		x[IDX(o, 0, j)] = x0[IDX(o, 0, j)];
    }
    if (o->rank == o->nranks - 1) {
	if (bc[1] == FARGO_BC_ZEROGRADIENT) /* zero_gradient.cpp:56-67 */
	    for (int j = 0; j < Nphi; ++j)
		x[IDX(o, Irad, j)] = x[IDX(o, Irad - 1, j)];
	else if (bc[1] == FARGO_BC_REFERENCE)
	    for (int j = 0; j < Nphi; ++j)
		x[IDX(o, Irad, j)] = x0[IDX(o, Irad, j)];
    }
}

static void bc_vrad(fargo_oracle *o)
{
    const int Nphi = o->ns, Irad = o->nr; /* vector grid: max_radial = Nrad */
    double *v = o->vrad;
    const int first = (o->rank == 0), last = (o->rank == o->nranks - 1);
    switch (o->p.bc_vrad[0]) {
    case FARGO_BC_ZEROGRADIENT: /* zero_gradient.cpp:29-40 */
	if (first)
	    for (int j = 0; j < Nphi; ++j) {
		v[IDX(o, 0, j)] = v[IDX(o, 2, j)];
		v[IDX(o, 1, j)] = v[IDX(o, 2, j)];
	    }
	break;
    case FARGO_BC_OUTFLOW: /* outflow.cpp:16-35 */
	if (first)
	    for (int j = 0; j < Nphi; ++j) {
		if (v[IDX(o, 2, j)] > 0.0) {
		    v[IDX(o, 1, j)] = 0.0;
		    v[IDX(o, 0, j)] = 0.0;
		} else {
		    v[IDX(o, 1, j)] = v[IDX(o, 2, j)];
		    v[IDX(o, 0, j)] = v[IDX(o, 2, j)];
		}
	    }
	break;
    case FARGO_BC_REFLECTING: /* reflecting.cpp:15-26 — NO rank guard in the reference */
	for (int j = 0; j < Nphi; ++j) {
	    v[IDX(o, 0, j)] = -v[IDX(o, 2, j)];
	    v[IDX(o, 1, j)] = 0;
	}
	break;
    case FARGO_BC_REFERENCE: /* reference.cpp:27-38 */
	if (first)
	    for (int j = 0; j < Nphi; ++j) {
		v[IDX(o, 0, j)] = o->vrad0[IDX(o, 0, j)];
		v[IDX(o, 1, j)] = o->vrad0[IDX(o, 1, j)];
	    }
	break;
    case FARGO_BC_KEPLERIAN: /* keplerian_radial.cpp:18-39 */
	if (first)
	    for (int j = 0; j < Nphi; ++j)
		for (int k = 0; k <= 1; k++) {
		    const double vKep = sqrt(o->p.G * o->p.hydro_center_mass / o->rmed[k]);
		    v[IDX(o, k, j)] = o->p.keplerian_radial_factor[0] * vKep;
		}
	break;
    case FARGO_BC_VISCOUS: /* viscous.cpp:18-46: the VISCOSITY grid as last stored */
	if (first)
	    for (int j = 0; j < Nphi; ++j) {
		const double s = o->p.viscous_outflow_speed;
		const double Nu0 = o->viscosity[IDX(o, 0, j)], Nu1 = o->viscosity[IDX(o, 1, j)];
		const double Nu = 0.5 * (Nu0 + Nu1);
		v[IDX(o, 1, j)] = -1.5 * s / o->rinf[1] * Nu;
		v[IDX(o, 0, j)] = -1.5 * s / o->rinf[0] * Nu;
	    }
	break;
    default:
	break;
    }
    switch (o->p.bc_vrad[1]) {
    case FARGO_BC_ZEROGRADIENT: /* zero_gradient.cpp:69-81 */
	if (last)
	    for (int j = 0; j < Nphi; ++j) {
		v[IDX(o, Irad, j)] = v[IDX(o, Irad - 2, j)];
		v[IDX(o, Irad - 1, j)] = v[IDX(o, Irad - 2, j)];
	    }
	break;
    case FARGO_BC_OUTFLOW: /* outflow.cpp:37-57 */
	if (last)
	    for (int j = 0; j < Nphi; ++j) {
		if (v[IDX(o, Irad - 2, j)] < 0.0) {
		    v[IDX(o, Irad - 1, j)] = 0.0;
		    v[IDX(o, Irad, j)] = 0.0;
		} else {
		    v[IDX(o, Irad - 1, j)] = v[IDX(o, Irad - 2, j)];
		    v[IDX(o, Irad, j)] = v[IDX(o, Irad - 2, j)];
		}
	    }
	break;
    case FARGO_BC_REFLECTING: /* reflecting.cpp:28-40 — NO rank guard */
	for (int j = 0; j < Nphi; ++j) {
	    v[IDX(o, Irad, j)] = -v[IDX(o, Irad - 2, j)];
	    v[IDX(o, Irad - 1, j)] = 0;
	}
	break;
    case FARGO_BC_REFERENCE:
	if (last)
	    for (int j = 0; j < Nphi; ++j) {
		v[IDX(o, Irad, j)] = o->vrad0[IDX(o, Irad, j)];
		v[IDX(o, Irad - 1, j)] = o->vrad0[IDX(o, Irad - 1, j)];
	    }
	break;
    default:
	break;
    }
}

static void bc_vazi(fargo_oracle *o)
{
    const int Nphi = o->ns, Irad = o->nr - 1;
    const double OmegaF = o->bodies.omega_frame;
    if (o->rank == 0) {
	if (o->p.bc_vazi[0] == FARGO_BC_KEPLERIAN) { /* keplerian_azimuthal.cpp:19-39 */
	    const double vKep = sqrt(o->p.G * o->p.hydro_center_mass / o->rmed[0]);
	    const double val = o->p.keplerian_azimuthal_factor[0] * vKep - o->rmed[0] * OmegaF;
	    for (int j = 0; j < Nphi; ++j)
		o->vazi[IDX(o, 0, j)] = val;
	} else if (o->p.bc_vazi[0] == FARGO_BC_ZEROGRADIENT) {
	    for (int j = 0; j < Nphi; ++j)
		o->vazi[IDX(o, 0, j)] = o->vazi[IDX(o, 1, j)];
	} else if (o->p.bc_vazi[0] == FARGO_BC_REFERENCE) {
	    for (int j = 0; j < Nphi; ++j)
		o->vazi[IDX(o, 0, j)] = o->vazi0[IDX(o, 0, j)];
	} else if (o->p.bc_vazi[0] == FARGO_BC_BALANCED) { /* balanced.cpp:23-75 */
	    double vaz_balanced = sqrt(o->p.balanced_vazi_sq[0]);
	    vaz_balanced -= o->rmed[0] * OmegaF;
	    for (int j = 0; j < Nphi; ++j)
		o->vazi[IDX(o, 0, j)] = vaz_balanced;
	} else if (o->p.bc_vazi[0] == FARGO_BC_ZEROSHEAR) { /* zero_shear.cpp:20-36 */
	    for (int j = 0; j < Nphi; ++j) {
		const double Omega_active = o->vazi[IDX(o, 1, j)] / o->rmed[1];
		o->vazi[IDX(o, 0, j)] = o->rmed[0] * Omega_active;
	    }
	}
    }
    if (o->rank == o->nranks - 1) {
	if (o->p.bc_vazi[1] == FARGO_BC_KEPLERIAN) { /* keplerian_azimuthal.cpp:41-60 */
	    const double vKep = sqrt(o->p.G * o->p.hydro_center_mass / o->rmed[Irad]);
	    const double val = o->p.keplerian_azimuthal_factor[1] * vKep - o->rmed[Irad] * OmegaF;
	    for (int j = 0; j < Nphi; ++j)
		o->vazi[IDX(o, Irad, j)] = val;
	} else if (o->p.bc_vazi[1] == FARGO_BC_ZEROGRADIENT) {
	    for (int j = 0; j < Nphi; ++j)
		o->vazi[IDX(o, Irad, j)] = o->vazi[IDX(o, Irad - 1, j)];
	} else if (o->p.bc_vazi[1] == FARGO_BC_REFERENCE) {
	    for (int j = 0; j < Nphi; ++j)
		o->vazi[IDX(o, Irad, j)] = o->vazi0[IDX(o, Irad, j)];
	} else if (o->p.bc_vazi[1] == FARGO_BC_BALANCED) {
	    double vaz_balanced = sqrt(o->p.balanced_vazi_sq[1]);
	    vaz_balanced -= o->rmed[Irad] * OmegaF;
	    for (int j = 0; j < Nphi; ++j)
		o->vazi[IDX(o, Irad, j)] = vaz_balanced;
	} else if (o->p.bc_vazi[1] == FARGO_BC_ZEROSHEAR) { /* zero_shear.cpp:38-54 */
	    for (int j = 0; j < Nphi; ++j) {
		const double Omega_active = o->vazi[IDX(o, Irad - 1, j)] / o->rmed[Irad - 1];
		o->vazi[IDX(o, Irad, j)] = o->rmed[Irad] * Omega_active;
	    }
	}
    }
}

int fargo_oracle_stage_boundary(fargo_oracle *o, double dt, int final_call)
{
    if (final_call && o->p.damping) { /* handle_damping + damping(), order vrad, vazi, sigma, energy (damping.cpp:204-270) */
	damp_field(o, o->vrad, o->vrad0, 1, 0, o->p.damp_vrad, dt);
	damp_field(o, o->vazi, o->vazi0, 0, 0, o->p.damp_vazi, dt);
	damp_field(o, o->sigma, o->sigma0, 0, 1, o->p.damp_sigma, dt);
	if (o->p.adiabatic) /* Interpret.cpp:559-565 */
	    damp_field(o, o->energy, o->energy0, 0, 0, o->p.damp_energy, dt);
    }
    bc_scalar(o, o->sigma, o->sigma0, o->p.bc_sigma);
    bc_scalar(o, o->energy, o->energy0, o->p.bc_energy);
    bc_vrad(o);
    bc_vazi(o);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Transport, TransportEuler.cpp:112-664 */
static inline double flux_limiter(const fargo_oracle *o, double a, double b)
{
    if (o->p.flux_limiter == FARGO_LIMITER_MC) { /* :314-323 */
	const double s = 0.5 * (a + b);
	double mm;
	if (a * b > 0.0)
	    mm = fabs(a) < fabs(b) ? a : b;
	else
	    mm = 0.0;
	const double t = 2.0 * mm;
	if (s * t > 0.0)
	    return fabs(s) < fabs(t) ? s : t;
	return 0.0;
    }
    if (a * b > 0.0) /* :306-312 */
	return 2.0 * a * b / (a + b);
    return 0;
}

/* compute_momenta_from_velocities :471-493 */
static void compute_momenta(fargo_oracle *o)
{
    const int Nr = o->nr, Nphi = o->ns;
    const double OmegaF = o->bodies.omega_frame;
#pragma omp parallel for
    for (int nr = 0; nr < Nr; ++nr) {
	for (int naz = 0; naz < Nphi; ++naz) {
	    const size_t c = IDX(o, nr, naz);
	    o->rmp[c] = o->sigma[c] * o->vrad[IDX(o, nr + 1, naz)];
	    o->rmm[c] = o->sigma[c] * o->vrad[c];
	    const int naz_ind = naz == (Nphi - 1) ? 0 : naz + 1;
	    const double vnext = o->vazi[IDX(o, nr, naz_ind)];
	    const double r = o->rmed[nr];
	    o->amp[c] = o->sigma[c] * (vnext + r * OmegaF) * r;
	    o->amm[c] = o->sigma[c] * (o->vazi[c] + r * OmegaF) * r;
	}
    }
}

/* compute_star_radial :349-406 */
static void compute_star_radial(fargo_oracle *o, const double *qbase, const double *vr, double *qstar, double dt)
{
    const int Nr = o->nr, Nphi = o->ns;
#pragma omp parallel for
    for (int nr = 0; nr < Nr; ++nr) {
	for (int naz = 0; naz < Nphi; ++naz) {
	    const size_t c = IDX(o, nr, naz);
	    if (nr == 0 || nr == Nr - 1) {
		o->dq[c] = 0.0;
	    } else {
		const double dqm = (qbase[c] - qbase[c - Nphi]) * o->invdiffrmed[nr];
		const double dqp = (qbase[c + Nphi] - qbase[c]) * o->invdiffrmed[nr + 1];
		o->dq[c] = flux_limiter(o, dqp, dqm);
	    }
	}
    }
#pragma omp parallel for
    for (int nr = 1; nr < Nr; ++nr) {
	for (int naz = 0; naz < Nphi; ++naz) {
	    const size_t c = IDX(o, nr, naz), cm = c - Nphi;
	    if (vr[c] > 0.0)
		qstar[c] = qbase[cm] + (o->rmed[nr] - o->rmed[nr - 1] - vr[c] * dt) * 0.5 * o->dq[cm];
	    else
		qstar[c] = qbase[c] - (o->rmed[nr + 1] - o->rmed[nr] + vr[c] * dt) * 0.5 * o->dq[c];
	}
    }
    for (int naz = 0; naz < Nphi; ++naz)
	qstar[naz] = 0.0;
}

/* VanLeerRadial :545-620 */
static void vanleer_radial(fargo_oracle *o, const double *vr, double *qbase, double dt, int is_density)
{
    const int Nr = o->nr, Nphi = o->ns;
    const size_t n = (size_t)Nr * Nphi;
#pragma omp parallel for
    for (size_t c = 0; c < n; ++c) /* divise_polargrid, SideEuler.cpp:27-43 */
	o->work[c] = qbase[c] / o->densint[c];
    compute_star_radial(o, o->work, vr, o->qrstar, dt);
#pragma omp parallel for
    for (int nr = 0; nr < Nr; ++nr) {
	for (int naz = 0; naz < Nphi; ++naz) {
	    const size_t c = IDX(o, nr, naz), lip = c + Nphi;
	    const double varq_inf = dt * o->dphi * o->rinf[nr] * o->qrstar[c] * o->densstar[c] * vr[c];
	    const double varq_sup = dt * o->dphi * o->rsup[nr] * o->qrstar[lip] * o->densstar[lip] * vr[lip];
	    qbase[c] += (varq_inf - varq_sup) * o->invsurf[nr];
	    if (is_density && o->track_bflow) { /* parameters::write_disk_quantities, :578-608 */
		if (o->rank == 0 && nr == 1) {
		    if (varq_inf > 0)
			o->bflow[0] += varq_inf;
		    else
			o->bflow[1] += -varq_inf;
		} else if (o->rank == o->nranks - 1 && nr == Nr - 2) {
		    if (varq_sup > 0)
			o->bflow[3] += varq_sup;
		    else
			o->bflow[2] += -varq_sup;
		}
	    }
	    if (is_density && o->massflow) { /* parameters::write_massflow, :610-616 */
		o->massflow[c] += varq_inf;
		if (o->rank == o->nranks - 1 && nr == Nr - 1)
		    o->massflow[c] += varq_sup;
	    }
	}
    }
}

/* ComputeStarTheta :416-466 */
static void compute_star_theta(fargo_oracle *o, const double *qbase, const double *vazi, double *qstar, double dt)
{
    const int Nr = o->nr, Nphi = o->ns;
#pragma omp parallel for
    for (int nr = 0; nr < Nr; ++nr) {
	const double dxtheta = o->dphi * o->rmed[nr];
	const double invdxtheta = 1.0 / dxtheta;
	for (int naz = 0; naz < Nphi; ++naz) {
	    const size_t c = IDX(o, nr, naz);
	    const size_t ljp = (naz == Nphi - 1) ? IDX(o, nr, 0) : c + 1;
	    const size_t ljm = (naz == 0) ? IDX(o, nr, Nphi - 1) : c - 1;
	    const double dqm = (qbase[c] - qbase[ljm]);
	    const double dqp = (qbase[ljp] - qbase[c]);
	    o->dq[c] = 0.5 * flux_limiter(o, dqp, dqm) * invdxtheta;
	}
    }
#pragma omp parallel for
    for (int nr = 0; nr < Nr; ++nr) {
	const double dxtheta = o->dphi * o->rmed[nr];
	for (int naz = 0; naz < Nphi; ++naz) {
	    const size_t c = IDX(o, nr, naz);
	    const size_t ljm = (naz == 0) ? IDX(o, nr, Nphi - 1) : c - 1;
	    const double ksi = vazi[c] * dt;
	    if (ksi > 0.0)
		qstar[c] = qbase[ljm] + (dxtheta - ksi) * o->dq[ljm];
	    else
		qstar[c] = qbase[c] - (dxtheta + ksi) * o->dq[c];
	}
    }
}

/* VanLeerTheta :630-664 */
static void vanleer_theta(fargo_oracle *o, const double *vazi, double *qbase, double dt, int uniform)
{
    const int Nr = o->nr, Nphi = o->ns;
    const size_t n = (size_t)Nr * Nphi;
#pragma omp parallel for
    for (size_t c = 0; c < n; ++c)
	o->work[c] = qbase[c] / o->densint[c];
    compute_star_theta(o, o->work, vazi, o->qrstar, dt);
    const int nosplit = !o->p.fast_transport; /* NoSplitAdvection[i], :225-234 */
#pragma omp parallel for
    for (int nr = 0; nr < Nr; ++nr) {
	const double dxrad = (o->rsup[nr] - o->rinf[nr]) * dt;
	const double invsurf = o->invsurf[nr];
	if (!uniform || !nosplit) {
	    for (int naz = 0; naz < Nphi; ++naz) {
		const size_t c = IDX(o, nr, naz);
		const size_t ljp = (naz == Nphi - 1) ? IDX(o, nr, 0) : c + 1;
		double varq = dxrad * o->qrstar[c] * o->densstar[c] * vazi[c];
		varq -= dxrad * o->qrstar[ljp] * o->densstar[ljp] * vazi[ljp];
		qbase[c] += varq * invsurf;
	    }
	}
    }
}

/* AdvectSHIFT :238-268 */
static void advect_shift(fargo_oracle *o, double *val)
{
    const int nr = o->nr, ns = o->ns;
#pragma omp parallel for
    for (int i = 0; i < nr; i++) {
	for (int j = 0; j < ns; j++) {
	    int ji = j - o->nshift[i];
	    while (ji < 0)
		ji += ns;
	    while (ji >= ns)
		ji -= ns;
	    o->tempshift[IDX(o, i, j)] = val[IDX(o, i, ji)];
	}
    }
    memcpy(val, o->tempshift, (size_t)nr * ns * sizeof(double));
}

/* QuantitiesAdvection :292-304 */
static void quantities_advection(fargo_oracle *o, const double *vazi, double dt, int uniform)
{
    const size_t n = (size_t)o->nr * o->ns;
    compute_star_theta(o, o->sigma, vazi, o->densstar, dt);
    memcpy(o->densint, o->sigma, n * sizeof(double));
    vanleer_theta(o, vazi, o->rmp, dt, uniform);
    vanleer_theta(o, vazi, o->rmm, dt, uniform);
    vanleer_theta(o, vazi, o->amp, dt, uniform);
    vanleer_theta(o, vazi, o->amm, dt, uniform);
    if (o->p.adiabatic)
	vanleer_theta(o, vazi, o->energy, dt, uniform);
    vanleer_theta(o, vazi, o->sigma, dt, uniform);
}

int fargo_oracle_stage_transport(fargo_oracle *o, double dt)
{
    const int Nr = o->nr, Nphi = o->ns;
    const size_t n = (size_t)Nr * Nphi;
    const double OmegaF = o->bodies.omega_frame;
    compute_momenta(o);
    /* OneWindRad :138-167 */
    memset(o->densstar + n, 0, (size_t)Nphi * sizeof(double)); /* ring Nr of the star grids stays 0 (:82-91) */
    memset(o->qrstar + n, 0, (size_t)Nphi * sizeof(double));
    compute_star_radial(o, o->sigma, o->vrad, o->densstar, dt);
    memcpy(o->densint, o->sigma, n * sizeof(double));
    vanleer_radial(o, o->vrad, o->rmp, dt, 0);
    vanleer_radial(o, o->vrad, o->rmm, dt, 0);
    vanleer_radial(o, o->vrad, o->amp, dt, 0);
    vanleer_radial(o, o->vrad, o->amm, dt, 0);
    if (o->p.adiabatic)
	vanleer_radial(o, o->vrad, o->energy, dt, 0);
    vanleer_radial(o, o->vrad, o->sigma, dt, 1);
    /* OneWindTheta :270-288 */
    for (int nr = 0; nr < Nr; ++nr) { /* compute_average_azimuthal_velocity :174-189 */
	double s = 0.0;
	for (int naz = 0; naz < Nphi; ++naz)
	    s += o->vazi[IDX(o, nr, naz)];
	o->vmean[nr] = s / (double)Nphi;
    }
#pragma omp parallel for
    for (int nr = 0; nr < Nr; ++nr) /* compute_residual_velocity :194-205 */
	for (int naz = 0; naz < Nphi; ++naz)
	    o->vres[IDX(o, nr, naz)] = o->vazi[IDX(o, nr, naz)] - o->vmean[nr];
    { /* ComputeConstantResidual :207-236 */
	const double invdt = 1.0 / dt;
	for (int i = 0; i < Nr; i++) {
	    const double Ntilde = o->vmean[i] * o->invrmed[i] * dt * o->invdphi;
	    const double Nround = floor(Ntilde + 0.5);
	    o->nshift[i] = (int)Nround;
	    for (int j = 0; j < Nphi; j++)
		o->vazi[IDX(o, i, j)] = (Ntilde - Nround) * o->rmed[i] * invdt * o->dphi;
	    if (!o->p.fast_transport) {
		for (int j = 0; j < Nphi; j++) {
		    const size_t l = IDX(o, i, j);
		    o->vres[l] = o->vazi[l] + o->vres[l];
		    o->vazi[l] = 0.0;
		}
	    }
	}
    }
    quantities_advection(o, o->vres, dt, 0);
    quantities_advection(o, o->vazi, dt, 1);
    advect_shift(o, o->rmp);
    advect_shift(o, o->rmm);
    advect_shift(o, o->amp);
    advect_shift(o, o->amm);
    if (o->p.adiabatic)
	advect_shift(o, o->energy);
    advect_shift(o, o->sigma);
    /* compute_velocities_from_momenta :498-535 */
#pragma omp parallel for
    for (int nr = 0; nr < Nr; ++nr) {
	for (int naz = 0; naz < Nphi; ++naz) {
	    const int nm = (naz == 0 ? Nphi - 1 : naz - 1);
	    const size_t c = IDX(o, nr, naz), cp = IDX(o, nr, nm);
	    if (nr == 0)
		o->vrad[c] = 0.0;
	    else
		o->vrad[c] = (o->rmp[c - Nphi] + o->rmm[c]) / (o->sigma[c - Nphi] + o->sigma[c]);
	    o->vazi[c] = (o->amp[cp] + o->amm[c]) / (o->sigma[cp] + o->sigma[c]) * o->invrmed[nr] - o->rmed[nr] * OmegaF;
	}
    }
    /* assure_minimum_value(SIGMA, sigma_floor*sigma0) :124-125, SourceEuler.cpp:102-134 */
    const double floorv = o->p.sigma_floor * o->p.sigma0;
#pragma omp parallel for
    for (size_t c = 0; c < n; ++c)
	if (o->sigma[c] < floorv)
	    o->sigma[c] = floorv;
    if (o->p.adiabatic)
	assure_temperature_range(o);
    return 0;
}

int fargo_oracle_get_nshift(fargo_oracle *o, int *out)
{
    memcpy(out, o->nshift, (size_t)o->nr * sizeof(int));
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * cfl::condition_cfl, cfl.cpp:185-382 (local part; the caller does the MIN over ranks) */
int fargo_oracle_condition_cfl(fargo_oracle *o, double *out)
{
    const int Nr = o->nr, Nphi = o->ns;
    const double CFL = o->p.cfl;
    const double C = o->p.artificial_viscosity_factor;
    for (int nr = 0; nr < Nr; ++nr) { /* :196-205 */
	double s = 0.0;
	for (int naz = 0; naz < Nphi; ++naz)
	    s += o->vazi[IDX(o, nr, naz)];
	o->vmean[nr] = s / (double)Nphi;
    }
    const double denom0 = fabs(o->vmean[0] * o->invrmed[0] - o->vmean[1] * o->invrmed[1]) + 1.0e-100;
    double dt_core = CFL * o->dphi / denom0;
    const double lf = o->p.leapfrog ? 0.6 : 1.0;
    for (int nr = o->first_active; nr < o->active_size; ++nr) {
	const double denom = fabs(o->vmean[nr] * o->invrmed[nr] - o->vmean[nr + 1] * o->invrmed[nr + 1]) + 1.0e-100;
	const double shear_dt = CFL * o->dphi / denom;
	if (shear_dt < dt_core)
	    dt_core = shear_dt;
	const double dxRadial = o->rsup[nr] - o->rinf[nr];
	const double dxAzimuthal = o->rmed[nr] * o->dphi;
	const double cell_size = stdmin(dxRadial, dxAzimuthal);
	for (int naz = 0; naz < Nphi; ++naz) {
	    const size_t c = IDX(o, nr, naz), cu = IDX(o, nr + 1, naz);
	    const int naz_next = naz == Nphi - 1 ? 0 : naz + 1;
	    const double vres = o->p.fast_transport ? o->vazi[c] - o->vmean[nr] : o->vazi[c];
	    const double invdt1 = o->soundspeed[c] / cell_size;
	    const double invdt2 = o->vrad[c] / dxRadial;
	    const double invdt3 = vres / dxAzimuthal;
	    double invdt4;
	    if (o->p.artificial_viscosity == FARGO_ARTVISC_SN) {
		double dvRadial = o->vrad[cu] - o->vrad[c];
		double dvAzimuthal = o->vazi[IDX(o, nr, naz_next)] - o->vazi[c];
		dvRadial = (dvRadial > 0.0) ? 0.0 : -dvRadial;
		dvAzimuthal = (dvAzimuthal > 0.0) ? 0.0 : -dvAzimuthal;
		invdt4 = 4.0 * (C * C) * stdmax(dvRadial / dxRadial, dvAzimuthal / dxAzimuthal) * lf;
	    } else {
		const double eps_rr = (o->vrad[cu] - o->vrad[c]) * o->invdiffrsup[nr];
		const double eps_pp =
		    o->invrmed[nr] * ((o->vazi[IDX(o, nr, naz_next)] - o->vazi[c]) * o->invdphi + 0.5 * (o->vrad[cu] + o->vrad[c]));
		const double mdiv_V = -stdmin(eps_rr + eps_pp, 0.0);
		invdt4 = 4.0 * (C * C) * mdiv_V * lf;
	    }
	    const double invdt5 = 4.0 * o->viscosity[c] / (cell_size * cell_size) * lf;
	    double invdt6;
	    if (o->p.adiabatic) {
		const double inv_limit = 1.0 / o->p.heating_cooling_cfl_limit;
		invdt6 = inv_limit * fabs((o->qplus[c] - o->qminus[c]) / o->energy[c]) * lf;
	    } else {
		invdt6 = 0.0;
	    }
	    double dt_cell = CFL / sqrt(invdt1 * invdt1 + invdt2 * invdt2 + invdt3 * invdt3 + invdt4 * invdt4 + invdt5 * invdt5 + invdt6 * invdt6);
	    if (o->p.stabilize_viscosity == 2) {
		const double cc = stdmin(o->cf_phi[c], o->cf_r[c]);
		if (cc != 0.0)
		    dt_cell = stdmin(dt_cell, -CFL / cc);
	    }
	    if (dt_cell < dt_core)
		dt_core = dt_cell;
	}
    }
    *out = dt_core;
    return 0;
}

/* sim::CalculateTimeStep, simulation.cpp:100-118 (single rank; multi-rank callers min-reduce
 * fargo_oracle_condition_cfl themselves and apply the same min) */
int fargo_oracle_cfl(fargo_oracle *o, double *last_dt, double *dt_out)
{
    double cfl_dt;
    fargo_oracle_condition_cfl(o, &cfl_dt);
    const double rv = stdmin(o->p.cfl_max_var * *last_dt, cfl_dt);
    *last_dt = rv;
    *dt_out = rv;
    return 0;
}

/* CommunicateBoundaries (commbound.cpp:98-182) for oracle slabs living in one process:
 * copies this slab's send rings into / out of caller-provided buffers of 4*CPUOVERLAP*naz doubles.
 * side 0 = towards rank-1 (inner), 1 = towards rank+1 (outer). */
int fargo_oracle_halo_pack(fargo_oracle *o, int side, double *buf)
{
    const size_t l = (size_t)FARGO_CPUOVERLAP * o->ns;
    const size_t off = side == 0 ? l : (size_t)(o->nr - 2 * FARGO_CPUOVERLAP) * o->ns;
    memcpy(buf, o->sigma + off, l * sizeof(double));
    memcpy(buf + l, o->vrad + off, l * sizeof(double));
    memcpy(buf + 2 * l, o->vazi + off, l * sizeof(double));
    memcpy(buf + 3 * l, o->energy + off, l * sizeof(double));
    return 0;
}
int fargo_oracle_halo_unpack(fargo_oracle *o, int side, const double *buf)
{
    const size_t l = (size_t)FARGO_CPUOVERLAP * o->ns;
    const size_t off = side == 0 ? 0 : (size_t)(o->nr - FARGO_CPUOVERLAP) * o->ns;
    memcpy(o->sigma + off, buf, l * sizeof(double));
    memcpy(o->vrad + off, buf + l, l * sizeof(double));
    memcpy(o->vazi + off, buf + 2 * l, l * sizeof(double));
    if (o->p.adiabatic)
	memcpy(o->energy + off, buf + 3 * l, l * sizeof(double));
    return 0;
}

/* gas part of step_Euler, simulation.cpp:167-175, 187-218, 230-266.  With nranks > 1 the caller must
 * split the step at the halo exchange: step_pre, exchange, step_post. */
int fargo_oracle_step_pre(fargo_oracle *o, double dt)
{
    fargo_oracle_stage_potential(o);
    fargo_oracle_stage_sources(o, dt);
    fargo_oracle_stage_artvisc(o, dt);
    fargo_oracle_stage_viscosity(o, dt);
    if (o->p.adiabatic)
	fargo_oracle_stage_substep3(o, dt);
    fargo_oracle_stage_boundary(o, 0.0, 0);
    fargo_oracle_stage_transport(o, dt);
    o->time += dt;
    return 0;
}
int fargo_oracle_step_post(fargo_oracle *o, double dt)
{
    fargo_oracle_stage_boundary(o, dt, 1);
    if (o->p.pvte) { /* simulation.cpp:256-262: the scale height after Transport, for the 3-D density of the lookup */
	compute_sound_speed(o);
	compute_scale_height(o);
    }
    fargo_oracle_stage_derived(o);
    return 0;
}
int fargo_oracle_step(fargo_oracle *o, double dt)
{
    fargo_oracle_step_pre(o, dt);
    fargo_oracle_step_post(o, dt);
    return 0;
}

/* accretion::AccreteOntoSinglePlanet (accretion.cpp:84-221, "kley" accretion): a fraction facc1 = facc / 3 of the gas within
 * frac * RHill of the planet and another facc2 = 2 facc / 3 within frac / 2 * RHill leaves the disk (never below the density
 * floor); the energy of an adiabatic disk is scaled with it (zone 2 by 1 - facc2, as the reference does).  facc = dt *
 * accretion efficiency / orbital period * ln 2 and RHill = dimensionless Roche radius * distance to the primary are the
 * N-body side's numbers.  The reference walks a window of cells around the planet that contains every cell inside
 * frac * RHill; testing every cell's distance gives the same cells.  out3 = mass and momentum taken from ACTIVE cells
 * (radial_first_active < i < radial_active_size, :171), summed in index order.  With more than one rank the reference's
 * condition also drops the first OWNED ring of every rank but the first (its only np-dependent result); slabs here count
 * that ring, so N slabs reproduce the np = 1 run. */
int fargo_oracle_accrete_sinkhole(fargo_oracle *o, double xp, double yp, double r_hill, double facc, double frac, double out3[3]);
int fargo_oracle_accrete_kley(fargo_oracle *o, double xp, double yp, double r_hill, double facc, double frac, double out3[3])
{
    const double OmegaF = o->bodies.omega_frame;
    const double density_floor = o->p.sigma_floor * o->p.sigma0;
    const double facc1 = 1.0 / 3.0 * facc, facc2 = 2.0 / 3.0 * facc;
    const double frac1 = frac, frac2 = 0.5 * frac;
    double dM = 0.0, dPx = 0.0, dPy = 0.0;
    for (int i = 0; i < o->nr; ++i) {
	for (int j = 0; j < o->ns; ++j) {
	    const size_t l = IDX(o, i, j);
	    const int jp = j == o->ns - 1 ? 0 : j + 1;
	    const double xc = o->rmed[i] * o->cosphi[j], yc = o->rmed[i] * o->sinphi[j];
	    const double dx = xp - xc, dy = yp - yc;
	    const double distance = sqrt(dx * dx + dy * dy);
	    if (!(distance < frac1 * r_hill))
		continue; /* zone 2 lies inside zone 1 */
	    const double vtcell = 0.5 * (o->vazi[l] + o->vazi[IDX(o, i, jp)]) + o->rmed[i] * OmegaF;
	    const double vrcell = 0.5 * (o->vrad[l] + o->vrad[IDX(o, i + 1, j)]);
	    const double vxcell = (vrcell * xc - vtcell * yc) / o->rmed[i];
	    const double vycell = (vrcell * yc + vtcell * xc) / o->rmed[i];
	    const double facc_max = 1 - density_floor / o->sigma[l];
	    const int active = (o->rank == 0 ? o->first_active < i : o->first_active <= i) && i < o->active_size;
	    {
		const double facc_ceil = facc_max < facc1 ? facc_max : facc1; /* std::min(facc1, facc_max) */
		const double deltaM = facc_ceil * o->sigma[l] * o->surf[i];
		o->sigma[l] *= 1.0 - facc_ceil;
		if (o->p.adiabatic)
		    o->energy[l] *= 1.0 - facc_ceil;
		if (active) {
		    dPx += deltaM * vxcell;
		    dPy += deltaM * vycell;
		    dM += deltaM;
		}
	    }
	    if (distance < frac2 * r_hill) {
		const double facc_ceil = facc_max < facc2 ? facc_max : facc2;
		const double deltaM = facc_ceil * o->sigma[l] * o->surf[i];
		o->sigma[l] *= 1.0 - facc_ceil;
		if (o->p.adiabatic)
		    o->energy[l] *= 1.0 - facc2;
		if (active) {
		    dPx += deltaM * vxcell;
		    dPy += deltaM * vycell;
		    dM += deltaM;
		}
	    }
	}
    }
    out3[0] = dM, out3[1] = dPx, out3[2] = dPy;
    return 0;
}

/* accretion::SinkHoleSinglePlanet (accretion.cpp:223-333): one zone of radius frac * RHill losing the fraction facc */
int fargo_oracle_accrete_sinkhole(fargo_oracle *o, double xp, double yp, double r_hill, double facc, double frac, double out3[3])
{
    const double OmegaF = o->bodies.omega_frame;
    const double density_floor = o->p.sigma_floor * o->p.sigma0;
    double dM = 0.0, dPx = 0.0, dPy = 0.0;
    for (int i = 0; i < o->nr; ++i) {
	for (int j = 0; j < o->ns; ++j) {
	    const size_t l = IDX(o, i, j);
	    const int jp = j == o->ns - 1 ? 0 : j + 1;
	    const double xc = o->rmed[i] * o->cosphi[j], yc = o->rmed[i] * o->sinphi[j];
	    const double dx = xp - xc, dy = yp - yc;
	    const double distance = sqrt(dx * dx + dy * dy);
	    if (!(distance < frac * r_hill))
		continue;
	    const double vtcell = 0.5 * (o->vazi[l] + o->vazi[IDX(o, i, jp)]) + o->rmed[i] * OmegaF;
	    const double vrcell = 0.5 * (o->vrad[l] + o->vrad[IDX(o, i + 1, j)]);
	    const double vxcell = (vrcell * xc - vtcell * yc) / o->rmed[i];
	    const double vycell = (vrcell * yc + vtcell * xc) / o->rmed[i];
	    const double facc_max = 1 - density_floor / o->sigma[l];
	    const double facc_ceil = facc_max < facc ? facc_max : facc; /* std::min(facc, facc_max) */
	    const double deltaM = facc_ceil * o->sigma[l] * o->surf[i];
	    o->sigma[l] *= 1.0 - facc_ceil;
	    if (o->p.adiabatic)
		o->energy[l] *= 1.0 - facc_ceil;
	    if ((o->rank == 0 ? o->first_active < i : o->first_active <= i) && i < o->active_size) {
		dPx += deltaM * vxcell;
		dPy += deltaM * vycell;
		dM += deltaM;
	    }
	}
    }
    out3[0] = dM, out3[1] = dPx, out3[2] = dPy;
    return 0;
}

/* accretion::AccreteOntoSinglePlanetViscous (accretion.cpp:335-480; "accretion method: viscous"): the fraction removed from a
 * cell is facc * nu * f(distance), f = 3 / (pi d_max^2) * (1 - distance / d_max), d_max = frac * RHill, with the VISCOSITY the
 * previous step stored.  facc = dt * 3 pi * accretion efficiency (the N-body side).  No CUDA counterpart yet: oracle only. */
int fargo_oracle_accrete_viscous(fargo_oracle *o, double xp, double yp, double r_hill, double facc, double frac, double out3[3])
{
    const double OmegaF = o->bodies.omega_frame;
    const double density_floor = o->p.sigma_floor * o->p.sigma0;
    const double dist_max = r_hill * frac;
    const double f_const = 3.0 / M_PI / pow(dist_max, 2);
    double dM = 0.0, dPx = 0.0, dPy = 0.0;
    for (int i = 0; i < o->nr; ++i) {
	for (int j = 0; j < o->ns; ++j) {
	    const size_t l = IDX(o, i, j);
	    const int jp = j == o->ns - 1 ? 0 : j + 1;
	    const double xc = o->rmed[i] * o->cosphi[j], yc = o->rmed[i] * o->sinphi[j];
	    const double dx = xp - xc, dy = yp - yc;
	    const double distance = sqrt(dx * dx + dy * dy);
	    if (!(distance < frac * r_hill))
		continue;
	    const double nu = o->viscosity[l];
	    const double spread = f_const * (1.0 - distance / dist_max);
	    const double vtcell = 0.5 * (o->vazi[l] + o->vazi[IDX(o, i, jp)]) + o->rmed[i] * OmegaF;
	    const double vrcell = 0.5 * (o->vrad[l] + o->vrad[IDX(o, i + 1, j)]);
	    const double vxcell = (vrcell * xc - vtcell * yc) / o->rmed[i];
	    const double vycell = (vrcell * yc + vtcell * xc) / o->rmed[i];
	    const double facc_max_dens = 1 - density_floor / o->sigma[l];
	    const double facc_tmp = facc * nu * spread;
	    const double facc_ceil = facc_max_dens < facc_tmp ? facc_max_dens : facc_tmp; /* std::min(facc_tmp, facc_max_dens) */
	    const double deltaM = facc_ceil * o->sigma[l] * o->surf[i];
	    o->sigma[l] *= 1.0 - facc_ceil;
	    if (o->p.adiabatic)
		o->energy[l] *= 1.0 - facc_ceil;
	    if ((o->rank == 0 ? o->first_active < i : o->first_active <= i) && i < o->active_size) {
		dPx += deltaM * vxcell;
		dPy += deltaM * vycell;
		dM += deltaM;
	    }
	}
    }
    out3[0] = dM, out3[1] = dPx, out3[2] = dPy;
    return 0;
}

/* Global disk quantities of monitor/Quantities.dat (output::write_quantities, output.cpp:326-520): serial sums in index
 * order, which is what the reference computes with OMP_NUM_THREADS=1 (its reductions have no defined order otherwise).
 * out8 = mass (quantities.cpp:51-78), angular momentum (:242-276), internal energy (:281-304), kinetic energy (:357-401),
 * radial kinetic (:406-438), azimuthal kinetic (:443-479), viscous dissipation (:306-328), luminosity (:330-352). */
int fargo_oracle_monitor_quantities(fargo_oracle *o, double radius_limit, double out8[8])
{
    const double OmegaF = o->bodies.omega_frame;
    const int ns = o->ns;
    double mass = 0.0, angmom = 0.0, eint = 0.0, ekin = 0.0, ekin_r = 0.0, ekin_a = 0.0, qp = 0.0, qm = 0.0;
    for (int i = o->first_active; i < o->active_size; ++i) {
	if (!(o->rmed[i] <= radius_limit))
	    continue;
	for (int j = 0; j < ns; ++j) {
	    const int jm = j == 0 ? ns - 1 : j - 1, jp = j == ns - 1 ? 0 : j + 1;
	    mass += o->surf[i] * o->sigma[IDX(o, i, j)];
	    angmom += o->surf[i] * 0.5 * (o->sigma[IDX(o, i, j)] + o->sigma[IDX(o, i, jm)]) * o->rmed[i] *
		      (o->vazi[IDX(o, i, j)] + OmegaF * o->rmed[i]);
	    if (o->p.adiabatic) {
		eint += o->surf[i] * o->energy[IDX(o, i, j)];
		qp += o->surf[i] * o->qplus[IDX(o, i, j)];
		qm += o->surf[i] * o->qminus[IDX(o, i, j)];
	    }
	    double v_radial_center =
		(o->rmed[i] - o->rinf[i]) * o->vrad[IDX(o, i + 1, j)] + (o->rsup[i] - o->rmed[i]) * o->vrad[IDX(o, i, j)];
	    v_radial_center /= (o->rsup[i] - o->rinf[i]);
	    const double v_azimuthal_center = 0.5 * (o->vazi[IDX(o, i, j)] + o->vazi[IDX(o, i, jp)]) + o->rmed[i] * OmegaF;
	    ekin += 0.5 * o->surf[i] * o->sigma[IDX(o, i, j)] *
		    (v_radial_center * v_radial_center + v_azimuthal_center * v_azimuthal_center);
	    ekin_r += 0.5 * o->surf[i] * o->sigma[IDX(o, i, j)] * (v_radial_center * v_radial_center);
	    ekin_a += 0.5 * o->surf[i] * o->sigma[IDX(o, i, j)] * (v_azimuthal_center * v_azimuthal_center);
	}
    }
    out8[0] = mass, out8[1] = angmom, out8[2] = eint, out8[3] = ekin, out8[4] = ekin_r, out8[5] = ekin_a, out8[6] = qp, out8[7] = qm;
    return 0;
}

/* The columns of monitor/Quantities.dat that are mass-weighted means or need the rings in order (output::write_quantities,
 * output.cpp:373-423): serial sums in index order, like fargo_oracle_monitor_quantities.
 * out5 = disk radius (quantities::gas_disk_radius, quantities.cpp:191-237: Rmed of the ring at which the running sum of the
 *        ring masses, ghost rings left out, first exceeds mass_fraction x the mass inside radius_limit),
 *        mass-weighted means of the cells' eccentricity vector rotated into the non-rotating frame (calculate_disk_ecc_vector
 *        :481-550, gas_reduce_mass_average :145-182; the caller forms sqrt(ex^2 + ey^2) and atan2(ey, ex), :552-567),
 *        mass-weighted mean aspect ratio SCALE_HEIGHT / Rb (compute_aspectratio mode 0 :784-806), the mass of those means,
 *        and the advection and viscous torques of the disk (gas_torques::calculate_advection_torque / calculate_viscous_torque,
 *        gas_torques.cpp:11-115, summed by gas_quantity_reduce quantities.cpp:80-105 in CalculateMonitorQuantitiesForOutput :1000-1018),
 *        the "potential energy" column -(mass-weighted mean of the POTENTIAL grid) (output.cpp:413-414) and the gravitational
 *        torque (gas_torques.cpp:122-153) — both from the POTENTIAL grid AS STORED, i.e. of the last kick's start (zeros before
 *        the first step); BodyForceFromPotential only.
 * One rank only (the reference gathers the ring masses on its root). */
int fargo_oracle_monitor_disk(fargo_oracle *o, double radius_limit, double mass_fraction, double frame_angle, double out9[9])
{
    if (o->nranks != 1)
	return 1;
    const int ns = o->ns;
    const double OmegaF = o->bodies.omega_frame;
    const double cms_mass = o->p.hydro_center_mass;
    const double sinF = sin(frame_angle), cosF = cos(frame_angle);
    double mass = 0.0, sum_ex = 0.0, sum_ey = 0.0, sum_h = 0.0, mass_lim = 0.0;
    for (int i = o->first_active; i < o->active_size; ++i) /* gas_total_mass (:51-78) */
	for (int j = 0; j < ns; ++j)
	    if (o->rmed[i] <= radius_limit)
		mass_lim += o->surf[i] * o->sigma[IDX(o, i, j)];
    /* the three means are separate loops in the reference, each with its own running mass; the masses are the same number */
    for (int i = o->first_active; i < o->active_size; ++i) {
	for (int j = 0; j < ns; ++j) {
	    if (!(o->rmed[i] <= radius_limit))
		continue;
	    const int jp = j == ns - 1 ? 0 : j + 1;
	    const size_t l = IDX(o, i, j);
	    const double cell_mass = o->sigma[l] * o->surf[i];
	    const double total_mass = cms_mass + o->sigma[l] * o->surf[i];
	    const double angle = (double)j * o->dphi;
	    const double r_x = o->rmed[i] * cos(angle), r_y = o->rmed[i] * sin(angle);
	    const double dist = sqrt(r_x * r_x + r_y * r_y);
	    const double v_xmed = cos(angle) * 0.5 * (o->vrad[l] + o->vrad[IDX(o, i + 1, j)]) -
				  sin(angle) * (0.5 * (o->vazi[l] + o->vazi[IDX(o, i, jp)]) + OmegaF * o->rmed[i]);
	    const double v_ymed = sin(angle) * 0.5 * (o->vrad[l] + o->vrad[IDX(o, i + 1, j)]) +
				  cos(angle) * (0.5 * (o->vazi[l] + o->vazi[IDX(o, i, jp)]) + OmegaF * o->rmed[i]);
	    const double jz = r_x * v_ymed - r_y * v_xmed;
	    const double e_x = jz * v_ymed / (o->p.G * total_mass) - r_x / dist;
	    const double e_y = -1.0 * jz * v_xmed / (o->p.G * total_mass) - r_y / dist;
	    const double e_x_frame = e_x * cosF - e_y * sinF;
	    const double e_y_frame = e_y * cosF + e_x * sinF;
	    mass += cell_mass;
	    sum_ex += e_x_frame * cell_mass;
	    sum_ey += e_y_frame * cell_mass;
	    sum_h += o->scale_height[l] / o->rmed[i] * cell_mass;
	}
    }
    double radius = 0.0, current = 0.0;
    /* the root walks the rings RootIMIN .. RootIMAX of every rank (split.cpp:339-343: the two ghost rings of the mesh are left
     * out) with a counter that starts at 1, so the ring that crosses the threshold reports its own GlobalRmed */
    for (int i = 1; i < o->nr - 1; ++i) {
	double ring = 0.0;
	for (int j = 0; j < ns; ++j)
	    ring += o->surf[i] * o->sigma[IDX(o, i, j)];
	current += ring;
	if (current > mass_fraction * mass_lim) {
	    radius = o->rmed[i];
	    break;
	}
    }
    double tadv = 0.0, tvisc = 0.0;
    for (int i = o->first_active; i < o->active_size; ++i) {
	if (!(o->rmed[i] <= radius_limit))
	    continue;
	const double r = o->rmed[i], inv_dr = o->invdiffrsup[i];
	for (int j = 0; j < ns; ++j) {
	    const int jp = j == ns - 1 ? 0 : j + 1;
	    const double sigma_cell = o->sigma[IDX(o, i, j)];
	    double vr_cell = (o->rmed[i] - o->rinf[i]) * o->vrad[IDX(o, i + 1, j)] + (o->rsup[i] - o->rmed[i]) * o->vrad[IDX(o, i, j)];
	    vr_cell *= inv_dr;
	    const double vazi_cell = 0.5 * (o->vazi[IDX(o, i, j)] + o->vazi[IDX(o, i, jp)]);
	    tadv += 0.0 + -pow(r, 2.0) * sigma_cell * vr_cell * vazi_cell * 1.0; /* worker array cleared, then += ... * dt (= 1) */
	}
    }
    for (int i = o->first_active; i < o->active_size; ++i) {
	if (!(o->rmed[i] <= radius_limit) || i < 1 || i >= o->nr - 1) /* calculate_viscous_torque fills rings 1 .. max_radial - 1 */
	    continue;
	const double r = o->rmed[i], inv_dr = o->invdiffrsup[i];
	const double inv_dr_med_top = o->invdiffrmed[i + 1], inv_dr_med_bot = o->invdiffrmed[i];
	for (int j = 0; j < ns; ++j) {
	    const int jp = j == ns - 1 ? 0 : j + 1, jm = j == 0 ? ns - 1 : j - 1;
	    const double sigma_cell = o->sigma[IDX(o, i, j)], viscosity_cell = o->viscosity[IDX(o, i, j)];
	    const double dvr_dphi_top = (o->vrad[IDX(o, i + 1, jp)] - o->vrad[IDX(o, i + 1, jm)]) * 0.5 * o->invdphi;
	    const double dvr_dphi_bot = (o->vrad[IDX(o, i, jp)] - o->vrad[IDX(o, i, jm)]) * 0.5 * o->invdphi;
	    double dvr_dphi = (o->rmed[i] - o->rinf[i]) * dvr_dphi_top + (o->rsup[i] - o->rmed[i]) * dvr_dphi_bot;
	    dvr_dphi *= inv_dr;
	    const double phi_dot_top = 0.5 * (o->vazi[IDX(o, i + 1, jp)] + o->vazi[IDX(o, i + 1, j)]) / o->rmed[i + 1];
	    const double phi_dot = 0.5 * (o->vazi[IDX(o, i, jp)] + o->vazi[IDX(o, i, j)]) / o->rmed[i];
	    const double phi_dot_bot = 0.5 * (o->vazi[IDX(o, i - 1, jp)] + o->vazi[IDX(o, i - 1, j)]) / o->rmed[i - 1];
	    const double dphi_dot_dr_top = (phi_dot_top - phi_dot) * inv_dr_med_top;
	    const double dphi_dot_dr_bot = (phi_dot - phi_dot_bot) * inv_dr_med_bot;
	    double dphi_dot_dr = (o->rmed[i] - o->rinf[i]) * dphi_dot_dr_top + (o->rsup[i] - o->rmed[i]) * dphi_dot_dr_bot;
	    dphi_dot_dr *= inv_dr;
	    tvisc += 0.0 + -pow(r, 3.0) * viscosity_cell * sigma_cell * (dphi_dot_dr + 1.0 / (pow(r, 2.0)) * dvr_dphi) * 1.0;
	}
    }
    double sum_pot = 0.0, mass_pot = 0.0, tgrav = 0.0;
    for (int i = o->first_active; i < o->active_size; ++i) {
	if (!(o->rmed[i] <= radius_limit))
	    continue;
	for (int j = 0; j < ns; ++j) {
	    const int jp = j == ns - 1 ? 0 : j + 1, jm = j == 0 ? ns - 1 : j - 1;
	    const size_t l = IDX(o, i, j);
	    const double cell_mass = o->sigma[l] * o->surf[i];
	    mass_pot += cell_mass;
	    sum_pot += o->potential[l] * cell_mass;
	    const double gradphi = (o->potential[IDX(o, i, jp)] - o->potential[IDX(o, i, jm)]) * o->invdphi * 0.5;
	    tgrav += 0.0 + -o->sigma[l] * gradphi * o->surf[i] * 1.0;
	}
    }
    out9[0] = radius;
    out9[1] = mass > 0.0 ? sum_ex / mass : 0.0;
    out9[2] = mass > 0.0 ? sum_ey / mass : 0.0;
    out9[3] = mass > 0.0 ? sum_h / mass : 0.0;
    out9[4] = mass;
    out9[5] = tadv;
    out9[6] = tvisc;
    const double nan = 0.0 / 0.0;
    out9[7] = o->p.body_force_from_potential ? -(mass_pot > 0.0 ? sum_pot / mass_pot : 0.0) : nan;
    out9[8] = o->p.body_force_from_potential ? tgrav : nan;
    return 0;
}

/* WriteMassFlow: the MASSFLOW grid VanLeerRadial accumulates (TransportEuler.cpp:610-616) */
int fargo_oracle_track_massflow(fargo_oracle *o, int on)
{
    if (on && !o->massflow)
	o->massflow = (double *)calloc((size_t)(o->nr + 1) * o->ns, sizeof(double));
    if (!on && o->massflow) {
	free(o->massflow);
	o->massflow = NULL;
    }
    return 0;
}
int fargo_oracle_clear_massflow(fargo_oracle *o)
{
    if (o->massflow)
	memset(o->massflow, 0, (size_t)(o->nr + 1) * o->ns * sizeof(double));
    return 0;
}

/* MassDelta's boundary flows (TransportEuler.cpp:578-608), summed in the reference's order */
int fargo_oracle_track_boundary_flow(fargo_oracle *o, int on)
{
    o->track_bflow = on != 0;
    return 0;
}
int fargo_oracle_boundary_flow(fargo_oracle *o, double out4[4], int reset)
{
    for (int q = 0; q < 4; ++q) {
	out4[q] = o->bflow[q];
	if (reset)
	    o->bflow[q] = 0.0;
    }
    return 0;
}

/* MassDelta's wave-damping terms (damping.cpp:335-357 and siblings), always kept */
int fargo_oracle_track_damping_mass(fargo_oracle *o, int on)
{
    (void)o, (void)on;
    return 0;
}
int fargo_oracle_damping_mass(fargo_oracle *o, double out4[4], int reset)
{
    for (int q = 0; q < 4; ++q) {
	out4[q] = o->dmass[q];
	if (reset)
	    o->dmass[q] = 0.0;
    }
    return 0;
}

/* the oracle keeps the POTENTIAL grid of every kick like the reference: nothing to arm */
int fargo_oracle_keep_potential(fargo_oracle *o, int on)
{
    (void)o, (void)on;
    return 0;
}

/* ComputeCircumPlanetaryMasses (circumplanetary_mass.cpp:11-51): mass of the active cells whose centre lies inside the body's
 * Roche radius (column "mdcp" of monitor/nbodyK.dat), serial sum in index order. */
int fargo_oracle_circumplanetary_mass(fargo_oracle *o, double x, double y, double roche_radius, double *out)
{
    double m = 0.0;
    for (int i = o->first_active; i < o->active_size; ++i)
	for (int j = 0; j < o->ns; ++j) {
	    const double cx = o->rmed[i] * o->cosphi[j], cy = o->rmed[i] * o->sinphi[j];
	    const double dist = sqrt((cx - x) * (cx - x) + (cy - y) * (cy - y));
	    if (dist < roche_radius)
		m += o->surf[i] * o->sigma[IDX(o, i, j)];
	}
    *out = m;
    return 0;
}

/* ComputeDiskOnPlanetAccel (Force.cpp:23-122), serial sum in index order (the reference's own order is undefined:
 * OpenMP reduction).  out4 = {axi, ayi, axo, ayo}. */
int fargo_oracle_disk_on_body_accel(fargo_oracle *o, int body, double klahr_factor, double out4[4])
{
    if (body < 0 || body >= o->bodies.n)
	return 1;
    const double x = o->bodies.x[body], y = o->bodies.y[body];
    const double a = sqrt(x * x + y * y);
    const double r_sm = o->bodies.cubic_smoothing_radius[body];
    double axi = 0.0, ayi = 0.0, axo = 0.0, ayo = 0.0;
    for (int n_rad = o->first_active; n_rad < o->active_size; ++n_rad) {
	double sigma1d = 0.0; /* ComputeAverageDensity (Pframeforce.cpp:174-188) */
	if (o->p.correct_disk_selfgravity) {
	    double sum = 0;
	    for (int n_az = 0; n_az < o->ns; ++n_az)
		sum += o->sigma[IDX(o, n_rad, n_az)];
	    sigma1d = sum / o->ns;
	}
	for (int n_az = 0; n_az < o->ns; ++n_az) {
	    const double smooth = o->p.thickness_smoothing * o->scale_height[IDX(o, n_rad, n_az)];
	    const double xc = o->rmed[n_rad] * o->cosphi[n_az];
	    const double yc = o->rmed[n_rad] * o->sinphi[n_az];
	    double cell_sigma = o->sigma[IDX(o, n_rad, n_az)];
	    if (o->p.correct_disk_selfgravity)
		cell_sigma -= sigma1d;
	    const double cellmass = o->surf[n_rad] * cell_sigma;
	    const double dx = xc - x;
	    const double dy = yc - y;
	    const double dist_2 = dx * dx + dy * dy;
	    const double dist_sm_2 = dist_2 + smooth * smooth;
	    const double dist_sm = sqrt(dist_sm_2);
	    const double dist_sm_3 = dist_sm_2 * dist_sm;
	    const double inv_dist_sm_3 = 1.0 / dist_sm_3;
	    double smooth_factor_klahr = 1.0;
	    if (klahr_factor > 0.0 && dist_sm < r_sm)
		smooth_factor_klahr = -(3.0 * pow(dist_sm / r_sm, 4.0) - 4.0 * pow(dist_sm / r_sm, 3.0));
	    if (o->rmed[n_rad] < a) {
		axi += o->p.G * cellmass * dx * inv_dist_sm_3 * smooth_factor_klahr;
		ayi += o->p.G * cellmass * dy * inv_dist_sm_3 * smooth_factor_klahr;
	    } else {
		axo += o->p.G * cellmass * dx * inv_dist_sm_3 * smooth_factor_klahr;
		ayo += o->p.G * cellmass * dy * inv_dist_sm_3 * smooth_factor_klahr;
	    }
	}
    }
    out4[0] = axi, out4[1] = ayi, out4[2] = axo, out4[3] = ayo;
    return 0;
}

/* kick / drift / finish_step: the pieces step_Euler and step_LeapFrog (simulation.cpp:148-267, 276-459) arrange */
int fargo_oracle_kick(fargo_oracle *o, double dt)
{
    /* step_LeapFrog recomputes the pressure before its SECOND kick (simulation.cpp:381) but NOT c_s / H: the potential
     * smoothing of that kick uses the scale height recalculate_viscosity left during the first kick.  The first kick (like
     * step_Euler's) reads the PRESSURE the previous step stored — which matters when AccreteOntoPlanets has changed Sigma / e in
     * between (simulation.cpp:302-303): the stored pressure is the pre-accretion one. */
    const int second = o->kicks_this_step > 0;
    o->kicks_this_step++;
    fargo_oracle_stage_potential(o);
    if (second) {
	if (o->p.pvte) { /* simulation.cpp:368-376: after the potential (which still read the stored scale height), c_s and H, the
			  * lookup, and c_s and H again */
	    compute_sound_speed(o);
	    compute_scale_height(o);
	    compute_gamma_mu(o);
	    compute_sound_speed(o);
	    compute_scale_height(o);
	}
	compute_pressure(o); /* :381 */
    }
    fargo_oracle_stage_sources(o, dt);
    fargo_oracle_stage_artvisc(o, dt);
    fargo_oracle_stage_viscosity(o, dt);
    if (o->p.adiabatic)
	fargo_oracle_stage_substep3(o, dt);
    return 0;
}
int fargo_oracle_drift(fargo_oracle *o, double dt)
{
    fargo_oracle_stage_boundary(o, 0.0, 0);
    fargo_oracle_stage_transport(o, dt);
    return 0;
}
int fargo_oracle_finish_step(fargo_oracle *o, double dt)
{
    o->kicks_this_step = 0;
    fargo_oracle_stage_boundary(o, dt, 1);
    fargo_oracle_stage_derived(o);
    return 0;
}
