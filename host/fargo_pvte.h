/* fargo_pvte.h — lookup tables of the PVTE equation of state (variable adiabatic index: hydrogen dissociation and
 * ionisation, ortho/para rotational + vibrational modes of H2), as pvte::initializeLookupTables builds them
 * (pvte_law.cpp:65-395; after Vaidya et al. 2015 and D'Angelo et al. 2013 / PLUTO's zeta tables).
 *
 * Plain C99 so that the C++ host driver (product) and the C oracle (test infrastructure, which may use product code — never
 * the other way round) build the SAME tables.  The tables must be the reference's bit for bit: gamma_eff, mu and Gamma_1 of a
 * cell are bilinear interpolations in them, so every operation below keeps the reference's order and calls the same libm
 * functions (exp, log, log10, pow, sqrt); the oracle's PVTE fixture pins the result through the fields the reference wrote.
 *
 * Grid: 1000 x 1000 points, log-spaced in density (1e-23 .. 1 g/cm^3) and specific energy (1e8 .. 1e15 erg/g). */
#ifndef FARGO_PVTE_H
#define FARGO_PVTE_H
#include <math.h>
#include <stdlib.h>

#include "../include/fargo_b200.h" /* fargo_pvte_consts */

#define FARGO_PVTE_NI 1000
#define FARGO_PVTE_NJ 1000
#define FARGO_PVTE_NZETA 5000
#define FARGO_PVTE_RHOMIN 1.0e-23
#define FARGO_PVTE_RHOMAX 1.0
#define FARGO_PVTE_EMIN 1.0e8
#define FARGO_PVTE_EMAX 1.0e15


typedef struct fargo_pvte_tables {
    double *rho, *e;          /* NI, NJ grid values */
    double *mu, *geff, *g1;   /* NI * NJ, index j + i * NJ */
    double dlogrho, dloge;
    double lnT[FARGO_PVTE_NZETA], funcdum[FARGO_PVTE_NZETA]; /* zeta tables */
    fargo_pvte_consts k;
} fargo_pvte_tables;

/* Saha equations: ionisation (x) and dissociation (y) fractions of hydrogen (pvte_law.cpp:443-495) */
static double fargo_pvte_saha(const fargo_pvte_consts *k, const int dissociation, const double rho, const double T)
{
    const double pi = 3.14159265358979323846;
    const double h_bar = k->h / (2.0 * pi);
    double rhs_exponent, rhs_constant;
    if (dissociation) {
	rhs_exponent = -4.48 * k->eV / k->k_B;
	rhs_constant = k->m_H / (2.0 * k->xMF) * pow(k->m_H * k->k_B / (4 * pi * h_bar * h_bar), 1.5);
    } else {
	rhs_exponent = -13.60 * k->eV / k->k_B;
	rhs_constant = k->m_H / k->xMF * pow(k->m_e * k->k_B / (2 * pi * h_bar * h_bar), 1.5);
    }
    double frac = 1.0;
    const double A = rhs_constant * pow(T, 1.5) * exp(rhs_exponent / T) / rho;
    if (A < 1.0e8)
	frac = 0.5 * (-A + sqrt(A * A + 4.0 * A));
    return frac;
}
static double fargo_pvte_mu_of(const fargo_pvte_consts *k, const double x, const double y)
{
    return 4.0 / (2.0 * k->xMF * (1.0 + y + 2.0 * y * x) + 1.0 - k->xMF);
}

/* rotational + vibrational energy of H2 from the zeta table (get_funcDum, pvte_law.cpp:77-100) */
static double fargo_pvte_funcdum(const fargo_pvte_tables *t, const double T)
{
    const double y = log(T);
    const int N = FARGO_PVTE_NZETA;
    if (y > t->lnT[N - 2])
	return t->funcdum[N - 2];
    if (y < t->lnT[0])
	return t->funcdum[0];
    const double dy = t->lnT[1] - t->lnT[0];
    const int indx = (int)(floor((y - t->lnT[0]) / dy));
    return (t->funcdum[indx] * (t->lnT[indx + 1] - y) + t->funcdum[indx + 1] * (y - t->lnT[indx])) / dy;
}

/* internal energy of the gas in units of k_B T per ... (gasEnergyContributions, pvte_law.cpp:103-130): molecular hydrogen,
 * ionisation, dissociation, helium, atomic hydrogen — summed in that order */
static double fargo_pvte_energy_terms(const fargo_pvte_tables *t, const double x, const double y, const double T)
{
    const fargo_pvte_consts *k = &t->k;
    const double epsHI = 1.5 * k->xMF * (1.0 + x) * y;
    const double epsHe = 0.375 * (1.0 - k->xMF);
    const double epsHH = 4.48 * k->eV * k->xMF * y / (2.0 * k->k_B * T);
    const double epsHII = 13.60 * k->eV * k->xMF * x * y / (k->k_B * T);
    const double epsH2 = 0.5 * k->xMF * (1.0 - y) * fargo_pvte_funcdum(t, T);
    return epsH2 + epsHII + epsHH + epsHe + epsHI;
}

static double fargo_pvte_gamma_eff(const fargo_pvte_tables *t, const double T, const double rho)
{
    const double x = fargo_pvte_saha(&t->k, 0, rho, T);
    const double y = fargo_pvte_saha(&t->k, 1, rho, T);
    const double mu = fargo_pvte_mu_of(&t->k, x, y);
    return 1.0 + 1.0 / (mu * fargo_pvte_energy_terms(t, x, y, T));
}

/* first adiabatic index from centred differences in T and rho (gamma1, pvte_law.cpp:151-212) */
static double fargo_pvte_gamma1(const fargo_pvte_tables *t, const double T, const double rho)
{
    const fargo_pvte_consts *k = &t->k;
    const double epsilon = 1.0e-4;
    const double TL = T * (1.0 - epsilon), TR = T * (1.0 + epsilon), dT = TL - TR;
    double xL = fargo_pvte_saha(k, 0, rho, TL), xR = fargo_pvte_saha(k, 0, rho, TR);
    const double xc = fargo_pvte_saha(k, 0, rho, T);
    double yL = fargo_pvte_saha(k, 1, rho, TL), yR = fargo_pvte_saha(k, 1, rho, TR);
    const double yc = fargo_pvte_saha(k, 1, rho, T);
    const double eps = fargo_pvte_energy_terms(t, xc, yc, T);
    const double eL = (fargo_pvte_energy_terms(t, xL, yL, TL)) * TL;
    const double eR = (fargo_pvte_energy_terms(t, xR, yR, TR)) * TR;
    const double e = eps * T;
    const double cv = (eL - eR) / dT;
    double muL = fargo_pvte_mu_of(k, xL, yL), muR = fargo_pvte_mu_of(k, xR, yR), muc = fargo_pvte_mu_of(k, xc, yc);
    const double gamma_eff = 1.0 + 1.0 / (muc * eps);
    const double p = (gamma_eff - 1.0) * e;
    const double chiT = 1.0 - T / muc * (muL - muR) / dT;
    const double rhoL = rho * (1.0 - epsilon), rhoR = rho * (1.0 + epsilon), drho = rhoL - rhoR;
    xL = fargo_pvte_saha(k, 0, rhoL, T), xR = fargo_pvte_saha(k, 0, rhoR, T);
    yL = fargo_pvte_saha(k, 1, rhoL, T), yR = fargo_pvte_saha(k, 1, rhoR, T);
    muL = fargo_pvte_mu_of(k, xL, yL), muR = fargo_pvte_mu_of(k, xR, yR), muc = fargo_pvte_mu_of(k, xc, yc);
    const double chiRho = 1.0 - rho / muc * (muL - muR) / drho;
    return p * pow(chiT, 2) / (cv * T) + chiRho;
}

/* T(e, rho): root of mu e (gamma - 1) / R - T (gamma_mu_root + energy_to_temperature, pvte_law.cpp:215-300).  The
 * reference's bracketing iteration is kept step for step: it never refreshes f(a), f(b), f(c) after the first evaluation, so
 * it is a bisection-like search that stops when |b - a| <= 1e-3 K — the table values depend on exactly where it stops. */
static double fargo_pvte_root_fn(const fargo_pvte_tables *t, const double T, const double rho, const double energy)
{
    const fargo_pvte_consts *k = &t->k;
    const double x = fargo_pvte_saha(k, 0, rho, T);
    const double y = fargo_pvte_saha(k, 1, rho, T);
    const double mu = fargo_pvte_mu_of(k, x, y);
    const double gamma = 1.0 + 1.0 / (mu * fargo_pvte_energy_terms(t, x, y, T));
    const double R = k->k_B / k->mp;
    const double temperature = mu * energy * (gamma - 1.0) / R;
    return temperature - T;
}
static double fargo_pvte_temperature(const fargo_pvte_tables *t, const double energy, const double rho)
{
    const double delta = 1.0e-3;
    double a = 1.0e0, b = 1.0e7, c, d = 0.0, s, tmp;
    double fa = fargo_pvte_root_fn(t, a, rho, energy);
    double fb = fargo_pvte_root_fn(t, b, rho, energy);
    double fs;
    volatile double fc;
    if (fabs(fa) < fabs(fb)) {
	tmp = a, a = b, b = tmp;
	tmp = fa, fa = fb, fb = tmp;
    }
    c = a;
    fc = fa;
    int mflag = 1;
    while (fabs(b - a) > delta) {
	if ((fa != fc) && (fb != fc))
	    s = a * fb * fc / ((fa - fb) * (fa - fc)) + b * fa * fc / ((fb - fa) * (fb - fc)) + c * fa * fb / ((fc - fa) * (fc - fb));
	else
	    s = b - fb * (b - a) / (fb - fa);
	const double q = (3.0 * a + b) / 4.0;
	const double lo = (b < q) ? b : q, hi = (q < b) ? b : q; /* std::min(q, b), std::max(q, b) */
	if (((s < lo) && (s > hi)) || (mflag && (fabs(s - b) >= fabs(b - c) / 2.0)) || (!mflag && (fabs(s - b) >= fabs(c - d) / 2.0)) ||
	    (mflag && (fabs(b - c) < delta)) || (!mflag && (fabs(c - d) < delta))) {
	    s = (a + b) / 2.0;
	    mflag = 1;
	} else {
	    mflag = 0;
	}
	fs = fargo_pvte_root_fn(t, s, rho, energy);
	d = c;
	c = b;
	if (fa * fs < 0.0)
	    b = s;
	else
	    a = s;
	if (fabs(fa) < fabs(fb)) {
	    tmp = a, a = b, b = tmp;
	    tmp = fa, fa = fb, fb = tmp;
	}
    }
    return b;
}

/* makeZetaTables (pvte_law.cpp:305-365), ORTHO_PARA_MODE 1: equilibrium mixture (alpha = 1, beta = 0, gamma = 1) */
static void fargo_pvte_zeta(fargo_pvte_tables *t)
{
    const double THETA_V = 6140.0, THETA_R = 85.5, Temp0 = 1.0, Tmax = 1.0e12;
    const double alpha = 1.0, beta = 0.0, gamma = 1.0;
    const double dy = log(Tmax / Temp0) * (1. / (double)FARGO_PVTE_NZETA);
    const double b1 = 2.0 * THETA_R;
    int j;
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (j = 0; j < FARGO_PVTE_NZETA; j++) {
	const double T = Temp0 * exp(j * dy);
	const double inv_T2 = 1.0 / (T * T);
	double zetaP = 0.0, dzetaP = 0.0, sum1 = 0.0, sum2 = 0.0;
	unsigned int i;
	for (i = 0; i <= 10000; i++) {
	    const double a = 2 * i + 1;
	    const double b = i * (i + 1) * THETA_R;
	    if ((i % 2) == 0) {
		const double scrh = a * exp(-b / T);
		zetaP += scrh;
		dzetaP += scrh * b;
	    } else {
		const double db = b - b1;
		const double scrh = a * exp(-db / T);
		sum1 += scrh;
		sum2 += scrh * db;
	    }
	}
	dzetaP *= inv_T2;
	const double zetaO = exp(-b1 / T) * sum1;
	const double dzetaO = exp(-b1 / T) * (b1 * sum1 + sum2) * inv_T2;
	const double dzO_zO_m = sum2 / sum1 * inv_T2;
	t->lnT[j] = log(T);
	const double scrh = zetaO * exp(2.0 * THETA_R / T);
	const double zetaR = pow(zetaP, alpha) * pow(scrh, beta) + 3.0 * gamma * zetaO;
	const double dzetaR = (zetaR - 3.0 * gamma * zetaO) * (alpha * (dzetaP / zetaP) + beta * dzO_zO_m) + 3.0 * gamma * dzetaO;
	const double dum1 = THETA_V / T;
	const double dum2 = dum1 * exp(-dum1) / (1.0 - exp(-dum1));
	const double dum3 = (T / zetaR) * dzetaR;
	t->funcdum[j] = 1.5 + dum2 + dum3;
    }
}

/* pvte::initializeLookupTables (pvte_law.cpp:371-394).  Returns NULL when out of memory; free with fargo_pvte_free. */
static fargo_pvte_tables *fargo_pvte_build(const fargo_pvte_consts *k)
{
    const int Ni = FARGO_PVTE_NI, Nj = FARGO_PVTE_NJ;
    fargo_pvte_tables *t = (fargo_pvte_tables *)calloc(1, sizeof(fargo_pvte_tables));
    if (!t)
	return NULL;
    t->k = *k;
    t->rho = (double *)malloc(sizeof(double) * Ni);
    t->e = (double *)malloc(sizeof(double) * Nj);
    t->mu = (double *)malloc(sizeof(double) * Ni * Nj);
    t->geff = (double *)malloc(sizeof(double) * Ni * Nj);
    t->g1 = (double *)malloc(sizeof(double) * Ni * Nj);
    if (!t->rho || !t->e || !t->mu || !t->geff || !t->g1)
	return NULL;
    t->dlogrho = log10(FARGO_PVTE_RHOMAX / FARGO_PVTE_RHOMIN) / (double)Ni;
    t->dloge = log10(FARGO_PVTE_EMAX / FARGO_PVTE_EMIN) / (double)Nj;
    fargo_pvte_zeta(t);
    int i;
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (i = 0; i < Ni; ++i) {
	int j;
	for (j = 0; j < Nj; ++j) {
	    const double rhoi = pow(10.0, (t->dlogrho * i)) * FARGO_PVTE_RHOMIN;
	    const double ej = pow(10.0, (t->dloge * j)) * FARGO_PVTE_EMIN;
	    const double T = fargo_pvte_temperature(t, ej, rhoi);
	    const double x = fargo_pvte_saha(&t->k, 0, rhoi, T), y = fargo_pvte_saha(&t->k, 1, rhoi, T);
	    const int index = j + i * Nj;
	    t->rho[i] = rhoi;
	    if (i == 0)
		t->e[j] = ej;
	    t->mu[index] = fargo_pvte_mu_of(&t->k, x, y);
	    t->geff[index] = fargo_pvte_gamma_eff(t, T, rhoi);
	    t->g1[index] = fargo_pvte_gamma1(t, T, rhoi);
	}
    }
    return t;
}
static void fargo_pvte_free(fargo_pvte_tables *t)
{
    if (!t)
	return;
    free(t->rho), free(t->e), free(t->mu), free(t->geff), free(t->g1);
    free(t);
}

/* pvte lookup (pvte_law.cpp:396-441): bilinear interpolation in (rho, e) [cgs]; out = gamma_eff, mu, Gamma_1 */
static void fargo_pvte_lookup(const fargo_pvte_tables *t, const double rho, const double e, double *geff, double *mu, double *g1)
{
    const int Ni = FARGO_PVTE_NI, Nj = FARGO_PVTE_NJ;
    int i = (int)(floor(log10(rho / FARGO_PVTE_RHOMIN) / t->dlogrho));
    int j = (int)(floor(log10(e / FARGO_PVTE_EMIN) / t->dloge));
    if (i >= Ni - 1)
	i = Ni - 2;
    if (i < 0)
	i = 0;
    if (j >= Nj - 1)
	j = Nj - 2;
    if (j < 0)
	j = 0;
    const double x = (rho - t->rho[i]) / (t->rho[i + 1] - t->rho[i]);
    const double y = (e - t->e[j]) / (t->e[j + 1] - t->e[j]);
    const int a = j + (i + 1) * Nj, b = j + i * Nj, c = j + 1 + (i + 1) * Nj, d = j + 1 + i * Nj;
    const double *tab[3] = {t->geff, t->mu, t->g1};
    double *out[3] = {geff, mu, g1};
    int q;
    for (q = 0; q < 3; ++q) {
	const double S_ij = tab[q][a] * x + tab[q][b] * (1.0 - x);
	const double S_ijp1 = tab[q][c] * x + tab[q][d] * (1.0 - x);
	*out[q] = S_ij * (1.0 - y) + S_ijp1 * y;
    }
}
#endif
