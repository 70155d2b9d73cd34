/* GSL shim — TEST INFRASTRUCTURE ONLY. The reference calls gsl_sf_bessel_Inu only for the
 * spreading-ring initial condition (init.cpp:381,398); libstdc++'s cyl_bessel_i is the same function. */
#ifndef ORACLE_SHIM_GSL_SF_BESSEL_H
#define ORACLE_SHIM_GSL_SF_BESSEL_H
#include <cmath>
static inline double gsl_sf_bessel_Inu(double nu, double x) { return std::cyl_bessel_i(nu, x); }
#endif
