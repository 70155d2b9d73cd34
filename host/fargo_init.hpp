// fargo_init.hpp — what the reference does between reading a setup YAML and the first hydro step, for `fargocpt_b200 start`:
//   units::set_baseunits / calculate_unit_factors          units.cpp:131-185, 270-377    -> UnitSystem
//   constants::initialize_constants / ..._in_code_units    constants.cpp:178-262         -> UnitSystem
//   write_code_units_file / write_code_constants_file      units.cpp:451-501, constants.cpp:332-362
//   init_radialarrays                                      init.cpp:78-150               -> make_radii
//   t_planetary_system::init_system / init_planet / initialize_planet_jacobi(_adjust_first_two) / move_to_hydro_frame_center /
//   calculate_orbital_elements / compute_dist_to_primary / init_roche_radii
//                                                          nbody/planetary_system.cpp:68-260, 483-578, 750-805, 941-1004
//   init_gas_density / init_gas_energy / init_gas_velocities (power-law profile branch)
//                                                          init.cpp:937-960, 1257-1300, 1717-1771, Theo.cpp:86-201,
//                                                          viscosity/viscous_radial_speed.cpp:19-200
// The arithmetic follows the reference expression by expression (same libm, -ffp-contract=off), so the initial state and
// the code-unit constants come out bit-identical to the reference's IEEE build for the setups this driver supports:
// power-law disks (`SigmaCondition` / `EnergyCondition`: profile), star-only or star + planets, HydroFrameCenter: primary.
// Anything else is refused by name instead of being silently ignored.
#pragma once
#include <cmath>
#include <cstdio>
#include <fstream>
#include <limits>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace finit
{

[[noreturn]] inline void refuse(const std::string &what)
{
    fprintf(stderr, "fargocpt_b200: %s\n", what.c_str());
    exit(1);
}

// ---------------------------------------------------------------------------------------------
// The units of measurement FargoCPT setups use, as SI multipliers (units.cpp:111-129 for the astronomical ones; llnl/units
// for the rest).  dim: 'L' length, 'M' mass, 'T' time, 'K' temperature, 'S' surface density.
struct UnitDef {
    double si;
    char dim;
};
inline const std::map<std::string, UnitDef> &unit_table()
{
    static const std::map<std::string, UnitDef> t = {
	{"au", {1.495978707e11, 'L'}},	     {"solRadius", {6.95700e8, 'L'}},	 {"jupiterRadius", {69911000, 'L'}},
	{"earthRadius", {6371000, 'L'}},     {"m", {1.0, 'L'}},			 {"cm", {0.01, 'L'}},
	{"km", {1000.0, 'L'}},		     {"solMass", {1.98847e30, 'M'}},	 {"jupiterMass", {1.8982e27, 'M'}},
	{"earthMass", {5.97217e24, 'M'}},    {"kg", {1.0, 'M'}},		 {"g", {0.001, 'M'}},
	{"s", {1.0, 'T'}},		     {"K", {1.0, 'K'}},			 {"g/cm2", {0.001 / (0.01 * 0.01), 'S'}},
	{"g/cm^2", {0.001 / (0.01 * 0.01), 'S'}},
	// kinematic viscosity (dim 'V' = L^2 / T) and frequency (dim 'F' = 1 / T)
	{"cm2/s", {0.01 * 0.01, 'V'}},	     {"cm^2/s", {0.01 * 0.01, 'V'}},	 {"m2/s", {1.0, 'V'}},
	{"m^2/s", {1.0, 'V'}},		     {"1/s", {1.0, 'F'}},		 {"s^-1", {1.0, 'F'}},
	{"Hz", {1.0, 'F'}},
    };
    return t;
}
// "<number> [unit]" -> (number, unit name or "")
inline bool split_value(const std::string &v, double &x, std::string &unit)
{
    std::istringstream is(v);
    unit.clear();
    if (!(is >> x))
	return false;
    is >> unit;
    return true;
}

struct UnitSystem {
    // base units as SI multipliers (units::L0, M0, T0, Temp0 are llnl precise_units; only their multipliers matter)
    double L0 = 0.01, M0 = 0.001, T0 = 1.0, Temp0 = 1.0;
    // code -> cgs factors (units.cpp:270-377), in the order of write_code_units_file
    double length, mass, time, temperature, energy, energy_density, density, surface_density, opacity, energy_flux, velocity,
	angular_momentum, kinematic_viscosity, dynamic_viscosity, acceleration, stress, pressure, power, potential, torque, force,
	mass_accretion_rate;
    struct Constant {
	const char *name, *symbol, *cgs_unit;
	double code, cgs;
    };
    Constant G, k_B, m_u, h, c, R, sigma, m_H, m_e, eV;

    // units::set_baseunits (units.cpp:131-185) with the default t0 / temp0 (derived from G, k_B, m_u)
    void set_baseunits(const std::string &l0s, const std::string &m0s)
    {
	double lv, mv;
	std::string lu, mu_;
	if (!split_value(l0s, lv, lu) || !split_value(m0s, mv, mu_))
	    refuse("l0 / m0 are not numbers: " + l0s + ", " + m0s);
	if (lu.empty() != mu_.empty())
	    refuse("l0 and m0 need to either all have a unit or all have no unit");
	if (lu.empty()) { // "Physical units implicitly applied": au and solMass
	    lu = "au", mu_ = "solMass";
	}
	const auto &t = unit_table();
	if (!t.count(lu) || t.at(lu).dim != 'L')
	    refuse("l0: unknown unit of length '" + lu + "'");
	if (!t.count(mu_) || t.at(mu_).dim != 'M')
	    refuse("m0: unknown unit of mass '" + mu_ + "'");
	L0 = lv * t.at(lu).si;	// measurement(value, unit).as_unit(): value * multiplier
	M0 = mv * t.at(mu_).si;
	// llnl constants (units/units.hpp:2033-2063), SI
	const double G_si = 6.67430e-11, k_si = 1.380649e-23, mu_si = 1.66053906660e-27;
	// T0 = sqrt((1 * L0 * L0 * L0) / (1 * M0 * G)).as_unit(): the measurement's value and the unit's multiplier take their
	// square roots separately (units/units.cpp:92-107)
	const double unit_mult = ((L0 * L0) * L0) / (M0 * 1.0);
	const double value = 1.0 / (1.0 * G_si);
	T0 = std::sqrt(value) * std::sqrt(unit_mult);
	// Temp0 = (1 * G * mu / kB * M0 / L0).as_unit()
	Temp0 = ((1.0 * G_si) * mu_si / k_si) * (M0 / L0);
    }

    // units::calculate_unit_factors (units.cpp:270-377) + constants::initialize_constants / calculate_constants_in_code_units
    void calculate()
    {
	length = L0 / 0.01;
	mass = M0 / 0.001;
	time = T0 / 1.0;
	energy = length * length * mass / (time * time);
	energy_density = mass / (time * time);
	temperature = Temp0 / 1.0;
	density = mass / (length * length * length);
	surface_density = mass / (length * length);
	opacity = length * length / mass;
	energy_flux = energy / (length * length * time);
	velocity = length / time;
	acceleration = length / (time * time);
	angular_momentum = length * mass * velocity;
	kinematic_viscosity = length * length / time;
	dynamic_viscosity = mass / (length * time);
	stress = mass / (time * time);
	pressure = mass / (time * time);
	power = mass * length * length / (time * time * time);
	potential = length * length / (time * time);
	torque = length * length * mass / (time * time);
	force = mass * length / (time * time);
	mass_accretion_rate = mass / time;
	// constants.cpp:48-85: value_as(cgs unit) = value * 1 / multiplier(cgs unit), the unit products left to right
	const double cm = 0.01, g = 0.001, s = 1.0, K = 1.0;
	const double cgs_G = 6.67430e-11 * 1.0 / (cm * cm * cm / (g * s * s));
	const double cgs_k_B = 1.380649e-23 * 1.0 / (g * cm * cm / (K * s * s));
	const double cgs_m_u = 1.66053906660e-27 * 1.0 / g;
	const double cgs_h = 6.62607015e-34 * 1.0 / (g * cm * cm / s);
	const double cgs_c = 299792458.0 * 1.0 / (cm / s);
	const double cgs_m_e = 9.1093837015e-31 * 1.0 / g;
	const double cgs_eV = 1.0e7 * 1.602176634e-19;
	const double cgs_m_H = 1.007825 * cgs_m_u;
	G = {"gravitational constant", "G", "cm^3 g^-1 s^-2", 1.0, cgs_G};
	k_B = {"Boltzmann constant", "k_B", "erg K^-1", 1.0, cgs_k_B};
	m_u = {"molecular mass", "m_u", "g", 1.0, cgs_m_u};
	h = {"Planck constant", "h", "erg s", 1.0, cgs_h};
	c = {"speed of light", "c", "cm s^-1", 1.0, cgs_c};
	R = {"specific gas constant", "R", "erg K^-1 g^-1", 1.0, cgs_k_B / cgs_m_u};
	eV = {"electron volt", "eV", "erg", 1.0, cgs_eV};
	m_e = {"electron mass", "m_e", "g", 1.0, cgs_m_e};
	m_H = {"hydrogen atom mass", "m_H", "g", 1.0, cgs_m_H};
	sigma = {"Stefan-Boltzmann constant", "sigma", "erg cm^-2 s^-1 K^-4", 1.0,
		 2. * pow(M_PI, 5) * pow(cgs_k_B, 4) / (15. * pow(cgs_h, 3) * pow(cgs_c, 2))};
	G.code = G.cgs / (length * length * length / (mass * time * time));
	k_B.code = k_B.cgs / (energy / temperature);
	m_u.code = m_u.cgs / (mass);
	h.code = h.cgs / (energy * time);
	c.code = c.cgs / (length / time);
	m_e.code = m_e.cgs / (mass);
	m_H.code = m_H.cgs / (mass);
	eV.code = eV.cgs / (energy);
	R.code = R.cgs / (energy / (temperature * mass));
	sigma.code = sigma.cgs / (energy / (length * length * time * temperature * temperature * temperature * temperature));
    }

    // config::Config::get<double>(key, default, unit) (config.cpp:333-383): a value with a unit is converted to the code
    // unit of dimension `dim` (value * multiplier / code multiplier), a bare number is taken as it is
    double in_code_units(const std::string &v, char dim) const
    {
	double x;
	std::string u;
	if (!split_value(v, x, u))
	    refuse("not a number: " + v);
	if (u.empty())
	    return x;
	const auto &t = unit_table();
	if (!t.count(u))
	    refuse("unit '" + u + "' is not known to this driver (value '" + v + "')");
	if (t.at(u).dim != dim)
	    refuse("unit '" + u + "' has the wrong dimension in '" + v + "'");
	const double target = dim == 'L' ? L0 : dim == 'M' ? M0 : dim == 'T' ? T0 : dim == 'K' ? Temp0 : dim == 'V' ? L0 * L0 / T0 :
			      dim == 'F' ? 1.0 / T0 : M0 / (L0 * L0);
	return x * t.at(u).si / target;
    }

    static std::string num17(double x)
    { // std::ostream with precision(max_digits10)
	char b[64];
	snprintf(b, sizeof b, "%.17g", x);
	return b;
    }
    void write_files(const std::string &outdir) const
    {
	{
	    std::ofstream of(outdir + "/units.yml");
	    of << "# code units file\n# version 0.2\n\n";
	    const std::pair<const char *, std::pair<double, const char *>> u[] = {
		{"length", {length, "cm"}},
		{"mass", {mass, "g"}},
		{"time", {time, "s"}},
		{"temperature", {temperature, "K"}},
		{"energy", {energy, "erg"}},
		{"energy surface density", {energy_density, "erg cm^-2"}},
		{"density", {density, "g cm^-3"}},
		{"mass surface density", {surface_density, "g cm^-2"}},
		{"opacity", {opacity, "g^-1 cm^2"}},
		{"energy flux", {energy_flux, "erg cm^-2 s^-1"}},
		{"velocity", {velocity, "cm s^-1"}},
		{"angular momentum", {angular_momentum, "cm^2 g s^-1"}},
		{"kinematic viscosity", {kinematic_viscosity, "cm^2 s^-1"}},
		{"dynamic viscosity", {dynamic_viscosity, "P"}},
		{"acceleration", {acceleration, "cm s^-2"}},
		{"stress", {stress, "g s^-2"}},
		{"pressure", {pressure, "dyn cm^-1"}},
		{"power", {power, "erg/s"}},
		{"potential", {potential, "erg/g"}},
		{"torque", {torque, "erg"}},
		{"force", {force, "dyn"}},
		{"mass accretion rate", {mass_accretion_rate, "g s^-1"}},
	    };
	    for (auto &e : u) {
		of << e.first << ":\n  cgs symbol: " << e.second.second << "\n  cgs value: " << num17(e.second.first)
		   << "\n  unit: " << num17(e.second.first) << " " << e.second.second << "\n\n";
	    }
	}
	{
	    std::ofstream of(outdir + "/constants.yml");
	    of << "# log output of physical constants file\n# version 0.1\n\n";
	    for (const Constant *k : {&G, &k_B, &m_u, &h, &c, &R, &sigma, &m_H, &m_e, &eV})
		of << k->name << ":\n  symbol: " << k->symbol << "\n  code value: " << num17(k->code) << "\n  cgs value: " << num17(k->cgs)
		   << "\n  cgs unit symbol: " << k->cgs_unit << "\n\n";
	}
    }
};

// ---------------------------------------------------------------------------------------------
// init_radialarrays (init.cpp:78-150): Radii[0 .. Nrad], one ghost cell on either side of [Rmin, Rmax]
inline std::vector<double> make_radii(int spacing /* fargo_params::radial_spacing */, int nrad, double rmin, double rmax,
				      double exp_cell_size_factor)
{
    std::vector<double> r((size_t)nrad + 1);
    if (spacing == 0 /* FARGO_SPACING_LOG */) {
	const double f = std::pow((rmax / rmin), 1.0 / ((double)nrad - 2.0));
	for (int n = 0; n <= nrad; ++n)
	    r[n] = rmin * std::pow(f, (double)n - 1.0);
    } else if (spacing == 1 /* FARGO_SPACING_ARITH */) {
	const double interval = (rmax - rmin) / (double)(nrad - 2.0);
	for (int n = 0; n <= nrad; ++n)
	    r[n] = rmin + interval * (double)(n - 1.0);
    } else if (spacing == 2 /* FARGO_SPACING_EXP */) {
	const double cgf = std::pow((rmax / rmin), 1.0 / ((double)nrad - 2.0));
	const double first = rmin * (cgf - 1.0) * exp_cell_size_factor;
	const double f = (rmax - rmin) / first;
	double egf = 1.02;
	const double Nr = (double)nrad - 2.0;
	for (int i = 0; i < 500000; ++i)
	    egf = egf - ((std::pow(egf, Nr) - egf * f + f - 1)) / (Nr * std::pow(egf, Nr - 1.0) - f);
	// On coarse grids (about Nrad < 40 with the default ExponentialCellSizeFactor) the reference's Newton iteration falls into the
	// trivial root 1: its grid then stops short of Rmax (or is 0 / 0) and its ring lookups (find_cell_id.cpp:236-246, 1 / log(g)) return
	// clamped garbage.  Nothing to be faithful to: refused.
	if (!(egf > 1.0 + 1.0e-9))
	    refuse("RadialSpacing: Exponential: the growth factor iteration of init.cpp:112-128 degenerates to 1 on this grid (too few rings)");
	for (int n = 0; n <= nrad; ++n)
	    r[n] = rmin + first * (std::pow(egf, (double)n - 1.0) - 1.0) / (egf - 1.0);
    } else {
	refuse("RadialSpacing: custom grids (radii.dat) are read by `restart`, not by `start`");
    }
    return r;
}

// ---------------------------------------------------------------------------------------------
// N-body initial state.  Only what the hydro step and the snapshot records need.
struct BodyInit {
    std::string name;
    double mass = 0, x = 0, y = 0, vx = 0, vy = 0;
    double cubic_smoothing_factor = 0, accretion_efficiency = 0, radius = 0, temperature = 0, irradiation_rampuptime = 0,
	   rampuptime = 0;
    double distance_to_primary = 0, roche = 0;
    double semi_major_axis = 0, eccentricity = 0, mean_anomaly = 0, true_anomaly = 0, eccentric_anomaly = 0, pericenter_angle = 0,
	   orbital_period = 0, omega = 0;
};

// Theo.cpp:251-279
inline double init_l1(const double central_star_mass, const double other_star_mass)
{
    const double q = central_star_mass / (central_star_mass + other_star_mass);
    double x = std::pow(other_star_mass / (3.0 * central_star_mass), 1.0 / 3.0);
    double f, df;
    int counter = 0;
    do {
	counter++;
	if (counter > 10)
	    break;
	f = q / std::pow(1.0 - x, 2) - (1.0 - q) / std::pow(x, 2) - q + x;
	df = 2.0 * q / std::pow(1.0 - x, 3) + 2.0 * (1.0 - q) / std::pow(x, 3) + 1.0;
	x = x - f / df;
    } while (std::fabs(f) > 1e-14);
    return x;
}

// Theo.cpp:288-303: one Newton-Raphson iteration from the previous value
inline double update_l1(const double central_star_mass, const double other_star_mass, double l1)
{
    const double q = central_star_mass / (central_star_mass + other_star_mass);
    double x = l1;
    double f = q / std::pow(1.0 - x, 2) - (1.0 - q) / std::pow(x, 2) - q + x;
    double df = 2.0 * q / std::pow(1.0 - x, 3) + 2.0 * (1.0 - q) / std::pow(x, 3) + 1.0;
    x = x - f / df;
    return x;
}

// t_planet::calculate_orbital_elements (nbody/planet.cpp:488-573)
inline void orbital_elements(BodyInit &b, double x, double y, double vx, double vy, double com_mass, double G)
{
    auto zero = [&]() {
	b.omega = b.orbital_period = b.semi_major_axis = b.eccentricity = b.mean_anomaly = b.true_anomaly = b.eccentric_anomaly =
	    b.pericenter_angle = 0.0;
    };
    double E, V, PerihelionPA, temp;
    const double m = com_mass + b.mass;
    const double h = x * vy - y * vx;
    const double d = std::sqrt(x * x + y * y);
    if (d * d < 1e-26 || h == 0.0) { // is_distance_zero (util.cpp:105-110)
	zero();
	return;
    }
    const double Ax = x * vy * vy - y * vx * vy - G * m * x / d;
    const double Ay = y * vx * vx - x * vx * vy - G * m * y / d;
    const double e = std::sqrt(Ax * Ax + Ay * Ay) / G / m;
    const double a = h * h / G / m / (1.0 - e * e);
    if (e > 1.0 || e < 0 || a < 0.0) {
	zero();
	return;
    }
    const double P = 2.0 * M_PI * std::sqrt(std::pow(a, 3) / (m * G));
    const double omega = std::sqrt((m * G) / std::pow(a, 3));
    if (e != 0.0) {
	temp = (1.0 - d / a) / e;
	E = temp > 1.0 ? 0.0 : temp < -1.0 ? M_PI : std::acos(temp);
    } else {
	E = 0.0;
    }
    if ((x * y * (vy * vy - vx * vx) + vx * vy * (x * x - y * y)) < 0)
	E = -E;
    const double M = E - e * std::sin(E);
    if (e != 0.0) {
	temp = (a * (1.0 - e * e) / d - 1.0) / e;
	V = temp > 1.0 ? 0.0 : temp < -1.0 ? M_PI : std::acos(temp);
    } else {
	V = 0.0;
    }
    if (E < 0.0)
	V = -V;
    PerihelionPA = e != 0.0 ? std::atan2(Ay, Ax) : std::atan2(y, x);
    b.omega = omega, b.orbital_period = P, b.semi_major_axis = a, b.eccentricity = e, b.mean_anomaly = M, b.true_anomaly = V;
    b.eccentric_anomaly = V; // sic (planet.cpp:571)
    b.pericenter_angle = PerihelionPA;
}

// t_planetary_system::init_system for HydroFrameCenter: primary (nbody/planetary_system.cpp:68-134).
// `nbody`: the YAML's list of maps, keys lower-cased.
inline std::vector<BodyInit> init_bodies(const std::vector<std::map<std::string, std::string>> &nbody, const UnitSystem &U, double rmax,
					 const std::vector<double> *cic_radii = nullptr, double rmin = 0.0, double klahr_smoothing_radius = 0.0,
					 unsigned n_center = 1 /* parameters::n_bodies_for_hydroframe_center, 0: all */)
{
    std::vector<BodyInit> B;
    const double G = U.G.code;
    auto get = [](const std::map<std::string, std::string> &m, const char *k, const char *def) {
	auto it = m.find(k);
	return it == m.end() ? std::string(def) : it->second;
    };
    for (const auto &cfg : nbody) {
	if (!cfg.count("semi-major axis") || !cfg.count("mass"))
	    refuse("One of the planets does not have all of: semi-major axis and mass!");
	BodyInit p;
	double a = U.in_code_units(cfg.at("semi-major axis"), 'L');
	const double mass = U.in_code_units(cfg.at("mass"), 'M');
	const double e = atof(get(cfg, "eccentricity", "0.0").c_str());
	p.cubic_smoothing_factor = atof(get(cfg, "cubic smoothing factor", "0.0").c_str());
	p.accretion_efficiency = atof(get(cfg, "accretion efficiency", "0.0").c_str());
	p.radius = U.in_code_units(get(cfg, "radius", "0.009304813 au"), 'L');
	p.temperature = U.in_code_units(get(cfg, "temperature", "0.0 K"), 'K');
	p.irradiation_rampuptime = U.in_code_units(get(cfg, "irradiation ramp-up time", "0.0"), 'T');
	const double nu = atof(get(cfg, "trueanomaly", "0.0").c_str());
	double omega = atof(get(cfg, "argument of pericenter", "0.0").c_str());
	p.rampuptime = atof(get(cfg, "ramp-up time", "0.0").c_str());
	p.name = get(cfg, "name", ("planet" + std::to_string(B.size())).c_str());
	if (cic_radii) { // CICPLANET: the planet starts at a cell centre, find_cell_center_radius (planetary_system.cpp:149-159, :199-205)
	    if (e > 0)
		refuse("Centering planet in cell and eccentricity > 0 are not supported at the same time.");
	    if (a < rmin || a > rmax)
		refuse("Can not find cell center radius outside the grid");
	    size_t j = 0;
	    while ((*cic_radii)[j] < a)
		j++;
	    const double r0 = (*cic_radii)[j - 1], r1 = (*cic_radii)[j];
	    a = 2.0 / 3.0 * (std::pow(r1, 3) - std::pow(r0, 3));
	    a = a / (std::pow(r1, 2) - std::pow(r0, 2));
	}
	const std::string method = get(cfg, "accretion method", "kley");
	if (p.accretion_efficiency > 0.0 && method != "kley" && method != "sinkhole" && method != "viscous" && method != "no" && method != "none")
	    refuse("accretion method '" + method + "' is not supported by this driver (kley, sinkhole, viscous)");
	// initialize_planet_jacobi (:539-578) around the centre of mass of the bodies added so far
	auto jacobi = [&](double om) {
	    p.mass = mass;
	    double cx = 0, cy = 0, cm = 0; // get_center_of_mass / get_mass over the previously added bodies (:592-623)
	    for (auto &q : B) {
		cx += q.x * q.mass, cy += q.y * q.mass;
		cm += q.mass;
	    }
	    if (cm > 0.0)
		cx /= cm, cy /= cm;
	    else
		cx = cy = 0.0;
	    const double cos_ota = std::cos(om + nu), sin_ota = std::sin(om + nu);
	    const double cos_o = std::cos(om), sin_o = std::sin(om), cos_ta = std::cos(nu), sin_ta = std::sin(nu);
	    const double r = a * (1 - e * e) / (1 + e * cos_ta);
	    p.x = cx + r * cos_ota;
	    p.y = cy + r * sin_ota;
	    double v = 0.0;
	    if (a > 0.0)
		v = sqrt(G * (cm + mass) / (a * (1 - e * e)));
	    p.vx = v * (-cos_o * sin_ta - sin_o * (e + cos_ta));
	    p.vy = v * (-sin_o * sin_ta + cos_o * (e + cos_ta));
	};
	if (B.empty()) { // first body always goes to the origin (:487-494)
	    p.mass = mass;
	} else if (B.size() == 1) { // the first two share the barycentre (:495-533)
	    if (mass > B[0].mass)
		omega += M_PI;
	    jacobi(omega);
	    const double m1 = B[0].mass, m2 = p.mass;
	    const double x = p.x, y = p.y, vx = p.vx, vy = p.vy;
	    const double k1 = m2 / (m1 + m2);
	    B[0].x = -k1 * x, B[0].y = -k1 * y, B[0].vx = -k1 * vx, B[0].vy = -k1 * vy;
	    const double k2 = m1 / (m1 + m2);
	    p.x = k2 * x, p.y = k2 * y, p.vx = k2 * vx, p.vy = k2 * vy;
	} else {
	    jacobi(omega);
	}
	B.push_back(p);
    }
    if (B.empty())
	refuse("config has no nbody entries");
    if (klahr_smoothing_radius > 0.0) // the deprecated global KlahrSmoothingRadius (planetary_system.cpp:95-117), before the recentring
	for (auto &b : B)
	    if (std::sqrt(b.x * b.x + b.y * b.y) > 1.0e-10 && b.cubic_smoothing_factor == 0.0)
		b.cubic_smoothing_factor = klahr_smoothing_radius;
    if (n_center == 0 || n_center > B.size()) // init_hydro_frame_center (:283-305)
	n_center = (unsigned)B.size();
    { // move_to_hydro_frame_center (:750-768): the centre of mass of the first n_center bodies (HydroFrameCenter)
	double cx = 0, cy = 0, cvx = 0, cvy = 0, cm = 0;
	for (unsigned k = 0; k < n_center; ++k) {
	    cm += B[k].mass;
	    cx += B[k].x * B[k].mass, cy += B[k].y * B[k].mass;
	    cvx += B[k].vx * B[k].mass, cvy += B[k].vy * B[k].mass;
	}
	if (cm > 0)
	    cx = cx / cm, cy = cy / cm, cvx = cvx / cm, cvy = cvy / cm;
	else
	    cx = cy = cvx = cvy = 0.0;
	for (auto &b : B) {
	    const double x = b.x, y = b.y, vx = b.vx, vy = b.vy;
	    b.x = x - cx, b.y = y - cy, b.vx = vx - cvx, b.vy = vy - cvy;
	}
    }
    // compute_dist_to_primary (:941-964), init_roche_radii (:966-1004)
    if (B.size() < 2) {
	B[0].roche = 1.0;
	B[0].distance_to_primary = rmax;
    } else {
	for (size_t i = 1; i < B.size(); ++i) {
	    const double dx = B[i].x - B[0].x, dy = B[i].y - B[0].y;
	    const double dist = std::sqrt(std::pow(dx, 2) + std::pow(dy, 2));
	    B[i].distance_to_primary = dist;
	    if (i == 1)
		B[0].distance_to_primary = dist;
	}
	const double M = B[0].mass;
	for (size_t i = 1; i < B.size(); ++i) {
	    const double m = B[i].mass;
	    if (m == 0) {
		B[i].roche = 0.0, B[0].roche = 1.0;
		break;
	    }
	    if (M == 0) {
		B[0].roche = 0.0, B[i].roche = 1.0;
		break;
	    }
	    B[i].roche = M > m ? init_l1(M, m) : 1.0 - init_l1(m, M);
	    if (i == 1) // boundary_conditions::rof_planet, default 1 (:1000-1002)
		B[0].roche = 1.0 - B[i].roche;
	}
    }
    // calculate_orbital_elements (:773-805) about the centre of mass of the bodies inside; body 0 has none when it is the
    // frame centre on its own
    for (size_t i = (n_center == 1 ? 1 : 0); i < B.size(); ++i) {
	double cx = 0, cy = 0, cvx = 0, cvy = 0, cm = 0;
	for (size_t k = 0; k < i; ++k) {
	    cx += B[k].x * B[k].mass, cy += B[k].y * B[k].mass;
	    cvx += B[k].vx * B[k].mass, cvy += B[k].vy * B[k].mass;
	    cm += B[k].mass;
	}
	if (cm > 0.0)
	    cx /= cm, cy /= cm, cvx /= cm, cvy /= cm;
	else
	    cx = cy = cvx = cvy = 0.0;
	orbital_elements(B[i], B[i].x - cx, B[i].y - cy, B[i].vx - cvx, B[i].vy - cvy, cm, G);
    }
    return B;
}

// ---------------------------------------------------------------------------------------------
// Gas initial conditions, power-law profile (parameters::initialize_condition_profile)
struct DiskModel {
    double sigma0, sigma_slope, sigma_floor, h0, flaring, gamma, mu, Rgas, G, viscous_alpha, constant_viscosity, thickness_smoothing,
	tmin, tmax, omega_frame, imposed_drift;
    bool adiabatic, vradial_zero;
    // VazimuthalConsidersQuadropoleMoment: the binary's quadrupole moment (init_binary_quadropole_moment, Theo.cpp:58-78) in the
    // initial azimuthal velocity outside twice the binary separation and in the viscous-speed model
    bool quadrupole_support = false;
    double quadrupole_moment = 0.0, quadrupole_from_radius = 0.0;
    // SigmaCondition / EnergyCondition: 2D — the profile is read from a raw double[nrad][naz] file (t_polargrid::read2D)
    const std::vector<double> *sigma_in = nullptr, *energy_in = nullptr;
    // SigmaCondition: Nbody — the profiles are centred on the centre of mass of ALL bodies and the gas orbits it
    // (initialize_condition_profile_Nbody_centered: init.cpp:962-997, 1302-1346, 1473-1604)
    bool nbody_centered = false, energy_nbody_centered = false; // SigmaCondition (also the velocities) / EnergyCondition
    double cms_x = 0, cms_y = 0, vcms_x = 0, vcms_y = 0, nbody_mass = 0, density_correction_factor = 1.0;
    // CircumBinaryRing: a Gaussian ring on top of the profiles (add_gaussian_density_ring / _energy_ring, init.cpp:889-935, 1208-1255)
    bool cbd_ring = false;
    double cbd_ring_position = 4.5, cbd_ring_width = 0.6, cbd_decay_width = 0.6 * 1.4, cbd_decay_exponent = 0.75, cbd_ring_factor = 2.5;
    bool shock_tube = false; // ShockTube: 1 — Sod's shock tube along the radius (init_shock_tube_test, init.cpp:423-522)
    bool pure_keplerian = false; // InitializePureKeplerian (init.cpp:1607-1627)
    // ProfileCutoffOuter / Inner (parameters.cpp:728-742): Fermi-function cut-offs of the initial profiles (util.cpp:69-93)
    bool cutoff_outer = false, cutoff_inner = false;
    double cutoff_point_outer = 1.0e300, cutoff_width_outer = 1.0, cutoff_point_inner = 0.0, cutoff_width_inner = 1.0;
    bool spreading_ring = false; // SpreadingRing: the Bessel-function ring of init_spreading_ring_test (init.cpp:358-412)
    bool set_sigma0 = false; // SetSigma0: rescale Sigma0 so that the disk holds DiskMass (renormalize_sigma_and_report, init.cpp:1150-1188)
    double diskmass = 0.0;
};

struct InitialState {
    std::vector<double> sigma, energy, vrad, vazi; // [nrad][naz], v_rad [nrad + 1][naz]
};

namespace detail
{
// util.cpp:69-93
inline double cutoff_outer(double point, double width, double x) { return 1.0 / (1.0 + exp((x - point) / width)); }
inline double cutoff_inner(double point, double width, double x) { return 1.0 / (1.0 + exp((point - x) / width)); }
// Theo.cpp:122-153
inline double support_azi_pressure(const DiskModel &d, const double R)
{
    const double h = d.h0 * std::pow(R, d.flaring);
    return (2.0 * d.flaring - 1.0 - d.sigma_slope) * std::pow(h, 2);
}
inline double support_azi_smoothing_derivative(const DiskModel &d, const double R)
{
    const double h = d.h0 * std::pow(R, d.flaring);
    const double eps = d.thickness_smoothing;
    return (1.0 + (d.flaring + 1.0) * std::pow(h * eps, 2)) / std::pow(std::sqrt(1 + std::pow(h * eps, 2)), 3);
}
// initial_locally_isothermal_smoothed_v_az (Theo.cpp:166-180)
inline double v_az(const DiskModel &d, const double R, const double M)
{
    const double smoothing_derivative_2 = support_azi_smoothing_derivative(d, R);
    const double pressure_support_2 = support_azi_pressure(d, R);
    const double support = smoothing_derivative_2 + pressure_support_2;
    const double vk_2 = d.G * M / R;
    return std::sqrt(vk_2 * support);
}
// viscosity/viscous_radial_speed.cpp:36-95 (get_nu2) and :97-119 (get_sigma), without profile cutoffs
inline double get_sigma(const DiskModel &d, const double R)
{
    double density = d.sigma0 * std::pow(R, -d.sigma_slope);
    const double density_floor = d.sigma_floor * d.sigma0;
    if (d.cutoff_outer)
	density *= cutoff_outer(d.cutoff_point_outer, d.cutoff_width_outer, R);
    if (d.cutoff_inner)
	density *= cutoff_inner(d.cutoff_point_inner, d.cutoff_width_inner, R);
    density = std::max(density, density_floor);
    return density;
}
inline double get_nu2(const DiskModel &d, const double R, const double M, const double Sigma)
{
    const double v_k = std::sqrt(d.G * M / R);
    const double h = d.h0 * std::pow(R, d.flaring);
    double cutoff = 1.0;
    if (d.cutoff_outer)
	cutoff *= cutoff_outer(d.cutoff_point_outer, d.cutoff_width_outer, R);
    if (d.cutoff_inner)
	cutoff *= cutoff_inner(d.cutoff_point_inner, d.cutoff_width_inner, R);
    double cs_adb, H;
    if (d.adiabatic) {
	const double gamma = d.gamma;
	double energy = cutoff * 1.0 / (gamma - 1.0) * Sigma * std::pow(h * v_k, 2);
	const double energy_floor = d.tmin * Sigma / d.mu * d.Rgas / (gamma - 1.0);
	const double energy_ceil = d.tmax * Sigma / d.mu * d.Rgas / (gamma - 1.0);
	energy = std::max(energy, energy_floor);
	energy = std::min(energy, energy_ceil);
	cs_adb = std::sqrt(gamma * (gamma - 1.0) * energy / Sigma);
	const double cs_iso = std::sqrt((gamma - 1.0) * energy / Sigma);
	const double omega_k = v_k / R;
	H = cs_iso / omega_k;
    } else {
	cs_adb = h * v_k;
	H = h * R;
    }
    return d.viscous_alpha * cs_adb * H;
}
typedef double (*fn2)(const DiskModel &, double, double);
// viscous_speed::derive (:129-143): five-point stencil with h = 8e-4 x
inline double derive(const DiskModel &d, const double r, const double mass, fn2 f)
{
    const double x = r;
    const double h = 8.0e-4 * x;
    const double f1 = -1.0 * f(d, x + 2.0 * h, mass);
    const double f2 = 8.0 * f(d, x + h, mass);
    const double f3 = -8.0 * f(d, x - h, mass);
    const double f4 = 1.0 * f(d, x - 2.0 * h, mass);
    return (f1 + f2 + f3 + f4) / (12.0 * h);
}
// initial_locally_isothermal_smoothed_v_az_with_quadropole_moment (Theo.cpp:183-199)
inline double v_az_quadrupole(const DiskModel &d, const double R, const double M)
{
    const double pressure_support_2 = support_azi_pressure(d, R);
    double quadropole_support = 0.0;
    if (d.quadrupole_moment > 0.0)
	quadropole_support = 3.0 * d.quadrupole_moment / std::pow(R, 2);
    const double smoothing_derivative_2 = support_azi_smoothing_derivative(d, R);
    const double support = quadropole_support + smoothing_derivative_2 + pressure_support_2;
    const double vk_2 = d.G * M / R;
    return std::sqrt(vk_2 * support);
}
inline double get_w(const DiskModel &d, const double r, const double mass)
{
    return (d.quadrupole_support ? v_az_quadrupole(d, r, mass) : v_az(d, r, mass)) / r;
}
inline double get_r2_w(const DiskModel &d, const double r, const double mass)
{
    const double omega = get_w(d, r, mass);
    return std::pow(r, 2) * omega;
}
inline double get_nu_S_r3_dwdr(const DiskModel &d, const double r, const double mass)
{
    const double dw_dr = derive(d, r, mass, get_w);
    const double Sigma = get_sigma(d, r);
    const double nu = get_nu2(d, r, mass, Sigma);
    return nu * Sigma * std::pow(r, 3) * dw_dr;
}
// get_vr_with_numerical_viscous_speed (:184-197)
inline double viscous_vr(const DiskModel &d, const double r, const double mass)
{
    const double num = 1.0 / r * derive(d, r, mass, get_nu_S_r3_dwdr);
    const double Sigma = get_sigma(d, r);
    const double den = Sigma * derive(d, r, mass, get_r2_w);
    return num / den;
}
} // namespace detail

// init_gas_density (init.cpp:937-960), init_gas_energy (:1257-1300), init_gas_velocities (:1717-1771) for a disk around the
// hydro frame centre of mass M.  v_rad ring nrad stays 0 (the boundary stage fills the ghost interfaces).
inline InitialState init_gas(DiskModel &d, const std::vector<double> &radii, int nrad, int naz, double M)
{
    InitialState s;
    const size_t ns = (size_t)nrad * naz;
    s.sigma.assign(ns, 0.0);
    s.vazi.assign(ns, 0.0);
    s.vrad.assign((size_t)(nrad + 1) * naz, 0.0);
    if (d.adiabatic)
	s.energy.assign(ns, 0.0);
    std::vector<double> rmed(nrad), sigmed(nrad), siginf(nrad);
    for (int i = 0; i < nrad; ++i) { // init.cpp:185-190
	rmed[i] = 2.0 / 3.0 * (std::pow(radii[i + 1], 3) - std::pow(radii[i], 3));
	rmed[i] = rmed[i] / (std::pow(radii[i + 1], 2) - std::pow(radii[i], 2));
    }
    for (int i = 0; i < nrad; ++i) {
	const double density = d.sigma0 * std::pow(rmed[i], -d.sigma_slope);
	const double density_floor = d.sigma_floor * d.sigma0;
	const double sig = std::max(density, density_floor);
	for (int j = 0; j < naz; ++j)
	    s.sigma[(size_t)i * naz + j] = sig;
	if (d.adiabatic) { // initial_energy (Theo.cpp:86-98)
	    const double energy = 1.0 / (d.gamma - 1.0) * d.sigma0 * std::pow(d.h0, 2) *
				  std::pow(rmed[i], -d.sigma_slope - 1.0 + 2.0 * d.flaring) * d.G * M;
	    const double energy_floor = d.tmin * sig / d.mu * d.Rgas / (d.gamma - 1.0);
	    const double en = std::max(energy, energy_floor);
	    for (int j = 0; j < naz; ++j)
		s.energy[(size_t)i * naz + j] = en;
	}
    }
    const double dphi = 2.0 * M_PI / (double)naz;
    if (d.shock_tube) { // replaces every other density / energy initialisation, renormalisation included (init.cpp:269-299)
	for (int i = 0; i < nrad; ++i) {
	    double density = 1.0, energy = 2.5;
	    if (rmed[i] - rmed[0] > 0.5) {
		density = 0.125;
		energy = 2.0 * 0.125;
	    }
	    for (int j = 0; j < naz; ++j) {
		s.sigma[(size_t)i * naz + j] = density;
		if (d.adiabatic)
		    s.energy[(size_t)i * naz + j] = energy;
	    }
	}
    }
    if (!d.shock_tube && (d.nbody_centered || (d.adiabatic && d.energy_nbody_centered))) {
	for (int i = 0; i < nrad; ++i)
	    for (int j = 0; j < naz; ++j) {
		const double phi = (double)j * dphi;
		if (d.nbody_centered) { // density at the cell INTERFACE radius (sic, init.cpp:979-981)
		    const double rm = radii[i];
		    const double x = rm * std::cos(phi) - d.cms_x, y = rm * std::sin(phi) - d.cms_y;
		    const double r = std::sqrt(x * x + y * y);
		    const double density = d.sigma0 * std::pow(r, -d.sigma_slope) * d.density_correction_factor;
		    s.sigma[(size_t)i * naz + j] = std::max(density, d.sigma_floor * d.sigma0);
		}
		if (d.adiabatic && d.energy_nbody_centered) {
		    const double rm = rmed[i];
		    const double x = rm * std::cos(phi) - d.cms_x, y = rm * std::sin(phi) - d.cms_y;
		    const double r = std::sqrt(x * x + y * y);
		    const double energy = 1.0 / (d.gamma - 1.0) * d.sigma0 * std::pow(d.h0, 2) *
					  std::pow(r, -d.sigma_slope - 1.0 + 2.0 * d.flaring) * d.G * d.nbody_mass;
		    const double energy_floor = d.tmin * s.sigma[(size_t)i * naz + j] / d.mu * d.Rgas / (d.gamma - 1.0);
		    s.energy[(size_t)i * naz + j] = std::max(energy, energy_floor);
		}
	    }
    }
    // radius a profile cut-off is evaluated at
    auto cut_radius = [&](int i, int j, bool centred) {
	if (!centred)
	    return rmed[i];
	const double phi = (double)j * dphi;
	const double x = rmed[i] * std::cos(phi) - d.cms_x, y = rmed[i] * std::sin(phi) - d.cms_y;
	return std::sqrt(x * x + y * y);
    };
    if (d.sigma_in) // initialize_condition_read2D (init.cpp:1013-1017)
	s.sigma = *d.sigma_in;
    if (d.adiabatic && d.energy_in) // init.cpp:1355-1358
	s.energy = *d.energy_in;
    if (d.spreading_ring) { // init_spreading_ring_test (init.cpp:358-412; Speith & Kley 2003), replaces the profile
	const double R0 = 1.0;
	int R0_id = 0;
	for (int i = 0; i < nrad; ++i)
	    if (radii[i + 1] > R0 && R0 > radii[i])
		R0_id = i;
	const double Disk_Mass = d.diskmass;
	const double tau0 = 0.016;
	double Sigma0;
	{
	    const double x = rmed[R0_id] / R0;
	    const double I = std::cyl_bessel_i(0.25, 2.0 * x / tau0); // gsl_sf_bessel_Inu
	    Sigma0 = Disk_Mass / (M_PI * R0 * R0) * 1.0 / (tau0 * std::pow(x, 0.25)) * I * std::exp(-(1.0 + x * x) / tau0);
	}
	for (int i = 0; i < nrad; ++i) {
	    const double density_floor = Sigma0 * d.sigma_floor;
	    const double x = rmed[i] / R0;
	    const double I = std::cyl_bessel_i(0.25, 2.0 * x / tau0);
	    double density = Disk_Mass / (M_PI * R0 * R0) * 1.0 / (tau0 * std::pow(x, 0.25)) * I * std::exp(-(1.0 + x * x) / tau0);
	    density = std::max(density, density_floor);
	    for (int j = 0; j < naz; ++j) {
		s.sigma[(size_t)i * naz + j] = density;
		if (d.adiabatic)
		    s.energy[(size_t)i * naz + j] = 0.0;
	    }
	}
    }
    // profile cut-offs (init.cpp:1064-1147 for Sigma, :1361-1464 for the energy): outer first, then inner, each with its floor
    for (int pass = 0; pass < 2; ++pass) {
	if (!(pass == 0 ? d.cutoff_outer : d.cutoff_inner))
	    continue;
	for (int i = 0; i < nrad; ++i)
	    for (int j = 0; j < naz; ++j) {
		const double r = cut_radius(i, j, d.nbody_centered);
		const double f = pass == 0 ? detail::cutoff_outer(d.cutoff_point_outer, d.cutoff_width_outer, r)
					   : detail::cutoff_inner(d.cutoff_point_inner, d.cutoff_width_inner, r);
		const size_t l = (size_t)i * naz + j;
		s.sigma[l] = std::max(s.sigma[l] * f, d.sigma_floor * d.sigma0);
	    }
    }
    if (d.adiabatic)
	for (int pass = 0; pass < 2; ++pass) {
	    if (!(pass == 0 ? d.cutoff_outer : d.cutoff_inner))
		continue;
	    for (int i = 0; i < nrad; ++i)
		for (int j = 0; j < naz; ++j) {
		    const double r = cut_radius(i, j, d.energy_nbody_centered);
		    const double f = pass == 0 ? detail::cutoff_outer(d.cutoff_point_outer, d.cutoff_width_outer, r)
					       : detail::cutoff_inner(d.cutoff_point_inner, d.cutoff_width_inner, r);
		    const size_t l = (size_t)i * naz + j;
		    const double energy_floor = d.tmin * s.sigma[l] / d.mu * d.Rgas / (d.gamma - 1.0);
		    s.energy[l] = std::max(s.energy[l] * f, energy_floor);
		}
	}
    if (d.set_sigma0) { // quantities::gas_total_mass over the active rings (quantities.cpp:51-75), summed in index order
	double total_mass = 0.0;
	for (int i = 1; i < nrad - 1; ++i) {
	    const double surf = M_PI * (std::pow(radii[i + 1], 2) - std::pow(radii[i], 2)) / (double)naz;
	    for (int j = 0; j < naz; ++j)
		total_mass += surf * s.sigma[(size_t)i * naz + j];
	}
	d.sigma0 *= d.diskmass / total_mass;
	for (size_t l = 0; l < ns; ++l) {
	    s.sigma[l] *= d.diskmass / total_mass;
	    if (d.adiabatic)
		s.energy[l] *= d.diskmass / total_mass; // keeps the temperature
	}
    }
    // compute_azi_avg_Sigma (Theo.cpp:30-48): SigmaMed = ring mean, SigmaInf = its interpolation to the inner interfaces
    for (int i = 0; i < nrad; ++i) {
	double sum = 0.0;
	for (int j = 0; j < naz; ++j)
	    sum += s.sigma[(size_t)i * naz + j];
	sigmed[i] = sum / (double)naz;
    }
    siginf[0] = sigmed[0];
    for (int i = 1; i < nrad; ++i) {
	const double dr = (rmed[i] - rmed[i - 1]);
	siginf[i] = (sigmed[i - 1] * (rmed[i] - radii[i]) + sigmed[i] * (radii[i] - rmed[i - 1])) / dr;
    }
    if (d.cbd_ring) { // after renormalize_sigma_and_report, i.e. after SigmaMed / SigmaInf were taken (init.cpp:294-297)
	if (!d.adiabatic)
	    s.energy.assign(ns, 0.0); // add_gaussian_energy_ring runs whatever the equation of state: an isothermal run writes it out
	for (int i = 0; i < nrad; ++i)
	    for (int j = 0; j < naz; ++j) {
		const size_t l = (size_t)i * naz + j;
		const double r = cut_radius(i, j, d.nbody_centered);
		const double mass = d.nbody_centered ? d.nbody_mass : M;
		const double sigma_ring = d.sigma0 * std::pow(r, -d.sigma_slope);
		const double energy_ring = 1.0 / (d.gamma - 1.0) * d.sigma0 * std::pow(d.h0, 2) *
					   std::pow(r, -d.sigma_slope - 1.0 + 2.0 * d.flaring) * d.G * mass;
		double g;
		if (r < d.cbd_ring_position)
		    g = std::exp(-std::pow(d.cbd_ring_position - r, 2) / (2.0 * std::pow(d.cbd_ring_width, 2)));
		else
		    g = std::exp(-std::pow(r - d.cbd_ring_position, d.cbd_decay_exponent) / (2.0 * std::pow(d.cbd_decay_width, 2)));
		s.sigma[l] += sigma_ring * (d.cbd_ring_factor - 1.0) * g;
		s.energy[l] += energy_ring * (d.cbd_ring_factor - 1.0) * g;
	    }
    }
    if (d.nbody_centered) { // init_gas_velocities, first branch (init.cpp:1473-1604)
	const double mass = d.nbody_mass;
	auto v0 = [&](double r_com, double &vazi0, double &vr0) {
	    if (d.pure_keplerian) {
		vazi0 = std::sqrt(d.G * mass / r_com);
		DiskModel k = d; // initial_viscous_radial_speed (Theo.cpp:220-245)
		if (d.viscous_alpha > 0) {
		    const double sqrt_gamma = d.adiabatic ? std::sqrt(d.gamma) : 1.0;
		    const double v_k = std::sqrt(d.G * mass / r_com);
		    const double h = d.h0 * std::pow(r_com, d.flaring);
		    const double nu = d.viscous_alpha * (sqrt_gamma * h * v_k) * (h * r_com);
		    vr0 = -3.0 * nu / r_com * (-d.sigma_slope + 2.0 * d.flaring + 1.0);
		} else {
		    vr0 = -3.0 * d.constant_viscosity / r_com * (-d.sigma_slope + .5);
		}
		(void)k;
	    } else {
		vazi0 = (d.quadrupole_support && r_com > d.quadrupole_from_radius) ? detail::v_az_quadrupole(d, r_com, mass)
										    : detail::v_az(d, r_com, mass);
		vr0 = detail::viscous_vr(d, r_com, mass);
	    }
	    if (d.vradial_zero)
		vr0 = 0.0;
	};
	for (int i = 0; i <= nrad; ++i) // v_rad has nrad + 1 rings; the last one sits at Rinf[nrad - 1] (sic, :1491-1495)
	    for (int j = 0; j < naz; ++j) {
		const double phi = (double)j * dphi;
		const double r = i == nrad ? radii[nrad - 1] : radii[i];
		const double cell_x = r * std::cos(phi), cell_y = r * std::sin(phi);
		const double x_com = cell_x - d.cms_x, y_com = cell_y - d.cms_y;
		const double r_com = std::sqrt(x_com * x_com + y_com * y_com);
		double vazi0, vr0;
		v0(r_com, vazi0, vr0);
		const double vx_com = (vr0 * x_com - vazi0 * y_com) / r_com, vy_com = (vr0 * y_com + vazi0 * x_com) / r_com;
		const double vx = vx_com + d.vcms_x, vy = vy_com + d.vcms_y;
		s.vrad[(size_t)i * naz + j] = vx * std::cos(phi) + vy * std::sin(phi);
	    }
	for (int i = 0; i < nrad; ++i)
	    for (int j = 0; j < naz; ++j) {
		const double phi = ((double)j - 0.5) * dphi;
		const double r = rmed[i];
		const double cell_x = r * std::cos(phi), cell_y = r * std::sin(phi);
		const double x_com = cell_x - d.cms_x, y_com = cell_y - d.cms_y;
		const double r_com = std::sqrt(x_com * x_com + y_com * y_com);
		double vazi0, vr0;
		v0(r_com, vazi0, vr0);
		const double vx_com = (vr0 * x_com - vazi0 * y_com) / r_com, vy_com = (vr0 * y_com + vazi0 * x_com) / r_com;
		const double vx = vx_com + d.vcms_x, vy = vy_com + d.vcms_y;
		const double vaz = vy * std::cos(phi) - vx * std::sin(phi);
		s.vazi[(size_t)i * naz + j] = vaz - d.omega_frame * r;
	    }
	return s;
    }
    for (int i = 0; i < nrad; ++i) {
	const double r = rmed[i], ri = radii[i];
	double vazi, vrad;
	if (d.pure_keplerian) { // init.cpp:1607-1627 (sic: both at Rmed), Theo.cpp:207-245
	    double vr;
	    if (d.viscous_alpha > 0) {
		const double sqrt_gamma = d.adiabatic ? std::sqrt(d.gamma) : 1.0;
		const double v_k = std::sqrt(d.G * M / r);
		const double h = d.h0 * std::pow(r, d.flaring);
		const double cs = sqrt_gamma * h * v_k;
		const double H = h * r;
		const double nu = d.viscous_alpha * cs * H;
		vr = -3.0 * nu / r * (-d.sigma_slope + 2.0 * d.flaring + 1.0);
	    } else {
		const double nu = d.constant_viscosity;
		vr = -3.0 * nu / r * (-d.sigma_slope + .5);
	    }
	    vrad = vr;
	    vazi = std::sqrt(d.G * M / r) - d.omega_frame * r;
	} else {
	    vazi = (d.quadrupole_support && r > d.quadrupole_from_radius) ? detail::v_az_quadrupole(d, r, M) : detail::v_az(d, r, M);
	    vazi -= d.omega_frame * r;
	    vrad = d.imposed_drift * d.sigma0 / siginf[i] / ri;
	    if (!d.vradial_zero)
		vrad += detail::viscous_vr(d, ri, M);
	    else
		vrad = 0.0;
	}
	for (int j = 0; j < naz; ++j) {
	    s.vazi[(size_t)i * naz + j] = vazi;
	    s.vrad[(size_t)i * naz + j] = vrad;
	}
    }
    return s;
}

} // namespace finit
