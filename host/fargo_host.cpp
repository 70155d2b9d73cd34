// fargo_host.cpp — C++ host driver above the C ABI of include/fargo_b200.h.
//
// Mirrors the part of the reference's host that sits around the hydro hot path (paths relative to the reference's src/):
//   main.cpp:48-164 + sim::run simulation.cpp:505-558   -> run()            (time loop, monitor / snapshot cadence)
//   sim::CalculateTimeStep simulation.cpp:100-118        -> fargo_cfl
//   step_Euler simulation.cpp:148-267                    -> step(): indirect term, N-body kick, fargo_step, N-body drift
//   config.cpp / parameters.cpp / boundary_conditions/config.cpp / damping.cpp:185-271 -> Config, make_params()
//   restart.cpp:19-131, polargrid.cpp:301-353            -> load_snapshot()   (raw double[Nrad(+1)][Naz], no header)
//   output.cpp:249-330, polargrid.cpp:135-180, output.h:16-24 (misc.bin), nbody/planet.h:11-45 (nbodyK.bin)
//                                                        -> write_snapshot()
// `start <setup.yml>` begins a run from a FargoCPT setup file like `fargocpt start` (units, constants, radial grid, N-body
// initial state and the power-law disk of init.cpp: host/fargo_init.hpp) and writes snapshot 0 exactly as the reference does.
// It is also a drop-in for an EXISTING FargoCPT output directory: `restart N <dir>` reads the directory the reference wrote
// (config.yml of the snapshot, constants.yml / units.yml for the code-unit constants, dimensions.dat, used_rad.dat, the
// snapshot's fields, misc.bin, nbodyK.bin and snapshots/reference/ for the damping targets) and continues the run on the
// GPU, writing snapshots in the same binary format, so Tools/compare_binary_output.py can diff the two runs file by file.
// Out of scope here: initial conditions other than the power-law profile (read-in files, N-body-centred
// disks, test problems), units beyond the table in fargo_init.hpp, REBOUND (bodies are advanced with RK4 sub-steps, see
// nbody_integrate), monitors other than timestepLogging.dat and Quantities.dat.
//
// Multi-GPU (`--ranks N`): one process per GPU like the reference's one process per MPI rank (main.cpp:57,
// split.cpp:21-87).  The first process forks N - 1 siblings before anything touches CUDA, rank 0 hands the ncclUniqueId over
// through a file in the output directory, every rank runs this same driver (the bodies are integrated redundantly and
// identically on every rank, as in the reference; reductions are all-reduced behind the ABI), uploads its radial slab and
// writes ITS rings of every field file at their offset (pwrite; the reference's stitched MPI-IO writes, polargrid.cpp:135-180);
// everything else is written by rank 0 only.  `--rank R --ranks N` without forking is for an external launcher.
//
// The hydro arithmetic all happens behind the C ABI; build with -DFARGO_HOST_ORACLE to bind the same driver to the CPU
// oracle (TEST INFRASTRUCTURE: lets the host logic be tested without a GPU; never shipped).
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <fcntl.h>
#include <signal.h>
#include <sys/prctl.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <thread>
#include <unistd.h>
#include <tuple>
#include <vector>

#include "../include/fargo_b200.h"
#include "fargo_init.hpp"

#ifdef FARGO_HOST_ORACLE
extern "C" {
typedef struct fargo_oracle fargo_oracle;
fargo_oracle *fargo_oracle_create(const fargo_params *, const double *, int, int);
void fargo_oracle_destroy(fargo_oracle *);
int fargo_oracle_upload_field(fargo_oracle *, int, const double *);
int fargo_oracle_download_field(fargo_oracle *, int, double *);
int fargo_oracle_copy_initial_values(fargo_oracle *);
int fargo_oracle_set_bodies(fargo_oracle *, const fargo_bodies *);
int fargo_oracle_set_time(fargo_oracle *, double);
int fargo_oracle_init_derived(fargo_oracle *);
int fargo_oracle_cfl(fargo_oracle *, double *, double *);
int fargo_oracle_step(fargo_oracle *, double);
int fargo_oracle_stage_boundary(fargo_oracle *, double, int);
int fargo_oracle_disk_on_body_accel(fargo_oracle *, int, double, double *);
int fargo_oracle_kick(fargo_oracle *, double);
int fargo_oracle_drift(fargo_oracle *, double);
int fargo_oracle_finish_step(fargo_oracle *, double);
int fargo_oracle_accrete_kley(fargo_oracle *, double, double, double, double, double, double *);
int fargo_oracle_monitor_quantities(fargo_oracle *, double, double *);
int fargo_oracle_monitor_disk(fargo_oracle *, double, double, double, double *);
int fargo_oracle_circumplanetary_mass(fargo_oracle *, double, double, double, double *);
int fargo_oracle_keep_potential(fargo_oracle *, int);
int fargo_oracle_track_massflow(fargo_oracle *, int);
int fargo_oracle_clear_massflow(fargo_oracle *);
int fargo_oracle_track_boundary_flow(fargo_oracle *, int);
int fargo_oracle_boundary_flow(fargo_oracle *, double *, int);
int fargo_oracle_track_damping_mass(fargo_oracle *, int);
int fargo_oracle_damping_mass(fargo_oracle *, double *, int);
int fargo_oracle_accrete_sinkhole(fargo_oracle *, double, double, double, double, double, double *);
int fargo_oracle_accrete_viscous(fargo_oracle *, double, double, double, double, double, double *);
int fargo_oracle_correct_vazi(fargo_oracle *, double);
int fargo_oracle_set_pvte(fargo_oracle *, const fargo_pvte_consts *);
}
typedef fargo_oracle backend_ctx;
#define BK(name) fargo_oracle_##name
static const char *backend_error() { return "oracle call failed"; }
#else
typedef fargo_ctx backend_ctx;
#define BK(name) fargo_##name
static const char *backend_error() { return fargo_last_error(); }
#endif

[[noreturn]] static void die(const char *fmt, const std::string &a = "")
{ // the reference die()s on every error (LowTasks.cpp:67-120)
    fprintf(stderr, "fargocpt_b200: ");
    fprintf(stderr, fmt, a.c_str());
    fprintf(stderr, "\n");
    exit(1);
}
#define CHECK(call)                                         \
    do {                                                    \
	if ((call) != 0)                                    \
	    die("%s", std::string(#call ": ") + backend_error()); \
    } while (0)

static std::string lower(std::string s)
{
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return std::tolower(c); });
    return s;
}
static std::string trim(const std::string &s)
{
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}
static std::string unquote(std::string v)
{
    v = trim(v);
    if (v.size() >= 2 && ((v.front() == '\'' && v.back() == '\'') || (v.front() == '"' && v.back() == '"')))
	v = v.substr(1, v.size() - 2);
    return v;
}

// ---------------------------------------------------------------------------------------------
// The YAML subset FargoCPT setups use: top-level `Key: value  # comment` pairs and one `nbody:` list of maps.
// Keys are case-insensitive like config::Config (config.h:33-72).
struct Config {
    std::map<std::string, std::string> kv;
    std::vector<std::map<std::string, std::string>> nbody;

    static std::string strip_comment(const std::string &line)
    {
	bool q1 = false, q2 = false;
	for (size_t i = 0; i < line.size(); ++i) {
	    if (line[i] == '\'' && !q2)
		q1 = !q1;
	    else if (line[i] == '"' && !q1)
		q2 = !q2;
	    else if (line[i] == '#' && !q1 && !q2 && (i == 0 || std::isspace((unsigned char)line[i - 1])))
		return line.substr(0, i);
	}
	return line;
    }
    void load(const std::string &path)
    {
	std::ifstream f(path);
	if (!f)
	    die("cannot open config %s", path);
	std::string line;
	bool in_nbody = false;
	while (std::getline(f, line)) {
	    line = strip_comment(line);
	    if (trim(line).empty())
		continue;
	    const size_t indent = line.find_first_not_of(" \t");
	    std::string body = trim(line);
	    if (indent == 0 && body[0] != '-') {
		const size_t c = body.find(':');
		if (c == std::string::npos)
		    continue;
		const std::string key = lower(trim(body.substr(0, c)));
		const std::string val = unquote(body.substr(c + 1));
		in_nbody = (key == "nbody");
		if (!in_nbody)
		    kv[key] = val;
		continue;
	    }
	    if (!in_nbody)
		continue; // nested maps other than nbody are not part of the hot path's surface
	    if (body[0] == '-') {
		nbody.emplace_back();
		body = trim(body.substr(1));
		if (body.empty())
		    continue;
	    }
	    const size_t c = body.find(':');
	    if (c == std::string::npos || nbody.empty())
		continue;
	    nbody.back()[lower(trim(body.substr(0, c)))] = unquote(body.substr(c + 1));
	}
    }
    bool has(const std::string &k) const { return kv.count(lower(k)) != 0; }
    std::string str(const std::string &k, const std::string &def) const
    {
	auto it = kv.find(lower(k));
	return it == kv.end() ? def : it->second;
    }
    // "<number> [unit]": with a unit the value is divided by unit_in_cgs (config::cfg.get<T>(key, default, unit))
    static double number(const std::string &v, double unit_in_cgs = 0.0)
    {
	std::istringstream is(v);
	double x = 0;
	std::string unit;
	if (!(is >> x))
	    die("not a number: %s", v);
	if (is >> unit) {
	    if (unit_in_cgs > 0.0)
		return x / unit_in_cgs;
	    // a value with a unit on a key this driver has no conversion for must not silently lose the unit
	    die("value '%s' carries a unit, and this driver converts no units for that key", v);
	}
	return x;
    }
    double num(const std::string &k, double def, double unit_in_cgs = 0.0) const
    {
	auto it = kv.find(lower(k));
	return it == kv.end() ? def : number(it->second, unit_in_cgs);
    }
    bool flag(const std::string &k, bool def) const
    { // config.h:55-72: first letter y / t / 1
	auto it = kv.find(lower(k));
	if (it == kv.end() || it->second.empty())
	    return def;
	const char c = (char)std::tolower((unsigned char)it->second[0]);
	return c == 'y' || c == 't' || c == '1';
    }
};

// The keys the reference's config reader knows (its own default_config.yml, written with `WriteDefaultValues: yes`, plus the
// ones it only visits later).  config::Config::exit_on_unknown_key (config.cpp) refuses a setup with any other key — a typo
// must not silently fall back to a default — and so does `start`.  Keys starting with '_' are this repo's fixture bookkeeping.
static const char *const KNOWN_KEYS =
	"accretewithoutdiskfeedback adiabatic adiabaticindex alphacold alphahot alphamode artificialviscosity "
	"artificialviscositydissipation artificialviscosityfactor aspectratio aspectratiomode bitwiseexactrestarting "
	"bodyforcefrompotential cartesianparticles centerprofiledensitycorrectionfactor cfl cflmaxvar cicplanet "
	"circumbinarydecayexponent circumbinarydecaywidth circumbinaryring circumbinaryringenhancementfactor "
	"circumbinaryringposition circumbinaryringwidth compatibilitynostarsmoothing compatibilitysmoothingplanetloc "
	"constantviscosity coolingbeta coolingbetalocal coolingbetarampup coolingbetareference coolingradiativefactor "
	"corotationreferencebody correctdiskselfgravity cps cvnr damping dampingenergyinner dampingenergyouter "
	"dampinginnerlimit dampingouterlimit dampingsurfacedensityinner dampingsurfacedensityouter dampingtimefactor "
	"dampingtimeradiusouter dampingvazimuthalinner dampingvazimuthalouter dampingvradialinner dampingvradialouter "
	"densityfactor disk diskfeedback diskmass diskradiusmassfraction dowrite1dfiles energycondition energyfilename "
	"equationofstate exponentialcellsizefactor featuresize firstdt flaringindex fluxlimiter frame "
	"heatingcoolingcfllimit heatingviscous heatingviscousfactor hydroframecenter hydrogenmassfraction "
	"imposeddiskdrift indirecttermmode initializepurekeplerian initializevradialzero innerboundary "
	"innerboundaryenergy innerboundarysigma innerboundaryvazi innerboundaryvazikeplerianfactor innerboundaryvrad "
	"innerboundaryvradkeplerianfactor integrateparticles integrator kappaconst kappafactor keepdiskmassconstant "
	"klahrsmoothingradius l0 logafterrealseconds logaftersteps m0 massaccretionradius maximumtemperature "
	"minimumtemperature monitortimestep mu naz nbody nmonitor nrad nsnapshots numberofparticles omegaframe opacity "
	"outerboundary outerboundaryenergy outerboundarysigma outerboundaryvazi outerboundaryvazikeplerianfactor "
	"outerboundaryvrad outerboundaryvradkeplerianfactor outputdir particledensity particlediskgravityenabled "
	"particledustdiffusion particleeccentricity particlegasdragenabled particleintegrator "
	"particlemaximumescaperadius particlemaximumradius particleminimumescaperadius particleminimumradius "
	"particleradius particleradiusincreasefactor particlespeciesnumber particlesurfacedensityslope "
	"planetorbitdisktest polytropicconstant profilecutoffinner profilecutoffouter profilecutoffpointinner "
	"profilecutoffpointouter profilecutoffwidthinner profilecutoffwidthouter quantitiesradiuslimit radialspacing "
	"radialviscosityfactor radiativediffusion radiativediffusionautoomega radiativediffusionchecksolution "
	"radiativediffusiondumpdata radiativediffusioninnerboundary radiativediffusionmaxiterations "
	"radiativediffusionomega radiativediffusionouterboundary radiativediffusiontest1d radiativediffusiontest2d "
	"radiativediffusiontest2ddensity radiativediffusiontest2dk radiativediffusiontest2dsteps "
	"radiativediffusiontolerance randomfactor randomseed randomsigma rmax rmin rochelobeoverflow rofaveragingtime "
	"rofgamma rofplanet roframpingtime roftemperature rofvalue rofvariabletransfer scurvetype secondarydisk "
	"selfgravity selfgravityaspectratiochangethreshold selfgravitymode selfgravitystepsbetweenkernelupdate "
	"setsigma0 shocktube sigma0 sigmacondition sigmafilename sigmafloor sigmaslope spreadingring stabilizeviscosity "
	"surfacecooling t0 taufactor taumin temp0 temperature0 thicknesssmoothing thicknesssmoothingsg transport "
	"vazimuthalconsidersquadropolemoment viscaccretmassflowtest viscousalpha viscousoutflowspeed writealpha "
	"writealphagrav writealphagravmean writealphareynolds writealphareynoldsmean writeaspectratio "
	"writeateverytimestep writedefaultvalues writedensity writediskquantities writedivv writeeccentricity "
	"writeeccentricitychange writeeffectivegamma writeenergy writefirstadiabaticindex writegastorques writekappa "
	"writelightcurves writelightcurvesradii writemassflow writemeanmolecularweight writepdv writepotential "
	"writepressure writeqminus writeqplus writeradialdissipation writeradialluminosity writescaleheight "
	"writesgaccelazi writesgaccelrad writesoundspeed writetau writetaucool writetemperature writetgravitational "
	"writetoomre writetorques writetreynolds writevelocity writeverticalopticaldepth writeviscosity writevisibility ";
static const char *const KNOWN_NBODY_KEYS =
	"accretion efficiency accretion method argument of pericenter cubic smoothing factor eccentricity irradiate "
	"irradiation ramp-up time mass name radius ramp-up time semi-major axis temperature trueanomaly ";
static bool key_is_known(const char *list, const std::string &k)
{
    const std::string padded = " " + std::string(list);
    return padded.find(" " + k + " ") != std::string::npos;
}

// constants.yml / units.yml as written by the reference (output.cpp, units.cpp:270-310): blocks of `symbol:` / `code value:`
struct CodeConstants {
    double G = 1.0, R = 1.0, sigma_sb = 0.0, c_light = 0.0, temperature_unit_K = 1.0;
    double length_cgs = 1.0, mass_cgs = 1.0, time_cgs = 1.0; // units.yml: code -> cgs factors of the base units
    double density_cgs = 1.0, opacity_cgs = 1.0;	     // ... and of the two the opacity tables need (opacity.cpp:13-14)
    double energy_flux_cgs = 1.0, sigma_cgs = 0.0, G_cgs = 0.0; // what the S-curve cooling fit needs (SourceEuler.cpp:726-831)
    void load(const std::string &dir)
    {
	std::ifstream f(dir + "/constants.yml");
	if (!f)
	    die("cannot open %s/constants.yml (needed for the code-unit constants)", dir);
	std::string line, sym;
	while (std::getline(f, line)) {
	    const std::string t = trim(line);
	    if (t.rfind("symbol:", 0) == 0)
		sym = trim(t.substr(7));
	    else if (t.rfind("code value:", 0) == 0) {
		const double v = atof(t.substr(11).c_str());
		if (sym == "G")
		    G = v;
		else if (sym == "R")
		    R = v;
		else if (sym == "sigma")
		    sigma_sb = v;
		else if (sym == "c")
		    c_light = v;
	    } else if (t.rfind("cgs value:", 0) == 0) {
		const double v = atof(t.substr(10).c_str());
		if (sym == "G")
		    G_cgs = v;
		else if (sym == "sigma")
		    sigma_cgs = v;
	    }
	}
	std::ifstream u(dir + "/units.yml");
	std::string block;
	while (std::getline(u, line)) {
	    const std::string t = trim(line);
	    if (!line.empty() && !std::isspace((unsigned char)line[0]))
		block = t;
	    else if (t.rfind("cgs value:", 0) == 0) {
		const double v = atof(t.substr(10).c_str());
		if (block == "temperature:")
		    temperature_unit_K = v;
		else if (block == "length:")
		    length_cgs = v;
		else if (block == "mass:")
		    mass_cgs = v;
		else if (block == "time:")
		    time_cgs = v;
		else if (block == "density:")
		    density_cgs = v;
		else if (block == "opacity:")
		    opacity_cgs = v;
		else if (block == "energy flux:")
		    energy_flux_cgs = v;
	    }
	}
    }
};

// ---------------------------------------------------------------------------------------------
// parameters.cpp / Interpret.cpp / boundary_conditions/config.cpp / damping.cpp -> fargo_params
static int enum_of(const std::string &v, const std::vector<std::pair<std::string, int>> &table, const char *what)
{
    const std::string l = lower(v);
    for (auto &p : table)
	if (p.first == l)
	    return p.second;
    die((std::string("unknown ") + what + " '%s'").c_str(), v);
}
static fargo_params make_params(const Config &c, const CodeConstants &k, int nrad, int naz)
{
    fargo_params p;
    memset(&p, 0, sizeof(p));
    p.abi_version = FARGO_ABI_VERSION;
    p.nrad = nrad, p.naz = naz;
    const char sp = (char)std::tolower((unsigned char)c.str("RadialSpacing", "Arithmetic")[0]);
    p.radial_spacing = sp == 'l' ? FARGO_SPACING_LOG : sp == 'a' ? FARGO_SPACING_ARITH : sp == 'e' ? FARGO_SPACING_EXP : FARGO_SPACING_CUSTOM;
    p.rmin = c.num("Rmin", 0.0), p.rmax = c.num("Rmax", 0.0);
    // Physics this path does not implement must not be dropped silently: a setup that switches it on is refused by name.
    {
	const std::string eos = lower(c.str("EquationOfState", "Isothermal")); // Interpret.cpp:391-470
	if (eos != "isothermal" && eos != "iso" && eos != "adiabatic" && eos != "ideal" && eos != "pvte" && eos != "pvtelaw")
	    die("EquationOfState: %s is not supported by this driver (isothermal, ideal, pvte)", eos);
	if (c.has("Adiabatic"))
	    die("%s", std::string("the deprecated 'Adiabatic' flag is not supported; use EquationOfState"));
	const bool energy_equation = eos == "adiabatic" || eos == "ideal" || eos == "pvte" || eos == "pvtelaw"; // SubStep3 only runs then (simulation.cpp:203-205)
	const std::string sc = lower(c.str("SurfaceCooling", "No")); // parameters.cpp:394-406
	if (energy_equation && !(sc == "no" || sc == "off" || sc == "false" || sc == "thermal" || sc == "scurve"))
	    die("SurfaceCooling: %s is not supported by this driver (no, thermal, scurve)", sc);
	if (c.num("AlphaMode", 0) != 0 && !(c.num("AlphaMode", 0) == 1 && energy_equation && c.num("ViscousAlpha", 0.0) > 0))
	    die("AlphaMode: %s is not supported by this driver (0; 1 with the energy equation and ViscousAlpha > 0)", c.str("AlphaMode", ""));
	const std::pair<const char *, double> zero_only[] = {{"AspectRatioMode", 0}};
	for (auto &k : zero_only)
	    if (c.num(k.first, k.second) != k.second)
		die((std::string(k.first) + ": %s is not supported by this driver").c_str(), c.str(k.first, ""));
	for (const char *k : {"SelfGravity", "RadiativeDiffusion", "RocheLobeOverflow", "KeepDiskMassConstant", "PlanetOrbitDiskTest",
			      "CompatibilityNoStarSmoothing", "CompatibilitySmoothingPlanetLoc", "IntegrateParticles",
			      "ViscAccretMassflowTest"})
	    if (c.flag(k, false))
		die((std::string(k) + ": %s is not supported by this driver").c_str(), c.str(k, ""));
    }
    const std::string eos = lower(c.str("EquationOfState", "Isothermal"));
    p.pvte = (eos == "pvte" || eos == "pvtelaw") ? 1 : 0; // Interpret.cpp:453-491: an ideal gas with a variable adiabatic index
    p.adiabatic = (eos == "ideal" || eos == "adiabatic" || p.pvte) ? 1 : 0;
    p.energy_density_cgs = k.mass_cgs / (k.time_cgs * k.time_cgs);	    // units.cpp:270-377 (energy per area of the 2-D disk)
    p.surface_density_cgs = k.mass_cgs / (k.length_cgs * k.length_cgs);
    p.gamma = c.num("AdiabaticIndex", 1.4);
    p.mu = c.num("mu", 1.0);
    p.aspectratio_ref = c.num("AspectRatio", 0.05);
    { // Interpret.cpp:194-197: a reference temperature at r = 1 replaces the aspect ratio
	const double T0 = c.num("Temperature0", -1.0, k.temperature_unit_K);
	if (T0 > 0.0)
	    p.aspectratio_ref = std::sqrt(T0 * k.R / p.mu);
    }
    p.flaring_index = c.num("FlaringIndex", 0.0);
    // default "173 g/cm2" (parameters.cpp:625); a value with a unit has been converted by the caller (`start`) already
    p.sigma0 = c.has("Sigma0") ? c.num("Sigma0", 0.0) : 173.0 / (k.mass_cgs / (k.length_cgs * k.length_cgs));
    p.sigma_floor = c.num("SigmaFloor", 1e-9);
    p.sigma_slope = c.num("SigmaSlope", 0.0);
    p.minimum_temperature = Config::number(c.str("MinimumTemperature", "3 K"), k.temperature_unit_K);
    p.maximum_temperature = Config::number(c.str("MaximumTemperature", "1.0e300 K"), k.temperature_unit_K);
    p.G = k.G, p.Rgas = k.R, p.sigma_sb = k.sigma_sb, p.c_light = k.c_light;
    p.hydro_center_mass = 1.0; // overwritten from the bodies (global.cpp:146)
    p.cfl = c.num("CFL", 0.5);
    p.cfl_max_var = c.num("CFLmaxVar", 1.1);
    p.heating_cooling_cfl_limit = c.num("HeatingCoolingCFLlimit", 10.0); // parameters.cpp:797
    p.leapfrog = std::tolower((unsigned char)c.str("Integrator", "Euler")[0]) == 'e' ? 0 : 1;
    p.fast_transport = std::tolower((unsigned char)c.str("Transport", "FARGO")[0]) == 'f' ? 1 : 0;
    const std::string fl = c.str("FluxLimiter", "VanLeer"); // Interpret.cpp:640-664 compares case-sensitively
    p.flux_limiter = (fl == "mc" || fl == "m") ? FARGO_LIMITER_MC : FARGO_LIMITER_VANLEER;
    { // parameters.cpp:637-650: first letter only ("No" is none)
	const std::string av = c.str("ArtificialViscosity", "SN");
	const char l = av.empty() ? 's' : (char)std::tolower((unsigned char)av[0]);
	if (l != 'n' && l != 't' && l != 's')
	    die("Invalid setting for ArtificialViscosity: %s", av);
	p.artificial_viscosity = l == 'n' ? 0 : l == 't' ? 1 : 2;
    }
    p.artificial_viscosity_factor = c.num("ArtificialViscosityFactor", 1.41);
    p.artificial_viscosity_dissipation = c.flag("ArtificialViscosityDissipation", true);
    p.viscous_alpha = c.num("ViscousAlpha", 0.0);
    p.alpha_mode = (int)c.num("AlphaMode", 0), p.alpha_cold = c.num("AlphaCold", 0.01), p.alpha_hot = c.num("AlphaHot", 0.1); // parameters.cpp:704-706
    p.constant_viscosity = c.num("ConstantViscosity", 0.0);
    p.stabilize_viscosity = (int)c.num("StabilizeViscosity", 0);
    p.radial_viscosity_factor = c.num("RadialViscosityFactor", 1.0);
    p.heating_viscous = c.flag("HeatingViscous", true); // parameters.cpp:561
    p.heating_viscous_factor = c.num("HeatingViscousFactor", 1.0);
    p.cooling_beta = c.flag("CoolingBetaLocal", false);
    p.cooling_beta_value = c.num("CoolingBeta", 1.0);
    p.cooling_beta_ramp_up = c.num("CoolingBetaRampUp", 0.0);
    p.cooling_beta_reference = enum_of(c.str("CoolingBetaReference", "zero"),
				       {{"zero", 0}, {"reference", 1}, {"diskmodel", 2}, {"floor", 4}}, "CoolingBetaReference"); // parameters.cpp:451-463
    p.body_force_from_potential = c.flag("BodyForceFromPotential", true);
    if (!p.body_force_from_potential) // SourceEuler.cpp:348-353, 406-413: the kicks would read the ACCEL_RADIAL / ACCEL_AZIMUTHAL grids
	die("BodyForceFromPotential: no (body forces from the acceleration grids) is outside this path");
    p.thickness_smoothing = c.num("ThicknessSmoothing", 0.6);
    p.imposed_disk_drift = c.num("ImposedDiskDrift", 0.0);
    const std::vector<std::pair<std::string, int>> BC = {{"none", 0}, {"zerogradient", 1}, {"zero_gradient", 1}, {"outflow", 2},
							  {"reflecting", 3}, {"keplerian", 4}, {"reference", 5}};
    const std::vector<std::pair<std::string, int>> BC_VRAD_INNER = {{"none", 0}, {"zerogradient", 1}, {"zero_gradient", 1}, {"outflow", 2},
								     {"reflecting", 3}, {"keplerian", 4}, {"reference", 5}, {"viscous", 8}};
    const std::vector<std::pair<std::string, int>> BC_VAZI = {{"none", 0}, {"zerogradient", 1}, {"zero_gradient", 1}, {"keplerian", 4},
							       {"reference", 5}, {"zeroshear", 6}, {"balanced", 7}};
    const char *sides[2] = {"Inner", "Outer"};
    // Composite names first (boundary_conditions/config.cpp:345-436), then the individual keys, which overwrite what the composite set
    // (get_type, config.cpp:75-94).  The reference infers the INNER energy type from the OUTER side's name and an explicit
    // InnerBoundaryEnergy overwrites that name too (config.cpp:147); reproduced, since a setup means what the reference makes of it.
    std::string name_sigma[2], name_energy[2], name_vrad[2];
    if (!c.has("OuterBoundary")) // Interpret.cpp:290-293
	die("OuterBoundary doesn't exist. Old parameter file?");
    for (int s = 0; s < 2; ++s) {
	const std::string comp = lower(c.str(std::string(sides[s]) + "Boundary", "individual"));
	if (comp == "zerogradient")
	    name_sigma[s] = name_energy[s] = name_vrad[s] = "zerogradient";
	else if (comp == "outflow")
	    name_sigma[s] = name_energy[s] = "zerogradient", name_vrad[s] = "outflow";
	else if (comp == "reflecting")
	    name_sigma[s] = name_energy[s] = "zerogradient", name_vrad[s] = "reflecting";
	else if (comp == "reference")
	    name_sigma[s] = name_energy[s] = name_vrad[s] = "reference";
	else if (comp == "viscous" && s == 0)
	    name_sigma[s] = name_energy[s] = "zerogradient", name_vrad[s] = "viscous";
	else if (comp != "individual")
	    die((std::string(sides[s]) + "Boundary: %s is outside this path").c_str(), comp);
    }
    auto get_type = [&](const std::string &key, std::string &name) {
	if (c.has(key))
	    name = lower(c.str(key, ""));
	else if (name.empty())
	    die("Can not infer '%s' when 'InnerBoundary/OuterBoundary' is set to 'individual'", key);
	return name;
    };
    const std::string type_sigma[2] = {get_type("InnerBoundarySigma", name_sigma[0]), get_type("OuterBoundarySigma", name_sigma[1])};
    const std::string type_energy_inner = get_type("InnerBoundaryEnergy", name_energy[1]); // sic: the outer name, and in this order
    const std::string type_energy[2] = {type_energy_inner, get_type("OuterBoundaryEnergy", name_energy[1])};
    const std::string type_vrad[2] = {get_type("InnerBoundaryVrad", name_vrad[0]), get_type("OuterBoundaryVrad", name_vrad[1])};
    for (int s = 0; s < 2; ++s) {
	const std::string &bs = type_sigma[s], &be = type_energy[s], &bvr = type_vrad[s];
	p.bc_sigma[s] = enum_of(bs, BC, "boundary");
	p.bc_energy[s] = enum_of(be, BC, "boundary");
	p.bc_vrad[s] = enum_of(bvr, s == 0 ? BC_VRAD_INNER : BC, "v_rad boundary");
	p.keplerian_radial_factor[s] = c.num(std::string(sides[s]) + "BoundaryVradKeplerianFactor", 0.1); // config.cpp:220-255
	p.bc_vazi[s] = enum_of(c.str(std::string(sides[s]) + "BoundaryVazi", "keplerian"), BC_VAZI, "v_azi boundary");
	p.keplerian_azimuthal_factor[s] = c.num(std::string(sides[s]) + "BoundaryVaziKeplerianFactor", 1.0);
    }
    p.viscous_outflow_speed = c.num("ViscousOutflowSpeed", 1.0); // config.cpp:498
    p.correct_disk_selfgravity = c.flag("CorrectDiskSelfgravity", !c.flag("SelfGravity", false)); // parameters.cpp:699
    // radiative surface cooling, opacity (parameters.cpp:389-435, 628-632); heating_star is set by the caller from the bodies
    // (t_planetary_system::derive_config, planetary_system.cpp:137-146)
    p.cooling_surface = (p.adiabatic && lower(c.str("SurfaceCooling", "No")) == "thermal") ? 1 : 0;
    if (p.adiabatic && lower(c.str("SurfaceCooling", "No")) == "scurve") // parameters.cpp:374-403: ScurveType Kimura (default) | Ichikawa
	p.cooling_scurve = enum_of(c.str("ScurveType", "Kimura"), {{"ichikawa", 1}, {"kimura", 2}}, "ScurveType");
    p.length_cgs = k.length_cgs, p.mass_cgs = k.mass_cgs, p.energy_flux_cgs = k.energy_flux_cgs, p.sigma_sb_cgs = k.sigma_cgs, p.G_cgs = k.G_cgs;
    p.surface_cooling_factor = c.num("CoolingRadiativeFactor", 1.0);
    p.heating_star = 0;
    p.opacity = enum_of(c.str("Opacity", "Lin"), {{"lin", FARGO_OPACITY_LIN}, {"bell", FARGO_OPACITY_BELL}, {"constant", FARGO_OPACITY_CONST},
						 {"simple", FARGO_OPACITY_SIMPLE}}, "Opacity");
    p.kappa_const = c.num("KappaConst", 1.0, k.opacity_cgs);
    p.kappa_factor = c.num("KappaFactor", 1.0);
    p.tau_factor = c.num("TauFactor", 0.5);
    p.tau_min = c.num("TauMin", 0.01);
    p.density_factor = c.num("DensityFactor", std::sqrt(2.0 * M_PI));
    p.temperature_cgs = k.temperature_unit_K;
    p.density_cgs = k.density_cgs;
    p.opacity_code = 1.0 / k.opacity_cgs;
    p.damping = c.flag("Damping", false);
    p.damping_inner_limit = c.num("DampingInnerLimit", 1.05);
    p.damping_outer_limit = c.num("DampingOuterLimit", 0.95);
    p.damping_time_factor = c.num("DampingTimeFactor", 1.0);
    p.damping_time_radius_outer = c.num("DampingTimeRadiusOuter", p.rmax);
    const std::vector<std::pair<std::string, int>> DAMP = {{"none", 0}, {"initial", 1}, {"zero", 2}, {"mean", 3}};
    for (int s = 0; s < 2; ++s) {
	p.damp_vrad[s] = enum_of(c.str(std::string("DampingVRadial") + sides[s], "None"), DAMP, "damping type");
	p.damp_vazi[s] = enum_of(c.str(std::string("DampingVAzimuthal") + sides[s], "None"), DAMP, "damping type");
	p.damp_sigma[s] = enum_of(c.str(std::string("DampingSurfaceDensity") + sides[s], "None"), DAMP, "damping type");
	p.damp_energy[s] = enum_of(c.str(std::string("DampingEnergy") + sides[s], "None"), DAMP, "damping type");
    }
    return p;
}

// ---------------------------------------------------------------------------------------------
// binary records of a snapshot directory
#pragma pack(push, 1)
struct MiscEntry { // output.h:16-24 (48 bytes with natural alignment)
    uint32_t timestep, nTimeStep;
    double time, OmegaFrame, FrameAngle, last_dt;
    uint64_t N_iter;
};
#pragma pack(pop)
static_assert(sizeof(MiscEntry) == 48, "misc.bin layout");

struct PlanetRecord { // nbody/planet.h:11-45 planet_member_variables with the compiler's natural alignment (256 bytes)
    uint32_t timestep;
    uint32_t pad0;
    double mass, x, y, vx, vy, cubic_smoothing_factor, acc, accreted_mass;
    uint32_t planet_number;
    uint32_t pad1;
    double temperature, radius;
    uint8_t irradiate;
    uint8_t pad2[7];
    double irradiation_rampuptime, rampuptime;
    double disk_on_planet_acceleration[2], nbody_on_planet_acceleration[2];
    double distance_to_primary, dimensionless_roche_radius, circumplanetary_mass;
    double semi_major_axis, eccentricity, mean_anomaly, true_anomaly, eccentric_anomaly, pericenter_angle;
    double torque, gas_torque_acc, accretion_torque_acc, indirect_torque_acc;
};
static_assert(sizeof(PlanetRecord) == 256, "nbodyK.bin layout");

static std::vector<double> read_doubles(const std::string &path, size_t n, bool required = true)
{
    std::vector<double> v;
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) {
	if (required)
	    die("cannot open %s", path);
	return v;
    }
    v.resize(n);
    const size_t got = fread(v.data(), sizeof(double), n, f);
    fclose(f);
    if (got != n)
	die("short read on %s", path);
    return v;
}
static void mkdirs(const std::string &p)
{
    std::string cur;
    for (size_t i = 0; i <= p.size(); ++i) {
	if (i == p.size() || p[i] == '/') {
	    if (!cur.empty())
		mkdir(cur.c_str(), 0755);
	}
	if (i < p.size())
	    cur += p[i];
    }
}
static bool exists(const std::string &p)
{
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}

// ---------------------------------------------------------------------------------------------
struct Body {
    PlanetRecord rec; // carried through so the records we write keep the fields we do not touch
    double orbital_period = 0.0;
    double omega = 0.0; // t_planet::m_omega: sqrt(G (M + m) / a^3) of the osculating orbit (planet.cpp:520-521)
};

// planet.get_rampup_mass (nbody/planet.cpp:166-179)
static double rampup_mass(const Body &b, double t)
{
    double ramping = 1.0;
    if (b.rec.rampuptime > 0 && t < b.rec.rampuptime * b.orbital_period) {
	const double cs = std::cos(t * M_PI_2 / (b.rec.rampuptime * b.orbital_period));
	ramping = 1.0 - cs * cs;
    }
    return b.rec.mass * ramping;
}

// ComputeNbodyOnNbodyAccel (Pframeforce.cpp:225-251) + ComputeIndirectTermNbodyEuler (frame_of_reference.cpp:112-132),
// hydro frame centred on the centre of mass of the first n_center bodies (HydroFrameCenter: primary / binary / ... / all)
static void indirect_term_euler(const std::vector<Body> &b, double G, unsigned n_center, double &ix, double &iy)
{
    ix = iy = 0.0;
    if (b.size() < 2)
	return;
    double mass_center = 0.0;
    for (unsigned n = 0; n < n_center; ++n) { // the bodies that make up the hydro frame centre
	const double x = b[n].rec.x, y = b[n].rec.y;
	double ax = 0.0, ay = 0.0;
	for (size_t o = 0; o < b.size(); ++o) {
	    if (o == n)
		continue;
	    const double xo = b[o].rec.x, yo = b[o].rec.y, mass = b[o].rec.mass;
	    const double dist = std::sqrt(std::pow(x - xo, 2) + std::pow(y - yo, 2));
	    ax -= G * mass / std::pow(dist, 3) * (x - xo);
	    ay -= G * mass / std::pow(dist, 3) * (y - yo);
	}
	ix -= b[n].rec.mass * ax;
	iy -= b[n].rec.mass * ay;
	mass_center += b[n].rec.mass;
    }
    ix /= mass_center;
    iy /= mass_center;
}

// planetary_system.integrate (nbody/planetary_system.cpp:878-889) advances the bodies under their mutual gravity with
// REBOUND's IAS15 (accurate to rounding).  REBOUND is out of scope; the same system is advanced here with classical RK4
// over sub-steps h with (orbital frequency * h) <= 5e-4 (truncation error (w h)^5 / 120 = 3e-19 of the orbit per sub-step) and
// the increments accumulated with compensated (Kahan) summation, so that what is left is one rounding of the state per hydro
// step — the level IAS15 itself works at.
static void nbody_integrate(std::vector<Body> &b, double G, double dt)
{
    const size_t n = b.size();
    if (n < 2)
	return; // a single particle that does not move (:880-883)
    double wmax = 0.0;
    for (size_t i = 0; i < n; ++i)
	for (size_t j = i + 1; j < n; ++j) {
	    const double dx = b[i].rec.x - b[j].rec.x, dy = b[i].rec.y - b[j].rec.y;
	    const double d = std::sqrt(dx * dx + dy * dy);
	    wmax = std::max(wmax, std::sqrt(G * (b[i].rec.mass + b[j].rec.mass) / (d * d * d)));
	}
    int nsub = (int)std::ceil(wmax * dt / 5e-4);
    nsub = std::max(1, std::min(nsub, 100000));
    const double h = dt / nsub;
    std::vector<double> s(4 * n), k1(4 * n), k2(4 * n), k3(4 * n), k4(4 * n), tmp(4 * n), comp(4 * n, 0.0);
    auto rhs = [&](const std::vector<double> &q, std::vector<double> &dq) {
	for (size_t i = 0; i < n; ++i) {
	    double ax = 0, ay = 0;
	    for (size_t j = 0; j < n; ++j) {
		if (j == i)
		    continue;
		const double dx = q[4 * i] - q[4 * j], dy = q[4 * i + 1] - q[4 * j + 1];
		const double d2 = dx * dx + dy * dy, d = std::sqrt(d2);
		const double f = G * b[j].rec.mass / (d2 * d);
		ax -= f * dx, ay -= f * dy;
	    }
	    dq[4 * i] = q[4 * i + 2], dq[4 * i + 1] = q[4 * i + 3], dq[4 * i + 2] = ax, dq[4 * i + 3] = ay;
	}
    };
    for (size_t i = 0; i < n; ++i)
	s[4 * i] = b[i].rec.x, s[4 * i + 1] = b[i].rec.y, s[4 * i + 2] = b[i].rec.vx, s[4 * i + 3] = b[i].rec.vy;
    for (int it = 0; it < nsub; ++it) {
	rhs(s, k1);
	for (size_t q = 0; q < 4 * n; ++q)
	    tmp[q] = s[q] + 0.5 * h * k1[q];
	rhs(tmp, k2);
	for (size_t q = 0; q < 4 * n; ++q)
	    tmp[q] = s[q] + 0.5 * h * k2[q];
	rhs(tmp, k3);
	for (size_t q = 0; q < 4 * n; ++q)
	    tmp[q] = s[q] + h * k3[q];
	rhs(tmp, k4);
	for (size_t q = 0; q < 4 * n; ++q) { // Kahan: comp carries what the previous additions rounded away
	    const double inc = h / 6.0 * (k1[q] + 2.0 * k2[q] + 2.0 * k3[q] + k4[q]) - comp[q];
	    const volatile double t = s[q] + inc;
	    comp[q] = (t - s[q]) - inc;
	    s[q] = t;
	}
    }
    for (size_t i = 0; i < n; ++i)
	b[i].rec.x = s[4 * i], b[i].rec.y = s[4 * i + 1], b[i].rec.vx = s[4 * i + 2], b[i].rec.vy = s[4 * i + 3];
}

// ---------------------------------------------------------------------------------------------
struct Run {
    Config cfg;
    CodeConstants consts;
    fargo_params params;
    std::vector<double> radii;
    int nrad = 0, naz = 0;
    backend_ctx *ctx = nullptr;
    std::vector<Body> bodies;
    double omega_frame = 0.0, frame_angle = 0.0;
    // sim:: state (simulation.cpp:27-41)
    double time = 0.0, last_dt = 0.0;
    uint64_t n_iter = 0;
    unsigned n_snapshot = 0, n_monitor = 0;
    double monitor_timestep = 0.0;
    unsigned nmonitor = 1, nsnapshots = 1;
    int indirect_mode = 1;
    std::string refdir, outdir;
    // multi-rank (see the file header): `outdir` is where this rank writes the small files (rank 0: the output directory; the
    // others: a scratch directory nobody reads), `fielddir` the shared output directory the field files live in
    int rank = 0, nranks = 1;
    std::string fielddir;
    int own_lo = 0, own_hi = 0; // global rings [own_lo, own_hi) are written by this rank (write2D, polargrid.cpp:150-176)

    size_t cells(bool vector) const { return (size_t)(nrad + (vector ? 1 : 0)) * naz; }

    void set_bodies_on_device()
    { // CalculateNbodyPotential's planet setup (Pframeforce.cpp:27-36) + refframe::IndirectTerm
	fargo_bodies fb;
	memset(&fb, 0, sizeof(fb));
	fb.n = (int)bodies.size();
	for (int k = 0; k < fb.n; ++k) {
	    const Body &b = bodies[k];
	    fb.x[k] = b.rec.x, fb.y[k] = b.rec.y, fb.mass[k] = rampup_mass(b, time);
	    fb.cubic_smoothing_radius[k] = b.rec.dimensionless_roche_radius * b.rec.distance_to_primary * b.rec.cubic_smoothing_factor;
	    // irradiation_single (SourceEuler.cpp:538-564)
	    fb.temperature[k] = b.rec.temperature, fb.radius[k] = b.rec.radius;
	    const double tr = b.rec.irradiation_rampuptime;
	    fb.irradiation_ramp[k] = time < tr ? 1.0 - std::pow(std::cos(time * M_PI / 2.0 / tr), 2) : 1.0;
	}
	combine_indirect();
	fb.indirect_x = ind_x;
	fb.indirect_y = ind_y;
	fb.omega_frame = omega_frame;
	CHECK(BK(set_bodies)(ctx, &fb));
	last_fb = fb;
    }
    fargo_bodies last_fb;
    // refframe::ComputeIndirectTermFully (frame_of_reference.cpp:166-169)
    void combine_indirect() { ind_x = ind_disk_x + ind_nbody_x, ind_y = ind_disk_y + ind_nbody_y; }
    double ind_x = 0.0, ind_y = 0.0, ind_disk_x = 0.0, ind_disk_y = 0.0, ind_nbody_x = 0.0, ind_nbody_y = 0.0;
    unsigned n_center = 1; // parameters::n_bodies_for_hydroframe_center (Interpret.cpp:324-345)
    void read_hydro_frame_center()
    {
	const char c = (char)std::tolower((unsigned char)cfg.str("HydroFrameCenter", "primary")[0]);
	n_center = c == 'p' ? 1 : c == 'b' ? 2 : c == 't' ? 3 : c == 'q' ? 4 : c == 'a' ? 0 : 99;
	if (n_center == 99)
	    die("Invalid setting for HydroFrameCenter: %s", cfg.str("HydroFrameCenter", ""));
	if (n_center == 0 || n_center > cfg.nbody.size())
	    n_center = (unsigned)cfg.nbody.size(); // init_hydro_frame_center (planetary_system.cpp:283-292)
    }
    double hydro_frame_center_mass() const
    {
	double m = 0.0;
	for (unsigned k = 0; k < n_center && k < bodies.size(); ++k)
	    m += bodies[k].rec.mass;
	return m;
    }

    // refframe::ComputeIndirectTermNbody (frame_of_reference.cpp:134-164): the acceleration of the hydro frame centre (body 0)
    // by the other bodies over the coming `dt`, forward looking, so it is taken while the bodies are still at the start of it.
    // IndirectTermMode 1: the instantaneous N-body acceleration (Euler); mode 0 (the default): the velocity change of the
    // centre over dt from a predictor integration of a copy of the system (:146-157, planetary_system.cpp:671-705) — with
    // this driver's integrator in place of REBOUND's, like the bodies themselves.
    void compute_indirect_nbody(double dt)
    {
	ind_nbody_x = ind_nbody_y = 0.0;
	if (n_center == bodies.size())
	    return; // every body belongs to the frame centre (:137-141)
	if (indirect_mode == 1) {
	    indirect_term_euler(bodies, consts.G, n_center, ind_nbody_x, ind_nbody_y);
	} else if (dt != 0.0) {
	    std::vector<Body> predictor = bodies;
	    nbody_integrate(predictor, consts.G, dt);
	    double vx_old = 0.0, vy_old = 0.0, vx_new = 0.0, vy_new = 0.0, mass = 0.0; // planetary_system.cpp:671-705
	    for (unsigned i = 0; i < n_center; ++i) {
		const double m = bodies[i].rec.mass;
		mass += m;
		vx_old += bodies[i].rec.vx * m, vy_old += bodies[i].rec.vy * m;
		vx_new += predictor[i].rec.vx * m, vy_new += predictor[i].rec.vy * m;
	    }
	    if (mass > 0) {
		const double dvx = (vx_new - vx_old) / mass, dvy = (vy_new - vy_old) / mass;
		ind_nbody_x = -(dvx / dt);
		ind_nbody_y = -(dvy / dt);
	    }
	}
    }
    bool disk_feedback = false;

    // ComputeDiskOnNbodyAccel (Pframeforce.cpp:194-220) + UpdatePlanetVelocitiesWithDiskForce (:257-275) +
    // refframe::ComputeIndirectTermDisk (frame_of_reference.cpp:69-90); only with DiskFeedback: yes (simulation.cpp:155-160)
    void disk_feedback_kick(double dt)
    {
	disk_feedback_compute(dt);
	disk_feedback_apply(dt);
    }
    // step_LeapFrog evaluates the disk's pull BEFORE AccreteOntoPlanets and applies it AFTER (simulation.cpp:294-305, 352-408):
    // the two halves of disk_feedback_kick
    void disk_feedback_compute(double dt)
    {
	ind_disk_x = ind_disk_y = 0.0;
	if (!disk_feedback)
	    return;
	set_bodies_on_device(); // positions for the force integral
	for (size_t k = 0; k < bodies.size(); ++k) {
	    double a4[4];
	    CHECK(BK(disk_on_body_accel)(ctx, (int)k, bodies[k].rec.cubic_smoothing_factor, a4));
	    Body &b = bodies[k];
	    b.rec.disk_on_planet_acceleration[0] = a4[0] + a4[2]; // Force.cpp:117-119
	    b.rec.disk_on_planet_acceleration[1] = a4[1] + a4[3];
	    b.rec.torque = (b.rec.x * b.rec.disk_on_planet_acceleration[1] - b.rec.y * b.rec.disk_on_planet_acceleration[0]) * b.rec.mass;
	}
	double mass_center = 0.0;
	for (unsigned n = 0; n < n_center; ++n) {
	    ind_disk_x -= bodies[n].rec.mass * bodies[n].rec.disk_on_planet_acceleration[0];
	    ind_disk_y -= bodies[n].rec.mass * bodies[n].rec.disk_on_planet_acceleration[1];
	    mass_center += bodies[n].rec.mass;
	}
	ind_disk_x /= mass_center;
	ind_disk_y /= mass_center;
	for (auto &b : bodies) // the indirect torque monitor (frame_of_reference.cpp:92-107)
	    b.rec.indirect_torque_acc += (b.rec.x * ind_disk_y - b.rec.y * ind_disk_x) * b.rec.mass * dt;
    }
    void disk_feedback_apply(double dt)
    { // UpdatePlanetVelocitiesWithDiskForce (Pframeforce.cpp:257-275)
	if (!disk_feedback)
	    return;
	for (auto &b : bodies) {
	    b.rec.vx = b.rec.vx + dt * b.rec.disk_on_planet_acceleration[0];
	    b.rec.vy = b.rec.vy + dt * b.rec.disk_on_planet_acceleration[1];
	    b.rec.gas_torque_acc += b.rec.torque * dt; // t_planet::add_torque (Pframeforce.cpp:272)
	}
    }

    void load(const std::string &dir, unsigned nsnap, int device)
    {
	refdir = dir;
	const std::string sd = dir + "/snapshots/" + std::to_string(nsnap);
	config_path = exists(sd + "/config.yml") ? sd + "/config.yml" : dir + "/parameters/cfg.yml";
	cfg.load(config_path);
	consts.load(dir);
	{ // the snapshot's config.yml is the setup as the user wrote it: the same unit conversion `start` does
	    finit::UnitSystem U;
	    U.set_baseunits(cfg.str("l0", "1.0"), cfg.str("m0", "1.0"));
	    U.calculate();
	    convert_units(U);
	}
	{ // dimensions.dat: RMIN RMAX PHIMIN PHIMAX NRAD NAZ NGHRAD NGHAZ Radial_spacing (init.cpp:227-247)
	    std::ifstream f(dir + "/dimensions.dat");
	    if (!f)
		die("cannot open %s/dimensions.dat", dir);
	    std::string line;
	    while (std::getline(f, line)) {
		if (line.empty() || line[0] == '#')
		    continue;
		std::istringstream is(line);
		double a, b, c, d;
		is >> a >> b >> c >> d >> nrad >> naz;
	    }
	}
	radii = std::vector<double>();
	{
	    std::ifstream f(dir + "/used_rad.dat");
	    double r;
	    while (f >> r)
		radii.push_back(r);
	    if ((int)radii.size() != nrad + 1)
		die("used_rad.dat does not hold Nrad + 1 radii in %s", dir);
	}
	params = make_params(cfg, consts, nrad, naz);
	if (params.bc_vazi[0] == FARGO_BC_BALANCED || params.bc_vazi[1] == FARGO_BC_BALANCED)
	    die("%s", std::string("a Balanced v_azi boundary needs the disk model of `start`: not supported on restart by this driver"));
	if (params.pvte)
	    die("%s", std::string("restart with EquationOfState: PVTE is not supported by this driver (start only)"));
	// bodies
	for (size_t k = 0; k < cfg.nbody.size(); ++k) {
	    Body b;
	    FILE *f = fopen((sd + "/nbody" + std::to_string(k) + ".bin").c_str(), "rb");
	    if (!f || fread(&b.rec, sizeof(PlanetRecord), 1, f) != 1)
		die("cannot read nbody record %s", sd + "/nbody" + std::to_string(k) + ".bin");
	    fclose(f);
	    if (b.rec.semi_major_axis > 0.0) // planet.cpp: T = 2 pi sqrt(a^3 / (G (M + m)))
		b.orbital_period = 2.0 * M_PI * std::sqrt(std::pow(b.rec.semi_major_axis, 3) / (consts.G * (1.0 + b.rec.mass)));
	    if (b.rec.semi_major_axis > 0.0)
		b.omega = std::sqrt((consts.G * (1.0 + b.rec.mass)) / std::pow(b.rec.semi_major_axis, 3));
	    bodies.push_back(b);
	}
	if (bodies.empty())
	    die("config has no nbody entries: %s", sd);
	if (bodies.size() > FARGO_MAX_BODIES)
	    die("too many bodies in %s", sd);
	read_hydro_frame_center();
	refresh_orbital_parameters(); // period / omega from the interior centre-of-mass mass, not from a primary of mass 1
	params.hydro_center_mass = hydro_frame_center_mass(); // update_global_hydro_frame_center_mass
	disk_feedback = cfg.flag("DiskFeedback", true); // parameters.cpp:755
	indirect_mode = (int)cfg.num("IndirectTermMode", 0);
	MiscEntry m;
	{
	    FILE *f = fopen((sd + "/misc.bin").c_str(), "rb");
	    if (!f || fread(&m, sizeof(m), 1, f) != 1)
		die("cannot read %s/misc.bin", sd);
	    fclose(f);
	}
	time = m.time, last_dt = m.last_dt, n_iter = m.N_iter, n_snapshot = m.timestep, n_monitor = m.nTimeStep;
	omega_frame = m.OmegaFrame, frame_angle = m.FrameAngle;
	read_frame_settings();
	if (corotating && corotation_body >= bodies.size())
	    die("%s", std::string("CorotationReferenceBody does not exist"));
	monitor_timestep = cfg.num("MonitorTimestep", 1.0);
	nmonitor = (unsigned)cfg.num("Nmonitor", 10); // Interpret.cpp:200-201
	nsnapshots = (unsigned)cfg.num("Nsnapshots", 1000);
	create_context(device);
	// restart_load (restart.cpp:19-131): the four state fields
	const std::pair<int, const char *> state[4] = {{FARGO_SIGMA, "Sigma"}, {FARGO_VRAD, "vrad"}, {FARGO_VAZI, "vazi"}, {FARGO_ENERGY, "energy"}};
	for (auto &s : state)
	    CHECK(BK(upload_field)(ctx, s.first, read_doubles(sd + "/" + s.second + ".dat", cells(s.first == FARGO_VRAD)).data()));
	set_bodies_on_device();
	CHECK(BK(set_time)(ctx, time));
	CHECK(BK(init_derived)(ctx));
	// Q+/- as the reference stored them (needed by the first CFL; `BitwiseExactRestarting`, output.cpp:258-266)
	if (params.adiabatic) {
	    auto qp = read_doubles(sd + "/Qplus.dat", cells(false), false), qm = read_doubles(sd + "/Qminus.dat", cells(false), false);
	    if (!qp.empty() && !qm.empty()) {
		CHECK(BK(upload_field)(ctx, FARGO_QPLUS, qp.data()));
		CHECK(BK(upload_field)(ctx, FARGO_QMINUS, qm.data()));
	    } else {
		// restart.cpp:73-88: the reference then estimates them (compute_heating_cooling_for_CFL) and says so; here the first
		// CFL after the restart simply has no heating / cooling limit.  Write them with BitwiseExactRestarting: yes.
		fprintf(stderr, "fargocpt_b200: cannot read Qplus / Qminus in %s, no bitwise identical restarting possible "
				"(BitwiseExactRestarting: yes writes them)\n", sd.c_str());
	    }
	}
	// damping / beta-cooling targets: snapshots/reference (simulation.cpp:43-48), else the loaded state itself
	const std::string rd = dir + "/snapshots/reference";
	if (exists(rd + "/Sigma.dat")) {
	    const std::pair<int, const char *> ref[4] = {{FARGO_SIGMA0, "Sigma"}, {FARGO_VRAD0, "vrad"}, {FARGO_VAZI0, "vazi"}, {FARGO_ENERGY0, "energy"}};
	    for (auto &s : ref)
		CHECK(BK(upload_field)(ctx, s.first, read_doubles(rd + "/" + s.second + ".dat", cells(s.first == FARGO_VRAD0)).data()));
	} else {
	    CHECK(BK(copy_initial_values)(ctx));
	}
    }

    // Interpret.cpp:311-323, parameters.cpp:548-549
    void read_frame_settings()
    {
	const char f = (char)std::tolower((unsigned char)cfg.str("Frame", "Fixed")[0]);
	if (f != 'f' && f != 'c')
	    die("Invalid setting for Frame: %s", cfg.str("Frame", ""));
	corotating = f == 'c';
	corotation_body = (unsigned)cfg.num("CorotationReferenceBody", 1);
    }

    void create_context(int device)
    {
	// t_planetary_system::derive_config (planetary_system.cpp:137-146): a body with a temperature irradiates
	params.heating_star = 0;
	for (const Body &b : bodies)
	    if (params.adiabatic && b.rec.temperature > 0)
		params.heating_star = 1;
	if (params.cooling_scurve && params.heating_star) // the reference's irradiation would read the TAU_EFF of the previous scurve_cooling call
	    die("%s", std::string("SurfaceCooling: scurve with an irradiating body is not supported by this driver"));
#ifdef FARGO_HOST_ORACLE
	(void)device;
	ctx = fargo_oracle_create(&params, radii.data(), 0, 1);
	if (!ctx)
	    die("fargo_oracle_create failed");
	own_lo = 0, own_hi = nrad;
#else
	unsigned char uid[128];
	memset(uid, 0, sizeof(uid));
	if (nranks > 1) { // rank 0 creates the id, the others wait for the file
	    mkdirs(fielddir);
	    const std::string idfile = fielddir + "/.nccl_id";
	    if (rank == 0) {
		if (fargo_get_unique_id(uid) != 0)
		    die("fargo_get_unique_id: %s", fargo_last_error());
		const std::string tmp = idfile + ".tmp";
		FILE *f = fopen(tmp.c_str(), "wb");
		if (!f || fwrite(uid, 1, sizeof(uid), f) != sizeof(uid))
		    die("cannot write %s", tmp);
		fclose(f);
		if (rename(tmp.c_str(), idfile.c_str()) != 0)
		    die("cannot publish %s", idfile);
	    } else {
		bool got = false;
		for (int tries = 0; tries < 1200 && !got; ++tries) { // two minutes
		    FILE *f = fopen(idfile.c_str(), "rb");
		    if (f) {
			got = fread(uid, 1, sizeof(uid), f) == sizeof(uid);
			fclose(f);
		    }
		    if (!got)
			usleep(100000);
		}
		if (!got)
		    die("rank 0 did not publish %s", idfile);
	    }
	}
	if (fargo_ctx_create(&ctx, &params, radii.data(), rank, nranks, nranks > 1 ? uid : nullptr, device) != 0)
	    die("fargo_ctx_create: %s", fargo_last_error());
	{ // the rings this rank writes: its slab without the ghost rings of interior sides (split.cpp:57-61)
	    const int imin = fargo_local_imin(ctx), nloc = fargo_local_nrad(ctx);
	    own_lo = imin + (rank == 0 ? 0 : FARGO_CPUOVERLAP);
	    own_hi = imin + nloc - (rank == nranks - 1 ? 0 : FARGO_CPUOVERLAP);
	}
	rank_barrier(); // everybody has read the id: rank 0 may remove the file
	if (nranks > 1 && rank == 0)
	    unlink((fielddir + "/.nccl_id").c_str());
#endif
	if (cfg.flag("WriteMassFlow", false)) // parameters.cpp:334-335
	    CHECK(BK(track_massflow)(ctx, 1));
	if (cfg.flag("WriteDiskQuantities", true)) { // MassDelta's boundary flows (TransportEuler.cpp:578-608) and wave-damping terms
	    CHECK(BK(track_boundary_flow)(ctx, 1));
	    if (params.damping)
		CHECK(BK(track_damping_mass)(ctx, 1));
	}
    }

    // All ranks meet here (host side, through the shared output directory: arrival files of a numbered barrier).
    unsigned barrier_count = 0;
    void rank_barrier()
    {
	if (nranks < 2)
	    return;
	const std::string base = fielddir + "/.barrier" + std::to_string(barrier_count++) + ".";
	{
	    FILE *f = fopen((base + std::to_string(rank)).c_str(), "w");
	    if (!f)
		die("cannot write %s", base + std::to_string(rank));
	    fclose(f);
	}
	for (int r = 0; r < nranks; ++r) {
	    int tries = 0;
	    while (!exists(base + std::to_string(r))) {
		if (++tries > 6000)
		    die("rank %s did not reach the barrier", std::to_string(r));
		usleep(20000);
	    }
	}
	// the files of the barrier before this one can go (everybody has left it)
	if (barrier_count >= 2)
	    unlink((fielddir + "/.barrier" + std::to_string(barrier_count - 2) + "." + std::to_string(rank)).c_str());
    }

    // One field file of the shared output directory: this rank's rings at their offset (t_polargrid::write2D,
    // polargrid.cpp:135-180: MPI_File_write_at of the rank's active rings).  `global` is a whole-grid array of which only
    // this rank's rings need to be valid.
    void write_field_file(const std::string &rel, const double *global, bool vector) const
    {
	const std::string path = fielddir + "/" + rel;
	const int hi = own_hi + ((vector && rank == nranks - 1) ? 1 : 0);
	const size_t off = (size_t)own_lo * naz, cnt = (size_t)(hi - own_lo) * naz;
	const int fd = open(path.c_str(), O_WRONLY | O_CREAT | (nranks == 1 ? O_TRUNC : 0), 0644);
	if (fd < 0)
	    die("cannot write %s", path);
	if (nranks > 1 && rank == 0 && ftruncate(fd, (off_t)(cells(vector) * sizeof(double))) != 0)
	    die("cannot size %s", path);
	const char *src = (const char *)(global + off);
	size_t left = cnt * sizeof(double);
	off_t at = (off_t)(off * sizeof(double));
	while (left > 0) {
	    const ssize_t w = pwrite(fd, src, left, at);
	    if (w <= 0)
		die("cannot write %s", path);
	    src += w, at += w, left -= (size_t)w;
	}
	close(fd);
    }

    // Values that may carry units become plain code-unit numbers (config::Config::get<double>(key, default, unit): Interpret.cpp,
    // parameters.cpp).  ONE place, used by `start` and by `restart` (whose config.yml is a verbatim copy of the setup).
    void convert_units(const finit::UnitSystem &U)
    {
	const std::pair<const char *, char> dims[] = {
	    {"Rmin", 'L'}, {"Rmax", 'L'}, {"Sigma0", 'S'}, {"MonitorTimestep", 'T'}, {"FirstDT", 'T'}, {"DampingTimeRadiusOuter", 'L'},
	    {"ConstantViscosity", 'V'} /* L0^2/T0, Interpret.cpp:586 */, {"CoolingBetaRampUp", 'T'} /* parameters.cpp:445 */,
	    {"OmegaFrame", 'F'} /* 1/T0, Interpret.cpp:323 */, {"QuantitiesRadiusLimit", 'L'}};
	if (!cfg.has("Sigma0"))
	    cfg.kv["sigma0"] = "173 g/cm2"; // parameters.cpp:625
	for (auto &k : dims)
	    if (cfg.has(k.first))
		cfg.kv[lower(k.first)] = finit::UnitSystem::num17(U.in_code_units(cfg.str(k.first, ""), k.second));
    }

    // main.cpp:48-164 up to the first output, for `start`: units and constants, grid, bodies, init_physics (init.cpp:255-345)
    std::string config_path;
    bool started_fresh = false;
    std::map<int, std::vector<double>> derived_at_init; // derived fields of snapshot 0 (see start())
    void start(const std::string &cfgfile, int device)
    {
	config_path = cfgfile;
	cfg.load(cfgfile);
	// what this driver's initial conditions do not cover is refused by name
	const std::pair<const char *, const char *> off[] = {
	    {"RandomSigma", "no"}, {"SelfGravity", "no"}, {"IntegrateParticles", "no"},
	    {"RadiativeDiffusion", "no"}, {"SecondaryDisk", "no"}, {"CbdRing", "no"}};
	for (auto &k : off) {
	    const std::string v = lower(cfg.str(k.first, k.second));
	    if (!(v.empty() || v[0] == 'n' || v[0] == 'f' || v[0] == '0'))
		die((std::string(k.first) + ": %s is not supported by `fargocpt_b200 start`").c_str(), v);
	}
	if (!cfg.flag("Disk", true))
	    die("%s", std::string("Disk: no (an N-body run without gas) is not what this driver is for"));
	for (const char *k : {"SigmaCondition", "EnergyCondition"}) {
	    const char c0 = (char)std::tolower((unsigned char)cfg.str(k, "Profile")[0]);
	    if (c0 != 'p' && c0 != '2' && c0 != 'n')
		die((std::string(k) + ": only 'Profile', 'Nbody' and '2D' are supported by `fargocpt_b200 start` (got %s)").c_str(), cfg.str(k, ""));
	}
	read_hydro_frame_center();
	for (auto &kv_ : cfg.kv) // config::Config::exit_on_unknown_key
	    if (kv_.first[0] != '_' && !key_is_known(KNOWN_KEYS, kv_.first))
		die("Unknown key in the setup file: '%s' (the reference refuses unknown keys too)", kv_.first);
	for (auto &b : cfg.nbody)
	    for (auto &kv_ : b)
		if (!key_is_known(KNOWN_NBODY_KEYS, kv_.first))
		    die("Unknown key in an nbody entry of the setup file: '%s'", kv_.first);
	read_frame_settings();
	finit::UnitSystem U;
	U.set_baseunits(cfg.str("l0", "1.0"), cfg.str("m0", "1.0"));
	U.calculate();
	consts.G = U.G.code, consts.R = U.R.code, consts.sigma_sb = U.sigma.code, consts.c_light = U.c.code;
	consts.temperature_unit_K = U.temperature;
	consts.length_cgs = U.length, consts.mass_cgs = U.mass, consts.time_cgs = U.time;
	consts.density_cgs = U.density, consts.opacity_cgs = U.opacity;
	consts.energy_flux_cgs = U.energy_flux, consts.sigma_cgs = U.sigma.cgs, consts.G_cgs = U.G.cgs;
	const int shock_tube = (int)cfg.num("ShockTube", 0);
	if (shock_tube != 0 && shock_tube != 1)
	    die("ShockTube: %s is not supported by this driver (1: the ideal-gas tube)", cfg.str("ShockTube", ""));
	convert_units(U);
	if (!cfg.has("Rmin") || !cfg.has("Rmax"))
	    die("%s: Rmin and Rmax are required", cfgfile);
	nrad = (int)cfg.num("Nrad", 64), naz = (int)cfg.num("Naz", 64);
	{ // Interpret.cpp:204-227: resolution from cells per scale height
	    const double cps = cfg.num("cps", -1.0), H = cfg.num("AspectRatio", 0.05), rmin = cfg.num("Rmin", 0), rmax = cfg.num("Rmax", 0);
	    const char sp = (char)std::tolower((unsigned char)cfg.str("RadialSpacing", "Arithmetic")[0]);
	    if (cps > 0) {
		if (sp == 'a') {
		    nrad = (int)std::round(cps * (rmax - rmin) / H);
		    naz = (int)std::round(2 * M_PI / (rmax - rmin) * nrad);
		} else if (sp == 'l') {
		    nrad = (int)std::round(std::log(rmax / (double)rmin) / std::log(1 + H / cps));
		    naz = (int)std::round(2 * M_PI / (std::pow(rmax / (double)rmin, 1.0 / (double)nrad) - 1));
		} else {
		    die("%s", std::string("Setting resolution is not supported for the selected radial grid spacing."));
		}
	    }
	}
	params = make_params(cfg, consts, nrad, naz);
	if (shock_tube) { // init_shock_tube_test sets G = R = 1 in memory AFTER the parameters were converted and the unit / constant
			  // files were written (init.cpp:506-513, main.cpp:85-86): the files keep the l0 / m0 values
	    consts.G = 1.0, consts.R = 1.0;
	    params.G = 1.0, params.Rgas = 1.0;
	}
	radii = finit::make_radii(params.radial_spacing, nrad, params.rmin, params.rmax, cfg.num("ExponentialCellSizeFactor", 1.41));
	// bodies
	const bool cic = cfg.flag("CICPLANET", false);
	const auto B = finit::init_bodies(cfg.nbody, U, params.rmax, cic ? &radii : nullptr, params.rmin, cfg.num("KlahrSmoothingRadius", 0.0), n_center);
	if (B.size() > FARGO_MAX_BODIES)
	    die("too many bodies in %s", cfgfile);
	for (size_t k = 0; k < B.size(); ++k) {
	    Body b;
	    memset(&b.rec, 0, sizeof(b.rec));
	    const finit::BodyInit &q = B[k];
	    b.rec.mass = q.mass, b.rec.x = q.x, b.rec.y = q.y, b.rec.vx = q.vx, b.rec.vy = q.vy;
	    b.rec.cubic_smoothing_factor = q.cubic_smoothing_factor, b.rec.acc = q.accretion_efficiency;
	    b.rec.planet_number = (uint32_t)k;
	    b.rec.temperature = q.temperature, b.rec.radius = q.radius;
	    b.rec.irradiation_rampuptime = q.irradiation_rampuptime, b.rec.rampuptime = q.rampuptime;
	    b.rec.distance_to_primary = q.distance_to_primary, b.rec.dimensionless_roche_radius = q.roche;
	    b.rec.semi_major_axis = q.semi_major_axis, b.rec.eccentricity = q.eccentricity, b.rec.mean_anomaly = q.mean_anomaly;
	    b.rec.true_anomaly = q.true_anomaly, b.rec.eccentric_anomaly = q.eccentric_anomaly, b.rec.pericenter_angle = q.pericenter_angle;
	    b.orbital_period = q.orbital_period;
	    b.omega = q.omega;
	    bodies.push_back(b);
	}
	if (bodies.size() == 2) { // a binary: both carry the secondary's elements (planetary_system.cpp:797-802)
	    PlanetRecord &p0 = bodies[0].rec;
	    const PlanetRecord &p1 = bodies[1].rec;
	    p0.semi_major_axis = p1.semi_major_axis, p0.eccentricity = p1.eccentricity, p0.mean_anomaly = p1.mean_anomaly;
	    p0.true_anomaly = p1.true_anomaly, p0.eccentric_anomaly = p1.eccentric_anomaly, p0.pericenter_angle = p1.pericenter_angle;
	    bodies[0].orbital_period = bodies[1].orbital_period;
	}
	params.hydro_center_mass = hydro_frame_center_mass(); // update_global_hydro_frame_center_mass (planetary_system.cpp:724-727)
	disk_feedback = cfg.flag("DiskFeedback", true);
	indirect_mode = (int)cfg.num("IndirectTermMode", 0);
	omega_frame = cfg.num("OmegaFrame", 0.0), frame_angle = 0.0;
	if (corotating) { // init_physics (init.cpp:259-263)
	    if (corotation_body >= bodies.size())
		die("%s", std::string("CorotationReferenceBody does not exist"));
	    omega_frame = bodies[corotation_body].omega;
	}
	monitor_timestep = cfg.num("MonitorTimestep", 1.0);
	nmonitor = (unsigned)cfg.num("Nmonitor", 10); // Interpret.cpp:201
	nsnapshots = (unsigned)cfg.num("Nsnapshots", 1000);
	// init_physics (init.cpp:255-345)
	finit::DiskModel d;
	d.sigma0 = params.sigma0, d.sigma_slope = params.sigma_slope, d.sigma_floor = params.sigma_floor;
	d.h0 = params.aspectratio_ref, d.flaring = params.flaring_index, d.gamma = params.gamma, d.mu = params.mu;
	d.Rgas = consts.R, d.G = consts.G, d.viscous_alpha = params.viscous_alpha, d.constant_viscosity = params.constant_viscosity;
	d.thickness_smoothing = params.thickness_smoothing, d.tmin = params.minimum_temperature, d.tmax = params.maximum_temperature;
	d.omega_frame = omega_frame, d.imposed_drift = params.imposed_disk_drift;
	d.adiabatic = params.adiabatic != 0, d.vradial_zero = cfg.flag("InitializeVradialZero", false);
	std::vector<double> sigma_in, energy_in; // parameters.cpp:586-611: SigmaFilename / EnergyFilename
	if (cfg.str("SigmaCondition", "Profile")[0] == '2') {
	    sigma_in = read_doubles(cfg.str("SigmaFilename", ""), cells(false));
	    d.sigma_in = &sigma_in;
	}
	if (params.adiabatic && cfg.str("EnergyCondition", "Profile")[0] == '2') {
	    energy_in = read_doubles(cfg.str("EnergyFilename", ""), cells(false));
	    d.energy_in = &energy_in;
	}
	d.cbd_ring = cfg.flag("CircumBinaryRing", false); // parameters.cpp:712-725
	if (cfg.has("CircumBinaryRingPosition"))
	    d.cbd_ring_position = U.in_code_units(cfg.str("CircumBinaryRingPosition", ""), 'L');
	if (cfg.has("CircumBinaryRingWidth"))
	    d.cbd_ring_width = U.in_code_units(cfg.str("CircumBinaryRingWidth", ""), 'L');
	d.cbd_decay_width = cfg.has("CircumBinaryDecayWidth") ? U.in_code_units(cfg.str("CircumBinaryDecayWidth", ""), 'L') : d.cbd_ring_width * 1.4;
	d.cbd_decay_exponent = cfg.num("CircumBinaryDecayExponent", 0.75);
	d.cbd_ring_factor = cfg.num("CircumBinaryRingEnhancementFactor", 2.5);
	d.pure_keplerian = cfg.flag("InitializePureKeplerian", false);
	if (std::tolower((unsigned char)cfg.str("SigmaCondition", "Profile")[0]) == 'n' ||
	    std::tolower((unsigned char)cfg.str("EnergyCondition", "Profile")[0]) == 'n') { // parameters.cpp:577-581, 600-604: either sets the density's
	    d.nbody_centered = true;
	    // parameters.cpp:577-611: 'Nbody' sets the energy condition too, but the EnergyCondition switch that follows (default
	    // Profile) overrides it again: the energy is N-body-centred only when EnergyCondition says so as well
	    d.energy_nbody_centered = std::tolower((unsigned char)cfg.str("EnergyCondition", "Profile")[0]) == 'n';
	    d.density_correction_factor = cfg.num("CenterProfileDensityCorrectionFactor", 1.0);
	    double m = 0, x = 0, y = 0, vx = 0, vy = 0; // get_center_of_mass / _velocity over all bodies (planetary_system.cpp:600-645)
	    for (auto &b : bodies) {
		m += b.rec.mass;
		x += b.rec.x * b.rec.mass, y += b.rec.y * b.rec.mass, vx += b.rec.vx * b.rec.mass, vy += b.rec.vy * b.rec.mass;
	    }
	    d.nbody_mass = m;
	    if (m > 0)
		d.cms_x = x / m, d.cms_y = y / m, d.vcms_x = vx / m, d.vcms_y = vy / m;
	}
	if (cfg.flag("VazimuthalConsidersQuadropoleMoment", false) && bodies.size() > 1) {
	    d.quadrupole_support = true;
	    d.quadrupole_from_radius = 2.0 * bodies[1].rec.distance_to_primary; // init.cpp:1728-1730
	    if (n_center == 2) { // init_binary_quadropole_moment (Theo.cpp:58-78)
		const double a_b = bodies[1].rec.semi_major_axis, m1 = bodies[0].rec.mass, m2 = bodies[1].rec.mass;
		const double q_b = m2 < m1 ? m2 / m1 : m1 / m2;
		const double e_b = bodies[1].rec.eccentricity;
		d.quadrupole_moment = std::pow(a_b, 2) / 4.0 * q_b / std::pow((1.0 + q_b), 2) * (1.0 + 3.0 / 2.0 * std::pow(e_b, 2));
	    }
	}
	d.cutoff_outer = cfg.flag("ProfileCutoffOuter", false), d.cutoff_inner = cfg.flag("ProfileCutoffInner", false);
	if (cfg.has("ProfileCutoffPointOuter"))
	    d.cutoff_point_outer = U.in_code_units(cfg.str("ProfileCutoffPointOuter", ""), 'L');
	if (cfg.has("ProfileCutoffWidthOuter"))
	    d.cutoff_width_outer = U.in_code_units(cfg.str("ProfileCutoffWidthOuter", ""), 'L');
	if (cfg.has("ProfileCutoffPointInner"))
	    d.cutoff_point_inner = U.in_code_units(cfg.str("ProfileCutoffPointInner", ""), 'L');
	if (cfg.has("ProfileCutoffWidthInner"))
	    d.cutoff_width_inner = U.in_code_units(cfg.str("ProfileCutoffWidthInner", ""), 'L');
	d.spreading_ring = cfg.flag("SpreadingRing", false);
	d.set_sigma0 = cfg.flag("SetSigma0", false);
	d.diskmass = cfg.has("DiskMass") ? U.in_code_units(cfg.str("DiskMass", "0.01"), 'M') : 0.01;
	if (shock_tube) { // the tube replaces every other density / energy initialisation (init.cpp:269-272)
	    d.shock_tube = true;
	    d.sigma_in = d.energy_in = nullptr;
	    d.spreading_ring = d.cutoff_outer = d.cutoff_inner = d.set_sigma0 = d.cbd_ring = d.nbody_centered = d.energy_nbody_centered = false;
	}
	const finit::InitialState s0 = finit::init_gas(d, radii, nrad, naz, params.hydro_center_mass);
	params.sigma0 = d.sigma0; // SetSigma0 rescales it; the density floor follows (init.cpp:1155)
	// FARGO_BC_BALANCED: v_sq of balanced_boundary (boundary_conditions/balanced.cpp:23-52) for the two ghost rings
	for (int side = 0; side < 2; ++side) {
	    const int i = side == 0 ? 0 : nrad - 1;
	    const double ri = radii[i], rs = radii[i + 1];
	    double R = 2.0 / 3.0 * (std::pow(rs, 3) - std::pow(ri, 3)); // Rb (init.cpp:178-179)
	    R = R / (std::pow(rs, 2) - std::pow(ri, 2));
	    const double vk_2 = std::pow(std::sqrt(consts.G * params.hydro_center_mass / R), 2); // pow(compute_v_kepler(R, M), 2)
	    double support = 0.0;
	    if (!d.cutoff_outer) {
		support += finit::detail::support_azi_pressure(d, R);
		support += finit::detail::support_azi_smoothing_derivative(d, R);
	    }
	    if (d.quadrupole_support && d.quadrupole_moment > 0.0) // support_azi_quadrupole (Theo.cpp:150-157)
		support += 3.0 * d.quadrupole_moment / std::pow(R, 2);
	    params.balanced_vazi_sq[side] = vk_2 * support;
	}
	create_context(device);
	// init_euler (SourceEuler.cpp:250-285) runs BEFORE the velocities exist: Q+/- of the first CFL see a gas at rest
	CHECK(BK(upload_field)(ctx, FARGO_SIGMA, s0.sigma.data()));
	if (!s0.energy.empty()) // adiabatic runs; isothermal ones only carry (and write out) the energy ring of CircumBinaryRing
	    CHECK(BK(upload_field)(ctx, FARGO_ENERGY, s0.energy.data()));
	set_bodies_on_device();
	CHECK(BK(set_time)(ctx, time));
	if (params.pvte) { // init_eos_arrays (init.cpp:290-292, 1190-1206) with the constants as constants.cpp forms them
	    fargo_pvte_consts pk;
	    pk.xMF = cfg.num("HydrogenMassFraction", 0.75); // Interpret.cpp:452
	    pk.m_H = U.m_H.cgs, pk.m_e = U.m_e.cgs, pk.eV = U.eV.cgs, pk.h = U.h.cgs, pk.k_B = U.k_B.cgs;
	    pk.mp = 1.67262192369e-27 * 1.0 / 0.001; // llnl constants::mp in g (pvte_law.cpp:232)
	    CHECK(BK(set_pvte)(ctx, &pk));
	}
	CHECK(BK(init_derived)(ctx));
	// The reference's derived grids of snapshot 0 date from init_euler, i.e. from before the first boundary conditions changed
	// the ghost rings; this path evaluates derived fields on download, so the ones a setup asks for are taken now.
	for (auto &s : {std::make_pair((int)FARGO_TEMPERATURE, "WriteTemperature"), std::make_pair((int)FARGO_PRESSURE, "WritePressure"),
			std::make_pair((int)FARGO_SOUNDSPEED, "WriteSoundSpeed"), std::make_pair((int)FARGO_SCALE_HEIGHT, "WriteScaleHeight"),
			std::make_pair((int)FARGO_VISCOSITY, "WriteViscosity")})
	    if (cfg.flag(s.second, false)) {
		std::vector<double> buf(cells(false), 0.0);
		CHECK(BK(download_field)(ctx, s.first, buf.data()));
		derived_at_init[s.first] = buf;
	    }
	CHECK(BK(upload_field)(ctx, FARGO_VRAD, s0.vrad.data()));
	CHECK(BK(upload_field)(ctx, FARGO_VAZI, s0.vazi.data()));
	CHECK(BK(copy_initial_values)(ctx));	  // init.cpp:340-344
	CHECK(BK(stage_boundary)(ctx, 0.0, 0));
	CHECK(BK(copy_initial_values)(ctx));
	last_dt = cfg.num("FirstDT", 1e-9);
	calculate_time_step(); // main.cpp:117
	started_fresh = true;
	write_static_files();
	U.write_files(outdir);
	write_snapshot(); // sim::handle_outputs before sim::run (main.cpp:150)
	finish_pending_snapshot();
	set_bodies_on_device();
	write_planet_monitor_files();
	write_quantities();
	rank_barrier(); // every rank's rings of snapshot 0 are in the files
	if (params.damping || params.cooling_beta_reference == 1) { // the damping reference (simulation.cpp:42-47)
	    const std::string from = outdir + "/snapshots/0", to = outdir + "/snapshots/reference";
	    mkdirs(to);
	    for (const char *f : {"Sigma.dat", "vrad.dat", "vazi.dat", "energy.dat"})
		if (exists(from + "/" + f)) {
		    std::ifstream in(from + "/" + f, std::ios::binary);
		    std::ofstream o(to + "/" + f, std::ios::binary);
		    o << in.rdbuf();
		}
	}
    }

    // sim::CalculateTimeStep (simulation.cpp:100-118)
    double calculate_time_step()
    {
	double dt = 0.0;
	CHECK(BK(cfl)(ctx, &last_dt, &dt));
	return dt;
    }

    // refframe::init_corotation (frame_of_reference.cpp:19-28)
    bool corotating = false;
    unsigned corotation_body = 1;
    double corot_old_x = 0.0, corot_old_y = 0.0;
    void init_corotation()
    {
	if (corotating)
	    corot_old_x = bodies[corotation_body].rec.x, corot_old_y = bodies[corotation_body].rec.y;
    }
    void rotate_frame(double dt)
    { // refframe::handle_corotation (frame_of_reference.cpp:30-60): a corotating frame (Frame: C) first follows its reference
      // body — new OmegaFrame from the angle the body has moved, v_azi of the gas corrected by the change — then the
      // bodies are rotated into the frame, t_planetary_system::rotate (nbody/planetary_system.cpp:409-432)
	if (corotating) {
	    const double x = bodies[corotation_body].rec.x, y = bodies[corotation_body].rec.y;
	    const double distance_new = std::sqrt(std::pow(x, 2) + std::pow(y, 2));
	    const double distance_old = std::sqrt(std::pow(corot_old_x, 2) + std::pow(corot_old_y, 2));
	    const double cross = corot_old_x * y - x * corot_old_y;
	    const double OmegaNew = std::asin(cross / (distance_new * distance_old)) / dt;
	    const double domega = (OmegaNew - omega_frame);
	    CHECK(BK(correct_vazi)(ctx, domega));
	    omega_frame = OmegaNew;
	    // the gas step that follows reads the new OmegaFrame; the bodies it sees stay the ones of the last set_bodies
	    last_fb.omega_frame = omega_frame;
	    CHECK(BK(set_bodies)(ctx, &last_fb));
	}
	const double angle = omega_frame * dt;
	for (auto &b : bodies) {
	    const double x = b.rec.x, y = b.rec.y, vx = b.rec.vx, vy = b.rec.vy;
	    b.rec.x = x * std::cos(angle) + y * std::sin(angle);
	    b.rec.y = -x * std::sin(angle) + y * std::cos(angle);
	    b.rec.vx = vx * std::cos(angle) + vy * std::sin(angle);
	    b.rec.vy = -vx * std::sin(angle) + vy * std::cos(angle);
	}
	frame_angle += omega_frame * dt;
    }
    void apply_indirect_term_on_nbody(double dt)
    { // nbody/planetary_system.cpp:730-744
	for (auto &b : bodies) {
	    b.rec.vx = b.rec.vx + dt * ind_x;
	    b.rec.vy = b.rec.vy + dt * ind_y;
	}
    }
    void integrate_and_recentre(double dt)
    { // planetary_system.integrate + move_to_hydro_frame_center (simulation.cpp:222-224)
	nbody_integrate(bodies, consts.G, dt);
	if (bodies.size() > 1) { // move_to_hydro_frame_center (:750-768)
	    double cx = 0, cy = 0, cvx = 0, cvy = 0, cm = 0;
	    for (unsigned k = 0; k < n_center; ++k) {
		const PlanetRecord &q = bodies[k].rec;
		cm += q.mass;
		cx += q.x * q.mass, cy += q.y * q.mass, cvx += q.vx * q.mass, cvy += q.vy * q.mass;
	    }
	    if (cm > 0)
		cx = cx / cm, cy = cy / cm, cvx = cvx / cm, cvy = cvy / cm;
	    else
		cx = cy = cvx = cvy = 0.0;
	    for (auto &b : bodies) {
		b.rec.x -= cx, b.rec.y -= cy, b.rec.vx -= cvx, b.rec.vy -= cvy;
	    }
	}
	refresh_orbital_parameters();
    }

    // move_to_hydro_center_and_update_orbital_parameters (nbody/planetary_system.cpp:861-872): the distance to the primary
    // (compute_dist_to_primary :941-964) and the osculating elements (calculate_orbital_elements :773-805) follow the bodies
    // every step; the accretion rate and the mass ramp-up read the orbital period off them
    void refresh_orbital_parameters()
    {
	if (bodies.size() < 2)
	    return;
	for (size_t i = 1; i < bodies.size(); ++i) {
	    const double dx = bodies[i].rec.x - bodies[0].rec.x, dy = bodies[i].rec.y - bodies[0].rec.y;
	    const double dist = std::sqrt(std::pow(dx, 2) + std::pow(dy, 2));
	    bodies[i].rec.distance_to_primary = dist;
	    if (i == 1)
		bodies[0].rec.distance_to_primary = dist;
	}
	for (size_t i = (n_center == 1 ? 1 : 0); i < bodies.size(); ++i) {
	    double cx = 0, cy = 0, cvx = 0, cvy = 0, cm = 0;
	    for (size_t k = 0; k < i; ++k) {
		const PlanetRecord &q = bodies[k].rec;
		cx += q.x * q.mass, cy += q.y * q.mass, cvx += q.vx * q.mass, cvy += q.vy * q.mass;
		cm += q.mass;
	    }
	    if (cm > 0.0)
		cx /= cm, cy /= cm, cvx /= cm, cvy /= cm;
	    else
		cx = cy = cvx = cvy = 0.0;
	    finit::BodyInit e;
	    PlanetRecord &r = bodies[i].rec;
	    e.mass = r.mass;
	    finit::orbital_elements(e, r.x - cx, r.y - cy, r.vx - cvx, r.vy - cvy, cm, consts.G);
	    r.semi_major_axis = e.semi_major_axis, r.eccentricity = e.eccentricity, r.mean_anomaly = e.mean_anomaly;
	    r.true_anomaly = e.true_anomaly, r.eccentric_anomaly = e.eccentric_anomaly, r.pericenter_angle = e.pericenter_angle;
	    bodies[i].orbital_period = e.orbital_period;
	    bodies[i].omega = e.omega;
	}
	if (bodies.size() == 2) { // a binary: both carry the secondary's elements (:797-802)
	    PlanetRecord &p0 = bodies[0].rec;
	    const PlanetRecord &p1 = bodies[1].rec;
	    p0.semi_major_axis = p1.semi_major_axis, p0.eccentricity = p1.eccentricity, p0.mean_anomaly = p1.mean_anomaly;
	    p0.true_anomaly = p1.true_anomaly, p0.eccentric_anomaly = p1.eccentric_anomaly, p0.pericenter_angle = p1.pericenter_angle;
	    bodies[0].orbital_period = bodies[1].orbital_period;
	}
    }

    // accretion::AccreteOntoPlanets (accretion.cpp:419-452), first thing in a step (simulation.cpp:150-153, :302-303, :403-404):
    // bodies with an accretion efficiency take gas out of their Hill sphere: "accretion method: kley" (the default, two
    // zones, :84-221), "sinkhole" (one zone, :223-333) or "viscous" (:335-417).  A body that feels the disk (DiskFeedback,
    // or AccreteWithoutDiskFeedback) also gains the mass and momentum of that gas: update_planet (:60-82).
    void accrete(double dt)
    {
	bool masses_changed = false;
	for (size_t k = 0; k < bodies.size(); ++k) {
	    masses_changed = accrete_onto(k, dt) || masses_changed;
	    // AccreteOntoPlanets refreshes the Roche radii INSIDE its loop over the bodies (accretion.cpp:486-515): once one body
	    // has gained mass, after that body and after every later one — so a later body's Hill radius, and every cubic
	    // smoothing radius, sees one Newton step of update_l1 per remaining body, not one per step
	    if (masses_changed && bodies.size() > 1)
		update_roche_radii();
	}
    }
    // one body of AccreteOntoPlanets; true if it gained mass
    bool accrete_onto(size_t k, double dt)
    {
	bool masses_changed = false;
	{
	    Body &b = bodies[k];
	    if (!(b.rec.acc > 0.0) || !(b.orbital_period > 0.0))
		return false;
	    std::string method = "kley";
	    if (k < cfg.nbody.size() && cfg.nbody[k].count("accretion method"))
		method = lower(cfg.nbody[k].at("accretion method"));
	    if (method == "no" || method == "none")
		return false;
	    if (method != "kley" && method != "sinkhole" && method != "viscous")
		die("accretion method '%s' is not supported by this driver (kley, sinkhole, viscous)", method);
	    if (method == "viscous" && cfg.flag("ViscAccretMassflowTest", false))
		die("%s", "ViscAccretMassflowTest is not supported by this driver");
	    const double facc = method == "viscous" ? dt * 3.0 * M_PI * b.rec.acc // accretion.cpp:355
						    : dt * b.rec.acc / b.orbital_period * std::log(2);
	    const double r_hill = b.rec.dimensionless_roche_radius * b.rec.distance_to_primary;
	    const double frac = cfg.num("MassAccretionRadius", 1.0);
	    double taken[3];
	    if (method == "kley")
		CHECK(BK(accrete_kley)(ctx, b.rec.x, b.rec.y, r_hill, facc, frac, taken));
	    else if (method == "viscous")
		CHECK(BK(accrete_viscous)(ctx, b.rec.x, b.rec.y, r_hill, facc, frac, taken));
	    else
		CHECK(BK(accrete_sinkhole)(ctx, b.rec.x, b.rec.y, r_hill, facc, frac, taken));
	    b.rec.accreted_mass += taken[0]; // monitoring only (accretion.cpp:201)
	    if (disk_feedback || cfg.flag("AccreteWithoutDiskFeedback", false)) { // update_planet (accretion.cpp:60-82)
		double Mplanet = b.rec.mass;
		double PxPlanet = Mplanet * b.rec.vx, PyPlanet = Mplanet * b.rec.vy;
		Mplanet += taken[0];
		PxPlanet += taken[1];
		PyPlanet += taken[2];
		b.rec.accretion_torque_acc += (b.rec.x * taken[2] - b.rec.y * taken[1]);
		b.rec.vx = PxPlanet / Mplanet;
		b.rec.vy = PyPlanet / Mplanet;
		b.rec.mass = Mplanet;
		masses_changed = masses_changed || taken[0] > 0;
	    }
	}
	return masses_changed;
    }
    // update_global_hydro_frame_center_mass + update_roche_radii (accretion.cpp:510-514, planetary_system.cpp:1006-1033)
    void update_roche_radii()
    {
	{
	    params_hydro_center_mass_changed();
	    const double M = bodies[0].rec.mass;
	    for (size_t i = 1; i < bodies.size(); ++i) {
		const double m = bodies[i].rec.mass;
		double x = bodies[i].rec.dimensionless_roche_radius;
		if (M > m) {
		    x = finit::update_l1(M, m, x);
		    bodies[i].rec.dimensionless_roche_radius = x;
		} else { // sic (planetary_system.cpp:1025-1029): the planet keeps its value
		    x = 1.0 - x;
		    x = finit::update_l1(m, M, x);
		    x = 1.0 - x;
		}
		if (i == 1)
		    bodies[0].rec.dimensionless_roche_radius = 1.0 - x;
	    }
	}
    }
    // hydro_center_mass is the primary's mass (HydroFrameCenter: primary); only an accreting primary could change it, and the
    // device holds it as a context constant
    void params_hydro_center_mass_changed()
    {
	if (hydro_frame_center_mass() != params.hydro_center_mass)
	    die("%s", std::string("an accreting frame-centre body (a changing hydro-frame-centre mass) is not supported by this driver"));
    }

    // step_Euler (simulation.cpp:148-267) around the gas part
    void step_euler(double dt)
    {
	accrete(dt);
	disk_feedback_kick(dt);
	compute_indirect_nbody(dt); // :161
	set_bodies_on_device();	    // indirect term (:160-162), potential inputs (:170)
	apply_indirect_term_on_nbody(dt); // :164
	rotate_frame(dt);		  // :184
	CHECK(BK(set_time)(ctx, time));
	CHECK(BK(step)(ctx, dt)); // :187-218, :230-266
	integrate_and_recentre(dt);
	time += dt;
	n_iter++;
    }

    // step_LeapFrog (simulation.cpp:276-459): the bodies drift dt/2 first, kick dt/2, gas drift dt, kick dt/2 with the
    // bodies at mid-step, bodies drift dt/2
    void step_leapfrog(double dt)
    {
	const double frog = dt / 2, start_time = time, mid_time = time + frog;
	compute_indirect_nbody(frog);	  // :285-287, while the bodies are still at the start of the step
	init_corotation();		  // :289
	integrate_and_recentre(frog);	  // :290-292
	disk_feedback_compute(frog);	  // :294-297 ComputeDiskOnNbodyAccel, ComputeIndirectTermDisk — before the accretion
	combine_indirect();		  // :299
	accrete(frog);			  // :302-303
	disk_feedback_apply(frog);	  // :304-305 UpdatePlanetVelocitiesWithDiskForce
	apply_indirect_term_on_nbody(frog); // :306
	rotate_frame(frog);		  // :313
	set_bodies_on_device();		  // unlike step_Euler, the potential of the first kick sees the bodies AFTER the frame rotation (:317-322)
	CHECK(BK(set_time)(ctx, start_time));
	CHECK(BK(kick)(ctx, frog));  // :326-345
	CHECK(BK(drift)(ctx, dt));   // :347-352
	time = mid_time;	       // the ramp-up mass and the beta-cooling ramp see the mid-step time (:364-398)
	disk_feedback_compute(frog); // :352-355
	compute_indirect_nbody(frog); // :356, bodies at mid-step
	set_bodies_on_device();
	CHECK(BK(set_time)(ctx, mid_time));
	CHECK(BK(kick)(ctx, frog));
	accrete(frog);			  // :403-404, after the gas's second kick
	disk_feedback_apply(frog);	  // :406-408
	apply_indirect_term_on_nbody(frog); // :410
	init_corotation();		  // :413
	integrate_and_recentre(frog);	  // :414-416
	rotate_frame(frog);		  // :428
	time = start_time + dt;
	n_iter++;
	CHECK(BK(finish_step)(ctx, dt)); // :437-458
    }

    void step(double dt)
    {
	if (params.leapfrog)
	    step_leapfrog(dt);
	else
	    step_euler(dt);
    }

#ifndef FARGO_HOST_ORACLE
    // The four state fields of a snapshot leave the device asynchronously (fargo_snapshot_async) into page-locked buffers
    // and are written to disk by a writer thread while the time loop goes on; the reference's loop stands still during
    // write_full_output.  One snapshot in flight: the next one joins the writer first.
    double *snap_host[4] = {nullptr, nullptr, nullptr, nullptr};
    std::thread snap_writer;
    void finish_pending_snapshot()
    {
	if (snap_writer.joinable())
	    snap_writer.join();
    }
#else
    void finish_pending_snapshot() {}
#endif

    // output::write_full_output (output.cpp:249-330) for the files the parity tooling reads
    void write_snapshot()
    {
	const std::string sd = outdir + "/snapshots/" + std::to_string(n_snapshot);
	mkdirs(sd);
	const std::pair<int, const char *> state[4] = {{FARGO_SIGMA, "Sigma"}, {FARGO_VRAD, "vrad"}, {FARGO_VAZI, "vazi"}, {FARGO_ENERGY, "energy"}};
	// isothermal runs of the reference write their (all-zero) energy grid too unless WriteEnergy says no
	const bool write_energy = params.adiabatic || (started_fresh ? cfg.flag("WriteEnergy", true) : exists(refdir + "/snapshots/0/energy.dat"));
	const std::string rel = "snapshots/" + std::to_string(n_snapshot) + "/";
	mkdirs(fielddir + "/" + rel);
#ifndef FARGO_HOST_ORACLE
	finish_pending_snapshot();
	// page-locked buffers for the rings this rank writes; the ABI addresses host arrays by GLOBAL ring, so it is handed the
	// address global ring 0 would have
	const size_t own_off = (size_t)own_lo * naz;
	for (int k = 0; k < 4; ++k)
	    if (!snap_host[k]) {
		const size_t n = (size_t)(own_hi - own_lo + (k == 1 ? 1 : 0)) * naz;
		snap_host[k] = (double *)fargo_pinned_alloc(n * sizeof(double));
		if (!snap_host[k])
		    die("%s", std::string("fargo_pinned_alloc: ") + backend_error());
		std::fill(snap_host[k], snap_host[k] + n, 0.0);
	    }
	auto as_global = [own_off](double *own) { return (double *)((uintptr_t)own - own_off * sizeof(double)); };
	CHECK(fargo_snapshot_async(ctx, as_global(snap_host[0]), as_global(snap_host[1]), as_global(snap_host[2]),
				   write_energy ? as_global(snap_host[3]) : nullptr));
	{
	    backend_ctx *cx = ctx;
	    std::vector<std::tuple<std::string, const double *, bool>> jobs;
	    for (int k = 0; k < 4; ++k)
		if (k < 3 || write_energy)
		    jobs.push_back(std::make_tuple(rel + state[k].second + ".dat", (const double *)as_global(snap_host[k]), k == 1));
	    const Run *self = this;
	    snap_writer = std::thread([cx, jobs, self]() {
		if (fargo_snapshot_wait(cx) != 0)
		    die("%s", std::string("fargo_snapshot_wait: ") + backend_error());
		for (auto &j : jobs)
		    self->write_field_file(std::get<0>(j), std::get<1>(j), std::get<2>(j));
	    });
	}
#else
	for (auto &s : state) {
	    if (s.first == FARGO_ENERGY && !write_energy)
		continue;
	    std::vector<double> buf(cells(s.first == FARGO_VRAD), 0.0);
	    CHECK(BK(download_field)(ctx, s.first, buf.data()));
	    write_field_file(rel + s.second + ".dat", buf.data(), s.first == FARGO_VRAD);
	}
#endif
	if (params.adiabatic) {
	    for (auto &s : {std::make_pair((int)FARGO_QPLUS, "Qplus"), std::make_pair((int)FARGO_QMINUS, "Qminus")}) {
		std::vector<double> buf(cells(false), 0.0);
		CHECK(BK(download_field)(ctx, s.first, buf.data()));
		write_field_file(rel + s.second + ".dat", buf.data(), false);
	    }
	}
	// optional derived outputs (data.cpp: WriteTemperature, WritePressure, ...): evaluated from the current state on download,
	// which is what recalculate_derived_disk_quantities left in the reference's grids at the end of the step
	for (auto &s : {std::make_tuple((int)FARGO_TEMPERATURE, "WriteTemperature", "Temperature"),
			std::make_tuple((int)FARGO_PRESSURE, "WritePressure", "pressure"),
			std::make_tuple((int)FARGO_SOUNDSPEED, "WriteSoundSpeed", "soundspeed"),
			std::make_tuple((int)FARGO_SCALE_HEIGHT, "WriteScaleHeight", "scale_height"),
			std::make_tuple((int)FARGO_VISCOSITY, "WriteViscosity", "viscosity"),
			std::make_tuple((int)FARGO_GAMMAEFF, "WriteEffectiveGamma", "gammaeff"),
			std::make_tuple((int)FARGO_GAMMA1, "WriteFirstAdiabaticIndex", "gamma1"),
			std::make_tuple((int)FARGO_MU, "WriteMeanMolecularWeight", "mu")}) {
	    if (!cfg.flag(std::get<1>(s), false))
		continue;
	    std::vector<double> buf(cells(false), 0.0);
	    auto kept = derived_at_init.find(std::get<0>(s));
	    if (n_snapshot == 0 && started_fresh && kept != derived_at_init.end())
		buf = kept->second; // snapshot 0: as init_euler left them, before the first boundary conditions touched the ghost rings
	    else
		CHECK(BK(download_field)(ctx, std::get<0>(s), buf.data()));
	    write_field_file(rel + std::get<2>(s) + ".dat", buf.data(), false);
	}
	derived_at_init.clear();
	if (cfg.flag("WriteMassFlow", false)) {
	    // MASSFLOW (data.cpp:273-278): mass through the inner interface of every cell since the last snapshot, divided by the
	    // time between snapshots before it is written (quantities::calculate_massflow, quantities.cpp:771-781), as the 2-D file
	    // and as the azimuthally integrated 1-D file (t_polargrid::write1D, polargrid.cpp:187-282: radius, sum, min, max per
	    // interface ring), then cleared
	    std::vector<double> buf(cells(true), 0.0);
	    CHECK(BK(download_field)(ctx, FARGO_MASSFLOW, buf.data()));
	    const double denom = nmonitor * monitor_timestep;
	    const double inv_c = 1.0 / denom; // t_polargrid::operator/= multiplies by the reciprocal (polargrid.cpp:510-521)
	    const int hi = own_hi + ((rank == nranks - 1) ? 1 : 0);
	    for (size_t k = (size_t)own_lo * naz; k < (size_t)hi * naz; ++k)
		buf[k] *= inv_c;
	    write_field_file(rel + "MassFlow.dat", buf.data(), true);
	    std::vector<double> one((size_t)(hi - own_lo) * 4);
	    for (int i = own_lo; i < hi; ++i) {
		double sum = 0.0, mn = std::numeric_limits<double>::max(), mx = std::numeric_limits<double>::lowest();
		for (int j = 0; j < naz; ++j) {
		    const double v = buf[(size_t)i * naz + j];
		    sum += v;
		    mn = std::min(mn, v), mx = std::max(mx, v);
		}
		double *o = &one[(size_t)(i - own_lo) * 4];
		o[0] = radii[i], o[1] = sum, o[2] = mn, o[3] = mx; // vector grid: the interface radius Ra
	    }
	    const std::string path = fielddir + "/" + rel + "MassFlow1D.dat";
	    const int fd = open(path.c_str(), O_WRONLY | O_CREAT | (nranks == 1 ? O_TRUNC : 0), 0644);
	    if (fd < 0 || pwrite(fd, one.data(), one.size() * sizeof(double), (off_t)own_lo * 4 * sizeof(double)) != (ssize_t)(one.size() * sizeof(double)))
		die("cannot write %s", path);
	    close(fd);
	    CHECK(BK(clear_massflow)(ctx));
	}
	MiscEntry m;
	m.timestep = n_snapshot, m.nTimeStep = n_monitor, m.time = time, m.OmegaFrame = omega_frame, m.FrameAngle = frame_angle;
	m.last_dt = last_dt, m.N_iter = n_iter;
	FILE *f = fopen((sd + "/misc.bin").c_str(), "wb");
	if (!f || fwrite(&m, sizeof(m), 1, f) != 1)
	    die("cannot write %s/misc.bin", sd);
	fclose(f);
	for (size_t k = 0; k < bodies.size(); ++k) {
	    bodies[k].rec.timestep = n_snapshot;
	    FILE *g = fopen((sd + "/nbody" + std::to_string(k) + ".bin").c_str(), "wb");
	    if (!g || fwrite(&bodies[k].rec, sizeof(PlanetRecord), 1, g) != 1)
		die("cannot write nbody record in %s", sd);
	    fclose(g);
	}
	if (!config_path.empty()) { // the setup travels with every snapshot (output.cpp:190-200), so `restart` works on our own output
	    std::ifstream in(config_path, std::ios::binary);
	    std::ofstream o(sd + "/config.yml", std::ios::binary);
	    o << in.rdbuf();
	}
	{ // snapshots/list.txt (output.cpp:332-350)
	    std::ofstream l(outdir + "/snapshots/list.txt", std::ios::app);
	    l << n_snapshot << "\n";
	}
    }

    // output::write_quantities (output.cpp:326-493): one row of monitor/Quantities.dat per monitor step, file version 2.4 with
    // the 35 columns of quantities_file_column_v2_5 (output.cpp:39-75).  The global sums come from fargo_monitor_quantities
    // and fargo_monitor_disk (device reductions); the one column this path does not evaluate (the density-floor
    // mass creation) are written as nan, never as made-up numbers.
    bool quantities_header_written = false;
    void write_quantities()
    {
	if (!cfg.flag("WriteDiskQuantities", true))
	    return;
	const std::string path = outdir + "/monitor/Quantities.dat";
	FILE *fd = fopen(path.c_str(), quantities_header_written ? "a" : "w");
	if (!fd)
	    die("cannot write %s", path);
	if (!quantities_header_written) {
	    const double L = consts.length_cgs, M = consts.mass_cgs, T = consts.time_cgs;
	    auto desc = [](double v, const char *sym) {
		char b[96];
		snprintf(b, sizeof b, "%.16e %s", v, sym);
		return std::string(b);
	    };
	    const std::string mass = desc(M, "g"), time_ = desc(T, "s"), length = desc(L, "cm"), energy = desc(L * L * M / (T * T), "erg"),
			      angmom = desc(L * M * (L / T), "cm^2 g s^-1"), power = desc(M * L * L / (T * T * T), "erg/s"),
			      accel = desc(L / (T * T), "cm s^-2"), freq = desc(1.0 / T, "1/s"), torque = desc(L * L * M / (T * T), "erg"),
			      ppt = desc(M / (T * T) / T, "dyn/cm/s"), one = "1";
	    const std::pair<const char *, const std::string *> cols[35] = {
		{"snapshot number", &one}, {"monitor number", &one}, {"time", &time_}, {"mass", &mass}, {"radius", &length},
		{"angular momentum", &angmom}, {"total energy", &energy}, {"internal energy", &energy}, {"kinematic energy", &energy},
		{"potential energy", &energy}, {"radial kinetic energy", &energy}, {"azimuthal kinetic energy", &energy},
		{"eccentricity", &one}, {"periastron", &one}, {"viscous dissipation", &power}, {"luminosity", &power}, {"pdivv", &ppt},
		{"inner boundary mass inflow", &mass}, {"inner boundary mass outflow", &mass}, {"outer boundary mass inflow", &mass},
		{"outer boundary mass outflow", &mass}, {"wave damping inner mass creation", &mass},
		{"wave damping inner mass removal", &mass}, {"wave damping outer mass creation", &mass},
		{"wave damping outer mass removal", &mass}, {"density floor mass creation", &mass}, {"aspect ratio", &one},
		{"indirect term nbody x", &accel}, {"indirect term nbody y", &accel}, {"indirect term disk x", &accel},
		{"indirect term disk y", &accel}, {"frame angle", &freq}, {"advection torque", &torque}, {"viscous torque", &torque},
		{"gravitational torque", &torque}};
	    fprintf(fd, "#FargoCPT quantities file\n#version: 2.4\n");
	    for (int k = 0; k < 35; ++k)
		fprintf(fd, "#variable: %d | %s | %s\n", k, cols[k].first, cols[k].second->c_str());
	    quantities_header_written = true;
	}
	// parameters::quantities_radius_limit (parameters.cpp:516-522)
	double limit = cfg.has("QuantitiesRadiusLimit") ? cfg.num("QuantitiesRadiusLimit", 0.0) : 2.0 * params.rmax;
	if (limit <= params.rmin)
	    limit = 2.0 * params.rmax;
	double q[8];
	CHECK(BK(monitor_quantities)(ctx, limit, q));
	const double nan = std::nan("");
	double row[33];
	for (double &x : row)
	    x = nan;
	row[0] = time, row[1] = q[0], row[3] = q[1], row[5] = q[2], row[6] = q[3], row[8] = q[4], row[9] = q[5];
	row[12] = q[6], row[13] = q[7];
	{ // disk radius, eccentricity / periastron, aspect ratio (output.cpp:373-423; AspectRatioMode 0 is all make_params lets through)
	    double d[9];
	    CHECK(BK(monitor_disk)(ctx, limit, cfg.num("DiskRadiusMassFraction", 0.99), frame_angle, d));
	    row[30] = d[5], row[31] = d[6], row[32] = d[8]; // advection, viscous, gravitational torque (quantities.cpp:1000-1018)
	    // pdivv_total is only formed when the P_DIVV grid is written (SourceEuler.cpp:881-899); without WritepDV the reference
	    // prints the 0 it was initialised with (data.cpp:302)
	    if (cfg.flag("WritepDV", false))
		die("%s", std::string("WritepDV is not supported by this driver"));
	    row[14] = 0.0;
	    double bf[4]; // MassDelta.Inner / OuterBoundaryInflow / Outflow since the last row (output.cpp:438-445, reset :493)
	    CHECK(BK(boundary_flow)(ctx, bf, 1));
	    row[15] = bf[0], row[16] = bf[1], row[17] = bf[2], row[18] = bf[3];
	    double dm[4] = {0.0, 0.0, 0.0, 0.0}; // Inner / OuterWaveDampingMassCreation / Removal (output.cpp:446-453): 0 without damping
	    if (params.damping)
		CHECK(BK(damping_mass)(ctx, dm, 1));
	    row[19] = dm[0], row[20] = dm[1], row[21] = dm[2], row[22] = dm[3];
	    row[7] = d[7];			     // gravitationalEnergy (output.cpp:413-414)
	    row[4] = q[2] + q[3] + d[7];	     // totalEnergy = internalEnergy + kinematicEnergy + gravitationalEnergy (:416-417)
	    row[2] = d[0];
	    row[10] = std::sqrt(std::pow(d[1], 2) + std::pow(d[2], 2)); // calculate_disk_ecc_peri (quantities.cpp:552-567)
	    row[11] = std::atan2(d[2], d[1]);
	    row[24] = d[3];
	}
	row[25] = ind_nbody_x, row[26] = ind_nbody_y, row[27] = ind_disk_x, row[28] = ind_disk_y, row[29] = frame_angle;
	fprintf(fd, "%u\t%u", n_monitor / nmonitor, n_monitor); // N_snapshot = N_monitor / Nmonitor (simulation.cpp:52)
	for (double x : row)
	    fprintf(fd, "\t%#.16e", x);
	fprintf(fd, "\n");
	fclose(fd);
    }

    // t_planet::create_planet_file / write_ascii (nbody/planet.cpp:279-372) through t_planetary_system::write_planets(1)
    // (sim::handle_outputs, simulation.cpp:83-84): monitor/nbodyK.dat, file version 2, 22 columns.
    bool planet_files_created = false;
    void write_planet_monitor_files()
    {
	const double L = consts.length_cgs, M = consts.mass_cgs, T = consts.time_cgs;
	auto desc = [](double v, const char *sym) {
	    char b[96];
	    snprintf(b, sizeof b, "%.16e %s", v, sym);
	    return std::string(b);
	};
	const double div = cfg.flag("WriteAtEveryTimestep", true) ? monitor_timestep : monitor_timestep * nmonitor;
	// ComputeCircumPlanetaryMasses (circumplanetary_mass.cpp:11-51) right before write_planets (simulation.cpp:83-84): every
	// body but the first; the value also travels in the NEXT snapshot's binary record, like the reference's
	for (size_t k = 1; k < bodies.size(); ++k) {
	    PlanetRecord &r = bodies[k].rec;
	    CHECK(BK(circumplanetary_mass)(ctx, r.x, r.y, r.distance_to_primary * r.dimensionless_roche_radius, &r.circumplanetary_mass));
	}
	for (size_t k = 0; k < bodies.size(); ++k) {
	    const std::string path = outdir + "/monitor/nbody" + std::to_string(k) + ".dat";
	    FILE *fd = fopen(path.c_str(), planet_files_created ? "a" : "w");
	    if (!fd)
		die("cannot write %s", path);
	    PlanetRecord &r = bodies[k].rec;
	    if (!planet_files_created) {
		const std::string one = "1", length = desc(L, "cm"), velocity = desc(L / T, "cm s^-1"), mass = desc(M, "g"), time_ = desc(T, "s"),
				  freq = desc(1.0 / T, "1/s"), angmom = desc(L * M * (L / T), "cm^2 g s^-1"), torque = desc(L * L * M / (T * T), "erg"),
				  mdot = desc(M / T, "g s^-1");
		const std::pair<const char *, const std::string *> cols[22] = {
		    {"snapshot number", &one}, {"monitor number", &one}, {"x", &length}, {"y", &length}, {"vx", &velocity}, {"vy", &velocity},
		    {"mass", &mass}, {"time", &time_}, {"omega frame", &freq}, {"mdcp", &mass}, {"eccentricity", &one},
		    {"angular momentum", &angmom}, {"semi-major axis", &length}, {"omega kepler", &freq}, {"mean anomaly", &one},
		    {"eccentric anomaly", &one}, {"true anomaly", &one}, {"pericenter angle", &one}, {"gas torque", &torque},
		    {"accretion torque", &torque}, {"indirect torque", &torque}, {"accretion rate", &mdot}};
		std::string name = "planet" + std::to_string(k);
		if (k < cfg.nbody.size() && cfg.nbody[k].count("name"))
		    name = cfg.nbody[k].at("name");
		fprintf(fd, "#FargoCPT planet file for planet: %s\n#version: 2\n", name.c_str());
		for (int c = 0; c < 22; ++c)
		    fprintf(fd, "#variable: %d | %s | %s\n", c, cols[c].first, cols[c].second->c_str());
	    }
	    double torque;
	    if (disk_feedback) {
		torque = r.gas_torque_acc / div;
	    } else { // sim::handle_outputs refreshes the disk's pull for the monitor (simulation.cpp:58-61)
		double a4[4];
		CHECK(BK(disk_on_body_accel)(ctx, (int)k, r.cubic_smoothing_factor, a4));
		r.disk_on_planet_acceleration[0] = a4[0] + a4[2], r.disk_on_planet_acceleration[1] = a4[1] + a4[3];
		r.torque = (r.x * r.disk_on_planet_acceleration[1] - r.y * r.disk_on_planet_acceleration[0]) * r.mass;
		torque = r.torque;
	    }
	    const double angular_momentum = r.mass * r.x * r.vy - r.mass * r.y * r.vx;
	    const double row[20] = {r.x, r.y, r.vx, r.vy, r.mass, time, omega_frame, r.circumplanetary_mass, r.eccentricity, angular_momentum,
				    r.semi_major_axis, bodies[k].omega, r.mean_anomaly, r.eccentric_anomaly, r.true_anomaly, r.pericenter_angle,
				    torque, r.accretion_torque_acc / div, r.indirect_torque_acc / div, r.accreted_mass / div};
	    fprintf(fd, "%u\t%u", n_monitor / nmonitor, n_monitor);
	    for (double x : row)
		fprintf(fd, "\t%#.18g", x);
	    fprintf(fd, "\n");
	    fclose(fd);
	    // t_planet::write(1) resets the accumulators (planet.cpp:323-327)
	    r.accreted_mass = 0.0, r.gas_torque_acc = 0.0, r.accretion_torque_acc = 0.0, r.indirect_torque_acc = 0.0;
	}
	planet_files_created = true;
    }

    void write_static_files()
    { // dimensions.dat / used_rad.dat (init.cpp:227-247) so python_module/fargocpt/data.py can load the directory
	mkdirs(outdir + "/snapshots");
	mkdirs(outdir + "/monitor");
	FILE *f = fopen((outdir + "/dimensions.dat").c_str(), "w");
	fprintf(f, "#RMIN\tRMAX\tPHIMIN\tPHIMAX          \tNRAD\tNAZ\tNGHRAD\tNGHAZ\tRadial_spacing\n");
	fprintf(f, "%.16g\t%.16g\t%d\t%.16g\t%d\t%d\t%d\t%d\t%s\n", params.rmin, params.rmax, 0, 2.0 * M_PI, nrad, naz, 1, 1,
		cfg.str("RadialSpacing", "Arithmetic").c_str());
	fclose(f);
	f = fopen((outdir + "/used_rad.dat").c_str(), "w");
	for (double r : radii)
	    fprintf(f, "%.18g\n", r);
	fclose(f);
	f = fopen((outdir + "/monitor/timestepLogging.dat").c_str(), "w");
	fprintf(f, "#version: 1.1-b200\n#variable: 0 | snapshot number | 1\n#variable: 1 | monitor number | 1\n#variable: 2 | hydrostep number | 1\n"
		   "#variable: 3 | time | code\n#variable: 4 | hydro dt | code\n");
	fclose(f);
    }

    // sim::run (simulation.cpp:505-558) until `until_snapshot` has been written
    void run(unsigned until_snapshot, long max_steps)
    {
	if (!started_fresh)
	    write_static_files();
	// main.cpp:117 + sim::init (simulation.cpp:462-470): first CFL, boundaries, CFL again
	if (n_iter == 0 && !started_fresh) {
	    last_dt = cfg.num("FirstDT", 1e-9);
	    calculate_time_step();
	}
	CHECK(BK(stage_boundary)(ctx, 0.0, 0));
	init_corotation();     // sim::init (simulation.cpp:463-464)
	if (started_fresh || n_iter == 0) // sim::init (simulation.cpp:465-467): not when restarting — the growth limiter
	    calculate_time_step();	  // (CFLmaxVar * last_dt) must be applied once per step, not twice before the first one
	long steps = 0;
	FILE *tl = fopen((outdir + "/monitor/timestepLogging.dat").c_str(), "a");
	while (n_snapshot < until_snapshot) {
	    if (max_steps >= 0 && steps >= max_steps)
		break; // -N: stops WITHOUT writing a snapshot (simulation.cpp:517-519)
	    const double cfl_dt = calculate_time_step();
	    const double time_next_monitor = (n_monitor + 1) * monitor_timestep;
	    const double left = time_next_monitor - time;
	    const bool overshoot = cfl_dt > left, almost_there = left < cfl_dt * (1 + 0.05);
	    const double step_dt = (overshoot || almost_there) ? left : cfl_dt; // :528-540
	    // the "potential energy" / "gravitational torque" columns of Quantities.dat read the POTENTIAL grid of the last step's
	    // start: a step that ends on a monitor time keeps it (the fused kernels otherwise hold the potential in registers)
	    if (cfg.flag("WriteDiskQuantities", true))
		CHECK(BK(keep_potential)(ctx, (overshoot || almost_there) ? 1 : 0));
	    step(step_dt);
	    ++steps;
	    fprintf(tl, "%u\t%u\t%llu\t%.17g\t%.17g\n", n_snapshot, n_monitor, (unsigned long long)n_iter, time, step_dt);
	    if (std::fabs(time_next_monitor - time) < 1e-6 * cfl_dt) { // :544-550
		n_monitor++;
		const bool snapshot_now = n_monitor % nmonitor == 0;
		if (snapshot_now) {
		    n_snapshot = n_monitor / nmonitor;
		    write_snapshot();
		}
		if (snapshot_now || cfg.flag("WriteAtEveryTimestep", true)) { // sim::handle_outputs (simulation.cpp:50-98)
		    if (!disk_feedback)
			set_bodies_on_device(); // positions for the force integral of the torque monitor
		    write_planet_monitor_files();
		    write_quantities();
		}
	    }
	}
	fclose(tl);
	finish_pending_snapshot(); // the last snapshot's files are complete when run() returns
    }
};

int main(int argc, char **argv)
{
    // options.cpp:42-189 subset: `start <setup.yml>`, `restart N <dir>`, -N <steps>, plus --out / --until / --device
    std::string mode, dir, out;
    long nrestart = -1, max_steps = -1, until = -1;
    int device = 0, nranks = 1, rank = -1;
    for (int i = 1; i < argc; ++i) {
	const std::string a = argv[i];
	if (a == "start" && i + 1 < argc) {
	    mode = a;
	    dir = argv[++i];
	} else if (a == "restart" && i + 2 < argc) {
	    mode = a;
	    nrestart = atol(argv[++i]);
	    dir = argv[++i];
	} else if (a == "-N" && i + 1 < argc)
	    max_steps = atol(argv[++i]);
	else if (a == "--out" && i + 1 < argc)
	    out = argv[++i];
	else if (a == "--until" && i + 1 < argc)
	    until = atol(argv[++i]);
	else if (a == "--device" && i + 1 < argc)
	    device = atoi(argv[++i]);
	else if (a == "--ranks" && i + 1 < argc)
	    nranks = atoi(argv[++i]);
	else if (a == "--rank" && i + 1 < argc)
	    rank = atoi(argv[++i]);
	else
	    die("unknown argument %s", a);
    }
    if ((mode != "restart" && mode != "start") || out.empty()) {
	fprintf(stderr, "usage: fargocpt_b200 start <setup.yml> --out <output dir> [--until <snapshot>] [-N <steps>] [--device <id>] [--ranks <GPUs>]\n"
			"       fargocpt_b200 restart <N> <fargocpt output dir> --out <new output dir> [--until <snapshot>] [-N <steps>] [--device <id>]\n");
	return 2;
    }
#ifdef FARGO_HOST_ORACLE
    if (nranks != 1)
	die("%s", std::string("the oracle-bound test driver runs one rank"));
    rank = 0;
#else
    // one process per GPU: rank r computes on device `--device` + r.  Without --rank this process is rank 0 and forks the others
    // NOW, before anything has touched CUDA; they die with it (PDEATHSIG), and it dies when one of them fails (SIGCHLD).
    std::vector<pid_t> children;
    if (nranks < 1 || rank >= nranks)
	die("bad --ranks / --rank");
    if (nranks > 1 && rank < 0) {
	mkdirs(out);
	unlink((out + "/.nccl_id").c_str());
	rank = 0;
	static volatile sig_atomic_t reaping_ok = 0;
	(void)reaping_ok;
	struct sigaction sa;
	memset(&sa, 0, sizeof(sa));
	sa.sa_handler = [](int) {
	    int st = 0;
	    pid_t p;
	    while ((p = waitpid(-1, &st, WNOHANG)) > 0)
		if (!(WIFEXITED(st) && WEXITSTATUS(st) == 0)) {
		    static const char msg[] = "fargocpt_b200: a rank failed, stopping\n";
		    if (write(2, msg, sizeof(msg) - 1) < 0) {
		    }
		    _exit(1);
		}
	};
	sa.sa_flags = SA_NOCLDSTOP | SA_RESTART;
	sigaction(SIGCHLD, &sa, nullptr);
	for (int k = 1; k < nranks; ++k) {
	    const pid_t p = fork();
	    if (p < 0)
		die("fork failed");
	    if (p == 0) {
		prctl(PR_SET_PDEATHSIG, SIGKILL);
		signal(SIGCHLD, SIG_DFL);
		rank = k;
		children.clear();
		if (!freopen("/dev/null", "w", stdout)) {
		}
		break;
	    }
	    children.push_back(p);
	}
    }
    if (rank < 0)
	rank = 0;
    device += rank;
#endif
    Run r;
    r.rank = rank, r.nranks = nranks;
    r.fielddir = out;
    r.outdir = rank == 0 ? out : out + "/.rank" + std::to_string(rank);
    if (mode == "start")
	r.start(dir, device);
    else
	r.load(dir, (unsigned)nrestart, device);
    r.run(until >= 0 ? (unsigned)until : r.nsnapshots, max_steps);
    r.rank_barrier(); // all field files are complete
    printf("-- Final: Total Hydrosteps %llu, time %.17g, last snapshot %u\n", (unsigned long long)r.n_iter, r.time, r.n_snapshot);
#ifdef FARGO_HOST_ORACLE
    fargo_oracle_destroy(r.ctx);
#else
    fargo_ctx_destroy(r.ctx);
    if (r.nranks > 1) {
	// arrival files: everybody has left the barrier before the last one; the last one's files may still be looked at by a
	// slower rank, so they go when all ranks have exited (below; an external launcher's rank 0 leaves them behind)
	if (r.barrier_count >= 2)
	    unlink((out + "/.barrier" + std::to_string(r.barrier_count - 2) + "." + std::to_string(rank)).c_str());
	if (rank > 0) { // the scratch directory of small files nobody reads
	    const std::string cmd = "rm -rf '" + r.outdir + "'";
	    if (system(cmd.c_str()) != 0) {
	    }
	}
    }
    if (!children.empty()) {
	signal(SIGCHLD, SIG_DFL);
	int rc = 0;
	for (pid_t p : children) {
	    int st = 0;
	    if (waitpid(p, &st, 0) == p && !(WIFEXITED(st) && WEXITSTATUS(st) == 0))
		rc = 1;
	}
	for (int k = 0; k < r.nranks; ++k)
	    unlink((out + "/.barrier" + std::to_string(r.barrier_count - 1) + "." + std::to_string(k)).c_str());
	return rc;
    }
#endif
    return 0;
}
