/* fargo_b200.h — C ABI of the B200-native FargoCPT hydro step (libfargo_b200.so).
 *
 * This is the drop-in boundary for ONE hot path of rometsch/fargocpt: the gas part of
 * sim::step_Euler (src/simulation.cpp:148-267) plus sim::CalculateTimeStep
 * (src/simulation.cpp:100-118).  FargoCPT has no plugin registry; the seam is the set of free
 * functions step_Euler calls with a `t_data&`.  Each entry point below names the reference
 * function(s) (file:line, relative to the reference's src/) it replaces.  INTEGRATION.md
 * shows the reference-side binding a maintainer would add.
 *
 * Conventions (mirroring the reference):
 *  - plain C, plain pointers and sizes, no torch / CUDA types in any signature;
 *  - one opaque context per GPU (== per MPI rank of the reference: one radial slab);
 *  - the caller is single-threaded per context (reference: MPI_THREAD_FUNNELED);
 *  - fields are updated in place on the device; host copies only via upload/download;
 *  - every function returns 0 on success, non-zero on error; fargo_last_error() gives the text.
 *    (The reference die()s; the host driver above this ABI turns non-zero into exit.)
 *  - there is NO CPU fallback: without a CUDA device fargo_ctx_create fails.
 *
 * Field layout == t_polargrid (src/polargrid.h:111-114): row-major double, azimuth contiguous,
 * index = naz + nrad*Naz.  "Scalar" grids have nrad rings, v_rad has nrad+1 rings
 * (ring i = inner interface of cell i), v_azi lives on the azimuthal interface j-1/2.
 */
#ifndef FARGO_B200_H
#define FARGO_B200_H

#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define FARGO_ABI_VERSION 4
#define FARGO_MAX_BODIES 8
/* src/constants.h:17 (CPUOVERLAP) and :19 (GHOSTCELLS_B) */
#define FARGO_CPUOVERLAP 7
#define FARGO_GHOSTCELLS_B 1

/* t_data::t_polargrid_type subset that crosses the boundary (src/data.h:17-86) */
enum fargo_field {
    FARGO_SIGMA = 0,      /* Sigma.dat   [nrad][naz]   */
    FARGO_VRAD = 1,       /* vrad.dat    [nrad+1][naz] */
    FARGO_VAZI = 2,       /* vazi.dat    [nrad][naz]   */
    FARGO_ENERGY = 3,     /* energy.dat  [nrad][naz]   */
    FARGO_SIGMA0 = 4,     /* damping / beta-cooling reference values (data.h:40-43) */
    FARGO_VRAD0 = 5,
    FARGO_VAZI0 = 6,
    FARGO_ENERGY0 = 7,
    FARGO_QPLUS = 8,      /* stored Q+/alpha_r, read by the next CFL (SourceEuler.cpp:926) */
    FARGO_QMINUS = 9,
    FARGO_TEMPERATURE = 10, /* derived, computed on download */
    FARGO_PRESSURE = 11,
    FARGO_SOUNDSPEED = 12,
    FARGO_SCALE_HEIGHT = 13,
    FARGO_VISCOSITY = 14,
    FARGO_POTENTIAL = 15,
    FARGO_T_REYNOLDS = 16, /* T_Reynolds.dat: stress::calculate_Reynolds_stress (stress.cpp:34-70), computed on download */
    FARGO_GAMMAEFF = 17,   /* gammaeff.dat, mu.dat, gamma1.dat: the PVTE grids (data.cpp:36-47), stored when params.pvte */
    FARGO_MU = 18,
    FARGO_GAMMA1 = 19,
    FARGO_MASSFLOW = 20,   /* MassFlow.dat [nrad+1][naz]: mass through the inner interface of every cell, summed over the steps
			    * since the last fargo_clear_massflow (VanLeerRadial, TransportEuler.cpp:610-616); fargo_track_massflow */
    FARGO_NFIELDS = 21
};

enum fargo_artvisc { FARGO_ARTVISC_NONE = 0, FARGO_ARTVISC_TW = 1, FARGO_ARTVISC_SN = 2 };
enum fargo_limiter { FARGO_LIMITER_VANLEER = 0, FARGO_LIMITER_MC = 1 };
enum fargo_spacing { FARGO_SPACING_LOG = 0, FARGO_SPACING_ARITH = 1, FARGO_SPACING_EXP = 2, FARGO_SPACING_CUSTOM = 3 };
/* per-variable boundary functions (src/boundary_conditions/boundary_conditions.h:13-22) */
enum fargo_bc {
    FARGO_BC_NONE = 0,
    FARGO_BC_ZEROGRADIENT = 1, /* zero_gradient.cpp */
    FARGO_BC_OUTFLOW = 2,      /* outflow.cpp (v_rad only) */
    FARGO_BC_REFLECTING = 3,   /* reflecting.cpp (v_rad only) */
    FARGO_BC_KEPLERIAN = 4,    /* keplerian_azimuthal.cpp (v_azi only, the default) */
    FARGO_BC_REFERENCE = 5,    /* reference.cpp: copy X0 into the ghost rings */
    FARGO_BC_ZEROSHEAR = 6,    /* zero_shear.cpp (v_azi only): the ghost ring rotates at the angular velocity of the first active ring */
    FARGO_BC_BALANCED = 7,     /* balanced.cpp (v_azi only): the equilibrium rotation sqrt(balanced_vazi_sq) - Rb OmegaFrame */
    FARGO_BC_VISCOUS = 8       /* viscous.cpp:18-46 (v_rad, inner side only): v_rad = -1.5 ViscousOutflowSpeed / Ra x the mean viscosity of
				* rings 0 and 1 (Kley, Papaloizou & Ogilvie 2008).  Device: viscosities that do not depend on the state
				* (constant, or alpha in a locally isothermal disk); the outer variant reads past its grid in the
				* reference and is not offered.  FARGO_BC_KEPLERIAN on the INNER v_rad is keplerian_radial.cpp:18-39:
				* the two ghost interfaces move at keplerian_radial_factor x v_K(Rmed) */
};
/* damping.cpp t_damping_type */
enum fargo_damping { FARGO_DAMP_NONE = 0, FARGO_DAMP_INITIAL = 1, FARGO_DAMP_ZERO = 2, FARGO_DAMP_MEAN = 3 };
/* beta-cooling reference (SourceEuler.cpp:656-683), bit flags */
enum fargo_beta_ref { FARGO_BETA_REF_NONE = 0, FARGO_BETA_REF_REFERENCE = 1, FARGO_BETA_REF_MODEL = 2, FARGO_BETA_REF_FLOOR = 4 };

/* POD mirror of the parameters::* globals the hot path reads (src/parameters.h, Interpret.cpp).
 * All dimensional values are in CODE units (the host converts, like config::cfg.get<T>(key, default, unit)). */
typedef struct fargo_params {
    int abi_version;  /* FARGO_ABI_VERSION */
    /* mesh: global grid (Interpret.cpp:196-231) */
    int nrad;         /* GlobalNRadial: rings including one ghost ring per side */
    int naz;          /* NAzimuthal */
    int radial_spacing; /* enum fargo_spacing: only used for the damping-ring lookup (find_cell_id.cpp) */
    double rmin, rmax;  /* RMIN, RMAX */
    /* equation of state */
    int adiabatic;    /* 1: EquationOfState Ideal (energy equation); 0: locally isothermal */
    double gamma;     /* ADIABATICINDEX */
    double mu;        /* MU */
    double aspectratio_ref; /* AspectRatio */
    double flaring_index;   /* FlaringIndex */
    double sigma0;          /* Sigma0 (code units) */
    double sigma_floor;     /* SigmaFloor (multiples of sigma0) */
    double sigma_slope;     /* SigmaSlope (imposed disk drift only) */
    double minimum_temperature, maximum_temperature; /* code units */
    /* physical constants in code units (constants.cpp) */
    double G, Rgas, sigma_sb, c_light;
    double hydro_center_mass; /* global.cpp:146 */
    /* time step (cfl.cpp, simulation.cpp:100-118) */
    double cfl;               /* CFL */
    double cfl_max_var;       /* CFLmaxVar */
    double heating_cooling_cfl_limit; /* HeatingCoolingCFLlimit */
    int leapfrog;             /* Integrator: 0 Euler, 1 Leapfrog (affects cfl factor 0.6 and step order) */
    /* transport */
    int fast_transport;       /* Transport: FARGO (1) | STANDARD (0) */
    int flux_limiter;         /* enum fargo_limiter */
    /* artificial viscosity */
    int artificial_viscosity;             /* enum fargo_artvisc */
    double artificial_viscosity_factor;   /* ArtificialViscosityFactor */
    int artificial_viscosity_dissipation; /* ArtificialViscosityDissipation */
    /* physical viscosity */
    double viscous_alpha;        /* ViscousAlpha (>0 selects alpha viscosity, AlphaMode const) */
    double constant_viscosity;   /* ConstantViscosity (code units) */
    int stabilize_viscosity;     /* StabilizeViscosity 0|1|2 */
    double radial_viscosity_factor; /* RadialViscosityFactor */
    /* energy sources (SubStep3) */
    int heating_viscous;         /* HeatingViscous */
    double heating_viscous_factor;
    int cooling_beta;            /* CoolingBetaLocal */
    double cooling_beta_value;   /* CoolingBeta */
    double cooling_beta_ramp_up; /* CoolingBetaRampUp */
    int cooling_beta_reference;  /* enum fargo_beta_ref flags */
    /* gravity coupling (Pframeforce.cpp) */
    int body_force_from_potential; /* BodyForceFromPotential (1: potential, 0: accelerations) */
    double thickness_smoothing;    /* ThicknessSmoothing */
    double imposed_disk_drift;     /* ImposedDiskDrift */
    /* boundaries: [0]=inner, [1]=outer (boundary_conditions/config.cpp) */
    int bc_sigma[2], bc_energy[2], bc_vrad[2], bc_vazi[2]; /* enum fargo_bc */
    double keplerian_azimuthal_factor[2];
    /* damping zones (damping.cpp:185-271) */
    int damping;
    double damping_inner_limit, damping_outer_limit, damping_time_factor, damping_time_radius_outer;
    int damp_vrad[2], damp_vazi[2], damp_sigma[2], damp_energy[2]; /* enum fargo_damping */
    /* disk -> body force (Force.cpp:64-66): subtract the ring-mean density; CorrectDiskSelfgravity, default yes without self-gravity */
    int correct_disk_selfgravity;
    /* radiative surface cooling and stellar irradiation (SubStep3: SourceEuler.cpp:538-612 irradiation_single, :693-723
     * thermal_cooling; compute.cpp:17-88 midplane_density / kappa_eff; opacity.cpp:11-298).  heating_star is derived by the
     * host like t_planetary_system::derive_config (a body with temperature > 0 irradiates). */
    int cooling_surface;           /* SurfaceCooling: thermal */
    double surface_cooling_factor; /* CoolingRadiativeFactor */
    int heating_star;
    int opacity;                   /* enum fargo_opacity */
    double kappa_const;            /* KappaConst (code units) */
    double kappa_factor, tau_factor, tau_min, density_factor; /* KappaFactor, TauFactor, TauMin, DensityFactor */
    double temperature_cgs, density_cgs, opacity_code; /* units::temperature / density code -> cgs, units::opacity cgs -> code */
    /* EquationOfState: PVTE (pvte_law.cpp:371-568): gamma_eff, mu, Gamma_1 per cell from lookup tables in (rho, e) [cgs]; the
     * tables come through fargo_set_pvte_tables.  density_factor and density_cgs above are shared with the opacity. */
    int pvte;
    double energy_density_cgs, surface_density_cgs; /* units::energy_density / surface_density code -> cgs */
    /* viscosity::get_alpha (viscosity/viscosity.cpp:31-95): AlphaMode 0 constant ViscousAlpha, 1 S-curve in the (stored)
     * temperature after Ichikawa & Osaki (1992); AlphaCold, AlphaHot.  ViscousAlpha must be > 0 for either
     * (update_viscosity :102).  Modes 2 and 3 are refused. */
    int alpha_mode;
    double alpha_cold, alpha_hot;
    /* SurfaceCooling: scurve (scurve_cooling, SourceEuler.cpp:726-831; parameters.cpp:374-403): 0 off, 1 ScurveType Ichikawa
     * (Ichikawa & Osaki 1992), 2 Kimura (Kimura et al. 2020, the default).  The fit is written in cgs: code -> cgs factors of
     * length, mass and energy flux (units.cpp), the Stefan-Boltzmann and gravitational constants in cgs (constants.cpp). */
    int cooling_scurve;
    double length_cgs, mass_cgs, energy_flux_cgs, sigma_sb_cgs, G_cgs;
    /* FARGO_BC_BALANCED (boundary_conditions/balanced.cpp:23-75): v_K^2 x (pressure + smoothing (+ quadrupole) support) of the
     * inner / outer ghost ring, formed by the host from the disk model (Theo.cpp:122-160); the frame rotation is subtracted per call */
    double balanced_vazi_sq[2];
    double keplerian_radial_factor[2]; /* Inner / OuterBoundaryVradKeplerianFactor (config.cpp:220-255), default 0.1; [0] is used */
    double viscous_outflow_speed;      /* ViscousOutflowSpeed (config.cpp:498), default 1 */
} fargo_params;
/* parameters::t_opacity (parameters.h), Opacity: Lin | Bell | Constant | Simple */
enum fargo_opacity { FARGO_OPACITY_LIN = 0, FARGO_OPACITY_BELL = 1, FARGO_OPACITY_CONST = 2, FARGO_OPACITY_SIMPLE = 3 };

/* Star/planets as seen by the gas for ONE step (Pframeforce.cpp:27-36, refframe::IndirectTerm).
 * The N-body integration stays on the host (planetary_system.cpp); the host refreshes this every step. */
typedef struct fargo_bodies {
    int n;
    double x[FARGO_MAX_BODIES], y[FARGO_MAX_BODIES];
    double mass[FARGO_MAX_BODIES];                   /* planet.get_rampup_mass(t) */
    double cubic_smoothing_radius[FARGO_MAX_BODIES]; /* g_cubic_smoothing_radius, 0 = disabled */
    double indirect_x, indirect_y;                   /* refframe::IndirectTerm */
    double omega_frame;                              /* refframe::OmegaFrame */
    /* irradiation_single (SourceEuler.cpp:538-564): temperature > 0 marks an irradiating body; radius = planet_radial_extend;
     * ramp = 1 - cos^2(t pi / 2 / rampuptime) before the ramp-up time, else 1 (formed on the host) */
    double temperature[FARGO_MAX_BODIES], radius[FARGO_MAX_BODIES], irradiation_ramp[FARGO_MAX_BODIES];
} fargo_bodies;

/* What pvte::initializeLookupTables reads besides its own constants (pvte_law.cpp:443-495, 215-241): the hydrogen mass
 * fraction and physical constants in cgs units exactly as constants.cpp:48-85 forms them. */
typedef struct fargo_pvte_consts {
    double xMF;                  /* HydrogenMassFraction */
    double m_H, m_e, eV, h, k_B; /* constants::*.get_cgs_value() */
    double mp;                   /* llnl units proton mass in g */
} fargo_pvte_consts;

typedef struct fargo_ctx fargo_ctx;

/* --- life cycle ---------------------------------------------------------------------------
 * Replaces SplitDomain (split.cpp:21-87), init_radialarrays (init.cpp:78-249), data.set_size
 * (data.cpp:308), InitTransport (TransportEuler.cpp:57-96), cfl::init (cfl.cpp:14).
 * radii: the nrad+1 GLOBAL interface radii (used_rad.dat).  rank/nranks select the radial slab.
 * nccl_unique_id: 128-byte ncclUniqueId shared by all ranks (NULL when nranks == 1).
 * device: CUDA ordinal. */
int fargo_ctx_create(fargo_ctx **out, const fargo_params *params, const double *radii, int rank, int nranks,
		     const void *nccl_unique_id, int device);
void fargo_ctx_destroy(fargo_ctx *ctx);
const char *fargo_last_error(void);
/* fills a 128-byte buffer with a fresh ncclUniqueId (rank 0 calls this, then broadcasts it) */
int fargo_get_unique_id(void *out128);

/* slab geometry (split.cpp:38-78): local ring count, global index of local ring 0 */
int fargo_local_nrad(const fargo_ctx *ctx);
int fargo_local_imin(const fargo_ctx *ctx);

/* --- field transfer -----------------------------------------------------------------------
 * Mirrors t_polargrid::read2D / write2D slab semantics (polargrid.cpp:135-180, 301-353):
 * `host_global` is the GLOBAL array ([nrad(+1)][naz]); upload takes this rank's slab (overlap
 * rings included); download writes the rings this rank owns (Zero_or_active..Max_or_active),
 * leaving the rest of host_global untouched, so ranks can fill one global buffer. */
int fargo_upload_field(fargo_ctx *ctx, int field, const double *host_global);
int fargo_download_field(fargo_ctx *ctx, int field, double *host_global);
/* raw slab copy (all local rings, overlap included) — used by tests */
int fargo_download_slab(fargo_ctx *ctx, int field, double *host_slab);

/* copy_initial_values (damping.cpp:287-296): X0 <- X for the four state fields */
int fargo_copy_initial_values(fargo_ctx *ctx);

/* --- per-step inputs from the host ---------------------------------------------------------- */
int fargo_set_bodies(fargo_ctx *ctx, const fargo_bodies *bodies);
int fargo_set_time(fargo_ctx *ctx, double time); /* sim::time, used by the beta-cooling ramp */
/* init_eos_arrays (init.cpp:1190-1206) for params.pvte: builds the lookup tables of pvte::initializeLookupTables on the host
 * (host/fargo_pvte.h; a few seconds, cached per process), uploads them and fills the GAMMAEFF / GAMMA1 / MU grids with the
 * constant gamma / mu.  Call before fargo_init_derived. */
int fargo_set_pvte(fargo_ctx *ctx, const fargo_pvte_consts *consts);

/* --- the hot path ---------------------------------------------------------------------------
 * init_euler's derived fields (SourceEuler.cpp:251-285): T, c_s, H, P, nu, Q+/- for the first CFL */
int fargo_init_derived(fargo_ctx *ctx);
/* sim::CalculateTimeStep (simulation.cpp:100-118) = cfl::condition_cfl (cfl.cpp:185-382) +
 * MPI_Allreduce(MIN) + min(CFLmaxVar*last_dt, cfl).  `last_dt` is in/out (sim::last_dt). */
int fargo_cfl(fargo_ctx *ctx, double *last_dt, double *dt_out);
/* raw cfl::condition_cfl result (global min over ranks) */
int fargo_condition_cfl(fargo_ctx *ctx, double *cfl_dt_out);
/* gas part of step_Euler (simulation.cpp:167-175, 187-218, 230-266): potential, source terms,
 * artificial viscosity, viscosity, SubStep3, boundary, Transport, halo exchange, boundary+damping,
 * derived quantities.  Uses the bodies/time set before. */
int fargo_step(fargo_ctx *ctx, double dt);

/* The same gas update in three pieces, for hosts that arrange them like step_LeapFrog (simulation.cpp:276-459):
 *   Euler:    fargo_kick(dt); fargo_drift(dt); fargo_finish_step(dt)                         == fargo_step(dt)
 *   Leapfrog: fargo_kick(dt/2); fargo_drift(dt); <bodies/time at mid-step>; fargo_kick(dt/2); fargo_finish_step(dt)
 * kick  = CalculateNbodyPotential, update_with_sourceterms, artificial viscosity, viscosity, SubStep3 (:167-175, :187-203)
 * drift = apply_boundary_condition(final = false) + Transport (:213-215)
 * finish_step = CommunicateBoundaries + apply_boundary_condition(final = true, damping over dt) + derived (:230-266) */
int fargo_kick(fargo_ctx *ctx, double dt);
int fargo_drift(fargo_ctx *ctx, double dt);
int fargo_finish_step(fargo_ctx *ctx, double dt);

/* per-stage entry points (same order as fargo_step), exported so parity can be bisected per
 * reference function */
int fargo_stage_potential(fargo_ctx *ctx);             /* CalculateNbodyPotential  Pframeforce.cpp:21 */
int fargo_stage_sources(fargo_ctx *ctx, double dt);    /* update_with_sourceterms  SourceEuler.cpp:435 */
int fargo_stage_artvisc(fargo_ctx *ctx, double dt);    /* art_visc::update_with_artificial_viscosity artificial_viscosity.cpp:11 */
int fargo_stage_viscosity(fargo_ctx *ctx, double dt);  /* recalculate_viscosity + compute_viscous_stress_tensor + update_velocities_with_viscosity */
int fargo_stage_substep3(fargo_ctx *ctx, double dt);   /* SubStep3 SourceEuler.cpp:859 (adiabatic only) */
int fargo_stage_boundary(fargo_ctx *ctx, double dt, int final_call); /* apply_boundary_condition boundary_conditions.cpp:65 */
int fargo_stage_transport(fargo_ctx *ctx, double dt);  /* Transport TransportEuler.cpp:112 */
int fargo_stage_halo(fargo_ctx *ctx);                  /* CommunicateBoundaries commbound.cpp:98 */
/* Where the radial slabs are cut (SplitDomain, split.cpp:38-55): cut[r] = first ring owned by rank r, cut[nranks] = nrad.
 * The reference gives every rank the same number of rings; these cuts balance cost instead (rings inside a damping zone
 * weigh more), which changes no result (constants.h:17).  FARGO_B200_SPLIT=equal in the environment restores the
 * reference's cuts.  Pure host arithmetic: needs no device and no context. */
int fargo_split_cuts(const fargo_params *params, const double *radii, int nranks, int *cut /* nranks + 1 */);

/* how the ghost rings travel: 0 = single rank, 1 = ncclSend / ncclRecv after Transport, 2 = stored into the neighbour's
 * inbox over NVLink peer memory by the transport kernel's edge launch while the interior rings are still being
 * transported (default when CUDA IPC maps on every rank; FARGO_B200_HALO=nccl in the environment forces 1) */
int fargo_halo_mode(const fargo_ctx *ctx);
int fargo_stage_derived(fargo_ctx *ctx);               /* recalculate_derived_disk_quantities SourceEuler.cpp:225 */

/* fargo_step normally runs the fused source-term kernels; on != 0 makes it go through the per-stage kernels above
 * instead (same results; used by the parity tests to cross-check both) */
int fargo_set_staged(fargo_ctx *ctx, int on);

/* ComputeDiskOnPlanetAccel (Force.cpp:23-122): acceleration of body `body` (index into the bodies set with
 * fargo_set_bodies; its cubic smoothing radius is taken from there) by the gas of this rank's active rings, summed over
 * all ranks: out4 = {ax_inner, ay_inner, ax_outer, ay_outer} (parts from inside / outside the body's orbit; the
 * reference adds them, Force.cpp:117-119).  klahr_factor = planet.get_cubic_smoothing_factor(). */
int fargo_disk_on_body_accel(fargo_ctx *ctx, int body, double klahr_factor, double out4[4]);

/* Asynchronous snapshot of the four state fields (the output half of sim::handle_outputs simulation.cpp:50-98 /
 * output::write_full_output output.cpp:249, during which the reference's time loop stands still): returns at once; the
 * fields as of the call travel to the host arrays (global layout as fargo_download_field, owned rings only; page-locked
 * memory for the copy to overlap; energy may be NULL) while later fargo_step calls run.  fargo_snapshot_wait blocks until
 * they have arrived. */
int fargo_snapshot_async(fargo_ctx *ctx, double *sigma, double *vrad, double *vazi, double *energy);
int fargo_snapshot_wait(fargo_ctx *ctx);
/* page-locked host memory for those arrays, for hosts that do not link the CUDA runtime themselves (NULL on failure) */
void *fargo_pinned_alloc(size_t bytes);
void fargo_pinned_free(void *p);

/* accretion::AccreteOntoSinglePlanet (accretion.cpp:84-221; "accretion method: kley", called first thing in a step,
 * simulation.cpp:150-153): gas within frac * r_hill of the body at (x, y) loses the fraction facc / 3, within frac / 2 * r_hill
 * another 2 facc / 3, never below the density floor; Sigma (and the energy of an adiabatic disk) are changed in place.
 * The N-body side provides r_hill = dimensionless Roche radius * distance to the primary, facc = dt * accretion efficiency /
 * orbital period * ln 2 and frac = MassAccretionRadius.  out3 = {mass, x-momentum, y-momentum} taken from the active cells of all
 * ranks (what update_planet, accretion.cpp:60-82, adds to the body when it feels the disk). */
int fargo_accrete_kley(fargo_ctx *ctx, double x, double y, double r_hill, double facc, double frac, double out3[3]);
/* accretion::SinkHoleSinglePlanet (accretion.cpp:223-333; "accretion method: sinkhole"): one zone — gas within frac * r_hill
 * loses the fraction facc (never below the density floor).  Same inputs and outputs as fargo_accrete_kley. */
int fargo_accrete_sinkhole(fargo_ctx *ctx, double x, double y, double r_hill, double facc, double frac, double out3[3]);
/* accretion::AccreteOntoSinglePlanetViscous (accretion.cpp:335-417; "accretion method: viscous"): one zone of radius
 * d_max = frac * r_hill; a cell at distance d loses the fraction facc * nu(cell) * 3 / (pi d_max^2) * (1 - d / d_max) (never below
 * the density floor), nu being the viscosity of the state before this step's accretion (the reference reads the VISCOSITY grid
 * stored by the previous step).  The N-body side provides facc = dt * 3 pi * accretion efficiency (:355).  Same outputs. */
int fargo_accrete_viscous(fargo_ctx *ctx, double x, double y, double r_hill, double facc, double frac, double out3[3]);

/* ComputeCircumPlanetaryMasses (circumplanetary_mass.cpp:11-51; column "mdcp" of monitor/nbodyK.dat): the mass sum(Surf Sigma)
 * of the active cells whose centre is closer to (x, y) than roche_radius (= distance to the primary x dimensionless Roche
 * radius, from the N-body side), all ranks. */
int fargo_circumplanetary_mass(fargo_ctx *ctx, double x, double y, double roche_radius, double *out);

/* Global disk quantities of monitor/Quantities.dat (output::write_quantities output.cpp:326-520 -> quantities.cpp):
 * sums over the active cells with Rmed <= radius_limit (QuantitiesRadiusLimit, default 2 Rmax), all ranks.
 * out8 = { mass (quantities.cpp:51-78), angular momentum (:242-276), internal energy (:281-304), kinetic energy (:357-401),
 * radial kinetic energy (:406-438), azimuthal kinetic energy (:443-479), viscous dissipation sum(Surf Q+) (:306-328),
 * luminosity sum(Surf Q-) (:330-352) }.  The reference adds with an OpenMP reduction (order undefined); the device adds in a
 * fixed order (reproducible run to run; agrees with a serial sum to rounding). */
int fargo_monitor_quantities(fargo_ctx *ctx, double radius_limit, double out8[8]);

/* The mass-weighted columns of monitor/Quantities.dat (output.cpp:373-423), all ranks:
 * out9 = { disk radius (quantities::gas_disk_radius quantities.cpp:191-237: Rmed of the ring at which the running sum of the ring
 *          masses, the mesh's two ghost rings left out, first exceeds mass_fraction (DiskRadiusMassFraction, default 0.99) x the
 *          mass inside radius_limit),
 *          mass-weighted mean of the cells' eccentricity vector, x and y, rotated by frame_angle into the non-rotating frame
 *          (calculate_disk_ecc_vector :481-550, gas_reduce_mass_average :145-182; the caller forms the columns "eccentricity" =
 *          sqrt(x^2 + y^2) and "periastron" = atan2(y, x), :552-567),
 *          mass-weighted mean aspect ratio H / Rb (compute_aspectratio, AspectRatioMode 0, :784-806),
 *          the mass these means are weighted with,
 *          advection torque and viscous torque of the disk (gas_torques::calculate_advection_torque / calculate_viscous_torque,
 *          gas_torques.cpp:11-115, summed over the active cells inside radius_limit: quantities.cpp:80-105, 1000-1018),
 *          "potential energy" = -(mass-weighted mean of the POTENTIAL grid) (output.cpp:413-414) and the gravitational torque
 *          (gas_torques::calculate_gravitational_torque, gas_torques.cpp:122-153) — both read the POTENTIAL grid as the last
 *          kick stored it, like the reference (zeros before the first step); NaN without BodyForceFromPotential, and NaN when
 *          the last kick did not store the grid (see fargo_keep_potential) }.
 * Per-ring sums on the device in a fixed order, rings added in order on the host. */
int fargo_monitor_disk(fargo_ctx *ctx, double radius_limit, double mass_fraction, double frame_angle, double out9[9]);

/* WriteMassFlow (parameters.cpp:334-335): VanLeerRadial adds the mass that crosses the inner interface of every cell in a step to
 * the MASSFLOW grid (TransportEuler.cpp:610-616); the output divides it by the time between snapshots and clears it
 * (quantities::calculate_massflow quantities.cpp:771-781, data.cpp:273-278).  on != 0: the radial transport sweep accumulates the
 * grid (a separate instantiation of the kernel; 16 bytes per cell more traffic), readable as FARGO_MASSFLOW;
 * fargo_clear_massflow zeroes it. */
int fargo_track_massflow(fargo_ctx *ctx, int on);
int fargo_clear_massflow(fargo_ctx *ctx);

/* MassDelta.InnerBoundaryInflow / InnerBoundaryOutflow / OuterBoundaryInflow / OuterBoundaryOutflow (TransportEuler.cpp:578-608:
 * the mass VanLeerRadial moves through the inner interface of ring 1 and the outer interface of ring nrad - 2, split by direction;
 * columns 17-20 of monitor/Quantities.dat, which resets them after every row, output.cpp:493).  on != 0: the radial sweep
 * accumulates them per column (a separate instantiation of the kernel); fargo_boundary_flow returns the four sums over the
 * columns and over the ranks — out4 = { inner inflow, inner outflow, outer inflow, outer outflow } — and zeroes them if `reset`. */
int fargo_track_boundary_flow(fargo_ctx *ctx, int on);
int fargo_boundary_flow(fargo_ctx *ctx, double out4[4], int reset);

/* MassDelta.InnerWaveDampingMassCreation / Removal, OuterWaveDampingMassCreation / Removal (damping.cpp:335-357 and its inner /
 * outer, initial / zero / mean siblings: the mass (Xnew - X) Surf the damping of Sigma adds to or takes from the active cells of a
 * zone; columns 21-24 of monitor/Quantities.dat, reset after every row).  While tracked the zones of Sigma are damped in their own
 * pass.  fargo_damping_mass: out4 = { inner creation, inner removal, outer creation, outer removal }, all ranks; zeroes if `reset`. */
int fargo_track_damping_mass(fargo_ctx *ctx, int on);
int fargo_damping_mass(fargo_ctx *ctx, double out4[4], int reset);

/* CalculateNbodyPotential stores the POTENTIAL grid (Pframeforce.cpp:21-86); the fused source-term kernel keeps the potential in
 * registers.  on != 0: every following fargo_kick also stores the grid (one extra pass) for fargo_monitor_disk's potential
 * columns; a host switches it on for the step that ends on a monitor time.  The staged kernels always store it. */
int fargo_keep_potential(fargo_ctx *ctx, int on);

/* integer FARGO shifts of the last transport (TransportEuler.cpp:49,220), local rings */
int fargo_get_nshift(fargo_ctx *ctx, int *out_local_nrad);
/* correct_v_azimuthal (SideEuler.cpp:79-95), called by refframe::handle_corotation (frame_of_reference.cpp:30-60) when a
 * corotating frame (Frame: C) changes its angular velocity by domega: v_azi -= domega * Rmed in every ring, ghost rings
 * included.  The new OmegaFrame itself reaches the device through fargo_set_bodies. */
int fargo_correct_vazi(fargo_ctx *ctx, double domega);

/* device self-test of the branch-free IEEE arithmetic the kernels use (csrc/fargo_math.h) against the plain
 * operators on random + adversarial operands: counts4 = {division mismatches, sqrt mismatches, exp mismatches,
 * divisions that took the fast path}; 256 * blocks * per_thread operand pairs; wide != 0 spans the full exponent range */
int fargo_selftest_math(fargo_ctx *ctx, unsigned long long seed, int blocks, int per_thread, int wide,
			unsigned long long *counts4);

/* the ring sums behind the CFL dt and the FARGO shifts (cfl.cpp:199-204, TransportEuler.cpp:215-219: a strictly sequential
 * sum per ring) on nrows caller-provided rows of ns doubles: by the scan kernel (csrc/kernels_ringsum.cuh) and by a plain
 * one-thread-per-row chain; the caller compares both with its own sequential sum, bit for bit */
int fargo_selftest_ringsum(fargo_ctx *ctx, int nrows, int ns, const double *x_host, double *sums_scan, double *sums_chain);

/* the device exp of the energy equation (csrc/fargo_math.h: glibc's algorithm, operation by operation) on n host-provided
 * arguments; the caller compares with its libm */
int fargo_selftest_exp(fargo_ctx *ctx, int n, const double *x_host, double *y_host);

/* stream sync + launch accounting (for bench.py) */
int fargo_sync(fargo_ctx *ctx);
long long fargo_launch_count(const fargo_ctx *ctx);

/* per-kernel device timing (CUDA events on the context stream); report = "kernel ms count" lines */
int fargo_profile_enable(fargo_ctx *ctx, int on);
int fargo_profile_report(fargo_ctx *ctx, char *buf, int buflen);
/* device-side wall clock on the context's stream: 4 event slots */
int fargo_event_record(fargo_ctx *ctx, int slot);
int fargo_event_elapsed_ms(fargo_ctx *ctx, int slot_a, int slot_b, double *ms_out);

#ifdef __cplusplus
}
#endif
#endif /* FARGO_B200_H */
