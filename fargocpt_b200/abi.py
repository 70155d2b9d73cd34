"""ctypes mirror of include/fargo_b200.h (the C ABI of libfargo_b200.so).

Only plumbing lives here: struct layouts and a thin handle class.  The same handle class is used
by the tests to drive the CPU oracle (oracle/libfargo_oracle.so, prefix ``fargo_oracle_``), which
exports the same entry points — so parity tests call both sides through identical code.
"""
import ctypes as C
import os

import numpy as np

FARGO_ABI_VERSION = 4
FARGO_MAX_BODIES = 8
CPUOVERLAP = 7

# enum fargo_field
(SIGMA, VRAD, VAZI, ENERGY, SIGMA0, VRAD0, VAZI0, ENERGY0, QPLUS, QMINUS, TEMPERATURE, PRESSURE, SOUNDSPEED,
 SCALE_HEIGHT, VISCOSITY, POTENTIAL, T_REYNOLDS, GAMMAEFF, MU, GAMMA1, MASSFLOW) = range(21)
FIELD_NAMES = {SIGMA: "Sigma", VRAD: "vrad", VAZI: "vazi", ENERGY: "energy", QPLUS: "Qplus", QMINUS: "Qminus"}
VECTOR_FIELDS = (VRAD, VRAD0, MASSFLOW)  # grids on the radial interfaces: nrad + 1 rings

ARTVISC = {"none": 0, "tw": 1, "sn": 2}
LIMITER = {"vanleer": 0, "mc": 1}
SPACING = {"logarithmic": 0, "arithmetic": 1, "exponential": 2, "custom": 3}
BC = {"none": 0, "zerogradient": 1, "outflow": 2, "reflecting": 3, "keplerian": 4, "reference": 5, "zeroshear": 6, "balanced": 7, "viscous": 8}
DAMP = {"none": 0, "initial": 1, "reference": 1, "zero": 2, "mean": 3}
OPACITY = {"lin": 0, "bell": 1, "constant": 2, "simple": 3}  # parameters.cpp:414-428
BETA_REF = {"zero": 0, "reference": 1, "diskmodel": 2, "floor": 4}  # parameters.cpp:451-463


class FargoParams(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int),
        ("nrad", C.c_int), ("naz", C.c_int), ("radial_spacing", C.c_int),
        ("rmin", C.c_double), ("rmax", C.c_double),
        ("adiabatic", C.c_int),
        ("gamma", C.c_double), ("mu", C.c_double), ("aspectratio_ref", C.c_double), ("flaring_index", C.c_double),
        ("sigma0", C.c_double), ("sigma_floor", C.c_double), ("sigma_slope", C.c_double),
        ("minimum_temperature", C.c_double), ("maximum_temperature", C.c_double),
        ("G", C.c_double), ("Rgas", C.c_double), ("sigma_sb", C.c_double), ("c_light", C.c_double),
        ("hydro_center_mass", C.c_double),
        ("cfl", C.c_double), ("cfl_max_var", C.c_double), ("heating_cooling_cfl_limit", C.c_double),
        ("leapfrog", C.c_int),
        ("fast_transport", C.c_int), ("flux_limiter", C.c_int),
        ("artificial_viscosity", C.c_int), ("artificial_viscosity_factor", C.c_double),
        ("artificial_viscosity_dissipation", C.c_int),
        ("viscous_alpha", C.c_double), ("constant_viscosity", C.c_double), ("stabilize_viscosity", C.c_int),
        ("radial_viscosity_factor", C.c_double),
        ("heating_viscous", C.c_int), ("heating_viscous_factor", C.c_double),
        ("cooling_beta", C.c_int), ("cooling_beta_value", C.c_double), ("cooling_beta_ramp_up", C.c_double),
        ("cooling_beta_reference", C.c_int),
        ("body_force_from_potential", C.c_int), ("thickness_smoothing", C.c_double),
        ("imposed_disk_drift", C.c_double),
        ("bc_sigma", C.c_int * 2), ("bc_energy", C.c_int * 2), ("bc_vrad", C.c_int * 2), ("bc_vazi", C.c_int * 2),
        ("keplerian_azimuthal_factor", C.c_double * 2),
        ("damping", C.c_int),
        ("damping_inner_limit", C.c_double), ("damping_outer_limit", C.c_double),
        ("damping_time_factor", C.c_double), ("damping_time_radius_outer", C.c_double),
        ("damp_vrad", C.c_int * 2), ("damp_vazi", C.c_int * 2), ("damp_sigma", C.c_int * 2),
        ("damp_energy", C.c_int * 2),
        ("correct_disk_selfgravity", C.c_int),
        ("cooling_surface", C.c_int), ("surface_cooling_factor", C.c_double), ("heating_star", C.c_int), ("opacity", C.c_int),
        ("kappa_const", C.c_double), ("kappa_factor", C.c_double), ("tau_factor", C.c_double), ("tau_min", C.c_double),
        ("density_factor", C.c_double),
        ("temperature_cgs", C.c_double), ("density_cgs", C.c_double), ("opacity_code", C.c_double),
        ("pvte", C.c_int), ("energy_density_cgs", C.c_double), ("surface_density_cgs", C.c_double),
        ("alpha_mode", C.c_int), ("alpha_cold", C.c_double), ("alpha_hot", C.c_double),
        ("cooling_scurve", C.c_int), ("length_cgs", C.c_double), ("mass_cgs", C.c_double), ("energy_flux_cgs", C.c_double),
        ("sigma_sb_cgs", C.c_double), ("G_cgs", C.c_double),
        ("balanced_vazi_sq", C.c_double * 2),
        ("keplerian_radial_factor", C.c_double * 2), ("viscous_outflow_speed", C.c_double),
    ]

    def as_dict(self):
        out = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            out[name] = list(v) if hasattr(v, "__len__") else v
        return out

    @classmethod
    def from_dict(cls, d):
        p = cls()
        for name, typ in cls._fields_:
            if name not in d:
                continue
            v = d[name]
            if isinstance(v, (list, tuple)):
                for k, x in enumerate(v):
                    getattr(p, name)[k] = x
            else:
                setattr(p, name, v)
        if "correct_disk_selfgravity" not in d:
            p.correct_disk_selfgravity = 1  # parameters.cpp:699 default without self-gravity
        p.abi_version = FARGO_ABI_VERSION
        return p


class PvteConsts(C.Structure):
    """fargo_pvte_consts: the cgs constants exactly as constants.cpp:48-85 forms them (same IEEE operations)."""
    _fields_ = [("xMF", C.c_double), ("m_H", C.c_double), ("m_e", C.c_double), ("eV", C.c_double), ("h", C.c_double),
                ("k_B", C.c_double), ("mp", C.c_double)]

    @classmethod
    def make(cls, hydrogen_mass_fraction=0.75):
        cm, g, s, K = 0.01, 0.001, 1.0, 1.0
        k = cls()
        k.xMF = hydrogen_mass_fraction
        m_u = 1.66053906660e-27 * 1.0 / g
        k.m_H = 1.007825 * m_u
        k.m_e = 9.1093837015e-31 * 1.0 / g
        k.eV = 1.0e7 * 1.602176634e-19
        k.h = 6.62607015e-34 * 1.0 / (g * cm * cm / s)
        k.k_B = 1.380649e-23 * 1.0 / (g * cm * cm / (K * s * s))
        k.mp = 1.67262192369e-27 * 1.0 / g
        return k


class FargoBodies(C.Structure):
    _fields_ = [
        ("n", C.c_int),
        ("x", C.c_double * FARGO_MAX_BODIES), ("y", C.c_double * FARGO_MAX_BODIES),
        ("mass", C.c_double * FARGO_MAX_BODIES), ("cubic_smoothing_radius", C.c_double * FARGO_MAX_BODIES),
        ("indirect_x", C.c_double), ("indirect_y", C.c_double), ("omega_frame", C.c_double),
        ("temperature", C.c_double * FARGO_MAX_BODIES), ("radius", C.c_double * FARGO_MAX_BODIES),
        ("irradiation_ramp", C.c_double * FARGO_MAX_BODIES),
    ]

    @classmethod
    def make(cls, x, y, mass, rsm=None, indirect=(0.0, 0.0), omega_frame=0.0, temperature=None, radius=None, ramp=None):
        b = cls()
        b.n = len(x)
        for k in range(b.n):
            b.x[k], b.y[k], b.mass[k] = x[k], y[k], mass[k]
            b.cubic_smoothing_radius[k] = 0.0 if rsm is None else rsm[k]
            b.temperature[k] = 0.0 if temperature is None else temperature[k]
            b.radius[k] = 0.0 if radius is None else radius[k]
            b.irradiation_ramp[k] = 1.0 if ramp is None else ramp[k]
        b.indirect_x, b.indirect_y = indirect
        b.omega_frame = omega_frame
        return b


_DP = C.POINTER(C.c_double)


def _dptr(a):
    return a.ctypes.data_as(_DP)


def _bind(lib, prefix):
    """Declare argtypes for every entry point of include/fargo_b200.h on `lib`."""
    vp = C.c_void_p
    sig = {
        "local_nrad": ([vp], C.c_int), "local_imin": ([vp], C.c_int),
        "upload_field": ([vp, C.c_int, _DP], C.c_int), "download_field": ([vp, C.c_int, _DP], C.c_int),
        "download_slab": ([vp, C.c_int, _DP], C.c_int),
        "copy_initial_values": ([vp], C.c_int),
        "set_bodies": ([vp, C.POINTER(FargoBodies)], C.c_int), "set_time": ([vp, C.c_double], C.c_int),
        "init_derived": ([vp], C.c_int), "set_pvte": ([vp, C.POINTER(PvteConsts)], C.c_int),
        "cfl": ([vp, _DP, _DP], C.c_int), "condition_cfl": ([vp, _DP], C.c_int),
        "step": ([vp, C.c_double], C.c_int),
        "kick": ([vp, C.c_double], C.c_int), "drift": ([vp, C.c_double], C.c_int), "finish_step": ([vp, C.c_double], C.c_int),
        "stage_potential": ([vp], C.c_int), "stage_sources": ([vp, C.c_double], C.c_int),
        "stage_artvisc": ([vp, C.c_double], C.c_int), "stage_viscosity": ([vp, C.c_double], C.c_int),
        "stage_substep3": ([vp, C.c_double], C.c_int),
        "stage_boundary": ([vp, C.c_double, C.c_int], C.c_int),
        "stage_transport": ([vp, C.c_double], C.c_int), "stage_derived": ([vp], C.c_int),
        "get_nshift": ([vp, C.POINTER(C.c_int)], C.c_int),
        "correct_vazi": ([vp, C.c_double], C.c_int),
    }
    for name, (args, res) in sig.items():
        fn = getattr(lib, prefix + name)
        fn.argtypes, fn.restype = args, res
    return lib


class Handle:
    """One hydro context (== one radial slab == one GPU / one MPI rank of the reference)."""

    def __init__(self, lib, prefix, ptr, params, rank, nranks):
        self.lib, self.prefix, self.ptr = lib, prefix, C.c_void_p(ptr)
        self.params, self.rank, self.nranks = params, rank, nranks
        self.nrad_global, self.naz = params.nrad, params.naz
        self.nr = self._call("local_nrad")
        self.imin = self._call("local_imin")

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def _call(self, name, *args):
        return self._fn(name)(self.ptr, *args)

    def _check(self, rc, name):
        if rc != 0:
            msg = ""
            if hasattr(self.lib, "fargo_last_error") and self.prefix == "fargo_":
                self.lib.fargo_last_error.restype = C.c_char_p
                msg = (self.lib.fargo_last_error() or b"").decode()
            raise RuntimeError(f"{self.prefix}{name} failed rc={rc}: {msg}")

    def global_shape(self, field):
        return (self.nrad_global + (1 if field in VECTOR_FIELDS else 0), self.naz)

    def slab_shape(self, field):
        return (self.nr + (1 if field in VECTOR_FIELDS else 0), self.naz)

    def upload(self, field, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        assert a.shape == self.global_shape(field), (a.shape, self.global_shape(field))
        self._check(self._call("upload_field", field, _dptr(a)), "upload_field")

    def download(self, field, out=None):
        if out is None:
            out = np.zeros(self.global_shape(field), dtype=np.float64)
        self._check(self._call("download_field", field, _dptr(out)), "download_field")
        return out

    def download_slab(self, field):
        out = np.zeros(self.slab_shape(field), dtype=np.float64)
        self._check(self._call("download_slab", field, _dptr(out)), "download_slab")
        return out

    def copy_initial_values(self):
        self._check(self._call("copy_initial_values"), "copy_initial_values")

    def set_bodies(self, bodies):
        self._bodies = bodies
        self._check(self._call("set_bodies", C.byref(bodies)), "set_bodies")

    def set_time(self, t):
        self._check(self._call("set_time", float(t)), "set_time")

    def set_pvte(self, hydrogen_mass_fraction=0.75):
        k = PvteConsts.make(hydrogen_mass_fraction)
        self._check(self._call("set_pvte", C.byref(k)), "set_pvte")

    def init_derived(self):
        self._check(self._call("init_derived"), "init_derived")

    def cfl(self, last_dt):
        l, d = C.c_double(last_dt), C.c_double(0.0)
        self._check(self._call("cfl", C.byref(l), C.byref(d)), "cfl")
        return d.value

    def condition_cfl(self):
        d = C.c_double(0.0)
        self._check(self._call("condition_cfl", C.byref(d)), "condition_cfl")
        return d.value

    def step(self, dt):
        self._check(self._call("step", float(dt)), "step")

    def kick(self, dt):
        self._check(self._call("kick", float(dt)), "kick")

    def drift(self, dt):
        self._check(self._call("drift", float(dt)), "drift")

    def finish_step(self, dt):
        self._check(self._call("finish_step", float(dt)), "finish_step")

    def step_leapfrog(self, time, dt, bodies_mid=None):
        """step_LeapFrog's gas part (simulation.cpp:276-459): kick dt/2, drift dt, (bodies at mid-step), kick dt/2."""
        frog = dt / 2
        self.set_time(time)
        self.kick(frog)
        self.drift(dt)
        if bodies_mid is not None:
            self.set_bodies(bodies_mid)
        self.set_time(time + frog)
        self.kick(frog)
        self.finish_step(dt)

    def stage(self, name, *args):
        self._check(self._call("stage_" + name, *args), "stage_" + name)

    def disk_on_body_accel(self, body, klahr_factor=0.0):
        """{ax_inner, ay_inner, ax_outer, ay_outer} of ComputeDiskOnPlanetAccel (Force.cpp:23-122)."""
        out = (C.c_double * 4)()
        fn = self._fn("disk_on_body_accel")
        fn.argtypes = [C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_double)]
        fn.restype = C.c_int
        self._check(fn(self.ptr, int(body), float(klahr_factor), out), "disk_on_body_accel")
        return np.array(list(out))

    def accrete_kley(self, x, y, r_hill, facc, frac=1.0, method="kley"):
        """accretion::AccreteOntoSinglePlanet (accretion.cpp:84-221) or, with method="sinkhole", SinkHoleSinglePlanet
        (:223-333); returns (dM, dPx, dPy) taken from the active cells."""
        out = (C.c_double * 3)()
        fn = self._fn("accrete_" + method)
        fn.argtypes = [C.c_void_p] + [C.c_double] * 5 + [C.POINTER(C.c_double)]
        fn.restype = C.c_int
        self._check(fn(self.ptr, float(x), float(y), float(r_hill), float(facc), float(frac), out), "accrete_kley")
        return tuple(out)

    def monitor_quantities(self, radius_limit=1e300):
        """Global sums of monitor/Quantities.dat (fargo_monitor_quantities): dict of mass, angular_momentum, internal_energy,
        kinetic_energy, radial_kinetic_energy, azimuthal_kinetic_energy, viscous_dissipation, luminosity."""
        out = (C.c_double * 8)()
        fn = self._fn("monitor_quantities")
        fn.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double)]
        fn.restype = C.c_int
        self._check(fn(self.ptr, float(radius_limit), out), "monitor_quantities")
        return dict(zip(MONITOR_QUANTITIES, list(out)))

    def circumplanetary_mass(self, x, y, roche_radius):
        """ComputeCircumPlanetaryMasses (circumplanetary_mass.cpp:11-51): mass inside the Roche radius around (x, y)."""
        out = C.c_double(0.0)
        fn = self._fn("circumplanetary_mass")
        fn.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double)]
        fn.restype = C.c_int
        self._check(fn(self.ptr, float(x), float(y), float(roche_radius), C.byref(out)), "circumplanetary_mass")
        return out.value

    def monitor_disk(self, radius_limit=1e300, mass_fraction=0.99, frame_angle=0.0):
        """The mass-weighted columns of monitor/Quantities.dat (fargo_monitor_disk): dict of radius, eccentricity, periastron,
        aspect_ratio, advection_torque, viscous_torque (and the raw ecc_x, ecc_y, mass the first three are formed from, output.cpp:373-423 / quantities.cpp:552-567)."""
        import math
        out = (C.c_double * 9)()
        fn = self._fn("monitor_disk")
        fn.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double)]
        fn.restype = C.c_int
        self._check(fn(self.ptr, float(radius_limit), float(mass_fraction), float(frame_angle), out), "monitor_disk")
        r, ex, ey, h, m, tadv, tvisc, epot, tgrav = list(out)
        return {"radius": r, "eccentricity": math.sqrt(ex ** 2 + ey ** 2), "periastron": math.atan2(ey, ex), "aspect_ratio": h,
                "ecc_x": ex, "ecc_y": ey, "mass": m, "advection_torque": tadv, "viscous_torque": tvisc,
                "potential_energy": epot, "gravitational_torque": tgrav}

    def track_massflow(self, on=True):
        """fargo_track_massflow: the radial sweep accumulates the MASSFLOW grid (download(abi.MASSFLOW), clear_massflow())."""
        self._check(self._call("track_massflow", int(bool(on))), "track_massflow")

    def clear_massflow(self):
        self._check(self._call("clear_massflow"), "clear_massflow")

    def track_boundary_flow(self, on=True):
        self._check(self._call("track_boundary_flow", int(bool(on))), "track_boundary_flow")

    def boundary_flow(self, reset=True):
        """fargo_boundary_flow: (inner inflow, inner outflow, outer inflow, outer outflow) since the last reset."""
        out = (C.c_double * 4)()
        fn = self._fn("boundary_flow")
        fn.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
        fn.restype = C.c_int
        self._check(fn(self.ptr, out, int(bool(reset))), "boundary_flow")
        return tuple(out)

    def track_damping_mass(self, on=True):
        self._check(self._call("track_damping_mass", int(bool(on))), "track_damping_mass")

    def damping_mass(self, reset=True):
        """fargo_damping_mass: (inner creation, inner removal, outer creation, outer removal) since the last reset."""
        out = (C.c_double * 4)()
        fn = self._fn("damping_mass")
        fn.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
        fn.restype = C.c_int
        self._check(fn(self.ptr, out, int(bool(reset))), "damping_mass")
        return tuple(out)

    def keep_potential(self, on=True):
        """fargo_keep_potential: the following kicks also store the POTENTIAL grid (for monitor_disk's potential columns)."""
        self._check(self._call("keep_potential", int(bool(on))), "keep_potential")

    def correct_vazi(self, domega):
        """correct_v_azimuthal (SideEuler.cpp:79-95): a corotating frame changed its angular velocity by domega."""
        self._check(self._call("correct_vazi", float(domega)), "correct_vazi")

    def nshift(self):
        out = np.zeros(self.nr, dtype=np.int32)
        self._check(self._call("get_nshift", out.ctypes.data_as(C.POINTER(C.c_int))), "get_nshift")
        return out


_HERE = os.path.dirname(os.path.abspath(__file__))
MONITOR_QUANTITIES = ("mass", "angular_momentum", "internal_energy", "kinetic_energy", "radial_kinetic_energy",
                      "azimuthal_kinetic_energy", "viscous_dissipation", "luminosity")
# FARGO_B200_LIB: experiment builds of the SAME CUDA library (kernel tuning variants); never a CPU path
LIB_PATH = os.environ.get("FARGO_B200_LIB") or os.path.join(_HERE, "csrc", "libfargo_b200.so")
_lib = None


def load_library():
    """Load the CUDA library.  Fails loudly if it was not built — there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        _bind(lib, "fargo_")
        lib.fargo_ctx_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(FargoParams), _DP, C.c_int, C.c_int,
                                         C.c_void_p, C.c_int]
        lib.fargo_ctx_create.restype = C.c_int
        lib.fargo_ctx_destroy.argtypes = [C.c_void_p]
        lib.fargo_ctx_destroy.restype = None
        lib.fargo_last_error.restype = C.c_char_p
        lib.fargo_get_unique_id.argtypes = [C.c_void_p]
        lib.fargo_get_unique_id.restype = C.c_int
        lib.fargo_stage_halo.argtypes = [C.c_void_p]
        lib.fargo_stage_halo.restype = C.c_int
        lib.fargo_set_staged.argtypes = [C.c_void_p, C.c_int]
        lib.fargo_set_staged.restype = C.c_int
        lib.fargo_selftest_math.argtypes = [C.c_void_p, C.c_ulonglong, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_ulonglong)]
        lib.fargo_selftest_math.restype = C.c_int
        lib.fargo_selftest_ringsum.argtypes = [C.c_void_p, C.c_int, C.c_int, _DP, _DP, _DP]
        lib.fargo_selftest_ringsum.restype = C.c_int
        lib.fargo_selftest_exp.argtypes = [C.c_void_p, C.c_int, _DP, _DP]
        lib.fargo_selftest_exp.restype = C.c_int
        lib.fargo_sync.argtypes = [C.c_void_p]
        lib.fargo_sync.restype = C.c_int
        lib.fargo_snapshot_async.argtypes = [C.c_void_p, _DP, _DP, _DP, _DP]
        lib.fargo_snapshot_async.restype = C.c_int
        lib.fargo_snapshot_wait.argtypes = [C.c_void_p]
        lib.fargo_snapshot_wait.restype = C.c_int
        lib.fargo_monitor_quantities.argtypes = [C.c_void_p, C.c_double, _DP]
        lib.fargo_monitor_quantities.restype = C.c_int
        lib.fargo_circumplanetary_mass.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, _DP]
        lib.fargo_circumplanetary_mass.restype = C.c_int
        lib.fargo_track_massflow.argtypes = [C.c_void_p, C.c_int]
        lib.fargo_track_massflow.restype = C.c_int
        lib.fargo_clear_massflow.argtypes = [C.c_void_p]
        lib.fargo_clear_massflow.restype = C.c_int
        lib.fargo_track_boundary_flow.argtypes = [C.c_void_p, C.c_int]
        lib.fargo_track_boundary_flow.restype = C.c_int
        lib.fargo_boundary_flow.argtypes = [C.c_void_p, _DP, C.c_int]
        lib.fargo_boundary_flow.restype = C.c_int
        lib.fargo_track_damping_mass.argtypes = [C.c_void_p, C.c_int]
        lib.fargo_track_damping_mass.restype = C.c_int
        lib.fargo_damping_mass.argtypes = [C.c_void_p, _DP, C.c_int]
        lib.fargo_damping_mass.restype = C.c_int
        lib.fargo_keep_potential.argtypes = [C.c_void_p, C.c_int]
        lib.fargo_keep_potential.restype = C.c_int
        lib.fargo_monitor_disk.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, _DP]
        lib.fargo_monitor_disk.restype = C.c_int
        lib.fargo_halo_mode.argtypes = [C.c_void_p]
        lib.fargo_halo_mode.restype = C.c_int
        lib.fargo_launch_count.argtypes = [C.c_void_p]
        lib.fargo_launch_count.restype = C.c_longlong
        lib.fargo_profile_enable.argtypes = [C.c_void_p, C.c_int]
        lib.fargo_profile_enable.restype = C.c_int
        lib.fargo_profile_report.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        lib.fargo_profile_report.restype = C.c_int
        lib.fargo_event_record.argtypes = [C.c_void_p, C.c_int]
        lib.fargo_event_record.restype = C.c_int
        lib.fargo_event_elapsed_ms.argtypes = [C.c_void_p, C.c_int, C.c_int, _DP]
        lib.fargo_event_elapsed_ms.restype = C.c_int
        _lib = lib
    return _lib


class HydroContext(Handle):
    """Device context of libfargo_b200.so (one per GPU)."""

    def __init__(self, params, radii, rank=0, nranks=1, unique_id=None, device=0):
        lib = load_library()
        radii = np.ascontiguousarray(radii, dtype=np.float64)
        assert radii.shape == (params.nrad + 1,)
        ptr = C.c_void_p()
        uid = None
        if unique_id is not None:
            self._uid = C.create_string_buffer(bytes(unique_id), 128)
            uid = C.cast(self._uid, C.c_void_p)
        rc = lib.fargo_ctx_create(C.byref(ptr), C.byref(params), _dptr(radii), rank, nranks, uid, device)
        if rc != 0:
            raise RuntimeError("fargo_ctx_create failed: " + (lib.fargo_last_error() or b"").decode())
        super().__init__(lib, "fargo_", ptr.value, params, rank, nranks)

    def halo(self):
        self._check(self.lib.fargo_stage_halo(self.ptr), "stage_halo")

    def sync(self):
        self._check(self.lib.fargo_sync(self.ptr), "sync")

    def selftest_math(self, seed=1, blocks=1024, per_thread=256, wide=False):
        out = (C.c_ulonglong * 4)()
        self._check(self.lib.fargo_selftest_math(self.ptr, seed, blocks, per_thread, int(wide), out), "selftest_math")
        return {"div_mismatch": out[0], "sqrt_mismatch": out[1], "exp_mismatch": out[2], "div_fast": out[3],
                "pairs": 256 * blocks * per_thread}

    def selftest_exp(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        self._check(self.lib.fargo_selftest_exp(self.ptr, x.size, _dptr(x), _dptr(y)), "selftest_exp")
        return y

    def selftest_ringsum(self, rows):
        """Ring sums of the rows of a 2-D float64 array by the scan kernel and by the plain chain (fargo_selftest_ringsum)."""
        x = np.ascontiguousarray(rows, dtype=np.float64)
        a, b = np.empty(x.shape[0]), np.empty(x.shape[0])
        self._check(self.lib.fargo_selftest_ringsum(self.ptr, x.shape[0], x.shape[1], _dptr(x), _dptr(a), _dptr(b)), "selftest_ringsum")
        return a, b

    def set_staged(self, on):
        """step() through the per-stage kernels (one per reference loop nest) instead of the fused ones."""
        self._check(self.lib.fargo_set_staged(self.ptr, int(on)), "set_staged")

    def snapshot_async(self, sigma, vrad, vazi, energy=None):
        """Start an asynchronous snapshot into global-shaped float64 arrays (pinned memory overlaps with later steps)."""
        for a in (sigma, vrad, vazi, energy):
            assert a is None or (a.dtype == np.float64 and a.flags["C_CONTIGUOUS"])
        self._check(self.lib.fargo_snapshot_async(self.ptr, _dptr(sigma), _dptr(vrad), _dptr(vazi),
                                                  _dptr(energy) if energy is not None else None), "snapshot_async")

    def snapshot_wait(self):
        self._check(self.lib.fargo_snapshot_wait(self.ptr), "snapshot_wait")

    def halo_mode(self):
        """0 single rank, 1 NCCL send/recv, 2 peer-memory stores from the transport kernel (fargo_halo_mode)."""
        return int(self.lib.fargo_halo_mode(self.ptr))

    def launch_count(self):
        return int(self.lib.fargo_launch_count(self.ptr))

    def event_record(self, slot):
        self._check(self.lib.fargo_event_record(self.ptr, slot), "event_record")

    def event_elapsed_ms(self, a, b):
        ms = C.c_double(0.0)
        self._check(self.lib.fargo_event_elapsed_ms(self.ptr, a, b, C.byref(ms)), "event_elapsed_ms")
        return ms.value

    def profile(self, on):
        self._check(self.lib.fargo_profile_enable(self.ptr, int(on)), "profile_enable")

    def profile_report(self):
        """{kernel: (total_ms, launches)} since profile(True)."""
        buf = C.create_string_buffer(1 << 16)
        self.lib.fargo_profile_report(self.ptr, buf, len(buf))
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, n = line.rsplit(" ", 2)
            out[name] = (float(ms), int(n))
        return out

    def close(self):
        if self.ptr:
            self.lib.fargo_ctx_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def get_unique_id():
    lib = load_library()
    buf = C.create_string_buffer(128)
    if lib.fargo_get_unique_id(C.cast(buf, C.c_void_p)) != 0:
        raise RuntimeError("fargo_get_unique_id failed: " + (lib.fargo_last_error() or b"").decode())
    return bytes(buf.raw)
