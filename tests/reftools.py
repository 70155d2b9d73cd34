"""Test-infrastructure helpers shared by tests/, bench.py (cpu_baseline leg) and smoke():
 - loading the CPU oracle (oracle/libfargo_oracle.so) behind the same Handle class as the CUDA lib,
 - turning a FargoCPT YAML config + the reference's constants/units output into a FargoParams,
 - a Python mirror of sim::run's time-loop logic (simulation.cpp:505-558) used to drive either side,
 - compare_binary_output-style statistics (Tools/compare_binary_output.py:16-44).
Nothing here is on the product path.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

from fargocpt_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "libfargo_oracle.so")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

_oracle = None


def build_oracle(force=False):
    src = os.path.join(ORACLE_DIR, "fargo_oracle.c")
    if force or not os.path.exists(ORACLE_LIB) or os.path.getmtime(ORACLE_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "libfargo_oracle.so"])
    return ORACLE_LIB


def load_oracle():
    global _oracle
    if _oracle is None:
        build_oracle()
        lib = C.CDLL(ORACLE_LIB)
        abi._bind(lib, "fargo_oracle_")
        lib.fargo_oracle_create.argtypes = [C.POINTER(abi.FargoParams), abi._DP, C.c_int, C.c_int]
        lib.fargo_oracle_create.restype = C.c_void_p
        lib.fargo_oracle_destroy.argtypes = [C.c_void_p]
        lib.fargo_oracle_upload_slab.argtypes = [C.c_void_p, C.c_int, abi._DP]
        lib.fargo_oracle_halo_pack.argtypes = [C.c_void_p, C.c_int, abi._DP]
        lib.fargo_oracle_halo_unpack.argtypes = [C.c_void_p, C.c_int, abi._DP]
        lib.fargo_oracle_step_pre.argtypes = [C.c_void_p, C.c_double]
        lib.fargo_oracle_step_post.argtypes = [C.c_void_p, C.c_double]
        _oracle = lib
    return _oracle


class OracleContext(abi.Handle):
    """CPU oracle slab with the same interface as fargocpt_b200.HydroContext."""

    def __init__(self, params, radii, rank=0, nranks=1):
        lib = load_oracle()
        radii = np.ascontiguousarray(radii, dtype=np.float64)
        ptr = lib.fargo_oracle_create(C.byref(params), abi._dptr(radii), rank, nranks)
        if not ptr:
            raise RuntimeError("fargo_oracle_create failed (mesh too narrow for this many ranks?)")
        super().__init__(lib, "fargo_oracle_", ptr, params, rank, nranks)

    def halo_pack(self, side):
        buf = np.zeros(4 * abi.CPUOVERLAP * self.naz)
        self.lib.fargo_oracle_halo_pack(self.ptr, side, abi._dptr(buf))
        return buf

    def halo_unpack(self, side, buf):
        buf = np.ascontiguousarray(buf, dtype=np.float64)
        self.lib.fargo_oracle_halo_unpack(self.ptr, side, abi._dptr(buf))

    def step_pre(self, dt):
        self.lib.fargo_oracle_step_pre(self.ptr, float(dt))

    def step_post(self, dt):
        self.lib.fargo_oracle_step_post(self.ptr, float(dt))

    def close(self):
        if self.ptr:
            self.lib.fargo_oracle_destroy(self.ptr)
            self.ptr = None


# ---------------------------------------------------------------------------------------------
# config -> params (mirror of the subset of parameters.cpp / Interpret.cpp / boundary_conditions/config.cpp
# the hot path reads).  `consts` = code-unit constants as the reference printed them (constants.yml).

def _flag(v, default=False):
    if v is None:
        return default
    if isinstance(v, bool):
        return v
    return str(v).strip().lower()[0] in ("y", "t", "1")


def _num(v, unit_cgs=None):
    """'3 K' with unit_cgs=<code temperature unit in K> -> 3/unit; plain numbers are code units."""
    if isinstance(v, (int, float)):
        return float(v)
    parts = str(v).split()
    x = float(parts[0])
    if len(parts) > 1 and unit_cgs is not None:
        return x / unit_cgs
    return x


def params_from_config(cfg, consts, nrad, naz, temp_unit_K=1.0):
    g = {k.lower(): v for k, v in cfg.items()}

    def get(key, default=None):
        return g.get(key.lower(), default)

    d = {}
    d["nrad"], d["naz"] = int(nrad), int(naz)
    spacing = {"l": "logarithmic", "a": "arithmetic", "e": "exponential"}.get(
        str(get("RadialSpacing", "Arithmetic")).lower()[:1], "custom")
    d["radial_spacing"] = abi.SPACING[spacing]
    d["rmin"], d["rmax"] = float(get("Rmin")), float(get("Rmax"))
    eos = str(get("EquationOfState", "Isothermal")).lower()
    d["adiabatic"] = 1 if eos in ("ideal", "adiabatic", "perfect") else 0
    d["gamma"] = float(get("AdiabaticIndex", 1.4))
    d["mu"] = float(get("mu", 1.0))
    d["aspectratio_ref"] = float(get("AspectRatio", 0.05))
    d["flaring_index"] = float(get("FlaringIndex", 0.0))
    d["sigma0"] = _num(get("Sigma0", 173.0))
    d["sigma_floor"] = float(get("SigmaFloor", 1e-9))
    d["sigma_slope"] = float(get("SigmaSlope", 0.0))
    d["minimum_temperature"] = _num(get("MinimumTemperature", "3 K"), temp_unit_K)
    d["maximum_temperature"] = _num(get("MaximumTemperature", "1e100 K"), temp_unit_K)
    d["G"], d["Rgas"], d["sigma_sb"], d["c_light"] = consts["G"], consts["R"], consts["sigma"], consts["c"]
    d["hydro_center_mass"] = consts.get("hydro_center_mass", 1.0)
    d["cfl"] = float(get("CFL", 0.5))
    d["cfl_max_var"] = float(get("CFLmaxVar", 1.1))
    d["heating_cooling_cfl_limit"] = float(get("HeatingCoolingCFLlimit", 1.0))
    integ = str(get("Integrator", "Euler")).lower()
    d["leapfrog"] = 0 if integ.startswith("e") else 1
    d["fast_transport"] = 1 if str(get("Transport", "FARGO")).lower().startswith("f") else 0
    # Interpret.cpp:640-664 compares case-sensitively: only the exact strings "mc" / "m" select MC
    d["flux_limiter"] = 1 if str(get("FluxLimiter", "VanLeer")) in ("mc", "m") else 0
    d["artificial_viscosity"] = abi.ARTVISC[str(get("ArtificialViscosity", "SN")).lower()]
    d["artificial_viscosity_factor"] = float(get("ArtificialViscosityFactor", 1.41))
    d["artificial_viscosity_dissipation"] = int(_flag(get("ArtificialViscosityDissipation"), True))
    d["viscous_alpha"] = float(get("ViscousAlpha", 0.0))
    d["constant_viscosity"] = _num(get("ConstantViscosity", 0.0))
    d["stabilize_viscosity"] = int(get("StabilizeViscosity", 0))
    d["radial_viscosity_factor"] = float(get("RadialViscosityFactor", 1.0))
    d["heating_viscous"] = int(_flag(get("HeatingViscous"), False))
    d["heating_viscous_factor"] = float(get("HeatingViscousFactor", 1.0))
    d["cooling_beta"] = int(_flag(get("CoolingBetaLocal"), False))
    d["cooling_beta_value"] = float(get("CoolingBeta", 1.0))
    d["cooling_beta_ramp_up"] = _num(get("CoolingBetaRampUp", 0.0))
    d["cooling_beta_reference"] = abi.BETA_REF[str(get("CoolingBetaReference", "zero")).lower()]
    d["body_force_from_potential"] = int(_flag(get("BodyForceFromPotential"), True))
    d["thickness_smoothing"] = float(get("ThicknessSmoothing", 0.0))
    d["imposed_disk_drift"] = float(get("ImposedDiskDrift", 0.0))

    # boundaries: composite names (boundary_conditions/config.cpp:345-436) or individual keys
    comp = {"zerogradient": ("zerogradient", "zerogradient", "zerogradient"),
            "outflow": ("zerogradient", "zerogradient", "outflow"),
            "reflecting": ("zerogradient", "zerogradient", "reflecting"),
            "reference": ("reference", "reference", "reference")}
    for side, name in ((0, "Inner"), (1, "Outer")):
        c = str(get(name + "Boundary", "individual")).lower()
        if c in comp:
            s, e, vr = comp[c]
        else:
            s = str(get(name + "BoundarySigma", "zerogradient")).lower()
            e = str(get(name + "BoundaryEnergy", "zerogradient")).lower()
            vr = str(get(name + "BoundaryVrad", "zerogradient")).lower()
        va = str(get(name + "BoundaryVazi", "keplerian")).lower()
        d.setdefault("bc_sigma", [0, 0])[side] = abi.BC[s]
        d.setdefault("bc_energy", [0, 0])[side] = abi.BC[e]
        d.setdefault("bc_vrad", [0, 0])[side] = abi.BC[vr]
        d.setdefault("bc_vazi", [0, 0])[side] = abi.BC[va]
        d.setdefault("keplerian_azimuthal_factor", [1.0, 1.0])[side] = float(
            get(name + "BoundaryVaziKeplerianFactor", 1.0))
    d["damping"] = int(_flag(get("Damping"), False))
    d["damping_inner_limit"] = float(get("DampingInnerLimit", 1.05))
    d["damping_outer_limit"] = float(get("DampingOuterLimit", 0.95))
    d["damping_time_factor"] = float(get("DampingTimeFactor", 1.0))
    d["damping_time_radius_outer"] = float(get("DampingTimeRadiusOuter", d["rmax"]))
    for key, name in (("damp_vrad", "VRadial"), ("damp_vazi", "VAzimuthal"), ("damp_sigma", "SurfaceDensity"),
                      ("damp_energy", "Energy")):
        d[key] = [abi.DAMP[str(get("Damping" + name + side, "None")).lower()] for side in ("Inner", "Outer")]
    return d


def make_params(d):
    return abi.FargoParams.from_dict(d)


# ---------------------------------------------------------------------------------------------
class TimeLoop:
    """Mirror of sim::CalculateTimeStep / sim::run bookkeeping (simulation.cpp:100-118, 505-558).

    `ctxs` is a list of slab handles (one per rank); single-rank callers pass one handle."""

    def __init__(self, ctx, first_dt, monitor_timestep, time=0.0):
        self.ctx = ctx
        self.last_dt = first_dt
        self.monitor_timestep = monitor_timestep
        self.time = time
        self.n_monitor = 0
        self.n_iter = 0
        self.dts = []

    def calculate_time_step(self):
        self.cfl_dt = self.ctx.cfl(self.last_dt)
        self.last_dt = self.cfl_dt
        return self.cfl_dt

    def next_dt(self):
        cfl_dt = self.calculate_time_step()
        time_next_monitor = (self.n_monitor + 1) * self.monitor_timestep
        left = time_next_monitor - self.time
        overshoot = cfl_dt > left
        almost_there = left < cfl_dt * (1 + 0.05)
        return (left if (overshoot or almost_there) else cfl_dt), cfl_dt, time_next_monitor

    def advance(self, bodies_fn=None):
        """One iteration of the for-loop in sim::run.  Returns True when a monitor boundary was hit."""
        step_dt, cfl_dt, time_next_monitor = self.next_dt()
        if bodies_fn is not None:
            self.ctx.set_bodies(bodies_fn(self.time, step_dt))
        self.ctx.set_time(self.time)
        self.ctx.step(step_dt)
        self.time += step_dt
        self.n_iter += 1
        self.dts.append(step_dt)
        if abs(time_next_monitor - self.time) < 1e-6 * cfl_dt:
            self.n_monitor += 1
            return True
        return False


def compare_stats(a, b):
    """Tools/compare_binary_output.py:16-44 statistics."""
    a, b = np.asarray(a).ravel(), np.asarray(b).ravel()
    diff = np.abs(a - b)
    nz = diff != 0
    denom = np.maximum(np.abs(a), np.abs(b))
    rel = np.where(denom > 0, diff / np.where(denom > 0, denom, 1), 0.0)
    return {"n_diff": int(nz.sum()), "n": int(a.size), "max_abs": float(diff.max(initial=0.0)),
            "max_rel": float(rel.max(initial=0.0))}


def load_golden(name):
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    z = np.load(path, allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    return meta, z
