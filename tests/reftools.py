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


from fargocpt_b200.config import params_from_config, _num, _flag  # noqa: E402,F401


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
        if self.ctx.params.leapfrog:
            self.ctx.step_leapfrog(self.time, step_dt)
        else:
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
