"""stress::calculate_Reynolds_stress (stress.cpp:34-70) and ComputeDiskOnPlanetAccel (Force.cpp:23-122).

CPU leg: the oracle restatements against values recorded from the unmodified reference (T_Reynolds.dat bit for bit; the
disk-on-planet acceleration, which the reference sums with an OpenMP reduction in no defined order, to 1e-12).
GPU leg: the CUDA kernels against the oracle (T_Reynolds bit for bit; the force components to 1e-12 — the device sums
in a fixed but different order)."""
import numpy as np
import pytest

import goldenrun
import reftools
from fargocpt_b200 import abi


def _load(ctx, meta, z, k):
    for fid, name in goldenrun.STATE:
        ctx.upload(fid, z[f"{name}_{k}"])
    omega = float(meta["config"].get("OmegaFrame", 0.0))
    ctx.set_bodies(goldenrun.bodies_at(meta, k, omega))
    ctx.set_time(meta["misc"][k]["time"])
    ctx.init_derived()


def _oracle(name):
    meta, z = reftools.load_golden(name)
    return meta, z, reftools.OracleContext(reftools.make_params(meta["params"]), z["radii"])


def test_oracle_reynolds_stress_matches_reference():
    meta, z, ctx = _oracle("rey_star")
    for k in (1, 2):
        _load(ctx, meta, z, k)
        st = reftools.compare_stats(ctx.download(abi.T_REYNOLDS), z[f"T_Reynolds_{k}"])
        assert st["n_diff"] == 0, (k, st)


@pytest.mark.parametrize("name,k", [("rey_star", 2), ("adia_planet_100", 50), ("iso_planet_100", 100)])
def test_oracle_disk_on_planet_accel_matches_reference(name, k):
    meta, z, ctx = _oracle(name)
    _load(ctx, meta, z, k)
    for body, rec in enumerate(meta["bodies"][k]):
        a = ctx.disk_on_body_accel(body)
        got = np.array([a[0] + a[2], a[1] + a[3]])  # Force.cpp:117-119
        ref = np.array(rec[5:7])
        assert np.allclose(got, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max()), (body, got, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("name,k", [("rey_star", 2), ("adia_planet_100", 50), ("iso_planet_100", 100)])
def test_gpu_diagnostics_vs_oracle(name, k):
    from fargocpt_b200 import HydroContext
    meta, z, cpu = _oracle(name)
    gpu = HydroContext(reftools.make_params(meta["params"]), z["radii"])
    for ctx in (cpu, gpu):
        _load(ctx, meta, z, k)
    st = reftools.compare_stats(gpu.download(abi.T_REYNOLDS), cpu.download(abi.T_REYNOLDS))
    assert st["n_diff"] == 0, st
    for body in range(len(meta["bodies"][k])):
        a, b = gpu.disk_on_body_accel(body), cpu.disk_on_body_accel(body)
        assert np.allclose(a, b, rtol=1e-12, atol=1e-13 * np.abs(b).max()), (body, a, b)
    # reproducible: the device sums in a fixed order
    assert np.array_equal(gpu.disk_on_body_accel(1), gpu.disk_on_body_accel(1))
