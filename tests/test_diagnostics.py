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


# ---------------------------------------------------------------------------------------------------------------------
# monitor/Quantities.dat sums (output::write_quantities output.cpp:326-520 -> quantities.cpp:51-480)
def _quantities_fixture():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "quantities.json")))["quantities"]


def _load_with_q(ctx, meta, z, k):
    _load(ctx, meta, z, k)
    if f"Qplus_{k}" in z.files and meta["params"]["adiabatic"]:  # the Q+- of the step that ended at snapshot k, as the reference kept them
        ctx.upload(abi.QPLUS, z[f"Qplus_{k}"])
        ctx.upload(abi.QMINUS, z[f"Qminus_{k}"])


QUANT_CASES = [("adia_planet_100", 50), ("adia_planet_100", 100), ("iso_planet_100", 50), ("iso_planet_100", 100), ("rey_star", 2),
               ("adia_star", 3), ("adia_star", 6)]


@pytest.mark.parametrize("name,k", QUANT_CASES)
def test_oracle_monitor_quantities_match_reference(name, k):
    """The oracle's serial sums against the reference's own Quantities.dat (recorded with OMP_NUM_THREADS=1, i.e. the same
    summation order): bit for bit for every quantity the fixture carries the inputs of."""
    meta, z, ctx = _oracle(name)
    _load_with_q(ctx, meta, z, k)
    got = ctx.monitor_quantities()
    ref = _quantities_fixture()[name][str(k)]
    names = ["mass", "angular_momentum", "kinetic_energy", "radial_kinetic_energy", "azimuthal_kinetic_energy"]
    if meta["params"]["adiabatic"]:
        names += ["internal_energy"]
        if f"Qplus_{k}" in z.files:
            names += ["viscous_dissipation", "luminosity"]
    for q in names:
        assert got[q] == ref[q], (q, got[q], ref[q])
    # the mass-weighted columns (fargo_monitor_disk): disk radius, eccentricity, periastron, aspect ratio — bit for bit too
    disk = ctx.monitor_disk(frame_angle=meta["misc"][k].get("frame_angle", 0.0))
    for q in ("radius", "eccentricity", "periastron", "aspect_ratio", "advection_torque", "viscous_torque"):
        assert disk[q] == ref[q], (q, disk[q], ref[q])


@pytest.mark.gpu
@pytest.mark.parametrize("name,k", QUANT_CASES)
def test_gpu_monitor_quantities_vs_oracle(name, k):
    """The device sums (fixed order, but not the serial one) against the oracle's: 1e-13 relative; and reproducible."""
    from fargocpt_b200 import HydroContext
    meta, z, cpu = _oracle(name)
    gpu = HydroContext(reftools.make_params(meta["params"]), z["radii"])
    for ctx in (cpu, gpu):
        _load_with_q(ctx, meta, z, k)
    a, b, a2 = gpu.monitor_quantities(), cpu.monitor_quantities(), gpu.monitor_quantities()
    assert a == a2
    for q in abi.MONITOR_QUANTITIES:
        assert abs(a[q] - b[q]) <= 1e-13 * max(abs(b[q]), 1e-300), (q, a[q], b[q])
    # a radius limit inside the grid: both leave out the same rings
    rl = float(0.5 * (z["radii"][len(z["radii"]) // 2] + z["radii"][len(z["radii"]) // 2 + 1]))
    a, b = gpu.monitor_quantities(rl), cpu.monitor_quantities(rl)
    for q in abi.MONITOR_QUANTITIES:
        assert abs(a[q] - b[q]) <= 1e-13 * max(abs(b[q]), 1e-300), (q, a[q], b[q])
    assert 0.0 < b["mass"] < cpu.monitor_quantities()["mass"]
    # ComputeCircumPlanetaryMasses: the mass inside a Roche radius around the planet (or a point of the disk)
    for x, y, roche in ((0.92, 0.39, 0.07), (-1.3, 0.2, 0.25), (0.45, 0.0, 0.1), (5.0, 0.0, 0.1)):
        a, b = gpu.circumplanetary_mass(x, y, roche), cpu.circumplanetary_mass(x, y, roche)
        assert abs(a - b) <= 1e-13 * abs(b), (x, y, roche, a, b)
        assert (b > 0.0) == (x < 3.0)
    # fargo_monitor_disk: per-ring device sums against the oracle's serial ones; the radius is a ring's Rmed (same ring)
    fa = meta["misc"][k].get("frame_angle", 0.0)
    for limit in (1e300, rl):
        a, b, a2 = gpu.monitor_disk(limit, 0.99, fa), cpu.monitor_disk(limit, 0.99, fa), gpu.monitor_disk(limit, 0.99, fa)
        assert a == a2
        assert a["radius"] == b["radius"], (a["radius"], b["radius"])
        for q in ("aspect_ratio", "mass", "viscous_torque"):
            assert abs(a[q] - b[q]) <= 1e-12 * abs(b[q]), (q, a[q], b[q])
        # no kick since the state was loaded: the POTENTIAL grid is the one init_derived / the oracle's derived stage left
        for q in ("potential_energy", "gravitational_torque"):
            assert np.isfinite(a[q]) == np.isfinite(b[q]), (q, a[q], b[q])
        # the advection torque is a sum of cell values of both signs: relative to the sum of their magnitudes (~ mass x r v_r v_phi)
        assert abs(a["advection_torque"] - b["advection_torque"]) <= 1e-12 * max(abs(b["advection_torque"]), 1e-3 * b["mass"]), (a, b)
        for q in ("ecc_x", "ecc_y"):  # means of O(h^2) cell values that cancel around the ring: absolute
            assert abs(a[q] - b[q]) <= 1e-15, (q, a[q], b[q])


# ---------------------------------------------------------------------------------------------------------------------
# accretion::AccreteOntoSinglePlanet (accretion.cpp:84-221)
@pytest.mark.parametrize("name", ["iso_accrete_20", "adia_accrete_20", "iso_sinkhole_20", "adia_viscacc_20"])
def test_oracle_accreted_mass_matches_reference(name):
    """The mass the oracle takes out of the Hill sphere in every step against the planet's recorded m_accreted_mass (the
    reference sums it with an OpenMP reduction: 1e-12).  The fields of the same runs are held bit for bit in
    test_oracle_vs_golden.py."""
    meta, z, ctx = _oracle(name)
    acc = []
    goldenrun.run_fixture(ctx, meta, z, accreted=acc)
    assert len(acc) == meta["nsnap"]
    for k, body, dm, dpx, dpy in acc:
        ref = meta["bodies"][k][body][8]
        assert dm > 0.0 and abs(dm - ref) <= 1e-12 * ref, (k, dm, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["iso_accrete_20", "adia_accrete_20", "iso_sinkhole_20", "adia_viscacc_20"])
def test_gpu_accretion_vs_oracle(name):
    """One accretion call from identical states: the cells changed bit for bit, the sums within 1e-13."""
    from fargocpt_b200 import HydroContext
    meta, z, cpu = _oracle(name)
    gpu = HydroContext(reftools.make_params(meta["params"]), z["radii"])
    res = []
    for ctx in (cpu, gpu):
        _load(ctx, meta, z, 10)
        method = meta["config"]["nbody"][1].get("accretion method", "kley")
        res.append(ctx.accrete_kley(*goldenrun.accretion_inputs(meta, 10, 1, meta["monitor_timestep"]), method=method))
    for fid in (abi.SIGMA, abi.ENERGY) if meta["params"]["adiabatic"] else (abi.SIGMA,):
        st = reftools.compare_stats(gpu.download(fid), cpu.download(fid))
        assert st["n_diff"] == 0, (fid, st)
    assert (gpu.download(abi.SIGMA) != z["Sigma_10"]).sum() > 4  # the Hill sphere covers cells on this grid
    for a, b in zip(res[1], res[0]):
        assert abs(a - b) <= 1e-13 * abs(b), (res)


# ---------------------------------------------------------------------------------------------------------------------
# correct_v_azimuthal (SideEuler.cpp:79-95) for corotating frames
def test_oracle_correct_vazi():
    meta, z, ctx = _oracle("iso_planet_100")
    _load(ctx, meta, z, 50)
    before = ctx.download(abi.VAZI)
    ctx.correct_vazi(0.125)
    radii = z["radii"]
    rmed = 2.0 / 3.0 * (radii[1:] ** 3 - radii[:-1] ** 3) / (radii[1:] ** 2 - radii[:-1] ** 2)
    assert np.allclose(ctx.download(abi.VAZI), before - 0.125 * rmed[:, None], rtol=1e-15, atol=0.0)


@pytest.mark.gpu
def test_gpu_correct_vazi_vs_oracle():
    from fargocpt_b200 import HydroContext
    meta, z, cpu = _oracle("iso_planet_100")
    gpu = HydroContext(reftools.make_params(meta["params"]), z["radii"])
    for ctx in (cpu, gpu):
        _load(ctx, meta, z, 50)
        ctx.correct_vazi(-3.0e-4)
    st = reftools.compare_stats(gpu.download(abi.VAZI), cpu.download(abi.VAZI))
    assert st["n_diff"] == 0, st
    # and a step on top of it (the marching kernels read the corrected buffer)
    for ctx in (cpu, gpu):
        ctx.set_time(0.0)
        ctx.step(1.0e-3)
    for fid in (abi.SIGMA, abi.VRAD, abi.VAZI):
        st = reftools.compare_stats(gpu.download(fid), cpu.download(fid))
        assert st["n_diff"] == 0, (fid, st)


# ---------------------------------------------------------------------------------------------------------------------
# grids with very few sectors (BASELINE configs[0] has Naz = 2): the reductions' per-block partials must not depend on
# the Nr x Nphi scratch field being large (round-1 failure: "scratch too small for the force partials")
@pytest.mark.gpu
@pytest.mark.parametrize("naz", [1, 2, 3, 5])
@pytest.mark.parametrize("physics", ["isothermal_planet", "adiabatic_planet"])
def test_gpu_reductions_on_grids_with_few_sectors(naz, physics):
    from fargocpt_b200 import HydroContext, synthetic
    cfg = synthetic.make_config(physics, 64, naz)
    radii = synthetic.radii_from_config(cfg)
    params = synthetic.params_from_config(cfg)
    fields = synthetic.disk_fields(cfg, radii, perturb=1e-2)
    orbit = synthetic.PlanetOrbit(cfg)
    gpu = HydroContext(params, radii)
    cpu = reftools.OracleContext(params, radii)
    res = {}
    for name, ctx in (("gpu", gpu), ("cpu", cpu)):
        for fid, key in goldenrun.STATE:
            if key == "energy" and not params.adiabatic:
                continue
            ctx.upload(fid, fields[key])
        ctx.set_bodies(orbit.bodies(0.0))
        ctx.set_time(0.0)
        ctx.copy_initial_values()  # the beta-cooling reference state Q- is evaluated against
        ctx.init_derived()
        force = [ctx.disk_on_body_accel(b) for b in (0, 1)]
        quant = ctx.monitor_quantities()
        acc = ctx.accrete_kley(1.0, 0.0, 0.3, 0.5)
        res[name] = (force, quant, acc, ctx.download(abi.SIGMA))
    g, c = res["gpu"], res["cpu"]
    # the pull on the star at the origin is a sum that cancels to rounding: compare on the scale of the terms (the planet's pull)
    fscale = max(np.abs(c[0][1]).max(), 1e-300)
    for a, b in zip(g[0], c[0]):
        assert np.allclose(a, b, rtol=1e-12, atol=1e-13 * fscale), (a, b)
    for q in abi.MONITOR_QUANTITIES:
        scale = abs(c[1][q])
        if q == "luminosity":  # sum of surf * Q-, Q- ~ (e - e_ref) = rounding residue at t = 0: judge it on the scale of Q+
            scale = max(scale, abs(c[1]["viscous_dissipation"]))
        assert np.isfinite(c[1][q]) and abs(g[1][q] - c[1][q]) <= 1e-13 * max(scale, 1e-300), (q, g[1][q], c[1][q])
    for a, b in zip(g[2], c[2]):
        assert abs(a - b) <= 1e-13 * max(abs(b), 1e-300), (g[2], c[2])
    assert c[2][0] > 0.0  # the zone covers cells: something was accreted
    assert reftools.compare_stats(g[3], c[3])["n_diff"] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("physics", ["adiabatic_planet", "isothermal_planet"])
def test_gpu_potential_columns_need_the_kept_grid(physics):
    """The "potential energy" / "gravitational torque" columns read the POTENTIAL grid of the last kick's start.  The fused kernels
    hold the potential in registers: after a step without fargo_keep_potential the columns are NaN (never a stale number), with
    it they equal the oracle's (which stores the grid in every kick like the reference); before the first step the grid is
    zeros on both sides."""
    from fargocpt_b200 import HydroContext, synthetic
    import test_gpu_fullsize as F
    nrad, naz = 64, 256
    cfg = synthetic.make_config(physics, nrad, naz)
    radii = synthetic.radii_from_config(cfg)
    params = synthetic.params_from_config(cfg)
    fields = synthetic.disk_fields(cfg, radii, perturb=1e-2)
    gpu, cpu = HydroContext(params, radii), reftools.OracleContext(params, radii)
    orbits = [F._start(c, cfg, fields) for c in (gpu, cpu)]
    a, b = gpu.monitor_disk(), cpu.monitor_disk()
    assert a["potential_energy"] == b["potential_energy"] == 0.0 and a["gravitational_torque"] == b["gravitational_torque"] == 0.0
    for c, o in zip((gpu, cpu), orbits):
        F._run(c, cfg, o, 2)
    a, b = gpu.monitor_disk(), cpu.monitor_disk()
    assert np.isnan(a["potential_energy"]) and np.isnan(a["gravitational_torque"]) and np.isfinite(b["potential_energy"])
    gpu.keep_potential(True)
    for c, o in zip((gpu, cpu), orbits):
        F._run(c, cfg, o, 1)
    a, b = gpu.monitor_disk(), cpu.monitor_disk()
    assert abs(a["potential_energy"] - b["potential_energy"]) <= 1e-13 * abs(b["potential_energy"]), (a, b)
    assert abs(a["gravitational_torque"] - b["gravitational_torque"]) <= 1e-12 * max(abs(b["gravitational_torque"]), 1e-6 * b["mass"]), (a, b)
    gpu.keep_potential(False)
    F._run(gpu, cfg, orbits[0], 1)
    assert np.isnan(gpu.monitor_disk()["potential_energy"])


@pytest.mark.gpu
def test_gpu_massflow_grid_vs_oracle():
    """fargo_track_massflow: the MASSFLOW grid the radial sweep accumulates (TransportEuler.cpp:610-616) against the oracle's, bit
    for bit, over several steps; fargo_clear_massflow zeroes it; an interface grid ([nrad + 1][naz])."""
    from fargocpt_b200 import HydroContext, synthetic
    import test_gpu_fullsize as F
    nrad, naz = 48, 131
    cfg = synthetic.make_config("adiabatic_planet", nrad, naz, InnerBoundary="outflow", OuterBoundary="outflow", Damping="No")
    radii = synthetic.radii_from_config(cfg)
    params = synthetic.params_from_config(cfg)
    fields = synthetic.disk_fields(cfg, radii, perturb=2e-2)
    fields["vrad"] = fields["vrad"] + 1e-3 * np.cos(np.arange(naz) * 2 * np.pi * 3 / naz)[None, :]
    out, flows = {}, {}
    for name, ctx in (("gpu", HydroContext(params, radii)), ("cpu", reftools.OracleContext(params, radii))):
        orbit = F._start(ctx, cfg, fields)
        ctx.track_massflow(True)
        ctx.track_boundary_flow(True)
        F._run(ctx, cfg, orbit, 4)
        mf = ctx.download(abi.MASSFLOW)
        assert mf.shape == (nrad + 1, naz)
        ctx.clear_massflow()
        out[name] = (mf, ctx.download(abi.MASSFLOW))
        flows[name] = (ctx.boundary_flow(reset=True), ctx.boundary_flow(reset=False))
        ctx.close()
    # MassDelta's boundary flows (TransportEuler.cpp:578-608): gas leaves through both outflow boundaries with the perturbed v_rad; per-column sums on
    # the device against the oracle's per-step sums: rounding
    a, b = flows["gpu"], flows["cpu"]
    assert b[0][1] > 0.0 and b[0][3] > 0.0 and a[1] == b[1] == (0.0, 0.0, 0.0, 0.0)  # outflow boundaries let nothing in
    assert np.allclose(a[0], b[0], rtol=1e-13, atol=0.0), (a[0], b[0])
    st = reftools.compare_stats(out["gpu"][0], out["cpu"][0])
    assert st["n_diff"] == 0, st
    assert np.abs(out["cpu"][0]).max() > 0 and not out["gpu"][1].any() and not out["cpu"][1].any()


@pytest.mark.gpu
@pytest.mark.parametrize("naz", [160, 131])
def test_gpu_damping_mass_vs_oracle(naz):
    """fargo_track_damping_mass: MassDelta's wave-damping mass creation / removal of both zones (damping.cpp:335-357 and siblings).
    While tracked, Sigma's zones are damped in their own pass (the other fields stay folded into the transport kernel): the fields
    must stay the oracle's bit for bit, the four sums agree to rounding (per-column sums on the device)."""
    from fargocpt_b200 import HydroContext, synthetic
    import test_gpu_fullsize as F
    nrad = 96
    cfg = synthetic.make_config("adiabatic_planet", nrad, naz, DampingInnerLimit=1.4, DampingOuterLimit=0.7, DampingSurfaceDensityOuter="Zero")
    radii = synthetic.radii_from_config(cfg)
    params = synthetic.params_from_config(cfg)
    fields = synthetic.disk_fields(cfg, radii, perturb=2e-2)
    out = {}
    for name, ctx in (("gpu", HydroContext(params, radii)), ("cpu", reftools.OracleContext(params, radii))):
        orbit = F._start(ctx, cfg, fields)
        ctx.track_damping_mass(True)
        dts, _ = F._run(ctx, cfg, orbit, 4)
        out[name] = (dts, {f: ctx.download(f) for f in (abi.SIGMA, abi.VRAD, abi.VAZI, abi.ENERGY)}, ctx.damping_mass(reset=True), ctx.damping_mass(reset=False))
        ctx.close()
    assert out["gpu"][0] == out["cpu"][0]
    for f in (abi.SIGMA, abi.VRAD, abi.VAZI, abi.ENERGY):
        assert reftools.compare_stats(out["gpu"][1][f], out["cpu"][1][f])["n_diff"] == 0, f
    a, b = out["gpu"][2], out["cpu"][2]
    assert max(b) > 0.0 and out["gpu"][3] == out["cpu"][3] == (0.0, 0.0, 0.0, 0.0)
    assert np.allclose(a, b, rtol=1e-12, atol=1e-20), (a, b)
