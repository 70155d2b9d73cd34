"""Parity at BASELINE.json's full sizes (configs[4]: 8192 x 16384, adiabatic planet-disk).

The oracle cannot run 1.3e8 cells in seconds, so the full grid is covered by size-independent properties:
  * a full-width annulus (all 16384 sectors: 142 warp windows per ring, the TMA ring-mean pipeline, the rotated
    FARGO loads) of 48 rings is checked against the oracle directly;
  * on the full 8192 x 16384 grid the fused marching kernels and the staged one-kernel-per-loop-nest path are two
    independent implementations of the same arithmetic and must agree bit for bit after several CFL-limited steps,
    the integer FARGO shifts included;
  * total mass changes only through the boundary rings (transport is conservative): interior mass is conserved
    to rounding.
"""
import numpy as np
import pytest

import reftools
from fargocpt_b200 import abi, synthetic

pytestmark = pytest.mark.gpu


def _start(ctx, cfg, fields):
    ctx.upload(abi.SIGMA, fields["Sigma"])
    ctx.upload(abi.ENERGY, fields["energy"])
    ctx.upload(abi.VRAD, fields["vrad"])
    ctx.upload(abi.VAZI, fields["vazi"])
    orbit = synthetic.PlanetOrbit(cfg)
    ctx.set_bodies(orbit.bodies(0.0))
    ctx.set_time(0.0)
    ctx.stage("boundary", 0.0, 0)
    ctx.copy_initial_values()  # before init_derived: the first CFL's Q- is evaluated against the beta-cooling reference state
    ctx.init_derived()
    return orbit


def _run(ctx, cfg, orbit, nsteps):
    last_dt, t, dts, shifts = float(cfg["FirstDT"]), 0.0, [], []
    for _ in range(nsteps):
        dt = ctx.cfl(last_dt)
        last_dt = dt
        dts.append(dt)
        ctx.set_bodies(orbit.bodies(t, dt))
        ctx.set_time(t)
        ctx.step(dt)
        shifts.append(ctx.nshift().copy())
        t += dt
    return dts, shifts


@pytest.mark.parametrize("physics", ["adiabatic_planet", "isothermal_planet"])
def test_full_width_annulus_vs_oracle(physics):
    from fargocpt_b200 import HydroContext
    nrad, naz = 48, 16384
    # same dr/r as the 8192-ring grid, annulus around the planet orbit
    g = (2.0 / 0.4) ** (1.0 / (8192 - 2.0))
    half = g ** ((nrad - 2) / 2.0)
    cfg = synthetic.make_config(physics, nrad, naz, Rmin=1.0 / half, Rmax=half, DampingInnerLimit=1.001, DampingOuterLimit=0.999)
    radii = synthetic.radii_from_config(cfg)
    params = synthetic.params_from_config(cfg)
    fields = synthetic.disk_fields(cfg, radii, perturb=1e-2)
    fields["vrad"] = fields["vrad"] + 1e-4 * np.cos(np.arange(naz) * 2 * np.pi * 5 / naz)[None, :]
    out = {}
    for name, ctx in (("gpu", HydroContext(params, radii)), ("cpu", reftools.OracleContext(params, radii))):
        orbit = _start(ctx, cfg, fields)
        dts, shifts = _run(ctx, cfg, orbit, 3)
        out[name] = (dts, shifts, {f: ctx.download(f) for f in (abi.SIGMA, abi.VRAD, abi.VAZI, abi.ENERGY)})
        ctx.close()
    adiabatic = physics == "adiabatic_planet"
    assert out["gpu"][0] == out["cpu"][0]  # the dt sequence, bit for bit
    for a, b in zip(out["gpu"][1], out["cpu"][1]):
        assert np.array_equal(a, b)
    for f in (abi.SIGMA, abi.VRAD, abi.VAZI, abi.ENERGY):
        if f == abi.ENERGY and not adiabatic:
            continue
        st = reftools.compare_stats(out["gpu"][2][f], out["cpu"][2][f])
        assert st["n_diff"] == 0, (f, st)


def test_full_grid_fused_equals_staged_and_conserves_mass():
    from fargocpt_b200 import HydroContext
    import torch
    free, _ = torch.cuda.mem_get_info()
    nrad, naz = (8192, 16384) if free > 90e9 else (2048, 4096)
    cfg = synthetic.make_config("adiabatic_planet", nrad, naz)
    radii = synthetic.radii_from_config(cfg)
    params = synthetic.params_from_config(cfg)
    fields = synthetic.disk_fields(cfg, radii, perturb=1e-2)
    surf = np.pi * (radii[1:] ** 2 - radii[:-1] ** 2) / naz
    res = {}
    for staged in (False, True):
        ctx = HydroContext(params, radii)
        ctx.set_staged(staged)
        orbit = _start(ctx, cfg, fields)
        dts, shifts = _run(ctx, cfg, orbit, 3)
        res[staged] = (dts, shifts, {f: ctx.download(f) for f in (abi.SIGMA, abi.VRAD, abi.VAZI, abi.ENERGY)})
        ctx.close()
    assert res[False][0] == res[True][0]
    for a, b in zip(res[False][1], res[True][1]):
        assert np.array_equal(a, b)
    for f in (abi.SIGMA, abi.VRAD, abi.VAZI, abi.ENERGY):
        assert np.array_equal(res[False][2][f], res[True][2][f]), f
    # mass: the rings well inside the damping zones' inner edges exchange mass only with their neighbours
    lo, hi = nrad // 3, 2 * nrad // 3
    sig0, sig1 = fields["Sigma"], res[False][2][abi.SIGMA]
    m0 = float((sig0[lo:hi] * surf[lo:hi, None]).sum())
    m1 = float((sig1[lo:hi] * surf[lo:hi, None]).sum())
    # flux through the two band edges over 3 steps is O(v_r dt / dr) of two rings out of (hi - lo)
    assert abs(m1 - m0) / m0 < 1e-5
    assert np.isfinite(sig1).all() and (sig1 > 0).all()


@pytest.mark.parametrize("physics", ["adiabatic_planet", "isothermal_planet"])
@pytest.mark.parametrize("naz", [160, 131])  # 131: odd sector count (no 16-byte vector path, ragged last window)
def test_damping_folded_into_transport_equals_own_pass_and_oracle(physics, naz, monkeypatch):
    """An Euler step applies the damping zones (damping.cpp:311-752) in the azimuthal transport kernel's epilogue; with
    FARGO_B200_FOLD_DAMPING=0 they run as their own pass (k_damping).  Both against the oracle, bit for bit, with wide zones,
    'Initial' and 'Zero' targets mixed (the density is damped towards the floor, damping.cpp), after several CFL-limited steps."""
    from fargocpt_b200 import HydroContext
    nrad = 96
    cfg = synthetic.make_config(physics, nrad, naz, DampingInnerLimit=1.4, DampingOuterLimit=0.7, DampingVRadialInner="Zero",
                                DampingVRadialOuter="Zero", DampingSurfaceDensityOuter="Zero", DampingEnergyInner="Zero")
    radii = synthetic.radii_from_config(cfg)
    params = synthetic.params_from_config(cfg)
    fields = synthetic.disk_fields(cfg, radii, perturb=2e-2)
    fields["vrad"] = fields["vrad"] + 1e-3 * np.cos(np.arange(naz) * 2 * np.pi * 3 / naz)[None, :]
    out = {}
    for name in ("folded", "own_pass", "oracle"):
        monkeypatch.setenv("FARGO_B200_FOLD_DAMPING", "0" if name == "own_pass" else "1")
        ctx = reftools.OracleContext(params, radii) if name == "oracle" else HydroContext(params, radii)
        orbit = _start(ctx, cfg, fields)
        l0 = ctx.launch_count() if name != "oracle" else 0
        dts, shifts = _run(ctx, cfg, orbit, 5)
        out[name] = (dts, {f: ctx.download(f) for f in (abi.SIGMA, abi.VRAD, abi.VAZI, abi.ENERGY)},
                     (ctx.launch_count() - l0) if name != "oracle" else 0)
        ctx.close()
    assert out["folded"][2] < out["own_pass"][2]  # the own pass is a launch per step more
    for name in ("folded", "own_pass"):
        assert out[name][0] == out["oracle"][0], name
        for f in (abi.SIGMA, abi.VRAD, abi.VAZI, abi.ENERGY):
            if f == abi.ENERGY and physics != "adiabatic_planet":
                continue
            st = reftools.compare_stats(out[name][1][f], out["oracle"][1][f])
            assert st["n_diff"] == 0, (name, f, st)
    # and the zones did something: the damped run differs from an undamped one
    cfg2 = dict(cfg, Damping="No")
    ctx = reftools.OracleContext(synthetic.params_from_config(cfg2), radii)
    orbit = _start(ctx, cfg2, fields)
    _run(ctx, cfg2, orbit, 5)
    assert reftools.compare_stats(ctx.download(abi.SIGMA), out["oracle"][1][abi.SIGMA])["n_diff"] > 100


def test_screened_cfl_equals_full_reduction_on_the_full_grid(monkeypatch):
    """fargo_condition_cfl bounds A of every cell with a cheap screen and evaluates the exact criterion only where the maximum
    can be (kernels_ring.cuh:k_cfl_screen / k_cfl_candidates).  FARGO_B200_CFL=check runs the full reduction (k_cfl) next to
    it on every call and fails the call unless the two dt agree bit for bit: BASELINE configs[4]'s grid (or 2048 x 4096 on a
    small device), perturbed disk + Jupiter, CFL-limited steps; the dt sequence also equals the full-only run's."""
    from fargocpt_b200 import HydroContext
    import torch
    free, _ = torch.cuda.mem_get_info()
    nrad, naz = (8192, 16384) if free > 90e9 else (2048, 4096)
    cfg = synthetic.make_config("adiabatic_planet", nrad, naz)
    radii = synthetic.radii_from_config(cfg)
    params = synthetic.params_from_config(cfg)
    fields = synthetic.disk_fields(cfg, radii, perturb=1e-2)
    seqs = {}
    for mode in ("check", "full"):
        monkeypatch.setenv("FARGO_B200_CFL", mode)
        ctx = HydroContext(params, radii)
        orbit = _start(ctx, cfg, fields)
        seqs[mode], _ = _run(ctx, cfg, orbit, 4)
        ctx.close()
    assert seqs["check"] == seqs["full"]


@pytest.mark.parametrize("physics", ["adiabatic_planet", "isothermal_planet"])
@pytest.mark.parametrize("nrad,naz", [(96, 131), (64, 2), (128, 1024)])
def test_screened_cfl_equals_full_reduction_small_grids(physics, nrad, naz, monkeypatch):
    """The same check on ragged / few-sector grids (no vector path, one partial block per ring) and against the oracle's dt."""
    from fargocpt_b200 import HydroContext
    cfg = synthetic.make_config(physics, nrad, naz)
    radii = synthetic.radii_from_config(cfg)
    params = synthetic.params_from_config(cfg)
    fields = synthetic.disk_fields(cfg, radii, perturb=2e-2)
    monkeypatch.setenv("FARGO_B200_CFL", "check")
    out = {}
    for name in ("gpu", "oracle"):
        ctx = reftools.OracleContext(params, radii) if name == "oracle" else HydroContext(params, radii)
        orbit = _start(ctx, cfg, fields)
        out[name], _ = _run(ctx, cfg, orbit, 5)
        ctx.close()
    assert out["gpu"] == out["oracle"]


@pytest.mark.parametrize("case", ["nan_cell", "zero_sigma", "negative_energy", "huge_velocity", "tiny_everything", "inf_q", "zero_energy",
                                  "uniform_state"])
def test_screened_cfl_on_adversarial_states(case, monkeypatch):
    """The screen of the two-pass CFL reduction must never hide the cell that owns the maximum, whatever the state: cells whose
    criterion is NaN (the reference ignores them: `dt_cell < dt` is false), zero / denormal densities (reciprocal seeds overflow),
    negative energies (c_s = sqrt(negative) is NaN in the reference, c_s^2 < 0 in the screen), velocities near overflow, values
    so small that every term underflows, and a perfectly uniform state (every block a candidate).  FARGO_B200_CFL=check compares
    with the full reduction inside the call; the oracle's condition_cfl must give the same double."""
    from fargocpt_b200 import HydroContext
    nrad, naz = 40, 1024
    cfg = synthetic.make_config("adiabatic_planet", nrad, naz)
    radii = synthetic.radii_from_config(cfg)
    params = synthetic.params_from_config(cfg)
    fields = {k: v.copy() for k, v in synthetic.disk_fields(cfg, radii, perturb=1e-2).items()}
    rng = np.random.default_rng(7)
    qp = 1e-7 * rng.random((nrad, naz))
    qm = 1e-7 * rng.random((nrad, naz))
    S, E, VR, VP = fields["Sigma"], fields["energy"], fields["vrad"], fields["vazi"]
    if case == "nan_cell":
        VR[7, 13] = np.nan
        E[9, 100] = np.nan
        qp[11, 5] = np.nan
    elif case == "zero_sigma":
        S[5, 17] = 0.0
        S[6, 18] = 5e-324
        S[20, 700] = 1e-300
    elif case == "negative_energy":
        E[8, 8] = -E[8, 8]
        E[30, 999] = -1e-30
    elif case == "huge_velocity":
        VR[12, 50] = 1e160
        VP[13, 51] = -3e153
        VR[14, 52] = 1e-160
    elif case == "tiny_everything":
        for a in (E, VR, VP):
            a *= 1e-170
        S *= 1e-20
        qp *= 1e-300
        qm *= 1e-300
    elif case == "inf_q":
        qp[15, 3] = np.inf
        qm[16, 4] = -np.inf
        qm[17, 5] = 1e300
    elif case == "zero_energy":
        E[18, 6] = 0.0
        qp[18, 6] = qm[18, 6] = 0.0  # 0 / 0 in invdt6
        E[19, 7] = 0.0
    elif case == "uniform_state":
        S[:] = S[:, :1]
        E[:] = E[:, :1]
        VR[:] = 0.0
        VP[:] = VP[:, :1]
        qp[:] = 0.0
        qm[:] = 0.0
    monkeypatch.setenv("FARGO_B200_CFL", "check")
    out = {}
    with np.errstate(all="ignore"):
        for name in ("gpu", "oracle"):
            ctx = reftools.OracleContext(params, radii) if name == "oracle" else HydroContext(params, radii)
            _start(ctx, cfg, fields)
            ctx.upload(abi.QPLUS, qp)
            ctx.upload(abi.QMINUS, qm)
            out[name] = ctx.condition_cfl()
            ctx.close()
    a, b = out["gpu"], out["oracle"]
    assert (a == b) or (np.isnan(a) and np.isnan(b)), (case, a, b)
