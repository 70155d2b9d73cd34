#!/usr/bin/env python
"""A few steps on a small grid through the entry points added late in round 2 (two-pass CFL, fargo_monitor_disk,
fargo_circumplanetary_mass, fargo_keep_potential, fargo_track_massflow, S-curve alpha / cooling), meant to run under
compute-sanitizer (memcheck / racecheck): tools/gpu_r2_sanitize.sh."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import goldenrun  # noqa: E402
import reftools  # noqa: E402
from fargocpt_b200 import HydroContext, abi, synthetic  # noqa: E402

for physics, nrad, naz in (("adiabatic_planet", 40, 131), ("isothermal_planet", 36, 256)):
    cfg = synthetic.make_config(physics, nrad, naz)
    radii = synthetic.radii_from_config(cfg)
    ctx = HydroContext(synthetic.params_from_config(cfg), radii)
    f = synthetic.disk_fields(cfg, radii, perturb=1e-2)
    for fid, k in ((abi.SIGMA, "Sigma"), (abi.ENERGY, "energy"), (abi.VRAD, "vrad"), (abi.VAZI, "vazi")):
        ctx.upload(fid, f[k])
    orbit = synthetic.PlanetOrbit(cfg)
    ctx.set_bodies(orbit.bodies(0.0))
    ctx.set_time(0.0)
    ctx.stage("boundary", 0.0, 0)
    ctx.copy_initial_values()
    ctx.init_derived()
    ctx.track_massflow(True)
    last_dt, t = float(cfg["FirstDT"]), 0.0
    for step in range(4):
        dt = ctx.cfl(last_dt)
        last_dt = dt
        ctx.keep_potential(step == 3)
        ctx.set_bodies(orbit.bodies(t, dt))
        ctx.set_time(t)
        ctx.step(dt)
        t += dt
    d = ctx.monitor_disk(1e300, 0.99, 0.1)
    m = ctx.circumplanetary_mass(0.9, 0.3, 0.2)
    mf = ctx.download(abi.MASSFLOW)
    ctx.clear_massflow()
    print(physics, "dt", last_dt, "radius", d["radius"], "E_pot", d["potential_energy"], "mdcp", m, "massflow", float(np.abs(mf).sum()))
    ctx.close()
for name in ("adia_scurve", "iso_sn_std"):
    meta, z = reftools.load_golden(name)
    ctx = HydroContext(reftools.make_params(meta["params"]), z["radii"])
    snaps = goldenrun.run_fixture(ctx, meta, z, nsteps=2)
    print(name, snaps[-1]["last_dt"] == meta["misc"][2]["last_dt"])
    ctx.close()
print("sanitize run done")
