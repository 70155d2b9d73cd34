#!/usr/bin/env python
"""The acceptance checks of the reference's own tests, run through `fargocpt_b200 start` on this repo's minimal
setups of the same physics (tests/golden/*_setup.yml) — no reference tree needed, so it runs on the GPU box too:

    python tests/checkers/run_reference_acceptance.py            # oracle-bound driver (CPU; host/fargocpt_b200_oracle_test)
    python tests/checkers/run_reference_acceptance.py --gpu      # the product: host/fargocpt_b200 on libfargo_b200.so

  test/shockTube           check_results.py:15-20     integrated |numerical - exact Sod| at t = 0.228 below 0.0073 / 0.0153 / 0.014 / 0.016
  test/spreading_ring      calc_deviation.py:38-66    mean |Sigma / Sigma_analytic - 1| < 0.007 at t = 314.159 (39 870 hydro steps)
  test/cold_disk_planet    calc_deviation.py:24-35    max |T(100 orbits) / T(0) - 1| < 0.1 (14 000 hydro steps)
  test/irradiation         check_results.py:40-120    max |T / T_theory - 1| < 0.03 for 2 < r < 15 au after 62 800 time units (168 596 steps)
  test/steady_state_accretion  check_results.py:69-116  max | |MassFlow| / (1e-8 solMass / yr) - 1 | < 2.2e-4 for 20 < r < 60 au after
                                                        3 141 526 time units (110 240 steps; the time-averaged mass flow of WriteMassFlow)
Prints one line per check and exits non-zero if one fails."""
import os
import struct
import subprocess
import sys
import tempfile
import time

import numpy as np
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def run(exe, cfg, until):
    tmp = tempfile.mkdtemp(prefix="accept_")
    yml, out = os.path.join(tmp, "setup.yml"), os.path.join(tmp, "out")
    yaml.safe_dump(cfg, open(yml, "w"), sort_keys=False)
    t0 = time.time()
    res = subprocess.run([exe, "start", yml, "--out", out, "--until", str(until)], capture_output=True, text=True)
    if res.returncode != 0:
        raise SystemExit(res.stdout[-2000:] + res.stderr[-2000:])
    steps = res.stdout.split("Total Hydrosteps")[1].split(",")[0].strip()
    return out, int(steps), time.time() - t0


def main():
    gpu = "--gpu" in sys.argv
    exe = os.path.join(ROOT, "host", "fargocpt_b200" if gpu else "fargocpt_b200_oracle_test")
    ok = True
    # --- shock tube
    from scipy import integrate
    from test_host_driver import _sod_exact
    base = yaml.safe_load(open(os.path.join(GOLDEN, "shock_tube_setup.yml")))
    for integ, av in (("Euler", "TW"), ("Euler", "SN"), ("Leapfrog", "TW"), ("Leapfrog", "SN")):
        cfg = dict(base, Integrator=integ, ArtificialViscosity=av)
        out, steps, secs = run(exe, cfg, 1)
        r12 = np.loadtxt(os.path.join(out, "used_rad.dat"))
        r1 = 0.5 * (r12[1:] + r12[:-1]) - r12[0]
        nr = len(r1)
        sig = np.fromfile(os.path.join(out, "snapshots", "1", "Sigma.dat")).reshape(nr, 2).mean(axis=1)
        en = np.fromfile(os.path.join(out, "snapshots", "1", "energy.dat")).reshape(nr, 2).mean(axis=1)
        vr = np.fromfile(os.path.join(out, "snapshots", "1", "vrad.dat")).reshape(nr + 1, 2).mean(axis=1)
        vr = 0.5 * (vr[1:] + vr[:-1])
        m = (r1 >= 0) & (r1 <= 1)
        x = r1[m]
        rho, vel, prs = _sod_exact(x, 0.228)
        dev = (integrate.simpson(np.abs(sig[m] - rho), x=x), integrate.simpson(np.abs(vr[m] - vel), x=x),
               integrate.simpson(np.abs(en[m] - prs / 0.4), x=x), integrate.simpson(np.abs(0.4 * en[m] / sig[m] - prs / rho), x=x))
        good = dev[0] < 0.0073 and dev[1] < 0.0153 and dev[2] < 0.014 and dev[3] < 0.016
        ok &= good
        print(f"shockTube {integ:8s} {av}: Sigma {dev[0]:.4f} vrad {dev[1]:.4f} energy {dev[2]:.4f} T {dev[3]:.4f}  "
              f"({steps} steps, {secs:.1f} s)  {'PASS' if good else 'FAIL'}")
    # --- spreading ring
    from scipy.special import iv
    cfg = yaml.safe_load(open(os.path.join(GOLDEN, "spreading_ring_setup.yml")))
    cfg["MonitorTimestep"], cfg["Nsnapshots"] = 314.159265359, 1
    out, steps, secs = run(exe, cfg, 1)
    ri = np.loadtxt(os.path.join(out, "used_rad.dat"))
    rc = 2.0 / 3.0 * (ri[1:] ** 3 - ri[:-1] ** 3) / (ri[1:] ** 2 - ri[:-1] ** 2)
    sigma = np.fromfile(os.path.join(out, "snapshots", "1", "Sigma.dat")).reshape(256, 2).mean(axis=1)
    t = struct.unpack("<IIddddQ", open(os.path.join(out, "snapshots", "1", "misc.bin"), "rb").read())[2]
    tau = 12 * 4.77e-5 * t + 0.016
    theo = 1.0 / np.pi / tau / rc ** 0.25 * iv(0.25, 2.0 * rc / tau) * np.exp(-(1 + rc ** 2) / tau)
    dev = float(np.mean(np.abs(sigma / theo - 1)))
    ok &= dev < 0.007
    print(f"spreading_ring: mean |Sigma / analytic - 1| = {dev:.5f} (< 0.007)  ({steps} steps, {secs:.1f} s)  {'PASS' if dev < 0.007 else 'FAIL'}")
    # --- cold disk + planet
    cfg = yaml.safe_load(open(os.path.join(GOLDEN, "cold_disk_planet_setup.yml")))
    out, steps, secs = run(exe, cfg, 10)
    dims = [l for l in open(os.path.join(out, "dimensions.dat")) if not l.startswith("#")][-1].split()
    nr, naz = int(dims[4]), int(dims[5])
    prof = [np.fromfile(os.path.join(out, "snapshots", str(n), "Temperature.dat")).reshape(nr, naz).mean(axis=1) for n in (0, 10)]
    dev = float(np.max(np.abs(prof[1] / prof[0] - 1)))
    ok &= dev < 0.1
    print(f"cold_disk_planet ({nr} x {naz}): max |T(100 orbits) / T(0) - 1| = {dev:.5f} (< 0.1)  ({steps} steps, {secs:.1f} s)  "
          f"{'PASS' if dev < 0.1 else 'FAIL'}")
    # --- irradiation: thermal surface cooling against stellar irradiation (D'Angelo & Marzari 2012, eq. 16, as check_results.py
    # states it: constants and the stellar temperature of the theory curve copied from there)
    cfg = yaml.safe_load(open(os.path.join(GOLDEN, "irradiation_setup.yml")))
    out, steps, secs = run(exe, cfg, 10)
    dims = [l for l in open(os.path.join(out, "dimensions.dat")) if not l.startswith("#")][-1].split()
    nr, naz = int(dims[4]), int(dims[5])
    r12 = np.loadtxt(os.path.join(out, "used_rad.dat"))
    r = 2.0 / 3.0 * (r12[1:] ** 3 - r12[:-1] ** 3) / (r12[1:] ** 2 - r12[:-1] ** 2)  # Rmed (init.cpp:188-195)
    units = yaml.safe_load(open(os.path.join(out, "units.yml")))
    T = np.fromfile(os.path.join(out, "snapshots", "10", "Temperature.dat")).reshape(nr, naz).mean(axis=1) * float(units["temperature"]["cgs value"])
    mu, m_H, k_B, l0, m0, G = 2.35, 1.66054e-24, 1.38065e-16, 14959787070000, 1.98847e+33, 6.6743e-08
    eta, eps, Rs, Ts = 2 / 7, 0.5, 4.6505e-05 * l0, 100000
    rc = r * l0
    htheo = (eta * (1 - eps) * (k_B * Ts / (mu * m_H)) ** 4 * (Rs / (G * m0)) ** 4 * (rc / Rs) ** 2) ** (1 / 7)
    Ttheo = Ts * np.sqrt(Rs / rc) * ((1 - eps) * (0.4 * (Rs / rc) + htheo * eta)) ** (1 / 4)
    sel = (r > 2) & (r < 15)
    dev = float(np.max(np.abs(T - Ttheo)[sel] / Ttheo[sel]))
    ok &= dev < 0.03
    print(f"irradiation ({nr} x {naz}): max |T / T_theory - 1| = {dev:.5f} (< 0.03)  ({steps} steps, {secs:.1f} s)  {'PASS' if dev < 0.03 else 'FAIL'}")
    # --- steady-state accretion: the mass flow through every interface, averaged over the last snapshot interval (WriteMassFlow:
    # MassFlow1D.dat = radius, azimuthal sum, min, max per interface), against the accretion rate of the initial profile
    cfg = yaml.safe_load(open(os.path.join(GOLDEN, "steady_state_accretion_setup.yml")))
    out, steps, secs = run(exe, cfg, int(cfg["Nsnapshots"]))
    units = yaml.safe_load(open(os.path.join(out, "units.yml")))
    to_msun_yr = float(units["mass"]["cgs value"]) / float(units["time"]["cgs value"]) / (1.98847e33 / 3.15576e7)
    ri = np.loadtxt(os.path.join(out, "used_rad.dat"))
    rc = 2.0 / 3.0 * (ri[1:] ** 3 - ri[:-1] ** 3) / (ri[1:] ** 2 - ri[:-1] ** 2)
    mf = np.fromfile(os.path.join(out, "snapshots", str(cfg["Nsnapshots"]), "MassFlow1D.dat")).reshape(-1, 4)
    assert np.array_equal(mf[:, 0], ri)
    diffval = np.abs(mf[1:-1, 1] * to_msun_yr) / 1e-8 - 1  # data[1:-1] of check_results.py
    inds = (rc[1:] > 20) & (rc[:-1] < 60)		   # x_[1:] > xmin and x_[:-1] < xmax, x_ the cell centres
    dev = float(np.max(np.abs(diffval[inds])))
    ok &= dev < 2.2e-4
    print(f"steady_state_accretion: max | |Mdot| / 1e-8 Msun/yr - 1 | = {dev:.3e} (< 2.2e-4)  ({steps} steps, {secs:.1f} s)  {'PASS' if dev < 2.2e-4 else 'FAIL'}")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
