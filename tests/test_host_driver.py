"""The C++ host driver (host/fargo_host.cpp) as a drop-in on a FargoCPT output directory.

A directory in the reference's own format (config.yml, constants.yml, units.yml, dimensions.dat, used_rad.dat,
snapshots/0/{Sigma,vrad,vazi,energy,Qplus,Qminus}.dat, misc.bin, nbodyK.bin, snapshots/reference/) is materialised from a
golden fixture (recorded from the unmodified reference), the host restarts from snapshot 0 and continues the run, and
the snapshot files it writes are compared with the ones the reference wrote:
  * star-only configs: every field file byte-identical, misc.bin (time, last_dt, N_iter) identical;
  * with a planet the bodies are advanced by the host's RK4 instead of REBOUND's IAS15 (out of scope), which agrees to
    rounding per step; fields then agree to the north_star tolerance (<= 1e-10) instead of bit for bit.
CPU leg: the same driver bound to the oracle (host logic without a GPU).  GPU leg: the real binary on libfargo_b200.so.
"""
import os
import struct
import subprocess

import numpy as np
import pytest
import yaml

import reftools

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def materialise(name, dest):
    """Write the fixture as a FargoCPT output directory (formats: SURVEY.md §5 'Checkpoint / resume')."""
    meta, z = reftools.load_golden(name)
    cfg = {k: v for k, v in meta["config"].items() if not k.startswith("_")}
    nrad, naz = meta["params"]["nrad"], meta["params"]["naz"]
    os.makedirs(os.path.join(dest, "snapshots", "0"))
    os.makedirs(os.path.join(dest, "snapshots", "reference"))
    c = meta["consts"]
    with open(os.path.join(dest, "constants.yml"), "w") as f:
        for sym in ("G", "R", "sigma", "c"):
            f.write(f"{sym} constant:\n  symbol: {sym}\n  code value: {c[sym]!r}\n\n")
    with open(os.path.join(dest, "units.yml"), "w") as f:
        f.write(f"temperature:\n  cgs symbol: K\n  cgs value: {meta['temperature_unit_K']!r}\n")
    with open(os.path.join(dest, "dimensions.dat"), "w") as f:
        f.write("#RMIN\tRMAX\tPHIMIN\tPHIMAX\tNRAD\tNAZ\tNGHRAD\tNGHAZ\tRadial_spacing\n")
        f.write(f"{cfg['Rmin']!r}\t{cfg['Rmax']!r}\t0\t{2 * np.pi!r}\t{nrad}\t{naz}\t1\t1\t{cfg['RadialSpacing']}\n")
    with open(os.path.join(dest, "used_rad.dat"), "w") as f:
        for r in z["radii"]:
            f.write(f"{float(r)!r}\n")
    for sub in ("0", "reference"):
        sd = os.path.join(dest, "snapshots", sub)
        yaml.safe_dump(cfg, open(os.path.join(sd, "config.yml"), "w"), sort_keys=False)
        for fname in ("Sigma", "vrad", "vazi", "energy", "Qplus", "Qminus"):
            if f"{fname}_0" in z:
                z[f"{fname}_0"].astype(np.float64).tofile(os.path.join(sd, fname + ".dat"))
        m = meta["misc"][0]
        with open(os.path.join(sd, "misc.bin"), "wb") as f:  # output.h:16-24
            f.write(struct.pack("<IIddddQ", 0, 0, m["time"], m["omega_frame"], m["frame_angle"], m["last_dt"], m["n_iter"]))
        for k, b in enumerate(meta["bodies"][0]):  # nbody/planet.h:11-45, 256 bytes
            rec = bytearray(256)
            struct.pack_into("<5d", rec, 8, *b[:5])
            if len(b) >= 12:  # accreting bodies: efficiency, distance to the primary, Roche radius, semi-major axis (planet.h:20-36)
                struct.pack_into("<d", rec, 56, b[7])
                struct.pack_into("<2d", rec, 152, b[9], b[10])
                struct.pack_into("<d", rec, 176, b[11])
            with open(os.path.join(sd, f"nbody{k}.bin"), "wb") as f:
                f.write(bytes(rec))
    return meta, z


def run_host(exe, name, tmp_path, until):
    src, out = str(tmp_path / "ref"), str(tmp_path / "out")
    meta, z = materialise(name, src)
    res = subprocess.run([exe, "restart", "0", src, "--out", out, "--until", str(until)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    return meta, z, out


def check(meta, z, out, k, exact):
    nrad, naz = meta["params"]["nrad"], meta["params"]["naz"]
    raw = open(os.path.join(out, "snapshots", str(k), "misc.bin"), "rb").read()
    ts, nts, time, omega, angle, last_dt, n_iter = struct.unpack("<IIddddQ", raw)
    m = meta["misc"][k]
    assert (ts, n_iter, time) == (k, m["n_iter"], m["time"])
    assert angle == m["frame_angle"]
    if exact:
        assert last_dt == m["last_dt"]
    else:
        assert last_dt == pytest.approx(m["last_dt"], rel=1e-10)
    for fname, rings in (("Sigma", nrad), ("vrad", nrad + 1), ("vazi", nrad), ("energy", nrad)):
        if f"{fname}_{k}" not in z:
            continue
        got = np.fromfile(os.path.join(out, "snapshots", str(k), fname + ".dat")).reshape(rings, naz)
        ref = z[f"{fname}_{k}"]
        if exact:
            assert got.tobytes() == ref.tobytes(), (fname, reftools.compare_stats(got, ref))
        else:  # Tools/compare_binary_output.py statistics, relative to the field's scale (v_rad crosses zero)
            st = reftools.compare_stats(got, ref)
            assert st["max_abs"] <= 1e-10 * float(np.abs(ref).max()), (fname, st)


def _oracle_exe():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "host"), "oracle_test"])
    return os.path.join(ROOT, "host", "fargocpt_b200_oracle_test")


@pytest.mark.parametrize("name", ["iso_star", "adia_star", "adia_sn_stab", "adia_leapfrog"])
def test_host_driver_star_only_bytes_identical_cpu(name, tmp_path):
    meta, z, out = run_host(_oracle_exe(), name, tmp_path, 6)
    for k in (1, 3, 6):
        check(meta, z, out, k, exact=True)
    assert open(os.path.join(out, "snapshots", "list.txt")).read().split() == [str(k) for k in range(1, 7)]


def test_host_driver_planet_within_tolerance_cpu(tmp_path):
    meta, z, out = run_host(_oracle_exe(), "iso_planet_100", tmp_path, 50)
    check(meta, z, out, 50, exact=False)


def test_host_driver_disk_feedback_cpu(tmp_path):
    """DiskFeedback: yes — ComputeDiskOnNbodyAccel, the velocity kick and the disk part of the indirect term every step."""
    meta, z, out = run_host(_oracle_exe(), "iso_feedback_20", tmp_path, 20)
    check(meta, z, out, 20, exact=False)
    raw = open(os.path.join(out, "snapshots", "20", "nbody1.bin"), "rb").read()
    got = np.array(struct.unpack("<5d", raw[8:48]))
    assert np.allclose(got, np.array(meta["bodies"][20][1][:5]), rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("name", ["iso_accrete_20", "adia_accrete_20", "adia_viscacc_20"])
def test_host_driver_accretion_cpu(name, tmp_path):
    """A planet that accretes (accretion.cpp:84-221) without feeling the disk: the driver takes gas out of its Hill sphere
    first thing in every step.  Its orbital period is the one of the restart record (the reference refreshes it every step),
    so the fields agree to the planet tolerance, and the gas the planet swallowed is in its record."""
    meta, z, out = run_host(_oracle_exe(), name, tmp_path, 20)
    check(meta, z, out, 20, exact=False)
    raw = open(os.path.join(out, "snapshots", "20", "nbody1.bin"), "rb").read()
    (accreted,) = struct.unpack("<d", raw[64:72])
    # like the reference, the counter is reset at every monitor output (t_planet::write, planet.cpp:323-327)
    assert accreted == pytest.approx(meta["bodies"][20][1][8], rel=1e-9)
    rows = [l.split("\t") for l in open(os.path.join(out, "monitor", "nbody1.dat")) if not l.startswith("#")]
    assert len(rows) == 20 and all(len(r) == 22 for r in rows)
    rate = [float(r[21]) * meta["monitor_timestep"] for r in rows]  # column 21: accretion rate = accreted mass / monitor step
    assert rate == pytest.approx([meta["bodies"][k][1][8] for k in range(1, 21)], rel=1e-9)
    # and it matters: without the accretion the surface density near the planet is off by far more than the tolerance
    nrad, naz = meta["params"]["nrad"], meta["params"]["naz"]
    got = np.fromfile(os.path.join(out, "snapshots", "20", "Sigma.dat")).reshape(nrad, naz)
    assert np.abs(got - z["Sigma_0"]).max() > 1e-6 * np.abs(z["Sigma_0"]).max()


@pytest.mark.gpu
@pytest.mark.parametrize("name,until,exact", [("iso_accrete_20", 20, False), ("adia_viscacc_20", 20, False), ("iso_star", 6, True), ("adia_star", 6, True), ("adia_leapfrog", 6, True), ("adia_planet_100", 100, False),
                                              ("iso_feedback_20", 20, False)])
def test_host_driver_on_gpu(name, until, exact, tmp_path):
    exe = os.path.join(ROOT, "host", "fargocpt_b200")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "host")])
    meta, z, out = run_host(exe, name, tmp_path, until)
    check(meta, z, out, until, exact)


# ---------------------------------------------------------------------------------------------------------------------
# `fargocpt_b200 start <setup.yml>`: units, constants, radial grid, N-body initial state and the power-law disk (host/fargo_init.hpp)
def start_host(exe, name, tmp_path, until):
    out = str(tmp_path / "out")
    yml = os.path.join(ROOT, "tests", "golden", name + ".yml")  # the very setup file the reference ran to record the fixture
    res = subprocess.run([exe, "start", yml, "--out", out, "--until", str(until)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    meta, z = reftools.load_golden(name)
    return meta, z, out


def check_start(meta, z, out, until, exact):
    """Snapshot 0 — the state the reference builds from the setup file — must be the reference's bit for bit: the code-unit
    constants, the radii, every field, misc.bin, and what the hydro path reads off the N-body records."""
    cyml = yaml.safe_load(open(os.path.join(out, "constants.yml")))
    consts = {v["symbol"]: float(v["code value"]) for v in cyml.values()}
    assert consts == {k: float(v) for k, v in meta["consts"].items() if not k.endswith("_cgs")}
    for k, v in meta["consts"].items():  # newer fixtures also carry the cgs values the S-curve cooling fit reads
        if k.endswith("_cgs"):
            assert [float(c["cgs value"]) for c in cyml.values() if c["symbol"] == k[:-4]] == [float(v)], k
    units = yaml.safe_load(open(os.path.join(out, "units.yml")))
    assert float(units["temperature"]["cgs value"]) == meta["temperature_unit_K"]
    assert np.array_equal(np.loadtxt(os.path.join(out, "used_rad.dat")), z["radii"])
    check(meta, z, out, 0, exact=True)
    nrad, naz = meta["params"]["nrad"], meta["params"]["naz"]
    for fname in ("Qplus", "Qminus"):  # the reference's first Q- is NaN where the beta-cooling reference does not exist yet
        if f"{fname}_0" in z and os.path.exists(os.path.join(out, "snapshots", "0", fname + ".dat")):
            got = np.fromfile(os.path.join(out, "snapshots", "0", fname + ".dat")).reshape(nrad, naz)
            assert np.array_equal(got, z[f"{fname}_0"], equal_nan=True), fname
    for k, b in enumerate(meta["bodies"][0]):
        raw = open(os.path.join(out, "snapshots", "0", f"nbody{k}.bin"), "rb").read()
        assert len(raw) == 256
        assert list(struct.unpack("<5d", raw[8:48])) == b[:5], (k, struct.unpack("<5d", raw[8:48]), b[:5])
        if len(b) >= 12:  # accretion efficiency, distance to the primary, Roche radius, semi-major axis
            assert struct.unpack("<d", raw[56:64])[0] == b[7]
            assert list(struct.unpack("<2d", raw[152:168])) == b[9:11]
            assert struct.unpack("<d", raw[176:184])[0] == b[11]
    if os.path.exists(os.path.join(out, "snapshots", "reference")):
        for fname in ("Sigma", "vrad", "vazi"):
            a = open(os.path.join(out, "snapshots", "reference", fname + ".dat"), "rb").read()
            assert a == open(os.path.join(out, "snapshots", "0", fname + ".dat"), "rb").read()
    check(meta, z, out, until, exact)


START_CASES = [("iso_star", 6, True), ("adia_star", 6, True), ("adia_cold", 6, True), ("adia_sn_stab", 6, True), ("iso_sn_std", 6, True),
               ("ring_like", 6, True), ("adia_leapfrog", 6, True), ("iso_planet_100", 50, False), ("adia_accrete_20", 10, False),
               ("iso_feedback_20", 20, False),
               # sinkhole accretion; an accreting planet that feels the disk (update_planet, Roche radius and orbital period refreshed)
               ("iso_sinkhole_20", 20, False), ("adia_accfb_20", 20, False),
               # SurfaceCooling: thermal, irradiating star (ramped), constant opacity and the two opacity tables
               ("adia_irrad", 6, True), ("adia_irrad_lf", 6, True), ("adia_cool_lin", 6, True), ("adia_cool_bell", 6, False),
               # EquationOfState: PVTE: lookup tables built by host/fargo_pvte.h, the reference's refresh order of gamma_eff / mu / Gamma_1
               ("adia_pvte", 6, True), ("adia_pvte_lf", 6, True),
               # AlphaMode 1: S-curve alpha in the stored temperature (viscosity/viscosity.cpp:31-49), Euler and Leapfrog
               ("adia_alpha_scurve", 6, True), ("adia_alpha_scurve_lf", 6, True),
               # SurfaceCooling: scurve (scurve_cooling, SourceEuler.cpp:726-831): ScurveType Kimura / Ichikawa
               ("adia_scurve", 6, True), ("adia_scurve_ichikawa_lf", 6, True),
               # v_azi boundaries Balanced (v_sq of the disk model formed by the host like balanced.cpp:23-52) and ZeroShear
               ("iso_bc_balanced", 6, True), ("adia_bc_zeroshear", 6, True),
               # inner v_rad boundaries Viscous (ViscousOutflowSpeed) and Keplerian (InnerBoundaryVradKeplerianFactor)
               ("iso_bc_viscous", 6, True), ("adia_bc_keplerian_vrad", 6, True),
               # ring-mean damping (the mean stays in column 0 of the initial-value grid) + Reference boundaries + beta cooling towards them
               ("adia_damp_mean_ref", 12, False)]


@pytest.mark.parametrize("name,until,exact", START_CASES)
def test_host_start_from_setup_file_cpu(name, until, exact, tmp_path):
    """From the setup YAML alone to the reference's snapshots: star-only runs byte-identical all the way, runs with a planet
    byte-identical at snapshot 0 and within the planet tolerance (RK4 instead of IAS15) afterwards."""
    meta, z, out = start_host(_oracle_exe(), name, tmp_path, until)
    check_start(meta, z, out, until, exact)


def test_host_start_output_restarts_cpu(tmp_path):
    """The directory `start` writes is a FargoCPT output directory: `restart` continues from it to the same bytes."""
    exe = _oracle_exe()
    meta, z, out = start_host(exe, "adia_star", tmp_path, 3)
    out2 = str(tmp_path / "out2")
    res = subprocess.run([exe, "restart", "3", out, "--out", out2, "--until", "6"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    check(meta, z, out2, 6, exact=True)


def test_host_refuses_a_unit_it_would_have_to_drop(tmp_path):
    """A value with a unit on a key this driver has no conversion for must not silently lose the unit (Config::number)."""
    cfg = yaml.safe_load(open(os.path.join(ROOT, "tests", "golden", "minimal_defaults_setup.yml")))
    cfg["DampingInnerLimit"] = "1.1 au"
    yml = str(tmp_path / "setup.yml")
    yaml.safe_dump(cfg, open(yml, "w"), sort_keys=False)
    res = subprocess.run([_oracle_exe(), "start", yml, "--out", str(tmp_path / "out"), "--until", "0"], capture_output=True, text=True)
    assert res.returncode != 0 and "carries a unit" in res.stderr, res.stderr


def test_host_start_refuses_what_it_does_not_cover(tmp_path):
    cfg = yaml.safe_load(open(os.path.join(ROOT, "tests", "golden", "iso_star.yml")))
    cfg["SigmaCondition"] = "1D"  # needs GSL splines in the reference; not restated
    yml = str(tmp_path / "setup.yml")
    yaml.safe_dump(cfg, open(yml, "w"), sort_keys=False)
    res = subprocess.run([_oracle_exe(), "start", yml, "--out", str(tmp_path / "out")], capture_output=True, text=True, timeout=60)
    assert res.returncode != 0 and "SigmaCondition" in res.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name,until,exact", [("adia_star", 6, True), ("iso_star", 6, True), ("iso_planet_100", 50, False),
                                              ("iso_sinkhole_20", 20, False), ("adia_accfb_20", 20, False)])
def test_host_start_from_setup_file_gpu(name, until, exact, tmp_path):
    exe = os.path.join(ROOT, "host", "fargocpt_b200")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "host")])
    meta, z, out = start_host(exe, name, tmp_path, until)
    check_start(meta, z, out, until, exact)


REFERENCE_SETUPS = [("/root/reference/test/cold_disk_planet/setup.yml", []), ("/root/reference/examples/config.yml", []),
                    # step_LeapFrog with a planet: forward-looking indirect term, potential after the frame rotation, rotating frame
                    ("/root/reference/test/cold_disk_planet/setup.yml", ["Integrator=Leapfrog"]),
                    ("/root/reference/examples/config.yml", ["Integrator=Leapfrog"]),
                    ("/root/reference/examples/config.yml", ["IndirectTermMode=1"]),
                    # four bodies: eccentric orbits from their elements (Jacobi coordinates), jupiterMass / earthMass units,
                    # cubic (Klahr) smoothing, ramp-up — this repo's own setup file, run through both codes
                    (os.path.join(ROOT, "tests", "golden", "multi_body_setup.yml"), ["--dt", "4e-3"]),
                    # ... two of them accreting (kley with a cubic smoothing radius, sinkhole) while they feel the disk: the Roche radii
                    # are refreshed inside AccreteOntoPlanets' loop over the bodies, after every body once one has gained mass
                    # (accretion.cpp:486-515) — one update per step instead left 2e-10 in the fields after 5 steps
                    (os.path.join(ROOT, "tests", "golden", "multi_body_accrete_setup.yml"), ["--dt", "4e-3", "--snapshots", "5"]),
                    # Frame: C — the frame follows the planet (refframe::handle_corotation: new OmegaFrame every step, v_azi
                    # corrected through fargo_correct_vazi), Euler / Leapfrog / with DiskFeedback and the predictor indirect term
                    # Fermi-function cut-offs of the initial profiles (also inside the numerically differentiated viscous speed), SetSigma0
                    (os.path.join(ROOT, "tests", "golden", "adia_planet_100.yml"),
                     ["--dt", "4e-3", "ProfileCutoffOuter=yes", "ProfileCutoffPointOuter=2.0", "ProfileCutoffWidthOuter=0.1",
                      "ProfileCutoffInner=yes", "ProfileCutoffPointInner=15 au", "ProfileCutoffWidthInner=0.05", "SetSigma0=yes", "DiskMass=0.02"]),
                    # a setup that sets almost nothing: the reference's DEFAULTS (Sigma0 173 g/cm2, l0 / m0, ThicknessSmoothing 0.6,
                    # HeatingViscous yes, HeatingCoolingCFLlimit 10, IndirectTermMode 0, ArtificialViscosity SN ...)
                    (os.path.join(ROOT, "tests", "golden", "minimal_defaults_setup.yml"), ["--dt", "5e-3"]),
                    # InitializePureKeplerian with alpha and with constant viscosity; the deprecated global KlahrSmoothingRadius
                    (os.path.join(ROOT, "tests", "golden", "adia_planet_100.yml"), ["--dt", "4e-3", "InitializePureKeplerian=yes", "KlahrSmoothingRadius=0.4"]),
                    (os.path.join(ROOT, "tests", "golden", "iso_planet_100.yml"),
                     ["--dt", "4e-3", "InitializePureKeplerian=yes", "ViscousAlpha=0", "ConstantViscosity=1e-5"]),
                    # circumbinary disk: HydroFrameCenter binary / all (frame centred on a centre of mass, indirect term over its bodies)
                    (os.path.join(ROOT, "tests", "golden", "circumbinary_setup.yml"), ["--dt", "2e-3"]),
                    (os.path.join(ROOT, "tests", "golden", "circumbinary_setup.yml"), ["--dt", "2e-3", "IndirectTermMode=0", "DiskFeedback=yes"]),
                    (os.path.join(ROOT, "tests", "golden", "circumbinary_setup.yml"), ["--dt", "2e-3", "HydroFrameCenter=all", "Integrator=Leapfrog"]),
                    (os.path.join(ROOT, "tests", "golden", "circumbinary_setup.yml"), ["--dt", "2e-3", "VazimuthalConsidersQuadropoleMoment=yes"]),
                    # SigmaCondition: Nbody — profiles and gas orbits centred on the centre of mass of all bodies
                    (os.path.join(ROOT, "tests", "golden", "circumbinary_setup.yml"), ["--dt", "2e-3", "SigmaCondition=Nbody"]),
                    (os.path.join(ROOT, "tests", "golden", "circumbinary_setup.yml"),
                     ["--dt", "2e-3", "SigmaCondition=Nbody", "HydroFrameCenter=primary", "ProfileCutoffOuter=yes", "ProfileCutoffPointOuter=2.2",
                      "ProfileCutoffWidthOuter=0.1"]),
                    (os.path.join(ROOT, "tests", "golden", "adia_planet_100.yml"), ["--dt", "2e-3", "SigmaCondition=Nbody"]),
                    (os.path.join(ROOT, "tests", "golden", "adia_planet_100.yml"), ["--dt", "2e-3", "EnergyCondition=Nbody"]),
                    # CircumBinaryRing: Gaussian ring on top of the profiles (density and energy; isothermal runs write the energy ring out too)
                    (os.path.join(ROOT, "tests", "golden", "circumbinary_setup.yml"),
                     ["--dt", "2e-3", "CircumBinaryRing=yes", "CircumBinaryRingPosition=1.5", "CircumBinaryRingWidth=0.2", "SigmaCondition=Nbody",
                      "SetSigma0=yes", "DiskMass=0.01"]),
                    (os.path.join(ROOT, "tests", "golden", "adia_planet_100.yml"),
                     ["--dt", "2e-3", "CircumBinaryRing=yes", "CircumBinaryRingPosition=1.5", "CircumBinaryRingWidth=0.2"]),
                    # step_LeapFrog with an accreting planet: the first kick reads the pressure stored BEFORE the accretion, the disk's
                    # pull is evaluated before AccreteOntoPlanets and applied after it (simulation.cpp:294-305, 352-408)
                    (os.path.join(ROOT, "tests", "golden", "iso_accrete_20.yml"), ["--dt", "4e-3", "Integrator=Leapfrog"]),
                    (os.path.join(ROOT, "tests", "golden", "adia_accfb_20.yml"), ["--dt", "4e-3", "Integrator=Leapfrog", "IndirectTermMode=0"]),
                    (os.path.join(ROOT, "tests", "golden", "corotating_setup.yml"), ["--dt", "4e-3"]),
                    (os.path.join(ROOT, "tests", "golden", "corotating_setup.yml"), ["--dt", "4e-3", "Integrator=Leapfrog"]),
                    (os.path.join(ROOT, "tests", "golden", "corotating_setup.yml"), ["--dt", "4e-3", "DiskFeedback=yes", "IndirectTermMode=0"])]


@pytest.mark.parametrize("setup,overrides", REFERENCE_SETUPS)
def test_host_start_on_the_references_own_setup_files(setup, overrides):
    """BASELINE configs[1] (test/cold_disk_planet/setup.yml: units, cps, ramped planet, damping, IndirectTermMode 0) and the
    physics of configs[3] (examples/config.yml: DiskFeedback) verbatim through the unmodified reference and through
    `fargocpt_b200 start`: identical constants / units / radii files, identical snapshot 0, fields within the north_star
    tolerance afterwards.  Needs the reference tree and oracle/_ref (build container only)."""
    if not (os.path.exists(setup) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "fargocpt_exe_ieee"))):
        pytest.skip("the reference tree / oracle/_ref are not available here")
    _oracle_exe()
    import importlib.util
    spec = importlib.util.spec_from_file_location("cmpstart", os.path.join(ROOT, "tests", "checkers", "compare_start_with_reference.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    import contextlib
    import io
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        worst = mod.main([setup, "--snapshots", "2", "--dt", "1e-3"] + overrides)
    text = buf.getvalue()
    for f in ("constants.yml", "units.yml", "used_rad.dat"):
        assert f + ": identical" in text, text
    snap0 = [l for l in text.splitlines() if l.startswith("snapshot 0:")][0]
    assert "ndiff=0" in snap0 and "DIFF" not in snap0 and snap0.count("identical") >= 2, snap0
    assert all("ndiff=0 " in part or "ndiff" not in part for part in snap0.split("max|d|")), snap0
    assert worst <= 1e-10, text


@pytest.mark.parametrize("name", ["adia_star", "iso_planet_100"])
def test_host_writes_quantities_dat_cpu(name, tmp_path):
    """monitor/Quantities.dat (output::write_quantities, output.cpp:326-493) from `start`: the reference's header and 35-column
    layout; the global sums equal the ones recorded from the reference's own Quantities.dat (tests/golden/quantities.json,
    bit for bit on the star-only run, to the planet tolerance otherwise), disk radius, eccentricity, periastron and aspect
    ratio included; columns this path does not evaluate are nan."""
    import json
    cfg = yaml.safe_load(open(os.path.join(ROOT, "tests", "golden", name + ".yml")))
    cfg["WriteDiskQuantities"] = "Yes"
    yml = str(tmp_path / "setup.yml")
    yaml.safe_dump(cfg, open(yml, "w"), sort_keys=False)
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "quantities.json")))["quantities"][name]
    until = max(int(k) for k in ref)
    out = str(tmp_path / "out")
    res = subprocess.run([_oracle_exe(), "start", yml, "--out", out, "--until", str(until)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    lines = open(os.path.join(out, "monitor", "Quantities.dat")).read().splitlines()
    assert lines[0] == "#FargoCPT quantities file" and lines[1] == "#version: 2.4"
    header = [l for l in lines if l.startswith("#variable:")]
    assert len(header) == 35 and header[3].startswith("#variable: 3 | mass | ") and header[34].startswith("#variable: 34 | gravitational torque | ")
    rows = [l.split("\t") for l in lines if not l.startswith("#")]
    assert [int(r[1]) for r in rows] == list(range(until + 1)) and all(len(r) == 35 for r in rows)
    cols = {"mass": 3, "angular_momentum": 5, "internal_energy": 7, "kinetic_energy": 8, "radial_kinetic_energy": 10,
            "azimuthal_kinetic_energy": 11, "viscous_dissipation": 14, "luminosity": 15,
            # the mass-weighted columns (fargo_monitor_disk)
            "radius": 4, "eccentricity": 12, "periastron": 13, "aspect_ratio": 26, "advection_torque": 32, "viscous_torque": 33,
            # from the POTENTIAL grid of the last step's start (fargo_keep_potential)
            "total_energy": 6, "potential_energy": 9, "gravitational_torque": 34,
            # MassDelta's boundary flows since the previous row (fargo_boundary_flow)
            "inner_boundary_inflow": 17, "inner_boundary_outflow": 18, "outer_boundary_inflow": 19, "outer_boundary_outflow": 20}
    for snap, want in ref.items():
        row = rows[int(snap)]
        assert int(row[0]) == int(snap)
        for q, c in cols.items():
            if name == "adia_star":
                assert float(row[c]) == want[q], (snap, q, row[c], want[q])
            else:
                assert float(row[c]) == pytest.approx(want[q], rel=1e-9, abs=1e-300), (snap, q)
        assert float(row[16]) == 0.0 and row[25] == "nan"  # pdivv without WritepDV; density-floor mass creation


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE configs[0]: test/spreading_ring (pressureless viscous ring, 256 x 2)
SPREADING_RING = os.path.join(ROOT, "tests", "golden", "spreading_ring_setup.yml")


def test_spreading_ring_setup_verbatim_cpu():
    """The reference's own test/spreading_ring/setup.yml through the unmodified reference and through `fargocpt_b200 start`
    (Bessel-function ring, SetSigma0, h = 0): every file identical, every field of every snapshot identical."""
    setup = "/root/reference/test/spreading_ring/setup.yml"
    if not (os.path.exists(setup) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "fargocpt_exe_ieee"))):
        pytest.skip("the reference tree / oracle/_ref are not available here")
    _oracle_exe()
    import contextlib
    import importlib.util
    import io
    spec = importlib.util.spec_from_file_location("cmpstart", os.path.join(ROOT, "tests", "checkers", "compare_start_with_reference.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        worst = mod.main([setup, "--snapshots", "3", "--dt", "0.05"])
    assert worst == 0.0 and buf.getvalue().count("misc identical") == 4, buf.getvalue()


def _run_start(exe, yml, out, until):
    res = subprocess.run([exe, "start", yml, "--out", out, "--until", str(until)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]


def test_spreading_ring_spreads_cpu(tmp_path):
    """Physics sanity of the ring on the committed minimal setup: mass is conserved to rounding while the peak drops."""
    out = str(tmp_path / "out")
    _run_start(_oracle_exe(), SPREADING_RING, out, 4)
    radii = np.loadtxt(os.path.join(out, "used_rad.dat"))
    surf = np.pi * (radii[1:] ** 2 - radii[:-1] ** 2) / 2
    s0 = np.fromfile(os.path.join(out, "snapshots", "0", "Sigma.dat")).reshape(256, 2)
    s4 = np.fromfile(os.path.join(out, "snapshots", "4", "Sigma.dat")).reshape(256, 2)
    m0, m4 = (surf[1:-1, None] * s0[1:-1]).sum(), (surf[1:-1, None] * s4[1:-1]).sum()
    assert m0 == pytest.approx(1.0, rel=1e-12) and m4 == pytest.approx(m0, rel=1e-9)
    assert s4.max() < s0.max() and np.array_equal(s4[:, 0], s4[:, 1])


@pytest.mark.gpu
def test_spreading_ring_gpu_equals_oracle(tmp_path):
    """BASELINE configs[0] on the B200: the CUDA path against the oracle-bound driver (the checker), byte for byte."""
    exe = os.path.join(ROOT, "host", "fargocpt_b200")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "host")])
    gpu, cpu = str(tmp_path / "gpu"), str(tmp_path / "cpu")
    _run_start(exe, SPREADING_RING, gpu, 4)
    _run_start(_oracle_exe(), SPREADING_RING, cpu, 4)
    for k in range(5):
        for f in ("Sigma.dat", "vrad.dat", "vazi.dat", "misc.bin"):
            a = open(os.path.join(gpu, "snapshots", str(k), f), "rb").read()
            b = open(os.path.join(cpu, "snapshots", str(k), f), "rb").read()
            assert a == b, (k, f)


def test_host_planet_monitor_files_cpu(tmp_path):
    """monitor/nbodyK.dat (t_planet::write_ascii, nbody/planet.cpp:279-372): header of file version 2 with 22 columns, one row per
    monitor step; positions / velocities / mass equal the snapshot records, the torque column is the disk's torque."""
    meta, z, out = start_host(_oracle_exe(), "iso_feedback_20", tmp_path, 20)
    lines = open(os.path.join(out, "monitor", "nbody1.dat")).read().splitlines()
    assert lines[0] == "#FargoCPT planet file for planet: planet" and lines[1] == "#version: 2"
    header = [l for l in lines if l.startswith("#variable:")]
    assert len(header) == 22 and header[12].startswith("#variable: 12 | semi-major axis | ")
    rows = [l.split("\t") for l in lines if not l.startswith("#")]
    assert [int(r[1]) for r in rows] == list(range(21))
    for k in (0, 20):
        raw = open(os.path.join(out, "snapshots", str(k), "nbody1.bin"), "rb").read()
        mass, x, y, vx, vy = struct.unpack("<5d", raw[8:48])
        assert [float(v) for v in rows[k][2:7]] == [x, y, vx, vy, mass]
    # DiskFeedback: the gas torque column is the accumulated torque per monitor step; it matches the recorded acceleration
    # (the record of snapshot 20 holds the acceleration computed at the START of step 20, where the planet was at its snapshot-19 position)
    b, at = meta["bodies"][20][1], meta["bodies"][19][1]
    torque = (at[1] * b[6] - at[2] * b[5]) * b[0]
    assert float(rows[20][18]) == pytest.approx(torque, rel=1e-9)
    assert 0.0 < float(rows[20][9]) < 1e-3  # circumplanetary mass (ComputeCircumPlanetaryMasses): pinned in tests/test_diagnostics.py


@pytest.mark.parametrize("name", ["adia_planet_100", "iso_planet_100"])
def test_host_circumplanetary_mass_is_the_references(name, tmp_path):
    """Column 9 (mdcp) of monitor/nbody1.dat — ComputeCircumPlanetaryMasses (circumplanetary_mass.cpp:11-51) with the Roche radius
    the N-body side keeps (update_roche_radii) — against the value the unmodified reference wrote (tests/golden/quantities.json,
    recorded with OMP_NUM_THREADS=1); the planet's position differs in its last bits (RK4 / IAS15), the set of cells does not."""
    import json
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "quantities.json")))["quantities"][name]
    until = max(int(k) for k in ref)
    meta, z, out = start_host(_oracle_exe(), name, tmp_path, until)
    rows = [l.split("\t") for l in open(os.path.join(out, "monitor", "nbody1.dat")).read().splitlines() if not l.startswith("#")]
    for snap, want in ref.items():
        assert int(rows[int(snap)][0]) == int(snap)
        assert float(rows[int(snap)][9]) == pytest.approx(want["mdcp"], rel=1e-10), (snap, rows[int(snap)][9], want["mdcp"])


def _massflow_check(exe, tmp_path):
    """WriteMassFlow (TransportEuler.cpp:610-616, quantities.cpp:771-781, polargrid.cpp:187-282): MassFlow.dat and MassFlow1D.dat of
    the steady-state accretion setup (198 x 1), shortened to two snapshots of three monitor steps, against the files the
    unmodified reference wrote for the same setup (tests/golden/steady_state_massflow.npz)."""
    cfg = yaml.safe_load(open(os.path.join(ROOT, "tests", "golden", "steady_state_accretion_setup.yml")))
    cfg["Nsnapshots"], cfg["Nmonitor"] = 2, 3
    yml, out = str(tmp_path / "setup.yml"), str(tmp_path / "out")
    yaml.safe_dump(cfg, open(yml, "w"), sort_keys=False)
    _run_start(exe, yml, out, 2)
    ref = np.load(os.path.join(ROOT, "tests", "golden", "steady_state_massflow.npz"))
    for k in (1, 2):
        for f in ("MassFlow", "MassFlow1D", "Sigma"):
            got = np.fromfile(os.path.join(out, "snapshots", str(k), f + ".dat"))
            assert np.array_equal(got, ref[f"{f}_{k}"]), (k, f, float(np.abs(got - ref[f"{f}_{k}"]).max()))
    assert np.abs(ref["MassFlow_2"]).max() > 0


def _steady_state_quantities_check(exe, tmp_path, rtol):
    """monitor/Quantities.dat of the shortened steady-state accretion run against the rows the unmodified reference wrote
    (tests/golden/steady_state_quantities.npz): every column this path evaluates — the inner-boundary outflow of the accreting
    disk (MassDelta, TransportEuler.cpp:578-608) among them — the wave-damping mass creation / removal of both zones too — and nan in the one it does not (density-floor mass creation)."""
    cfg = yaml.safe_load(open(os.path.join(ROOT, "tests", "golden", "steady_state_accretion_setup.yml")))
    cfg["Nsnapshots"], cfg["Nmonitor"] = 2, 3
    yml, out = str(tmp_path / "setup.yml"), str(tmp_path / "out")
    yaml.safe_dump(cfg, open(yml, "w"), sort_keys=False)
    _run_start(exe, yml, out, 2)
    ref = np.load(os.path.join(ROOT, "tests", "golden", "steady_state_quantities.npz"))["rows"]
    got = np.array([[float(x) for x in l.split()] for l in open(os.path.join(out, "monitor", "Quantities.dat")) if not l.startswith("#")])
    assert got.shape == ref.shape == (7, 35)
    nan_cols = [25]
    assert np.isnan(got[:, nan_cols]).all()
    cols = [c for c in range(35) if c not in nan_cols]
    assert ref[-1, 18] > 0  # the disk accretes through the inner boundary
    if rtol == 0.0:
        assert np.array_equal(got[:, cols], ref[:, cols]), [c for c in cols if not np.array_equal(got[:, c], ref[:, c])]
    else:
        scale = np.maximum(np.abs(ref[:, cols]).max(axis=0), 1e-6 * np.abs(ref[:, 3]).max())
        assert (np.abs(got[:, cols] - ref[:, cols]) / scale).max() <= rtol


def test_host_quantities_of_an_accreting_disk_are_the_references_cpu(tmp_path):
    _steady_state_quantities_check(_oracle_exe(), tmp_path, 0.0)


@pytest.mark.gpu
def test_host_quantities_of_an_accreting_disk_are_the_references_gpu(tmp_path):
    _steady_state_quantities_check(os.path.join(ROOT, "host", "fargocpt_b200"), tmp_path, 1e-12)


def test_host_writes_the_references_massflow_files_cpu(tmp_path):
    _massflow_check(_oracle_exe(), tmp_path)


@pytest.mark.gpu
def test_host_writes_the_references_massflow_files_gpu(tmp_path):
    _massflow_check(os.path.join(ROOT, "host", "fargocpt_b200"), tmp_path)


def test_host_start_reads_2d_profiles_like_the_reference(tmp_path):
    """SigmaCondition / EnergyCondition: 2D — non-axisymmetric profiles read from raw files (t_polargrid::read2D), through the
    reference and through `fargocpt_b200 start`."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "fargocpt_exe_ieee")):
        pytest.skip("oracle/_ref is not available here")
    _oracle_exe()
    meta, z = reftools.load_golden("adia_planet_100")
    s, e = z["Sigma_0"].copy(), z["energy_0"].copy()
    i, j = np.meshgrid(np.arange(s.shape[0]), np.arange(s.shape[1]), indexing="ij")
    s *= 1 + 0.05 * np.sin(3 * j * 2 * np.pi / s.shape[1]) * np.exp(-((i - 24) / 6.0) ** 2)
    e *= 1 + 0.02 * np.cos(2 * j * 2 * np.pi / s.shape[1])
    sf, ef = str(tmp_path / "sig2d.dat"), str(tmp_path / "en2d.dat")
    s.tofile(sf)
    e.tofile(ef)
    import contextlib
    import importlib.util
    import io
    spec = importlib.util.spec_from_file_location("cmpstart", os.path.join(ROOT, "tests", "checkers", "compare_start_with_reference.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        worst = mod.main([os.path.join(ROOT, "tests", "golden", "adia_planet_100.yml"), "--snapshots", "2", "--dt", "4e-3", "SigmaCondition=2D",
                          f"SigmaFilename={sf}", "EnergyCondition=2D", f"EnergyFilename={ef}"])
    snap0 = [l for l in buf.getvalue().splitlines() if l.startswith("snapshot 0:")][0]
    assert snap0.count("ndiff=0 ") == 4 and "misc identical" in snap0, snap0
    assert worst <= 1e-10


def test_host_start_refuses_unknown_keys_like_the_reference(tmp_path):
    """config::Config::exit_on_unknown_key: a misspelt key must not silently fall back to a default."""
    cfg = yaml.safe_load(open(os.path.join(ROOT, "tests", "golden", "iso_star.yml")))
    cfg["ViscousAlfa"] = cfg.pop("ViscousAlpha")
    yml = str(tmp_path / "setup.yml")
    yaml.safe_dump(cfg, open(yml, "w"), sort_keys=False)
    res = subprocess.run([_oracle_exe(), "start", yml, "--out", str(tmp_path / "out")], capture_output=True, text=True, timeout=60)
    assert res.returncode != 0 and "viscousalfa" in res.stderr.lower()


def test_host_refuses_physics_it_does_not_implement(tmp_path):
    for key, value in (("EquationOfState", "Polytropic"), ("SurfaceCooling", "fld"), ("SelfGravity", "yes"), ("AlphaMode", 2),
                       ("BodyForceFromPotential", "no")):
        cfg = yaml.safe_load(open(os.path.join(ROOT, "tests", "golden", "adia_star.yml")))
        cfg[key] = value
        yml = str(tmp_path / f"setup_{key}.yml")
        yaml.safe_dump(cfg, open(yml, "w"), sort_keys=False)
        res = subprocess.run([_oracle_exe(), "start", yml, "--out", str(tmp_path / ("out_" + key))], capture_output=True, text=True, timeout=60)
        assert res.returncode != 0 and key in res.stderr, (key, res.stderr)


def test_boundary_types_are_inferred_like_the_reference(tmp_path):
    """boundary_conditions/config.cpp:75-94, 145-147: individual keys overwrite what a composite set; the inner energy type is inferred
    from the OUTER side's name, and an explicit InnerBoundaryEnergy also becomes the outer default.  Checked on the Python mirror of
    the host's parser; the side-by-side runs of tests/checkers/fuzz_against_reference.py cover the host against the reference."""
    import sys
    sys.path.insert(0, ROOT)
    from fargocpt_b200 import abi, config
    base = {k: v for k, v in yaml.safe_load(open(os.path.join(ROOT, "tests", "golden", "adia_star.yml"))).items()}
    consts = {"G": 1.0, "R": 1.0, "sigma": 1.0, "c": 1.0}

    def bc(**over):
        cfg = dict(base, InnerBoundary="individual", OuterBoundary="individual")
        cfg.update(over)
        d = config.params_from_config(cfg, consts, 24, 48)
        return d["bc_sigma"], d["bc_energy"], d["bc_vrad"]

    ZG, REF, RFL, OUT = abi.BC["zerogradient"], abi.BC["reference"], abi.BC["reflecting"], abi.BC["outflow"]
    # inner composite Reference, outer Reflecting: the inner ENERGY follows the outer composite (zerogradient)
    assert bc(InnerBoundary="Reference", OuterBoundary="Reflecting") == ([REF, ZG], [ZG, ZG], [REF, RFL])
    # an individual key overwrites the composite's choice
    assert bc(InnerBoundary="Reflecting", OuterBoundary="Reflecting", InnerBoundaryVrad="outflow")[2] == [OUT, RFL]
    # an explicit InnerBoundaryEnergy also becomes the outer side's type unless that is given too
    assert bc(InnerBoundary="Reflecting", OuterBoundary="Reflecting", InnerBoundaryEnergy="reference")[1] == [REF, REF]
    # inner composite + outer individual: the inner energy cannot be inferred (the reference throws)
    with pytest.raises(ValueError, match="InnerBoundaryEnergy"):
        bc(InnerBoundary="Reflecting", OuterBoundarySigma="zerogradient", OuterBoundaryEnergy="zerogradient", OuterBoundaryVrad="outflow")
    cfg = dict(base, InnerBoundary="Reflecting", OuterBoundary="individual")
    cfg.update(OuterBoundarySigma="zerogradient", OuterBoundaryEnergy="zerogradient", OuterBoundaryVrad="outflow")
    yml = str(tmp_path / "setup.yml")
    yaml.safe_dump(cfg, open(yml, "w"), sort_keys=False)
    res = subprocess.run([_oracle_exe(), "start", yml, "--out", str(tmp_path / "out")], capture_output=True, text=True, timeout=60)
    assert res.returncode != 0 and "InnerBoundaryEnergy" in res.stderr, res.stderr


# ---------------------------------------------------------------------------------------------------------------------
# Known answer of the reference's test/shockTube (check_results.py: integrated |numerical - analytic| over the tube, thresholds below)
def _sod_exact(x, t, gamma=1.4, left=(1.0, 1.0), right=(0.125, 0.1), x0=0.5):
    """Exact solution of Sod's Riemann problem (rho, u, p) at positions x and time t (Toro, ch. 4)."""
    rl, pl = left
    rr, pr = right
    cl, cr = np.sqrt(gamma * pl / rl), np.sqrt(gamma * pr / rr)
    g1, g2 = (gamma - 1) / (2 * gamma), (gamma + 1) / (2 * gamma)

    def f(p, rk, pk, ck):  # pressure functions and their derivatives
        if p > pk:
            a, b = 2 / ((gamma + 1) * rk), (gamma - 1) / (gamma + 1) * pk
            return (p - pk) * np.sqrt(a / (p + b)), np.sqrt(a / (b + p)) * (1 - (p - pk) / (2 * (b + p)))
        return 2 * ck / (gamma - 1) * ((p / pk) ** g1 - 1), (p / pk) ** (-g2) / (rk * ck)

    p = 0.5 * (pl + pr)
    for _ in range(60):
        fl, dfl = f(p, rl, pl, cl)
        fr, dfr = f(p, rr, pr, cr)
        p -= (fl + fr) / (dfl + dfr)
    u = 0.5 * (f(p, rr, pr, cr)[0] - f(p, rl, pl, cl)[0])
    rsl = rl * (p / pl) ** (1 / gamma)  # behind the left rarefaction
    rsr = rr * ((p / pr + (gamma - 1) / (gamma + 1)) / ((gamma - 1) / (gamma + 1) * p / pr + 1))  # behind the right shock
    csl = cl * (p / pl) ** g1
    s_shock = cr * np.sqrt(g2 * p / pr + g1)
    xi = (x - x0) / t
    rho, vel, prs = np.empty_like(x), np.empty_like(x), np.empty_like(x)
    for k, s in enumerate(xi):
        if s < -cl:
            rho[k], vel[k], prs[k] = rl, 0.0, pl
        elif s < u - csl:  # inside the fan
            c = 2 / (gamma + 1) * (cl - (gamma - 1) / 2 * s)
            vel[k] = 2 / (gamma + 1) * (cl + s)
            rho[k] = rl * (c / cl) ** (2 / (gamma - 1))
            prs[k] = pl * (c / cl) ** (2 * gamma / (gamma - 1))
        elif s < u:
            rho[k], vel[k], prs[k] = rsl, u, p
        elif s < s_shock:
            rho[k], vel[k], prs[k] = rsr, u, p
        else:
            rho[k], vel[k], prs[k] = rr, 0.0, pr
    return rho, vel, prs


@pytest.mark.parametrize("integrator,artvisc", [("Euler", "TW"), ("Euler", "SN"), ("Leapfrog", "TW"), ("Leapfrog", "SN")])
def test_shock_tube_matches_the_exact_riemann_solution_cpu(integrator, artvisc, tmp_path):
    """The reference's test/shockTube acceptance (check_results.py:15-20, its four setups TW / SN x Euler / Leapfrog):
    Simpson-integrated absolute deviation from the exact Sod solution at t = 0.228 (membrane at x = 0.5 as in the reference's
    analytic_shock.dat) below 0.0073 (Sigma), 0.0153 (v_rad), 0.014 (energy), 0.016 (temperature)."""
    from scipy import integrate
    cfg = yaml.safe_load(open(os.path.join(ROOT, "tests", "golden", "shock_tube_setup.yml")))
    cfg["Integrator"], cfg["ArtificialViscosity"] = integrator, artvisc
    yml, out = str(tmp_path / "setup.yml"), str(tmp_path / "out")
    yaml.safe_dump(cfg, open(yml, "w"), sort_keys=False)
    _run_start(_oracle_exe(), yml, out, 1)
    r12 = np.loadtxt(os.path.join(out, "used_rad.dat"))
    r1 = 0.5 * (r12[1:] + r12[:-1]) - r12[0]
    nr = len(r1)
    sig = np.fromfile(os.path.join(out, "snapshots", "1", "Sigma.dat")).reshape(nr, 2).mean(axis=1)
    en = np.fromfile(os.path.join(out, "snapshots", "1", "energy.dat")).reshape(nr, 2).mean(axis=1)
    vr = np.fromfile(os.path.join(out, "snapshots", "1", "vrad.dat")).reshape(nr + 1, 2).mean(axis=1)
    vr = 0.5 * (vr[1:] + vr[:-1])
    inside = (r1 >= 0) & (r1 <= 1)
    x = r1[inside]
    rho, vel, prs = _sod_exact(x, 0.228)
    dev = {"Sigma": integrate.simpson(np.abs(sig[inside] - rho), x=x), "vrad": integrate.simpson(np.abs(vr[inside] - vel), x=x),
           "energy": integrate.simpson(np.abs(en[inside] - prs / 0.4), x=x),
           "Temperature": integrate.simpson(np.abs(0.4 * en[inside] / sig[inside] - prs / rho), x=x)}
    limits = {"vrad": 0.0153, "Sigma": 0.0073, "Temperature": 0.016, "energy": 0.014}
    for q, lim in limits.items():
        assert dev[q] < lim, (q, dev[q], lim)


@pytest.mark.parametrize("setup", ["shocktube_TW.yml", "shocktube_SN.yml", "shocktube_TW_LF.yml", "shocktube_SN_LF.yml"])
def test_shock_tube_setups_of_the_reference_verbatim_cpu(setup):
    """test/shockTube/setups/*.yml through the unmodified reference and through `fargocpt_b200 start`: ~150 CFL-limited hydro steps
    with strong shocks (TW / SN artificial viscosity, Euler / Leapfrog), every double of every snapshot identical."""
    path = os.path.join("/root/reference/test/shockTube/setups", setup)
    if not (os.path.exists(path) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "fargocpt_exe_ieee"))):
        pytest.skip("the reference tree / oracle/_ref are not available here")
    _oracle_exe()
    import contextlib
    import importlib.util
    import io
    spec = importlib.util.spec_from_file_location("cmpstart", os.path.join(ROOT, "tests", "checkers", "compare_start_with_reference.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        worst = mod.main([path, "--snapshots", "3", "--dt", "0.01"])
    assert worst == 0.0 and buf.getvalue().count("misc identical") == 4, buf.getvalue()


def test_spreading_ring_meets_the_references_acceptance_cpu(tmp_path):
    """test/spreading_ring/calc_deviation.py:38-66: after t = 314.159 (39 870 hydro steps on the 256 x 2 grid) the ring must
    follow the analytic viscous-spreading solution (Speith & Kley 2003) with a mean relative deviation below 0.007."""
    from scipy.special import iv
    cfg = yaml.safe_load(open(SPREADING_RING))
    cfg["MonitorTimestep"], cfg["Nsnapshots"] = 314.159265359, 1
    yml, out = str(tmp_path / "setup.yml"), str(tmp_path / "out")
    yaml.safe_dump(cfg, open(yml, "w"), sort_keys=False)
    _run_start(_oracle_exe(), yml, out, 1)
    ri = np.loadtxt(os.path.join(out, "used_rad.dat"))
    rinf, rsup = ri[:-1], ri[1:]
    rc = 2.0 / 3.0 * (rsup ** 3 - rinf ** 3) / (rsup ** 2 - rinf ** 2)
    sigma = np.fromfile(os.path.join(out, "snapshots", "1", "Sigma.dat")).reshape(256, 2).mean(axis=1)
    raw = open(os.path.join(out, "snapshots", "1", "misc.bin"), "rb").read()
    t = struct.unpack("<IIddddQ", raw)[2]
    tau = 12 * 4.77e-5 * t + 0.016
    theo = 1.0 / np.pi / tau / rc ** 0.25 * iv(0.25, 2.0 * rc / tau) * np.exp(-(1 + rc ** 2) / tau)
    assert np.mean(np.abs(sigma / theo - 1)) < 0.007


def test_host_writes_derived_fields_on_request_cpu(tmp_path):
    """WriteTemperature / WritePressure / WriteSoundSpeed: the optional derived outputs of a snapshot (data.cpp), evaluated from the
    snapshot's own state (identical to the reference's files, tests/checkers/compare_start_with_reference.py --keep)."""
    cfg = yaml.safe_load(open(os.path.join(ROOT, "tests", "golden", "adia_star.yml")))
    cfg.update({"WriteTemperature": "yes", "WritePressure": "yes", "WriteSoundSpeed": "yes"})
    yml, out = str(tmp_path / "setup.yml"), str(tmp_path / "out")
    yaml.safe_dump(cfg, open(yml, "w"), sort_keys=False)
    _run_start(_oracle_exe(), yml, out, 2)
    meta, z = reftools.load_golden("adia_star")
    sd = os.path.join(out, "snapshots", "2")
    sigma, energy = np.fromfile(os.path.join(sd, "Sigma.dat")), np.fromfile(os.path.join(sd, "energy.dat"))
    gamma, mu, rgas = meta["params"]["gamma"], meta["params"]["mu"], meta["consts"]["R"]
    pressure = np.fromfile(os.path.join(sd, "pressure.dat"))
    assert np.array_equal(pressure, (gamma - 1.0) * energy)
    assert np.allclose(np.fromfile(os.path.join(sd, "Temperature.dat")), mu / rgas * pressure / sigma, rtol=1e-15, atol=0)
    assert np.allclose(np.fromfile(os.path.join(sd, "soundspeed.dat")), np.sqrt(gamma * (gamma - 1.0) * energy / sigma), rtol=1e-15, atol=0)
    assert not os.path.exists(os.path.join(sd, "viscosity.dat"))


@pytest.mark.parametrize("setup,extra", [("/root/reference/test/cold_disk_planet/setup.yml", []),
                                         ("/root/reference/examples/config.yml", ["Integrator=Leapfrog"]),
                                         # a corotating frame: the reference's own restart is not seamless there (init_corotation takes the
                                         # restart position as the frame's reference position, frame_of_reference.cpp:19-28, so its first
                                         # step after a restart sees OmegaFrame = 0) — the driver reproduces the REFERENCE'S RESTART, quirk included
                                         (os.path.join(ROOT, "tests", "golden", "corotating_setup.yml"), ["--dt", "4e-3", "--vs-reference-restart"]),
                                         ("/root/reference/test/cold_disk_planet/setup.yml", ["--vs-reference-restart"]),
                                         # values with units in the snapshot's config.yml (a verbatim copy of the setup): `restart` runs the same
                                         # unit conversion as `start` (l0 = 30 au: Rmin = 0.4, Rmax = 2)
                                         ("/root/reference/test/cold_disk_planet/setup.yml", ["Rmin=12 au", "Rmax=60 au"]),
                                         # restarted while dt is still limited by CFLmaxVar * last_dt (default FirstDT 1e-9): the reference skips
                                         # sim::init's CalculateTimeStep when restarting (simulation.cpp:465), so the limiter acts once per step
                                         (os.path.join(ROOT, "tests", "golden", "minimal_defaults_setup.yml"), ["FirstDT=1e-9", "--vs-reference-restart"]),
                                         # a unit on a key whose dimension is L0^2 / T0 (Interpret.cpp:586)
                                         ("/root/reference/test/cold_disk_planet/setup.yml", ["ConstantViscosity=1e15 cm2/s", "ViscousAlpha=0"])])
def test_host_restarts_from_a_directory_the_reference_wrote(setup, extra):
    """`restart 2 <dir>` on an output directory written by the unmodified reference itself (its real constants.yml, units.yml,
    dimensions.dat, 256-byte nbody records, misc.bin, snapshots/reference): snapshots 3 and 4 against the reference's own."""
    if not (os.path.exists(setup) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "fargocpt_exe_ieee"))):
        pytest.skip("the reference tree / oracle/_ref are not available here")
    _oracle_exe()
    import contextlib
    import importlib.util
    import io
    spec = importlib.util.spec_from_file_location("cmpstart", os.path.join(ROOT, "tests", "checkers", "compare_start_with_reference.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        worst = mod.main([setup, "--snapshots", "4", "--dt", "1e-3", "--restart-from", "2"] + extra)
    text = buf.getvalue()
    assert worst <= 1e-10 and sum(l.startswith("snapshot ") for l in text.splitlines()) == 2, text


@pytest.mark.parametrize("k", [2, 3, 4])
def test_baseline_config_setups_against_the_reference_at_reduced_resolution(k, tmp_path):
    """tests/golden/baseline_config{2,3,4}_setup.yml are BASELINE.json's configs[2..4] as setup files for `fargocpt_b200 start`
    (2048 x 4096, 4096 x 8192, 8192 x 16384).  Here: the same physics at 1/16 of the resolution per direction through the
    reference and the driver."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "fargocpt_exe_ieee")):
        pytest.skip("oracle/_ref is not available here")
    _oracle_exe()
    cfg = yaml.safe_load(open(os.path.join(ROOT, "tests", "golden", f"baseline_config{k}_setup.yml")))
    assert (cfg["Nrad"], cfg["Naz"]) == {2: (2048, 4096), 3: (4096, 8192), 4: (8192, 16384)}[k]
    shrink = 16 if k < 4 else 64
    cfg["Nrad"], cfg["Naz"] = cfg["Nrad"] // shrink, cfg["Naz"] // shrink
    yml = str(tmp_path / "setup.yml")
    yaml.safe_dump(cfg, open(yml, "w"), sort_keys=False)
    import contextlib
    import importlib.util
    import io
    spec = importlib.util.spec_from_file_location("cmpstart", os.path.join(ROOT, "tests", "checkers", "compare_start_with_reference.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        worst = mod.main([yml, "--snapshots", "2", "--dt", "1e-3"])
    snap0 = [l for l in buf.getvalue().splitlines() if l.startswith("snapshot 0:")][0]
    assert "misc identical" in snap0 and snap0.count("ndiff=0 ") >= 3, snap0
    assert worst <= 1e-10, buf.getvalue()


# ---------------------------------------------------------------------------------------------------------------------
# The product driver (GPU) against the oracle-bound driver: the same host code on two backends whose hydro results agree bit for
# bit, so every file they write must be identical — frames, circumbinary disks, initial-condition variants, monitor files.
GPU_VS_ORACLE_DRIVER = [
    ("circumbinary_setup", {}),
    # an isothermal run that carries an energy grid (the Gaussian ring): it must reach the snapshot through fargo_snapshot_async
    ("circumbinary_setup", {"CircumBinaryRing": "yes", "CircumBinaryRingPosition": 1.5, "CircumBinaryRingWidth": 0.2, "SigmaCondition": "Nbody",
                            "SetSigma0": "yes", "DiskMass": 0.01}),
    ("circumbinary_setup", {"HydroFrameCenter": "all", "Integrator": "Leapfrog", "DiskFeedback": "yes", "IndirectTermMode": 0}),
    ("corotating_setup", {"DiskFeedback": "yes", "IndirectTermMode": 0}),
    ("multi_body_setup", {}),
    ("minimal_defaults_setup", {}),
    ("adia_irrad_lf", {"WriteTemperature": "yes", "WritePressure": "yes", "WriteSoundSpeed": "yes", "WriteScaleHeight": "yes", "WriteViscosity": "yes"}),
    ("adia_pvte", {"WriteTemperature": "yes", "WriteScaleHeight": "yes"}),
]


@pytest.mark.gpu
@pytest.mark.parametrize("setup,over", GPU_VS_ORACLE_DRIVER)
def test_gpu_driver_writes_what_the_oracle_bound_driver_writes(setup, over, tmp_path):
    cfg = {k: v for k, v in yaml.safe_load(open(os.path.join(ROOT, "tests", "golden", setup + ".yml"))).items() if not k.startswith("_")}
    cfg.update({"MonitorTimestep": 2e-3, "Nmonitor": 1, "Nsnapshots": 3, "WriteAtEveryTimestep": "yes"})
    cfg.update(over)
    yml = str(tmp_path / "setup.yml")
    yaml.safe_dump(cfg, open(yml, "w"), sort_keys=False)
    outs = {}
    for kind, exe in (("gpu", os.path.join(ROOT, "host", "fargocpt_b200")), ("oracle", _oracle_exe())):
        outs[kind] = str(tmp_path / kind)
        _run_start(exe, yml, outs[kind], 3)
    nfiles = 0
    for dirpath, _, files in os.walk(outs["oracle"]):
        for f in files:
            po = os.path.join(dirpath, f)
            pg = os.path.join(outs["gpu"], os.path.relpath(po, outs["oracle"]))
            assert os.path.exists(pg), pg
            if f == "timestepLogging.dat" or f.startswith(("nbody", "Quantities")) and f.endswith(".dat") and "monitor" in dirpath:
                # monitor sums: device reductions against the oracle's sequential sums agree to rounding, not to the bit
                a = np.array([[float(x) for x in l.split()] for l in open(po) if not l.startswith("#") and l.strip()])
                b = np.array([[float(x) for x in l.split()] for l in open(pg) if not l.startswith("#") and l.strip()])
                assert a.shape == b.shape, f
                assert np.array_equal(np.isnan(a), np.isnan(b)), f
                a, b = np.nan_to_num(a), np.nan_to_num(b)
                if f.startswith("Quantities"):
                    # disk eccentricity (column 12): a mean of O(h^2) cell values that cancel around an axisymmetric disk —
                    # absolute; its periastron (13) is the angle of that mean: compared only where there is an eccentricity
                    assert np.max(np.abs(a[:, 12] - b[:, 12])) < 1e-13, f
                    has_ecc = a[:, 12] > 1e-9
                    dper = np.abs(np.angle(np.exp(1j * (a[:, 13] - b[:, 13]))))
                    assert not has_ecc.any() or np.max(dper[has_ecc]) < 1e-6, f
                    a[:, 12:14] = b[:, 12:14] = 0.0
                scale = np.maximum(np.abs(a).max(axis=0), 1e-300)
                if f.startswith("Quantities"):
                    # torques of an axisymmetric disk are sums that cancel to rounding noise: measured against the disk's mass
                    scale[32:35] = np.maximum(scale[32:35], 1e-6 * np.abs(a[:, 3]).max())
                assert np.max(np.abs(a - b) / scale) < 1e-9, f
                continue
            rel = os.path.relpath(po, outs["oracle"])
            a, b = open(po, "rb").read(), open(pg, "rb").read()
            feedback = str(cfg.get("DiskFeedback", "no")).lower().startswith("y")
            if f.startswith("nbody") and f.endswith(".bin"):
                # the record carries the disk's pull on the body (offsets 120-152) and torque sums: device reductions, rounding-level
                x, y = np.frombuffer(a[8:72]), np.frombuffer(b[8:72])  # mass, x, y, vx, vy, smoothing, accretion efficiency, accreted mass
                assert np.allclose(x, y, rtol=1e-12 if feedback else 0.0, atol=0.0), (rel, x, y)
            elif feedback and f.endswith(".dat") and "snapshots" in dirpath:
                # the reduced pull moves the bodies, so the gas follows to rounding
                x, y = np.nan_to_num(np.frombuffer(a)), np.nan_to_num(np.frombuffer(b))
                assert x.shape == y.shape and np.abs(x - y).max() <= 1e-12 * max(np.abs(x).max(), 1e-300), rel
            elif f == "misc.bin" and feedback:
                assert np.allclose(np.frombuffer(a[8:40]), np.frombuffer(b[8:40]), rtol=1e-12), rel
            else:
                assert a == b, rel
            nfiles += 1
    assert nfiles >= 20
