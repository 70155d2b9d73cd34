#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE
(oracle/_ref/fargocpt_exe_ieee, built by `make -f oracle/Makefile.ref` from /root/reference).

Runs only in the build container (needs oracle/_ref); the .npz outputs are committed.
Each case = one small FargoCPT YAML config.  To capture the state after EVERY hydro step the
configs use Nmonitor: 1 and a MonitorTimestep smaller than the CFL dt, so every hydro step is
exactly one monitor step and one snapshot (simulation.cpp:528-550).

A fixture holds: the YAML text, the derived FargoParams (as JSON), used_rad.dat, and for every
snapshot k the raw arrays Sigma/vrad/vazi/energy(/Qplus/Qminus), misc.bin (time, last_dt, N_iter)
and each body's (mass, x, y, vx, vy).
"""
import json
import os
import shutil
import struct
import subprocess
import sys
import tempfile

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import reftools  # noqa: E402

EXE = os.path.join(ROOT, "oracle", "_ref", "fargocpt_exe_ieee")

BASE = {
    "DiskFeedback": "no", "MonitorTimestep": 1.0e-3, "Nmonitor": 1, "Nsnapshots": 6, "FirstDT": 1.0e-1,
    "CFLmaxVar": 1.1, "ShockTube": 0, "Sigma0": 0.005743125733951172, "SigmaSlope": 1.0, "SigmaFloor": 1e-7,
    "AspectRatio": 0.05, "FlaringIndex": 0.2857142857142857, "ViscousAlpha": 0.0, "ArtificialViscosity": "TW",
    "ArtificialViscosityDissipation": "Yes", "ArtificialViscosityFactor": 3.0, "SelfGravity": "No",
    "EquationOfState": "Ideal", "AdiabaticIndex": 1.4, "CoolingBetaLocal": "No", "CoolingBetaReference": "reference",
    "CoolingBeta": 100, "RadiativeDiffusion": "No", "HeatingViscous": "no", "MinimumTemperature": "3 K",
    "MaximumTemperature": "1e100 K", "CFL": 0.5, "HeatingCoolingCFLlimit": 1.0,
    "l0": "30 au", "m0": "1 solMass", "mu": 2.35, "ThicknessSmoothing": 0.6, "Transport": "FARGO",
    "Integrator": "Euler", "IndirectTermMode": 0, "InnerBoundary": "Reflecting", "OuterBoundary": "Reflecting",
    "Damping": "No", "Disk": "yes", "OmegaFrame": 0, "Frame": "F",
    "Nrad": 24, "Naz": 48, "cps": -1, "Rmin": 0.4, "Rmax": 2.0, "RadialSpacing": "Logarithmic",
    "DoWrite1DFiles": "No", "WriteAtEveryTimestep": "Yes", "WriteDensity": "Yes", "WriteEnergy": "Yes",
    "WriteVelocity": "Yes", "WriteQMinus": "Yes", "WriteQPlus": "Yes", "WriteDiskQuantities": "No",
    "RandomSigma": "No", "IntegrateParticles": "no", "HydroFrameCenter": "primary", "BodyForceFromPotential": "Yes",
    "LogAfterSteps": 0, "LogAfterRealSeconds": 600,
    "nbody": [{"name": "Star", "semi-major axis": 0.0, "mass": "1 solMass", "eccentricity": 0,
               "radius": "1 solRadius", "temperature": 0}],
}
DAMP_ALL = {k: "Initial" for k in (
    "DampingEnergyInner", "DampingVRadialInner", "DampingVAzimuthalInner", "DampingSurfaceDensityInner",
    "DampingEnergyOuter", "DampingVRadialOuter", "DampingVAzimuthalOuter", "DampingSurfaceDensityOuter")}

CASES = {
    # adiabatic, TW art-visc + dissipation, alpha viscosity + viscous heating + beta cooling, reflecting + damping
    "adia_star": dict(ViscousAlpha=1e-3, HeatingViscous="yes", CoolingBetaLocal="yes", CoolingBeta=10,
                      Damping="Yes", DampingInnerLimit=1.311, DampingOuterLimit=0.763, DampingTimeFactor=0.05,
                      **DAMP_ALL),
    # cold_disk-like: adiabatic, no physical viscosity, no heating/cooling (test/cold_disk/setup.yml physics)
    "adia_cold": dict(Damping="Yes", DampingInnerLimit=1.311, DampingOuterLimit=0.763, DampingTimeFactor=0.05,
                      **DAMP_ALL),
    # locally isothermal, alpha viscosity, TW, outflow boundaries, corotating-frame style OmegaFrame != 0
    "iso_star": dict(EquationOfState="Isothermal", ViscousAlpha=1e-2, ArtificialViscosityFactor=1.41,
                     InnerBoundary="Outflow", OuterBoundary="Outflow", OmegaFrame=1.0, FlaringIndex=0.0),
    # locally isothermal, Stone-Norman art-visc, constant viscosity, standard (non-FARGO) transport, MC limiter,
    # zero-gradient boundaries, arithmetic grid
    "iso_sn_std": dict(EquationOfState="Isothermal", ArtificialViscosity="SN", ArtificialViscosityFactor=1.41,
                       ConstantViscosity=1.0e-5, Transport="Standard", FluxLimiter="mc", MonitorTimestep=2.0e-4,
                       InnerBoundary="ZeroGradient", OuterBoundary="ZeroGradient", RadialSpacing="Arithmetic"),
    # adiabatic with SN art-visc + dissipation, StabilizeViscosity 1, damping to zero / mean
    "adia_sn_stab": dict(ArtificialViscosity="SN", ArtificialViscosityFactor=1.41, ViscousAlpha=5e-3,
                         StabilizeViscosity=1, HeatingViscous="yes", InnerBoundary="Outflow",
                         OuterBoundary="Reflecting", Damping="Yes", DampingInnerLimit=1.2, DampingOuterLimit=0.8,
                         DampingTimeFactor=0.1, DampingVRadialInner="Zero", DampingVRadialOuter="Zero",
                         DampingSurfaceDensityOuter="Mean", DampingEnergyOuter="Mean"),
    # pressureless viscous ring geometry of test/spreading_ring (Naz = 2): art-visc None, h = 0 disabled here
    # because the Bessel initial condition is out of scope; same grid/viscosity/boundaries on a power-law disk
    "ring_like": dict(EquationOfState="Isothermal", AspectRatio=0.01, FlaringIndex=0.0, ArtificialViscosity="None",
                      ArtificialViscosityFactor=1.41, ConstantViscosity=4.77e-5, InnerBoundary="Outflow",
                      OuterBoundary="Outflow", Nrad=64, Naz=2, Rmin=0.2, Rmax=1.8, MonitorTimestep=1.0e-4,
                      ThicknessSmoothing=0.0),
    # stress::calculate_Reynolds_stress (stress.cpp:34-70): T_Reynolds.dat is filled when the alpha-Reynolds output runs
    "rey_star": dict(ViscousAlpha=1e-3, HeatingViscous="yes", CoolingBetaLocal="yes", CoolingBeta=10, Nsnapshots=2,
                     WriteAlphaReynolds="yes", WriteTReynolds="yes", _planet=3e-4, IndirectTermMode=1),
    # Integrator: Leapfrog (step_LeapFrog simulation.cpp:276-459: kick dt/2, drift dt, kick dt/2; CFL factor 0.6)
    "adia_leapfrog": dict(Integrator="Leapfrog", ViscousAlpha=1e-3, HeatingViscous="yes", CoolingBetaLocal="yes", CoolingBeta=10,
                          Damping="Yes", DampingInnerLimit=1.311, DampingOuterLimit=0.763, DampingTimeFactor=0.05, **DAMP_ALL),
    # DiskFeedback: yes — the disk's pull on star and planet (Force.cpp:23-122) enters the bodies' velocities and the
    # indirect term every step (simulation.cpp:155-165)
    "iso_feedback_20": dict(Nrad=48, Naz=160, Rmax=2.5, Nsnapshots=20, MonitorTimestep=4.0e-3, IndirectTermMode=1,
                            EquationOfState="Isothermal", ViscousAlpha=1e-3, ArtificialViscosityFactor=1.41,
                            FlaringIndex=0.0, DiskFeedback="yes", _planet=1e-3, _keep=(0, 20)),
    # 100 hydro steps with a Jupiter-mass planet (north_star: fields <= 1e-10 after 100 steps; dt and Nshift bit-exact).
    # IndirectTermMode 1 (Euler): the indirect term is a closed formula of the recorded body states
    # (frame_of_reference.cpp:112-132, Pframeforce.cpp:225-251) which tests/goldenrun.py restates.
    "adia_planet_100": dict(Nrad=48, Naz=160, Rmax=2.5, Nsnapshots=100, MonitorTimestep=4.0e-3, IndirectTermMode=1,
                            ViscousAlpha=1e-3, HeatingViscous="yes", CoolingBetaLocal="yes", CoolingBeta=10,
                            Damping="Yes", DampingInnerLimit=1.25, DampingOuterLimit=0.84, DampingTimeFactor=0.1,
                            _planet=1e-3, _keep=(0, 50, 100), **DAMP_ALL),
    # accretion::AccreteOntoSinglePlanet (accretion.cpp:84-221): a planet that accretes ("kley", efficiency 5 per orbit) but
    # does not feel the disk (DiskFeedback: no), so the only feedback of the accretion is the gas it removes
    "iso_accrete_20": dict(Nrad=48, Naz=160, Rmax=2.5, Nsnapshots=20, MonitorTimestep=4.0e-3, IndirectTermMode=1,
                           EquationOfState="Isothermal", ViscousAlpha=1e-3, ArtificialViscosityFactor=1.41, FlaringIndex=0.0,
                           _planet=3e-3, _accretion=5.0, _keep=(0, 10, 20)),
    "adia_accrete_20": dict(Nrad=48, Naz=160, Rmax=2.5, Nsnapshots=20, MonitorTimestep=4.0e-3, IndirectTermMode=1,
                            ViscousAlpha=1e-3, HeatingViscous="yes", CoolingBetaLocal="yes", CoolingBeta=10,
                            _planet=3e-3, _accretion=5.0, _keep=(0, 10, 20)),
    # "accretion method: sinkhole" (accretion.cpp:223-333): one zone of radius MassAccretionRadius * R_Hill
    "iso_sinkhole_20": dict(Nrad=48, Naz=160, Rmax=2.5, Nsnapshots=20, MonitorTimestep=4.0e-3, IndirectTermMode=1,
                            EquationOfState="Isothermal", ViscousAlpha=1e-3, ArtificialViscosityFactor=1.41, FlaringIndex=0.0,
                            MassAccretionRadius=0.75, _planet=3e-3, _accretion=5.0, _accretion_method="sinkhole", _keep=(0, 10, 20)),
    # "accretion method: viscous" (accretion.cpp:335-480): the removed fraction scales with the local viscosity and a cone profile
    "adia_viscacc_20": dict(Nrad=48, Naz=160, Rmax=2.5, Nsnapshots=20, MonitorTimestep=4.0e-3, IndirectTermMode=1,
                            ViscousAlpha=1e-2, HeatingViscous="yes", CoolingBetaLocal="yes", CoolingBeta=10,
                            _planet=3e-3, _accretion=2.0e3, _accretion_method="viscous", _keep=(0, 10, 20)),
    # an accreting planet that feels the disk: update_planet (accretion.cpp:60-82) adds the accreted mass and momentum to it
    "adia_accfb_20": dict(Nrad=48, Naz=160, Rmax=2.5, Nsnapshots=20, MonitorTimestep=4.0e-3, IndirectTermMode=1,
                          ViscousAlpha=1e-3, HeatingViscous="yes", CoolingBetaLocal="yes", CoolingBeta=10, DiskFeedback="yes",
                          _planet=3e-3, _accretion=5.0, _keep=(0, 10, 20)),
    # SurfaceCooling: thermal + an irradiating star (SourceEuler.cpp:538-723, compute.cpp:17-88) with a constant opacity, as
    # test/irradiation sets them up (there with Naz = 2); viscous heating on, no beta cooling
    "adia_irrad": dict(ViscousAlpha=1e-3, HeatingViscous="yes", SurfaceCooling="thermal", Opacity="Constant", KappaConst="2.0e-6",
                       TauFactor=1.0, HeatingCoolingCFLlimit=1000.0, ArtificialViscosity="None", ArtificialViscosityDissipation="No",
                       _star={"temperature": "10000 K", "irradiation ramp-up time": 5.0e-3}),
    # the same physics through step_LeapFrog, where SubStep3 runs in both kicks
    "adia_irrad_lf": dict(Integrator="Leapfrog", ViscousAlpha=1e-3, HeatingViscous="yes", SurfaceCooling="thermal", Opacity="Constant",
                          KappaConst="2.0e-6", TauFactor=1.0, _star={"temperature": "10000 K"}),
    # thermal cooling alone (the non-irradiated tau_eff) through the Lin & Papaloizou and the Bell & Lin opacity tables
    "adia_cool_lin": dict(ViscousAlpha=1e-3, HeatingViscous="yes", SurfaceCooling="thermal", Opacity="Lin"),
    "adia_cool_bell": dict(ViscousAlpha=1e-3, HeatingViscous="yes", SurfaceCooling="thermal", Opacity="Bell", KappaFactor=2.0,
                           _planet=3e-4, IndirectTermMode=1),
    # EquationOfState: PVTE (pvte_law.cpp): gamma_eff, mu, Gamma_1 per cell from the lookup tables; the three grids are recorded
    "adia_pvte": dict(EquationOfState="PVTE", ViscousAlpha=1e-3, HeatingViscous="yes", CoolingBetaLocal="yes", CoolingBeta=10,
                      WriteEffectiveGamma="yes", WriteFirstAdiabaticIndex="yes", WriteMeanMolecularWeight="yes", WriteScaleHeight="yes",
                      Damping="Yes", DampingInnerLimit=1.311, DampingOuterLimit=0.763, DampingTimeFactor=0.05, **DAMP_ALL),
    # PVTE with Leapfrog: the second kick refreshes c_s, H and the lookup after its potential (simulation.cpp:368-376)
    "adia_pvte_lf": dict(EquationOfState="PVTE", Integrator="Leapfrog", ViscousAlpha=1e-3, HeatingViscous="yes", CoolingBetaLocal="yes", CoolingBeta=10,
                         WriteEffectiveGamma="yes", WriteFirstAdiabaticIndex="yes", WriteMeanMolecularWeight="yes", WriteScaleHeight="yes"),
    # AlphaMode 1: S-curve alpha in the (stored) temperature (viscosity.cpp:36-49); l0 = 0.06 au puts the disk around 1e4 K,
    # where alpha switches between AlphaCold and AlphaHot
    "adia_alpha_scurve": dict(AlphaMode=1, ViscousAlpha=1e-3, AlphaCold=0.01, AlphaHot=0.1, HeatingViscous="yes", CoolingBetaLocal="yes",
                              CoolingBeta=10, l0="0.06 au", Sigma0=0.001, WriteTemperature="yes", WriteViscosity="yes",
                              Damping="Yes", DampingInnerLimit=1.311, DampingOuterLimit=0.763, DampingTimeFactor=0.05, **DAMP_ALL),
    "adia_alpha_scurve_lf": dict(Integrator="Leapfrog", AlphaMode=1, ViscousAlpha=1e-3, AlphaCold=0.01, AlphaHot=0.1, HeatingViscous="yes",
                                 l0="0.06 au", Sigma0=0.001, WriteTemperature="yes"),
    # v_azi boundaries: Balanced (balanced.cpp: equilibrium rotation in the ghost rings) and ZeroShear (zero_shear.cpp), also in a
    # rotating frame
    "iso_bc_balanced": dict(EquationOfState="Isothermal", InnerBoundaryVazi="Balanced", OuterBoundaryVazi="Balanced", ViscousAlpha=1e-3,
                            ThicknessSmoothing=0.4, OmegaFrame=0.3),
    "adia_bc_zeroshear": dict(InnerBoundaryVazi="ZeroShear", OuterBoundaryVazi="ZeroShear", ViscousAlpha=1e-3, HeatingViscous="yes"),
    # inner v_rad boundaries Viscous (viscous.cpp: outflow at 1.5 s nu / r) and Keplerian (keplerian_radial.cpp)
    "iso_bc_viscous": dict(EquationOfState="Isothermal", InnerBoundary="individual", InnerBoundarySigma="zerogradient",
                           InnerBoundaryEnergy="zerogradient", InnerBoundaryVrad="viscous", ViscousOutflowSpeed=5.0, ViscousAlpha=1e-2),
    "adia_bc_keplerian_vrad": dict(InnerBoundary="individual", InnerBoundarySigma="zerogradient", InnerBoundaryEnergy="zerogradient",
                                   InnerBoundaryVrad="keplerian", InnerBoundaryVradKeplerianFactor=-0.01, ViscousAlpha=0.0, ConstantViscosity=1e-5),
    # SurfaceCooling: scurve (scurve_cooling, SourceEuler.cpp:726-831) together with the S-curve alpha: a dwarf-nova disk
    "adia_scurve": dict(SurfaceCooling="scurve", ScurveType="Kimura", AlphaMode=1, ViscousAlpha=1e-3, AlphaCold=0.01, AlphaHot=0.1,
                        HeatingViscous="yes", l0="0.06 au", Sigma0=0.001, WriteTemperature="yes", WriteQminus="yes", WriteQplus="yes"),
    "adia_scurve_ichikawa_lf": dict(SurfaceCooling="scurve", ScurveType="Ichikawa", Integrator="Leapfrog", ViscousAlpha=1e-3,
                                    HeatingViscous="yes", l0="0.06 au", Sigma0=0.001, WriteTemperature="yes"),
    # damping towards the ring mean leaves the mean in column 0 of the initial-value grid (damping.cpp:578-585, 706-713), which the
    # Reference boundaries and the beta cooling towards the reference state read afterwards: all of it in one run, with a planet
    "adia_damp_mean_ref": dict(InnerBoundary="Reference", OuterBoundary="Reference", InnerBoundaryVazi="Reference", OuterBoundaryVazi="Reference",
                               ViscousAlpha=1e-2, HeatingViscous="yes", CoolingBetaLocal="yes", CoolingBeta=1, CoolingBetaReference="reference",
                               Damping="Yes", DampingInnerLimit=1.6, DampingOuterLimit=0.6, DampingTimeFactor=0.01, IndirectTermMode=1,
                               Nsnapshots=12, _planet=1e-3, **{k: "Mean" for k in DAMP_ALL}),
    "iso_planet_100": dict(Nrad=48, Naz=160, Rmax=2.5, Nsnapshots=100, MonitorTimestep=4.0e-3, IndirectTermMode=1,
                           EquationOfState="Isothermal", ViscousAlpha=1e-3, ArtificialViscosityFactor=1.41, OmegaFrame=1.0,
                           FlaringIndex=0.0, Damping="Yes", DampingInnerLimit=1.25, DampingOuterLimit=0.84,
                           DampingTimeFactor=0.1, _planet=1e-3, _keep=(0, 50, 100), **DAMP_ALL),
}


def parse_constants(outdir):
    c = yaml.safe_load(open(os.path.join(outdir, "constants.yml")))
    u = yaml.safe_load(open(os.path.join(outdir, "units.yml")))
    consts = {v["symbol"]: float(v["code value"]) for v in c.values()}
    for sym in ("sigma", "G"):  # the S-curve cooling fit is written in cgs (SourceEuler.cpp:726-831)
        consts[sym + "_cgs"] = [float(v["cgs value"]) for v in c.values() if v["symbol"] == sym][0]
    return consts, float(u["temperature"]["cgs value"]), {k: float(u[k]["cgs value"]) for k in ("density", "opacity", "energy surface density", "mass surface density",
                                                                                                "length", "mass", "energy flux")}


def read_misc(path):
    raw = open(path, "rb").read()
    timestep, ntimestep, time, omega_frame, frame_angle, last_dt, n_iter = struct.unpack("<IIddddQ", raw[:48])
    return dict(time=time, omega_frame=omega_frame, frame_angle=frame_angle, last_dt=last_dt, n_iter=n_iter)


def read_body(path):
    raw = open(path, "rb").read()
    mass, x, y, vx, vy = struct.unpack("<5d", raw[8:48])  # planet_member_variables (nbody/planet.h:11-17)
    dax, day = struct.unpack("<2d", raw[120:136])  # m_disk_on_planet_acceleration (:27), refreshed at every monitor output
    # what accretion::AccreteOntoSinglePlanet reads off the body (accretion.cpp:104-118): accretion efficiency (:20),
    # accreted mass so far (:21), distance to the primary and dimensionless Roche radius (:31-32), semi-major axis (:36)
    acc_eff, accreted = struct.unpack("<2d", raw[56:72])
    dist_primary, roche = struct.unpack("<2d", raw[152:168])
    (semi_major,) = struct.unpack("<d", raw[176:184])
    # irradiation_single (SourceEuler.cpp:538-564): temperature and radius (:22-23), irradiation ramp-up time (:25), cubic
    # smoothing factor (:18)
    temperature, radius = struct.unpack("<2d", raw[80:96])
    (irr_rampup,) = struct.unpack("<d", raw[104:112])
    (cubic,) = struct.unpack("<d", raw[48:56])
    return [mass, x, y, vx, vy, dax, day, acc_eff, accreted, dist_primary, roche, semi_major, temperature, radius, irr_rampup, cubic]


def run_case(name, overrides, keep=False):
    cfg = dict(BASE)
    cfg.update(overrides)
    planet = cfg.pop("_planet", 0.0)
    star = cfg.pop("_star", None)
    if star:
        cfg["nbody"] = [dict(cfg["nbody"][0], **star)]
    keep_snaps = cfg.pop("_keep", None)
    accretion = cfg.pop("_accretion", 0.0)
    accretion_method = cfg.pop("_accretion_method", "kley")
    if planet > 0:
        cfg["nbody"] = list(cfg["nbody"]) + [{"name": "planet", "semi-major axis": 1.0, "mass": float(planet),
                                               "accretion efficiency": float(accretion), "eccentricity": 0.0, "radius": "0.01 solRadius",
                                               "temperature": "0 K", "ramp-up time": 0}]
        if accretion > 0:
            cfg["nbody"][-1]["accretion method"] = accretion_method
    tmp = tempfile.mkdtemp(prefix="golden_" + name + "_")
    cfg["OutputDir"] = os.path.join(tmp, "out")
    ypath = os.path.join(tmp, "cfg.yml")
    with open(ypath, "w") as f:
        yaml.safe_dump(cfg, f, sort_keys=False)
    env = dict(os.environ, OMP_NUM_THREADS="4")
    res = subprocess.run([EXE, "start", ypath], cwd=tmp, env=env, capture_output=True, text=True)
    if res.returncode != 0:
        print(res.stdout[-3000:], res.stderr[-3000:])
        raise SystemExit(f"reference run failed for {name}")
    out = cfg["OutputDir"]
    consts, temp_unit, units = parse_constants(out)
    dims = [l for l in open(os.path.join(out, "dimensions.dat")) if not l.startswith("#")][-1].split()
    nrad, naz = int(dims[4]), int(dims[5])
    radii = np.loadtxt(os.path.join(out, "used_rad.dat"))
    assert radii.shape == (nrad + 1,)
    pdict = reftools.params_from_config(cfg, consts, nrad, naz, temp_unit, units)
    from fargocpt_b200.config import balanced_vazi_sq
    pdict["balanced_vazi_sq"] = balanced_vazi_sq(pdict, cfg, radii)  # boundary_conditions/balanced.cpp:23-52
    nsnap = int(cfg["Nsnapshots"])
    nb = len(cfg["nbody"])
    arrays = {"radii": radii}
    misc = []
    bodies = []
    for k in range(nsnap + 1):
        sd = os.path.join(out, "snapshots", str(k))
        for fname, rings in (("Sigma", nrad), ("vrad", nrad + 1), ("vazi", nrad), ("energy", nrad),
                             ("Qplus", nrad), ("Qminus", nrad), ("T_Reynolds", nrad), ("gammaeff", nrad), ("mu", nrad), ("gamma1", nrad), ("scale_height", nrad), ("Temperature", nrad), ("viscosity", nrad)):
            p = os.path.join(sd, fname + ".dat")
            if keep_snaps is not None and k not in keep_snaps:
                continue
            if os.path.exists(p):
                arrays[f"{fname}_{k}"] = np.fromfile(p, dtype=np.float64).reshape(rings, naz)
        misc.append(read_misc(os.path.join(sd, "misc.bin")))
        bodies.append([read_body(os.path.join(sd, f"nbody{b}.bin")) for b in range(nb)])
    if keep_snaps is not None:
        cfg["_keep"] = list(keep_snaps)
    meta = dict(name=name, config=cfg, params=pdict, consts=consts, temperature_unit_K=temp_unit, nsnap=nsnap,
                misc=misc, bodies=bodies, first_dt=reftools._num(cfg["FirstDT"]),
                monitor_timestep=float(cfg["MonitorTimestep"]))
    arrays["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    with open(os.path.join(HERE, name + ".yml"), "w") as f:
        cfg2 = dict(cfg)
        cfg2["OutputDir"] = "out/" + name
        yaml.safe_dump(cfg2, f, sort_keys=False)
    print(f"{name}: {nrad}x{naz}, {nsnap} snapshots, dt[1]={misc[1]['time'] - misc[0]['time']:.6e}, "
          f"N_iter={[m['n_iter'] for m in misc]}")
    if keep:
        print("  kept", tmp)
    else:
        shutil.rmtree(tmp)


if __name__ == "__main__":
    if not os.path.exists(EXE):
        raise SystemExit("build the reference first: make -f oracle/Makefile.ref -j8")
    names = [a for a in sys.argv[1:] if not a.startswith("-")] or list(CASES)
    for n in names:
        run_case(n, CASES[n], keep="--keep" in sys.argv)
