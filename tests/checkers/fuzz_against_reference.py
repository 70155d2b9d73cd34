#!/usr/bin/env python
"""Randomised side-by-side sweep: draw setups from the option space this path accepts, run each through the unmodified reference
(oracle/_ref/fargocpt_exe_ieee) and through `fargocpt_b200 start` (oracle-bound test binary, or --gpu for the product), and report
the worst field deviation per draw (compare_start_with_reference.py does the comparison).  Build container only.

    python tests/checkers/fuzz_against_reference.py [--seeds 0:40] [--wide] [--gpu] [--snapshots 3] [--dt 0.05] [-- <arguments of compare_start_with_reference.py>]

A draw that the host refuses by name is reported as "refused" (that is the contract for physics outside the path); a draw the
reference itself rejects is skipped."""
import contextlib
import io
import os
import random
import sys

import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import compare_start_with_reference as cmp  # noqa: E402

BASE = os.path.join(ROOT, "tests", "golden", "cold_disk_planet_setup.yml")


def draw(seed, wide=False):
    rng = random.Random(seed)
    pick = rng.choice
    cfg = {k: v for k, v in yaml.safe_load(open(BASE)).items() if not k.startswith("_")}
    cfg.pop("cps", None)
    cfg["Nrad"], cfg["Naz"] = pick([24, 33, 48]), pick([32, 48, 96])
    cfg["RadialSpacing"] = pick(["Logarithmic", "Arithmetic", "Exponential"])
    if cfg["RadialSpacing"] == "Exponential":  # coarser grids send the reference's growth-factor iteration into its trivial root (init.cpp:112-128)
        cfg["Nrad"] = 48
    cfg["Rmin"], cfg["Rmax"] = pick([0.3, 0.4, 0.5]), pick([2.0, 2.5, 3.1])
    eos = pick(["Isothermal", "Ideal", "Ideal", "PVTE"])
    cfg["EquationOfState"] = eos
    cfg["AdiabaticIndex"] = pick([1.4, 5.0 / 3.0])
    cfg["Integrator"] = pick(["Euler", "Leapfrog"])
    cfg["FluxLimiter"] = pick(["VanLeer", "mc"])
    cfg["Transport"] = pick(["FARGO", "FARGO", "Standard"])
    cfg["ArtificialViscosity"] = pick(["SN", "TW", "TW_LOCAL", "None"])
    cfg["ArtificialViscosityFactor"] = pick([1.41, 2.0, 3.0])
    cfg["ArtificialViscosityDissipation"] = pick(["Yes", "No"])
    visc = pick(["alpha", "const", "none"])
    cfg["ViscousAlpha"] = pick([1e-3, 1e-2]) if visc == "alpha" else 0.0
    if visc == "const":
        cfg["ConstantViscosity"] = pick([1e-5, 4e-6])
    if visc == "alpha" and eos != "Isothermal" and rng.random() < 0.3:
        cfg["AlphaMode"], cfg["AlphaCold"], cfg["AlphaHot"] = 1, 0.01, 0.1
    cfg["StabilizeViscosity"] = pick([0, 0, 1, 2])
    cfg["HeatingViscous"] = pick(["yes", "no"])
    cfg["SigmaSlope"], cfg["FlaringIndex"] = pick([0.5, 1.0, 1.5]), pick([0.0, 0.25])
    cfg["AspectRatio"] = pick([0.03, 0.05, 0.1])
    cfg["CFL"] = pick([0.3, 0.5])
    cfg["ThicknessSmoothing"] = pick([0.4, 0.6])
    cfg["IndirectTermMode"] = pick([0, 1])
    cfg["DiskFeedback"] = pick(["yes", "no"])
    rng.random()  # (a former draw; kept so that the seeds keep their meaning)
    frame = pick(["F", "F", "R", "C"])
    cfg["Frame"] = "C" if frame == "C" else "F"  # Interpret.cpp:313-321: Fixed (rotating with OmegaFrame) | Corotating
    cfg["OmegaFrame"] = 1.0 if frame == "R" else 0
    if eos != "Isothermal":
        cool = pick(["none", "beta", "thermal", "scurve"])
        if cool == "beta":
            cfg["CoolingBetaLocal"], cfg["CoolingBeta"] = "Yes", pick([1.0, 10.0])
            cfg["CoolingBetaReference"] = pick(["zero", "reference", "diskmodel", "floor"])
        elif cool in ("thermal", "scurve"):
            cfg["SurfaceCooling"] = cool
            if cool == "scurve":
                cfg["ScurveType"] = pick(["Kimura", "Ichikawa"])
            cfg["Opacity"] = pick(["Lin", "Bell", "Constant"])
            cfg["KappaConst"] = 1.0
    # boundaries
    individual = rng.random() < 0.2  # both sides or none: an inner composite cannot be combined with individual outer keys (config.cpp:147)
    for side in ("Inner", "Outer"):
        comp = "individual" if individual else pick(["Reflecting", "Outflow", "Zerogradient", "Reference"])
        if comp == "individual":
            cfg[side + "Boundary"] = "individual"  # the key itself must exist (Interpret.cpp:286-293)
            cfg[side + "BoundarySigma"] = pick(["zerogradient", "reference"])
            cfg[side + "BoundaryEnergy"] = pick(["zerogradient", "reference"])
            vr = ["zerogradient", "reflecting", "outflow", "reference"] + (["viscous", "keplerian"] if side == "Inner" else [])
            cfg[side + "BoundaryVrad"] = pick(vr)
        else:
            cfg[side + "Boundary"] = comp
        cfg[side + "BoundaryVazi"] = pick(["keplerian", "zerogradient", "zeroshear", "balanced", "reference"])
    damp = pick([True, True, False])
    cfg["Damping"] = "Yes" if damp else "No"
    for q in ("Energy", "VRadial", "VAzimuthal", "SurfaceDensity"):
        for side in ("Inner", "Outer"):
            cfg["Damping" + q + side] = pick(["Initial", "Mean", "Zero", "None"]) if q == "VRadial" else pick(["Initial", "Mean", "None"])
    # bodies
    planet = cfg["nbody"][1]
    planet["mass"] = pick([2e-5, 1e-3])
    planet["eccentricity"] = pick([0.0, 0.1])
    planet["ramp-up time"] = pick([0, 10])
    planet["accretion efficiency"] = pick([0.0, 0.0, 1.0])
    if planet["accretion efficiency"] > 0:
        planet["accretion method"] = pick(["kley", "sinkhole", "viscous"])
    if rng.random() < 0.25:
        cfg["nbody"] = cfg["nbody"][:1]
        if cfg["Frame"] == "C":
            cfg["Frame"] = "F"
    cfg["WriteDiskQuantities"] = pick(["Yes", "No"])
    cfg["WriteMassFlow"] = pick(["yes", "no"])
    if wide:  # a second round of switches: disk model, frame centre, several bodies, irradiation, profile cut-offs, limits
        cfg["CFLmaxVar"] = pick([1.1, 1.5])
        cfg["SigmaFloor"] = pick([1e-9, 1e-7])
        cfg["RadialViscosityFactor"] = pick([1.0, 1.0, 0.5])
        cfg["HeatingViscousFactor"] = pick([1.0, 0.5])
        cfg["ImposedDiskDrift"] = pick([0.0, 0.0, 1e-3])
        cfg["InitializePureKeplerian"] = pick(["no", "no", "yes"])
        cfg["InitializeVradialZero"] = pick(["no", "yes"])
        cfg["HeatingCoolingCFLlimit"] = pick([1.0, 10.0, 1000.0])
        cfg["MinimumTemperature"] = pick(["3 K", "10 K"])
        cfg["MaximumTemperature"] = pick(["1e100 K", "5000 K"])
        cfg["InnerBoundaryVaziKeplerianFactor"] = pick([1.0, 0.98])
        cfg["DampingTimeFactor"] = pick([0.05, 1.0])
        cfg["DampingInnerLimit"], cfg["DampingOuterLimit"] = pick([1.1, 1.311]), pick([0.763, 0.9])
        if rng.random() < 0.3:
            cfg["ProfileCutoffOuter"], cfg["ProfileCutoffPointOuter"], cfg["ProfileCutoffWidthOuter"] = "yes", 0.8 * cfg["Rmax"], 0.1
        if rng.random() < 0.3:
            cfg["ProfileCutoffInner"], cfg["ProfileCutoffPointInner"], cfg["ProfileCutoffWidthInner"] = "yes", 1.5 * cfg["Rmin"], 0.05
        if cfg.get("CoolingBetaLocal") == "Yes":
            cfg["CoolingBetaRampUp"] = pick([0.0, 0.3])
        if cfg.get("SurfaceCooling") == "thermal":
            cfg["KappaFactor"], cfg["TauFactor"], cfg["TauMin"] = pick([1.0, 2.0]), pick([0.5, 1.0]), pick([0.01, 0.1])
            cfg["CoolingRadiativeFactor"] = pick([1.0, 0.5])
            if rng.random() < 0.5:
                cfg["nbody"][0]["temperature"] = "5000 K"
                cfg["nbody"][0]["irradiation ramp-up time"] = pick([0.0, 0.2])
        if eos == "PVTE":
            cfg["HydrogenMassFraction"] = pick([0.75, 0.7])
        if len(cfg["nbody"]) > 1:
            planet = cfg["nbody"][1]
            planet["semi-major axis"] = pick([1, 1.3])
            planet["cubic smoothing factor"] = pick([0.0, 0.5])
            planet["argument of pericenter"] = pick([0.0, 1.0])
            planet["trueanomaly"] = pick([0.0, 2.0])
            if rng.random() < 0.4:
                second = dict(planet, name="second", mass=pick([1e-4, 5e-4]))
                second["semi-major axis"] = 1.7
                second["accretion efficiency"] = 0.0
                second.pop("accretion method", None)
                cfg["nbody"].append(second)
            cfg["HydroFrameCenter"] = pick(["primary", "primary", "binary", "all"]) if planet["mass"] >= 1e-3 else "primary"
            if cfg["HydroFrameCenter"] != "primary":
                cfg["Frame"] = "F"
    return cfg


def main():
    args = sys.argv[1:]
    lo, hi = 0, 40
    if "--seeds" in args:
        lo, hi = (int(x) for x in args[args.index("--seeds") + 1].split(":"))
    extra = [a for a in ("--gpu",) if a in args]
    if "--" in args:  # everything after "--" goes to compare_start_with_reference.py (e.g. -- --restart-from 1 --vs-reference-restart)
        extra += args[args.index("--") + 1:]
    dt = args[args.index("--dt") + 1] if "--dt" in args else "0.05"
    nsnap = args[args.index("--snapshots") + 1] if "--snapshots" in args else "3"
    import tempfile
    os.environ.setdefault("CMPSTART_TIMEOUT", "120")  # a draw whose time step collapses in the reference is skipped
    tmp = tempfile.mkdtemp(prefix="fuzz_")
    worst_all = 0.0
    for seed in range(lo, hi):
        cfg = draw(seed, wide="--wide" in args)
        path = os.path.join(tmp, f"draw_{seed}.yml")
        yaml.safe_dump(cfg, open(path, "w"), sort_keys=False)
        out = io.StringIO()
        status = "ok"
        worst = float("nan")
        try:
            with contextlib.redirect_stdout(out):
                worst = cmp.main([path, "--snapshots", nsnap, "--dt", dt] + extra)
        except SystemExit as e:
            status = str(e)
        text = out.getvalue()
        tag = " ".join(f"{k}={cfg.get(k)}" for k in ("EquationOfState", "Integrator", "Frame", "Transport", "ArtificialViscosity",
                                                     "SurfaceCooling", "AlphaMode", "InnerBoundary", "OuterBoundary") if cfg.get(k) is not None)
        if status != "ok":
            reason = [l for l in text.splitlines() if l.strip()][-3:]
            print(f"seed {seed}: {status}: {' | '.join(reason)[-400:]}  [{tag}]", flush=True)
            continue
        notes = [l for l in text.splitlines() if "DIFFERENT" in l or "SHAPE" in l or ("columns off" in l and not l.rstrip().endswith("none"))]
        worst_all = max(worst_all, worst)
        print(f"seed {seed}: worst {worst:.2e}  [{tag}]" + ("".join("\n    " + n for n in notes)), flush=True)
    print(f"worst over all draws: {worst_all:.2e}; setups kept under {tmp}")


if __name__ == "__main__":
    main()
