#!/usr/bin/env python
"""BASELINE.json configs[1..3] on the PRODUCT (host/fargocpt_b200 on libfargo_b200.so, a B200) against the
unmodified reference (oracle/_ref/fargocpt_exe_ieee, built by oracle/Makefile.ref; it travels to the GPU box with the snapshot)
run on the same box's host cores, from the same setup file, over >= 100 CFL-limited hydro steps of natural time stepping
(MonitorTimestep = T, Nmonitor 1, Nsnapshots 1: both codes step with their own CFL dt until t = T).  Compared with the logic of the
reference's Tools/compare_binary_output.py:16-44: every double of Sigma, vrad, vazi, energy of the final snapshot, and misc.bin
(N_iter, time, last dt).  Writes gpurun_out/<tag>_baseline_configs.json and a markdown table for DESIGN.md.

    python tests/checkers/baseline_configs_on_gpu.py <tag> [--cpu]      (--cpu: the oracle-bound driver instead, reduced sizes)"""
import contextlib
import importlib.util
import io
import json
import os
import re
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
G = os.path.join(ROOT, "tests", "golden")
# (name, setup, overrides, T): T chosen so that the run takes a bit more than 100 hydro steps at that resolution
STAR_ONLY = "nbody=[{name: Star, semi-major axis: 0.0, mass: 1 solMass, eccentricity: 0, radius: 1 solRadius, temperature: 0}]"
# FirstDT: the reference's first step is CFLmaxVar^3 * FirstDT whatever the CFL says when beta cooling relaxes towards the reference
# state (its Q- of the first CFL is NaN: SourceEuler.cpp:284 runs before the reference state exists) — reproduced bit for bit, but a
# first step of 0.13 wrecks a disk with a Jupiter in it, so the larger cases start from 1e-3 like a careful user would
CASES = [
    ("C2 test/cold_disk_planet verbatim (cps 3: 97x376)", "cold_disk_planet_setup.yml", [], 5.0),
    ("C2 cold_disk_planet 512x1024", "cold_disk_planet_setup.yml", ["cps=-1", "Nrad=512", "Naz=1024", "FirstDT=1e-3"], 1.0),
    ("C3 adiabatic + viscous heating + beta cooling 2048x4096", "baseline_config2_setup.yml", ["FirstDT=1e-3"], 0.3),
    # without the planet nothing on the host integrates an orbit (ours: RK4, the reference: REBOUND IAS15): every double must agree
    ("C3 2048x4096, star only (bit-exact expected)", "baseline_config2_setup.yml", [STAR_ONLY, "FirstDT=1e-3"], 0.3),
    ("C4 examples/config.yml physics 1024x2048", "baseline_config3_setup.yml", ["Nrad=1024", "Naz=2048", "FirstDT=1e-3"], 0.22),
    ("C5 physics (adiabatic, Jupiter) 1024x2048", "baseline_config4_setup.yml", ["Nrad=1024", "Naz=2048", "FirstDT=1e-3"], 0.5),
]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "baseline"
    cpu = "--cpu" in sys.argv
    spec = importlib.util.spec_from_file_location("cmpstart", os.path.join(ROOT, "tests", "checkers", "compare_start_with_reference.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rows = []
    for name, setup, over, T in CASES:
        if cpu:  # reduced sizes so the oracle-bound driver finishes: /8 per direction, T x 8
            over = [o for o in over if not o.startswith(("Nrad", "Naz"))]
            if "verbatim" not in name:
                over += ["Nrad=128", "Naz=256"]
                T = T * (2048 if "2048x4096" in name else 512 if "512x" in name else 1024) / 128
        args = [os.path.join(G, setup), "--snapshots", "1", "--dt", repr(T), "--ref-threads", str(os.cpu_count() or 1), "--ulp-sensitivity"] + over
        if not cpu:
            args.append("--gpu")
        buf = io.StringIO()
        t0 = time.time()
        try:
            with contextlib.redirect_stdout(buf):
                worst = mod.main(args)
        except SystemExit as e:
            rows.append({"case": name, "error": str(e), "log": buf.getvalue()[-1500:]})
            print(name, "FAILED", e, buf.getvalue()[-1500:], flush=True)
            continue
        text = buf.getvalue()
        snap = [l for l in text.splitlines() if l.startswith("snapshot 1:")][0]
        snap0 = [l for l in text.splitlines() if l.startswith("snapshot 0:")][0]
        m = re.search(r"Total Hydrosteps (\d+)", text)
        fields = {f: (int(n), float(d)) for f, n, d in re.findall(r"(\w+) ndiff=(\d+) max\|d\|/scale=([0-9.eE+-]+|nan)", snap)}
        ulp = {f: float(d) for f, d in re.findall(r"(\w+) ulpdev=([0-9.eE+-]+|nan)", text)}
        rows.append({"case": name, "overrides": over, "T": T, "hydro_steps": int(m.group(1)) if m else None,
                     "worst_rel_to_field_scale": worst, "fields_ndiff_maxrel": fields, "reference_one_ulp_sensitivity": ulp,
                     "misc_identical": "misc identical" in snap, "snapshot0_identical": snap0.count("ndiff=0 ") >= 3,
                     "seconds_both_runs": round(time.time() - t0, 1), "line": snap})
        print(json.dumps(rows[-1]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = os.path.join(ROOT, "gpurun_out", f"{tag}_baseline_configs")
    json.dump({"backend": "oracle-bound driver (CPU)" if cpu else "host/fargocpt_b200 on a B200", "rows": rows}, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write("| config | hydro steps | snapshot 0 identical | misc.bin (N_iter, t, last dt) identical | Sigma | vrad | vazi | energy | worst | "
                "the reference against itself with the planet 1 ulp further out (Sigma / vrad / vazi / energy) |\n|---|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            if "error" in r:
                f.write(f"| {r['case']} | FAILED: {r['error']} |\n")
                continue
            fl = r["fields_ndiff_maxrel"]
            cell = lambda k: f"{fl[k][1]:.1e} ({fl[k][0]} differ)" if k in fl else "-"  # noqa: E731
            f.write(f"| {r['case']} | {r['hydro_steps']} | {r['snapshot0_identical']} | {r['misc_identical']} | {cell('Sigma')} | {cell('vrad')} | "
                    f"{cell('vazi')} | {cell('energy')} | {r['worst_rel_to_field_scale']:.1e} | "
                    + (" / ".join(f"{r['reference_one_ulp_sensitivity'].get(k, float('nan')):.1e}" for k in ("Sigma", "vrad", "vazi", "energy"))
                       if r["reference_one_ulp_sensitivity"] else "-") + " |\n")
    print(open(out + ".md").read())
    def ok(r):
        if "error" in r:
            return False
        for f, (n, d) in r["fields_ndiff_maxrel"].items():
            if not d <= max(1e-10, 3.0 * r["reference_one_ulp_sensitivity"].get(f, 0.0)):
                return False
        return True
    bad = [r for r in rows if not ok(r)]
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
