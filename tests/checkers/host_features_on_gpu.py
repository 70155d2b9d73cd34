#!/usr/bin/env python
"""Every host-driver feature that test_host_driver.py checks on the oracle-bound driver against the unmodified reference, run once
more on the PRODUCT (host/fargocpt_b200 on libfargo_b200.so, a B200) against the reference binary on the same box
(oracle/_ref/fargocpt_exe_ieee travels with the snapshot): the setups of test_host_driver.REFERENCE_SETUPS that live in this repo
(frames, circumbinary disks, initial-condition variants, Leapfrog + accretion, corotation), the radiative fixtures' setups and the
monitor files.  One line per case and a markdown table (gpurun_out/<tag>_host_features.md).

    python tests/checkers/host_features_on_gpu.py <tag> [--cpu]"""
import contextlib
import importlib.util
import io
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "hostfeat"
    cpu = "--cpu" in sys.argv
    import test_host_driver as T
    spec = importlib.util.spec_from_file_location("cmpstart", os.path.join(ROOT, "tests", "checkers", "compare_start_with_reference.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cases = [(s, o) for s, o in T.REFERENCE_SETUPS if os.path.exists(s)]
    G = os.path.join(ROOT, "tests", "golden")
    cases += [(os.path.join(G, n + ".yml"), ["--dt", "1e-3"]) for n in ("adia_irrad", "adia_irrad_lf", "adia_cool_lin", "adia_cool_bell")]
    cases += [(os.path.join(G, "multi_body_setup.yml"), ["--dt", "2e-3"]), (os.path.join(G, "shock_tube_setup.yml"), ["--dt", "0.02"])]
    rows, ok = [], True
    for setup, over in cases:
        buf = io.StringIO()
        try:
            with contextlib.redirect_stdout(buf):
                worst = mod.main([setup, "--snapshots", "2"] + (over if "--dt" in over else ["--dt", "1e-3"] + over) + ([] if cpu else ["--gpu"]))
        except SystemExit as e:  # a run failed
            rows.append((os.path.basename(setup), " ".join(over), "FAILED: " + str(e), "", ""))
            ok = False
            print(rows[-1])
            continue
        text = buf.getvalue()
        files = all(f + ": identical" in text for f in ("constants.yml", "units.yml", "used_rad.dat"))
        snap0 = [l for l in text.splitlines() if l.startswith("snapshot 0:")][0]
        s0 = "DIFF" not in snap0 and not re.search(r"ndiff=[1-9]", snap0)
        # per monitor file: header identical?  columns that deviate by more than 1e-9 of their scale (column:deviation)
        mon = "; ".join(l.split(":", 1)[0].replace("monitor/", "") + (" header ok" if "header identical" in l else " HEADER DIFFERS") +
                        ", off: " + l.rsplit("scale:", 1)[-1].strip() for l in text.splitlines() if l.startswith("monitor/"))
        good = files and s0 and worst <= 1e-10
        ok &= good
        rows.append((os.path.basename(setup), " ".join(a for a in over if a not in ("--dt",) and not re.fullmatch(r"[0-9.e-]+", a)),
                     "identical" if (files and s0) else "DIFFERENT", f"{worst:.1e}", mon))
        print(("PASS " if good else "FAIL ") + " | ".join(rows[-1]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", tag + "_host_features.md"), "w") as f:
        f.write("| setup (tests/golden) | overrides | constants / units / radii / snapshot 0 | worst field deviation / scale after 2 snapshots | monitor files |\n|---|---|---|---|---|\n")
        for r in rows:
            f.write("| " + " | ".join(r) + " |\n")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
