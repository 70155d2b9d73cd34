#!/usr/bin/env python
"""Run a FargoCPT setup file through the unmodified reference (oracle/_ref/fargocpt_exe_ieee) AND through
`fargocpt_b200 start` (bound to the oracle: host/fargocpt_b200_oracle_test, or --gpu for the real binary) and compare the
output directories file by file (Tools/compare_binary_output.py statistics).  Build container only (needs oracle/_ref).

    python tests/checkers/compare_start_with_reference.py /root/reference/test/cold_disk_planet/setup.yml [--snapshots 3] [--dt 1e-3] [key=value ...]

The setup is run with MonitorTimestep = --dt, Nmonitor 1, so every snapshot is one short monitor step."""
import os
import shutil
import struct
import subprocess
import sys
import tempfile

import numpy as np
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref", "fargocpt_exe_ieee")


def main(argv=None):
    args = list(sys.argv[1:] if argv is None else argv)
    gpu = "--gpu" in args
    nsnap, dt, over, restart_from, ref_threads = 3, 1e-3, {}, None, 1
    setup = None
    i = 0
    while i < len(args):
        a = args[i]
        if a == "--snapshots":
            nsnap = int(args[i + 1]); i += 1
        elif a == "--dt":
            dt = float(args[i + 1]); i += 1
        elif a == "--restart-from":  # ours does not start from the YAML but restarts from the REFERENCE's snapshot K
            restart_from = int(args[i + 1]); i += 1
        elif a == "--ref-threads":  # OpenMP threads of the reference run (its fields do not depend on the thread count)
            ref_threads = int(args[i + 1]); i += 1
        elif a in ("--gpu", "--keep", "--vs-reference-restart", "--ulp-sensitivity"):
            pass
        elif "=" in a:
            k, v = a.split("=", 1)
            over[k] = yaml.safe_load(v)
        else:
            setup = a
        i += 1
    cfg = {k: v for k, v in yaml.safe_load(open(setup)).items() if not k.startswith("_")}  # "_keep" etc.: fixture bookkeeping
    cfg.update({"MonitorTimestep": dt, "Nmonitor": 1, "Nsnapshots": nsnap, "WriteAtEveryTimestep": "yes"})
    cfg.update(over)
    tmp = tempfile.mkdtemp(prefix="cmpstart_")
    cfg["OutputDir"] = os.path.join(tmp, "ref")
    ypath = os.path.join(tmp, "setup.yml")
    yaml.safe_dump(cfg, open(ypath, "w"), sort_keys=False)
    env = dict(os.environ, OMP_NUM_THREADS=str(ref_threads))
    try:
        r = subprocess.run([REF, "start", ypath], cwd=tmp, env=env, capture_output=True, text=True, timeout=float(os.environ.get("CMPSTART_TIMEOUT", 3600)))
    except subprocess.TimeoutExpired:
        raise SystemExit("reference run timed out")
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-2000:])
        raise SystemExit("reference run failed")
    exe = os.path.join(ROOT, "host", "fargocpt_b200" if gpu else "fargocpt_b200_oracle_test")
    ours = os.path.join(tmp, "ours")
    if restart_from is None:
        r = subprocess.run([exe, "start", ypath, "--out", ours, "--until", str(nsnap)], capture_output=True, text=True)
    else:
        r = subprocess.run([exe, "restart", str(restart_from), cfg["OutputDir"], "--out", ours, "--until", str(nsnap)], capture_output=True, text=True)
    print(r.stdout[-300:], r.stderr[-600:])
    if r.returncode != 0:
        raise SystemExit("fargocpt_b200 start failed")
    ref = cfg["OutputDir"]
    if restart_from is not None and "--vs-reference-restart" in args:
        # compare with what the REFERENCE does when it restarts from the same snapshot (not with its uninterrupted run: the two
        # differ where the reference's restart is not seamless, e.g. in a corotating frame, frame_of_reference.cpp:19-28)
        ref2 = os.path.join(tmp, "ref_restarted")
        shutil.copytree(ref, ref2)
        for k in range(restart_from + 1, nsnap + 1):
            shutil.rmtree(os.path.join(ref2, "snapshots", str(k)))
        cfg2 = dict(cfg, OutputDir=ref2)
        ypath2 = os.path.join(tmp, "setup_restart.yml")
        yaml.safe_dump(cfg2, open(ypath2, "w"), sort_keys=False)
        r = subprocess.run([REF, "restart", str(restart_from), ypath2], cwd=tmp, env=env, capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stdout[-2000:], r.stderr[-2000:])
            raise SystemExit("reference restart failed")
        ref = ref2
    worst = 0.0
    for f in ("constants.yml", "units.yml", "used_rad.dat"):
        if restart_from is not None and not os.path.exists(os.path.join(ours, f)):
            continue
        same = open(os.path.join(ref, f)).read() == open(os.path.join(ours, f)).read()
        print(f"{f}: {'identical' if same else 'DIFFERENT'}")
    for k in range(0 if restart_from is None else restart_from + 1, nsnap + 1):
        sd_r, sd_o = os.path.join(ref, "snapshots", str(k)), os.path.join(ours, "snapshots", str(k))
        line = [f"snapshot {k}:"]
        for f in ("Sigma", "vrad", "vazi", "energy"):
            pr, po = os.path.join(sd_r, f + ".dat"), os.path.join(sd_o, f + ".dat")
            if not (os.path.exists(pr) and os.path.exists(po)):
                continue
            a, b = np.fromfile(pr), np.fromfile(po)
            if a.shape != b.shape:
                line.append(f"{f} SHAPE {a.shape} vs {b.shape}")
                continue
            d = np.abs(a - b)
            scale = np.abs(a).max() or 1.0
            worst = max(worst, float(d.max() / scale))
            line.append(f"{f} ndiff={int((d != 0).sum())} max|d|/scale={d.max() / scale:.2e}")
        mr = struct.unpack("<IIddddQ", open(os.path.join(sd_r, "misc.bin"), "rb").read()[:48])
        mo = struct.unpack("<IIddddQ", open(os.path.join(sd_o, "misc.bin"), "rb").read()[:48])
        line.append(f"misc {'identical' if mr == mo else f'ref={mr} ours={mo}'}")
        nb = 0
        while os.path.exists(os.path.join(sd_r, f"nbody{nb}.bin")):
            br = open(os.path.join(sd_r, f"nbody{nb}.bin"), "rb").read()
            bo = open(os.path.join(sd_o, f"nbody{nb}.bin"), "rb").read()
            sr, so = np.array(struct.unpack("<5d", br[8:48])), np.array(struct.unpack("<5d", bo[8:48]))
            line.append(f"body{nb} {'identical' if np.array_equal(sr, so) else f'max|d|={np.abs(sr - so).max():.1e}'}")
            nb += 1
        print(" ".join(line))
    # monitor files: same header, same columns (nan in ours = a column this path does not evaluate)
    def table(path):
        head = [l for l in open(path) if l.startswith("#")]
        rows = np.array([[float(x) for x in l.split()] for l in open(path) if not l.startswith("#") and l.strip()])
        return head, rows
    mon = [] if restart_from is not None else [f for f in sorted(os.listdir(os.path.join(ours, "monitor"))) if f.startswith(("Quantities", "nbody"))]
    for f in mon:
        pr, po = os.path.join(ref, "monitor", f), os.path.join(ours, "monitor", f)
        if not os.path.exists(pr):
            continue
        (hr, a), (ho, b) = table(pr), table(po)
        n = min(len(a), len(b))
        if n == 0 or a.shape[1] != b.shape[1]:
            print(f"monitor/{f}: SHAPE {a.shape} vs {b.shape}")
            continue
        with np.errstate(all="ignore"):
            scale = np.maximum(np.abs(a[:n]).max(axis=0), 1e-300)
            d = np.abs(a[:n] - b[:n]).max(axis=0) / scale
        cols = ", ".join(f"{k}:{v:.0e}" for k, v in enumerate(d) if np.isfinite(v) and v > 1e-9)
        print(f"monitor/{f}: header {'identical' if hr == ho else 'DIFFERENT'}, rows {len(a)} vs {len(b)}, columns not evaluated: "
              f"{[k for k, v in enumerate(d) if not np.isfinite(v)]}, columns off by more than 1e-9 of their scale: {cols or 'none'}")
    if "--ulp-sensitivity" in args and len(cfg.get("nbody", [])) > 1:
        # How far do two runs of the REFERENCE ITSELF drift apart when the planet starts one ulp further out?  The bodies are advanced
        # on the host (ours: compensated RK4, the reference: REBOUND IAS15) and agree to ~1e-16; this is the yardstick for what such
        # a difference does to the gas after the same number of steps.
        import copy
        cfg2 = copy.deepcopy(cfg)
        nb = cfg2["nbody"][1]
        a0 = float(str(nb.get("semi-major axis", 1.0)).split()[0])
        unit = " ".join(str(nb.get("semi-major axis", 1.0)).split()[1:])
        nb["semi-major axis"] = (repr(float(np.nextafter(a0, 2 * a0))) + " " + unit).strip()
        cfg2["OutputDir"] = os.path.join(tmp, "ref_ulp")
        y2 = os.path.join(tmp, "setup_ulp.yml")
        yaml.safe_dump(cfg2, open(y2, "w"), sort_keys=False)
        r = subprocess.run([REF, "start", y2], cwd=tmp, env=env, capture_output=True, text=True)
        if r.returncode == 0:
            line = [f"reference vs reference with the planet one ulp further out, snapshot {nsnap}:"]
            for f in ("Sigma", "vrad", "vazi", "energy"):
                pr, po = os.path.join(ref, "snapshots", str(nsnap), f + ".dat"), os.path.join(cfg2["OutputDir"], "snapshots", str(nsnap), f + ".dat")
                if os.path.exists(pr) and os.path.exists(po):
                    a, b = np.fromfile(pr), np.fromfile(po)
                    line.append(f"{f} ulpdev={np.abs(a - b).max() / (np.abs(a).max() or 1.0):.2e}")
            print(" ".join(line))
    print(f"worst field deviation relative to the field scale: {worst:.2e}")
    if "--keep" not in args:
        shutil.rmtree(tmp, ignore_errors=True)
    else:
        print("kept", tmp)
    return worst
    if "--keep" not in args:
        shutil.rmtree(tmp)
    else:
        print("kept", tmp)


if __name__ == "__main__":
    main()
