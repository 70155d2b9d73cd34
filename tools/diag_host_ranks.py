#!/usr/bin/env python
"""Diagnostic: `fargocpt_b200 start <setup> --ranks N` against one rank, every snapshot file: max deviation / field scale and
where it sits.  usage: diag_host_ranks.py <setup.yml in tests/golden> <N> <until> [Key=value ...]"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_multi import _setup_with  # noqa: E402

setup, n, until = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
over = dict(a.split("=", 1) for a in sys.argv[4:])
tmp = tempfile.mkdtemp(prefix="diag_ranks_")
yml = os.path.join(tmp, "setup.yml")
_setup_with(os.path.join(ROOT, "tests", "golden", setup + ".yml"), yml, **over)
exe = os.path.join(ROOT, "host", "fargocpt_b200")
outs = {}
for ranks in (1, n):
    out = os.path.join(tmp, f"out{ranks}")
    cmd = [exe, "start", yml, "--out", out, "--until", str(until)] + (["--ranks", str(ranks)] if ranks > 1 else [])
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(ranks, "rank(s): rc", res.returncode, res.stdout.strip().splitlines()[-1:] , res.stderr[-300:])
    outs[ranks] = out
dims = [l for l in open(os.path.join(outs[1], "dimensions.dat")) if not l.startswith("#")][-1].split()
nr, naz = int(dims[4]), int(dims[5])
for snap in range(until + 1):
    d1, dn = (os.path.join(outs[r], "snapshots", str(snap)) for r in (1, n))
    for f in sorted(os.listdir(d1)):
        if not f.endswith(".dat"):
            continue
        x, y = np.fromfile(os.path.join(d1, f)), np.fromfile(os.path.join(dn, f))
        if x.shape != y.shape:
            print(snap, f, "SHAPE", x.shape, y.shape)
            continue
        x, y = np.nan_to_num(x), np.nan_to_num(y)
        d = np.abs(x - y)
        k = int(d.argmax())
        print(f"snapshot {snap} {f:12s} differ {int((x != y).sum()):8d}  max dev / scale {d.max() / max(np.abs(x).max(), 1e-300):.3e}  at ring {k // naz} col {k % naz}"
              f"  rings with differences: {sorted(set((np.nonzero(x != y)[0] // naz).tolist()))[:12]}")
