#!/usr/bin/env python
"""Per-stage GPU-vs-oracle difference report (diagnostics, not a test).  Run on the GPU box."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import goldenrun  # noqa: E402
import reftools  # noqa: E402
from fargocpt_b200 import HydroContext, abi  # noqa: E402
from test_gpu_parity import STAGES, CASES  # noqa: E402

for name in (sys.argv[1:] or CASES):
    meta, z = reftools.load_golden(name)
    params = reftools.make_params(meta["params"])
    gpu, cpu = HydroContext(params, z["radii"]), reftools.OracleContext(params, z["radii"])
    la = goldenrun.start_from_snapshot0(gpu, meta, z)[0]
    lb = goldenrun.start_from_snapshot0(cpu, meta, z)[0]
    print(f"== {name}: first dt gpu={la.last_dt!r} cpu={lb.last_dt!r} equal={la.last_dt == lb.last_dt}")
    dt = meta["monitor_timestep"]
    for step in range(2):
        for stage, args in STAGES:
            a = tuple(dt if x == "dt" else x for x in args)
            gpu.stage(stage, *a)
            cpu.stage(stage, *a)
            if stage in ("potential", "derived"):
                continue
            fields = [(abi.SIGMA, "Sigma"), (abi.ENERGY, "energy")]
            if stage == "transport" or (stage == "boundary" and args[-1] == 1):
                fields += [(abi.VRAD, "vrad"), (abi.VAZI, "vazi")]
            msg = []
            for fid, fn in fields:
                A, B = gpu.download_slab(fid), cpu.download_slab(fid)
                st = reftools.compare_stats(A, B)
                rows = np.unique(np.nonzero(A != B)[0])
                msg.append(f"{fn}: ndiff={st['n_diff']} maxrel={st['max_rel']:.2e} rings={rows[:6].tolist()}{'...' if len(rows) > 6 else ''}")
            print(f"  step {step} {stage:10s} " + " | ".join(msg))
        if stage == "derived":
            pass
        print(f"  step {step} nshift equal={np.array_equal(gpu.nshift(), cpu.nshift())} cfl gpu={gpu.condition_cfl()!r} cpu={cpu.condition_cfl()!r}")
