#!/usr/bin/env python
"""Lock-step GPU-vs-oracle report over a fixture with an accreting planet (diagnostics, not a test).  Run on the GPU box:
    python tests/checkers/gpu_diag_accrete.py [fixture ...]
Three passes per fixture: (A) fargo_step (fused kernels) with accretion, fields compared after every accretion call and
after every step; (B) the same without the accretion calls (is it the planet or the accretion?); (C) the per-stage entry
points with accretion, compared after every stage.  FARGO_DIAG_CPU=1 binds both sides to the oracle (syntax check)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import goldenrun  # noqa: E402
import reftools  # noqa: E402
from fargocpt_b200 import abi  # noqa: E402

STAGES = [("potential", ()), ("sources", ("dt",)), ("artvisc", ("dt",)), ("viscosity", ("dt",)), ("substep3", ("dt",)),
          ("boundary", (0.0, 0)), ("transport", ("dt",)), ("boundary", ("dt", 1)), ("derived", ())]
FIELDS = [(abi.SIGMA, "Sigma"), (abi.VRAD, "vrad"), (abi.VAZI, "vazi"), (abi.ENERGY, "energy")]


def make_pair(meta, z):
    params = reftools.make_params(meta["params"])
    cpu = reftools.OracleContext(params, z["radii"])
    if os.environ.get("FARGO_DIAG_CPU"):
        return reftools.OracleContext(params, z["radii"]), cpu
    from fargocpt_b200 import HydroContext
    return HydroContext(params, z["radii"]), cpu


def diff(gpu, cpu, fields, tag):
    bad = False
    msg = []
    for fid, fn in fields:
        A, B = gpu.download_slab(fid), cpu.download_slab(fid)
        st = reftools.compare_stats(A, B)
        if st["n_diff"]:
            bad = True
            ii, jj = np.nonzero(A != B)
            w = int(np.argmax(np.abs(A - B)[ii, jj]))
            msg.append(f"{fn}: ndiff={st['n_diff']} maxrel={st['max_rel']:.2e} rings={np.unique(ii)[:8].tolist()} "
                       f"worst=({ii[w]},{jj[w]}) gpu={A[ii[w], jj[w]]!r} cpu={B[ii[w], jj[w]]!r}")
    print(f"    {tag}: " + (" | ".join(msg) if bad else "identical"))
    return bad


def run(name, accrete, staged, max_steps=8):
    meta, z = reftools.load_golden(name)
    gpu, cpu = make_pair(meta, z)
    adia = bool(meta["params"]["adiabatic"])
    fields = FIELDS if adia else FIELDS[:3]
    loops = []
    for ctx in (gpu, cpu):
        loop, omega = goldenrun.start_from_snapshot0(ctx, meta, z)
        loops.append(loop)
    print(f"== {name} accrete={accrete} staged={staged}: first dt equal={loops[0].last_dt == loops[1].last_dt}")
    accretors = [b for b, rec in enumerate(meta["bodies"][0]) if len(rec) > 7 and rec[7] > 0.0]
    nbad = 0
    for k in range(1, max_steps + 1):
        dts = [loop.next_dt() for loop in loops]
        print(f"  step {k}: cfl gpu={dts[0][1]!r} cpu={dts[1][1]!r} equal={dts[0][1] == dts[1][1]}")
        step_dt = dts[1][0]
        if accrete:
            res = []
            for ctx in (gpu, cpu):
                for b in accretors:
                    res.append(ctx.accrete_kley(*goldenrun.accretion_inputs(meta, k - 1, b, step_dt)))
            print(f"    accreted gpu={res[0]} cpu={res[-1]}")
            nbad += diff(gpu, cpu, [f for f in fields if f[0] in (abi.SIGMA, abi.ENERGY)], "after accretion")
        bodies = goldenrun.bodies_at(meta, k - 1, omega)
        for ctx, loop in zip((gpu, cpu), loops):
            ctx.set_bodies(bodies)
            ctx.set_time(loop.time)
        if staged:
            for stage, args in STAGES:
                a = tuple(step_dt if x == "dt" else x for x in args)
                gpu.stage(stage, *a)
                cpu.stage(stage, *a)
                if stage in ("potential", "derived"):
                    continue
                f = [x for x in fields if x[0] in (abi.SIGMA, abi.ENERGY)]
                if stage == "transport" or (stage == "boundary" and args[-1] == 1):
                    f = fields
                nbad += diff(gpu, cpu, f, f"after {stage}{args[-1:] if stage == 'boundary' else ''}")
        else:
            gpu.step(step_dt)
            cpu.step(step_dt)
            nbad += diff(gpu, cpu, fields, "after step")
        for loop in loops:
            loop.time += step_dt
            loop.n_iter += 1
            loop.n_monitor += 1
        if nbad >= 3:
            break


for name in (sys.argv[1:] or ["iso_accrete_20"]):
    run(name, accrete=True, staged=False)
    run(name, accrete=False, staged=False)
    run(name, accrete=True, staged=True, max_steps=3)
