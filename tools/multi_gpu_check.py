#!/usr/bin/env python
"""N-GPU vs 1-GPU parity of the radial-slab path (split.cpp:21-87 + commbound.cpp:98-182 + the dt all-reduce).

Run under torchrun, one rank per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py [--physics isothermal_planet] [--nrad 256] [--naz 512] [--steps 12]
Every rank runs its slab for K CFL-limited steps; the owned rings are stitched with an all-reduce and rank 0
compares them with a single-slab run of the same library on its own GPU.  The reference states that its result
does not depend on the number of ranks (constants.h:17); so: dt sequence bit-equal, fields bit-equal.
Prints one JSON line; exit code 1 on mismatch.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run_slab(ctx, cfg, radii, fields, nsteps, synthetic, abi):
    ctx.upload(abi.SIGMA, fields["Sigma"])
    ctx.upload(abi.ENERGY, fields["energy"])
    ctx.upload(abi.VRAD, fields["vrad"])
    ctx.upload(abi.VAZI, fields["vazi"])
    orbit = synthetic.PlanetOrbit(cfg)
    ctx.set_bodies(orbit.bodies(0.0))
    ctx.set_time(0.0)
    ctx.stage("boundary", 0.0, 0)
    ctx.copy_initial_values()  # before init_derived: Q- of the first CFL is evaluated against the beta-cooling reference state
    ctx.init_derived()
    ctx.track_massflow(True)      # the tracking instantiations of the radial sweep and Sigma's own damping pass, slab by slab
    ctx.track_damping_mass(True)
    last_dt, t, dts = float(cfg["FirstDT"]), 0.0, []
    for _ in range(nsteps):
        dt = ctx.cfl(last_dt)
        last_dt = dt
        dts.append(dt)
        ctx.set_bodies(orbit.bodies(t, dt))
        ctx.set_time(t)
        ctx.step(dt)
        t += dt
    out = {}
    for fid, name in ((abi.SIGMA, "Sigma"), (abi.VRAD, "vrad"), (abi.VAZI, "vazi"), (abi.ENERGY, "energy")):
        out[name] = ctx.download(fid)  # only the rings this rank owns are written, the rest stays 0
    # the monitor reductions are collective: global sums (all-reduced) and the per-ring sums behind disk radius / eccentricity
    mon = dict(ctx.monitor_quantities(), **{"disk_" + k: v for k, v in ctx.monitor_disk(1e300, 0.99, 0.3).items()})
    mon["circumplanetary_mass"] = ctx.circumplanetary_mass(0.9, 0.3, 0.3)
    for k, v in zip(("inner_creation", "inner_removal", "outer_creation", "outer_removal"), ctx.damping_mass()):
        mon["damping_" + k] = v  # per-column sums over the rings of a rank, then over the ranks: rounding
    for k, v in zip(("inner_in", "inner_out", "outer_in", "outer_out"), ctx.boundary_flow()):
        mon["boundary_" + k] = v
    out["MassFlow"] = ctx.download(abi.MASSFLOW)
    return dts, out, mon


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--physics", default="isothermal_planet")
    ap.add_argument("--nrad", type=int, default=256)
    ap.add_argument("--naz", type=int, default=512)
    ap.add_argument("--steps", type=int, default=12)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from fargocpt_b200 import HydroContext, abi, synthetic

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = torch.frombuffer(bytearray(abi.get_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(buf, 0)
    uid = bytes(buf.cpu().numpy().tobytes())

    cfg = synthetic.make_config(args.physics, args.nrad, args.naz)
    radii = synthetic.radii_from_config(cfg)
    params = synthetic.params_from_config(cfg)
    fields = synthetic.disk_fields(cfg, radii, perturb=2e-2)
    fields["vrad"] = fields["vrad"] + 1e-3 * np.sin(np.arange(args.naz) * 2 * np.pi * 3 / args.naz)[None, :]

    ctx = HydroContext(params, radii, rank=rank, nranks=world, unique_id=uid, device=local)
    halo_mode = ctx.halo_mode()
    dts, out, mon = run_slab(ctx, cfg, radii, fields, args.steps, synthetic, abi)
    ctx.close()
    stitched = {}
    for name, a in out.items():
        t = torch.from_numpy(a).cuda()
        dist.all_reduce(t)  # owned ring sets are disjoint, the rest is exactly 0: the sum stitches bit-exactly
        stitched[name] = t.cpu().numpy()
    rc = 0
    if rank == 0:
        one = HydroContext(params, radii, rank=0, nranks=1, device=local)
        dts1, out1, mon1 = run_slab(one, cfg, radii, fields, args.steps, synthetic, abi)
        import reftools
        res = {"n_gpus": world, "physics": args.physics, "grid": [args.nrad, args.naz], "steps": args.steps,
               "dt_bit_equal": dts == dts1, "fields": {},
               "halo_mode": {1: "nccl send/recv", 2: "peer-memory stores from the transport kernel"}.get(halo_mode, halo_mode)}
        adiabatic = bool(params.adiabatic)
        for name in stitched:
            if name == "energy" and not adiabatic:
                continue
            st = reftools.compare_stats(stitched[name], out1[name])
            res["fields"][name] = st
            if st["n_diff"] != 0:
                rc = 1
        if not res["dt_bit_equal"]:
            rc = 1
        # per-ring sums are formed by the same kernel in the same order on whichever rank owns the ring and added in ring order:
        # identical; the global sums of fargo_monitor_quantities are added per rank first: rounding
        same = lambda a, b: a == b or (a != a and b != b)  # the potential columns are NaN on both sides unless a kick kept the grid
        res["monitor_disk_equal"] = all(same(mon[k], mon1[k]) for k in mon if k.startswith("disk_"))
        res["monitor_sums_max_rel_dev"] = max(abs(mon[k] - mon1[k]) / max(abs(mon1[k]), 1e-300) for k in mon if not k.startswith("disk_"))
        if not res["monitor_disk_equal"] or res["monitor_sums_max_rel_dev"] > 1e-12:
            rc = 1
            res["monitor"] = {k: (mon[k], mon1[k]) for k in mon}
        res["ok"] = rc == 0
        print(json.dumps(res))
    flag = torch.tensor([rc], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(int(flag.item()))


if __name__ == "__main__":
    main()
