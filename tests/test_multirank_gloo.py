"""world_size-2 CPU test (gloo) of the multi-rank host logic: radial slabs as SplitDomain makes them (split.cpp:21-87),
the 7-ring ghost exchange of CommunicateBoundaries (commbound.cpp:98-182) and the dt all-reduce (cfl.cpp:379), driven
exactly like the GPU ranks are (one process per slab, torch.distributed for the plumbing) but with the CPU oracle as
the slab engine.  The 2-rank result must equal the 1-rank result bit for bit (the reference's np-independence,
constants.h:17).  This pins the slab / halo semantics the CUDA path implements with ncclSend/ncclRecv."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, name, nsteps, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import goldenrun
    import reftools
    from fargocpt_b200 import abi

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    meta, z = reftools.load_golden(name)
    params = reftools.make_params(meta["params"])
    ctx = reftools.OracleContext(params, z["radii"], rank=rank, nranks=world)
    for fid, fname in goldenrun.STATE:
        ctx.upload(fid, z[fname + "_0"])
    omega = float(meta["config"].get("OmegaFrame", 0.0))
    ctx.set_bodies(goldenrun.bodies_at(meta, 0, omega))
    ctx.set_time(0.0)
    ctx.init_derived()
    ctx.copy_initial_values()
    ctx.stage("boundary", 0.0, 0)
    n = abi.CPUOVERLAP * ctx.naz

    def exchange():
        # even/odd ordered pairwise exchange like commbound.cpp:130-158 (blocking gloo send/recv)
        for side, peer in ((0, rank - 1), (1, rank + 1)):
            if peer < 0 or peer >= world:
                continue
            out = torch.from_numpy(ctx.halo_pack(side))
            inc = torch.zeros(4 * n, dtype=torch.float64)
            if rank % 2 == 0:
                dist.send(out, peer)
                dist.recv(inc, peer)
            else:
                dist.recv(inc, peer)
                dist.send(out, peer)
            ctx.halo_unpack(side, inc.numpy())

    accretors = [b for b, rec in enumerate(meta["bodies"][0]) if len(rec) > 7 and rec[7] > 0.0]
    ctx.track_massflow(True)  # MASSFLOW grid, MassDelta's boundary flows and wave-damping terms: every slab keeps its own share
    ctx.track_boundary_flow(True)
    ctx.track_damping_mass(True)
    last_dt, t, dts, accreted = meta["first_dt"], 0.0, [], []
    for k in range(nsteps):
        loc = torch.tensor([ctx.condition_cfl()], dtype=torch.float64)
        dist.all_reduce(loc, op=dist.ReduceOp.MIN)  # MPI_Allreduce(MIN), cfl.cpp:379
        dt = min(params.cfl_max_var * last_dt, float(loc.item()))  # simulation.cpp:100-118
        last_dt = dt
        dts.append(dt)
        for b in accretors:  # AccreteOntoPlanets first thing in the step; every slab takes its own cells, MPI_Allreduce(SUM) of
            # what the ACTIVE cells gave (accretion.cpp:199-213) — the overlap rings change on both ranks but count once
            method = meta["config"]["nbody"][b].get("accretion method", "kley")
            d = torch.tensor(ctx.accrete_kley(*goldenrun.accretion_inputs(meta, k, b, dt), method=method), dtype=torch.float64)
            dist.all_reduce(d)
            accreted.append(d.tolist())
        ctx.set_bodies(goldenrun.bodies_at(meta, k, omega))
        ctx.set_time(t)
        ctx.step_pre(dt)
        exchange()
        ctx.step_post(dt)
        t += dt
    res = {}
    for fid, fname in goldenrun.STATE:
        part = torch.from_numpy(ctx.download(fid))  # owned rings only, zeros elsewhere
        dist.all_reduce(part)
        res[fname] = part.numpy()
    part = torch.from_numpy(ctx.download(abi.MASSFLOW))
    dist.all_reduce(part)
    res["MassFlow"] = part.numpy()
    sums = torch.tensor(list(ctx.damping_mass()) + list(ctx.boundary_flow()), dtype=torch.float64)  # MPI_Reduce(SUM), output.cpp:438-453
    dist.all_reduce(sums)
    res["mass_delta"] = sums.numpy()
    if rank == 0:
        q.put((dts, res, accreted))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["iso_planet_100", "adia_planet_100", "iso_accrete_20", "iso_sinkhole_20"])
def test_two_ranks_equal_one_rank(name):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import goldenrun
    import reftools
    nsteps = 8
    ctx_mp = mp.get_context("spawn")
    q = ctx_mp.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx_mp.Process(target=_worker, args=(r, 2, port, name, nsteps, q)) for r in range(2)]
    for p in procs:
        p.start()
    dts2, res2, acc2 = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single rank, same driver
    meta, z = reftools.load_golden(name)
    params = reftools.make_params(meta["params"])
    one = reftools.OracleContext(params, z["radii"])
    for fid, fname in goldenrun.STATE:
        one.upload(fid, z[fname + "_0"])
    omega = float(meta["config"].get("OmegaFrame", 0.0))
    one.set_bodies(goldenrun.bodies_at(meta, 0, omega))
    one.set_time(0.0)
    one.init_derived()
    one.copy_initial_values()
    one.stage("boundary", 0.0, 0)
    accretors = [b for b, rec in enumerate(meta["bodies"][0]) if len(rec) > 7 and rec[7] > 0.0]
    one.track_massflow(True)
    one.track_boundary_flow(True)
    one.track_damping_mass(True)
    last_dt, t, dts1, acc1 = meta["first_dt"], 0.0, [], []
    for k in range(nsteps):
        dt = min(params.cfl_max_var * last_dt, one.condition_cfl())
        last_dt = dt
        dts1.append(dt)
        for b in accretors:
            method = meta["config"]["nbody"][b].get("accretion method", "kley")
            acc1.append(list(one.accrete_kley(*goldenrun.accretion_inputs(meta, k, b, dt), method=method)))
        one.set_bodies(goldenrun.bodies_at(meta, k, omega))
        one.set_time(t)
        one.step(dt)
        t += dt
    assert dts1 == dts2
    assert len(acc1) == len(acc2) == (nsteps if accretors else 0)
    # The Hill sphere straddles the cut between the two slabs on this grid.  The fields below are np-independent (both slabs change
    # their copies of the overlap rings identically), and so is what the planet is credited with: the reference's condition
    # `radial_first_active < i` (accretion.cpp:186-187, strictly) would also drop the first OWNED ring of rank 1; the slabs count it,
    # so two slabs sum to the one-slab (np = 1) value up to the order of the additions.
    for a, b in zip(acc1, acc2):
        assert a[0] > 0 and abs(b[0] - a[0]) <= 1e-12 * a[0], (a, b)
    for fid, fname in goldenrun.STATE:
        if fname == "energy" and not params.adiabatic:
            continue
        st = reftools.compare_stats(res2[fname], one.download(fid))
        assert st["n_diff"] == 0, (fname, st)
    # the mass-flow grid is np-independent ring by ring; MassDelta's sums only count active cells (sum_without_ghost_cells), so the
    # slabs' shares add up to the one-slab value up to the order of the additions
    from fargocpt_b200 import abi
    assert reftools.compare_stats(res2["MassFlow"], one.download(abi.MASSFLOW))["n_diff"] == 0
    want = np.array(list(one.damping_mass()) + list(one.boundary_flow()))
    if params.damping:
        assert np.abs(want[:4]).max() > 0
    assert np.allclose(res2["mass_delta"], want, rtol=1e-12, atol=0.0), (res2["mass_delta"], want)


def test_slab_partition_matches_split_domain():
    """split.cpp:38-61: contiguous slabs, +1 ring for the first `remainder` ranks, 7 overlap rings on interior sides."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import reftools
    meta, z = reftools.load_golden("iso_planet_100")
    params = reftools.make_params(meta["params"])
    nrad = params.nrad
    for world in (2, 3):
        owned = np.zeros(nrad, dtype=int)
        for r in range(world):
            c = reftools.OracleContext(params, z["radii"], rank=r, nranks=world)
            low, rem = nrad // world, nrad % world
            size = low + 1 if r < rem else low
            start = (low + 1) * r if r < rem else (low + 1) * rem + (r - rem) * low
            imin = start - (7 if r > 0 else 0)
            assert c.imin == imin and c.nr == size + (7 if r > 0 else 0) + (7 if r < world - 1 else 0)
            owned[start:start + size] += 1
            c.close()
        assert (owned == 1).all()
