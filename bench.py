#!/usr/bin/env python
"""bench.py — hydro cell-updates/s (FP64) of the B200-native FargoCPT hydro step.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU build (oracle/_ref)

A "step" is one iteration of sim::run's loop (simulation.cpp:515-553): CalculateTimeStep (CFL reduction, dt read
back by the host) followed by the gas part of step_Euler (sources, artificial + physical viscosity, SubStep3,
boundaries, Transport, halo exchange, damping).  Workload at every N: BASELINE.json configs[4], the
8192 x 16384 adiabatic planet-disk (strong scaling: the global grid is fixed and split radially over N GPUs).
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import math
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

METRIC = "hydro cell-updates/s (FP64)"
E2E_INTERVALS = 3  # snapshot intervals in the end-to-end leg
UNIT = "cell-updates/s"
# algorithmic bytes per cell-update, SURVEY.md §8d / DESIGN.md: adiabatic 400 B (+96 f in damping zones)
B_ALG = {"adiabatic_planet": 400.0, "cold_disk_planet": 312.0, "isothermal_planet": 280.0}
# algorithmic bytes per cell of each kernel (reads + writes of live state arrays only; DESIGN.md "Kernels")
KERNEL_BYTES = {
    "k_transport_azimuthal<ADI>": 88.0, "k_transport_azimuthal<ISO>": 72.0,
    "(k_transport_radial<LIM, true>)": 80.0, "(k_transport_radial<LIM, false>)": 64.0,
    "(k_fused_sources<ADI, false>)": 56.0, "k_fused_artvisc<ADI>": 56.0, "k_fused_viscosity<ADI>": 88.0,  # 72 + Sigma0, e0 (beta cooling)
    "k_potential": 24.0, "k_sources_velocity": 56.0, "k_compression_heating": 32.0, "k_artvisc_q": 56.0,
    "k_artvisc_v": 56.0, "k_viscosity_nu": 24.0, "k_stress": 64.0, "k_viscosity_v": 64.0, "k_substep3": 96.0,
    "k_cfl": 48.0, "k_ring_mean[cfl]": 8.0, "k_ring_mean[transport,side-stream]": 8.0,
}
# measured DRAM traffic per cell of the same kernels: dram__bytes_read.sum + dram__bytes_write.sum of one
# `ncu --set full` capture at the bench size (profiles/r02_v13_ncu_full_c5_8192x16384.md, 8192 x 16384 adiabatic_planet) / 134.2e6
# cells.  The azimuthal kernel's 100 B against 88 algorithmic: the 10 overlap columns of neighbouring 64-column windows are staged
# by different warps at the same moment, so they come from DRAM twice (the kernel is FP64-bound, DRAM at 33 %).
NCU_TRAFFIC_B_PER_CELL = {
    "k_transport_azimuthal<ADI>": 100.4, "(k_transport_radial<LIM, true>)": 81.3, "(k_fused_sources<ADI, false>)": 57.4,
    "k_fused_artvisc<ADI>": 57.3, "k_fused_viscosity<ADI>": 89.7, "k_cfl": 48.1,  # k_cfl: the screen pass (k_cfl_screen)
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([x.strip() for x in line.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def global_nrad(args, world):
    """Strong scaling: the global grid is fixed.  Weak scaling: 1024 * g x Naz, a fixed slab per GPU."""
    return args.nrad if args.scaling == "strong" else args.weak_nrad_per_gpu * world


def workload_config(args, nrad=None):
    from fargocpt_b200 import synthetic
    # FirstDT: the reference evaluates Q- for its first CFL before the beta-cooling reference state exists (NaN, SourceEuler.cpp:284),
    # so its first step is CFLmaxVar^3 * FirstDT whatever the CFL says; a small FirstDT keeps both arms' disks physical
    return synthetic.make_config(args.physics, nrad or args.nrad, args.naz, FirstDT=1.0e-5)


# ------------------------------------------------------------------------------------------------------
def run_reference(args):
    """Times the UNMODIFIED reference (oracle/_ref/fargocpt_exe_fast: the reference's own sources and -Ofast flags, OpenMP over
    all host cores; np = 1 because the image has no MPI) through its stock `start` code path, on the bench workload itself when
    the host has the memory for it (the reference allocates 76 grids: 82 GB at 8192 x 16384), else on an annulus of the same grid
    (same dr/r, same physics, fewer rings).  Per-step wall times are the reference's own: `LogAfterSteps: 1` makes
    logging::print_runtime_info (logging.cpp:204-262) print `timeperstep` after every hydro step; the first W are warm-up."""
    import yaml
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(args.gpus, 1)
    nrad_g = global_nrad(args, world)
    exe = os.path.join(ROOT, "oracle", "_ref", "fargocpt_exe_fast")
    kind = "reference"
    cfg = workload_config(args, nrad_g)
    cores = os.cpu_count() or 1
    steps, warm = max(args.steps, 1), max(args.warmup, 0)
    nrad_s = args.ref_nrad
    if nrad_s <= 0:  # auto: the whole grid if it fits in RAM (0.62 kB per cell: 76 grids + slack), else 1/8 of the rings
        try:
            avail = int(re.search(r"MemAvailable:\s+(\d+)", open("/proc/meminfo").read()).group(1)) * 1024
        except Exception:  # noqa: BLE001
            avail = 0
        nrad_s = nrad_g if nrad_g * args.naz * 8 * 80 < 0.8 * avail else max(nrad_g // 8, 64)
    nrad_s = min(nrad_s, nrad_g)
    # keep the logarithmic cell aspect ratio: same dr/r, annulus centred on the planet orbit
    if nrad_s == nrad_g:
        rmin_s, rmax_s = float(cfg["Rmin"]), float(cfg["Rmax"])
    else:
        growth = math.pow(float(cfg["Rmax"]) / float(cfg["Rmin"]), 1.0 / (nrad_g - 2.0))
        half = growth ** ((nrad_s - 2) / 2.0)
        rmin_s, rmax_s = 1.0 / half, half
    if not os.path.exists(exe):
        # fall back to the oracle port (the one other place bench.py may execute oracle/)
        val, sample = time_oracle_port(args, min(nrad_s, 256), rmin_s, rmax_s, steps + warm)
        nrad_s = min(nrad_s, 256)
        kind = "port"
        ms = nrad_s * args.naz / val * 1e3
    else:
        ycfg = {
            "DiskFeedback": "no", "MonitorTimestep": 1.0e9, "Nmonitor": 1, "Nsnapshots": 1,
            "l0": "30 au", "m0": "1 solMass", "SelfGravity": "No", "RadiativeDiffusion": "No", "Disk": "yes", "Frame": "F",
            "cps": -1, "DoWrite1DFiles": "No", "WriteAtEveryTimestep": "No", "RandomSigma": "No", "IntegrateParticles": "no",
            "HydroFrameCenter": "primary", "LogAfterSteps": 1, "LogAfterRealSeconds": 36000, "IndirectTermMode": 0,
            "ShockTube": 0, "WriteDensity": "no", "WriteVelocity": "no", "WriteEnergy": "no",  # no field output inside the run
        }
        for k, v in cfg.items():
            if k not in ("planet_mass",):
                ycfg[k] = v
        ycfg.update({"Nrad": nrad_s, "Naz": args.naz, "Rmin": rmin_s, "Rmax": rmax_s})
        nb = [{"name": "Star", "semi-major axis": 0.0, "mass": "1 solMass", "eccentricity": 0, "radius": "1 solRadius",
               "temperature": 0}]
        if cfg.get("planet_mass", 0) > 0:
            nb.append({"name": "planet", "semi-major axis": 1, "mass": float(cfg["planet_mass"]), "accretion efficiency": 0.0,
                       "eccentricity": 0.0, "radius": "0.01 solRadius", "temperature": "0 K", "ramp-up time": 0})
        ycfg["nbody"] = nb
        tmp = tempfile.mkdtemp(prefix="bench_ref_")
        ycfg["OutputDir"] = os.path.join(tmp, "out")
        ypath = os.path.join(tmp, "cfg.yml")
        yaml.safe_dump(ycfg, open(ypath, "w"), sort_keys=False)
        env = dict(os.environ, OMP_NUM_THREADS=str(cores), OMP_PROC_BIND="close", OMP_PLACES="cores")
        res = subprocess.run([exe, "-N", str(warm + steps), "start", ypath], cwd=tmp, env=env, capture_output=True, text=True)
        per_step = [float(x) for x in re.findall(r"hydrostep \d+, .*?timeperstep ([0-9.eE+-]+) ms", res.stdout)]
        if res.returncode != 0 or len(per_step) < warm + steps:
            raise RuntimeError("reference run failed: " + res.stdout[-800:] + res.stderr[-800:])
        timed = per_step[warm:warm + steps]
        ms = sum(timed) / len(timed)
        val = nrad_s * args.naz / (ms * 1e-3)
        what = ("the whole grid" if nrad_s == nrad_g else
                f"annulus r=[{rmin_s:.4f},{rmax_s:.4f}] of the {nrad_g}x{args.naz} grid (same dr/r, same physics)")
        sample = (f"unmodified reference, -Ofast -march=x86-64-v3 (oracle/Makefile.ref), np=1 (no MPI in image) x nt={cores}, "
                  f"{nrad_s}x{args.naz}: {what}; {len(timed)} timed hydro steps after {warm} warm-up steps, per-step wall time from the "
                  f"reference's own log (LogAfterSteps: 1)")
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, world), "cells": nrad_g * args.naz, "sample_nrad": nrad_s, "same_grid": nrad_s == nrad_g},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_name(args, world):
    nrad_g = global_nrad(args, world)
    which = "BASELINE configs[4]" if (args.physics, nrad_g, args.naz) == ("adiabatic_planet", 8192, 16384) else "BASELINE configs[4] physics"
    return (f"{args.physics} {nrad_g}x{args.naz} ({which}), " +
            ("fixed global grid" if args.scaling == "strong" else f"{args.weak_nrad_per_gpu} rings per GPU (weak scaling)"))


def time_oracle_port(args, nrad_s, rmin_s, rmax_s, nsteps):
    import reftools
    from fargocpt_b200 import synthetic
    cfg = synthetic.make_config(args.physics, nrad_s, args.naz, Rmin=rmin_s, Rmax=rmax_s)
    radii = synthetic.radii_from_config(cfg)
    params = synthetic.params_from_config(cfg)
    ctx = reftools.OracleContext(params, radii)
    t = drive(ctx, cfg, radii, nsteps, timed_from=1)
    return nrad_s * args.naz * (nsteps - 1) / t["wall"], f"oracle port, {nrad_s}x{args.naz} annulus, {nsteps - 1} steps"


def init_state(ctx, cfg, radii, fields=None):
    from fargocpt_b200 import abi, synthetic
    if fields is None:
        fields = synthetic.disk_fields(cfg, radii)
    ctx.upload(abi.SIGMA, fields["Sigma"])
    ctx.upload(abi.ENERGY, fields["energy"])
    ctx.upload(abi.VRAD, fields["vrad"])
    ctx.upload(abi.VAZI, fields["vazi"])
    orbit = synthetic.PlanetOrbit(cfg)
    ctx.set_bodies(orbit.bodies(0.0))
    ctx.set_time(0.0)
    ctx.stage("boundary", 0.0, 0)
    ctx.copy_initial_values()  # damping / beta-cooling reference state: before init_derived evaluates Q- against it for the first CFL
    ctx.init_derived()
    return orbit, fields


def drive(ctx, cfg, radii, nsteps, timed_from=0):
    """Host time loop (sim::run) on an oracle context; returns wall seconds of steps [timed_from, nsteps)."""
    orbit, _ = init_state(ctx, cfg, radii)
    last_dt, t = float(cfg["FirstDT"]), 0.0
    t0 = None
    for k in range(nsteps):
        if k == timed_from:
            t0 = time.time()
        dt = ctx.cfl(last_dt)
        last_dt = dt
        ctx.set_bodies(orbit.bodies(t, dt))
        ctx.set_time(t)
        ctx.step(dt)
        t += dt
    return {"wall": time.time() - t0}


# ------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from fargocpt_b200 import HydroContext, abi, synthetic

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf = torch.frombuffer(bytearray(abi.get_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(buf, 0)
        uid = bytes(buf.cpu().numpy().tobytes())

    nrad_g = global_nrad(args, world)
    cfg = workload_config(args, nrad_g)
    radii = synthetic.radii_from_config(cfg)
    params = synthetic.params_from_config(cfg)
    ctx = HydroContext(params, radii, rank=rank, nranks=world, unique_id=uid, device=local)
    ncell = nrad_g * args.naz

    # host copy of the initial state in PINNED memory (e2e leg uploads from it)
    fields = synthetic.disk_fields(cfg, radii)
    pinned = {}
    for k, v in fields.items():
        t = torch.from_numpy(v).pin_memory()
        pinned[k] = t.numpy()
    orbit, _ = init_state(ctx, cfg, radii, pinned)
    state = {"last_dt": float(cfg["FirstDT"]), "t": 0.0}

    def one_step():
        dt = ctx.cfl(state["last_dt"])
        state["last_dt"] = dt
        ctx.set_bodies(orbit.bodies(state["t"], dt))
        ctx.set_time(state["t"])
        ctx.step(dt)
        state["t"] += dt

    def barrier():
        ctx.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        one_step()
    # ---- device-resident timed region (value) -------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = ctx.launch_count()
    ctx.event_record(0)
    for _ in range(args.steps):
        one_step()
    ctx.event_record(1)
    ms = ctx.event_elapsed_ms(0, 1)
    barrier()
    launches = ctx.launch_count() - l0
    # per-kernel device times (CUDA events around every launch, on the launching stream) in a separate pass so that
    # the event records do not sit inside the timed region above
    prof_steps = min(args.steps, 5)
    ctx.profile(True)
    for _ in range(prof_steps):
        one_step()
    barrier()
    prof = ctx.profile_report()
    ctx.profile(False)
    if world > 1:
        tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    value = ncell * args.steps / (ms * 1e-3)
    # ---- checksum of the state after the W + K + P device-resident steps -------------------------------
    # The result of the reference does not depend on the number of ranks (constants.h:17) and neither does ours, bit for bit:
    # every N of a strong-scaling sweep must print the SAME checksum (same grid, same start, same number of steps) — the
    # driver-visible proof that the ghost-ring exchange (CommunicateBoundaries, commbound.cpp:98-182) delivers the right rings.
    checksum = state_checksum(ctx, world, torch, dist if world > 1 else None)
    steps_done = max(args.warmup, 3) + args.steps + prof_steps

    # ---- end-to-end leg: host buffers in, host buffers out -------------------------------------------
    # A restart + E2E_INTERVALS snapshot intervals as a user of the C ABI runs them: upload the four state fields from
    # pinned host memory, then per interval K steps (each with its dt read-back) and one snapshot of the four state fields
    # into pinned host memory (fargo_snapshot_async: the copy overlaps the next interval's steps); the region ends when
    # the last snapshot is on the host.  Bytes per step = totals / (E2E_INTERVALS * K).
    out_host = [{fid: torch.empty(ctx.global_shape(fid), dtype=torch.float64).pin_memory().numpy()
                 for fid in (abi.SIGMA, abi.VRAD, abi.VAZI, abi.ENERGY)} for _ in range(2)]
    barrier()
    ctx.event_record(2)
    t_e2e0 = time.time()
    ctx.upload(abi.SIGMA, pinned["Sigma"])
    ctx.upload(abi.ENERGY, pinned["energy"])
    ctx.upload(abi.VRAD, pinned["vrad"])
    ctx.upload(abi.VAZI, pinned["vazi"])
    state["t"] = 0.0
    for it in range(E2E_INTERVALS):
        for _ in range(args.steps):
            one_step()
        o = out_host[it % 2]
        ctx.snapshot_async(o[abi.SIGMA], o[abi.VRAD], o[abi.VAZI], o[abi.ENERGY])
    ctx.snapshot_wait()
    ctx.event_record(3)
    ms_e2e = ctx.event_elapsed_ms(2, 3)
    barrier()
    if world > 1:
        tms = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms_e2e = float(tms.item())
    # the same leg at a cadence nearer to production runs: one restart upload, 10 K steps, one snapshot,
    # region ends when the snapshot is on the host.  (The reference's own configs write a snapshot every 1e3-1e4 hydro steps.)
    long_steps = 10 * args.steps
    barrier()
    ctx.event_record(2)
    ctx.upload(abi.SIGMA, pinned["Sigma"])
    ctx.upload(abi.ENERGY, pinned["energy"])
    ctx.upload(abi.VRAD, pinned["vrad"])
    ctx.upload(abi.VAZI, pinned["vazi"])
    state["t"] = 0.0
    for _ in range(long_steps):
        one_step()
    o = out_host[0]
    ctx.snapshot_async(o[abi.SIGMA], o[abi.VRAD], o[abi.VAZI], o[abi.ENERGY])
    ctx.snapshot_wait()
    ctx.event_record(3)
    ms_e2e_long = ctx.event_elapsed_ms(2, 3)
    barrier()
    if world > 1:
        tms = torch.tensor([ms_e2e_long], dtype=torch.float64, device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms_e2e_long = float(tms.item())
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    slab_cells = ctx.nr * ctx.naz
    e2e_steps = E2E_INTERVALS * args.steps
    h2d = (4 * slab_cells + ctx.naz) * 8 / e2e_steps
    owned = sum(int(np.prod(ctx.global_shape(f))) for f in out_host[0]) / world
    d2h = owned * 8 / args.steps + 8  # one snapshot per interval + the dt scalar every step
    e2e_val = ncell * e2e_steps / (ms_e2e * 1e-3)

    if rank != 0:
        return
    # ---- roofline of the dominant kernel ---------------------------------------------------------------
    peaks, peak_kind = measured_peaks()
    top = max(prof.items(), key=lambda kv: kv[1][0]) if prof else (None, (0.0, 0))
    total_kernel_ms = sum(v[0] for v in prof.values())
    roof = None
    if top[0]:
        kname, (kms, kn) = top
        kname = kname.replace("[+halo push]", "")  # the multi-GPU launch of the same kernel (edge rings mirrored to the neighbours)
        bytes_per_cell = KERNEL_BYTES.get(kname, 0.0)
        per_launch_bytes = bytes_per_cell * slab_cells
        avg_s = kms / max(kn, 1) * 1e-3
        achieved = per_launch_bytes / avg_s / 1e9 if avg_s > 0 else 0.0
        roof = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)",
                "traffic": (NCU_TRAFFIC_B_PER_CELL[kname] * slab_cells if kname in NCU_TRAFFIC_B_PER_CELL else None),
                "traffic_source": "ncu --set full, profiles/r02_v13_ncu_full_c5_8192x16384.md (bytes per cell x cells of this launch)",
                "bytes_per_launch_algorithmic": per_launch_bytes, "avg_launch_ms": avg_s * 1e3,
                "share_of_step": kms / total_kernel_ms if total_kernel_ms else None}
    step_roof = B_ALG[args.physics] * value / world / 1e9  # whole-step algorithmic GB/s per GPU
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "checksum": {"sha256": checksum, "after_steps": steps_done,
                     "of": "Sigma, v_rad, v_azi, e of the whole grid: per-ring 64-bit sums of the bit patterns (plain and column-weighted), "
                           "owned rings of every rank stitched in ring order; identical for every N of a strong-scaling sweep"},
        "config": {"workload": workload_name(args, world) + f", radial slabs over {world} GPU(s)",
                   "cells": ncell, "l2": "inputs larger than L2 (1.07 GB per field)" if ncell * 8 > 126e6 else "inputs fit L2",
                   "b_alg_bytes_per_cell_update": B_ALG[args.physics],
                   "halo_exchange": {0: "none (1 GPU)", 1: "ncclSend/ncclRecv after Transport",
                                     2: "NVLink peer-memory stores from the transport kernel's edge launch, overlapped with the interior rings"}[ctx.halo_mode()]},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "what": f"restart (upload 4 state fields from pinned host memory) + {E2E_INTERVALS} intervals of {args.steps} steps, each ending in an "
                        "asynchronous snapshot of the 4 state fields into pinned host memory (overlaps the next interval); ends when the last snapshot is on the host. "
                        "A stress cadence: PCIe-bound once a step takes a few ms (N > 2); the reference's own configs write a snapshot every 1e3-1e4 hydro "
                        "steps, where the end-to-end rate is `value` (every step of `value` already reads its dt back through the ABI)",
                "at_one_snapshot_per_10K_steps": {
                    "value": ncell * long_steps / (ms_e2e_long * 1e-3), "unit": UNIT, "steps": long_steps,
                    "h2d_bytes_per_step": (4 * slab_cells + ctx.naz) * 8 / long_steps, "d2h_bytes_per_step": owned * 8 / long_steps + 8,
                    "what": f"the same leg with one restart upload, {long_steps} steps and one snapshot (still 5-50x the reference configs' cadence)"}},
        "gpu_launches": int(launches),
        "roofline": roof,
        "step_roofline": {"achieved_gbs_per_gpu": step_roof, "frac_of_measured_peak": step_roof / peaks["hbm_gbs"],
                          "frac_of_8tbs": step_roof / 8000.0},
        "kernels_ms_per_step": {k: round(v[0] / prof_steps, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    if world == 1 and not args.no_tolerance_mode and args.scaling == "strong":
        line["tolerance_mode"] = tolerance_mode(args)
    print(json.dumps(line))


def state_checksum(ctx, world, torch, dist):
    """SHA-256 over per-ring hashes of the four state fields (partition-independent: a rank contributes the rings it owns)."""
    import hashlib
    from fargocpt_b200 import abi
    parts = []
    for fid in (abi.SIGMA, abi.VRAD, abi.VAZI, abi.ENERGY):
        a = ctx.download(fid)  # global shape; only the rings this rank owns are written, the rest stays 0
        bits = a.view(np.uint64)
        w = (2 * np.arange(bits.shape[1], dtype=np.uint64) + 1)[None, :]
        h = np.stack([bits.sum(axis=1, dtype=np.uint64), (bits * w).sum(axis=1, dtype=np.uint64)], axis=1)  # wraps mod 2^64
        del a, bits
        if dist is not None:
            t = torch.from_numpy(h.view(np.int64)).cuda()
            dist.all_reduce(t)  # owned ring sets are disjoint, the other rows are 0: the (wrapping) sum stitches them
            h = t.cpu().numpy().view(np.uint64)
        parts.append(np.ascontiguousarray(h))
    return hashlib.sha256(b"".join(p.tobytes() for p in parts)).hexdigest()


TOL_LIB = os.path.join(ROOT, "fargocpt_b200", "csrc", "libfargo_b200_tol.so")


def run_leg(args):
    """Helper legs of the tolerance-mode report, each in its own process (a process loads ONE build of the library):
    --leg time: ms/step of K device-resident steps after W warm-up steps;  --leg dump: N steps from the synthetic start, final
    fields + dt sequence + Nshift of every step written to --dump."""
    import torch
    from fargocpt_b200 import HydroContext, abi, synthetic
    torch.cuda.set_device(0)
    cfg = workload_config(args)
    radii = synthetic.radii_from_config(cfg)
    params = synthetic.params_from_config(cfg)
    ctx = HydroContext(params, radii)
    orbit, _ = init_state(ctx, cfg, radii)
    state = {"last_dt": float(cfg["FirstDT"]), "t": 0.0}
    dts, shifts = [], []

    def one_step(record=False):
        dt = ctx.cfl(state["last_dt"])
        state["last_dt"] = dt
        ctx.set_bodies(orbit.bodies(state["t"], dt))
        ctx.set_time(state["t"])
        ctx.step(dt)
        state["t"] += dt
        if record:
            dts.append(dt)
            shifts.append(ctx.nshift().copy())
    if args.leg == "time":
        for _ in range(max(args.warmup, 3)):
            one_step()
        ctx.sync()
        ctx.event_record(0)
        for _ in range(args.steps):
            one_step()
        ctx.event_record(1)
        ms = ctx.event_elapsed_ms(0, 1)
        print(json.dumps({"leg": "time", "lib": os.path.basename(abi.LIB_PATH), "ms_per_step": ms / args.steps}))
    else:
        for _ in range(args.steps):
            one_step(record=True)
        np.savez(args.dump, dts=np.array(dts), shifts=np.array(shifts), Sigma=ctx.download(abi.SIGMA), vrad=ctx.download(abi.VRAD),
                 vazi=ctx.download(abi.VAZI), energy=ctx.download(abi.ENERGY))
        print(json.dumps({"leg": "dump", "lib": os.path.basename(abi.LIB_PATH), "steps": args.steps}))


def tolerance_mode(args):
    """The tolerance build (fargo_math.h, -DFARGO_TOL: quotients without the final correction step, no validity keys, fused
    multiply-adds in the transport kernels) timed on the bench workload, and its measured deviation from the exact build (which is
    bit-identical to the reference) after 100 CFL-limited steps on a 1024 x 2048 grid of the same physics.  A second number,
    never `value`."""
    if not os.path.exists(TOL_LIB):
        return {"error": "libfargo_b200_tol.so not built"}
    me = [sys.executable, os.path.abspath(__file__), "--physics", args.physics]

    def leg(lib, extra):
        env = dict(os.environ)
        if lib:
            env["FARGO_B200_LIB"] = lib
        else:
            env.pop("FARGO_B200_LIB", None)
        res = subprocess.run(me + extra, env=env, capture_output=True, text=True, timeout=900)
        for ln in res.stdout.splitlines()[::-1]:
            if ln.startswith("{"):
                return json.loads(ln)
        raise RuntimeError((res.stdout + res.stderr)[-600:])
    try:
        t = leg(TOL_LIB, ["--leg", "time", "--nrad", str(args.nrad), "--naz", str(args.naz), "--steps", str(args.steps),
                          "--warmup", str(args.warmup)])
        tmp = tempfile.mkdtemp(prefix="bench_tol_")
        nsteps, nr, na = 100, 1024, 2048
        for name, lib in (("exact", None), ("tol", TOL_LIB)):
            leg(lib, ["--leg", "dump", "--nrad", str(nr), "--naz", str(na), "--steps", str(nsteps), "--dump", os.path.join(tmp, name + ".npz")])
        a, b = np.load(os.path.join(tmp, "exact.npz")), np.load(os.path.join(tmp, "tol.npz"))
        dev = {f: float(np.abs(a[f] - b[f]).max() / np.abs(a[f]).max()) for f in ("Sigma", "vrad", "vazi", "energy")}
        dev["vrad_relative_to_sound_speed_scale"] = float(np.abs(a["vrad"] - b["vrad"]).max() / np.sqrt((0.4 * a["energy"][1:-1] / a["Sigma"][1:-1] * 1.4).max()))
        out = {"ms_per_step": t["ms_per_step"], "value": args.nrad * args.naz / (t["ms_per_step"] * 1e-3), "unit": UNIT,
               "deviation_from_exact_after_steps": nsteps, "deviation_grid": [nr, na],
               "max_abs_deviation_over_field_max": dev,
               "dt_step0_bit_equal": bool(a["dts"][0] == b["dts"][0]),
               "dt_max_rel_deviation": float(np.abs(a["dts"] / b["dts"] - 1.0).max()),
               "nshift_steps_equal": int((a["shifts"] == b["shifts"]).all(axis=1).sum()), "nshift_steps": nsteps,
               "what": "second build of the library (-DFARGO_TOL), see fargocpt_b200/csrc/fargo_math.h; the exact build stays the default and the parity gate"}
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
        return out
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[-600:]}


def cpu_baseline(args):
    """The reference arm on a bounded sample (256 rings of the same grid: 10-30 s of CPU work), run in a subprocess so its
    threads do not disturb this process."""
    try:
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "6", "--warmup", "2",
                              "--physics", args.physics, "--nrad", str(args.nrad), "--naz", str(args.naz),
                              "--ref-nrad", str(args.cpu_sample_nrad)], capture_output=True, text=True, timeout=900)
        for ln in res.stdout.splitlines()[::-1]:
            if ln.startswith("{"):
                return json.loads(ln)["cpu_baseline"]
        return {"error": (res.stdout + res.stderr)[-400:]}
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--physics", default="adiabatic_planet", choices=list(B_ALG))
    ap.add_argument("--nrad", type=int, default=8192)
    ap.add_argument("--naz", type=int, default=16384)
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: the global grid --nrad x --naz is split over the GPUs; weak: --weak-nrad-per-gpu rings per GPU")
    ap.add_argument("--weak-nrad-per-gpu", type=int, default=1024)
    ap.add_argument("--ref-nrad", type=int, default=0,
                    help="--impl reference: rings of the grid the reference runs (0 = the whole grid if the host has the RAM, else 1/8)")
    ap.add_argument("--cpu-sample-nrad", type=int, default=256, help="rings of the annulus of the cpu_baseline leg of the GPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tolerance-mode", action="store_true")
    ap.add_argument("--leg", default=None, choices=["time", "dump"], help="internal: helper legs of the tolerance-mode report")
    ap.add_argument("--dump", default=None)
    args = ap.parse_args()
    if args.leg:
        run_leg(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
