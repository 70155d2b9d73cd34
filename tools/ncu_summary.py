#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` report (read on the CPU box).

    ncu_summary.py report.ncu-rep [cells] > profiles/xxx.md

Table 1: time, DRAM bytes and DRAM throughput % of peak, issue-slot utilisation, FP64-pipe utilisation, occupancy,
registers, eligible warps per cycle, L1 / L2 hit rates.  Table 2 (when `cells` is given: cells per launch): warp and thread
instructions per cell, FP64 thread instructions per cell (DADD + DMUL + DFMA), the share of FP64 in all issued instructions.
Table 3: the top-3 warp stall reasons (average warps stalled per issue-active cycle).  Then, per distinct kernel, the dynamic
opcode mix from the source page (instructions executed per cell, top opcodes)."""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
cells = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: k for k, n in enumerate(hdr)}


def num(r, m, default=float("nan")):
    if m not in col:
        return default
    try:
        return float(r[col[m]].replace(",", ""))
    except ValueError:
        return default


def scaled(r, m, table):
    v = num(r, m)
    return v * table.get(units[col[m]].lower(), 1.0) if m in col else float("nan")


TIME = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "nsecond": 1e-6, "second": 1e3}
BYTES = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}


def kname(r):
    return re.sub(r"^void ", "", re.sub(r"\(.*", "", r[col["Kernel Name"]]))


def fmt(x, d=1):
    return "-" if x != x else (f"{x:.{d}f}")


print(f"source: {rep} (`ncu --set full --clock-control none`, one row per launch; durations are under replay: cold cache, serialised)\n")
print("| id | kernel | ms | DRAM rd MB | DRAM wr MB | DRAM throughput % of peak | issue slots busy % | FP64 pipe % | occupancy % | regs | "
      "eligible warps / cycle / scheduler | L1 hit % | L2 hit % |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in data:
    print(f"| {r[col['ID']]} | {kname(r)} | {fmt(scaled(r, 'gpu__time_duration.sum', TIME), 3)} | "
          f"{fmt(scaled(r, 'dram__bytes_read.sum', BYTES) / 1e6)} | {fmt(scaled(r, 'dram__bytes_write.sum', BYTES) / 1e6)} | "
          f"{fmt(num(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'))} | "
          f"{fmt(num(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'))} | "
          f"{fmt(num(r, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'))} | "
          f"{fmt(num(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'))} | {fmt(num(r, 'launch__registers_per_thread'), 0)} | "
          f"{fmt(num(r, 'smsp__warps_eligible.avg.per_cycle_active'), 2)} | {fmt(num(r, 'l1tex__t_sector_hit_rate.pct'))} | "
          f"{fmt(num(r, 'lts__t_sector_hit_rate.pct'))} |")

if cells:
    print(f"\nInstruction budget per cell ({cells:.0f} cells per launch; kernels that touch a subset of the rings are still divided by all cells):\n")
    print("| id | kernel | warp inst x 32 / cell | thread inst / cell | FP64 thread inst / cell (DADD + DMUL + DFMA) | FP64 share of thread inst % | DRAM bytes / cell |")
    print("|---|---|---|---|---|---|---|")
    tot = collections.Counter()
    seen = set()
    for r in data:
        winst = num(r, "smsp__inst_executed.sum")
        tinst = num(r, "thread_inst_executed_true", num(r, "smsp__thread_inst_executed.sum"))
        cyc = num(r, "smsp__cycles_elapsed.avg") if "smsp__cycles_elapsed.avg" in col else float("nan")
        # per-cycle-elapsed FP64 op rates x elapsed cycles = totals (the report stores the rates)
        f64 = float("nan")
        ops = ["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
               "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
               "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed"]
        if all(o in col for o in ops) and cyc == cyc:
            f64 = sum(num(r, o) for o in ops) * cyc
        dram = scaled(r, "dram__bytes_read.sum", BYTES) + scaled(r, "dram__bytes_write.sum", BYTES)
        print(f"| {r[col['ID']]} | {kname(r)} | {fmt(winst * 32 / cells)} | {fmt(tinst / cells)} | {fmt(f64 / cells)} | "
              f"{fmt(100 * f64 / tinst) if tinst == tinst and tinst > 0 else '-'} | {fmt(dram / cells)} |")
        if kname(r) not in seen:
            seen.add(kname(r))
            tot["w"] += winst * 32 / cells
            tot["t"] += tinst / cells
            tot["f"] += (f64 / cells) if f64 == f64 else 0.0
    print(f"\nSum over the distinct kernels of the step: {tot['w']:.0f} (warp inst x 32) / {tot['t']:.0f} thread instructions per cell-step, "
          f"of which {tot['f']:.0f} FP64.\n")

STALL = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
print("\nTop warp stall reasons (average warps per scheduler stalled for that reason per issue-active cycle; `selected` = issuing):\n")
print("| id | kernel | 1st | 2nd | 3rd | 4th |")
print("|---|---|---|---|---|---|")
for r in data:
    s = sorted(((num(r, h), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in STALL), reverse=True)
    s = [x for x in s if x[1] != "selected"][:4]
    print(f"| {r[col['ID']]} | {kname(r)} | " + " | ".join(f"{n} {v:.2f}" for v, n in s) + " |")

# dynamic opcode mix from the source page, first launch of each distinct kernel
if cells:
    print("\nDynamic opcode mix (source page; thread instructions per cell = warp-level executions x 32 / cells; first launch of each kernel):\n")
    done = set()
    for r in data:
        k = kname(r)
        if k in done or num(r, "smsp__inst_executed.sum") * 32 / cells < 20:
            continue
        done.add(k)
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{int(r[col['ID']]) + 1}"],
                             capture_output=True, text=True).stdout
        srows = list(csv.reader(io.StringIO(src)))
        try:
            h = next(x for x in srows if x and x[0] == "Address")
        except StopIteration:
            continue
        ix = {n: i for i, n in enumerate(h)}
        ex = collections.Counter()
        for x in srows:
            if len(x) < len(h) or not x[0].startswith("0x"):
                continue
            op = re.sub(r"^@!?U?P\w+\s+", "", x[ix["Source"]].strip()).split()[0].split(".")[0].rstrip(";")
            ex[op] += int(x[ix["Instructions Executed"]])
        te = sum(ex.values())
        if te == 0:
            continue
        print(f"* `{k}` ({te * 32 / cells:.0f} / cell): " + ", ".join(f"{op} {n * 32 / cells:.0f}" for op, n in ex.most_common(18)))
