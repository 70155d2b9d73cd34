#!/usr/bin/env python
"""Per-kernel summary of an ncu --set full report (read on the CPU box).
usage: ncu_summary.py report.ncu-rep > profiles/xxx.md"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: k for k, n in enumerate(hdr)}
WANT = [
    ("gpu__time_duration.sum", "ms", 1e-6),
    ("dram__bytes_read.sum", "rd MB", None),
    ("dram__bytes_write.sum", "wr MB", None),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %", 1),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("smsp__inst_executed.sum", "warp inst M", 1e-6),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %", 1),
    ("lts__t_sector_hit_rate.pct", "L2 hit %", 1),
]


def tobytes(v, unit):
    v = float(v.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


names = [w[1] for w in WANT]
print(f"source: {rep} (ncu --set full, per launch; times are under replay, cold cache)\n")
print("| id | kernel | " + " | ".join(names) + " |")
print("|---|---|" + "---|" * len(names))
for r in data:
    out = []
    for m, label, scale in WANT:
        if m not in col:
            out.append("-")
            continue
        v, u = r[col[m]], units[col[m]]
        try:
            if scale is None:
                out.append(f"{tobytes(v, u) / 1e6:.1f}")
            else:
                x = float(v.replace(",", "")) * scale
                if m == "gpu__time_duration.sum":
                    x = float(v.replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
                out.append(f"{x:.3f}" if x < 100 else f"{x:.1f}")
        except ValueError:
            out.append(v)
    kname = re.sub(r"\(.*", "", r[col["Kernel Name"]])
    print(f"| {r[col['ID']]} | {kname} | " + " | ".join(out) + " |")
