#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total and share.
usage: summarize_launches.py launches.csv [skip_first_n_launches] > profiles/xxx.md"""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    rows.append((int(r["ID"]), re.sub(r"\(.*", "", r["Kernel Name"]), r["Grid Size"], r["Block Size"], float(r["Metric Value"]) / 1e6))
rows = [r for r in rows if r[0] >= skip]
agg = OrderedDict()
for _, name, grid, block, ms in rows:
    a = agg.setdefault(name, [0, 0.0, grid, block])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
print(f"source: {path} (launch ids >= {skip}); ncu per-launch times are cold-cache and serialised: compare SHARES\n")
print("| kernel | launches | total ms | avg ms | share | grid | block |")
print("|---|---|---|---|---|---|---|")
for name, (n, ms, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {name} | {n} | {ms:.3f} | {ms / n:.3f} | {100 * ms / tot:.1f}% | {grid} | {block} |")
print(f"| total | {sum(a[0] for a in agg.values())} | {tot:.3f} | | | | |")
