#!/usr/bin/env python
"""Dynamic instruction mix + stall samples per opcode from `ncu -i rep --page source --csv --kernel-name regex:K`.
usage: ncu_source_hist.py source.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(hdr)}
STALLS = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot, execd, stall = collections.Counter(), collections.Counter(), collections.Counter()
samples = 0
for r in rows:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    op = re.sub(r"^@!?U?P\w+\s+", "", r[ix["Source"]].strip()).split()[0].split(".")[0].rstrip(";")
    n = int(r[ix["# Samples"]])
    samples += n
    tot[op] += n
    execd[op] += int(r[ix["Instructions Executed"]])
    for s in STALLS:
        stall[s] += int(r[ix[s]])
te = sum(execd.values())
print("samples", samples, "warp-inst executed", te)
print("stalls:", ", ".join(f"{k[6:]}={100 * v / samples:.1f}%" for k, v in stall.most_common(10)))
for op, n in tot.most_common(22):
    print(f"{op:10s} samples {100 * n / samples:5.1f}%   executed {100 * execd[op] / te:5.1f}%")
