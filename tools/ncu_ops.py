#!/usr/bin/env python
"""Opcode histogram (warp instructions executed, per unit of work) from an ncu source-page CSV.
usage: ncu_ops.py src.csv [units]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
per = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
si, ei, ni = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
ops, samp, tot, seen = collections.Counter(), collections.Counter(), 0, set()
for r in rows[hi + 1:]:
    if r[0] in seen:
        continue
    seen.add(r[0])
    try:
        e, s = int(r[ei]), int(r[ni])
    except ValueError:
        continue
    op = r[si].split()
    name = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
    ops[name] += e
    samp[name] += s
    tot += e
print("total warp instr", tot, "per unit", tot / per)
for k, v in ops.most_common(28):
    print(f"{k:10s} {v:12d} per-unit {v / per:8.1f}  samples {samp[k]}")
