#!/usr/bin/env python
"""Top stall sites of one kernel from `ncu -i rep --page source --csv` output.
usage: ncu_top.py src.csv [N] [context]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {k: hdr.index(k) for k in ("Source", "# Samples", "stall_long_sb", "stall_barrier", "stall_wait",
                                  "stall_short_sb", "stall_mio", "stall_math", "Instructions Executed")}
data = []
for n, r in enumerate(rows[hi + 1:]):
    if len(r) <= ix["# Samples"] or r[0] == "Address":
        continue
    try:
        s = int(r[ix["# Samples"]] or 0)
    except ValueError:
        continue
    data.append((s, n, r))
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
top = sorted(data, key=lambda d: -d[0])[:N]
for s, n, r in top:
    print(f"{s:6d} {s / tot * 100:5.1f}% i{n:5d} long={r[ix['stall_long_sb']]:>5s} bar={r[ix['stall_barrier']]:>5s} "
          f"wait={r[ix['stall_wait']]:>5s} ssb={r[ix['stall_short_sb']]:>5s} mio={r[ix['stall_mio']]:>4s} "
          f"math={r[ix['stall_math']]:>4s} ex={r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:80]}")
if ctx:
    for s, n, r in top[:3]:
        print("---- context around i%d" % n)
        for s2, n2, r2 in data[max(0, n - ctx): n + ctx]:
            print(f"   {s2:6d} i{n2:5d} {r2[ix['Source']].strip()[:100]}")
