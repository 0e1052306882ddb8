#!/usr/bin/env python
"""SASS opcode histogram of the in-tree library -> profiles/r02_sass_opcodes.txt (no GPU needed: cuobjdump -sass).
Shows which kernels really carry tcgen05 (UTCHMMA, LDTM/STTM, UTCBAR), TMA (UTMALDG), cp.async (LDGSTS) and
mma.sync (HMMA.16816) instructions."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "scratchpad_b200", "csrc", "libb200attn.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn, per = None, collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and fn:
        per[fn][m.group(1)] += 1


def demangle(n):
    return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]


KEY = re.compile(r"^(UTCHMMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UTMALDG|UTMAPF|LDGSTS|HMMA|MOVM|LDSM|MUFU\.EX2|SYNCS|UCGABAR|USETMAXREG|FENCE\.VIEW|F2FP|UBLKCP|ACQBULK)")
KERNELS = ("decode_mma_kernel", "decode_paged_kernel", "merge_kernel", "extend_tc", "extend_mma", "kv_store", "rope_kv",
           "plan_kernel", "group_kernel", "lcp_kernel")
tot, lines = collections.Counter(), []
for f, c in sorted(per.items(), key=lambda kv: demangle(kv[0])):
    name = demangle(f)
    if not any(k in name for k in KERNELS):
        continue
    sel = collections.Counter()
    for op, n in c.items():
        if not KEY.match(op):
            continue
        if op.startswith("HMMA"):  # HMMA.16816.F32 = f16 operands (fp8 pools, fp16 models), .F32.BF16 = bf16 operands
            key = ".".join(op.split(".")[:4])
        elif op.startswith(("UTCHMMA", "UTCBAR", "UTMALDG", "LDGSTS", "SYNCS", "F2FP")):
            key = ".".join(op.split(".")[:3])
        elif op.startswith("MUFU"):
            key = "MUFU.EX2"
        else:
            key = op.split(".")[0]
        sel[key] += n
        tot[key] += n
    lines.append(f"{name}\n    instructions {sum(c.values())}: " + ", ".join(f"{k} {v}" for k, v in sorted(sel.items())))
hdr = [
    "# SASS opcode histogram of scratchpad_b200/csrc/libb200attn.so (sm_100a), round 2 (final code: bf16 and fp16 model dtypes)",
    "# how: cuobjdump -sass libb200attn.so, opcodes counted per kernel; regenerate with `python tools/sass_hist.py`",
    "# what to look for: UTCHMMA = tcgen05.mma (UTCHMMA.2CTA = cta_group::2), LDTM/STTM = tcgen05.ld/st (TMEM), UTCBAR = tcgen05.commit,",
    "#   UTMALDG = cp.async.bulk.tensor (TMA), LDGSTS = cp.async 16-byte copies, HMMA.16816 = mma.sync m16n8k16, MOVM = movmatrix,",
    "#   LDSM = ldmatrix, SYNCS = mbarrier ops, UCGABAR = barrier.cluster, USETMAXREG = setmaxnreg",
    "", "## totals over the listed kernels", ", ".join(f"{k} {v}" for k, v in sorted(tot.items())), "", "## per kernel"]
dst = os.path.join(ROOT, "profiles", "r02_sass_opcodes.txt")
open(dst, "w").write("\n".join(hdr + lines) + "\n")
print(f"{dst}: {len(lines)} kernels")
