#!/usr/bin/env python
"""Per-CTA timeline of the split-KV decode kernel (library built with B200_DEC_TRACE=1 into libb200attn_trace.so).
usage: dec_trace.py [bench args...]   e.g. dec_trace.py --workload C2 --shard 8"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
os.environ["B200_ATTN_LIB"] = os.path.abspath("scratchpad_b200/csrc/libb200attn_trace.so")
from scratchpad_b200 import _lib  # noqa: E402
import bench  # noqa: E402

lib = _lib.load()
sys.argv = ["bench.py"] + sys.argv[1:] + ["--steps", "8", "--warmup", "4", "--layers", "1", "--no-e2e", "--no-cpu-baseline", "--mode", "eager"]
bench.main()
torch.cuda.synchronize()
buf = np.zeros((1024, 6), dtype=np.uint64)
assert lib.b200_debug_dec_trace_read(buf.ctypes.data_as(C.c_void_p)) == 0
b = buf[buf[:, 5] > 0].astype(np.int64)
t0 = b[:, 0].min()
b = b - t0
names = ["enter", "after pdl_wait", "prologue issued + Q", "first tile landed", "loop done", "exit"]
print(f"{len(b)} CTAs; ns from the first CTA's entry: min / median / max")
for i, n in enumerate(names):
    print(f"  {n:22s} {b[:, i].min():8d} {int(np.median(b[:, i])):8d} {b[:, i].max():8d}")
ends = np.sort(b[:, 4])
print("loop-done percentiles (ns): " + "  ".join(f"p{q}={int(np.percentile(ends, q))}" for q in (1, 10, 25, 50, 75, 90, 99, 100)))
print(f"mean loop-done {ends.mean():.0f} ns; if every CTA finished at the mean the kernel would end {ends.max() - ends.mean():.0f} ns earlier")
d = b[:, 1:] - b[:, :-1]
print("per-CTA phase lengths (ns): median / max")
for i, n in enumerate(["pdl_wait", "metadata + prologue issue + Q load", "wait for first tile", "streaming loop", "drain"]):
    print(f"  {n:36s} {int(np.median(d[:, i])):8d} {d[:, i].max():8d}")
