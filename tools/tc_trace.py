#!/usr/bin/env python
"""Timeline of one CTA of the tcgen05 extend kernel (library built with B200_TC_TRACE=1 bash csrc/build.sh).
Runs config C4 once and prints per-KV-tile clock deltas of the producer / MMA issuer / softmax warpgroups."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from scratchpad_b200 import _lib  # noqa: E402
import bench  # noqa: E402

_lib.LIB_PATH = _lib.LIB_PATH.replace("libb200attn.so", "libb200attn_trace.so")
lib = _lib.load()
nrows = int(sys.argv[1]) if len(sys.argv) > 1 else 40
workload = sys.argv[2] if len(sys.argv) > 2 else "C4"
sys.argv = ["bench.py", "--workload", workload, "--steps", "3", "--warmup", "3", "--layers", "1", "--no-e2e", "--no-cpu-baseline",
            "--mode", "eager"]
bench.main()
torch.cuda.synchronize()
buf = np.zeros((20, 128), dtype=np.int64)
rc = lib.b200_debug_tc_trace_read(buf.ctypes.data_as(C.c_void_p))
assert rc == 0
names = ["prod:empty", "prod:issued", "mma:fullV", "mma:P0", "mma:fullK", "mma:P1", "sm0:S", "sm0:ld", "sm0:Pdone", "sm1:S",
         "sm1:ld", "sm1:Pdone", "sm0:max", "sm0:resc", "sm0:exp", "sm0:st"]
t0 = buf[buf > 0].min()
np.set_printoptions(linewidth=250)
print("kernel entry", buf[16, 0] - t0, "after griddepcontrol.wait", buf[16, 1] - t0, "tile begin (TMEM/barriers ready)", buf[17, 0] - t0,
      "Q loads issued", buf[17, 1] - t0, "first Q chunk converted", buf[17, 2] - t0, "Q stored", buf[17, 3] - t0,
      "Q loaded", buf[18, 0] - t0, "epilogue done", buf[19, 0] - t0)
print("clock offsets from the first event; rows = KV tile j (producer rows: n = 2j / 2j+1 shown as K,V)")
print("j   " + " ".join(f"{n:>11s}" for n in ["prodK:empty", "prodK:iss", "prodV:empty", "prodV:iss"] + names[2:]))
for j in range(nrows):
    row = [buf[0, 2 * j], buf[1, 2 * j], buf[0, 2 * j + 1], buf[1, 2 * j + 1]] if 2 * j + 1 < 128 else [0, 0, 0, 0]
    row += [buf[e, j] for e in range(2, 16)]
    print(f"{j:3d} " + " ".join(f"{(x - t0) if x > 0 else -1:11d}" for x in row))
