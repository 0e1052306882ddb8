"""Wait-time accounting of the CTA-pair extend kernel (cluster 0 of one launch): build the trace library with
    B200_OUT=$PWD/_ab/lib_tc2_trace.so B200_OBJDIR=$PWD/_ab/o4 B200_EXTRA_FLAGS=-DB200_TC2_TRACE bash scratchpad_b200/csrc/build.sh
then  B200_ATTN_LIB=$PWD/_ab/lib_tc2_trace.so python tools/tc2_trace.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gpu_util as G
from scratchpad_b200 import _lib

lib = _lib.load()
buf = (ctypes.c_longlong * 32)()
pre, ext = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, int(sys.argv[2]) if len(sys.argv) > 2 else 2048
c = G.random_extend_case(32, 8, 128, [pre] * 8, [ext] * 8, seed=1)
G.run_extend_case(c, force_path=_lib.PATH_TC2)           # warm-up (also fills L2 like the steady state)
lib.b200_debug_tc2_trace_read(buf, 1)
G.run_extend_case(c, force_path=_lib.PATH_TC2)
lib.b200_debug_tc2_trace_read(buf, 0)
names = ["producer warp 12: wait slot empty", "producer warp 12: cp.async.wait_group", "MMA: wait K tile full", "MMA: wait V tile full",
         "MMA: wait P (both halves)", "softmax warp 0: wait S ready", "softmax warp 0: pair barrier", "softmax loop total (thread 0)", "tiles"]
for rank in range(2):
    v = list(buf)[16 * rank:16 * rank + 16]
    nt = max(v[8], 1)
    print(f"cluster 0, CTA rank {rank}: {nt} tiles, softmax loop {v[7]} clk = {v[7] / nt:.0f} clk per tile")
    for i, n in enumerate(names[:7]):
        print(f"   {n:40s} {v[i]:>10d} clk  = {v[i] / nt:7.0f} per tile")
