#!/usr/bin/env python
"""Where the end-to-end decode step (bench.py's e2e leg, C2) spends its time: copies, metadata hook, graph replay, sync."""
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from harness import ForwardMode  # noqa: E402

sys.argv = ["bench.py", "--workload", sys.argv[1] if len(sys.argv) > 1 else "C2"]
args = bench.parse_args()
st = bench.build_state(args, 0, 1)
w, backend, fb, layers, L = st["w"], st["backend"], st["fb"], st["layers"], st["L"]
B, hq, hkv, d = w.batch, st["hq"], st["hkv"], st["d"]
dev = torch.device("cuda:0")
qkv = torch.randn((B, (hq + 2 * hkv) * d), device=dev).to(torch.bfloat16)
q, k, v = qkv[:, : hq * d], qkv[:, hq * d: (hq + hkv) * d], qkv[:, (hq + hkv) * d:]
backend.init_cuda_graph_state(B)
backend.init_forward_metadata_capture_cuda_graph(B, B, fb.req_pool_indices, fb.seq_lens, None, ForwardMode.DECODE, None)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    layers[0](q, k, v, fb, save_kv_cache=True)
torch.cuda.current_stream().wait_stream(side)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = layers[0](q, k, v, fb, save_kv_cache=True)
h_qkv, h_sl, h_rpi, h_loc = qkv.cpu().pin_memory(), fb.seq_lens.cpu().pin_memory(), fb.req_pool_indices.cpu().pin_memory(), fb.out_cache_loc.cpu().pin_memory()
h_o = torch.empty_like(out, device="cpu").pin_memory()
sync = torch.cuda.current_stream().synchronize


def h2d():
    qkv.copy_(h_qkv, non_blocking=True)
    fb.seq_lens.copy_(h_sl, non_blocking=True)
    fb.req_pool_indices.copy_(h_rpi, non_blocking=True)
    fb.out_cache_loc.copy_(h_loc, non_blocking=True)


meta_dev = torch.stack([fb.seq_lens.to(torch.int64), fb.req_pool_indices.to(torch.int64), fb.out_cache_loc.to(torch.int64)]).contiguous()
h_meta = meta_dev.cpu().pin_memory()


def h2d_packed():
    qkv.copy_(h_qkv, non_blocking=True)
    meta_dev.copy_(h_meta, non_blocking=True)


def hook():
    backend.init_forward_metadata_replay_cuda_graph(B, fb.req_pool_indices, fb.seq_lens, fb.seq_lens_sum, None, ForwardMode.DECODE, None, seq_lens_cpu=None)


def full():
    h2d(); hook(); g.replay(); h_o.copy_(out, non_blocking=True)


for name, fn in (("sync only", lambda: None), ("h2d", h2d), ("h2d packed meta", h2d_packed), ("h2d qkv only", lambda: qkv.copy_(h_qkv, non_blocking=True)), ("hook (plan)", hook), ("graph replay", g.replay),
                 ("d2h", lambda: h_o.copy_(out, non_blocking=True)), ("h2d+hook+graph", lambda: (h2d(), hook(), g.replay())), ("full step", full)):
    for _ in range(20):
        fn(); sync()
    t0 = time.perf_counter()
    n = 300
    for _ in range(n):
        fn(); sync()
    print(f"{name:18s} {(time.perf_counter() - t0) / n * 1e6:8.1f} us per step (including the stream sync)")
