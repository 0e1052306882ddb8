#!/usr/bin/env python
"""bf16 vs fp16 model dtype at the headline decode shape (C3B: Llama-3.1-8B heads, batch 256, 4k context; scattered and
consecutive slots) and at the C4 extend shape (8 x 2048 new tokens on 8192 cached), same process, eager launches through
the AttentionBackend surface, CUDA events.  The fp16 kernels are the bf16 ones with f16 MMAs, so the times should agree.
  gpurun -- 'python tools/dtype_timing.py > gpurun_out/dtype_timing.txt'"""
import sys

import torch

sys.path.insert(0, ".")
from harness import synth  # noqa: E402
from scratchpad_b200 import CudaB200AttnBackend  # noqa: E402


def time_cuda(fn, iters, warmup=4):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters  # us


def fill(pool):
    for buf in pool.k_buffer + pool.v_buffer:
        buf.normal_()
        buf[0].zero_()


def decode(dtype, mode, layers=3):
    hq, hkv, d, bsz, ctx = 32, 8, 128, 256, 4096
    runner = synth.make_model_runner(hq, hkv, d, layer_num=layers, pool_tokens=bsz * ctx, max_reqs=bsz, context_len=ctx,
                                     kv_dtype=dtype, dtype=dtype)
    fill(runner.token_to_kv_pool)
    backend = CudaB200AttnBackend(runner)
    fb = synth.make_decode_batch(runner, [ctx] * bsz, torch.Generator().manual_seed(1), mode=mode, backend=backend)
    lyr = [synth.make_layer(runner, i) for i in range(layers)]
    q = torch.randn((bsz, hq * d), device="cuda").to(dtype)
    backend.init_forward_metadata(fb)
    us = time_cuda(lambda i: lyr[i % layers](q, None, None, fb, save_kv_cache=False), 30)
    nbytes = synth.decode_algorithmic_bytes([ctx] * bsz, hq, hkv, d, d, 2)
    return us, nbytes / us / 1e3  # us, GB/s


def extend(dtype, layers=2):
    hq, hkv, d, bsz, pre, ext = 32, 8, 128, 8, 8192, 2048
    runner = synth.make_model_runner(hq, hkv, d, layer_num=layers, pool_tokens=bsz * (pre + ext), max_reqs=bsz,
                                     context_len=pre + ext, kv_dtype=dtype, dtype=dtype)
    fill(runner.token_to_kv_pool)
    backend = CudaB200AttnBackend(runner)
    fb = synth.make_extend_batch(runner, [pre] * bsz, [ext] * bsz, torch.Generator().manual_seed(2), backend=backend)
    lyr = [synth.make_layer(runner, i) for i in range(layers)]
    t = bsz * ext
    q, k, v = (torch.randn((t, n * d), device="cuda").to(dtype) for n in (hq, hkv, hkv))
    backend.init_forward_metadata(fb)
    us = time_cuda(lambda i: lyr[i % layers](q, k, v, fb, save_kv_cache=False), 8, warmup=2)
    return us, synth.extend_algorithmic_flops([pre] * bsz, [ext] * bsz, hq, d) / us / 1e6  # us, TFLOP/s


for mode in ("random", "contiguous"):
    for dt in (torch.bfloat16, torch.float16):
        us, gbs = decode(dt, mode)
        print(f"decode C3B shape, {mode:10s} slots, {str(dt):15s}: {us:7.1f} us per layer-call (attention only), {gbs:6.0f} GB/s = {gbs / 80:.1f} % of 8 TB/s")
        torch.cuda.empty_cache()
for dt in (torch.bfloat16, torch.float16):
    us, tf = extend(dt)
    print(f"extend C4 shape, random slots, {str(dt):15s}: {us:7.0f} us per layer-call, {tf:5.0f} TFLOP/s")
    torch.cuda.empty_cache()
