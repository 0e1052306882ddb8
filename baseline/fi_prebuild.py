#!/usr/bin/env python
"""Compile (no GPU needed) the FlashInfer JIT modules the reference's FlashInferAttnBackend ends up loading for the
benchmark and parity shapes, into baseline/_ref/fi_cache.  See baseline/fi_env.py.

    python baseline/fi_prebuild.py
"""
from __future__ import annotations

import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baseline import fi_env  # noqa: E402  (sets the JIT cache location before flashinfer is imported)

import torch  # noqa: E402


def main() -> int:
    fi_env.patch_jit()
    from flashinfer.cascade import get_cascade_module
    from flashinfer.decode import get_batch_decode_module
    from flashinfer.prefill import get_batch_prefill_module
    from flashinfer.utils import PosEncodingMode

    none = PosEncodingMode["NONE"].value
    bf16, fp8, i32 = torch.bfloat16, torch.float8_e4m3fn, torch.int32
    jobs = []
    for hd in (64, 128):
        for cap in (False, True):
            # BatchDecodeWithPagedKVCacheWrapper(use_tensor_cores=False): bf16 pools at group size <= 4
            # (flashinfer_backend.py:936-981)
            jobs.append(("decode", get_batch_decode_module, (bf16, bf16, bf16, i32, hd, hd, none, False, cap)))
            # tensor-core decode (fp8 pools, group size > 4) and both prefill wrappers run the fa2 prefill module
            for kv in (bf16, fp8):
                jobs.append(("prefill", get_batch_prefill_module,
                             ("fa2", bf16, kv, bf16, i32, hd, hd, none, False, cap, False)))
    # fp16 model dtype (fp16 or fp8 pool), no soft cap: the fp16 comparator cases of tests/test_fp16_gpu.py
    f16 = torch.float16
    for hd in (64, 128):
        jobs.append(("decode", get_batch_decode_module, (f16, f16, f16, i32, hd, hd, none, False, False)))
        for kv in (f16, fp8):
            jobs.append(("prefill", get_batch_prefill_module, ("fa2", f16, kv, f16, i32, hd, hd, none, False, False, False)))
    jobs.append(("cascade", get_cascade_module, ()))
    # the trtllm-gen FMHA launcher (its kernels are prebuilt cubins from flashinfer-cubin): compile only -- loading the
    # shared object needs libcuda, which the build container does not have
    from flashinfer.decode import gen_trtllm_gen_fmha_module
    jobs.append(("trtllm-gen fmha", lambda: gen_trtllm_gen_fmha_module().build(False), ()))
    t0 = time.time()
    for name, fn, args in jobs:
        t = time.time()
        fn(*args)
        print(f"fi_prebuild: {name} {[str(a) for a in args]} {time.time() - t:.1f}s", flush=True)
    # object files are not needed to load the modules; keep the shipped cache small
    for root, _, files in os.walk(fi_env.CACHE_BASE):
        for f in files:
            if f.endswith(".o"):
                os.remove(os.path.join(root, f))
    print(f"fi_prebuild: {len(jobs)} modules in {time.time() - t0:.0f}s -> {fi_env.CACHE_BASE}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
