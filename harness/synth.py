"""Synthetic engine state for tests and bench: pools, page tables and batches
with the exact dtypes/strides the engine produces.

Nothing here is on the product path; it stands in for the scheduler / model
runner that own these objects in the real engine
(``scratchpad/model_executor/model_runner.py:360-442``,
``scratchpad/scheduler/schedule_batch.py:901-1071,1230-1308``).
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from types import SimpleNamespace
from typing import List, Optional, Sequence

import torch

from .forward_info import ForwardBatch, ForwardMode, compute_extend_start_loc
from scratchpad_b200.pool import MHATokenToKVPool
from .pools import ReqToTokenPool
from .radix_attention import RadixAttention


@dataclass
class Workload:
    """One BASELINE.json config (per-GPU view after KV-head sharding)."""
    name: str
    hq_total: int
    hkv_total: int
    head_dim: int
    batch: int
    kv_dtype: torch.dtype
    description: str
    ctx: int = 0                 # fixed context length (decode)
    shared_prefix: int = 0       # RadixAttention-style shared prefix length
    suffix_max: int = 0          # private suffix ~ U{1..suffix_max}
    ctx_min: int = 0             # mixed contexts: log-uniform in [ctx_min, ctx_max]
    ctx_max: int = 0
    extend_len: int = 0          # extend workloads
    prefix_len: int = 0
    layers_total: int = 1


WORKLOADS = {
    "C1": Workload("C1", 32, 8, 64, 8, torch.bfloat16,
                   "Llama-3.2-1B decode, batch 8, 1k context, bf16 paged KV (page=1)", ctx=1024, layers_total=16),
    "C2": Workload("C2", 32, 8, 64, 256, torch.bfloat16,
                   "Llama-3.2-1B batched decode, batch 256, 2k context, GQA 32q/8kv, bf16 KV", ctx=2048, layers_total=16),
    "C3": Workload("C3", 32, 8, 128, 256, torch.float8_e4m3fn,
                   "Llama-3.1-8B decode, shared 4k prefix, batch 256, FP8-e4m3 KV", shared_prefix=4096,
                   suffix_max=256, layers_total=32),
    "C3U": Workload("C3U", 32, 8, 128, 256, torch.float8_e4m3fn,
                    "Llama-3.1-8B decode, unshared 4k context control, batch 256, FP8-e4m3 KV", ctx=4096, layers_total=32),
    "C3B": Workload("C3B", 32, 8, 128, 256, torch.bfloat16,
                    "Llama-3.1-8B decode, batch 256, 4k context, bf16 KV (north_star target shape)", ctx=4096, layers_total=32),
    "C4": Workload("C4", 32, 8, 128, 8, torch.bfloat16,
                   "Llama-3.1-8B chunked extend: 8x2k new tokens on 8k cached prefix, causal, bf16",
                   extend_len=2048, prefix_len=8192, layers_total=32),
    "C4S": Workload("C4S", 32, 8, 64, 8, torch.bfloat16,
                    "Llama-3.2-1B chunked extend: 8x2k new tokens on 8k cached prefix, causal, bf16",
                    extend_len=2048, prefix_len=8192, layers_total=16),
    "C4SF": Workload("C4SF", 32, 8, 64, 8, torch.float8_e4m3fn,
                     "Llama-3.2-1B chunked extend: 8x2k new tokens on 8k cached prefix, causal, FP8-e4m3 KV pool",
                     extend_len=2048, prefix_len=8192, layers_total=16),
    "C4F": Workload("C4F", 32, 8, 128, 8, torch.float8_e4m3fn,
                    "Llama-3.1-8B chunked extend: 8x2k new tokens on 8k cached prefix, causal, FP8-e4m3 KV pool",
                    extend_len=2048, prefix_len=8192, layers_total=32),
    "C4P": Workload("C4P", 32, 8, 128, 8, torch.bfloat16,
                    "Llama-3.1-8B prefill without a cached prefix: 8x4k new tokens, causal, bf16",
                    extend_len=4096, prefix_len=0, layers_total=32),
    "C5": Workload("C5", 64, 8, 128, 512, torch.float8_e4m3fn,
                   "Llama-3.1-70B decode TP=8, batch 512, mixed 1k-32k contexts, FP8 KV", ctx_min=1024,
                   ctx_max=32768, layers_total=80),
}


def workload_seq_lens(w: Workload, gen: torch.Generator) -> List[int]:
    if w.ctx:
        return [w.ctx] * w.batch
    if w.shared_prefix:
        suf = torch.randint(1, w.suffix_max + 1, (w.batch,), generator=gen)
        return [w.shared_prefix + int(s) for s in suf]
    if w.ctx_max:
        u = torch.rand(w.batch, generator=gen)
        lens = torch.exp(math.log(w.ctx_min) + u * (math.log(w.ctx_max) - math.log(w.ctx_min)))
        return [int(x) for x in lens.round().clamp(w.ctx_min, w.ctx_max)]
    raise ValueError("not a decode workload")


def make_model_runner(hq_total: int, hkv_total: int, head_dim: int, tp_size: int = 1, layer_num: int = 1,
                      pool_tokens: int = 4096, max_reqs: int = 64, context_len: int = 4096,
                      kv_dtype: torch.dtype = torch.bfloat16, device: str = "cuda", fused_kv: bool = False,
                      page_size: int = 1, dtype: torch.dtype = None):
    """The attributes AttentionBackend.__init__ may read over freshly allocated pools.  ``dtype`` = the
    model dtype (default: the pool's when that is a 16-bit type -- kv_cache_dtype "auto" -- else bfloat16)."""
    if dtype is None:
        dtype = kv_dtype if kv_dtype in (torch.bfloat16, torch.float16) else torch.bfloat16
    hkv_local = max(1, hkv_total // tp_size)  # ModelConfig.get_num_kv_heads (model_config.py:159-166)
    model_config = SimpleNamespace(
        num_attention_heads=hq_total, head_dim=head_dim, context_len=context_len,
        get_num_kv_heads=lambda tp: max(1, hkv_total // tp),
    )
    r2t = ReqToTokenPool(max_reqs + 1, context_len + 4, device)
    kv = MHATokenToKVPool(pool_tokens, page_size, kv_dtype, hkv_local, head_dim, layer_num, device, fused_kv=fused_kv)
    return SimpleNamespace(
        model_config=model_config, tp_size=tp_size, device=device, dtype=dtype,
        kv_cache_dtype="auto" if kv_dtype == dtype else "fp8_e4m3",
        server_args=SimpleNamespace(enable_dp_attention=False, attention_backend="cuda_b200"),
        req_to_token_pool=r2t, token_to_kv_pool=kv, sliding_window_size=None,
    )


def make_layer(runner, layer_id: int = 0, logit_cap: float = 0.0) -> RadixAttention:
    cfg = runner.model_config
    hq = cfg.num_attention_heads // runner.tp_size
    hkv = cfg.get_num_kv_heads(runner.tp_size)
    return RadixAttention(hq, cfg.head_dim, cfg.head_dim ** -0.5, hkv, layer_id, logit_cap=logit_cap)


def build_paged_table(req_to_token: torch.Tensor, seq_lens: Sequence[int], pool_tokens: int, page_size: int,
                      gen: torch.Generator, shared_prefix: int = 0, stale: bool = True):
    """``build_page_table`` for ``page_size > 1``: every request owns randomly chosen pages (ids >= 1; page 0 holds
    the dummy slots), position t sits at slot ``page(t // page_size) * page_size + t % page_size``.  A shared prefix
    shares whole pages only (paged radix matching is page-aligned, radix_cache.py:57-66): ``shared_prefix`` is
    rounded down to a page multiple.  ``pool_tokens`` is the pool's ``size``; the buffers have size + page_size rows."""
    rows, cols = req_to_token.shape
    b, ps = len(seq_lens), page_size
    assert b <= rows and max(seq_lens, default=0) <= cols
    if stale:
        req_to_token.copy_(torch.randint(0, pool_tokens + 1, (rows, cols), generator=gen, dtype=torch.int32))
    req_rows = torch.randperm(rows, generator=gen)[:b]
    shared_pages = shared_prefix // ps
    need = shared_pages + sum((n + ps - 1) // ps - (shared_pages if n >= shared_pages * ps and shared_pages else 0)
                              for n in seq_lens)
    num_pages = pool_tokens // ps
    assert need <= num_pages, f"pool too small: need {need} pages, have {num_pages}"
    pages = (torch.randperm(num_pages, generator=gen)[:need] + 1).to(torch.int32)
    offs = torch.arange(ps, dtype=torch.int32)
    shared = (pages[:shared_pages, None] * ps + offs).reshape(-1)
    cur = shared_pages
    for i, n in enumerate(seq_lens):
        r = req_rows[i]
        npg = (n + ps - 1) // ps
        if shared_pages and n >= shared_pages * ps:
            own = pages[cur:cur + npg - shared_pages]
            cur += npg - shared_pages
            ids = torch.cat([shared, (own[:, None] * ps + offs).reshape(-1)])
        else:
            own = pages[cur:cur + npg]
            cur += npg
            ids = (own[:, None] * ps + offs).reshape(-1)
        req_to_token[r, :n] = ids[:n]
    return req_rows.to(torch.int64)


def build_page_table(req_to_token: torch.Tensor, seq_lens: Sequence[int], pool_tokens: int, gen: torch.Generator,
                     mode: str = "random", shared_prefix: int = 0, stale: bool = True):
    """Fill ``req_to_token`` (CPU int32) like the scheduler would.  Returns ``req_pool_indices`` (int64).

    mode "random": slot ids are a random permutation of 1..P (worst case for the gather);
    mode "contiguous": each request owns a consecutive run (what a fresh allocator yields,
    pool.py:251-253).  ``shared_prefix`` > 0 makes the first positions of every row equal ids
    (RadixAttention prefix sharing, schedule_batch.py:938-941).  Slot 0 is never handed out."""
    rows, cols = req_to_token.shape
    b = len(seq_lens)
    assert b <= rows and max(seq_lens, default=0) <= cols
    if stale:
        req_to_token.copy_(torch.randint(0, pool_tokens + 1, (rows, cols), generator=gen, dtype=torch.int32))
    req_rows = torch.randperm(rows, generator=gen)[:b]
    need = shared_prefix + sum(max(0, n - shared_prefix) if shared_prefix and n >= shared_prefix else n for n in seq_lens)
    assert need <= pool_tokens, f"pool too small: need {need}, have {pool_tokens}"
    if mode == "random":
        ids = (torch.randperm(pool_tokens, generator=gen)[:need] + 1).to(torch.int32)
    elif mode == "contiguous":
        ids = torch.arange(1, need + 1, dtype=torch.int32)
    else:
        raise ValueError(mode)
    shared = ids[:shared_prefix]
    cur = shared_prefix
    for i, n in enumerate(seq_lens):
        r = req_rows[i]
        if shared_prefix and n >= shared_prefix:
            req_to_token[r, :shared_prefix] = shared
            req_to_token[r, shared_prefix:n] = ids[cur:cur + n - shared_prefix]
            cur += n - shared_prefix
        else:
            req_to_token[r, :n] = ids[cur:cur + n]
            cur += n
    return req_rows.to(torch.int64)


def fill_pool_random(pool: MHATokenToKVPool, gen_seed: int, per_head_scale: bool = True, on_device: bool = True):
    """Random N(0,1) KV in every layer buffer; fp8 pools are quantised with scale = amax/448."""
    dev = pool.device if on_device else "cpu"
    g = torch.Generator(device=dev).manual_seed(gen_seed)
    for layer_id in range(pool.layer_num):
        for which, buf in (("k", pool.k_buffer[layer_id]), ("v", pool.v_buffer[layer_id])):
            x = torch.randn(buf.shape, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
            if pool.dtype == torch.bfloat16:
                buf.copy_(x)
            else:
                xf = x.float()
                amax = xf.abs().amax(dim=(0, 2)) if per_head_scale else xf.abs().amax().expand(pool.head_num)
                scale = (amax / 448.0).clamp_min(1e-8)
                qx = (xf / scale.view(1, -1, 1)).to(torch.bfloat16).to(pool.dtype).view(torch.uint8)
                buf.copy_(qx)
                (pool.k_scale if which == "k" else pool.v_scale)[layer_id].copy_(scale)
            buf[0].zero_()


def make_decode_batch(runner, seq_lens: Sequence[int], gen: torch.Generator, mode: str = "random",
                      shared_prefix: int = 0, idx_dtype: torch.dtype = torch.int64, backend=None) -> ForwardBatch:
    r2t_pool = runner.req_to_token_pool
    table = torch.zeros(r2t_pool.req_to_token.shape, dtype=torch.int32)
    ps = runner.token_to_kv_pool.page_size
    if ps > 1:
        rows = build_paged_table(table, seq_lens, runner.token_to_kv_pool.size, ps, gen, shared_prefix)
    else:
        rows = build_page_table(table, seq_lens, runner.token_to_kv_pool.size, gen, mode, shared_prefix)
    r2t_pool.req_to_token.copy_(table)
    dev = runner.device
    sl = torch.tensor(list(seq_lens), dtype=idx_dtype)
    out_loc = torch.stack([table[rows[i], n - 1] if n > 0 else torch.tensor(0, dtype=torch.int32)
                           for i, n in enumerate(seq_lens)]).to(torch.int64)
    return ForwardBatch(
        forward_mode=ForwardMode.DECODE, batch_size=len(seq_lens),
        req_pool_indices=rows.to(idx_dtype).to(dev), seq_lens=sl.to(dev), out_cache_loc=out_loc.to(dev),
        seq_lens_sum=int(sum(seq_lens)), req_to_token_pool=r2t_pool,
        token_to_kv_pool=runner.token_to_kv_pool, attn_backend=backend,
    )


def make_extend_batch(runner, prefix_lens: Sequence[int], extend_lens: Sequence[int], gen: torch.Generator,
                      mode: str = "random", idx_dtype: torch.dtype = torch.int64, backend=None) -> ForwardBatch:
    seq_lens = [p + e for p, e in zip(prefix_lens, extend_lens)]
    r2t_pool = runner.req_to_token_pool
    table = torch.zeros(r2t_pool.req_to_token.shape, dtype=torch.int32)
    rows = build_page_table(table, seq_lens, runner.token_to_kv_pool.size, gen, mode)
    r2t_pool.req_to_token.copy_(table)
    dev = runner.device
    out_loc = torch.cat([table[rows[i], p:p + e] for i, (p, e) in enumerate(zip(prefix_lens, extend_lens))]).to(torch.int64)
    el = torch.tensor(list(extend_lens), dtype=torch.int32)
    return ForwardBatch(
        forward_mode=ForwardMode.EXTEND, batch_size=len(seq_lens),
        req_pool_indices=rows.to(idx_dtype).to(dev), seq_lens=torch.tensor(seq_lens, dtype=idx_dtype).to(dev),
        out_cache_loc=out_loc.to(dev), seq_lens_sum=int(sum(seq_lens)),
        extend_num_tokens=int(sum(extend_lens)), extend_seq_lens=el.to(dev),
        extend_prefix_lens=torch.tensor(list(prefix_lens), dtype=torch.int32).to(dev),
        extend_start_loc=compute_extend_start_loc(el).to(dev),
        extend_prefix_lens_cpu=list(prefix_lens), extend_seq_lens_cpu=list(extend_lens),
        req_to_token_pool=r2t_pool, token_to_kv_pool=runner.token_to_kv_pool, attn_backend=backend,
    )


# ---- roofline accounting: algorithmic bytes / flops of one launch ---------------------------------
def decode_algorithmic_bytes(seq_lens: Sequence[int], hq: int, hkv: int, d: int, dv: int,
                             kv_elt: int, q_elt: int = 2, idx_elt: int = 8, len_elt: int = 8) -> int:
    """K/V rows of every context token + Q read + O write + one int32 slot id per token + the batch index tensors.
    Split-KV partials and LSE are implementation traffic and excluded."""
    tot = int(sum(int(s) for s in seq_lens))
    b = len(seq_lens)
    return (tot * hkv * (d + dv) * kv_elt + b * hq * d * q_elt + b * hq * dv * q_elt
            + tot * 4 + b * (idx_elt + len_elt))


def extend_algorithmic_flops(prefix_lens: Sequence[int], extend_lens: Sequence[int], hq: int, d: int) -> int:
    """2 (Q.K^T) + 2 (P.V) flops per (query, visible key, head, dim); the causal half is counted exactly."""
    return 4 * d * hq * sum(int(e) * int(p) + int(e) * (int(e) + 1) // 2 for p, e in zip(prefix_lens, extend_lens))
