"""CPU oracle for the paged decode / extend attention hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``scratchpad_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and there only as the
checker or as the reported CPU baseline -- never as the product path.

What it restates (all paths relative to /root/reference):

* the backend contract -- KV write first, then attend, output ``[T, Hq*Dv]``:
  ``scratchpad/nn/attention/triton_backend.py:118-196``
* the pool layout and the fp8 store semantics (``x / scale`` -> cast -> uint8
  view, slot ids index dim 0 of ``[P+1, Hkv, D]``):
  ``scratchpad/memory/pool.py:298-317,373-424``
* decode math -- ``softmax(q . K^T * sm_scale [softcap]) . V`` over the slots
  ``req_to_token[req_pool_indices[i], :seq_lens[i]]``:
  ``scratchpad/nn/attention/triton_attn/decode_attention.py:251-434``
* extend math -- query row t of request i attends, non-causally, to the
  prefix slots ``req_to_token[r, :pre_i]`` read from the pool and, causally,
  to the new tokens ``k_new[start_i : start_i + t + 1]`` read from the call
  arguments (never re-read from the pool):
  ``scratchpad/nn/attention/triton_attn/extend_attention.py:53-226``
* ``extend_start_loc`` = exclusive cumsum of ``extend_seq_lens``:
  ``scratchpad/model_executor/forward_info.py:452-466``

Parity pin: the reference has no golden vectors for this path (SURVEY.md
section 8c).  The oracle is pinned against the reference's *own* in-tree
kernels executed in this container under ``TRITON_INTERPRET=1``
(``oracle/gen_golden.py`` -> ``tests/golden/*.npz``; checked by
``tests/test_oracle_golden.py``).  The fp8-e4m3 KV variant has no working
Triton path in the reference (its e5m2 path is broken, SURVEY.md section 0
fact 4); its semantics here follow ``MHATokenToKVPool.set_kv_buffer`` and the
scales the reference's FlashInfer backend passes
(flashinfer_backend.py:387-388,457-458,475-482), and are pinned on the GPU box
against that backend itself -- the unmodified ``FlashInferAttnBackend`` from
``baseline/_ref`` over the installed FlashInfer, on the same quantised pool
bytes (``tests/test_comparators_gpu.py``: oracle, reference and CUDA path
agree pairwise within the fp8 tolerance).

All math is fp32 (the tolerance in BASELINE.json is stated against an fp32
reference).  ``decode_torch_native`` / ``extend_torch_native`` are the
bf16-storage flavour (per-request gather + ``scaled_dot_product_attention``)
that ``bench.py`` times as the CPU baseline.
"""

from __future__ import annotations

import math
from typing import Optional, Sequence, Union

import torch

ScaleT = Union[None, float, torch.Tensor]

FP8_DTYPES = (torch.float8_e4m3fn, torch.float8_e5m2)


# --------------------------------------------------------------------------
# pool semantics (pool.py:373-424)
# --------------------------------------------------------------------------
def _per_head(scale: ScaleT, hkv: int) -> Optional[torch.Tensor]:
    """Scale -> fp32 ``[Hkv]`` vector (per-tensor scalars are broadcast)."""
    if scale is None:
        return None
    if isinstance(scale, torch.Tensor):
        s = scale.detach().to(torch.float32).reshape(-1).cpu()
        if s.numel() == 1:
            s = s.expand(hkv)
        assert s.numel() == hkv, "scale must be a scalar or one value per kv head"
        return s.contiguous()
    return torch.full((hkv,), float(scale), dtype=torch.float32)


def quantize_kv_ref(x: torch.Tensor, kv_dtype: torch.dtype, scale: ScaleT) -> torch.Tensor:
    """``x [T, Hkv, D]`` (model dtype) -> pool store representation.

    pool.py:404-412: ``cache_k.div_(k_scale)`` happens *in the model dtype*
    (so the quotient is rounded to bf16 before the fp8 cast), then ``.to(dtype)``
    and, for fp8, ``.view(uint8)``.
    """
    if x.dtype == kv_dtype:
        return x.clone()
    y = x.clone()
    s = _per_head(scale, x.shape[1])
    if s is not None:
        # fp32 division rounded back to the model dtype == Tensor.div_ on CPU
        y = (y.to(torch.float32) / s.view(1, -1, 1)).to(x.dtype)
    y = y.to(kv_dtype)
    if kv_dtype in FP8_DTYPES:
        y = y.view(torch.uint8)
    return y


def set_kv_buffer_ref(
    k_buf: torch.Tensor,
    v_buf: torch.Tensor,
    loc: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    kv_dtype: torch.dtype,
    k_scale: ScaleT = None,
    v_scale: ScaleT = None,
) -> None:
    """``k_buf[loc] = quant(k)`` (pool.py:392-424); later rows win on duplicate
    ``loc`` entries exactly as ``index_put_`` without accumulate does on CPU."""
    k_buf[loc.long()] = quantize_kv_ref(k, kv_dtype, k_scale).view(k_buf.dtype)
    v_buf[loc.long()] = quantize_kv_ref(v, kv_dtype, v_scale).view(v_buf.dtype)


def dequantize_kv_ref(rows: torch.Tensor, kv_dtype: torch.dtype, scale: ScaleT) -> torch.Tensor:
    """Pool rows ``[S, Hkv, D]`` (store dtype) -> fp32 (pool.py:377-379 view +
    the ``* scale`` the FlashInfer path applies, flashinfer_backend.py:475-482)."""
    if kv_dtype in FP8_DTYPES and rows.dtype == torch.uint8:
        rows = rows.view(kv_dtype)
    x = rows.to(torch.float32)
    s = _per_head(scale, rows.shape[1])
    if s is not None:
        x = x * s.view(1, -1, 1)
    return x


def gather_rows_ref(
    buf: torch.Tensor, req_to_token: torch.Tensor, req_pool_index: int, length: int
) -> torch.Tensor:
    """The bit-exact gather contract: rows ``buf[req_to_token[r, :length]]``."""
    idx = req_to_token[int(req_pool_index), : int(length)].long()
    return buf[idx]


# --------------------------------------------------------------------------
# math helpers
# --------------------------------------------------------------------------
def _softcap(s: torch.Tensor, logit_cap: Optional[float]) -> torch.Tensor:
    # decode_attention.py:343-344 / extend_attention.py:10-13,138-139
    if logit_cap is not None and logit_cap > 0:
        return logit_cap * torch.tanh(s / logit_cap)
    return s


def _attend(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, sm_scale: float,
            logit_cap: Optional[float], mask: Optional[torch.Tensor]) -> torch.Tensor:
    """q [Tq,Hq,D], k [S,Hkv,D], v [S,Hkv,Dv] fp32 -> [Tq,Hq,Dv] fp32.
    kv heads are shared by ``g = Hq // Hkv`` consecutive q heads
    (decode_attention.py:275, extend_attention.py:55)."""
    tq, hq, d = q.shape
    s_len, hkv, dv = v.shape
    if s_len == 0:
        return torch.zeros((tq, hq, dv), dtype=torch.float32)
    g = hq // hkv
    qg = q.reshape(tq, hkv, g, d)
    # scores [Hkv, g, Tq, S]
    s = torch.einsum("thgd,shd->hgts", qg, k) * sm_scale
    s = _softcap(s, logit_cap)
    if mask is not None:
        s = s.masked_fill(~mask.view(1, 1, tq, s_len), float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = torch.einsum("hgts,shd->thgd", p, v)
    return o.reshape(tq, hq, dv)


# --------------------------------------------------------------------------
# decode (triton_backend.py:157-196)
# --------------------------------------------------------------------------
def decode_ref(
    q: torch.Tensor,                 # [B, Hq, D] model dtype
    k_buf: torch.Tensor,             # [P+1, Hkv, D] store dtype
    v_buf: torch.Tensor,             # [P+1, Hkv, Dv]
    req_to_token: torch.Tensor,      # [R, C] int32
    req_pool_indices: torch.Tensor,  # [B] int32/int64
    seq_lens: torch.Tensor,          # [B] int32/int64
    sm_scale: float,
    logit_cap: Optional[float] = 0.0,
    kv_dtype: Optional[torch.dtype] = None,
    k_scale: ScaleT = None,
    v_scale: ScaleT = None,
) -> torch.Tensor:
    """fp32 oracle of one layer's decode attention. Returns ``[B, Hq, Dv]`` fp32.
    A request with ``seq_len == 0`` yields zeros."""
    kv_dtype = kv_dtype or k_buf.dtype
    b, hq, _ = q.shape
    dv = v_buf.shape[-1]
    out = torch.zeros((b, hq, dv), dtype=torch.float32)
    qf = q.to(torch.float32)
    for i in range(b):
        n = int(seq_lens[i])
        if n == 0:
            continue
        r = int(req_pool_indices[i])
        k = dequantize_kv_ref(gather_rows_ref(k_buf, req_to_token, r, n), kv_dtype, k_scale)
        v = dequantize_kv_ref(gather_rows_ref(v_buf, req_to_token, r, n), kv_dtype, v_scale)
        out[i] = _attend(qf[i : i + 1], k, v, sm_scale, logit_cap, None)[0]
    return out


# --------------------------------------------------------------------------
# extend (triton_backend.py:118-155, extend_attention.py:53-226)
# --------------------------------------------------------------------------
def extend_start_loc_ref(extend_seq_lens: torch.Tensor) -> torch.Tensor:
    """forward_info.py:452-466 (compute_position_torch)."""
    out = torch.zeros_like(extend_seq_lens)
    out[1:] = torch.cumsum(extend_seq_lens[:-1], dim=0)
    return out


def extend_ref(
    q: torch.Tensor,                 # [T, Hq, D]
    k_new: torch.Tensor,             # [T, Hkv, D]  model dtype, contiguous new tokens
    v_new: torch.Tensor,             # [T, Hkv, Dv]
    k_buf: torch.Tensor,
    v_buf: torch.Tensor,
    req_to_token: torch.Tensor,
    req_pool_indices: torch.Tensor,
    seq_lens: torch.Tensor,          # prefix + extend, per request
    extend_seq_lens: torch.Tensor,
    extend_start_loc: Optional[torch.Tensor],
    sm_scale: float,
    logit_cap: Optional[float] = 0.0,
    kv_dtype: Optional[torch.dtype] = None,
    k_scale: ScaleT = None,
    v_scale: ScaleT = None,
) -> torch.Tensor:
    """fp32 oracle of one layer's extend attention. Returns ``[T, Hq, Dv]`` fp32.

    Prefix K/V come from the pool (dequantised when the pool is fp8); the new
    tokens' K/V come from ``k_new``/``v_new`` unquantised, exactly as both
    reference backends do (extend_attention.py:159-216;
    flashinfer_backend.py:419-439)."""
    kv_dtype = kv_dtype or k_buf.dtype
    t, hq, _ = q.shape
    dv = v_buf.shape[-1]
    if extend_start_loc is None:
        extend_start_loc = extend_start_loc_ref(extend_seq_lens)
    out = torch.zeros((t, hq, dv), dtype=torch.float32)
    qf, kf, vf = q.to(torch.float32), k_new.to(torch.float32), v_new.to(torch.float32)
    for i in range(len(seq_lens)):
        ext = int(extend_seq_lens[i])
        if ext == 0:
            continue
        pre = int(seq_lens[i]) - ext
        st = int(extend_start_loc[i])
        r = int(req_pool_indices[i])
        kp = dequantize_kv_ref(gather_rows_ref(k_buf, req_to_token, r, pre), kv_dtype, k_scale)
        vp = dequantize_kv_ref(gather_rows_ref(v_buf, req_to_token, r, pre), kv_dtype, v_scale)
        k = torch.cat([kp, kf[st : st + ext]], dim=0)
        v = torch.cat([vp, vf[st : st + ext]], dim=0)
        # row t sees prefix (all) + new tokens 0..t  (extend_attention.py:195-199)
        cols = torch.arange(pre + ext).view(1, -1)
        rows = torch.arange(ext).view(-1, 1)
        mask = cols <= (rows + pre)
        out[st : st + ext] = _attend(qf[st : st + ext], k, v, sm_scale, logit_cap, mask)
    return out


def redundant_attention_ref(q, k_buf, v_buf, req_to_token, req_pool_indices, seq_lens,
                            extend_seq_lens, extend_start_loc, sm_scale):
    """SGLang's cross-check (extend_attention.py:330-364): full causal attention
    over prefix+extend read *from the pool* (which must already hold the new
    tokens), sliced to the extend rows. Used to test write-then-attend."""
    t, hq, _ = q.shape
    dv = v_buf.shape[-1]
    out = torch.zeros((t, hq, dv), dtype=torch.float32)
    qf = q.to(torch.float32)
    for i in range(len(seq_lens)):
        ext, n = int(extend_seq_lens[i]), int(seq_lens[i])
        if ext == 0:
            continue
        pre, st, r = n - ext, int(extend_start_loc[i]), int(req_pool_indices[i])
        k = gather_rows_ref(k_buf, req_to_token, r, n).to(torch.float32)
        v = gather_rows_ref(v_buf, req_to_token, r, n).to(torch.float32)
        mask = torch.arange(n).view(1, -1) <= (torch.arange(ext).view(-1, 1) + pre)
        out[st : st + ext] = _attend(qf[st : st + ext], k, v, sm_scale, 0.0, mask)
    return out


# --------------------------------------------------------------------------
# the timed CPU baseline: bf16 storage, per-request gather + SDPA
# (the "torch-native backend" shape named in BASELINE.md section 3)
# --------------------------------------------------------------------------
def decode_torch_native(q, k_buf, v_buf, req_to_token, req_pool_indices, seq_lens,
                        sm_scale, math_dtype=torch.float32, kv_dtype=None,
                        k_scale: ScaleT = None, v_scale: ScaleT = None) -> torch.Tensor:
    b, hq, d = q.shape
    hkv, dv = v_buf.shape[1], v_buf.shape[2]
    out = torch.empty((b, hq, dv), dtype=q.dtype)
    kv_dtype = kv_dtype or k_buf.dtype
    for i in range(b):
        n = int(seq_lens[i])
        idx = req_to_token[int(req_pool_indices[i]), :n].long()
        if kv_dtype in FP8_DTYPES:
            k = dequantize_kv_ref(k_buf[idx], kv_dtype, k_scale).to(math_dtype)
            v = dequantize_kv_ref(v_buf[idx], kv_dtype, v_scale).to(math_dtype)
        else:
            k, v = k_buf[idx].to(math_dtype), v_buf[idx].to(math_dtype)
        o = torch.nn.functional.scaled_dot_product_attention(
            q[i].to(math_dtype).view(1, hq, 1, d),
            k.permute(1, 0, 2).unsqueeze(0),
            v.permute(1, 0, 2).unsqueeze(0),
            scale=sm_scale, enable_gqa=(hq != hkv),
        )
        out[i] = o.view(hq, dv).to(q.dtype)
    return out


def extend_torch_native(q, k_new, v_new, k_buf, v_buf, req_to_token, req_pool_indices,
                        seq_lens, extend_seq_lens, extend_start_loc, sm_scale,
                        math_dtype=torch.float32) -> torch.Tensor:
    t, hq, d = q.shape
    hkv, dv = v_buf.shape[1], v_buf.shape[2]
    out = torch.empty((t, hq, dv), dtype=q.dtype)
    for i in range(len(seq_lens)):
        ext = int(extend_seq_lens[i])
        pre, st = int(seq_lens[i]) - ext, int(extend_start_loc[i])
        idx = req_to_token[int(req_pool_indices[i]), :pre].long()
        k = torch.cat([k_buf[idx], k_new[st : st + ext]], 0).to(math_dtype)
        v = torch.cat([v_buf[idx], v_new[st : st + ext]], 0).to(math_dtype)
        mask = torch.arange(pre + ext).view(1, -1) <= (torch.arange(ext).view(-1, 1) + pre)
        o = torch.nn.functional.scaled_dot_product_attention(
            q[st : st + ext].to(math_dtype).permute(1, 0, 2).unsqueeze(0),
            k.permute(1, 0, 2).unsqueeze(0), v.permute(1, 0, 2).unsqueeze(0),
            attn_mask=mask, scale=sm_scale, enable_gqa=(hq != hkv),
        )
        out[st : st + ext] = o[0].permute(1, 0, 2).to(q.dtype)
    return out


# --------------------------------------------------------------------------
# algorithmic work -- shared by bench.py and DESIGN.md
# --------------------------------------------------------------------------
def decode_algorithmic_bytes(seq_lens: Sequence[int], hq: int, hkv: int, d: int, dv: int,
                             kv_elt: int, q_elt: int = 2, idx_elt: int = 8, len_elt: int = 8) -> int:
    tot = int(sum(int(s) for s in seq_lens))
    b = len(seq_lens)
    return (tot * hkv * (d + dv) * kv_elt + b * hq * d * q_elt + b * hq * dv * q_elt
            + tot * 4 + b * (idx_elt + len_elt))


def extend_algorithmic_flops(prefix_lens: Sequence[int], extend_lens: Sequence[int],
                             hq: int, d: int) -> int:
    return 4 * d * hq * sum(int(e) * int(p) + int(e) * (int(e) + 1) // 2
                            for p, e in zip(prefix_lens, extend_lens))


def default_sm_scale(d: int) -> float:
    return 1.0 / math.sqrt(d)
