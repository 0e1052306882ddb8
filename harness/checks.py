"""Device-side cross-checks shared by tests/ and bench.py (no oracle import here: callers bring their own checker).

* ``naive_decode``: the library's one-warp-per-(request, head) fp32 kernel (``b200_debug_decode_naive``) on the same
  device buffers -- an independent twin of the tiled kernels at sizes the CPU oracle cannot reach.
* ``sample_rows_to_cpu``: cuts ``rows`` of a decode batch out of the device pool into a small self-contained CPU
  problem (own pool, own page table) that the CPU oracle finishes in seconds.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from scratchpad_b200 import _lib


def naive_decode(q3: torch.Tensor, kb: torch.Tensor, vb: torch.Tensor, req_to_token: torch.Tensor,
                 req_pool_indices: torch.Tensor, seq_lens: torch.Tensor, sm_scale: float, logit_cap: float,
                 kv_code: int, k_scale: Optional[torch.Tensor], v_scale: Optional[torch.Tensor]) -> torch.Tensor:
    """q3 [B, Hq, D] bf16 or fp16 (any token / head stride); returns o [B, Hq, Dv] in q3's dtype."""
    b, hq, d = q3.shape
    hkv, dv = kb.shape[1], vb.shape[2]
    o = torch.empty((b, hq, dv), dtype=q3.dtype, device=q3.device)
    if req_pool_indices.dtype != seq_lens.dtype:
        seq_lens = seq_lens.to(req_pool_indices.dtype)
    p = _lib.DecodeParams(
        q=q3.data_ptr(), q_stride_t=q3.stride(0), q_stride_h=q3.stride(1),
        k_buf=kb.data_ptr(), v_buf=vb.data_ptr(), k_stride_t=kb.stride(0), v_stride_t=vb.stride(0),
        o=o.data_ptr(), o_stride_t=o.stride(0), o_stride_h=o.stride(1),
        req_to_token=req_to_token.data_ptr(), req_to_token_stride=req_to_token.stride(0),
        bs=b, hq=hq, hkv=hkv, d=d, dv=dv, kv_dtype=kv_code, sm_scale=float(sm_scale), logit_cap=float(logit_cap),
        k_scale=None if k_scale is None else k_scale.data_ptr(), v_scale=None if v_scale is None else v_scale.data_ptr(),
        workspace=None, workspace_bytes=0, max_bs=b, force_path=-1,
        act_dtype=_lib.ACT_F16 if q3.dtype == torch.float16 else _lib.ACT_BF16)
    code = _lib.IDX_I64 if req_pool_indices.dtype == torch.int64 else _lib.IDX_I32
    rc = _lib.load().b200_debug_decode_naive(p, req_pool_indices.data_ptr(), seq_lens.data_ptr(), code,
                                             torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "b200_debug_decode_naive")
    return o


def sample_rows_to_cpu(rows: Sequence[int], q3: torch.Tensor, kb: torch.Tensor, vb: torch.Tensor,
                       req_to_token: torch.Tensor, req_pool_indices: torch.Tensor, seq_lens: torch.Tensor):
    """Returns CPU tensors (q [n,Hq,D], k_buf, v_buf [1+sum S, Hkv, D], req_to_token [n, max S], req_pool_indices,
    seq_lens) holding exactly the bytes requests ``rows`` of the batch read."""
    lens = [int(seq_lens[r]) for r in rows]
    width = max(max(lens), 1)
    table = torch.zeros((len(rows), width), dtype=torch.int32)
    ks, vs = [torch.zeros_like(kb[:1]).cpu()], [torch.zeros_like(vb[:1]).cpu()]
    cur = 1
    for i, (r, n) in enumerate(zip(rows, lens)):
        slots = req_to_token[int(req_pool_indices[r]), :n].long()
        ks.append(kb[slots].cpu())
        vs.append(vb[slots].cpu())
        table[i, :n] = torch.arange(cur, cur + n, dtype=torch.int32)
        cur += n
    idx = torch.tensor(list(rows), device=q3.device)
    return (q3[idx].cpu().contiguous(), torch.cat(ks), torch.cat(vs), table, torch.arange(len(rows), dtype=torch.int64),
            torch.tensor(lens, dtype=torch.int64))
