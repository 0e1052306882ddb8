"""The KV pool with the layout the reference engine uses, plus the fp8-e4m3 scale
metadata the new backend needs, and a page-granular allocator.

``MHATokenToKVPool`` keeps the reference's layout (``scratchpad/memory/pool.py:258-424``:
per-layer ``[size+page, Hkv, D]`` buffers, fp8 stored as uint8, slot 0 = padding).  The
backend itself only duck-types a pool (``k_buffer`` / ``v_buffer`` / ``dtype``), so the
reference's own unmodified class works too (tests/test_engine_boundary_gpu.py); this class is what
the reference's would become for the cuda_b200 backend
(BASELINE.json north_star, "KV pool layout/scale metadata"):

* ``kv_cache_dtype`` may be ``torch.float8_e4m3fn``; the pool then carries
  per-layer, per-kv-head fp32 dequantisation scales (``k_scale``/``v_scale``
  ``[layer_num, Hkv]`` on the device, default 1.0) which the kernels read
  directly -- per-tensor scaling is the special case of equal entries;
* ``set_kv_buffer`` is one fused CUDA launch (``b200_kv_store``) instead of
  two ``index_put_`` plus ``div_``/cast, and does not modify the caller's k/v.

Only CUDA tensors are accepted by ``set_kv_buffer``: the product path has no
CPU fallback.
"""

from __future__ import annotations

from typing import List, Optional, Tuple, Union

import torch

from . import _lib

FP8_DTYPES = (torch.float8_e4m3fn,)
PAGE_SIZES = (16, 32, 64, 128)


class PagedTokenToKVPoolAllocator:
    """Page-granular free list for ``page_size > 1`` -- the allocator the reference's paged call sites expect
    (``alloc_extend(prefix_lens, seq_lens, last_loc, extend_num_tokens)`` / ``alloc_decode(seq_lens, last_loc)`` /
    ``page_size`` / ``backup_state``: schedule_batch.py:754-828) but does not ship.

    Pages are numbered from 1 (page 0 = the dummy slots 0..page_size-1 padded tokens write to); slot =
    page * page_size + offset, and request position t always sits at offset ``t % page_size`` of its page.  That
    invariant is what lets the decode kernel fetch a page with one TMA box copy."""

    def __init__(self, size: int, page_size: int, dtype: torch.dtype, device: str, kvcache: "MHATokenToKVPool"):
        if page_size not in PAGE_SIZES:
            raise NotImplementedError(f"page_size must be one of {PAGE_SIZES}, got {page_size}")
        self.size = size
        self.page_size = page_size
        self.num_pages = size // page_size
        self.dtype = dtype
        self.device = device
        self._kvcache = kvcache
        self.debug_checks = False  # True: a double free raises instead of being ignored
        self.clear()

    def available_size(self):
        return len(self.free_pages) * self.page_size

    def get_kvcache(self):
        return self._kvcache

    def clear(self):
        self.free_pages = torch.arange(1, self.num_pages + 1, dtype=torch.int64, device=self.device)

    def backup_state(self):
        return self.free_pages

    def restore_state(self, free_pages):
        self.free_pages = free_pages

    def _take(self, n: int):
        pages, self.free_pages = self.free_pages[:n], self.free_pages[n:]
        return pages

    def alloc(self, need_size: int):
        """Whole pages for ``need_size`` tokens (page-aligned callers only)."""
        assert need_size % self.page_size == 0, "token-granular alloc on a paged pool"
        n = need_size // self.page_size
        if n > len(self.free_pages):
            return None
        pages = self._take(n)
        return (pages[:, None] * self.page_size + torch.arange(self.page_size, device=self.device)).reshape(-1)

    def alloc_extend(self, prefix_lens: torch.Tensor, seq_lens: torch.Tensor, last_loc: torch.Tensor,
                     extend_num_tokens: int):
        """Slots for positions ``prefix_lens[i] .. seq_lens[i]-1`` of every request: first the rest of the page
        ``last_loc[i]`` (slot of position ``prefix_lens[i]-1``, -1 if none) lives in, then fresh pages."""
        ps = self.page_size
        pre, seq, last = prefix_lens.tolist(), seq_lens.tolist(), last_loc.tolist()
        need_pages = sum((s + ps - 1) // ps - (p + ps - 1) // ps for p, s in zip(pre, seq))
        if need_pages > len(self.free_pages):
            return None
        pages = self._take(need_pages).tolist()
        out, cur = [], 0
        for p, s, l in zip(pre, seq, last):
            t = p
            while t < s and t % ps != 0:  # finish the partially filled page
                assert l >= 0 and (l + 1) % ps == t % ps, "last_loc is not the slot of position prefix_len - 1"
                l += 1
                out.append(l)
                t += 1
            while t < s:
                n = min(ps, s - t)
                out.extend(range(pages[cur] * ps, pages[cur] * ps + n))
                cur += 1
                t += n
        assert len(out) == extend_num_tokens
        return torch.tensor(out, dtype=torch.int64, device=self.device)

    def alloc_decode(self, seq_lens: torch.Tensor, last_loc: torch.Tensor):
        """One slot per request for position ``seq_lens[i]-1``: the next offset of the page ``last_loc[i]`` lives
        in, or offset 0 of a fresh page when that position starts a page."""
        ps = self.page_size
        pos = (seq_lens.to(torch.int64) - 1).to(last_loc.device)
        fresh = (pos % ps) == 0
        n = int(fresh.sum().item())
        if n > len(self.free_pages):
            return None
        out = last_loc.to(torch.int64) + 1
        out[fresh] = self._take(n).to(out.device) * ps
        return out

    def free(self, free_index: torch.Tensor):
        """Slots back to the free list.  The engine frees whole requests / radix nodes, i.e. runs that cover every
        used slot of a page; a page returns once, whichever of its slots are named (``unique(slot // page_size)``),
        and a page that is already free is not added twice."""
        if free_index.numel() == 0:
            return
        pages = torch.unique(free_index.to(torch.int64) // self.page_size).to(self.free_pages.device)
        pages = pages[pages > 0]  # page 0 holds the dummy slots and is never handed out
        if self.debug_checks:
            assert not torch.isin(pages, self.free_pages).any(), "double free of a KV page"
        else:
            pages = pages[~torch.isin(pages, self.free_pages)]
        self.free_pages = torch.cat((self.free_pages, pages))


class MHATokenToKVPool:
    def __init__(
        self,
        size: int,
        page_size: int,
        dtype: torch.dtype,
        head_num: int,
        head_dim: int,
        layer_num: int,
        device: str,
        enable_memory_saver: bool = False,
        v_head_dim: Optional[int] = None,
        fused_kv: bool = False,
    ):
        # The reference refuses page_size > 1 at init (model_runner.py:431-432) although its scheduler and radix cache
        # carry the paged code paths (schedule_batch.py:754-828, radix_cache.py:57-66).  The cuda_b200 backend takes
        # pages of 16..128 tokens: req_to_token keeps one slot id per token, a page's slots are consecutive, and the
        # decode kernel fetches whole pages with TMA box copies.
        if page_size != 1 and page_size not in PAGE_SIZES:
            raise NotImplementedError(f"page_size must be 1 or one of {PAGE_SIZES}, got {page_size}")
        self.size = size
        self.page_size = page_size
        self.dtype = dtype
        self.device = device
        self.store_dtype = torch.uint8 if dtype in FP8_DTYPES + (torch.float8_e5m2,) else dtype
        self.head_num = head_num
        self.head_dim = head_dim
        self.v_head_dim = v_head_dim or head_dim
        self.layer_num = layer_num
        # [size + page_size, head_num, head_dim] per layer; slot 0 = padded tokens
        self.fused_kv = bool(fused_kv)
        if self.fused_kv:
            # Layout option of the cuda_b200 backend: K and V rows of a token are adjacent in one
            # [size + page, 2, head_num, head_dim] allocation, exposed as strided k_buffer/v_buffer views
            # (same slot indexing, token stride doubled).  At TP8 a token row is only 128 B per buffer
            # (fp8 D=128 / bf16 D=64); random 128 B gathers cap at ~5.0 TB/s on B200 while 256 B ones
            # reach ~7.1 TB/s (tools/gather_bw.cu), so fusing the rows lifts the gather roofline.
            if self.v_head_dim != head_dim:
                raise ValueError("fused_kv needs head_dim == v_head_dim")
            self.kv_buffer = [
                torch.zeros((size + page_size, 2, head_num, head_dim), dtype=self.store_dtype, device=device)
                for _ in range(layer_num)
            ]
            self.k_buffer = [kv[:, 0] for kv in self.kv_buffer]
            self.v_buffer = [kv[:, 1] for kv in self.kv_buffer]
        else:
            self.k_buffer = [
                torch.zeros((size + page_size, head_num, head_dim), dtype=self.store_dtype, device=device)
                for _ in range(layer_num)
            ]
            self.v_buffer = [
                torch.zeros((size + page_size, head_num, self.v_head_dim), dtype=self.store_dtype, device=device)
                for _ in range(layer_num)
            ]
        # fp8 metadata: x ~= x_hat * scale[layer, head]
        self.k_scale = self.v_scale = None
        if dtype in FP8_DTYPES:
            self.k_scale = torch.ones((layer_num, head_num), dtype=torch.float32, device=device)
            self.v_scale = torch.ones((layer_num, head_num), dtype=torch.float32, device=device)
        self.capture_mode = False

    # ---- scale metadata ---------------------------------------------------
    def set_kv_scales(self, layer_id: int, k_scale, v_scale) -> None:
        """``k_scale``/``v_scale``: python float (per-tensor) or ``[Hkv]`` tensor (per-head)."""
        if self.k_scale is None:
            raise ValueError("kv scales only exist for an fp8 pool")
        for dst, s in ((self.k_scale, k_scale), (self.v_scale, v_scale)):
            if isinstance(s, torch.Tensor):
                dst[layer_id].copy_(s.to(torch.float32).reshape(-1).expand(self.head_num))
            else:
                dst[layer_id].fill_(float(s))

    def get_kv_scales(self, layer_id: int) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        if self.k_scale is None:
            return None, None
        return self.k_scale[layer_id], self.v_scale[layer_id]

    # ---- buffers ------------------------------------------------------------
    def get_key_buffer(self, layer_id: int):
        if self.store_dtype != self.dtype:
            return self.k_buffer[layer_id].view(self.dtype)
        return self.k_buffer[layer_id]

    def get_value_buffer(self, layer_id: int):
        if self.store_dtype != self.dtype:
            return self.v_buffer[layer_id].view(self.dtype)
        return self.v_buffer[layer_id]

    def get_kv_buffer(self, layer_id: int):
        return self.get_key_buffer(layer_id), self.get_value_buffer(layer_id)

    def get_kv_size_bytes(self):
        k = sum(b.numel() * b.element_size() for b in self.k_buffer)
        v = sum(b.numel() * b.element_size() for b in self.v_buffer)
        return k, v

    def kv_dtype_code(self) -> int:
        return kv_dtype_code(self.dtype)

    def set_kv_buffer(self, layer, loc: torch.Tensor, cache_k: torch.Tensor, cache_v: torch.Tensor,
                      k_scale=None, v_scale=None) -> None:
        """``k_buffer[layer][loc] = quant(cache_k)`` (pool.py:392-424) as one CUDA launch.

        ``k_scale``/``v_scale`` override the pool's stored scales when given
        (floats, the reference's calling convention)."""
        layer_id = layer.layer_id
        if k_scale is not None or v_scale is not None:
            self.set_kv_scales(layer_id, 1.0 if k_scale is None else k_scale, 1.0 if v_scale is None else v_scale)
        ks, vs = self.get_kv_scales(layer_id)
        kv_store(self.k_buffer[layer_id], self.v_buffer[layer_id], loc, cache_k, cache_v, self.kv_dtype_code(), ks, vs)


def kv_dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.bfloat16:
        return _lib.KV_BF16
    if dtype == torch.float16:
        return _lib.KV_F16
    if dtype == torch.float8_e4m3fn:
        return _lib.KV_FP8_E4M3
    raise NotImplementedError(f"kv cache dtype {dtype} is not supported by the cuda_b200 backend")


def act_dtype_code(dtype: torch.dtype) -> int:
    """The model dtype (q, the new tokens' k / v, o): bf16, or fp16 -- what the engine picks for fp16 and, by
    default, fp32 checkpoints (config/utils.py:41-43)."""
    if dtype == torch.bfloat16:
        return _lib.ACT_BF16
    if dtype == torch.float16:
        return _lib.ACT_F16
    raise NotImplementedError(f"model dtype {dtype} is not supported by the cuda_b200 backend (bfloat16 or float16)")


def check_dtype_pair(kv_code: int, act_code: int) -> None:
    """bf16 activations go with a bf16 or fp8-e4m3 pool, fp16 activations with an fp16 or fp8-e4m3 pool
    (kv_cache_dtype "auto" = the model dtype, model_runner.py:366-367)."""
    ok = {_lib.ACT_BF16: (_lib.KV_BF16, _lib.KV_FP8_E4M3), _lib.ACT_F16: (_lib.KV_F16, _lib.KV_FP8_E4M3)}[act_code]
    if kv_code not in ok:
        raise NotImplementedError("cuda_b200 backend: a 16-bit KV pool must have the model dtype "
                                  f"(kv dtype code {kv_code}, activation dtype code {act_code})")


def kv_store(kb: torch.Tensor, vb: torch.Tensor, loc: torch.Tensor, cache_k: torch.Tensor, cache_v: torch.Tensor,
             kv_code: int, k_scale: Optional[torch.Tensor], v_scale: Optional[torch.Tensor]) -> None:
    """``kb[loc] = quant(cache_k); vb[loc] = quant(cache_v)`` -- one ``b200_kv_store`` launch on the current stream.
    ``kb``/``vb``: one layer's ``[P+page, Hkv, D]`` pool buffers (bf16, or uint8 / fp8 storage); ``cache_k``/``cache_v``
    bf16 / fp16 ``[T, Hkv, D]`` (or ``[T, Hkv*D]``), possibly column slices of the qkv GEMM output; scales: device fp32
    ``[Hkv]`` or None.  The caller's k/v are not modified (the reference divides them in place, pool.py:404-412)."""
    if not cache_k.is_cuda:
        raise RuntimeError("cuda_b200 kv_store needs CUDA tensors (no CPU fallback)")
    if cache_k.dtype != cache_v.dtype:
        raise TypeError(f"k and v must have one dtype, got {cache_k.dtype} and {cache_v.dtype}")
    act_code = act_dtype_code(cache_k.dtype)
    check_dtype_pair(kv_code, act_code)
    if loc.dtype not in (torch.int64, torch.int32) or not loc.is_contiguous():
        raise TypeError(f"out_cache_loc must be a contiguous int64/int32 tensor, got {loc.dtype}")
    hkv, d, dv = kb.shape[1], kb.shape[2], vb.shape[2]
    cache_k = cache_k.view(-1, hkv, d)
    cache_v = cache_v.view(-1, hkv, dv)
    if cache_k.stride(2) != 1 or cache_v.stride(2) != 1:
        cache_k, cache_v = cache_k.contiguous(), cache_v.contiguous()
    if loc.numel() != cache_k.shape[0]:
        raise ValueError(f"out_cache_loc has {loc.numel()} entries for {cache_k.shape[0]} tokens")
    loc_code = _lib.IDX_I64 if loc.dtype == torch.int64 else _lib.IDX_I32
    rc = _lib.load().b200_kv_store(
        kb.data_ptr(), vb.data_ptr(), kb.stride(0), vb.stride(0),
        loc.data_ptr(), loc_code, cache_k.data_ptr(), cache_v.data_ptr(),
        cache_k.stride(0), cache_k.stride(1), cache_v.stride(0), cache_v.stride(1),
        cache_k.shape[0], hkv, d, dv, kv_code, act_code,
        k_scale.data_ptr() if k_scale is not None else None, v_scale.data_ptr() if v_scale is not None else None,
        torch.cuda.current_stream().cuda_stream,
    )
    _lib.check(rc, "b200_kv_store")
