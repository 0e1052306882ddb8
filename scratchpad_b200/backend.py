"""`cuda_b200` attention backend: the drop-in for the engine's AttentionBackend.

``AttentionBackend`` restates the reference ABC
(``scratchpad/nn/attention/backend.py:11-105``: same method names, argument
order and dispatch rule).  ``CudaB200AttnBackend`` implements it the way
``TritonAttnBackend`` does (``scratchpad/nn/attention/triton_backend.py:17-196``:
KV write first, then attend; output ``[T, Hq*Dv]`` in q's dtype) but over the
hand-written sm_100a kernels behind the C ABI in ``include/b200_attn.h``.

Selection site in the engine (one ``elif`` in
``scratchpad/model_executor/model_runner.py:453-470``)::

    elif self.server_args.attention_backend == "cuda_b200":
        from scratchpad_b200 import CudaB200AttnBackend
        self.attn_backend = CudaB200AttnBackend(self)

No Triton, no dispatch to another backend, no CPU fallback: every forward call
goes through ``libb200attn.so`` and raises if it is missing.
"""

from __future__ import annotations

import os
from abc import ABC, abstractmethod
from typing import Optional

import torch

from . import _lib
from .pool import act_dtype_code, check_dtype_pair, kv_dtype_code, kv_store


class AttentionBackend(ABC):
    """The base class of attention backends (reference backend.py:11-105)."""

    @abstractmethod
    def init_forward_metadata(self, forward_batch):
        raise NotImplementedError()

    def init_cuda_graph_state(self, max_bs: int):
        raise NotImplementedError()

    def init_forward_metadata_capture_cuda_graph(self, bs, num_tokens, req_pool_indices, seq_lens,
                                                 encoder_lens, forward_mode, spec_info):
        raise NotImplementedError()

    def init_forward_metadata_replay_cuda_graph(self, bs, req_pool_indices, seq_lens, seq_lens_sum,
                                                encoder_lens, forward_mode, spec_info, seq_lens_cpu=None):
        raise NotImplementedError()

    def get_cuda_graph_seq_len_fill_value(self):
        raise NotImplementedError()

    def forward(self, q, k, v, layer, forward_batch, save_kv_cache: bool = True, rope=None):
        # rope is this backend's one extension of the reference signature: (rotary_emb, positions) asks the backend
        # to apply the rotary embedding itself, fused with the KV write (rotary_embedding.py); None = the reference's
        # contract, q and k arrive rotated
        kw = {} if rope is None else {"rope": rope}
        if forward_batch.forward_mode.is_decode():
            return self.forward_decode(q, k, v, layer, forward_batch, save_kv_cache=save_kv_cache, **kw)
        else:
            return self.forward_extend(q, k, v, layer, forward_batch, save_kv_cache=save_kv_cache, **kw)

    def forward_decode(self, q, k, v, layer, forward_batch, save_kv_cache: bool = True):
        raise NotImplementedError()

    def forward_extend(self, q, k, v, layer, forward_batch, save_kv_cache: bool = True):
        raise NotImplementedError()


def _idx_code(t: torch.Tensor) -> int:
    if t.dtype == torch.int64:
        return _lib.IDX_I64
    if t.dtype == torch.int32:
        return _lib.IDX_I32
    raise TypeError(f"index tensors must be int32 or int64, got {t.dtype}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _i32(t: torch.Tensor, what: str) -> torch.Tensor:
    """The C ABI reads these as contiguous int32 (forward_info.py:258-263,408-410 produce exactly that); another
    integration's int64 or strided tensor is converted here instead of being misread."""
    if t.dtype != torch.int32 or not t.is_contiguous():
        if t.dtype not in (torch.int32, torch.int64):
            raise TypeError(f"{what} must be an integer tensor, got {t.dtype}")
        t = t.to(torch.int32).contiguous()
    return t


def _idx(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.dtype not in (torch.int32, torch.int64):
        raise TypeError(f"{what} must be int32 or int64, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


class CudaB200AttnBackend(AttentionBackend):
    def __init__(self, model_runner, force_path: int = _lib.PATH_AUTO, cascade: Optional[bool] = None,
                 tma_contiguous_runs: Optional[bool] = None):
        super().__init__()
        self.lib = _lib.load()  # raises when the extension is not built
        cfg = model_runner.model_config
        tp = model_runner.tp_size
        if getattr(model_runner.server_args, "enable_dp_attention", False):
            self.num_head = cfg.num_attention_heads
        else:
            self.num_head = cfg.num_attention_heads // tp
        self.num_kv_head = cfg.get_num_kv_heads(tp)
        self.head_dim = cfg.head_dim
        self.device = model_runner.device
        self.req_to_token_pool = model_runner.req_to_token_pool
        self.token_to_kv_pool = model_runner.token_to_kv_pool
        self.v_head_dim = self.token_to_kv_pool.get_value_buffer(0).shape[-1]
        # any pool with the reference's layout works (k_buffer / v_buffer lists of [P+page, Hkv, D], .dtype); the
        # fp8 scale metadata of scratchpad_b200.pool.MHATokenToKVPool is used when the pool has it
        self.kv_dtype = kv_dtype_code(self.token_to_kv_pool.dtype)
        self._layer_scales = {}
        self.force_path = force_path
        # Shared-prefix cascade decode: the prefix every request of a decode batch shares (same slot ids in
        # req_to_token, which is how RadixAttention represents a shared system prompt) is attended once per
        # kv head as a dense tcgen05 contraction; detection is on device, so it costs nothing to leave on.
        if cascade is None:
            cascade = bool(getattr(model_runner.server_args, "b200_shared_prefix_cascade", True))
        self.cascade = 1 if cascade else 0
        # TMA for KV tiles whose slots are consecutive in the token-granular pool (an unfragmented allocator hands out
        # runs, memory/pool.py:240-253): on by default -- a CTA whose first tile is scattered runs the pure row-gather
        # loop at the plain kernel's speed, so the setting costs nothing on a fragmented pool
        if tma_contiguous_runs is None:
            tma_contiguous_runs = bool(getattr(model_runner.server_args, "b200_tma_contiguous_runs", True))
        self.tma_contiguous_runs = bool(tma_contiguous_runs)
        if getattr(model_runner, "sliding_window_size", None) not in (None, -1):
            raise NotImplementedError("cuda_b200 backend: sliding-window attention is not supported")
        # the model dtype: bf16, or fp16 (what the engine serves fp16 and, by default, fp32 checkpoints in,
        # config/utils.py:41-43); a 16-bit pool has the model dtype (kv_cache_dtype "auto", model_runner.py:366-367)
        self.act_dtype = act_dtype_code(model_runner.dtype)
        check_dtype_pair(self.kv_dtype, self.act_dtype)
        self._graph_state_frozen = False
        self.cuda_graph_max_seq_len = cfg.context_len
        self._check_pages = os.environ.get("B200_CHECK_PAGES", "0") == "1"
        self.max_bs = 0
        self.workspace = None
        self.config = None
        self.forward_metadata = None
        self._planned_for = None
        # bench.py only: (begin, end) raw cudaEvent_t handles recorded around the split-KV kernel
        self.profile_events = None
        self._ensure_workspace(int(getattr(self.req_to_token_pool, "size", 256)))

    # ---- workspace / launch configuration ---------------------------------
    def _ensure_workspace(self, max_bs: int) -> None:
        if self.workspace is not None and max_bs <= self.max_bs:
            return
        if self.workspace is not None and self._graph_state_frozen:
            # captured graphs hold this workspace's address and a layout that depends on max_bs: growing would leave
            # them reading freed memory (the reference sizes its graph state once, cuda_graph_runner.py:178-181)
            raise RuntimeError(f"cuda_b200 backend: batch size {max_bs} exceeds the workspace sized for {self.max_bs} "
                               "after CUDA-graph capture; call init_cuda_graph_state(max_bs) with the largest batch "
                               "before capturing")
        cfg = _lib.AttnConfig()
        rc = self.lib.b200_attn_get_config(max_bs, self.num_head, self.num_kv_head, self.head_dim,
                                           self.v_head_dim, self.kv_dtype, self.force_path, cfg)
        _lib.check(rc, "b200_attn_get_config")
        self.config = cfg
        self.max_bs = max_bs
        self.workspace = torch.empty(int(cfg.workspace_bytes), dtype=torch.uint8, device=self.device)

    def _plan(self, bs: int, req_pool_indices: torch.Tensor, seq_lens: torch.Tensor) -> None:
        req_pool_indices = _idx(req_pool_indices, "req_pool_indices")
        seq_lens = _idx(seq_lens, "seq_lens")
        if req_pool_indices.dtype != seq_lens.dtype:
            seq_lens = seq_lens.to(req_pool_indices.dtype)
        rc = self.lib.b200_attn_plan(
            self.workspace.data_ptr(), self.workspace.numel(), self.max_bs,
            req_pool_indices.data_ptr(), seq_lens.data_ptr(), _idx_code(req_pool_indices), bs,
            self.num_head, self.num_kv_head, self.head_dim, self.v_head_dim, self.kv_dtype,
            self.force_path, self.req_to_token_pool.req_to_token.data_ptr(),
            self.req_to_token_pool.req_to_token.stride(0), self.cascade,
            torch.cuda.current_stream().cuda_stream,
        )
        _lib.check(rc, "b200_attn_plan")
        self._planned_for = bs
        if self._check_pages and int(getattr(self.token_to_kv_pool, "page_size", 1)) > 1:
            self._verify_paged_table(bs, req_pool_indices, seq_lens)

    def _verify_paged_table(self, bs: int, req_pool_indices: torch.Tensor, seq_lens: torch.Tensor) -> None:
        """B200_CHECK_PAGES=1 only (synchronises): the paged decode path trusts the allocator's invariant
        slot(t) = page_start + t % page_size; a table that breaks it would be read wrongly, so fail loudly."""
        r2t = self.req_to_token_pool.req_to_token
        bad = torch.zeros(1, dtype=torch.int32, device=r2t.device)
        rc = self.lib.b200_debug_check_paged(r2t.data_ptr(), r2t.stride(0), req_pool_indices.data_ptr(),
                                             seq_lens.data_ptr(), _idx_code(req_pool_indices), bs,
                                             int(self.token_to_kv_pool.page_size), bad.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "b200_debug_check_paged")
        n = int(bad.item())
        if n:
            raise RuntimeError(f"cuda_b200 backend: req_to_token breaks the page_size={self.token_to_kv_pool.page_size} "
                               f"layout contract at {n} positions (slot != page_start + position % page_size)")

    # ---- AttentionBackend surface -------------------------------------------
    def init_forward_metadata(self, forward_batch):
        """Once per forward batch: balance the batch's KV tiles over the persistent
        grid on device.  No kv_indices flatten, no host sync (cf.
        triton_backend.py:48-68, flashinfer_backend.py:623-670)."""
        if forward_batch.forward_mode.is_decode():
            bs = forward_batch.batch_size
            self._ensure_workspace(bs)
            self._plan(bs, forward_batch.req_pool_indices, forward_batch.seq_lens)
            self.forward_metadata = ("decode", bs)
        else:
            lens = forward_batch.extend_seq_lens_cpu
            if lens is None:  # the reference's Triton backend syncs here too (triton_backend.py:66)
                lens = forward_batch.extend_seq_lens.tolist()
            self.forward_metadata = ("extend", max(lens) if len(lens) else 0)

    def init_cuda_graph_state(self, max_bs: int):
        self._ensure_workspace(max_bs)

    def init_forward_metadata_capture_cuda_graph(self, bs, num_tokens, req_pool_indices, seq_lens,
                                                 encoder_lens, forward_mode, spec_info):
        assert encoder_lens is None, "Not supported"
        assert forward_mode.is_decode(), "Not supported"
        assert spec_info is None, "Not supported"
        assert bs <= self.max_bs, "init_cuda_graph_state(max_bs) must cover every captured batch size"
        self._graph_state_frozen = True
        self._plan(bs, req_pool_indices[:bs], seq_lens[:bs])
        self.forward_metadata = ("decode", bs)

    def init_forward_metadata_replay_cuda_graph(self, bs, req_pool_indices, seq_lens, seq_lens_sum,
                                                encoder_lens, forward_mode, spec_info, seq_lens_cpu=None):
        # same static workspace the captured kernels read; refreshed in place, no sync
        self._plan(bs, req_pool_indices[:bs], seq_lens[:bs])
        self.forward_metadata = ("decode", bs)

    def get_cuda_graph_seq_len_fill_value(self):
        return 1

    def _scales(self, layer, pool=None):
        """Device fp32 [Hkv] dequantisation scales of this layer's K and V, or (None, None) = 1.0."""
        pool = pool if pool is not None else self.token_to_kv_pool
        if hasattr(pool, "get_kv_scales"):
            return pool.get_kv_scales(layer.layer_id)
        if self.kv_dtype != _lib.KV_FP8_E4M3:
            return None, None
        # a pool without scale metadata (the reference's class): per-tensor floats carried by the layer, the way
        # the reference's FlashInfer backend reads them (flashinfer_backend.py:387-388,457-458)
        ks = getattr(layer, "k_scale_float", None)
        vs = getattr(layer, "v_scale_float", None)
        key = (layer.layer_id, ks, vs)
        hit = self._layer_scales.get(layer.layer_id)
        if hit is None or hit[0] != key:
            mk = lambda x: torch.full((self.num_kv_head,), 1.0 if x is None else float(x), dtype=torch.float32,
                                      device=self.device)
            hit = (key, mk(ks), mk(vs))
            self._layer_scales[layer.layer_id] = hit
        return hit[1], hit[2]

    def _store_kv(self, pool, layer, loc, k, v) -> None:
        """MHATokenToKVPool.set_kv_buffer (pool.py:392-424) as one launch, on whatever pool class the engine uses."""
        ks, vs = self._scales(layer, pool)
        kv_store(pool.k_buffer[layer.layer_id], pool.v_buffer[layer.layer_id], loc, k, v, self.kv_dtype, ks, vs)

    def _page_hint(self, pool) -> int:
        """b200_decode_params.page_size: the pool's page size, or PAGE_AUTO for a token-granular pool when
        ``tma_contiguous_runs`` is on (tiles whose slots happen to be consecutive are then fetched by TMA)."""
        ps = int(getattr(pool, "page_size", 1))
        if ps > 1:
            return ps
        return _lib.PAGE_AUTO if self.tma_contiguous_runs else 1

    def _logit_cap(self, layer) -> float:
        cap = getattr(layer, "logit_cap", 0.0)
        return float(cap) if cap else 0.0

    def _act_code(self, *tensors) -> int:
        """The kernels read q / k / v as the model dtype the backend was built for; anything else would be misread."""
        for t in tensors:
            if act_dtype_code(t.dtype) != self.act_dtype:
                raise TypeError(f"cuda_b200 backend: built for model dtype code {self.act_dtype}, got a {t.dtype} tensor")
        return self.act_dtype

    def _check_layer(self, layer) -> None:
        if getattr(layer, "is_cross_attention", False):
            raise NotImplementedError("cuda_b200 backend: cross attention is not supported")
        if getattr(layer, "sliding_window_size", -1) not in (None, -1):
            raise NotImplementedError("cuda_b200 backend: sliding-window attention is not supported")

    def _rope_and_store(self, rope, q, k, v, layer, forward_batch, save_kv_cache: bool, write_back_k: bool) -> None:
        """q (in place) and k rotated, (k, v) written to the pool: one launch instead of the model's rotary kernel
        (llama.py:149-152) plus set_kv_buffer's two index_put_ (pool.py:392-424)."""
        rotary_emb, positions = rope
        if save_kv_cache:
            pool = forward_batch.token_to_kv_pool
            rotary_emb.rope_kv_store(positions, q, k, v, pool, layer.layer_id, forward_batch.out_cache_loc,
                                     write_back_k, scales=self._scales(layer, pool))
        else:
            rotary_emb(positions, q, k)

    def forward_decode(self, q, k, v, layer, forward_batch, save_kv_cache: bool = True, rope=None):
        self._check_layer(layer)
        hq, d, dv = layer.tp_q_head_num, layer.qk_head_dim, layer.v_head_dim
        q = q.reshape(-1, hq * d)
        if q.stride(1) != 1:
            q = q.contiguous()
        o = q.new_empty((q.shape[0], hq * dv))
        pool = forward_batch.token_to_kv_pool
        if rope is not None:
            # decode reads the new token's K from the pool, so the rotated k is not written back
            self._rope_and_store(rope, q, k, v, layer, forward_batch, save_kv_cache, write_back_k=not save_kv_cache)
        elif save_kv_cache:
            self._store_kv(pool, layer, forward_batch.out_cache_loc, k, v)
        kind, bs = self.forward_metadata
        assert kind == "decode" and bs == q.shape[0], "init_forward_metadata was not called for this batch"
        kb, vb = pool.k_buffer[layer.layer_id], pool.v_buffer[layer.layer_id]
        r2t = forward_batch.req_to_token_pool.req_to_token
        ks, vs = self._scales(layer, pool)
        p = _lib.DecodeParams(
            q=q.data_ptr(), q_stride_t=q.stride(0), q_stride_h=d,
            k_buf=kb.data_ptr(), v_buf=vb.data_ptr(), k_stride_t=kb.stride(0), v_stride_t=vb.stride(0),
            o=o.data_ptr(), o_stride_t=o.stride(0), o_stride_h=dv,
            req_to_token=r2t.data_ptr(), req_to_token_stride=r2t.stride(0),
            bs=bs, hq=hq, hkv=layer.tp_k_head_num, d=d, dv=dv, kv_dtype=self.kv_dtype,
            sm_scale=float(layer.scaling), logit_cap=self._logit_cap(layer),
            k_scale=_ptr(ks), v_scale=_ptr(vs),
            workspace=self.workspace.data_ptr(), workspace_bytes=self.workspace.numel(),
            max_bs=self.max_bs, force_path=self.force_path, cascade=self.cascade,
            page_size=self._page_hint(pool), kv_pool_rows=kb.shape[0], act_dtype=self._act_code(q),
        )
        if self.profile_events is not None:
            p.ev_kernel_begin, p.ev_kernel_end = self.profile_events
        rc = self.lib.b200_attn_decode(p, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "b200_attn_decode")
        return o

    def forward_extend(self, q, k, v, layer, forward_batch, save_kv_cache: bool = True, rope=None):
        self._check_layer(layer)
        hq, hkv, d, dv = layer.tp_q_head_num, layer.tp_k_head_num, layer.qk_head_dim, layer.v_head_dim
        q3 = q.view(-1, hq, d) if q.dim() == 2 else q
        if q3.stride(2) != 1:
            q3 = q3.contiguous()
        o = q.new_empty((q3.shape[0], hq * dv))
        pool = forward_batch.token_to_kv_pool
        if rope is not None:
            # the extend kernels read the new tokens' K from the k argument: rotated k is written back in place
            self._rope_and_store(rope, q3, k, v, layer, forward_batch, save_kv_cache, write_back_k=True)
        elif save_kv_cache:
            self._store_kv(pool, layer, forward_batch.out_cache_loc, k, v)
        kind, max_extend_len = self.forward_metadata
        assert kind == "extend", "init_forward_metadata was not called for this batch"
        k3, v3 = k.view(-1, hkv, d), v.view(-1, hkv, dv)
        if k3.stride(2) != 1 or v3.stride(2) != 1:
            k3, v3 = k3.contiguous(), v3.contiguous()
        kb, vb = pool.k_buffer[layer.layer_id], pool.v_buffer[layer.layer_id]
        r2t = forward_batch.req_to_token_pool.req_to_token
        ks, vs = self._scales(layer, pool)
        rpi, sl = _idx(forward_batch.req_pool_indices, "req_pool_indices"), _idx(forward_batch.seq_lens, "seq_lens")
        if rpi.dtype != sl.dtype:
            sl = sl.to(rpi.dtype)
        ext_lens = _i32(forward_batch.extend_seq_lens, "extend_seq_lens")
        ext_start = _i32(forward_batch.extend_start_loc, "extend_start_loc")
        p = _lib.ExtendParams(
            q=q3.data_ptr(), q_stride_t=q3.stride(0), q_stride_h=q3.stride(1),
            k_new=k3.data_ptr(), v_new=v3.data_ptr(),
            kn_stride_t=k3.stride(0), kn_stride_h=k3.stride(1), vn_stride_t=v3.stride(0), vn_stride_h=v3.stride(1),
            k_buf=kb.data_ptr(), v_buf=vb.data_ptr(), k_stride_t=kb.stride(0), v_stride_t=vb.stride(0),
            o=o.data_ptr(), o_stride_t=o.stride(0), o_stride_h=dv,
            req_to_token=r2t.data_ptr(), req_to_token_stride=r2t.stride(0),
            req_pool_indices=rpi.data_ptr(), seq_lens=sl.data_ptr(), idx_dtype=_idx_code(rpi),
            extend_seq_lens=ext_lens.data_ptr(), extend_start_loc=ext_start.data_ptr(),
            bs=forward_batch.batch_size, hq=hq, hkv=hkv, d=d, dv=dv,
            total_tokens=q3.shape[0], max_extend_len=int(max_extend_len), kv_dtype=self.kv_dtype,
            sm_scale=float(layer.scaling), logit_cap=self._logit_cap(layer),
            k_scale=_ptr(ks), v_scale=_ptr(vs), force_path=self.force_path,
            page_size=self._page_hint(pool), kv_pool_rows=kb.shape[0], act_dtype=self._act_code(q3, k3, v3),
        )
        rc = self.lib.b200_attn_extend(p, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "b200_attn_extend")
        return o
