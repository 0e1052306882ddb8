"""Rotary position embedding for the cuda_b200 path: the reference's module surface over one fused CUDA launch.

Mirrors ``scratchpad/nn/layers/rotary_embedding.py`` (reference): ``RotaryEmbedding`` (:52-170: constructor
arguments, ``cos_sin_cache`` buffer = ``[cos | sin]`` per position in the model dtype, in-place ``forward`` on
q and k), ``Llama3RotaryEmbedding`` (the Llama-3.1 / 3.2 frequency rescaling) and ``get_rope`` (:918-975, the
``None`` / ``"default"`` / ``"llama3"`` branches -- the ones the Llama family reaches).

``forward(positions, q, k)`` is the drop-in for the reference's ``forward_cuda`` (triteia_cuda's in-place
``rotary_embedding`` kernel, :131-164).  ``rope_kv_store`` is the fused form: the same
rotation plus the KV write ``MHATokenToKVPool.set_kv_buffer`` does (memory/pool.py:392-424) in ONE launch of
``b200_rope_kv_store`` -- the attention backend calls it when ``RadixAttention.forward`` is given
``rope=(rotary_emb, positions)`` instead of pre-rotated q/k.  Results are bit-identical to the reference's
``forward_native`` (tests/test_rope_gpu.py against oracle/rope_oracle.py and tests/golden/rope_*.npz).

CUDA tensors only; there is no CPU fallback.
"""

from __future__ import annotations

import math
from typing import Any, Dict, Optional, Tuple

import torch
from torch import nn

from . import _lib
from .pool import act_dtype_code, check_dtype_pair, kv_dtype_code


class RotaryEmbedding(nn.Module):
    def __init__(self, head_size: int, rotary_dim: int, max_position_embeddings: int, base: float,
                 is_neox_style: bool, dtype: torch.dtype) -> None:
        super().__init__()
        if rotary_dim % 16 or rotary_dim > head_size or head_size % 16:
            raise NotImplementedError("cuda_b200 rotary embedding: rotary_dim and head_size must be multiples of 16")
        self.head_size = head_size
        self.rotary_dim = rotary_dim
        self.max_position_embeddings = max_position_embeddings
        self.base = base
        self.is_neox_style = is_neox_style
        self.dtype = dtype
        self.register_buffer("cos_sin_cache", self._compute_cos_sin_cache().to(dtype), persistent=False)

    def _compute_inv_freq(self, base: float) -> torch.Tensor:
        return 1.0 / (base ** (torch.arange(0, self.rotary_dim, 2, dtype=torch.float) / self.rotary_dim))

    def _compute_cos_sin_cache(self) -> torch.Tensor:
        inv_freq = self._compute_inv_freq(self.base)
        t = torch.arange(self.max_position_embeddings, dtype=torch.float)
        freqs = torch.einsum("i,j -> ij", t, inv_freq)
        return torch.cat((freqs.cos(), freqs.sin()), dim=-1)

    # ---- the one launcher -------------------------------------------------------
    def _launch(self, positions, q, k, v=None, pool=None, layer_id=None, loc=None, write_back_k=True, scales=None) -> None:
        if not k.is_cuda:
            raise RuntimeError("cuda_b200 rotary embedding needs CUDA tensors (no CPU fallback)")
        if q is not None and q.dtype != k.dtype:
            raise TypeError(f"q and k must have one dtype, got {q.dtype} and {k.dtype}")
        act_code = act_dtype_code(k.dtype)  # bf16 or fp16: the arithmetic is rounded to this dtype
        cache = self.cos_sin_cache
        if cache.device != k.device:
            self.cos_sin_cache = cache = cache.to(k.device)
        if cache.dtype == torch.float32:
            cache_code = _lib.CACHE_F32
        elif cache.dtype == k.dtype:
            cache_code = _lib.CACHE_F16 if k.dtype == torch.float16 else _lib.CACHE_BF16
        else:  # the module builds its cache in the model dtype (reference rotary_embedding.py:90-99)
            raise NotImplementedError(f"cos_sin_cache dtype {cache.dtype} with {k.dtype} activations")
        hs = self.head_size
        positions = positions.flatten()
        if positions.dtype not in (torch.int64, torch.int32):
            raise TypeError(f"positions must be int32 or int64, got {positions.dtype}")
        t = positions.shape[0]
        k3 = k.view(t, -1, hs)
        q3 = q.view(t, -1, hs) if q is not None else None
        if k3.stride(2) != 1 or (q3 is not None and q3.stride(2) != 1):
            raise ValueError("q/k must be contiguous in the head dimension (rotated in place)")
        p = _lib.RopeStoreParams(
            q=q3.data_ptr() if q3 is not None else None,
            q_stride_t=q3.stride(0) if q3 is not None else 0, q_stride_h=q3.stride(1) if q3 is not None else 0,
            k=k3.data_ptr(), k_src_stride_t=k3.stride(0), k_src_stride_h=k3.stride(1),
            positions=positions.data_ptr(), cos_sin_cache=cache.data_ptr(), cache_stride=cache.stride(0),
            cache_dtype=cache_code, pos_dtype=_lib.IDX_I64 if positions.dtype == torch.int64 else _lib.IDX_I32,
            rotary_dim=self.rotary_dim, is_neox=int(self.is_neox_style),
            num_tokens=t, hq=q3.shape[1] if q3 is not None else 0, hkv=k3.shape[1], d=hs, dv=hs,
            kv_dtype=_lib.KV_BF16, write_back_k=int(write_back_k), act_dtype=act_code,
        )
        if pool is not None:
            kb, vb = pool.k_buffer[layer_id], pool.v_buffer[layer_id]
            v3 = v.view(t, k3.shape[1], -1)
            if v3.stride(2) != 1:
                v3 = v3.contiguous()
            if scales is not None:
                ks, vs = scales
            elif hasattr(pool, "get_kv_scales"):
                ks, vs = pool.get_kv_scales(layer_id)
            else:  # a pool without scale metadata (the reference's own class): unit scales
                ks = vs = None
            if loc.dtype not in (torch.int64, torch.int32) or not loc.is_contiguous():
                raise TypeError(f"out_cache_loc must be a contiguous int64/int32 tensor, got {loc.dtype}")
            p.v, p.v_src_stride_t, p.v_src_stride_h = v3.data_ptr(), v3.stride(0), v3.stride(1)
            p.k_buf, p.v_buf, p.k_stride_t, p.v_stride_t = kb.data_ptr(), vb.data_ptr(), kb.stride(0), vb.stride(0)
            p.loc = loc.data_ptr()
            p.loc_dtype = {torch.int64: _lib.IDX_I64, torch.int32: _lib.IDX_I32}[loc.dtype]
            p.dv = v3.shape[2]
            p.kv_dtype = kv_dtype_code(pool.dtype)
            check_dtype_pair(p.kv_dtype, act_code)
            p.k_scale = ks.data_ptr() if ks is not None else None
            p.v_scale = vs.data_ptr() if vs is not None else None
        rc = _lib.load().b200_rope_kv_store(p, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "b200_rope_kv_store")

    def forward(self, positions: torch.Tensor, query: torch.Tensor, key: torch.Tensor,
                offsets: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """In-place on query and key, like the reference's forward_cuda (:131-164)."""
        if offsets is not None:
            positions = positions + offsets
        self._launch(positions, query, key)
        return query, key

    def rope_kv_store(self, positions, query, key, value, pool, layer_id: int, loc, write_back_k: bool,
                      scales=None) -> None:
        """Rotate q in place, rotate k and write (k, v) into the pool at ``loc`` -- one launch.  ``scales``:
        (k_scale, v_scale) device fp32 ``[Hkv]`` tensors overriding the pool's own metadata."""
        self._launch(positions, query, key, value, pool, layer_id, loc, write_back_k, scales)

    def extra_repr(self) -> str:
        return (f"head_size={self.head_size}, rotary_dim={self.rotary_dim}, "
                f"max_position_embeddings={self.max_position_embeddings}, base={self.base}, "
                f"is_neox_style={self.is_neox_style}")


class Llama3RotaryEmbedding(RotaryEmbedding):
    def __init__(self, head_size, rotary_dim, max_position_embeddings, base, is_neox_style, dtype,
                 scaling_factor: float, low_freq_factor: float, high_freq_factor: float, orig_max_position: int):
        self.scaling_factor = scaling_factor
        self.low_freq_factor = low_freq_factor
        self.high_freq_factor = high_freq_factor
        self.orig_max_position = orig_max_position
        super().__init__(head_size, rotary_dim, max_position_embeddings, base, is_neox_style, dtype)

    def _compute_inv_freq(self, base: float) -> torch.Tensor:
        inv_freqs = super()._compute_inv_freq(base)
        low_freq_wavelen = self.orig_max_position / self.low_freq_factor
        high_freq_wavelen = self.orig_max_position / self.high_freq_factor
        wave_len = 2 * math.pi / inv_freqs
        if self.low_freq_factor != self.high_freq_factor:
            smooth = (self.orig_max_position / wave_len - self.low_freq_factor) / (
                self.high_freq_factor - self.low_freq_factor)
        else:
            smooth = 0
        return torch.where(
            wave_len < high_freq_wavelen, inv_freqs,
            torch.where(wave_len > low_freq_wavelen, inv_freqs / self.scaling_factor,
                        (1 - smooth) * inv_freqs / self.scaling_factor + smooth * inv_freqs))


_ROPE_DICT: Dict[Tuple, RotaryEmbedding] = {}


def get_rope(head_size: int, rotary_dim: int, max_position: int, base: float, is_neox_style: bool = True,
             rope_scaling: Optional[Dict[str, Any]] = None, dtype: Optional[torch.dtype] = None,
             partial_rotary_factor: float = 1.0) -> RotaryEmbedding:
    if dtype is None:
        dtype = torch.get_default_dtype()
    args = None
    if rope_scaling is not None:
        args = tuple((k, tuple(v) if isinstance(v, list) else v) for k, v in rope_scaling.items())
    if partial_rotary_factor < 1.0:
        rotary_dim = int(rotary_dim * partial_rotary_factor)
    key = (head_size, rotary_dim, max_position, base, is_neox_style, args, dtype)
    if key in _ROPE_DICT:
        return _ROPE_DICT[key]
    kind = "default" if rope_scaling is None else rope_scaling.get("rope_type", rope_scaling.get("type"))
    if kind == "default":
        if rope_scaling is not None and "mrope_section" in rope_scaling:
            raise NotImplementedError("cuda_b200 rotary embedding: mrope is out of scope (vision models)")
        emb = RotaryEmbedding(head_size, rotary_dim, max_position, base, is_neox_style, dtype)
    elif kind == "llama3":
        emb = Llama3RotaryEmbedding(head_size, rotary_dim, max_position, base, is_neox_style, dtype,
                                    rope_scaling["factor"], rope_scaling["low_freq_factor"],
                                    rope_scaling["high_freq_factor"],
                                    rope_scaling["original_max_position_embeddings"])
    else:
        raise NotImplementedError(f"cuda_b200 rotary embedding: rope_type {kind!r} is not built "
                                  "(any module exposing cos_sin_cache/rotary_dim/is_neox_style can be passed instead)")
    _ROPE_DICT[key] = emb
    return emb
