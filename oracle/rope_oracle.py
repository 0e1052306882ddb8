"""CPU oracle for rotary position embedding + KV write (the fused ``b200_rope_kv_store`` launch).

TEST INFRASTRUCTURE ONLY (same rule as oracle/attn_oracle.py: imported by ``tests/`` only, never by
``scratchpad_b200/``).

What it restates (paths relative to /root/reference):

* ``RotaryEmbedding._compute_cos_sin_cache`` -- ``cache[pos] = [cos(pos*inv_freq) | sin(pos*inv_freq)]`` cast to the
  model dtype: ``scratchpad/nn/layers/rotary_embedding.py:78-99,73``
* ``Llama3RotaryEmbedding._compute_inv_freq`` (the Llama-3.1/3.2 frequency rescaling): ``rotary_embedding.py``
  class ``Llama3RotaryEmbedding``
* ``RotaryEmbedding.forward_native`` + ``_apply_rotary_emb`` -- rotate the first ``rotary_dim`` elements of every
  head of q and k, neox (halves) or GPT-J (interleaved) pairing, pass the rest through:
  ``rotary_embedding.py:24-49,101-129``.  The arithmetic is torch bf16 elementwise: every product and every
  sum is rounded to bf16.  Restated here with explicit fp32 ops + explicit roundings (no torch bf16 kernels),
  so the check against the golden vectors is a real one.
* the call site ``q, k = self.rotary_emb(positions, q, k)`` followed by attention's KV write:
  ``scratchpad/nn/models/llama/llama.py:149-152``, ``scratchpad/memory/pool.py:392-424``.

Parity pin: ``oracle/gen_golden_rope.py`` imports the reference's ``rotary_embedding.py`` in this container, runs
``forward_native`` on seeded bf16 inputs (plain, Llama-3-scaled, GPT-J-style and partial-rotary modules) and
commits inputs + outputs as ``tests/golden/rope_*.npz``; ``tests/test_rope_oracle.py`` holds this file to them
bit-exactly.
"""

from __future__ import annotations

import math

import torch


def _round_to(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """fp32 -> nearest-even model dtype (bf16 or fp16) -> fp32."""
    return x.to(dtype).to(torch.float32)


def inv_freq_ref(rotary_dim: int, base: float) -> torch.Tensor:
    return 1.0 / (base ** (torch.arange(0, rotary_dim, 2, dtype=torch.float) / rotary_dim))


def llama3_inv_freq_ref(rotary_dim: int, base: float, factor: float, low_freq_factor: float,
                        high_freq_factor: float, orig_max_position: int) -> torch.Tensor:
    inv = inv_freq_ref(rotary_dim, base)
    low_wavelen = orig_max_position / low_freq_factor
    high_wavelen = orig_max_position / high_freq_factor
    wavelen = 2 * math.pi / inv
    if low_freq_factor != high_freq_factor:
        smooth = (orig_max_position / wavelen - low_freq_factor) / (high_freq_factor - low_freq_factor)
    else:
        smooth = 0
    return torch.where(wavelen < high_wavelen, inv,
                       torch.where(wavelen > low_wavelen, inv / factor,
                                   (1 - smooth) * inv / factor + smooth * inv))


def cos_sin_cache_ref(inv_freq: torch.Tensor, max_position: int, dtype=torch.bfloat16) -> torch.Tensor:
    t = torch.arange(max_position, dtype=torch.float)
    freqs = t[:, None] * inv_freq[None, :]
    return torch.cat((freqs.cos(), freqs.sin()), dim=-1).to(dtype)


def rope_ref(positions: torch.Tensor, x: torch.Tensor, cos_sin_cache: torch.Tensor, rotary_dim: int,
             is_neox: bool) -> torch.Tensor:
    """x: bf16 or fp16 [T, H, D] -> rotated [T, H, D] in the same dtype (forward_native's arithmetic, explicit
    roundings to x.dtype: torch evaluates 16-bit elementwise ops in fp32 and rounds each result once)."""
    t, h, d = x.shape
    rnd = lambda y: _round_to(y, x.dtype)
    cs = cos_sin_cache.index_select(0, positions.long().flatten())
    half = rotary_dim // 2
    cos = rnd(cs[:, :half].float())[:, None, :]  # .to(x.dtype)
    sin = rnd(cs[:, half:rotary_dim].float())[:, None, :]
    xf = x.float()
    rot, rest = xf[..., :rotary_dim], xf[..., rotary_dim:]
    if is_neox:
        x1, x2 = rot[..., :half], rot[..., half:]
    else:
        x1, x2 = rot[..., 0::2], rot[..., 1::2]
    o1 = rnd(rnd(x1 * cos) - rnd(x2 * sin))
    o2 = rnd(rnd(x2 * cos) + rnd(x1 * sin))
    if is_neox:
        out = torch.cat((o1, o2), dim=-1)
    else:
        out = torch.stack((o1, o2), dim=-1).flatten(-2)
    return torch.cat((out, rest), dim=-1).to(x.dtype)
