"""Per-layer attention module: reshapes k/v and delegates to the backend.

Same attributes and forward contract as the reference's
``scratchpad/nn/attention/radix_attention.py:6-57``; kept here so the backend
can be exercised without importing the engine.
"""

from __future__ import annotations

from torch import nn


class RadixAttention(nn.Module):
    def __init__(
        self,
        num_heads: int,
        head_dim: int,
        scaling: float,
        num_kv_heads: int,
        layer_id: int,
        logit_cap: float = 0.0,
        v_head_dim: int = -1,
        sliding_window_size: int = -1,
        is_cross_attention: bool = False,
        prefix: str = "",
        use_irope: bool = False,
    ):
        super().__init__()
        self.tp_q_head_num = num_heads
        self.tp_k_head_num = num_kv_heads
        self.tp_v_head_num = num_kv_heads
        self.head_dim = head_dim
        self.qk_head_dim = head_dim
        self.v_head_dim = v_head_dim if v_head_dim != -1 else head_dim
        self.scaling = scaling
        self.layer_id = layer_id
        self.logit_cap = logit_cap
        self.sliding_window_size = sliding_window_size or -1
        self.is_cross_attention = is_cross_attention
        self.k_scale = None
        self.v_scale = None
        self.use_irope = use_irope

    def forward(self, q, k, v, forward_batch, save_kv_cache: bool = True, rope=None):
        """``rope=(rotary_emb, positions)``: q and k arrive UNrotated and the backend applies the rotary
        embedding fused with its KV write (cuda_b200 only; see rotary_embedding.py).  Default: the reference's
        contract (radix_attention.py:41-57)."""
        if k is not None:
            assert v is not None
            k = k.view(-1, self.tp_k_head_num, self.qk_head_dim)
            v = v.view(-1, self.tp_v_head_num, self.v_head_dim)
        if rope is None:
            return forward_batch.attn_backend.forward(q, k, v, self, forward_batch, save_kv_cache)
        return forward_batch.attn_backend.forward(q, k, v, self, forward_batch, save_kv_cache, rope=rope)
