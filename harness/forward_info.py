"""The slice of the engine's forward-pass types the attention backend consumes.

Mirrors ``scratchpad/model_executor/forward_info.py`` (reference): ``ForwardMode``
(:19-66, same integer values and predicates) and the ``ForwardBatch`` fields an
``AttentionBackend`` reads (:85-177).  The real engine passes its own
``ForwardBatch``; the backend only duck-types these attributes, so this module
exists for the standalone tests/bench and as documentation of the contract.
"""

from __future__ import annotations

from dataclasses import dataclass
from enum import IntEnum
from typing import Any, List, Optional

import torch


class ForwardMode(IntEnum):
    EXTEND = 1
    DECODE = 2
    MIXED = 3
    IDLE = 4
    TARGET_VERIFY = 5
    DRAFT_EXTEND = 6
    DUMMY_FIRST = 7

    def is_extend(self):
        return self == ForwardMode.EXTEND or self == ForwardMode.MIXED

    def is_decode(self):
        return self == ForwardMode.DECODE

    def is_mixed(self):
        return self == ForwardMode.MIXED

    def is_idle(self):
        return self == ForwardMode.IDLE

    def is_target_verify(self):
        return self == ForwardMode.TARGET_VERIFY

    def is_draft_extend(self):
        return self == ForwardMode.DRAFT_EXTEND

    def is_cuda_graph(self):
        return self.is_decode() or self.is_target_verify() or self.is_idle()

    def is_dummy_first(self):
        return self == ForwardMode.DUMMY_FIRST

    def is_decode_or_idle(self):
        return self.is_decode() or self.is_idle()


def compute_extend_start_loc(extend_seq_lens: torch.Tensor) -> torch.Tensor:
    """Exclusive cumsum, as ``compute_position_torch`` (forward_info.py:452-466)."""
    out = torch.zeros_like(extend_seq_lens)
    out[1:] = torch.cumsum(extend_seq_lens[:-1], dim=0)
    return out


@dataclass
class ForwardBatch:
    forward_mode: ForwardMode
    batch_size: int
    req_pool_indices: torch.Tensor      # [B] int64 (eager) / int32 (graph buffers)
    seq_lens: torch.Tensor              # [B]
    out_cache_loc: torch.Tensor         # [T] int64 KV slots of the tokens in this call
    seq_lens_sum: int = 0
    input_ids: Optional[torch.Tensor] = None
    seq_lens_cpu: Optional[torch.Tensor] = None
    positions: Optional[torch.Tensor] = None
    # extend
    extend_num_tokens: Optional[int] = None
    extend_seq_lens: Optional[torch.Tensor] = None     # int32
    extend_prefix_lens: Optional[torch.Tensor] = None  # int32
    extend_start_loc: Optional[torch.Tensor] = None    # int32
    extend_prefix_lens_cpu: Optional[List[int]] = None
    extend_seq_lens_cpu: Optional[List[int]] = None
    encoder_lens: Optional[torch.Tensor] = None
    # pools / backend
    req_to_token_pool: Any = None
    token_to_kv_pool: Any = None
    attn_backend: Any = None
    spec_info: Any = None
