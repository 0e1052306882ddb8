"""KV-head sharding of the attention path across the GPUs of one box.

Mirrors the engine's tensor-parallel head layout: ``QKVParallelLinear`` gives each TP rank
``Hq/tp`` q heads and ``Hkv/tp`` kv heads, or *replicates* kv heads when ``tp > Hkv``
(``scratchpad/nn/layers/linear.py:716-722``; ``ModelConfig.get_num_kv_heads``,
``scratchpad/config/model_config.py:159-166``), and the KV pool is sized with the local head count
(``scratchpad/model_executor/model_runner.py:420-429``).  Attention needs no collective: a rank's
output ``[T, Hq_local*Dv]`` feeds the row-parallel ``o_proj`` whose all-reduce is the engine's.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class HeadShard:
    rank: int
    tp_size: int
    hq_local: int
    hkv_local: int
    q_head_start: int    # first global q head owned by this rank
    kv_head_start: int   # first global kv head this rank reads
    kv_replicas: int     # ranks sharing the same kv head (tp > Hkv)

    @property
    def q_heads(self) -> slice:
        return slice(self.q_head_start, self.q_head_start + self.hq_local)

    @property
    def kv_heads(self) -> slice:
        return slice(self.kv_head_start, self.kv_head_start + self.hkv_local)


def head_shard(hq_total: int, hkv_total: int, tp_size: int, rank: int) -> HeadShard:
    if tp_size < 1 or not 0 <= rank < tp_size:
        raise ValueError(f"bad rank {rank} / tp_size {tp_size}")
    if hq_total % tp_size:
        raise ValueError("num_attention_heads must be divisible by tp_size (llama.py:86-88)")
    hq_local = hq_total // tp_size
    if hkv_total >= tp_size:
        if hkv_total % tp_size:
            raise ValueError("num_key_value_heads must be divisible by tp_size (llama.py:90-94)")
        hkv_local, reps = hkv_total // tp_size, 1
        kv_start = rank * hkv_local
    else:
        if tp_size % hkv_total:
            raise ValueError("tp_size must be divisible by num_key_value_heads (llama.py:95-98)")
        hkv_local, reps = 1, tp_size // hkv_total
        kv_start = rank // reps
    return HeadShard(rank, tp_size, hq_local, hkv_local, rank * hq_local, kv_start, reps)


def shard_decode_inputs(q, k_buf, v_buf, shard: HeadShard):
    """Slices ``q [B, Hq, D]`` and pool buffers ``[P+1, Hkv, D]`` of the *unsharded* model down to one rank."""
    return q[:, shard.q_heads], k_buf[:, shard.kv_heads], v_buf[:, shard.kv_heads]


def max_over_ranks(seconds: float, dist=None, device="cpu") -> float:
    """Job time of a step = the slowest rank's time (all ranks launched concurrently)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return seconds
    import torch

    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
