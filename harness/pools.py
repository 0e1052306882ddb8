"""Stand-ins for the engine's request / slot pools (tests and bench only; not product code).

Condensed restatements of the reference's ``ReqToTokenPool`` (``scratchpad/memory/pool.py:13-72``) and
``TokenToKVPoolAllocator`` (``:189-255``: page size 1, slot 0 reserved for padded tokens).  In deployment the engine's
own classes are used (tests/test_engine_boundary_gpu.py drives the backend through the reference's real ones).
"""

from __future__ import annotations

from typing import List, Optional, Union

import torch


class ReqToTokenPool:
    """Request row -> KV slot ids (``int32 [size, max_context_len]``)."""

    def __init__(self, size: int, max_context_len: int, device: str, use_records: bool = False):
        self.size = size
        self.max_context_len = max_context_len
        self.device = device
        self.req_to_token = torch.zeros((size, max_context_len), dtype=torch.int32, device=device)
        self.free_slots = list(range(size))
        self.write_records = []
        self.use_records = use_records

    def write(self, indices, values):
        self.req_to_token[indices] = values
        if self.use_records:
            self.write_records.append((indices, values))

    def available_size(self):
        return len(self.free_slots)

    def alloc(self, need_size: int) -> Optional[List[int]]:
        if need_size > len(self.free_slots):
            return None
        select_index = self.free_slots[:need_size]
        self.free_slots = self.free_slots[need_size:]
        return select_index

    def free(self, free_index: Union[int, List[int]]):
        if isinstance(free_index, int):
            self.free_slots.append(free_index)
        else:
            self.free_slots.extend(free_index)

    def clear(self):
        self.free_slots = list(range(self.size))
        self.write_records = []


class TokenToKVPoolAllocator:
    """Free list of int64 KV slot ids; slot 0 is the dummy slot for padded tokens."""

    def __init__(self, size: int, dtype: torch.dtype, device: str, kvcache: "MHATokenToKVPool"):
        self.size = size
        self.dtype = dtype
        self.device = device
        self.page_size = 1
        self._kvcache = kvcache
        self.clear()

    def available_size(self):
        return len(self.free_slots)

    def get_kvcache(self):
        return self._kvcache

    def alloc(self, need_size: int):
        if need_size > len(self.free_slots):
            return None
        select_index = self.free_slots[:need_size]
        self.free_slots = self.free_slots[need_size:]
        return select_index

    def free(self, free_index: torch.Tensor):
        if free_index.numel() == 0:
            return
        self.free_slots = torch.cat((self.free_slots, free_index))

    def clear(self):
        self.free_slots = torch.arange(1, self.size + 1, dtype=torch.int64, device=self.device)
