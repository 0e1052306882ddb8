"""Stand-ins for the engine around the attention backend -- scheduler-side pools, ``ForwardBatch`` / ``ForwardMode``,
``RadixAttention`` and synthetic batches of the BASELINE.json shapes.  Used by ``tests/``, ``bench.py`` and
``__graft_entry__.smoke()`` only; the product package ``scratchpad_b200`` never imports it, and in deployment the
engine's own classes take their place (``tests/test_engine_boundary_gpu.py`` runs the backend under the reference's
real ones)."""

from .forward_info import ForwardBatch, ForwardMode, compute_extend_start_loc
from .pools import ReqToTokenPool, TokenToKVPoolAllocator
from .radix_attention import RadixAttention

__all__ = ["ForwardBatch", "ForwardMode", "compute_extend_start_loc", "ReqToTokenPool", "TokenToKVPoolAllocator",
           "RadixAttention"]
