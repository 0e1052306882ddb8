"""scratchpad_b200 -- B200-native paged decode/extend attention for eth-easl/Scratchpad.

One hot path only: the `cuda_b200` AttentionBackend and the C-ABI kernel
library behind it.  See DESIGN.md / INTEGRATION.md.
"""

from .backend import AttentionBackend, CudaB200AttnBackend
from .pool import MHATokenToKVPool, PagedTokenToKVPoolAllocator

__all__ = [
    "AttentionBackend",
    "CudaB200AttnBackend",
    "MHATokenToKVPool",
    "PagedTokenToKVPoolAllocator",
]
