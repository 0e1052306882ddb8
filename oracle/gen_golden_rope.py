"""Generate tests/golden/rope_*.npz by running the REFERENCE'S OWN RotaryEmbedding.forward_native.

TEST INFRASTRUCTURE ONLY.  Imports /root/reference/scratchpad/nn/layers/rotary_embedding.py by path (its only
engine import, ``scratchpad.model_executor.custom_op.CustomOp``, is stubbed by a plain nn.Module with the same
dispatch shape), builds the module the way ``get_rope`` does for Llama-3.x (``rotary_embedding.py:918-975``) and
stores seeded inputs + outputs.  Runs only in the build container; the fixtures are committed.

    python oracle/gen_golden_rope.py
"""

from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference/scratchpad/nn/layers/rotary_embedding.py"
OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_reference_rope():
    class CustomOp(nn.Module):  # scratchpad/model_executor/custom_op.py:5-33 without the platform imports
        def __init__(self, *a, **k):
            super().__init__()
            self._forward_method = self.forward_native

        def forward(self, *a, **k):
            return self._forward_method(*a, **k)

    for name in ("scratchpad", "scratchpad.model_executor", "scratchpad.model_executor.custom_op"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["scratchpad.model_executor.custom_op"].CustomOp = CustomOp
    spec = importlib.util.spec_from_file_location("ref_rotary_embedding", REF)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def bits(x: torch.Tensor) -> np.ndarray:
    return x.contiguous().view(torch.int16).numpy().astype(np.uint16)


LLAMA3_SCALING = {"rope_type": "llama3", "factor": 32.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                  "original_max_position_embeddings": 8192}

# name: (head_size, rotary_dim, max_position, base, is_neox, rope_scaling, hq, hkv, tokens); "ropeh_*" = fp16 model dtype
CASES = {
    "rope_llama1b_neox": (64, 64, 2048, 500000.0, True, LLAMA3_SCALING, 32, 8, 24),
    "rope_llama8b_neox": (128, 128, 2048, 500000.0, True, LLAMA3_SCALING, 8, 2, 17),
    "rope_plain_neox": (128, 128, 2048, 10000.0, True, None, 4, 1, 9),
    "rope_gptj": (64, 64, 2048, 10000.0, False, None, 4, 2, 11),
    "rope_partial_neox": (128, 64, 2048, 10000.0, True, None, 4, 2, 7),
    "rope_partial_gptj": (64, 32, 2048, 10000.0, False, None, 2, 2, 5),
    "ropeh_llama8b_neox": (128, 128, 2048, 500000.0, True, LLAMA3_SCALING, 8, 2, 17),
    "ropeh_llama1b_neox": (64, 64, 2048, 500000.0, True, LLAMA3_SCALING, 32, 8, 24),
    "ropeh_gptj": (64, 64, 2048, 10000.0, False, None, 4, 2, 11),
    "ropeh_partial_neox": (128, 64, 2048, 10000.0, True, None, 4, 2, 7),
}


def main():
    os.makedirs(OUT_DIR, exist_ok=True)
    ref = load_reference_rope()
    for i, (name, (hs, rd, maxpos, base, neox, scaling, hq, hkv, t)) in enumerate(CASES.items()):
        gen = torch.Generator().manual_seed(3000 + i)
        dtype = torch.float16 if name.startswith("ropeh_") else torch.bfloat16
        mod = ref.get_rope(hs, rotary_dim=rd, max_position=maxpos, base=base, is_neox_style=neox,
                           rope_scaling=scaling, dtype=dtype)
        positions = torch.randint(0, maxpos, (t,), generator=gen, dtype=torch.int64)
        positions[0], positions[-1] = 0, maxpos - 1
        q = torch.randn((t, hq * hs), generator=gen).to(dtype)
        k = (torch.randn((t, hkv * hs), generator=gen) * 3).to(dtype)
        q_out, k_out = mod.forward_native(positions, q.clone(), k.clone())
        np.savez_compressed(
            os.path.join(OUT_DIR, name + ".npz"), kind="rope", dtype="f16" if dtype == torch.float16 else "bf16", head_size=hs, rotary_dim=rd, max_position=maxpos,
            base=np.float64(base), is_neox=int(neox), llama3=int(scaling is not None), hq=hq, hkv=hkv,
            positions=positions.numpy(), q=bits(q), k=bits(k), cos_sin_cache=bits(mod.cos_sin_cache),
            q_out=bits(q_out), k_out=bits(k_out),
        )
        print(name, type(mod).__name__, tuple(q_out.shape), tuple(mod.cos_sin_cache.shape))


if __name__ == "__main__":
    main()
