"""Generate tests/golden/*.npz by running the REFERENCE'S OWN in-tree kernels.

TEST INFRASTRUCTURE ONLY (see oracle/attn_oracle.py header).

The reference (/root/reference, eth-easl/Scratchpad v0.1.6) is pure Python; its
attention math for this path lives in
``scratchpad/nn/attention/triton_attn/{decode,extend}_attention.py``.  Those
files load by path and execute on CPU under ``TRITON_INTERPRET=1`` with fp32
tensors.  This script feeds them
seeded bf16-valued inputs and stores inputs + outputs as small fixtures, so
the oracle (and through it the CUDA path) is pinned to outputs of the
reference itself.  It only runs in the build container -- /root/reference is
not present on the GPU box -- hence the committed fixtures.

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz
"""

from __future__ import annotations

import os
import sys
import types
import importlib.util

os.environ["TRITON_INTERPRET"] = "1"  # before importing triton

import numpy as np
import torch

REF_DIR = "/root/reference/scratchpad/nn/attention/triton_attn"
OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_reference_kernels():
    # extend_attention.py:7 / prefill_attention.py:5 query the device at import
    torch.cuda.get_device_capability = lambda *a, **k: (8, 0)
    pkg = types.ModuleType("ref_triton_attn")
    pkg.__path__ = [REF_DIR]
    sys.modules["ref_triton_attn"] = pkg

    def load(n):
        s = importlib.util.spec_from_file_location(f"ref_triton_attn.{n}", f"{REF_DIR}/{n}.py")
        m = importlib.util.module_from_spec(s)
        sys.modules[s.name] = m
        s.loader.exec_module(m)
        return m

    load("prefill_attention")
    return load("decode_attention"), load("extend_attention")


def bf16_valued(shape, gen, std=1.0):
    return (torch.randn(shape, generator=gen) * std).to(torch.bfloat16).to(torch.float32)


def bf16_bits(x: torch.Tensor) -> np.ndarray:
    return x.to(torch.bfloat16).view(torch.int16).numpy().astype(np.uint16)


def build_state(gen, n_req_rows, max_ctx, pool, hkv, d, dv, seq_lens, shared_prefix=0):
    """Pool + page table the way the engine lays them out (pool.py:16-22,298-317):
    slot 0 reserved, random slot ids, stale garbage beyond seq_len, optional
    RadixAttention-style shared prefix (equal ids in different rows)."""
    b = len(seq_lens)
    k_buf = bf16_valued((pool + 1, hkv, d), gen)
    v_buf = bf16_valued((pool + 1, hkv, dv), gen)
    k_buf[0].zero_(); v_buf[0].zero_()
    req_to_token = torch.randint(0, pool + 1, (n_req_rows, max_ctx), generator=gen, dtype=torch.int32)  # stale
    rows = torch.randperm(n_req_rows, generator=gen)[:b]
    perm = (torch.randperm(pool, generator=gen) + 1).to(torch.int32)
    shared = perm[:shared_prefix]
    cur = shared_prefix
    for i, n in enumerate(seq_lens):
        if shared_prefix and n >= shared_prefix:
            req_to_token[rows[i], :shared_prefix] = shared
            rest = n - shared_prefix
            req_to_token[rows[i], shared_prefix:n] = perm[cur:cur + rest]
            cur += rest
        else:
            req_to_token[rows[i], :n] = perm[cur:cur + n]
            cur += n
    assert cur <= pool
    return k_buf, v_buf, req_to_token, rows.to(torch.int32)


DECODE_CASES = {
    # name: (hq, hkv, d, seq_lens, pool, shared_prefix, logit_cap)
    "decode_gqa4_d64":        (8, 2, 64, [5, 17, 33], 128, 0, 0.0),
    "decode_mha_d64":         (4, 4, 64, [1, 40], 96, 0, 0.0),
    "decode_gqa8_d128":       (8, 1, 128, [130, 7, 64], 256, 0, 0.0),
    "decode_shared_cap":      (8, 2, 64, [70, 70, 95, 3], 256, 64, 30.0),
    "decode_llama1b_small":   (32, 8, 64, [300, 1, 129, 64, 257, 16, 200, 33], 1100, 0, 0.0),
}

EXTEND_CASES = {
    # name: (hq, hkv, d, prefix_lens, extend_lens, pool, logit_cap)
    "extend_small":           (4, 2, 64, [7, 0], [5, 9], 64, 0.0),
    "extend_gqa4_d128":       (8, 2, 128, [130, 64, 1], [1, 33, 140], 512, 0.0),
    "extend_gqa8_cap":        (8, 1, 64, [0, 40, 200], [70, 3, 17], 400, 50.0),
}


def gen_decode(dec, name, cfg, seed):
    hq, hkv, d, seq_lens, pool, shared, cap = cfg
    gen = torch.Generator().manual_seed(seed)
    max_ctx = max(seq_lens) + 4
    k_buf, v_buf, r2t, rows = build_state(gen, len(seq_lens) + 3, max_ctx, pool, hkv, d, d, seq_lens, shared)
    b = len(seq_lens)
    q = bf16_valued((b, hq, d), gen)
    sl = torch.tensor(seq_lens, dtype=torch.int32)
    start = torch.zeros(b, dtype=torch.int32); start[1:] = torch.cumsum(sl[:-1], 0)
    o = torch.zeros((b, hq, d), dtype=torch.float32)
    logits = torch.empty((hq, int(sl.sum())), dtype=torch.float32)
    sm_scale = 1.0 / d ** 0.5
    # decode_attention.py:547-560 (older non-split signature; the in-tree kernel)
    dec.decode_attention_fwd(q, k_buf, v_buf, o, r2t, rows, start, sl, logits, int(sl.max()), sm_scale, cap)
    np.savez_compressed(
        os.path.join(OUT_DIR, name + ".npz"), kind="decode", seed=seed,
        q=bf16_bits(q), k_buf=bf16_bits(k_buf), v_buf=bf16_bits(v_buf), req_to_token=r2t.numpy(),
        req_pool_indices=rows.numpy(), seq_lens=sl.numpy(), sm_scale=np.float32(sm_scale),
        logit_cap=np.float32(cap), out=o.numpy(),
    )
    return o


def gen_extend(ext, name, cfg, seed):
    hq, hkv, d, pre_lens, ext_lens, pool, cap = cfg
    gen = torch.Generator().manual_seed(seed)
    seq_lens = [p + e for p, e in zip(pre_lens, ext_lens)]
    max_ctx = max(seq_lens) + 4
    k_buf, v_buf, r2t, rows = build_state(gen, len(seq_lens) + 2, max_ctx, pool, hkv, d, d, seq_lens)
    t = sum(ext_lens)
    q = bf16_valued((t, hq, d), gen)
    k_new = bf16_valued((t, hkv, d), gen)
    v_new = bf16_valued((t, hkv, d), gen)
    sl = torch.tensor(seq_lens, dtype=torch.int32)
    el = torch.tensor(ext_lens, dtype=torch.int32)
    st = torch.zeros(len(seq_lens), dtype=torch.int32); st[1:] = torch.cumsum(el[:-1], 0)
    # triton_backend.py:133-136: the backend writes the new KV into the pool first
    for i in range(len(seq_lens)):
        loc = r2t[rows[i], pre_lens[i]:seq_lens[i]].long()
        k_buf[loc] = k_new[st[i]:st[i] + el[i]]
        v_buf[loc] = v_new[st[i]:st[i] + el[i]]
    o = torch.zeros((t, hq, d), dtype=torch.float32)
    sm_scale = 1.0 / d ** 0.5
    ext.extend_attention_fwd(q, k_new, v_new, o, k_buf, v_buf, r2t, rows, sl, el, st, int(el.max()), sm_scale, cap)
    np.savez_compressed(
        os.path.join(OUT_DIR, name + ".npz"), kind="extend", seed=seed,
        q=bf16_bits(q), k_new=bf16_bits(k_new), v_new=bf16_bits(v_new), k_buf=bf16_bits(k_buf),
        v_buf=bf16_bits(v_buf), req_to_token=r2t.numpy(), req_pool_indices=rows.numpy(), seq_lens=sl.numpy(),
        extend_seq_lens=el.numpy(), extend_start_loc=st.numpy(), sm_scale=np.float32(sm_scale),
        logit_cap=np.float32(cap), out=o.numpy(),
    )
    return o


def main():
    os.makedirs(OUT_DIR, exist_ok=True)
    dec, ext = load_reference_kernels()
    for i, (name, cfg) in enumerate(DECODE_CASES.items()):
        o = gen_decode(dec, name, cfg, 1000 + i)
        print(name, tuple(o.shape), float(o.abs().max()))
    for i, (name, cfg) in enumerate(EXTEND_CASES.items()):
        o = gen_extend(ext, name, cfg, 2000 + i)
        print(name, tuple(o.shape), float(o.abs().max()))


if __name__ == "__main__":
    main()
