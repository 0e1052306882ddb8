"""Debug aid: which rows of an extend case differ between the CTA-pair kernel and the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import gpu_util as G
from oracle import attn_oracle as O
from scratchpad_b200 import _lib

CASES = [
    (8, 2, 128, [130, 64, 1], [1, 33, 140]),
    (32, 8, 128, [130, 0, 64, 300], [1, 70, 129, 33]),
    (8, 1, 128, [200, 5], [17, 300]),
    (4, 4, 128, [33], [300]),
    (6, 2, 128, [100, 256], [90, 128]),
    (32, 8, 128, [1000, 17], [1, 1]),
    (32, 8, 128, [640], [1030]),
    (8, 1, 128, [0], [64]),
    (8, 1, 128, [0], [300]),
    (8, 1, 128, [256], [32]),
]
for hq, hkv, d, pre, ext in CASES:
    c = G.random_extend_case(hq, hkv, d, pre, ext, seed=hq * 11 + hkv + len(pre))
    want = O.extend_ref(c["q"], c["k_new"], c["v_new"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"],
                        c["seq_lens"], c["extend_seq_lens"], c["extend_start_loc"], c["sm_scale"], 0.0)
    for path in (_lib.PATH_TC2, _lib.PATH_TC):
        got = G.run_extend_case(c, force_path=path)[0].float()
        err = (got - want).abs().amax(dim=-1)  # [T, hq]
        bad = (err > 0.02).nonzero()
        print(f"case {hq}/{hkv} pre={pre} ext={ext} path={path}: max {err.max():.4f} bad rows {len(bad)} of {err.numel()}")
        if len(bad) and path == _lib.PATH_TC2:
            g = hq // hkv
            start = 0
            for i, e in enumerate(ext):
                sel = bad[(bad[:, 0] >= start) & (bad[:, 0] < start + e)]
                if len(sel):
                    toks = sorted(set((sel[:, 0] - start).tolist()))
                    heads = sorted(set(sel[:, 1].tolist()))
                    bm_t = 256 // g
                    print(f"   req {i} (pre {pre[i]}, ext {e}): bad tokens {toks[:40]}{'...' if len(toks) > 40 else ''} heads {heads}")
                    print(f"      blocks {sorted(set(t // bm_t for t in toks))}, in-block tokens {sorted(set(t % bm_t for t in toks))[:40]}")
                start += e
