"""-m gpu randomized parity: seeded random shapes / lengths / prefix groups through the C ABI vs the fp32 oracle.
Complements the hand-picked cases of test_decode_gpu.py / test_extend_gpu.py (tile boundaries, dead rows, ragged
batches, groups around the cascade thresholds); bounded to a few seconds of GPU time."""
import random

import pytest
import torch

from oracle import attn_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import gpu_util as G
    from scratchpad_b200 import _lib

HEADS = [(32, 8), (8, 1), (4, 4), (16, 4), (8, 2), (6, 2), (7, 1), (4, 1)]


def _decode_oracle(c):
    return O.decode_ref(c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"], c["seq_lens"],
                        c["sm_scale"], 0.0, c.get("kv_dtype"), c.get("k_scale"), c.get("v_scale"))


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_decode_with_and_without_shared_prefix(seed):
    rng = random.Random(1000 + seed)
    hq, hkv = rng.choice(HEADS)
    d = rng.choice([64, 128])
    fp8 = rng.random() < 0.5
    kv_dtype = torch.float8_e4m3fn if fp8 else torch.bfloat16
    tol = (G.FP8_ABS, G.FP8_REL) if fp8 else (G.BF16_ABS, G.BF16_REL)
    bs = rng.choice([1, 3, 17, 40, 70])
    shared = rng.choice([0, 0, 130, 1024, 1100, 1290])
    lens = [shared + rng.choice([1, 2, 15, 16, 17, 127, 128, 129, rng.randint(1, 600)]) for _ in range(bs)]
    c = G.random_decode_case(hq, hkv, d, lens, seed=seed, kv_dtype=kv_dtype, shared_prefix=shared)
    got = G.run_decode_case(c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"], c["seq_lens"],
                            c["sm_scale"], 0.0, kv_dtype, c.get("k_scale"), c.get("v_scale"),
                            idx_dtype=rng.choice([torch.int32, torch.int64]), cascade=True)[0]
    G.assert_close(got, _decode_oracle(c), *tol, what=f"fuzz decode seed {seed}: {hq}/{hkv}/{d} fp8={fp8} bs={bs} shared={shared}")


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_extend_both_kernel_families(seed):
    rng = random.Random(2000 + seed)
    hq, hkv = rng.choice(HEADS)
    d = rng.choice([64, 128])
    fp8 = rng.random() < 0.5
    kv_dtype = torch.float8_e4m3fn if fp8 else torch.bfloat16
    tol = (G.FP8_ABS, G.FP8_REL) if fp8 else (G.BF16_ABS, G.BF16_REL)
    bs = rng.choice([1, 2, 4])
    pre = [rng.choice([0, 1, 63, 64, 65, 127, 128, 129, rng.randint(0, 700)]) for _ in range(bs)]
    ext = [rng.choice([1, 2, 31, 64, 65, 128, 130, rng.randint(1, 400)]) for _ in range(bs)]
    c = G.random_extend_case(hq, hkv, d, pre, ext, seed=seed, kv_dtype=kv_dtype)
    want = O.extend_ref(c["q"], c["k_new"], c["v_new"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"],
                        c["seq_lens"], c["extend_seq_lens"], c["extend_start_loc"], c["sm_scale"], 0.0, c.get("kv_dtype"),
                        c.get("k_scale"), c.get("v_scale"))
    for path in ([_lib.PATH_MMA, _lib.PATH_TC] + ([_lib.PATH_TC2] if d == 128 and not fp8 else [])):
        got = G.run_extend_case(c, force_path=path, save_kv_cache=rng.random() < 0.5)[0]
        G.assert_close(got, want, *tol, what=f"fuzz extend seed {seed} path {path}: {hq}/{hkv}/{d} fp8={fp8} pre={pre} ext={ext}")


@pytest.mark.parametrize("kv", ["bf16", "fp8"])
def test_large_logits_decode_extend_cascade(kv):
    """Scores of magnitude ~1e2 (sharp softmax, huge dynamic range): running maxima, lazy rescale, LSE merge of
    split-KV / shared-prefix partials must stay finite and match the oracle."""
    kv_dtype = torch.float8_e4m3fn if kv == "fp8" else torch.bfloat16
    tol = (G.FP8_ABS, G.FP8_REL) if kv == "fp8" else (G.BF16_ABS, G.BF16_REL)
    # decode + cascade
    lens = [1300 + 17 * i for i in range(20)]
    c = G.random_decode_case(32, 8, 128, lens, seed=5, kv_dtype=kv_dtype, shared_prefix=1290)
    c["q"] = (c["q"].float() * 6.0).to(torch.bfloat16)
    got = G.run_decode_case(c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"], c["seq_lens"],
                            c["sm_scale"], 0.0, kv_dtype, c.get("k_scale"), c.get("v_scale"), cascade=True)[0]
    want = O.decode_ref(c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"], c["seq_lens"],
                        c["sm_scale"], 0.0, c.get("kv_dtype"), c.get("k_scale"), c.get("v_scale"))
    G.assert_close(got, want, *tol, what=f"large logits decode {kv}")
    # extend on both kernel families, scores growing along the sequence
    e = G.random_extend_case(16, 4, 128, [300, 0], [200, 333], seed=6, kv_dtype=kv_dtype)
    e["q"] = (e["q"].float() * 5.0).to(torch.bfloat16)
    ramp = torch.linspace(0.5, 2.5, e["k_new"].shape[0]).view(-1, 1, 1)
    e["k_new"] = (e["k_new"].float() * ramp).to(torch.bfloat16)
    want = O.extend_ref(e["q"], e["k_new"], e["v_new"], e["k_buf"], e["v_buf"], e["req_to_token"], e["req_pool_indices"],
                        e["seq_lens"], e["extend_seq_lens"], e["extend_start_loc"], e["sm_scale"], 0.0, e.get("kv_dtype"),
                        e.get("k_scale"), e.get("v_scale"))
    for path in ((_lib.PATH_MMA, _lib.PATH_TC, _lib.PATH_TC2) if kv == "bf16" else (_lib.PATH_MMA, _lib.PATH_TC)):
        got = G.run_extend_case(e, force_path=path)[0]
        G.assert_close(got, want, *tol, what=f"large logits extend {kv} path {path}")
