"""-m gpu parity tests of the extend / chunked-prefill path: CUDA kernel (through the Python
backend surface and the C ABI) vs the fp32 oracle, the golden fixtures written from the
reference's own kernel, and the reference's redundant_attention cross-check."""
import pytest
import torch

from oracle import attn_oracle as O
from golden_util import golden_names, load_golden

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import gpu_util as G
    from harness import ForwardMode


def _oracle(c, logit_cap=0.0):
    return O.extend_ref(c["q"], c["k_new"], c["v_new"], c["k_buf"], c["v_buf"], c["req_to_token"],
                        c["req_pool_indices"], c["seq_lens"], c["extend_seq_lens"], c["extend_start_loc"],
                        c["sm_scale"], logit_cap, c.get("kv_dtype"), c.get("k_scale"), c.get("v_scale"))


@pytest.mark.parametrize("name", golden_names("extend"))
def test_extend_matches_reference_golden(name):
    g = load_golden(name)
    c = dict(g, req_pool_indices=g["req_pool_indices"].long(), seq_lens=g["seq_lens"].long())
    got = G.run_extend_case(c, logit_cap=g["logit_cap"], save_kv_cache=False)[0]
    G.assert_close(got, g["out"], G.BF16_ABS, G.BF16_REL, name)


CASES = [
    # hq, hkv, d, prefix_lens, extend_lens
    (8, 2, 64, [7, 0], [5, 9]),
    (32, 8, 128, [130, 0, 64, 300], [1, 70, 129, 33]),          # 8B shapes, ragged, prefix 0, ext 1
    (32, 8, 64, [63, 64, 65], [64, 63, 200]),                    # 1B shapes, tile-boundary prefixes
    (8, 1, 128, [200, 5], [17, 300]),                            # g = 8 (70B TP8 shard)
    (4, 4, 64, [33], [150]),                                     # MHA
    (7, 1, 64, [40, 0], [30, 129]),                              # g = 7: 18 tokens x 7 heads per CTA, 2 dead rows
    (6, 2, 128, [100], [90]),                                    # g = 3
    (32, 8, 128, [1000, 2000, 17, 0], [1, 1, 1, 1]),             # MIXED-mode decode rows (ext_len 1, long prefix)
]


@pytest.mark.parametrize("hq,hkv,d,pre,ext", CASES)
@pytest.mark.parametrize("kv", ["bf16", "fp8"])
def test_extend_vs_oracle(hq, hkv, d, pre, ext, kv):
    kv_dtype = torch.bfloat16 if kv == "bf16" else torch.float8_e4m3fn
    tol = (G.BF16_ABS, G.BF16_REL) if kv == "bf16" else (G.FP8_ABS, G.FP8_REL)
    c = G.random_extend_case(hq, hkv, d, pre, ext, seed=hq * 7 + hkv + d, kv_dtype=kv_dtype)
    got = G.run_extend_case(c)[0]
    G.assert_close(got, _oracle(c), *tol, what=f"{hq}/{hkv}/{d}/{kv}/{pre}/{ext}")


@pytest.mark.parametrize("idx_dtype", [torch.int32, torch.int64])
def test_extend_index_dtypes_logit_cap_mixed_mode(idx_dtype):
    c = G.random_extend_case(16, 4, 128, [50, 0, 129], [40, 65, 1], seed=3)
    got = G.run_extend_case(c, logit_cap=30.0, idx_dtype=idx_dtype, forward_mode=ForwardMode.MIXED)[0]
    G.assert_close(got, _oracle(c, 30.0), G.BF16_ABS, G.BF16_REL)


def test_extend_writes_kv_then_matches_redundant_attention():
    """save_kv_cache=True stores the new tokens (bit-exact) and extend == full causal attention over the
    pool sliced to the new rows (the reference's redundant_attention fixture, extend_attention.py:330-364)."""
    c = G.random_extend_case(32, 8, 64, [64, 0, 33], [100, 31, 257], seed=4)
    got, runner = G.run_extend_case(c, save_kv_cache=True)
    kb, vb = c["k_buf"].clone(), c["v_buf"].clone()
    O.set_kv_buffer_ref(kb, vb, c["out_cache_loc"], c["k_new"], c["v_new"], torch.bfloat16)
    pool = runner.token_to_kv_pool
    assert torch.equal(pool.k_buffer[0].cpu().view(torch.int16), kb.view(torch.int16))
    assert torch.equal(pool.v_buffer[0].cpu().view(torch.int16), vb.view(torch.int16))
    red = O.redundant_attention_ref(c["q"], kb, vb, c["req_to_token"], c["req_pool_indices"], c["seq_lens"],
                                    c["extend_seq_lens"], c["extend_start_loc"], c["sm_scale"])
    G.assert_close(got, red, G.BF16_ABS, G.BF16_REL, "extend vs redundant attention")


def test_extend_long_prefix_chunk_vs_decode_consistency():
    """Size-independent property: the last row of an extend over (prefix P, ext E) equals a decode over the
    P+E tokens once they are in the pool."""
    c = G.random_extend_case(32, 8, 128, [1500], [600], seed=5)
    got, runner = G.run_extend_case(c, save_kv_cache=True)
    kb, vb = c["k_buf"].clone(), c["v_buf"].clone()
    O.set_kv_buffer_ref(kb, vb, c["out_cache_loc"], c["k_new"], c["v_new"], torch.bfloat16)
    dec = G.run_decode_case(c["q"][-1:], kb, vb, c["req_to_token"], c["req_pool_indices"], c["seq_lens"],
                            c["sm_scale"])[0]
    G.assert_close(got[-1:], dec, G.BF16_ABS, G.BF16_REL, "extend last row vs decode")
    G.assert_close(got, _oracle(c), G.BF16_ABS, G.BF16_REL, "extend vs oracle")


# ---- tcgen05 / TMEM path (head_dim 128, bf16 pool), forced so that small cases reach it too ----------------
TC_CASES = [
    (8, 2, 128, [130, 64, 1], [1, 33, 140]),                     # the golden shape family
    (32, 8, 128, [130, 0, 64, 300], [1, 70, 129, 33]),           # 8B shapes, ragged, prefix 0, ext 1
    (8, 1, 128, [200, 5], [17, 300]),                            # g = 8
    (4, 4, 128, [33], [300]),                                    # MHA, several query blocks
    (6, 2, 128, [100, 256], [90, 128]),                          # g = 3 (dead rows), tile-aligned prefix/extend
    (32, 8, 128, [1000, 17], [1, 1]),                            # MIXED-mode decode rows
    (32, 8, 128, [640], [1030]),                                 # many tiles: ring wrap-around, lazy rescale
]


@pytest.mark.parametrize("hq,hkv,d,pre,ext", TC_CASES)
@pytest.mark.parametrize("kv", ["bf16", "fp8"])
def test_extend_tcgen05_path_vs_oracle(hq, hkv, d, pre, ext, kv):
    from scratchpad_b200 import _lib
    kv_dtype = torch.bfloat16 if kv == "bf16" else torch.float8_e4m3fn
    tol = (G.BF16_ABS, G.BF16_REL) if kv == "bf16" else (G.FP8_ABS, G.FP8_REL)
    c = G.random_extend_case(hq, hkv, d, pre, ext, seed=hq * 11 + hkv + len(pre), kv_dtype=kv_dtype)
    ref_path = G.run_extend_case(c, force_path=_lib.PATH_MMA)[0]
    # PATH_TC: the one-CTA tcgen05 kernel; PATH_TC2: CTA pairs (cta_group::2), bf16 pools at head_dim 128
    for path in ((_lib.PATH_TC, _lib.PATH_TC2) if kv == "bf16" else (_lib.PATH_TC,)):
        got = G.run_extend_case(c, force_path=path)[0]
        G.assert_close(got, _oracle(c), *tol, what=f"tc path {path} {hq}/{hkv}/{d}/{kv}/{pre}/{ext}")
        G.assert_close(got, ref_path, *tol, what=f"tcgen05 path {path} vs mma.sync path")


def test_extend_tcgen05_fp8_pool_per_tensor_scale_and_no_prefix():
    """fp8-e4m3 pool on the tcgen05 path: per-tensor scales, a pure prefill (every key is a new bf16 token that is
    rescaled by 1/scale on its way in) and a prefix|new boundary that is not a multiple of 16."""
    from scratchpad_b200 import _lib
    c = G.random_extend_case(32, 8, 128, [0, 77, 519], [260, 130, 9], seed=21, kv_dtype=torch.float8_e4m3fn, per_head=False)
    got = G.run_extend_case(c, force_path=_lib.PATH_TC, save_kv_cache=True)[0]
    G.assert_close(got, _oracle(c), G.FP8_ABS, G.FP8_REL, what="tc fp8 per-tensor")


def test_extend_tcgen05_golden_and_logit_cap():
    from scratchpad_b200 import _lib
    g = load_golden("extend_gqa4_d128")
    c = dict(g, req_pool_indices=g["req_pool_indices"].long(), seq_lens=g["seq_lens"].long())
    c2 = G.random_extend_case(16, 4, 128, [50, 0, 129], [40, 65, 1], seed=3)
    for path in (_lib.PATH_TC, _lib.PATH_TC2):
        got = G.run_extend_case(c, logit_cap=g["logit_cap"], save_kv_cache=False, force_path=path)[0]
        G.assert_close(got, g["out"], G.BF16_ABS, G.BF16_REL, f"tc golden path {path}")
        got = G.run_extend_case(c2, logit_cap=30.0, idx_dtype=torch.int32, force_path=path)[0]
        G.assert_close(got, _oracle(c2, 30.0), G.BF16_ABS, G.BF16_REL, f"tc logit cap path {path}")


def test_extend_tcgen05_large_outliers_trigger_rescale():
    """Scores grow along the sequence so that row maxima jump by more than the lazy-rescale threshold."""
    from scratchpad_b200 import _lib
    c = G.random_extend_case(8, 2, 128, [512], [256], seed=9)
    ramp = torch.linspace(0.2, 3.0, c["k_new"].shape[0]).view(-1, 1, 1)
    c["k_new"] = (c["k_new"].float() * ramp).to(torch.bfloat16)
    c["q"] = (c["q"].float() * 2.0).to(torch.bfloat16)
    for path in (_lib.PATH_TC, _lib.PATH_TC2):
        got = G.run_extend_case(c, force_path=path)[0]
        G.assert_close(got, _oracle(c), G.BF16_ABS, G.BF16_REL, f"tc rescale path {path}")


TC64_CASES = [
    (32, 8, 64, [63, 64, 65], [64, 63, 200]),       # Llama-3.2-1B shapes, tile-boundary prefixes
    (8, 2, 64, [7, 0, 300], [5, 190, 140]),
    (4, 4, 64, [33], [300]),                        # MHA
    (32, 8, 64, [700], [900]),                      # many tiles: ring wrap-around
]


@pytest.mark.parametrize("hq,hkv,d,pre,ext", TC64_CASES)
def test_extend_tcgen05_head_dim_64(hq, hkv, d, pre, ext):
    from scratchpad_b200 import _lib
    c = G.random_extend_case(hq, hkv, d, pre, ext, seed=hq * 13 + hkv + len(pre))
    got = G.run_extend_case(c, force_path=_lib.PATH_TC)[0]
    G.assert_close(got, _oracle(c), G.BF16_ABS, G.BF16_REL, what=f"tc d64 {hq}/{hkv}/{pre}/{ext}")
    # head_dim 64 over an fp8-e4m3 pool (round 2): raw 64-byte rows widened to f16 on chip, per-head and per-tensor scales
    for per_head in (True, False):
        c8 = G.random_extend_case(hq, hkv, d, pre, ext, seed=1, kv_dtype=torch.float8_e4m3fn, per_head=per_head)
        got8 = G.run_extend_case(c8, force_path=_lib.PATH_TC)[0]
        G.assert_close(got8, _oracle(c8), G.FP8_ABS, G.FP8_REL, what=f"tc d64 fp8 {hq}/{hkv}/{pre}/{ext}")
        G.assert_close(got8, G.run_extend_case(c8, force_path=_lib.PATH_MMA)[0], G.FP8_ABS, G.FP8_REL, what="tc d64 fp8 vs mma.sync")


def test_extend_tcgen05_tma_tiles_new_tokens_and_consecutive_prefix():
    """The tcgen05 extend kernel fetches a K / V tile by ONE TMA box copy when its 128 keys are 128 consecutive rows:
    tiles of new tokens always (rows of k_new / v_new), cached-prefix tiles when the slots are a run (unfragmented
    token-granular allocator, or pages of >= 128 tokens) -- checked per tile on device.  Same oracle; the debug counter
    proves which path ran."""
    from scratchpad_b200 import _lib
    lib = _lib.load()
    hq, hkv, d = 8, 2, 128
    pre, ext = [640, 130], [300, 520]

    def run(c):
        lib.b200_debug_extend_tma_tiles(1)
        got = G.run_extend_case(c, force_path=_lib.PATH_TC)[0]
        n = lib.b200_debug_extend_tma_tiles(1)
        G.assert_close(got, _oracle(c), G.BF16_ABS, G.BF16_REL, "tc extend with TMA tiles")
        return n

    n_random = run(G.random_extend_case(hq, hkv, d, pre, ext, seed=31))                       # new-token tiles only
    n_contig = run(G.random_extend_case(hq, hkv, d, pre, ext, seed=32, mode="contiguous"))    # + prefix runs
    n_paged = run(G.random_extend_case(hq, hkv, d, pre, ext, seed=33, page_size=128))         # + 128-token pages
    assert n_random > 0, "tiles of new tokens must arrive by TMA"
    assert n_contig > n_random and n_paged > n_random, (n_random, n_contig, n_paged)
    # fp8 pools widen their rows on chip: row gather only
    lib.b200_debug_extend_tma_tiles(1)
    c8 = G.random_extend_case(hq, hkv, d, pre, ext, seed=34, kv_dtype=torch.float8_e4m3fn, mode="contiguous")
    got = G.run_extend_case(c8, force_path=_lib.PATH_TC)[0]
    G.assert_close(got, _oracle(c8), G.FP8_ABS, G.FP8_REL, "tc extend fp8")
    assert lib.b200_debug_extend_tma_tiles(1) == 0
