"""-m gpu parity tests of the fp16 model dtype (``model_runner.dtype == torch.float16``: what the engine serves fp16
and, by default, fp32 checkpoints in -- config/utils.py:41-43 -- with the KV pool in fp16, kv_cache_dtype "auto",
model_runner.py:366-367, or in fp8-e4m3).  Every path of the bf16 suite once more with fp16 q / k / v / o: split-KV
decode (row gather, TMA tiles, paged pools, CUDA-graph replay), the shared-prefix cascade, both extend kernel
families, the KV store and the fused rotary store -- against the fp32 oracle, the golden fixtures written from the
reference's own kernels, the library's naive kernel and the reference's own GPU path (FlashInfer, Triton)."""
import pytest
import torch

from oracle import attn_oracle as O
from oracle import rope_oracle as R
from golden_util import golden_names, load_golden
from scratchpad_b200 import _lib

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import gpu_util as G
    from harness import ForwardBatch, ForwardMode, checks, synth
    from scratchpad_b200 import CudaB200AttnBackend

F16, FP8 = torch.float16, torch.float8_e4m3fn


def _tol(kv_dtype):
    return (G.F16_ABS, G.F16_REL) if kv_dtype == F16 else (G.FP8_ABS, G.FP8_REL)


def _dec_oracle(c, logit_cap=0.0):
    return O.decode_ref(c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"], c["seq_lens"],
                        c["sm_scale"], logit_cap, c.get("kv_dtype"), c.get("k_scale"), c.get("v_scale"))


def _dec_run(c, logit_cap=0.0, **kw):
    return G.run_decode_case(c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"],
                             c["seq_lens"], c["sm_scale"], logit_cap, c["kv_dtype"], c.get("k_scale"), c.get("v_scale"), **kw)


def _ext_oracle(c, logit_cap=0.0):
    return O.extend_ref(c["q"], c["k_new"], c["v_new"], c["k_buf"], c["v_buf"], c["req_to_token"],
                        c["req_pool_indices"], c["seq_lens"], c["extend_seq_lens"], c["extend_start_loc"],
                        c["sm_scale"], logit_cap, c.get("kv_dtype"), c.get("k_scale"), c.get("v_scale"))


# ---------------------------------------------------------------------------------------------------------------
# decode
# ---------------------------------------------------------------------------------------------------------------
DECODE_SHAPES = [
    # hq, hkv, d -> kernel variant (see tests/test_decode_gpu.py)
    (32, 8, 64),    # Llama-3.2-1B heads: 8 kv heads per CTA
    (32, 8, 128),   # 8B TP1: fp16 -> HPC 4 x 2 head groups; fp8 -> HPC 8
    (8, 1, 128),    # 70B TP8 shard, g = 8, token sub-slices combined inside the CTA
    (4, 1, 64),     # 1B TP8 shard
    (14, 2, 64),    # g = 7 (Qwen2-0.5B heads: an fp16-era family)
    (16, 1, 128),   # g = 16: two q tiles per kv head
    (4, 4, 128),    # MHA
]
LENS = [[1], [15, 16, 17], [127, 128, 129, 1, 300], [513, 64, 2, 1000]]


@pytest.mark.parametrize("hq,hkv,d", DECODE_SHAPES)
@pytest.mark.parametrize("kv", ["fp16", "fp8"])
def test_fp16_decode_vs_oracle_shapes(hq, hkv, d, kv):
    kv_dtype = F16 if kv == "fp16" else FP8
    for i, lens in enumerate(LENS):
        c = G.random_decode_case(hq, hkv, d, lens, seed=50 * hq + 10 * hkv + i, kv_dtype=kv_dtype, act_dtype=F16)
        got, backend, runner, fb, layer = _dec_run(c)
        assert got.dtype == F16 and backend.act_dtype == _lib.ACT_F16
        G.assert_close(got, _dec_oracle(c), *_tol(kv_dtype), what=f"fp16 decode {hq}/{hkv}/{d}/{kv}/{lens}")


@pytest.mark.parametrize("name", golden_names("decode"))
def test_fp16_decode_matches_reference_golden(name):
    """The fixtures written from the reference's own decode kernel; their bf16-valued inputs convert to fp16 without
    loss (8 significand bits; only values below 2^-17 land between fp16 subnormals), so the same `out` is the target."""
    g = load_golden(name)
    q, kb, vb = g["q"].to(F16), g["k_buf"].to(F16), g["v_buf"].to(F16)
    assert (q.float() - g["q"].float()).abs().max() < 1e-7 and (kb.float() - g["k_buf"].float()).abs().max() < 1e-7
    got = G.run_decode_case(q, kb, vb, g["req_to_token"], g["req_pool_indices"].long(), g["seq_lens"].long(),
                            g["sm_scale"], g["logit_cap"], kv_dtype=F16)[0]
    G.assert_close(got, g["out"], G.F16_ABS, G.F16_REL, name)


def test_fp16_decode_int32_indices_logit_cap_qkv_slice_and_naive_kernel():
    c = G.random_decode_case(32, 8, 128, [300, 1, 129, 64, 1000], seed=7, kv_dtype=F16)
    got, backend, runner, fb, layer = _dec_run(c, logit_cap=30.0, idx_dtype=torch.int32, q_slice_of_qkv=True)
    G.assert_close(got, _dec_oracle(c, 30.0), G.F16_ABS, G.F16_REL, "fp16 decode, cap, int32, q slice")
    # the library's one-warp-per-(request, head) fp32 kernel on the same device buffers (fp16 in, fp16 out)
    pool = runner.token_to_kv_pool
    naive = checks.naive_decode(c["q"].cuda(), pool.k_buffer[0], pool.v_buffer[0], runner.req_to_token_pool.req_to_token,
                                fb.req_pool_indices, fb.seq_lens, c["sm_scale"], 30.0, _lib.KV_F16, None, None)
    torch.cuda.synchronize()
    assert naive.dtype == F16
    G.assert_close(got, naive.cpu(), G.F16_ABS, G.F16_REL, "fp16 decode vs naive kernel")


@pytest.mark.parametrize("kv", ["fp16", "fp8"])
@pytest.mark.parametrize("page_size,mode", [(1, "contiguous"), (16, "random"), (64, "random")])
def test_fp16_decode_tma_tiles_and_paged_pools(kv, page_size, mode):
    """TMA box copies of fp16 tiles: consecutive slots of the token-granular pool (PAGE_AUTO) and 16 / 64-token
    pages; the tile counter proves the TMA branch ran, and the result is bit-identical to the pure row gather's."""
    kv_dtype = F16 if kv == "fp16" else FP8
    lib = _lib.load()
    lens = [700, 33, 256, 1025, 16]
    c = G.random_decode_case(32, 8, 128, lens, seed=91, kv_dtype=kv_dtype, act_dtype=F16, mode=mode, page_size=page_size)
    lib.b200_debug_tma_tiles(1)
    got = _dec_run(c, page_size=page_size)[0]
    n = lib.b200_debug_tma_tiles(1)
    assert n > 0, "no tile took the TMA branch"
    G.assert_close(got, _dec_oracle(c), *_tol(kv_dtype), what=f"fp16 decode TMA ps={page_size} {kv}")
    plain = _dec_run(c, page_size=page_size, tma_contiguous_runs=False)[0] if page_size == 1 else None
    if plain is not None:
        assert lib.b200_debug_tma_tiles(1) == 0
        assert torch.equal(got.view(torch.int16), plain.view(torch.int16)), "TMA tiles and row gather disagree"


@pytest.mark.parametrize("kv", ["fp16", "fp8"])
@pytest.mark.parametrize("hq,hkv,d,bs", [(32, 8, 128, 24), (32, 8, 64, 40), (8, 1, 128, 16)])
def test_fp16_decode_shared_prefix_cascade(kv, hq, hkv, d, bs):
    """The dense tcgen05 leg (f16 x f16 MMAs on the pool bytes as they are / widened from e4m3) + suffix split-KV +
    merge, against the oracle and against the same backend with the cascade off."""
    kv_dtype = F16 if kv == "fp16" else FP8
    gen = torch.Generator().manual_seed(5)
    lens = [1280 + int(x) for x in torch.randint(1, 200, (bs,), generator=gen)]
    c = G.random_decode_case(hq, hkv, d, lens, seed=17, kv_dtype=kv_dtype, act_dtype=F16, shared_prefix=1280)
    on, backend_on = _dec_run(c, cascade=True)[:2]
    off, backend_off = _dec_run(c, cascade=False)[:2]
    shared_len = lambda b: int(b.workspace[:64].view(torch.int32)[7].item())  # plan header: what the dense leg takes
    assert shared_len(backend_on) == 1280 and shared_len(backend_off) == 0
    want = _dec_oracle(c)
    G.assert_close(on, want, *_tol(kv_dtype), what=f"fp16 cascade on {kv}")
    G.assert_close(off, want, *_tol(kv_dtype), what=f"fp16 cascade off {kv}")
    G.assert_close(on, off, *_tol(kv_dtype), what="cascade on vs off")


def test_fp16_decode_cuda_graph_replay():
    """Capture at a padded batch size with static int32 buffers, replay with other lengths after the replay hook
    (cuda_graph_runner.py:394-423,443-510), fp16 buffers."""
    hq, hkv, d, bs = 32, 8, 128, 8
    lens_a = [100, 200, 17, 64, 1, 1, 1, 1]
    lens_b = [31, 500, 333, 128, 250, 1, 1, 1]
    c = G.random_decode_case(hq, hkv, d, [600] * bs, seed=41, kv_dtype=F16)
    runner = G.make_runner_from_case(c["k_buf"], c["v_buf"], c["req_to_token"], F16, tp_hq=hq, dtype=F16)
    backend = CudaB200AttnBackend(runner)
    layer = synth.make_layer(runner)
    backend.init_cuda_graph_state(16)
    s_rpi = torch.zeros(16, dtype=torch.int32, device="cuda")
    s_sl = torch.full((16,), backend.get_cuda_graph_seq_len_fill_value(), dtype=torch.int32, device="cuda")
    s_q = torch.zeros((bs, hq * d), dtype=F16, device="cuda")
    fb = ForwardBatch(ForwardMode.DECODE, bs, s_rpi[:bs], s_sl[:bs], torch.zeros(bs, dtype=torch.int64, device="cuda"),
                      req_to_token_pool=runner.req_to_token_pool, token_to_kv_pool=runner.token_to_kv_pool,
                      attn_backend=backend)
    backend.init_forward_metadata_capture_cuda_graph(bs, bs, s_rpi, s_sl, None, ForwardMode.DECODE, None)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            layer(s_q, None, None, fb, save_kv_cache=False)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        s_o = layer(s_q, None, None, fb, save_kv_cache=False)
    assert s_o.dtype == F16
    for lens in (lens_a, lens_b):
        s_rpi[:bs].copy_(c["req_pool_indices"].to(torch.int32))
        s_sl[:bs].copy_(torch.tensor(lens, dtype=torch.int32))
        s_q.copy_(c["q"].reshape(bs, -1))
        backend.init_forward_metadata_replay_cuda_graph(bs, s_rpi, s_sl, sum(lens), None, ForwardMode.DECODE, None,
                                                        seq_lens_cpu=None)
        graph.replay()
        torch.cuda.synchronize()
        cc = dict(c, seq_lens=torch.tensor(lens))
        G.assert_close(s_o.view(bs, hq, d).cpu(), _dec_oracle(cc), G.F16_ABS, G.F16_REL, what=f"fp16 graph replay {lens}")


# ---------------------------------------------------------------------------------------------------------------
# extend
# ---------------------------------------------------------------------------------------------------------------
EXT_CASES = [
    # hq, hkv, d, prefix_lens, extend_lens
    (8, 2, 64, [7, 0], [5, 9]),
    (32, 8, 128, [130, 0, 64, 300], [1, 70, 129, 33]),
    (32, 8, 64, [63, 64, 65], [64, 63, 200]),
    (14, 2, 64, [40, 0], [30, 129]),                  # g = 7: dead rows
    (32, 8, 128, [1000, 17], [1, 1]),                 # MIXED-mode decode rows
]


@pytest.mark.parametrize("hq,hkv,d,pre,ext", EXT_CASES)
@pytest.mark.parametrize("kv", ["fp16", "fp8"])
def test_fp16_extend_mma_path_vs_oracle(hq, hkv, d, pre, ext, kv):
    kv_dtype = F16 if kv == "fp16" else FP8
    c = G.random_extend_case(hq, hkv, d, pre, ext, seed=hq * 5 + hkv + d, kv_dtype=kv_dtype, act_dtype=F16)
    got = G.run_extend_case(c, force_path=_lib.PATH_MMA)[0]
    assert got.dtype == F16
    G.assert_close(got, _ext_oracle(c), *_tol(kv_dtype), what=f"fp16 extend mma {hq}/{hkv}/{d}/{kv}")


TC_CASES = [
    (32, 8, 128, [300, 0], [257, 200]),
    (8, 1, 128, [128, 5], [130, 300]),                # g = 8
    (32, 8, 64, [63, 64, 65], [64, 63, 200]),         # head_dim 64
    (32, 8, 128, [640], [1030]),                      # many tiles: ring wrap-around, lazy rescale
    (14, 2, 128, [100, 256], [90, 128]),              # g = 7 on tcgen05: 36 tokens x 7 heads, 4 dead rows per tile
]


@pytest.mark.parametrize("hq,hkv,d,pre,ext", TC_CASES)
@pytest.mark.parametrize("kv", ["fp16", "fp8"])
def test_fp16_extend_tcgen05_path_vs_oracle(hq, hkv, d, pre, ext, kv):
    kv_dtype = F16 if kv == "fp16" else FP8
    c = G.random_extend_case(hq, hkv, d, pre, ext, seed=hq * 3 + hkv + len(pre), kv_dtype=kv_dtype, act_dtype=F16)
    got = G.run_extend_case(c, force_path=_lib.PATH_TC)[0]
    assert got.dtype == F16
    G.assert_close(got, _ext_oracle(c), *_tol(kv_dtype), what=f"fp16 extend tcgen05 {hq}/{hkv}/{d}/{kv}")
    G.assert_close(got, G.run_extend_case(c, force_path=_lib.PATH_MMA)[0], *_tol(kv_dtype), what="tcgen05 vs mma.sync")


def test_fp16_extend_tcgen05_tma_tiles_golden_and_logit_cap():
    lib = _lib.load()
    # new-token tiles and consecutive prefix runs arrive by TMA (fp16 bytes through the same tensor maps)
    c = G.random_extend_case(8, 2, 128, [640, 130], [300, 520], seed=32, kv_dtype=F16, mode="contiguous")
    lib.b200_debug_extend_tma_tiles(1)
    got = G.run_extend_case(c, force_path=_lib.PATH_TC)[0]
    assert lib.b200_debug_extend_tma_tiles(1) > 0
    G.assert_close(got, _ext_oracle(c), G.F16_ABS, G.F16_REL, "fp16 tc extend with TMA tiles")
    # the reference kernel's fixture, inputs converted exactly to fp16
    g = load_golden("extend_gqa4_d128")
    cg = dict(g, req_pool_indices=g["req_pool_indices"].long(), seq_lens=g["seq_lens"].long(), kv_dtype=F16)
    for k in ("q", "k_new", "v_new", "k_buf", "v_buf"):
        cg[k] = g[k].to(F16)
    for path in (_lib.PATH_TC, _lib.PATH_MMA):
        got = G.run_extend_case(cg, logit_cap=g["logit_cap"], save_kv_cache=False, force_path=path)[0]
        G.assert_close(got, g["out"], G.F16_ABS, G.F16_REL, f"fp16 extend golden path {path}")
    c2 = G.random_extend_case(16, 4, 128, [50, 0, 129], [140, 165, 1], seed=3, kv_dtype=F16)
    for path in (_lib.PATH_TC, _lib.PATH_MMA):
        got = G.run_extend_case(c2, logit_cap=30.0, idx_dtype=torch.int32, force_path=path)[0]
        G.assert_close(got, _ext_oracle(c2, 30.0), G.F16_ABS, G.F16_REL, f"fp16 extend logit cap path {path}")
    with pytest.raises(_lib.B200AttnError):  # the CTA-pair kernel is bf16-only (opt-in)
        G.run_extend_case(c2, force_path=_lib.PATH_TC2)


def test_fp16_extend_writes_pool_then_redundant_attention():
    """save_kv_cache=True stores the fp16 new tokens; full attention read back from the pool agrees
    (the reference's redundant_attention cross-check, extend_attention.py:330-364)."""
    c = G.random_extend_case(8, 2, 128, [130, 0, 77], [64, 150, 3], seed=31, kv_dtype=F16)
    got, runner = G.run_extend_case(c, save_kv_cache=True)
    pool = runner.token_to_kv_pool
    kb, vb = pool.k_buffer[0].cpu(), pool.v_buffer[0].cpu()
    loc = c["out_cache_loc"]
    assert torch.equal(kb[loc].view(torch.int16), c["k_new"].view(torch.int16))  # raw fp16 bits, no conversion
    assert torch.equal(vb[loc].view(torch.int16), c["v_new"].view(torch.int16))
    ref = O.redundant_attention_ref(c["q"], kb, vb, c["req_to_token"], c["req_pool_indices"], c["seq_lens"],
                                    c["extend_seq_lens"], c["extend_start_loc"], c["sm_scale"])
    G.assert_close(got, ref, G.F16_ABS, G.F16_REL, "fp16 extend vs redundant attention")


# ---------------------------------------------------------------------------------------------------------------
# KV store and the fused rotary store: bit-exact
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kv_dtype", [F16, FP8])
def test_fp16_kv_store_bit_exact(kv_dtype):
    """fp16 pool: the activation's bits; fp8 pool: e4m3(fp16(x / scale)) -- the reference's in-place div_ on the fp16
    tensor then .to(fp8) (pool.py:404-412) -- per-head scales, overflow to NaN included, caller's k / v untouched."""
    from scratchpad_b200.pool import kv_dtype_code, kv_store
    hkv, d, t, pool_tokens = 4, 128, 333, 1000
    gen = torch.Generator().manual_seed(8)
    k = (torch.randn((t, hkv, d), generator=gen) * 3).to(F16)
    v = torch.randn((t, hkv, d), generator=gen).to(F16)
    k[0, 0, 0], k[1, 1, 1], k[2, 2, 2] = 60000.0, -60000.0, 6e-8   # overflow after / scale; an fp16 subnormal
    loc = (torch.randperm(pool_tokens, generator=gen)[:t] + 1).to(torch.int64)
    ks = vs = None
    if kv_dtype == FP8:
        ks = torch.tensor([0.037, 0.5, 1.7, 0.0021])
        vs = torch.tensor([0.011, 2.0, 0.3, 1.0])
    store = torch.uint8 if kv_dtype == FP8 else F16
    kb_ref = torch.zeros((pool_tokens + 1, hkv, d), dtype=store)
    vb_ref = torch.zeros_like(kb_ref)
    O.set_kv_buffer_ref(kb_ref, vb_ref, loc, k.clone(), v.clone(), kv_dtype, ks, vs)
    kb, vb = torch.zeros_like(kb_ref).cuda(), torch.zeros_like(vb_ref).cuda()
    kd, vd = k.cuda(), v.cuda()
    kv_store(kb, vb, loc.cuda(), kd, vd, kv_dtype_code(kv_dtype), None if ks is None else ks.cuda(),
             None if vs is None else vs.cuda())
    torch.cuda.synchronize()
    as_bits = lambda x: x.view(torch.int16) if x.dtype == F16 else x
    assert torch.equal(as_bits(kb.cpu()), as_bits(kb_ref)) and torch.equal(as_bits(vb.cpu()), as_bits(vb_ref))
    assert torch.equal(kd.cpu().view(torch.int16), k.view(torch.int16))  # not divided in place
    with pytest.raises(NotImplementedError):  # a 16-bit pool has the model dtype
        kv_store(torch.zeros_like(kb_ref, dtype=torch.bfloat16).cuda(), torch.zeros_like(vb_ref, dtype=torch.bfloat16).cuda(),
                 loc.cuda(), kd, vd, _lib.KV_BF16, None, None)


@pytest.mark.parametrize("kv_dtype", [F16, FP8])
@pytest.mark.parametrize("neox,d,rot", [(True, 128, 128), (False, 64, 64), (True, 128, 64)])
def test_fp16_fused_rope_store_matches_oracle_bit_exact(kv_dtype, neox, d, rot):
    """b200_rope_kv_store in the fp16 model dtype: q rotated in place, pool rows = quant(rope(k)), quant(v) at loc.
    The rotation itself is pinned to the reference module by tests/golden/ropeh_*.npz (test_rope_gpu.py)."""
    from scratchpad_b200.rotary_embedding import get_rope
    llama3 = {"rope_type": "llama3", "factor": 32.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
              "original_max_position_embeddings": 8192}
    hq, hkv, t, pool_tokens = 8, 2, 67, 300
    gen = torch.Generator().manual_seed(12)
    emb = get_rope(d, rot, 4096, 500000.0, neox, llama3, dtype=F16)
    assert emb.cos_sin_cache.dtype == F16
    qkv = torch.randn((t, (hq + 2 * hkv) * d), generator=gen).to(F16)
    pos = torch.randint(0, 4096, (t,), generator=gen)
    loc = (torch.randperm(pool_tokens, generator=gen)[:t] + 1).to(torch.int64)
    q, k, v = qkv.split([hq * d, hkv * d, hkv * d], dim=-1)
    q_ref = R.rope_ref(pos, q.reshape(t, hq, d), emb.cos_sin_cache.cpu(), rot, neox)
    k_ref = R.rope_ref(pos, k.reshape(t, hkv, d), emb.cos_sin_cache.cpu(), rot, neox)
    ks = vs = None
    if kv_dtype == FP8:
        ks = k_ref.float().abs().amax(dim=(0, 2)) / 448.0
        vs = v.reshape(t, hkv, d).float().abs().amax(dim=(0, 2)) / 448.0
    store = torch.uint8 if kv_dtype == FP8 else F16
    kb_ref = torch.zeros((pool_tokens + 1, hkv, d), dtype=store)
    vb_ref = torch.zeros_like(kb_ref)
    O.set_kv_buffer_ref(kb_ref, vb_ref, loc, k_ref.clone(), v.reshape(t, hkv, d).clone(), kv_dtype, ks, vs)
    bits = lambda x: x.contiguous().view(torch.int16) if x.dtype == F16 else x
    for write_back in (False, True):
        runner = synth.make_model_runner(hq, hkv, d, tp_size=1, layer_num=2, pool_tokens=pool_tokens, max_reqs=4,
                                         context_len=64, kv_dtype=kv_dtype, dtype=F16)
        pool = runner.token_to_kv_pool
        if ks is not None:
            pool.set_kv_scales(1, ks, vs)
        dev = qkv.cuda()
        qd, kd, vd = dev.split([hq * d, hkv * d, hkv * d], dim=-1)
        emb.rope_kv_store(pos.cuda(), qd, kd, vd, pool, 1, loc.cuda(), write_back)
        torch.cuda.synchronize()
        assert torch.equal(bits(qd.cpu()), bits(q_ref.reshape(t, -1)))
        assert torch.equal(bits(kd.cpu()), bits(k_ref.reshape(t, -1)) if write_back else bits(k))
        assert torch.equal(bits(pool.k_buffer[1].cpu()), bits(kb_ref)) and torch.equal(bits(pool.v_buffer[1].cpu()), bits(vb_ref))


def test_fp16_backend_refuses_mismatched_dtypes():
    c = G.random_decode_case(8, 2, 64, [40, 3], seed=1, kv_dtype=F16)
    runner = G.make_runner_from_case(c["k_buf"], c["v_buf"], c["req_to_token"], F16, tp_hq=8, dtype=torch.bfloat16)
    with pytest.raises(NotImplementedError):  # bf16 model over an fp16 pool
        CudaB200AttnBackend(runner)
    runner = G.make_runner_from_case(c["k_buf"], c["v_buf"], c["req_to_token"], F16, tp_hq=8, dtype=F16)
    backend = CudaB200AttnBackend(runner)
    layer = synth.make_layer(runner, 0, 0.0)
    fb = ForwardBatch(forward_mode=ForwardMode.DECODE, batch_size=2, req_pool_indices=c["req_pool_indices"].cuda(),
                      seq_lens=c["seq_lens"].cuda(), out_cache_loc=torch.zeros(2, dtype=torch.int64, device="cuda"),
                      seq_lens_sum=43, req_to_token_pool=runner.req_to_token_pool,
                      token_to_kv_pool=runner.token_to_kv_pool, attn_backend=backend)
    backend.init_forward_metadata(fb)
    with pytest.raises(TypeError):  # a bf16 q handed to a backend built for an fp16 model would be misread
        layer(c["q"].reshape(2, -1).to(torch.bfloat16).cuda(), None, None, fb, save_kv_cache=False)


# ---------------------------------------------------------------------------------------------------------------
# the reference's own GPU path on the same fp16 bytes (baseline/_ref; skipped when it is not installed)
# ---------------------------------------------------------------------------------------------------------------
from baseline import ref_loader  # noqa: E402

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="baseline/_ref not installed (python baseline/install_ref.py)")


def _flashinfer_ok():
    try:
        import flashinfer  # noqa: F401
        return True
    except Exception:
        return False


needs_fi = pytest.mark.skipif(not _flashinfer_ok(), reason="flashinfer is not importable")


def _ref_state(runner, hq, kv_dtype, k_scale=None, v_scale=None, sm_scale=None):
    from baseline import comparators as C
    pool = runner.token_to_kv_pool
    return C.RefState(req_to_token=runner.req_to_token_pool.req_to_token, k_buffers=pool.k_buffer,
                      v_buffers=pool.v_buffer, kv_dtype=kv_dtype, hq=hq, hkv=pool.head_num, head_dim=pool.head_dim,
                      k_scale=k_scale, v_scale=v_scale, sm_scale=sm_scale, model_dtype=F16)


@needs_ref
@needs_fi
@pytest.mark.parametrize("hq,hkv,d,lens", [(32, 8, 64, [2048, 1, 17, 300, 1025]), (32, 8, 128, [4100, 129, 64, 2, 999, 31])])
@pytest.mark.parametrize("kv_dtype", [F16, FP8])
def test_fp16_decode_matches_reference_flashinfer_backend(hq, hkv, d, lens, kv_dtype):
    from baseline import comparators as C
    fp8 = kv_dtype == FP8
    c = G.random_decode_case(hq, hkv, d, lens, seed=13 + d, kv_dtype=kv_dtype, act_dtype=F16, per_head=False)
    ours, backend, runner, fb, layer = _dec_run(c)
    ks = float(c["k_scale"][0]) if fp8 else None
    vs = float(c["v_scale"][0]) if fp8 else None
    st = _ref_state(runner, hq, kv_dtype, ks, vs, sm_scale=c["sm_scale"])
    ref = C.RefFlashInfer(st, context_len=runner.model_config.context_len)
    rfb = st.decode_batch(fb.req_pool_indices, fb.seq_lens, fb.out_cache_loc, ref.backend)
    ref.init_forward_metadata(rfb)
    theirs = ref.forward(0, c["q"].reshape(len(lens), hq * d).cuda(), None, None, rfb, save_kv_cache=False)
    torch.cuda.synchronize()
    assert theirs.dtype == F16
    theirs = theirs.view(len(lens), hq, d).float().cpu()
    tol = _tol(kv_dtype)
    G.assert_close(ours, theirs, *tol, what=f"fp16: ours vs reference FlashInfer backend ({kv_dtype})")
    G.assert_close(theirs, _dec_oracle(c), *tol, what="fp16: reference FlashInfer backend vs oracle")
    assert ((ours.float() - theirs).abs() <= tol[0] + tol[1] * theirs.abs()).all()


@needs_ref
@needs_fi
@pytest.mark.parametrize("force_path", [-1, 0])
def test_fp16_extend_matches_reference_flashinfer_backend(force_path):
    from baseline import comparators as C
    hq, hkv, d, pre, ext = 32, 8, 128, [300, 0, 129, 1000], [130, 257, 1, 64]
    c = G.random_extend_case(hq, hkv, d, pre, ext, seed=6, kv_dtype=F16)
    ours, runner = G.run_extend_case(c, force_path=force_path, save_kv_cache=False)
    st = _ref_state(runner, hq, F16, sm_scale=c["sm_scale"])
    ref = C.RefFlashInfer(st, context_len=runner.model_config.context_len)
    el = c["extend_seq_lens"].to(torch.int64).cuda()
    sl = c["seq_lens"].to(torch.int64).cuda()
    rfb = st.extend_batch(c["req_pool_indices"].cuda(), sl, c["out_cache_loc"].cuda(), el, sl - el, ref.backend)
    ref.init_forward_metadata(rfb)
    t = c["q"].shape[0]
    theirs = ref.forward(0, c["q"].reshape(t, -1).cuda(), c["k_new"].reshape(t, -1).cuda(),
                         c["v_new"].reshape(t, -1).cuda(), rfb, save_kv_cache=False)
    torch.cuda.synchronize()
    G.assert_close(ours, theirs.view(t, hq, d).float().cpu(), G.F16_ABS, G.F16_REL, "fp16 extend: ours vs reference FlashInfer")


@needs_ref
def test_fp16_decode_and_extend_match_reference_triton_kernels():
    from baseline import comparators as C
    c = G.random_decode_case(32, 8, 128, [4100, 129, 64, 2, 999, 31], seed=15, kv_dtype=F16)
    ours, backend, runner, fb, layer = _dec_run(c)
    pool = runner.token_to_kv_pool
    theirs = C.triton_decode(c["q"].cuda(), pool.k_buffer[0], pool.v_buffer[0], runner.req_to_token_pool.req_to_token,
                             fb.req_pool_indices, fb.seq_lens, c["sm_scale"], 0.0)
    torch.cuda.synchronize()
    G.assert_close(ours, theirs.float().cpu(), G.F16_ABS, G.F16_REL, "fp16: ours vs in-tree Triton decode")
    e = G.random_extend_case(32, 8, 128, [300, 0, 129], [130, 257, 1], seed=16, kv_dtype=F16)
    ours, runner = G.run_extend_case(e, save_kv_cache=False)
    pool = runner.token_to_kv_pool
    theirs = C.triton_extend(e["q"].cuda(), e["k_new"].cuda(), e["v_new"].cuda(), pool.k_buffer[0], pool.v_buffer[0],
                             runner.req_to_token_pool.req_to_token, e["req_pool_indices"].cuda(), e["seq_lens"].cuda(),
                             e["extend_seq_lens"].cuda(), e["extend_start_loc"].cuda(), e["sm_scale"])
    torch.cuda.synchronize()
    G.assert_close(ours, theirs.float().cpu(), G.F16_ABS, G.F16_REL, "fp16: ours vs in-tree Triton extend")
