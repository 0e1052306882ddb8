"""-m gpu parity tests of the decode path: CUDA kernels (through the Python
backend surface and the C ABI) vs the fp32 oracle, the golden fixtures written
from the reference's own kernels, and size-independent properties."""
import pytest
import torch

from oracle import attn_oracle as O
from golden_util import golden_names, load_golden
from scratchpad_b200 import _lib

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import gpu_util as G


def _oracle(c, logit_cap=0.0):
    return O.decode_ref(c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"], c["seq_lens"],
                        c["sm_scale"], logit_cap, c.get("kv_dtype"), c.get("k_scale"), c.get("v_scale"))


def _run(c, logit_cap=0.0, **kw):
    return G.run_decode_case(c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"],
                             c["seq_lens"], c["sm_scale"], logit_cap, c.get("kv_dtype", torch.bfloat16),
                             c.get("k_scale"), c.get("v_scale"), **kw)[0]


@pytest.mark.parametrize("name", golden_names("decode"))
def test_decode_matches_reference_golden(name):
    g = load_golden(name)
    got = G.run_decode_case(g["q"], g["k_buf"], g["v_buf"], g["req_to_token"], g["req_pool_indices"].long(),
                            g["seq_lens"].long(), g["sm_scale"], g["logit_cap"])[0]
    G.assert_close(got, g["out"], G.BF16_ABS, G.BF16_REL, name)


SHAPES = [
    # hq, hkv, d   -> kernel variant exercised
    (32, 8, 64),    # C2: HPC=8, 8 warps, one slice
    (8, 2, 64),     # HPC=2, 4 token sub-slices
    (4, 1, 64),     # HPC=1 (TP=8 shard of Llama-3.2-1B)
    (32, 8, 128),   # 8B TP1: bf16 -> HPC=4 x 2 head groups; fp8 -> HPC=8
    (8, 1, 128),    # 70B TP8 shard, g=8 (dense N=8 tile)
    (16, 4, 128),
    (7, 1, 64),     # g=7 (Qwen2-7B shard), odd group
    (4, 4, 128),    # MHA g=1
    (6, 2, 64),     # g=3
    (16, 1, 128),   # g=16 (Llama-3.1-405B TP8 shard): two N=8 q tiles per kv head, HPC 1
    (64, 4, 128),   # g=16, several kv heads per CTA
    (12, 1, 128),   # g=12: second q tile half empty
]
LENS = [[1], [15, 16, 17], [127, 128, 129, 1, 300], [513, 64, 2, 1000]]


@pytest.mark.parametrize("hq,hkv,d", SHAPES)
@pytest.mark.parametrize("kv", ["bf16", "fp8"])
def test_decode_vs_oracle_shapes(hq, hkv, d, kv):
    kv_dtype = torch.bfloat16 if kv == "bf16" else torch.float8_e4m3fn
    tol = (G.BF16_ABS, G.BF16_REL) if kv == "bf16" else (G.FP8_ABS, G.FP8_REL)
    for i, lens in enumerate(LENS):
        c = G.random_decode_case(hq, hkv, d, lens, seed=100 * hq + 10 * hkv + i, kv_dtype=kv_dtype)
        got = _run(c)
        G.assert_close(got, _oracle(c), *tol, what=f"{hq}/{hkv}/{d}/{kv}/{lens}")


@pytest.mark.parametrize("idx_dtype", [torch.int32, torch.int64])
def test_decode_index_dtypes_shared_prefix_and_stale_tails(idx_dtype):
    # equal slot ids in different rows (RadixAttention prefix sharing), stale garbage past seq_len,
    # shuffled req_pool_indices, both index dtypes the engine produces
    c = G.random_decode_case(32, 8, 64, [200, 200, 231, 199, 64, 260], seed=5, shared_prefix=192)
    got = _run(c, idx_dtype=idx_dtype)
    G.assert_close(got, _oracle(c), G.BF16_ABS, G.BF16_REL)


def test_decode_logit_cap_and_qkv_slice():
    c = G.random_decode_case(32, 8, 128, [77, 300, 5], seed=6)
    got = _run(c, logit_cap=30.0, q_slice_of_qkv=True)
    G.assert_close(got, _oracle(c, 30.0), G.BF16_ABS, G.BF16_REL)


def test_decode_fp8_per_tensor_scale():
    c = G.random_decode_case(32, 8, 128, [333, 48], seed=8, kv_dtype=torch.float8_e4m3fn, per_head=False)
    G.assert_close(_run(c), _oracle(c), G.FP8_ABS, G.FP8_REL)


def test_decode_zero_length_and_padded_rows():
    # seq_len 0 -> zeros; graph-padded rows arrive as seq_len = fill value 1 with req_pool_index 0
    c = G.random_decode_case(8, 2, 64, [0, 40, 1, 1], seed=9)
    c["req_pool_indices"][2:] = 0
    got = _run(c)
    ref = _oracle(c)
    assert got[0].abs().max().item() == 0.0
    G.assert_close(got, ref, G.BF16_ABS, G.BF16_REL)


def test_decode_many_requests_many_ctas_vs_naive_kernel():
    """Larger than the CPU oracle likes: compare the tiled kernels with the
    one-warp-per-head fp32 kernel on the same device buffers, and the oracle on a sample."""
    lens = [int(x) for x in torch.randint(1, 1500, (96,), generator=torch.Generator().manual_seed(3))]
    for kv_dtype, tol in ((torch.bfloat16, (G.BF16_ABS, G.BF16_REL)), (torch.float8_e4m3fn, (G.FP8_ABS, G.FP8_REL))):
        c = G.random_decode_case(32, 8, 128, lens, seed=11, kv_dtype=kv_dtype)
        got, backend, runner, fb, layer = G.run_decode_case(
            c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"], c["seq_lens"], c["sm_scale"],
            0.0, kv_dtype, c["k_scale"], c["v_scale"])
        pool = runner.token_to_kv_pool
        q = c["q"].cuda()
        o = torch.empty_like(q)
        ks, vs = pool.get_kv_scales(0)
        p = _lib.DecodeParams(
            q=q.data_ptr(), q_stride_t=q.stride(0), q_stride_h=q.stride(1),
            k_buf=pool.k_buffer[0].data_ptr(), v_buf=pool.v_buffer[0].data_ptr(),
            k_stride_t=pool.k_buffer[0].stride(0), v_stride_t=pool.v_buffer[0].stride(0),
            o=o.data_ptr(), o_stride_t=o.stride(0), o_stride_h=o.stride(1),
            req_to_token=runner.req_to_token_pool.req_to_token.data_ptr(),
            req_to_token_stride=runner.req_to_token_pool.req_to_token.stride(0),
            bs=len(lens), hq=32, hkv=8, d=128, dv=128, kv_dtype=pool.kv_dtype_code(), sm_scale=c["sm_scale"],
            logit_cap=0.0, k_scale=None if ks is None else ks.data_ptr(), v_scale=None if vs is None else vs.data_ptr(),
            workspace=None, workspace_bytes=0, max_bs=len(lens), force_path=-1)
        rc = _lib.load().b200_debug_decode_naive(p, fb.req_pool_indices.data_ptr(), fb.seq_lens.data_ptr(),
                                                 _lib.IDX_I64, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "naive")
        torch.cuda.synchronize()
        G.assert_close(got, o.cpu(), *tol, what="tiled vs naive")
        sub = {k: (v[:6] if k in ("q", "req_pool_indices", "seq_lens") else v) for k, v in c.items()}
        G.assert_close(got[:6], _oracle(sub), *tol, what="tiled vs oracle sample")


def test_gather_is_bit_exact():
    """rows consumed == k_buf[req_to_token[req_pool_indices[i], t]] bit for bit."""
    lib = _lib.load()
    for idx_dtype, code in ((torch.int32, _lib.IDX_I32), (torch.int64, _lib.IDX_I64)):
        for kv_dtype in (torch.bfloat16, torch.float8_e4m3fn):
            c = G.random_decode_case(8, 2, 128, [33, 1, 200, 64], seed=21, kv_dtype=kv_dtype, shared_prefix=32)
            c["req_to_token"][c["req_pool_indices"][1], 0] = 0  # slot 0 is an ordinary row for the gather
            buf = c["k_buf"].cuda()
            r2t = c["req_to_token"].cuda()
            rpi, sl = c["req_pool_indices"].to(idx_dtype).cuda(), c["seq_lens"].to(idx_dtype).cuda()
            offs = torch.zeros(4, dtype=torch.int64)
            offs[1:] = torch.cumsum(c["seq_lens"][:-1], 0)
            row_bytes = buf.shape[1] * buf.shape[2] * buf.element_size()
            out = torch.zeros((int(c["seq_lens"].sum()), row_bytes), dtype=torch.uint8, device="cuda")
            rc = lib.b200_debug_gather(buf.data_ptr(), buf.stride(0) * buf.element_size(), row_bytes, r2t.data_ptr(),
                                       r2t.stride(0), rpi.data_ptr(), sl.data_ptr(), code, offs.cuda().data_ptr(), 4,
                                       out.data_ptr(), torch.cuda.current_stream().cuda_stream)
            _lib.check(rc, "gather")
            torch.cuda.synchronize()
            ref = torch.cat([O.gather_rows_ref(c["k_buf"], c["req_to_token"], c["req_pool_indices"][i], c["seq_lens"][i])
                             for i in range(4)])
            assert torch.equal(out.cpu().view(-1), ref.contiguous().view(torch.uint8).view(-1))


@pytest.mark.parametrize("kv", ["bf16", "fp8"])
def test_kv_store_bit_exact_then_decode(kv):
    """set_kv_buffer (fused store kernel) is bit-exact vs pool.py:392-424 semantics, does not touch the
    caller's k/v, and the decode that follows sees the new token (write-then-attend)."""
    kv_dtype = torch.bfloat16 if kv == "bf16" else torch.float8_e4m3fn
    c = G.random_decode_case(32, 8, 64, [50, 9, 130], seed=31, kv_dtype=kv_dtype)
    gen = torch.Generator().manual_seed(32)
    b, hkv, d = 3, 8, 64
    k_new = torch.randn((b, hkv, d), generator=gen).to(torch.bfloat16)
    v_new = torch.randn((b, hkv, d), generator=gen).to(torch.bfloat16)
    loc = torch.stack([c["req_to_token"][c["req_pool_indices"][i], c["seq_lens"][i] - 1] for i in range(b)]).long()
    runner = G.make_runner_from_case(c["k_buf"], c["v_buf"], c["req_to_token"], kv_dtype, c["k_scale"], c["v_scale"], 32)
    from scratchpad_b200 import CudaB200AttnBackend
    from harness import ForwardBatch, ForwardMode, synth
    backend = CudaB200AttnBackend(runner)
    layer = synth.make_layer(runner)
    fb = ForwardBatch(ForwardMode.DECODE, b, c["req_pool_indices"].cuda(), c["seq_lens"].cuda(), loc.cuda(),
                      req_to_token_pool=runner.req_to_token_pool, token_to_kv_pool=runner.token_to_kv_pool,
                      attn_backend=backend)
    # k, v arrive as slices of the qkv buffer
    qkv = torch.zeros((b, (32 + 16) * d), dtype=torch.bfloat16, device="cuda")
    qkv[:, : 32 * d] = c["q"].reshape(b, -1).cuda()
    qkv[:, 32 * d: 40 * d] = k_new.reshape(b, -1).cuda()
    qkv[:, 40 * d:] = v_new.reshape(b, -1).cuda()
    qd, kd, vd = qkv[:, : 32 * d], qkv[:, 32 * d: 40 * d], qkv[:, 40 * d:]
    backend.init_forward_metadata(fb)
    o = layer(qd, kd, vd, fb, save_kv_cache=True)
    torch.cuda.synchronize()
    assert torch.equal(kd.cpu().view(b, hkv, d), k_new), "caller's k was modified"
    kb, vb = c["k_buf"].clone(), c["v_buf"].clone()
    O.set_kv_buffer_ref(kb, vb, loc, k_new, v_new, kv_dtype, c["k_scale"], c["v_scale"])
    pool = runner.token_to_kv_pool
    assert torch.equal(pool.k_buffer[0].cpu().view(torch.uint8), kb.view(torch.uint8))
    assert torch.equal(pool.v_buffer[0].cpu().view(torch.uint8), vb.view(torch.uint8))
    c2 = dict(c, k_buf=kb, v_buf=vb)
    tol = (G.BF16_ABS, G.BF16_REL) if kv == "bf16" else (G.FP8_ABS, G.FP8_REL)
    G.assert_close(o.view(b, 32, d).cpu(), _oracle(c2), *tol)


def test_decode_under_cuda_graph_capture_and_replay():
    """Capture at a padded batch size with static int32 buffers, then replay twice with different
    sequence lengths after the replay hook (cuda_graph_runner.py:394-423,443-510)."""
    from scratchpad_b200 import CudaB200AttnBackend
    from harness import ForwardBatch, ForwardMode, synth
    hq, hkv, d, bs = 32, 8, 64, 8
    lens_a = [100, 200, 17, 64, 1, 1, 1, 1]
    lens_b = [31, 500, 333, 128, 250, 1, 1, 1]
    c = G.random_decode_case(hq, hkv, d, [600] * bs, seed=41)
    runner = G.make_runner_from_case(c["k_buf"], c["v_buf"], c["req_to_token"], torch.bfloat16, tp_hq=hq)
    backend = CudaB200AttnBackend(runner)
    layer = synth.make_layer(runner)
    backend.init_cuda_graph_state(16)
    s_rpi = torch.zeros(16, dtype=torch.int32, device="cuda")
    s_sl = torch.full((16,), backend.get_cuda_graph_seq_len_fill_value(), dtype=torch.int32, device="cuda")
    s_q = torch.zeros((bs, hq * d), dtype=torch.bfloat16, device="cuda")
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
    for lens in (lens_a, lens_b):
        s_rpi[:bs].copy_(c["req_pool_indices"].to(torch.int32))
        s_sl[:bs].copy_(torch.tensor(lens, dtype=torch.int32))
        s_q.copy_(c["q"].reshape(bs, -1))
        backend.init_forward_metadata_replay_cuda_graph(bs, s_rpi, s_sl, sum(lens), None, ForwardMode.DECODE, None,
                                                        seq_lens_cpu=None)
        graph.replay()
        torch.cuda.synchronize()
        cc = dict(c, seq_lens=torch.tensor(lens))
        G.assert_close(s_o.view(bs, hq, d).cpu(), _oracle(cc), G.BF16_ABS, G.BF16_REL, what=f"graph replay {lens}")


def test_unsupported_configs_raise():
    from scratchpad_b200 import CudaB200AttnBackend
    from harness import synth
    runner = synth.make_model_runner(32, 8, 96, pool_tokens=64, max_reqs=4, context_len=64)
    with pytest.raises(_lib.B200AttnError):
        CudaB200AttnBackend(runner)
    runner = synth.make_model_runner(32, 2, 64, pool_tokens=64, max_reqs=4, context_len=64)  # g = 16 at head_dim 64
    with pytest.raises(_lib.B200AttnError):
        CudaB200AttnBackend(runner)
    runner = synth.make_model_runner(64, 2, 128, pool_tokens=64, max_reqs=4, context_len=64)  # g = 32
    with pytest.raises(_lib.B200AttnError):
        CudaB200AttnBackend(runner)


@pytest.mark.parametrize("kv", ["bf16", "fp8"])
@pytest.mark.parametrize("hq,hkv", [(16, 1), (128, 8)])
def test_decode_group_size_16_many_ctas_paged_and_graph(hq, hkv, kv):
    """Llama-3.1-405B head layout (128 q / 8 kv heads; 16 / 1 at TP8): two q tiles per kv head share every staged K/V
    tile.  Enough requests that CTAs cross request boundaries (flush, in-CTA sub-slice combination where it fits),
    over a token-granular pool, over 128-token pages (TMA tiles) and with int32 indices."""
    kv_dtype = torch.bfloat16 if kv == "bf16" else torch.float8_e4m3fn
    tol = (G.BF16_ABS, G.BF16_REL) if kv == "bf16" else (G.FP8_ABS, G.FP8_REL)
    lens = [int(x) for x in torch.randint(1, 1200, (40,), generator=torch.Generator().manual_seed(hq))]
    c = G.random_decode_case(hq, hkv, 128, lens, seed=70 + hq, kv_dtype=kv_dtype)
    got, backend = G.run_decode_case(c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"], c["seq_lens"],
                                     c["sm_scale"], 0.0, kv_dtype, c["k_scale"], c["v_scale"], idx_dtype=torch.int32)[:2]
    assert backend.config.q_tiles == 2
    G.assert_close(got, _oracle(c), *tol, what=f"g=16 {hq}/{hkv}/{kv}")
    cp = G.random_decode_case(hq, hkv, 128, lens, seed=71 + hq, kv_dtype=kv_dtype, page_size=128)
    gotp = G.run_decode_case(cp["q"], cp["k_buf"], cp["v_buf"], cp["req_to_token"], cp["req_pool_indices"], cp["seq_lens"],
                             cp["sm_scale"], 0.0, kv_dtype, cp["k_scale"], cp["v_scale"], page_size=128)[0]
    G.assert_close(gotp, _oracle(cp), *tol, what=f"g=16 paged {hq}/{hkv}/{kv}")


@pytest.mark.parametrize("kv", ["bf16", "fp8"])
def test_fused_kv_pool_layout_decode_store_extend(kv, monkeypatch):
    """The fused [P+1, 2, Hkv, D] pool option (strided k/v views, same slot indexing) gives identical results."""
    monkeypatch.setattr(G, "FUSED_KV", True)
    kv_dtype = torch.bfloat16 if kv == "bf16" else torch.float8_e4m3fn
    tol = (G.BF16_ABS, G.BF16_REL) if kv == "bf16" else (G.FP8_ABS, G.FP8_REL)
    for hq, hkv, d in ((8, 1, 128), (4, 1, 64), (32, 8, 64)):
        c = G.random_decode_case(hq, hkv, d, [129, 1, 700, 64], seed=51, kv_dtype=kv_dtype)
        got, backend, runner, fb, layer = G.run_decode_case(
            c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"], c["seq_lens"], c["sm_scale"],
            0.0, kv_dtype, c["k_scale"], c["v_scale"])
        pool = runner.token_to_kv_pool
        assert pool.fused_kv and pool.k_buffer[0].stride(0) == 2 * hkv * d
        assert pool.v_buffer[0].data_ptr() - pool.k_buffer[0].data_ptr() == hkv * d * pool.k_buffer[0].element_size()
        G.assert_close(got, _oracle(c), *tol, what=f"fused decode {hq}/{hkv}/{d}")
    ce = G.random_extend_case(8, 1, 128, [100, 0], [37, 70], seed=52, kv_dtype=kv_dtype)
    got, runner = G.run_extend_case(ce, save_kv_cache=True)
    ref = O.extend_ref(ce["q"], ce["k_new"], ce["v_new"], ce["k_buf"], ce["v_buf"], ce["req_to_token"],
                       ce["req_pool_indices"], ce["seq_lens"], ce["extend_seq_lens"], ce["extend_start_loc"],
                       ce["sm_scale"], 0.0, kv_dtype, ce["k_scale"], ce["v_scale"])
    G.assert_close(got, ref, *tol, what="fused extend")
    kb, vb = ce["k_buf"].clone(), ce["v_buf"].clone()
    O.set_kv_buffer_ref(kb, vb, ce["out_cache_loc"], ce["k_new"], ce["v_new"], kv_dtype, ce["k_scale"], ce["v_scale"])
    pool = runner.token_to_kv_pool
    assert torch.equal(pool.k_buffer[0].cpu().contiguous().view(torch.uint8), kb.view(torch.uint8))
    assert torch.equal(pool.v_buffer[0].cpu().contiguous().view(torch.uint8), vb.view(torch.uint8))


# ---- shared-prefix cascade: dense tcgen05 leg over the prefix the whole batch shares ----
def _shared_len(backend):
    return int(backend.workspace[:64].view(torch.int32)[7].item())


def _run_full(c, **kw):
    return G.run_decode_case(c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"],
                             c["seq_lens"], c["sm_scale"], 0.0, c.get("kv_dtype", torch.bfloat16),
                             c.get("k_scale"), c.get("v_scale"), **kw)


CASCADE_SHAPES = [
    # hq, hkv, bs
    (32, 8, 24),   # Llama-3.1-8B TP1, g=4: 96 query rows per kv head (one partial 256-row block)
    (8, 1, 40),    # 70B TP8 shard, g=8: 320 rows -> two row blocks, the second partial
    (4, 4, 70),    # MHA
]


@pytest.mark.parametrize("hq,hkv,bs", CASCADE_SHAPES)
@pytest.mark.parametrize("kv", ["bf16", "fp8"])
def test_decode_cascade_matches_oracle_and_split_kv(hq, hkv, bs, kv):
    kv_dtype = torch.bfloat16 if kv == "bf16" else torch.float8_e4m3fn
    tol = (G.BF16_ABS, G.BF16_REL) if kv == "bf16" else (G.FP8_ABS, G.FP8_REL)
    shared = 1300  # -> 1280 tokens through the dense leg, 20 + suffix through the split-KV kernel
    gen = torch.Generator().manual_seed(bs)
    lens = [shared + int(x) for x in torch.randint(1, 400, (bs,), generator=gen)]
    lens[0], lens[1] = shared + 1, shared + 333
    c = G.random_decode_case(hq, hkv, 128, lens, seed=7 * hq + bs, kv_dtype=kv_dtype, shared_prefix=shared)
    got, backend = _run_full(c, cascade=True)[:2]
    assert _shared_len(backend) == 1280, "the shared prefix was not detected"
    ref = _oracle(c)
    G.assert_close(got, ref, *tol, what=f"cascade {hq}/{hkv}/{bs}/{kv}")
    off, backend_off = _run_full(c, cascade=False)[:2]
    assert _shared_len(backend_off) == 0
    G.assert_close(got, off, *tol, what="cascade vs split-KV only")


def test_decode_cascade_partial_sharing_and_off_cases():
    # one request leaves the common prefix early: the detected length is the batch minimum, rounded down to 128
    c = G.random_decode_case(32, 8, 128, [2100] * 20, seed=11, shared_prefix=2048)
    r = int(c["req_pool_indices"][5])
    c["req_to_token"][r, 1100:2048] = c["req_to_token"][r, 1100:2048].flip(0)  # same slots, other order: prefix ends at 1100
    got, backend = _run_full(c, cascade=True)[:2]
    assert _shared_len(backend) == 1024
    G.assert_close(got, _oracle(c), G.BF16_ABS, G.BF16_REL, what="partial sharing")
    # below the 1024-token threshold, or too few query rows, the cascade leaves the batch alone: bit-identical output
    for lens, sp in (([900 + i for i in range(20)], 800), ([1500, 1501, 1502], 1408)):
        c = G.random_decode_case(32, 8, 128, lens, seed=12, shared_prefix=sp)
        on, backend = _run_full(c, cascade=True)[:2]
        assert _shared_len(backend) == 0
        assert torch.equal(on, _run_full(c, cascade=False)[0])


def test_decode_cascade_logit_cap_and_int32_indices():
    c = G.random_decode_case(32, 8, 128, [1200 + 7 * i for i in range(18)], seed=13, shared_prefix=1152)
    got, backend = G.run_decode_case(c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"],
                                     c["seq_lens"], c["sm_scale"], 20.0, idx_dtype=torch.int32, cascade=True)[:2]
    assert _shared_len(backend) == 1152
    G.assert_close(got, _oracle(c, 20.0), G.BF16_ABS, G.BF16_REL, what="cascade + softcap")


def test_decode_cascade_under_cuda_graph_with_padded_rows():
    """One captured graph serves a batch that shares a prefix (cascade on, detected on device), with the rows the
    graph runner pads the batch with (seq_len = fill value, req_pool_index 0), and then a batch that shares nothing."""
    from scratchpad_b200 import CudaB200AttnBackend
    from harness import ForwardBatch, ForwardMode, synth
    hq, hkv, d, bs, real = 32, 8, 128, 24, 20
    shared = 1300
    lens_shared = [shared + 3 + 11 * i for i in range(real)]
    c = G.random_decode_case(hq, hkv, d, lens_shared, seed=43, shared_prefix=shared)
    runner = G.make_runner_from_case(c["k_buf"], c["v_buf"], c["req_to_token"], torch.bfloat16, tp_hq=hq)
    backend = CudaB200AttnBackend(runner, cascade=True)
    layer = synth.make_layer(runner)
    layer.scaling = c["sm_scale"]
    backend.init_cuda_graph_state(32)
    s_rpi = torch.zeros(32, dtype=torch.int32, device="cuda")
    s_sl = torch.full((32,), backend.get_cuda_graph_seq_len_fill_value(), dtype=torch.int32, device="cuda")
    s_q = torch.zeros((bs, hq * d), dtype=torch.bfloat16, device="cuda")
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
    lens_unshared = [40 + 37 * i for i in range(real)]  # prefixes of the same rows, shorter than the 1024-token threshold
    for lens, want_shared in ((lens_shared, 1280), (lens_unshared, 0)):
        s_rpi.zero_()
        s_sl.fill_(backend.get_cuda_graph_seq_len_fill_value())
        s_rpi[:real].copy_(c["req_pool_indices"].to(torch.int32))
        s_sl[:real].copy_(torch.tensor(lens, dtype=torch.int32))
        s_q.zero_()
        s_q[:real].copy_(c["q"].reshape(real, -1))
        backend.init_forward_metadata_replay_cuda_graph(bs, s_rpi, s_sl, sum(lens) + bs - real, None, ForwardMode.DECODE,
                                                        None, seq_lens_cpu=None)
        graph.replay()
        torch.cuda.synchronize()
        assert _shared_len(backend) == want_shared
        cc = dict(c, seq_lens=torch.tensor(lens))
        G.assert_close(s_o.view(bs, hq, d)[:real].cpu(), _oracle(cc), G.BF16_ABS, G.BF16_REL, what=f"graph + cascade {want_shared}")
        assert torch.isfinite(s_o.float()).all()


def test_decode_cascade_head_dim_64():
    """Llama-3.2-1B shapes (head_dim 64) through the shared-prefix leg, bf16 and (round 2) fp8-e4m3 pools."""
    lens = [1100 + 9 * i for i in range(20)]
    c = G.random_decode_case(32, 8, 64, lens, seed=17, shared_prefix=1090)
    got, backend = _run_full(c, cascade=True)[:2]
    assert _shared_len(backend) == 1024
    G.assert_close(got, _oracle(c), G.BF16_ABS, G.BF16_REL, what="cascade d64")
    c = G.random_decode_case(32, 8, 64, lens, seed=17, shared_prefix=1090, kv_dtype=torch.float8_e4m3fn)
    got, backend = _run_full(c, cascade=True)[:2]
    assert _shared_len(backend) == 1024
    G.assert_close(got, _oracle(c), G.FP8_ABS, G.FP8_REL, what="cascade d64 fp8")
    off = _run_full(c, cascade=False)[0]
    G.assert_close(got, off, G.FP8_ABS, G.FP8_REL, what="cascade d64 fp8 vs split-KV only")


def _multi_group_case(hq, hkv, d, groups, seed, kv_dtype=torch.bfloat16):
    """Decode case with several prefix groups: groups = [(n_requests, shared_len, max_suffix), ...]; requests of a
    group share their first shared_len slot ids (what RadixAttention does for a common prompt), rows are shuffled."""
    gen = torch.Generator().manual_seed(seed)
    lens, specs = [], []
    for gi, (n, sh, suf) in enumerate(groups):
        for _ in range(n):
            s = sh + int(torch.randint(1, suf + 1, (1,), generator=gen))
            lens.append(s)
            specs.append((gi, sh))
    order = torch.randperm(len(lens), generator=gen).tolist()
    lens = [lens[i] for i in order]
    specs = [specs[i] for i in order]
    need = sum(sh for _, sh, _ in groups) + sum(s - sh for s, (_, sh) in zip(lens, specs))
    c = G.random_decode_case(hq, hkv, d, lens, seed=seed + 1, kv_dtype=kv_dtype)   # pools big enough? rebuild below
    pool = need + 33
    ids = (torch.randperm(pool, generator=gen) + 1).to(torch.int32)
    r2t = torch.zeros((len(lens) + 3, max(lens) + 4), dtype=torch.int32)
    r2t[:] = ids[torch.randint(0, pool, r2t.shape, generator=gen)]                # stale garbage everywhere
    cur = 0
    shared_ids = []
    for _, sh, _ in groups:
        shared_ids.append(ids[cur:cur + sh])
        cur += sh
    rows = torch.randperm(len(lens) + 3, generator=gen)[: len(lens)]
    for i, (s, (gi, sh)) in enumerate(zip(lens, specs)):
        r2t[rows[i], :sh] = shared_ids[gi]
        r2t[rows[i], sh:s] = ids[cur:cur + s - sh]
        cur += s - sh
    kf = torch.randn((pool + 1, hkv, d), generator=gen).to(torch.bfloat16)
    vf = torch.randn((pool + 1, hkv, d), generator=gen).to(torch.bfloat16)
    ks = vs = None
    if kv_dtype != torch.bfloat16:
        ks = kf.float().abs().amax(dim=(0, 2)) / 448.0
        vs = vf.float().abs().amax(dim=(0, 2)) / 448.0
        kf, vf = O.quantize_kv_ref(kf, kv_dtype, ks), O.quantize_kv_ref(vf, kv_dtype, vs)
    c.update(k_buf=kf, v_buf=vf, req_to_token=r2t, req_pool_indices=rows.to(torch.int64),
             seq_lens=torch.tensor(lens, dtype=torch.int64), k_scale=ks, v_scale=vs)
    return c, specs


@pytest.mark.parametrize("kv", ["bf16", "fp8"])
def test_decode_cascade_prefix_groups(kv):
    """Several radix subtrees in one batch: each sufficiently large group gets its own dense shared-prefix leg, small
    groups and singletons stay on the split-KV kernel; everything merges to the oracle's result."""
    kv_dtype = torch.bfloat16 if kv == "bf16" else torch.float8_e4m3fn
    tol = (G.BF16_ABS, G.BF16_REL) if kv == "bf16" else (G.FP8_ABS, G.FP8_REL)
    groups = [(70, 1300, 200),   # 280 query rows per kv head: two row tiles (64 + 6 requests), shared 1280
              (18, 2100, 60),    # one partial row tile, shared 2048
              (5, 1500, 40),     # 20 rows: below the 64-row threshold -> split-KV only
              (1, 0, 900), (1, 0, 33), (1, 0, 1400), (1, 0, 2)]
    c, specs = _multi_group_case(32, 8, 128, groups, seed=71, kv_dtype=kv_dtype)
    got, backend = _run_full(c, cascade=True)[:2]
    ws = backend.workspace[:64].view(torch.int32)
    assert int(ws[7]) == 2048 and int(ws[8]) == 3, (int(ws[7]), int(ws[8]))   # longest shared length, row tiles
    G.assert_close(got, _oracle(c), *tol, what=f"prefix groups {kv}")
    off = _run_full(c, cascade=False)[0]
    G.assert_close(got, off, *tol, what="prefix groups vs split-KV only")


def test_decode_cascade_duplicate_rows_and_large_batch():
    # the same request row twice in a batch (and hence a "shared prefix" of seq - 1 tokens): still the oracle's answer
    c = G.random_decode_case(32, 8, 128, [1400] * 24, seed=51, shared_prefix=1300)
    c["req_pool_indices"][1::2] = c["req_pool_indices"][0::2]          # 12 distinct rows, each listed twice
    got, backend = _run_full(c, cascade=True)[:2]
    assert _shared_len(backend) == 1280                                 # every pair still shares the 1300-token prompt
    G.assert_close(got, _oracle(c), G.BF16_ABS, G.BF16_REL, what="duplicate rows")
    # more requests than the single-CTA grouping handles (1024): the cascade switches itself off, results unchanged
    lens = [3 + (i % 29) for i in range(1100)]
    c = G.random_decode_case(4, 1, 128, lens, seed=52)
    got, backend = _run_full(c, cascade=True)[:2]
    assert _shared_len(backend) == 0
    G.assert_close(got, _oracle(c), G.BF16_ABS, G.BF16_REL, what="bs > 1024")
