"""-m gpu parity tests of decode over a paged pool (page_size 16..128): whole KV tiles arrive by
TMA box copies, ragged request tails by the row gather.  Same oracle as page size 1 -- req_to_token still holds one
slot id per token -- so every case is also compared with the page_size=1 kernel on the same bytes."""
import pytest
import torch

from oracle import attn_oracle as O
from scratchpad_b200 import CudaB200AttnBackend, _lib
from harness import ForwardBatch, ForwardMode, synth

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import gpu_util as G


def _oracle(c, logit_cap=0.0):
    return O.decode_ref(c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"], c["seq_lens"],
                        c["sm_scale"], logit_cap, c.get("kv_dtype"), c.get("k_scale"), c.get("v_scale"))


def _run(c, page_size, logit_cap=0.0, **kw):
    return G.run_decode_case(c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"],
                             c["seq_lens"], c["sm_scale"], logit_cap, c.get("kv_dtype", torch.bfloat16),
                             c.get("k_scale"), c.get("v_scale"), page_size=page_size, **kw)


SHAPES = [
    # hq, hkv, d -> (heads per CTA, tile tokens): every (segment size, swizzle mode, boxes-per-tile) combination
    (32, 8, 64),    # bf16: 128-byte rows, HPC 8, tile 16 | fp8: 64-byte rows (SWIZZLE_64B)
    (4, 1, 64),     # HPC 1, tile 128 tokens: several page boxes per tile
    (32, 8, 128),   # bf16: 256-byte segments as two column halves, HPC 4, tile 16 | fp8: HPC 8
    (8, 1, 128),    # 70B TP8 shard: bf16 tile 64, fp8 tile 128
    (8, 2, 128),
    (6, 2, 64),     # g = 3
]
LENS = [[1, 15, 16, 17], [127, 128, 129, 300, 64], [513, 1000, 2, 256, 255, 257]]


@pytest.mark.parametrize("hq,hkv,d", SHAPES)
@pytest.mark.parametrize("kv", ["bf16", "fp8"])
@pytest.mark.parametrize("page_size", [16, 64, 128])
def test_paged_decode_vs_oracle_and_token_granular_kernel(hq, hkv, d, kv, page_size):
    kv_dtype = torch.bfloat16 if kv == "bf16" else torch.float8_e4m3fn
    tol = (G.BF16_ABS, G.BF16_REL) if kv == "bf16" else (G.FP8_ABS, G.FP8_REL)
    for i, lens in enumerate(LENS):
        c = G.random_decode_case(hq, hkv, d, lens, seed=100 * hq + 10 * hkv + i + page_size, kv_dtype=kv_dtype,
                                 page_size=page_size)
        lib = _lib.load()
        lib.b200_debug_tma_tiles(1)
        got, backend = _run(c, page_size)[:2]
        # page >= tile: one TMA box per K (V) tile; smaller pages: the row gather over the paged table
        tn = backend.config.tile_tokens
        assert lib.b200_debug_last_tma_mode() == (1 if page_size >= tn else 0)
        # ... and the TMA branch is really taken: every full tile of a paged table has consecutive slots.  A CTA picks
        # its loop (TMA-capable / row gather only) from its first tile, so with one tile per CTA the count is exact
        n_tma = lib.b200_debug_tma_tiles(1)
        full_tiles = sum(n // tn for n in lens) * (hkv // backend.config.heads_per_cta)
        tiles_per_cta = int(backend.workspace[:64].view(torch.int32)[1].item())
        if page_size >= tn:
            assert n_tma <= full_tiles and (n_tma == full_tiles if tiles_per_cta == 1 else n_tma > 0 or full_tiles == 0), \
                (n_tma, full_tiles, tiles_per_cta)
        else:
            assert n_tma == 0
        G.assert_close(got, _oracle(c), *tol, what=f"paged {page_size} {hq}/{hkv}/{d}/{kv}/{lens}")
        # the same bytes through the row-gather kernel (page_size 1 over the same buffers): identical arithmetic on
        # identical tiles -> bit-identical output
        cc = dict(c)
        plain = G.run_decode_case(c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"], c["seq_lens"],
                                  c["sm_scale"], 0.0, kv_dtype, c.get("k_scale"), c.get("v_scale"),
                                  tma_contiguous_runs=False)[0]
        assert _lib.load().b200_debug_last_tma_mode() == 0
        assert torch.equal(got.view(torch.int16), plain.view(torch.int16)), f"paged {page_size} != row gather"


def test_paged_decode_page_32_int32_indices_logit_cap_qkv_slice():
    c = G.random_decode_case(32, 8, 128, [700, 33, 1, 320], seed=5, page_size=32)
    got = _run(c, 32, logit_cap=30.0, idx_dtype=torch.int32, q_slice_of_qkv=True)[0]
    G.assert_close(got, _oracle(c, 30.0), G.BF16_ABS, G.BF16_REL, what="paged 32")


@pytest.mark.parametrize("kv", ["bf16", "fp8"])
def test_paged_decode_many_ctas_long_mixed_contexts(kv):
    """Enough tiles that every CTA of the persistent grid streams several TMA stages and crosses request
    boundaries (ring reuse: TMA after LDGSTS after TMA on the same stage)."""
    kv_dtype = torch.bfloat16 if kv == "bf16" else torch.float8_e4m3fn
    tol = (G.BF16_ABS, G.BF16_REL) if kv == "bf16" else (G.FP8_ABS, G.FP8_REL)
    gen = torch.Generator().manual_seed(9)
    lens = [int(x) for x in torch.randint(1, 3000, (48,), generator=gen)]
    c = G.random_decode_case(8, 1, 128, lens, seed=77, kv_dtype=kv_dtype, page_size=16)
    got = _run(c, 16)[0]
    G.assert_close(got, _oracle(c), *tol, what=f"paged long {kv}")


@pytest.mark.parametrize("kv", ["bf16", "fp8"])
def test_paged_decode_with_shared_prefix_cascade(kv):
    """Page-aligned shared prefix (paged radix matching shares whole pages): dense tcgen05 leg over the shared
    pages, TMA split-KV over the private suffixes starting at a 128-token (= page multiple) offset."""
    kv_dtype = torch.bfloat16 if kv == "bf16" else torch.float8_e4m3fn
    tol = (G.BF16_ABS, G.BF16_REL) if kv == "bf16" else (G.FP8_ABS, G.FP8_REL)
    shared = 1344  # 84 pages of 16; the dense leg takes 1280
    gen = torch.Generator().manual_seed(3)
    lens = [shared + int(x) for x in torch.randint(1, 300, (64,), generator=gen)]
    c = G.random_decode_case(32, 8, 128, lens, seed=15, kv_dtype=kv_dtype, shared_prefix=shared, page_size=16)
    got, backend = _run(c, 16, cascade=True)[:2]
    assert int(backend.workspace[:64].view(torch.int32)[7].item()) == 1280, "the shared prefix was not detected"
    G.assert_close(got, _oracle(c), *tol, what=f"paged cascade {kv}")


def test_paged_store_then_decode_with_allocator_and_graph_replay():
    """The paged allocator hands out the decode slots, b200_kv_store writes them, and a captured graph is replayed
    for two successive decode steps (tensor maps ride in the captured kernel parameters)."""
    from scratchpad_b200 import PagedTokenToKVPoolAllocator
    hq, hkv, d, bs, ps = 32, 8, 64, 6, 16
    lens0 = [100, 16, 47, 1, 255, 32]
    c = G.random_decode_case(hq, hkv, d, [l + 2 for l in lens0], seed=23, page_size=ps)
    runner = G.make_runner_from_case(c["k_buf"], c["v_buf"], c["req_to_token"], torch.bfloat16, tp_hq=hq, page_size=ps)
    backend = CudaB200AttnBackend(runner)
    layer = synth.make_layer(runner)
    backend.init_cuda_graph_state(8)
    s_rpi = c["req_pool_indices"].to(torch.int32).cuda()
    s_sl = torch.ones(bs, dtype=torch.int32, device="cuda")
    s_loc = torch.zeros(bs, dtype=torch.int64, device="cuda")
    s_q = torch.zeros((bs, hq * d), dtype=torch.bfloat16, device="cuda")
    s_k = torch.zeros((bs, hkv * d), dtype=torch.bfloat16, device="cuda")
    s_v = torch.zeros_like(s_k)
    fb = ForwardBatch(ForwardMode.DECODE, bs, s_rpi, s_sl, s_loc, req_to_token_pool=runner.req_to_token_pool,
                      token_to_kv_pool=runner.token_to_kv_pool, attn_backend=backend)
    backend.init_forward_metadata_capture_cuda_graph(bs, bs, s_rpi, s_sl, None, ForwardMode.DECODE, None)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        layer(s_q, s_k, s_v, fb)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        s_o = layer(s_q, s_k, s_v, fb)
    # restore the pool bytes the warm-up stores touched (slot 0)
    runner.token_to_kv_pool.k_buffer[0].copy_(c["k_buf"])
    runner.token_to_kv_pool.v_buffer[0].copy_(c["v_buf"])
    kb, vb = c["k_buf"].clone(), c["v_buf"].clone()
    gen = torch.Generator().manual_seed(1)
    for step in (1, 2):
        lens = [l + step for l in lens0]
        # the slot of the new position as the table (built by the paged allocator logic) holds it
        loc = torch.stack([c["req_to_token"][int(c["req_pool_indices"][i]), lens[i] - 1] for i in range(bs)]).long()
        pos = torch.tensor(lens) - 1
        assert torch.equal(loc % ps, pos % ps)  # the allocator invariant the TMA path relies on
        k_new = torch.randn((bs, hkv, d), generator=gen).to(torch.bfloat16)
        v_new = torch.randn((bs, hkv, d), generator=gen).to(torch.bfloat16)
        kb[loc], vb[loc] = k_new, v_new
        s_sl.copy_(torch.tensor(lens, dtype=torch.int32))
        s_loc.copy_(loc)
        s_q.copy_(c["q"].reshape(bs, -1))
        s_k.copy_(k_new.reshape(bs, -1))
        s_v.copy_(v_new.reshape(bs, -1))
        backend.init_forward_metadata_replay_cuda_graph(bs, s_rpi, s_sl, sum(lens), None, ForwardMode.DECODE, None)
        graph.replay()
        torch.cuda.synchronize()
        ref = O.decode_ref(c["q"], kb, vb, c["req_to_token"], c["req_pool_indices"], torch.tensor(lens), c["sm_scale"])
        G.assert_close(s_o.view(bs, hq, d).cpu(), ref, G.BF16_ABS, G.BF16_REL, what=f"paged graph step {step}")
    # PagedTokenToKVPoolAllocator.alloc_decode produces slots with the same invariant
    alloc = PagedTokenToKVPoolAllocator(runner.token_to_kv_pool.size, ps, torch.bfloat16, "cpu", runner.token_to_kv_pool)
    new = alloc.alloc_decode(torch.tensor([17, 18]), torch.tensor([99, 5 * ps]))
    assert new[0] % ps == 0 and new[1] == 5 * ps + 1


def test_unsupported_page_sizes_raise():
    c = G.random_decode_case(8, 2, 64, [40], seed=1)
    runner = G.make_runner_from_case(c["k_buf"], c["v_buf"], c["req_to_token"], torch.bfloat16, tp_hq=8)
    runner.token_to_kv_pool.page_size = 8  # not a supported page size
    backend = CudaB200AttnBackend(runner)
    layer = synth.make_layer(runner)
    fb = ForwardBatch(ForwardMode.DECODE, 1, c["req_pool_indices"].cuda(), c["seq_lens"].cuda(),
                      torch.zeros(1, dtype=torch.int64, device="cuda"), req_to_token_pool=runner.req_to_token_pool,
                      token_to_kv_pool=runner.token_to_kv_pool, attn_backend=backend)
    backend.init_forward_metadata(fb)
    with pytest.raises(_lib.B200AttnError):
        layer(c["q"].reshape(1, -1).cuda(), None, None, fb, save_kv_cache=False)


def test_paged_contract_check_catches_a_token_granular_table(monkeypatch):
    """B200_CHECK_PAGES=1: a table that is not page-structured (random token slots) is rejected at metadata time
    instead of being read through page copies; a well-formed one passes."""
    monkeypatch.setenv("B200_CHECK_PAGES", "1")
    good = G.random_decode_case(8, 2, 64, [100, 33], seed=2, page_size=16)
    _run(good, 16)
    bad = G.random_decode_case(8, 2, 64, [100, 33], seed=2)             # token-granular random slots ...
    pad = torch.zeros((15,) + tuple(bad["k_buf"].shape[1:]), dtype=bad["k_buf"].dtype)
    bad["k_buf"], bad["v_buf"] = torch.cat([bad["k_buf"], pad]), torch.cat([bad["v_buf"], pad])
    with pytest.raises(RuntimeError, match="layout contract"):
        _run(bad, 16)                                                   # ... declared as a 16-token-page pool


@pytest.mark.parametrize("kv", ["bf16", "fp8"])
@pytest.mark.parametrize("force_path", [_lib.PATH_MMA, _lib.PATH_TC, _lib.PATH_TC2])
def test_extend_over_a_paged_pool(kv, force_path):
    """The extend kernels gather the cached prefix by token slot and the KV store scatters by out_cache_loc, so a
    paged table needs nothing new from them: same oracle, pool rows written at the paged slots."""
    if force_path == _lib.PATH_TC2 and kv != "bf16":
        pytest.skip("the CTA-pair kernel is built for bf16 pools")
    kv_dtype = torch.bfloat16 if kv == "bf16" else torch.float8_e4m3fn
    tol = (G.BF16_ABS, G.BF16_REL) if kv == "bf16" else (G.FP8_ABS, G.FP8_REL)
    c = G.random_extend_case(8, 2, 128, [130, 0, 77, 256], [140, 150, 3, 129], seed=61, kv_dtype=kv_dtype, page_size=16)
    got, runner = G.run_extend_case(c, force_path=force_path)
    ref = O.extend_ref(c["q"], c["k_new"], c["v_new"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"],
                       c["seq_lens"], c["extend_seq_lens"], c["extend_start_loc"], c["sm_scale"], 0.0, kv_dtype,
                       c["k_scale"], c["v_scale"])
    G.assert_close(got, ref, *tol, what=f"paged extend {kv}")
    kb, vb = c["k_buf"].clone(), c["v_buf"].clone()
    O.set_kv_buffer_ref(kb, vb, c["out_cache_loc"], c["k_new"].clone(), c["v_new"].clone(), kv_dtype, c["k_scale"], c["v_scale"])
    assert torch.equal(runner.token_to_kv_pool.k_buffer[0].cpu(), kb) and torch.equal(runner.token_to_kv_pool.v_buffer[0].cpu(), vb)


@pytest.mark.parametrize("kv", ["bf16", "fp8"])
def test_wrong_page_hint_costs_speed_not_correctness(kv):
    """The kernel checks every tile's slot ids before fetching it as page runs: a token-granular table with random
    slots declared as a 16-token-page pool is still read correctly (every tile falls back to the row gather), and so
    is a table where only some runs are consecutive."""
    kv_dtype = torch.bfloat16 if kv == "bf16" else torch.float8_e4m3fn
    tol = (G.BF16_ABS, G.BF16_REL) if kv == "bf16" else (G.FP8_ABS, G.FP8_REL)
    c = G.random_decode_case(32, 8, 64, [100, 333, 16, 64, 257], seed=12, kv_dtype=kv_dtype)   # random token slots
    pad = torch.zeros((15,) + tuple(c["k_buf"].shape[1:]), dtype=c["k_buf"].dtype)
    c["k_buf"], c["v_buf"] = torch.cat([c["k_buf"], pad]), torch.cat([c["v_buf"], pad])
    G.assert_close(_run(c, 16)[0], _oracle(c), *tol, what="random slots, page hint 16")
    # a paged table with a few entries swapped inside pages: those tiles must take the row gather, the rest TMA
    c = G.random_decode_case(32, 8, 64, [100, 333, 16, 64, 257], seed=13, kv_dtype=kv_dtype, page_size=16)
    row = int(c["req_pool_indices"][1])
    for t in (5, 40, 41, 200, 319):
        a, b = c["req_to_token"][row, t].item(), c["req_to_token"][row, t + 1].item()
        c["req_to_token"][row, t], c["req_to_token"][row, t + 1] = b, a
    G.assert_close(_run(c, 16)[0], _oracle(c), *tol, what="paged table with swapped entries")


def test_token_granular_pool_with_consecutive_slots_uses_tma():
    """PAGE_AUTO (tma_contiguous_runs=True): a page_size=1 pool whose allocator handed out consecutive runs
    (pool.py:240-253 on an unfragmented free list) gets whole-tile TMA copies; output identical to the row gather."""
    gen = torch.Generator().manual_seed(3)
    lens = [int(x) for x in torch.randint(1, 900, (24,), generator=gen)]
    c = G.random_decode_case(32, 8, 128, lens, seed=14, mode="contiguous")
    _lib.load().b200_debug_tma_tiles(1)
    got = _run(c, 1, tma_contiguous_runs=True)[0]
    assert _lib.load().b200_debug_last_tma_mode() == 1
    assert _lib.load().b200_debug_tma_tiles(1) >= sum(n // 16 for n in lens) // 2, "consecutive slots must be fetched by TMA"
    G.assert_close(got, _oracle(c), G.BF16_ABS, G.BF16_REL, what="auto TMA on consecutive slots")
    plain = _run(c, 1, tma_contiguous_runs=False)[0]
    assert _lib.load().b200_debug_last_tma_mode() == 0
    # and it is the default: a backend built without the argument fetches consecutive slots by TMA
    _lib.load().b200_debug_tma_tiles(1)
    dflt = _run(c, 1)[0]
    assert _lib.load().b200_debug_last_tma_mode() == 1 and _lib.load().b200_debug_tma_tiles(1) > 0
    assert torch.equal(dflt.view(torch.int16), plain.view(torch.int16))
    assert torch.equal(got.view(torch.int16), plain.view(torch.int16))
    c = G.random_decode_case(32, 8, 128, lens, seed=15)                     # random slots: nothing consecutive
    G.assert_close(_run(c, 1, tma_contiguous_runs=True)[0], _oracle(c), G.BF16_ABS, G.BF16_REL, what="auto TMA, random slots")
    assert _lib.load().b200_debug_tma_tiles(1) == 0, "scattered slots must all take the row gather"


@pytest.mark.parametrize("hq,hkv,d,kv,page_size", [(32, 8, 64, "bf16", 16), (8, 1, 128, "fp8", 128), (8, 1, 128, "bf16", 64),
                                                   (32, 8, 128, "fp8", 32)])
def test_paged_decode_over_the_fused_kv_layout(hq, hkv, d, kv, page_size, monkeypatch):
    """Fused [P + page, 2, Hkv, D] pool (K and V rows of a token adjacent): the tensor maps take the doubled token
    stride and the V view's base pointer; whole-tile boxes and per-run boxes both land where the row gather would."""
    monkeypatch.setattr(G, "FUSED_KV", True)
    kv_dtype = torch.bfloat16 if kv == "bf16" else torch.float8_e4m3fn
    tol = (G.BF16_ABS, G.BF16_REL) if kv == "bf16" else (G.FP8_ABS, G.FP8_REL)
    c = G.random_decode_case(hq, hkv, d, [300, 128, 17, 640, 129], seed=31 + page_size, kv_dtype=kv_dtype, page_size=page_size)
    got, backend, runner = _run(c, page_size)[:3]
    assert runner.token_to_kv_pool.fused_kv and _lib.load().b200_debug_last_tma_mode() in (0, 1)
    G.assert_close(got, _oracle(c), *tol, what=f"paged fused {hq}/{hkv}/{d}/{kv}/{page_size}")
