"""-m gpu tests of the split-KV kernel's tile schedule with a dynamically claimed tail (DESIGN 3.2 "Dynamic tail"):
7/8 of a launch's KV tiles go to the CTAs as equal static ranges, the rest is cut into small chunks that CTAs claim
with one atomicAdd each as they run dry.  The product turns that on for launches of ~360 us and more whose slots the
plan finds handed out as runs (whole tiles by TMA: there the CTAs of a launch finish up to 8 % apart); here
``b200_debug_set_dyn_min_base`` lowers the threshold so that batches the CPU oracle can check run it -- with
one-tile chunks, i.e. as many chunk switches, partial flushes and claims as the schedule can produce."""
import pytest
import torch

from oracle import attn_oracle as O
from scratchpad_b200 import _lib

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import gpu_util as G
    from harness import checks

F16, FP8, BF16 = torch.float16, torch.float8_e4m3fn, torch.bfloat16


@pytest.fixture
def forced_dynamic():
    lib = _lib.load()
    prev = lib.b200_debug_set_dyn_min_base(8)
    yield lib
    lib.b200_debug_set_dyn_min_base(prev)


def _config(hq, hkv, d, kv_dtype):
    from scratchpad_b200.pool import kv_dtype_code
    cfg = _lib.AttnConfig()
    _lib.check(_lib.load().b200_attn_get_config(64, hq, hkv, d, d, kv_dtype_code(kv_dtype), -1, cfg), "get_config")
    return cfg


def _lens_for(cfg, tiles_per_cta, n_req, seed):
    """Ragged lengths whose tiles give every CTA of the persistent grid about `tiles_per_cta` tiles."""
    gen = torch.Generator().manual_seed(seed)
    total = cfg.grid_ctas * tiles_per_cta * cfg.tile_tokens
    w = torch.rand(n_req, generator=gen) + 0.2
    lens = [max(1, int(x)) for x in (w / w.sum() * total)]
    lens[1] = 1          # a one-token request in the middle
    lens[-1] += 3        # a ragged last tile
    return lens


def _hdr(backend):
    return backend.workspace[:64].view(torch.int32).tolist()


def _run(c, **kw):
    return G.run_decode_case(c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"], c["seq_lens"],
                             c["sm_scale"], 0.0, c["kv_dtype"], c.get("k_scale"), c.get("v_scale"), **kw)


def _oracle(c):
    return O.decode_ref(c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"], c["seq_lens"],
                        c["sm_scale"], 0.0, c.get("kv_dtype"), c.get("k_scale"), c.get("v_scale"))


SHAPES = [
    # hq, hkv, d
    (32, 8, 128),   # two head groups (blockIdx.y): two claim counters
    (32, 8, 64),
    (8, 1, 128),    # one kv head per CTA: token sub-slices combined inside the CTA before every chunk's flush
    (16, 1, 128),   # two q tiles per kv head
]


@pytest.mark.parametrize("hq,hkv,d", SHAPES)
@pytest.mark.parametrize("kv_dtype", [BF16, FP8, F16])
def test_dynamic_tail_matches_oracle_and_static_schedule(forced_dynamic, hq, hkv, d, kv_dtype):
    lib = forced_dynamic
    cfg = _config(hq, hkv, d, kv_dtype)
    lens = _lens_for(cfg, 10, 7, seed=hq + d)
    c = G.random_decode_case(hq, hkv, d, lens, seed=3 * hq + d, kv_dtype=kv_dtype, mode="contiguous")
    got, backend, runner, fb, layer = _run(c)
    h = _hdr(backend)
    grid, nch = h[5], h[2]
    assert nch > grid and h[9] >= 1 and h[10] == grid * h[1], f"the dynamic tail is not active: {h[:11]}"
    tol = {BF16: (G.BF16_ABS, G.BF16_REL), F16: (G.F16_ABS, G.F16_REL), FP8: (G.FP8_ABS, G.FP8_REL)}[kv_dtype]
    G.assert_close(got, _oracle(c), *tol, what=f"dynamic tail {hq}/{hkv}/{d}/{kv_dtype}")
    # three more launches on the same plan (the next layers of a model step): the merge has reset the claim counters
    q2 = c["q"].reshape(len(lens), -1).cuda()
    for _ in range(3):
        again = layer(q2, None, None, fb, save_kv_cache=False)
    torch.cuda.synchronize()
    G.assert_close(again.view(len(lens), hq, d).cpu(), got, 1e-3, 1e-3, what="repeated launches on one plan")
    # the static schedule (one equal range per CTA) on the same bytes
    lib.b200_debug_set_dyn_min_base(1 << 30)
    static, backend_s = _run(c)[:2]
    hs = _hdr(backend_s)
    assert hs[2] <= hs[5], "static schedule expected"
    G.assert_close(got, static, 2e-3, 1e-3, what="dynamic tail vs static ranges")
    lib.b200_debug_set_dyn_min_base(8)


@pytest.mark.parametrize("kv_dtype", [BF16, FP8])
@pytest.mark.parametrize("page_size,mode", [(1, "contiguous"), (16, "random"), (64, "random")])
def test_dynamic_tail_with_tma_tiles_and_paged_pools(forced_dynamic, kv_dtype, page_size, mode):
    lib = forced_dynamic
    hq, hkv, d = 32, 8, 128
    cfg = _config(hq, hkv, d, kv_dtype)
    lens = _lens_for(cfg, 9, 6, seed=page_size)
    c = G.random_decode_case(hq, hkv, d, lens, seed=77, kv_dtype=kv_dtype, mode=mode, page_size=page_size)
    lib.b200_debug_tma_tiles(1)
    got, backend = _run(c, page_size=page_size)[:2]
    assert lib.b200_debug_tma_tiles(1) > 0, "no tile took the TMA branch"
    h = _hdr(backend)
    assert h[2] > h[5], "the dynamic tail is not active"
    tol = (G.BF16_ABS, G.BF16_REL) if kv_dtype == BF16 else (G.FP8_ABS, G.FP8_REL)
    G.assert_close(got, _oracle(c), *tol, what=f"dynamic tail, TMA tiles, page {page_size}")


@pytest.mark.parametrize("kv_dtype", [BF16, FP8])
def test_dynamic_tail_with_shared_prefix_cascade_and_naive_kernel(forced_dynamic, kv_dtype):
    """The suffix tiles behind a cascaded shared prefix run the same schedule (positions relative to the prefix)."""
    hq, hkv, d, bs = 32, 8, 128, 24
    cfg = _config(hq, hkv, d, kv_dtype)
    per_req = cfg.grid_ctas * 10 * cfg.tile_tokens // bs
    gen = torch.Generator().manual_seed(9)
    lens = [1280 + per_req // 2 + int(x) for x in torch.randint(1, per_req, (bs,), generator=gen)]
    c = G.random_decode_case(hq, hkv, d, lens, seed=19, kv_dtype=kv_dtype, shared_prefix=1280, mode="contiguous")
    got, backend, runner, fb, layer = _run(c, cascade=True)
    h = _hdr(backend)
    assert h[7] == 1280 and h[2] > h[5], h[:11]
    pool = runner.token_to_kv_pool
    ks, vs = pool.get_kv_scales(0)
    naive = checks.naive_decode(c["q"].cuda(), pool.k_buffer[0], pool.v_buffer[0], runner.req_to_token_pool.req_to_token,
                                fb.req_pool_indices, fb.seq_lens, c["sm_scale"], 0.0, pool.kv_dtype_code(), ks, vs)
    torch.cuda.synchronize()
    tol = (G.BF16_ABS, G.BF16_REL) if kv_dtype == BF16 else (G.FP8_ABS, G.FP8_REL)
    G.assert_close(got, naive.cpu(), *tol, what="dynamic tail + cascade vs naive kernel")
    G.assert_close(got, _oracle(c), *tol, what="dynamic tail + cascade vs oracle")


def test_dynamic_tail_under_cuda_graph_replay(forced_dynamic):
    """Captured launches claim from counters that every merge leaves at zero: replaying a graph of several layer-calls
    (one plan, as a model step) gives the same result every time."""
    from scratchpad_b200 import CudaB200AttnBackend
    from harness import ForwardBatch, ForwardMode, synth
    hq, hkv, d, bs = 32, 8, 128, 8
    cfg = _config(hq, hkv, d, BF16)
    lens = _lens_for(cfg, 10, bs, seed=5)
    c = G.random_decode_case(hq, hkv, d, lens, seed=41, mode="contiguous")
    runner = G.make_runner_from_case(c["k_buf"], c["v_buf"], c["req_to_token"], BF16, tp_hq=hq)
    backend = CudaB200AttnBackend(runner)
    layer = synth.make_layer(runner)
    backend.init_cuda_graph_state(16)
    s_rpi = torch.zeros(16, dtype=torch.int32, device="cuda")
    s_sl = torch.full((16,), 1, dtype=torch.int32, device="cuda")
    s_q = torch.zeros((bs, hq * d), dtype=BF16, device="cuda")
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
        outs = [layer(s_q, None, None, fb, save_kv_cache=False) for _ in range(3)]  # three layer-calls, one plan
    s_rpi[:bs].copy_(c["req_pool_indices"].to(torch.int32))
    s_sl[:bs].copy_(torch.tensor(lens, dtype=torch.int32))
    s_q.copy_(c["q"].reshape(bs, -1))
    want = _oracle(c)
    for rep in range(3):
        backend.init_forward_metadata_replay_cuda_graph(bs, s_rpi, s_sl, sum(lens), None, ForwardMode.DECODE, None,
                                                        seq_lens_cpu=None)
        graph.replay()
        torch.cuda.synchronize()
        h = _hdr(backend)
        assert h[2] > h[5], "the dynamic tail is not active"
        for i, o in enumerate(outs):
            G.assert_close(o.view(bs, hq, d).cpu(), want, G.BF16_ABS, G.BF16_REL, what=f"replay {rep}, layer-call {i}")


def test_scattered_slots_keep_the_static_schedule(forced_dynamic):
    """The plan probes the first slots of every request: where the allocator scattered them (row gather, CTAs finish
    < 4 % apart) the chunk switches would cost what the balance gains, so the ranges stay static."""
    hq, hkv, d = 32, 8, 128
    cfg = _config(hq, hkv, d, BF16)
    lens = _lens_for(cfg, 10, 7, seed=2)
    for mode, dynamic in (("random", False), ("contiguous", True)):
        c = G.random_decode_case(hq, hkv, d, lens, seed=11, mode=mode)
        got, backend = _run(c)[:2]
        h = _hdr(backend)
        assert (h[2] > h[5]) == dynamic, (mode, h[:11])
        G.assert_close(got, _oracle(c), G.BF16_ABS, G.BF16_REL, what=f"{mode} slots")
