"""-m gpu tests of the C ABI's failure behaviour and host-side caches (VERDICT r1 "Harden the C ABI"):
an unplanned workspace is refused (B200_ERR_NOT_PLANNED), a plan header that does not describe the batch poisons
the output instead of being read, the per-kernel attributes are set per device, and the cached tensor maps of a
paged pool are keyed by buffer."""
import pytest
import torch

from oracle import attn_oracle as O
from scratchpad_b200 import _lib

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import gpu_util as G
    from harness import ForwardBatch, ForwardMode, synth
    from scratchpad_b200 import CudaB200AttnBackend


def _setup(lens, hq=8, hkv=2, d=128, seed=5, cascade=False, page_size=1):
    c = G.random_decode_case(hq, hkv, d, lens, seed=seed, page_size=page_size)
    runner = G.make_runner_from_case(c["k_buf"], c["v_buf"], c["req_to_token"], torch.bfloat16, tp_hq=hq,
                                     page_size=page_size)
    backend = CudaB200AttnBackend(runner, cascade=cascade)
    layer = synth.make_layer(runner, 0)
    b = len(lens)
    fb = ForwardBatch(forward_mode=ForwardMode.DECODE, batch_size=b,
                      req_pool_indices=c["req_pool_indices"].cuda(), seq_lens=c["seq_lens"].cuda(),
                      out_cache_loc=torch.zeros(b, dtype=torch.int64, device="cuda"), seq_lens_sum=sum(lens),
                      req_to_token_pool=runner.req_to_token_pool, token_to_kv_pool=runner.token_to_kv_pool,
                      attn_backend=backend)
    q = c["q"].reshape(b, hq * d).cuda()
    return c, backend, layer, fb, q


def _oracle(c):
    return O.decode_ref(c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"], c["seq_lens"],
                        c["sm_scale"])


def test_decode_on_unplanned_workspace_is_refused():
    c, backend, layer, fb, q = _setup([40, 3, 200])
    backend.init_forward_metadata(fb)
    # a workspace of the right size that b200_attn_plan has never seen
    backend.workspace = torch.zeros_like(backend.workspace)
    with pytest.raises(_lib.B200AttnError) as e:
        layer(q, None, None, fb, save_kv_cache=False)
    assert e.value.code == _lib.ERR_NOT_PLANNED
    # planning it makes the same call succeed
    backend.init_forward_metadata(fb)
    got = layer(q, None, None, fb, save_kv_cache=False).view(3, 8, 128).cpu()
    G.assert_close(got, _oracle(c), G.BF16_ABS, G.BF16_REL, "after plan")


def test_plan_for_another_cascade_setting_is_refused():
    c, backend, layer, fb, q = _setup([40, 3, 200, 17])
    backend.init_forward_metadata(fb)
    backend.cascade = 1  # the plan in the workspace was made with cascade off
    with pytest.raises(_lib.B200AttnError) as e:
        layer(q, None, None, fb, save_kv_cache=False)
    assert e.value.code == _lib.ERR_NOT_PLANNED


def test_overwritten_plan_header_poisons_the_output():
    """The host-side record says 'planned', but the header on the device no longer describes this batch (somebody
    overwrote the workspace; a graph replayed without its plan refresh): the kernels refuse to read it."""
    c, backend, layer, fb, q = _setup([40, 3, 200])
    backend.init_forward_metadata(fb)
    good = layer(q, None, None, fb, save_kv_cache=False).clone()
    backend.workspace[:64].zero_()
    bad = layer(q, None, None, fb, save_kv_cache=False)
    torch.cuda.synchronize()
    assert torch.isnan(bad.float()).all(), "an output computed from a missing plan must not look valid"
    backend.init_forward_metadata(fb)
    again = layer(q, None, None, fb, save_kv_cache=False)
    assert torch.equal(again, good)


def test_cached_tensor_maps_follow_the_buffer():
    """Paged pool (TMA box copies): the second layer-call reuses the encoded maps; another pool at another address
    (and the same pool after its bytes change) must still be read correctly."""
    for seed in (11, 12):
        c, backend, layer, fb, q = _setup([300, 17, 129, 64], hq=32, hkv=8, d=128, seed=seed, page_size=16)
        backend.init_forward_metadata(fb)
        ref = _oracle(c)
        for _ in range(2):
            got = layer(q, None, None, fb, save_kv_cache=False).view(4, 32, 128).cpu()
            G.assert_close(got, ref, G.BF16_ABS, G.BF16_REL, f"paged seed {seed}")
        assert backend.lib.b200_debug_last_tma_mode() == 1
        # same buffers, new bytes: the cached map describes addresses, not contents
        pool = fb.token_to_kv_pool
        pool.v_buffer[0].mul_(0.5)
        got = layer(q, None, None, fb, save_kv_cache=False).view(4, 32, 128).cpu()
        G.assert_close(got, ref * 0.5, G.BF16_ABS, G.BF16_REL, "paged, halved V")


def test_workspace_of_a_large_request_pool_stays_small():
    """the cascade partials are sized for the batches the cascade can run on (<= 1024), not for max_bs."""
    lib = _lib.load()
    small, big = _lib.AttnConfig(), _lib.AttnConfig()
    _lib.check(lib.b200_attn_get_config(1024, 32, 8, 128, 128, _lib.KV_BF16, -1, small), "get_config")
    _lib.check(lib.b200_attn_get_config(4096, 32, 8, 128, 128, _lib.KV_BF16, -1, big), "get_config")
    cas_1024 = 8 * 1024 * 32 * (128 + 1) * 4
    # growing max_bs 4x grows only the split-KV partial slots (one per extra request row), not the cascade's
    per_row = 32 * (128 + 1) * 4  # one partial per (cta, request) and q head
    assert big.workspace_bytes - small.workspace_bytes < 3072 * per_row + (1 << 20)
    # split-KV partials: one per (chunk, request); at most 5 chunks per CTA (+1) when the tail of a long launch is
    # claimed dynamically
    assert big.workspace_bytes < cas_1024 + (5 * big.grid_ctas + 1 + 4096) * per_row + (4 << 20)


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_devices_in_one_process():
    """The >48 KiB shared-memory attribute and the occupancy are per device: a second GPU driven from the same
    process must launch the same kernels (decode, paged decode, extend on both families)."""
    for dev in (0, 1, 0):
        with torch.cuda.device(dev):
            c = G.random_decode_case(32, 8, 128, [300, 17, 1, 129], seed=20 + dev)
            got = G.run_decode_case(c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"],
                                    c["seq_lens"], c["sm_scale"])[0]
            G.assert_close(got, _oracle(c), G.BF16_ABS, G.BF16_REL, f"decode on cuda:{dev}")
            c = G.random_decode_case(32, 8, 128, [300, 17, 1, 129], seed=30 + dev, page_size=16)
            got = G.run_decode_case(c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"],
                                    c["seq_lens"], c["sm_scale"], page_size=16)[0]
            G.assert_close(got, _oracle(c), G.BF16_ABS, G.BF16_REL, f"paged decode on cuda:{dev}")
            for path in (0, 1):
                e = G.random_extend_case(8, 2, 128, [200, 0], [140, 130], seed=40 + dev)
                got = G.run_extend_case(e, force_path=path)[0]
                ref = O.extend_ref(e["q"], e["k_new"], e["v_new"], e["k_buf"], e["v_buf"], e["req_to_token"],
                                   e["req_pool_indices"], e["seq_lens"], e["extend_seq_lens"], e["extend_start_loc"],
                                   e["sm_scale"], 0.0)
                G.assert_close(got, ref, G.BF16_ABS, G.BF16_REL, f"extend path {path} on cuda:{dev}")


def test_workspace_cannot_grow_under_captured_graphs():
    """graphs captured earlier hold the workspace's address and a layout that depends on max_bs; a larger
    batch after capture must be refused, not served from a re-allocated workspace."""
    c, backend, layer, fb, q = _setup([40, 3, 200])
    small = backend.max_bs
    backend.init_cuda_graph_state(small)
    backend.init_forward_metadata_capture_cuda_graph(3, 3, fb.req_pool_indices, fb.seq_lens, None, ForwardMode.DECODE, None)
    ws = backend.workspace.data_ptr()
    big = ForwardBatch(forward_mode=ForwardMode.DECODE, batch_size=small + 1,
                       req_pool_indices=torch.zeros(small + 1, dtype=torch.int64, device="cuda"),
                       seq_lens=torch.ones(small + 1, dtype=torch.int64, device="cuda"),
                       out_cache_loc=torch.zeros(small + 1, dtype=torch.int64, device="cuda"), seq_lens_sum=small + 1,
                       req_to_token_pool=fb.req_to_token_pool, token_to_kv_pool=fb.token_to_kv_pool, attn_backend=backend)
    with pytest.raises(RuntimeError, match="after CUDA-graph capture"):
        backend.init_forward_metadata(big)
    assert backend.workspace.data_ptr() == ws and backend.max_bs == small
    # the captured batch size still works
    backend.init_forward_metadata(fb)
    got = layer(q, None, None, fb, save_kv_cache=False).view(3, 8, 128).cpu()
    G.assert_close(got, _oracle(c), G.BF16_ABS, G.BF16_REL, "after the refused growth")
