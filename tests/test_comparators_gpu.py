"""Parity against the reference's own GPU attention path on the same pool bytes (SURVEY.md 8d "GPU comparators";
VERDICT r1 items 1-2): the UNMODIFIED ``FlashInferAttnBackend`` of the reference over the installed FlashInfer, and
the reference's in-tree Triton kernels -- both loaded from ``baseline/_ref`` (``baseline/ref_loader.py``).

This is also the pin of the fp8-e4m3 KV semantics: FlashInfer takes ``k_scale`` / ``v_scale`` on e4m3 pools exactly
as the reference passes them (``flashinfer_backend.py:387-388,457-458,475-482``), an implementation independent of
this repo's oracle."""
import pytest
import torch

import gpu_util as G

pytestmark = pytest.mark.gpu

from baseline import ref_loader  # noqa: E402

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="baseline/_ref not installed (python baseline/install_ref.py)")


def _flashinfer_ok():
    try:
        import flashinfer  # noqa: F401
        return True
    except Exception:
        return False


needs_fi = pytest.mark.skipif(not _flashinfer_ok(), reason="flashinfer is not importable")


def _ref_state(runner, hq, kv_dtype, k_scale=None, v_scale=None, logit_cap=0.0, sm_scale=None):
    from baseline import comparators as C

    pool = runner.token_to_kv_pool
    return C.RefState(req_to_token=runner.req_to_token_pool.req_to_token, k_buffers=pool.k_buffer,
                      v_buffers=pool.v_buffer, kv_dtype=kv_dtype, hq=hq, hkv=pool.head_num, head_dim=pool.head_dim,
                      k_scale=k_scale, v_scale=v_scale, logit_cap=logit_cap, sm_scale=sm_scale)


DECODE_SHAPES = [
    # hq, hkv, d, seq_lens
    (32, 8, 64, [2048, 1, 17, 300, 1025]),       # Llama-3.2-1B heads
    (32, 8, 128, [4100, 129, 64, 2, 999, 31]),   # Llama-3.1-8B heads
    (8, 1, 128, [1500, 3000, 77]),               # 70B TP8 shard, group 8 -> the reference picks tensor cores
]


@needs_ref
@needs_fi
@pytest.mark.parametrize("hq,hkv,d,lens", DECODE_SHAPES)
@pytest.mark.parametrize("kv_dtype", [torch.bfloat16, torch.float8_e4m3fn])
def test_decode_matches_reference_flashinfer_backend(hq, hkv, d, lens, kv_dtype):
    from baseline import comparators as C
    from oracle import attn_oracle as O

    fp8 = kv_dtype != torch.bfloat16
    c = G.random_decode_case(hq, hkv, d, lens, seed=11 + d + hq, kv_dtype=kv_dtype, per_head=False)
    ours, backend, runner, fb, layer = G.run_decode_case(
        c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"], c["seq_lens"], c["sm_scale"], 0.0,
        kv_dtype, c["k_scale"], c["v_scale"])
    ks = float(c["k_scale"][0]) if fp8 else None
    vs = float(c["v_scale"][0]) if fp8 else None
    st = _ref_state(runner, hq, kv_dtype, ks, vs, sm_scale=c["sm_scale"])
    ref = C.RefFlashInfer(st, context_len=runner.model_config.context_len)
    assert ref.use_tensor_cores == (fp8 or hq // hkv > 4)  # flashinfer_backend.py:936-981
    rfb = st.decode_batch(fb.req_pool_indices, fb.seq_lens, fb.out_cache_loc, ref.backend)
    ref.init_forward_metadata(rfb)
    q2 = c["q"].reshape(len(lens), hq * d).cuda()
    theirs = ref.forward(0, q2, None, None, rfb, save_kv_cache=False)
    torch.cuda.synchronize()
    theirs = theirs.view(len(lens), hq, d).float().cpu()
    tol = (G.FP8_ABS, G.FP8_REL) if fp8 else (G.BF16_ABS, G.BF16_REL)
    mx, rl = G.assert_close(ours, theirs, *tol, what=f"ours vs reference FlashInfer backend ({kv_dtype})")
    # and both against the fp32 oracle on the same bytes: the three agree pairwise
    want = O.decode_ref(c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"], c["seq_lens"],
                        c["sm_scale"], 0.0, kv_dtype, c["k_scale"], c["v_scale"])
    G.assert_close(theirs, want, *tol, what="reference FlashInfer backend vs oracle")
    G.assert_close(ours, want, *tol, what="ours vs oracle")
    # element-wise, not only in norm: |ours - theirs| <= abs_tol + rel_tol * |theirs|
    assert ((ours.float() - theirs).abs() <= tol[0] + tol[1] * theirs.abs()).all()


@needs_ref
@pytest.mark.parametrize("hq,hkv,d,lens", DECODE_SHAPES[:2])
@pytest.mark.parametrize("cap", [0.0, 30.0])
def test_decode_matches_reference_triton_kernel(hq, hkv, d, lens, cap):
    from baseline import comparators as C

    c = G.random_decode_case(hq, hkv, d, lens, seed=5 + d)
    ours, backend, runner, fb, layer = G.run_decode_case(
        c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"], c["seq_lens"], c["sm_scale"], cap)
    pool = runner.token_to_kv_pool
    theirs = C.triton_decode(c["q"].cuda(), pool.k_buffer[0], pool.v_buffer[0], runner.req_to_token_pool.req_to_token,
                             fb.req_pool_indices, fb.seq_lens, c["sm_scale"], cap)
    torch.cuda.synchronize()
    G.assert_close(ours, theirs.float().cpu(), G.BF16_ABS, G.BF16_REL, what="ours vs in-tree Triton decode")


@needs_fi
@pytest.mark.parametrize("page_size", [16, 64])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_paged_decode_matches_flashinfer_trtllm_gen(page_size, dtype):
    """SURVEY 8d stretch comparator: FlashInfer's trtllm-gen decode (NVIDIA's Blackwell-native FMHA cubins) on the same
    paged pool bytes, the [P + page, Hkv, D] buffer viewed as [pages, page, Hkv, D] and the block table cut out of
    req_to_token (baseline/comparators.py TrtllmGenDecode).  Skipped when the cubins / launcher cannot be loaded."""
    from baseline import comparators as C
    hq, hkv, d, lens = 32, 8, 128, [4100, 129, 64, 2, 999, 31, 2048]
    c = G.random_decode_case(hq, hkv, d, lens, seed=23, kv_dtype=dtype, page_size=page_size)
    ours, backend, runner, fb, layer = G.run_decode_case(
        c["q"], c["k_buf"], c["v_buf"], c["req_to_token"], c["req_pool_indices"], c["seq_lens"], c["sm_scale"], 0.0,
        dtype, page_size=page_size)
    pool = runner.token_to_kv_pool
    try:
        tg = C.TrtllmGenDecode(runner.req_to_token_pool.req_to_token, fb.req_pool_indices, fb.seq_lens, page_size)
        theirs = tg(c["q"].cuda(), pool.k_buffer[0], pool.v_buffer[0], c["sm_scale"])
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001  (no cubins, no launcher, unsupported page size: not this repo's code)
        pytest.skip(f"trtllm-gen decode unavailable: {e!r}"[:200])
    tol = (G.BF16_ABS, G.BF16_REL) if dtype == torch.bfloat16 else (G.F16_ABS, G.F16_REL)
    tol = (tol[0], max(tol[1], G.BF16_REL))  # their P is rounded to the 16-bit type as ours is; compare at the bf16 bound
    G.assert_close(ours, theirs.float().cpu(), *tol, what=f"ours vs trtllm-gen decode, page {page_size}, {dtype}")


EXTEND_SHAPES = [
    # hq, hkv, d, prefix_lens, extend_lens
    (32, 8, 128, [300, 0, 129, 1000], [130, 257, 1, 64]),
    (32, 8, 64, [256, 77], [200, 31]),
    (8, 2, 128, [0, 0, 0], [300, 128, 5]),   # no prefix anywhere: the reference's extend_no_prefix fast path
]


@needs_ref
@needs_fi
@pytest.mark.parametrize("hq,hkv,d,pre,ext", EXTEND_SHAPES)
@pytest.mark.parametrize("kv_dtype", [torch.bfloat16, torch.float8_e4m3fn])
@pytest.mark.parametrize("force_path", [-1, 0])
def test_extend_matches_reference_flashinfer_backend(hq, hkv, d, pre, ext, kv_dtype, force_path):
    """ragged (new tokens, causal) + paged (cached prefix) + merge_state, flashinfer_backend.py:378-446."""
    from baseline import comparators as C

    fp8 = kv_dtype != torch.bfloat16
    c = G.random_extend_case(hq, hkv, d, pre, ext, seed=3 + d, kv_dtype=kv_dtype, per_head=False)
    if fp8:
        # the reference's extend path hands no scales to the paged (prefix) wrapper (flashinfer_backend.py:426-433):
        # its fp8 prefix semantics are "scale 1.0", so that is what is compared
        from oracle import attn_oracle as O
        one = torch.ones(hkv)
        gen = torch.Generator().manual_seed(99)
        kf = torch.randn(c["k_buf"].shape, generator=gen).to(torch.bfloat16)
        vf = torch.randn(c["v_buf"].shape, generator=gen).to(torch.bfloat16)
        c["k_buf"], c["v_buf"] = O.quantize_kv_ref(kf, kv_dtype, one), O.quantize_kv_ref(vf, kv_dtype, one)
        c["k_scale"], c["v_scale"] = one, one.clone()
    ours, runner = G.run_extend_case(c, force_path=force_path, save_kv_cache=False)
    st = _ref_state(runner, hq, kv_dtype, 1.0 if fp8 else None, 1.0 if fp8 else None, sm_scale=c["sm_scale"])
    ref = C.RefFlashInfer(st, context_len=runner.model_config.context_len)
    dev = "cuda"
    el = c["extend_seq_lens"].to(torch.int64).to(dev)
    sl = c["seq_lens"].to(torch.int64).to(dev)
    rfb = st.extend_batch(c["req_pool_indices"].to(dev), sl, c["out_cache_loc"].to(dev), el, sl - el, ref.backend)
    ref.init_forward_metadata(rfb)
    t = c["q"].shape[0]
    q = c["q"].reshape(t, -1).cuda()
    k = c["k_new"].reshape(t, -1).cuda()
    v = c["v_new"].reshape(t, -1).cuda()
    theirs = ref.forward(0, q, k, v, rfb, save_kv_cache=False)
    torch.cuda.synchronize()
    theirs = theirs.view(t, hq, d).float().cpu()
    tol = (G.FP8_ABS, G.FP8_REL) if fp8 else (G.BF16_ABS, G.BF16_REL)
    G.assert_close(ours, theirs, *tol, what=f"extend: ours vs reference FlashInfer backend ({kv_dtype})")


@needs_ref
@pytest.mark.parametrize("hq,hkv,d,pre,ext", EXTEND_SHAPES[:2])
@pytest.mark.parametrize("force_path", [-1, 0])
def test_extend_matches_reference_triton_kernel(hq, hkv, d, pre, ext, force_path):
    from baseline import comparators as C

    c = G.random_extend_case(hq, hkv, d, pre, ext, seed=8 + d)
    ours, runner = G.run_extend_case(c, force_path=force_path, save_kv_cache=False)
    pool = runner.token_to_kv_pool
    theirs = C.triton_extend(c["q"].cuda(), c["k_new"].cuda(), c["v_new"].cuda(), pool.k_buffer[0], pool.v_buffer[0],
                             runner.req_to_token_pool.req_to_token, c["req_pool_indices"].cuda(), c["seq_lens"].cuda(),
                             c["extend_seq_lens"].cuda(), c["extend_start_loc"].cuda(), c["sm_scale"])
    torch.cuda.synchronize()
    G.assert_close(ours, theirs.float().cpu(), G.BF16_ABS, G.BF16_REL, what="ours vs in-tree Triton extend")


@needs_ref
def test_kv_store_matches_reference_set_kv_buffer():
    """The reference's own MHATokenToKVPool.set_kv_buffer (pool.py:392-424: div_ in place, cast, two index_put_)
    against b200_kv_store on the same inputs.  bf16 pools: identical bytes.  fp8 pools: identical bytes except where
    ATen's CUDA ``div_`` by a Python scalar -- a multiplication by the fp32 reciprocal, "may lose one bit" says its
    own comment -- rounds differently from the true division ``b200_kv_store`` and the CPU path of the same reference
    code perform; such bytes must be adjacent e4m3 codes and rarer than 1e-4."""
    from harness import synth
    from scratchpad_b200.pool import kv_dtype_code, kv_store

    hkv, d, t = 8, 128, 1024
    gen = torch.Generator().manual_seed(2)
    k = torch.randn((t, hkv, d), generator=gen).to(torch.bfloat16).cuda()
    v = torch.randn((t, hkv, d), generator=gen).to(torch.bfloat16).cuda()
    loc = (torch.randperm(4000, generator=gen)[:t] + 1).cuda()
    for kv_dtype, ks, vs in ((torch.float8_e4m3fn, 0.037, 0.021), (torch.float8_e4m3fn, 0.5, 2.0),
                             (torch.bfloat16, None, None)):
        runner = synth.make_model_runner(32, hkv, d, layer_num=1, pool_tokens=4096, max_reqs=2, context_len=16,
                                         kv_dtype=kv_dtype)
        pool = runner.token_to_kv_pool
        st = _ref_state(runner, 32, kv_dtype, ks, vs)
        # the reference writes through ITS pool object (same tensors); it divides k/v in place, so hand it copies
        st.kv_pool.set_kv_buffer(st.layers[0], loc, k.clone(), v.clone(), ks, vs)
        torch.cuda.synchronize()
        want = [pool.k_buffer[0].clone(), pool.v_buffer[0].clone()]
        pool.k_buffer[0].zero_()
        pool.v_buffer[0].zero_()
        sk = None if ks is None else torch.full((hkv,), ks, device="cuda")
        sv = None if vs is None else torch.full((hkv,), vs, device="cuda")
        kv_store(pool.k_buffer[0], pool.v_buffer[0], loc, k, v, kv_dtype_code(kv_dtype), sk, sv)
        torch.cuda.synchronize()
        for got, ref in zip((pool.k_buffer[0], pool.v_buffer[0]), want):
            if kv_dtype == torch.bfloat16 or ks in (0.5,):  # power-of-two scales: the reciprocal is exact
                assert torch.equal(got, ref), (kv_dtype, ks)
                continue
            diff = got != ref
            n = int(diff.sum())
            assert n <= 1e-4 * got.numel(), f"{n} of {got.numel()} bytes differ"
            if n:
                a = got[diff].view(torch.float8_e4m3fn).float()
                b = ref[diff].view(torch.float8_e4m3fn).float()
                assert ((a - b).abs() <= 0.13 * b.abs().clamp_min(2 ** -9)).all()  # neighbouring e4m3 codes
