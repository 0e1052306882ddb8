"""Helpers for the -m gpu parity tests: run a CPU-defined case through the
cuda_b200 backend (Python surface -> ctypes -> C ABI -> sm_100a kernels)."""
import torch

from scratchpad_b200 import CudaB200AttnBackend
from harness import ForwardBatch, ForwardMode, compute_extend_start_loc, synth

BF16_ABS, BF16_REL = 2e-2, 1e-2   # BASELINE.json north_star tolerance, bf16 KV
F16_ABS, F16_REL = 1e-2, 5e-3     # fp16 model dtype over an fp16 pool: 3 more mantissa bits than bf16, tighter bound
FP8_ABS, FP8_REL = 5e-2, 5e-2     # fp8-e4m3 KV


def rel_err(got, ref):
    return ((got.float() - ref.float()).norm() / ref.float().norm().clamp_min(1e-12)).item()


def assert_close(got, ref, abs_tol, rel_tol, what=""):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    mx = (got - ref).abs().max().item()
    rl = rel_err(got, ref)
    assert mx <= abs_tol and rl <= rel_tol, f"{what}: max-abs {mx:.4g} (tol {abs_tol}), rel {rl:.4g} (tol {rel_tol})"
    return mx, rl


FUSED_KV = False  # tests flip this to run every helper over the fused [P+1, 2, Hkv, D] pool layout


def make_runner_from_case(k_buf, v_buf, r2t, kv_dtype, k_scale=None, v_scale=None, tp_hq=None, page_size=1, dtype=None):
    """Pools on the GPU holding exactly the CPU case's bytes.  ``dtype``: the model dtype (see make_model_runner)."""
    p1, hkv, d = k_buf.shape
    runner = synth.make_model_runner(tp_hq, hkv, d, tp_size=1, layer_num=1, pool_tokens=p1 - page_size,
                                     max_reqs=r2t.shape[0] - 1, context_len=r2t.shape[1] - 4, kv_dtype=kv_dtype,
                                     fused_kv=FUSED_KV, page_size=page_size, dtype=dtype)
    pool = runner.token_to_kv_pool
    pool.k_buffer[0].copy_(k_buf.view(pool.store_dtype))
    pool.v_buffer[0].copy_(v_buf.view(pool.store_dtype))
    runner.req_to_token_pool.req_to_token.copy_(r2t)
    if k_scale is not None:
        pool.set_kv_scales(0, k_scale, v_scale)
    return runner


def run_decode_case(q, k_buf, v_buf, r2t, rpi, sl, sm_scale, logit_cap=0.0, kv_dtype=torch.bfloat16,
                    k_scale=None, v_scale=None, idx_dtype=torch.int64, q_slice_of_qkv=False, force_path=-1,
                    cascade=None, page_size=1, tma_contiguous_runs=None):
    b, hq, d = q.shape
    runner = make_runner_from_case(k_buf, v_buf, r2t, kv_dtype, k_scale, v_scale, tp_hq=hq, page_size=page_size,
                                   dtype=q.dtype)
    backend = CudaB200AttnBackend(runner, force_path=force_path, cascade=cascade, tma_contiguous_runs=tma_contiguous_runs)
    layer = synth.make_layer(runner, 0, logit_cap)
    layer.scaling = sm_scale
    fb = ForwardBatch(forward_mode=ForwardMode.DECODE, batch_size=b,
                      req_pool_indices=rpi.to(idx_dtype).cuda(), seq_lens=sl.to(idx_dtype).cuda(),
                      out_cache_loc=torch.zeros(b, dtype=torch.int64, device="cuda"), seq_lens_sum=int(sl.sum()),
                      req_to_token_pool=runner.req_to_token_pool, token_to_kv_pool=runner.token_to_kv_pool,
                      attn_backend=backend)
    hkv = k_buf.shape[1]
    if q_slice_of_qkv:  # q as a column slice of the fused qkv GEMM output (llama.py:149-152)
        qkv = torch.zeros((b, (hq + 2 * hkv) * d), dtype=q.dtype, device="cuda")
        qkv[:, : hq * d] = q.reshape(b, hq * d).cuda()
        qd = qkv[:, : hq * d]
    else:
        qd = q.reshape(b, hq * d).cuda()
    backend.init_forward_metadata(fb)
    o = layer(qd, None, None, fb, save_kv_cache=False)
    torch.cuda.synchronize()
    return o.view(b, hq, -1).cpu(), backend, runner, fb, layer


def random_decode_case(hq, hkv, d, seq_lens, seed, kv_dtype=torch.bfloat16, shared_prefix=0, per_head=True,
                       extra_rows=3, mode="random", page_size=1, act_dtype=None):
    """CPU tensors for one decode case (values in the model dtype ``act_dtype`` -- default: the pool's when 16-bit, else
    bf16; fp8 pools quantised with amax/448 scales).
    page_size > 1: page-granular tables (synth.build_paged_table) over a [pool + page_size, Hkv, D] buffer."""
    if act_dtype is None:
        act_dtype = kv_dtype if kv_dtype in (torch.bfloat16, torch.float16) else torch.bfloat16
    gen = torch.Generator().manual_seed(seed)
    b = len(seq_lens)
    r2t = torch.zeros((b + extra_rows, max(max(seq_lens), 1) + 4), dtype=torch.int32)
    if page_size > 1:
        ps = page_size
        pool = (shared_prefix // ps + sum((n + ps - 1) // ps for n in seq_lens) + 3) * ps
        rpi = synth.build_paged_table(r2t, seq_lens, pool, ps, gen, shared_prefix)
    else:
        need = shared_prefix + sum(n - shared_prefix if shared_prefix and n >= shared_prefix else n for n in seq_lens)
        pool = need + 17
        rpi = synth.build_page_table(r2t, seq_lens, pool, gen, mode, shared_prefix)
    kf = torch.randn((pool + page_size, hkv, d), generator=gen).to(act_dtype)
    vf = torch.randn((pool + page_size, hkv, d), generator=gen).to(act_dtype)
    q = torch.randn((b, hq, d), generator=gen).to(act_dtype)
    ks = vs = None
    if kv_dtype != act_dtype:
        from oracle import attn_oracle as O
        if per_head:
            ks = kf.float().abs().amax(dim=(0, 2)) / 448.0
            vs = vf.float().abs().amax(dim=(0, 2)) / 448.0
        else:
            ks = (kf.float().abs().amax() / 448.0).expand(hkv).clone()
            vs = (vf.float().abs().amax() / 448.0).expand(hkv).clone()
        kf = O.quantize_kv_ref(kf, kv_dtype, ks)
        vf = O.quantize_kv_ref(vf, kv_dtype, vs)
    sl = torch.tensor(seq_lens, dtype=torch.int64)
    return dict(q=q, k_buf=kf, v_buf=vf, req_to_token=r2t, req_pool_indices=rpi, seq_lens=sl,
                sm_scale=d ** -0.5, k_scale=ks, v_scale=vs, kv_dtype=kv_dtype)


def random_extend_case(hq, hkv, d, prefix_lens, extend_lens, seed, kv_dtype=torch.bfloat16, per_head=True,
                       extra_rows=2, mode="random", page_size=1, act_dtype=None):
    """CPU tensors for one extend case: pool holding the prefixes (quantised for fp8), new q/k/v in the model dtype."""
    if act_dtype is None:
        act_dtype = kv_dtype if kv_dtype in (torch.bfloat16, torch.float16) else torch.bfloat16
    gen = torch.Generator().manual_seed(seed)
    seq_lens = [p + e for p, e in zip(prefix_lens, extend_lens)]
    b = len(seq_lens)
    r2t = torch.zeros((b + extra_rows, max(max(seq_lens), 1) + 4), dtype=torch.int32)
    if page_size > 1:
        pool = (sum((n + page_size - 1) // page_size for n in seq_lens) + 3) * page_size
        rpi = synth.build_paged_table(r2t, seq_lens, pool, page_size, gen)
    else:
        pool = sum(seq_lens) + 19
        rpi = synth.build_page_table(r2t, seq_lens, pool, gen, mode)
    kf = torch.randn((pool + page_size, hkv, d), generator=gen).to(act_dtype)
    vf = torch.randn((pool + page_size, hkv, d), generator=gen).to(act_dtype)
    t = sum(extend_lens)
    q = torch.randn((t, hq, d), generator=gen).to(act_dtype)
    k_new = torch.randn((t, hkv, d), generator=gen).to(act_dtype)
    v_new = torch.randn((t, hkv, d), generator=gen).to(act_dtype)
    ks = vs = None
    if kv_dtype != act_dtype:
        from oracle import attn_oracle as O
        if per_head:
            ks = kf.float().abs().amax(dim=(0, 2)) / 448.0
            vs = vf.float().abs().amax(dim=(0, 2)) / 448.0
        else:
            ks = (kf.float().abs().amax() / 448.0).expand(hkv).clone()
            vs = (vf.float().abs().amax() / 448.0).expand(hkv).clone()
        kf = O.quantize_kv_ref(kf, kv_dtype, ks)
        vf = O.quantize_kv_ref(vf, kv_dtype, vs)
    el = torch.tensor(extend_lens, dtype=torch.int32)
    out_loc = torch.cat([r2t[rpi[i], p:p + e] for i, (p, e) in enumerate(zip(prefix_lens, extend_lens))]).long()
    return dict(q=q, k_new=k_new, v_new=v_new, k_buf=kf, v_buf=vf, req_to_token=r2t, req_pool_indices=rpi,
                seq_lens=torch.tensor(seq_lens, dtype=torch.int64), extend_seq_lens=el,
                extend_prefix_lens=torch.tensor(prefix_lens, dtype=torch.int32),
                extend_start_loc=compute_extend_start_loc(el), out_cache_loc=out_loc, sm_scale=d ** -0.5,
                k_scale=ks, v_scale=vs, kv_dtype=kv_dtype, page_size=page_size)


def run_extend_case(c, logit_cap=0.0, idx_dtype=torch.int64, save_kv_cache=True, qkv_slices=True,
                    forward_mode=ForwardMode.EXTEND, force_path=-1):
    """Runs one extend case through CudaB200AttnBackend.forward_extend.  Returns (o [T,Hq,Dv] cpu, runner)."""
    t, hq, d = c["q"].shape
    hkv = c["k_buf"].shape[1]
    kv_dtype = c.get("kv_dtype", torch.bfloat16)
    runner = make_runner_from_case(c["k_buf"], c["v_buf"], c["req_to_token"], kv_dtype, c.get("k_scale"),
                                   c.get("v_scale"), tp_hq=hq, page_size=c.get("page_size", 1), dtype=c["q"].dtype)
    backend = CudaB200AttnBackend(runner, force_path=force_path)
    layer = synth.make_layer(runner, 0, logit_cap)
    layer.scaling = c["sm_scale"]
    el = c["extend_seq_lens"].to(torch.int32)
    pre = (c["seq_lens"].to(torch.int64) - el.to(torch.int64)).to(torch.int32)
    if "out_cache_loc" in c:
        out_loc = c["out_cache_loc"]
    else:
        out_loc = torch.cat([c["req_to_token"][int(c["req_pool_indices"][i]), int(pre[i]):int(c["seq_lens"][i])]
                             for i in range(len(el))]).long()
    fb = ForwardBatch(forward_mode=forward_mode, batch_size=len(el),
                      req_pool_indices=c["req_pool_indices"].to(idx_dtype).cuda(),
                      seq_lens=c["seq_lens"].to(idx_dtype).cuda(), out_cache_loc=out_loc.cuda(),
                      seq_lens_sum=int(c["seq_lens"].sum()), extend_num_tokens=t, extend_seq_lens=el.cuda(),
                      extend_prefix_lens=pre.cuda(), extend_start_loc=c["extend_start_loc"].to(torch.int32).cuda(),
                      extend_prefix_lens_cpu=pre.tolist(), extend_seq_lens_cpu=el.tolist(),
                      req_to_token_pool=runner.req_to_token_pool, token_to_kv_pool=runner.token_to_kv_pool,
                      attn_backend=backend)
    if qkv_slices:  # q/k/v as column slices of the fused qkv GEMM output (llama.py:149-152)
        qkv = torch.zeros((t, (hq + 2 * hkv) * d), dtype=c["q"].dtype, device="cuda")
        qkv[:, : hq * d] = c["q"].reshape(t, -1).cuda()
        qkv[:, hq * d: (hq + hkv) * d] = c["k_new"].reshape(t, -1).cuda()
        qkv[:, (hq + hkv) * d:] = c["v_new"].reshape(t, -1).cuda()
        qd, kd, vd = qkv[:, : hq * d], qkv[:, hq * d: (hq + hkv) * d], qkv[:, (hq + hkv) * d:]
    else:
        qd, kd, vd = c["q"].reshape(t, -1).cuda(), c["k_new"].reshape(t, -1).cuda(), c["v_new"].reshape(t, -1).cuda()
    backend.init_forward_metadata(fb)
    o = layer(qd, kd, vd, fb, save_kv_cache=save_kv_cache)
    torch.cuda.synchronize()
    return o.view(t, hq, -1).cpu(), runner
