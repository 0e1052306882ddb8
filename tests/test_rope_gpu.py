"""GPU parity of the fused rotary-embedding + KV-write launch (b200_rope_kv_store) through the Python surface ->
ctypes -> C ABI: bit-exact against the reference's own RotaryEmbedding.forward_native outputs
(tests/golden/rope_*.npz) and against oracle/rope_oracle.py + the pool-store oracle at larger shapes."""
import pytest
import torch

from oracle import attn_oracle as O
from oracle import rope_oracle as R
from golden_util import golden_names, load_golden
from scratchpad_b200 import CudaB200AttnBackend
from harness import ForwardBatch, ForwardMode, synth
from scratchpad_b200.rotary_embedding import RotaryEmbedding, get_rope

import gpu_util as G

pytestmark = pytest.mark.gpu

LLAMA3 = {"rope_type": "llama3", "factor": 32.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
          "original_max_position_embeddings": 8192}


def bits(x):
    return x.contiguous().view(torch.int16)


@pytest.mark.parametrize("name", golden_names("rope"))  # rope_*: bf16 model dtype, ropeh_*: fp16
@pytest.mark.parametrize("cache", ["model_dtype", "f32"])
def test_rotation_matches_reference_golden_bit_exact(name, cache):
    g = load_golden(name)
    dtype = torch.float16 if g.get("dtype", "bf16") == "f16" else torch.bfloat16
    as_model = lambda a: a.view(torch.int16).view(dtype)  # the fixtures hold 16-bit patterns
    emb = RotaryEmbedding(g["head_size"], g["rotary_dim"], g["max_position"], g["base"], bool(g["is_neox"]), dtype)
    ref_cache = as_model(g["cos_sin_cache"])  # the reference module's own cache
    emb.cos_sin_cache = (ref_cache if cache == "model_dtype" else ref_cache.float()).cuda()
    q, k = as_model(g["q"]).cuda(), as_model(g["k"]).cuda()
    q2, k2 = emb(g["positions"].cuda(), q, k)
    assert q2.data_ptr() == q.data_ptr() and k2.data_ptr() == k.data_ptr()  # in place, like forward_cuda
    assert torch.equal(bits(q.cpu()), bits(as_model(g["q_out"])))
    assert torch.equal(bits(k.cpu()), bits(as_model(g["k_out"])))


@pytest.mark.parametrize("kv_dtype", [torch.bfloat16, torch.float8_e4m3fn])
@pytest.mark.parametrize("neox,d,rot", [(True, 128, 128), (True, 64, 64), (False, 64, 64), (True, 128, 64)])
@pytest.mark.parametrize("pos_dtype", [torch.int64, torch.int32])
def test_fused_store_matches_oracle_bit_exact(kv_dtype, neox, d, rot, pos_dtype):
    """q rotated in place; pool rows = quant(rope(k)), quant(v) at loc; k untouched unless write_back_k;
    q/k/v as column slices of a fused qkv buffer (llama.py:149-152)."""
    hq, hkv, t, pool_tokens = 8, 2, 67, 300
    gen = torch.Generator().manual_seed(11)
    emb = get_rope(d, rot, 4096, 500000.0, neox, LLAMA3, dtype=torch.bfloat16)
    qkv = torch.randn((t, (hq + 2 * hkv) * d), generator=gen).to(torch.bfloat16)
    pos = torch.randint(0, 4096, (t,), generator=gen).to(pos_dtype)
    loc = (torch.randperm(pool_tokens, generator=gen)[:t] + 1).to(torch.int64)
    q, k, v = qkv.split([hq * d, hkv * d, hkv * d], dim=-1)
    q_ref = R.rope_ref(pos, q.reshape(t, hq, d), emb.cos_sin_cache.cpu(), rot, neox)
    k_ref = R.rope_ref(pos, k.reshape(t, hkv, d), emb.cos_sin_cache.cpu(), rot, neox)
    ks = vs = None
    if kv_dtype != torch.bfloat16:
        ks = k_ref.float().abs().amax(dim=(0, 2)) / 448.0
        vs = v.reshape(t, hkv, d).float().abs().amax(dim=(0, 2)) / 448.0
    store = torch.uint8 if kv_dtype != torch.bfloat16 else torch.bfloat16
    kb_ref = torch.zeros((pool_tokens + 1, hkv, d), dtype=store)
    vb_ref = torch.zeros_like(kb_ref)
    O.set_kv_buffer_ref(kb_ref, vb_ref, loc, k_ref.clone(), v.reshape(t, hkv, d).clone(), kv_dtype, ks, vs)

    for write_back in (False, True):
        runner = synth.make_model_runner(hq, hkv, d, tp_size=1, layer_num=2, pool_tokens=pool_tokens, max_reqs=4,
                                         context_len=64, kv_dtype=kv_dtype)
        pool = runner.token_to_kv_pool
        if ks is not None:
            pool.set_kv_scales(1, ks, vs)
        dev = qkv.cuda()
        qd, kd, vd = dev.split([hq * d, hkv * d, hkv * d], dim=-1)
        emb.rope_kv_store(pos.cuda(), qd, kd, vd, pool, 1, loc.cuda(), write_back)
        torch.cuda.synchronize()
        assert torch.equal(bits(qd.cpu()), bits(q_ref.reshape(t, -1)))
        assert torch.equal(bits(kd.cpu()), bits(k_ref.reshape(t, -1)) if write_back else bits(k))
        assert torch.equal(vd.cpu(), v)
        assert torch.equal(pool.k_buffer[1].cpu(), kb_ref) and torch.equal(pool.v_buffer[1].cpu(), vb_ref)
        assert pool.k_buffer[0].abs().sum().item() == 0  # other layers untouched


def _decode_fb(runner, backend, c, out_loc):
    b = c["q"].shape[0]
    return ForwardBatch(forward_mode=ForwardMode.DECODE, batch_size=b, req_pool_indices=c["req_pool_indices"].cuda(),
                        seq_lens=c["seq_lens"].cuda(), out_cache_loc=out_loc.cuda(), seq_lens_sum=int(c["seq_lens"].sum()),
                        req_to_token_pool=runner.req_to_token_pool, token_to_kv_pool=runner.token_to_kv_pool,
                        attn_backend=backend)


@pytest.mark.parametrize("kv_dtype", [torch.bfloat16, torch.float8_e4m3fn])
def test_decode_with_fused_rope_equals_rotate_then_decode(kv_dtype):
    """RadixAttention.forward(q, k, v, fb, rope=(emb, positions)) == the reference's sequence
    rotary_emb(positions, q, k); attn(q, k, v) -- same kernels downstream, so outputs are bit-identical; both
    within tolerance of the fp32 oracle on the rotated inputs."""
    hq, hkv, d = 32, 8, 128
    seq_lens = [257, 1, 64, 300, 129, 33]
    c = G.random_decode_case(hq, hkv, d, seq_lens, seed=21, kv_dtype=kv_dtype)
    b = len(seq_lens)
    gen = torch.Generator().manual_seed(22)
    k_new = torch.randn((b, hkv * d), generator=gen).to(torch.bfloat16)
    v_new = torch.randn((b, hkv * d), generator=gen).to(torch.bfloat16)
    pos = c["seq_lens"] - 1
    out_loc = torch.stack([c["req_to_token"][int(c["req_pool_indices"][i]), seq_lens[i] - 1] for i in range(b)]).long()
    emb = get_rope(d, d, 8192, 500000.0, True, LLAMA3, dtype=torch.bfloat16)
    outs = []
    for fused in (True, False):
        runner = G.make_runner_from_case(c["k_buf"], c["v_buf"], c["req_to_token"], kv_dtype, c["k_scale"],
                                         c["v_scale"], tp_hq=hq)
        backend = CudaB200AttnBackend(runner)
        layer = synth.make_layer(runner, 0, 0.0)
        fb = _decode_fb(runner, backend, c, out_loc)
        backend.init_forward_metadata(fb)
        q, k, v = c["q"].reshape(b, -1).cuda(), k_new.cuda(), v_new.cuda()
        if fused:
            o = layer(q, k, v, fb, rope=(emb, pos.cuda()))
        else:
            q, k = emb(pos.cuda(), q, k)
            o = layer(q, k, v, fb)
        torch.cuda.synchronize()
        outs.append((o.cpu(), runner.token_to_kv_pool.k_buffer[0].cpu(), runner.token_to_kv_pool.v_buffer[0].cpu()))
    assert torch.equal(bits(outs[0][0]), bits(outs[1][0]))
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    # oracle: rotate, store, attend in fp32
    q_rot = R.rope_ref(pos, c["q"], emb.cos_sin_cache.cpu(), d, True)
    k_rot = R.rope_ref(pos, k_new.view(b, hkv, d), emb.cos_sin_cache.cpu(), d, True)
    kb, vb = c["k_buf"].clone(), c["v_buf"].clone()
    O.set_kv_buffer_ref(kb, vb, out_loc, k_rot, v_new.view(b, hkv, d).clone(), kv_dtype, c["k_scale"], c["v_scale"])
    ref = O.decode_ref(q_rot.float(), kb, vb, c["req_to_token"], c["req_pool_indices"], c["seq_lens"], c["sm_scale"],
                       kv_dtype=kv_dtype, k_scale=c["k_scale"], v_scale=c["v_scale"])
    tol = (G.BF16_ABS, G.BF16_REL) if kv_dtype == torch.bfloat16 else (G.FP8_ABS, G.FP8_REL)
    G.assert_close(outs[0][0].view(b, hq, d), ref, *tol, what="decode with fused rope")


def test_extend_with_fused_rope_equals_rotate_then_extend():
    hq, hkv, d = 8, 2, 128
    c = G.random_extend_case(hq, hkv, d, [130, 0, 77], [64, 150, 3], seed=31)
    t = c["q"].shape[0]
    pos = torch.cat([torch.arange(p, p + e) for p, e in zip([130, 0, 77], [64, 150, 3])])
    emb = get_rope(d, d, 8192, 500000.0, True, LLAMA3, dtype=torch.bfloat16)
    c_rot = dict(c)
    c_rot["q"] = R.rope_ref(pos, c["q"], emb.cos_sin_cache.cpu(), d, True)
    c_rot["k_new"] = R.rope_ref(pos, c["k_new"], emb.cos_sin_cache.cpu(), d, True)
    want, runner_a = G.run_extend_case(c_rot)

    runner = G.make_runner_from_case(c["k_buf"], c["v_buf"], c["req_to_token"], torch.bfloat16, tp_hq=hq)
    backend = CudaB200AttnBackend(runner)
    layer = synth.make_layer(runner, 0, 0.0)
    layer.scaling = c["sm_scale"]
    el = c["extend_seq_lens"]
    pre = c["extend_prefix_lens"]
    fb = ForwardBatch(forward_mode=ForwardMode.EXTEND, batch_size=len(el), req_pool_indices=c["req_pool_indices"].cuda(),
                      seq_lens=c["seq_lens"].cuda(), out_cache_loc=c["out_cache_loc"].cuda(),
                      seq_lens_sum=int(c["seq_lens"].sum()), extend_num_tokens=t, extend_seq_lens=el.cuda(),
                      extend_prefix_lens=pre.cuda(), extend_start_loc=c["extend_start_loc"].to(torch.int32).cuda(),
                      extend_prefix_lens_cpu=pre.tolist(), extend_seq_lens_cpu=el.tolist(),
                      req_to_token_pool=runner.req_to_token_pool, token_to_kv_pool=runner.token_to_kv_pool,
                      attn_backend=backend)
    qkv = torch.cat([c["q"].reshape(t, -1), c["k_new"].reshape(t, -1), c["v_new"].reshape(t, -1)], dim=1).cuda()
    qd, kd, vd = qkv.split([hq * d, hkv * d, hkv * d], dim=-1)
    backend.init_forward_metadata(fb)
    o = layer(qd, kd, vd, fb, rope=(emb, pos.cuda()))
    torch.cuda.synchronize()
    assert torch.equal(bits(o.cpu().view(t, hq, d)), bits(want))
    assert torch.equal(bits(kd.cpu()), bits(c_rot["k_new"].reshape(t, -1)))  # written back for the extend kernel
    assert torch.equal(runner.token_to_kv_pool.k_buffer[0].cpu(), runner_a.token_to_kv_pool.k_buffer[0].cpu())
