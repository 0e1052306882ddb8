"""The backend under the reference's REAL engine-side classes (VERDICT r1 "Missing 5").

``ReqToTokenPool``, ``TokenToKVPoolAllocator``, ``MHATokenToKVPool`` (``memory/pool.py:13-72,189-255,258-424``),
``RadixAttention`` (``nn/attention/radix_attention.py:6-57``), ``ForwardMode`` / ``ForwardBatch`` /
``compute_position_torch`` (``model_executor/forward_info.py:19-66,85-177,452-466``) are loaded unmodified from
``baseline/_ref`` and drive ``CudaB200AttnBackend`` through a small serving schedule the way
``ScheduleBatch.prepare_for_extend`` / ``prepare_for_decode`` (``scheduler/schedule_batch.py:901-1071,1230-1308``)
write ``req_to_token`` and ``out_cache_loc``:

    prefill of request A  ->  prefill of B and C with a radix hit on A's prompt  ->  3 decode steps of {A, B, C}

Every layer-call is checked against the fp32 oracle on the pool bytes, and (when FlashInfer imports) against the
reference's own FlashInferAttnBackend run through the same schedule on its own pools.
"""
from types import SimpleNamespace

import pytest
import torch

import gpu_util as G

pytestmark = pytest.mark.gpu

from baseline import ref_loader  # noqa: E402

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="baseline/_ref not installed")

HQ, HKV, D, LAYERS = 32, 8, 128, 2
CTX, POOL, MAX_REQS = 2048, 8192, 16


class Engine:
    """Minimal scheduler + model runner around the reference's pool / batch classes (one instance per backend)."""

    def __init__(self, kv_dtype, make_backend, k_scale=None, v_scale=None, graph=False):
        self.P = ref_loader.load("memory.pool")
        self.F = ref_loader.load("model_executor.forward_info")
        RA = ref_loader.load("nn.attention.radix_attention")
        dev = "cuda"
        self.kv_dtype = kv_dtype
        self.r2t = self.P.ReqToTokenPool(MAX_REQS, CTX + 4, dev, False)
        self.kv = self.P.MHATokenToKVPool(POOL, 1, kv_dtype, HKV, D, LAYERS, dev, False)
        self.alloc = self.P.TokenToKVPoolAllocator(POOL, kv_dtype, dev, self.kv)
        model_config = SimpleNamespace(num_attention_heads=HQ, head_dim=D, context_len=CTX,
                                       get_num_kv_heads=lambda tp: HKV, is_multimodal=False, is_encoder_decoder=False,
                                       hf_config=SimpleNamespace(architectures=["LlamaForCausalLM"]))
        server_args = SimpleNamespace(kv_cache_dtype="auto" if kv_dtype == torch.bfloat16 else "fp8_e4m3",
                                      enable_dp_attention=False, attention_backend="cuda_b200")
        self.runner = SimpleNamespace(model_config=model_config, tp_size=1, device=dev, dtype=torch.bfloat16,
                                      kv_cache_dtype=kv_dtype, server_args=server_args, req_to_token_pool=self.r2t,
                                      token_to_kv_pool=self.kv, sliding_window_size=None)
        self.backend = make_backend(self.runner)
        self.layers = [RA.RadixAttention(HQ, D, D ** -0.5, HKV, i) for i in range(LAYERS)]
        for l in self.layers:
            l.k_scale_float, l.v_scale_float = k_scale, v_scale
        self.rows = {}      # request name -> req_pool_idx
        self.lens = {}      # request name -> tokens in the pool

    # -- ScheduleBatch.prepare_for_extend (schedule_batch.py:901-1071), page size 1, no chunking ------------------
    def extend(self, reqs):
        """reqs: list of (name, prefix_indices int64 tensor from a radix match, extend_len)."""
        F, dev = self.F, "cuda"
        bs = len(reqs)
        req_pool_indices = self.r2t.alloc(bs)
        prefix_lens = [int(r[1].numel()) for r in reqs]
        extend_lens = [r[2] for r in reqs]
        seq_lens = [p + e for p, e in zip(prefix_lens, extend_lens)]
        for i, (name, prefix, _) in enumerate(reqs):
            self.rows[name], self.lens[name] = req_pool_indices[i], seq_lens[i]
            if prefix_lens[i] > 0:
                self.r2t.write((req_pool_indices[i], slice(0, prefix_lens[i])), prefix)
        rpi = torch.tensor(req_pool_indices, dtype=torch.int64).to(dev)
        sl = torch.tensor(seq_lens, dtype=torch.int64).to(dev)
        out_cache_loc = self.alloc.alloc(sum(extend_lens))
        pt = 0
        for i in range(bs):  # the torch branch of the req_to_token write (:1052-1058)
            self.r2t.write((req_pool_indices[i], slice(prefix_lens[i], seq_lens[i])), out_cache_loc[pt:pt + extend_lens[i]])
            pt += extend_lens[i]
        # ForwardBatch.init_new (forward_info.py:257-279)
        esl = torch.tensor(extend_lens, dtype=torch.int32).to(dev)
        epl = torch.tensor(prefix_lens, dtype=torch.int32).to(dev)
        positions, start = F.compute_position_torch(epl, esl)
        return F.ForwardBatch(forward_mode=F.ForwardMode.EXTEND, batch_size=bs, input_ids=None, req_pool_indices=rpi,
                              seq_lens=sl, out_cache_loc=out_cache_loc, seq_lens_sum=sum(seq_lens), positions=positions,
                              extend_num_tokens=sum(extend_lens), extend_seq_lens=esl, extend_prefix_lens=epl,
                              extend_start_loc=start, extend_prefix_lens_cpu=prefix_lens, extend_seq_lens_cpu=extend_lens,
                              req_to_token_pool=self.r2t, token_to_kv_pool=self.kv, attn_backend=self.backend)

    # -- ScheduleBatch.prepare_for_decode (schedule_batch.py:1230-1308) ------------------------------------------
    def decode(self, names):
        F, dev = self.F, "cuda"
        bs = len(names)
        rpi = torch.tensor([self.rows[n] for n in names], dtype=torch.int64).to(dev)
        seq_lens = torch.tensor([self.lens[n] for n in names], dtype=torch.int64).to(dev)
        locs = seq_lens.clone()
        seq_lens = seq_lens + 1  # overlap mode: a fresh tensor (:1287-1289)
        out_cache_loc = self.alloc.alloc(bs)
        self.r2t.write((rpi, locs), out_cache_loc.to(torch.int32))
        for n in names:
            self.lens[n] += 1
        return F.ForwardBatch(forward_mode=F.ForwardMode.DECODE, batch_size=bs, input_ids=None, req_pool_indices=rpi,
                              seq_lens=seq_lens, out_cache_loc=out_cache_loc, seq_lens_sum=int(seq_lens.sum()),
                              positions=(seq_lens - 1), req_to_token_pool=self.r2t, token_to_kv_pool=self.kv,
                              attn_backend=self.backend)

    def forward(self, fb, qkv):
        """ModelRunner.forward_decode / forward_extend (model_runner.py:506-529): metadata once, then every layer."""
        self.backend.init_forward_metadata(fb)
        outs = []
        for i, layer in enumerate(self.layers):
            q, k, v = qkv[i].split([HQ * D, HKV * D, HKV * D], dim=-1)  # column slices of the qkv GEMM output
            outs.append(layer(q, k.clone(), v.clone(), fb))  # the reference's fp8 store divides k/v in place
        torch.cuda.synchronize()
        return outs

    def radix_match(self, name, n):
        """What RadixCache.match_prefix returns for a prompt sharing its first n tokens with `name`."""
        return self.r2t.req_to_token[self.rows[name], :n].to(torch.int64).clone()


def _oracle(engine, fb, qkv, layer_id):
    from oracle import attn_oracle as O

    kv, F = engine.kv, engine.F
    q, k, v = qkv[layer_id].cpu().split([HQ * D, HKV * D, HKV * D], dim=-1)
    kb = kv.k_buffer[layer_id].cpu()
    vb = kv.v_buffer[layer_id].cpu()
    r2t = engine.r2t.req_to_token.cpu()
    fp8 = engine.kv_dtype != torch.bfloat16
    ks = vs = None
    if fp8:
        ks = torch.full((HKV,), engine.layers[layer_id].k_scale_float)
        vs = torch.full((HKV,), engine.layers[layer_id].v_scale_float)
    t = q.shape[0]
    if fb.forward_mode.is_decode():
        return O.decode_ref(q.reshape(t, HQ, D), kb, vb, r2t, fb.req_pool_indices.cpu(), fb.seq_lens.cpu(), D ** -0.5,
                            0.0, engine.kv_dtype, ks, vs)
    return O.extend_ref(q.reshape(t, HQ, D), k.reshape(t, HKV, D), v.reshape(t, HKV, D), kb, vb, r2t,
                        fb.req_pool_indices.cpu(), fb.seq_lens.cpu(), fb.extend_seq_lens.cpu(), fb.extend_start_loc.cpu(),
                        D ** -0.5, 0.0, engine.kv_dtype, ks, vs)


def _schedule(engine, seed):
    """Yields (tag, forward_batch, qkv per layer, outputs per layer) for the whole schedule."""
    gen = torch.Generator(device="cuda").manual_seed(seed)

    def qkv(t):
        return [torch.randn((t, (HQ + 2 * HKV) * D), generator=gen, device="cuda", dtype=torch.float32).to(torch.bfloat16)
                for _ in range(LAYERS)]

    empty = torch.empty(0, dtype=torch.int64, device="cuda")
    fb = engine.extend([("A", empty, 700)])
    x = qkv(700)
    yield "prefill A", fb, x, engine.forward(fb, x)
    shared = 640  # B and C hit the radix cache on A's first 640 tokens: same slot ids in their rows
    fb = engine.extend([("B", engine.radix_match("A", shared), 150), ("C", engine.radix_match("A", shared), 33)])
    x = qkv(183)
    yield "prefill B,C on A's prefix", fb, x, engine.forward(fb, x)
    for step in range(3):
        fb = engine.decode(["A", "B", "C"])
        x = qkv(3)
        yield f"decode {step}", fb, x, engine.forward(fb, x)


@needs_ref
@pytest.mark.parametrize("kv_dtype,ks,vs", [(torch.bfloat16, None, None), (torch.float8_e4m3fn, 1.0, 1.0)])
def test_backend_under_the_reference_engine_classes(kv_dtype, ks, vs):
    from scratchpad_b200 import CudaB200AttnBackend

    ours = Engine(kv_dtype, lambda r: CudaB200AttnBackend(r), ks, vs)
    theirs = None
    try:
        from baseline import comparators  # noqa: F401
        mod = ref_loader.flashinfer_backend_module()
        mod.global_workspace_buffer = None
        theirs = Engine(kv_dtype, lambda r: mod.FlashInferAttnBackend(r), ks, vs)
    except ImportError:
        pass
    tol = (G.FP8_ABS, G.FP8_REL) if kv_dtype != torch.bfloat16 else (G.BF16_ABS, G.BF16_REL)
    ref_iter = _schedule(theirs, 7) if theirs is not None else None
    for tag, fb, x, outs in _schedule(ours, 7):
        assert fb.req_pool_indices.dtype == torch.int64 and fb.out_cache_loc.dtype == torch.int64
        for i, o in enumerate(outs):
            assert o.shape == (x[i].shape[0], HQ * D) and o.dtype == torch.bfloat16
            want = _oracle(ours, fb, x, i)
            G.assert_close(o.view(-1, HQ, D), want, *tol, what=f"{tag}, layer {i} vs oracle")
        if ref_iter is not None:
            rtag, rfb, rx, routs = next(ref_iter)
            assert rtag == tag and torch.equal(rfb.out_cache_loc, fb.out_cache_loc)  # same schedule, same slots
            for i in range(LAYERS):
                assert torch.equal(rx[i], x[i])
                G.assert_close(outs[i], routs[i], *tol, what=f"{tag}, layer {i} vs the reference FlashInfer backend")
    # the two engines hold the same KV pool bytes at the end (bf16; fp8 with unit scales)
    if theirs is not None:
        for i in range(LAYERS):
            assert torch.equal(ours.kv.k_buffer[i], theirs.kv.k_buffer[i])
            assert torch.equal(ours.kv.v_buffer[i], theirs.kv.v_buffer[i])
    # radix sharing really is equal slot ids
    a, b = ours.rows["A"], ours.rows["B"]
    assert torch.equal(ours.r2t.req_to_token[a, :640], ours.r2t.req_to_token[b, :640])
