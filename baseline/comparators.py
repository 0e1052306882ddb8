"""The reference's own GPU attention path on the tensors the cuda_b200 backend is given (SURVEY.md 8d "GPU
comparators (same box, same inputs)").

* ``RefFlashInfer`` instantiates the reference's **unmodified** ``FlashInferAttnBackend``
  (``scratchpad/nn/attention/flashinfer_backend.py``, loaded from ``baseline/_ref``) over the reference's own
  ``ReqToTokenPool`` / ``MHATokenToKVPool`` / ``RadixAttention`` / ``ForwardBatch`` classes, whose tensors alias the
  ones the caller built.  Metadata (``kv_indices`` flatten + ``begin_forward`` plan,
  ``flashinfer_backend.py:623-670``) and the per-layer call (``forward_decode`` ``:448-484``, ``forward_extend``
  ``:378-446``: ragged + paged + ``merge_state``) run exactly as the engine would run them.
* ``triton_decode`` / ``triton_extend`` call the reference's in-tree Triton kernels
  (``triton_attn/decode_attention.py:547-608``, ``extend_attention.py:229-327``), JIT-compiled for the device.

Environment shims, none of them edits reference code: (1) ``RadixAttention`` in this fork lacks the
``k_scale_float`` / ``v_scale_float`` attributes its own FlashInfer backend reads for fp8 pools
(``flashinfer_backend.py:387-388,457-458``) -- they are set on the layer objects here, as
upstream SGLang's class has them; (2) absent engine modules are stubbed by ``baseline/ref_loader.py``.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch

from . import ref_loader


def _ref_runner(hq_total: int, hkv_total: int, head_dim: int, tp_size: int, context_len: int, kv_dtype: torch.dtype,
                device, r2t_pool, kv_pool, max_reqs: int, model_dtype: torch.dtype = torch.bfloat16):
    """What FlashInferAttnBackend.__init__ reads from ModelRunner (flashinfer_backend.py:62-74,107)."""
    model_config = SimpleNamespace(
        num_attention_heads=hq_total, head_dim=head_dim, context_len=context_len,
        get_num_kv_heads=lambda tp: max(1, hkv_total // tp), is_multimodal=False, is_encoder_decoder=False,
        hf_config=SimpleNamespace(architectures=["LlamaForCausalLM"]),
    )
    server_args = SimpleNamespace(
        kv_cache_dtype="auto" if kv_dtype == model_dtype else "fp8_e4m3", enable_dp_attention=False,
        triton_attention_num_kv_splits=8, attention_backend="flashinfer",
    )
    return SimpleNamespace(model_config=model_config, tp_size=tp_size, device=device, dtype=model_dtype,
                           kv_cache_dtype=kv_dtype, server_args=server_args, req_to_token_pool=r2t_pool,
                           token_to_kv_pool=kv_pool, sliding_window_size=None)


class RefState:
    """Reference pool / layer / batch objects aliasing the caller's tensors."""

    def __init__(self, *, req_to_token: torch.Tensor, k_buffers, v_buffers, kv_dtype: torch.dtype, hq: int, hkv: int,
                 head_dim: int, k_scale: Optional[float] = None, v_scale: Optional[float] = None,
                 logit_cap: float = 0.0, sm_scale: Optional[float] = None, model_dtype: torch.dtype = torch.bfloat16):
        pool_mod = ref_loader.load("memory.pool")
        self.model_dtype = model_dtype  # ModelRunner.dtype: bf16, or fp16 (model_runner.py:366-367 makes it the pool's too)
        self.fi_mod = None
        dev = req_to_token.device
        self.device = dev
        self.hq, self.hkv, self.d = hq, hkv, head_dim
        rows, cols = req_to_token.shape
        # the reference's own classes; their tensors are then pointed at the caller's (same bytes, no copies)
        self.r2t_pool = pool_mod.ReqToTokenPool(1, 1, str(dev), False)
        self.r2t_pool.size, self.r2t_pool.max_context_len = rows, cols
        self.r2t_pool.req_to_token = req_to_token
        self.r2t_pool.free_slots = list(range(rows))
        self.kv_pool = pool_mod.MHATokenToKVPool(0, 1, kv_dtype, hkv, head_dim, len(k_buffers), str(dev), False)
        self.kv_pool.size = k_buffers[0].shape[0] - 1
        self.kv_pool.k_buffer = list(k_buffers)
        self.kv_pool.v_buffer = list(v_buffers)
        self.kv_dtype = kv_dtype
        ra_mod = ref_loader.load("nn.attention.radix_attention")
        scaling = head_dim ** -0.5 if sm_scale is None else sm_scale
        self.layers = []
        for i in range(len(k_buffers)):
            layer = ra_mod.RadixAttention(hq, head_dim, scaling, hkv, i, logit_cap=logit_cap)
            layer.k_scale_float, layer.v_scale_float = k_scale, v_scale  # shim (1) of the module docstring
            self.layers.append(layer)
        self.fwd_mod = ref_loader.load("model_executor.forward_info")

    def runner(self, context_len: int):
        return _ref_runner(self.hq, self.hkv, self.d, 1, context_len, self.kv_dtype, str(self.device), self.r2t_pool,
                           self.kv_pool, self.r2t_pool.size, self.model_dtype)

    def decode_batch(self, req_pool_indices, seq_lens, out_cache_loc, backend):
        F = self.fwd_mod
        return F.ForwardBatch(forward_mode=F.ForwardMode.DECODE, batch_size=len(seq_lens), input_ids=None,
                              req_pool_indices=req_pool_indices, seq_lens=seq_lens, out_cache_loc=out_cache_loc,
                              seq_lens_sum=int(seq_lens.sum().item()), req_to_token_pool=self.r2t_pool,
                              token_to_kv_pool=self.kv_pool, attn_backend=backend)

    def extend_batch(self, req_pool_indices, seq_lens, out_cache_loc, extend_seq_lens, extend_prefix_lens, backend):
        F = self.fwd_mod
        _, start = F.compute_position_torch(extend_prefix_lens, extend_seq_lens)
        return F.ForwardBatch(forward_mode=F.ForwardMode.EXTEND, batch_size=len(seq_lens), input_ids=None,
                              req_pool_indices=req_pool_indices, seq_lens=seq_lens, out_cache_loc=out_cache_loc,
                              seq_lens_sum=int(seq_lens.sum().item()), extend_num_tokens=int(extend_seq_lens.sum().item()),
                              extend_seq_lens=extend_seq_lens, extend_prefix_lens=extend_prefix_lens,
                              extend_start_loc=start, extend_prefix_lens_cpu=extend_prefix_lens.tolist(),
                              extend_seq_lens_cpu=extend_seq_lens.tolist(), req_to_token_pool=self.r2t_pool,
                              token_to_kv_pool=self.kv_pool, attn_backend=backend)


class RefFlashInfer:
    """The reference's FlashInferAttnBackend, constructed and driven as ModelRunner / RadixAttention do."""

    def __init__(self, state: RefState, context_len: int):
        mod = ref_loader.flashinfer_backend_module()
        mod.global_workspace_buffer = None  # module-level cache keyed by nothing: one buffer per comparator instance
        self.state = state
        self.backend = mod.FlashInferAttnBackend(state.runner(context_len))
        self.use_tensor_cores = bool(self.backend.decode_use_tensor_cores)

    def init_forward_metadata(self, fb):
        self.backend.init_forward_metadata(fb)

    def forward(self, layer_id: int, q, k, v, fb, save_kv_cache: bool = True):
        # RadixAttention.forward (radix_attention.py:41-57) -> AttentionBackend.forward (backend.py:54-81)
        return self.state.layers[layer_id](q, k, v, fb, save_kv_cache)


def triton_decode(q, k_buf, v_buf, req_to_token, req_pool_indices, seq_lens, sm_scale, logit_cap: float = 0.0,
                  attn_logits=None):
    """In-tree ``decode_attention_fwd`` (decode_attention.py:547-608): q [B,Hq,D] bf16; returns o [B,Hq,D]."""
    dec, _ = ref_loader.triton_kernels()
    b, hq, _ = q.shape
    sl32 = seq_lens.to(torch.int32)
    start = torch.zeros_like(sl32)
    start[1:] = torch.cumsum(sl32[:-1], 0)
    total = int(sl32.sum().item())
    if attn_logits is None:
        attn_logits = torch.empty((hq, total), dtype=torch.float32, device=q.device)
    o = torch.empty_like(q)
    dec.decode_attention_fwd(q, k_buf, v_buf, o, req_to_token, req_pool_indices.to(torch.int32), start, sl32, attn_logits,
                             int(sl32.max().item()), sm_scale, logit_cap)
    return o


def triton_extend(q, k_new, v_new, k_buf, v_buf, req_to_token, req_pool_indices, seq_lens, extend_seq_lens,
                  extend_start_loc, sm_scale, logit_cap: float = 0.0):
    """In-tree ``extend_attention_fwd`` (extend_attention.py:229-327): q [T,Hq,D], k_new/v_new [T,Hkv,D] contiguous."""
    _, ext = ref_loader.triton_kernels()
    o = torch.empty_like(q)
    ext.extend_attention_fwd(q, k_new, v_new, o, k_buf, v_buf, req_to_token, req_pool_indices.to(torch.int32),
                             seq_lens.to(torch.int32), extend_seq_lens.to(torch.int32), extend_start_loc.to(torch.int32),
                             int(extend_seq_lens.max().item()), sm_scale, logit_cap)
    return o


class TrtllmGenDecode:
    """FlashInfer's ``trtllm-gen`` paged decode (SURVEY.md 8d "stretch" comparator): NVIDIA's own Blackwell-native
    (tcgen05 / TMA) FMHA cubins shipped in ``flashinfer-cubin``, reached through
    ``flashinfer.decode.trtllm_batch_decode_with_kv_cache``.  Not a path the reference engine takes -- it needs pages of
    >= 16 tokens, the reference runs page size 1 -- but the strongest library kernel for this op on this GPU, so it is
    timed on the very pool bytes the b200 backend reads when the pool is paged: the ``[P + page, Hkv, D]`` buffer is
    viewed (no copy) as ``[pages, page, Hkv, D]`` (``kv_layout="NHD"``), the block table is every ``page``-th entry of
    the request's ``req_to_token`` row divided by the page size."""

    def __init__(self, req_to_token: torch.Tensor, req_pool_indices: torch.Tensor, seq_lens: torch.Tensor, page_size: int):
        from . import fi_env
        fi_env.patch_jit()
        import flashinfer.decode as fd
        self.fn = fd.trtllm_batch_decode_with_kv_cache
        self.page = int(page_size)
        self.seq_lens = seq_lens.to(torch.int32).contiguous()
        self.max_seq_len = int(seq_lens.max().item())
        npages = (self.max_seq_len + self.page - 1) // self.page
        rows = req_to_token[req_pool_indices.long()]
        self.block_tables = (rows[:, : npages * self.page : self.page] // self.page).to(torch.int32).contiguous()
        self.workspace = torch.zeros(128 << 20, dtype=torch.uint8, device=req_to_token.device)

    def paged_view(self, buf: torch.Tensor) -> torch.Tensor:
        rows, hkv, d = buf.shape
        assert rows % self.page == 0 and buf.is_contiguous()
        return buf.view(rows // self.page, self.page, hkv, d)

    def __call__(self, q3: torch.Tensor, k_buf: torch.Tensor, v_buf: torch.Tensor, sm_scale: float) -> torch.Tensor:
        return self.fn(q3, (self.paged_view(k_buf), self.paged_view(v_buf)), self.workspace, self.block_tables,
                       self.seq_lens, self.max_seq_len, bmm1_scale=float(sm_scale), bmm2_scale=1.0, kv_layout="NHD",
                       backend="trtllm-gen")


def time_cuda(fn, iters: int, warmup: int = 3):
    """Mean seconds per call of ``fn(i)``, CUDA events on the current stream."""
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / iters
