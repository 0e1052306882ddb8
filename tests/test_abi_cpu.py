"""CPU-side checks: the C-ABI library loads and exports every symbol include/b200_attn.h
declares (no compute calls without a GPU), and the host-side mirror keeps the reference's surface."""
import ctypes
import inspect
import os
import re

import pytest
import torch

import harness
import scratchpad_b200 as sb
from scratchpad_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "b200_attn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _header_functions()
    assert set(names) == set(_lib.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.b200_attn_abi_version() == _lib.ABI_VERSION
    assert b"unsupported" in lib.b200_attn_error_string(_lib.ERR_UNSUPPORTED)


def test_struct_sizes_match_the_header_layout():
    lib = _lib.load()
    for which, cls in enumerate((_lib.AttnConfig, _lib.DecodeParams, _lib.ExtendParams, _lib.RopeStoreParams)):
        assert lib.b200_attn_struct_size(which) == ctypes.sizeof(cls), cls


def test_calls_fail_loudly_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = _lib.AttnConfig()
    rc = _lib.load().b200_attn_get_config(8, 32, 8, 64, 64, _lib.KV_BF16, -1, cfg)
    assert rc >= 1000  # a cudaError_t, reported, not swallowed
    with pytest.raises(_lib.B200AttnError):
        _lib.check(rc, "b200_attn_get_config")


def test_backend_surface_matches_reference_abc():
    # names and argument order of scratchpad/nn/attention/backend.py:11-105
    want = {
        "init_forward_metadata": ["self", "forward_batch"],
        "init_cuda_graph_state": ["self", "max_bs"],
        "init_forward_metadata_capture_cuda_graph": ["self", "bs", "num_tokens", "req_pool_indices", "seq_lens",
                                                     "encoder_lens", "forward_mode", "spec_info"],
        "init_forward_metadata_replay_cuda_graph": ["self", "bs", "req_pool_indices", "seq_lens", "seq_lens_sum",
                                                    "encoder_lens", "forward_mode", "spec_info", "seq_lens_cpu"],
        "get_cuda_graph_seq_len_fill_value": ["self"],
        "forward": ["self", "q", "k", "v", "layer", "forward_batch", "save_kv_cache"],
        "forward_decode": ["self", "q", "k", "v", "layer", "forward_batch", "save_kv_cache"],
        "forward_extend": ["self", "q", "k", "v", "layer", "forward_batch", "save_kv_cache"],
    }
    ext = ("forward", "forward_decode", "forward_extend")  # CudaB200AttnBackend appends one optional keyword: rope
    for cls in (sb.AttentionBackend, sb.CudaB200AttnBackend):
        for name, args in want.items():
            got = list(inspect.signature(getattr(cls, name)).parameters)
            if name in ext and got[-1] == "rope":
                assert inspect.signature(getattr(cls, name)).parameters["rope"].default is None
                got = got[:-1]
            assert got == args, (cls, name)


def test_rope_store_argument_checks_run_before_any_launch():
    """b200_rope_kv_store validates on the host first: these return without touching a device."""
    lib = _lib.load()
    buf = (ctypes.c_uint8 * 4096)()
    base = (ctypes.addressof(buf) + 15) & ~15

    def params(**kw):
        p = _lib.RopeStoreParams(k=base, k_src_stride_t=128, k_src_stride_h=64, positions=base, cos_sin_cache=base,
                                 cache_stride=64, cache_dtype=_lib.CACHE_BF16, pos_dtype=_lib.IDX_I64, rotary_dim=64,
                                 is_neox=1, num_tokens=1, hq=0, hkv=2, d=64, dv=64, kv_dtype=_lib.KV_BF16)
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    assert lib.b200_rope_kv_store(None, None) == _lib.ERR_INVALID
    assert lib.b200_rope_kv_store(params(num_tokens=0), None) == _lib.OK          # nothing to do
    assert lib.b200_rope_kv_store(params(rotary_dim=24), None) == _lib.ERR_UNSUPPORTED
    assert lib.b200_rope_kv_store(params(rotary_dim=128), None) == _lib.ERR_UNSUPPORTED  # > head size
    assert lib.b200_rope_kv_store(params(k=base + 2), None) == _lib.ERR_INVALID   # misaligned
    assert lib.b200_rope_kv_store(params(k_buf=base), None) == _lib.ERR_INVALID   # k_buf without v_buf
    assert lib.b200_rope_kv_store(params(cache_dtype=5), None) == _lib.ERR_UNSUPPORTED
    # a 16-bit cos/sin cache is in the model dtype; a pool is the model dtype or fp8-e4m3
    assert lib.b200_rope_kv_store(params(act_dtype=_lib.ACT_F16), None) == _lib.ERR_UNSUPPORTED        # fp16 q/k, bf16 cache
    assert lib.b200_rope_kv_store(params(cache_dtype=_lib.CACHE_F16), None) == _lib.ERR_UNSUPPORTED    # bf16 q/k, fp16 cache
    assert lib.b200_rope_kv_store(params(act_dtype=7, cache_dtype=_lib.CACHE_F32), None) == _lib.ERR_UNSUPPORTED
    pooled = dict(k_buf=base, v_buf=base, v=base, loc=base, v_src_stride_t=128, v_src_stride_h=64, k_stride_t=128,
                  v_stride_t=128, cache_dtype=_lib.CACHE_F32)
    assert lib.b200_rope_kv_store(params(act_dtype=_lib.ACT_F16, kv_dtype=_lib.KV_BF16, **pooled), None) == _lib.ERR_UNSUPPORTED
    assert lib.b200_rope_kv_store(params(act_dtype=_lib.ACT_BF16, kv_dtype=_lib.KV_F16, **pooled), None) == _lib.ERR_UNSUPPORTED


def test_model_dtype_and_pool_dtype_pairing():
    """bf16 activations go with a bf16 or fp8-e4m3 pool, fp16 activations with an fp16 or fp8-e4m3 pool
    (kv_cache_dtype "auto" = the model dtype, model_runner.py:366-367); checked in Python and again behind the C ABI."""
    from scratchpad_b200.pool import act_dtype_code, check_dtype_pair, kv_dtype_code
    assert kv_dtype_code(torch.float16) == _lib.KV_F16 and act_dtype_code(torch.float16) == _lib.ACT_F16
    assert kv_dtype_code(torch.bfloat16) == _lib.KV_BF16 and act_dtype_code(torch.bfloat16) == _lib.ACT_BF16
    for kv, act in ((_lib.KV_BF16, _lib.ACT_BF16), (_lib.KV_FP8_E4M3, _lib.ACT_BF16), (_lib.KV_F16, _lib.ACT_F16),
                    (_lib.KV_FP8_E4M3, _lib.ACT_F16)):
        check_dtype_pair(kv, act)
    for kv, act in ((_lib.KV_F16, _lib.ACT_BF16), (_lib.KV_BF16, _lib.ACT_F16)):
        with pytest.raises(NotImplementedError):
            check_dtype_pair(kv, act)
    with pytest.raises(NotImplementedError):
        act_dtype_code(torch.float32)
    # b200_kv_store / b200_attn_decode / b200_attn_extend refuse the pairing on the host, before any launch
    lib = _lib.load()
    buf = (ctypes.c_uint8 * 4096)()
    base = (ctypes.addressof(buf) + 15) & ~15
    rc = lib.b200_kv_store(base, base, 128, 128, base, _lib.IDX_I64, base, base, 128, 64, 128, 64, 1, 2, 64, 64,
                           _lib.KV_BF16, _lib.ACT_F16, None, None, None)
    assert rc == _lib.ERR_UNSUPPORTED
    dp = _lib.DecodeParams(q=base, k_buf=base, v_buf=base, o=base, req_to_token=base, workspace=base, bs=1, max_bs=1,
                           hq=8, hkv=2, d=64, dv=64, kv_dtype=_lib.KV_F16, act_dtype=_lib.ACT_BF16)
    assert lib.b200_attn_decode(dp, None) == _lib.ERR_UNSUPPORTED
    ep = _lib.ExtendParams(q=base, k_new=base, v_new=base, k_buf=base, v_buf=base, o=base, req_to_token=base,
                           req_pool_indices=base, seq_lens=base, extend_seq_lens=base, extend_start_loc=base, bs=1,
                           hq=8, hkv=2, d=64, dv=64, total_tokens=4, max_extend_len=4, kv_dtype=_lib.KV_BF16,
                           act_dtype=_lib.ACT_F16)
    assert lib.b200_attn_extend(ep, None) == _lib.ERR_UNSUPPORTED


def test_rotary_module_builds_the_reference_cache():
    """get_rope -> cos_sin_cache identical to the reference module's (tests/golden/rope_*.npz hold the caches the
    reference built in the build container)."""
    from scratchpad_b200.rotary_embedding import get_rope
    from golden_util import golden_names, load_golden
    scaling = {"rope_type": "llama3", "factor": 32.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
               "original_max_position_embeddings": 8192}
    for name in golden_names("rope"):
        g = load_golden(name)
        dtype = torch.float16 if g.get("dtype", "bf16") == "f16" else torch.bfloat16  # ropeh_*: fp16 model dtype
        emb = get_rope(g["head_size"], g["rotary_dim"], g["max_position"], g["base"], bool(g["is_neox"]),
                       scaling if g["llama3"] else None, dtype=dtype)
        want = g["cos_sin_cache"].view(torch.int16)
        assert torch.equal(emb.cos_sin_cache.cpu().view(torch.int16), want), name
    with pytest.raises(RuntimeError):  # no CPU fallback
        emb(torch.zeros(1, dtype=torch.int64), torch.zeros(1, 64, dtype=torch.bfloat16),
            torch.zeros(1, 64, dtype=torch.bfloat16))


def test_forward_mode_values_and_dispatch():
    fm = harness.ForwardMode
    assert (fm.EXTEND, fm.DECODE, fm.MIXED, fm.IDLE) == (1, 2, 3, 4)
    assert fm.MIXED.is_extend() and not fm.MIXED.is_decode() and fm.DECODE.is_cuda_graph()

    class Probe(sb.AttentionBackend):
        def init_forward_metadata(self, fb): pass
        def forward_decode(self, *a, **k): return "decode"
        def forward_extend(self, *a, **k): return "extend"

    fb = harness.ForwardBatch(fm.MIXED, 1, torch.zeros(1), torch.zeros(1), torch.zeros(1))
    assert Probe().forward(None, None, None, None, fb) == "extend"   # EXTEND and MIXED both go to extend
    fb.forward_mode = fm.DECODE
    assert Probe().forward(None, None, None, None, fb) == "decode"


def test_pool_layout_and_no_cpu_fallback():
    pool = sb.MHATokenToKVPool(32, 1, torch.float8_e4m3fn, 2, 64, 2, "cpu")
    assert pool.k_buffer[0].shape == (33, 2, 64) and pool.k_buffer[0].dtype == torch.uint8
    assert pool.get_key_buffer(0).dtype == torch.float8_e4m3fn
    pool.set_kv_scales(1, 0.5, torch.tensor([0.25, 2.0]))
    ks, vs = pool.get_kv_scales(1)
    assert ks.tolist() == [0.5, 0.5] and vs.tolist() == [0.25, 2.0]
    layer = harness.RadixAttention(4, 64, 0.125, 2, 0)
    with pytest.raises(RuntimeError):
        pool.set_kv_buffer(layer, torch.tensor([1]), torch.zeros(1, 2, 64, dtype=torch.bfloat16),
                           torch.zeros(1, 2, 64, dtype=torch.bfloat16))
    alloc = harness.TokenToKVPoolAllocator(32, torch.bfloat16, "cpu", pool)
    got = alloc.alloc(4)
    assert got.dtype == torch.int64 and got.tolist() == [1, 2, 3, 4]  # slot 0 is reserved
    with pytest.raises(NotImplementedError):
        sb.MHATokenToKVPool(32, 24, torch.bfloat16, 2, 64, 1, "cpu")  # pages are 16, 32, 64 or 128 tokens


def test_synth_page_table_properties():
    from harness import synth
    gen = torch.Generator().manual_seed(0)
    r2t = torch.zeros((9, 300), dtype=torch.int32)
    lens = [260, 260, 290, 5]
    rows = synth.build_page_table(r2t, lens, 2000, gen, "random", shared_prefix=256)
    assert len(set(rows.tolist())) == 4
    assert torch.equal(r2t[rows[0], :256], r2t[rows[1], :256])       # shared prefix = equal ids
    priv = torch.cat([r2t[rows[i], 256:lens[i]] for i in range(3)] + [r2t[rows[3], :5]])
    assert priv.min() >= 1 and len(set(priv.tolist())) == priv.numel()  # slot 0 never handed out, no reuse
