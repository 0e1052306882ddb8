"""CPU-side checks: the C-ABI library loads and exports every symbol include/b200_attn.h
declares (no compute calls without a GPU), and the host-side mirror keeps the reference's surface."""
import ctypes
import inspect
import os
import re

import pytest
import torch

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
    for which, cls in enumerate((_lib.AttnConfig, _lib.DecodeParams, _lib.ExtendParams)):
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
    for cls in (sb.AttentionBackend, sb.CudaB200AttnBackend):
        for name, args in want.items():
            assert list(inspect.signature(getattr(cls, name)).parameters) == args, (cls, name)


def test_forward_mode_values_and_dispatch():
    fm = sb.ForwardMode
    assert (fm.EXTEND, fm.DECODE, fm.MIXED, fm.IDLE) == (1, 2, 3, 4)
    assert fm.MIXED.is_extend() and not fm.MIXED.is_decode() and fm.DECODE.is_cuda_graph()

    class Probe(sb.AttentionBackend):
        def init_forward_metadata(self, fb): pass
        def forward_decode(self, *a, **k): return "decode"
        def forward_extend(self, *a, **k): return "extend"

    fb = sb.ForwardBatch(fm.MIXED, 1, torch.zeros(1), torch.zeros(1), torch.zeros(1))
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
    layer = sb.RadixAttention(4, 64, 0.125, 2, 0)
    with pytest.raises(RuntimeError):
        pool.set_kv_buffer(layer, torch.tensor([1]), torch.zeros(1, 2, 64, dtype=torch.bfloat16),
                           torch.zeros(1, 2, 64, dtype=torch.bfloat16))
    alloc = sb.TokenToKVPoolAllocator(32, torch.bfloat16, "cpu", pool)
    got = alloc.alloc(4)
    assert got.dtype == torch.int64 and got.tolist() == [1, 2, 3, 4]  # slot 0 is reserved
    with pytest.raises(NotImplementedError):
        sb.MHATokenToKVPool(32, 16, torch.bfloat16, 2, 64, 1, "cpu")


def test_synth_page_table_properties():
    from scratchpad_b200 import synth
    gen = torch.Generator().manual_seed(0)
    r2t = torch.zeros((9, 300), dtype=torch.int32)
    lens = [260, 260, 290, 5]
    rows = synth.build_page_table(r2t, lens, 2000, gen, "random", shared_prefix=256)
    assert len(set(rows.tolist())) == 4
    assert torch.equal(r2t[rows[0], :256], r2t[rows[1], :256])       # shared prefix = equal ids
    priv = torch.cat([r2t[rows[i], 256:lens[i]] for i in range(3)] + [r2t[rows[3], :5]])
    assert priv.min() >= 1 and len(set(priv.tolist())) == priv.numel()  # slot 0 never handed out, no reuse
