"""CPU checks against the UNMODIFIED reference files in baseline/_ref: this repo's host-side mirrors (the
AttentionBackend ABC in the product package; ForwardMode / pools / RadixAttention stand-ins in harness/) agree with the
classes they restate.  Skipped where baseline/_ref is not installed (python baseline/install_ref.py)."""
import inspect

import pytest
import torch

import harness
import scratchpad_b200 as sb
from baseline import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="baseline/_ref not installed")


def test_backend_abc_signatures_equal_the_reference_file():
    ref = ref_loader.load("nn.attention.backend").AttentionBackend
    for name, fn in inspect.getmembers(ref, inspect.isfunction):
        if name.startswith("_"):
            continue
        assert hasattr(sb.AttentionBackend, name), name
        want = list(inspect.signature(fn).parameters)
        got = list(inspect.signature(getattr(sb.AttentionBackend, name)).parameters)
        # `rope` is this backend's one optional extension of forward(); everything else is the reference's order
        assert [p for p in got if p != "rope"][:len(want)] == want, (name, got, want)
    assert inspect.isabstract(ref) and inspect.isabstract(sb.AttentionBackend)


def test_forward_mode_and_positions_equal_the_reference():
    F = ref_loader.load("model_executor.forward_info")
    for name in ("EXTEND", "DECODE", "MIXED", "IDLE", "TARGET_VERIFY", "DRAFT_EXTEND", "DUMMY_FIRST"):
        a, b = getattr(F.ForwardMode, name), getattr(harness.ForwardMode, name)
        assert int(a) == int(b)
        for pred in ("is_extend", "is_decode", "is_mixed", "is_idle", "is_cuda_graph", "is_decode_or_idle"):
            assert bool(getattr(a, pred)()) == bool(getattr(b, pred)()), (name, pred)
    pre, ext = torch.tensor([3, 0, 11], dtype=torch.int32), torch.tensor([2, 4, 1], dtype=torch.int32)
    _, start = F.compute_position_torch(pre, ext)
    assert torch.equal(start, harness.compute_extend_start_loc(ext)) and start.dtype == torch.int32
    fields = {f for f in F.ForwardBatch.__dataclass_fields__}
    for f in harness.ForwardBatch.__dataclass_fields__:
        assert f in fields, f  # the stand-in never invents a field the engine's batch lacks


def test_pool_stand_ins_behave_like_the_reference_pools():
    P = ref_loader.load("memory.pool")
    a, b = P.ReqToTokenPool(5, 12, "cpu", False), harness.ReqToTokenPool(5, 12, "cpu")
    assert a.req_to_token.shape == b.req_to_token.shape and a.req_to_token.dtype == b.req_to_token.dtype
    assert a.alloc(3) == b.alloc(3) and a.alloc(9) is None and b.alloc(9) is None
    a.free([1]); b.free([1])
    assert a.available_size() == b.available_size()
    kv_ref = P.MHATokenToKVPool(32, 1, torch.float8_e4m3fn, 2, 64, 2, "cpu", False)
    kv_own = sb.MHATokenToKVPool(32, 1, torch.float8_e4m3fn, 2, 64, 2, "cpu")
    for x, y in ((kv_ref.k_buffer[1], kv_own.k_buffer[1]), (kv_ref.get_value_buffer(0), kv_own.get_value_buffer(0))):
        assert x.shape == y.shape and x.dtype == y.dtype and x.stride() == y.stride()
    r, o = P.TokenToKVPoolAllocator(32, torch.bfloat16, "cpu", kv_ref), harness.TokenToKVPoolAllocator(32, torch.bfloat16, "cpu", kv_own)
    assert torch.equal(r.alloc(4), o.alloc(4)) and r.alloc(40) is None and o.alloc(40) is None
    r.free(torch.tensor([2, 3])); o.free(torch.tensor([2, 3]))
    assert torch.equal(r.free_slots, o.free_slots)


def test_radix_attention_stand_in_has_the_reference_attributes():
    RA = ref_loader.load("nn.attention.radix_attention").RadixAttention
    a, b = RA(8, 128, 0.1, 2, 3, logit_cap=30.0), harness.RadixAttention(8, 128, 0.1, 2, 3, logit_cap=30.0)
    for attr in ("tp_q_head_num", "tp_k_head_num", "tp_v_head_num", "head_dim", "qk_head_dim", "v_head_dim", "scaling",
                 "layer_id", "logit_cap", "sliding_window_size", "is_cross_attention"):
        assert getattr(a, attr) == getattr(b, attr), attr


def test_reference_flashinfer_policy_matches_design_table():
    """should_use_tensor_core (flashinfer_backend.py:936-981), the reference's own evidence for 'dense enough'."""
    try:
        m = ref_loader.flashinfer_backend_module()
    except ImportError:
        pytest.skip("flashinfer is not importable")
    assert m.should_use_tensor_core(torch.bfloat16, 32, 8) is False      # C2 / C3B: group 4
    assert m.should_use_tensor_core(torch.float8_e4m3fn, 32, 8) is True  # C3: fp8
    assert m.should_use_tensor_core(torch.bfloat16, 8, 1) is True        # C5 heads: group 8
