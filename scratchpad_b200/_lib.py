"""ctypes binding of libb200attn.so (the C ABI declared in include/b200_attn.h).

The library is built in-tree by ``scratchpad_b200/csrc/build.sh`` (nvcc,
sm_100a).  There is no fallback: if the shared object is missing or a symbol
is absent, importing the hot path fails loudly.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200_ATTN_LIB selects another build of the same library (kernel-tuning variants, the trace build); never a fallback
LIB_PATH = os.environ.get("B200_ATTN_LIB") or os.path.join(_HERE, "csrc", "libb200attn.so")
BUILD_SCRIPT = os.path.join(_HERE, "csrc", "build.sh")

ABI_VERSION = 10

KV_BF16, KV_FP8_E4M3, KV_F16 = 0, 1, 2
ACT_BF16, ACT_F16 = 0, 1
IDX_I32, IDX_I64 = 0, 1
PATH_AUTO, PATH_MMA, PATH_TC, PATH_TC2 = -1, 0, 1, 2
CACHE_BF16, CACHE_F32, CACHE_F16 = 0, 1, 2
PAGE_AUTO = -1

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_WORKSPACE, ERR_NOT_PLANNED = 0, 1, 2, 3, 4

EXPORTS = (
    "b200_attn_abi_version",
    "b200_attn_struct_size",
    "b200_attn_error_string",
    "b200_attn_get_config",
    "b200_attn_plan",
    "b200_attn_decode",
    "b200_attn_extend",
    "b200_kv_store",
    "b200_rope_kv_store",
    "b200_debug_last_tma_mode",
    "b200_debug_set_dyn_min_base",
    "b200_debug_sched",
    "b200_debug_tma_tiles",
    "b200_debug_extend_tma_tiles",
    "b200_debug_check_paged",
    "b200_debug_gather",
    "b200_debug_decode_naive",
)


class AttnConfig(C.Structure):
    _fields_ = [
        ("path", C.c_int32),
        ("tile_tokens", C.c_int32),
        ("grid_ctas", C.c_int32),
        ("heads_per_cta", C.c_int32),
        ("sub_splits", C.c_int32),
        ("stages", C.c_int32),
        ("smem_bytes", C.c_int32),
        ("num_sms", C.c_int32),
        ("partials_per_slot", C.c_int32),
        ("q_tiles", C.c_int32),
        ("workspace_bytes", C.c_uint64),
    ]


class DecodeParams(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("q_stride_t", C.c_int64), ("q_stride_h", C.c_int64),
        ("k_buf", C.c_void_p), ("v_buf", C.c_void_p),
        ("k_stride_t", C.c_int64), ("v_stride_t", C.c_int64),
        ("o", C.c_void_p), ("o_stride_t", C.c_int64), ("o_stride_h", C.c_int64),
        ("req_to_token", C.c_void_p), ("req_to_token_stride", C.c_int64),
        ("bs", C.c_int32), ("hq", C.c_int32), ("hkv", C.c_int32), ("d", C.c_int32), ("dv", C.c_int32),
        ("kv_dtype", C.c_int32),
        ("sm_scale", C.c_float), ("logit_cap", C.c_float),
        ("k_scale", C.c_void_p), ("v_scale", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64),
        ("max_bs", C.c_int32), ("force_path", C.c_int32),
        ("ev_kernel_begin", C.c_void_p), ("ev_kernel_end", C.c_void_p),
        ("cascade", C.c_int32), ("page_size", C.c_int32), ("kv_pool_rows", C.c_int64),
        ("act_dtype", C.c_int32), ("reserved0", C.c_int32),
    ]


class ExtendParams(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("q_stride_t", C.c_int64), ("q_stride_h", C.c_int64),
        ("k_new", C.c_void_p), ("v_new", C.c_void_p),
        ("kn_stride_t", C.c_int64), ("kn_stride_h", C.c_int64),
        ("vn_stride_t", C.c_int64), ("vn_stride_h", C.c_int64),
        ("k_buf", C.c_void_p), ("v_buf", C.c_void_p),
        ("k_stride_t", C.c_int64), ("v_stride_t", C.c_int64),
        ("o", C.c_void_p), ("o_stride_t", C.c_int64), ("o_stride_h", C.c_int64),
        ("req_to_token", C.c_void_p), ("req_to_token_stride", C.c_int64),
        ("req_pool_indices", C.c_void_p), ("seq_lens", C.c_void_p), ("idx_dtype", C.c_int32),
        ("extend_seq_lens", C.c_void_p), ("extend_start_loc", C.c_void_p),
        ("bs", C.c_int32), ("hq", C.c_int32), ("hkv", C.c_int32), ("d", C.c_int32), ("dv", C.c_int32),
        ("total_tokens", C.c_int32), ("max_extend_len", C.c_int32),
        ("kv_dtype", C.c_int32),
        ("sm_scale", C.c_float), ("logit_cap", C.c_float),
        ("k_scale", C.c_void_p), ("v_scale", C.c_void_p),
        ("force_path", C.c_int32), ("page_size", C.c_int32), ("kv_pool_rows", C.c_int64),
        ("act_dtype", C.c_int32), ("reserved0", C.c_int32),
    ]


class RopeStoreParams(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("q_stride_t", C.c_int64), ("q_stride_h", C.c_int64),
        ("k", C.c_void_p), ("k_src_stride_t", C.c_int64), ("k_src_stride_h", C.c_int64),
        ("v", C.c_void_p), ("v_src_stride_t", C.c_int64), ("v_src_stride_h", C.c_int64),
        ("positions", C.c_void_p), ("cos_sin_cache", C.c_void_p), ("cache_stride", C.c_int64),
        ("cache_dtype", C.c_int32), ("pos_dtype", C.c_int32), ("rotary_dim", C.c_int32), ("is_neox", C.c_int32),
        ("k_buf", C.c_void_p), ("v_buf", C.c_void_p), ("k_stride_t", C.c_int64), ("v_stride_t", C.c_int64),
        ("loc", C.c_void_p), ("loc_dtype", C.c_int32),
        ("num_tokens", C.c_int32), ("hq", C.c_int32), ("hkv", C.c_int32), ("d", C.c_int32), ("dv", C.c_int32),
        ("kv_dtype", C.c_int32),
        ("k_scale", C.c_void_p), ("v_scale", C.c_void_p),
        ("write_back_k", C.c_int32), ("act_dtype", C.c_int32),
    ]


class B200AttnError(RuntimeError):
    def __init__(self, code: int, what: str, where: str):
        super().__init__(f"{where} failed: [{code}] {what}")
        self.code = code


_lib = None


def build(verbose: bool = False) -> str:
    """Compile the extension in-tree (nvcc, -gencode arch=compute_100a,code=sm_100a)."""
    env = dict(os.environ)
    if verbose:
        env["B200_PTXAS_V"] = "1"
    subprocess.run(["bash", BUILD_SCRIPT], check=True, env=env)
    return LIB_PATH


def load() -> C.CDLL:
    """Load libb200attn.so; raises if it is absent or stale (no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `bash {BUILD_SCRIPT}` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "The cuda_b200 attention backend has no CPU or Triton fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise ImportError(f"{LIB_PATH} does not export {name}; rebuild it")
    lib.b200_attn_abi_version.restype = C.c_int
    lib.b200_attn_error_string.restype = C.c_char_p
    lib.b200_attn_error_string.argtypes = [C.c_int]
    if lib.b200_attn_abi_version() != ABI_VERSION:
        raise ImportError(
            f"{LIB_PATH} has ABI {lib.b200_attn_abi_version()}, binding expects {ABI_VERSION}; rebuild it"
        )
    lib.b200_attn_struct_size.restype = C.c_uint64
    lib.b200_attn_struct_size.argtypes = [C.c_int]
    for which, cls in enumerate((AttnConfig, DecodeParams, ExtendParams, RopeStoreParams)):
        if lib.b200_attn_struct_size(which) != C.sizeof(cls):
            raise ImportError(f"{cls.__name__}: ctypes layout ({C.sizeof(cls)} B) != compiled layout "
                              f"({lib.b200_attn_struct_size(which)} B)")
    i32, i64, u64, vp, f = C.c_int32, C.c_int64, C.c_uint64, C.c_void_p, C.c_float
    lib.b200_attn_get_config.argtypes = [i32, i32, i32, i32, i32, i32, i32, C.POINTER(AttnConfig)]
    lib.b200_attn_plan.argtypes = [vp, u64, i32, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, i64, i32, vp]
    lib.b200_attn_decode.argtypes = [C.POINTER(DecodeParams), vp]
    lib.b200_attn_extend.argtypes = [C.POINTER(ExtendParams), vp]
    lib.b200_kv_store.argtypes = [vp, vp, i64, i64, vp, i32, vp, vp, i64, i64, i64, i64, i32, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.b200_rope_kv_store.argtypes = [C.POINTER(RopeStoreParams), vp]
    lib.b200_debug_check_paged.argtypes = [vp, i64, vp, vp, i32, i32, i32, vp, vp]
    lib.b200_debug_gather.argtypes = [vp, i64, i32, vp, i64, vp, vp, i32, vp, i32, vp, vp]
    lib.b200_debug_decode_naive.argtypes = [C.POINTER(DecodeParams), vp, vp, i32, vp]
    for name in EXPORTS[3:]:
        getattr(lib, name).restype = C.c_int
    lib.b200_debug_set_dyn_min_base.argtypes = [i32]
    lib.b200_debug_sched.argtypes = [i32, i32, i32, i32, C.POINTER(C.c_int32)]
    lib.b200_debug_tma_tiles.restype = C.c_int64
    lib.b200_debug_tma_tiles.argtypes = [i32]
    lib.b200_debug_extend_tma_tiles.restype = C.c_int64
    lib.b200_debug_extend_tma_tiles.argtypes = [i32]
    _lib = lib
    return lib


def check(code: int, where: str) -> None:
    if code != 0:
        raise B200AttnError(code, load().b200_attn_error_string(code).decode(), where)
