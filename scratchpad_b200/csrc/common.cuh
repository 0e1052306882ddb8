// Shared device helpers for the B200 (sm_100a) paged-attention kernels.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "b200_attn.h"

namespace b200 {

#define B200_CUDA_TRY(expr)                                   \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return 1000 + static_cast<int>(_e); \
  } while (0)

constexpr float kLog2e = 1.4426950408889634f;

// plan header words shared by decode.cu and extend_tc.cu: 7 = longest prefix any prefix group hands to the dense
// shared-prefix leg (0 = nothing to do), 8 = number of row tiles of that leg
constexpr int kHdrShared = 7, kHdrCasTiles = 8;

// device tables the plan writes for the shared-prefix leg (decode.cu) and extend_tc.cu reads
struct CasTables {
  const int* hdr;
  const int* req32;      // request -> req_to_token row
  const int* perm;       // sorted position -> request (requests of a prefix group are adjacent)
  const int* ct_first;   // per row tile: first sorted position, number of requests, a request whose row holds the
  const int* ct_cnt;     //   group's shared slots, and the shared length (multiple of 128)
  const int* ct_leader;
  const int* ct_shared;
  int max_tiles;
};

// ---- dtype pairing ------------------------------------------------------------
// The pool holds the model dtype (kv_cache_dtype "auto", model_runner.py:366-367) or fp8-e4m3: bf16 activations go
// with a bf16 or fp8 pool, fp16 activations with an fp16 or fp8 pool.
static inline bool dtype_pair_ok(int kv_dtype, int act_dtype) {
  if (act_dtype == B200_ACT_BF16) return kv_dtype == B200_KV_BF16 || kv_dtype == B200_KV_FP8_E4M3;
  if (act_dtype == B200_ACT_F16) return kv_dtype == B200_KV_F16 || kv_dtype == B200_KV_FP8_E4M3;
  return false;
}

// ---- index helpers ---------------------------------------------------------
__device__ __forceinline__ int64_t load_idx(const void* p, int idx64, int i) {
  return idx64 ? static_cast<const int64_t*>(p)[i] : static_cast<int64_t>(static_cast<const int32_t*>(p)[i]);
}

// The one place a KV slot id becomes an address.  Shared by the decode/extend
// kernels and by b200_debug_gather, which is how the bit-exact gather contract
//  is tested: row = buf + slot * token_stride_bytes.
__device__ __forceinline__ const uint8_t* kv_row_ptr(const uint8_t* buf, int32_t slot, int64_t stride_t_bytes) {
  return buf + static_cast<int64_t>(slot) * stride_t_bytes;
}

// ---- programmatic dependent launch (PDL) ---------------------------------------
// Every kernel of the path is launched with programmaticStreamSerializationAllowed and starts with
// griddepcontrol.wait (returns once the preceding grid has completed and flushed), so the launch
// latency and CTA scheduling of kernel N+1 overlap the tail of kernel N.  No-ops without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline int launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                             Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static const bool no_pdl = getenv("B200_NO_PDL") != nullptr;  // measurement switch: plain stream-ordered launches
  cfg.numAttrs = no_pdl ? 0 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  if (e != cudaSuccess) return 1000 + static_cast<int>(e);
  return B200_OK;
}

// ---- per-device one-time function attributes ------------------------------------------------------------------
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and occupancy are properties of (kernel, device): a process that
// drives several GPUs (one backend instance per device) must set / query them once per device, not once per process.
constexpr int kMaxDevices = 64;
struct PerDevice {
  bool attr_done[kMaxDevices] = {};
  int occupancy[kMaxDevices] = {};  // 0 = not queried yet
};
static inline int current_device(int* dev) {
  B200_CUDA_TRY(cudaGetDevice(dev));
  if (*dev < 0 || *dev >= kMaxDevices) return B200_ERR_UNSUPPORTED;
  return B200_OK;
}
template <class K>
static inline int ensure_smem_attr(K kern, int smem_bytes, PerDevice& pd) {
  int dev = 0;
  if (int rc = current_device(&dev)) return rc;
  if (!pd.attr_done[dev]) {
    B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    pd.attr_done[dev] = true;
  }
  return B200_OK;
}

// ---- PTX wrappers ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 16-byte async global->shared copy (LDGSTS), L2-only; src_bytes = 0 zero-fills.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n"
               : "=r"(r[0]), "=r"(r[1])
               : "r"(addr));
}
// 8x8 b16 transpose across the warp: lane i holds (row i/4, cols 2(i%4),+1).
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;\n" : "=r"(d) : "r"(a));
  return d;
}

// D(16x8,f32) += A(16x16) * B(16x8); A row-major, B "col" (k pairs contiguous).
template <bool F16>
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (F16) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
// cap * tanh(x / cap), the reference's soft cap (extend_attention.py:10-13,138-139)
__device__ __forceinline__ float softcap(float x, float cap) {
  // tanh(y) = 1 - 2 / (exp(2y) + 1); exact-form keeps us within 1e-6 of torch.tanh
  float y = x / cap;
  float e = __expf(2.0f * y);
  float t = 1.0f - __fdividef(2.0f, e + 1.0f);
  return cap * t;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <bool F16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  if constexpr (F16) return pack_f16x2(lo, hi);
  else return pack_bf16x2(lo, hi);
}
// two e4m3 bytes (low 16 bits of x) -> f16x2, in register (F2FP.F16.E4M3.UNPACK_B)
__device__ __forceinline__ uint32_t e4m3x2_to_f16x2(uint16_t x) {
  uint32_t d;
  asm("cvt.rn.f16x2.e4m3x2 %0, %1;\n" : "=r"(d) : "h"(x));
  return d;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;\n" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) { return __uint_as_float(static_cast<uint32_t>(b) << 16); }
__device__ __forceinline__ float e4m3_to_float(uint8_t b) {
  __half_raw h = __nv_cvt_fp8_to_halfraw(static_cast<__nv_fp8_storage_t>(b), __NV_E4M3);
  return __half2float(*reinterpret_cast<__half*>(&h));
}

__device__ __forceinline__ float f16_bits_to_float(uint16_t b) { return __half2float(__ushort_as_half(b)); }
// 16-bit activation (HALF: fp16, else bf16) <-> float; round_act = the value rounded to the model dtype
template <bool HALF>
__device__ __forceinline__ float act_bits_to_float(uint16_t b) {
  if constexpr (HALF) return f16_bits_to_float(b);
  else return bf16_bits_to_float(b);
}
template <bool HALF>
__device__ __forceinline__ float round_act(float x) {
  if constexpr (HALF) return __half2float(__float2half_rn(x));
  else return __bfloat162float(__float2bfloat16_rn(x));
}

// activation -> pool byte: e4m3_rn(act_rn(x / scale)), the double rounding of the reference's in-place div_ on the
// model-dtype tensor followed by .to(fp8) (memory/pool.py:404-412); overflow becomes NaN exactly as torch's cast does
// (no saturation)
template <bool HALF>
__device__ __forceinline__ uint8_t quant_e4m3(uint16_t act_bits, float scale) {
  const float x = act_bits_to_float<HALF>(act_bits);
  const float y = round_act<HALF>(__fdiv_rn(x, scale));
  return static_cast<uint8_t>(__nv_cvt_float_to_fp8(y, __NV_NOSAT, __NV_E4M3));
}
// 8 activations (one uint4) -> 8 e4m3 bytes
template <bool HALF>
__device__ __forceinline__ uint2 quant_e4m3x8(uint4 x, float scale) {
  const uint32_t w[4] = {x.x, x.y, x.z, x.w};
  uint32_t out[2] = {0u, 0u};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint16_t bits = static_cast<uint16_t>(j & 1 ? w[j >> 1] >> 16 : w[j >> 1] & 0xffffu);
    out[j >> 2] |= static_cast<uint32_t>(quant_e4m3<HALF>(bits, scale)) << (8 * (j & 3));
  }
  return make_uint2(out[0], out[1]);
}

// XOR swizzle of 16-byte chunks inside one (token, head) segment so that the 8
// rows an ldmatrix phase touches land in 8 different bank groups.
// CPS = 16-byte chunks per segment (4, 8 or 16).
template <int CPS>
__device__ __forceinline__ int swz_chunk(int chunk, int row) {
  if constexpr (CPS >= 8) return chunk ^ (row & 7);
  else return chunk ^ ((row >> 1) & 3);
}

}  // namespace b200
