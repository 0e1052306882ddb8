// KV write into the paged pool: k_buf[loc[t]] = quant(k[t]).
// Replaces MHATokenToKVPool.set_kv_buffer (scratchpad/memory/pool.py:392-424):
// two index_put_ launches plus, for fp8, an in-place div_ on the caller's
// tensor and a cast.  Here K and V go in one launch, the source is not
// modified, and the fp8 value is e4m3_rn(bf16_rn(x / scale[head])) -- the same
// double rounding the reference's bf16 div_ followed by .to(fp8) produces;
// overflow becomes NaN exactly as torch's cast does (no saturation).

#include "common.cuh"

namespace b200 {

template <bool FP8>
__global__ void __launch_bounds__(256) kv_store_kernel(uint8_t* __restrict__ kbuf, uint8_t* __restrict__ vbuf,
                                                       int64_t k_st_bytes, int64_t v_st_bytes,
                                                       const void* __restrict__ loc, int loc64,
                                                       const uint16_t* __restrict__ k, const uint16_t* __restrict__ v,
                                                       int64_t ks_t, int64_t ks_h, int64_t vs_t, int64_t vs_h,
                                                       int num_tokens, int hkv, int d, int dv,
                                                       const float* __restrict__ k_scale,
                                                       const float* __restrict__ v_scale) {
  pdl_wait();
  pdl_launch_dependents();
  const int vk = d / 8, vv = dv / 8;          // 8-element vectors per head row
  const int per_tok = hkv * (vk + vv);
  const int64_t total = static_cast<int64_t>(num_tokens) * per_tok;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i / per_tok);
    int r = static_cast<int>(i % per_tok);
    const bool is_v = r >= hkv * vk;
    if (is_v) r -= hkv * vk;
    const int nvec = is_v ? vv : vk;
    const int h = r / nvec, c = r % nvec;
    const int dd = is_v ? dv : d;
    const uint16_t* src = is_v ? v + t * vs_t + h * vs_h + c * 8 : k + t * ks_t + h * ks_h + c * 8;
    const uint4 x = __ldg(reinterpret_cast<const uint4*>(src));
    const int64_t slot = load_idx(loc, loc64, t);
    uint8_t* row = (is_v ? vbuf + slot * v_st_bytes : kbuf + slot * k_st_bytes);
    if constexpr (!FP8) {
      *reinterpret_cast<uint4*>(row + (static_cast<int64_t>(h) * dd + c * 8) * 2) = x;
    } else {
      const float* sp = is_v ? v_scale : k_scale;
      const float s = sp ? __ldg(sp + h) : 1.0f;
      *reinterpret_cast<uint2*>(row + static_cast<int64_t>(h) * dd + c * 8) = quant_e4m3x8(x, s);
    }
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_kv_store(void* k_buf, void* v_buf, int64_t k_stride_t, int64_t v_stride_t, const void* loc,
                             int32_t loc_dtype, const void* k, const void* v, int64_t k_src_stride_t,
                             int64_t k_src_stride_h, int64_t v_src_stride_t, int64_t v_src_stride_h,
                             int32_t num_tokens, int32_t hkv, int32_t d, int32_t dv, int32_t kv_dtype,
                             const float* k_scale, const float* v_scale, b200_stream_t stream) {
  if (!k_buf || !v_buf || !loc || !k || !v || num_tokens < 0 || hkv <= 0) return B200_ERR_INVALID;
  if (num_tokens == 0) return B200_OK;
  if (d % 8 || dv % 8 || k_src_stride_t % 8 || k_src_stride_h % 8 || v_src_stride_t % 8 || v_src_stride_h % 8 ||
      (reinterpret_cast<uintptr_t>(k) & 15) || (reinterpret_cast<uintptr_t>(v) & 15) ||
      (reinterpret_cast<uintptr_t>(k_buf) & 15) || (reinterpret_cast<uintptr_t>(v_buf) & 15))
    return B200_ERR_INVALID;
  const bool fp8 = kv_dtype == B200_KV_FP8_E4M3;
  if (!fp8 && kv_dtype != B200_KV_BF16) return B200_ERR_UNSUPPORTED;
  const int elt = fp8 ? 1 : 2;
  if ((k_stride_t * elt) % 16 || (v_stride_t * elt) % 16) return B200_ERR_INVALID;
  const int64_t total = static_cast<int64_t>(num_tokens) * hkv * (d / 8 + dv / 8);
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto kern = fp8 ? kv_store_kernel<true> : kv_store_kernel<false>;
  return launch_pdl(kern, dim3(blocks), dim3(256), 0, st, static_cast<uint8_t*>(k_buf), static_cast<uint8_t*>(v_buf),
                    static_cast<int64_t>(k_stride_t * elt), static_cast<int64_t>(v_stride_t * elt), loc,
                    static_cast<int>(loc_dtype == B200_IDX_I64), static_cast<const uint16_t*>(k),
                    static_cast<const uint16_t*>(v), k_src_stride_t, k_src_stride_h, v_src_stride_t, v_src_stride_h,
                    static_cast<int>(num_tokens), static_cast<int>(hkv), static_cast<int>(d), static_cast<int>(dv),
                    k_scale, v_scale);
}
