// KV write into the paged pool: k_buf[loc[t]] = quant(k[t]).
// Replaces MHATokenToKVPool.set_kv_buffer (scratchpad/memory/pool.py:392-424):
// two index_put_ launches plus, for fp8, an in-place div_ on the caller's
// tensor and a cast.  Here K and V go in one launch, the source is not
// modified, and the fp8 value is e4m3_rn(act_rn(x / scale[head])) -- the same
// double rounding the reference's in-place div_ on the model-dtype tensor (bf16
// or fp16) followed by .to(fp8) produces; overflow becomes NaN exactly as
// torch's cast does (no saturation).

#include "common.cuh"

namespace b200 {

template <bool FP8, bool HALF>
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
      *reinterpret_cast<uint2*>(row + static_cast<int64_t>(h) * dd + c * 8) = quant_e4m3x8<HALF>(x, s);
    }
  }
}

// ---------------------------------------------------------------------------
// Rotary embedding fused with the KV write (b200_rope_kv_store).
// Work item = one 16-element unit of one (token, head) row: q heads are rotated in place, k heads are rotated and
// stored (quantised for fp8 pools), v heads are stored.  Arithmetic follows RotaryEmbedding.forward_native
// (rotary_embedding.py:24-49): products and sums each rounded to bf16, so the result is bit-identical.
// ---------------------------------------------------------------------------
// HALF: the model dtype is fp16 (else bf16).  torch computes 16-bit elementwise ops in fp32 and rounds the result to
// the tensor dtype, which is what round_act(__fmul_rn / __fsub_rn / __fadd_rn) restates for either dtype.
struct Vec8 {
  float f[8];
};
template <bool HALF>
__device__ __forceinline__ Vec8 unpack8(uint4 x) {
  const uint32_t w[4] = {x.x, x.y, x.z, x.w};
  Vec8 r;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    r.f[2 * j] = act_bits_to_float<HALF>(static_cast<uint16_t>(w[j] & 0xffffu));
    r.f[2 * j + 1] = act_bits_to_float<HALF>(static_cast<uint16_t>(w[j] >> 16));
  }
  return r;
}
template <bool HALF>
__device__ __forceinline__ uint4 pack8(const Vec8& v) {  // values are already representable in the model dtype
  return make_uint4(pack2<HALF>(v.f[0], v.f[1]), pack2<HALF>(v.f[2], v.f[3]), pack2<HALF>(v.f[4], v.f[5]),
                    pack2<HALF>(v.f[6], v.f[7]));
}
// 8 cos (or sin) values of a cache row starting at element e, as model-dtype-valued floats
template <bool HALF>
__device__ __forceinline__ Vec8 load_cs8(const void* row, int f32, int e) {
  if (!f32) return unpack8<HALF>(__ldg(reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(row) + e)));
  const float4 a = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(row) + e));
  const float4 b = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(row) + e + 4));
  Vec8 r = {{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
#pragma unroll
  for (int j = 0; j < 8; ++j) r.f[j] = round_act<HALF>(r.f[j]);  // cos.to(x.dtype)
  return r;
}
// (x1, x2) -> (x1*cos - x2*sin, x2*cos + x1*sin), each op rounded to the model dtype (no fma contraction)
template <bool HALF>
__device__ __forceinline__ void rotate_pair(float& x1, float& x2, float c, float s) {
  const float a = round_act<HALF>(__fmul_rn(x1, c)), b = round_act<HALF>(__fmul_rn(x2, s));
  const float e = round_act<HALF>(__fmul_rn(x2, c)), f = round_act<HALF>(__fmul_rn(x1, s));
  x1 = round_act<HALF>(__fsub_rn(a, b));
  x2 = round_act<HALF>(__fadd_rn(e, f));
}

template <bool FP8, bool HALF>
__global__ void __launch_bounds__(256) rope_kv_store_kernel(const b200_rope_store_params p, int64_t k_st_bytes,
                                                            int64_t v_st_bytes) {
  pdl_wait();
  pdl_launch_dependents();
  const int uq = p.q ? p.hq * (p.d / 16) : 0, uk = p.hkv * (p.d / 16), uv = p.v_buf ? p.hkv * (p.dv / 16) : 0;
  const int per_tok = uq + uk + uv;
  const int64_t total = static_cast<int64_t>(p.num_tokens) * per_tok;
  const int half = p.rotary_dim / 2, rot_units = p.rotary_dim / 16;
  const int units_per_row = p.d / 16;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i / per_tok);
    int r = static_cast<int>(i % per_tok);
    if (r >= uq + uk) {  // ---- v: plain store of 16 elements
      r -= uq + uk;
      const int h = r / (p.dv / 16), c = r % (p.dv / 16);
      const uint16_t* src = static_cast<const uint16_t*>(p.v) + t * p.v_src_stride_t + h * p.v_src_stride_h + c * 16;
      const uint4 x0 = __ldg(reinterpret_cast<const uint4*>(src)), x1 = __ldg(reinterpret_cast<const uint4*>(src + 8));
      uint8_t* row = static_cast<uint8_t*>(p.v_buf) + load_idx(p.loc, p.loc_dtype == B200_IDX_I64, t) * v_st_bytes;
      if constexpr (!FP8) {
        uint4* dst = reinterpret_cast<uint4*>(row + (static_cast<int64_t>(h) * p.dv + c * 16) * 2);
        dst[0] = x0;
        dst[1] = x1;
      } else {
        const float s = p.v_scale ? __ldg(p.v_scale + h) : 1.0f;
        const uint2 a = quant_e4m3x8<HALF>(x0, s), b = quant_e4m3x8<HALF>(x1, s);
        *reinterpret_cast<uint4*>(row + static_cast<int64_t>(h) * p.dv + c * 16) = make_uint4(a.x, a.y, b.x, b.y);
      }
      continue;
    }
    const bool is_k = r >= uq;
    if (is_k) r -= uq;
    const int h = r / units_per_row, u = r % units_per_row;
    uint16_t* base = is_k ? static_cast<uint16_t*>(p.k) + t * p.k_src_stride_t + h * p.k_src_stride_h
                          : static_cast<uint16_t*>(p.q) + t * p.q_stride_t + h * p.q_stride_h;
    // the two 8-element vectors of this unit: a rotated pair of vectors, or 16 pass-through elements
    int e0, e1;
    const bool rot = u < rot_units;
    if (!rot) {
      e0 = p.rotary_dim + (u - rot_units) * 16;
      e1 = e0 + 8;
    } else if (p.is_neox) {
      e0 = u * 8;
      e1 = half + u * 8;
    } else {
      e0 = u * 16;
      e1 = e0 + 8;
    }
    uint4 r0 = *reinterpret_cast<const uint4*>(base + e0), r1 = *reinterpret_cast<const uint4*>(base + e1);
    if (rot) {
      const int64_t pos = load_idx(p.positions, p.pos_dtype == B200_IDX_I64, t);
      const int f32 = p.cache_dtype == B200_CACHE_F32;
      const void* crow = static_cast<const uint8_t*>(p.cos_sin_cache) + pos * p.cache_stride * (f32 ? 4 : 2);
      const Vec8 c = load_cs8<HALF>(crow, f32, u * 8), s = load_cs8<HALF>(crow, f32, half + u * 8);
      Vec8 a = unpack8<HALF>(r0), b = unpack8<HALF>(r1);
      if (p.is_neox) {
#pragma unroll
        for (int j = 0; j < 8; ++j) rotate_pair<HALF>(a.f[j], b.f[j], c.f[j], s.f[j]);
      } else {  // pairs (2j, 2j+1) of 16 consecutive elements; cos/sin index u*8 + j
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          rotate_pair<HALF>(a.f[2 * j], a.f[2 * j + 1], c.f[j], s.f[j]);
          rotate_pair<HALF>(b.f[2 * j], b.f[2 * j + 1], c.f[4 + j], s.f[4 + j]);
        }
      }
      r0 = pack8<HALF>(a);
      r1 = pack8<HALF>(b);
      if (!is_k || p.write_back_k) {
        *reinterpret_cast<uint4*>(base + e0) = r0;
        *reinterpret_cast<uint4*>(base + e1) = r1;
      }
    }
    if (is_k && p.k_buf) {
      uint8_t* row = static_cast<uint8_t*>(p.k_buf) + load_idx(p.loc, p.loc_dtype == B200_IDX_I64, t) * k_st_bytes;
      if constexpr (!FP8) {
        uint8_t* dst = row + static_cast<int64_t>(h) * p.d * 2;
        *reinterpret_cast<uint4*>(dst + e0 * 2) = r0;
        *reinterpret_cast<uint4*>(dst + e1 * 2) = r1;
      } else {
        const float s = p.k_scale ? __ldg(p.k_scale + h) : 1.0f;
        uint8_t* dst = row + static_cast<int64_t>(h) * p.d;
        *reinterpret_cast<uint2*>(dst + e0) = quant_e4m3x8<HALF>(r0, s);
        *reinterpret_cast<uint2*>(dst + e1) = quant_e4m3x8<HALF>(r1, s);
      }
    }
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_kv_store(void* k_buf, void* v_buf, int64_t k_stride_t, int64_t v_stride_t, const void* loc,
                             int32_t loc_dtype, const void* k, const void* v, int64_t k_src_stride_t,
                             int64_t k_src_stride_h, int64_t v_src_stride_t, int64_t v_src_stride_h,
                             int32_t num_tokens, int32_t hkv, int32_t d, int32_t dv, int32_t kv_dtype,
                             int32_t act_dtype, const float* k_scale, const float* v_scale, b200_stream_t stream) {
  if (!k_buf || !v_buf || !loc || !k || !v || num_tokens < 0 || hkv <= 0) return B200_ERR_INVALID;
  if (num_tokens == 0) return B200_OK;
  if (d % 8 || dv % 8 || k_src_stride_t % 8 || k_src_stride_h % 8 || v_src_stride_t % 8 || v_src_stride_h % 8 ||
      (reinterpret_cast<uintptr_t>(k) & 15) || (reinterpret_cast<uintptr_t>(v) & 15) ||
      (reinterpret_cast<uintptr_t>(k_buf) & 15) || (reinterpret_cast<uintptr_t>(v_buf) & 15))
    return B200_ERR_INVALID;
  const bool fp8 = kv_dtype == B200_KV_FP8_E4M3, half = act_dtype == B200_ACT_F16;
  if (!dtype_pair_ok(kv_dtype, act_dtype)) return B200_ERR_UNSUPPORTED;
  const int elt = fp8 ? 1 : 2;
  if ((k_stride_t * elt) % 16 || (v_stride_t * elt) % 16) return B200_ERR_INVALID;
  const int64_t total = static_cast<int64_t>(num_tokens) * hkv * (d / 8 + dv / 8);
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // a 16-bit pool takes the activation's bits as they are, whichever 16-bit type that is
  auto kern = fp8 ? (half ? kv_store_kernel<true, true> : kv_store_kernel<true, false>) : kv_store_kernel<false, false>;
  return launch_pdl(kern, dim3(blocks), dim3(256), 0, st, static_cast<uint8_t*>(k_buf), static_cast<uint8_t*>(v_buf),
                    static_cast<int64_t>(k_stride_t * elt), static_cast<int64_t>(v_stride_t * elt), loc,
                    static_cast<int>(loc_dtype == B200_IDX_I64), static_cast<const uint16_t*>(k),
                    static_cast<const uint16_t*>(v), k_src_stride_t, k_src_stride_h, v_src_stride_t, v_src_stride_h,
                    static_cast<int>(num_tokens), static_cast<int>(hkv), static_cast<int>(d), static_cast<int>(dv),
                    k_scale, v_scale);
}

extern "C" int b200_rope_kv_store(const b200_rope_store_params* p, b200_stream_t stream) {
  if (!p || !p->k || !p->positions || !p->cos_sin_cache || p->num_tokens < 0 || p->hkv <= 0) return B200_ERR_INVALID;
  if ((p->k_buf == nullptr) != (p->v_buf == nullptr)) return B200_ERR_INVALID;
  if (p->k_buf && (!p->loc || !p->v)) return B200_ERR_INVALID;
  if (p->q && p->hq <= 0) return B200_ERR_INVALID;
  if (p->num_tokens == 0) return B200_OK;
  const bool fp8 = p->kv_dtype == B200_KV_FP8_E4M3, half = p->act_dtype == B200_ACT_F16;
  // rotate-only calls (k_buf == NULL) carry no pool: only the activation dtype matters
  if (p->k_buf ? !dtype_pair_ok(p->kv_dtype, p->act_dtype) : (p->act_dtype != B200_ACT_BF16 && !half)) return B200_ERR_UNSUPPORTED;
  // a 16-bit cos/sin cache is in the model dtype (the module builds it with .to(dtype), rotary_embedding.py:90-99)
  if (p->cache_dtype != B200_CACHE_F32 && p->cache_dtype != (half ? B200_CACHE_F16 : B200_CACHE_BF16)) return B200_ERR_UNSUPPORTED;
  if (p->d % 16 || p->dv % 16 || p->rotary_dim % 16 || p->rotary_dim <= 0 || p->rotary_dim > p->d) return B200_ERR_UNSUPPORTED;
  auto mis = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) != 0; };
  if (mis(p->k) || mis(p->q) || mis(p->v) || mis(p->k_buf) || mis(p->v_buf) || mis(p->cos_sin_cache) ||
      p->k_src_stride_t % 8 || p->k_src_stride_h % 8 || (p->q && (p->q_stride_t % 8 || p->q_stride_h % 8)) ||
      (p->v && (p->v_src_stride_t % 8 || p->v_src_stride_h % 8)) || p->cache_stride % 8)
    return B200_ERR_INVALID;
  const int elt = fp8 ? 1 : 2;
  if (p->k_buf && ((p->k_stride_t * elt) % 16 || (p->v_stride_t * elt) % 16)) return B200_ERR_INVALID;
  const int per_tok = (p->q ? p->hq * (p->d / 16) : 0) + p->hkv * (p->d / 16) + (p->v_buf ? p->hkv * (p->dv / 16) : 0);
  const int64_t total = static_cast<int64_t>(p->num_tokens) * per_tok;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  auto kern = fp8 ? (half ? rope_kv_store_kernel<true, true> : rope_kv_store_kernel<true, false>)
                  : (half ? rope_kv_store_kernel<false, true> : rope_kv_store_kernel<false, false>);
  return launch_pdl(kern, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), *p,
                    static_cast<int64_t>(p->k_stride_t * elt), static_cast<int64_t>(p->v_stride_t * elt));
}
