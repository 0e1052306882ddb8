// Extend / chunked-prefill attention over prefix-cached chunks for B200 (sm_100a).
//
// Replaces, behind AttentionBackend.forward_extend, what the reference does with
// triteia/Triton `extend_attention_fwd` (scratchpad/nn/attention/triton_backend.py:118-155;
// math as in triton_attn/extend_attention.py:53-226) and FlashInfer's ragged + paged
// prefill wrappers merged by LSE (flashinfer_backend.py:419-439).  Semantics: query
// row t of request i attends, non-causally, to the prefix slots req_to_token[r, :pre_i]
// gathered from the pool and, causally, to the new tokens k_new/v_new[start_i : start_i+t+1]
// taken from the call arguments (never re-read from the pool; for an fp8 pool the prefix
// is dequantised, the new tokens stay in model precision).
//
// Design (this file: the generic mma.sync path, any group size / D in {64,128} / fp8 pool):
//   * GQA query-group packing: a CTA owns 128 query rows = (128/g tokens) x (g q heads of
//     one kv head), so each K/V tile staged in shared memory is used by all g heads --
//     the reference's kernel re-streams K/V once per q head (extend_attention.py:267-288).
//   * One pass over [prefix | new] keys in 64-key tiles through a 3-stage cp.async ring,
//     16-byte chunks XOR-swizzled for conflict-free ldmatrix; prefix rows are gathered
//     through the token index (slot ids prefetched one tile ahead), new rows are linear.
//   * FA2-style register pipeline per warp (16 rows): S = Q.K^T (m16n8k16, fp32 accum),
//     online softmax in the exp2 domain, P re-used in registers as the A operand of P.V.
//   * Heavy (late, long-causal) query blocks are scheduled first.

#include "common.cuh"

namespace b200 {

constexpr int kExtStages = 3;
constexpr int kExtRows = 128;  // query rows per CTA (tokens x group heads)
constexpr int kExtBN = 64;     // keys per tile
constexpr int kExtNW = 8;
constexpr int kExtNT = kExtNW * 32;

struct ExtendArgs {
  const uint16_t* q;  // bf16 or fp16 bits (act_dtype), like k_new / v_new / o
  int64_t q_st, q_sh;
  const uint16_t* k_new;
  const uint16_t* v_new;
  int64_t kn_st, kn_sh, vn_st, vn_sh;
  const uint8_t* kbuf;
  const uint8_t* vbuf;
  int64_t k_st_bytes, v_st_bytes;
  uint16_t* o;
  int64_t o_st, o_sh;
  const int32_t* r2t;
  int64_t r2t_stride;
  const void* req_idx;
  const void* seq_lens;
  int idx64;
  const int32_t* ext_lens;
  const int32_t* ext_start;
  int hq, hkv, g, bm_t;  // bm_t = tokens per CTA = 128 / g
  float sm_scale, logit_cap;
  const float* k_scale;
  const float* v_scale;
};

// HALF_: the model dtype (q, new tokens, o, and a 16-bit pool) is fp16 instead of bf16
template <int D_, bool FP8_, bool HALF_ = false>
struct ExtCfg {
  static constexpr int D = D_;
  static constexpr bool FP8 = FP8_;
  static constexpr bool HALF = HALF_;
  static constexpr int SEG = D * 2;        // bytes of one smem row (always the model dtype in smem)
  static constexpr int CPS = SEG / 16;     // 16-byte chunks per smem row
  static constexpr int KC = D / 16;        // k-steps of Q.K^T
  static constexpr int TILE_BYTES = kExtBN * SEG;
  static constexpr int STAGE_BYTES = 2 * TILE_BYTES;
  static constexpr int SMEM_BYTES = kExtStages * STAGE_BYTES;
  static constexpr int PASSES = kExtBN * CPS / kExtNT;       // cp.async passes per tile (bf16 rows)
  static constexpr int PSEG = D;                              // fp8 pool row bytes per head
  static constexpr int PCH = PSEG / 16;                       // 16-byte chunks of an fp8 pool row
  static constexpr int FPASSES = (kExtBN * PCH + kExtNT - 1) / kExtNT;  // staged fp8 loads per thread
  static_assert(kExtRows * SEG <= STAGE_BYTES, "Q tile is staged in stage 0");
  static_assert(PASSES >= 1 && (kExtBN * CPS) % kExtNT == 0, "copy mapping");
};

// 16 e4m3 bytes -> 16 model-dtype values (x scale), as two 16-byte chunks
template <bool HALF>
__device__ __forceinline__ void dequant16(const uint4& raw, float scale, uint4& lo, uint4& hi) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
  uint32_t out[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t h01 = e4m3x2_to_f16x2(static_cast<uint16_t>(w[i] & 0xffffu));
    const uint32_t h23 = e4m3x2_to_f16x2(static_cast<uint16_t>(w[i] >> 16));
    const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&h01));
    const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&h23));
    out[2 * i] = pack2<HALF>(f01.x * scale, f01.y * scale);
    out[2 * i + 1] = pack2<HALF>(f23.x * scale, f23.y * scale);
  }
  lo = make_uint4(out[0], out[1], out[2], out[3]);
  hi = make_uint4(out[4], out[5], out[6], out[7]);
}

template <class C>
__global__ void __launch_bounds__(kExtNT, 1) extend_mma_kernel(const ExtendArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int D = C::D, SEG = C::SEG, CPS = C::CPS, BN = kExtBN;
  constexpr bool FP8 = C::FP8, HALF = C::HALF;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, cid = lane & 3;
  const int b = blockIdx.z, kvh = blockIdx.y;
  pdl_wait();
  pdl_launch_dependents();
  const int G = a.g, bm_t = a.bm_t;

  const int ext = __ldg(a.ext_lens + b);
  const int nqb = (ext + bm_t - 1) / bm_t;
  const int qb = nqb - 1 - static_cast<int>(blockIdx.x);  // heavy (late) blocks first
  if (qb < 0) return;
  const int seq = static_cast<int>(load_idx(a.seq_lens, a.idx64, b));
  const int pre = seq - ext;
  const int start = __ldg(a.ext_start + b);
  const int32_t* row = a.r2t + load_idx(a.req_idx, a.idx64, b) * a.r2t_stride;
  const int t0 = qb * bm_t;
  const int t_end = min(ext, t0 + bm_t);
  const int kv_len = pre + t_end;  // keys visible to this CTA
  const int ntiles = (kv_len + BN - 1) / BN;
  const int live_rows = (t_end - t0) * G;

  const uint32_t smem_base = smem_u32(smem);
  const float ks = (FP8 && a.k_scale) ? __ldg(a.k_scale + kvh) : 1.0f;
  const float vs = (FP8 && a.v_scale) ? __ldg(a.v_scale + kvh) : 1.0f;

  // ---- Q tile -> stage 0 -> A fragments -------------------------------------------------
  for (int idx = tid; idx < kExtRows * CPS; idx += kExtNT) {
    const int r = idx / CPS, ch = idx % CPS;
    const bool live = r < live_rows;
    const int t = t0 + (live ? r / G : 0), h = kvh * G + (live ? r % G : 0);
    const uint16_t* src = a.q + static_cast<int64_t>(start + t) * a.q_st + static_cast<int64_t>(h) * a.q_sh + ch * 8;
    cp_async16(smem_base + (r * CPS + swz_chunk<CPS>(ch, r)) * 16, src, live ? 16 : 0);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  uint32_t qf[C::KC][4];
  {
    const int r = warp * 16 + (lane & 15);
#pragma unroll
    for (int kc = 0; kc < C::KC; ++kc)
      ldmatrix_x4(qf[kc], smem_base + (r * CPS + swz_chunk<CPS>(2 * kc + (lane >> 4), r)) * 16);
  }
  __syncthreads();

  // ---- loader ----------------------------------------------------------------------------
  int slot_reg[2];
  auto prefetch_slots = [&](int t) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int p = t * BN + j * 32 + lane;
      slot_reg[j] = (p < pre) ? __ldg(row + p) : 0;
    }
  };
  const int64_t pool_head_off = static_cast<int64_t>(kvh) * (FP8 ? C::PSEG : SEG);
  uint4 stg_k[FP8 ? C::FPASSES : 1], stg_v[FP8 ? C::FPASSES : 1];

  auto issue_tile = [&](int t, int stage) {
    const uint32_t sk = smem_base + stage * C::STAGE_BYTES;
    const uint32_t sv = sk + C::TILE_BYTES;
#pragma unroll
    for (int i = 0; i < C::PASSES; ++i) {
      const int idx = tid + i * kExtNT;
      const int kr = idx / CPS, ch = idx % CPS;
      const int p = t * BN + kr;
      const int slot = __shfl_sync(0xffffffffu, (kr >> 5) ? slot_reg[1] : slot_reg[0], kr & 31);
      const uint32_t doff = (kr * CPS + swz_chunk<CPS>(ch, kr)) * 16;
      if (p >= pre || !FP8) {
        const void* ksrc;
        const void* vsrc;
        if (p < pre) {
          ksrc = kv_row_ptr(a.kbuf, slot, a.k_st_bytes) + pool_head_off + ch * 16;
          vsrc = kv_row_ptr(a.vbuf, slot, a.v_st_bytes) + pool_head_off + ch * 16;
        } else {
          const int64_t n = start + min(p, kv_len - 1) - pre;
          ksrc = a.k_new + n * a.kn_st + static_cast<int64_t>(kvh) * a.kn_sh + ch * 8;
          vsrc = a.v_new + n * a.vn_st + static_cast<int64_t>(kvh) * a.vn_sh + ch * 8;
        }
        const int nb = p < kv_len ? 16 : 0;
        cp_async16(sk + doff, ksrc, nb);
        cp_async16(sv + doff, vsrc, nb);
      }
    }
    if constexpr (FP8) {  // prefix rows of an fp8 pool: staged through registers, stored by store_staged
#pragma unroll
      for (int i = 0; i < C::FPASSES; ++i) {
        const int idx = tid + i * kExtNT;
        const int kr = idx / C::PCH, ch = idx % C::PCH;
        const int p = t * BN + kr;
        const int slot = __shfl_sync(0xffffffffu, ((kr >> 5) & 1) ? slot_reg[1] : slot_reg[0], kr & 31);
        if (kr < BN && p < pre) {
          stg_k[i] = __ldg(reinterpret_cast<const uint4*>(kv_row_ptr(a.kbuf, slot, a.k_st_bytes) + pool_head_off + ch * 16));
          stg_v[i] = __ldg(reinterpret_cast<const uint4*>(kv_row_ptr(a.vbuf, slot, a.v_st_bytes) + pool_head_off + ch * 16));
        }
      }
    }
  };
  auto store_staged = [&](int t, int stage) {
    if constexpr (FP8) {
      uint8_t* sk = smem + stage * C::STAGE_BYTES;
      uint8_t* sv = sk + C::TILE_BYTES;
#pragma unroll
      for (int i = 0; i < C::FPASSES; ++i) {
        const int idx = tid + i * kExtNT;
        const int kr = idx / C::PCH, ch = idx % C::PCH;
        const int p = t * BN + kr;
        if (kr < BN && p < pre) {
          uint4 lo, hi;
          dequant16<HALF>(stg_k[i], ks, lo, hi);
          *reinterpret_cast<uint4*>(sk + (kr * CPS + swz_chunk<CPS>(2 * ch, kr)) * 16) = lo;
          *reinterpret_cast<uint4*>(sk + (kr * CPS + swz_chunk<CPS>(2 * ch + 1, kr)) * 16) = hi;
          dequant16<HALF>(stg_v[i], vs, lo, hi);
          *reinterpret_cast<uint4*>(sv + (kr * CPS + swz_chunk<CPS>(2 * ch, kr)) * 16) = lo;
          *reinterpret_cast<uint4*>(sv + (kr * CPS + swz_chunk<CPS>(2 * ch + 1, kr)) * 16) = hi;
        }
      }
    }
  };

  // ---- per-warp softmax state -------------------------------------------------------------
  const int r0 = warp * 16 + gid, r1 = r0 + 8;
  // causal limit of each row: key position <= pre + token index (dead rows behave like the last live token)
  const int lim0 = pre + min(t0 + r0 / G, t_end - 1);
  const int lim1 = pre + min(t0 + r1 / G, t_end - 1);
  const float cap = a.logit_cap;
  const float sc_log2 = a.sm_scale * kLog2e;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;

  auto compute_tile = [&](int t, int stage) {
    const uint32_t sk = smem_base + stage * C::STAGE_BYTES;
    const uint32_t sv = sk + C::TILE_BYTES;
    float s[BN / 8][4];
#pragma unroll
    for (int i = 0; i < BN / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    {
      const int krow = (lane & 7) + ((lane >> 4) << 3);
      const int kch = (lane >> 3) & 1;
#pragma unroll
      for (int kc = 0; kc < C::KC; ++kc) {
#pragma unroll
        for (int np = 0; np < BN / 16; ++np) {
          uint32_t bf[4];
          const int kr = np * 16 + krow;
          ldmatrix_x4(bf, sk + (kr * CPS + swz_chunk<CPS>(2 * kc + kch, kr)) * 16);
          mma_16816<HALF>(s[2 * np], qf[kc], bf[0], bf[1]);
          mma_16816<HALF>(s[2 * np + 1], qf[kc], bf[2], bf[3]);
        }
      }
    }
    if (cap > 0.f) {
#pragma unroll
      for (int i = 0; i < BN / 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = softcap(s[i][j] * a.sm_scale, cap) * kLog2e;
    } else {
#pragma unroll
      for (int i = 0; i < BN / 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] *= sc_log2;
    }
    const int pbase = t * BN + 2 * cid;
    if ((t + 1) * BN - 1 > pre + t0) {  // tile reaches the causal diagonal (or the ragged end): mask
#pragma unroll
      for (int i = 0; i < BN / 8; ++i) {
        const int p = pbase + i * 8;
        if (p > lim0) s[i][0] = -INFINITY;
        if (p + 1 > lim0) s[i][1] = -INFINITY;
        if (p > lim1) s[i][2] = -INFINITY;
        if (p + 1 > lim1) s[i][3] = -INFINITY;
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int i = 0; i < BN / 8; ++i) {
      mx0 = fmaxf(mx0, fmaxf(s[i][0], s[i][1]));
      mx1 = fmaxf(mx1, fmaxf(s[i][2], s[i][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    // a row can be fully masked in a tile only after it has seen key 0, so mn is finite from tile 0 on
    const float al0 = ex2(m0 - mn0), al1 = ex2(m1 - mn1);
    m0 = mn0;
    m1 = mn1;
    l0 *= al0;
    l1 *= al1;
#pragma unroll
    for (int i = 0; i < BN / 8; ++i) {
      s[i][0] = ex2(s[i][0] - mn0);
      s[i][1] = ex2(s[i][1] - mn0);
      s[i][2] = ex2(s[i][2] - mn1);
      s[i][3] = ex2(s[i][3] - mn1);
      l0 += s[i][0] + s[i][1];
      l1 += s[i][2] + s[i][3];
    }
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      o[i][0] *= al0;
      o[i][1] *= al0;
      o[i][2] *= al1;
      o[i][3] *= al1;
    }
    {
      const int vrow = (lane & 7) + (((lane >> 3) & 1) << 3);
      const int vch = lane >> 4;
#pragma unroll
      for (int ksx = 0; ksx < BN / 16; ++ksx) {
        uint32_t pa[4];
        pa[0] = pack2<HALF>(s[2 * ksx][0], s[2 * ksx][1]);
        pa[1] = pack2<HALF>(s[2 * ksx][2], s[2 * ksx][3]);
        pa[2] = pack2<HALF>(s[2 * ksx + 1][0], s[2 * ksx + 1][1]);
        pa[3] = pack2<HALF>(s[2 * ksx + 1][2], s[2 * ksx + 1][3]);
        const int kr = ksx * 16 + vrow;
#pragma unroll
        for (int dp = 0; dp < D / 16; ++dp) {
          uint32_t bf[4];
          ldmatrix_x4_trans(bf, sv + (kr * CPS + swz_chunk<CPS>(2 * dp + vch, kr)) * 16);
          mma_16816<HALF>(o[2 * dp], pa, bf[0], bf[1]);
          mma_16816<HALF>(o[2 * dp + 1], pa, bf[2], bf[3]);
        }
      }
    }
  };

  // ---- pipeline -----------------------------------------------------------------------------
  prefetch_slots(0);
#pragma unroll
  for (int s = 0; s < kExtStages - 1; ++s) {
    if (s < ntiles) {
      issue_tile(s, s);
      store_staged(s, s);
      prefetch_slots(s + 1);
    }
    cp_async_commit();
  }
  for (int t = 0; t < ntiles; ++t) {
    cp_async_wait<kExtStages - 2>();
    __syncthreads();
    const int tn = t + kExtStages - 1;
    const int sn = tn % kExtStages;
    if (tn < ntiles) {
      issue_tile(tn, sn);
      prefetch_slots(tn + 1);
    }
    cp_async_commit();
    compute_tile(t, t % kExtStages);
    if (tn < ntiles) store_staged(tn, sn);
  }
  cp_async_wait<0>();

  // ---- epilogue -------------------------------------------------------------------------------
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
  if (r0 < live_rows) {
    uint16_t* op = a.o + static_cast<int64_t>(start + t0 + r0 / G) * a.o_st + static_cast<int64_t>(kvh * G + r0 % G) * a.o_sh + 2 * cid;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) *reinterpret_cast<uint32_t*>(op + i * 8) = pack2<HALF>(o[i][0] * inv0, o[i][1] * inv0);
  }
  if (r1 < live_rows) {
    uint16_t* op = a.o + static_cast<int64_t>(start + t0 + r1 / G) * a.o_st + static_cast<int64_t>(kvh * G + r1 % G) * a.o_sh + 2 * cid;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) *reinterpret_cast<uint32_t*>(op + i * 8) = pack2<HALF>(o[i][2] * inv1, o[i][3] * inv1);
  }
}

template <class C>
static int launch_extend_t(const ExtendArgs& a, dim3 grid, cudaStream_t st) {
  auto kern = extend_mma_kernel<C>;
  static PerDevice pd;  // per instantiation
  if (int rc = ensure_smem_attr(kern, C::SMEM_BYTES, pd)) return rc;
  return launch_pdl(kern, grid, dim3(kExtNT), C::SMEM_BYTES, st, a);
}

int launch_extend_tc(const b200_extend_params* p, cudaStream_t st);   // extend_tc.cu: one CTA per 256-row tile
int launch_extend_tc2(const b200_extend_params* p, cudaStream_t st);  // extend_tc2.cu: CTA pairs (cta_group::2)

}  // namespace b200

using namespace b200;

extern "C" int b200_attn_extend(const b200_extend_params* p, b200_stream_t stream) {
  if (!p || !p->q || !p->k_new || !p->v_new || !p->k_buf || !p->v_buf || !p->o || !p->req_to_token ||
      !p->req_pool_indices || !p->seq_lens || !p->extend_seq_lens || !p->extend_start_loc)
    return B200_ERR_INVALID;
  if (p->bs <= 0 || p->hq <= 0 || p->hkv <= 0 || p->hq % p->hkv != 0 || p->total_tokens < 0) return B200_ERR_INVALID;
  if (p->total_tokens == 0 || p->max_extend_len <= 0) return B200_OK;
  if (p->d != p->dv || (p->d != 64 && p->d != 128)) return B200_ERR_UNSUPPORTED;
  const int g = p->hq / p->hkv;
  if (g > kExtRows) return B200_ERR_UNSUPPORTED;
  const bool fp8 = p->kv_dtype == B200_KV_FP8_E4M3, half = p->act_dtype == B200_ACT_F16;
  if (!dtype_pair_ok(p->kv_dtype, p->act_dtype)) return B200_ERR_UNSUPPORTED;
  const int elt = fp8 ? 1 : 2;
  auto mis16 = [](const void* x) { return (reinterpret_cast<uintptr_t>(x) & 15) != 0; };
  if ((p->k_stride_t * elt) % 16 || (p->v_stride_t * elt) % 16 || p->q_stride_t % 8 || p->q_stride_h % 8 ||
      p->kn_stride_t % 8 || p->kn_stride_h % 8 || p->vn_stride_t % 8 || p->vn_stride_h % 8 || p->o_stride_t % 2 ||
      p->o_stride_h % 2 || mis16(p->q) || mis16(p->k_new) || mis16(p->v_new) || mis16(p->k_buf) || mis16(p->v_buf) ||
      (reinterpret_cast<uintptr_t>(p->o) & 3))
    return B200_ERR_INVALID;
  // tcgen05/TMEM path: head_dim 128 or 64 (bf16 or fp8-e4m3 pool), 16-byte aligned output rows; auto-selected once the
  // longest extend fills at least half of its 256-row CTA tile (short MIXED-mode rows stay on mma.sync)
  // (k_new / v_new heads contiguous inside a token row, as in the fused qkv GEMM output)
  const bool tc_ok = (p->d == 128 || p->d == 64) && p->o_stride_t % 8 == 0 && p->o_stride_h % 8 == 0 && !mis16(p->o) && g <= 128 &&
                     p->kn_stride_h == p->d && p->vn_stride_h == p->dv;
  const bool pair_ok = tc_ok && p->d == 128 && !fp8 && !half;
  if (p->force_path == B200_PATH_TC && !tc_ok) return B200_ERR_UNSUPPORTED;
  if (p->force_path == B200_PATH_TC2 && !pair_ok) return B200_ERR_UNSUPPORTED;
  if (tc_ok && (p->force_path == B200_PATH_TC || p->force_path == B200_PATH_TC2 ||
                (p->force_path < 0 && static_cast<int64_t>(p->max_extend_len) * g >= 128))) {
    // The CTA-pair kernel (tcgen05 cta_group::2, extend_tc2.cu) is correct and tested but measured SLOWER than the
    // one-CTA kernel (C4: 925 vs 1180 TFLOP/s), so it is opt-in: B200_PATH_TC2, or B200_EXTEND_PAIR=1 to
    // make the automatic choice take it where it exists (head_dim 128 over a bf16 pool).
    static const bool pair_auto = getenv("B200_EXTEND_PAIR") != nullptr && getenv("B200_EXTEND_PAIR")[0] == '1';
    const bool pair = pair_ok && (p->force_path == B200_PATH_TC2 || (p->force_path < 0 && pair_auto));
    return pair ? launch_extend_tc2(p, static_cast<cudaStream_t>(stream)) : launch_extend_tc(p, static_cast<cudaStream_t>(stream));
  }
  ExtendArgs a;
  a.q = static_cast<const uint16_t*>(p->q);
  a.q_st = p->q_stride_t;
  a.q_sh = p->q_stride_h;
  a.k_new = static_cast<const uint16_t*>(p->k_new);
  a.v_new = static_cast<const uint16_t*>(p->v_new);
  a.kn_st = p->kn_stride_t;
  a.kn_sh = p->kn_stride_h;
  a.vn_st = p->vn_stride_t;
  a.vn_sh = p->vn_stride_h;
  a.kbuf = static_cast<const uint8_t*>(p->k_buf);
  a.vbuf = static_cast<const uint8_t*>(p->v_buf);
  a.k_st_bytes = p->k_stride_t * elt;
  a.v_st_bytes = p->v_stride_t * elt;
  a.o = static_cast<uint16_t*>(p->o);
  a.o_st = p->o_stride_t;
  a.o_sh = p->o_stride_h;
  a.r2t = p->req_to_token;
  a.r2t_stride = p->req_to_token_stride;
  a.req_idx = p->req_pool_indices;
  a.seq_lens = p->seq_lens;
  a.idx64 = p->idx_dtype == B200_IDX_I64;
  a.ext_lens = p->extend_seq_lens;
  a.ext_start = p->extend_start_loc;
  a.hq = p->hq;
  a.hkv = p->hkv;
  a.g = g;
  a.bm_t = kExtRows / g;
  a.sm_scale = p->sm_scale;
  a.logit_cap = p->logit_cap;
  a.k_scale = p->k_scale;
  a.v_scale = p->v_scale;
  const int nqb = (p->max_extend_len + a.bm_t - 1) / a.bm_t;
  const dim3 grid(nqb, p->hkv, p->bs);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (half) {
    if (p->d == 64) return fp8 ? launch_extend_t<ExtCfg<64, true, true>>(a, grid, st) : launch_extend_t<ExtCfg<64, false, true>>(a, grid, st);
    return fp8 ? launch_extend_t<ExtCfg<128, true, true>>(a, grid, st) : launch_extend_t<ExtCfg<128, false, true>>(a, grid, st);
  }
  if (p->d == 64) return fp8 ? launch_extend_t<ExtCfg<64, true>>(a, grid, st) : launch_extend_t<ExtCfg<64, false>>(a, grid, st);
  return fp8 ? launch_extend_t<ExtCfg<128, true>>(a, grid, st) : launch_extend_t<ExtCfg<128, false>>(a, grid, st);
}
