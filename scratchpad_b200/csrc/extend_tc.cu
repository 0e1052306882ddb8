// Extend / chunked-prefill attention, and the shared-prefix leg of decode, on the 5th-generation tensor cores
// (tcgen05 + TMEM), sm_100a.
//
// extend_tc_kernel<MODE, F16, D, HALF>:
//   MODE kModeExtend  the Blackwell-native path of b200_attn_extend (BASELINE.json config 4: Llama-3.1-8B, 8 x 2048
//                     new tokens on an 8192-token cached prefix).  Same semantics as extend.cu / the reference's
//                     extend_attention_fwd (scratchpad/nn/attention/triton_attn/extend_attention.py:53-226): query
//                     row t of request i sees the gathered prefix slots req_to_token[r, :pre_i] (non-causal) and
//                     k_new/v_new[start_i : start_i+t+1].
//   MODE kModeShared  the decode queries of a whole batch against the prefix all its requests share.
//   F16               the pool is fp8-e4m3: rows are widened to f16 on chip and the MMAs run f16 x f16.
//   D                 head_dim 128 or 64.
//   HALF              the model dtype (q, new tokens, o, a 16-bit pool) is fp16 instead of bf16: the MMAs run f16 x f16
//                     on the bytes as they are; over an fp8 pool q needs no narrowing and the new tokens are scaled in f16.
//
// One CTA per SM, 16 warps in 4 warpgroups, warp-specialised:
//   warps 0-3 / 4-7 : softmax warpgroups for Q tile 0 / 1 (thread = one query row = one TMEM lane);
//                     a thread holds its whole 128-column S row in registers (one TMEM read per tile)
//   warp  8         : MMA issuer (the warp runs the loop uniformly, lane 0 issues tcgen05.mma / tcgen05.commit)
//   warps 9-11      : idle (they only pad warpgroup 2 so that setmaxnreg can move its registers)
//   warps 12-15     : producers -- gather K/V token rows through the page table with 16-byte cp.async
//                     (LDGSTS) straight into the UMMA SWIZZLE_128B canonical layout.  Page size is 1, so
//                     the gather unit is one (token, kv-head) row; a thread always works on rows with the
//                     same (row & 7), so its swizzled destination is a per-thread constant + a fixed step
//                     per row and a copy costs ~3 instructions.
//                     (The TMA row gather, cp.async.bulk.tensor.2d.tile::gather4 + SWIZZLE_128B, was built
//                     and measured first: correct -- tools/gather4_probe.cu -- but the TMA unit retires one
//                     128-byte row segment per ~18 clk, 4500 clk per 32 KiB tile, 2x slower than the whole
//                     K/V budget of a tile; see DESIGN.md.)
// Registers: 512 threads x 128 at launch; setmaxnreg gives the softmax warpgroups 192 and leaves 64 to the
// MMA/producer warpgroups.
// A CTA owns 256 query rows = 2 tiles x (128/g tokens x g q-heads of one kv head): GQA packing makes
// every staged K/V tile serve the whole group and halves K/V shared-memory traffic per flop.
// TMEM (512 columns): S0 | S1 | O0 | O1, 128 fp32 columns each; P (bf16/f16) overwrites S in place and is
// the A operand of the P.V MMA straight from TMEM, so P never touches shared memory or registers of
// another thread.  Ping-pong schedule on the tensor pipe:
//   S0(j) S1(j) | PV0(j) S0(j+1) | PV1(j) S1(j+1) | ...
// so the pipe works on tile 1 while warpgroup 0 does softmax and vice versa; P is released in two 64-key
// halves so the first P.V k-steps overlap the rest of the exponentials.  The O rescale is lazy (only when a
// row maximum grows by more than 2^8), done by the softmax threads themselves.
// K/V ring: 5 x 32 KiB slots (K_j, V_j alternate), mbarrier full (producers) / empty (tcgen05.commit).

#include "common.cuh"
#include "tc_ptx.cuh"
#include "tensor_map.cuh"

namespace b200 {
namespace tc {

constexpr int kD = 128;                    // widest head_dim; ring slots and Q regions are sized for it (head_dim 64 uses half)
constexpr int kBN = 128;                   // keys per tile
constexpr int kTileRows = 128;             // rows per Q tile (UMMA M)
constexpr int kRows = 2 * kTileRows;       // query rows per CTA
constexpr int kSlots = 5;
constexpr int kTileBytes = kBN * kD * 2;   // 32 KiB: one K or V tile, or one Q tile
constexpr int kColBlockBytes = 128 * 128;  // [128 rows x 64 elements] SWIZZLE_128B block
constexpr int kSoftmaxWarps = 8;
constexpr int kMmaWarp = 8;
constexpr int kProdWarp0 = 12;
constexpr int kProdThreads = 128;
constexpr int kThreads = 512;              // 4 warpgroups: softmax 0, softmax 1, MMA (+3 idle warps), producers
constexpr int kLag = 2;                    // cp.async groups a producer thread keeps in flight
#ifndef B200_TC_POLY
#define B200_TC_POLY 0
#endif
#ifndef B200_TC_POLY1
#define B200_TC_POLY1 B200_TC_POLY
#endif
// of every 4 score pairs, this many take the FMA-pipe exp2.  Measured on C4 (gpurun_out/r1s,r1t,r1u): 0 -> 1049,
// 1 -> 1021, 2 -> 940 TFLOP/s: FFMA2/FADD2 retire 2 clk per warp instruction, so the polynomial costs the FMA
// pipe what the MUFU saves; kept as a build option (B200_TC_POLY) for re-tuning.
constexpr int kPolyTile0 = B200_TC_POLY, kPolyTile1 = B200_TC_POLY1;  // per Q tile
template <int V>
struct IntTag {
  static constexpr int value = V;
};
constexpr int kRegsSoftmax = 192, kRegsOther = 64;  // 256*192 + 256*64 = 64 Ki registers

constexpr int kSmemQ = 0;
constexpr int kSmemRing = 2 * kTileBytes;
constexpr int kSmemBar = kSmemRing + kSlots * kTileBytes;  // 229376
constexpr int kNumBars = 2 * kSlots + 1 + 2 + 4 + 2 + 1;
constexpr int kSmemTmemPtr = kSmemBar + kNumBars * 8;
constexpr int kSmemBytes = kSmemTmemPtr + 16 + 1024;       // + alignment slack

constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColS0 = 0, kColS1 = 128, kColO0 = 256, kColO1 = 384;
constexpr float kRescaleThreshold = 8.0f;  // log2 domain

// ---- optional timeline trace of CTA (0,0,0) (build with -DB200_TC_TRACE; tools/tc_trace.py) ----------------
#ifdef B200_TC_TRACE
constexpr int kTraceEvents = 20, kTraceLen = 128;
__device__ long long g_trace[kTraceEvents][kTraceLen];
#define TC_TRACE(ev, idx)                                                                                  \
  do {                                                                                                     \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (idx) < kTraceLen) g_trace[ev][idx] = clock64(); \
  } while (0)
#else
#define TC_TRACE(ev, idx) \
  do {                    \
  } while (0)
#endif

struct Args {
  const uint16_t* q;  // bf16 or fp16 bits (the model dtype), like k_new / v_new / o
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
  int hq, hkv, g, bm_t;  // bm_t = tokens (extend) or requests (shared-prefix decode) per CTA = 256 / g
  float sm_scale, logit_cap;
  // shared-prefix decode (kModeShared) only
  CasTables cas;           // prefix-group row tiles (plan)
  float* cas_o;            // [nsplit, bs, hq, 128] fp32 partial outputs, normalised
  float* cas_lse;          // [nsplit, bs, hq] log2-domain log-sum-exp
  const float* k_scale;    // per kv head (fp8 pools), or null
  const float* v_scale;
  int bs;
  int tma;  // extend, bf16 pools: bit 0 = tiles of 128 new tokens, bit 1 = prefix tiles whose 128 slots are consecutive: TMA
};

__device__ unsigned long long g_tc_tma_tiles;  // debug counter: K / V tiles the extend kernel fetched by TMA since the last reset

// tensor maps of the launch (extend mode; zero-filled and unused otherwise): pool K / V, new-token K / V
struct TcMaps {
  CUtensorMap pool_k, pool_v, new_k, new_v;
};

constexpr int kModeExtend = 0;  // causal extend over [gathered prefix | new tokens], bf16 out
constexpr int kModeShared = 1;  // decode queries of the whole batch against the prefix they all share, fp32 partials out

// byte offset of 16-byte chunk (row r, element block cb of 64, chunk c of 8) inside a [128 x 128] bf16 tile
__device__ __forceinline__ uint32_t tile_off(int r, int cb, int c) {
  return cb * kColBlockBytes + r * 128 + ((c ^ (r & 7)) << 4);
}

// MODE kModeExtend: grid (query block, kv head, request).  MODE kModeShared: grid (block of 256/g requests, kv head,
// KV split); F16 = the pool is fp8-e4m3: K/V are widened to f16 on their way into shared memory (exact), Q is
// narrowed bf16 -> f16 (exact for |q| >= 2^-14) and the MMAs run as f16 x f16; k_scale folds into the softmax
// scale and v_scale into the epilogue, like the split-KV kernel does.
template <int MODE, bool F16, int D, bool HALF>
__global__ void __launch_bounds__(kThreads, 1) extend_tc_kernel(const __grid_constant__ Args a, const __grid_constant__ TcMaps tm) {
  static_assert(D == 128 || D == 64, "head_dim 128 or 64");
  constexpr bool MATH16 = F16 || HALF;  // MMA operand type f16 (else bf16)
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) TC_TRACE(16, 0);
  pdl_wait();
  pdl_launch_dependents();
  if (tid == 0) TC_TRACE(16, 1);

  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024-byte alignment
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar0 = sbase + kSmemBar;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (kSlots + s); };
  const uint32_t bar_q = bar0 + 8u * (2 * kSlots);
  auto bar_s = [&](int i) { return bar0 + 8u * (2 * kSlots + 1 + i); };
  auto bar_p = [&](int i, int half) { return bar0 + 8u * (2 * kSlots + 3 + 2 * i + half); };  // P_i keys [64 half, +64) written
  auto bar_pv = [&](int i) { return bar0 + 8u * (2 * kSlots + 7 + i); };
  const uint32_t bar_fin = bar0 + 8u * (2 * kSlots + 9);  // every MMA and every commit of the tile has landed
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(sgen + kSmemTmemPtr);

  // extend: grid (query block, kv head, request), one tile per CTA.  Shared-prefix decode: grid (kv head, KV split,
  // row-tile lane): the grid has ceil(bs / requests-per-tile) lanes -- exactly the tiles of one group that covers the
  // batch -- and a CTA loops `tile += gridDim.z` over whatever further tiles the plan produced for more groups.
  // Every role runs the same tile loop inside its own branch (so that ptxas sees one register budget per role):
  //   get_tile -> [MMA warp: barrier init, TMEM alloc] -> tile_begin (CTA barrier) -> role body -> tile_end (CTA barrier)
  const int kvh = MODE == kModeExtend ? blockIdx.y : blockIdx.x;
  const int G = a.g, bm_t = a.bm_t;
  struct Tile {
    int pre, start, t0, t_end, kv_len, nt, cas_split, live_rows;
    const int32_t* row;
  };
  // 0: no more tiles for this CTA, 1: run the tile, 2: nothing to compute (skip)
  auto get_tile = [&](int it, Tile& t) -> int {
    t.cas_split = 0;
    if constexpr (MODE == kModeExtend) {
      if (it > 0) return 0;
      const int b = blockIdx.z;
      const int ext = __ldg(a.ext_lens + b);
      const int nqb = (ext + bm_t - 1) / bm_t;
      const int qb = nqb - 1 - static_cast<int>(blockIdx.x);  // heavy (late) blocks first
      if (qb < 0) return 0;
      const int seq = static_cast<int>(load_idx(a.seq_lens, a.idx64, b));
      t.pre = seq - ext;
      t.start = __ldg(a.ext_start + b);
      t.row = a.r2t + load_idx(a.req_idx, a.idx64, b) * a.r2t_stride;
      t.t0 = qb * bm_t;
      t.t_end = min(ext, t.t0 + bm_t);
      t.kv_len = t.pre + t.t_end;
      t.nt = (t.kv_len + kBN - 1) / kBN;
    } else {
      // one row tile of one prefix group: requests perm[t0 .. t_end) share `shared` leading slots (a multiple of kBN)
      const int tile = blockIdx.z + it * gridDim.z;
      if (tile >= __ldg(a.cas.hdr + kHdrCasTiles)) return 0;
      const int shared = __ldg(a.cas.ct_shared + tile);
      t.t0 = __ldg(a.cas.ct_first + tile);
      t.t_end = t.t0 + __ldg(a.cas.ct_cnt + tile);
      t.cas_split = blockIdx.y;
      const int ntot = shared / kBN, ns = gridDim.y;
      const int j0 = static_cast<int>(static_cast<int64_t>(t.cas_split) * ntot / ns);
      t.nt = static_cast<int>(static_cast<int64_t>(t.cas_split + 1) * ntot / ns) - j0;
      const int leader = __ldg(a.cas.ct_leader + tile);
      t.row = a.r2t + static_cast<int64_t>(__ldg(a.cas.req32 + leader)) * a.r2t_stride + j0 * kBN;
      t.start = 0;
      t.pre = t.kv_len = t.nt * kBN;
      if (t.nt <= 0) {  // more splits than tiles (the plan never produces this): an empty partial
        const int R = tid;
        if (R < (t.t_end - t.t0) * G)
          a.cas_lse[(static_cast<int64_t>(t.cas_split) * a.bs + __ldg(a.cas.perm + t.t0 + R / G)) * a.hq + kvh * G + R % G] = -INFINITY;
        return 2;
      }
    }
    t.live_rows = (t.t_end - t.t0) * G;
    return 1;
  };
  auto tile_begin = [&]() -> uint32_t {  // barriers initialised, TMEM allocated: visible to the whole CTA
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return *tmem_ptr;
  };
  auto tile_end = [&]() {  // every role is done with shared memory, TMEM and the barriers of this tile
    tc_fence_before();
    __syncthreads();
  };
#define TC_TILE_LOOP_BEGIN                                                                                   \
  for (int it = 0;; ++it) {                                                                                  \
    Tile tl;                                                                                                 \
    const int tile_state = get_tile(it, tl);                                                                 \
    if (tile_state == 0) break;                                                                              \
    if (tile_state == 2) continue;                                                                           \
    const int pre = tl.pre, start = tl.start, t0 = tl.t0, t_end = tl.t_end, kv_len = tl.kv_len, nt = tl.nt,   \
              cas_split = tl.cas_split, live_rows = tl.live_rows;                                            \
    const int32_t* row = tl.row;                                                                             \
    (void)pre; (void)start; (void)t0; (void)t_end; (void)kv_len; (void)nt; (void)cas_split; (void)live_rows; (void)row;
#define TC_TILE_LOOP_END \
    tile_end();          \
  }

  if (warp >= kProdWarp0) {
    // =========================== producers ===========================================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsOther));
    TC_TILE_LOOP_BEGIN
    tile_begin();
    if constexpr (F16) {
      // fp8-e4m3 pool.  A 128-key tile is 16 KiB of e4m3 and 32 KiB of f16:
      //  phase A  the raw pool rows are gathered with LDGSTS into the UPPER half of the tile's ring slot (row r
      //           at 16 KiB + 128 r), kLag tiles ahead, so the page-table and DRAM latencies overlap;
      //  phase B  each thread widens the chunks it copied itself (cvt.rn.f16x2.e4m3x2, exact) and stores them
      //           in the SWIZZLE_128B layout.  The f16 row r of element block 1 lands exactly on raw row r, which
      //           the 8 lanes that share the row (same warp) have all read before any of them stores.
      //  New tokens (extend mode; bf16, never read from the pool) are copied straight to their final place in
      //  phase A and converted in place in phase B: bf16 -> x 1/scale -> f16, so that the per-head scale folds
      //  into the softmax scale (K) and the epilogue (V) for pool rows and new rows alike.
      // A raw row is D bytes = CPRR 16-byte chunks, so a warp instruction covers 32 / CPRR rows and a thread owns raw
      // chunk c8 (= f16 chunks 2c, 2c+1) of rows rr + ROWSTEP q, q < NQ (ROWSTEP is a multiple of 8: constant row & 7
      // per thread).  head_dim 128: 8 chunks, rows 16q + 4pw + lane/8, the f16 tile fills the slot and the raw rows
      // alias its second element block (see phase B); head_dim 64: 4 chunks, rows 32q + 8pw + lane/4, the f16 tile is
      // the slot's lower 16 KiB and the raw rows ([128 x 64 B]) sit above it, no aliasing.
      constexpr int CPRR = D / 16, ROWSTEP = 4 * (32 / CPRR), NQ = kBN / ROWSTEP;
      constexpr uint32_t kRawStep = ROWSTEP * D, kOutStep = ROWSTEP * 128;  // bytes between a thread's rows: raw / f16
      const int pw = warp - kProdWarp0, c8 = lane % CPRR, rr = (32 / CPRR) * pw + lane / CPRR, rr7 = rr & 7;
      const int64_t col8 = static_cast<int64_t>(kvh) * D + c8 * 16;
      const uint32_t ring = sbase + kSmemRing;
      const uint32_t raw_thread = kColBlockBytes + rr * D + c8 * 16;
      const uint32_t out_thread = (c8 >> 2) * kColBlockBytes + rr * 128;
      const uint32_t o0 = (((2 * c8) & 7) ^ rr7) << 4, o1 = (((2 * c8 + 1) & 7) ^ rr7) << 4;
      const float inv_ks = a.k_scale ? 1.f / __ldg(a.k_scale + kvh) : 1.f;
      const float inv_vs = a.v_scale ? 1.f / __ldg(a.v_scale + kvh) : 1.f;
      auto widen = [&](int n) {
        const uint32_t slot_base = ring + (n % kSlots) * kTileBytes;
        const int p_base = (n >> 1) * kBN + rr;
        const float inv = (n & 1) ? inv_vs : inv_ks;
#pragma unroll
        for (int q0 = 0; q0 < NQ; q0 += 4) {  // 4 rows at a time: this warpgroup runs on 64 registers
          uint32_t raw[4][4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (MODE == kModeShared || p_base + ROWSTEP * (q0 + q) < pre) ld_shared_v4(slot_base + raw_thread + (q0 + q) * kRawStep, raw[q]);
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t dst = slot_base + out_thread + (q0 + q) * kOutStep;
            if (MODE == kModeShared || p_base + ROWSTEP * (q0 + q) < pre) {
              uint32_t h[8];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                h[2 * e] = e4m3x2_to_f16x2(static_cast<uint16_t>(raw[q][e] & 0xffffu));
                h[2 * e + 1] = e4m3x2_to_f16x2(static_cast<uint16_t>(raw[q][e] >> 16));
              }
              st_shared_v4(dst + o0, h[0], h[1], h[2], h[3]);
              st_shared_v4(dst + o1, h[4], h[5], h[6], h[7]);
            } else {
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                uint32_t w[4];
                ld_shared_v4(dst + (half ? o1 : o0), w);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if constexpr (HALF) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
                    w[e] = pack_f16x2(f.x * inv, f.y * inv);
                  } else {
                    w[e] = pack_f16x2(__uint_as_float(w[e] << 16) * inv, __uint_as_float(w[e] & 0xffff0000u) * inv);
                  }
                }
                st_shared_v4(dst + (half ? o1 : o0), w[0], w[1], w[2], w[3]);
              }
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(bar_full(n % kSlots));
      };
      const int total = 2 * nt;
      for (int n = 0; n < total; ++n) {
        const int j = n >> 1, is_v = n & 1, s = n % kSlots;
        const uint8_t* pool8 = (is_v ? a.vbuf : a.kbuf) + col8;
        const int64_t pool_st = is_v ? a.v_st_bytes : a.k_st_bytes;
        mbar_wait(bar_empty(s), ((n / kSlots) & 1) ^ 1);
        const uint32_t slot_base = ring + s * kTileBytes;
        if (MODE == kModeShared || (j + 1) * kBN <= pre) {
          int slot[NQ];
#pragma unroll
          for (int q = 0; q < NQ; ++q) slot[q] = __ldg(row + j * kBN + ROWSTEP * q + rr);
#pragma unroll
          for (int q = 0; q < NQ; ++q) cp_async16(slot_base + raw_thread + q * kRawStep, kv_row_ptr(pool8, slot[q], pool_st), 16);
        } else {
          const uint8_t* fresh = reinterpret_cast<const uint8_t*>(is_v ? a.v_new : a.k_new) + static_cast<int64_t>(kvh) * (D * 2) + c8 * 32;
          const int64_t f_st = (is_v ? a.vn_st : a.kn_st) * 2;
#pragma unroll 2
          for (int q = 0; q < NQ; ++q) {
            const int p = j * kBN + ROWSTEP * q + rr, pc = min(p, kv_len - 1);
            if (pc < pre) {
              cp_async16(slot_base + raw_thread + q * kRawStep, kv_row_ptr(pool8, __ldg(row + pc), pool_st), 16);
            } else {
              const uint8_t* src = fresh + static_cast<int64_t>(start + pc - pre) * f_st;
              const int nb = p < kv_len ? 16 : 0;  // keys past the end: zeros (masked by the softmax anyway)
              cp_async16(slot_base + out_thread + q * kOutStep + o0, src, nb);
              cp_async16(slot_base + out_thread + q * kOutStep + o1, src + 16, nb);
            }
          }
        }
        cp_async_commit();
        if (n >= kLag) {
          cp_async_wait<kLag>();
          widen(n - kLag);
        }
      }
      cp_async_wait<0>();
      for (int n = max(total - kLag, 0); n < total; ++n) widen(n);
    } else {
    // A (token, kv-head) row is D*2 bytes = CPR 16-byte chunks, so a warp instruction covers RPI = 32/CPR rows
    // (head_dim 128: 2 rows, head_dim 64: 4 rows) and a thread copies chunk `ch` of rows STEP*q + RPI*pw + lane/CPR,
    // q = 0..NQ-1 (STEP = 4*RPI is a multiple of 8, so row & 7 is a per-thread constant).
    constexpr int CPR = D / 8, RPI = 32 / CPR, STEP = 4 * RPI, NQ = kBN / STEP;
    const int pw = warp - kProdWarp0, ch = lane % CPR;
    const int rr = RPI * pw + lane / CPR, r7 = rr & 7;
    const uint32_t dst_thread = (ch >> 3) * kColBlockBytes + rr * 128 + (((ch & 7) ^ r7) << 4);
    const int64_t col_bytes = static_cast<int64_t>(kvh) * (D * 2) + ch * 16;
    const int total = 2 * nt;
    // TMA where 128 keys are 128 consecutive rows (north_star: "TMA where pages are contiguous"): the new tokens of the
    // chunk always are (rows of k_new / v_new), cached-prefix tiles are when the allocator handed the slots out as a
    // run -- pages of >= 128 tokens, or an unfragmented token-granular pool -- which every producer warp checks for
    // itself on the tile's slot ids (4 per lane, one vote: all warps reach the same verdict without talking).  Such a
    // tile is ONE box copy for K and one for V, issued by one thread and completing on the slot's full barrier through
    // its tx count; the 128 lagged thread arrivals stay as they are.  Everything else -- scattered slots, the tile
    // that straddles prefix and new tokens, the ragged last tile -- takes the row gather below.
    bool tile_by_tma = false;
    int tile_row0 = 0, n_tma = 0;
    // Page-table lookups run one tile ahead of the copies they feed (-DB200_TC_NO_SLOT_PREFETCH: looked up in place)
#ifndef B200_TC_NO_SLOT_PREFETCH
    int slot_nx[NQ];
    auto fetch_slots = [&](int n_) {
      const int j_ = n_ >> 1;
      if (n_ < total && (j_ + 1) * kBN <= pre) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) slot_nx[q] = __ldg(row + j_ * kBN + rr + STEP * q);
      }
    };
    fetch_slots(0);
#endif
    for (int n = 0; n < total; ++n) {
      const int j = n >> 1, is_v = n & 1, s = n % kSlots;
      const uint8_t* pool = (is_v ? a.vbuf : a.kbuf) + col_bytes;
      const int64_t pool_st = is_v ? a.v_st_bytes : a.k_st_bytes;
      const int p0 = j * kBN + rr;
      mbar_wait(bar_empty(s), ((n / kSlots) & 1) ^ 1);
      if (tid == kProdWarp0 * 32) TC_TRACE(0, n);
      const uint32_t dst0 = sbase + kSmemRing + s * kTileBytes + dst_thread;
#ifndef B200_TC_NO_SLOT_PREFETCH
      int slot[NQ];  // this tile's slot ids (looked up one tile ago; meaningful for prefix tiles), then the next tile's
#pragma unroll
      for (int q = 0; q < NQ; ++q) slot[q] = slot_nx[q];
      fetch_slots(n + 1);
#endif
      if constexpr (MODE == kModeExtend) {
        if (a.tma != 0 && !is_v) {  // decided on the K tile, reused for the V tile of the same keys
          tile_by_tma = false;
          if ((a.tma & 1) && j * kBN >= pre && (j + 1) * kBN <= kv_len) {
            tile_by_tma = true;
            tile_row0 = start + j * kBN - pre;
          } else if ((a.tma & 2) && (j + 1) * kBN <= pre) {
            const int32_t* rp = row + j * kBN + 4 * lane;
            const int s0 = __ldg(rp), s1 = __ldg(rp + 1), s2 = __ldg(rp + 2), s3 = __ldg(rp + 3);
            tile_row0 = __shfl_sync(0xffffffffu, s0, 0);
            tile_by_tma = __all_sync(0xffffffffu, s1 == s0 + 1 && s2 == s0 + 2 && s3 == s0 + 3 && s0 == tile_row0 + 4 * lane);
          }
        }
      }
      if (MODE == kModeExtend && tile_by_tma) {
        if (tid == kProdWarp0 * 32) {
          const bool fresh_tile = j * kBN >= pre;
          const CUtensorMap* map = fresh_tile ? (is_v ? &tm.new_v : &tm.new_k) : (is_v ? &tm.pool_v : &tm.pool_k);
          mbar_expect_tx(bar_full(s), kBN * D * 2);
          tma_load_4d(sbase + kSmemRing + s * kTileBytes, map, bar_full(s), 0, tile_row0, 0, kvh);
          ++n_tma;
        }
      } else if ((j + 1) * kBN <= pre) {
        // tile of cached-prefix keys only: page-table lookup (L1-resident after the first touch) + one copy
#ifdef B200_TC_NO_SLOT_PREFETCH
        int slot[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) slot[q] = __ldg(row + p0 + STEP * q);
#endif
#pragma unroll
        for (int q = 0; q < NQ; ++q) cp_async16(dst0 + q * (STEP * 128), kv_row_ptr(pool, slot[q], pool_st), 16);
      } else {
        const uint8_t* fresh = reinterpret_cast<const uint8_t*>(is_v ? a.v_new : a.k_new) + col_bytes;
        const int64_t f_st = (is_v ? a.vn_st : a.kn_st) * 2;
#pragma unroll 4
        for (int q = 0; q < NQ; ++q) {
          const int p = p0 + STEP * q, pc = min(p, kv_len - 1);
          const uint8_t* src = pc < pre ? kv_row_ptr(pool, __ldg(row + pc), pool_st) : fresh + static_cast<int64_t>(start + pc - pre) * f_st;
          cp_async16(dst0 + q * (STEP * 128), src, p < kv_len ? 16 : 0);
        }
      }
      cp_async_commit();
      if (tid == kProdWarp0 * 32) TC_TRACE(1, n);
      if (n >= kLag) {
        cp_async_wait<kLag>();
        fence_proxy_async();
        mbar_arrive(bar_full((n - kLag) % kSlots));
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (int n = max(total - kLag, 0); n < total; ++n) mbar_arrive(bar_full(n % kSlots));
    if (n_tma) atomicAdd(&g_tc_tma_tiles, static_cast<unsigned long long>(n_tma));
    }  // !F16
    TC_TILE_LOOP_END
  } else if (warp == kMmaWarp) {
    // =========================== MMA issuer ============================================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsOther));
    bool have_tmem = false;
    uint32_t tmem = 0;
    TC_TILE_LOOP_BEGIN
    if (lane == 0) {  // (re-)initialise: every barrier starts each tile at phase 0
      for (int s = 0; s < kSlots; ++s) {
        mbar_init(bar_full(s), kProdThreads);
        mbar_init(bar_empty(s), 1);
      }
      mbar_init(bar_q, kSoftmaxWarps * 32);
      for (int i = 0; i < 2; ++i) {
        mbar_init(bar_s(i), 1);
        mbar_init(bar_p(i, 0), 128);
        mbar_init(bar_p(i, 1), 128);
        mbar_init(bar_pv(i), 1);
      }
      mbar_init(bar_fin, 1);
      fence_barrier_init();
    }
    __syncwarp();
    if (!have_tmem) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + kSmemTmemPtr), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
      have_tmem = true;
    }
    tmem = tile_begin();
    // The whole warp runs this loop so that every address / descriptor is warp-uniform (uniform registers,
    // no R2UR in front of each UTCHMMA); only the tcgen05.mma / tcgen05.commit themselves are issued by lane 0.
    const bool leader = lane == 0;
    {
      constexpr uint32_t idesc_qk = instr_desc(kTileRows, kBN, false, MATH16);
      constexpr uint32_t idesc_pv = instr_desc(kTileRows, D, true, MATH16);
      const uint32_t q_base = sbase + kSmemQ;
      const uint32_t ring = sbase + kSmemRing;
      auto issue_qk = [&](int i, uint32_t k_base) {  // S_i = Q_i . K^T
        const uint32_t d = tmem + (i ? kColS1 : kColS0);
        const uint32_t qlo = desc_lo(q_base + i * kTileBytes, 16), klo = desc_lo(k_base, 16);
#pragma unroll
        for (int kc = 0; kc < D / 16; ++kc) {
          const uint32_t off16 = ((kc >> 2) * kColBlockBytes + (kc & 3) * 32) >> 4;
          if (leader) tc_mma_ss(d, desc_from(qlo + off16, 1024), desc_from(klo + off16, 1024), idesc_qk, kc > 0);
        }
        if (leader) tc_commit(bar_s(i));
      };
      // O_i (+)= P_i . V over keys [64 half, +64): P is released in two halves, so the first four k-steps run
      // while the softmax threads are still exponentiating the second half of the row
      auto issue_pv = [&](int i, uint32_t v_base, bool first, int half) {
        const uint32_t d = tmem + (i ? kColO1 : kColO0);
        const uint32_t p = tmem + (i ? kColS1 : kColS0);
        const uint32_t vlo = desc_lo(v_base, kColBlockBytes);
#pragma unroll
        for (int k4 = 0; k4 < kBN / 32; ++k4) {
          const int ks = half * (kBN / 32) + k4;
          if (leader) tc_mma_ts(d, p + ks * 8, desc_from(vlo + ((ks * 2048) >> 4), 1024), idesc_pv, !(first && ks == 0));
        }
        if (half && leader) tc_commit(bar_pv(i));
      };
      auto slot_of = [&](int n) { return n % kSlots; };
      auto wait_full = [&](int n) {
        mbar_wait(bar_full(slot_of(n)), (n / kSlots) & 1);
        tc_fence_after();
      };
      mbar_wait(bar_q, 0);
      tc_fence_after();
      wait_full(0);
      issue_qk(0, ring + slot_of(0) * kTileBytes);
      issue_qk(1, ring + slot_of(0) * kTileBytes);
      if (leader) tc_commit(bar_empty(slot_of(0)));
      for (int j = 0; j < nt; ++j) {
        const int nv = 2 * j + 1, nk = 2 * j + 2;
        const bool more = j + 1 < nt;
        wait_full(nv);
        if (leader) TC_TRACE(2, j);
        const uint32_t v_base = ring + slot_of(nv) * kTileBytes;
        mbar_wait(bar_p(0, 0), j & 1);
        tc_fence_after();
        if (leader) TC_TRACE(3, j);
        issue_pv(0, v_base, j == 0, 0);
        mbar_wait(bar_p(0, 1), j & 1);
        tc_fence_after();
        issue_pv(0, v_base, j == 0, 1);
        if (more) {
          wait_full(nk);
          if (leader) TC_TRACE(4, j);
          issue_qk(0, ring + slot_of(nk) * kTileBytes);
        }
        mbar_wait(bar_p(1, 0), j & 1);
        tc_fence_after();
        if (leader) TC_TRACE(5, j);
        issue_pv(1, v_base, j == 0, 0);
        mbar_wait(bar_p(1, 1), j & 1);
        tc_fence_after();
        issue_pv(1, v_base, j == 0, 1);
        if (leader) tc_commit(bar_empty(slot_of(nv)));
        if (more) {
          issue_qk(1, ring + slot_of(nk) * kTileBytes);
          if (leader) tc_commit(bar_empty(slot_of(nk)));
        }
      }
      // all asynchronous arrivals of this tile (the commits above) are in before the barriers are re-initialised
      if (leader) tc_commit(bar_fin);
      mbar_wait(bar_fin, 0);
      tc_fence_after();
    }
    __syncwarp();
    TC_TILE_LOOP_END
    if (have_tmem) {
      tc_fence_after();
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
    }
  } else if (warp > kMmaWarp) {
    // idle padding warps of warpgroup 2: they only take part in the CTA barriers
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsOther));
    TC_TILE_LOOP_BEGIN
    tile_begin();
    TC_TILE_LOOP_END
  } else {
    // =========================== softmax warpgroups =======================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsSoftmax));
    TC_TILE_LOOP_BEGIN
    const uint32_t tmem = tile_begin();
    if (tid == 0) TC_TRACE(17, it);
    // Q tiles first (the 256 softmax threads are idle until the first S lands):
    // row R = i*128 + r -> (token t0 + R/G, head kvh*G + R%G)
    // Q tiles: chunk `ch` (constant per thread) of rows R = r_base + RSTEP * it_q.  (token, head) of a row advance
    // incrementally -- one division per thread instead of two per chunk: with 1300 dependent integer instructions
    // in front of the first MMA this loop was 10 K clocks of every CTA's prologue.
    constexpr int CPRQ = D / 8;                                  // 16-byte chunks per Q row
    constexpr int RSTEP = kSoftmaxWarps * 32 / CPRQ;              // rows between two chunks of a thread
    constexpr int kQIters = kRows / RSTEP;
    const int ch = tid % CPRQ, r_base = tid / CPRQ;
    const uint32_t q_dst0 = sbase + kSmemQ + (ch >> 3) * kColBlockBytes + (((ch & 7) ^ (r_base & 7)) << 4);
    static_assert(RSTEP % 8 == 0, "row & 7 must be a per-thread constant");
    int q_rowidx[kQIters];  // extend: token start + t of the flattened batch; shared-prefix decode: request perm[t]
    int q_head[kQIters];
    {
      int t = r_base / G, h = r_base % G;
      const int dt = RSTEP / G, dh = RSTEP % G;
#pragma unroll
      for (int it_q = 0; it_q < kQIters; ++it_q) {  // (all perm loads first: they are a dependent hop of their own)
        const bool live = r_base + RSTEP * it_q < live_rows;
        q_head[it_q] = live ? h : 0;
        const int tt = t0 + (live ? t : 0);
        q_rowidx[it_q] = MODE == kModeExtend ? start + tt : __ldg(a.cas.perm + tt);
        t += dt;
        h += dh;
        if (h >= G) {
          h -= G;
          ++t;
        }
      }
    }
    auto q_chunk = [&](int it_q, const uint16_t*& src, uint32_t& dst, bool& live) {
      const int R = r_base + RSTEP * it_q;
      live = R < live_rows;
      src = a.q + static_cast<int64_t>(q_rowidx[it_q]) * a.q_st + static_cast<int64_t>(kvh * G + q_head[it_q]) * a.q_sh + ch * 8;
      dst = q_dst0 + ((RSTEP * it_q) >> 7) * kTileBytes + (r_base + ((RSTEP * it_q) & 127)) * 128;
    };
    if constexpr (F16 && !HALF) {
      // bf16 -> f16 on the way in: all of the thread's loads are issued before the first one is converted
      // (one memory round trip; one load per iteration was 10 us of the shared-prefix leg's 40)
      uint4 x[kQIters];
      uint32_t dsts[kQIters];
#pragma unroll
      for (int it_q = 0; it_q < kQIters; ++it_q) {
        const uint16_t* src;
        bool live;
        q_chunk(it_q, src, dsts[it_q], live);
        x[it_q] = live ? __ldg(reinterpret_cast<const uint4*>(src)) : make_uint4(0u, 0u, 0u, 0u);
      }
      if (tid == 0) TC_TRACE(17, 1);
#pragma unroll
      for (int it_q = 0; it_q < kQIters; ++it_q) {
        if (tid == 0 && it_q == 1) TC_TRACE(17, 2);
        const uint32_t w[4] = {x[it_q].x, x[it_q].y, x[it_q].z, x[it_q].w};
        uint32_t h2[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) h2[e] = pack_f16x2(__uint_as_float(w[e] << 16), __uint_as_float(w[e] & 0xffff0000u));
        st_shared_v4(dsts[it_q], h2[0], h2[1], h2[2], h2[3]);
      }
    } else {
#pragma unroll
      for (int it_q = 0; it_q < kQIters; ++it_q) {
        const uint16_t* src;
        uint32_t dst;
        bool live;
        q_chunk(it_q, src, dst, live);
        cp_async16(dst, src, live ? 16 : 0);
      }
    }
    if (tid == 0) TC_TRACE(17, 3);
    cp_async_commit();
    cp_async_wait<0>();
    fence_proxy_async();
    mbar_arrive(bar_q);
    if (tid == 0) TC_TRACE(18, it);

    const int i = warp >> 2;                       // Q tile
    const int r = (warp & 3) * 32 + lane;          // row inside the tile == TMEM lane
    const int R = i * kTileRows + r;
    const uint32_t lane_base = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t s_addr = lane_base + (i ? kColS1 : kColS0);
    const uint32_t o_addr = lane_base + (i ? kColO1 : kColO0);
    // causal limit (dead rows behave like the last live token); the shared prefix is visible to every row
    const int lim = MODE == kModeExtend ? pre + min(t0 + R / G, t_end - 1) : kv_len;
    const float cap = a.logit_cap;
    const float sm_scale = (F16 && a.k_scale) ? a.sm_scale * __ldg(a.k_scale + kvh) : a.sm_scale;
    const float sc_log2 = sm_scale * kLog2e;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < nt; ++j) {
      mbar_wait(bar_s(i), j & 1);
      tc_fence_after();
      if (r == 0) TC_TRACE(6 + 3 * i, j);
      const int climit = lim - j * kBN;  // columns <= climit are visible
      const bool masked = climit < kBN - 1;
      // interior tile for the whole warp: no mask, no soft cap.  Must be warp-uniform: both branches
      // below issue .sync.aligned tcgen05.ld/st, which a divergent warp would deadlock on.
      const bool fast = !__any_sync(0xffffffffu, masked) && !(cap > 0.f);
      uint32_t v[kBN];  // this row of S, raw scores
      tc_ld32(s_addr, v);
      tc_ld32(s_addr + 32, v + 32);
      tc_ld32(s_addr + 64, v + 64);
      tc_ld32(s_addr + 96, v + 96);
      tc_wait_ld();
      if (r == 0) TC_TRACE(7 + 3 * i, j);
      // Optimistic order for interior tiles behind the first: the exponentials of keys 0..63 FIRST, against the running
      // maximum m, with the maximum of the whole row computed in their shadow (FMNMX on the ALU pipe while the MUFU
      // works).  The rescale is lazy anyway -- m may lag the true maximum by 2^8 -- so in all but a few tiles of a row the
      // check finds nothing to do and the separate max pass is gone.  If a row's maximum did jump, O is rescaled and the
      // half recomputed from the scores still in registers (an overflowed P is simply overwritten) before it is
      // released.  Measured (same box, gpurun_out/r2l_*): head_dim 64 (exponentials per flop doubled, the MUFU-bound
      // variant) 681 -> 705 TFLOP/s at C4S; head_dim 128 within +-1 % (C4 1158 / 1147, C4F 1047 / 1035, C4P 897 / 906,
      // C3 cascade 66.9 / 66.9 us) -- there the max pass already ran under the other Q tile's exponentials -- so the
      // order is used for head_dim 64 only (-DB200_TC_OPTIMISTIC_ALL / -DB200_TC_NO_OPTIMISTIC for re-measuring).
#if defined(B200_TC_NO_OPTIMISTIC)
      constexpr bool kOptimistic = false;
#elif defined(B200_TC_OPTIMISTIC_ALL)
      constexpr bool kOptimistic = kPolyTile0 == 0 && kPolyTile1 == 0;
#else
      constexpr bool kOptimistic = D == 64 && kPolyTile0 == 0 && kPolyTile1 == 0;
#endif
      const bool optimistic = kOptimistic && fast && j > 0;
      if (optimistic) {
        const float l_before = l;
        float c0 = -INFINITY, c1 = -INFINITY, c2 = -INFINITY, c3 = -INFINITY;
        const uint64_t sc2 = pack_f32x2(sc_log2, sc_log2);
        uint64_t nm2 = pack_f32x2(-m, -m);
        uint64_t l2 = pack_f32x2(l, 0.f);
#pragma unroll
        for (int c4 = 0; c4 < 2; ++c4) {
          uint32_t pk[16];
#pragma unroll
          for (int k = 0; k < 32; k += 2) {
            const float x0 = __uint_as_float(v[c4 * 32 + k]), x1 = __uint_as_float(v[c4 * 32 + k + 1]);
            c0 = fmaxf(c0, x0);
            c1 = fmaxf(c1, x1);
            c2 = fmaxf(c2, __uint_as_float(v[64 + c4 * 32 + k]));
            c3 = fmaxf(c3, __uint_as_float(v[64 + c4 * 32 + k + 1]));
            const uint64_t y = fma_f32x2(pack_f32x2(x0, x1), sc2, nm2);
            float y0, y1;
            unpack_f32x2(y, y0, y1);
            const float p0 = ex2(y0), p1 = ex2(y1);
            l2 = add_f32x2(l2, pack_f32x2(p0, p1));
            pk[k >> 1] = pack2<MATH16>(p0, p1);
          }
          tc_st16(s_addr + c4 * 16, pk);
        }
        const float mxs = fmaxf(fmaxf(c0, c1), fmaxf(c2, c3)) * sc_log2;  // no soft cap on this path
        const bool need = mxs > m + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          // lazy rescale: O_i *= 2^(m_old - m_new); P.V(j-1) must have landed, P.V(j) waits for our arrive
          mbar_wait(bar_pv(i), (j - 1) & 1);
          tc_fence_after();
          const float mn = fmaxf(m, mxs);
          const float alpha = ex2(m - mn);
          m = mn;
#pragma unroll 1
          for (int c4 = 0; c4 < D / 32; ++c4) {
            uint32_t w[32];
            tc_ld32(o_addr + c4 * 32, w);
            tc_wait_ld();
#pragma unroll
            for (int k = 0; k < 32; ++k) w[k] = __float_as_uint(__uint_as_float(w[k]) * alpha);
            tc_st32(o_addr + c4 * 32, w);
          }
          nm2 = pack_f32x2(-m, -m);
          l2 = pack_f32x2(l_before * alpha, 0.f);
#pragma unroll
          for (int c4 = 0; c4 < 2; ++c4) {  // keys 0..63 again, against the new maximum
            uint32_t pk[16];
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
              const uint64_t y = fma_f32x2(pack_f32x2(__uint_as_float(v[c4 * 32 + k]), __uint_as_float(v[c4 * 32 + k + 1])), sc2, nm2);
              float y0, y1;
              unpack_f32x2(y, y0, y1);
              const float p0 = ex2(y0), p1 = ex2(y1);
              l2 = add_f32x2(l2, pack_f32x2(p0, p1));
              pk[k >> 1] = pack2<MATH16>(p0, p1);
            }
            tc_st16(s_addr + c4 * 16, pk);
          }
        }
        tc_wait_st();  // keys 0..63 of P are in TMEM: release them to the MMA warp
        tc_fence_before();
        mbar_arrive(bar_p(i, 0));
#pragma unroll
        for (int c4 = 2; c4 < 4; ++c4) {
          uint32_t pk[16];
#pragma unroll
          for (int k = 0; k < 32; k += 2) {
            const uint64_t y = fma_f32x2(pack_f32x2(__uint_as_float(v[c4 * 32 + k]), __uint_as_float(v[c4 * 32 + k + 1])), sc2, nm2);
            float y0, y1;
            unpack_f32x2(y, y0, y1);
            const float p0 = ex2(y0), p1 = ex2(y1);
            l2 = add_f32x2(l2, pack_f32x2(p0, p1));
            pk[k >> 1] = pack2<MATH16>(p0, p1);
          }
          tc_st16(s_addr + c4 * 16, pk);
        }
        float la, lb;
        unpack_f32x2(l2, la, lb);
        l = la + lb;
      } else {
      float mx;
      if (masked) {
        mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < kBN; ++k)
          if (k <= climit) mx = fmaxf(mx, __uint_as_float(v[k]));
      } else {
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
        for (int k = 0; k < kBN; k += 8) {
          m0 = fmaxf(m0, fmaxf(__uint_as_float(v[k]), __uint_as_float(v[k + 1])));
          m1 = fmaxf(m1, fmaxf(__uint_as_float(v[k + 2]), __uint_as_float(v[k + 3])));
          m2 = fmaxf(m2, fmaxf(__uint_as_float(v[k + 4]), __uint_as_float(v[k + 5])));
          m3 = fmaxf(m3, fmaxf(__uint_as_float(v[k + 6]), __uint_as_float(v[k + 7])));
        }
        mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      }
      if (R == 0) TC_TRACE(12, j);
      // softcap is monotone, so the maximum of the capped scores is the capped maximum
      float mxs = cap > 0.f ? softcap(mx * sm_scale, cap) * kLog2e : mx * sc_log2;
      if (j == 0) {
        m = mxs;  // key 0 is visible to every row, so this is finite
      } else {
        const bool need = mxs > m + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          // lazy rescale: O_i *= 2^(m_old - m_new); P.V(j-1) must have landed, P.V(j) waits for our arrive
          mbar_wait(bar_pv(i), (j - 1) & 1);
          tc_fence_after();
          const float mn = fmaxf(m, mxs);
          const float alpha = ex2(m - mn);
          m = mn;
          l *= alpha;
#pragma unroll 1
          for (int c4 = 0; c4 < D / 32; ++c4) {
            uint32_t w[32];
            tc_ld32(o_addr + c4 * 32, w);
            tc_wait_ld();
#pragma unroll
            for (int k = 0; k < 32; ++k) w[k] = __float_as_uint(__uint_as_float(w[k]) * alpha);
            tc_st32(o_addr + c4 * 32, w);
          }
          tc_wait_st();
        }
      }
      if (R == 0) TC_TRACE(13, j);
      // P = 2^(s - m) -> bf16, written over S in place (P is the TMEM A operand of the P.V MMA)
      if (fast) {
        const uint64_t sc2 = pack_f32x2(sc_log2, sc_log2), nm2 = pack_f32x2(-m, -m);
        uint64_t l2 = pack_f32x2(l, 0.f);
        // NP of every 4 score pairs take the FMA-pipe exp2 (a compile-time constant per Q tile, see kPolyTile0/1)
        auto exp_store = [&](auto np_tag) {
          constexpr int NP = decltype(np_tag)::value;
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            uint32_t pk[16];
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
              const uint64_t y =
                  fma_f32x2(pack_f32x2(__uint_as_float(v[c4 * 32 + k]), __uint_as_float(v[c4 * 32 + k + 1])), sc2, nm2);
              float p0, p1;
              if ((k >> 1) % 4 < NP) {
                exp2_poly_x2(y, p0, p1);
              } else {
                float y0, y1;
                unpack_f32x2(y, y0, y1);
                p0 = ex2(y0);
                p1 = ex2(y1);
              }
              l2 = add_f32x2(l2, pack_f32x2(p0, p1));
              pk[k >> 1] = pack2<MATH16>(p0, p1);
            }
            tc_st16(s_addr + c4 * 16, pk);
            if (c4 == 1) {  // keys 0..63 of P are in TMEM: release them to the MMA warp
              tc_wait_st();
              tc_fence_before();
              mbar_arrive(bar_p(i, 0));
            }
          }
        };
        if (kPolyTile0 == kPolyTile1 || i == 0) exp_store(IntTag<kPolyTile0>{});
        else exp_store(IntTag<kPolyTile1>{});
        float la, lb;
        unpack_f32x2(l2, la, lb);
        l = la + lb;
      } else {
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          uint32_t pk[16];
#pragma unroll
          for (int k = 0; k < 32; k += 2) {
            float x0 = __uint_as_float(v[c4 * 32 + k]), x1 = __uint_as_float(v[c4 * 32 + k + 1]);
            float p0, p1;
            if (cap > 0.f) {
              p0 = ex2(softcap(x0 * sm_scale, cap) * kLog2e - m);
              p1 = ex2(softcap(x1 * sm_scale, cap) * kLog2e - m);
            } else {
              p0 = ex2(fmaf(x0, sc_log2, -m));
              p1 = ex2(fmaf(x1, sc_log2, -m));
            }
            if (c4 * 32 + k > climit) p0 = 0.f;
            if (c4 * 32 + k + 1 > climit) p1 = 0.f;
            l += p0 + p1;
            pk[k >> 1] = pack2<MATH16>(p0, p1);
          }
          tc_st16(s_addr + c4 * 16, pk);
          if (c4 == 1) {
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(bar_p(i, 0));
          }
        }
      }
      }  // !optimistic
      if (R == 0) TC_TRACE(14, j);
      tc_wait_st();
      if (R == 0) TC_TRACE(15, j);
      tc_fence_before();
      mbar_arrive(bar_p(i, 1));
      if (r == 0) TC_TRACE(8 + 3 * i, j);
    }
    // epilogue (thread = row): extend -> O_i / l as bf16; shared prefix -> fp32 partial (O_i v_scale / l, m + log2 l)
    mbar_wait(bar_pv(i), (nt - 1) & 1);
    tc_fence_after();
    const bool live = R < live_rows;
    if constexpr (MODE == kModeExtend) {
      const float inv = l > 0.f ? ((F16 && a.v_scale) ? __ldg(a.v_scale + kvh) : 1.f) / l : 0.f;
      uint16_t* op = a.o + static_cast<int64_t>(start + t0 + (live ? R / G : 0)) * a.o_st +
                          static_cast<int64_t>(kvh * G + (live ? R % G : 0)) * a.o_sh;
#pragma unroll 1
      for (int c4 = 0; c4 < D / 32; ++c4) {
        uint32_t w[32];
        tc_ld32(o_addr + c4 * 32, w);
        tc_wait_ld();
        if (live) {
#pragma unroll
          for (int k = 0; k < 32; k += 8) {
            uint4 u;
            u.x = pack2<HALF>(__uint_as_float(w[k]) * inv, __uint_as_float(w[k + 1]) * inv);
            u.y = pack2<HALF>(__uint_as_float(w[k + 2]) * inv, __uint_as_float(w[k + 3]) * inv);
            u.z = pack2<HALF>(__uint_as_float(w[k + 4]) * inv, __uint_as_float(w[k + 5]) * inv);
            u.w = pack2<HALF>(__uint_as_float(w[k + 6]) * inv, __uint_as_float(w[k + 7]) * inv);
            *reinterpret_cast<uint4*>(op + c4 * 32 + k) = u;
          }
        }
      }
    } else {
      const float vs = (F16 && a.v_scale) ? __ldg(a.v_scale + kvh) : 1.f;
      const float inv = l > 0.f ? vs / l : 0.f;
      const int req = __ldg(a.cas.perm + t0 + (live ? R / G : 0));
      const int64_t pi = (static_cast<int64_t>(cas_split) * a.bs + req) * a.hq + kvh * G + (live ? R % G : 0);
      if (live) a.cas_lse[pi] = l > 0.f ? m + lg2(l) : -INFINITY;
      float* op = a.cas_o + pi * D;
#pragma unroll 1
      for (int c4 = 0; c4 < D / 32; ++c4) {
        uint32_t w[32];
        tc_ld32(o_addr + c4 * 32, w);
        tc_wait_ld();
        if (live) {
#pragma unroll
          for (int k = 0; k < 32; k += 4)
            *reinterpret_cast<float4*>(op + c4 * 32 + k) =
                make_float4(__uint_as_float(w[k]) * inv, __uint_as_float(w[k + 1]) * inv, __uint_as_float(w[k + 2]) * inv,
                            __uint_as_float(w[k + 3]) * inv);
        }
      }
    }
    if (tid == 0) TC_TRACE(19, it);
    TC_TILE_LOOP_END
  }

#undef TC_TILE_LOOP_BEGIN
#undef TC_TILE_LOOP_END
}

}  // namespace tc

// Test hook (synchronises): K / V tiles the tcgen05 extend kernel fetched by TMA on the current device since the last call
// with reset != 0 (-1 on a CUDA error).
extern "C" int64_t b200_debug_extend_tma_tiles(int32_t reset) {
  unsigned long long n = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(&n, tc::g_tc_tma_tiles, sizeof(n)) != cudaSuccess) return -1;
  if (reset) {
    const unsigned long long zero = 0;
    if (cudaMemcpyToSymbol(tc::g_tc_tma_tiles, &zero, sizeof(zero)) != cudaSuccess) return -1;
  }
  return static_cast<int64_t>(n);
}

#ifdef B200_TC_TRACE
extern "C" int b200_debug_tc_trace_read(long long* host_dst) {
  return cudaMemcpyFromSymbol(host_dst, tc::g_trace, sizeof(tc::g_trace)) == cudaSuccess ? 0 : 1;
}
#endif

template <int MODE>
static int dispatch_tc(bool fp8, int d, bool half, const tc::Args& a, const tc::TcMaps& maps, dim3 grid, cudaStream_t st);

template <int MODE, bool F16, int D, bool HALF>
static int launch_tc(const tc::Args& a, const tc::TcMaps& maps, dim3 grid, cudaStream_t st) {
  auto kern = tc::extend_tc_kernel<MODE, F16, D, HALF>;
  static PerDevice pd;  // per instantiation
  if (int rc = ensure_smem_attr(kern, tc::kSmemBytes, pd)) return rc;
  return launch_pdl(kern, grid, dim3(tc::kThreads), tc::kSmemBytes, st, a, maps);
}

template <int MODE>
static int dispatch_tc(bool fp8, int d, bool half, const tc::Args& a, const tc::TcMaps& maps, dim3 grid, cudaStream_t st) {
  if (half) {
    if (d == 64) return fp8 ? launch_tc<MODE, true, 64, true>(a, maps, grid, st) : launch_tc<MODE, false, 64, true>(a, maps, grid, st);
    return fp8 ? launch_tc<MODE, true, 128, true>(a, maps, grid, st) : launch_tc<MODE, false, 128, true>(a, maps, grid, st);
  }
  if (d == 64) return fp8 ? launch_tc<MODE, true, 64, false>(a, maps, grid, st) : launch_tc<MODE, false, 64, false>(a, maps, grid, st);
  return fp8 ? launch_tc<MODE, true, 128, false>(a, maps, grid, st) : launch_tc<MODE, false, 128, false>(a, maps, grid, st);
}

// Called by b200_attn_extend (extend.cu) for head_dim 128 and 64 (bf16 or fp8 pool).  Strides and pointers have been validated there.
int launch_extend_tc(const b200_extend_params* p, cudaStream_t st) {
  const bool fp8 = p->kv_dtype == B200_KV_FP8_E4M3;
  tc::Args a{};
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
  a.k_st_bytes = p->k_stride_t * (fp8 ? 1 : 2);
  a.v_st_bytes = p->v_stride_t * (fp8 ? 1 : 2);
  a.k_scale = p->k_scale;
  a.v_scale = p->v_scale;
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
  a.g = p->hq / p->hkv;
  a.bm_t = tc::kRows / a.g;
  a.sm_scale = p->sm_scale;
  a.logit_cap = p->logit_cap;
  const int nqb = (p->max_extend_len + a.bm_t - 1) / a.bm_t;
  const dim3 grid(nqb, p->hkv, p->bs);
  // TMA for tiles of 128 consecutive rows (bf16 pools; fp8 pools widen their rows on chip and keep the row gather):
  // the new tokens' K / V are described per call (the activations move), the pool's maps are cached per buffer.  A map
  // the driver refuses only switches the feature off.
  tc::TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  static const bool no_tma = getenv("B200_EXTEND_NO_TMA") != nullptr;  // measurement switch
  if (!fp8 && !no_tma) {
    if (p->total_tokens >= tc::kBN &&
        encode_pool_map(&maps.new_k, p->k_new, p->kn_stride_t * 2, p->total_tokens, p->hkv, p->d, false, tc::kBN, 1) == B200_OK &&
        encode_pool_map(&maps.new_v, p->v_new, p->vn_stride_t * 2, p->total_tokens, p->hkv, p->dv, false, tc::kBN, 1) == B200_OK)
      a.tma |= 1;
    // prefix tiles: a paged pool with pages of >= 128 tokens, or a token-granular pool whose allocator hands out runs
    // (page_size B200_PAGE_AUTO); every tile's slot ids are checked on device before they are trusted
    if ((p->page_size >= tc::kBN || p->page_size == B200_PAGE_AUTO) && p->kv_pool_rows >= tc::kBN &&
        cached_pool_map(&maps.pool_k, p->k_buf, a.k_st_bytes, p->kv_pool_rows, p->hkv, p->d, false, tc::kBN, 1) == B200_OK &&
        cached_pool_map(&maps.pool_v, p->v_buf, a.v_st_bytes, p->kv_pool_rows, p->hkv, p->dv, false, tc::kBN, 1) == B200_OK)
      a.tma |= 2;
  }
  return dispatch_tc<tc::kModeExtend>(fp8, p->d, p->act_dtype == B200_ACT_F16, a, maps, grid, st);
}

// Shared-prefix ("cascade") leg of b200_attn_decode (decode.cu): the decode queries of the whole batch against the
// hdr[kHdrShared] leading tokens that every request shares (same slot ids), as one dense tcgen05 contraction per kv
// head split `nsplit` ways along the keys; writes fp32 partials that merge_kernel folds in with the per-request
// suffix partials.  head_dim 128 only; returns B200_ERR_UNSUPPORTED otherwise (the caller then leaves cascade off).
int launch_shared_prefix_tc(const b200_decode_params* p, const CasTables& t, float* cas_o, float* cas_lse, int nsplit,
                            cudaStream_t st) {
  const int g = p->hq / p->hkv;
  const bool fp8 = p->kv_dtype == B200_KV_FP8_E4M3;
  if (p->d != p->dv || tc::kRows % g != 0 || !(p->d == 128 || p->d == 64)) return B200_ERR_UNSUPPORTED;
  tc::Args a{};
  a.q = static_cast<const uint16_t*>(p->q);
  a.q_st = p->q_stride_t;
  a.q_sh = p->q_stride_h;
  a.kbuf = static_cast<const uint8_t*>(p->k_buf);
  a.vbuf = static_cast<const uint8_t*>(p->v_buf);
  a.k_st_bytes = p->k_stride_t * (fp8 ? 1 : 2);
  a.v_st_bytes = p->v_stride_t * (fp8 ? 1 : 2);
  a.r2t = p->req_to_token;
  a.r2t_stride = p->req_to_token_stride;
  a.hq = p->hq;
  a.hkv = p->hkv;
  a.g = g;
  a.bm_t = tc::kRows / g;
  a.sm_scale = p->sm_scale;
  a.logit_cap = p->logit_cap;
  a.cas = t;
  a.cas_o = cas_o;
  a.cas_lse = cas_lse;
  a.k_scale = p->k_scale;
  a.v_scale = p->v_scale;
  a.bs = p->bs;
  const dim3 grid(p->hkv, nsplit, (p->bs + a.bm_t - 1) / a.bm_t);  // z = row-tile lanes; CTAs loop over further tiles
  tc::TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  return dispatch_tc<tc::kModeShared>(fp8, p->d, p->act_dtype == B200_ACT_F16, a, maps, grid, st);
}

}  // namespace b200
