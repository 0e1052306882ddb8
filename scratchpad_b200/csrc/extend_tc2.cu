// Extend / chunked-prefill attention on CTA PAIRS: tcgen05.mma.cta_group::2 (UMMA M = 256 across two SMs), sm_100a.
//
// Same semantics as extend_tc.cu / extend.cu / the reference's extend_attention_fwd
// (scratchpad/nn/attention/triton_attn/extend_attention.py:53-226): query row t of request i sees the gathered prefix
// slots req_to_token[r, :pre_i] (non-causal) and k_new/v_new[start_i : start_i+t+1].  bf16 pools, head_dim 128.
//
// Why pairs.  The one-CTA kernel keeps two 128-row Q tiles per SM with ONE score buffer each
// (S0 | S1 | O0 | O1 = all 512 TMEM columns), so per Q tile the chain  S ready -> softmax -> P.V + next Q.K^T -> S ready
// is serial (3630 clk per 128 keys measured; the tensor pipe and the MUFU are each busy ~56 % of it).  Here a CLUSTER
// of two CTAs owns the 256 query rows: CTA r holds rows [128 r, 128 r + 128) -- one Q tile, one softmax warpgroup -- and
// the MMAs are issued once, by the leader CTA, for both SMs:
//   * every K/V tile is staged HALF per CTA (K: 64 of the 128 keys; V: 64 of the 128 head-dim columns), which is what
//     cta_group::2 reads as the B operand -- half the shared-memory fill and operand traffic per flop and SM;
//   * with one Q tile per CTA the TMEM holds TWO score buffers (Sa | Sb | O = 384 columns): Q.K^T of tile j+1 is issued
//     before the softmax of tile j has finished, S(j+1) is waiting when the softmax warps come back, and the period is
//     the softmax itself, not the sum of the chain.  P (bf16) overwrites its S buffer in place and is the TMEM A operand
//     of P.V, as in the one-CTA kernel.
// Roles per CTA (512 threads): warps 0-7 softmax (two threads per query row = TMEM lane: keys 0..63 / 64..127 of every
// tile), warp 8 MMA issuer (leader CTA only; the peer's warp 8 only allocates TMEM), warps 9-11 idle padding, warps
// 12-15 producers (page-size-1 row gather with 16-byte LDGSTS straight into the UMMA SWIZZLE_128B layout).
// Cross-CTA protocol: barriers the leader's MMA warp waits on (K/V slot full, Q loaded, P written) live in the LEADER's
// shared memory and are arrived on by the threads of both CTAs (mapa + mbarrier.arrive.shared::cluster); barriers the
// CTAs wait on locally (slot empty, S ready, P.V done) are signalled by tcgen05.commit.cta_group::2 ... multicast to the
// same offset in both CTAs.  barrier.cluster brackets the kernel so neither CTA touches the other's barriers, shared
// memory or TMEM before they exist or after they are gone.

#include "common.cuh"
#include "tc_ptx.cuh"

namespace b200 {
namespace tc2 {

using namespace tc;

constexpr int kD = 128;
constexpr int kBN = 128;                    // keys per tile
constexpr int kRowsCta = 128;               // query rows per CTA (one UMMA M half)
constexpr int kRowsPair = 256;              // query rows per cluster
#ifndef B200_TC2_SLOTS
#define B200_TC2_SLOTS 10
#endif
#ifndef B200_TC2_LAG
#define B200_TC2_LAG 6
#endif
constexpr int kSlots = B200_TC2_SLOTS;      // ring of half tiles: K_j (64 keys x 128 d), V_j (128 keys x 64 d), alternating
constexpr int kSlotBytes = 64 * kD * 2;     // 16 KiB
constexpr int kQBytes = kRowsCta * kD * 2;  // 32 KiB
constexpr int kQColBlock = 128 * 128;       // [128 rows x 64 el] SWIZZLE_128B block of the Q tile
constexpr int kKColBlock = 64 * 128;        // [64 keys x 64 el] block of a K half tile
constexpr int kSoftmaxWarps = 8;            // two threads per query row: warps 0-3 keys 0..63 of a tile, warps 4-7 keys 64..127
constexpr int kMmaWarp = 8;
constexpr int kProdWarp0 = 12;
constexpr int kProdThreads = 128;
constexpr int kThreads = 512;
// cp.async groups (of 8 row segments, 16 KiB per CTA) a producer thread keeps in flight.  Measured at C4 with one
// softmax thread per row: 3 groups / 10 slots 922, 5 / 10 912, 8 / 12 896 TFLOP/s -- the gather depth is not what
// bounds the kernel
constexpr int kLag = B200_TC2_LAG;
constexpr int kRegsSoftmax = 168, kRegsOther = 64;  // 256*168 + 256*64 = 59392 <= 64 Ki registers
static_assert(kLag <= kSlots - 3, "the MMA warp consumes K(j+1) before V(j): the producers may not run a whole ring ahead");

constexpr int kSmemQ = 0;
constexpr int kSmemRing = kQBytes;
constexpr int kSmemBar = kSmemRing + kSlots * kSlotBytes;
constexpr int kNumBars = 2 * kSlots + 1 + 2 + 2 + 1 + 1;  // full, empty, q, s[2], p[2 halves], pv, fin
constexpr int kSmemTmemPtr = kSmemBar + kNumBars * 8;
constexpr int kSmemXch = kSmemTmemPtr + 16;           // float [2 parities][2 halves][128 rows]: row-max / row-sum exchange
constexpr int kSmemBytes = kSmemXch + 2 * 2 * kRowsCta * 4 + 1024;  // + alignment slack
static_assert(kSmemBytes <= 232448, "227 KiB of dynamic shared memory per CTA");

constexpr uint32_t kTmemCols = 512;  // power of two >= Sa | Sb | O
constexpr uint32_t kColSa = 0, kColSb = 128, kColO = 256;
constexpr float kRescaleThreshold = 8.0f;  // log2 domain

// ---- cluster-scope PTX -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// this CTA's shared address -> the same offset in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Default semantics (.release.cta), as cutlass::arch::ClusterBarrier::arrive(cta_id) issues it: what the arrivals order
// here is never generic-proxy memory of another CTA -- it is this CTA's own shared memory already made visible to the
// async proxy (cp.async.wait_group + fence.proxy.async) or TMEM (tcgen05.wait::st + tcgen05.fence) -- and a
// cluster-scope release would cost a MEMBAR.ALL.GPU per arrive on the softmax critical path.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Every wait of this kernel is bounded: a barrier that never flips (a protocol bug across the pair) traps instead of
// hanging both SMs.  try_wait suspends the thread in hardware for up to kSuspendHintNs per poll.
__device__ __forceinline__ void mbar_wait2(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  int polls = 0;
  while (!done) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(kSuspendHintNs)
        : "memory");
    if (!done && ++polls > (1 << 20)) __trap();
  }
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// completion of every tcgen05 op issued so far by this thread -> the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc2_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}
__device__ __forceinline__ void tc2_mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc2_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}

// ---- optional wait-time accounting of cluster 0 (build with -DB200_TC2_TRACE; tools/tc2_trace.py) -------------------
// g_tc2_trace[rank][i]: clocks one representative thread of each role spent waiting at each barrier, over the tile.
// Accumulated in registers (tr_[]) and written once when the role ends: a global read-modify-write per wait stalls the
// warp for an L2 round trip and was most of what the first version of this trace measured.
#ifdef B200_TC2_TRACE
__device__ long long g_tc2_trace[2][16];
#define TC2_DECL() unsigned tr_[3] = {0u, 0u, 0u}
#define TC2_T0() const unsigned t_wait0_ = static_cast<unsigned>(clock64())
#define TC2_ACC(k) tr_[k] += static_cast<unsigned>(clock64()) - t_wait0_
#define TC2_FLUSH(base, n)                                                                 \
  do {                                                                                     \
    if (blockIdx.x < 2 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0)                 \
      for (int k_ = 0; k_ < (n); ++k_) g_tc2_trace[rank][(base) + k_] = tr_[k_];           \
  } while (0)
#else
#define TC2_DECL() do {} while (0)
#define TC2_T0() do {} while (0)
#define TC2_ACC(k) do {} while (0)
#define TC2_FLUSH(base, n) do {} while (0)
#endif

struct Args {
  const __nv_bfloat16* q;
  int64_t q_st, q_sh;
  const __nv_bfloat16* k_new;
  const __nv_bfloat16* v_new;
  int64_t kn_st, vn_st;
  const uint8_t* kbuf;
  const uint8_t* vbuf;
  int64_t k_st_bytes, v_st_bytes;
  __nv_bfloat16* o;
  int64_t o_st, o_sh;
  const int32_t* r2t;
  int64_t r2t_stride;
  const void* req_idx;
  const void* seq_lens;
  int idx64;
  const int32_t* ext_lens;
  const int32_t* ext_start;
  int hq, hkv, g, bm_t;  // bm_t = tokens per cluster = 256 / g
  float sm_scale, logit_cap;
};

// grid (2 * query blocks, kv head, request), cluster (2, 1, 1): CTAs 2k and 2k+1 share query block k
__global__ void __launch_bounds__(kThreads, 1) extend_tc2_kernel(const Args a) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_wait();
  pdl_launch_dependents();

  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024-byte alignment
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t rank = cluster_ctarank();
  const bool leader_cta = rank == 0;
  const uint32_t bar0 = sbase + kSmemBar;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };            // leader's copy is the live one
  auto bar_empty = [&](int s) { return bar0 + 8u * (kSlots + s); };  // per CTA (multicast commit)
  const uint32_t bar_q = bar0 + 8u * (2 * kSlots);                  // leader's copy
  auto bar_s = [&](int buf) { return bar0 + 8u * (2 * kSlots + 1 + buf); };      // per CTA
  auto bar_p = [&](int half) { return bar0 + 8u * (2 * kSlots + 3 + half); };    // leader's copy
  const uint32_t bar_pv = bar0 + 8u * (2 * kSlots + 5);             // per CTA
  const uint32_t bar_fin = bar0 + 8u * (2 * kSlots + 6);            // per CTA
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(sgen + kSmemTmemPtr);

  // ---- the cluster's tile (both CTAs compute the same values) ----------------------------------------------------
  const int kvh = blockIdx.y, b = blockIdx.z;
  const int G = a.g, bm_t = a.bm_t;
  const int ext = __ldg(a.ext_lens + b);
  const int nqb = (ext + bm_t - 1) / bm_t;
  const int qb = nqb - 1 - static_cast<int>(blockIdx.x >> 1);  // heavy (late) blocks first
  if (qb < 0) return;                                          // the whole cluster leaves: nothing is shared yet
  const int seq = static_cast<int>(load_idx(a.seq_lens, a.idx64, b));
  const int pre = seq - ext;
  const int start = __ldg(a.ext_start + b);
  const int32_t* row = a.r2t + load_idx(a.req_idx, a.idx64, b) * a.r2t_stride;
  const int t0 = qb * bm_t;
  const int t_end = min(ext, t0 + bm_t);
  const int kv_len = pre + t_end;
  const int nt = (kv_len + kBN - 1) / kBN;
  const int live_rows = (t_end - t0) * G;  // of the cluster's 256

  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < kSlots; ++s) {
        mbar_init(bar_full(s), 2 * kProdThreads / 32);  // one elected arrive per producer warp, both CTAs
        mbar_init(bar_empty(s), 1);
      }
      mbar_init(bar_q, 2 * kSoftmaxWarps);             // one per softmax warp, both CTAs
      mbar_init(bar_s(0), 1);
      mbar_init(bar_s(1), 1);
      mbar_init(bar_p(0), 2 * kSoftmaxWarps / 2);      // one per warp of the key half, both CTAs
      mbar_init(bar_p(1), 2 * kSoftmaxWarps / 2);
      mbar_init(bar_pv, 1);
      mbar_init(bar_fin, 1);
      fence_barrier_init();
    }
    __syncwarp();
    // both CTAs' MMA warp, same shared-memory offset for the result (cute::TMEM::Allocator2Sm's contract)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + kSmemTmemPtr), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  // barriers initialised and TMEM allocated in BOTH CTAs before anyone signals or issues across the pair
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  if (warp >= kProdWarp0) {
    // =========================== producers ===========================================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsOther));
    const int pw = warp - kProdWarp0;
    // K half tile: keys [64 rank, +64) of the tile, whole 256-byte rows: 16 chunks per row, a warp instruction
    // covers 2 rows; thread = chunk kch of rows krr + 8 q (row & 7 constant)
    const int kch = lane & 15, krr = 2 * pw + (lane >> 4);
    const uint32_t k_dst_thread = (kch >> 3) * kKColBlock + krr * 128 + (((kch & 7) ^ (krr & 7)) << 4);
    const int64_t k_col_bytes = static_cast<int64_t>(kvh) * (kD * 2) + kch * 16;
    // V half tile: all 128 keys, head-dim columns [64 rank, +64) = one 128-byte half of each row: 8 chunks per
    // row, a warp instruction covers 4 rows; thread = chunk vch of rows vrr + 16 q
    const int vch = lane & 7, vrr = 4 * pw + (lane >> 3);
    const uint32_t v_dst_thread = vrr * 128 + ((vch ^ (vrr & 7)) << 4);
    const int64_t v_col_bytes = static_cast<int64_t>(kvh) * (kD * 2) + rank * 128 + vch * 16;
    const uint32_t full_remote0 = mapa(bar_full(0), 0);  // the leader's full barriers (cluster address)
    const int total = 2 * nt;
    TC2_DECL();
    // Page-table lookups run ONE group ahead of the copies they feed: a group's 8 slot ids are loaded while the previous
    // group's copies are issued, so the warp never sits out an L2 round trip between lookup and copy (it did: 2 x ~900
    // clk per tile, the whole kernel ran at the producers' pace -- tools/tc2_trace.py).
    int slot_nx[8];
    auto fetch_slots = [&](int n) {
      const int j = n >> 1, is_v = n & 1;
      if (n < total && (j + 1) * kBN <= pre) {
        const int p0 = j * kBN + (is_v ? vrr : 64 * static_cast<int>(rank) + krr), step = is_v ? 16 : 8;
#pragma unroll
        for (int q = 0; q < 8; ++q) slot_nx[q] = __ldg(row + p0 + step * q);
      }
    };
    fetch_slots(0);
    for (int n = 0; n < total; ++n) {
      const int j = n >> 1, is_v = n & 1, s = n % kSlots;
      {
        TC2_T0();
        mbar_wait2(bar_empty(s), ((n / kSlots) & 1) ^ 1);
        TC2_ACC(0);
      }
      const uint32_t dst0 = sbase + kSmemRing + s * kSlotBytes + (is_v ? v_dst_thread : k_dst_thread);
      const uint8_t* pool = is_v ? a.vbuf + v_col_bytes : a.kbuf + k_col_bytes;
      const int64_t pool_st = is_v ? a.v_st_bytes : a.k_st_bytes;
      const int step = is_v ? 16 : 8;
      if ((j + 1) * kBN <= pre) {
        // tile of cached-prefix keys only: one copy per row segment, slot ids fetched one group ago
        int slot[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) slot[q] = slot_nx[q];
        fetch_slots(n + 1);
#pragma unroll
        for (int q = 0; q < 8; ++q) cp_async16(dst0 + q * (step * 128), kv_row_ptr(pool, slot[q], pool_st), 16);
      } else {
        const uint8_t* fresh = reinterpret_cast<const uint8_t*>(is_v ? a.v_new : a.k_new) + (is_v ? v_col_bytes : k_col_bytes);
        const int64_t f_st = (is_v ? a.vn_st : a.kn_st) * 2;
        const int p0 = j * kBN + (is_v ? vrr : 64 * static_cast<int>(rank) + krr);
#pragma unroll 4
        for (int q = 0; q < 8; ++q) {
          const int p = p0 + step * q, pc = min(p, kv_len - 1);
          const uint8_t* src = pc < pre ? kv_row_ptr(pool, __ldg(row + pc), pool_st) : fresh + static_cast<int64_t>(start + pc - pre) * f_st;
          cp_async16(dst0 + q * (step * 128), src, p < kv_len ? 16 : 0);  // keys past the end: zeros
        }
      }
      cp_async_commit();
      // One arrive per WARP: every lane has waited for its own copies and fenced them towards the async proxy, the
      // warp barrier orders that before lane 0's arrive.  (One arrive per thread -- 512 remote mbarrier updates per
      // tile from the peer CTA -- serialised on the leader's barriers: 2540 clk per tile whatever else was changed.)
      if (n >= kLag) {
        {
          TC2_T0();
          cp_async_wait<kLag>();
          TC2_ACC(1);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(full_remote0 + 8u * ((n - kLag) % kSlots));
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0)
      for (int n = max(total - kLag, 0); n < total; ++n) mbar_arrive_cluster(full_remote0 + 8u * (n % kSlots));
    if (warp == kProdWarp0) TC2_FLUSH(0, 2);
  } else if (warp == kMmaWarp) {
    // =========================== MMA issuer (leader CTA) ================================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsOther));
    if (leader_cta) {
      // the whole warp runs the loop so that addresses / descriptors stay warp-uniform; lane 0 issues
      const bool leader = lane == 0;
      TC2_DECL();
      constexpr uint32_t idesc_qk = instr_desc(kRowsPair, kBN, false);
      constexpr uint32_t idesc_pv = instr_desc(kRowsPair, kD, true);
      const uint32_t q_base = sbase + kSmemQ;
      const uint32_t ring = sbase + kSmemRing;
      auto slot_of = [&](int n) { return n % kSlots; };
      auto wait_full = [&](int n) {
        TC2_T0();
        mbar_wait2(bar_full(slot_of(n)), (n / kSlots) & 1);
        TC2_ACC(n & 1);  // 0: K tiles, 1: V tiles
        tc_fence_after();
      };
      auto issue_qk = [&](int j) {  // S[j & 1] = Q . K_j^T   (A: 128 rows per CTA; B: 64 keys per CTA)
        const int n = 2 * j;
        wait_full(n);
        const uint32_t d = tmem + ((j & 1) ? kColSb : kColSa);
        const uint32_t qlo = desc_lo(q_base, 16), klo = desc_lo(ring + slot_of(n) * kSlotBytes, 16);
#pragma unroll
        for (int kc = 0; kc < kD / 16; ++kc) {
          const uint32_t qoff = ((kc >> 2) * kQColBlock + (kc & 3) * 32) >> 4;
          const uint32_t koff = ((kc >> 2) * kKColBlock + (kc & 3) * 32) >> 4;
          if (leader) tc2_mma_ss(d, desc_from(qlo + qoff, 1024), desc_from(klo + koff, 1024), idesc_qk, kc > 0);
        }
        if (leader) {
          tc2_commit(bar_s(j & 1));
          tc2_commit(bar_empty(slot_of(n)));
        }
      };
      // O (+)= P_j . V_j over keys [64 half, +64)
      auto issue_pv = [&](int j, int half) {
        const uint32_t p = tmem + ((j & 1) ? kColSb : kColSa);
        const uint32_t vlo = desc_lo(ring + slot_of(2 * j + 1) * kSlotBytes, 16);
#pragma unroll
        for (int k4 = 0; k4 < kBN / 32; ++k4) {
          const int ks = half * (kBN / 32) + k4;
          if (leader) tc2_mma_ts(tmem + kColO, p + ks * 8, desc_from(vlo + ((ks * 2048) >> 4), 1024), idesc_pv, !(j == 0 && ks == 0));
        }
      };
      mbar_wait2(bar_q, 0);
      tc_fence_after();
      issue_qk(0);
      for (int j = 0; j < nt; ++j) {
        // S(j+1) into the other score buffer while the softmax warps work on S(j): its previous content, P(j-1),
        // was consumed by P.V(j-1), issued before this and executed in order
        if (j + 1 < nt) issue_qk(j + 1);
        wait_full(2 * j + 1);
        // both key halves of P(j) -- and with them both partners' share of a lazy rescale of O, which P.V writes in
        // full width -- before the first k-step
        {
          TC2_T0();
          mbar_wait2(bar_p(0), j & 1);
          mbar_wait2(bar_p(1), j & 1);
          TC2_ACC(2);
        }
        tc_fence_after();
        issue_pv(j, 0);
        issue_pv(j, 1);
        if (leader) {
          tc2_commit(bar_pv);
          tc2_commit(bar_empty(slot_of(2 * j + 1)));
        }
      }
      if (leader) tc2_commit(bar_fin);
      TC2_FLUSH(2, 3);
    }
    // every asynchronous arrival of the tile has landed in this CTA before it leaves
    mbar_wait2(bar_fin, 0);
    tc_fence_after();
    __syncwarp();
  } else if (warp > kMmaWarp) {
    // idle padding warps of the MMA warpgroup: they only take part in the CTA / cluster barriers
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsOther));
  } else {
    // =========================== softmax warpgroups ======================================================
    // Two threads per query row: thread (warp q, lane) of warpgroup 0 owns keys 0..63 of every tile of row 32 q + lane,
    // the same thread of warpgroup 1 keys 64..127.  With one softmax warp per SM sub-partition the row pass was
    // latency-bound (2540 clk per tile, 922 TFLOP/s at C4: nothing overlaps a warp's own TMEM-load / MUFU / store
    // latencies); two warps per sub-partition on the same tile halve each warp's work and overlap each other.  The
    // partners exchange one float per tile (their partial row maximum) through shared memory and keep everything else
    // private: partial row sums are added in the epilogue, O's columns are split between them for the lazy rescale and
    // the output, and P's two key halves are exactly the two halves the MMA warp waits for.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsSoftmax));
    // Q tile of this CTA: rows R = 128 rank + r -> (token t0 + R / G, head kvh G + R % G); chunk `ch` of rows
    // r_base + 16 it (row & 7 constant per thread)
    constexpr int kQIters = kRowsCta / 16;
    const int ch = tid & 15, r_base = tid >> 4;
    const uint32_t q_dst0 = sbase + kSmemQ + (ch >> 3) * kQColBlock + (((ch & 7) ^ (r_base & 7)) << 4);
    {
      const int R0 = kRowsCta * static_cast<int>(rank) + r_base;
      int t = R0 / G, h = R0 % G;
      const int dt = 16 / G, dh = 16 % G;
#pragma unroll
      for (int it = 0; it < kQIters; ++it) {
        const bool live = R0 + 16 * it < live_rows;
        const __nv_bfloat16* src = a.q + static_cast<int64_t>(start + t0 + (live ? t : 0)) * a.q_st +
                                   static_cast<int64_t>(kvh * G + (live ? h : 0)) * a.q_sh + ch * 8;
        cp_async16(q_dst0 + (r_base + 16 * it) * 128, src, live ? 16 : 0);
        t += dt;
        h += dh;
        if (h >= G) {
          h -= G;
          ++t;
        }
      }
    }
    cp_async_commit();
    cp_async_wait<0>();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(mapa(bar_q, 0));

    const int half = warp >> 2, wq = warp & 3;  // key half of the tile; TMEM lane quarter
    const uint32_t p_remote = mapa(bar_p(half), 0);
    const int r = wq * 32 + lane;  // row inside the CTA's tile == TMEM lane
    const int R = kRowsCta * static_cast<int>(rank) + r;
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(wq * 32) << 16);
    const uint32_t o_addr = lane_base + kColO + 64 * half;  // this thread's half of the O row
    float* xch = reinterpret_cast<float*>(sgen + kSmemXch);   // [parity][half][row]
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + wq) : "memory"); };  // warps wq and wq + 4
    // causal limit (dead rows behave like the last live token)
    const int lim = pre + min(t0 + R / G, t_end - 1);
    const float cap = a.logit_cap;
    const float sm_scale = a.sm_scale;
    const float sc_log2 = sm_scale * kLog2e;
    float m = -INFINITY, l = 0.f;  // m: of the whole row (both partners hold the same value); l: of this thread's keys
    TC2_DECL();
#ifdef B200_TC2_TRACE
    const long long t_loop0 = clock64();
#endif
    for (int j = 0; j < nt; ++j) {
      const uint32_t s_addr = lane_base + ((j & 1) ? kColSb : kColSa);
      {
        TC2_T0();
        mbar_wait2(bar_s(j & 1), (j >> 1) & 1);
        TC2_ACC(0);
      }
      tc_fence_after();
      const int climit = lim - j * kBN - 64 * half;  // this thread's columns <= climit are visible
      const bool masked = climit < 63;
      // interior tile for the whole warp: no mask, no soft cap.  Must be warp-uniform (.sync.aligned tcgen05.ld/st)
      const bool fast = !__any_sync(0xffffffffu, masked) && !(cap > 0.f);
      uint32_t v[64];  // this thread's half of the row of S, raw scores
      tc_ld32(s_addr + 64 * half, v);
      tc_ld32(s_addr + 64 * half + 32, v + 32);
      tc_wait_ld();
      // P = 2^(s - m) -> bf16, written over S in place (P is the TMEM A operand of the P.V MMA): keys [64 half, +64)
      // are bf16 pairs in columns [32 half, +32) of the score buffer.  Interior tiles: packed f32x2 arithmetic; the
      // row maximum of the tile (four independent chains) rides along in the shadow of the exponentials.
      auto exp_store_fast = [&](float mcur, float l_in, float& mx_out) -> float {
        const uint64_t sc2 = pack_f32x2(sc_log2, sc_log2), nm2 = pack_f32x2(-mcur, -mcur);
        uint64_t l2 = pack_f32x2(l_in, 0.f);
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
        for (int c4 = 0; c4 < 2; ++c4) {
          uint32_t pk[16];
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            const float x0 = __uint_as_float(v[c4 * 32 + k]), x1 = __uint_as_float(v[c4 * 32 + k + 1]);
            const float x2 = __uint_as_float(v[c4 * 32 + k + 2]), x3 = __uint_as_float(v[c4 * 32 + k + 3]);
            m0 = fmaxf(m0, x0);
            m1 = fmaxf(m1, x1);
            m2 = fmaxf(m2, x2);
            m3 = fmaxf(m3, x3);
            const uint64_t ya = fma_f32x2(pack_f32x2(x0, x1), sc2, nm2), yb = fma_f32x2(pack_f32x2(x2, x3), sc2, nm2);
            float y0, y1, y2, y3;
            unpack_f32x2(ya, y0, y1);
            unpack_f32x2(yb, y2, y3);
            const float p0 = ex2(y0), p1 = ex2(y1), p2 = ex2(y2), p3 = ex2(y3);
            l2 = add_f32x2(l2, add_f32x2(pack_f32x2(p0, p1), pack_f32x2(p2, p3)));
            pk[k >> 1] = pack_bf16x2(p0, p1);
            pk[(k >> 1) + 1] = pack_bf16x2(p2, p3);
          }
          tc_st16(s_addr + 32 * half + c4 * 16, pk);
        }
        mx_out = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        float la, lb;
        unpack_f32x2(l2, la, lb);
        return la + lb;
      };
      // Optimistic order for interior tiles behind the first: exponentials FIRST, against the running maximum m, and the
      // tile's own maximum afterwards.  The rescale is lazy anyway (m may lag the true maximum by 2^8), so in all but a
      // few tiles of a row the check below finds nothing to do and the separate max pass -- which kept both partner
      // warps of a sub-partition off the MUFU at the same time -- is gone.  If a row's maximum did jump, O is rescaled
      // and P recomputed from the scores still in registers (an overflowed P is simply overwritten).
      const bool optimistic = fast && j > 0;
      const float l_before = l;
      float mx = -INFINITY;
      if (optimistic) {
        l = exp_store_fast(m, l, mx);
      } else if (masked) {
#pragma unroll
        for (int k = 0; k < 64; ++k)
          if (k <= climit) mx = fmaxf(mx, __uint_as_float(v[k]));
      } else {
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
        for (int k = 0; k < 64; k += 8) {
          m0 = fmaxf(m0, fmaxf(__uint_as_float(v[k]), __uint_as_float(v[k + 1])));
          m1 = fmaxf(m1, fmaxf(__uint_as_float(v[k + 2]), __uint_as_float(v[k + 3])));
          m2 = fmaxf(m2, fmaxf(__uint_as_float(v[k + 4]), __uint_as_float(v[k + 5])));
          m3 = fmaxf(m3, fmaxf(__uint_as_float(v[k + 6]), __uint_as_float(v[k + 7])));
        }
        mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      }
      // row maximum = max of the partners' halves: one float each way through shared memory.  Slots alternate with the
      // tile parity: a slot is rewritten two tiles later, behind the next tile's pair barrier, i.e. after it was read
      xch[((j & 1) * 2 + half) * kRowsCta + r] = mx;
      {
        TC2_T0();
        pair_sync();
        TC2_ACC(1);
      }
      mx = fmaxf(mx, xch[((j & 1) * 2 + (half ^ 1)) * kRowsCta + r]);
      // softcap is monotone, so the maximum of the capped scores is the capped maximum
      const float mxs = cap > 0.f ? softcap(mx * sm_scale, cap) * kLog2e : mx * sc_log2;
      bool write_p = !optimistic;
      if (j == 0) {
        m = mxs;  // key 0 is visible to every row, so this is finite
      } else {
        const bool need = mxs > m + kRescaleThreshold;  // the same decision in both partner warps (same rows, same values)
        if (__any_sync(0xffffffffu, need)) {
          // lazy rescale: O *= 2^(m_old - m_new), this thread's 64 columns; P.V(j-1) must have landed, P.V(j) waits
          // for both partners' arrives
          mbar_wait2(bar_pv, (j - 1) & 1);
          tc_fence_after();
          const float mn = fmaxf(m, mxs);
          const float alpha = ex2(m - mn);
          m = mn;
          l = l_before * alpha;
          write_p = true;
#pragma unroll 1
          for (int c4 = 0; c4 < 2; ++c4) {
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
      if (write_p) {  // warp-uniform
        if (fast) {
          float unused;
          l = exp_store_fast(m, l, unused);
        } else {
#pragma unroll
          for (int c4 = 0; c4 < 2; ++c4) {
            uint32_t pk[16];
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
              const float x0 = __uint_as_float(v[c4 * 32 + k]), x1 = __uint_as_float(v[c4 * 32 + k + 1]);
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
              pk[k >> 1] = pack_bf16x2(p0, p1);
            }
            tc_st16(s_addr + 32 * half + c4 * 16, pk);
          }
        }
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(p_remote);  // this warp's rows of this key half of P(j) are in TMEM
    }
    // epilogue (thread = row and half of the head dim): O / l as bf16.  A parity wait only tells the two most recent
    // phases apart, and with two score buffers the softmax of the last tile can finish before P.V(nt-2) has even
    // completed (S(nt-1) was ready long before): wait for the phases in order, or the wait for P.V(nt-1) would pass on
    // the state "P.V(nt-3) complete".
#ifdef B200_TC2_TRACE
    if (blockIdx.x < 2 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0) {
      g_tc2_trace[rank][7] = clock64() - t_loop0;  // the softmax loop, all tiles
      g_tc2_trace[rank][8] = nt;
    }
    if (warp == 0) TC2_FLUSH(5, 2);
#endif
    xch[(((nt & 1) * 2) + half) * kRowsCta + r] = l;  // the slot parity the last tile did not use
    pair_sync();
    l += xch[(((nt & 1) * 2) + (half ^ 1)) * kRowsCta + r];
    if (nt >= 2) mbar_wait2(bar_pv, (nt - 2) & 1);
    mbar_wait2(bar_pv, (nt - 1) & 1);
    tc_fence_after();
    const bool live = R < live_rows;
    const float inv = l > 0.f ? 1.f / l : 0.f;
    __nv_bfloat16* op = a.o + static_cast<int64_t>(start + t0 + (live ? R / G : 0)) * a.o_st +
                        static_cast<int64_t>(kvh * G + (live ? R % G : 0)) * a.o_sh + 64 * half;
#pragma unroll 1
    for (int c4 = 0; c4 < 2; ++c4) {
      uint32_t w[32];
      tc_ld32(o_addr + c4 * 32, w);
      tc_wait_ld();
      if (live) {
#pragma unroll
        for (int k = 0; k < 32; k += 8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(w[k]) * inv, __uint_as_float(w[k + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(w[k + 2]) * inv, __uint_as_float(w[k + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(w[k + 4]) * inv, __uint_as_float(w[k + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(w[k + 6]) * inv, __uint_as_float(w[k + 7]) * inv);
          *reinterpret_cast<uint4*>(op + c4 * 32 + k) = u;
        }
      }
    }
  }

  // neither CTA may release its shared memory, barriers or TMEM while the other can still reach them
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }
}

}  // namespace tc2

#ifdef B200_TC2_TRACE
extern "C" int b200_debug_tc2_trace_read(long long* host_dst, int reset) {
  if (cudaMemcpyFromSymbol(host_dst, tc2::g_tc2_trace, sizeof(tc2::g_tc2_trace)) != cudaSuccess) return 1;
  if (reset) {
    long long zero[2][16] = {};
    if (cudaMemcpyToSymbol(tc2::g_tc2_trace, zero, sizeof(zero)) != cudaSuccess) return 1;
  }
  return 0;
}
#endif

// Called by b200_attn_extend (extend.cu): head_dim 128 over a bf16 pool.  Strides and pointers have been validated there.
int launch_extend_tc2(const b200_extend_params* p, cudaStream_t st) {
  tc2::Args a{};
  a.q = static_cast<const __nv_bfloat16*>(p->q);
  a.q_st = p->q_stride_t;
  a.q_sh = p->q_stride_h;
  a.k_new = static_cast<const __nv_bfloat16*>(p->k_new);
  a.v_new = static_cast<const __nv_bfloat16*>(p->v_new);
  a.kn_st = p->kn_stride_t;
  a.vn_st = p->vn_stride_t;
  a.kbuf = static_cast<const uint8_t*>(p->k_buf);
  a.vbuf = static_cast<const uint8_t*>(p->v_buf);
  a.k_st_bytes = p->k_stride_t * 2;
  a.v_st_bytes = p->v_stride_t * 2;
  a.o = static_cast<__nv_bfloat16*>(p->o);
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
  a.bm_t = tc2::kRowsPair / a.g;
  a.sm_scale = p->sm_scale;
  a.logit_cap = p->logit_cap;
  const int nqb = (p->max_extend_len + a.bm_t - 1) / a.bm_t;
  static PerDevice pd;
  if (int rc = ensure_smem_attr(tc2::extend_tc2_kernel, tc2::kSmemBytes, pd)) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * nqb, p->hkv, p->bs);
  cfg.blockDim = dim3(tc2::kThreads);
  cfg.dynamicSmemBytes = tc2::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static const bool no_pdl = getenv("B200_NO_PDL") != nullptr;
  cfg.numAttrs = no_pdl ? 1 : 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, tc2::extend_tc2_kernel, a);
  if (e != cudaSuccess) return 1000 + static_cast<int>(e);
  return B200_OK;
}

}  // namespace b200
