// Paged decode attention for B200 (sm_100a): plan -> persistent split-KV
// flash-decoding kernel -> log-sum-exp merge.
//
// Replaces, behind the same AttentionBackend surface, what the reference does
// with triteia/Triton `decode_attention_fwd` (scratchpad/nn/attention/
// triton_backend.py:183-195; math as in triton_attn/decode_attention.py:251-434)
// and FlashInfer's BatchDecodeWithPagedKVCacheWrapper
// (flashinfer_backend.py:475-482, 623-670).  Not a port of either: see DESIGN.md.
//
// Work decomposition
//   * A KV "tile" is TN consecutive token positions of one request, all
//     HPC kv heads of the CTA's head group: for every token one contiguous
//     HPC*D*elt-byte span of the pool row (page size is 1, so a token row is
//     the gather unit; keeping all local heads of a row in one CTA makes the
//     DRAM access 512 B - 1 KiB contiguous instead of 128 B).
//   * plan_kernel numbers all tiles of the batch 0..T-1 (prefix sum of
//     ceil(seq_len/TN)); CTA c of the persistent grid owns the contiguous tile
//     range [c*tpc, (c+1)*tpc).  Work is balanced to one tile regardless of
//     how sequence lengths are distributed, with no host sync and a grid that
//     does not depend on device data (CUDA-graph safe).
//   * A CTA streams its tiles through a 3-stage cp.async (LDGSTS, 16 B per
//     thread, XOR-swizzled) pipeline that keeps running across request
//     boundaries; softmax state is flushed as one partial (O/l, lse) per
//     (CTA, request, token sub-slice) into slot (c + b) of the workspace.
//   * merge_kernel combines the <= few partials of each (request, q head).
//
// Tensor-core mapping (GQA query-group packing, "swapped" orientation):
//   S^T[16 tokens x 8 q-heads] = K[16 x D] . Q^T[D x 8]      (A = K via ldmatrix)
//   O^T[D x 8 q-heads]        += V^T[D x 16] . P^T[16 x 8]    (A = V via ldmatrix.trans)
// so the g q-heads that share a kv head fill the N=8 side of m16n8k16 (g=8:
// dense, g=4: half) and the softmax reductions run over tokens with warp
// shuffles.  P^T moves from accumulator to B-operand layout with movmatrix.
// fp8-e4m3 KV is dequantised in registers (cvt.rn.f16x2.e4m3x2) and multiplied
// in f16; per-head scales are folded into the softmax scale (K) and the
// epilogue (V).

#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include <cstring>
#include <cstdio>
#include <mutex>
#include <type_traits>

#include "common.cuh"
#include "tensor_map.cuh"

namespace b200 {

// Ring depth and CTAs per SM of the split-KV kernels.  3 stages of 32 KiB x 2 CTAs per SM is the shipped point; deeper
// rings with one CTA per SM (-DB200_DEC_STAGES=5..7 -DB200_DEC_MINBLOCKS=1: up to 192 KiB in flight per SM) are a build
// option for A/B runs
#ifndef B200_DEC_STAGES
#define B200_DEC_STAGES 3
#endif
#ifndef B200_DEC_MINBLOCKS
#define B200_DEC_MINBLOCKS 2
#endif
constexpr int kStages = B200_DEC_STAGES;
constexpr int kMinBlocks = B200_DEC_MINBLOCKS;
// Partials per (cta, request): the token sub-slices of a tile are combined inside the CTA, so one.  The two build
// switches below exist for A/B measurements of exactly these two design decisions (tools/gpu_ab_lib.sh):
//   -DB200_NO_CTA_REDUCE  every sub-slice flushes its own partial (round-1 behaviour)
//   -DB200_ONE_LOOP       the TMA-capable kernel always runs its two-producer loop (round-1 behaviour)
#ifdef B200_NO_CTA_REDUCE
constexpr bool kCtaReduce = false;
#else
constexpr bool kCtaReduce = true;
#endif
#ifdef B200_ONE_LOOP
constexpr bool kTwoLoops = false;
#else
constexpr bool kTwoLoops = true;
#endif
// bytes of reduction scratch the in-CTA combination of `nsub` sub-slice states needs (first tree round: nsub / 2 writers
// per virtual head, per lane 4 * (d / 16) accumulator floats + m[2] + l[2]) -- it must fit the ring stage just consumed
constexpr int reduce_scratch_bytes(int nsub, int vh, int d) { return (nsub / 2) * vh * (4 * (d / 16) + 4) * 32 * 4; }
constexpr bool cta_reduce_fits(int nsub, int vh, int d, int stage_bytes) {
  return kCtaReduce && nsub > 1 && reduce_scratch_bytes(nsub, vh, d) <= stage_bytes;
}
constexpr int kPlanMagic = 0x62323030;  // "b200"
constexpr int kHdrInts = 16;
// header words: 0 total tiles, 1 tiles per static chunk (= per CTA when nothing is dynamic), 2 chunks, 3 bs, 4 tile
// tokens, 5 grid, 6 magic, 7 / 8: kHdrShared / kHdrCasTiles (common.cuh), 9 tiles per dynamic chunk, 10 first dynamic tile
//
// Tile schedule.  The batch's KV tiles are one sequence cut into CHUNKS of consecutive tiles; a partial (O, lse) is
// flushed per (chunk, request), slot chunk + request -- unique along the monotone staircase, no atomics.  Chunk x < grid
// is CTA x's own (no claim); when a CTA's share is long enough (>= kDynMinBase tiles, ~360 us of streaming) AND the plan
// finds the slots of most requests handed out as runs (whole tiles by TMA), only 7/8 of
// the tiles are handed out that way and the last eighth is cut into chunks of 1/32 share which CTAs CLAIM as they run
// dry (one atomicAdd per chunk and head group).  Why: with equal static ranges the CTAs of one launch finish up to 8 %
// apart -- per-SM bandwidth differs, not the work (tools/dec_trace.py: C3B over 64-token pages, loop done at 545 .. 623
// us, mean 575) -- and the kernel ends with its slowest CTA.  Chunk boundaries are closed-form (sched_*), so the merge
// finds a request's partials without a table.
constexpr int kDynMinBase = 256;     // tiles per CTA below which everything stays static (a chunk switch costs ~3 us;
                                     // measured: C3B, 443 tiles per CTA, gains 2-3 %; C3U, 222, loses 3 % to the merge)
constexpr int kDynChunksPerCta = 4;  // dynamic chunks per CTA on average: (1/8 share) / (1/32 share)
constexpr int kMaxHeadGroups = 64;   // claim counters: one per blockIdx.y
#ifdef B200_STATIC_TILES
constexpr bool kDynamicTail = false;  // A/B switch: equal static ranges only (round 1 / early round 2)
#else
constexpr bool kDynamicTail = true;
#endif
__host__ __device__ inline int max_chunks(int grid) { return grid * (1 + kDynChunksPerCta) + 1; }
struct Sched {
  int total, grid, s_tiles, d_tiles, dyn0, nch;
};
// dyn: {tiles per CTA from which the tail is dynamic, denominator of the dynamic fraction, denominator of the chunk size}
struct DynParams {
  int min_base, frac_den, chunk_den;
};
__host__ __device__ inline Sched make_sched(int total, int grid, DynParams dyn) {
  const int dyn_min_base = dyn.min_base;
  Sched sc;
  sc.total = total;
  sc.grid = grid;
  int base = (total + grid - 1) / grid;
  if (base < 1) base = 1;
  sc.s_tiles = base;
  sc.d_tiles = 1;
  if (kDynamicTail && base >= dyn_min_base && base >= 8) {
    sc.s_tiles = base - base / dyn.frac_den;
    sc.d_tiles = (base + dyn.chunk_den - 1) / dyn.chunk_den;  // ceil: at most chunk_den / frac_den (<= 4) * grid dynamic chunks (+1)
  }
  sc.dyn0 = grid * sc.s_tiles;
  if (sc.dyn0 >= total) {
    sc.dyn0 = total;
    sc.nch = (total + sc.s_tiles - 1) / sc.s_tiles;
  } else {
    sc.nch = grid + (total - sc.dyn0 + sc.d_tiles - 1) / sc.d_tiles;
  }
  return sc;
}
// the dynamic tail only where the plan's probe found most requests' slots handed out as runs
__device__ __forceinline__ DynParams runs_dyn(DynParams dyn, int probed, int runs) {
  if (!(probed > 0 && 2 * runs >= probed)) dyn.min_base = 1 << 30;
  return dyn;
}
__host__ __device__ __forceinline__ int sched_chunk_begin(int c, int grid, int s_tiles, int d_tiles) {
  return c < grid ? c * s_tiles : grid * s_tiles + (c - grid) * d_tiles;
}
__host__ __device__ __forceinline__ int sched_chunk_of(int t, int dyn0, int grid, int s_tiles, int d_tiles) {
  return t < dyn0 ? t / s_tiles : grid + (t - dyn0) / d_tiles;
}
constexpr int kExtraCasTiles = 4;    // row tiles beyond ceil(bs / requests-per-tile): room for that many more prefix groups
constexpr int kMaxGroupBs = 1024;    // prefix groups are found by one CTA; larger batches run without the cascade
constexpr int kMaxCasSplits = 8;     // KV splits of the shared-prefix leg (workspace is sized for this many)
constexpr int kCasTile = 128;        // its key-tile size (extend_tc.cu)
constexpr int kMinShared = 1024;     // shorter common prefixes stay on the split-KV kernel
constexpr int kMinCasRows = 64;      // bs * g below this: the dense leg would be mostly padding

struct WsLayout {
  size_t tile_off, seq_len, req_idx, cta_b0, claim, part_lse, part_o, cas_lse, cas_o, total;
  // prefix groups (cascade): per request / per sorted position / per row tile of the shared-prefix leg
  size_t shared32, perm, grp_of, grp_min, grp_size, grp_fin, ct_first, ct_cnt, ct_leader, ct_shared;
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static WsLayout ws_layout(int max_bs, int grid, int nsub, int hq, int dv) {
  WsLayout L;
  size_t off = kHdrInts * sizeof(int);
  L.tile_off = off;
  off = align_up(off + (size_t)(max_bs + 1) * sizeof(int), 16);
  L.seq_len = off;
  off = align_up(off + (size_t)max_bs * sizeof(int), 16);
  L.req_idx = off;
  off = align_up(off + (size_t)max_bs * sizeof(int), 16);
  L.cta_b0 = off;  // request holding the first tile of each chunk
  off = align_up(off + (size_t)max_chunks(grid) * sizeof(int), 256);
  L.claim = off;   // claim counters of the dynamic chunks, one per head group (zeroed by the plan and by every merge)
  off = align_up(off + (size_t)kMaxHeadGroups * sizeof(int), 256);
  size_t nparts = (size_t)(max_chunks(grid) + max_bs) * nsub * hq;
  L.part_lse = off;
  off = align_up(off + nparts * sizeof(float), 256);
  L.part_o = off;
  off = align_up(off + nparts * dv * sizeof(float), 256);
  // shared-prefix ("cascade") partials: [split, request, q head], indexed with a stride of bs (not max_bs); the
  // cascade only runs for bs <= kMaxGroupBs, so a backend sized for thousands of request rows does not pin
  // kMaxCasSplits * max_bs * hq * dv floats it can never use
  const int cas_bs = max_bs < kMaxGroupBs ? max_bs : kMaxGroupBs;
  const size_t ncas = (size_t)kMaxCasSplits * cas_bs * hq;
  L.cas_lse = off;
  off = align_up(off + ncas * sizeof(float), 256);
  L.cas_o = off;
  off = align_up(off + ncas * dv * sizeof(float), 256);
  size_t* ints[] = {&L.shared32, &L.perm, &L.grp_of, &L.grp_min, &L.grp_size, &L.grp_fin, &L.ct_first, &L.ct_cnt, &L.ct_leader, &L.ct_shared};
  for (size_t* f : ints) {
    *f = off;
    off = align_up(off + (size_t)cas_bs * sizeof(int), 16);
  }
  L.total = align_up(off, 256);
  return L;
}

// ---------------------------------------------------------------------------
// shared-prefix detection.  RadixAttention makes requests that share a prompt prefix share its slot ids
// (memory/radix_cache.py:105-135), so equal slot ids in req_to_token ARE the sharing information and reading it
// here needs nothing from the scheduler.
//   group_kernel (1 CTA): requests whose FIRST slot id is equal form a prefix group (a slot belongs to one radix
//     node, so two requests share a prefix iff they share position 0).  Stable rank of the keys -> perm (sorted
//     position -> request), group id = sorted position of the group's first member, group sizes.  Rows the CUDA
//     graph runner pads the batch with (seq_len <= 1, cuda_graph_runner.py:457-465) get unique keys.
//   lcp_kernel (1 CTA per request): length of the common prefix of the request's row and its group leader's row,
//     at most seq - 1 (the newest token always stays in the split-KV leg); atomicMin into the group's slot.
//   plan_kernel then rounds a group's length down to 128 tokens, drops it below kMinShared or when the group has
//     fewer than kMinCasRows query rows per kv head, and cuts the group into row tiles for the dense leg.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) group_kernel(int* __restrict__ perm, int* __restrict__ grp_of, int* __restrict__ grp_min,
                                                     int* __restrict__ grp_size, const int32_t* __restrict__ r2t,
                                                     int64_t r2t_stride, const void* __restrict__ req,
                                                     const void* __restrict__ seq, int idx64, int bs) {
  __shared__ int key[kMaxGroupBs];
  __shared__ int sorted_key[kMaxGroupBs];
  __shared__ int sorted_req[kMaxGroupBs];
  __shared__ int gstart[kMaxGroupBs];
  __shared__ int gsize[kMaxGroupBs];
  const int tid = threadIdx.x;  // bs <= kMaxGroupBs == blockDim.x: one request / sorted position per thread
  if (tid < bs) {
    const int64_t s = load_idx(seq, idx64, tid);
    // real slot ids are > 0; padded / empty rows get distinct negative keys
    key[tid] = s >= 2 ? __ldg(r2t + load_idx(req, idx64, tid) * r2t_stride) : -1 - tid;
    gsize[tid] = 0;
  }
  __syncthreads();
  if (tid < bs) {
    const int k = key[tid];
    int rank = 0;
    for (int j = 0; j < bs; ++j) rank += (key[j] < k) || (key[j] == k && j < tid);
    sorted_key[rank] = k;
    sorted_req[rank] = tid;
  }
  __syncthreads();
  if (tid < bs) {
    // first sorted position of my group: binary search for the first equal key (keys are sorted)
    const int k = sorted_key[tid];
    int lo = 0, hi = tid;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (sorted_key[mid] < k) lo = mid + 1;
      else hi = mid;
    }
    gstart[tid] = lo;
    atomicAdd(&gsize[lo], 1);  // shared-memory atomic
  }
  __syncthreads();
  if (tid < bs) {
    perm[tid] = sorted_req[tid];
    grp_of[sorted_req[tid]] = gstart[tid];
    grp_size[tid] = gsize[tid];
    grp_min[tid] = 0x7fffffff;
  }
}

__global__ void __launch_bounds__(256) lcp_kernel(const int* __restrict__ perm, const int* __restrict__ grp_of,
                                                  int* __restrict__ grp_min, const int32_t* __restrict__ r2t,
                                                  int64_t r2t_stride, const void* __restrict__ req,
                                                  const void* __restrict__ seq, int idx64) {
  __shared__ int s_first;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int g = grp_of[b], leader = perm[g];
  int64_t n = min(load_idx(seq, idx64, b), load_idx(seq, idx64, leader)) - 1;
  if (n < 0) n = 0;
  const int32_t* mine = r2t + load_idx(req, idx64, b) * r2t_stride;
  const int32_t* ref = r2t + load_idx(req, idx64, leader) * r2t_stride;
  if (tid == 0) s_first = static_cast<int>(n);
  __syncthreads();
  if (b != leader) {
    for (int base = 0; base < n; base += 256 * 8) {
      int first = 0x7fffffff;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int t = base + u * 256 + tid;
        if (t < n && __ldg(mine + t) != __ldg(ref + t)) first = min(first, t);
      }
      if (first != 0x7fffffff) atomicMin(&s_first, first);
      __syncthreads();
      if (s_first < base + 256 * 8) break;  // a mismatch inside this block of positions: done
    }
  }
  __syncthreads();
  if (tid == 0) atomicMin(grp_min + g, s_first);
}

// ---------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------
struct CasPlan {  // prefix-group tables of the cascade (null perm = cascade off for this batch)
  const int* perm;
  const int* grp_of;
  const int* grp_min;
  const int* grp_size;
  int* grp_fin;
  int max_tiles;       // capacity of the row-tile table = grid.z of the dense leg
  int* shared32;
  int* ct_first;
  int* ct_cnt;
  int* ct_leader;
  int* ct_shared;
  int rows_per_tile;  // requests per 256-row tile of the dense leg = 256 / g
  int g;
};

__global__ void __launch_bounds__(1024) plan_kernel(int* __restrict__ hdr, int* __restrict__ tile_off,
                                                    int* __restrict__ seq32, int* __restrict__ req32,
                                                    int* __restrict__ cta_b0, int* __restrict__ claim,
                                                    const void* __restrict__ req, const void* __restrict__ seq,
                                                    int idx64, int bs, int tile_tokens, int grid, DynParams dyn,
                                                    const int32_t* __restrict__ r2t, int64_t r2t_stride, const CasPlan cp) {
  __shared__ int warp_sums[32];
  __shared__ int s_running;
  __shared__ int s_runs, s_probed;  // requests whose first tile's slots are / were probed for being consecutive
  pdl_wait();
  pdl_launch_dependents();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ int s_cas_tiles, s_max_shared;
  if (tid == 0) s_running = s_cas_tiles = s_max_shared = s_runs = s_probed = 0;
  __syncthreads();
  // Prefix groups (group_kernel / lcp_kernel ran before us): the part of a group's common prefix that the dense
  // tcgen05 leg takes over; the split-KV tiles of its requests start behind it.
  // cp.grp_fin[group] = what the dense leg takes over for the group: its common prefix rounded down to 128 tokens,
  // 0 when that is short, the group has too few query rows, or the fixed-size tile table is full
  if (cp.perm != nullptr) {
    for (int p0 = tid; p0 < bs; p0 += 1024) {  // sorted positions that start a group: cut it into row tiles
      const int b = cp.perm[p0];
      if (cp.grp_of[b] != p0) continue;
      int sh = cp.grp_min[p0] / kCasTile * kCasTile;
      const int size = cp.grp_size[p0], nt = (size + cp.rows_per_tile - 1) / cp.rows_per_tile;
      if (sh < kMinShared || size * cp.g < kMinCasRows) sh = 0;
      if (sh > 0) {
        const int base = atomicAdd(&s_cas_tiles, nt);
        if (base + nt > cp.max_tiles) {
          atomicSub(&s_cas_tiles, nt);  // no room: this group stays on the split-KV kernel
          sh = 0;
        } else {
          for (int k = 0; k < nt; ++k) {
            cp.ct_first[base + k] = p0 + k * cp.rows_per_tile;
            cp.ct_cnt[base + k] = min(cp.rows_per_tile, size - k * cp.rows_per_tile);
            cp.ct_leader[base + k] = b;
            cp.ct_shared[base + k] = sh;
          }
          atomicMax(&s_max_shared, sh);
        }
      }
      cp.grp_fin[p0] = sh;
    }
  }
  __syncthreads();
  for (int base = 0; base < bs; base += 1024) {
    const int i = base + tid;
    int n = 0;
    if (i < bs) {
      int64_t s = load_idx(seq, idx64, i);
      if (s < 0) s = 0;
      seq32[i] = static_cast<int>(s);
      req32[i] = static_cast<int>(load_idx(req, idx64, i));
      int shared = 0;
      if (cp.perm != nullptr) {
        shared = cp.grp_fin[cp.grp_of[i]];  // <= s - 1 by construction (lcp_kernel)
        cp.shared32[i] = shared;
      }
      n = static_cast<int>((s - shared + tile_tokens - 1) / tile_tokens);
      // Does the allocator hand out runs (a paged pool, or an unfragmented token-granular one)?  Probe the first 16
      // slots behind the shared prefix of every request that has them.  The dynamically claimed tail (make_sched) is
      // only worth its chunk switches and extra partials where whole tiles arrive by TMA: there the CTAs of a launch
      // finish up to 8 % apart, over scattered slots less than 4 %.
      if (r2t != nullptr && s - shared >= 16) {
        const int32_t* row = r2t + static_cast<int64_t>(req32[i]) * r2t_stride + shared;
        bool run = true;
        const int s0 = __ldg(row);
        for (int k = 1; k < 16; ++k) run = run && __ldg(row + k) == s0 + k;
        atomicAdd(&s_probed, 1);
        if (run) atomicAdd(&s_runs, 1);
      }
    }
    int x = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      warp_sums[lane] = w;  // inclusive
    }
    __syncthreads();
    const int running = s_running;
    const int incl = running + x + (warp > 0 ? warp_sums[warp - 1] : 0);
    if (i < bs) tile_off[i] = incl - n;
    __syncthreads();
    if (tid == 1023) s_running = incl;
    __syncthreads();
  }
  {
    // first request of every chunk's tile range: last b with tile_off[b] <= first tile of the chunk (removes a
    // dependent-load binary search from the decode kernel's prologue)
    const Sched sc = make_sched(s_running, grid, runs_dyn(dyn, s_probed, s_runs));
    if (tid < kMaxHeadGroups) claim[tid] = 0;
    for (int c = tid; c < sc.nch; c += 1024) {
      const int gt = sched_chunk_begin(c, grid, sc.s_tiles, sc.d_tiles);
      int lo = 0, hi = bs;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (tile_off[mid] <= gt) lo = mid;  // written above by this CTA, ordered by the __syncthreads
        else hi = mid;
      }
      cta_b0[c] = lo;
    }
  }
  if (tid == 0) {
    const int total = s_running;
    const Sched sc = make_sched(total, grid, runs_dyn(dyn, s_probed, s_runs));
    tile_off[bs] = total;
    hdr[0] = total;
    hdr[1] = sc.s_tiles;
    hdr[2] = sc.nch;
    hdr[9] = sc.d_tiles;
    hdr[10] = sc.dyn0;
    hdr[3] = bs;
    hdr[4] = tile_tokens;
    hdr[5] = grid;
    hdr[6] = kPlanMagic;
    hdr[kHdrShared] = s_max_shared;
    hdr[kHdrCasTiles] = s_cas_tiles;
  }
}

// ---------------------------------------------------------------------------
// decode kernel
// ---------------------------------------------------------------------------
struct DecodeArgs {
  const uint16_t* q;  // bf16 or fp16 bits (act_dtype)
  int64_t q_st, q_sh;
  const uint8_t* kbuf;
  const uint8_t* vbuf;
  int64_t k_st_bytes, v_st_bytes;
  const int32_t* r2t;
  int64_t r2t_stride;
  const int* hdr;
  const int* tile_off;
  const int* seq_len;
  const int* req_idx;
  const int* cta_b0;
  int* claim;  // per head group: next dynamic chunk
  const int* shared32;  // per request: tokens the shared-prefix leg handles (null = cascade off)
  float* part_o;
  float* part_lse;
  int hq, hkv, g;
  int bs;  // what the caller says the batch is; must equal what the plan in the workspace was made for
  float sm_scale, logit_cap;
  const float* k_scale;
  const float* v_scale;
};

// ---- optional per-CTA timeline (build with -DB200_DEC_TRACE; tools/dec_trace.py): %globaltimer at six points
#ifdef B200_DEC_TRACE
__device__ unsigned long long g_dec_trace[1024][6];
#define DEC_TRACE(i)                                                                                   \
  do {                                                                                                 \
    if (threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.x < 1024) {                                    \
      unsigned long long t_;                                                                           \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                           \
      g_dec_trace[blockIdx.x][i] = t_;                                                                 \
    }                                                                                                  \
  } while (0)
#else
#define DEC_TRACE(i) \
  do {               \
  } while (0)
#endif

__device__ unsigned long long g_tma_tiles;  // debug counter: KV tiles fetched by TMA box copies since the last reset

struct Cursor {
  int gt, gt_end;  // global tile index / end of this CTA's range
  int b;           // request (batch row) the tile belongs to
  int b_begin, b_end;  // tile range of request b
  int seq_len;
  const int32_t* row;  // req_to_token row of request b
};

__device__ __forceinline__ void cursor_load_req(Cursor& c, const DecodeArgs& a) {
  c.b_begin = __ldg(a.tile_off + c.b);
  c.b_end = __ldg(a.tile_off + c.b + 1);
  // positions are relative to the end of the prefix the cascade leg handles for this request's group (if any)
  const int shared = a.shared32 ? __ldg(a.shared32 + c.b) : 0;
  c.seq_len = __ldg(a.seq_len + c.b) - shared;
  c.row = a.r2t + static_cast<int64_t>(__ldg(a.req_idx + c.b)) * a.r2t_stride + shared;
}
__device__ __forceinline__ bool cursor_valid(const Cursor& c) { return c.gt < c.gt_end; }
__device__ __forceinline__ void cursor_advance(Cursor& c, const DecodeArgs& a) {
  c.gt++;
  if (c.gt < c.gt_end && c.gt >= c.b_end) {
    do {
      c.b++;
    } while (c.gt >= __ldg(a.tile_off + c.b + 1));
    cursor_load_req(c, a);
  }
}

// QT_ = N=8 tiles of q heads per kv head: 1 for group sizes <= 8, 2 for 9..16 (Llama-3.1-405B: 128 q / 8 kv heads).  The
// QT warps of a kv head read the SAME K/V rows of the staged tile against their own 8 q heads, so a larger group costs
// warps (fewer token sub-slices per tile), never a second pass over the KV bytes.
// HALF_ = the model dtype (q, and the pool unless it is fp8) is fp16 instead of bf16: a 16-bit pool runs the same kernel
// with f16 MMAs and P rounded to f16; an fp8 pool (widened to f16 either way) takes q as it is instead of narrowing it.
template <int D_, bool FP8_, int HPC_, int NW_, bool PAGED_ = false, int QT_ = 1, bool HALF_ = false>
struct DecodeCfg {
  static constexpr int D = D_;
  static constexpr bool FP8 = FP8_;
  static constexpr bool HALF = HALF_;
  static constexpr bool F16MATH = FP8_ || HALF_;  // MMA operand type: f16 (fp8 pools are widened to f16), else bf16
  static constexpr bool PAGED = PAGED_;  // page_size >= 16: whole tiles arrive by TMA box copies (one per page)
  static constexpr int HPC = HPC_;     // kv heads per CTA
  static constexpr int NW = NW_;       // warps per CTA
  static constexpr int NT = NW * 32;
  static constexpr int ELT = FP8 ? 1 : 2;
  static constexpr int SEG = D * ELT;  // bytes of one (token, head) K or V row
  static constexpr int CPS = SEG / 16; // 16-byte chunks per segment
  static constexpr int QT = QT_;
  static constexpr int VH = HPC * QT;    // "virtual" heads per CTA: (kv head, q tile) pairs, one warp each per sub-slice
  static constexpr int NSUB = NW / VH;   // token sub-slices (one warp each per virtual head)
  static constexpr int TN = 16 * NSUB;   // tokens per tile
  static constexpr int CPT = HPC * CPS;  // chunks per token (contiguous in the pool row)
  static constexpr int PASSES = TN * CPT / NT;
  static constexpr int TOK_PER_PASS = NT / CPT;
  static constexpr int TILE_BYTES = HPC * TN * SEG;  // K (or V) bytes per stage
  static constexpr int STAGE_BYTES = 2 * TILE_BYTES;
  static constexpr int ROWB = SEG >= 128 ? 128 : SEG;  // bytes of one shared-memory row = inner extent of a TMA box
  static constexpr int HALVES = SEG / ROWB;            // 256-byte segments are kept as two 128-byte column halves
  // paged: + kStages mbarriers behind the ring.  No alignment slack (the dynamic segment is declared 1024-aligned):
  // two CTAs of 98,368 B (+1 KiB each that the system reserves) stay inside the 196 KiB shared-memory carve-out
  // + 64 bytes behind the ring: the paged kernels' kStages mbarriers, and (every kernel) the word the claimed chunk id
  // is broadcast through
  static constexpr int SMEM_BYTES = kStages * STAGE_BYTES + 64;
  static constexpr int NEXT_CHUNK_OFF = kStages * STAGE_BYTES + 56;
  static_assert(kStages <= 7, "the mbarriers and the chunk word share 64 bytes");
  // the token sub-slices of a tile are combined inside the CTA where the scratch fits a ring stage (everywhere but the
  // fp8 two-q-tile variants): then ONE partial per (cta, request), else one per sub-slice
  static constexpr bool REDUCE = cta_reduce_fits(NSUB, VH, D, STAGE_BYTES);
  static constexpr int PARTS = REDUCE ? 1 : NSUB;
  static constexpr int NSLOT = (TN + 31) / 32;  // slot-id registers per lane
  static constexpr int KC = D / 16;             // MMA k-chunks (bf16) == 16-elt chunks (fp8)
  static constexpr int MT = D / 16;             // PV m-tiles
  static_assert(NW % VH == 0 && NSUB >= 1, "warps must tile the (kv head, q tile) pairs");
  static_assert(NT % CPT == 0 && PASSES >= 1, "copy mapping");
  static_assert(CPS == 4 || CPS == 8 || CPS == 16, "segment size");
};

// Byte offset of 16-byte chunk `chunk` of row (head, tok) inside a K (or V) tile.  Rows are 128 bytes (64 for
// 64-byte segments), XOR-swizzled by 16-byte chunk exactly the way TMA's SWIZZLE_128B (SWIZZLE_64B) mode lays down
// a [tokens x 128 B] box at a 1024-byte-aligned address, so the LDGSTS row gather and the TMA page copy fill ONE
// layout and ldmatrix reads it conflict-free.  A 256-byte segment (bf16, head_dim 128) is two 128-byte column halves.
template <int CPS, int TN>
__device__ __forceinline__ int tile_chunk_off(int head, int tok, int chunk) {
  if constexpr (CPS == 16) return (((head * 2 + (chunk >> 3)) * TN + tok) * 8 + ((chunk & 7) ^ (tok & 7))) * 16;
  else return ((head * TN + tok) * CPS + swz_chunk<CPS>(chunk, tok)) * 16;
}

// x == y ? a : b as an opaque setp + selp: a chain of these over the elements of a small per-thread array keeps the array in
// registers (written as C++ selects the compiler may turn the chain into an indexed load from local memory --
// it did for the 4-register slot-id array of the 128-token tiles, +13-27 % kernel time at the TP8 shards)
__device__ __forceinline__ int select_eq(int a, int b, int x, int y) {  // x == y ? a : b
  int d;
  asm("{\n.reg .pred p;\nsetp.eq.s32 p, %3, %4;\nselp.b32 %0, %1, %2, p;\n}" : "=r"(d) : "r"(a), "r"(b), "r"(x), "r"(y));
  return d;
}

// ---- mbarrier / TMA wrappers (paged pools) -------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  int spins = 0;
  while (!done) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && ++spins > (1 << 22)) __trap();  // a copy that never lands must not hang the GPU
  }
}
// one whole K (V) tile of the pool viewed as (128-byte column block, token rows, column halves, kv heads)
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <class C>
__device__ __forceinline__ void decode_body(const DecodeArgs& a, const CUtensorMap* tmk, const CUtensorMap* tmv) {
  extern __shared__ __align__(1024) uint8_t smem[];  // TMA's swizzle pattern is a function of the address
  constexpr int D = C::D, HPC = C::HPC, TN = C::TN, SEG = C::SEG, CPS = C::CPS, CPT = C::CPT;
  constexpr bool FP8 = C::FP8, PAGED = C::PAGED;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x;
  const int h0 = blockIdx.y * HPC;  // first kv head of this CTA
  DEC_TRACE(0);
  uint32_t smem_base = smem_u32(smem);
  uint32_t bar_base = 0, phases = 0, tma_bits = 0;  // paged: one "tile has landed" mbarrier per stage; its parity bit;
                                                   // bit s of tma_bits: stage s was filled by TMA
  if constexpr (PAGED) {
    if (smem_base & 1023u) __trap();  // the ring must start on a swizzle atom (it does: no static shared memory)
    bar_base = smem_base + kStages * C::STAGE_BYTES;
    if (tid == 0) {
#pragma unroll
      for (int s = 0; s < kStages; ++s) mbar_init(bar_base + 8 * s, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  pdl_wait();               // everything below reads what earlier launches (plan, kv store, qkv GEMM) wrote
  pdl_launch_dependents();  // let the merge grid get scheduled behind us; it waits for our completion
  DEC_TRACE(1);

  // (the schedule words are re-read from the header at every chunk switch -- asm volatile, so that the compiler does not
  // keep five more values alive across the tile loop; the fp8 kernels run at their 128-register cap)
  auto hdr_word = [&](int i) {
    int v;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(a.hdr + i));
    return v;
  };
  const int grid = gridDim.x;
  // Belt and braces behind the host-side B200_ERR_NOT_PLANNED check (a graph replayed without its plan refresh, a
  // workspace overwritten by someone else): without a plan for exactly this batch there is nothing safe to read.
  // The merge kernel makes the same test and writes NaN, so the failure is visible in the output.
  if (__ldg(a.hdr + 6) != kPlanMagic || __ldg(a.hdr + 3) != a.bs || __ldg(a.hdr + 5) != grid) return;
  // chunk x < grid is CTA x's own; further chunks (the dynamic tail of a long launch) are claimed at the bottom
  int chunk = cta;
  if (chunk >= hdr_word(2)) return;
  Cursor pc, cc;  // producer / consumer cursor over the current chunk's tiles
  auto enter_chunk = [&](int c) {
    const int total = hdr_word(0), s_tiles = hdr_word(1), d_tiles = hdr_word(9);
    const int gt_begin = sched_chunk_begin(c, grid, s_tiles, d_tiles);
    pc.gt = gt_begin;
    pc.gt_end = min(gt_begin + (c < grid ? s_tiles : d_tiles), total);
    pc.b = __ldg(a.cta_b0 + c);
    cursor_load_req(pc, a);
    cc = pc;
  };
  enter_chunk(chunk);

  // ---- producer: per-thread constants of the copy mapping -------------------
  const int p_tok = tid / CPT;  // token of pass 0; pass i adds i * TOK_PER_PASS
  const int p_wc = tid % CPT;   // 16-byte chunk inside the token's HPC*SEG span
  const int p_head = p_wc / CPS, p_col = p_wc % CPS;
  const int64_t k_head_off = static_cast<int64_t>(h0) * SEG + p_wc * 16;  // byte offset inside a pool row

  int slot_reg[C::NSLOT];
  auto prefetch_slots = [&](const Cursor& c) {
    const int tok0 = (c.gt - c.b_begin) * TN;
#pragma unroll
    for (int j = 0; j < C::NSLOT; ++j) {
      const int t = tok0 + j * 32 + lane;
      slot_reg[j] = (j * 32 + lane < TN && t < c.seq_len) ? __ldg(c.row + t) : 0;
    }
  };
  // slots of the first `nvalid` tokens of the tile in slot_reg consecutive?  (one shuffle, one compare, one vote)
  auto slots_consecutive = [&](int nvalid) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < C::NSLOT; ++j) {
      int prev = __shfl_up_sync(0xffffffffu, slot_reg[j], 1);
      if (j > 0) {
        const int carry = __shfl_sync(0xffffffffu, slot_reg[j - 1], 31);
        if (lane == 0) prev = carry;
      }
      const int t = j * 32 + lane;
      ok = ok && (t >= nvalid || t == 0 || slot_reg[j] == prev + 1);
    }
    return __all_sync(0xffffffffu, ok);
  };
  int n_tma_tiles = 0;  // tiles this CTA fetched by TMA (debug counter, summed into g_tma_tiles at exit)
  auto issue_tile = [&](auto tma_tag, const Cursor& c, int stage) {
    constexpr bool TMA = decltype(tma_tag)::value;
    const int tok0 = (c.gt - c.b_begin) * TN;
    const uint32_t sk = smem_base + stage * C::STAGE_BYTES;
    const uint32_t sv = sk + C::TILE_BYTES;
    if constexpr (TMA) {
      // A tile that lies wholly inside the sequence and whose TN slots are consecutive (pages of >= TN tokens from a
      // paged allocator: slot = page * page_size + offset, request position t at offset t % page_size; or simply an
      // unfragmented token-granular allocator) is fetched by TMA: one box for K, one for V, issued by thread 0,
      // signalling the stage's mbarrier.  The check is made HERE, on the slot ids every warp holds anyway (one
      // shuffle, one vote), so the caller's page-size hint can cost speed but never correctness; any other tile --
      // the ragged last tile of a request included -- takes the row gather below (which zero-fills rows past the end;
      // the PV product relies on that).  The decision is remembered per stage for the consumer side.
      const bool by_tma = tok0 + TN <= c.seq_len && slots_consecutive(TN);
      const bool prev_by_tma = (tma_bits >> stage) & 1u;  // how this ring stage was filled three tiles ago
      tma_bits = by_tma ? (tma_bits | (1u << stage)) : (tma_bits & ~(1u << stage));
      if (by_tma) {
        if (tid == 0) {
          const uint32_t bar = bar_base + 8 * stage;
          // a stage last written by the row gather (generic proxy) changes hands to the async proxy: fence once
          // (MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC); TMA after TMA needs none -- the reads in between were ordered by
          // the barrier at the top of the tile loop
          if (!prev_by_tma) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive_expect_tx(bar, C::STAGE_BYTES);
          // the whole K (V) tile is ONE box (all heads of the CTA, both column halves): two copies per tile
          tma_load_4d(sk, tmk, bar, 0, slot_reg[0], 0, h0);  // thread 0 holds the slot id of the tile's first token
          tma_load_4d(sv, tmv, bar, 0, slot_reg[0], 0, h0);
          ++n_tma_tiles;
        }
        return;
      }
    }
#pragma unroll
    for (int i = 0; i < C::PASSES; ++i) {
      const int tok = p_tok + i * C::TOK_PER_PASS;
      // which slot-id register holds token tok = p_tok + i * TOK_PER_PASS (p_tok < TOK_PER_PASS, powers of two):
      // known at compile time when a pass spans <= 32 tokens, else one of TOK_PER_PASS / 32 consecutive registers
      int sreg;
      if constexpr (C::TOK_PER_PASS <= 32) {
        sreg = slot_reg[(i * C::TOK_PER_PASS) >> 5];
      } else {
        constexpr int R = C::TOK_PER_PASS / 32;
        sreg = slot_reg[i * R];
#pragma unroll
        for (int r = 1; r < R; ++r) sreg = select_eq(slot_reg[i * R + r], sreg, p_tok >> 5, r);
      }
      const int slot = __shfl_sync(0xffffffffu, sreg, tok & 31);
      const int nbytes = (tok0 + tok < c.seq_len) ? 16 : 0;
      const uint32_t doff = tile_chunk_off<CPS, TN>(p_head, tok, p_col);
      cp_async16(sk + doff, kv_row_ptr(a.kbuf, slot, a.k_st_bytes) + k_head_off, nbytes);
      cp_async16(sv + doff, kv_row_ptr(a.vbuf, slot, a.v_st_bytes) + k_head_off, nbytes);
    }
  };

  // (Round-1 experiment, removed: prefetch.global.L2 / cp.async.bulk.prefetch.L2 of the rows 2-16
  // tiles ahead of the cp.async front, meant to keep more bytes in flight than shared memory allows.
  // Measured slower at every distance -- C2 169 -> 187 us at 2 tiles, 295 us at 8 -- so it is gone.)

  // ---- consumer: warp -> (kv head, token sub-slice) ---------------------------
  const int w_vh = warp % C::VH;       // (kv head, q tile) pair of this warp
  const int w_head = w_vh / C::QT;     // kv head inside the CTA's staged tile
  const int q0 = (w_vh % C::QT) * 8;   // first q head (of the kv head's group) of this warp's N=8 tile
  const int w_sub = warp / C::VH;
  const int kvh = h0 + w_head;
  const int gid = lane >> 2, cid = lane & 3;  // mma "groupID" / "threadID_in_group"
  const int G = a.g;
  const float ks = a.k_scale ? __ldg(a.k_scale + kvh) : 1.0f;
  const float vs = a.v_scale ? __ldg(a.v_scale + kvh) : 1.0f;
  const float cap = a.logit_cap;
  const float sc_raw = a.sm_scale * ks;           // logits = s * sc_raw
  const float sc_log2 = sc_raw * kLog2e;          // exp2 domain

  uint32_t qf[C::KC][2];
  float m[2], l[2];
  float o[C::MT][4];
  auto reset_state = [&]() {
    m[0] = m[1] = -INFINITY;
    l[0] = l[1] = 0.f;
#pragma unroll
    for (int i = 0; i < C::MT; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  };
  // B fragments of Q^T: lane holds q-head n = gid (of this kv head's group)
  auto load_q = [&](int b) {
    const bool live = q0 + gid < G;
    const uint16_t* qp = a.q + static_cast<int64_t>(b) * a.q_st + static_cast<int64_t>(kvh * G + (live ? q0 + gid : 0)) * a.q_sh;
#pragma unroll
    for (int kc = 0; kc < C::KC; ++kc) {
      if constexpr (!FP8) {
        // slots (2c,2c+1) <-> d = 16kc+2c+{0,1};  (2c+8,2c+9) <-> d+8
        const uint32_t* p = reinterpret_cast<const uint32_t*>(qp + kc * 16 + cid * 2);
        qf[kc][0] = live ? __ldg(p) : 0u;
        qf[kc][1] = live ? __ldg(p + 4) : 0u;
        // Pass the fragments through a shuffle so the loop's HMMAs read SHFL results, not LDG
        // results: otherwise ptxas ties them to the scoreboard slot the per-tile slot-id LDG
        // also uses and every tile's first HMMA waits a DRAM latency (ncu r1a: 22% long_scoreboard).
        qf[kc][0] = __shfl_sync(0xffffffffu, qf[kc][0], lane);
        qf[kc][1] = __shfl_sync(0xffffffffu, qf[kc][1], lane);
      } else {
        // 16-element chunk kc: slots (2c,2c+1) <-> d = 16kc+4c+{0,1}; (2c+8,2c+9) <-> d = 16kc+4c+{2,3}
        const uint2 raw = live ? __ldg(reinterpret_cast<const uint2*>(qp + kc * 16 + cid * 4)) : make_uint2(0u, 0u);
        if constexpr (C::HALF) {  // fp16 activations: already the MMA's operand type (through a shuffle, as above:
          qf[kc][0] = __shfl_sync(0xffffffffu, raw.x, lane);  // the loop's HMMAs must not read an LDG result directly)
          qf[kc][1] = __shfl_sync(0xffffffffu, raw.y, lane);
        } else {
          const __nv_bfloat162 x01 = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
          const __nv_bfloat162 x23 = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
          qf[kc][0] = pack_f16x2(__low2float(x01), __high2float(x01));
          qf[kc][1] = pack_f16x2(__low2float(x23), __high2float(x23));
        }
      }
    }
  };

  auto compute_tile = [&](int stage, int nvalid) {
    // nvalid: valid tokens in this warp's 16-token slice (>= 1)
    const uint32_t sk = smem_base + stage * C::STAGE_BYTES;
    const uint32_t sv = sk + C::TILE_BYTES;
    const int wtok = w_sub * 16;  // this warp's 16-token slice of the tile
    const int lrow = lane & 15;   // ldmatrix row (token) this lane addresses
    const int lhi = lane >> 4;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    if constexpr (!FP8) {
#pragma unroll
      for (int kc = 0; kc < C::KC; ++kc) {
        uint32_t af[4];
        ldmatrix_x4(af, sk + tile_chunk_off<CPS, TN>(w_head, wtok + lrow, 2 * kc + lhi));
        mma_16816<C::F16MATH>(s, af, qf[kc][0], qf[kc][1]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < C::KC / 2; ++j) {
        uint32_t r[4];  // r0:(tok 0-7, chunk 2j) r1:(tok 8-15, 2j) r2:(tok 0-7, 2j+1) r3:(tok 8-15, 2j+1)
        ldmatrix_x4(r, sk + tile_chunk_off<CPS, TN>(w_head, wtok + lrow, 2 * j + lhi));
        uint32_t af[4];
        af[0] = e4m3x2_to_f16x2(static_cast<uint16_t>(r[0] & 0xffffu));
        af[1] = e4m3x2_to_f16x2(static_cast<uint16_t>(r[1] & 0xffffu));
        af[2] = e4m3x2_to_f16x2(static_cast<uint16_t>(r[0] >> 16));
        af[3] = e4m3x2_to_f16x2(static_cast<uint16_t>(r[1] >> 16));
        mma_16816<true>(s, af, qf[2 * j][0], qf[2 * j][1]);
        af[0] = e4m3x2_to_f16x2(static_cast<uint16_t>(r[2] & 0xffffu));
        af[1] = e4m3x2_to_f16x2(static_cast<uint16_t>(r[3] & 0xffffu));
        af[2] = e4m3x2_to_f16x2(static_cast<uint16_t>(r[2] >> 16));
        af[3] = e4m3x2_to_f16x2(static_cast<uint16_t>(r[3] >> 16));
        mma_16816<true>(s, af, qf[2 * j + 1][0], qf[2 * j + 1][1]);
      }
    }
    // s[0],s[1]: token gid, heads 2cid,2cid+1; s[2],s[3]: token gid+8
    if (cap > 0.f) {
#pragma unroll
      for (int i = 0; i < 4; ++i) s[i] = softcap(s[i] * sc_raw, cap) * kLog2e;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) s[i] *= sc_log2;
    }
    if (gid >= nvalid) s[0] = s[1] = -INFINITY;
    if (gid + 8 >= nvalid) s[2] = s[3] = -INFINITY;
    float mx0 = fmaxf(s[0], s[2]), mx1 = fmaxf(s[1], s[3]);
#pragma unroll
    for (int x = 4; x < 32; x <<= 1) {
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, x));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, x));
    }
    const float mn0 = fmaxf(m[0], mx0), mn1 = fmaxf(m[1], mx1);  // finite: nvalid >= 1
    const float al0 = ex2(m[0] - mn0), al1 = ex2(m[1] - mn1);
    m[0] = mn0;
    m[1] = mn1;
    const float p0 = ex2(s[0] - mn0), p1 = ex2(s[1] - mn1), p2 = ex2(s[2] - mn0), p3 = ex2(s[3] - mn1);
    l[0] = l[0] * al0 + p0 + p2;  // per-lane partial sums, reduced over lanes at flush
    l[1] = l[1] * al1 + p1 + p3;
#pragma unroll
    for (int i = 0; i < C::MT; ++i) {
      o[i][0] *= al0;
      o[i][1] *= al1;
      o[i][2] *= al0;
      o[i][3] *= al1;
    }
    // P^T (tokens x heads, accumulator layout) -> B fragment (k = tokens, n = heads)
    const uint32_t pb0 = movmatrix_trans(pack2<C::F16MATH>(p0, p1));
    const uint32_t pb1 = movmatrix_trans(pack2<C::F16MATH>(p2, p3));
    const int vtok = (lane & 7) + ((lane >> 4) << 3);  // ldmatrix.trans row (token)
    const int vhalf = (lane >> 3) & 1;
    if constexpr (!FP8) {
#pragma unroll
      for (int mt = 0; mt < C::MT; ++mt) {
        uint32_t af[4];  // V^T tile: rows d = 16mt + {gid, gid+8}, cols tokens
        ldmatrix_x4_trans(af, sv + tile_chunk_off<CPS, TN>(w_head, wtok + vtok, 2 * mt + vhalf));
        mma_16816<C::F16MATH>(o[mt], af, pb0, pb1);
      }
    } else {
#pragma unroll
      for (int j = 0; j < C::MT / 2; ++j) {
        uint32_t r[4];  // transposed: lane holds bytes {V[2c][d0],V[2c][d0+1],V[2c+1][d0],V[2c+1][d0+1]}, d0 = 16ch+2gid
        ldmatrix_x4_trans(r, sv + tile_chunk_off<CPS, TN>(w_head, wtok + lrow, 2 * j + lhi));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t t0 = prmt(r[2 * h], 0u, 0x3120u);      // tokens 0-7 : {d0:(2c,2c+1) | d0+1:(2c,2c+1)}
          const uint32_t t1 = prmt(r[2 * h + 1], 0u, 0x3120u);  // tokens 8-15
          uint32_t af[4];
          af[0] = e4m3x2_to_f16x2(static_cast<uint16_t>(t0 & 0xffffu));  // row gid   <-> d0,   k 2c..
          af[1] = e4m3x2_to_f16x2(static_cast<uint16_t>(t0 >> 16));      // row gid+8 <-> d0+1
          af[2] = e4m3x2_to_f16x2(static_cast<uint16_t>(t1 & 0xffffu));  // k 2c+8..
          af[3] = e4m3x2_to_f16x2(static_cast<uint16_t>(t1 >> 16));
          mma_16816<true>(o[2 * j + h], af, pb0, pb1);
        }
      }
    }
  };

  // one partial per (chunk, request): slot = chunk + b is unique along the monotone staircase of (chunk, request)
  // pairs.  (A fused "last CTA to arrive merges" variant was measured in round 1 and was slower -- 184 vs 169 us
  // at C2: the fence + atomic round trip + merge reads stall the copy pipeline at every request boundary -- so the
  // merge stays a separate, programmatically-dependent launch.)
  // With several token sub-slices per tile (NSUB > 1: the one- and two-head CTAs of tensor-parallel shards) the
  // NSUB softmax states of a head are first combined inside the CTA: a tree over the sub-slice index through the
  // ring stage that has just been consumed (free until the next iteration's copy is issued into it).  Partner warps
  // hold the same (head, dim) fragment in the same lane, so the exchange is lane-local: no shuffles, log2(NSUB)
  // rounds, and the workspace and the merge kernel see ONE partial per (cta, request) instead of NSUB.
  auto flush = [&](int b, int free_stage) {
    if constexpr (C::REDUCE) {
      constexpr int kStateF = C::MT * 4 + 4;  // floats of softmax state per lane: o, m[2], l[2]
      float* red = reinterpret_cast<float*>(smem + free_stage * C::STAGE_BYTES);
#pragma unroll
      for (int half = C::NSUB / 2; half >= 1; half >>= 1) {
        __syncthreads();  // round 1: every warp is done reading the stage; later rounds: the previous round's reads
        if (w_sub >= half && w_sub < 2 * half) {
          float* dst = red + (static_cast<size_t>((w_sub - half) * C::VH + w_vh) * kStateF) * 32 + lane;
#pragma unroll
          for (int mt = 0; mt < C::MT; ++mt)
#pragma unroll
            for (int e = 0; e < 4; ++e) dst[(mt * 4 + e) * 32] = o[mt][e];
          dst[(C::MT * 4 + 0) * 32] = m[0];
          dst[(C::MT * 4 + 1) * 32] = m[1];
          dst[(C::MT * 4 + 2) * 32] = l[0];
          dst[(C::MT * 4 + 3) * 32] = l[1];
        }
        __syncthreads();
        if (w_sub < half) {
          const float* src = red + (static_cast<size_t>(w_sub * C::VH + w_vh) * kStateF) * 32 + lane;
          const float pm0 = src[(C::MT * 4 + 0) * 32], pm1 = src[(C::MT * 4 + 1) * 32];
          const float mn0 = fmaxf(m[0], pm0), mn1 = fmaxf(m[1], pm1);
          // a slice that saw no token has m = -inf (and l = 0, o = 0): weight 0, never exp2(-inf - -inf)
          const float a0 = m[0] == -INFINITY ? 0.f : ex2(m[0] - mn0), a1 = m[1] == -INFINITY ? 0.f : ex2(m[1] - mn1);
          const float b0 = pm0 == -INFINITY ? 0.f : ex2(pm0 - mn0), b1 = pm1 == -INFINITY ? 0.f : ex2(pm1 - mn1);
#pragma unroll
          for (int mt = 0; mt < C::MT; ++mt) {
            o[mt][0] = o[mt][0] * a0 + src[(mt * 4 + 0) * 32] * b0;
            o[mt][1] = o[mt][1] * a1 + src[(mt * 4 + 1) * 32] * b1;
            o[mt][2] = o[mt][2] * a0 + src[(mt * 4 + 2) * 32] * b0;
            o[mt][3] = o[mt][3] * a1 + src[(mt * 4 + 3) * 32] * b1;
          }
          l[0] = l[0] * a0 + src[(C::MT * 4 + 2) * 32] * b0;
          l[1] = l[1] * a1 + src[(C::MT * 4 + 3) * 32] * b1;
          m[0] = mn0;
          m[1] = mn1;
        }
      }
      if constexpr (PAGED) tma_bits &= ~(1u << free_stage);  // generic-proxy writes: the next TMA fill of this stage fences
      if (w_sub != 0) return;
    }
    float l0 = l[0], l1 = l[1];
#pragma unroll
    for (int x = 4; x < 32; x <<= 1) {
      l0 += __shfl_xor_sync(0xffffffffu, l0, x);
      l1 += __shfl_xor_sync(0xffffffffu, l1, x);
    }
    const float inv0 = l0 > 0.f ? vs / l0 : 0.f, inv1 = l1 > 0.f ? vs / l1 : 0.f;
    const int64_t slot = C::REDUCE ? static_cast<int64_t>(chunk + b) : static_cast<int64_t>(chunk + b) * C::NSUB + w_sub;
    const int n0 = q0 + 2 * cid, n1 = n0 + 1;  // q-head index inside the group
    const int64_t pi0 = slot * a.hq + kvh * G + n0;
    const int64_t pi1 = pi0 + 1;
    if (gid == 0) {
      if (n0 < G) a.part_lse[pi0] = l0 > 0.f ? m[0] + lg2(l0) : -INFINITY;
      if (n1 < G) a.part_lse[pi1] = l1 > 0.f ? m[1] + lg2(l1) : -INFINITY;
    }
#pragma unroll
    for (int mt = 0; mt < C::MT; ++mt) {
      const int da = FP8 ? 16 * mt + 2 * gid : 16 * mt + gid;
      const int db = FP8 ? da + 1 : da + 8;
      if (n0 < G) {
        a.part_o[pi0 * D + da] = o[mt][0] * inv0;
        a.part_o[pi0 * D + db] = o[mt][2] * inv0;
      }
      if (n1 < G) {
        a.part_o[pi1 * D + da] = o[mt][1] * inv1;
        a.part_o[pi1 * D + db] = o[mt][3] * inv1;
      }
    }
  };

  // ---- pipeline ---------------------------------------------------------------
  // The tile loop exists in two copies inside the TMA-capable kernel: one that can fetch a tile either way (TMA box
  // copy when its slots are consecutive, row gather otherwise) and one that only gathers rows -- the very loop of
  // decode_mma_kernel.  A row gather inside the two-producer loop runs 6 % below its speed in a loop of its own
  // (the second producer's code in the loop body changes how the gather is scheduled), so a CTA picks
  // ONE copy for its whole tile range from what it sees in its first tile: consecutive slots (a paged pool, or an
  // unfragmented token-granular allocator) -> the TMA-capable loop; scattered slots -> the pure row-gather loop, at
  // the plain kernel's speed.  Either loop reads any table correctly; the choice is about speed only.
  auto pipeline = [&](auto tma_tag) {
  constexpr bool TMA = decltype(tma_tag)::value;
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (cursor_valid(pc)) {
      issue_tile(tma_tag, pc, s);
      cursor_advance(pc, a);
      if (cursor_valid(pc)) prefetch_slots(pc);
    }
    cp_async_commit();
  }
  reset_state();
  load_q(cc.b);
  DEC_TRACE(2);

  int stage = 0;
#ifdef B200_DEC_TRACE
  bool first_tile = true;
#endif
  // Nested loops: the inner loop streams the tiles of ONE request inside this CTA's range and contains nothing but
  // wait -> issue -> compute; the partial flush (with its in-CTA reduction) and the switch to the next request sit
  // between two inner loops.  (-DB200_FLAT_LOOP keeps the round-1 single loop with the flush inside, for A/B.)
#ifndef B200_FLAT_LOOP
  while (cursor_valid(cc)) {
    const int seg_end = min(cc.b_end, cc.gt_end);
    int tok0 = (cc.gt - cc.b_begin) * TN + w_sub * 16;
    int last_stage = stage;
    for (; cc.gt < seg_end; ++cc.gt, tok0 += TN) {
      cp_async_wait<kStages - 2>();
      if constexpr (TMA) {
        if ((tma_bits >> stage) & 1u) {  // the producer side fetched this stage's tile by TMA
          mbar_wait(bar_base + 8 * stage, (phases >> stage) & 1u);
          phases ^= 1u << stage;
        }
      }
      __syncthreads();
#ifdef B200_DEC_TRACE
      if (first_tile) {
        DEC_TRACE(3);
        first_tile = false;
      }
#endif
      {
        int pstage = stage + kStages - 1;
        if (pstage >= kStages) pstage -= kStages;
        if (cursor_valid(pc)) {
          issue_tile(tma_tag, pc, pstage);
          cursor_advance(pc, a);
          if (cursor_valid(pc)) prefetch_slots(pc);
        }
        cp_async_commit();
      }
      const int nvalid = min(16, cc.seq_len - tok0);
      if (nvalid > 0) compute_tile(stage, nvalid);
      last_stage = stage;
      stage = stage + 1 == kStages ? 0 : stage + 1;
    }
    flush(cc.b, last_stage);
    if (cc.gt < cc.gt_end) {  // next request with tiles in this CTA's range (requests without tiles are skipped)
      do {
        cc.b++;
      } while (cc.gt >= __ldg(a.tile_off + cc.b + 1));
      cursor_load_req(cc, a);
      reset_state();
      load_q(cc.b);
    }
  }
#else
  while (cursor_valid(cc)) {
    cp_async_wait<kStages - 2>();
    if constexpr (TMA) {
      if ((tma_bits >> stage) & 1u) {  // the producer side fetched this stage's tile by TMA
        mbar_wait(bar_base + 8 * stage, (phases >> stage) & 1u);
        phases ^= 1u << stage;
      }
    }
    __syncthreads();
    {
      int pstage = stage + kStages - 1;
      if (pstage >= kStages) pstage -= kStages;
      if (cursor_valid(pc)) {
        issue_tile(tma_tag, pc, pstage);
        cursor_advance(pc, a);
        if (cursor_valid(pc)) prefetch_slots(pc);
      }
      cp_async_commit();
    }
    const int tok0 = (cc.gt - cc.b_begin) * TN + w_sub * 16;
    const int nvalid = min(16, cc.seq_len - tok0);
    if (nvalid > 0) compute_tile(stage, nvalid);

    const int b_cur = cc.b;
    const bool req_done = (cc.gt + 1 >= cc.gt_end) || (cc.gt + 1 >= cc.b_end);
    cursor_advance(cc, a);
    if (req_done) {
      flush(b_cur, stage);
      if (cursor_valid(cc)) {
        reset_state();
        load_q(cc.b);
      }
    }
    stage = stage + 1 == kStages ? 0 : stage + 1;
  }
#endif
  };  // pipeline

  volatile int* next_chunk = reinterpret_cast<volatile int*>(smem + C::NEXT_CHUNK_OFF);
  for (;;) {
    prefetch_slots(pc);
    if constexpr (PAGED) {
      // consecutive slots among the valid tokens of this chunk's first tile?  (a single valid token says nothing)
      const int nvalid0 = min(TN, pc.seq_len - (pc.gt - pc.b_begin) * TN);
      if (!kTwoLoops || (nvalid0 >= 2 && slots_consecutive(nvalid0))) pipeline(std::true_type{});
      else pipeline(std::false_type{});
    } else {
      pipeline(std::false_type{});
    }
    const int nch = hdr_word(2);
    if (nch <= grid) break;  // nothing dynamic in this launch: one chunk per CTA
    // claim the next chunk of the dynamic tail (one atomic per chunk and head group); the barrier also keeps the next
    // chunk's first copies out of ring stages a slower warp is still reading
    if (tid == 0) *next_chunk = grid + atomicAdd(a.claim + blockIdx.y, 1);
    __syncthreads();
    chunk = *next_chunk;
    if (chunk >= nch) break;
    enter_chunk(chunk);
  }
  if constexpr (PAGED) {
    if (tid == 0 && n_tma_tiles) atomicAdd(&g_tma_tiles, static_cast<unsigned long long>(n_tma_tiles));
  }
  DEC_TRACE(4);
  cp_async_wait<0>();
  DEC_TRACE(5);
}

template <class C>
__global__ void __launch_bounds__(C::NT, kMinBlocks) decode_mma_kernel(const DecodeArgs a) {
  decode_body<C>(a, nullptr, nullptr);
}
// paged pools: the two tensor maps (K and V pool of this layer) ride in the kernel parameters
template <class C>
__global__ void __launch_bounds__(C::NT, kMinBlocks) decode_paged_kernel(const __grid_constant__ DecodeArgs a,
                                                                const __grid_constant__ CUtensorMap tmk,
                                                                const __grid_constant__ CUtensorMap tmv) {
  decode_body<C>(a, &tmk, &tmv);
}

// ---------------------------------------------------------------------------
// merge: o[b, h, :] = sum_i w_i * O_i,  w_i = 2^(lse_i - max) / sum_j 2^(lse_j - max)
// One warp per (request, q head): lanes own DV/32 consecutive dims; the <= few partial
// slots of the pair are consecutive ((c + b) * nsub + sub), still L2-resident.
// ---------------------------------------------------------------------------
template <int DV>
#ifdef B200_MERGE_DEEP  // A/B: twice the loads in flight per warp at 3 instead of 4 CTAs per SM
#define B200_MERGE_MINBLOCKS 3
#else
#define B200_MERGE_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(256, B200_MERGE_MINBLOCKS) merge_kernel(const int* __restrict__ hdr, const int* __restrict__ tile_off,
                                                    const float* __restrict__ part_o,
                                                    const float* __restrict__ part_lse, uint16_t* __restrict__ o, int o_half,
                                                    int64_t o_st, int64_t o_sh, int bs, int hq, int nsub,
                                                    const float* __restrict__ cas_o, const float* __restrict__ cas_lse,
                                                    int cas_splits, const int* __restrict__ shared32, int* __restrict__ claim) {
  pdl_wait();
  pdl_launch_dependents();
  // the split-KV grid has completed: its claim counters start the next launch (next layer, same plan) at zero
  if (blockIdx.x == 0 && threadIdx.x < kMaxHeadGroups) claim[threadIdx.x] = 0;
  constexpr int VPL = DV / 32;  // floats per lane
  const int lane = threadIdx.x & 31;
  const int pair = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (pair >= bs * hq) return;
  const int b = pair / hq, h = pair % hq;
  const int s_tiles = __ldg(hdr + 1), d_tiles = __ldg(hdr + 9), dyn0 = __ldg(hdr + 10), grid = __ldg(hdr + 5);
  if (__ldg(hdr + 6) != kPlanMagic || __ldg(hdr + 3) != bs) {  // no plan for this batch in the workspace: poison o
    uint16_t* bad = o + static_cast<int64_t>(b) * o_st + static_cast<int64_t>(h) * o_sh + lane * (DV / 32);
#pragma unroll
    for (int j = 0; j < DV / 32; ++j) bad[j] = o_half ? 0x7e00 : 0x7fc0;  // quiet NaN in fp16 / bf16
    return;
  }
  const int t0 = __ldg(tile_off + b), t1 = __ldg(tile_off + b + 1);
  // partials of the shared-prefix leg (one per KV split), present when the plan found a shared prefix
  const int ncas = (cas_splits > 0 && __ldg(shared32 + b) > 0) ? cas_splits : 0;
  uint16_t* op = o + static_cast<int64_t>(b) * o_st + static_cast<int64_t>(h) * o_sh + lane * VPL;
  float acc[VPL];
#pragma unroll
  for (int j = 0; j < VPL; ++j) acc[j] = 0.f;
  float den = 0.f;
  int64_t first = 0;
  int n = 0;
  if (t1 > t0) {
    const int c0 = sched_chunk_of(t0, dyn0, grid, s_tiles, d_tiles), c1 = sched_chunk_of(t1 - 1, dyn0, grid, s_tiles, d_tiles);
    first = static_cast<int64_t>(c0 + b) * nsub;
    n = (c1 - c0 + 1) * nsub;
  }
  float mx = -INFINITY;
  for (int i = lane; i < n; i += 32) mx = fmaxf(mx, __ldcg(part_lse + (first + i) * hq + h));
  const int64_t cas_i = static_cast<int64_t>(b) * hq + h, cas_st = static_cast<int64_t>(bs) * hq;
  if (lane < ncas) mx = fmaxf(mx, __ldcg(cas_lse + lane * cas_st + cas_i));
#pragma unroll
  for (int x = 16; x > 0; x >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, x));
  if (mx > -INFINITY) {
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      const float w_lane = i < n ? ex2(__ldcg(part_lse + (first + i) * hq + h) - mx) : 0.f;
      const int cnt = min(32, n - base);
      // several partials per round trip: all loads of a group are issued before the first FMA consumes one (with one
      // warp per (request, head) and few warps per SM at small shards, a load-use chain per partial was the
      // whole kernel: 16-24 partials x L2 latency)
#ifdef B200_MERGE_DEEP
      constexpr int kGroup = 32 / VPL;        // 32 registers: the requests in a dynamically claimed tail have ~18 partials
#else
      constexpr int kGroup = 32 / (VPL * 2);  // 16 registers of loads in flight: 8 partials (DV 64) or 4 (DV 128)
#endif
      for (int j0 = 0; j0 < cnt; j0 += kGroup) {
        float vbuf[kGroup][VPL];
#pragma unroll
        for (int u = 0; u < kGroup; ++u) {
          if (j0 + u < cnt) {
            const float* src = part_o + ((first + base + j0 + u) * hq + h) * DV + lane * VPL;
            if constexpr (VPL == 4) {
              const float4 v = __ldcg(reinterpret_cast<const float4*>(src));
              vbuf[u][0] = v.x; vbuf[u][1] = v.y; vbuf[u][2] = v.z; vbuf[u][3] = v.w;
            } else {
              const float2 v = __ldcg(reinterpret_cast<const float2*>(src));
              vbuf[u][0] = v.x; vbuf[u][1] = v.y;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kGroup; ++u) {
          if (j0 + u < cnt) {
            const float w = __shfl_sync(0xffffffffu, w_lane, j0 + u);
            den += w;
#pragma unroll
            for (int e = 0; e < VPL; ++e) acc[e] += w * vbuf[u][e];
          }
        }
      }
    }
    if (ncas > 0) {
      // the shared-prefix leg's partials, four loads in flight at a time (one dependent L2 round trip per partial was
      // 4-8 of them in a row at the end of every warp)
      const float w_lane = lane < ncas ? ex2(__ldcg(cas_lse + lane * cas_st + cas_i) - mx) : 0.f;
      for (int j0 = 0; j0 < ncas; j0 += 4) {
        float vb4[4][VPL];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (j0 + u < ncas) {
            const float* src = cas_o + ((j0 + u) * cas_st + cas_i) * DV + lane * VPL;
            if constexpr (VPL == 4) {
              const float4 v = __ldcg(reinterpret_cast<const float4*>(src));
              vb4[u][0] = v.x; vb4[u][1] = v.y; vb4[u][2] = v.z; vb4[u][3] = v.w;
            } else {
              const float2 v = __ldcg(reinterpret_cast<const float2*>(src));
              vb4[u][0] = v.x; vb4[u][1] = v.y;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (j0 + u < ncas) {
            const float w = __shfl_sync(0xffffffffu, w_lane, j0 + u);
            den += w;
#pragma unroll
            for (int e = 0; e < VPL; ++e) acc[e] += w * vb4[u][e];
          }
        }
      }
    }
  }
  const float r = den > 0.f ? 1.f / den : 0.f;  // seq_len == 0 -> zeros
  auto pack_o = [&](float lo, float hi) { return o_half ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi); };
  if constexpr (VPL == 4) {
    *reinterpret_cast<uint2*>(op) = make_uint2(pack_o(acc[0] * r, acc[1] * r), pack_o(acc[2] * r, acc[3] * r));
  } else {
    *reinterpret_cast<uint32_t*>(op) = pack_o(acc[0] * r, acc[1] * r);
  }
}

// ---------------------------------------------------------------------------
// debug / test hooks
// ---------------------------------------------------------------------------
__global__ void gather_kernel(const uint8_t* __restrict__ buf, int64_t stride_t_bytes, int row_bytes,
                              const int32_t* __restrict__ r2t, int64_t r2t_stride, const void* __restrict__ req,
                              const void* __restrict__ seq, int idx64, const int64_t* __restrict__ out_off,
                              uint8_t* __restrict__ out) {
  const int b = blockIdx.y;
  const int64_t n = load_idx(seq, idx64, b);
  const int32_t* row = r2t + load_idx(req, idx64, b) * r2t_stride;
  for (int64_t t = blockIdx.x; t < n; t += gridDim.x) {
    const uint8_t* src = kv_row_ptr(buf, __ldg(row + t), stride_t_bytes);
    uint8_t* dst = out + (out_off[b] + t) * row_bytes;
    for (int i = threadIdx.x * 16; i < row_bytes; i += blockDim.x * 16)
      *reinterpret_cast<uint4*>(dst + i) = __ldg(reinterpret_cast<const uint4*>(src + i));
  }
}

// paged-table contract check: counts positions t < seq_len whose slot is not page_start + t % page_size with
// page_start = slot of position t - t % page_size a multiple of page_size
__global__ void check_paged_kernel(const int32_t* __restrict__ r2t, int64_t r2t_stride, const void* __restrict__ req,
                                   const void* __restrict__ seq, int idx64, int page_size, int* __restrict__ violations) {
  const int b = blockIdx.x;
  const int64_t n = load_idx(seq, idx64, b);
  const int32_t* row = r2t + load_idx(req, idx64, b) * r2t_stride;
  int bad = 0;
  for (int64_t t = threadIdx.x; t < n; t += blockDim.x) {
    const int off = static_cast<int>(t % page_size);
    const int32_t first = row[t - off];
    bad += (first % page_size != 0) || (row[t] != first + off);
  }
  if (bad) atomicAdd(violations, bad);
}

// one warp per (request, q head); fp32 everywhere; the obviously-correct twin
__global__ void __launch_bounds__(32) naive_decode_kernel(const b200_decode_params p, const void* __restrict__ req,
                                                          const void* __restrict__ seq, int idx64) {
  const int b = blockIdx.x, h = blockIdx.y, lane = threadIdx.x;
  const int g = p.hq / p.hkv, kvh = h / g;
  const int64_t n = load_idx(seq, idx64, b);
  const int32_t* row = p.req_to_token + load_idx(req, idx64, b) * p.req_to_token_stride;
  const bool fp8 = p.kv_dtype == B200_KV_FP8_E4M3, kv_half = p.kv_dtype == B200_KV_F16, act_half = p.act_dtype == B200_ACT_F16;
  const int elt = fp8 ? 1 : 2;
  const float ks = p.k_scale ? p.k_scale[kvh] : 1.f, vs = p.v_scale ? p.v_scale[kvh] : 1.f;
  const uint16_t* q = static_cast<const uint16_t*>(p.q) + b * p.q_stride_t + h * p.q_stride_h;
  auto ld16 = [](const uint16_t* x, int i, bool half) { return half ? f16_bits_to_float(x[i]) : bf16_bits_to_float(x[i]); };
  float acc[8];  // dv <= 256
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  float m = -INFINITY, l = 0.f;
  for (int64_t t = 0; t < n; ++t) {
    const int32_t slot = row[t];
    const uint8_t* kr = kv_row_ptr(static_cast<const uint8_t*>(p.k_buf), slot, p.k_stride_t * elt) + kvh * p.d * elt;
    const uint8_t* vr = kv_row_ptr(static_cast<const uint8_t*>(p.v_buf), slot, p.v_stride_t * elt) + kvh * p.dv * elt;
    float s = 0.f;
    for (int i = lane; i < p.d; i += 32) {
      const float kx = fp8 ? e4m3_to_float(kr[i]) * ks : ld16(reinterpret_cast<const uint16_t*>(kr), i, kv_half);
      s += ld16(q, i, act_half) * kx;
    }
#pragma unroll
    for (int x = 16; x > 0; x >>= 1) s += __shfl_xor_sync(0xffffffffu, s, x);
    s *= p.sm_scale;
    if (p.logit_cap > 0.f) s = p.logit_cap * tanhf(s / p.logit_cap);
    const float mn = fmaxf(m, s);
    const float al = expf(m - mn), pr = expf(s - mn);
    l = l * al + pr;
    for (int i = lane, j = 0; i < p.dv; i += 32, ++j) {
      const float vx = fp8 ? e4m3_to_float(vr[i]) * vs : ld16(reinterpret_cast<const uint16_t*>(vr), i, kv_half);
      acc[j] = acc[j] * al + pr * vx;
    }
    m = mn;
  }
  uint16_t* op = static_cast<uint16_t*>(p.o) + b * p.o_stride_t + h * p.o_stride_h;
  for (int i = lane, j = 0; i < p.dv; i += 32, ++j) {
    const float x = l > 0.f ? acc[j] / l : 0.f;
    op[i] = act_half ? __half_as_ushort(__float2half_rn(x)) : __bfloat16_as_ushort(__float2bfloat16_rn(x));
  }
}

// ---------------------------------------------------------------------------
// host side: config selection + launchers
// ---------------------------------------------------------------------------
struct KernelChoice {
  int d, fp8, hpc, nw, qt, half;  // half: fp16 activations (and pool, unless it is fp8)
};
// partials per (cta, request) slot of the workspace: must mirror DecodeCfg::PARTS
static int parts_per_slot(const KernelChoice& kc) {
  const int vh = kc.hpc * kc.qt, nsub = kc.nw / vh, elt = kc.fp8 ? 1 : 2;
  const int stage_bytes = 2 * kc.hpc * (16 * nsub) * kc.d * elt;
  return cta_reduce_fits(nsub, vh, kc.d, stage_bytes) ? 1 : nsub;
}

static int pick_hpc(int hkv, int nw) {
  int hpc = 1;
  while (hpc * 2 <= nw && hkv % (hpc * 2) == 0) hpc *= 2;
  return hpc;
}

static int choose(int hq, int hkv, int d, int dv, int kv_dtype, KernelChoice* kc) {
  if (hq <= 0 || hkv <= 0 || hq % hkv != 0) return B200_ERR_INVALID;
  if (d != dv) return B200_ERR_UNSUPPORTED;            // MLA-style split dims: out of scope
  if (d != 64 && d != 128) return B200_ERR_UNSUPPORTED;
  if (kv_dtype != B200_KV_BF16 && kv_dtype != B200_KV_FP8_E4M3 && kv_dtype != B200_KV_F16) return B200_ERR_UNSUPPORTED;
  // one N=8 mma tile of q heads per kv head, or two for group sizes 9..16 (built for head_dim 128)
  if (hq / hkv > 16 || (hq / hkv > 8 && d != 128)) return B200_ERR_UNSUPPORTED;
  kc->d = d;
  kc->fp8 = kv_dtype == B200_KV_FP8_E4M3;
  kc->half = kv_dtype == B200_KV_F16;
  kc->nw = (d == 128 && !kc->fp8) ? 4 : 8;  // keep a stage at 32 KiB
  kc->qt = hq / hkv > 8 ? 2 : 1;
  kc->hpc = pick_hpc(hkv, kc->nw / kc->qt);
  return B200_OK;
}

struct DeviceInfo {
  int num_sms = 0;
};
static int device_info(DeviceInfo* di) {
  static DeviceInfo cached[kMaxDevices];
  int dev = 0;
  if (int rc = current_device(&dev)) return rc;
  if (cached[dev].num_sms == 0)
    B200_CUDA_TRY(cudaDeviceGetAttribute(&cached[dev].num_sms, cudaDevAttrMultiProcessorCount, dev));
  *di = cached[dev];
  return B200_OK;
}

// Which workspaces hold a plan, and for what: b200_attn_plan records it, b200_attn_decode refuses a workspace that
// was never planned, or was planned for another batch size / shape / cascade setting (B200_ERR_NOT_PLANNED).  Host
// bookkeeping only -- no sync; the kernels re-check the plan header on device.
struct PlanRecord {
  const void* ws = nullptr;
  int bs = 0, hq = 0, hkv = 0, d = 0, kv_dtype = 0, cascade = 0, max_bs = 0;
};
constexpr int kPlanSlots = 64;
static PlanRecord g_plans[kPlanSlots];
static int g_plan_next = 0;
static std::mutex g_plan_mu;
static void plan_record(const PlanRecord& r) {
  std::lock_guard<std::mutex> lock(g_plan_mu);
  for (int i = 0; i < kPlanSlots; ++i)
    if (g_plans[i].ws == r.ws) {
      g_plans[i] = r;
      return;
    }
  g_plans[g_plan_next] = r;
  g_plan_next = (g_plan_next + 1) % kPlanSlots;
}
static bool plan_matches(const PlanRecord& r) {
  std::lock_guard<std::mutex> lock(g_plan_mu);
  for (int i = 0; i < kPlanSlots; ++i) {
    const PlanRecord& x = g_plans[i];
    if (x.ws == r.ws)
      return x.bs == r.bs && x.hq == r.hq && x.hkv == r.hkv && x.d == r.d && x.kv_dtype == r.kv_dtype &&
             x.cascade == r.cascade && x.max_bs == r.max_bs;
  }
  return false;
}

static int g_dyn_min_base = kDynMinBase;  // b200_debug_set_dyn_min_base: tests force the dynamic tail onto small batches
// B200_SCHED="frac_den,chunk_den" (measurement switch): the last 1/frac_den of the tiles in chunks of 1/chunk_den share;
// chunk_den <= 4 * frac_den keeps the chunk count inside the workspace (max_chunks)
static DynParams dyn_params() {
  static int frac_den = 8, chunk_den = 32;
  static bool parsed = false, off = false;  // "0,0": equal static ranges only
  if (!parsed) {
    parsed = true;
    if (const char* e = getenv("B200_SCHED")) {
      int a = 0, b = 0;
      if (sscanf(e, "%d,%d", &a, &b) == 2 && a >= 2 && b >= a && b <= kDynChunksPerCta * a) {
        frac_den = a;
        chunk_den = b;
      } else if (a == 0) {
        off = true;
      }
    }
  }
  return DynParams{off ? (1 << 30) : g_dyn_min_base, frac_den, chunk_den};
}
static int g_last_tma_mode = 0;  // test / bench introspection: 0 row gather only, 1 TMA for tiles of consecutive slots

template <class C>
static int launch_decode_t(const DecodeArgs& a, const PagedMaps* maps, int grid_x, int grid_y, cudaStream_t st, bool query_only,
                           int* occ_out) {
  static PerDevice pd;  // per instantiation: shared-memory attribute and occupancy, remembered per device
  auto prepare = [&](auto kern) -> int {
    if (int rc = ensure_smem_attr(kern, C::SMEM_BYTES, pd)) return rc;
    if (query_only) {
      int dev = 0;
      if (int rc = current_device(&dev)) return rc;
      if (pd.occupancy[dev] == 0) {
        int occ = 0;
        B200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::NT, C::SMEM_BYTES));
        pd.occupancy[dev] = occ > 0 ? occ : -1;
      }
      *occ_out = pd.occupancy[dev] > 0 ? pd.occupancy[dev] : 0;
    }
    return B200_OK;
  };
  if constexpr (C::PAGED) {
    auto kern = decode_paged_kernel<C>;
    if (int rc = prepare(kern)) return rc;
    if (query_only) return B200_OK;
    return launch_pdl(kern, dim3(grid_x, grid_y), dim3(C::NT), C::SMEM_BYTES, st, a, maps->k, maps->v);
  } else {
    auto kern = decode_mma_kernel<C>;
    if (int rc = prepare(kern)) return rc;
    if (query_only) return B200_OK;
    return launch_pdl(kern, dim3(grid_x, grid_y), dim3(C::NT), C::SMEM_BYTES, st, a);
  }
}

template <int D, bool FP8, bool PAGED, bool HALF = false>
static int dispatch_hpc(const KernelChoice& kc, const DecodeArgs& a, const PagedMaps* m, int gx, int gy, cudaStream_t st, bool q,
                        int* occ) {
  constexpr int NW = (D == 128 && !FP8) ? 4 : 8;
  if constexpr (D == 128) {
    if (kc.qt == 2) {  // group sizes 9..16: two q tiles per kv head
      switch (kc.hpc) {
        case 1: return launch_decode_t<DecodeCfg<D, FP8, 1, NW, PAGED, 2, HALF>>(a, m, gx, gy, st, q, occ);
        case 2: return launch_decode_t<DecodeCfg<D, FP8, 2, NW, PAGED, 2, HALF>>(a, m, gx, gy, st, q, occ);
        case 4:
          if constexpr (NW >= 8) return launch_decode_t<DecodeCfg<D, FP8, 4, NW, PAGED, 2, HALF>>(a, m, gx, gy, st, q, occ);
      }
      return B200_ERR_UNSUPPORTED;
    }
  }
  switch (kc.hpc) {
    case 1: return launch_decode_t<DecodeCfg<D, FP8, 1, NW, PAGED, 1, HALF>>(a, m, gx, gy, st, q, occ);
    case 2: return launch_decode_t<DecodeCfg<D, FP8, 2, NW, PAGED, 1, HALF>>(a, m, gx, gy, st, q, occ);
    case 4: return launch_decode_t<DecodeCfg<D, FP8, 4, NW, PAGED, 1, HALF>>(a, m, gx, gy, st, q, occ);
    case 8:
      if constexpr (NW >= 8) return launch_decode_t<DecodeCfg<D, FP8, 8, NW, PAGED, 1, HALF>>(a, m, gx, gy, st, q, occ);
  }
  return B200_ERR_UNSUPPORTED;
}

template <bool PAGED>
static int dispatch_p(const KernelChoice& kc, const DecodeArgs& a, const PagedMaps* m, int gx, int gy, cudaStream_t st, bool q,
                      int* occ) {
  if (kc.half) {
    if (kc.d == 64)
      return kc.fp8 ? dispatch_hpc<64, true, PAGED, true>(kc, a, m, gx, gy, st, q, occ) : dispatch_hpc<64, false, PAGED, true>(kc, a, m, gx, gy, st, q, occ);
    return kc.fp8 ? dispatch_hpc<128, true, PAGED, true>(kc, a, m, gx, gy, st, q, occ) : dispatch_hpc<128, false, PAGED, true>(kc, a, m, gx, gy, st, q, occ);
  }
  if (kc.d == 64)
    return kc.fp8 ? dispatch_hpc<64, true, PAGED>(kc, a, m, gx, gy, st, q, occ) : dispatch_hpc<64, false, PAGED>(kc, a, m, gx, gy, st, q, occ);
  return kc.fp8 ? dispatch_hpc<128, true, PAGED>(kc, a, m, gx, gy, st, q, occ) : dispatch_hpc<128, false, PAGED>(kc, a, m, gx, gy, st, q, occ);
}

// maps == nullptr: page size 1, row gather only
static int dispatch(const KernelChoice& kc, const DecodeArgs& a, const PagedMaps* maps, int gx, int gy, cudaStream_t st, bool q,
                    int* occ) {
  return maps ? dispatch_p<true>(kc, a, maps, gx, gy, st, q, occ) : dispatch_p<false>(kc, a, nullptr, gx, gy, st, q, occ);
}

static int get_config(int max_bs, int hq, int hkv, int d, int dv, int kv_dtype, int force_path, b200_attn_config* out) {
  (void)force_path;
  if (!out || max_bs <= 0) return B200_ERR_INVALID;
  KernelChoice kc;
  int rc = choose(hq, hkv, d, dv, kv_dtype, &kc);
  if (rc) return rc;
  DeviceInfo di;
  rc = device_info(&di);
  if (rc) return rc;
  int occ = 0;
  DecodeArgs dummy{};
  rc = dispatch(kc, dummy, nullptr, 0, 0, nullptr, true, &occ);
  if (rc) return rc;
  if (occ < 1) return B200_ERR_UNSUPPORTED;
  if (occ > 2) occ = 2;
  const int elt = kc.fp8 ? 1 : 2;
  out->path = B200_PATH_MMA;
  out->heads_per_cta = kc.hpc;
  out->sub_splits = kc.nw / (kc.hpc * kc.qt);
  out->partials_per_slot = parts_per_slot(kc);
  out->q_tiles = kc.qt;
  out->tile_tokens = 16 * out->sub_splits;
  const int gy = hkv / kc.hpc;
  if (gy > kMaxHeadGroups) return B200_ERR_UNSUPPORTED;  // one claim counter per head group
  // persistent grid: every resident CTA slot of the machine, shared by the head groups
  int gx = di.num_sms * occ / gy;
  if (gx < 1) gx = 1;
  out->grid_ctas = gx;
  out->stages = kStages;
  out->smem_bytes = kStages * 2 * kc.hpc * out->tile_tokens * d * elt + 64;
  out->num_sms = di.num_sms;
  out->workspace_bytes = ws_layout(max_bs, gx, out->partials_per_slot, hq, dv).total;
  return B200_OK;
}

int launch_shared_prefix_tc(const b200_decode_params* p, const CasTables& t, float* cas_o, float* cas_lse, int nsplit,
                            cudaStream_t st);  // extend_tc.cu

}  // namespace b200

using namespace b200;

extern "C" {

int b200_attn_abi_version(void) { return B200_ATTN_ABI_VERSION; }

uint64_t b200_attn_struct_size(int which) {
  switch (which) {
    case 0: return sizeof(b200_attn_config);
    case 1: return sizeof(b200_decode_params);
    case 2: return sizeof(b200_extend_params);
    case 3: return sizeof(b200_rope_store_params);
    default: return 0;
  }
}

const char* b200_attn_error_string(int code) {
  switch (code) {
    case B200_OK: return "ok";
    case B200_ERR_INVALID: return "invalid argument (null pointer, bad size or misaligned stride)";
    case B200_ERR_UNSUPPORTED: return "unsupported configuration (head_dim must be 64 or 128 with d == dv, group size <= 8 -- <= 16 at head_dim 128 --, bf16 activations over a bf16 or fp8-e4m3 pool or fp16 activations over an fp16 or fp8-e4m3 pool)";
    case B200_ERR_WORKSPACE: return "workspace too small for this configuration";
    case B200_ERR_NOT_PLANNED: return "workspace holds no plan for this batch (call b200_attn_plan with the same bs, shape and cascade flag first)";
    default:
      if (code >= 1000) return cudaGetErrorString(static_cast<cudaError_t>(code - 1000));
      return "unknown error";
  }
}

int b200_attn_get_config(int32_t max_bs, int32_t hq, int32_t hkv, int32_t d, int32_t dv, int32_t kv_dtype,
                         int32_t force_path, b200_attn_config* out) {
  return get_config(max_bs, hq, hkv, d, dv, kv_dtype, force_path, out);
}

// Row-tile capacity of the dense leg (its grid is fixed at launch, the plan fills the table on device): one group
// covering the batch needs ceil(bs / requests-per-tile); every further prefix group can add one partial tile.
static int cas_max_tiles(int bs, int g) {
  const int rows_per_tile = 256 / g;
  int n = (bs + rows_per_tile - 1) / rows_per_tile + kExtraCasTiles;
  return n > bs ? bs : n;
}

// cascade is honoured only where the dense leg exists and is worth launching
static bool cascade_supported(int bs, int hq, int hkv, int d, int dv, int kv_dtype) {
  const bool shape_ok = d == dv && (d == 128 || d == 64);
  return shape_ok && hkv > 0 && 256 % (hq / hkv) == 0 && bs * (hq / hkv) >= kMinCasRows;
}

int b200_attn_plan(void* workspace, uint64_t workspace_bytes, int32_t max_bs, const void* req_pool_indices,
                   const void* seq_lens, int32_t idx_dtype, int32_t bs, int32_t hq, int32_t hkv, int32_t d, int32_t dv,
                   int32_t kv_dtype, int32_t force_path, const int32_t* req_to_token, int64_t req_to_token_stride,
                   int32_t cascade, b200_stream_t stream) {
  if (!workspace || !req_pool_indices || !seq_lens || bs <= 0 || bs > max_bs) return B200_ERR_INVALID;
  if (cascade && !req_to_token) return B200_ERR_INVALID;
  b200_attn_config cfg;
  int rc = get_config(max_bs, hq, hkv, d, dv, kv_dtype, force_path, &cfg);
  if (rc) return rc;
  if (workspace_bytes < cfg.workspace_bytes) return B200_ERR_WORKSPACE;
  const WsLayout L = ws_layout(max_bs, cfg.grid_ctas, cfg.partials_per_slot, hq, dv);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int use_cascade = cascade && bs <= kMaxGroupBs && cascade_supported(bs, hq, hkv, d, dv, kv_dtype);
  auto wi = [&](size_t off) { return reinterpret_cast<int*>(ws + off); };
  CasPlan cp{};
  if (use_cascade) {
    const bool i64 = idx_dtype == B200_IDX_I64;
    group_kernel<<<1, 1024, 0, st>>>(wi(L.perm), wi(L.grp_of), wi(L.grp_min), wi(L.grp_size), req_to_token, req_to_token_stride,
                                     req_pool_indices, seq_lens, i64, bs);
    lcp_kernel<<<bs, 256, 0, st>>>(wi(L.perm), wi(L.grp_of), wi(L.grp_min), req_to_token, req_to_token_stride, req_pool_indices,
                                   seq_lens, i64);
    B200_CUDA_TRY(cudaGetLastError());
    cp.perm = wi(L.perm);
    cp.grp_of = wi(L.grp_of);
    cp.grp_min = wi(L.grp_min);
    cp.grp_size = wi(L.grp_size);
    cp.grp_fin = wi(L.grp_fin);
    cp.max_tiles = cas_max_tiles(bs, hq / hkv);
    cp.shared32 = wi(L.shared32);
    cp.ct_first = wi(L.ct_first);
    cp.ct_cnt = wi(L.ct_cnt);
    cp.ct_leader = wi(L.ct_leader);
    cp.ct_shared = wi(L.ct_shared);
    cp.g = hq / hkv;
    cp.rows_per_tile = 256 / cp.g;
  }
  plan_kernel<<<1, 1024, 0, st>>>(
      reinterpret_cast<int*>(ws), reinterpret_cast<int*>(ws + L.tile_off), reinterpret_cast<int*>(ws + L.seq_len),
      reinterpret_cast<int*>(ws + L.req_idx), reinterpret_cast<int*>(ws + L.cta_b0), reinterpret_cast<int*>(ws + L.claim), req_pool_indices, seq_lens, idx_dtype == B200_IDX_I64, bs,
      cfg.tile_tokens, cfg.grid_ctas, dyn_params(), req_to_token, req_to_token_stride, cp);
  B200_CUDA_TRY(cudaGetLastError());
  PlanRecord rec;
  rec.ws = workspace; rec.bs = bs; rec.hq = hq; rec.hkv = hkv; rec.d = d; rec.kv_dtype = kv_dtype;
  rec.cascade = cascade != 0; rec.max_bs = max_bs;
  plan_record(rec);
  return B200_OK;
}

int b200_attn_decode(const b200_decode_params* p, b200_stream_t stream) {
  if (!p || !p->q || !p->k_buf || !p->v_buf || !p->o || !p->req_to_token || !p->workspace) return B200_ERR_INVALID;
  if (p->bs <= 0 || p->bs > p->max_bs) return B200_ERR_INVALID;
  if (!dtype_pair_ok(p->kv_dtype, p->act_dtype)) return B200_ERR_UNSUPPORTED;
  KernelChoice kc;
  int rc = choose(p->hq, p->hkv, p->d, p->dv, p->kv_dtype, &kc);
  if (rc) return rc;
  kc.half = p->act_dtype == B200_ACT_F16;  // fp8 pools: the q dtype picks the kernel (same shape, same launch configuration)
  const int elt = kc.fp8 ? 1 : 2;
  // 16-byte vector access on pool rows, 8-byte on q
  if ((p->k_stride_t * elt) % 16 || (p->v_stride_t * elt) % 16 || p->q_stride_t % 4 || p->q_stride_h % 4 ||
      (reinterpret_cast<uintptr_t>(p->k_buf) & 15) || (reinterpret_cast<uintptr_t>(p->v_buf) & 15) ||
      (reinterpret_cast<uintptr_t>(p->q) & 7) || (reinterpret_cast<uintptr_t>(p->o) & 15) || p->o_stride_t % 8 ||
      p->o_stride_h % 8)
    return B200_ERR_INVALID;
  b200_attn_config cfg;
  rc = get_config(p->max_bs, p->hq, p->hkv, p->d, p->dv, p->kv_dtype, p->force_path, &cfg);
  if (rc) return rc;
  if (p->workspace_bytes < cfg.workspace_bytes) return B200_ERR_WORKSPACE;
  {
    PlanRecord rec;
    rec.ws = p->workspace; rec.bs = p->bs; rec.hq = p->hq; rec.hkv = p->hkv; rec.d = p->d; rec.kv_dtype = p->kv_dtype;
    rec.cascade = p->cascade != 0; rec.max_bs = p->max_bs;
    if (!plan_matches(rec)) return B200_ERR_NOT_PLANNED;
  }
  const WsLayout L = ws_layout(p->max_bs, cfg.grid_ctas, cfg.partials_per_slot, p->hq, p->dv);
  uint8_t* ws = static_cast<uint8_t*>(p->workspace);
  DecodeArgs a;
  a.q = static_cast<const uint16_t*>(p->q);
  a.q_st = p->q_stride_t;
  a.q_sh = p->q_stride_h;
  a.kbuf = static_cast<const uint8_t*>(p->k_buf);
  a.vbuf = static_cast<const uint8_t*>(p->v_buf);
  a.k_st_bytes = p->k_stride_t * elt;
  a.v_st_bytes = p->v_stride_t * elt;
  a.r2t = p->req_to_token;
  a.r2t_stride = p->req_to_token_stride;
  a.hdr = reinterpret_cast<const int*>(ws);
  a.tile_off = reinterpret_cast<const int*>(ws + L.tile_off);
  a.seq_len = reinterpret_cast<const int*>(ws + L.seq_len);
  a.req_idx = reinterpret_cast<const int*>(ws + L.req_idx);
  a.cta_b0 = reinterpret_cast<const int*>(ws + L.cta_b0);
  a.claim = reinterpret_cast<int*>(ws + L.claim);
  a.part_lse = reinterpret_cast<float*>(ws + L.part_lse);
  a.part_o = reinterpret_cast<float*>(ws + L.part_o);
  a.hq = p->hq;
  a.hkv = p->hkv;
  a.g = p->hq / p->hkv;
  a.bs = p->bs;
  a.sm_scale = p->sm_scale;
  a.logit_cap = p->logit_cap;
  a.k_scale = p->k_scale;
  a.v_scale = p->v_scale;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the optional event pair brackets both attention legs (shared-prefix + split-KV), not the merge
  if (p->ev_kernel_begin) B200_CUDA_TRY(cudaEventRecord(static_cast<cudaEvent_t>(p->ev_kernel_begin), st));
  // shared-prefix leg first (it exits at once when the plan found nothing shared); must match what plan decided
  int cas_splits = 0;
  float* cas_o = reinterpret_cast<float*>(ws + L.cas_o);
  float* cas_lse = reinterpret_cast<float*>(ws + L.cas_lse);
  a.shared32 = nullptr;
  if (p->cascade && p->bs <= kMaxGroupBs && cascade_supported(p->bs, p->hq, p->hkv, p->d, p->dv, p->kv_dtype)) {
    const int g = p->hq / p->hkv, rows_per_tile = 256 / g, row_blocks = (p->bs + rows_per_tile - 1) / rows_per_tile;
    cas_splits = cfg.num_sms / (row_blocks * p->hkv);
    cas_splits = cas_splits < 1 ? 1 : (cas_splits > kMaxCasSplits ? kMaxCasSplits : cas_splits);
    auto wi = [&](size_t off) { return reinterpret_cast<const int*>(ws + off); };
    CasTables t;
    t.hdr = a.hdr;
    t.req32 = a.req_idx;
    t.perm = wi(L.perm);
    t.ct_first = wi(L.ct_first);
    t.ct_cnt = wi(L.ct_cnt);
    t.ct_leader = wi(L.ct_leader);
    t.ct_shared = wi(L.ct_shared);
    t.max_tiles = cas_max_tiles(p->bs, g);
    a.shared32 = wi(L.shared32);
    rc = launch_shared_prefix_tc(p, t, cas_o, cas_lse, cas_splits, st);
    if (rc) return rc;
  }
  // page_size > 1: slots of a page are consecutive and request position t sits at offset
  // t % page_size of its page, so a tile inside a page is fetched with one TMA box copy for K and one for V (the kernel checks
  // every tile's slot ids before it trusts that: a wrong page_size costs speed, not correctness)
  PagedMaps maps;
  const PagedMaps* mp = nullptr;
  if (p->page_size > 1 || p->page_size == B200_PAGE_AUTO) {
    // B200_PAGE_AUTO: a token-granular pool whose allocator often hands out consecutive slots
    const int ps = p->page_size == B200_PAGE_AUTO ? cfg.tile_tokens : p->page_size;
    if (ps < 16 || ps > 128 || (ps & (ps - 1)) || p->kv_pool_rows <= 0) return B200_ERR_UNSUPPORTED;
    // Measured: one box per K (V) tile (16 KB) beats the row gather by 2-5 %; page runs smaller than the
    // tile, fetched as one box per head and run (2-8 KB), never did -- per-copy cost of the TMA unit -- while the
    // row gather over a paged table gains the pages' DRAM locality by itself.  So: TMA when a page covers the tile,
    // the row gather (correct on any table: one slot id per token) otherwise.
    if (ps >= cfg.tile_tokens) {
      rc = cached_pool_map(&maps.k, p->k_buf, a.k_st_bytes, p->kv_pool_rows, p->hkv, p->d, kc.fp8, cfg.tile_tokens, kc.hpc);
      if (!rc) rc = cached_pool_map(&maps.v, p->v_buf, a.v_st_bytes, p->kv_pool_rows, p->hkv, p->dv, kc.fp8, cfg.tile_tokens, kc.hpc);
      if (!rc) mp = &maps;
      else if (rc != B200_ERR_INVALID) return rc;  // a driver that refuses the map: row gather
    }
  }
  g_last_tma_mode = mp ? 1 : 0;
  rc = dispatch(kc, a, mp, cfg.grid_ctas, p->hkv / kc.hpc, st, false, nullptr);
  if (rc) return rc;
  if (p->ev_kernel_end) B200_CUDA_TRY(cudaEventRecord(static_cast<cudaEvent_t>(p->ev_kernel_end), st));
  uint16_t* o = static_cast<uint16_t*>(p->o);
  const int o_half = p->act_dtype == B200_ACT_F16;
  const dim3 mgrid((p->bs * p->hq + 7) / 8);
  if (p->dv == 64)
    return launch_pdl(merge_kernel<64>, mgrid, dim3(256), 0, st, a.hdr, a.tile_off, static_cast<const float*>(a.part_o),
                      static_cast<const float*>(a.part_lse), o, o_half, static_cast<int64_t>(p->o_stride_t),
                      static_cast<int64_t>(p->o_stride_h), static_cast<int>(p->bs), static_cast<int>(p->hq),
                      static_cast<int>(cfg.partials_per_slot), static_cast<const float*>(cas_o), static_cast<const float*>(cas_lse), cas_splits, reinterpret_cast<const int*>(ws + L.shared32), a.claim);
  return launch_pdl(merge_kernel<128>, mgrid, dim3(256), 0, st, a.hdr, a.tile_off, static_cast<const float*>(a.part_o),
                    static_cast<const float*>(a.part_lse), o, o_half, static_cast<int64_t>(p->o_stride_t),
                    static_cast<int64_t>(p->o_stride_h), static_cast<int>(p->bs), static_cast<int>(p->hq),
                    static_cast<int>(cfg.partials_per_slot), static_cast<const float*>(cas_o), static_cast<const float*>(cas_lse), cas_splits, reinterpret_cast<const int*>(ws + L.shared32), a.claim);
}

#ifdef B200_DEC_TRACE
int b200_debug_dec_trace_read(unsigned long long* host_dst) {
  return cudaMemcpyFromSymbol(host_dst, g_dec_trace, sizeof(g_dec_trace)) == cudaSuccess ? 0 : 1;
}
#endif

int b200_debug_last_tma_mode(void) { return g_last_tma_mode; }

// Test hook (host only, no device needed): the tile schedule b200_attn_plan would write for `total` tiles on a persistent
// grid of `grid` CTAs when the slots are runs -- out[0..5] = tiles per static chunk, tiles per dynamic chunk, first dynamic
// tile, number of chunks, capacity of the workspace's chunk tables, and, for the tile `probe`, the chunk that owns it;
// out[6], out[7] = first tile of that chunk and of the next one.
int b200_debug_sched(int32_t total, int32_t grid, int32_t min_base, int32_t probe, int32_t* out) {
  if (!out || total < 0 || grid <= 0) return B200_ERR_INVALID;
  DynParams dyn = dyn_params();
  if (min_base > 0) dyn.min_base = min_base;
  const Sched sc = make_sched(total, grid, dyn);
  out[0] = sc.s_tiles;
  out[1] = sc.d_tiles;
  out[2] = sc.dyn0;
  out[3] = sc.nch;
  out[4] = max_chunks(grid);
  const int c = (probe >= 0 && probe < total) ? sched_chunk_of(probe, sc.dyn0, grid, sc.s_tiles, sc.d_tiles) : -1;
  out[5] = c;
  out[6] = c >= 0 ? sched_chunk_begin(c, grid, sc.s_tiles, sc.d_tiles) : -1;
  out[7] = c >= 0 ? sched_chunk_begin(c + 1, grid, sc.s_tiles, sc.d_tiles) : -1;
  return B200_OK;
}

// Test hook: tiles per CTA from which b200_attn_plan cuts the last eighth of the tiles into dynamically claimed chunks
// (<= 0 restores the default, kDynMinBase); returns the previous value.  Plans made afterwards use it.
int b200_debug_set_dyn_min_base(int32_t tiles) {
  const int prev = g_dyn_min_base;
  g_dyn_min_base = tiles > 0 ? tiles : kDynMinBase;
  return prev;
}

// Test hook (synchronises): number of KV tiles the decode kernels fetched by TMA box copies on the current device since
// the last call with reset != 0; proves that the TMA branch is taken, not merely that the TMA-capable kernel ran.
int64_t b200_debug_tma_tiles(int32_t reset) {
  unsigned long long n = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(&n, g_tma_tiles, sizeof(n)) != cudaSuccess) return -1;
  if (reset) {
    const unsigned long long zero = 0;
    if (cudaMemcpyToSymbol(g_tma_tiles, &zero, sizeof(zero)) != cudaSuccess) return -1;
  }
  return static_cast<int64_t>(n);
}

int b200_debug_check_paged(const int32_t* req_to_token, int64_t req_to_token_stride, const void* req_pool_indices,
                           const void* seq_lens, int32_t idx_dtype, int32_t bs, int32_t page_size, int32_t* violations,
                           b200_stream_t stream) {
  if (!req_to_token || !req_pool_indices || !seq_lens || !violations || bs <= 0 || page_size <= 0) return B200_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B200_CUDA_TRY(cudaMemsetAsync(violations, 0, sizeof(int32_t), st));
  check_paged_kernel<<<bs, 256, 0, st>>>(req_to_token, req_to_token_stride, req_pool_indices, seq_lens,
                                         idx_dtype == B200_IDX_I64, page_size, violations);
  B200_CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

int b200_debug_gather(const void* buf, int64_t stride_t_bytes, int32_t row_bytes, const int32_t* req_to_token,
                      int64_t req_to_token_stride, const void* req_pool_indices, const void* seq_lens,
                      int32_t idx_dtype, const int64_t* out_offsets, int32_t bs, void* out, b200_stream_t stream) {
  if (!buf || !req_to_token || !req_pool_indices || !seq_lens || !out_offsets || !out || bs <= 0) return B200_ERR_INVALID;
  if (row_bytes % 16 || stride_t_bytes % 16) return B200_ERR_INVALID;
  gather_kernel<<<dim3(64, bs), 64, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(buf), stride_t_bytes, row_bytes, req_to_token, req_to_token_stride, req_pool_indices,
      seq_lens, idx_dtype == B200_IDX_I64, out_offsets, static_cast<uint8_t*>(out));
  B200_CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

int b200_debug_decode_naive(const b200_decode_params* p, const void* req_pool_indices, const void* seq_lens,
                            int32_t idx_dtype, b200_stream_t stream) {
  if (!p || !req_pool_indices || !seq_lens || p->bs <= 0 || p->dv > 256) return B200_ERR_INVALID;
  naive_decode_kernel<<<dim3(p->bs, p->hq), 32, 0, static_cast<cudaStream_t>(stream)>>>(*p, req_pool_indices, seq_lens,
                                                                                        idx_dtype == B200_IDX_I64);
  B200_CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

}  // extern "C"
