/*
 * b200_attn.h -- C ABI of the B200-native paged attention kernels.
 *
 * This is the drop-in boundary for eth-easl/Scratchpad's attention hot path:
 * the entry points below are what a new `cuda_b200` AttentionBackend
 * (scratchpad/nn/attention/backend.py:11-105) binds through ctypes.  Plain
 * pointers, sizes and a raw cudaStream_t only; no torch types.  Every call
 *   - returns 0 on success, a B200_ERR_* code or 1000+cudaError_t otherwise,
 *   - never throws, never synchronises the host, never allocates device memory,
 *   - launches on the stream it is given with grids that do not depend on
 *     device-side values (capture-safe for CUDA graphs,
 *     scratchpad/model_executor/cuda_graph_runner.py:394-423,483-510).
 *
 * Data layout contract (all paths relative to /root/reference):
 *   req_to_token  int32 [rows, cols] row-major, row stride given in elements
 *                 (scratchpad/memory/pool.py:16-22)
 *   k_buf/v_buf   [P+1, Hkv, D] contiguous in the last two dims, token stride
 *                 given in elements; bf16 / fp16 (the model dtype), or fp8-e4m3 stored as bytes
 *                 (pool.py:298-317,373-390); slot 0 is the padding slot
 *   q / o         bf16 or fp16 (act_dtype) [T, Hq, D] with last dim contiguous, token and head
 *                 strides in elements (q may be a column slice of the fused
 *                 qkv GEMM output, scratchpad/nn/models/llama/llama.py:149-152)
 *   req_pool_indices, seq_lens  int32 or int64 (eager batches are int64,
 *                 schedule_batch.py:916-924; graph buffers int32,
 *                 cuda_graph_runner.py:194-197) -- see idx_dtype fields
 *   page size is 1 in the reference (model_runner.py:431-432): the gather unit is one token row.  Decode also
 *                 accepts page sizes 16..128 (b200_decode_params.page_size): req_to_token keeps one slot id per
 *                 token, slots of a page are consecutive, and the gather unit becomes a page.
 */
#ifndef B200_ATTN_H_
#define B200_ATTN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ATTN_ABI_VERSION 10

/* kv_dtype */
#define B200_KV_BF16 0
#define B200_KV_FP8_E4M3 1
#define B200_KV_F16 2
/* act_dtype: the model dtype -- q, the new tokens' k / v and o (model_runner.dtype; `o` is "same dtype as q",
 * triton_backend.py:171-176).  Pools that go with it: bf16 activations over a bf16 or fp8-e4m3 pool, fp16 activations
 * over an fp16 or fp8-e4m3 pool (kv_cache_dtype "auto" = the model dtype, model_runner.py:366-367; fp32 checkpoints
 * are served in fp16 by default, config/utils.py:41-43); any other pairing returns B200_ERR_UNSUPPORTED. */
#define B200_ACT_BF16 0
#define B200_ACT_F16 1
/* b200_decode_params.page_size */
#define B200_PAGE_AUTO (-1)
/* idx_dtype */
#define B200_IDX_I32 0
#define B200_IDX_I64 1
/* compute path (b200_attn_config.path) */
#define B200_PATH_MMA 0  /* GQA-packed register-level tensor-core path (mma.sync m16n8k16): decode, generic extend */
#define B200_PATH_TC 1   /* extend only: tcgen05 + TMEM path, one CTA per 256-row tile (head_dim 128 or 64, bf16 or fp8 pool) */
#define B200_PATH_TC2 2  /* extend only: tcgen05 cta_group::2 path, CTA pairs (head_dim 128 over a bf16 pool).  Opt-in: correct
                          * and tested, measured slower than B200_PATH_TC */

/* error codes (cuda errors are reported as 1000 + cudaError_t) */
#define B200_OK 0
#define B200_ERR_INVALID 1     /* null pointer, bad size, misaligned stride */
#define B200_ERR_UNSUPPORTED 2 /* head_dim / group size / dtype not built */
#define B200_ERR_WORKSPACE 3   /* workspace smaller than b200_attn_get_config says */
#define B200_ERR_NOT_PLANNED 4 /* decode called before b200_attn_plan on this workspace */

typedef void* b200_stream_t; /* cudaStream_t */

/* Static launch configuration of the decode path for one model shape.
 * Replaces TritonAttnBackend.__init__'s num_kv_splits / attn_logits sizing
 * (scratchpad/nn/attention/triton_backend.py:32-61). */
typedef struct b200_attn_config {
  int32_t path;            /* B200_PATH_* */
  int32_t tile_tokens;     /* KV tokens per pipeline stage of one CTA */
  int32_t grid_ctas;       /* persistent grid size (x); y = hkv / heads_per_cta */
  int32_t heads_per_cta;   /* kv heads processed by one CTA */
  int32_t sub_splits;      /* token sub-slices per tile that keep separate softmax state */
  int32_t stages;          /* cp.async pipeline depth */
  int32_t smem_bytes;      /* dynamic shared memory per CTA */
  int32_t num_sms;
  int32_t partials_per_slot; /* split-KV partials per (cta, request): 1 where the sub-slices are combined inside the CTA */
  int32_t q_tiles;           /* N=8 tiles of q heads per kv head: 1 (group size <= 8) or 2 (9..16, head_dim 128) */
  uint64_t workspace_bytes; /* for the max_bs passed to b200_attn_get_config */
} b200_attn_config;

/* One layer's batched decode attention (replaces decode_attention_fwd as called
 * from triton_backend.py:183-195 and the FlashInfer decode wrapper call,
 * flashinfer_backend.py:475-482). */
typedef struct b200_decode_params {
  const void* q;            /* act_dtype [bs, hq, d] */
  int64_t q_stride_t, q_stride_h;
  const void* k_buf;        /* [P+1, hkv, d]  bf16 / fp16 or fp8 bytes */
  const void* v_buf;        /* [P+1, hkv, dv] */
  int64_t k_stride_t, v_stride_t; /* token stride in elements */
  void* o;                  /* act_dtype [bs, hq, dv] */
  int64_t o_stride_t, o_stride_h;
  const int32_t* req_to_token;
  int64_t req_to_token_stride;
  int32_t bs, hq, hkv, d, dv;
  int32_t kv_dtype;
  float sm_scale;
  float logit_cap;          /* <= 0: off (radix_attention.py:18; gemma3 passes None) */
  const float* k_scale;     /* device fp32 [hkv] or NULL (=1.0): dequant x_hat * scale */
  const float* v_scale;
  void* workspace;          /* planned by b200_attn_plan for this batch */
  uint64_t workspace_bytes;
  int32_t max_bs;           /* the max_bs the workspace was sized for */
  int32_t force_path;       /* -1 auto, else B200_PATH_* (tests) */
  void* ev_kernel_begin;    /* optional cudaEvent_t pair recorded on `stream` around the */
  void* ev_kernel_end;      /* split-KV kernel alone (bench.py roofline timing); NULL = off */
  int32_t cascade;          /* must equal the `cascade` b200_attn_plan was called with for this batch */
  int32_t page_size;        /* 0 / 1: token-granular pool, row gather only (the reference today, model_runner.py:431-432).
                             * 16, 32, 64 or 128: paged pool -- for every request row r and position t < seq_len,
                             * req_to_token[r, t] = req_to_token[r, t - t % page_size] + t % page_size (what a paged
                             * allocator yields: schedule_batch.py:754-828, radix_cache.py:57-66 page-aligned matching).
                             * When a page covers a KV tile (b200_attn_config.tile_tokens; always at TP1), tiles whose
                             * slot ids really are consecutive are fetched by one TMA box copy each for K and V; the
                             * kernel checks that per tile on device and takes the row gather for any other tile, so
                             * the value is a performance hint, never a correctness contract.  Smaller pages use the
                             * row gather (which sees the pages' DRAM locality).
                             * B200_PAGE_AUTO: token-granular pool, TMA for every tile whose slots happen to be
                             * consecutive (an unfragmented allocator hands out runs, memory/pool.py:240-253). */
  int64_t kv_pool_rows;     /* paged only: token rows of k_buf / v_buf (bounds of the tensor maps) */
  int32_t act_dtype;        /* B200_ACT_*: dtype of q and o (0 = bf16) */
  int32_t reserved0;
} b200_decode_params;

/* One layer's extend / chunked-prefill attention over prefix-cached chunks
 * (replaces extend_attention_fwd as called from triton_backend.py:139-154):
 * query row t of request i attends to the pool slots req_to_token[r, :pre_i]
 * (non-causal) and to k_new/v_new[start_i : start_i + t + 1] (causal). */
typedef struct b200_extend_params {
  const void* q;            /* act_dtype [T, hq, d] */
  int64_t q_stride_t, q_stride_h;
  const void* k_new;        /* act_dtype [T, hkv, d]  (model dtype, unquantised) */
  const void* v_new;        /* act_dtype [T, hkv, dv] */
  int64_t kn_stride_t, kn_stride_h, vn_stride_t, vn_stride_h;
  const void* k_buf;
  const void* v_buf;
  int64_t k_stride_t, v_stride_t;
  void* o;                  /* act_dtype [T, hq, dv] */
  int64_t o_stride_t, o_stride_h;
  const int32_t* req_to_token;
  int64_t req_to_token_stride;
  const void* req_pool_indices; /* [bs] */
  const void* seq_lens;         /* [bs] prefix + extend */
  int32_t idx_dtype;            /* dtype of the two above */
  const int32_t* extend_seq_lens;  /* int32 [bs] (forward_info.py:258-263) */
  const int32_t* extend_start_loc; /* int32 [bs] */
  int32_t bs, hq, hkv, d, dv;
  int32_t total_tokens;         /* T */
  int32_t max_extend_len;       /* upper bound of extend_seq_lens (host value) */
  int32_t kv_dtype;
  float sm_scale;
  float logit_cap;
  const float* k_scale;
  const float* v_scale;
  int32_t force_path;           /* -1 auto, else B200_PATH_MMA / B200_PATH_TC / B200_PATH_TC2 (tests, bench) */
  int32_t page_size;            /* hint as in b200_decode_params: >= 128 or B200_PAGE_AUTO lets the tcgen05 kernel fetch prefix
                                 * tiles whose 128 slots are consecutive by TMA (checked per tile on device); 0 / 1: row gather.
                                 * Tiles of 128 new tokens (rows of k_new / v_new) are fetched by TMA whatever the value. */
  int64_t kv_pool_rows;         /* token rows of k_buf / v_buf (bounds of the pool's tensor maps); 0 = unknown, no prefix TMA */
  int32_t act_dtype;            /* B200_ACT_*: dtype of q, k_new, v_new and o (0 = bf16) */
  int32_t reserved0;
} b200_extend_params;

int b200_attn_abi_version(void);
/* sizeof() of the structs above as compiled (which: 0 config, 1 decode, 2 extend,
 * 3 rope-store), so a binding can verify its own layout before the first call. */
uint64_t b200_attn_struct_size(int which);
const char* b200_attn_error_string(int code);

/* Launch configuration + workspace size for (max_bs, shape) on the current device. */
int b200_attn_get_config(int32_t max_bs, int32_t hq, int32_t hkv, int32_t d, int32_t dv,
                         int32_t kv_dtype, int32_t force_path, b200_attn_config* out);

/* Per-step metadata, once per forward batch (replaces init_forward_metadata's
 * attn_logits allocation, triton_backend.py:48-68, and FlashInfer's kv_indices
 * flatten + host-side plan, flashinfer_backend.py:623-670): normalises the index
 * dtypes and balances KV tiles over the persistent grid, entirely on device. */
/* cascade != 0 additionally looks for shared prompt prefixes: requests whose req_to_token rows start with the same
 * slot id form a prefix group (RadixAttention gives requests with a common prompt the same slot ids,
 * memory/radix_cache.py:105-135); a group's common prefix, rounded down to 128 tokens, is used when it is >= 1024
 * tokens and the group has >= 64 query rows per kv head.  b200_attn_decode then computes that prefix once per kv
 * head as a dense tcgen05 contraction over all queries of the group and only the per-request suffixes through the
 * split-KV kernel; the LSE merge combines both.  Detection is on device (no host sync, CUDA-graph safe); head_dim
 * 128 or 64 and bs <= 1024; otherwise the flag is ignored.
 * req_to_token may be NULL when cascade == 0. */
int b200_attn_plan(void* workspace, uint64_t workspace_bytes, int32_t max_bs,
                   const void* req_pool_indices, const void* seq_lens, int32_t idx_dtype,
                   int32_t bs, int32_t hq, int32_t hkv, int32_t d, int32_t dv, int32_t kv_dtype,
                   int32_t force_path, const int32_t* req_to_token, int64_t req_to_token_stride,
                   int32_t cascade, b200_stream_t stream);

int b200_attn_decode(const b200_decode_params* p, b200_stream_t stream);
int b200_attn_extend(const b200_extend_params* p, b200_stream_t stream);

/* KV write: k_buf[loc[t]] = quant(k[t]) (replaces MHATokenToKVPool.set_kv_buffer's
 * two index_put_ launches + div_/cast, pool.py:392-424).  k/v are act_dtype (bf16 or fp16)
 * [T, hkv, d] with strides in elements; loc int32/int64 [T]; for fp8 the value
 * stored is e4m3(act(x / scale[h])) -- the division rounded to the model dtype first, as the reference's
 * in-place div_ does -- and k/v are NOT modified in place. */
int b200_kv_store(void* k_buf, void* v_buf, int64_t k_stride_t, int64_t v_stride_t,
                  const void* loc, int32_t loc_dtype, const void* k, const void* v,
                  int64_t k_src_stride_t, int64_t k_src_stride_h, int64_t v_src_stride_t,
                  int64_t v_src_stride_h, int32_t num_tokens, int32_t hkv, int32_t d, int32_t dv,
                  int32_t kv_dtype, int32_t act_dtype, const float* k_scale, const float* v_scale,
                  b200_stream_t stream);

/* Rotary position embedding fused with the KV write: one launch replaces the model's
 * `q, k = self.rotary_emb(positions, q, k)` (scratchpad/nn/models/llama/llama.py:149-152 ->
 * RotaryEmbedding.forward_cuda, scratchpad/nn/layers/rotary_embedding.py:131-164, an in-place kernel on q and k)
 * and MHATokenToKVPool.set_kv_buffer (memory/pool.py:392-424, two index_put_ + div_/cast).
 * Arithmetic is RotaryEmbedding.forward_native's (rotary_embedding.py:24-49,101-129) in the model dtype: every
 * product and every sum is rounded to bf16 (fp16 when act_dtype says so), so results are bit-identical to the reference's:
 *   o1 = bf16(bf16(x1*cos) - bf16(x2*sin)),  o2 = bf16(bf16(x2*cos) + bf16(x1*sin)),
 * (x1, x2) = (x[i], x[i + rotary_dim/2]) for neox style, (x[2i], x[2i+1]) for GPT-J style; elements
 * >= rotary_dim pass through.  cos/sin come from the module's cos_sin_cache[positions[t]] = [cos | sin]
 * (rotary_embedding.py:90-99), whatever scaling variant built it (llama3, linear, dynamic NTK, ...).
 * q is rotated in place.  k is rotated and stored as k_buf[loc[t]] = quant(rot(k[t])); it is written back
 * in place only when write_back_k != 0 (the extend kernels read the new tokens' K from the k argument;
 * decode reads it from the pool).  v is stored as b200_kv_store does. */
typedef struct b200_rope_store_params {
  void* q;                  /* act_dtype [T, hq, d], rotated in place; NULL = no q */
  int64_t q_stride_t, q_stride_h;
  void* k;                  /* act_dtype [T, hkv, d] */
  int64_t k_src_stride_t, k_src_stride_h;
  const void* v;            /* act_dtype [T, hkv, dv] */
  int64_t v_src_stride_t, v_src_stride_h;
  const void* positions;    /* [T] int32 / int64 (pos_dtype = B200_IDX_*) */
  const void* cos_sin_cache; /* [max_position, rotary_dim] = [cos(rotary_dim/2) | sin(rotary_dim/2)] */
  int64_t cache_stride;     /* row stride in elements */
  int32_t cache_dtype;      /* B200_CACHE_BF16 / B200_CACHE_F16 (must be the model dtype) / B200_CACHE_F32 (rounded to the
                             * model dtype first: cos.to(x.dtype)) */
  int32_t pos_dtype;
  int32_t rotary_dim;       /* multiple of 16, <= d */
  int32_t is_neox;
  void* k_buf;              /* NULL (with v_buf) = rotate only, nothing stored (save_kv_cache = False) */
  void* v_buf;
  int64_t k_stride_t, v_stride_t; /* pool token strides in elements */
  const void* loc;          /* [T] slot ids (out_cache_loc) */
  int32_t loc_dtype;
  int32_t num_tokens, hq, hkv, d, dv;
  int32_t kv_dtype;
  const float* k_scale;
  const float* v_scale;
  int32_t write_back_k;
  int32_t act_dtype;        /* B200_ACT_*: dtype of q, k and v (0 = bf16) */
} b200_rope_store_params;
#define B200_CACHE_BF16 0
#define B200_CACHE_F32 1
#define B200_CACHE_F16 2

int b200_rope_kv_store(const b200_rope_store_params* p, b200_stream_t stream);

/* Test hook for the bit-exact gather contract: copies, with the kernels' own
 * slot -> row address arithmetic, rows buf[req_to_token[req_pool_indices[i], t]]
 * for t < seq_lens[i] to out + out_offsets[i] * row_bytes (rows packed). */
int b200_debug_gather(const void* buf, int64_t stride_t_bytes, int32_t row_bytes,
                      const int32_t* req_to_token, int64_t req_to_token_stride,
                      const void* req_pool_indices, const void* seq_lens, int32_t idx_dtype,
                      const int64_t* out_offsets, int32_t bs, void* out, b200_stream_t stream);

/* Test / bench introspection: how the last b200_attn_decode call of this process fetched KV tiles --
 * 0 row gather only, 1 TMA (one box per K / V tile) for every tile whose slots are consecutive, row gather for the rest. */
int b200_debug_last_tma_mode(void);

/* Test hook: the split-KV kernel hands 7/8 of a launch's KV tiles to its CTAs as equal static ranges and lets them claim
 * the rest in small chunks (one atomicAdd each) once a CTA's share reaches `tiles` tiles -- by default only launches of
 * ~360 us and more whose slots the plan finds handed out as runs, where the per-SM bandwidth spread costs more than the
 * chunk switches and the extra partials.  Tests lower the threshold to
 * run that schedule on small batches; <= 0 restores the default.  Returns the previous value; affects later plans. */
int b200_debug_set_dyn_min_base(int32_t tiles);

/* Test hook (host only): the tile schedule b200_attn_plan writes for `total` KV tiles on a persistent grid of `grid` CTAs
 * when the slots are runs (min_base <= 0: the library's threshold).  out[8] = tiles per static chunk, tiles per dynamic
 * chunk, first dynamically claimed tile, number of chunks, chunk capacity of the workspace, the chunk owning tile
 * `probe`, the first tile of that chunk and of the next one (-1 when probe is out of range). */
int b200_debug_sched(int32_t total, int32_t grid, int32_t min_base, int32_t probe, int32_t* out);

/* Test hook (synchronises the device): KV tiles the decode kernels fetched by TMA box copies on the current device
 * since the last call with reset != 0 (-1 on a CUDA error).  b200_debug_last_tma_mode() says the TMA-capable kernel
 * was launched; this says how often its TMA branch was actually taken. */
int64_t b200_debug_tma_tiles(int32_t reset);
/* The same for the tcgen05 extend kernel: K / V tiles (128 keys of one kv head) fetched by one TMA box copy each. */
int64_t b200_debug_extend_tma_tiles(int32_t reset);

/* Debug check of the paged-table contract (b200_decode_params.page_size): writes to *violations (device int32)
 * the number of positions t < seq_lens[i] with req_to_token[r_i, t] != page_start + t % page_size, page_start =
 * req_to_token[r_i, t - t % page_size] a multiple of page_size.  CudaB200AttnBackend runs it once per forward batch
 * when B200_CHECK_PAGES=1 (it synchronises to read the count; never on the normal path). */
int b200_debug_check_paged(const int32_t* req_to_token, int64_t req_to_token_stride,
                           const void* req_pool_indices, const void* seq_lens, int32_t idx_dtype,
                           int32_t bs, int32_t page_size, int32_t* violations, b200_stream_t stream);

/* Test hook: straightforward one-warp-per-(request, q head) fp32 decode used to
 * cross-check the tiled kernels at sizes the CPU oracle cannot reach. */
int b200_debug_decode_naive(const b200_decode_params* p, const void* req_pool_indices,
                            const void* seq_lens, int32_t idx_dtype, b200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_ATTN_H_ */
