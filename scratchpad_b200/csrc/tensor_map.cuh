// Tensor maps (TMA descriptors) of KV pools and of the new tokens' K / V, shared by decode.cu and extend_tc.cu.
#pragma once

#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include <cstring>
#include <mutex>

#include "common.cuh"

namespace b200 {

// ---- tensor maps of a paged pool ---------------------------------------------------------------------------------
// The pool buffer [rows, Hkv, D] (token stride given, so the fused K|V layout works too) as a 4-d tensor
// (128-byte column block, token rows, column halves of a head, kv heads); shared-memory side swizzled by 128 bytes
// (64 for 64-byte segments) -- the layout tile_chunk_off() describes.
struct PagedMaps {
  CUtensorMap k, v;
};
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int encode_pool_map(CUtensorMap* out, const void* buf, int64_t stride_t_bytes, int64_t rows, int hkv, int d, bool fp8,
                           int tile_tokens, int hpc) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    B200_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres));
    if (!sym || qres != cudaDriverEntryPointSuccess) return B200_ERR_UNSUPPORTED;
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  const int elt = fp8 ? 1 : 2, seg = d * elt, rowb = seg >= 128 ? 128 : seg, halves = seg / rowb;
  // (128-byte column block, token rows, column halves, kv heads): a box over all heads of the CTA and both halves
  // lands as [head][half][token][128 B] -- the tile layout -- so a tile of consecutive slots is ONE copy.  The strides
  // (row, 128 B, D*elt) are not monotonic; the driver accepts that.
  const cuuint64_t gdim[4] = {static_cast<cuuint64_t>(rowb / elt), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(halves),
                              static_cast<cuuint64_t>(hkv)};
  const cuuint64_t gstride[3] = {static_cast<cuuint64_t>(stride_t_bytes), static_cast<cuuint64_t>(rowb), static_cast<cuuint64_t>(seg)};
  const cuuint32_t box[4] = {static_cast<cuuint32_t>(rowb / elt), static_cast<cuuint32_t>(tile_tokens),
                             static_cast<cuuint32_t>(halves), static_cast<cuuint32_t>(hpc)};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const CUresult r = fn(out, fp8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(buf), gdim,
                        gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? B200_OK : B200_ERR_INVALID;
}

// The two maps of a layer's pool depend only on (buffer, stride, rows, shape, tile): encoded once and kept, so a
// layer-call over a paged pool does no driver work on the host after the first one.
struct MapKey {
  const void* buf;
  int64_t stride_t_bytes, rows;
  int hkv, d, fp8, tile_tokens, hpc, dev;
};
struct MapSlot {
  MapKey key;
  CUtensorMap map;
  bool used = false;
};
static int cached_pool_map(CUtensorMap* out, const void* buf, int64_t stride_t_bytes, int64_t rows, int hkv, int d, bool fp8,
                           int tile_tokens, int hpc) {
  constexpr int kSlots = 1024;  // direct-mapped: K and V of a few hundred layers
  static MapSlot slots[kSlots];
  static std::mutex mu;
  int dev = 0;
  if (int rc = current_device(&dev)) return rc;
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.buf = buf; key.stride_t_bytes = stride_t_bytes; key.rows = rows;
  key.hkv = hkv; key.d = d; key.fp8 = fp8; key.tile_tokens = tile_tokens; key.hpc = hpc; key.dev = dev;
  const uint64_t h = (reinterpret_cast<uint64_t>(buf) >> 8) * 0x9E3779B97F4A7C15ull;
  MapSlot& s = slots[(h >> 40) % kSlots];
  std::lock_guard<std::mutex> lock(mu);
  if (!s.used || memcmp(&s.key, &key, sizeof(key)) != 0) {
    if (int rc = encode_pool_map(&s.map, buf, stride_t_bytes, rows, hkv, d, fp8, tile_tokens, hpc)) {
      s.used = false;
      return rc;
    }
    s.key = key;
    s.used = true;
  }
  *out = s.map;
  return B200_OK;
}

}  // namespace b200
