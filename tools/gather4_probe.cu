// Probe of the sm_100 TMA row-gather (cp.async.bulk.tensor.2d ... tile::gather4): checks on a real B200 that
//  (1) a 2-D tensor map with box {64 x 1} bf16 and SWIZZLE_128B loads 4 rows x 128 bytes per instruction,
//  (2) the rows land consecutively (128 B apart) with the 16-byte chunk XOR (chunk ^ (row & 7)) the UMMA
//      SWIZZLE_128B canonical layout expects when the destination is 1024-byte aligned,
//  (3) out-of-range row indices are zero-filled.
// This is the access pattern extend_tc.cu uses to gather page-size-1 K/V token rows.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gather4_probe gather4_probe.cu && ./gather4_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap map, const int* idx, int col0, uint16_t* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (static_cast<uint32_t>(__cvta_generic_to_shared(raw)) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - static_cast<uint32_t>(__cvta_generic_to_shared(raw)));
  const uint32_t bar = base + 16384;
  const int lane = threadIdx.x;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = lane; i < 16384 / 2; i += 32) reinterpret_cast<uint16_t*>(gen)[i] = 0xdead;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(16384) : "memory");
  __syncwarp();
  const int i0 = idx[4 * lane], i1 = idx[4 * lane + 1], i2 = idx[4 * lane + 2], i3 = idx[4 * lane + 3];
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          base + lane * 512),
      "l"(&map), "r"(bar), "r"(col0), "r"(i0), "r"(i1), "r"(i2), "r"(i3)
      : "memory");
  uint32_t done = 0;
  while (!done)
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(bar)
        : "memory");
  for (int i = lane; i < 16384 / 2; i += 32) out[i] = reinterpret_cast<uint16_t*>(gen)[i];
}

int main() {
  const int R = 1000, C = 1024;
  std::vector<uint16_t> h(static_cast<size_t>(R) * C);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c) h[static_cast<size_t>(r) * C + c] = static_cast<uint16_t>((r * 131 + c * 7 + 1) & 0x7fff) | 1;
  uint16_t* d;
  cudaMalloc(&d, h.size() * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  std::vector<int> idx(128);
  srand(1);
  for (int i = 0; i < 128; ++i) idx[i] = rand() % R;
  idx[5] = R;        // just past the end
  idx[6] = R + 77;   // far past the end
  idx[7] = -1;       // negative
  idx[40] = idx[41]; // duplicates
  int* didx;
  cudaMalloc(&didx, 512);
  cudaMemcpy(didx, idx.data(), 512, cudaMemcpyHostToDevice);
  uint16_t* dout;
  cudaMalloc(&dout, 16384);

  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&encode), cudaEnableDefault, &qres) != cudaSuccess ||
      !encode) {
    printf("FAIL: no cuTensorMapEncodeTiled\n");
    return 1;
  }
  CUtensorMap map;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(R)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(C) * 2};
  cuuint32_t box[2] = {64, 1};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    printf("FAIL: encode rc=%d\n", static_cast<int>(cr));
    return 1;
  }
  const int col0 = 192;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 64 + 1024);
  probe<<<1, 32, 16384 + 64 + 1024>>>(map, didx, col0, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("FAIL: kernel %s\n", cudaGetErrorString(e));
    return 1;
  }
  std::vector<uint16_t> out(8192);
  cudaMemcpy(out.data(), dout, 16384, cudaMemcpyDeviceToHost);
  int bad = 0, oob_zero = 0, oob_other = 0;
  for (int r = 0; r < 128; ++r) {
    const bool oob = idx[r] < 0 || idx[r] >= R;
    for (int c = 0; c < 8; ++c)
      for (int e2 = 0; e2 < 8; ++e2) {
        const uint16_t got = out[(r * 128 + ((c ^ (r & 7)) << 4)) / 2 + e2];
        if (oob) {
          if (got == 0) ++oob_zero;
          else ++oob_other;
        } else {
          const uint16_t want = h[static_cast<size_t>(idx[r]) * C + col0 + c * 8 + e2];
          if (got != want && bad++ < 8) printf("mismatch row %d chunk %d elem %d: got %04x want %04x\n", r, c, e2, got, want);
        }
      }
  }
  printf("gather4 probe: mismatches=%d oob_zero=%d oob_other=%d (first oob value %04x)\n", bad, oob_zero, oob_other,
         out[(5 * 128 + ((0 ^ 5) << 4)) / 2]);
  printf(bad == 0 ? "PASS\n" : "FAIL\n");
  return bad != 0;
}
