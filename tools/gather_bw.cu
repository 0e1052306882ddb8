// Gather-bandwidth microbenchmark: the ceiling for reading `rows` random rows of `row_bytes`
// contiguous bytes (stride = stride_bytes) from a large HBM buffer.  Used to put the page=1
// token-row gather of the decode kernel in context.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bw gather_bw.cu && ./gather_bw
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>

template <int ROWB>
__global__ void __launch_bounds__(256) gather_read(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                   const int* __restrict__ idx, int64_t stride, int nrows,
                                                   uint4* __restrict__ sink) {
  constexpr int LPR = ROWB / 16;          // lanes per row
  constexpr int RPI = 256 / LPR;          // rows per CTA iteration
  constexpr int U = 8;                    // iterations in flight
  const int sub = threadIdx.x % LPR, r_in = threadIdx.x / LPR;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (int64_t base = (int64_t)blockIdx.x * RPI * U; base < nrows; base += (int64_t)gridDim.x * RPI * U) {
    uint4 va[U], vb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t r = base + u * RPI + r_in;
      int id = r < nrows ? __ldg(idx + r) : 0;
      const uint4* pa = reinterpret_cast<const uint4*>(a + (int64_t)id * stride) + sub;
      const uint4* pb = reinterpret_cast<const uint4*>(b + (int64_t)id * stride) + sub;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(va[u].x), "=r"(va[u].y), "=r"(va[u].z), "=r"(va[u].w) : "l"(pa));
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(vb[u].x), "=r"(vb[u].y), "=r"(vb[u].z), "=r"(vb[u].w) : "l"(pb));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) { acc.x ^= va[u].x ^ vb[u].x; acc.y ^= va[u].y ^ vb[u].y; acc.z ^= va[u].z ^ vb[u].z; acc.w ^= va[u].w ^ vb[u].w; }
  }
  if (acc.x == 0x12345678u && acc.y == 0x9abcdef0u) sink[blockIdx.x * 256 + threadIdx.x] = acc;
}

template <int ROWB>
static void run(const uint8_t* a, const uint8_t* b, const int* idx, int nrows, uint4* sink, const char* tag) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int ctas_per_sm : {2, 4, 8}) {
    int grid = 148 * ctas_per_sm;
    for (int i = 0; i < 3; ++i) gather_read<ROWB><<<grid, 256>>>(a, b, idx, ROWB, nrows, sink);
    cudaEventRecord(e0);
    const int it = 10;
    for (int i = 0; i < it; ++i) gather_read<ROWB><<<grid, 256>>>(a, b, idx, ROWB, nrows, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double bytes = 2.0 * nrows * ROWB * it;
    printf("%s row_bytes=%4d ctas/sm=%d  %.1f us/launch  %.0f GB/s\n", tag, ROWB, ctas_per_sm, ms * 1e3 / it, bytes / (ms * 1e-3) / 1e9);
  }
}

int main() {
  const size_t buf = 2ull << 30;  // two 2 GiB buffers (K and V)
  uint8_t *a, *b; uint4* sink; int* idx;
  cudaMalloc(&a, buf); cudaMalloc(&b, buf); cudaMalloc(&sink, 148 * 8 * 256 * 16);
  cudaMemset(a, 1, buf); cudaMemset(b, 2, buf);
  for (int mode = 0; mode < 2; ++mode) {
    for (int rowb : {128, 256, 512, 1024, 2048}) {
      int nrows = (int)(buf / rowb);
      std::vector<int> h(nrows);
      for (int i = 0; i < nrows; ++i) h[i] = i;
      if (mode == 0) { std::mt19937 g(1); std::shuffle(h.begin(), h.end(), g); }
      cudaMalloc(&idx, (size_t)nrows * 4);
      cudaMemcpy(idx, h.data(), (size_t)nrows * 4, cudaMemcpyHostToDevice);
      const char* tag = mode == 0 ? "random    " : "contiguous";
      switch (rowb) {
        case 128: run<128>(a, b, idx, nrows, sink, tag); break;
        case 256: run<256>(a, b, idx, nrows, sink, tag); break;
        case 512: run<512>(a, b, idx, nrows, sink, tag); break;
        case 1024: run<1024>(a, b, idx, nrows, sink, tag); break;
        case 2048: run<2048>(a, b, idx, nrows, sink, tag); break;
      }
      cudaFree(idx);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
