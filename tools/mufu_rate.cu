// MUFU ex2 throughput microbenchmark (sm_100a): f32, f16, f16x2, bf16x2 -- results per clock and SM.
// The tcgen05 extend kernel is co-bound by the MUFU (128 x 128 exp2 per 128-key tile and Q tile); this answers whether a
// half-precision exp2 would retire faster on B200 (it does on B300, whose SFU rate was doubled).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mufu_rate mufu_rate.cu && ./mufu_rate
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

template <int MODE>
__global__ void __launch_bounds__(256) k(int reps, uint32_t seed, uint32_t* out, long long* clk) {
  uint32_t x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = seed + threadIdx.x * 8 + i;  // arbitrary bit patterns; only throughput matters
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if constexpr (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(x[i]));
      if constexpr (MODE == 1) { uint16_t h = static_cast<uint16_t>(x[i]); asm volatile("ex2.approx.f16 %0, %0;" : "+h"(h)); x[i] = h; }
      if constexpr (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[i]));
      if constexpr (MODE == 3) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(x[i]));
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_instr) {
  const int ctas = 148, reps = 2048;
  uint32_t* out;
  long long* clk;
  cudaMalloc(&out, ctas * 256 * 4);
  cudaMalloc(&clk, ctas * 8);
  for (int it = 0; it < 2; ++it) k<MODE><<<ctas, 256>>>(reps, 12345u, out, clk);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, ctas * 8, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < ctas; ++i) avg += h[i];
  avg /= ctas;
  const double results = 256.0 * 8 * reps * per_instr;  // per CTA = per SM (one CTA of 8 warps per SM)
  printf("%-22s %8.0f clk  %.2f exp2 results / clk / SM\n", name, avg, results / avg);
  cudaFree(out);
  cudaFree(clk);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16", 1);
  run<2>("ex2.approx.f16x2", 2);
  run<3>("ex2.approx.ftz.bf16x2", 2);
  return 0;
}
