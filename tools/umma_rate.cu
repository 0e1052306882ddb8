// tcgen05.mma issue/throughput microbenchmark (sm_100a): clocks per instruction for the UMMA shapes the
// extend kernel can choose from (kind::f16, bf16 x bf16 -> fp32, cta_group::1, M = 128):
//   SS (A and B from shared memory) with N = 64 / 128 / 256, TS (A from TMEM) with N = 128 / 256.
// One CTA per SM on all SMs (so the chip is at its sustained tensor power state), one thread issues `reps`
// back-to-back MMAs on the same accumulator, then commit + wait.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_rate umma_rate.cu && ./umma_rate
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3fffu);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3fffu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t instr_desc(int m, int n, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(b_mn) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

template <int N, bool TS, int NACC, int M = 128>
__global__ void __launch_bounds__(128, 1) rate(int reps, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < (64 + 64) * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(gen)[i] = 0;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = instr_desc(M, N, false);
    const uint64_t adesc = smem_desc(base, 16, 1024);            // A: 128 rows x 16 k (one 32-byte k-step of a 128B-swizzled block)
    const uint64_t bdesc = smem_desc(base + 65536, 16, 1024);    // B: N rows x 16 k
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      // accumulators rotate over NACC column blocks; TS reads A from TMEM columns 448..455
      const uint32_t d = tmem + (r % NACC) * N;
      if constexpr (TS) {
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d),
                     "r"(tmem + 448), "l"(bdesc), "r"(idesc), "r"(1)
                     : "memory");
      } else {
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
                     "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1)
                     : "memory");
      }
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    const long long t2 = clock64();
    out[2 * blockIdx.x] = t1 - t0;
    out[2 * blockIdx.x + 1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int N, bool TS, int NACC, int M = 128>
void run(const char* name, int ctas) {
  const int reps = 4096, smem = 128 * 1024 + 1024 + 64;
  long long* d;
  cudaMalloc(&d, ctas * 16);
  cudaFuncSetAttribute(rate<N, TS, NACC, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int it = 0; it < 3; ++it) rate<N, TS, NACC, M><<<ctas, 128, smem>>>(reps, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%s: %s\n", name, cudaGetErrorString(e));
    return;
  }
  long long h[2 * 148];
  cudaMemcpy(h, d, ctas * 16, cudaMemcpyDeviceToHost);
  double issue = 0, total = 0;
  for (int i = 0; i < ctas; ++i) {
    issue += h[2 * i];
    total += h[2 * i + 1];
  }
  issue /= ctas * reps;
  total /= ctas * reps;
  const double flops_per_clk = 2.0 * M * N * 16 / total;
  printf("%-28s ctas=%3d  issue %.1f clk/instr  complete %.1f clk/instr  %.0f flop/clk/SM (%.0f%% of 8192)\n", name, ctas, issue, total,
         flops_per_clk, flops_per_clk / 81.92);
  cudaFree(d);
}

int main() {
  for (int ctas : {148}) {
    run<64, false, 1>("SS M128 N64  1 acc", ctas);
    run<64, false, 4>("SS M128 N64  4 acc", ctas);
    run<128, false, 1>("SS M128 N128 1 acc", ctas);
    run<128, false, 2>("SS M128 N128 2 acc", ctas);
    run<256, false, 1>("SS M128 N256 1 acc", ctas);
    run<128, true, 1>("TS M128 N128 1 acc", ctas);
    run<128, true, 2>("TS M128 N128 2 acc", ctas);
    run<256, true, 1>("TS M128 N256 1 acc", ctas);
    run<128, false, 1, 64>("SS M64  N128 1 acc", ctas);
    run<128, false, 2, 64>("SS M64  N128 2 acc", ctas);
    run<256, false, 1, 64>("SS M64  N256 1 acc", ctas);
    run<128, true, 1, 64>("TS M64  N128 1 acc", ctas);
    run<128, true, 2, 64>("TS M64  N128 2 acc", ctas);
  }
  return 0;
}
