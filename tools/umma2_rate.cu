// tcgen05.mma.cta_group::2 throughput microbenchmark (sm_100a): clocks per M = 256 instruction issued by the leader CTA
// of a 2-CTA cluster (kind::f16, bf16 x bf16 -> fp32), SS and TS, N = 64 / 128 / 256.  74 clusters = every SM busy.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma2_rate umma2_rate.cu && ./umma2_rate
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
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int N, bool TS, int NACC, bool BMN>
__global__ void __launch_bounds__(128, 1) rate2(int reps, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < (64 + 64) * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(gen)[i] = 0;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    long long t0 = 0, t1 = 0;
    if (rank == 0) {
      constexpr uint32_t idesc = instr_desc(256, N, BMN);
      const uint64_t adesc = smem_desc(base, 16, 1024);          // A: 128 rows per CTA x 16 k
      const uint64_t bdesc = smem_desc(base + 65536, 16, 1024);  // B: N / 2 rows per CTA x 16 k (or MN-major)
      t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        const uint32_t d = tmem + (r % NACC) * N;
        if constexpr (TS) {
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d),
                       "r"(tmem + 448), "l"(bdesc), "r"(idesc), "r"(1)
                       : "memory");
        } else {
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
                       "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1)
                       : "memory");
        }
      }
      t1 = clock64();
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)),
                   "h"(static_cast<uint16_t>(3))
                   : "memory");
    }
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    const long long t2 = clock64();
    if (rank == 0) {
      out[blockIdx.x] = t1 - t0;
      out[blockIdx.x + 1] = t2 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int N, bool TS, int NACC, bool BMN = false>
void run(const char* name) {
  const int ctas = 148, reps = 4096, smem = 128 * 1024 + 1024 + 64;
  long long* d;
  cudaMalloc(&d, ctas * 8);
  cudaMemset(d, 0, ctas * 8);
  auto kern = rate2<N, TS, NACC, BMN>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  for (int it = 0; it < 3; ++it) cudaLaunchKernelEx(&cfg, kern, reps, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%s: %s\n", name, cudaGetErrorString(e));
    return;
  }
  long long h[148];
  cudaMemcpy(h, d, ctas * 8, cudaMemcpyDeviceToHost);
  double issue = 0, total = 0;
  for (int i = 0; i < ctas; i += 2) {
    issue += h[i];
    total += h[i + 1];
  }
  issue /= (ctas / 2) * reps;
  total /= (ctas / 2) * reps;
  const double flops_per_clk_sm = 2.0 * 256 * N * 16 / total / 2;
  printf("%-30s issue %.1f clk/instr  complete %.1f clk/instr  %.0f flop/clk/SM (%.0f%% of 8192)\n", name, issue, total, flops_per_clk_sm,
         flops_per_clk_sm / 81.92);
  cudaFree(d);
}

int main() {
  run<64, false, 1>("2CTA SS M256 N64  1 acc");
  run<128, false, 1>("2CTA SS M256 N128 1 acc");
  run<128, false, 2>("2CTA SS M256 N128 2 acc");
  run<256, false, 1>("2CTA SS M256 N256 1 acc");
  run<128, false, 1, true>("2CTA SS M256 N128 B MN-major");
  run<128, true, 1>("2CTA TS M256 N128 1 acc");
  run<128, true, 1, true>("2CTA TS M256 N128 B MN-major");
  run<256, true, 1>("2CTA TS M256 N256 1 acc");
  return 0;
}
