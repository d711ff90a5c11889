// Micro-benchmark: does a cta_group::2 tcgen05.mma (M = 256 over a pair of SMs, A from tensor memory) cost the same
// ~45 issue cycles as the cta_group::1 instruction (M = 128)?  If so the per-row MMA floor of the sparse conv halves.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_2cta_bench mma_2cta_bench.cu && ./mma_2cta_bench
// 74 clusters of 2 CTAs; the leader CTA's thread 0 issues 16 x rounds MMAs back to back, commits (multicast) and waits.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) bench2_kernel(long long* out, int rounds) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // B region: 4 chunks x (N/2 rows per CTA) x 128 bytes
  for (uint32_t i = threadIdx.x; i < (4 * (N / 2) * 128) / 4; i += blockDim.x)
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + i * 4), "r"(0) : "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster.sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  // M = 256 (128 rows per CTA), N, K = 16
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  if (rank == 0 && threadIdx.x == 0) {
    long long t0, t1;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0));
    for (int r = 0; r < rounds; ++r) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int sub = i / 4, kk = i % 4;
        const uint64_t bdesc = make_desc_sw128(base + sub * (N / 2) * 128) + 2 * kk;
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(tmem),
                     "r"(tmem + 256 + sub * 32 + 8 * kk), "l"(bdesc), "r"(idesc), "r"(1u)
                     : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)),
                 "h"((uint16_t)1)
                 : "memory");
    uint32_t done;
    do {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    } while (!done);
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1));
    out[blockIdx.x / 2] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster.sync();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <int N>
void run() {
  long long* d;
  cudaMalloc(&d, 74 * sizeof(long long));
  cudaMemset(d, 0, 74 * sizeof(long long));
  const int smem = 1024 + 4 * (N / 2) * 128;
  cudaFuncSetAttribute(bench2_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int rounds = 64;
  for (int it = 0; it < 2; ++it) bench2_kernel<N><<<148, 128, smem>>>(d, rounds);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[74];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < 74; ++i) s += (double)h[i];
  printf("cta_group::2, A in TMEM, M=256 N=%3d K=16: %6.1f cycles per MMA (= per 256 rows)  (%s)\n", N, s / 74 / (rounds * 16),
         cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<16>();
  run<32>();
  run<64>();
  run<128>();
  return 0;
}
