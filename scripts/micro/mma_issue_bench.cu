// Micro-benchmark: what does ONE tcgen05.mma (M=128, K=16, bf16) cost when issued back to back with no other work?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_issue_bench mma_issue_bench.cu && ./mma_issue_bench
// Variants: A operand from tensor memory (TS) or shared memory (SS); N = 16..128; one accumulator (dependent chain)
// or NACC accumulators used round-robin.  Reports cycles per MMA measured with clock64 in the issuing thread
// (issue of 256 MMAs + commit + wait for completion), one CTA per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int N, bool TS, int NACC, int ORDER>
__global__ void __launch_bounds__(128, 1) bench_kernel(long long* out, int rounds) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // A region (SS): 4 chunks x 16 KB; B region: 4 chunks x N*128 bytes
  const uint32_t a_base = base, b_base = base + 4 * 16384;
  for (uint32_t i = threadIdx.x; i < (4 * 16384 + 4 * N * 128) / 4; i += blockDim.x)
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + i * 4), "r"(0) : "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (threadIdx.x == 0) {
    long long t0, t1;
    uint32_t parity = 0;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0));
    for (int r = 0; r < rounds; ++r) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int sub = ORDER == 0 ? i / 4 : i % 4, kk = ORDER == 0 ? i % 4 : i / 4;
        const uint32_t d = tmem + (uint32_t)((i / 4) % NACC) * N;
        const uint64_t bdesc = make_desc_sw128(b_base + sub * N * 128) + 2 * kk;
        if (TS) mma_ts(d, tmem + 256 + sub * 32 + 8 * kk, bdesc, idesc, 1u);
        else mma_ss(d, make_desc_sw128(a_base + sub * 16384) + 2 * kk, bdesc, idesc, 1u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done;
    do {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
    } while (!done);
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1));
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <int N, bool TS, int NACC, int ORDER>
void run(const char* name) {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  const int smem = 1024 + 4 * 16384 + 4 * N * 128;
  cudaFuncSetAttribute(bench_kernel<N, TS, NACC, ORDER>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int rounds = 64;
  for (int it = 0; it < 2; ++it) bench_kernel<N, TS, NACC, ORDER><<<148, 128, smem>>>(d, rounds);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < 148; ++i) s += (double)h[i];
  printf("%-44s N=%3d  %6.1f cycles/MMA  (%s)\n", name, N, s / 148 / (rounds * 16), cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<16, true, 1, 0>("TS, 1 accumulator, chunk-major");
  run<32, true, 1, 0>("TS, 1 accumulator, chunk-major");
  run<64, true, 1, 0>("TS, 1 accumulator, chunk-major");
  run<128, true, 1, 0>("TS, 1 accumulator, chunk-major");
  run<16, true, 4, 0>("TS, 4 accumulators (per chunk), chunk-major");
  run<64, true, 4, 0>("TS, 4 accumulators (per chunk), chunk-major");
  run<16, true, 1, 1>("TS, 1 accumulator, k-slice-major");
  run<16, false, 1, 0>("SS, 1 accumulator, chunk-major");
  run<32, false, 1, 0>("SS, 1 accumulator, chunk-major");
  run<64, false, 1, 0>("SS, 1 accumulator, chunk-major");
  run<128, false, 1, 0>("SS, 1 accumulator, chunk-major");
  run<16, false, 4, 0>("SS, 4 accumulators (per chunk), chunk-major");
  return 0;
}
