// Micro-benchmark: where do the ~300 cycles per stage of a tcgen05 producer/consumer ring go?  Clean code: one elected
// thread, every operand a compile-time constant or loop counter, the rest of the CTA parked or producing.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ring_bench ring_bench.cu && ./ring_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)); return t; }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.b32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

// mode 0: loop {commit}            1: loop {fence::after; commit}     2: loop {wait(done); fence; commit}
// mode 3: ring NS=4 with one producer warp per slot (4 warps), consumer = elected thread; out[1] = cycles spent in wait
// mode 4: same, consumer arrives with plain mbarrier.arrive instead of tcgen05.commit
// mode 5: ring, consumer does not wait for full at all (commits only), producers wait empty/arrive full
__global__ void __launch_bounds__(768, 1) k(int mode, int rounds, long long* out) {
  __shared__ uint64_t bars[16];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full = smem_u32(&bars[0]), empty = smem_u32(&bars[4]), fin = smem_u32(&bars[8]), done = smem_u32(&bars[9]);
  if (threadIdx.x == 0) {
    const int P6 = mode == 6 ? 4 : mode == 7 ? 8 : mode == 8 ? 16 : 1;
    for (int s = 0; s < 4; ++s) { mbar_init(full + 8 * s, mode >= 6 ? P6 : 1); mbar_init(empty + 8 * s, 1); }
    mbar_init(fin, 1); mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  if (threadIdx.x == 0) arrive(done);      // phase 0 of `done` is complete from here on
  __syncthreads();
  if (warp == 7 && mode < 100) {
    if (elect_one()) {
      long long t0 = clk(), tw = 0;
      if (mode == 0) {
#pragma unroll 1
        for (int r = 0; r < rounds; ++r) commit(empty + 8 * (r & 3));
      } else if (mode == 1) {
#pragma unroll 1
        for (int r = 0; r < rounds; ++r) { fence_after(); commit(empty + 8 * (r & 3)); }
      } else if (mode == 2) {
#pragma unroll 1
        for (int r = 0; r < rounds; ++r) { mbar_wait(done, 0); fence_after(); commit(empty + 8 * (r & 3)); }
      } else {
#pragma unroll 1
        for (int r = 0; r < rounds; ++r) {
          const uint32_t s = mode >= 6 ? (r & 1) : (r & 3), ph = mode >= 6 ? ((r >> 1) & 1) : ((r >> 2) & 1);
          if (mode != 5) {
            long long a = clk();
            mbar_wait(full + 8 * s, ph);
            tw += clk() - a;
          }
          fence_after();
          if (mode == 4) arrive(empty + 8 * s); else commit(empty + 8 * s);
        }
      }
      commit(fin);
      mbar_wait(fin, 0);
      out[blockIdx.x * 2] = clk() - t0;
      out[blockIdx.x * 2 + 1] = tw;
    }
    __syncwarp();
  } else if (mode >= 6 && warp < 7 && warp < ((mode - 6) % 3 == 0 ? 4 : (mode - 6) % 3 == 1 ? 7 : 7)) {
    // handled below (kept for structure)
  }
  if (mode >= 6 && mode <= 11) {
    // ring of 2 slots, ALL P producer warps arrive on every slot's full barrier (conv_tr v2 / conv_ts pattern).
    // modes 6,7,8: P = 4, 8, 16 direct arrivals (barrier count P); modes 9,10,11: same P, but the producers first meet
    // on a named barrier and ONE thread arrives (barrier count 1).
    const int P = mode == 6 || mode == 9 ? 4 : mode == 7 || mode == 10 ? 8 : 16;
    const bool tree = mode >= 9;
    if (warp >= 8 && warp < 8 + P) {
#pragma unroll 1
      for (int r = 0; r < rounds; ++r) {
        const uint32_t s = r & 1, u = r >> 1;
        if (lane == 0) mbar_wait(empty + 8 * s, (u & 1) ^ 1);
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        fence_before();
        if (tree) {
          asm volatile("bar.sync 1, %0;" ::"r"(P * 32) : "memory");
          if (warp == 8 && lane == 0) arrive(full + 8 * s);
        } else if (lane == 0) {
          arrive(full + 8 * s);
        }
      }
    }
  } else if (mode >= 3 && warp < 4) {
    // producer of slot `warp`
    const int n = (rounds - warp + 3) / 4;
#pragma unroll 1
    for (int u = 0; u < n; ++u) {
      mbar_wait(empty + 8 * warp, (u & 1) ^ 1);
      fence_before();
      if (lane == 0) arrive(full + 8 * warp);
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 0) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
  }
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  long long* d;
  cudaMalloc(&d, 148 * 16);
  const char* names[12] = {"loop {commit}", "loop {fence::after; commit}", "loop {wait(done phase); fence; commit}",
                          "ring of 4, consumer commits", "ring of 4, consumer uses mbarrier.arrive", "ring of 4, consumer never waits (commit only)",
                          "ring of 2, 4 producer warps arrive on every slot", "ring of 2, 8 producer warps arrive on every slot",
                          "ring of 2, 16 producer warps arrive on every slot", "ring of 2, 4 warps meet on bar.sync, one arrives",
                          "ring of 2, 8 warps meet on bar.sync, one arrives", "ring of 2, 16 warps meet on bar.sync, one arrives"};
  for (int mode = 0; mode < 12; ++mode) {
    const int rounds = 1024;
    cudaMemset(d, 0, 148 * 16);
    k<<<148, 768>>>(mode, rounds, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[296];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0, w = 0;
    for (int i = 0; i < 148; ++i) { s += (double)h[2 * i]; w += (double)h[2 * i + 1]; }
    printf("%-48s %7.1f cycles per round (of which %.1f in the wait)  %s\n", names[mode], s / 148 / rounds, w / 148 / rounds, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
