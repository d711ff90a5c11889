// Micro-benchmark: mbarrier costs on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mbar_bench mbar_bench.cu && ./mbar_bench
// 1. try_wait / test_wait on an already completed phase: cycles per call (one warp, lane 0 or all lanes).
// 2. wake-up latency: warp A blocks in try_wait (or spins on test_wait, or try_wait + nanosleep back-off), warp B
//    arrives; cycles from the arrive to A's first instruction after the wait.  With 0 / 15 other warps spinning.
// 3. ping-pong between two warps over two barriers: cycles per round trip.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ bool test_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)); return t; }

// mode 0: completed-phase cost.  kind 0 try_wait, 1 test_wait.
// mode 1: wake-up latency.  kind 0 try_wait loop, 1 test_wait spin, 2 try_wait + nanosleep(40) back-off.  spin = other warps spinning on a never-completing barrier
// mode 2: ping-pong.  kind as mode 1
__global__ void __launch_bounds__(1024, 1) k(int mode, int kind, int spin, int rounds, long long* out) {
  __shared__ uint64_t bars[8];
  __shared__ long long t_arrive;
  const uint32_t b0 = smem_u32(&bars[0]), b1 = smem_u32(&bars[1]), bn = smem_u32(&bars[2]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(b0, 1); mbar_init(b1, 1); mbar_init(bn, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    t_arrive = 0;
  }
  __syncthreads();
  auto wait = [&](uint32_t bar, uint32_t par) {
    if (kind == 0) { while (!try_wait(bar, par)) {} }
    else if (kind == 1) { while (!test_wait(bar, par)) {} }
    else { while (!try_wait(bar, par)) __nanosleep(40); }
  };
  if (mode == 0) {
    if (warp == 0) {
      if (lane == 0) arrive(b0);
      __syncwarp();
      long long t0 = clk();
      uint32_t acc = 0;
      for (int r = 0; r < rounds; ++r) acc += kind == 0 ? try_wait(b0, 0) : test_wait(b0, 0);
      long long t1 = clk();
      if (lane == 0) out[blockIdx.x] = (t1 - t0) + (acc == 12345 ? 1 : 0);
    }
  } else if (mode == 1) {
    // warp 0 waits, warp 4 (same scheduler) or 1 arrives after a delay; warps 8.. spin if asked
    long long tot = 0;
    for (int r = 0; r < rounds; ++r) {
      const uint32_t par = r & 1;
      if (warp == 0) {
        wait(b0, par);
        long long t = clk();
        if (lane == 0) tot += t - *(volatile long long*)&t_arrive;
      } else if (warp == 1) {
        long long t = clk();
        while (clk() - t < 3000) {}
        if (lane == 0) { *(volatile long long*)&t_arrive = clk(); arrive(b0); }
      } else if (warp >= 8 && warp < 8 + spin) {
        // spin for about the same time on a barrier that never completes
        long long t = clk();
        while (clk() - t < 2500) { if (kind == 1) test_wait(bn, 0); else if (kind == 0) try_wait(bn, 0); else { try_wait(bn, 0); __nanosleep(40); } }
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = tot;
  } else if (mode == 2) {
    if (warp == 0) {
      long long t0 = clk();
      for (int r = 0; r < rounds; ++r) {
        if (lane == 0) arrive(b0);
        wait(b1, r & 1);
      }
      if (lane == 0) out[blockIdx.x] = clk() - t0;
    } else if (warp == 1) {
      for (int r = 0; r < rounds; ++r) {
        wait(b0, r & 1);
        if (lane == 0) arrive(b1);
      }
    } else if (warp >= 8 && warp < 8 + spin) {
      wait(bn, 0 ^ 0);   // never completes ... until the end
    }
    __syncthreads_or(0);
    if (threadIdx.x == 0) arrive(bn);
  }
}

double run(int mode, int kind, int spin, int rounds, int threads) {
  long long* d;
  cudaMalloc(&d, 148 * 8);
  cudaMemset(d, 0, 148 * 8);
  k<<<148, threads>>>(mode, kind, spin, rounds, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < 148; ++i) s += (double)h[i];
  cudaFree(d);
  return s / 148 / rounds;
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  printf("1. wait on a completed phase, cycles per call: try_wait %.1f  test_wait %.1f\n", run(0, 0, 0, 1000, 64), run(0, 1, 0, 1000, 64));
  const char* names[3] = {"try_wait loop", "test_wait spin", "try_wait + nanosleep(40)"};
  for (int kind = 0; kind < 3; ++kind)
    for (int spin : {0, 8, 16})
      printf("2. wake-up latency, %-26s %2d other warps waiting the same way: %7.1f cycles\n", names[kind], spin, run(1, kind, spin, 50, 1024));
  for (int kind = 0; kind < 3; ++kind)
    printf("3. ping-pong round trip (2 hand-offs), %-26s: %7.1f cycles\n", names[kind], run(2, kind, 0, 1000, 64));
  return 0;
}
