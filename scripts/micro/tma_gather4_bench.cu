// Micro-benchmark: throughput and placement of TMA tile::gather4 (cp.async.bulk.tensor.2d ... tile::gather4) on B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather4_bench tma_gather4_bench.cu && ./tma_gather4_bench
// One warp per CTA (148 CTAs) gathers 128-byte bf16 rows of a (N x 64) matrix into a ring of 16 KB chunk buffers in
// shared memory (SWIZZLE_128B tensor map, box {64,1}): 32 gather4 per chunk (lane l fetches chunk rows 4l..4l+3),
// an mbarrier with expect_tx = 16 KB per chunk.  Reports cycles per 16 KB chunk for several index patterns and
// fractions of out-of-range (-1) rows, and checks WHERE the bytes land (the canonical K-major SW128 layout?).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int kRing = 4;

__global__ void __launch_bounds__(32, 1) gather_kernel(const __grid_constant__ CUtensorMap tmap, const int* __restrict__ idx,
                                                       int chunks, long long* out_cycles, uint4* dump) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar[kRing];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int lane = threadIdx.x;
  if (lane == 0) {
    for (int i = 0; i < kRing; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int* my = idx + (size_t)blockIdx.x * chunks * 128;
  long long t0, t1;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0));
  for (int c = 0; c < chunks + kRing; ++c) {
    if (c >= kRing) {   // wait for chunk c - kRing (frees its buffer)
      const int w = c - kRing;
      uint32_t done;
      do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(&bar[w % kRing])), "r"((uint32_t)((w / kRing) & 1)) : "memory");
      } while (!done);
      if (dump != nullptr && blockIdx.x == 0 && w == 0) {   // dump the first chunk as it sits in shared memory
        for (int i = lane; i < 1024; i += 32) {
          uint4 v;
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(base + i * 16));
          dump[i] = v;
        }
      }
      __syncwarp();
    }
    if (c < chunks) {
      const uint32_t b = smem_u32(&bar[c % kRing]);
      if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(16384u) : "memory");
      __syncwarp();
      const int4 r = *reinterpret_cast<const int4*>(my + (size_t)c * 128 + lane * 4);
      const uint32_t dst = base + (uint32_t)(c % kRing) * 16384u + (uint32_t)lane * 512u;
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
          ::"r"(dst), "l"(&tmap), "r"(0), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w), "r"(b)
          : "memory");
    }
  }
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1));
  if (lane == 0) out_cycles[blockIdx.x] = t1 - t0;
}

int main() {
  const int N = 131072, C = 64, chunks = 256, ctas = 148;
  std::vector<__nv_bfloat16> h((size_t)N * C);
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < C; ++c) h[(size_t)r * C + c] = __float2bfloat16((float)((r * 7 + c) % 251));
  __nv_bfloat16* d;
  cudaMalloc(&d, h.size() * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  EncodeTiled enc = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres);
  if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  CUtensorMap tmap;
  cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)N};
  cuuint64_t strides[1] = {(cuuint64_t)C * 2};
  cuuint32_t box[2] = {(cuuint32_t)C, 1};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)cr);
  if (cr != CUDA_SUCCESS) return 1;
  int* didx;
  long long* dcyc;
  uint4* ddump;
  cudaMalloc(&didx, (size_t)ctas * chunks * 128 * 4);
  cudaMalloc(&dcyc, ctas * 8);
  cudaMalloc(&ddump, 16384);
  const int smem = 1024 + kRing * 16384;
  cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[4] = {"consecutive rows (key-order-like)", "random rows", "consecutive, 50% absent (-1)", "random, 70% absent (-1)"};
  for (int pat = 0; pat < 4; ++pat) {
    std::vector<int> hidx((size_t)ctas * chunks * 128);
    srand(1234 + pat);
    for (int b = 0; b < ctas; ++b)
      for (int c = 0; c < chunks; ++c)
        for (int r = 0; r < 128; ++r) {
          int v = (pat == 0 || pat == 2) ? (int)(((size_t)b * 911 + (size_t)c * 131 + r) % N) : rand() % N;
          if (pat == 2 && (rand() % 100) < 50) v = -1;
          if (pat == 3 && (rand() % 100) < 70) v = -1;
          hidx[((size_t)b * chunks + c) * 128 + r] = v;
        }
    cudaMemcpy(didx, hidx.data(), hidx.size() * 4, cudaMemcpyHostToDevice);
    for (int it = 0; it < 2; ++it) gather_kernel<<<ctas, 32, smem>>>(tmap, didx, chunks, dcyc, pat == 0 && it == 1 ? ddump : nullptr);
    cudaError_t e = cudaDeviceSynchronize();
    long long hc[148];
    cudaMemcpy(hc, dcyc, sizeof(hc), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < ctas; ++i) s += (double)hc[i];
    printf("%-36s %7.1f cycles per 16 KB chunk (32 gather4)  -> %5.1f B/clk/SM  (%s)\n", names[pat], s / ctas / chunks,
           16384.0 / (s / ctas / chunks), cudaGetErrorString(e));
    if (pat == 0) {
      // placement check: chunk 0 of CTA 0, chunk row r = source row hidx[r]; canonical SW128: row r at (r>>3)*1024 +
      // (r&7)*128, 16-byte piece q at ((q ^ (r&7)) << 4)
      std::vector<uint4> dump(1024);
      cudaMemcpy(dump.data(), ddump, 16384, cudaMemcpyDeviceToHost);
      const uint8_t* bytes = (const uint8_t*)dump.data();
      int bad = 0;
      for (int r = 0; r < 128 && bad < 5; ++r)
        for (int q = 0; q < 8; ++q) {
          const __nv_bfloat16* got = (const __nv_bfloat16*)(bytes + (r >> 3) * 1024 + (r & 7) * 128 + ((q ^ (r & 7)) << 4));
          const int src = hidx[r];
          for (int e2 = 0; e2 < 8; ++e2)
            if (__bfloat162float(got[e2]) != __bfloat162float(h[(size_t)src * C + q * 8 + e2])) { ++bad; break; }
        }
      printf("  placement vs canonical K-major SWIZZLE_128B layout (rows 4l..4l+3 at +512*l): %s\n", bad ? "MISMATCH" : "matches");
    }
  }
  return 0;
}
