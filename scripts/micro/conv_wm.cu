// a7 — EXPERIMENT (COMB_CONV_NARROW=wm, not the default path): sparse convolution forward for the narrow levels
// (Cin, Cout in {16, 32}) as a warp-level gather + mma.sync kernel.  r1 result: parity-green but instruction-bound
// (ncu: ~1150 warp instructions per 16-row tile, issue slots 57 % busy, HMMA pipe 12-28 %), 45 / 87 us per 16x16 /
// 32x32 layer against 33 / 58 us for the tcgen05 kernel of conv_ts.cu — kept for A/B measurements only.
//
// Replaces the gather-GEMM-scatter of spconv's SubMConv3d / SparseConv3d forward
// (pcdet/models/backbones_3d/spconv_backbone.py:191-205, conv_input / conv1 / conv2 levels).
//
// Why not tcgen05 here (r1 measurements, profiles/r1_c_conv_ts_trace.txt): one tcgen05.mma (M=128, K=16) costs
// ~60-75 cycles whatever N is, so at N = 16 / 32 the tensor pipe does 8 / 16 cycles of math per instruction and a
// 128-row tile still pays the dense-K instruction count (27 taps, 70 % of them zero rows at level 1).  A warp that
// owns 16 output rows can instead skip every kernel offset none of its 16 rows has (warp-uniform test on a ballot),
// keeps A in registers straight from the global load (no staging in shared or tensor memory, no barriers) and runs
// m16n8k16 bf16 HMMA with fp32 accumulators.  The 64- and 128-channel levels stay on tcgen05 (conv_ts.cu).
//
// Fragment mapping (mma.m16n8k16, q = lane % 4, g = lane / 4): a lane loads CIN/2 contiguous bytes of feature rows
// g and g+8 (one LDG.64 / LDG.128 per row and offset), so the K order inside an offset is permuted: source channel
// e = (CIN/4)*q + 4*s + w  (k-step s, w = 0..3)  feeds A registers {a0a1 | a4a5}[w/2] of k-step s; the packed B
// fragments use the same permutation.  Output channels are permuted the same way: accumulator (n-tile j, column
// 2q+i) is real channel (COUT/4)*q + 2j + i, so a lane stores COUT/2 contiguous bytes per row.
#include <stdlib.h>
#include "common.cuh"
#include "conv_impl.cuh"

namespace comb {
namespace {

constexpr int kWarps = 8;
constexpr int kThreadsWM = kWarps * 32;

__device__ __forceinline__ void hmma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                           uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int CIN>
struct RowVec;
template <>
struct RowVec<16> {
  using T = uint2;
  static __device__ __forceinline__ T zero() { return make_uint2(0u, 0u); }
};
template <>
struct RowVec<32> {
  using T = uint4;
  static __device__ __forceinline__ T zero() { return make_uint4(0u, 0u, 0u, 0u); }
};

// A registers of k-step s from the lane's row pieces (lo = row g, hi = row g + 8)
template <int CIN>
__device__ __forceinline__ void a_regs(const typename RowVec<CIN>::T& lo, const typename RowVec<CIN>::T& hi, int s,
                                       uint32_t (&a)[4]);
template <>
__device__ __forceinline__ void a_regs<16>(const uint2& lo, const uint2& hi, int, uint32_t (&a)[4]) {
  a[0] = lo.x; a[1] = hi.x; a[2] = lo.y; a[3] = hi.y;
}
template <>
__device__ __forceinline__ void a_regs<32>(const uint4& lo, const uint4& hi, int s, uint32_t (&a)[4]) {
  a[0] = s ? lo.z : lo.x; a[1] = s ? hi.z : hi.x; a[2] = s ? lo.w : lo.y; a[3] = s ? hi.w : hi.y;
}


template <int CIN, int COUT>
__global__ void __launch_bounds__(kThreadsWM, 2) spconv_wm_kernel(ConvFwdArgs p) {
  constexpr int kBatch = CIN == 16 ? 4 : 2;   // kernel offsets gathered per batch (two batches in flight)
  constexpr int KS = CIN / 16;       // k-steps per kernel offset
  constexpr int NT = COUT / 8;       // n-tiles
  constexpr int NP = COUT / 16;      // n-tile pairs (one uint4 of B fragments each)
  constexpr int CPL = COUT / 4;      // output channels per lane
  using AV = typename RowVec<CIN>::T;
  using OV = typename RowVec<COUT>::T;
  extern __shared__ uint4 wsm[];     // [K][KS][NP][32 lanes] weights, then index tiles [2][kWarps][K][16] ints

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = lane & 3, g = lane >> 2;
  const int no = eff_n(p.no_max, p.no_dev);
  const int K = p.K;
  const int nvec = K * KS * NP * 32;
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.wpacked);
    for (int i = tid; i < nvec; i += kThreadsWM) wsm[i] = __ldg(src + i);
  }
  int* idx_all = reinterpret_cast<int*>(wsm + nvec);
  const int tile_ints = K * 16;
  __syncthreads();

  // per-lane epilogue constants: channels CPL*q .. CPL*q + CPL-1
  float bias[CPL], scale[CPL], shift[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    bias[i] = (p.epi & COMB_EPI_BIAS) ? __ldg(p.bias + CPL * q + i) : 0.0f;
    scale[i] = (p.epi & COMB_EPI_AFFINE) ? __ldg(p.scale + CPL * q + i) : 1.0f;
    shift[i] = (p.epi & COMB_EPI_AFFINE) ? __ldg(p.shift + CPL * q + i) : 0.0f;
  }

  const int ntiles = (no + 15) >> 4;
  const int* __restrict__ nbr = p.nbr;
  const size_t ld = (size_t)p.ld;
  const bool vec_ok = (p.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.nbr) & 15) == 0);
  const uint8_t* __restrict__ in = reinterpret_cast<const uint8_t*>(p.in) + (size_t)q * (CIN / 2);
  const int stride = (int)gridDim.x * kWarps;

  // index tile of 16 rows x K offsets -> shared memory (cp.async); rows >= no read as -1
  auto prefetch = [&](int tile, int* dst) {
    const int row0 = tile << 4;
    if (vec_ok && row0 + 16 <= no) {
      for (int i = lane; i < K * 4; i += 32) {
        const int k = i >> 2, part = i & 3;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + k * 16 + part * 4)),
                     "l"(nbr + (size_t)k * ld + row0 + part * 4)
                     : "memory");
      }
    } else {
      for (int i = lane; i < K * 16; i += 32) {
        const int k = i >> 4, r = i & 15;
        dst[i] = row0 + r < no ? __ldg(nbr + (size_t)k * ld + row0 + r) : -1;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int tile = (int)blockIdx.x * kWarps + warp;
  int buf = 0;
  if (tile < ntiles) prefetch(tile, idx_all + warp * tile_ints);
  for (; tile < ntiles; tile += stride, buf ^= 1) {
    const int row0 = tile << 4;
    const int* idx = idx_all + (buf * kWarps + warp) * tile_ints;
    if (tile + stride < ntiles) {
      prefetch(tile + stride, idx_all + ((buf ^ 1) * kWarps + warp) * tile_ints);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();

    // warp-uniform mask of the kernel offsets at least one of the 16 rows has (lane k looks at offset k)
    bool any = false;
    if (lane < K) {
      const int4* v = reinterpret_cast<const int4*>(idx + lane * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int4 t = v[i];
        any = any || ((t.x & t.y & t.z & t.w) >= 0);   // some entry has a clear sign bit
      }
    }
    unsigned mask = __ballot_sync(0xffffffffu, any);

    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = 0.0f;

    // residual rows (g and g+8) issued up front; consumed in the epilogue
    OV r_lo = RowVec<COUT>::zero(), r_hi = RowVec<COUT>::zero();
    if (p.epi & COMB_EPI_RESIDUAL) {
      const uint8_t* rb = reinterpret_cast<const uint8_t*>(p.residual) + (size_t)q * (COUT / 2);
      if (row0 + g < no) r_lo = __ldg(reinterpret_cast<const OV*>(rb + (size_t)(row0 + g) * (COUT * 2)));
      if (row0 + g + 8 < no) r_hi = __ldg(reinterpret_cast<const OV*>(rb + (size_t)(row0 + g + 8) * (COUT * 2)));
    }

    // batches of kBatch offsets, two batches in flight (A is computed while B's loads fly, and vice versa)
    int kA[kBatch], kB[kBatch];
    AV loA[kBatch], hiA[kBatch], loB[kBatch], hiB[kBatch];
    auto issue = [&](int (&kk)[kBatch], AV (&lo)[kBatch], AV (&hi)[kBatch]) {
#pragma unroll
      for (int b = 0; b < kBatch; ++b) {
        kk[b] = -1;
        lo[b] = RowVec<CIN>::zero();
        hi[b] = RowVec<CIN>::zero();
        if (mask) {                                  // warp-uniform
          const int k = __ffs(mask) - 1;
          mask &= mask - 1;
          kk[b] = k;
          const int rl = idx[k * 16 + g], rh = idx[k * 16 + g + 8];
          if (rl >= 0) lo[b] = __ldg(reinterpret_cast<const AV*>(in + (size_t)(uint32_t)rl * (CIN * 2)));
          if (rh >= 0) hi[b] = __ldg(reinterpret_cast<const AV*>(in + (size_t)(uint32_t)rh * (CIN * 2)));
        }
      }
    };
    auto compute = [&](const int (&kk)[kBatch], const AV (&lo)[kBatch], const AV (&hi)[kBatch]) {
#pragma unroll
      for (int b = 0; b < kBatch; ++b) {
        if (kk[b] >= 0) {                            // warp-uniform
#pragma unroll
          for (int s = 0; s < KS; ++s) {
            uint32_t a[4];
            a_regs<CIN>(lo[b], hi[b], s, a);
#pragma unroll
            for (int np = 0; np < NP; ++np) {
              const uint4 bw = wsm[((kk[b] * KS + s) * NP + np) * 32 + lane];
              hmma_16816(acc[2 * np], a[0], a[1], a[2], a[3], bw.x, bw.y);
              hmma_16816(acc[2 * np + 1], a[0], a[1], a[2], a[3], bw.z, bw.w);
            }
          }
        }
      }
    };
    issue(kA, loA, hiA);
    while (true) {
      if (kA[0] < 0) break;
      issue(kB, loB, hiB);
      compute(kA, loA, hiA);
      if (kB[0] < 0) break;
      issue(kA, loA, hiA);
      compute(kB, loB, hiB);
    }

    // epilogue: lane owns channels CPL*q + 2j + i of rows g (acc[j][0..1]) and g+8 (acc[j][2..3])
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int row = row0 + g + 8 * hh;
      if (row < no) {
        float f[CPL];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          f[2 * j] = acc[j][2 * hh];
          f[2 * j + 1] = acc[j][2 * hh + 1];
        }
#pragma unroll
        for (int i = 0; i < CPL; ++i) f[i] = fmaf(f[i] + bias[i], scale[i], shift[i]);
        if (p.epi & COMB_EPI_RESIDUAL) {
          const OV rv = hh ? r_hi : r_lo;
          const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
          for (int i = 0; i < CPL / 2; ++i) {
            const float2 t = __bfloat1622float2(r2[i]);
            f[2 * i] += t.x;
            f[2 * i + 1] += t.y;
          }
        }
        if (p.epi & COMB_EPI_RELU) {
#pragma unroll
          for (int i = 0; i < CPL; ++i) f[i] = fmaxf(f[i], 0.0f);
        }
        if (p.out_f32) {
          float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (size_t)row * COUT + CPL * q);
#pragma unroll
          for (int i = 0; i < CPL / 4; ++i) op[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
        } else {
          OV o;
          __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
          for (int i = 0; i < CPL / 2; ++i) o2[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
          *reinterpret_cast<OV*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * COUT + CPL * q) = o;
        }
      }
    }
    __syncwarp();    // every lane is done with idx[buf] before the next iteration prefetches into it
  }
}

// Packed weights: uint4 index ((k*KS + s)*NP + np)*32 + lane holds the B fragments of n-tiles 2np, 2np+1 for k-step s
// of kernel offset k: {b0b1, b2b3} of each tile.  Fragment element (k-slot, n = lane/4 =: m within n-tile j):
//   b0b1: source channels ci = (CIN/4)*q + 4*s + {0,1};  b2b3: ci = (CIN/4)*q + 4*s + {2,3}     (q = lane % 4)
//   logical column m of n-tile j = accumulator column; real output channel = (COUT/4)*(m/2) + 2j + (m%2)
template <int CIN, int COUT>
__global__ void __launch_bounds__(256) wm_pack_kernel(const float* __restrict__ w, int K, int Cin_real,
                                                       __nv_bfloat16* __restrict__ out) {
  constexpr int KS = CIN / 16, NP = COUT / 16;
  const int total = K * KS * NP * 32 * 8;     // bf16 elements
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int el = e & 7;                     // element inside the uint4: tile (el/4), reg ((el/2)%2), half (el%2)
    const int lane = (e >> 3) & 31;
    int rest = e >> 8;
    const int np = rest % NP;
    rest /= NP;
    const int s = rest % KS;
    const int k = rest / KS;
    const int q = lane & 3, m = lane >> 2;
    const int j = 2 * np + (el >> 2);
    const int ci = (CIN / 4) * q + 4 * s + 2 * ((el >> 1) & 1) + (el & 1);
    const int co = (COUT / 4) * (m >> 1) + 2 * j + (m & 1);
    float v = 0.0f;
    if (ci < Cin_real) v = w[((size_t)co * K + k) * Cin_real + ci];
    out[e] = __float2bfloat16(v);
  }
}

template <int CIN, int COUT>
int launch_wm(const ConvFwdArgs& p, cudaStream_t stream) {
  const size_t smem = (size_t)p.K * (CIN / 16) * (COUT / 16) * 512 + (size_t)2 * kWarps * p.K * 16 * 4;
  static thread_local DevOnce configured;   // per device: the attribute is a per-device property
  if (configured.first()) {
    COMB_CUDA(cudaFuncSetAttribute(spconv_wm_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  }
  COMB_CHECK_ARG(smem <= 100 * 1024, "comb_spconv_fwd_bf16: weight image %zu too large for the warp-MMA kernel", smem);
  const int ntiles = cdiv(p.no_max, 16);
  const int per_sm = 2;   // 128 registers x 256 threads: two blocks per SM
  int grid = cdiv(ntiles, kWarps);
  if (grid > sm_count() * per_sm) grid = sm_count() * per_sm;
  spconv_wm_kernel<CIN, COUT><<<grid, kThreadsWM, smem, stream>>>(p);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

}  // namespace

bool wm_supported(int Cin_p, int Cout) { return (Cin_p == 16 || Cin_p == 32) && (Cout == 16 || Cout == 32); }

int wm_fwd_bf16(const ConvFwdArgs& p, int Cin_p, int Cout, cudaStream_t stream) {
  if (Cin_p == 16 && Cout == 16) return launch_wm<16, 16>(p, stream);
  if (Cin_p == 16 && Cout == 32) return launch_wm<16, 32>(p, stream);
  if (Cin_p == 32 && Cout == 16) return launch_wm<32, 16>(p, stream);
  if (Cin_p == 32 && Cout == 32) return launch_wm<32, 32>(p, stream);
  set_error("comb_spconv_fwd_bf16: (%d,%d) not served by the warp-MMA kernel", Cin_p, Cout);
  return COMB_EINVAL;
}

int wm_pack_weight(const float* weight, int Cout, int K, int Cin, int Cin_p, void* wpacked, cudaStream_t stream) {
  __nv_bfloat16* out = (__nv_bfloat16*)wpacked;
  const int total = K * (Cin_p / 16) * (Cout / 16) * 256;
  const int grid = cdiv(total, 256);
  if (Cin_p == 16 && Cout == 16) wm_pack_kernel<16, 16><<<grid, 256, 0, stream>>>(weight, K, Cin, out);
  else if (Cin_p == 16 && Cout == 32) wm_pack_kernel<16, 32><<<grid, 256, 0, stream>>>(weight, K, Cin, out);
  else if (Cin_p == 32 && Cout == 16) wm_pack_kernel<32, 16><<<grid, 256, 0, stream>>>(weight, K, Cin, out);
  else if (Cin_p == 32 && Cout == 32) wm_pack_kernel<32, 32><<<grid, 256, 0, stream>>>(weight, K, Cin, out);
  else COMB_CHECK_ARG(false, "comb_spconv_pack_weight_bf16: (%d,%d) not served by the warp-MMA kernel", Cin_p, Cout);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

}  // namespace comb
