// a7/a8 — sparse convolution in fp32 "check mode" (CUDA cores, fixed summation order) and the
// backward kernels (dgrad / wgrad) used by the autograd path.
//
// Replaces the gather-GEMM-scatter inside spconv's SparseConvolution forward/backward (call sites
// pcdet/models/backbones_3d/spconv_backbone.py:12-15,38-45,191-232).  The kernels are
// OUTPUT-STATIONARY: a block owns 32 output rows, walks the kernel offsets k in ascending order,
// gathers the (<=32) contributing input rows of offset k into shared memory and contracts them
// with W_k.  There is no scatter and no atomic in fwd/dgrad; every output row is written once.
#include "common.cuh"

namespace comb {
namespace {

constexpr int kRows = 32;      // output rows per block
constexpr int kThreads = 256;

// generic weight addressing: w(oc, k, ic) = weight[oc*s_oc + k*s_k + ic*s_ic]
struct WStride {
  int s_oc, s_k, s_ic;
};

// OC = output channels of this contraction (Cout for fwd, Cin for dgrad), IC likewise.
__global__ void __launch_bounds__(kThreads) spconv_os_f32_kernel(
    const float* __restrict__ in, int IC, const float* __restrict__ weight, WStride ws, int K, int OC,
    const int* __restrict__ nbr, int ld, int no_max, const int* __restrict__ no_dev, int epi,
    const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift,
    const float* __restrict__ residual, float* __restrict__ out, int OCp /* pow2 >= OC, <= 128 */) {
  extern __shared__ float smem[];
  float* sW = smem;                        // [OC][IC+1]
  float* sA = smem + (size_t)OC * (IC + 1);  // [kRows][IC]
  __shared__ int sIdx[kRows];

  const int no = eff_n(no_max, no_dev);
  const int row0 = blockIdx.x * kRows;
  if (row0 >= no) return;
  const int tid = threadIdx.x;
  const int oc = tid % OCp;
  const bool oc_live = oc < OC;
  const int rstep = kThreads / OCp;  // rows handled per pass
  const int r_first = tid / OCp;
  constexpr int kMaxAcc = 16;        // kRows / rstep <= 16 when OCp <= 128
  float acc[kMaxAcc];
#pragma unroll
  for (int j = 0; j < kMaxAcc; ++j) acc[j] = 0.0f;

  for (int k = 0; k < K; ++k) {
    int my = -1;
    if (tid < kRows) {
      int o = row0 + tid;
      my = (o < no) ? __ldg(nbr + (size_t)k * ld + o) : -1;
      sIdx[tid] = my;
    }
    if (!__syncthreads_or(my >= 0)) continue;  // uniform: whole tile has no neighbour at this offset
    for (int e = tid; e < OC * IC; e += kThreads) {
      int c_o = e / IC, c_i = e - c_o * IC;
      sW[c_o * (IC + 1) + c_i] = __ldg(weight + (size_t)c_o * ws.s_oc + (size_t)k * ws.s_k + (size_t)c_i * ws.s_ic);
    }
    for (int e = tid; e < kRows * IC; e += kThreads) {
      int r = e / IC, c = e - r * IC;
      int i = sIdx[r];
      sA[e] = (i >= 0) ? __ldg(in + (size_t)i * IC + c) : 0.0f;
    }
    __syncthreads();
    const float* w = sW + (oc_live ? oc : 0) * (IC + 1);
#pragma unroll
    for (int j = 0; j < kMaxAcc; ++j) {
      int r = r_first + j * rstep;
      if (oc_live && r < kRows && sIdx[r] >= 0) {
        const float* a = sA + r * IC;
        float s = acc[j];
        for (int c = 0; c < IC; ++c) s = fmaf(a[c], w[c], s);
        acc[j] = s;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < kMaxAcc; ++j) {
    int r = r_first + j * rstep;
    int o = row0 + r;
    if (oc_live && r < kRows && o < no) {
      float v = acc[j];
      if (epi & COMB_EPI_BIAS) v += bias[oc];
      if (epi & COMB_EPI_AFFINE) v = fmaf(v, scale[oc], shift[oc]);
      if (epi & COMB_EPI_RESIDUAL) v += residual[(size_t)o * OC + oc];
      if (epi & COMB_EPI_RELU) v = fmaxf(v, 0.0f);
      out[(size_t)o * OC + oc] = v;
    }
  }
}

// wgrad: grid (row chunks, K).  Each block reduces its chunk into a register tile of dW_k and
// adds it to global memory with fp32 atomics (dW is zeroed by the host wrapper first).
constexpr int kWgChunk = 2048;

__global__ void __launch_bounds__(kThreads) spconv_wgrad_f32_kernel(
    const float* __restrict__ in, int Cin, const float* __restrict__ dout, int Cout, int K,
    const int* __restrict__ nbr, int ld, int no_max, const int* __restrict__ no_dev, float* __restrict__ dweight) {
  extern __shared__ float smem[];
  float* sA = smem;                 // [kRows][Cin]
  float* sG = smem + kRows * Cin;   // [kRows][Cout]
  __shared__ int sIdx[kRows];
  const int no = eff_n(no_max, no_dev);
  const int k = blockIdx.y;
  const int chunk0 = blockIdx.x * kWgChunk;
  if (chunk0 >= no) return;
  const int chunk1 = min(no, chunk0 + kWgChunk);
  const int tid = threadIdx.x;
  const int total = Cout * Cin;
  constexpr int kMaxAcc = 64;  // 128*128/256
  float acc[kMaxAcc];
#pragma unroll
  for (int j = 0; j < kMaxAcc; ++j) acc[j] = 0.0f;
  bool touched = false;

  for (int r0 = chunk0; r0 < chunk1; r0 += kRows) {
    int my = -1;
    if (tid < kRows) {
      int o = r0 + tid;
      my = (o < chunk1) ? __ldg(nbr + (size_t)k * ld + o) : -1;
      sIdx[tid] = my;
    }
    if (!__syncthreads_or(my >= 0)) continue;
    touched = true;
    for (int e = tid; e < kRows * Cin; e += kThreads) {
      int r = e / Cin, c = e - r * Cin;
      int i = sIdx[r];
      sA[e] = (i >= 0) ? __ldg(in + (size_t)i * Cin + c) : 0.0f;
    }
    for (int e = tid; e < kRows * Cout; e += kThreads) {
      int r = e / Cout, c = e - r * Cout;
      sG[e] = (sIdx[r] >= 0) ? __ldg(dout + (size_t)(r0 + r) * Cout + c) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kMaxAcc; ++j) {
      int e = tid + j * kThreads;
      if (e < total) {
        int co = e / Cin, ci = e - co * Cin;
        float s = acc[j];
#pragma unroll 8
        for (int r = 0; r < kRows; ++r) s = fmaf(sG[r * Cout + co], sA[r * Cin + ci], s);
        acc[j] = s;
      }
    }
    __syncthreads();
  }
  if (!touched) return;
#pragma unroll
  for (int j = 0; j < kMaxAcc; ++j) {
    int e = tid + j * kThreads;
    if (e < total) {
      int co = e / Cin, ci = e - co * Cin;
      atomicAdd(dweight + ((size_t)co * K + k) * Cin + ci, acc[j]);
    }
  }
}

static int pow2_ge(int c) {
  int p = 1;
  while (p < c) p <<= 1;
  return p;
}

static int launch_os(const float* in, int IC, const float* weight, WStride ws, int K, int OC, const int* nbr, int ld,
                     int no_max, const int* no_dev, int epi, const float* bias, const float* scale, const float* shift,
                     const float* residual, float* out, cudaStream_t stream) {
  size_t smem = ((size_t)OC * (IC + 1) + (size_t)kRows * IC) * sizeof(float);
  static thread_local size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    COMB_CUDA(cudaFuncSetAttribute(spconv_os_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = 160 * 1024;
  }
  spconv_os_f32_kernel<<<cdiv(no_max, kRows), kThreads, smem, stream>>>(in, IC, weight, ws, K, OC, nbr, ld, no_max,
                                                                         no_dev, epi, bias, scale, shift, residual,
                                                                         out, pow2_ge(OC));
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

}  // namespace
}  // namespace comb

using namespace comb;

extern "C" int comb_spconv_fwd_f32(const float* in_feats, int Cin, const float* weight, int K, int Cout,
                                   const int* nbr, int ld, int no_max, const int* no_dev, int epi_flags,
                                   const float* bias, const float* scale, const float* shift, const float* residual,
                                   float* out, void* stream_) {
  COMB_CHECK_ARG(Cin >= 1 && Cin <= 128, "comb_spconv_fwd_f32: Cin %d outside [1,128]", Cin);
  COMB_CHECK_ARG(Cout >= 1 && Cout <= 128, "comb_spconv_fwd_f32: Cout %d outside [1,128]", Cout);
  COMB_CHECK_ARG(K >= 1 && ld >= no_max && no_max >= 0, "comb_spconv_fwd_f32: bad K/ld/no");
  COMB_CHECK_ARG(!(epi_flags & COMB_EPI_BIAS) || bias, "comb_spconv_fwd_f32: bias flag without pointer");
  COMB_CHECK_ARG(!(epi_flags & COMB_EPI_AFFINE) || (scale && shift), "comb_spconv_fwd_f32: affine flag without pointers");
  COMB_CHECK_ARG(!(epi_flags & COMB_EPI_RESIDUAL) || residual, "comb_spconv_fwd_f32: residual flag without pointer");
  if (no_max == 0) return COMB_OK;
  COMB_CHECK_ARG(in_feats && weight && nbr && out, "comb_spconv_fwd_f32: null pointer");
  WStride ws{K * Cin, Cin, 1};
  return launch_os(in_feats, Cin, weight, ws, K, Cout, nbr, ld, no_max, no_dev, epi_flags, bias, scale, shift,
                   residual, out, (cudaStream_t)stream_);
}

extern "C" int comb_spconv_dgrad_f32(const float* dout, int Cout, const float* weight, int K, int Cin,
                                     const int* nbr_t, int ld_t, int ni_max, const int* ni_dev, float* din,
                                     void* stream_) {
  COMB_CHECK_ARG(Cout >= 1 && Cout <= 128, "comb_spconv_dgrad_f32: Cout %d outside [1,128]", Cout);
  COMB_CHECK_ARG(Cin >= 1 && Cin <= 128, "comb_spconv_dgrad_f32: Cin %d outside [1,128]", Cin);
  COMB_CHECK_ARG(K >= 1 && ld_t >= ni_max && ni_max >= 0, "comb_spconv_dgrad_f32: bad K/ld/ni");
  if (ni_max == 0) return COMB_OK;
  COMB_CHECK_ARG(dout && weight && nbr_t && din, "comb_spconv_dgrad_f32: null pointer");
  // contraction over co: "output channel" = ci (stride 1), "input channel" = co (stride K*Cin)
  WStride ws{1, Cin, K * Cin};
  return launch_os(dout, Cout, weight, ws, K, Cin, nbr_t, ld_t, ni_max, ni_dev, 0, nullptr, nullptr, nullptr, nullptr,
                   din, (cudaStream_t)stream_);
}

extern "C" int comb_spconv_wgrad_f32(const float* in_feats, int Cin, const float* dout, int Cout, int K,
                                     const int* nbr, int ld, int no_max, const int* no_dev, float* dweight,
                                     void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(Cin >= 1 && Cin <= 128 && Cout >= 1 && Cout <= 128, "comb_spconv_wgrad_f32: channels outside [1,128]");
  COMB_CHECK_ARG(K >= 1 && ld >= no_max && no_max >= 0 && dweight, "comb_spconv_wgrad_f32: bad arguments");
  COMB_CUDA(cudaMemsetAsync(dweight, 0, (size_t)Cout * K * Cin * sizeof(float), stream));
  if (no_max == 0) return COMB_OK;
  COMB_CHECK_ARG(in_feats && dout && nbr, "comb_spconv_wgrad_f32: null pointer");
  size_t smem = (size_t)kRows * (Cin + Cout) * sizeof(float);
  dim3 grid(cdiv(no_max, kWgChunk), K);
  spconv_wgrad_f32_kernel<<<grid, kThreads, smem, stream>>>(in_feats, Cin, dout, Cout, K, nbr, ld, no_max, no_dev,
                                                            dweight);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}
