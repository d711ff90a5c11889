// a9, training form — nn.BatchNorm1d in TRAIN mode over the active rows of a sparse tensor, fused with the ReLU and the
// residual add that follow it, forward and backward.
//
// Replaces, for the fused train-mode step, the eager chain the reference runs between two sparse convolutions
//   SparseSequential(conv, BatchNorm1d(eps=1e-3, momentum=0.01), ReLU)        pcdet/models/backbones_3d/spconv_backbone.py:21-25
//   SparseBasicBlock.forward: bn1 -> relu -> conv2 -> bn2 -> (+ identity) -> relu                          ...:50-66
// (three to five elementwise / reduction kernels of torch per convolution, each a full pass over N x C).
//
// Forward   mean/var over rows (two-pass deterministic reduction: per-block partial sums in fp64, one finishing block),
//           running statistics updated like torch (momentum, unbiased variance), out = relu(x*scale + shift (+ res)).
// Backward  g = dy * (out > 0);  dgamma = sum g*xhat;  dbeta = sum g;
//           dx = gamma*invstd * (g - dbeta/n - xhat*dgamma/n);  the residual branch receives g.
// Activations and gradients are stored in bf16 (what the tensor-core convolutions consume), statistics in fp32/fp64.
#include "common.cuh"

namespace comb {
namespace {

constexpr int kBnThreads = 256;
constexpr int kBnMaxBlocks = 592;   // 4 per SM

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

// 8 consecutive channels of element group e of the convolution output (bf16: one 16-byte load, fp32: two)
template <typename XT>
__device__ __forceinline__ void load8(const XT* __restrict__ x, size_t e, float (&f)[8]);
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* __restrict__ x, size_t e, float (&f)[8]) {
  unpack8(__ldg(reinterpret_cast<const uint4*>(x) + e), f);
}
template <>
__device__ __forceinline__ void load8<float>(const float* __restrict__ x, size_t e, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(x) + 2 * e), b = __ldg(reinterpret_cast<const float4*>(x) + 2 * e + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// Column reductions.  A thread owns one 8-channel group (16 bytes) and walks rows with a grid stride; MODE 0: sum x and
// sum x^2 (forward statistics, also the plain column sum for the bias gradient); MODE 1: sum g and sum g*xhat (backward).
// part[block][2][C] in fp64; blocks that own no rows write zeros.
template <int MODE, typename XT>
__global__ void __launch_bounds__(kBnThreads) bn_reduce_kernel(const XT* __restrict__ x, const uint4* __restrict__ act,
                                                                const uint4* __restrict__ dy, int n_max,
                                                                const int* __restrict__ n_dev, int C,
                                                                const float* __restrict__ mean,
                                                                const float* __restrict__ invstd, int relu,
                                                                double* __restrict__ part) {
  // fp64 accumulators from the first product on: var = E[x^2] - mean^2 cancels, and BatchNorm of a near-constant
  // channel (var << mean^2, eps = 1e-3) must not see fp32 rounding of x^2 (B200 has full-rate-enough fp64 for a
  // memory-bound pass)
  __shared__ double red[2][kBnThreads][8];
  const int n = eff_n(n_max, n_dev);
  const int groups = C >> 3;                        // threads per row
  const int rows_per_iter = kBnThreads / groups;
  const int g = threadIdx.x % groups, r0 = threadIdx.x / groups;
  double a[8], b[8];
  float mu[8], is[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = b[i] = 0.0;
    mu[i] = MODE == 1 ? __ldg(mean + g * 8 + i) : 0.f;
    is[i] = MODE == 1 ? __ldg(invstd + g * 8 + i) : 0.f;
  }
  for (long long row = (long long)blockIdx.x * rows_per_iter + r0; row < n; row += (long long)gridDim.x * rows_per_iter) {
    const size_t e = (size_t)row * groups + g;
    float xv[8];
    load8<XT>(x, e, xv);
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a[i] += (double)xv[i];
        b[i] += (double)xv[i] * (double)xv[i];
      }
    } else {
      float gv[8], ov[8];
      unpack8(__ldg(dy + e), gv);
      if (relu) {
        unpack8(__ldg(act + e), ov);
#pragma unroll
        for (int i = 0; i < 8; ++i) gv[i] = ov[i] > 0.f ? gv[i] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a[i] += (double)gv[i];
        b[i] += (double)(gv[i] * ((xv[i] - mu[i]) * is[i]));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[0][threadIdx.x][i] = a[i];
    red[1][threadIdx.x][i] = b[i];
  }
  __syncthreads();
  // thread t < 2*C finishes one (which, channel) column in fp64
  if (threadIdx.x < 2 * C) {
    const int which = threadIdx.x / C, c = threadIdx.x % C;
    const int cg = c >> 3, ci = c & 7;
    double s = 0.0;
    for (int r = 0; r < rows_per_iter; ++r) s += red[which][r * groups + cg][ci];
    part[((size_t)blockIdx.x * 2 + which) * C + c] = s;
  }
}

// Sum of the per-block partials for channel c = threadIdx.x % C by ALL threads of a 1024-thread block: slice s =
// threadIdx.x / C adds blocks s, s + nsl, ..., then a fixed-shape tree over the slices (deterministic).  Returns the
// totals to the threads of slice 0.  (r2: the first version walked up to 592 partials serially in C threads — 53 us per
// call, 2.2 ms of a 6.5 ms training step.)
constexpr int kFinThreads = 1024;
__device__ __forceinline__ void sum_partials(const double* __restrict__ part, int nblocks, int C, double& s, double& q) {
  __shared__ double sh[2][kFinThreads];
  const int c = threadIdx.x % C, slice = threadIdx.x / C, nsl = kFinThreads / C;
  s = 0.0;
  q = 0.0;
  for (int b = slice; b < nblocks; b += nsl) {
    s += part[((size_t)b * 2 + 0) * C + c];
    q += part[((size_t)b * 2 + 1) * C + c];
  }
  sh[0][threadIdx.x] = s;
  sh[1][threadIdx.x] = q;
  __syncthreads();
  for (int stride = nsl >> 1; stride > 0; stride >>= 1) {
    if (slice < stride) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + stride * C];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + stride * C];
    }
    __syncthreads();
  }
  s = sh[0][c];
  q = sh[1][c];
}

// Finishing block of the forward statistics.
__global__ void __launch_bounds__(kFinThreads) bn_fwd_finalize_kernel(const double* __restrict__ part, int nblocks, int n_max,
                                                               const int* __restrict__ n_dev, int C,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float eps, float momentum,
                                                               float* __restrict__ running_mean,
                                                               float* __restrict__ running_var, float* __restrict__ save_mean,
                                                               float* __restrict__ save_invstd, float* __restrict__ scale,
                                                               float* __restrict__ shift) {
  double s, q;
  sum_partials(part, nblocks, C, s, q);
  const int c = threadIdx.x;
  if (c >= C) return;
  const int n = eff_n(n_max, n_dev);
  const double inv_n = n > 0 ? 1.0 / (double)n : 0.0;
  const double mu = s * inv_n;
  double var = q * inv_n - mu * mu;
  if (var < 0.0) var = 0.0;
  const double is = 1.0 / sqrt(var + (double)eps);
  save_mean[c] = (float)mu;
  save_invstd[c] = (float)is;
  const float sc = (float)((double)__ldg(gamma + c) * is);
  scale[c] = sc;
  shift[c] = (float)((double)__ldg(beta + c) - mu * (double)__ldg(gamma + c) * is);
  if (running_mean != nullptr && n > 0) {           // torch: running = (1-m)*running + m*batch, unbiased batch variance
    const double unbiased = n > 1 ? var * (double)n / (double)(n - 1) : var;
    running_mean[c] = (float)((1.0 - (double)momentum) * (double)running_mean[c] + (double)momentum * mu);
    running_var[c] = (float)((1.0 - (double)momentum) * (double)running_var[c] + (double)momentum * unbiased);
  }
}

// out = relu(x*scale + shift (+ residual)), 8 channels per thread
template <typename XT>
__global__ void __launch_bounds__(kBnThreads) bn_apply_kernel(const XT* __restrict__ x, int n_max,
                                                               const int* __restrict__ n_dev, int C,
                                                               const float* __restrict__ scale,
                                                               const float* __restrict__ shift,
                                                               const uint4* __restrict__ residual, int relu,
                                                               uint4* __restrict__ out) {
  const int groups = C >> 3;
  const long long total = (long long)eff_n(n_max, n_dev) * groups;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(e % groups);
    float v[8];
    load8<XT>(x, (size_t)e, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], __ldg(scale + g * 8 + i), __ldg(shift + g * 8 + i));
    if (residual != nullptr) {
      float r[8];
      unpack8(__ldg(residual + e), r);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += r[i];
    }
    if (relu) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    out[e] = pack8(v);
  }
}

// dgamma, dbeta and the two per-channel coefficients of the input gradient
__global__ void __launch_bounds__(kFinThreads) bn_bwd_finalize_kernel(const double* __restrict__ part, int nblocks, int n_max,
                                                               const int* __restrict__ n_dev, int C,
                                                               float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                               float* __restrict__ coef) {
  double s, q;
  sum_partials(part, nblocks, C, s, q);
  const int c = threadIdx.x;
  if (c >= C) return;
  const int n = eff_n(n_max, n_dev);
  dbeta[c] = (float)s;
  dgamma[c] = (float)q;
  const double inv_n = n > 0 ? 1.0 / (double)n : 0.0;
  coef[c] = (float)(s * inv_n);
  coef[C + c] = (float)(q * inv_n);
}

// dx = gamma*invstd*(g - a - xhat*b), g = dy * (act > 0); optionally also stores g (gradient of the residual branch)
template <typename XT>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_apply_kernel(const XT* __restrict__ x, const uint4* __restrict__ act,
                                                                   const uint4* __restrict__ dy, int n_max,
                                                                   const int* __restrict__ n_dev, int C,
                                                                   const float* __restrict__ gamma,
                                                                   const float* __restrict__ mean,
                                                                   const float* __restrict__ invstd,
                                                                   const float* __restrict__ coef, int relu,
                                                                   uint4* __restrict__ dx, uint4* __restrict__ g_out) {
  const int groups = C >> 3;
  const long long total = (long long)eff_n(n_max, n_dev) * groups;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(e % groups);
    float xv[8], gv[8];
    load8<XT>(x, (size_t)e, xv);
    unpack8(__ldg(dy + e), gv);
    if (relu) {
      float ov[8];
      unpack8(__ldg(act + e), ov);
#pragma unroll
      for (int i = 0; i < 8; ++i) gv[i] = ov[i] > 0.f ? gv[i] : 0.f;
    }
    if (g_out != nullptr) g_out[e] = pack8(gv);
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = g * 8 + i;
      const float is = __ldg(invstd + c);
      const float xhat = (xv[i] - __ldg(mean + c)) * is;
      o[i] = __ldg(gamma + c) * is * (gv[i] - __ldg(coef + c) - xhat * __ldg(coef + C + c));
    }
    dx[e] = pack8(o);
  }
}

__global__ void __launch_bounds__(kFinThreads) colsum_finalize_kernel(const double* __restrict__ part, int nblocks, int C,
                                                               float* __restrict__ sum) {
  double s, q;
  sum_partials(part, nblocks, C, s, q);
  if (threadIdx.x < C) sum[threadIdx.x] = (float)s;
}

int reduce_blocks(int n_max, int C) {
  const int rows_per_iter = kBnThreads / (C >> 3);
  int b = cdiv(n_max > 0 ? n_max : 1, rows_per_iter);
  return b < kBnMaxBlocks ? b : kBnMaxBlocks;
}
int ew_blocks(long long total) {
  long long b = (total + kBnThreads - 1) / kBnThreads;
  if (b < 1) b = 1;
  return (int)(b < 148 * 16 ? b : 148 * 16);
}
bool bn_c_ok(int C) { return C == 16 || C == 32 || C == 64 || C == 128; }

}  // namespace
}  // namespace comb

using namespace comb;

extern "C" size_t comb_bn_workspace_bytes(int C) {
  if (!bn_c_ok(C)) return 0;
  return (size_t)kBnMaxBlocks * 2 * C * sizeof(double) + (size_t)4 * C * sizeof(float);
}

extern "C" int comb_bn_train_fwd(const void* x, int x_dtype, int n_max, const int* n_dev, int C, const float* gamma, const float* beta,
                                 float eps, float momentum, float* running_mean, float* running_var, const void* residual,
                                 int relu, void* out, float* save_mean, float* save_invstd, void* workspace,
                                 size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(bn_c_ok(C), "comb_bn_train_fwd: C %d not in {16,32,64,128}", C);
  COMB_CHECK_ARG(x_dtype == COMB_DT_F32 || x_dtype == COMB_DT_BF16, "comb_bn_train_fwd: bad x dtype");
  COMB_CHECK_ARG(n_max >= 0, "comb_bn_train_fwd: bad n");
  COMB_CHECK_ARG(workspace && workspace_bytes >= comb_bn_workspace_bytes(C), "comb_bn_train_fwd: workspace too small");
  COMB_CHECK_ARG(gamma && beta && save_mean && save_invstd, "comb_bn_train_fwd: null pointer");
  COMB_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "comb_bn_train_fwd: running stats come in pairs");
  if (n_max == 0) return COMB_OK;
  COMB_CHECK_ARG(x && out, "comb_bn_train_fwd: null x/out");
  double* part = (double*)workspace;
  float* scale = (float*)(part + (size_t)kBnMaxBlocks * 2 * C);
  float* shift = scale + C;
  const int nb = reduce_blocks(n_max, C);
  if (x_dtype == COMB_DT_F32)
    bn_reduce_kernel<0, float><<<nb, kBnThreads, 0, stream>>>((const float*)x, nullptr, nullptr, n_max, n_dev, C, nullptr,
                                                               nullptr, 0, part);
  else
    bn_reduce_kernel<0, __nv_bfloat16><<<nb, kBnThreads, 0, stream>>>((const __nv_bfloat16*)x, nullptr, nullptr, n_max,
                                                                       n_dev, C, nullptr, nullptr, 0, part);
  COMB_LAUNCH_CHECK();
  bn_fwd_finalize_kernel<<<1, kFinThreads, 0, stream>>>(part, nb, n_max, n_dev, C, gamma, beta, eps, momentum, running_mean,
                                                running_var, save_mean, save_invstd, scale, shift);
  COMB_LAUNCH_CHECK();
  const int eb = ew_blocks((long long)n_max * (C >> 3));
  if (x_dtype == COMB_DT_F32)
    bn_apply_kernel<float><<<eb, kBnThreads, 0, stream>>>((const float*)x, n_max, n_dev, C, scale, shift,
                                                           (const uint4*)residual, relu, (uint4*)out);
  else
    bn_apply_kernel<__nv_bfloat16><<<eb, kBnThreads, 0, stream>>>((const __nv_bfloat16*)x, n_max, n_dev, C, scale, shift,
                                                                   (const uint4*)residual, relu, (uint4*)out);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

extern "C" int comb_bn_train_bwd(const void* dy, const void* act, const void* x, int x_dtype, int n_max, const int* n_dev, int C,
                                 const float* gamma, const float* save_mean, const float* save_invstd, int relu, void* dx,
                                 void* g_out, float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes,
                                 void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(bn_c_ok(C), "comb_bn_train_bwd: C %d not in {16,32,64,128}", C);
  COMB_CHECK_ARG(x_dtype == COMB_DT_F32 || x_dtype == COMB_DT_BF16, "comb_bn_train_bwd: bad x dtype");
  COMB_CHECK_ARG(n_max >= 0, "comb_bn_train_bwd: bad n");
  COMB_CHECK_ARG(workspace && workspace_bytes >= comb_bn_workspace_bytes(C), "comb_bn_train_bwd: workspace too small");
  COMB_CHECK_ARG(gamma && save_mean && save_invstd && dgamma && dbeta, "comb_bn_train_bwd: null pointer");
  COMB_CHECK_ARG(!relu || act, "comb_bn_train_bwd: relu needs the forward output");
  double* part = (double*)workspace;
  float* coef = (float*)(part + (size_t)kBnMaxBlocks * 2 * C);
  if (n_max == 0) {
    COMB_CUDA(cudaMemsetAsync(dgamma, 0, (size_t)C * 4, stream));
    COMB_CUDA(cudaMemsetAsync(dbeta, 0, (size_t)C * 4, stream));
    return COMB_OK;
  }
  COMB_CHECK_ARG(dy && x && dx, "comb_bn_train_bwd: null dy/x/dx");
  const int nb = reduce_blocks(n_max, C);
  if (x_dtype == COMB_DT_F32)
    bn_reduce_kernel<1, float><<<nb, kBnThreads, 0, stream>>>((const float*)x, (const uint4*)act, (const uint4*)dy, n_max,
                                                               n_dev, C, save_mean, save_invstd, relu, part);
  else
    bn_reduce_kernel<1, __nv_bfloat16><<<nb, kBnThreads, 0, stream>>>((const __nv_bfloat16*)x, (const uint4*)act,
                                                                       (const uint4*)dy, n_max, n_dev, C, save_mean,
                                                                       save_invstd, relu, part);
  COMB_LAUNCH_CHECK();
  bn_bwd_finalize_kernel<<<1, kFinThreads, 0, stream>>>(part, nb, n_max, n_dev, C, dgamma, dbeta, coef);
  COMB_LAUNCH_CHECK();
  const int eb = ew_blocks((long long)n_max * (C >> 3));
  if (x_dtype == COMB_DT_F32)
    bn_bwd_apply_kernel<float><<<eb, kBnThreads, 0, stream>>>((const float*)x, (const uint4*)act, (const uint4*)dy, n_max,
                                                               n_dev, C, gamma, save_mean, save_invstd, coef, relu,
                                                               (uint4*)dx, (uint4*)g_out);
  else
    bn_bwd_apply_kernel<__nv_bfloat16><<<eb, kBnThreads, 0, stream>>>((const __nv_bfloat16*)x, (const uint4*)act,
                                                                       (const uint4*)dy, n_max, n_dev, C, gamma, save_mean,
                                                                       save_invstd, coef, relu, (uint4*)dx, (uint4*)g_out);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

extern "C" int comb_col_sum(const void* x, int n_max, const int* n_dev, int C, float* sum, void* workspace,
                            size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(bn_c_ok(C), "comb_col_sum: C %d not in {16,32,64,128}", C);
  COMB_CHECK_ARG(workspace && workspace_bytes >= comb_bn_workspace_bytes(C) && sum, "comb_col_sum: bad workspace / sum");
  if (n_max == 0) {
    COMB_CUDA(cudaMemsetAsync(sum, 0, (size_t)C * 4, stream));
    return COMB_OK;
  }
  COMB_CHECK_ARG(x, "comb_col_sum: null x");
  double* part = (double*)workspace;
  const int nb = reduce_blocks(n_max, C);
  bn_reduce_kernel<0, __nv_bfloat16><<<nb, kBnThreads, 0, stream>>>((const __nv_bfloat16*)x, nullptr, nullptr, n_max, n_dev,
                                                                     C, nullptr, nullptr, 0, part);
  COMB_LAUNCH_CHECK();
  colsum_finalize_kernel<<<1, kFinThreads, 0, stream>>>(part, nb, C, sum);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}
