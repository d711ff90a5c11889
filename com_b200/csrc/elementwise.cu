// a9/a10 — per-row elementwise epilogues between convolutions and the dense BEV scatter.
//
// Replaces
//   nn.BatchNorm1d(eval) + nn.ReLU + residual add on `.features`   pcdet/models/backbones_3d/spconv_backbone.py:21-25,50-66
//   SparseConvTensor.dense() used by HeightCompression.forward     pcdet/models/backbones_2d/map_to_bev/height_compression.py:21-23
#include "common.cuh"

namespace comb {
namespace {

template <typename T>
__device__ __forceinline__ float ld_as_float(const T* p, size_t i);
template <>
__device__ __forceinline__ float ld_as_float<float>(const float* p, size_t i) { return __ldg(p + i); }
template <>
__device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p, size_t i) {
  return __bfloat162float(p[i]);
}
template <typename T>
__device__ __forceinline__ void st_from_float(T* p, size_t i, float v);
template <>
__device__ __forceinline__ void st_from_float<float>(float* p, size_t i, float v) { p[i] = v; }
template <>
__device__ __forceinline__ void st_from_float<__nv_bfloat16>(__nv_bfloat16* p, size_t i, float v) {
  p[i] = __float2bfloat16(v);
}

template <typename T>
__global__ void __launch_bounds__(256) affine_relu_kernel(const T* __restrict__ x, int n_max,
                                                           const int* __restrict__ n_dev, int C,
                                                           const float* __restrict__ scale,
                                                           const float* __restrict__ shift,
                                                           const T* __restrict__ residual, int relu,
                                                           T* __restrict__ out) {
  const long long n = (long long)eff_n(n_max, n_dev) * C;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int c = (int)(e % C);
    float v = ld_as_float(x, e);
    if (scale) v = fmaf(v, __ldg(scale + c), __ldg(shift + c));
    if (residual) v += ld_as_float(residual, e);
    if (relu) v = fmaxf(v, 0.0f);
    st_from_float(out, e, v);
  }
}

__global__ void __launch_bounds__(256) cast_pad_kernel(const float* __restrict__ x, int n_max,
                                                        const int* __restrict__ n_dev, int C,
                                                        __nv_bfloat16* __restrict__ out, int ld) {
  const long long n = (long long)eff_n(n_max, n_dev) * ld;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int row = (int)(e / ld), c = (int)(e - (long long)row * ld);
    out[e] = __float2bfloat16(c < C ? __ldg(x + (size_t)row * C + c) : 0.0f);
  }
}

// one thread per 4-byte word of a row
__global__ void __launch_bounds__(256) permute_rows_kernel(const uint32_t* __restrict__ in,
                                                            const int* __restrict__ row_map, int n_max,
                                                            const int* __restrict__ n_dev, int row_words, int scatter,
                                                            uint32_t* __restrict__ out) {
  const long long n = (long long)eff_n(n_max, n_dev) * row_words;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / row_words), w = (int)(e - (long long)r * row_words);
    const int m = __ldg(row_map + r);
    if (scatter) {
      if (m >= 0) out[(size_t)m * row_words + w] = __ldg(in + e);
    } else {
      out[e] = m >= 0 ? __ldg(in + (size_t)m * row_words + w) : 0u;
    }
  }
}

__global__ void __launch_bounds__(256) dense_index_kernel(const int4* __restrict__ coords, int n_max,
                                                           const int* __restrict__ n_dev, int batch, int D, int H,
                                                           int W, int* __restrict__ cell_row) {
  const int n = eff_n(n_max, n_dev);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int4 c = __ldg(coords + i);
    if (c.x < 0 || c.x >= batch || c.y < 0 || c.y >= D || c.z < 0 || c.z >= H || c.w < 0 || c.w >= W) continue;
    cell_row[(((size_t)c.x * D + c.y) * H + c.z) * W + c.w] = i;
  }
}

// Writes the WHOLE dense tensor: a block owns 32 consecutive cells of one frame (contiguous along x
// in NCDHW for every channel), stages the present feature rows in shared memory (row-major, read
// coalesced) and stores channel-major 128-byte lines.
constexpr int kCells = 32;

template <typename T>
__global__ void __launch_bounds__(256) dense_write_kernel(const T* __restrict__ feats,
                                                           const int* __restrict__ cell_row, int C, int DHW,
                                                           int tiles_per_frame, float* __restrict__ out) {
  extern __shared__ float tile[];  // [kCells][C+1]
  __shared__ int rows[kCells];
  const int b = blockIdx.x / tiles_per_frame;
  const int cell0 = (blockIdx.x - b * tiles_per_frame) * kCells;
  const int tid = threadIdx.x;
  int my = -1;
  if (tid < kCells) {
    int cell = cell0 + tid;
    my = (cell < DHW) ? cell_row[(size_t)b * DHW + cell] : -1;
    rows[tid] = my;
  }
  const int any = __syncthreads_or(my >= 0);
  if (any) {
    for (int e = tid; e < kCells * C; e += 256) {
      int r = e / C, c = e - r * C;
      int row = rows[r];
      tile[r * (C + 1) + c] = row >= 0 ? ld_as_float(feats, (size_t)row * C + c) : 0.0f;
    }
    __syncthreads();
  }
  const int lane = tid & 31, warp = tid >> 5;
  const int cell = cell0 + lane;
  if (cell >= DHW) return;
  float* o = out + (size_t)b * C * DHW + cell;
  for (int c = warp; c < C; c += 8) o[(size_t)c * DHW] = any ? tile[lane * (C + 1) + c] : 0.0f;
}

// Vector variant (needs DHW % 4 == 0): a block owns 128 consecutive cells of one frame = 512 contiguous bytes
// per channel.  Lane l of every warp owns cells 4l..4l+3 (their feature rows are read once, as one int4),
// warp w writes channels w, w+8, ... with one 16-byte streaming store per lane: the whole tensor is written
// exactly once in full 512-byte segments; feature rows of occupied cells are re-read from L1 (2 lines per row).
constexpr int kCellsV = 128;

template <typename T>
__global__ void __launch_bounds__(256) dense_write_vec_kernel(const T* __restrict__ feats,
                                                               const int* __restrict__ cell_row, int C, int DHW,
                                                               int tiles_per_frame, float* __restrict__ out) {
  const int b = blockIdx.x / tiles_per_frame;
  const int cell0 = (blockIdx.x - b * tiles_per_frame) * kCellsV;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cell = cell0 + lane * 4;
  if (cell >= DHW) return;   // DHW % 4 == 0: a lane's four cells are all inside or all outside
  const int4 r = __ldg(reinterpret_cast<const int4*>(cell_row + (size_t)b * DHW + cell));
  float* o = out + (size_t)b * C * DHW + cell;
  if ((r.x & r.y & r.z & r.w) < 0 && r.x < 0 && r.y < 0 && r.z < 0 && r.w < 0) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = warp; c < C; c += 8) __stcs(reinterpret_cast<float4*>(o + (size_t)c * DHW), z);
    return;
  }
  for (int c = warp; c < C; c += 8) {
    float4 v;
    v.x = r.x >= 0 ? ld_as_float(feats, (size_t)r.x * C + c) : 0.0f;
    v.y = r.y >= 0 ? ld_as_float(feats, (size_t)r.y * C + c) : 0.0f;
    v.z = r.z >= 0 ? ld_as_float(feats, (size_t)r.z * C + c) : 0.0f;
    v.w = r.w >= 0 ? ld_as_float(feats, (size_t)r.w * C + c) : 0.0f;
    __stcs(reinterpret_cast<float4*>(o + (size_t)c * DHW), v);
  }
}

static int ew_grid(long long work) {
  int g = cdiv(work, 256), cap = sm_count() * 16;
  return g < cap ? (g > 0 ? g : 1) : cap;
}

}  // namespace
}  // namespace comb

using namespace comb;

extern "C" int comb_affine_relu(const void* x, int dtype, int n_max, const int* n_dev, int C, const float* scale,
                                const float* shift, const void* residual, int relu, void* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(n_max >= 0 && C >= 1, "comb_affine_relu: bad shape");
  COMB_CHECK_ARG((scale == nullptr) == (shift == nullptr), "comb_affine_relu: scale and shift go together");
  if (n_max == 0) return COMB_OK;
  COMB_CHECK_ARG(x && out, "comb_affine_relu: null pointer");
  int grid = ew_grid((long long)n_max * C);
  if (dtype == COMB_DT_F32)
    affine_relu_kernel<float><<<grid, 256, 0, stream>>>((const float*)x, n_max, n_dev, C, scale, shift,
                                                        (const float*)residual, relu, (float*)out);
  else if (dtype == COMB_DT_BF16)
    affine_relu_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, n_max, n_dev, C, scale, shift,
                                                                (const __nv_bfloat16*)residual, relu,
                                                                (__nv_bfloat16*)out);
  else
    COMB_CHECK_ARG(false, "comb_affine_relu: unknown dtype %d", dtype);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

extern "C" int comb_cast_pad(const float* x, int n_max, const int* n_dev, int C, void* out_bf16, int ld,
                             void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(n_max >= 0 && C >= 1 && ld >= C, "comb_cast_pad: bad shape");
  if (n_max == 0) return COMB_OK;
  COMB_CHECK_ARG(x && out_bf16, "comb_cast_pad: null pointer");
  cast_pad_kernel<<<ew_grid((long long)n_max * ld), 256, 0, stream>>>(x, n_max, n_dev, C, (__nv_bfloat16*)out_bf16, ld);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

extern "C" int comb_permute_rows(const void* in, const int* row_map, int n_max, const int* n_dev, int row_bytes,
                                 int scatter, void* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(n_max >= 0 && row_bytes >= 4 && row_bytes % 4 == 0, "comb_permute_rows: bad shape");
  if (n_max == 0) return COMB_OK;
  COMB_CHECK_ARG(in && row_map && out, "comb_permute_rows: null pointer");
  const int rw = row_bytes / 4;
  permute_rows_kernel<<<ew_grid((long long)n_max * rw), 256, 0, stream>>>((const uint32_t*)in, row_map, n_max, n_dev, rw,
                                                                         scatter, (uint32_t*)out);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

// Scatter form of dense(): `out` is already zero (the caller clears it early, off the critical path, e.g. with
// cudaMemsetAsync on a side stream) and only the active cells are written.  One thread per (row, 8-channel group):
// one 16-byte load of the row's channels, 8 stores into 8 channel planes; lanes of a warp are CONSECUTIVE rows, and
// with rows in key order consecutive rows are x-neighbours, so a warp's stores into one plane land in one or two
// sectors.  Algorithmic bytes: n*C*(in + 4) + n*16 instead of the full B*C*D*H*W*4 rewrite.
template <typename T>
__global__ void __launch_bounds__(256) dense_scatter_kernel(const T* __restrict__ feats, const int4* __restrict__ coords,
                                                             int n_max, const int* __restrict__ n_dev, int batch, int C,
                                                             int D, int H, int W, float* __restrict__ out) {
  const int n = eff_n(n_max, n_dev);
  const int groups = (C + 7) >> 3;
  const long long total = (long long)groups * ((n + 31) & ~31);
  const long long DHW = (long long)D * H * W;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    // 32 consecutive rows x one channel group per warp
    const int row = (int)((e >> 5) / groups) * 32 + (int)(e & 31);
    const int g = (int)((e >> 5) % groups);
    if (row >= n) continue;
    const int4 c = __ldg(coords + row);
    if ((unsigned)c.x >= (unsigned)batch || (unsigned)c.y >= (unsigned)D || (unsigned)c.z >= (unsigned)H ||
        (unsigned)c.w >= (unsigned)W)
      continue;
    float v[8];
    const int c0 = g * 8;
    if (sizeof(T) == 2 && (C & 7) == 0) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(feats + (size_t)row * C + c0));
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h2[i]);
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = c0 + i < C ? (float)feats[(size_t)row * C + c0 + i] : 0.0f;
    }
    float* o = out + ((size_t)c.x * C + c0) * DHW + ((size_t)c.y * H + c.z) * W + c.w;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (c0 + i < C) o[(size_t)i * DHW] = v[i];
  }
}

// Backward of dense(): grad_feats[row, c] = grad_dense[b, c, z, y, x] (the adjoint of the scatter; cells without a
// row receive no gradient).  Same thread mapping as dense_scatter_kernel: lanes = consecutive rows (x-neighbours in
// key order), one 8-channel group per thread, 8 plane reads -> one 16/32-byte row store.
template <typename T>
__global__ void __launch_bounds__(256) dense_gather_kernel(const float* __restrict__ dense, const int4* __restrict__ coords,
                                                            int n_max, const int* __restrict__ n_dev, int batch, int C,
                                                            int D, int H, int W, T* __restrict__ out) {
  const int n = eff_n(n_max, n_dev);
  const int groups = (C + 7) >> 3;
  const long long total = (long long)groups * ((n + 31) & ~31);
  const long long DHW = (long long)D * H * W;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int row = (int)((e >> 5) / groups) * 32 + (int)(e & 31);
    const int g = (int)((e >> 5) % groups);
    if (row >= n) continue;
    const int4 c = __ldg(coords + row);
    const int c0 = g * 8;
    const bool ok = (unsigned)c.x < (unsigned)batch && (unsigned)c.y < (unsigned)D && (unsigned)c.z < (unsigned)H &&
                    (unsigned)c.w < (unsigned)W;
    const float* o = dense + ((size_t)(ok ? c.x : 0) * C + c0) * DHW + ((size_t)(ok ? c.y : 0) * H + (ok ? c.z : 0)) * W +
                     (ok ? c.w : 0);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (c0 + i < C) out[(size_t)row * C + c0 + i] = (T)(ok ? __ldg(o + (size_t)i * DHW) : 0.0f);
  }
}

extern "C" size_t comb_dense_workspace_bytes(int batch, int D, int H, int W) {
  if (batch < 1 || D < 1 || H < 1 || W < 1) return 0;
  return align_up((size_t)batch * D * H * W * 4, 256);
}

extern "C" int comb_dense(const void* feats, int dtype, const int* coords, int n_max, const int* n_dev, int batch,
                          int C, int D, int H, int W, float* out, void* workspace, size_t workspace_bytes,
                          void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(batch >= 1 && C >= 1 && D >= 1 && H >= 1 && W >= 1 && n_max >= 0, "comb_dense: bad shape");
  COMB_CHECK_ARG(out && workspace, "comb_dense: null pointer");
  COMB_CHECK_ARG(workspace_bytes >= comb_dense_workspace_bytes(batch, D, H, W), "comb_dense: workspace too small");
  COMB_CHECK_ARG((long long)D * H * W < (1ll << 31) / 1, "comb_dense: frame volume too large");
  const int DHW = D * H * W;
  int* cell_row = (int*)workspace;
  COMB_CUDA(cudaMemsetAsync(cell_row, 0xFF, (size_t)batch * DHW * 4, stream));
  if (n_max > 0) {
    COMB_CHECK_ARG(feats && coords, "comb_dense: null feats/coords");
    dense_index_kernel<<<cdiv(n_max, 256), 256, 0, stream>>>((const int4*)coords, n_max, n_dev, batch, D, H, W,
                                                             cell_row);
    COMB_LAUNCH_CHECK();
  }
  if (DHW % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (dtype == COMB_DT_F32 || dtype == COMB_DT_BF16)) {
    const int tpf = cdiv(DHW, kCellsV);
    const long long nb = (long long)tpf * batch;
    COMB_CHECK_ARG(nb < (1ll << 31), "comb_dense: too many tiles");
    if (dtype == COMB_DT_F32)
      dense_write_vec_kernel<float><<<(unsigned)nb, 256, 0, stream>>>((const float*)feats, cell_row, C, DHW, tpf, out);
    else
      dense_write_vec_kernel<__nv_bfloat16><<<(unsigned)nb, 256, 0, stream>>>((const __nv_bfloat16*)feats, cell_row, C,
                                                                              DHW, tpf, out);
    COMB_LAUNCH_CHECK();
    return COMB_OK;
  }
  const int tiles_per_frame = cdiv(DHW, kCells);
  const long long blocks = (long long)tiles_per_frame * batch;
  COMB_CHECK_ARG(blocks < (1ll << 31), "comb_dense: too many tiles");
  size_t smem = (size_t)kCells * (C + 1) * sizeof(float);
  COMB_CHECK_ARG(smem <= 48 * 1024, "comb_dense: C %d too large", C);
  if (dtype == COMB_DT_F32)
    dense_write_kernel<float><<<(unsigned)blocks, 256, smem, stream>>>((const float*)feats, cell_row, C, DHW,
                                                                      tiles_per_frame, out);
  else if (dtype == COMB_DT_BF16)
    dense_write_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, smem, stream>>>((const __nv_bfloat16*)feats, cell_row,
                                                                              C, DHW, tiles_per_frame, out);
  else
    COMB_CHECK_ARG(false, "comb_dense: unknown dtype %d", dtype);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

extern "C" int comb_dense_scatter(const void* feats, int dtype, const int* coords, int n_max, const int* n_dev,
                                  int batch, int C, int D, int H, int W, float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(batch >= 1 && C >= 1 && D >= 1 && H >= 1 && W >= 1 && n_max >= 0, "comb_dense_scatter: bad shape");
  COMB_CHECK_ARG(out, "comb_dense_scatter: null pointer");
  if (n_max == 0) return COMB_OK;
  COMB_CHECK_ARG(feats && coords, "comb_dense_scatter: null feats/coords");
  const long long total = (long long)((C + 7) / 8) * ((n_max + 31) / 32 * 32);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (dtype == COMB_DT_F32)
    dense_scatter_kernel<float><<<(unsigned)blocks, 256, 0, stream>>>((const float*)feats, (const int4*)coords, n_max, n_dev,
                                                                       batch, C, D, H, W, out);
  else if (dtype == COMB_DT_BF16)
    dense_scatter_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, stream>>>((const __nv_bfloat16*)feats, (const int4*)coords,
                                                                                n_max, n_dev, batch, C, D, H, W, out);
  else
    COMB_CHECK_ARG(false, "comb_dense_scatter: unknown dtype %d", dtype);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

extern "C" int comb_dense_gather(const float* dense, const int* coords, int n_max, const int* n_dev, int batch, int C,
                                 int D, int H, int W, void* out, int dtype, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(batch >= 1 && C >= 1 && D >= 1 && H >= 1 && W >= 1 && n_max >= 0, "comb_dense_gather: bad shape");
  if (n_max == 0) return COMB_OK;
  COMB_CHECK_ARG(dense && coords && out, "comb_dense_gather: null pointer");
  const long long total = (long long)((C + 7) / 8) * ((n_max + 31) / 32 * 32);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (dtype == COMB_DT_F32)
    dense_gather_kernel<float><<<(unsigned)blocks, 256, 0, stream>>>(dense, (const int4*)coords, n_max, n_dev, batch, C, D,
                                                                      H, W, (float*)out);
  else if (dtype == COMB_DT_BF16)
    dense_gather_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, stream>>>(dense, (const int4*)coords, n_max, n_dev,
                                                                               batch, C, D, H, W, (__nv_bfloat16*)out);
  else
    COMB_CHECK_ARG(false, "comb_dense_gather: unknown dtype %d", dtype);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

// ---- f4: the BEV tensor in channels-last bf16 -------------------------------------------------------------------------
// HeightCompression (pcdet/models/backbones_2d/map_to_bev/height_compression.py:21-24) turns the encoded sparse tensor
// into (N, C*D, H, W) for the 2D backbone (pcdet/models/backbones_2d/base_bev_backbone.py:81-112), channel index
// c*D + z.  Here the rows are scattered straight into the NHWC image out[b][y][x][c*D + z] in bf16 (pre-zeroed by the
// caller): the 2D backbone then runs its convolutions on bf16 tensor cores without a layout or dtype conversion pass,
// and the image is half the bytes of the fp32 NCDHW one.  One warp per row, lanes walk the channels.
namespace comb {
namespace {
template <typename T>
__global__ void __launch_bounds__(256) dense_scatter_nhwc_kernel(const T* __restrict__ feats, const int4* __restrict__ coords,
                                                                  int n_max, const int* __restrict__ n_dev, int batch, int C,
                                                                  int D, int H, int W, __nv_bfloat16* __restrict__ out) {
  const int n = eff_n(n_max, n_dev);
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < n; row += gridDim.x * wpb) {
    const int4 c = __ldg(coords + row);
    if ((unsigned)c.x >= (unsigned)batch || (unsigned)c.y >= (unsigned)D || (unsigned)c.z >= (unsigned)H ||
        (unsigned)c.w >= (unsigned)W)
      continue;
    __nv_bfloat16* o = out + (((size_t)c.x * H + c.z) * W + c.w) * ((size_t)C * D) + c.y;
    for (int ch = lane; ch < C; ch += 32) o[(size_t)ch * D] = (__nv_bfloat16)feats[(size_t)row * C + ch];
  }
}
// adjoint: rows[row][ch] = grad[b][y][x][ch*D + z]
template <typename G, typename T>
__global__ void __launch_bounds__(256) dense_gather_nhwc_kernel(const G* __restrict__ grad, const int4* __restrict__ coords,
                                                                 int n_max, const int* __restrict__ n_dev, int batch, int C,
                                                                 int D, int H, int W, T* __restrict__ out) {
  const int n = eff_n(n_max, n_dev);
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < n; row += gridDim.x * wpb) {
    const int4 c = __ldg(coords + row);
    const bool ok = (unsigned)c.x < (unsigned)batch && (unsigned)c.y < (unsigned)D && (unsigned)c.z < (unsigned)H &&
                    (unsigned)c.w < (unsigned)W;
    const G* g = grad + (((size_t)(ok ? c.x : 0) * H + (ok ? c.z : 0)) * W + (ok ? c.w : 0)) * ((size_t)C * D) + (ok ? c.y : 0);
    for (int ch = lane; ch < C; ch += 32) out[(size_t)row * C + ch] = (T)(ok ? (float)g[(size_t)ch * D] : 0.0f);
  }
}
}  // namespace
}  // namespace comb

extern "C" int comb_dense_scatter_nhwc_bf16(const void* feats, int dtype, const int* coords, int n_max, const int* n_dev,
                                            int batch, int C, int D, int H, int W, void* out, void* stream_) {
  using namespace comb;
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(batch >= 1 && C >= 1 && D >= 1 && H >= 1 && W >= 1 && n_max >= 0, "comb_dense_scatter_nhwc_bf16: bad shape");
  if (n_max == 0) return COMB_OK;
  COMB_CHECK_ARG(feats && coords && out, "comb_dense_scatter_nhwc_bf16: null pointer");
  int blocks = cdiv(n_max, 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (dtype == COMB_DT_F32)
    dense_scatter_nhwc_kernel<float><<<blocks, 256, 0, stream>>>((const float*)feats, (const int4*)coords, n_max, n_dev, batch,
                                                                  C, D, H, W, (__nv_bfloat16*)out);
  else if (dtype == COMB_DT_BF16)
    dense_scatter_nhwc_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)feats, (const int4*)coords, n_max,
                                                                          n_dev, batch, C, D, H, W, (__nv_bfloat16*)out);
  else
    COMB_CHECK_ARG(false, "comb_dense_scatter_nhwc_bf16: unknown dtype %d", dtype);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

extern "C" int comb_dense_gather_nhwc(const void* grad, int grad_dtype, const int* coords, int n_max, const int* n_dev,
                                      int batch, int C, int D, int H, int W, void* out, int dtype, void* stream_) {
  using namespace comb;
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(batch >= 1 && C >= 1 && D >= 1 && H >= 1 && W >= 1 && n_max >= 0, "comb_dense_gather_nhwc: bad shape");
  if (n_max == 0) return COMB_OK;
  COMB_CHECK_ARG(grad && coords && out, "comb_dense_gather_nhwc: null pointer");
  COMB_CHECK_ARG((grad_dtype == COMB_DT_F32 || grad_dtype == COMB_DT_BF16) && (dtype == COMB_DT_F32 || dtype == COMB_DT_BF16),
                 "comb_dense_gather_nhwc: unknown dtype");
  int blocks = cdiv(n_max, 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  const int4* cd = (const int4*)coords;
  if (grad_dtype == COMB_DT_BF16 && dtype == COMB_DT_BF16)
    dense_gather_nhwc_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)grad, cd, n_max, n_dev, batch, C, D, H, W, (__nv_bfloat16*)out);
  else if (grad_dtype == COMB_DT_BF16)
    dense_gather_nhwc_kernel<__nv_bfloat16, float><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)grad, cd, n_max, n_dev, batch, C, D, H, W, (float*)out);
  else if (dtype == COMB_DT_BF16)
    dense_gather_nhwc_kernel<float, __nv_bfloat16><<<blocks, 256, 0, stream>>>((const float*)grad, cd, n_max, n_dev, batch, C, D, H, W, (__nv_bfloat16*)out);
  else
    dense_gather_nhwc_kernel<float, float><<<blocks, 256, 0, stream>>>((const float*)grad, cd, n_max, n_dev, batch, C, D, H, W, (float*)out);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}
