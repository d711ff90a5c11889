// a1/a2/a4 — hard voxelization with sequential ("first come") semantics on a parallel machine,
// plus fused MeanVFE.
//
// Replaces: spconv.utils.Point2VoxelCPU3d.point_to_voxel / VoxelGeneratorV2.generate as called from
//   VoxelGeneratorWrapper.generate        pcdet/datasets/processor/data_processor.py:44-60
//   DataProcessor.transform_points_to_voxels  data_processor.py:125-153
//   MeanVFE.forward                        pcdet/models/backbones_3d/vfe/mean_vfe.py:14-31
//
// The CPU generator walks the points once and is order dependent (voxel id = order of first
// appearance, first `max_voxels` voxels win, first `max_points` points of a voxel win).  The same
// result is obtained without any sequential pass:
//   K1 insert : every in-range point claims the hash slot of its voxel key and pushes its own index
//               through an atomicMin chain, so slot.cand[0..T) ends up holding the T smallest point
//               indices of the voxel in ascending order, whatever the interleaving.
//   K2 count  : a point is "first" iff cand[0] == its index; per 1024-point chunk count the firsts.
//   K3 scan   : exclusive scan of the chunk counts per frame -> voxel id of every first point is its
//               rank among the firsts = order of first appearance; ids >= max_voxels are dropped.
//   K4 assign : first points write coords / slot of their voxel row.
//   K5 fill   : rows gather their <=T points (zero padded), the count and the channel mean.
// All tables (keys 4 B/slot, cand 4*T B/slot, 2 slots per point) stay resident in the 126 MB L2.
#include "common.cuh"

namespace comb {
namespace {

constexpr int kMaxBatch = 32;
constexpr int kChunk = 1024;          // points per block in K2/K4
constexpr int kCandInf = 0x7F7F7F7F;  // memset(0x7F) pattern, larger than any point index

// Frame offsets: either captured by value from host integers, or read from device memory (`dev`), which keeps
// the launch sequence independent of the per-frame point counts (CUDA-graph replay with new inputs).
struct Frames {
  int off[kMaxBatch + 1];
  const int* dev;
  __device__ __forceinline__ int at(int b) const { return dev ? __ldg(dev + b) : off[b]; }
};

struct VoxGeom {
  float rmin[3];
  float vs[3];
  int grid[3];  // x, y, z
};

__device__ __forceinline__ bool point_key(const float* __restrict__ p, const VoxGeom& g, int frame, uint32_t& key) {
  int c[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    // fp32 IEEE subtract and divide, no reciprocal, no contraction (matches the CPU generator)
    float q = floorf(__fdiv_rn(__fsub_rn(p[j], g.rmin[j]), g.vs[j]));
    if (!(q >= 0.0f && q < (float)g.grid[j])) return false;  // also rejects NaN
    c[j] = (int)q;
  }
  key = (uint32_t)(((frame * g.grid[2] + c[2]) * g.grid[1] + c[1]) * g.grid[0] + c[0]);
  return true;
}

__global__ void __launch_bounds__(256) vox_insert_kernel(const float* __restrict__ points, Frames fr, int C, VoxGeom g,
                                                          int T, uint32_t* __restrict__ keys, int* __restrict__ cand,
                                                          uint32_t mask, int* __restrict__ slot_of_point, int batch) {
  // one thread per point of the concatenated cloud; its frame = the last offset <= i (batch <= 32: linear search).
  // (A (points-per-frame, batch) grid sized for the worst case "one frame holds every point" launches batch x more
  // blocks than there are points when the offsets live on the device.)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= fr.at(batch)) return;
  int frame = 0;
  while (frame + 1 < batch && i >= fr.at(frame + 1)) ++frame;
  const float* p = points + (size_t)i * C;
  float xyz[3] = {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
  uint32_t key;
  if (!point_key(xyz, g, frame, key)) {
    slot_of_point[i] = -1;
    return;
  }
  uint32_t s = hash_u32(key) & mask;
  while (true) {
    uint32_t cur = *((volatile uint32_t*)(keys + s));
    if (cur == key) break;
    if (cur == kEmptyKey) {
      uint32_t prev = atomicCAS(keys + s, kEmptyKey, key);
      if (prev == kEmptyKey || prev == key) break;
    }
    s = (s + 1) & mask;
  }
  slot_of_point[i] = (int)s;
  int* c = cand + (size_t)s * T;
  // cand entries only ever decrease: a stale read that is already < i proves i is not among the T smallest
  if (*((volatile int*)(c + T - 1)) < i) return;
  int x = i;
  for (int t = 0; t < T; ++t) {
    int old = atomicMin(c + t, x);
    if (old == kCandInf) break;
    x = old > x ? old : x;
  }
}

__device__ __forceinline__ int block_exclusive_scan_1024(int v, int* total, int* smem /*33 ints*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += n;
  }
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = (lane < (blockDim.x >> 5)) ? smem[lane] : 0;
    int wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += n;
    }
    smem[lane] = wi - w;  // exclusive warp offsets
    if (lane == 31) smem[32] = wi;
  }
  __syncthreads();
  int res = smem[warp] + incl - v;
  *total = smem[32];
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(kChunk) vox_count_kernel(Frames fr, int T, const int* __restrict__ cand,
                                                            const int* __restrict__ slot_of_point,
                                                            int* __restrict__ chunk_counts, int chunks_per_frame) {
  const int frame = blockIdx.y;
  const int i = fr.at(frame) + blockIdx.x * kChunk + threadIdx.x;
  bool first = false;
  if (i < fr.at(frame + 1)) {
    int s = slot_of_point[i];
    first = (s >= 0) && (cand[(size_t)s * T] == i);
  }
  int cnt = __syncthreads_count(first);
  if (threadIdx.x == 0) chunk_counts[frame * chunks_per_frame + blockIdx.x] = cnt;
}

// One block: per frame exclusive scan of chunk counts; frame totals clamped to max_voxels; frame bases.
__global__ void __launch_bounds__(1024) vox_scan_kernel(Frames fr, int batch, int chunks_per_frame, int max_voxels,
                                                         const int* __restrict__ chunk_counts,
                                                         int* __restrict__ chunk_offsets, int* __restrict__ frame_base,
                                                         int* __restrict__ counts) {
  __shared__ int smem[33];
  __shared__ int s_base;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int b = 0; b < batch; ++b) {
    const int nchunks = (fr.at(b + 1) - fr.at(b) + kChunk - 1) / kChunk;
    int running = 0;
    for (int c0 = 0; c0 < nchunks; c0 += 1024) {
      int c = c0 + threadIdx.x;
      int v = (c < nchunks) ? chunk_counts[b * chunks_per_frame + c] : 0;
      int tot;
      int ex = block_exclusive_scan_1024(v, &tot, smem);
      if (c < nchunks) chunk_offsets[b * chunks_per_frame + c] = running + ex;
      running += tot;
    }
    if (threadIdx.x == 0) {
      int m = running < max_voxels ? running : max_voxels;
      counts[b] = m;
      frame_base[b] = s_base;
      s_base += m;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    counts[batch] = s_base;
    frame_base[batch] = s_base;
  }
}

__global__ void __launch_bounds__(kChunk) vox_assign_kernel(Frames fr, VoxGeom g, int T, int max_voxels,
                                                             const uint32_t* __restrict__ keys,
                                                             const int* __restrict__ cand,
                                                             const int* __restrict__ slot_of_point,
                                                             const int* __restrict__ chunk_offsets,
                                                             const int* __restrict__ frame_base, int chunks_per_frame,
                                                             int* __restrict__ voxel_slot, int* __restrict__ coords) {
  __shared__ int smem[33];
  const int frame = blockIdx.y;
  const int i = fr.at(frame) + blockIdx.x * kChunk + threadIdx.x;
  bool first = false;
  int s = -1;
  if (i < fr.at(frame + 1)) {
    s = slot_of_point[i];
    first = (s >= 0) && (cand[(size_t)s * T] == i);
  }
  int tot;
  int rank = block_exclusive_scan_1024(first ? 1 : 0, &tot, smem);
  if (!first) return;
  int vid = chunk_offsets[frame * chunks_per_frame + blockIdx.x] + rank;
  if (vid >= max_voxels) return;
  int row = frame_base[frame] + vid;
  voxel_slot[row] = s;
  uint32_t key = keys[s];
  int x = key % g.grid[0];
  uint32_t r = key / g.grid[0];
  int y = r % g.grid[1];
  r /= g.grid[1];
  int z = r % g.grid[2];
  reinterpret_cast<int4*>(coords)[row] = make_int4(frame, z, y, x);
}

// voxels[row][t][c] (zero padded) — one thread per output element, coalesced stores.
__global__ void __launch_bounds__(256) vox_fill_kernel(const float* __restrict__ points, int C, int T,
                                                        const int* __restrict__ cand,
                                                        const int* __restrict__ voxel_slot,
                                                        const int* __restrict__ total_ptr, float* __restrict__ voxels) {
  const long long n = (long long)(*total_ptr) * T * C;
  const int TC = T * C;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int row = (int)(e / TC);
    int r = (int)(e - (long long)row * TC);
    int t = r / C, c = r - t * C;
    int idx = cand[(size_t)voxel_slot[row] * T + t];
    voxels[e] = (idx != kCandInf) ? __ldg(points + (size_t)idx * C + c) : 0.0f;
  }
}

// num_points[row] and the MeanVFE feature row.  One thread per (row, out column).
template <typename OutT>
__global__ void __launch_bounds__(256) vox_mean_kernel(const float* __restrict__ points, int C, int T,
                                                        const int* __restrict__ cand,
                                                        const int* __restrict__ voxel_slot,
                                                        const int* __restrict__ total_ptr, int* __restrict__ num_points,
                                                        OutT* __restrict__ mean_out, int c0, int ld) {
  const int cols = mean_out ? ld : 1;
  const long long n = (long long)(*total_ptr) * cols;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int row = (int)(e / cols);
    int c = (int)(e - (long long)row * cols);
    const int* cd = cand + (size_t)voxel_slot[row] * T;
    int num = 0;
    float s = 0.0f;
    const bool live = mean_out && (c0 + c < C);
    for (int t = 0; t < T; ++t) {
      int idx = cd[t];
      if (idx == kCandInf) break;
      ++num;
      if (live) s = __fadd_rn(s, __ldg(points + (size_t)idx * C + c0 + c));
    }
    if (c == 0) num_points[row] = num;
    if (mean_out) {
      float m = live ? __fdiv_rn(s, (float)(num > 1 ? num : 1)) : 0.0f;
      if constexpr (sizeof(OutT) == 2)
        mean_out[e] = __float2bfloat16(m);
      else
        mean_out[e] = m;
    }
  }
}

// Same result for the pipeline's layout (bf16 rows padded to 16 columns): ONE thread per voxel row walks the
// candidate list once and keeps the <= 16 channel sums in registers (the per-(row, column) form above walks the list
// 16 times per row, 11 of them for padding columns: r1 ncu 23 M warp instructions, 39 us, issue-bound).
// Summation order per channel is unchanged (t = 0..T-1, fp32 round-to-nearest adds), so the bits are the same.
__global__ void __launch_bounds__(256) vox_mean_row16_kernel(const float* __restrict__ points, int C, int T,
                                                              const int* __restrict__ cand,
                                                              const int* __restrict__ voxel_slot,
                                                              const int* __restrict__ total_ptr,
                                                              int* __restrict__ num_points,
                                                              __nv_bfloat16* __restrict__ mean_out, int c0) {
  const int n = *total_ptr;
  const int live = C - c0 < 16 ? C - c0 : 16;      // real channels among the 16 output columns
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n; row += gridDim.x * blockDim.x) {
    const int* cd = cand + (size_t)voxel_slot[row] * T;
    float s[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) s[c] = 0.0f;
    int num = 0;
    for (int t = 0; t < T; ++t) {
      const int idx = cd[t];
      if (idx == kCandInf) break;
      ++num;
      const float* pt = points + (size_t)idx * C + c0;
#pragma unroll
      for (int c = 0; c < 16; ++c)
        if (c < live) s[c] = __fadd_rn(s[c], __ldg(pt + c));
    }
    num_points[row] = num;
    const float den = (float)(num > 1 ? num : 1);
    uint4 o[2];
    __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(o);
#pragma unroll
    for (int c = 0; c < 16; ++c) ob[c] = __float2bfloat16(c < live ? __fdiv_rn(s[c], den) : 0.0f);
    uint4* dst = reinterpret_cast<uint4*>(mean_out + (size_t)row * 16);
    dst[0] = o[0];
    dst[1] = o[1];
  }
}

template <typename NumT>
__global__ void __launch_bounds__(256) mean_vfe_kernel(const float* __restrict__ voxels, const NumT* __restrict__ num,
                                                        int M, int T, int C, float* __restrict__ out) {
  const long long n = (long long)M * C;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int row = (int)(e / C), c = (int)(e - (long long)row * C);
    const float* v = voxels + (size_t)row * T * C + c;
    float s = 0.0f;
    for (int t = 0; t < T; ++t) s = __fadd_rn(s, __ldg(v + (size_t)t * C));
    float d = (float)num[row];
    d = d < 1.0f ? 1.0f : d;  // torch.clamp_min(num, 1.0)
    out[e] = __fdiv_rn(s, d);
  }
}

struct VoxWs {
  uint32_t* keys;
  int* cand;
  int* slot_of_point;
  int* chunk_counts;
  int* chunk_offsets;
  int* frame_base;
  int* voxel_slot;
  uint32_t slots;
  int chunks_per_frame;
  size_t bytes;
};

static uint32_t vox_slots(int n_total) {
  uint32_t s = 1024;
  while (s < 2ull * (uint32_t)n_total) s <<= 1;
  return s;
}

static VoxWs carve(void* base, int n_total, int batch, int max_voxels, int T) {
  VoxWs w;
  w.slots = vox_slots(n_total);
  // worst case one frame holds every point
  w.chunks_per_frame = cdiv(n_total > 0 ? n_total : 1, kChunk);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return (char*)base + o;
  };
  w.keys = (uint32_t*)take((size_t)w.slots * 4);
  w.cand = (int*)take((size_t)w.slots * T * 4);
  w.slot_of_point = (int*)take((size_t)(n_total > 0 ? n_total : 1) * 4);
  w.chunk_counts = (int*)take((size_t)batch * w.chunks_per_frame * 4);
  w.chunk_offsets = (int*)take((size_t)batch * w.chunks_per_frame * 4);
  w.frame_base = (int*)take((size_t)(batch + 1) * 4);
  w.voxel_slot = (int*)take((size_t)batch * max_voxels * 4);
  w.bytes = off;
  return w;
}

}  // namespace
}  // namespace comb

using namespace comb;

extern "C" size_t comb_voxelize_workspace_bytes(int n_total, int batch, int max_voxels, int max_points) {
  if (n_total < 0 || batch <= 0 || max_voxels <= 0 || max_points <= 0) return 0;
  return carve(nullptr, n_total, batch, max_voxels, max_points).bytes;
}

extern "C" int comb_voxelize(const float* points, const int* frame_offsets_host, const int* frame_offsets_dev,
                             int n_cap, int batch, int C,
                             const float* vsize, const float* range, int max_points, int max_voxels, float* voxels,
                             int* coords, int* num_points, void* mean_out, int mean_dtype, int mean_c0, int mean_ld,
                             int* counts, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(batch >= 1 && batch <= kMaxBatch, "comb_voxelize: batch %d outside [1,%d]", batch, kMaxBatch);
  COMB_CHECK_ARG(C >= 3, "comb_voxelize: points need >= 3 channels, got %d", C);
  COMB_CHECK_ARG(max_points >= 1 && max_voxels >= 1, "comb_voxelize: max_points/max_voxels must be >= 1");
  COMB_CHECK_ARG(coords && num_points && counts && workspace, "comb_voxelize: null output/workspace pointer");
  COMB_CHECK_ARG((frame_offsets_host || frame_offsets_dev) && vsize && range,
                 "comb_voxelize: null host parameter pointer");
  Frames fr;
  int max_frame = 0, n_total = 0;
  fr.dev = frame_offsets_dev;
  if (frame_offsets_dev) {
    // device-side offsets: every launch is sized for n_cap points per frame, surplus threads exit
    COMB_CHECK_ARG(n_cap >= 0, "comb_voxelize: n_cap must be given with device-side frame offsets");
    for (int b = 0; b <= batch; ++b) fr.off[b] = 0;
    max_frame = n_total = n_cap;
  } else {
    for (int b = 0; b <= batch; ++b) fr.off[b] = frame_offsets_host[b];
    COMB_CHECK_ARG(fr.off[0] == 0, "comb_voxelize: frame_offsets[0] must be 0");
    for (int b = 0; b < batch; ++b) {
      COMB_CHECK_ARG(fr.off[b + 1] >= fr.off[b], "comb_voxelize: frame_offsets must be non-decreasing");
      max_frame = fr.off[b + 1] - fr.off[b] > max_frame ? fr.off[b + 1] - fr.off[b] : max_frame;
    }
    n_total = fr.off[batch];
  }
  COMB_CHECK_ARG(n_total == 0 || points, "comb_voxelize: null points");
  VoxGeom g;
  for (int j = 0; j < 3; ++j) {
    g.rmin[j] = range[j];
    g.vs[j] = vsize[j];
    COMB_CHECK_ARG(vsize[j] > 0.f, "comb_voxelize: voxel size must be positive");
    // grid = round((max - min) / vsize), data_processor.py:127-128
    g.grid[j] = (int)lrintf((range[3 + j] - range[j]) / vsize[j]);
    COMB_CHECK_ARG(g.grid[j] >= 1, "comb_voxelize: empty grid on axis %d", j);
  }
  const unsigned long long vol = (unsigned long long)g.grid[0] * g.grid[1] * g.grid[2] * batch;
  if (vol >= 0xFFFFFFFFull) {
    set_error("comb_voxelize: batch*grid volume %llu exceeds the 32-bit key space", vol);
    return COMB_ERANGE;
  }
  VoxWs w = carve(workspace, n_total, batch, max_voxels, max_points);
  COMB_CHECK_ARG(workspace_bytes >= w.bytes, "comb_voxelize: workspace %zu < required %zu", workspace_bytes, w.bytes);
  COMB_CHECK_ARG(mean_out == nullptr || (mean_ld >= 1 && mean_c0 >= 0 && mean_c0 < C),
                 "comb_voxelize: bad mean_c0/mean_ld");

  COMB_CUDA(cudaMemsetAsync(w.keys, 0xFF, (size_t)w.slots * 4, stream));
  COMB_CUDA(cudaMemsetAsync(w.cand, 0x7F, (size_t)w.slots * max_points * 4, stream));
  const int T = max_points;
  if (max_frame > 0) {
    vox_insert_kernel<<<cdiv(n_total, 256), 256, 0, stream>>>(points, fr, C, g, T, w.keys, w.cand, w.slots - 1,
                                                              w.slot_of_point, batch);
    COMB_LAUNCH_CHECK();
    dim3 gc(cdiv(max_frame, kChunk), batch);
    vox_count_kernel<<<gc, kChunk, 0, stream>>>(fr, T, w.cand, w.slot_of_point, w.chunk_counts, w.chunks_per_frame);
    COMB_LAUNCH_CHECK();
  }
  vox_scan_kernel<<<1, 1024, 0, stream>>>(fr, batch, w.chunks_per_frame, max_voxels, w.chunk_counts, w.chunk_offsets,
                                          w.frame_base, counts);
  COMB_LAUNCH_CHECK();
  if (max_frame > 0) {
    dim3 gc(cdiv(max_frame, kChunk), batch);
    vox_assign_kernel<<<gc, kChunk, 0, stream>>>(fr, g, T, max_voxels, w.keys, w.cand, w.slot_of_point,
                                                  w.chunk_offsets, w.frame_base, w.chunks_per_frame, w.voxel_slot,
                                                  coords);
    COMB_LAUNCH_CHECK();
    const int grid = sm_count() * 8;
    if (voxels) {
      vox_fill_kernel<<<grid, 256, 0, stream>>>(points, C, T, w.cand, w.voxel_slot, counts + batch, voxels);
      COMB_LAUNCH_CHECK();
    }
    if (mean_out && mean_dtype == COMB_DT_BF16 && mean_ld == 16 && (reinterpret_cast<uintptr_t>(mean_out) & 15) == 0)
      vox_mean_row16_kernel<<<grid, 256, 0, stream>>>(points, C, T, w.cand, w.voxel_slot, counts + batch, num_points,
                                                      (__nv_bfloat16*)mean_out, mean_c0);
    else if (mean_out && mean_dtype == COMB_DT_BF16)
      vox_mean_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(points, C, T, w.cand, w.voxel_slot, counts + batch,
                                                               num_points, (__nv_bfloat16*)mean_out, mean_c0, mean_ld);
    else
      vox_mean_kernel<float><<<grid, 256, 0, stream>>>(points, C, T, w.cand, w.voxel_slot, counts + batch, num_points,
                                                       (float*)mean_out, mean_c0, mean_ld);
    COMB_LAUNCH_CHECK();
  }
  return COMB_OK;
}

extern "C" int comb_mean_vfe(const float* voxels, const void* num_points, int num_is_float, int M, int T, int C,
                             float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(M >= 0 && T >= 1 && C >= 1, "comb_mean_vfe: bad shape");
  if (M == 0) return COMB_OK;
  COMB_CHECK_ARG(voxels && num_points && out, "comb_mean_vfe: null pointer");
  int grid = cdiv((long long)M * C, 256);
  int cap = sm_count() * 16;
  grid = grid < cap ? grid : cap;
  if (num_is_float)
    mean_vfe_kernel<float><<<grid, 256, 0, stream>>>(voxels, (const float*)num_points, M, T, C, out);
  else
    mean_vfe_kernel<int><<<grid, 256, 0, stream>>>(voxels, (const int*)num_points, M, T, C, out);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}
