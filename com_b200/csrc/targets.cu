// f1 — CenterHead target assignment and the COM (curriculum) loss re-weighting on the device.
//
// Replaces the per-object Python loops of the reference (every object costs host arithmetic, `.item()` reads and a
// handful of tiny tensor kernels; the targets are built on the CPU and copied back):
//   CurriculumCenterHead.cluster                      pcdet/models/dense_heads/curriculum_center_head.py:431-473
//   CurriculumCenterHead.assign_targets               curriculum_center_head.py:203-296
//     .assign_target_of_single_head                   curriculum_center_head.py:120-201
//     centernet_utils.gaussian_radius / draw_gaussian_to_heatmap   pcdet/models/model_utils/centernet_utils.py:48-108
//   FocalLossCenterCurriculum.confidence_of_all_groups / group_confifence   pcdet/utils/loss_utils.py:1131-1178
//   FocalLossCenterCurriculum.neg_loss, the object loop             loss_utils.py:1222-1287
//     centernet_utils.draw_mask_to_heatmap                          centernet_utils.py:110-131
// (CenterHead.assign_targets, pcdet/models/dense_heads/center_head.py:119-236, is the same assignment without the
// point-count filter and the group column.)
//
// Parity.  The integer outputs (inds, radius_map, masks) and the heat maps are bit-exact: the fp32 arithmetic follows
// the reference operation by operation with separately rounded multiplies / divides / square roots (torch runs them as
// separate fp32 kernels on the CPU), and the Gaussian values come from a table the HOST builds with the reference's own
// numpy formula (float64 exp, eps cut, cast to fp32) for every radius up to the table size; beyond it the kernel
// evaluates the same formula in fp64 on the device.  log / cos / sin of the regression targets use the device fp32
// functions (1e-6).  The re-weighting follows the reference's float64 host arithmetic in fp64.
#include "common.cuh"

namespace comb {
namespace {

constexpr int kTgtThreads = 256;
constexpr int kMaxObjsCap = 1024;

__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }

// centernet_utils.gaussian_radius (centernet_utils.py:48-75), height = dx, width = dy as the caller passes them
__device__ float gaussian_radius_f32(float height, float width, float ov_minus /*1-ov*/, float ov_plus /*1+ov*/,
                                     float a3 /*4*ov*/, float b3c /*-2*ov*/, float c3c /*ov-1*/, float a3x4 /*4*a3*/) {
  const float b1 = fadd(height, width);
  const float c1 = fdiv(fmul(fmul(width, height), ov_minus), ov_plus);
  const float sq1 = __fsqrt_rn(fsub(fmul(b1, b1), fmul(4.0f, c1)));
  const float r1 = fdiv(fadd(b1, sq1), 2.0f);
  const float b2 = fmul(2.0f, fadd(height, width));
  const float c2 = fmul(fmul(ov_minus, width), height);
  const float sq2 = __fsqrt_rn(fsub(fmul(b2, b2), fmul(16.0f, c2)));
  const float r2 = fdiv(fadd(b2, sq2), 2.0f);
  const float b3 = fmul(b3c, fadd(height, width));
  const float c3 = fmul(fmul(c3c, width), height);
  const float sq3 = __fsqrt_rn(fsub(fmul(b3, b3), fmul(a3x4, c3)));
  const float r3 = fdiv(fadd(b3, sq3), 2.0f);
  (void)a3;
  return fminf(fminf(r1, r2), r3);      // torch.min propagates NaN; NaN radii are skipped by the caller (dx, dy <= 0)
}

struct AssignArgs {
  const float* gt;          // [B][M][C]; class id (1-based, 0 = padding row) in column C-1
  float* gt_rw;             // the same tensor, written when relabel is set
  int relabel;
  const float* npgt;        // [B][M] points inside each box
  const long long* group;   // [B][M] curriculum group, or NULL (radius_map then has 4 columns)
  const int* cls_map;       // [n_cls + 1]: global class id -> index inside this head, -1 = another head
  int n_cls;
  int B, M, C;
  float x0, y0, vx, vy, stride;
  int W, H;                 // feature map size
  int max_objs;
  float ov_minus, ov_plus, a3, b3c, c3c, a3x4;
  int min_radius;
  int filter_points;        // epoch <= EPOCH_THRED
  float min_points;
  const float* gtab;        // Gaussian tables, radius 0..rmax back to back
  const int* gtab_off;      // [rmax + 2]
  int rmax;
  float* heatmap;           // [B][Ch][H][W]   zero-filled by the caller
  int Ch;
  float* ret_boxes;         // [B][max_objs][C] zero-filled
  long long* inds;          // [B][max_objs]    zero-filled
  float* mask;              // [B][max_objs]    zero-filled
  long long* radius_map;    // [B][max_objs][R] zero-filled
  int R;
};

// one block per frame
__global__ void __launch_bounds__(kTgtThreads) assign_targets_kernel(AssignArgs a) {
  __shared__ int sel[kMaxObjsCap];        // object index (row of gt) of the k-th object of this head
  __shared__ int sel_cls[kMaxObjsCap];    // its class inside the head (read before the in-place relabelling)
  __shared__ int s_warp[kTgtThreads / 32];
  __shared__ int s_base;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* gt = a.gt + (size_t)b * a.M * a.C;
  if (tid == 0) s_base = 0;
  __syncthreads();
  // ---- objects of this head, in their original order (curriculum_center_head.py:249-262)
  for (int m0 = 0; m0 < a.M; m0 += kTgtThreads) {
    const int m = m0 + tid;
    bool in = false;
    if (m < a.M) {
      const long long cls = (long long)gt[(size_t)m * a.C + a.C - 1];     // .long(): truncation
      in = cls >= 0 && cls <= a.n_cls && a.cls_map[cls] >= 0;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, in);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    const int k = off + __popc(bal & ((1u << lane) - 1u));
    int local_cls = -1;
    if (in) local_cls = a.cls_map[(long long)gt[(size_t)m * a.C + a.C - 1]];
    if (in && k < a.max_objs) {
      sel[k] = m;
      sel_cls[k] = local_cls;
    }
    // The reference relabels the boxes of this head IN PLACE in the caller's gt_boxes while it collects them
    // (`temp_box[-1] = cur_class_names.index(name) + 1`, curriculum_center_head.py:252-254): later heads — and the
    // caller — see the head-local ids.  Reproduced when asked for (multi-head configurations).
    if (in && a.relabel) a.gt_rw[((size_t)b * a.M + m) * a.C + a.C - 1] = (float)(local_cls + 1);
    __syncthreads();
    if (tid == 0) {
      int t = s_base;
      for (int w = 0; w < kTgtThreads / 32; ++w) t += s_warp[w];
      s_base = t;
    }
    __syncthreads();
  }
  const int nobj = s_base < a.max_objs ? s_base : a.max_objs;
  // ---- one thread per object (curriculum_center_head.py:146-199)
  for (int k = tid; k < nobj; k += kTgtThreads) {
    const int m = sel[k];
    const float* g = gt + (size_t)m * a.C;
    const float x = g[0], y = g[1], z = g[2];
    float cx = fdiv(fdiv(fsub(x, a.x0), a.vx), a.stride);
    float cy = fdiv(fdiv(fsub(y, a.y0), a.vy), a.stride);
    // torch.clamp(min=0, max=size-0.5): NaN stays NaN (then fails the range test below)
    cx = cx < 0.0f ? 0.0f : (cx > (float)a.W - 0.5f ? (float)a.W - 0.5f : cx);
    cy = cy < 0.0f ? 0.0f : (cy > (float)a.H - 0.5f ? (float)a.H - 0.5f : cy);
    const int ix = (int)cx, iy = (int)cy;
    const float dx = fdiv(fdiv(g[3], a.vx), a.stride);
    const float dy = fdiv(fdiv(g[4], a.vy), a.stride);
    if (!(dx > 0.0f) || !(dy > 0.0f)) {
      if (dx <= 0.0f || dy <= 0.0f) continue;      // the reference's test; NaN sizes fall through as there
    }
    if (!(0 <= ix && ix <= a.W && 0 <= iy && iy <= a.H)) continue;
    if (a.filter_points && a.npgt[(size_t)b * a.M + m] < a.min_points) continue;
    const float rf = gaussian_radius_f32(dx, dy, a.ov_minus, a.ov_plus, a.a3, a.b3c, a.c3c, a.a3x4);
    int radius = (int)rf;                            // .int(): truncation
    if (radius < a.min_radius) radius = a.min_radius;
    const int cls = sel_cls[k];                            // (gt[-1] - 1) of the head-local 1-based id

    // draw_gaussian_to_heatmap (centernet_utils.py:86-108): element-wise max with the Gaussian window
    {
      float* hm = a.heatmap + ((size_t)b * a.Ch + cls) * a.H * a.W;
      const int left = ix < radius ? ix : radius, right = a.W - ix < radius + 1 ? a.W - ix : radius + 1;
      const int top = iy < radius ? iy : radius, bottom = a.H - iy < radius + 1 ? a.H - iy : radius + 1;
      if (right + left > 0 && bottom + top > 0 && right > -left && bottom > -top) {
        const int diam = 2 * radius + 1;
        const float* tab = radius <= a.rmax ? a.gtab + a.gtab_off[radius] : nullptr;
        const double sigma = (double)diam / 6.0;
        for (int yy = -top; yy < bottom; ++yy)
          for (int xx = -left; xx < right; ++xx) {
            float v;
            if (tab != nullptr) {
              v = tab[(radius + yy) * diam + (radius + xx)];
            } else {
              double h = exp(-((double)(xx * xx) + (double)(yy * yy)) / (2.0 * sigma * sigma));
              if (h < 2.220446049250313e-16) h = 0.0;       // eps * max, max = 1 at the centre
              v = (float)h;
            }
            // values are >= 0: the integer image of the float orders like the float
            atomicMax(reinterpret_cast<int*>(hm + (size_t)(iy + yy) * a.W + (ix + xx)), __float_as_int(v));
          }
      }
    }
    a.inds[(size_t)b * a.max_objs + k] = (long long)iy * a.W + ix;
    a.mask[(size_t)b * a.max_objs + k] = 1.0f;
    float* rb = a.ret_boxes + ((size_t)b * a.max_objs + k) * a.C;
    rb[0] = fsub(cx, (float)ix);
    rb[1] = fsub(cy, (float)iy);
    rb[2] = z;
    rb[3] = logf(g[3]);
    rb[4] = logf(g[4]);
    rb[5] = logf(g[5]);
    rb[6] = cosf(g[6]);
    rb[7] = sinf(g[6]);
    for (int c = 8; c < a.C; ++c) rb[c] = g[c - 1];       // ret_boxes[8:] = gt[7:-1]
    long long* rm = a.radius_map + ((size_t)b * a.max_objs + k) * a.R;
    rm[0] = cls;
    rm[1] = ix;
    rm[2] = iy;
    rm[3] = radius;
    if (a.R >= 5) rm[4] = a.group[(size_t)b * a.M + m];
  }
}

// CurriculumCenterHead.cluster: the curriculum group of every ground-truth box (0 = none)
__global__ void __launch_bounds__(256) cluster_groups_kernel(const float* __restrict__ gt, int n, int C,
                                                              const float* __restrict__ true_object,
                                                              const float* __restrict__ occupancy,
                                                              const float* __restrict__ facade, long long* __restrict__ group) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* g = gt + (size_t)i * C;
  const float x = g[0], y = g[1], length = g[3], cls = g[C - 1];
  const float dist = __fsqrt_rn(fadd(fmul(x, x), fmul(y, y)));
  const float occ = occupancy[i], fac = facade[i];
  long long out = 0;
  if (true_object[i] == 1.0f) {
    int di = -1;
    if (dist <= 30.0f) di = 0;
    else if (dist > 30.0f && dist <= 50.0f) di = 1;
    else if (dist > 50.0f) di = 2;
    if (cls == 1.0f) {
      int li = -1, fi = -1, oi = -1;
      if (length <= 6.0f) li = 0;
      else if (length > 6.0f) li = 1;
      if (fac == 3.0f) fi = 0;
      else if (fac == 2.0f) fi = 1;
      else if (fac == 1.0f) fi = 2;
      else if (fac == 0.0f) fi = 3;
      if (occ > 0.7f) oi = 0;
      else if (occ <= 0.7f && occ > 0.5f) oi = 1;
      else if (occ <= 0.5f && occ > 0.25f) oi = 2;
      else if (occ <= 0.25f) oi = 3;
      if (di >= 0 && li >= 0 && fi >= 0 && oi >= 0) out = ((di * 2 + li) * 4 + fi) * 4 + oi + 1;
    } else if (cls == 2.0f || cls == 3.0f) {
      // thresholds 0.21 * 5 / 12 ... evaluated in float64 by Python, compared in fp32 by torch
      const float t21 = (float)(0.21 * 5 / 12), t41 = (float)(0.41 * 5 / 12), t61 = (float)(0.61 * 5 / 12),
                  t81 = (float)(0.81 * 5 / 12);
      int oi = -1;
      if (occ > t81) oi = 0;
      else if (occ <= t81 && occ > t61) oi = 1;
      else if (occ <= t61 && occ > t41) oi = 2;
      else if (occ <= t41 && occ > t21) oi = 3;
      else if (occ <= t21) oi = 4;
      if (di >= 0 && oi >= 0) out = di * 5 + oi + 1;
    }
  }
  group[i] = out;
}

// confidence_of_all_groups: conf[c][g-1] = sum of pred at the centres of the objects of class c and group g, num = count.
// One thread per (class, group), objects visited in (frame, object) order: deterministic.
__global__ void __launch_bounds__(256) group_confidence_kernel(const float* __restrict__ pred, int B, int Ch, int H, int W,
                                                                const long long* __restrict__ rmap, int nobj, int R,
                                                                int n_class, int n_group, float* __restrict__ conf,
                                                                float* __restrict__ num) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_class * n_group) return;
  const int c = t / n_group, gsel = t % n_group + 1;
  float s = 0.0f;
  int cnt = 0;
  for (int b = 0; b < B; ++b)
    for (int k = 0; k < nobj; ++k) {
      const long long* rm = rmap + ((size_t)b * nobj + k) * R;
      if (rm[R - 1] == gsel && rm[0] == c) {
        if (c < Ch) s += pred[(((size_t)b * Ch + c) * H + rm[2]) * W + rm[1]];
        ++cnt;
      }
    }
  conf[t] = s;
  num[t] = (float)cnt;
}

struct ReweightArgs {
  const float* pred;        // [B][Ch][H][W]  clamped sigmoid
  const long long* rmap;    // [B][nobj][R]
  int B, Ch, H, W, nobj, R;
  double threshold, elongation, height, K;
  int mode;                 // 0 logistic (default), 1 straight, 2 tuning
  int fixed_radius, add_radius, only_center, active;
  float* box_mask;          // [B][nobj]       in/out (a clone of the assignment mask)
  float* mask;              // [B][Ch][H][W]   in/out (heatmap_mask)
};

// one block per frame; objects in index order (a later object's window overwrites an earlier one's)
__global__ void __launch_bounds__(256) comloss_reweight_kernel(ReweightArgs a) {
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int k = 0; k < a.nobj; ++k) {
    const long long* rm = a.rmap + ((size_t)b * a.nobj + k) * a.R;
    if (rm[3] <= 0) continue;                      // uniform over the block
    const int cls = (int)rm[0], cx = (int)rm[1], cy = (int)rm[2];
    const int radius = a.fixed_radius != 0 ? a.fixed_radius : (int)rm[3] + a.add_radius;
    const double conf = (double)a.pred[(((size_t)b * a.Ch + cls) * a.H + cy) * a.W + cx];
    double w;
    if (a.mode == 1) w = a.K * (conf - a.threshold) + 1.0;
    else if (a.mode == 2) w = 1.0;
    else w = a.height / (1.0 + exp(a.elongation * (conf - a.threshold))) + 1.0 - a.height / 2.0;
    if (!a.active) continue;
    const float wf = (float)w;
    if (tid == 0) a.box_mask[(size_t)b * a.nobj + k] = wf;
    float* mk = a.mask + ((size_t)b * a.Ch + cls) * a.H * a.W;
    if (a.only_center) {
      if (tid == 0) mk[(size_t)cy * a.W + cx] = wf;
    } else {
      // draw_mask_to_heatmap (centernet_utils.py:110-131): assignment of the constant window
      const int left = cx < radius ? cx : radius, right = a.W - cx < radius + 1 ? a.W - cx : radius + 1;
      const int top = cy < radius ? cy : radius, bottom = a.H - cy < radius + 1 ? a.H - cy : radius + 1;
      const int ww = left + right, hh = top + bottom;
      if (ww > 0 && hh > 0)
        for (int e = tid; e < ww * hh; e += blockDim.x) {
          const int yy = e / ww - top, xx = e % ww - left;
          mk[(size_t)(cy + yy) * a.W + (cx + xx)] = wf;
        }
    }
    __syncthreads();       // the next object may overwrite these cells
  }
}

}  // namespace
}  // namespace comb

using namespace comb;

extern "C" int comb_centerhead_assign_targets(const float* gt_boxes, const float* npgt, const long long* group,
                                              const int* cls_map, int n_cls, int B, int M, int C, float x0, float y0,
                                              float vx, float vy, float stride, int W, int H, int max_objs,
                                              double overlap, int min_radius, int filter_points, float min_points,
                                              const float* gtab, const int* gtab_off, int rmax, int Ch, float* heatmap,
                                              float* ret_boxes, long long* inds, float* mask, long long* radius_map,
                                              int R, int relabel_in_place, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(B >= 0 && M >= 0 && C >= 8, "comb_centerhead_assign_targets: gt_boxes must be (B, M, >= 8)");
  COMB_CHECK_ARG(max_objs >= 1 && max_objs <= kMaxObjsCap, "comb_centerhead_assign_targets: NUM_MAX_OBJS %d outside [1,%d]",
                 max_objs, kMaxObjsCap);
  COMB_CHECK_ARG(R == 4 || R == 5, "comb_centerhead_assign_targets: radius_map has 4 or 5 columns");
  if (B == 0 || M == 0) return COMB_OK;           // no ground truth at all: the zero-filled outputs are the answer
  COMB_CHECK_ARG(R == 4 || group != nullptr, "comb_centerhead_assign_targets: 5 columns need the group tensor");
  COMB_CHECK_ARG(gt_boxes && npgt && cls_map && heatmap && ret_boxes && inds && mask && radius_map && gtab && gtab_off,
                 "comb_centerhead_assign_targets: null pointer");
  AssignArgs a;
  a.gt = gt_boxes; a.gt_rw = const_cast<float*>(gt_boxes); a.relabel = relabel_in_place; a.npgt = npgt; a.group = group; a.cls_map = cls_map; a.n_cls = n_cls;
  a.B = B; a.M = M; a.C = C;
  a.x0 = x0; a.y0 = y0; a.vx = vx; a.vy = vy; a.stride = stride;
  a.W = W; a.H = H; a.max_objs = max_objs;
  // Python evaluates the scalar expressions in float64, torch casts each scalar to fp32 at the operation
  a.ov_minus = (float)(1.0 - overlap);
  a.ov_plus = (float)(1.0 + overlap);
  a.a3 = (float)(4.0 * overlap);
  a.b3c = (float)(-2.0 * overlap);
  a.c3c = (float)(overlap - 1.0);
  a.a3x4 = (float)(4.0 * (4.0 * overlap));
  a.min_radius = min_radius; a.filter_points = filter_points; a.min_points = min_points;
  a.gtab = gtab; a.gtab_off = gtab_off; a.rmax = rmax;
  a.heatmap = heatmap; a.Ch = Ch; a.ret_boxes = ret_boxes; a.inds = inds; a.mask = mask; a.radius_map = radius_map; a.R = R;
  assign_targets_kernel<<<B, kTgtThreads, 0, stream>>>(a);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

extern "C" int comb_centerhead_cluster_groups(const float* gt_boxes, int n, int C, const float* true_object,
                                              const float* occupancy_ratio, const float* facade_type, long long* group,
                                              void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(n >= 0 && C >= 8, "comb_centerhead_cluster_groups: gt_boxes must be (n, >= 8)");
  if (n == 0) return COMB_OK;
  COMB_CHECK_ARG(gt_boxes && true_object && occupancy_ratio && facade_type && group, "comb_centerhead_cluster_groups: null pointer");
  cluster_groups_kernel<<<cdiv(n, 256), 256, 0, stream>>>(gt_boxes, n, C, true_object, occupancy_ratio, facade_type, group);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

extern "C" int comb_comloss_group_confidence(const float* pred, int B, int Ch, int H, int W, const long long* radius_map,
                                             int nobj, int R, int n_class, int n_group, float* conf, float* num,
                                             void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(n_class >= 1 && n_group >= 1 && R >= 4, "comb_comloss_group_confidence: bad shape");
  COMB_CHECK_ARG(pred && radius_map && conf && num, "comb_comloss_group_confidence: null pointer");
  group_confidence_kernel<<<cdiv(n_class * n_group, 256), 256, 0, stream>>>(pred, B, Ch, H, W, radius_map, nobj, R, n_class,
                                                                             n_group, conf, num);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

extern "C" int comb_comloss_reweight(const float* pred, int B, int Ch, int H, int W, const long long* radius_map, int nobj,
                                     int R, double threshold, double elongation, double height, double K, int mode,
                                     int fixed_radius, int add_radius, int only_center, int active, float* box_mask,
                                     float* mask, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(R >= 4 && mode >= 0 && mode <= 2, "comb_comloss_reweight: bad arguments");
  if (B == 0 || nobj == 0) return COMB_OK;
  COMB_CHECK_ARG(pred && radius_map && box_mask && mask, "comb_comloss_reweight: null pointer");
  ReweightArgs a;
  a.pred = pred; a.rmap = radius_map; a.B = B; a.Ch = Ch; a.H = H; a.W = W; a.nobj = nobj; a.R = R;
  a.threshold = threshold; a.elongation = elongation; a.height = height; a.K = K; a.mode = mode;
  a.fixed_radius = fixed_radius; a.add_radius = add_radius; a.only_center = only_center; a.active = active;
  a.box_mask = box_mask; a.mask = mask;
  comloss_reweight_kernel<<<B, 256, 0, stream>>>(a);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}
