// a11-a16 — box ops used by COMAug GT sampling and CenterHead post-processing.
//
// Replaces (all under pcdet/ops/):
//   points_in_boxes_cpu      roiaware_pool3d/src/roiaware_pool3d.cpp:121-168   (MARGIN 1e-2, (Nb,P) 0/1 mask)
//   points_in_boxes_gpu      roiaware_pool3d/src/roiaware_pool3d_kernel.cu:15-36,313-336 (MARGIN 1e-5, first hit)
//   boxes_iou_bev_cpu        iou3d_nms/src/iou3d_cpu.cpp:128-252
//   boxes_iou_bev_gpu / boxes_overlap_bev_gpu   iou3d_nms/src/iou3d_nms.cpp:49-88, iou3d_nms_kernel.cu:104-265
//   nms_gpu / nms_normal_gpu iou3d_nms/src/iou3d_nms.cpp:90-188, iou3d_nms_kernel.cu:267-372
//
// Two arithmetic "flavours" exist because the reference has two implementations whose results are
// not bit-identical to each other:
//   CPUF = true : the g++/x86-64 build — IEEE fp32, NO fused multiply-add, glibc cosf/sinf.  The
//                 per-box trigonometry is taken from the caller (host libm, O(boxes) work); every
//                 product/sum below goes through __fmul_rn/__fadd_rn so nvcc cannot contract.
//   CPUF = false: the nvcc build — device cosf/sinf and default contraction; expressions are
//                 written in the reference's association order so nvcc contracts them alike.
// The geometry itself (16 edge/edge intersections, 8 corner-in-box tests, angular sort, shoelace)
// is restated here per pair with per-box work (corners, extents, rotation) hoisted into a
// shared-memory record computed once per block.
#include <math.h>
#include "common.cuh"

namespace comb {
namespace {

template <bool CPUF>
struct Ar {
  static __device__ __forceinline__ float mul(float a, float b) { return CPUF ? __fmul_rn(a, b) : a * b; }
  static __device__ __forceinline__ float add(float a, float b) { return CPUF ? __fadd_rn(a, b) : a + b; }
  static __device__ __forceinline__ float sub(float a, float b) { return CPUF ? __fsub_rn(a, b) : a - b; }
  static __device__ __forceinline__ float div(float a, float b) { return CPUF ? __fdiv_rn(a, b) : a / b; }
};

struct P2 {
  float x, y;
};

// Per-box record shared by all pairs that involve the box.
struct BoxRec {
  P2 c[4];        // rotated corners
  float cx, cy;   // centre
  float ncos, nsin;  // cos(-rz), sin(-rz)
  float hx, hy;   // dx/2 + MARGIN, dy/2 + MARGIN (fp32, MARGIN 1e-2)
  float area;     // dx*dy
  float rad;      // conservative radius for the exact early-out
};

template <bool CPUF>
__device__ __forceinline__ void make_box(const float* __restrict__ b, const float* __restrict__ trig, BoxRec& r) {
  using A = Ar<CPUF>;
  const float x = b[0], y = b[1], dx = b[3], dy = b[4], rz = b[6];
  float cs, sn, ncs, nsn;
  if (CPUF) {
    cs = trig[0]; sn = trig[1]; ncs = trig[2]; nsn = trig[3];
  } else {
    cs = cosf(rz); sn = sinf(rz); ncs = cosf(-rz); nsn = sinf(-rz);
  }
  const float hx = dx / 2, hy = dy / 2;
  const float x1 = A::sub(x, hx), y1 = A::sub(y, hy), x2 = A::add(x, hx), y2 = A::add(y, hy);
  const float px[4] = {x1, x2, x2, x1}, py[4] = {y1, y1, y2, y2};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    // (p.x - c.x)*cos + (p.y - c.y)*(-sin) + c.x ; (p.x - c.x)*sin + (p.y - c.y)*cos + c.y
    const float ux = A::sub(px[k], x), uy = A::sub(py[k], y);
    r.c[k].x = A::add(A::add(A::mul(ux, cs), A::mul(uy, -sn)), x);
    r.c[k].y = A::add(A::add(A::mul(ux, sn), A::mul(uy, cs)), y);
  }
  r.cx = x; r.cy = y;
  r.ncos = ncs; r.nsin = nsn;
  const float MARGIN = 1e-2f;
  r.hx = A::add(hx, MARGIN);
  r.hy = A::add(hy, MARGIN);
  r.area = A::mul(dx, dy);
  // any point of the margin-inflated box is within rad of the centre (1% + 1cm slack on top)
  r.rad = sqrtf(r.hx * r.hx + r.hy * r.hy) * 1.01f + 0.02f;
}

template <bool CPUF>
__device__ __forceinline__ float cross3(const P2& p1, const P2& p2, const P2& p0) {
  using A = Ar<CPUF>;
  return A::sub(A::mul(A::sub(p1.x, p0.x), A::sub(p2.y, p0.y)), A::mul(A::sub(p2.x, p0.x), A::sub(p1.y, p0.y)));
}

__device__ __forceinline__ float fminr(float a, float b) { return a > b ? b : a; }
__device__ __forceinline__ float fmaxr(float a, float b) { return a > b ? a : b; }

// segment p0->p1 against q0->q1; writes the crossing into ans when they properly intersect
template <bool CPUF>
__device__ __forceinline__ bool seg_cross(const P2& p1, const P2& p0, const P2& q1, const P2& q0, P2& ans) {
  using A = Ar<CPUF>;
  const bool boxes_touch = fminr(p0.x, p1.x) <= fmaxr(q0.x, q1.x) && fminr(q0.x, q1.x) <= fmaxr(p0.x, p1.x) &&
                           fminr(p0.y, p1.y) <= fmaxr(q0.y, q1.y) && fminr(q0.y, q1.y) <= fmaxr(p0.y, p1.y);
  if (!boxes_touch) return false;
  const float s1 = cross3<CPUF>(q0, p1, p0);
  const float s2 = cross3<CPUF>(p1, q1, p0);
  const float s3 = cross3<CPUF>(p0, q1, q0);
  const float s4 = cross3<CPUF>(q1, p1, q0);
  if (!(A::mul(s1, s2) > 0 && A::mul(s3, s4) > 0)) return false;
  const float s5 = cross3<CPUF>(q1, p1, p0);
  const float EPS = 1e-8f;
  if (fabsf(A::sub(s5, s1)) > EPS) {
    ans.x = A::div(A::sub(A::mul(s5, q0.x), A::mul(s1, q1.x)), A::sub(s5, s1));
    ans.y = A::div(A::sub(A::mul(s5, q0.y), A::mul(s1, q1.y)), A::sub(s5, s1));
  } else {
    const float a0 = A::sub(p0.y, p1.y), b0 = A::sub(p1.x, p0.x), c0 = A::sub(A::mul(p0.x, p1.y), A::mul(p1.x, p0.y));
    const float a1 = A::sub(q0.y, q1.y), b1 = A::sub(q1.x, q0.x), c1 = A::sub(A::mul(q0.x, q1.y), A::mul(q1.x, q0.y));
    const float D = A::sub(A::mul(a0, b1), A::mul(a1, b0));
    ans.x = A::div(A::sub(A::mul(b0, c1), A::mul(b1, c0)), D);
    ans.y = A::div(A::sub(A::mul(a1, c0), A::mul(a0, c1)), D);
  }
  return true;
}

template <bool CPUF>
__device__ __forceinline__ bool corner_inside(const BoxRec& box, const P2& p) {
  using A = Ar<CPUF>;
  const float ux = A::sub(p.x, box.cx), uy = A::sub(p.y, box.cy);
  const float rx = A::add(A::mul(ux, box.ncos), A::mul(uy, -box.nsin));
  const float ry = A::add(A::mul(ux, box.nsin), A::mul(uy, box.ncos));
  return fabsf(rx) < box.hx && fabsf(ry) < box.hy;
}

template <bool CPUF>
__device__ float overlap_area(const BoxRec& a, const BoxRec& b) {
  using A = Ar<CPUF>;
  {  // exact early-out: far apart => no crossing, no contained corner => area +0
    const float ddx = a.cx - b.cx, ddy = a.cy - b.cy, rr = a.rad + b.rad;
    if (ddx * ddx + ddy * ddy > rr * rr) return 0.0f;
  }
  P2 pts[16];
  float ang[16];
  int cnt = 0;
  float sx = 0.0f, sy = 0.0f;
#pragma unroll 1
  for (int i = 0; i < 4; ++i) {
    const P2 a0 = a.c[i], a1 = a.c[(i + 1) & 3];
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      P2 hit;
      if (seg_cross<CPUF>(a1, a0, b.c[(j + 1) & 3], b.c[j], hit)) {
        sx = A::add(sx, hit.x);
        sy = A::add(sy, hit.y);
        pts[cnt++] = hit;
      }
    }
  }
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    if (corner_inside<CPUF>(a, b.c[k])) {
      sx = A::add(sx, b.c[k].x);
      sy = A::add(sy, b.c[k].y);
      pts[cnt++] = b.c[k];
    }
    if (corner_inside<CPUF>(b, a.c[k])) {
      sx = A::add(sx, a.c[k].x);
      sy = A::add(sy, a.c[k].y);
      pts[cnt++] = a.c[k];
    }
  }
  if (cnt == 0) return 0.0f;
  const float mx = A::div(sx, (float)cnt), my = A::div(sy, (float)cnt);
  for (int i = 0; i < cnt; ++i) ang[i] = atan2f(A::sub(pts[i].y, my), A::sub(pts[i].x, mx));
  // bubble sort ascending by angle (same swap rule as the reference so ties keep their order)
  for (int j = 0; j < cnt - 1; ++j)
    for (int i = 0; i < cnt - j - 1; ++i)
      if (ang[i] > ang[i + 1]) {
        const P2 tp = pts[i]; pts[i] = pts[i + 1]; pts[i + 1] = tp;
        const float ta = ang[i]; ang[i] = ang[i + 1]; ang[i + 1] = ta;
      }
  float area = 0.0f;
  for (int k = 0; k < cnt - 1; ++k) {
    const float ax = A::sub(pts[k].x, pts[0].x), ay = A::sub(pts[k].y, pts[0].y);
    const float bx = A::sub(pts[k + 1].x, pts[0].x), by = A::sub(pts[k + 1].y, pts[0].y);
    area = A::add(area, A::sub(A::mul(ax, by), A::mul(ay, bx)));
  }
  return fabsf(area) / 2.0f;
}

template <bool CPUF>
__device__ __forceinline__ float iou_from_overlap(const BoxRec& a, const BoxRec& b, float ov) {
  using A = Ar<CPUF>;
  return A::div(ov, fmaxf(A::sub(A::add(a.area, b.area), ov), 1e-8f));
}

// ---------------------------------------------------------------- pairwise matrix
// Persistent blocks walk 16 x 128 (64 x 128 for large matrices) tiles of the (Na, Nb) matrix.  Per 16-row step:
//   A. every thread runs the exact far-apart test on 8 consecutive columns of one row, writes the zeros of the step
//      with 16-byte stores and appends the (i, j) of the pairs that may overlap to a queue in shared memory;
//   B. whenever the queue holds a whole round (256 pairs) the block evaluates it with EVERY lane busy: the expensive
//      geometry (16 segment crossings, 8 corner tests, angular sort, shoelace: ~3000 instructions with local-memory
//      arrays) runs only for queued pairs, and a round is never paid for a handful of them — the queue carries over
//      from tile to tile and is flushed once at the end.
// (r1 history on the COMAug 10k x 10k case, 0.4 % of the pairs overlap: one thread per pair 1.55 ms — 12 % of the warps
// held one near pair and ran the geometry for a single lane; per-tile queues 0.75 ms — every tile still paid one
// round for its ~55 queued pairs.  r2: clearing the output with a memset first and writing only the near pairs from
// the kernel, with 32 rows per barrier round, is NOT faster — 0.547 vs 0.51 ms, bit-identical — so the zero stores are
// not what bounds this kernel; what is left is the geometry rounds and their imbalance between blocks.)
constexpr int kTB = 128, kPairsPerThread = 8, kRound = 256;
constexpr int kQueueCap = kRound + 16 * kTB;      // a 16-row step adds at most 16 x 128 pairs to a remainder < 256

template <bool CPUF>
__device__ __forceinline__ float4 centre_rad(const float* __restrict__ b) {
  using A = Ar<CPUF>;
  // the same conservative radius as make_box(): any point of the margin-inflated box is within rad of the centre
  const float hx = A::add(b[3] / 2, 1e-2f), hy = A::add(b[4] / 2, 1e-2f);
  return make_float4(b[0], b[1], sqrtf(hx * hx + hy * hy) * 1.01f + 0.02f, 0.0f);
}

template <bool CPUF>
__device__ __forceinline__ void eval_pair(const float* __restrict__ boxes_a, const float* __restrict__ trig_a,
                                          const float* __restrict__ boxes_b, const float* __restrict__ trig_b, int nb,
                                          int what, uint2 e, float* __restrict__ out) {
  BoxRec ra, rb;
  make_box<CPUF>(boxes_a + (size_t)e.x * 7, CPUF ? trig_a + (size_t)e.x * 4 : nullptr, ra);
  make_box<CPUF>(boxes_b + (size_t)e.y * 7, CPUF ? trig_b + (size_t)e.y * 4 : nullptr, rb);
  const float ov = overlap_area<CPUF>(ra, rb);
  out[(size_t)e.x * nb + e.y] = what ? ov : iou_from_overlap<CPUF>(ra, rb, ov);
}

template <bool CPUF, int kTA>
__global__ void __launch_bounds__(kRound) boxes_bev_kernel(const float* __restrict__ boxes_a,
                                                            const float* __restrict__ trig_a, int na,
                                                            const float* __restrict__ boxes_b,
                                                            const float* __restrict__ trig_b, int nb, int what,
                                                            float* __restrict__ out, int vec_ok, int tiles_x,
                                                            int ntiles) {
  static_assert(kTA % 16 == 0 && kTA + kTB <= kRound, "one thread prepares one box of the tile");
  __shared__ float4 sa[kTA], sb[kTB];       // centre x, y, conservative radius
  __shared__ uint2 queue[kQueueCap];
  __shared__ int qn;
  const int tid = threadIdx.x;
  if (tid == 0) qn = 0;
  const int c0 = (tid & 15) * kPairsPerThread;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int a0 = (tile / tiles_x) * kTA, b0 = (tile % tiles_x) * kTB;
    __syncthreads();                        // the previous tile's readers of sa / sb are done (and qn = 0 is visible)
    if (tid < kTB) {
      if (b0 + tid < nb) sb[tid] = centre_rad<CPUF>(boxes_b + (size_t)(b0 + tid) * 7);
    } else if (tid < kTB + kTA) {
      if (a0 + tid - kTB < na) sa[tid - kTB] = centre_rad<CPUF>(boxes_a + (size_t)(a0 + tid - kTB) * 7);
    }
    __syncthreads();
#pragma unroll 1
    for (int r = tid >> 4; r < kTA; r += 16) {           // uniform trip count: every thread reaches the barriers
      // ---- phase A: far-apart test + zeros for 16 rows x 128 columns
      const int i = a0 + r;
      if (i < na && b0 + c0 < nb) {
        const float4 ra = sa[r];
        unsigned near = 0;
#pragma unroll
        for (int u = 0; u < kPairsPerThread; ++u) {
          if (b0 + c0 + u < nb) {
            // the same exact early-out as overlap_area: further apart than the sum of the conservative radii => +0
            const float4 rb = sb[c0 + u];
            const float ddx = ra.x - rb.x, ddy = ra.y - rb.y, rr = ra.z + rb.z;
            if (!(ddx * ddx + ddy * ddy > rr * rr)) near |= 1u << u;
          }
        }
        float* o = out + (size_t)i * nb + b0 + c0;
        if (vec_ok && b0 + c0 + kPairsPerThread <= nb) {
          reinterpret_cast<float4*>(o)[0] = make_float4(0.f, 0.f, 0.f, 0.f);
          reinterpret_cast<float4*>(o)[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
#pragma unroll
          for (int u = 0; u < kPairsPerThread; ++u)
            if (b0 + c0 + u < nb) o[u] = 0.0f;
        }
        if (near) {
          const int base = atomicAdd(&qn, __popc(near));
          int k = 0;
#pragma unroll
          for (int u = 0; u < kPairsPerThread; ++u)
            if (near & (1u << u)) queue[base + k++] = make_uint2((unsigned)i, (unsigned)(b0 + c0 + u));
        }
      }
      // ---- phase B: whole rounds only (taken from the top of the queue)
      __syncthreads();                      // the pushes (and the zeros) of this step are done
      const int n = qn;
      const int take = n & ~(kRound - 1);
      for (int q = tid; q < take; q += kRound)
        eval_pair<CPUF>(boxes_a, trig_a, boxes_b, trig_b, nb, what, queue[n - take + q], out);
      __syncthreads();                      // everybody has read qn and its queue entries
      if (tid == 0) qn = n - take;
      __syncthreads();
    }
  }
  __syncthreads();
  // ---- flush: the remainder (< 256 pairs, or everything when the block saw less than one round)
  const int n = qn;
  for (int q = tid; q < n; q += kRound)
    eval_pair<CPUF>(boxes_a, trig_a, boxes_b, trig_b, nb, what, queue[q], out);
}

// ---------------------------------------------------------------- NMS
__device__ __forceinline__ float iou_axis_aligned(const float* a, const float* b) {
  float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
  float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
  float width = fmaxf(right - left, 0.f), height = fmaxf(bottom - top, 0.f);
  float inter = width * height;
  float sa = a[3] * a[4], sb = b[3] * b[4];
  return inter / fmaxf(sa + sb - inter, 1e-8f);
}

// One thread per (row, column) PAIR of the upper triangle: block = 4 rows x 64 columns (256 threads), the 64 column
// boxes and the 4 row boxes are prepared once in shared memory, every warp assembles 32 bits of a suppression word
// with one ballot.  (r1: one thread per row looping over 64 columns — the reference's shape — left the 500-box NMS at
// 64 blocks x 64 threads and ~0.25 ms.)
template <bool CPUF, bool ROT>
__global__ void __launch_bounds__(256) nms_mask_kernel(const float* __restrict__ boxes, const float* __restrict__ trig,
                                                        int n_max, const int* __restrict__ n_dev, float thresh,
                                                        unsigned long long* __restrict__ mask, int col_blocks) {
  const int n = eff_n(n_max, n_dev);
  const int cb = blockIdx.x, r0 = blockIdx.y * 4;
  if (cb < (r0 >> 6)) return;                       // whole block left of the diagonal (4 | 64: one row block per block)
  __shared__ BoxRec scol[64], srow[4];
  __shared__ float sraw[68 * 7];
  const int t = threadIdx.x;
  if (t < 68) {
    const int b = t < 64 ? cb * 64 + t : r0 + (t - 64);
    if (b < n) {
      if (ROT)
        make_box<CPUF>(boxes + (size_t)b * 7, CPUF ? trig + (size_t)b * 4 : nullptr, t < 64 ? scol[t] : srow[t - 64]);
      else
        for (int q = 0; q < 7; ++q) sraw[t * 7 + q] = boxes[(size_t)b * 7 + q];
    }
  }
  __syncthreads();
  const int rl = t >> 6, jl = t & 63;
  const int ri = r0 + rl, cj = cb * 64 + jl;
  bool hit = false;
  if (ri < n && cj < n && cj > ri) {
    if (ROT) {
      const float ov = overlap_area<CPUF>(srow[rl], scol[jl]);
      hit = iou_from_overlap<CPUF>(srow[rl], scol[jl], ov) > thresh;
    } else {
      hit = iou_axis_aligned(sraw + (64 + rl) * 7, sraw + jl * 7) > thresh;
    }
  }
  const unsigned bits = __ballot_sync(0xffffffffu, hit);
  if ((t & 31) == 0 && ri < n)
    reinterpret_cast<unsigned*>(mask)[((size_t)ri * col_blocks + cb) * 2 + ((t >> 5) & 1)] = bits;
}

// Greedy sweep (iou3d_nms.cpp:121-133) on the device: one warp, suppression words in shared memory.
__global__ void __launch_bounds__(32) nms_sweep_kernel(const unsigned long long* __restrict__ mask, int n_max,
                                                        const int* __restrict__ n_dev, int col_blocks,
                                                        long long* __restrict__ keep, int* __restrict__ num_keep) {
  extern __shared__ unsigned long long remv[];
  const int n = eff_n(n_max, n_dev);
  const int lane = threadIdx.x;
  for (int j = lane; j < col_blocks; j += 32) remv[j] = 0ull;
  __syncwarp();
  int kept = 0;
  for (int i = 0; i < n; ++i) {
    const int nblock = i >> 6, inblock = i & 63;
    const unsigned long long w = remv[nblock];
    if (!((w >> inblock) & 1ull)) {
      if (lane == 0) keep[kept] = i;
      ++kept;
      const unsigned long long* row = mask + (size_t)i * col_blocks;
      for (int j = nblock + lane; j < col_blocks; j += 32) remv[j] |= row[j];
    }
    __syncwarp();
  }
  if (lane == 0) *num_keep = kept;
}

// Same sweep with the whole suppression matrix staged in shared memory first (n * col_blocks * 8 bytes <= ~200 KB, i.e.
// n <= ~1200: the CenterHead case of <= 500 boxes per frame is 32 KB).  The global-memory form above pays one
// dependent L2 round trip per KEPT box (r1: 0.36 ms for 500 boxes at thresh 0.7, almost all of it the sweep).
__global__ void __launch_bounds__(256) nms_sweep_smem_kernel(const unsigned long long* __restrict__ mask, int n_max,
                                                              const int* __restrict__ n_dev, int col_blocks,
                                                              long long* __restrict__ keep, int* __restrict__ num_keep) {
  extern __shared__ unsigned long long sm[];      // [col_blocks] removed bits, then [n][col_blocks] matrix
  const int n = eff_n(n_max, n_dev);
  unsigned long long* remv = sm;
  unsigned long long* rows = sm + col_blocks;
  const int total = n * col_blocks;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int i = e / col_blocks, j = e - i * col_blocks;
    if (j >= (i >> 6)) rows[e] = __ldg(mask + e);        // the mask kernel only writes the upper triangle
  }
  for (int j = threadIdx.x; j < col_blocks; j += blockDim.x) remv[j] = 0ull;
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  int kept = 0;
  if (col_blocks <= 32) {
    // removed bits in REGISTERS (lane j owns word j), the tested word by shuffle, row i+1 prefetched from shared
    // memory while row i is decided: the serial chain per box is a shuffle and a few ALU ops
    unsigned long long myrem = 0ull;
    unsigned long long nextw = (lane < col_blocks) ? rows[lane] : 0ull;
    for (int i = 0; i < n; ++i) {
      const unsigned long long cur = nextw;
      if (i + 1 < n) nextw = (lane < col_blocks && lane >= ((i + 1) >> 6)) ? rows[(size_t)(i + 1) * col_blocks + lane] : 0ull;
      const unsigned long long w = __shfl_sync(0xffffffffu, myrem, i >> 6);
      if (!((w >> (i & 63)) & 1ull)) {
        if (lane == 0) keep[kept] = i;
        ++kept;
        if (lane >= (i >> 6)) myrem |= cur;
      }
    }
    if (lane == 0) *num_keep = kept;
    return;
  }
  for (int i = 0; i < n; ++i) {
    const int nblock = i >> 6, inblock = i & 63;
    const unsigned long long w = remv[nblock];
    if (!((w >> inblock) & 1ull)) {
      if (lane == 0) keep[kept] = i;
      ++kept;
      const unsigned long long* row = rows + (size_t)i * col_blocks;
      for (int j = nblock + lane; j < col_blocks; j += 32) remv[j] |= row[j];
    }
    __syncwarp();
  }
  if (lane == 0) *num_keep = kept;
}

// ---------------------------------------------------------------- points in boxes
struct PibBox {
  float cx, cy, cz, tz;  // tz: |z-cz| > tz rejects
  float ncos, nsin, tx, ty;  // |lx| < tx && |ly| < ty accepts
};

// fp64 thresholds of the reference folded to fp32 exactly:
//   float a <  double t  <=>  a <  (float)t rounded up      (dx/2.0 + MARGIN)
//   float a >  double t  <=>  a >  (float)t rounded down    (dz/2.0)
__device__ __forceinline__ void make_pib(const float* __restrict__ b, float ncos, float nsin, float margin_f,
                                         PibBox& r) {
  r.cx = b[0]; r.cy = b[1]; r.cz = b[2];
  const double m = (double)margin_f;
  r.tx = __double2float_ru((double)b[3] / 2.0 + m);
  r.ty = __double2float_ru((double)b[4] / 2.0 + m);
  r.tz = __double2float_rd((double)b[5] / 2.0);
  r.ncos = ncos; r.nsin = nsin;
}

template <bool CPUF>
__device__ __forceinline__ bool pt_in_box(float x, float y, float z, const PibBox& b) {
  using A = Ar<CPUF>;
  if (fabsf(A::sub(z, b.cz)) > b.tz) return false;
  const float sx = A::sub(x, b.cx), sy = A::sub(y, b.cy);
  const float lx = A::add(A::mul(sx, b.ncos), A::mul(sy, -b.nsin));
  const float ly = A::add(A::mul(sx, b.nsin), A::mul(sy, b.ncos));
  return (fabsf(lx) < b.tx) & (fabsf(ly) < b.ty);
}

constexpr int kPibBoxes = 64;   // boxes per block (shared memory)
constexpr int kPibPts = 4;      // points per thread

// grid (point tiles, box tiles); each thread keeps kPibPts points in registers and streams the
// block's boxes from shared memory; mask rows are written coalesced along the point axis.
__global__ void __launch_bounds__(256) pib_mask_kernel(const float* __restrict__ points, int P, int pstride,
                                                        const float* __restrict__ boxes,
                                                        const float* __restrict__ trig, int nb,
                                                        int* __restrict__ mask) {
  __shared__ PibBox sb[kPibBoxes];
  const int b0 = blockIdx.y * kPibBoxes;
  const int nbl = min(kPibBoxes, nb - b0);
  if (threadIdx.x < nbl) {
    const int b = b0 + threadIdx.x;
    make_pib(boxes + (size_t)b * 7, trig[b * 2], trig[b * 2 + 1], 1e-2f, sb[threadIdx.x]);
  }
  __syncthreads();
  const int p0 = blockIdx.x * (256 * kPibPts) + threadIdx.x;
  float x[kPibPts], y[kPibPts], z[kPibPts];
#pragma unroll
  for (int q = 0; q < kPibPts; ++q) {
    const int p = p0 + q * 256;
    if (p < P) {
      const float* pp = points + (size_t)p * pstride;
      x[q] = __ldg(pp); y[q] = __ldg(pp + 1); z[q] = __ldg(pp + 2);
    } else {
      x[q] = y[q] = z[q] = 0.f;
    }
  }
  for (int j = 0; j < nbl; ++j) {
    const PibBox bx = sb[j];
    int* row = mask + (size_t)(b0 + j) * P;
#pragma unroll
    for (int q = 0; q < kPibPts; ++q) {
      const int p = p0 + q * 256;
      if (p < P) __stcs(row + p, pt_in_box<true>(x[q], y[q], z[q], bx) ? 1 : 0);
    }
  }
}

// any-box form of the same test (remove_points_in_boxes3d, pcdet/utils/box_utils.py:117-131, only needs
// `mask.sum(0) != 0`): one byte per point instead of Nb ints, boxes walked in shared-memory tiles, a thread stops
// testing a point at its first hit.  Same arithmetic as pib_mask_kernel, so any[p] == OR_b mask[b][p] bit for bit.
__global__ void __launch_bounds__(256) pib_any_kernel(const float* __restrict__ points, int P, int pstride,
                                                       const float* __restrict__ boxes,
                                                       const float* __restrict__ trig, int nb,
                                                       unsigned char* __restrict__ any) {
  __shared__ PibBox sb[kPibBoxes];
  const int p0 = blockIdx.x * (256 * kPibPts) + threadIdx.x;
  float x[kPibPts], y[kPibPts], z[kPibPts];
  bool hit[kPibPts];
#pragma unroll
  for (int q = 0; q < kPibPts; ++q) {
    const int p = p0 + q * 256;
    hit[q] = false;
    if (p < P) {
      const float* pp = points + (size_t)p * pstride;
      x[q] = __ldg(pp); y[q] = __ldg(pp + 1); z[q] = __ldg(pp + 2);
    } else {
      x[q] = y[q] = z[q] = 0.f;
    }
  }
  for (int b0 = 0; b0 < nb; b0 += kPibBoxes) {
    const int nbl = min(kPibBoxes, nb - b0);
    __syncthreads();
    if (threadIdx.x < nbl) {
      const int b = b0 + threadIdx.x;
      make_pib(boxes + (size_t)b * 7, trig[b * 2], trig[b * 2 + 1], 1e-2f, sb[threadIdx.x]);
    }
    __syncthreads();
    for (int j = 0; j < nbl; ++j) {
      const PibBox bx = sb[j];
#pragma unroll
      for (int q = 0; q < kPibPts; ++q) hit[q] = hit[q] || pt_in_box<true>(x[q], y[q], z[q], bx);
    }
  }
#pragma unroll
  for (int q = 0; q < kPibPts; ++q) {
    const int p = p0 + q * 256;
    if (p < P) any[p] = hit[q] ? 1 : 0;
  }
}

// first-hit index, reference device arithmetic (MARGIN 1e-5, device trig, default contraction)
__global__ void __launch_bounds__(256) pib_index_kernel(const float* __restrict__ points,
                                                         const float* __restrict__ boxes, int P, int T,
                                                         int* __restrict__ idx) {
  extern __shared__ PibBox sbx[];
  const int f = blockIdx.y;
  const float* fb = boxes + (size_t)f * T * 7;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float rz = fb[(size_t)t * 7 + 6];
    make_pib(fb + (size_t)t * 7, cosf(-rz), sinf(-rz), 1e-5f, sbx[t]);
  }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float* pp = points + ((size_t)f * P + p) * 3;
  const float x = pp[0], y = pp[1], z = pp[2];
  for (int t = 0; t < T; ++t) {
    if (pt_in_box<false>(x, y, z, sbx[t])) {
      idx[(size_t)f * P + p] = t;
      return;
    }
  }
}

}  // namespace
}  // namespace comb

using namespace comb;

extern "C" void comb_box_trig_host(const float* boxes, int nb, float* trig) {
  for (int i = 0; i < nb; ++i) {
    const float rz = boxes[(size_t)i * 7 + 6];
    trig[i * 2 + 0] = cosf(-rz);
    trig[i * 2 + 1] = sinf(-rz);
  }
}

extern "C" void comb_box_trig4_host(const float* boxes, int n, float* trig) {
  for (int i = 0; i < n; ++i) {
    const float rz = boxes[(size_t)i * 7 + 6];
    trig[i * 4 + 0] = cosf(rz);
    trig[i * 4 + 1] = sinf(rz);
    trig[i * 4 + 2] = cosf(-rz);
    trig[i * 4 + 3] = sinf(-rz);
  }
}

extern "C" int comb_points_in_boxes_mask(const float* points, int P, int point_stride, const float* boxes,
                                         const float* box_trig, int nb, int* mask, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(P >= 0 && nb >= 0 && point_stride >= 3, "comb_points_in_boxes_mask: bad shape");
  if (P == 0 || nb == 0) return COMB_OK;
  COMB_CHECK_ARG(points && boxes && box_trig && mask, "comb_points_in_boxes_mask: null pointer");
  dim3 grid(cdiv(P, 256 * kPibPts), cdiv(nb, kPibBoxes));
  COMB_CHECK_ARG(grid.y <= 65535, "comb_points_in_boxes_mask: too many boxes (%d)", nb);
  pib_mask_kernel<<<grid, 256, 0, stream>>>(points, P, point_stride, boxes, box_trig, nb, mask);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

extern "C" int comb_points_in_any_box(const float* points, int P, int point_stride, const float* boxes,
                                     const float* box_trig, int nb, unsigned char* any, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(P >= 0 && nb >= 0 && point_stride >= 3, "comb_points_in_any_box: bad shape");
  if (P == 0) return COMB_OK;
  COMB_CHECK_ARG(points && any, "comb_points_in_any_box: null pointer");
  if (nb == 0) {
    COMB_CUDA(cudaMemsetAsync(any, 0, (size_t)P, stream));
    return COMB_OK;
  }
  COMB_CHECK_ARG(boxes && box_trig, "comb_points_in_any_box: null pointer");
  pib_any_kernel<<<cdiv(P, 256 * kPibPts), 256, 0, stream>>>(points, P, point_stride, boxes, box_trig, nb, any);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

extern "C" int comb_points_in_boxes_index(const float* points, const float* boxes, int batch, int P, int T, int* idx,
                                          void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(batch >= 0 && P >= 0 && T >= 0, "comb_points_in_boxes_index: bad shape");
  if (batch == 0 || P == 0) return COMB_OK;
  COMB_CHECK_ARG(idx, "comb_points_in_boxes_index: null idx");
  COMB_CUDA(cudaMemsetAsync(idx, 0xFF, (size_t)batch * P * 4, stream));
  if (T == 0) return COMB_OK;
  COMB_CHECK_ARG(points && boxes, "comb_points_in_boxes_index: null pointer");
  size_t smem = (size_t)T * sizeof(PibBox);
  COMB_CHECK_ARG(smem <= 200 * 1024, "comb_points_in_boxes_index: T=%d boxes per frame exceed shared memory", T);
  if (smem > 48 * 1024)
    COMB_CUDA(cudaFuncSetAttribute(pib_index_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(P, 256), batch);
  pib_index_kernel<<<grid, 256, smem, stream>>>(points, boxes, P, T, idx);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

extern "C" int comb_boxes_bev(const float* boxes_a, const float* trig_a, int na, const float* boxes_b,
                              const float* trig_b, int nb, int flavour, int what, float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(na >= 0 && nb >= 0, "comb_boxes_bev: bad shape");
  COMB_CHECK_ARG(flavour == 0 || flavour == 1, "comb_boxes_bev: flavour must be 0 (cpu) or 1 (gpu)");
  if (na == 0 || nb == 0) return COMB_OK;
  COMB_CHECK_ARG(boxes_a && boxes_b && out, "comb_boxes_bev: null pointer");
  COMB_CHECK_ARG(flavour == 1 || (trig_a && trig_b), "comb_boxes_bev: cpu flavour needs host-libm trig tables");
  const int vec_ok = (nb % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const bool big = (long long)na * nb >= (4ll << 20);        // enough 64-row tiles to fill the GPU several times
  const int ta = big ? 64 : 16;
  const int tiles_x = cdiv(nb, kTB);
  const long long ntiles = (long long)tiles_x * cdiv(na, ta);
  COMB_CHECK_ARG(ntiles < (1ll << 31), "comb_boxes_bev: matrix %d x %d too large", na, nb);
  auto launch = [&](auto kernel) -> int {
    int per_sm = 0;      // persistent blocks: exactly as many as are resident at once
    COMB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kRound, 0));
    const long long cap = (long long)sm_count() * (per_sm > 0 ? per_sm : 1);
    const int grid = (int)(ntiles < cap ? ntiles : cap);
    kernel<<<grid, kRound, 0, stream>>>(boxes_a, trig_a, na, boxes_b, trig_b, nb, what, out, vec_ok, tiles_x, (int)ntiles);
    return COMB_OK;
  };
  int rc;
  if (flavour == 0 && big) rc = launch(boxes_bev_kernel<true, 64>);
  else if (flavour == 0) rc = launch(boxes_bev_kernel<true, 16>);
  else if (big) rc = launch(boxes_bev_kernel<false, 64>);
  else rc = launch(boxes_bev_kernel<false, 16>);
  if (rc != COMB_OK) return rc;
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

extern "C" size_t comb_nms_workspace_bytes(int n) {
  if (n <= 0) return 256;
  size_t cb = (size_t)(n + 63) / 64;
  return align_up((size_t)n * cb * 8, 256);
}

extern "C" int comb_nms_dev(const float* boxes, const float* trig, int n, const int* n_dev, float thresh, int rotated,
                            int flavour, long long* keep, int* num_keep, void* workspace, size_t workspace_bytes,
                            void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(n >= 0 && keep && num_keep, "comb_nms: bad arguments");
  COMB_CHECK_ARG(flavour == 0 || flavour == 1, "comb_nms: flavour must be 0 (cpu) or 1 (gpu)");
  if (n == 0) {
    COMB_CUDA(cudaMemsetAsync(num_keep, 0, 4, stream));
    return COMB_OK;
  }
  COMB_CHECK_ARG(boxes && workspace && workspace_bytes >= comb_nms_workspace_bytes(n), "comb_nms: workspace too small");
  COMB_CHECK_ARG(!(rotated && flavour == 0) || trig, "comb_nms: cpu flavour needs host-libm trig table");
  const int cb = cdiv(n, 64);
  COMB_CHECK_ARG(cb <= 65535, "comb_nms: too many boxes (%d)", n);
  unsigned long long* mask = (unsigned long long*)workspace;
  dim3 grid(cb, cdiv(n, 4));       // 4 rows x 64 columns per block
  COMB_CHECK_ARG(grid.y <= 65535, "comb_nms: too many boxes (%d)", n);
  if (!rotated)
    nms_mask_kernel<false, false><<<grid, 256, 0, stream>>>(boxes, trig, n, n_dev, thresh, mask, cb);
  else if (flavour == 0)
    nms_mask_kernel<true, true><<<grid, 256, 0, stream>>>(boxes, trig, n, n_dev, thresh, mask, cb);
  else
    nms_mask_kernel<false, true><<<grid, 256, 0, stream>>>(boxes, trig, n, n_dev, thresh, mask, cb);
  COMB_LAUNCH_CHECK();
  const size_t staged = ((size_t)n * cb + cb) * 8;
  if (staged <= 200 * 1024) {
    static thread_local DevOnce configured;   // per device: the attribute is a per-device property
    if (configured.first()) {
      COMB_CUDA(cudaFuncSetAttribute(nms_sweep_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    nms_sweep_smem_kernel<<<1, 256, staged, stream>>>(mask, n, n_dev, cb, keep, num_keep);
  } else {
    nms_sweep_kernel<<<1, 32, (size_t)cb * 8, stream>>>(mask, n, n_dev, cb, keep, num_keep);
  }
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

extern "C" int comb_nms(const float* boxes, const float* trig, int n, float thresh, int rotated, int flavour,
                        long long* keep, int* num_keep, void* workspace, size_t workspace_bytes, void* stream_) {
  return comb_nms_dev(boxes, trig, n, nullptr, thresh, rotated, flavour, keep, num_keep, workspace, workspace_bytes, stream_);
}

// ---- f3: COMAug placement test -------------------------------------------------------------------------------------
// valid[i] = (max_j iou1[i][j] + max_j iou2[i][j], diagonal of iou2 zeroed) == 0 — the collision test of the COMAug
// database sampler (pcdet/datasets/augmentor/database_sampler_v2.py:600-604) — computed from the two IoU matrices
// where they lie on the device: one byte per sampled box goes back to the host instead of S x (E + S) floats.
// numpy's max propagates NaN (a degenerate box): NaN + x == 0 is false, the box is rejected as there.
namespace comb {
namespace {
__global__ void __launch_bounds__(128) comaug_valid_kernel(const float* __restrict__ iou1, const float* __restrict__ iou2,
                                                            int S, int E, unsigned char* __restrict__ valid) {
  const int i = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ float s_m1[4], s_m2[4];
  __shared__ int s_nan[4];
  float m1 = 0.0f, m2 = 0.0f;         // IoU >= 0, and a row of iou2 always holds its zeroed diagonal
  int nan = 0;
  for (int j = threadIdx.x; j < E; j += blockDim.x) {
    const float v = iou1[(size_t)i * E + j];
    nan |= (v != v);
    m1 = fmaxf(m1, v);
  }
  for (int j = threadIdx.x; j < S; j += blockDim.x) {
    const float v = j == i ? 0.0f : iou2[(size_t)i * S + j];
    nan |= (v != v);
    m2 = fmaxf(m2, v);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, d));
    m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, d));
    nan |= __shfl_xor_sync(0xffffffffu, nan, d);
  }
  if (lane == 0) { s_m1[warp] = m1; s_m2[warp] = m2; s_nan[warp] = nan; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 4; ++w) { m1 = fmaxf(m1, s_m1[w]); m2 = fmaxf(m2, s_m2[w]); nan |= s_nan[w]; }
    if (E == 0) m1 = m2;              // `iou1 = iou1 if iou1.shape[1] > 0 else iou2`
    valid[i] = (!nan && (m1 + m2) == 0.0f) ? 1 : 0;
  }
}
}  // namespace
}  // namespace comb

extern "C" int comb_comaug_valid_mask(const float* iou1, const float* iou2, int S, int E, unsigned char* valid,
                                      void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(S >= 0 && E >= 0, "comb_comaug_valid_mask: negative size");
  if (S == 0) return COMB_OK;
  COMB_CHECK_ARG((iou1 || E == 0) && iou2 && valid, "comb_comaug_valid_mask: null pointer");
  comb::comaug_valid_kernel<<<S, 128, 0, stream>>>(iou1, iou2, S, E, valid);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}
