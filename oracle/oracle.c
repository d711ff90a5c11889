/*
 * oracle.c — CPU restatement of the COM voxel-detector hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (com_b200) never calls it.
 *
 * What is restated and what pins it:
 *  - box ops (points_in_boxes_cpu, boxes_iou_bev_cpu, greedy NMS sweep): restatement of the
 *    reference's own C++ (pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:121-168,
 *    pcdet/ops/iou3d_nms/src/iou3d_cpu.cpp:32-252, iou3d_nms.cpp:117-133).  PINNED bit-exactly
 *    against the reference compiled from /root/reference (oracle/_ref, tests/test_oracle_ref.py)
 *    and against golden vectors generated from it (tests/golden/box_ops_ref.npz).
 *  - voxel generator, rulebook, sparse conv, dense(): the arithmetic lives in spconv (+cumm), a
 *    third-party dependency the reference does not vendor or pin ("spconv v1.0 (commit 8da6f96)
 *    or v1.2 or v2.x", docs/INSTALL.md:9; docker/Dockerfile:55 installs spconv-cu102 unpinned) and
 *    that is not installed here.  The reference has no tests or golden vectors for it.  The
 *    functions below restate spconv's published CPU semantics —
 *    PARITY UNPINNED for this segment.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared oracle.c -o liboracle.so -lm   (see Makefile)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ============================================================ voxel generator (spconv semantics)
 * Point2VoxelCPU3d.point_to_voxel / points_to_voxel_3d_np as called by
 * VoxelGeneratorWrapper.generate (pcdet/datasets/processor/data_processor.py:44-60):
 * single pass in point order; lut = dense coor->voxel id table of the grid volume (caller
 * provides it filled with -1; it is restored to -1 before returning). coords are (z,y,x). */
int orc_voxelize(const float* pts, int n, int C, const float* vsize, const float* range, int T, int cap,
                 float* voxels, int* coords, int* num, int* lut) {
  int grid[3];
  for (int j = 0; j < 3; ++j) grid[j] = (int)lrintf((range[3 + j] - range[j]) / vsize[j]);
  int nvox = 0;
  for (int i = 0; i < n; ++i) {
    int c[3], ok = 1;
    for (int j = 0; j < 3; ++j) {
      float q = floorf((pts[(size_t)i * C + j] - range[j]) / vsize[j]);
      if (!(q >= 0.0f && q < (float)grid[j])) { ok = 0; break; }
      c[j] = (int)q;
    }
    if (!ok) continue;
    size_t cell = ((size_t)c[2] * grid[1] + c[1]) * grid[0] + c[0];
    int vid = lut[cell];
    if (vid < 0) {
      if (nvox >= cap) continue;
      vid = nvox++;
      lut[cell] = vid;
      coords[vid * 3 + 0] = c[2]; coords[vid * 3 + 1] = c[1]; coords[vid * 3 + 2] = c[0];
      num[vid] = 0;
      memset(voxels + (size_t)vid * T * C, 0, sizeof(float) * T * C);
    }
    int m = num[vid];
    if (m < T) {
      memcpy(voxels + ((size_t)vid * T + m) * C, pts + (size_t)i * C, sizeof(float) * C);
      num[vid] = m + 1;
    }
  }
  for (int v = 0; v < nvox; ++v)
    lut[((size_t)coords[v * 3] * grid[1] + coords[v * 3 + 1]) * grid[0] + coords[v * 3 + 2]] = -1;
  return nvox;
}

/* MeanVFE.forward (pcdet/models/backbones_3d/vfe/mean_vfe.py:26-29) */
void orc_mean_vfe(const float* voxels, const int* num, int M, int T, int C, float* out) {
  for (int m = 0; m < M; ++m)
    for (int c = 0; c < C; ++c) {
      float s = 0.0f;
      for (int t = 0; t < T; ++t) s += voxels[((size_t)m * T + t) * C + c];
      float d = (float)(num[m] < 1 ? 1 : num[m]);
      out[(size_t)m * C + c] = s / d;
    }
}

/* ============================================================ rulebook (spconv semantics) */
typedef struct { int64_t key; int row; } KeyRow;
static int cmp_keyrow(const void* a, const void* b) {
  int64_t ka = ((const KeyRow*)a)->key, kb = ((const KeyRow*)b)->key;
  return ka < kb ? -1 : (ka > kb ? 1 : 0);
}
static int64_t lin_key(int b, int z, int y, int x, int D, int H, int W) { return (((int64_t)b * D + z) * H + y) * W + x; }
static int find_row(const KeyRow* t, int n, int64_t key) {
  int lo = 0, hi = n - 1;
  while (lo <= hi) {
    int mid = (lo + hi) >> 1;
    if (t[mid].key == key) return t[mid].row;
    if (t[mid].key < key) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

/* Output coordinate set of a strided SparseConv3d in canonical (ascending linear key) order.
 * o = (i + p - k*d) / s when divisible and inside out_shape.  Returns the count (<= cap). */
int orc_conv_out_coords(const int* in_coords, int n, const int* out_shape, const int* ks, const int* st,
                        const int* pd, const int* dl, int* out_coords, int cap) {
  int K = ks[0] * ks[1] * ks[2];
  int64_t* keys = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1) * K);
  size_t m = 0;
  for (int i = 0; i < n; ++i) {
    const int* c = in_coords + (size_t)i * 4;
    for (int kz = 0; kz < ks[0]; ++kz) for (int ky = 0; ky < ks[1]; ++ky) for (int kx = 0; kx < ks[2]; ++kx) {
      int t[3] = {c[1] + pd[0] - kz * dl[0], c[2] + pd[1] - ky * dl[1], c[3] + pd[2] - kx * dl[2]};
      int ok = 1, o[3];
      for (int j = 0; j < 3; ++j) {
        if (t[j] < 0 || t[j] % st[j]) { ok = 0; break; }
        o[j] = t[j] / st[j];
        if (o[j] >= out_shape[j]) { ok = 0; break; }
      }
      if (ok) keys[m++] = lin_key(c[0], o[0], o[1], o[2], out_shape[0], out_shape[1], out_shape[2]);
    }
  }
  /* sort + unique */
  KeyRow* kr = (KeyRow*)malloc(sizeof(KeyRow) * (m > 0 ? m : 1));
  for (size_t i = 0; i < m; ++i) { kr[i].key = keys[i]; kr[i].row = 0; }
  qsort(kr, m, sizeof(KeyRow), cmp_keyrow);
  int cnt = 0;
  for (size_t i = 0; i < m; ++i) {
    if (i > 0 && kr[i].key == kr[i - 1].key) continue;
    if (cnt < cap) {
      int64_t k = kr[i].key;
      int x = (int)(k % out_shape[2]); k /= out_shape[2];
      int y = (int)(k % out_shape[1]); k /= out_shape[1];
      int z = (int)(k % out_shape[0]); k /= out_shape[0];
      out_coords[cnt * 4 + 0] = (int)k; out_coords[cnt * 4 + 1] = z; out_coords[cnt * 4 + 2] = y; out_coords[cnt * 4 + 3] = x;
    }
    ++cnt;
  }
  free(kr); free(keys);
  return cnt < cap ? cnt : cap;
}

/* Gather-form rulebook: nbr[k*no + o] = row of input coordinate o*s - p + k*d, else -1.
 * SubM: out_coords == in_coords, s = 1, p = (k/2)*d. */
void orc_nbrmap(const int* out_coords, int no, const int* in_coords, int ni, const int* in_shape, const int* ks,
                const int* st, const int* pd, const int* dl, int* nbr) {
  KeyRow* t = (KeyRow*)malloc(sizeof(KeyRow) * (size_t)(ni > 0 ? ni : 1));
  for (int i = 0; i < ni; ++i) {
    const int* c = in_coords + (size_t)i * 4;
    t[i].key = lin_key(c[0], c[1], c[2], c[3], in_shape[0], in_shape[1], in_shape[2]);
    t[i].row = i;
  }
  qsort(t, ni, sizeof(KeyRow), cmp_keyrow);
  for (int o = 0; o < no; ++o) {
    const int* c = out_coords + (size_t)o * 4;
    int k = 0;
    for (int kz = 0; kz < ks[0]; ++kz) for (int ky = 0; ky < ks[1]; ++ky) for (int kx = 0; kx < ks[2]; ++kx, ++k) {
      int z = c[1] * st[0] - pd[0] + kz * dl[0], y = c[2] * st[1] - pd[1] + ky * dl[1], x = c[3] * st[2] - pd[2] + kx * dl[2];
      int row = -1;
      if (z >= 0 && z < in_shape[0] && y >= 0 && y < in_shape[1] && x >= 0 && x < in_shape[2])
        row = find_row(t, ni, lin_key(c[0], z, y, x, in_shape[0], in_shape[1], in_shape[2]));
      nbr[(size_t)k * no + o] = row;
    }
  }
  free(t);
}

/* ============================================================ sparse conv (fp64 accumulate = "truth")
 * out[o,co] = sum_k sum_ci in[nbr[k][o],ci] * W[co,k,ci] (+bias).  W layout (Cout,K,Cin). */
void orc_conv_fwd(const float* in, int Cin, const float* W, int K, int Cout, const int* nbr, int no, const float* bias,
                  float* out) {
  double* acc = (double*)malloc(sizeof(double) * Cout);
  for (int o = 0; o < no; ++o) {
    for (int co = 0; co < Cout; ++co) acc[co] = bias ? (double)bias[co] : 0.0;
    for (int k = 0; k < K; ++k) {
      int i = nbr[(size_t)k * no + o];
      if (i < 0) continue;
      const float* a = in + (size_t)i * Cin;
      for (int co = 0; co < Cout; ++co) {
        const float* w = W + ((size_t)co * K + k) * Cin;
        double s = 0.0;
        for (int ci = 0; ci < Cin; ++ci) s += (double)a[ci] * (double)w[ci];
        acc[co] += s;
      }
    }
    for (int co = 0; co < Cout; ++co) out[(size_t)o * Cout + co] = (float)acc[co];
  }
  free(acc);
}

/* din[i,ci] = sum_{k,o: nbr[k][o]==i} sum_co dout[o,co] * W[co,k,ci] */
void orc_conv_dgrad(const float* dout, int Cout, const float* W, int K, int Cin, const int* nbr, int no, int ni,
                    float* din) {
  double* acc = (double*)calloc((size_t)(ni > 0 ? ni : 1) * Cin, sizeof(double));
  for (int k = 0; k < K; ++k)
    for (int o = 0; o < no; ++o) {
      int i = nbr[(size_t)k * no + o];
      if (i < 0) continue;
      for (int co = 0; co < Cout; ++co) {
        double g = dout[(size_t)o * Cout + co];
        const float* w = W + ((size_t)co * K + k) * Cin;
        for (int ci = 0; ci < Cin; ++ci) acc[(size_t)i * Cin + ci] += g * (double)w[ci];
      }
    }
  for (size_t e = 0; e < (size_t)ni * Cin; ++e) din[e] = (float)acc[e];
  free(acc);
}

/* dW[co,k,ci] = sum_o dout[o,co] * in[nbr[k][o],ci] */
void orc_conv_wgrad(const float* in, int Cin, const float* dout, int Cout, int K, const int* nbr, int no, float* dW) {
  double* acc = (double*)calloc((size_t)Cout * K * Cin, sizeof(double));
  for (int k = 0; k < K; ++k)
    for (int o = 0; o < no; ++o) {
      int i = nbr[(size_t)k * no + o];
      if (i < 0) continue;
      for (int co = 0; co < Cout; ++co) {
        double g = dout[(size_t)o * Cout + co];
        for (int ci = 0; ci < Cin; ++ci) acc[((size_t)co * K + k) * Cin + ci] += g * (double)in[(size_t)i * Cin + ci];
      }
    }
  for (size_t e = 0; e < (size_t)Cout * K * Cin; ++e) dW[e] = (float)acc[e];
  free(acc);
}

/* SparseConvTensor.dense() -> (B,C,D,H,W) (pcdet/models/backbones_2d/map_to_bev/height_compression.py:21) */
void orc_dense(const float* feats, const int* coords, int n, int B, int C, int D, int H, int W, float* out) {
  memset(out, 0, sizeof(float) * (size_t)B * C * D * H * W);
  for (int i = 0; i < n; ++i) {
    const int* c = coords + (size_t)i * 4;
    for (int ch = 0; ch < C; ++ch)
      out[((((size_t)c[0] * C + ch) * D + c[1]) * H + c[2]) * W + c[3]] = feats[(size_t)i * C + ch];
  }
}

/* ============================================================ box ops (reference CPU arithmetic)
 * points_in_boxes_cpu — roiaware_pool3d.cpp:121-168: fp32 rotate with cosf/sinf(-rz), fp64 compares,
 * MARGIN = (float)1e-2. */
static int pt_in_box_cpu(const float* pt, const float* box) {
  const float MARGIN = 1e-2f;
  float x = pt[0], y = pt[1], z = pt[2];
  float cx = box[0], cy = box[1], cz = box[2], dx = box[3], dy = box[4], dz = box[5], rz = box[6];
  if ((double)fabsf(z - cz) > (double)dz / 2.0) return 0;
  float sx = x - cx, sy = y - cy;
  float cosa = cosf(-rz), sina = sinf(-rz);
  float lx = sx * cosa + sy * (-sina);
  float ly = sx * sina + sy * cosa;
  return ((double)fabsf(lx) < (double)dx / 2.0 + (double)MARGIN) & ((double)fabsf(ly) < (double)dy / 2.0 + (double)MARGIN);
}

void orc_points_in_boxes_cpu(const float* boxes, int nb, const float* pts, int P, int* out) {
  for (int b = 0; b < nb; ++b)
    for (int p = 0; p < P; ++p) out[(size_t)b * P + p] = pt_in_box_cpu(pts + (size_t)p * 3, boxes + (size_t)b * 7);
}

/* rotated BEV overlap — iou3d_cpu.cpp:32-229 */
typedef struct { float x, y; } Pt;
static float fmin2(float a, float b) { return a > b ? b : a; }
static float fmax2(float a, float b) { return a > b ? a : b; }
static float cr3(Pt p1, Pt p2, Pt p0) { return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y); }

static int seg_x(Pt p1, Pt p0, Pt q1, Pt q0, Pt* ans) {
  if (!(fmin2(p0.x, p1.x) <= fmax2(q0.x, q1.x) && fmin2(q0.x, q1.x) <= fmax2(p0.x, p1.x) &&
        fmin2(p0.y, p1.y) <= fmax2(q0.y, q1.y) && fmin2(q0.y, q1.y) <= fmax2(p0.y, p1.y)))
    return 0;
  float s1 = cr3(q0, p1, p0), s2 = cr3(p1, q1, p0), s3 = cr3(p0, q1, q0), s4 = cr3(q1, p1, q0);
  if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
  float s5 = cr3(q1, p1, p0);
  const float EPS = 1e-8f;
  if (fabsf(s5 - s1) > EPS) {
    ans->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
    ans->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
  } else {
    float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
    float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
    float D = a0 * b1 - a1 * b0;
    ans->x = (b0 * c1 - b1 * c0) / D;
    ans->y = (a1 * c0 - a0 * c1) / D;
  }
  return 1;
}

static int in_box2d(const float* box, Pt p) {
  const float MARGIN = 1e-2f;
  float ac = cosf(-box[6]), as = sinf(-box[6]);
  float rx = (p.x - box[0]) * ac + (p.y - box[1]) * (-as);
  float ry = (p.x - box[0]) * as + (p.y - box[1]) * ac;
  return fabsf(rx) < box[3] / 2 + MARGIN && fabsf(ry) < box[4] / 2 + MARGIN;
}

static void corners_of(const float* b, Pt* c) {
  float hx = b[3] / 2, hy = b[4] / 2;
  float x1 = b[0] - hx, y1 = b[1] - hy, x2 = b[0] + hx, y2 = b[1] + hy;
  float cs = cosf(b[6]), sn = sinf(b[6]);
  float px[4] = {x1, x2, x2, x1}, py[4] = {y1, y1, y2, y2};
  for (int k = 0; k < 4; ++k) {
    float nx = (px[k] - b[0]) * cs + (py[k] - b[1]) * (-sn) + b[0];
    float ny = (px[k] - b[0]) * sn + (py[k] - b[1]) * cs + b[1];
    c[k].x = nx; c[k].y = ny;
  }
  c[4] = c[0];
}

static float overlap_cpu(const float* a, const float* b) {
  Pt ca[5], cb[5], pts[16], ctr = {0.0f, 0.0f};
  corners_of(a, ca);
  corners_of(b, cb);
  int cnt = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      if (seg_x(ca[i + 1], ca[i], cb[j + 1], cb[j], &pts[cnt])) {
        ctr.x = ctr.x + pts[cnt].x; ctr.y = ctr.y + pts[cnt].y; ++cnt;
      }
  for (int k = 0; k < 4; ++k) {
    if (in_box2d(a, cb[k])) { ctr.x = ctr.x + cb[k].x; ctr.y = ctr.y + cb[k].y; pts[cnt++] = cb[k]; }
    if (in_box2d(b, ca[k])) { ctr.x = ctr.x + ca[k].x; ctr.y = ctr.y + ca[k].y; pts[cnt++] = ca[k]; }
  }
  ctr.x /= cnt; ctr.y /= cnt;
  for (int j = 0; j < cnt - 1; ++j)
    for (int i = 0; i < cnt - j - 1; ++i)
      if (atan2f(pts[i].y - ctr.y, pts[i].x - ctr.x) > atan2f(pts[i + 1].y - ctr.y, pts[i + 1].x - ctr.x)) {
        Pt t = pts[i]; pts[i] = pts[i + 1]; pts[i + 1] = t;
      }
  float area = 0.0f;
  for (int k = 0; k < cnt - 1; ++k) {
    Pt u = {pts[k].x - pts[0].x, pts[k].y - pts[0].y}, v = {pts[k + 1].x - pts[0].x, pts[k + 1].y - pts[0].y};
    area += u.x * v.y - u.y * v.x;
  }
  return (float)(fabsf(area) / 2.0);
}

static float iou_cpu(const float* a, const float* b) {
  float sa = a[3] * a[4], sb = b[3] * b[4], so = overlap_cpu(a, b);
  return so / fmaxf(sa + sb - so, 1e-8f);
}

/* boxes_iou_bev_cpu — iou3d_cpu.cpp:232-252; what: 0 = IoU, 1 = overlap area */
void orc_boxes_bev_cpu(const float* a, int na, const float* b, int nb, int what, float* out) {
  for (int i = 0; i < na; ++i)
    for (int j = 0; j < nb; ++j)
      out[(size_t)i * nb + j] = what ? overlap_cpu(a + (size_t)i * 7, b + (size_t)j * 7) : iou_cpu(a + (size_t)i * 7, b + (size_t)j * 7);
}

/* Greedy NMS over score-sorted boxes: suppression bit matrix (iou > thresh, j > i) + the host sweep
 * of iou3d_nms.cpp:117-133, with the CPU IoU above.  Returns the number kept. */
int orc_nms_cpu(const float* boxes, int n, float thresh, long long* keep) {
  int cb = (n + 63) / 64;
  unsigned long long* remv = (unsigned long long*)calloc(cb > 0 ? cb : 1, sizeof(unsigned long long));
  int kept = 0;
  for (int i = 0; i < n; ++i) {
    if (remv[i >> 6] & (1ull << (i & 63))) continue;
    keep[kept++] = i;
    for (int j = i + 1; j < n; ++j)
      if (iou_cpu(boxes + (size_t)i * 7, boxes + (size_t)j * 7) > thresh) remv[j >> 6] |= 1ull << (j & 63);
  }
  free(remv);
  return kept;
}
