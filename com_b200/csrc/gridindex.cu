// a6 — bitmap-rank grid index: spatially ORDERED row sets and their rulebooks without hashing.
//
// Replaces the indice-pair generation inside spconv's SubMConv3d / SparseConv3d forward (call sites
// pcdet/models/backbones_3d/spconv_backbone.py:191-232) for the fused pipeline, where every level keeps its
// rows in canonical order (ascending linear key ((b*D+z)*H+y)*W+x, SURVEY hard part 1).
//
// An index of one level is
//   bitmap  : 1 bit per grid cell, 32-bit words, padded to whole 8-word (32-byte = one DRAM sector) blocks
//   bprefix : number of set bits before each 8-word block
// so that  row(key) = bprefix[key>>8] + popc(words of the block before key>>5) + popc(word & below-mask)
// costs ONE sector of bitmap and one int.  Because rows are in key order, neighbouring voxels sit in
// neighbouring words and neighbouring rows: the 27 probes of an output row touch ~9 sectors that its
// neighbours in the warp touch as well (the 8-byte-slot hash table it replaces costs 1.5 random L2 lines
// per probe; r1 profile: nbrmap 34% and hash/outset 17% of the step).
//
//   comb_index_build : memset -> mark (the level's own coords, or the output set of a strided conv of the
//                      previous level) -> block popcounts + local scan -> scan of super-block sums ->
//                      emit coords in key order + finalise bprefix.
//   comb_index_rank  : row of every coordinate (-1 when absent) — the permutation voxel order -> key order.
//   comb_nbrmap_build_indexed : gather-form rulebook nbr[k][o] by bitmap test + rank.
#include "common.cuh"

namespace comb {
namespace {

struct Conv3 {
  int k[3], s[3], p[3], d[3];  // z, y, x
};

constexpr int kBlkWords = 8;              // words per rank block (one 32-byte sector)
constexpr int kSuperBlks = 1024;          // rank blocks per super-block (one CUDA block of 256 threads)
constexpr int kSuperWords = kBlkWords * kSuperBlks;

struct IndexDims {
  long long vol;      // cells
  int nwords;         // padded to kBlkWords
  int nblks;          // rank blocks
  int nsuper;         // super-blocks
};

static IndexDims index_dims(int batch, int D, int H, int W) {
  IndexDims d;
  d.vol = (long long)batch * D * H * W;
  long long w = (d.vol + 31) / 32;
  w = (w + kBlkWords - 1) / kBlkWords * kBlkWords;
  d.nwords = (int)w;
  d.nblks = d.nwords / kBlkWords;
  d.nsuper = (d.nblks + kSuperBlks - 1) / kSuperBlks;
  return d;
}

__device__ __forceinline__ void set_bit(uint32_t* __restrict__ bitmap, uint32_t key) {
  const uint32_t bit = 1u << (key & 31);
  uint32_t* w = bitmap + (key >> 5);
  if (!(*((volatile uint32_t*)w) & bit)) atomicOr(w, bit);
}

// ---- mark ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) index_mark_self_kernel(const int4* __restrict__ coords, int n_max,
                                                               const int* __restrict__ n_dev, int batch, int D, int H,
                                                               int W, uint32_t* __restrict__ bitmap) {
  const int n = eff_n(n_max, n_dev);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 c = __ldg(coords + i);  // (b, z, y, x)
  if ((unsigned)c.x >= (unsigned)batch || (unsigned)c.y >= (unsigned)D || (unsigned)c.z >= (unsigned)H ||
      (unsigned)c.w >= (unsigned)W)
    return;
  set_bit(bitmap, (uint32_t)(((c.x * D + c.y) * H + c.z) * W + c.w));
}

// Output set of a strided conv: input voxel i contributes to o = (i + p - k*d)/s for every kernel tap k that
// divides; only the taps that can divide are visited (<= ceil(k/s) per axis, 8 in total for k=3, s=2).
__global__ void __launch_bounds__(256) index_mark_conv_kernel(const int4* __restrict__ coords, int n_max,
                                                               const int* __restrict__ n_dev, Conv3 cv, int oD, int oH,
                                                               int oW, uint32_t* __restrict__ bitmap) {
  const int n = eff_n(n_max, n_dev);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 c = __ldg(coords + i);
  const int in[3] = {c.y, c.z, c.w};
  const int on[3] = {oD, oH, oW};
  int cand[3][4], nc[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    nc[a] = 0;
    for (int k = 0; k < cv.k[a]; ++k) {
      const int t = in[a] + cv.p[a] - k * cv.d[a];
      if (t < 0 || (t % cv.s[a]) != 0) continue;
      const int o = t / cv.s[a];
      if (o >= on[a]) continue;
      bool dup = false;  // with dilation two taps can land on the same output
      for (int q = 0; q < nc[a]; ++q) dup |= (cand[a][q] == o);
      if (!dup && nc[a] < 4) cand[a][nc[a]++] = o;
    }
  }
  for (int a = 0; a < nc[0]; ++a)
    for (int b = 0; b < nc[1]; ++b)
      for (int e = 0; e < nc[2]; ++e)
        set_bit(bitmap, (uint32_t)(((c.x * oD + cand[0][a]) * oH + cand[1][b]) * oW + cand[2][e]));
}

// ---- scan ------------------------------------------------------------------------------------
// One CUDA block per super-block: popcount of each 8-word rank block, exclusive scan inside the super-block.
__global__ void __launch_bounds__(256) index_count_kernel(const uint32_t* __restrict__ bitmap, int nblks,
                                                           int* __restrict__ bprefix, int* __restrict__ super_sums) {
  __shared__ int cnt[kSuperBlks];
  __shared__ int wsum[8];
  const int sb = blockIdx.x, t = threadIdx.x;
  const int b0 = sb * kSuperBlks;
#pragma unroll
  for (int j = 0; j < kSuperBlks / 256; ++j) {
    const int b = b0 + j * 256 + t;
    int c = 0;
    if (b < nblks) {
      const uint4* p = reinterpret_cast<const uint4*>(bitmap + (size_t)b * kBlkWords);
      const uint4 u = __ldg(p), v = __ldg(p + 1);
      c = __popc(u.x) + __popc(u.y) + __popc(u.z) + __popc(u.w) + __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
    }
    cnt[j * 256 + t] = c;
  }
  __syncthreads();
  // thread t owns rank blocks 4t..4t+3
  const int c0 = cnt[4 * t], c1 = cnt[4 * t + 1], c2 = cnt[4 * t + 2], c3 = cnt[4 * t + 3];
  const int mine = c0 + c1 + c2 + c3;
  int incl = mine;
  const int lane = t & 31, warp = t >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  int woff = 0, total = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const int s = wsum[w];
    if (w < warp) woff += s;
    total += s;
  }
  int ex = woff + incl - mine;
  if (total == 0) {
    if (t == 0) super_sums[sb] = 0;
    return;  // bprefix of an empty super-block is finalised by the emit kernel (uniform value)
  }
  const int b = b0 + 4 * t;
  if (b + 3 < nblks) {
    reinterpret_cast<int4*>(bprefix)[b >> 2] = make_int4(ex, ex + c0, ex + c0 + c1, ex + c0 + c1 + c2);
  } else {
    if (b < nblks) bprefix[b] = ex;
    if (b + 1 < nblks) bprefix[b + 1] = ex + c0;
    if (b + 2 < nblks) bprefix[b + 2] = ex + c0 + c1;
  }
  if (t == 0) super_sums[sb] = total;
}

// Single block: exclusive scan of the super-block sums (in place, n+1 entries), total -> out_count.
__global__ void __launch_bounds__(1024) index_scan_kernel(int* __restrict__ super_sums, int nsuper, int out_cap,
                                                           int* __restrict__ out_count) {
  __shared__ int warp_tot[32];
  __shared__ int s_running;
  if (threadIdx.x == 0) s_running = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b0 = 0; b0 < nsuper; b0 += 1024) {
    const int b = b0 + threadIdx.x;
    const int v = b < nsuper ? super_sums[b] : 0;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int w = warp_tot[lane];
      int wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, wi, d);
        if (lane >= d) wi += o;
      }
      warp_tot[lane] = wi - w;
    }
    __syncthreads();
    const int excl = s_running + warp_tot[warp] + incl - v;
    if (b < nsuper) super_sums[b] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_running = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int total = s_running;
    super_sums[nsuper] = total;
    *out_count = total < out_cap ? total : out_cap;
  }
}

// One thread per bitmap word: position of the word's first set bit = super-block offset + local prefix of its
// rank block + bits of the preceding words of the block; the thread of a block's first word also publishes the
// final (global) block prefix that comb_index_rank / the rulebook kernel use.  Coalesced word loads, almost all
// threads leave at once (the bitmaps are sparse), set bits are enumerated with one division per word.
__global__ void __launch_bounds__(256) index_emit_kernel(const uint32_t* __restrict__ bitmap, int nwords,
                                                          const int* __restrict__ bprefix_local,
                                                          int* __restrict__ bprefix, const int* __restrict__ super_off,
                                                          int oD, int oH, int oW, int out_cap,
                                                          int4* __restrict__ out_coords) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nwords) return;
  const int b = w >> 3, wi = w & 7;
  const int sb = b / kSuperBlks;
  const int off = __ldg(super_off + sb);
  const bool sb_empty = (__ldg(super_off + sb + 1) == off);   // bprefix_local of an empty super-block is not written
  int pos = off + (sb_empty ? 0 : __ldg(bprefix_local + b));
  if (wi == 0) bprefix[b] = pos;
  if (sb_empty || out_coords == nullptr) return;
  uint32_t bits = __ldg(bitmap + w);
  if (bits == 0u) return;
  for (int q = 0; q < wi; ++q) pos += __popc(__ldg(bitmap + (size_t)b * kBlkWords + q));
  // decode the word's first cell once, then walk along x (a word spans at most 32 cells)
  const uint32_t key0 = (uint32_t)w << 5;
  int x = (int)(key0 % (uint32_t)oW);
  uint32_t r = key0 / (uint32_t)oW;
  int y = (int)(r % (uint32_t)oH);
  r /= (uint32_t)oH;
  int z = (int)(r % (uint32_t)oD);
  int bb = (int)(r / (uint32_t)oD);
  int prev = 0;
  while (bits) {
    const int bp = __ffs(bits) - 1;
    bits &= bits - 1;
    x += bp - prev;
    prev = bp;
    while (x >= oW) {
      x -= oW;
      if (++y == oH) {
        y = 0;
        if (++z == oD) { z = 0; ++bb; }
      }
    }
    if (pos < out_cap) out_coords[pos] = make_int4(bb, z, y, x);
    ++pos;
  }
}

// Prefix finalisation alone (no coordinates wanted): one thread per rank block.  Used for the level-1 index, whose
// rows in key order are produced by comb_index_rank_scatter from the voxel list instead of by enumerating a 46 MB
// bitmap (r1 ncu: the enumerating emit kernel took 51 us on that bitmap, the rank kernel 9 us).
__global__ void __launch_bounds__(256) index_finalize_kernel(int nblks, const int* __restrict__ bprefix_local,
                                                              int* __restrict__ bprefix,
                                                              const int* __restrict__ super_off) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblks) return;
  const int sb = b / kSuperBlks;
  const int off = __ldg(super_off + sb);
  const bool sb_empty = (__ldg(super_off + sb + 1) == off);
  bprefix[b] = off + (sb_empty ? 0 : __ldg(bprefix_local + b));
}

// row of `key` (its bit is known to be set, or use only if you tested it)
__device__ __forceinline__ int index_rank(const uint32_t* __restrict__ bitmap, const int* __restrict__ bprefix,
                                          uint32_t key) {
  const uint32_t w = key >> 5, b = w >> 3, wi = w & 7;
  const uint4* p = reinterpret_cast<const uint4*>(bitmap + (size_t)b * kBlkWords);
  const uint4 u = __ldg(p);
  uint32_t words[8] = {u.x, u.y, u.z, u.w, 0, 0, 0, 0};
  if (wi >= 4) {
    const uint4 v = __ldg(p + 1);
    words[4] = v.x; words[5] = v.y; words[6] = v.z; words[7] = v.w;
  }
  int r = __ldg(bprefix + b);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    if (q < (int)wi) r += __popc(words[q]);
    else if (q == (int)wi) r += __popc(words[q] & ((1u << (key & 31)) - 1u));
  }
  return r;
}

__global__ void __launch_bounds__(256) index_rank_kernel(const int4* __restrict__ coords, int n_max,
                                                          const int* __restrict__ n_dev, int batch, int D, int H, int W,
                                                          const uint32_t* __restrict__ bitmap,
                                                          const int* __restrict__ bprefix, int* __restrict__ rows,
                                                          int4* __restrict__ sorted_coords, int sorted_cap) {
  const int n = eff_n(n_max, n_dev);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 c = __ldg(coords + i);
  int r = -1;
  if ((unsigned)c.x < (unsigned)batch && (unsigned)c.y < (unsigned)D && (unsigned)c.z < (unsigned)H &&
      (unsigned)c.w < (unsigned)W) {
    const uint32_t key = (uint32_t)(((c.x * D + c.y) * H + c.z) * W + c.w);
    if (__ldg(bitmap + (key >> 5)) & (1u << (key & 31))) r = index_rank(bitmap, bprefix, key);
  }
  rows[i] = r;
  // unique coordinates land on distinct rows: the coordinate list in key order, without enumerating the bitmap
  if (sorted_coords != nullptr && r >= 0 && r < sorted_cap) sorted_coords[r] = c;
}

// nbr[k][o] = row of coordinate o*s - p + k*d in the input level, by bitmap test + rank.  One thread per output
// row.  For a fixed (kz, ky) the kx taps are bits of ONE 32-bit window of the bitmap that starts at the first
// in-range tap (anchor): the thread first issues the loads of ALL its (kz, ky) lines — the anchor's 32-byte rank
// block, its prefix and (only when the anchor sits in the last word of a block) the first word of the next block —
// i.e. up to 27 independent loads whose sectors are shared with the neighbouring rows of the warp, then resolves
// every tap with popcounts:  row(tap) = prefix + popc(words before the anchor word) + popc(anchor word below the
// anchor bit) + popc(window below the tap bit).
// (r1: the first version walked the 27 taps one dependent load chain at a time — 30 us per level, 8 % of the HBM
// roofline, latency-bound; the second ranked every tap separately against its block — ALU-bound, ~1100
// instructions per row.)
template <int KZ, int KY, int KX>
__global__ void __launch_bounds__(256) nbrmap_indexed_kernel(const int4* __restrict__ out_coords, int no_max,
                                                              const int* __restrict__ no_dev,
                                                              const uint32_t* __restrict__ bitmap,
                                                              const int* __restrict__ bprefix, int nwords, int iD,
                                                              int iH, int iW, Conv3 cv, int* __restrict__ nbr, int ld) {
  const int no = eff_n(no_max, no_dev);
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= no) return;
  const int4 c = __ldg(out_coords + o);
  const int bz = c.y * cv.s[0] - cv.p[0], by = c.z * cv.s[1] - cv.p[1], bx = c.w * cv.s[2] - cv.p[2];
  const int xa = bx < 0 ? 0 : bx;                      // anchor: every in-range tap is at an offset in [0, 32) from it
  constexpr int NL = KZ * KY;
  uint4 lo[NL], hi[NL];
  uint32_t extra[NL];
  int prefix[NL];
  uint32_t akey[NL];
  bool ok[NL];
#pragma unroll
  for (int l = 0; l < NL; ++l) {
    const int z = bz + (l / KY) * cv.d[0], y = by + (l % KY) * cv.d[1];
    ok[l] = (unsigned)z < (unsigned)iD && (unsigned)y < (unsigned)iH && xa < iW;
    akey[l] = (uint32_t)(((c.x * iD + z) * iH + y) * iW + xa);
    lo[l] = hi[l] = make_uint4(0u, 0u, 0u, 0u);
    extra[l] = 0u;
    prefix[l] = 0;
    if (ok[l]) {
      const uint32_t w0 = akey[l] >> 5, b = w0 >> 3;
      const uint4* p = reinterpret_cast<const uint4*>(bitmap + (size_t)b * kBlkWords);
      lo[l] = __ldg(p);
      hi[l] = __ldg(p + 1);
      prefix[l] = __ldg(bprefix + b);
      if ((w0 & 7u) == 7u && (int)(w0 + 1) < nwords) extra[l] = __ldg(bitmap + w0 + 1);
    }
  }
#pragma unroll
  for (int l = 0; l < NL; ++l) {
    const uint32_t wi = (akey[l] >> 5) & 7u, sh = akey[l] & 31u;
    const uint32_t w[9] = {lo[l].x, lo[l].y, lo[l].z, lo[l].w, hi[l].x, hi[l].y, hi[l].z, hi[l].w, extra[l]};
    uint32_t w_lo = 0u, w_hi = 0u;
    int base = prefix[l];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      base += q < (int)wi ? __popc(w[q]) : 0;
      w_lo = q == (int)wi ? w[q] : w_lo;
      w_hi = q == (int)wi ? w[q + 1] : w_hi;
    }
    base += __popc(w_lo & ((1u << sh) - 1u));
    const uint32_t win = __funnelshift_r(w_lo, w_hi, sh);   // bit i = cell (anchor + i)
#pragma unroll
    for (int kx = 0; kx < KX; ++kx) {
      const int x = bx + kx * cv.d[2];
      const uint32_t off = (uint32_t)(x - xa);
      int row = -1;
      if (ok[l] && (unsigned)x < (unsigned)iW && ((win >> off) & 1u)) row = base + __popc(win & ((1u << off) - 1u));
      nbr[(size_t)(l * KX + kx) * ld + o] = row;
    }
  }
}

static int fill_conv3(Conv3& cv, const int* ksize, const int* stride, const int* pad, const int* dil) {
  for (int j = 0; j < 3; ++j) {
    cv.k[j] = ksize[j];
    cv.s[j] = stride ? stride[j] : 1;
    cv.p[j] = pad ? pad[j] : 0;
    cv.d[j] = dil ? dil[j] : 1;
    if (cv.k[j] < 1 || cv.s[j] < 1 || cv.p[j] < 0 || cv.d[j] < 1) return -1;
  }
  return 0;
}

static int check_volume(const char* who, int batch, int D, int H, int W) {
  if (batch < 1 || D < 1 || H < 1 || W < 1) {
    set_error("%s: bad grid %d x (%d,%d,%d)", who, batch, D, H, W);
    return COMB_EINVAL;
  }
  const unsigned long long vol = (unsigned long long)batch * D * H * W;
  if (vol >= 0x7FFFFF00ull) {
    set_error("%s: batch*volume %llu exceeds the 31-bit key space of the bitmap index", who, vol);
    return COMB_ERANGE;
  }
  return COMB_OK;
}

}  // namespace
}  // namespace comb

using namespace comb;

extern "C" size_t comb_index_bitmap_bytes(int batch, int D, int H, int W) {
  if (batch < 1 || D < 1 || H < 1 || W < 1) return 0;
  return align_up((size_t)index_dims(batch, D, H, W).nwords * 4, 256);
}

extern "C" size_t comb_index_prefix_bytes(int batch, int D, int H, int W) {
  if (batch < 1 || D < 1 || H < 1 || W < 1) return 0;
  const IndexDims d = index_dims(batch, D, H, W);
  // bprefix (padded to 4 ints), the super-block sums (n+1), and the per-super-block local prefixes (scratch)
  return 2 * align_up((size_t)(d.nblks + 4) * 4, 256) + align_up((size_t)(d.nsuper + 1) * 4, 256);
}

extern "C" int comb_index_build(const int* coords, int n_max, const int* n_dev, int batch, int D, int H, int W,
                                const int* ksize, const int* stride, const int* pad, const int* dil, void* bitmap,
                                void* prefix, int* out_coords, int out_cap, int* out_count, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(bitmap && prefix && out_count && n_max >= 0 && out_cap >= 0, "comb_index_build: bad arguments");
  int rc = check_volume("comb_index_build", batch, D, H, W);
  if (rc) return rc;
  const IndexDims d = index_dims(batch, D, H, W);
  uint32_t* bm = (uint32_t*)bitmap;
  int* bprefix = (int*)prefix;
  int* super = (int*)((char*)prefix + align_up((size_t)(d.nblks + 4) * 4, 256));
  int* blocal = (int*)((char*)super + align_up((size_t)(d.nsuper + 1) * 4, 256));
  COMB_CUDA(cudaMemsetAsync(bm, 0, (size_t)d.nwords * 4, stream));
  if (n_max > 0) {
    COMB_CHECK_ARG(coords, "comb_index_build: null coords");
    if (ksize) {
      Conv3 cv;
      COMB_CHECK_ARG(fill_conv3(cv, ksize, stride, pad, dil) == 0, "comb_index_build: bad conv parameters");
      COMB_CHECK_ARG(cv.k[0] <= 4 * cv.s[0] && cv.k[1] <= 4 * cv.s[1] && cv.k[2] <= 4 * cv.s[2],
                     "comb_index_build: kernel larger than 4 strides is not supported");
      index_mark_conv_kernel<<<cdiv(n_max, 256), 256, 0, stream>>>((const int4*)coords, n_max, n_dev, cv, D, H, W, bm);
    } else {
      index_mark_self_kernel<<<cdiv(n_max, 256), 256, 0, stream>>>((const int4*)coords, n_max, n_dev, batch, D, H, W,
                                                                   bm);
    }
    COMB_LAUNCH_CHECK();
  }
  index_count_kernel<<<d.nsuper, 256, 0, stream>>>(bm, d.nblks, blocal, super);
  COMB_LAUNCH_CHECK();
  index_scan_kernel<<<1, 1024, 0, stream>>>(super, d.nsuper, out_cap, out_count);
  COMB_LAUNCH_CHECK();
  if (out_coords == nullptr)
    index_finalize_kernel<<<cdiv(d.nblks, 256), 256, 0, stream>>>(d.nblks, blocal, bprefix, super);
  else
    index_emit_kernel<<<cdiv(d.nwords, 256), 256, 0, stream>>>(bm, d.nwords, blocal, bprefix, super, D, H, W, out_cap,
                                                               (int4*)out_coords);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

extern "C" int comb_index_rank(const int* coords, int n_max, const int* n_dev, int batch, int D, int H, int W,
                               const void* bitmap, const void* prefix, int* rows, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(n_max >= 0, "comb_index_rank: bad arguments");
  int rc = check_volume("comb_index_rank", batch, D, H, W);
  if (rc) return rc;
  if (n_max == 0) return COMB_OK;
  COMB_CHECK_ARG(coords && bitmap && prefix && rows, "comb_index_rank: null pointer");
  index_rank_kernel<<<cdiv(n_max, 256), 256, 0, stream>>>((const int4*)coords, n_max, n_dev, batch, D, H, W,
                                                          (const uint32_t*)bitmap, (const int*)prefix, rows, nullptr, 0);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

extern "C" int comb_index_rank_scatter(const int* coords, int n_max, const int* n_dev, int batch, int D, int H, int W,
                                       const void* bitmap, const void* prefix, int* rows, int* sorted_coords,
                                       int sorted_cap, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(n_max >= 0 && sorted_cap >= 0, "comb_index_rank_scatter: bad arguments");
  int rc = check_volume("comb_index_rank_scatter", batch, D, H, W);
  if (rc) return rc;
  if (n_max == 0) return COMB_OK;
  COMB_CHECK_ARG(coords && bitmap && prefix && rows && sorted_coords, "comb_index_rank_scatter: null pointer");
  index_rank_kernel<<<cdiv(n_max, 256), 256, 0, stream>>>((const int4*)coords, n_max, n_dev, batch, D, H, W,
                                                          (const uint32_t*)bitmap, (const int*)prefix, rows,
                                                          (int4*)sorted_coords, sorted_cap);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

extern "C" int comb_nbrmap_build_indexed(const int* out_coords, int no_max, const int* no_dev, const void* bitmap,
                                         const void* prefix, int batch, int iD, int iH, int iW, const int* ksize,
                                         const int* stride, const int* pad, const int* dil, int* nbr, int ld,
                                         void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(ksize && ld >= no_max && no_max >= 0, "comb_nbrmap_build_indexed: bad arguments");
  Conv3 cv;
  COMB_CHECK_ARG(fill_conv3(cv, ksize, stride, pad, dil) == 0, "comb_nbrmap_build_indexed: bad conv parameters");
  int rc = check_volume("comb_nbrmap_build_indexed", batch, iD, iH, iW);
  if (rc) return rc;
  if (no_max == 0) return COMB_OK;
  COMB_CHECK_ARG(out_coords && bitmap && prefix && nbr, "comb_nbrmap_build_indexed: null pointer");
  COMB_CHECK_ARG((cv.k[2] - 1) * cv.d[2] < 32, "comb_nbrmap_build_indexed: x extent of the kernel must be < 32 cells");
#define COMB_NBRMAP(KZ, KY, KX)                                                                                   \
  nbrmap_indexed_kernel<KZ, KY, KX><<<cdiv(no_max, 256), 256, 0, stream>>>(                                         \
      (const int4*)out_coords, no_max, no_dev, (const uint32_t*)bitmap, (const int*)prefix,                         \
      index_dims(batch, iD, iH, iW).nwords, iD, iH, iW, cv, nbr, ld)
  if (cv.k[0] == 3 && cv.k[1] == 3 && cv.k[2] == 3) COMB_NBRMAP(3, 3, 3);
  else if (cv.k[0] == 3 && cv.k[1] == 1 && cv.k[2] == 1) COMB_NBRMAP(3, 1, 1);
  else if (cv.k[0] == 1 && cv.k[1] == 1 && cv.k[2] == 1) COMB_NBRMAP(1, 1, 1);
  else if (cv.k[0] == 2 && cv.k[1] == 2 && cv.k[2] == 2) COMB_NBRMAP(2, 2, 2);
  else if (cv.k[0] == 1 && cv.k[1] == 3 && cv.k[2] == 3) COMB_NBRMAP(1, 3, 3);
  else {
    set_error("comb_nbrmap_build_indexed: kernel %dx%dx%d not instantiated (3x3x3, 3x1x1, 1x3x3, 2x2x2, 1x1x1)", cv.k[0],
              cv.k[1], cv.k[2]);
    return COMB_EINVAL;
  }
#undef COMB_NBRMAP
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}
