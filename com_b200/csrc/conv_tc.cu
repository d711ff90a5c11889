// a7 — sparse convolution forward on the 5th-gen tensor cores (tcgen05 / TMEM), bf16 in, fp32 accumulate.
//
// Replaces the gather-GEMM-scatter of spconv's SubMConv3d / SparseConv3d forward
// (pcdet/models/backbones_3d/spconv_backbone.py:191-232) with ONE persistent, warp-specialised,
// output-stationary implicit GEMM per layer:
//
//   D[128 rows, Cout] (TMEM, fp32)  +=  A_chunk[128, 64] (smem, bf16)  x  B_chunk[Cout, 64]^T (smem, bf16)
//
// The GEMM K dimension is the concatenation of all kernel offsets: K_total = K * Cin, walked in
// chunks of 64 elements (= one 128-byte swizzle row).  A chunk holds 64/Cin kernel offsets when
// Cin <= 64 (4 offsets at Cin=16) or half an offset at Cin=128.  Row r of an A chunk is the
// concatenation of the feature rows of r's neighbours at those offsets — gathered straight from
// global/L2 with 16-byte cp.async into the canonical K-major SWIZZLE_128B layout (missing
// neighbours are zero-filled by cp.async with src-size 0, costing no memory traffic).
// There is no scatter and no atomic: each output row is owned by one CTA and written once with
// bias / folded BatchNorm / residual / ReLU applied in the epilogue.
//
// Warp roles (768 threads, one CTA per SM, persistent over row tiles):
//   warps 0-15 producers : neighbour-index tile prefetch (cp.async 4 B, double buffered) and the A gather through
//                          registers: 2 x LDG.128 (only for neighbours that exist) -> 2 x STS.128 per thread per
//                          chunk, software-pipelined one chunk deep; a warp instruction covers 32/PPO consecutive
//                          rows of ONE kernel offset (few distinct 128-byte lines when rows are spatially ordered).
//                          r1 profiles: one producer warp per sub-partition was issue-latency bound (~1100
//                          cycles/chunk); cp.async tops out at ~17 B/clk/SM zero-fill included (~900 cycles/chunk).
//   warps 16-19 epilogue : tcgen05.ld 32 lanes x Cout columns -> registers -> epilogue -> global.
//   warp  20   MMA       : lane 0 issues 4 x tcgen05.mma (M=128, N=Cout, K=16) per chunk, commits to the
//                          stage's empty barrier; allocates / frees TMEM (2 accumulators).
//   warp  21   weights   : lane 0 streams the pre-swizzled weight chunk of every stage with cp.async.bulk, or,
//                          when the whole packed image is <= 64 KB (Cin,Cout <= 32), loads it once and keeps it
//                          resident.
#include <stdlib.h>
#include "common.cuh"
#include "conv_impl.cuh"
#include "tc_util.cuh"

namespace comb {
namespace {

constexpr int kBM = 128;          // rows per tile (UMMA M)
constexpr int kChunkK = 64;       // bf16 elements per K chunk (128 bytes)
constexpr int kABytes = kBM * 128;
constexpr int kProdWarps = 16;    // gather warps: 4 groups of 4 warps; group g owns the chunks G with G % 4 == g
constexpr int kProducers = kProdWarps * 32;
constexpr int kGroups = 4;
constexpr int kGroupThreads = kProducers / kGroups;       // 128: one thread per 16-byte column pair of 8 pieces
constexpr int kGroupWarps = kProdWarps / kGroups;
constexpr int kEpiWarp0 = kProdWarps;      // warps 16..19: (warp & 3) = TMEM lane quarter
constexpr int kMmaWarp = kProdWarps + 4;   // warp 20
// warp 21 (kProdWarps + 5): weight streamer
constexpr int kIdxWarp = kProdWarps + 6;   // warp 22: neighbour-index tile prefetcher
constexpr int kSentinelWarp = kProdWarps + 7;   // warp 23: stage sentinel (barrier wait + proxy fence ahead of the MMA warp)
constexpr int kThreadsTC = (kProdWarps + 8) * 32;
constexpr int kPiecesPerThread = (kBM * 8) / kGroupThreads;  // 8 x 16-byte pieces of one A chunk per group thread
constexpr int kMaxK = 32;         // kernel offsets supported by the index tile
constexpr int kBResidentMax = 64 * 1024;   // packed weights up to this size stay in shared memory for the whole kernel

using namespace tcu;


template <int CIN, int COUT>
struct TcCfg {
  static constexpr int kOffPerChunk = CIN <= 64 ? 64 / CIN : 1;  // kernel offsets per K chunk
  static constexpr int kChunksPerOff = CIN <= 64 ? 1 : CIN / 64;
  static constexpr int kPPO = CIN <= 64 ? CIN / 8 : 8;            // 16-byte pieces of one offset inside a chunk
  static constexpr int kBBytes = COUT * 128;
  static constexpr int kTmemCols = 2 * COUT < 32 ? 32 : 2 * COUT;
  static __host__ __device__ int num_chunks(int K) {
    return CIN <= 64 ? (K + kOffPerChunk - 1) / kOffPerChunk : K * kChunksPerOff;
  }
  static __host__ __device__ int k_pad(int K) { return CIN <= 64 ? num_chunks(K) * kOffPerChunk : K; }
  static __host__ __device__ bool b_resident(int K) { return (size_t)num_chunks(K) * kBBytes <= (size_t)kBResidentMax; }
  // A pipeline stage holds S consecutive K chunks (A: S x 16 KB, plus their weight chunks when streamed) and is
  // handed to the MMA warp as a whole: one barrier wait, 4*S tcgen05.mma and one commit per stage.  (r1 trace:
  // with S = 1 the MMA warp's own loop — try_wait, descriptor arithmetic, 4 UTCHMMA, commit — took ~700 cycles
  // per chunk and bounded every layer, whatever the producers did.)
  static __host__ __device__ int sub_bytes(int K) { return kABytes + (b_resident(K) ? 0 : kBBytes); }
  static __host__ __device__ int fixed_bytes(int K) {
    return 1024 + 2 * k_pad(K) * kBM * 4 + (b_resident(K) ? num_chunks(K) * kBBytes : 0) + 512;
  }
  static __host__ __device__ int subs(int K) {   // S: largest of 4, 2, 1 that still leaves two stages
    const int avail = 225 * 1024 - fixed_bytes(K);
    for (int S = 4; S > 1; S >>= 1)
      if (avail / (S * sub_bytes(K)) >= 2 && num_chunks(K) >= S) return S;
    return 1;
  }
  static __host__ __device__ int stages(int K, int S) {
    int n = (225 * 1024 - fixed_bytes(K)) / (S * sub_bytes(K));
    const int cap = 8 / S;           // at most 8 chunks in flight
    return n > cap ? cap : n;
  }
  static size_t smem_bytes(int K, int S) { return (size_t)fixed_bytes(K) + (size_t)stages(K, S) * S * sub_bytes(K); }
};

struct TcParams {
  const __nv_bfloat16* in;
  const uint8_t* wpacked;
  const int* nbr;
  int ld, no_max;
  const int* no_dev;
  int K;
  int epi;
  const float* bias;
  const float* scale;
  const float* shift;
  const __nv_bfloat16* residual;
  void* out;
  int out_f32;
  long long* dbg;   // optional trace buffer (comb_debug_conv_trace), NULL in production
  int S;            // K chunks per pipeline stage (1, 2 or 4), chosen on the host
};

// trace layout: dbg[(G * 8 + slot)] for G < kDbgChunks, CTA 0 only; slots: 0 mma:full seen, 1 mma:issued+committed,
// 2 prod(group of G, warp 0 of the group):loads issued, 3 prod:empty seen, 4 prod:stored+arrived, 5 epi: tile done (G = tile)
constexpr int kDbgChunks = 512;
__device__ __forceinline__ void dbg_stamp(long long* dbg, int G, int slot) {
  if (dbg != nullptr && blockIdx.x == 0 && G < kDbgChunks) {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
    dbg[G * 8 + slot] = t;
  }
}

__device__ __forceinline__ void dbg_cta_time(long long* dbg, int slot) {   // per-CTA wall clock (ns), slots 0..3
  if (dbg != nullptr && blockIdx.x < 256) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    dbg[kDbgChunks * 8 + blockIdx.x * 4 + slot] = t;
  }
}

constexpr int kMaxStages = 8;

template <int CIN, int COUT>
__global__ void __launch_bounds__(kThreadsTC, 1) spconv_tc_kernel(TcParams p) {
  using Cfg = TcCfg<CIN, COUT>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_full[kMaxStages], bar_ready[kMaxStages], bar_empty[kMaxStages], bar_tfull[2], bar_tempty[2],
      bar_idx[2], bar_idx_free[2], bar_b;
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) dbg_cta_time(p.dbg, 0);
  const int no = eff_n(p.no_max, p.no_dev);
  const int ntiles = (no + kBM - 1) / kBM;
  const int K = p.K;
  const int nchunks = Cfg::num_chunks(K);
  const int kpad = Cfg::k_pad(K);
  const bool bres = Cfg::b_resident(K);
  const int S = p.S;                                // K chunks per pipeline stage
  const int NS = Cfg::stages(K, S);
  const int sub_bytes = Cfg::sub_bytes(K);
  const int stage_bytes = S * sub_bytes;
  const int nstages_tile = (nchunks + S - 1) / S;   // pipeline stages per row tile
  const int nsub_pad = nstages_tile * S;            // chunk slots per tile (the last stage may be partial)

  // shared memory map (1024-byte aligned):
  //   [NS stages: S x A chunk (16 KB) | S x B chunk if streamed] [resident B] [index tiles 2 x kpad x 128 ints]
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bres_base = smem_base + NS * stage_bytes;
  const uint32_t idx_base = bres_base + (bres ? nchunks * Cfg::kBBytes : 0);
  const uint32_t idx_buf_bytes = (uint32_t)kpad * kBM * 4;
  const int my_tiles = ntiles > (int)blockIdx.x ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(smem_u32(&bar_full[s]), S * kGroupWarps + (bres ? 0 : 1));
      mbar_init(smem_u32(&bar_empty[s]), 1);
      mbar_init(smem_u32(&bar_ready[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&bar_tfull[a]), 1);
      mbar_init(smem_u32(&bar_tempty[a]), 4);
      mbar_init(smem_u32(&bar_idx[a]), 32);
      mbar_init(smem_u32(&bar_idx_free[a]), kGroups);
    }
    mbar_init(smem_u32(&bar_b), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                 "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // index-tile rows of the padding offsets (k in [K, kpad)) stay -1 for the whole kernel
  for (int e = tid; e < 2 * (kpad - K) * kBM; e += kThreadsTC) {
    const int buf = e / ((kpad - K) * kBM), r = e - buf * (kpad - K) * kBM;
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(idx_base + buf * idx_buf_bytes + (K * kBM + r) * 4), "r"(-1) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  if (tid == 0) dbg_cta_time(p.dbg, 1);

  if (warp < kProdWarps) {
    // ===================== producers: A gather (global/L2 -> registers -> swizzled shared memory) ==========
    // Chunk slots are numbered G = it*nsub_pad + c over this CTA's tiles; group g (4 warps, 128 threads) owns the
    // slots with G % 4 == g and moves the whole 128 x 128 B chunk: 8 x LDG.128 per thread (only for neighbours
    // that exist) -> 8 x STS.128.  While one group waits for its loads the other three issue theirs.
    // Piece e = j*128 + t:  piece-in-offset = t % PPO, row = ((j % PPO)*128 + t) / PPO, offset slot = j / PPO:
    // a warp instruction covers 32/PPO CONSECUTIVE rows of ONE kernel offset — with rows in key order the
    // gathered addresses are neighbours in memory (few 128-byte lines per instruction, L1 reuse across dx).
    const int grp = warp / kGroupWarps, t = tid % kGroupThreads;
    const uint8_t* src_base = reinterpret_cast<const uint8_t*>(p.in) + (t % Cfg::kPPO) * 16;
    const int log2S = S == 4 ? 2 : (S == 2 ? 1 : 0);
    int gs_prev = 0, s = 0;      // stage bookkeeping without divisions: gs = G >> log2S only ever grows
    uint32_t ph = 0;
    for (int it = 0; it < my_tiles; ++it) {
      const int buf = it & 1;
      const uint32_t idx_tile = idx_base + (uint32_t)buf * idx_buf_bytes;
      mbar_wait(smem_u32(&bar_idx[buf]), (it >> 1) & 1);
      int c = (grp - it * nsub_pad) & (kGroups - 1);   // first chunk slot of this tile owned by the group
      for (; c < nsub_pad; c += kGroups) {
        const int G = it * nsub_pad + c;
        const int gs = G >> log2S;               // global stage counter
        const int sub = G & (S - 1);
        s += gs - gs_prev;
        gs_prev = gs;
        while (s >= NS) { s -= NS; ph ^= 1u; }
        uint4 v[kPiecesPerThread];
        if (c < nchunks) {
          const uint32_t idx_c =
              idx_tile + (uint32_t)(CIN <= 64 ? c * Cfg::kOffPerChunk : c / Cfg::kChunksPerOff) * kBM * 4;
          const uint32_t half = CIN > 64 ? (uint32_t)(c % Cfg::kChunksPerOff) * 128u : 0u;
#pragma unroll
          for (int j = 0; j < kPiecesPerThread; ++j) {
            const int row = ((j % Cfg::kPPO) * kGroupThreads + t) / Cfg::kPPO, slot = j / Cfg::kPPO;
            int rw;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(rw) : "r"(idx_c + (uint32_t)(slot * kBM + row) * 4) : "memory");
            v[j] = make_uint4(0u, 0u, 0u, 0u);
            if (rw >= 0) v[j] = __ldg(reinterpret_cast<const uint4*>(src_base + (size_t)(uint32_t)rw * (CIN * 2) + half));
          }
        }
        if (t == 0) dbg_stamp(p.dbg, G, 2);
        mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
        if (t == 0) dbg_stamp(p.dbg, G, 3);
        if (c < nchunks) {   // padding slots of a partial last stage only arrive
          const uint32_t a_base = smem_base + s * stage_bytes + sub * kABytes;
#pragma unroll
          for (int j = 0; j < kPiecesPerThread; ++j) {
            const int row = ((j % Cfg::kPPO) * kGroupThreads + t) / Cfg::kPPO, slot = j / Cfg::kPPO;
            const int q = slot * Cfg::kPPO + (t % Cfg::kPPO);           // 16-byte column of the 128-byte smem row
            const uint32_t dst = a_base + (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((q ^ (row & 7)) << 4));
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v[j].x), "r"(v[j].y), "r"(v[j].z),
                         "r"(v[j].w)
                         : "memory");
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bar_full[s]));   // release: the warp's stores happen-before the MMA warp's wait
        if (t == 0) dbg_stamp(p.dbg, G, 4);
      }
      // the group has issued (and consumed) all its index reads of this tile: hand the buffer back
      asm volatile("bar.sync %0, %1;" ::"r"(2 + grp), "n"(kGroupThreads) : "memory");
      if (t == 0) mbar_arrive(smem_u32(&bar_idx_free[buf]));
    }
  } else if (warp < kEpiWarp0 + 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const int a = it & 1;
      mbar_wait(smem_u32(&bar_tfull[a]), (it >> 1) & 1);
      tc_fence_after();
      const int row = tile * kBM + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * COUT;
#pragma unroll
      for (int c0 = 0; c0 < COUT; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
        if (row < no) {
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
          if (p.epi & COMB_EPI_BIAS) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] += __ldg(p.bias + c0 + i);
          }
          if (p.epi & COMB_EPI_AFFINE) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fmaf(f[i], __ldg(p.scale + c0 + i), __ldg(p.shift + c0 + i));
          }
          if (p.epi & COMB_EPI_RESIDUAL) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + (size_t)row * COUT + c0);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint4 rv = __ldg(rp + h);
              const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float2 tt = __bfloat1622float2(r2[i]);
                f[h * 8 + 2 * i] += tt.x;
                f[h * 8 + 2 * i + 1] += tt.y;
              }
            }
          }
          if (p.epi & COMB_EPI_RELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.0f);
          }
          if (p.out_f32) {
            float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (size_t)row * COUT + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i) op[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
          } else {
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * COUT + c0);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint4 o;
              __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int i = 0; i < 4; ++i) o2[i] = __floats2bfloat162_rn(f[h * 8 + 2 * i], f[h * 8 + 2 * i + 1]);
              op[h] = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[a]));
      if (warp == kEpiWarp0 && lane == 0) dbg_stamp(p.dbg, it, 5);
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    // instruction descriptor: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, K-major A/B, N>>3 [17,23), M>>4 [24,29)
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(COUT >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
    int s = 0;
    uint32_t ph = 0;
    if (bres) mbar_wait(smem_u32(&bar_b), 0);
    for (int it = 0; it < my_tiles; ++it) {
      const int a = it & 1;
      mbar_wait(smem_u32(&bar_tempty[a]), ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + a * COUT;
      for (int st = 0; st < nstages_tile; ++st) {
        // the sentinel warp has already waited for the stage and issued the generic->async proxy fence
        mbar_wait(smem_u32(&bar_ready[s]), ph);
        if (lane == 0) dbg_stamp(p.dbg, (it * nstages_tile + st) * S, 0);
        tc_fence_after();
        if (lane == 0) dbg_stamp(p.dbg, (it * nstages_tile + st) * S, 6);
        const uint32_t stage_addr = smem_base + s * stage_bytes;
        const int c0 = st * S;
        const int nsub = nchunks - c0 < S ? nchunks - c0 : S;
        if (elect_one_sync()) {
          for (int sub = 0; sub < nsub; ++sub) {
            const uint64_t adesc = make_desc_sw128(stage_addr + sub * kABytes);
            const uint64_t bdesc = make_desc_sw128(bres ? bres_base + (c0 + sub) * Cfg::kBBytes
                                                        : stage_addr + S * kABytes + sub * Cfg::kBBytes);
#pragma unroll
            for (int kk = 0; kk < kChunkK / 16; ++kk) {
              // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in the (>>4) address field
              umma_bf16(tmem_d, adesc + 2 * kk, bdesc + 2 * kk, idesc, (c0 | sub | kk) != 0 ? 1u : 0u);
            }
          }
          umma_commit(smem_u32(&bar_empty[s]));
          if (st == nstages_tile - 1) umma_commit(smem_u32(&bar_tfull[a]));
        }
        __syncwarp();
        if (lane == 0) dbg_stamp(p.dbg, (it * nstages_tile + st) * S, 1);
        if (++s == NS) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == kSentinelWarp) {
    // ===================== stage sentinel (warp 23, one lane) =====================
    // Waits for a stage to be full, makes the producers' st.shared (generic proxy, acquired through the barrier)
    // visible to the tensor core's async-proxy reads with ONE proxy fence, and hands the stage to the MMA warp.
    // This keeps barrier-wake-up and fence latency off the MMA warp, whose issue loop bounds the kernel
    // (r1 trace: ~950 cycles of tcgen05.mma issue per 4-chunk stage + ~600 of wait/fence in between; a fence in
    // every producer thread instead cost each gather group ~600 cycles per chunk).
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const int total = my_tiles * nstages_tile;
      for (int g = 0; g < total; ++g) {
        mbar_wait(smem_u32(&bar_full[s]), ph);
        fence_proxy_async();
        mbar_arrive(smem_u32(&bar_ready[s]));
        if (++s == NS) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == kIdxWarp) {
    // ===================== neighbour-index prefetcher (warp 22) =====================
    // idx[buf][k][r] = nbr[k][tile*128 + r] (-1 beyond the last row), double buffered one tile ahead.
    const bool vec_ok = (p.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.nbr) & 15) == 0);
    for (int it = 0; it < my_tiles; ++it) {
      const int buf = it & 1;
      if (it >= 2) mbar_wait(smem_u32(&bar_idx_free[buf]), ((it - 2) >> 1) & 1);
      const int row0 = ((int)blockIdx.x + it * (int)gridDim.x) * kBM + lane * 4;   // this lane: 4 consecutive rows
      const uint32_t dst = idx_base + (uint32_t)buf * idx_buf_bytes + lane * 16;
      if (vec_ok && row0 + 3 < no) {
        for (int k = 0; k < K; ++k)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + k * kBM * 4),
                       "l"(p.nbr + (size_t)k * p.ld + row0)
                       : "memory");
      } else {
        for (int k = 0; k < K; ++k)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (row0 + q < no)
              cp_async4(dst + k * kBM * 4 + q * 4, p.nbr + (size_t)k * p.ld + row0 + q);
            else
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + k * kBM * 4 + q * 4), "r"(-1) : "memory");
          }
      }
      cp_async_mbar_arrive_noinc(smem_u32(&bar_idx[buf]));   // fires when this lane's copies have landed
    }
  } else if (warp == kMmaWarp + 1 && lane == 0) {
    // ===================== weight streamer (warp 21, one lane) =====================
    if (bres) {
      // small layers: the whole pre-swizzled weight image stays resident
      const uint32_t bytes = (uint32_t)nchunks * Cfg::kBBytes;
      mbar_arrive_expect_tx(smem_u32(&bar_b), bytes);
      bulk_copy_g2s(bres_base, p.wpacked, bytes, smem_u32(&bar_b));
    } else {
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < my_tiles; ++it) {
        for (int st = 0; st < nstages_tile; ++st) {
          const int c0 = st * S;
          const int nsub = nchunks - c0 < S ? nchunks - c0 : S;
          mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
          mbar_arrive_expect_tx(smem_u32(&bar_full[s]), (uint32_t)nsub * Cfg::kBBytes);
          bulk_copy_g2s(smem_base + s * stage_bytes + S * kABytes, p.wpacked + (size_t)c0 * Cfg::kBBytes,
                        (uint32_t)nsub * Cfg::kBBytes, smem_u32(&bar_full[s]));
          if (++s == NS) { s = 0; ph ^= 1; }
        }
      }
    }
  }

  if (warp == kMmaWarp && lane == 0) dbg_cta_time(p.dbg, 2);
  tc_fence_before();
  __syncthreads();
  if (tid == 0) dbg_cta_time(p.dbg, 3);
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
  }
}

// Pre-swizzle the weights into the shared-memory image of every B chunk:
// chunk c, row n (= output channel), K element kk -> W[n][k][ci]
template <int CIN>
__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ w, int Cout, int K, int Cin_real,
                                                           int nchunks, __nv_bfloat16* __restrict__ out) {
  const long long total = (long long)nchunks * Cout * kChunkK;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(e % kChunkK);
    const int n = (int)((e / kChunkK) % Cout);
    const int c = (int)(e / ((long long)kChunkK * Cout));
    int k, ci;
    if constexpr (CIN <= 64) {
      k = c * (64 / CIN) + kk / CIN;
      ci = kk % CIN;
    } else {
      k = c / (CIN / 64);
      ci = (c % (CIN / 64)) * 64 + kk;
    }
    float v = 0.0f;
    if (k < K && ci < Cin_real) v = w[((size_t)n * K + k) * Cin_real + ci];
    const int q = kk >> 3, within = kk & 7;
    const size_t byte_off = (size_t)c * Cout * 128 + (size_t)(n >> 3) * 1024 + (n & 7) * 128 + ((q ^ (n & 7)) << 4) + within * 2;
    out[byte_off / 2] = __float2bfloat16(v);
  }
}

template <int CIN, int COUT>
int launch_tc(const TcParams& p_in, cudaStream_t stream) {
  using Cfg = TcCfg<CIN, COUT>;
  TcParams p = p_in;
  // chunks per stage: COMB_TC_S (tuning knob) or the largest of 4, 2, 1 that still leaves two stages
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("COMB_TC_S");
    forced = e ? atoi(e) : 0;
  }
  p.S = 1;
  for (int S = (forced == 1 || forced == 2 || forced == 4) ? forced : 4; S > 1; S >>= 1)
    if (Cfg::stages(p.K, S) >= 2 && Cfg::num_chunks(p.K) >= S) {
      p.S = S;
      break;
    }
  const size_t smem = Cfg::smem_bytes(p.K, p.S);
  static thread_local DevOnce configured;   // per device: the attribute is a per-device property
  if (configured.first()) {
    COMB_CUDA(cudaFuncSetAttribute(spconv_tc_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
  }
  if (smem > 227 * 1024 - 2048 || Cfg::stages(p.K, p.S) < 2) {
    set_error("comb_spconv_fwd_bf16: shared memory %zu exceeds the per-CTA limit", smem);
    return COMB_EINVAL;
  }
  const int ntiles = cdiv(p.no_max, kBM);
  int grid = ntiles < sm_count() ? ntiles : sm_count();
  spconv_tc_kernel<CIN, COUT><<<grid, kThreadsTC, smem, stream>>>(p);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

template <int CIN>
int dispatch_cout(int Cout, const TcParams& p, cudaStream_t stream) {
  switch (Cout) {
    case 16: return launch_tc<CIN, 16>(p, stream);
    case 32: return launch_tc<CIN, 32>(p, stream);
    case 64: return launch_tc<CIN, 64>(p, stream);
    case 128: return launch_tc<CIN, 128>(p, stream);
  }
  set_error("comb_spconv_fwd_bf16: Cout %d not in {16,32,64,128}", Cout);
  return COMB_EINVAL;
}

static long long* g_conv_trace = nullptr;

static int chunks_for(int Cin_p, int K) {
  return Cin_p <= 64 ? (K + 64 / Cin_p - 1) / (64 / Cin_p) : K * (Cin_p / 64);
}

// COMB_CONV_IMPL: "ss" = the shared-memory-A kernel of this file, "ts" = tensor-memory A gathered in 16x256b fragments
// (conv_ts.cu) for every layer, "tr" = the row-per-thread kernel (conv_tr.cu) for every layer; unset = per layer, see
// use_tr().  Read once per process: the packed weight image depends on it.
static int conv_impl() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("COMB_CONV_IMPL");
    v = 0;                                              // auto
    if (e && e[0] == 's' && e[1] == 's') v = 1;
    else if (e && e[0] == 't' && e[1] == 's') v = 2;
    else if (e && e[0] == 't' && e[1] == 'r') v = 3;
  }
  return v;
}
static bool use_ts() { return conv_impl() != 1; }       // "one of the tensor-memory kernels"
// which tensor-memory kernel runs a layer; the same predicate decides the layout of its packed weights
static bool use_tr(int Cin_p, int Cout, int K) {
  if (conv_impl() == 3) return tr_supported(Cin_p, Cout, K);
  if (conv_impl() == 2) return false;
  // auto: the row-per-thread kernel for the narrow inputs (r2, bench frames: 22.5 vs 29.4 us at 16x16, 45.6 vs 49.6 at
  // 32x32; at Cin = 64 the 16x256b gather of conv_ts is faster, 48 vs 67 us)
  return Cin_p <= 32 && tr_supported(Cin_p, Cout, K);
}

}  // namespace
}  // namespace comb

using namespace comb;

// Debug hook (not part of the product path): when set, CTA 0 of every following comb_spconv_fwd_bf16 launch
// records clock64 stamps of its pipeline events for the first 512 chunks into buf (512*8 int64).
extern "C" int comb_debug_conv_trace(void* buf) {
  g_conv_trace = (long long*)buf;
  return COMB_OK;
}

extern "C" size_t comb_spconv_packed_bytes(int Cin_p, int K, int Cout) {
  if (!(Cin_p == 16 || Cin_p == 32 || Cin_p == 64 || Cin_p == 128) || K < 1 || Cout < 8) return 0;
  return (size_t)chunks_for(Cin_p, K) * Cout * 128;
}

extern "C" int comb_spconv_pack_weight_bf16(const float* weight, int Cout, int K, int Cin, int Cin_p, void* wpacked,
                                            void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(weight && wpacked, "comb_spconv_pack_weight_bf16: null pointer");
  COMB_CHECK_ARG(Cin >= 1 && Cin <= Cin_p, "comb_spconv_pack_weight_bf16: Cin %d > padded %d", Cin, Cin_p);
  COMB_CHECK_ARG(Cout % 8 == 0 && Cout >= 8 && K >= 1, "comb_spconv_pack_weight_bf16: bad Cout/K");
  const int nchunks = chunks_for(Cin_p, K);
  if (use_ts()) return ts_pack_weight(weight, Cout, K, Cin, Cin_p, nchunks, use_tr(Cin_p, Cout, K) ? 1 : 0, wpacked, stream);
  const long long total = (long long)nchunks * Cout * kChunkK;
  const int grid = cdiv(total, 256);
  __nv_bfloat16* out = (__nv_bfloat16*)wpacked;
  switch (Cin_p) {
    case 16: pack_weight_kernel<16><<<grid, 256, 0, stream>>>(weight, Cout, K, Cin, nchunks, out); break;
    case 32: pack_weight_kernel<32><<<grid, 256, 0, stream>>>(weight, Cout, K, Cin, nchunks, out); break;
    case 64: pack_weight_kernel<64><<<grid, 256, 0, stream>>>(weight, Cout, K, Cin, nchunks, out); break;
    case 128: pack_weight_kernel<128><<<grid, 256, 0, stream>>>(weight, Cout, K, Cin, nchunks, out); break;
    default: COMB_CHECK_ARG(false, "comb_spconv_pack_weight_bf16: Cin_p %d not in {16,32,64,128}", Cin_p);
  }
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

extern "C" int comb_spconv_fwd_bf16(const void* in_feats, int Cin_p, const void* wpacked, int K, int Cout,
                                    const int* nbr, int ld, int no_max, const int* no_dev, int epi_flags,
                                    const float* bias, const float* scale, const float* shift, const void* residual,
                                    void* out, int out_dtype, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(K >= 1 && K <= kMaxK, "comb_spconv_fwd_bf16: K %d outside [1,%d]", K, kMaxK);
  COMB_CHECK_ARG(ld >= no_max && no_max >= 0, "comb_spconv_fwd_bf16: bad ld/no");
  COMB_CHECK_ARG(!(epi_flags & COMB_EPI_BIAS) || bias, "comb_spconv_fwd_bf16: bias flag without pointer");
  COMB_CHECK_ARG(!(epi_flags & COMB_EPI_AFFINE) || (scale && shift), "comb_spconv_fwd_bf16: affine flag without pointers");
  COMB_CHECK_ARG(!(epi_flags & COMB_EPI_RESIDUAL) || residual, "comb_spconv_fwd_bf16: residual flag without pointer");
  COMB_CHECK_ARG(out_dtype == COMB_DT_F32 || out_dtype == COMB_DT_BF16, "comb_spconv_fwd_bf16: bad out dtype");
  if (no_max == 0) return COMB_OK;
  COMB_CHECK_ARG(in_feats && wpacked && nbr && out, "comb_spconv_fwd_bf16: null pointer");
  if (use_ts()) {
    ConvFwdArgs a;
    a.in = (const __nv_bfloat16*)in_feats;
    a.wpacked = (const uint8_t*)wpacked;
    a.nbr = nbr;
    a.ld = ld;
    a.no_max = no_max;
    a.no_dev = no_dev;
    a.K = K;
    a.epi = epi_flags;
    a.bias = bias;
    a.scale = scale;
    a.shift = shift;
    a.residual = (const __nv_bfloat16*)residual;
    a.out = out;
    a.out_f32 = out_dtype == COMB_DT_F32;
    a.dbg = g_conv_trace;
    if (use_tr(Cin_p, Cout, K)) return tr_fwd_bf16(a, Cin_p, Cout, stream);
    return ts_fwd_bf16(a, Cin_p, Cout, stream);
  }
  TcParams p;
  p.in = (const __nv_bfloat16*)in_feats;
  p.wpacked = (const uint8_t*)wpacked;
  p.nbr = nbr;
  p.ld = ld;
  p.no_max = no_max;
  p.no_dev = no_dev;
  p.K = K;
  p.epi = epi_flags;
  p.bias = bias;
  p.scale = scale;
  p.shift = shift;
  p.residual = (const __nv_bfloat16*)residual;
  p.out = out;
  p.out_f32 = out_dtype == COMB_DT_F32;
  p.dbg = g_conv_trace;
  switch (Cin_p) {
    case 16: return dispatch_cout<16>(Cout, p, stream);
    case 32: return dispatch_cout<32>(Cout, p, stream);
    case 64: return dispatch_cout<64>(Cout, p, stream);
    case 128: return dispatch_cout<128>(Cout, p, stream);
  }
  set_error("comb_spconv_fwd_bf16: Cin_p %d not in {16,32,64,128}", Cin_p);
  return COMB_EINVAL;
}
