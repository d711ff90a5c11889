// a8 — weight gradient of the sparse convolution on the 5th-gen tensor cores (tcgen05 / TMEM), bf16 operands, fp32
// accumulate.
//
// Replaces the wgrad GEMMs of spconv's SparseConvolution backward (autograd Function behind SubMConv3d / SparseConv3d,
// call sites pcdet/models/backbones_3d/spconv_backbone.py:12-15,38-45,191-232):
//
//     dW[co][k][ci] = sum_o dout[o][co] * in[nbr[k][o]][ci]
//
// As a GEMM the reduction runs over the OUTPUT ROWS o, and the gathered operand is stored the way it is gathered —
// one feature row per GEMM-K index, channels contiguous — which is the "MN-major" canonical layout of the tcgen05
// shared-memory descriptors.  No transposition pass is needed, both operands are MN-major (instruction descriptor
// bits 15 / 16):
//
//     D[(tap, ci), co] (TMEM, fp32, M = 128)  +=  A[(tap, ci), o]  x  B[o, co]            o = 16 rows per tcgen05.mma
//       A: gathered input rows of 128 / Cin kernel offsets side by side   (shared memory, 64 rows x 128 B blocks)
//       B: the dout rows of the same 64 output rows, shared by every kernel offset of the CTA
//
// Work split: grid = (row chunks, kernel-offset groups).  A CTA owns a group of kernel offsets small enough for its
// accumulators to stay in tensor memory for the whole kernel (<= 4 M-tiles, <= 512 columns) and one contiguous chunk
// of 64-row tiles; it writes its partial dW to a workspace with plain stores and a second kernel sums the row chunks
// in a fixed order (deterministic, no atomics — the fp32 check kernel of conv_f32.cu uses fp32 atomics).
//
// Warp roles (544 threads, one CTA per SM): warps 0-15 gather (LDG.128 -> STS.128 into SWIZZLE_128B blocks, 8 loads in
// flight per thread, missing neighbours and rows beyond the end are zero filled); warp 16 issues the MMAs (one
// elected lane) and owns the TMEM allocation; warps 0-3 read the accumulators back at the end (tcgen05.ld).
#include "common.cuh"
#include "tc_util.cuh"

namespace comb {
namespace {

using namespace tcu;

constexpr int kR = 64;                       // rows (GEMM K) per pipeline stage: 4 MMAs of K = 16 per M-tile
constexpr int kCbBytes = kR * 128;           // one column block: 64 rows x 64 bf16
constexpr int kWgProdWarps = 16;
constexpr int kWgProducers = kWgProdWarps * 32;
constexpr int kWgMmaWarp = kWgProdWarps;
constexpr int kWgThreads = (kWgProdWarps + 1) * 32;
constexpr int kWgMaxStages = 4;
constexpr int kWgUnroll = 8;                 // gathered 16-byte pieces in flight per thread

struct WgParams {
  const __nv_bfloat16* in;     // [ni, CIN] bf16 (CIN = padded channel count)
  const __nv_bfloat16* dout;   // [no, Cout] bf16
  const int* nbr;
  int ld, no_max;
  const int* no_dev;
  int K, Cin_real, Cout;
  int taps_per_group;          // kernel offsets per blockIdx.y
  int mt_alloc;                // M-tiles a stage is laid out for (that of a full group)
  int tiles_per_chunk, ntiles; // 64-row tiles per blockIdx.x / in total
  int NS;                      // pipeline stages
  int tmem_cols;
  float* partial;              // [gridDim.x][Cout * K * Cin_real]
};

// MN-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor; canonical layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): 64 contiguous bf16 along M/N per 128-byte row, one row per K
// index, 8 rows = one 1024-byte swizzle atom; SBO = bytes between 8-row groups along K (1024: rows are packed),
// LBO = bytes between 64-element blocks along M/N (one column block).
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ uint32_t swz_off(int row, int q) {   // byte offset of 16-byte piece q of a block row
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((q ^ (row & 7)) << 4));
}

template <int CIN, int NB>
__global__ void __launch_bounds__(kWgThreads, 1) spconv_wgrad_tc_kernel(WgParams p) {
  constexpr int PPO = CIN / 8;                      // 16-byte pieces of one gathered row
  constexpr int TPC = CIN <= 64 ? 64 / CIN : 1;     // kernel offsets per column block
  constexpr int LOG_RP = CIN == 16 ? 7 : (CIN == 32 ? 8 : (CIN == 64 ? 9 : 10));   // log2(kR * PPO)
  constexpr int LOG_PPO = LOG_RP - 6;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_full[kWgMaxStages], bar_empty[kWgMaxStages], bar_done;
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int no = eff_n(p.no_max, p.no_dev);
  const int k0 = (int)blockIdx.y * p.taps_per_group;
  const int ntap = min(p.taps_per_group, p.K - k0);
  const int ncb = CIN <= 64 ? (ntap + TPC - 1) / TPC : ntap * 2;   // column blocks that carry data
  const int MT = (ncb + 1) >> 1;                                   // M-tiles of this CTA (<= p.mt_alloc)
  const int tile0 = (int)blockIdx.x * p.tiles_per_chunk;
  const int my_tiles = min(p.tiles_per_chunk, p.ntiles - tile0);   // >= 1 by construction of the grid
  const int NS = p.NS;
  const uint32_t b_off = (uint32_t)(2 * p.mt_alloc) * kCbBytes;    // B block(s) behind the A blocks of a stage
  const uint32_t stage_bytes = b_off + (NB / 64) * kCbBytes;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int Cout = p.Cout;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(smem_u32(&bar_full[s]), kWgProdWarps);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWgMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero the stages once: pieces that are never written (padding kernel offsets of the last M-tile, dout columns
  // beyond Cout) must hold finite values — they only reach accumulator lanes / columns nobody reads
  for (uint32_t off = (uint32_t)tid * 16; off < (uint32_t)NS * stage_bytes; off += kWgThreads * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(smem_base + off), "r"(0) : "memory");
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  if (warp < kWgProdWarps) {
    // ===================== producers =====================
    const uint8_t* in_b = reinterpret_cast<const uint8_t*>(p.in);
    const uint8_t* dout_b = reinterpret_cast<const uint8_t*>(p.dout);
    const int nA = ntap << LOG_RP;                   // 16-byte pieces of the A blocks of one stage
    const int log_ppb = Cout == 16 ? 1 : (Cout == 32 ? 2 : (Cout == 64 ? 3 : 4));
    const int nB = kR << log_ppb;
    // One global-memory latency per stage on the critical path instead of three: the neighbour indices of a tile's
    // first batch of pieces are loaded one tile ahead (idx_next), the dout pieces are loaded together with the
    // gathered rows, and the wait for the stage to be free comes after the loads have been issued.
    auto load_idx = [&](int r0, int e0, int (&idx)[kWgUnroll]) {
#pragma unroll
      for (int u = 0; u < kWgUnroll; ++u) {
        const int e = e0 + u * kWgProducers;
        const int t = e >> LOG_RP, row = (e >> LOG_PPO) & (kR - 1);
        idx[u] = -1;
        if (e < nA && r0 + row < no) idx[u] = __ldg(p.nbr + (size_t)(k0 + t) * p.ld + r0 + row);
      }
    };
    int idx_next[kWgUnroll];
    load_idx(tile0 * kR, tid, idx_next);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < my_tiles; ++it) {
      const int r0 = (tile0 + it) * kR;
      const uint32_t stage = smem_base + (uint32_t)s * stage_bytes;
      uint4 vb[2];                                   // nB <= 1024 pieces: at most two per thread
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int e = tid + u * kWgProducers;
        const int row = e >> log_ppb, pc = e & ((1 << log_ppb) - 1);
        vb[u] = make_uint4(0u, 0u, 0u, 0u);
        if (e < nB && r0 + row < no)
          vb[u] = __ldg(reinterpret_cast<const uint4*>(dout_b + (size_t)(r0 + row) * (Cout * 2) + pc * 16));
      }
      bool waited = false;
      for (int e0 = tid; e0 < nA; e0 += kWgProducers * kWgUnroll) {
        int idx[kWgUnroll];
        if (e0 == tid) {
#pragma unroll
          for (int u = 0; u < kWgUnroll; ++u) idx[u] = idx_next[u];
        } else {
          load_idx(r0, e0, idx);
        }
        uint4 v[kWgUnroll];
#pragma unroll
        for (int u = 0; u < kWgUnroll; ++u) {
          const int pc = (e0 + u * kWgProducers) & (PPO - 1);
          v[u] = make_uint4(0u, 0u, 0u, 0u);
          if (idx[u] >= 0) v[u] = __ldg(reinterpret_cast<const uint4*>(in_b + (size_t)(uint32_t)idx[u] * (CIN * 2) + pc * 16));
        }
        if (e0 == tid && it + 1 < my_tiles) load_idx(r0 + kR, tid, idx_next);
        if (!waited) {
          mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
          waited = true;
        }
#pragma unroll
        for (int u = 0; u < kWgUnroll; ++u) {
          const int e = e0 + u * kWgProducers;
          if (e < nA) {
            const int t = e >> LOG_RP, row = (e >> LOG_PPO) & (kR - 1), pc = e & (PPO - 1);
            int cb, q;
            if (CIN <= 64) {
              cb = t / TPC;
              q = (t % TPC) * PPO + pc;
            } else {
              cb = t * 2 + (pc >> 3);
              q = pc & 7;
            }
            const uint32_t dst = stage + (uint32_t)cb * kCbBytes + swz_off(row, q);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v[u].x), "r"(v[u].y), "r"(v[u].z),
                         "r"(v[u].w)
                         : "memory");
          }
        }
      }
      if (!waited) mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int e = tid + u * kWgProducers;
        if (e < nB) {
          const int row = e >> log_ppb, pc = e & ((1 << log_ppb) - 1);
          const uint32_t dst = stage + b_off + (uint32_t)(pc >> 3) * kCbBytes + swz_off(row, pc & 7);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(vb[u].x), "r"(vb[u].y), "r"(vb[u].z),
                       "r"(vb[u].w)
                       : "memory");
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_full[s]));   // release: the warp's stores happen-before the MMA warp's wait
      if (++s == NS) { s = 0; ph ^= 1; }
    }
    if (warp < 4) {
      // ===================== accumulator read-back (warps 0-3 = TMEM lane quarters) =====================
      mbar_wait_sleep(smem_u32(&bar_done), 0);
      tc_fence_after();
      const int L = warp * 32 + lane;                // accumulator row = (kernel offset, input channel)
      const size_t kc = (size_t)p.K * p.Cin_real;
      float* part = p.partial + (size_t)blockIdx.x * ((size_t)Cout * kc);
      for (int mt = 0; mt < MT; ++mt) {
        int slot, ci;
        if (CIN <= 64) {
          const int e = L & 63;
          slot = (2 * mt + (L >> 6)) * TPC + e / CIN;
          ci = e % CIN;
        } else {
          slot = mt;
          ci = L;
        }
        const bool ok = slot < ntap && ci < p.Cin_real;
        float* dst = part + (size_t)(k0 + slot) * p.Cin_real + ci;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mt * NB);
        for (int c0 = 0; c0 < Cout; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
          if (ok) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dst[(size_t)(c0 + i) * kc] = __uint_as_float(v[i]);
          }
        }
      }
    }
  } else {
    // ===================== MMA issuer (warp 16) =====================
    // instruction descriptor: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, A MN-major [15], B MN-major [16],
    // N>>3 [17,23), M>>4 [24,29)
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NB >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < my_tiles; ++it) {
      mbar_wait(smem_u32(&bar_full[s]), ph);
      fence_proxy_async();     // the producers' st.shared (generic proxy) -> the tensor core's async-proxy reads
      tc_fence_after();
      const uint32_t stage = smem_base + (uint32_t)s * stage_bytes;
      if (elect_one_sync()) {
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int j = 0; j < kR / 16; ++j) {
            // 16 rows = two 8-row swizzle atoms = 2048 bytes per K step
            const uint64_t adesc = make_desc_mn_sw128(stage + (uint32_t)(2 * mt) * kCbBytes + j * 2048, kCbBytes);
            const uint64_t bdesc = make_desc_mn_sw128(stage + b_off + j * 2048, kCbBytes);
            umma_bf16(tmem_base + (uint32_t)(mt * NB), adesc, bdesc, idesc, (it | j) != 0 ? 1u : 0u);
          }
        }
        umma_commit(smem_u32(&bar_empty[s]));
        if (it == my_tiles - 1) umma_commit(smem_u32(&bar_done));
      }
      __syncwarp();
      if (++s == NS) { s = 0; ph ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWgMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

// dW[e] = sum over the row chunks of partial[c][e] in a FIXED order (deterministic): a block owns 64 elements, its
// four thread rows sum the chunks c = j, j+4, j+8, ... (independent loads, four in flight) and the four slice sums
// are combined as (s0 + s1) + (s2 + s3).
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int nchunks, int n,
                                                            float* __restrict__ dw) {
  __shared__ float sm[4][64];
  const int i = blockIdx.x * 64 + (threadIdx.x & 63), j = threadIdx.x >> 6;
  float s = 0.0f;
  if (i < n) {
    int c = j;
    for (; c + 12 < nchunks; c += 16) {
      const float a0 = __ldg(partial + (size_t)c * n + i), a1 = __ldg(partial + (size_t)(c + 4) * n + i);
      const float a2 = __ldg(partial + (size_t)(c + 8) * n + i), a3 = __ldg(partial + (size_t)(c + 12) * n + i);
      s += a0;
      s += a1;
      s += a2;
      s += a3;
    }
    for (; c < nchunks; c += 4) s += __ldg(partial + (size_t)c * n + i);
  }
  sm[j][threadIdx.x & 63] = s;
  __syncthreads();
  if (j == 0 && i < n) dw[i] = (sm[0][threadIdx.x] + sm[1][threadIdx.x]) + (sm[2][threadIdx.x] + sm[3][threadIdx.x]);
}

struct WgPlan {
  int ngroups, tpg, mt_alloc, NB, ntiles, nchunks, tpc, NS, tmem_cols;
  size_t stage_bytes, smem;
};

bool wg_supported(int Cin_p, int Cout) {
  return (Cin_p == 16 || Cin_p == 32 || Cin_p == 64 || Cin_p == 128) && (Cout == 16 || Cout == 32 || Cout == 64 || Cout == 128);
}

WgPlan wg_plan(int Cin_p, int Cout, int K, int no_max) {
  WgPlan pl;
  const int max_taps = Cin_p <= 64 ? 8 * (64 / Cin_p) : 4;      // 8 column blocks = 4 M-tiles per CTA
  pl.ngroups = cdiv(K, max_taps);
  pl.tpg = cdiv(K, pl.ngroups);
  pl.ngroups = cdiv(K, pl.tpg);
  const int ncb = Cin_p <= 64 ? cdiv(pl.tpg, 64 / Cin_p) : 2 * pl.tpg;
  pl.mt_alloc = (ncb + 1) / 2;
  pl.NB = Cout <= 64 ? 64 : 128;
  pl.stage_bytes = (size_t)(2 * pl.mt_alloc + pl.NB / 64) * kCbBytes;
  pl.NS = (int)((225 * 1024 - 1024) / pl.stage_bytes);
  if (pl.NS > kWgMaxStages) pl.NS = kWgMaxStages;
  pl.smem = (size_t)pl.NS * pl.stage_bytes + 1024;
  pl.tmem_cols = 32;
  while (pl.tmem_cols < pl.mt_alloc * pl.NB) pl.tmem_cols <<= 1;
  pl.ntiles = cdiv(no_max, kR);
  int chunks = sm_count() / pl.ngroups;
  if (chunks < 1) chunks = 1;
  if (chunks > pl.ntiles) chunks = pl.ntiles;
  if (chunks < 1) chunks = 1;
  pl.tpc = pl.ntiles > 0 ? cdiv(pl.ntiles, chunks) : 1;
  pl.nchunks = pl.ntiles > 0 ? cdiv(pl.ntiles, pl.tpc) : 1;
  return pl;
}

template <int CIN, int NB>
int launch_wg(const WgParams& p, const WgPlan& pl, cudaStream_t stream) {
  static thread_local DevOnce configured;   // per device: the attribute is a per-device property
  if (configured.first()) {
    COMB_CUDA(cudaFuncSetAttribute(spconv_wgrad_tc_kernel<CIN, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   227 * 1024 - 2048));
  }
  dim3 grid(pl.nchunks, pl.ngroups);
  spconv_wgrad_tc_kernel<CIN, NB><<<grid, kWgThreads, pl.smem, stream>>>(p);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

template <int CIN>
int dispatch_nb(const WgParams& p, const WgPlan& pl, cudaStream_t stream) {
  return pl.NB == 64 ? launch_wg<CIN, 64>(p, pl, stream) : launch_wg<CIN, 128>(p, pl, stream);
}

}  // namespace
}  // namespace comb

using namespace comb;

extern "C" size_t comb_spconv_wgrad_bf16_workspace_bytes(int Cin_p, int Cin, int Cout, int K, int no_max) {
  if (!wg_supported(Cin_p, Cout) || Cin < 1 || Cin > Cin_p || K < 1 || no_max < 0) return 0;
  const WgPlan pl = wg_plan(Cin_p, Cout, K, no_max);
  return (size_t)pl.nchunks * Cout * K * Cin * sizeof(float);
}

extern "C" int comb_spconv_wgrad_bf16(const void* in_feats, int Cin_p, int Cin, const void* dout, int Cout, int K,
                                      const int* nbr, int ld, int no_max, const int* no_dev, float* dweight,
                                      void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  COMB_CHECK_ARG(wg_supported(Cin_p, Cout), "comb_spconv_wgrad_bf16: Cin_p %d / Cout %d not in {16,32,64,128}", Cin_p, Cout);
  COMB_CHECK_ARG(Cin >= 1 && Cin <= Cin_p, "comb_spconv_wgrad_bf16: Cin %d > padded %d", Cin, Cin_p);
  COMB_CHECK_ARG(K >= 1 && ld >= no_max && no_max >= 0 && dweight, "comb_spconv_wgrad_bf16: bad arguments");
  const size_t n = (size_t)Cout * K * Cin;
  if (no_max == 0) {
    COMB_CUDA(cudaMemsetAsync(dweight, 0, n * sizeof(float), stream));
    return COMB_OK;
  }
  COMB_CHECK_ARG(in_feats && dout && nbr, "comb_spconv_wgrad_bf16: null pointer");
  COMB_CHECK_ARG(((uintptr_t)in_feats & 15) == 0 && ((uintptr_t)dout & 15) == 0, "comb_spconv_wgrad_bf16: unaligned operand");
  const WgPlan pl = wg_plan(Cin_p, Cout, K, no_max);
  COMB_CHECK_ARG(pl.NS >= 2 && pl.tmem_cols <= 512, "comb_spconv_wgrad_bf16: no valid plan");
  COMB_CHECK_ARG(workspace && workspace_bytes >= (size_t)pl.nchunks * n * sizeof(float),
                 "comb_spconv_wgrad_bf16: workspace of %zu bytes is too small", workspace_bytes);
  WgParams p;
  p.in = (const __nv_bfloat16*)in_feats;
  p.dout = (const __nv_bfloat16*)dout;
  p.nbr = nbr;
  p.ld = ld;
  p.no_max = no_max;
  p.no_dev = no_dev;
  p.K = K;
  p.Cin_real = Cin;
  p.Cout = Cout;
  p.taps_per_group = pl.tpg;
  p.mt_alloc = pl.mt_alloc;
  p.tiles_per_chunk = pl.tpc;
  p.ntiles = pl.ntiles;
  p.NS = pl.NS;
  p.tmem_cols = pl.tmem_cols;
  p.partial = (float*)workspace;
  int rc;
  switch (Cin_p) {
    case 16: rc = dispatch_nb<16>(p, pl, stream); break;
    case 32: rc = dispatch_nb<32>(p, pl, stream); break;
    case 64: rc = dispatch_nb<64>(p, pl, stream); break;
    default: rc = dispatch_nb<128>(p, pl, stream); break;
  }
  if (rc != COMB_OK) return rc;
  wgrad_reduce_kernel<<<cdiv((long long)n, 64), 256, 0, stream>>>(p.partial, pl.nchunks, (int)n, dweight);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}
