// a7 — sparse convolution forward on tcgen05, "row-per-thread" form: one gather thread owns one output row.
//
// Replaces the gather-GEMM-scatter of spconv's SubMConv3d / SparseConv3d forward
// (pcdet/models/backbones_3d/spconv_backbone.py:191-232), same output-stationary implicit GEMM as conv_ts.cu
//
//   D[128 rows, Cout] (TMEM, fp32)  +=  A_slab[128, taps x Cin] (TMEM, bf16)  x  B[Cout, taps x Cin]^T (smem, bf16)
//
// but built around what the r2 micro-benchmarks measured (scripts/micro/commit_bench.cu, profiles/r2_commit_bench.txt):
//  * the tensor pipe retires an M=128, K=16 MMA in 128*N/256 cycles (9 at N=16, 16 at N=32, 32 at N=64, 64 at N=128)
//    when the issuing thread feeds it from uniform registers; tcgen05.commit costs nothing on top;
//  * ONE producer/consumer barrier round costs the issuing thread ~300 cycles whatever the ring depth.  conv_ts.cu pays
//    a round per 4 chunk slots with 16 gather warps arriving on every one of them: with everything but the hand-shakes
//    switched off (COMB_TS_ABLATE=29) its 16x16 layer still takes 22 of its 32 us.
// So here the unit of hand-shake is a SLAB: a run of whole kernel taps of one 128-row tile (14 taps at Cin = 16, 7 at
// Cin = 32, 3 at Cin = 64, 1 at Cin = 128) that one group of four gather warps fills on its own:
//  * thread t of gather warp w owns output row 32*(w%4) + t of the tile (= TMEM lane): it reads its neighbour indices
//    straight from the rulebook (coalesced: 32 consecutive rows of one tap are 128 contiguous bytes; no index tile in
//    shared memory, no index warp), loads each present neighbour's feature row with 32-byte loads (LDG.256; an absent
//    neighbour reads a shared zero row instead of predicating and zero-filling registers) and writes it to its lane
//    with tcgen05.st.32x32b — K stays in its natural order (tap-major, channel-minor), the weight image is the plain
//    K-major SWIZZLE_128B image;
//  * gather group g (warps 4g..4g+3) fills slot g of a ring of 4 slabs; the next slab's indices are fetched while the
//    current slab's rows are in flight;
//  * one elected thread issues the slab's MMAs back to back and commits once per slab (to the slot's empty barrier;
//    after the last slab of a tile also to the accumulator's full barrier); accumulators are double buffered so the
//    epilogue of tile i overlaps the MMAs of tile i+1.
//
// TMEM map (512 columns): [0, 2*Cout) two accumulators, then 4 slots of taps_per_slab * Cin/2 columns.
// Weights: resident in shared memory when the image fits (<= 216 KB: everything up to 64x64), else streamed per slab
// through a ring of cp.async.bulk stages.
#include <stdlib.h>
#include "common.cuh"
#include "conv_impl.cuh"
#include "tc_util.cuh"

namespace comb {
namespace {

using namespace tcu;

constexpr int kBM = 128;
constexpr int kGatherWarps = 16;
constexpr int kEpiWarp0 = 16;
constexpr int kMmaWarp = 20;
constexpr int kBWarp = 21;
constexpr int kThreads = 22 * 32;      // 704 threads: 88 registers each
constexpr int kSlots = 4;
constexpr int kSmemMax = 227 * 1024;
constexpr int kMaxBStages = 6;

// 256 zero bytes: the "feature row" of an absent neighbour
__device__ __align__(256) uint4 g_zero_row[16];

template <int CIN, int COUT>
struct TrCfg {
  static constexpr int kTapCols = CIN / 2;                     // TMEM columns of one tap of one row
  static constexpr int kAccCols = 2 * COUT;                    // two accumulators
  static constexpr int kTpsMax = (512 - kAccCols) / kSlots / kTapCols;   // taps of a slab that fit in a slot
  static constexpr int kSlotCols = kTpsMax * kTapCols;
  static constexpr int kBBytes = COUT * 128;                   // one 64-wide K chunk of the weight image
  static constexpr int kTapRegs = CIN / 2;                     // registers of one tap of one row
  static constexpr int kBatchRegs = CIN == 16 ? 40 : CIN == 32 ? 48 : 64;      // data registers of one batch of loads
  static constexpr int kBatch = kTapRegs >= 64 ? 1 : kBatchRegs / kTapRegs > kTpsMax ? kTpsMax : kBatchRegs / kTapRegs;   // taps loaded together
  static_assert(kTpsMax >= 1, "slab does not fit");
  static __host__ __device__ int num_chunks(int K) { return CIN <= 64 ? (K * CIN + 63) / 64 : K * (CIN / 64); }
  static __host__ __device__ int slabs(int K) { return (K + kTpsMax - 1) / kTpsMax; }
  // taps per slab, balanced over the slabs of a tile; with streamed weights a slab must cover whole 64-wide K chunks
  static __host__ __device__ int tps(int K, bool resident) {
    const int s = slabs(K);
    int t = (K + s - 1) / s;
    if (!resident) {
      const int gran = CIN >= 64 ? 1 : 64 / CIN;
      t = (t + gran - 1) / gran * gran;
      if (t > kTpsMax) t = kTpsMax / gran * gran;
    }
    return t;
  }
  static __host__ __device__ bool resident(int K) { return (size_t)num_chunks(K) * kBBytes + 2048 <= (size_t)kSmemMax; }
};

constexpr uint32_t kBarFull = 0, kBarEmpty = 32, kBarAccFull = 64, kBarAccEmpty = 80, kBarB = 96, kBarBFull = 128,
                   kBarBEmpty = 192, kTmemSlot = 256;

// 32 bytes of a feature row (read-only path)
__device__ __forceinline__ void ldg256(const void* p, uint32_t (&r)[8]) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}

// tcgen05.st.32x32b.xN: thread t writes N consecutive 32-bit columns of lane (32*(warp%4) + t)
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_st_row(uint32_t taddr, const uint32_t* v) {
  if constexpr (N == 8) {
    tmem_st_x8(taddr, v);
  } else {
#pragma unroll
    for (int i = 0; i < N; i += 16) tmem_st_x16(taddr + i, v + i);
  }
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(kThreads, 1) spconv_tr_kernel(ConvFwdArgs p) {
  using Cfg = TrCfg<CIN, COUT>;
  extern __shared__ uint8_t smem_raw[];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int no = eff_n(p.no_max, p.no_dev);             // real row count: only the gather and the epilogue look at it
  const int ntiles = (p.no_max + kBM - 1) / kBM;        // loop structure of every role: the launch bound (uniform)
  const int K = p.K;
  const bool bres = Cfg::resident(K);
  const int TPS = p.sc;                              // taps per slab (host: Cfg::tps)
  const int SPT = (K + TPS - 1) / TPS;               // slabs per tile
  const int NB = p.nb;                               // streamed-weight stages
  const int my_tiles = ntiles > (int)blockIdx.x ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int total = my_tiles * SPT;                  // slabs of this CTA

  const uint32_t bars = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = bars + 1024u;
  const uint32_t slab_wbytes = (uint32_t)(TPS * CIN / 64) * Cfg::kBBytes;   // streamed: weight bytes of a full slab

  if (tid == 0) {
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(bars + kBarFull + 8 * s, 4);           // the four warps of the slot's gather group
      mbar_init(bars + kBarEmpty + 8 * s, 1);          // tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bars + kBarAccFull + 8 * a, 1);
      mbar_init(bars + kBarAccEmpty + 8 * a, 4);
    }
    mbar_init(bars + kBarB, 1);
    for (int s = 0; s < 8; ++s) {
      mbar_init(bars + kBarBFull + 8 * s, 1);
      mbar_init(bars + kBarBEmpty + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars + kTmemSlot), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  {
    // The CTA owns all 512 columns of its SM's tensor memory, so the allocation starts at lane 0, column 0; addresses
    // are then plain constants (uniform registers in the issue loop).  Anything else is a broken assumption: stop.
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(bars + kTmemSlot) : "memory");
    if (tmem_base != 0) __trap();
  }

  if (p.ablate & 64) {
    // timing experiment: set-up and tear-down only
  } else if (warp < kGatherWarps) {
    // ===================== gather: rulebook + rows -> registers -> TMEM =====================
    const int q = warp & 3, g = warp >> 2;
    const uint32_t t_slot = ((uint32_t)(q * 32) << 16) + (uint32_t)(Cfg::kAccCols + g * Cfg::kSlotCols);
    const char* in = reinterpret_cast<const char*>(p.in);
    const char* zrow = reinterpret_cast<const char*>(g_zero_row);
    const int ld = p.ld;
    constexpr int TPSMAX = Cfg::kTpsMax, TB = Cfg::kBatch;
    constexpr int LASTB0 = (TPSMAX - 1) / TB * TB;                // first tap of the last batch
    const int AB = p.ablate;             // COMB_TS_ABLATE: timing experiments only (results are garbage)
    const uint32_t bar_empty = bars + kBarEmpty + 8 * g, bar_full = bars + kBarFull + 8 * g;

    // position of this group's current slab (ti, sl) and of the one whose indices are being fetched (nti, nsl):
    // the group takes every fourth slab of the CTA's sequence, so the position advances by 4 slabs, no division
    auto advance = [&](int& ti, int& sl) {
      sl += kSlots;
      while (sl >= SPT) {
        sl -= SPT;
        ++ti;
      }
    };
    // indices of slab (ti, sl) for this thread's row: nbr[k0 + t][tile * 128 + row]; -1 = absent / past the end
    int idx[TPSMAX];
    auto load_idx = [&](int ti, int sl) {
      const int row = ((int)blockIdx.x + ti * (int)gridDim.x) * kBM + q * 32 + lane;
      const int k0 = sl * TPS;
      const int nt = K - k0 < TPS ? K - k0 : TPS;
      const int* src = p.nbr + (size_t)k0 * ld + row;
      const bool live = row < no && ti < my_tiles && !(AB & 16);
#pragma unroll
      for (int t = 0; t < TPSMAX; ++t) {
        idx[t] = -1;
        if (t < nt && live) idx[t] = __ldg(src + (size_t)t * ld);
      }
    };
    if (AB & 256) {
      // timing experiment: the bare hand-shake of a gather group
      uint32_t use = 0;
      for (int j = g; j < total; j += kSlots, ++use) {
        mbar_wait(bar_empty, (use & 1u) ^ 1u);
        if (lane == 0) mbar_arrive(bar_full);
      }
    } else {
      int ti = 0, sl = g;
      while (sl >= SPT) {
        sl -= SPT;
        ++ti;
      }
      int nti = ti, nsl = sl;
      load_idx(ti, sl);
      uint32_t par = 1;                   // parity to wait for on the slot's empty barrier
#pragma unroll 1
      for (; ti < my_tiles; advance(ti, sl), par ^= 1u) {
        const int k0 = sl * TPS;
        const int nt = K - k0 < TPS ? K - k0 : TPS;       // taps of this slab
        advance(nti, nsl);
#pragma unroll
        for (int b0 = 0; b0 < TPSMAX; b0 += TB) {
          uint32_t v[TB][Cfg::kTapRegs];
#pragma unroll
          for (int t = 0; t < TB; ++t) {
            if (b0 + t < TPSMAX && b0 + t < nt && !(AB & 4)) {
              const char* src = idx[b0 + t] >= 0 ? in + (size_t)(uint32_t)idx[b0 + t] * (CIN * 2) : zrow;
#pragma unroll
              for (int i = 0; i < CIN / 16; ++i) ldg256(src + 32 * i, *reinterpret_cast<uint32_t(*)[8]>(&v[t][8 * i]));
            }
          }
          // the next slab's indices ride behind the last batch of loads of this one (past the end: all -1, no loads)
          if (b0 == LASTB0) load_idx(nti, nsl);
          if (b0 == 0) {
            mbar_wait(bar_empty, par);
            tc_fence_after();
          }
#pragma unroll
          for (int t = 0; t < TB; ++t)
            if (b0 + t < TPSMAX && b0 + t < nt && !(AB & 6)) tmem_st_row<Cfg::kTapRegs>(t_slot + (uint32_t)((b0 + t) * Cfg::kTapCols), v[t]);
        }
        tmem_st_wait();
        tc_fence_before();
        if (lane == 0) mbar_arrive(bar_full);
      }
    }
  } else if (warp < kEpiWarp0 + 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;
    for (int ti = 0; ti < my_tiles && !(p.ablate & 128); ++ti) {
      const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
      const int a = ti & 1;
      const int row = tile * kBM + q * 32 + lane;
      const bool live = row < no;
      uint4 rv[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
      const uint4* rp = reinterpret_cast<const uint4*>(p.residual + (size_t)row * COUT);
      if ((p.epi & COMB_EPI_RESIDUAL) && live) {   // issued before the accumulator is ready
        rv[0] = __ldg(rp);
        rv[1] = __ldg(rp + 1);
      }
      mbar_wait_sleep(bars + kBarAccFull + 8 * a, (uint32_t)(ti >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = ((uint32_t)(q * 32) << 16) + a * COUT;
      if (!(p.ablate & 8))
#pragma unroll
      for (int c0 = 0; c0 < COUT; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        uint4 rn[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
        if (c0 + 16 < COUT && (p.epi & COMB_EPI_RESIDUAL) && live) {
          rn[0] = __ldg(rp + (c0 + 16) / 8);
          rn[1] = __ldg(rp + (c0 + 16) / 8 + 1);
        }
        tmem_ld_wait();
        if (live) {
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
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rv[h]);
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
        rv[0] = rn[0];
        rv[1] = rn[1];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + kBarAccEmpty + 8 * a);
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    // Everything this thread touches is derived from kernel parameters, blockIdx and constants (the tile count comes
    // from the launch bound no_max, not from the device-side row count: a tile past the real end is multiplied like
    // any other and simply not stored), so the whole loop lives in uniform registers.
    // instruction descriptor: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, K-major A/B, N>>3 [17,23), M>>4 [24,29)
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(COUT >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
    constexpr uint32_t kChunkStep = (uint32_t)Cfg::kBBytes >> 4;      // descriptor units between 64-wide K chunks
    const int AB = p.ablate;
    if (elect_one_sync()) {
      if (bres) mbar_wait(bars + kBarB, 0);
      uint32_t j = 0, bs = 0, bph = 0;
      const uint64_t desc0 = make_desc_sw128(w_base);
#pragma unroll 1
      for (int ti = 0; ti < my_tiles; ++ti) {
        const uint32_t a = (uint32_t)ti & 1u;
        if (!(AB & 128)) mbar_wait(bars + kBarAccEmpty + 8 * a, (((uint32_t)ti >> 1) & 1u) ^ 1u);
        const uint32_t tmem_d = a * COUT;
        uint32_t dlo = (uint32_t)desc0, kk = 0, acc = 0;      // weight descriptor walks through the tile's K range
        int k_left = K;
#pragma unroll 1
        for (int sl = 0; sl < SPT; ++sl, ++j) {
          const uint32_t slot = j & (kSlots - 1);
          const int nt = k_left < TPS ? k_left : TPS;
          k_left -= nt;
          const int nsteps = nt * (CIN / 16);                       // K = 16 per MMA
          if (!bres) {
            mbar_wait(bars + kBarBFull + 8 * bs, bph);
            dlo = (uint32_t)desc0 + bs * (slab_wbytes >> 4);
            kk = 0;
          }
          mbar_wait(bars + kBarFull + 8 * slot, (j >> 2) & 1u);
          tc_fence_after();
          uint32_t tmem_a = (uint32_t)Cfg::kAccCols + slot * Cfg::kSlotCols;
          if (!(AB & 1)) {
#pragma unroll 4
            for (int s = 0; s < nsteps; ++s) {
              umma_bf16_ts(tmem_d, tmem_a, (desc0 & 0xFFFFFFFF00000000ull) | dlo, idesc, acc);
              acc = 1;
              tmem_a += 8;
              dlo += 2;
              if (++kk == 4) {
                kk = 0;
                dlo += kChunkStep - 8;
              }
            }
          }
          umma_commit(bars + kBarEmpty + 8 * slot);
          if (!bres) {
            umma_commit(bars + kBarBEmpty + 8 * bs);
            if (++bs == (uint32_t)NB) { bs = 0; bph ^= 1u; }
          }
        }
        if (!(AB & 128)) umma_commit(bars + kBarAccFull + 8 * a);
      }
    }
    __syncwarp();
  } else if (warp == kBWarp && lane == 0) {
    // ===================== weights =====================
    if (bres) {
      const uint32_t bytes = (uint32_t)Cfg::num_chunks(K) * Cfg::kBBytes;
      mbar_arrive_expect_tx(bars + kBarB, bytes);
      bulk_copy_g2s(w_base, p.wpacked, bytes, bars + kBarB);
    } else {
      const int nchunks = Cfg::num_chunks(K), cps = TPS * CIN / 64;      // chunks per full slab
      int bs = 0;
      uint32_t bph = 0;
      for (int ti = 0; ti < my_tiles; ++ti)
        for (int sl = 0; sl < SPT; ++sl) {
          const int c0 = sl * cps;
          const uint32_t bytes = (uint32_t)(nchunks - c0 < cps ? nchunks - c0 : cps) * Cfg::kBBytes;
          mbar_wait(bars + kBarBEmpty + 8 * bs, bph ^ 1u);
          mbar_arrive_expect_tx(bars + kBarBFull + 8 * bs, bytes);
          bulk_copy_g2s(w_base + (uint32_t)bs * slab_wbytes, p.wpacked + (size_t)c0 * Cfg::kBBytes, bytes, bars + kBarBFull + 8 * bs);
          if (++bs == NB) { bs = 0; bph ^= 1u; }
        }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(0u), "r"(512u) : "memory");
  }
}

template <int CIN, int COUT>
int launch_tr(const ConvFwdArgs& p_in, cudaStream_t stream) {
  using Cfg = TrCfg<CIN, COUT>;
  ConvFwdArgs p = p_in;
  const bool bres = Cfg::resident(p.K);
  p.sc = Cfg::tps(p.K, bres);
  {
    const char* ab = getenv("COMB_TS_ABLATE");     // read per launch: timing experiments flip it inside one process
    p.ablate = ab ? atoi(ab) : 0;
  }
  size_t smem = 2048;
  if (bres) {
    smem += (size_t)Cfg::num_chunks(p.K) * Cfg::kBBytes;
    p.nb = 0;
  } else {
    const size_t stage = (size_t)(p.sc * CIN / 64) * Cfg::kBBytes;
    int nb = (int)((kSmemMax - 2048) / stage);
    if (nb > kMaxBStages) nb = kMaxBStages;
    if (nb < 2 || p.sc < 1) {
      set_error("comb_spconv_fwd_bf16: weight stage of %zu bytes does not fit twice in shared memory", stage);
      return COMB_EINVAL;
    }
    p.nb = nb;
    smem += (size_t)nb * stage;
  }
  static thread_local DevOnce configured;   // per device: the attribute is a per-device property
  if (configured.first())
    COMB_CUDA(cudaFuncSetAttribute(spconv_tr_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
  const int ntiles = cdiv(p.no_max, kBM);
  const int grid = ntiles < sm_count() ? ntiles : sm_count();
  spconv_tr_kernel<CIN, COUT><<<grid, kThreads, smem, stream>>>(p);
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

template <int CIN>
int dispatch_cout(int Cout, const ConvFwdArgs& p, cudaStream_t stream) {
  switch (Cout) {
    case 16: return launch_tr<CIN, 16>(p, stream);
    case 32: return launch_tr<CIN, 32>(p, stream);
    case 64: return launch_tr<CIN, 64>(p, stream);
    case 128: return launch_tr<CIN, 128>(p, stream);
  }
  set_error("comb_spconv_fwd_bf16: Cout %d not in {16,32,64,128}", Cout);
  return COMB_EINVAL;
}

}  // namespace

int tr_fwd_bf16(const ConvFwdArgs& p, int Cin_p, int Cout, cudaStream_t stream) {
  switch (Cin_p) {
    case 16: return dispatch_cout<16>(Cout, p, stream);
    case 32: return dispatch_cout<32>(Cout, p, stream);
    case 64: return dispatch_cout<64>(Cout, p, stream);
    case 128: return dispatch_cout<128>(Cout, p, stream);
  }
  set_error("comb_spconv_fwd_bf16: Cin_p %d not in {16,32,64,128}", Cin_p);
  return COMB_EINVAL;
}

}  // namespace comb
