// a7 — sparse convolution forward on tcgen05, "row-per-thread" form: one gather thread owns one output row.
//
// Replaces the gather-GEMM-scatter of spconv's SubMConv3d / SparseConv3d forward
// (pcdet/models/backbones_3d/spconv_backbone.py:191-232), same output-stationary implicit GEMM as conv_ts.cu
//
//   D[128 rows, Cout] (TMEM, fp32)  +=  A_slab[128, up to 448] (TMEM, bf16)  x  B[Cout, up to 448]^T (smem, bf16)
//
// but built around what the r2 micro-benchmarks measured (scripts/micro/*.cu, profiles/r2_micro_*.txt):
//  * the tensor pipe retires an M=128, K=16 MMA in 128*N/256 cycles (9 at N=16, 16 at N=32, 32 at N=64, 64 at N=128)
//    when the issuing thread feeds it from uniform registers, and tcgen05.commit costs nothing on top;
//  * a barrier round costs the issuing thread 180-240 cycles of try_wait latency however deep the ring is, so the
//    thread that issues the MMAs — not the tensor pipe, not the gather — bounds a kernel that hand-shakes often:
//    conv_ts.cu pays a round per 4 chunk slots; with everything but its hand-shakes switched off (COMB_TS_ABLATE=29)
//    its 16x16 layer still takes 17 of its 29 us;
//  * what bounds the gather is bytes in flight (registers), not instructions.
// So the unit of hand-shake here is a SLAB of up to 7 K-chunks (448 K positions: a whole 128-row tile at Cin = 16,
// half a tile at Cin = 32), filled by ALL 16 gather warps at once, each with ONE batch of loads:
//  * thread t of gather warp w owns output row 32*(w%4) + t of the tile (= TMEM lane) and quarter w/4 of the slab's
//    K range: up to 7 pieces of 32 bytes (16 channels of one kernel tap).  It reads the neighbour index of each
//    piece's tap straight from the rulebook (coalesced: 32 consecutive rows of a tap are 128 contiguous bytes; no
//    index tile in shared memory, no index warp; the next slab's indices are fetched behind this slab's loads),
//    loads the pieces with LDG.256 (an absent neighbour reads a shared zero row instead of predicating and
//    zero-filling registers) and writes them to its lane with tcgen05.st.32x32b — K stays in its natural order
//    (tap-major, channel-minor), the weight image is the plain K-major SWIZZLE_128B image;
//  * one elected thread issues the slab's MMAs back to back (4 per chunk at constant descriptor offsets), commits the
//    slot once per slab and probes the next slab's barrier while it issues; accumulators are double buffered so the
//    epilogue of tile i overlaps the MMAs of tile i+1.
//
// TMEM map (512 columns): [0, 2*Cout) two accumulators, then 2 slots of 32 columns per chunk.
// Weights: the whole image is resident in shared memory (one bulk copy per CTA): everything up to 64x64x27 fits in
// 227 KB; the three wider shapes (64x128, 128x64, 128x128) stay on conv_ts.cu, which streams them (tr_supported()).
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
constexpr int kThreads = 21 * 32;      // registers are handed out per 4 warps: 21 warps cost as much as 24, 80 registers each
constexpr int kSlots = 2;
constexpr int kSmemMax = 227 * 1024;
constexpr int kMaxPieces = 7;          // pieces (8 registers each) one gather thread holds in flight

// 256 zero bytes: the "feature row" of an absent neighbour
__device__ __align__(256) uint4 g_zero_row[16];

template <int CIN, int COUT>
struct TrCfg {
  static constexpr int kPPT = CIN / 16;                        // 32-byte pieces per tap
  static constexpr int kLogPPT = CIN == 16 ? 0 : CIN == 32 ? 1 : CIN == 64 ? 2 : 3;
  static constexpr int kAccCols = 2 * COUT;                    // two accumulators
  static constexpr int kSlabChunksMax = (512 - kAccCols) / kSlots / 32 > kMaxPieces ? kMaxPieces : (512 - kAccCols) / kSlots / 32;
  static constexpr int kSlotCols = kSlabChunksMax * 32;
  static constexpr int kBBytes = COUT * 128;                   // one 64-wide K chunk of the weight image
  static __host__ __device__ int num_chunks(int K) { return CIN <= 64 ? (K * CIN + 63) / 64 : K * (CIN / 64); }
  static __host__ __device__ int slabs(int K) { return (num_chunks(K) + kSlabChunksMax - 1) / kSlabChunksMax; }
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
// tcgen05.st.32x32b.x8: thread t writes 8 consecutive 32-bit columns of lane (32*(warp%4) + t)
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
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
  pdl_launch_dependents();                           // the next kernel of the stream may set itself up behind our tail
  const int K = p.K;
  const int NCHT = Cfg::num_chunks(K);               // chunks of a tile's K range
  constexpr int NCH = Cfg::kSlabChunksMax;           // chunks per slab
  const int SPT = (NCHT + NCH - 1) / NCH;            // slabs per tile

  const uint32_t bars = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = bars + 1024u;

  if (tid == 0) {
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(bars + kBarFull + 8 * s, kGatherWarps);
      mbar_init(bars + kBarEmpty + 8 * s, 1);          // tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bars + kBarAccFull + 8 * a, 1);
      mbar_init(bars + kBarAccEmpty + 8 * a, 4);
    }
    mbar_init(bars + kBarB, 1);
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
  if (warp == kMmaWarp && elect_one_sync()) {
    // the weight image (packed once per weight version, long before this step): one bulk copy, resident for the CTA's life
    const uint32_t bytes = (uint32_t)NCHT * Cfg::kBBytes;
    mbar_arrive_expect_tx(bars + kBarB, bytes);
    bulk_copy_g2s(w_base, p.wpacked, bytes, bars + kBarB);
  }
  // ---- everything above touched only this CTA's own state and constant data; from here on the earlier kernels of the
  // stream (the producer of p.in / p.nbr / p.no_dev / p.residual) are complete
  pdl_wait();
  const int no = eff_n(p.no_max, p.no_dev);             // real row count (device side): capacities run 1.5-2x above it
  const int ntiles = (no + kBM - 1) / kBM;
  const int my_tiles = ntiles > (int)blockIdx.x ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (p.ablate & 64) {
    // timing experiment: set-up and tear-down only
  } else if (warp < kGatherWarps) {
    // ===================== gather: rulebook + rows -> registers -> TMEM =====================
    // Straight-line code: a slab always has NP pieces per thread (the slab size is a compile-time constant; pieces
    // past the tile's K range read the zero row and land in columns no MMA reads), so the NP index loads, the NP row
    // loads and the NP tcgen05.st of a slab are issued back to back with nothing but address arithmetic between them.
    // (r2 A/B: a rotating window — store piece i of slab n, immediately request piece i of slab n+1 into the freed
    // registers, indices two slabs ahead — keeps the same NP rows in flight and measured SLOWER, 27.1 vs 22.5 us on
    // the 16x16 level-1 layer: the window cannot hold more than one slab, and ptxas moves the loads behind the stores.
    // Also measured: TWO register sets of 3 pieces with slabs of 3 chunks in a ring of 4 slots — the rows of slab n+1
    // requested before slab n is stored, so a batch of loads is always in flight — 27.9 vs 22.5 us at 16x16, 48.9 vs
    // 45.6 at 32x32: each slab costs ~900 cycles of wait / tcgen05.wait::st / fence / arrive latencies whatever its
    // size, and three slabs per tile cost more than the hidden round trip saves.)
    constexpr int NP = Cfg::kSlabChunksMax;           // pieces per thread per slab = chunks per slab
    const int q = warp & 3, r = warp >> 2;            // row quarter, quarter of the slab's K range
    const char* in = reinterpret_cast<const char*>(p.in);
    const char* zrow = reinterpret_cast<const char*>(g_zero_row);
    const int ld = p.ld;
    // slab sl, this warp: global pieces (4 * sl + r) * NP + i; piece gp -> tap gp >> kLogPPT, channels 16 * (gp % kPPT)
    int idx[NP];
    // (32-bit element offsets into the rulebook — K * ld < 2^31 is checked on the host — and one base per slab: the
    // 64-bit tap * ld products of the first version were a third of the loop's instructions; same-box A/B 24.0 vs 25.2 us
    // at 16x16, 45.9 vs 48.8 at 32x32.  Cache hints measured the same way and rejected: L1::no_allocate on the rulebook
    // loads 23.9 vs 23.2 us, plus L1::evict_last on the row loads 23.8)
    const int* nbr_q = p.nbr + (q * 32 + lane);
    auto load_idx = [&](int ti, int sl) {
      const int tile_row = ((int)blockIdx.x + ti * (int)gridDim.x) * kBM;
      const int gp0 = (4 * sl + r) * NP;
      const int t0 = gp0 >> Cfg::kLogPPT, sub = gp0 & (Cfg::kPPT - 1);
      const bool live = tile_row + q * 32 + lane < no && ti < my_tiles;
      const int* src = nbr_q + (t0 * ld + tile_row);
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int dt = (sub + i) >> Cfg::kLogPPT;           // tap of piece i relative to t0 (compile-time at Cin = 16)
        idx[i] = -1;
        if (t0 + dt < K && live) idx[i] = __ldg(src + dt * ld);
      }
    };
    int ti = 0, sl = 0;
    load_idx(0, 0);
    uint32_t j = 0;
    const uint32_t t_base = ((uint32_t)(q * 32) << 16) + (uint32_t)Cfg::kAccCols + (uint32_t)(r * NP * 8);
#pragma unroll 1
    for (; ti < my_tiles; ++j) {
      const uint32_t slot = j & 1u;
      const int gp0 = (4 * sl + r) * NP;
      uint32_t v[NP][8];
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int ch = (gp0 + i) & (Cfg::kPPT - 1);
        const char* src = idx[i] >= 0 ? in + ((size_t)(uint32_t)idx[i] * (CIN * 2) + ch * 32) : zrow;
        ldg256(src, v[i]);
      }
      // next slab: its indices ride behind this slab's loads (past the end: all -1, no loads)
      if (++sl == SPT) {
        sl = 0;
        ++ti;
      }
      load_idx(ti, sl);
      mbar_wait(bars + kBarEmpty + 8 * slot, ((j >> 1) & 1u) ^ 1u);      // every lane polls: see mbar_wait_warp in tc_util.cuh
      tc_fence_after();
      const uint32_t t_dst = t_base + slot * Cfg::kSlotCols;
#pragma unroll
      for (int i = 0; i < NP; ++i) tmem_st_x8(t_dst + 8 * i, v[i]);
      tmem_st_wait();
      tc_fence_before();
      if (lane == 0) mbar_arrive(bars + kBarFull + 8 * slot);
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
    // Addresses and descriptors are derived from kernel parameters, blockIdx and constants only, so they live in
    // uniform registers; only the trip count (from the device-side row count) sits in an ordinary register.
    // instruction descriptor: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, K-major A/B, N>>3 [17,23), M>>4 [24,29)
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(COUT >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
    constexpr uint32_t kChunkStep = (uint32_t)Cfg::kBBytes >> 4;      // descriptor units between 64-wide K chunks
    const int AB = p.ablate;
    if (elect_one_sync()) {
      mbar_wait(bars + kBarB, 0);          // the weight image (requested before the grid dependency wait)
      uint32_t j = 0;
      const uint64_t desc0 = make_desc_sw128(w_base);
      const uint64_t desc_hi = desc0 & 0xFFFFFFFF00000000ull;
      bool ready = false;                  // the next slab's full barrier was seen complete by the look-ahead probe
#pragma unroll 1
      for (int ti = 0; ti < my_tiles; ++ti) {
        const uint32_t a = (uint32_t)ti & 1u;
        if (!(AB & 128)) mbar_wait(bars + kBarAccEmpty + 8 * a, (((uint32_t)ti >> 1) & 1u) ^ 1u);
        const uint32_t tmem_d = a * COUT;
        uint32_t dlo = (uint32_t)desc0, acc = 0;      // the weight descriptor walks through the tile's K range
        int c_left = NCHT;
#pragma unroll 1
        for (int sl = 0; sl < SPT; ++sl, ++j) {
          const uint32_t slot = j & 1u;
          const int nchs = c_left < NCH ? c_left : NCH;
          c_left -= nchs;
          if (!ready) mbar_wait(bars + kBarFull + 8 * slot, (j >> 1) & 1u);
          tc_fence_after();
          uint32_t tmem_a = (uint32_t)Cfg::kAccCols + slot * Cfg::kSlotCols;
          // look-ahead: the probe's latency (~150 cycles) hides behind the MMAs issued below
          const uint32_t j2 = j + 1;
          ready = !(AB & 32) && mbar_test(bars + kBarFull + 8 * (j2 & 1u), (j2 >> 1) & 1u);
          if (!(AB & 1)) {
#pragma unroll 1
            for (int c = 0; c < nchs; ++c) {
              umma_bf16_ts(tmem_d, tmem_a, desc_hi | dlo, idesc, acc);
              umma_bf16_ts(tmem_d, tmem_a + 8, desc_hi | (dlo + 2), idesc, 1u);
              umma_bf16_ts(tmem_d, tmem_a + 16, desc_hi | (dlo + 4), idesc, 1u);
              umma_bf16_ts(tmem_d, tmem_a + 24, desc_hi | (dlo + 6), idesc, 1u);
              acc = 1;
              tmem_a += 32;
              dlo += kChunkStep;
            }
          }
          umma_commit(bars + kBarEmpty + 8 * slot);
        }
        if (!(AB & 128)) umma_commit(bars + kBarAccFull + 8 * a);
      }
    }
    __syncwarp();
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
  COMB_CHECK_ARG((long long)p.K * p.ld < (1ll << 31), "comb_spconv_fwd_bf16: rulebook of %d x %d entries exceeds 32-bit offsets", p.K, p.ld);
  if (!Cfg::resident(p.K)) {
    set_error("comb_spconv_fwd_bf16: the row-per-thread kernel keeps the weight image in shared memory; %d x %d x %d does not fit", p.K, CIN, COUT);
    return COMB_EINVAL;
  }
  p.sc = Cfg::kSlabChunksMax;        // chunks per slab: fixed, the last slab of a tile may be partly padding
  {
    const char* ab = getenv("COMB_TS_ABLATE");     // read per launch: timing experiments flip it inside one process
    p.ablate = ab ? atoi(ab) : 0;
  }
  const size_t smem = 2048 + (size_t)Cfg::num_chunks(p.K) * Cfg::kBBytes;
  static thread_local DevOnce configured;   // per device: the attribute is a per-device property
  if (configured.first())
    COMB_CUDA(cudaFuncSetAttribute(spconv_tr_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
  const int ntiles = cdiv(p.no_max, kBM);
  const int grid = ntiles < sm_count() ? ntiles : sm_count();
  COMB_CUDA(launch_pdl(spconv_tr_kernel<CIN, COUT>, grid, kThreads, smem, stream, p));
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

bool tr_supported(int Cin_p, int Cout, int K) {
  const long long chunks = Cin_p <= 64 ? ((long long)K * Cin_p + 63) / 64 : (long long)K * (Cin_p / 64);
  return chunks * Cout * 128 + 2048 <= kSmemMax;
}

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
