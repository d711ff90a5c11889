// a7 — sparse convolution forward on tcgen05 with the gathered A operand staged in TENSOR MEMORY (TS form).
//
// Replaces the gather-GEMM-scatter of spconv's SubMConv3d / SparseConv3d forward
// (pcdet/models/backbones_3d/spconv_backbone.py:191-232).  Same output-stationary implicit GEMM as conv_tc.cu
//
//   D[128 rows, Cout] (TMEM, fp32)  +=  A_chunk[128, 64] (TMEM, bf16)  x  B_chunk[Cout, 64]^T (smem, bf16)
//
// but the gathered rows never touch shared memory: the gather warps load neighbour rows from global/L2 into
// registers (LDG.128, only for neighbours that exist) and write them straight into TMEM with tcgen05.st, and
// the MMA reads A from TMEM (tcgen05.mma [d], [a], b-desc).  r1 trace of the SS kernel: every tcgen05.mma
// (M=128,K=16) cost >= 57 cycles at issue whatever N was, because the 4 KB A operand was fetched through the
// shared-memory port that the 16 KB/chunk of gather stores also went through; with A in TMEM the instruction
// floor is 128*N/256 cycles and the shared-memory port only carries the weights.
//
// TMEM map (512 columns): [0,256) accumulators (256/Cout buffers, ring), [256,512) eight A chunk slots of 32
// columns (= 64 bf16 per row).  A row m of a tile lives in TMEM lane m; a warp may only touch lanes
// 32*(warp%4)..+31, so gather warp w serves row quarter w%4 and chunk slots G with G%4 == w/4.
//
// K permutation.  tcgen05.st.16x256b hands thread t of a warp columns 8g+2(t%4)+{0,1} of rows t/4 and t/4+8
// (g = repeat).  To keep the global loads 16 bytes per lane and 64 contiguous bytes per 4 lanes, the thread's
// 16-byte piece p = 4*half + (t%4) of the 128-byte chunk row goes to column groups 2*half and 2*half+1.  So the
// source element e = 8p + w of a chunk sits at K position  kappa(e) = 32*half + 16*(w/4) + 4*(t%4) + (w%4);
// the weight image is packed with the same permutation (ts_pack_weight), so the contraction is unchanged.
//
// Streamed weights (images that do not fit in shared memory: 64x64, 64x128, 128x128) are consumed by TWO row
// tiles per stage (a CTA walks 256-row super-tiles), which halves the L2->SM weight traffic that bounded those
// layers (one 128-row tile re-read the full 216-864 KB image).
//
// Warp roles (768 threads, one CTA per SM, persistent over super-tiles):
//   warps 0-15  gather    : index LDS -> predicated LDG.128 -> tcgen05.st -> wait::st -> arrive on the slot
//   warps 16-19 epilogue  : tcgen05.ld -> bias / BN affine / residual / ReLU -> 16-byte stores
//   warp  20    MMA       : elected lane issues 4 x tcgen05.mma (K=16) per chunk, commits slot / weight stage / tile
//   warp  21    weights   : resident image (one bulk copy) or a ring of per-chunk stages (cp.async.bulk)
//   warp  22    indices   : neighbour-index tiles (27 x 128 ints), 2*T buffers, 16-byte cp.async
#include <stdlib.h>
#include "common.cuh"
#include "conv_impl.cuh"
#include "tc_util.cuh"

namespace comb {
namespace {

using namespace tcu;

constexpr int kBM = 128;
constexpr int kChunkK = 64;
constexpr int kGatherWarps = 16;
constexpr int kEpiWarp0 = 16;
constexpr int kMmaWarp = 20;       // issues the MMAs of row tile 0 of every pass (and owns the TMEM allocation)
constexpr int kMmaWarp2 = 23;      // issues the MMAs of row tile 1
constexpr int kBWarp = 21;
constexpr int kIdxWarp = 22;
constexpr int kThreads = 24 * 32;
constexpr int kMaxK = 32;
constexpr int kMaxBStages = 8;
constexpr int kSmemBudget = 225 * 1024;
constexpr int kBResidentMax = 160 * 1024;
constexpr bool kWideDefault = true;    // COMB_TS_WIDE unset: one 32-byte load per row (0: the two 16-byte pieces of r1)

template <int CIN, int COUT>
struct TsCfg {
  static constexpr int kOffPerChunk = CIN <= 64 ? 64 / CIN : 1;
  static constexpr int kChunksPerOff = CIN <= 64 ? 1 : CIN / 64;
  static constexpr int kBBytes = COUT * 128;
  static __host__ __device__ int num_chunks(int K) {
    return CIN <= 64 ? (K + kOffPerChunk - 1) / kOffPerChunk : K * kChunksPerOff;
  }
  static __host__ __device__ int k_pad(int K) { return CIN <= 64 ? num_chunks(K) * kOffPerChunk : K; }
  static __host__ __device__ bool b_resident(int K) { return num_chunks(K) * kBBytes <= kBResidentMax; }
  static __host__ __device__ int tiles_per_pass(int) { return 2; }   // T: every pass walks two row tiles, one per MMA-issuing thread
  // TMEM: accumulator ring in [0, d_cols), A stages of 128 columns (4 chunks) behind it
  static __host__ __device__ int d_cols(int) { return COUT >= 128 ? 256 : 128; }   // 2 x 128, 2 x 64, 4 x 32, 4 x 16
  static __host__ __device__ int n_acc(int K) {
    const int n = d_cols(K) / COUT;
    return n > 4 ? 4 : n;
  }
  static __host__ __device__ int a_stages(int K) { return (512 - d_cols(K)) / 128; }   // 2 or 3
  // index-tile ring: 8 buffers (4 passes ahead of the gather) when shared memory allows, else 4.  (r1 ncu source
  // view: with 2 buffers the gather warps' most executed instructions were the spin on the index barrier.)
  static __host__ __device__ int n_idx(int K) {
    const int big = 2048 + 8 * k_pad(K) * kBM * 4;
    if (b_resident(K)) return num_chunks(K) * kBBytes + big <= kSmemBudget ? 8 : 4;
    return (kSmemBudget - big) / (2 * kBBytes) >= 4 ? 8 : 4;
  }
  static __host__ __device__ int idx_bytes(int K) { return n_idx(K) * k_pad(K) * kBM * 4; }
  static __host__ __device__ int b_stages(int K) {   // streamed weights: ring of 2-chunk stages
    int n = (kSmemBudget - 2048 - idx_bytes(K)) / (2 * kBBytes);
    return n > kMaxBStages ? kMaxBStages : n;
  }
  static size_t smem_bytes(int K) {
    return 2048 + (size_t)idx_bytes(K) +
           (b_resident(K) ? (size_t)num_chunks(K) * kBBytes : (size_t)b_stages(K) * 2 * kBBytes);
  }
};

// barrier block at the start of the (1024-byte aligned) dynamic shared memory: 32-bit shared addresses are plain
// arithmetic on one base register (taking the address of a static __shared__ array costs an S2R + LEA every time)
constexpr uint32_t kBarFull = 0, kBarEmpty = 32, kBarBFull = 64, kBarBEmpty = 128, kBarTFull = 192, kBarTEmpty = 256,
                   kBarIdx = 320, kBarIdxFree = 384, kBarB = 448, kTmemSlot = 456;

__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint4& a0, const uint4& b0, const uint4& a1,
                                                   const uint4& b1) {
  // a = row t/4, b = row t/4 + 8; 0/1 = first / second 16-byte piece of the thread (see the K permutation above)
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(a0.x), "r"(a0.y), "r"(b0.x), "r"(b0.y), "r"(a0.z), "r"(a0.w), "r"(b0.z), "r"(b0.w), "r"(a1.x), "r"(a1.y),
      "r"(b1.x), "r"(b1.y), "r"(a1.z), "r"(a1.w), "r"(b1.z), "r"(b1.w)
      : "memory");
}
// trace layout: dbg[G*16 + slot] for chunk slot G < 512 of CTA 0.  slots: 0 mma saw the
// chunk full, 1 mma issued + committed, 2 gather (row quarter 0 of the owning group) start, 3 gather saw the slot
// empty (loads issued before), 4 gather stored + arrived, 5 epilogue finished tile G, 6 mma loop top, 7 gather saw idx
constexpr int kDbgChunks = 512;
__device__ __forceinline__ void dbg_stamp(long long* dbg, int G, int slot) {
  if (dbg != nullptr && blockIdx.x == 0 && G < kDbgChunks) {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
    dbg[G * 16 + slot] = t;
  }
}
__device__ __forceinline__ void dbg_cta_time(long long* dbg, int slot) {   // per-CTA wall clock (ns), slots 0..3
  if (dbg != nullptr && blockIdx.x < 256) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    dbg[kDbgChunks * 16 + blockIdx.x * 4 + slot] = t;
  }
}

// wait of a warp that is not on the MMA issue path: PIPE = 2 (COMB_TS_PIPE=2) polls with plain try_wait (the
// instruction suspends the warp in hardware for a bounded time), everything else backs off with nanosleep
template <int PIPE>
__device__ __forceinline__ void wait_bg(uint32_t bar, uint32_t parity) {
  if (PIPE == 2) mbar_wait(bar, parity);
  else mbar_wait_sleep(bar, parity);
}

// 256 zero bytes: the "feature row" of an absent neighbour in the wide-load gather (PIPE == 3)
__device__ __align__(256) uint4 g_ts_zero_row[16];

// 32 bytes of a feature row (read-only path), as the two 16-byte halves the tcgen05.st fragment wants
__device__ __forceinline__ void ldg256(const void* p, uint4& lo, uint4& hi) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
               : "l"(p));
}

__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// PIPE: software-pipelined gather (two batches of loads in flight per gather warp, register budget moved to the gather
// warps with setmaxnreg); PIPE = false is the r1 loop (one batch in flight), kept for A/B runs (COMB_TS_PIPE=0).
template <int CIN, int COUT, bool TRACE, int PIPE>
__global__ void __launch_bounds__(kThreads, 1) spconv_ts_kernel(ConvFwdArgs p) {
  using Cfg = TsCfg<CIN, COUT>;
  extern __shared__ uint8_t smem_raw[];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();                           // the next kernel of the stream may set itself up behind our tail
  if (TRACE && tid == 0) dbg_cta_time(p.dbg, 0);
  const int K = p.K;
  const int nchunks = Cfg::num_chunks(K);
  const int kpad = Cfg::k_pad(K);
  const bool bres = Cfg::b_resident(K);
  constexpr int T = 2, lT = 1;                       // row tiles per pass (256 rows): one per MMA-issuing thread
  const int NI = p.ni;                               // index-tile buffers (8 or 4), chosen on the host (launch_ts)
  const int lNI = NI == 8 ? 3 : 2;
  const int NB = p.nb;                               // streamed-weight stages, chosen on the host
  const int ND = Cfg::n_acc(K);                      // accumulator ring (power of two, >= T)
  // A stage = SC chunks of EACH of the two row tiles of the pass (2*SC chunk slots of 32 TMEM columns).  r2: with
  // every neighbour ABSENT — no global load at all — the 16x16 layer still takes 27.9 us against 28.6 us on the real
  // rulebook (scripts/conv_floor.py): the hand-shakes of the pipeline, not the gather loads, bound the kernel.  Larger
  // stages (SC = 3: 6 slots, 2 stages) were measured and did not help; SC = 2 (4 slots, 3 stages at Cout <= 64) is the
  // default.
  const int SC = p.sc;
  const uint32_t stage_cols = (uint32_t)(2 * SC * 32);
  const int NS = (512 - Cfg::d_cols(K)) / (int)stage_cols;     // A stages
  const uint32_t colA = (uint32_t)Cfg::d_cols(K);
  const int nst = (nchunks + SC - 1) / SC;           // stages per super-tile (the last one may be partial)
  // Single-tile passes (the split last wave below, or an odd tile at the end) used to leave the second half of the
  // pipeline idle: 8 of the 16 gather warps, one of the two MMA threads.  At level 4 of the bench frames (101 row tiles
  // on 148 SMs) EVERY pass is such a pass.  With p.split the idle half takes the UPPER half of the tile's stages
  // (K split inside the CTA): gather groups 1 and 3 read the same index tile, MMA thread 1 accumulates stages
  // [nst_p, nst) into the second accumulator, the streamed-weight ring carries the two stage sequences interleaved, and
  // the epilogue adds the two accumulators (fixed order: deterministic).
  // r2 A/B on the bench frames, same box, two runs each (profiles/r2_bench_split{0,1}.json): 64x64 67.9 -> 65.1 us per
  // launch (its last wave is made of single-tile passes), step 1.0767 -> 1.0700 ms.  Level 4 does NOT gain: its dense-K
  // row count is 246 tiles, i.e. 123 full two-tile passes, not 101 single ones as first assumed from the FLOP count.
  const bool split_ok = PIPE != 1 && p.split != 0 && nst >= 2;

  // shared memory map (1024-byte aligned): [barriers 1 KB] [weights: resident image | NB streamed 2-chunk stages]
  // [index tiles NI x kpad x 128]
  const uint32_t bars = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = bars + 1024u;
  const uint32_t idx_base = w_base + (bres ? nchunks * Cfg::kBBytes : NB * SC * Cfg::kBBytes);
  const uint32_t idx_buf_bytes = (uint32_t)kpad * kBM * 4;

  if (tid == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(bars + kBarFull + 8 * s, kGatherWarps);
      mbar_init(bars + kBarEmpty + 8 * s, 2);          // both MMA threads commit
    }
    for (int s = 0; s < 8; ++s) {
      mbar_init(bars + kBarBFull + 8 * s, 1);
      mbar_init(bars + kBarBEmpty + 8 * s, 2);
      mbar_init(bars + kBarTFull + 8 * s, 1);
      mbar_init(bars + kBarTEmpty + 8 * s, 4);
      mbar_init(bars + kBarIdx + 8 * s, 32);
      mbar_init(bars + kBarIdxFree + 8 * s, kGatherWarps);
    }
    mbar_init(bars + kBarB, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars + kTmemSlot), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // index-tile rows of the padding offsets (k in [K, kpad)) stay -1 for the whole kernel
  for (int e = tid; e < NI * (kpad - K) * kBM; e += kThreads) {
    const int buf = e / ((kpad - K) * kBM), r = e - buf * (kpad - K) * kBM;
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(idx_base + buf * idx_buf_bytes + (K * kBM + r) * 4), "r"(-1) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(bars + kTmemSlot) : "memory");
  if (TRACE && tid == 0) dbg_cta_time(p.dbg, 1);

  // ---- everything above touched only this CTA's own state; from here on the earlier kernels of the stream (the
  // producers of p.in / p.nbr / p.no_dev / p.residual) are complete (programmatic dependent launch, tc_util.cuh)
  pdl_wait();
  const int no = eff_n(p.no_max, p.no_dev);
  const int ntiles = (no + kBM - 1) / kBM;
  // Passes.  A pass walks T row tiles; with T = 2 the last wave of passes is split into single-tile passes when
  // that shortens it (r <= grid/2 leftover super-tiles become 2r half passes on 2r CTAs: the absent second tile
  // is neither gathered nor multiplied).
  // Two assignments of super-tiles (pairs of row tiles) to CTAs:
  //  * strided (default): CTA b owns b, b+grid, ...;
  //  * blocked (COMB_TS_BLOCKED=1): CTA b owns a CONTIGUOUS run of super-tiles — with rows in key order the tiles of one
  //    CTA are spatial neighbours in the same z-plane, so the rows a tile gathers for its dy = +1 taps are the rows the
  //    next tile gathers for dy = 0, -1 and could still be in the SM's L1.  r2 A/B on the bench frames: 0.94-0.96 ms
  //    per step against 0.93 strided, with 8 or 4 index buffers and 8, 4 or 2 weight stages (i.e. 0-100 KB of L1):
  //    cross-tile L1 reuse does not bound the gather.  Kept as a switch.
  int nsuper = (ntiles + T - 1) >> lT, nfull = nsuper;
  const bool blocked = p.blocked != 0;
  {
    // the last wave of passes is split into single-tile passes when that shortens it (r <= grid/2 leftover super-tiles
    // become 2r half passes on 2r CTAs: the absent second tile is neither gathered nor multiplied)
    const int r = nsuper % (int)gridDim.x;
    if (T == 2 && r > 0 && 2 * r <= (int)gridDim.x) {
      nfull = nsuper - r;
      nsuper = nfull + (ntiles - 2 * nfull);
    }
  }
  // blocked: the nfull full super-tiles are dealt out in contiguous runs (base or base+1 per CTA), the half passes of
  // the split last wave go one each to the first CTAs
  const int base = nfull / (int)gridDim.x, rem = nfull % (int)gridDim.x;
  const int b_lo = (int)blockIdx.x * base + ((int)blockIdx.x < rem ? (int)blockIdx.x : rem);
  const int b_cnt = base + ((int)blockIdx.x < rem ? 1 : 0);
  // super-tile of pass `it` of this CTA (>= nfull: a single-tile pass), and whether its second row tile is absent
  auto sidx_of = [&](int it) -> int {
    if (!blocked) return (int)blockIdx.x + it * (int)gridDim.x;
    return it < b_cnt ? b_lo + it : nfull + (int)blockIdx.x;
  };
  auto absent_of = [&](int it) -> bool { return sidx_of(it) >= nfull || 2 * sidx_of(it) + 1 >= ntiles; };
  // row tile of tile-sequence number n of this CTA (a tile index past the end reads as "no rows")
  auto tile_of = [&](int n) -> int {
    const int sidx = sidx_of(n >> lT);
    if (sidx < nfull) return (sidx << lT) + (n & (T - 1));
    return (n & (T - 1)) == 0 ? (nfull << lT) + (sidx - nfull) : 0x00FFFFFF;
  };
  int my_super;
  if (blocked) {
    my_super = b_cnt + ((int)blockIdx.x < nsuper - nfull ? 1 : 0);
  } else {
    my_super = nsuper > (int)blockIdx.x ? (nsuper - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  }
  if (PIPE == 1) {
    // Register budget: the CTA owns 768 x 80 = 61440 registers for its whole life (setmaxnreg only moves registers
    // inside the CTA's pool: asking for more than the other warpgroups release never completes — r2 lesson, a 104-register
    // request deadlocked).  The gather warps (warpgroups 0-3) hold two batches of 32 registers of loads in flight; the
    // epilogue warpgroup gives up 16 and the MMA / weight / index warpgroup 48 registers per thread:
    // 512 x 96 + 128 x 64 + 128 x 32 = 61440.
    if (warp < kGatherWarps) asm volatile("setmaxnreg.inc.sync.aligned.u32 96;" ::: "memory");
    else if (warp < kEpiWarp0 + 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;" ::: "memory");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 32;" ::: "memory");
  }

  if (PIPE == 1 && warp < kGatherWarps) {
    // ===================== gather, software pipelined =====================
    // Same work split as below (group g = chunk sub-slot g of every stage, warp q of the group = row quarter q), but
    // the loads of iteration j+1 are issued BEFORE the rows of iteration j are stored to tensor memory, so every warp
    // has up to two batches of loads in flight and the load latency (r1 trace: 300-700 cycles of every ~1100-cycle
    // iteration) overlaps the slot wait and the tcgen05.st of the previous batch.
    // Loads are 8 bytes wide and land DIRECTLY in the registers of the tcgen05.st.16x256b fragment (thread t, repeat
    // g: registers 4g, 4g+1 = bytes 32g + 8(t%4) .. +7 of row t/4, registers 4g+2, 4g+3 = the same bytes of row
    // t/4 + 8): no staging registers and no register moves, and K keeps its natural order (the weight image of this
    // variant is packed without the permutation).  r2 measurement of the first pipelined version (16-byte loads +
    // staging block): 96 registers were not enough, ~20 spill instructions per iteration went to L2 (the shared memory
    // carve-out leaves no L1 at 64x64: 2.9 M local-load sector misses) and the kernel was 1.6x SLOWER.
    const int q = warp & 3, grp = warp >> 2;
    const int j = lane & 3, r8 = lane >> 2;
    const __nv_bfloat16* in = p.in + 4 * j;
    const int t = grp & (T - 1);
    const uint32_t row_off = (uint32_t)(q * 32 + r8) * 4;           // + (h*16 + rr*8)*4 per row of the thread
    const uint32_t ta0 = tmem_base + ((uint32_t)(q * 32) << 16) + colA + (uint32_t)(grp * 32);   // SC == 2 in this variant
    const int total = my_super * nst;

    int lit = 0, lst = 0;
    uint32_t l_idx_tile = 0;
    bool l_absent = false;
    auto issue = [&](uint2 (&v)[2][4][2]) -> bool {      // [16-lane half h][repeat g][row, row + 8]
      if (lst == 0) {                                 // first stage of a pass: its index tile must have landed
        const int n = (lit << lT) + t;
        const int buf = n & (NI - 1);
        l_absent = t == 1 && absent_of(lit);
        l_idx_tile = idx_base + (uint32_t)buf * idx_buf_bytes;
        mbar_wait_sleep(bars + kBarIdx + 8 * buf, (uint32_t)(n >> lNI) & 1u);
      }
      const int c = 2 * lst + (grp >> 1);
      const bool have = c < nchunks && !l_absent;
      if (have) {
        const uint32_t idx_c = l_idx_tile + row_off +
                               (uint32_t)(CIN <= 64 ? c * Cfg::kOffPerChunk : c / Cfg::kChunksPerOff) * kBM * 4;
        const __nv_bfloat16* src = in + (CIN > 64 ? (c % Cfg::kChunksPerOff) * 64 : 0);
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            constexpr int kSlots = CIN <= 64 ? 64 / CIN : 1;      // kernel offsets inside the chunk
            int rw[kSlots];
#pragma unroll
            for (int sl = 0; sl < kSlots; ++sl)
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(rw[sl]) : "r"(idx_c + (uint32_t)(sl * kBM + h * 16 + rr * 8) * 4) : "memory");
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int sl = CIN <= 64 ? (16 * g) / CIN : 0;      // element 16g + 4j of the chunk row
              const int eo = CIN <= 64 ? (16 * g) % CIN : 16 * g;
              v[h][g][rr] = make_uint2(0u, 0u);
              if (rw[sl] >= 0) v[h][g][rr] = __ldg(reinterpret_cast<const uint2*>(src + (size_t)(uint32_t)rw[sl] * CIN + eo));
            }
          }
      }
      if (++lst == nst) { lst = 0; ++lit; }
      return have;
    };
    int s = 0, sst = 0, free_buf = 0;
    uint32_t ph = 0;
    auto store = [&](const uint2 (&v)[2][4][2], bool have) {
      mbar_wait_sleep(bars + kBarEmpty + 8 * s, ph ^ 1u);
      if (have) {
        tc_fence_after();
        const uint32_t ta = ta0 + (uint32_t)(s * 128);
#pragma unroll
        for (int h = 0; h < 2; ++h)
          asm volatile(
              "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
                  ta + ((uint32_t)(h * 16) << 16)),
              "r"(v[h][0][0].x), "r"(v[h][0][0].y), "r"(v[h][0][1].x), "r"(v[h][0][1].y), "r"(v[h][1][0].x), "r"(v[h][1][0].y),
              "r"(v[h][1][1].x), "r"(v[h][1][1].y), "r"(v[h][2][0].x), "r"(v[h][2][0].y), "r"(v[h][2][1].x), "r"(v[h][2][1].y),
              "r"(v[h][3][0].x), "r"(v[h][3][0].y), "r"(v[h][3][1].x), "r"(v[h][3][1].y)
              : "memory");
        tmem_st_wait();
        tc_fence_before();
      }
      if (lane == 0) mbar_arrive(bars + kBarFull + 8 * s);
      if (++s == NS) { s = 0; ph ^= 1u; }
      if (++sst == nst) {                             // all index reads of this pass were issued one iteration ago
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bars + kBarIdxFree + 8 * free_buf);
          mbar_arrive(bars + kBarIdxFree + 8 * (free_buf + 1));
        }
        free_buf = (free_buf + 2) & (NI - 1);
        sst = 0;
      }
    };
    uint2 va[2][4][2], vb[2][4][2];
    bool ha = false, hb = false;
    if (total > 0) ha = issue(va);
    for (int jj = 0; jj < total; jj += 2) {
      if (jj + 1 < total) hb = issue(vb);
      store(va, ha);
      if (jj + 1 >= total) break;
      if (jj + 2 < total) ha = issue(va);
      store(vb, hb);
    }
  } else if (warp < kGatherWarps) {
    // ===================== gather: global/L2 -> registers -> TMEM =====================
    // group g (4 warps, one per row quarter) fills chunk sub-slot g of every stage
    const int q = warp & 3, grp = warp >> 2;
    const int j = lane & 3, r8 = lane >> 2;
    // this thread's two 16-byte pieces of the 128-byte chunk row: p0 = j (bytes 16j..), p1 = 4 + j (bytes 64+16j..)
    int slot0, slot1, eo0, eo1;      // kernel-offset slot inside the chunk and element offset inside the feature row
    if (CIN <= 64) {
      slot0 = (8 * j) / CIN;
      eo0 = (8 * j) % CIN;
      slot1 = (32 + 8 * j) / CIN;
      eo1 = (32 + 8 * j) % CIN;
    } else {
      slot0 = slot1 = 0;
      eo0 = 8 * j;
      eo1 = 32 + 8 * j;
    }
    // WIDE (PIPE == 3, late r2): ONE 32-byte load per row instead of two 16-byte pieces 64 bytes apart.  Lane j of a row's
    // four lanes reads bytes 32j..32j+31 of the 128-byte chunk row, so a warp request covers 8 rows x 128 contiguous
    // bytes = 8 full lines and half the load instructions.  (The hoped-for halving of the L1 data-pipe wavefronts per
    // byte — l1tex__data_pipe_lsu_wavefronts is the busiest unit of the gather, 43 % at 64x64, 74 % in conv_tr<32,32> —
    // did not happen: ncu counts 15.5 wavefronts per LDG.256 request against 8.7 per LDG.128 request, the 32-byte load
    // is served in two 16-byte phases; profiles/r2wide_ncu_summary.json.)
    // An absent neighbour reads a zero row instead of predicating and zero-filling 8 registers.  The K permutation
    // changes with it (source element 16j + w -> K position 16*(w/4) + 4j + w%4; ts_pack_weight mode 2).
    // Same-box A/B on the bench frames, two runs each (profiles/r2_bench_wide{0,1}.json): 64x64 65.3 -> 64.0 us per launch,
    // 128x128 45.0 -> 44.5, step 1.0725 -> 1.0682 ms — small (these layers are latency-bound, their L1 data pipe was at
    // 43 / 27 %), consistent, and fewer instructions: the default.  Forcing conv_ts on the narrow levels stays slower
    // than conv_tr (16x16 42 vs 32 us, 32x32 65 vs 60).
    constexpr bool WIDE = PIPE == 3;
    const int slotw = CIN <= 64 ? (16 * j) / CIN : 0;
    const int eow = CIN <= 64 ? (16 * j) % CIN : 16 * j;
    const char* zrow = reinterpret_cast<const char*>(g_ts_zero_row);
    const __nv_bfloat16* in = p.in;
    const bool tr = q == 0 && grp == 0 && lane == 0;
    int s = 0, gst = 0;
    uint32_t ph = 0;
    const int t = grp & (T - 1);          // chunk slot x = 4*st + grp of a stage belongs to row tile x & 1 = grp & 1
    const int AB = p.ablate;
    for (int it = 0; it < my_super; ++it) {
      // per pass: this warp's tile, its index buffer (waited for once), whether the tile exists
      const int n = (it << lT) + t;                    // tile sequence number of this CTA
      const int buf = n & (NI - 1);
      const bool single = absent_of(it);
      const bool split = split_ok && single;           // single-tile pass: my half of the warps takes half of its stages
      const bool absent = t == 1 && single && !split;
      const int nst_p = split ? (nst + 1) >> 1 : nst;  // stage steps of this pass
      const int st_off = split && t == 1 ? nst_p : 0;  // first stage of my half
      uint32_t idx_tile = idx_base + (uint32_t)buf * idx_buf_bytes;
      wait_bg<PIPE>(bars + kBarIdx + 8 * buf, (uint32_t)(n >> lNI) & 1u);   // also for an absent tile: its fill must have landed before the buffer is handed back
      if (split && t == 1) {                           // the rows are tile 0's: read ITS index tile
        const int n0 = n - 1, buf0 = n0 & (NI - 1);
        wait_bg<PIPE>(bars + kBarIdx + 8 * buf0, (uint32_t)(n0 >> lNI) & 1u);
        idx_tile = idx_base + (uint32_t)buf0 * idx_buf_bytes;
      }
      for (int st = 0; st < nst_p; ++st, ++gst) {
        // this group's chunks of the stage: offsets hc = grp>>1, grp>>1 + 2, ... < SC inside the stage (tile t = grp & 1)
        bool waited = false;
        for (int hc = grp >> 1; hc < SC; hc += 2) {
          const int c = SC * (st + st_off) + hc;
          const bool have = c < nchunks && !absent;
          uint4 v[2][2][2];     // [16-lane half][row, row+8][piece]
          if (TRACE && tr) dbg_stamp(p.dbg, gst, 2);
          if (have && !(AB & 4)) {
            if (TRACE && tr) dbg_stamp(p.dbg, gst, 7);
            const uint32_t idx_c = idx_tile +
                                   (uint32_t)(CIN <= 64 ? c * Cfg::kOffPerChunk : c / Cfg::kChunksPerOff) * kBM * 4;
            const int ehalf = CIN > 64 ? (c % Cfg::kChunksPerOff) * 64 : 0;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int rr = 0; rr < 2; ++rr) {
                const int row = q * 32 + h * 16 + rr * 8 + r8;
                if (WIDE) {
                  int rw;
                  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(rw) : "r"(idx_c + (uint32_t)(slotw * kBM + row) * 4) : "memory");
                  const void* src = rw >= 0 ? reinterpret_cast<const char*>(in + (size_t)(uint32_t)rw * CIN + ehalf + eow) : zrow;
                  ldg256(src, v[h][rr][0], v[h][rr][1]);
                  continue;
                }
                int rw0, rw1;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(rw0) : "r"(idx_c + (uint32_t)(slot0 * kBM + row) * 4) : "memory");
                if (CIN <= 32) {
                  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(rw1) : "r"(idx_c + (uint32_t)(slot1 * kBM + row) * 4) : "memory");
                } else {
                  rw1 = rw0;
                }
                v[h][rr][0] = make_uint4(0u, 0u, 0u, 0u);
                v[h][rr][1] = make_uint4(0u, 0u, 0u, 0u);
                // (r1 A/B: 8-byte loads that land directly in the register pairs of the tcgen05.st fragment remove the 32
                // register moves per chunk but double the load instructions: 5 % SLOWER, the LSU is the next limit)
                if (rw0 >= 0) v[h][rr][0] = __ldg(reinterpret_cast<const uint4*>(in + (size_t)(uint32_t)rw0 * CIN + ehalf + eo0));
                if (rw1 >= 0) v[h][rr][1] = __ldg(reinterpret_cast<const uint4*>(in + (size_t)(uint32_t)rw1 * CIN + ehalf + eo1));
              }
          }
          if (!waited) {
            wait_bg<PIPE>(bars + kBarEmpty + 8 * s, ph ^ 1u);
            waited = true;
          }
          if (TRACE && tr) dbg_stamp(p.dbg, gst, 3);
          if (have && !(AB & 6)) {
            tc_fence_after();
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + colA + (uint32_t)s * stage_cols + (uint32_t)((2 * hc + t) * 32);
            tmem_st_16x256b_x4(ta, v[0][0][0], v[0][1][0], v[0][0][1], v[0][1][1]);
            tmem_st_16x256b_x4(ta + (16u << 16), v[1][0][0], v[1][1][0], v[1][0][1], v[1][1][1]);
          }
        }
        if (!waited) wait_bg<PIPE>(bars + kBarEmpty + 8 * s, ph ^ 1u);     // a group without a chunk in this stage (SC = 1)
        tmem_st_wait();
        tc_fence_before();
        if (lane == 0) mbar_arrive(bars + kBarFull + 8 * s);
        if (TRACE && tr) dbg_stamp(p.dbg, gst, 4);
        if (++s == NS) { s = 0; ph ^= 1u; }
      }
      // all index reads of this super-tile are consumed: hand its buffers back
      __syncwarp();
      if (lane == 0)
        for (int t = 0; t < T; ++t) mbar_arrive(bars + kBarIdxFree + 8 * (((it << lT) + t) & (NI - 1)));
    }
  } else if (warp < kEpiWarp0 + 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int ntile_seq = my_super << lT;
    for (int n = 0; n < ntile_seq; ++n) {
      const bool split = split_ok && absent_of(n >> lT);
      if (split && (n & 1)) continue;                  // the second accumulator of a split pass is added to the first below
      const int tile = tile_of(n);
      const int a = n & (ND - 1);
      const int row = tile * kBM + q * 32 + lane;
      const bool live = row < no;
      uint4 rv[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
      const uint4* rp = reinterpret_cast<const uint4*>(p.residual + (size_t)row * COUT);
      if ((p.epi & COMB_EPI_RESIDUAL) && live) {   // issued before the accumulator is ready
        rv[0] = __ldg(rp);
        rv[1] = __ldg(rp + 1);
      }
      wait_bg<PIPE>(bars + kBarTFull + 8 * a, (uint32_t)(n / ND) & 1u);
      if (split) wait_bg<PIPE>(bars + kBarTFull + 8 * (a + 1), (uint32_t)((n + 1) / ND) & 1u);   // n even, ND even: a + 1 < ND
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * COUT;
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
        if (split) {
          // upper-half partial sums from the second accumulator.  tcgen05.ld is .sync.aligned: every lane of the warp
          // must execute it, so it stays OUTSIDE the per-row `live` branch (a partially live last tile hung the kernel
          // when it sat inside)
          uint32_t w2[16];
          tmem_ld16(taddr + COUT + c0, w2);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(w2[i]));
        }
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
      if (lane == 0) {
        mbar_arrive(bars + kBarTEmpty + 8 * a);
        if (split) mbar_arrive(bars + kBarTEmpty + 8 * (a + 1));
      }
      if (TRACE && warp == kEpiWarp0 && lane == 0) dbg_stamp(p.dbg, n, 5);
    }
  } else if (warp == kMmaWarp || warp == kMmaWarp2) {
    // ===================== MMA issuers =====================
    // TWO issuing threads, one per row tile of the pass (separate accumulators, so no ordering between them is needed).
    // Micro-benchmark (scripts/micro/mma_issue_bench.cu): back-to-back tcgen05.mma (M=128, K=16) cost 45 cycles each
    // for N <= 64 and 64 at N = 128, from tensor or shared memory alike; r1 traces of a single issuing thread showed
    // 74-90 cycles per MMA because its barrier checks and operand set-up (~100 cycles per check even on a completed
    // phase) starve the shallow MMA queue.  With two threads one issues while the other looks at barriers.
    // One barrier round per STAGE of 4 chunks = (c, tile 0), (c, tile 1), (c+1, tile 0), (c+1, tile 1); each thread
    // issues the two chunks of its tile, both commit to the stage's empty barriers (count 2).
    // instruction descriptor: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, K-major A/B, N>>3 [17,23), M>>4 [24,29)
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(COUT >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
    const int my_t = warp == kMmaWarp ? 0 : 1;
    const int AB = p.ablate;
    if (elect_one_sync()) {
      int s = 0, bs = 0, gst = 0;
      uint32_t ph = 0, bph = 0;
      bool ready = false;
      if (bres) mbar_wait(bars + kBarB, 0);
      auto ring_adv = [&](int& b, uint32_t& phs) {
        if (++b == NB) { b = 0; phs ^= 1u; }
      };
      auto last_pass_of = [&](int it) -> bool { return it == my_super - 1; };
      // per pass: my tile's accumulator and whether the tile exists; per stage only ring arithmetic (the loop nest keeps
      // the live state of this thread small: its issue loop is the critical path of the kernel)
      for (int it = 0; it < my_super; ++it) {
        const int n = (it << lT) + my_t;              // tile sequence number of my tile in this pass
        const bool single = absent_of(it);            // single-tile pass
        const bool split = split_ok && single;        // ... whose upper stages are mine (my_t == 1), see split_ok
        const bool absent = my_t == 1 && single && !split;
        const int nst_p = split ? (nst + 1) >> 1 : nst;
        const int st_off = split && my_t == 1 ? nst_p : 0;
        const bool split_nx = !last_pass_of(it) && split_ok && absent_of(it + 1);
        const uint32_t acc = (uint32_t)(n & (ND - 1));
        const uint32_t tmem_d = tmem_base + acc * COUT;
        const uint32_t acc_ph = (uint32_t)(n / ND) & 1u;
        const bool last_pass = it == my_super - 1;
        for (int st = 0; st < nst_p; ++st, ++gst) {
          if (TRACE && my_t == 0) dbg_stamp(p.dbg, gst, 6);
          // streamed weights: (bs, bph) is the ring entry of tile 0's chunks of this stage; in a split pass the entry
          // behind it carries the chunks of the upper half (mine if my_t == 1) and a stage consumes two entries
          int mb = bs;
          uint32_t mph = bph;
          if (split && my_t == 1) ring_adv(mb, mph);
          if (!ready) {
            if (st == 0) mbar_wait(bars + kBarTEmpty + 8 * acc, acc_ph ^ 1u);
            if (!bres) mbar_wait(bars + kBarBFull + 8 * mb, mph);
            mbar_wait(bars + kBarFull + 8 * s, ph);
          }
          if (TRACE && my_t == 0) dbg_stamp(p.dbg, gst, 0);
          tc_fence_after();
          // ring positions of the next stage
          const int s2 = s + 1 == NS ? 0 : s + 1;
          const uint32_t ph2 = s + 1 == NS ? ph ^ 1u : ph;
          const bool last_st = st == nst_p - 1;
          int bs2 = bs;
          uint32_t bph2 = bph;
          ring_adv(bs2, bph2);
          if (split) ring_adv(bs2, bph2);
          int mb2 = bs2;                                // my entry of the NEXT stage (which may open the next pass)
          uint32_t mph2 = bph2;
          if ((last_st ? split_nx : split) && my_t == 1) ring_adv(mb2, mph2);
          const uint32_t a_stage = tmem_base + colA + (uint32_t)s * stage_cols + (uint32_t)(my_t * 32);
          for (int h = 0; h < SC; ++h) {
            const int c = SC * (st + st_off) + h;       // my chunk of the stage: slot 2*h + my_t
            if (c < nchunks && !absent && !(AB & 1)) {
              const uint32_t tmem_a = a_stage + (uint32_t)(h * 64);
              const uint64_t bdesc = make_desc_sw128(w_base + (uint32_t)(bres ? c : mb * SC + h) * Cfg::kBBytes);
#pragma unroll
              for (int kk = 0; kk < kChunkK / 16; ++kk)
                umma_bf16_ts(tmem_d, tmem_a + 8 * kk, bdesc + 2 * kk, idesc, (st | h | kk) != 0 ? 1u : 0u);
            }
            if (h == 0) {   // probe the next stage while the other chunks of this one are still to be issued
              ready = !(last_pass && last_st) && !(AB & 32) && mbar_test(bars + kBarFull + 8 * s2, ph2);
              if (!bres) ready = ready && mbar_test(bars + kBarBFull + 8 * mb2, mph2);
              if (last_st) {
                const int n2 = n + T;
                ready = ready && mbar_test(bars + kBarTEmpty + 8 * (n2 & (ND - 1)), ((uint32_t)(n2 / ND) & 1u) ^ 1u);
              }
            }
          }
          umma_commit(bars + kBarEmpty + 8 * s);
          if (!bres) {
            // Every ring entry expects two arrivals.  In a split pass an entry has ONE reader; by default both threads
            // still commit on both entries of the stage (an entry is then handed back when the MMAs of both threads up to
            // this stage have retired — conservative, and the two-committers-per-barrier pattern of the normal passes);
            // p.split & 2: the reader commits twice on its own entry instead.
            if (!split) {
              umma_commit(bars + kBarBEmpty + 8 * mb);
            } else if (p.split & 2) {
              umma_commit(bars + kBarBEmpty + 8 * mb);
              umma_commit(bars + kBarBEmpty + 8 * mb);
            } else {
              int ob = bs;                                   // the other entry of the stage
              if (my_t == 0) { uint32_t dummy = 0; ring_adv(ob, dummy); }
              umma_commit(bars + kBarBEmpty + 8 * mb);
              umma_commit(bars + kBarBEmpty + 8 * ob);
            }
          }
          if (last_st) umma_commit(bars + kBarTFull + 8 * acc);
          if (TRACE && my_t == 0) dbg_stamp(p.dbg, gst, 1);
          s = s2;
          ph = ph2;
          if (!bres) { bs = bs2; bph = bph2; }
        }
      }
    }
    __syncwarp();
  } else if (warp == kIdxWarp) {
    // ===================== neighbour-index prefetcher =====================
    // idx[buf][k][r] = nbr[k][tile*128 + r] (-1 beyond the last row), NI buffers ahead of the gather warps.
    const bool vec_ok = (p.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.nbr) & 15) == 0);
    const int ntile_seq = my_super << lT;
    for (int n = 0; n < ntile_seq; ++n) {
      const int buf = n & (NI - 1);
      const int use = n >> lNI;            // how many times this buffer has been filled before
      if (use > 0) wait_bg<PIPE>(bars + kBarIdxFree + 8 * buf, (uint32_t)(use - 1) & 1u);
      const int tile = tile_of(n);
      const int row0 = tile * kBM + lane * 4;   // this lane: 4 consecutive rows
      const uint32_t dst = idx_base + (uint32_t)buf * idx_buf_bytes + lane * 16;
      if (p.ablate & 16) {
      } else if (vec_ok && row0 + 3 < no) {
        for (int k = 0; k < K; ++k)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + k * kBM * 4),
                       "l"(p.nbr + (size_t)k * p.ld + row0)
                       : "memory");
      } else {
        for (int k = 0; k < K; ++k)
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            if (row0 + qq < no)
              cp_async4(dst + k * kBM * 4 + qq * 4, p.nbr + (size_t)k * p.ld + row0 + qq);
            else
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + k * kBM * 4 + qq * 4), "r"(-1) : "memory");
          }
      }
      cp_async_mbar_arrive_noinc(bars + kBarIdx + 8 * buf);
    }
  } else if (warp == kBWarp && lane == 0) {
    // ===================== weights =====================
    if (bres) {
      const uint32_t bytes = (uint32_t)nchunks * Cfg::kBBytes;
      mbar_arrive_expect_tx(bars + kBarB, bytes);
      bulk_copy_g2s(w_base, p.wpacked, bytes, bars + kBarB);
    } else {
      // one SC-chunk weight stage per A stage (a stage is chunks SC*st .. SC*st+SC-1 for both row tiles)
      int bs = 0;
      uint32_t bph = 0;
      for (int it = 0; it < my_super; ++it) {
        // a split pass (see split_ok) interleaves the stage sequences of its two halves: entry 2*st = stage st,
        // entry 2*st + 1 = stage nst_p + st (possibly past the end: an empty entry, completed by a plain arrive)
        const bool split = split_ok && absent_of(it);
        const int nst_p = split ? (nst + 1) >> 1 : nst;
        for (int e = 0; e < (split ? 2 * nst_p : nst_p); ++e) {
          const int c0 = SC * (split ? (e >> 1) + (e & 1) * nst_p : e);
          const int left = nchunks - c0;
          const uint32_t bytes = (uint32_t)(left < SC ? (left > 0 ? left : 0) : SC) * Cfg::kBBytes;
          mbar_wait(bars + kBarBEmpty + 8 * bs, bph ^ 1u);
          if (bytes != 0) {
            mbar_arrive_expect_tx(bars + kBarBFull + 8 * bs, bytes);
            bulk_copy_g2s(w_base + bs * SC * Cfg::kBBytes, p.wpacked + (size_t)c0 * Cfg::kBBytes, bytes, bars + kBarBFull + 8 * bs);
          } else {
            mbar_arrive(bars + kBarBFull + 8 * bs);
          }
          if (++bs == NB) { bs = 0; bph ^= 1u; }
        }
      }
    }
  }

  if (TRACE && warp == kMmaWarp && lane == 0) dbg_cta_time(p.dbg, 2);
  tc_fence_before();
  __syncthreads();
  if (TRACE && tid == 0) dbg_cta_time(p.dbg, 3);
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// Weight image: chunk c, output channel n, K position kappa (see the permutation above) -> W[n][k][ci], in the
// K-major SWIZZLE_128B layout the B descriptor expects.
template <int CIN>
__global__ void __launch_bounds__(256) ts_pack_kernel(const float* __restrict__ w, int Cout, int K, int Cin_real,
                                                       int nchunks, int natural, __nv_bfloat16* __restrict__ out) {
  const long long total = (long long)nchunks * Cout * kChunkK;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int kappa = (int)(i % kChunkK);
    const int n = (int)((i / kChunkK) % Cout);
    const int c = (int)(i / ((long long)kChunkK * Cout));
    const int half = kappa >> 5, wq = (kappa >> 4) & 1, j = (kappa >> 2) & 3, wl = kappa & 3;
    // source element of the chunk row: natural order (1), the wide-load fragment order (2) or the two-piece order (0)
    const int e = natural == 1 ? kappa : natural == 2 ? 16 * j + 4 * (kappa >> 4) + wl : 8 * (4 * half + j) + 4 * wq + wl;
    int k, ci;
    if constexpr (CIN <= 64) {
      k = c * (64 / CIN) + e / CIN;
      ci = e % CIN;
    } else {
      k = c / (CIN / 64);
      ci = (c % (CIN / 64)) * 64 + e;
    }
    float v = 0.0f;
    if (k < K && ci < Cin_real) v = w[((size_t)n * K + k) * Cin_real + ci];
    const int qq = kappa >> 3, within = kappa & 7;
    const size_t byte_off = (size_t)c * Cout * 128 + (size_t)(n >> 3) * 1024 + (n & 7) * 128 + ((qq ^ (n & 7)) << 4) + within * 2;
    out[byte_off / 2] = __float2bfloat16(v);
  }
}

static int ts_pipe() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("COMB_TS_PIPE");
    v = e ? atoi(e) : 0;      // 0: r1 loop; 1: pipelined gather; 2: sleep-free background waits
#ifndef COMB_TS_EXPERIMENTS
    v = 0;                    // the variants are not part of the product library (nvcc -DCOMB_TS_EXPERIMENTS builds them)
#endif
    if (v == 0) {             // 3: the r1 loop with one 32-byte load per row (COMB_TS_WIDE, see the gather)
      const char* w = getenv("COMB_TS_WIDE");
      if (w ? atoi(w) != 0 : kWideDefault) v = 3;
    }
  }
  return v;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

template <int CIN, int COUT>
int launch_ts(const ConvFwdArgs& p_in, cudaStream_t stream) {
  using Cfg = TsCfg<CIN, COUT>;
  ConvFwdArgs p = p_in;
  // Shared-memory rings.  Whatever the kernel does not take stays L1 cache, and the gather lives on L1 hits, so the
  // rings are not made as deep as the 227 KB allow: COMB_TS_NI index-tile buffers (8 or 4; default 4) and at most
  // COMB_TS_NB streamed-weight stages (default 4).
  static const int ni_env = env_int("COMB_TS_NI", 4), nb_env = env_int("COMB_TS_NB", 4), blocked_env = env_int("COMB_TS_BLOCKED", 0);
  p.blocked = blocked_env;
  p.ni = (ni_env == 8 && Cfg::n_idx(p.K) == 8) ? 8 : 4;
  // chunks of each row tile per A stage.  r2 A/B (gpu_r2_m): SC = 3 where tensor memory allows (6 slots, 2 stages)
  // against SC = 2 (4 slots, 3 stages): 1.018 vs 0.997 ms of conv per step — the larger stage does not pay, the default
  // stays 2 (COMB_TS_SC=1..3 to override; the pipelined-gather variant is written for SC = 2)
  static const int sc_env = env_int("COMB_TS_SC", 2);
  int sc_max = (512 - Cfg::d_cols(p.K)) / (2 * 2 * 32);      // 2 stages x 2 tiles x 32 columns per chunk slot
  if (sc_max > 3) sc_max = 3;
  int sc = sc_env >= 1 && sc_env <= sc_max ? sc_env : 2;
  if (ts_pipe() == 1) sc = 2;
  p.sc = sc;
  {
    const char* ab = getenv("COMB_TS_ABLATE");     // read per launch: timing experiments flip it inside one process
    p.ablate = ab ? atoi(ab) : 0;
  }
  const int idx_bytes = p.ni * Cfg::k_pad(p.K) * kBM * 4;
  const bool bres = Cfg::b_resident(p.K);
  int nb = bres ? 0 : (kSmemBudget - 2048 - idx_bytes) / (sc * Cfg::kBBytes);
  if (nb > kMaxBStages) nb = kMaxBStages;
  if (!bres && nb > nb_env && nb_env >= 2) nb = nb_env;
  p.nb = nb;
  const size_t smem = 2048 + (size_t)idx_bytes + (bres ? (size_t)Cfg::num_chunks(p.K) * Cfg::kBBytes : (size_t)nb * sc * Cfg::kBBytes);
  static thread_local DevOnce configured;   // per device: the attribute is a per-device property
  if (configured.first()) {
    COMB_CUDA(cudaFuncSetAttribute(spconv_ts_kernel<CIN, COUT, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
    COMB_CUDA(cudaFuncSetAttribute(spconv_ts_kernel<CIN, COUT, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
    COMB_CUDA(cudaFuncSetAttribute(spconv_ts_kernel<CIN, COUT, false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
    COMB_CUDA(cudaFuncSetAttribute(spconv_ts_kernel<CIN, COUT, true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
#ifdef COMB_TS_EXPERIMENTS
    COMB_CUDA(cudaFuncSetAttribute(spconv_ts_kernel<CIN, COUT, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
    COMB_CUDA(cudaFuncSetAttribute(spconv_ts_kernel<CIN, COUT, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024));
#endif
  }
  if (smem > 227 * 1024 - 1024 || (!bres && nb < 2)) {
    set_error("comb_spconv_fwd_bf16: shared memory %zu exceeds the per-CTA limit", smem);
    return COMB_EINVAL;
  }
  // one CTA per ROW TILE up to the SM count (not per super-tile): the kernel turns a last wave of r <= grid/2 super-tiles
  // into 2r single-tile passes, which needs the CTAs to exist
  static const int split_env = env_int("COMB_TS_SPLIT", 1);
  p.split = split_env;
  const int ntiles_max = cdiv(p.no_max, kBM);
  const int grid = ntiles_max < sm_count() ? ntiles_max : sm_count();
  if (p.dbg != nullptr && ts_pipe() == 3) COMB_CUDA(launch_pdl(spconv_ts_kernel<CIN, COUT, true, 3>, grid, kThreads, smem, stream, p));
  else if (p.dbg != nullptr) COMB_CUDA(launch_pdl(spconv_ts_kernel<CIN, COUT, true, 0>, grid, kThreads, smem, stream, p));   // pipeline trace build
  else if (ts_pipe() == 3) COMB_CUDA(launch_pdl(spconv_ts_kernel<CIN, COUT, false, 3>, grid, kThreads, smem, stream, p));
#ifdef COMB_TS_EXPERIMENTS     // the two measured-and-rejected gather variants (software-pipelined gather, sleep-free waits): built on demand only
  else if (ts_pipe() == 1) COMB_CUDA(launch_pdl(spconv_ts_kernel<CIN, COUT, false, 1>, grid, kThreads, smem, stream, p));
  else if (ts_pipe() == 2) COMB_CUDA(launch_pdl(spconv_ts_kernel<CIN, COUT, false, 2>, grid, kThreads, smem, stream, p));
#endif
  else COMB_CUDA(launch_pdl(spconv_ts_kernel<CIN, COUT, false, 0>, grid, kThreads, smem, stream, p));
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

template <int CIN>
int dispatch_cout(int Cout, const ConvFwdArgs& p, cudaStream_t stream) {
  switch (Cout) {
    case 16: return launch_ts<CIN, 16>(p, stream);
    case 32: return launch_ts<CIN, 32>(p, stream);
    case 64: return launch_ts<CIN, 64>(p, stream);
    case 128: return launch_ts<CIN, 128>(p, stream);
  }
  set_error("comb_spconv_fwd_bf16: Cout %d not in {16,32,64,128}", Cout);
  return COMB_EINVAL;
}

}  // namespace

int ts_fwd_bf16(const ConvFwdArgs& p, int Cin_p, int Cout, cudaStream_t stream) {
  COMB_CHECK_ARG(p.K >= 1 && p.K <= kMaxK, "comb_spconv_fwd_bf16: K %d outside [1,%d]", p.K, kMaxK);
  switch (Cin_p) {
    case 16: return dispatch_cout<16>(Cout, p, stream);
    case 32: return dispatch_cout<32>(Cout, p, stream);
    case 64: return dispatch_cout<64>(Cout, p, stream);
    case 128: return dispatch_cout<128>(Cout, p, stream);
  }
  set_error("comb_spconv_fwd_bf16: Cin_p %d not in {16,32,64,128}", Cin_p);
  return COMB_EINVAL;
}

int ts_pack_weight(const float* weight, int Cout, int K, int Cin, int Cin_p, int nchunks, int natural_in, void* wpacked,
                   cudaStream_t stream) {
  const long long total = (long long)nchunks * Cout * kChunkK;
  const int grid = cdiv(total, 256);
  __nv_bfloat16* out = (__nv_bfloat16*)wpacked;
  // the pipelined gather (and conv_tr) keep K in its natural order; the wide-load gather has its own fragment order
  const int natural = (natural_in || ts_pipe() == 1) ? 1 : ts_pipe() == 3 ? 2 : 0;
  switch (Cin_p) {
    case 16: ts_pack_kernel<16><<<grid, 256, 0, stream>>>(weight, Cout, K, Cin, nchunks, natural, out); break;
    case 32: ts_pack_kernel<32><<<grid, 256, 0, stream>>>(weight, Cout, K, Cin, nchunks, natural, out); break;
    case 64: ts_pack_kernel<64><<<grid, 256, 0, stream>>>(weight, Cout, K, Cin, nchunks, natural, out); break;
    case 128: ts_pack_kernel<128><<<grid, 256, 0, stream>>>(weight, Cout, K, Cin, nchunks, natural, out); break;
    default: COMB_CHECK_ARG(false, "comb_spconv_pack_weight_bf16: Cin_p %d not in {16,32,64,128}", Cin_p);
  }
  COMB_LAUNCH_CHECK();
  return COMB_OK;
}

}  // namespace comb
