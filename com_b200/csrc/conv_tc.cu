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
// Warp roles (288 threads):
//   warps 0-3  producers : neighbour-index tile prefetch (cp.async 4 B, double buffered) and the
//                          A gather (8 x 16 B cp.async per thread per chunk); thread 0 also streams
//                          the pre-swizzled weight chunk with one cp.async.bulk (TMA engine).
//   warps 4-7  epilogue  : tcgen05.ld 32 lanes x Cout columns -> registers -> epilogue -> global.
//   warp  8    MMA       : lane 0 issues 4 x tcgen05.mma (M=128, N=Cout, K=16) per chunk, commits
//                          to the stage's empty barrier; allocates / frees TMEM (2 accumulators).
#include "common.cuh"

namespace comb {
namespace {

constexpr int kBM = 128;          // rows per tile (UMMA M)
constexpr int kChunkK = 64;       // bf16 elements per K chunk (128 bytes)
constexpr int kABytes = kBM * 128;
constexpr int kProducers = 128;
constexpr int kThreadsTC = 288;
constexpr int kMaxK = 32;         // kernel offsets supported by the index tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) (unused for swizzled K-major, 1) | SBO>>4 [32,46) = 1024 B between
// 8-row groups | version=1 [46,48) | layout_type=2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

template <int CIN, int COUT>
struct TcCfg {
  static constexpr int kOffPerChunk = CIN <= 64 ? 64 / CIN : 1;  // kernel offsets per K chunk
  static constexpr int kChunksPerOff = CIN <= 64 ? 1 : CIN / 64;
  static constexpr int kPiecesPerOff = CIN <= 64 ? CIN / 8 : 8;   // 16-byte pieces of one offset inside a chunk
  static constexpr int kBBytes = COUT * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = COUT >= 128 ? 5 : 6;
  static constexpr int kTmemCols = 2 * COUT < 32 ? 32 : 2 * COUT;
  static __host__ __device__ int num_chunks(int K) {
    return CIN <= 64 ? (K + kOffPerChunk - 1) / kOffPerChunk : K * kChunksPerOff;
  }
  static size_t smem_bytes(int K) {
    return 1024 /*align slack*/ + (size_t)kStages * kStageBytes + 2 * (size_t)K * kBM * 4 + 256;
  }
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
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(kThreadsTC, 1) spconv_tc_kernel(TcParams p) {
  using Cfg = TcCfg<CIN, COUT>;
  constexpr int NS = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_full[NS], bar_empty[NS], bar_tfull[2], bar_tempty[2], bar_idx[2];
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int no = eff_n(p.no_max, p.no_dev);
  const int ntiles = (no + kBM - 1) / kBM;
  const int K = p.K;
  const int nchunks = Cfg::num_chunks(K);

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t idx_base = smem_base + NS * Cfg::kStageBytes;  // [2][K][128] ints
  auto stageA = [&](int s) { return smem_base + s * Cfg::kStageBytes; };
  auto stageB = [&](int s) { return smem_base + s * Cfg::kStageBytes + kABytes; };

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(smem_u32(&bar_full[s]), kProducers + 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&bar_tfull[a]), 1);
      mbar_init(smem_u32(&bar_tempty[a]), 4);
      mbar_init(smem_u32(&bar_idx[a]), kProducers);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                 "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  if (warp < 4) {
    // ===================== producers =====================
    const uint8_t* in_bytes = reinterpret_cast<const uint8_t*>(p.in);
    auto prefetch_idx = [&](int tile, int buf) {
      const int row = tile * kBM + tid;
      const uint32_t dst = idx_base + (uint32_t)buf * K * kBM * 4 + tid * 4;
      if (tile < ntiles && row < no) {
        for (int k = 0; k < K; ++k) cp_async4(dst + k * kBM * 4, p.nbr + (size_t)k * p.ld + row);
      } else {
        for (int k = 0; k < K; ++k) asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + k * kBM * 4), "r"(-1) : "memory");
      }
      // arrives once the thread's cp.asyncs have landed (immediately when it only stored -1)
      cp_async_mbar_arrive_noinc(smem_u32(&bar_idx[buf]));
    };
    int it = 0;
    uint32_t g = 0;  // global chunk counter -> stage / phase
    prefetch_idx(blockIdx.x, 0);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      // all producers are done reading the other index buffer (used by the previous tile)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      prefetch_idx(tile + gridDim.x, buf ^ 1);
      mbar_wait(smem_u32(&bar_idx[buf]), (it >> 1) & 1);
      const uint32_t idx_tile = idx_base + (uint32_t)buf * K * kBM * 4;
      for (int c = 0; c < nchunks; ++c, ++g) {
        const int s = g % NS;
        const uint32_t ph = (g / NS) & 1;
        mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
        const uint32_t a_base = stageA(s);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int e = j * kProducers + tid;
          const int r = e >> 3, q = e & 7;
          int k, piece;
          if constexpr (CIN <= 64) {
            k = c * Cfg::kOffPerChunk + q / Cfg::kPiecesPerOff;
            piece = q % Cfg::kPiecesPerOff;
          } else {
            k = c / Cfg::kChunksPerOff;
            piece = (c % Cfg::kChunksPerOff) * 8 + q;
          }
          int src_row = -1;
          if (k < K) {
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(src_row) : "r"(idx_tile + (k * kBM + r) * 4) : "memory");
          }
          const uint32_t dst = a_base + (r >> 3) * 1024 + (r & 7) * 128 + ((q ^ (r & 7)) << 4);
          const uint8_t* src = in_bytes + (src_row >= 0 ? ((size_t)src_row * CIN * 2 + piece * 16) : 0);
          cp_async16(dst, src, src_row >= 0 ? 16u : 0u);
        }
        cp_async_mbar_arrive_noinc(smem_u32(&bar_full[s]));
        if (tid == 0) {
          mbar_arrive_expect_tx(smem_u32(&bar_full[s]), Cfg::kBBytes);
          bulk_copy_g2s(stageB(s), p.wpacked + (size_t)c * Cfg::kBBytes, Cfg::kBBytes, smem_u32(&bar_full[s]));
        }
      }
    }
  } else if (warp < 8) {
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
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
                float2 t = __bfloat1622float2(r2[i]);
                f[h * 8 + 2 * i] += t.x;
                f[h * 8 + 2 * i + 1] += t.y;
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
    }
  } else {
    // ===================== MMA issuer =====================
    // instruction descriptor: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, K-major A/B, N>>3 [17,23), M>>4 [24,29)
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(COUT >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
    int it = 0;
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int a = it & 1;
      mbar_wait(smem_u32(&bar_tempty[a]), ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + a * COUT;
      for (int c = 0; c < nchunks; ++c, ++g) {
        const int s = g % NS;
        mbar_wait(smem_u32(&bar_full[s]), (g / NS) & 1);
        fence_proxy_async();
        tc_fence_after();
        if (lane == 0) {
          const uint64_t adesc = make_desc_sw128(stageA(s));
          const uint64_t bdesc = make_desc_sw128(stageB(s));
#pragma unroll
          for (int kk = 0; kk < kChunkK / 16; ++kk) {
            // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in the (>>4) address field
            umma_bf16(tmem_d, adesc + 2 * kk, bdesc + 2 * kk, idesc, (c | kk) != 0 ? 1u : 0u);
          }
          umma_commit(smem_u32(&bar_empty[s]));
          if (c == nchunks - 1) umma_commit(smem_u32(&bar_tfull[a]));
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
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
int launch_tc(const TcParams& p, cudaStream_t stream) {
  using Cfg = TcCfg<CIN, COUT>;
  const size_t smem = Cfg::smem_bytes(p.K);
  static thread_local bool configured = false;
  if (!configured) {
    COMB_CUDA(cudaFuncSetAttribute(spconv_tc_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
    configured = true;
  }
  if (smem > 227 * 1024 - 2048) {
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

static int chunks_for(int Cin_p, int K) {
  return Cin_p <= 64 ? (K + 64 / Cin_p - 1) / (64 / Cin_p) : K * (Cin_p / 64);
}

}  // namespace
}  // namespace comb

using namespace comb;

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
  COMB_CHECK_ARG((ld % 1) == 0, "comb_spconv_fwd_bf16: ld");
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
  switch (Cin_p) {
    case 16: return dispatch_cout<16>(Cout, p, stream);
    case 32: return dispatch_cout<32>(Cout, p, stream);
    case 64: return dispatch_cout<64>(Cout, p, stream);
    case 128: return dispatch_cout<128>(Cout, p, stream);
  }
  set_error("comb_spconv_fwd_bf16: Cin_p %d not in {16,32,64,128}", Cin_p);
  return COMB_EINVAL;
}
