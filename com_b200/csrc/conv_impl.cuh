// Internal interface between the C-ABI entry points of the bf16 sparse convolution (conv_tc.cu) and its
// implementations, the two tcgen05 kernels: "ts" (A operand gathered into tensor memory, conv_ts.cu — the default) and
// "ss" (A operand gathered into shared memory, conv_tc.cu — kept for A/B measurements, COMB_CONV_IMPL=ss).  (The r1
// warp-level mma.sync experiment for the narrow levels, measured slower, lives in scripts/micro/conv_wm.cu and is no
// longer part of the product library.)
// The packed weight image differs between the two (the ts form permutes K inside a chunk), so the choice is
// made once per process and used by both comb_spconv_pack_weight_bf16 and comb_spconv_fwd_bf16.
#pragma once
#include <stdlib.h>
#include "common.cuh"

namespace comb {

struct ConvFwdArgs {
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
  int blocked = 0;  // conv_ts: contiguous super-tile range per CTA (1) or strided assignment (0, default)
  int ni = 4;       // conv_ts: index-tile ring depth (set by launch_ts)
  int nb = 0;       // conv_ts: streamed-weight stages (set by launch_ts)
  int sc = 2;       // conv_ts: chunks of each row tile per A stage (set by launch_ts)
  int split = 1;    // conv_ts: single-tile passes split their K range over the two halves of the pipeline (COMB_TS_SPLIT=0: r1 behaviour)
  int ablate = 0;   // conv_ts: COMB_TS_ABLATE bit mask — pipeline pieces switched off for timing experiments (results are garbage)
};

// Launch, optionally with programmatic stream serialization (COMB_PDL=1; the kernels call griddepcontrol.wait before
// they read anything an earlier kernel wrote, so their set-up can overlap the tail of the previous kernel of the stream,
// also inside a captured graph).  r2 A/B in the bench, same box, two runs each: 1.102 / 1.107 ms per step with it against
// 1.099 / 1.093 without — the 2.2 us of set-up per launch are not what separates the 21 convs of the chain — so the
// default is a plain launch.
template <typename Kernel>
static inline cudaError_t launch_pdl(Kernel kernel, int grid, int block, size_t smem, cudaStream_t stream, const ConvFwdArgs& p) {
  static const bool on = [] {
    const char* e = getenv("COMB_PDL");
    return e && e[0] == '1';
  }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = on ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, p);
}

int ts_fwd_bf16(const ConvFwdArgs& p, int Cin_p, int Cout, cudaStream_t stream);
// natural: K in tap-major / channel-minor order (conv_tr.cu and the pipelined conv_ts variant); otherwise the permuted
// order of the tcgen05.st.16x256b fragment (conv_ts.cu)
int ts_pack_weight(const float* weight, int Cout, int K, int Cin, int Cin_p, int nchunks, int natural, void* wpacked,
                   cudaStream_t stream);
// row-per-thread form (conv_tr.cu)
int tr_fwd_bf16(const ConvFwdArgs& p, int Cin_p, int Cout, cudaStream_t stream);
bool tr_supported(int Cin_p, int Cout, int K);     // weight image fits in shared memory


}  // namespace comb
