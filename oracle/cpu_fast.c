/*
 * cpu_fast.c — multi-threaded fp32 CPU implementation of the sparse-conv forward, used ONLY as the
 * CPU BASELINE that bench.py times (cpu_baseline / --impl reference legs).  TEST INFRASTRUCTURE.
 *
 * Same semantics as orc_conv_fwd in oracle.c (spconv's gather-GEMM-scatter, restated — spconv is not
 * vendored by the reference: docs/INSTALL.md:9, pcdet/models/backbones_3d/spconv_backbone.py:191-232),
 * but written the way a CPU implementation would be shipped: fp32 accumulation, weights transposed
 * to [K][Cin][Cout] so the inner loop vectorises over Cout, OpenMP over output rows, fused
 * per-channel affine (eval BatchNorm1d) + residual + ReLU (spconv_backbone.py:21-25,50-66).
 * Checked against orc_conv_fwd in tests/test_oracle_spconv.py.
 *
 * Build: gcc -O3 -mavx2 -mfma -fopenmp -fPIC -shared cpu_fast.c -o liboracle_fast.so
 */
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_fast_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_fast_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* W is (Cout, K, Cin); nbr is (K, no).  scale/shift/residual may be NULL. */
void orc_fast_conv_fwd(const float* in, int Cin, const float* W, int K, int Cout, const int* nbr, int no,
                       const float* bias, const float* scale, const float* shift, const float* residual, int relu,
                       float* out) {
  float* Wt = (float*)malloc(sizeof(float) * (size_t)K * Cin * Cout);
  for (int co = 0; co < Cout; ++co)
    for (int k = 0; k < K; ++k)
      for (int ci = 0; ci < Cin; ++ci) Wt[((size_t)k * Cin + ci) * Cout + co] = W[((size_t)co * K + k) * Cin + ci];
#pragma omp parallel for schedule(dynamic, 256)
  for (int o = 0; o < no; ++o) {
    float acc[256];
    for (int co = 0; co < Cout; ++co) acc[co] = bias ? bias[co] : 0.0f;
    for (int k = 0; k < K; ++k) {
      int i = nbr[(size_t)k * no + o];
      if (i < 0) continue;
      const float* a = in + (size_t)i * Cin;
      const float* wk = Wt + (size_t)k * Cin * Cout;
      for (int ci = 0; ci < Cin; ++ci) {
        const float av = a[ci];
        const float* w = wk + (size_t)ci * Cout;
#pragma omp simd
        for (int co = 0; co < Cout; ++co) acc[co] += av * w[co];
      }
    }
    float* dst = out + (size_t)o * Cout;
    for (int co = 0; co < Cout; ++co) {
      float v = acc[co];
      if (scale) v = v * scale[co] + shift[co];
      if (residual) v += residual[(size_t)o * Cout + co];
      if (relu && v < 0.0f) v = 0.0f;
      dst[co] = v;
    }
  }
  free(Wt);
}
