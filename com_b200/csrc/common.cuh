// Shared helpers for libcomb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/comb200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libcomb200 is written for sm_100a (B200) only"
#endif

namespace comb {

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
void count_launch();

#define COMB_CHECK_ARG(cond, ...)      \
  do {                                 \
    if (!(cond)) {                     \
      comb::set_error(__VA_ARGS__);    \
      return COMB_EINVAL;              \
    }                                  \
  } while (0)

#define COMB_CUDA(call)                                  \
  do {                                                   \
    int _rc = comb::check_cuda((call), #call);           \
    if (_rc != COMB_OK) return _rc;                      \
  } while (0)

// every kernel launch of the library is followed by exactly one COMB_LAUNCH_CHECK: it also feeds the
// launch counter behind comb_launch_count() (bench.py's "gpu_launches").
#define COMB_LAUNCH_CHECK(name)           \
  do {                                    \
    comb::count_launch();                 \
    COMB_CUDA(cudaPeekAtLastError());     \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// One-time-per-DEVICE guard for cudaFuncSetAttribute (a per-device property): `static thread_local DevOnce once;
// if (once.first()) cudaFuncSetAttribute(...)`.  A process that drives several GPUs from one thread sets it on each.
struct DevOnce {
  unsigned long long seen[2] = {0ull, 0ull};
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    if (d < 0 || d >= 128) return true;
    const unsigned long long bit = 1ull << (d & 63);
    if (seen[d >> 6] & bit) return false;
    seen[d >> 6] |= bit;
    return true;
  }
};
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int sm_count();

// Effective row count: device-side count (if given) clamped to the launch bound.
__device__ __forceinline__ int eff_n(int n_max, const int* n_dev) {
  if (n_dev == nullptr) return n_max;
  int n = __ldg(n_dev);
  return n < n_max ? n : n_max;
}

constexpr uint32_t kEmptyKey = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t hash_u32(uint32_t k) {
  // murmur3 finaliser: cheap, good avalanche for linear voxel keys
  k ^= k >> 16;
  k *= 0x85ebca6bu;
  k ^= k >> 13;
  k *= 0xc2b2ae35u;
  k ^= k >> 16;
  return k;
}

// Open-addressing table of (key, value) pairs packed in one 8-byte slot.
__device__ __forceinline__ int hash_lookup(const uint2* __restrict__ table, uint32_t mask, uint32_t key) {
  uint32_t s = hash_u32(key) & mask;
  while (true) {
    uint2 kv = __ldg(table + s);
    if (kv.x == key) return (int)kv.y;
    if (kv.x == kEmptyKey) return -1;
    s = (s + 1) & mask;
  }
}

}  // namespace comb
