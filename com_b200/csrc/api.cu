// Library-level entry points and error plumbing of libcomb200.
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include "common.cuh"

namespace comb {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return COMB_OK;
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return COMB_ECUDA;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace comb

extern "C" {

int comb_version(void) { return 100; }

long long comb_launch_count(void) { return comb::g_launches.load(std::memory_order_relaxed); }

const char* comb_last_error(void) { return comb::g_err; }

int comb_sm_count(void) {
  int dev = 0, n = 0;
  COMB_CUDA(cudaGetDevice(&dev));
  COMB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  return n;
}

}  // extern "C"
