// Process-level pieces of the C ABI: version, last error, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace hn {

static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int fail_arg(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return -1;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return static_cast<int>(e);
}

}  // namespace hn

extern "C" {

int hn_version(void) { return HN_ABI_VERSION; }

const char* hn_last_error(void) { return hn::g_error; }

int64_t hn_launch_count(void) { return hn::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
