// Shared helpers for the humannerf_b200 CUDA sources (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/humannerf_b200.h"

namespace hn {

// thread-local last-error text + process-wide launch counter (api.cu)
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int fail_arg(const char* fmt, ...);          // sets the message, returns -1
int check_cuda(cudaError_t e, const char* what);   // 0 or the cudaError_t (message set)

// call after every kernel launch
inline int launched(const char* what) {
  count_launch();
  return check_cuda(cudaGetLastError(), what);
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

constexpr int kMaxJoints = 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace hn

#define HN_TRY(expr)            \
  do {                          \
    int _hn_rc = (expr);        \
    if (_hn_rc != 0) return _hn_rc; \
  } while (0)
