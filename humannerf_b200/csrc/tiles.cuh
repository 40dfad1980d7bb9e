// T16 tile image: the bf16 activation / gradient layout shared by the tcgen05 kernels
// (definition in tile_wgrad.cu and include/humannerf_b200.h).
#pragma once

#include "common.cuh"

namespace hn {

using WgradJob = hn_wgrad_job;

// byte offset of element (sample s, feature f) in an image with f_pad features per sample
__host__ __device__ __forceinline__ size_t tile_offset(int64_t s, int f, int f_pad) {
  const int64_t tile = s >> 7;
  const int r = (int)(s & 127);
  return (size_t)tile * 128 * f_pad * 2 + (size_t)(r >> 4) * 16 * f_pad * 2 + (size_t)(f >> 3) * 256 +
         (size_t)(r & 15) * 16 + (size_t)(f & 7) * 2;
}

constexpr int kMaxWgradJobs = 48;   // per launch: the table is a kernel parameter (48 x 72 B < 4 KB)

// jobs_host: HOST array of n_jobs <= kMaxWgradJobs jobs
// n_samples_dev (nullable): device-side count of the samples actually present (<= n_samples)
int launch_tile_wgrad(const WgradJob* jobs_host, int n_jobs, int64_t n_samples, const int64_t* n_samples_dev,
                      cudaStream_t st);

}  // namespace hn
