// Micro-benchmark: TMEM read throughput.  W warps (each restricted to its lane quarter) read all 512 columns with
// tcgen05.ld 32x32b.x32 / .x8, ITERS times; reports bytes per clock per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -I humannerf_b200/csrc scripts/micro/ldtm_rate.cu
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"
using namespace hn;
namespace hn { void set_error(const char*, ...) {} void count_launch(int) {} int fail_arg(const char*, ...) { return -1; } int check_cuda(cudaError_t e, const char*) { return (int)e; } }

template <int X>
__global__ void __launch_bounds__(512, 1) k(int iters, int warps, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  const int warp = threadIdx.x >> 5;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < warps) {
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    for (int i = 0; i < iters; ++i) {
      for (int c = 0; c < 512; c += X) {
        if (X == 32) { uint32_t v[32]; tmem_ld32(base + c, v); tmem_ld_wait(); acc += v[0] ^ v[31]; }
        else { uint32_t v[16]; tmem_ld16(base + c, v); tmem_ld_wait(); acc += v[0] ^ v[15]; }
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* out; cudaMallocManaged(&out, 64);
  uint32_t* sink; cudaMalloc(&sink, 148 * 512 * 4);
  const int iters = 200;
  for (int warps : {1, 4, 8, 16}) {
    k<32><<<148, 512>>>(iters, warps, out, sink); cudaDeviceSynchronize();
    double bytes = (double)warps * iters * 512 * 32 * 4;
    printf("x32, %2d warps: %.1f B/clk/SM (%.0f clk per 128x256 fp32 accumulator)\n", warps, bytes / out[0], 131072.0 / (bytes / out[0]));
    k<16><<<148, 512>>>(iters, warps, out, sink); cudaDeviceSynchronize();
    printf("x16, %2d warps: %.1f B/clk/SM\n", warps, bytes / out[0]);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
