// Micro-benchmark of the forward epilogue of mlp_tc.cu in isolation: 16 warps (2 tiles x 4 lane quarters x 2
// column halves) turn two 128x256 fp32 accumulators in TMEM into bf16 activations in shared memory
// (tcgen05.ld 32x32b.x32 -> cvt.rn.relu.bf16x2.f32 -> st.shared.v4), ITERS times.  VARIANT bit0: skip the
// conversions (store raw words), bit1: skip the shared-memory stores, bit2: skip the TMEM loads.
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"
using namespace hn;
namespace hn { void set_error(const char*, ...) {} void count_launch(int) {} int fail_arg(const char*, ...) { return -1; } int check_cuda(cudaError_t e, const char*) { return (int)e; } }

template <int VARIANT>
__global__ void __launch_bounds__(512, 1) k(int iters, long long* out, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3, tile = warp >> 3, half = (warp >> 2) & 1;
  const int row = quarter * 32 + lane;
  const uint32_t t_acc = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)tile * 256u;
  const uint32_t a_h = smem_u32(smem) + (uint32_t)tile * 65536u;
  uint32_t accx = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int c_begin = half * 128;
    uint32_t v[2][32];
    if (!(VARIANT & 4)) tmem_ld32(t_acc + (uint32_t)c_begin, v[0]);
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 64) {
      if (!(VARIANT & 4)) tmem_ld_wait();
      if (!(VARIANT & 4)) tmem_ld32(t_acc + (uint32_t)(c_begin + c0 + 32), v[1]);
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        if (b == 1) { if (!(VARIANT & 4)) { tmem_ld_wait(); if (c0 + 64 < 128) tmem_ld32(t_acc + (uint32_t)(c_begin + c0 + 64), v[0]); } }
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const uint32_t* u = v[b] + g8 * 8;
          const int chunk = ((c_begin + c0 + 32 * b) >> 3) + g8;
          uint4 pk;
          if (VARIANT & 1) pk = make_uint4(u[0] ^ u[1], u[2] ^ u[3], u[4] ^ u[5], u[6] ^ u[7]);
          else pk = make_uint4(pack_relu_bf16(__uint_as_float(u[0]), __uint_as_float(u[1])), pack_relu_bf16(__uint_as_float(u[2]), __uint_as_float(u[3])),
                               pack_relu_bf16(__uint_as_float(u[4]), __uint_as_float(u[5])), pack_relu_bf16(__uint_as_float(u[6]), __uint_as_float(u[7])));
          if (VARIANT & 2) accx += pk.x ^ pk.y ^ pk.z ^ pk.w;
          else st_shared_v4(a_h + (uint32_t)chunk * 2048u + (uint32_t)row * 16u, pk.x, pk.y, pk.z, pk.w);
        }
      }
    }
    tc_fence_before();
    fence_proxy_async();
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = accx;
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int V> void run(long long* out, uint32_t* sink) {
  const int iters = 500;
  cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  k<V><<<148, 512, 131072>>>(iters, out, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  printf("no-cvt %d no-sts %d no-ldtm %d: %.0f clk per pair of 128x256 tiles\n", V & 1, (V >> 1) & 1, (V >> 2) & 1, (double)out[0] / iters);
}
int main() {
  long long* out; cudaMallocManaged(&out, 64);
  uint32_t* sink; cudaMalloc(&sink, 148 * 512 * 4);
  run<0>(out, sink); run<1>(out, sink); run<2>(out, sink); run<3>(out, sink); run<4>(out, sink); run<6>(out, sink);
  return 0;
}
