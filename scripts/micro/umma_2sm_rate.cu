// Micro-benchmark: cycles per tcgen05.mma.cta_group::2 (kind::f16, M = 256, K = 16) for a given N, issued by ISSUERS
// warps of the leader CTA in the warp-uniform style of mlp_tc_2sm.cu (whole warp runs the loop, one elected lane
// issues), chunks of PER MMAs each followed by a multicast commit, optionally (WAIT) preceded by a try_wait on a
// completed barrier as the kernel's `full` probe.
// nvcc -gencode arch=compute_100a,code=sm_100a -I humannerf_b200/csrc scripts/micro/umma_2sm_rate.cu
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"
using namespace hn;
namespace hn { void set_error(const char*, ...) {} void count_launch(int) {} int fail_arg(const char*, ...) { return -1; } int check_cuda(cudaError_t e, const char*) { return (int)e; } }

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred;
}

template <int PER, bool WAIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(int n, int chunks, int issuers, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t s_base = smem_u32(smem);
  const uint32_t bar = s_base + 200 * 1024, slot = bar + 128;
  volatile uint32_t* tslot = reinterpret_cast<volatile uint32_t*>(smem + 200 * 1024 + 128);
  for (int i = threadIdx.x; i < 50 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int b = 0; b < 8; ++b) mbar_init(bar + 8 * b, b < 2 ? 1 : 1000000); fence_barrier_init(); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
  const uint32_t tmem = *tslot;
  const int w = threadIdx.x >> 5;
  if (cluster_ctarank() == 0 && w < issuers) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)n >> 3) << 17) | ((256u >> 4) << 24);
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lo0 = ((s_base & 0x3FFFFu) >> 4) | ((2048u >> 4) << 16);
    const uint32_t b_lo0 = (((s_base + 65536u) & 0x3FFFFu) >> 4) | ((uint32_t)(n / 2) << 16);
    const long long t0 = clock64();
    for (int c = 0; c < chunks; ++c) {
      if (WAIT) mbar_wait(bar + 16 + 8 * w, 1, 1);      // initial phase: parity 1 passes at once
      if (WAIT) tc_fence_after();
      uint32_t a_lo = a_lo0, b_lo = b_lo0;
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < PER; ++i) {
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                       ::"r"(tmem + w * 256), "l"(((uint64_t)desc_hi << 32) | a_lo), "l"(((uint64_t)desc_hi << 32) | b_lo), "r"(idesc), "r"((uint32_t)(c + i > 0)) : "memory");
          a_lo += (4096u >> 4); b_lo += (uint32_t)n;
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar + 32 + 8 * w), "h"((uint16_t)3) : "memory");
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar + 8 * w), "h"((uint16_t)1) : "memory");
    __syncwarp();
    mbar_wait(bar + 8 * w, 0, 9);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { out[w * 2] = t1 - t0; out[w * 2 + 1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory"); }
}

template <int PER, bool WAIT> void run(long long* out, int n, int issuers) {
  const int chunks = 2048;
  cudaFuncSetAttribute(k<PER, WAIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
  k<PER, WAIT><<<148, 128, 201 * 1024>>>(n, chunks, issuers, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  const double mmas = (double)chunks * PER * issuers;
  printf("cta_group::2 N %3d, %d MMAs/chunk, probe %d, %d issuer(s): issue done after %.1f clk/MMA, complete %.1f clk/MMA (pipe, all issuers)\n",
         n, PER, (int)WAIT, issuers, (double)out[0] * issuers / mmas, (double)out[1] / mmas);
}

int main() {
  long long* out; cudaMallocManaged(&out, 64);
  for (int n : {256, 128}) {
    run<16, false>(out, n, 1);
    run<2, false>(out, n, 1); run<2, true>(out, n, 1); run<2, true>(out, n, 2);
    run<4, true>(out, n, 1); run<4, true>(out, n, 2);
  }
  return 0;
}
