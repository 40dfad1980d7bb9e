// Micro-test of tcgen05.mma.cta_group::2 (kind::f16, bf16 operands, no-swizzle K-major shared-memory operands):
// a CTA pair computes D (256 x N) = A (256 x K) B^T (N x K).  CTA r holds rows [128 r, 128 r + 128) of A and
// rows [N/2 r, N/2 r + N/2) of B in ITS shared memory (same offsets in both CTAs), the leader CTA issues the
// MMAs and commits with .multicast::cluster to both CTAs, and each CTA reads its 128 rows of D from its own
// TMEM.  Checks the result against the host and reports what works.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"
using namespace hn;
namespace hn { void set_error(const char*, ...) {} void count_launch(int) {} int fail_arg(const char*, ...) { return -1; } int check_cuda(cudaError_t e, const char*) { return (int)e; } }

constexpr int K = 64;

__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(int N, float* D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar_mem;
  const uint32_t r = ctarank();
  const int nh = N / 2;
  const uint32_t s_a = smem_u32(smem), s_b = s_a + 128 * K * 2;
  const uint32_t bar = smem_u32(&bar_mem);
  // A_r[m][k] = ((m + 128 r) % 7 - 3) * 0.25 + (k % 5) * 0.125 ; B_r[n][k] = ((n + nh r) % 5 - 2) * 0.5 - (k % 3) * 0.25
  for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) {
    const int m = i / K, kk = i % K;
    const float v = (float)(((m + 128 * (int)r) % 7) - 3) * 0.25f + (float)(kk % 5) * 0.125f;
    *reinterpret_cast<__nv_bfloat16*>(smem + (kk / 8) * (128 * 16) + m * 16 + (kk % 8) * 2) = __float2bfloat16(v);
  }
  for (int i = threadIdx.x; i < nh * K; i += blockDim.x) {
    const int n = i / K, kk = i % K;
    const float v = (float)(((n + nh * (int)r) % 5) - 2) * 0.5f - (float)(kk % 3) * 0.25f;
    *reinterpret_cast<__nv_bfloat16*>(smem + 128 * K * 2 + (kk / 8) * (nh * 16) + n * 16 + (kk % 8) * 2) = __float2bfloat16(v);
  }
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (r == 0 && threadIdx.x == 0) {
    // instruction descriptor: D = F32, A = B = BF16, K-major, N, M = 256
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((256u >> 4) << 24);
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint32_t a_lo = (((s_a + ks * 2 * (128 * 16)) & 0x3FFFFu) >> 4) | (((128u * 16u) >> 4) << 16);
      const uint32_t b_lo = (((s_b + ks * 2 * (nh * 16)) & 0x3FFFFu) >> 4) | ((((uint32_t)nh * 16u) >> 4) << 16);
      const uint64_t da = ((uint64_t)desc_hi << 32) | a_lo, db = ((uint64_t)desc_hi << 32) | b_lo;
      const uint32_t acc = ks > 0;
      asm volatile(
          "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
          "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
  }
  mbar_wait(bar, 0, 7);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(size_t)(r * 128 + warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
  }
}

static float bf(float x) { __nv_bfloat16 b = __float2bfloat16(x); return __bfloat162float(b); }

int main() {
  for (int N : {256, 128, 32, 16}) {
    float* D; cudaMallocManaged(&D, 256 * N * sizeof(float));
    for (int i = 0; i < 256 * N; ++i) D[i] = -12345.f;
    const int smem = 128 * K * 2 + (N / 2) * K * 2;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<<<2, 128, smem>>>(N, D);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d: error %s\n", N, cudaGetErrorString(e)); return 1; }
    double worst = 0; int bad = 0;
    for (int m = 0; m < 256; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int kk = 0; kk < K; ++kk)
          ref += (double)bf((float)((m % 7) - 3) * 0.25f + (float)(kk % 5) * 0.125f) * (double)bf((float)((n % 5) - 2) * 0.5f - (float)(kk % 3) * 0.25f);
        const double err = fabs(ref - D[m * N + n]);
        if (err > worst) worst = err;
        if (err > 1e-3) ++bad;
      }
    printf("cta_group::2 M=256 N=%3d K=%d: max abs err %.3e, %d of %d wrong; D[0][0]=%.4f D[200][N-1]=%.4f\n", N, K, worst, bad, 256 * N, D[0],
           D[200 * N + N - 1]);
    cudaFree(D);
  }
  return 0;
}
