// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128, K=16) for a given N when one
// thread issues them back to back, with WORK dependent integer operations and optionally a commit between
// consecutive MMAs, and with one or two issuing threads (separate accumulators).
// nvcc -gencode arch=compute_100a,code=sm_100a -I humannerf_b200/csrc scripts/micro/umma_rate.cu
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"
using namespace hn;
namespace hn { void set_error(const char*, ...) {} void count_launch(int) {} int fail_arg(const char*, ...) { return -1; } int check_cuda(cudaError_t e, const char*) { return (int)e; } }

template <int WORK, int PER_COMMIT>
__global__ void __launch_bounds__(128, 1) k(int n, int iters, int issuers, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t s_base = smem_u32(smem);
  const uint32_t bar = s_base + 200 * 1024, slot = bar + 64;
  volatile uint32_t* tslot = reinterpret_cast<volatile uint32_t*>(smem + 200 * 1024 + 64);
  for (int i = threadIdx.x; i < 50 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int b = 0; b < 4; ++b) mbar_init(bar + 8 * b, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *tslot;
  fence_proxy_async();
  __syncthreads();
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < issuers) {
    const uint32_t idesc = make_idesc(n);
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lo0 = ((s_base & 0x3FFFFu) >> 4) | ((2048u >> 4) << 16);
    const uint32_t b_lo0 = (((s_base + 65536u) & 0x3FFFFu) >> 4) | ((uint32_t)n << 16);
    const long long t0 = clock64();
    uint32_t a_lo = a_lo0, b_lo = b_lo0, dummy = (uint32_t)iters;
    for (int i = 0; i < iters; ++i) {
      umma_bf16(tmem + w * 256, ((uint64_t)desc_hi << 32) | a_lo, ((uint64_t)desc_hi << 32) | b_lo, idesc, i > 0);
      a_lo += (4096u >> 4); b_lo += 2u * (uint32_t)n;
      if ((i & 15) == 15) { a_lo = a_lo0; b_lo = b_lo0; }
#pragma unroll
      for (int j = 0; j < WORK; ++j) asm volatile("mad.lo.u32 %0, %0, %0, %0;" : "+r"(dummy));
      if (PER_COMMIT > 0 && (i % PER_COMMIT) == PER_COMMIT - 1) umma_commit(bar + 16 + 8 * w);
    }
    const long long t1 = clock64();
    umma_commit(bar + 8 * w);
    mbar_wait(bar + 8 * w, 0, 9);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[w * 2] = t1 - t0; out[w * 2 + 1] = t2 - t0 + (dummy == 12345u); }
  }
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// cluster of two CTAs: commit with .multicast::cluster to the same barrier offset in both CTAs
template <int PER_COMMIT, bool kMulticast>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k2(int n, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t s_base = smem_u32(smem);
  const uint32_t bar = s_base + 200 * 1024, slot = bar + 64;
  volatile uint32_t* tslot = reinterpret_cast<volatile uint32_t*>(smem + 200 * 1024 + 64);
  for (int i = threadIdx.x; i < 50 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 8, kMulticast ? 2 : 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  tc_fence_after();
  const uint32_t tmem = *tslot;
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(n);
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lo0 = ((s_base & 0x3FFFFu) >> 4) | ((2048u >> 4) << 16);
    const uint32_t b_lo0 = (((s_base + 65536u) & 0x3FFFFu) >> 4) | ((uint32_t)n << 16);
    const long long t0 = clock64();
    uint32_t a_lo = a_lo0, b_lo = b_lo0;
    for (int i = 0; i < iters; ++i) {
      umma_bf16(tmem, ((uint64_t)desc_hi << 32) | a_lo, ((uint64_t)desc_hi << 32) | b_lo, idesc, i > 0);
      a_lo += (4096u >> 4); b_lo += 2u * (uint32_t)n;
      if ((i & 15) == 15) { a_lo = a_lo0; b_lo = b_lo0; }
      if ((i % PER_COMMIT) == PER_COMMIT - 1) {
        if (kMulticast)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar + 8), "h"((uint16_t)3) : "memory");
        else
          umma_commit(bar + 8);
      }
    }
    const long long t1 = clock64();
    umma_commit(bar);
    mbar_wait(bar, 0, 9);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int PC, bool MC>
void run2(long long* out, int n) {
  const int iters = 4096;
  cudaFuncSetAttribute(k2<PC, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
  k2<PC, MC><<<148, 128, 201 * 1024>>>(n, iters, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  printf("cluster2 N %3d commit/%d %s: issue %.1f clk/MMA, complete %.1f clk/MMA\n", n, PC, MC ? "multicast" : "local    ",
         (double)out[0] / iters, (double)out[1] / iters);
}

// the issuer loop of the field kernels in miniature: per chunk [VARIANT bit0: try_wait on a completed barrier x2]
// [bit1: tcgen05.fence::after_thread_sync] 2 MMAs, 1 commit [bit2: second commit]
template <int VARIANT>
__global__ void __launch_bounds__(640, 1) k3(int n, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t s_base = smem_u32(smem);
  const uint32_t bar = s_base + 200 * 1024, slot = bar + 64;
  volatile uint32_t* tslot = reinterpret_cast<volatile uint32_t*>(smem + 200 * 1024 + 64);
  for (int i = threadIdx.x; i < 50 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  volatile uint32_t* stop = reinterpret_cast<volatile uint32_t*>(smem + 200 * 1024 + 256);
  if (threadIdx.x == 0) { for (int b = 0; b < 8; ++b) mbar_init(bar + 8 * b, 1); *stop = 0; fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *tslot;
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x >= 128) {
    // VARIANT bit 5: 16 warps poll a barrier that never completes (like epilogue warps waiting for an
    // accumulator); bit 6: the same with a suspend-time hint
    if (VARIANT & 32) {
      while (*stop == 0) {
        uint32_t ok;
        if (VARIANT & 64)
          asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar + 56), "r"(0u), "r"(20000u) : "memory");
        else
          ok = mbar_try_wait(bar + 56, 0);
        if (ok) break;
      }
    }
  } else if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(n);
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lo0 = ((s_base & 0x3FFFFu) >> 4) | ((2048u >> 4) << 16);
    const uint32_t b_lo0 = (((s_base + 65536u) & 0x3FFFFu) >> 4) | ((uint32_t)n << 16);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if ((VARIANT & 1) && !(VARIANT & 8)) {
        // barriers 2 and 3 are in their initial phase: waiting for parity 1 returns at once
        mbar_wait(bar + 16, 1, 1);
        mbar_wait(bar + 24, 1, 1);
      }
      if (VARIANT & 8) {          // non-blocking test_wait instead of try_wait
        uint32_t ok1, ok2;
        asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok1) : "r"(bar + 16), "r"(1u) : "memory");
        asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok2) : "r"(bar + 24), "r"(1u) : "memory");
        if (!(ok1 & ok2)) break;
      }
      if (VARIANT & 16) {         // plain volatile shared-memory flag reads instead of mbarrier waits
        volatile uint32_t* f = reinterpret_cast<volatile uint32_t*>(smem + 200 * 1024 + 128);
        if (f[0] == 0xdeadbeefu || f[1] == 0xdeadbeefu) break;
      }
      if (VARIANT & 2) tc_fence_after();
      umma_bf16(tmem, ((uint64_t)desc_hi << 32) | a_lo0, ((uint64_t)desc_hi << 32) | b_lo0, idesc, i > 0);
      umma_bf16(tmem, ((uint64_t)desc_hi << 32) | (a_lo0 + 256), ((uint64_t)desc_hi << 32) | (b_lo0 + 2 * n), idesc, 1);
      umma_commit(bar + 32);
      if (VARIANT & 4) umma_commit(bar + 40);
    }
    const long long t1 = clock64();
    umma_commit(bar);
    mbar_wait(bar, 0, 9);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    *stop = 1;
  }
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}
template <int V>
void run3(long long* out, int n) {
  const int iters = 2048;
  cudaFuncSetAttribute(k3<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
  k3<V><<<148, 640, 201 * 1024>>>(n, iters, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  printf("pollers %d hint %d | chunk loop N %3d waits %d fence %d commits %d: %.1f clk per 2-MMA chunk issued, %.1f complete\n", (V >> 5) & 1, (V >> 6) & 1, n, V & 1, (V >> 1) & 1,
         1 + ((V >> 2) & 1), (double)out[0] / iters, (double)out[1] / iters);
}

template <int WORK, int PC>
void run(long long* out, int n, int issuers) {
  const int iters = 4096;
  cudaFuncSetAttribute(k<WORK, PC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
  k<WORK, PC><<<148, 128, 201 * 1024>>>(n, iters, issuers, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  printf("issuers %d N %3d work %3d commit/%d: issue %.1f clk/MMA, complete %.1f clk/MMA", issuers, n, WORK, PC,
         (double)out[0] / iters, (double)out[1] / iters);
  if (issuers == 2) printf("  (2nd: %.1f / %.1f)", (double)out[2] / iters, (double)out[3] / iters);
  printf("\n");
}

int main() {
  long long* out; cudaMallocManaged(&out, 64);
  for (int n : {128, 256}) { run3<0>(out, n); run3<1>(out, n); run3<2>(out, n); run3<3>(out, n); run3<32>(out, n); run3<35>(out, n); run3<96>(out, n); run3<99>(out, n); }
  for (int n : {128, 256}) {
    run2<1, false>(out, n); run2<2, false>(out, n); run2<1, true>(out, n); run2<2, true>(out, n); run2<4, true>(out, n); run2<8, true>(out, n);
  }
  for (int n : {128, 256}) {
    run<0, 0>(out, n, 1); run<8, 0>(out, n, 1); run<16, 0>(out, n, 1); run<32, 0>(out, n, 1); run<64, 0>(out, n, 1);
    run<0, 1>(out, n, 1); run<0, 2>(out, n, 1); run<0, 4>(out, n, 1); run<16, 2>(out, n, 1);
    run<0, 0>(out, n, 2); run<32, 0>(out, n, 2); run<0, 2>(out, n, 2); run<32, 2>(out, n, 2);
  }
  return 0;
}
