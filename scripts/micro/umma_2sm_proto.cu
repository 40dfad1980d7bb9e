// Micro-test of the cross-CTA protocol the cta_group::2 field kernel (mlp_tc_2sm.cu) relies on:
//   * tcgen05.alloc.cta_group::2 of all 512 columns, executed collectively by the same warp of both CTAs;
//   * the B operand halves arrive by 1-D bulk copies, each CTA fetching ITS half into ITS ring slot;
//     MODE 0: the peer's copy completes on its own barrier and a relay thread forwards a remote arrive to the
//             leader; MODE 1: the peer's copy names the LEADER's barrier as its completion barrier
//             (shared::cluster address from mapa) and the leader expects the bytes of both halves;
//   * the A operand is written by generic stores of 4 warps per CTA, fence.proxy.async, and ONE
//     mbarrier.arrive.release.cluster per warp on the leader's barrier (remote for the peer);
//   * the leader issues M = 256 MMAs and commits with .multicast::cluster to both CTAs;
//   * ROUNDS rounds with different data check the phase bookkeeping.
// nvcc -gencode arch=compute_100a,code=sm_100a -I humannerf_b200/csrc scripts/micro/umma_2sm_proto.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"
using namespace hn;
namespace hn { void set_error(const char*, ...) {} void count_launch(int) {} int fail_arg(const char*, ...) { return -1; } int check_cuda(cudaError_t e, const char*) { return (int)e; } }

constexpr int K = 64, N = 256, NH = N / 2, ROUNDS = 5;

__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void wait_cluster(uint32_t bar, uint32_t parity, int who) {
  const long long t0 = clock64();
  while (!try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 2000000000LL) { printf("timeout who %d block %d thread %d\n", who, (int)blockIdx.x, (int)threadIdx.x); __trap(); }
  }
}

// Bimg: per round, [rank][k/8][n'][8] bf16 (the split image); a(m, k, round) is computed on the fly
__device__ __host__ inline float a_val(int m, int k, int round) { return (float)(((m + round) % 7) - 3) * 0.25f + (float)(k % 5) * 0.125f; }
__device__ __host__ inline float b_val(int n, int k, int round) { return (float)(((n + 2 * round) % 5) - 2) * 0.5f - (float)(k % 3) * 0.25f; }

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(160, 1) k(const uint8_t* __restrict__ Bimg, float* D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bars[4];       // 0 full (leader), 1 full_local (peer, MODE 0), 2 a_ready (leader), 3 acc (both)
  const uint32_t r = ctarank();
  const uint32_t s_a = smem_u32(smem), s_b = s_a + 128 * K * 2;
  const uint32_t b_full = smem_u32(&bars[0]), b_local = smem_u32(&bars[1]), b_ready = smem_u32(&bars[2]), b_acc = smem_u32(&bars[3]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(b_full, MODE == 0 ? 2 : 1);
    mbar_init(b_local, 1);
    mbar_init(b_ready, 8);
    mbar_init(b_acc, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t half_bytes = NH * K * 2;

  for (int round = 0; round < ROUNDS; ++round) {
    const uint32_t ph = round & 1;
    const uint32_t d_col = (round & 1) * 256u;           // alternate the two accumulators
    if (warp == 4) {
      // ---- producer + issuer warp ----
      if (lane == 0) {
        if (round > 0) mbar_wait(b_acc, (round - 1) & 1, 5);   // the previous round's MMAs no longer read the B slot
        const uint8_t* src = Bimg + (size_t)round * 2 * half_bytes + (size_t)r * half_bytes;
        if (MODE == 0) {
          if (r == 0) { mbar_expect_tx(b_full, half_bytes); tma_bulk_g2s(s_b, src, half_bytes, b_full); }
          else { mbar_expect_tx(b_local, half_bytes); tma_bulk_g2s(s_b, src, half_bytes, b_local); }
        } else {
          if (r == 0) mbar_expect_tx(b_full, 2 * half_bytes);
          tma_bulk_g2s(s_b, src, half_bytes, mapa(b_full, 0));
        }
      }
      if (MODE == 0 && r == 1 && lane == 1) {            // relay
        mbar_wait(b_local, ph, 1);
        arrive_cluster(mapa(b_full, 0));
      }
      if (r == 0 && lane == 0) {
        wait_cluster(b_ready, ph, 2);
        wait_cluster(b_full, ph, 3);
        tc_fence_after();
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((256u >> 4) << 24);
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        for (int ks = 0; ks < K / 16; ++ks) {
          const uint32_t a_lo = (((s_a + ks * 2 * (128 * 16)) & 0x3FFFFu) >> 4) | (((128u * 16u) >> 4) << 16);
          const uint32_t b_lo = (((s_b + ks * 2 * (NH * 16)) & 0x3FFFFu) >> 4) | ((((uint32_t)NH * 16u) >> 4) << 16);
          const uint64_t da = ((uint64_t)desc_hi << 32) | a_lo, db = ((uint64_t)desc_hi << 32) | b_lo;
          const uint32_t acc = ks > 0;
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                       ::"r"(tmem + d_col), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(b_acc), "h"((uint16_t)3) : "memory");
      }
    } else {
      // ---- 4 "epilogue" warps: write A (row = thread), hand over, then read D ----
      const int m = warp * 32 + lane;
      for (int kg = 0; kg < K / 8; ++kg) {
        float v[8];
        for (int i = 0; i < 8; ++i) v[i] = a_val(m + 128 * (int)r, kg * 8 + i, round);
        st_shared_v4(s_a + kg * (128 * 16) + m * 16, pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
      }
      asm volatile("fence.proxy.async;" ::: "memory");
      __syncwarp();
      if (lane == 0) arrive_cluster(mapa(b_ready, 0));
      mbar_wait(b_acc, ph, 4);
      tc_fence_after();
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + d_col + c0, v);
        tmem_ld_wait();
        if (round == ROUNDS - 1)
          for (int j = 0; j < 16; ++j) D[(size_t)(r * 128 + m) * N + c0 + j] = __uint_as_float(v[j]);
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

static float bf(float x) { __nv_bfloat16 b = __float2bfloat16(x); return __bfloat162float(b); }

template <int MODE> int run(const uint8_t* Bimg) {
  float* D; cudaMallocManaged(&D, 256 * N * sizeof(float));
  for (int i = 0; i < 256 * N; ++i) D[i] = -12345.f;
  const int smem = 128 * K * 2 + NH * K * 2;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<MODE><<<2, 160, smem>>>(Bimg, D);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("MODE %d: error %s\n", MODE, cudaGetErrorString(e)); return 1; }
  const int round = ROUNDS - 1;
  double worst = 0; int bad = 0;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int kk = 0; kk < K; ++kk) ref += (double)bf(a_val(m, kk, round)) * (double)bf(b_val(n, kk, round));
      const double err = fabs(ref - D[m * N + n]);
      if (err > worst) worst = err;
      if (err > 1e-3) ++bad;
    }
  printf("MODE %d (%s): max abs err %.3e, %d of %d wrong\n", MODE, MODE == 0 ? "relay" : "cross-CTA complete_tx", worst, bad, 256 * N);
  cudaFree(D);
  return bad != 0;
}

int main() {
  const size_t half = (size_t)NH * K * 2;
  std::vector<__nv_bfloat16> h(ROUNDS * 2 * half / 2);
  for (int round = 0; round < ROUNDS; ++round)
    for (int r = 0; r < 2; ++r)
      for (int kg = 0; kg < K / 8; ++kg)
        for (int n = 0; n < NH; ++n)
          for (int i = 0; i < 8; ++i)
            h[((size_t)round * 2 + r) * (half / 2) + ((size_t)kg * NH + n) * 8 + i] = __float2bfloat16(b_val(n + NH * r, kg * 8 + i, round));
  uint8_t* Bimg; cudaMalloc(&Bimg, h.size() * 2);
  cudaMemcpy(Bimg, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  int rc = run<0>(Bimg);
  rc |= run<1>(Bimg);
  return rc;
}
