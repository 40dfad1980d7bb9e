// Micro-test + micro-benchmark: tcgen05.mma.cta_group::2 with the A operand in TENSOR MEMORY ("TS" form).
//   1. correctness: A (256 rows = 128 per CTA, K = 64, small integers, bf16) is written to TMEM with tcgen05.st
//      (lane = row, one 32-bit column = two consecutive K elements) and, for comparison, to shared memory in the
//      K-major core-matrix layout of the production kernels; B (N = 128, half per CTA) sits in shared memory.
//      D_ts = A_tmem x B^T and D_ss = A_smem x B^T are read back with tcgen05.ld and compared with the exact integer
//      result -> verifies the TMEM layout of A, the instruction form and the st -> fence -> mma ordering.
//   2. rate: clocks per TS MMA (M = 256, K = 16) for N = 128 and N = 256, 64 back-to-back MMAs issued by one
//      elected lane, against the SS form.
// nvcc -gencode arch=compute_100a,code=sm_100a -I humannerf_b200/csrc scripts/micro/umma_ts.cu -o /tmp/umma_ts
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"
using namespace hn;
namespace hn { void set_error(const char*, ...) {} void count_launch(int) {} int fail_arg(const char*, ...) { return -1; } int check_cuda(cudaError_t e, const char*) { return (int)e; } }

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void mma2_ts(uint32_t d, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
               ::"r"(d), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma2_ss(uint32_t d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit2(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__host__ __device__ inline int a_val(int r, int k) { return ((r * 7 + k * 3) % 13) - 6; }
__host__ __device__ inline int b_val(int n, int k) { return ((n * 5 + k) % 11) - 5; }
__device__ __forceinline__ uint16_t bf16_bits(int v) { return (uint16_t)(__float_as_uint((float)v) >> 16); }   // small ints: exact

constexpr int K = 64, N = 128;
constexpr uint32_t kColsA = 0, kColsDts = 128, kColsDss = 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
ts_kernel(float* __restrict__ d_ts, float* __restrict__ d_ss, long long* __restrict__ clk) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_a = s_base, s_b = s_base + 65536, bar = s_base + 200 * 1024, slot = bar + 64;
  volatile uint32_t* tslot = reinterpret_cast<volatile uint32_t*>(smem + 200 * 1024 + 64);
  const uint32_t rank = cluster_ctarank();
  const int row = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // operands in shared memory: A rows of THIS CTA (K-major core matrices, LBO = 2048), B rows [rank * 64, +64)
  for (int k = 0; k < K; ++k) {
    *reinterpret_cast<uint16_t*>(smem + (k / 8) * 2048 + row * 16 + (k % 8) * 2) = bf16_bits(a_val((int)rank * 128 + row, k));
    if (row < N / 2)
      *reinterpret_cast<uint16_t*>(smem + 65536 + (k / 8) * (N / 2 * 16) + row * 16 + (k % 8) * 2) = bf16_bits(b_val((int)rank * (N / 2) + row, k));
  }
  // rate test operand: a full-width B block (N = 256: 128 rows per CTA, 256 K) of ones at +96 KB
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + 98304)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int b = 0; b < 4; ++b) mbar_init(bar + 8 * b, 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
  const uint32_t tmem = *tslot;
  const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
  // A -> tensor memory: lane = row, column c = (k = 2c | k = 2c + 1 << 16)
  for (int c0 = 0; c0 < K / 2; c0 += 8) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = 2 * (c0 + j);
      v[j] = (uint32_t)bf16_bits(a_val((int)rank * 128 + row, k)) | ((uint32_t)bf16_bits(a_val((int)rank * 128 + row, k + 1)) << 16);
    }
    tmem_st8(t_lane + kColsA + (uint32_t)c0, v);
  }
  tmem_st_wait();
  tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();

  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((256u >> 4) << 24);
  const uint32_t desc_hi = (128u >> 4) | (1u << 14);
  if (rank == 0 && warp == 0) {
    const uint32_t a_lo0 = ((s_a & 0x3FFFFu) >> 4) | ((2048u >> 4) << 16);
    const uint32_t b_lo0 = ((s_b & 0x3FFFFu) >> 4) | ((uint32_t)(N / 2) << 16);
    if (elect_one()) {
      for (int j = 0; j < K / 16; ++j) {
        mma2_ts(tmem + kColsDts, tmem + kColsA + 8 * j, ((uint64_t)desc_hi << 32) | (b_lo0 + (uint32_t)j * N), idesc, j > 0);
        mma2_ss(tmem + kColsDss, ((uint64_t)desc_hi << 32) | (a_lo0 + (uint32_t)j * (4096u >> 4)),
                ((uint64_t)desc_hi << 32) | (b_lo0 + (uint32_t)j * N), idesc, j > 0);
      }
      commit2(bar, 3);
    }
    __syncwarp();
  }
  mbar_wait(bar, 0, 1);
  tc_fence_after();
  // read both results back: this thread's row, N columns
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(t_lane + kColsDts + (uint32_t)c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) d_ts[((size_t)rank * 128 + row) * N + c0 + j] = __uint_as_float(v[j]);
    tmem_ld32(t_lane + kColsDss + (uint32_t)c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) d_ss[((size_t)rank * 128 + row) * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();

  // ---- rate: 64 MMAs back to back, TS and SS, N = 128 and N = 256 (B of ones; results are not looked at)
  if (rank == 0 && warp == 0) {
    int slot_i = 0;
    for (int n = 128; n <= 256; n += 128) {
      const uint32_t id = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)n >> 3) << 17) | ((256u >> 4) << 24);
      const uint32_t b_lo0 = (((s_base + 98304u) & 0x3FFFFu) >> 4) | ((uint32_t)(n / 2) << 16);
      const uint32_t a_lo0 = ((s_a & 0x3FFFFu) >> 4) | ((2048u >> 4) << 16);
      for (int ts = 1; ts >= 0; --ts) {
        const long long t0 = clock64();
        if (elect_one()) {
          for (int i = 0; i < 64; ++i) {
            const uint32_t j = (uint32_t)(i & 3);
            if (ts) mma2_ts(tmem + 256, tmem + kColsA + 8 * j, ((uint64_t)desc_hi << 32) | (b_lo0 + j * (uint32_t)n), id, 1);
            else mma2_ss(tmem + 256, ((uint64_t)desc_hi << 32) | (a_lo0 + j * (4096u >> 4)), ((uint64_t)desc_hi << 32) | (b_lo0 + j * (uint32_t)n), id, 1);
          }
          commit2(bar + 8 + 8 * (slot_i & 1), 1);
        }
        __syncwarp();
        const long long t1 = clock64();
        mbar_wait(bar + 8 + 8 * (slot_i & 1), (slot_i >> 1) & 1, 2);
        const long long t2 = clock64();
        if (lane == 0) { clk[slot_i * 2] = t1 - t0; clk[slot_i * 2 + 1] = t2 - t0; }
        ++slot_i;
      }
    }
  }
  tc_fence_before(); __syncthreads(); cluster_sync_all();
  if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory"); }
}

int main() {
  float *d_ts, *d_ss;
  long long* clk;
  cudaMalloc(&d_ts, 256 * N * 4); cudaMalloc(&d_ss, 256 * N * 4); cudaMalloc(&clk, 64 * 8);
  cudaMemset(d_ts, 0xff, 256 * N * 4); cudaMemset(d_ss, 0xff, 256 * N * 4); cudaMemset(clk, 0, 64 * 8);
  cudaFuncSetAttribute(ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
  ts_kernel<<<2, 128, 201 * 1024>>>(d_ts, d_ss, clk);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<float> ts(256 * N), ss(256 * N);
  long long c[16];
  cudaMemcpy(ts.data(), d_ts, ts.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(ss.data(), d_ss, ss.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(c, clk, sizeof(c), cudaMemcpyDeviceToHost);
  int bad_ts = 0, bad_ss = 0;
  for (int r = 0; r < 256; ++r)
    for (int n = 0; n < N; ++n) {
      int want = 0;
      for (int k = 0; k < K; ++k) want += a_val(r, k) * b_val(n, k);
      bad_ts += ts[r * N + n] != (float)want;
      bad_ss += ss[r * N + n] != (float)want;
      if ((r == 0 || r == 129) && n < 6) printf("  D[%d][%d]: ts %g ss %g want %d\n", r, n, ts[r * N + n], ss[r * N + n], want);
    }
  printf("mismatches: TS %d, SS %d of %d\n", bad_ts, bad_ss, 256 * N);
  const char* names[4] = {"TS N=128", "SS N=128", "TS N=256", "SS N=256"};
  for (int i = 0; i < 4; ++i) printf("%s: issue %.1f clk/MMA, complete %.1f clk/MMA\n", names[i], c[2 * i] / 64.0, c[2 * i + 1] / 64.0);
  return (bad_ts || bad_ss) ? 1 : 0;
}
