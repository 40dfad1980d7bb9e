// fp32 SIMT building blocks of the canonical-field MLPs (HN_PRECISION_FP32 path): a tiled
// SGEMM with fused bias / ReLU / ReLU-mask / split-K epilogues, a column sum (bias gradients)
// and the positional-encoding kernels.  This is the 1e-4 parity path; the throughput path is
// the tcgen05 kernel in mlp_tc.cu.
#pragma once

#include "common.cuh"

namespace hn {

// C[M,N] (op)= sum_k A(m,k) * B(k,n)
//   A(m,k) = A_KC ? A[m*lda + k] : A[k*lda + m]
//   B(k,n) = B_KC ? B[n*ldb + k] : B[k*ldb + n]
struct GemmArgs {
  int64_t M;
  int N, K;
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  const float* bias;        // nullable, per n
  const float* mask_src;    // nullable: C = (mask_src[m*ld_mask + n] > 0) ? C : 0
  int64_t ld_mask;
  int relu;                 // C = max(C, 0)
  int accumulate;           // C += result (plain read-modify-write)
  int atomic;               // atomicAdd result into C (split-K)
  int64_t k_chunk;          // K range handled per blockIdx.y when atomic (split-K)
};

constexpr int kBM = 128;
constexpr int kBK = 16;

template <int TN, bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256)
sgemm_kernel(GemmArgs g) {
  constexpr int BN = 16 * TN;
  constexpr int ASTR = kBM + 4;
  constexpr int BSTR = BN + 4;
  __shared__ __align__(16) float As[kBK * ASTR];
  __shared__ __align__(16) float Bs[kBK * BSTR];

  const int tid = threadIdx.x;
  const int n_tiles = (g.N + BN - 1) / BN;
  const int64_t tile = blockIdx.x;
  const int64_t m0 = (tile / n_tiles) * kBM;
  const int n0 = (int)(tile % n_tiles) * BN;
  int64_t k_begin = 0, k_end = g.K;
  if (g.atomic) {
    k_begin = (int64_t)blockIdx.y * g.k_chunk;
    k_end = min((int64_t)g.K, k_begin + g.k_chunk);
  }
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  const bool a_vec = ((g.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0);
  const bool b_vec = ((g.ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0);

  for (int64_t k0 = k_begin; k0 < k_end; k0 += kBK) {
    // ---- A tile -> As[k][m]
    if (A_KC) {
      const int m = tid >> 1, kk = (tid & 1) * 8;
      const int64_t gm = m0 + m;
      float v[8];
      const float* src = g.A + gm * g.lda + k0 + kk;
      if (gm < g.M && a_vec && k0 + kk + 8 <= k_end && ((k0 + kk) & 3) == 0) {
        float4 p = *reinterpret_cast<const float4*>(src);
        float4 q = *reinterpret_cast<const float4*>(src + 4);
        v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w; v[4] = q.x; v[5] = q.y; v[6] = q.z; v[7] = q.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (gm < g.M && k0 + kk + i < k_end) ? src[i] : 0.0f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) As[(kk + i) * ASTR + m] = v[i];
    } else {
      const int kk = tid >> 4, m = (tid & 15) * 8;
      const int64_t gk = k0 + kk;
      const float* src = g.A + gk * g.lda + m0 + m;
      float v[8];
      if (gk < k_end && a_vec && m0 + m + 8 <= g.M && ((m0 + m) & 3) == 0) {
        float4 p = *reinterpret_cast<const float4*>(src);
        float4 q = *reinterpret_cast<const float4*>(src + 4);
        v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w; v[4] = q.x; v[5] = q.y; v[6] = q.z; v[7] = q.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (gk < k_end && m0 + m + i < g.M) ? src[i] : 0.0f;
      }
      *reinterpret_cast<float4*>(&As[kk * ASTR + m]) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(&As[kk * ASTR + m + 4]) = make_float4(v[4], v[5], v[6], v[7]);
    }
    // ---- B tile -> Bs[k][n]
    if (B_KC) {
      const int n = tid >> 1, kk = (tid & 1) * 8;
      if (n < BN) {
        const int gn = n0 + n;
        const float* src = g.B + (int64_t)gn * g.ldb + k0 + kk;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          Bs[(kk + i) * BSTR + n] = (gn < g.N && k0 + kk + i < k_end) ? __ldg(src + i) : 0.0f;
      }
    } else {
      if (TN == 8) {
        const int kk = tid >> 4, n = (tid & 15) * 8;
        const int64_t gk = k0 + kk;
        const float* src = g.B + gk * g.ldb + n0 + n;
        float v[8];
        if (gk < k_end && b_vec && n0 + n + 8 <= g.N && ((n0 + n) & 3) == 0) {
          float4 p = *reinterpret_cast<const float4*>(src);
          float4 q = *reinterpret_cast<const float4*>(src + 4);
          v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w; v[4] = q.x; v[5] = q.y; v[6] = q.z; v[7] = q.w;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = (gk < k_end && n0 + n + i < g.N) ? __ldg(src + i) : 0.0f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) Bs[kk * BSTR + n + i] = v[i];
      } else {
        const int kk = tid >> 4, n = tid & 15;
        const int64_t gk = k0 + kk;
        Bs[kk * BSTR + n] = (gk < k_end && n0 + n < g.N) ? __ldg(g.B + gk * g.ldb + n0 + n) : 0.0f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      float a[8], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk * ASTR + ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk * ASTR + 64 + ty * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      if (TN == 8) {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk * BSTR + tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk * BSTR + 64 + tx * 4]);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
        b[4 % TN] = b1.x; b[5 % TN] = b1.y; b[6 % TN] = b1.z; b[7 % TN] = b1.w;
      } else {
        b[0] = Bs[kk * BSTR + tx];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue: rows {ty*4+i, 64+ty*4+i}, cols {tx*4+j, 64+tx*4+j} (TN=8) or {tx} (TN=1)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + (TN == 8 ? (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4)) : tx);
      if (gn >= g.N) continue;
      float v = acc[i][j];
      float* dst = g.C + gm * g.ldc + gn;
      if (g.atomic) {
        atomicAdd(dst, v);
        continue;
      }
      if (g.bias) v += __ldg(g.bias + gn);
      if (g.relu) v = fmaxf(v, 0.0f);
      if (g.mask_src && !(g.mask_src[gm * g.ld_mask + gn] > 0.0f)) v = 0.0f;
      if (g.accumulate) v += *dst;
      *dst = v;
    }
  }
}

template <bool A_KC, bool B_KC>
inline int launch_sgemm(const GemmArgs& g, cudaStream_t st, const char* what) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return 0;
  const bool small_n = g.N <= 16;
  const int BN = small_n ? 16 : 128;
  const int64_t tiles = ceil_div<int64_t>(g.M, kBM) * ceil_div<int>(g.N, BN);
  unsigned splits = 1;
  if (g.atomic) splits = (unsigned)ceil_div<int64_t>(g.K, g.k_chunk);
  dim3 grid((unsigned)tiles, splits, 1);
  if (small_n)
    sgemm_kernel<1, A_KC, B_KC><<<grid, 256, 0, st>>>(g);
  else
    sgemm_kernel<8, A_KC, B_KC><<<grid, 256, 0, st>>>(g);
  return launched(what);
}

// Y[M,N] = act(X[M,K] W[N,K]^T + b)
inline int linear_fwd(int64_t M, int N, int K, const float* X, int64_t ldx, const float* W, int64_t ldw,
                      const float* bias, bool relu, float* Y, int64_t ldy, cudaStream_t st) {
  GemmArgs g{};
  g.M = M; g.N = N; g.K = K; g.A = X; g.lda = ldx; g.B = W; g.ldb = ldw; g.C = Y; g.ldc = ldy;
  g.bias = bias; g.relu = relu ? 1 : 0;
  return launch_sgemm<true, true>(g, st, "linear_fwd");
}

// dX[M,K] (+)= dY[M,N] W[N,K]  (W may be a column window: pass W + col0, ldw unchanged);
// optional ReLU mask from the saved post-activation input.
inline int linear_bwd_data(int64_t M, int N, int K, const float* dY, int64_t lddy, const float* W, int64_t ldw,
                           float* dX, int64_t lddx, const float* mask_src, int64_t ld_mask, bool accumulate,
                           cudaStream_t st) {
  GemmArgs g{};
  g.M = M; g.N = K; g.K = N; g.A = dY; g.lda = lddy; g.B = W; g.ldb = ldw; g.C = dX; g.ldc = lddx;
  g.mask_src = mask_src; g.ld_mask = ld_mask; g.accumulate = accumulate ? 1 : 0;
  return launch_sgemm<true, false>(g, st, "linear_bwd_data");
}

// dW[N,K] += dY[M,N]^T X[M,K]   (split over M, atomics)
inline int linear_bwd_weight(int64_t M, int N, int K, const float* dY, int64_t lddy, const float* X,
                             int64_t ldx, float* dW, int64_t ldw, cudaStream_t st) {
  GemmArgs g{};
  g.M = N; g.N = K; g.K = (int)M; g.A = dY; g.lda = lddy; g.B = X; g.ldb = ldx; g.C = dW; g.ldc = ldw;
  g.atomic = 1;
  // aim for ~600 CTAs; chunks are multiples of the K tile
  int64_t tiles = ceil_div<int64_t>(N, kBM) * ceil_div<int>(K, K <= 16 ? 16 : 128);
  int64_t want = ceil_div<int64_t>(592, tiles);
  int64_t chunk = ceil_div<int64_t>(ceil_div<int64_t>(M, want), kBK) * kBK;
  if (chunk < 256) chunk = 256;
  g.k_chunk = chunk;
  if (M > 0x7fffffffLL) return fail_arg("linear_bwd_weight: too many samples");
  return launch_sgemm<false, false>(g, st, "linear_bwd_weight");
}

}  // namespace hn
