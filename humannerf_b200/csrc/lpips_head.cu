// LPIPS "head" of one trunk stage, fused (a16: trainer.py:97-113 get_img_rebuild_loss -> third_parties/lpips/lpips.py
// :78-92: normalize_tensor (third_parties/lpips/__init__.py:40-42 - this fork's variant with eps INSIDE and outside
// the root: f / (sqrt(sum_c f^2 + eps) + eps), eps = 1e-10), squared difference, 1x1 non-negative `lin` layer,
// spatial_average (lpips.py:14-15)):
//   out[i] += (1 / HW) * sum_p sum_c lin[c] * ( fx[i,c,p] / nx - fy[i,c,p] / ny )^2,   nx = sqrt(||fx[i,:,p]||^2 + eps) + eps
// for n image pairs whose features are the two halves of ONE (2n, C, H, W) tensor (prediction rows first), any
// dense strides (cuDNN hands back NHWC).  One warp per (pair, position): lanes stride over the channels, two
// passes over C (norms, then the weighted squared difference), warp-shuffle reductions, one atomicAdd per warp.
// Replaces ~10 elementwise / reduction launches per stage in the forward and ~15 in the backward.
#include "common.cuh"

namespace hn {

constexpr float kLpipsEps = 1e-10f;

__global__ void __launch_bounds__(256)
lpips_head_fwd_kernel(int n, int C, int HW, const float* __restrict__ f, int64_t s_n, int64_t s_c, int64_t s_p,
                      const float* __restrict__ lin, float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n * HW) return;
  const int i = warp / HW, p = warp - i * HW;
  const float* fx = f + (int64_t)i * s_n + (int64_t)p * s_p;
  const float* fy = f + (int64_t)(n + i) * s_n + (int64_t)p * s_p;
  float sx = 0.f, sy = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float a = fx[(int64_t)c * s_c], b = fy[(int64_t)c * s_c];
    sx += a * a; sy += b * b;
  }
  sx = warp_sum(sx); sy = warp_sum(sy);
  const float ix = 1.0f / (sqrtf(sx + kLpipsEps) + kLpipsEps), iy = 1.0f / (sqrtf(sy + kLpipsEps) + kLpipsEps);
  float acc = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = fx[(int64_t)c * s_c] * ix - fy[(int64_t)c * s_c] * iy;
    acc += __ldg(lin + c) * d * d;
  }
  acc = warp_sum(acc);
  if (lane == 0) atomicAdd(out + i, acc / (float)HW);
}

// df = g[i] / HW * d(out[i]) / df for both halves:
//   d_c = fx_c / nx - fy_c / ny,  A = sum_k 2 lin_k d_k fx_k,  B = sum_k 2 lin_k d_k fy_k,  rx = sqrt(||fx||^2 + eps), nx = rx + eps
//   dfx_c = g ( 2 lin_c d_c / nx - fx_c A / (nx^2 rx) ),  dfy_c = g ( -2 lin_c d_c / ny + fy_c B / (ny^2 ry) )
// (rx >= 1e-5: an all-zero feature vector - common after a ReLU - has the finite slope 1 / nx, as in the reference)
__global__ void __launch_bounds__(256)
lpips_head_bwd_kernel(int n, int C, int HW, const float* __restrict__ f, int64_t s_n, int64_t s_c, int64_t s_p,
                      const float* __restrict__ lin, const float* __restrict__ g_out, float* __restrict__ df) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n * HW) return;
  const int i = warp / HW, p = warp - i * HW;
  const int64_t ox = (int64_t)i * s_n + (int64_t)p * s_p, oy = (int64_t)(n + i) * s_n + (int64_t)p * s_p;
  float sx = 0.f, sy = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float a = f[ox + (int64_t)c * s_c], b = f[oy + (int64_t)c * s_c];
    sx += a * a; sy += b * b;
  }
  sx = warp_sum(sx); sy = warp_sum(sy);
  const float rx = sqrtf(sx + kLpipsEps), ry = sqrtf(sy + kLpipsEps);
  const float ix = 1.0f / (rx + kLpipsEps), iy = 1.0f / (ry + kLpipsEps);
  float A = 0.f, B = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float a = f[ox + (int64_t)c * s_c], b = f[oy + (int64_t)c * s_c];
    const float t = 2.0f * __ldg(lin + c) * (a * ix - b * iy);
    A += t * a; B += t * b;
  }
  A = warp_sum(A); B = warp_sum(B);
  const float g = __ldg(g_out + i) / (float)HW;
  const float kx = A * ix * ix / rx, ky = B * iy * iy / ry;
  for (int c = lane; c < C; c += 32) {
    const float a = f[ox + (int64_t)c * s_c], b = f[oy + (int64_t)c * s_c];
    const float t = 2.0f * __ldg(lin + c) * (a * ix - b * iy);
    df[ox + (int64_t)c * s_c] = g * (t * ix - a * kx);
    df[oy + (int64_t)c * s_c] = g * (-t * iy + b * ky);
  }
}

}  // namespace hn

using namespace hn;

extern "C" {

int hn_lpips_head_fwd(int n_pairs, int channels, int positions, const float* feat, int64_t stride_n, int64_t stride_c,
                      int64_t stride_p, const float* lin, float* out, void* stream) {
  if (n_pairs == 0 || positions == 0) return 0;
  if (n_pairs < 0 || channels <= 0 || positions < 0 || !feat || !lin || !out) return fail_arg("hn_lpips_head_fwd: bad argument");
  const int64_t warps = (int64_t)n_pairs * positions;
  lpips_head_fwd_kernel<<<(unsigned)ceil_div<int64_t>(warps * 32, 256), 256, 0, as_stream(stream)>>>(
      n_pairs, channels, positions, feat, stride_n, stride_c, stride_p, lin, out);
  return launched("hn_lpips_head_fwd");
}

int hn_lpips_head_bwd(int n_pairs, int channels, int positions, const float* feat, int64_t stride_n, int64_t stride_c,
                      int64_t stride_p, const float* lin, const float* d_out, float* d_feat, void* stream) {
  if (n_pairs == 0 || positions == 0) return 0;
  if (n_pairs < 0 || channels <= 0 || positions < 0 || !feat || !lin || !d_out || !d_feat)
    return fail_arg("hn_lpips_head_bwd: bad argument");
  const int64_t warps = (int64_t)n_pairs * positions;
  lpips_head_bwd_kernel<<<(unsigned)ceil_div<int64_t>(warps * 32, 256), 256, 0, as_stream(stream)>>>(
      n_pairs, channels, positions, feat, stride_n, stride_c, stride_p, lin, d_out, d_feat);
  return launched("hn_lpips_head_bwd");
}

}  // extern "C"
