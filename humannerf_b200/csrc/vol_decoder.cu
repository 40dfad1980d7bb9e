// a7: the motion-weight volume decoder, forward and backward, hand-written (fp32 SIMT, deterministic).
//   reference: core/nets/human_nerf/mweight_vol_decoders/deconv_vol_decoder.py:25-33 (softmax(decoder(const_embedding)
//   + log(prior))), core/utils/network_util.py:12-50 ConvDecoder3D (Linear + LeakyReLU(0.2), then ConvTranspose3d(k4, s2,
//   p1) layers with LeakyReLU(0.2) in between).
//
// A transposed convolution with kernel 4, stride 2, padding 1 maps input voxel i and tap k to output voxel o = 2 i - 1 + k
// (per axis).  Output voxels of one PARITY class p = (o mod 2) are reached by exactly two taps per axis,
//     p = 0 (o = 2m):     k = 1 from i = m,   k = 3 from i = m - 1
//     p = 1 (o = 2m + 1): k = 2 from i = m,   k = 0 from i = m + 1
// so the layer is 8 independent GEMMs  out_p[m][co] = sum_{t in 8 taps} sum_ci in[i(m, t)][ci] W[ci][k(p, t)][co]  with
// K = 8 Cin (forward), its transpose over (tap, co) for the data gradient and one small GEMM per tap for the weight
// gradient.  Activations are channel-last ([z][y][x][C]) so that a GEMM row is contiguous; the weights are read and
// their gradients written in the parameters' own NDHWC strides ([Cin][kz][ky][kx][Cout], network.py keeps them in
// torch.channels_last_3d), so nothing is transposed or re-strided per step.
//
// One tiled SIMT kernel (64 x 64 x 16 tiles, 4 x 4 per thread) with three operand loaders (forward / data gradient /
// weight gradient); layers whose GEMM has few rows (the 1^3, 2^3, 4^3 volumes: weight-bandwidth bound) are split over K
// into a partial buffer that a finishing kernel reduces in a fixed order (bias / LeakyReLU / LeakyReLU' fused) - no
// atomics anywhere, results are bit-reproducible.  The last layer's finishing step is softmax(logits + log prior)
// over the 25 channels, written channel-first as the motion field reads it.
#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace hn {

constexpr int kDecMaxConv = 8;
constexpr float kLeaky = 0.2f;
constexpr int kTile = 64, kTileK = 16, kTileStride = kTile + 4;

struct DeconvArgs {
  int din;                 // input volume side (output side = 2 din)
  int cin, cout;           // channels
  int ld_out;              // row stride (floats) of the output-side activation / gradient buffer (>= cout)
  const float* in;         // input-side activation [din^3][cin]            (forward A, wgrad A)
  const float* outside;    // output-side tensor [(2 din)^3][ld_out]         (dgrad A, wgrad B: the gradient of the pre-activation)
  const float* w;          // [cin][64][cout]
  float* dst;              // forward: output activation or partial; dgrad: input gradient or partial; wgrad: dW
  const float* bias;       // forward, direct epilogue
  const float* act_in;     // dgrad, direct epilogue: input-side post-activation (LeakyReLU' mask); null: no mask
  int k_splits;            // > 1: dst = partial[split][row][col]
  int k_per_split;         // multiple of kTileK
  int leaky;               // forward, direct epilogue: apply LeakyReLU
};

__device__ __forceinline__ int tap_k(int p, int t) { return p == 0 ? (t == 0 ? 1 : 3) : (t == 0 ? 2 : 0); }
__device__ __forceinline__ int tap_off(int p, int t) { return t == 0 ? 0 : (p == 0 ? -1 : 1); }

// acc[4][4] += As[k][ty*4 + i] * Bs[k][tx*4 + j]
__device__ __forceinline__ void tile_fma(const float* __restrict__ As, const float* __restrict__ Bs, int ty, int tx, float (&acc)[4][4]) {
#pragma unroll
  for (int kk = 0; kk < kTileK; ++kk) {
    const float4 a = *reinterpret_cast<const float4*>(As + kk * kTileStride + ty * 4);
    const float4 b = *reinterpret_cast<const float4*>(Bs + kk * kTileStride + tx * 4);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

// MODE 0: forward.  grid = (m_tiles * n_tiles, 8 parity classes, k_splits); GEMM rows m = input-grid voxel, cols = cout.
// MODE 1: data gradient.  grid = (m_tiles * n_tiles, 1, k_splits); rows = input voxel i, cols = cin, K = (tap, cout).
// MODE 2: weight gradient.  grid = (m_tiles * n_tiles, 64 taps, 1); rows = cin, cols = cout, K = input voxel.
template <int MODE>
__global__ void __launch_bounds__(256)
deconv_gemm_kernel(const DeconvArgs g) {
  __shared__ __align__(16) float As[kTileK * kTileStride];
  __shared__ __align__(16) float Bs[kTileK * kTileStride];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int din = g.din, din3 = din * din * din, dout = 2 * din;
  const int M = MODE == 2 ? g.cin : din3;
  const int N = MODE == 0 ? g.cout : (MODE == 1 ? g.cin : g.cout);
  const int n_tiles = (N + kTile - 1) / kTile;
  const int m0 = (int)(blockIdx.x / n_tiles) * kTile, n0 = (int)(blockIdx.x % n_tiles) * kTile;
  // K geometry
  const int cout_pad = (g.cout + kTileK - 1) / kTileK * kTileK;          // dgrad: K per tap, zero padded
  const int n_taps = MODE == 0 ? (din == 1 ? 1 : 8) : (din == 1 ? 8 : 64);
  const int k_total = MODE == 0 ? n_taps * g.cin : (MODE == 1 ? n_taps * cout_pad : din3);
  int k_begin = 0, k_end = k_total;
  if (MODE != 2 && g.k_splits > 1) {
    k_begin = (int)blockIdx.z * g.k_per_split;
    k_end = min(k_total, k_begin + g.k_per_split);
  }
  const int pz = (blockIdx.y >> 2) & 1, py = (blockIdx.y >> 1) & 1, px = blockIdx.y & 1;      // MODE 0: parity class
  const int wz = blockIdx.y >> 4, wy = (blockIdx.y >> 2) & 3, wx = blockIdx.y & 3;           // MODE 2: tap

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader roles: A tile = 64 rows x 16 k (thread: row tid >> 2, k quad (tid & 3) * 4) for row-contiguous-in-k operands,
  // or 16 k x 64 rows (thread: k tid >> 4, row quad (tid & 15) * 4) for k-major operands; B tile = 16 k x 64 cols
  for (int k0 = k_begin; k0 < k_end; k0 += kTileK) {
    if (MODE == 0) {
      // A(m, kk) = in[i(m, tap)][ci .. ci + 3]: contiguous in k
      const int r = tid >> 2, kq = (tid & 3) * 4;
      const int m = m0 + r, kk = k0 + kq;
      const int tap = kk / g.cin, ci = kk - tap * g.cin;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < din3 && kk < k_end) {
        const int mz = m / (din * din), my = (m / din) % din, mx = m % din;
        const int iz = mz + tap_off(pz, (tap >> 2) & 1), iy = my + tap_off(py, (tap >> 1) & 1), ix = mx + tap_off(px, tap & 1);
        if (iz >= 0 && iz < din && iy >= 0 && iy < din && ix >= 0 && ix < din)
          v = *reinterpret_cast<const float4*>(g.in + ((size_t)(iz * din + iy) * din + ix) * g.cin + ci);
      }
      As[(kq + 0) * kTileStride + r] = v.x; As[(kq + 1) * kTileStride + r] = v.y;
      As[(kq + 2) * kTileStride + r] = v.z; As[(kq + 3) * kTileStride + r] = v.w;
      // B(kk, n) = w[ci][k(p, tap)][n]: contiguous in n
      const int kr = tid >> 4, nq = (tid & 15) * 4;
      const int kb = k0 + kr;
      const int tapb = kb / g.cin, cib = kb - tapb * g.cin;
      const int kidx = (tap_k(pz, (tapb >> 2) & 1) * 4 + tap_k(py, (tapb >> 1) & 1)) * 4 + tap_k(px, tapb & 1);
      const float* src = g.w + ((size_t)cib * 64 + kidx) * g.cout + n0 + nq;
      float b[4] = {0.f, 0.f, 0.f, 0.f};
      if (kb < k_end) {
        if ((g.cout & 3) == 0 && n0 + nq + 4 <= N) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(src));
          b[0] = q.x; b[1] = q.y; b[2] = q.z; b[3] = q.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (n0 + nq + j < N) b[j] = __ldg(src + j);
        }
      }
      *reinterpret_cast<float4*>(Bs + kr * kTileStride + nq) = make_float4(b[0], b[1], b[2], b[3]);
    } else if (MODE == 1) {
      // A(i, kk) = outside[o(i, tap)][co .. co + 3], kk = tap * cout_pad + co: contiguous in k
      const int r = tid >> 2, kq = (tid & 3) * 4;
      const int i = m0 + r, kk = k0 + kq;
      const int tl = kk / cout_pad, co = kk - tl * cout_pad;
      // tap list: all 64 taps, or (din == 1) the 8 taps with k in {1, 2} per axis - the only ones that reach the 2^3 output
      const int kz = din == 1 ? 1 + ((tl >> 2) & 1) : tl >> 4, ky = din == 1 ? 1 + ((tl >> 1) & 1) : (tl >> 2) & 3,
                kx = din == 1 ? 1 + (tl & 1) : tl & 3;
      float a[4] = {0.f, 0.f, 0.f, 0.f};
      if (i < din3 && kk < k_end) {
        const int iz = i / (din * din), iy = (i / din) % din, ix = i % din;
        const int oz = 2 * iz - 1 + kz, oy = 2 * iy - 1 + ky, ox = 2 * ix - 1 + kx;
        if (oz >= 0 && oz < dout && oy >= 0 && oy < dout && ox >= 0 && ox < dout) {
          const float* src = g.outside + ((size_t)(oz * dout + oy) * dout + ox) * g.ld_out + co;
          if ((g.ld_out & 3) == 0 && co + 4 <= g.cout) {
            const float4 q = *reinterpret_cast<const float4*>(src);
            a[0] = q.x; a[1] = q.y; a[2] = q.z; a[3] = q.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (co + j < g.cout) a[j] = src[j];
          }
        }
      }
      As[(kq + 0) * kTileStride + r] = a[0]; As[(kq + 1) * kTileStride + r] = a[1];
      As[(kq + 2) * kTileStride + r] = a[2]; As[(kq + 3) * kTileStride + r] = a[3];
      // B(kk, n = ci) = w[ci][tap][co]: contiguous in k for a fixed ci -> thread: column ci = tid >> 2, k quad
      const int c = tid >> 2, ci = n0 + c;
      const int kidx = (kz * 4 + ky) * 4 + kx;
      float b[4] = {0.f, 0.f, 0.f, 0.f};
      if (ci < N && kk < k_end) {
        const float* src = g.w + ((size_t)ci * 64 + kidx) * g.cout + co;
        if ((g.cout & 3) == 0 && co + 4 <= g.cout) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(src));
          b[0] = q.x; b[1] = q.y; b[2] = q.z; b[3] = q.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (co + j < g.cout) b[j] = __ldg(src + j);
        }
      }
      Bs[(kq + 0) * kTileStride + c] = b[0]; Bs[(kq + 1) * kTileStride + c] = b[1];
      Bs[(kq + 2) * kTileStride + c] = b[2]; Bs[(kq + 3) * kTileStride + c] = b[3];
    } else {
      // A(ci, i) = in[i][ci]: contiguous in the row index -> thread: k = tid >> 4, row quad
      const int kr = tid >> 4, rq = (tid & 15) * 4;
      const int i = k0 + kr;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < k_end && m0 + rq + 4 <= M) v = *reinterpret_cast<const float4*>(g.in + (size_t)i * g.cin + m0 + rq);
      *reinterpret_cast<float4*>(As + kr * kTileStride + rq) = v;
      // B(i, co) = outside[o(i, tap)][co]
      float b[4] = {0.f, 0.f, 0.f, 0.f};
      if (i < k_end) {
        const int iz = i / (din * din), iy = (i / din) % din, ix = i % din;
        const int oz = 2 * iz - 1 + wz, oy = 2 * iy - 1 + wy, ox = 2 * ix - 1 + wx;
        if (oz >= 0 && oz < dout && oy >= 0 && oy < dout && ox >= 0 && ox < dout) {
          const float* src = g.outside + ((size_t)(oz * dout + oy) * dout + ox) * g.ld_out + n0 + rq;
          if ((g.ld_out & 3) == 0 && n0 + rq + 4 <= N) {
            const float4 q = *reinterpret_cast<const float4*>(src);
            b[0] = q.x; b[1] = q.y; b[2] = q.z; b[3] = q.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (n0 + rq + j < N) b[j] = src[j];
          }
        }
      }
      *reinterpret_cast<float4*>(Bs + kr * kTileStride + rq) = make_float4(b[0], b[1], b[2], b[3]);
    }
    __syncthreads();
    tile_fma(As, Bs, ty, tx, acc);
    __syncthreads();
  }

  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (MODE == 0) {
        const int mz = m / (din * din), my = (m / din) % din, mx = m % din;
        const size_t o = ((size_t)((2 * mz + pz) * dout + 2 * my + py)) * dout + 2 * mx + px;
        if (g.k_splits > 1) {
          g.dst[((size_t)blockIdx.z * dout * dout * dout + o) * g.cout + n] = v;
        } else {
          v += __ldg(g.bias + n);
          if (g.leaky) v = v > 0.f ? v : kLeaky * v;
          g.dst[o * g.ld_out + n] = v;
        }
      } else if (MODE == 1) {
        if (g.k_splits > 1) {
          g.dst[((size_t)blockIdx.z * din3 + m) * g.cin + n] = v;
        } else {
          if (g.act_in) v *= g.act_in[(size_t)m * g.cin + n] > 0.f ? 1.0f : kLeaky;
          g.dst[(size_t)m * g.cin + n] = v;
        }
      } else {
        g.dst[((size_t)m * 64 + blockIdx.y) * g.cout + n] = v;
      }
    }
  }
}

// out[row][col] = f(sum_s partial[s][row][col]): forward (bias + optional LeakyReLU) or data gradient (LeakyReLU' mask)
__global__ void deconv_finish_kernel(const float* __restrict__ partial, int splits, int64_t rows, int cols, int ld_out,
                                     const float* __restrict__ bias, int leaky, const float* __restrict__ act_in,
                                     float* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const int64_t r = idx / cols;
  const int c = (int)(idx - r * cols);
  float v = 0.f;
  for (int s = 0; s < splits; ++s) v += partial[((size_t)s * rows + r) * cols + c];
  if (bias) v += __ldg(bias + c);
  if (leaky) v = v > 0.f ? v : kLeaky * v;
  if (act_in) v *= act_in[idx] > 0.f ? 1.0f : kLeaky;
  out[(size_t)r * ld_out + c] = v;
}

// h[j] = LeakyReLU(b[j] + sum_i W[j][i] e[i]): one warp per output
__global__ void linear_lrelu_fwd_kernel(int n_out, int n_in, const float* __restrict__ w, const float* __restrict__ b,
                                        const float* __restrict__ e, float* __restrict__ h) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (j >= n_out) return;
  float s = 0.f;
  for (int i = lane; i < n_in; i += 32) s = fmaf(__ldg(w + (size_t)j * n_in + i), __ldg(e + i), s);
  s = warp_sum(s);
  if (lane == 0) {
    s += __ldg(b + j);
    h[j] = s > 0.f ? s : kLeaky * s;
  }
}
// d_w[j][i] = dpre[j] e[i]; d_b[j] = dpre[j]   (dpre already carries LeakyReLU')
__global__ void linear_wgrad_kernel(int n_out, int n_in, const float* __restrict__ dpre, const float* __restrict__ e,
                                    float* __restrict__ d_w, float* __restrict__ d_b) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_out * n_in) return;
  const int j = idx / n_in, i = idx - j * n_in;
  d_w[idx] = dpre[j] * __ldg(e + i);
  if (i == 0) d_b[j] = dpre[j];
}
// d_e[i] = sum_j W[j][i] dpre[j]: one thread per input, fixed order
__global__ void linear_dgrad_kernel(int n_out, int n_in, const float* __restrict__ w, const float* __restrict__ dpre,
                                    float* __restrict__ d_e) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_in) return;
  float s = 0.f;
  for (int j = 0; j < n_out; ++j) s = fmaf(__ldg(w + (size_t)j * n_in + i), dpre[j], s);
  d_e[i] = s;
}
// d_b[c] = sum_rows g[row][c]: one block of 256 threads per column, fixed-order tree
__global__ void __launch_bounds__(256) column_sum_kernel(int64_t rows, int ld, const float* __restrict__ g, float* __restrict__ out) {
  __shared__ float red[256];
  const int c = blockIdx.x;
  float s = 0.f;
  for (int64_t r = threadIdx.x; r < rows; r += 256) s += g[r * ld + c];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[c] = red[0];
}

// vol[c][voxel] = softmax_c(logits[voxel][c] + log prior[c][voxel]); one thread per voxel
__global__ void softmax_prior_fwd_kernel(int n_vox, int C, int ld, const float* __restrict__ logits, const float* __restrict__ prior,
                                         float* __restrict__ vol) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vox) return;
  float x[32];
  float mx = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < C; ++c) {
    x[c] = logits[(size_t)v * ld + c] + logf(prior[(size_t)c * n_vox + v]);
    mx = fmaxf(mx, x[c]);
  }
  float s = 0.f;
#pragma unroll 1
  for (int c = 0; c < C; ++c) { x[c] = expf(x[c] - mx); s += x[c]; }
  const float inv = 1.0f / s;
#pragma unroll 1
  for (int c = 0; c < C; ++c) vol[(size_t)c * n_vox + v] = x[c] * inv;
}
// d_logits[voxel][c] = y_c (g_c - sum_k g_k y_k)
__global__ void softmax_prior_bwd_kernel(int n_vox, int C, int ld, const float* __restrict__ vol, const float* __restrict__ d_vol,
                                         float* __restrict__ d_logits) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vox) return;
  float dot = 0.f;
#pragma unroll 1
  for (int c = 0; c < C; ++c) dot = fmaf(d_vol[(size_t)c * n_vox + v], vol[(size_t)c * n_vox + v], dot);
#pragma unroll 1
  for (int c = 0; c < C; ++c) {
    const float y = vol[(size_t)c * n_vox + v];
    d_logits[(size_t)v * ld + c] = y * (d_vol[(size_t)c * n_vox + v] - dot);
  }
#pragma unroll 1
  for (int c = C; c < ld; ++c) d_logits[(size_t)v * ld + c] = 0.f;
}

// ------------------------------------------------------------------------------------------------------------------
struct DecPlan {
  int n_conv, embed, hidden;
  int ch[kDecMaxConv + 1];
  int din[kDecMaxConv];
  int ld_logits;
  int fwd_splits[kDecMaxConv], dgrad_splits[kDecMaxConv];                 // K splits of the forward / data-gradient GEMMs
  size_t off_h, off_act[kDecMaxConv + 1], off_logits, ws_floats;          // saved for backward
  size_t off_partial, off_g[2], off_dlogits, scratch_floats;
};

// K-splits for a GEMM with `rows` rows: the low layers are weight-bandwidth bound and need many CTAs
static int pick_splits(int64_t tiles, int k_total) {
  int splits = 1;
  while (tiles * splits < 296 && splits < 16 && k_total / (splits * 2) >= 64) splits *= 2;
  return splits;
}

static int make_dec_plan(int embed, int hidden, int n_conv, const int* channels, DecPlan& p) {
  if (embed <= 0 || hidden <= 0 || n_conv < 1 || n_conv > kDecMaxConv || !channels)
    return fail_arg("hn_vol_decoder: bad geometry");
  if (channels[0] != hidden) return fail_arg("hn_vol_decoder: the first convolution must take the linear layer's %d features", hidden);
  p.n_conv = n_conv; p.embed = embed; p.hidden = hidden;
  for (int l = 0; l <= n_conv; ++l) {
    p.ch[l] = channels[l];
    if (channels[l] <= 0 || (l < n_conv && (channels[l] % 16))) return fail_arg("hn_vol_decoder: channel counts must be multiples of 16 (except the output)");
  }
  if (p.ch[n_conv] > 32) return fail_arg("hn_vol_decoder: at most 32 output channels");
  for (int l = 0; l < n_conv; ++l) p.din[l] = 1 << l;
  p.ld_logits = 32;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += (n + 63) / 64 * 64; return o; };
  p.off_h = 0;
  // act[l] = input of convolution l (act[0] = h), channel-last
  for (int l = 0; l < n_conv; ++l) p.off_act[l] = take((size_t)p.din[l] * p.din[l] * p.din[l] * p.ch[l]);
  const size_t n_vox = (size_t)8 * p.din[n_conv - 1] * p.din[n_conv - 1] * p.din[n_conv - 1];
  p.off_logits = take(n_vox * p.ld_logits);
  p.ws_floats = off;
  // scratch: split-K partials (forward and data gradient), two gradient buffers, d_logits
  size_t part = 0, gmax = 0;
  for (int l = 0; l < n_conv; ++l) {
    const size_t d3 = (size_t)p.din[l] * p.din[l] * p.din[l];
    const int64_t ft = (int64_t)ceil_div<int64_t>((int64_t)d3, kTile) * ceil_div(p.ch[l + 1], kTile) * 8;
    p.fwd_splits[l] = pick_splits(ft, (p.din[l] == 1 ? 1 : 8) * p.ch[l]);
    const int64_t dt = (int64_t)ceil_div<int64_t>((int64_t)d3, kTile) * ceil_div(p.ch[l], kTile);
    p.dgrad_splits[l] = pick_splits(dt, (p.din[l] == 1 ? 8 : 64) * (ceil_div(p.ch[l + 1], kTileK) * kTileK));
    if (p.fwd_splits[l] > 1) part = std::max(part, (size_t)p.fwd_splits[l] * 8 * d3 * p.ch[l + 1]);
    if (p.dgrad_splits[l] > 1) part = std::max(part, (size_t)p.dgrad_splits[l] * d3 * p.ch[l]);
    gmax = std::max(gmax, std::max(d3 * p.ch[l], 8 * d3 * (size_t)std::max(p.ch[l + 1], l == n_conv - 1 ? p.ld_logits : 0)));
  }
  off = 0;
  p.off_partial = take(std::max<size_t>(part, 64));
  p.off_g[0] = take(gmax);
  p.off_g[1] = take(gmax);
  p.off_dlogits = take(n_vox * p.ld_logits);
  p.scratch_floats = off;
  return 0;
}

static int run_forward_layer(const DecPlan& p, int l, const float* in, const float* w, const float* b, float* out, int ld_out,
                             bool leaky, float* partial, cudaStream_t st) {
  DeconvArgs g;
  memset(&g, 0, sizeof(g));
  const int din = p.din[l], d3 = din * din * din;
  g.din = din; g.cin = p.ch[l]; g.cout = p.ch[l + 1]; g.ld_out = ld_out; g.in = in; g.w = w; g.bias = b; g.leaky = leaky ? 1 : 0;
  const int k_total = (din == 1 ? 1 : 8) * g.cin;
  const int64_t tiles = (int64_t)ceil_div(d3, kTile) * ceil_div(g.cout, kTile);
  const int splits = p.fwd_splits[l];
  g.k_splits = splits;
  g.k_per_split = ceil_div(ceil_div(k_total, splits), kTileK) * kTileK;
  g.dst = splits > 1 ? partial : out;
  dim3 grid((unsigned)tiles, 8, (unsigned)splits);
  deconv_gemm_kernel<0><<<grid, 256, 0, st>>>(g);
  HN_TRY(launched("hn_vol_decoder_fwd/deconv"));
  if (splits > 1) {
    const int64_t rows = (int64_t)8 * d3;
    deconv_finish_kernel<<<(unsigned)ceil_div<int64_t>(rows * g.cout, 256), 256, 0, st>>>(partial, splits, rows, g.cout, ld_out, b,
                                                                                       leaky ? 1 : 0, nullptr, out);
    HN_TRY(launched("hn_vol_decoder_fwd/finish"));
  }
  return 0;
}

}  // namespace hn

using namespace hn;

extern "C" {

int64_t hn_vol_decoder_workspace_bytes(int embed, int hidden, int n_conv, const int* channels) {
  DecPlan p;
  if (make_dec_plan(embed, hidden, n_conv, channels, p) != 0) return -1;
  return (int64_t)p.ws_floats * 4 + 256;
}
int64_t hn_vol_decoder_scratch_bytes(int embed, int hidden, int n_conv, const int* channels) {
  DecPlan p;
  if (make_dec_plan(embed, hidden, n_conv, channels, p) != 0) return -1;
  return (int64_t)p.scratch_floats * 4 + 256;
}

int hn_vol_decoder_fwd(int embed, int hidden, int n_conv, const int* channels, const float* embedding, const float* lin_w,
                       const float* lin_b, const float* const* conv_w, const float* const* conv_b, const float* prior,
                       float* vol, void* workspace, int64_t workspace_bytes, void* scratch, int64_t scratch_bytes, void* stream) {
  DecPlan p;
  HN_TRY(make_dec_plan(embed, hidden, n_conv, channels, p));
  if (!embedding || !lin_w || !lin_b || !conv_w || !conv_b || !prior || !vol || !workspace || !scratch)
    return fail_arg("hn_vol_decoder_fwd: null pointer");
  for (int l = 0; l < n_conv; ++l)
    if (!conv_w[l] || !conv_b[l]) return fail_arg("hn_vol_decoder_fwd: null layer %d", l);
  if (workspace_bytes < (int64_t)p.ws_floats * 4 || scratch_bytes < (int64_t)p.scratch_floats * 4)
    return fail_arg("hn_vol_decoder_fwd: workspace / scratch too small");
  cudaStream_t st = as_stream(stream);
  float* ws = static_cast<float*>(workspace);
  float* sc = static_cast<float*>(scratch);
  linear_lrelu_fwd_kernel<<<ceil_div(hidden * 32, 256), 256, 0, st>>>(hidden, embed, lin_w, lin_b, embedding, ws + p.off_act[0]);
  HN_TRY(launched("hn_vol_decoder_fwd/linear"));
  for (int l = 0; l < n_conv; ++l) {
    const bool last = l == n_conv - 1;
    float* out = last ? ws + p.off_logits : ws + p.off_act[l + 1];
    HN_TRY(run_forward_layer(p, l, ws + p.off_act[l], conv_w[l], conv_b[l], out, last ? p.ld_logits : p.ch[l + 1], !last,
                             sc + p.off_partial, st));
  }
  const int side = 2 * p.din[n_conv - 1], n_vox = side * side * side;
  softmax_prior_fwd_kernel<<<ceil_div(n_vox, 128), 128, 0, st>>>(n_vox, p.ch[n_conv], p.ld_logits, ws + p.off_logits, prior, vol);
  return launched("hn_vol_decoder_fwd/softmax");
}

int hn_vol_decoder_bwd(int embed, int hidden, int n_conv, const int* channels, const float* embedding, const float* lin_w,
                       const float* const* conv_w, const float* vol, const float* d_vol, const void* workspace,
                       int64_t workspace_bytes, void* scratch, int64_t scratch_bytes, float* d_embedding, float* d_lin_w,
                       float* d_lin_b, float* const* d_conv_w, float* const* d_conv_b, void* stream) {
  DecPlan p;
  HN_TRY(make_dec_plan(embed, hidden, n_conv, channels, p));
  if (!embedding || !lin_w || !conv_w || !vol || !d_vol || !workspace || !scratch || !d_embedding || !d_lin_w || !d_lin_b ||
      !d_conv_w || !d_conv_b)
    return fail_arg("hn_vol_decoder_bwd: null pointer");
  for (int l = 0; l < n_conv; ++l)
    if (!conv_w[l] || !d_conv_w[l] || !d_conv_b[l]) return fail_arg("hn_vol_decoder_bwd: null layer %d", l);
  if (workspace_bytes < (int64_t)p.ws_floats * 4 || scratch_bytes < (int64_t)p.scratch_floats * 4)
    return fail_arg("hn_vol_decoder_bwd: workspace / scratch too small");
  cudaStream_t st = as_stream(stream);
  const float* ws = static_cast<const float*>(workspace);
  float* sc = static_cast<float*>(scratch);
  const int side = 2 * p.din[n_conv - 1], n_vox = side * side * side;
  float* g_out = sc + p.off_dlogits;                       // gradient of the pre-activation on the output side of layer l
  int ld_g = p.ld_logits;
  softmax_prior_bwd_kernel<<<ceil_div(n_vox, 128), 128, 0, st>>>(n_vox, p.ch[n_conv], p.ld_logits, vol, d_vol, g_out);
  HN_TRY(launched("hn_vol_decoder_bwd/softmax"));
  for (int l = n_conv - 1; l >= 0; --l) {
    const int din = p.din[l], d3 = din * din * din, cin = p.ch[l], cout = p.ch[l + 1];
    const int64_t rows_out = (int64_t)8 * d3;
    // bias gradient
    column_sum_kernel<<<cout, 256, 0, st>>>(rows_out, ld_g, g_out, d_conv_b[l]);
    HN_TRY(launched("hn_vol_decoder_bwd/bias"));
    // weight gradient: one small GEMM per tap, written in the parameter's [cin][64][cout] layout
    DeconvArgs g;
    memset(&g, 0, sizeof(g));
    g.din = din; g.cin = cin; g.cout = cout; g.ld_out = ld_g; g.in = ws + p.off_act[l]; g.outside = g_out; g.dst = d_conv_w[l];
    g.k_splits = 1;
    dim3 wgrid((unsigned)(ceil_div(cin, kTile) * ceil_div(cout, kTile)), 64, 1);
    deconv_gemm_kernel<2><<<wgrid, 256, 0, st>>>(g);
    HN_TRY(launched("hn_vol_decoder_bwd/wgrad"));
    // data gradient -> gradient of the pre-activation of the previous layer (LeakyReLU' from its stored output)
    float* g_in = sc + p.off_g[l & 1];
    g.w = conv_w[l]; g.act_in = ws + p.off_act[l];
    const int k_total = (din == 1 ? 8 : 64) * ceil_div(cout, kTileK) * kTileK;
    const int64_t tiles = (int64_t)ceil_div(d3, kTile) * ceil_div(cin, kTile);
    const int splits = p.dgrad_splits[l];
    g.k_splits = splits;
    g.k_per_split = ceil_div(ceil_div(k_total, splits), kTileK) * kTileK;
    g.dst = splits > 1 ? sc + p.off_partial : g_in;
    dim3 dgrid((unsigned)tiles, 1, (unsigned)splits);
    deconv_gemm_kernel<1><<<dgrid, 256, 0, st>>>(g);
    HN_TRY(launched("hn_vol_decoder_bwd/dgrad"));
    if (splits > 1) {
      deconv_finish_kernel<<<(unsigned)ceil_div<int64_t>((int64_t)d3 * cin, 256), 256, 0, st>>>(sc + p.off_partial, splits, d3, cin, cin, nullptr, 0,
                                                                                          ws + p.off_act[l], g_in);
      HN_TRY(launched("hn_vol_decoder_bwd/finish"));
    }
    g_out = g_in;
    ld_g = cin;
  }
  // linear layer: g_out = d(pre-activation of h) (hidden)
  linear_wgrad_kernel<<<ceil_div(hidden * embed, 256), 256, 0, st>>>(hidden, embed, g_out, embedding, d_lin_w, d_lin_b);
  HN_TRY(launched("hn_vol_decoder_bwd/linear wgrad"));
  linear_dgrad_kernel<<<ceil_div(embed, 128), 128, 0, st>>>(hidden, embed, lin_w, g_out, d_embedding);
  return launched("hn_vol_decoder_bwd/linear dgrad");
}

}  // extern "C"
