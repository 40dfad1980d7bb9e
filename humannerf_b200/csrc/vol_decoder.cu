// a7: the motion-weight volume decoder, forward and backward, hand-written (fp32 storage, deterministic).
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
// One tiled kernel (64 x 64 x 16 tiles; warp-level mma.sync m16n8k8 TF32 - 3xTF32 with fp32-level accuracy on the
// HN_PRECISION_FP32 parity path, plain TF32 on the HN_PRECISION_BF16 path) with three operand loaders (forward / data gradient /
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
constexpr int kTile = 64, kTileK = 16;

struct DeconvArgs {
  int din;                 // input volume side (output side = 2 din), a power of two
  int log2_din;
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

// Tensor-core tile product (warp-level mma.sync m16n8k8, TF32 operands, fp32 accumulate).  8 warps = 4 (rows) x 2 (cols):
// warp (wm, wn) owns rows [16 wm, +16) x cols [kTN/2 wn, +kTN/2) of the 64 x kTN tile = kTN/16 fragments of 16 x 8;
// acc[f][i]: fragment f, i = 0..3 -> (row g + 8 (i >> 1), col 2 t + (i & 1)) with g = lane >> 2, t = lane & 3.
// kX3 = true: every operand is split into two TF32 values (hi + lo) and the product is hi*hi + hi*lo + lo*hi ("3xTF32"):
// fp32-level accuracy for the 1e-4 parity path; kX3 = false: plain TF32 (the bf16 path, like cuDNN's TF32 engines).
// Operand tiles sit in shared memory in the layout their source is contiguous in, so every global -> shared copy is a
// 16-byte cp.async: "row-major in k" tiles [rows][16 k] with a row stride of 20 floats, "k-major" tiles [16 k][cols]
// with a row stride of cols + 8 floats - both strides make the fragment loads bank-conflict free.
constexpr int kStrideRK = kTileK + 4;      // 20
constexpr int kStrideKM = kTile + 8;       // 72 (a 32-column k-major tile uses 40)
constexpr int kTileFloats = kTile * kStrideRK > kTileK * kStrideKM ? kTile * kStrideRK : kTileK * kStrideKM;   // 1280
constexpr int kDecStages = 4;

// fp32 -> TF32 operand bits.  sm_100a has no native cvt.rna.tf32.f32 (ptxas expands it into ~8 integer / compare
// instructions per element, which made the tile product instruction-issue bound: 150 M warp instructions for the last
// layer's forward, 2.8 % of them MMAs).  The tensor core reads only the upper 19 bits of a TF32 operand, so adding half
// an ulp of the 10-bit mantissa to the bit pattern rounds to nearest (ties away from zero, like .rna) in one integer add.
__device__ __forceinline__ uint32_t to_tf32(float x) { return __float_as_uint(x) + 0x1000u; }
// the same value with the ignored bits cleared: what the tensor core multiplies with (needed to form the 3xTF32 residual)
__device__ __forceinline__ uint32_t to_tf32_exact(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// A_RK / B_RK: the tile is row-major in k ([row][k], stride 20) instead of k-major ([k][row], stride 72 / kTN + 8)
template <bool kX3, bool A_RK, bool B_RK, int kTN>
__device__ __forceinline__ void tile_mma(const float* __restrict__ As, const float* __restrict__ Bs, int warp, int lane,
                                         float (&acc)[kTN / 16][4]) {
  constexpr int kSB = kTN + 8, kWN = kTN / 2;
  const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int k8 = 0; k8 < kTileK; k8 += 8) {
    float af[4];
    if (A_RK) {
      af[0] = As[(wm * 16 + g) * kStrideRK + k8 + t];     af[1] = As[(wm * 16 + g + 8) * kStrideRK + k8 + t];
      af[2] = As[(wm * 16 + g) * kStrideRK + k8 + t + 4]; af[3] = As[(wm * 16 + g + 8) * kStrideRK + k8 + t + 4];
    } else {
      af[0] = As[(k8 + t) * kStrideKM + wm * 16 + g];     af[1] = As[(k8 + t) * kStrideKM + wm * 16 + g + 8];
      af[2] = As[(k8 + t + 4) * kStrideKM + wm * 16 + g]; af[3] = As[(k8 + t + 4) * kStrideKM + wm * 16 + g + 8];
    }
    uint32_t a_hi[4], a_lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a_hi[i] = kX3 ? to_tf32_exact(af[i]) : to_tf32(af[i]);
      if (kX3) a_lo[i] = to_tf32(af[i] - __uint_as_float(a_hi[i]));
    }
#pragma unroll
    for (int f = 0; f < kTN / 16; ++f) {
      float bf[2];
      if (B_RK) {
        bf[0] = Bs[(wn * kWN + f * 8 + g) * kStrideRK + k8 + t]; bf[1] = Bs[(wn * kWN + f * 8 + g) * kStrideRK + k8 + t + 4];
      } else {
        bf[0] = Bs[(k8 + t) * kSB + wn * kWN + f * 8 + g]; bf[1] = Bs[(k8 + t + 4) * kSB + wn * kWN + f * 8 + g];
      }
      uint32_t b_hi[2], b_lo[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        b_hi[i] = kX3 ? to_tf32_exact(bf[i]) : to_tf32(bf[i]);
        if (kX3) b_lo[i] = to_tf32(bf[i] - __uint_as_float(b_hi[i]));
      }
      if (kX3) {
        mma_tf32(acc[f], a_lo, b_hi);          // small terms first
        mma_tf32(acc[f], a_hi, b_lo);
      }
      mma_tf32(acc[f], a_hi, b_hi);
    }
  }
}

// 16 bytes global -> shared (zero-filled when !valid); sources that are not 16-byte aligned (rows of 25 floats) go
// element by element with `n_valid` leading elements copied and the rest zero-filled
__device__ __forceinline__ void cp_async16(float* dst, const float* src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async4x4(float* dst, const float* src, int n_valid) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + j)), "l"(j < n_valid ? src + j : src),
                 "r"(j < n_valid ? 4 : 0) : "memory");
}

// MODE 0: forward.  grid = (m_tiles * n_tiles, 8 parity classes, k_splits); GEMM rows m = input-grid voxel, cols = cout.
// MODE 1: data gradient.  grid = (m_tiles * n_tiles, 1, k_splits); rows = input voxel i, cols = cin, K = (tap, cout).
// MODE 2: weight gradient.  grid = (m_tiles * n_tiles, 64 taps, k_splits); rows = cin, cols = cout, K = input voxel.
// Tiles are 64 rows x kTN columns (kTN = 32 for the 25-channel output layer) x 16 k.  Four-stage cp.async pipeline: the
// copies of tile k + 3 are issued before the MMAs of tile k.  The K axis of MODE 0 / 1 is (tap, channel) with the channel
// count a multiple of 16, so a k-tile lies inside ONE tap: everything that depends on the tap (neighbour voxel, bounds
// test, weight slice) is recomputed only when the tap advances and a step costs two pointer additions; volume sides are
// powers of two (the decoder doubles from 1^3), so voxel <-> (z, y, x) is shifts and masks.
template <int MODE, bool kX3, int kTN>
__global__ void __launch_bounds__(256)
deconv_gemm_kernel(const DeconvArgs g) {
  __shared__ __align__(16) float As[kDecStages][kTileFloats];
  __shared__ __align__(16) float Bs[kDecStages][kTileFloats];
  constexpr bool A_RK = MODE != 2, B_RK = MODE == 1;
  constexpr int kSB = kTN + 8;                       // row stride of a k-major B tile
  static_assert(MODE != 1 || kTN == 64, "the data gradient always has >= 64 columns (cin)");
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sh = g.log2_din, din = 1 << sh, dmask = din - 1, din3 = 1 << (3 * sh), dout = 2 * din;
  const int M = MODE == 2 ? g.cin : din3;
  const int N = MODE == 0 ? g.cout : (MODE == 1 ? g.cin : g.cout);
  const int n_tiles = (N + kTN - 1) / kTN;
  const int m0 = (int)(blockIdx.x / n_tiles) * kTile, n0 = (int)(blockIdx.x % n_tiles) * kTN;
  // K geometry
  const int cout_pad = (g.cout + kTileK - 1) / kTileK * kTileK;          // dgrad: K per tap, zero padded
  const int n_taps = MODE == 0 ? (din == 1 ? 1 : 8) : (din == 1 ? 8 : 64);
  const int kpt = MODE == 0 ? g.cin : cout_pad;                          // K per tap (MODE 0 / 1)
  const int k_total = MODE == 2 ? din3 : n_taps * kpt;
  int k_begin = 0, k_end = k_total;
  if (g.k_splits > 1) {
    k_begin = (int)blockIdx.z * g.k_per_split;
    k_end = min(k_total, k_begin + g.k_per_split);
  }
  const int pz = (blockIdx.y >> 2) & 1, py = (blockIdx.y >> 1) & 1, px = blockIdx.y & 1;      // MODE 0: parity class
  const int wz = blockIdx.y >> 4, wy = (blockIdx.y >> 2) & 3, wx = blockIdx.y & 3;           // MODE 2: tap
  // copy roles.  row-major-in-k tiles: thread = (row tid >> 2, k quad (tid & 3) * 4); k-major A tile (64 columns):
  // (k tid >> 4, quad (tid & 15) * 4); k-major B tile: the same for 64 columns, (k tid >> 3, quad (tid & 7) * 4) of the
  // first 128 threads for 32 columns
  const int r4 = tid >> 2, kq = (tid & 3) * 4, kr = tid >> 4, q16 = (tid & 15) * 4;
  const int krb = kTN == 64 ? kr : tid >> 3, qb = kTN == 64 ? q16 : (tid & 7) * 4;
  const bool b_thread = kTN == 64 || tid < 128;
  const bool w_rows_aligned = (g.cout & 3) == 0;
  const int vm = m0 + r4;                            // MODE 0 / 1: voxel of this thread's A row
  const int vz = vm >> (2 * sh), vy = (vm >> sh) & dmask, vx = vm & dmask;
  const bool row_ok = vm < din3;

  // issue cursor (uniform over the block): tap and channel offset of the next k-tile to copy
  int tap_i = 0, c_i = 0, k_i = k_begin;
  if (MODE != 2) { tap_i = k_begin / kpt; c_i = k_begin - tap_i * kpt; }
  const float* a_tap = nullptr;                      // A row of this thread for the current tap (+ its k quad)
  const float* b_tap = nullptr;
  bool a_ok = false;
  int b_cols = 0;                                    // MODE 0 / 2: valid columns of this thread's B quad
  if (MODE != 1) b_cols = b_thread ? max(0, min(4, N - (n0 + qb))) : 0;
  const bool b_row_ok = MODE == 1 && n0 + r4 < N;
  auto set_tap = [&](int tap) {
    if (MODE == 0) {
      const int iz = vz + tap_off(pz, (tap >> 2) & 1), iy = vy + tap_off(py, (tap >> 1) & 1), ix = vx + tap_off(px, tap & 1);
      a_ok = row_ok && (unsigned)iz < (unsigned)din && (unsigned)iy < (unsigned)din && (unsigned)ix < (unsigned)din;
      a_tap = g.in + (size_t)((((iz << sh) + iy) << sh) + ix) * g.cin + kq;
      const int kidx = (tap_k(pz, (tap >> 2) & 1) * 4 + tap_k(py, (tap >> 1) & 1)) * 4 + tap_k(px, tap & 1);
      b_tap = g.w + ((size_t)krb * 64 + kidx) * g.cout + n0 + qb;
    } else if (MODE == 1) {
      // tap list: all 64 taps, or (din == 1) the 8 taps with k in {1, 2} per axis - the only ones that reach the 2^3 output
      const int kz = din == 1 ? 1 + ((tap >> 2) & 1) : tap >> 4, ky = din == 1 ? 1 + ((tap >> 1) & 1) : (tap >> 2) & 3,
                kx = din == 1 ? 1 + (tap & 1) : tap & 3;
      const int oz = 2 * vz - 1 + kz, oy = 2 * vy - 1 + ky, ox = 2 * vx - 1 + kx;
      a_ok = row_ok && (unsigned)oz < (unsigned)dout && (unsigned)oy < (unsigned)dout && (unsigned)ox < (unsigned)dout;
      a_tap = g.outside + (size_t)((((oz << (sh + 1)) + oy) << (sh + 1)) + ox) * g.ld_out + kq;
      b_tap = g.w + ((size_t)(n0 + r4) * 64 + (kz * 4 + ky) * 4 + kx) * g.cout + kq;
    }
  };
  if (MODE != 2) set_tap(tap_i);

  auto issue = [&](int stage) {
    float* as = As[stage];
    float* bs = Bs[stage];
    if (MODE == 0) {
      // A(m, kk) = in[i(m, tap)][ci .. ci + 3];  B(kk, n) = w[ci][k(p, tap)][n .. n + 3]
      cp_async16(as + r4 * kStrideRK + kq, a_ok ? a_tap + c_i : g.in, a_ok);
      if (b_thread) {
        const float* src = b_tap + (size_t)c_i * 64 * g.cout;
        if (w_rows_aligned) cp_async16(bs + krb * kSB + qb, b_cols == 4 ? src : g.w, b_cols == 4);
        else cp_async4x4(bs + krb * kSB + qb, b_cols > 0 ? src : g.w, b_cols);
      }
    } else if (MODE == 1) {
      // A(i, kk) = outside[o(i, tap)][co .. co + 3], kk = tap * cout_pad + co;  B(ci, kk) = w[ci][tap][co .. co + 3]
      // ld_out is a multiple of 4 and columns [cout, ld_out) of the gradient buffers are zero
      const int co = c_i + kq;
      const bool av = a_ok && co + 4 <= g.ld_out;
      cp_async16(as + r4 * kStrideRK + kq, av ? a_tap + c_i : g.outside, av);
      const int bv = b_row_ok ? max(0, min(4, g.cout - co)) : 0;
      if (w_rows_aligned) cp_async16(bs + r4 * kStrideRK + kq, bv == 4 ? b_tap + c_i : g.w, bv == 4);
      else cp_async4x4(bs + r4 * kStrideRK + kq, bv > 0 ? b_tap + c_i : g.w, bv);
    } else {
      // A(ci, i) = in[i][ci .. ci + 3];  B(i, co) = outside[o(i, tap)][co .. co + 3]
      const int ia = k_i + kr;
      const bool aok = ia < k_end && m0 + q16 + 4 <= M;
      cp_async16(as + kr * kStrideKM + q16, aok ? g.in + (size_t)ia * g.cin + m0 + q16 : g.in, aok);
      if (b_thread) {
        const int i = k_i + krb;
        const int oz = 2 * (i >> (2 * sh)) - 1 + wz, oy = 2 * ((i >> sh) & dmask) - 1 + wy, ox = 2 * (i & dmask) - 1 + wx;
        const bool bok = i < k_end && (unsigned)oz < (unsigned)dout && (unsigned)oy < (unsigned)dout && (unsigned)ox < (unsigned)dout &&
                         n0 + qb + 4 <= g.ld_out;
        cp_async16(bs + krb * kSB + qb, bok ? g.outside + (size_t)((((oz << (sh + 1)) + oy) << (sh + 1)) + ox) * g.ld_out + n0 + qb : g.outside,
                   bok);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    k_i += kTileK;
    if (MODE != 2) {
      c_i += kTileK;
      if (c_i == kpt) {
        c_i = 0;
        if (++tap_i < n_taps) set_tap(tap_i);
      }
    }
  };

  float acc[kTN / 16][4];
#pragma unroll
  for (int i = 0; i < kTN / 16; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int n_steps = (k_end - k_begin + kTileK - 1) / kTileK;
#pragma unroll
  for (int s = 0; s < kDecStages - 1; ++s) {
    if (s < n_steps) issue(s);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
#pragma unroll 1
  for (int step = 0; step < n_steps; ++step) {
    asm volatile("cp.async.wait_group %0;" ::"n"(kDecStages - 2) : "memory");
    __syncthreads();                                   // tile `step` has landed for every thread; tile step - 1 is no longer read
    const int nxt = step + kDecStages - 1;
    if (nxt < n_steps) issue(nxt % kDecStages);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    const float* as = As[step % kDecStages];
    const float* bs = Bs[step % kDecStages];
    if (kX3) {
      // tensor-core accumulation truncates: keep its error inside one 16-long partial sum and add the tiles up in
      // round-to-nearest fp32
      float part[kTN / 16][4];
#pragma unroll
      for (int i = 0; i < kTN / 16; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
      tile_mma<true, A_RK, B_RK, kTN>(as, bs, warp, lane, part);
#pragma unroll
      for (int i = 0; i < kTN / 16; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
    } else {
      tile_mma<false, A_RK, B_RK, kTN>(as, bs, warp, lane, acc);
    }
  }

  // ---- epilogue: fragment f, element e of this lane -> (row, col) of the tile
  const int wm = warp >> 1, wn = warp & 1, fg = lane >> 2, ft = lane & 3;
#pragma unroll
  for (int i = 0; i < kTN / 16; ++i) {                // i = fragment
#pragma unroll
    for (int j = 0; j < 4; ++j) {                     // j = element of the fragment
      const int m = m0 + wm * 16 + fg + 8 * (j >> 1);
      const int n = n0 + wn * (kTN / 2) + i * 8 + 2 * ft + (j & 1);
      if (m >= M || n >= N) continue;
      float v = acc[i][j];
      if (MODE == 0) {
        const int mz = m >> (2 * sh), my = (m >> sh) & dmask, mx = m & dmask;
        const size_t o = (size_t)(((((2 * mz + pz) << (sh + 1)) + 2 * my + py) << (sh + 1)) + 2 * mx + px);
        if (g.k_splits > 1) {
          g.dst[((size_t)blockIdx.z * 8 * din3 + o) * g.cout + n] = v;
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
        // k_splits > 1: partial[split] has the layout of dW itself
        g.dst[(size_t)blockIdx.z * g.cin * 64 * g.cout + ((size_t)m * 64 + blockIdx.y) * g.cout + n] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The same three GEMMs with 128 x kTN x 32 tiles for the layers with >= 512 rows / K (the 8^3 -> 16^3 and 16^3 -> 32^3
// layers hold 85 % of the decoder's FLOPs).  8 warps = 4 (rows) x 2 (cols), warp tile 32 x kTN/2 = 2 x kTN/16 fragments:
// per 8-wide k step a warp issues 8 + kTN/8 shared loads for 2 kTN/16 MMAs (the 64 x 64 x 16 kernel above: 12 for 4),
// a k-tile is 32 deep (half the barriers), three stages of 27 KB in dynamic shared memory, two blocks per SM.
// Operand layouts, tap handling and epilogues are those of deconv_gemm_kernel.
constexpr int kBM = 128, kBK = 32, kBStages = 3;
constexpr int kBStrideRK = kBK + 4;        // 36: rows of a row-major-in-k tile
constexpr int kBStrideAKM = kBM + 8;       // 136: rows of the k-major A tile of the weight gradient
constexpr int kBAFloats = kBM * kBStrideRK;                        // 4608 >= 32 * 136
__host__ __device__ constexpr int big_b_floats(int tn) { return tn == 64 ? 64 * kBStrideRK : kBK * (tn + 8); }   // 2304 (= 32 * 72) / 1280
__host__ __device__ constexpr int big_smem_bytes(int tn) { return kBStages * (kBAFloats + big_b_floats(tn)) * 4; }

template <int MODE, bool kX3, int kTN>
__global__ void __launch_bounds__(256, 2)
deconv_gemm_big_kernel(const DeconvArgs g) {
  extern __shared__ __align__(16) float big_smem[];
  constexpr bool A_RK = MODE != 2, B_RK = MODE == 1;
  constexpr int kSB = kTN + 8, kNF = kTN / 16, kBF = big_b_floats(kTN);
  static_assert(MODE != 1 || kTN == 64, "the data gradient always has >= 64 columns (cin)");
  float* const As = big_smem;
  float* const Bs = big_smem + kBStages * kBAFloats;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sh = g.log2_din, din = 1 << sh, dmask = din - 1, din3 = 1 << (3 * sh), dout = 2 * din;
  const int M = MODE == 2 ? g.cin : din3;
  const int N = MODE == 0 ? g.cout : (MODE == 1 ? g.cin : g.cout);
  const int n_tiles = (N + kTN - 1) / kTN;
  const int m0 = (int)(blockIdx.x / n_tiles) * kBM, n0 = (int)(blockIdx.x % n_tiles) * kTN;
  const int cout_pad = (g.cout + kBK - 1) / kBK * kBK;                    // dgrad: K per tap, zero padded
  const int n_taps = MODE == 0 ? (din == 1 ? 1 : 8) : (din == 1 ? 8 : 64);
  const int kpt = MODE == 0 ? g.cin : cout_pad;                          // K per tap (MODE 0 / 1), a multiple of 32
  const int k_total = MODE == 2 ? din3 : n_taps * kpt;
  int k_begin = 0, k_end = k_total;
  if (g.k_splits > 1) {
    k_begin = (int)blockIdx.z * g.k_per_split;
    k_end = min(k_total, k_begin + g.k_per_split);
  }
  const int pz = (blockIdx.y >> 2) & 1, py = (blockIdx.y >> 1) & 1, px = blockIdx.y & 1;      // MODE 0: parity class
  const int wz = blockIdx.y >> 4, wy = (blockIdx.y >> 2) & 3, wx = blockIdx.y & 3;           // MODE 2: tap
  // copy roles.  Row-major-in-k tiles (8 quads per row): thread = (row (tid >> 3) + 32 j, quad (tid & 7) * 4);
  // k-major A tile (32 quads per k row): (k (tid >> 5) + 8 j, quad (tid & 31) * 4); k-major B tile: 64 columns (k (tid >> 4)
  // + 16 j, quad (tid & 15) * 4), 32 columns (k tid >> 3, quad (tid & 7) * 4)
  const int r8 = tid >> 3, kq = (tid & 7) * 4;
  const int krb = kTN == 64 ? tid >> 4 : tid >> 3, qb = kTN == 64 ? (tid & 15) * 4 : (tid & 7) * 4;
  constexpr int kBJ = kTN == 64 ? 2 : 1;                                  // k rows of the k-major B tile per thread
  const bool w_rows_aligned = (g.cout & 3) == 0;

  int tap_i = 0, c_i = 0, k_i = k_begin;
  if (MODE != 2) { tap_i = k_begin / kpt; c_i = k_begin - tap_i * kpt; }
  const float* a_tap[4] = {nullptr, nullptr, nullptr, nullptr};          // A rows of this thread for the current tap (+ k quad)
  const float* b_tap[2] = {nullptr, nullptr};
  uint32_t a_ok = 0u;                                                    // bit j: row j of this thread is inside the volume
  int b_cols = 0;                                                        // MODE 0 / 2: valid columns of this thread's B quad
  if (MODE != 1) b_cols = max(0, min(4, N - (n0 + qb)));
  auto set_tap = [&](int tap) {
    if (MODE == 2) return;
    a_ok = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int vm = m0 + r8 + 32 * j;
      const int vz = vm >> (2 * sh), vy = (vm >> sh) & dmask, vx = vm & dmask;
      if (MODE == 0) {
        const int iz = vz + tap_off(pz, (tap >> 2) & 1), iy = vy + tap_off(py, (tap >> 1) & 1), ix = vx + tap_off(px, tap & 1);
        if (vm < din3 && (unsigned)iz < (unsigned)din && (unsigned)iy < (unsigned)din && (unsigned)ix < (unsigned)din) a_ok |= 1u << j;
        a_tap[j] = g.in + (size_t)((((iz << sh) + iy) << sh) + ix) * g.cin + kq;
      } else {
        const int kz = din == 1 ? 1 + ((tap >> 2) & 1) : tap >> 4, ky = din == 1 ? 1 + ((tap >> 1) & 1) : (tap >> 2) & 3,
                  kx = din == 1 ? 1 + (tap & 1) : tap & 3;
        const int oz = 2 * vz - 1 + kz, oy = 2 * vy - 1 + ky, ox = 2 * vx - 1 + kx;
        if (vm < din3 && (unsigned)oz < (unsigned)dout && (unsigned)oy < (unsigned)dout && (unsigned)ox < (unsigned)dout) a_ok |= 1u << j;
        a_tap[j] = g.outside + (size_t)((((oz << (sh + 1)) + oy) << (sh + 1)) + ox) * g.ld_out + kq;
      }
    }
    if (MODE == 0) {
      const int kidx = (tap_k(pz, (tap >> 2) & 1) * 4 + tap_k(py, (tap >> 1) & 1)) * 4 + tap_k(px, tap & 1);
#pragma unroll
      for (int j = 0; j < kBJ; ++j) b_tap[j] = g.w + ((size_t)(krb + 16 * j) * 64 + kidx) * g.cout + n0 + qb;
    } else {
      const int kz = din == 1 ? 1 + ((tap >> 2) & 1) : tap >> 4, ky = din == 1 ? 1 + ((tap >> 1) & 1) : (tap >> 2) & 3,
                kx = din == 1 ? 1 + (tap & 1) : tap & 3;
#pragma unroll
      for (int j = 0; j < 2; ++j) b_tap[j] = g.w + ((size_t)(n0 + r8 + 32 * j) * 64 + (kz * 4 + ky) * 4 + kx) * g.cout + kq;
    }
  };
  if (MODE != 2) set_tap(tap_i);

  auto issue = [&](int stage) {
    float* as = As + stage * kBAFloats;
    float* bs = Bs + stage * kBF;
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool ok = (a_ok >> j) & 1u;
        cp_async16(as + (r8 + 32 * j) * kBStrideRK + kq, ok ? a_tap[j] + c_i : g.in, ok);
      }
#pragma unroll
      for (int j = 0; j < kBJ; ++j) {
        const float* src = b_tap[j] + (size_t)c_i * 64 * g.cout;
        float* dst = bs + (krb + 16 * j) * kSB + qb;
        if (w_rows_aligned) cp_async16(dst, b_cols == 4 ? src : g.w, b_cols == 4);
        else cp_async4x4(dst, b_cols > 0 ? src : g.w, b_cols);
      }
    } else if (MODE == 1) {
      const int co = c_i + kq;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool av = ((a_ok >> j) & 1u) && co + 4 <= g.ld_out;
        cp_async16(as + (r8 + 32 * j) * kBStrideRK + kq, av ? a_tap[j] + c_i : g.outside, av);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int bv = (n0 + r8 + 32 * j < N) ? max(0, min(4, g.cout - co)) : 0;
        float* dst = bs + (r8 + 32 * j) * kBStrideRK + kq;
        if (w_rows_aligned) cp_async16(dst, bv == 4 ? b_tap[j] + c_i : g.w, bv == 4);
        else cp_async4x4(dst, bv > 0 ? b_tap[j] + c_i : g.w, bv);
      }
    } else {
      // A(ci, i) = in[i][ci .. ci + 3];  B(i, co) = outside[o(i, tap)][co .. co + 3]
      const int ka = tid >> 5, qa = (tid & 31) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ia = k_i + ka + 8 * j;
        const bool aok = ia < k_end && m0 + qa + 4 <= M;
        cp_async16(as + (ka + 8 * j) * kBStrideAKM + qa, aok ? g.in + (size_t)ia * g.cin + m0 + qa : g.in, aok);
      }
#pragma unroll
      for (int j = 0; j < kBJ; ++j) {
        const int i = k_i + krb + 16 * j;
        const int oz = 2 * (i >> (2 * sh)) - 1 + wz, oy = 2 * ((i >> sh) & dmask) - 1 + wy, ox = 2 * (i & dmask) - 1 + wx;
        const bool bok = i < k_end && (unsigned)oz < (unsigned)dout && (unsigned)oy < (unsigned)dout && (unsigned)ox < (unsigned)dout &&
                         n0 + qb + 4 <= g.ld_out;
        cp_async16(bs + (krb + 16 * j) * kSB + qb,
                   bok ? g.outside + (size_t)((((oz << (sh + 1)) + oy) << (sh + 1)) + ox) * g.ld_out + n0 + qb : g.outside, bok);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    k_i += kBK;
    if (MODE != 2) {
      c_i += kBK;
      if (c_i == kpt) {
        c_i = 0;
        if (++tap_i < n_taps) set_tap(tap_i);
      }
    }
  };

  float acc[2][kNF][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int i = 0; i < kNF; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[a][i][j] = 0.f;

  const int wm = warp >> 1, wn = warp & 1, fg = lane >> 2, ft = lane & 3;
  const int n_steps = (k_end - k_begin + kBK - 1) / kBK;
#pragma unroll
  for (int s = 0; s < kBStages - 1; ++s) {
    if (s < n_steps) issue(s);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
#pragma unroll 1
  for (int step = 0; step < n_steps; ++step) {
    asm volatile("cp.async.wait_group %0;" ::"n"(kBStages - 2) : "memory");
    __syncthreads();
    const int nxt = step + kBStages - 1;
    if (nxt < n_steps) issue(nxt % kBStages);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    const float* as = As + (step % kBStages) * kBAFloats;
    const float* bs = Bs + (step % kBStages) * kBF;
    float part[2][kNF][4];
    if (kX3) {
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int i = 0; i < kNF; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) part[a][i][j] = 0.f;
    }
#pragma unroll
    for (int k8 = 0; k8 < kBK; k8 += 8) {
      uint32_t a_hi[2][4], a_lo[2][4];
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int r = wm * 32 + a * 16 + fg;
        float af[4];
        if (A_RK) {
          af[0] = as[r * kBStrideRK + k8 + ft];     af[1] = as[(r + 8) * kBStrideRK + k8 + ft];
          af[2] = as[r * kBStrideRK + k8 + ft + 4]; af[3] = as[(r + 8) * kBStrideRK + k8 + ft + 4];
        } else {
          af[0] = as[(k8 + ft) * kBStrideAKM + r];     af[1] = as[(k8 + ft) * kBStrideAKM + r + 8];
          af[2] = as[(k8 + ft + 4) * kBStrideAKM + r]; af[3] = as[(k8 + ft + 4) * kBStrideAKM + r + 8];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a_hi[a][i] = kX3 ? to_tf32_exact(af[i]) : to_tf32(af[i]);
          if (kX3) a_lo[a][i] = to_tf32(af[i] - __uint_as_float(a_hi[a][i]));
        }
      }
#pragma unroll
      for (int f = 0; f < kNF; ++f) {
        const int c = wn * (kTN / 2) + f * 8 + fg;
        float bf[2];
        if (B_RK) { bf[0] = bs[c * kBStrideRK + k8 + ft]; bf[1] = bs[c * kBStrideRK + k8 + ft + 4]; }
        else { bf[0] = bs[(k8 + ft) * kSB + c]; bf[1] = bs[(k8 + ft + 4) * kSB + c]; }
        uint32_t b_hi[2], b_lo[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          b_hi[i] = kX3 ? to_tf32_exact(bf[i]) : to_tf32(bf[i]);
          if (kX3) b_lo[i] = to_tf32(bf[i] - __uint_as_float(b_hi[i]));
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          if (kX3) {
            mma_tf32(part[a][f], a_lo[a], b_hi);
            mma_tf32(part[a][f], a_hi[a], b_lo);
            mma_tf32(part[a][f], a_hi[a], b_hi);
          } else {
            mma_tf32(acc[a][f], a_hi[a], b_hi);
          }
        }
      }
    }
    if (kX3) {
      // tensor-core accumulation truncates: its error stays inside one 32-long partial sum, the tiles add up in fp32
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int i = 0; i < kNF; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[a][i][j] += part[a][i][j];
    }
  }

#pragma unroll
  for (int a = 0; a < 2; ++a) {
#pragma unroll
    for (int i = 0; i < kNF; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int m = m0 + wm * 32 + a * 16 + fg + 8 * (j >> 1);
        const int n = n0 + wn * (kTN / 2) + i * 8 + 2 * ft + (j & 1);
        if (m >= M || n >= N) continue;
        float v = acc[a][i][j];
        if (MODE == 0) {
          const int mz = m >> (2 * sh), my = (m >> sh) & dmask, mx = m & dmask;
          const size_t o = (size_t)(((((2 * mz + pz) << (sh + 1)) + 2 * my + py) << (sh + 1)) + 2 * mx + px);
          if (g.k_splits > 1) {
            g.dst[((size_t)blockIdx.z * 8 * din3 + o) * g.cout + n] = v;
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
          g.dst[(size_t)blockIdx.z * g.cin * 64 * g.cout + ((size_t)m * 64 + blockIdx.y) * g.cout + n] = v;
        }
      }
    }
  }
}

// rows of a GEMM tile: the 128-row kernel for GEMMs with >= 512 rows (forward / data gradient: input voxels) or a K of
// >= 512 voxels and >= 128 rows (weight gradient), when the K extent per tap is a multiple of its 32-deep k-tile
static bool use_big_tiles(int mode, int din, int cin, int cout) {
  const int d3 = din * din * din;
  if (d3 < 512) return false;
  if (mode == 0) return cin % kBK == 0;
  if (mode == 1) return true;                 // K per tap is padded to 32 by the kernel itself
  return cin >= kBM;
}
static int tile_m_rows(bool big) { return big ? kBM : kTile; }
static int tile_k(bool big) { return big ? kBK : kTileK; }

template <int MODE, bool kX3, int kTN>
static int launch_big(const DeconvArgs& g, dim3 grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    HN_TRY(check_cuda(cudaFuncSetAttribute(deconv_gemm_big_kernel<MODE, kX3, kTN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           big_smem_bytes(kTN)), "hn_vol_decoder/smem attr"));
    attr_set = true;
  }
  deconv_gemm_big_kernel<MODE, kX3, kTN><<<grid, 256, big_smem_bytes(kTN), st>>>(g);
  return 0;
}

template <int MODE>
static int launch_deconv_gemm(const DeconvArgs& g, dim3 grid, bool x3, int tn, bool big, cudaStream_t st, const char* what) {
  constexpr int M32 = MODE == 1 ? 0 : MODE;          // the data gradient has no 32-column variant
  if (big) {
    if (MODE != 1 && tn == 32) {
      if (x3) HN_TRY((launch_big<M32, true, 32>(g, grid, st)));
      else HN_TRY((launch_big<M32, false, 32>(g, grid, st)));
    } else {
      if (x3) HN_TRY((launch_big<MODE, true, 64>(g, grid, st)));
      else HN_TRY((launch_big<MODE, false, 64>(g, grid, st)));
    }
    return launched(what);
  }
  if (MODE != 1 && tn == 32) {
    if (x3) deconv_gemm_kernel<M32, true, 32><<<grid, 256, 0, st>>>(g);
    else deconv_gemm_kernel<M32, false, 32><<<grid, 256, 0, st>>>(g);
  } else {
    if (x3) deconv_gemm_kernel<MODE, true, 64><<<grid, 256, 0, st>>>(g);
    else deconv_gemm_kernel<MODE, false, 64><<<grid, 256, 0, st>>>(g);
  }
  return launched(what);
}
// column tile of a GEMM with `n` columns
static int tile_n(int n) { return n <= 32 ? 32 : 64; }

// Weight gradient of the FIRST transposed convolution (1^3 input): dW[ci][k][co] = in[ci] * g[o(k)][co] for the 8 taps
// k in {1, 2}^3 that reach the 2^3 output (o = k - 1 per axis) and 0 for the other 56 - an outer product written once,
// 16 bytes per thread in the parameter's [cin][64][cout] order (the generic kernel ran 8192 blocks with K = 1 for it).
__global__ void __launch_bounds__(256) deconv_wgrad_din1_kernel(int cin, int cout, int ld_g, const float* __restrict__ in,
                                                                const float* __restrict__ g, float4* __restrict__ dw) {
  const int q4 = cout >> 2;                                   // float4 per (ci, tap) row
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)cin * 64 * q4) return;
  const int c4 = (int)(idx % q4);
  const int tap = (int)((idx / q4) & 63);
  const int ci = (int)(idx / ((int64_t)q4 * 64));
  const int kz = tap >> 4, ky = (tap >> 2) & 3, kx = tap & 3;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (((kz - 1) | (ky - 1) | (kx - 1)) >= 0 && kz <= 2 && ky <= 2 && kx <= 2) {
    const int o = ((kz - 1) * 2 + (ky - 1)) * 2 + (kx - 1);
    const float a = __ldg(in + ci);
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g + (size_t)o * ld_g) + c4);
    v = make_float4(a * gg.x, a * gg.y, a * gg.z, a * gg.w);
  }
  dw[idx] = v;
}

// dst[i] = sum_s partial[s][i] (weight gradients of the layers whose K = voxel count was split)
__global__ void sum_partials_kernel(const float4* __restrict__ partial, int splits, int64_t n4, float4* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 s = partial[i];
  for (int k = 1; k < splits; ++k) {
    const float4 p = partial[(size_t)k * n4 + i];
    s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
  }
  dst[i] = s;
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
// d_e[i] = sum_j W[j][i] dpre[j]: block = 32 inputs x 32 slices of j (coalesced rows of W), slices added in a fixed order
__global__ void __launch_bounds__(1024) linear_dgrad_kernel(int n_out, int n_in, const float* __restrict__ w, const float* __restrict__ dpre,
                                                            float* __restrict__ d_e) {
  __shared__ float red[32][33];
  const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (i < n_in) {
#pragma unroll 4
    for (int j = sl; j < n_out; j += 32) s = fmaf(__ldg(w + (size_t)j * n_in + i), dpre[j], s);
  }
  red[sl][lane] = s;
  __syncthreads();
  if (sl == 0 && i < n_in) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += red[k][lane];
    d_e[i] = t;
  }
}
// d_b[c] = sum_rows g[row][c] for c < n_cols.  grid = (row slabs, chunks of 128 columns).  Block (b, y) sums its slab of
// rows with coalesced row reads (thread = column lane x row lane), writes partial[b][.]; the block of a chunk that finishes
// last adds the partials in slab order: no floating-point atomics, bit-reproducible.  `counter[y]` must be zero at launch
// (cleared by the caller once per backward pass).
constexpr int kColSumMaxLd = 128;          // columns per chunk
constexpr int kColSumMaxChunks = 4;        // => at most 512 channels on the output side of a convolution
__global__ void __launch_bounds__(256) column_sum_kernel(int64_t rows, int ld, int n_cols, const float* __restrict__ g,
                                                         float* __restrict__ partial, unsigned int* __restrict__ counter,
                                                         float* __restrict__ out) {
  __shared__ float red[8][kColSumMaxLd];
  __shared__ bool is_last;
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5, n_slabs = gridDim.x, c0 = blockIdx.y * kColSumMaxLd;
  const int64_t r0 = rows * blockIdx.x / n_slabs, r1 = rows * (blockIdx.x + 1) / n_slabs;
  float s[kColSumMaxLd / 32] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int64_t r = r0 + rl; r < r1; r += 8) {
#pragma unroll
    for (int q = 0; q < kColSumMaxLd / 32; ++q)
      if (c0 + cl + 32 * q < ld) s[q] += g[r * ld + c0 + cl + 32 * q];
  }
#pragma unroll
  for (int q = 0; q < kColSumMaxLd / 32; ++q) red[rl][cl + 32 * q] = s[q];
  __syncthreads();
  const int c = c0 + (int)threadIdx.x;
  if (threadIdx.x < kColSumMaxLd && c < ld) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    partial[(size_t)blockIdx.x * ld + c] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(counter + blockIdx.y, 1u) == (unsigned)(n_slabs - 1);
  __syncthreads();
  if (is_last) {
    // the same two-level order for the partials: row lane rl takes slabs rl, rl + 8, ..., then the 8 lanes in order
    __threadfence();
#pragma unroll
    for (int q = 0; q < kColSumMaxLd / 32; ++q) s[q] = 0.f;
#pragma unroll 4
    for (int b = rl; b < n_slabs; b += 8) {
#pragma unroll
      for (int q = 0; q < kColSumMaxLd / 32; ++q)
        if (c0 + cl + 32 * q < ld) s[q] += __ldcg(partial + (size_t)b * ld + c0 + cl + 32 * q);
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kColSumMaxLd / 32; ++q) red[rl][cl + 32 * q] = s[q];
    __syncthreads();
    if (threadIdx.x < kColSumMaxLd && c < n_cols) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
      out[c] = t;
    }
  }
}

// vol[c][voxel] = softmax_c(logits[voxel][c] + log prior[c][voxel]); one thread per voxel, logits rows of kLdLogits = 32
// floats read as 8 x 16 bytes, everything in registers
constexpr int kLdLogits = 32;
__global__ void __launch_bounds__(128) softmax_prior_fwd_kernel(int n_vox, int C, const float* __restrict__ logits,
                                                                const float* __restrict__ prior, float* __restrict__ vol) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vox) return;
  float x[kLdLogits];
  const float4* row = reinterpret_cast<const float4*>(logits + (size_t)v * kLdLogits);
#pragma unroll
  for (int q = 0; q < kLdLogits / 4; ++q) {
    const float4 t = row[q];
    x[4 * q] = t.x; x[4 * q + 1] = t.y; x[4 * q + 2] = t.z; x[4 * q + 3] = t.w;
  }
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < kLdLogits; ++c)
    if (c < C) {
      x[c] += logf(prior[(size_t)c * n_vox + v]);
      mx = fmaxf(mx, x[c]);
    }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kLdLogits; ++c)
    if (c < C) { x[c] = expf(x[c] - mx); s += x[c]; }
  const float inv = 1.0f / s;
#pragma unroll
  for (int c = 0; c < kLdLogits; ++c)
    if (c < C) vol[(size_t)c * n_vox + v] = x[c] * inv;
}
// d_logits[voxel][c] = y_c (g_c - sum_k g_k y_k); columns [C, 32) are written as zeros
__global__ void __launch_bounds__(128) softmax_prior_bwd_kernel(int n_vox, int C, const float* __restrict__ vol,
                                                                const float* __restrict__ d_vol, float* __restrict__ d_logits) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vox) return;
  float y[kLdLogits], gy[kLdLogits];
  float dot = 0.f;
#pragma unroll
  for (int c = 0; c < kLdLogits; ++c) {
    y[c] = 0.f; gy[c] = 0.f;
    if (c < C) {
      y[c] = vol[(size_t)c * n_vox + v];
      gy[c] = d_vol[(size_t)c * n_vox + v];
      dot = fmaf(gy[c], y[c], dot);
    }
  }
  float4* row = reinterpret_cast<float4*>(d_logits + (size_t)v * kLdLogits);
#pragma unroll
  for (int q = 0; q < kLdLogits / 4; ++q)
    row[q] = make_float4(y[4 * q] * (gy[4 * q] - dot), y[4 * q + 1] * (gy[4 * q + 1] - dot), y[4 * q + 2] * (gy[4 * q + 2] - dot),
                         y[4 * q + 3] * (gy[4 * q + 3] - dot));
}

// ------------------------------------------------------------------------------------------------------------------
struct DecPlan {
  int n_conv, embed, hidden;
  int ch[kDecMaxConv + 1];
  int din[kDecMaxConv];
  int ld_logits;
  int fwd_splits[kDecMaxConv], dgrad_splits[kDecMaxConv], wgrad_splits[kDecMaxConv];   // K splits of the three GEMMs per layer
  size_t off_h, off_act[kDecMaxConv + 1], off_logits, ws_floats;          // saved for backward
  size_t off_partial, off_partial_w, off_g[3], off_dlogits, off_colsum, off_counters, scratch_floats;
};

constexpr int kColSumSlabs = 128;
// K-splits for a GEMM with `tiles` output tiles: the low layers are weight-bandwidth bound and need many CTAs; the large
// layers also run faster in ~4 blocks per SM than in one wave (measured: 74 against 83 us for the last layer's forward)
static int pick_splits(int64_t tiles, int k_total) {
  int splits = 1;
  while (tiles * splits < 592 && splits < 16 && k_total / (splits * 2) >= 64) splits *= 2;
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
    if (l > 0 && channels[l] > kColSumMaxLd * kColSumMaxChunks)
      return fail_arg("hn_vol_decoder: at most %d output channels per convolution", kColSumMaxLd * kColSumMaxChunks);
  }
  if (p.ch[n_conv] > 32) return fail_arg("hn_vol_decoder: at most 32 output channels");
  for (int l = 0; l < n_conv; ++l) p.din[l] = 1 << l;
  p.ld_logits = kLdLogits;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += (n + 63) / 64 * 64; return o; };
  p.off_h = 0;
  // act[l] = input of convolution l (act[0] = h), channel-last
  for (int l = 0; l < n_conv; ++l) p.off_act[l] = take((size_t)p.din[l] * p.din[l] * p.din[l] * p.ch[l]);
  const size_t n_vox = (size_t)8 * p.din[n_conv - 1] * p.din[n_conv - 1] * p.din[n_conv - 1];
  p.off_logits = take(n_vox * p.ld_logits);
  p.ws_floats = off;
  // scratch: split-K partials (forward and data gradient), two gradient buffers, d_logits
  size_t part = 0, part_w = 0, gmax = 0;
  for (int l = 0; l < n_conv; ++l) {
    const size_t d3 = (size_t)p.din[l] * p.din[l] * p.din[l];
    const bool fb = use_big_tiles(0, p.din[l], p.ch[l], p.ch[l + 1]), db = use_big_tiles(1, p.din[l], p.ch[l], p.ch[l + 1]),
               wb = use_big_tiles(2, p.din[l], p.ch[l], p.ch[l + 1]);
    const int64_t ft = (int64_t)ceil_div<int64_t>((int64_t)d3, tile_m_rows(fb)) * ceil_div(p.ch[l + 1], tile_n(p.ch[l + 1])) * 8;
    p.fwd_splits[l] = pick_splits(ft, (p.din[l] == 1 ? 1 : 8) * p.ch[l]);
    const int64_t dt = (int64_t)ceil_div<int64_t>((int64_t)d3, tile_m_rows(db)) * ceil_div(p.ch[l], kTile);
    p.dgrad_splits[l] = pick_splits(dt, (p.din[l] == 1 ? 8 : 64) * (ceil_div(p.ch[l + 1], tile_k(db)) * tile_k(db)));
    const int64_t wt = (int64_t)ceil_div(p.ch[l], tile_m_rows(wb)) * ceil_div(p.ch[l + 1], tile_n(p.ch[l + 1])) * 64;
    p.wgrad_splits[l] = pick_splits(wt, (int)d3);
    if (p.wgrad_splits[l] > 1) part_w = std::max(part_w, (size_t)p.wgrad_splits[l] * p.ch[l] * 64 * p.ch[l + 1]);
    if (p.fwd_splits[l] > 1) part = std::max(part, (size_t)p.fwd_splits[l] * 8 * d3 * p.ch[l + 1]);
    if (p.dgrad_splits[l] > 1) part = std::max(part, (size_t)p.dgrad_splits[l] * d3 * p.ch[l]);
    gmax = std::max(gmax, std::max(d3 * p.ch[l], 8 * d3 * (size_t)std::max(p.ch[l + 1], l == n_conv - 1 ? p.ld_logits : 0)));
  }
  off = 0;
  p.off_partial = take(std::max<size_t>(part, 64));
  p.off_partial_w = take(std::max<size_t>(part_w, 64));        // the weight side runs next to the data side (own partials)
  for (int k = 0; k < 3; ++k) p.off_g[k] = take(gmax);
  p.off_dlogits = take(n_vox * p.ld_logits);
  p.off_colsum = take((size_t)kColSumSlabs * kColSumMaxLd * kColSumMaxChunks);
  p.off_counters = take(kDecMaxConv * kColSumMaxChunks);
  p.scratch_floats = off;
  return 0;
}

static int run_forward_layer(const DecPlan& p, int l, const float* in, const float* w, const float* b, float* out, int ld_out,
                             bool leaky, float* partial, bool x3, cudaStream_t st) {
  DeconvArgs g;
  memset(&g, 0, sizeof(g));
  const int din = p.din[l], d3 = din * din * din;
  g.din = din; g.log2_din = l; g.cin = p.ch[l]; g.cout = p.ch[l + 1]; g.ld_out = ld_out; g.in = in; g.w = w; g.bias = b; g.leaky = leaky ? 1 : 0;
  const int k_total = (din == 1 ? 1 : 8) * g.cin;
  const int tn = tile_n(g.cout);
  const bool big = use_big_tiles(0, din, g.cin, g.cout);
  const int64_t tiles = (int64_t)ceil_div(d3, tile_m_rows(big)) * ceil_div(g.cout, tn);
  const int splits = p.fwd_splits[l];
  g.k_splits = splits;
  g.k_per_split = ceil_div(ceil_div(k_total, splits), tile_k(big)) * tile_k(big);
  g.dst = splits > 1 ? partial : out;
  dim3 grid((unsigned)tiles, 8, (unsigned)splits);
  HN_TRY(launch_deconv_gemm<0>(g, grid, x3, tn, big, st, "hn_vol_decoder_fwd/deconv"));
  if (splits > 1) {
    const int64_t rows = (int64_t)8 * d3;
    deconv_finish_kernel<<<(unsigned)ceil_div<int64_t>(rows * g.cout, 256), 256, 0, st>>>(partial, splits, rows, g.cout, ld_out, b,
                                                                                       leaky ? 1 : 0, nullptr, out);
    HN_TRY(launched("hn_vol_decoder_fwd/finish"));
  }
  return 0;
}

// Library-owned stream and events of the backward's weight side, one set per device (created on first use, never freed)
struct BwdStreams {
  cudaStream_t side = nullptr;
  cudaEvent_t fork[kDecMaxConv], done[kDecMaxConv];
};
static int get_bwd_streams(BwdStreams** out) {
  static BwdStreams per_device[64];
  int dev = 0;
  HN_TRY(check_cuda(cudaGetDevice(&dev), "hn_vol_decoder_bwd/device"));
  if (dev < 0 || dev >= 64) return fail_arg("hn_vol_decoder_bwd: device index %d", dev);
  BwdStreams& b = per_device[dev];
  if (!b.side) {
    HN_TRY(check_cuda(cudaStreamCreateWithFlags(&b.side, cudaStreamNonBlocking), "hn_vol_decoder_bwd/stream"));
    for (int l = 0; l < kDecMaxConv; ++l) {
      HN_TRY(check_cuda(cudaEventCreateWithFlags(&b.fork[l], cudaEventDisableTiming), "hn_vol_decoder_bwd/event"));
      HN_TRY(check_cuda(cudaEventCreateWithFlags(&b.done[l], cudaEventDisableTiming), "hn_vol_decoder_bwd/event"));
    }
  }
  *out = &b;
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
                       float* vol, void* workspace, int64_t workspace_bytes, void* scratch, int64_t scratch_bytes, int precision,
                       void* stream) {
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
                             sc + p.off_partial, precision == HN_PRECISION_FP32, st));
  }
  const int side = 2 * p.din[n_conv - 1], n_vox = side * side * side;
  softmax_prior_fwd_kernel<<<ceil_div(n_vox, 128), 128, 0, st>>>(n_vox, p.ch[n_conv], ws + p.off_logits, prior, vol);
  return launched("hn_vol_decoder_fwd/softmax");
}

int hn_vol_decoder_bwd(int embed, int hidden, int n_conv, const int* channels, const float* embedding, const float* lin_w,
                       const float* const* conv_w, const float* vol, const float* d_vol, const void* workspace,
                       int64_t workspace_bytes, void* scratch, int64_t scratch_bytes, float* d_embedding, float* d_lin_w,
                       float* d_lin_b, float* const* d_conv_w, float* const* d_conv_b, int precision, void* stream) {
  const bool x3 = precision == HN_PRECISION_FP32;
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
  // Two branches per layer, both fed by g_out: the DATA side (data gradient -> next layer's g_out) stays on `stream`, the
  // WEIGHT side (bias sums, weight gradient) runs on a library-owned stream forked and joined with events - capturable
  // (the side stream joins the capture through the fork event and leaves it through the join event).  The three g
  // buffers rotate, and the data gradient that re-uses a buffer waits for the weight side that read it two layers up.
  BwdStreams* bs = nullptr;
  HN_TRY(get_bwd_streams(&bs));
  cudaStream_t sw = bs->side;
  const float* ws = static_cast<const float*>(workspace);
  float* sc = static_cast<float*>(scratch);
  const int side = 2 * p.din[n_conv - 1], n_vox = side * side * side;
  float* g_out = sc + p.off_dlogits;                       // gradient of the pre-activation on the output side of layer l
  int ld_g = p.ld_logits;
  unsigned int* counters = reinterpret_cast<unsigned int*>(sc + p.off_counters);
  HN_TRY(check_cuda(cudaMemsetAsync(counters, 0, kDecMaxConv * kColSumMaxChunks * sizeof(unsigned int), st), "hn_vol_decoder_bwd/memset"));
  softmax_prior_bwd_kernel<<<ceil_div(n_vox, 128), 128, 0, st>>>(n_vox, p.ch[n_conv], vol, d_vol, g_out);
  HN_TRY(launched("hn_vol_decoder_bwd/softmax"));
  for (int l = n_conv - 1; l >= 0; --l) {
    const int din = p.din[l], d3 = din * din * din, cin = p.ch[l], cout = p.ch[l + 1];
    const int64_t rows_out = (int64_t)8 * d3;
    // fork: g_out of this layer is complete on `st`
    HN_TRY(check_cuda(cudaEventRecord(bs->fork[l], st), "hn_vol_decoder_bwd/fork"));
    HN_TRY(check_cuda(cudaStreamWaitEvent(sw, bs->fork[l], 0), "hn_vol_decoder_bwd/fork wait"));
    // bias gradient
    const int slabs = (int)std::max<int64_t>(1, std::min<int64_t>(kColSumSlabs, rows_out / 32));
    column_sum_kernel<<<dim3((unsigned)slabs, (unsigned)ceil_div(ld_g, kColSumMaxLd)), 256, 0, sw>>>(rows_out, ld_g, cout, g_out, sc + p.off_colsum,
                                                                                                  counters + l * kColSumMaxChunks, d_conv_b[l]);
    HN_TRY(launched("hn_vol_decoder_bwd/bias"));
    // weight gradient: one small GEMM per tap, written in the parameter's [cin][64][cout] layout
    DeconvArgs g;
    memset(&g, 0, sizeof(g));
    g.din = din; g.log2_din = l; g.cin = cin; g.cout = cout; g.ld_out = ld_g; g.in = ws + p.off_act[l]; g.outside = g_out; g.dst = d_conv_w[l];
    const bool wbig = use_big_tiles(2, din, cin, cout), dbig = use_big_tiles(1, din, cin, cout);
    const bool outer = din == 1 && (cout & 3) == 0 && (ld_g & 3) == 0;      // first layer: an outer product per tap
    if (outer) {
      const int64_t n4 = (int64_t)cin * 64 * cout / 4;
      deconv_wgrad_din1_kernel<<<(unsigned)ceil_div<int64_t>(n4, 256), 256, 0, sw>>>(cin, cout, ld_g, ws + p.off_act[l], g_out,
                                                                                   reinterpret_cast<float4*>(d_conv_w[l]));
      HN_TRY(launched("hn_vol_decoder_bwd/wgrad outer"));
    }
    const int wsp = p.wgrad_splits[l];
    g.k_splits = wsp;
    g.k_per_split = ceil_div(ceil_div(d3, wsp), tile_k(wbig)) * tile_k(wbig);
    if (wsp > 1) g.dst = sc + p.off_partial_w;
    const int tn = tile_n(cout);
    dim3 wgrid((unsigned)(ceil_div(cin, tile_m_rows(wbig)) * ceil_div(cout, tn)), 64, (unsigned)wsp);
    if (!outer) HN_TRY(launch_deconv_gemm<2>(g, wgrid, x3, tn, wbig, sw, "hn_vol_decoder_bwd/wgrad"));
    if (wsp > 1 && !outer) {
      const int64_t n4 = (int64_t)cin * 64 * cout / 4;
      sum_partials_kernel<<<(unsigned)ceil_div<int64_t>(n4, 256), 256, 0, sw>>>(reinterpret_cast<const float4*>(sc + p.off_partial_w), wsp, n4,
                                                                              reinterpret_cast<float4*>(d_conv_w[l]));
      HN_TRY(launched("hn_vol_decoder_bwd/wgrad sum"));
    }
    HN_TRY(check_cuda(cudaEventRecord(bs->done[l], sw), "hn_vol_decoder_bwd/weight side done"));
    // data gradient -> gradient of the pre-activation of the previous layer (LeakyReLU' from its stored output).  Its
    // output buffer was the g_out of layer l + 2: that layer's weight side must have finished reading it
    float* g_in = sc + p.off_g[l % 3];
    if (l + 2 < n_conv) HN_TRY(check_cuda(cudaStreamWaitEvent(st, bs->done[l + 2], 0), "hn_vol_decoder_bwd/buffer wait"));
    g.w = conv_w[l]; g.act_in = ws + p.off_act[l];
    const int k_total = (din == 1 ? 8 : 64) * ceil_div(cout, tile_k(dbig)) * tile_k(dbig);
    const int64_t tiles = (int64_t)ceil_div(d3, tile_m_rows(dbig)) * ceil_div(cin, kTile);
    const int splits = p.dgrad_splits[l];
    g.k_splits = splits;
    g.k_per_split = ceil_div(ceil_div(k_total, splits), tile_k(dbig)) * tile_k(dbig);
    g.dst = splits > 1 ? sc + p.off_partial : g_in;
    dim3 dgrid((unsigned)tiles, 1, (unsigned)splits);
    HN_TRY(launch_deconv_gemm<1>(g, dgrid, x3, 64, dbig, st, "hn_vol_decoder_bwd/dgrad"));
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
  linear_dgrad_kernel<<<ceil_div(embed, 32), 1024, 0, st>>>(hidden, embed, lin_w, g_out, d_embedding);
  HN_TRY(launched("hn_vol_decoder_bwd/linear dgrad"));
  // join: everything the weight side wrote is complete for whatever follows on `stream`
  for (int l = std::min(n_conv - 1, 1); l >= 0; --l)
    HN_TRY(check_cuda(cudaStreamWaitEvent(st, bs->done[l], 0), "hn_vol_decoder_bwd/join"));
  return 0;
}

}  // extern "C"
