// Device helpers shared by the tcgen05 forward kernels (mlp_tc.cu, mlp_tc_2sm.cu): positional encodings
// written straight into the A operand block, and the T16 image addressing.
// Reference semantics: embedders/hannw_fourier.py:38-49, embedders/fourier.py:9-38.
#pragma once

#include "mlp_tc_plan.cuh"

namespace hn {

// ------------------------------------------------------------------------------------------
// positional encodings written straight into the A operand block (bf16, 8 columns per store);
// `half` selects which of the two threads of a row stores a given 16-byte chunk
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void store8(uint32_t region, int row, int chunk, const float* v, uint8_t* img_row) {
  const uint32_t addr = region + (uint32_t)chunk * kLboA + (uint32_t)row * 16u;
  const uint32_t a = pack_bf16(v[0], v[1]), b = pack_bf16(v[2], v[3]), c = pack_bf16(v[4], v[5]), d = pack_bf16(v[6], v[7]);
  st_shared_v4(addr, a, b, c, d);
  if (img_row) *reinterpret_cast<uint4*>(img_row + (size_t)chunk * 256) = make_uint4(a, b, c, d);
}

// byte offset of (row r of tile t, feature 0) inside a T16 image with F features per sample
__device__ __forceinline__ size_t image_row_offset(int64_t tile, int r, int F) {
  return (size_t)tile * 128 * F * 2 + (size_t)(r >> 4) * 16 * F * 2 + (size_t)(r & 15) * 16;
}

// columns: [optional x (3)], then per frequency sin(x,y,z), cos(x,y,z); zero padding up to `pad`
// columns with a constant 1 in the last one (the bias column).  The two threads of a row (kHalf) store alternate
// 16-byte chunks, and each evaluates only the columns it stores (one MUFU per column; the column -> (frequency,
// axis, sin / cos) map is static after unrolling).
template <bool kIncludeInput, int kHalf>
__device__ __forceinline__ void write_encoding_half(uint32_t region, int row, const float x[3], int n_freqs,
                                                    const float* __restrict__ weights, int pad, uint8_t* img_row) {
  constexpr int base = kIncludeInput ? 3 : 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int ch = 2 * j + kHalf;
    if (ch * 8 < pad) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = ch * 8 + i;
        float val = 0.0f;
        if (c < base) {
          val = x[c < 3 ? c : 0];
        } else {
          const int cc = c - base, k = cc / 6, r = cc - 6 * k, a = r >= 3 ? r - 3 : r;
          if (k < 10 && k < n_freqs) {
            const float arg = x[a] * (float)(1 << k);
            const float t = r >= 3 ? __cosf(arg) : __sinf(arg);
            val = weights ? weights[k] * t : t;
          }
        }
        v[i] = val;
      }
      if (ch * 8 + 8 == pad) v[7] = 1.0f;
      store8(region, row, ch, v, img_row);
    }
  }
}

template <bool kIncludeInput>
__device__ __forceinline__ void write_encoding(uint32_t region, int row, int half, const float x[3], int n_freqs,
                                               const float* __restrict__ weights, int pad, uint8_t* img_row) {
  if (half == 0) write_encoding_half<kIncludeInput, 0>(region, row, x, n_freqs, weights, pad, img_row);
  else write_encoding_half<kIncludeInput, 1>(region, row, x, n_freqs, weights, pad, img_row);
}

// acc masked by a packed pair of bf16 activations (ReLU outputs: positive <=> bit pattern != 0)
__device__ __forceinline__ uint32_t mask_pack(uint32_t h2, uint32_t a_lo, uint32_t a_hi) {
  const float lo = (h2 & 0xFFFFu) ? __uint_as_float(a_lo) : 0.0f;
  const float hi = (h2 >> 16) ? __uint_as_float(a_hi) : 0.0f;
  return pack_bf16(lo, hi);
}


// ReLU-mask words (relu_bits_* in mlp_tc_plan.cuh).  h2 = two non-negative bf16 (cvt.relu: sign bit clear), so
// h + 0x7FFF carries into bit 15 exactly when h != 0 and never into the other half.  After the 16 words of a
// 32-column group have been shifted in, word j sits at bit j (low half = even column) and bit 16 + j (odd column).
__device__ __forceinline__ uint32_t relu_bits_step(uint32_t m, uint32_t h2) {
  return (m >> 1) | ((h2 + 0x7FFF7FFFu) & 0x80008000u);
}
// acc masked by bits j / 16 + j of a mask word
__device__ __forceinline__ uint32_t mask_pack_bits(uint32_t m, int j, uint32_t a_lo, uint32_t a_hi) {
  const float lo = (m & (1u << j)) ? __uint_as_float(a_lo) : 0.0f;
  const float hi = (m & (0x10000u << j)) ? __uint_as_float(a_hi) : 0.0f;
  return pack_bf16(lo, hi);
}

}  // namespace hn
