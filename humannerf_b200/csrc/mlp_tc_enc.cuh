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
// columns with a constant 1 in the last one (the bias column).
template <bool kIncludeInput>
__device__ __forceinline__ void write_encoding(uint32_t region, int row, int half, const float x[3], int n_freqs,
                                               const float* __restrict__ weights, int pad, uint8_t* img_row) {
  float pe[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) pe[i] = 0.0f;
  int base = 0;
  if (kIncludeInput) { pe[0] = x[0]; pe[1] = x[1]; pe[2] = x[2]; base = 3; }
  float freq = 1.0f;
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    if (k < n_freqs) {
      const float w = weights ? weights[k] : 1.0f;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        float s, c;
        __sincosf(x[a] * freq, &s, &c);
        pe[base + k * 6 + a] = w * s;
        pe[base + k * 6 + 3 + a] = w * c;
      }
    }
    freq *= 2.0f;
  }
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    if (ch * 8 < pad && (ch & 1) == half) {
      if (ch * 8 + 8 == pad) pe[ch * 8 + 7] = 1.0f;
      store8(region, row, ch, pe + ch * 8, img_row);
    }
  }
}

}  // namespace hn
