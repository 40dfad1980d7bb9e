// Shared by the TS-form forward kernels (mlp_tc_ts.cu: one tile per CTA, rendering and training; mlp_tc_ts2.cu: two tiles per
// CTA, rendering): the chunk table passed as a kernel parameter, the TS-form MMA / tcgen05.st wrappers and the static
// positional-encoding chunk writers.
#pragma once

#include "mlp_tc_plan.cuh"
#include "mlp_tc_enc.cuh"
#include "mlp_tc_2sm.cuh"

namespace hn {

// Chunk schedule as a KERNEL PARAMETER (constant bank, warp-uniform loads; see IssueTab in mlp_tc_2sm.cuh):
// x = n | flags << 16 | flags2 << 24, y = ss_a_off | ss_ksteps << 8 | ts0_ksteps << 16 | ts1_ksteps << 24, z = w_off, w = bytes
struct TsTab {
  uint4 e[kMaxTsChunks];
};

__device__ __forceinline__ void umma2_ts_bf16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
        "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
        "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// one 8-column chunk kCh of an encoding row (columns as in write_encoding_half of mlp_tc_enc.cuh; the column ->
// (frequency, axis, sin / cos) map is static) -> the shared-memory operand block and, when training, the T16 image row
template <bool kIncludeInput, int kCh>
__device__ __forceinline__ void write_encoding_chunk(uint32_t region, int row, const float x[3], int n_freqs,
                                                     const float* __restrict__ weights, int pad, uint8_t* img_row) {
  constexpr int base = kIncludeInput ? 3 : 0;
  if (kCh * 8 >= pad) return;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = kCh * 8 + i;
    float val = 0.0f;
    if (c < base) {
      val = x[c < 3 ? c : 0];
    } else {
      const int cc = c - base, k = cc / 6, r = cc - 6 * k, a = r >= 3 ? r - 3 : r;
      if (k < 10 && k < n_freqs) {
        const float arg = x[a] * (float)(1 << k);
        const float t = r >= 3 ? __cosf(arg) : __sinf(arg);
        val = weights ? __ldg(weights + k) * t : t;
      }
    }
    v[i] = val;
  }
  if (kCh * 8 + 8 == pad) v[7] = 1.0f;
  store8(region, row, kCh, v, img_row);
}
// the share of warp `part` (0..3) of a row's encoding: chunks part and part + 4
template <bool kIncludeInput>
__device__ __forceinline__ void write_encoding_part(uint32_t region, int row, int part, const float x[3], int n_freqs,
                                                    const float* __restrict__ weights, int pad, uint8_t* img_row) {
  switch (part) {
    case 0: write_encoding_chunk<kIncludeInput, 0>(region, row, x, n_freqs, weights, pad, img_row);
            write_encoding_chunk<kIncludeInput, 4>(region, row, x, n_freqs, weights, pad, img_row); break;
    case 1: write_encoding_chunk<kIncludeInput, 1>(region, row, x, n_freqs, weights, pad, img_row);
            write_encoding_chunk<kIncludeInput, 5>(region, row, x, n_freqs, weights, pad, img_row); break;
    case 2: write_encoding_chunk<kIncludeInput, 2>(region, row, x, n_freqs, weights, pad, img_row);
            write_encoding_chunk<kIncludeInput, 6>(region, row, x, n_freqs, weights, pad, img_row); break;
    default: write_encoding_chunk<kIncludeInput, 3>(region, row, x, n_freqs, weights, pad, img_row);
             write_encoding_chunk<kIncludeInput, 7>(region, row, x, n_freqs, weights, pad, img_row); break;
  }
}


// host: the plan's TS chunks as the kernel-parameter table; *nr_chunks = number of leading chunks of the non-rigid MLP
inline void fill_ts_tab(const TcPlan& plan, TsTab* tab, int* nr_chunks) {
  memset(tab, 0, sizeof(*tab));
  *nr_chunks = 0;
  for (size_t c = 0; c < plan.ts_chunks.size(); ++c) {
    const TsChunk& ch = plan.ts_chunks[c];
    tab->e[c] = make_uint4((uint32_t)ch.n | ((uint32_t)ch.flags << 16) | ((uint32_t)ch.flags2 << 24),
                           (uint32_t)ch.ss_a_off | ((uint32_t)ch.ss_ksteps << 8) | ((uint32_t)ch.ts0_ksteps << 16) |
                               ((uint32_t)ch.ts1_ksteps << 24),
                           ch.w_off, ch.bytes);
    if (ch.flags2 & 2) *nr_chunks = (int)c + 1;
  }
}

}  // namespace hn
