// Canonical field, tensor-core path (HN_PRECISION_BF16): one persistent sm_100a kernel runs
//   hann-PE -> non-rigid MLP -> xyz = x_skel + offset -> PE -> canonical MLP -> raw
// for 128-sample tiles with every activation held on chip:
//   * bf16 operands, fp32 accumulation in TMEM (tcgen05.mma, cta_group::1, M = 128, N = layer width);
//   * the A operand (activations of the tile) lives in shared memory in the UMMA no-swizzle
//     K-major core-matrix layout and is rewritten in place by the epilogue warps
//     (TMEM -> registers -> +bias -> ReLU -> bf16 -> st.shared);
//   * the B operand (weights) is pre-tiled on the device by hn_field_pack_bf16 into exactly the
//     shared-memory image the MMA wants, so the producer streams it with plain 1-D TMA bulk
//     copies (cp.async.bulk + mbarrier complete_tx) through a 4-stage ring - no tensor maps;
//   * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 =
//     epilogue (one TMEM lane quarter each).  Two CTAs are resident per SM (113 KB shared memory,
//     256 TMEM columns each) so one CTA's epilogue overlaps the other's MMAs.
//
// Reference semantics: embedders/hannw_fourier.py:38-49, embedders/fourier.py:9-38,
// non_rigid_motion_mlps/mlp_offset.py:74-114, canonical_mlps/mlp_rgb_sigma.py:132-206.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include <cuda_bf16.h>

#include "common.cuh"

namespace hn {

constexpr int kTileM = 128;
constexpr int kStages = 4;
constexpr uint32_t kStageBytes = 8192;
constexpr uint32_t kActHBytes = kTileM * 256 * 2;      // 64 KB: hidden activations, up to 256 wide
constexpr uint32_t kActPeBytes = kTileM * 64 * 2;      // 16 KB: positional encoding, up to 64 wide
constexpr uint32_t kActBytes = kActHBytes + kActPeBytes;
constexpr uint32_t kLboA = kTileM * 16;                // bytes between K-adjacent core matrices of A
constexpr uint32_t kSmemBar = kActBytes + kStages * kStageBytes;
constexpr uint32_t kSmemTotal = kSmemBar + 256;
constexpr int kTmemCols = 256;
constexpr int kThreads = 192;
constexpr uint32_t kMagic = 0x484e5443u;               // "HNTC"
constexpr int kMaxTcLayers = 2 * HN_MAX_LAYERS;

struct TcChunk {
  uint32_t w_off;     // byte offset of the chunk inside the weight image
  uint32_t bytes;     // multiple of 16
  uint32_t a_off;     // byte offset of the chunk's first A core matrix inside the activation area
  uint16_t n;         // MMA N
  uint8_t ksteps;     // K / 16
  uint8_t flags;      // bit0 first chunk of its layer, bit1 last chunk of its layer
};
static_assert(sizeof(TcChunk) == 16, "TcChunk must be 16 bytes");

struct TcHeader {
  uint32_t magic, n_chunks, table_off, image_off, bias_off, total_bytes;
  uint32_t nr_layers, cn_layers, nr_width, cn_width;
  uint32_t nr_freqs, cn_freqs, nr_pe_pad, cn_pe_pad;
  uint32_t layer_n[kMaxTcLayers];          // MMA N per layer (non-rigid layers first)
  uint32_t layer_bias[kMaxTcLayers];       // float offset of the layer's bias inside the bias block
};

// ------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
// Bounded wait: a protocol bug traps with a message instead of hanging the GPU (no legitimate
// wait in this kernel lasts anywhere near a second).
constexpr long long kWatchdogCycles = 4000000000LL;
__device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity, int who) {
  printf("humannerf_b200 field_tc_kernel: mbarrier wait timed out (block %d thread %d role %d bar 0x%x parity %u)\n",
         (int)blockIdx.x, (int)threadIdx.x, who, bar, parity);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int who = 0) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > kWatchdogCycles) mbar_timeout(bar, parity, who);
  }
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T ; bf16 x bf16 -> fp32, issued by one thread
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// relu(x) rounded to bf16, packed {hi = b, lo = a}
__device__ __forceinline__ uint32_t pack_relu_bf16(float a, float b) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// UMMA shared-memory descriptor, no swizzle, K-major (cute::UMMA::SmemDescriptor):
// [0,14) addr>>4, [16,30) LBO>>4 (K-adjacent core matrices), [32,46) SBO>>4 (8-row groups),
// [46,48) version = 1, [61,64) layout = 0 (SWIZZLE_NONE)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         (1ull << 46);
}
// UMMA instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16, both K-major
__device__ __forceinline__ uint32_t make_idesc(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

// writes `count` (multiple of 8) bf16 values of one row into the A area at column col0 (multiple of 8)
template <int COUNT>
__device__ __forceinline__ void store_row_bf16(uint32_t region, int row, int col0, const float* v) {
#pragma unroll
  for (int c = 0; c < COUNT; c += 8) {
    uint32_t addr = region + (uint32_t)((col0 + c) >> 3) * kLboA + (uint32_t)row * 16u;
    st_shared_v4(addr, pack_bf16(v[c], v[c + 1]), pack_bf16(v[c + 2], v[c + 3]), pack_bf16(v[c + 4], v[c + 5]),
                 pack_bf16(v[c + 6], v[c + 7]));
  }
}

// ------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 2)
field_tc_kernel(const uint8_t* __restrict__ packed, int64_t n_samples, const float* __restrict__ x_skel,
                const float* __restrict__ hann_w, int ignore_non_rigid, float* __restrict__ raw,
                float* __restrict__ xyz_out, float* __restrict__ off_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const TcHeader* hdr = reinterpret_cast<const TcHeader*>(packed);
  const TcChunk* table = reinterpret_cast<const TcChunk*>(packed + hdr->table_off);
  const uint8_t* image = packed + hdr->image_off;
  const float* biases = reinterpret_cast<const float*>(packed + hdr->bias_off);

  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_act_h = s_base;
  const uint32_t s_act_pe = s_base + kActHBytes;
  const uint32_t s_ring = s_base + kActBytes;
  const uint32_t s_full = s_base + kSmemBar;                 // kStages x 8 B
  const uint32_t s_empty = s_full + kStages * 8;             // kStages x 8 B
  const uint32_t s_acc_full = s_empty + kStages * 8;         // 8 B
  const uint32_t s_a_ready = s_acc_full + 8;                 // 8 B
  const uint32_t s_tmem_slot = s_a_ready + 8;                // 4 B
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kSmemBar + (2 * kStages + 2) * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (n_samples + kTileM - 1) / kTileM;
  const int first_layer = ignore_non_rigid ? (int)hdr->nr_layers : 0;
  const int n_layers = (int)(hdr->nr_layers + hdr->cn_layers);
  // chunks of the skipped non-rigid MLP are skipped too
  uint32_t first_chunk = 0;
  if (ignore_non_rigid) {
    uint32_t seen = 0;
    for (uint32_t c = 0; c < hdr->n_chunks && seen < hdr->nr_layers; ++c) {
      if (table[c].flags & 2) ++seen;
      first_chunk = c + 1;
    }
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(s_full + s * 8, 1);
      mbar_init(s_empty + s * 8, 1);
    }
    mbar_init(s_acc_full, 1);
    mbar_init(s_a_ready, 128);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(s_tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (uint32_t c = first_chunk; c < hdr->n_chunks; ++c) {
          const TcChunk ch = table[c];
          mbar_wait(s_empty + stage * 8, phase ^ 1, 1);
          mbar_expect_tx(s_full + stage * 8, ch.bytes);
          tma_bulk_g2s(s_ring + stage * kStageBytes, image + ch.w_off, ch.bytes, s_full + stage * 8);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, a_phase = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (uint32_t c = first_chunk; c < hdr->n_chunks; ++c) {
          const TcChunk ch = table[c];
          if (ch.flags & 1) {
            // the A operand of this layer is complete (and the previous accumulator drained)
            mbar_wait(s_a_ready, a_phase, 2);
            a_phase ^= 1;
            tc_fence_after();
          }
          mbar_wait(s_full + stage * 8, phase, 3);
          tc_fence_after();
          const uint32_t idesc = make_idesc(ch.n);
          const uint32_t b_lbo = (uint32_t)ch.n * 16u;
          const uint32_t b_base = s_ring + stage * kStageBytes;
          const uint32_t a_base = s_base + ch.a_off;
#pragma unroll 1
          for (uint32_t k = 0; k < ch.ksteps; ++k) {
            const uint64_t a_desc = make_desc(a_base + k * 2u * kLboA, kLboA, 128u);
            const uint64_t b_desc = make_desc(b_base + k * 2u * b_lbo, b_lbo, 128u);
            umma_bf16(tmem_base, a_desc, b_desc, idesc, ((ch.flags & 1) && k == 0) ? 0u : 1u);
          }
          umma_commit(s_empty + stage * 8);            // frees the ring slot when these MMAs retire
          if (ch.flags & 2) umma_commit(s_acc_full);   // accumulator of the layer is complete
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;          // row of the tile == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t acc_phase = 0;
    float hw[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) hw[k] = (k < (int)hdr->nr_freqs && hann_w) ? __ldg(hann_w + k) : 0.0f;
    const int nr_freqs = (int)hdr->nr_freqs, cn_freqs = (int)hdr->cn_freqs;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t sample = tile * kTileM + row;
      const bool valid = sample < n_samples;
      float x[3] = {0.f, 0.f, 0.f};
      if (valid) { x[0] = x_skel[sample * 3 + 0]; x[1] = x_skel[sample * 3 + 1]; x[2] = x_skel[sample * 3 + 2]; }
      float q[3] = {x[0], x[1], x[2]};            // position fed to the canonical encoding

      if (!ignore_non_rigid) {
        // hann-windowed encoding, no raw x: per frequency sin(x,y,z) then cos(x,y,z)
        float pe[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) pe[i] = 0.0f;
        float freq = 1.0f;
#pragma unroll
        for (int k = 0; k < 10; ++k) {
          if (k < nr_freqs) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              float s, c;
              __sincosf(x[a] * freq, &s, &c);
              pe[k * 6 + a] = hw[k] * s;
              pe[k * 6 + 3 + a] = hw[k] * c;
            }
          }
          freq *= 2.0f;
        }
        if (hdr->nr_pe_pad <= 16) store_row_bf16<16>(s_act_pe, row, 0, pe);
        else if (hdr->nr_pe_pad <= 32) store_row_bf16<32>(s_act_pe, row, 0, pe);
        else if (hdr->nr_pe_pad <= 48) store_row_bf16<48>(s_act_pe, row, 0, pe);
        else store_row_bf16<64>(s_act_pe, row, 0, pe);
        fence_proxy_async();
        mbar_arrive(s_a_ready);
      }

      for (int layer = first_layer; layer < n_layers; ++layer) {
        const bool is_nr = layer < (int)hdr->nr_layers;
        const bool is_out = (layer == (int)hdr->nr_layers - 1) || (layer == n_layers - 1);
        if (layer == (int)hdr->nr_layers) {
          // canonical encoding of q: [q, sin(2^k q), cos(2^k q)]_k
          float pe[64];
#pragma unroll
          for (int i = 0; i < 64; ++i) pe[i] = 0.0f;
          pe[0] = q[0]; pe[1] = q[1]; pe[2] = q[2];
          float freq = 1.0f;
#pragma unroll
          for (int k = 0; k < 10; ++k) {
            if (k < cn_freqs) {
#pragma unroll
              for (int a = 0; a < 3; ++a) {
                float s, c;
                __sincosf(q[a] * freq, &s, &c);
                pe[3 + k * 6 + a] = s;
                pe[3 + k * 6 + 3 + a] = c;
              }
            }
            freq *= 2.0f;
          }
          if (hdr->cn_pe_pad <= 16) store_row_bf16<16>(s_act_pe, row, 0, pe);
          else if (hdr->cn_pe_pad <= 32) store_row_bf16<32>(s_act_pe, row, 0, pe);
          else if (hdr->cn_pe_pad <= 48) store_row_bf16<48>(s_act_pe, row, 0, pe);
          else store_row_bf16<64>(s_act_pe, row, 0, pe);
          fence_proxy_async();
          mbar_arrive(s_a_ready);
        }

        // wait for the layer's accumulator
        mbar_wait(s_acc_full, acc_phase, 4);
        acc_phase ^= 1;
        tc_fence_after();
        const float* bias = biases + hdr->layer_bias[layer];
        if (!is_out) {
          const int n_cols = (int)hdr->layer_n[layer];
          for (int c0 = 0; c0 < n_cols; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(t_lane + (uint32_t)c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              float f[8];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c0 + g8 * 8 + h * 4));
                f[h * 4 + 0] = __uint_as_float(v[g8 * 8 + h * 4 + 0]) + b.x;
                f[h * 4 + 1] = __uint_as_float(v[g8 * 8 + h * 4 + 1]) + b.y;
                f[h * 4 + 2] = __uint_as_float(v[g8 * 8 + h * 4 + 2]) + b.z;
                f[h * 4 + 3] = __uint_as_float(v[g8 * 8 + h * 4 + 3]) + b.w;
              }
              const uint32_t addr = s_act_h + (uint32_t)((c0 >> 3) + g8) * kLboA + (uint32_t)row * 16u;
              st_shared_v4(addr, pack_relu_bf16(f[0], f[1]), pack_relu_bf16(f[2], f[3]),
                           pack_relu_bf16(f[4], f[5]), pack_relu_bf16(f[6], f[7]));
            }
          }
          tc_fence_before();
          fence_proxy_async();
          mbar_arrive(s_a_ready);
        } else {
          uint32_t v[16];
          tmem_ld16(t_lane, v);
          tmem_ld_wait();
          tc_fence_before();
          if (is_nr) {
            // offsets -> xyz = x_skel + offset (fp32); feeds the canonical encoding
            float o0 = __uint_as_float(v[0]) + __ldg(bias + 0);
            float o1 = __uint_as_float(v[1]) + __ldg(bias + 1);
            float o2 = __uint_as_float(v[2]) + __ldg(bias + 2);
            q[0] = x[0] + o0; q[1] = x[1] + o1; q[2] = x[2] + o2;
            if (valid) {
              off_out[sample * 3 + 0] = o0; off_out[sample * 3 + 1] = o1; off_out[sample * 3 + 2] = o2;
            }
          } else {
            if (valid) {
              float4 r;
              r.x = __uint_as_float(v[0]) + __ldg(bias + 0);
              r.y = __uint_as_float(v[1]) + __ldg(bias + 1);
              r.z = __uint_as_float(v[2]) + __ldg(bias + 2);
              r.w = __uint_as_float(v[3]) + __ldg(bias + 3);
              reinterpret_cast<float4*>(raw)[sample] = r;
              xyz_out[sample * 3 + 0] = q[0]; xyz_out[sample * 3 + 1] = q[1]; xyz_out[sample * 3 + 2] = q[2];
              if (ignore_non_rigid) {
                off_out[sample * 3 + 0] = 0.f; off_out[sample * 3 + 1] = 0.f; off_out[sample * 3 + 2] = 0.f;
              }
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------
// Weight packing
// ------------------------------------------------------------------------------------------
struct PackJob {            // one per chunk, consumed by pack_weights_kernel
  const float* src;         // weight matrix (rows = out features), nullptr = all zero
  int32_t src_ld;
  int32_t src_col0;         // source column of the chunk's k = 0
  int32_t valid_k;          // source columns available from src_col0 (rest is zero padding)
  int32_t valid_n;          // rows of the weight matrix (rest is zero padding)
  int32_t n;                // chunk N
  int32_t kc;               // chunk K
  uint32_t w_off;           // destination byte offset inside the image
};

// one thread per 16-byte core-matrix row: chunk image[(k/8) * (N*16) + n*16 .. +16] = W[n][k0+k .. +8]
__global__ void pack_weights_kernel(const PackJob* __restrict__ jobs, int n_jobs, uint8_t* __restrict__ image) {
  const int j = blockIdx.y;
  if (j >= n_jobs) return;
  const PackJob job = jobs[j];
  const int units = job.n * (job.kc / 8);
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < units; u += gridDim.x * blockDim.x) {
    const int kg = u / job.n, n = u - kg * job.n;
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = kg * 8 + i;
      f[i] = (job.src && n < job.valid_n && k < job.valid_k)
                 ? job.src[(int64_t)n * job.src_ld + job.src_col0 + k] : 0.0f;
    }
    uint4 o;
    o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]); o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(image + job.w_off + (size_t)kg * job.n * 16 + (size_t)n * 16) = o;
  }
}

struct BiasJob {
  const float* bias;        // nullable
  const float* fold_w;      // nullable: + fold_w[n, 0:fold_k] . cond
  int32_t fold_ld, fold_k;
  int32_t valid_n, n;       // rows available / padded N
  uint32_t dst;             // float offset inside the bias block
};

__global__ void pack_bias_kernel(const BiasJob* __restrict__ jobs, int n_jobs, const float* __restrict__ cond,
                                 float* __restrict__ out) {
  const int j = blockIdx.x;
  if (j >= n_jobs) return;
  const BiasJob job = jobs[j];
  for (int n = threadIdx.x; n < job.n; n += blockDim.x) {
    float v = 0.0f;
    if (n < job.valid_n) {
      if (job.bias) v = job.bias[n];
      if (job.fold_w && cond)
        for (int k = 0; k < job.fold_k; ++k) v += job.fold_w[(int64_t)n * job.fold_ld + k] * cond[k];
    }
    out[job.dst + n] = v;
  }
}

static inline int round16(int v) { return (v + 15) & ~15; }

struct TcPlan {
  TcHeader hdr;
  std::vector<TcChunk> chunks;
  std::vector<PackJob> jobs;
  std::vector<BiasJob> biases;
  uint32_t jobs_off = 0, bias_jobs_off = 0;   // scratch tables placed after the bias block
};

static int make_tc_plan(const hn_field_params* p, TcPlan& plan) {
  if (!p) return fail_arg("hn_field(bf16): null params");
  if (p->nr_layers < 2 || p->nr_layers > HN_MAX_LAYERS || p->cn_layers < 2 || p->cn_layers > HN_MAX_LAYERS)
    return fail_arg("hn_field(bf16): layer counts outside [2,%d]", HN_MAX_LAYERS);
  if (p->nr_width % 32 || p->cn_width % 32 || p->nr_width < 32 || p->cn_width < 32 || p->nr_width > 256 ||
      p->cn_width > 256)
    return fail_arg("hn_field(bf16): widths must be multiples of 32 in [32,256]");
  const int nr_pe = 6 * p->nr_freqs, cn_pe = 3 + 6 * p->cn_freqs;
  if (p->nr_freqs < 1 || p->cn_freqs < 1 || p->nr_freqs > 10 || p->cn_freqs > 10 || round16(nr_pe) > 64 ||
      round16(cn_pe) > 64)
    return fail_arg("hn_field(bf16): encodings wider than 64 columns are not supported");
  if (p->nr_skip == 0 || p->nr_skip >= p->nr_layers - 1 || p->cn_skip == 0 || p->cn_skip >= p->cn_layers - 1)
    return fail_arg("hn_field(bf16): skip layer must be a hidden layer (or -1)");
  TcHeader& h = plan.hdr;
  memset(&h, 0, sizeof(h));
  h.magic = kMagic;
  h.nr_layers = p->nr_layers; h.cn_layers = p->cn_layers; h.nr_width = p->nr_width; h.cn_width = p->cn_width;
  h.nr_freqs = p->nr_freqs; h.cn_freqs = p->cn_freqs; h.nr_pe_pad = round16(nr_pe); h.cn_pe_pad = round16(cn_pe);
  uint32_t w_off = 0, bias_floats = 0;
  auto add_segment = [&](const float* W, int ldw, int col0, int valid_k, int k_pad, int valid_n, int n,
                         uint32_t a_region_off, bool first, bool last) {
    // split k_pad (multiple of 16) into chunks of at most kStageBytes
    int kc_max = (int)(kStageBytes / (uint32_t)(n * 2)) / 16 * 16;
    if (kc_max < 16) kc_max = 16;
    for (int k0 = 0; k0 < k_pad; k0 += kc_max) {
      const int kc = std::min(kc_max, k_pad - k0);
      TcChunk c;
      c.w_off = w_off;
      c.bytes = (uint32_t)(n * kc * 2);
      c.a_off = a_region_off + (uint32_t)(k0 / 8) * kLboA;
      c.n = (uint16_t)n;
      c.ksteps = (uint8_t)(kc / 16);
      c.flags = (uint8_t)(((first && k0 == 0) ? 1 : 0) | ((last && k0 + kc >= k_pad) ? 2 : 0));
      plan.chunks.push_back(c);
      PackJob j;
      j.src = W; j.src_ld = ldw; j.src_col0 = col0 + k0; j.valid_k = std::max(0, valid_k - k0);
      j.valid_n = valid_n; j.n = n; j.kc = kc; j.w_off = w_off;
      plan.jobs.push_back(j);
      w_off += c.bytes;
    }
  };
  int layer_idx = 0;
  auto add_mlp = [&](int layers, int width, int skip, int pe, int pe_pad, int lead, bool pe_first,
                     const float* const* W, const float* const* B, int out_dim) {
    for (int l = 0; l < layers; ++l) {
      const bool out = l == layers - 1;
      const int n = out ? 16 : width;
      const int valid_n = out ? out_dim : width;
      const int in_dim = (l == 0) ? lead + pe : (l == skip ? width + pe : width);
      if (l == 0) {
        add_segment(W[l], in_dim, lead, pe, pe_pad, valid_n, n, kActHBytes, true, true);
      } else if (l == skip) {
        if (pe_first) {
          add_segment(W[l], in_dim, 0, pe, pe_pad, valid_n, n, kActHBytes, true, false);
          add_segment(W[l], in_dim, pe, width, width, valid_n, n, 0, false, true);
        } else {
          add_segment(W[l], in_dim, 0, width, width, valid_n, n, 0, true, false);
          add_segment(W[l], in_dim, width, pe, pe_pad, valid_n, n, kActHBytes, false, true);
        }
      } else {
        add_segment(W[l], in_dim, 0, width, width, valid_n, n, 0, true, true);
      }
      BiasJob bj;
      bj.bias = B[l];
      bj.fold_w = (l == 0 && lead > 0) ? W[l] : nullptr;
      bj.fold_ld = in_dim; bj.fold_k = lead; bj.valid_n = valid_n; bj.n = n; bj.dst = bias_floats;
      plan.biases.push_back(bj);
      h.layer_n[layer_idx] = (uint32_t)n;
      h.layer_bias[layer_idx] = bias_floats;
      bias_floats += (uint32_t)n;
      ++layer_idx;
    }
  };
  add_mlp(p->nr_layers, p->nr_width, p->nr_skip, nr_pe, round16(nr_pe), p->cond_dim, false, p->nr_w, p->nr_b, 3);
  add_mlp(p->cn_layers, p->cn_width, p->cn_skip, cn_pe, round16(cn_pe), 0, true, p->cn_w, p->cn_b, 4);
  for (const TcChunk& c : plan.chunks)
    if (c.bytes > kStageBytes || c.bytes % 16) return fail_arg("hn_field(bf16): internal chunk size error");
  h.n_chunks = (uint32_t)plan.chunks.size();
  auto align = [](uint32_t v, uint32_t a) { return (v + a - 1) / a * a; };
  h.table_off = align((uint32_t)sizeof(TcHeader), 16);
  h.image_off = align(h.table_off + h.n_chunks * (uint32_t)sizeof(TcChunk), 128);
  h.bias_off = align(h.image_off + w_off, 128);
  plan.jobs_off = align(h.bias_off + bias_floats * 4u, 128);
  plan.bias_jobs_off = align(plan.jobs_off + (uint32_t)(plan.jobs.size() * sizeof(PackJob)), 128);
  h.total_bytes = align(plan.bias_jobs_off + (uint32_t)(plan.biases.size() * sizeof(BiasJob)), 128);
  return 0;
}

}  // namespace hn

using namespace hn;

extern "C" {

int hn_has_tcgen05(void) { return 1; }

int64_t hn_field_packed_bytes(const hn_field_params* p) {
  TcPlan plan;
  if (make_tc_plan(p, plan) != 0) return -1;
  return (int64_t)plan.hdr.total_bytes;
}

int hn_field_pack_bf16(const hn_field_params* p, const float* cond, void* packed, void* stream) {
  TcPlan plan;
  HN_TRY(make_tc_plan(p, plan));
  if (!packed) return fail_arg("hn_field_pack_bf16: null destination");
  if (p->cond_dim > 0 && !cond) return fail_arg("hn_field_pack_bf16: cond missing");
  for (int l = 0; l < p->nr_layers; ++l)
    if (!p->nr_w[l] || !p->nr_b[l]) return fail_arg("hn_field_pack_bf16: null non-rigid layer %d", l);
  for (int l = 0; l < p->cn_layers; ++l)
    if (!p->cn_w[l] || !p->cn_b[l]) return fail_arg("hn_field_pack_bf16: null canonical layer %d", l);
  cudaStream_t st = as_stream(stream);
  uint8_t* dst = static_cast<uint8_t*>(packed);
  // header, chunk table and the two job tables are staged from pageable host memory: the copies
  // are complete (staged) when cudaMemcpyAsync returns, so the vectors may die with this frame
  HN_TRY(check_cuda(cudaMemcpyAsync(dst, &plan.hdr, sizeof(TcHeader), cudaMemcpyHostToDevice, st), "pack/header"));
  HN_TRY(check_cuda(cudaMemcpyAsync(dst + plan.hdr.table_off, plan.chunks.data(),
                                    plan.chunks.size() * sizeof(TcChunk), cudaMemcpyHostToDevice, st), "pack/table"));
  HN_TRY(check_cuda(cudaMemcpyAsync(dst + plan.jobs_off, plan.jobs.data(), plan.jobs.size() * sizeof(PackJob),
                                    cudaMemcpyHostToDevice, st), "pack/jobs"));
  HN_TRY(check_cuda(cudaMemcpyAsync(dst + plan.bias_jobs_off, plan.biases.data(),
                                    plan.biases.size() * sizeof(BiasJob), cudaMemcpyHostToDevice, st), "pack/bias jobs"));
  dim3 grid(8, (unsigned)plan.jobs.size());
  pack_weights_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const PackJob*>(dst + plan.jobs_off),
                                            (int)plan.jobs.size(), dst + plan.hdr.image_off);
  HN_TRY(launched("hn_field_pack_bf16/weights"));
  pack_bias_kernel<<<(unsigned)plan.biases.size(), 256, 0, st>>>(
      reinterpret_cast<const BiasJob*>(dst + plan.bias_jobs_off), (int)plan.biases.size(), cond,
      reinterpret_cast<float*>(dst + plan.hdr.bias_off));
  return launched("hn_field_pack_bf16/bias");
}

int hn_field_fwd_fp32(const hn_field_params* p, int64_t n_samples, const float* x_skel, const float* cond,
                      const float* hann_w, int ignore_non_rigid, float* raw, float* xyz, float* offsets,
                      void* workspace, int64_t workspace_bytes, int save_for_backward, void* stream);

int hn_field_fwd(const hn_field_params* p, int64_t n_samples, const float* x_skel, const float* cond,
                 const float* hann_w, int ignore_non_rigid, float* raw, float* xyz, float* offsets,
                 void* workspace, int64_t workspace_bytes, int save_for_backward, int precision,
                 const void* packed, void* stream) {
  if (precision == HN_PRECISION_FP32)
    return hn_field_fwd_fp32(p, n_samples, x_skel, cond, hann_w, ignore_non_rigid, raw, xyz, offsets, workspace,
                             workspace_bytes, save_for_backward, stream);
  if (precision != HN_PRECISION_BF16) return fail_arg("hn_field_fwd: unknown precision %d", precision);
  if (save_for_backward)
    return fail_arg("hn_field_fwd: the bf16 tensor-core path is forward-only; train with HN_PRECISION_FP32");
  if (!packed) return fail_arg("hn_field_fwd: bf16 path needs the image from hn_field_pack_bf16");
  if (n_samples < 0 || !x_skel || !raw || !xyz || !offsets) return fail_arg("hn_field_fwd: null pointer");
  if (!ignore_non_rigid && !hann_w) return fail_arg("hn_field_fwd: hann_w missing");
  if (n_samples == 0) return 0;
  TcPlan plan;
  HN_TRY(make_tc_plan(p, plan));
  int dev = 0, sms = 0;
  HN_TRY(check_cuda(cudaGetDevice(&dev), "hn_field_fwd/device"));
  HN_TRY(check_cuda(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "hn_field_fwd/attr"));
  HN_TRY(check_cuda(cudaFuncSetAttribute(field_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kSmemTotal), "hn_field_fwd/smem attr"));
  const int64_t n_tiles = ceil_div<int64_t>(n_samples, kTileM);
  const unsigned grid = (unsigned)std::min<int64_t>(n_tiles, (int64_t)sms * 2);
  field_tc_kernel<<<grid, kThreads, kSmemTotal, as_stream(stream)>>>(
      static_cast<const uint8_t*>(packed), n_samples, x_skel, hann_w, ignore_non_rigid ? 1 : 0, raw, xyz, offsets);
  return launched("hn_field_fwd/tcgen05");
}

}  // extern "C"
