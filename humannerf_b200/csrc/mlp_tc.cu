// Canonical field, tensor-core path (HN_PRECISION_BF16): one persistent sm_100a kernel runs
//   hann-PE -> non-rigid MLP -> xyz = x_skel + offset -> PE -> canonical MLP -> raw
// with every activation held on chip.
//
//   * one CTA per SM works on PAIRS of 128-sample tiles (256 samples): both tiles consume each
//     weight chunk back to back, so every byte fetched from L2 feeds 256 rows of MMA;
//   * bf16 operands, fp32 accumulation in TMEM (tcgen05.mma cta_group::1, M = 128, N = layer width;
//     tile X accumulates in TMEM columns [0,256), tile Y in [256,512));
//   * the A operand (activations) lives in shared memory in the UMMA no-swizzle K-major
//     core-matrix layout and is rewritten in place by 16 epilogue warps
//     (TMEM -> registers -> ReLU -> bf16 -> st.shared);
//   * biases ride in the GEMM: the positional-encoding block of A carries a constant-1 column and
//     the weight image carries the bias in the matching K row, so the epilogue has no loads;
//   * the B operand (weights) is pre-tiled by hn_field_pack_bf16 into exactly the shared-memory
//     image the MMA wants and streamed with 1-D TMA bulk copies (cp.async.bulk + mbarrier
//     complete_tx) through a 4 x 16 KB ring - no tensor maps;
//   * warp roles: warp 0 = TMA producer, warps 1 and 18 = MMA issuers (one per tile: a single thread
//     issues ~1 MMA per 130 cycles, slower than the 128-cycle MMA itself), warps 2..17 = epilogue
//     (two warps per TMEM lane quarter and tile, each owning half of the columns);
//   * the two tiles have their own accumulator-full / operand-ready barriers, so they drift apart by
//     up to the depth of the weight ring and one tile's epilogue overlaps the other tile's MMAs.
//
// Reference semantics: embedders/hannw_fourier.py:38-49, embedders/fourier.py:9-38,
// non_rigid_motion_mlps/mlp_offset.py:74-114, canonical_mlps/mlp_rgb_sigma.py:132-206.
#include <cstdlib>

#include "mlp_tc_plan.cuh"
#include "mlp_tc_enc.cuh"

namespace hn {

// ------------------------------------------------------------------------------------------
// Cycle accounting of block 0 (measurement aid, active only with HN_TC_ABLATE & 8):
//  [0..3]  producer: total, wait-empty
//  [4..9]  issuer tile 0: total, wait a_ready, wait full, issue ; [10..15] issuer tile 1
//  [16..21] epilogue thread (tile 0, quarter 0, half 0, lane 0): total, wait acc, encoding, hidden, out
// ------------------------------------------------------------------------------------------
// Compiled in only with -DHN_TC_PROFILE (HN_TC_PROFILE=1 ./build.sh): the counters cost registers.
__device__ unsigned long long g_tc_prof[32];
#ifdef HN_TC_PROFILE
// counters accumulate in registers (pacc[4] per role) and are flushed once when the role finishes
#define HN_PROF_T(var) const long long var = prof ? clock64() : 0
#define HN_PROF_ADD(slot, t0) do { if (prof) pacc[slot] += (unsigned long long)(clock64() - (t0)); } while (0)
#define HN_PROF_FLUSH(base) do { if (prof) { for (int _i = 0; _i < 4; ++_i) g_tc_prof[(base) + _i] += pacc[_i]; } } while (0)
#define HN_PROF_DECL unsigned long long pacc[4] = {0, 0, 0, 0}
#else
#define HN_PROF_T(var) do { } while (0)
#define HN_PROF_ADD(slot, t0) do { } while (0)
#define HN_PROF_FLUSH(base) do { } while (0)
#define HN_PROF_DECL do { } while (0)
#endif

// ------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------
template <bool kStore>
__global__ void __launch_bounds__(kThreads, 1)
field_tc_kernel(const uint8_t* __restrict__ packed, int64_t n_cap, const int64_t* __restrict__ n_dev,
                const float* __restrict__ x_skel,
                const float* __restrict__ hann_w, int ignore_non_rigid, float* __restrict__ raw,
                float* __restrict__ xyz_out, float* __restrict__ off_out, const TrainBufs bufs, int ablate) {
  // n_dev (nullable): the number of samples actually present, read from device memory (empty-space skipping
  // compacts the samples without telling the host); n_cap bounds it and sizes the buffers
  const int64_t n_samples = n_dev ? min(max(__ldg(n_dev), (int64_t)0), n_cap) : n_cap;
  extern __shared__ __align__(1024) uint8_t smem[];
  const TcHeader* hdr = reinterpret_cast<const TcHeader*>(packed);
  const uint8_t* image = packed + hdr->image_off;

  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_act_pe = s_base + kActHBytes;
  const uint32_t s_ring = s_base + kRingOff;
  const uint32_t s_full = s_base + kBarOff;                  // kStages x 8 B
  const uint32_t s_empty = s_full + kStages * 8;             // kStages x 8 B
  const uint32_t s_acc_full = s_empty + kStages * 8;         // kTiles x 8 B
  const uint32_t s_a_ready = s_acc_full + kTiles * 8;        // kTiles x 8 B
  const uint32_t s_tmem_slot = s_a_ready + kTiles * 8;       // 4 B
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(smem + kBarOff + (2 * kStages + 2 * kTiles) * 8);
  const TcChunk* table = reinterpret_cast<const TcChunk*>(smem + kTableOff);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef HN_TC_PROFILE
  const bool prof = (ablate & 8) && blockIdx.x == 0;
#endif
  const int64_t n_pairs = (n_samples + kTiles * kTileM - 1) / (kTiles * kTileM);
  const int nr_layers = (int)hdr->nr_layers;
  const int n_layers = nr_layers + (int)hdr->cn_layers;
  const int first_layer = ignore_non_rigid ? nr_layers : 0;
  const uint32_t n_chunks = hdr->n_chunks;

  // chunk table -> shared memory (every role reads it once per chunk)
  {
    const uint4* src = reinterpret_cast<const uint4*>(packed + hdr->table_off);
    uint4* dst = reinterpret_cast<uint4*>(smem + kTableOff);
    for (uint32_t i = threadIdx.x; i < n_chunks; i += blockDim.x) dst[i] = src[i];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(s_full + s * 8, 1);
      mbar_init(s_empty + s * 8, kTiles);          // one tcgen05.commit per MMA issuer
    }
    for (int t = 0; t < kTiles; ++t) {
      mbar_init(s_acc_full + t * 8, 1);
      mbar_init(s_a_ready + t * 8, kEpiWarps / kTiles);   // one arrive per epilogue warp of the tile
    }
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

  // chunks of the skipped non-rigid MLP are skipped too
  uint32_t first_chunk = 0;
  if (ignore_non_rigid) {
    int seen = 0;
    for (uint32_t c = 0; c < n_chunks && seen < nr_layers; ++c) {
      if (table[c].flags & 2) ++seen;
      first_chunk = c + 1;
    }
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      HN_PROF_DECL;
      HN_PROF_T(p_t0);
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        for (uint32_t c = first_chunk; c < n_chunks; ++c) {
          const TcChunk ch = table[c];
          HN_PROF_T(p_w);
          mbar_wait(s_empty + stage * 8, phase ^ 1, 1);
          HN_PROF_ADD(1, p_w);
          mbar_expect_tx(s_full + stage * 8, ch.bytes);
          tma_bulk_g2s(s_ring + stage * kStageBytes, image + ch.w_off, ch.bytes, s_full + stage * 8);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      HN_PROF_ADD(0, p_t0);
      HN_PROF_FLUSH(0);
    }
  } else if (warp == kMmaWarp0 || warp == kMmaWarp1) {
    // ===================== MMA issuers: warp 1 drives tile 0, warp 18 drives tile 1 =====================
    if (lane == 0) {
      const uint32_t t = (warp == kMmaWarp0) ? 0u : 1u;
      const uint32_t d_tmem = tmem_base + t * 256u;
      const uint32_t bar_ready = s_a_ready + t * 8, bar_acc = s_acc_full + t * 8;
      // descriptor halves: hi = SBO (128 B >> 4) | version 1 (bit 46 -> bit 14 of hi); lo = addr >> 4 | LBO >> 4 << 16
      const uint32_t desc_hi = (128u >> 4) | (1u << 14);
      const uint32_t a_lo_const = ((kLboA >> 4) << 16);
      uint32_t stage = 0, phase = 0, a_phase = 0;
      const int pbase = 4 + 6 * (int)t;
      HN_PROF_DECL;
      HN_PROF_T(m_t0);
      // A dependent shared-memory round trip of this thread costs ~120 clk while MMAs are in flight
      // (scripts/micro/umma_rate.cu): the NEXT chunk's table entry is fetched before this chunk's waits.
      const uint32_t s_table = smem_u32(table);
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        TcChunk ch_next = table[first_chunk];
        for (uint32_t c = first_chunk; c < n_chunks; ++c) {
          const TcChunk ch = ch_next;
          {
            const uint32_t cn = c + 1 < n_chunks ? c + 1 : c;
            uint32_t w0, w1, w2, w3;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(s_table + cn * 16u));
            ch_next.w_off = w0; ch_next.bytes = w1; ch_next.a_off = w2;
            ch_next.n = (uint16_t)(w3 & 0xFFFFu); ch_next.ksteps = (uint8_t)((w3 >> 16) & 0xFFu); ch_next.flags = (uint8_t)(w3 >> 24);
          }
          if (ch.flags & 1) {
            // the A operand of this layer is complete (and the previous accumulator drained)
            HN_PROF_T(m_w);
            mbar_wait(bar_ready, a_phase, 2);
            HN_PROF_ADD(1, m_w);
            a_phase ^= 1;
            tc_fence_after();
          }
          HN_PROF_T(m_f);
          mbar_wait(s_full + stage * 8, phase, 3);
          HN_PROF_ADD(2, m_f);
          tc_fence_after();
          HN_PROF_T(m_i);
          const uint32_t idesc = make_idesc(ch.n);
          const bool in_pe = (ch.flags & 4) != 0;
          const uint32_t a_addr = (in_pe ? s_act_pe + t * kActPeTile : s_base + t * kActHTile) + ch.a_off;
          const uint32_t b_addr = s_ring + stage * kStageBytes;
          uint32_t a_lo = ((a_addr & 0x3FFFFu) >> 4) | a_lo_const;
          uint32_t b_lo = ((b_addr & 0x3FFFFu) >> 4) | ((uint32_t)ch.n << 16);      // LBO = n*16 bytes -> n
          const uint32_t a_inc = (2u * kLboA) >> 4, b_inc = 2u * (uint32_t)ch.n;     // one K step = two core matrices
          uint32_t acc = (ch.flags & 1) ? 0u : 1u;
#pragma unroll 2
          for (uint32_t k = 0; k < ch.ksteps; ++k) {
            umma_bf16(d_tmem, ((uint64_t)desc_hi << 32) | a_lo, ((uint64_t)desc_hi << 32) | b_lo, idesc, acc);
            acc = 1u;
            a_lo += a_inc;
            b_lo += b_inc;
          }
          umma_commit(s_empty + stage * 8);            // ring slot is free once both issuers' MMAs retire
          if (ch.flags & 2) umma_commit(bar_acc);      // accumulator of the layer is complete
          HN_PROF_ADD(3, m_i);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      HN_PROF_ADD(0, m_t0);
      HN_PROF_FLUSH(pbase);
    }
  } else {
    // ===================== epilogue warps (2 .. 17) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int tile = ew >> 3;                     // 0: rows 0..127 of the pair, 1: rows 128..255
    const int half = (ew >> 2) & 1;               // which half of the layer's columns / PE chunks
    const int row = quarter * 32 + lane;          // row inside the tile == TMEM lane
    const uint32_t t_acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)tile * 256u;
    const uint32_t a_h = s_base + (uint32_t)tile * kActHTile;
    const uint32_t a_pe = s_act_pe + (uint32_t)tile * kActPeTile;
    const uint32_t bar_ready = s_a_ready + (uint32_t)tile * 8, bar_acc = s_acc_full + (uint32_t)tile * 8;
    uint32_t acc_phase = 0;
    float hw[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) hw[k] = (k < (int)hdr->nr_freqs && hann_w) ? __ldg(hann_w + k) : 0.0f;
    const int nr_freqs = (int)hdr->nr_freqs, cn_freqs = (int)hdr->cn_freqs;
    const int nr_pad = (int)hdr->nr_pe_pad, cn_pad = (int)hdr->cn_pe_pad;
#ifdef HN_TC_PROFILE
    const bool eprof = prof && warp == 4 && lane == 0;      // warp 4: tile 0, quarter 0, half 0
#define HN_EPROF_T(var) const long long var = eprof ? clock64() : 0
#define HN_EPROF_ADD(slot, t0) do { if (eprof) pacc[(slot) - 16] += (unsigned long long)(clock64() - (t0)); } while (0)
    unsigned long long pacc[4] = {0, 0, 0, 0};
#else
#define HN_EPROF_T(var) do { } while (0)
#define HN_EPROF_ADD(slot, t0) do { } while (0)
#endif
    HN_EPROF_T(e_t0);

    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const int64_t tile_idx = pair * kTiles + tile;
      const int64_t sample = tile_idx * kTileM + row;
      const bool valid = sample < n_samples;
      const bool store = kStore && tile_idx < bufs.n_tiles;      // the training images hold ceil(P/128) tiles
      float x[3] = {0.f, 0.f, 0.f};
      if (valid) { x[0] = x_skel[sample * 3 + 0]; x[1] = x_skel[sample * 3 + 1]; x[2] = x_skel[sample * 3 + 2]; }
      float q[3] = {x[0], x[1], x[2]};            // position fed to the canonical encoding

      if (!ignore_non_rigid) {
        HN_EPROF_T(e_pe);
        write_encoding<false>(a_pe, row, half, x, nr_freqs, hw, nr_pad,
                              store ? bufs.pe[0] + image_row_offset(tile_idx, row, nr_pad) : nullptr);
        fence_proxy_async();
        warp_arrive(bar_ready, lane);
        HN_EPROF_ADD(18, e_pe);
      }

      for (int layer = first_layer; layer < n_layers; ++layer) {
        const bool is_nr = layer < nr_layers;
        const bool is_out = (layer == nr_layers - 1) || (layer == n_layers - 1);
        if (layer == nr_layers) {
          HN_EPROF_T(e_pe);
          write_encoding<true>(a_pe, row, half, q, cn_freqs, nullptr, cn_pad,
                               store ? bufs.pe[1] + image_row_offset(tile_idx, row, cn_pad) : nullptr);
          fence_proxy_async();
          warp_arrive(bar_ready, lane);
          HN_EPROF_ADD(18, e_pe);
        }
        HN_EPROF_T(e_w);
        mbar_wait(bar_acc, acc_phase, 4);
        HN_EPROF_ADD(17, e_w);
        acc_phase ^= 1;
        tc_fence_after();
        HN_EPROF_T(e_h);
        if (!is_out && (ablate & 1)) {
          // measurement aid (HN_TC_ABLATE=1): barrier handshake only, accumulator never read
          tc_fence_before();
          fence_proxy_async();
          warp_arrive(bar_ready, lane);
        } else if (!is_out) {
          const int n_full = (int)hdr->layer_n[layer];
          const int n_half = n_full >> 1;                          // columns owned by this thread
          const int c_begin = half * n_half;
          uint8_t* h_img = nullptr;
          if (store)
            h_img = (is_nr ? bufs.h[0][layer] : bufs.h[1][layer - nr_layers]) + image_row_offset(tile_idx, row, n_full);
          uint32_t v[2][32];
          tmem_ld32(t_acc + (uint32_t)c_begin, v[0]);
#pragma unroll 1
          for (int c0 = 0; c0 < n_half; c0 += 64) {
            // two 32-column groups per iteration so that the register buffers are indexed statically
            tmem_ld_wait();
            if (c0 + 32 < n_half) tmem_ld32(t_acc + (uint32_t)(c_begin + c0 + 32), v[1]);
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              const uint32_t* u = v[0] + g8 * 8;
              const int chunk = ((c_begin + c0) >> 3) + g8;
              const uint4 pk = make_uint4(pack_relu_bf16(__uint_as_float(u[0]), __uint_as_float(u[1])),
                                          pack_relu_bf16(__uint_as_float(u[2]), __uint_as_float(u[3])),
                                          pack_relu_bf16(__uint_as_float(u[4]), __uint_as_float(u[5])),
                                          pack_relu_bf16(__uint_as_float(u[6]), __uint_as_float(u[7])));
              st_shared_v4(a_h + (uint32_t)chunk * kLboA + (uint32_t)row * 16u, pk.x, pk.y, pk.z, pk.w);
              if (store) *reinterpret_cast<uint4*>(h_img + (size_t)chunk * 256) = pk;
            }
            if (c0 + 32 < n_half) {
              tmem_ld_wait();
              if (c0 + 64 < n_half) tmem_ld32(t_acc + (uint32_t)(c_begin + c0 + 64), v[0]);
#pragma unroll
              for (int g8 = 0; g8 < 4; ++g8) {
                const uint32_t* u = v[1] + g8 * 8;
                const int chunk = ((c_begin + c0 + 32) >> 3) + g8;
                const uint4 pk = make_uint4(pack_relu_bf16(__uint_as_float(u[0]), __uint_as_float(u[1])),
                                            pack_relu_bf16(__uint_as_float(u[2]), __uint_as_float(u[3])),
                                            pack_relu_bf16(__uint_as_float(u[4]), __uint_as_float(u[5])),
                                            pack_relu_bf16(__uint_as_float(u[6]), __uint_as_float(u[7])));
                st_shared_v4(a_h + (uint32_t)chunk * kLboA + (uint32_t)row * 16u, pk.x, pk.y, pk.z, pk.w);
                if (store) *reinterpret_cast<uint4*>(h_img + (size_t)chunk * 256) = pk;
              }
            }
          }
          tc_fence_before();
          fence_proxy_async();
          warp_arrive(bar_ready, lane);
          HN_EPROF_ADD(19, e_h);
        } else {
          uint32_t v[16];
          tmem_ld16(t_acc, v);
          tmem_ld_wait();
          tc_fence_before();
          if (is_nr) {
            // offsets -> xyz = x_skel + offset (fp32); feeds the canonical encoding
            const float o0 = __uint_as_float(v[0]), o1 = __uint_as_float(v[1]), o2 = __uint_as_float(v[2]);
            q[0] = x[0] + o0; q[1] = x[1] + o1; q[2] = x[2] + o2;
            if (valid && half == 0) {
              off_out[sample * 3 + 0] = o0; off_out[sample * 3 + 1] = o1; off_out[sample * 3 + 2] = o2;
            }
          } else if (valid && half == 0) {
            float4 r;
            r.x = __uint_as_float(v[0]); r.y = __uint_as_float(v[1]);
            r.z = __uint_as_float(v[2]); r.w = __uint_as_float(v[3]);
            reinterpret_cast<float4*>(raw)[sample] = r;
            xyz_out[sample * 3 + 0] = q[0]; xyz_out[sample * 3 + 1] = q[1]; xyz_out[sample * 3 + 2] = q[2];
            if (ignore_non_rigid) {
              off_out[sample * 3 + 0] = 0.f; off_out[sample * 3 + 1] = 0.f; off_out[sample * 3 + 2] = 0.f;
            }
          }
        }
      }
    }
    HN_EPROF_ADD(16, e_t0);
#ifdef HN_TC_PROFILE
    if (eprof) { for (int i = 0; i < 4; ++i) g_tc_prof[16 + i] += pacc[i]; }
#endif
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
// one thread per 16-byte core-matrix row: image[(k/8) * (N*16) + n*16 .. +16] = W[n][k0+k .. +8]
__global__ void pack_weights_kernel(const PackJob* __restrict__ jobs, int n_jobs, const float* __restrict__ cond,
                                    uint8_t* __restrict__ image) {
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
      float v = 0.0f;
      if (n < job.valid_n) {
        if (job.src && k < job.valid_k)
          v = job.transpose ? job.src[(int64_t)(job.src_row0 + k) * job.src_ld + job.src_col0 + n]
                            : job.src[(int64_t)n * job.src_ld + job.src_col0 + k];
        if (k == job.bias_row) {
          v = job.bias ? job.bias[n] : 0.0f;
          if (job.fold_w && cond)
            for (int c = 0; c < job.fold_k; ++c) v += job.fold_w[(int64_t)n * job.fold_ld + c] * cond[c];
        }
      }
      f[i] = v;
    }
    uint4 o;
    o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]); o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
    size_t dst = (size_t)kg * job.n * 16 + (size_t)n * 16;
    if (job.split) {
      const int nh = job.n >> 1, r = n >= nh ? 1 : 0;
      dst = (size_t)r * (job.kc / 8) * nh * 16 + (size_t)kg * nh * 16 + (size_t)(n - r * nh) * 16;
    }
    *reinterpret_cast<uint4*>(image + job.w_off + dst) = o;
  }
}

// mlp_tc_cl.cu
int launch_field_tcc(const TcPlan& plan, const uint8_t* packed, int64_t n_samples, const float* x_skel,
                     const float* hann_w, int ignore_non_rigid, float* raw, float* xyz, float* offsets,
                     const TrainBufs* bufs, int sms, cudaStream_t st);

// mlp_tc_2sm.cu
int launch_field_tc2(const TcPlan& plan, const uint8_t* packed, int64_t n_samples, const int64_t* n_samples_dev, const float* x_skel,
                     const float* hann_w, int ignore_non_rigid, float* raw, float* xyz, float* offsets,
                     const TrainBufs* bufs, int sms, int ablate, cudaStream_t st);

int64_t tc_train_workspace_bytes(const hn_field_params* p, int64_t n_samples) {
  TcPlan plan;
  if (make_tc_plan(p, plan) != 0) return -1;
  return carve_train_bufs(plan, n_samples, nullptr, nullptr) + 256;
}

}  // namespace hn

using namespace hn;

extern "C" {

int hn_has_tcgen05(void) { return 1; }

/* measurement aid: copies (and clears) the 32 cycle counters of block 0 (HN_TC_ABLATE & 8) */
int hn_debug_tc_prof(unsigned long long* out_host) {
  if (!out_host) return fail_arg("hn_debug_tc_prof: null");
  HN_TRY(check_cuda(cudaMemcpyFromSymbol(out_host, g_tc_prof, sizeof(unsigned long long) * 32), "hn_debug_tc_prof/read"));
  unsigned long long zero[32] = {0};
  return check_cuda(cudaMemcpyToSymbol(g_tc_prof, zero, sizeof(zero)), "hn_debug_tc_prof/clear");
}

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
  // header, chunk table and job table are staged from pageable host memory: the copies are complete
  // (staged) when cudaMemcpyAsync returns, so the vectors may die with this frame
  HN_TRY(check_cuda(cudaMemcpyAsync(dst, &plan.hdr, sizeof(TcHeader), cudaMemcpyHostToDevice, st), "pack/header"));
  HN_TRY(check_cuda(cudaMemcpyAsync(dst + plan.hdr.table_off, plan.chunks.data(),
                                    plan.chunks.size() * sizeof(TcChunk), cudaMemcpyHostToDevice, st), "pack/table"));
  HN_TRY(check_cuda(cudaMemcpyAsync(dst + plan.hdr.bwd_table_off, plan.bwd_chunks.data(),
                                    plan.bwd_chunks.size() * sizeof(TcChunk), cudaMemcpyHostToDevice, st), "pack/bwd table"));
  HN_TRY(check_cuda(cudaMemcpyAsync(dst + plan.jobs_off, plan.jobs.data(), plan.jobs.size() * sizeof(PackJob),
                                    cudaMemcpyHostToDevice, st), "pack/jobs"));
  dim3 grid(8, (unsigned)plan.jobs.size());
  pack_weights_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const PackJob*>(dst + plan.jobs_off),
                                            (int)plan.jobs.size(), cond, dst + plan.hdr.image_off);
  return launched("hn_field_pack_bf16");
}

int hn_field_repack_bf16(const hn_field_params* p, const float* cond, void* packed, void* stream) {
  TcPlan plan;
  HN_TRY(make_tc_plan(p, plan));
  if (!packed) return fail_arg("hn_field_repack_bf16: null image");
  if (p->cond_dim > 0 && !cond) return fail_arg("hn_field_repack_bf16: cond missing");
  uint8_t* dst = static_cast<uint8_t*>(packed);
  dim3 grid(8, (unsigned)plan.jobs.size());
  pack_weights_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const PackJob*>(dst + plan.jobs_off),
                                                           (int)plan.jobs.size(), cond, dst + plan.hdr.image_off);
  return launched("hn_field_repack_bf16");
}

int hn_field_fwd_fp32(const hn_field_params* p, int64_t n_samples, const float* x_skel, const float* cond,
                      const float* hann_w, int ignore_non_rigid, float* raw, float* xyz, float* offsets,
                      void* workspace, int64_t workspace_bytes, int save_for_backward, void* stream);

int hn_field_fwd(const hn_field_params* p, int64_t n_samples, const float* x_skel, const float* cond,
                 const float* hann_w, int ignore_non_rigid, float* raw, float* xyz, float* offsets,
                 void* workspace, int64_t workspace_bytes, int save_for_backward, int precision,
                 const void* packed, const int64_t* n_samples_dev, void* stream) {
  if (precision == HN_PRECISION_FP32 && n_samples_dev)
    return fail_arg("hn_field_fwd: a device-side sample count needs HN_PRECISION_BF16");
  if (precision == HN_PRECISION_FP32)
    return hn_field_fwd_fp32(p, n_samples, x_skel, cond, hann_w, ignore_non_rigid, raw, xyz, offsets, workspace,
                             workspace_bytes, save_for_backward, stream);
  if (precision != HN_PRECISION_BF16) return fail_arg("hn_field_fwd: unknown precision %d", precision);
  if (!packed) return fail_arg("hn_field_fwd: bf16 path needs the image from hn_field_pack_bf16");
  if (n_samples == 0) return 0;
  if (n_samples < 0 || !x_skel || !raw || !xyz || !offsets) return fail_arg("hn_field_fwd: null pointer");
  if (!ignore_non_rigid && !hann_w) return fail_arg("hn_field_fwd: hann_w missing");
  TcPlan plan;
  HN_TRY(make_tc_plan(p, plan));
  int dev = 0, sms = 0;
  HN_TRY(check_cuda(cudaGetDevice(&dev), "hn_field_fwd/device"));
  HN_TRY(check_cuda(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "hn_field_fwd/attr"));
  const int64_t n_pairs = ceil_div<int64_t>(n_samples, kTiles * kTileM);
  const unsigned grid = (unsigned)std::min<int64_t>(n_pairs, (int64_t)sms);
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  // measurement aid, never set in production: HN_TC_ABLATE=1 skips the hidden-layer epilogue work
  // (results are garbage) to expose how much of the kernel time the epilogue accounts for
  const char* abl = getenv("HN_TC_ABLATE");
  const int ablate = abl ? atoi(abl) : 0;
  // default: the cta_group::2 kernel of mlp_tc_2sm.cu (CTA pairs share every weight chunk)
  if (use_2sm_kernel()) {
    if (sms < 2) return fail_arg("hn_field_fwd: the tcgen05 forward kernel needs a device with thread-block clusters");
    TrainBufs bufs;
    if (save_for_backward) {
      const int64_t need = carve_train_bufs(plan, n_samples, static_cast<uint8_t*>(workspace), &bufs);
      if (!workspace || workspace_bytes < need) return fail_arg("hn_field_fwd: bf16 training workspace too small");
      if (!use_relu_bits()) memset(bufs.bits, 0, sizeof(bufs.bits));
    }
    return launch_field_tc2(plan, pk, n_samples, n_samples_dev, x_skel, hann_w, ignore_non_rigid ? 1 : 0, raw, xyz, offsets,
                            save_for_backward ? &bufs : nullptr, sms, ablate, as_stream(stream));
  }
  // HN_TC_KERNEL=pair: the tile-pair kernel below; HN_TC_KERNEL=cluster: the experimental cluster kernel
  // (mlp_tc_cl.cu: correct, but slower - its single tile per CTA exposes the ~2000-clk accumulator ->
  // epilogue -> operand hand-over latency of every layer)
  if (use_cluster_kernel() && sms >= 2 && !n_samples_dev) {
    TrainBufs bufs;
    if (save_for_backward) {
      const int64_t need = carve_train_bufs(plan, n_samples, static_cast<uint8_t*>(workspace), &bufs);
      if (!workspace || workspace_bytes < need) return fail_arg("hn_field_fwd: bf16 training workspace too small");
    }
    return launch_field_tcc(plan, pk, n_samples, x_skel, hann_w, ignore_non_rigid ? 1 : 0, raw, xyz, offsets,
                            save_for_backward ? &bufs : nullptr, sms, as_stream(stream));
  }
  if (save_for_backward) {
    TrainBufs bufs;
    const int64_t need = carve_train_bufs(plan, n_samples, static_cast<uint8_t*>(workspace), &bufs);
    if (!workspace || workspace_bytes < need) return fail_arg("hn_field_fwd: bf16 training workspace too small");
    HN_TRY(check_cuda(cudaFuncSetAttribute(field_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)kSmemTotal), "hn_field_fwd/smem attr"));
    field_tc_kernel<true><<<grid, kThreads, kSmemTotal, as_stream(stream)>>>(
        pk, n_samples, n_samples_dev, x_skel, hann_w, ignore_non_rigid ? 1 : 0, raw, xyz, offsets, bufs, ablate);
  } else {
    TrainBufs bufs;
    memset(&bufs, 0, sizeof(bufs));
    HN_TRY(check_cuda(cudaFuncSetAttribute(field_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)kSmemTotal), "hn_field_fwd/smem attr"));
    field_tc_kernel<false><<<grid, kThreads, kSmemTotal, as_stream(stream)>>>(
        pk, n_samples, n_samples_dev, x_skel, hann_w, ignore_non_rigid ? 1 : 0, raw, xyz, offsets, bufs, ablate);
  }
  return launched("hn_field_fwd/tcgen05");
}

}  // extern "C"
