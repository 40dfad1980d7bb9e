// Canonical field, tensor-core forward with the hidden activations in TENSOR MEMORY ("TS" form of tcgen05.mma):
//
//   hann-PE -> non-rigid MLP -> xyz = x_skel + offset -> PE -> canonical MLP -> raw      (one persistent kernel)
//
// The SS-form kernel (mlp_tc_2sm.cu) keeps the activations of two tile sets in shared memory and is bound by the
// shared-memory bandwidth: per 256-wide layer and SM the tensor core reads 128 KB of A and 128 KB of B, the epilogue
// writes 128 KB, the weight ring takes 64 KB (+ 128 KB read by the image stores when training) - 4000-5000 clk at
// 128 B/clk against 4096 clk of MMA time, and an epilogue that overlaps MMAs stretches both (measured).
// Here the A operand of every hidden layer is read from tensor memory and written there by the epilogue
// (tcgen05.ld -> ReLU -> bf16 -> tcgen05.st), training images leave the registers with 16-byte global stores (a warp
// writes two full 256-byte runs of the T16 image), and shared memory only holds the weight ring and the 16 KB
// encoding blocks: 128 KB of B reads + 64 KB of ring writes per layer and SM.
//
//   * CTA pair (tcgen05 cta_group::2), ONE 128-row tile per CTA: M = 256 MMAs, each CTA streams half of every weight
//     chunk (rows [0, N/2) / [N/2, N)) through a 4 x 40 KB ring; a chunk is everything one accumulator half of one layer
//     needs (9 - 20 MMAs: a single issuer warp spends ~750 clk per chunk whatever its size - measured with 1 - 8 MMAs per
//     chunk the issuer, not the tensor pipe, set the pace);
//   * tensor memory per CTA (512 columns): A0 [0,128) | A1 [128,256) (bf16 pairs: column c = features 2c, 2c + 1;
//     layer l reads the buffer layer l - 1 wrote and its epilogue writes the other one) | D0 [256,384) | D1 [384,512);
//   * a 256-wide layer is two N = 128 halves: half 0 accumulates into D0, half 1 into D1.  The epilogue of half 0 (8 of
//     the 16 epilogue warps) runs under the MMAs of half 1, and the epilogue of half 1 under the first K = 128 steps
//     of the NEXT layer's half 0, which only need features [0,128) - written by the epilogue of half 0 long before.
//     Two mbarriers a_ready[0/1] (features [0,128) / [128,256) stored, accumulator half drained) gate the issuer;
//   * encoding / bias steps stay SS: the encoding block of the current MLP is an ordinary shared-memory operand
//     (constant-1 column x bias row, as in the SS kernel);
//   * 128-wide layers (non-rigid MLP) and the N = 16 output layers use D0 only.
//
// Outputs are those of the SS kernel bit for bit in layout (raw / xyz / offsets, T16 activation and encoding images,
// 1-bit ReLU mask images), so the backward chain and the weight-gradient kernel are unchanged.
// Reference semantics: embedders/hannw_fourier.py:38-49, embedders/fourier.py:9-38,
// non_rigid_motion_mlps/mlp_offset.py:74-114, canonical_mlps/mlp_rgb_sigma.py:132-206.
#include <cstdlib>

#include "mlp_tc_plan.cuh"
#include "mlp_tc_enc.cuh"
#include "mlp_tc_2sm.cuh"
#include "mlp_tc_ts.cuh"

namespace hn {

namespace ts {
constexpr int kStagesTs = 4;
constexpr uint32_t kStageBytesTs = kTsStageBytes;                       // per CTA: half of one (layer, accumulator half) chunk
constexpr uint32_t kRingOffTs = 2 * kActPeTile;                        // two encoding blocks of 16 KB in front
constexpr uint32_t kBarOffTs = kRingOffTs + kStagesTs * kStageBytesTs;
// barriers: full[4] floc[4] empty[4] acc_full[2] a_ready[2] pe_ready[2] out_full, then the tmem slot
constexpr uint32_t kSmemTs = kBarOffTs + 512;
constexpr int kEpiWarpsTs = 16;
constexpr int kThreadsTs = (2 + kEpiWarpsTs) * 32;                     // producer, issuer (relay in the peer), epilogue
constexpr uint32_t kColA = 0, kColD = 256;
static_assert(kSmemTs <= 232448, "shared memory budget");
}  // namespace ts

// Timeline of one pass of cluster 0 (measurement builds only: HN_EXTRA_FLAGS=-DHN_TS_TRACE ./build.sh, read with
// scripts/ts_trace.py): clock64 of the leader SM.  [0][c]: issuer, chunk c: {waits done, MMAs + commits issued};
// [1][l] / [2][l]: epilogue warp 2 (half 0) / warp 10 (half 1), layer slot l: {accumulator seen, arrive done}.
#ifdef HN_TS_TRACE
__device__ long long g_ts_trace[3][128][2];
#define HN_TS_T(role, idx, k) do { if (trace_on && lane == 0) g_ts_trace[role][idx][k] = clock64(); } while (0)
#else
#define HN_TS_T(role, idx, k) do { } while (0)
#endif

template <bool kStore>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ts::kThreadsTs, 1)
field_ts_kernel(const __grid_constant__ TsTab tab, int n_chunks, int first_chunk,
                const uint8_t* __restrict__ packed, int64_t n_cap, const int64_t* __restrict__ n_dev,
                const float* __restrict__ x_skel, const float* __restrict__ hann_w, int ignore_non_rigid,
                float* __restrict__ raw, float* __restrict__ xyz_out, float* __restrict__ off_out, const TrainBufs bufs) {
  using namespace ts;
  const int64_t n_samples = n_dev ? min(max(__ldg(n_dev), (int64_t)0), n_cap) : n_cap;
  extern __shared__ __align__(1024) uint8_t smem[];
  const TcHeader* hdr = reinterpret_cast<const TcHeader*>(packed);
  const uint8_t* image = packed + hdr->image_off;

  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_ring = s_base + kRingOffTs;
  const uint32_t s_full = s_base + kBarOffTs;                // leader: own bulk copy + forwarded arrive of the peer
  const uint32_t s_floc = s_full + kStagesTs * 8;            // peer: its own bulk copy
  const uint32_t s_empty = s_floc + kStagesTs * 8;           // both: multicast commit of the issuer
  const uint32_t s_acc_full = s_empty + kStagesTs * 8;       // both: accumulator half complete (multicast commit)
  const uint32_t s_a_ready = s_acc_full + 2 * 8;             // leader: 8 epilogue warps x 2 CTAs per half
  const uint32_t s_pe_ready = s_a_ready + 2 * 8;             // leader: 16 epilogue warps x 2 CTAs per encoding block
  const uint32_t s_out_full = s_pe_ready + 2 * 8;            // both: accumulator of an OUTPUT layer complete
  const uint32_t s_tmem_slot = s_out_full + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kBarOffTs + (3 * kStagesTs + 7) * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int64_t n_duos = (n_samples + 2 * kTileM - 1) / (2 * kTileM);           // 2 tiles = 256 samples per pass
  const int64_t n_clusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;
  const int nr_layers = (int)hdr->nr_layers, cn_layers = (int)hdr->cn_layers;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStagesTs; ++s) {
      mbar_init(s_full + s * 8, 2);
      mbar_init(s_floc + s * 8, 1);
      mbar_init(s_empty + s * 8, 1);
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(s_acc_full + h * 8, 1);
      mbar_init(s_a_ready + h * 8, 2 * (kEpiWarpsTs / 2));
      mbar_init(s_pe_ready + h * 8, 2 * kEpiWarpsTs);
    }
    mbar_init(s_out_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {            // the same warp of both CTAs allocates collectively
    tmem_alloc2(s_tmem_slot, kTmemCols);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();         // barrier inits visible before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: this CTA's half of every chunk =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t s_done = leader ? s_full : s_floc;
      for (int64_t duo = cluster_id; duo < n_duos; duo += n_clusters) {
        for (int c = first_chunk; c < n_chunks; ++c) {
          const uint32_t w_off = tab.e[c].z, half = tab.e[c].w >> 1;
          mbar_wait(s_empty + stage * 8, phase ^ 1, 1);
          mbar_expect_tx(s_done + stage * 8, half);
          tma_bulk_g2s(s_ring + stage * kStageBytesTs, image + w_off + rank * half, half, s_done + stage * 8);
          if (++stage == (uint32_t)kStagesTs) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===================== MMA issuer (leader): the whole warp runs the loop, one elected lane issues =====================
      const uint32_t desc_hi = (128u >> 4) | (1u << 14);
      const uint32_t pe_lo0 = ((s_base & 0x3FFFFu) >> 4) | ((kLboA >> 4) << 16);
      const uint32_t pe_lo1 = (((s_base + kActPeTile) & 0x3FFFFu) >> 4) | ((kLboA >> 4) << 16);
      const uint32_t ring_lo = (s_ring & 0x3FFFFu) >> 4;
      uint32_t stage = 0, phase = 0, full_ok = 0;
      uint32_t a_phase0 = 0, a_phase1 = 0, pe_phase0 = 0, pe_phase1 = 0;
      int pass = 0;
      for (int64_t duo = cluster_id; duo < n_duos; duo += n_clusters, ++pass) {
#ifdef HN_TS_TRACE
        const bool trace_on = cluster_id == 0 && pass == 3;
#endif
        for (int c = first_chunk; c < n_chunks; ++c) {
          const uint32_t ex = tab.e[c].x, ey = tab.e[c].y;
          const uint32_t n = ex & 0xFFFFu, flags = (ex >> 16) & 0xFFu, flags2 = ex >> 24;
          const uint32_t ss_a_off = ey & 0xFFu, ss_k = (ey >> 8) & 0xFFu, ts0_k = (ey >> 16) & 0xFFu, ts1_k = ey >> 24;
          if (flags & kTsWaitPe) {
            if (flags & kTsPeCn) { mbar_wait(s_pe_ready + 8, pe_phase1, 6); pe_phase1 ^= 1; }
            else { mbar_wait(s_pe_ready, pe_phase0, 6); pe_phase0 ^= 1; }
          }
          if (flags & kTsWaitA0) { mbar_wait(s_a_ready, a_phase0, 2); a_phase0 ^= 1; }
          if (!full_ok) mbar_wait(s_full + stage * 8, phase, 3);
          tc_fence_after();
          HN_TS_T(0, c, 0);
          {   // probe the NEXT chunk's weights now: the round trip overlaps the MMA issue below
            const uint32_t ns = stage + 1 == (uint32_t)kStagesTs ? 0u : stage + 1;
            full_ok = mbar_test_wait(s_full + ns * 8, ns == 0u ? phase ^ 1u : phase);
          }
          const uint32_t idesc = make_idesc2(n);
          const uint32_t d_tmem = tmem_base + kColD + ((flags & kTsD1) ? 128u : 0u);
          const uint32_t a_t0 = tmem_base + kColA + ((flags2 & 1u) ? 128u : 0u);
          const uint32_t b_lo0 = (ring_lo + stage * (kStageBytesTs >> 4)) | ((n >> 1) << 16);    // LBO = (n/2) * 16 bytes
          const uint32_t b_lo1 = b_lo0 + (ss_k + ts0_k) * n;          // one K step = two core-matrix groups of n/2 rows x 16 B, >> 4
          if (elect_one()) {
            // the chunk is a whole accumulator half of a layer: the first MMA overwrites
            uint32_t acc = 0u;
            uint32_t b_lo = b_lo0;
            uint32_t a_lo = ((flags & kTsPeCn) ? pe_lo1 : pe_lo0) + ss_a_off * (kLboA >> 4);
            for (uint32_t k = 0; k < ss_k; ++k) {
              umma2_bf16(d_tmem, ((uint64_t)desc_hi << 32) | a_lo, ((uint64_t)desc_hi << 32) | b_lo, idesc, acc);
              acc = 1u;
              a_lo += (2u * kLboA) >> 4;
              b_lo += n;
            }
            uint32_t a_t = a_t0;
            for (uint32_t k = 0; k < ts0_k; ++k) {
              umma2_ts_bf16(d_tmem, a_t, ((uint64_t)desc_hi << 32) | b_lo, idesc, 1u);
              a_t += 8u;                                     // 16 bf16 of K = 8 columns
              b_lo += n;
            }
          }
          __syncwarp();
          if (flags & kTsWaitA1) {                           // features [128,256) of the previous layer stored, D1 drained
            mbar_wait(s_a_ready + 8, a_phase1, 2);
            a_phase1 ^= 1;
            tc_fence_after();
          }
          if (elect_one()) {
            uint32_t a_t = a_t0 + 64u;
            uint32_t b_lo = b_lo1;
            for (uint32_t k = 0; k < ts1_k; ++k) {
              umma2_ts_bf16(d_tmem, a_t, ((uint64_t)desc_hi << 32) | b_lo, idesc, 1u);
              a_t += 8u;
              b_lo += n;
            }
            umma2_commit_mc(s_empty + stage * 8);            // the slot is free in both CTAs once these MMAs retire
            umma2_commit_mc((flags & kTsOut) ? s_out_full : s_acc_full + ((flags & kTsD1) ? 8u : 0u));
          }
          __syncwarp();
          HN_TS_T(0, c, 1);
          if (++stage == (uint32_t)kStagesTs) { stage = 0; phase ^= 1; }
        }
      }
    } else if (lane == 0) {
      // ===================== relay (peer): forwards "my half of the chunk has landed" to the leader =====================
      const uint32_t full_remote = mapa_cluster(s_full, 0);
      uint32_t stage = 0, phase = 0;
      for (int64_t duo = cluster_id; duo < n_duos; duo += n_clusters) {
        for (int c = first_chunk; c < n_chunks; ++c) {
          mbar_wait(s_floc + stage * 8, phase, 5);
          mbar_arrive_remote(full_remote + stage * 8);
          if (++stage == (uint32_t)kStagesTs) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue warps (2 .. 17) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int nh = ew >> 3;                       // accumulator half this warp drains (features [128 nh, 128 nh + 128))
    const int sub = (ew >> 2) & 1;                // which 64 of those 128 columns
    const int part = nh * 2 + sub;                // 0..3: share of the per-row work among the four warps of a quarter
    const int row = quarter * 32 + lane;          // row inside the tile == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t pe_nr = s_base, pe_cn = s_base + kActPeTile;
    const uint32_t bar_a = mapa_cluster(s_a_ready + (uint32_t)nh * 8, 0);            // in the leader
    const uint32_t bar_pe0 = mapa_cluster(s_pe_ready, 0), bar_pe1 = mapa_cluster(s_pe_ready + 8, 0);
    uint32_t acc_phase0 = 0, acc_phase1 = 0, out_phase = 0;
    const int nr_freqs = (int)hdr->nr_freqs, cn_freqs = (int)hdr->cn_freqs;
    const int nr_pad = (int)hdr->nr_pe_pad, cn_pad = (int)hdr->cn_pe_pad;
    const int nr_width = (int)hdr->nr_width, cn_width = (int)hdr->cn_width;

    // encoding block of one MLP for this thread's row: the four warps of a quarter write two 8-column chunks each
    auto encode = [&](bool cn, const float p[3], int64_t tile_idx, bool store) {
      const int pad = cn ? cn_pad : nr_pad;
      uint8_t* img = (kStore && store) ? bufs.pe[cn ? 1 : 0] + image_row_offset(tile_idx, row, pad) : nullptr;
      if (cn) write_encoding_part<true>(pe_cn, row, part, p, cn_freqs, nullptr, pad, img);
      else write_encoding_part<false>(pe_nr, row, part, p, nr_freqs, hann_w, pad, img);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(cn ? bar_pe1 : bar_pe0);
    };
    auto load_x = [&](int64_t duo, float p[3]) {
      const int64_t sample = (duo * 2 + (int64_t)rank) * kTileM + row;
      const bool ok = duo < n_duos && sample < n_samples;
      p[0] = ok ? __ldg(x_skel + sample * 3 + 0) : 0.f;
      p[1] = ok ? __ldg(x_skel + sample * 3 + 1) : 0.f;
      p[2] = ok ? __ldg(x_skel + sample * 3 + 2) : 0.f;
    };
    // hidden layer `l` of MLP `m` (width `width`): drain this warp's 64 accumulator columns into the A buffer (l & 1)
    int tr_slot = 0;
    bool trace_on = false;
    (void)tr_slot; (void)trace_on;
    auto hidden = [&](int m, int l, int width, int64_t tile_idx, bool store) {
      const bool two = width > 128;
      // Each barrier is waited for ONLY by the warps whose arrivals gate its next phase (acc_full[0]: group 0 through
      // a_ready[0]; acc_full[1]: group 1 through a_ready[1]), so no waiter can fall two phases behind and alias the
      // parity.  (A version in which group 0 also waited for acc_full[1] "to stay in step" dead-locked under the
      // training variant's stores: a warp delayed between its a_ready[0] arrive and that wait missed two phases.)
      if (nh == 0) { mbar_wait(s_acc_full, acc_phase0, 4); acc_phase0 ^= 1; }
      else if (two) { mbar_wait(s_acc_full + 8, acc_phase1, 4); acc_phase1 ^= 1; }
      if (two || nh == 0) {
        tc_fence_after();
        HN_TS_T(1 + nh, tr_slot, 0);
        const int col0 = nh * 128 + sub * 64;                                  // first feature this warp owns
        const uint32_t t_d = t_lane + kColD + (uint32_t)col0;
        const uint32_t t_a = t_lane + kColA + ((l & 1) ? 128u : 0u) + (uint32_t)(col0 >> 1);
        uint8_t* img = nullptr;
        uint8_t* bits_rec = nullptr;
        if (kStore && store) {
          img = bufs.h[m][l] + image_row_offset(tile_idx, row, width) + (size_t)(col0 >> 3) * 256;
          uint8_t* b = bufs.bits[m][l];
          // record (tile, half, row) of width / 2 columns: this warp's 64 columns are words [0,2) of half `sub`
          // (width 128) or words [2 sub, 2 sub + 2) of half `nh` (width 256)
          if (b) bits_rec = two ? b + relu_bits_offset(tile_idx, nh, row, 128) + sub * 8
                                : b + relu_bits_offset(tile_idx, sub, row, 64);
        }
        uint32_t pk[32];
        uint32_t m01[2] = {0u, 0u};
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t v[32];
          tmem_ld32(t_d + (uint32_t)(g * 32), v);
          tmem_ld_wait();
          uint32_t mm = 0u;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t h2 = pack_relu_bf16(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
            pk[g * 16 + i] = h2;
            if (kStore) mm = relu_bits_step(mm, h2);
          }
          m01[g] = mm;
        }
        tmem_st32(t_a, pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(bar_a);
        HN_TS_T(1 + nh, tr_slot, 1);
        if (kStore && img) {
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8)
            *reinterpret_cast<uint4*>(img + (size_t)c8 * 256) = make_uint4(pk[4 * c8], pk[4 * c8 + 1], pk[4 * c8 + 2], pk[4 * c8 + 3]);
          if (bits_rec) *reinterpret_cast<uint2*>(bits_rec) = make_uint2(m01[0], m01[1]);
        }
      }
      ++tr_slot;
    };

    float x[3];
    load_x(cluster_id, x);
    if (cluster_id < n_duos) {
      const int64_t t0 = cluster_id * 2 + (int64_t)rank;
      encode(ignore_non_rigid != 0, x, t0, t0 < bufs.n_tiles);
    }
    int pass = 0;
    for (int64_t duo = cluster_id; duo < n_duos; duo += n_clusters, ++pass) {
      tr_slot = 0;
      trace_on = blockIdx.x == 0 && pass == 3 && (ew == 0 || ew == 8);
      const int64_t tile_idx = duo * 2 + (int64_t)rank;
      const int64_t sample = tile_idx * kTileM + row;
      const bool valid = sample < n_samples;
      const bool store = kStore && tile_idx < bufs.n_tiles;      // the training images hold ceil(P/128) tiles
      float q[3] = {x[0], x[1], x[2]};                           // position fed to the canonical encoding
      float xn[3];
      load_x(duo + n_clusters, xn);                              // the next pass's positions: ~1 us of latency hidden under this pass
      if (!ignore_non_rigid) {
        for (int l = 0; l + 1 < nr_layers; ++l) hidden(0, l, nr_width, tile_idx, store);
        // output layer: offsets in accumulator columns [0,3) of D0; every warp reads them (all need q for the encoding).
        // out_full is waited for by all 16 warps, and all of them arrive on pe_ready between two of its phases.
        mbar_wait(s_out_full, out_phase, 4);
        out_phase ^= 1;
        tc_fence_after();
        HN_TS_T(1 + nh, tr_slot, 0);
        uint32_t v[16];
        tmem_ld16(t_lane + kColD, v);
        tmem_ld_wait();
        tc_fence_before();
        const float o0 = __uint_as_float(v[0]), o1 = __uint_as_float(v[1]), o2 = __uint_as_float(v[2]);
        q[0] = x[0] + o0; q[1] = x[1] + o1; q[2] = x[2] + o2;
        if (valid && part == 0) {
          off_out[sample * 3 + 0] = o0; off_out[sample * 3 + 1] = o1; off_out[sample * 3 + 2] = o2;
        }
        encode(true, q, tile_idx, store);
        HN_TS_T(1 + nh, tr_slot, 1);
        ++tr_slot;
      }
      for (int l = 0; l + 1 < cn_layers; ++l) hidden(1, l, cn_width, tile_idx, store);
      {
        mbar_wait(s_out_full, out_phase, 4);
        out_phase ^= 1;
        tc_fence_after();
        HN_TS_T(1 + nh, tr_slot, 0);
        if (part == 0) {
          uint32_t v[16];
          tmem_ld16(t_lane + kColD, v);
          tmem_ld_wait();
          if (valid) {
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
        tc_fence_before();
      }
      // next pass: positions and first encoding block (its arrive also tells the issuer that D0 has been drained)
      const int64_t next = duo + n_clusters;
      x[0] = xn[0]; x[1] = xn[1]; x[2] = xn[2];
      if (next < n_duos) {
        const int64_t tn = next * 2 + (int64_t)rank;
        encode(ignore_non_rigid != 0, x, tn, kStore && tn < bufs.n_tiles);
      }
      HN_TS_T(1 + nh, tr_slot, 1);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // the peer may still arrive on / commit into this CTA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, kTmemCols);
  }
}

// Launch: persistent, as many clusters as can be co-resident.  Returns <0 / cudaError_t like every entry point.
int launch_field_ts(const TcPlan& plan, const uint8_t* packed, int64_t n_samples, const int64_t* n_samples_dev, const float* x_skel,
                    const float* hann_w, int ignore_non_rigid, float* raw, float* xyz, float* offsets,
                    const TrainBufs* bufs, int sms, cudaStream_t st) {
  using namespace ts;
  if (plan.ts_chunks.empty()) return fail_arg("hn_field_fwd: the TS kernel needs MLP widths of 128 or 256");
  static int max_clusters_cached = 0;
  if (max_clusters_cached == 0) {
    HN_TRY(check_cuda(cudaFuncSetAttribute(field_ts_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTs),
                      "hn_field_fwd/ts smem attr"));
    HN_TRY(check_cuda(cudaFuncSetAttribute(field_ts_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTs),
                      "hn_field_fwd/ts smem attr"));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * (sms / 2)));
    cfg.blockDim = dim3(kThreadsTs);
    cfg.dynamicSmemBytes = kSmemTs;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    HN_TRY(check_cuda(cudaOccupancyMaxActiveClusters(&n, field_ts_kernel<false>, &cfg), "hn_field_fwd/ts occupancy"));
    if (n <= 0) return fail_arg("hn_field_fwd: no 2-CTA cluster of the TS kernel fits on this device");
    max_clusters_cached = n;
  }
  TsTab tab;
  int nr_chunks = 0;
  fill_ts_tab(plan, &tab, &nr_chunks);
  const int first_chunk = ignore_non_rigid ? nr_chunks : 0;      // chunks of the skipped MLP are skipped too
  const int64_t n_duos = ceil_div<int64_t>(n_samples, 2 * kTileM);
  const unsigned grid = 2u * (unsigned)std::min<int64_t>(n_duos, (int64_t)max_clusters_cached);
  const int n_chunks = (int)plan.ts_chunks.size();
  if (bufs) {
    field_ts_kernel<true><<<grid, kThreadsTs, kSmemTs, st>>>(tab, n_chunks, first_chunk, packed, n_samples, n_samples_dev,
                                                             x_skel, hann_w, ignore_non_rigid, raw, xyz, offsets, *bufs);
  } else {
    TrainBufs none;
    memset(&none, 0, sizeof(none));
    field_ts_kernel<false><<<grid, kThreadsTs, kSmemTs, st>>>(tab, n_chunks, first_chunk, packed, n_samples, n_samples_dev,
                                                              x_skel, hann_w, ignore_non_rigid, raw, xyz, offsets, none);
  }
  return launched("hn_field_fwd/tcgen05 ts");
}

#ifdef HN_TS_TRACE
extern "C" int hn_ts_trace_read(long long* out_host) {      // measurement builds only; not declared in the public header
  return (int)cudaMemcpyFromSymbol(out_host, g_ts_trace, sizeof(long long) * 3 * 128 * 2);
}
#endif

}  // namespace hn
