// Canonical field, tensor-core forward for RENDERING (nothing saved for backward), second TS-form kernel: like
// mlp_tc_ts.cu (read its header first: hidden activations in tensor memory, CTA pairs, N-half pipelining of the 256-wide
// layers), but a CTA works on TWO 128-row tiles per pass so that the non-rigid MLP - 128-wide layers, 17 % of the FLOPs,
// but one serial [MMAs -> epilogue -> hand-over] chain per layer: a third of the time of the one-tile kernel (timeline in
// DESIGN 4.1) - is evaluated for both tiles interleaved: tile set X's epilogue runs under tile set Y's MMAs and vice
// versa, and the non-rigid weight chunks are fetched once for both.
//
//   pass = 4 tiles: CTA r owns X = tile 4q + r and Y = tile 4q + 2 + r (set X = {X of both CTAs}, one M = 256 MMA)
//   N phase (non-rigid MLP, both sets):  tensor memory  set s: A0 [256 s, +64) | A1 [256 s + 64, +64) | D [256 s + 128, +128)
//        per layer: chunk fetched once; issue X (after n_ready[X]) then Y (after n_ready[Y]); epilogue group s (8 warps)
//        drains set s.  The output layer's epilogue writes offsets / xyz and the canonical encoding block of its set.
//   C phase (canonical MLP), set X then set Y, exactly the schedule of mlp_tc_ts.cu (all 16 epilogue warps on one set):
//        A0 [0,128) | A1 [128,256) | D0 [256,384) | D1 [384,512)
//   barriers (leader unless noted): n_acc[s] (both CTAs), n_ready[s], n_pe[s] (non-rigid encoding of set s written and
//        the accumulators of the previous pass drained), c_pe[s] (canonical encoding of set s written = N phase of set s
//        complete), c_done (output accumulator of set X drained), plus acc_full / a_ready / out_full of the C phase.
//        Every barrier is waited for only by parties whose own arrivals gate its next phase (no parity aliasing).
// Needs the reference widths (non-rigid 128, canonical 256); other widths use the one-tile kernel.
// Reference semantics: embedders/hannw_fourier.py:38-49, embedders/fourier.py:9-38,
// non_rigid_motion_mlps/mlp_offset.py:74-114, canonical_mlps/mlp_rgb_sigma.py:132-206.
#include <cstdlib>

#include "mlp_tc_plan.cuh"
#include "mlp_tc_enc.cuh"
#include "mlp_tc_2sm.cuh"
#include "mlp_tc_ts.cuh"

namespace hn {

namespace ts2 {
constexpr int kStages2T = 4;
constexpr uint32_t kStageBytes2T = kTsStageBytes;                        // per CTA: half of one (layer, accumulator half) chunk
constexpr uint32_t kPeOff2T = 0;                                         // encoding blocks: nr X, nr Y, cn X, cn Y (16 KB each)
constexpr uint32_t kRingOff2T = 4 * kActPeTile;
constexpr uint32_t kBarOff2T = kRingOff2T + kStages2T * kStageBytes2T;
// barriers: full[4] floc[4] empty[4] | n_acc[2] n_ready[2] n_pe[2] c_pe[2] | acc_full[2] a_ready[2] out_full c_done | tmem slot
constexpr uint32_t kSmem2T = kBarOff2T + 512;
constexpr int kEpiWarps2T = 16;
constexpr int kThreads2T = (2 + kEpiWarps2T) * 32;
static_assert(kSmem2T <= 232448, "shared memory budget");
}  // namespace ts2

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ts2::kThreads2T, 1)
field_ts2_kernel(const __grid_constant__ TsTab tab, int n_chunks, int nr_chunks, const uint8_t* __restrict__ packed, int64_t n_cap,
                 const int64_t* __restrict__ n_dev, const float* __restrict__ x_skel, const float* __restrict__ hann_w,
                 int ignore_non_rigid, float* __restrict__ raw, float* __restrict__ xyz_out, float* __restrict__ off_out) {
  using namespace ts2;
  const int64_t n_samples = n_dev ? min(max(__ldg(n_dev), (int64_t)0), n_cap) : n_cap;
  extern __shared__ __align__(1024) uint8_t smem[];
  const TcHeader* hdr = reinterpret_cast<const TcHeader*>(packed);
  const uint8_t* image = packed + hdr->image_off;

  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_ring = s_base + kRingOff2T;
  const uint32_t s_full = s_base + kBarOff2T;
  const uint32_t s_floc = s_full + kStages2T * 8;
  const uint32_t s_empty = s_floc + kStages2T * 8;
  const uint32_t s_n_acc = s_empty + kStages2T * 8;
  const uint32_t s_n_ready = s_n_acc + 16;
  const uint32_t s_n_pe = s_n_ready + 16;
  const uint32_t s_c_pe = s_n_pe + 16;
  const uint32_t s_acc_full = s_c_pe + 16;
  const uint32_t s_a_ready = s_acc_full + 16;
  const uint32_t s_out_full = s_a_ready + 16;
  const uint32_t s_c_done = s_out_full + 8;
  const uint32_t s_tmem_slot = s_c_done + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kBarOff2T + (3 * kStages2T + 14) * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int64_t n_quads = (n_samples + 4 * kTileM - 1) / (4 * kTileM);            // 4 tiles = 512 samples per pass
  const int64_t n_clusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;
  const int nr_layers = (int)hdr->nr_layers, cn_layers = (int)hdr->cn_layers;
  const int c_first = ignore_non_rigid ? nr_chunks : 0;                             // first chunk of a pass

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages2T; ++s) {
      mbar_init(s_full + s * 8, 2);
      mbar_init(s_floc + s * 8, 1);
      mbar_init(s_empty + s * 8, 1);
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(s_n_acc + h * 8, 1);
      mbar_init(s_n_ready + h * 8, 2 * (kEpiWarps2T / 2));
      mbar_init(s_n_pe + h * 8, 2 * (kEpiWarps2T / 2));
      mbar_init(s_c_pe + h * 8, 2 * (kEpiWarps2T / 2));
      mbar_init(s_acc_full + h * 8, 1);
      mbar_init(s_a_ready + h * 8, 2 * (kEpiWarps2T / 2));
    }
    mbar_init(s_out_full, 1);
    mbar_init(s_c_done, 2 * kEpiWarps2T);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc2(s_tmem_slot, kTmemCols);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: non-rigid chunks once per pass, canonical chunks once per tile set =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t s_done = leader ? s_full : s_floc;
      auto load = [&](int c) {
        const uint32_t w_off = tab.e[c].z, half = tab.e[c].w >> 1;
        mbar_wait(s_empty + stage * 8, phase ^ 1, 1);
        mbar_expect_tx(s_done + stage * 8, half);
        tma_bulk_g2s(s_ring + stage * kStageBytes2T, image + w_off + rank * half, half, s_done + stage * 8);
        if (++stage == (uint32_t)kStages2T) { stage = 0; phase ^= 1; }
      };
      for (int64_t quad = cluster_id; quad < n_quads; quad += n_clusters) {
        for (int c = c_first; c < nr_chunks; ++c) load(c);
        for (int set = 0; set < 2; ++set)
          for (int c = nr_chunks; c < n_chunks; ++c) load(c);
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===================== MMA issuer (leader) =====================
      const uint32_t desc_hi = (128u >> 4) | (1u << 14);
      const uint32_t ring_lo = (s_ring & 0x3FFFFu) >> 4;
      uint32_t stage = 0, phase = 0, full_ok = 0;
      uint32_t ph_n_ready[2] = {0, 0}, ph_n_pe[2] = {0, 0}, ph_c_pe[2] = {0, 0}, ph_a[2] = {0, 0}, ph_c_done = 0;
      // one MMA group: `n_ss` SS steps from the encoding block at shared offset pe_off (+ first core matrix ss_a_off), then
      // ts_k TS steps from tensor-memory columns a_t; the first MMA overwrites the accumulator when `fresh`
      auto issue = [&](uint32_t d_tmem, uint32_t idesc, uint32_t n, uint32_t b_lo, uint32_t pe_off, uint32_t ss_a_off, uint32_t ss_k,
                       uint32_t a_t, uint32_t ts_k, bool fresh) -> uint32_t {
        if (elect_one()) {
          uint32_t acc = fresh ? 0u : 1u;
          uint32_t a_lo = (((s_base + pe_off) & 0x3FFFFu) >> 4 | ((kLboA >> 4) << 16)) + ss_a_off * (kLboA >> 4);
          uint32_t b = b_lo;
          for (uint32_t k = 0; k < ss_k; ++k) {
            umma2_bf16(d_tmem, ((uint64_t)desc_hi << 32) | a_lo, ((uint64_t)desc_hi << 32) | b, idesc, acc);
            acc = 1u; a_lo += (2u * kLboA) >> 4; b += n;
          }
          for (uint32_t k = 0; k < ts_k; ++k) {
            umma2_ts_bf16(d_tmem, a_t, ((uint64_t)desc_hi << 32) | b, idesc, acc);
            acc = 1u; a_t += 8u; b += n;
          }
        }
        __syncwarp();
        return b_lo + (ss_k + ts_k) * n;
      };
      auto wait_full = [&]() {
        if (!full_ok) mbar_wait(s_full + stage * 8, phase, 3);
        tc_fence_after();
        const uint32_t ns = stage + 1 == (uint32_t)kStages2T ? 0u : stage + 1;
        full_ok = mbar_test_wait(s_full + ns * 8, ns == 0u ? phase ^ 1u : phase);      // probe the next chunk's weights
      };
      auto release = [&](uint32_t extra_bar) {
        if (elect_one()) {
          umma2_commit_mc(s_empty + stage * 8);
          if (extra_bar) umma2_commit_mc(extra_bar);
        }
        __syncwarp();
        if (++stage == (uint32_t)kStages2T) { stage = 0; phase ^= 1; }
      };
      for (int64_t quad = cluster_id; quad < n_quads; quad += n_clusters) {
        // ---------------- N phase: every non-rigid chunk feeds set X, then set Y ----------------
        for (int c = c_first; c < nr_chunks; ++c) {
          const uint32_t ex = tab.e[c].x, ey = tab.e[c].y;
          const uint32_t n = ex & 0xFFFFu, flags = (ex >> 16) & 0xFFu, flags2 = ex >> 24;
          const uint32_t ss_a_off = ey & 0xFFu, ss_k = (ey >> 8) & 0xFFu, ts0_k = (ey >> 16) & 0xFFu;
          const uint32_t idesc = make_idesc2(n);
          for (uint32_t s = 0; s < 2; ++s) {
            if (flags & kTsWaitPe) { mbar_wait(s_n_pe + s * 8, ph_n_pe[s], 6); ph_n_pe[s] ^= 1; }
            else { mbar_wait(s_n_ready + s * 8, ph_n_ready[s], 2); ph_n_ready[s] ^= 1; }
            if (s == 0) wait_full();
            else tc_fence_after();
            const uint32_t b_lo = (ring_lo + stage * (kStageBytes2T >> 4)) | ((n >> 1) << 16);
            issue(tmem_base + 256u * s + 128u, idesc, n, b_lo, kPeOff2T + s * kActPeTile, ss_a_off, ss_k,
                  tmem_base + 256u * s + ((flags2 & 1u) ? 64u : 0u), ts0_k, true);
            if (s == 0) {
              if (elect_one()) umma2_commit_mc(s_n_acc);
              __syncwarp();
            } else {
              release(s_n_acc + 8);
            }
          }
        }
        // ---------------- C phase: set X, then set Y ----------------
        for (uint32_t s = 0; s < 2; ++s) {
          for (int c = nr_chunks; c < n_chunks; ++c) {
            const uint32_t ex = tab.e[c].x, ey = tab.e[c].y;
            const uint32_t n = ex & 0xFFFFu, flags = (ex >> 16) & 0xFFu, flags2 = ex >> 24;
            const uint32_t ss_a_off = ey & 0xFFu, ss_k = (ey >> 8) & 0xFFu, ts0_k = (ey >> 16) & 0xFFu, ts1_k = ey >> 24;
            if (flags & kTsWaitPe) {
              if (s == 0) {
                // both sets have left the N phase: their tensor-memory regions are free, the encoding blocks written
                mbar_wait(s_c_pe, ph_c_pe[0], 6); ph_c_pe[0] ^= 1;
                mbar_wait(s_c_pe + 8, ph_c_pe[1], 6); ph_c_pe[1] ^= 1;
              } else {
                mbar_wait(s_c_done, ph_c_done, 6); ph_c_done ^= 1;              // set X's output accumulator drained
              }
            }
            if (flags & kTsWaitA0) { mbar_wait(s_a_ready, ph_a[0], 2); ph_a[0] ^= 1; }
            wait_full();
            const uint32_t idesc = make_idesc2(n);
            const uint32_t d_tmem = tmem_base + 256u + ((flags & kTsD1) ? 128u : 0u);
            const uint32_t a_t0 = tmem_base + ((flags2 & 1u) ? 128u : 0u);
            const uint32_t b_lo0 = (ring_lo + stage * (kStageBytes2T >> 4)) | ((n >> 1) << 16);
            const uint32_t b_lo1 = issue(d_tmem, idesc, n, b_lo0, kPeOff2T + (2 + s) * kActPeTile, ss_a_off, ss_k, a_t0, ts0_k, true);
            if (flags & kTsWaitA1) {
              mbar_wait(s_a_ready + 8, ph_a[1], 2); ph_a[1] ^= 1;
              tc_fence_after();
            }
            issue(d_tmem, idesc, n, b_lo1, 0, 0, 0, a_t0 + 64u, ts1_k, false);
            release((flags & kTsOut) ? s_out_full : s_acc_full + ((flags & kTsD1) ? 8u : 0u));
          }
        }
      }
    } else if (lane == 0) {
      // ===================== relay (peer) =====================
      const uint32_t full_remote = mapa_cluster(s_full, 0);
      uint32_t stage = 0, phase = 0;
      const int per_pass = (nr_chunks - c_first) + 2 * (n_chunks - nr_chunks);
      for (int64_t quad = cluster_id; quad < n_quads; quad += n_clusters) {
        for (int c = 0; c < per_pass; ++c) {
          mbar_wait(s_floc + stage * 8, phase, 5);
          mbar_arrive_remote(full_remote + stage * 8);
          if (++stage == (uint32_t)kStages2T) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue warps (2 .. 17) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int grp = ew >> 3;                      // N phase: tile set this warp serves; C phase: accumulator half it drains
    const int sub = (ew >> 2) & 1;                // which 64 of the 128 columns
    const int part = grp * 2 + sub;
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t pe_nr = s_base + kPeOff2T + (uint32_t)grp * kActPeTile;             // this group's set (N phase)
    const uint32_t pe_cn_own = s_base + kPeOff2T + (uint32_t)(2 + grp) * kActPeTile;
    const uint32_t bar_n_ready = mapa_cluster(s_n_ready + (uint32_t)grp * 8, 0);
    const uint32_t bar_n_pe = mapa_cluster(s_n_pe + (uint32_t)grp * 8, 0);
    const uint32_t bar_c_pe = mapa_cluster(s_c_pe + (uint32_t)grp * 8, 0);
    const uint32_t bar_a = mapa_cluster(s_a_ready + (uint32_t)grp * 8, 0);
    const uint32_t bar_c_done = mapa_cluster(s_c_done, 0);
    uint32_t ph_n_acc = 0, ph_acc = 0, ph_out = 0;
    const int nr_freqs = (int)hdr->nr_freqs, cn_freqs = (int)hdr->cn_freqs;
    const int nr_pad = (int)hdr->nr_pe_pad, cn_pad = (int)hdr->cn_pe_pad;

    // this group's share of an encoding row: the two warps (sub) of a quarter write alternate 8-column chunks
    auto encode_nr = [&](const float p[3]) {
      if (sub == 0) {
        write_encoding_chunk<false, 0>(pe_nr, row, p, nr_freqs, hann_w, nr_pad, nullptr); write_encoding_chunk<false, 2>(pe_nr, row, p, nr_freqs, hann_w, nr_pad, nullptr);
        write_encoding_chunk<false, 4>(pe_nr, row, p, nr_freqs, hann_w, nr_pad, nullptr); write_encoding_chunk<false, 6>(pe_nr, row, p, nr_freqs, hann_w, nr_pad, nullptr);
      } else {
        write_encoding_chunk<false, 1>(pe_nr, row, p, nr_freqs, hann_w, nr_pad, nullptr); write_encoding_chunk<false, 3>(pe_nr, row, p, nr_freqs, hann_w, nr_pad, nullptr);
        write_encoding_chunk<false, 5>(pe_nr, row, p, nr_freqs, hann_w, nr_pad, nullptr); write_encoding_chunk<false, 7>(pe_nr, row, p, nr_freqs, hann_w, nr_pad, nullptr);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(bar_n_pe);
    };
    auto encode_cn = [&](const float p[3]) {
      if (sub == 0) {
        write_encoding_chunk<true, 0>(pe_cn_own, row, p, cn_freqs, nullptr, cn_pad, nullptr); write_encoding_chunk<true, 2>(pe_cn_own, row, p, cn_freqs, nullptr, cn_pad, nullptr);
        write_encoding_chunk<true, 4>(pe_cn_own, row, p, cn_freqs, nullptr, cn_pad, nullptr); write_encoding_chunk<true, 6>(pe_cn_own, row, p, cn_freqs, nullptr, cn_pad, nullptr);
      } else {
        write_encoding_chunk<true, 1>(pe_cn_own, row, p, cn_freqs, nullptr, cn_pad, nullptr); write_encoding_chunk<true, 3>(pe_cn_own, row, p, cn_freqs, nullptr, cn_pad, nullptr);
        write_encoding_chunk<true, 5>(pe_cn_own, row, p, cn_freqs, nullptr, cn_pad, nullptr); write_encoding_chunk<true, 7>(pe_cn_own, row, p, cn_freqs, nullptr, cn_pad, nullptr);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(bar_c_pe);
    };
    // sample index of this thread's row in this group's OWN tile (N phase): tile 4 quad + 2 grp + rank
    auto own_sample = [&](int64_t quad) { return ((quad * 4 + 2 * grp + (int64_t)rank) * kTileM + row); };
    auto load_x = [&](int64_t quad, float p[3]) {
      const int64_t sample = own_sample(quad);
      const bool ok = quad < n_quads && sample < n_samples;
      p[0] = ok ? __ldg(x_skel + sample * 3 + 0) : 0.f;
      p[1] = ok ? __ldg(x_skel + sample * 3 + 1) : 0.f;
      p[2] = ok ? __ldg(x_skel + sample * 3 + 2) : 0.f;
    };
    // start of a pass for this group's set: the non-rigid encoding, or (non-rigid MLP skipped) outputs + canonical encoding
    auto open_pass = [&](int64_t quad, const float p[3]) {
      if (!ignore_non_rigid) { encode_nr(p); return; }
      const int64_t sample = own_sample(quad);
      if (sub == 0 && sample < n_samples) {
        xyz_out[sample * 3 + 0] = p[0]; xyz_out[sample * 3 + 1] = p[1]; xyz_out[sample * 3 + 2] = p[2];
        off_out[sample * 3 + 0] = 0.f; off_out[sample * 3 + 1] = 0.f; off_out[sample * 3 + 2] = 0.f;
      }
      encode_cn(p);
    };
    // drain 64 accumulator columns at t_d into 32 operand columns at t_a (ReLU, bf16 pairs), then arrive on `bar`
    auto drain = [&](uint32_t t_d, uint32_t t_a, uint32_t bar) {
      uint32_t pk[32];
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t v[32];
        tmem_ld32(t_d + (uint32_t)(g * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[g * 16 + i] = pack_relu_bf16(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
      }
      tmem_st32(t_a, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(bar);
    };

    float x[3];
    load_x(cluster_id, x);
    if (cluster_id < n_quads) open_pass(cluster_id, x);
    for (int64_t quad = cluster_id; quad < n_quads; quad += n_clusters) {
      float xn[3];
      load_x(quad + n_clusters, xn);                             // next pass's positions, hidden under this pass
      // ---------------- N phase: this group's own set ----------------
      if (!ignore_non_rigid) {
        const uint32_t t_set = t_lane + 256u * (uint32_t)grp;
        for (int l = 0; l + 1 < nr_layers; ++l) {
          mbar_wait(s_n_acc + grp * 8, ph_n_acc, 4);
          ph_n_acc ^= 1;
          tc_fence_after();
          drain(t_set + 128u + (uint32_t)(sub * 64), t_set + ((l & 1) ? 64u : 0u) + (uint32_t)(sub * 32), bar_n_ready);
        }
        mbar_wait(s_n_acc + grp * 8, ph_n_acc, 4);
        ph_n_acc ^= 1;
        tc_fence_after();
        uint32_t v[16];
        tmem_ld16(t_set + 128u, v);
        tmem_ld_wait();
        tc_fence_before();
        const float o0 = __uint_as_float(v[0]), o1 = __uint_as_float(v[1]), o2 = __uint_as_float(v[2]);
        const float q[3] = {x[0] + o0, x[1] + o1, x[2] + o2};
        const int64_t sample = own_sample(quad);
        if (sub == 0 && sample < n_samples) {
          off_out[sample * 3 + 0] = o0; off_out[sample * 3 + 1] = o1; off_out[sample * 3 + 2] = o2;
          xyz_out[sample * 3 + 0] = q[0]; xyz_out[sample * 3 + 1] = q[1]; xyz_out[sample * 3 + 2] = q[2];
        }
        encode_cn(q);
      }
      // ---------------- C phase: set X, then set Y; this warp drains accumulator half `grp` ----------------
      for (int s = 0; s < 2; ++s) {
        for (int l = 0; l + 1 < cn_layers; ++l) {
          mbar_wait(s_acc_full + grp * 8, ph_acc, 4);
          ph_acc ^= 1;
          tc_fence_after();
          const int col0 = grp * 128 + sub * 64;
          drain(t_lane + 256u + (uint32_t)col0, t_lane + ((l & 1) ? 128u : 0u) + (uint32_t)(col0 >> 1), bar_a);
        }
        mbar_wait(s_out_full, ph_out, 4);
        ph_out ^= 1;
        tc_fence_after();
        if (part == 0) {
          uint32_t v[16];
          tmem_ld16(t_lane + 256u, v);
          tmem_ld_wait();
          const int64_t sample = (quad * 4 + 2 * s + (int64_t)rank) * kTileM + row;
          if (sample < n_samples) {
            float4 r;
            r.x = __uint_as_float(v[0]); r.y = __uint_as_float(v[1]);
            r.z = __uint_as_float(v[2]); r.w = __uint_as_float(v[3]);
            reinterpret_cast<float4*>(raw)[sample] = r;
          }
        }
        tc_fence_before();
        if (s == 0) {
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(bar_c_done);         // set Y's first layer may overwrite the accumulators
        }
      }
      // next pass (its arrives also tell the issuer that this pass's accumulators are drained)
      const int64_t next = quad + n_clusters;
      x[0] = xn[0]; x[1] = xn[1]; x[2] = xn[2];
      if (next < n_quads) open_pass(next, x);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, kTmemCols);
  }
}

int launch_field_ts2(const TcPlan& plan, const uint8_t* packed, int64_t n_samples, const int64_t* n_samples_dev, const float* x_skel,
                     const float* hann_w, int ignore_non_rigid, float* raw, float* xyz, float* offsets, int sms, cudaStream_t st) {
  using namespace ts2;
  if (plan.ts_chunks.empty() || plan.width[0] != 128 || plan.width[1] != 256)
    return fail_arg("hn_field_fwd: the two-tile TS kernel needs MLP widths 128 / 256");
  static int max_clusters_cached = 0;
  if (max_clusters_cached == 0) {
    HN_TRY(check_cuda(cudaFuncSetAttribute(field_ts2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem2T), "hn_field_fwd/ts2 smem attr"));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * (sms / 2)));
    cfg.blockDim = dim3(kThreads2T);
    cfg.dynamicSmemBytes = kSmem2T;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    HN_TRY(check_cuda(cudaOccupancyMaxActiveClusters(&n, field_ts2_kernel, &cfg), "hn_field_fwd/ts2 occupancy"));
    if (n <= 0) return fail_arg("hn_field_fwd: no 2-CTA cluster of the two-tile TS kernel fits on this device");
    max_clusters_cached = n;
  }
  TsTab tab;
  int nr_chunks = 0;
  fill_ts_tab(plan, &tab, &nr_chunks);
  const int64_t n_quads = ceil_div<int64_t>(n_samples, 4 * kTileM);
  const unsigned grid = 2u * (unsigned)std::min<int64_t>(n_quads, (int64_t)max_clusters_cached);
  field_ts2_kernel<<<grid, kThreads2T, kSmem2T, st>>>(tab, (int)plan.ts_chunks.size(), nr_chunks, packed, n_samples, n_samples_dev, x_skel,
                                                 hann_w, ignore_non_rigid, raw, xyz, offsets);
  return launched("hn_field_fwd/tcgen05 ts2");
}

}  // namespace hn
