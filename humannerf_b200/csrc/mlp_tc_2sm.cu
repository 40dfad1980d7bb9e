// Canonical field, tensor-core path (HN_PRECISION_BF16), default forward kernel: CTA PAIRS with
// tcgen05.mma.cta_group::2.
//
// One persistent kernel runs  hann-PE -> non-rigid MLP -> xyz = x_skel + offset -> PE -> canonical MLP -> raw  with
// every activation held on chip:
//   * bf16 operands, fp32 accumulation in TMEM; the A operand (activations) lives in shared memory in the UMMA
//     no-swizzle K-major core-matrix layout and is rewritten in place by 16 epilogue warps
//     (TMEM -> registers -> ReLU -> bf16 -> st.shared);
//   * biases ride in the GEMM: the positional-encoding block of A carries a constant-1 column and the weight image
//     the bias in the matching K row, so the epilogue has no loads;
//   * the B operand (weights) is pre-tiled by hn_field_pack_bf16 into exactly the shared-memory image the MMA wants
//     and streamed with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx) through a ring - no tensor maps.
// The predecessor (round 1, one CTA per SM, cta_group::1, two tiles sharing a 4 x 16 KB ring) ran its tiles in
// lock-step (39.8 % pipe active).  Here
//
//   * a cluster of two CTAs (two SMs) works on FOUR 128-sample tiles: tile set X = {tile 0 of CTA 0, tile 0 of
//     CTA 1}, tile set Y = {tile 1 of CTA 0, tile 1 of CTA 1}.  One M = 256 MMA covers a tile set: each SM's tensor
//     core multiplies ITS 128 rows of A (its own shared memory) with the whole B operand, of which each SM holds
//     HALF (rows [0, N/2) in CTA 0, [N/2, N) in CTA 1, same offsets) - the pair exchanges the halves in hardware;
//   * each CTA therefore streams only its half of every chunk (hn_field_pack_bf16 stores forward chunks as two
//     half images): the same 64 KB of ring per SM now hold 8 stages = 256 K-rows = a WHOLE layer, the L2 -> SM
//     traffic per sample halves, and tile set X can run a layer ahead of Y: X's epilogue and hand-over overlap
//     Y's MMAs and vice versa;
//   * CTA 0 (the leader) issues all MMAs (warp 1: set X into TMEM columns [0,256), warp 18: set Y into [256,512)
//     of BOTH CTAs' tensor memory) and commits with .multicast::cluster, so `empty` (ring slot free) and
//     `acc_full` (accumulator complete) fire in both CTAs;
//   * what the issuers wait for lives in the leader: `a_ready[t]` collects one mbarrier.arrive.release.cluster per
//     epilogue warp of BOTH CTAs (remote for the peer), `full[s]` collects the leader's own bulk copy
//     (complete_tx) and a forwarded arrive for the peer's half: a 1-D cp.async.bulk can only complete on a barrier
//     of the CTA it writes to (scripts/micro/umma_2sm_proto.cu: naming the leader's barrier hangs), so the peer's
//     copy completes on its own `floc[s]` and a relay thread (the peer's idle issuer warp) forwards it.
//
// Protocol pieces were checked in isolation by scripts/micro/umma_2sm.cu and umma_2sm_proto.cu.
// Reference semantics: embedders/hannw_fourier.py:38-49, embedders/fourier.py:9-38,
// non_rigid_motion_mlps/mlp_offset.py:74-114, canonical_mlps/mlp_rgb_sigma.py:132-206.
#include <cstdlib>

#include "mlp_tc_plan.cuh"
#include "mlp_tc_enc.cuh"
#include "mlp_tc_2sm.cuh"

namespace hn {

// kStore: training (activation / mask images are written); kSt: image store mode of store_blocks (1 or 4)
template <bool kStore, int kSt>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
field_tc2_kernel(const __grid_constant__ IssueTab itab, const uint8_t* __restrict__ packed, int64_t n_cap, const int64_t* __restrict__ n_dev,
                 const float* __restrict__ x_skel, const float* __restrict__ hann_w, int ignore_non_rigid,
                 float* __restrict__ raw, float* __restrict__ xyz_out, float* __restrict__ off_out, const TrainBufs bufs,
                 const __grid_constant__ ImageMaps maps, int n_stages) {
  using namespace s2;
  const uint32_t stage_bytes = kRingBytes2 / (uint32_t)n_stages;
  // n_dev (nullable): the number of samples actually present, read from device memory (empty-space skipping
  // compacts the samples without telling the host); n_cap bounds it and sizes the buffers
  const int64_t n_samples = n_dev ? min(max(__ldg(n_dev), (int64_t)0), n_cap) : n_cap;
  extern __shared__ __align__(1024) uint8_t smem[];
  const TcHeader* hdr = reinterpret_cast<const TcHeader*>(packed);
  const uint8_t* image = packed + hdr->image_off;

  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_act_pe = s_base + kActHBytes;
  const uint32_t s_ring = s_base + kRingOff;
  const uint32_t s_full = s_base + kBarOff2;                 // leader: own bulk copy + forwarded arrive of the peer
  const uint32_t s_floc = s_full + kStages2 * 8;             // peer: its own bulk copy
  const uint32_t s_empty = s_floc + kStages2 * 8;            // both: multicast commits of the two issuers
  const uint32_t s_acc_full = s_empty + kStages2 * 8;        // both: multicast commit at the end of a layer
  const uint32_t s_a_ready = s_acc_full + kTiles * 8;        // leader: epilogue warps of both CTAs
  const uint32_t s_tmem_slot = s_a_ready + kTiles * 8;
  const uint32_t s_lead = s_tmem_slot + 8;                   // leader: tile set X has issued the first chunk of a layer
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kBarOff2 + (3 * kStages2 + 2 * kTiles) * 8);
  const TcChunk* table = reinterpret_cast<const TcChunk*>(smem + kTableOff2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int64_t n_quads = (n_samples + 2 * kSetRows - 1) / (2 * kSetRows);       // 4 tiles = 512 samples per pass
  const int64_t n_clusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;
  const int nr_layers = (int)hdr->nr_layers;
  const int n_layers = nr_layers + (int)hdr->cn_layers;
  const int first_layer = ignore_non_rigid ? nr_layers : 0;
  const uint32_t n_chunks = hdr->n_chunks;

  // chunk table -> shared memory (every role reads it once per chunk)
  {
    const uint4* src = reinterpret_cast<const uint4*>(packed + hdr->table_off);
    uint4* dst = reinterpret_cast<uint4*>(smem + kTableOff2);
    for (uint32_t i = threadIdx.x; i < n_chunks; i += blockDim.x) dst[i] = src[i];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages2; ++s) {
      mbar_init(s_full + s * 8, 2);
      mbar_init(s_floc + s * 8, 1);
      mbar_init(s_empty + s * 8, kTiles);                    // one commit per MMA issuer (tile set)
    }
    for (int t = 0; t < kTiles; ++t) {
      mbar_init(s_acc_full + t * 8, 1);
      mbar_init(s_a_ready + t * 8, 2 * (kEpiWarps / kTiles));   // one arrive per epilogue warp of the tile set
    }
    mbar_init(s_lead, 1);
    mbar_init(s_lead + 8, 1);                                  // ... and Y has seen it (X never runs two layers ahead)
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
    // ===================== TMA producer: this CTA's half of every chunk =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t s_done = leader ? s_full : s_floc;
      for (int64_t quad = cluster_id; quad < n_quads; quad += n_clusters) {
        for (uint32_t c = first_chunk; c < n_chunks; ++c) {
          const TcChunk ch = table[c];
          const uint32_t half = ch.bytes >> 1;
          mbar_wait(s_empty + stage * 8, phase ^ 1, 1);
          mbar_expect_tx(s_done + stage * 8, half);
          tma_bulk_g2s(s_ring + stage * stage_bytes, image + ch.w_off + rank * half, half, s_done + stage * 8);
          if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kMmaWarp0 || warp == kMmaWarp1) {
    if (leader) {
      // ===================== MMA issuers (leader): warp 1 drives tile set X, warp 18 tile set Y =====================
      // the whole warp runs this loop convergently; one elected lane issues the tcgen05 instructions
      const uint32_t t = (warp == kMmaWarp0) ? 0u : 1u;
      const uint32_t d_tmem = tmem_base + t * 256u;
      const uint32_t bar_ready = s_a_ready + t * 8, bar_acc = s_acc_full + t * 8;
      // descriptor halves: hi = SBO (128 B >> 4) | version 1 (bit 46 -> bit 14 of hi); lo = addr >> 4 | LBO >> 4 << 16
      const uint32_t desc_hi = (128u >> 4) | (1u << 14);
      const uint32_t h_lo = (((s_base + t * kActHTile) & 0x3FFFFu) >> 4) | ((kLboA >> 4) << 16);
      const uint32_t pe_lo = (((s_act_pe + t * kActPeTile) & 0x3FFFFu) >> 4) | ((kLboA >> 4) << 16);
      const uint32_t ring_lo = (s_ring & 0x3FFFFu) >> 4;
      const uint32_t a_inc = (2u * kLboA) >> 4;                                      // one K step = two core matrices
      uint32_t stage = 0, phase = 0, a_phase = 0, full_ok = 0;
      // tile set X leads Y by one chunk (measured: +3.5 %; 2 / 3 / 4 chunks: +1.6 / -1.2 / +1.9 %: the 4-stage ring
      // then stalls X on refills)
      constexpr uint32_t lead_chunks = 1u;
      constexpr bool lead = true;
      uint32_t in_layer = 0;
      uint32_t n_layers_issued = 0;
      for (int64_t quad = cluster_id; quad < n_quads; quad += n_clusters) {
        for (uint32_t c = first_chunk; c < n_chunks; ++c) {
          const uint2 e = itab.e[c];
          const uint32_t n = e.y & 0xFFFFu, ksteps = (e.y >> 16) & 0xFFu, flags = e.y >> 24;
          in_layer = (flags & 1u) ? 1u : in_layer + 1u;
          if (flags & 1u) {
            // the A operand of this layer is complete in both CTAs (and the previous accumulator drained)
            mbar_wait(bar_ready, a_phase, 2);
            // tile set Y starts every layer one chunk behind X (lead / ack mbarrier pair; X never runs two layers
            // ahead): the two sets then leave the MMA phase staggered instead of neck and neck
            if (lead && t == 1u) {
              mbar_wait(s_lead, a_phase, 6);
              if (elect_one()) mbar_arrive(s_lead + 8);
              __syncwarp();
            }
            if (lead && t == 0u && n_layers_issued > 0) mbar_wait(s_lead + 8, a_phase ^ 1u, 7);
            ++n_layers_issued;
            a_phase ^= 1;
          }
          if (!full_ok) mbar_wait(s_full + stage * 8, phase, 3);
          tc_fence_after();
          // probe the NEXT chunk's weights now: the round trip overlaps the MMA issue below
          {
            const uint32_t ns = stage + 1 == (uint32_t)n_stages ? 0u : stage + 1;
            full_ok = mbar_test_wait(s_full + ns * 8, ns == 0u ? phase ^ 1u : phase);
          }
          const uint32_t idesc = make_idesc2(n);
          uint32_t a_lo = ((flags & 4u) ? pe_lo : h_lo) + (e.x >> 4);
          uint32_t b_lo = (ring_lo + stage * (stage_bytes >> 4)) | ((n >> 1) << 16);   // LBO = (n/2)*16 bytes
          if (elect_one()) {
            uint32_t acc = (flags & 1u) ? 0u : 1u;
            for (uint32_t k = 0; k < ksteps; ++k) {
              umma2_bf16(d_tmem, ((uint64_t)desc_hi << 32) | a_lo, ((uint64_t)desc_hi << 32) | b_lo, idesc, acc);
              acc = 1u;
              a_lo += a_inc;
              b_lo += n;                                   // two core-matrix groups of n/2 rows x 16 B, >> 4
            }
            umma2_commit_mc(s_empty + stage * 8);          // the slot is free in both CTAs once both issuers' MMAs retire
            if (flags & 2u) umma2_commit_mc(bar_acc);      // accumulator of the layer is complete (both CTAs)
            if (lead && t == 0u && (in_layer == lead_chunks || ((flags & 2u) && in_layer < lead_chunks))) mbar_arrive(s_lead);
          }
          __syncwarp();
          if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (lane == 0 && warp == kMmaWarp0) {
      // ===================== relay (peer): forwards "my half of the chunk has landed" to the leader =====================
      const uint32_t full_remote = mapa_cluster(s_full, 0);
      uint32_t stage = 0, phase = 0;
      for (int64_t quad = cluster_id; quad < n_quads; quad += n_clusters) {
        for (uint32_t c = first_chunk; c < n_chunks; ++c) {
          mbar_wait(s_floc + stage * 8, phase, 5);
          mbar_arrive_remote(full_remote + stage * 8);
          if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue warps (2 .. 17) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int tile = ew >> 3;                     // tile set: 0 = X, 1 = Y
    const int half = (ew >> 2) & 1;               // which half of the layer's columns / PE chunks
    const int row = quarter * 32 + lane;          // row inside the tile == TMEM lane
    const uint32_t t_acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)tile * 256u;
    const uint32_t a_h = s_base + (uint32_t)tile * kActHTile;
    const uint32_t a_pe = s_act_pe + (uint32_t)tile * kActPeTile;
    const uint32_t bar_ready = mapa_cluster(s_a_ready + (uint32_t)tile * 8, 0);      // in the leader
    const uint32_t bar_acc = s_acc_full + (uint32_t)tile * 8;
    uint32_t acc_phase = 0;
    constexpr int st_mode = kSt;                                     // image store mode (store_blocks)
    const uint32_t st_bar = 1u + (uint32_t)(tile * 2 + half);        // named barrier of this (tile, column half)
    float hw[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) hw[k] = (k < (int)hdr->nr_freqs && hann_w) ? __ldg(hann_w + k) : 0.0f;
    const int nr_freqs = (int)hdr->nr_freqs, cn_freqs = (int)hdr->cn_freqs;
    const int nr_pad = (int)hdr->nr_pe_pad, cn_pad = (int)hdr->cn_pe_pad;

    float xn[3] = {0.f, 0.f, 0.f};                // positions of the coming pass
    {
      const int64_t sample0 = (cluster_id * 4 + tile * 2 + (int64_t)rank) * kTileM + row;
      if (cluster_id < n_quads && sample0 < n_samples) {
        xn[0] = __ldg(x_skel + sample0 * 3 + 0); xn[1] = __ldg(x_skel + sample0 * 3 + 1); xn[2] = __ldg(x_skel + sample0 * 3 + 2);
      }
    }
    for (int64_t quad = cluster_id; quad < n_quads; quad += n_clusters) {
      // tile set X = tiles 4q, 4q+1 (CTA 0, CTA 1); tile set Y = tiles 4q+2, 4q+3
      const int64_t tile_idx = quad * 4 + tile * 2 + (int64_t)rank;
      const int64_t sample = tile_idx * kTileM + row;
      const bool valid = sample < n_samples;
      const bool store = kStore && tile_idx < bufs.n_tiles;      // the training images hold ceil(P/128) tiles
      const float x[3] = {xn[0], xn[1], xn[2]};
      float q[3] = {x[0], x[1], x[2]};            // position fed to the canonical encoding
      {
        // the next pass's positions are fetched now: ~1 us of HBM latency would otherwise open every pass
        const int64_t sample_n = ((quad + n_clusters) * 4 + tile * 2 + (int64_t)rank) * kTileM + row;
        const bool valid_n = quad + n_clusters < n_quads && sample_n < n_samples;
        xn[0] = valid_n ? __ldg(x_skel + sample_n * 3 + 0) : 0.f;
        xn[1] = valid_n ? __ldg(x_skel + sample_n * 3 + 1) : 0.f;
        xn[2] = valid_n ? __ldg(x_skel + sample_n * 3 + 2) : 0.f;
      }

      if (!ignore_non_rigid) {
        if (kStore) block_reusable(st_mode, st_bar);           // the block's previous image copy has left shared memory
        write_encoding<false>(a_pe, row, half, x, nr_freqs, hw, nr_pad, nullptr);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(bar_ready);
        if (store) store_blocks(st_mode, &maps.m[2], -1, st_bar, bufs.pe[0], tile_idx, nr_pad, a_pe, quarter, lane, half, 2, nr_pad >> 4);
      }

      for (int layer = first_layer; layer < n_layers; ++layer) {
        const bool is_nr = layer < nr_layers;
        const bool is_out = (layer == nr_layers - 1) || (layer == n_layers - 1);
        if (layer == nr_layers) {
          if (kStore) block_reusable(st_mode, st_bar);
          write_encoding<true>(a_pe, row, half, q, cn_freqs, nullptr, cn_pad, nullptr);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(bar_ready);
          if (store) store_blocks(st_mode, &maps.m[3], -1, st_bar, bufs.pe[1], tile_idx, cn_pad, a_pe, quarter, lane, half, 2, cn_pad >> 4);
        }
        // training: the image copies of the block this layer overwrites have left shared memory - checked (and, with
        // tensor-map stores, agreed among the four warps) while the layer's MMAs run, not after them
        if (kStore && !is_out) block_reusable(st_mode, st_bar);
        mbar_wait(bar_acc, acc_phase, 4);
        acc_phase ^= 1;
        tc_fence_after();
        if (!is_out) {
          const int n_full = (int)hdr->layer_n[layer];
          const int n_half = n_full >> 1;                          // columns owned by this thread
          const int c_begin = half * n_half;
          // training: this thread's record of the layer's 1-bit ReLU mask image (the backward chain masks with it)
          uint8_t* bits_rec = nullptr;
          if (kStore && store) {
            uint8_t* b = is_nr ? bufs.bits[0][layer] : bufs.bits[1][layer - nr_layers];
            if (b) bits_rec = b + relu_bits_offset(tile_idx, half, row, n_half);
          }
          uint32_t v[2][32];
          tmem_ld32(t_acc + (uint32_t)c_begin, v[0]);
#pragma unroll 1
          for (int c0 = 0; c0 < n_half; c0 += 64) {
            // two 32-column groups per iteration so that the register buffers are indexed statically
            uint32_t m0 = 0u, m1 = 0u;
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
              if (kStore) m0 = relu_bits_step(relu_bits_step(relu_bits_step(relu_bits_step(m0, pk.x), pk.y), pk.z), pk.w);
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
                if (kStore) m1 = relu_bits_step(relu_bits_step(relu_bits_step(relu_bits_step(m1, pk.x), pk.y), pk.z), pk.w);
              }
            }
            if (kStore && bits_rec) *reinterpret_cast<uint2*>(bits_rec + (c0 >> 6) * 8) = make_uint2(m0, m1);
          }
          tc_fence_before();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(bar_ready);
          if (store)
            store_blocks(st_mode, &maps.m[is_nr ? 0 : 1], 2 * (is_nr ? layer : layer - nr_layers), st_bar,
                         is_nr ? bufs.h[0][layer] : bufs.h[1][layer - nr_layers], tile_idx, n_full, a_h, quarter, lane,
                         half * (n_half >> 3), 1, n_half >> 3);
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
    if (kStore) bulk_wait0();                       // image copies complete before the CTA gives up its shared memory
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // the peer may still arrive on / commit into this CTA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, kTmemCols);
  }
}

// Launch: persistent, exactly as many clusters as can be co-resident (GPC boundaries may leave a few SMs without
// a partner, so this can be below sms / 2).  Returns <0 / cudaError_t like every entry point.
int launch_field_tc2(const TcPlan& plan, const uint8_t* packed, int64_t n_samples, const int64_t* n_samples_dev, const float* x_skel,
                     const float* hann_w, int ignore_non_rigid, float* raw, float* xyz, float* offsets,
                     const TrainBufs* bufs, int sms, cudaStream_t st) {
  using namespace s2;
  static int max_clusters_cached = 0;
  if (max_clusters_cached == 0) {
    HN_TRY(check_cuda(cudaFuncSetAttribute(field_tc2_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTotal2),
                      "hn_field_fwd/2sm smem attr"));
    HN_TRY(check_cuda(cudaFuncSetAttribute(field_tc2_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTotal2),
                      "hn_field_fwd/2sm smem attr"));
    HN_TRY(check_cuda(cudaFuncSetAttribute(field_tc2_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTotal2),
                      "hn_field_fwd/2sm smem attr"));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * (sms / 2)));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemTotal2;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    HN_TRY(check_cuda(cudaOccupancyMaxActiveClusters(&n, field_tc2_kernel<false, 1>, &cfg), "hn_field_fwd/2sm occupancy"));
    if (n <= 0) return fail_arg("hn_field_fwd: no 2-CTA cluster of the tcgen05 kernel fits on this device");
    max_clusters_cached = n;
    if (getenv("HN_DEBUG")) fprintf(stderr, "humannerf_b200: %d co-resident 2-CTA clusters (%d SMs)\n", n, sms);
  }
  IssueTab itab;
  memset(&itab, 0, sizeof(itab));
  for (size_t c = 0; c < plan.chunks.size(); ++c) {
    const TcChunk& ch = plan.chunks[c];
    itab.e[c] = make_uint2(ch.a_off, (uint32_t)ch.n | ((uint32_t)ch.ksteps << 16) | ((uint32_t)ch.flags << 24));
  }
  const int64_t n_quads = ceil_div<int64_t>(n_samples, 2 * kSetRows);
  const unsigned grid = 2u * (unsigned)std::min<int64_t>(n_quads, (int64_t)max_clusters_cached);
  ImageMaps maps;
  memset(&maps, 0, sizeof(maps));
  if (bufs && bufs->leader_stores == 4) {
    HN_TRY(make_image_maps(plan, *bufs, &maps));
    field_tc2_kernel<true, 4><<<grid, kThreads, kSmemTotal2, st>>>(itab, packed, n_samples, n_samples_dev, x_skel, hann_w,
                                                                  ignore_non_rigid, raw, xyz, offsets, *bufs, maps,
                                                                  tc2_stages());
  } else if (bufs) {
    field_tc2_kernel<true, 1><<<grid, kThreads, kSmemTotal2, st>>>(itab, packed, n_samples, n_samples_dev, x_skel, hann_w,
                                                                  ignore_non_rigid, raw, xyz, offsets, *bufs, maps,
                                                                  tc2_stages());
  } else {
    TrainBufs none;
    memset(&none, 0, sizeof(none));
    field_tc2_kernel<false, 1><<<grid, kThreads, kSmemTotal2, st>>>(itab, packed, n_samples, n_samples_dev, x_skel, hann_w,
                                                                   ignore_non_rigid, raw, xyz, offsets, none, maps,
                                                                   tc2_stages());
  }
  return launched("hn_field_fwd/tcgen05 2sm");
}

}  // namespace hn
