// Canonical field, tensor-core backward-DATA chain on CTA pairs (tcgen05.mma.cta_group::2): the default chain
// kernel of hn_field_bwd_bf16.  Same math and epilogues as field_tc_bwd_kernel (mlp_tc_bwd.cu, read its header
// first), same pair structure as the forward kernel of mlp_tc_2sm.cu: a cluster of two CTAs works on four
// 128-sample tiles, one M = 256 MMA covers a tile set (dZ rows of both CTAs against the W^T chunk of which each CTA
// streams half), the leader CTA's two issuer warps run warp-uniform with the chunk schedule in kernel parameters,
// and what the issuers wait for (`a_ready`, `full`) collects remote arrives from the peer.
//   * dZ of the previous layer is written in place as the next A operand; its T16 image for the weight-gradient
//     kernel leaves shared memory through the TMA (store_blocks in mlp_tc_2sm.cuh: one 2 KB tensor-map copy per
//     8-column chunk of a tile by default, kSt = 4; 256-byte bulk copies with kSt = 1) instead of a 16-byte store per
//     thread and chunk;
//   * ReLU': kBits = true (default) masks the accumulator with the 1-bit mask image the forward kernel wrote
//     (relu_bits_* in mlp_tc_plan.cuh): one 4..16-byte record per thread and round, fetched a round ahead, instead
//     of re-reading the bf16 activations (16 bytes per thread and 8-column chunk, 4.4 GB per 786 k samples);
//     kBits = false (HN_TC_RELU_BITS=0) keeps the activation read.
// Autograd of reference mlp_offset.py:74-114 and mlp_rgb_sigma.py:132-206 on the bf16 path.
#include "mlp_tc_plan.cuh"
#include "mlp_tc_enc.cuh"
#include "mlp_tc_2sm.cuh"

namespace hn {

template <bool kBits, int kSt>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
field_tc2_bwd_kernel(const __grid_constant__ IssueTab itab, const uint8_t* __restrict__ packed, int64_t n_cap,
                     const int64_t* __restrict__ n_dev, int mlp, const float* __restrict__ d_out, int ld_out, int n_out,
                     const TrainBufs bufs, const __grid_constant__ ImageMaps maps, const float* __restrict__ pos,
                     const float* __restrict__ hann_w, float* __restrict__ d_pos, int ld_dpos, int n_stages) {
  using namespace s2;
  const int64_t n_samples = n_dev ? min(max(__ldg(n_dev), (int64_t)0), n_cap) : n_cap;   // see field_tc2_kernel
  const uint32_t stage_bytes = kRingBytes2 / (uint32_t)n_stages;
  extern __shared__ __align__(1024) uint8_t smem[];
  const TcHeader* hdr = reinterpret_cast<const TcHeader*>(packed);
  const uint8_t* image = packed + hdr->image_off;

  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_ring = s_base + kRingOff;
  const uint32_t s_full = s_base + kBarOff2;                 // leader: own bulk copy + forwarded arrive of the peer
  const uint32_t s_floc = s_full + kStages2 * 8;             // peer: its own bulk copy
  const uint32_t s_empty = s_floc + kStages2 * 8;            // both: multicast commits of the two issuers
  const uint32_t s_acc_full = s_empty + kStages2 * 8;        // both: multicast commit at the end of a round
  const uint32_t s_a_ready = s_acc_full + kTiles * 8;        // leader: epilogue warps of both CTAs
  const uint32_t s_tmem_slot = s_a_ready + kTiles * 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kBarOff2 + (3 * kStages2 + 2 * kTiles) * 8);
  const TcChunk* table = reinterpret_cast<const TcChunk*>(smem + kTableOff2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int64_t n_quads = (n_samples + 2 * kSetRows - 1) / (2 * kSetRows);
  const int64_t n_clusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;
  const uint32_t c_first = hdr->bwd_chunk0[mlp], c_last = hdr->bwd_chunk1[mlp];
  const int r_first = (int)hdr->bwd_round0[mlp], r_last = (int)hdr->bwd_round1[mlp];
  const int out_layer = (int)(mlp == 0 ? hdr->nr_layers : hdr->cn_layers) - 1;

  {
    const uint4* src = reinterpret_cast<const uint4*>(packed + hdr->bwd_table_off);
    uint4* dst = reinterpret_cast<uint4*>(smem + kTableOff2);
    for (uint32_t i = threadIdx.x; i < hdr->bwd_n_chunks; i += blockDim.x) dst[i] = src[i];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages2; ++s) {
      mbar_init(s_full + s * 8, 2);
      mbar_init(s_floc + s * 8, 1);
      mbar_init(s_empty + s * 8, kTiles);
    }
    for (int t = 0; t < kTiles; ++t) {
      mbar_init(s_acc_full + t * 8, 1);
      mbar_init(s_a_ready + t * 8, 2 * (kEpiWarps / kTiles));   // one arrive per epilogue warp of the tile set
    }
    fence_barrier_init();
  }
  if (warp == 1) {            // the same warp of both CTAs allocates collectively
    tmem_alloc2(s_tmem_slot, kTmemCols);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: this CTA's half of every W^T chunk =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t s_done = leader ? s_full : s_floc;
      for (int64_t quad = cluster_id; quad < n_quads; quad += n_clusters) {
        for (uint32_t c = c_first; c < c_last; ++c) {
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
      // ===================== MMA issuers (leader), warp-uniform: warp 1 tile set X, warp 18 tile set Y =====================
      const uint32_t t = (warp == kMmaWarp0) ? 0u : 1u;
      const uint32_t d_tmem = tmem_base + t * 256u;
      const uint32_t bar_ready = s_a_ready + t * 8, bar_acc = s_acc_full + t * 8;
      const uint32_t desc_hi = (128u >> 4) | (1u << 14);
      const uint32_t h_lo = (((s_base + t * kActHTile) & 0x3FFFFu) >> 4) | ((kLboA >> 4) << 16);
      const uint32_t ring_lo = (s_ring & 0x3FFFFu) >> 4;
      const uint32_t a_inc = (2u * kLboA) >> 4;
      uint32_t stage = 0, phase = 0, a_phase = 0, full_ok = 0;
      for (int64_t quad = cluster_id; quad < n_quads; quad += n_clusters) {
        for (uint32_t c = c_first; c < c_last; ++c) {
          const uint2 e = itab.e[c];
          const uint32_t n = e.y & 0xFFFFu, ksteps = (e.y >> 16) & 0xFFu, flags = e.y >> 24;
          if (flags & 1u) {
            mbar_wait(bar_ready, a_phase, 2);
            a_phase ^= 1;
          }
          if (!full_ok) mbar_wait(s_full + stage * 8, phase, 3);
          tc_fence_after();
          {
            const uint32_t ns = stage + 1 == (uint32_t)n_stages ? 0u : stage + 1;
            full_ok = mbar_test_wait(s_full + ns * 8, ns == 0u ? phase ^ 1u : phase);
          }
          const uint32_t idesc = make_idesc2(n);
          uint32_t a_lo = h_lo + (e.x >> 4);
          uint32_t b_lo = (ring_lo + stage * (stage_bytes >> 4)) | ((n >> 1) << 16);
          if (elect_one()) {
            uint32_t acc = (flags & 1u) ? 0u : 1u;
            for (uint32_t k = 0; k < ksteps; ++k) {
              umma2_bf16(d_tmem, ((uint64_t)desc_hi << 32) | a_lo, ((uint64_t)desc_hi << 32) | b_lo, idesc, acc);
              acc = 1u;
              a_lo += a_inc;
              b_lo += n;
            }
            umma2_commit_mc(s_empty + stage * 8);
            if (flags & 2u) umma2_commit_mc(bar_acc);
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
        for (uint32_t c = c_first; c < c_last; ++c) {
          mbar_wait(s_floc + stage * 8, phase, 5);
          mbar_arrive_remote(full_remote + stage * 8);
          if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int tile = ew >> 3;
    const int half = (ew >> 2) & 1;
    const int row = quarter * 32 + lane;
    const uint32_t t_acc = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)tile * 256u;
    const uint32_t a_h = s_base + (uint32_t)tile * kActHTile;
    const uint32_t bar_ready = mapa_cluster(s_a_ready + (uint32_t)tile * 8, 0);      // in the leader
    const uint32_t bar_acc = s_acc_full + (uint32_t)tile * 8;
    uint32_t acc_phase = 0;
    constexpr int st_mode = kSt;                                     // image store mode (store_blocks: 1 or 4)
    const uint32_t st_bar = 1u + (uint32_t)(tile * 2 + half);        // named barrier of this (tile, column half)

    for (int64_t quad = cluster_id; quad < n_quads; quad += n_clusters) {
      const int64_t tile_idx = quad * 4 + tile * 2 + (int64_t)rank;
      const int64_t sample = tile_idx * kTileM + row;
      const bool valid = sample < n_samples;
      const bool have_img = tile_idx < bufs.n_tiles;

      // ---- dZ of the output layer: 16 columns (n_out valid), chunk `half` ----
      {
        float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (valid && half == 0)
          for (int c = 0; c < n_out && c < 8; ++c) g[c] = d_out[sample * ld_out + c];
        const uint4 pk = make_uint4(pack_bf16(g[0], g[1]), pack_bf16(g[2], g[3]), pack_bf16(g[4], g[5]), pack_bf16(g[6], g[7]));
        block_reusable(st_mode, st_bar);                         // image copies of the previous pass have left the block
        st_shared_v4(a_h + (uint32_t)half * kLboA + (uint32_t)row * 16u, pk.x, pk.y, pk.z, pk.w);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(bar_ready);
        if (have_img) store_blocks(st_mode, &maps.m[4 + mlp], -1, st_bar, bufs.dz[mlp][out_layer], tile_idx, 16, a_h, quarter, lane, half, 1, 1);
      }

      // The stored activations that mask the accumulators come from HBM (~1 us): the two 16-row groups of this
      // warp's rows of the NEXT masking round are pulled into L2 one round ahead (one 8 KB bulk prefetch per group)
      auto prefetch_mask_rows = [&](int r_from) {
        if (!have_img || half != 0 || lane >= 2) return;
        for (int rr = r_from; rr < r_last; ++rr) {
          const BwdRound nx = hdr->rounds[rr];
          if (nx.epi != 0) continue;
          const int F = nx.n;
          bulk_prefetch_l2(bufs.h[mlp][nx.layer] + (size_t)tile_idx * 128 * F * 2 + (size_t)(2 * quarter + lane) * F * 32,
                           (uint32_t)F * 32u);
          return;
        }
      };
      // kBits: this thread's record of a masking round's 1-bit ReLU mask image (one word per 32 columns).  It is
      // pulled into L2 a round ahead (no registers held) and loaded before the wait for the round's accumulator.
      uint32_t mb[4] = {0u, 0u, 0u, 0u};
      auto bits_record = [&](int r_from) -> const uint32_t* {
        if (!have_img) return nullptr;
        for (int rr = r_from; rr < r_last; ++rr) {
          const BwdRound nx = hdr->rounds[rr];
          if (nx.epi != 0) continue;
          return reinterpret_cast<const uint32_t*>(bufs.bits[mlp][nx.layer] + relu_bits_offset(tile_idx, half, row, nx.n >> 1));
        }
        return nullptr;
      };
      auto prefetch_bits = [&](int r_from) {
        const uint32_t* rec = bits_record(r_from);
        if (rec) asm volatile("prefetch.global.L2 [%0];" ::"l"(rec));
      };
      if (kBits) prefetch_bits(r_first);
      else prefetch_mask_rows(r_first);
      for (int r = r_first; r < r_last; ++r) {
        const BwdRound rd = hdr->rounds[r];
        const bool last = r == r_last - 1;
        if (!kBits && rd.epi == 0) prefetch_mask_rows(r + 1);
        if (kBits && rd.epi == 0) {
          const int n_full = rd.n, n_half = n_full >> 1, c_begin = half * n_half;
          {
            const uint32_t* rec = bits_record(r);
#pragma unroll
            for (int i = 0; i < 4; ++i) mb[i] = (rec && i * 32 < n_half) ? __ldg(rec + i) : 0u;
          }
          prefetch_bits(r + 1);
          block_reusable(st_mode, st_bar);                       // the block's previous image copy has left shared memory
          mbar_wait(bar_acc, acc_phase, 4);                      // (checked while the round's MMAs run)
          acc_phase ^= 1;
          tc_fence_after();
          uint32_t v[2][32];
          tmem_ld32(t_acc + (uint32_t)c_begin, v[0]);
#pragma unroll 1
          for (int c0 = 0; c0 < n_half; c0 += 64) {
            // two 32-column groups per iteration (static register indexing); mb[0] / mb[1] are their mask words
            tmem_ld_wait();
            if (c0 + 32 < n_half) tmem_ld32(t_acc + (uint32_t)(c_begin + c0 + 32), v[1]);
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              const uint32_t* u = v[0] + g8 * 8;
              const int chunk = ((c_begin + c0) >> 3) + g8;
              st_shared_v4(a_h + (uint32_t)chunk * kLboA + (uint32_t)row * 16u,
                           mask_pack_bits(mb[0], g8 * 4 + 0, u[0], u[1]), mask_pack_bits(mb[0], g8 * 4 + 1, u[2], u[3]),
                           mask_pack_bits(mb[0], g8 * 4 + 2, u[4], u[5]), mask_pack_bits(mb[0], g8 * 4 + 3, u[6], u[7]));
            }
            if (c0 + 32 < n_half) {
              tmem_ld_wait();
              if (c0 + 64 < n_half) tmem_ld32(t_acc + (uint32_t)(c_begin + c0 + 64), v[0]);
#pragma unroll
              for (int g8 = 0; g8 < 4; ++g8) {
                const uint32_t* u = v[1] + g8 * 8;
                const int chunk = ((c_begin + c0 + 32) >> 3) + g8;
                st_shared_v4(a_h + (uint32_t)chunk * kLboA + (uint32_t)row * 16u,
                             mask_pack_bits(mb[1], g8 * 4 + 0, u[0], u[1]), mask_pack_bits(mb[1], g8 * 4 + 1, u[2], u[3]),
                             mask_pack_bits(mb[1], g8 * 4 + 2, u[4], u[5]), mask_pack_bits(mb[1], g8 * 4 + 3, u[6], u[7]));
              }
            }
            mb[0] = mb[2]; mb[1] = mb[3];
          }
          tc_fence_before();
          fence_proxy_async();
          __syncwarp();
          if (!last && lane == 0) mbar_arrive_remote(bar_ready);
          if (have_img)
            store_blocks(st_mode, &maps.m[mlp], 2 * (int)rd.layer + 1, st_bar, bufs.dz[mlp][rd.layer], tile_idx, n_full, a_h,
                         quarter, lane, half * (n_half >> 3), 1, n_half >> 3);
        } else if (rd.epi == 0) {
          const int n_full = rd.n, n_half = n_full >> 1, c_begin = half * n_half;
          const size_t row_off = image_row_offset(tile_idx, row, n_full);
          const uint8_t* h_img = have_img ? bufs.h[mlp][rd.layer] + row_off : nullptr;
          uint4 hq[4], hn[4];
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8)
            hq[g8] = h_img ? __ldg(reinterpret_cast<const uint4*>(h_img + (size_t)((c_begin >> 3) + g8) * 256))
                           : make_uint4(0u, 0u, 0u, 0u);
          block_reusable(st_mode, st_bar);                       // the block's previous image copy has left shared memory
          mbar_wait(bar_acc, acc_phase, 4);
          acc_phase ^= 1;
          tc_fence_after();
#pragma unroll 1
          for (int c0 = 0; c0 < n_half; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(t_acc + (uint32_t)(c_begin + c0), v);
            if (c0 + 32 < n_half) {
#pragma unroll
              for (int g8 = 0; g8 < 4; ++g8)
                hn[g8] = h_img ? __ldg(reinterpret_cast<const uint4*>(h_img + (size_t)(((c_begin + c0 + 32) >> 3) + g8) * 256))
                               : make_uint4(0u, 0u, 0u, 0u);
            }
            tmem_ld_wait();
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              const uint32_t* u = v + g8 * 8;
              const int chunk = ((c_begin + c0) >> 3) + g8;
              const uint4 pk = make_uint4(mask_pack(hq[g8].x, u[0], u[1]), mask_pack(hq[g8].y, u[2], u[3]),
                                          mask_pack(hq[g8].z, u[4], u[5]), mask_pack(hq[g8].w, u[6], u[7]));
              st_shared_v4(a_h + (uint32_t)chunk * kLboA + (uint32_t)row * 16u, pk.x, pk.y, pk.z, pk.w);
            }
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) hq[g8] = hn[g8];
          }
          tc_fence_before();
          fence_proxy_async();
          __syncwarp();
          if (!last && lane == 0) mbar_arrive_remote(bar_ready);
          if (have_img)
            store_blocks(st_mode, &maps.m[mlp], 2 * (int)rd.layer + 1, st_bar, bufs.dz[mlp][rd.layer], tile_idx, n_full, a_h,
                         quarter, lane, half * (n_half >> 3), 1, n_half >> 3);
        } else {
          // gradient of the encoding block, contracted with the encoding's derivative:
          //   d_pos[a] += [raw x] g[a] + sum_k w_k 2^k (cos(2^k x_a) g_sin(k,a) - sin(2^k x_a) g_cos(k,a))
          // (embedders/fourier.py:22-25 column order: per frequency sin(x,y,z) then cos(x,y,z)).
          // The two threads of a row take alternate 16-column groups and add their partial sums atomically.
          mbar_wait(bar_acc, acc_phase, 4);
          acc_phase ^= 1;
          tc_fence_after();
          const int n_full = rd.n;
          const int n_freqs = (int)(mlp == 0 ? hdr->nr_freqs : hdr->cn_freqs);
          const int raw_cols = mlp == 0 ? 0 : 3;                  // the canonical encoding starts with x itself
          const int pe_dim = raw_cols + 6 * n_freqs;
          float px = 0.f, py = 0.f, pz = 0.f;
          if (valid) { px = pos[sample * 3 + 0]; py = pos[sample * 3 + 1]; pz = pos[sample * 3 + 2]; }
          float ax = 0.f, ay = 0.f, az = 0.f;
          for (int c0 = half * 16; c0 < n_full; c0 += 32) {
            uint32_t v[16];
            tmem_ld16(t_acc + (uint32_t)c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int c = c0 + j;
              if (c >= pe_dim) continue;                          // zero padding and the constant-1 column
              const float g = __uint_as_float(v[j]);
              if (c < raw_cols) {
                if (c == 0) ax += g; else if (c == 1) ay += g; else az += g;
                continue;
              }
              const int cc = c - raw_cols;
              const int k = cc / 6, rr = cc - 6 * k;
              const int a = rr >= 3 ? rr - 3 : rr;
              const float freq = (float)(1 << k);
              const float w = (hann_w ? __ldg(hann_w + k) : 1.0f) * freq;
              const float xa = a == 0 ? px : (a == 1 ? py : pz);
              float sn, cs;
              __sincosf(xa * freq, &sn, &cs);
              const float tt = w * g * (rr >= 3 ? -sn : cs);
              if (a == 0) ax += tt; else if (a == 1) ay += tt; else az += tt;
            }
          }
          if (valid) {
            atomicAdd(d_pos + sample * ld_dpos + 0, ax);
            atomicAdd(d_pos + sample * ld_dpos + 1, ay);
            atomicAdd(d_pos + sample * ld_dpos + 2, az);
          }
          tc_fence_before();
          __syncwarp();
          if (!last && lane == 0) mbar_arrive_remote(bar_ready);
        }
      }
    }
    bulk_wait0();                                   // image copies complete before the CTA gives up its shared memory
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // the peer may still arrive on / commit into this CTA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, kTmemCols);
  }
}

int launch_field_tc2_bwd(const TcPlan& plan, const uint8_t* packed, int64_t n_samples, const int64_t* n_samples_dev, int mlp,
                         const float* d_out, int ld_out, int n_out, const TrainBufs& bufs, const float* pos,
                         const float* hann_w, float* d_pos, int ld_dpos, int sms, cudaStream_t st) {
  using namespace s2;
  static int max_clusters_cached = 0;
  if (max_clusters_cached == 0) {
    const void* variants[4] = {(const void*)field_tc2_bwd_kernel<true, 4>, (const void*)field_tc2_bwd_kernel<true, 1>,
                               (const void*)field_tc2_bwd_kernel<false, 4>, (const void*)field_tc2_bwd_kernel<false, 1>};
    for (const void* f : variants)
      HN_TRY(check_cuda(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTotal2),
                        "hn_field_bwd_bf16/2sm smem attr"));
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
    HN_TRY(check_cuda(cudaOccupancyMaxActiveClusters(&n, field_tc2_bwd_kernel<false, 1>, &cfg), "hn_field_bwd_bf16/2sm occupancy"));
    if (n <= 0) return fail_arg("hn_field_bwd_bf16: no 2-CTA cluster of the tcgen05 kernel fits on this device");
    max_clusters_cached = n;
  }
  IssueTab itab;
  memset(&itab, 0, sizeof(itab));
  for (size_t c = 0; c < plan.bwd_chunks.size(); ++c) {
    const TcChunk& ch = plan.bwd_chunks[c];
    itab.e[c] = make_uint2(ch.a_off, (uint32_t)ch.n | ((uint32_t)ch.ksteps << 16) | ((uint32_t)ch.flags << 24));
  }
  const int64_t n_quads = ceil_div<int64_t>(n_samples, 2 * kSetRows);
  const unsigned grid = 2u * (unsigned)std::min<int64_t>(n_quads, (int64_t)max_clusters_cached);
  ImageMaps maps;
  memset(&maps, 0, sizeof(maps));
  if (bufs.leader_stores == 4) HN_TRY(make_image_maps(plan, bufs, &maps));
  // the 1-bit mask images exist when the cta_group::2 forward wrote them (hn_field_bwd_bf16 nulls the pointers otherwise)
  const bool bits = bufs.bits[mlp][0] != nullptr, maps4 = bufs.leader_stores == 4;
#define HN_BWD2_LAUNCH(B, S)                                                                                              \
  field_tc2_bwd_kernel<B, S><<<grid, kThreads, kSmemTotal2, st>>>(itab, packed, n_samples, n_samples_dev, mlp, d_out, ld_out, \
                                                                 n_out, bufs, maps, pos, hann_w, d_pos, ld_dpos, tc2_stages())
  if (bits && maps4) HN_BWD2_LAUNCH(true, 4);
  else if (bits) HN_BWD2_LAUNCH(true, 1);
  else if (maps4) HN_BWD2_LAUNCH(false, 4);
  else HN_BWD2_LAUNCH(false, 1);
#undef HN_BWD2_LAUNCH
  return launched(mlp == 1 ? "hn_field_bwd_bf16/canonical chain (2sm)" : "hn_field_bwd_bf16/non-rigid chain (2sm)");
}

}  // namespace hn
