// Canonical field, tensor-core backward (HN_PRECISION_BF16).
//
//   1. field_tc_bwd_kernel: the backward-DATA chain of one MLP, fused like the forward kernel.  Per
//      pair of 128-sample tiles, dZ of the current layer sits in shared memory as the A operand (bf16,
//      UMMA K-major core-matrix layout), W^T chunks stream through the TMA ring, the accumulator
//      dH = dZ W lands in TMEM, and the epilogue masks it with the stored activation of the previous
//      layer (T16 image written by the forward kernel) to get dZ of that layer - written in place as
//      the next A operand AND to a T16 image in HBM for the weight-gradient kernel.  The gradient of
//      the positional-encoding columns is contracted with the encoding's derivative right in the
//      epilogue (d_pos += PE'(pos) . d_pe, two threads per row, fp32 atomics), so the (P, pe_pad)
//      encoding gradients never reach HBM,
//   2. hn_tile_wgrad jobs produce every dW and db from the dZ / activation images,
//   3. the condition-code columns of the first non-rigid layer get dW0[:, :69] += db0 (x) cond.
//
// Autograd of reference mlp_offset.py:74-114 and mlp_rgb_sigma.py:132-206 on the bf16 path.
#include "mlp_tc_plan.cuh"
#include "mlp_tc_enc.cuh"

namespace hn {

// field_fp32.cu
int add3(int64_t n, const float* a, int64_t lda, const float* b, int64_t ldb, float* out, int64_t ldo,
         cudaStream_t st);
int copy_cols(int64_t n, int n_cols, const float* src, int64_t src_ld, float* dst, int64_t dst_ld,
              cudaStream_t st);
// mlp_tc_bwd_2sm.cu
int launch_field_tc2_bwd(const TcPlan& plan, const uint8_t* packed, int64_t n_samples, const int64_t* n_samples_dev, int mlp,
                         const float* d_out, int ld_out, int n_out, const TrainBufs& bufs, const float* pos,
                         const float* hann_w, float* d_pos, int ld_dpos, int sms, cudaStream_t st);

__device__ __forceinline__ size_t image_row_offset_b(int64_t tile, int r, int F) {
  return (size_t)tile * 128 * F * 2 + (size_t)(r >> 4) * 16 * F * 2 + (size_t)(r & 15) * 16;
}

__global__ void __launch_bounds__(kThreads, 1)
field_tc_bwd_kernel(const uint8_t* __restrict__ packed, int64_t n_cap, const int64_t* __restrict__ n_dev, int mlp,
                    const float* __restrict__ d_out,
                    int ld_out, int n_out, const TrainBufs bufs, const float* __restrict__ pos,
                    const float* __restrict__ hann_w, float* __restrict__ d_pos, int ld_dpos) {
  const int64_t n_samples = n_dev ? min(max(__ldg(n_dev), (int64_t)0), n_cap) : n_cap;   // see field_tc_kernel
  extern __shared__ __align__(1024) uint8_t smem[];
  const TcHeader* hdr = reinterpret_cast<const TcHeader*>(packed);
  const uint8_t* image = packed + hdr->image_off;

  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_ring = s_base + kRingOff;
  const uint32_t s_full = s_base + kBarOff;
  const uint32_t s_empty = s_full + kStages * 8;
  const uint32_t s_acc_full = s_empty + kStages * 8;
  const uint32_t s_a_ready = s_acc_full + kTiles * 8;
  const uint32_t s_tmem_slot = s_a_ready + kTiles * 8;
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(smem + kBarOff + (2 * kStages + 2 * kTiles) * 8);
  const TcChunk* table = reinterpret_cast<const TcChunk*>(smem + kTableOff);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_pairs = (n_samples + kTiles * kTileM - 1) / (kTiles * kTileM);
  const uint32_t c_first = hdr->bwd_chunk0[mlp], c_last = hdr->bwd_chunk1[mlp];
  const int r_first = (int)hdr->bwd_round0[mlp], r_last = (int)hdr->bwd_round1[mlp];
  const int out_layer = (int)(mlp == 0 ? hdr->nr_layers : hdr->cn_layers) - 1;

  {
    const uint4* src = reinterpret_cast<const uint4*>(packed + hdr->bwd_table_off);
    uint4* dst = reinterpret_cast<uint4*>(smem + kTableOff);
    for (uint32_t i = threadIdx.x; i < hdr->bwd_n_chunks; i += blockDim.x) dst[i] = src[i];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(s_full + s * 8, 1);
      mbar_init(s_empty + s * 8, kTiles);
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

  if (warp == 0) {
    // ===================== TMA producer: W^T chunks =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        for (uint32_t c = c_first; c < c_last; ++c) {
          const TcChunk ch = table[c];
          mbar_wait(s_empty + stage * 8, phase ^ 1, 1);
          mbar_expect_tx(s_full + stage * 8, ch.bytes);
          tma_bulk_g2s(s_ring + stage * kStageBytes, image + ch.w_off, ch.bytes, s_full + stage * 8);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kMmaWarp0 || warp == kMmaWarp1) {
    // ===================== MMA issuers (one per tile) =====================
    if (lane == 0) {
      const uint32_t t = (warp == kMmaWarp0) ? 0u : 1u;
      const uint32_t d_tmem = tmem_base + t * 256u;
      const uint32_t bar_ready = s_a_ready + t * 8, bar_acc = s_acc_full + t * 8;
      const uint32_t desc_hi = (128u >> 4) | (1u << 14);
      const uint32_t a_lo_const = ((kLboA >> 4) << 16);
      uint32_t stage = 0, phase = 0, a_phase = 0;
      // the NEXT chunk's table entry is fetched before this chunk's waits (a dependent shared-memory round
      // trip of the issuing thread costs ~120 clk while MMAs are in flight, scripts/micro/umma_rate.cu)
      const uint32_t s_table = smem_u32(table);
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        TcChunk ch_next = table[c_first];
        for (uint32_t c = c_first; c < c_last; ++c) {
          const TcChunk ch = ch_next;
          {
            const uint32_t cn = c + 1 < c_last ? c + 1 : c;
            uint32_t w0, w1, w2, w3;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(s_table + cn * 16u));
            ch_next.w_off = w0; ch_next.bytes = w1; ch_next.a_off = w2;
            ch_next.n = (uint16_t)(w3 & 0xFFFFu); ch_next.ksteps = (uint8_t)((w3 >> 16) & 0xFFu); ch_next.flags = (uint8_t)(w3 >> 24);
          }
          if (ch.flags & 1) {
            mbar_wait(bar_ready, a_phase, 2);
            a_phase ^= 1;
            tc_fence_after();
          }
          mbar_wait(s_full + stage * 8, phase, 3);
          tc_fence_after();
          const uint32_t idesc = make_idesc(ch.n);
          const uint32_t a_addr = s_base + t * kActHTile + ch.a_off;
          const uint32_t b_addr = s_ring + stage * kStageBytes;
          uint32_t a_lo = ((a_addr & 0x3FFFFu) >> 4) | a_lo_const;
          uint32_t b_lo = ((b_addr & 0x3FFFFu) >> 4) | ((uint32_t)ch.n << 16);
          const uint32_t a_inc = (2u * kLboA) >> 4, b_inc = 2u * (uint32_t)ch.n;
          uint32_t acc = (ch.flags & 1) ? 0u : 1u;
#pragma unroll 2
          for (uint32_t k = 0; k < ch.ksteps; ++k) {
            umma_bf16(d_tmem, ((uint64_t)desc_hi << 32) | a_lo, ((uint64_t)desc_hi << 32) | b_lo, idesc, acc);
            acc = 1u;
            a_lo += a_inc;
            b_lo += b_inc;
          }
          umma_commit(s_empty + stage * 8);
          if (ch.flags & 2) umma_commit(bar_acc);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
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
    const uint32_t bar_ready = s_a_ready + (uint32_t)tile * 8, bar_acc = s_acc_full + (uint32_t)tile * 8;
    uint32_t acc_phase = 0;

    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const int64_t tile_idx = pair * kTiles + tile;
      const int64_t sample = tile_idx * kTileM + row;
      const bool valid = sample < n_samples;
      const bool have_img = tile_idx < bufs.n_tiles;

      // ---- dZ of the output layer: 16 columns (n_out valid), chunk `half` ----
      {
        float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (valid && half == 0)
          for (int c = 0; c < n_out && c < 8; ++c) g[c] = d_out[sample * ld_out + c];
        const uint4 pk = make_uint4(pack_bf16(g[0], g[1]), pack_bf16(g[2], g[3]), pack_bf16(g[4], g[5]), pack_bf16(g[6], g[7]));
        st_shared_v4(a_h + (uint32_t)half * kLboA + (uint32_t)row * 16u, pk.x, pk.y, pk.z, pk.w);
        if (have_img)
          *reinterpret_cast<uint4*>(bufs.dz[mlp][out_layer] + image_row_offset_b(tile_idx, row, 16) + (size_t)half * 256) = pk;
        fence_proxy_async();
        warp_arrive(bar_ready, lane);
      }

      for (int r = r_first; r < r_last; ++r) {
        const BwdRound rd = hdr->rounds[r];
        const bool last = r == r_last - 1;
        if (rd.epi == 0) {
          const int n_full = rd.n, n_half = n_full >> 1, c_begin = half * n_half;
          const size_t row_off = image_row_offset_b(tile_idx, row, n_full);
          const uint8_t* h_img = have_img ? bufs.h[mlp][rd.layer] + row_off : nullptr;
          uint8_t* dz_img = have_img ? bufs.dz[mlp][rd.layer] + row_off : nullptr;
          uint4 hq[4], hn[4];
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8)
            hq[g8] = h_img ? __ldg(reinterpret_cast<const uint4*>(h_img + (size_t)((c_begin >> 3) + g8) * 256))
                           : make_uint4(0u, 0u, 0u, 0u);
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
              if (dz_img) *reinterpret_cast<uint4*>(dz_img + (size_t)chunk * 256) = pk;
            }
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) hq[g8] = hn[g8];
          }
          tc_fence_before();
          fence_proxy_async();
          if (!last) warp_arrive(bar_ready, lane);
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
              const int k = cc / 6, r = cc - 6 * k;
              const int a = r >= 3 ? r - 3 : r;
              const float freq = (float)(1 << k);
              const float w = (hann_w ? __ldg(hann_w + k) : 1.0f) * freq;
              const float xa = a == 0 ? px : (a == 1 ? py : pz);
              float sn, cs;
              __sincosf(xa * freq, &sn, &cs);
              const float t = w * g * (r >= 3 ? -sn : cs);
              if (a == 0) ax += t; else if (a == 1) ay += t; else az += t;
            }
          }
          if (valid) {
            atomicAdd(d_pos + sample * ld_dpos + 0, ax);
            atomicAdd(d_pos + sample * ld_dpos + 1, ay);
            atomicAdd(d_pos + sample * ld_dpos + 2, az);
          }
          tc_fence_before();
          if (!last) warp_arrive(bar_ready, lane);
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

// dW[r, c] += a[r] * b[c]
__global__ void outer_add_kernel(float* __restrict__ dw, int ld, int rows, int cols, const float* __restrict__ a,
                                 const float* __restrict__ b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int r = i / cols, c = i - r * cols;
  dw[(int64_t)r * ld + c] += a[r] * b[c];
}

static int64_t bwd_scratch_bytes(int64_t n) {
  const int64_t rows = ceil_div<int64_t>(std::max<int64_t>(n, 1), 128) * 128;
  return 2 * rows * 4 * (int64_t)sizeof(float) + 1024;
}

}  // namespace hn

using namespace hn;

extern "C" {

int64_t hn_field_bwd_bf16_scratch_bytes(const hn_field_params* p, int64_t n_samples) {
  TcPlan plan;
  if (make_tc_plan(p, plan) != 0) return -1;
  return bwd_scratch_bytes(n_samples);
}

int hn_field_bwd_bf16(const hn_field_params* p, const hn_field_grads* g, int64_t n_samples, const float* x_skel,
                      const float* xyz, const float* cond, const float* hann_w, int ignore_non_rigid,
                      const float* d_raw, const float* d_xyz, const float* d_offsets, float* d_x_skel,
                      void* workspace, int64_t workspace_bytes, const void* packed, void* scratch,
                      int64_t scratch_bytes, const int64_t* n_samples_dev, void* stream) {
  TcPlan plan;
  HN_TRY(make_tc_plan(p, plan));
  if (n_samples == 0) return 0;
  if (!g || n_samples < 0 || !x_skel || !xyz || !d_raw || !d_x_skel || !workspace || !packed || !scratch)
    return fail_arg("hn_field_bwd_bf16: null pointer or negative size");
  if (!ignore_non_rigid && ((p->cond_dim > 0 && !cond) || !hann_w)) return fail_arg("hn_field_bwd_bf16: cond / hann_w missing");
  TrainBufs bufs;
  const int64_t need = carve_train_bufs(plan, n_samples, static_cast<uint8_t*>(workspace), &bufs, true);
  if (workspace_bytes < need) return fail_arg("hn_field_bwd_bf16: workspace is not a bf16 training workspace of this size");
  if (scratch_bytes < bwd_scratch_bytes(n_samples)) return fail_arg("hn_field_bwd_bf16: scratch too small");
  cudaStream_t st = as_stream(stream);
  const int64_t n = n_samples;
  const int64_t rows = ceil_div<int64_t>(n, 128) * 128;
  float* d_xyz_tot = static_cast<float*>(scratch);             // (n, 4)
  float* dz_nr = d_xyz_tot + rows * 4;                         // (n, 4)
  float* dummy = d_xyz_tot;                                    // never written (n_valid = 0 jobs)

  int dev = 0, sms = 0;
  HN_TRY(check_cuda(cudaGetDevice(&dev), "hn_field_bwd_bf16/device"));
  HN_TRY(check_cuda(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "hn_field_bwd_bf16/attr"));
  HN_TRY(check_cuda(cudaFuncSetAttribute(field_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kSmemTotal), "hn_field_bwd_bf16/smem attr"));
  const int64_t n_pairs = ceil_div<int64_t>(n, kTiles * kTileM);
  const unsigned grid = (unsigned)std::min<int64_t>(n_pairs, (int64_t)sms);
  const uint8_t* pk = static_cast<const uint8_t*>(packed);

  // 1. canonical MLP: d_raw -> dZ images; d_xyz_tot = upstream d_xyz + PE'(xyz) . d(encoding), accumulated
  //    by the chain's epilogue (both encoding reads: skip layer and layer 0)
  if (d_xyz) HN_TRY(copy_cols(n, 3, d_xyz, 3, d_xyz_tot, 4, st));
  else HN_TRY(check_cuda(cudaMemsetAsync(d_xyz_tot, 0, (size_t)rows * 4 * sizeof(float), st), "hn_field_bwd_bf16/memset"));
  // default: the cta_group::2 chain kernel (mlp_tc_bwd_2sm.cu); HN_TC_KERNEL=pair / cluster: the tile-pair kernel above
  const bool two_sm = use_2sm_kernel();
  if (two_sm && sms < 2) return fail_arg("hn_field_bwd_bf16: the tcgen05 chain kernel needs a device with thread-block clusters");
  // only the cta_group::2 forward writes the 1-bit ReLU mask images (and only unless HN_TC_RELU_BITS=0)
  if (!two_sm || !use_relu_bits()) memset(bufs.bits, 0, sizeof(bufs.bits));
  if (two_sm) {
    HN_TRY(launch_field_tc2_bwd(plan, pk, n, n_samples_dev, 1, d_raw, 4, 4, bufs, xyz, nullptr, d_xyz_tot, 4, sms, st));
  } else {
    field_tc_bwd_kernel<<<grid, kThreads, kSmemTotal, st>>>(pk, n, n_samples_dev, 1, d_raw, 4, 4, bufs, xyz, nullptr, d_xyz_tot, 4);
    HN_TRY(launched("hn_field_bwd_bf16/canonical chain"));
  }
  // 2. non-rigid MLP: its output gradient is d_xyz_tot + d_offsets; d_x_skel = d_xyz_tot + PE'(x_skel) . d(encoding)
  HN_TRY(copy_cols(n, 3, d_xyz_tot, 4, d_x_skel, 3, st));
  if (!ignore_non_rigid) {
    HN_TRY(add3(n, d_xyz_tot, 4, d_offsets, 3, dz_nr, 4, st));
    if (two_sm) {
      HN_TRY(launch_field_tc2_bwd(plan, pk, n, n_samples_dev, 0, dz_nr, 4, 3, bufs, x_skel, hann_w, d_x_skel, 3, sms, st));
    } else {
      field_tc_bwd_kernel<<<grid, kThreads, kSmemTotal, st>>>(pk, n, n_samples_dev, 0, dz_nr, 4, 3, bufs, x_skel, hann_w, d_x_skel, 3);
      HN_TRY(launched("hn_field_bwd_bf16/non-rigid chain"));
    }
  }
  // 3. weight / bias gradients from the images
  std::vector<WgradJob> jobs;
  auto add_job = [&](const uint8_t* dz, int f_dz, int m, const uint8_t* x, int f_x, int x_f0, int nn, int n_valid,
                     float* dw, int dw_ld, int dw_col0, int bias_col, float* db) {
    if (!dw && !db) return;
    WgradJob j;
    memset(&j, 0, sizeof(j));
    j.dz = dz; j.x = x; j.dw = dw ? dw : dummy; j.db = db; j.f_dz = f_dz; j.f_x = f_x; j.m = m; j.x_f0 = x_f0;
    j.n = nn; j.n_valid = dw ? n_valid : 0; j.dw_ld = dw_ld; j.dw_col0 = dw_col0; j.bias_col = db ? bias_col : -1;
    jobs.push_back(j);
  };
  for (int m = ignore_non_rigid ? 1 : 0; m < 2; ++m) {
    float* const* gW = m == 0 ? g->nr_w : g->cn_w;
    float* const* gB = m == 0 ? g->nr_b : g->cn_b;
    const int L = plan.layers[m], width = plan.width[m], skip = plan.skip[m], pe = plan.pe_dim[m];
    const int pe_pad = plan.pe_pad[m], lead = plan.lead[m];
    const bool pe_first = m == 1;
    for (int l = 0; l < L; ++l) {
      const bool out = l == L - 1;
      const int f_dz = out ? 16 : width, mrows = out ? plan.out_dim[m] : width;
      const int in_dim = (l == 0) ? lead + pe : (l == skip ? width + pe : width);
      const uint8_t* dz = bufs.dz[m][l];
      if (l == 0) {
        add_job(dz, f_dz, mrows, bufs.pe[m], pe_pad, 0, pe_pad, pe, gW[l], in_dim, lead, pe_pad - 1, gB[l]);
      } else if (l == skip) {
        add_job(dz, f_dz, mrows, bufs.h[m][l - 1], width, 0, width, width, gW[l], in_dim, pe_first ? pe : 0, -1, nullptr);
        add_job(dz, f_dz, mrows, bufs.pe[m], pe_pad, 0, pe_pad, pe, gW[l], in_dim, pe_first ? 0 : width, pe_pad - 1, gB[l]);
      } else if (gW[l]) {
        // weights and bias in one job: the bias gradient is the column sum of the dZ stages (bias_col = -2)
        add_job(dz, f_dz, mrows, bufs.h[m][l - 1], width, 0, width, width, gW[l], in_dim, 0, gB[l] ? -2 : -1, gB[l]);
      } else {
        add_job(dz, f_dz, mrows, bufs.pe[m], pe_pad, pe_pad - 16, 16, 0, nullptr, 16, 0, 15, gB[l]);
      }
    }
  }
  if ((int)jobs.size() > kMaxWgradJobs) return fail_arg("hn_field_bwd_bf16: too many weight-gradient jobs");
  if (!jobs.empty()) HN_TRY(launch_tile_wgrad(jobs.data(), (int)jobs.size(), n, n_samples_dev, st));
  // 4. condition-code columns of the first non-rigid layer: the input is the same vector for every sample
  if (!ignore_non_rigid && p->cond_dim > 0 && g->nr_w[0] && g->nr_b[0]) {
    const int rows_w = p->nr_width, cols = p->cond_dim;
    outer_add_kernel<<<ceil_div(rows_w * cols, 256), 256, 0, st>>>(g->nr_w[0], p->cond_dim + plan.pe_dim[0], rows_w, cols,
                                                                  g->nr_b[0], cond);
    HN_TRY(launched("hn_field_bwd_bf16/cond outer"));
  }
  return 0;
}

}  // extern "C"
