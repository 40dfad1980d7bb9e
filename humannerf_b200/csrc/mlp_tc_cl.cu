// Canonical field forward, EXPERIMENTAL cluster variant of the tcgen05 kernel (mlp_tc.cu), selected with
// HN_TC_KERNEL=cluster.  It passes the same parity tests as the tile-pair kernel but is slower (78 ms vs
// 42 ms per 512x512 frame): with ONE tile per CTA nothing hides the accumulator-complete -> epilogue ->
// operand-ready hand-over of each layer (~2000 clk, measured with the HN_TC_PROFILE counters), which the
// tile-pair kernel overlaps with the other tile's MMAs.  Kept because it carries the cluster machinery
// (multicast weight stream, cross-CTA ring release, block-granular operand hand-over, pre-decoded issuer
// table) that the round-2 kernel (two tiles per CTA pair with cta_group::2) builds on - DESIGN.md 4.1.
//
// Why a second structure: in the tile-pair kernel the two tiles of a CTA share every weight chunk and
// the 64 KB ring holds only half a layer, so the tiles run in lock-step and their epilogues (TMEM read
// port: 2048 clk per 128x256 fp32 accumulator) leave the tensor pipe idle.  Here
//
//   * a CLUSTER of two CTAs shares the weight stream: each CTA fetches half of every chunk and
//     TMA-multicasts it into both CTAs' rings (cp.async.bulk ... .multicast::cluster), so the L2 -> SM
//     traffic per SM is that of the tile-pair kernel while each CTA works on ONE 128-sample tile;
//   * with one tile per CTA the activations take 80 KB and the ring grows to 9 x 16 KB = 288 K-rows -
//     more than a whole layer - so the two CTAs of a cluster (and the producer) never stall each other
//     within a layer; ring slots are released by tcgen05.commit ... .multicast::cluster arriving on the
//     `empty` barrier of BOTH CTAs;
//   * the TMEM accumulator is double buffered (2 x 256 columns): while the 16 epilogue warps drain
//     layer l (TMEM -> registers -> ReLU -> bf16 -> shared memory) the MMAs of layer l+1 already run
//     into the other buffer;
//   * the operand hand-over is per 32-column block: four threads per row convert 8 columns each
//     (tcgen05.ld 32x32b.x8 -> one 16-byte st.shared), a warp-level arrive on the block's mbarrier
//     follows, and the issuer starts the K steps of layer l+1 that read this block (weight chunks
//     never span more than one block, mlp_tc_plan.cuh).  The MMA of layer l+1 therefore trails the
//     epilogue of layer l by one block instead of one layer.
//
// Everything else (weight image, bias folded into the GEMM through the constant-1 column of the
// encoding block, T16 image stores for training) is as in mlp_tc.cu.
// Reference semantics: embedders/hannw_fourier.py:38-49, embedders/fourier.py:9-38,
// non_rigid_motion_mlps/mlp_offset.py:74-114, canonical_mlps/mlp_rgb_sigma.py:132-206.
#include <cstdlib>

#include "mlp_tc_plan.cuh"

namespace hn {

namespace cl {
constexpr int kStagesC = 9;
constexpr uint32_t kActH = kTileM * 256 * 2;                  // 64 KB
constexpr uint32_t kActPe = kTileM * 64 * 2;                  // 16 KB
constexpr uint32_t kRingOffC = kActH + kActPe;
constexpr uint32_t kBarOffC = kRingOffC + kStagesC * kStageBytes;
constexpr int kBlocks = 8;                                    // 32-column blocks of the hidden operand
// barriers: full[9] empty[9] acc_full[2] a_blk[8] pe_ready[1] + tmem slot
constexpr uint32_t kBarBytes = (2 * kStagesC + 2 + kBlocks + 1) * 8 + 8;
constexpr uint32_t kTableOffC = kBarOffC + 256;
constexpr uint32_t kSmemTotalC = kTableOffC + kMaxChunks * 16;
constexpr int kEpiWarpsC = 16;
constexpr int kThreadsC = (2 + kEpiWarpsC) * 32;              // producer, issuer, 16 epilogue warps
static_assert(kBarBytes <= 256, "barrier block");
static_assert(kSmemTotalC <= 232448, "shared memory budget");
}  // namespace cl

// half of a weight chunk into the same ring slot of BOTH CTAs of the cluster; completes `bytes` on the
// `full` barrier (same offset) of each destination CTA
__device__ __forceinline__ void tma_bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(mask)
      : "memory");
}
// arrive on the barrier at this offset in every CTA of `mask` once all previously issued MMAs retire
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
// D[tmem] (+)= A * B^T with the two 64-bit shared-memory descriptors given as (low, high) words
__device__ __forceinline__ void umma_bf16_hl(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "setp.ne.b32 p, %6, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

// value of column c of an encoding block: [optional x (3)], then per frequency k sin(2^k x_a) (a = 0..2)
// and cos(2^k x_a); zero padding; constant 1 in the last column of the block (the bias column)
template <bool kIncludeInput>
__device__ __forceinline__ float encoding_col(int c, const float x[3], int n_freqs, const float* hw, int pad) {
  if (c == pad - 1) return 1.0f;
  int cc = c;
  if (kIncludeInput) {
    if (c < 3) return c == 0 ? x[0] : (c == 1 ? x[1] : x[2]);
    cc = c - 3;
  }
  const int k = cc / 6;
  if (k >= n_freqs) return 0.0f;
  const int r = cc - 6 * k;
  const int a = r >= 3 ? r - 3 : r;
  const float xa = a == 0 ? x[0] : (a == 1 ? x[1] : x[2]);
  const float arg = xa * (float)(1 << k);
  const float v = r >= 3 ? __cosf(arg) : __sinf(arg);
  return hw ? hw[k] * v : v;
}

// the four threads of a row write the 16-byte chunks ch = sub, sub + 4, ... of the encoding block
template <bool kIncludeInput>
__device__ __forceinline__ void write_encoding_cl(uint32_t region, int row, int sub, const float x[3], int n_freqs,
                                                  const float* hw, int pad, uint8_t* img_row) {
  for (int ch = sub; ch * 8 < pad; ch += 4) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = encoding_col<kIncludeInput>(ch * 8 + i, x, n_freqs, hw, pad);
    const uint32_t a = pack_bf16(v[0], v[1]), b = pack_bf16(v[2], v[3]), c = pack_bf16(v[4], v[5]), d = pack_bf16(v[6], v[7]);
    st_shared_v4(region + (uint32_t)ch * kLboA + (uint32_t)row * 16u, a, b, c, d);
    if (img_row) *reinterpret_cast<uint4*>(img_row + (size_t)ch * 256) = make_uint4(a, b, c, d);
  }
}

__device__ __forceinline__ size_t image_row_offset_c(int64_t tile, int r, int F) {
  return (size_t)tile * 128 * F * 2 + (size_t)(r >> 4) * 16 * F * 2 + (size_t)(r & 15) * 16;
}

// cycle accounting of block 0 (HN_TC_PROFILE builds with HN_TC_ABLATE & 8):
//  [0..3] producer: total, wait empty;  [4..7] issuer: total, wait block/pe, wait full, issue;
//  [8..11] epilogue thread (warp 2 lane 0): total, wait acc, encoding, hidden blocks
__device__ unsigned long long g_tcc_prof[16];
#ifdef HN_TC_PROFILE
#define HC_T(var) const long long var = prof ? clock64() : 0
#define HC_ADD(slot, t0) do { if (prof) pacc[slot] += (unsigned long long)(clock64() - (t0)); } while (0)
#define HC_DECL unsigned long long pacc[4] = {0, 0, 0, 0}
#define HC_FLUSH(base) do { if (prof) { for (int _i = 0; _i < 4; ++_i) g_tcc_prof[(base) + _i] += pacc[_i]; } } while (0)
#else
#define HC_T(var) do { } while (0)
#define HC_ADD(slot, t0) do { } while (0)
#define HC_DECL do { } while (0)
#define HC_FLUSH(base) do { } while (0)
#endif

template <bool kStore>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(cl::kThreadsC, 1)
field_tcc_kernel(const uint8_t* __restrict__ packed, int64_t n_samples, const float* __restrict__ x_skel,
                 const float* __restrict__ hann_w, int ignore_non_rigid, float* __restrict__ raw,
                 float* __restrict__ xyz_out, float* __restrict__ off_out, const TrainBufs bufs, int dbg) {
  using namespace cl;
#ifdef HN_TC_PROFILE
  bool prof = (dbg & 8) && blockIdx.x == 0;
#endif
  extern __shared__ __align__(1024) uint8_t smem[];
  const TcHeader* hdr = reinterpret_cast<const TcHeader*>(packed);
  const uint8_t* image = packed + hdr->image_off;

  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_act_pe = s_base + kActH;
  const uint32_t s_ring = s_base + kRingOffC;
  const uint32_t s_full = s_base + kBarOffC;
  const uint32_t s_empty = s_full + kStagesC * 8;
  const uint32_t s_acc_full = s_empty + kStagesC * 8;          // 2 barriers
  const uint32_t s_a_blk = s_acc_full + 2 * 8;                 // kBlocks barriers
  const uint32_t s_pe_ready = s_a_blk + kBlocks * 8;
  const uint32_t s_tmem_slot = s_pe_ready + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kBarOffC + (2 * kStagesC + 2 + kBlocks + 1) * 8);
  uint2* p_tab = reinterpret_cast<uint2*>(smem + kTableOffC);                       // producer: (w_off, bytes)
  uint2* i_tab = reinterpret_cast<uint2*>(smem + kTableOffC + kMaxChunks * 8);      // issuer: (a_lo, control)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t n_tiles = (n_samples + kTileM - 1) / kTileM;
  const int64_t n_pairs = (n_tiles + 1) / 2;                   // one tile per CTA of the cluster
  const int64_t n_clusters = gridDim.x / 2, cluster_id = blockIdx.x / 2;
  const int nr_layers = (int)hdr->nr_layers;
  const int n_layers = nr_layers + (int)hdr->cn_layers;
  const int first_layer = ignore_non_rigid ? nr_layers : 0;
  const uint32_t n_chunks = hdr->n_chunks;

  // The chunk table is decoded once into two 8-byte-per-chunk tables in shared memory: what the producer
  // needs (offset, bytes) and what the issuer needs (the A descriptor's low word and a packed control
  // word), so that the issuer - a single thread whose dependent instruction chain sits on the critical
  // path of every MMA - does no address arithmetic per chunk.
  //   control word: bit0 first chunk of its layer, bit1 last, bit2 operand in the encoding block,
  //   bit3 wait for the encoding (first chunk of each MLP), bits 4-5 N code (0: 16, 1: 128, 2: 256, 3: other),
  //   bits 6-7 K steps, bits 8-10 operand block, bits 16-31 N
  {
    const TcChunk* src = reinterpret_cast<const TcChunk*>(packed + hdr->table_off);
    for (uint32_t c = threadIdx.x; c < n_chunks; c += blockDim.x) {
      const TcChunk ch = src[c];
      int layers_before = 0;
      for (uint32_t i = 0; i < c; ++i) layers_before += (src[i].flags >> 1) & 1;
      const bool in_pe = (ch.flags & 4) != 0;
      const bool wait_pe = (ch.flags & 1) && (layers_before == 0 || layers_before == nr_layers);
      const uint32_t a_addr = (in_pe ? s_act_pe : s_base) + ch.a_off;
      const uint32_t a_lo = ((a_addr & 0x3FFFFu) >> 4) | ((kLboA >> 4) << 16);
      const uint32_t ncode = ch.n == 16 ? 0u : (ch.n == 128 ? 1u : (ch.n == 256 ? 2u : 3u));
      const uint32_t ctrl = (uint32_t)(ch.flags & 7) | (wait_pe ? 8u : 0u) | (ncode << 4) | ((uint32_t)ch.ksteps << 6) |
                            (((ch.a_off / kLboA) >> 2) << 8) | ((uint32_t)ch.n << 16);
      p_tab[c] = make_uint2(ch.w_off, ch.bytes);
      i_tab[c] = make_uint2(a_lo, ctrl);
    }
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStagesC; ++s) {
      mbar_init(s_full + s * 8, 1);
      mbar_init(s_empty + s * 8, 2);                 // one multicast commit from each CTA's issuer
    }
    mbar_init(s_acc_full, 1);
    mbar_init(s_acc_full + 8, 1);
    for (int b = 0; b < kBlocks; ++b) mbar_init(s_a_blk + b * 8, kEpiWarpsC);
    mbar_init(s_pe_ready, kEpiWarpsC);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(s_tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                // barrier inits visible before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  uint32_t first_chunk = 0;
  if (ignore_non_rigid) {
    int seen = 0;
    for (uint32_t c = 0; c < n_chunks && seen < nr_layers; ++c) {
      if (i_tab[c].y & 2u) ++seen;
      first_chunk = c + 1;
    }
  }

  if (warp == 0) {
    // ===================== TMA producer: this CTA's half of every chunk, multicast to both =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      HC_DECL;
      HC_T(p0);
      for (int64_t pair = cluster_id; pair < n_pairs; pair += n_clusters) {
        for (uint32_t c = first_chunk; c < n_chunks; ++c) {
          const uint2 ch = p_tab[c];                              // (w_off, bytes)
          const uint32_t half = ch.y >> 1;
          HC_T(pw);
          mbar_wait(s_empty + stage * 8, phase ^ 1, 1);          // both CTAs have consumed the slot
          HC_ADD(1, pw);
          mbar_expect_tx(s_full + stage * 8, ch.y);              // my half + the peer's half
          if (dbg & 16)   // measurement aid: no multicast, every CTA fetches whole chunks
            tma_bulk_g2s(s_ring + stage * kStageBytes, image + ch.x, ch.y, s_full + stage * 8);
          else
            tma_bulk_g2s_mc(s_ring + stage * kStageBytes + rank * half, image + ch.x + rank * half, half,
                            s_full + stage * 8, (uint16_t)3);
          if (++stage == kStagesC) { stage = 0; phase ^= 1; }
        }
      }
      HC_ADD(0, p0);
      HC_FLUSH(0);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t desc_hi = (128u >> 4) | (1u << 14);
      const uint32_t idesc16 = make_idesc(16), idesc128 = make_idesc(128), idesc256 = make_idesc(256);
      const uint32_t a_inc = (2u * kLboA) >> 4;
      const uint32_t b_ring_lo = (s_ring & 0x3FFFFu) >> 4, b_stage_inc = kStageBytes >> 4;
      uint32_t stage = 0, phase = 0, pe_phase = 0, g = 0;       // g: running layer counter (accumulator buffer = g & 1)
      uint32_t b_stage_lo = b_ring_lo;
      uint32_t blk_phase = 0;                                    // bit b: parity to wait for on a_blk[b]
      HC_DECL;
      HC_T(m0);
      // Every dependent shared-memory round trip of this thread costs ~120 clk while MMAs are in flight
      // (scripts/micro/umma_rate.cu), as much as one MMA: the control word of the NEXT chunk is therefore
      // fetched before this chunk's waits, and the two barrier probes are issued back to back.
      const uint32_t s_itab = smem_u32(i_tab);
      for (int64_t pair = cluster_id; pair < n_pairs; pair += n_clusters) {
        uint2 e_next = i_tab[first_chunk];
        for (uint32_t c = first_chunk; c < n_chunks; ++c) {
          const uint2 e = e_next;
          {
            const uint32_t cn = c + 1 < n_chunks ? c + 1 : c;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e_next.x), "=r"(e_next.y) : "r"(s_itab + cn * 8u));
          }
          const uint32_t ctrl = e.y;
          HC_T(mw);
          uint32_t bar_a = 0, par_a = 0;
          if (ctrl & 4u) {
            if (ctrl & 8u) {                                       // first chunk of an MLP: its encoding block
              bar_a = s_pe_ready;
              par_a = pe_phase;
              pe_phase ^= 1;
            }
          } else {
            const uint32_t b = (ctrl >> 8) & 7u;                   // 32-column block this chunk reads
            bar_a = s_a_blk + b * 8;
            par_a = (blk_phase >> b) & 1u;
            blk_phase ^= 1u << b;
          }
          const uint32_t bar_f = s_full + stage * 8;
          const uint32_t ok_a = bar_a ? mbar_try_wait(bar_a, par_a) : 1u;
          const uint32_t ok_f = mbar_try_wait(bar_f, phase);
          if (!ok_a) mbar_wait(bar_a, par_a, 2);
          HC_ADD(1, mw);
          HC_T(mf);
          if (!ok_f) mbar_wait(bar_f, phase, 3);
          HC_ADD(2, mf);
          tc_fence_after();
          HC_T(mi);
          const uint32_t ncode = (ctrl >> 4) & 3u, n = ctrl >> 16;
          const uint32_t idesc = ncode == 2u ? idesc256 : (ncode == 1u ? idesc128 : (ncode == 0u ? idesc16 : make_idesc(n)));
          const uint32_t d_tmem = tmem_base + (g & 1u) * 256u;
          const uint32_t b_lo = b_stage_lo | (n << 16);
          umma_bf16_hl(d_tmem, e.x, desc_hi, b_lo, desc_hi, idesc, (ctrl & 1u) ^ 1u);
          if (ctrl & 0x80u) umma_bf16_hl(d_tmem, e.x + a_inc, desc_hi, b_lo + 2u * n, desc_hi, idesc, 1u);
          umma_commit_mc(s_empty + stage * 8, (uint16_t)3);       // frees the slot in BOTH CTAs
          if (ctrl & 2u) {
            umma_commit(s_acc_full + (g & 1u) * 8);
            ++g;
          }
          HC_ADD(3, mi);
          b_stage_lo += b_stage_inc;
          if (++stage == kStagesC) { stage = 0; phase ^= 1; b_stage_lo = b_ring_lo; }
        }
      }
      HC_ADD(0, m0);
      HC_FLUSH(4);
    }
  } else {
    // ===================== epilogue warps (2 .. 17): four threads per row =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int sub = ew >> 2;                      // which 8 columns of every 32-column block
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t acc_phase = 0;                       // bit i: parity to wait for on acc_full[i]
    uint32_t g = 0;
    float hw[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) hw[k] = (k < (int)hdr->nr_freqs && hann_w) ? __ldg(hann_w + k) : 0.0f;
    const int nr_freqs = (int)hdr->nr_freqs, cn_freqs = (int)hdr->cn_freqs;
    const int nr_pad = (int)hdr->nr_pe_pad, cn_pad = (int)hdr->cn_pe_pad;
#ifdef HN_TC_PROFILE
    prof = prof && warp == 2 && lane == 0;
#endif
    HC_DECL;
    HC_T(e0);

    for (int64_t pair = cluster_id; pair < n_pairs; pair += n_clusters) {
      const int64_t tile_idx = pair * 2 + rank;
      const int64_t sample = tile_idx * kTileM + row;
      const bool valid = sample < n_samples;
      const bool store = kStore && tile_idx < bufs.n_tiles;
      float x[3] = {0.f, 0.f, 0.f};
      if (valid) { x[0] = x_skel[sample * 3 + 0]; x[1] = x_skel[sample * 3 + 1]; x[2] = x_skel[sample * 3 + 2]; }
      float q[3] = {x[0], x[1], x[2]};

      if (!ignore_non_rigid) {
        write_encoding_cl<false>(s_act_pe, row, sub, x, nr_freqs, hw, nr_pad,
                                 store ? bufs.pe[0] + image_row_offset_c(tile_idx, row, nr_pad) : nullptr);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_pe_ready);
      }

      for (int layer = first_layer; layer < n_layers; ++layer) {
        const bool is_nr = layer < nr_layers;
        const bool is_out = (layer == nr_layers - 1) || (layer == n_layers - 1);
        if (layer == nr_layers) {
          write_encoding_cl<true>(s_act_pe, row, sub, q, cn_freqs, nullptr, cn_pad,
                                  store ? bufs.pe[1] + image_row_offset_c(tile_idx, row, cn_pad) : nullptr);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_pe_ready);
        }
        const uint32_t buf = g & 1u;
        HC_T(ew0);
        mbar_wait(s_acc_full + buf * 8, (acc_phase >> buf) & 1u, 4);
        HC_ADD(1, ew0);
        acc_phase ^= 1u << buf;
        ++g;
        tc_fence_after();
        const uint32_t t_acc = t_lane + buf * 256u;
        HC_T(eh);
        if (!is_out) {
          const int n_full = (int)hdr->layer_n[layer];
          const int n_blocks = n_full >> 5;
          uint8_t* h_img = nullptr;
          if (store)
            h_img = (is_nr ? bufs.h[0][layer] : bufs.h[1][layer - nr_layers]) + image_row_offset_c(tile_idx, row, n_full);
          uint32_t v[2][8] = {};
          if (!(dbg & 2)) tmem_ld8(t_acc + (uint32_t)(sub * 8), v[0]);
#pragma unroll 1
          for (int b = 0; b < n_blocks; b += 2) {
            // two blocks per iteration so that the register buffers are indexed statically
            if (!(dbg & 2)) tmem_ld_wait();
            if (b + 1 < n_blocks && !(dbg & 2)) tmem_ld8(t_acc + (uint32_t)((b + 1) * 32 + sub * 8), v[1]);
            {
              const uint32_t* u = v[0];
              const int chunk = b * 4 + sub;
              const uint4 pk = make_uint4(pack_relu_bf16(__uint_as_float(u[0]), __uint_as_float(u[1])),
                                          pack_relu_bf16(__uint_as_float(u[2]), __uint_as_float(u[3])),
                                          pack_relu_bf16(__uint_as_float(u[4]), __uint_as_float(u[5])),
                                          pack_relu_bf16(__uint_as_float(u[6]), __uint_as_float(u[7])));
              if (!(dbg & 6)) st_shared_v4(s_base + (uint32_t)chunk * kLboA + (uint32_t)row * 16u, pk.x, pk.y, pk.z, pk.w);
              if (store) *reinterpret_cast<uint4*>(h_img + (size_t)chunk * 256) = pk;
              tc_fence_before();
              if (!(dbg & 1)) fence_proxy_async();
              __syncwarp();
              if (lane == 0) mbar_arrive(s_a_blk + b * 8);
            }
            if (b + 1 < n_blocks) {
              if (!(dbg & 2)) tmem_ld_wait();
              if (b + 2 < n_blocks && !(dbg & 2)) tmem_ld8(t_acc + (uint32_t)((b + 2) * 32 + sub * 8), v[0]);
              const uint32_t* u = v[1];
              const int chunk = (b + 1) * 4 + sub;
              const uint4 pk = make_uint4(pack_relu_bf16(__uint_as_float(u[0]), __uint_as_float(u[1])),
                                          pack_relu_bf16(__uint_as_float(u[2]), __uint_as_float(u[3])),
                                          pack_relu_bf16(__uint_as_float(u[4]), __uint_as_float(u[5])),
                                          pack_relu_bf16(__uint_as_float(u[6]), __uint_as_float(u[7])));
              if (!(dbg & 6)) st_shared_v4(s_base + (uint32_t)chunk * kLboA + (uint32_t)row * 16u, pk.x, pk.y, pk.z, pk.w);
              if (store) *reinterpret_cast<uint4*>(h_img + (size_t)chunk * 256) = pk;
              tc_fence_before();
              if (!(dbg & 1)) fence_proxy_async();
              __syncwarp();
              if (lane == 0) mbar_arrive(s_a_blk + (b + 1) * 8);
            }
          }
          HC_ADD(3, eh);
        } else {
          uint32_t v[8];
          tmem_ld8(t_acc, v);
          tmem_ld_wait();
          tc_fence_before();
          if (is_nr) {
            const float o0 = __uint_as_float(v[0]), o1 = __uint_as_float(v[1]), o2 = __uint_as_float(v[2]);
            q[0] = x[0] + o0; q[1] = x[1] + o1; q[2] = x[2] + o2;
            if (valid && sub == 0) {
              off_out[sample * 3 + 0] = o0; off_out[sample * 3 + 1] = o1; off_out[sample * 3 + 2] = o2;
            }
          } else if (valid && sub == 0) {
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
    HC_ADD(0, e0);
    HC_FLUSH(8);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // the peer may still multicast into / arrive on this CTA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

int launch_field_tcc(const TcPlan& plan, const uint8_t* packed, int64_t n_samples, const float* x_skel,
                     const float* hann_w, int ignore_non_rigid, float* raw, float* xyz, float* offsets,
                     const TrainBufs* bufs, int sms, cudaStream_t st) {
  (void)plan;
  const char* dbg_env = getenv("HN_TC_ABLATE");
  const int dbg = dbg_env ? atoi(dbg_env) : 0;
  const int64_t n_tiles = ceil_div<int64_t>(n_samples, kTileM);
  const int64_t n_pairs = (n_tiles + 1) / 2;
  HN_TRY(check_cuda(cudaFuncSetAttribute(field_tcc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)cl::kSmemTotalC), "hn_field_fwd/cluster smem attr"));
  HN_TRY(check_cuda(cudaFuncSetAttribute(field_tcc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)cl::kSmemTotalC), "hn_field_fwd/cluster smem attr"));
  // persistent kernel: exactly as many clusters as can be co-resident (GPC boundaries may leave a few
  // SMs without a partner, so this can be below sms / 2)
  static int max_clusters_cached = 0;
  if (max_clusters_cached == 0) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * (sms / 2)));
    cfg.blockDim = dim3(cl::kThreadsC);
    cfg.dynamicSmemBytes = cl::kSmemTotalC;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    HN_TRY(check_cuda(cudaOccupancyMaxActiveClusters(&n, field_tcc_kernel<false>, &cfg), "hn_field_fwd/cluster occupancy"));
    if (n <= 0) return fail_arg("hn_field_fwd: no 2-CTA cluster of the tcgen05 kernel fits on this device");
    max_clusters_cached = n;
    if (getenv("HN_DEBUG")) fprintf(stderr, "humannerf_b200: %d co-resident 2-CTA clusters (%d SMs)\n", n, sms);
  }
  const unsigned grid = 2u * (unsigned)std::min<int64_t>(n_pairs, (int64_t)max_clusters_cached);
  if (bufs) {
    field_tcc_kernel<true><<<grid, cl::kThreadsC, cl::kSmemTotalC, st>>>(packed, n_samples, x_skel, hann_w, ignore_non_rigid,
                                                                       raw, xyz, offsets, *bufs, dbg);
  } else {
    TrainBufs none;
    memset(&none, 0, sizeof(none));
    field_tcc_kernel<false><<<grid, cl::kThreadsC, cl::kSmemTotalC, st>>>(packed, n_samples, x_skel, hann_w, ignore_non_rigid,
                                                                        raw, xyz, offsets, none, dbg);
  }
  return launched("hn_field_fwd/tcgen05 cluster");
}

}  // namespace hn

extern "C" int hn_debug_tcc_prof(unsigned long long* out_host) {
  using namespace hn;
  if (!out_host) return fail_arg("hn_debug_tcc_prof: null");
  HN_TRY(check_cuda(cudaMemcpyFromSymbol(out_host, g_tcc_prof, sizeof(unsigned long long) * 16), "hn_debug_tcc_prof/read"));
  unsigned long long zero[16] = {0};
  return check_cuda(cudaMemcpyToSymbol(g_tcc_prof, zero, sizeof(zero)), "hn_debug_tcc_prof/clear");
}
