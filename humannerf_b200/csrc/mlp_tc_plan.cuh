// Host-side plan of the tcgen05 canonical-field kernels: shared-memory geometry, the packed weight
// image (forward chunks W, backward-data chunks W^T), the round table of the backward chain and the
// layout of the training buffers (T16 tile images).  Shared by mlp_tc.cu and mlp_tc_bwd.cu.
#pragma once

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda.h>

#include "tc_common.cuh"
#include "tiles.cuh"

namespace hn {

constexpr int kTileM = 128;
constexpr int kTiles = 2;                               // tiles per CTA pass
constexpr int kStages = 4;
constexpr uint32_t kStageBytes = 16384;
constexpr uint32_t kActHTile = kTileM * 256 * 2;        // 64 KB per tile: hidden activations, <= 256 wide
constexpr uint32_t kActPeTile = kTileM * 64 * 2;        // 16 KB per tile: positional encoding, <= 64 wide
constexpr uint32_t kActHBytes = kTiles * kActHTile;
constexpr uint32_t kActPeBytes = kTiles * kActPeTile;
constexpr uint32_t kLboA = kTileM * 16;                 // bytes between K-adjacent core matrices of A
constexpr uint32_t kRingOff = kActHBytes + kActPeBytes;
constexpr uint32_t kBarOff = kRingOff + kStages * kStageBytes;
constexpr int kMaxChunks = 128;
constexpr uint32_t kTableOff = kBarOff + 128;
constexpr uint32_t kSmemTotal = kTableOff + kMaxChunks * 16;
constexpr int kTmemCols = 512;
constexpr int kEpiWarps = 16;
constexpr int kThreads = (3 + kEpiWarps) * 32;          // producer, 2 MMA issuers (one per tile), epilogue
constexpr int kMmaWarp0 = 1, kMmaWarp1 = 2 + kEpiWarps;
constexpr uint32_t kMagic = 0x484e5433u;                // "HNT3"
constexpr int kMaxTcLayers = 2 * HN_MAX_LAYERS;
constexpr int kMaxRounds = 2 * HN_MAX_LAYERS + 4;
static_assert(kSmemTotal <= 232448, "shared memory budget");

struct TcChunk {
  uint32_t w_off;     // byte offset of the chunk inside the weight image
  uint32_t bytes;     // multiple of 16, <= kStageBytes
  uint32_t a_off;     // byte offset of the chunk's first A core matrix inside tile 0's region
  uint16_t n;         // MMA N
  uint8_t ksteps;     // K / 16
  uint8_t flags;      // bit0 first chunk of its layer / round, bit1 last chunk, bit2 A operand in the PE block
};
static_assert(sizeof(TcChunk) == 16, "TcChunk must be 16 bytes");

// One accumulator round of the backward-data chain: MMAs of chunks [chunk0, chunk1) then an epilogue.
struct BwdRound {
  uint16_t chunk0, chunk1;
  uint16_t n;          // accumulator columns
  uint8_t epi;         // 0: dZ_prev = acc * [H_prev > 0] -> A block + dZ image;  1: acc -> d_pe (fp32)
  uint8_t layer;       // epi 0: hidden layer whose activation masks / whose dZ is produced; epi 1: d_pe slot
};
static_assert(sizeof(BwdRound) == 8, "BwdRound must be 8 bytes");

// One weight chunk of the TS-mode forward kernel (mlp_tc_ts.cu) = everything one accumulator half of one layer needs:
// the A operand of the hidden layers lives in TENSOR memory (written by the epilogue with tcgen05.st), only the
// encoding / bias step reads A from shared memory.  A layer of width 256 is computed as two N = 128 halves into the
// accumulator halves D0 / D1.  K rows of the chunk, in issue order: [encoding block incl. the bias row, or the 16-row
// bias step: SS] [hidden features [0,128): TS] [hidden features [128,256): TS, issued after a_ready[1]].
struct TsChunk {
  uint32_t w_off;     // byte offset inside the weight image (two half images, one per CTA of the pair)
  uint32_t bytes;     // both halves together
  uint16_t n;         // MMA N (128 or 16)
  uint8_t ss_a_off;   // index of the first core matrix (8 K columns) of the SS step inside the encoding block
  uint8_t ss_ksteps;  // K / 16 of the SS step
  uint8_t ts0_ksteps; // K / 16 of the hidden features [0,128)
  uint8_t ts1_ksteps; // K / 16 of the hidden features [128,256) (0: none)
  uint8_t flags;      // kTs* below
  uint8_t flags2;     // bit0: A buffer (TS), bit1: chunk of the non-rigid MLP
};
static_assert(sizeof(TsChunk) == 16, "TsChunk must be 16 bytes");
constexpr uint8_t kTsOut = 1;      // output layer: its completion is signalled on out_full (all epilogue warps wait there)
constexpr uint8_t kTsD1 = 8;       // accumulator half 1
constexpr uint8_t kTsWaitA0 = 16;  // issue only after a_ready[0] (previous layer's epilogue of features [0,128) / D0 drained)
constexpr uint8_t kTsWaitA1 = 32;  // the [128,256) steps wait for a_ready[1]
constexpr uint8_t kTsWaitPe = 64;  // ... pe_ready[mlp] (encoding block written; previous output accumulator drained)
constexpr uint8_t kTsPeCn = 128;   // the SS step reads the canonical encoding block (else the non-rigid one)
constexpr int kMaxTsChunks = 40;
constexpr uint32_t kTsStageBytes = 40960;      // per CTA: half of the largest chunk (N = 128, K = 64 + 256)

struct TcHeader {
  uint32_t magic, n_chunks, table_off, image_off, total_bytes, reserved;
  uint32_t nr_layers, cn_layers, nr_width, cn_width;
  uint32_t nr_freqs, cn_freqs, nr_pe_pad, cn_pe_pad;
  uint32_t layer_n[kMaxTcLayers];          // MMA N per layer (non-rigid layers first)
  // backward-data chain (per MLP: 0 = non-rigid, 1 = canonical)
  uint32_t bwd_table_off, bwd_n_chunks;    // chunk table of W^T chunks (w_off relative to image_off)
  uint32_t bwd_chunk0[2], bwd_chunk1[2];
  uint32_t bwd_round0[2], bwd_round1[2];
  BwdRound rounds[kMaxRounds];
};

// Training buffers: T16 tile images of the encoding blocks, hidden activations and pre-activation
// gradients.  Passed to the kernels by value.
struct TrainBufs {
  uint8_t* pe[2];                    // [mlp]            F = pe_pad
  uint8_t* h[2][HN_MAX_LAYERS];      // [mlp][layer]     F = width      (hidden layers)
  uint8_t* dz[2][HN_MAX_LAYERS];     // [mlp][layer]     F = width, 16 for the output layer
  uint8_t* bits[2][HN_MAX_LAYERS];   // [mlp][layer]     ReLU masks of the hidden layers, 1 bit per unit (relu_bits_* below);
                                     //                  null: the backward chain reads the mask from h instead
  int64_t n_tiles;                   // tiles that exist in the images
  int32_t leader_stores;             // cta_group::2 kernels: 1 = one elected lane per warp issues the warp's image copies
  int32_t reserved;
};

// Tensor maps over the T16 images for the 2 KB image stores of the cta_group::2 training kernels (store mode 4): an
// operand chunk in shared memory ([128 rows][16 B] = 8 sixteen-row groups x 256 B) is one dense TMA box whose 8
// pieces land F * 32 bytes apart in the image.  View of an image: dim0 = the 128 bf16 of a 256-byte block, dim1 = column
// chunk (stride 256 B), dim2 = sixteen-row group (stride F * 32 B); the hidden-layer images of one MLP (h0, dz0, h1,
// dz1, ... - carve_train_bufs places them back to back) share one rank-4 map with dim3 = image slot.
//   [m] hidden images of MLP m (slot 2l = h[m][l], 2l + 1 = dz[m][l]);  [2 + m] encoding image pe[m];
//   [4 + m] output-layer dZ image (F = 16)
struct alignas(64) ImageMaps {
  CUtensorMap m[6];
};

// ReLU-mask image of a hidden layer (cta_group::2 kernels).  The epilogue thread (row, half) of the forward kernel and
// of the backward chain own the same n_half = F / 2 columns of a row; its record holds one 32-bit word per 32 columns:
// bit j = column 2j of the group is active, bit 16 + j = column 2j + 1 (the order the packed bf16 pairs produce).
// Records are [tile][half][row] so that a warp reads / writes contiguous memory: 1/16 of the bytes of the bf16 image.
__host__ __device__ inline int relu_bits_record_bytes(int n_half) { return 8 * ((n_half + 63) / 64); }
__host__ __device__ inline size_t relu_bits_offset(int64_t tile, int half, int row, int n_half) {
  return (size_t)((tile * 2 + half) * 128 + row) * (size_t)relu_bits_record_bytes(n_half);
}

struct PackJob {            // one per chunk, consumed by pack_weights_kernel
  const float* src;         // weight matrix (rows = out features), nullptr = all zero
  const float* bias;        // bias vector placed in K row `bias_row` (nullptr / -1: none)
  const float* fold_w;      // + fold_w[n, 0:fold_k] . cond added to the bias (condition-code fold)
  int32_t src_ld;
  int32_t src_col0;         // source column of the chunk's k = 0 (n = 0 when transposed)
  int32_t valid_k;          // valid K extent from src_col0 (rest is zero padding)
  int32_t valid_n;          // valid N extent (rest is zero padding)
  int32_t n;                // chunk N
  int32_t kc;               // chunk K
  int32_t bias_row;
  int32_t fold_ld, fold_k;
  int32_t transpose;        // 0: B[n][k] = src[n][col0 + k];  1: B[n][k] = src[row0 + k][col0 + n]
  int32_t src_row0;
  uint32_t w_off;           // destination byte offset inside the image
  int32_t dst_k0, dst_kc;   // dst_kc > 0: the job fills K rows [dst_k0, dst_k0 + kc) of a chunk of dst_kc K rows (TS chunks)
  int32_t src_n0;           // non-transposed: first source row (B[n][k] = src[src_n0 + n][col0 + k], bias[src_n0 + n])
  int32_t split;            // 1: the chunk is stored as two half images (rows [0,n/2) then [n/2,n)), one per CTA of a
                            // cta_group::2 pair: addr = r*(kc/8)*(n/2)*16 + (k/8)*(n/2)*16 + n'*16
};

static inline int round16(int v) { return (v + 15) & ~15; }

// Ring geometry of the cta_group::2 kernel: 64 KB per CTA as 4 stages x 16 KB (default: chunks of up to 32 KB per
// pair = 4 MMAs at N = 256, so the ~500 clk an issuer spends per chunk stay below the chunk's tensor-pipe time) or,
// with HN_TC_2SM_STAGES=8, 8 stages x 8 KB (the chunk table of the tile-pair kernel).
static inline int tc2_stages() {
  static const int v = [] {
    const char* e = getenv("HN_TC_2SM_STAGES");
    return (e != nullptr && atoi(e) == 8) ? 8 : 4;
  }();
  return v;
}
// cta_group::2 training kernels: the forward writes 1-bit ReLU masks next to the bf16 activation images and the
// backward chain masks with them instead of re-reading the activations (16 x fewer bytes on that read);
// HN_TC_RELU_BITS=0 keeps the activation read (measurement aid / A-B switch)
static inline bool use_relu_bits() {
  static const bool v = [] {
    const char* e = getenv("HN_TC_RELU_BITS");
    return e == nullptr || atoi(e) != 0;
  }();
  return v;
}
// cta_group::2 training kernels, image store mode (store_blocks in mlp_tc_2sm.cuh): 4 = one 2 KB tensor-map copy per
// 8-column chunk of a tile, 1 = 256-byte bulk copies issued by one elected lane per warp.  Default: the backward chains
// use 4 (3.40 -> 3.32 ms per step), the forward 1 (1.44 ms; 1.59 ms with 4: its epilogue already spills, and the two
// named barriers per layer cost more than the 8 x fewer TMA requests save).  HN_TC_LEADER_STORES=1 / 4 forces both.
static inline int use_leader_stores(bool backward) {
  static const int forced = [] {
    const char* e = getenv("HN_TC_LEADER_STORES");
    return e == nullptr ? 0 : (atoi(e) == 1 ? 1 : 4);
  }();
  return forced != 0 ? forced : (backward ? 4 : 1);
}
// largest weight chunk (bytes, both halves) for the selected kernel family (forward and backward-data chain)
static inline uint32_t fwd_chunk_bytes() {
  return 2u * (65536u / (uint32_t)tc2_stages());
}

struct TcPlan {
  TcHeader hdr;
  std::vector<TcChunk> chunks, bwd_chunks;
  std::vector<TsChunk> ts_chunks;     // TS-mode forward (empty: widths other than 128 / 256)
  std::vector<PackJob> jobs;
  uint32_t jobs_off = 0;   // scratch table placed after the image
  int pe_dim[2], pe_pad[2], width[2], layers[2], skip[2], lead[2], out_dim[2];
};

inline int make_tc_plan(const hn_field_params* p, TcPlan& plan) {
  if (!p) return fail_arg("hn_field(bf16): null params");
  if (p->nr_layers < 2 || p->nr_layers > HN_MAX_LAYERS || p->cn_layers < 2 || p->cn_layers > HN_MAX_LAYERS)
    return fail_arg("hn_field(bf16): layer counts outside [2,%d]", HN_MAX_LAYERS);
  if (p->nr_width % 64 || p->cn_width % 64 || p->nr_width < 64 || p->cn_width < 64 || p->nr_width > 256 ||
      p->cn_width > 256)
    return fail_arg("hn_field(bf16): widths must be multiples of 64 in [64,256]");
  const int nr_pe = 6 * p->nr_freqs, cn_pe = 3 + 6 * p->cn_freqs;
  // one spare column of the encoding block carries the constant 1 that multiplies the bias row
  const int nr_pad = round16(nr_pe + 1), cn_pad = round16(cn_pe + 1);
  if (p->nr_freqs < 1 || p->cn_freqs < 1 || p->nr_freqs > 10 || p->cn_freqs > 10 || nr_pad > 64 || cn_pad > 64)
    return fail_arg("hn_field(bf16): encodings wider than 63 columns are not supported");
  if (p->nr_skip == 0 || p->nr_skip >= p->nr_layers - 1 || p->cn_skip == 0 || p->cn_skip >= p->cn_layers - 1)
    return fail_arg("hn_field(bf16): skip layer must be a hidden layer (or -1)");
  TcHeader& h = plan.hdr;
  memset(&h, 0, sizeof(h));
  h.magic = kMagic;
  h.nr_layers = p->nr_layers; h.cn_layers = p->cn_layers; h.nr_width = p->nr_width; h.cn_width = p->cn_width;
  h.nr_freqs = p->nr_freqs; h.cn_freqs = p->cn_freqs; h.nr_pe_pad = nr_pad; h.cn_pe_pad = cn_pad;
  plan.pe_dim[0] = nr_pe; plan.pe_dim[1] = cn_pe; plan.pe_pad[0] = nr_pad; plan.pe_pad[1] = cn_pad;
  plan.width[0] = p->nr_width; plan.width[1] = p->cn_width; plan.layers[0] = p->nr_layers; plan.layers[1] = p->cn_layers;
  plan.skip[0] = p->nr_skip; plan.skip[1] = p->cn_skip; plan.lead[0] = p->cond_dim; plan.lead[1] = 0;
  plan.out_dim[0] = 3; plan.out_dim[1] = 4;
  uint32_t w_off = 0;

  // ---- forward: one segment = a run of K columns of one A block against W[:, cols] ----------------
  auto add_segment = [&](std::vector<TcChunk>& table, const float* W, int ldw, int col0, int valid_k, int k_pad,
                         int valid_n, int n, bool in_pe, int a_col0, bool first, bool last, int bias_col,
                         const float* bias, const float* fold_w, int fold_ld, int fold_k, bool transpose, int row0) {
    const uint32_t chunk_cap = fwd_chunk_bytes();      // forward and backward tables share the kernel family
    int kc_max = (int)(chunk_cap / (uint32_t)(n * 2)) / 16 * 16;
    if (kc_max > 240) kc_max = 240;                       // ksteps is a uint8
    for (int k0 = 0; k0 < k_pad; k0 += kc_max) {
      const int kc = std::min(kc_max, k_pad - k0);
      TcChunk c;
      c.w_off = w_off;
      c.bytes = (uint32_t)(n * kc * 2);
      c.a_off = (uint32_t)((a_col0 + k0) / 8) * kLboA;
      c.n = (uint16_t)n;
      c.ksteps = (uint8_t)(kc / 16);
      c.flags = (uint8_t)(((first && k0 == 0) ? 1 : 0) | ((last && k0 + kc >= k_pad) ? 2 : 0) | (in_pe ? 4 : 0));
      table.push_back(c);
      PackJob j;
      memset(&j, 0, sizeof(j));
      j.src = W; j.src_ld = ldw; j.valid_n = valid_n; j.n = n; j.kc = kc; j.w_off = w_off;
      j.valid_k = std::max(0, valid_k - k0);
      j.transpose = transpose ? 1 : 0;
      j.split = 1;                                        // two half images, one per CTA of the pair
      if (transpose) { j.src_col0 = col0; j.src_row0 = row0 + k0; }
      else { j.src_col0 = col0 + k0; j.src_row0 = 0; }
      j.bias_row = -1;
      if (bias_col >= k0 && bias_col < k0 + kc) {
        j.bias_row = bias_col - k0; j.bias = bias; j.fold_w = fold_w; j.fold_ld = fold_ld; j.fold_k = fold_k;
      }
      plan.jobs.push_back(j);
      w_off += c.bytes;
    }
  };
  int layer_idx = 0;
  auto add_mlp = [&](int m, const float* const* W, const float* const* B) {
    const int layers = plan.layers[m], width = plan.width[m], skip = plan.skip[m], pe = plan.pe_dim[m];
    const int pe_pad = plan.pe_pad[m], lead = plan.lead[m];
    const bool pe_first = m == 1;
    const int ones_col = pe_pad - 1;
    for (int l = 0; l < layers; ++l) {
      const bool out = l == layers - 1;
      const int n = out ? 16 : width;
      const int valid_n = out ? plan.out_dim[m] : width;
      const int in_dim = (l == 0) ? lead + pe : (l == skip ? width + pe : width);
      const float* fold = (l == 0 && lead > 0) ? W[l] : nullptr;
      auto& T = plan.chunks;
      if (l == 0) {
        add_segment(T, W[l], in_dim, lead, pe, pe_pad, valid_n, n, true, 0, true, true, ones_col, B[l], fold, in_dim, lead, false, 0);
      } else if (l == skip) {
        if (pe_first) {
          add_segment(T, W[l], in_dim, 0, pe, pe_pad, valid_n, n, true, 0, true, false, ones_col, B[l], nullptr, 0, 0, false, 0);
          add_segment(T, W[l], in_dim, pe, width, width, valid_n, n, false, 0, false, true, -1, nullptr, nullptr, 0, 0, false, 0);
        } else {
          add_segment(T, W[l], in_dim, 0, width, width, valid_n, n, false, 0, true, false, -1, nullptr, nullptr, 0, 0, false, 0);
          add_segment(T, W[l], in_dim, width, pe, pe_pad, valid_n, n, true, 0, false, true, ones_col, B[l], nullptr, 0, 0, false, 0);
        }
      } else {
        add_segment(T, W[l], in_dim, 0, width, width, valid_n, n, false, 0, true, false, -1, nullptr, nullptr, 0, 0, false, 0);
        // bias step: the last 16 columns of the encoding block against an all-zero weight slab whose
        // last K row is the bias
        add_segment(T, nullptr, 0, 0, 0, 16, valid_n, n, true, pe_pad - 16, false, true, 15, B[l], nullptr, 0, 0, false, 0);
      }
      h.layer_n[layer_idx++] = (uint32_t)n;
    }
  };
  add_mlp(0, p->nr_w, p->nr_b);
  add_mlp(1, p->cn_w, p->cn_b);

  // ---- TS-mode forward (mlp_tc_ts.cu): see TsChunk -------------------------------------------------------
  if ((p->nr_width == 128 || p->nr_width == 256) && (p->cn_width == 128 || p->cn_width == 256)) {
    // one pack job per K segment of a chunk
    auto add_seg = [&](const float* W, int ldw, int n0, int valid_n, int n, int col0, int valid_k, int kc, int dst_k0,
                       int dst_kc, int bias_row, const float* bias, const float* fold_w, int fold_ld, int fold_k) {
      PackJob j;
      memset(&j, 0, sizeof(j));
      j.src = W; j.src_ld = ldw; j.src_n0 = n0; j.valid_n = std::max(0, std::min(n, valid_n - n0)); j.n = n; j.kc = kc;
      j.w_off = w_off; j.src_col0 = col0; j.valid_k = valid_k; j.split = 1; j.bias_row = bias_row;
      j.dst_k0 = dst_k0; j.dst_kc = dst_kc;
      if (bias_row >= 0) { j.bias = bias; j.fold_w = fold_w; j.fold_ld = fold_ld; j.fold_k = fold_k; }
      plan.jobs.push_back(j);
    };
    for (int m = 0; m < 2; ++m) {
      const float* const* W = m == 0 ? p->nr_w : p->cn_w;
      const float* const* B = m == 0 ? p->nr_b : p->cn_b;
      const int layers = plan.layers[m], width = plan.width[m], skip = plan.skip[m], pe = plan.pe_dim[m];
      const int pe_pad = plan.pe_pad[m], lead = plan.lead[m];
      const bool pe_first = m == 1;
      const uint8_t f2m = m == 0 ? 2 : 0, pe_sel = m == 1 ? kTsPeCn : 0;
      for (int l = 0; l < layers; ++l) {
        const bool out = l == layers - 1;
        const int n_total = out ? 16 : width, valid_total = out ? plan.out_dim[m] : width;
        const int halves = n_total > 128 ? 2 : 1, nn = out ? 16 : 128;
        const int in_dim = (l == 0) ? lead + pe : (l == skip ? width + pe : width);
        const bool in_pe = l == 0 || l == skip;
        const int pe_col = (l == 0) ? lead : (pe_first ? 0 : width);
        const int h_col = (l == skip && pe_first) ? pe : 0;
        const int k_ss = in_pe ? pe_pad : 16, k_ts = l > 0 ? width : 0, k_tot = k_ss + k_ts;
        for (int dh = 0; dh < halves; ++dh) {
          const int n0 = dh * 128;
          TsChunk c;
          memset(&c, 0, sizeof(c));
          c.w_off = w_off; c.bytes = (uint32_t)(nn * k_tot * 2); c.n = (uint16_t)nn;
          c.ss_a_off = (uint8_t)(in_pe ? 0 : (pe_pad - 16) / 8); c.ss_ksteps = (uint8_t)(k_ss / 16);
          c.ts0_ksteps = (uint8_t)(std::min(k_ts, 128) / 16); c.ts1_ksteps = (uint8_t)(std::max(k_ts - 128, 0) / 16);
          c.flags = (uint8_t)((dh ? kTsD1 : 0) | pe_sel | (out ? kTsOut : 0));
          if (dh == 0) c.flags |= l == 0 ? kTsWaitPe : kTsWaitA0;
          if (dh == 0 && c.ts1_ksteps) c.flags |= kTsWaitA1;
          c.flags2 = (uint8_t)(f2m | ((l + 1) & 1));                 // layer l - 1 wrote A buffer (l - 1) & 1
          plan.ts_chunks.push_back(c);
          if (in_pe)
            add_seg(W[l], in_dim, n0, valid_total, nn, pe_col, pe, pe_pad, 0, k_tot, pe_pad - 1, B[l],
                    (l == 0 && lead > 0) ? W[l] : nullptr, in_dim, lead);
          else
            add_seg(nullptr, 0, n0, valid_total, nn, 0, 0, 16, 0, k_tot, 15, B[l], nullptr, 0, 0);
          if (k_ts > 0) add_seg(W[l], in_dim, n0, valid_total, nn, h_col, k_ts, k_ts, k_ss, k_tot, -1, nullptr, nullptr, 0, 0);
          w_off += c.bytes;
        }
      }
    }
    if ((int)plan.ts_chunks.size() > kMaxTsChunks) return fail_arg("hn_field(bf16): too many TS weight chunks");
    for (const TsChunk& c : plan.ts_chunks)
      if (c.bytes > 2 * kTsStageBytes || c.bytes % 32) return fail_arg("hn_field(bf16): internal TS chunk size error");
  }

  // ---- backward data: for layer l (last to first) dX = dZ_l W_l restricted to a column window.
  //      A = dZ_l (K = out features of layer l, in the hidden A block), B = (W_l window)^T -----------
  int n_rounds = 0;
  auto add_round = [&](int chunk0, int n, int epi, int layer) {
    BwdRound r;
    r.chunk0 = (uint16_t)chunk0; r.chunk1 = (uint16_t)plan.bwd_chunks.size(); r.n = (uint16_t)n;
    r.epi = (uint8_t)epi; r.layer = (uint8_t)layer;
    h.rounds[n_rounds++] = r;
  };
  auto add_mlp_bwd = [&](int m, const float* const* W) {
    const int layers = plan.layers[m], width = plan.width[m], skip = plan.skip[m], pe = plan.pe_dim[m];
    const int pe_pad = plan.pe_pad[m], lead = plan.lead[m];
    const bool pe_first = m == 1;
    h.bwd_chunk0[m] = (uint32_t)plan.bwd_chunks.size();
    h.bwd_round0[m] = (uint32_t)n_rounds;
    for (int l = layers - 1; l >= 0; --l) {
      const bool out = l == layers - 1;
      const int k_pad = out ? 16 : width;                        // K = out features of layer l
      const int valid_k = out ? plan.out_dim[m] : width;
      const int in_dim = (l == 0) ? lead + pe : (l == skip ? width + pe : width);
      const int pe_col = (l == 0) ? lead : (pe_first ? 0 : width);
      const int h_col = (l == skip && pe_first) ? pe : 0;
      if (l == 0 || l == skip) {
        // gradient of the encoding columns -> fp32 buffer (slot 0: skip layer, slot 1: layer 0)
        const int c0 = (int)plan.bwd_chunks.size();
        add_segment(plan.bwd_chunks, W[l], in_dim, pe_col, valid_k, k_pad, pe, pe_pad, false, 0, true, true, -1, nullptr,
                    nullptr, 0, 0, true, 0);
        add_round(c0, pe_pad, 1, l == 0 ? 1 : 0);
      }
      if (l > 0) {
        // gradient of the previous hidden layer -> masked by its activation -> dZ_{l-1}
        const int c0 = (int)plan.bwd_chunks.size();
        add_segment(plan.bwd_chunks, W[l], in_dim, h_col, valid_k, k_pad, width, width, false, 0, true, true, -1, nullptr,
                    nullptr, 0, 0, true, 0);
        add_round(c0, width, 0, l - 1);
      }
    }
    h.bwd_chunk1[m] = (uint32_t)plan.bwd_chunks.size();
    h.bwd_round1[m] = (uint32_t)n_rounds;
  };
  add_mlp_bwd(0, p->nr_w);
  add_mlp_bwd(1, p->cn_w);
  if (n_rounds > kMaxRounds) return fail_arg("hn_field(bf16): too many backward rounds");

  if ((int)plan.chunks.size() > kMaxChunks || (int)plan.bwd_chunks.size() > kMaxChunks)
    return fail_arg("hn_field(bf16): too many weight chunks");
  for (const auto* tab : {&plan.chunks, &plan.bwd_chunks})
    for (const TcChunk& c : *tab)
      if (c.bytes > fwd_chunk_bytes() || c.bytes % 32 || c.ksteps == 0)
        return fail_arg("hn_field(bf16): internal chunk size error");
  h.n_chunks = (uint32_t)plan.chunks.size();
  h.bwd_n_chunks = (uint32_t)plan.bwd_chunks.size();
  auto align = [](uint32_t v, uint32_t a) { return (v + a - 1) / a * a; };
  h.table_off = align((uint32_t)sizeof(TcHeader), 16);
  h.bwd_table_off = align(h.table_off + h.n_chunks * (uint32_t)sizeof(TcChunk), 16);
  h.image_off = align(h.bwd_table_off + h.bwd_n_chunks * (uint32_t)sizeof(TcChunk), 128);
  plan.jobs_off = align(h.image_off + w_off, 128);
  h.total_bytes = align(plan.jobs_off + (uint32_t)(plan.jobs.size() * sizeof(PackJob)), 128);
  return 0;
}

// Carves the training buffers out of `base` (may be null to only measure). Returns the bytes used.
inline int64_t carve_train_bufs(const TcPlan& plan, int64_t n_samples, uint8_t* base, TrainBufs* out, bool backward = false) {
  const int64_t n_tiles = ceil_div<int64_t>(std::max<int64_t>(n_samples, 1), 128);
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    uint8_t* p = base ? base + off : nullptr;
    off += (bytes + 255) / 256 * 256;
    return p;
  };
  TrainBufs b;
  memset(&b, 0, sizeof(b));
  b.n_tiles = n_tiles;
  b.leader_stores = use_leader_stores(backward);
  for (int m = 0; m < 2; ++m) {
    b.pe[m] = take(n_tiles * 128 * plan.pe_pad[m] * 2);
    for (int l = 0; l < plan.layers[m]; ++l) {
      const bool outl = l == plan.layers[m] - 1;
      if (!outl) b.h[m][l] = take(n_tiles * 128 * plan.width[m] * 2);
      b.dz[m][l] = take(n_tiles * 128 * (outl ? 16 : plan.width[m]) * 2);
    }
  }
  // after everything else: the offsets of the images above do not depend on the mask images
  for (int m = 0; m < 2; ++m)
    for (int l = 0; l + 1 < plan.layers[m]; ++l)
      b.bits[m][l] = take(n_tiles * 2 * 128 * relu_bits_record_bytes(plan.width[m] / 2));
  if (out) *out = b;
  return off;
}

// Encodes the ImageMaps of a carved training workspace (host; the driver entry point comes from the runtime, libcuda
// is not linked).  Returns 0 or a negative argument error.
inline int make_image_maps(const TcPlan& plan, const TrainBufs& b, ImageMaps* out) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return reinterpret_cast<EncodeFn>(fn);
  }();
  if (!encode) return fail_arg("tensor-map image stores: cuTensorMapEncodeTiled is not available from this driver");
  memset(out, 0, sizeof(*out));
  const cuuint64_t groups = (cuuint64_t)b.n_tiles * 8;
  auto enc = [&](CUtensorMap* m, void* base, int F, int slots) -> int {
    if (!base || F < 8 || (F & 7) || (reinterpret_cast<uintptr_t>(base) & 255)) return fail_arg("tensor-map image stores: bad image (F = %d)", F);
    const cuuint64_t gdim[4] = {128, (cuuint64_t)(F / 8), groups, (cuuint64_t)std::max(slots, 1)};
    const cuuint64_t gstr[3] = {256, (cuuint64_t)F * 32, groups * (cuuint64_t)F * 32};
    const cuuint32_t box[4] = {128, 1, 8, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, slots > 0 ? 4 : 3, base, gdim, gstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail_arg("tensor-map image stores: cuTensorMapEncodeTiled failed (%d, F = %d)", (int)r, F);
    return 0;
  };
  for (int m = 0; m < 2; ++m) {
    const int L = plan.layers[m], F = plan.width[m];
    if (L < 1) continue;
    const int64_t img = b.n_tiles * 128 * (int64_t)F * 2;
    if (L > 1) {
      // the hidden images must sit back to back: h0, dz0, h1, dz1, ...
      for (int l = 0; l + 1 < L; ++l)
        if (b.h[m][l] != b.h[m][0] + (int64_t)(2 * l) * img || b.dz[m][l] != b.h[m][0] + (int64_t)(2 * l + 1) * img)
          return fail_arg("tensor-map image stores: unexpected layout of the training workspace");
      HN_TRY(enc(&out->m[m], b.h[m][0], F, 2 * (L - 1)));
    }
    HN_TRY(enc(&out->m[2 + m], b.pe[m], plan.pe_pad[m], 0));
    HN_TRY(enc(&out->m[4 + m], b.dz[m][L - 1], 16, 0));
  }
  return 0;
}

}  // namespace hn
