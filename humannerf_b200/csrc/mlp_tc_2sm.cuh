// Shared by the cta_group::2 kernels (mlp_tc_2sm.cu forward, mlp_tc_bwd_2sm.cu backward-data chain): shared-memory
// geometry, the tcgen05 cta_group::2 wrappers, the issuers' parameter table and the bulk image stores.
#pragma once

#include "mlp_tc_plan.cuh"

namespace hn {

namespace s2 {
constexpr int kStages2 = 8;                                            // barrier slots (the ring has 4 or 8 stages)
constexpr uint32_t kRingBytes2 = 65536;                                // per CTA: 4 x 16 KB or 8 x 8 KB half chunks
constexpr uint32_t kBarOff2 = kRingOff + kRingBytes2;
// barriers: full[8] floc[8] empty[8] acc_full[2] a_ready[2] + tmem slot
constexpr uint32_t kBarBytes2 = (3 * kStages2 + 2 * kTiles) * 8 + 8 + 16;   // + lead / ack barriers of the forward kernel
constexpr uint32_t kTableOff2 = kBarOff2 + 256;
constexpr uint32_t kSmemTotal2 = kTableOff2 + kMaxChunks * 16;
constexpr int kSetRows = 2 * kTileM;                                   // rows of one tile set (MMA M)
static_assert(kBarBytes2 <= 256, "barrier block");
static_assert(kSmemTotal2 <= 232448, "shared memory budget");
}  // namespace s2

__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem of both CTAs] (+)= A (256 rows, 128 per CTA) * B^T (N rows, N/2 per CTA); issued by one thread of the leader
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs once all previously issued MMAs of the pair retire
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
// instruction descriptor: D = F32, A = B = BF16, both K-major, M = 256 (pair), N
__device__ __forceinline__ uint32_t make_idesc2(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

// What the MMA issuers need per chunk, as a KERNEL PARAMETER (constant bank, warp-uniform loads): x = byte offset of
// the chunk's first A core matrix inside its operand block, y = N | K steps << 16 | flags << 24 (TcChunk's last word).
// The issuer warps run convergently with every operand in uniform registers and elect one lane per MMA: a lone
// diverged thread pays a ~15-instruction ELECT / R2UR.BROADCAST waterfall per tcgen05 instruction plus ~100 dependent
// vector instructions per chunk (measured: 580 clk per chunk + 61 clk per MMA, against 256 clk of tensor pipe time).
struct IssueTab {
  uint2 e[kMaxChunks];
};

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(pred));
  return pred;
}

// Training: the operand block a warp has just written IS the T16 image of its rows (16 rows x 8 columns = 256
// contiguous bytes in both).  A warp owns 2 sixteen-row groups (2 * quarter, 2 * quarter + 1) x `n_chunks` column
// chunks chunk0, chunk0 + stride, ...; the four warps (quarters) of a (tile, column half) own those chunks completely.
//   mode = 4 (default): a chunk of all 128 rows ([128][16 B] = 2 KB, dense in shared memory) leaves with ONE tensor copy
//                   through `tmap` (ImageMaps: its 8 sixteen-row pieces land F * 32 bytes apart); the four warps meet
//                   on the named barrier `bar_id`, then an elected lane of quarter 0 issues the group's copies.
//                   Measured with equal bytes: 8 x fewer TMA requests are worth 0.36 ms per training step;
//   mode = 1:       256-byte bulk copies (the largest piece contiguous on both sides without a tensor map), all of a
//                   warp's copies issued by one elected lane in a loop over uniform addresses;
//                   (one copy per lane, the first version, made ptxas wrap every UBLKCP in a 9-instruction R2UR
//                   waterfall iteration: 7.23 -> 6.96 ms per training step when the elected lane took over).
// Callers: fence.proxy.async by every writer first; the source is reused only after block_reusable() below.
__device__ __forceinline__ void store_blocks(int mode, const CUtensorMap* tmap, int slot, uint32_t bar_id, uint8_t* image,
                                             int64_t tile_idx, int F, uint32_t s_block, int quarter, int lane, int chunk0,
                                             int stride, int n_chunks) {
  if (mode == 4) {
    named_bar_sync(bar_id, 128);                       // all 128 rows of the group's chunks are written and fenced
    if (quarter == 0 && elect_one()) {
      for (int i = 0; i < n_chunks; ++i) {
        const int chunk = chunk0 + i * stride;
        if (slot >= 0) tma_store_4d(tmap, s_block + (uint32_t)chunk * kLboA, 0, chunk, (int)(tile_idx * 8), slot);
        else tma_store_3d(tmap, s_block + (uint32_t)chunk * kLboA, 0, chunk, (int)(tile_idx * 8));
      }
      bulk_commit();
    }
    __syncwarp();
    return;
  }
  uint8_t* g0 = image + (size_t)tile_idx * 128 * F * 2 + (size_t)(2 * quarter) * F * 32;
  const uint32_t s0 = s_block + (uint32_t)(2 * quarter) * 256u;
  if (elect_one()) {
    for (int i = 0; i < n_chunks; ++i) {
      const int chunk = chunk0 + i * stride;
      bulk_s2g(g0 + (size_t)chunk * 256, s0 + (uint32_t)chunk * kLboA, 256u);
      bulk_s2g(g0 + (size_t)F * 32 + (size_t)chunk * 256, s0 + (uint32_t)chunk * kLboA + 256u, 256u);
    }
    bulk_commit();
  }
  __syncwarp();
}

// The image copies that read this warp's rows of an operand block have left shared memory (call before overwriting).
// Mode 1: every lane waits for its own copies; mode 4: the issuing lane (quarter 0) waits, the group's named
// barrier tells the other three warps.
__device__ __forceinline__ void block_reusable(int mode, uint32_t bar_id) {
  bulk_wait_read0();
  if (mode == 4) named_bar_sync(bar_id, 128);
  else __syncwarp();
}

}  // namespace hn
