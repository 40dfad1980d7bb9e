// Weight gradients on the tensor cores: dW[n_out][k_in] += sum_samples dZ[s][n_out] * X[s][k_in]
// with both operands read from "T16 tile images" in HBM (bf16), the layout the tcgen05 forward /
// backward-data kernels write their activations and pre-activation gradients in:
//
//   image(X)[s][f] at  (s/128) * 128*F*2  +  ((s%128)/16) * 16*F*2  +  (f/8) * 256  +  (s%16) * 16  +  (f%8) * 2
//
// i.e. per 16 samples and 8 features one 256-byte block of 16 rows x 16 bytes.  A block of 16 samples
// with all F features is contiguous (F*32 bytes), so a pipeline stage of 64 samples is one or a few
// 1-D TMA bulk copies, and the same bytes are a valid UMMA *MN-major* no-swizzle operand for a K = 16
// step (K = samples): 8 K rows of 16 bytes per core matrix, LBO (next 8 samples) = 128 B, SBO (next
// 8 features) = 256 B.  D = dZ^T X accumulates in TMEM over a slab of samples; the epilogue adds it
// into the fp32 gradient with atomics (split over samples across CTAs).
//
// Replaces the autograd of nn.Linear in reference non_rigid_motion_mlps/mlp_offset.py:74-84 and
// canonical_mlps/mlp_rgb_sigma.py:163-166,198 on the bf16 path.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tc_common.cuh"
#include "tiles.cuh"

namespace hn {

constexpr int kWgStages = 3;
constexpr int kWgStageSamples = 64;
constexpr uint32_t kWgOperandBytes = 256 * kWgStageSamples * 2;       // 32 KB: up to 256 features x 64 samples
constexpr uint32_t kWgStageBytes = 2 * kWgOperandBytes;               // A (dZ) + B (X)
constexpr uint32_t kWgBarOff = kWgStages * kWgStageBytes;
constexpr uint32_t kWgSmem = kWgBarOff + 128;
constexpr int kWgItemsPerSm = 3;
constexpr int kWgThreads = 6 * 32;                                    // producer, MMA, 4 epilogue warps

// MN-major for both operands: bits 15 / 16 of the instruction descriptor
__device__ __forceinline__ uint32_t make_idesc_mn(uint32_t n) { return make_idesc(n) | (1u << 15) | (1u << 16); }

// The job list travels as a kernel parameter (constant bank): no host->device copy, so a launch can be
// captured in a CUDA graph.
struct WgradJobTable {
  WgradJob jobs[kMaxWgradJobs];
  int slab0[kMaxWgradJobs + 1];      // work items of job j: [slab0[j], slab0[j + 1]) - one item = one slab of samples
};

// Work decomposition: job j is cut into n_slabs_j = slab0[j + 1] - slab0[j] slabs of samples, n_slabs_j proportional to
// the bytes the job streams per sample (f_dz + n features), so every (job, slab) item costs about the same and the
// item count is a whole multiple of the grid: the static round-robin over the CTAs is balanced (a common slab count
// for all jobs left the slowest CTA 16 % above the mean with the 144 ... 512-byte jobs of the training step).
struct WgradItem { int job, slab, n_slabs; };
// even split of the tiles over a job's slabs (slabs differ by at most one tile; empty when n_tiles < n_slabs)
__device__ __forceinline__ int slab_begin(int n_tiles, int slab, int n_slabs) {
  return (int)((int64_t)slab * n_tiles / n_slabs);
}
__device__ __forceinline__ WgradItem find_item(const WgradJobTable& tab, int n_jobs, int item) {
  int j = 0;
  while (j + 1 < n_jobs && item >= tab.slab0[j + 1]) ++j;
  return {j, item - tab.slab0[j], tab.slab0[j + 1] - tab.slab0[j]};
}

__global__ void __launch_bounds__(kWgThreads, 1)
tile_wgrad_kernel(const __grid_constant__ WgradJobTable tab, int n_jobs, int n_tiles_cap,
                  const int64_t* __restrict__ n_dev) {
  const WgradJob* jobs = tab.jobs;
  // n_dev (nullable): number of samples actually present (device-side count, see field_tc_kernel); the slab
  // decomposition keeps the host's slab COUNTS and shrinks the slabs
  int n_tiles = n_tiles_cap;
  if (n_dev) n_tiles = (int)min((max(__ldg(n_dev), (int64_t)0) + 127) / 128, (int64_t)n_tiles_cap);
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_full = s_base + kWgBarOff;
  const uint32_t s_empty = s_full + kWgStages * 8;
  const uint32_t s_acc = s_empty + kWgStages * 8;
  const uint32_t s_drained = s_acc + 8;
  const uint32_t s_tmem_slot = s_drained + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kWgBarOff + (2 * kWgStages + 2) * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = tab.slab0[n_jobs];

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(s_full + s * 8, 1);
      mbar_init(s_empty + s * 8, 1 + 4);         // the MMA commit + one arrive per epilogue warp (bias column sums)
    }
    mbar_init(s_acc, 1);
    mbar_init(s_drained, 128);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(s_tmem_slot, 512);
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
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const WgradItem it = find_item(tab, n_jobs, item);
        const WgradJob job = jobs[it.job];
        const int t0 = slab_begin(n_tiles, it.slab, it.n_slabs), t1 = slab_begin(n_tiles, it.slab + 1, it.n_slabs);
        const uint32_t a_blk = (uint32_t)job.f_dz * 32u;       // bytes per 16-sample block of the dZ image
        const uint32_t x_blk = (uint32_t)job.f_x * 32u;
        const uint32_t b_blk = (uint32_t)job.n * 32u;          // bytes of the X window per 16-sample block
        const bool x_whole = (job.x_f0 == 0 && job.n == job.f_x);
        for (int t = t0; t < t1; ++t) {
          const uint8_t* a_src = static_cast<const uint8_t*>(job.dz) + (size_t)t * 128 * job.f_dz * 2;
          const uint8_t* x_src = static_cast<const uint8_t*>(job.x) + (size_t)t * 128 * job.f_x * 2 + (size_t)(job.x_f0 / 8) * 256;
          for (int q = 0; q < 128 / kWgStageSamples; ++q) {
            const uint32_t a_dst = s_base + stage * kWgStageBytes, b_dst = a_dst + kWgOperandBytes;
            mbar_wait(s_empty + stage * 8, phase ^ 1, 1);
            mbar_expect_tx(s_full + stage * 8, 4u * (a_blk + b_blk));
            tma_bulk_g2s(a_dst, a_src + (size_t)q * 4 * a_blk, 4u * a_blk, s_full + stage * 8);
            if (x_whole) {
              tma_bulk_g2s(b_dst, x_src + (size_t)q * 4 * x_blk, 4u * b_blk, s_full + stage * 8);
            } else {
              for (int kb = 0; kb < 4; ++kb)
                tma_bulk_g2s(b_dst + kb * b_blk, x_src + (size_t)(q * 4 + kb) * x_blk, b_blk, s_full + stage * 8);
            }
            if (++stage == kWgStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, drained_phase = 0;
      const uint32_t desc_hi = (256u >> 4) | (1u << 14);      // SBO = 256 B (next 8 features), version 1
      const uint32_t lo_const = ((128u >> 4) << 16);          // LBO = 128 B (next 8 samples)
      bool first_item = true;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const WgradItem it = find_item(tab, n_jobs, item);
        const WgradJob job = jobs[it.job];
        const int t0 = slab_begin(n_tiles, it.slab, it.n_slabs), t1 = slab_begin(n_tiles, it.slab + 1, it.n_slabs);
        const int m_tiles = (job.m + 127) / 128;
        const uint32_t idesc = make_idesc_mn((uint32_t)job.n);
        const uint32_t a_blk = (uint32_t)job.f_dz * 32u, b_blk = (uint32_t)job.n * 32u;
        if (!first_item) {
          // the epilogue must have drained the previous item's accumulators
          mbar_wait(s_drained, drained_phase, 2);
          drained_phase ^= 1;
          tc_fence_after();
        }
        first_item = false;
        uint32_t acc = 0;
        for (int t = t0; t < t1; ++t) {
          for (int q = 0; q < 128 / kWgStageSamples; ++q) {
            mbar_wait(s_full + stage * 8, phase, 3);
            tc_fence_after();
            const uint32_t a_addr = s_base + stage * kWgStageBytes, b_addr = a_addr + kWgOperandBytes;
#pragma unroll 1
            for (uint32_t kb = 0; kb < 4; ++kb) {
              const uint32_t b_lo = (((b_addr + kb * b_blk) & 0x3FFFFu) >> 4) | lo_const;
              for (int mt = 0; mt < m_tiles; ++mt) {
                const uint32_t a_lo = (((a_addr + kb * a_blk + (uint32_t)mt * 16u * 256u) & 0x3FFFFu) >> 4) | lo_const;
                umma_bf16(tmem_base + (uint32_t)mt * 256u, ((uint64_t)desc_hi << 32) | a_lo,
                          ((uint64_t)desc_hi << 32) | b_lo, idesc, acc);
              }
              acc = 1u;
            }
            umma_commit(s_empty + stage * 8);
            if (++stage == kWgStages) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit(s_acc);
      }
    }
  } else {
    // ===================== epilogue (warps 2..5): TMEM -> fp32 atomics =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    uint32_t acc_phase = 0;
    uint32_t e_stage = 0, e_phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const WgradItem it = find_item(tab, n_jobs, item);
      const WgradJob job = jobs[it.job];
      const int m_tiles = (job.m + 127) / 128;
      const int t0 = slab_begin(n_tiles, it.slab, it.n_slabs), t1 = slab_begin(n_tiles, it.slab + 1, it.n_slabs);
      const bool has_tiles = t1 > t0;                              // an empty slab left the accumulators undefined
      // While the MMA warp works through the stages, these four otherwise idle warps read the dZ operand of
      // every stage from shared memory and keep its column sums = the bias gradient (bias_col == -2), so no
      // second job has to stream dZ from HBM again just for db.  Thread (warp w, lane l) owns the 8 features
      // of 16-byte chunk l and the 16 samples of block w of the 64-sample stage.
      {
        const bool direct = job.bias_col == -2 && job.db != nullptr;
        const int ew = warp - 2;
        const bool own = direct && lane * 8 < job.m;
        const uint32_t a_blk = (uint32_t)job.f_dz * 32u;
        float bsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int t = t0; t < t1; ++t) {
          for (int q = 0; q < 128 / kWgStageSamples; ++q) {
            {
              // always wait for the stage, also when this job needs no sums: it keeps these warps in step with
              // the ring (an early second arrive on the same `empty` phase would corrupt the barrier)
              mbar_wait(s_full + e_stage * 8, e_phase, 5);
              if (own) {
                const uint32_t src = s_base + e_stage * kWgStageBytes + (uint32_t)ew * a_blk + (uint32_t)lane * 256u;
#pragma unroll 4
                for (int r = 0; r < 16; ++r) {
                  uint32_t w0, w1, w2, w3;
                  // lane l starts at sample l: the 8 lanes of a quarter warp (256 bytes apart) hit 8 different 16-byte bank
                  // groups instead of one (the straight order made every load a 32-way conflict, 2048 shared-memory
                  // wavefronts per 64 KB stage - more than the MMAs and the TMA writes of the stage together)
                  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                               : "r"(src + (((uint32_t)(r + lane) & 15u) << 4)));
                  bsum[0] += __uint_as_float(w0 << 16); bsum[1] += __uint_as_float(w0 & 0xFFFF0000u);
                  bsum[2] += __uint_as_float(w1 << 16); bsum[3] += __uint_as_float(w1 & 0xFFFF0000u);
                  bsum[4] += __uint_as_float(w2 << 16); bsum[5] += __uint_as_float(w2 & 0xFFFF0000u);
                  bsum[6] += __uint_as_float(w3 << 16); bsum[7] += __uint_as_float(w3 & 0xFFFF0000u);
                }
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty + e_stage * 8);
            if (++e_stage == kWgStages) { e_stage = 0; e_phase ^= 1; }
          }
        }
        if (own && has_tiles) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (lane * 8 + i < job.m) atomicAdd(job.db + lane * 8 + i, bsum[i]);
        }
      }
      mbar_wait(s_acc, acc_phase, 4);
      acc_phase ^= 1;
      tc_fence_after();
      for (int mt = 0; has_tiles && mt < m_tiles; ++mt) {
        const int r = mt * 128 + row;
        const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)mt * 256u;
        for (int c0 = 0; c0 < job.n; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(t_addr + (uint32_t)c0, v);
          tmem_ld_wait();
          if (r < job.m) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int c = c0 + j;
              const float val = __uint_as_float(v[j]);
              if (c < job.n_valid) atomicAdd(job.dw + (int64_t)r * job.dw_ld + job.dw_col0 + c, val);
              else if (c == job.bias_col && job.bias_col >= 0 && job.db) atomicAdd(job.db + r, val);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(s_drained);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// fp32 row-major <-> T16 tile image (tests, and the d_raw / d_xyz entry of the backward chain)
// ------------------------------------------------------------------------------------------
__global__ void tile_pack_kernel(int64_t n_samples, int n_feat, int f_pad, const float* __restrict__ src,
                                 int64_t ld, uint8_t* __restrict__ image) {
  const int64_t n_tiles = (n_samples + 127) / 128;
  const int64_t units = n_tiles * 128 * (f_pad / 8);
  for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < units; u += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(u % (f_pad / 8));
    const int64_t s = u / (f_pad / 8);
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = g * 8 + i;
      f[i] = (s < n_samples && c < n_feat) ? src[s * ld + c] : 0.0f;
    }
    uint4 o;
    o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]); o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(image + tile_offset(s, g * 8, f_pad)) = o;
  }
}

__global__ void tile_unpack_kernel(int64_t n_samples, int n_feat, int f_pad, const uint8_t* __restrict__ image,
                                   float* __restrict__ dst, int64_t ld) {
  const int64_t total = n_samples * n_feat;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / n_feat;
    const int c = (int)(i - s * n_feat);
    const __nv_bfloat16 v = *reinterpret_cast<const __nv_bfloat16*>(image + tile_offset(s, c, f_pad));
    dst[s * ld + c] = __bfloat162float(v);
  }
}

int launch_tile_wgrad(const WgradJob* jobs_host, int n_jobs, int64_t n_samples, const int64_t* n_samples_dev,
                      cudaStream_t st) {
  if (n_jobs <= 0 || n_samples <= 0) return 0;
  if (n_jobs > kMaxWgradJobs) return fail_arg("tile_wgrad: at most %d jobs per launch", kMaxWgradJobs);
  WgradJobTable tab;
  memset(&tab, 0, sizeof(tab));
  memcpy(tab.jobs, jobs_host, (size_t)n_jobs * sizeof(WgradJob));
  int dev = 0, sms = 0;
  HN_TRY(check_cuda(cudaGetDevice(&dev), "tile_wgrad/device"));
  HN_TRY(check_cuda(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "tile_wgrad/attr"));
  HN_TRY(check_cuda(cudaFuncSetAttribute(tile_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kWgSmem), "tile_wgrad/smem attr"));
  const int n_tiles = (int)ceil_div<int64_t>(n_samples, 128);
  // ~kWgItemsPerSm equal-cost work items per SM (few enough to keep the atomic traffic of the epilogues small): slabs per job
  // proportional to the bytes per sample it streams, then the largest jobs take up the rounding slack so that the
  // item count is a multiple of the grid
  const int want = kWgItemsPerSm * sms;
  double cost_sum = 0.0;
  for (int j = 0; j < n_jobs; ++j) cost_sum += jobs_host[j].f_dz + jobs_host[j].n;
  int slabs[kMaxWgradJobs], total = 0;
  for (int j = 0; j < n_jobs; ++j) {
    const double share = (jobs_host[j].f_dz + jobs_host[j].n) / cost_sum * want;
    slabs[j] = std::max(1, std::min(n_tiles, (int)(share + 0.5)));
    total += slabs[j];
  }
  const int grid = std::min(total, sms);
  for (int guard = 0; total % grid != 0 && guard < 4 * kMaxWgradJobs; ++guard) {
    // add slabs to the jobs with the most samples per slab until the count is a multiple of the grid
    int best = -1;
    double best_load = 0.0;
    for (int j = 0; j < n_jobs; ++j) {
      const double load = (double)(jobs_host[j].f_dz + jobs_host[j].n) / slabs[j];
      if (slabs[j] < n_tiles && load > best_load) { best_load = load; best = j; }
    }
    if (best < 0) break;
    ++slabs[best];
    ++total;
  }
  tab.slab0[0] = 0;
  for (int j = 0; j < n_jobs; ++j) tab.slab0[j + 1] = tab.slab0[j] + slabs[j];
  tile_wgrad_kernel<<<grid, kWgThreads, kWgSmem, st>>>(tab, n_jobs, n_tiles, n_samples_dev);
  return launched("tile_wgrad");
}

static int check_job(const hn_wgrad_job& j) {
  if (!j.dz || !j.x || !j.dw) return fail_arg("hn_tile_wgrad: null pointer in job");
  if (j.f_dz % 16 || j.f_dz < 16 || j.f_dz > 256) return fail_arg("hn_tile_wgrad: f_dz=%d must be a multiple of 16 in [16,256]", j.f_dz);
  if (j.f_x % 16 || j.f_x < 16 || j.f_x > 256) return fail_arg("hn_tile_wgrad: f_x=%d must be a multiple of 16 in [16,256]", j.f_x);
  if (j.n % 16 || j.n < 16 || j.n > 256 || j.x_f0 % 8 || j.x_f0 + j.n > j.f_x)
    return fail_arg("hn_tile_wgrad: bad X window (x_f0=%d, n=%d, f_x=%d)", j.x_f0, j.n, j.f_x);
  if (j.m < 1 || j.m > j.f_dz || j.n_valid < 0 || j.n_valid > j.n) return fail_arg("hn_tile_wgrad: bad m / n_valid");
  return 0;
}

}  // namespace hn

using namespace hn;

extern "C" {

int64_t hn_tile_image_bytes(int64_t n_samples, int n_features_padded) {
  if (n_samples < 0 || n_features_padded <= 0 || n_features_padded % 8) return -1;
  return ceil_div<int64_t>(n_samples, 128) * 128 * n_features_padded * 2;
}

int hn_tile_pack(int64_t n_samples, int n_features, int n_features_padded, const float* src, int64_t ld,
                 void* image, void* stream) {
  if (n_samples == 0) return 0;
  if (n_samples < 0 || n_features <= 0 || n_features_padded % 8 || n_features > n_features_padded || !src || !image)
    return fail_arg("hn_tile_pack: bad argument");
  tile_pack_kernel<<<1024, 256, 0, as_stream(stream)>>>(n_samples, n_features, n_features_padded, src, ld,
                                                        static_cast<uint8_t*>(image));
  return launched("hn_tile_pack");
}

int hn_tile_unpack(int64_t n_samples, int n_features, int n_features_padded, const void* image, float* dst,
                   int64_t ld, void* stream) {
  if (n_samples == 0) return 0;
  if (n_samples < 0 || n_features <= 0 || n_features_padded % 8 || n_features > n_features_padded || !dst || !image)
    return fail_arg("hn_tile_unpack: bad argument");
  tile_unpack_kernel<<<1024, 256, 0, as_stream(stream)>>>(n_samples, n_features, n_features_padded,
                                                          static_cast<const uint8_t*>(image), dst, ld);
  return launched("hn_tile_unpack");
}

int hn_tile_wgrad(const hn_wgrad_job* jobs_host, int n_jobs, int64_t n_samples, void* jobs_scratch_dev,
                  void* stream) {
  if (n_jobs == 0 || n_samples == 0) return 0;
  (void)jobs_scratch_dev;
  if (n_jobs < 0 || n_samples < 0 || !jobs_host) return fail_arg("hn_tile_wgrad: bad argument");
  for (int i = 0; i < n_jobs; ++i) HN_TRY(check_job(jobs_host[i]));
  cudaStream_t st = as_stream(stream);
  for (int i = 0; i < n_jobs; i += kMaxWgradJobs)
    HN_TRY(launch_tile_wgrad(jobs_host + i, std::min(kMaxWgradJobs, n_jobs - i), n_samples, nullptr, st));
  return 0;
}

}  // extern "C"
