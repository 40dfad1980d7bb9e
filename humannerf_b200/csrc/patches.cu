// a3: patch selection on the device (reference core/data/human_nerf/train.py:235-340 get_patch_ray_indices /
// _get_patch_ray_indices, :226-233 select_rays, :450-476 sample_patch_rays).
//
// The reference picks, per patch, (1) `np.random.rand(1)[0] < sample_subject_ratio` -> candidates = subject mask or
// bbox-minus-subject mask (:263-266), (2) `np.random.choice(n_candidates)` -> the k-th candidate pixel in row-major
// order is the patch centre (:303-309), then clips the window to the image (:312-321), intersects it with the ray mask
// (:330), maps the hit pixels to compacted ray ids with cumsum(ray_mask) - 1 (:333-334) and crops the target (:463-468).
// Here the two host randoms per patch (u, k - the only thing the host still produces per step: 16 bytes per patch) are
// consumed by ONE kernel that writes everything the train step reads: select_inds, patch masks, div indices, window
// origins, target crops and the gathered, zero-padded ray batch with its scatter index.  Integer / index work,
// bit-exact against the reference's outputs for the same draws (tests/golden/patches.npz).
//
//   hn_patch_prepare     per frame: inclusive prefix counts of the ray / subject / background masks
//   hn_select_patches    per step : one block, one warp per patch for the k-th-candidate search (32-ary search in
//                                   the prefix counts), then the block walks the windows with a running scan
#include "common.cuh"

namespace hn {

constexpr int kPatchThreads = 1024;
constexpr int kMaxPatches = 32;

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// block-wide inclusive scan of one int per thread (1024 threads); returns the inclusive value, *total = block sum
__device__ __forceinline__ int block_incl_scan(int v, int* warp_sums, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int inc = warp_incl_scan(v, lane);
  __syncthreads();                       // warp_sums may still be read from the previous call
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int s = warp_incl_scan(warp_sums[lane], lane);
    warp_sums[lane] = s;
  }
  __syncthreads();
  *total = warp_sums[31];
  return inc + (warp > 0 ? warp_sums[warp - 1] : 0);
}

// one block walks the image in tiles of 1024 pixels (coalesced), carrying the three running counts
__global__ void __launch_bounds__(kPatchThreads)
patch_prepare_kernel(int n_pix, const uint8_t* __restrict__ ray_mask, const uint8_t* __restrict__ subject,
                     int32_t* __restrict__ ray_cum, int32_t* __restrict__ subj_cum, int32_t* __restrict__ bg_cum,
                     int32_t* __restrict__ counts) {
  __shared__ int ws[3][32];
  int carry_r = 0, carry_s = 0, carry_b = 0;
  for (int base = 0; base < n_pix; base += kPatchThreads) {
    const int i = base + threadIdx.x;
    const int r = (i < n_pix && ray_mask[i]) ? 1 : 0;
    const int s = (i < n_pix && subject[i]) ? 1 : 0;
    const int b = (r && !s) ? 1 : 0;                        // np.bitwise_and(bbox_mask, np.bitwise_not(subject_mask))
    int tr, ts, tb;
    const int ir = block_incl_scan(r, ws[0], &tr);
    const int is = block_incl_scan(s, ws[1], &ts);
    const int ib = block_incl_scan(b, ws[2], &tb);
    if (i < n_pix) {
      ray_cum[i] = carry_r + ir;
      subj_cum[i] = carry_s + is;
      bg_cum[i] = carry_b + ib;
    }
    carry_r += tr; carry_s += ts; carry_b += tb;
  }
  if (threadIdx.x == 0) { counts[0] = carry_r; counts[1] = carry_s; counts[2] = carry_b; }
}

// smallest pixel with cum[pixel] >= want (want >= 1): 32-ary search by one warp; n_pix if there is none
__device__ __forceinline__ int kth_set_pixel(const int32_t* __restrict__ cum, int n_pix, int want, int lane) {
  int lo = 0, hi = n_pix;                                   // answer in [lo, hi]; cum[hi - 1] >= want assumed checked
  while (hi - lo > 1) {
    const int step = (hi - lo + 31) / 32;
    int idx = lo + (lane + 1) * step - 1;
    if (idx > hi - 1) idx = hi - 1;
    const bool ge = __ldg(cum + idx) >= want;
    const unsigned m = __ballot_sync(0xffffffffu, ge);
    if (m == 0u) return n_pix;
    const int first = __ffs(m) - 1;
    int new_hi = lo + (first + 1) * step - 1;
    if (new_hi > hi - 1) new_hi = hi - 1;
    const int new_lo = first == 0 ? lo : lo + first * step;  // the probe before `first` was < want
    lo = new_lo;
    hi = new_hi + 1;
  }
  return lo;
}

__global__ void __launch_bounds__(kPatchThreads)
select_patches_kernel(int H, int W, int n_patches, int P, double ratio, const uint8_t* __restrict__ ray_mask,
                      const int32_t* __restrict__ ray_cum, const int32_t* __restrict__ subj_cum,
                      const int32_t* __restrict__ bg_cum, const double* __restrict__ u, const int64_t* __restrict__ choice,
                      const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ near,
                      const float* __restrict__ far, const float* __restrict__ image, int32_t* __restrict__ select_inds,
                      uint8_t* __restrict__ patch_masks, int32_t* __restrict__ div_indices, int32_t* __restrict__ xy_min,
                      float* __restrict__ targets, float* __restrict__ out_rays, float* __restrict__ out_near,
                      float* __restrict__ out_far, int64_t* __restrict__ scatter_index, int64_t* __restrict__ n_rays) {
  __shared__ int ws[32];
  __shared__ int s_x0[kMaxPatches], s_y0[kMaxPatches];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_pix = H * W;
  const int cap = n_patches * P * P;
  // ---- phase 1: warp p finds patch p's centre = the choice[p]-th candidate pixel in row-major order
  if (warp < n_patches) {
    const bool on_subject = u[warp] < ratio;                                       // train.py:263
    const int32_t* cum = on_subject ? subj_cum : bg_cum;
    const int n_cand = __ldg(cum + n_pix - 1);
    long long k = choice[warp];
    if (k < 0) k = 0;
    if (k > (long long)n_cand - 1) k = (long long)n_cand - 1;                     // host contract: 0 <= k < n_cand
    int centre = n_cand > 0 ? kth_set_pixel(cum, n_pix, (int)k + 1, lane) : 0;
    if (centre >= n_pix) centre = n_pix - 1;
    const int cy = centre / W, cx = centre - cy * W;
    if (lane == 0) {
      const int half = P / 2;
      s_x0[warp] = min(max(cx - half, 0), W - P);                                 // np.clip(center - half, 0, W - patch)
      s_y0[warp] = min(max(cy - half, 0), H - P);
      xy_min[warp * 2 + 0] = s_x0[warp];
      xy_min[warp * 2 + 1] = s_y0[warp];
    }
  }
  __syncthreads();
  // ---- phase 2: walk the windows in row-major order; `running` = rays selected so far (patch_div_indices)
  int running = 0;
  if (threadIdx.x == 0) div_indices[0] = 0;
  for (int p = 0; p < n_patches; ++p) {
    const int x0 = s_x0[p], y0 = s_y0[p];
    for (int base = 0; base < P * P; base += kPatchThreads) {
      const int t = base + threadIdx.x;
      const bool in = t < P * P;
      const int wy = in ? t / P : 0, wx = in ? t - wy * P : 0;
      const int pix = (y0 + wy) * W + x0 + wx;
      const int hit = (in && ray_mask[pix]) ? 1 : 0;                              // inter_mask (:330)
      int total;
      const int inc = block_incl_scan(hit, ws, &total);
      if (in) {
        patch_masks[(size_t)p * P * P + t] = (uint8_t)hit;
        const float* px = image + (size_t)pix * 3;
        float* tg = targets + ((size_t)p * P * P + t) * 3;
        tg[0] = px[0]; tg[1] = px[1]; tg[2] = px[2];                              // img[y_min:y_max, x_min:x_max] (:463-468)
      }
      if (hit) {
        const int row = running + inc - 1;
        const int ray = __ldg(ray_cum + pix) - 1;                                 // np.cumsum(ray_mask) - 1 (:333)
        select_inds[row] = ray;
        scatter_index[row] = (int64_t)p * P * P + t;                              // patch_imgs[i, patch_masks[i]] = rgbs[...]
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float o = rays_o[(size_t)ray * 3 + c], d = rays_d[(size_t)ray * 3 + c];
          out_rays[((size_t)0 * cap + row) * 3 + c] = o;
          out_rays[((size_t)1 * cap + row) * 3 + c] = d;
          out_rays[((size_t)2 * cap + row) * 3 + c] = d;
        }
        out_near[row] = near[ray];
        out_far[row] = far[ray];
      }
      running += total;
    }
    if (threadIdx.x == 0) div_indices[p + 1] = running;
  }
  // ---- padding rows: zero rays, near = far = 0, scatter index -1 (skipped by the scatter: no gradient)
  for (int row = running + threadIdx.x; row < cap; row += kPatchThreads) {
    select_inds[row] = -1;
    scatter_index[row] = -1;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      out_rays[((size_t)0 * cap + row) * 3 + c] = 0.f;
      out_rays[((size_t)1 * cap + row) * 3 + c] = 0.f;
      out_rays[((size_t)2 * cap + row) * 3 + c] = 0.f;
    }
    out_near[row] = 0.f;
    out_far[row] = 0.f;
  }
  if (threadIdx.x == 0 && n_rays) *n_rays = running;
}

}  // namespace hn

using namespace hn;

extern "C" {

int hn_patch_prepare(int H, int W, const uint8_t* ray_mask, const uint8_t* subject_mask, int32_t* ray_cum,
                     int32_t* subject_cum, int32_t* background_cum, int32_t* counts, void* stream) {
  if (H <= 0 || W <= 0 || (int64_t)H * W > (int64_t)1 << 30) return fail_arg("hn_patch_prepare: bad image size %d x %d", H, W);
  if (!ray_mask || !subject_mask || !ray_cum || !subject_cum || !background_cum || !counts)
    return fail_arg("hn_patch_prepare: null pointer");
  patch_prepare_kernel<<<1, kPatchThreads, 0, as_stream(stream)>>>(H * W, ray_mask, subject_mask, ray_cum, subject_cum,
                                                                  background_cum, counts);
  return launched("hn_patch_prepare");
}

int hn_select_patches(int H, int W, int n_patches, int patch_size, double subject_ratio, const uint8_t* ray_mask,
                      const int32_t* ray_cum, const int32_t* subject_cum, const int32_t* background_cum,
                      const double* u, const int64_t* choice, const float* rays_o, const float* rays_d,
                      const float* near, const float* far, const float* image, int32_t* select_inds,
                      uint8_t* patch_masks, int32_t* div_indices, int32_t* xy_min, float* target_patches,
                      float* out_rays, float* out_near, float* out_far, int64_t* scatter_index, int64_t* n_rays,
                      void* stream) {
  if (n_patches == 0) return 0;
  if (H <= 0 || W <= 0 || n_patches < 0 || n_patches > kMaxPatches || patch_size <= 0 || patch_size > H || patch_size > W)
    return fail_arg("hn_select_patches: bad geometry (H %d, W %d, %d patches of %d; at most %d patches, patch <= image)",
                    H, W, n_patches, patch_size, kMaxPatches);
  if (!ray_mask || !ray_cum || !subject_cum || !background_cum || !u || !choice || !rays_o || !rays_d || !near || !far ||
      !image || !select_inds || !patch_masks || !div_indices || !xy_min || !target_patches || !out_rays || !out_near ||
      !out_far || !scatter_index)
    return fail_arg("hn_select_patches: null pointer");
  select_patches_kernel<<<1, kPatchThreads, 0, as_stream(stream)>>>(
      H, W, n_patches, patch_size, subject_ratio, ray_mask, ray_cum, subject_cum, background_cum, u, choice, rays_o,
      rays_d, near, far, image, select_inds, patch_masks, div_indices, xy_min, target_patches, out_rays, out_near,
      out_far, scatter_index, n_rays);
  return launched("hn_select_patches");
}

}  // extern "C"
