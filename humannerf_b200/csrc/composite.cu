// a15: alpha compositing (reference network.py:355-388 _raw2outputs), forward and backward.
// One warp per ray; lane l owns the contiguous samples [l*SPL, (l+1)*SPL); the exclusive
// transmittance product and the backward recurrence are warp-shuffle scans.
#include "common.cuh"

namespace hn {

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

// Per-lane forward quantities of one ray.
template <int SPL>
struct RayLane {
  float alpha[SPL];
  float trans[SPL];   // exclusive transmittance T_s
  float zz[SPL];
  float dist[SPL];    // delta_s * |d|
  float e[SPL];       // exp(-relu(sigma) * dist)
  float4 raw[SPL];
  float m[SPL];
};

template <int SPL>
__device__ __forceinline__ void ray_forward(int64_t ray, int S, int lane, const float* __restrict__ raw,
                                            const float* __restrict__ mask, const float* __restrict__ z,
                                            const float* __restrict__ rays_d, RayLane<SPL>& L) {
  const int s0 = lane * SPL;
  const float4* raw4 = reinterpret_cast<const float4*>(raw) + ray * S;
#pragma unroll
  for (int k = 0; k < SPL; ++k) {
    int s = s0 + k;
    bool in = s < S;
    L.raw[k] = in ? raw4[s] : make_float4(0.f, 0.f, 0.f, 0.f);
    L.m[k] = in ? mask[ray * S + s] : 0.0f;
    L.zz[k] = in ? z[ray * S + s] : 0.0f;
  }
  float dx = rays_d[ray * 3 + 0], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
  float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
  // z of the next lane's first sample closes this lane's last interval
  float z_next_lane = __shfl_down_sync(0xffffffffu, L.zz[0], 1);
  float prod = 1.0f;
#pragma unroll
  for (int k = 0; k < SPL; ++k) {
    int s = s0 + k;
    float zn = (k + 1 < SPL) ? L.zz[(k + 1 < SPL) ? k + 1 : k] : z_next_lane;
    float delta = (s >= S - 1) ? 1e10f : (zn - L.zz[k]);
    L.dist[k] = delta * dnorm;
    float sig = fmaxf(L.raw[k].w, 0.0f);
    L.e[k] = expf(-sig * L.dist[k]);
    float a = (1.0f - L.e[k]) * L.m[k];
    if (s >= S) a = 0.0f;
    L.alpha[k] = a;
    L.trans[k] = prod;                       // lane-local exclusive product for now
    if (s < S) prod *= __fadd_rn(__fsub_rn(1.0f, a), 1e-10f);
  }
  // exclusive multiplicative scan of the lane products
  float incl = prod;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl *= t;
  }
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 1.0f;
#pragma unroll
  for (int k = 0; k < SPL; ++k) L.trans[k] *= excl;
}

template <int SPL>
__global__ void __launch_bounds__(128)
composite_fwd_kernel(int64_t n_rays, int S, const float* __restrict__ raw, const float* __restrict__ mask,
                     const float* __restrict__ z, const float* __restrict__ rays_d,
                     const float* __restrict__ xyz, const float* __restrict__ bgcolor,
                     float* __restrict__ rgb, float* __restrict__ acc, float* __restrict__ depth,
                     float* __restrict__ weights, float* __restrict__ cnl_xyz,
                     float* __restrict__ cnl_rgb, float* __restrict__ w_max,
                     float* __restrict__ rgb_on_rays) {
  const int lane = threadIdx.x & 31;
  int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  RayLane<SPL> L;
  ray_forward<SPL>(ray, S, lane, raw, mask, z, rays_d, L);
  float sr = 0.f, sg = 0.f, sb = 0.f, sd = 0.f, sa = 0.f;
  float best = -1.0f;      // weights are >= 0 unless mask > 1 pushes 1-alpha negative; see below
  int best_s = 0x7fffffff;
  bool have = false;
  const int s0 = lane * SPL;
#pragma unroll
  for (int k = 0; k < SPL; ++k) {
    int s = s0 + k;
    if (s >= S) continue;
    float w = L.alpha[k] * L.trans[k];
    float cr = sigmoidf(L.raw[k].x), cg = sigmoidf(L.raw[k].y), cb = sigmoidf(L.raw[k].z);
    sr += w * cr; sg += w * cg; sb += w * cb;
    sd += w * L.zz[k];
    sa += w;
    if (!have || w > best) { best = w; best_s = s; have = true; }
    if (weights) weights[ray * S + s] = w;
    if (rgb_on_rays) {
      float* o = rgb_on_rays + (ray * S + s) * 3;
      o[0] = cr; o[1] = cg; o[2] = cb;
    }
  }
  sr = warp_sum(sr); sg = warp_sum(sg); sb = warp_sum(sb); sd = warp_sum(sd); sa = warp_sum(sa);
  // arg-max with first-index tie break (torch.max on CPU)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ob = __shfl_xor_sync(0xffffffffu, best, o);
    int os = __shfl_xor_sync(0xffffffffu, best_s, o);
    bool oh = __shfl_xor_sync(0xffffffffu, have ? 1 : 0, o) != 0;
    if (oh && (!have || ob > best || (ob == best && os < best_s))) { best = ob; best_s = os; have = true; }
  }
  if (lane == 0) {
    float rest = 1.0f - sa;
    rgb[ray * 3 + 0] = sr + rest * (bgcolor[0] / 255.0f);
    rgb[ray * 3 + 1] = sg + rest * (bgcolor[1] / 255.0f);
    rgb[ray * 3 + 2] = sb + rest * (bgcolor[2] / 255.0f);
    acc[ray] = sa;
    depth[ray] = sd;
    if (w_max) w_max[ray] = best;
    const float4 r4 = reinterpret_cast<const float4*>(raw)[ray * S + best_s];
    if (cnl_rgb) {
      cnl_rgb[ray * 3 + 0] = sigmoidf(r4.x);
      cnl_rgb[ray * 3 + 1] = sigmoidf(r4.y);
      cnl_rgb[ray * 3 + 2] = sigmoidf(r4.z);
    }
    if (cnl_xyz && xyz) {
      const float* q = xyz + (ray * S + best_s) * 3;
      cnl_xyz[ray * 3 + 0] = q[0]; cnl_xyz[ray * 3 + 1] = q[1]; cnl_xyz[ray * 3 + 2] = q[2];
    }
  }
}

// Backward.  With b_s = 1 - alpha_s + 1e-10, T_s = prod_{i<s} b_i, w_s = alpha_s T_s and
// G_s = dL/dw_s = d_rgb . (rgb_s - bg) + d_depth z_s + d_acc + d_w_s, define
//   Q_s = sum_{j>s} G_j alpha_j prod_{s<i<j} b_i,   Q_{S-1} = 0,  Q_{s-1} = G_s alpha_s + b_s Q_s.
// Then dL/dalpha_s = T_s (G_s - Q_s): no division by b_s, stable when alpha -> 1.
template <int SPL>
__global__ void __launch_bounds__(128)
composite_bwd_kernel(int64_t n_rays, int S, const float* __restrict__ raw, const float* __restrict__ mask,
                     const float* __restrict__ z, const float* __restrict__ rays_d,
                     const float* __restrict__ bgcolor, const float* __restrict__ d_rgb,
                     const float* __restrict__ d_acc, const float* __restrict__ d_depth,
                     const float* __restrict__ d_weights, float* __restrict__ d_raw,
                     float* __restrict__ d_mask) {
  const int lane = threadIdx.x & 31;
  int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= n_rays) return;
  RayLane<SPL> L;
  ray_forward<SPL>(ray, S, lane, raw, mask, z, rays_d, L);
  float gr = 0.f, gg = 0.f, gb = 0.f, ga = 0.f, gd = 0.f;
  if (d_rgb) { gr = d_rgb[ray * 3 + 0]; gg = d_rgb[ray * 3 + 1]; gb = d_rgb[ray * 3 + 2]; }
  if (d_acc) ga = d_acc[ray];
  if (d_depth) gd = d_depth[ray];
  float bg_dot = gr * (bgcolor[0] / 255.0f) + gg * (bgcolor[1] / 255.0f) + gb * (bgcolor[2] / 255.0f);
  const int s0 = lane * SPL;
  float G[SPL], cr[SPL], cg[SPL], cb[SPL];
  // lane-local composition of Q_{first-1} = A + B * Q_{last}
  float A = 0.0f, B = 1.0f;
#pragma unroll
  for (int k = SPL - 1; k >= 0; --k) {
    int s = s0 + k;
    cr[k] = sigmoidf(L.raw[k].x); cg[k] = sigmoidf(L.raw[k].y); cb[k] = sigmoidf(L.raw[k].z);
    float g = gr * cr[k] + gg * cg[k] + gb * cb[k] - bg_dot + gd * L.zz[k] + ga;
    if (d_weights && s < S) g += d_weights[ray * S + s];
    if (s >= S) g = 0.0f;
    G[k] = g;
    float a_s = g * L.alpha[k];
    float b_s = (s < S) ? __fadd_rn(__fsub_rn(1.0f, L.alpha[k]), 1e-10f) : 1.0f;
    // prepend f_s to the running map: new(q) = f_s(old(q)) ... applied from the back, so
    // running map so far covers samples k+1..SPL-1; f_s o running:
    A = a_s + b_s * A;
    B = b_s * B;
  }
  // inclusive suffix scan over lanes: F_l = f_l o f_{l+1} o ... o f_31
  float FA = A, FB = B;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float oa = __shfl_down_sync(0xffffffffu, FA, o);
    float ob = __shfl_down_sync(0xffffffffu, FB, o);
    if (lane + o < 32) { FA = FA + FB * oa; FB = FB * ob; }
  }
  // Q after this lane's last sample = F_{l+1}(0) = FA of the next lane (0 for the last lane)
  float q = __shfl_down_sync(0xffffffffu, FA, 1);
  if (lane == 31) q = 0.0f;
#pragma unroll
  for (int k = SPL - 1; k >= 0; --k) {
    int s = s0 + k;
    if (s < S) {
      float dalpha = L.trans[k] * (G[k] - q);
      float w = L.alpha[k] * L.trans[k];
      float sig = L.raw[k].w;
      float dsig = (sig > 0.0f) ? dalpha * L.m[k] * L.dist[k] * L.e[k] : 0.0f;
      float4 o;
      o.x = w * gr * cr[k] * (1.0f - cr[k]);
      o.y = w * gg * cg[k] * (1.0f - cg[k]);
      o.z = w * gb * cb[k] * (1.0f - cb[k]);
      o.w = dsig;
      reinterpret_cast<float4*>(d_raw)[ray * S + s] = o;
      if (d_mask) d_mask[ray * S + s] = dalpha * (1.0f - L.e[k]);
      float b_s = __fadd_rn(__fsub_rn(1.0f, L.alpha[k]), 1e-10f);
      q = G[k] * L.alpha[k] + b_s * q;
    }
  }
}

template <int SPL>
static int launch_fwd(int64_t n_rays, int S, const float* raw, const float* mask, const float* z,
                      const float* rays_d, const float* xyz, const float* bg, float* rgb, float* acc,
                      float* depth, float* weights, float* cnl_xyz, float* cnl_rgb, float* w_max,
                      float* rgb_on_rays, cudaStream_t st) {
  unsigned grid = (unsigned)ceil_div<int64_t>(n_rays, 4);
  composite_fwd_kernel<SPL><<<grid, 128, 0, st>>>(n_rays, S, raw, mask, z, rays_d, xyz, bg, rgb, acc, depth,
                                                  weights, cnl_xyz, cnl_rgb, w_max, rgb_on_rays);
  return launched("hn_composite_fwd");
}

template <int SPL>
static int launch_bwd(int64_t n_rays, int S, const float* raw, const float* mask, const float* z,
                      const float* rays_d, const float* bg, const float* d_rgb, const float* d_acc,
                      const float* d_depth, const float* d_weights, float* d_raw, float* d_mask,
                      cudaStream_t st) {
  unsigned grid = (unsigned)ceil_div<int64_t>(n_rays, 4);
  composite_bwd_kernel<SPL><<<grid, 128, 0, st>>>(n_rays, S, raw, mask, z, rays_d, bg, d_rgb, d_acc, d_depth,
                                                  d_weights, d_raw, d_mask);
  return launched("hn_composite_bwd");
}

}  // namespace hn

using namespace hn;

extern "C" {

int hn_composite_fwd(int64_t n_rays, int n_samples, const float* raw, const float* mask, const float* z,
                     const float* rays_d, const float* xyz, const float* bgcolor, float* rgb, float* acc,
                     float* depth, float* weights, float* cnl_xyz, float* cnl_rgb, float* w_max,
                     float* rgb_on_rays, void* stream) {
  if (n_rays < 0 || n_samples <= 0 || n_samples > 512)
    return fail_arg("hn_composite_fwd: n_samples=%d outside [1,512]", n_samples);
  if (n_rays == 0) return 0;
  if (!raw || !mask || !z || !rays_d || !bgcolor || !rgb || !acc || !depth)
    return fail_arg("hn_composite_fwd: null required pointer");
  if (cnl_xyz && !xyz) return fail_arg("hn_composite_fwd: cnl_xyz requested without xyz");
  if (n_rays == 0) return 0;
  cudaStream_t st = as_stream(stream);
  int spl = (n_samples + 31) / 32;
#define HN_GO(N) return launch_fwd<N>(n_rays, n_samples, raw, mask, z, rays_d, xyz, bgcolor, rgb, acc, depth, \
                                      weights, cnl_xyz, cnl_rgb, w_max, rgb_on_rays, st)
  if (spl <= 1) HN_GO(1);
  if (spl <= 2) HN_GO(2);
  if (spl <= 4) HN_GO(4);
  if (spl <= 8) HN_GO(8);
  HN_GO(16);
#undef HN_GO
}

int hn_composite_bwd(int64_t n_rays, int n_samples, const float* raw, const float* mask, const float* z,
                     const float* rays_d, const float* bgcolor, const float* d_rgb, const float* d_acc,
                     const float* d_depth, const float* d_weights, float* d_raw, float* d_mask,
                     void* stream) {
  if (n_rays < 0 || n_samples <= 0 || n_samples > 512)
    return fail_arg("hn_composite_bwd: n_samples=%d outside [1,512]", n_samples);
  if (n_rays == 0) return 0;
  if (!raw || !mask || !z || !rays_d || !bgcolor || !d_raw)
    return fail_arg("hn_composite_bwd: null required pointer");
  if (n_rays == 0) return 0;
  cudaStream_t st = as_stream(stream);
  int spl = (n_samples + 31) / 32;
#define HN_GO(N) return launch_bwd<N>(n_rays, n_samples, raw, mask, z, rays_d, bgcolor, d_rgb, d_acc, d_depth, \
                                      d_weights, d_raw, d_mask, st)
  if (spl <= 1) HN_GO(1);
  if (spl <= 2) HN_GO(2);
  if (spl <= 4) HN_GO(4);
  if (spl <= 8) HN_GO(8);
  HN_GO(16);
#undef HN_GO
}

}  // extern "C"
