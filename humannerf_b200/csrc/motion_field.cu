// a10: motion field = 24-channel trilinear lookup of the blend-weight volume at the per-joint
// warped positions + inverse linear-blend-skinning (reference network.py:392-444), forward and
// backward.  One thread per sample; the per-joint transforms live in shared memory; the volume
// (J * V^3 floats = 3.1 MB at V = 32) is read through L1/L2 - a (z, y) row of V = 32 floats is
// exactly one 128-byte line.  Joints whose warped position lies a full cell outside the volume
// contribute an exact zero and skip their eight loads.
#include <algorithm>

#include "common.cuh"

namespace hn {

struct Trilinear {
  int x0, y0, z0;
  float tx, ty, tz;
  bool any;   // at least one corner inside the volume
};

// grid_sample, align_corners=True: ix = ((g + 1) / 2) * (V - 1)   (unnormalise, ATen GridSampler.h)
__device__ __forceinline__ Trilinear make_trilinear(float gx, float gy, float gz, int V) {
  float fx = ((gx + 1.0f) * 0.5f) * (float)(V - 1);
  float fy = ((gy + 1.0f) * 0.5f) * (float)(V - 1);
  float fz = ((gz + 1.0f) * 0.5f) * (float)(V - 1);
  Trilinear t;
  // every corner is out of range when f <= -1 or f >= V (also rejects NaN / inf)
  t.any = (fx > -1.0f) && (fx < (float)V) && (fy > -1.0f) && (fy < (float)V) &&
          (fz > -1.0f) && (fz < (float)V);
  float flx = floorf(fx), fly = floorf(fy), flz = floorf(fz);
  t.tx = fx - flx; t.ty = fy - fly; t.tz = fz - flz;
  t.x0 = t.any ? (int)flx : 0; t.y0 = t.any ? (int)fly : 0; t.z0 = t.any ? (int)flz : 0;
  return t;
}

// loads the 8 corners (zero outside), c[dz*4 + dy*2 + dx]
__device__ __forceinline__ void load_corners(const float* __restrict__ vj, int V, const Trilinear& t,
                                             float c[8]) {
#pragma unroll
  for (int dz = 0; dz < 2; ++dz) {
    int zi = t.z0 + dz;
    bool zin = zi >= 0 && zi < V;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      int yi = t.y0 + dy;
      bool yin = zin && yi >= 0 && yi < V;
      const float* row = vj + ((int64_t)zi * V + yi) * V;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        int xi = t.x0 + dx;
        bool in = yin && xi >= 0 && xi < V;
        c[dz * 4 + dy * 2 + dx] = in ? __ldg(row + xi) : 0.0f;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Corner table: for every joint and every cell (z0, y0, x0) in [-1, V-1]^3 the 8 corner values of the
// trilinear stencil, zero where a corner falls outside the volume, stored as 8 consecutive floats
// (one 32-byte sector).  A lookup is then two 16-byte loads and no per-corner bounds test, instead of
// 8 scalar loads from up to 4 different 128-byte lines.  J * (V+1)^3 * 32 B = 27.6 MB at J = 24,
// V = 32: resident in the 126 MB L2.  Rebuilt per call from the (J,V,V,V) volume (~10 us).
// ------------------------------------------------------------------------------------------
__global__ void pack_corners_kernel(int J, int V, const float* __restrict__ vol, float4* __restrict__ table) {
  const int C = V + 1;
  const int64_t cells = (int64_t)J * C * C * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int cx = (int)(r % C); r /= C;
    const int cy = (int)(r % C); r /= C;
    const int cz = (int)(r % C);
    const int j = (int)(r / C);
    const float* vj = vol + (int64_t)j * V * V * V;
    float c[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int zi = cz - 1 + (k >> 2), yi = cy - 1 + ((k >> 1) & 1), xi = cx - 1 + (k & 1);
      const bool in = zi >= 0 && zi < V && yi >= 0 && yi < V && xi >= 0 && xi < V;
      c[k] = in ? __ldg(vj + ((int64_t)zi * V + yi) * V + xi) : 0.0f;
    }
    table[2 * i] = make_float4(c[0], c[1], c[2], c[3]);
    table[2 * i + 1] = make_float4(c[4], c[5], c[6], c[7]);
  }
}

// corners of the cell of `t` for joint j from the table (t.any must hold)
__device__ __forceinline__ void load_corners_packed(const float4* __restrict__ table, int j, int V, const Trilinear& t,
                                                    float c[8]) {
  const int C = V + 1;
  const int64_t cell = (((int64_t)j * C + (t.z0 + 1)) * C + (t.y0 + 1)) * C + (t.x0 + 1);
  const float4 a = __ldg(table + 2 * cell), b = __ldg(table + 2 * cell + 1);
  c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
}

__device__ __forceinline__ float interp(const float c[8], const Trilinear& t) {
  float ux = 1.0f - t.tx, uy = 1.0f - t.ty, uz = 1.0f - t.tz;
  float w = 0.0f;
  w += c[0] * (ux * uy * uz); w += c[1] * (t.tx * uy * uz);
  w += c[2] * (ux * t.ty * uz); w += c[3] * (t.tx * t.ty * uz);
  w += c[4] * (ux * uy * t.tz); w += c[5] * (t.tx * uy * t.tz);
  w += c[6] * (ux * t.ty * t.tz); w += c[7] * (t.tx * t.ty * t.tz);
  return w;
}

__device__ __forceinline__ void load_point(int64_t i, int S, const float* __restrict__ pts,
                                           const float* __restrict__ rays_o,
                                           const float* __restrict__ rays_d,
                                           const float* __restrict__ z, float p[3]) {
  if (pts != nullptr) {
    p[0] = pts[i * 3 + 0]; p[1] = pts[i * 3 + 1]; p[2] = pts[i * 3 + 2];
  } else {
    int64_t r = i / S;
    float zz = z[i];
    // pts = o + d * z with separate roundings (eager reference, network.py:499)
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = __fadd_rn(rays_o[r * 3 + a], __fmul_rn(rays_d[r * 3 + a], zz));
  }
}

struct FieldConsts {
  float bmin[3], bscale[3];
};

template <bool kWriteWeights>
__global__ void __launch_bounds__(256)
motion_field_fwd_kernel(int64_t n_pts, int S, int J, int V, const float* __restrict__ pts,
                        const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                        const float* __restrict__ z, const float* __restrict__ Rs,
                        const float* __restrict__ Ts, const float* __restrict__ vol,
                        const float* __restrict__ bbox_min, const float* __restrict__ bbox_scale,
                        float* __restrict__ x_skel, float* __restrict__ mask,
                        float* __restrict__ weights, const float4* __restrict__ table) {
  __shared__ float s_R[kMaxJoints * 9];
  __shared__ float s_T[kMaxJoints * 3];
  for (int k = threadIdx.x; k < J * 9; k += blockDim.x) s_R[k] = Rs[k];
  for (int k = threadIdx.x; k < J * 3; k += blockDim.x) s_T[k] = Ts[k];
  __syncthreads();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pts) return;
  float bm[3] = {__ldg(bbox_min), __ldg(bbox_min + 1), __ldg(bbox_min + 2)};
  float bs[3] = {__ldg(bbox_scale), __ldg(bbox_scale + 1), __ldg(bbox_scale + 2)};
  float p[3];
  load_point(i, S, pts, rays_o, rays_d, z, p);
  float nx = 0.f, ny = 0.f, nz = 0.f, wsum = 0.f;
  const int64_t vstride = (int64_t)V * V * V;
  for (int j = 0; j < J; ++j) {
    const float* R = s_R + j * 9;
    float px = R[0] * p[0] + R[1] * p[1] + R[2] * p[2] + s_T[j * 3 + 0];
    float py = R[3] * p[0] + R[4] * p[1] + R[5] * p[2] + s_T[j * 3 + 1];
    float pz = R[6] * p[0] + R[7] * p[1] + R[8] * p[2] + s_T[j * 3 + 2];
    Trilinear t = make_trilinear((px - bm[0]) * bs[0] - 1.0f, (py - bm[1]) * bs[1] - 1.0f,
                                 (pz - bm[2]) * bs[2] - 1.0f, V);
    float w = 0.0f;
    if (t.any) {
      float c[8];
      if (table) load_corners_packed(table, j, V, t, c);
      else load_corners(vol + j * vstride, V, t, c);
      w = interp(c, t);
    }
    if (kWriteWeights) weights[i * J + j] = w;
    nx += w * px; ny += w * py; nz += w * pz;
    wsum += w;
  }
  float inv = 1.0f / fmaxf(wsum, 1e-4f);
  x_skel[i * 3 + 0] = nx * inv;
  x_skel[i * 3 + 1] = ny * inv;
  x_skel[i * 3 + 2] = nz * inv;
  mask[i] = wsum;
}

// Backward.  With N = sum_j w_j pos_j, W = sum_j w_j, c = max(W, 1e-4), x = N / c:
//   dN = gx / c ;  dW = gm - [W >= 1e-4] (gx . x) / c ;  dw_j = dW + dN . pos_j + gw_j
//   dpos_j = w_j dN + dw_j * grad_pos(w_j) ;  dvol_j[corner] += dw_j * corner_weight
//   dR_j += dpos_j (x) p ;  dT_j += dpos_j
__global__ void __launch_bounds__(256)
motion_field_bwd_kernel(int64_t n_pts, int S, int J, int V, const float* __restrict__ pts,
                        const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                        const float* __restrict__ z, const float* __restrict__ Rs,
                        const float* __restrict__ Ts, const float* __restrict__ vol,
                        const float* __restrict__ bbox_min, const float* __restrict__ bbox_scale,
                        const float* __restrict__ d_x_skel, const float* __restrict__ d_mask,
                        const float* __restrict__ d_weights, const float* __restrict__ x_skel_fwd,
                        const float* __restrict__ mask_fwd, float* __restrict__ d_vol,
                        float* __restrict__ d_Rs, float* __restrict__ d_Ts, const float4* __restrict__ table,
                        float4* __restrict__ d_table) {
  __shared__ float s_R[kMaxJoints * 9];
  __shared__ float s_T[kMaxJoints * 3];
  __shared__ float s_acc[kMaxJoints * 12];   // per-block dR (9) + dT (3) per joint
  for (int k = threadIdx.x; k < J * 9; k += blockDim.x) s_R[k] = Rs[k];
  for (int k = threadIdx.x; k < J * 3; k += blockDim.x) s_T[k] = Ts[k];
  for (int k = threadIdx.x; k < J * 12; k += blockDim.x) s_acc[k] = 0.0f;
  __syncthreads();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  float bm[3] = {__ldg(bbox_min), __ldg(bbox_min + 1), __ldg(bbox_min + 2)};
  float bs[3] = {__ldg(bbox_scale), __ldg(bbox_scale + 1), __ldg(bbox_scale + 2)};
  float p[3] = {0.f, 0.f, 0.f};
  float gx[3] = {0.f, 0.f, 0.f};
  float gm = 0.0f;
  bool live = i < n_pts;
  if (live) {
    if (d_x_skel) { gx[0] = d_x_skel[i * 3 + 0]; gx[1] = d_x_skel[i * 3 + 1]; gx[2] = d_x_skel[i * 3 + 2]; }
    if (d_mask) gm = d_mask[i];
    // a sample without upstream gradient (e.g. one that empty-space skipping kept out of the MLPs: sigma = 0
    // gives d mask = 0 exactly) contributes nothing: skip its 24 lookups
    if (!d_weights && gx[0] == 0.0f && gx[1] == 0.0f && gx[2] == 0.0f && gm == 0.0f) live = false;
  }
  if (live) load_point(i, S, pts, rays_o, rays_d, z, p);
  const int64_t vstride = (int64_t)V * V * V;
  // pass 1: W and N - skipped when the caller hands back the forward outputs (x = N / max(W, 1e-4), W)
  float nx = 0.f, ny = 0.f, nz = 0.f, wsum = 0.f;
  const bool have_fwd = x_skel_fwd != nullptr && mask_fwd != nullptr;
  if (live && have_fwd) {
    wsum = mask_fwd[i];
    const float c = fmaxf(wsum, 1e-4f);
    nx = x_skel_fwd[i * 3 + 0] * c; ny = x_skel_fwd[i * 3 + 1] * c; nz = x_skel_fwd[i * 3 + 2] * c;
  } else if (live) {
    for (int j = 0; j < J; ++j) {
      const float* R = s_R + j * 9;
      float px = R[0] * p[0] + R[1] * p[1] + R[2] * p[2] + s_T[j * 3 + 0];
      float py = R[3] * p[0] + R[4] * p[1] + R[5] * p[2] + s_T[j * 3 + 1];
      float pz = R[6] * p[0] + R[7] * p[1] + R[8] * p[2] + s_T[j * 3 + 2];
      Trilinear t = make_trilinear((px - bm[0]) * bs[0] - 1.0f, (py - bm[1]) * bs[1] - 1.0f,
                                   (pz - bm[2]) * bs[2] - 1.0f, V);
      if (t.any) {
        float c[8];
        if (table) load_corners_packed(table, j, V, t, c);
        else load_corners(vol + j * vstride, V, t, c);
        float w = interp(c, t);
        nx += w * px; ny += w * py; nz += w * pz; wsum += w;
      }
    }
  }
  float cden = fmaxf(wsum, 1e-4f);
  float inv = 1.0f / cden;
  float dN[3] = {gx[0] * inv, gx[1] * inv, gx[2] * inv};
  float xdotg = (nx * inv) * gx[0] + (ny * inv) * gx[1] + (nz * inv) * gx[2];
  float dW = gm - (wsum >= 1e-4f ? xdotg * inv : 0.0f);
  const float half_v = 0.5f * (float)(V - 1);
  // pass 2: per-joint gradients
  for (int j = 0; j < J; ++j) {
    float dp[3] = {0.f, 0.f, 0.f};
    bool contributes = false;
    if (live) {
      const float* R = s_R + j * 9;
      float px = R[0] * p[0] + R[1] * p[1] + R[2] * p[2] + s_T[j * 3 + 0];
      float py = R[3] * p[0] + R[4] * p[1] + R[5] * p[2] + s_T[j * 3 + 1];
      float pz = R[6] * p[0] + R[7] * p[1] + R[8] * p[2] + s_T[j * 3 + 2];
      Trilinear t = make_trilinear((px - bm[0]) * bs[0] - 1.0f, (py - bm[1]) * bs[1] - 1.0f,
                                   (pz - bm[2]) * bs[2] - 1.0f, V);
      if (t.any) {
        float c[8];
        if (table) load_corners_packed(table, j, V, t, c);
        else load_corners(vol + j * vstride, V, t, c);
        float w = interp(c, t);
        float dw = dW + dN[0] * px + dN[1] * py + dN[2] * pz;
        if (d_weights) dw += d_weights[i * J + j];
        float ux = 1.0f - t.tx, uy = 1.0f - t.ty, uz = 1.0f - t.tz;
        // spatial gradient of the interpolant in voxel units
        float gxv = uz * (uy * (c[1] - c[0]) + t.ty * (c[3] - c[2])) + t.tz * (uy * (c[5] - c[4]) + t.ty * (c[7] - c[6]));
        float gyv = uz * (ux * (c[2] - c[0]) + t.tx * (c[3] - c[1])) + t.tz * (ux * (c[6] - c[4]) + t.tx * (c[7] - c[5]));
        float gzv = uy * (ux * (c[4] - c[0]) + t.tx * (c[5] - c[1])) + t.ty * (ux * (c[6] - c[2]) + t.tx * (c[7] - c[3]));
        dp[0] = w * dN[0] + dw * gxv * (bs[0] * half_v);
        dp[1] = w * dN[1] + dw * gyv * (bs[1] * half_v);
        dp[2] = w * dN[2] + dw * gzv * (bs[2] * half_v);
        contributes = true;
        if (dw != 0.0f && d_table) {
          // gradient of the 8 corners of the cell as two 16-byte vector reductions into the corner-table layout
          // (red.global.add.v4.f32: a quarter of the atomic operations); unpack_corners_grad_kernel folds it into d_vol
          const int C = V + 1;
          const int64_t cell = (((int64_t)j * C + (t.z0 + 1)) * C + (t.y0 + 1)) * C + (t.x0 + 1);
          const float a0 = dw * uz, a1 = dw * t.tz;
          const float b00 = a0 * uy, b01 = a0 * t.ty, b10 = a1 * uy, b11 = a1 * t.ty;
          atomicAdd(d_table + 2 * cell, make_float4(b00 * ux, b00 * t.tx, b01 * ux, b01 * t.tx));
          atomicAdd(d_table + 2 * cell + 1, make_float4(b10 * ux, b10 * t.tx, b11 * ux, b11 * t.tx));
        } else if (dw != 0.0f) {
          float* gj = d_vol + j * vstride;
#pragma unroll
          for (int dz = 0; dz < 2; ++dz) {
            int zi = t.z0 + dz;
            if (zi < 0 || zi >= V) continue;
            float wz = dz ? t.tz : uz;
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
              int yi = t.y0 + dy;
              if (yi < 0 || yi >= V) continue;
              float wzy = wz * (dy ? t.ty : uy);
#pragma unroll
              for (int dx = 0; dx < 2; ++dx) {
                int xi = t.x0 + dx;
                if (xi < 0 || xi >= V) continue;
                atomicAdd(gj + ((int64_t)zi * V + yi) * V + xi, dw * (wzy * (dx ? t.tx : ux)));
              }
            }
          }
        }
      }
    }
    // warp-reduce the 12 transform gradients with 18 shuffles instead of 60: two transposing steps
    // (lanes trade half of their values, so the value count halves as the lane groups merge), then a
    // plain butterfly over the last 3 values; lanes 0, 8, 16, 24 end up with 3 totals each
    if (__any_sync(0xffffffffu, contributes)) {
      float v[12] = {dp[0] * p[0], dp[0] * p[1], dp[0] * p[2], dp[1] * p[0], dp[1] * p[1], dp[1] * p[2],
                     dp[2] * p[0], dp[2] * p[1], dp[2] * p[2], dp[0], dp[1], dp[2]};
      const bool up = (lane & 16) != 0;
      float a[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const float recv = __shfl_xor_sync(0xffffffffu, up ? v[k] : v[k + 6], 16);
        a[k] = (up ? v[k + 6] : v[k]) + recv;
      }
      const bool up2 = (lane & 8) != 0;
      float b[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float recv = __shfl_xor_sync(0xffffffffu, up2 ? a[k] : a[k + 3], 8);
        b[k] = (up2 ? a[k + 3] : a[k]) + recv;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        b[k] += __shfl_xor_sync(0xffffffffu, b[k], 4);
        b[k] += __shfl_xor_sync(0xffffffffu, b[k], 2);
        b[k] += __shfl_xor_sync(0xffffffffu, b[k], 1);
      }
      if ((lane & 7) == 0) {
        const int base = j * 12 + (up ? 6 : 0) + (up2 ? 3 : 0);
#pragma unroll
        for (int k = 0; k < 3; ++k) atomicAdd(&s_acc[base + k], b[k]);
      }
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < J * 12; k += blockDim.x) {
    float v = s_acc[k];
    if (v != 0.0f) {
      int j = k / 12, e = k % 12;
      if (e < 9) atomicAdd(d_Rs + j * 9 + e, v);
      else atomicAdd(d_Ts + j * 3 + (e - 9), v);
    }
  }
}

static int check_common(const char* who, int64_t n_rays, int S, int J, int V, const float* pts,
                        const float* rays_o, const float* rays_d, const float* z, const float* Rs,
                        const float* Ts, const float* vol, const float* bmin, const float* bscale) {
  if (n_rays < 0 || S <= 0) return fail_arg("%s: bad sizes (rays=%lld, samples=%d)", who, (long long)n_rays, S);
  if (n_rays == 0) return 0;
  if (J <= 0 || J > kMaxJoints) return fail_arg("%s: n_joints=%d outside [1,%d]", who, J, kMaxJoints);
  if (V < 2 || V > 1024) return fail_arg("%s: vol_size=%d unsupported", who, V);
  if (!pts && (!rays_o || !rays_d || !z)) return fail_arg("%s: need pts or (rays_o, rays_d, z)", who);
  if (!Rs || !Ts || !vol || !bmin || !bscale) return fail_arg("%s: null frame input", who);
  return 0;
}

}  // namespace hn

using namespace hn;

namespace hn {
// d_vol[j][z][y][x] += the 8 table entries that hold voxel (z, y, x) as a corner: corner k of cell
// (z + 1 - (k >> 2), y + 1 - ((k >> 1) & 1), x + 1 - (k & 1)) (the inverse of pack_corners_kernel)
__global__ void unpack_corners_grad_kernel(int J, int V, const float* __restrict__ d_table, float* __restrict__ d_vol) {
  const int C = V + 1;
  const int64_t voxels = (int64_t)J * V * V * V;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < voxels; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int x = (int)(r % V); r /= V;
    const int y = (int)(r % V); r /= V;
    const int z = (int)(r % V);
    const int j = (int)(r / V);
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int64_t cell = (((int64_t)j * C + (z + 1 - (k >> 2))) * C + (y + 1 - ((k >> 1) & 1))) * C + (x + 1 - (k & 1));
      acc += __ldg(d_table + cell * 8 + k);
    }
    d_vol[i] += acc;
  }
}
}  // namespace hn

extern "C" {

int64_t hn_motion_field_workspace_bytes(int n_joints, int vol_size) {
  if (n_joints <= 0 || n_joints > kMaxJoints || vol_size < 2 || vol_size > 1024) return -1;
  const int64_t c = vol_size + 1;
  return (int64_t)n_joints * c * c * c * 32;
}

static const float4* build_table(int J, int V, const float* vol, void* workspace, int64_t workspace_bytes,
                                 cudaStream_t st, int* rc) {
  *rc = 0;
  if (!workspace) return nullptr;
  if (workspace_bytes < hn_motion_field_workspace_bytes(J, V)) {
    *rc = fail_arg("hn_motion_field: workspace smaller than hn_motion_field_workspace_bytes");
    return nullptr;
  }
  if (reinterpret_cast<uintptr_t>(workspace) & 15) {
    *rc = fail_arg("hn_motion_field: workspace must be 16-byte aligned");
    return nullptr;
  }
  const int64_t c = V + 1, cells = (int64_t)J * c * c * c;
  pack_corners_kernel<<<(unsigned)std::min<int64_t>(ceil_div<int64_t>(cells, 256), 148 * 16), 256, 0, st>>>(
      J, V, vol, static_cast<float4*>(workspace));
  *rc = launched("hn_motion_field/pack_corners");
  return static_cast<const float4*>(workspace);
}

int hn_motion_field_fwd(int64_t n_rays, int n_samples, int n_joints, int vol_size, const float* pts,
                        const float* rays_o, const float* rays_d, const float* z, const float* Rs,
                        const float* Ts, const float* vol, const float* bbox_min,
                        const float* bbox_scale, float* x_skel, float* mask, float* weights,
                        void* workspace, int64_t workspace_bytes, void* stream) {
  HN_TRY(check_common("hn_motion_field_fwd", n_rays, n_samples, n_joints, vol_size, pts, rays_o, rays_d, z,
                      Rs, Ts, vol, bbox_min, bbox_scale));
  if (!x_skel || !mask) return fail_arg("hn_motion_field_fwd: null output");
  int64_t n = n_rays * n_samples;
  if (n == 0) return 0;
  int rc = 0;
  const float4* table = build_table(n_joints, vol_size, vol, workspace, workspace_bytes, as_stream(stream), &rc);
  HN_TRY(rc);
  unsigned grid = (unsigned)ceil_div<int64_t>(n, 256);
  if (weights)
    motion_field_fwd_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(
        n, n_samples, n_joints, vol_size, pts, rays_o, rays_d, z, Rs, Ts, vol, bbox_min, bbox_scale, x_skel,
        mask, weights, table);
  else
    motion_field_fwd_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(
        n, n_samples, n_joints, vol_size, pts, rays_o, rays_d, z, Rs, Ts, vol, bbox_min, bbox_scale, x_skel,
        mask, nullptr, table);
  return launched("hn_motion_field_fwd");
}

int hn_motion_field_bwd(int64_t n_rays, int n_samples, int n_joints, int vol_size, const float* pts,
                        const float* rays_o, const float* rays_d, const float* z, const float* Rs,
                        const float* Ts, const float* vol, const float* bbox_min,
                        const float* bbox_scale, const float* d_x_skel, const float* d_mask,
                        const float* d_weights, const float* x_skel_fwd, const float* mask_fwd,
                        float* d_vol, float* d_Rs, float* d_Ts, const void* table, void* grad_workspace,
                        int64_t grad_workspace_bytes, void* stream) {
  HN_TRY(check_common("hn_motion_field_bwd", n_rays, n_samples, n_joints, vol_size, pts, rays_o, rays_d, z,
                      Rs, Ts, vol, bbox_min, bbox_scale));
  if (!d_vol || !d_Rs || !d_Ts) return fail_arg("hn_motion_field_bwd: null gradient output");
  int64_t n = n_rays * n_samples;
  if (n == 0) return 0;
  cudaStream_t st = as_stream(stream);
  // with a scratch buffer of hn_motion_field_workspace_bytes the volume gradient is scattered as 16-byte vector
  // reductions in the corner-table layout and folded into d_vol afterwards; without one: 8 scalar atomics per joint
  float4* d_table = nullptr;
  const int64_t tbytes = hn_motion_field_workspace_bytes(n_joints, vol_size);
  if (grad_workspace) {
    if (grad_workspace_bytes < tbytes) return fail_arg("hn_motion_field_bwd: gradient workspace too small");
    d_table = static_cast<float4*>(grad_workspace);
    HN_TRY(check_cuda(cudaMemsetAsync(d_table, 0, (size_t)tbytes, st), "hn_motion_field_bwd/memset"));
  }
  unsigned grid = (unsigned)ceil_div<int64_t>(n, 256);
  motion_field_bwd_kernel<<<grid, 256, 0, st>>>(
      n, n_samples, n_joints, vol_size, pts, rays_o, rays_d, z, Rs, Ts, vol, bbox_min, bbox_scale, d_x_skel,
      d_mask, d_weights, x_skel_fwd, mask_fwd, d_vol, d_Rs, d_Ts, static_cast<const float4*>(table), d_table);
  if (d_table) {
    HN_TRY(launched("hn_motion_field_bwd"));
    const int64_t voxels = (int64_t)n_joints * vol_size * vol_size * vol_size;
    unpack_corners_grad_kernel<<<(unsigned)std::min<int64_t>(ceil_div<int64_t>(voxels, 256), 4096), 256, 0, st>>>(
        n_joints, vol_size, reinterpret_cast<const float*>(d_table), d_vol);
    return launched("hn_motion_field_bwd/unpack");
  }
  return launched("hn_motion_field_bwd");
}

}  // extern "C"
