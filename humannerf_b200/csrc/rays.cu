// Ray set-up kernels: pixel rays (a1), ray/bbox intersection (a2), stream compaction,
// depth samples along the ray (a9), and the row scatter / gather used to place rays back
// into images and patches (a16).  HBM-bound byte / index work: one thread per element,
// coalesced, no tensor cores.
#include "common.cuh"

namespace hn {

// ---------------------------------------------------------------------------------------
// a1: reference core/utils/camera_util.py:132-159.  float64 throughout.
// cam = (x, y, 1) Kinv^T ; world = (cam - T) R ; o = -R^T T ; d = world - o
// ---------------------------------------------------------------------------------------
struct CameraF64 {
  double kinv[9];   // row-major K^-1
  double rot[9];    // row-major R (world -> camera)
  double trans[3];  // T
  double origin[3];
};

__global__ void gen_rays_kernel(int H, int W, CameraF64 cam, double* __restrict__ rays_o,
                                double* __restrict__ rays_d) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)H * W) return;
  double px = (double)(float)(i % W), py = (double)(float)(i / W);
  double c[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    // dot of (px, py, 1) with row a of K^-1, summed left to right without contraction
    double s = __dmul_rn(px, cam.kinv[a * 3 + 0]);
    s = __dadd_rn(s, __dmul_rn(py, cam.kinv[a * 3 + 1]));
    s = __dadd_rn(s, cam.kinv[a * 3 + 2]);
    c[a] = __dadd_rn(s, -cam.trans[a]);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    double s = __dmul_rn(c[0], cam.rot[0 * 3 + a]);
    s = __dadd_rn(s, __dmul_rn(c[1], cam.rot[1 * 3 + a]));
    s = __dadd_rn(s, __dmul_rn(c[2], cam.rot[2 * 3 + a]));
    rays_d[i * 3 + a] = __dadd_rn(s, -cam.origin[a]);
    rays_o[i * 3 + a] = cam.origin[a];
  }
}

static bool invert3x3(const double* m, double* out) {
  double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
  double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
  if (det == 0.0) return false;
  double r = 1.0 / det;
  out[0] = (e * i - f * h) * r; out[1] = (c * h - b * i) * r; out[2] = (b * f - c * e) * r;
  out[3] = (f * g - d * i) * r; out[4] = (a * i - c * g) * r; out[5] = (c * d - a * f) * r;
  out[6] = (d * h - e * g) * r; out[7] = (b * g - a * h) * r; out[8] = (a * e - b * d) * r;
  return true;
}

// ---------------------------------------------------------------------------------------
// a2: reference core/utils/camera_util.py:162-208.  Bit-exact float64 restatement:
// every product and sum is rounded separately (numpy evaluates them as separate ufuncs).
// ---------------------------------------------------------------------------------------
struct BoundsF64 {
  double lo[3], hi[3];          // padded by 0.01
  double lo_eps[3], hi_eps[3];  // +- 1e-6 for the inside test
};

__global__ void rays_bbox_kernel(int64_t n, BoundsF64 b, const double* __restrict__ rays_o,
                                 double* __restrict__ rays_d, double* __restrict__ near,
                                 double* __restrict__ far, uint8_t* __restrict__ mask) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double o[3], d[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    o[a] = rays_o[i * 3 + a];
    double v = rays_d[i * 3 + a];
    if (fabs(v) < 1e-5) {
      v = 1e-5;
      rays_d[i * 3 + a] = v;     // the reference patches the caller's array in place
    }
    d[a] = v;
  }
  int hits = 0;
  double p0[3] = {0, 0, 0}, p1[3] = {0, 0, 0};
  // plane order of the reference's reshape(-1, 6): min x,y,z then max x,y,z
#pragma unroll
  for (int plane = 0; plane < 6; ++plane) {
    int a = plane % 3;
    double bound = plane < 3 ? b.lo[a] : b.hi[a];
    double t = __ddiv_rn(__dadd_rn(bound, -o[a]), d[a]);
    double p[3];
    bool in = true;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      p[c] = __dadd_rn(__dmul_rn(t, d[c]), o[c]);
      in = in && (p[c] >= b.lo_eps[c]) && (p[c] <= b.hi_eps[c]);
    }
    if (in) {
      if (hits == 0) { p0[0] = p[0]; p0[1] = p[1]; p0[2] = p[2]; }
      else if (hits == 1) { p1[0] = p[0]; p1[1] = p[1]; p1[2] = p[2]; }
      ++hits;
    }
  }
  bool ok = hits == 2;
  mask[i] = ok ? 1 : 0;
  double nr = 0.0, fr = 0.0;
  if (ok) {
    auto norm3 = [](double x, double y, double z) {
      return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
    };
    double dn = norm3(d[0], d[1], d[2]);
    double d0 = __ddiv_rn(norm3(__dadd_rn(p0[0], -o[0]), __dadd_rn(p0[1], -o[1]), __dadd_rn(p0[2], -o[2])), dn);
    double d1 = __ddiv_rn(norm3(__dadd_rn(p1[0], -o[0]), __dadd_rn(p1[1], -o[1]), __dadd_rn(p1[2], -o[2])), dn);
    nr = fmin(d0, d1);
    fr = fmax(d0, d1);
  }
  near[i] = nr;
  far[i] = fr;
}

// ---------------------------------------------------------------------------------------
// Order-preserving stream compaction (three small kernels: block counts, scan, scatter).
// ---------------------------------------------------------------------------------------
constexpr int kCompactBlock = 1024;

__global__ void compact_count_kernel(int64_t n, const uint8_t* __restrict__ mask,
                                     int32_t* __restrict__ block_counts) {
  __shared__ int s_warp[32];
  int64_t i = (int64_t)blockIdx.x * kCompactBlock + threadIdx.x;
  int v = (i < n && mask[i]) ? 1 : 0;
  unsigned ballot = __ballot_sync(0xffffffffu, v);
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_warp[warp] = __popc(ballot);
  __syncthreads();
  if (warp == 0) {
    int c = s_warp[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) block_counts[blockIdx.x] = c;
  }
}

// single block: exclusive scan of block counts in place; total -> count
__global__ void compact_scan_kernel(int n_blocks, int32_t* __restrict__ block_counts,
                                    int64_t* __restrict__ count) {
  __shared__ int s_part[1024];
  int per = (n_blocks + blockDim.x - 1) / blockDim.x;
  int begin = threadIdx.x * per, end = min(begin + per, n_blocks);
  int sum = 0;
  for (int i = begin; i < end; ++i) sum += block_counts[i];
  s_part[threadIdx.x] = sum;
  __syncthreads();
  // Hillis-Steele inclusive scan over blockDim.x partials
  for (int o = 1; o < (int)blockDim.x; o <<= 1) {
    int v = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
    __syncthreads();
    s_part[threadIdx.x] += v;
    __syncthreads();
  }
  int run = s_part[threadIdx.x] - sum;
  for (int i = begin; i < end; ++i) {
    int c = block_counts[i];
    block_counts[i] = run;
    run += c;
  }
  if (threadIdx.x == blockDim.x - 1) *count = (int64_t)s_part[threadIdx.x];
}

__global__ void compact_scatter_kernel(int64_t n, const uint8_t* __restrict__ mask,
                                       const int32_t* __restrict__ block_offsets,
                                       const double* __restrict__ rays_o, const double* __restrict__ rays_d,
                                       const double* __restrict__ near, const double* __restrict__ far,
                                       float* __restrict__ out_o, float* __restrict__ out_d,
                                       float* __restrict__ out_near, float* __restrict__ out_far,
                                       int32_t* __restrict__ pixel_index) {
  __shared__ int s_warp[32];
  int64_t i = (int64_t)blockIdx.x * kCompactBlock + threadIdx.x;
  int v = (i < n && mask[i]) ? 1 : 0;
  unsigned ballot = __ballot_sync(0xffffffffu, v);
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_warp[warp] = __popc(ballot);
  __syncthreads();
  if (warp == 0) {
    int c = s_warp[lane], incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    s_warp[lane] = incl - c;
  }
  __syncthreads();
  if (!v) return;
  int64_t dst = (int64_t)block_offsets[blockIdx.x] + s_warp[warp] + __popc(ballot & ((1u << lane) - 1u));
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (out_o) out_o[dst * 3 + a] = (float)rays_o[i * 3 + a];
    if (out_d) out_d[dst * 3 + a] = (float)rays_d[i * 3 + a];
  }
  if (out_near) out_near[dst] = (float)near[i];
  if (out_far) out_far[dst] = (float)far[i];
  if (pixel_index) pixel_index[dst] = (int32_t)i;
}

// ---------------------------------------------------------------------------------------
// a9: reference network.py:455-471.  __fmul_rn/__fadd_rn keep the eager rounding sequence.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float z_plain(float nr, float fr, float t) {
  return __fadd_rn(__fmul_rn(nr, __fsub_rn(1.0f, t)), __fmul_rn(fr, t));
}

__global__ void sample_z_kernel(int64_t n_rays, int S, const float* __restrict__ near,
                                const float* __restrict__ far, const float* __restrict__ t_table,
                                const float* __restrict__ t_rand, float* __restrict__ z) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays * S) return;
  int64_t r = i / S;
  int s = (int)(i - r * S);
  float nr = near[r], fr = far[r];
  float zc = z_plain(nr, fr, t_table[s]);
  if (t_rand == nullptr) {
    z[i] = zc;
    return;
  }
  float lower = zc, upper = zc;
  if (s > 0) lower = __fmul_rn(0.5f, __fadd_rn(zc, z_plain(nr, fr, t_table[s - 1])));
  if (s < S - 1) upper = __fmul_rn(0.5f, __fadd_rn(z_plain(nr, fr, t_table[s + 1]), zc));
  z[i] = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t_rand[i]));
}

// ---------------------------------------------------------------------------------------
// a16: row scatter / gather
// ---------------------------------------------------------------------------------------
__global__ void scatter_rows_kernel(int64_t n_rows, int C, const float* __restrict__ src,
                                    const int64_t* __restrict__ idx, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * C) return;
  int64_t r = i / C;
  int c = (int)(i - r * C);
  if (idx[r] >= 0) out[idx[r] * C + c] = src[i];       // negative index = padding row, skipped
}

__global__ void gather_rows_kernel(int64_t n_rows, int C, const float* __restrict__ src,
                                   const int64_t* __restrict__ idx, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * C) return;
  int64_t r = i / C;
  int c = (int)(i - r * C);
  out[i] = idx[r] >= 0 ? src[idx[r] * C + c] : 0.0f;   // negative index = padding row
}

}  // namespace hn

using namespace hn;

extern "C" {

int hn_gen_rays(int H, int W, const double* K_host, const double* E_host, double* rays_o,
                double* rays_d, void* stream) {
  if (H <= 0 || W <= 0 || !K_host || !E_host || !rays_o || !rays_d)
    return fail_arg("hn_gen_rays: bad argument (H=%d W=%d)", H, W);
  CameraF64 cam;
  if (!invert3x3(K_host, cam.kinv)) return fail_arg("hn_gen_rays: singular intrinsics");
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) cam.rot[r * 3 + c] = E_host[r * 4 + c];
    cam.trans[r] = E_host[r * 4 + 3];
  }
  for (int a = 0; a < 3; ++a) {
    double s = 0.0;
    for (int b = 0; b < 3; ++b) s += cam.rot[b * 3 + a] * cam.trans[b];
    cam.origin[a] = -s;
  }
  int64_t n = (int64_t)H * W;
  gen_rays_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, as_stream(stream)>>>(H, W, cam, rays_o, rays_d);
  return launched("hn_gen_rays");
}

int hn_rays_bbox(int64_t n_rays, const double* bounds_host, const double* rays_o, double* rays_d,
                 double* near, double* far, uint8_t* mask, void* stream) {
  if (n_rays == 0) return 0;
  if (n_rays < 0 || !bounds_host || !rays_o || !rays_d || !near || !far || !mask)
    return fail_arg("hn_rays_bbox: bad argument");
  BoundsF64 b;
  for (int a = 0; a < 3; ++a) {
    b.lo[a] = bounds_host[a] + (-0.01);
    b.hi[a] = bounds_host[3 + a] + 0.01;
    b.lo_eps[a] = b.lo[a] - 1e-6;
    b.hi_eps[a] = b.hi[a] + 1e-6;
  }
  rays_bbox_kernel<<<(unsigned)ceil_div<int64_t>(n_rays, 256), 256, 0, as_stream(stream)>>>(
      n_rays, b, rays_o, rays_d, near, far, mask);
  return launched("hn_rays_bbox");
}

int64_t hn_compact_rays_workspace_bytes(int64_t n_rays) {
  return (ceil_div<int64_t>(n_rays > 0 ? n_rays : 1, kCompactBlock) + 1) * (int64_t)sizeof(int32_t);
}

int hn_compact_rays(int64_t n_rays, const uint8_t* mask, const double* rays_o, const double* rays_d,
                    const double* near, const double* far, float* out_o, float* out_d,
                    float* out_near, float* out_far, int32_t* pixel_index, int64_t* count,
                    void* workspace, void* stream) {
  if (n_rays < 0 || !mask || !count || !workspace) return fail_arg("hn_compact_rays: bad argument");
  if ((out_o && !rays_o) || (out_d && !rays_d) || (out_near && !near) || (out_far && !far))
    return fail_arg("hn_compact_rays: output requested without its source");
  cudaStream_t st = as_stream(stream);
  if (n_rays == 0) return check_cuda(cudaMemsetAsync(count, 0, sizeof(int64_t), st), "hn_compact_rays");
  int n_blocks = (int)ceil_div<int64_t>(n_rays, kCompactBlock);
  int32_t* counts = static_cast<int32_t*>(workspace);
  compact_count_kernel<<<n_blocks, kCompactBlock, 0, st>>>(n_rays, mask, counts);
  HN_TRY(launched("hn_compact_rays/count"));
  compact_scan_kernel<<<1, 1024, 0, st>>>(n_blocks, counts, count);
  HN_TRY(launched("hn_compact_rays/scan"));
  compact_scatter_kernel<<<n_blocks, kCompactBlock, 0, st>>>(n_rays, mask, counts, rays_o, rays_d, near,
                                                              far, out_o, out_d, out_near, out_far,
                                                              pixel_index);
  return launched("hn_compact_rays/scatter");
}

int hn_sample_z(int64_t n_rays, int n_samples, const float* near, const float* far,
                const float* t_table, const float* t_rand, float* z, void* stream) {
  if (n_rays == 0 && n_samples > 0) return 0;
  if (n_rays < 0 || n_samples <= 0 || !near || !far || !t_table || !z)
    return fail_arg("hn_sample_z: bad argument");
  int64_t n = n_rays * n_samples;
  sample_z_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, as_stream(stream)>>>(
      n_rays, n_samples, near, far, t_table, t_rand, z);
  return launched("hn_sample_z");
}

int hn_scatter_rows(int64_t n_rows, int n_cols, const float* src, const int64_t* dst_index,
                    float* out, void* stream) {
  if (n_rows == 0 && n_cols > 0) return 0;
  if (n_rows < 0 || n_cols <= 0 || !src || !dst_index || !out) return fail_arg("hn_scatter_rows: bad argument");
  int64_t n = n_rows * n_cols;
  scatter_rows_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, as_stream(stream)>>>(n_rows, n_cols, src,
                                                                                         dst_index, out);
  return launched("hn_scatter_rows");
}

int hn_gather_rows(int64_t n_rows, int n_cols, const float* src, const int64_t* src_index,
                   float* out, void* stream) {
  if (n_rows == 0 && n_cols > 0) return 0;
  if (n_rows < 0 || n_cols <= 0 || !src || !src_index || !out) return fail_arg("hn_gather_rows: bad argument");
  int64_t n = n_rows * n_cols;
  gather_rows_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, as_stream(stream)>>>(n_rows, n_cols, src,
                                                                                        src_index, out);
  return launched("hn_gather_rows");
}

}  // extern "C"
