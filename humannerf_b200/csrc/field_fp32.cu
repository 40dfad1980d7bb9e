// Canonical field, fp32 path (HN_PRECISION_FP32): hann-PE -> non-rigid MLP -> xyz -> PE ->
// canonical MLP (reference embedders/*.py, mlp_offset.py:74-114, mlp_rgb_sigma.py:132-206,
// network.py:233-327), forward and backward, sequenced on one stream from host code.
//
// Workspace (floats per sample, all row strides multiples of 4 so rows stay 16-byte aligned):
//   non-rigid : in0 [cond | pe] (ld = pad4(cond+pe)), h_1 .. h_{L-2} (ld = width; the input of the
//               skip layer is [h | pe], ld = pad4(width + pe)), offsets (ld 4)
//   canonical : pe (ld = pad4(3 + 6 F)), h_1 .. h_{L-2} (skip input is [pe | h],
//               ld = pad4(pe + width))
// With save_for_backward every buffer is kept; otherwise the sample range is processed in
// sub-chunks that reuse a small ring.
#include <algorithm>
#include <vector>

#include "sgemm.cuh"

namespace hn {

static inline int pad4(int v) { return (v + 3) & ~3; }

int64_t tc_train_workspace_bytes(const hn_field_params* p, int64_t n_samples);   // mlp_tc.cu

// ---------------------------------------------------------------------------------------
// positional encodings
// out[:, col0 + ...] = [x (if include_input), w_k sin(2^k x), w_k cos(2^k x)]_k ; per k the
// three sines come first, then the three cosines (embedders/fourier.py:22-25).
// ---------------------------------------------------------------------------------------
__global__ void pe_fwd_kernel(int64_t n, const float* __restrict__ x, int64_t ldx, int n_freqs,
                              const float* __restrict__ fw, int include_input, float* __restrict__ out,
                              int64_t ldo) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v[3] = {x[i * ldx + 0], x[i * ldx + 1], x[i * ldx + 2]};
  float* o = out + i * ldo;
  int c = 0;
  if (include_input) { o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; c = 3; }
  float freq = 1.0f;
  for (int k = 0; k < n_freqs; ++k) {
    float w = fw ? __ldg(fw + k) : 1.0f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float s, co;
      sincosf(v[a] * freq, &s, &co);
      o[c + a] = fw ? w * s : s;
      o[c + 3 + a] = fw ? w * co : co;
    }
    c += 6;
    freq *= 2.0f;
  }
}

// d_x[i, a] (+)= [include_input] d_out[a] + sum_k w_k 2^k (cos(2^k x) d_sin - sin(2^k x) d_cos)
__global__ void pe_bwd_kernel(int64_t n, const float* __restrict__ x, int64_t ldx, int n_freqs,
                              const float* __restrict__ fw, int include_input,
                              const float* __restrict__ d_out, int64_t ldo, float* __restrict__ d_x,
                              int64_t lddx, int accumulate) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* g = d_out + i * ldo;
  float v[3] = {x[i * ldx + 0], x[i * ldx + 1], x[i * ldx + 2]};
  float acc[3] = {0.f, 0.f, 0.f};
  int c = 0;
  if (include_input) { acc[0] = g[0]; acc[1] = g[1]; acc[2] = g[2]; c = 3; }
  float freq = 1.0f;
  for (int k = 0; k < n_freqs; ++k) {
    float w = (fw ? __ldg(fw + k) : 1.0f) * freq;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float s, co;
      sincosf(v[a] * freq, &s, &co);
      acc[a] += w * (co * g[c + a] - s * g[c + 3 + a]);
    }
    c += 6;
    freq *= 2.0f;
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float r = acc[a];
    if (accumulate) r += d_x[i * lddx + a];
    d_x[i * lddx + a] = r;
  }
}

// dst[i, c0:c0+n_cols] = src[i, 0:n_cols]   (or a broadcast row when src_ld == 0)
__global__ void copy_cols_kernel(int64_t n, int n_cols, const float* __restrict__ src, int64_t src_ld,
                                 float* __restrict__ dst, int64_t dst_ld) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * n_cols) return;
  int64_t r = i / n_cols;
  int c = (int)(i - r * n_cols);
  dst[r * dst_ld + c] = src[r * src_ld + c];
}

// out[i, 0:3] = a[i, 0:3] + b[i, 0:3]   (b nullable)
__global__ void add3_kernel(int64_t n, const float* __restrict__ a, int64_t lda, const float* __restrict__ b,
                            int64_t ldb, float* __restrict__ out, int64_t ldo) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 3) return;
  int64_t r = i / 3;
  int c = (int)(i - r * 3);
  float v = a[r * lda + c];
  if (b) v += b[r * ldb + c];
  out[r * ldo + c] = v;
}

// db[n] += sum_m dY[m, n]
__global__ void colsum_kernel(int64_t M, int N, const float* __restrict__ dY, int64_t ld,
                              float* __restrict__ db, int rows_per_block) {
  __shared__ float s_part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(M, r0 + rows_per_block);
  for (int n0 = 0; n0 < N; n0 += 32) {
    int n = n0 + tx;
    float s = 0.0f;
    if (n < N)
      for (int64_t r = r0 + ty; r < r1; r += 8) s += dY[r * ld + n];
    s_part[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && n < N) {
      float t = 0.0f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += s_part[k][tx];
      atomicAdd(db + n, t);
    }
    __syncthreads();
  }
}

template <typename F>
static int launch_1d(int64_t n, cudaStream_t st, const char* what, F&& f) {
  if (n <= 0) return 0;
  f((unsigned)ceil_div<int64_t>(n, 256), 256, st);
  return launched(what);
}

static int pe_fwd(int64_t n, const float* x, int64_t ldx, int nf, const float* fw, bool incl, float* out,
                  int64_t ldo, cudaStream_t st) {
  return launch_1d(n, st, "pe_fwd", [&](unsigned g, unsigned b, cudaStream_t s) {
    pe_fwd_kernel<<<g, b, 0, s>>>(n, x, ldx, nf, fw, incl ? 1 : 0, out, ldo);
  });
}

int pe_bwd(int64_t n, const float* x, int64_t ldx, int nf, const float* fw, bool incl,
                  const float* d_out, int64_t ldo, float* d_x, int64_t lddx, bool accumulate,
                  cudaStream_t st) {
  return launch_1d(n, st, "pe_bwd", [&](unsigned g, unsigned b, cudaStream_t s) {
    pe_bwd_kernel<<<g, b, 0, s>>>(n, x, ldx, nf, fw, incl ? 1 : 0, d_out, ldo, d_x, lddx, accumulate ? 1 : 0);
  });
}

int copy_cols(int64_t n, int n_cols, const float* src, int64_t src_ld, float* dst, int64_t dst_ld,
                     cudaStream_t st) {
  return launch_1d(n * n_cols, st, "copy_cols", [&](unsigned g, unsigned b, cudaStream_t s) {
    copy_cols_kernel<<<g, b, 0, s>>>(n, n_cols, src, src_ld, dst, dst_ld);
  });
}

int add3(int64_t n, const float* a, int64_t lda, const float* b, int64_t ldb, float* out, int64_t ldo,
                cudaStream_t st) {
  return launch_1d(n * 3, st, "add3", [&](unsigned g, unsigned bl, cudaStream_t s) {
    add3_kernel<<<g, bl, 0, s>>>(n, a, lda, b, ldb, out, ldo);
  });
}

static int colsum(int64_t M, int N, const float* dY, int64_t ld, float* db, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  const int rows = 2048;
  colsum_kernel<<<(unsigned)ceil_div<int64_t>(M, rows), 256, 0, st>>>(M, N, dY, ld, db, rows);
  return launched("colsum");
}

// ---------------------------------------------------------------------------------------
// Layer plan shared by forward and backward
// ---------------------------------------------------------------------------------------
struct Buf {
  int64_t off;   // float offset of the buffer inside the workspace (per-chunk base added later)
  int ld;
};

struct MlpPlan {
  int layers = 0, width = 0, skip = -1, pe_dim = 0, pe_ld = 0;
  bool pe_first = false;                // canonical: skip input is [pe | h]; non-rigid: [h | pe]
  int in_dim[HN_MAX_LAYERS];            // K of layer l
  int out_dim[HN_MAX_LAYERS];           // N of layer l
  Buf in[HN_MAX_LAYERS];                // input buffer of layer l
  int h_col[HN_MAX_LAYERS];             // column where the previous layer's output sits in in[l]
  int pe_col[HN_MAX_LAYERS];            // column of the (re-)appended encoding in in[l], -1 if none
  Buf out_last;                         // output of the last layer (ld 4)
  int64_t floats_per_sample = 0;
};

struct FieldPlan {
  MlpPlan nr, cn;
  int64_t floats_per_sample = 0;
};

static int make_plan(const hn_field_params* p, FieldPlan& fp) {
  if (!p) return fail_arg("hn_field: null params");
  if (p->nr_layers < 2 || p->nr_layers > HN_MAX_LAYERS || p->cn_layers < 2 || p->cn_layers > HN_MAX_LAYERS)
    return fail_arg("hn_field: layer counts (%d, %d) outside [2,%d]", p->nr_layers, p->cn_layers, HN_MAX_LAYERS);
  if (p->nr_width <= 0 || p->cn_width <= 0 || p->nr_width % 4 || p->cn_width % 4)
    return fail_arg("hn_field: widths must be positive multiples of 4");
  if (p->nr_freqs < 0 || p->nr_freqs > 16 || p->cn_freqs < 0 || p->cn_freqs > 16 || p->cond_dim < 0)
    return fail_arg("hn_field: bad encoding sizes");
  if (p->nr_skip == 0 || p->nr_skip >= p->nr_layers - 1 || p->cn_skip == 0 || p->cn_skip >= p->cn_layers - 1)
    return fail_arg("hn_field: skip layer must be a hidden layer (or -1)");
  int64_t cur = 0;
  auto build = [&](MlpPlan& m, int layers, int width, int skip, int pe_dim, int lead_dim, bool pe_first,
                   int out_last_dim) {
    m.layers = layers; m.width = width; m.skip = skip; m.pe_dim = pe_dim; m.pe_first = pe_first;
    for (int l = 0; l < layers; ++l) {
      int k;
      if (l == 0) {
        k = lead_dim + pe_dim;
        m.h_col[l] = -1;
        m.pe_col[l] = lead_dim;
      } else if (l == skip) {
        k = width + pe_dim;
        m.h_col[l] = pe_first ? pe_dim : 0;
        m.pe_col[l] = pe_first ? 0 : width;
      } else {
        k = width;
        m.h_col[l] = 0;
        m.pe_col[l] = -1;
      }
      m.in_dim[l] = k;
      m.out_dim[l] = (l == layers - 1) ? out_last_dim : width;
      m.in[l].ld = pad4(k);
      m.in[l].off = cur;
      cur += m.in[l].ld;
    }
    m.pe_ld = m.in[0].ld;
    m.out_last.ld = 4;
    m.out_last.off = cur;
    cur += 4;
  };
  build(fp.nr, p->nr_layers, p->nr_width, p->nr_skip, 6 * p->nr_freqs, p->cond_dim, false, 3);
  fp.nr.floats_per_sample = cur;
  int64_t before = cur;
  build(fp.cn, p->cn_layers, p->cn_width, p->cn_skip, 3 + 6 * p->cn_freqs, 0, true, 4);
  fp.cn.floats_per_sample = cur - before;
  fp.floats_per_sample = cur;
  return 0;
}

constexpr int64_t kEvalChunk = 262144;   // samples per sub-chunk when nothing is saved

static int64_t chunk_samples(int64_t n, bool save) { return save ? std::max<int64_t>(n, 1) : std::min<int64_t>(std::max<int64_t>(n, 1), kEvalChunk); }

// forward of one MLP over samples [0, n) with workspace base `ws` (row-major buffers of n rows each)
static int mlp_forward(const MlpPlan& m, const float* const* W, const float* const* B, int64_t n, float* ws,
                       int64_t rows, cudaStream_t st) {
  for (int l = 0; l < m.layers; ++l) {
    const bool last = l == m.layers - 1;
    const float* X = ws + m.in[l].off * rows;
    float* Y;
    int64_t ldy;
    if (last) {
      Y = ws + m.out_last.off * rows;
      ldy = m.out_last.ld;
    } else {
      Y = ws + m.in[l + 1].off * rows + m.h_col[l + 1];
      ldy = m.in[l + 1].ld;
    }
    HN_TRY(linear_fwd(n, m.out_dim[l], m.in_dim[l], X, m.in[l].ld, W[l], m.in_dim[l], B[l], !last, Y, ldy, st));
  }
  return 0;
}

}  // namespace hn

using namespace hn;

extern "C" {

int64_t hn_field_workspace_bytes(const hn_field_params* p, int64_t n_samples, int save_for_backward,
                                 int precision) {
  FieldPlan fp;
  if (make_plan(p, fp) != 0) return -1;
  if (precision == HN_PRECISION_BF16)                // forward-only: activations stay on chip
    return save_for_backward ? tc_train_workspace_bytes(p, n_samples) : 256;
  int64_t rows = chunk_samples(n_samples, save_for_backward != 0);
  return fp.floats_per_sample * rows * (int64_t)sizeof(float) + 256;
}

int64_t hn_field_bwd_scratch_bytes(const hn_field_params* p, int64_t n_samples) {
  FieldPlan fp;
  if (make_plan(p, fp) != 0) return -1;
  int wide = std::max(fp.nr.in[0].ld, fp.cn.in[0].ld);
  for (int l = 0; l < fp.nr.layers; ++l) wide = std::max(wide, fp.nr.in[l].ld);
  for (int l = 0; l < fp.cn.layers; ++l) wide = std::max(wide, fp.cn.in[l].ld);
  // two ping-pong gradient buffers + d_offsets_total(4) + d_pe + d_xyz_total(4)
  int64_t per = 2 * (int64_t)wide + 4 + std::max(fp.nr.pe_ld, fp.cn.pe_ld) + 4;
  return per * std::max<int64_t>(n_samples, 1) * (int64_t)sizeof(float) + 256;
}

int hn_field_fwd_fp32(const hn_field_params* p, int64_t n_samples, const float* x_skel, const float* cond,
                      const float* hann_w, int ignore_non_rigid, float* raw, float* xyz, float* offsets,
                      void* workspace, int64_t workspace_bytes, int save_for_backward, void* stream) {
  FieldPlan fp;
  HN_TRY(make_plan(p, fp));
  if (n_samples == 0) return 0;        // an empty chunk (all samples culled) has empty tensors: nothing to check or do
  if (n_samples < 0 || !x_skel || !raw || !xyz || !offsets || !workspace)
    return fail_arg("hn_field_fwd: null pointer or negative size");
  if (!ignore_non_rigid && ((p->cond_dim > 0 && !cond) || (p->nr_freqs > 0 && !hann_w)))
    return fail_arg("hn_field_fwd: cond / hann_w missing");
  const bool save = save_for_backward != 0;
  const int64_t rows = chunk_samples(n_samples, save);
  if (workspace_bytes < fp.floats_per_sample * rows * (int64_t)sizeof(float))
    return fail_arg("hn_field_fwd: workspace too small (%lld bytes)", (long long)workspace_bytes);
  cudaStream_t st = as_stream(stream);
  float* ws = static_cast<float*>(workspace);
  for (int64_t a = 0; a < n_samples; a += rows) {
    const int64_t n = std::min(rows, n_samples - a);
    const float* xs = x_skel + a * 3;
    float* xyz_c = xyz + a * 3;
    float* off_c = offsets + a * 3;
    if (!ignore_non_rigid) {
      const MlpPlan& m = fp.nr;
      float* in0 = ws + m.in[0].off * rows;
      if (p->cond_dim > 0) HN_TRY(copy_cols(n, p->cond_dim, cond, 0, in0, m.in[0].ld, st));
      HN_TRY(pe_fwd(n, xs, 3, p->nr_freqs, hann_w, false, in0 + p->cond_dim, m.in[0].ld, st));
      if (m.skip > 0)
        HN_TRY(copy_cols(n, m.pe_dim, in0 + p->cond_dim, m.in[0].ld,
                         ws + m.in[m.skip].off * rows + m.pe_col[m.skip], m.in[m.skip].ld, st));
      HN_TRY(mlp_forward(m, p->nr_w, p->nr_b, n, ws, rows, st));
      const float* off = ws + m.out_last.off * rows;
      HN_TRY(copy_cols(n, 3, off, 4, off_c, 3, st));
      HN_TRY(add3(n, xs, 3, off, 4, xyz_c, 3, st));
    } else {
      HN_TRY(check_cuda(cudaMemsetAsync(off_c, 0, (size_t)n * 3 * sizeof(float), st), "hn_field_fwd/memset"));
      HN_TRY(add3(n, xs, 3, nullptr, 0, xyz_c, 3, st));
    }
    {
      const MlpPlan& m = fp.cn;
      float* pe = ws + m.in[0].off * rows;
      HN_TRY(pe_fwd(n, xyz_c, 3, p->cn_freqs, nullptr, true, pe, m.in[0].ld, st));
      if (m.skip > 0)
        HN_TRY(copy_cols(n, m.pe_dim, pe, m.in[0].ld, ws + m.in[m.skip].off * rows + m.pe_col[m.skip],
                         m.in[m.skip].ld, st));
      HN_TRY(mlp_forward(m, p->cn_w, p->cn_b, n, ws, rows, st));
      HN_TRY(copy_cols(n, 4, ws + m.out_last.off * rows, 4, raw + a * 4, 4, st));
    }
  }
  return 0;
}

// backward of one MLP.  dz: gradient w.r.t. the last layer's output (n x out_last_dim, ld 4).
// Leaves d(encoding) accumulated in d_pe (n x pe_ld, pre-zeroed by the caller through the first
// non-accumulating write) and parameter gradients in gW / gB.
static int mlp_backward(const MlpPlan& m, const float* const* W, float* const* gW, float* const* gB, int64_t n,
                        const float* ws, int64_t rows, const float* dz_last, int64_t ld_dz_last, float* bufA,
                        float* bufB, int64_t ld_buf, float* d_pe, cudaStream_t st) {
  const float* dz = dz_last;
  int64_t ld_dz = ld_dz_last;
  bool pe_written = false;
  float* next = bufA;
  for (int l = m.layers - 1; l >= 0; --l) {
    const float* X = ws + m.in[l].off * rows;
    const int N = m.out_dim[l], K = m.in_dim[l];
    if (gW[l]) HN_TRY(linear_bwd_weight(n, N, K, dz, ld_dz, X, m.in[l].ld, gW[l], K, st));
    if (gB[l]) HN_TRY(colsum(n, N, dz, ld_dz, gB[l], st));
    // gradient of the encoding columns of this layer's input
    if (m.pe_col[l] >= 0 && m.pe_dim > 0) {
      HN_TRY(linear_bwd_data(n, N, m.pe_dim, dz, ld_dz, W[l] + m.pe_col[l], K, d_pe, m.pe_ld, nullptr, 0,
                             pe_written, st));
      pe_written = true;
    }
    if (l > 0) {
      // gradient w.r.t. the previous layer's pre-activation: (dz W_l[:, h cols]) * [h > 0]
      HN_TRY(linear_bwd_data(n, N, m.width, dz, ld_dz, W[l] + m.h_col[l], K, next, ld_buf,
                             X + m.h_col[l], m.in[l].ld, false, st));
      dz = next;
      ld_dz = ld_buf;
      next = (next == bufA) ? bufB : bufA;
    }
  }
  return 0;
}

int hn_field_bwd(const hn_field_params* p, const hn_field_grads* g, int64_t n_samples, const float* x_skel,
                 const float* cond, const float* hann_w, int ignore_non_rigid, const float* d_raw,
                 const float* d_xyz, const float* d_offsets, float* d_x_skel, const void* workspace,
                 int64_t workspace_bytes, void* scratch, int64_t scratch_bytes, void* stream) {
  FieldPlan fp;
  HN_TRY(make_plan(p, fp));
  if (n_samples == 0) return 0;
  if (!g || n_samples < 0 || !x_skel || !d_raw || !d_x_skel || !workspace || !scratch)
    return fail_arg("hn_field_bwd: null pointer or negative size");
  const int64_t n = n_samples, rows = n_samples;
  if (workspace_bytes < fp.floats_per_sample * rows * (int64_t)sizeof(float))
    return fail_arg("hn_field_bwd: workspace is not a save_for_backward workspace of this size");
  if (scratch_bytes < hn_field_bwd_scratch_bytes(p, n_samples) - 256)
    return fail_arg("hn_field_bwd: scratch too small");
  cudaStream_t st = as_stream(stream);
  const float* ws = static_cast<const float*>(workspace);
  int wide = 0;
  for (int l = 0; l < fp.nr.layers; ++l) wide = std::max(wide, fp.nr.in[l].ld);
  for (int l = 0; l < fp.cn.layers; ++l) wide = std::max(wide, fp.cn.in[l].ld);
  float* bufA = static_cast<float*>(scratch);
  float* bufB = bufA + (int64_t)wide * n;
  float* dz_off = bufB + (int64_t)wide * n;                                     // (n, 4)
  float* d_pe = dz_off + 4 * n;
  float* d_xyz_tot = d_pe + (int64_t)std::max(fp.nr.pe_ld, fp.cn.pe_ld) * n;   // (n, 4)

  // canonical MLP: d_raw -> d(pe(xyz)) -> d_xyz
  HN_TRY(mlp_backward(fp.cn, p->cn_w, g->cn_w, g->cn_b, n, ws, rows, d_raw, 4, bufA, bufB, wide, d_pe, st));
  // the encoding input xyz = x_skel + offsets is rebuilt (ld 4) in bufA, which is free again here
  float* xyz_re = bufA;
  if (!ignore_non_rigid) HN_TRY(add3(n, x_skel, 3, ws + fp.nr.out_last.off * rows, 4, xyz_re, 4, st));
  else HN_TRY(add3(n, x_skel, 3, nullptr, 0, xyz_re, 4, st));
  if (d_xyz) HN_TRY(copy_cols(n, 3, d_xyz, 3, d_xyz_tot, 4, st));
  HN_TRY(pe_bwd(n, xyz_re, 4, p->cn_freqs, nullptr, true, d_pe, fp.cn.pe_ld, d_xyz_tot, 4, d_xyz != nullptr, st));
  if (ignore_non_rigid) {
    return copy_cols(n, 3, d_xyz_tot, 4, d_x_skel, 3, st);
  }
  // non-rigid MLP: d_offsets_total = d_xyz_tot + d_offsets
  HN_TRY(add3(n, d_xyz_tot, 4, d_offsets, 3, dz_off, 4, st));
  HN_TRY(mlp_backward(fp.nr, p->nr_w, g->nr_w, g->nr_b, n, ws, rows, dz_off, 4, bufA, bufB, wide, d_pe, st));
  // d_x_skel = d_xyz_tot + PE'(x_skel) . d_pe
  HN_TRY(copy_cols(n, 3, d_xyz_tot, 4, d_x_skel, 3, st));
  return pe_bwd(n, x_skel, 3, p->nr_freqs, hann_w, false, d_pe, fp.nr.pe_ld, d_x_skel, 3, true, st);
}

}  // extern "C"
