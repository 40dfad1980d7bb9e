// Canonical field, tensor-core backward (HN_PRECISION_BF16): the C-ABI entry point.
//
//   1. field_tc2_bwd_kernel (mlp_tc_bwd_2sm.cu): the backward-DATA chain of one MLP on CTA pairs.  dZ of the
//      current layer sits in shared memory as the A operand, W^T chunks stream through the TMA ring, dH = dZ W
//      lands in TMEM, and the epilogue masks it with ReLU' (1-bit mask images written by the forward) to get dZ of
//      the previous layer - written in place as the next A operand AND to a T16 image in HBM for the
//      weight-gradient kernel.  The gradient of the positional-encoding columns is contracted with the
//      encoding's derivative in the epilogue (d_pos += PE'(pos) . d_pe), so the (P, pe_pad) encoding gradients
//      never reach HBM,
//   2. hn_tile_wgrad jobs (tile_wgrad.cu) produce every dW and db from the dZ / activation images,
//   3. the condition-code columns of the first non-rigid layer get dW0[:, :69] += db0 (x) cond.
//
// Autograd of reference mlp_offset.py:74-114 and mlp_rgb_sigma.py:132-206 on the bf16 path.
#include "mlp_tc_plan.cuh"

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
  return hn_field_bwd_bf16_phases(p, g, n_samples, x_skel, xyz, cond, hann_w, ignore_non_rigid, d_raw, d_xyz, d_offsets, d_x_skel,
                                  workspace, workspace_bytes, packed, scratch, scratch_bytes, n_samples_dev,
                                  HN_BWD_DATA | HN_BWD_WEIGHTS, stream);
}

int hn_field_bwd_bf16_phases(const hn_field_params* p, const hn_field_grads* g, int64_t n_samples, const float* x_skel,
                             const float* xyz, const float* cond, const float* hann_w, int ignore_non_rigid,
                             const float* d_raw, const float* d_xyz, const float* d_offsets, float* d_x_skel,
                             void* workspace, int64_t workspace_bytes, const void* packed, void* scratch,
                             int64_t scratch_bytes, const int64_t* n_samples_dev, int phases, void* stream) {
  TcPlan plan;
  HN_TRY(make_tc_plan(p, plan));
  if (n_samples == 0) return 0;
  if ((phases & ~(HN_BWD_DATA | HN_BWD_WEIGHTS)) || !phases) return fail_arg("hn_field_bwd_bf16_phases: phases = %d", phases);
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
  const uint8_t* pk = static_cast<const uint8_t*>(packed);

  if (phases & HN_BWD_DATA) {
  // 1. canonical MLP: d_raw -> dZ images; d_xyz_tot = upstream d_xyz + PE'(xyz) . d(encoding), accumulated
  //    by the chain's epilogue (both encoding reads: skip layer and layer 0)
  if (d_xyz) HN_TRY(copy_cols(n, 3, d_xyz, 3, d_xyz_tot, 4, st));
  else HN_TRY(check_cuda(cudaMemsetAsync(d_xyz_tot, 0, (size_t)rows * 4 * sizeof(float), st), "hn_field_bwd_bf16/memset"));
  if (sms < 2) return fail_arg("hn_field_bwd_bf16: the tcgen05 chain kernel needs a device with thread-block clusters");
  if (!use_relu_bits()) memset(bufs.bits, 0, sizeof(bufs.bits));     // HN_TC_RELU_BITS=0: ReLU' from the activation images
  HN_TRY(launch_field_tc2_bwd(plan, pk, n, n_samples_dev, 1, d_raw, 4, 4, bufs, xyz, nullptr, d_xyz_tot, 4, sms, st));
  // 2. non-rigid MLP: its output gradient is d_xyz_tot + d_offsets; d_x_skel = d_xyz_tot + PE'(x_skel) . d(encoding)
  HN_TRY(copy_cols(n, 3, d_xyz_tot, 4, d_x_skel, 3, st));
  if (!ignore_non_rigid) {
    HN_TRY(add3(n, d_xyz_tot, 4, d_offsets, 3, dz_nr, 4, st));
    HN_TRY(launch_field_tc2_bwd(plan, pk, n, n_samples_dev, 0, dz_nr, 4, 3, bufs, x_skel, hann_w, d_x_skel, 3, sms, st));
  }
  }
  if (!(phases & HN_BWD_WEIGHTS)) return 0;
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
