// Canonical field, tensor-core path (HN_PRECISION_BF16): weight packing and the C-ABI entry points of the forward.
//
//   hn_field_pack_bf16 / hn_field_repack_bf16   fp32 nn.Linear weights -> the bf16 UMMA shared-memory images the
//                                               tcgen05 kernels stream (layout: DESIGN.md section 3)
//   hn_field_fwd                                fp32 SIMT path (field_fp32.cu) or the cta_group::2 kernel (mlp_tc_2sm.cu)
//
// Reference semantics: embedders/hannw_fourier.py:38-49, embedders/fourier.py:9-38,
// non_rigid_motion_mlps/mlp_offset.py:74-114, canonical_mlps/mlp_rgb_sigma.py:132-206.
#include <cstdlib>

#include "mlp_tc_plan.cuh"

namespace hn {


// ------------------------------------------------------------------------------------------
// Weight packing
// ------------------------------------------------------------------------------------------
// one thread per 16-byte core-matrix row: image[(k/8) * (N*16) + n*16 .. +16] = W[n][k0+k .. +8]
__global__ void pack_weights_kernel(const PackJob* __restrict__ jobs, int n_jobs, const float* __restrict__ cond,
                                    uint8_t* __restrict__ image) {
  const int j = blockIdx.y;
  if (j >= n_jobs) return;
  const PackJob job = jobs[j];
  const int units = job.n * (job.kc / 8);
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < units; u += gridDim.x * blockDim.x) {
    const int kg = u / job.n, n = u - kg * job.n;
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = kg * 8 + i;
      float v = 0.0f;
      if (n < job.valid_n) {
        if (job.src && k < job.valid_k)
          v = job.transpose ? job.src[(int64_t)(job.src_row0 + k) * job.src_ld + job.src_col0 + n]
                            : job.src[(int64_t)(job.src_n0 + n) * job.src_ld + job.src_col0 + k];
        if (k == job.bias_row) {
          v = job.bias ? job.bias[job.src_n0 + n] : 0.0f;
          if (job.fold_w && cond)
            for (int c = 0; c < job.fold_k; ++c) v += job.fold_w[(int64_t)(job.src_n0 + n) * job.fold_ld + c] * cond[c];
        }
      }
      f[i] = v;
    }
    uint4 o;
    o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]); o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
    size_t dst = (size_t)kg * job.n * 16 + (size_t)n * 16;
    if (job.split) {
      // two half images (rows [0, n/2) then [n/2, n)); a TS job fills K rows [dst_k0, dst_k0 + kc) of a dst_kc-row chunk
      const int nh = job.n >> 1, r = n >= nh ? 1 : 0;
      const int kc_img = job.dst_kc > 0 ? job.dst_kc : job.kc, kg_img = kg + job.dst_k0 / 8;
      dst = (size_t)r * (kc_img / 8) * nh * 16 + (size_t)kg_img * nh * 16 + (size_t)(n - r * nh) * 16;
    }
    *reinterpret_cast<uint4*>(image + job.w_off + dst) = o;
  }
}

// mlp_tc_2sm.cu
int launch_field_tc2(const TcPlan& plan, const uint8_t* packed, int64_t n_samples, const int64_t* n_samples_dev, const float* x_skel,
                     const float* hann_w, int ignore_non_rigid, float* raw, float* xyz, float* offsets,
                     const TrainBufs* bufs, int sms, cudaStream_t st);

// mlp_tc_ts.cu
int launch_field_ts(const TcPlan& plan, const uint8_t* packed, int64_t n_samples, const int64_t* n_samples_dev, const float* x_skel,
                    const float* hann_w, int ignore_non_rigid, float* raw, float* xyz, float* offsets,
                    const TrainBufs* bufs, int sms, cudaStream_t st);

// mlp_tc_ts2.cu
int launch_field_ts2(const TcPlan& plan, const uint8_t* packed, int64_t n_samples, const int64_t* n_samples_dev, const float* x_skel,
                     const float* hann_w, int ignore_non_rigid, float* raw, float* xyz, float* offsets, int sms, cudaStream_t st);

// Forward kernel selection, read once per process.  Default ("auto"): rendering (nothing saved for backward) runs the TS
// kernel (hidden activations in tensor memory, mlp_tc_ts.cu: 28.7 instead of 33.5 ms per 512 x 512 frame), training the
// SS kernel of mlp_tc_2sm.cu (activations in shared memory, image stores through the TMA: 1.40 against 1.45 ms per
// 6144-ray step - with the image stores both kernels are paced by the store path and the L2).
// HN_TC_FWD=ts / ss forces one kernel for both.
static inline int ts_forward_mode() {        // 0 auto, 1 ts, 2 ss
  static const int v = [] {
    const char* e = getenv("HN_TC_FWD");
    if (e == nullptr) return 0;
    return e[0] == 't' ? 1 : (e[0] == 's' ? 2 : 0);
  }();
  return v;
}
static inline bool use_ts_forward(bool save_for_backward) {
  const int m = ts_forward_mode();
  return m == 1 || (m == 0 && !save_for_backward);
}

int64_t tc_train_workspace_bytes(const hn_field_params* p, int64_t n_samples) {
  TcPlan plan;
  if (make_tc_plan(p, plan) != 0) return -1;
  return carve_train_bufs(plan, n_samples, nullptr, nullptr) + 256;
}

}  // namespace hn

using namespace hn;

extern "C" {

int hn_has_tcgen05(void) { return 1; }

int64_t hn_field_packed_bytes(const hn_field_params* p) {
  TcPlan plan;
  if (make_tc_plan(p, plan) != 0) return -1;
  return (int64_t)plan.hdr.total_bytes;
}

int hn_field_pack_bf16(const hn_field_params* p, const float* cond, void* packed, void* stream) {
  TcPlan plan;
  HN_TRY(make_tc_plan(p, plan));
  if (!packed) return fail_arg("hn_field_pack_bf16: null destination");
  if (p->cond_dim > 0 && !cond) return fail_arg("hn_field_pack_bf16: cond missing");
  for (int l = 0; l < p->nr_layers; ++l)
    if (!p->nr_w[l] || !p->nr_b[l]) return fail_arg("hn_field_pack_bf16: null non-rigid layer %d", l);
  for (int l = 0; l < p->cn_layers; ++l)
    if (!p->cn_w[l] || !p->cn_b[l]) return fail_arg("hn_field_pack_bf16: null canonical layer %d", l);
  cudaStream_t st = as_stream(stream);
  uint8_t* dst = static_cast<uint8_t*>(packed);
  // header, chunk table and job table are staged from pageable host memory: the copies are complete
  // (staged) when cudaMemcpyAsync returns, so the vectors may die with this frame
  HN_TRY(check_cuda(cudaMemcpyAsync(dst, &plan.hdr, sizeof(TcHeader), cudaMemcpyHostToDevice, st), "pack/header"));
  HN_TRY(check_cuda(cudaMemcpyAsync(dst + plan.hdr.table_off, plan.chunks.data(),
                                    plan.chunks.size() * sizeof(TcChunk), cudaMemcpyHostToDevice, st), "pack/table"));
  HN_TRY(check_cuda(cudaMemcpyAsync(dst + plan.hdr.bwd_table_off, plan.bwd_chunks.data(),
                                    plan.bwd_chunks.size() * sizeof(TcChunk), cudaMemcpyHostToDevice, st), "pack/bwd table"));
  HN_TRY(check_cuda(cudaMemcpyAsync(dst + plan.jobs_off, plan.jobs.data(), plan.jobs.size() * sizeof(PackJob),
                                    cudaMemcpyHostToDevice, st), "pack/jobs"));
  dim3 grid(8, (unsigned)plan.jobs.size());
  pack_weights_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const PackJob*>(dst + plan.jobs_off),
                                            (int)plan.jobs.size(), cond, dst + plan.hdr.image_off);
  return launched("hn_field_pack_bf16");
}

int hn_field_repack_bf16(const hn_field_params* p, const float* cond, void* packed, void* stream) {
  TcPlan plan;
  HN_TRY(make_tc_plan(p, plan));
  if (!packed) return fail_arg("hn_field_repack_bf16: null image");
  if (p->cond_dim > 0 && !cond) return fail_arg("hn_field_repack_bf16: cond missing");
  uint8_t* dst = static_cast<uint8_t*>(packed);
  dim3 grid(8, (unsigned)plan.jobs.size());
  pack_weights_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const PackJob*>(dst + plan.jobs_off),
                                                           (int)plan.jobs.size(), cond, dst + plan.hdr.image_off);
  return launched("hn_field_repack_bf16");
}

int hn_field_fwd_fp32(const hn_field_params* p, int64_t n_samples, const float* x_skel, const float* cond,
                      const float* hann_w, int ignore_non_rigid, float* raw, float* xyz, float* offsets,
                      void* workspace, int64_t workspace_bytes, int save_for_backward, void* stream);

int hn_field_fwd(const hn_field_params* p, int64_t n_samples, const float* x_skel, const float* cond,
                 const float* hann_w, int ignore_non_rigid, float* raw, float* xyz, float* offsets,
                 void* workspace, int64_t workspace_bytes, int save_for_backward, int precision,
                 const void* packed, const int64_t* n_samples_dev, void* stream) {
  if (precision == HN_PRECISION_FP32 && n_samples_dev)
    return fail_arg("hn_field_fwd: a device-side sample count needs HN_PRECISION_BF16");
  if (precision == HN_PRECISION_FP32)
    return hn_field_fwd_fp32(p, n_samples, x_skel, cond, hann_w, ignore_non_rigid, raw, xyz, offsets, workspace,
                             workspace_bytes, save_for_backward, stream);
  if (precision != HN_PRECISION_BF16) return fail_arg("hn_field_fwd: unknown precision %d", precision);
  if (!packed) return fail_arg("hn_field_fwd: bf16 path needs the image from hn_field_pack_bf16");
  if (n_samples == 0) return 0;
  if (n_samples < 0 || !x_skel || !raw || !xyz || !offsets) return fail_arg("hn_field_fwd: null pointer");
  if (!ignore_non_rigid && !hann_w) return fail_arg("hn_field_fwd: hann_w missing");
  TcPlan plan;
  HN_TRY(make_tc_plan(p, plan));
  int dev = 0, sms = 0;
  HN_TRY(check_cuda(cudaGetDevice(&dev), "hn_field_fwd/device"));
  HN_TRY(check_cuda(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "hn_field_fwd/attr"));
  // the cta_group::2 kernel of mlp_tc_2sm.cu: CTA pairs share every weight chunk
  if (sms < 2) return fail_arg("hn_field_fwd: the tcgen05 forward kernel needs a device with thread-block clusters");
  TrainBufs bufs;
  if (save_for_backward) {
    const int64_t need = carve_train_bufs(plan, n_samples, static_cast<uint8_t*>(workspace), &bufs);
    if (!workspace || workspace_bytes < need) return fail_arg("hn_field_fwd: bf16 training workspace too small");
    if (!use_relu_bits()) memset(bufs.bits, 0, sizeof(bufs.bits));
  }
  // rendering with the reference widths: the two-tile TS kernel (HN_TC_TS2=0 keeps the one-tile kernel)
  static const bool ts2_on = [] { const char* v = getenv("HN_TC_TS2"); return v == nullptr || atoi(v) != 0; }();
  if (use_ts_forward(save_for_backward != 0) && !save_for_backward && ts2_on && !plan.ts_chunks.empty() && plan.width[0] == 128 &&
      plan.width[1] == 256)
    return launch_field_ts2(plan, static_cast<const uint8_t*>(packed), n_samples, n_samples_dev, x_skel, hann_w,
                            ignore_non_rigid ? 1 : 0, raw, xyz, offsets, sms, as_stream(stream));
  if (use_ts_forward(save_for_backward != 0) && !plan.ts_chunks.empty())
    return launch_field_ts(plan, static_cast<const uint8_t*>(packed), n_samples, n_samples_dev, x_skel, hann_w,
                           ignore_non_rigid ? 1 : 0, raw, xyz, offsets, save_for_backward ? &bufs : nullptr, sms,
                           as_stream(stream));
  return launch_field_tc2(plan, static_cast<const uint8_t*>(packed), n_samples, n_samples_dev, x_skel, hann_w,
                          ignore_non_rigid ? 1 : 0, raw, xyz, offsets, save_for_backward ? &bufs : nullptr, sms,
                          as_stream(stream));
}

}  // extern "C"
