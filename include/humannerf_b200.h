/*
 * humannerf_b200 - C ABI of the B200-native HumanNeRF per-ray rendering hot path.
 *
 * One shared library (libhumannerf_b200.so), extern "C", plain pointers and sizes only.
 * The reference (ChenYutongTHU/humannerf) is pure Python and has no FFI of its own; each
 * entry point below replaces one group of eager PyTorch / numpy calls of the reference and
 * cites it (paths relative to the reference root).  A Python host binds these with ctypes
 * (humannerf_b200/_lib.py); INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in _host;
 *  - the caller owns all memory (inputs, outputs, workspaces); nothing is allocated or freed here;
 *    every op that needs scratch memory has a *_workspace_bytes query;
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*) and never synchronises;
 *  - return value: 0 = ok, < 0 = argument error, > 0 = cudaError_t; text via hn_last_error()
 *    (thread local).  No entry point exits or aborts the process;
 *  - float = IEEE fp32 row-major; "nullable" pointers may be NULL to skip that input / output;
 *  - entry points keep no mutable global state (only function attributes set once per device).
 */
#ifndef HUMANNERF_B200_H
#define HUMANNERF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HN_ABI_VERSION 6
#define HN_MAX_LAYERS 16

/* ABI version of the loaded library (compare with HN_ABI_VERSION). */
int hn_version(void);
/* Message of the last failing call on this thread ("" if none). */
const char* hn_last_error(void);
/* Number of kernels this library has launched in this process (all threads). */
int64_t hn_launch_count(void);
/* 1 if the tcgen05/TMEM fused-MLP kernel was compiled in (sm_100a build). */
int hn_has_tcgen05(void);

/* ------------------------------------------------------------------------------------------
 * a1  core/utils/camera_util.py:132-159  get_rays_from_KRT
 * Pixel (x, y, 1) -> K^-1 -> camera -> world; origin = -R^T T; d = pixel_world - origin (not
 * normalised).  float64 arithmetic.  K_host: 9 doubles row-major; E_host: 16 doubles row-major
 * (world->camera 4x4).  rays_o, rays_d: (H*W, 3) double, row-major pixel order.
 */
int hn_gen_rays(int H, int W, const double* K_host, const double* E_host,
                double* rays_o, double* rays_d, void* stream);

/* a2  core/utils/camera_util.py:162-208  rays_intersect_3d_bbox
 * bounds_host: 6 doubles (min xyz, max xyz), padded by 0.01 inside.  |d| < 1e-5 components are
 * overwritten with 1e-5 IN PLACE in rays_d (the reference mutates the array the network later
 * consumes).  mask[i] = 1 iff exactly two of the six plane hits lie in the eps-grown box; near /
 * far (double, dense, defined where mask = 1) = |p - o| / |d| min / max.  Bit-exact float64.
 */
int hn_rays_bbox(int64_t n_rays, const double* bounds_host, const double* rays_o, double* rays_d,
                 double* near, double* far, uint8_t* mask, void* stream);

/* Stream compaction of the rays selected by `mask`, keeping row-major pixel order
 * (core/data/human_nerf/train.py:592-616: rays[ray_mask], near/far .astype(float32)).
 * Outputs have capacity n_rays; *count (device int64) receives R.  pixel_index[r] = source pixel.
 */
int64_t hn_compact_rays_workspace_bytes(int64_t n_rays);
int hn_compact_rays(int64_t n_rays, const uint8_t* mask, const double* rays_o, const double* rays_d,
                    const double* near, const double* far,
                    float* out_o, float* out_d, float* out_near, float* out_far,
                    int32_t* pixel_index, int64_t* count, void* workspace, void* stream);

/* a9  core/nets/human_nerf/network.py:455-471, 499
 * z[r][s] = near*(1-t_s) + far*t_s with t = t_table (S floats, torch.linspace(0,1,S));
 * if t_rand (R,S) != NULL: stratified jitter between neighbouring mid points.  Separate
 * multiply / add roundings (no FMA contraction) -> bit-exact with the eager reference.
 */
int hn_sample_z(int64_t n_rays, int n_samples, const float* near, const float* far,
                const float* t_table, const float* t_rand /*nullable*/, float* z, void* stream);

/* a10  core/nets/human_nerf/network.py:392-444  _sample_motion_fields  (+ pts = o + d*z, :499)
 * Either pts (P,3) is given, or pts == NULL and (rays_o, rays_d (R,3), z (R,S)) are.
 * Rs (J,3,3), Ts (J,3): motion bases; vol (>=J, V, V, V) blend-weight volume indexed [z][y][x];
 * bbox_min, bbox_scale: 3 floats each (device).  Trilinear, align_corners=True, zero padding.
 * Outputs: x_skel (P,3), mask (P) = sum_j w_j, weights (P,J) nullable.  J <= 32.
 * workspace (nullable, 16-byte aligned, hn_motion_field_workspace_bytes): when given, the call first
 * builds the corner table (per joint and stencil cell the 8 corner values, zero outside the volume, in
 * one 32-byte sector) and the lookups read it instead of the volume; the filled workspace can be handed
 * to hn_motion_field_bwd as `table`.
 */
int64_t hn_motion_field_workspace_bytes(int n_joints, int vol_size);
int hn_motion_field_fwd(int64_t n_rays, int n_samples, int n_joints, int vol_size,
                        const float* pts, const float* rays_o, const float* rays_d, const float* z,
                        const float* Rs, const float* Ts, const float* vol,
                        const float* bbox_min, const float* bbox_scale,
                        float* x_skel, float* mask, float* weights,
                        void* workspace, int64_t workspace_bytes, void* stream);

/* Backward of the above.  d_x_skel (P,3), d_mask (P), d_weights (P,J) are nullable upstream
 * gradients.  x_skel_fwd (P,3) and mask_fwd (P) are the forward's outputs (both nullable together):
 * when given, the kernel skips re-evaluating the 24 lookups that only rebuild them.
 * table (nullable): the corner table a forward call left in its workspace (same vol).
 * grad_workspace (nullable, hn_motion_field_workspace_bytes): scratch in which the volume gradient is scattered
 * as 16-byte vector reductions (corner-table layout) before it is folded into d_vol; without it the kernel
 * issues 8 scalar atomics per sample and joint.
 * Accumulates (+=) into d_vol (J,V,V,V), d_Rs (J,3,3), d_Ts (J,3): caller zeroes.
 */
int hn_motion_field_bwd(int64_t n_rays, int n_samples, int n_joints, int vol_size,
                        const float* pts, const float* rays_o, const float* rays_d, const float* z,
                        const float* Rs, const float* Ts, const float* vol,
                        const float* bbox_min, const float* bbox_scale,
                        const float* d_x_skel, const float* d_mask, const float* d_weights,
                        const float* x_skel_fwd, const float* mask_fwd,
                        float* d_vol, float* d_Rs, float* d_Ts, const void* table,
                        void* grad_workspace, int64_t grad_workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * a11-a14  embedders/{hannw_fourier,fourier}.py, non_rigid_motion_mlps/mlp_offset.py:74-114,
 *          canonical_mlps/mlp_rgb_sigma.py:132-206, network.py:233-327 (_apply_mlp_kernals)
 *
 * "Canonical field": x_skel -> hann-PE -> non-rigid MLP -> xyz = x_skel + offset -> PE ->
 * canonical MLP -> raw (r,g,b,sigma).  Weights are the reference's nn.Linear tensors, (out,in)
 * row-major fp32, in module order.  Layer layout (reference defaults in parentheses):
 *   non-rigid: layer 0 takes [cond (69) | pe (36)]; layer nr_skip (4) takes [h | pe]; the last
 *              layer (index nr_layers-1) maps width -> 3 without activation;
 *   canonical: layer 0 takes pe (63); layer cn_skip (5) takes [pe | h]; the last layer maps
 *              width -> 4 without activation.
 */
typedef struct hn_field_params {
  int32_t nr_layers;   /* Linear layers of the non-rigid MLP incl. output (7) */
  int32_t nr_width;    /* 128 */
  int32_t nr_skip;     /* index of the layer that re-reads the encoding (4); -1 = none */
  int32_t nr_freqs;    /* 6  -> 36-d hann-windowed encoding, no raw x */
  int32_t cond_dim;    /* 69 */
  int32_t cn_layers;   /* Linear layers of the canonical MLP incl. output (9) */
  int32_t cn_width;    /* 256 */
  int32_t cn_skip;     /* 5; -1 = none */
  int32_t cn_freqs;    /* 10 -> 63-d encoding incl. raw x */
  int32_t reserved;
  const float* nr_w[HN_MAX_LAYERS];
  const float* nr_b[HN_MAX_LAYERS];
  const float* cn_w[HN_MAX_LAYERS];
  const float* cn_b[HN_MAX_LAYERS];
} hn_field_params;

/* Gradient destinations, same shapes as the weights; accumulated (+=); caller zeroes. */
typedef struct hn_field_grads {
  float* nr_w[HN_MAX_LAYERS];
  float* nr_b[HN_MAX_LAYERS];
  float* cn_w[HN_MAX_LAYERS];
  float* cn_b[HN_MAX_LAYERS];
} hn_field_grads;

#define HN_PRECISION_FP32 0   /* SIMT fp32 kernels, 1e-4 parity path                       */
#define HN_PRECISION_BF16 1   /* tcgen05 / TMEM fused kernels, bf16 operands, fp32 accumulate */

/* Bytes of workspace hn_field_fwd needs for n_samples.  save_for_backward != 0 keeps every
 * activation hn_field_bwd needs (the same workspace must then be passed to it untouched). */
int64_t hn_field_workspace_bytes(const hn_field_params* p, int64_t n_samples,
                                 int save_for_backward, int precision);

/* cond (cond_dim), hann_w (nr_freqs) are device vectors.  ignore_non_rigid != 0 skips the
 * non-rigid MLP (cfg.ignore_non_rigid_motions, network.py:264,276-277): xyz = x_skel, offsets = 0.
 * Outputs: raw (P,4), xyz (P,3), offsets (P,3).
 * For HN_PRECISION_BF16, `packed` is the bf16 weight image made by hn_field_pack_bf16.
 * n_samples_dev (nullable, HN_PRECISION_BF16 only): device int64 holding the number of samples actually
 * present in the first rows of x_skel (<= n_samples, which then is the capacity every buffer is sized
 * for).  Lets a caller compact samples on the device (empty-space skipping) without a host round trip,
 * e.g. inside a CUDA graph; rows past the count are neither read nor written. */
int hn_field_fwd(const hn_field_params* p, int64_t n_samples, const float* x_skel,
                 const float* cond, const float* hann_w, int ignore_non_rigid,
                 float* raw, float* xyz, float* offsets,
                 void* workspace, int64_t workspace_bytes, int save_for_backward,
                 int precision, const void* packed, const int64_t* n_samples_dev, void* stream);

/* Backward (HN_PRECISION_FP32 path).  d_raw (P,4), d_xyz (P,3) nullable, d_offsets (P,3) nullable.
 * Writes d_x_skel (P,3) and accumulates parameter gradients into `g`. `scratch` must hold
 * hn_field_bwd_scratch_bytes. */
int64_t hn_field_bwd_scratch_bytes(const hn_field_params* p, int64_t n_samples);
int hn_field_bwd(const hn_field_params* p, const hn_field_grads* g, int64_t n_samples,
                 const float* x_skel, const float* cond, const float* hann_w, int ignore_non_rigid,
                 const float* d_raw, const float* d_xyz, const float* d_offsets,
                 float* d_x_skel, const void* workspace, int64_t workspace_bytes,
                 void* scratch, int64_t scratch_bytes, void* stream);

/* Backward on the tensor cores (HN_PRECISION_BF16).  `workspace` is the one hn_field_fwd filled with
 * save_for_backward != 0 and precision BF16 (T16 images of every activation plus, from the cta_group::2
 * forward kernel, a 1-bit ReLU mask image per hidden layer that the chains mask with), `packed` the image the
 * forward used, `xyz` the forward's xyz output.  Runs the fused backward-data chains (tcgen05), the
 * encoding backward, and hn_tile_wgrad jobs for every weight / bias; accumulates into `g`. */
int64_t hn_field_bwd_bf16_scratch_bytes(const hn_field_params* p, int64_t n_samples);
int hn_field_bwd_bf16(const hn_field_params* p, const hn_field_grads* g, int64_t n_samples,
                      const float* x_skel, const float* xyz, const float* cond, const float* hann_w,
                      int ignore_non_rigid, const float* d_raw, const float* d_xyz, const float* d_offsets,
                      float* d_x_skel, void* workspace, int64_t workspace_bytes, const void* packed,
                      void* scratch, int64_t scratch_bytes, const int64_t* n_samples_dev /*nullable, as in
                      hn_field_fwd*/, void* stream);

/* The same in two calls: HN_BWD_DATA runs the backward-data chains (writes d_x_skel and the dZ images into `workspace`),
 * HN_BWD_WEIGHTS the weight / bias gradient kernel that reads those images (same arguments, any later time on a stream
 * ordered after the first call).  Nothing between the loss and the optimizer needs the MLPs' weight gradients, so a
 * multi-GPU trainer issues the all-reduce of the boundary gradients (which the data phase feeds) before the 1.7 ms
 * weight phase and hides it there (parallel.CutBReducer); autograd of mlp_offset.py:74-114 / mlp_rgb_sigma.py:132-206. */
#define HN_BWD_DATA 1
#define HN_BWD_WEIGHTS 2
int hn_field_bwd_bf16_phases(const hn_field_params* p, const hn_field_grads* g, int64_t n_samples,
                             const float* x_skel, const float* xyz, const float* cond, const float* hann_w,
                             int ignore_non_rigid, const float* d_raw, const float* d_xyz, const float* d_offsets,
                             float* d_x_skel, void* workspace, int64_t workspace_bytes, const void* packed,
                             void* scratch, int64_t scratch_bytes, const int64_t* n_samples_dev, int phases, void* stream);

/* bf16 weight image for the tcgen05 kernel: weights re-tiled into the UMMA shared-memory
 * core-matrix layout, in consumption order, with the condition code folded into the first
 * non-rigid bias (W0[:, :cond_dim] . cond + b0 is constant per frame, mlp_offset.py:75-77). */
int64_t hn_field_packed_bytes(const hn_field_params* p);
int hn_field_pack_bf16(const hn_field_params* p, const float* cond, void* packed, void* stream);
/* Refresh the image after the weights / cond changed IN PLACE (same pointers as the hn_field_pack_bf16
 * call that initialised `packed`): one conversion kernel, no host->device copies, so it can be captured
 * in a CUDA graph (a training step re-packs the weights the optimizer just updated). */
int hn_field_repack_bf16(const hn_field_params* p, const float* cond, void* packed, void* stream);

/* ------------------------------------------------------------------------------------------
 * a15  core/nets/human_nerf/network.py:355-388  _raw2outputs
 * raw (R,S,4), mask (R,S), z (R,S), rays_d (R,3), xyz (R,S,3), bgcolor (3, in 0..255).
 * Outputs: rgb (R,3), acc (R), depth (R), weights (R,S) nullable, cnl_xyz (R,3), cnl_rgb (R,3),
 * w_max (R), rgb_on_rays (R,S,3) nullable.  arg-max ties resolve to the first sample.
 */
int hn_composite_fwd(int64_t n_rays, int n_samples, const float* raw, const float* mask,
                     const float* z, const float* rays_d, const float* xyz, const float* bgcolor,
                     float* rgb, float* acc, float* depth, float* weights,
                     float* cnl_xyz, float* cnl_rgb, float* w_max, float* rgb_on_rays, void* stream);

/* Backward w.r.t. raw and mask.  Upstream grads d_rgb (R,3), d_acc (R), d_depth (R),
 * d_weights (R,S) are nullable.  Writes d_raw (R,S,4) and d_mask (R,S). */
int hn_composite_bwd(int64_t n_rays, int n_samples, const float* raw, const float* mask,
                     const float* z, const float* rays_d, const float* bgcolor,
                     const float* d_rgb, const float* d_acc, const float* d_depth,
                     const float* d_weights, float* d_raw, float* d_mask, void* stream);

/* ------------------------------------------------------------------------------------------
 * a16  core/train/trainers/human_nerf/trainer.py:28-37 _unpack_imgs and run.py:49-65
 * Scatter per-ray values (R,C) into a dense image / patch stack pre-filled with `fill` (C floats,
 * device): out[dst_index[r]*C + c] = src[r*C + c].  And the matching gather for the backward.
 * A negative index marks a padding row: the scatter skips it, the gather returns 0 for it.
 */
int hn_scatter_rows(int64_t n_rows, int n_cols, const float* src, const int64_t* dst_index,
                    float* out, void* stream);
int hn_gather_rows(int64_t n_rows, int n_cols, const float* src, const int64_t* src_index,
                   float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * a7  core/nets/human_nerf/mweight_vol_decoders/deconv_vol_decoder.py:25-33 + core/utils/network_util.py:12-50
 * vol = softmax_c(ConvDecoder3D(const_embedding) + log(prior)): Linear(embed, hidden) + LeakyReLU(0.2), then n_conv
 * ConvTranspose3d(k4, s2, p1) layers with channels[0] = hidden -> ... -> channels[n_conv] (LeakyReLU(0.2) between
 * them), volume side 2^n_conv.  fp32 storage, deterministic (no atomics); precision = HN_PRECISION_FP32: products as
 * 3xTF32 (fp32-level accuracy, the 1e-4 parity path), HN_PRECISION_BF16: plain TF32 tensor-core products.  lin_w (hidden, embed) row-major; conv_w[l] is the
 * ConvTranspose3d weight (C_in, C_out, 4, 4, 4) in NDHWC (torch.channels_last_3d) strides, i.e. memory order
 * [C_in][kz][ky][kx][C_out]; d_conv_w[l] is written in the same order.  prior / vol / d_vol: (channels[n_conv], V, V, V)
 * contiguous.  workspace (hn_vol_decoder_workspace_bytes) keeps the activations between forward and backward;
 * scratch (hn_vol_decoder_scratch_bytes) is free after each call.
 */
int64_t hn_vol_decoder_workspace_bytes(int embed, int hidden, int n_conv, const int* channels);
int64_t hn_vol_decoder_scratch_bytes(int embed, int hidden, int n_conv, const int* channels);
int hn_vol_decoder_fwd(int embed, int hidden, int n_conv, const int* channels, const float* embedding, const float* lin_w,
                       const float* lin_b, const float* const* conv_w, const float* const* conv_b, const float* prior,
                       float* vol, void* workspace, int64_t workspace_bytes, void* scratch, int64_t scratch_bytes,
                       int precision, void* stream);
int hn_vol_decoder_bwd(int embed, int hidden, int n_conv, const int* channels, const float* embedding, const float* lin_w,
                       const float* const* conv_w, const float* vol, const float* d_vol, const void* workspace,
                       int64_t workspace_bytes, void* scratch, int64_t scratch_bytes, float* d_embedding, float* d_lin_w,
                       float* d_lin_b, float* const* d_conv_w, float* const* d_conv_b, int precision, void* stream);

/* ------------------------------------------------------------------------------------------
 * a3  core/data/human_nerf/train.py:235-340 get_patch_ray_indices / _get_patch_ray_indices,
 *     :226-233 select_rays, :450-476 sample_patch_rays - patch selection on the device.
 * Per frame: hn_patch_prepare turns ray_mask (H*W uint8, = bbox_mask) and subject_mask (H*W uint8) into inclusive
 * prefix counts of the ray mask, the subject mask and bbox-minus-subject (int32 each, H*W) and counts[3] = their
 * totals (the host needs counts[1], counts[2] to draw `np.random.choice(n)` like the reference).
 * Per step: hn_select_patches consumes the reference's two host randoms per patch - u[i] = np.random.rand(1)[0]
 * (subject if u < subject_ratio, :263) and choice[i] = np.random.choice(n_candidates) (:307) - and writes, for
 * cap = n_patches * patch_size^2 rows: select_inds (compacted ray id per selected ray, -1 padding),
 * patch_masks (n_patches, P, P) uint8, div_indices (n_patches + 1), xy_min (n_patches, 2) = (x_min, y_min),
 * target_patches (n_patches, P, P, 3) cropped from image (H, W, 3), and the gathered batch: out_rays (3, cap, 3) =
 * [o, d, d] from the frame's compacted fp32 rays_o / rays_d (R, 3), out_near / out_far (cap), scatter_index (cap) =
 * flat (patch, y, x) of every ray (-1 padding; zero rays, near = far = 0), n_rays (1, optional) = div_indices[-1].
 * Integer outputs are bit-exact with the reference for the same draws.  n_patches <= 32.
 */
int hn_patch_prepare(int H, int W, const uint8_t* ray_mask, const uint8_t* subject_mask, int32_t* ray_cum,
                     int32_t* subject_cum, int32_t* background_cum, int32_t* counts, void* stream);
int hn_select_patches(int H, int W, int n_patches, int patch_size, double subject_ratio, const uint8_t* ray_mask,
                      const int32_t* ray_cum, const int32_t* subject_cum, const int32_t* background_cum,
                      const double* u, const int64_t* choice, const float* rays_o, const float* rays_d,
                      const float* near, const float* far, const float* image, int32_t* select_inds,
                      uint8_t* patch_masks, int32_t* div_indices, int32_t* xy_min, float* target_patches,
                      float* out_rays, float* out_near, float* out_far, int64_t* scatter_index, int64_t* n_rays,
                      void* stream);

/* ------------------------------------------------------------------------------------------
 * a16  trainer.py:97-113 get_img_rebuild_loss -> third_parties/lpips/__init__.py:40-42, lpips.py:14-15, 78-92
 * The LPIPS head of one trunk stage, fused: unit-normalise the channel vector of every position
 * (x / (sqrt(||x||^2 + 1e-10) + 1e-10), the fork's normalize_tensor), squared difference of prediction and target, non-negative 1x1 `lin` weights (C),
 * spatial mean.  feat: (2 * n_pairs, C, positions) features of one trunk stage with element strides
 * (stride_n, stride_c, stride_p) - prediction images first, then their targets.
 *   out[i] += lpips_stage(feat[i], feat[n_pairs + i])           (accumulates: one call per stage)
 * Backward: d_feat (same strides as feat, both halves) = d_out[i] * d out[i] / d feat.
 */
int hn_lpips_head_fwd(int n_pairs, int channels, int positions, const float* feat, int64_t stride_n,
                      int64_t stride_c, int64_t stride_p, const float* lin, float* out, void* stream);
int hn_lpips_head_bwd(int n_pairs, int channels, int positions, const float* feat, int64_t stride_n,
                      int64_t stride_c, int64_t stride_p, const float* lin, const float* d_out,
                      float* d_feat, void* stream);

/* ------------------------------------------------------------------------------------------
 * T16 tile images (bf16) and tensor-core weight gradients
 *
 * The tcgen05 kernels keep activations H and pre-activation gradients dZ in HBM as "T16 tile
 * images": element (sample s, feature f) of a (P x F) matrix, F a multiple of 8, lives at byte
 *   (s/128)*128*F*2 + ((s%128)/16)*16*F*2 + (f/8)*256 + (s%16)*16 + (f%8)*2
 * (per 16 samples x 8 features one 256-byte block; 16 samples x all features are contiguous).
 * hn_tile_pack / hn_tile_unpack convert from / to fp32 row-major.
 *
 * hn_tile_wgrad runs a list of jobs  dW[0:m, col0:col0+n_valid] += dZ[:, 0:m]^T X[:, x_f0:x_f0+n]
 * (nn.Linear weight gradients, reference mlp_offset.py:74-84, mlp_rgb_sigma.py:163-166,198) on the
 * tensor cores; column `bias_col` of the X window (the constant-1 column of the encoding block)
 * can be routed to a bias gradient db[0:m].  dW / db are accumulated with fp32 atomics.
 */
typedef struct hn_wgrad_job {
  const void* dz;      /* tile image of dZ, f_dz features per sample */
  const void* x;       /* tile image of the layer input, f_x features per sample */
  float* dw;           /* (m, dw_ld) fp32, accumulated */
  float* db;           /* nullable, (m) fp32, accumulated */
  int32_t f_dz, f_x;   /* multiples of 16 in [16,256] */
  int32_t m;           /* rows of dW = leading features of dZ used (<= f_dz) */
  int32_t x_f0;        /* first feature of the X window (multiple of 8) */
  int32_t n;           /* width of the X window = MMA N (multiple of 16, <= 256) */
  int32_t n_valid;     /* columns of the window written to dW */
  int32_t dw_ld, dw_col0;
  int32_t bias_col;    /* >= 0: window column routed to db (the constant-1 column of an encoding block);
                          -1: no db; -2: db = column sums of dZ, accumulated from the operand stages in shared memory */
  int32_t reserved;
} hn_wgrad_job;

int64_t hn_tile_image_bytes(int64_t n_samples, int n_features_padded);
int hn_tile_pack(int64_t n_samples, int n_features, int n_features_padded, const float* src, int64_t ld,
                 void* image, void* stream);
int hn_tile_unpack(int64_t n_samples, int n_features, int n_features_padded, const void* image, float* dst,
                   int64_t ld, void* stream);
/* jobs_host: host array (travels to the kernel as a launch parameter, at most 48 jobs per launch;
 * longer lists are split).  jobs_scratch_dev is unused since ABI 2 and may be NULL. */
int hn_tile_wgrad(const hn_wgrad_job* jobs_host, int n_jobs, int64_t n_samples, void* jobs_scratch_dev,
                  void* stream);

/* ------------------------------------------------------------------------------------------
 * a5 + a6  per-frame front end: RodriguesModule (core/utils/network_util.py:57-83), corrected rotations
 *          Rs[1:] = dst_Rs[1:] @ dR (core/nets/human_nerf/network.py:612-616) and the forward-kinematics
 *          chain of MotionBasisComputer (network_util.py:125-146), one launch forward, one backward.
 * parent_host: n_joints ints on the HOST (parent of every joint, parent[i] < i, parent[0] ignored).
 * rvec (J-1,3) axis-angle corrections of joints 1..J-1 (nullable in the forward: no refinement),
 * dst_Rs (J,3,3), dst_Ts (J,3).  Outputs: Rs_out (J,3,3) corrected local rotations, G (J,4,4) global
 * joint transforms.  The caller inverts G (the reference uses an LU inverse) and multiplies with the
 * canonical transforms.  Backward: d_G (J,4,4) and d_Rs_out (J,3,3, nullable) -> d_rvec (J-1,3).
 */
int hn_pose_fk_fwd(int n_joints, const int32_t* parent_host, const float* rvec, const float* dst_Rs,
                   const float* dst_Ts, float* Rs_out, float* G, void* stream);
int hn_pose_fk_bwd(int n_joints, const int32_t* parent_host, const float* rvec, const float* dst_Rs,
                   const float* Rs_corr, const float* dst_Ts, const float* G, const float* d_G,
                   const float* d_Rs_out, float* d_rvec, void* stream);

/* ------------------------------------------------------------------------------------------
 * a17  core/train/optimizers/human_nerf/optimizer.py:12-43 (Adam, one group per tensor) and
 *      core/train/lr_updaters/exp_decay.py:7-16 (per-step learning-rate decay)
 * Fused multi-tensor Adam: one launch for all tensors.  hn_adam_plan writes a plan (tensor table +
 * chunk table) into HOST memory; the caller copies it to the device once.  The step reads, from DEVICE
 * memory, lr_dev[n_tensors] (per-tensor learning rates), lr_scale_dev (nullable global factor),
 * step_dev[n_tensors] (1-based step count of each tensor, as floats) and active_dev[n_tensors]
 * (nullable; 0 = leave the tensor and its moments untouched, like torch skipping a parameter whose
 * .grad is None), so a captured launch follows the schedule.
 * Update rule = torch.optim.Adam(amsgrad=False, weight_decay=0).
 */
typedef struct hn_adam_tensor {
  float* param;
  const float* grad;      /* may be NULL when hn_adam_step gets grad_ptrs_dev */
  float* exp_avg;
  float* exp_avg_sq;
  int64_t numel;
} hn_adam_tensor;

int64_t hn_adam_plan_bytes(const hn_adam_tensor* tensors_host, int n_tensors);
int hn_adam_plan(const hn_adam_tensor* tensors_host, int n_tensors, void* plan_host, int64_t plan_bytes);
/* n_chunks = (plan_bytes - chunk table offset) / 8 is returned by hn_adam_n_chunks. */
int64_t hn_adam_n_chunks(const hn_adam_tensor* tensors_host, int n_tensors);
/* grad_ptrs_dev (nullable): device array of n_tensors gradient pointers that overrides the plan's `grad`
 * fields for this step - autograd allocates fresh gradient tensors every backward pass, and stealing them
 * is cheaper than accumulating into persistent buffers (no zero-fill, no `grad += new` pass). */
int hn_adam_step(const void* plan_dev, int64_t n_chunks, const void* grad_ptrs_dev, const float* lr_dev,
                 const float* lr_scale_dev, const float* step_dev, const float* active_dev, float beta1,
                 float beta2, float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HUMANNERF_B200_H */
