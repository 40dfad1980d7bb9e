// Fused multi-tensor Adam: ONE launch updates every parameter tensor of the network.
//
// The reference builds torch.optim.Adam with one parameter group per tensor (55 groups,
// core/train/optimizers/human_nerf/optimizer.py:12-43) and rescales every group's learning rate
// after each step (lr_updaters/exp_decay.py:7-16); on a GPU that is ~400 tiny launches per step for
// 64.4 M parameters.  Here a plan (tensor table + chunk table, built once on the host by
// hn_adam_plan and uploaded by the caller) drives one HBM-bound kernel: 28 bytes per element
// (read p, g, m, v; write p, m, v; 12 bytes for elements whose gradient and moments are all zero).  The step count, the per-tensor learning rates and an optional
// global learning-rate scale are read from DEVICE memory, so the launch can be captured in a CUDA
// graph and replayed while the schedule advances.
//
// Arithmetic = torch.optim.Adam (amsgrad=False, weight_decay=0, maximize=False):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace hn {

constexpr int kAdamChunk = 8192;      // elements per chunk (multiple of 4 * 256)
constexpr int kAdamThreads = 256;

struct AdamHeader {
  int32_t n_tensors;
  int32_t reserved;
  int64_t n_chunks;
  int64_t tensors_off;   // byte offset of the hn_adam_tensor table
  int64_t chunks_off;    // byte offset of the int2 (tensor, chunk-in-tensor) table
};

__global__ void __launch_bounds__(kAdamThreads)
adam_kernel(const uint8_t* __restrict__ plan, const float* const* __restrict__ grad_ptrs,
            const float* __restrict__ lr, const float* __restrict__ lr_scale,
            const float* __restrict__ step, const float* __restrict__ active, float beta1, float beta2, float eps) {
  const AdamHeader* hdr = reinterpret_cast<const AdamHeader*>(plan);
  const hn_adam_tensor* tensors = reinterpret_cast<const hn_adam_tensor*>(plan + hdr->tensors_off);
  const int2* chunks = reinterpret_cast<const int2*>(plan + hdr->chunks_off);
  __shared__ float s_c[2];
  const float scale = lr_scale ? lr_scale[0] : 1.0f;
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
  for (int64_t c = blockIdx.x; c < hdr->n_chunks; c += gridDim.x) {
    const int2 ch = chunks[c];
    if (active && active[ch.x] == 0.0f) continue;
    hn_adam_tensor t = tensors[ch.x];
    if (grad_ptrs) t.grad = grad_ptrs[ch.x];          // this step's gradient buffers (autograd re-allocates them)
    // bias corrections in double, like the host arithmetic of torch.optim.Adam (one thread per chunk;
    // the chunk loop and the `continue` above are block-uniform)
    __syncthreads();
    if (threadIdx.x == 0) {
      const double tt = (double)step[ch.x];
      s_c[0] = (float)(1.0 / (1.0 - pow((double)beta1, tt)));          // 1 / bias_correction1
      s_c[1] = (float)(1.0 / sqrt(1.0 - pow((double)beta2, tt)));      // 1 / sqrt(bias_correction2)
    }
    __syncthreads();
    const float s_c1 = s_c[0], s_c2 = s_c[1];
    const float step_size = lr[ch.x] * scale * s_c1;
    const int64_t begin = (int64_t)ch.y * kAdamChunk;
    const int64_t end = min(begin + kAdamChunk, t.numel);
    const bool vec = ((reinterpret_cast<uintptr_t>(t.param) | reinterpret_cast<uintptr_t>(t.grad) |
                       reinterpret_cast<uintptr_t>(t.exp_avg) | reinterpret_cast<uintptr_t>(t.exp_avg_sq)) & 15) == 0;
    if (vec && end - begin == kAdamChunk) {
      float4* p4 = reinterpret_cast<float4*>(t.param + begin);
      const float4* g4 = reinterpret_cast<const float4*>(t.grad + begin);
      float4* m4 = reinterpret_cast<float4*>(t.exp_avg + begin);
      float4* v4 = reinterpret_cast<float4*>(t.exp_avg_sq + begin);
#pragma unroll 4
      for (int i = threadIdx.x; i < kAdamChunk / 4; i += kAdamThreads) {
        float4 m = m4[i], v = v4[i];
        const float4 g = __ldg(g4 + i);
        // g = m = v = +0 for all four elements: the update is exactly the identity (m' = v' = 0, p' = p - 0), so neither
        // p is read nor anything written: 12 instead of 28 bytes per element.  That is the steady state of 29.4 M of the
        // 64.4 M parameters: 56 of the 64 taps of the first transposed convolution never reach its 2^3 output
        // (network_util.py:27-29, kernel 4 / stride 2 / padding 1 on a 1^3 input) and have a zero gradient for ever.
        if (((__float_as_uint(g.x) | __float_as_uint(g.y) | __float_as_uint(g.z) | __float_as_uint(g.w)) |
             (__float_as_uint(m.x) | __float_as_uint(m.y) | __float_as_uint(m.z) | __float_as_uint(m.w)) |
             (__float_as_uint(v.x) | __float_as_uint(v.y) | __float_as_uint(v.z) | __float_as_uint(v.w))) == 0u)
          continue;
        float4 p = p4[i];
        m.x = beta1 * m.x + omb1 * g.x; m.y = beta1 * m.y + omb1 * g.y;
        m.z = beta1 * m.z + omb1 * g.z; m.w = beta1 * m.w + omb1 * g.w;
        v.x = beta2 * v.x + omb2 * g.x * g.x; v.y = beta2 * v.y + omb2 * g.y * g.y;
        v.z = beta2 * v.z + omb2 * g.z * g.z; v.w = beta2 * v.w + omb2 * g.w * g.w;
        p.x -= step_size * (m.x / (sqrtf(v.x) * s_c2 + eps));
        p.y -= step_size * (m.y / (sqrtf(v.y) * s_c2 + eps));
        p.z -= step_size * (m.z / (sqrtf(v.z) * s_c2 + eps));
        p.w -= step_size * (m.w / (sqrtf(v.w) * s_c2 + eps));
        p4[i] = p; m4[i] = m; v4[i] = v;
      }
    } else {
      for (int64_t i = begin + threadIdx.x; i < end; i += kAdamThreads) {
        const float g = t.grad[i];
        const float m = beta1 * t.exp_avg[i] + omb1 * g;
        const float v = beta2 * t.exp_avg_sq[i] + omb2 * g * g;
        t.exp_avg[i] = m;
        t.exp_avg_sq[i] = v;
        t.param[i] -= step_size * (m / (sqrtf(v) * s_c2 + eps));
      }
    }
  }
}

static int64_t adam_chunks(const hn_adam_tensor* t, int n) {
  int64_t c = 0;
  for (int i = 0; i < n; ++i) c += ceil_div<int64_t>(t[i].numel, kAdamChunk);
  return c;
}

}  // namespace hn

using namespace hn;

extern "C" {

int64_t hn_adam_plan_bytes(const hn_adam_tensor* tensors_host, int n_tensors) {
  if (!tensors_host || n_tensors <= 0) return -1;
  for (int i = 0; i < n_tensors; ++i)
    if (tensors_host[i].numel < 0) return -1;
  const int64_t tab = ceil_div<int64_t>((int64_t)sizeof(AdamHeader) + (int64_t)n_tensors * sizeof(hn_adam_tensor), 16) * 16;
  return tab + adam_chunks(tensors_host, n_tensors) * (int64_t)sizeof(int2);
}

int64_t hn_adam_n_chunks(const hn_adam_tensor* tensors_host, int n_tensors) {
  if (!tensors_host || n_tensors <= 0) return -1;
  return adam_chunks(tensors_host, n_tensors);
}

int hn_adam_plan(const hn_adam_tensor* tensors_host, int n_tensors, void* plan_host, int64_t plan_bytes) {
  const int64_t need = hn_adam_plan_bytes(tensors_host, n_tensors);
  if (need < 0 || !plan_host || plan_bytes < need) return fail_arg("hn_adam_plan: bad argument or buffer too small");
  for (int i = 0; i < n_tensors; ++i) {
    const hn_adam_tensor& t = tensors_host[i];
    if (t.numel > 0 && (!t.param || !t.exp_avg || !t.exp_avg_sq))
      return fail_arg("hn_adam_plan: null pointer in tensor %d", i);
    if (ceil_div<int64_t>(t.numel, kAdamChunk) > 0x7fffffff) return fail_arg("hn_adam_plan: tensor %d too large", i);
  }
  uint8_t* out = static_cast<uint8_t*>(plan_host);
  AdamHeader h;
  memset(&h, 0, sizeof(h));
  h.n_tensors = n_tensors;
  h.n_chunks = adam_chunks(tensors_host, n_tensors);
  h.tensors_off = sizeof(AdamHeader);
  h.chunks_off = ceil_div<int64_t>((int64_t)sizeof(AdamHeader) + (int64_t)n_tensors * sizeof(hn_adam_tensor), 16) * 16;
  memcpy(out, &h, sizeof(h));
  memcpy(out + h.tensors_off, tensors_host, (size_t)n_tensors * sizeof(hn_adam_tensor));
  int2* chunks = reinterpret_cast<int2*>(out + h.chunks_off);
  int64_t c = 0;
  for (int i = 0; i < n_tensors; ++i)
    for (int64_t k = 0; k < ceil_div<int64_t>(tensors_host[i].numel, kAdamChunk); ++k) chunks[c++] = make_int2(i, (int)k);
  return 0;
}

int hn_adam_step(const void* plan_dev, int64_t n_chunks, const void* grad_ptrs_dev, const float* lr_dev,
                 const float* lr_scale_dev, const float* step_dev, const float* active_dev, float beta1,
                 float beta2, float eps, void* stream) {
  if (n_chunks == 0) return 0;
  if (!plan_dev || n_chunks < 0 || !lr_dev || !step_dev) return fail_arg("hn_adam_step: null pointer");
  int dev = 0, sms = 0;
  HN_TRY(check_cuda(cudaGetDevice(&dev), "hn_adam_step/device"));
  HN_TRY(check_cuda(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "hn_adam_step/attr"));
  const unsigned grid = (unsigned)std::min<int64_t>(n_chunks, (int64_t)sms * 8);
  adam_kernel<<<grid, kAdamThreads, 0, as_stream(stream)>>>(static_cast<const uint8_t*>(plan_dev),
                                                            static_cast<const float* const*>(grad_ptrs_dev), lr_dev, lr_scale_dev,
                                                            step_dev, active_dev, beta1, beta2, eps);
  return launched("hn_adam_step");
}

}  // extern "C"
