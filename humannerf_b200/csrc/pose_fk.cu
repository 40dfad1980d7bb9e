// a5 + a6 (per-frame front end): axis-angle corrections -> rotation matrices (RodriguesModule,
// core/utils/network_util.py:57-83), corrected joint rotations Rs[1:] = dst_Rs[1:] @ dR
// (network.py:612-616, :667-695) and forward kinematics over the SMPL tree, G_i = G_parent(i) [R_i T_i; 0 1]
// (MotionBasisComputer, network_util.py:125-146), forward and backward, in ONE tiny launch each.
// In eager PyTorch this is ~300 launches of 1-element kernels per training step (fwd + bwd).
//
// The forward replicates the reference's fp32 expression trees with separate roundings
// (__fmul_rn / __fadd_rn) where eager PyTorch evaluates elementwise, so G agrees with the torch
// implementation to the last bits; the 4x4 inverse and the product with the canonical transforms stay
// torch ops (linalg.inv_ex - same LU as the reference - and one batched matmul).
#include "common.cuh"

namespace hn {

__constant__ int c_parent[kMaxJoints];

struct Rod {
  float theta, x, y, z, c, s;
};

__device__ __forceinline__ Rod rodrigues_terms(float r0, float r1, float r2) {
  Rod t;
  // theta = sqrt(1e-5 + sum(r^2)); torch.sum over 3 elements adds left to right
  const float ss = __fadd_rn(__fadd_rn(__fmul_rn(r0, r0), __fmul_rn(r1, r1)), __fmul_rn(r2, r2));
  t.theta = sqrtf(__fadd_rn(1e-5f, ss));
  t.x = r0 / t.theta; t.y = r1 / t.theta; t.z = r2 / t.theta;
  t.c = cosf(t.theta);
  t.s = sinf(t.theta);
  return t;
}

// the nine entries in the reference's expression order, e.g. x*x + (1 - x*x)*c and x*y*(1 - c) - z*s
__device__ __forceinline__ void rodrigues_matrix(const Rod& t, float R[9]) {
  const float x = t.x, y = t.y, z = t.z, c = t.c, s = t.s;
  const float omc = __fadd_rn(1.0f, -c);
  auto diag = [&](float u) {
    const float uu = __fmul_rn(u, u);
    return __fadd_rn(uu, __fmul_rn(__fadd_rn(1.0f, -uu), c));
  };
  auto off = [&](float u, float v, float w, float sign) {   // u*v*(1-c) + sign*w*s
    const float a = __fmul_rn(__fmul_rn(u, v), omc);
    const float b = __fmul_rn(w, s);
    return sign > 0.0f ? __fadd_rn(a, b) : __fadd_rn(a, -b);
  };
  R[0] = diag(x);             R[1] = off(x, y, z, -1.f);  R[2] = off(x, z, y, +1.f);
  R[3] = off(x, y, z, +1.f);  R[4] = diag(y);             R[5] = off(y, z, x, -1.f);
  R[6] = off(x, z, y, -1.f);  R[7] = off(y, z, x, +1.f);  R[8] = diag(z);
}

__device__ __forceinline__ void mat3_mul(const float* A, const float* B, float* C) {   // C = A B, k ascending
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float acc = A[i * 3 + 0] * B[0 * 3 + j];
      acc = fmaf(A[i * 3 + 1], B[1 * 3 + j], acc);
      acc = fmaf(A[i * 3 + 2], B[2 * 3 + j], acc);
      C[i * 3 + j] = acc;
    }
}

// rvec (J-1,3) nullable (no pose refinement), dst_Rs (J,3,3), dst_Ts (J,3) -> Rs_out (J,3,3) corrected
// rotations, G (J,4,4) global transforms
__global__ void pose_fk_fwd_kernel(int J, const float* __restrict__ rvec, const float* __restrict__ dst_Rs,
                                   const float* __restrict__ dst_Ts, float* __restrict__ Rs_out, float* __restrict__ G) {
  __shared__ float s_L[kMaxJoints][12];     // local [R | T], row major 3x4
  __shared__ float s_G[kMaxJoints][12];
  const int j = threadIdx.x;
  if (j < J) {
    float R[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = dst_Rs[j * 9 + k];
    if (j > 0 && rvec != nullptr) {
      float dR[9], C[9];
      const Rod t = rodrigues_terms(rvec[(j - 1) * 3 + 0], rvec[(j - 1) * 3 + 1], rvec[(j - 1) * 3 + 2]);
      rodrigues_matrix(t, dR);
      mat3_mul(R, dR, C);
#pragma unroll
      for (int k = 0; k < 9; ++k) R[k] = C[k];
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) Rs_out[j * 9 + k] = R[k];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      s_L[j][r * 4 + 0] = R[r * 3 + 0]; s_L[j][r * 4 + 1] = R[r * 3 + 1]; s_L[j][r * 4 + 2] = R[r * 3 + 2];
      s_L[j][r * 4 + 3] = dst_Ts[j * 3 + r];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 0; k < 12; ++k) s_G[0][k] = s_L[0][k];
    for (int i = 1; i < J; ++i) {
      const float* P = s_G[c_parent[i]];
      const float* L = s_L[i];
      // [Pr Pt; 0 1] [Lr Lt; 0 1] = [Pr Lr, Pr Lt + Pt; 0 1]; dot products over k = 0..3 like a 4x4 matmul
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) {
          float acc = P[r * 4 + 0] * L[0 * 4 + c];
          acc = fmaf(P[r * 4 + 1], L[1 * 4 + c], acc);
          acc = fmaf(P[r * 4 + 2], L[2 * 4 + c], acc);
          if (c == 3) acc = fmaf(P[r * 4 + 3], 1.0f, acc);
          s_G[i][r * 4 + c] = acc;
        }
    }
  }
  __syncthreads();
  if (j < J) {
#pragma unroll
    for (int k = 0; k < 12; ++k) G[j * 16 + k] = s_G[j][k];
    G[j * 16 + 12] = 0.f; G[j * 16 + 13] = 0.f; G[j * 16 + 14] = 0.f; G[j * 16 + 15] = 1.f;
  }
}

// d_G (J,4,4) and d_Rs_out (J,3,3, nullable) -> d_rvec (J-1,3).  Reverse sweep over the tree:
//   dL_i = G_parent^T dG_i (rotation / translation blocks), dG_parent += dG_i L_i^T,
// then dR_i -> d(dR_i) = dst_R_i^T dR_i -> Rodrigues backward.
__global__ void pose_fk_bwd_kernel(int J, const float* __restrict__ rvec, const float* __restrict__ dst_Rs,
                                   const float* __restrict__ Rs_corr, const float* __restrict__ dst_Ts,
                                   const float* __restrict__ G, const float* __restrict__ d_G,
                                   const float* __restrict__ d_Rs_out, float* __restrict__ d_rvec) {
  __shared__ float s_dG[kMaxJoints][12];
  __shared__ float s_dR[kMaxJoints][9];
  const int j = threadIdx.x;
  if (j < J)
    for (int k = 0; k < 12; ++k) s_dG[j][k] = d_G[j * 16 + k];
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = J - 1; i >= 1; --i) {
      const int p = c_parent[i];
      const float* Gp = G + p * 16;
      const float* dGi = s_dG[i];
      float Lr[9], Lt[3];
      for (int k = 0; k < 9; ++k) Lr[k] = Rs_corr[i * 9 + k];
      for (int k = 0; k < 3; ++k) Lt[k] = dst_Ts[i * 3 + k];
      // dLr = Pr^T dGr   (only the rotation block of the local transform carries a gradient)
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
          s_dR[i][r * 3 + c] = Gp[0 * 4 + r] * dGi[0 * 4 + c] + Gp[1 * 4 + r] * dGi[1 * 4 + c] + Gp[2 * 4 + r] * dGi[2 * 4 + c];
      // dPr += dGr Lr^T + dGt Lt^T ; dPt += dGt
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c)
          s_dG[p][r * 4 + c] += dGi[r * 4 + 0] * Lr[c * 3 + 0] + dGi[r * 4 + 1] * Lr[c * 3 + 1] + dGi[r * 4 + 2] * Lr[c * 3 + 2] +
                                dGi[r * 4 + 3] * Lt[c];
        s_dG[p][r * 4 + 3] += dGi[r * 4 + 3];
      }
    }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) s_dR[0][r * 3 + c] = s_dG[0][r * 4 + c];
  }
  __syncthreads();
  if (j >= 1 && j < J) {
    float dR[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) dR[k] = s_dR[j][k] + (d_Rs_out ? d_Rs_out[j * 9 + k] : 0.0f);
    // R_corr = dst_R @ dRod  =>  d(dRod) = dst_R^T dR
    float D[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        D[r * 3 + c] = dst_Rs[j * 9 + 0 * 3 + r] * dR[0 * 3 + c] + dst_Rs[j * 9 + 1 * 3 + r] * dR[1 * 3 + c] +
                       dst_Rs[j * 9 + 2 * 3 + r] * dR[2 * 3 + c];
    const float r0 = rvec[(j - 1) * 3 + 0], r1 = rvec[(j - 1) * 3 + 1], r2 = rvec[(j - 1) * 3 + 2];
    const Rod t = rodrigues_terms(r0, r1, r2);
    const float x = t.x, y = t.y, z = t.z, c = t.c, s = t.s, omc = 1.0f - c;
    const float s01 = D[1] + D[3], s02 = D[2] + D[6], s12 = D[5] + D[7];
    const float d_c = D[0] * (1.f - x * x) + D[4] * (1.f - y * y) + D[8] * (1.f - z * z) - s01 * x * y - s02 * x * z - s12 * y * z;
    const float d_s = -z * D[1] + y * D[2] + z * D[3] - x * D[5] - y * D[6] + x * D[7];
    const float d_x = D[0] * (2.f * x * omc) + s01 * y * omc + s02 * z * omc + (D[7] - D[5]) * s;
    const float d_y = D[4] * (2.f * y * omc) + s01 * x * omc + s12 * z * omc + (D[2] - D[6]) * s;
    const float d_z = D[8] * (2.f * z * omc) + s02 * x * omc + s12 * y * omc + (D[3] - D[1]) * s;
    const float d_theta = -s * d_c + c * d_s - (x * d_x + y * d_y + z * d_z) / t.theta;
    d_rvec[(j - 1) * 3 + 0] = d_x / t.theta + d_theta * x;
    d_rvec[(j - 1) * 3 + 1] = d_y / t.theta + d_theta * y;
    d_rvec[(j - 1) * 3 + 2] = d_z / t.theta + d_theta * z;
  }
}

static int upload_parents(const int32_t* parent_host, int J) {
  static int cached[kMaxJoints];
  static int cached_n = 0;
  bool same = cached_n == J;
  for (int i = 0; same && i < J; ++i) same = cached[i] == parent_host[i];
  if (same) return 0;
  for (int i = 1; i < J; ++i)
    if (parent_host[i] < 0 || parent_host[i] >= i) return fail_arg("hn_pose_fk: parent[%d] = %d is not an earlier joint", i, parent_host[i]);
  int tmp[kMaxJoints] = {0};
  for (int i = 0; i < J; ++i) tmp[i] = parent_host[i] < 0 ? 0 : parent_host[i];
  HN_TRY(check_cuda(cudaMemcpyToSymbol(c_parent, tmp, sizeof(int) * kMaxJoints), "hn_pose_fk/parents"));
  for (int i = 0; i < J; ++i) cached[i] = parent_host[i];
  cached_n = J;
  return 0;
}

}  // namespace hn

using namespace hn;

extern "C" {

int hn_pose_fk_fwd(int n_joints, const int32_t* parent_host, const float* rvec, const float* dst_Rs, const float* dst_Ts,
                   float* Rs_out, float* G, void* stream) {
  if (n_joints < 1 || n_joints > kMaxJoints) return fail_arg("hn_pose_fk_fwd: n_joints=%d outside [1,%d]", n_joints, kMaxJoints);
  if (!parent_host || !dst_Rs || !dst_Ts || !Rs_out || !G) return fail_arg("hn_pose_fk_fwd: null pointer");
  HN_TRY(upload_parents(parent_host, n_joints));
  pose_fk_fwd_kernel<<<1, 32, 0, as_stream(stream)>>>(n_joints, rvec, dst_Rs, dst_Ts, Rs_out, G);
  return launched("hn_pose_fk_fwd");
}

int hn_pose_fk_bwd(int n_joints, const int32_t* parent_host, const float* rvec, const float* dst_Rs, const float* Rs_corr,
                   const float* dst_Ts, const float* G, const float* d_G, const float* d_Rs_out, float* d_rvec,
                   void* stream) {
  if (n_joints < 1 || n_joints > kMaxJoints) return fail_arg("hn_pose_fk_bwd: n_joints=%d outside [1,%d]", n_joints, kMaxJoints);
  if (!parent_host || !rvec || !dst_Rs || !Rs_corr || !dst_Ts || !G || !d_G || !d_rvec) return fail_arg("hn_pose_fk_bwd: null pointer");
  HN_TRY(upload_parents(parent_host, n_joints));
  pose_fk_bwd_kernel<<<1, 32, 0, as_stream(stream)>>>(n_joints, rvec, dst_Rs, Rs_corr, dst_Ts, G, d_G, d_Rs_out, d_rvec);
  return launched("hn_pose_fk_bwd");
}

}  // extern "C"
