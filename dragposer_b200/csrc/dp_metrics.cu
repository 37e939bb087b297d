// dp_metrics.cu -- batched accuracy metrics on the device.
//
// Reference: eval_pos_error (python/src/eval_metrics.py:6-32): forward kinematics of the ground-truth and of the result
// skeleton with the root at the origin, mean joint distance (MPJPE) and mean distance of the sparse end effectors
// (MPEEPE, joints 4, 8, 13, 17, 21).  Poses come in the engine's own output format: (n,88) standardised root-space
// quaternions whose root slot holds the standardised WORLD root rotation, so results can be scored without leaving the GPU
// format; the kinematics are the closed form of the frame kernels (R_j = R_0 M(q_j), p_j = p_parent + R_parent o_j).
// One warp per pose pair, lane == joint.
#include "dp_common.cuh"
#include "dp_internal.h"

namespace {

__device__ __forceinline__ void pose_positions(const DpModelImage& M, const float* __restrict__ pose, int lane, float p[3]) {
  const bool is_joint = lane < DP_J;
  const float4 y = is_joint ? reinterpret_cast<const float4*>(pose)[lane] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 mq = is_joint ? reinterpret_cast<const float4*>(M.mean_q)[lane] : make_float4(1.f, 0.f, 0.f, 0.f);
  const float4 sq = is_joint ? reinterpret_cast<const float4*>(M.std_q)[lane] : make_float4(0.f, 0.f, 0.f, 0.f);
  float q[4] = {fmaf(y.x, sq.x, mq.x), fmaf(y.y, sq.y, mq.y), fmaf(y.z, sq.z, mq.z), fmaf(y.w, sq.w, mq.w)};
  const float inv = 1.0f / (sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]) + 1e-8f);
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] *= inv;
  float r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = __shfl_sync(0xffffffffu, q[i], 0);  // world root rotation
  float R0[9], Mj[9], R[9];
  quat_to_mat(r, R0);
  const float ident[4] = {1.f, 0.f, 0.f, 0.f};
  quat_to_mat(lane == 0 ? ident : q, Mj);
  mat_mul(R0, Mj, R);
  const int par = lane == 0 ? 0 : M.parent[lane];
  float Rp[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) Rp[i] = __shfl_sync(0xffffffffu, R[i], par);
  const float ov[3] = {M.off[lane][0], M.off[lane][1], M.off[lane][2]};
  mat_vec(Rp, ov, p);
  if (lane == 0) p[0] = p[1] = p[2] = 0.0f;  // root at the origin
  const int n_jump = M.pad[0];
  for (int rd = 0; rd < DP_JUMP_ROUNDS && rd < n_jump; ++rd) {
    const int a = M.jump[rd][lane];
    const int src = a >= 0 ? a : lane;
    const float t0 = __shfl_sync(0xffffffffu, p[0], src), t1 = __shfl_sync(0xffffffffu, p[1], src), t2 = __shfl_sync(0xffffffffu, p[2], src);
    if (a >= 0) { p[0] += t0; p[1] += t1; p[2] += t2; }
  }
}

__global__ void __launch_bounds__(128) dp_pose_error_kernel(const DpModelImage* __restrict__ model, const float* __restrict__ pose_a,
                                                            const float* __restrict__ pose_b, int n, float* __restrict__ err) {
  const int lane = threadIdx.x & 31, row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= n) return;
  const DpModelImage& M = *model;
  float pa[3], pb[3];
  pose_positions(M, pose_a + (size_t)row * (DP_J * 4), lane, pa);
  pose_positions(M, pose_b + (size_t)row * (DP_J * 4), lane, pb);
  const float dx = pa[0] - pb[0], dy = pa[1] - pb[1], dz = pa[2] - pb[2];
  const float d = lane < DP_J ? sqrtf(dx * dx + dy * dy + dz * dz) : 0.0f;
  const bool ee = lane == 4 || lane == 8 || lane == 13 || lane == 17 || lane == 21;
  const float all = warp_sum(d), sparse = warp_sum(ee ? d : 0.0f);
  if (lane == 0) {
    err[row * 2] = all * (1.0f / DP_J);
    err[row * 2 + 1] = sparse * (1.0f / 5.0f);
  }
}

}  // namespace

cudaError_t dp_pose_error_launch(const DpModelImage* model, const float* pose_a, const float* pose_b, int n, float* err, cudaStream_t st) {
  dp_pose_error_kernel<<<(n + 3) / 4, 128, 0, st>>>(model, pose_a, pose_b, n, err);
  return cudaGetLastError();
}
