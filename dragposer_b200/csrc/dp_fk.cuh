// dp_fk.cuh -- forward kinematics + masked tracker loss + reverse-mode adjoint for ONE clip per warp
// (lane == joint), shared by the fp32 CUDA-core frame kernel and the tcgen05 frame kernel.
// Arithmetic: SURVEY.md appendix B, i.e. the closed form of python/src/utils.py:80-149 and
// python/src/drag_pose.py:66-194.  `M` only needs the statistics / skeleton-table members.
#pragma once
#include "dp_common.cuh"

struct ClipTrackers {   // per-clip, per-lane(joint) tracker row in shared memory
  float4 pw;            // tp.xyz, w_pos
  float4 r0;            // TR row 0, w_rot
  float4 r1;            // TR row 1, -
  float4 r2;            // TR row 2, -
};

__device__ __forceinline__ float lrelu(float x) { return x > 0.0f ? x : 0.2f * x; }

struct FkOut {
  float lp, lr;        // weighted position loss, lambda-scaled rotation loss (warp-uniform)
  float le;            // sum of the enabled extension losses (0 when none)
};

// The reference's "Additional Losses" (python/src/drag_pose.py:129-183, commented out as shipped: the documented extension point for
// constraints expressed as losses; y axis = 1, joints of the 22-joint body).  mask bits: 1 feet on the floor (toes 4, 8),
// 2 head / hips facing the same way, 4 head above the hips, 8 hips between the feet (ankles 3, 7).  mask == 0 costs nothing.
struct FkExtra {
  int mask;
  float floor_level;
  float gpos_y;        // y of the clip's current global position (the positions here are relative to the previous root)
};

// default sink of the adjoint: dL/dy overwrites y in place (all reads of y precede the writes)
struct FkEmitInPlace {
  float* ybuf;
  int lane;
  __device__ __forceinline__ void operator()(const float4& yb, const float4& db) const {
    if (lane < DP_J) reinterpret_cast<float4*>(ybuf)[lane] = yb;
    if (lane == 0) reinterpret_cast<float4*>(ybuf)[DP_J] = db;
    __syncwarp();
  }
};

// Forward kinematics + masked tracker loss (+ adjoint) for ONE clip; lane == joint.
// The adjoint hands dL/dy to `emit(yb, db)`, called convergently by all 32 lanes: yb = dL/dy[4 lane .. 4 lane + 3] (zeros on
// lanes >= 22), db = dL/d(displacement outputs y[88..91]) on lane 0 (zeros elsewhere).
template <bool ADJOINT, bool EPILOGUE, class MODEL, class EMIT>
__device__ __forceinline__ FkOut fk_loss(const MODEL& M, const float* __restrict__ ybuf, const ClipTrackers* __restrict__ trk,
                                         const float g[4], float inv3e, float lrot9e, int lane,
                                         // epilogue outputs
                                         float q_out[4], float r_out[4], float p_out[3], float d_out[3], EMIT emit,
                                         const FkExtra X = FkExtra{0, 0.0f, 0.0f}) {
  const bool is_joint = lane < DP_J;
  const bool is_root = lane == 0;
  const float4 yv = is_joint ? reinterpret_cast<const float4*>(ybuf)[lane] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 yd = reinterpret_cast<const float4*>(ybuf)[DP_J];
  const float4 mq = is_joint ? reinterpret_cast<const float4*>(M.mean_q)[lane] : make_float4(1.f, 0.f, 0.f, 0.f);
  const float4 sq = is_joint ? reinterpret_cast<const float4*>(M.std_q)[lane] : make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
  float u[4] = {fmaf(yv.x, sq.x, mq.x), fmaf(yv.y, sq.y, mq.y), fmaf(yv.z, sq.z, mq.z), fmaf(yv.w, sq.w, mq.w)};
  const float n = fast_sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2] + u[3] * u[3]);
  const float inv = fast_rcp(n + 1e-8f);
  float q[4] = {u[0] * inv, u[1] * inv, u[2] * inv, u[3] * inv};
  float q0[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) q0[i] = __shfl_sync(0xffffffffu, q[i], 0);
  float r[4];
  quat_mul(g, q0, r);  // world root rotation (drag_pose.py:88-92)
  float R0[9], Mj[9], R[9];
  quat_to_mat(r, R0);
  {
    const float ident[4] = {1.f, 0.f, 0.f, 0.f};
    quat_to_mat(is_root ? ident : q, Mj);
  }
  mat_mul(R0, Mj, R);  // closed form of utils.py:80-149: R_j = R_0 M(q_j)
  const float d[3] = {fmaf(yd.x, M.std_d[0], M.mean_d[0]), fmaf(yd.y, M.std_d[1], M.mean_d[1]),
                      fmaf(yd.z, M.std_d[2], M.mean_d[2])};
  float p0[3];
  mat_vec(R0, d, p0);  // == quat.mul_vec(world_rotation, displacement) (drag_pose.py:102)
  // c_j = R_parent o_j ; p_j = sum of c over the ancestor chain (log-step pointer jumping)
  const int par = M.parent[lane];
  float Rp[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) Rp[i] = __shfl_sync(0xffffffffu, R[i], par);
  const float4 off = *reinterpret_cast<const float4*>(M.off[lane]);
  const float ov[3] = {off.x, off.y, off.z};
  float p[3];
  mat_vec(Rp, ov, p);
  if (is_root) { p[0] = p0[0]; p[1] = p0[1]; p[2] = p0[2]; }
  const int n_jump = M.pad[0], n_child = M.pad[1];  // rounds this skeleton needs (3 and 3 for the 22-joint body)
#pragma unroll
  for (int rd = 0; rd < DP_JUMP_ROUNDS; ++rd) {
    if (rd >= n_jump) break;
    const int a = M.jump[rd][lane];
    const int src = a >= 0 ? a : lane;
    const float t0 = __shfl_sync(0xffffffffu, p[0], src);
    const float t1 = __shfl_sync(0xffffffffu, p[1], src);
    const float t2 = __shfl_sync(0xffffffffu, p[2], src);
    if (a >= 0) { p[0] += t0; p[1] += t1; p[2] += t2; }
  }
  // masked tracker loss (drag_pose.py:116-124); untracked lanes carry zero weights
  const ClipTrackers tk = trk[lane];
  const float ep[3] = {p[0] - tk.pw.x, p[1] - tk.pw.y, p[2] - tk.pw.z};
  const float wp = tk.pw.w, wr = tk.r0.w;
  float eR[9] = {R[0] - tk.r0.x, R[1] - tk.r0.y, R[2] - tk.r0.z, R[3] - tk.r1.x, R[4] - tk.r1.y,
                 R[5] - tk.r1.z, R[6] - tk.r2.x, R[7] - tk.r2.y, R[8] - tk.r2.z};
  float sp = ep[0] * ep[0] + ep[1] * ep[1] + ep[2] * ep[2];
  float sr = 0.0f;
#pragma unroll
  for (int i = 0; i < 9; ++i) sr = fmaf(eR[i], eR[i], sr);
  sp *= wp;
  sr *= wr;
  {  // two warp sums with 6 shuffles: lanes 0-15 finish the position sum, lanes 16-31 the rotation sum, then swap
    const bool hi = lane & 16;
    float k = (hi ? sr : sp) + __shfl_xor_sync(0xffffffffu, hi ? sp : sr, 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) k += __shfl_xor_sync(0xffffffffu, k, o);
    const float other = __shfl_xor_sync(0xffffffffu, k, 16);
    sp = hi ? other : k;
    sr = hi ? k : other;
  }
  FkOut out;
  out.lp = sp * inv3e;
  out.lr = sr * lrot9e;
  out.le = 0.0f;
  // extension losses: value, and seeds of the adjoint (pe -> pbar of this lane, re02 / re22 -> Rbar[0][2], Rbar[2][2])
  float pe[3] = {0.f, 0.f, 0.f}, re02 = 0.f, re22 = 0.f;
  if (X.mask) {  // warp-uniform
    float le = 0.0f;
    const float p0x = __shfl_sync(0xffffffffu, p[0], 0), p0z = __shfl_sync(0xffffffffu, p[2], 0);
    if ((X.mask & 1) && (lane == 4 || lane == 8)) {  // mean over the two toes of (global y + p_y - floor)^2  (:133-135)
      const float e = X.gpos_y + (p[1] - X.floor_level);
      le = fmaf(0.5f * e, e, le);
      pe[1] = e;
    }
    float sx = 0.f, sz = 0.f;  // ground-plane seeds that the hips receive with the opposite sign
    if ((X.mask & 4) && lane == 13) {  // |head - hips|^2 on the ground plane (:159-164)
      const float dx = p[0] - p0x, dz = p[2] - p0z;
      le += dx * dx + dz * dz;
      sx = 2.0f * dx; sz = 2.0f * dz;
    }
    if ((X.mask & 8) && (lane == 3 || lane == 7)) {  // max(|hips - ankle|^2 - 0.2^2, 0) on the ground plane (:166-176)
      const float dx = p[0] - p0x, dz = p[2] - p0z;
      const float d2 = dx * dx + dz * dz - 0.2f * 0.2f;
      if (d2 > 0.0f) { le += d2; sx = 2.0f * dx; sz = 2.0f * dz; }
    }
    pe[0] = sx; pe[2] = sz;
    {
      const float tx = __shfl_sync(0xffffffffu, sx, 13) + __shfl_sync(0xffffffffu, sx, 3) + __shfl_sync(0xffffffffu, sx, 7);
      const float tz = __shfl_sync(0xffffffffu, sz, 13) + __shfl_sync(0xffffffffu, sz, 3) + __shfl_sync(0xffffffffu, sz, 7);
      if (is_root) { pe[0] -= tx; pe[2] -= tz; }
    }
    if (X.mask & 2) {  // (1 - min(1, fwd_head . fwd_hips + 0.2))^2 with the forward axes (third columns) flattened to the ground (:137-157)
      const float hx = __shfl_sync(0xffffffffu, R[2], 13), hz = __shfl_sync(0xffffffffu, R[8], 13);
      const float gx = R0[2], gz = R0[8];
      const float nh = sqrtf(hx * hx + hz * hz), ng = sqrtf(gx * gx + gz * gz);
      if (nh > 0.5f) {
        const float ihx = hx / nh, ihz = hz / nh, igx = gx / ng, igz = gz / ng;
        const float s = ihx * igx + ihz * igz, c = s + 0.2f;
        if (c < 1.0f) {
          const float dl = -2.0f * (1.0f - c);  // dL/ds
          if (lane == 13) { re02 = dl * (igx - s * ihx) / nh; re22 = dl * (igz - s * ihz) / nh; }
          if (is_root) { re02 = dl * (ihx - s * igx) / ng; re22 = dl * (ihz - s * igz) / ng; le += (1.0f - c) * (1.0f - c); }
        }
      }
    }
    out.le = warp_sum(le);
  }
  if (EPILOGUE) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { q_out[i] = q[i]; r_out[i] = r[i]; }
#pragma unroll
    for (int i = 0; i < 3; ++i) { p_out[i] = p[i]; d_out[i] = d[i]; }
  }
  if (ADJOINT) {
    // seeds
    const float kp = 2.0f * wp * inv3e, kr = 2.0f * wr * lrot9e;
    float pb[3] = {fmaf(ep[0], kp, pe[0]), fmaf(ep[1], kp, pe[1]), fmaf(ep[2], kp, pe[2])};
    float Rb[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) Rb[i] = eR[i] * kr;
    Rb[2] += re02;
    Rb[8] += re22;
    // subtree sums of pbar over the pre-order numbering: inclusive scan, then a range difference
    float P[3] = {pb[0], pb[1], pb[2]};
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t0 = __shfl_up_sync(0xffffffffu, P[0], o);
      const float t1 = __shfl_up_sync(0xffffffffu, P[1], o);
      const float t2 = __shfl_up_sync(0xffffffffu, P[2], o);
      if (lane >= o) { P[0] += t0; P[1] += t1; P[2] += t2; }
    }
    const int last = M.last[lane];
    float cb[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float hi = __shfl_sync(0xffffffffu, P[i], last);
      float lo = __shfl_up_sync(0xffffffffu, P[i], 1);
      if (lane == 0) lo = 0.0f;
      cb[i] = hi - lo;  // cbar_j = sum of pbar over subtree(j)
    }
    // Rbar_j += sum_children cbar_c o_c^T   (p_c = p_j + R_j o_c)
#pragma unroll
    for (int k = 0; k < DP_MAX_CHILD; ++k) {
      if (k >= n_child) break;
      const int ch = M.child[k][lane];
      const int src = ch >= 0 ? ch : lane;
      const float t0 = __shfl_sync(0xffffffffu, cb[0], src);
      const float t1 = __shfl_sync(0xffffffffu, cb[1], src);
      const float t2 = __shfl_sync(0xffffffffu, cb[2], src);
      if (ch >= 0) {
        const float4 co = *reinterpret_cast<const float4*>(M.coff[k][lane]);
        Rb[0] = fmaf(t0, co.x, Rb[0]); Rb[1] = fmaf(t0, co.y, Rb[1]); Rb[2] = fmaf(t0, co.z, Rb[2]);
        Rb[3] = fmaf(t1, co.x, Rb[3]); Rb[4] = fmaf(t1, co.y, Rb[4]); Rb[5] = fmaf(t1, co.z, Rb[5]);
        Rb[6] = fmaf(t2, co.x, Rb[6]); Rb[7] = fmaf(t2, co.y, Rb[7]); Rb[8] = fmaf(t2, co.z, Rb[8]);
      }
    }
    if (is_root) {  // p_0 = R_0 d
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Rb[3 * i + j] = fmaf(cb[i], d[j], Rb[3 * i + j]);
    }
    // every joint contributes Rbar_j M_j^T to Rbar_0 (M_0 = I); reduce in quaternion space (4 values, not 9)
    float X[9], rbp[4];
    mat_mul_bt(Rb, Mj, X);
    mat_bar_to_quat(r, X, rbp);
    {  // only lane 0 needs the four totals: 6-shuffle scatter reduction, then lane 0 collects from lanes 0, 8, 16, 24
      const float k = warp_sum4_scatter(rbp, lane);
      rbp[0] = k;
      rbp[1] = __shfl_sync(0xffffffffu, k, 8);
      rbp[2] = __shfl_sync(0xffffffffu, k, 16);
      rbp[3] = __shfl_sync(0xffffffffu, k, 24);
    }
    float qb[4];
    if (is_root) {
      const float gc[4] = {g[0], -g[1], -g[2], -g[3]};
      quat_mul(gc, rbp, qb);  // r = g (x) q_0  ->  q0bar = conj(g) (x) rbar
    } else {
      float G[9];
      mat_mul_at(R0, Rb, G);
      mat_bar_to_quat(q, G, qb);
    }
    // adjoint of q = u / (|u| + 1e-8)
    const float dt = u[0] * qb[0] + u[1] * qb[1] + u[2] * qb[2] + u[3] * qb[3];
    const float kk = dt * inv * inv * fast_rcp(n);
    float4 yb;
    yb.x = (qb[0] * inv - u[0] * kk) * sq.x;
    yb.y = (qb[1] * inv - u[1] * kk) * sq.y;
    yb.z = (qb[2] * inv - u[2] * kk) * sq.z;
    yb.w = (qb[3] * inv - u[3] * kk) * sq.w;
    float4 db4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (is_root) {
      float db[3];
      mat_t_vec(R0, cb, db);
      db4 = make_float4(db[0] * M.std_d[0], db[1] * M.std_d[1], db[2] * M.std_d[2], 0.0f);
    }
    emit(yb, db4);
  }
  return out;
}
template <bool ADJOINT, bool EPILOGUE, class MODEL>
__device__ __forceinline__ FkOut fk_loss(const MODEL& M, float* __restrict__ ybuf, const ClipTrackers* __restrict__ trk,
                                         const float g[4], float inv3e, float lrot9e, int lane, float q_out[4], float r_out[4],
                                         float p_out[3], float d_out[3]) {
  return fk_loss<ADJOINT, EPILOGUE>(M, ybuf, trk, g, inv3e, lrot9e, lane, q_out, r_out, p_out, d_out, FkEmitInPlace{ybuf, lane});
}
template <bool ADJOINT, bool EPILOGUE, class MODEL>
__device__ __forceinline__ FkOut fk_loss(const MODEL& M, float* __restrict__ ybuf, const ClipTrackers* __restrict__ trk,
                                         const float g[4], float inv3e, float lrot9e, int lane, float q_out[4], float r_out[4],
                                         float p_out[3], float d_out[3], const FkExtra X) {
  return fk_loss<ADJOINT, EPILOGUE>(M, ybuf, trk, g, inv3e, lrot9e, lane, q_out, r_out, p_out, d_out, FkEmitInPlace{ybuf, lane}, X);
}

