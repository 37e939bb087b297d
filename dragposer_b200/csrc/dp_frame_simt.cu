// dp_frame_simt.cu -- persistent per-frame optimisation kernel, fp32 CUDA-core decoder.
//
// One launch runs a whole frame of DragPose.run (python/src/drag_pose.py:196-414) for
// every clip: up to max_iter x { decoder forward, forward kinematics, masked tracker
// loss, reverse-mode adjoint, decoder backward, latent Adam step } with per-clip early
// stopping, then the frame epilogue (root update, joint adjustment, ring-buffer push,
// output pose).  No host round trips inside the frame.
//
// Mapping: one warp owns CPW clips.  Decoder phases process the CPW clips together
// (each weight fetched from shared memory feeds CPW FMAs); kinematics phases run one
// clip at a time with lane == joint: the joint hierarchy lives in shared memory
// (DpModelImage), positions accumulate along the parent chain with log-step warp
// shuffles (ancestor pointer jumping), subtree sums of the adjoint use a warp scan over
// the pre-order joint numbering.  Arithmetic follows SURVEY.md appendix B (closed form of
// python/src/utils.py:80-149 + python/src/drag_pose.py:66-194).
#include "dp_common.cuh"
#include "dp_internal.h"

namespace {

struct ClipTrackers {   // per-clip, per-lane(joint) tracker row in shared memory
  float4 pw;            // tp.xyz, w_pos
  float4 r0;            // TR row 0, w_rot
  float4 r1;            // TR row 1, -
  float4 r2;            // TR row 2, -
};

template <int CPW, int K, int N>
__device__ __forceinline__ void dense_pairs(const float* __restrict__ W, const float* __restrict__ bias,
                                            const float* __restrict__ in, float (&acc)[CPW][3], int lane) {
  // out[o] = bias[o] + sum_k W[k][o] in[k];  lane owns o = 2*lane, 2*lane+1 (< min(N,64)) and 64+lane (< N)
  constexpr bool THIRD = (N > 64);
  const bool pa = (2 * lane < N) && (lane < 32);
  const int o = pa ? 2 * lane : 0;
  const bool pt = THIRD && (64 + lane < N);
  const int o3 = pt ? 64 + lane : 0;
#pragma unroll
  for (int c = 0; c < CPW; ++c) {
    acc[c][0] = bias ? bias[o] : 0.0f;
    acc[c][1] = bias ? bias[o + 1] : 0.0f;
    acc[c][2] = (THIRD && bias) ? bias[o3] : 0.0f;
  }
#pragma unroll 2
  for (int k = 0; k < K; k += 4) {
    float4 a[CPW];
#pragma unroll
    for (int c = 0; c < CPW; ++c) a[c] = *reinterpret_cast<const float4*>(in + c * 2 * DP_SCRATCH + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float2 w = *reinterpret_cast<const float2*>(W + (k + kk) * N + o);
      float w3 = 0.0f;
      if (THIRD) w3 = W[(k + kk) * N + o3];
#pragma unroll
      for (int c = 0; c < CPW; ++c) {
        const float av = kk == 0 ? a[c].x : kk == 1 ? a[c].y : kk == 2 ? a[c].z : a[c].w;
        acc[c][0] = fmaf(w.x, av, acc[c][0]);
        acc[c][1] = fmaf(w.y, av, acc[c][1]);
        if (THIRD) acc[c][2] = fmaf(w3, av, acc[c][2]);
      }
    }
  }
}

__device__ __forceinline__ float lrelu(float x) { return x > 0.0f ? x : 0.2f * x; }

struct FkOut {
  float lp, lr;        // weighted position loss, lambda-scaled rotation loss (warp-uniform)
};

// Forward kinematics + masked tracker loss (+ adjoint) for ONE clip; lane == joint.
// y / ybar alias the same 96-float shared buffer (all reads of y precede the writes).
template <bool ADJOINT, bool EPILOGUE>
__device__ __forceinline__ FkOut fk_loss(const DpModelImage& M, float* __restrict__ ybuf, const ClipTrackers* __restrict__ trk,
                                         const float g[4], float inv3e, float lrot9e, int lane,
                                         // epilogue outputs
                                         float q_out[4], float r_out[4], float p_out[3], float d_out[3]) {
  const bool is_joint = lane < DP_J;
  const bool is_root = lane == 0;
  const float4 yv = is_joint ? reinterpret_cast<const float4*>(ybuf)[lane] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 yd = reinterpret_cast<const float4*>(ybuf)[DP_J];
  const float4 mq = is_joint ? reinterpret_cast<const float4*>(M.mean_q)[lane] : make_float4(1.f, 0.f, 0.f, 0.f);
  const float4 sq = is_joint ? reinterpret_cast<const float4*>(M.std_q)[lane] : make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
  float u[4] = {fmaf(yv.x, sq.x, mq.x), fmaf(yv.y, sq.y, mq.y), fmaf(yv.z, sq.z, mq.z), fmaf(yv.w, sq.w, mq.w)};
  const float n = sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2] + u[3] * u[3]);
  const float inv = 1.0f / (n + 1e-8f);
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
#pragma unroll
  for (int rd = 0; rd < DP_JUMP_ROUNDS; ++rd) {
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
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sp += __shfl_xor_sync(0xffffffffu, sp, o);
    sr += __shfl_xor_sync(0xffffffffu, sr, o);
  }
  FkOut out;
  out.lp = sp * inv3e;
  out.lr = sr * lrot9e;
  if (EPILOGUE) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { q_out[i] = q[i]; r_out[i] = r[i]; }
#pragma unroll
    for (int i = 0; i < 3; ++i) { p_out[i] = p[i]; d_out[i] = d[i]; }
  }
  if (ADJOINT) {
    // seeds
    const float kp = 2.0f * wp * inv3e, kr = 2.0f * wr * lrot9e;
    float pb[3] = {ep[0] * kp, ep[1] * kp, ep[2] * kp};
    float Rb[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) Rb[i] = eR[i] * kr;
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
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) rbp[i] += __shfl_xor_sync(0xffffffffu, rbp[i], o);
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
    const float kk = dt * inv * inv / n;
    float4 yb;
    yb.x = (qb[0] * inv - u[0] * kk) * sq.x;
    yb.y = (qb[1] * inv - u[1] * kk) * sq.y;
    yb.z = (qb[2] * inv - u[2] * kk) * sq.z;
    yb.w = (qb[3] * inv - u[3] * kk) * sq.w;
    if (is_joint) reinterpret_cast<float4*>(ybuf)[lane] = yb;
    if (is_root) {
      float db[3];
      mat_t_vec(R0, cb, db);
      reinterpret_cast<float4*>(ybuf)[DP_J] = make_float4(db[0] * M.std_d[0], db[1] * M.std_d[1], db[2] * M.std_d[2], 0.0f);
    }
    __syncwarp();
  }
  return out;
}

template <int CPW>
__global__ void __launch_bounds__(512, 1) dp_frame_simt_kernel(const __grid_constant__ DpFrameArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DpModelImage& M = *reinterpret_cast<DpModelImage*>(smem_raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + sizeof(DpModelImage));
  float* scratch_base = reinterpret_cast<float*>(smem_raw + sizeof(DpModelImage) + 16);
  const int warps = blockDim.x >> 5;
  ClipTrackers* trk_base = reinterpret_cast<ClipTrackers*>(scratch_base + warps * CPW * 2 * DP_SCRATCH);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---- model image -> shared memory: one bulk TMA copy per CTA
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    constexpr uint32_t kBytes = (uint32_t)sizeof(DpModelImage);
    constexpr uint32_t kChunk = 32768;
    mbar_expect_tx(bar, kBytes);
    for (uint32_t o = 0; o < kBytes; o += kChunk)
      tma_bulk_g2s(smem_raw + o, reinterpret_cast<const unsigned char*>(A.model) + o, min(kChunk, kBytes - o), bar);
  }

  float* sa = scratch_base + (warp * CPW) * 2 * DP_SCRATCH;  // clip c: ping = sa + c*2*S, pong = +S
  float* sb = sa + DP_SCRATCH;
  ClipTrackers* trk = trk_base + (warp * CPW) * 32;
  const int clip0 = (blockIdx.x * warps + warp) * CPW;

  // ---- per-clip frame inputs
  bool valid[CPW];
  float g[CPW][4], inv3e[CPW], lrot9e[CPW];
  float2 z[CPW], tl[CPW], am[CPW], av[CPW], zlast[CPW];
#pragma unroll
  for (int c = 0; c < CPW; ++c) {
    const int clip = clip0 + c;
    valid[c] = clip < A.n_clips;
    const int cc = valid[c] ? clip : 0;
    const int ne = A.n_ee ? A.n_ee[cc] : A.ee_stride;
    inv3e[c] = 1.0f / (3.0f * (float)ne);
    lrot9e[c] = A.lambda_rot / (9.0f * (float)ne);
#pragma unroll
    for (int i = 0; i < 4; ++i) g[c][i] = A.grot[cc * 4 + i];
    z[c] = tl[c] = make_float2(0.f, 0.f);
    if (lane < DP_L / 2) {
      z[c] = reinterpret_cast<const float2*>(A.latent + (size_t)cc * DP_L)[lane];
      tl[c] = reinterpret_cast<const float2*>(A.target_buf + ((size_t)cc * A.target_rows + A.target_index) * DP_L)[lane];
    }
    am[c] = av[c] = make_float2(0.f, 0.f);
    zlast[c] = z[c];
    // tracker rows: lane j picks the slot that tracks joint j (joints are unique per clip)
    ClipTrackers row;
    row.pw = row.r0 = row.r1 = row.r2 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int32_t* jn = A.joints + (A.shared_trackers ? 0 : (size_t)cc * A.ee_stride);
    const float* wt = A.weights + (A.shared_trackers ? 0 : (size_t)cc * A.ee_stride * 2);
    for (int e = 0; e < ne; ++e) {
      if (jn[e] == lane) {
        const float* tp = A.tgt_pos + ((size_t)cc * A.ee_stride + e) * 3;
        const float* tr = A.tgt_rot + ((size_t)cc * A.ee_stride + e) * 9;
        row.pw = make_float4(tp[0], tp[1], tp[2], wt[2 * e]);
        row.r0 = make_float4(tr[0], tr[1], tr[2], wt[2 * e + 1]);
        row.r1 = make_float4(tr[3], tr[4], tr[5], 0.f);
        row.r2 = make_float4(tr[6], tr[7], tr[8], 0.f);
      }
    }
    trk[c * 32 + lane] = row;
  }
  mbar_wait(bar, 0);  // model image has landed
  __syncwarp();

  // ---- optimisation loop (drag_pose.py:296-355)
  bool active[CPW];
  double prev[CPW], incr[CPW];
  float lp[CPW], lr[CPW], lt[CPW];
  int iters[CPW];
#pragma unroll
  for (int c = 0; c < CPW; ++c) {
    active[c] = valid[c];
    prev[c] = 10000000.0;
    incr[c] = 1.0;
    lp[c] = lr[c] = lt[c] = __int_as_float(0x7f800000);
    iters[c] = 0;
  }
  const float lt_scale = A.lambda_t * (1.0f / (float)DP_L);
  for (int it = 0; it < A.max_iter; ++it) {
    bool any = false;
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
      active[c] = active[c] && ((double)lp[c] > A.eps_pos || (double)lr[c] > A.eps_rot) && (incr[c] > A.min_incr);
      any = any || active[c];
    }
    if (!any) break;
    // decoder forward 24 -> 40 -> 60 -> 92
    float acc[CPW][3];
    unsigned neg0[CPW], neg1[CPW];  // LeakyReLU slope bits of the lane's pair
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
      if (active[c]) zlast[c] = z[c];
      if (lane < DP_L / 2) reinterpret_cast<float2*>(sa + c * 2 * DP_SCRATCH)[lane] = z[c];
    }
    __syncwarp();
    dense_pairs<CPW, DP_L, DP_H0>(M.W0t, M.b0, sa, acc, lane);
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
      neg0[c] = (acc[c][0] > 0.f ? 0u : 1u) | (acc[c][1] > 0.f ? 0u : 2u);
      if (2 * lane < DP_H0) reinterpret_cast<float2*>(sb + c * 2 * DP_SCRATCH)[lane] = make_float2(lrelu(acc[c][0]), lrelu(acc[c][1]));
    }
    __syncwarp();
    dense_pairs<CPW, DP_H0, DP_H1>(M.W1t, M.b1, sb, acc, lane);
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
      neg1[c] = (acc[c][0] > 0.f ? 0u : 1u) | (acc[c][1] > 0.f ? 0u : 2u);
      if (2 * lane < DP_H1) reinterpret_cast<float2*>(sa + c * 2 * DP_SCRATCH)[lane] = make_float2(lrelu(acc[c][0]), lrelu(acc[c][1]));
    }
    __syncwarp();
    dense_pairs<CPW, DP_H1, DP_Y>(M.W2t, M.b2, sa, acc, lane);
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
      float* yb = sb + c * 2 * DP_SCRATCH;
      reinterpret_cast<float2*>(yb)[lane] = make_float2(acc[c][0], acc[c][1]);
      if (64 + lane < DP_Y) yb[64 + lane] = acc[c][2];
    }
    __syncwarp();
    // kinematics + loss + adjoint, one clip at a time (lane == joint); ybar overwrites y
    float nlp[CPW], nlr[CPW];
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
      nlp[c] = lp[c];
      nlr[c] = lr[c];
      if (active[c]) {
        const FkOut o = fk_loss<true, false>(M, sb + c * 2 * DP_SCRATCH, trk + c * 32, g[c], inv3e[c], lrot9e[c], lane,
                                             nullptr, nullptr, nullptr, nullptr);
        nlp[c] = o.lp;
        nlr[c] = o.lr;
      }
    }
    __syncwarp();
    // decoder backward 92 -> 60 -> 40 -> 24 (data gradient only; the decoder is frozen)
    dense_pairs<CPW, DP_Y, DP_H1>(M.W2, nullptr, sb, acc, lane);
#pragma unroll
    for (int c = 0; c < CPW; ++c)
      if (2 * lane < DP_H1)
        reinterpret_cast<float2*>(sa + c * 2 * DP_SCRATCH)[lane] =
            make_float2(acc[c][0] * ((neg1[c] & 1u) ? 0.2f : 1.0f), acc[c][1] * ((neg1[c] & 2u) ? 0.2f : 1.0f));
    __syncwarp();
    dense_pairs<CPW, DP_H1, DP_H0>(M.W1, nullptr, sa, acc, lane);
#pragma unroll
    for (int c = 0; c < CPW; ++c)
      if (2 * lane < DP_H0)
        reinterpret_cast<float2*>(sb + c * 2 * DP_SCRATCH)[lane] =
            make_float2(acc[c][0] * ((neg0[c] & 1u) ? 0.2f : 1.0f), acc[c][1] * ((neg0[c] & 2u) ? 0.2f : 1.0f));
    __syncwarp();
    dense_pairs<CPW, DP_H0, DP_L>(M.W0, nullptr, sb, acc, lane);
    // temporal term, loss bookkeeping, Adam (torch/optim/adam.py single-tensor path)
    const float step_size = A.adam_tab[it], bc2s = A.adam_tab[A.max_iter + it];
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
      const float dx = z[c].x - tl[c].x, dy = z[c].y - tl[c].y;
      float s = (lane < DP_L / 2) ? fmaf(dx, dx, dy * dy) : 0.0f;
      s = warp_sum(s);
      const float nlt = s * lt_scale;
      const float gx = fmaf(2.0f * lt_scale, dx, acc[c][0]);
      const float gy = fmaf(2.0f * lt_scale, dy, acc[c][1]);
      if (A.trace && active[c]) {
        float* row = A.trace + ((size_t)(clip0 + c) * A.trace_iters + it) * 52;
        if (lane < DP_L / 2) {
          reinterpret_cast<float2*>(row)[lane] = z[c];
          reinterpret_cast<float2*>(row + DP_L)[lane] = make_float2(gx, gy);
        }
        if (lane == 0) { row[48] = nlp[c]; row[49] = nlr[c]; row[50] = nlt; row[51] = 1.0f; }
      }
      if (A.eval_only) {
        if (active[c] && lane < DP_L / 2) reinterpret_cast<float2*>(A.eval_grad + (size_t)(clip0 + c) * DP_L)[lane] = make_float2(gx, gy);
      } else if (active[c]) {
        am[c].x = fmaf(0.1f, gx - am[c].x, am[c].x);
        am[c].y = fmaf(0.1f, gy - am[c].y, am[c].y);
        av[c].x = av[c].x * 0.999f + (0.001f * gx) * gx;
        av[c].y = av[c].y * 0.999f + (0.001f * gy) * gy;
        z[c].x += (-step_size * am[c].x) / (sqrtf(av[c].x) / bc2s + 1e-8f);
        z[c].y += (-step_size * am[c].y) / (sqrtf(av[c].y) / bc2s + 1e-8f);
      }
      if (active[c]) {
        lp[c] = nlp[c];
        lr[c] = nlr[c];
        lt[c] = nlt;
        const float total = (nlp[c] + nlr[c]) + nlt;
        incr[c] = prev[c] - (double)total;
        prev[c] = (double)total;
        iters[c] += 1;
      }
    }
  }

  // ---- frame epilogue (drag_pose.py:369-414) from the LAST EVALUATED latent (pre-step)
#pragma unroll
  for (int c = 0; c < CPW; ++c)
    if (lane < DP_L / 2) reinterpret_cast<float2*>(sa + c * 2 * DP_SCRATCH)[lane] = zlast[c];
  __syncwarp();
  {
    float acc[CPW][3];
    dense_pairs<CPW, DP_L, DP_H0>(M.W0t, M.b0, sa, acc, lane);
#pragma unroll
    for (int c = 0; c < CPW; ++c)
      if (2 * lane < DP_H0) reinterpret_cast<float2*>(sb + c * 2 * DP_SCRATCH)[lane] = make_float2(lrelu(acc[c][0]), lrelu(acc[c][1]));
    __syncwarp();
    dense_pairs<CPW, DP_H0, DP_H1>(M.W1t, M.b1, sb, acc, lane);
#pragma unroll
    for (int c = 0; c < CPW; ++c)
      if (2 * lane < DP_H1) reinterpret_cast<float2*>(sa + c * 2 * DP_SCRATCH)[lane] = make_float2(lrelu(acc[c][0]), lrelu(acc[c][1]));
    __syncwarp();
    dense_pairs<CPW, DP_H1, DP_Y>(M.W2t, M.b2, sa, acc, lane);
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
      float* yb = sb + c * 2 * DP_SCRATCH;
      reinterpret_cast<float2*>(yb)[lane] = make_float2(acc[c][0], acc[c][1]);
      if (64 + lane < DP_Y) yb[64 + lane] = acc[c][2];
    }
    __syncwarp();
  }
#pragma unroll
  for (int c = 0; c < CPW; ++c) {
    if (!valid[c]) continue;
    const int clip = clip0 + c;
    float q[4], r[4], p[3], d[3];
    fk_loss<false, true>(M, sb + c * 2 * DP_SCRATCH, trk + c * 32, g[c], inv3e[c], lrot9e[c], lane, q, r, p, d);
    if (A.eval_only) {
      if (lane < DP_J && A.eval_pos) {
        float* o = A.eval_pos + ((size_t)clip * DP_J + lane) * 3;
        o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
      }
      if (lane == 0 && A.out_losses) { A.out_losses[clip * 3] = lp[c]; A.out_losses[clip * 3 + 1] = lr[c]; A.out_losses[clip * 3 + 2] = lt[c]; }
      continue;
    }
    // root update: current_global_pos += world_displacement (= p_0), current_global_rot = world_rotation
    float p0[3], gp[3], adj[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      p0[i] = __shfl_sync(0xffffffffu, p[i], 0);
      gp[i] = A.gpos[clip * 3 + i] + p0[i];
    }
    if (A.adj_joint >= 0) {  // joint adjustment (drag_pose.py:374-381)
      const float* tp = A.tgt_pos + ((size_t)clip * A.ee_stride + A.adj_slot) * 3;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float pj = __shfl_sync(0xffffffffu, p[i], A.adj_joint);
        adj[i] = (tp[i] - pj) * A.adj_w;
        gp[i] += adj[i];
      }
    }
    __syncwarp();
    const int hs = M.height_slot[lane];
    if (hs >= 0) A.height_buf[((size_t)clip * DP_PAST + A.ring_head) * DP_NH + hs] = p[1] + gp[1];
    if (lane < DP_L / 2) {
      reinterpret_cast<float2*>(A.latent_buf + ((size_t)clip * DP_PAST + A.ring_head) * DP_L)[lane] = zlast[c];
      reinterpret_cast<float2*>(A.latent + (size_t)clip * DP_L)[lane] = z[c];
    }
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        A.disp_buf[((size_t)clip * DP_PAST + A.ring_head) * 3 + i] = d[i] + adj[i];
        A.gpos[clip * 3 + i] = gp[i];
        A.out_gpos[clip * 3 + i] = gp[i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) A.grot[clip * 4 + i] = r[i];
      A.out_iters[clip] = iters[c];
      A.out_losses[clip * 3] = lp[c];
      A.out_losses[clip * 3 + 1] = lr[c];
      A.out_losses[clip * 3 + 2] = lt[c];
    }
    // output pose: standardised root-space quats, root slot := standardised world rotation
    if (lane < DP_J) {
      const float4 mq = reinterpret_cast<const float4*>(M.mean_q)[lane];
      const float4 sq = reinterpret_cast<const float4*>(M.std_q)[lane];
      const float* s = (lane == 0) ? r : q;
      reinterpret_cast<float4*>(A.out_pose + (size_t)clip * 88)[lane] =
          make_float4((s[0] - mq.x) / sq.x, (s[1] - mq.y) / sq.y, (s[2] - mq.z) / sq.z, (s[3] - mq.w) / sq.w);
    }
  }
}

}  // namespace

size_t dp_frame_simt_smem_bytes(int warps, int cpw) {
  return sizeof(DpModelImage) + 16 + (size_t)warps * cpw * (2 * DP_SCRATCH * sizeof(float) + 32 * sizeof(ClipTrackers));
}

cudaError_t dp_frame_simt_launch(const DpFrameArgs& args, int num_sms, cudaStream_t stream) {
  // 4096 clips on 148 SMs: two clips per warp, ~14 warps per CTA, one CTA per SM (single wave).
  const int cpw = args.n_clips > 2 * num_sms ? 2 : 1;
  const int total_warps = (args.n_clips + cpw - 1) / cpw;
  int wpc = (total_warps + num_sms - 1) / num_sms;
  wpc = wpc < 1 ? 1 : (wpc > 16 ? 16 : wpc);
  const int grid = (total_warps + wpc - 1) / wpc;
  const size_t smem = dp_frame_simt_smem_bytes(wpc, cpw);
  cudaError_t err;
  if (cpw == 2) {
    err = cudaFuncSetAttribute(dp_frame_simt_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    dp_frame_simt_kernel<2><<<grid, wpc * 32, smem, stream>>>(args);
  } else {
    err = cudaFuncSetAttribute(dp_frame_simt_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    dp_frame_simt_kernel<1><<<grid, wpc * 32, smem, stream>>>(args);
  }
  return cudaGetLastError();
}
