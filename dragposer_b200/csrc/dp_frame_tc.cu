// dp_frame_tc.cu -- persistent per-frame optimisation kernel with the decoder on tcgen05 tensor cores.
//
// Same contract as dp_frame_simt.cu (one launch == one frame of DragPose.run, python/src/drag_pose.py:196-414,
// for every clip) but the six dense layers of an iteration (decoder forward 24->40->60->92 and its
// data-gradient 92->60->40->24, python/src/autoencoder.py:224-256 folded) run as tcgen05.mma tiles:
//
//   D[features(M=128) x clips(N=32)] = W[features x K] . X[K x clips]          accumulator in tensor memory
//
// * transposed mapping: the WEIGHTS are the 128-row A operand, the CTA's 32 clips are the N dimension, so
//   4096 clips occupy 128 SMs (clips-as-M would fill only 32).
// * precision: split products on kind::f16 (see PREC below): fp16x2 (default) or bf16x3, both ~fp32-exact
//   (gradient 2.5e-6 / 3.0e-6 relative vs 1.8e-6 for the fp32 kernel; plain TF32 would be 1.8e-3 and bf16x2
//   1e-4 -- both break the 1e-4 bar).  kind::f16 rather than kind::tf32 because tf32 produces zeros for
//   MN-major operands in the no-swizzle layout (measured) while f16/bf16 accept them: ONE weight image
//   serves the forward GEMM (A K-major) and the transposed backward GEMM (A MN-major), 42 KB instead of
//   155 KB of shared memory.
// * activations never leave the SM: epilogue warps read the accumulator with tcgen05.ld (lane == feature),
//   apply bias / LeakyReLU (slope bits stay in registers for the backward pass), split to bf16 pieces and
//   write the next layer's B operand (MN-major image, one STS.128 per 8 clips).
// * kinematics, loss, adjoint, Adam, early stopping and the frame epilogue are the warp-per-clip code shared
//   with the fp32 kernel (dp_fk.cuh).
#include "dp_fk.cuh"
#include "dp_internal.h"
#include "dp_umma.cuh"

namespace {

constexpr int NC = DP_TC_CLIPS;   // 32 clips per CTA == UMMA N
constexpr int kWarps = 16;
constexpr int kEpiWarps = 8;      // warps 0..7: TMEM lane quarter = warp % 4, clip half = warp / 4
constexpr int kIssueWarp = 8;
constexpr uint32_t kB_LBO = 128 * (NC / 8);  // activation image: K 8-groups 512 B apart, clip 8-groups 128 B apart
constexpr uint32_t kB_SBO = 128;
constexpr uint32_t kPingBytes = (64 / 8) * kB_LBO, kPongBytes = (96 / 8) * kB_LBO;  // per bf16 piece
constexpr int kPieces = 3;        // buffers are sized for three pieces
// PREC 0: bf16x3 -- x = x1 + x2 + x3 (bf16), six products (3,1)(2,2)(1,3)(2,1)(1,2)(1,1): 1.3e-7 relative per GEMM.
// PREC 1: fp16x2 -- x = x1 + x2 (fp16, 22 mantissa bits), three products (2,1)(1,2)(1,1): ~5e-7 relative per GEMM at half the
//         tensor and epilogue work.  fp16's narrow exponent is handled by exact power-of-two scalings: the weight image stores
//         16 W (undone in every epilogue), and the backward pass carries dL/dy scaled per clip so that its largest component
//         is in [16, 32) (undone when dL/dz is written); forward activations (|a| < 2^6 here) need no scaling.
constexpr float kWScale = 16.0f;

struct SmemTC {
  DpModelImageTC M;
  __align__(16) unsigned char ping[kPieces][kPingBytes];  // [piece]  z (24) / a1 (60) / dL/dh1 (60)
  __align__(16) unsigned char pong[kPieces][kPongBytes];  // [piece]  a0 (40) / dL/dy (92) / dL/dh0 (40)
  __align__(16) float ybuf[NC][96];                 // y, then dL/dy in place (fp32, one row per clip)
  float zgrad[NC][25];                              // dL/dz from the decoder (fp32)
  float bscale[NC];                                 // fp16 path: 1 / (per-clip power-of-two scale of dL/dy)
  __align__(16) ClipTrackers trk[NC][32];
  uint64_t bar_w, bar_mma;
  uint32_t tmem_base;
};

template <bool ACC>
__device__ __forceinline__ void umma_f16_c(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
  if (ACC)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                 "l"(a_desc), "l"(b_desc), "r"(idesc));
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                 "l"(a_desc), "l"(b_desc), "r"(idesc));
}

// layer geometry (DpModelImageTC): input width padded to 16, output rows padded to 16, both zero filled
template <int L> struct Lay;
template <> struct Lay<0> { static constexpr int kin = 32, kout = 48; static constexpr uint32_t off = DP_TC_W0_OFF; };
template <> struct Lay<1> { static constexpr int kin = 48, kout = 64; static constexpr uint32_t off = DP_TC_W1_OFF; };
template <> struct Lay<2> { static constexpr int kin = 64, kout = 96; static constexpr uint32_t off = DP_TC_W2_OFF; };

// the split products of one K step, smallest first
template <int PREC>
__device__ __forceinline__ void issue_terms(uint32_t tmem, const UmmaDescBase (&a)[kPieces], const UmmaDescBase (&b)[kPieces], uint32_t ao,
                                            uint32_t bo, uint32_t idesc, bool first) {
  if (PREC == 0) {  // (3,1) (2,2) (1,3) | (2,1) (1,2) | (1,1)
    if (first) umma_f16_c<false>(tmem, umma_desc_at(a[2], ao), umma_desc_at(b[0], bo), idesc);
    else umma_f16_c<true>(tmem, umma_desc_at(a[2], ao), umma_desc_at(b[0], bo), idesc);
    umma_f16_c<true>(tmem, umma_desc_at(a[1], ao), umma_desc_at(b[1], bo), idesc);
    umma_f16_c<true>(tmem, umma_desc_at(a[0], ao), umma_desc_at(b[2], bo), idesc);
    umma_f16_c<true>(tmem, umma_desc_at(a[1], ao), umma_desc_at(b[0], bo), idesc);
  } else {          // (2,1) | (1,2) | (1,1)
    if (first) umma_f16_c<false>(tmem, umma_desc_at(a[1], ao), umma_desc_at(b[0], bo), idesc);
    else umma_f16_c<true>(tmem, umma_desc_at(a[1], ao), umma_desc_at(b[0], bo), idesc);
  }
  umma_f16_c<true>(tmem, umma_desc_at(a[0], ao), umma_desc_at(b[1], bo), idesc);
  umma_f16_c<true>(tmem, umma_desc_at(a[0], ao), umma_desc_at(b[0], bo), idesc);
}

// forward layer L: A = W_L (K-major: LBO 128 between input 8-groups, SBO between output-row 8-groups)
template <int L, int PREC>
__device__ __forceinline__ void issue_fwd(const SmemTC& S, uint32_t tmem, const unsigned char* src, uint32_t piece_stride) {
  constexpr uint32_t sbo = 128 * (Lay<L>::kin / 8);
  constexpr uint32_t fmt = PREC ? 0u : 1u;  // kind::f16 operand format: 0 = fp16, 1 = bf16
  constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 16) | ((uint32_t)(NC >> 3) << 17) | (8u << 24);  // B MN-major
  UmmaDescBase a[kPieces], b[kPieces];
#pragma unroll
  for (int p = 0; p < kPieces; ++p) {
    a[p] = umma_desc_base(smem_u32(S.M.w[p]) + Lay<L>::off, 128, sbo);
    b[p] = umma_desc_base(smem_u32(src) + p * piece_stride, kB_LBO, kB_SBO);
  }
#pragma unroll
  for (int k = 0; k < Lay<L>::kin / 16; ++k) {
    const uint32_t ao = k * 256, bo = k * 2 * kB_LBO;
    issue_terms<PREC>(tmem, a, b, ao, bo, idesc, k == 0);
  }
}
// backward of layer L: the SAME weight image read MN-major (M = inputs, K = outputs): LBO / SBO swap roles
template <int L, int PREC>
__device__ __forceinline__ void issue_bwd(const SmemTC& S, uint32_t tmem, const unsigned char* src, uint32_t piece_stride) {
  constexpr uint32_t wsbo = 128 * (Lay<L>::kin / 8);
  constexpr uint32_t fmt = PREC ? 0u : 1u;
  constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NC >> 3) << 17) | (8u << 24);
  UmmaDescBase a[kPieces], b[kPieces];
#pragma unroll
  for (int p = 0; p < kPieces; ++p) {
    a[p] = umma_desc_base(smem_u32(S.M.w[p]) + Lay<L>::off, wsbo, 128);
    b[p] = umma_desc_base(smem_u32(src) + p * piece_stride, kB_LBO, kB_SBO);
  }
#pragma unroll
  for (int k = 0; k < Lay<L>::kout / 16; ++k) {
    const uint32_t ao = k * 2 * wsbo, bo = k * 2 * kB_LBO;
    issue_terms<PREC>(tmem, a, b, ao, bo, idesc, k == 0);
  }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {  // lower address <- lo_elem
  uint32_t p;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(hi_elem), "f"(lo_elem));
  return p;
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo_elem, float hi_elem) {
  uint32_t p;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(hi_elem), "f"(lo_elem));
  return p;
}
__device__ __forceinline__ void unpack_f16x2(uint32_t p, float& lo_elem, float& hi_elem) {
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}\n" : "=f"(lo_elem), "=f"(hi_elem) : "r"(p));
}
// 16 fp32 values (feature k, clips 16*half .. +15) -> split pieces in an MN-major activation image
template <int PREC>
__device__ __forceinline__ void store_pieces(unsigned char* img, uint32_t piece_stride, int k, int half, const float (&v)[16]) {
  unsigned char* dst = img + (k >> 3) * kB_LBO + (k & 7) * 16 + (2 * half) * kB_SBO;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    uint32_t p1[4], p2[4], p3[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float x0 = v[8 * g + 2 * i], x1 = v[8 * g + 2 * i + 1];
      if (PREC == 0) {
        p1[i] = pack_bf16x2(x0, x1);
        const float r0 = x0 - __uint_as_float(p1[i] << 16), r1 = x1 - __uint_as_float(p1[i] & 0xffff0000u);
        p2[i] = pack_bf16x2(r0, r1);
        const float s0 = r0 - __uint_as_float(p2[i] << 16), s1 = r1 - __uint_as_float(p2[i] & 0xffff0000u);
        p3[i] = pack_bf16x2(s0, s1);
      } else {
        p1[i] = pack_f16x2(x0, x1);
        float h0, h1;
        unpack_f16x2(p1[i], h0, h1);
        p2[i] = pack_f16x2(x0 - h0, x1 - h1);
      }
    }
    *reinterpret_cast<uint4*>(dst + g * kB_SBO) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
    *reinterpret_cast<uint4*>(dst + piece_stride + g * kB_SBO) = make_uint4(p2[0], p2[1], p2[2], p2[3]);
    if (PREC == 0) *reinterpret_cast<uint4*>(dst + 2 * piece_stride + g * kB_SBO) = make_uint4(p3[0], p3[1], p3[2], p3[3]);
  }
}
// a single fp32 value (feature k, clip n) -> its pieces (used by the Adam lanes for the latent)
template <int PREC>
__device__ __forceinline__ void store_piece_scalar(unsigned char* img, uint32_t piece_stride, int k, int n, float x) {
  unsigned char* dst = img + (k >> 3) * kB_LBO + (k & 7) * 16 + (n >> 3) * kB_SBO + (n & 7) * 2;
  if (PREC == 0) {
    const uint32_t p = pack_bf16x2(x, 0.0f);
    const float r = x - __uint_as_float(p << 16);
    const uint32_t q = pack_bf16x2(r, 0.0f);
    const float s = r - __uint_as_float(q << 16);
    const uint32_t t = pack_bf16x2(s, 0.0f);
    *reinterpret_cast<unsigned short*>(dst) = (unsigned short)(p & 0xffffu);
    *reinterpret_cast<unsigned short*>(dst + piece_stride) = (unsigned short)(q & 0xffffu);
    *reinterpret_cast<unsigned short*>(dst + 2 * piece_stride) = (unsigned short)(t & 0xffffu);
  } else {
    const uint32_t p = pack_f16x2(x, 0.0f);
    float h0, h1;
    unpack_f16x2(p, h0, h1);
    const uint32_t q = pack_f16x2(x - h0, 0.0f);
    *reinterpret_cast<unsigned short*>(dst) = (unsigned short)(p & 0xffffu);
    *reinterpret_cast<unsigned short*>(dst + piece_stride) = (unsigned short)(q & 0xffffu);
  }
}

struct Ctx {
  SmemTC* S;
  uint32_t tmem;
  int warp, lane;
  uint32_t phase;  // parity of the next MMA completion
};

// one dense layer on the tensor pipe + its epilogue; every thread of the CTA calls this (ends with __syncthreads)
template <int L, bool FWD, int PREC, class Epi>
__device__ __forceinline__ void tc_layer(Ctx& c, const unsigned char* src, uint32_t src_stride, int out_rows, Epi epi) {
  SmemTC& S = *c.S;
  if (c.warp == kIssueWarp) {
    tc_fence_after();
    if (elect_one()) {
      if (FWD) issue_fwd<L, PREC>(S, c.tmem, src, src_stride);
      else issue_bwd<L, PREC>(S, c.tmem, src, src_stride);
      umma_commit(&S.bar_mma);
    }
    __syncwarp();
  }
  if (c.warp < kEpiWarps) {
    const int quarter = c.warp & 3, half = c.warp >> 2;
    if (quarter * 32 < out_rows) {
      mbar_wait(&S.bar_mma, c.phase);
      tc_fence_after();
      float v[16];
      tmem_ld16(c.tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(16 * half), v);
      tmem_ld_wait();
      const int k = quarter * 32 + c.lane;
      if (k < out_rows) epi(k, half, v);
      tc_fence_before();
    }
    fence_proxy_async();
  }
  c.phase ^= 1u;
  __syncthreads();
}

template <int PREC>
__global__ void __launch_bounds__(kWarps * 32, 1) dp_frame_tc_kernel(const __grid_constant__ DpFrameArgs A) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemTC& S = *reinterpret_cast<SmemTC*>(smem_raw);
  const DpModelImageTC& M = S.M;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&S.bar_w, 1);
    mbar_init(&S.bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == kIssueWarp) tmem_alloc(&S.tmem_base, 32);
  for (int i = threadIdx.x; i < (int)(sizeof(S.ping) + sizeof(S.pong)) / 16; i += kWarps * 32)
    reinterpret_cast<uint4*>(&S.ping[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);  // ping and pong are contiguous; pad rows must stay finite
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    constexpr uint32_t kBytes = (uint32_t)sizeof(DpModelImageTC);
    constexpr uint32_t kChunk = 32768;
    mbar_expect_tx(&S.bar_w, kBytes);
    for (uint32_t o = 0; o < kBytes; o += kChunk)
      tma_bulk_g2s(reinterpret_cast<unsigned char*>(&S.M) + o, reinterpret_cast<const unsigned char*>(PREC ? A.model_tc16 : A.model_tc) + o, min(kChunk, kBytes - o), &S.bar_w);
  }
  Ctx ctx{&S, S.tmem_base, warp, lane, 0u};
  constexpr int CPW = NC / kWarps;  // 2 clips per warp in the per-clip phases
  const int n0 = warp * CPW;        // local clip index (tile column) of this warp's first clip
  const int clip0 = blockIdx.x * A.clips_per_cta + n0;

  // ---- per-clip frame inputs (same as the fp32 kernel)
  bool valid[CPW];
  float g[CPW][4], inv3e[CPW], lrot9e[CPW];
  float2 z[CPW], tl[CPW], am[CPW], av[CPW], zlast[CPW];
#pragma unroll
  for (int c = 0; c < CPW; ++c) {
    const int clip = clip0 + c;
    valid[c] = clip < A.n_clips && n0 + c < A.clips_per_cta;
    const int cc = valid[c] ? clip : 0;
    const int ne = A.n_ee ? A.n_ee[cc] : A.ee_stride;
    inv3e[c] = 1.0f / (3.0f * (float)ne);
    lrot9e[c] = A.lambda_rot / (9.0f * (float)ne);
#pragma unroll
    for (int i = 0; i < 4; ++i) g[c][i] = A.grot[cc * 4 + i];
    z[c] = tl[c] = make_float2(0.f, 0.f);
    if (lane < DP_L / 2 && valid[c]) {
      z[c] = reinterpret_cast<const float2*>(A.latent + (size_t)cc * DP_L)[lane];
      tl[c] = reinterpret_cast<const float2*>(A.target_buf + ((size_t)cc * A.target_rows + A.target_index) * DP_L)[lane];
    }
    am[c] = av[c] = make_float2(0.f, 0.f);
    zlast[c] = z[c];
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
    S.trk[n0 + c][lane] = row;
    if (lane < DP_L / 2) {  // latent -> B operand of the first layer
      store_piece_scalar<PREC>(&S.ping[0][0], kPingBytes, 2 * lane, n0 + c, z[c].x);
      store_piece_scalar<PREC>(&S.ping[0][0], kPingBytes, 2 * lane + 1, n0 + c, z[c].y);
    }
  }
  fence_proxy_async();
  mbar_wait(&S.bar_w, 0);  // model image (weights, statistics, skeleton tables) has landed

  constexpr float wsc = PREC ? 1.0f / kWScale : 1.0f;  // undoes the weight-image scaling of the fp16 path
  unsigned neg0 = 0, neg1 = 0;  // LeakyReLU slope bits of (feature k, this thread's 16 clips) for the backward pass
  auto forward = [&]() {
    tc_layer<0, true, PREC>(ctx, &S.ping[0][0], kPingBytes, DP_H0, [&](int k, int half, float (&v)[16]) {
      const float b = M.b0[k];
      neg0 = 0;
#pragma unroll
      for (int i = 0; i < 16; ++i) { v[i] = fmaf(v[i], wsc, b); neg0 |= (v[i] > 0.f ? 0u : 1u) << i; v[i] = lrelu(v[i]); }
      store_pieces<PREC>(&S.pong[0][0], kPongBytes, k, half, v);
    });
    tc_layer<1, true, PREC>(ctx, &S.pong[0][0], kPongBytes, DP_H1, [&](int k, int half, float (&v)[16]) {
      const float b = M.b1[k];
      neg1 = 0;
#pragma unroll
      for (int i = 0; i < 16; ++i) { v[i] = fmaf(v[i], wsc, b); neg1 |= (v[i] > 0.f ? 0u : 1u) << i; v[i] = lrelu(v[i]); }
      store_pieces<PREC>(&S.ping[0][0], kPingBytes, k, half, v);
    });
    tc_layer<2, true, PREC>(ctx, &S.ping[0][0], kPingBytes, DP_Y, [&](int k, int half, float (&v)[16]) {
      const float b = M.b2[k];
#pragma unroll
      for (int i = 0; i < 16; ++i) S.ybuf[16 * half + i][k] = fmaf(v[i], wsc, b);
    });
  };
  auto backward = [&]() {
    if (warp < kEpiWarps) {  // dL/dy (fp32 rows written by the kinematics warps) -> B operand
      const int k = (warp & 3) * 32 + lane, half = warp >> 2;
      if (k < DP_Y) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = S.ybuf[16 * half + i][k];
        store_pieces<PREC>(&S.pong[0][0], kPongBytes, k, half, v);
      }
      fence_proxy_async();
    }
    __syncthreads();
    tc_layer<2, false, PREC>(ctx, &S.pong[0][0], kPongBytes, DP_H1, [&](int k, int half, float (&v)[16]) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] *= ((neg1 >> i) & 1u) ? 0.2f * wsc : wsc;
      store_pieces<PREC>(&S.ping[0][0], kPingBytes, k, half, v);
    });
    tc_layer<1, false, PREC>(ctx, &S.ping[0][0], kPingBytes, DP_H0, [&](int k, int half, float (&v)[16]) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] *= ((neg0 >> i) & 1u) ? 0.2f * wsc : wsc;
      store_pieces<PREC>(&S.pong[0][0], kPongBytes, k, half, v);
    });
    tc_layer<0, false, PREC>(ctx, &S.pong[0][0], kPongBytes, DP_L, [&](int k, int half, float (&v)[16]) {
#pragma unroll
      for (int i = 0; i < 16; ++i) S.zgrad[16 * half + i][k] = v[i] * (PREC ? wsc * S.bscale[16 * half + i] : 1.0f);
    });
  };

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
    if (!__syncthreads_or(any ? 1 : 0)) break;  // also publishes the latent pieces written by the Adam lanes
#pragma unroll
    for (int c = 0; c < CPW; ++c)
      if (active[c]) zlast[c] = z[c];
    forward();
    float nlp[CPW], nlr[CPW];
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
      nlp[c] = lp[c];
      nlr[c] = lr[c];
      if (active[c]) {
        const FkOut o = fk_loss<true, false>(M, &S.ybuf[n0 + c][0], &S.trk[n0 + c][0], g[c], inv3e[c], lrot9e[c], lane, nullptr, nullptr,
                                             nullptr, nullptr);
        nlp[c] = o.lp;
        nlr[c] = o.lr;
        if (PREC) {  // bring the largest |dL/dy| component of this clip into [16, 32) with an exact power of two
          float4* row = reinterpret_cast<float4*>(&S.ybuf[n0 + c][0]);
          float4 yb = lane < 23 ? row[lane] : make_float4(0.f, 0.f, 0.f, 0.f);
          float mx = fmaxf(fmaxf(fabsf(yb.x), fabsf(yb.y)), fmaxf(fabsf(yb.z), fabsf(yb.w)));
#pragma unroll
          for (int sh = 16; sh > 0; sh >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, sh));
          int e = (int)((__float_as_uint(mx) >> 23) & 0xffu) - 127;  // floor(log2(mx)) for normal mx
          e = mx > 0.f ? max(-100, min(100, e)) : 4;
          const float sc = __uint_as_float((uint32_t)(127 + 4 - e) << 23), isc = __uint_as_float((uint32_t)(127 - 4 + e) << 23);
          if (lane < 23) row[lane] = make_float4(yb.x * sc, yb.y * sc, yb.z * sc, yb.w * sc);
          if (lane == 0) S.bscale[n0 + c] = isc;
          __syncwarp();
        }
      }
    }
    __syncthreads();
    backward();
    const float step_size = A.adam_tab[it], inv_bc2s = A.adam_tab[A.max_iter + it];  // lr/(1-b1^k), 1/sqrt(1-b2^k)
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
      const float dx = z[c].x - tl[c].x, dy = z[c].y - tl[c].y;
      float s = (lane < DP_L / 2) ? fmaf(dx, dx, dy * dy) : 0.0f;
      s = warp_sum(s);
      const float nlt = s * lt_scale;
      float gx = 0.f, gy = 0.f;
      if (lane < DP_L / 2) {
        gx = fmaf(2.0f * lt_scale, dx, S.zgrad[n0 + c][2 * lane]);
        gy = fmaf(2.0f * lt_scale, dy, S.zgrad[n0 + c][2 * lane + 1]);
      }
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
        z[c].x += __fdividef(-step_size * am[c].x, fmaf(fast_sqrt(av[c].x), inv_bc2s, 1e-8f));
        z[c].y += __fdividef(-step_size * am[c].y, fmaf(fast_sqrt(av[c].y), inv_bc2s, 1e-8f));
      }
      // the latent rows of the ping image were overwritten by the a1 / dL/dh1 pieces of this iteration: restore them for
      // EVERY clip (stopped and padding clips included) so that no column ever feeds back on its own garbage -- a
      // non-finite value in a K-padding row would poison the column through 0 x NaN
      if (lane < DP_L / 2) {
        store_piece_scalar<PREC>(&S.ping[0][0], kPingBytes, 2 * lane, n0 + c, z[c].x);
        store_piece_scalar<PREC>(&S.ping[0][0], kPingBytes, 2 * lane + 1, n0 + c, z[c].y);
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
    fence_proxy_async();
  }

  // ---- frame epilogue (drag_pose.py:369-414) from the LAST EVALUATED latent (pre-step)
#pragma unroll
  for (int c = 0; c < CPW; ++c)
    if (lane < DP_L / 2) {
      store_piece_scalar<PREC>(&S.ping[0][0], kPingBytes, 2 * lane, n0 + c, zlast[c].x);
      store_piece_scalar<PREC>(&S.ping[0][0], kPingBytes, 2 * lane + 1, n0 + c, zlast[c].y);
    }
  fence_proxy_async();
  __syncthreads();
  forward();
#pragma unroll
  for (int c = 0; c < CPW; ++c) {
    if (!valid[c]) continue;
    const int clip = clip0 + c;
    float q[4], r[4], p[3], d[3];
    fk_loss<false, true>(M, &S.ybuf[n0 + c][0], &S.trk[n0 + c][0], g[c], inv3e[c], lrot9e[c], lane, q, r, p, d);
    if (A.eval_only) {
      if (lane < DP_J && A.eval_pos) {
        float* o = A.eval_pos + ((size_t)clip * DP_J + lane) * 3;
        o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
      }
      if (lane == 0 && A.out_losses) { A.out_losses[clip * 3] = lp[c]; A.out_losses[clip * 3 + 1] = lr[c]; A.out_losses[clip * 3 + 2] = lt[c]; }
      continue;
    }
    float p0[3], gp[3], adj[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      p0[i] = __shfl_sync(0xffffffffu, p[i], 0);
      gp[i] = A.gpos[clip * 3 + i] + p0[i];
    }
    if (A.adj_joint >= 0) {
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
    if (lane < DP_J) {
      const float4 mq = reinterpret_cast<const float4*>(M.mean_q)[lane];
      const float4 sq = reinterpret_cast<const float4*>(M.std_q)[lane];
      const float* s = (lane == 0) ? r : q;
      reinterpret_cast<float4*>(A.out_pose + (size_t)clip * 88)[lane] =
          make_float4((s[0] - mq.x) / sq.x, (s[1] - mq.y) / sq.y, (s[2] - mq.z) / sq.z, (s[3] - mq.w) / sq.w);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kIssueWarp) tmem_dealloc(ctx.tmem, 32);
}

}  // namespace

cudaError_t dp_frame_tc_launch(const DpFrameArgs& args, int num_sms, bool fp16, cudaStream_t stream) {
  static bool configured = false;
  const size_t smem = sizeof(SmemTC) + 1024;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dp_frame_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(dp_frame_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  // 4096 clips: 28 real clips per 32-column tile -> 147 CTAs, one per SM, instead of 128 CTAs of 32
  DpFrameArgs a = args;
  int cpc = (args.n_clips + num_sms - 1) / num_sms;
  cpc = cpc < 1 ? 1 : (cpc > NC ? NC : cpc);
  if (args.n_clips > num_sms * NC) cpc = NC;  // several waves anyway: use full tiles
  a.clips_per_cta = cpc;
  const int grid = (args.n_clips + cpc - 1) / cpc;
  if (fp16) dp_frame_tc_kernel<1><<<grid, kWarps * 32, smem, stream>>>(a);
  else dp_frame_tc_kernel<0><<<grid, kWarps * 32, smem, stream>>>(a);
  return cudaGetLastError();
}
