// dp_frame_tc.cu -- persistent per-frame optimisation kernel with the decoder on tcgen05 tensor cores, bf16x3 split products,
// weights in shared memory (decoder path 2; the default batch path is dp_frame_tc16.cu, which this design preceded).
//
// Same contract as dp_frame_simt.cu (one launch == one frame of DragPose.run, python/src/drag_pose.py:196-414,
// for every clip) but the six dense layers of an iteration (decoder forward 24->40->60->92 and its
// data-gradient 92->60->40->24, python/src/autoencoder.py:224-256 folded) run as tcgen05.mma tiles:
//
//   D[features(M=128) x clips(N=32)] = W[features x K] . X[K x clips]          accumulator in tensor memory
//
// * transposed mapping: the WEIGHTS are the 128-row A operand, the CTA's 32 clips are the N dimension, so
//   4096 clips occupy 128 SMs (clips-as-M would fill only 32).
// * precision: bf16x3 split products on kind::f16 -- x = x1 + x2 + x3 (bf16), six products
//   (3,1)(2,2)(1,3)(2,1)(1,2)(1,1) accumulated in fp32: 1.3e-7 relative per GEMM, gradient 3.0e-6 relative (the fp32
//   kernel: 1.8e-6; plain TF32 would be 1.8e-3 and bf16x2 1e-4 -- both break the 1e-4 bar).  kind::f16 rather than
//   kind::tf32 because tf32 produces zeros for MN-major operands in the no-swizzle layout (measured) while bf16 accepts
//   them: ONE weight image serves the forward GEMM (A K-major) and the transposed backward GEMM (A MN-major), 63 KB.
// * activations never leave the SM: epilogue warps read the accumulator with tcgen05.ld (lane == feature),
//   apply bias / LeakyReLU (slope bits stay in registers for the backward pass), split to bf16 pieces and
//   write the next layer's B operand (MN-major image, one STS.128 per 8 clips).
// * kinematics, loss and adjoint: the packed two-clip pass of dp_fk2.cuh; Adam and early stopping per warp.
#include "dp_fk2.cuh"
#include "dp_internal.h"
#include "dp_umma.cuh"

namespace {

// Tile geometry.  NC clips (the UMMA N dimension) per CTA, two clips per warp in the per-clip phases, eight epilogue warps
// (TMEM lane quarter = warp % 4, clip half = warp / 4).  Instantiated for NC = 32: 16 warps, one CTA per SM (the three-piece
// weight image does not fit twice).
template <int NC>
struct Geo {
  static constexpr int kWarps = NC / 2;
  static constexpr int EC = NC / 2;                      // clips per epilogue thread
  static constexpr int kIssueWarp = kWarps > 8 ? 8 : 0;  // a non-epilogue warp when there is one
  static constexpr uint32_t kB_LBO = 128 * (NC / 8);     // activation image: K 8-groups LBO apart, clip 8-groups 128 B apart
  static constexpr uint32_t kPingBytes = (64 / 8) * kB_LBO, kPongBytes = (96 / 8) * kB_LBO;  // per 16-bit piece
};
constexpr int kEpiWarps = 8;
constexpr uint32_t kB_SBO = 128;
constexpr int kMaxPieces = 3;
enum { ST_Z = 0, ST_TL = 1, ST_M = 2, ST_V = 3, ST_ZLAST = 4 };

template <int NC>
struct SmemTC {
  static constexpr int NP = 3;  // bf16 pieces
  // the first NP pieces of the model image (weights are its last member, see DpModelImageTC)
  __align__(16) unsigned char model[DP_TC_IMAGE_BYTES(NP)];
  __align__(16) unsigned char ping[NP][Geo<NC>::kPingBytes];  // [piece]  z (24) / a1 (60) / dL/dh1 (60)
  __align__(16) unsigned char pong[NP][Geo<NC>::kPongBytes];  // [piece]  a0 (40) / dL/dy (92) / dL/dh0 (40)
  __align__(16) float ybuf[NC][96];                 // y, then dL/dy in place (fp32, one row per clip)
  float zgrad[NC][25];                              // dL/dz from the decoder (fp32)
  __align__(16) ClipTrackers trk[NC][32];
  // per-clip optimiser state lives here, not in registers: the packed kinematics pass needs the register file
  __align__(16) float2 st[NC][5][DP_L / 2];         // [ST_Z latent | ST_TL target latent | ST_M, ST_V Adam moments | ST_ZLAST]
  __align__(16) float groot[NC][4];                 // previous world root rotation g (wxyz)
  __align__(16) float2 fkscr[NC / 2][16];           // per clip pair: R_0, r, d parked between the two halves of the kinematics pass
  double prev[NC];                                  // previous total loss (early stopping compares in double)
  float loss[NC][3];                                // last evaluated lp, lr, lt
  int iters[NC];
  uint64_t bar_w, bar_mma;
  uint32_t tmem_base;
  __device__ __forceinline__ const DpModelImageTC& M() const { return *reinterpret_cast<const DpModelImageTC*>(model); }
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
__device__ __forceinline__ void issue_terms(uint32_t tmem, const UmmaDescBase (&a)[kMaxPieces], const UmmaDescBase (&b)[kMaxPieces], uint32_t ao,
                                            uint32_t bo, uint32_t idesc, bool first) {
  // (3,1) (2,2) (1,3) | (2,1) (1,2) | (1,1)
  if (first) umma_f16_c<false>(tmem, umma_desc_at(a[2], ao), umma_desc_at(b[0], bo), idesc);
  else umma_f16_c<true>(tmem, umma_desc_at(a[2], ao), umma_desc_at(b[0], bo), idesc);
  umma_f16_c<true>(tmem, umma_desc_at(a[1], ao), umma_desc_at(b[1], bo), idesc);
  umma_f16_c<true>(tmem, umma_desc_at(a[0], ao), umma_desc_at(b[2], bo), idesc);
  umma_f16_c<true>(tmem, umma_desc_at(a[1], ao), umma_desc_at(b[0], bo), idesc);
  umma_f16_c<true>(tmem, umma_desc_at(a[0], ao), umma_desc_at(b[1], bo), idesc);
  umma_f16_c<true>(tmem, umma_desc_at(a[0], ao), umma_desc_at(b[0], bo), idesc);
}

// forward layer L: A = W_L (K-major: LBO 128 between input 8-groups, SBO between output-row 8-groups)
template <int L, int NC>
__device__ __forceinline__ void issue_fwd(const SmemTC<NC>& S, uint32_t tmem, const unsigned char* src, uint32_t piece_stride) {
  constexpr uint32_t kB_LBO = Geo<NC>::kB_LBO;
  constexpr int NP = SmemTC<NC>::NP;
  constexpr uint32_t sbo = 128 * (Lay<L>::kin / 8);
  constexpr uint32_t fmt = 1u;  // kind::f16 operand format: 1 = bf16
  constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 16) | ((uint32_t)(NC >> 3) << 17) | (8u << 24);  // B MN-major
  UmmaDescBase a[kMaxPieces], b[kMaxPieces];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    a[p] = umma_desc_base(smem_u32(S.M().w[p]) + Lay<L>::off, 128, sbo);
    b[p] = umma_desc_base(smem_u32(src) + p * piece_stride, kB_LBO, kB_SBO);
  }
#pragma unroll
  for (int k = 0; k < Lay<L>::kin / 16; ++k) {
    const uint32_t ao = k * 256, bo = k * 2 * kB_LBO;
    issue_terms(tmem, a, b, ao, bo, idesc, k == 0);
  }
}
// backward of layer L: the SAME weight image read MN-major (M = inputs, K = outputs): LBO / SBO swap roles
template <int L, int NC>
__device__ __forceinline__ void issue_bwd(const SmemTC<NC>& S, uint32_t tmem, const unsigned char* src, uint32_t piece_stride) {
  constexpr uint32_t kB_LBO = Geo<NC>::kB_LBO;
  constexpr int NP = SmemTC<NC>::NP;
  constexpr uint32_t wsbo = 128 * (Lay<L>::kin / 8);
  constexpr uint32_t fmt = 1u;
  constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NC >> 3) << 17) | (8u << 24);
  UmmaDescBase a[kMaxPieces], b[kMaxPieces];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    a[p] = umma_desc_base(smem_u32(S.M().w[p]) + Lay<L>::off, wsbo, 128);
    b[p] = umma_desc_base(smem_u32(src) + p * piece_stride, kB_LBO, kB_SBO);
  }
#pragma unroll
  for (int k = 0; k < Lay<L>::kout / 16; ++k) {
    const uint32_t ao = k * 2 * wsbo, bo = k * 2 * kB_LBO;
    issue_terms(tmem, a, b, ao, bo, idesc, k == 0);
  }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {  // lower address <- lo_elem
  uint32_t p;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(hi_elem), "f"(lo_elem));
  return p;
}
// EC fp32 values (feature k, clips EC*half .. +EC-1) -> split pieces in an MN-major activation image
template <int NC>
__device__ __forceinline__ void store_pieces(unsigned char* img, uint32_t piece_stride, int k, int half, const float (&v)[Geo<NC>::EC]) {
  constexpr int EC = Geo<NC>::EC;
  unsigned char* dst = img + (k >> 3) * Geo<NC>::kB_LBO + (k & 7) * 16 + ((EC / 8) * half) * kB_SBO;
#pragma unroll
  for (int g = 0; g < EC / 8; ++g) {
    uint32_t p1[4], p2[4], p3[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float x0 = v[8 * g + 2 * i], x1 = v[8 * g + 2 * i + 1];
      p1[i] = pack_bf16x2(x0, x1);
      const float r0 = x0 - __uint_as_float(p1[i] << 16), r1 = x1 - __uint_as_float(p1[i] & 0xffff0000u);
      p2[i] = pack_bf16x2(r0, r1);
      const float s0 = r0 - __uint_as_float(p2[i] << 16), s1 = r1 - __uint_as_float(p2[i] & 0xffff0000u);
      p3[i] = pack_bf16x2(s0, s1);
    }
    *reinterpret_cast<uint4*>(dst + g * kB_SBO) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
    *reinterpret_cast<uint4*>(dst + piece_stride + g * kB_SBO) = make_uint4(p2[0], p2[1], p2[2], p2[3]);
    *reinterpret_cast<uint4*>(dst + 2 * piece_stride + g * kB_SBO) = make_uint4(p3[0], p3[1], p3[2], p3[3]);
  }
}
// a single fp32 value (feature k, clip n) -> its pieces (used by the Adam lanes for the latent)
template <int NC>
__device__ __forceinline__ void store_piece_scalar(unsigned char* img, uint32_t piece_stride, int k, int n, float x) {
  unsigned char* dst = img + (k >> 3) * Geo<NC>::kB_LBO + (k & 7) * 16 + (n >> 3) * kB_SBO + (n & 7) * 2;
  const uint32_t p = pack_bf16x2(x, 0.0f);
  const float r = x - __uint_as_float(p << 16);
  const uint32_t q = pack_bf16x2(r, 0.0f);
  const float s = r - __uint_as_float(q << 16);
  const uint32_t t = pack_bf16x2(s, 0.0f);
  *reinterpret_cast<unsigned short*>(dst) = (unsigned short)(p & 0xffffu);
  *reinterpret_cast<unsigned short*>(dst + piece_stride) = (unsigned short)(q & 0xffffu);
  *reinterpret_cast<unsigned short*>(dst + 2 * piece_stride) = (unsigned short)(t & 0xffffu);
}

template <int NC>
struct Ctx {
  SmemTC<NC>* S;
  uint32_t tmem;
  int warp, lane;
  uint32_t phase;  // parity of the next MMA completion
};

// one dense layer on the tensor pipe + its epilogue; every thread of the CTA calls this (ends with __syncthreads)
template <int L, bool FWD, int NC, class Epi>
__device__ __forceinline__ void tc_layer(Ctx<NC>& c, const unsigned char* src, uint32_t src_stride, int out_rows, Epi epi) {
  SmemTC<NC>& S = *c.S;
  constexpr int EC = Geo<NC>::EC;
  if (c.warp == Geo<NC>::kIssueWarp) {
    tc_fence_after();
    if (elect_one()) {
      if (FWD) issue_fwd<L, NC>(S, c.tmem, src, src_stride);
      else issue_bwd<L, NC>(S, c.tmem, src, src_stride);
      umma_commit(&S.bar_mma);
    }
    __syncwarp();
  }
  if (c.warp < kEpiWarps) {
    const int quarter = c.warp & 3, half = c.warp >> 2;
    if (quarter * 32 < out_rows) {
      mbar_wait(&S.bar_mma, c.phase);
      tc_fence_after();
      float v[EC];
      if constexpr (EC == 16) tmem_ld16(c.tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(EC * half), reinterpret_cast<float (&)[16]>(v));
      else tmem_ld8(c.tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(EC * half), reinterpret_cast<float (&)[8]>(v));
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

template <int NC>
__global__ void __launch_bounds__(Geo<NC>::kWarps * 32, 1) dp_frame_tc_kernel(const __grid_constant__ DpFrameArgs A) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  using Smem = SmemTC<NC>;
  constexpr int kWarps = Geo<NC>::kWarps, EC = Geo<NC>::EC;
  constexpr uint32_t kPingBytes = Geo<NC>::kPingBytes, kPongBytes = Geo<NC>::kPongBytes;
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const DpModelImageTC& M = S.M();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&S.bar_w, 1);
    mbar_init(&S.bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == Geo<NC>::kIssueWarp) tmem_alloc(&S.tmem_base, 32);
  for (int i = threadIdx.x; i < (int)(sizeof(S.ping) + sizeof(S.pong)) / 16; i += kWarps * 32)
    reinterpret_cast<uint4*>(&S.ping[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);  // ping and pong are contiguous; pad rows must stay finite
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    constexpr uint32_t kBytes = (uint32_t)sizeof(S.model);
    constexpr uint32_t kChunk = 32768;
    mbar_expect_tx(&S.bar_w, kBytes);
    for (uint32_t o = 0; o < kBytes; o += kChunk)
      tma_bulk_g2s(S.model + o, reinterpret_cast<const unsigned char*>(A.model_tc) + o, min(kChunk, kBytes - o), &S.bar_w);
  }
  Ctx<NC> ctx{&S, S.tmem_base, warp, lane, 0u};
  constexpr int CPW = NC / kWarps;  // 2 clips per warp in the per-clip phases
  const int n0 = warp * CPW;        // local clip index (tile column) of this warp's first clip
  const int clip0 = blockIdx.x * A.clips_per_cta + n0;

  // ---- per-clip frame inputs (same as the fp32 kernel)
  static_assert(CPW == 2, "the kinematics pass packs exactly two clips per warp");
  bool valid[CPW];
  float inv3e[CPW], lrot9e[CPW];
#pragma unroll
  for (int c = 0; c < CPW; ++c) {
    const int clip = clip0 + c;
    valid[c] = clip < A.n_clips && n0 + c < A.clips_per_cta;
    const int cc = valid[c] ? clip : 0;
    const int ne = A.n_ee ? A.n_ee[cc] : A.ee_stride;
    inv3e[c] = 1.0f / (3.0f * (float)ne);
    lrot9e[c] = A.lambda_rot / (9.0f * (float)ne);
    if (lane < 4) S.groot[n0 + c][lane] = A.grot[cc * 4 + lane];
    float2 z = make_float2(0.f, 0.f), tl = z;
    if (lane < DP_L / 2 && valid[c]) {
      z = reinterpret_cast<const float2*>(A.latent + (size_t)cc * DP_L)[lane];
      tl = reinterpret_cast<const float2*>(A.target_buf + ((size_t)cc * A.target_rows + A.target_index) * DP_L)[lane];
    }
    if (lane < DP_L / 2) {
      S.st[n0 + c][ST_Z][lane] = z;
      S.st[n0 + c][ST_TL][lane] = tl;
      S.st[n0 + c][ST_M][lane] = make_float2(0.f, 0.f);
      S.st[n0 + c][ST_V][lane] = make_float2(0.f, 0.f);
      S.st[n0 + c][ST_ZLAST][lane] = z;
    }
    if (lane == 0) {
      S.prev[n0 + c] = 10000000.0;
      S.loss[n0 + c][0] = S.loss[n0 + c][1] = S.loss[n0 + c][2] = __int_as_float(0x7f800000);
      S.iters[n0 + c] = 0;
    }
    ClipTrackers row;
    row.pw = row.r0 = row.r1 = row.r2 = make_float4(0.f, 0.f, 0.f, 0.f);
    float origin[3] = {0.f, 0.f, 0.f};  // world-absolute targets are taken relative to the clip's current global position
    if (A.targets_world) { origin[0] = A.gpos[cc * 3]; origin[1] = A.gpos[cc * 3 + 1]; origin[2] = A.gpos[cc * 3 + 2]; }
    const int32_t* jn = A.joints + (A.shared_trackers ? 0 : (size_t)cc * A.ee_stride);
    const float* wt = A.weights + (A.shared_trackers ? 0 : (size_t)cc * A.ee_stride * 2);
    for (int e = 0; e < ne; ++e) {
      if (jn[e] == lane) {
        const float* tp = A.tgt_pos + ((size_t)cc * A.ee_stride + e) * 3;
        const float* tr = A.tgt_rot + ((size_t)cc * A.ee_stride + e) * 9;
        row.pw = make_float4(tp[0] - origin[0], tp[1] - origin[1], tp[2] - origin[2], wt[2 * e]);
        row.r0 = make_float4(tr[0], tr[1], tr[2], wt[2 * e + 1]);
        row.r1 = make_float4(tr[3], tr[4], tr[5], 0.f);
        row.r2 = make_float4(tr[6], tr[7], tr[8], 0.f);
      }
    }
    S.trk[n0 + c][lane] = row;
    if (lane < DP_L / 2) {  // latent -> B operand of the first layer
      store_piece_scalar<NC>(&S.ping[0][0], kPingBytes, 2 * lane, n0 + c, z.x);
      store_piece_scalar<NC>(&S.ping[0][0], kPingBytes, 2 * lane + 1, n0 + c, z.y);
    }
  }
  const P2 inv3e2 = mk2(inv3e[0], inv3e[1]), lrot9e2 = mk2(lrot9e[0], lrot9e[1]);
  __syncwarp();
  fence_proxy_async();
  mbar_wait(&S.bar_w, 0);  // model image (weights, statistics, skeleton tables) has landed

  unsigned neg0 = 0, neg1 = 0;  // LeakyReLU slope bits of (feature k, this thread's 16 clips) for the backward pass
  auto forward = [&]() {
    tc_layer<0, true, NC>(ctx, &S.ping[0][0], kPingBytes, DP_H0, [&](int k, int half, float (&v)[EC]) {
      const float b = M.b0[k];
      neg0 = 0;
#pragma unroll
      for (int i = 0; i < EC; ++i) { v[i] += b; neg0 |= (v[i] > 0.f ? 0u : 1u) << i; v[i] = lrelu(v[i]); }
      store_pieces<NC>(&S.pong[0][0], kPongBytes, k, half, v);
    });
    tc_layer<1, true, NC>(ctx, &S.pong[0][0], kPongBytes, DP_H1, [&](int k, int half, float (&v)[EC]) {
      const float b = M.b1[k];
      neg1 = 0;
#pragma unroll
      for (int i = 0; i < EC; ++i) { v[i] += b; neg1 |= (v[i] > 0.f ? 0u : 1u) << i; v[i] = lrelu(v[i]); }
      store_pieces<NC>(&S.ping[0][0], kPingBytes, k, half, v);
    });
    tc_layer<2, true, NC>(ctx, &S.ping[0][0], kPingBytes, DP_Y, [&](int k, int half, float (&v)[EC]) {
      const float b = M.b2[k];
#pragma unroll
      for (int i = 0; i < EC; ++i) S.ybuf[EC * half + i][k] = v[i] + b;
    });
  };
  auto backward = [&]() {
    if (warp < kEpiWarps) {  // dL/dy (fp32 rows written by the kinematics warps) -> B operand
      const int k = (warp & 3) * 32 + lane, half = warp >> 2;
      if (k < DP_Y) {
        float v[EC];
#pragma unroll
        for (int i = 0; i < EC; ++i) v[i] = S.ybuf[EC * half + i][k];
        store_pieces<NC>(&S.pong[0][0], kPongBytes, k, half, v);
      }
      fence_proxy_async();
    }
    __syncthreads();
    tc_layer<2, false, NC>(ctx, &S.pong[0][0], kPongBytes, DP_H1, [&](int k, int half, float (&v)[EC]) {
#pragma unroll
      for (int i = 0; i < EC; ++i) v[i] *= ((neg1 >> i) & 1u) ? 0.2f : 1.0f;
      store_pieces<NC>(&S.ping[0][0], kPingBytes, k, half, v);
    });
    tc_layer<1, false, NC>(ctx, &S.ping[0][0], kPingBytes, DP_H0, [&](int k, int half, float (&v)[EC]) {
#pragma unroll
      for (int i = 0; i < EC; ++i) v[i] *= ((neg0 >> i) & 1u) ? 0.2f : 1.0f;
      store_pieces<NC>(&S.pong[0][0], kPongBytes, k, half, v);
    });
    tc_layer<0, false, NC>(ctx, &S.pong[0][0], kPongBytes, DP_L, [&](int k, int half, float (&v)[EC]) {
#pragma unroll
      for (int i = 0; i < EC; ++i) S.zgrad[EC * half + i][k] = v[i];
    });
  };

  // ---- optimisation loop (drag_pose.py:296-355)
  // early-stop test of the NEXT iteration, evaluated right after each Adam step (initial losses are +inf, initial increment 1)
  bool active[CPW] = {valid[0] && 1.0 > A.min_incr, valid[1] && 1.0 > A.min_incr};
  const float lt_scale = A.lambda_t * (1.0f / (float)DP_L);
  const bool clocked = A.phase_cycles != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  long long tick = clocked ? clock64() : 0;
  auto phase_done = [&](int i) {
    if (clocked) {
      const long long now = clock64();
      A.phase_cycles[i] += (unsigned long long)(now - tick);
      tick = now;
    }
  };
  for (int it = 0; it < A.max_iter; ++it) {
    if (!__syncthreads_or((active[0] || active[1]) ? 1 : 0)) break;  // also publishes the latent pieces written by the Adam lanes
    if (lane == 0) {  // the Adam phase reads two table entries: pull their lines into L1 now instead of stalling there
      asm volatile("prefetch.global.L1 [%0];" ::"l"(A.adam_tab + it));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(A.adam_tab + A.max_iter + it));
    }
    phase_done(3);
    forward();
    phase_done(0);
    float nlp[CPW] = {0.f, 0.f}, nlr[CPW] = {0.f, 0.f};
    if (active[0] || active[1]) {  // both clips of the warp in one packed pass; results of a stopped clip are discarded
      const FkOut2 o = fk_loss2<true, false>(M, &S.ybuf[n0][0], &S.ybuf[n0 + 1][0], &S.trk[n0][0], &S.trk[n0 + 1][0], &S.groot[n0][0], &S.fkscr[warp][0], inv3e2,
                                             lrot9e2, lane, nullptr, nullptr, nullptr, nullptr);
      phase_done(4);
      nlp[0] = o.lp.v.x; nlr[0] = o.lr.v.x;
      nlp[1] = o.lp.v.y; nlr[1] = o.lr.v.y;
    }
    __syncthreads();
    phase_done(1);
    backward();
    phase_done(2);
    const float step_size = A.adam_tab[it], inv_bc2s = A.adam_tab[A.max_iter + it];  // lr/(1-b1^k), 1/sqrt(1-b2^k)
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
      const int n = n0 + c;
      float2 z = make_float2(0.f, 0.f), tl = z;
      if (lane < DP_L / 2) { z = S.st[n][ST_Z][lane]; tl = S.st[n][ST_TL][lane]; }
      const float dx = z.x - tl.x, dy = z.y - tl.y;
      const float nlt = warp_sum(fmaf(dx, dx, dy * dy)) * lt_scale;
      float gx = 0.f, gy = 0.f;
      if (lane < DP_L / 2) {
        gx = fmaf(2.0f * lt_scale, dx, S.zgrad[n][2 * lane]);
        gy = fmaf(2.0f * lt_scale, dy, S.zgrad[n][2 * lane + 1]);
      }
      if (A.trace && active[c]) {
        float* row = A.trace + ((size_t)(clip0 + c) * A.trace_iters + it) * 52;
        if (lane < DP_L / 2) {
          reinterpret_cast<float2*>(row)[lane] = z;
          reinterpret_cast<float2*>(row + DP_L)[lane] = make_float2(gx, gy);
        }
        if (lane == 0) { row[48] = nlp[c]; row[49] = nlr[c]; row[50] = nlt; row[51] = 1.0f; }
      }
      if (A.eval_only) {
        if (active[c] && lane < DP_L / 2) reinterpret_cast<float2*>(A.eval_grad + (size_t)(clip0 + c) * DP_L)[lane] = make_float2(gx, gy);
      } else if (active[c] && lane < DP_L / 2) {
        float2 am = S.st[n][ST_M][lane], av = S.st[n][ST_V][lane];
        am.x = fmaf(0.1f, gx - am.x, am.x);
        am.y = fmaf(0.1f, gy - am.y, am.y);
        av.x = av.x * 0.999f + (0.001f * gx) * gx;
        av.y = av.y * 0.999f + (0.001f * gy) * gy;
        S.st[n][ST_M][lane] = am;
        S.st[n][ST_V][lane] = av;
        S.st[n][ST_ZLAST][lane] = z;  // the frame's output is decoded from the last EVALUATED latent
        z.x += __fdividef(-step_size * am.x, fmaf(fast_sqrt(av.x), inv_bc2s, 1e-8f));
        z.y += __fdividef(-step_size * am.y, fmaf(fast_sqrt(av.y), inv_bc2s, 1e-8f));
        S.st[n][ST_Z][lane] = z;
      }
      // the latent rows of the ping image were overwritten by the a1 / dL/dh1 pieces of this iteration: restore them for
      // EVERY clip (stopped and padding clips included) so that no column ever feeds back on its own garbage -- a
      // non-finite value in a K-padding row would poison the column through 0 x NaN
      if (lane < DP_L / 2) {
        store_piece_scalar<NC>(&S.ping[0][0], kPingBytes, 2 * lane, n, z.x);
        store_piece_scalar<NC>(&S.ping[0][0], kPingBytes, 2 * lane + 1, n, z.y);
      }
      if (active[c]) {
        const float total = (nlp[c] + nlr[c]) + nlt;
        const double incr = S.prev[n] - (double)total;
        __syncwarp();
        if (lane == 0) {
          S.prev[n] = (double)total;
          S.loss[n][0] = nlp[c]; S.loss[n][1] = nlr[c]; S.loss[n][2] = nlt;
          S.iters[n] += 1;
        }
        active[c] = ((double)nlp[c] > A.eps_pos || (double)nlr[c] > A.eps_rot) && (incr > A.min_incr);
      }
    }
    fence_proxy_async();
  }

  // ---- frame epilogue (drag_pose.py:369-414) from the LAST EVALUATED latent (pre-step)
  __syncwarp();
#pragma unroll
  for (int c = 0; c < CPW; ++c)
    if (lane < DP_L / 2) {
      const float2 zl = S.st[n0 + c][ST_ZLAST][lane];
      store_piece_scalar<NC>(&S.ping[0][0], kPingBytes, 2 * lane, n0 + c, zl.x);
      store_piece_scalar<NC>(&S.ping[0][0], kPingBytes, 2 * lane + 1, n0 + c, zl.y);
    }
  fence_proxy_async();
  __syncthreads();
  forward();
  P2 q2[4], r2[4], p2[3], d2[3];
  if (valid[0] || valid[1])
    fk_loss2<false, true>(M, &S.ybuf[n0][0], &S.ybuf[n0 + 1][0], &S.trk[n0][0], &S.trk[n0 + 1][0], &S.groot[n0][0], &S.fkscr[warp][0], inv3e2, lrot9e2, lane, q2, r2, p2, d2);
#pragma unroll
  for (int c = 0; c < CPW; ++c) {
    if (!valid[c]) continue;
    const int clip = clip0 + c;
    float q[4], r[4], p[3], d[3];
#pragma unroll
    for (int i = 0; i < 4; ++i) { q[i] = c ? q2[i].v.y : q2[i].v.x; r[i] = c ? r2[i].v.y : r2[i].v.x; }
#pragma unroll
    for (int i = 0; i < 3; ++i) { p[i] = c ? p2[i].v.y : p2[i].v.x; d[i] = c ? d2[i].v.y : d2[i].v.x; }
    if (A.eval_only) {
      if (lane < DP_J && A.eval_pos) {
        float* o = A.eval_pos + ((size_t)clip * DP_J + lane) * 3;
        o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
      }
      if (lane < 3 && A.out_losses) A.out_losses[clip * 3 + lane] = S.loss[n0 + c][lane];
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
        adj[i] = ((tp[i] - (A.targets_world ? A.gpos[clip * 3 + i] : 0.0f)) - pj) * A.adj_w;
        gp[i] += adj[i];
      }
    }
    __syncwarp();
    const int hs = M.height_slot[lane];
    if (hs >= 0) A.height_buf[((size_t)clip * DP_PAST + A.ring_head) * DP_NH + hs] = p[1] + gp[1];
    if (lane < DP_L / 2) {
      reinterpret_cast<float2*>(A.latent_buf + ((size_t)clip * DP_PAST + A.ring_head) * DP_L)[lane] = S.st[n0 + c][ST_ZLAST][lane];
      reinterpret_cast<float2*>(A.latent + (size_t)clip * DP_L)[lane] = S.st[n0 + c][ST_Z][lane];
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
      A.out_iters[clip] = S.iters[n0 + c];
      A.out_losses[clip * 3] = S.loss[n0 + c][0];
      A.out_losses[clip * 3 + 1] = S.loss[n0 + c][1];
      A.out_losses[clip * 3 + 2] = S.loss[n0 + c][2];
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
  if (warp == Geo<NC>::kIssueWarp) tmem_dealloc(ctx.tmem, 32);
}

}  // namespace

cudaError_t dp_frame_tc_launch(const DpFrameArgs& args, int num_sms, cudaStream_t stream) {
  static bool configured = false;
  const size_t smem = sizeof(SmemTC<32>) + 1024;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dp_frame_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  // 4096 clips: 28 real clips per 32-column tile -> 147 CTAs, one per SM, instead of 128 CTAs of 32
  DpFrameArgs a = args;
  int cpc = (args.n_clips + num_sms - 1) / num_sms;
  cpc = cpc < 1 ? 1 : (cpc > 32 ? 32 : cpc);
  if (args.n_clips > num_sms * 32) cpc = 32;  // several waves anyway: use full tiles
  a.clips_per_cta = cpc;
  dp_frame_tc_kernel<32><<<(args.n_clips + cpc - 1) / cpc, Geo<32>::kWarps * 32, smem, stream>>>(a);
  return cudaGetLastError();
}
