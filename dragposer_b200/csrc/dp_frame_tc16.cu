// dp_frame_tc16.cu -- persistent per-frame optimisation kernel, fp16x2 split products on tcgen05 with the WEIGHTS
// resident in tensor memory (default path for batches).
//
// Same contract and arithmetic as dp_frame_tc.cu (one launch == one frame of DragPose.run, python/src/drag_pose.py:196-414,
// for every clip; decoder forward 24->40->60->92 and its data-gradient as D[features x clips] = W . X tiles), restructured
// around what the phase clock and the MMA-rate probes measured on B200:
//   * an MMA whose A operand comes from shared memory costs ~41 cycles at N = 16 whatever the math; with A in tensor memory
//     (TS mode) it costs ~12.  The six weight matrices (forward W_l and backward W_l^T, two fp16 pieces each, 16 W) fit in
//     352 of the 512 tensor-memory columns, so they are loaded there once per launch and shared memory only carries the
//     activations (B operand, MN-major).
//   * the CUDA cores idle during the tensor phases and vice versa, so ONE CTA per SM runs TWO independent groups of
//     8 warps x 16 clips.  Each group has its own accumulator columns, mbarrier, named block barrier and loop exit; they
//     share the tensor-memory weights and drift apart naturally, one group's kinematics overlapping the other's layers.
//   * kinematics / loss / adjoint: two clips per warp in packed fp32x2 registers (dp_fk2.cuh); optimiser state in shared
//     memory.
#include "dp_fk2.cuh"
#include "dp_internal.h"
#include "dp_umma.cuh"

namespace {

constexpr int NC = 32;            // clip columns per CTA: two groups of 16 (UMMA N = 16)
constexpr int kWarps = 16;     // two groups of 8 warps
constexpr int EC = 8;             // clips per epilogue thread: a group's 8 warps = 4 TMEM lane quarters x 2 clip halves
// activation image (B operand, MN-major, no swizzle): element (feature k, clip column n) of a piece at
//   (k / 8) * kB_LBO + (k % 8) * 16 + (n / 8) * kB_SBO + (n % 8) * 2 bytes.
// The K 8-groups are padded from 512 to 528 bytes: at 512 every 8-group starts in the same bank and the 32 lanes of an epilogue
// warp (consecutive features, one 16-byte store each) collide four ways; 528 moves each group on by four banks.
constexpr uint32_t kB_LBO = 128 * (NC / 8) + 16;
constexpr uint32_t kB_SBO = 128;
constexpr uint32_t kPingBytes = (64 / 8) * kB_LBO, kPongBytes = (96 / 8) * kB_LBO;  // per fp16 piece
// dL/dy image (B operand of the first backward layer only), K-MAJOR: element (clip column n, feature k) of a piece at
//   (n / 8) * kD_SBO + (k / 8) * kD_LBO + (n % 8) * 16 + (k % 8) * 2 bytes,
// so that the kinematics lane of joint j stores its four values dL/dy[4j .. 4j+3] of a clip as ONE 8-byte word per piece -- the
// adjoint writes the tensor-core operand itself and the backward layers start after a single group barrier (round 1 wrote fp32
// rows, and the epilogue warps re-read them transposed, split them and stored them behind a second barrier).  kD_LBO is padded
// from 128 to 144 bytes for the same bank reason as kB_LBO.
constexpr uint32_t kD_LBO = 144, kD_SBO = (96 / 8) * kD_LBO, kDyBytes = (NC / 8) * kD_SBO;
constexpr float kWScale = 16.0f;  // the weight image holds 16 W; dL/dy is scaled per clip into [16, 32) (see dp_frame_tc.cu)
enum { ST_Z = 0, ST_TL = 1, ST_M = 2, ST_V = 3, ST_ZLAST = 4 };
// tensor-memory columns: accumulators of group g at 16 g; then the weight pieces (two K elements per 32-bit word)
constexpr uint32_t kT_D = 0, kT_W = 32, kT_COLS = 512;
static_assert(kT_W + DP_TC_TMEM_WORDS <= kT_COLS, "weights must fit in tensor memory");

struct SmemT {
  __align__(16) unsigned char model[DP_TC_IMAGE_BYTES(0)];   // biases, statistics, skeleton tables (no weight pieces)
  __align__(16) unsigned char ping[2][kPingBytes];  // [piece]  z (24) / a1 (60) / dL/dh1 (60)
  __align__(16) unsigned char pong[2][kPongBytes];  // [piece]  a0 (40) / dL/dy (92) / dL/dh0 (40)
  __align__(16) unsigned char dyimg[2][kDyBytes];   // [piece]  dL/dy (92), K-major
  __align__(16) float ybuf[NC][96];                 // y (fp32, one row per clip)
  float zgrad[NC][25];                              // dL/dz from the decoder (fp32)
  float bscale[NC];                                 // 1 / (per-clip power-of-two scale of dL/dy)
  __align__(16) float4 trk[NC][4][32];              // tracker tables, structure of arrays (dp_fk2.cuh)
  __align__(16) float2 st[NC][5][DP_L / 2];         // [ST_Z latent | ST_TL target latent | ST_M, ST_V Adam moments | ST_ZLAST]
  __align__(16) float groot[NC][4];                 // previous world root rotation g (wxyz)
  __align__(16) float2 fkscr[NC / 2][16];           // per clip pair: R_0, r, d parked between the two halves of the kinematics pass
  double prev[NC];                                  // previous total loss (early stopping compares in double)
  float loss[NC][3];                                // last evaluated lp, lr, lt
  int iters[NC];
  uint64_t bar_w, bar_mma[2];
  uint32_t tmem_base;
  __device__ __forceinline__ const DpModelImageTC& M() const { return *reinterpret_cast<const DpModelImageTC*>(model); }
};

// named block barriers of the two groups (ids 1, 2; id 0 is __syncthreads)
__device__ __forceinline__ void group_sync(int gid) { asm volatile("bar.sync %0, 256;" ::"r"(gid + 1) : "memory"); }
__device__ __forceinline__ bool group_or(int gid, bool p) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbar.red.or.pred q, %2, 256, p;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
      : "=r"(r)
      : "r"((uint32_t)p), "r"(gid + 1)
      : "memory");
  return r != 0;
}

__device__ __forceinline__ uint32_t pack_f16x2(float lo_elem, float hi_elem) {  // lower address <- lo_elem
  uint32_t p;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(hi_elem), "f"(lo_elem));
  return p;
}
__device__ __forceinline__ void unpack_f16x2(uint32_t p, float& lo_elem, float& hi_elem) {
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}\n" : "=f"(lo_elem), "=f"(hi_elem) : "r"(p));
}
// 8 fp32 values (feature k, clip 8-group cg of the tile) -> the two fp16 pieces of an MN-major activation image
__device__ __forceinline__ void store_pieces(unsigned char* img, uint32_t piece_stride, int k, int cg, const float (&v)[EC]) {
  unsigned char* dst = img + (k >> 3) * kB_LBO + (k & 7) * 16 + cg * kB_SBO;
  uint32_t p1[4], p2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float x0 = v[2 * i], x1 = v[2 * i + 1];
    p1[i] = pack_f16x2(x0, x1);
    float h0, h1;
    unpack_f16x2(p1[i], h0, h1);
    p2[i] = pack_f16x2(x0 - h0, x1 - h1);
  }
  *reinterpret_cast<uint4*>(dst) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
  *reinterpret_cast<uint4*>(dst + piece_stride) = make_uint4(p2[0], p2[1], p2[2], p2[3]);
}
// a single fp32 value (feature k, clip column n) -> its pieces (used by the Adam lanes for the latent)
__device__ __forceinline__ void store_piece_scalar(unsigned char* img, uint32_t piece_stride, int k, int n, float x) {
  unsigned char* dst = img + (k >> 3) * kB_LBO + (k & 7) * 16 + (n >> 3) * kB_SBO + (n & 7) * 2;
  const uint32_t p = pack_f16x2(x, 0.0f);
  float h0, h1;
  unpack_f16x2(p, h0, h1);
  const uint32_t q = pack_f16x2(x - h0, 0.0f);
  *reinterpret_cast<unsigned short*>(dst) = (unsigned short)(p & 0xffffu);
  *reinterpret_cast<unsigned short*>(dst + piece_stride) = (unsigned short)(q & 0xffffu);
}

// weight pieces in tensor memory (word offsets inside kT_W): layer l forward = W_l (rows = outputs), backward = W_l^T
template <int L, bool FWD> struct WT;
template <> struct WT<0, true>  { static constexpr uint32_t p1 = 0,   p2 = 16,  ksteps = 2; };   // K = 32 (24 latent + pad)
template <> struct WT<1, true>  { static constexpr uint32_t p1 = 32,  p2 = 56,  ksteps = 3; };   // K = 48
template <> struct WT<2, true>  { static constexpr uint32_t p1 = 80,  p2 = 112, ksteps = 4; };   // K = 64
template <> struct WT<2, false> { static constexpr uint32_t p1 = 144, p2 = 192, ksteps = 6; };   // K = 96
template <> struct WT<1, false> { static constexpr uint32_t p1 = 240, p2 = 272, ksteps = 4; };   // K = 64
template <> struct WT<0, false> { static constexpr uint32_t p1 = 304, p2 = 328, ksteps = 3; };   // K = 48
static_assert(WT<0, false>::p2 + 24 == DP_TC_TMEM_WORDS, "tensor-memory weight map");

// sink of the kinematics adjoint: the (scaled) dL/dy of the warp's two clips as fp16 pieces of the K-major operand image
struct EmitDyPieces {
  unsigned char* img;
  int n0, lane;
  static __device__ __forceinline__ void put(unsigned char* dst, float x0, float x1, float x2, float x3) {
    const uint32_t a = pack_f16x2(x0, x1), b = pack_f16x2(x2, x3);
    float h0, h1, h2, h3;
    unpack_f16x2(a, h0, h1);
    unpack_f16x2(b, h2, h3);
    *reinterpret_cast<uint2*>(dst) = make_uint2(a, b);
    *reinterpret_cast<uint2*>(dst + kDyBytes) = make_uint2(pack_f16x2(x0 - h0, x1 - h1), pack_f16x2(x2 - h2, x3 - h3));
  }
  __device__ __forceinline__ void operator()(const P2 (&o)[4], const P2 (&db)[3]) const {
    // clips n0 and n0 + 1 share an 8-group (n0 is even): rows (n0 % 8) and (n0 % 8) + 1 of the same core matrices
    unsigned char* base = img + (n0 >> 3) * kD_SBO + (n0 & 7) * 16;
    if (lane < DP_J) {
      unsigned char* dst = base + (lane >> 1) * kD_LBO + (lane & 1) * 8;  // k = 4 lane: 8-group lane / 2, first or second half
      put(dst, o[0].v.x, o[1].v.x, o[2].v.x, o[3].v.x);
      put(dst + 16, o[0].v.y, o[1].v.y, o[2].v.y, o[3].v.y);
    }
    if (lane == 0) {  // k = 88 .. 91 (91 is the unused pad output: zero)
      unsigned char* dst = base + 11 * kD_LBO;
      put(dst, db[0].v.x, db[1].v.x, db[2].v.x, 0.0f);
      put(dst + 16, db[0].v.y, db[1].v.y, db[2].v.y, 0.0f);
    }
  }
};

struct Ctx {
  SmemT* S;
  uint32_t tmem;
  int gid, wg, lane;
  uint32_t phase;  // parity of the group's next MMA completion
};

// one dense layer of one group on the tensor pipe + its epilogue; every thread of the group calls this (ends with the
// group barrier).  A = weight pieces in tensor memory, B = the group's 16 clip columns of the activation image.
// B_KMAJOR: the B operand is the K-major dL/dy image instead of an MN-major activation image.
template <int L, bool FWD, bool B_KMAJOR = false, class Epi>
__device__ __forceinline__ void tc_layer(Ctx& c, const unsigned char* src, uint32_t src_stride, int out_rows, Epi epi) {
  SmemT& S = *c.S;
  if (c.wg == 0) {
    tc_fence_after();
    if (elect_one()) {
      // fp16 x fp16 -> fp32, N 16, M 128; bit 16: B is MN-major
      constexpr uint32_t idesc = (1u << 4) | (B_KMAJOR ? 0u : (1u << 16)) | ((uint32_t)(16 >> 3) << 17) | (8u << 24);
      constexpr uint32_t lbo = B_KMAJOR ? kD_LBO : kB_LBO, sbo = B_KMAJOR ? kD_SBO : kB_SBO;
      const uint32_t b_base = smem_u32(src) + (uint32_t)c.gid * 2 * sbo;  // the group's 16 clip columns = two 8-groups
      const UmmaDescBase b1 = umma_desc_base(b_base, lbo, sbo), b2 = umma_desc_base(b_base + src_stride, lbo, sbo);
      const uint32_t d = c.tmem + kT_D + 16 * c.gid, a1 = c.tmem + kT_W + WT<L, FWD>::p1, a2 = c.tmem + kT_W + WT<L, FWD>::p2;
#pragma unroll
      for (int k = 0; k < (int)WT<L, FWD>::ksteps; ++k) {  // (2,1) | (1,2) | (1,1), smallest first
        const uint32_t bo = k * 2 * lbo;  // K = 16 per instruction: two 8-groups of K
        if (k == 0) umma_f16_ts_c<false>(d, a2 + 8 * k, umma_desc_at(b1, bo), idesc);
        else umma_f16_ts_c<true>(d, a2 + 8 * k, umma_desc_at(b1, bo), idesc);
        umma_f16_ts_c<true>(d, a1 + 8 * k, umma_desc_at(b2, bo), idesc);
        umma_f16_ts_c<true>(d, a1 + 8 * k, umma_desc_at(b1, bo), idesc);
      }
      umma_commit(&S.bar_mma[c.gid]);
    }
    __syncwarp();
  }
  const int quarter = c.wg & 3, half = c.wg >> 2;
  if (quarter * 32 < out_rows) {
    mbar_wait(&S.bar_mma[c.gid], c.phase);
    tc_fence_after();
    float v[EC];
    tmem_ld8(c.tmem + ((uint32_t)(quarter * 32) << 16) + kT_D + (uint32_t)(16 * c.gid + EC * half), v);
    tmem_ld_wait();
    const int k = quarter * 32 + c.lane;
    if (k < out_rows) epi(k, half, v);
    tc_fence_before();
  }
  fence_proxy_async();
  c.phase ^= 1u;
  group_sync(c.gid);
}

__global__ void __launch_bounds__(kWarps * 32, 1) dp_frame_tc16_kernel(const __grid_constant__ DpFrameArgs A) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemT& S = *reinterpret_cast<SmemT*>(smem_raw);
  const DpModelImageTC& M = S.M();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gid = warp >> 3, wg = warp & 7;  // group, warp within the group
  if (threadIdx.x == 0) {
    mbar_init(&S.bar_w, 1);
    mbar_init(&S.bar_mma[0], 1);
    mbar_init(&S.bar_mma[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&S.tmem_base, kT_COLS);
  static_assert(offsetof(SmemT, dyimg) == offsetof(SmemT, ping) + sizeof(S.ping) + sizeof(S.pong), "operand images are contiguous");
  for (int i = threadIdx.x; i < (int)(sizeof(S.ping) + sizeof(S.pong) + sizeof(S.dyimg)) / 16; i += kWarps * 32)
    reinterpret_cast<uint4*>(&S.ping[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);  // pad rows / columns must stay finite
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&S.bar_w, (uint32_t)sizeof(S.model));
    tma_bulk_g2s(S.model, A.model_tc, (uint32_t)sizeof(S.model), &S.bar_w);
  }
  const uint32_t tmem = S.tmem_base;
  {  // weight pieces -> tensor memory: warp w fills lanes 32 (w % 4) .. +31, columns 88 (w / 4) .. +87 (coalesced reads)
    const uint32_t* img = A.model_tmem;
    const int row = (warp & 3) * 32 + lane, c0 = (warp >> 2) * (DP_TC_TMEM_WORDS / 4);
#pragma unroll 1
    for (int j = 0; j < DP_TC_TMEM_WORDS / 4; j += 8) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(img[(size_t)(c0 + j + i) * 128 + row]);
      tmem_st8(tmem + ((uint32_t)((warp & 3) * 32) << 16) + kT_W + (uint32_t)(c0 + j), v);
    }
    tmem_st_wait();
  }
  Ctx ctx{&S, tmem, gid, wg, lane, 0u};
  constexpr int CPW = 2;            // clips per warp in the per-clip phases
  const int n0 = warp * CPW;        // clip column of this warp's first clip (group g owns columns 16 g .. 16 g + 15)
  const int cpg = (A.clips_per_cta + 1) / 2;  // real clips per group
  const int clip0 = blockIdx.x * A.clips_per_cta + gid * cpg + 2 * wg;

  // ---- per-clip frame inputs (same as the fp32 kernel)
  static_assert(CPW == 2, "the kinematics pass packs exactly two clips per warp");
  bool valid[CPW];
  float inv3e[CPW], lrot9e[CPW];
#pragma unroll
  for (int c = 0; c < CPW; ++c) {
    const int clip = clip0 + c;
    valid[c] = clip < A.n_clips && 2 * wg + c < cpg && gid * cpg + 2 * wg + c < A.clips_per_cta;
    const int cc = valid[c] ? clip : 0;
    int ne = A.n_ee ? A.n_ee[cc] : A.ee_stride;
    ne = max(1, min(ne, A.ee_stride));
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
    // tracker stream of the clip: lane e loads slot e (all slots in flight at once, neighbouring lanes read neighbouring
    // addresses) and scatters its row to the lane of the joint it tracks; untracked joints keep zero weights
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) S.trk[n0 + c][i][lane] = zero4;
    __syncwarp();
    if (lane < ne) {
      float origin[3] = {0.f, 0.f, 0.f};  // world-absolute targets are taken relative to the clip's current global position
      if (A.targets_world) { origin[0] = A.gpos[cc * 3]; origin[1] = A.gpos[cc * 3 + 1]; origin[2] = A.gpos[cc * 3 + 2]; }
      const int j = A.joints[(A.shared_trackers ? 0 : (size_t)cc * A.ee_stride) + lane];
      const float2 wt = reinterpret_cast<const float2*>(A.weights + (A.shared_trackers ? 0 : (size_t)cc * A.ee_stride * 2))[lane];
      const float* tp = A.tgt_pos + ((size_t)cc * A.ee_stride + lane) * 3;
      const float* tr = A.tgt_rot + ((size_t)cc * A.ee_stride + lane) * 9;
      if (j >= 0 && j < DP_J) {
        S.trk[n0 + c][0][j] = make_float4(tp[0] - origin[0], tp[1] - origin[1], tp[2] - origin[2], wt.x);
        S.trk[n0 + c][1][j] = make_float4(tr[0], tr[1], tr[2], wt.y);
        S.trk[n0 + c][2][j] = make_float4(tr[3], tr[4], tr[5], 0.f);
        S.trk[n0 + c][3][j] = make_float4(tr[6], tr[7], tr[8], 0.f);
      }
    }
    if (lane < DP_L / 2) {  // latent -> B operand of the first layer
      store_piece_scalar(&S.ping[0][0], kPingBytes, 2 * lane, n0 + c, z.x);
      store_piece_scalar(&S.ping[0][0], kPingBytes, 2 * lane + 1, n0 + c, z.y);
    }
  }
  const P2 inv3e2 = mk2(inv3e[0], inv3e[1]), lrot9e2 = mk2(lrot9e[0], lrot9e[1]);
  __syncwarp();
  fence_proxy_async();
  mbar_wait(&S.bar_w, 0);  // model image (biases, statistics, skeleton tables) has landed
  tc_fence_before();
  __syncthreads();         // tensor-memory weights written by all warps; latent pieces and tracker rows of both groups in place
  tc_fence_after();

  const FkLaneIdx lane_idx = fk_lane_idx(M, lane);  // skeleton indices of this lane, in registers for the whole launch
  constexpr float wsc = 1.0f / kWScale;  // undoes the weight-image scaling
  const int cg0 = 2 * gid;               // first clip 8-group of this group in the activation images
  unsigned neg0 = 0, neg1 = 0;           // LeakyReLU slope bits of (feature k, this thread's 8 clips) for the backward pass
  auto forward = [&]() {
    tc_layer<0, true>(ctx, &S.ping[0][0], kPingBytes, DP_H0, [&](int k, int half, float (&v)[EC]) {
      const float b = M.b0[k];
      neg0 = 0;
#pragma unroll
      for (int i = 0; i < EC; ++i) { v[i] = fmaf(v[i], wsc, b); neg0 |= (v[i] > 0.f ? 0u : 1u) << i; v[i] = lrelu(v[i]); }
      store_pieces(&S.pong[0][0], kPongBytes, k, cg0 + half, v);
    });
    tc_layer<1, true>(ctx, &S.pong[0][0], kPongBytes, DP_H1, [&](int k, int half, float (&v)[EC]) {
      const float b = M.b1[k];
      neg1 = 0;
#pragma unroll
      for (int i = 0; i < EC; ++i) { v[i] = fmaf(v[i], wsc, b); neg1 |= (v[i] > 0.f ? 0u : 1u) << i; v[i] = lrelu(v[i]); }
      store_pieces(&S.ping[0][0], kPingBytes, k, cg0 + half, v);
    });
    tc_layer<2, true>(ctx, &S.ping[0][0], kPingBytes, DP_Y, [&](int k, int half, float (&v)[EC]) {
      const float b = M.b2[k];
#pragma unroll
      for (int i = 0; i < EC; ++i) S.ybuf[8 * (cg0 + half) + i][k] = fmaf(v[i], wsc, b);
    });
  };
  auto backward = [&]() {  // dL/dy pieces were written by the kinematics warps (EmitDyPieces)
    tc_layer<2, false, true>(ctx, &S.dyimg[0][0], kDyBytes, DP_H1, [&](int k, int half, float (&v)[EC]) {
#pragma unroll
      for (int i = 0; i < EC; ++i) v[i] *= ((neg1 >> i) & 1u) ? 0.2f * wsc : wsc;
      store_pieces(&S.ping[0][0], kPingBytes, k, cg0 + half, v);
    });
    tc_layer<1, false>(ctx, &S.ping[0][0], kPingBytes, DP_H0, [&](int k, int half, float (&v)[EC]) {
#pragma unroll
      for (int i = 0; i < EC; ++i) v[i] *= ((neg0 >> i) & 1u) ? 0.2f * wsc : wsc;
      store_pieces(&S.pong[0][0], kPongBytes, k, cg0 + half, v);
    });
    tc_layer<0, false>(ctx, &S.pong[0][0], kPongBytes, DP_L, [&](int k, int half, float (&v)[EC]) {
#pragma unroll
      for (int i = 0; i < EC; ++i) S.zgrad[8 * (cg0 + half) + i][k] = v[i] * (wsc * S.bscale[8 * (cg0 + half) + i]);
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
    if (!group_or(gid, active[0] || active[1])) break;  // also publishes the latent pieces written by the Adam lanes
    if (lane == 0) {  // the Adam phase reads two table entries: pull their lines into L1 now instead of stalling there
      asm volatile("prefetch.global.L1 [%0];" ::"l"(A.adam_tab + it));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(A.adam_tab + A.max_iter + it));
    }
    phase_done(3);
    forward();
    phase_done(0);
    float nlp[CPW] = {0.f, 0.f}, nlr[CPW] = {0.f, 0.f};
    if (active[0] || active[1]) {  // both clips of the warp in one packed pass; results of a stopped clip are discarded
      // one packed pass; dL/dy leaves it scaled per clip into [16, 32) (exact powers of two, undone when dL/dz is written)
      const FkOut2 o = fk_loss2<true, false, true>(M, lane_idx, &S.ybuf[n0][0], &S.ybuf[n0 + 1][0], &S.trk[n0][0][0], &S.trk[n0 + 1][0][0], &S.groot[n0][0],
                                                   &S.fkscr[warp][0], inv3e2, lrot9e2, lane, nullptr, nullptr, nullptr, nullptr, &S.bscale[n0],
                                                   EmitDyPieces{&S.dyimg[0][0], n0, lane});
      phase_done(4);
      if (active[0]) { nlp[0] = o.lp.v.x; nlr[0] = o.lr.v.x; }
      if (active[1]) { nlp[1] = o.lp.v.y; nlr[1] = o.lr.v.y; }
    }
    fence_proxy_async();  // the dL/dy pieces are read by the tensor core (async proxy)
    group_sync(gid);
    phase_done(1);
    backward();
    phase_done(2);
    const float step_size = A.adam_tab[it], inv_bc2s = A.adam_tab[A.max_iter + it];  // lr/(1-b1^k), 1/sqrt(1-b2^k)
    {  // latent loss, Adam and early-stop bookkeeping of BOTH clips in one pass: clip 0 on lanes 0..11, clip 1 on lanes 16..27
      const int c = lane >> 4, j = lane & 15, n = n0 + c;
      const bool lat = j < DP_L / 2;
      const bool act = c ? active[1] : active[0];
      const float my_lp = c ? nlp[1] : nlp[0], my_lr = c ? nlr[1] : nlr[0];
      float2 z = make_float2(0.f, 0.f), tl = z;
      if (lat) { z = S.st[n][ST_Z][j]; tl = S.st[n][ST_TL][j]; }
      const float dx = z.x - tl.x, dy = z.y - tl.y;
      float ssq = fmaf(dx, dx, dy * dy);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, o);  // sums stay inside each 16-lane half
      const float nlt = ssq * lt_scale;
      float gx = 0.f, gy = 0.f;
      if (lat) {
        gx = fmaf(2.0f * lt_scale, dx, S.zgrad[n][2 * j]);
        gy = fmaf(2.0f * lt_scale, dy, S.zgrad[n][2 * j + 1]);
      }
      if (A.trace && act) {
        float* row = A.trace + ((size_t)(clip0 + c) * A.trace_iters + it) * 52;
        if (lat) {
          reinterpret_cast<float2*>(row)[j] = z;
          reinterpret_cast<float2*>(row + DP_L)[j] = make_float2(gx, gy);
        }
        if (j == 0) { row[48] = my_lp; row[49] = my_lr; row[50] = nlt; row[51] = 1.0f; }
      }
      if (A.eval_only) {
        if (act && lat) reinterpret_cast<float2*>(A.eval_grad + (size_t)(clip0 + c) * DP_L)[j] = make_float2(gx, gy);
      } else if (act && lat) {
        float2 am = S.st[n][ST_M][j], av = S.st[n][ST_V][j];
        am.x = fmaf(0.1f, gx - am.x, am.x);
        am.y = fmaf(0.1f, gy - am.y, am.y);
        av.x = av.x * 0.999f + (0.001f * gx) * gx;
        av.y = av.y * 0.999f + (0.001f * gy) * gy;
        S.st[n][ST_M][j] = am;
        S.st[n][ST_V][j] = av;
        S.st[n][ST_ZLAST][j] = z;  // the frame's output is decoded from the last EVALUATED latent
        z.x += __fdividef(-step_size * am.x, fmaf(fast_sqrt(av.x), inv_bc2s, 1e-8f));
        z.y += __fdividef(-step_size * am.y, fmaf(fast_sqrt(av.y), inv_bc2s, 1e-8f));
        S.st[n][ST_Z][j] = z;
      }
      // the latent rows of the ping image were overwritten by the a1 / dL/dh1 pieces of this iteration: restore them for
      // EVERY clip (stopped and padding clips included) so that no column ever feeds back on its own garbage -- a
      // non-finite value in a K-padding row would poison the column through 0 x NaN
      if (lat) {
        store_piece_scalar(&S.ping[0][0], kPingBytes, 2 * j, n, z.x);
        store_piece_scalar(&S.ping[0][0], kPingBytes, 2 * j + 1, n, z.y);
      }
      const float total = (my_lp + my_lr) + nlt;
      const double incr = S.prev[n] - (double)total;
      __syncwarp();
      if (act && j == 0) {
        S.prev[n] = (double)total;
        S.loss[n][0] = my_lp; S.loss[n][1] = my_lr; S.loss[n][2] = nlt;
        S.iters[n] += 1;
      }
      const bool next = act && ((double)my_lp > A.eps_pos || (double)my_lr > A.eps_rot) && (incr > A.min_incr);
      const unsigned votes = __ballot_sync(0xffffffffu, next);
      active[0] = votes & 1u;
      active[1] = (votes >> 16) & 1u;
    }
    fence_proxy_async();
  }

  // ---- frame epilogue (drag_pose.py:369-414) from the LAST EVALUATED latent (pre-step)
  __syncwarp();
#pragma unroll
  for (int c = 0; c < CPW; ++c)
    if (lane < DP_L / 2) {
      const float2 zl = S.st[n0 + c][ST_ZLAST][lane];
      store_piece_scalar(&S.ping[0][0], kPingBytes, 2 * lane, n0 + c, zl.x);
      store_piece_scalar(&S.ping[0][0], kPingBytes, 2 * lane + 1, n0 + c, zl.y);
    }
  fence_proxy_async();
  group_sync(gid);
  forward();
  P2 q2[4], r2[4], p2[3], d2[3];
  if (valid[0] || valid[1])
    fk_loss2<false, true>(M, lane_idx, &S.ybuf[n0][0], &S.ybuf[n0 + 1][0], &S.trk[n0][0][0], &S.trk[n0 + 1][0][0], &S.groot[n0][0], &S.fkscr[warp][0], inv3e2, lrot9e2, lane, q2, r2, p2, d2);
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
        A.out_gpos[(size_t)clip * A.out_gpos_stride + i] = gp[i];
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
      reinterpret_cast<float4*>(A.out_pose + (size_t)clip * A.out_pose_stride)[lane] =
          make_float4((s[0] - mq.x) / sq.x, (s[1] - mq.y) / sq.y, (s[2] - mq.z) / sq.z, (s[3] - mq.w) / sq.w);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(ctx.tmem, kT_COLS);
}

}  // namespace

cudaError_t dp_frame_tc16_launch(const DpFrameArgs& args, int num_sms, cudaStream_t stream) {
  static std::atomic<unsigned long long> configured{0};
  const size_t smem = sizeof(SmemT) + 1024;
  if (cudaError_t e = dp_ensure_smem(dp_frame_tc16_kernel, smem, configured); e != cudaSuccess) return e;
  // spread the clips over every SM: 4096 clips -> 28 per CTA (two groups of 14) on 147 SMs
  DpFrameArgs a = args;
  int cpc = (args.n_clips + num_sms - 1) / num_sms;
  cpc = cpc < 1 ? 1 : (cpc > NC ? NC : cpc);
  if (args.n_clips > num_sms * NC) cpc = NC;  // several waves anyway: use full tiles
  a.clips_per_cta = cpc;
  dp_frame_tc16_kernel<<<(args.n_clips + cpc - 1) / cpc, kWarps * 32, smem, stream>>>(a);
  return cudaGetLastError();
}
