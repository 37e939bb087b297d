// dp_temporal_tc.cu -- feed-forward block of the temporal predictor on tcgen05 tensor cores.
//
//   out = LayerNorm(x + W2 relu(W1 x + b1) + b2)   [+ optional second LayerNorm]
// (torch nn.TransformerEncoderLayer / DecoderLayer FF sub-block, post-norm, d_model 48,
// dim_feedforward 2048; python/src/temporal_transformer.py:26-33).  These two GEMMs are 95%
// of the predictor's FLOPs.  One CTA owns a tile of 128 tokens (the UMMA M dimension) and
// walks the 2048 hidden units in chunks of 64:
//   MMA1  H[128x64]  = X[128x48]  . W1c^T      (tf32, K = 48)       accumulator in TMEM
//   epi   H -> +b1, relu, hi/lo split -> shared memory (UMMA A-operand layout)
//   MMA2  O[128x48] += H[128x64]  . W2c^T      (tf32, K = 64)       accumulator in TMEM
// Precision: 3xTF32 error compensation (x = hi + lo; lo.hi + hi.lo + hi.hi, fp32 accumulate)
// -- measured 2.7e-7 relative on B200, the same as an fp32 GEMM; plain TF32 gives 7.7e-4.
// Weights are pre-split and pre-tiled on the host into the exact shared-memory image, so a
// chunk is ONE contiguous bulk-TMA copy (cp.async.bulk); chunks are double buffered and the
// MMA of chunk c+1 is issued behind MMA2 of chunk c so the tensor pipe stays busy while the
// 8 epilogue warps convert the next H tile.
#include <cstdlib>
#include <cstring>

#include "dp_internal.h"
#include "dp_umma.cuh"

namespace {

constexpr int kTM = 128;
constexpr int kHC = FFT_HC;                    // 64 hidden units per chunk
constexpr int kChunks = TP_FF / kHC;           // 32
constexpr uint32_t kW1Bytes = kHC * TP_D * 4;  // 12288: one tf32 image of W1c [64][48]
constexpr uint32_t kW2Bytes = TP_D * kHC * 4;  // 12288: one tf32 image of W2c [48][64]
constexpr uint32_t kPart1 = 2 * kW1Bytes + kHC * 4;  // W1 hi | W1 lo | b1 chunk
constexpr uint32_t kPart2 = 2 * kW2Bytes;            // W2 hi | W2 lo
static_assert(kPart1 + kPart2 == FFT_CHUNK_BYTES, "chunk size");
// shared-memory operand geometry (bytes), see dp_umma.cuh
constexpr uint32_t kW1_LBO = 128 * (kHC / 8), kW2_LBO = 128 * (TP_D / 8), kB_SBO = 128;

constexpr int kW1Stages = 4, kW2Stages = 4;
struct Smem {
  unsigned char w1[kW1Stages][kPart1];
  unsigned char w2[kW2Stages][kPart2];
  uint64_t w1full[kW1Stages], w2full[kW2Stages], hfull[2], hready[2];
  uint32_t tmem_base;
};
// tensor-memory columns (fp32 words per lane; lane = token row)
constexpr uint32_t kT_XHI = 0, kT_XLO = 48, kT_H0 = 96, kT_H1 = 160, kT_L0 = 224, kT_L1 = 288, kT_OUT = 352, kT_COLS = 512;

// H(c)[128x64] = X . W1c^T : A = X from tensor memory (hi | lo), B = W1c from shared memory
__device__ __forceinline__ void issue_mma1(const Smem& S, int stage, uint32_t tmem, uint32_t d_col) {
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kHC >> 3) << 17) | ((uint32_t)(kTM >> 4) << 24);
  const UmmaDescBase wh = umma_desc_base(smem_u32(S.w1[stage]), kW1_LBO, kB_SBO);
  const UmmaDescBase wl = umma_desc_base(smem_u32(S.w1[stage]) + kW1Bytes, kW1_LBO, kB_SBO);
  const uint32_t d = tmem + d_col, xh = tmem + kT_XHI, xl = tmem + kT_XLO;
#pragma unroll
  for (int k = 0; k < TP_D / 8; ++k) {
    const uint32_t bo = k * 2 * kW1_LBO;
    if (k == 0) umma_tf32_ts_c<false>(d, xl + 8 * k, umma_desc_at(wh, bo), idesc);
    else umma_tf32_ts_c<true>(d, xl + 8 * k, umma_desc_at(wh, bo), idesc);
    umma_tf32_ts_c<true>(d, xh + 8 * k, umma_desc_at(wl, bo), idesc);
    umma_tf32_ts_c<true>(d, xh + 8 * k, umma_desc_at(wh, bo), idesc);
  }
}
// O[128x48] += H(c) . W2c^T : A = relu(H) hi (in place of the accumulator) | lo from tensor memory
template <bool FIRST>
__device__ __forceinline__ void issue_mma2(const Smem& S, int stage, uint32_t tmem, uint32_t hi_col, uint32_t lo_col) {
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TP_D >> 3) << 17) | ((uint32_t)(kTM >> 4) << 24);
  const UmmaDescBase wh = umma_desc_base(smem_u32(S.w2[stage]), kW2_LBO, kB_SBO);
  const UmmaDescBase wl = umma_desc_base(smem_u32(S.w2[stage]) + kW2Bytes, kW2_LBO, kB_SBO);
  const uint32_t d = tmem + kT_OUT, hh = tmem + hi_col, hl = tmem + lo_col;
#pragma unroll
  for (int k = 0; k < kHC / 8; ++k) {
    const uint32_t bo = k * 2 * kW2_LBO;
    if (FIRST && k == 0) umma_tf32_ts_c<false>(d, hl + 8 * k, umma_desc_at(wh, bo), idesc);
    else umma_tf32_ts_c<true>(d, hl + 8 * k, umma_desc_at(wh, bo), idesc);
    umma_tf32_ts_c<true>(d, hh + 8 * k, umma_desc_at(wl, bo), idesc);
    umma_tf32_ts_c<true>(d, hh + 8 * k, umma_desc_at(wh, bo), idesc);
  }
}

__device__ __forceinline__ void ln48(float (&v)[TP_D], const float* __restrict__ w, const float* __restrict__ b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < TP_D; ++i) s += v[i];
  const float mean = s * (1.0f / TP_D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < TP_D; ++i) { v[i] -= mean; q = fmaf(v[i], v[i], q); }
  const float rstd = rsqrtf(q * (1.0f / TP_D) + 1e-5f);
#pragma unroll
  for (int i = 0; i < TP_D; ++i) v[i] = v[i] * rstd * w[i] + b[i];
}

__device__ __forceinline__ void load_w1(Smem& S, const unsigned char* wtiles, int c) {
  const int st = c % kW1Stages;
  mbar_expect_tx(&S.w1full[st], kPart1);
  tma_bulk_g2s(S.w1[st], wtiles + (size_t)c * FFT_CHUNK_BYTES, kPart1, &S.w1full[st]);
}
__device__ __forceinline__ void load_w2(Smem& S, const unsigned char* wtiles, int c) {
  const int st = c % kW2Stages;
  mbar_expect_tx(&S.w2full[st], kPart2);
  tma_bulk_g2s(S.w2[st], wtiles + (size_t)c * FFT_CHUNK_BYTES + kPart1, kPart2, &S.w2full[st]);
}

constexpr int kEpiThreads = 256, kThreads = kEpiThreads + 32;  // 8 epilogue warps + 1 MMA/TMA issuer warp

__global__ void __launch_bounds__(kThreads, 1)
tp_ff_tc_kernel(const unsigned char* __restrict__ wtiles, const float* __restrict__ blob, TpFF F, TpNorm N1, TpNorm N2, int has_n2,
                const float* __restrict__ x_g, int n_rows, int T, int row_stride, float* __restrict__ out_g, int dbg) {
  extern __shared__ __align__(1024) unsigned char raw[];
  Smem& S = *reinterpret_cast<Smem*>(raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * kTM;
  if (tid == 0) {
    for (int i = 0; i < kW1Stages; ++i) mbar_init(&S.w1full[i], 1);
    for (int i = 0; i < kW2Stages; ++i) mbar_init(&S.w2full[i], 1);
    mbar_init(&S.hfull[0], 1); mbar_init(&S.hfull[1], 1);
    mbar_init(&S.hready[0], kEpiThreads); mbar_init(&S.hready[1], kEpiThreads);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(&S.tmem_base, kT_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;
  if (warp == 8 && elect_one()) {  // weight pipeline prologue (bulk TMA): W1 chunks 0..3, W2 chunks 0..3
    for (int c = 0; c < kW1Stages; ++c) load_w1(S, wtiles, c);
    for (int c = 0; c < kW2Stages; ++c) load_w2(S, wtiles, c);
  }
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const int m = (warp & 3) * 32 + lane;          // token row == TMEM lane owned by this thread
  const int chalf = warp >> 2;                   // which 32 of a chunk's 64 hidden columns
  const int row = row0 + m;
  const size_t g = row < n_rows ? ((size_t)(row / T) * row_stride + row % T) * TP_D : 0;
  // X tile -> tensor memory as the A operand (hi | lo); warps 0-3 take columns 0..23, warps 4-7 columns 24..47
  if (warp < 8) {
    float h[24], l[24];
#pragma unroll
    for (int j = 0; j < 24; j += 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < n_rows) v = *reinterpret_cast<const float4*>(x_g + g + chalf * 24 + j);
      split_tf32(v.x, h[j], l[j]); split_tf32(v.y, h[j + 1], l[j + 1]);
      split_tf32(v.z, h[j + 2], l[j + 2]); split_tf32(v.w, h[j + 3], l[j + 3]);
    }
    tmem_st8(tmem + lane_base + kT_XHI + chalf * 24, h + 0); tmem_st8(tmem + lane_base + kT_XHI + chalf * 24 + 8, h + 8);
    tmem_st8(tmem + lane_base + kT_XHI + chalf * 24 + 16, h + 16);
    tmem_st8(tmem + lane_base + kT_XLO + chalf * 24, l + 0); tmem_st8(tmem + lane_base + kT_XLO + chalf * 24 + 8, l + 8);
    tmem_st8(tmem + lane_base + kT_XLO + chalf * 24 + 16, l + 16);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    // ===== MMA + TMA issuer warp: warp-uniform control flow, one elected lane issues (tcgen05.mma is a
    // single-thread instruction); decoupled from the epilogue warps through mbarriers only
    tc_fence_after();
    mbar_wait(&S.w1full[0], 0);
    mbar_wait(&S.w1full[1], 0);
    if (elect_one()) {
      issue_mma1(S, 0, tmem, kT_H0);
      umma_commit(&S.hfull[0]);
      issue_mma1(S, 1, tmem, kT_H1);
      umma_commit(&S.hfull[1]);
    }
    __syncwarp();
    for (int c = 0; c < kChunks; ++c) {
      const int b = c & 1;
      const uint32_t hcol = b ? kT_H1 : kT_H0, lcol = b ? kT_L1 : kT_L0;
      mbar_wait(&S.hready[b], (c >> 1) & 1);  // all epilogue threads converted H(c) (and finished chunk c-1)
      tc_fence_after();
      if (elect_one()) {
        // refill two chunks ahead of use: stage of chunk c (MMA1(c) retired, b1(c) consumed by the epilogue) and stage of
        // chunk c-2 (MMA2(c-2) retired -- the epilogue could only signal after seeing the commit that covered it)
        if (c + kW1Stages < kChunks) load_w1(S, wtiles, c + kW1Stages);
        if (c >= 2 && c - 2 + kW2Stages < kChunks) load_w2(S, wtiles, c - 2 + kW2Stages);
      }
      __syncwarp();
      mbar_wait(&S.w2full[c % kW2Stages], (c / kW2Stages) & 1);
      if (c + 2 < kChunks) mbar_wait(&S.w1full[(c + 2) % kW1Stages], ((c + 2) / kW1Stages) & 1);
      if (elect_one()) {
        if (!(dbg & 2)) {
          if (c == 0) issue_mma2<true>(S, c % kW2Stages, tmem, hcol, lcol);
          else issue_mma2<false>(S, c % kW2Stages, tmem, hcol, lcol);
        }
        if (c + 2 < kChunks && !(dbg & 1)) issue_mma1(S, (c + 2) % kW1Stages, tmem, hcol);
        umma_commit(&S.hfull[b]);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue warps: H(c) accumulator -> relu(H + b1) -> hi (in place) | lo, all inside tensor memory
    for (int c = 0; c < kChunks; ++c) {
      const int b = c & 1;
      const uint32_t hcol = b ? kT_H1 : kT_H0, lcol = b ? kT_L1 : kT_L0;
      mbar_wait(&S.hfull[b], (c >> 1) & 1);  // H(c) accumulated; MMA2(c-2) has released this buffer pair
      tc_fence_after();
      mbar_wait(&S.w1full[c % kW1Stages], (c / kW1Stages) & 1);  // acquire the TMA-written b1 slice
      const float* b1 = reinterpret_cast<const float*>(S.w1[c % kW1Stages] + 2 * kW1Bytes) + chalf * 32;
      if (!(dbg & 4)) {  // one 32-column load, one wait, all the math, two 32-column stores, one wait
        float v[32], lo[32];
        tmem_ld32(tmem + lane_base + hcol + (uint32_t)(chalf * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) split_tf32(fmaxf(v[j] + b1[j], 0.f), v[j], lo[j]);
        tmem_st32(tmem + lane_base + hcol + (uint32_t)(chalf * 32), v);
        tmem_st32(tmem + lane_base + lcol + (uint32_t)(chalf * 32), lo);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&S.hready[b]);
    }
  }
  // the last two commits cover MMA2(30) and MMA2(31)
  mbar_wait(&S.hfull[0], (kChunks >> 1) & 1);
  mbar_wait(&S.hfull[1], (kChunks >> 1) & 1);
  tc_fence_after();
  if (warp < 4) {
    float o[TP_D];
#pragma unroll
    for (int j0 = 0; j0 < TP_D; j0 += 16) {
      float v[16];
      tmem_ld16(tmem + lane_base + kT_OUT + (uint32_t)j0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j0 + j] = v[j];
    }
    if (row < n_rows) {
      const float* b2 = blob + F.b2;
#pragma unroll
      for (int j = 0; j < TP_D; j += 4) {
        const float4 xv = *reinterpret_cast<const float4*>(x_g + g + j);
        o[j] += b2[j] + xv.x; o[j + 1] += b2[j + 1] + xv.y; o[j + 2] += b2[j + 2] + xv.z; o[j + 3] += b2[j + 3] + xv.w;
      }
      ln48(o, blob + N1.w, blob + N1.b);
      if (has_n2) ln48(o, blob + N2.w, blob + N2.b);
#pragma unroll
      for (int j = 0; j < TP_D; j += 4) *reinterpret_cast<float4*>(out_g + g + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, kT_COLS);
}

}  // namespace

// Host: build the pre-split, pre-tiled weight image of one FF block (kChunks x FFT_CHUNK_BYTES).
// w1t is [48][2048] (in, out), w2t is [2048][48] (in, out) -- the transposed layout of the blob.
void dp_ff_tc_pack(const float* w1t, const float* b1, const float* w2t, unsigned char* dst) {
  auto tf32_hi = [](float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;  // round-to-nearest (ties away) to 10 explicit mantissa bits == cvt.rna.tf32
    float r;
    memcpy(&r, &u, 4);
    return r;
  };
  for (int c = 0; c < kChunks; ++c) {
    unsigned char* base = dst + (size_t)c * FFT_CHUNK_BYTES;
    float* w1hi = reinterpret_cast<float*>(base);
    float* w1lo = reinterpret_cast<float*>(base + kW1Bytes);
    float* bb = reinterpret_cast<float*>(base + 2 * kW1Bytes);
    float* w2hi = reinterpret_cast<float*>(base + kPart1);
    float* w2lo = reinterpret_cast<float*>(base + kPart1 + kW2Bytes);
    for (int n = 0; n < kHC; ++n) {      // W1c as B operand [N = hidden][K = 48]
      bb[n] = b1[c * kHC + n];
      for (int k = 0; k < TP_D; ++k) {
        const float w = w1t[(size_t)k * TP_FF + c * kHC + n];
        const uint32_t off = ((n >> 3) * kB_SBO + (k >> 2) * kW1_LBO + (n & 7) * 16 + (k & 3) * 4) / 4;
        const float h = tf32_hi(w);
        w1hi[off] = h;
        w1lo[off] = w - h;
      }
    }
    for (int n = 0; n < TP_D; ++n)       // W2c as B operand [N = out 48][K = hidden chunk]
      for (int k = 0; k < kHC; ++k) {
        const float w = w2t[(size_t)(c * kHC + k) * TP_D + n];
        const uint32_t off = ((n >> 3) * kB_SBO + (k >> 2) * kW2_LBO + (n & 7) * 16 + (k & 3) * 4) / 4;
        const float h = tf32_hi(w);
        w2hi[off] = h;
        w2lo[off] = w - h;
      }
  }
}

cudaError_t dp_ff_tc_launch(const unsigned char* wtiles, const float* blob, const TpFF& F, const TpNorm& N1, const TpNorm& N2,
                            int has_n2, const float* x, int n_rows, int T, int row_stride, float* out, cudaStream_t st) {
  static bool configured = false;
  const size_t smem = sizeof(Smem) + 1024;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tp_ff_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  static int dbg = getenv("DP_FF_DBG") ? atoi(getenv("DP_FF_DBG")) : 0;
  tp_ff_tc_kernel<<<(n_rows + kTM - 1) / kTM, kThreads, smem, st>>>(wtiles, blob, F, N1, N2, has_n2, x, n_rows, T, row_stride, out, dbg);
  return cudaGetLastError();
}
