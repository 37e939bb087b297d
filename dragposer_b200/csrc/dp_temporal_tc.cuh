#pragma once
// dp_temporal_tc.cuh -- feed-forward block of the temporal predictor on tcgen05 tensor cores.
//
//   out = LayerNorm(x + W2 relu(W1 x + b1) + b2)   [+ optional second LayerNorm]
// (torch nn.TransformerEncoderLayer / DecoderLayer FF sub-block, post-norm, d_model 48,
// dim_feedforward 2048; python/src/temporal_transformer.py:26-33).  These two GEMMs are 95%
// of the predictor's FLOPs.  One CTA owns a tile of 128 tokens (the UMMA M dimension) and
// walks the 2048 hidden units in chunks of 64:
//   MMA1  H[128x64]  = X[128x48]  . W1c^T      (K = 48)       accumulator in TMEM
//   epi   H -> +b1, relu, split -> TENSOR MEMORY (A operand of MMA2, two K elements per column)
//   MMA2  O[128x48] += H[128x64]  . W2c^T      (K = 64)       accumulator in TMEM
// Precision: fp16x2 split products on kind::f16 -- every fp32 value is two fp16 pieces (22 mantissa bits),
// three products (2,1)(1,2)(1,1) accumulate in fp32; the weight image stores 64 W (exact power of two, undone in
// the epilogues) so that the second piece of the small FF weights stays a normal fp16.  Measured on B200: same
// predictor output error as the fp32 CUDA-core kernel to ~1e-6; plain TF32 would be 7.7e-4 per GEMM.
// (History: 3xTF32 on kind::tf32 needed twice the MMAs and twice the weight bytes for the same accuracy.)
// A operands live in tensor memory (TS mode: ~N/2 cycles per MMA instead of the ~45-cycle shared-memory A fetch):
// X is split once into TMEM, the 8 epilogue warps read the H accumulator with one tcgen05.ld.x32 and write the
// packed pieces back with tcgen05.st, so activations never touch shared memory.  Weights are pre-split and
// pre-tiled on the host into the exact shared-memory operand image, so a chunk is ONE contiguous bulk-TMA copy
// (cp.async.bulk) through a 4+4 stage mbarrier pipeline; a dedicated issuer warp (elect.sync lane) feeds the tensor
// pipe -- MMA2(c) then MMA1(c+2) per commit -- while the epilogue warps convert chunk c+1.
#include "dp_common.cuh"
#include "dp_internal.h"
#include "dp_umma.cuh"

namespace tpf {


constexpr int kTM = 128;
constexpr int kHC = FFT_HC;                    // 64 hidden units per chunk
constexpr int kChunks = TP_FF / kHC;           // 32
constexpr int kLag = 2;                        // step t carries W2 of chunk t - kLag
constexpr int kCtasPerSm = 2;
constexpr float kFfWScale = 64.0f;             // weight image holds 64 W
constexpr uint32_t kW1Bytes = kHC * TP_D * 2;  // 6144: one fp16 image of W1c [64][48]
constexpr uint32_t kW2Bytes = TP_D * kHC * 2;  // 6144: one fp16 image of W2c [48][64]
// Weight image of a layer: b1 (2048 fp32) followed by kChunks + 2 "steps"; step t = [W2(t-2) pieces | W1(t) pieces], exactly
// what the issuer needs in iteration t-2 (MMA2 of chunk t-2, MMA1 of chunk t): one bulk copy and one barrier per iteration.
constexpr uint32_t kStepBytes = 2 * kW2Bytes + 2 * kW1Bytes;
static_assert(kStepBytes == FFT_STEP_BYTES && TP_FF * 4 + (kChunks + 2) * kStepBytes == FFT_LAYER_BYTES, "image size");
// shared-memory B-operand geometry (bytes), K-major no-swizzle fp16: element (n,k) at (n/8)*128 + (k/8)*LBO + (n%8)*16 + (k%8)*2
constexpr uint32_t kW1_LBO = 128 * (kHC / 8), kW2_LBO = 128 * (TP_D / 8), kB_SBO = 128;

constexpr int kStages = 4;
struct Smem {
  unsigned char w[kStages][kStepBytes];
  float b1[TP_FF];
  uint64_t wfull[kStages], wfree[kStages], hfull[2], hready[2], b1full;
  uint32_t tmem_base;
};
// tensor-memory columns (32-bit words per lane; lane = token row; fp16 operands hold two K elements per word)
constexpr uint32_t kT_X1 = 0, kT_X2 = 24, kT_H0 = 48, kT_H1 = 112, kT_OUT = 176, kT_COLS = 256;
// The packed pieces of relu(H) overwrite the accumulator they came from: the epilogue thread that owns hidden columns
// [32h, 32h+32) of a chunk writes piece 1 to words [32h, 32h+16) and piece 2 to [32h+16, 32h+32) of the same buffer.
// 256 columns per CTA -> two CTAs share an SM (and its tensor pipe), one converting while the other multiplies.
constexpr uint32_t kIdescF16 = (1u << 4);  // fp32 accumulate, fp16 A/B, both K-major; N and M added below

// H(c)[128x64] = X . W1c^T : A = X pieces from tensor memory, B = W1c pieces from shared memory
__device__ __forceinline__ void issue_mma1(const Smem& S, int stage, uint32_t tmem, uint32_t d_col) {
  constexpr uint32_t idesc = kIdescF16 | ((uint32_t)(kHC >> 3) << 17) | ((uint32_t)(kTM >> 4) << 24);
  const uint32_t base = smem_u32(S.w[stage]) + 2 * kW2Bytes;
  const UmmaDescBase w1 = umma_desc_base(base, kW1_LBO, kB_SBO);
  const UmmaDescBase w2 = umma_desc_base(base + kW1Bytes, kW1_LBO, kB_SBO);
  const uint32_t d = tmem + d_col, x1 = tmem + kT_X1, x2 = tmem + kT_X2;
#pragma unroll
  for (int k = 0; k < TP_D / 16; ++k) {
    const uint32_t bo = k * 2 * kW1_LBO;
    if (k == 0) umma_f16_ts_c<false>(d, x2 + 8 * k, umma_desc_at(w1, bo), idesc);
    else umma_f16_ts_c<true>(d, x2 + 8 * k, umma_desc_at(w1, bo), idesc);
    umma_f16_ts_c<true>(d, x1 + 8 * k, umma_desc_at(w2, bo), idesc);
    umma_f16_ts_c<true>(d, x1 + 8 * k, umma_desc_at(w1, bo), idesc);
  }
}
// O[128x48] += relu(H(c)) . W2c^T : A = pieces of relu(H) from tensor memory
__device__ __forceinline__ void issue_mma2(const Smem& S, int stage, uint32_t tmem, uint32_t p_col, bool first) {
  constexpr uint32_t idesc = kIdescF16 | ((uint32_t)(TP_D >> 3) << 17) | ((uint32_t)(kTM >> 4) << 24);
  const UmmaDescBase w1 = umma_desc_base(smem_u32(S.w[stage]), kW2_LBO, kB_SBO);
  const UmmaDescBase w2 = umma_desc_base(smem_u32(S.w[stage]) + kW2Bytes, kW2_LBO, kB_SBO);
  const uint32_t d = tmem + kT_OUT;
#pragma unroll
  for (int k = 0; k < kHC / 16; ++k) {
    const uint32_t bo = k * 2 * kW2_LBO;
    const uint32_t h1 = tmem + p_col + (k >> 1) * 32 + (k & 1) * 8, h2 = h1 + 16;
    if (k == 0) {
      if (first) umma_f16_ts_c<false>(d, h2, umma_desc_at(w1, bo), idesc);
      else umma_f16_ts_c<true>(d, h2, umma_desc_at(w1, bo), idesc);
    } else {
      umma_f16_ts_c<true>(d, h2, umma_desc_at(w1, bo), idesc);
    }
    umma_f16_ts_c<true>(d, h1, umma_desc_at(w2, bo), idesc);
    umma_f16_ts_c<true>(d, h1, umma_desc_at(w1, bo), idesc);
  }
}

constexpr int kEpiThreads = 256, kThreads = kEpiThreads + 64;  // 8 epilogue warps + MMA issuer warp (8) + TMA producer warp (9)

// One feed-forward block for the 128-row tile starting at row0 (rows >= m_limit of the tile are padding: the fused encoder kernel
// works on tiles of nine whole clips = 126 rows).  (split, n_split): this call takes the hidden chunks [split, split + 1) * kChunks /
// n_split; with n_split > 1 the raw partial sums go to `part` ([n_split][n_rows][48]) and tp_ff_finish_kernel adds bias, residual
// and the LayerNorms (few-row launches: a 4096-row decoder step would otherwise occupy 32 SMs for 32 serial chunks).
// Called by all kThreads threads with kT_COLS columns of tensor memory at `tmem` and the mbarriers of S freshly initialised and
// published by a block barrier; returns after a block barrier.  x_g is read with ld.global.cg (see dp_temporal_attn_tc.cuh).
// bulk copy of local step j of a CTA's chunk range into ring stage st.  The first two steps are only read for their W1 half (MMA1 of
// chunks 0 and 1) and the last two only for their W2 half: the unused halves stay in L2 -- 24 KB less in the fill burst with which every
// CTA of a wave starts at the same moment (profiles/r2_ff_pipeline_clock.md), 48 KB of 816 KB less per tile.
__device__ __forceinline__ void load_step(Smem& S, const unsigned char* steps, int c0, int j, int n_loc, int st) {
  const unsigned char* src = steps + (size_t)(c0 + j) * kStepBytes;
  uint32_t off = 0, bytes = kStepBytes;
  if (j < 2) { off = 2 * kW2Bytes; bytes = 2 * kW1Bytes; }
  else if (j >= n_loc) bytes = 2 * kW2Bytes;
  mbar_expect_tx(&S.wfull[st], bytes);
  tma_bulk_g2s(S.w[st] + off, src + off, bytes, &S.wfull[st]);
}
__device__ __forceinline__ void ff_init_barriers(Smem& S) {
  for (int i = 0; i < kStages; ++i) { mbar_init(&S.wfull[i], 1); mbar_init(&S.wfree[i], 1); }
  mbar_init(&S.hfull[0], 1); mbar_init(&S.hfull[1], 1);
  mbar_init(&S.hready[0], kEpiThreads); mbar_init(&S.hready[1], kEpiThreads);
  mbar_init(&S.b1full, 1);
}
__device__ __forceinline__ void ff_inval_barriers(Smem& S) {
  for (int i = 0; i < kStages; ++i) { mbar_inval(&S.wfull[i]); mbar_inval(&S.wfree[i]); }
  mbar_inval(&S.hfull[0]); mbar_inval(&S.hfull[1]);
  mbar_inval(&S.hready[0]); mbar_inval(&S.hready[1]);
  mbar_inval(&S.b1full);
}
__device__ __forceinline__ void ff_tile(Smem& S, const uint32_t tmem, const unsigned char* wimg, const float* blob, TpFF F, TpNorm N1, TpNorm N2, int has_n2,
                                        const float* x_g, int n_rows, int T, int row_stride, int row0, int m_limit, int split, int n_split,
                                        float* out_g, float* part, long long* trace) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_loc = kChunks / n_split, c0 = split * n_loc;  // this call's chunks: c0 .. c0 + n_loc - 1
  const unsigned char* steps = wimg + TP_FF * 4;
  const bool trace0 = trace && blockIdx.x == (unsigned)trace[3] && blockIdx.y == 0;  // trace[3]: the CTA to clock (DP_FF_TRACE_CTA)
  if (warp == 9 && elect_one()) {  // TMA producer, part 1 (no waits before the block barrier below): bias slice + first stages
    mbar_expect_tx(&S.b1full, (uint32_t)(n_loc * kHC * 4));
    tma_bulk_g2s(S.b1 + c0 * kHC, wimg + (size_t)c0 * kHC * 4, (uint32_t)(n_loc * kHC * 4), &S.b1full);
    for (int j = 0; j < kStages && j < n_loc + 2; ++j) {  // local step j == image step c0 + j
      load_step(S, steps, c0, j, n_loc, j);
    }
  }
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const int m = (warp & 3) * 32 + lane;          // token row == TMEM lane owned by this thread
  const int chalf = (warp >> 2) & 1;             // which 32 of a chunk's 64 hidden columns
  const int row = row0 + m;
  const bool row_ok = row < n_rows && m < m_limit;
  const size_t g = row_ok ? ((size_t)(row / T) * row_stride + row % T) * TP_D : 0;
  // X tile -> tensor memory as the A operand (two fp16 pieces, two K elements per word); warps 0-3 own the 128 rows
  if (warp < 4) {
    float p1[24], p2[24];
#pragma unroll
    for (int j = 0; j < TP_D; j += 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_ok) v = __ldcg(reinterpret_cast<const float4*>(x_g + g + j));
      split_h2(v.x, v.y, p1[j / 2], p2[j / 2]);
      split_h2(v.z, v.w, p1[j / 2 + 1], p2[j / 2 + 1]);
    }
    tmem_st8(tmem + lane_base + kT_X1, reinterpret_cast<float (&)[8]>(p1[0]));
    tmem_st8(tmem + lane_base + kT_X1 + 8, reinterpret_cast<float (&)[8]>(p1[8]));
    tmem_st8(tmem + lane_base + kT_X1 + 16, reinterpret_cast<float (&)[8]>(p1[16]));
    tmem_st8(tmem + lane_base + kT_X2, reinterpret_cast<float (&)[8]>(p2[0]));
    tmem_st8(tmem + lane_base + kT_X2 + 8, reinterpret_cast<float (&)[8]>(p2[8]));
    tmem_st8(tmem + lane_base + kT_X2 + 16, reinterpret_cast<float (&)[8]>(p2[16]));
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    // ===== MMA issuer warp: warp-uniform control flow, one elected lane issues (tcgen05.mma is a single-thread
    // instruction); it only waits on barriers and feeds the tensor pipe -- iteration i: MMA2(i), MMA1(i+2), two commits
    tc_fence_after();
    for (int j = 0; j < 2 && j < n_loc; ++j) {  // prologue: H(0), H(1) from the W1 halves of steps 0, 1
      mbar_wait(&S.wfull[j], 0);
      if (elect_one()) {
        issue_mma1(S, j, tmem, j ? kT_H1 : kT_H0);
        umma_commit(&S.hfull[j]);
        umma_commit(&S.wfree[j]);
      }
      __syncwarp();
    }
    for (int i = 0; i < n_loc; ++i) {
      const int b = i & 1, st = (i + 2) % kStages;
      const uint32_t hcol = b ? kT_H1 : kT_H0;
      mbar_wait(&S.wfull[st], ((i + 2) / kStages) & 1);  // usually long complete
      const bool traced = trace0 && lane == 0;
      if (traced) trace[i * 8 + 0] = clock64();
      mbar_wait(&S.hready[b], (i >> 1) & 1);             // all epilogue threads converted H(i) into its pieces
      tc_fence_after();
      if (traced) trace[i * 8 + 1] = clock64();
      if (elect_one()) {
        issue_mma2(S, st, tmem, hcol, i == 0);
        if (i + 2 < n_loc) issue_mma1(S, st, tmem, hcol);
        umma_commit(&S.hfull[b]);
        umma_commit(&S.wfree[st]);
      }
      __syncwarp();
      if (traced) trace[i * 8 + 2] = clock64();
    }
  } else if (warp == 9) {
    // ===== TMA producer warp, part 2: one step per iteration through the kStages ring; a stage is free again when the
    // MMAs that read it have retired (tcgen05.commit by the issuer)
    if (elect_one()) {
      for (int j = kStages; j < n_loc + 2; ++j) {
        const int st = j % kStages;
        mbar_wait(&S.wfree[st], ((j / kStages) - 1) & 1);
        load_step(S, steps, c0, j, n_loc, st);
      }
    }
    __syncwarp();
  } else if (warp < 8) {
    // ===== epilogue warps: H(i) accumulator -> relu(H / 64 + b1) -> two packed fp16 pieces, all inside tensor memory
    mbar_wait(&S.b1full, 0);
    for (int i = 0; i < n_loc; ++i) {
      const int b = i & 1;
      const uint32_t hcol = b ? kT_H1 : kT_H0;
      const float* b1 = S.b1 + (c0 + i) * kHC + chalf * 32;
      mbar_wait(&S.hfull[b], (i >> 1) & 1);  // H(i) accumulated; MMA2(i-2) has released this buffer
      tc_fence_after();
      const bool traced = trace0 && tid == 0;
      if (traced) trace[i * 8 + 4] = clock64();
      float v[32], p1[16], p2[16];
      tmem_ld32(tmem + lane_base + hcol + (uint32_t)(chalf * 32), v);
      tmem_ld_wait();
      if (traced) trace[i * 8 + 5] = clock64();
      // Three CUDA-core instructions per hidden unit instead of five (the epilogue warps co-limit this kernel with the tensor pipe):
      // packed fp32x2 arithmetic (FFMA2 / FADD2, sm_100) for scale + bias and for the residual of the split, and the ReLU folded into
      // the conversions.  Piece 1 is relu(h) rounded TOWARDS ZERO, so the residual h - p1 of a positive h is never negative, while a
      // negative h (p1 = 0) leaves a negative residual that the second relu-conversion clamps to zero: p1 + p2 = relu(h) to 2^-21.
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float2 h = __ffma2_rn(make_float2(v[j], v[j + 1]), make_float2(1.0f / kFfWScale, 1.0f / kFfWScale), make_float2(b1[j], b1[j + 1]));
        uint32_t w, w2;
        asm("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(h.y), "f"(h.x));
        float2 f;
        asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}\n" : "=f"(f.x), "=f"(f.y) : "r"(w));
        const float2 d = __fadd2_rn(h, make_float2(-f.x, -f.y));
        asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(w2) : "f"(d.y), "f"(d.x));
        p1[j / 2] = __uint_as_float(w);
        p2[j / 2] = __uint_as_float(w2);
      }
      if (traced && i == 0) trace[11] = clock64();  // chunk 0: pieces computed, before the tensor-memory stores
      tmem_st16(tmem + lane_base + hcol + (uint32_t)(chalf * 32), p1);
      tmem_st16(tmem + lane_base + hcol + (uint32_t)(chalf * 32) + 16u, p2);
      tmem_st_wait();
      if (traced) trace[i * 8 + 6] = clock64();
      tc_fence_before();
      mbar_arrive(&S.hready[b]);
    }
  }
  if (warp < 4) {
    // the last two commits cover MMA2(n_loc - 2) and MMA2(n_loc - 1)
    if (n_loc >= 2) mbar_wait(&S.hfull[n_loc & 1], (n_loc >> 1) & 1);
    mbar_wait(&S.hfull[(n_loc - 1) & 1], (((n_loc - 1) >> 1) + 1) & 1);
    tc_fence_after();
    float o[TP_D];
#pragma unroll
    for (int j0 = 0; j0 < TP_D; j0 += 16) {
      float v[16];
      tmem_ld16(tmem + lane_base + kT_OUT + (uint32_t)j0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j0 + j] = v[j] * (1.0f / kFfWScale);
    }
    if (row_ok) {
      if (n_split > 1) {
        float* dst = part + ((size_t)split * n_rows + row) * TP_D;
#pragma unroll
        for (int j = 0; j < TP_D; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
      } else {
        const float* b2 = blob + F.b2;
#pragma unroll
        for (int j = 0; j < TP_D; j += 4) {
          const float4 xv = __ldcg(reinterpret_cast<const float4*>(x_g + g + j));
          o[j] += b2[j] + xv.x; o[j + 1] += b2[j + 1] + xv.y; o[j + 2] += b2[j + 2] + xv.z; o[j + 3] += b2[j + 3] + xv.w;
        }
        ln48(o, blob + N1.w, blob + N1.b);
        if (has_n2) ln48(o, blob + N2.w, blob + N2.b);
#pragma unroll
        for (int j = 0; j < TP_D; j += 4) *reinterpret_cast<float4*>(out_g + g + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (trace0 && tid == 0) trace[15] = clock64();
}


}  // namespace tpf
