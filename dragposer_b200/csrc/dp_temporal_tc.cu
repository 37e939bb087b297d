// dp_temporal_tc.cu -- feed-forward block of the temporal predictor on tcgen05 tensor cores.
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
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "dp_internal.h"
#include "dp_umma.cuh"

namespace {

constexpr int kTM = 128;
constexpr int kHC = FFT_HC;                    // 64 hidden units per chunk
constexpr int kChunks = TP_FF / kHC;           // 32
constexpr float kFfWScale = 64.0f;             // weight image holds 64 W
constexpr uint32_t kW1Bytes = kHC * TP_D * 2;  // 6144: one fp16 image of W1c [64][48]
constexpr uint32_t kW2Bytes = TP_D * kHC * 2;  // 6144: one fp16 image of W2c [48][64]
constexpr uint32_t kPart1 = 2 * kW1Bytes + kHC * 4;  // W1 piece 1 | W1 piece 2 | b1 chunk (fp32)
constexpr uint32_t kPart2 = 2 * kW2Bytes;            // W2 piece 1 | W2 piece 2
static_assert(kPart1 + kPart2 == FFT_CHUNK_BYTES, "chunk size");
// shared-memory B-operand geometry (bytes), K-major no-swizzle fp16: element (n,k) at (n/8)*128 + (k/8)*LBO + (n%8)*16 + (k%8)*2
constexpr uint32_t kW1_LBO = 128 * (kHC / 8), kW2_LBO = 128 * (TP_D / 8), kB_SBO = 128;

constexpr int kW1Stages = 4, kW2Stages = 4;
struct Smem {
  unsigned char w1[kW1Stages][kPart1];
  unsigned char w2[kW2Stages][kPart2];
  uint64_t w1full[kW1Stages], w2full[kW2Stages], hfull[2], hready[2];
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
  const UmmaDescBase w1 = umma_desc_base(smem_u32(S.w1[stage]), kW1_LBO, kB_SBO);
  const UmmaDescBase w2 = umma_desc_base(smem_u32(S.w1[stage]) + kW1Bytes, kW1_LBO, kB_SBO);
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
template <bool FIRST>
__device__ __forceinline__ void issue_mma2(const Smem& S, int stage, uint32_t tmem, uint32_t p_col) {
  constexpr uint32_t idesc = kIdescF16 | ((uint32_t)(TP_D >> 3) << 17) | ((uint32_t)(kTM >> 4) << 24);
  const UmmaDescBase w1 = umma_desc_base(smem_u32(S.w2[stage]), kW2_LBO, kB_SBO);
  const UmmaDescBase w2 = umma_desc_base(smem_u32(S.w2[stage]) + kW2Bytes, kW2_LBO, kB_SBO);
  const uint32_t d = tmem + kT_OUT;
#pragma unroll
  for (int k = 0; k < kHC / 16; ++k) {
    const uint32_t bo = k * 2 * kW2_LBO;
    const uint32_t h1 = tmem + p_col + (k >> 1) * 32 + (k & 1) * 8, h2 = h1 + 16;
    if (FIRST && k == 0) umma_f16_ts_c<false>(d, h2, umma_desc_at(w1, bo), idesc);
    else umma_f16_ts_c<true>(d, h2, umma_desc_at(w1, bo), idesc);
    umma_f16_ts_c<true>(d, h1, umma_desc_at(w2, bo), idesc);
    umma_f16_ts_c<true>(d, h1, umma_desc_at(w1, bo), idesc);
  }
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

__global__ void __launch_bounds__(kThreads, 2)
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
  // X tile -> tensor memory as the A operand (two fp16 pieces, two K elements per word); warps 0-3 own the 128 rows
  if (warp < 4) {
    float p1[24], p2[24];
#pragma unroll
    for (int j = 0; j < TP_D; j += 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < n_rows) v = *reinterpret_cast<const float4*>(x_g + g + j);
      split_h2(v.x, v.y, p1[j / 2], p2[j / 2]);
      split_h2(v.z, v.w, p1[j / 2 + 1], p2[j / 2 + 1]);
    }
    tmem_st8(tmem + lane_base + kT_X1, p1 + 0); tmem_st8(tmem + lane_base + kT_X1 + 8, p1 + 8); tmem_st8(tmem + lane_base + kT_X1 + 16, p1 + 16);
    tmem_st8(tmem + lane_base + kT_X2, p2 + 0); tmem_st8(tmem + lane_base + kT_X2 + 8, p2 + 8); tmem_st8(tmem + lane_base + kT_X2 + 16, p2 + 16);
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
      const uint32_t hcol = b ? kT_H1 : kT_H0, pcol = hcol;
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
          if (c == 0) issue_mma2<true>(S, c % kW2Stages, tmem, pcol);
          else issue_mma2<false>(S, c % kW2Stages, tmem, pcol);
        }
        if (c + 2 < kChunks && !(dbg & 1)) issue_mma1(S, (c + 2) % kW1Stages, tmem, hcol);
        umma_commit(&S.hfull[b]);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue warps: H(c) accumulator -> relu(H / 64 + b1) -> two packed fp16 pieces, all inside tensor memory
    for (int c = 0; c < kChunks; ++c) {
      const int b = c & 1;
      const uint32_t hcol = b ? kT_H1 : kT_H0, pcol = hcol;
      mbar_wait(&S.hfull[b], (c >> 1) & 1);  // H(c) accumulated; MMA2(c-2) has released this buffer pair
      tc_fence_after();
      mbar_wait(&S.w1full[c % kW1Stages], (c / kW1Stages) & 1);  // acquire the TMA-written b1 slice
      const float* b1 = reinterpret_cast<const float*>(S.w1[c % kW1Stages] + 2 * kW1Bytes) + chalf * 32;
      if (!(dbg & 4)) {  // one 32-column load, bias + relu, split into packed fp16 pieces, two 16-word stores
        float v[32], p1[16], p2[16];
        tmem_ld32(tmem + lane_base + hcol + (uint32_t)(chalf * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 2)
          split_h2(fmaxf(fmaf(v[j], 1.0f / kFfWScale, b1[j]), 0.f), fmaxf(fmaf(v[j + 1], 1.0f / kFfWScale, b1[j + 1]), 0.f), p1[j / 2], p2[j / 2]);
        tmem_st16(tmem + lane_base + pcol + (uint32_t)(chalf * 32), p1);
        tmem_st16(tmem + lane_base + pcol + (uint32_t)(chalf * 32) + 16u, p2);
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
        o[j] = fmaf(o[j], 1.0f / kFfWScale, b2[j] + xv.x); o[j + 1] = fmaf(o[j + 1], 1.0f / kFfWScale, b2[j + 1] + xv.y);
        o[j + 2] = fmaf(o[j + 2], 1.0f / kFfWScale, b2[j + 2] + xv.z); o[j + 3] = fmaf(o[j + 3], 1.0f / kFfWScale, b2[j + 3] + xv.w);
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
  for (int c = 0; c < kChunks; ++c) {
    unsigned char* base = dst + (size_t)c * FFT_CHUNK_BYTES;
    __half* w1p[2] = {reinterpret_cast<__half*>(base), reinterpret_cast<__half*>(base + kW1Bytes)};
    float* bb = reinterpret_cast<float*>(base + 2 * kW1Bytes);
    __half* w2p[2] = {reinterpret_cast<__half*>(base + kPart1), reinterpret_cast<__half*>(base + kPart1 + kW2Bytes)};
    for (int n = 0; n < kHC; ++n) {      // W1c as B operand [N = hidden][K = 48]
      bb[n] = b1[c * kHC + n];
      for (int k = 0; k < TP_D; ++k) {
        float r = kFfWScale * w1t[(size_t)k * TP_FF + c * kHC + n];
        const uint32_t off = ((n >> 3) * kB_SBO + (k >> 3) * kW1_LBO + (n & 7) * 16 + (k & 7) * 2) / 2;
        for (int p = 0; p < 2; ++p) {
          w1p[p][off] = __float2half_rn(r);
          r -= __half2float(w1p[p][off]);
        }
      }
    }
    for (int n = 0; n < TP_D; ++n)       // W2c as B operand [N = out 48][K = hidden chunk]
      for (int k = 0; k < kHC; ++k) {
        float r = kFfWScale * w2t[(size_t)(c * kHC + k) * TP_D + n];
        const uint32_t off = ((n >> 3) * kB_SBO + (k >> 3) * kW2_LBO + (n & 7) * 16 + (k & 7) * 2) / 2;
        for (int p = 0; p < 2; ++p) {
          w2p[p][off] = __float2half_rn(r);
          r -= __half2float(w2p[p][off]);
        }
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
