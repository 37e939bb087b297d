// dp_temporal_attn_tc.cu -- encoder self-attention block of the temporal predictor with its two projections on
// tcgen05 tensor cores.
//
//   out = LayerNorm(x + W_o . MHA(x W_in + b_in) + b_o)
// (torch nn.TransformerEncoderLayer self-attention sub-block, post-norm, d_model 48, 4 heads of 12, 14 tokens per
// clip; python/src/temporal_transformer.py:26-33).  The encoder always sees 14 tokens per clip, so one CTA takes a
// tile of 9 whole clips = 126 token rows (the UMMA M dimension is 128; two rows are padding):
//   MMA1  QKV[128x144] = X[128x48] . W_in^T      accumulator in tensor memory (144 columns)
//   epi   K | V -> shared memory (fp32, + bias); each thread keeps the Q of its token for two heads in registers
//   SIMT  one thread per (token, head pair): 14 scores, softmax, P.V  -- the keys of a clip are broadcast reads
//   A     attention output -> fp16 pieces packed into tensor memory (A operand of the output projection)
//   MMA2  O[128x48]    = A[128x48] . W_o^T       accumulator in tensor memory
//   epi   + b_o + residual, LayerNorm, store
// Same fp16x2 split-product scheme as the feed-forward kernel (dp_temporal_tc.cu): every fp32 operand is two fp16
// pieces, three products accumulate in fp32, the weight image holds 64 W.  Replaces the fp32 CUDA-core attention kernel
// for the encoder layers (57 k tokens at 4096 clips: 118 us -> see profiles/); the decoder keeps the CUDA-core kernel.
#include <cuda_fp16.h>

#include "dp_common.cuh"
#include "dp_internal.h"
#include "dp_temporal.cuh"
#include "dp_umma.cuh"

namespace {

constexpr int kTM = 128;
constexpr int kClips = ATT_TILE_CLIPS;          // 9 clips per tile
constexpr int kRows = kClips * TP_S;            // 126 token rows
constexpr float kWScale = 64.0f;                // weight image holds 64 W
constexpr uint32_t kWinBytes = 3 * TP_D * TP_D * 2;  // one fp16 image of W_in as B operand [N = 144][K = 48]
constexpr uint32_t kWoBytes = TP_D * TP_D * 2;       // one fp16 image of W_o  as B operand [N = 48][K = 48]
constexpr uint32_t kOffWin = 0, kOffWo = 2 * kWinBytes, kOffBin = kOffWo + 2 * kWoBytes, kOffBo = kOffBin + 3 * TP_D * 4;
static_assert(kOffBo + TP_D * 4 == ATT_LAYER_BYTES, "attention weight image size");
// K-major no-swizzle fp16 B operand: element (n,k) at (n/8)*128 + (k/8)*LBO + (n%8)*16 + (k%8)*2
constexpr uint32_t kWin_LBO = 128 * (3 * TP_D / 8), kWo_LBO = 128 * (TP_D / 8), kSBO = 128;
constexpr int kKvStride = 100;  // floats per K|V row: 14 rows apart is 24 banks apart, so the <= 4 clips of a warp never collide

struct Smem {
  __align__(16) unsigned char w[ATT_LAYER_BYTES];
  __align__(16) float kv[kRows][kKvStride];  // K (48) | V (48) | pad
  uint64_t bar_w, bar_mma;
  uint32_t tmem_base;
};
// tensor-memory columns (32-bit words per lane; lane = token row)
constexpr uint32_t kT_A1 = 0, kT_A2 = 24;  // the two fp16 pieces of X, later of the attention output (two K elements per word)
constexpr uint32_t kT_QKV = 48;            // 144 fp32 columns; the first 48 are reused for the output projection
constexpr uint32_t kT_COLS = 256;
constexpr uint32_t kIdescF16 = (1u << 4);  // fp32 accumulate, fp16 A/B, both K-major

template <int N>
__device__ __forceinline__ void issue_proj(uint32_t tmem, uint32_t d_col, uint32_t w_smem, uint32_t piece_bytes, uint32_t lbo) {
  constexpr uint32_t idesc = kIdescF16 | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTM >> 4) << 24);
  const UmmaDescBase w1 = umma_desc_base(w_smem, lbo, kSBO), w2 = umma_desc_base(w_smem + piece_bytes, lbo, kSBO);
  const uint32_t d = tmem + d_col, a1 = tmem + kT_A1, a2 = tmem + kT_A2;
#pragma unroll
  for (int k = 0; k < TP_D / 16; ++k) {
    const uint32_t bo = k * 2 * lbo;
    if (k == 0) umma_f16_ts_c<false>(d, a2 + 8 * k, umma_desc_at(w1, bo), idesc);
    else umma_f16_ts_c<true>(d, a2 + 8 * k, umma_desc_at(w1, bo), idesc);
    umma_f16_ts_c<true>(d, a1 + 8 * k, umma_desc_at(w2, bo), idesc);
    umma_f16_ts_c<true>(d, a1 + 8 * k, umma_desc_at(w1, bo), idesc);
  }
}

constexpr int kWorkThreads = 256, kThreads = kWorkThreads + 32;  // 8 worker warps + 1 MMA/TMA issuer warp

__global__ void __launch_bounds__(kThreads, 2)
tp_attn_tc_kernel(const unsigned char* __restrict__ wimg, const float* __restrict__ blob, TpNorm N1, const float* __restrict__ x_g,
                  int n_clips, float* __restrict__ out_g) {
  extern __shared__ __align__(1024) unsigned char raw[];
  Smem& S = *reinterpret_cast<Smem*>(raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int clip0 = blockIdx.x * kClips;
  if (tid == 0) {
    mbar_init(&S.bar_w, 1);
    mbar_init(&S.bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(&S.tmem_base, kT_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;
  if (warp == 8 && elect_one()) {  // the whole weight image of this layer: two bulk copies
    constexpr uint32_t kHalf = 2 * kWinBytes;
    mbar_expect_tx(&S.bar_w, ATT_LAYER_BYTES);
    tma_bulk_g2s(S.w, wimg, kHalf, &S.bar_w);
    tma_bulk_g2s(S.w + kHalf, wimg + kHalf, ATT_LAYER_BYTES - kHalf, &S.bar_w);
  }
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const int m = (warp & 3) * 32 + lane;  // token row == TMEM lane owned by this thread
  const int h = (warp >> 2) & 1;         // worker warps 4..7 take the second head pair / the V half
  const bool valid = m < kRows && clip0 + m / TP_S < n_clips;
  const size_t g = ((size_t)clip0 * TP_S + m) * TP_D;
  if (warp < 4) {  // X tile -> tensor memory as the A operand
    float p1[24], p2[24];
#pragma unroll
    for (int j = 0; j < TP_D; j += 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) v = *reinterpret_cast<const float4*>(x_g + g + j);
      split_h2(v.x, v.y, p1[j / 2], p2[j / 2]);
      split_h2(v.z, v.w, p1[j / 2 + 1], p2[j / 2 + 1]);
    }
    tmem_st8(tmem + lane_base + kT_A1, reinterpret_cast<float (&)[8]>(p1[0]));
    tmem_st8(tmem + lane_base + kT_A1 + 8, reinterpret_cast<float (&)[8]>(p1[8]));
    tmem_st8(tmem + lane_base + kT_A1 + 16, reinterpret_cast<float (&)[8]>(p1[16]));
    tmem_st8(tmem + lane_base + kT_A2, reinterpret_cast<float (&)[8]>(p2[0]));
    tmem_st8(tmem + lane_base + kT_A2 + 8, reinterpret_cast<float (&)[8]>(p2[8]));
    tmem_st8(tmem + lane_base + kT_A2 + 16, reinterpret_cast<float (&)[8]>(p2[16]));
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    mbar_wait(&S.bar_w, 0);
    if (elect_one()) {
      issue_proj<3 * TP_D>(tmem, kT_QKV, smem_u32(S.w) + kOffWin, kWinBytes, kWin_LBO);
      umma_commit(&S.bar_mma);
    }
    __syncwarp();
  }
  float q[24];  // Q of this token for heads 2h, 2h+1, already scaled by 1/sqrt(head_dim)
  if (warp < 8) {
    mbar_wait(&S.bar_mma, 0);
    tc_fence_after();
    mbar_wait(&S.bar_w, 0);  // acquire the TMA-written bias vectors
    const float* b_in = reinterpret_cast<const float*>(S.w + kOffBin);
    // K (h = 0) or V (h = 1) of this token -> shared memory
#pragma unroll
    for (int j0 = 0; j0 < TP_D; j0 += 16) {
      float v[16];
      tmem_ld16(tmem + lane_base + kT_QKV + (uint32_t)(TP_D + TP_D * h + j0), v);
      tmem_ld_wait();
      if (m < kRows) {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(&S.kv[m][TP_D * h + j0 + j]) =
              make_float4(fmaf(v[j], 1.0f / kWScale, b_in[TP_D + TP_D * h + j0 + j]), fmaf(v[j + 1], 1.0f / kWScale, b_in[TP_D + TP_D * h + j0 + j + 1]),
                          fmaf(v[j + 2], 1.0f / kWScale, b_in[TP_D + TP_D * h + j0 + j + 2]), fmaf(v[j + 3], 1.0f / kWScale, b_in[TP_D + TP_D * h + j0 + j + 3]));
      }
    }
    const float qs = rsqrtf((float)TP_HD);
#pragma unroll
    for (int j0 = 0; j0 < 24; j0 += 8) {
      float v[8];
      tmem_ld8(tmem + lane_base + kT_QKV + (uint32_t)(24 * h + j0), v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) q[j0 + j] = fmaf(v[j], 1.0f / kWScale, b_in[24 * h + j0 + j]) * qs;
    }
    tc_fence_before();
  }
  __syncthreads();  // K | V of the tile visible; the QKV accumulator columns are free again
  if (warp < 8) {
    float o[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) o[i] = 0.0f;
    if (valid) {
      const float* kv0 = &S.kv[(m / TP_S) * TP_S][0];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int col = (2 * h + hh) * TP_HD;
        float s[TP_S], mx = -3.0e38f;
#pragma unroll
        for (int j = 0; j < TP_S; ++j) {
          const float4* kr = reinterpret_cast<const float4*>(kv0 + j * kKvStride + col);
          const float4 k0 = kr[0], k1 = kr[1], k2 = kr[2];
          const float* qq = q + TP_HD * hh;
          float a = qq[0] * k0.x;
          a = fmaf(qq[1], k0.y, a); a = fmaf(qq[2], k0.z, a); a = fmaf(qq[3], k0.w, a);
          a = fmaf(qq[4], k1.x, a); a = fmaf(qq[5], k1.y, a); a = fmaf(qq[6], k1.z, a); a = fmaf(qq[7], k1.w, a);
          a = fmaf(qq[8], k2.x, a); a = fmaf(qq[9], k2.y, a); a = fmaf(qq[10], k2.z, a); a = fmaf(qq[11], k2.w, a);
          s[j] = a;
          mx = fmaxf(mx, a);
        }
        float sum = 0.0f;
#pragma unroll
        for (int j = 0; j < TP_S; ++j) { s[j] = expf(s[j] - mx); sum += s[j]; }
        const float inv = 1.0f / sum;
        float* oo = o + TP_HD * hh;
#pragma unroll
        for (int j = 0; j < TP_S; ++j) {
          const float4* vr = reinterpret_cast<const float4*>(kv0 + j * kKvStride + TP_D + col);
          const float4 v0 = vr[0], v1 = vr[1], v2 = vr[2];
          const float pj = s[j] * inv;
          oo[0] = fmaf(pj, v0.x, oo[0]); oo[1] = fmaf(pj, v0.y, oo[1]); oo[2] = fmaf(pj, v0.z, oo[2]); oo[3] = fmaf(pj, v0.w, oo[3]);
          oo[4] = fmaf(pj, v1.x, oo[4]); oo[5] = fmaf(pj, v1.y, oo[5]); oo[6] = fmaf(pj, v1.z, oo[6]); oo[7] = fmaf(pj, v1.w, oo[7]);
          oo[8] = fmaf(pj, v2.x, oo[8]); oo[9] = fmaf(pj, v2.y, oo[9]); oo[10] = fmaf(pj, v2.z, oo[10]); oo[11] = fmaf(pj, v2.w, oo[11]);
        }
      }
    }
    // attention output (features 24h .. 24h+23 of this token) -> fp16 pieces, words 12h .. 12h+11 of each piece
    float p1[12], p2[12];
#pragma unroll
    for (int j = 0; j < 24; j += 2) split_h2(o[j], o[j + 1], p1[j / 2], p2[j / 2]);
    tc_fence_after();
    tmem_st8(tmem + lane_base + kT_A1 + (uint32_t)(12 * h), reinterpret_cast<float (&)[8]>(p1[0]));
    tmem_st4(tmem + lane_base + kT_A1 + (uint32_t)(12 * h + 8), reinterpret_cast<float (&)[4]>(p1[8]));
    tmem_st8(tmem + lane_base + kT_A2 + (uint32_t)(12 * h), reinterpret_cast<float (&)[8]>(p2[0]));
    tmem_st4(tmem + lane_base + kT_A2 + (uint32_t)(12 * h + 8), reinterpret_cast<float (&)[4]>(p2[8]));
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    if (elect_one()) {
      issue_proj<TP_D>(tmem, kT_QKV, smem_u32(S.w) + kOffWo, kWoBytes, kWo_LBO);
      umma_commit(&S.bar_mma);
    }
    __syncwarp();
  }
  if (warp < 4) {
    mbar_wait(&S.bar_mma, 1);
    tc_fence_after();
    float o[TP_D];
#pragma unroll
    for (int j0 = 0; j0 < TP_D; j0 += 16) {
      float v[16];
      tmem_ld16(tmem + lane_base + kT_QKV + (uint32_t)j0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j0 + j] = v[j];
    }
    if (valid) {
      const float* b_o = reinterpret_cast<const float*>(S.w + kOffBo);
#pragma unroll
      for (int j = 0; j < TP_D; j += 4) {
        const float4 xv = *reinterpret_cast<const float4*>(x_g + g + j);
        o[j] = fmaf(o[j], 1.0f / kWScale, b_o[j] + xv.x); o[j + 1] = fmaf(o[j + 1], 1.0f / kWScale, b_o[j + 1] + xv.y);
        o[j + 2] = fmaf(o[j + 2], 1.0f / kWScale, b_o[j + 2] + xv.z); o[j + 3] = fmaf(o[j + 3], 1.0f / kWScale, b_o[j + 3] + xv.w);
      }
      ln48(o, blob + N1.w, blob + N1.b);
#pragma unroll
      for (int j = 0; j < TP_D; j += 4) *reinterpret_cast<float4*>(out_g + g + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, kT_COLS);
}

// ---------------------------------------------------------------------------------------------------------------------
// K | V projections of the encoder memory for the cross-attention of ALL decoder layers, computed once per predictor call
// (the reference, and the CUDA-core kernel, redo them in every decoder layer of every autoregressive pass):
//   KV_l[rows x 96] = memory[rows x 48] . [W_k | W_v]_l^T + b     l = 0..2, one 128-row tile per CTA, 27 MMAs, one commit.
constexpr uint32_t kKvWBytes = 2 * TP_D * TP_D * 2;   // one fp16 image of [W_k | W_v] as B operand [N = 96][K = 48]
constexpr uint32_t kKv_LBO = 128 * (2 * TP_D / 8);
constexpr uint32_t kKvOffBias = TP_NDEC * 2 * kKvWBytes;
static_assert(kKvOffBias + TP_NDEC * 2 * TP_D * 4 == KV_IMAGE_BYTES, "kv image size");
constexpr uint32_t kT_KV = 48, kT_KV_COLS = 512;      // three 96-column accumulators after the X pieces

struct SmemKv {
  __align__(16) unsigned char w[KV_IMAGE_BYTES];
  uint64_t bar_w, bar_mma;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kThreads, 1)
tp_kv_tc_kernel(const unsigned char* __restrict__ wimg, const float* __restrict__ x_g, int n_rows, float* __restrict__ kv_g) {
  extern __shared__ __align__(1024) unsigned char raw[];
  SmemKv& S = *reinterpret_cast<SmemKv*>(raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&S.bar_w, 1);
    mbar_init(&S.bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(&S.tmem_base, kT_KV_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;
  if (warp == 8 && elect_one()) {
    constexpr uint32_t kPart = 2 * kKvWBytes;  // one layer's two pieces
    mbar_expect_tx(&S.bar_w, KV_IMAGE_BYTES);
    for (int l = 0; l < TP_NDEC; ++l) tma_bulk_g2s(S.w + l * kPart, wimg + l * kPart, kPart, &S.bar_w);
    tma_bulk_g2s(S.w + kKvOffBias, wimg + kKvOffBias, KV_IMAGE_BYTES - kKvOffBias, &S.bar_w);
  }
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const int m = (warp & 3) * 32 + lane, h = (warp >> 2) & 1;
  const int row = blockIdx.x * kTM + m;
  const bool valid = row < n_rows;
  if (warp < 4) {
    float p1[24], p2[24];
#pragma unroll
    for (int j = 0; j < TP_D; j += 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) v = *reinterpret_cast<const float4*>(x_g + (size_t)row * TP_D + j);
      split_h2(v.x, v.y, p1[j / 2], p2[j / 2]);
      split_h2(v.z, v.w, p1[j / 2 + 1], p2[j / 2 + 1]);
    }
    tmem_st8(tmem + lane_base + kT_A1, reinterpret_cast<float (&)[8]>(p1[0]));
    tmem_st8(tmem + lane_base + kT_A1 + 8, reinterpret_cast<float (&)[8]>(p1[8]));
    tmem_st8(tmem + lane_base + kT_A1 + 16, reinterpret_cast<float (&)[8]>(p1[16]));
    tmem_st8(tmem + lane_base + kT_A2, reinterpret_cast<float (&)[8]>(p2[0]));
    tmem_st8(tmem + lane_base + kT_A2 + 8, reinterpret_cast<float (&)[8]>(p2[8]));
    tmem_st8(tmem + lane_base + kT_A2 + 16, reinterpret_cast<float (&)[8]>(p2[16]));
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    mbar_wait(&S.bar_w, 0);
    if (elect_one()) {
      for (int l = 0; l < TP_NDEC; ++l)
        issue_proj<2 * TP_D>(tmem, kT_KV + 2 * TP_D * l, smem_u32(S.w) + l * 2 * kKvWBytes, kKvWBytes, kKv_LBO);
      umma_commit(&S.bar_mma);
    }
    __syncwarp();
  } else {
    mbar_wait(&S.bar_mma, 0);
    tc_fence_after();
    mbar_wait(&S.bar_w, 0);
    const float* bias = reinterpret_cast<const float*>(S.w + kKvOffBias);
#pragma unroll 1
    for (int l = 0; l < TP_NDEC; ++l) {
      float* dst = kv_g + ((size_t)l * n_rows + row) * (2 * TP_D) + TP_D * h;
#pragma unroll
      for (int j0 = 0; j0 < TP_D; j0 += 16) {
        float v[16];
        tmem_ld16(tmem + lane_base + kT_KV + (uint32_t)(2 * TP_D * l + TP_D * h + j0), v);
        tmem_ld_wait();
        if (valid) {
          const float* b = bias + 2 * TP_D * l + TP_D * h + j0;
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(dst + j0 + j) = make_float4(fmaf(v[j], 1.0f / kWScale, b[j]), fmaf(v[j + 1], 1.0f / kWScale, b[j + 1]),
                                                                  fmaf(v[j + 2], 1.0f / kWScale, b[j + 2]), fmaf(v[j + 3], 1.0f / kWScale, b[j + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, kT_KV_COLS);
}

}  // namespace

// Host: build the pre-split, pre-tiled weight image of one self-attention block (ATT_LAYER_BYTES).
// w_in_t is [48][144] (in, out = q|k|v), w_out_t is [48][48] (in, out) -- the transposed layout of the blob.
void dp_attn_tc_pack(const float* w_in_t, const float* b_in, const float* w_out_t, const float* b_out, unsigned char* dst) {
  auto put = [&](unsigned char* base, uint32_t piece_bytes, uint32_t lbo, int N, const float* wt) {
    __half* pc[2] = {reinterpret_cast<__half*>(base), reinterpret_cast<__half*>(base + piece_bytes)};
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < TP_D; ++k) {
        float r = kWScale * wt[(size_t)k * N + n];
        const uint32_t off = ((n >> 3) * kSBO + (k >> 3) * lbo + (n & 7) * 16 + (k & 7) * 2) / 2;
        for (int p = 0; p < 2; ++p) {
          pc[p][off] = __float2half_rn(r);
          r -= __half2float(pc[p][off]);
        }
      }
  };
  put(dst + kOffWin, kWinBytes, kWin_LBO, 3 * TP_D, w_in_t);
  put(dst + kOffWo, kWoBytes, kWo_LBO, TP_D, w_out_t);
  memcpy(dst + kOffBin, b_in, 3 * TP_D * sizeof(float));
  memcpy(dst + kOffBo, b_out, TP_D * sizeof(float));
}

cudaError_t dp_attn_tc_launch(const unsigned char* wimg, const float* blob, const TpNorm& N1, const float* x, int n_clips, float* out,
                              cudaStream_t st) {
  static bool configured = false;
  const size_t smem = sizeof(Smem) + 1024;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tp_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  tp_attn_tc_kernel<<<(n_clips + kClips - 1) / kClips, kThreads, smem, st>>>(wimg, blob, N1, x, n_clips, out);
  return cudaGetLastError();
}

// Host: [W_k | W_v] images + biases of the TP_NDEC cross-attention blocks (KV_IMAGE_BYTES). w_in_t[l] is that block's
// [48][144] transposed in-projection, b_in[l] its [144] bias.
void dp_kv_tc_pack(const float* const* w_in_t, const float* const* b_in, unsigned char* dst) {
  for (int l = 0; l < TP_NDEC; ++l) {
    __half* pc[2] = {reinterpret_cast<__half*>(dst + l * 2 * kKvWBytes), reinterpret_cast<__half*>(dst + l * 2 * kKvWBytes + kKvWBytes)};
    for (int n = 0; n < 2 * TP_D; ++n)
      for (int k = 0; k < TP_D; ++k) {
        float r = kWScale * w_in_t[l][(size_t)k * 3 * TP_D + TP_D + n];
        const uint32_t off = ((n >> 3) * kSBO + (k >> 3) * kKv_LBO + (n & 7) * 16 + (k & 7) * 2) / 2;
        for (int p = 0; p < 2; ++p) {
          pc[p][off] = __float2half_rn(r);
          r -= __half2float(pc[p][off]);
        }
      }
    memcpy(dst + kKvOffBias + l * 2 * TP_D * sizeof(float), b_in[l] + TP_D, 2 * TP_D * sizeof(float));
  }
}

cudaError_t dp_kv_tc_launch(const unsigned char* wimg, const float* x, int n_rows, float* kv, cudaStream_t st) {
  static bool configured = false;
  const size_t smem = sizeof(SmemKv) + 1024;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tp_kv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  tp_kv_tc_kernel<<<(n_rows + kTM - 1) / kTM, kThreads, smem, st>>>(wimg, x, n_rows, kv);
  return cudaGetLastError();
}
