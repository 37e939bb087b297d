#pragma once
// dp_temporal_attn_tc.cuh -- attention blocks of the temporal predictor with their projections on tcgen05 tensor cores.
//
//   out = LayerNorm(x_q + W_o . MHA(x_q W_q + b_q, x_kv W_k + b_k, x_kv W_v + b_v) + b_o)
// (torch nn.TransformerEncoderLayer / DecoderLayer attention sub-blocks, post-norm, d_model 48, 4 heads of 12;
// python/src/temporal_transformer.py:26-33).  One kernel serves the encoder self-attention (14 tokens per clip), the
// decoder self-attention (T = 1..30 tokens) and the decoder cross-attention (T queries against the 14 memory tokens).
// A CTA takes a tile of G whole clips, G = 128 / max(T, S) (the UMMA M dimension is 128 rows):
//   MMA   self : QKV[128x144] = X[128x48] . W_in^T                          one accumulator, 144 columns of tensor memory
//         cross: Q[128x48] = X_q . W_q^T  and  K|V[128x96] = X_kv . [W_k|W_v]^T   (same columns, two A tiles)
//   epi   K | V -> shared memory (fp32, + bias); each thread keeps the Q of its row for two heads in registers
//   SIMT  one thread per (query row, head pair): scores against the clip's S keys, softmax, P.V -- the keys of a clip are
//         broadcast reads
//   A     attention output -> fp16 pieces packed into tensor memory (A operand of the output projection)
//   MMA   O[128x48] = A[128x48] . W_o^T
//   epi   + b_o + residual, LayerNorm, store
// Same fp16x2 split-product scheme as the feed-forward kernel (dp_temporal_tc.cu): every fp32 operand is two fp16
// pieces, three products accumulate in fp32, the weight image holds 64 W.  The fp32 CUDA-core kernel (dp_temporal.cu)
// stays as the on-device cross-check (predictor path 1).
#include "dp_common.cuh"
#include "dp_internal.h"
#include "dp_temporal.cuh"
#include "dp_umma.cuh"

namespace tpa {


constexpr int kTM = 128;
constexpr float kWScale = 64.0f;                // weight image holds 64 W
constexpr uint32_t kWinBytes = 3 * TP_D * TP_D * 2;  // one fp16 image of W_in as B operand [N = 144][K = 48]
constexpr uint32_t kWoBytes = TP_D * TP_D * 2;       // one fp16 image of W_o  as B operand [N = 48][K = 48]
constexpr uint32_t kOffWin = 0, kOffWo = 2 * kWinBytes, kOffBin = kOffWo + 2 * kWoBytes, kOffBo = kOffBin + 3 * TP_D * 4;
static_assert(kOffBo + TP_D * 4 == ATT_LAYER_BYTES, "attention weight image size");
// K-major no-swizzle fp16 B operand: element (n,k) at (n/8)*128 + (k/8)*LBO + (n%8)*16 + (k%8)*2
constexpr uint32_t kWin_LBO = 128 * (3 * TP_D / 8), kWo_LBO = 128 * (TP_D / 8), kSBO = 128;
constexpr uint32_t kWkvRowOff = (TP_D / 8) * kSBO;  // rows 48.. of the W_in image: the [W_k | W_v] sub-matrix
constexpr int kKvStride = 100;  // floats per K|V row: rows 14 apart are 24 banks apart, neighbours 4 banks apart

struct Smem {
  __align__(16) unsigned char w[ATT_LAYER_BYTES];
  __align__(16) float kv[kTM][kKvStride];  // K (48) | V (48) | pad
  uint64_t bar_w, bar_mma;
  uint32_t tmem_base;
};
// tensor-memory columns (32-bit words per lane; lane = token row; fp16 operands hold two K elements per word)
constexpr uint32_t kT_KV1 = 0, kT_KV2 = 24;   // pieces of the key/value-side tokens (self-attention: THE tokens)
constexpr uint32_t kT_Q1 = 48, kT_Q2 = 72;    // pieces of the query-side tokens (cross only), later of the attention output
constexpr uint32_t kT_QKV = 96;               // Q (48) | K (48) | V (48) fp32; the Q columns are reused by the output projection
constexpr uint32_t kT_COLS = 256;
constexpr uint32_t kIdescF16 = (1u << 4);     // fp32 accumulate, fp16 A/B, both K-major

template <int N>
__device__ __forceinline__ void issue_proj(uint32_t tmem, uint32_t d_col, uint32_t a1_col, uint32_t a2_col, uint32_t w_smem, uint32_t piece_bytes,
                                           uint32_t lbo) {
  constexpr uint32_t idesc = kIdescF16 | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTM >> 4) << 24);
  const UmmaDescBase w1 = umma_desc_base(w_smem, lbo, kSBO), w2 = umma_desc_base(w_smem + piece_bytes, lbo, kSBO);
  const uint32_t d = tmem + d_col, a1 = tmem + a1_col, a2 = tmem + a2_col;
#pragma unroll
  for (int k = 0; k < TP_D / 16; ++k) {
    const uint32_t bo = k * 2 * lbo;
    if (k == 0) umma_f16_ts_c<false>(d, a2 + 8 * k, umma_desc_at(w1, bo), idesc);
    else umma_f16_ts_c<true>(d, a2 + 8 * k, umma_desc_at(w1, bo), idesc);
    umma_f16_ts_c<true>(d, a1 + 8 * k, umma_desc_at(w2, bo), idesc);
    umma_f16_ts_c<true>(d, a1 + 8 * k, umma_desc_at(w1, bo), idesc);
  }
}

// one 48-float token row -> its two fp16 pieces in tensor memory (zeros for rows outside the tile)
__device__ __forceinline__ void row_to_tmem(const float* src, bool valid, uint32_t t1, uint32_t t2) {
  float p1[24], p2[24];
#pragma unroll
  for (int j = 0; j < TP_D; j += 4) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) v = __ldcg(reinterpret_cast<const float4*>(src + j));
    split_h2(v.x, v.y, p1[j / 2], p2[j / 2]);
    split_h2(v.z, v.w, p1[j / 2 + 1], p2[j / 2 + 1]);
  }
  tmem_st8(t1, reinterpret_cast<float (&)[8]>(p1[0]));
  tmem_st8(t1 + 8, reinterpret_cast<float (&)[8]>(p1[8]));
  tmem_st8(t1 + 16, reinterpret_cast<float (&)[8]>(p1[16]));
  tmem_st8(t2, reinterpret_cast<float (&)[8]>(p2[0]));
  tmem_st8(t2 + 8, reinterpret_cast<float (&)[8]>(p2[8]));
  tmem_st8(t2 + 16, reinterpret_cast<float (&)[8]>(p2[16]));
  tmem_st_wait();
}

// Attention of one query row for the head pair 2h, 2h+1 over the S keys of its clip (rows of the K|V tile in shared memory):
// scores, max-subtracted softmax, P.V.  Packed fp32x2 arithmetic (FFMA2): a float4 of a K or V row is two aligned pairs.
// NS > 0 fixes the key count at compile time so that the loads of all keys can be issued ahead of the arithmetic.
template <int NS, int SMAX>
__device__ __forceinline__ void attend(const float* __restrict__ kv0, const float (&q)[24], int h, int S, float (&o)[24]) {
  constexpr int N = NS > 0 ? NS : SMAX;
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    const int col = (2 * h + hh) * TP_HD;
    float2 qp[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) qp[i] = make_float2(q[TP_HD * hh + 2 * i], q[TP_HD * hh + 2 * i + 1]);
    float sc[N], mx = -3.0e38f;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      sc[j] = -3.0e38f;
      if (NS > 0 || j < S) {
        const float4* kr = reinterpret_cast<const float4*>(kv0 + j * kKvStride + col);
        const float4 k0 = kr[0], k1 = kr[1], k2 = kr[2];
        float2 a = __fmul2_rn(qp[0], make_float2(k0.x, k0.y));
        a = __ffma2_rn(qp[1], make_float2(k0.z, k0.w), a);
        a = __ffma2_rn(qp[2], make_float2(k1.x, k1.y), a);
        a = __ffma2_rn(qp[3], make_float2(k1.z, k1.w), a);
        a = __ffma2_rn(qp[4], make_float2(k2.x, k2.y), a);
        a = __ffma2_rn(qp[5], make_float2(k2.z, k2.w), a);
        sc[j] = a.x + a.y;
        mx = fmaxf(mx, sc[j]);
      }
    }
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (NS > 0 || j < S) { sc[j] = __expf(sc[j] - mx); sum += sc[j]; }
    const float inv = 1.0f / sum;
    float2 op[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) op[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (NS > 0 || j < S) {
        const float4* vr = reinterpret_cast<const float4*>(kv0 + j * kKvStride + TP_D + col);
        const float4 v0 = vr[0], v1 = vr[1], v2 = vr[2];
        const float pj = sc[j] * inv;
        const float2 pp = make_float2(pj, pj);
        op[0] = __ffma2_rn(pp, make_float2(v0.x, v0.y), op[0]);
        op[1] = __ffma2_rn(pp, make_float2(v0.z, v0.w), op[1]);
        op[2] = __ffma2_rn(pp, make_float2(v1.x, v1.y), op[2]);
        op[3] = __ffma2_rn(pp, make_float2(v1.z, v1.w), op[3]);
        op[4] = __ffma2_rn(pp, make_float2(v2.x, v2.y), op[4]);
        op[5] = __ffma2_rn(pp, make_float2(v2.z, v2.w), op[5]);
      }
#pragma unroll
    for (int i = 0; i < 6; ++i) { o[TP_HD * hh + 2 * i] = op[i].x; o[TP_HD * hh + 2 * i + 1] = op[i].y; }
  }
}

constexpr int kWorkThreads = 256, kThreads = kWorkThreads + 32;  // 8 worker warps + 1 MMA/TMA issuer warp

// Rows of a tile: query row m = g*T + t, key/value row m = g*S + s (g = clip within the tile).  xq == xkv (and T == S) for
// self-attention.  SMAX bounds S (the score array lives in registers).
// One attention block for the tile of G clips starting at clip0.  Called by every thread of the CTA (>= kThreads threads: warps
// 0..7 work, warp 8 issues TMA and MMA, further warps only take part in the block barriers) with kT_COLS columns of tensor memory
// at `tmem` and the two mbarriers of S_ freshly initialised and published by a block barrier; returns after a block barrier.
// The token rows are read with ld.global.cg and written with plain stores: in the fused encoder kernel a thread re-reads rows it
// wrote itself earlier in the same launch.
__device__ __forceinline__ void attn_init_barriers(Smem& S_) {
  mbar_init(&S_.bar_w, 1);
  mbar_init(&S_.bar_mma, 1);
}
__device__ __forceinline__ void attn_inval_barriers(Smem& S_) {
  mbar_inval(&S_.bar_w);
  mbar_inval(&S_.bar_mma);
}
template <int SMAX>
__device__ __forceinline__ void attn_tile(Smem& S_, const uint32_t tmem, const unsigned char* wimg, const float* blob, TpNorm N1, const float* xq_g, int T,
                                          int q_stride, const float* xkv_g, int S, int kv_stride, int n_clips, int G, int clip0, float* out_g,
                                          long long* trace) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (trace && blockIdx.x == 0 && tid == 0) trace[1] = clock64();
  const int g_here = min(G, n_clips - clip0);
  const bool cross = xq_g != xkv_g;
  if (warp == 8 && elect_one()) {  // the whole weight image of this block: two bulk copies
    constexpr uint32_t kHalf = 2 * kWinBytes;
    mbar_expect_tx(&S_.bar_w, ATT_LAYER_BYTES);
    tma_bulk_g2s(S_.w, wimg, kHalf, &S_.bar_w);
    tma_bulk_g2s(S_.w + kHalf, wimg + kHalf, ATT_LAYER_BYTES - kHalf, &S_.bar_w);
  }
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const int m = (warp & 3) * 32 + lane;  // tile row == TMEM lane owned by this thread
  const int h = (warp >> 2) & 1;         // worker warps 4..7 take the second head pair / the V half
  const bool q_valid = m < g_here * T, kv_valid = m < g_here * S;
  const size_t gq = q_valid ? ((size_t)(clip0 + m / T) * q_stride + m % T) * TP_D : 0;
  const size_t gkv = kv_valid ? ((size_t)(clip0 + m / S) * kv_stride + m % S) * TP_D : 0;
  if (warp < 4) row_to_tmem(xkv_g + gkv, kv_valid, tmem + lane_base + kT_KV1, tmem + lane_base + kT_KV2);
  else if (warp < 8 && cross) row_to_tmem(xq_g + gq, q_valid, tmem + lane_base + kT_Q1, tmem + lane_base + kT_Q2);
  tc_fence_before();
  __syncthreads();
  if (trace && blockIdx.x == 0 && tid == 0) trace[2] = clock64();
  if (warp == 8) {
    tc_fence_after();
    mbar_wait(&S_.bar_w, 0);
    if (elect_one()) {
      const uint32_t w_in = smem_u32(S_.w) + kOffWin;
      if (cross) {
        issue_proj<TP_D>(tmem, kT_QKV, kT_Q1, kT_Q2, w_in, kWinBytes, kWin_LBO);
        issue_proj<2 * TP_D>(tmem, kT_QKV + TP_D, kT_KV1, kT_KV2, w_in + kWkvRowOff, kWinBytes, kWin_LBO);
      } else {
        issue_proj<3 * TP_D>(tmem, kT_QKV, kT_KV1, kT_KV2, w_in, kWinBytes, kWin_LBO);
      }
      umma_commit(&S_.bar_mma);
    }
    __syncwarp();
  }
  float q[24];  // Q of this row for heads 2h, 2h+1, already scaled by 1/sqrt(head_dim)
  if (warp < 8) {
    mbar_wait(&S_.bar_mma, 0);
    tc_fence_after();
    mbar_wait(&S_.bar_w, 0);  // acquire the TMA-written bias vectors
    const float* b_in = reinterpret_cast<const float*>(S_.w + kOffBin);
    // K (h = 0) or V (h = 1) of key/value row m -> shared memory; Q of heads 2h, 2h+1 -> registers (loads batched: two waits)
    {
      float v[TP_D];
      tmem_ld16(tmem + lane_base + kT_QKV + (uint32_t)(TP_D + TP_D * h), reinterpret_cast<float (&)[16]>(v[0]));
      tmem_ld16(tmem + lane_base + kT_QKV + (uint32_t)(TP_D + TP_D * h + 16), reinterpret_cast<float (&)[16]>(v[16]));
      tmem_ld16(tmem + lane_base + kT_QKV + (uint32_t)(TP_D + TP_D * h + 32), reinterpret_cast<float (&)[16]>(v[32]));
      tmem_ld_wait();
      const float* bb = b_in + TP_D + TP_D * h;
#pragma unroll
      for (int j = 0; j < TP_D; j += 4)
        *reinterpret_cast<float4*>(&S_.kv[m][TP_D * h + j]) =
            make_float4(fmaf(v[j], 1.0f / kWScale, bb[j]), fmaf(v[j + 1], 1.0f / kWScale, bb[j + 1]), fmaf(v[j + 2], 1.0f / kWScale, bb[j + 2]),
                        fmaf(v[j + 3], 1.0f / kWScale, bb[j + 3]));
    }
    {
      const float qs = rsqrtf((float)TP_HD);
      float v[24];
      tmem_ld8(tmem + lane_base + kT_QKV + (uint32_t)(24 * h), reinterpret_cast<float (&)[8]>(v[0]));
      tmem_ld8(tmem + lane_base + kT_QKV + (uint32_t)(24 * h + 8), reinterpret_cast<float (&)[8]>(v[8]));
      tmem_ld8(tmem + lane_base + kT_QKV + (uint32_t)(24 * h + 16), reinterpret_cast<float (&)[8]>(v[16]));
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 24; ++j) q[j] = fmaf(v[j], 1.0f / kWScale, b_in[24 * h + j]) * qs;
    }
    tc_fence_before();
  }
  __syncthreads();  // K | V of the tile visible; the accumulator columns are free again
  if (trace && blockIdx.x == 0 && tid == 0) trace[3] = clock64();
  if (warp < 8) {
    float o[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) o[i] = 0.0f;
    if (q_valid && S == 1) {  // a single key (first decoder pass): softmax is 1, the output is that key's V row
      const float4* vr = reinterpret_cast<const float4*>(&S_.kv[m / T][TP_D + 24 * h]);
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const float4 t = vr[i];
        o[4 * i] = t.x; o[4 * i + 1] = t.y; o[4 * i + 2] = t.z; o[4 * i + 3] = t.w;
      }
    } else if (q_valid) {
      const float* kv0 = &S_.kv[(m / T) * S][0];
      if (S == TP_S) attend<TP_S, TP_S>(kv0, q, h, TP_S, o);  // 14 keys (encoder self-, every cross-attention): compile-time trip count
      else attend<0, SMAX>(kv0, q, h, S, o);
    }
    // attention output (features 24h .. 24h+23 of this row) -> fp16 pieces, words 12h .. 12h+11 of each piece
    float p1[12], p2[12];
#pragma unroll
    for (int j = 0; j < 24; j += 2) split_h2(o[j], o[j + 1], p1[j / 2], p2[j / 2]);
    tc_fence_after();
    tmem_st8(tmem + lane_base + kT_Q1 + (uint32_t)(12 * h), reinterpret_cast<float (&)[8]>(p1[0]));
    tmem_st4(tmem + lane_base + kT_Q1 + (uint32_t)(12 * h + 8), reinterpret_cast<float (&)[4]>(p1[8]));
    tmem_st8(tmem + lane_base + kT_Q2 + (uint32_t)(12 * h), reinterpret_cast<float (&)[8]>(p2[0]));
    tmem_st4(tmem + lane_base + kT_Q2 + (uint32_t)(12 * h + 8), reinterpret_cast<float (&)[4]>(p2[8]));
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  if (trace && blockIdx.x == 0 && tid == 0) trace[4] = clock64();
  if (warp == 8) {
    tc_fence_after();
    if (elect_one()) {
      issue_proj<TP_D>(tmem, kT_QKV, kT_Q1, kT_Q2, smem_u32(S_.w) + kOffWo, kWoBytes, kWo_LBO);
      umma_commit(&S_.bar_mma);
    }
    __syncwarp();
  }
  if (warp < 4) {
    mbar_wait(&S_.bar_mma, 1);
    tc_fence_after();
    float o[TP_D];
    tmem_ld16(tmem + lane_base + kT_QKV, reinterpret_cast<float (&)[16]>(o[0]));  // three loads in flight, one wait
    tmem_ld16(tmem + lane_base + kT_QKV + 16u, reinterpret_cast<float (&)[16]>(o[16]));
    tmem_ld16(tmem + lane_base + kT_QKV + 32u, reinterpret_cast<float (&)[16]>(o[32]));
    tmem_ld_wait();
    if (q_valid) {
      const float* b_o = reinterpret_cast<const float*>(S_.w + kOffBo);
#pragma unroll
      for (int j = 0; j < TP_D; j += 4) {
        const float4 xv = __ldcg(reinterpret_cast<const float4*>(xq_g + gq + j));
        o[j] = fmaf(o[j], 1.0f / kWScale, b_o[j] + xv.x); o[j + 1] = fmaf(o[j + 1], 1.0f / kWScale, b_o[j + 1] + xv.y);
        o[j + 2] = fmaf(o[j + 2], 1.0f / kWScale, b_o[j + 2] + xv.z); o[j + 3] = fmaf(o[j + 3], 1.0f / kWScale, b_o[j + 3] + xv.w);
      }
      ln48(o, blob + N1.w, blob + N1.b);
#pragma unroll
      for (int j = 0; j < TP_D; j += 4) *reinterpret_cast<float4*>(out_g + gq + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (trace && blockIdx.x == 0 && tid == 0) trace[5] = clock64();
}


}  // namespace tpa
