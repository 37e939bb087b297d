// dp_temporal.cuh -- weight blob layout of the temporal predictor (host + device).
//
// Reference: python/src/temporal_transformer.py:6-78 (Temporal), torch nn.Transformer
// (d_model 48, 4 heads, 3+3 post-norm layers, FF 2048, ReLU, final LayerNorms),
// python/src/positional_encoding.py:15-32, python/src/train_temporal.py:17-37.
// Every Linear weight is stored TRANSPOSED ([in][out]) so that consecutive threads
// (consecutive output features) read consecutive floats.
#pragma once
#include <stddef.h>

#define TP_D 48
#define TP_H 4
#define TP_HD 12
#define TP_FF 2048
#define TP_ENC_IN 33
#define TP_LAT 24
#define TP_S 14        // encoder tokens
#define TP_MAXT 32     // decoder tokens <= 1 + DP_MAX_WINDOW/4 = 30
#define TP_PE 30
#define TP_NENC 3
#define TP_NDEC 3
#define FFT_HC 64                 // hidden units per tensor-core FF chunk
#define FFT_STEP_BYTES 24576      // one pipeline step: W2c pieces of chunk t-2 (2 x 6144 fp16) + W1c pieces of chunk t (2 x 6144 fp16)
#define FFT_LAYER_BYTES (TP_FF * 4 + (TP_FF / FFT_HC + 2) * FFT_STEP_BYTES)   // b1 (fp32) + 34 steps
#define ATT_LAYER_BYTES 37632      // W_in pieces (2 x 13824 fp16) + W_o pieces (2 x 4608 fp16) + b_in (576) + b_o (192)

struct TpAttn {   // offsets (floats) into the blob
  size_t w_in;    // [48][144]  (q | k | v columns)
  size_t b_in;    // [144]
  size_t w_out;   // [48][48]
  size_t b_out;   // [48]
};
struct TpFF {
  size_t w1;      // [48][2048]
  size_t b1;      // [2048]
  size_t w2;      // [2048][48]
  size_t b2;      // [48]
};
struct TpNorm { size_t w, b; };

struct TpLayout {
  size_t enc_in_w, enc_in_b;   // [33][48], [48]
  size_t dec_in_w, dec_in_b;   // [24][48], [48]
  size_t pe;                   // [30][48]
  struct { TpAttn sa; TpFF ff; TpNorm n1, n2; } enc[TP_NENC];
  TpNorm enc_norm;
  struct { TpAttn sa, ca; TpFF ff; TpNorm n1, n2, n3; } dec[TP_NDEC];
  TpNorm dec_norm;
  size_t out_w, out_b;         // [48][24], [24]
  size_t total;
};

inline TpLayout tp_layout() {
  TpLayout L;
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 3) & ~size_t(3); return r; };  // keep 16-byte alignment
  auto attn = [&](TpAttn& a) { a.w_in = take(TP_D * 3 * TP_D); a.b_in = take(3 * TP_D); a.w_out = take(TP_D * TP_D); a.b_out = take(TP_D); };
  auto ff = [&](TpFF& f) { f.w1 = take(TP_D * TP_FF); f.b1 = take(TP_FF); f.w2 = take(TP_FF * TP_D); f.b2 = take(TP_D); };
  auto norm = [&](TpNorm& n) { n.w = take(TP_D); n.b = take(TP_D); };
  L.enc_in_w = take(TP_ENC_IN * TP_D); L.enc_in_b = take(TP_D);
  L.dec_in_w = take(TP_LAT * TP_D); L.dec_in_b = take(TP_D);
  L.pe = take(TP_PE * TP_D);
  for (int l = 0; l < TP_NENC; ++l) { attn(L.enc[l].sa); ff(L.enc[l].ff); norm(L.enc[l].n1); norm(L.enc[l].n2); }
  norm(L.enc_norm);
  for (int l = 0; l < TP_NDEC; ++l) { attn(L.dec[l].sa); attn(L.dec[l].ca); ff(L.dec[l].ff); norm(L.dec[l].n1); norm(L.dec[l].n2); norm(L.dec[l].n3); }
  norm(L.dec_norm);
  L.out_w = take(TP_D * TP_LAT); L.out_b = take(TP_LAT);
  L.total = o;
  return L;
}

// Optional work fused behind the feed-forward block of a SINGLE-TOKEN decoder pass (T == 1: the first autoregressive step, the only
// one when temporal_future_window == 0): the finished row goes straight through the next decoder layer's self-attention block
// (one key: softmax == 1, so the block is LN(x + W_o (W_v x + b_v) + b_o), a row-local map) and / or through the prediction head
// of drag_pose.py:275-289.  Saves the 32-CTA attention launch and the head launch of every layer of the pass.
struct TpFfTail {
  int next_self_attn;
  TpAttn sa;
  TpNorm n1;
  int out_head;
  size_t out_w, out_b;
  const float* mu;
  const float* sigma;
  float* dec_lat;
  float* target_buf;
  int step_i, window;
};

#ifdef __CUDACC__
// ---- one warp owns one 48-wide row: lane holds features lane and lane + 32 (< 48)
__device__ __forceinline__ float tp_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void tp_ln_row_warp(float& v0, float& v1, const float* __restrict__ w, const float* __restrict__ b, int lane) {
  const bool has1 = lane + 32 < TP_D;
  const float mean = tp_warp_sum(v0 + (has1 ? v1 : 0.0f)) * (1.0f / TP_D);
  const float d0 = v0 - mean, d1 = has1 ? v1 - mean : 0.0f;
  const float rstd = rsqrtf(tp_warp_sum(d0 * d0 + d1 * d1) * (1.0f / TP_D) + 1e-5f);
  v0 = d0 * rstd * w[lane] + b[lane];
  v1 = has1 ? d1 * rstd * w[lane + 32] + b[lane + 32] : 0.0f;
}
// y = bias + x W for a [K][ldw] weight block starting at column col0 (blob layout: [in][out]); K <= 48 inputs held like a row
template <int K>
__device__ __forceinline__ void tp_warp_matvec(float x0, float x1, const float* __restrict__ W, int ldw, int col0, const float* __restrict__ bias,
                                               int n_out, int lane, float& y0, float& y1) {
  const bool has0 = lane < n_out, has1 = lane + 32 < n_out;
  y0 = has0 ? bias[lane] : 0.0f;
  y1 = has1 ? bias[lane + 32] : 0.0f;
#pragma unroll 8
  for (int i = 0; i < K; ++i) {
    const float xi = __shfl_sync(0xffffffffu, i < 32 ? x0 : x1, i & 31);
    const float* w = W + (size_t)i * ldw + col0;
    if (has0) y0 = fmaf(xi, w[lane], y0);
    if (has1) y1 = fmaf(xi, w[lane + 32], y1);
  }
}
// self-attention block of a single token (its only key is itself): x <- LN(x + W_o (W_v x + b_v) + b_o)
__device__ __forceinline__ void tp_self_attn_single(const float* __restrict__ blob, const TpAttn& A, const TpNorm& N, int lane, float& x0, float& x1) {
  float v0, v1, o0, o1;
  tp_warp_matvec<TP_D>(x0, x1, blob + A.w_in, 3 * TP_D, 2 * TP_D, blob + A.b_in + 2 * TP_D, TP_D, lane, v0, v1);
  tp_warp_matvec<TP_D>(v0, v1, blob + A.w_out, TP_D, 0, blob + A.b_out, TP_D, lane, o0, o1);
  x0 += o0;
  x1 += o1;
  tp_ln_row_warp(x0, x1, blob + N.w, blob + N.b, lane);
}
// prediction head on a finished last-token row (drag_pose.py:275-289): appends the standardised prediction to the decoder inputs
// and writes the de-standardised one into target_buf with the step-function up-sampling
__device__ __forceinline__ void tp_out_head_row(const float* __restrict__ blob, const TpFfTail& t, int b, int T, int lane, float x0, float x1) {
  float a, unused;
  tp_warp_matvec<TP_D>(x0, x1, blob + t.out_w, TP_LAT, 0, blob + t.out_b, TP_LAT, lane, a, unused);
  if (lane >= TP_LAT) return;
  if (T < TP_MAXT) t.dec_lat[((size_t)b * TP_MAXT + T) * TP_LAT + lane] = a;
  const float val = a * t.sigma[lane] + t.mu[lane];
  float* tb = t.target_buf + (size_t)b * (t.window + 1) * TP_LAT;
  if (t.window == 0) {
    tb[lane] = val;
  } else if (t.step_i >= 4) {
    for (int r = t.step_i - 4; r < t.step_i; ++r) tb[r * TP_LAT + lane] = val;
    if (t.step_i == t.window) tb[t.window * TP_LAT + lane] = val;
  }
}

// LayerNorm of one 48-wide row held by one thread (eps 1e-5, torch nn.LayerNorm)
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
#endif
