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

#ifdef __CUDACC__
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
