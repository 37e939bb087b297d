// dp_encoder.cu -- clip start-up on the device: folded pose-VAE encoder + reparameterisation for a batch of clips.
//
// Reference: Encoder.forward (python/src/autoencoder.py:136-143: three masked skeleton convolutions + poolings, LeakyReLU 0.2,
// then f_mu / f_logvar) and Autoencoder.reparameterize (autoencoder.py:19-27): latent = mu + eps * exp(0.5 logvar).  The three
// conv + pool pairs fold exactly into dense layers 176 -> 112 -> 72 -> 48 (model.fold_generator_state), the heads are 48 -> 24.
// eps comes from the caller (torch's RNG stream cannot be reproduced on the device); eps == null gives the mean.
// One CTA per clip, every weight stored [in][out] so that consecutive threads read consecutive floats.  ~33 k MAC per clip:
// this runs once per clip, it only has to keep the start-up of thousands of clips off the host.
#include "dp_common.cuh"
#include "dp_internal.h"

namespace {

template <int IN, int OUT, bool ACT>
__device__ __forceinline__ void dense(const float* __restrict__ wt, const float* __restrict__ b, const float* x, float* y, int tid) {
  if (tid < OUT) {
    float a = b[tid];
#pragma unroll 8
    for (int i = 0; i < IN; ++i) a = fmaf(x[i], wt[i * OUT + tid], a);
    y[tid] = ACT ? (a > 0.0f ? a : 0.2f * a) : a;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(128) dp_encode_kernel(const float* __restrict__ blob, const float* __restrict__ dqs, const float* __restrict__ eps,
                                                        float* __restrict__ latent) {
  __shared__ float x0[DP_ENC_IN], x1[DP_ENC_H0], x2[DP_ENC_H1], x3[DP_ENC_H2], head[2 * DP_L];
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < DP_ENC_IN; i += blockDim.x) x0[i] = dqs[(size_t)b * DP_ENC_IN + i];
  __syncthreads();
  const float* p = blob;
  dense<DP_ENC_IN, DP_ENC_H0, true>(p, p + DP_ENC_IN * DP_ENC_H0, x0, x1, tid);
  p += DP_ENC_IN * DP_ENC_H0 + DP_ENC_H0;
  dense<DP_ENC_H0, DP_ENC_H1, true>(p, p + DP_ENC_H0 * DP_ENC_H1, x1, x2, tid);
  p += DP_ENC_H0 * DP_ENC_H1 + DP_ENC_H1;
  dense<DP_ENC_H1, DP_ENC_H2, true>(p, p + DP_ENC_H1 * DP_ENC_H2, x2, x3, tid);
  p += DP_ENC_H1 * DP_ENC_H2 + DP_ENC_H2;
  dense<DP_ENC_H2, 2 * DP_L, false>(p, p + DP_ENC_H2 * 2 * DP_L, x3, head, tid);  // [mu | logvar]
  if (tid < DP_L) {
    const float mu = head[tid], logvar = head[DP_L + tid];
    latent[(size_t)b * DP_L + tid] = eps ? fmaf(eps[(size_t)b * DP_L + tid], expf(0.5f * logvar), mu) : mu;
  }
}

}  // namespace

cudaError_t dp_encode_launch(const float* blob, const float* dqs, const float* eps, float* latent, int n, cudaStream_t st) {
  dp_encode_kernel<<<n, 128, 0, st>>>(blob, dqs, eps, latent);
  return cudaGetLastError();
}
