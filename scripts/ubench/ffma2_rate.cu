// FFMA vs FFMA2 issue-rate microbenchmark (B200): the same number of fp32 FMAs through scalar and packed instructions.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096, CH = 8;
__global__ void __launch_bounds__(1024) k_scalar(float* out, float a, float b) {
  float acc[2 * CH];
  for (int i = 0; i < 2 * CH; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < 2 * CH; ++i) acc[i] = fmaf(acc[i], a, b);
  float s = 0;
  for (int i = 0; i < 2 * CH; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(1024) k_packed(float* out, float a, float b) {
  float2 acc[CH];
  for (int i = 0; i < CH; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 1e-3f - i);
  const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = __ffma2_rn(acc[i], a2, b2);
  float s = 0;
  for (int i = 0; i < CH; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  float* out;
  cudaMalloc(&out, sms * 1024 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k_scalar<<<sms, 1024>>>(out, 0.999f, 0.001f);
      else k_packed<<<sms, 1024>>>(out, 0.999f, 0.001f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double fma = (double)sms * 1024 * ITERS * 2 * CH;
      if (rep == 2)
        printf("%s: %.3f ms  %.1f fp32 FMA/clk/SM (at %d MHz nominal)  %.1f TFLOP/s\n", mode ? "FFMA2" : "FFMA ", ms,
               fma / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000, 2 * fma / (ms * 1e-3) / 1e12);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
