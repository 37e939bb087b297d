// tcgen05.ld / tcgen05.st throughput on B200: bytes per clock per SM for the 32x32b shape, x32 (what the predictor's feed-forward
// epilogue issues), with 4, 8 and 16 warps per SM (16 = two co-resident 8-warp epilogues).  One CTA per SM, 512 columns allocated.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../dragposer_b200/csrc -o ldtm_rate ldtm_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "dp_umma.cuh"
constexpr int ITERS = 2048;
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, long long* cyc) {
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
  float v[32], acc = 0.f;
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
  tmem_st32(t, v);
  tmem_st32(t + 32, v);
  tmem_st_wait();
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) {  // loads only, two in flight
      float a[32], b[32];
      tmem_ld32(t, a);
      tmem_ld32(t + 32, b);
      tmem_ld_wait();
      acc += a[0] + a[31] + b[7] + b[16];
    } else {  // load 32 columns, store 32 columns back (the epilogue's pattern)
      float a[32];
      tmem_ld32(t, a);
      tmem_ld_wait();
      a[0] += 1.0f;
      tmem_st32(t + 32, a);
      tmem_st_wait();
      acc += a[3] + a[30];
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}
int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; long long* cyc;
  cudaMalloc(&out, sms * 512 * sizeof(float));
  cudaMalloc(&cyc, 8);
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {4, 8, 16}) {
      if (mode == 0) k<0><<<sms, warps * 32>>>(out, cyc); else k<1><<<sms, warps * 32>>>(out, cyc);
      cudaDeviceSynchronize();
      long long c;
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double bytes = (double)ITERS * warps * 32 * 32 * 4 * (mode == 0 ? 2 : 1);
      printf("%s, %2d warps/SM: %.1f cycles per warp-iteration, %.1f B/clk/SM read%s\n", mode ? "ld32 + st32" : "2 x ld32   ", warps,
             (double)c / ITERS, bytes / c, mode ? " (and as much written)" : "");
    }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
