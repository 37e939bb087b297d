// Per-SMSP issue cost of the instruction classes of the packed kinematics pass (B200), measured with clock64 inside one CTA per SM:
// cycles per warp-instruction for 1, 2 and 4 warps per SM sub-partition, independent chains (8 per warp).
//   SHFL.IDX, SHFL.BFLY, FFMA2, FMUL2, FSEL, LDS.128, and a 1:1 FFMA2 / SHFL mix.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smsp_rate smsp_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 2048, CH = 8;
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, long long* cyc, int src, float a) {
  __shared__ float4 sm[512];
  sm[threadIdx.x] = make_float4(threadIdx.x, 1.f, 2.f, 3.f);
  float v[CH];
  float2 p[CH];
  for (int i = 0; i < CH; ++i) { v[i] = threadIdx.x * 1e-3f + i; p[i] = make_float2(v[i], -v[i]); }
  const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(0.001f, 0.002f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (MODE == 0) v[i] = __shfl_sync(0xffffffffu, v[i], (src + i) & 31);
      if (MODE == 1) v[i] = __shfl_xor_sync(0xffffffffu, v[i], 1 << (i & 3));
      if (MODE == 2) p[i] = __ffma2_rn(p[i], a2, b2);
      if (MODE == 3) p[i] = __fmul2_rn(p[i], a2);
      if (MODE == 4) v[i] = (src + it) & 1 ? v[i] : v[(i + 1) % CH];
      if (MODE == 5) { const float4 t = sm[(threadIdx.x + (int)v[i]) & 511]; v[i] = t.x * 0.f + (float)((i + it) & 7); p[i].x += t.y; }
      if (MODE == 6) { p[i] = __ffma2_rn(p[i], a2, b2); v[i] = __shfl_sync(0xffffffffu, v[i], (src + i) & 31); }
      if (MODE == 7) v[i] = fmaf(v[i], a, 0.001f);
    }
  }
  const long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < CH; ++i) s += v[i] + p[i].x + p[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char* name, int per_instr, float* out, long long* cyc, int sms) {
  printf("%-22s", name);
  for (int warps : {1, 4, 8, 16}) {  // warps per CTA = per SM: 1 -> one SMSP busy; 4 -> one warp per SMSP; 8 -> two; 16 -> four
    k<MODE><<<sms, warps * 32>>>(out, cyc, 3, 0.999f);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const int per_smsp = warps < 4 ? 1 : warps / 4;
    printf("  %2d warps/SM: %6.2f cyc per warp-instr per SMSP", warps, (double)c / ((double)ITERS * CH * per_instr * per_smsp));
  }
  printf("\n");
}
int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; long long* cyc;
  cudaMalloc(&out, sms * 512 * sizeof(float));
  cudaMalloc(&cyc, 8);
  run<0>("SHFL.IDX", 1, out, cyc, sms);
  run<1>("SHFL.BFLY", 1, out, cyc, sms);
  run<2>("FFMA2", 1, out, cyc, sms);
  run<3>("FMUL2", 1, out, cyc, sms);
  run<7>("FFMA", 1, out, cyc, sms);
  run<4>("FSEL", 1, out, cyc, sms);
  run<5>("LDS.128 (+cvt, add)", 1, out, cyc, sms);
  run<6>("FFMA2 + SHFL pair", 2, out, cyc, sms);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
