// dp_selftest.cu -- hardware self-test of the tcgen05 building block: one CTA computes
// D[128 x N] = A * B^T from raw shared-memory images supplied by the host, with the operand
// layout (LBO / SBO / major-ness / per-k-step start offsets) given at run time.  Used by the
// GPU tests to pin the descriptor conventions that the decoder and predictor kernels rely on.
#include <cuda_runtime.h>

#include <string>

#include "../../include/dp_engine.h"
#include "dp_umma.cuh"

namespace {

struct SelftestArgs {
  const unsigned char* a_img;  // raw bytes copied to shared memory
  const unsigned char* b_img;
  float* d_out;                // [128][N] row-major
  long long* cycles;           // clock64 ticks from first issue to completion
  uint32_t a_bytes, b_bytes;
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  uint32_t a_kstep, b_kstep;   // start-address advance per K=8 step, bytes
  int N, ksteps, a_mn, b_mn, passes, kind, a_tmem, nacc;
};

__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const SelftestArgs T) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  unsigned char* sa = smem;
  unsigned char* sb = smem + ((T.a_bytes + 1023u) & ~1023u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (uint32_t i = tid * 16; i < T.a_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(sa + i) = *reinterpret_cast<const uint4*>(T.a_img + i);
  for (uint32_t i = tid * 16; i < T.b_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(sb + i) = *reinterpret_cast<const uint4*>(T.b_img + i);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (T.a_tmem == 2) {  // A given row-major [128][16*ksteps] 16-bit: row m -> TMEM lane m, two K elements per 32-bit column
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(T.a_img) + (size_t)tid * 8 * T.ksteps;
    for (int c0 = 0; c0 < 8 * T.ksteps; c0 += 16) {
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = (c0 + i < 8 * T.ksteps) ? __uint_as_float(arow[c0 + i]) : 0.f;
      tmem_st16(tmem_base + ((uint32_t)(warp * 32) << 16) + 256u + (uint32_t)c0, v);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  } else if (T.a_tmem) {  // A given row-major [128][8*ksteps] fp32: row m -> TMEM lane m, columns 256..
    const float* arow = reinterpret_cast<const float*>(T.a_img) + (size_t)tid * 8 * T.ksteps;
    for (int c0 = 0; c0 < 8 * T.ksteps; c0 += 16) {
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = (c0 + i < 8 * T.ksteps) ? arow[c0 + i] : 0.f;
      tmem_st16(tmem_base + ((uint32_t)(warp * 32) << 16) + 256u + (uint32_t)c0, v);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  long long t0 = 0;
  if (tid == 0) {
    t0 = clock64();
    uint32_t idesc = T.kind ? umma_idesc_bf16(128, T.N, T.a_mn, T.b_mn) : umma_idesc_tf32(128, T.N, T.a_mn, T.b_mn);
    if (T.kind == 2) idesc &= ~((1u << 7) | (1u << 10));  // fp16 operands (format 0)
    for (int p = 0; p < T.passes; ++p)  // passes > 1 re-accumulates the same product (tests the accumulate flag)
      for (int k = 0; k < T.ksteps; ++k) {
        const uint64_t da = umma_smem_desc(smem_u32(sa) + k * T.a_kstep, T.a_lbo, T.a_sbo);
        const uint64_t db = umma_smem_desc(smem_u32(sb) + k * T.b_kstep, T.b_lbo, T.b_sbo);
        // nacc > 1: rotate over independent accumulators (timing probe only; results of accumulators 1.. are discarded)
        const uint32_t dcol = tmem_base + (T.nacc > 1 ? (uint32_t)(((p * T.ksteps + k) % T.nacc) * 32) : 0u);
        const uint32_t acc = T.nacc > 1 ? ((p * T.ksteps + k) >= T.nacc ? 1u : 0u) : ((p | k) ? 1u : 0u);
        if (T.a_tmem == 2) {
          if (acc) umma_f16_ts_c<true>(dcol, tmem_base + 256u + 8u * k, db, idesc);
          else umma_f16_ts_c<false>(dcol, tmem_base + 256u + 8u * k, db, idesc);
        } else if (T.a_tmem) umma_tf32_ts(dcol, tmem_base + 256u + 8u * k, db, idesc, acc);
        else if (T.kind) umma_bf16_ss(dcol, da, db, idesc, acc);
        else umma_tf32_ss(dcol, da, db, idesc, acc);
      }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  if (tid == 0 && T.cycles) *T.cycles = clock64() - t0;
  tc_fence_after();
  for (int c0 = 0; c0 < T.N; c0 += 8) {
    float v[8];
    tmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) T.d_out[(size_t)(warp * 32 + lane) * T.N + c0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---- clean issue-rate probe: one converged warp, elect.sync lane, 64 fully unrolled MMAs with precomputed descriptors
template <int KIND, bool TS, int NACC>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int N, int reps, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.0f;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 1) {
    const uint32_t idesc = KIND ? umma_idesc_bf16(128, N, 0, 0) : umma_idesc_tf32(128, N, 0, 0);
    const UmmaDescBase da = umma_desc_base(smem_u32(smem), 128, 1024);
    const UmmaDescBase db = umma_desc_base(smem_u32(smem) + 16384, 128 * (N / 8), 128);
    long long t0 = 0;
    uint32_t phase = 0;
    for (int r = 0; r < reps; ++r) {
      if (r == 1) t0 = clock64();
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const uint32_t d = tmem + (uint32_t)((i % NACC) * 64);
          const uint32_t boff = (uint32_t)((i % 8) * 256);
          if (TS) umma_tf32_ts_c<true>(d, tmem + 384u + 8u * (i % 8), umma_desc_at(db, boff), idesc);
          else if (KIND) umma_bf16_ss(d, umma_desc_at(da, boff), umma_desc_at(db, boff), idesc, 1u);
          else umma_tf32_ss(d, umma_desc_at(da, boff), umma_desc_at(db, boff), idesc, 1u);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, phase);
      phase ^= 1;
    }
    if (elect_one()) *cycles = (clock64() - t0) / ((reps - 1) * 64);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace

extern "C" __attribute__((visibility("default"))) long long dp_selftest_umma_rate(int kind, int ts, int nacc, int N) {
  long long* dc = nullptr;
  long long h = -1;
  if (cudaMalloc(&dc, 8) != cudaSuccess) return -1;
  const size_t smem = 64 * 1024;
#define DP_RATE(K, T, A)                                                                              \
  {                                                                                                   \
    cudaFuncSetAttribute(umma_rate_kernel<K, T, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    umma_rate_kernel<K, T, A><<<1, 128, smem>>>(N, 9, dc);                                            \
  }
  if (kind == 0 && !ts && nacc == 1) DP_RATE(0, false, 1)
  else if (kind == 0 && !ts) DP_RATE(0, false, 2)
  else if (kind == 0 && ts && nacc == 1) DP_RATE(0, true, 1)
  else if (kind == 0 && ts) DP_RATE(0, true, 2)
  else if (nacc == 1) DP_RATE(1, false, 1)
  else DP_RATE(1, false, 2)
#undef DP_RATE
  if (cudaDeviceSynchronize() == cudaSuccess) cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
  cudaFree(dc);
  return h;
}

extern "C" __attribute__((visibility("default"))) int dp_selftest_umma(const void* a_img, uint32_t a_bytes, const void* b_img,
                                                                        uint32_t b_bytes, uint32_t a_lbo, uint32_t a_sbo,
                                                                        uint32_t b_lbo, uint32_t b_sbo, uint32_t a_kstep,
                                                                        uint32_t b_kstep, int N, int ksteps, int a_mn, int b_mn,
                                                                        int passes, int kind, int a_tmem, int nacc, float* d_host, long long* cycles_host) {
  if (!a_img || !b_img || !d_host || N < 8 || N > 256 || (N % 8) || ksteps < 1 || (a_bytes % 16) || (b_bytes % 16)) return DP_ERR_ARG;
  SelftestArgs T{};
  unsigned char *da = nullptr, *db = nullptr;
  float* dd = nullptr;
  long long* dc = nullptr;
  const size_t smem = ((a_bytes + 1023u) & ~1023u) + ((b_bytes + 1023u) & ~1023u) + 1024;
  if (smem > 200 * 1024) return DP_ERR_ARG;
  if (cudaMalloc(&da, a_bytes) || cudaMalloc(&db, b_bytes) || cudaMalloc(&dd, 128 * (size_t)N * 4) || cudaMalloc(&dc, 8)) return DP_ERR_CUDA;
  cudaMemcpy(da, a_img, a_bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b_img, b_bytes, cudaMemcpyHostToDevice);
  T.a_img = da; T.b_img = db; T.d_out = dd; T.cycles = dc; T.a_bytes = a_bytes; T.b_bytes = b_bytes;
  T.a_lbo = a_lbo; T.a_sbo = a_sbo; T.b_lbo = b_lbo; T.b_sbo = b_sbo; T.a_kstep = a_kstep; T.b_kstep = b_kstep;
  T.N = N; T.ksteps = ksteps; T.a_mn = a_mn; T.b_mn = b_mn; T.passes = passes < 1 ? 1 : passes; T.kind = kind; T.a_tmem = a_tmem; T.nacc = nacc < 1 ? 1 : nacc;
  cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  umma_selftest_kernel<<<1, 128, smem>>>(T);
  cudaError_t err = cudaDeviceSynchronize();
  if (err == cudaSuccess) err = cudaMemcpy(d_host, dd, 128 * (size_t)N * 4, cudaMemcpyDeviceToHost);
  if (err == cudaSuccess && cycles_host) err = cudaMemcpy(cycles_host, dc, 8, cudaMemcpyDeviceToHost);
  cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(dc);
  return err == cudaSuccess ? DP_OK : DP_ERR_CUDA;
}
