// Microbenchmark: issue rate of packed fp32 (FFMA2 / FADD2 / FMUL2, PTX *.f32x2) against scalar FFMA on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pk(float a, float b) {
  float2 v = make_float2(a, b);
  return *reinterpret_cast<unsigned long long*>(&v);
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float ffma1(float a, float b, float c) {
  float d;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

constexpr int ITER = 4096;

template <int MODE>
__global__ void bench(float* out, long long* cycles, float seed) {
  float a = seed + threadIdx.x, b = 1.0001f;
  long long t0, t1;
  if (MODE == 0) {   // 16 independent scalar FFMA chains
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = j;
    t0 = clock64();
    for (int it = 0; it < ITER; ++it)
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = ffma1(a, b, acc[j]);
    t1 = clock64();
    float s = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else if (MODE == 1) {   // 8 independent FFMA2 chains: the same number of flops
    unsigned long long acc[8], pa = pk(a, a + 1), pb = pk(b, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = pk(j, j + 0.5f);
    t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = ffma2(pa, pb, acc[j]);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = ffma2(pa, pb, acc[j]);
    }
    t1 = clock64();
    unsigned long long s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s ^= acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(s & 0xffff);
  } else {   // 16 FADD2 per iteration
    unsigned long long acc[8], pa = pk(a, a + 1);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = pk(j, j + 0.5f);
    t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fadd2(pa, acc[j]);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fadd2(pa, acc[j]);
    }
    t1 = clock64();
    unsigned long long s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s ^= acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(s & 0xffff);
  }
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  bench<MODE><<<148, threads>>>(out, cyc, 1.0f);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bench<MODE><<<148, threads>>>(out, cyc, 1.0f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double instr_per_warp = 16.0 * ITER;   // instructions issued per warp (scalar: 16 FFMA, packed: 16 x2-ops)
  const double warps_per_smsp = threads / 32 / 4.0;
  printf("%-8s threads/SM %4d: %lld cycles, %.2f cycles per warp-instruction per SMSP, %.3f ms, flops/clk/SM %.0f\n", name,
         threads, h[0], h[0] / (instr_per_warp * (warps_per_smsp < 1 ? 1 : warps_per_smsp)), ms,
         (MODE == 0 ? 2.0 : 4.0) * 32 * instr_per_warp * (threads / 32) / h[0]);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int threads : {128, 256, 512, 1024}) {
    run<0>("FFMA", threads);
    run<1>("FFMA2", threads);
    run<2>("FADD2", threads);
  }
  return 0;
}
