// HBM floor for the traffic mixes of the three hot kernels: R reads and W streaming writes of one 226 MB tensor each
// per launch (K1: 1r1w, K2: 2r1w, K3: 1r2w), plain grid-stride float4 kernels.  Build: nvcc -O3 -arch=sm_100a.
#include <cstdio>
#include <cuda_runtime.h>

template <int R, int W, bool CS>
__global__ void mix(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ o0,
                    float4* __restrict__ o1, size_t n) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float4 v = __ldg(a + i);
    if (R == 2) {
      const float4 w = __ldg(b + i);
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    if (CS) {
      __stcs(o0 + i, v);
      if (W == 2) __stcs(o1 + i, make_float4(v.y, v.x, v.w, v.z));
    } else {
      o0[i] = v;
      if (W == 2) o1[i] = make_float4(v.y, v.x, v.w, v.z);
    }
  }
}

template <int R, int W, bool CS>
void run(const char* name, const float4* a, const float4* b, float4* o0, float4* o1, size_t n, int ctas_per_sm) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = ctas_per_sm > 0 ? 148 * ctas_per_sm : int((n + 255) / 256);
  for (int i = 0; i < 3; ++i) mix<R, W, CS><<<grid, 256>>>(a, b, o0, o1, n);
  cudaEventRecord(e0);
  for (int i = 0; i < 20; ++i) mix<R, W, CS><<<grid, 256>>>(a, b, o0, o1, n);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 20;
  const double bytes = double(n) * 16 * (R + W);
  printf("%-28s grid %7d  %.4f ms  %.0f GB/s\n", name, grid, ms, bytes / ms * 1e-6);
}

int main() {
  const size_t n = 56459264 / 4;   // float4 elements of one cfg2 tensor (225.8 MB)
  float4 *a, *b, *o0, *o1;
  cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMalloc(&o0, n * 16); cudaMalloc(&o1, n * 16);
  cudaMemset(a, 0, n * 16); cudaMemset(b, 0, n * 16);
  for (int cps : {0, 8, 16}) {
    run<1, 1, false>("1r1w", a, b, o0, o1, n, cps);
    run<1, 1, true>("1r1w .cs", a, b, o0, o1, n, cps);
    run<2, 1, true>("2r1w .cs", a, b, o0, o1, n, cps);
    run<1, 2, false>("1r2w", a, b, o0, o1, n, cps);
    run<1, 2, true>("1r2w .cs", a, b, o0, o1, n, cps);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
