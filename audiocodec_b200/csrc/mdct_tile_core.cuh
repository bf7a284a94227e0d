// Device-side building blocks of the MDCT tile kernels (mdct_tile_kernels.cu) shared with the fused encoder
// (psycho_mma_kernels.cu): the two-sequence register FFT with its in-place shared-memory exchange, the post-twiddle
// and the tile shape.  See the header comment of mdct_tile_kernels.cu for the design.
#pragma once

#include "kernels.h"
#include "fft_core.cuh"

namespace ac {

namespace {

// the T threads of group g meet: a warp-level sync when the group is (part of) one warp, else a named barrier
template <int T>
__device__ __forceinline__ void group_sync(int g) {
  if constexpr (T <= 32) {
    __syncwarp();
  } else {
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(T) : "memory");
  }
}

// element `a` of the thread's two sequences: the two channels of one row (C == 2) or the same element of
// two adjacent rows (C == 1)
template <int C, int ROW>
__device__ __forceinline__ float2 ld2(const float* base, int a) {
  if constexpr (C == 2) {
    return *reinterpret_cast<const float2*>(base + 2 * a);
  } else {
    return make_float2(base[a], base[a + ROW]);
  }
}
template <int C, int ROW>
__device__ __forceinline__ void st2(float* base, int a, float2 v) {
  if constexpr (C == 2) {
    *reinterpret_cast<float2*>(base + 2 * a) = v;
  } else {
    base[a] = v.x;
    base[a + ROW] = v.y;
  }
}

template <typename Plan>
__device__ __forceinline__ int swz(int pos) {
  constexpr int SH = Plan::R0 == 16 ? 4 : (Plan::R0 == 8 ? 3 : (Plan::R0 == 4 ? 2 : 1));
  return pos ^ ((pos >> SH) & 7);
}

// ---- the FFT of two sequences; exchange through `scratch` (M float4: re0, im0, re1, im1) ---------------------
template <typename Plan, int R, int NS>
__device__ __forceinline__ void pass_to_scratch(const float2* v0, const float2* v1, float4* scratch, int t) {
  constexpr int E = Plan::E, T = Plan::T;
#pragma unroll
  for (int q = 0; q < E / R; ++q) {
    const int j = t + T * q;
    const int j0 = (j / NS) * NS * R + (j % NS);
#pragma unroll
    for (int r = 0; r < R; ++r)
      scratch[swz<Plan>(j0 + r * NS)] = make_float4(v0[q * R + r].x, v0[q * R + r].y, v1[q * R + r].x, v1[q * R + r].y);
  }
}

template <typename Plan, int R, int NS>
__device__ __forceinline__ void pass_from_scratch(float2* v0, float2* v1, const float4* scratch, int t,
                                                  const float2* __restrict__ tw) {
  constexpr int M = Plan::M, E = Plan::E, T = Plan::T;
#pragma unroll
  for (int q = 0; q < E / R; ++q) {
    const int j = t + T * q;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float4 x = scratch[swz<Plan>(j + r * (M / R))];
      v0[q * R + r] = make_float2(x.x, x.y);
      v1[q * R + r] = make_float2(x.z, x.w);
    }
  }
#pragma unroll
  for (int q = 0; q < E / R; ++q) {
    const int j = t + T * q;
#pragma unroll
    for (int r = 1; r < R; ++r) {
      const float2 w = __ldg(&tw[(r - 1) * (M / R) + j]);      // exp(-2 pi i r (j mod NS) / (NS R)), contiguous in j
      v0[q * R + r] = cmul(v0[q * R + r], w);
      v1[q * R + r] = cmul(v1[q * R + r], w);
    }
    dft<R>(&v0[q * R]);
    dft<R>(&v1[q * R]);
  }
}

// On entry v0 / v1 hold the pass-0 inputs in Plan::in_index order and nobody reads `scratch` any more; on exit
// they hold the spectra in Plan::out_index order and every read of `scratch` by this group has completed.
template <typename Plan>
__device__ __forceinline__ void fft2(float2* v0, float2* v1, float4* scratch, int t, int g, const float2* __restrict__ tw1,
                                     const float2* __restrict__ tw2) {
  constexpr int E = Plan::E, T = Plan::T, R0 = Plan::R0, R1 = Plan::R1, R2 = Plan::R2;
#pragma unroll
  for (int q = 0; q < E / R0; ++q) {
    dft<R0>(&v0[q * R0]);
    dft<R0>(&v1[q * R0]);
  }
  if constexpr (R1 > 1) {
    pass_to_scratch<Plan, R0, 1>(v0, v1, scratch, t);
    group_sync<T>(g);
    pass_from_scratch<Plan, R1, R0>(v0, v1, scratch, t, tw1);
    if constexpr (R2 > 1) {
      group_sync<T>(g);
      pass_to_scratch<Plan, R1, R0>(v0, v1, scratch, t);
      group_sync<T>(g);
      pass_from_scratch<Plan, R2, R0 * R1>(v0, v1, scratch, t, tw2);
    }
    group_sync<T>(g);
  }
}

// post-twiddle: the two outputs of spectrum bin k go to positions 2k and N-1-2k (order set by the variant)
template <typename Plan, int C, int ROW>
__device__ __forceinline__ void post_store(const float2* v0, const float2* v1, float* out, int t, int variant,
                                           const float4* __restrict__ post) {
  constexpr int M = Plan::M, N = 2 * M, E = Plan::E;
#pragma unroll
  for (int s = 0; s < E; ++s) {
    const int k = Plan::out_index(t, s);
    const float4 c4 = __ldg(&post[variant * M + k]);
    const int i1 = variant ? N - 1 - 2 * k : 2 * k;
    const int i2 = (N - 1) - i1;
    st2<C, ROW>(out, i1, make_float2(fmaf(v0[s].y, c4.y, v0[s].x * c4.x), fmaf(v1[s].y, c4.y, v1[s].x * c4.x)));
    st2<C, ROW>(out, i2, make_float2(fmaf(v0[s].y, c4.w, v0[s].x * c4.z), fmaf(v1[s].y, c4.w, v1[s].x * c4.z)));
  }
}

template <typename Plan, int C, int THREADS>
struct TileShape {
  static constexpr int M = Plan::M, N = 2 * M, T = Plan::T;
  static constexpr int G = THREADS / T;              // groups per CTA
  static constexpr int FP = G * (2 / C);             // frames transformed per tile
  static constexpr int ROW = N * C;                  // floats per frame / block row
  static_assert(THREADS % T == 0 && G >= 1, "tile shape");
};


using Plan64 = FftPlan<32, 8, 8, 4, 1>;
using Plan128 = FftPlan<64, 8, 8, 8, 1>;
using Plan256 = FftPlan<128, 16, 16, 8, 1>;
using Plan512 = FftPlan<256, 16, 16, 16, 1>;
using Plan1024 = FftPlan<512, 8, 8, 8, 8>;      // 64 threads per transform pair: named barriers, 16 warps per SM

}  // namespace

}  // namespace ac
