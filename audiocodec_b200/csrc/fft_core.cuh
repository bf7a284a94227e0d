// Register-resident mixed-radix Stockham FFT shared by the forward and inverse MDCT kernels.
//
// One "group" of T = M / E threads transforms M complex points; every thread keeps E points in
// registers, does E / R radix-R butterflies per pass and the passes exchange data through a padded
// shared-memory buffer private to the group.  M = N / 2 where N is the MDCT size: an N-point DCT-IV is
// one M-point complex FFT between a pre- and a post-twiddle (see mdct_kernels.cu).
#pragma once

#include <cuda_runtime.h>

namespace ac {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// multiply by -i
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }

// ---- radix-R DFTs, forward sign (e^{-2 pi i nk / R}), natural-order in / natural-order out -----------
template <int R>
__device__ __forceinline__ void dft(float2* v);

template <>
__device__ __forceinline__ void dft<2>(float2* v) {
  const float2 a = v[0], b = v[1];
  v[0] = cadd(a, b);
  v[1] = csub(a, b);
}

__device__ __forceinline__ void dft4_regs(float2& v0, float2& v1, float2& v2, float2& v3) {
  const float2 a = cadd(v0, v2), b = csub(v0, v2), c = cadd(v1, v3), d = cmul_mi(csub(v1, v3));
  v0 = cadd(a, c);
  v1 = cadd(b, d);
  v2 = csub(a, c);
  v3 = csub(b, d);
}

template <>
__device__ __forceinline__ void dft<4>(float2* v) {
  dft4_regs(v[0], v[1], v[2], v[3]);
}

template <>
__device__ __forceinline__ void dft<8>(float2* v) {
  constexpr float kH = 0.70710678118654752440f;
  float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
  float2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
  dft4_regs(e0, e1, e2, e3);
  dft4_regs(o0, o1, o2, o3);
  // o_k *= w8^k
  o1 = make_float2((o1.x + o1.y) * kH, (o1.y - o1.x) * kH);
  o2 = cmul_mi(o2);
  o3 = make_float2((o3.y - o3.x) * kH, -(o3.x + o3.y) * kH);
  v[0] = cadd(e0, o0);
  v[1] = cadd(e1, o1);
  v[2] = cadd(e2, o2);
  v[3] = cadd(e3, o3);
  v[4] = csub(e0, o0);
  v[5] = csub(e1, o1);
  v[6] = csub(e2, o2);
  v[7] = csub(e3, o3);
}

template <>
__device__ __forceinline__ void dft<16>(float2* v) {
  constexpr float kH = 0.70710678118654752440f;   // cos(pi/4)
  constexpr float kC = 0.92387953251128675613f;   // cos(pi/8)
  constexpr float kS = 0.38268343236508977173f;   // sin(pi/8)
  // n = n1 + 4 n2, k = 4 k1 + k2:  A[n1][k2] = DFT4_{n2} x[n1 + 4 n2]
  float2 a[4][4];
#pragma unroll
  for (int n1 = 0; n1 < 4; ++n1) {
    a[n1][0] = v[n1];
    a[n1][1] = v[n1 + 4];
    a[n1][2] = v[n1 + 8];
    a[n1][3] = v[n1 + 12];
    dft4_regs(a[n1][0], a[n1][1], a[n1][2], a[n1][3]);
  }
  // A[n1][k2] *= w16^(n1 k2)
  a[1][1] = cmul(a[1][1], make_float2(kC, -kS));
  a[1][2] = make_float2((a[1][2].x + a[1][2].y) * kH, (a[1][2].y - a[1][2].x) * kH);   // w16^2 = w8
  a[1][3] = cmul(a[1][3], make_float2(kS, -kC));
  a[2][1] = make_float2((a[2][1].x + a[2][1].y) * kH, (a[2][1].y - a[2][1].x) * kH);   // w16^2
  a[2][2] = cmul_mi(a[2][2]);                                                             // w16^4
  a[2][3] = make_float2((a[2][3].y - a[2][3].x) * kH, -(a[2][3].x + a[2][3].y) * kH);  // w16^6
  a[3][1] = cmul(a[3][1], make_float2(kS, -kC));                                          // w16^3
  a[3][2] = make_float2((a[3][2].y - a[3][2].x) * kH, -(a[3][2].x + a[3][2].y) * kH);  // w16^6
  a[3][3] = cmul(a[3][3], make_float2(-kC, kS));                                          // w16^9
  // X[4 k1 + k2] = DFT4_{n1} A[n1][k2]
#pragma unroll
  for (int k2 = 0; k2 < 4; ++k2) {
    dft4_regs(a[0][k2], a[1][k2], a[2][k2], a[3][k2]);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) v[4 * k1 + k2] = a[k1][k2];
  }
}

// One padding slot every 16 complex values keeps the stride-R writes of the first pass conflict-free.
__device__ __forceinline__ int fft_pad(int i) { return i + (i >> 4); }
template <int M>
struct FftBuf {
  static constexpr int kSlots = M + (M >> 4) + 1;
};

// ---- compile-time FFT plan ------------------------------------------------------------------------------
// M complex points, E per thread, up to three passes with radices R0, R1, R2 (1 = pass absent).
template <int M_, int E_, int R0_, int R1_, int R2_>
struct FftPlan {
  static constexpr int M = M_, E = E_, R0 = R0_, R1 = R1_, R2 = R2_;
  static constexpr int T = M / E;                 // threads per transform
  static constexpr int RL = (R2 > 1) ? R2 : ((R1 > 1) ? R1 : R0);   // radix of the last pass
  static_assert(R0 * R1 * R2 == M, "radices must multiply to M");
  static_assert(E % R0 == 0 && E % R1 == 0 && E % R2 == 0, "each thread owns whole butterflies");
  // index of the point held in v[q * R + r] before a pass of radix R: (t + T q) + r * (M / R)
  __device__ static __forceinline__ int in_index(int t, int slot) {
    return (t + T * (slot / R0)) + (slot % R0) * (M / R0);
  }
  // index of the point held in v[q * RL + r] after the last pass
  __device__ static __forceinline__ int out_index(int t, int slot) {
    return (t + T * (slot / RL)) + (slot % RL) * (M / RL);
  }
};

// One middle/last pass: read from buf (natural Stockham input order), twiddle, butterflies.
template <int M, int E, int T, int R, int NS>
__device__ __forceinline__ void fft_pass_from_buf(float2* v, const float2* buf, int t, const float2* __restrict__ roots) {
#pragma unroll
  for (int q = 0; q < E / R; ++q) {
    const int j = t + T * q;
#pragma unroll
    for (int r = 0; r < R; ++r) v[q * R + r] = buf[fft_pad(j + r * (M / R))];
  }
#pragma unroll
  for (int q = 0; q < E / R; ++q) {
    const int j = t + T * q;
    const int k = j % NS;
#pragma unroll
    for (int r = 1; r < R; ++r) {
      // exp(-2 pi i r k / (NS R)) = roots[r k M / (NS R)],  r k < NS R
      const float2 w = __ldg(&roots[r * k * (M / (NS * R))]);
      v[q * R + r] = cmul(v[q * R + r], w);
    }
    dft<R>(&v[q * R]);
  }
}

// Scatter the outputs of a pass (radix R, stride NS) to buf in Stockham order.
template <int M, int E, int T, int R, int NS>
__device__ __forceinline__ void fft_pass_to_buf(const float2* v, float2* buf, int t) {
#pragma unroll
  for (int q = 0; q < E / R; ++q) {
    const int j = t + T * q;
    const int j0 = (j / NS) * NS * R + (j % NS);
#pragma unroll
    for (int r = 0; r < R; ++r) buf[fft_pad(j0 + r * NS)] = v[q * R + r];
  }
}

// Full transform.  On entry v[] holds the pass-0 inputs in Plan::in_index order; on exit it holds the
// spectrum in Plan::out_index order.  `buf` is this group's exchange buffer (FftBuf<M>::kSlots float2).
// All threads of the CTA must call this together (block-wide barriers).
template <typename Plan>
__device__ __forceinline__ void fft_forward(float2* v, float2* buf, int t, const float2* __restrict__ roots) {
  constexpr int M = Plan::M, E = Plan::E, T = Plan::T, R0 = Plan::R0, R1 = Plan::R1, R2 = Plan::R2;
#pragma unroll
  for (int q = 0; q < E / R0; ++q) dft<R0>(&v[q * R0]);
  if constexpr (R1 > 1) {
    fft_pass_to_buf<M, E, T, R0, 1>(v, buf, t);
    __syncthreads();
    fft_pass_from_buf<M, E, T, R1, R0>(v, buf, t, roots);
    if constexpr (R2 > 1) {
      __syncthreads();
      fft_pass_to_buf<M, E, T, R1, R0>(v, buf, t);
      __syncthreads();
      fft_pass_from_buf<M, E, T, R2, R0 * R1>(v, buf, t, roots);
    }
  }
}

}  // namespace ac
