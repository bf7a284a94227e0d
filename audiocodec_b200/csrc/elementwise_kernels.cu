// Element-wise kernels around the hot path (sm_100a): bfloat16 <-> float32 conversion for the bfloat16 compute dtype
// and the dB utilities of the psychoacoustic model.
//
// Reference behaviour: /root/reference/audiocodec/psychoacoustic.py:71-85 (amplitude_to_dB), :87-100
// (amplitude_to_dB_norm); compute_dtype=tf.bfloat16 at psychoacoustic.py:42-44, 65-69 (tables cast to the compute dtype)
// and mdctransformer.py:58-59, 326-344 (up-cast to float32 around the DCT).  All of them are one read and one write per
// element: a grid-stride loop over 16-byte vectors, plain scalar loop for unaligned tensors and tails.
#include "kernels.h"

#include <cuda_bf16.h>

namespace ac {

namespace {

constexpr int kEwThreads = 256;

inline int ew_grid(int64_t work) {
  const int64_t want = (work + kEwThreads - 1) / kEwThreads;
  const int64_t cap = 148LL * 16;
  return static_cast<int>(want < 1 ? 1 : (want < cap ? want : cap));
}

__global__ void __launch_bounds__(kEwThreads) bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in,
                                                                 float* __restrict__ out, int64_t n, int vec) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t n8 = vec ? n / 8 : 0;
  for (int64_t i = i0; i < n8; i += stride) {                // 8 elements: one 16-byte load, two 16-byte stores
    const uint4 raw = __ldcs(reinterpret_cast<const uint4*>(in) + i);
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
    float4 a, b;
    a.x = __uint_as_float(w[0] << 16);
    a.y = __uint_as_float(w[0] & 0xffff0000u);
    a.z = __uint_as_float(w[1] << 16);
    a.w = __uint_as_float(w[1] & 0xffff0000u);
    b.x = __uint_as_float(w[2] << 16);
    b.y = __uint_as_float(w[2] & 0xffff0000u);
    b.z = __uint_as_float(w[3] << 16);
    b.w = __uint_as_float(w[3] & 0xffff0000u);
    reinterpret_cast<float4*>(out)[2 * i] = a;
    reinterpret_cast<float4*>(out)[2 * i + 1] = b;
  }
  for (int64_t i = 8 * n8 + i0; i < n; i += stride) out[i] = __bfloat162float(in[i]);
}

__global__ void __launch_bounds__(kEwThreads) f32_to_bf16_kernel(const float* __restrict__ in,
                                                                 __nv_bfloat16* __restrict__ out, int64_t n, int vec) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t n8 = vec ? n / 8 : 0;
  for (int64_t i = i0; i < n8; i += stride) {
    const float4 a = reinterpret_cast<const float4*>(in)[2 * i], b = reinterpret_cast<const float4*>(in)[2 * i + 1];
    const __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);   // round to nearest even
    const __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
    uint4 raw;
    raw.x = *reinterpret_cast<const uint32_t*>(&p0);
    raw.y = *reinterpret_cast<const uint32_t*>(&p1);
    raw.z = *reinterpret_cast<const uint32_t*>(&p2);
    raw.w = *reinterpret_cast<const uint32_t*>(&p3);
    __stcs(reinterpret_cast<uint4*>(out) + i, raw);
  }
  for (int64_t i = 8 * n8 + i0; i < n; i += stride) out[i] = __float2bfloat16_rn(in[i]);
}

// amplitude_to_dB (psychoacoustic.py:83-84): 10 log10(max(eps, a^2)) + dB_MAX, computed as the reference does
// (10 ln(.) / ln 10); NORM: (dB - dB_MIN) / (dB_MAX - dB_MIN)  (:99-100)
template <typename T, bool NORM>
__global__ void __launch_bounds__(kEwThreads) amplitude_to_db_kernel(const T* __restrict__ a, T* __restrict__ out,
                                                                     int64_t n, T eps, T db_max, T db_min) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const T ln10 = static_cast<T>(2.302585092994045684);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const T v = a[i];
    const T i2 = v * v;
    T db = static_cast<T>(10) * log(i2 > eps ? i2 : eps) / ln10 + db_max;
    if (NORM) db = (db - db_min) / (db_max - db_min);
    out[i] = db;
  }
}

}  // namespace

cudaError_t bf16_to_f32(const void* in, float* out, int64_t n, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const int vec = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 ? 1 : 0;
  bf16_to_f32_kernel<<<ew_grid(n / 8 + 1), kEwThreads, 0, stream>>>(static_cast<const __nv_bfloat16*>(in), out, n, vec);
  count_launch();
  return cudaGetLastError();
}

cudaError_t f32_to_bf16(const float* in, void* out, int64_t n, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const int vec = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 ? 1 : 0;
  f32_to_bf16_kernel<<<ew_grid(n / 8 + 1), kEwThreads, 0, stream>>>(in, static_cast<__nv_bfloat16*>(out), n, vec);
  count_launch();
  return cudaGetLastError();
}

cudaError_t amplitude_to_db_f32(const float* a, float* out, int64_t n, float eps, float db_max, float db_min, bool norm,
                                cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  if (norm) amplitude_to_db_kernel<float, true><<<ew_grid(n), kEwThreads, 0, stream>>>(a, out, n, eps, db_max, db_min);
  else amplitude_to_db_kernel<float, false><<<ew_grid(n), kEwThreads, 0, stream>>>(a, out, n, eps, db_max, db_min);
  count_launch();
  return cudaGetLastError();
}

cudaError_t amplitude_to_db_f64(const double* a, double* out, int64_t n, double eps, double db_max, double db_min, bool norm,
                                cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  if (norm) amplitude_to_db_kernel<double, true><<<ew_grid(n), kEwThreads, 0, stream>>>(a, out, n, eps, db_max, db_min);
  else amplitude_to_db_kernel<double, false><<<ew_grid(n), kEwThreads, 0, stream>>>(a, out, n, eps, db_max, db_min);
  count_launch();
  return cudaGetLastError();
}

}  // namespace ac
