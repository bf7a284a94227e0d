// float64 compute dtype: the reference accepts compute_dtype=tf.float64 for both classes
// (/root/reference/audiocodec/mdctransformer.py:13-23, psychoacoustic.py:42-44).  Functional kernels, not the tuned
// fp32 path: B200's fp64 rate is a fraction of its fp32 rate and the north star is fp32; these exist so that a float64
// caller of the reference finds the same API.  One CTA per (frame, channel) for the MDCT (direct O(N^2) DCT-IV from a
// cos(pi m / 4N) table, any even N), one warp per (frame, channel) for the psychoacoustic model.
//
// Reference behaviour: mdctransformer.py:61-125 (transform), :127-153 (inverse_transform); psychoacoustic.py:102-120
// (tonality), :122-148, :169-210, :301-331 (global_masking_threshold); quantiser = SURVEY.md 8a row Q.
#include "kernels.h"

#include <algorithm>
#include <cstdint>

namespace ac {

namespace {

constexpr int kWarps = 8;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// window + fold (sparse F, tables.h), then X[k] = scale * sum_j u[j] cos(pi (2j+1)(2k+1) / 4N)   (mdctransformer.py:314)
__global__ void mdct_forward_f64_kernel(MdctDeviceTables64 tb, const double* __restrict__ x, double* __restrict__ y,
                                        int blocks_n, int C) {
  extern __shared__ double u64[];   // [N]
  const int n = tb.n, h = n / 2;
  const int frames = blocks_n + 1;
  const int64_t bf = blockIdx.x;                 // b * frames + f
  const int64_t b = bf / frames;
  const int f = static_cast<int>(bf - b * frames);
  const int c = blockIdx.y;
  const double* xb = x + b * static_cast<int64_t>(blocks_n) * n * C + c;
  for (int p = threadIdx.x; p < h; p += blockDim.x) {
    const double* a = tb.fold + 4 * p;
    double xp0 = 0., xp1 = 0., xc0 = 0., xc1 = 0.;
    if (f >= 1) {                                // the block in front of the signal is zero (mdctransformer.py:366)
      xp0 = xb[(static_cast<int64_t>(f - 1) * n + p) * C];
      xp1 = xb[(static_cast<int64_t>(f - 1) * n + n - 1 - p) * C];
    }
    if (f < blocks_n) {
      xc0 = xb[(static_cast<int64_t>(f) * n + p) * C];
      xc1 = xb[(static_cast<int64_t>(f) * n + n - 1 - p) * C];
    }
    u64[h - 1 - p] = fma(a[0], xp0, a[1] * xp1);
    u64[h + p] = fma(a[2], xc0, a[3] * xc1);
  }
  __syncthreads();
  const int period = 8 * n;
  double* yr = y + (bf * n) * C + c;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    double acc = 0.;
    int idx = (2 * k + 1) % period;            // (2j+1)(2k+1) mod 8N, stepped by 2(2k+1)
    const int step = (2 * (2 * k + 1)) % period;
    for (int j = 0; j < n; ++j) {
      acc = fma(u64[j], tb.cos_table[idx], acc);
      idx += step;
      if (idx >= period) idx -= period;
    }
    yr[static_cast<int64_t>(k) * C] = acc * tb.scale_fwd;
  }
}

// DCT-IV of frames blk and blk - 1, synthesis window + TDAC overlap-add into output block blk   (:138-153)
__global__ void mdct_inverse_f64_kernel(MdctDeviceTables64 tb, const double* __restrict__ y, double* __restrict__ x,
                                        int frames_n, int C) {
  extern __shared__ double sm64[];   // yn[N], yp[N], vn_low[h], vp_high[h]
  const int n = tb.n, h = n / 2;
  double* yn = sm64;
  double* yp = sm64 + n;
  double* vlow = sm64 + 2 * n;
  double* vhigh = vlow + h;
  const int out_blocks = frames_n + 1;
  const int64_t bb = blockIdx.x;
  const int64_t b = bb / out_blocks;
  const int blk = static_cast<int>(bb - b * out_blocks);
  const int c = blockIdx.y;
  const int64_t base = b * static_cast<int64_t>(frames_n) * n * C + c;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    yn[k] = blk < frames_n ? y[base + (static_cast<int64_t>(blk) * n + k) * C] : 0.;
    yp[k] = blk >= 1 ? y[base + (static_cast<int64_t>(blk - 1) * n + k) * C] : 0.;
  }
  __syncthreads();
  const int period = 8 * n;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const double* src = k < h ? yn : yp;     // lower half of v_n, upper half of v_{n-1}
    double acc = 0.;
    int idx = (2 * k + 1) % period;
    const int step = (2 * (2 * k + 1)) % period;
    for (int j = 0; j < n; ++j) {
      acc = fma(src[j], tb.cos_table[idx], acc);
      idx += step;
      if (idx >= period) idx -= period;
    }
    if (k < h) vlow[k] = acc * tb.scale_inv; else vhigh[k - h] = acc * tb.scale_inv;
  }
  __syncthreads();
  double* xo = x + (b * out_blocks + blk) * static_cast<int64_t>(n) * C + c;
  for (int p = threadIdx.x; p < h; p += blockDim.x) {
    const double* s = tb.unfold + 4 * p;
    const double vn = vlow[h - 1 - p], vp = vhigh[p];
    xo[static_cast<int64_t>(p) * C] = fma(s[0], vn, s[1] * vp);
    xo[static_cast<int64_t>(n - 1 - p) * C] = fma(s[2], vn, s[3] * vp);
  }
}

// tonality from the two frame sums (psychoacoustic.py:113-118)
__device__ __forceinline__ double tonality_from_sums(double sum_i, double sum_log, int n, double eps) {
  const double mean_log = sum_log / n;
  const double am = sum_i / n + eps;
  const double sfm = 10. * log(exp(mean_log) / am) / log(10.);
  return fmin(sfm / -60., 1.0);
}

__global__ void __launch_bounds__(kWarps * 32)
pa_tonality_f64_kernel(PaDeviceTables64 tb, const double* __restrict__ y, double* __restrict__ ton, int64_t items, int C) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kWarps;
  const int n = tb.n;
  for (int64_t item = warp0; item < items; item += stride) {
    const int64_t r = item / C;
    const int c = static_cast<int>(item - r * C);
    const double* base = y + r * n * C + c;
    double sum_i = 0., sum_log = 0.;
    for (int k = lane; k < n; k += 32) {
      const double a = base[static_cast<int64_t>(k) * C];
      const double in = a * a;
      sum_i += in;
      sum_log += log(fmax(tb.eps, in));
    }
    sum_i = warp_sum(sum_i);
    sum_log = warp_sum(sum_log);
    if (lane == 0) ton[item] = tonality_from_sums(sum_i, sum_log, n, tb.eps);
  }
}

// thr for one (frame, channel) per warp iteration; the masking matrix of the reference (:195-197) is never materialised
__global__ void __launch_bounds__(kWarps * 32)
pa_threshold_f64_kernel(PaDeviceTables64 tb, const double* __restrict__ y, const double* __restrict__ ton_in,
                        double one_minus_drown, double* __restrict__ thr_out, int64_t items, int C) {
  extern __shared__ double smem64[];
  const int n = tb.n, nb = tb.nb;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* inten = smem64 + warp * (n + 2 * nb);   // [N]  A^2
  double* pw = inten + n;                         // [nb] P = max(eps, I_bark)^alpha
  double* gm = pw + nb;                           // [nb] max(masking, quiet)
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * kWarps + warp;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kWarps;
  for (int64_t item = warp0; item < items; item += stride) {
    const int64_t r = item / C;
    const int c = static_cast<int>(item - r * C);
    const int64_t base = r * n * C + c;
    double sum_i = 0., sum_log = 0.;
    for (int k = lane; k < n; k += 32) {
      const double a = y[base + static_cast<int64_t>(k) * C];
      const double in = a * a;
      inten[k] = in;
      sum_i += in;
      sum_log += log(fmax(tb.eps, in));
    }
    double ton;
    if (ton_in != nullptr) {
      ton = ton_in[item];
    } else {
      sum_i = warp_sum(sum_i);
      sum_log = warp_sum(sum_log);
      ton = tonality_from_sums(sum_i, sum_log, n, tb.eps);
    }
    __syncwarp();
    for (int i = lane; i < nb; i += 32) {                  // (psychoacoustic.py:204-206, :313)
      const int k0 = tb.band_k0[i], cnt = tb.band_cnt[i], ptr = tb.band_ptr[i];
      double acc = 0.;
      for (int t = 0; t < cnt; ++t) acc = fma(inten[k0 + t], tb.band_w[ptr + t], acc);
      pw[i] = pow(fmax(tb.eps, acc), tb.alpha);
    }
    __syncwarp();
    for (int j = lane; j < nb; j += 32) {                  // (:185-208, :144)
      double acc = 0.;
      const double* sf = tb.spread_fn + nb + j;            // S[i][j] = spread_fn[nb - i + j]
      for (int i = 0; i < nb; ++i) acc = fma(pw[i], sf[-i], acc);
      const double offset = one_minus_drown * ((ton * tb.lin[j] + 9. * ton) + 5.5);
      const double gain = pow(10., -tb.alpha * offset / 10.);
      const double msk = pow(fmax(tb.eps, acc * gain), tb.inv_alpha);
      gm[j] = fmax(msk, tb.quiet[j]);
    }
    __syncwarp();
    for (int k = lane; k < n; k += 32) {                   // (:330-331)
      const int b0 = tb.filt_b0[k], cnt = tb.filt_cnt[k], ptr = tb.filt_ptr[k];
      double acc = 0.;
      for (int t = 0; t < cnt; ++t) acc = fma(gm[b0 + t], tb.filt_w[ptr + t], acc);
      thr_out[base + static_cast<int64_t>(k) * C] = sqrt(fmax(tb.eps, acc));
    }
    __syncwarp();
  }
}

__global__ void quantize_f64_kernel(const double* __restrict__ y, const double* __restrict__ thr, int32_t* __restrict__ q,
                                    int64_t n) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    q[i] = __double2int_rn(y[i] / thr[i]);                 // IEEE divide, round-half-even (== tf.round)
}

__global__ void dequantize_f64_kernel(const int32_t* __restrict__ q, const double* __restrict__ thr, double* __restrict__ y,
                                      int64_t n) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = static_cast<double>(q[i]) * thr[i];
}

unsigned grid_1d(int64_t work, int per_cta) {
  const int64_t want = (work + per_cta - 1) / per_cta;
  return static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>(want, 148 * 16)));
}

}  // namespace

cudaError_t mdct_forward_f64(const MdctDeviceTables64& tb, const double* x, double* y, int64_t batches, int64_t blocks_n,
                             int channels, cudaStream_t stream) {
  const int64_t ctas = batches * (blocks_n + 1);
  if (ctas == 0) return cudaSuccess;
  if (ctas > 2147483647LL || channels > 65535) return cudaErrorInvalidConfiguration;
  const size_t smem = static_cast<size_t>(tb.n) * sizeof(double);
  cudaError_t err = cudaFuncSetAttribute(mdct_forward_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
  if (err != cudaSuccess) return err;
  mdct_forward_f64_kernel<<<dim3(static_cast<unsigned>(ctas), channels), 128, smem, stream>>>(
      tb, x, y, static_cast<int>(blocks_n), channels);
  count_launch();
  return cudaGetLastError();
}

cudaError_t mdct_inverse_f64(const MdctDeviceTables64& tb, const double* y, double* x, int64_t batches, int64_t frames_n,
                             int channels, cudaStream_t stream) {
  const int64_t ctas = batches * (frames_n + 1);
  if (ctas == 0) return cudaSuccess;
  if (ctas > 2147483647LL || channels > 65535) return cudaErrorInvalidConfiguration;
  const size_t smem = static_cast<size_t>(3) * tb.n * sizeof(double);
  cudaError_t err = cudaFuncSetAttribute(mdct_inverse_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
  if (err != cudaSuccess) return err;
  mdct_inverse_f64_kernel<<<dim3(static_cast<unsigned>(ctas), channels), 128, smem, stream>>>(
      tb, y, x, static_cast<int>(frames_n), channels);
  count_launch();
  return cudaGetLastError();
}

cudaError_t pa_tonality_f64(const PaDeviceTables64& tb, const double* y, double* ton, int64_t rows, int channels,
                            cudaStream_t stream) {
  const int64_t items = rows * channels;
  if (items == 0) return cudaSuccess;
  pa_tonality_f64_kernel<<<grid_1d(items, kWarps), kWarps * 32, 0, stream>>>(tb, y, ton, items, channels);
  count_launch();
  return cudaGetLastError();
}

cudaError_t pa_threshold_f64(const PaDeviceTables64& tb, const double* y, const double* ton_in, double drown, double* thr,
                             int64_t rows, int channels, cudaStream_t stream) {
  const int64_t items = rows * channels;
  if (items == 0) return cudaSuccess;
  const size_t smem = static_cast<size_t>(kWarps) * (tb.n + 2 * tb.nb) * sizeof(double);
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t err = cudaFuncSetAttribute(pa_threshold_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
  if (err != cudaSuccess) return err;
  pa_threshold_f64_kernel<<<grid_1d(items, kWarps), kWarps * 32, smem, stream>>>(tb, y, ton_in, 1.0 - drown, thr, items,
                                                                                  channels);
  count_launch();
  return cudaGetLastError();
}

cudaError_t quantize_f64(const double* y, const double* thr, int32_t* q, int64_t n, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  quantize_f64_kernel<<<grid_1d(n, 256 * 4), 256, 0, stream>>>(y, thr, q, n);
  count_launch();
  return cudaGetLastError();
}

cudaError_t dequantize_f64(const int32_t* q, const double* thr, double* y, int64_t n, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  dequantize_f64_kernel<<<grid_1d(n, 256 * 4), 256, 0, stream>>>(q, thr, y, n);
  count_launch();
  return cudaGetLastError();
}

}  // namespace ac
