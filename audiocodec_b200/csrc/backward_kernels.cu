// Backward pass of the psychoacoustic model (sm_100a): vector-Jacobian products of tonality and
// global_masking_threshold, so that the two classes work as differentiable layers.
//
// Reference behaviour: every method of /root/reference/audiocodec/psychoacoustic.py is a @tf.function of TensorFlow ops
// (:102, :122), i.e. differentiable by construction, and the comment at :311 speaks of the gradient; the gradients here
// are those of the formulas at :113-118 (tonality) and :139-146, :185-208, :312-313, :330-331 (threshold), with the
// sub-gradient 0 where a max(eps, .) or min(., 1) clamp is active and the quiet threshold's branch of max(., quiet)
// passing nothing.  (The MDCT needs no kernel of its own: for the orthogonal windows the adjoint of transform is
// inverse_transform / 4N and vice versa - audiocodec_b200/autograd.py.)
//
// One warp per (frame, channel) item, the structure of pa_threshold_kernel (psycho_kernels.cu): the forward
// intermediates are recomputed in shared memory, then the chain runs backwards.  Functional, not tuned.
#include "kernels.h"

namespace ac {

namespace {

constexpr int kBwdWarps = 4;

__device__ __forceinline__ float bwd_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// d tonality / d A  (psychoacoustic.py:113-118):
//   I = A^2, L = mean ln max(eps, I), AM = mean I + eps, sfm = 10 (L - ln AM) / ln 10, ton = min(sfm / -60, 1)
__global__ void __launch_bounds__(kBwdWarps * 32)
pa_tonality_backward_kernel(PaDeviceTables tb, const float* __restrict__ y, const float* __restrict__ grad_ton,
                            float* __restrict__ grad_y, int64_t items, int C) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * kBwdWarps + (threadIdx.x >> 5);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kBwdWarps;
  const int n = tb.n;
  const float eps = tb.eps;
  for (int64_t item = warp0; item < items; item += stride) {
    const int64_t r = item / C;
    const int c = static_cast<int>(item - r * C);
    const int64_t base = r * n * C + c;
    float sum_i = 0.f, sum_log = 0.f;
    for (int k = lane; k < n; k += 32) {
      const float a = y[base + static_cast<int64_t>(k) * C];
      const float in = a * a;
      sum_i += in;
      sum_log += logf(fmaxf(eps, in));
    }
    sum_i = bwd_warp_sum(sum_i);
    sum_log = bwd_warp_sum(sum_log);
    const float inv_n = 1.0f / static_cast<float>(n);
    const float am = sum_i * inv_n + eps;
    const float sfm = 10.f * (sum_log * inv_n - logf(am)) / 2.302585092994046f;
    // min(sfm / -60, 1): the clamp passes no gradient
    const float d_sfm = (sfm / -60.f < 1.0f) ? grad_ton[item] * (-1.0f / 60.f) : 0.f;
    const float d_l = d_sfm * (10.f / 2.302585092994046f) * inv_n;      // per-element factor of d mean ln
    const float d_am = -d_sfm * (10.f / 2.302585092994046f) / am * inv_n;
    for (int k = lane; k < n; k += 32) {
      const float a = y[base + static_cast<int64_t>(k) * C];
      const float in = a * a;
      const float d_i = (in > eps ? d_l / in : 0.f) + d_am;
      grad_y[base + static_cast<int64_t>(k) * C] = 2.f * a * d_i;
    }
  }
}

// d threshold / d (A, tonality)  (psychoacoustic.py:139-146 and the helpers it calls)
__global__ void __launch_bounds__(kBwdWarps * 32)
pa_threshold_backward_kernel(PaDeviceTables tb, const float* __restrict__ y, const float* __restrict__ ton_in,
                             float one_minus_drown, const float* __restrict__ grad_thr, float* __restrict__ grad_y,
                             float* __restrict__ grad_ton, int64_t items, int C) {
  extern __shared__ float smem[];
  const int n = tb.n, nb = tb.nb;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* inten = smem + warp * (2 * n + 5 * nb);   // [N]  A^2
  float* dv = inten + n;                           // [N]  d loss / d V_k, V = G . W_inv
  float* bark = dv + n;                            // [nb] I_bark
  float* pw = bark + nb;                           // [nb] P = max(eps, I_bark)^alpha
  float* gm = pw + nb;                             // [nb] G = max(masking, quiet)
  float* dm = gm + nb;                             // [nb] d loss / d (sum_i P_i S_ij)
  float* db = dm + nb;                             // [nb] d loss / d I_bark
  const float eps = tb.eps;
  const float ln10 = 2.302585092994046f;
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * kBwdWarps + warp;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kBwdWarps;

  for (int64_t item = warp0; item < items; item += stride) {
    const int64_t r = item / C;
    const int c = static_cast<int>(item - r * C);
    const int64_t base = r * n * C + c;
    const float ton = ton_in[item];

    // ---- forward, as pa_threshold_kernel ------------------------------------------------------------------
    for (int k = lane; k < n; k += 32) {
      const float a = y[base + static_cast<int64_t>(k) * C];
      inten[k] = a * a;
    }
    __syncwarp();
    for (int i = lane; i < nb; i += 32) {
      const int k0 = tb.band_k0[i], cnt = tb.band_cnt[i], ptr = tb.band_ptr[i];
      float acc = 0.f;
      for (int t = 0; t < cnt; ++t) acc = fmaf(inten[k0 + t], tb.band_w[ptr + t], acc);
      bark[i] = acc;
      pw[i] = powf(fmaxf(eps, acc), tb.alpha);
    }
    __syncwarp();
    for (int j = lane; j < nb; j += 32) {
      float acc = 0.f;
      const float* sf = tb.spread_fn + nb + j;          // S[i][j] = spread_fn[nb - i + j]
      for (int i = 0; i < nb; ++i) acc = fmaf(pw[i], sf[-i], acc);
      const float offset = one_minus_drown * ((ton * tb.lin[j] + 9.f * ton) + 5.5f);
      const float gain = powf(10.f, tb.neg_alpha * offset / 10.f);
      const float msk = powf(fmaxf(eps, acc * gain), tb.inv_alpha);
      gm[j] = fmaxf(msk, tb.quiet[j]);
    }
    __syncwarp();

    // ---- backward --------------------------------------------------------------------------------------------
    // thr_k = sqrt(max(eps, V_k)), V_k = sum_j G_j W_inv[j][k]                                    (:330-331)
    for (int k = lane; k < n; k += 32) {
      const int b0 = tb.filt_b0[k], cnt = tb.filt_cnt[k], ptr = tb.filt_ptr[k];
      float v = 0.f;
      for (int t = 0; t < cnt; ++t) v = fmaf(gm[b0 + t], tb.filt_w[ptr + t], v);
      const float g = grad_thr[base + static_cast<int64_t>(k) * C];
      dv[k] = v > eps ? 0.5f * g / sqrtf(v) : 0.f;
    }
    __syncwarp();
    float d_ton = 0.f;
    for (int j = lane; j < nb; j += 32) {
      // d G_j = sum over the filters of band j's support of dV_k W_inv[j][k]
      const int k0 = tb.band_k0[j], cnt = tb.band_cnt[j];
      float d_g = 0.f;
      for (int t = 0; t < cnt; ++t) {
        const int k = k0 + t, s = j - tb.filt_b0[k];
        if (s >= 0 && s < tb.filt_cnt[k]) d_g = fmaf(dv[k], tb.filt_w[tb.filt_ptr[k] + s], d_g);
      }
      // G = max(Mk, quiet), Mk = max(eps, M)^(1/alpha), M = gain * sum_i P_i S_ij                 (:144, :205-208)
      float acc = 0.f;
      const float* sf = tb.spread_fn + nb + j;
      for (int i = 0; i < nb; ++i) acc = fmaf(pw[i], sf[-i], acc);
      const float lin9 = tb.lin[j] + 9.f;
      const float offset = one_minus_drown * (ton * lin9 + 5.5f);
      const float gain = powf(10.f, tb.neg_alpha * offset / 10.f);
      const float m = acc * gain;
      const float mk = powf(fmaxf(eps, m), tb.inv_alpha);
      const float d_mk = mk > tb.quiet[j] ? d_g : 0.f;
      const float d_m = m > eps ? d_mk * tb.inv_alpha * mk / m : 0.f;
      dm[j] = d_m * gain;
      // gain = 10^(-alpha offset / 10), offset = (1 - drown) (ton (lin_j + 9) + 5.5)               (:185-197)
      d_ton += d_m * acc * gain * (tb.neg_alpha * ln10 / 10.f) * one_minus_drown * lin9;
    }
    d_ton = bwd_warp_sum(d_ton);
    if (lane == 0 && grad_ton != nullptr) grad_ton[item] = d_ton;
    __syncwarp();
    for (int i = lane; i < nb; i += 32) {
      // d P_i = sum_j dM'_j S[i][j];  P = max(eps, I_bark)^alpha                                    (:206)
      float d_p = 0.f;
      const float* sf = tb.spread_fn + nb - i;
      for (int j = 0; j < nb; ++j) d_p = fmaf(dm[j], sf[j], d_p);
      db[i] = bark[i] > eps ? d_p * tb.alpha * pw[i] / bark[i] : 0.f;
    }
    __syncwarp();
    for (int k = lane; k < n; k += 32) {
      // d I_k = sum_i dB_i W[k][i];  I = A^2                                                       (:312-313)
      const int b0 = tb.filt_b0[k], cnt = tb.filt_cnt[k];
      float d_i = 0.f;
      for (int t = 0; t < cnt; ++t) {
        const int i = b0 + t, s = k - tb.band_k0[i];
        if (s >= 0 && s < tb.band_cnt[i]) d_i = fmaf(db[i], tb.band_w[tb.band_ptr[i] + s], d_i);
      }
      const float a = y[base + static_cast<int64_t>(k) * C];
      grad_y[base + static_cast<int64_t>(k) * C] = 2.f * a * d_i;
    }
    __syncwarp();
  }
}

int bwd_grid(int64_t items) {
  const int64_t want = (items + kBwdWarps - 1) / kBwdWarps;
  const int64_t cap = 148LL * 8;
  return static_cast<int>(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace

cudaError_t pa_tonality_backward(const PaDeviceTables& tb, const float* y, const float* grad_ton, float* grad_y,
                                 int64_t rows, int channels, cudaStream_t stream) {
  const int64_t items = rows * channels;
  if (items == 0) return cudaSuccess;
  pa_tonality_backward_kernel<<<bwd_grid(items), kBwdWarps * 32, 0, stream>>>(tb, y, grad_ton, grad_y, items, channels);
  count_launch();
  return cudaGetLastError();
}

cudaError_t pa_threshold_backward(const PaDeviceTables& tb, const float* y, const float* ton, float drown,
                                  const float* grad_thr, float* grad_y, float* grad_ton, int64_t rows, int channels,
                                  cudaStream_t stream) {
  const int64_t items = rows * channels;
  if (items == 0) return cudaSuccess;
  const size_t smem = static_cast<size_t>(kBwdWarps) * (2 * tb.n + 5 * tb.nb) * sizeof(float);
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t err = cudaFuncSetAttribute(pa_threshold_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
  if (err != cudaSuccess) return err;
  pa_threshold_backward_kernel<<<bwd_grid(items), kBwdWarps * 32, smem, stream>>>(
      tb, y, ton, static_cast<float>(1.0 - static_cast<double>(drown)), grad_thr, grad_y, grad_ton, items, channels);
  count_launch();
  return cudaGetLastError();
}

}  // namespace ac
