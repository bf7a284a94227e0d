// MDCT analysis / synthesis kernels for sm_100a.
//
// Reference behaviour: /root/reference/audiocodec/mdctransformer.py:61-125 (transform) and :127-153
// (inverse_transform).  The reference evaluates window + time-domain-aliasing fold as a dense 2-tap
// [2, N, N] convolution and the DCT-IV through a zero-interleaved 4N-point FFT; here each frame is
//   fold (4 multiply-adds per sample pair, sparse F)  ->  pre-twiddle  ->  N/2-point complex FFT held in
//   registers (fft_core.cuh)  ->  post-twiddle + scale,
// and the TDAC overlap-add of the inverse is resolved inside the CTA's shared memory (no atomics): a
// CTA walks consecutive frames of one batch row and carries the previous frame across tiles.
//
// DCT-IV by FFT (M = N/2, u the folded block):  z[n] = (u[2n] + i u[N-1-2n]) e^{-i pi (n + 1/8) / N},
//   D[k] = e^{-i pi (k + 1/8) / N} FFT_M(z)[k],   X[2k] = Re D[k],   X[N-1-2k] = -Im D[k].
#include "kernels.h"
#include "fft_core.cuh"

#include <algorithm>
#include <cstdlib>

namespace ac {

namespace {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ------------------------------------------------------------------------------------------ forward
template <typename Plan, int THREADS>
__global__ void __launch_bounds__(THREADS, 2)
mdct_forward_kernel(MdctDeviceTables tb, const float* __restrict__ x, float* __restrict__ y,
                    int blocks_n, int C, int L, int tiles_per_cta, int chunks_per_row) {
  constexpr int M = Plan::M, N = 2 * M, H = M, T = Plan::T, E = Plan::E, G = THREADS / T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int row = N * C;                                     // floats per block / per frame
  float* xin = reinterpret_cast<float*>(smem_raw);           // [L + 1][N][C], slot 0 = block f0 - 1
  float* yout = xin + (L + 1) * row;                         // [L][N][C]
  float2* bufs = reinterpret_cast<float2*>(yout + L * row);  // [G][FftBuf<M>::kSlots]

  const int tid = threadIdx.x, g = tid / T, t = tid % T;
  float2* buf = bufs + g * FftBuf<M>::kSlots;
  const int64_t b = blockIdx.x / chunks_per_row;
  const int chunk = blockIdx.x % chunks_per_row;
  const int frames = blocks_n + 1;
  const float* xb = x + b * static_cast<int64_t>(blocks_n) * row;
  float* yb = y + b * static_cast<int64_t>(frames) * row;

  int f0 = chunk * tiles_per_cta * L;
  // halo: block f0 - 1 -> slot 0 (zeros in front of the signal, mdctransformer.py:366)
  for (int i = tid * 4; i < row; i += THREADS * 4) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (f0 >= 1 && f0 - 1 < blocks_n) v = ld4(xb + static_cast<int64_t>(f0 - 1) * row + i);
    st4(xin + i, v);
  }
  for (int tile = 0; tile < tiles_per_cta; ++tile, f0 += L) {
    if (f0 >= frames) break;
    const int nf = min(L, frames - f0);
    for (int i = tid * 4; i < L * row; i += THREADS * 4) {
      const int blk = f0 + i / row;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (blk < blocks_n) v = ld4(xb + static_cast<int64_t>(f0) * row + i);
      st4(xin + row + i, v);
    }
    __syncthreads();

    for (int item0 = 0; item0 < L * C; item0 += G) {
      int item = item0 + g;
      const bool valid = item < L * C;
      if (!valid) item = 0;
      const int fl = item / C, c = item - fl * C;
      const float* xp = xin + fl * row + c;   // previous block
      const float* xc = xp + row;             // current block
      float2 v[E];
#pragma unroll
      for (int s = 0; s < E; ++s) {
        const int n = Plan::in_index(t, s);
        const bool low = n < N / 4;
        const int p = low ? (H - 1 - 2 * n) : (2 * n - H);
        const float4 a = __ldg(&tb.fold[p]);
        const float alpha = fmaf(a.x, xp[p * C], a.y * xp[(N - 1 - p) * C]);   // delayed half  (H[1])
        const float beta = fmaf(a.z, xc[p * C], a.w * xc[(N - 1 - p) * C]);    // current half  (H[0])
        const float2 z = low ? make_float2(alpha, beta) : make_float2(beta, alpha);
        v[s] = cmul(z, __ldg(&tb.tw_pre[n]));
      }
      fft_forward<Plan>(v, buf, t, tb.roots);
      if (valid) {
        float* yr = yout + fl * row + c;
#pragma unroll
        for (int s = 0; s < E; ++s) {
          const int k = Plan::out_index(t, s);
          const float2 d = cmul(v[s], __ldg(&tb.tw_post_fwd[k]));
          yr[(2 * k) * C] = d.x;
          yr[(N - 1 - 2 * k) * C] = -d.y;
        }
      }
      __syncthreads();
    }

    float* yt = yb + static_cast<int64_t>(f0) * row;
    for (int i = tid * 4; i < nf * row; i += THREADS * 4) st4(yt + i, ld4(yout + i));
    for (int i = tid * 4; i < row; i += THREADS * 4) st4(xin + i, ld4(xin + L * row + i));   // carry
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------ inverse
template <typename Plan, int THREADS, bool DEQUANT>
__global__ void __launch_bounds__(THREADS, 2)
mdct_inverse_kernel(MdctDeviceTables tb, const float* __restrict__ y, const int32_t* __restrict__ q,
                    const float* __restrict__ thr, float* __restrict__ x,
                    int frames_n, int C, int L, int tiles_per_cta, int chunks_per_row) {
  constexpr int M = Plan::M, N = 2 * M, H = M, T = Plan::T, E = Plan::E, G = THREADS / T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int row = N * C;
  float* vbuf = reinterpret_cast<float*>(smem_raw);                 // [L + 1][N][C], slot 0 = frame f0 - 1
  float2* bufs = reinterpret_cast<float2*>(vbuf + (L + 1) * row);   // [G][kSlots]

  const int tid = threadIdx.x, g = tid / T, t = tid % T;
  float2* buf = bufs + g * FftBuf<M>::kSlots;
  const int64_t b = blockIdx.x / chunks_per_row;
  const int chunk = blockIdx.x % chunks_per_row;
  const int64_t in_off = b * static_cast<int64_t>(frames_n) * row;
  float* xb = x + b * static_cast<int64_t>(frames_n + 1) * row;

  // CTA `chunk` transforms frames [f0, f0 + K L) and emits output blocks [first_out, f0 + K L); chunks
  // overlap by one frame so that every block sees both of its frames inside one CTA.
  const int kl = tiles_per_cta * L;
  int f0 = chunk == 0 ? 0 : chunk * (kl - 1);
  const int first_out = chunk == 0 ? 0 : f0 + 1;
  for (int i = tid * 4; i < row; i += THREADS * 4) st4(vbuf + i, make_float4(0.f, 0.f, 0.f, 0.f));

  for (int tile = 0; tile < tiles_per_cta; ++tile, f0 += L) {
    if (f0 > frames_n) break;
    for (int i = tid * 4; i < L * row; i += THREADS * 4) {
      const int fr = f0 + i / row;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (fr < frames_n) {
        const int64_t off = in_off + static_cast<int64_t>(f0) * row + i;
        if constexpr (DEQUANT) {
          const int4 qi = *reinterpret_cast<const int4*>(q + off);
          const float4 th = ld4(thr + off);
          v = make_float4(static_cast<float>(qi.x) * th.x, static_cast<float>(qi.y) * th.y,
                          static_cast<float>(qi.z) * th.z, static_cast<float>(qi.w) * th.w);
        } else {
          v = ld4(y + off);
        }
      }
      st4(vbuf + row + i, v);
    }
    __syncthreads();

    for (int item0 = 0; item0 < L * C; item0 += G) {
      int item = item0 + g;
      const bool valid = item < L * C;
      if (!valid) item = 0;
      const int fl = item / C, c = item - fl * C;
      float* yr = vbuf + (fl + 1) * row + c;
      float2 v[E];
#pragma unroll
      for (int s = 0; s < E; ++s) {
        const int n = Plan::in_index(t, s);
        const float2 z = make_float2(yr[(2 * n) * C], yr[(N - 1 - 2 * n) * C]);
        v[s] = cmul(z, __ldg(&tb.tw_pre[n]));
      }
      if constexpr (Plan::R1 == 1) __syncthreads();   // single-pass plans: separate the reads from the in-place writes
      fft_forward<Plan>(v, buf, t, tb.roots);
      if (valid) {
#pragma unroll
        for (int s = 0; s < E; ++s) {
          const int k = Plan::out_index(t, s);
          const float2 d = cmul(v[s], __ldg(&tb.tw_post_inv[k]));
          yr[(2 * k) * C] = d.x;            // v = sqrt(4N) * DCT-IV(Y)   (mdctransformer.py:145-148)
          yr[(N - 1 - 2 * k) * C] = -d.y;
        }
      }
      __syncthreads();
    }

    // synthesis window + TDAC overlap-add (H_inv, mdctransformer.py:148,176-190): block n takes the lower
    // half of v_n and the upper half of v_{n-1}.
    const int pairs = H * C;
    for (int idx = tid; idx < L * pairs; idx += THREADS) {
      const int bl = idx / pairs, rem = idx - bl * pairs;
      const int p = rem / C, c = rem - p * C;
      const int blk = f0 + bl;
      if (blk >= first_out && blk <= frames_n) {
        const float vn = vbuf[((bl + 1) * N + (H - 1 - p)) * C + c];
        const float vp = vbuf[(bl * N + (H + p)) * C + c];
        const float4 s = __ldg(&tb.unfold[p]);
        float* xo = xb + static_cast<int64_t>(blk) * row;
        xo[p * C + c] = fmaf(s.x, vn, s.y * vp);
        xo[(N - 1 - p) * C + c] = fmaf(s.z, vn, s.w * vp);
      }
    }
    __syncthreads();
    for (int i = tid * 4; i < row; i += THREADS * 4) st4(vbuf + i, ld4(vbuf + L * row + i));   // carry
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------- generic N (any even N)
// Direct O(N^2) DCT-IV from a cos(pi m / 4N) table; used when N is not a power of two in [16, 4096].
__global__ void mdct_forward_generic_kernel(MdctDeviceTables tb, const float* __restrict__ x, float* __restrict__ y,
                                            int blocks_n, int C) {
  extern __shared__ float u[];   // [N]
  const int n = tb.n, h = n / 2;
  const int frames = blocks_n + 1;
  const int64_t bf = blockIdx.x;                 // b * frames + f
  const int64_t b = bf / frames;
  const int f = static_cast<int>(bf - b * frames);
  const int c = blockIdx.y;
  const float* xb = x + b * static_cast<int64_t>(blocks_n) * n * C + c;
  for (int p = threadIdx.x; p < h; p += blockDim.x) {
    const float4 a = tb.fold[p];
    float xp0 = 0.f, xp1 = 0.f, xc0 = 0.f, xc1 = 0.f;
    if (f >= 1) {
      xp0 = xb[(static_cast<int64_t>(f - 1) * n + p) * C];
      xp1 = xb[(static_cast<int64_t>(f - 1) * n + n - 1 - p) * C];
    }
    if (f < blocks_n) {
      xc0 = xb[(static_cast<int64_t>(f) * n + p) * C];
      xc1 = xb[(static_cast<int64_t>(f) * n + n - 1 - p) * C];
    }
    u[h - 1 - p] = fmaf(a.x, xp0, a.y * xp1);
    u[h + p] = fmaf(a.z, xc0, a.w * xc1);
  }
  __syncthreads();
  const int period = 8 * n;
  float* yr = y + (bf * n) * C + c;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    float acc = 0.f;
    int idx = (2 * k + 1) % period;            // (2j+1)(2k+1) mod 8N, stepped by 2(2k+1)
    const int step = (2 * (2 * k + 1)) % period;
    for (int j = 0; j < n; ++j) {
      acc = fmaf(u[j], __ldg(&tb.cos_table[idx]), acc);
      idx += step;
      if (idx >= period) idx -= period;
    }
    yr[static_cast<int64_t>(k) * C] = acc * tb.scale_fwd;
  }
}

template <bool DEQUANT>
__global__ void mdct_inverse_generic_kernel(MdctDeviceTables tb, const float* __restrict__ y,
                                            const int32_t* __restrict__ q, const float* __restrict__ thr,
                                            float* __restrict__ x, int frames_n, int C) {
  extern __shared__ float sm[];   // yn[N], yp[N], vn_low[h], vp_high[h]
  const int n = tb.n, h = n / 2;
  float* yn = sm;
  float* yp = sm + n;
  float* vlow = sm + 2 * n;
  float* vhigh = vlow + h;
  const int out_blocks = frames_n + 1;
  const int64_t bb = blockIdx.x;
  const int64_t b = bb / out_blocks;
  const int blk = static_cast<int>(bb - b * out_blocks);
  const int c = blockIdx.y;
  const int64_t base = b * static_cast<int64_t>(frames_n) * n * C + c;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    float a = 0.f, bprev = 0.f;
    if (blk < frames_n) {
      const int64_t off = base + (static_cast<int64_t>(blk) * n + k) * C;
      a = DEQUANT ? static_cast<float>(q[off]) * thr[off] : y[off];
    }
    if (blk >= 1) {
      const int64_t off = base + (static_cast<int64_t>(blk - 1) * n + k) * C;
      bprev = DEQUANT ? static_cast<float>(q[off]) * thr[off] : y[off];
    }
    yn[k] = a;
    yp[k] = bprev;
  }
  __syncthreads();
  const int period = 8 * n;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const float* src = k < h ? yn : yp;     // lower half of v_n, upper half of v_{n-1}
    float acc = 0.f;
    int idx = (2 * k + 1) % period;
    const int step = (2 * (2 * k + 1)) % period;
    for (int j = 0; j < n; ++j) {
      acc = fmaf(src[j], __ldg(&tb.cos_table[idx]), acc);
      idx += step;
      if (idx >= period) idx -= period;
    }
    if (k < h) vlow[k] = acc * tb.scale_inv; else vhigh[k - h] = acc * tb.scale_inv;
  }
  __syncthreads();
  float* xo = x + (b * out_blocks + blk) * static_cast<int64_t>(n) * C + c;
  for (int p = threadIdx.x; p < h; p += blockDim.x) {
    const float4 s = tb.unfold[p];
    const float vn = vlow[h - 1 - p], vp = vhigh[p];
    xo[static_cast<int64_t>(p) * C] = fmaf(s.x, vn, s.y * vp);
    xo[static_cast<int64_t>(n - 1 - p) * C] = fmaf(s.z, vn, s.w * vp);
  }
}

// ------------------------------------------------------------------------------------------ launchers
struct TileShape {
  int L, tiles_per_cta, chunks;
  size_t smem;
};

template <typename Plan, int THREADS>
TileShape choose_tiles(int units, int C, bool forward) {
  constexpr int M = Plan::M, N = 2 * M, G = THREADS / Plan::T;
  TileShape ts;
  ts.L = std::max(1, G / C);
  const size_t row = static_cast<size_t>(N) * C * sizeof(float);
  const size_t bufs = static_cast<size_t>(G) * FftBuf<M>::kSlots * sizeof(float2);
  ts.smem = forward ? (2 * ts.L + 1) * row + bufs : (ts.L + 1) * row + bufs;
  // ~8 tiles per CTA keeps the halo (one extra block / frame per CTA) around 1 %, but never fewer than 2 frames
  ts.tiles_per_cta = std::max(8, (2 + ts.L - 1) / ts.L);
  const int kl = ts.tiles_per_cta * ts.L;
  if (forward) {
    ts.chunks = (units + kl - 1) / kl;                                     // units = frames
  } else {
    ts.chunks = 1 + (units > kl ? (units - kl + (kl - 2)) / (kl - 1) : 0);   // units = output blocks
  }
  return ts;
}

template <typename Plan, int THREADS>
cudaError_t launch_forward(const MdctDeviceTables& tb, const float* x, float* y, int64_t batches, int blocks_n,
                           int C, cudaStream_t stream) {
  const TileShape ts = choose_tiles<Plan, THREADS>(blocks_n + 1, C, true);
  auto kernel = mdct_forward_kernel<Plan, THREADS>;
  if (ts.smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(ts.smem));
  if (err != cudaSuccess) return err;
  const int64_t grid = batches * ts.chunks;
  if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
  kernel<<<static_cast<unsigned>(grid), THREADS, ts.smem, stream>>>(tb, x, y, blocks_n, C, ts.L, ts.tiles_per_cta, ts.chunks);
  count_launch();
  return cudaGetLastError();
}

template <typename Plan, int THREADS>
cudaError_t launch_inverse(const MdctDeviceTables& tb, const float* y, const int32_t* q, const float* thr, float* x,
                           int64_t batches, int frames_n, int C, cudaStream_t stream) {
  const TileShape ts = choose_tiles<Plan, THREADS>(frames_n + 1, C, false);
  if (ts.smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  const int64_t grid = batches * ts.chunks;
  if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
  cudaError_t err;
  if (q != nullptr) {
    auto kernel = mdct_inverse_kernel<Plan, THREADS, true>;
    err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(ts.smem));
    if (err != cudaSuccess) return err;
    kernel<<<static_cast<unsigned>(grid), THREADS, ts.smem, stream>>>(tb, y, q, thr, x, frames_n, C, ts.L, ts.tiles_per_cta, ts.chunks);
  } else {
    auto kernel = mdct_inverse_kernel<Plan, THREADS, false>;
    err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(ts.smem));
    if (err != cudaSuccess) return err;
    kernel<<<static_cast<unsigned>(grid), THREADS, ts.smem, stream>>>(tb, y, q, thr, x, frames_n, C, ts.L, ts.tiles_per_cta, ts.chunks);
  }
  count_launch();
  return cudaGetLastError();
}

//                      M    E  R0  R1 R2
using Plan16 = FftPlan<8, 8, 8, 1, 1>;
using Plan32 = FftPlan<16, 8, 8, 2, 1>;
using Plan64 = FftPlan<32, 8, 8, 4, 1>;
using Plan128 = FftPlan<64, 8, 8, 8, 1>;
using Plan256 = FftPlan<128, 16, 16, 8, 1>;
using Plan512 = FftPlan<256, 16, 16, 16, 1>;
using Plan1024 = FftPlan<512, 16, 8, 8, 8>;
using Plan2048 = FftPlan<1024, 16, 16, 8, 8>;
using Plan4096 = FftPlan<2048, 16, 16, 16, 8>;

}  // namespace

// AC_MDCT_LEGACY=1 routes 1- and 2-channel signals through the any-channel-count kernels of this file too
// (A/B measurements and tests of that path)
static bool legacy_path() {
  static const bool legacy = [] {
    const char* e = std::getenv("AC_MDCT_LEGACY");
    return e != nullptr && e[0] == '1';
  }();
  return legacy;
}

bool mdct_has_fast_path(int n) {
  return n == 16 || n == 32 || n == 64 || n == 128 || n == 256 || n == 512 || n == 1024 || n == 2048 || n == 4096;
}

cudaError_t mdct_forward(const MdctDeviceTables& tb, const float* x, float* y, int64_t batches, int64_t blocks_n,
                         int C, cudaStream_t stream) {
  if (batches == 0) return cudaSuccess;
  if (!legacy_path() && mdct_tile_forward_supported(tb.n, C)) return mdct_forward_tile(tb, x, y, batches, blocks_n, C, stream);
  const int bn = static_cast<int>(blocks_n);
  cudaError_t tiled = cudaErrorInvalidConfiguration;   // a tile that does not fit in shared memory: generic kernel below
  switch (tb.n) {
    case 16: tiled = launch_forward<Plan16, 128>(tb, x, y, batches, bn, C, stream); break;
    case 32: tiled = launch_forward<Plan32, 128>(tb, x, y, batches, bn, C, stream); break;
    case 64: tiled = launch_forward<Plan64, 256>(tb, x, y, batches, bn, C, stream); break;
    case 128: tiled = launch_forward<Plan128, 256>(tb, x, y, batches, bn, C, stream); break;
    case 256: tiled = launch_forward<Plan256, 256>(tb, x, y, batches, bn, C, stream); break;
    case 512: tiled = launch_forward<Plan512, 256>(tb, x, y, batches, bn, C, stream); break;
    case 1024: tiled = launch_forward<Plan1024, 256>(tb, x, y, batches, bn, C, stream); break;
    case 2048: tiled = launch_forward<Plan2048, 256>(tb, x, y, batches, bn, C, stream); break;
    case 4096: tiled = launch_forward<Plan4096, 256>(tb, x, y, batches, bn, C, stream); break;
    default: break;
  }
  if (tiled != cudaErrorInvalidConfiguration) return tiled;
  const int64_t rows = batches * (blocks_n + 1);
  if (rows > 2147483647LL || C > 65535) return cudaErrorInvalidConfiguration;
  dim3 grid(static_cast<unsigned>(rows), static_cast<unsigned>(C));
  mdct_forward_generic_kernel<<<grid, 128, tb.n * sizeof(float), stream>>>(tb, x, y, bn, C);
  count_launch();
  return cudaGetLastError();
}

cudaError_t mdct_inverse(const MdctDeviceTables& tb, const float* y, const int32_t* q, const float* thr, float* x,
                         int64_t batches, int64_t frames_n, int C, cudaStream_t stream) {
  if (batches == 0) return cudaSuccess;
  if (!legacy_path() && mdct_tile_inverse_supported(tb.n, C))
    return mdct_inverse_tile(tb, y, q, thr, x, batches, frames_n, C, stream);
  const int fn = static_cast<int>(frames_n);
  cudaError_t tiled = cudaErrorInvalidConfiguration;
  switch (tb.n) {
    case 16: tiled = launch_inverse<Plan16, 128>(tb, y, q, thr, x, batches, fn, C, stream); break;
    case 32: tiled = launch_inverse<Plan32, 128>(tb, y, q, thr, x, batches, fn, C, stream); break;
    case 64: tiled = launch_inverse<Plan64, 256>(tb, y, q, thr, x, batches, fn, C, stream); break;
    case 128: tiled = launch_inverse<Plan128, 256>(tb, y, q, thr, x, batches, fn, C, stream); break;
    case 256: tiled = launch_inverse<Plan256, 256>(tb, y, q, thr, x, batches, fn, C, stream); break;
    case 512: tiled = launch_inverse<Plan512, 256>(tb, y, q, thr, x, batches, fn, C, stream); break;
    case 1024: tiled = launch_inverse<Plan1024, 256>(tb, y, q, thr, x, batches, fn, C, stream); break;
    case 2048: tiled = launch_inverse<Plan2048, 256>(tb, y, q, thr, x, batches, fn, C, stream); break;
    case 4096: tiled = launch_inverse<Plan4096, 256>(tb, y, q, thr, x, batches, fn, C, stream); break;
    default: break;
  }
  if (tiled != cudaErrorInvalidConfiguration) return tiled;
  const int64_t rows = batches * (frames_n + 1);
  if (rows > 2147483647LL || C > 65535) return cudaErrorInvalidConfiguration;
  dim3 grid(static_cast<unsigned>(rows), static_cast<unsigned>(C));
  const size_t smem = 3 * static_cast<size_t>(tb.n) * sizeof(float);   // above 48 KB for filters_n > 4096: opt in
  cudaError_t err;
  if (q != nullptr) {
    err = cudaFuncSetAttribute(mdct_inverse_generic_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem));
    if (err != cudaSuccess) return err;
    mdct_inverse_generic_kernel<true><<<grid, 128, smem, stream>>>(tb, y, q, thr, x, fn, C);
  } else {
    err = cudaFuncSetAttribute(mdct_inverse_generic_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem));
    if (err != cudaSuccess) return err;
    mdct_inverse_generic_kernel<false><<<grid, 128, smem, stream>>>(tb, y, q, thr, x, fn, C);
  }
  count_launch();
  return cudaGetLastError();
}

}  // namespace ac
