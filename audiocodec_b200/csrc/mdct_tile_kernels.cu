// MDCT analysis / synthesis for 1 or 2 channels: tile kernels with bulk-async staging (sm_100a).
//
// Reference behaviour: /root/reference/audiocodec/mdctransformer.py:61-125 (transform), :127-153
// (inverse_transform); same mathematics as mdct_kernels.cu (fold -> pre-twiddle -> N/2-point complex FFT ->
// post-twiddle), reorganised around what bounds it on a B200: instruction issue and shared-memory wavefronts.
//
//   * a tile (a run of consecutive frames of one batch row) moves between HBM and shared memory with ONE
//     cp.async.bulk each way (TMA 1-D, mbarrier completion): no load / store instructions for global data;
//   * every thread transforms TWO sequences at once - the two channels of a stereo frame (float2 shared
//     loads), or two adjacent frames of a mono signal - so tables, twiddles and index arithmetic are shared;
//   * fold + pre-twiddle are one 4-term dot product per component with host-merged coefficients, the
//     post-twiddle writes its two outputs straight into the transposed positions;
//   * a frame's shared-memory row is reused in place: input block -> FFT exchange scratch (XOR-swizzled
//     instead of padded) -> output frame, so a tile needs (frames + 1) rows and several CTAs fit per SM;
//   * the two quarter-warps of a half-warp make their two strided accesses in opposite order ("variant"), which
//     puts them on complementary bank classes: the stride-2 access pattern of the DCT-IV costs no conflicts;
//   * the FFT passes of a group synchronise with __syncwarp (groups of up to a warp) or a named barrier.
//
// tools/emulate_mdct_tile.py is a NumPy emulation of exactly this index arithmetic (development aid).
#include "kernels.h"
#include "fft_core.cuh"
#include "async_copy.cuh"
#include "mdct_tile_core.cuh"

#include <algorithm>
#include <cstdlib>

namespace ac {

namespace {

// ------------------------------------------------------------------------------------------ forward
// Two tile buffers: the bulk load of tile i + 1 is in flight while tile i is transformed.
template <typename Plan, int C, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
mdct_forward_tile_kernel(MdctDeviceTables tb, const float* __restrict__ x, float* __restrict__ y, int blocks_n,
                         int tiles_per_row, int64_t total_tiles) {
  using S = TileShape<Plan, C, THREADS>;
  constexpr int M = S::M, N = S::N, H = M, T = S::T, E = Plan::E, FP = S::FP, ROW = S::ROW, R0 = Plan::R0;
  constexpr int BUF = (FP + 1) * ROW;                                 // floats per tile buffer
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* bufs = reinterpret_cast<float*>(smem_raw);                   // [2][FP + 1][ROW]: row r = block f0 - 1 + r
  uint64_t* mbar = reinterpret_cast<uint64_t*>(bufs + 2 * BUF);       // [2]

  const int tid = threadIdx.x, g = tid / T, t = tid % T, variant = (tid >> 3) & 1;
  const int frames = blocks_n + 1;
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  pdl_launch_dependents();                  // the next kernel on the stream may start its prologue
  pdl_wait();                               // the kernel before this one has completed: its tensors are visible

  // rows r in [r_lo, r_hi) of a tile hold real blocks, the others are the zero padding (mdctransformer.py:366)
  auto issue_load = [&](int64_t tile, int slot) {        // thread 0 only
    const int64_t b = tile / tiles_per_row;
    const int f0 = static_cast<int>(tile - b * tiles_per_row) * FP;
    const int r_lo = f0 == 0 ? 1 : 0;
    const int r_hi = min(FP + 1, blocks_n - f0 + 1);
    if (r_hi > r_lo) {
      const uint32_t bytes = static_cast<uint32_t>(r_hi - r_lo) * ROW * sizeof(float);
      mbar_arrive_expect_tx(&mbar[slot], bytes);
      bulk_load(bufs + slot * BUF + r_lo * ROW, x + (b * blocks_n + (f0 - 1 + r_lo)) * static_cast<int64_t>(ROW), bytes,
                &mbar[slot]);
    } else {
      mbar_arrive(&mbar[slot]);
    }
  };
  if (tid == 0 && blockIdx.x < total_tiles) issue_load(blockIdx.x, 0);

  int it = 0;
  for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
    const int slot = it & 1;
    float* buf = bufs + slot * BUF;
    float* prev = buf + g * (2 / C) * ROW;      // block row before this group's (first) frame
    float* cur = prev + ROW;                    // this group's frame row: input block, FFT scratch, output frame
    const int64_t b = tile / tiles_per_row;
    const int f0 = static_cast<int>(tile - b * tiles_per_row) * FP;
    const int r_lo = f0 == 0 ? 1 : 0;
    const int r_hi = min(FP + 1, blocks_n - f0 + 1);
    mbar_wait(&mbar[slot], (it >> 1) & 1);
    if (r_lo > 0 || r_hi < FP + 1) {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r_lo > 0)
        for (int i = tid * 4; i < ROW; i += THREADS * 4) *reinterpret_cast<float4*>(buf + i) = z;
      for (int i = max(r_hi, r_lo) * ROW + tid * 4; i < BUF; i += THREADS * 4) *reinterpret_cast<float4*>(buf + i) = z;
      __syncthreads();
    }

    // ---- window + fold + pre-twiddle: one 4-term dot product per component        (mdctransformer.py:118, H)
    float2 v0[E], v1[E];
#pragma unroll
    for (int s = 0; s < E; ++s) {
      const int n = Plan::in_index(t, s);
      constexpr int kHalf = R0 / 2;
      const bool low = (s % R0) < kHalf;      // n < N/4: the pair comes from odd positions
      const int p = low ? H - 1 - 2 * n : 2 * n - H;
      const int a1 = variant ? N - 1 - p : p;
      const int a2 = (N - 1) - a1;
      const float2 l0 = ld2<C, ROW>(prev, a1), l1 = ld2<C, ROW>(prev, a2);
      const float2 l2 = ld2<C, ROW>(cur, a1), l3 = ld2<C, ROW>(cur, a2);
      const float4 kr = __ldg(&tb.pre_fwd[(variant * 2) * M + n]);
      const float4 ki = __ldg(&tb.pre_fwd[(variant * 2 + 1) * M + n]);
      v0[s].x = fmaf(l3.x, kr.w, fmaf(l2.x, kr.z, fmaf(l1.x, kr.y, l0.x * kr.x)));
      v0[s].y = fmaf(l3.x, ki.w, fmaf(l2.x, ki.z, fmaf(l1.x, ki.y, l0.x * ki.x)));
      v1[s].x = fmaf(l3.y, kr.w, fmaf(l2.y, kr.z, fmaf(l1.y, kr.y, l0.y * kr.x)));
      v1[s].y = fmaf(l3.y, ki.w, fmaf(l2.y, ki.z, fmaf(l1.y, ki.y, l0.y * ki.x)));
    }
    __syncthreads();                          // every block row has been read: rows become scratch / output
    if (tid == 0 && tile + gridDim.x < total_tiles) {
      bulk_wait_read<0>();                    // the store that last read the other buffer has drained it
      issue_load(tile + gridDim.x, slot ^ 1);
    }

    fft2<Plan>(v0, v1, reinterpret_cast<float4*>(cur), t, g, tb.tw_pass1, tb.tw_pass2);
    post_store<Plan, C, ROW>(v0, v1, cur, t, variant, tb.post_fwd);
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      const int nf = min(FP, frames - f0);
      bulk_store(y + (b * frames + f0) * static_cast<int64_t>(ROW), buf + ROW, static_cast<uint32_t>(nf) * ROW * sizeof(float));
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait<0>();
}

// ------------------------------------------------------------------------------------------ inverse
// A tile transforms FP consecutive frames (the first one is the halo: frame nb0 - 1) and emits the FP - 1
// output blocks nb0 .. nb0 + FP - 2, each of which needs two adjacent frames (TDAC overlap-add, H_inv).
//
// Dequantising inverse, fully overlapped loads: the steps are double-buffered, the integers have one buffer that
// is re-armed as soon as every thread has consumed them, and the output goes straight to global memory.  96 KB of
// shared memory per CTA at N = 256 (two CTAs per SM): the kernel is nowhere near issue-bound, so eight warps are
// enough, and no CTA waits for a tile (single-buffered, 38 % of the stall samples sat in the mbarrier wait).
template <typename Plan, int C, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
mdct_inverse_dequant_tile_kernel(MdctDeviceTables tb, const int32_t* __restrict__ q, const float* __restrict__ thr,
                                 float* __restrict__ x, int frames_n, int tiles_per_row, int64_t total_tiles) {
  using S = TileShape<Plan, C, THREADS>;
  constexpr int M = S::M, N = S::N, H = M, T = S::T, E = Plan::E, FP = S::FP, ROW = S::ROW;
  constexpr int BUF = FP * ROW;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tbufs = reinterpret_cast<float*>(smem_raw);                 // [2][FP][ROW]: steps -> scratch -> v
  float* qbuf = tbufs + 2 * BUF;                                     // [FP][ROW]: quantised integers
  uint64_t* mbar = reinterpret_cast<uint64_t*>(qbuf + BUF);          // [0..1] steps, [2] integers

  const int tid = threadIdx.x, g = tid / T, t = tid % T, variant = (tid >> 3) & 1;
  const int32_t* qrow = reinterpret_cast<const int32_t*>(qbuf) + g * (2 / C) * ROW;
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_init(&mbar[2], 1);
    mbar_fence_init();
  }
  __syncthreads();
  pdl_launch_dependents();                  // the next kernel on the stream may start its prologue
  pdl_wait();                               // the kernel before this one has completed: its tensors are visible

  auto issue_load = [&](int64_t tile, float* dst, const void* src, uint64_t* bar) {        // thread 0 only
    const int64_t b = tile / tiles_per_row;
    const int fs = static_cast<int>(tile - b * tiles_per_row) * (FP - 1) - 1;
    const int r_lo = fs < 0 ? 1 : 0;
    const int r_hi = min(FP, frames_n - fs);
    if (r_hi > r_lo) {
      const uint32_t bytes = static_cast<uint32_t>(r_hi - r_lo) * ROW * sizeof(float);
      mbar_arrive_expect_tx(bar, bytes);
      bulk_load(dst + r_lo * ROW, static_cast<const float*>(src) + (b * frames_n + (fs + r_lo)) * static_cast<int64_t>(ROW), bytes, bar);
    } else {
      mbar_arrive(bar);
    }
  };
  if (tid == 0 && blockIdx.x < total_tiles) {
    issue_load(blockIdx.x, tbufs, thr, &mbar[0]);
    issue_load(blockIdx.x, qbuf, q, &mbar[2]);
  }

  int it = 0;
  for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
    const int slot = it & 1;
    float* abuf = tbufs + slot * BUF;
    float* arow = abuf + g * (2 / C) * ROW;
    const int64_t b = tile / tiles_per_row;
    const int nb0 = static_cast<int>(tile - b * tiles_per_row) * (FP - 1);   // first output block
    const int fs = nb0 - 1;
    const int r_lo = fs < 0 ? 1 : 0;
    const int r_hi = min(FP, frames_n - fs);
    const bool more = tile + gridDim.x < total_tiles;
    // the other step buffer was last read by the overlap-add of the previous tile, which ended in a barrier
    if (tid == 0 && more) issue_load(tile + gridDim.x, tbufs + (slot ^ 1) * BUF, thr, &mbar[slot ^ 1]);
    mbar_wait(&mbar[slot], (it >> 1) & 1);
    mbar_wait(&mbar[2], it & 1);
    if (r_lo > 0 || r_hi < FP) {               // frames outside the signal are zero (mdctransformer.py:366)
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r_lo > 0)
        for (int i = tid * 4; i < ROW; i += THREADS * 4) {
          *reinterpret_cast<float4*>(abuf + i) = z;
          *reinterpret_cast<float4*>(qbuf + i) = z;
        }
      for (int i = max(r_hi, r_lo) * ROW + tid * 4; i < BUF; i += THREADS * 4) {
        *reinterpret_cast<float4*>(abuf + i) = z;
        *reinterpret_cast<float4*>(qbuf + i) = z;
      }
      __syncthreads();
    }

    // ---- dequantise, pre-twiddle                                                   (mdctransformer.py:141-148)
    float2 v0[E], v1[E];
#pragma unroll
    for (int s = 0; s < E; ++s) {
      const int n = Plan::in_index(t, s);
      const int a1 = variant ? N - 1 - 2 * n : 2 * n;
      const int a2 = (N - 1) - a1;
      float2 l1 = ld2<C, ROW>(arow, a1), l2 = ld2<C, ROW>(arow, a2);
      const float2 q1 = ld2<C, ROW>(reinterpret_cast<const float*>(qrow), a1);
      const float2 q2 = ld2<C, ROW>(reinterpret_cast<const float*>(qrow), a2);
      l1.x *= static_cast<float>(__float_as_int(q1.x));
      l1.y *= static_cast<float>(__float_as_int(q1.y));
      l2.x *= static_cast<float>(__float_as_int(q2.x));
      l2.y *= static_cast<float>(__float_as_int(q2.y));
      const float4 k4 = __ldg(&tb.pre_inv[variant * M + n]);
      v0[s] = make_float2(fmaf(l2.x, k4.y, l1.x * k4.x), fmaf(l2.x, k4.w, l1.x * k4.z));
      v1[s] = make_float2(fmaf(l2.y, k4.y, l1.y * k4.x), fmaf(l2.y, k4.w, l1.y * k4.z));
    }
    __syncthreads();                           // integers consumed by everybody: their buffer takes the next tile
    if (tid == 0 && more) issue_load(tile + gridDim.x, qbuf, q, &mbar[2]);

    fft2<Plan>(v0, v1, reinterpret_cast<float4*>(arow), t, g, tb.tw_pass1, tb.tw_pass2);
    post_store<Plan, C, ROW>(v0, v1, arow, t, variant, tb.post_inv);
    __syncthreads();

    // ---- synthesis window + TDAC overlap-add (H_inv, mdctransformer.py:148,176-190), straight to global memory
    const int nblk = min(FP - 1, frames_n + 1 - nb0);
    float* xb = x + (b * (frames_n + 1) + nb0) * static_cast<int64_t>(ROW);
    for (int idx = tid; idx < nblk * H; idx += THREADS) {
      const int bl = idx / H, p = idx % H;
      const float4 s = __ldg(&tb.unfold[p]);
      const float* vn = abuf + (bl + 1) * ROW + (H - 1 - p) * C;
      const float* vp = abuf + bl * ROW + (H + p) * C;
      float* xo = xb + static_cast<int64_t>(bl) * ROW;
      if constexpr (C == 2) {
        const float2 a = *reinterpret_cast<const float2*>(vn), c = *reinterpret_cast<const float2*>(vp);
        *reinterpret_cast<float2*>(xo + 2 * p) = make_float2(fmaf(s.x, a.x, s.y * c.x), fmaf(s.x, a.y, s.y * c.y));
        *reinterpret_cast<float2*>(xo + 2 * (N - 1 - p)) = make_float2(fmaf(s.z, a.x, s.w * c.x), fmaf(s.z, a.y, s.w * c.y));
      } else {
        xo[p] = fmaf(s.x, vn[0], s.y * vp[0]);
        xo[N - 1 - p] = fmaf(s.z, vn[0], s.w * vp[0]);
      }
    }
    __syncthreads();       // this step buffer is the target of the bulk load issued at the top of the next iteration
  }
}

// Dequantising inverse on the COMPACT side information: instead of one step per coefficient the
// kernel loads the 64 bark-domain thresholds of each (frame, channel) - bark [frames][64][C], 1 / (N / 64) of the
// bytes - and rebuilds the step of coefficient k as sqrt(G[b] w0 + G[b + 1] w1 + G[b + 2] w2) from the filter table
// of the masking model (filt4[k] = { w0, w1, w2, b }, psychoacoustic.py:330-331), with the operations of phase D of
// the masking kernel (psycho_mma_kernels.cu): the same bits as the step the encoder divided by.  One scratch tile,
// one tile of integers and one small G tile, each re-armed for the next tile as soon as every thread has consumed it
// (72 KB at N = 256: three CTAs per SM).
__device__ __forceinline__ float rsqrt_approx_ftz(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

template <typename Plan, int C, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
mdct_inverse_dequant_compact_tile_kernel(MdctDeviceTables tb, const int32_t* __restrict__ q,
                                         const float* __restrict__ bark, const float4* __restrict__ filt4,
                                         const float eps_s2, float* __restrict__ x, int frames_n, int tiles_per_row,
                                         int64_t total_tiles) {
  using S = TileShape<Plan, C, THREADS>;
  constexpr int M = S::M, N = S::N, H = M, T = S::T, E = Plan::E, FP = S::FP, ROW = S::ROW;
  constexpr int BUF = FP * ROW, GROW = 64 * C, GBUF = FP * GROW;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* abuf = reinterpret_cast<float*>(smem_raw);                  // [FP][ROW]: scratch -> v
  float* qbuf = abuf + BUF;                                          // [FP][ROW]: quantised integers
  float* gbuf = qbuf + BUF;                                          // [FP][64][C]: bark-domain thresholds
  uint64_t* mbar = reinterpret_cast<uint64_t*>(gbuf + GBUF);         // [0] thresholds, [1] integers

  const int tid = threadIdx.x, g = tid / T, t = tid % T, variant = (tid >> 3) & 1;
  float* arow = abuf + g * (2 / C) * ROW;
  const int32_t* qrow = reinterpret_cast<const int32_t*>(qbuf) + g * (2 / C) * ROW;
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  pdl_launch_dependents();                  // the next kernel on the stream may start its prologue
  pdl_wait();                               // the kernel before this one has completed: its tensors are visible

  auto issue_load = [&](int64_t tile, float* dst, const void* src, int row_floats, uint64_t* bar) {   // thread 0 only
    const int64_t b = tile / tiles_per_row;
    const int fs = static_cast<int>(tile - b * tiles_per_row) * (FP - 1) - 1;
    const int r_lo = fs < 0 ? 1 : 0;
    const int r_hi = min(FP, frames_n - fs);
    if (r_hi > r_lo) {
      const uint32_t bytes = static_cast<uint32_t>(r_hi - r_lo) * row_floats * sizeof(float);
      mbar_arrive_expect_tx(bar, bytes);
      bulk_load(dst + r_lo * row_floats,
                static_cast<const float*>(src) + (b * frames_n + (fs + r_lo)) * static_cast<int64_t>(row_floats), bytes, bar);
    } else {
      mbar_arrive(bar);
    }
  };
  if (tid == 0 && blockIdx.x < total_tiles) {
    issue_load(blockIdx.x, gbuf, bark, GROW, &mbar[0]);
    issue_load(blockIdx.x, qbuf, q, ROW, &mbar[1]);
  }

  int it = 0;
  for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
    const int64_t b = tile / tiles_per_row;
    const int nb0 = static_cast<int>(tile - b * tiles_per_row) * (FP - 1);   // first output block
    const int fs = nb0 - 1;
    const int r_lo = fs < 0 ? 1 : 0;
    const int r_hi = min(FP, frames_n - fs);
    const bool more = tile + gridDim.x < total_tiles;
    mbar_wait(&mbar[0], it & 1);
    mbar_wait(&mbar[1], it & 1);
    if (r_lo > 0 || r_hi < FP) {               // frames outside the signal are zero (mdctransformer.py:366)
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r_lo > 0) {
        for (int i = tid * 4; i < ROW; i += THREADS * 4) *reinterpret_cast<float4*>(qbuf + i) = z;
        for (int i = tid * 4; i < GROW; i += THREADS * 4) *reinterpret_cast<float4*>(gbuf + i) = z;
      }
      for (int i = max(r_hi, r_lo) * ROW + tid * 4; i < BUF; i += THREADS * 4) *reinterpret_cast<float4*>(qbuf + i) = z;
      for (int i = max(r_hi, r_lo) * GROW + tid * 4; i < GBUF; i += THREADS * 4) *reinterpret_cast<float4*>(gbuf + i) = z;
      __syncthreads();
    }

    // ---- the quantiser steps of the tile: abuf[frame][k][c] = sqrt(G W_inv), thread <-> filter k (its table entry
    // stays in registers for all frames; neighbouring filters read the same bands of G: shared-memory broadcasts)
    {
      static_assert(N % THREADS == 0, "filters per thread");
#pragma unroll
      for (int kk = 0; kk < N / THREADS; ++kk) {
        const int k = tid + kk * THREADS;
        const float4 f4 = __ldg(filt4 + k);
        const float* gk = gbuf + __float_as_int(f4.w) * C;
#pragma unroll 4
        for (int f = 0; f < FP; ++f) {
          const float* gr = gk + f * GROW;
          if constexpr (C == 2) {
            const float2 g0 = *reinterpret_cast<const float2*>(gr), g1 = *reinterpret_cast<const float2*>(gr + 2);
            const float2 g2 = *reinterpret_cast<const float2*>(gr + 4);
            const float vx = fmaxf(eps_s2, fmaf(g2.x, f4.z, fmaf(g1.x, f4.y, g0.x * f4.x)));
            const float vy = fmaxf(eps_s2, fmaf(g2.y, f4.z, fmaf(g1.y, f4.y, g0.y * f4.x)));
            *reinterpret_cast<float2*>(abuf + f * ROW + 2 * k) =
                make_float2(vx * rsqrt_approx_ftz(vx), vy * rsqrt_approx_ftz(vy));
          } else {
            const float v = fmaxf(eps_s2, fmaf(gr[2], f4.z, fmaf(gr[1], f4.y, gr[0] * f4.x)));
            abuf[f * ROW + k] = v * rsqrt_approx_ftz(v);
          }
        }
      }
      __syncthreads();                         // thresholds consumed by everybody: their buffer takes the next tile
      if (tid == 0 && more) issue_load(tile + gridDim.x, gbuf, bark, GROW, &mbar[0]);
    }

    // ---- dequantise, pre-twiddle                               (mdctransformer.py:141-148)
    float2 v0[E], v1[E];
#pragma unroll
    for (int s = 0; s < E; ++s) {
      const int n = Plan::in_index(t, s);
      const int a1 = variant ? N - 1 - 2 * n : 2 * n;
      const int a2 = (N - 1) - a1;
      float2 l1 = ld2<C, ROW>(arow, a1), l2 = ld2<C, ROW>(arow, a2);
      const float2 q1 = ld2<C, ROW>(reinterpret_cast<const float*>(qrow), a1);
      const float2 q2 = ld2<C, ROW>(reinterpret_cast<const float*>(qrow), a2);
      l1.x *= static_cast<float>(__float_as_int(q1.x));
      l1.y *= static_cast<float>(__float_as_int(q1.y));
      l2.x *= static_cast<float>(__float_as_int(q2.x));
      l2.y *= static_cast<float>(__float_as_int(q2.y));
      const float4 k4 = __ldg(&tb.pre_inv[variant * M + n]);
      v0[s] = make_float2(fmaf(l2.x, k4.y, l1.x * k4.x), fmaf(l2.x, k4.w, l1.x * k4.z));
      v1[s] = make_float2(fmaf(l2.y, k4.y, l1.y * k4.x), fmaf(l2.y, k4.w, l1.y * k4.z));
    }
    __syncthreads();                           // integers consumed by everybody: their buffer takes the next tile
    if (tid == 0 && more) issue_load(tile + gridDim.x, qbuf, q, ROW, &mbar[1]);

    fft2<Plan>(v0, v1, reinterpret_cast<float4*>(arow), t, g, tb.tw_pass1, tb.tw_pass2);
    post_store<Plan, C, ROW>(v0, v1, arow, t, variant, tb.post_inv);
    __syncthreads();

    // ---- synthesis window + TDAC overlap-add (H_inv, mdctransformer.py:148,176-190), straight to global memory
    const int nblk = min(FP - 1, frames_n + 1 - nb0);
    float* xb = x + (b * (frames_n + 1) + nb0) * static_cast<int64_t>(ROW);
    for (int idx = tid; idx < nblk * H; idx += THREADS) {
      const int bl = idx / H, p = idx % H;
      const float4 sw = __ldg(&tb.unfold[p]);
      const float* vn = abuf + (bl + 1) * ROW + (H - 1 - p) * C;
      const float* vp = abuf + bl * ROW + (H + p) * C;
      float* xo = xb + static_cast<int64_t>(bl) * ROW;
      if constexpr (C == 2) {
        const float2 a = *reinterpret_cast<const float2*>(vn), c = *reinterpret_cast<const float2*>(vp);
        *reinterpret_cast<float2*>(xo + 2 * p) = make_float2(fmaf(sw.x, a.x, sw.y * c.x), fmaf(sw.x, a.y, sw.y * c.y));
        *reinterpret_cast<float2*>(xo + 2 * (N - 1 - p)) = make_float2(fmaf(sw.z, a.x, sw.w * c.x), fmaf(sw.z, a.y, sw.w * c.y));
      } else {
        xo[p] = fmaf(sw.x, vn[0], sw.y * vp[0]);
        xo[N - 1 - p] = fmaf(sw.z, vn[0], sw.w * vp[0]);
      }
    }
    __syncthreads();       // the scratch tile is rewritten by the transform of the next tile
  }
}

template <typename Plan, int C, int THREADS, int MINB>
cudaError_t launch_inverse_compact_tile(const MdctDeviceTables& tb, const int32_t* q, const float* bark,
                                        const float4* filt4, float eps_s2, float* x, int64_t batches, int frames_n,
                                        cudaStream_t stream) {
  using S = TileShape<Plan, C, THREADS>;
  static_assert(S::FP >= 2, "an inverse tile needs two frames");
  constexpr size_t kSmem = (2 * static_cast<size_t>(S::FP) * S::ROW + static_cast<size_t>(S::FP) * 64 * C) * sizeof(float) + 32;
  if (kSmem > 227 * 1024) return cudaErrorInvalidConfiguration;
  const int tiles_per_row = (frames_n + 1 + S::FP - 2) / (S::FP - 1);
  const int64_t total = batches * tiles_per_row;
  constexpr int kFit = static_cast<int>((227 * 1024) / (kSmem + 1024));
  constexpr int kMinB = kFit < 1 ? 1 : (kFit < MINB ? kFit : MINB);
  const int64_t cap = static_cast<int64_t>(tile_sm_count()) * kMinB;
  const unsigned grid = static_cast<unsigned>(std::min(total, cap));
  auto kernel = mdct_inverse_dequant_compact_tile_kernel<Plan, C, THREADS, kMinB>;
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmem));
  if (err != cudaSuccess) return err;
  err = launch_pdl(4, kernel, grid, THREADS, kSmem, stream, tb, q, bark, filt4, eps_s2, x, frames_n, tiles_per_row, total);
  count_launch();
  return cudaGetLastError();
}

// Plain inverse (no fused dequantisation): the amplitudes are the only input, so the second tile buffer that the
// dequantising kernel spends on the integers is free to double-buffer the bulk loads, and the overlap-add writes
// its output blocks straight to global memory (coalesced 8-byte stores) instead of staging them.
template <typename Plan, int C, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
mdct_inverse_plain_tile_kernel(MdctDeviceTables tb, const float* __restrict__ y, float* __restrict__ x, int frames_n,
                               int tiles_per_row, int64_t total_tiles) {
  using S = TileShape<Plan, C, THREADS>;
  constexpr int M = S::M, N = S::N, H = M, T = S::T, E = Plan::E, FP = S::FP, ROW = S::ROW;
  constexpr int BUF = FP * ROW;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* bufs = reinterpret_cast<float*>(smem_raw);                  // [2][FP][ROW]: amplitudes -> scratch -> v
  uint64_t* mbar = reinterpret_cast<uint64_t*>(bufs + 2 * BUF);      // [2]

  const int tid = threadIdx.x, g = tid / T, t = tid % T, variant = (tid >> 3) & 1;
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  pdl_launch_dependents();                  // the next kernel on the stream may start its prologue
  pdl_wait();                               // the kernel before this one has completed: its tensors are visible

  auto issue_load = [&](int64_t tile, int slot) {        // thread 0 only
    const int64_t b = tile / tiles_per_row;
    const int fs = static_cast<int>(tile - b * tiles_per_row) * (FP - 1) - 1;
    const int r_lo = fs < 0 ? 1 : 0;
    const int r_hi = min(FP, frames_n - fs);
    if (r_hi > r_lo) {
      const uint32_t bytes = static_cast<uint32_t>(r_hi - r_lo) * ROW * sizeof(float);
      mbar_arrive_expect_tx(&mbar[slot], bytes);
      bulk_load(bufs + slot * BUF + r_lo * ROW, y + (b * frames_n + (fs + r_lo)) * static_cast<int64_t>(ROW), bytes, &mbar[slot]);
    } else {
      mbar_arrive(&mbar[slot]);
    }
  };
  if (tid == 0 && blockIdx.x < total_tiles) issue_load(blockIdx.x, 0);

  int it = 0;
  for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
    const int slot = it & 1;
    float* abuf = bufs + slot * BUF;
    float* arow = abuf + g * (2 / C) * ROW;
    const int64_t b = tile / tiles_per_row;
    const int nb0 = static_cast<int>(tile - b * tiles_per_row) * (FP - 1);   // first output block
    const int fs = nb0 - 1;
    const int r_lo = fs < 0 ? 1 : 0;
    const int r_hi = min(FP, frames_n - fs);
    // the other buffer was last read by the overlap-add of the previous tile, which ended in a block-wide barrier
    if (tid == 0 && tile + gridDim.x < total_tiles) issue_load(tile + gridDim.x, slot ^ 1);
    mbar_wait(&mbar[slot], (it >> 1) & 1);
    if (r_lo > 0 || r_hi < FP) {               // frames outside the signal are zero (mdctransformer.py:366)
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r_lo > 0)
        for (int i = tid * 4; i < ROW; i += THREADS * 4) *reinterpret_cast<float4*>(abuf + i) = z;
      for (int i = max(r_hi, r_lo) * ROW + tid * 4; i < BUF; i += THREADS * 4) *reinterpret_cast<float4*>(abuf + i) = z;
      __syncthreads();
    }

    float2 v0[E], v1[E];
#pragma unroll
    for (int s = 0; s < E; ++s) {
      const int n = Plan::in_index(t, s);
      const int a1 = variant ? N - 1 - 2 * n : 2 * n;
      const int a2 = (N - 1) - a1;
      const float2 l1 = ld2<C, ROW>(arow, a1), l2 = ld2<C, ROW>(arow, a2);
      const float4 k4 = __ldg(&tb.pre_inv[variant * M + n]);
      v0[s] = make_float2(fmaf(l2.x, k4.y, l1.x * k4.x), fmaf(l2.x, k4.w, l1.x * k4.z));
      v1[s] = make_float2(fmaf(l2.y, k4.y, l1.y * k4.x), fmaf(l2.y, k4.w, l1.y * k4.z));
    }
    group_sync<T>(g);
    fft2<Plan>(v0, v1, reinterpret_cast<float4*>(arow), t, g, tb.tw_pass1, tb.tw_pass2);
    post_store<Plan, C, ROW>(v0, v1, arow, t, variant, tb.post_inv);
    __syncthreads();

    const int nblk = min(FP - 1, frames_n + 1 - nb0);
    float* xb = x + (b * (frames_n + 1) + nb0) * static_cast<int64_t>(ROW);
    for (int idx = tid; idx < nblk * H; idx += THREADS) {
      const int bl = idx / H, p = idx % H;
      const float4 s = __ldg(&tb.unfold[p]);
      const float* vn = abuf + (bl + 1) * ROW + (H - 1 - p) * C;
      const float* vp = abuf + bl * ROW + (H + p) * C;
      float* xo = xb + static_cast<int64_t>(bl) * ROW;
      if constexpr (C == 2) {
        const float2 a = *reinterpret_cast<const float2*>(vn), c = *reinterpret_cast<const float2*>(vp);
        *reinterpret_cast<float2*>(xo + 2 * p) = make_float2(fmaf(s.x, a.x, s.y * c.x), fmaf(s.x, a.y, s.y * c.y));
        *reinterpret_cast<float2*>(xo + 2 * (N - 1 - p)) = make_float2(fmaf(s.z, a.x, s.w * c.x), fmaf(s.z, a.y, s.w * c.y));
      } else {
        xo[p] = fmaf(s.x, vn[0], s.y * vp[0]);
        xo[N - 1 - p] = fmaf(s.z, vn[0], s.w * vp[0]);
      }
    }
    __syncthreads();       // this buffer is the target of the bulk load issued at the top of the next iteration
  }
}

// ------------------------------------------------------------------------------------------ launchers
template <typename Plan, int C, int THREADS, int MINB>
cudaError_t launch_forward_tile(const MdctDeviceTables& tb, const float* x, float* y, int64_t batches, int blocks_n,
                                cudaStream_t stream) {
  using S = TileShape<Plan, C, THREADS>;
  const size_t smem = static_cast<size_t>(2 * (S::FP + 1)) * S::ROW * sizeof(float) + 16;
  auto kernel = mdct_forward_tile_kernel<Plan, C, THREADS, MINB>;
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (err != cudaSuccess) return err;
  const int tiles_per_row = (blocks_n + 1 + S::FP - 1) / S::FP;
  const int64_t total = batches * tiles_per_row;
  int per_sm = static_cast<int>((227 * 1024) / (smem + 1024));
  per_sm = std::max(1, std::min(per_sm, MINB));
  const int64_t cap = static_cast<int64_t>(tile_sm_count()) * per_sm;
  err = launch_pdl(1, kernel, static_cast<unsigned>(std::min(total, cap)), THREADS, smem, stream, tb, x, y, blocks_n, tiles_per_row, total);
  count_launch();
  return cudaGetLastError();
}

template <typename Plan, int C, int THREADS, int MINB>
cudaError_t launch_inverse_tile(const MdctDeviceTables& tb, const float* y, const int32_t* q, const float* thr, float* x,
                                int64_t batches, int frames_n, cudaStream_t stream) {
  using S = TileShape<Plan, C, THREADS>;
  static_assert(S::FP >= 2, "an inverse tile needs two frames");
  const int buffers = q != nullptr ? 3 : 2;          // steps x 2 + integers, or amplitudes x 2
  const size_t smem = static_cast<size_t>(buffers * S::FP) * S::ROW * sizeof(float) + 32;
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  const int tiles_per_row = (frames_n + 1 + S::FP - 2) / (S::FP - 1);
  const int64_t total = batches * tiles_per_row;
  int per_sm = static_cast<int>((227 * 1024) / (smem + 1024));
  per_sm = std::max(1, std::min(per_sm, MINB));
  const int64_t cap = static_cast<int64_t>(tile_sm_count()) * per_sm;
  const unsigned grid = static_cast<unsigned>(std::min(total, cap));
  cudaError_t err;
  if (q != nullptr) {
    // the register budget follows the CTAs that actually fit: three tile buffers often allow fewer than MINB
    constexpr int kFit = static_cast<int>((227 * 1024) / (3 * S::FP * S::ROW * sizeof(float) + 32 + 1024));
    constexpr int kMinB = kFit < 1 ? 1 : (kFit < MINB ? kFit : MINB);
    auto kernel = mdct_inverse_dequant_tile_kernel<Plan, C, THREADS, kMinB>;
    err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (err != cudaSuccess) return err;
    err = launch_pdl(4, kernel, grid, THREADS, smem, stream, tb, q, thr, x, frames_n, tiles_per_row, total);
  } else {
    auto kernel = mdct_inverse_plain_tile_kernel<Plan, C, THREADS, MINB>;
    err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (err != cudaSuccess) return err;
    err = launch_pdl(4, kernel, grid, THREADS, smem, stream, tb, y, x, frames_n, tiles_per_row, total);
  }
  count_launch();
  return cudaGetLastError();
}

//                       M    E  R0  R1 R2

}  // namespace

int tile_sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
      cached = sms;
    else
      return 148;
  }
  return cached;
}

bool mdct_tile_forward_supported(int n, int channels) {
  return (channels == 1 || channels == 2) && (n == 64 || n == 128 || n == 256 || n == 512 || n == 1024);
}

bool mdct_tile_inverse_supported(int n, int channels) {
  return (channels == 1 || channels == 2) && (n == 64 || n == 128 || n == 256 || n == 512 || n == 1024);
}

cudaError_t mdct_forward_tile(const MdctDeviceTables& tb, const float* x, float* y, int64_t batches, int64_t blocks_n,
                              int C, cudaStream_t stream) {
  const int bn = static_cast<int>(blocks_n);
#define AC_FWD(PLAN, THREADS, MINB)                                                                   \
  return C == 2 ? launch_forward_tile<PLAN, 2, THREADS, MINB>(tb, x, y, batches, bn, stream)         \
                : launch_forward_tile<PLAN, 1, THREADS, MINB>(tb, x, y, batches, bn, stream)
  switch (tb.n) {
    case 64: AC_FWD(Plan64, 128, 4);
    case 128: AC_FWD(Plan128, 128, 4);
    case 256: AC_FWD(Plan256, 128, 3);
    case 512: AC_FWD(Plan512, 128, 3);
    case 1024: AC_FWD(Plan1024, 512, 1);
    default: return cudaErrorInvalidConfiguration;
  }
#undef AC_FWD
}

cudaError_t mdct_inverse_tile(const MdctDeviceTables& tb, const float* y, const int32_t* q, const float* thr, float* x,
                              int64_t batches, int64_t frames_n, int C, cudaStream_t stream) {
  const int fn = static_cast<int>(frames_n);
#define AC_INV(PLAN, THREADS, MINB)                                                                          \
  return C == 2 ? launch_inverse_tile<PLAN, 2, THREADS, MINB>(tb, y, q, thr, x, batches, fn, stream)        \
                : launch_inverse_tile<PLAN, 1, THREADS, MINB>(tb, y, q, thr, x, batches, fn, stream)
  switch (tb.n) {
    case 64: AC_INV(Plan64, 128, 4);
    case 128: AC_INV(Plan128, 128, 4);
    case 256: AC_INV(Plan256, 128, 3);
    case 512: AC_INV(Plan512, 128, 3);
    case 1024: AC_INV(Plan1024, 512, 1);
    default: return cudaErrorInvalidConfiguration;
  }
#undef AC_INV
}

cudaError_t mdct_inverse_compact_tile(const MdctDeviceTables& tb, const int32_t* q, const float* bark, const float4* filt4,
                                      float eps_s2, float* x, int64_t batches, int64_t frames_n, int C,
                                      cudaStream_t stream) {
  const int fn = static_cast<int>(frames_n);
#define AC_INVC(PLAN, THREADS, MINB)                                                                                 \
  return C == 2 ? launch_inverse_compact_tile<PLAN, 2, THREADS, MINB>(tb, q, bark, filt4, eps_s2, x, batches, fn, stream) \
                : launch_inverse_compact_tile<PLAN, 1, THREADS, MINB>(tb, q, bark, filt4, eps_s2, x, batches, fn, stream)
  switch (tb.n) {
    case 256: AC_INVC(Plan256, 128, 3);
    case 512: AC_INVC(Plan512, 128, 3);
    case 1024: AC_INVC(Plan1024, 512, 1);
    default: return cudaErrorInvalidConfiguration;
  }
#undef AC_INVC
}

}  // namespace ac
