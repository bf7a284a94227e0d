// Psychoacoustic tile kernel, second generation (sm_100a): global masking threshold + fused quantiser with the
// 64 x 64 spreading contraction on the tensor cores.
//
// Reference behaviour: /root/reference/audiocodec/psychoacoustic.py:102-120 (tonality), :122-148
// (global_masking_threshold), :169-210 (_masking_intensity_in_bark), :301-331 (bark mappings); quantiser = SURVEY.md
// 8a row Q.  Same mathematics as pa_tile_kernel (psycho_kernels.cu), restructured around its instruction budget:
//
//   A1  lane <-> filter k   coalesced read of y, I = y^2 written TRANSPOSED to T[k][item], tonality sums
//   A2  lane <-> item       one job per (bark band, chunk): steps of four filters, then P = max(eps, I_bark)^alpha
//                           through an exponent table (x^a = 2^(a lg2 mantissa + r[E]) * 2^n[E]: no int <-> float
//                           conversions, the exponent keeps fp32 precision), P stored XOR-swizzled
//   B   mma.sync m16n8k8    acc[item][j] = sum_i P[item][i] S[i][j] as an error-compensated TF32 product
//                           (P = hi + lo, S = hi + lo, three MMAs: lo hi + hi lo + hi hi; every term is positive, the
//                           dropped lo lo term is 2^-22 relative) with fp32 accumulators: 32 items x 32 bands per warp.
//                           The Toeplitz S is read from two 128-entry tables; fragment (k-step, n-tile) only depends
//                           on n-tile - k-step, so one new fragment per k-step is loaded and the rest rotate.
//       epilogue            in the accumulator layout: the masking offset joins the exponent of ^(1/alpha)
//                           (10^(-alpha offset / 10))^(1/alpha) = 2^(offset_log2 offset)), quiet threshold, scale^2
//   D   lane <-> filter k   thr = sqrt(sum_b G[b] W_inv[b][k]) as v * rsqrt(v); the same rsqrt seeds the division
//                           q = rint(y / thr) (two exact-residual corrections: the IEEE quotient), coalesced stores
//
// y is read from HBM in A1 and again (an L2 hit) in D: HBM sees one read of y and one write each of thr and q.
#include "kernels.h"

#include <algorithm>
#include <cstdint>

namespace ac {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = 8;
constexpr int kTI = 64;       // items (frame, channel) per tile; lane l of the band sums owns items 2l, 2l + 1
constexpr int kTS = 66;       // row stride of T (words): 8-byte aligned item pairs, half-warps conflict-free
constexpr int kGS = 68;       // row stride of G: conflict-free stores from the accumulator layout (8 t + g), 16-byte rows
constexpr int kPS = 64;       // row stride of P
constexpr int kNB = 64;       // bark bands
constexpr int kPartS = 65;    // row stride of the tonality partials

template <int C> struct Vec;
template <> struct Vec<1> { using F = float; using I = int32_t; };
template <> struct Vec<2> { using F = float2; using I = int2; };
template <> struct Vec<4> { using F = float4; using I = int4; };

typedef unsigned long long u64;

__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// D += A B, m16n8k8, A row-major (items x masker bands), B column-major (masker x maskee bands), TF32 in, fp32 accumulate.
// Fragment layout (PTX ISA, mma.m16n8k8 .tf32), g = lane / 4, t = lane % 4:
//   a0 (g, t)  a1 (g + 8, t)  a2 (g, t + 4)  a3 (g + 8, t + 4);  b0 (k = t, n = g)  b1 (k = t + 4, n = g);
//   d0 (g, 2t)  d1 (g, 2t + 1)  d2 (g + 8, 2t)  d3 (g + 8, 2t + 1)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// v = hi + lo with hi the nearest TF32 value (10 explicit mantissa bits; valid for finite positive v) and lo the
// remainder cut to TF32: |v - hi - lo| <= 2^-21 |v|
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(v - __uint_as_float(hi)) & 0xffffe000u;
}

// x^a for x > 0 through the exponent table of a (kernels.h): tab[E] = { 2^rint(a e), a e - rint(a e) }, e = E - 127
__device__ __forceinline__ float pow_tab(float x, float a, const float2* tab) {
  const uint32_t bits = __float_as_uint(x);
  const float2 en = tab[bits >> 23];
  const float m = __uint_as_float((bits & 0x007fffffu) | 0x3f800000u);   // [1, 2)
  return ex2_approx(fmaf(a, lg2_approx(m), en.y)) * en.x;
}

__device__ __forceinline__ float tonality_from_log2_sums(float sum_i, float sum_log2, int n, float eps) {
  // psychoacoustic.py:113-118 with 10 log10(GM / AM) = 10 log10(2) (mean log2 - log2 AM)
  const float inv_n = 1.0f / static_cast<float>(n);
  const float am = fmaf(sum_i, inv_n, eps);
  const float sfm = 3.010299956639812f * (sum_log2 * inv_n - log2f(am));
  return fminf(sfm * (-1.0f / 60.0f), 1.0f);
}

__host__ __device__ inline bool filt_in_smem(const PaDeviceTables& tb) { return tb.n <= 512; }


// shared-memory accesses of the band-sum loop by 32-bit shared address + immediate: the address arithmetic stays one
// add per job (generic pointers made the compiler rebuild the shared window base inside the loop)
template <int OFF>
__device__ __forceinline__ u64 lds_b64(uint32_t addr) {
  u64 v;
  asm volatile("ld.shared.b64 %0, [%1+%2];" : "=l"(v) : "r"(addr), "n"(OFF));
  return v;
}
template <int OFF>
__device__ __forceinline__ void lds_2b64(uint32_t addr, u64& x, u64& y) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2+%3];" : "=l"(x), "=l"(y) : "r"(addr), "n"(OFF));
}
__device__ __forceinline__ void sts_b64(uint32_t addr, u64 v) {
  asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}

struct Layout2 {
  int powa, powia, t, p, part, uv, sfh, sfl, quiet, lin, bw8, filt4, total;
};

__host__ __device__ inline Layout2 layout2(const PaDeviceTables& tb) {
  Layout2 L;
  const int kc = tb.n < tb.chunk_k ? tb.n : tb.chunk_k;
  const int t_rows = (kc + 3) * kTS;                    // 3 zero rows behind the chunk for the 4-filter steps
  const int g_words = kNB * kGS;                        // G aliases T (dead once the last band sum is done)
  int o = 0;
  L.powa = o;    o += 512;                              // the exponent tables come first: an index with the sign bit
  L.powia = o;   o += 512;                              //   set (NaN input) still reads inside the allocation
  L.t = o;       o += ((t_rows > g_words ? t_rows : g_words) + 3) & ~3;
  L.p = o;       o += kNB * kPS;
  L.part = o;    o += 8 * kPartS;
  L.uv = o;      o += 3 * kTI;
  L.sfh = o;     o += 128;
  L.sfl = o;     o += 128;
  L.quiet = o;   o += kNB;
  L.lin = o;     o += kNB;
  L.bw8 = o;     o += (2 * tb.n_band_w4 + 3) & ~3;       // every weight twice: a packed pair for both items of a lane
  L.filt4 = o;   o += filt_in_smem(tb) ? 4 * tb.n : 0;   // long filter tables stay in global memory (L1 / L2)
  L.total = o;
  return L;
}

// Phase D of the tile kernel for one frame row (all channels) per call: thr (and q) of the row's n filters.
// G carries scale^2, so sqrt gives the scaled step directly; thr = v rsqrt(v) (relative error 2^-22), and the same
// rsqrt is the reciprocal estimate of thr: two exact-residual corrections make q the IEEE quotient.  A warp writes
// whole rows (n C contiguous floats of thr and of q): long DRAM bursts.
template <int C, bool QUANT, bool THR, bool FILT_SMEM, int NFIX>
__device__ __forceinline__ void phase_d_row(const float4* __restrict__ filt4, const int n_runtime, const int lane,
                                            const float* __restrict__ yrow, float* __restrict__ trow,
                                            int32_t* __restrict__ qrow, const float* gr, const float eps_s2) {
  using VF = typename Vec<C>::F;
  using VI = typename Vec<C>::I;
  constexpr int GS = kGS;
  const int n = NFIX > 0 ? NFIX : n_runtime;
  const VF* yv = reinterpret_cast<const VF*>(yrow);
  VF* tv = reinterpret_cast<VF*>(trow);
  VI* qv = reinterpret_cast<VI*>(qrow);
  if constexpr (C == 2) {
    const u64 k_neg = pack2(-1.f, -1.f);
#pragma unroll 4
    for (int k = lane; k < n; k += 32) {
      const float4 f4 = FILT_SMEM ? filt4[k] : __ldg(filt4 + k);
      const float* gp = gr + __float_as_int(f4.w) * GS;
      const u64 g0 = *reinterpret_cast<const u64*>(gp);
      const u64 g1 = *reinterpret_cast<const u64*>(gp + GS);
      const u64 g2 = *reinterpret_cast<const u64*>(gp + 2 * GS);
      float vx, vy;
      unpack2(ffma2(g2, pack2(f4.z, f4.z), ffma2(g1, pack2(f4.y, f4.y), fmul2(g0, pack2(f4.x, f4.x)))), vx, vy);
      vx = fmaxf(eps_s2, vx);
      vy = fmaxf(eps_s2, vy);
      const u64 r2 = pack2(rsqrt_approx(vx), rsqrt_approx(vy));
      const u64 th2 = fmul2(pack2(vx, vy), r2);
      if (QUANT) {
        const u64 nd = fmul2(th2, k_neg), a2 = __ldg(reinterpret_cast<const u64*>(yv) + k);
        u64 qq = fmul2(a2, r2);
        qq = ffma2(ffma2(nd, qq, a2), r2, qq);
        qq = ffma2(ffma2(nd, qq, a2), r2, qq);
        float qx, qy;
        unpack2(qq, qx, qy);
        __stcs(&qv[k], make_int2(__float2int_rn(qx), __float2int_rn(qy)));   // streaming: keep y in L2, not q
      }
      if (THR) __stcs(reinterpret_cast<u64*>(&tv[k]), th2);
    }
  } else {
#pragma unroll 4
    for (int k = lane; k < n; k += 32) {
      const float4 f4 = FILT_SMEM ? filt4[k] : __ldg(filt4 + k);
      const float* gp = gr + __float_as_int(f4.w) * GS;
      const VF g0 = *reinterpret_cast<const VF*>(gp);
      const VF g1 = *reinterpret_cast<const VF*>(gp + GS);
      const VF g2 = *reinterpret_cast<const VF*>(gp + 2 * GS);
      const float* a0 = reinterpret_cast<const float*>(&g0);
      const float* a1 = reinterpret_cast<const float*>(&g1);
      const float* a2 = reinterpret_cast<const float*>(&g2);
      VF thr_v;
      float* th = reinterpret_cast<float*>(&thr_v);
      float rs[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float v = fmaxf(eps_s2, fmaf(a2[c], f4.z, fmaf(a1[c], f4.y, a0[c] * f4.x)));
        rs[c] = rsqrt_approx(v);
        th[c] = v * rs[c];
      }
      if (QUANT) {
        const VF yl = __ldg(yv + k);
        const float* ya = reinterpret_cast<const float*>(&yl);
        VI q_v;
        int32_t* qa = reinterpret_cast<int32_t*>(&q_v);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          float qq = ya[c] * rs[c];
          qq = fmaf(fmaf(-th[c], qq, ya[c]), rs[c], qq);
          qq = fmaf(fmaf(-th[c], qq, ya[c]), rs[c], qq);
          qa[c] = __float2int_rn(qq);
        }
        __stcs(&qv[k], q_v);
      }
      if (THR) __stcs(&tv[k], thr_v);
    }
  }
}

template <int C, bool QUANT, int NFIX>
__global__ void __launch_bounds__(kThreads, 3)
pa_mma_tile_kernel(const __grid_constant__ PaDeviceTables tb, const __grid_constant__ PaJobParams jp,
                   const float* __restrict__ y, const float* __restrict__ ton_in, float one_minus_drown, float thr_scale,
                   float* __restrict__ thr_out, int32_t* __restrict__ q_out, int64_t frames_total, int64_t tiles) {
  using VF = typename Vec<C>::F;
  using VI = typename Vec<C>::I;
  constexpr int TI = kTI, TS = kTS, GS = kGS, PS = kPS;
  constexpr int FT = TI / C;                    // frames per tile
  constexpr int ROWS = FT / kWarps;             // frame rows per warp
  static_assert(FT % kWarps == 0, "tile shape");
  extern __shared__ __align__(16) float sm[];
  const Layout2 L = layout2(tb);
  const int n = NFIX > 0 ? NFIX : tb.n, kc = n < tb.chunk_k ? n : tb.chunk_k;
  float2* s_powa = reinterpret_cast<float2*>(sm + L.powa);
  float2* s_powia = reinterpret_cast<float2*>(sm + L.powia);
  float* T = sm + L.t;                          // [kc + 3][TS]: I[k][item]
  float* G = sm + L.t;                          // [64][GS]
  float* P = sm + L.p;                          // [64][PS], column item ^ ((band & 3) << 3)
  float* s_part = sm + L.part;                  // [2][4][kPartS]
  float* s_u = sm + L.uv;                       // per item: offset_log2 (1 - drown) tonality
  float* s_v = s_u + TI;                        //           offset_log2 (1 - drown) (9 tonality + 5.5) + log2 scale^2
  float* s_ton = s_v + TI;
  uint32_t* s_sfh = reinterpret_cast<uint32_t*>(sm + L.sfh);   // TF32 halves of spread_fn[0 .. 127]
  uint32_t* s_sfl = reinterpret_cast<uint32_t*>(sm + L.sfl);
  float* s_quiet = sm + L.quiet;
  float* s_lin = sm + L.lin;
  float* s_bw8 = sm + L.bw8;
  const bool filt_smem = filt_in_smem(tb);
  float4* s_filt4 = reinterpret_cast<float4*>(sm + L.filt4);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);        // the same value, known to be warp-uniform
  const float scale = QUANT ? thr_scale : 1.0f;
  const float scale2 = scale * scale;
  for (int i = tid; i < 256; i += kThreads) {
    s_powa[i] = tb.pow_alpha[i];
    s_powia[i] = tb.pow_inv_alpha[i];
  }
  for (int i = tid; i < 128; i += kThreads) {
    uint32_t hi, lo;
    split_tf32(tb.spread_fn[i], hi, lo);
    s_sfh[i] = hi;
    s_sfl[i] = lo;
  }
  for (int i = tid; i < kNB; i += kThreads) {
    s_quiet[i] = tb.quiet[i] * scale2;
    s_lin[i] = tb.lin[i];
  }
  for (int i = tid; i < tb.n_band_w4; i += kThreads) {
    const float w = tb.band_w4[i];
    s_bw8[2 * i] = w;
    s_bw8[2 * i + 1] = w;
  }
  if (filt_smem)
    for (int i = tid; i < n; i += kThreads) s_filt4[i] = tb.filt4[i];
  __syncthreads();

  const uint32_t sm_base = static_cast<uint32_t>(__cvta_generic_to_shared(sm));
  const uint32_t t_lane = sm_base + static_cast<uint32_t>(L.t) * 4u + static_cast<uint32_t>(lane) * 8u;
  const uint32_t w_base = sm_base + static_cast<uint32_t>(L.bw8) * 4u;
  const uint32_t p_base = sm_base + static_cast<uint32_t>(L.p) * 4u;
  const float eps = tb.eps;
  const float eps_s2 = eps * scale2;
  const float log2_s2 = 2.0f * log2f(scale);
  // tiles are walked from the END of the tensor: the producer of y (the forward MDCT) wrote its last ~100 MB into
  // L2 most recently, and the consumer of thr / q (the inverse MDCT) starts at the front, where this kernel ends
  for (int64_t tile_i = blockIdx.x; tile_i < tiles; tile_i += gridDim.x) {
    const int64_t tile = tiles - 1 - tile_i;
    const int64_t f0 = tile * FT;
    const int nf = static_cast<int>(frames_total - f0 < FT ? frames_total - f0 : FT);

    float t_sum[ROWS][C], t_log[ROWS][C];       // tonality sums of this warp's frame rows (psychoacoustic.py:113-116)
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int c = 0; c < C; ++c) t_sum[r][c] = t_log[r][c] = 0.f;

    for (int chunk = 0; chunk < tb.n_chunks; ++chunk) {
      const int kc0 = chunk * tb.chunk_k;
      const int kcn = (n - kc0 < kc ? n - kc0 : kc);          // filters in this chunk
      // ---- A1: I = y^2, transposed; tonality sums                         (psychoacoustic.py:113, :312)
      // frame rows behind the end of the tensor re-read the last frame: their results are never stored
      if (kcn < kc || chunk == 0)                              // zero rows behind a short (or the first) chunk
        for (int i = tid; i < 3 * TS; i += kThreads) T[kcn * TS + i] = 0.f;
      if ((kcn & 127) == 0) {
        for (int kb = 0; kb < kcn; kb += 128) {         // whole 128-filter pieces: all loads first
          VF v[ROWS][4];
#pragma unroll
          for (int r = 0; r < ROWS; ++r) {
            const int fl = min(warp * ROWS + r, nf - 1);
            const VF* row = reinterpret_cast<const VF*>(y) + ((f0 + fl) * static_cast<int64_t>(n) + kc0 + kb + lane);
#pragma unroll
            for (int u = 0; u < 4; ++u) v[r][u] = __ldg(row + u * 32);
          }
#pragma unroll
          for (int r = 0; r < ROWS; ++r) {
            float* tp = T + (kb + lane) * TS + (warp * ROWS + r) * C;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float* a = reinterpret_cast<const float*>(&v[r][u]);
              if constexpr (C == 2) {      // both channels in one packed multiply / add (same IEEE results per half)
                const u64 a2 = pack2(a[0], a[1]);
                const u64 in2 = fmul2(a2, a2);
                float ix, iy;
                unpack2(in2, ix, iy);
                *reinterpret_cast<u64*>(tp + u * 32 * TS) = in2;
                u64 s2 = fadd2(pack2(t_sum[r][0], t_sum[r][1]), in2);
                unpack2(s2, t_sum[r][0], t_sum[r][1]);
                u64 l2 = fadd2(pack2(t_log[r][0], t_log[r][1]), pack2(lg2_approx(fmaxf(eps, ix)), lg2_approx(fmaxf(eps, iy))));
                unpack2(l2, t_log[r][0], t_log[r][1]);
              } else {
                float in[C];
#pragma unroll
                for (int c = 0; c < C; ++c) {
                  in[c] = a[c] * a[c];
                  t_sum[r][c] += in[c];
                  t_log[r][c] += lg2_approx(fmaxf(eps, in[c]));
                }
                if constexpr (C == 4) {
                  *reinterpret_cast<float2*>(tp + u * 32 * TS) = make_float2(in[0], in[1]);
                  *reinterpret_cast<float2*>(tp + u * 32 * TS + 2) = make_float2(in[2], in[3]);
                } else {
                  tp[u * 32 * TS] = in[0];
                }
              }
            }
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          const int fl = warp * ROWS + r;
          const VF* row = reinterpret_cast<const VF*>(y) + ((f0 + min(fl, nf - 1)) * static_cast<int64_t>(n) + kc0);
          float* tcol = T + fl * C;
          for (int kb = 0; kb < kcn; kb += 32) {
            const int k = kb + lane;
            if (k < kcn) {
              const VF v = __ldg(row + k);
              const float* a = reinterpret_cast<const float*>(&v);
#pragma unroll
              for (int c = 0; c < C; ++c) {
                const float in = a[c] * a[c];
                tcol[k * TS + c] = in;
                t_sum[r][c] += in;
                t_log[r][c] += lg2_approx(fmaxf(eps, in));
              }
            }
          }
        }
      }
      const bool last_chunk = chunk == tb.n_chunks - 1;
      if (last_chunk && ton_in == nullptr) {
        // fold the 32 lane partials of every row to 4 and park them for the lane <-> item pass
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
#pragma unroll
          for (int c = 0; c < C; ++c) {
            float a = t_sum[r][c], b = t_log[r][c];
#pragma unroll
            for (int o = 16; o >= 4; o >>= 1) {
              a += __shfl_xor_sync(0xffffffffu, a, o);
              b += __shfl_xor_sync(0xffffffffu, b, o);
            }
            if (lane < 4) {
              const int item = (warp * ROWS + r) * C + c;
              s_part[lane * kPartS + item] = a;
              s_part[(4 + lane) * kPartS + item] = b;
            }
          }
      }
      __syncthreads();

      // ---- per-item constants of the masking offset (psychoacoustic.py:185-191), once per tile
      if (last_chunk && warp < TI / 32) {
        const int it = warp * 32 + lane;
        float ton;
        if (ton_in == nullptr) {                         // tonality of the item (psychoacoustic.py:113-118)
          const float* sp = s_part + it;
          const float s_i = (sp[0] + sp[kPartS]) + (sp[2 * kPartS] + sp[3 * kPartS]);
          const float s_l = (sp[4 * kPartS] + sp[5 * kPartS]) + (sp[6 * kPartS] + sp[7 * kPartS]);
          ton = tonality_from_log2_sums(s_i, s_l, n, eps);
        } else {
          const int64_t item = f0 * C + it;
          ton = item < frames_total * C ? __ldg(ton_in + item) : 0.f;
        }
        const float ko = tb.offset_log2 * one_minus_drown;
        s_u[it] = ko * ton;
        s_v[it] = fmaf(ko, fmaf(9.f, ton, 5.5f), log2_s2);
        s_ton[it] = ton;
      }

      // ---- A2: band energies of this chunk; P = max(eps, I_bark)^alpha when a band is complete  (:204-206, :313)
      // lane l owns the item pair (2l, 2l + 1) as packed fp32: one LDS.64 and one FFMA2 per filter for two items
      {
        const int j0 = jp.start[chunk * 9 + warp_u], j1 = jp.start[chunk * 9 + warp_u + 1];
#pragma unroll 1
        for (int j = j0; j < j1; ++j) {
          const int4 jb = jp.job[j];           // { T byte offset, weight byte offset, steps, P byte offset | flags << 16 }
          uint32_t tp = t_lane + static_cast<uint32_t>(jb.x);
          uint32_t wp = w_base + static_cast<uint32_t>(jb.y);
          u64 a0 = 0ull, a1 = 0ull;           // two chains: even and odd steps
          int s = jb.z;
#pragma unroll 1
          for (; s >= 2; s -= 2) {
            u64 w0, w1, w2, w3, w4, w5, w6, w7;
            lds_2b64<0>(wp, w0, w1);
            lds_2b64<16>(wp, w2, w3);
            lds_2b64<32>(wp, w4, w5);
            lds_2b64<48>(wp, w6, w7);
            a0 = ffma2(lds_b64<0>(tp), w0, a0);
            a1 = ffma2(lds_b64<4 * TS * 4>(tp), w4, a1);
            a0 = ffma2(lds_b64<1 * TS * 4>(tp), w1, a0);
            a1 = ffma2(lds_b64<5 * TS * 4>(tp), w5, a1);
            a0 = ffma2(lds_b64<2 * TS * 4>(tp), w2, a0);
            a1 = ffma2(lds_b64<6 * TS * 4>(tp), w6, a1);
            a0 = ffma2(lds_b64<3 * TS * 4>(tp), w3, a0);
            a1 = ffma2(lds_b64<7 * TS * 4>(tp), w7, a1);
            tp += 8 * TS * 4;
            wp += 64;
          }
          if (s) {
            u64 w0, w1, w2, w3;
            lds_2b64<0>(wp, w0, w1);
            lds_2b64<16>(wp, w2, w3);
            a0 = ffma2(lds_b64<0>(tp), w0, a0);
            a1 = ffma2(lds_b64<1 * TS * 4>(tp), w1, a1);
            a0 = ffma2(lds_b64<2 * TS * 4>(tp), w2, a0);
            a1 = ffma2(lds_b64<3 * TS * 4>(tp), w3, a1);
          }
          u64 acc2 = fadd2(a0, a1);
          const uint32_t pp = p_base + ((static_cast<uint32_t>(jb.w) & 0xffffu) ^ (static_cast<uint32_t>(lane) << 3));
          if (jb.w & 0x10000) acc2 = fadd2(acc2, lds_b64<0>(pp));
          if (jb.w & 0x20000) {
            float ax, ay;
            unpack2(acc2, ax, ay);
            acc2 = pack2(pow_tab(fmaxf(eps, ax), tb.alpha, s_powa), pow_tab(fmaxf(eps, ay), tb.alpha, s_powa));
          }
          sts_b64(pp, acc2);
        }
      }
      __syncthreads();
    }

    // ---- B: spreading on the tensor cores                                 (psychoacoustic.py:195-206)
    {
      // pull the next tile of y towards L2 while this phase only computes
      const int64_t next0 = (tile - gridDim.x) * FT;
      if (next0 >= 0) {
        const int64_t next_floats = (frames_total - next0 < FT ? frames_total - next0 : FT) * static_cast<int64_t>(n) * C;
        const float* np = y + next0 * static_cast<int64_t>(n) * C;
        for (int64_t o = static_cast<int64_t>(tid) * 32; o < next_floats; o += kThreads * 32)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(np + o));
      }
      const int g = lane >> 2, t = lane & 3;
      const int m0 = (warp & 3) * 16, nq = warp >> 2;            // 16 items x 32 maskee bands per warp
      const int col0 = (m0 + g) ^ (t << 3), col1 = (m0 + g + 8) ^ (t << 3);
      const float* pa = P + t * PS;                               // masker band 8 ks + t (and + 4)
      // S[i][j] = spread_fn[64 - i + j]: b0 of (k-step ks, n-tile nt) sits at lb + 8 (nt - ks), b1 four entries below
      const int lb = 64 + g - t + 32 * nq;
      float acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
      uint32_t bh0[4], bh1[4], bl0[4], bl1[4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        bh0[nt] = s_sfh[lb + 8 * nt];
        bh1[nt] = s_sfh[lb + 8 * nt - 4];
        bl0[nt] = s_sfl[lb + 8 * nt];
        bl1[nt] = s_sfl[lb + 8 * nt - 4];
      }
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        if (ks > 0) {
#pragma unroll
          for (int nt = 3; nt > 0; --nt) {
            bh0[nt] = bh0[nt - 1];
            bh1[nt] = bh1[nt - 1];
            bl0[nt] = bl0[nt - 1];
            bl1[nt] = bl1[nt - 1];
          }
          bh0[0] = s_sfh[lb - 8 * ks];
          bh1[0] = s_sfh[lb - 8 * ks - 4];
          bl0[0] = s_sfl[lb - 8 * ks];
          bl1[0] = s_sfl[lb - 8 * ks - 4];
        }
        const float* pk = pa + ks * 8 * PS;
        uint32_t ah[4], al[4];
        split_tf32(pk[col0], ah[0], al[0]);
        split_tf32(pk[col1], ah[1], al[1]);
        split_tf32(pk[4 * PS + col0], ah[2], al[2]);
        split_tf32(pk[4 * PS + col1], ah[3], al[3]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          mma_tf32(acc[nt], al, bh0[nt], bh1[nt]);
          mma_tf32(acc[nt], ah, bl0[nt], bl1[nt]);
          mma_tf32(acc[nt], ah, bh0[nt], bh1[nt]);
        }
      }

      // masking offset, non-linear superposition, quiet threshold            (psychoacoustic.py:185-208, :144)
      const int ma = m0 + g, mb = m0 + g + 8;
      if (!tb.clamp_needed) {
        const float ua = s_u[ma], va = s_v[ma], ub = s_u[mb], vb = s_v[mb];
        const float inva = tb.inv_alpha;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int j = 32 * nq + 8 * nt + 2 * t;
          const float2 lin2 = *reinterpret_cast<const float2*>(s_lin + j);
          const float2 q2 = *reinterpret_cast<const float2*>(s_quiet + j);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float lin = (e & 1) ? lin2.y : lin2.x, qt = (e & 1) ? q2.y : q2.x;
            const float u = (e & 2) ? ub : ua, v = (e & 2) ? vb : va;
            const uint32_t bits = __float_as_uint(acc[nt][e]);
            const float2 en = s_powia[bits >> 23];
            const float mant = __uint_as_float((bits & 0x007fffffu) | 0x3f800000u);
            // (acc 2^(gain_log2 offset))^(1/alpha) scale^2 = 2^(log2(acc) / alpha + offset_log2 offset + log2 scale^2)
            const float f = fmaf(u, lin, fmaf(inva, lg2_approx(mant), en.y)) + v;
            G[(j + (e & 1)) * GS + ((e & 2) ? mb : ma)] = fmaxf(ex2_approx(f) * en.x, qt);
          }
        }
      } else {
        const float ta = s_ton[ma], t9a = 9.f * ta, tn = s_ton[mb], t9b = 9.f * tn;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int j = 32 * nq + 8 * nt + 2 * t;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int jj = j + (e & 1);
            const float ton = (e & 2) ? tn : ta, t9 = (e & 2) ? t9b : t9a;
            const float offset = one_minus_drown * ((ton * s_lin[jj] + t9) + 5.5f);
            const float gain = ex2_approx(tb.gain_log2 * offset);
            const float msk = pow_tab(fmaxf(eps, acc[nt][e] * gain), tb.inv_alpha, s_powia);
            G[jj * GS + ((e & 2) ? mb : ma)] = fmaxf(msk * scale2, s_quiet[jj]);
          }
        }
      }
    }
    __syncthreads();

    // ---- D: back to the filter bands, amplitude, optional quantiser      (:330-331; quantiser: SURVEY 8a row Q)
    {
      const bool thr = thr_out != nullptr;
#pragma unroll 1
      for (int r = 0; r < ROWS; ++r) {
        const int fl = warp * ROWS + r;
        if (fl >= nf) break;
        const int64_t off = (f0 + fl) * static_cast<int64_t>(n) * C;
        const float* gr = G + fl * C;
#define AC_PHASE_D(THR_, FS_) \
  phase_d_row<C, QUANT, THR_, FS_, NFIX>(FS_ ? s_filt4 : tb.filt4, n, lane, y + off, thr_out + off, q_out + off, gr, eps_s2)
        if (filt_smem) {
          if (thr) AC_PHASE_D(true, true); else AC_PHASE_D(false, true);
        } else {
          if (thr) AC_PHASE_D(true, false); else AC_PHASE_D(false, false);
        }
#undef AC_PHASE_D
      }
    }
    __syncthreads();       // G (aliasing T) and P are rewritten by the next tile
  }
}

int mma_sm_count() {
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

template <int C, bool QUANT, int NFIX>
cudaError_t launch_mma_tile_n(const PaDeviceTables& tb, const float* y, const float* ton_in, float omd, float thr_scale,
                              float* thr_out, int32_t* q_out, int64_t frames, cudaStream_t stream) {
  constexpr int FT = kTI / C;
  const size_t smem = static_cast<size_t>(layout2(tb).total) * sizeof(float);
  const int64_t tiles = (frames + FT - 1) / FT;
  int per_sm = static_cast<int>((227 * 1024) / (smem + 1024));
  per_sm = per_sm > 3 ? 3 : (per_sm < 1 ? 1 : per_sm);
  const int64_t cap = static_cast<int64_t>(mma_sm_count()) * per_sm;
  const unsigned grid = static_cast<unsigned>(tiles < cap ? tiles : cap);
  auto kernel = pa_mma_tile_kernel<C, QUANT, NFIX>;
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (err != cudaSuccess) return err;
  kernel<<<grid, kThreads, smem, stream>>>(tb, *tb.jobs_host, y, ton_in, omd, thr_scale, thr_out, q_out, frames, tiles);
  count_launch();
  return cudaGetLastError();
}

// filters_n of the headline configurations are compile-time values (row strides become immediates)
template <int C, bool QUANT>
cudaError_t launch_mma_tile_q(const PaDeviceTables& tb, const float* y, const float* ton_in, float omd, float thr_scale,
                              float* thr_out, int32_t* q_out, int64_t frames, cudaStream_t stream) {
  if (C <= 2 && tb.n == 256)
    return launch_mma_tile_n<C, QUANT, (C <= 2 ? 256 : 0)>(tb, y, ton_in, omd, thr_scale, thr_out, q_out, frames, stream);
  if (C == 2 && tb.n == 1024)
    return launch_mma_tile_n<C, QUANT, (C == 2 ? 1024 : 0)>(tb, y, ton_in, omd, thr_scale, thr_out, q_out, frames, stream);
  return launch_mma_tile_n<C, QUANT, 0>(tb, y, ton_in, omd, thr_scale, thr_out, q_out, frames, stream);
}

template <int C>
cudaError_t launch_mma_tile(const PaDeviceTables& tb, const float* y, const float* ton_in, float omd, float thr_scale,
                            float* thr_out, int32_t* q_out, int64_t frames, cudaStream_t stream) {
  if (q_out != nullptr) return launch_mma_tile_q<C, true>(tb, y, ton_in, omd, thr_scale, thr_out, q_out, frames, stream);
  return launch_mma_tile_q<C, false>(tb, y, ton_in, omd, thr_scale, thr_out, q_out, frames, stream);
}

}  // namespace

bool pa_mma_tile_supported(const PaDeviceTables& tb, int channels) {
  if (!tb.tile_ok || tb.nb != kNB || tb.jobs_host == nullptr) return false;
  if (!(channels == 1 || channels == 2 || channels == 4)) return false;
  return static_cast<size_t>(layout2(tb).total) * sizeof(float) <= 200 * 1024;
}

cudaError_t pa_threshold_mma_tile(const PaDeviceTables& tb, const float* y, const float* ton_in, float one_minus_drown,
                                  float thr_scale, float* thr_out, int32_t* q_out, int64_t frames, int channels,
                                  cudaStream_t stream) {
  switch (channels) {
    case 1: return launch_mma_tile<1>(tb, y, ton_in, one_minus_drown, thr_scale, thr_out, q_out, frames, stream);
    case 2: return launch_mma_tile<2>(tb, y, ton_in, one_minus_drown, thr_scale, thr_out, q_out, frames, stream);
    case 4: return launch_mma_tile<4>(tb, y, ton_in, one_minus_drown, thr_scale, thr_out, q_out, frames, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace ac
