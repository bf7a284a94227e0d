// Psychoacoustic tile kernel, second generation (sm_100a): global masking threshold + fused quantiser with the
// 64 x 64 spreading contraction on the 5th-generation tensor cores (tcgen05.mma, accumulator in tensor memory).
//
// Reference behaviour: /root/reference/audiocodec/psychoacoustic.py:102-120 (tonality), :122-148
// (global_masking_threshold), :169-210 (_masking_intensity_in_bark), :301-331 (bark mappings); quantiser = SURVEY.md
// 8a row Q.  Same mathematics as pa_tile_kernel (psycho_kernels.cu), restructured around its instruction budget.
// An 8-warp CTA (three per SM) walks tiles of 64 (frame, channel) items; the filter axis goes through shared memory in
// chunks of 64 filters, two chunk buffers (one lands while the other is read):
//
//   load  cp.async          y[frame][k][c] -> T[k][item] by 8-byte (two channels) asynchronous copies: the copy itself
//                           transposes; chunk c + 1 (or chunk 0 of the next tile) is in flight while chunk c is used
//   A1  lane <-> item pair  tonality sums over the chunk (packed fp32, two filters per logarithm); the chunk's filters are
//                           dealt to the warps so that they level the band-sum jobs (PaJobParams::ton_start)
//   A2  lane <-> item pair  one job per (bark band, chunk): steps of four filters (LDS.64, square, FFMA2 per filter and
//                           item pair); the job list is a by-value kernel parameter = constant bank = warp-uniform
//                           control; P = max(eps, I_bark)^alpha as sqrt(x) x^(alpha - 1/2) (alpha near 1/2; else through
//                           an exponent table: x^a = 2^(a lg2 mantissa + r[E]) * 2^n[E]), stored straight into the
//                           operand layout of the product (the job descriptor carries the band's row)
//   B   tcgen05.mma         acc[item][j] = sum_i P[item][i] S[i][j] as an error-compensated TF32 product (P = hi + lo,
//       kind::tf32          S = hi + lo: hi lo + hi hi + lo hi; every term is positive, the dropped lo lo term is 2^-22
//       M 64, N 64, K 8     relative) with the fp32 accumulator in tensor memory: 24 single-thread instructions per tile.
//                           A = P in place (MN-major, SWIZZLE_128B_BASE32B; the unsplit value is the high-order term, the
//                           low-order term goes to the dead chunk buffer), B = the Toeplitz S as a table of 30 K-major core
//                           matrices addressed as overlapping windows; tcgen05.ld.16x256b returns the accumulator in the
//                           mma.sync fragment layout (see tc_desc below).  mma.sync m16n8k8 form of the same product
//                           (fragments from two 128-entry tables, rotated in registers) for plans whose operand layout
//                           would cost a resident CTA and under AC_PA_MMA=sync.
//       epilogue            in the accumulator layout: the masking offset joins the exponent of ^(1/alpha)
//                           (10^(-alpha offset / 10))^(1/alpha) = 2^(offset_log2 offset)), quiet threshold, scale^2
//   D   lane <-> filter k   thr = sqrt(sum_b G[b] W_inv[b][k]) as v * rsqrt(v); the same rsqrt seeds the division
//                           q = rint(y / thr) (two exact-residual corrections: the IEEE quotient); the four item pairs
//                           of a warp share one unrolled body per 64 filters (table entry, slot pattern, band rows)
//
// y is read from HBM by the copies and again (an L2 hit) in D: HBM sees one read of y and one write each of thr and q.
// tools/emulate_pa_mma.py checks the fragment / swizzle index maps on the CPU; profiles/README.md has the measurements.
#include "kernels.h"
#include "async_copy.cuh"
#include "mdct_tile_core.cuh"

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdlib>

namespace ac {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = 8;
constexpr int kTI = 64;       // items (frame, channel) per tile; lane l of the band sums owns items 2l, 2l + 1
constexpr int kTS = 66;       // row stride of T (words): 8-byte aligned item pairs, half-warps conflict-free
constexpr int kGS = 68;       // row stride of G: conflict-free stores from the accumulator layout (8 t + g), 16-byte rows
constexpr int kPS = 64;       // row stride of P
constexpr int kNB = 64;       // bark bands
constexpr int kChunk = 64;    // filters per chunk of T (the job lists are built for it: capi.cu, build_mma_jobs)

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
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
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
  int powa, powia, t, tbuf, p, part, uv, sfh, sfl, quiet, lin, bw8, filt4, misc, total;
};

// the exponent tables are only read by the table powers: plans on the split powers without the clamp do not stage them
__host__ __device__ inline bool pow_tables_needed(const PaDeviceTables& tb) { return !tb.pow_split || tb.clamp_needed; }

// TC (the spreading product on tcgen05.mma, see the kernel): both T chunk buffers and P start on 1024-byte boundaries
// (P and - in the chunk buffer that is dead during the product - its low-order term are SWIZZLE_128B_BASE32B operand
// tiles), the spreading function is staged as two tables of 30 K-major core matrices instead of two 128-entry tables
constexpr int kTcCores = 30, kTcTable = kTcCores * 32;
constexpr int kTcIssuer = 64;         // the thread that issues the product: lane 0 of a warp without per-item constants to compute

// fused: the single-pass encoder (x -> q): T holds ALL filters of the tile's frames (the forward MDCT leaves them
// there, transposed) instead of two chunk buffers
// n_fix: filters_n as a compile-time constant of the calling kernel (0: tb.n)
__host__ __device__ inline Layout2 layout2(const PaDeviceTables& tb, const bool fused = false, const bool tc = false,
                                           const int n_fix = 0) {
  Layout2 L;
  const int n = n_fix > 0 ? n_fix : tb.n;
  const int kc = n < kChunk ? n : kChunk;
  const int t_rows = ((fused ? n : kc) + 3) * kTS;      // 3 zero rows behind the chunk for the 4-filter steps
  int o = 0;
  L.tbuf = tc ? (t_rows + 255) & ~255 : (t_rows + 3) & ~3;
  L.t = o;       o += (fused ? 1 : 2) * L.tbuf;         // two chunk buffers: one is filled while the other is read
  // G [64][kGS] aliases P and the tonality partials behind it: both are dead once the MMA loop and the per-item
  // constants are done, and are next written behind the first barrier of the next tile
  static_assert(kNB * kPS + 2 * kWarps * kTI >= kNB * kGS, "G must fit into P + the tonality partials");
  L.p = o;       o += kNB * kPS;
  L.part = o;    o += 2 * kWarps * kTI;
  L.uv = o;      o += 3 * kTI;
  L.sfh = o;     o += tc ? kTcTable : 128;
  L.sfl = o;     o += tc ? kTcTable : 128;
  L.quiet = o;   o += kNB;
  L.lin = o;     o += kNB;
  // mbarrier of the tcgen05 product (8 bytes), tensor-memory address, mbarrier of the fused encoder's bulk copy, the two
  // ticket slots of the tile scheduler: no static shared memory, so the dynamic window starts 1024-byte aligned
  L.misc = o;    o += 12;
  L.bw8 = o;     o += (2 * tb.n_mma_w4 + 3) & ~3;        // every weight twice: a packed pair for both items of a lane
  L.filt4 = o;   o += filt_in_smem(tb) ? 4 * n : 0;      // long filter tables stay in global memory (L1 / L2)
  // the exponent tables come last (every offset above is a compile-time constant when filters_n is): 256 entries each,
  // and 256 more words so that an index with the sign bit set (NaN input) still reads inside the allocation
  const int pow_words = pow_tables_needed(tb) ? 512 : 0;
  L.powa = o;    o += pow_words;
  L.powia = o;   o += pow_words + (pow_words ? 512 : 0);
  L.total = o;
  return L;
}

// the tensor-memory port needs a dead chunk buffer that holds the 16 KB low-order term of P
__host__ __device__ inline bool tc_layout_ok(const PaDeviceTables& tb) {
  return tb.mma_chunk_k == kChunk && layout2(tb, false, true).tbuf >= kNB * kTI && tb.mma_n_chunks >= 2;
}

// asynchronous global -> shared copies of 4 or 8 bytes (LDGSTS): the destination address is free, so the copy
// itself transposes y[frame][filter][channel] into T[filter][item]
template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(src), "n"(BYTES) : "memory");
}
// with an L2 eviction-priority hint (createpolicy): y is read again from L2 in phase D
template <int BYTES>
__device__ __forceinline__ void cp_async_hint(uint32_t dst, const void* src, uint64_t policy) {
  asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], %2, %3;" ::"r"(dst), "l"(src), "n"(BYTES), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

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

// The same for KI x 32 consecutive filters of a row (filt4 points at this lane's first entry) with the y values yv[i]
// (filter k0 + 32 i + lane) loaded ahead of time by the caller.  masks: three bits
// per 32-filter group, bit s set when any filter of the group has a non-zero weight for band slot s - the shared-memory
// loads of unused slots are skipped (warp-uniform predicates; a skipped term is an exact + 0).
template <int C, bool QUANT, bool THR, bool FILT_SMEM, int KI>
__device__ __forceinline__ void phase_d_unit(const float4* __restrict__ filt4, const unsigned masks,
                                             const typename Vec<C>::F (&yv)[KI], float* __restrict__ trow,
                                             int32_t* __restrict__ qrow, const float* gr, const float eps_s2) {
  using VF = typename Vec<C>::F;
  using VI = typename Vec<C>::I;
  constexpr int GS = kGS;
  VF* tv = reinterpret_cast<VF*>(trow);
  VI* qv = reinterpret_cast<VI*>(qrow);
  const u64 k_neg = pack2(-1.f, -1.f);
#pragma unroll
  for (int i = 0; i < KI; ++i) {
    const float4 f4i = FILT_SMEM ? filt4[32 * i] : __ldg(filt4 + 32 * i);
    const float* gp = gr + __float_as_int(f4i.w) * GS;
    const unsigned m = masks >> (3 * i);
    if constexpr (C == 2) {
      u64 v2;                                    // warp-uniform branch on the slot pattern of the 32-filter group
      if ((m & 7u) == 3u) {
        v2 = ffma2(*reinterpret_cast<const u64*>(gp + GS), pack2(f4i.y, f4i.y),
                   fmul2(*reinterpret_cast<const u64*>(gp), pack2(f4i.x, f4i.x)));
      } else if ((m & 7u) == 6u) {
        v2 = ffma2(*reinterpret_cast<const u64*>(gp + 2 * GS), pack2(f4i.z, f4i.z),
                   fmul2(*reinterpret_cast<const u64*>(gp + GS), pack2(f4i.y, f4i.y)));
      } else {
        v2 = ffma2(*reinterpret_cast<const u64*>(gp + 2 * GS), pack2(f4i.z, f4i.z),
                   ffma2(*reinterpret_cast<const u64*>(gp + GS), pack2(f4i.y, f4i.y),
                         fmul2(*reinterpret_cast<const u64*>(gp), pack2(f4i.x, f4i.x))));
      }
      float vx, vy;
      unpack2(v2, vx, vy);
      vx = fmaxf(eps_s2, vx);
      vy = fmaxf(eps_s2, vy);
      const u64 r2 = pack2(rsqrt_approx(vx), rsqrt_approx(vy));
      const u64 th2 = fmul2(pack2(vx, vy), r2);
      if (QUANT) {
        const u64 nd = fmul2(th2, k_neg), a2 = pack2(yv[i].x, yv[i].y);
        u64 qq = fmul2(a2, r2);
        qq = ffma2(ffma2(nd, qq, a2), r2, qq);
        qq = ffma2(ffma2(nd, qq, a2), r2, qq);
        float qx, qy;
        unpack2(qq, qx, qy);
        __stcs(&qv[32 * i], make_int2(__float2int_rn(qx), __float2int_rn(qy)));   // streaming: keep y in L2, not q
      }
      if (THR) __stcs(reinterpret_cast<u64*>(&tv[32 * i]), th2);
    } else {
      float v[C];
#pragma unroll
      for (int c = 0; c < C; ++c) v[c] = 0.f;
#pragma unroll
      for (int sl = 0; sl < 3; ++sl) {
        if (m & (1u << sl)) {
          const VF g = *reinterpret_cast<const VF*>(gp + sl * GS);
          const float* ga = reinterpret_cast<const float*>(&g);
          const float w = sl == 0 ? f4i.x : (sl == 1 ? f4i.y : f4i.z);
#pragma unroll
          for (int c = 0; c < C; ++c) v[c] = fmaf(ga[c], w, v[c]);
        }
      }
      const float* ya = reinterpret_cast<const float*>(&yv[i]);
      VF thr_v;
      VI q_v;
      float* th = reinterpret_cast<float*>(&thr_v);
      int32_t* qa = reinterpret_cast<int32_t*>(&q_v);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float vc = fmaxf(eps_s2, v[c]);
        const float rs = rsqrt_approx(vc);
        th[c] = vc * rs;
        if (QUANT) {
          float qq = ya[c] * rs;
          qq = fmaf(fmaf(-th[c], qq, ya[c]), rs, qq);
          qq = fmaf(fmaf(-th[c], qq, ya[c]), rs, qq);
          qa[c] = __float2int_rn(qq);
        }
      }
      if (QUANT) __stcs(&qv[32 * i], q_v);
      if (THR) __stcs(&tv[32 * i], thr_v);
    }
  }
}

// Mono: TWO consecutive frame rows per call - items (2p, 2p + 1) are adjacent in G, so the pair runs on the packed
// fp32 path of the stereo code (half the arithmetic instructions of two scalar rows); y / thr / q of the two rows are
// separate 4-byte accesses.  row_b: the second row exists (the last row of a ragged tile may be alone).
template <bool QUANT, bool THR, bool FILT_SMEM, int KI>
__device__ __forceinline__ void phase_d_unit_mono2(const float4* __restrict__ filt4, const unsigned masks,
                                                   const float (&ya)[KI], const float (&yb)[KI], float* __restrict__ ta,
                                                   float* __restrict__ tb_, int32_t* __restrict__ qa,
                                                   int32_t* __restrict__ qb, const float* gr, const float eps_s2,
                                                   const bool row_b) {
  constexpr int GS = kGS;
  const u64 k_neg = pack2(-1.f, -1.f);
#pragma unroll
  for (int i = 0; i < KI; ++i) {
    const float4 f4i = FILT_SMEM ? filt4[32 * i] : __ldg(filt4 + 32 * i);
    const float* gp = gr + __float_as_int(f4i.w) * GS;
    const unsigned m = masks >> (3 * i);
    u64 v2;                                    // warp-uniform branch on the slot pattern of the 32-filter group
    if ((m & 7u) == 3u) {
      v2 = ffma2(*reinterpret_cast<const u64*>(gp + GS), pack2(f4i.y, f4i.y),
                 fmul2(*reinterpret_cast<const u64*>(gp), pack2(f4i.x, f4i.x)));
    } else if ((m & 7u) == 6u) {
      v2 = ffma2(*reinterpret_cast<const u64*>(gp + 2 * GS), pack2(f4i.z, f4i.z),
                 fmul2(*reinterpret_cast<const u64*>(gp + GS), pack2(f4i.y, f4i.y)));
    } else {
      v2 = ffma2(*reinterpret_cast<const u64*>(gp + 2 * GS), pack2(f4i.z, f4i.z),
                 ffma2(*reinterpret_cast<const u64*>(gp + GS), pack2(f4i.y, f4i.y),
                       fmul2(*reinterpret_cast<const u64*>(gp), pack2(f4i.x, f4i.x))));
    }
    float vx, vy;
    unpack2(v2, vx, vy);
    vx = fmaxf(eps_s2, vx);
    vy = fmaxf(eps_s2, vy);
    const u64 r2 = pack2(rsqrt_approx(vx), rsqrt_approx(vy));
    const u64 th2 = fmul2(pack2(vx, vy), r2);
    float tx, ty;
    unpack2(th2, tx, ty);
    if (QUANT) {
      const u64 nd = fmul2(th2, k_neg), a2 = pack2(ya[i], yb[i]);
      u64 qq = fmul2(a2, r2);
      qq = ffma2(ffma2(nd, qq, a2), r2, qq);
      qq = ffma2(ffma2(nd, qq, a2), r2, qq);
      float qx, qy;
      unpack2(qq, qx, qy);
      __stcs(qa + 32 * i, __float2int_rn(qx));          // streaming: keep y in L2, not q
      if (row_b) __stcs(qb + 32 * i, __float2int_rn(qy));
    }
    if (THR) {
      __stcs(ta + 32 * i, tx);
      if (row_b) __stcs(tb_ + 32 * i, ty);
    }
  }
}

// R item PAIRS per call (stereo: the two channels of R consecutive frame rows; mono: R pairs of consecutive rows) share
// the filter-table entry, the slot-pattern branch and the address of the band rows of G (pair h's items sit 2 h floats
// further), and their dependency chains interleave: fewer shared-memory wavefronts and control instructions per
// coefficient than one pair at a time.  row_stride in floats; yv[h][i] = the pair's two amplitudes of filter 32 i + lane.
// GUARD: pair h is stored only when h < pairs_live (ragged last tile of the fused encoder).
// CLAMP: max(eps, .) before the square root (false when the plan's quiet threshold keeps every filter above eps)
template <int C, bool QUANT, bool THR, bool FILT_SMEM, int KI, int R, bool GUARD = false, bool CLAMP = true>
__device__ __forceinline__ void phase_d_unit_pairs(const float4* __restrict__ filt4, const unsigned masks,
                                                    const float2 (&yv)[R][KI], float* __restrict__ t0,
                                                    int32_t* __restrict__ q0, const size_t row_stride, const float* gr,
                                                    const float eps_s2, const int pairs_live = R) {
  constexpr int GS = kGS;
  const u64 k_neg = pack2(-1.f, -1.f);
#pragma unroll
  for (int i = 0; i < KI; ++i) {
    const float4 f4i = FILT_SMEM ? filt4[32 * i] : __ldg(filt4 + 32 * i);
    const float* gp = gr + __float_as_int(f4i.w) * GS;
    const unsigned m = masks >> (3 * i);
    u64 v[R];                                  // warp-uniform branch on the slot pattern of the 32-filter group
    if ((m & 7u) == 3u) {
      const u64 w0 = pack2(f4i.x, f4i.x), w1 = pack2(f4i.y, f4i.y);
#pragma unroll
      for (int h = 0; h < R; ++h)
        v[h] = ffma2(*reinterpret_cast<const u64*>(gp + GS + 2 * h), w1, fmul2(*reinterpret_cast<const u64*>(gp + 2 * h), w0));
    } else if ((m & 7u) == 6u) {
      const u64 w1 = pack2(f4i.y, f4i.y), w2 = pack2(f4i.z, f4i.z);
#pragma unroll
      for (int h = 0; h < R; ++h)
        v[h] = ffma2(*reinterpret_cast<const u64*>(gp + 2 * GS + 2 * h), w2,
                     fmul2(*reinterpret_cast<const u64*>(gp + GS + 2 * h), w1));
    } else {
      const u64 w0 = pack2(f4i.x, f4i.x), w1 = pack2(f4i.y, f4i.y), w2 = pack2(f4i.z, f4i.z);
#pragma unroll
      for (int h = 0; h < R; ++h)
        v[h] = ffma2(*reinterpret_cast<const u64*>(gp + 2 * GS + 2 * h), w2,
                     ffma2(*reinterpret_cast<const u64*>(gp + GS + 2 * h), w1,
                           fmul2(*reinterpret_cast<const u64*>(gp + 2 * h), w0)));
    }
#pragma unroll
    for (int h = 0; h < R; ++h) {
      if (GUARD && h >= pairs_live) break;
      float vx, vy;
      unpack2(v[h], vx, vy);
      if constexpr (CLAMP) {
        vx = fmaxf(eps_s2, vx);
        vy = fmaxf(eps_s2, vy);
      }
      const u64 r2 = pack2(rsqrt_approx(vx), rsqrt_approx(vy));
      const u64 th2 = fmul2(pack2(vx, vy), r2);
      if (QUANT) {
        const u64 nd = fmul2(th2, k_neg), a2 = pack2(yv[h][i].x, yv[h][i].y);
        u64 qq = fmul2(a2, r2);
        qq = ffma2(ffma2(nd, qq, a2), r2, qq);
        qq = ffma2(ffma2(nd, qq, a2), r2, qq);
        float qx, qy;
        unpack2(qq, qx, qy);
        if constexpr (C == 2) {
          __stcs(reinterpret_cast<int2*>(q0 + h * row_stride) + 32 * i, make_int2(__float2int_rn(qx), __float2int_rn(qy)));
        } else {
          __stcs(q0 + (2 * h) * row_stride + 32 * i, __float2int_rn(qx));
          __stcs(q0 + (2 * h + 1) * row_stride + 32 * i, __float2int_rn(qy));
        }
      }
      if (THR) {                                // streaming stores: keep y in L2, not thr / q
        if constexpr (C == 2) {
          __stcs(reinterpret_cast<u64*>(t0 + h * row_stride) + 32 * i, th2);
        } else {
          float tx, ty;
          unpack2(th2, tx, ty);
          __stcs(t0 + (2 * h) * row_stride + 32 * i, tx);
          __stcs(t0 + (2 * h + 1) * row_stride + 32 * i, ty);
        }
      }
    }
  }
}

// ---- tcgen05 (5th-generation tensor cores): the spreading product of a tile as 24 single-thread instructions ----------
// acc[item][j] = sum_i P[item][i] S[i][j] is M = 64 items x N = 64 bands x K = 64 bands: eight k-steps of
// tcgen05.mma.cta_group::1.kind::tf32 (M = 64, N = 64, K = 8), three instructions per k-step for the error-compensated
// product (P_lo S_hi + P_hi S_lo + P_hi S_hi), accumulator in 64 columns of tensor memory.  Operands
// (tools/microbench/tcgen05_toeplitz.cu measured both forms on a B200: 6.6e-7 / 1.1e-6 against float64):
//   A = P, MN-major (items contiguous, as the band-sum phase stores it), SWIZZLE_128B_BASE32B: band position kp, item i at
//       byte (i >> 5) 8192 + (kp >> 2) 512 + (kp & 3) 128 + (((i & 31) 4) ^ ((kp & 3) << 5)); descriptor: leading byte
//       offset 8192 (the next 32 items), stride byte offset 512 (the next four bands).  P_hi is P as stored (the tensor
//       core ignores the low 13 mantissa bits), P_lo = P - trunc(P) is written to the chunk buffer that is dead during the
//       product.
//   B = S, K-major, no swizzle.  S is Toeplitz (S[i][j] = f[64 - i + j]): the core matrix (8 n-rows x 4 k) of column
//       block nb and k-block kb only depends on c = 2 nb - kb, so the 128 cores of the dense operand are 30 distinct ones,
//       and because a descriptor addresses cores affinely (start + nb SBO + kb LBO) a 3840-byte table serves: with the
//       two K-cores of a k-step in DESCENDING band order (band i sits at position kp = i ^ 4 of A) c grows with both
//       steps: LBO = 128 B, SBO = 256 B, start = table + (14 - 2 ks) 128.  Core c, row r, element e = f[4 + 4 c + r - e].
// The accumulator comes back with tcgen05.ld.16x256b in the mma.sync fragment layout (thread (g, t): rows g, g + 8,
// columns 2t, 2t + 1 of every 8-column block), so the epilogue is the one of the mma.sync path.
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return static_cast<uint64_t>((saddr >> 4) & 0x3fffu) | (static_cast<uint64_t>((lbo >> 4) & 0x3fffu) << 16) |
         (static_cast<uint64_t>((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46) | (static_cast<uint64_t>(layout & 7u) << 61);
}
// kind::tf32 (a / b format 2), fp32 accumulate (c format 1), A MN-major (bit 15), B K-major, N = 64 (>> 3), M = 64 (>> 4)
constexpr uint32_t kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((kNB >> 3) << 17) | ((kTI >> 4) << 24);
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kTcIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_mbar(uint32_t mbar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n"
      "@!q bra WAIT_%=;\n"
      "}\n" ::"r"(mbar),
      "r"(parity)
      : "memory");
}
// 16 rows x 32 columns of the accumulator: v[4 i + e] = mma.sync element e of n-tile i
__device__ __forceinline__ void tc_ld_16x256b_x4(uint32_t taddr, float (&acc)[4][4]) {
  uint32_t v[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[i][e] = __uint_as_float(v[4 * i + e]);
}

#ifdef AC_PA_TRACE
// development build only (tools/k3_trace.py): per CTA and tile, the global timer at the start of the chunk loop, of the
// MMA phase, of phase D and at the end of phase D; slot 0 of a CTA holds its SM id
__device__ unsigned long long* g_pa_trace = nullptr;
__device__ __forceinline__ void pa_trace(int slot, int ev) {
  if (threadIdx.x == 0 && g_pa_trace != nullptr && slot < 15) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_pa_trace[(static_cast<size_t>(blockIdx.x) * 16 + 1 + slot) * 4 + ev] = t;
  }
}
#define PA_TRACE(slot, ev) pa_trace(slot, ev)
#else
#define PA_TRACE(slot, ev)
#endif

// FUSED (the single-pass encoder; stereo, filters_n = 256): `y` is the SIGNAL x [B, S, 2].  A tile is
// 32 consecutive frames of one batch row: their 33 blocks arrive by one bulk copy, the forward MDCT of
// mdct_forward_tile_kernel (mdctransformer.py:61-125) runs in place on the block rows and its post-twiddle stores the
// amplitudes TRANSPOSED into the same region (T[filter][item], all filters resident) once every group has finished its
// FFT exchanges - Y never goes to global memory.  The phases below then read T exactly like a chunk buffer, phase D
// takes its amplitudes from T instead of L2, and the next tile's blocks are requested when phase D is done.
template <int C, bool QUANT, int NFIX, int MINB, bool FUSED = false, bool TC = false>
__global__ void __launch_bounds__(kThreads, MINB)
pa_mma_tile_kernel(const __grid_constant__ PaDeviceTables tb, const __grid_constant__ PaJobParams jp,
                   const float* __restrict__ y, const float* __restrict__ ton_in, float one_minus_drown, float thr_scale,
                   float* __restrict__ thr_out, int32_t* __restrict__ q_out, int64_t frames_total, int64_t tiles,
                   unsigned* __restrict__ sched, const int ablate, const MdctDeviceTables mt, const int blocks_n,
                   const int tiles_per_row) {
  using VF = typename Vec<C>::F;
  using VI = typename Vec<C>::I;
  constexpr int TI = kTI, TS = kTS, GS = kGS, PS = kPS;
  constexpr int FT = TI / C;                    // frames per tile
  constexpr int ROWS = FT / kWarps;             // frame rows per warp
  static_assert(FT % kWarps == 0, "tile shape");
  static_assert(!FUSED || (C == 2 && NFIX == 256), "the fused encoder is built for stereo, filters_n = 256");
  extern __shared__ __align__(1024) float sm[];  // TC: P and the chunk buffers are swizzled operand tiles (1024-byte atoms)
  const Layout2 L = layout2(tb, FUSED, TC, NFIX);
  const int n = NFIX > 0 ? NFIX : tb.n, kc = n < kChunk ? n : kChunk;
  const int n_chunks = tb.mma_n_chunks;
  float2* s_powa = reinterpret_cast<float2*>(sm + L.powa);
  float2* s_powia = reinterpret_cast<float2*>(sm + L.powia);
  float* P = sm + L.p;                          // [64][PS], column item ^ ((band & 3) << 3)
  float* s_part = sm + L.part;                  // [2][kWarps][TI]: per-warp partial tonality sums
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
  {
    // the tables: every global load of a thread is issued before its first shared-memory store (one round trip to L2
    // instead of one per table); kThreads == 256 == entries of the exponent tables
    static_assert(kThreads == 256, "one exponent-table entry per thread");
    const bool pow_tabs = pow_tables_needed(tb);
    const float2 pa = pow_tabs ? tb.pow_alpha[tid] : make_float2(0.f, 0.f);
    const float2 pia = pow_tabs ? tb.pow_inv_alpha[tid] : make_float2(0.f, 0.f);
    const float sfv = tb.spread_fn[tid & 127];
    float zv[4];                                // TC: this thread's entries of the Toeplitz core table (see tc_desc)
    if constexpr (TC) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = tid + u * kThreads;
        zv[u] = i < kTcTable ? tb.spread_fn[4 + 4 * (i >> 5) + ((i >> 2) & 7) - (i & 3)] : 0.f;
      }
    }
    const float qv = tb.quiet[tid & (kNB - 1)], lv = tb.lin[tid & (kNB - 1)];
    const int nw = tb.n_mma_w4;
    float wv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) wv[u] = tid + u * kThreads < nw ? tb.mma_w4[tid + u * kThreads] : 0.f;
    float4 fv[2];
#pragma unroll
    for (int u = 0; u < 2; ++u)
      fv[u] = filt_smem && tid + u * kThreads < n ? tb.filt4[tid + u * kThreads] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (pow_tabs) {
      s_powa[tid] = pa;
      s_powia[tid] = pia;
    }
    if constexpr (TC) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = tid + u * kThreads;
        if (i < kTcTable) {
          uint32_t hi, lo;
          split_tf32(zv[u], hi, lo);
          s_sfh[i] = hi;
          s_sfl[i] = lo;
        }
      }
    } else if (tid < 128) {
      uint32_t hi, lo;
      split_tf32(sfv, hi, lo);
      s_sfh[tid] = hi;
      s_sfl[tid] = lo;
    }
    if (tid < kNB) {
      s_quiet[tid] = qv * scale2;
      s_lin[tid] = lv;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (tid + u * kThreads < nw) {
        s_bw8[2 * (tid + u * kThreads)] = wv[u];
        s_bw8[2 * (tid + u * kThreads) + 1] = wv[u];
      }
    for (int i = tid + 4 * kThreads; i < nw; i += kThreads) {      // long weight lists (large filters_n)
      const float w = tb.mma_w4[i];
      s_bw8[2 * i] = w;
      s_bw8[2 * i + 1] = w;
    }
    if (filt_smem) {
#pragma unroll
      for (int u = 0; u < 2; ++u)
        if (tid + u * kThreads < n) s_filt4[tid + u * kThreads] = fv[u];
    }
  }

  // the three zero rows behind a whole chunk are never written by the copies: zeroed once, in both buffers (a short
  // last chunk zeroes its own rows in load_chunk; the barriers of the chunk loop order these stores before any read)
  if constexpr (FUSED) {
    for (int i = tid; i < 3 * TS; i += kThreads) sm[L.t + n * TS + i] = 0.f;   // behind the last filter; never overwritten
  } else {
    for (int i = tid; i < 2 * 3 * TS; i += kThreads)
      sm[L.t + (i / (3 * TS)) * L.tbuf + kc * TS + (i % (3 * TS))] = 0.f;
  }
  uint64_t& s_mbar = *reinterpret_cast<uint64_t*>(sm + L.misc + 4);     // FUSED: completion of the tile's bulk copy
  if (FUSED && tid == 0) {
    mbar_init(&s_mbar, 1);
    mbar_fence_init();
  }
  if constexpr (FUSED) __syncthreads();
  // TC: 64 columns of tensor memory for the accumulator of the spreading product (three CTAs per SM: 192 of 512), the
  // mbarrier its completion arrives on
  uint32_t tc_tmem = 0, tc_parity = 0;
  const uint32_t tc_mbar = static_cast<uint32_t>(__cvta_generic_to_shared(sm + L.misc));
  if constexpr (TC) {
    if (tid == 0) {
      mbar_init(reinterpret_cast<uint64_t*>(sm + L.misc), 1);
      mbar_fence_init();
    }
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(tc_mbar + 8u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    tc_tmem = *reinterpret_cast<volatile uint32_t*>(sm + L.misc + 2);
  }

  const uint32_t sm_base = static_cast<uint32_t>(__cvta_generic_to_shared(sm));
  const uint32_t t_base = sm_base + static_cast<uint32_t>(L.t) * 4u;
  const uint32_t tbuf_bytes = static_cast<uint32_t>(L.tbuf) * 4u;
  const uint32_t w_base = sm_base + static_cast<uint32_t>(L.bw8) * 4u;
  const uint32_t p_base = sm_base + static_cast<uint32_t>(L.p) * 4u;
  const uint32_t tc_lane_off = ((static_cast<uint32_t>(lane) & 15u) << 3) | ((static_cast<uint32_t>(lane) >> 4) << 13);
  const float eps = tb.eps;
  const bool pow_split = tb.pow_split != 0;
  const float eps_s2 = eps * scale2;
  const float log2_s2 = 2.0f * log2f(scale);

  // y[f0 .. f0 + FT)[kc0 .. kc0 + kcn) -> T[buffer][filter][item], asynchronously.  Warp w copies its ROWS frame rows,
  // lanes run along the filters (coalesced reads); frame rows behind the end of the tensor re-read the last frame
  // (their results are never stored).  The three rows behind the chunk are zeroed for the 4-filter steps.
  // the copies of y ask L2 to keep their lines (evict_last) until phase D has re-read them (evict_first)
  uint64_t l2_policy, l2_policy_first;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(l2_policy));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(l2_policy_first));
  auto load_chunk = [&](int64_t f0, int nf, int chunk, int buf) {
    if (ablate & 32) return;
    const int kc0 = chunk * kChunk;
    const int kcn = (n - kc0 < kc ? n - kc0 : kc);
    if (kcn != kc) {                            // a short last chunk: its zero rows sit inside the data rows of the others
      float* tz = sm + L.t + buf * L.tbuf + kcn * TS;
      for (int i = tid; i < 3 * TS; i += kThreads) tz[i] = 0.f;
    }
    const uint32_t dst = t_base + static_cast<uint32_t>(buf) * tbuf_bytes + static_cast<uint32_t>(lane) * (TS * 4u) +
                         static_cast<uint32_t>(warp * ROWS * C) * 4u;
    const float* src = y + ((f0 + warp * ROWS) * static_cast<int64_t>(n) + kc0 + lane) * C;
    const size_t rs = static_cast<size_t>(n) * C;          // row stride in floats (an immediate when N is fixed)
    if (nf == FT && kcn == 64) {                // whole tile, whole 64-filter chunk: straight-line copies
#pragma unroll
      for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int kk = 0; kk < 64; kk += 32) {
          if constexpr (C == 1) {
            cp_async_hint<4>(dst + r * 4 + kk * (TS * 4), src + r * rs + kk, l2_policy);
          } else {
#pragma unroll
            for (int c2 = 0; c2 < C; c2 += 2)
              cp_async_hint<8>(dst + (r * C + c2) * 4 + kk * (TS * 4), src + r * rs + kk * C + c2, l2_policy);
          }
        }
    } else if (nf == FT && kcn == 32) {         // whole tile, whole 32-filter chunk
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        if constexpr (C == 1) {
          cp_async<4>(dst + r * 4, src + r * rs);
        } else {
#pragma unroll
          for (int c2 = 0; c2 < C; c2 += 2) cp_async<8>(dst + (r * C + c2) * 4, src + r * rs + c2);
        }
      }
    } else {
#pragma unroll 1
      for (int r = 0; r < ROWS; ++r) {
        const int fl = warp * ROWS + r;
        const float* srow = src + (fl < nf ? r : nf - 1 - warp * ROWS) * static_cast<int64_t>(rs);
        for (int kk = 0; kk + lane < kcn; kk += 32) {
          if constexpr (C == 1) {
            cp_async<4>(dst + r * 4 + kk * (TS * 4), srow + kk);
          } else {
#pragma unroll
            for (int c2 = 0; c2 < C; c2 += 2) cp_async<8>(dst + (r * C + c2) * 4 + kk * (TS * 4), srow + kk * C + c2);
          }
        }
      }
    }
    cp_async_commit();
  };

  // tiles are walked from the END of the tensor: the producer of y (the forward MDCT) wrote its last ~100 MB into
  // L2 most recently, and the consumer of thr / q (the inverse MDCT) starts at the front, where this kernel ends
  // Tile scheduling: the first tile of a CTA is its block index, every further one comes from a global ticket
  // counter (sched[0]) - the CTAs of an SM do not progress at the same rate (the warp schedulers favour the older
  // ones: measured 18.5 against 21.9 us per tile), and with a static split the SM idles while its slowest CTA
  // finishes.  A ticket is drawn at the start of a tile and published through shared memory before the barrier of the
  // tile's last chunk, behind which the first chunk of the next tile is requested.  The last CTA to finish
  // (sched[1]) re-arms both counters for the next launch.
  volatile long long* s_next = reinterpret_cast<volatile long long*>(sm + L.misc + 8);
  int par = 0;                                  // buffer of the chunk that is processed next
  int tpar = 0;                                 // parity of the tile (slot of s_next)
  int64_t tile_i = blockIdx.x;
  // FUSED: the blocks f0l - 1 .. f0l + FT - 1 of batch row b -> rows 0 .. FT of the T region (2 KB each, as in
  // mdct_forward_tile_kernel); rows outside the signal are zero-filled by the consumer.  Thread 0 only.
  const int frames_row = blocks_n + 1;          // FUSED: frames per batch row
  auto issue_x_load = [&](int64_t ti) {
    if constexpr (FUSED) {
      constexpr int ROWF = 256 * C;             // floats per block row
      const int64_t tile = tiles - 1 - ti;
      const int64_t b = tile / tiles_per_row;
      const int f0l = static_cast<int>(tile - b * tiles_per_row) * FT;
      const int r_lo = f0l == 0 ? 1 : 0;
      const int r_hi = min(FT + 1, blocks_n - f0l + 1);
      if (r_hi > r_lo) {
        const uint32_t bytes = static_cast<uint32_t>(r_hi - r_lo) * ROWF * sizeof(float);
        mbar_arrive_expect_tx(&s_mbar, bytes);
        bulk_load(sm + L.t + r_lo * ROWF, y + (b * blocks_n + (f0l - 1 + r_lo)) * static_cast<int64_t>(ROWF), bytes, &s_mbar);
      } else {
        mbar_arrive(&s_mbar);
      }
    }
  };
  if (!(ablate & 128)) pdl_launch_dependents(); // the next kernel on the stream may start its prologue (128: experiment, at the end)
  pdl_wait();                                   // the producer of y (x) has completed; the tables above are plan constants
  if (tile_i < tiles) {
    if constexpr (FUSED) {
      if (tid == 0) issue_x_load(tile_i);
    } else {
      const int64_t f0 = (tiles - 1 - tile_i) * FT;
      load_chunk(f0, static_cast<int>(frames_total - f0 < FT ? frames_total - f0 : FT), 0, 0);
    }
  }
  int slot_i = 0;
  while (tile_i < tiles) {
    const int64_t tile = tiles - 1 - tile_i;
    int64_t f0;                                 // first frame of the tile in the flattened [B F] frame axis
    int nf;                                     // frames of the tile inside the tensor
    if constexpr (FUSED) {
      const int64_t b = tile / tiles_per_row;
      const int f0l = static_cast<int>(tile - b * tiles_per_row) * FT;
      f0 = b * frames_row + f0l;
      nf = min(FT, frames_row - f0l);
    } else {
      f0 = tile * FT;
      nf = static_cast<int>(frames_total - f0 < FT ? frames_total - f0 : FT);
    }
    long long ticket = 0;
    if (tid == 0) ticket = static_cast<long long>(gridDim.x) + atomicAdd(sched, 1u);
    int64_t next_i = tiles;

    if constexpr (FUSED) {
      // ---- forward MDCT of the tile's frames, in place (mdct_forward_tile_kernel with 32 groups of 8 threads)
      using Plan = Plan256;
      constexpr int M = Plan::M, N = 2 * M, H = M, T = Plan::T, E = Plan::E, R0 = Plan::R0, ROWF = N * C;
      static_assert(kThreads / T == FT, "one group of threads per frame");
      const int64_t b = tile / tiles_per_row;
      const int f0l = static_cast<int>(tile - b * tiles_per_row) * FT;
      const int g = tid / T, t = tid % T, variant = (tid >> 3) & 1;
      float* buf = sm + L.t;
      float* prev = buf + g * ROWF;             // block row before this group's frame
      float* cur = prev + ROWF;                 // the frame's own block: input, FFT scratch
      const int r_lo = f0l == 0 ? 1 : 0;
      const int r_hi = min(FT + 1, blocks_n - f0l + 1);
      mbar_wait(&s_mbar, slot_i & 1);
      if (r_lo > 0 || r_hi < FT + 1) {          // blocks outside the signal are zero (mdctransformer.py:366)
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r_lo > 0)
          for (int i = tid * 4; i < ROWF; i += kThreads * 4) *reinterpret_cast<float4*>(buf + i) = z;
        for (int i = max(r_hi, r_lo) * ROWF + tid * 4; i < (FT + 1) * ROWF; i += kThreads * 4)
          *reinterpret_cast<float4*>(buf + i) = z;
        __syncthreads();
      }
      float2 v0[E], v1[E];
#pragma unroll
      for (int s = 0; s < E; ++s) {             // window + fold + pre-twiddle (mdctransformer.py:118, H)
        const int nn = Plan::in_index(t, s);
        constexpr int kHalf = R0 / 2;
        const bool low = (s % R0) < kHalf;
        const int p = low ? H - 1 - 2 * nn : 2 * nn - H;
        const int a1 = variant ? N - 1 - p : p;
        const int a2 = (N - 1) - a1;
        const float2 l0 = ld2<C, ROWF>(prev, a1), l1 = ld2<C, ROWF>(prev, a2);
        const float2 l2 = ld2<C, ROWF>(cur, a1), l3 = ld2<C, ROWF>(cur, a2);
        const float4 kr = __ldg(&mt.pre_fwd[(variant * 2) * M + nn]);
        const float4 ki = __ldg(&mt.pre_fwd[(variant * 2 + 1) * M + nn]);
        v0[s].x = fmaf(l3.x, kr.w, fmaf(l2.x, kr.z, fmaf(l1.x, kr.y, l0.x * kr.x)));
        v0[s].y = fmaf(l3.x, ki.w, fmaf(l2.x, ki.z, fmaf(l1.x, ki.y, l0.x * ki.x)));
        v1[s].x = fmaf(l3.y, kr.w, fmaf(l2.y, kr.z, fmaf(l1.y, kr.y, l0.y * kr.x)));
        v1[s].y = fmaf(l3.y, ki.w, fmaf(l2.y, ki.z, fmaf(l1.y, ki.y, l0.y * ki.x)));
      }
      __syncthreads();                          // every block row has been read: the rows become FFT scratch
      fft2<Plan>(v0, v1, reinterpret_cast<float4*>(cur), t, g, mt.tw_pass1, mt.tw_pass2);
      __syncthreads();                          // every exchange is over: the region becomes T[filter][item]
      float* tcol = buf + 2 * g;                // items 2 g, 2 g + 1 = the two channels of frame g
#pragma unroll
      for (int s = 0; s < E; ++s) {             // post-twiddle; bin k -> filters 2k and N-1-2k
        const int k = Plan::out_index(t, s);
        const float4 c4 = __ldg(&mt.post_fwd[variant * M + k]);
        const int i1 = variant ? N - 1 - 2 * k : 2 * k;
        const int i2 = (N - 1) - i1;
        *reinterpret_cast<float2*>(tcol + i1 * TS) =
            make_float2(fmaf(v0[s].y, c4.y, v0[s].x * c4.x), fmaf(v1[s].y, c4.y, v1[s].x * c4.x));
        *reinterpret_cast<float2*>(tcol + i2 * TS) =
            make_float2(fmaf(v0[s].y, c4.w, v0[s].x * c4.z), fmaf(v1[s].y, c4.w, v1[s].x * c4.z));
      }
      // no barrier here: the first barrier of the chunk loop follows
    }

    PA_TRACE(slot_i, 0);
    u64 ton_i2 = 0ull, ton_l2 = 0ull;           // tonality sums (psychoacoustic.py:113-116) of the item pair of this lane
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
      if (chunk + 1 == n_chunks && tid == 0) s_next[tpar] = ticket;
      if constexpr (!FUSED) cp_async_wait_all();
      __syncthreads();                          // the chunk has landed; the other buffer is free
      if constexpr (FUSED) {                    // (T complete / the P partials of the previous chunk are visible)
        if (chunk + 1 == n_chunks) next_i = s_next[tpar];
      } else if (chunk + 1 < n_chunks) {
        load_chunk(f0, nf, chunk + 1, par ^ 1);
      } else {
        next_i = s_next[tpar];
        if (next_i < tiles) {
          const int64_t nf0 = (tiles - 1 - next_i) * FT;
          load_chunk(nf0, static_cast<int>(frames_total - nf0 < FT ? frames_total - nf0 : FT), 0, par ^ 1);
        }
      }
      const uint32_t t_lane = FUSED ? t_base + static_cast<uint32_t>(chunk * kChunk) * (TS * 4u) + static_cast<uint32_t>(lane) * 8u
                                    : t_base + static_cast<uint32_t>(par) * tbuf_bytes + static_cast<uint32_t>(lane) * 8u;

      // ---- A1: tonality sums over this warp's share of the filters: sum I and sum log2 max(eps, I)   (:113, :312)
      // two filters per logarithm: log2 a + log2 b = log2(a b), a b >= eps^2 (finite for |y| < 1e9)
      if (ton_in == nullptr && !(ablate & 1)) {
        auto lds_t = [&](int k) {
          u64 a2;
          asm volatile("ld.shared.b64 %0, [%1];" : "=l"(a2) : "r"(t_lane + static_cast<uint32_t>(k) * (TS * 4u)));
          return a2;
        };
        auto clamp2 = [&](u64 in2) {
          float ix, iy;
          unpack2(in2, ix, iy);
          return pack2(fmaxf(eps, ix), fmaxf(eps, iy));
        };
        // the filters of the chunk are dealt to the warps in contiguous runs that even out the band-sum jobs below
        // (PaJobParams::ton_start): a warp with few or no jobs in this chunk takes more of the tonality pass
        int k = jp.ton_start[chunk * 9 + warp_u];
        const int k1 = jp.ton_start[chunk * 9 + warp_u + 1];
#pragma unroll 2
        for (; k + 1 < k1; k += 2) {
          const u64 a2 = lds_t(k), b2 = lds_t(k + 1);
          const u64 ia = fmul2(a2, a2), ib = fmul2(b2, b2);
          ton_i2 = fadd2(ton_i2, fadd2(ia, ib));
          float px, py;
          unpack2(fmul2(clamp2(ia), clamp2(ib)), px, py);
          ton_l2 = fadd2(ton_l2, pack2(lg2_approx(px), lg2_approx(py)));
        }
        if (k < k1) {
          const u64 a2 = lds_t(k);
          const u64 in2 = fmul2(a2, a2);
          float ix, iy;
          unpack2(clamp2(in2), ix, iy);
          ton_i2 = fadd2(ton_i2, in2);
          ton_l2 = fadd2(ton_l2, pack2(lg2_approx(ix), lg2_approx(iy)));
        }
      }

      // ---- A2: band energies of this chunk; P = max(eps, I_bark)^alpha when a band is complete  (:204-206, :313)
      // lane l owns the item pair (2l, 2l + 1) as packed fp32: one LDS.64, a square and one FFMA2 per filter
      {
        const int j0 = jp.start[chunk * 9 + warp_u], j1 = (ablate & 2) ? j0 : jp.start[chunk * 9 + warp_u + 1];
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
            const u64 y0 = lds_b64<0>(tp), y4 = lds_b64<4 * TS * 4>(tp), y1 = lds_b64<1 * TS * 4>(tp);
            const u64 y5 = lds_b64<5 * TS * 4>(tp), y2 = lds_b64<2 * TS * 4>(tp), y6 = lds_b64<6 * TS * 4>(tp);
            const u64 y3 = lds_b64<3 * TS * 4>(tp), y7 = lds_b64<7 * TS * 4>(tp);
            a0 = ffma2(fmul2(y0, y0), w0, a0);
            a1 = ffma2(fmul2(y4, y4), w4, a1);
            a0 = ffma2(fmul2(y1, y1), w1, a0);
            a1 = ffma2(fmul2(y5, y5), w5, a1);
            a0 = ffma2(fmul2(y2, y2), w2, a0);
            a1 = ffma2(fmul2(y6, y6), w6, a1);
            a0 = ffma2(fmul2(y3, y3), w3, a0);
            a1 = ffma2(fmul2(y7, y7), w7, a1);
            tp += 8 * TS * 4;
            wp += 64;
          }
          if (s) {
            u64 w0, w1, w2, w3;
            lds_2b64<0>(wp, w0, w1);
            lds_2b64<16>(wp, w2, w3);
            const u64 y0 = lds_b64<0>(tp), y1 = lds_b64<1 * TS * 4>(tp), y2 = lds_b64<2 * TS * 4>(tp);
            const u64 y3 = lds_b64<3 * TS * 4>(tp);
            a0 = ffma2(fmul2(y0, y0), w0, a0);
            a1 = ffma2(fmul2(y1, y1), w1, a1);
            a0 = ffma2(fmul2(y2, y2), w2, a0);
            a1 = ffma2(fmul2(y3, y3), w3, a1);
          }
          u64 acc2 = fadd2(a0, a1);
          // TC: the operand layout of tc_desc (low 16 bits of the descriptor word; lanes 16 - 31 own items 32 - 63);
          // mma.sync: [band][item ^ swizzle] (bits 18 ..)
          const uint32_t pp = TC ? p_base + ((static_cast<uint32_t>(jb.w) & 0xffffu) ^ tc_lane_off)
                                 : p_base + ((static_cast<uint32_t>(jb.w) >> 18) ^ (static_cast<uint32_t>(lane) << 3));
          if (jb.w & 0x10000) acc2 = fadd2(acc2, lds_b64<0>(pp));
          if (jb.w & 0x20000) {
            float ax, ay;
            unpack2(acc2, ax, ay);
            ax = fmaxf(eps, ax);
            ay = fmaxf(eps, ay);
            if (pow_split) {                    // x^alpha = sqrt(x) x^(alpha - 1/2)
              const u64 e2 = fmul2(pack2(lg2_approx(ax), lg2_approx(ay)), pack2(tb.pow_c1, tb.pow_c1));
              float ex, ey;
              unpack2(e2, ex, ey);
              acc2 = fmul2(pack2(sqrt_approx(ax), sqrt_approx(ay)), pack2(ex2_approx(ex), ex2_approx(ey)));
            } else {
              acc2 = pack2(pow_tab(ax, tb.alpha, s_powa), pow_tab(ay, tb.alpha, s_powa));
            }
          }
          sts_b64(pp, acc2);
        }
      }
      par ^= 1;
    }
    float* G = sm + L.p;                        // [64][GS] over P and the tonality partials (dead after the MMA loop)
    if (ton_in == nullptr) {
      *reinterpret_cast<u64*>(s_part + warp * TI + 2 * lane) = ton_i2;
      *reinterpret_cast<u64*>(s_part + (kWarps + warp) * TI + 2 * lane) = ton_l2;
    }
    // The product is issued as two groups on the same accumulator: P_hi (S_lo + S_hi) over all k-steps right behind the
    // barrier that completes P, then P_lo S_hi.  P as it stands is P_hi (the tensor core ignores the low 13 mantissa bits);
    // P_lo = P - trunc(P) is made while the first group runs and goes
    //   chain kernel: into the chunk buffer that is dead until the next tile's second chunk is requested (`par` is the
    //     buffer the next tile's first chunk is landing in);
    //   single-pass encoder (no dead buffer, T holds the amplitudes): over P itself once the first group has read it.
    // The instruction order is the same, so both kernels produce the same bits.
    constexpr bool INPLACE = FUSED;
    const uint32_t lo_base = INPLACE ? p_base : t_base + static_cast<uint32_t>(par ^ 1) * tbuf_bytes;
    auto split_lo = [&](float4* l4) {
      const float4* p4 = reinterpret_cast<const float4*>(P);
#pragma unroll
      for (int u = 0; u < kNB * kTI / 4 / kThreads; ++u) {
        const float4 v = p4[tid + u * kThreads];
        float4 r;
        r.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
        r.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
        r.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
        r.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
        l4[tid + u * kThreads] = r;
      }
    };
    // descriptors of k-step 0; a k-step moves the start address (16-byte units in the low word) by +1024 bytes in P and
    // by -256 bytes in the core table - the issuing thread spends an add per operand, not a descriptor build
    const uint32_t zh = sm_base + static_cast<uint32_t>(L.sfh) * 4u, zl = sm_base + static_cast<uint32_t>(L.sfl) * 4u;
    const uint64_t d_phi = tc_desc(p_base, 8192, 512, 1), d_plo = tc_desc(lo_base, 8192, 512, 1);
    const uint64_t d_zh = tc_desc(zh + 14 * 128, 128, 256, 0), d_zl = tc_desc(zl + 14 * 128, 128, 256, 0);
    auto issue_hi = [&]() {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        tc_mma_tf32(tc_tmem, d_phi + static_cast<uint64_t>(ks * 64), d_zl - static_cast<uint64_t>(ks * 16), ks > 0 ? 1u : 0u);
        tc_mma_tf32(tc_tmem, d_phi + static_cast<uint64_t>(ks * 64), d_zh - static_cast<uint64_t>(ks * 16), 1u);
      }
    };
    auto issue_lo = [&]() {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        tc_mma_tf32(tc_tmem, d_plo + static_cast<uint64_t>(ks * 64), d_zh - static_cast<uint64_t>(ks * 16), 1u);
    };
    if constexpr (TC) {
      fence_async_smem();                       // this thread's stores of P (the band sums) -> async proxy
      tc_fence_before();                        // its reads of the previous tile's accumulator are over
    }
    __syncthreads();                            // P and the tonality partials are complete
    PA_TRACE(slot_i, 1);
    if constexpr (TC) {                         // the first group runs while the low-order term / the constants are made
      if (tid == kTcIssuer && !(ablate & 4)) {
        tc_fence_after();
        issue_hi();
        if constexpr (INPLACE) tc_commit(tc_mbar);
      }
    }

    // ---- per-item constants of the masking offset (psychoacoustic.py:185-191), once per tile
    if (warp < TI / 32) {
      const int it = warp * 32 + lane;
      float ton;
      if (ton_in == nullptr) {                         // tonality of the item (psychoacoustic.py:113-118)
        float s_i = 0.f, s_l = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
          s_i += s_part[w * TI + it];
          s_l += s_part[(kWarps + w) * TI + it];
        }
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

    // ---- B: spreading on the tensor cores                                 (psychoacoustic.py:195-206)
    {
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
      if constexpr (TC) {
        if constexpr (INPLACE) {
          if (!(ablate & 4)) {
            tc_wait_mbar(tc_mbar, tc_parity);   // the first group has read P
            tc_parity ^= 1u;
          }
          split_lo(reinterpret_cast<float4*>(P));
        } else {
          split_lo(reinterpret_cast<float4*>(sm + L.t + (par ^ 1) * L.tbuf));
        }
        fence_async_smem();                     // this thread's part of the low-order term -> async proxy
        __syncthreads();                        // the low-order term is complete; the per-item constants are visible
        if (!(ablate & 4)) {
          if (tid == kTcIssuer) {
            tc_fence_after();
            issue_lo();
            tc_commit(tc_mbar);
          }
          tc_wait_mbar(tc_mbar, tc_parity);
          tc_parity ^= 1u;
          tc_fence_after();
          tc_ld_16x256b_x4(tc_tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>(32 * nq), acc);
        }
      } else {
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
        if (ablate & 4) break;
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

      __syncthreads();                          // the per-item constants are visible; P is free
      }
      // masking offset, non-linear superposition, quiet threshold            (psychoacoustic.py:185-208, :144)
      const int ma = m0 + g, mb = m0 + g + 8;
      if (ablate & 8) {
      } else if (!tb.clamp_needed && pow_split) {
        // acc^(1/alpha) 2^(offset terms) = acc^2 2^((1/alpha - 2) lg2(acc) + offset_log2 offset + log2 scale^2)
        const float ua = s_u[ma], va = s_v[ma], ub = s_u[mb], vb = s_v[mb];
        const float c2 = tb.pow_c2;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int j = 32 * nq + 8 * nt + 2 * t;
          const float2 lin2 = *reinterpret_cast<const float2*>(s_lin + j);
          const float2 q2 = *reinterpret_cast<const float2*>(s_quiet + j);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float lin = (e & 1) ? lin2.y : lin2.x, qt = (e & 1) ? q2.y : q2.x;
            const float u = (e & 2) ? ub : ua, v = (e & 2) ? vb : va;
            const float a = acc[nt][e];
            const float f = fmaf(c2, lg2_approx(a), fmaf(u, lin, v));
            G[(j + (e & 1)) * GS + ((e & 2) ? mb : ma)] = fmaxf((a * a) * ex2_approx(f), qt);
          }
        }
      } else if (!tb.clamp_needed) {
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

    PA_TRACE(slot_i, 2);
    // compact side information: the tile's bark-domain thresholds G[band][frame, channel] -> bark_out[frame][band][channel]
    if (tb.bark_out != nullptr) {
#pragma unroll 1
      for (int fl = warp * ROWS; fl < (warp + 1) * ROWS && fl < nf; ++fl) {
        float* go = tb.bark_out + (f0 + fl) * (kNB * C);
        *reinterpret_cast<VF*>(go + lane * C) = *reinterpret_cast<const VF*>(G + lane * GS + fl * C);
        *reinterpret_cast<VF*>(go + (lane + 32) * C) = *reinterpret_cast<const VF*>(G + (lane + 32) * GS + fl * C);
      }
    }

    // ---- D: back to the filter bands, amplitude, optional quantiser      (:330-331; quantiser: SURVEY 8a row Q)
    // units of KI x 32 filters: the unit's filter-table entries stay in registers for all rows of the warp; a row
    // of a unit is KI x 32 x C contiguous floats of thr and of q (long DRAM bursts)
    {
      const bool thr = thr_out != nullptr;
      constexpr int KI = C == 4 ? 4 : 8;
      const int rows_live = (ablate & 16) ? 0 : min(ROWS, nf - warp * ROWS);
      const bool whole = C <= 2 && n % 64 == 0 && (FUSED || rows_live == ROWS);   // the common case, see below
      if (C == 1 && n % (32 * KI) == 0 && !whole) {
        // mono: pairs of consecutive rows on the packed path (phase_d_unit_mono2)
        if constexpr (C == 1) {
          const int upr = n / (32 * KI);
#pragma unroll 1
          for (int r = 0; r < rows_live; r += 2) {
            const bool row_b = r + 1 < rows_live;
            const int64_t offa = ((f0 + warp * ROWS + r) * static_cast<int64_t>(n) + lane);
            const int64_t offb = row_b ? offa + n : offa;          // a lone last row is read twice, stored once
            const float* gr = G + warp * ROWS + r;
#pragma unroll 1
            for (int u = 0; u < upr; ++u) {
              const int k0 = u * (32 * KI);
              float ya[KI], yb[KI];
              if (QUANT) {
#pragma unroll
                for (int i = 0; i < KI; ++i) {
                  ya[i] = __ldg(y + offa + k0 + 32 * i);
                  yb[i] = __ldg(y + offb + k0 + 32 * i);
                }
              }
              unsigned masks = 0;
#pragma unroll
              for (int i = 0; i < KI; ++i) masks |= static_cast<unsigned>(tb.filt_mask[k0 / 32 + i]) << (3 * i);
#define AC_PHASE_D(THR_, FS_)                                                                                       \
  phase_d_unit_mono2<QUANT, THR_, FS_, KI>((FS_ ? s_filt4 : tb.filt4) + k0 + lane, masks, ya, yb, thr_out + offa + k0, \
                                           thr_out + offb + k0, q_out + offa + k0, q_out + offb + k0, gr, eps_s2, row_b)
              if (filt_smem) {
                if (thr) AC_PHASE_D(true, true); else AC_PHASE_D(false, true);
              } else {
                if (thr) AC_PHASE_D(true, false); else AC_PHASE_D(false, false);
              }
#undef AC_PHASE_D
            }
          }
        }
      } else if (whole) {
        // whole tile, stereo / mono: the warp's four item pairs together (stereo: its four rows; mono: four of its eight
        // rows, as pairs), 64 filters at a time (phase_d_unit_pairs); measured against 2 pairs x 64 / 128 / 256 filters
        // and 4 pairs x 32 / 128 filters (profiles/README.md)
        if constexpr (C <= 2) {
          constexpr int R = 4, K2 = 2;
          const size_t rs = static_cast<size_t>(n) * C;
          const int upr = n / (32 * K2);
#pragma unroll 1
          for (int r0 = 0; r0 < ROWS; r0 += R * (2 / C)) {          // stereo: all four rows at once; mono: rows 0-3, 4-7... in pairs
            const int64_t off0 = ((f0 + warp * ROWS + r0) * static_cast<int64_t>(n) + lane) * C;
            const float* gr = G + (warp * ROWS + r0) * C;
#pragma unroll 1
            for (int u = 0; u < upr; ++u) {
              const int k0 = u * (32 * K2);
              const int64_t off = off0 + C * k0;
              float2 yv[R][K2];
              if constexpr (FUSED) {            // the amplitudes are still in T[filter][item]
                const float* tp = sm + L.t + (k0 + lane) * TS + (warp * ROWS + r0) * C;
#pragma unroll
                for (int h = 0; h < R; ++h)
#pragma unroll
                  for (int i = 0; i < K2; ++i) yv[h][i] = *reinterpret_cast<const float2*>(tp + 32 * i * TS + 2 * h);
              } else if (QUANT) {               // an L2 hit; loading a unit ahead changes nothing (measured)
#pragma unroll
                for (int h = 0; h < R; ++h)
#pragma unroll
                  for (int i = 0; i < K2; ++i) {
                    if constexpr (C == 2) {
                      asm volatile("ld.global.cg.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;"     // L2 only, last use
                                   : "=f"(yv[h][i].x), "=f"(yv[h][i].y)
                                   : "l"(reinterpret_cast<const float2*>(y + off + h * rs) + 32 * i), "l"(l2_policy_first));
                    } else {
                      asm volatile("ld.global.cg.L2::cache_hint.f32 %0, [%1], %2;"
                                   : "=f"(yv[h][i].x) : "l"(y + off + (2 * h) * rs + 32 * i), "l"(l2_policy_first));
                      asm volatile("ld.global.cg.L2::cache_hint.f32 %0, [%1], %2;"
                                   : "=f"(yv[h][i].y) : "l"(y + off + (2 * h + 1) * rs + 32 * i), "l"(l2_policy_first));
                    }
                  }
              }
              unsigned masks = 0;
#pragma unroll
              for (int i = 0; i < K2; ++i) masks |= static_cast<unsigned>(tb.filt_mask[k0 / 32 + i]) << (3 * i);
#define AC_PHASE_D(THR_, FS_, CL_)                                                                                           \
  phase_d_unit_pairs<C, QUANT, THR_, FS_, K2, R, FUSED, CL_>((FS_ ? s_filt4 : tb.filt4) + k0 + lane, masks, yv, thr_out + off, \
                                                             q_out + off, rs, gr, eps_s2, rows_live - r0)
              if (filt_smem && !tb.thr_clamp_needed) {      // the headline plans: no clamp, table in shared memory
                if (thr) AC_PHASE_D(true, true, false); else AC_PHASE_D(false, true, false);
              } else if (filt_smem) {
                if (thr) AC_PHASE_D(true, true, true); else AC_PHASE_D(false, true, true);
              } else if (!tb.thr_clamp_needed) {
                if (thr) AC_PHASE_D(true, false, false); else AC_PHASE_D(false, false, false);
              } else {
                if (thr) AC_PHASE_D(true, false, true); else AC_PHASE_D(false, false, true);
              }
#undef AC_PHASE_D
            }
          }
        }
      } else if (n % (32 * KI) == 0) {
        // the rows of a warp are contiguous in memory: its units are one flat sequence; the y values of unit u + 1
        // are loaded (an L2 hit) before unit u is computed
        const int upr = n / (32 * KI);          // units per row
        const int units = rows_live > 0 ? rows_live * upr : 0;
        const int64_t off0 = ((f0 + warp * ROWS) * static_cast<int64_t>(n) + lane) * C;
        const VF* yp = reinterpret_cast<const VF*>(y + off0);
        VF ynext[KI];
        if (QUANT && units > 0) {
#pragma unroll
          for (int i = 0; i < KI; ++i) ynext[i] = __ldg(yp + 32 * i);
        }
#pragma unroll 1
        for (int u = 0; u < units; ++u) {
          VF ycur[KI];
#pragma unroll
          for (int i = 0; i < KI; ++i) ycur[i] = ynext[i];
          if (QUANT && u + 1 < units) {
#pragma unroll
            for (int i = 0; i < KI; ++i) ynext[i] = __ldg(yp + (u + 1) * (32 * KI) + 32 * i);
          }
          const int r = u / upr, k0 = (u - r * upr) * (32 * KI);
          unsigned masks = 0;
#pragma unroll
          for (int i = 0; i < KI; ++i) masks |= static_cast<unsigned>(tb.filt_mask[k0 / 32 + i]) << (3 * i);
          const int64_t off = off0 + static_cast<int64_t>(u) * (32 * KI * C);
          const float* gr = G + (warp * ROWS + r) * C;
#define AC_PHASE_D(THR_, FS_) \
  phase_d_unit<C, QUANT, THR_, FS_, KI>((FS_ ? s_filt4 : tb.filt4) + k0 + lane, masks, ycur, thr_out + off, q_out + off, gr, eps_s2)
          if (filt_smem) {
            if (thr) AC_PHASE_D(true, true); else AC_PHASE_D(false, true);
          } else {
            if (thr) AC_PHASE_D(true, false); else AC_PHASE_D(false, false);
          }
#undef AC_PHASE_D
        }
      } else {
#pragma unroll 1
        for (int r = 0; r < rows_live; ++r) {
          const int fl = warp * ROWS + r;
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
    }
    // no barrier here: G's buffer and P are next written behind the first barrier of the next tile
#ifdef AC_PA_TRACE
    __syncthreads();
    PA_TRACE(slot_i, 3);
#endif
    if constexpr (FUSED) {
      fence_async_smem();                       // this thread's accesses to T are ordered before the next bulk copy
      __syncthreads();                          // phase D has read its amplitudes: T may receive the next blocks
      if (tid == 0 && next_i < tiles) issue_x_load(next_i);
    }
    tile_i = next_i;
    tpar ^= 1;
    ++slot_i;
  }
  if (ablate & 128) pdl_launch_dependents();
  if (tid == 0 && atomicAdd(sched + 1, 1u) == gridDim.x - 1) {   // every CTA has drawn its last ticket
    sched[0] = 0u;
    sched[1] = 0u;
  }
  if constexpr (TC) {
    tc_fence_before();
    __syncthreads();                            // every warp has read its last accumulator
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tc_tmem) : "memory");
  }
#ifdef AC_PA_TRACE
  if (threadIdx.x == 0 && g_pa_trace != nullptr) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_pa_trace[static_cast<size_t>(blockIdx.x) * 64] = smid;
  }
#endif
}

#ifdef AC_PA_TRACE
}  // namespace
}  // namespace ac
extern "C" __attribute__((visibility("default"))) int ac_debug_set_pa_trace(unsigned long long* buf) {
  return static_cast<int>(cudaMemcpyToSymbol(ac::g_pa_trace, &buf, sizeof(buf)));
}
namespace ac {
namespace {
#endif

// Ticket counters of the tile scheduler: a ring of {next ticket, finished CTAs} pairs owned by the plan (zeroed at
// creation, re-armed by the last CTA of every launch); launches take the slots round robin, so launches of one plan
// that overlap on different streams do not share a counter (up to kPaSchedSlots in flight)
unsigned* pa_sched_slot(const PaDeviceTables& tb) {
  static std::atomic<unsigned> next{0};
  return tb.sched + 2 * (next.fetch_add(1, std::memory_order_relaxed) % kPaSchedSlots);
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

template <int C, bool QUANT, int NFIX, int MINB, bool TC>
cudaError_t launch_mma_tile_nb(const PaDeviceTables& tb, const float* y, const float* ton_in, float omd, float thr_scale,
                               float* thr_out, int32_t* q_out, int64_t frames, int per_sm, cudaStream_t stream) {
  constexpr int FT = kTI / C;
  const size_t smem = static_cast<size_t>(layout2(tb, false, TC).total) * sizeof(float);
  const int64_t tiles = (frames + FT - 1) / FT;
  const int64_t cap = static_cast<int64_t>(mma_sm_count()) * per_sm;
  const unsigned grid = static_cast<unsigned>(tiles < cap ? tiles : cap);
  int ablate = 0;
  // experiments (profiles/README.md, ablation table): bit 0 skips the tonality pass, 1 the band sums, 2 the MMA loop,
  // 3 its epilogue, 4 phase D, 5 the asynchronous copies of y - the results are then wrong, only the time is of interest
  if (const char* e = std::getenv("AC_PA_ABLATE")) ablate = std::atoi(e);
  auto kernel = pa_mma_tile_kernel<C, QUANT, NFIX, MINB, false, TC>;
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (err != cudaSuccess) return err;
  err = launch_pdl(2, kernel, grid, kThreads, smem, stream, tb, *tb.jobs_host, y, ton_in, omd, thr_scale, thr_out, q_out, frames,
                   tiles, pa_sched_slot(tb), ablate, MdctDeviceTables{}, 0, 0);
  count_launch();
  return err != cudaSuccess ? err : cudaGetLastError();
}

// CTAs per SM: four when the shared memory of the plan allows it (32-filter chunks; the kernel is then compiled for 64
// registers), else three
template <int C, bool QUANT, int NFIX>
cudaError_t launch_mma_tile_n(const PaDeviceTables& tb, const float* y, const float* ton_in, float omd, float thr_scale,
                              float* thr_out, int32_t* q_out, int64_t frames, cudaStream_t stream) {
  auto ctas_per_sm = [](size_t smem) {
    const int v = static_cast<int>((227 * 1024) / (smem + 1024));
    return v > 4 ? 4 : (v < 1 ? 1 : v);
  };
  int per_sm = ctas_per_sm(static_cast<size_t>(layout2(tb).total) * sizeof(float));
  // the spreading product on tcgen05 (tensor memory) when its operand layout costs no resident CTA; AC_PA_MMA=sync keeps
  // the mma.sync product (A/B runs, cross-check tests)
  bool tc = false;
  if (tc_layout_ok(tb)) {
    const int per_sm_tc = ctas_per_sm(static_cast<size_t>(layout2(tb, false, true).total) * sizeof(float));
    const char* e = std::getenv("AC_PA_MMA");
    tc = per_sm_tc >= (per_sm > 3 ? 3 : per_sm) && !(e != nullptr && e[0] == 's');
    if (tc) per_sm = per_sm_tc > 3 ? 3 : per_sm_tc;
  }
  if (const char* e = std::getenv("AC_PA_CTAS")) per_sm = std::max(1, std::min(per_sm, std::atoi(e)));   // experiments
  if (tc)
    return launch_mma_tile_nb<C, QUANT, NFIX, 3, true>(tb, y, ton_in, omd, thr_scale, thr_out, q_out, frames, per_sm, stream);
  if (per_sm >= 4)
    return launch_mma_tile_nb<C, QUANT, NFIX, 4, false>(tb, y, ton_in, omd, thr_scale, thr_out, q_out, frames, per_sm, stream);
  return launch_mma_tile_nb<C, QUANT, NFIX, 3, false>(tb, y, ton_in, omd, thr_scale, thr_out, q_out, frames, per_sm, stream);
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

// Decoder side of the compact side information: one warp per frame row rebuilds thr[row][k][c] from G[row][64][c].
// v = G[b] w0 + G[b + 1] w1 + G[b + 2] w2 in the order of phase D (its two-slot patterns are the same sums with a zero
// weight), the same clamp and the same rsqrt: bit-identical to the thr the encoder quantised with.
template <int C>
__global__ void __launch_bounds__(256) pa_expand_threshold_kernel(const float4* __restrict__ filt4,
                                                                  const float* __restrict__ bark, const float eps_s2,
                                                                  float* __restrict__ thr, const int64_t rows,
                                                                  const int n) {
  __shared__ float s_g[8][kNB * C];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp; row < rows; row += static_cast<int64_t>(gridDim.x) * 8) {
    __syncwarp();
    const float* g = bark + row * (kNB * C);
#pragma unroll
    for (int i = 0; i < 2 * C; ++i) s_g[warp][32 * i + lane] = __ldcs(g + 32 * i + lane);
    __syncwarp();
    float* out = thr + row * static_cast<int64_t>(n) * C;
    for (int k = lane; k < n; k += 32) {
      const float4 f4 = __ldg(filt4 + k);
      const float* gb = &s_g[warp][__float_as_int(f4.w) * C];
      float t[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float v = fmaxf(eps_s2, fmaf(gb[2 * C + c], f4.z, fmaf(gb[C + c], f4.y, gb[c] * f4.x)));
        t[c] = v * rsqrt_approx(v);
      }
      if constexpr (C == 2) {
        __stcs(reinterpret_cast<float2*>(out) + k, make_float2(t[0], t[1]));
      } else if constexpr (C == 4) {
        __stcs(reinterpret_cast<float4*>(out) + k, make_float4(t[0], t[1], t[2], t[3]));
      } else {
#pragma unroll
        for (int c = 0; c < C; ++c) __stcs(out + static_cast<int64_t>(k) * C + c, t[c]);
      }
    }
  }
}

}  // namespace

cudaError_t pa_encode_compact(const PaDeviceTables& tb, const float* y, float drown, float thr_scale, float* bark_out,
                              int32_t* q_out, int64_t rows, int channels, cudaStream_t stream) {
  if (!pa_mma_tile_supported(tb, channels)) return cudaErrorNotSupported;
  if ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(q_out) | reinterpret_cast<uintptr_t>(bark_out)) & 15)
    return cudaErrorMisalignedAddress;
  PaDeviceTables with_out = tb;
  with_out.bark_out = bark_out;
  return pa_threshold_mma_tile(with_out, y, nullptr, static_cast<float>(1.0 - static_cast<double>(drown)), thr_scale, nullptr,
                               q_out, rows, channels, stream);
}

cudaError_t pa_expand_threshold(const PaDeviceTables& tb, const float* bark, float thr_scale, float* thr, int64_t rows,
                                int channels, cudaStream_t stream) {
  if (tb.nb != kNB || tb.filt4 == nullptr || !tb.tile_ok) return cudaErrorNotSupported;
  if ((reinterpret_cast<uintptr_t>(thr) | reinterpret_cast<uintptr_t>(bark)) & 15) return cudaErrorMisalignedAddress;
  const float eps_s2 = tb.eps * (thr_scale * thr_scale);
  const int64_t want = (rows + 7) / 8;
  const int grid = static_cast<int>(want < 16LL * tile_sm_count() ? want : 16LL * tile_sm_count());
  count_launch();
  switch (channels) {
    case 1: pa_expand_threshold_kernel<1><<<grid, 256, 0, stream>>>(tb.filt4, bark, eps_s2, thr, rows, tb.n); break;
    case 2: pa_expand_threshold_kernel<2><<<grid, 256, 0, stream>>>(tb.filt4, bark, eps_s2, thr, rows, tb.n); break;
    case 4: pa_expand_threshold_kernel<4><<<grid, 256, 0, stream>>>(tb.filt4, bark, eps_s2, thr, rows, tb.n); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// ---- the single-pass encoder: x -> (q, step | bark thresholds) without the amplitudes in global memory ------------
bool pa_encode_fused_supported(const PaDeviceTables& tb, const MdctDeviceTables& mt, int channels) {
  return channels == 2 && tb.n == 256 && mt.n == 256 && mt.pre_fwd != nullptr && pa_mma_tile_supported(tb, channels) &&
         static_cast<size_t>(layout2(tb, true, true).total) * sizeof(float) <= 112 * 1024;
}

cudaError_t pa_encode_fused(const PaDeviceTables& tb_in, const MdctDeviceTables& mt, const float* x, float drown,
                            float thr_scale, float* thr_out, float* bark_out, int32_t* q_out, int64_t batches,
                            int64_t blocks_n, int channels, cudaStream_t stream) {
  if (!pa_encode_fused_supported(tb_in, mt, channels)) return cudaErrorNotSupported;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(q_out) | reinterpret_cast<uintptr_t>(thr_out) |
       reinterpret_cast<uintptr_t>(bark_out)) & 15)
    return cudaErrorMisalignedAddress;
  PaDeviceTables tb = tb_in;
  tb.bark_out = bark_out;
  constexpr int C = 2, FT = kTI / C;
  const int64_t frames = blocks_n + 1;
  const int tiles_per_row = static_cast<int>((frames + FT - 1) / FT);
  const int64_t tiles = batches * tiles_per_row;
  // the same spreading product as the chain kernel (tcgen05, or mma.sync under AC_PA_MMA=sync): the two stay bit-identical
  const char* e = std::getenv("AC_PA_MMA");
  const bool tc = !(e != nullptr && e[0] == 's');
  const size_t smem = static_cast<size_t>(layout2(tb, true, tc).total) * sizeof(float);
  int per_sm = static_cast<int>((227 * 1024) / (smem + 1024));
  per_sm = per_sm > 2 ? 2 : (per_sm < 1 ? 1 : per_sm);
  const int64_t cap = static_cast<int64_t>(mma_sm_count()) * per_sm;
  const unsigned grid = static_cast<unsigned>(tiles < cap ? tiles : cap);
  auto kernel = tc ? pa_mma_tile_kernel<C, true, 256, 2, true, true> : pa_mma_tile_kernel<C, true, 256, 2, true, false>;
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (err != cudaSuccess) return err;
  err = launch_pdl(2, kernel, grid, kThreads, smem, stream, tb, *tb.jobs_host, x, nullptr,
                   static_cast<float>(1.0 - static_cast<double>(drown)), thr_scale, thr_out, q_out, batches * frames, tiles,
                   pa_sched_slot(tb), 0, mt, static_cast<int>(blocks_n), tiles_per_row);
  count_launch();
  return err != cudaSuccess ? err : cudaGetLastError();
}

bool pa_mma_tile_supported(const PaDeviceTables& tb, int channels) {
  if (!tb.tile_ok || tb.nb != kNB || tb.jobs_host == nullptr || tb.mma_chunk_k != kChunk) return false;
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
