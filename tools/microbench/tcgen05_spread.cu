// Microbenchmark: the 64 x 64 x 64 spreading product of the masking kernel (acc[item][j] = sum_i P[item][i] S[i][j],
// psychoacoustic.py:195-206) on the 5th-generation tensor cores: tcgen05.mma kind::tf32, M = 64 items, N = 64 bands,
// K = 8 per instruction, operands in shared memory (canonical K-major, no swizzle), accumulator in tensor memory, issued
// by ONE thread; error-compensated 3xTF32 (P = hi + lo, S = hi + lo: lo hi + hi lo + hi hi = 24 instructions per tile).
// Checks the result against a float64 product on the host and reports cycles per tile.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tcgen05_spread tcgen05_spread.cu ; run on a B200.
// What it is for: DESIGN.md section 4 (why the shipped kernel still uses mma.sync): the instruction works at M = 64 and
// the numbers below are what a port would gain; the cost is 48 KB more shared memory per CTA (S hi / lo as dense 64 x 64
// operand tiles - a Toeplitz table cannot be expressed by a matrix descriptor - and P hi / lo instead of one P).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CHECK(x)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (x);                                                                         \
    if (e_ != cudaSuccess) {                                                                      \
      std::fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));     \
      std::exit(1);                                                                               \
    }                                                                                             \
  } while (0)

constexpr int kM = 64, kN = 64, kK = 64;
#ifndef A_MN_MAJOR
#define A_MN_MAJOR 0        // 1: the A operand as the masking kernel would write it - [band][item], items contiguous
#endif
#ifndef LD_16x256
#define LD_16x256 0         // 1: read the accumulator in the mma.sync fragment layout (tcgen05.ld.16x256b)
#endif
#ifndef A_SBO
#define A_SBO 144
#endif
#ifndef A_SWAP
#define A_SWAP 0
#endif
constexpr int kASbo = A_SBO;            // MN-major A: bytes between cores of 4 items x 8 bands (128 B of data + 16 B pad:
constexpr int kAKs = 16 * kASbo;      //   the lanes' 8-byte stores then fall on different banks), bytes per k-step

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// canonical K-major, no swizzle: core matrix = 8 rows (MN) x 16 bytes (4 tf32 along K), 128 contiguous bytes;
// the two K-cores of one K = 8 instruction are adjacent (LBO = 128 B), MN-cores follow at SBO = 256 B, k-steps at 2 KB
__host__ __device__ inline int tile_offset_words(int mn, int k) {
  const int ks = k >> 3, kc = (k >> 2) & 1, e = k & 3, mc = mn >> 3, r = mn & 7;
  return ks * 512 + mc * 64 + kc * 32 + r * 4 + e;
}

// MN-major, no swizzle: core = 8 K-rows of 16 bytes (4 items); cores along M at kASbo
__host__ __device__ inline int a_mn_offset_words(int m, int k) {
  return ((k >> 3) * kAKs + (m >> 2) * kASbo + (k & 7) * 16 + (m & 3) * 4) / 4;
}

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3fffu);
  const uint32_t mn_stride = static_cast<uint32_t>(kASbo) >> 4, k_stride = static_cast<uint32_t>(kAKs) >> 4;
  d |= static_cast<uint64_t>((A_SWAP ? mn_stride : k_stride) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((A_SWAP ? k_stride : mn_stride) & 0x3fffu) << 32;
  d |= 1ull << 46;
  return d;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3fffu);             // start address
  d |= static_cast<uint64_t>((128u >> 4) & 0x3fffu) << 16;        // leading byte offset: the next core along K
  d |= static_cast<uint64_t>((256u >> 4) & 0x3fffu) << 32;        // stride byte offset: the next core along M / N
  d |= 1ull << 46;                                                // descriptor version (Blackwell)
  return d;                                                       // layout type 0: no swizzle
}

// kind::tf32, fp32 accumulate, both operands K-major, M = 64, N = 64
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(A_MN_MAJOR) << 15) |
                            ((kN >> 3) << 17) | ((kM >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128) spread_tcgen05_kernel(const float* __restrict__ p, const float* __restrict__ s,
                                                             float* __restrict__ out, long long* __restrict__ cycles,
                                                             int reps) {
  extern __shared__ __align__(128) uint32_t dyn[];     // 4 operand tiles (A tiles padded in the MN-major variant)
  constexpr int kATile = A_MN_MAJOR ? 8 * kAKs / 4 : kM * kK;
  uint32_t *a_hi = dyn, *a_lo = dyn + kATile, *b_hi = dyn + 2 * kATile, *b_lo = dyn + 2 * kATile + kN * kK;
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // operands: A[m][k] = P[item m][band k], B[n][k] = S[k][n] (the product needs B^T in K-major form), split hi / lo
  for (int i = tid; i < kM * kK; i += 128) {
    const int m = i / kK, k = i % kK;
    const float v = p[m * kK + k];
    const uint32_t hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
    const int ao = A_MN_MAJOR ? a_mn_offset_words(m, k) : tile_offset_words(m, k);
    a_hi[ao] = hi;
    a_lo[ao] = __float_as_uint(v - __uint_as_float(hi)) & 0xffffe000u;
    const float w = s[k * kN + m];                 // S[k][n = m]
    const uint32_t whi = (__float_as_uint(w) + 0x1000u) & 0xffffe000u;
    b_hi[tile_offset_words(m, k)] = whi;
    b_lo[tile_offset_words(m, k)] = __float_as_uint(w - __uint_as_float(whi)) & 0xffffe000u;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes of the operands -> async proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;

  long long t0 = 0, t1 = 0;
  uint32_t parity = 0;
  if (tid == 0) t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    if (tid == 0) {
      const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
      for (int ks = 0; ks < kK / 8; ++ks) {
        const uint32_t o = ks * 2048, oa = A_MN_MAJOR ? ks * kAKs : o;
        const uint64_t dal = A_MN_MAJOR ? make_desc_mn(al + oa) : make_desc(al + oa);
        const uint64_t dah = A_MN_MAJOR ? make_desc_mn(ah + oa) : make_desc(ah + oa);
        mma_tf32(tmem, dal, make_desc(bh + o), ks > 0 ? 1u : 0u);
        mma_tf32(tmem, dah, make_desc(bl + o), 1u);
        mma_tf32(tmem, dah, make_desc(bh + o), 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar))
                   : "memory");
    }
    // everybody waits for the accumulator
    uint32_t done = 0, spins = 0;
    while (!done && ++spins < (1u << 24)) {      // bounded: a wrong descriptor must not hang the GPU
      asm volatile(
          "{\n"
          ".reg .pred q;\n"
          "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n"
          "selp.u32 %0, 1, 0, q;\n"
          "}\n"
          : "=r"(done)
          : "r"(smem_u32(&mbar)), "r"(parity)
          : "memory");
    }
    if (!done) {
      if (tid == 0) cycles[1] = -1;
      break;
    }
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#if LD_16x256
    // the mma.sync accumulator layout: thread (g = lane / 4, t = lane % 4) gets rows g, g + 8 and columns 2t, 2t + 1 of
    // every block of 8 columns: v[4 i + 0 .. 3] = (g, 8i + 2t), (g, 8i + 2t + 1), (g + 8, 8i + 2t), (g + 8, 8i + 2t + 1)
    uint32_t v[32];
    const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
        "%25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
          "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
          "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (rep == reps - 1) {
      const int g = lane >> 2, t = lane & 3;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        out[(warp * 16 + g) * kN + 8 * i + 2 * t] = __uint_as_float(v[4 * i]);
        out[(warp * 16 + g) * kN + 8 * i + 2 * t + 1] = __uint_as_float(v[4 * i + 1]);
        out[(warp * 16 + g + 8) * kN + 8 * i + 2 * t] = __uint_as_float(v[4 * i + 2]);
        out[(warp * 16 + g + 8) * kN + 8 * i + 2 * t + 1] = __uint_as_float(v[4 * i + 3]);
      }
    }
#else
    // M = 64: row m sits in lane (m % 16) + 32 (m / 16), i.e. lanes 0 .. 15 of warp m / 16; 64 fp32 columns per row
    uint32_t v[64];
    const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
        "%25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
          "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
          "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]),
          "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]),
          "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]),
          "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]),
          "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (rep == reps - 1 && lane < 16) {
      const int m = warp * 16 + lane;
#pragma unroll
      for (int n = 0; n < kN; ++n) out[m * kN + n] = __uint_as_float(v[n]);
    }
#endif
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                      // the accumulator has been read: the next product may overwrite it
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (tid == 0) {
    t1 = clock64();
    cycles[0] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main() {
  std::vector<float> p(kM * kK), s(kK * kN), out(kM * kN);
  // P: positive, nine decades of dynamic range (as max(eps, I_bark)^alpha is); S: the Toeplitz spreading prototype's shape
  srand(7);
  for (int i = 0; i < kM * kK; ++i) p[i] = std::pow(10.0, -9.0 * rand() / RAND_MAX) * (0.5 + 0.5 * rand() / RAND_MAX);
  std::vector<double> f(2 * kK);
  for (int t = 0; t < 2 * kK; ++t) {
    const double z = -25.8 + t * (51.6 / (2 * kK - 1)) + 0.474;
    f[t] = std::pow(10.0, 0.6 * (15.81 + 7.5 * z - 17.5 * std::sqrt(1 + z * z)) / 10.0);
  }
  for (int i = 0; i < kK; ++i)
    for (int j = 0; j < kN; ++j) s[i * kN + j] = static_cast<float>(f[kK - i + j]);
  float *dp, *ds, *dout;
  long long* dcyc;
  CHECK(cudaMalloc(&dp, p.size() * 4));
  CHECK(cudaMalloc(&ds, s.size() * 4));
  CHECK(cudaMalloc(&dout, out.size() * 4));
  CHECK(cudaMalloc(&dcyc, 16));
  CHECK(cudaMemset(dcyc, 0, 16));
  CHECK(cudaMemcpy(dp, p.data(), p.size() * 4, cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(ds, s.data(), s.size() * 4, cudaMemcpyHostToDevice));
  CHECK(cudaMemset(dout, 0, out.size() * 4));
  const int smem = (A_MN_MAJOR ? 2 * 8 * kAKs : 2 * kM * kK * 4) + 2 * kN * kK * 4;
  CHECK(cudaFuncSetAttribute(spread_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int reps : {1, 1000}) {
    spread_tcgen05_kernel<<<1, 128, smem>>>(dp, ds, dout, dcyc, reps);
    CHECK(cudaGetLastError());
    CHECK(cudaDeviceSynchronize());
    long long cyc2[2] = {0, 0};
    CHECK(cudaMemcpy(cyc2, dcyc, 16, cudaMemcpyDeviceToHost));
    const long long cyc = cyc2[0];
    if (cyc2[1] != 0) std::printf("TIMEOUT waiting for the MMA completion barrier\n");
    CHECK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0, worst_fp32 = 0;
    for (int m = 0; m < kM; ++m)
      for (int n = 0; n < kN; ++n) {
        double ref = 0;
        float ref32 = 0.f;
        for (int k = 0; k < kK; ++k) {
          ref += static_cast<double>(p[m * kK + k]) * s[k * kN + n];
          ref32 = std::fmaf(p[m * kK + k], s[k * kN + n], ref32);
        }
        worst = std::fmax(worst, std::fabs(out[m * kN + n] - ref) / ref);
        worst_fp32 = std::fmax(worst_fp32, std::fabs(ref32 - ref) / ref);
      }
    std::printf("A %s, accumulator read %s; reps %d: %.1f cycles per 64 x 64 x 64 3xTF32 product (24 tcgen05.mma + commit + wait + tcgen05.ld of 64 columns); "
                "max relative error vs float64 %.3g (a sequential fp32 FMA sum: %.3g)\n",
                A_MN_MAJOR ? "MN-major (items contiguous, 144 B cores)" : "K-major", LD_16x256 ? "16x256b.x8" : "32x32b.x64", reps,
                static_cast<double>(cyc) / reps, worst, worst_fp32);
  }
  return 0;
}
