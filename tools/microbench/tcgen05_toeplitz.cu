// Microbenchmark, follow-up of tcgen05_spread.cu: the two operand forms that let the 64 x 64 x 64 spreading product of the
// masking kernel (acc[item][j] = sum_i P[item][i] S[i][j], psychoacoustic.py:195-206) run on tcgen05.mma kind::tf32 WITHOUT
// 48 KB of extra shared memory:
//   B  S is Toeplitz (S[i][j] = f[64 - i + j]): a K-major / no-swizzle core matrix (8 n-rows x 4 k) of S only depends on
//      c = 2 nb - kb, so the 128 cores of the dense operand are 30 distinct ones.  A matrix descriptor addresses cores
//      affinely (start + nb SBO + kb LBO): with the two K-cores of an instruction taken in descending order (A's bands are
//      stored with the two 4-groups of every 8-group swapped) c grows with both steps, LBO = 128 B, SBO = 256 B, and the
//      operand is a 3840-byte table instead of a 16 KB tile.
//   A  P as the band-sum phase produces it: [band][item], items contiguous = MN-major, which for 32-bit operands exists
//      only as the SWIZZLE_128B_BASE32B layout (atoms of 4 k-rows x 128 B, 32-byte chunks XORed with the row).
//   raw_hi: the `hi` term of A is the unsplit fp32 value (does the tensor core ignore the low 13 mantissa bits?).
// Every combination is checked against a float64 product.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tcgen05_toeplitz tcgen05_toeplitz.cu
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

struct Config {
  int a_mode;       // 0: K-major, no swizzle   1: MN-major, SWIZZLE_128B_BASE32B
  int b_mode;       // 0: dense K-major tile    1: Toeplitz core table (K order swapped inside every k-step)
  int a_lbo, a_sbo; // descriptor fields of A in a_mode 1 (bytes)
  int a_mnb, a_kb;  // a_mode 1: bytes between 32-item blocks / between 4-band atoms in the layout actually written
  int raw_hi;       // 1: A hi = the unsplit fp32 value
  int a_layout;     // descriptor layout type of A in a_mode 1
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// canonical K-major, no swizzle: core = 8 rows (MN) x 16 bytes (4 tf32 along K); the two K-cores of a k-step adjacent
__host__ __device__ inline int tile_offset_words(int mn, int k) {
  const int ks = k >> 3, kc = (k >> 2) & 1, e = k & 3, mc = mn >> 3, r = mn & 7;
  return ks * 512 + mc * 64 + kc * 32 + r * 4 + e;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3fffu);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3fffu) << 32;
  d |= 1ull << 46;                                                // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(layout & 7u) << 61;
  return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128) spread_kernel(const float* __restrict__ p, const float* __restrict__ f,
                                                     float* __restrict__ out, long long* __restrict__ cycles, int reps,
                                                     const Config cfg) {
  extern __shared__ __align__(1024) uint32_t dyn_raw[];
  uint32_t* dyn = dyn_raw + ((1024u - (smem_u32(dyn_raw) & 1023u)) & 1023u) / 4;     // swizzled atoms: 1024-byte aligned
  uint32_t *a_hi = dyn, *a_lo = dyn + 4096, *b_hi = dyn + 8192, *b_lo = dyn + 12288, *z_hi = dyn + 16384, *z_lo = dyn + 16384 + 960;
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < 2 * 4096; i += 128) dyn[i] = 0u;
  __syncthreads();
  for (int i = tid; i < kM * kK; i += 128) {
    const int m = i / kK, k = i % kK;
    const int kp = cfg.b_mode ? (k ^ 4) : k;           // position of band k in the K order of the product
    const float v = p[m * kK + k];
    const uint32_t hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
    int ao;
    if (cfg.a_mode == 0) {
      ao = tile_offset_words(m, kp);
    } else {
      const int byte = (m >> 5) * cfg.a_mnb + (kp >> 2) * cfg.a_kb + (kp & 3) * 128 + ((((m & 31) * 4) ^ ((kp & 3) << 5)));
      ao = byte / 4;
    }
    if (cfg.raw_hi) {
      const uint32_t tr = __float_as_uint(v) & 0xffffe000u;
      a_hi[ao] = __float_as_uint(v);
      a_lo[ao] = __float_as_uint(v - __uint_as_float(tr));
    } else {
      a_hi[ao] = hi;
      a_lo[ao] = __float_as_uint(v - __uint_as_float(hi)) & 0xffffe000u;
    }
    const float w = f[kK - k + m];                     // S[k][n = m]
    const uint32_t whi = (__float_as_uint(w) + 0x1000u) & 0xffffe000u;
    b_hi[tile_offset_words(m, kp)] = whi;
    b_lo[tile_offset_words(m, kp)] = __float_as_uint(w - __uint_as_float(whi)) & 0xffffe000u;
  }
  for (int i = tid; i < 960; i += 128) {               // Toeplitz core table: core c, row r, element e = f[4 + 4c + r - e]
    const int c = i >> 5, r = (i >> 2) & 7, e = i & 3;
    const float w = f[4 + 4 * c + r - e];
    const uint32_t whi = (__float_as_uint(w) + 0x1000u) & 0xffffe000u;
    z_hi[i] = whi;
    z_lo[i] = __float_as_uint(w - __uint_as_float(whi)) & 0xffffe000u;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(cfg.a_mode) << 15) | ((kN >> 3) << 17) |
                         ((kM >> 4) << 24);

  long long t0 = 0, t1 = 0;
  uint32_t parity = 0;
  if (tid == 0) t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    if (tid == 0) {
      const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
      const uint32_t zh = smem_u32(z_hi), zl = smem_u32(z_lo);
#pragma unroll
      for (int ks = 0; ks < kK / 8; ++ks) {
        uint64_t dah, dal, dbh, dbl;
        if (cfg.a_mode == 0) {
          dah = make_desc(ah + ks * 2048, 128, 256, 0);
          dal = make_desc(al + ks * 2048, 128, 256, 0);
        } else {
          dah = make_desc(ah + ks * 2 * cfg.a_kb, cfg.a_lbo, cfg.a_sbo, cfg.a_layout);
          dal = make_desc(al + ks * 2 * cfg.a_kb, cfg.a_lbo, cfg.a_sbo, cfg.a_layout);
        }
        if (cfg.b_mode == 0) {
          dbh = make_desc(bh + ks * 2048, 128, 256, 0);
          dbl = make_desc(bl + ks * 2048, 128, 256, 0);
        } else {
          dbh = make_desc(zh + (14 - 2 * ks) * 128, 128, 256, 0);
          dbl = make_desc(zl + (14 - 2 * ks) * 128, 128, 256, 0);
        }
        mma_tf32(tmem, dal, dbh, idesc, ks > 0 ? 1u : 0u);
        mma_tf32(tmem, dah, dbl, idesc, 1u);
        mma_tf32(tmem, dah, dbh, idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar))
                   : "memory");
    }
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
    // the mma.sync accumulator layout: thread (g = lane / 4, t = lane % 4) gets rows g, g + 8 and columns 2t, 2t + 1 of
    // every block of 8 columns
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
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
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
  std::vector<float> p(kM * kK), out(kM * kN), f32(2 * kK);
  srand(7);
  for (int i = 0; i < kM * kK; ++i) p[i] = std::pow(10.0, -9.0 * rand() / RAND_MAX) * (0.5 + 0.5 * rand() / RAND_MAX);
  for (int t = 0; t < 2 * kK; ++t) {
    const double z = -25.8 + t * (51.6 / (2 * kK - 1)) + 0.474;
    f32[t] = static_cast<float>(std::pow(10.0, 0.6 * (15.81 + 7.5 * z - 17.5 * std::sqrt(1 + z * z)) / 10.0));
  }
  float *dp, *df, *dout;
  long long* dcyc;
  CHECK(cudaMalloc(&dp, p.size() * 4));
  CHECK(cudaMalloc(&df, f32.size() * 4));
  CHECK(cudaMalloc(&dout, out.size() * 4));
  CHECK(cudaMalloc(&dcyc, 16));
  CHECK(cudaMemcpy(dp, p.data(), p.size() * 4, cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(df, f32.data(), f32.size() * 4, cudaMemcpyHostToDevice));
  const int smem = (16384 + 2 * 960) * 4 + 1024;
  CHECK(cudaFuncSetAttribute(spread_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  struct Named { const char* name; Config c; };
  const Named runs[] = {
      {"A K-major, B dense (the known-good form)", {0, 0, 0, 0, 0, 0, 0, 0}},
      {"A K-major, B Toeplitz table", {0, 1, 0, 0, 0, 0, 0, 0}},
      {"A K-major raw hi, B dense", {0, 0, 0, 0, 0, 0, 1, 0}},
      {"A MN SW128_32B (blocks 8192, atoms 512) lbo = blocks, sbo = atoms; B dense", {1, 0, 8192, 512, 8192, 512, 0, 1}},
      {"A MN SW128_32B (blocks 8192, atoms 512) lbo = atoms, sbo = blocks; B dense", {1, 0, 512, 8192, 8192, 512, 0, 1}},
      {"A MN SW128_32B (blocks 512, atoms 1024) lbo = blocks, sbo = atoms; B dense", {1, 0, 512, 1024, 512, 1024, 0, 1}},
      {"A MN SW128_32B (blocks 512, atoms 1024) lbo = atoms, sbo = blocks; B dense", {1, 0, 1024, 512, 512, 1024, 0, 1}},
      {"A MN SW128 layout type 2 (blocks 8192, atoms 512) lbo = blocks, sbo = atoms; B dense", {1, 0, 8192, 512, 8192, 512, 0, 2}},
      {"A MN SW128_32B (blocks 8192, atoms 512) lbo = blocks, sbo = atoms; B Toeplitz", {1, 1, 8192, 512, 8192, 512, 0, 1}},
      {"A MN SW128_32B (blocks 8192, atoms 512) lbo = atoms, sbo = blocks; B Toeplitz", {1, 1, 512, 8192, 8192, 512, 0, 1}},
      {"A MN SW128_32B raw hi (blocks 8192, atoms 512) lbo = blocks, sbo = atoms; B Toeplitz", {1, 1, 8192, 512, 8192, 512, 1, 1}},
      {"A MN SW128_32B raw hi (blocks 8192, atoms 512) lbo = atoms, sbo = blocks; B Toeplitz", {1, 1, 512, 8192, 8192, 512, 1, 1}},
  };
  for (const Named& run : runs) {
    for (int reps : {1, 1000}) {
      CHECK(cudaMemset(dcyc, 0, 16));
      CHECK(cudaMemset(dout, 0, out.size() * 4));
      spread_kernel<<<1, 128, smem>>>(dp, df, dout, dcyc, reps, run.c);
      CHECK(cudaGetLastError());
      CHECK(cudaDeviceSynchronize());
      long long cyc2[2] = {0, 0};
      CHECK(cudaMemcpy(cyc2, dcyc, 16, cudaMemcpyDeviceToHost));
      if (cyc2[1] != 0) std::printf("TIMEOUT waiting for the MMA completion barrier\n");
      CHECK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
      double worst = 0;
      int zeros = 0;
      for (int m = 0; m < kM; ++m)
        for (int n = 0; n < kN; ++n) {
          double ref = 0;
          for (int k = 0; k < kK; ++k) ref += static_cast<double>(p[m * kK + k]) * f32[kK - k + n];
          worst = std::fmax(worst, std::fabs(out[m * kN + n] - ref) / ref);
          zeros += out[m * kN + n] == 0.f;
        }
      if (reps == 1000 || worst > 1e-5)
        std::printf("%-90s reps %4d: %7.1f cycles per product, max relative error %.3g, zeros %d\n", run.name, reps,
                    static_cast<double>(cyc2[0]) / reps, worst, zeros);
    }
  }
  return 0;
}
