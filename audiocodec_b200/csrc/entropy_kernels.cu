// Entropy-coded bitstream of the quantised integers (sm_100a): adaptive Golomb-Rice, one independent byte range per
// frame row.
//
// No reference symbol: korneelvdbroek/audiocodec has no quantiser and no bitstream (SURVEY.md finding 2, 8f row 4);
// the format is build-defined; the tests hold a CPU restatement that reproduces the stream bit for bit.
//
// Format.  A "row" is row_len consecutive integers of q (one frame: filters x channels, interleaved as stored), cut
// into groups of 16.  Per group: a 5-bit header k (31 = the group is all zero, no payload) followed by the 16 values,
// each zigzag-mapped u = (q << 1) ^ (q >> 31) and written as (u >> k) zero bits, a one bit, then the k low bits of u
// (least significant bit first).  k is the candidate of {k0 - 1, k0, k0 + 1}, k0 = floor(log2(mean u + 1)), with the
// fewest payload bits (lowest k on ties).  Bits fill little-endian 32-bit words from bit 0 upwards; a row's bit string
// is padded with zeros to a multiple of 32 bits, so every row starts on a 4-byte boundary of the stream and
// offsets[row] (bytes) addresses it directly - rows decode independently, which is what the GPU decoder uses.
//
// Three kernels, one thread per row (the hot path is elsewhere; 110 k rows on cfg2): sizes, an exclusive scan of the
// sizes (single CTA), and the writer; the decoder is a fourth.
#include "kernels.h"

namespace ac {

namespace {

constexpr int kGroup = 16;

__device__ __forceinline__ uint32_t zigzag(int32_t q) {
  return (static_cast<uint32_t>(q) << 1) ^ static_cast<uint32_t>(q >> 31);
}

__device__ __forceinline__ int32_t unzigzag(uint32_t u) {
  return static_cast<int32_t>(u >> 1) ^ -static_cast<int32_t>(u & 1u);
}

// payload bits of a group's 16 values with Rice parameter k
__device__ __forceinline__ uint64_t rice_bits(const uint32_t (&u)[kGroup], int k) {
  uint64_t bits = 0;
#pragma unroll
  for (int i = 0; i < kGroup; ++i) bits += (u[i] >> k) + 1u + static_cast<uint32_t>(k);
  return bits;
}

// loads a group, returns its Rice parameter (31: all zero) and payload bits
__device__ __forceinline__ int choose_k(const int32_t* __restrict__ g, uint32_t (&u)[kGroup], uint64_t& payload) {
  uint64_t sum = 0;
#pragma unroll
  for (int i = 0; i < kGroup; ++i) {
    u[i] = zigzag(g[i]);
    sum += u[i];
  }
  if (sum == 0) {
    payload = 0;
    return 31;
  }
  const uint64_t mean1 = sum / kGroup + 1;
  const int k0 = 63 - __clzll(static_cast<long long>(mean1));          // floor(log2(mean + 1)), 0 .. 31
  int best = -1;
  uint64_t best_bits = 0;
  for (int k = max(k0 - 1, 0); k <= min(k0 + 1, 30); ++k) {
    const uint64_t b = rice_bits(u, k);
    if (best < 0 || b < best_bits) {
      best = k;
      best_bits = b;
    }
  }
  payload = best_bits;
  return best;
}

__global__ void __launch_bounds__(128) rice_sizes_kernel(const int32_t* __restrict__ q, int64_t rows, int row_len,
                                                        int64_t* __restrict__ sizes) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const int32_t* r = q + row * row_len;
  uint64_t bits = 0;
  for (int g = 0; g < row_len / kGroup; ++g) {
    uint32_t u[kGroup];
    uint64_t payload;
    choose_k(r + g * kGroup, u, payload);
    bits += 5 + payload;
  }
  sizes[row] = static_cast<int64_t>((bits + 31) / 32) * 4;              // bytes, a multiple of four
}

// exclusive scan of sizes[0 .. rows) into offsets[0 .. rows] (single CTA); offsets may be the same array as sizes: a
// thread reads sizes[i] before it writes offsets[i], and its chunk is its own
__global__ void __launch_bounds__(1024) scan_sizes_kernel(const int64_t* sizes, int64_t rows, int64_t* offsets) {
  __shared__ int64_t s_part[1024];
  const int tid = threadIdx.x;
  const int64_t per = (rows + 1023) / 1024;
  const int64_t lo = min(rows, tid * per), hi = min(rows, lo + per);
  int64_t sum = 0;
  for (int64_t i = lo; i < hi; ++i) sum += sizes[i];
  s_part[tid] = sum;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {                                   // Hillis-Steele inclusive scan of the partials
    const int64_t v = tid >= d ? s_part[tid - d] : 0;
    __syncthreads();
    s_part[tid] += v;
    __syncthreads();
  }
  int64_t run = tid > 0 ? s_part[tid - 1] : 0;
  for (int64_t i = lo; i < hi; ++i) {
    const int64_t sz = sizes[i];
    offsets[i] = run;
    run += sz;
  }
  if (tid == 1023) offsets[rows] = s_part[1023];
}

struct BitWriter {
  uint32_t* out;
  uint64_t acc = 0;
  int fill = 0;
  __device__ __forceinline__ void put(uint32_t value, int nbits) {       // nbits <= 32
    acc |= static_cast<uint64_t>(value) << fill;
    fill += nbits;
    if (fill >= 32) {
      *out++ = static_cast<uint32_t>(acc);
      acc >>= 32;
      fill -= 32;
    }
  }
  __device__ __forceinline__ void zeros(uint32_t n) {                     // a run of zero bits of any length
    while (n >= 32) {
      put(0u, 32);
      n -= 32;
    }
    if (n) put(0u, static_cast<int>(n));
  }
  __device__ __forceinline__ void flush() {
    if (fill > 0) *out++ = static_cast<uint32_t>(acc);
  }
};

__global__ void __launch_bounds__(128) rice_write_kernel(const int32_t* __restrict__ q, int64_t rows, int row_len,
                                                        const int64_t* __restrict__ offsets, uint8_t* __restrict__ bytes) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const int32_t* r = q + row * row_len;
  BitWriter w;
  w.out = reinterpret_cast<uint32_t*>(bytes + offsets[row]);
  for (int g = 0; g < row_len / kGroup; ++g) {
    uint32_t u[kGroup];
    uint64_t payload;
    const int k = choose_k(r + g * kGroup, u, payload);
    w.put(static_cast<uint32_t>(k), 5);
    if (k == 31) continue;
    for (int i = 0; i < kGroup; ++i) {
      w.zeros(u[i] >> k);
      w.put(1u, 1);
      if (k) w.put(u[i] & ((1u << k) - 1u), k);
    }
  }
  w.flush();
}

struct BitReader {
  const uint32_t* in;
  uint64_t acc = 0;
  int fill = 0;
  __device__ __forceinline__ void refill() {
    if (fill <= 32) {
      acc |= static_cast<uint64_t>(*in++) << fill;
      fill += 32;
    }
  }
  __device__ __forceinline__ uint32_t get(int nbits) {                    // nbits <= 32
    refill();
    const uint32_t v = static_cast<uint32_t>(acc & ((nbits == 32) ? 0xffffffffull : ((1ull << nbits) - 1ull)));
    acc >>= nbits;
    fill -= nbits;
    return v;
  }
  __device__ __forceinline__ uint32_t unary() {                           // zeros up to and including the next one bit
    uint32_t n = 0;
    for (;;) {
      refill();
      const uint32_t low = static_cast<uint32_t>(acc);
      const int avail = fill < 32 ? fill : 32;
      const int z = low ? __ffs(static_cast<int>(low)) - 1 : 32;
      if (z < avail) {
        acc >>= (z + 1);
        fill -= z + 1;
        return n + static_cast<uint32_t>(z);
      }
      acc >>= avail;
      fill -= avail;
      n += static_cast<uint32_t>(avail);
    }
  }
};

// The word behind a row's last one is read ahead but never used: the stream buffer carries 4 bytes of slack.
__global__ void __launch_bounds__(128) rice_read_kernel(const uint8_t* __restrict__ bytes, const int64_t* __restrict__ offsets,
                                                       int64_t rows, int row_len, int32_t* __restrict__ q) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  int32_t* r = q + row * row_len;
  BitReader rd;
  rd.in = reinterpret_cast<const uint32_t*>(bytes + offsets[row]);
  for (int g = 0; g < row_len / kGroup; ++g) {
    const int k = static_cast<int>(rd.get(5));
    if (k == 31) {
#pragma unroll
      for (int i = 0; i < kGroup; ++i) r[g * kGroup + i] = 0;
      continue;
    }
    for (int i = 0; i < kGroup; ++i) {
      uint32_t u = rd.unary() << k;
      if (k) u |= rd.get(k);
      r[g * kGroup + i] = unzigzag(u);
    }
  }
}

}  // namespace

cudaError_t entropy_plan(const int32_t* q, int64_t rows, int row_len, int64_t* offsets, cudaStream_t stream) {
  // the row sizes are written to offsets[0 .. rows) and scanned in place
  if (rows > 0) {
    rice_sizes_kernel<<<static_cast<unsigned>((rows + 127) / 128), 128, 0, stream>>>(q, rows, row_len, offsets);
    count_launch();
  }
  scan_sizes_kernel<<<1, 1024, 0, stream>>>(offsets, rows, offsets);
  count_launch();
  return cudaGetLastError();
}

cudaError_t entropy_encode(const int32_t* q, int64_t rows, int row_len, const int64_t* offsets, uint8_t* bytes,
                           cudaStream_t stream) {
  if (rows == 0) return cudaSuccess;
  rice_write_kernel<<<static_cast<unsigned>((rows + 127) / 128), 128, 0, stream>>>(q, rows, row_len, offsets, bytes);
  count_launch();
  return cudaGetLastError();
}

cudaError_t entropy_decode(const uint8_t* bytes, const int64_t* offsets, int64_t rows, int row_len, int32_t* q,
                           cudaStream_t stream) {
  if (rows == 0) return cudaSuccess;
  rice_read_kernel<<<static_cast<unsigned>((rows + 127) / 128), 128, 0, stream>>>(bytes, offsets, rows, row_len, q);
  count_launch();
  return cudaGetLastError();
}

}  // namespace ac
