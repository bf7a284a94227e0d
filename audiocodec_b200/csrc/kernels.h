// Internal C++ interface between the C ABI (capi.cu) and the kernel translation units.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace ac {

// Device-resident MDCT tables (all fp32, built in float64 on the host; tables.h documents the entries).
struct MdctDeviceTables {
  int n = 0;
  float scale_fwd = 0.f;             // 1 / (N sqrt 2)  = (1 / sqrt(4N)) * sqrt(2 / N)   (mdctransformer.py:125, :314)
  float scale_inv = 0.f;             // 2 sqrt 2        = sqrt(4N) * sqrt(2 / N)         (mdctransformer.py:145, :314)
  const float4* fold = nullptr;      // [N/2]
  const float4* unfold = nullptr;    // [N/2]
  const float2* tw_pre = nullptr;    // [N/2]  exp(-i pi (j + 1/8) / N)
  const float2* tw_post_fwd = nullptr;   // tw_pre * scale_fwd
  const float2* tw_post_inv = nullptr;   // tw_pre * scale_inv
  const float2* roots = nullptr;     // [N/2]  exp(-2 pi i j / (N/2))
  const float* cos_table = nullptr;  // [8N]   cos(pi m / (4N)), generic-N path only
  // tile kernels (mdct_tile_kernels.cu), [variant][n]: host-merged coefficients, see capi.cu
  const float4* pre_fwd = nullptr;   // [2][2][N/2]  fold x pre-twiddle: Re / Im as dot products of the four samples
  const float4* post_fwd = nullptr;  // [2][N/2]     post-twiddle x 1 / (N sqrt 2), as the two stored outputs
  const float4* pre_inv = nullptr;   // [2][N/2]     pre-twiddle of the inverse
  const float4* post_inv = nullptr;  // [2][N/2]     post-twiddle x 2 sqrt 2
  // inter-pass twiddles of the tile FFT, one contiguous row per butterfly input: tw_pass1[(r - 1) * (M / R1) + j] =
  // exp(-2 pi i r (j mod R0) / (R0 R1)), tw_pass2 likewise for the third pass (radices from tile_fft_radices)
  const float2* tw_pass1 = nullptr;
  const float2* tw_pass2 = nullptr;
};

// radices of the tile kernels' FFT plan for filters_n = n (mdct_tile_kernels.cu); false when n has no tile plan
inline bool tile_fft_radices(int n, int* r0, int* r1, int* r2) {
  switch (n) {
    case 64: *r0 = 8; *r1 = 4; *r2 = 1; return true;
    case 128: *r0 = 8; *r1 = 8; *r2 = 1; return true;
    case 256: *r0 = 16; *r1 = 8; *r2 = 1; return true;
    case 512: *r0 = 16; *r1 = 16; *r2 = 1; return true;
    case 1024: *r0 = 8; *r1 = 8; *r2 = 8; return true;
    default: return false;
  }
}

// Band-sum job list of the tensor-core tile kernel, passed BY VALUE as a kernel parameter: it then lives in the
// constant bank, job descriptors and loop bounds are read through the uniform datapath and every branch on them is
// warp-uniform by construction (no convergence barriers, no shared-memory traffic for control).
constexpr int kPaMaxJobs = 112, kPaMaxChunks = 32;
struct PaJobParams {
  int4 job[kPaMaxJobs];
  int32_t start[9 * kPaMaxChunks + 1];
  // tonality pass: warp w of chunk c sums the filters ton_start[9 c + w] .. ton_start[9 c + w + 1] of the chunk
  int16_t ton_start[9 * kPaMaxChunks + 2];
};

// Device-resident psychoacoustic tables (psychoacoustic.py:52-69, sparse forms from tables.h).
struct PaDeviceTables {
  int n = 0, nb = 0;
  float alpha = 0.f;        // fp32(alpha), exponent of the non-linear superposition (psychoacoustic.py:206)
  float neg_alpha = 0.f;    // fp32(-alpha)                                           (psychoacoustic.py:197)
  float inv_alpha = 0.f;    // fp32(1 / alpha)                                        (psychoacoustic.py:208)
  float eps = 1e-14f;       // _INTENSITY_EPS                                         (psychoacoustic.py:56)
  int max_band_cnt = 0, max_filt_cnt = 0;
  const int32_t* band_k0 = nullptr;
  const int32_t* band_cnt = nullptr;
  const int32_t* band_ptr = nullptr;
  const float* band_w = nullptr;
  const int32_t* filt_b0 = nullptr;
  const int32_t* filt_cnt = nullptr;
  const int32_t* filt_ptr = nullptr;
  const float* filt_w = nullptr;
  const float* quiet = nullptr;       // [nb]
  const float* spread_fn = nullptr;   // [2 nb]
  const float* lin = nullptr;         // [nb]
  // tile kernel (nb == 64, <= 3 bands per filter): the filter axis is processed in chunks of chunk_k filters.
  //   band_desc[d] = one step of four filters of a band sum: { row of T, offset of its four (zero-padded)
  //                  weights in band_w4, band | add-earlier-partial << 8 | band-complete << 9 | first-step << 10 |
  //                  last-step-in-chunk << 11, 0 }
  //   desc_start[5 c + w] .. desc_start[5 c + w + 1]: the steps warp w of a 4-warp CTA runs in chunk c
  //   filt4[k] = { w0, w1, w2, first band (as int bits) }: W_inv weights of the three bands from `first band`
  int chunk_k = 0, n_chunks = 0, tile_ok = 0;
  int n_desc = 0, n_band_w4 = 0;
  const int4* band_desc = nullptr;
  const float* band_w4 = nullptr;
  const int32_t* desc_start = nullptr;
  const float4* filt4 = nullptr;
  float gain_log2 = 0.f;    // fp32(-alpha * log2(10) / 10): gain = 2^(gain_log2 * offset)   (psychoacoustic.py:197)
  // tensor-core tile kernel (psycho_mma_kernels.cu): the filter axis is processed in chunks of mma_chunk_k filters
  //   jobs_host->job[j] = one bark band inside one chunk: { byte offset of its first row of T (66 words per row), byte
  //                 offset of its (zero-padded, duplicated) weights, steps of four filters, byte offset of P[band] (64
  //                 items per band, XOR swizzle (band & 3) << 3 items folded in) | add-earlier-partial << 16 |
  //                 band-complete << 17 }
  //   jobs_host->start[9 c + w] .. start[9 c + w + 1]: the jobs warp w of an 8-warp CTA runs in chunk c (9 entries per chunk)
  //   pow_alpha[E] / pow_inv_alpha[E], E = biased exponent of x: { 2^rint(a (E - 127)), a (E - 127) - rint(a (E - 127)) },
  //                 so that x^a = 2^(a log2(mantissa) + second) * first with full fp32 precision in the exponent
  int n_jobs = 0, mma_chunk_k = 0, mma_n_chunks = 0, n_mma_w4 = 0;
  const float* mma_w4 = nullptr;   // the (zero-padded) weights of the jobs' steps
  // three bits per group of 32 filters: bit s set when a filter of the group has a non-zero filt4 weight for band slot s
  uint8_t filt_mask[64] = {};
  int clamp_needed = 1;     // 0: max(eps, .) before ^(1/alpha) can never win against the quiet threshold
  // 0: max(eps, .) before the square root of the filter-domain threshold (psychoacoustic.py:330-331) can never act - every
  // filter's share of the quiet threshold alone is above 4 eps (the bark-domain threshold is >= the quiet threshold)
  int thr_clamp_needed = 1;
  const PaJobParams* jobs_host = nullptr;   // HOST pointer (owned by the plan); null when the list does not fit
  const float2* pow_alpha = nullptr;
  const float2* pow_inv_alpha = nullptr;
  float offset_log2 = 0.f;  // fp32(-log2(10) / 10) = gain_log2 / alpha: the masking offset in the log2 domain after ^(1/alpha)
  // compact side information: when set, the tensor-core tile kernel also writes the bark-domain
  // thresholds G [rows][64][channels] (intensity, thr_scale^2 folded in) from which pa_expand_threshold rebuilds thr
  float* bark_out = nullptr;
  // split powers of the tensor-core tile kernel, for alpha near 1/2 (the default 0.6): x^alpha = sqrt(x) x^(alpha - 1/2)
  // and x^(1/alpha) = x^2 x^(1/alpha - 2) - the exponent that multiplies lg2(x) is small (pow_c1 = alpha - 1/2,
  // pow_c2 = 1/alpha - 2), so the rounding of lg2(x) (|lg2 x| up to 47) no longer limits the result and the exponent
  // tables are not needed; pow_split == 0 keeps the table powers (any alpha)
  int pow_split = 0;
  float pow_c1 = 0.f, pow_c2 = 0.f;
  // tile scheduler of the tensor-core tile kernel: kPaSchedSlots pairs of {next ticket, finished CTAs}, zero between launches
  unsigned* sched = nullptr;
};
constexpr int kPaSchedSlots = 64;

// psycho_mma_kernels.cu: nb == 64, <= 3 bands per filter, 1 / 2 / 4 channels
bool pa_mma_tile_supported(const PaDeviceTables& tb, int channels);
cudaError_t pa_threshold_mma_tile(const PaDeviceTables& tb, const float* y, const float* ton_in, float one_minus_drown,
                                  float thr_scale, float* thr_out, int32_t* q_out, int64_t frames, int channels,
                                  cudaStream_t stream);

// psycho_mma_kernels.cu, single-pass encoder: forward MDCT + masking + quantiser in one kernel,
// x [B, S, 2] -> q (and the steps thr_out and / or the bark-domain thresholds bark_out, either may be null); the
// amplitudes never reach global memory.  Stereo, filters_n = 256; cudaErrorNotSupported otherwise.
bool pa_encode_fused_supported(const PaDeviceTables& tb, const MdctDeviceTables& mt, int channels);
cudaError_t pa_encode_fused(const PaDeviceTables& tb, const MdctDeviceTables& mt, const float* x, float drown,
                            float thr_scale, float* thr_out, float* bark_out, int32_t* q_out, int64_t batches,
                            int64_t blocks_n, int channels, cudaStream_t stream);

// elementwise_kernels.cu: bfloat16 <-> float32 (round to nearest even) and the dB utilities (psychoacoustic.py:71-100)
cudaError_t bf16_to_f32(const void* in, float* out, int64_t n, cudaStream_t stream);
cudaError_t f32_to_bf16(const float* in, void* out, int64_t n, cudaStream_t stream);
cudaError_t amplitude_to_db_f32(const float* a, float* out, int64_t n, float eps, float db_max, float db_min, bool norm,
                                cudaStream_t stream);
cudaError_t amplitude_to_db_f64(const double* a, double* out, int64_t n, double eps, double db_max, double db_min, bool norm,
                                cudaStream_t stream);

// backward_kernels.cu: vector-Jacobian products of tonality and global_masking_threshold (warp per item, functional)
cudaError_t pa_tonality_backward(const PaDeviceTables& tb, const float* y, const float* grad_ton, float* grad_y,
                                 int64_t rows, int channels, cudaStream_t stream);
cudaError_t pa_threshold_backward(const PaDeviceTables& tb, const float* y, const float* ton, float drown,
                                  const float* grad_thr, float* grad_y, float* grad_ton, int64_t rows, int channels,
                                  cudaStream_t stream);

// entropy_kernels.cu: adaptive Golomb-Rice bitstream of the quantised integers, one independent byte range per row of
// row_len integers (row_len a multiple of 16).  entropy_plan: offsets[0 .. rows] = byte offset of every row, offsets[rows]
// = size of the stream; entropy_encode writes it; entropy_decode reads it back (the stream buffer needs 4 bytes of slack).
cudaError_t entropy_plan(const int32_t* q, int64_t rows, int row_len, int64_t* offsets, cudaStream_t stream);
cudaError_t entropy_encode(const int32_t* q, int64_t rows, int row_len, const int64_t* offsets, uint8_t* bytes,
                           cudaStream_t stream);
cudaError_t entropy_decode(const uint8_t* bytes, const int64_t* offsets, int64_t rows, int row_len, int32_t* q,
                           cudaStream_t stream);

// float64 compute dtype (f64_kernels.cu): the same tables in double, sparse forms shared with the fp32 plan
struct MdctDeviceTables64 {
  int n = 0;
  double scale_fwd = 0., scale_inv = 0.;
  const double* fold = nullptr;        // [4 N/2]
  const double* unfold = nullptr;      // [4 N/2]
  const double* cos_table = nullptr;   // [8N] cos(pi m / (4N))
};
struct PaDeviceTables64 {
  int n = 0, nb = 0;
  double alpha = 0., inv_alpha = 0., eps = 1e-14;
  const int32_t* band_k0 = nullptr;
  const int32_t* band_cnt = nullptr;
  const int32_t* band_ptr = nullptr;
  const double* band_w = nullptr;
  const int32_t* filt_b0 = nullptr;
  const int32_t* filt_cnt = nullptr;
  const int32_t* filt_ptr = nullptr;
  const double* filt_w = nullptr;
  const double* quiet = nullptr;
  const double* spread_fn = nullptr;
  const double* lin = nullptr;
};
cudaError_t mdct_forward_f64(const MdctDeviceTables64& tb, const double* x, double* y, int64_t batches, int64_t blocks_n,
                             int channels, cudaStream_t stream);
cudaError_t mdct_inverse_f64(const MdctDeviceTables64& tb, const double* y, double* x, int64_t batches, int64_t frames_n,
                             int channels, cudaStream_t stream);
cudaError_t pa_tonality_f64(const PaDeviceTables64& tb, const double* y, double* ton, int64_t rows, int channels,
                            cudaStream_t stream);
cudaError_t pa_threshold_f64(const PaDeviceTables64& tb, const double* y, const double* ton_in, double drown, double* thr,
                             int64_t rows, int channels, cudaStream_t stream);
cudaError_t quantize_f64(const double* y, const double* thr, int32_t* q, int64_t n, cudaStream_t stream);
cudaError_t dequantize_f64(const int32_t* q, const double* thr, double* y, int64_t n, cudaStream_t stream);

void count_launch();
int tile_sm_count();

// Programmatic dependent launch (the tile kernels of the encode / decode chain): the launch carries the programmatic
// stream-serialisation attribute, the kernel stages its tables, allocates tensor memory and initialises its barriers,
// then executes griddepcontrol.wait before it touches a tensor - so its launch latency and prologue overlap the tail of
// the kernel before it on the stream.  AC_PDL=0 launches without the attribute (A/B runs).
// kind: 1 forward MDCT, 2 masking / single-pass encoder, 4 inverse MDCT (AC_PDL is a mask of the kinds that use it)
bool pdl_enabled(int kind);
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(int kind, void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream,
                       Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled(kind) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

bool mdct_tile_forward_supported(int n, int channels);
bool mdct_tile_inverse_supported(int n, int channels);
cudaError_t mdct_forward_tile(const MdctDeviceTables& tb, const float* x, float* y, int64_t batches, int64_t blocks_n,
                              int channels, cudaStream_t stream);
cudaError_t mdct_inverse_tile(const MdctDeviceTables& tb, const float* y, const int32_t* q, const float* thr, float* x,
                              int64_t batches, int64_t frames_n, int channels, cudaStream_t stream);

bool mdct_has_fast_path(int n);
cudaError_t mdct_forward(const MdctDeviceTables& tb, const float* x, float* y, int64_t batches, int64_t blocks_n,
                         int channels, cudaStream_t stream);
// q == nullptr: plain inverse of y.  q != nullptr: inverse of q * thr (y ignored).
cudaError_t mdct_inverse(const MdctDeviceTables& tb, const float* y, const int32_t* q, const float* thr, float* x,
                         int64_t batches, int64_t frames_n, int channels, cudaStream_t stream);

cudaError_t pa_tonality(const PaDeviceTables& tb, const float* y, float* ton, int64_t rows, int channels,
                        cudaStream_t stream);
// thr_out and q_out may each be null (not both).  ton_in may be null (tonality computed internally).
cudaError_t pa_threshold(const PaDeviceTables& tb, const float* y, const float* ton_in, float drown, float thr_scale,
                         float* thr_out, int32_t* q_out, int64_t rows, int channels, cudaStream_t stream);

cudaError_t quantize(const float* y, const float* thr, int32_t* q, int64_t n, cudaStream_t stream);
cudaError_t dequantize(const int32_t* q, const float* thr, float* y, int64_t n, cudaStream_t stream);
// stats[0..2] += { n, non-zero integers, sum of log2(2|q|+1) in 16.16 fixed point }
cudaError_t codec_stats(const int32_t* q, int64_t n, unsigned long long* stats, cudaStream_t stream);
cudaError_t add_noise(const float* y, const float* thr, float* out, int64_t n, uint64_t seed, cudaStream_t stream);

// compact side information: q and G [rows][64][channels] instead of q and thr [rows][N][channels]; only on the
// tensor-core tile kernel (cudaErrorNotSupported otherwise).  pa_expand_threshold: thr = sqrt(G W_inv), the same
// operations in the same order as phase D of the tile kernel (bit-identical step sizes in encoder and decoder).
cudaError_t pa_encode_compact(const PaDeviceTables& tb, const float* y, float drown, float thr_scale, float* bark_out,
                              int32_t* q_out, int64_t rows, int channels, cudaStream_t stream);
// mdct_tile_kernels.cu: inverse MDCT of q * sqrt(G W_inv), the expansion fused into the dequantisation (N = 256 / 512 /
// 1024, 1 or 2 channels; cudaErrorInvalidConfiguration otherwise)
cudaError_t mdct_inverse_compact_tile(const MdctDeviceTables& tb, const int32_t* q, const float* bark, const float4* filt4,
                                      float eps_s2, float* x, int64_t batches, int64_t frames_n, int channels,
                                      cudaStream_t stream);
cudaError_t pa_expand_threshold(const PaDeviceTables& tb, const float* bark, float thr_scale, float* thr, int64_t rows,
                                int channels, cudaStream_t stream);

}  // namespace ac
