/* audiocodec_b200 — C ABI of the B200-native audiocodec hot path.
 *
 * The reference (korneelvdbroek/audiocodec, /root/reference) has no FFI layer: its boundary is two Python
 * classes whose methods dispatch TensorFlow ops.  This header is the boundary a binding would target
 * instead; each entry point names the reference interface it replaces (file:line under
 * /root/reference/audiocodec/).  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types; every function returns an ac_status (0 = OK,
 *     negative = error) and records a thread-local message retrievable with ac_last_error();
 *   - the library never allocates caller-visible buffers and never synchronises: all device work is
 *     enqueued on the caller's cudaStream_t (passed as void*; NULL = legacy default stream);
 *   - data pointers are DEVICE pointers on the current CUDA device, fp32, C-contiguous, 16-byte aligned;
 *     layouts are the reference's channels-last ones:  signal x [B, S, C],  amplitudes Y [B, M, N, C],
 *     tonality [B, M, 1, C];
 *   - plans are immutable after creation (device tables only) and may be shared between threads/streams
 *     of the device they were created on;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     AC_ERR_CUDA.  Only the *_host table builders run without a GPU.
 */
#ifndef AUDIOCODEC_B200_H_
#define AUDIOCODEC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

#define AC_ABI_VERSION 1

typedef enum ac_status {
  AC_OK = 0,
  AC_ERR_INVALID = -1,      /* bad argument (shape, alignment, null pointer, odd filters_n, ...) */
  AC_ERR_CUDA = -2,         /* CUDA runtime error (message carries cudaGetErrorString) */
  AC_ERR_UNSUPPORTED = -3,  /* valid in the reference but not built here (e.g. bf16 compute) */
  AC_ERR_ALLOC = -4
} ac_status;

typedef enum ac_window {     /* mdctransformer.py:199-211 */
  AC_WINDOW_ONES = 0,        /* any string other than 'sine'/'vorbis' */
  AC_WINDOW_SINE = 1,
  AC_WINDOW_VORBIS = 2
} ac_window;

typedef enum ac_compute_dtype {   /* compute_dtype of the two reference classes (mdctransformer.py:13, psychoacoustic.py:14) */
  AC_DTYPE_F32 = 0,                /* tf.float32 (default); float64 tensors run on the same plans through the _f64 entry points */
  AC_DTYPE_BF16 = 1                /* tf.bfloat16: tables and constants rounded to bfloat16, see the _bf16 entry points */
} ac_compute_dtype;

typedef struct ac_mdct_plan ac_mdct_plan;
typedef struct ac_pa_plan ac_pa_plan;
struct DLManagedTensor;      /* dlpack.h, DLPack v0.x ABI (the capsule named "dltensor") */

/* ---------------------------------------------------------------------------------------- library */
const char* ac_last_error(void);
int ac_abi_version(void);
/* Number of launches of this library's own kernels since load (all plans, this process). */
int64_t ac_kernel_launch_count(void);

/* ------------------------------------------------------------------------- host-side table builders */
/* Sparse form of the fold matrices (mdctransformer.py:192-229 F, :176-190 inv(F)); h = N/2 pairs.
 *   fold[4p..4p+3]   = { F[p, h-1-p], F[N-1-p, h-1-p], F[p, h+p], F[N-1-p, h+p] }
 *   unfold[4p..4p+3] = { Finv[h-1-p, p], Finv[h+p, p], Finv[h-1-p, N-1-p], Finv[h+p, N-1-p] }
 * precompute_f32 != 0 evaluates the window in float32 (precompute_dtype=tf.float32). */
int ac_mdct_tables_host(int filters_n, int window_type, int precompute_f32, double* fold, double* unfold);

/* Dense tables of PsychoacousticModel.__init__ (psychoacoustic.py:61-69), float64 precompute, cast to fp32:
 *   W [N, nb], W_inv [nb, N], quiet [nb], spreading [nb, nb];
 *   scalars[0..3] = { max_frequency, max_bark, bark_band_width, dB_MIN }.  Any output may be NULL. */
int ac_pa_tables_host(double sample_rate, int filter_bands_n, int bark_bands_n, double alpha,
                      float* W, float* W_inv, float* quiet, float* spreading, double* scalars);

/* Host-side job list of the tensor-core masking kernel (no CUDA needed; for tests and tooling): the filter axis in
 * chunks of counts[0] filters; jobs[4j..4j+3] = { byte offset of the first row of the chunk buffer (264 bytes per
 * filter), byte offset of the job's zero-padded weights (each stored twice), steps of four filters, byte offset of
 * the band's row of P (256 bytes per band, swizzle folded in) | starts-in-an-earlier-chunk << 16 | band-complete << 17 };
 * job_start[9c + w] .. job_start[9c + w + 1]: jobs of warp w in chunk c; ton_start likewise for the filters of the
 * tonality pass; weights[counts[3]]; counts = { chunk, chunks, jobs, weights, fits-the-kernel-parameter }.
 * Arrays may be NULL; size them for 112 jobs and 32 chunks. */
int ac_pa_mma_jobs_host(double sample_rate, int filter_bands_n, int bark_bands_n, double alpha, int32_t* jobs,
                        int32_t* job_start, int16_t* ton_start, float* weights, int32_t* counts);

/* ------------------------------------------------------------------------------------------- MDCT */
/* MDCTransformer.__init__ (mdctransformer.py:13-59).  filters_n must be even (AC_ERR_INVALID otherwise,
 * the reference asserts at :26).  Tables go to the current device. */
int ac_mdct_plan_create(int filters_n, int window_type, int precompute_f32, ac_mdct_plan** out);
/* The same with the compute dtype: AC_DTYPE_BF16 casts H / H_inv and the scale constants to bfloat16 as
 * MDCTransformer(compute_dtype=tf.bfloat16) does (mdctransformer.py:58-59, 125, 145, 347). */
int ac_mdct_plan_create_ex(int filters_n, int window_type, int precompute_f32, int compute_dtype, ac_mdct_plan** out);
int ac_mdct_plan_destroy(ac_mdct_plan* plan);

/* MDCTransformer.transform (mdctransformer.py:61-125):  x [B, S, C] -> y [B, S/N + 1, N, C].
 * S must be a multiple of filters_n (the reference raises at :287). */
int ac_mdct_forward_f32(const ac_mdct_plan* plan, const float* x, float* y,
                        int64_t batches, int64_t samples, int channels, void* stream);

/* MDCTransformer.inverse_transform (mdctransformer.py:127-153):  y [B, M, N, C] -> x [B, (M+1) N, C]. */
int ac_mdct_inverse_f32(const ac_mdct_plan* plan, const float* y, float* x,
                        int64_t batches, int64_t blocks, int channels, void* stream);

/* Decoder fusion: inverse_transform(q * thr) without materialising the dequantised amplitudes.
 * q int32 [B, M, N, C], thr fp32 [B, M, N, C]. */
int ac_mdct_inverse_dequant_f32(const ac_mdct_plan* plan, const int32_t* q, const float* thr, float* x,
                                int64_t batches, int64_t blocks, int channels, void* stream);

/* ---------------------------------------------------------------------------------- psychoacoustics */
/* PsychoacousticModel.__init__ (psychoacoustic.py:14-69); compute dtype fp32, precompute float64. */
int ac_pa_plan_create(double sample_rate, int filter_bands_n, int bark_bands_n, double alpha, ac_pa_plan** out);
/* The same with the compute dtype: AC_DTYPE_BF16 casts W, W_inv, the quiet threshold, the spreading matrix, the
 * bark-axis linspace, eps, alpha and 1 / alpha to bfloat16 (psychoacoustic.py:56, 65-69, 187-189, 197, 206-208). */
int ac_pa_plan_create_ex(double sample_rate, int filter_bands_n, int bark_bands_n, double alpha, int compute_dtype,
                         ac_pa_plan** out);
int ac_pa_plan_destroy(ac_pa_plan* plan);

/* PsychoacousticModel.tonality (psychoacoustic.py:102-120):  y [B, M, N, C] -> ton [B, M, 1, C]. */
int ac_pa_tonality_f32(const ac_pa_plan* plan, const float* y, float* ton,
                       int64_t batches, int64_t blocks, int channels, void* stream);

/* PsychoacousticModel.global_masking_threshold (psychoacoustic.py:122-148, 169-210, 301-331):
 * y [B, M, N, C], ton [B, M, 1, C] (NULL = compute the tonality of y internally) -> thr [B, M, N, C]. */
int ac_pa_threshold_f32(const ac_pa_plan* plan, const float* y, const float* ton, float drown, float* thr,
                        int64_t batches, int64_t blocks, int channels, void* stream);

/* Encoder fusion: tonality -> threshold -> q = rint(y / (thr_scale * thr)) in one pass over y.
 * thr_out (may be NULL) receives thr_scale * thr, i.e. the step actually used; q int32 [B, M, N, C]. */
int ac_pa_encode_f32(const ac_pa_plan* plan, const float* y, float drown, float thr_scale,
                     float* thr_out, int32_t* q,
                     int64_t batches, int64_t blocks, int channels, void* stream);

/* Compact side information (the decoder-side mapping is psychoacoustic.py:330-331): the encoder
 * writes q and the 64 bark-domain thresholds per (frame, channel) - bark_thr [B, M, 64, C], intensities with
 * thr_scale^2 folded in - instead of one threshold per coefficient; ac_pa_expand_threshold_f32 rebuilds
 * step [B, M, N, C] = sqrt(bark_thr W_inv) with the operations of the encoder, bit-identical to the step q was
 * quantised with.  AC_ERR_UNSUPPORTED unless bark_bands_n == 64 with <= 3 bands per filter and 1/2/4 channels. */
int ac_pa_encode_compact_f32(const ac_pa_plan* plan, const float* y, float drown, float thr_scale,
                             float* bark_thr, int32_t* q,
                             int64_t batches, int64_t blocks, int channels, void* stream);
int ac_pa_expand_threshold_f32(const ac_pa_plan* plan, const float* bark_thr, float thr_scale, float* thr,
                               int64_t batches, int64_t blocks, int channels, void* stream);
/* Decoder fusion on the compact side information: x = inverse MDCT of q * sqrt(bark_thr W_inv), the expansion done
 * inside the dequantising inverse kernel (it reads q and N / 64 times fewer threshold bytes).  Same result, bit for
 * bit, as ac_pa_expand_threshold_f32 followed by ac_mdct_inverse_dequant_f32.  filters_n 256 / 512 / 1024, 1 or 2
 * channels; AC_ERR_UNSUPPORTED otherwise. */
int ac_mdct_inverse_dequant_compact_f32(const ac_mdct_plan* plan, const ac_pa_plan* pa_plan, const int32_t* q,
                                        const float* bark_thr, float thr_scale, float* x,
                                        int64_t batches, int64_t blocks, int channels, void* stream);

/* Single-pass encoder: MDCTransformer.transform (mdctransformer.py:61-125) -> tonality ->
 * global_masking_threshold (psychoacoustic.py:102-148) -> quantiser in ONE kernel,
 *   x [B, S, C] -> q int32 [B, S/N + 1, N, C], step_out fp32 same shape (may be NULL) and / or
 *   bark_thr_out [B, S/N + 1, 64, C] (may be NULL; the compact side information above);
 * the amplitudes stay in shared memory and never reach global memory: 4 bytes read and 4 + (4 | 1) bytes written per
 * sample instead of 20.  Bit-identical to ac_mdct_forward_f32 + ac_pa_encode_f32 / ac_pa_encode_compact_f32.
 * Fused for stereo signals with filters_n = 256 (ac_codec_encode_workspace_bytes returns 0: workspace may be NULL);
 * every other shape runs the two kernels through `workspace`, device memory of that many bytes (the amplitudes). */
int64_t ac_codec_encode_workspace_bytes(const ac_mdct_plan* mdct, const ac_pa_plan* pa, int64_t batches,
                                        int64_t samples, int channels);
int ac_codec_encode_f32(const ac_mdct_plan* mdct, const ac_pa_plan* pa, const float* x, float drown, float thr_scale,
                        float* step_out, float* bark_thr_out, int32_t* q,
                        int64_t batches, int64_t samples, int channels, void* workspace, void* stream);

/* PsychoacousticModel.add_noise (psychoacoustic.py:150-167): out = y + thr * N(0, 1/6), counter-based RNG. */
int ac_pa_add_noise_f32(const float* y, const float* thr, float* out, int64_t n, uint64_t seed, void* stream);

/* -------------------------------------------------------------------------------------- quantiser */
/* Build-defined (the reference has no quantiser; spec from add_noise, psychoacoustic.py:150-167):
 * q = rint(y / thr) with IEEE division and round-half-even;  y_hat = q * thr. */
int ac_quantize_f32(const float* y, const float* thr, int32_t* q, int64_t n, void* stream);
int ac_dequantize_f32(const int32_t* q, const float* thr, float* y, int64_t n, void* stream);

/* Bitstream statistics of a tensor of quantised integers (the quantities gathered across ranks at the end of a job,
 * SURVEY.md 8e; no reference symbol): stats_dev[0..2] += { n, non-zero integers, sum log2(2|q|+1) in 16.16 fixed
 * point }.  stats_dev is DEVICE memory (three 64-bit counters the caller zeroes); integer accumulation, so the
 * result does not depend on the order of the atomics.  Basis of the rate loop in the Python layer. */
int ac_codec_stats_i32(const int32_t* q, int64_t n, uint64_t* stats_dev, void* stream);

/* ------------------------------------------------------------------------- float64 compute dtype */
/* The reference accepts compute_dtype=tf.float64 for both classes (mdctransformer.py:13-23,
 * psychoacoustic.py:42-44).  Same operations and layouts as the _f32 entry points on float64 tensors with the
 * unrounded float64 tables; functional kernels (any even filters_n), not the tuned fp32 path. */
int ac_mdct_forward_f64(const ac_mdct_plan* plan, const double* x, double* y,
                        int64_t batches, int64_t samples, int channels, void* stream);
int ac_mdct_inverse_f64(const ac_mdct_plan* plan, const double* y, double* x,
                        int64_t batches, int64_t blocks, int channels, void* stream);
int ac_pa_tonality_f64(const ac_pa_plan* plan, const double* y, double* ton,
                       int64_t batches, int64_t blocks, int channels, void* stream);
int ac_pa_threshold_f64(const ac_pa_plan* plan, const double* y, const double* ton, double drown, double* thr,
                        int64_t batches, int64_t blocks, int channels, void* stream);
int ac_quantize_f64(const double* y, const double* thr, int32_t* q, int64_t n, void* stream);
int ac_dequantize_f64(const int32_t* q, const double* thr, double* y, int64_t n, void* stream);

/* ------------------------------------------------------------------------------- backward pass */
/* The reference's methods are @tf.function graphs of differentiable ops (psychoacoustic.py:102, :122; the comment at
 * :311 speaks of the gradient): these are the vector-Jacobian products that make the drop-in classes differentiable
 * layers too.  grad_y [B, M, N, C] is written (not accumulated); grad_ton [B, M, 1, C] may be NULL.  Clamps
 * (max(eps, .), min(., 1), the quiet-threshold branch) pass no gradient where they are active.  The MDCT needs no entry
 * point: for the orthogonal windows (sine, vorbis) the adjoint of ac_mdct_forward_f32 is ac_mdct_inverse_f32 / 4N
 * cropped by one block at both ends, and the adjoint of ac_mdct_inverse_f32 is 4N ac_mdct_forward_f32 without its
 * first and last frame (audiocodec_b200/autograd.py). */
int ac_pa_tonality_backward_f32(const ac_pa_plan* plan, const float* y, const float* grad_ton, float* grad_y,
                                int64_t batches, int64_t blocks, int channels, void* stream);
int ac_pa_threshold_backward_f32(const ac_pa_plan* plan, const float* y, const float* ton, float drown,
                                 const float* grad_thr, float* grad_y, float* grad_ton,
                                 int64_t batches, int64_t blocks, int channels, void* stream);

/* ----------------------------------------------------------------------- entropy-coded bitstream */
/* Adaptive Golomb-Rice coding of the quantised integers (no reference symbol - the reference has no quantiser and no
 * bitstream; the format is build-defined: csrc/entropy_kernels.cu, restated bit for bit by
 * oracle/entropy_oracle.py).  q is seen as `rows` rows of `row_len` integers (row_len a multiple of 16; the codec uses
 * one frame = filters_n x channels_n); every row becomes an independent byte range of the stream that starts on a 4-byte
 * boundary, so rows decode in parallel.  Per 16 values: a 5-bit Rice parameter (31 = all zero), then zigzag(q) as
 * (u >> k) zeros, a one, k low bits.
 *   ac_entropy_plan_i32   offsets[0 .. rows] (device, int64): byte offset of every row; offsets[rows] = stream size
 *   ac_entropy_encode_i32 writes the stream into bytes (device, 4-byte aligned, offsets[rows] + 4 bytes)
 *   ac_entropy_decode_i32 reads it back (reads up to 4 bytes behind the last row) */
int ac_entropy_plan_i32(const int32_t* q, int64_t rows, int64_t row_len, int64_t* offsets, void* stream);
int ac_entropy_encode_i32(const int32_t* q, int64_t rows, int64_t row_len, const int64_t* offsets, uint8_t* bytes, void* stream);
int ac_entropy_decode_i32(const uint8_t* bytes, const int64_t* offsets, int64_t rows, int64_t row_len, int32_t* q, void* stream);

/* ------------------------------------------------------------------------ bfloat16 compute dtype */
/* compute_dtype=tf.bfloat16 (psychoacoustic.py:42-44; mdctransformer.py:326-344 up-casts to float32 around the DCT):
 * tensors are bfloat16 at the boundary (void* = device pointers to bfloat16), the plan (created with AC_DTYPE_BF16)
 * holds bfloat16-valued tables and constants, and the arithmetic in between runs in float32 - the reference's rule for
 * the DCT applied to the whole path; TensorFlow's per-op rounding of intermediates to bfloat16 is NOT reproduced
 * (it cannot be pinned without TensorFlow; the result is the closer one to the float64 value).  The float32 copies of
 * inputs and outputs live in `workspace`: device memory of ac_bf16_workspace_bytes(input elements, output elements)
 * bytes (for ac_pa_threshold_bf16 the inputs are the amplitudes plus, padded to a multiple of 4, the tonality).
 * Functional path (three launches), not tuned: the north star is float32. */
int64_t ac_bf16_workspace_bytes(int64_t in_elems, int64_t out_elems);
int ac_mdct_forward_bf16(const ac_mdct_plan* plan, const void* x, void* y,
                         int64_t batches, int64_t samples, int channels, void* workspace, void* stream);
int ac_mdct_inverse_bf16(const ac_mdct_plan* plan, const void* y, void* x,
                         int64_t batches, int64_t blocks, int channels, void* workspace, void* stream);
int ac_pa_tonality_bf16(const ac_pa_plan* plan, const void* y, void* ton,
                        int64_t batches, int64_t blocks, int channels, void* workspace, void* stream);
int ac_pa_threshold_bf16(const ac_pa_plan* plan, const void* y, const void* ton, float drown, void* thr,
                         int64_t batches, int64_t blocks, int channels, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------ dB utilities */
/* PsychoacousticModel.amplitude_to_dB (psychoacoustic.py:71-85): out = 10 ln(max(eps, a^2)) / ln 10 + 120;
 * normalised != 0: amplitude_to_dB_norm (:87-100), (dB - dB_MIN) / (dB_MAX - dB_MIN) in [0, 1].  n elements. */
int ac_pa_amplitude_to_db_f32(const ac_pa_plan* plan, const float* a, float* out, int64_t n, int normalised, void* stream);
int ac_pa_amplitude_to_db_f64(const ac_pa_plan* plan, const double* a, double* out, int64_t n, int normalised, void* stream);

/* ------------------------------------------------------------------------- host-buffer streaming */
/* encode + decode of clips that live in HOST memory (the call a file / network front end makes; no reference
 * symbol - the reference leaves data movement to TensorFlow).  The pipeline owns three streams, a ring of four
 * chunk-sized device buffers for x and for x_hat and one chunk of amplitudes / steps / integers - its device
 * footprint does not depend on the size of the host batch (cfg5's 87 GB stream through one GPU);
 * x_host [B, S, C] flows in chunks of `chunk_clips` clips through
 * H2D -> ac_mdct_forward -> ac_pa_encode -> ac_mdct_inverse_dequant -> D2H into xhat_host [B, S + 2N, C] with
 * both PCIe directions and the kernels overlapped.  Pinned host memory gives asynchronous copies (pageable
 * memory works, serialised).  ac_codec_roundtrip_host_f32 orders itself after the work already enqueued on
 * `stream` and returns when xhat_host is complete.  stats (may be NULL) receives
 * { coefficients, non-zero integers, sum log2(2|q|+1) } accumulated over the call, as doubles. */
typedef struct ac_codec_pipeline ac_codec_pipeline;
int ac_codec_pipeline_create(const ac_mdct_plan* mdct, const ac_pa_plan* pa, int64_t chunk_clips, int64_t samples,
                             int channels, ac_codec_pipeline** out);
int ac_codec_pipeline_destroy(ac_codec_pipeline* pipe);
int ac_codec_roundtrip_host_f32(ac_codec_pipeline* pipe, const float* x_host, float* xhat_host, int64_t batches,
                                float drown, float thr_scale, double* stats, void* stream);

/* ------------------------------------------------------------------------------------------ DLPack */
/* Same operations on DLManagedTensor* (what `tensor.__dlpack__()` / tf.experimental.dlpack.to_dlpack put
 * in the "dltensor" capsule).  Tensors are validated (kDLCUDA, float32 / int32, rank, compact strides,
 * 16-byte alignment incl. byte_offset) and never copied or consumed: the caller keeps ownership. */
int ac_mdct_forward_dl(const ac_mdct_plan* plan, struct DLManagedTensor* x, struct DLManagedTensor* y, void* stream);
int ac_mdct_inverse_dl(const ac_mdct_plan* plan, struct DLManagedTensor* y, struct DLManagedTensor* x, void* stream);
int ac_pa_tonality_dl(const ac_pa_plan* plan, struct DLManagedTensor* y, struct DLManagedTensor* ton, void* stream);
int ac_pa_threshold_dl(const ac_pa_plan* plan, struct DLManagedTensor* y, struct DLManagedTensor* ton_or_null,
                       float drown, struct DLManagedTensor* thr, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* AUDIOCODEC_B200_H_ */
