// C ABI (include/audiocodec_b200.h): plan objects, argument validation, DLPack unwrapping.
#include "../../include/audiocodec_b200.h"

#include "kernels.h"
#include "tables.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

// ---- minimal DLPack v0.x ABI (struct layout of dlpack.h; only what is needed to read a tensor) ----------
extern "C" {
typedef enum { ac_kDLCPU = 1, ac_kDLCUDA = 2, ac_kDLCUDAHost = 3, ac_kDLCUDAManaged = 13 } ac_DLDeviceType;
typedef struct { int32_t device_type; int32_t device_id; } ac_DLDevice;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } ac_DLDataType;   // code: 0 int, 1 uint, 2 float
typedef struct {
  void* data;
  ac_DLDevice device;
  int32_t ndim;
  ac_DLDataType dtype;
  int64_t* shape;
  int64_t* strides;   // in elements; NULL = compact row-major
  uint64_t byte_offset;
} ac_DLTensor;
struct DLManagedTensor {
  ac_DLTensor dl_tensor;
  void* manager_ctx;
  void (*deleter)(struct DLManagedTensor* self);
};
}

namespace {

thread_local char g_error[512] = "";
std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t err, const char* what) {
  return fail(AC_ERR_CUDA, "%s: %s", what, cudaGetErrorString(err));
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
cudaError_t upload(const std::vector<T>& host, const T** dev, std::vector<void*>& owned) {
  void* p = nullptr;
  cudaError_t err = cudaMalloc(&p, std::max<size_t>(host.size(), 1) * sizeof(T));
  if (err != cudaSuccess) return err;
  owned.push_back(p);
  if (!host.empty()) {
    err = cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice);
    if (err != cudaSuccess) return err;
  }
  *dev = static_cast<const T*>(p);
  return cudaSuccess;
}

}  // namespace

namespace ac {
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
bool pdl_enabled(int kind) {
  static const int mask = [] {
    const char* e = std::getenv("AC_PDL");
    // measured on cfg2 (tools/chain_probe.py): the masking kernel behind the forward MDCT gains 2.5 us per step, the
    // forward MDCT behind the inverse nothing, the inverse MDCT behind the masking kernel LOSES 15 us - off by default
    return e != nullptr ? std::atoi(e) : 3;
  }();
  return (mask & kind) != 0;
}
}  // namespace ac

struct ac_mdct_plan {
  int device = 0;
  int compute_dtype = 0;     // ac_compute_dtype
  ac::MdctDeviceTables tb;
  ac::MdctDeviceTables64 tb64;
  std::vector<void*> owned;
};

struct ac_pa_plan {
  int device = 0;
  int compute_dtype = 0;     // ac_compute_dtype
  ac::PaDeviceTables tb;
  ac::PaDeviceTables64 tb64;
  ac::PaJobParams jobs;
  ac::PaTables host;
  std::vector<void*> owned;
};

namespace {

void free_all(std::vector<void*>& owned) {
  for (void* p : owned) cudaFree(p);
  owned.clear();
}

int check_window(int window_type) {
  return (window_type == AC_WINDOW_ONES || window_type == AC_WINDOW_SINE || window_type == AC_WINDOW_VORBIS) ? 0 : -1;
}

// nearest-even rounding of a double to a bfloat16 value (through float32, as tf.cast(float64 -> bfloat16) does)
double bf16_round(double v) {
  const float f = static_cast<float>(v);
  uint32_t u;
  std::memcpy(&u, &f, sizeof(u));
  if ((u & 0x7f800000u) == 0x7f800000u) return f;      // inf / nan
  u += 0x7fffu + ((u >> 16) & 1u);
  u &= 0xffff0000u;
  float r;
  std::memcpy(&r, &u, sizeof(r));
  return r;
}

// validates a DLPack tensor and returns its data pointer (incl. byte_offset)
int unwrap_dl(struct DLManagedTensor* m, const char* name, int ndim, uint8_t code, void** data, const int64_t** shape) {
  if (m == nullptr) return fail(AC_ERR_INVALID, "%s: null DLManagedTensor", name);
  const ac_DLTensor& t = m->dl_tensor;
  if (t.device.device_type != ac_kDLCUDA && t.device.device_type != ac_kDLCUDAManaged)
    return fail(AC_ERR_INVALID, "%s: tensor is not on a CUDA device (DLDeviceType %d); there is no CPU path", name,
                t.device.device_type);
  int dev = -1;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return cuda_fail(err, "cudaGetDevice");
  if (t.device.device_id != dev)
    return fail(AC_ERR_INVALID, "%s: tensor lives on cuda:%d but the current device is cuda:%d", name, t.device.device_id, dev);
  if (t.dtype.code != code || t.dtype.bits != 32 || t.dtype.lanes != 1)
    return fail(AC_ERR_INVALID, "%s: dtype must be %s32 (input dtype must equal compute dtype, no implicit cast)", name,
                code == 2 ? "float" : "int");
  if (t.ndim != ndim) return fail(AC_ERR_INVALID, "%s: expected rank %d, got %d", name, ndim, t.ndim);
  if (t.strides != nullptr) {
    int64_t expect = 1;
    for (int d = ndim - 1; d >= 0; --d) {
      if (t.shape[d] != 1 && t.strides[d] != expect) return fail(AC_ERR_INVALID, "%s: tensor must be C-contiguous", name);
      expect *= t.shape[d];
    }
  }
  *data = static_cast<char*>(t.data) + t.byte_offset;
  *shape = t.shape;
  return AC_OK;
}

}  // namespace

static int __float_as_int_host(float f) {
  int i;
  std::memcpy(&i, &f, sizeof(i));
  return i;
}

// Band-sum job list of the tensor-core tile kernel (psycho_mma_kernels.cu; PaJobParams in kernels.h): chunks of 64
// filters, one job per (band, chunk) = steps of four filters with zero-padded weights; the jobs of a chunk are dealt
// to the eight warps in contiguous runs of equal cost, and the filters of the chunk's tonality pass in contiguous runs
// that level the warps' totals.  Host-only (no CUDA): ac_pa_mma_jobs_host exposes it to the CPU tests.  Returns
// false when the list does not fit the kernel parameter (the kernel is then not used).
static bool build_mma_jobs_chunk(const ac::PaTables& t, ac::PaJobParams& jp, std::vector<float>& mma_w4, int* chunk_k,
                                 int* n_chunks_out, int* n_jobs_out, int mma_chunk_k);

// filters per chunk of the masking kernel's double-buffered tile: 64, a compile-time constant of the kernel (kChunk in
// psycho_mma_kernels.cu; 32-filter chunks with four CTAs per SM measured slower on B200, profiles/README.md)
static bool build_mma_jobs(const ac::PaTables& t, ac::PaJobParams& jp, std::vector<float>& mma_w4, int* chunk_k,
                           int* n_chunks_out, int* n_jobs_out) {
  mma_w4.clear();
  return build_mma_jobs_chunk(t, jp, mma_w4, chunk_k, n_chunks_out, n_jobs_out, 64);
}

// instructions per lane of a step of four filters, of the tail of a job that completes / does not complete its band, and
// of one filter of the tonality pass: what the split of a chunk's work between the eight warps is levelled by
// (AC_PA_COST="step,final,partial,ton" overrides it when a plan is created: experiments)
struct PaCostModel {
  double step = 13.5, final = 45.0, partial = 25.0, ton_row = 9.0;
};

static bool build_mma_jobs_chunk(const ac::PaTables& t, ac::PaJobParams& jp, std::vector<float>& mma_w4, int* chunk_k,
                                 int* n_chunks_out, int* n_jobs_out, const int mma_chunk_k) {
  PaCostModel cm;
  if (const char* e = std::getenv("AC_PA_COST")) std::sscanf(e, "%lf,%lf,%lf,%lf", &cm.step, &cm.final, &cm.partial, &cm.ton_row);
  const int mma_n_chunks = (t.n + mma_chunk_k - 1) / mma_chunk_k;
  std::vector<int4> job_desc;
  std::vector<int32_t> job_start(static_cast<size_t>(mma_n_chunks) * 9 + 1, 0);
  std::vector<int16_t> ton_start(static_cast<size_t>(mma_n_chunks) * 9 + 1, 0);
  for (int c = 0; c < mma_n_chunks; ++c) {
    const int kc0 = c * mma_chunk_k, kc1 = std::min(t.n, kc0 + mma_chunk_k);
    const size_t first_job = job_desc.size();
    std::vector<double> cost;
    for (int i = 0; i < t.nb; ++i) {
      const int ka = std::max(t.band_k0[i], kc0), kb = std::min(t.band_k0[i] + t.band_cnt[i], kc1);
      if (kb <= ka) continue;
      const int steps = (kb - ka + 3) / 4;
      const bool final = t.band_k0[i] + t.band_cnt[i] <= kc1, partial = t.band_k0[i] < kc0;
      int4 jb;
      jb.x = (ka - kc0) * 66 * 4;                      // byte offset of the first row of T (66 words per row)
      jb.y = static_cast<int>(mma_w4.size()) * 8;      // byte offset of the weights, each stored twice (packed pairs)
      jb.z = steps;
      // low 16 bits: byte offset of the band's row of P in the operand layout of the tcgen05 product (psycho_mma_kernels.cu,
      // tc_desc): band i at position kp = i ^ 4, atoms of four bands 512 bytes apart, 128 bytes per band, 32-byte chunks
      // XORed with kp & 3 (the kernel xors the lane's part: a lane owns the item pair 2 lane, 2 lane + 1).  Bits 18 ..: byte
      // offset of P[band][0] in the layout of the mma.sync product (64 items per band, XOR swizzle of the band folded in)
      const int kp = i ^ 4;
      const unsigned tc_off = static_cast<unsigned>((kp >> 2) * 512 + (kp & 3) * 128 + ((kp & 3) << 5));
      const unsigned sync_off = static_cast<unsigned>(i * 256 + ((i & 3) << 5));
      jb.w = static_cast<int>(tc_off | (partial ? 0x10000u : 0u) | (final ? 0x20000u : 0u) | (sync_off << 18));
      for (int k = ka; k < ka + 4 * steps; ++k)
        mma_w4.push_back(k < kb ? t.band_w[t.band_ptr[i] + (k - t.band_k0[i])] : 0.f);
      job_desc.push_back(jb);
      cost.push_back(cm.step * steps + (final ? cm.final : cm.partial));   // ~instructions per lane
    }
    double total = 0, run = 0;
    for (double v : cost) total += v;
    int w = 0;
    job_start[static_cast<size_t>(c) * 9] = static_cast<int32_t>(first_job);
    for (size_t j = 0; j < cost.size(); ++j) {
      run += cost[j];
      while (w < 7 && run >= total * (w + 1) / 8.0)
        job_start[static_cast<size_t>(c) * 9 + (++w)] = static_cast<int32_t>(first_job + j + 1);
    }
    while (w < 8) job_start[static_cast<size_t>(c) * 9 + (++w)] = static_cast<int32_t>(job_desc.size());
    // tonality pass of the chunk (~9 instructions per filter): water-filling, the level T with
    // sum_w max(0, T - load_w) = rows * row_cost
    double load[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int ww = 0; ww < 8; ++ww)
      for (int j = job_start[static_cast<size_t>(c) * 9 + ww]; j < job_start[static_cast<size_t>(c) * 9 + ww + 1]; ++j)
        load[ww] += cost[static_cast<size_t>(j) - first_job];
    const int rows = kc1 - kc0;
    const double row_cost = cm.ton_row;
    double lo = 0, hi = total + rows * row_cost;
    for (int it = 0; it < 60; ++it) {
      const double mid = 0.5 * (lo + hi);
      double fill = 0;
      for (double l : load) fill += std::max(0.0, mid - l);
      (fill < rows * row_cost ? lo : hi) = mid;
    }
    int at = 0;
    for (int ww = 0; ww < 8; ++ww) {
      ton_start[static_cast<size_t>(c) * 9 + ww] = static_cast<int16_t>(at);
      int take = static_cast<int>(std::lround(std::max(0.0, hi - load[ww]) / row_cost));
      take = std::min(take, rows - at);
      if (ww == 7) take = rows - at;
      at += take;
    }
    ton_start[static_cast<size_t>(c) * 9 + 8] = static_cast<int16_t>(rows);
  }
  job_start[static_cast<size_t>(mma_n_chunks) * 9] = static_cast<int32_t>(job_desc.size());
  *chunk_k = mma_chunk_k;
  *n_chunks_out = mma_n_chunks;
  *n_jobs_out = static_cast<int>(job_desc.size());
  if (static_cast<int>(job_desc.size()) > ac::kPaMaxJobs || mma_n_chunks > ac::kPaMaxChunks) return false;
  std::copy(job_desc.begin(), job_desc.end(), jp.job);
  std::copy(job_start.begin(), job_start.end(), jp.start);
  std::copy(ton_start.begin(), ton_start.end(), jp.ton_start);
  return true;
}

extern "C" {

const char* ac_last_error(void) { return g_error; }
int ac_abi_version(void) { return AC_ABI_VERSION; }
int64_t ac_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// ------------------------------------------------------------------------------------ host-side tables
int ac_mdct_tables_host(int filters_n, int window_type, int precompute_f32, double* fold, double* unfold) {
  if (filters_n < 2 || (filters_n % 2) != 0)
    return fail(AC_ERR_INVALID, "number of filters used in mdct transformation needs to be even (got %d)", filters_n);
  if (check_window(window_type) != 0) return fail(AC_ERR_INVALID, "unknown window type %d", window_type);
  const ac::MdctTables t = ac::build_mdct_tables(filters_n, window_type, precompute_f32 != 0);
  if (fold != nullptr) std::memcpy(fold, t.fold.data(), t.fold.size() * sizeof(double));
  if (unfold != nullptr) std::memcpy(unfold, t.unfold.data(), t.unfold.size() * sizeof(double));
  return AC_OK;
}

int ac_pa_tables_host(double sample_rate, int filter_bands_n, int bark_bands_n, double alpha, float* W, float* W_inv,
                      float* quiet, float* spreading, double* scalars) {
  if (!(sample_rate > 0) || filter_bands_n < 1 || bark_bands_n < 1 || !(alpha > 0))
    return fail(AC_ERR_INVALID, "invalid psychoacoustic parameters (sample_rate %g, filter_bands_n %d, bark_bands_n %d, alpha %g)",
                sample_rate, filter_bands_n, bark_bands_n, alpha);
  const ac::PaTables t = ac::build_pa_tables(sample_rate, filter_bands_n, bark_bands_n, alpha);
  const int n = t.n, nb = t.nb;
  if (W != nullptr)
    for (size_t i = 0; i < t.w.size(); ++i) W[i] = static_cast<float>(t.w[i]);
  if (W_inv != nullptr)
    for (size_t i = 0; i < t.w_inv.size(); ++i) W_inv[i] = static_cast<float>(t.w_inv[i]);
  if (quiet != nullptr)
    for (int i = 0; i < nb; ++i) quiet[i] = static_cast<float>(t.quiet[i]);
  if (spreading != nullptr)
    for (int i = 0; i < nb; ++i)
      for (int j = 0; j < nb; ++j) spreading[i * nb + j] = static_cast<float>(t.spread_fn[nb - i + j]);
  if (scalars != nullptr) {
    scalars[0] = t.max_frequency;
    scalars[1] = t.max_bark;
    scalars[2] = t.bark_band_width;
    scalars[3] = t.db_min;
  }
  (void)n;
  return AC_OK;
}

int ac_pa_mma_jobs_host(double sample_rate, int filter_bands_n, int bark_bands_n, double alpha, int32_t* jobs,
                        int32_t* job_start, int16_t* ton_start, float* weights, int32_t* counts) {
  if (!(sample_rate > 0) || filter_bands_n < 1 || bark_bands_n < 1 || !(alpha > 0))
    return fail(AC_ERR_INVALID, "invalid psychoacoustic parameters");
  if (counts == nullptr) return fail(AC_ERR_INVALID, "counts is null");
  const ac::PaTables t = ac::build_pa_tables(sample_rate, filter_bands_n, bark_bands_n, alpha);
  std::vector<float> w4;
  ac::PaJobParams* jp = new (std::nothrow) ac::PaJobParams();
  if (jp == nullptr) return fail(AC_ERR_ALLOC, "out of host memory");
  int chunk_k = 0, n_chunks = 0, n_jobs = 0;
  const bool fits = build_mma_jobs(t, *jp, w4, &chunk_k, &n_chunks, &n_jobs);
  counts[0] = chunk_k;
  counts[1] = n_chunks;
  counts[2] = n_jobs;
  counts[3] = static_cast<int32_t>(w4.size());
  counts[4] = fits ? 1 : 0;
  if (fits) {
    if (jobs != nullptr)
      for (int j = 0; j < n_jobs; ++j) {
        jobs[4 * j + 0] = jp->job[j].x;
        jobs[4 * j + 1] = jp->job[j].y;
        jobs[4 * j + 2] = jp->job[j].z;
        jobs[4 * j + 3] = jp->job[j].w;
      }
    if (job_start != nullptr) std::copy(jp->start, jp->start + 9 * n_chunks + 1, job_start);
    if (ton_start != nullptr) std::copy(jp->ton_start, jp->ton_start + 9 * n_chunks + 1, ton_start);
    if (weights != nullptr) std::copy(w4.begin(), w4.end(), weights);
  }
  delete jp;
  return AC_OK;
}

// ------------------------------------------------------------------------------------------------ MDCT
int ac_mdct_plan_create(int filters_n, int window_type, int precompute_f32, ac_mdct_plan** out) {
  return ac_mdct_plan_create_ex(filters_n, window_type, precompute_f32, AC_DTYPE_F32, out);
}

int ac_mdct_plan_create_ex(int filters_n, int window_type, int precompute_f32, int compute_dtype, ac_mdct_plan** out) {
  if (out == nullptr) return fail(AC_ERR_INVALID, "out is null");
  *out = nullptr;
  if (compute_dtype != AC_DTYPE_F32 && compute_dtype != AC_DTYPE_BF16)
    return fail(AC_ERR_INVALID, "compute_dtype must be AC_DTYPE_F32 or AC_DTYPE_BF16 (float64 runs on any plan)");
  if (filters_n < 2 || (filters_n % 2) != 0)
    return fail(AC_ERR_INVALID, "number of filters used in mdct transformation needs to be even (got %d)", filters_n);
  if (check_window(window_type) != 0) return fail(AC_ERR_INVALID, "unknown window type %d", window_type);
  const bool fast = ac::mdct_has_fast_path(filters_n);
  if (!fast && filters_n > 8192)
    return fail(AC_ERR_UNSUPPORTED, "filters_n = %d: only powers of two in [16, 4096] or any even value <= 8192 are built", filters_n);
  ac_mdct_plan* plan = new (std::nothrow) ac_mdct_plan();
  if (plan == nullptr) return fail(AC_ERR_ALLOC, "out of host memory");
  cudaError_t err = cudaGetDevice(&plan->device);
  if (err != cudaSuccess) {
    delete plan;
    return cuda_fail(err, "cudaGetDevice (no CUDA device? this library has no CPU path)");
  }
  const int n = filters_n, h = n / 2;
  ac::MdctTables t = ac::build_mdct_tables(n, window_type, precompute_f32 != 0);
  const double pi = 3.14159265358979323846;
  double scale_fwd = 1.0 / (n * std::sqrt(2.0)), scale_inv = 2.0 * std::sqrt(2.0);
  plan->compute_dtype = compute_dtype;
  if (compute_dtype == AC_DTYPE_BF16) {
    // compute_dtype=tf.bfloat16: H and H_inv are cast to bfloat16 (mdctransformer.py:58-59), and so are the constants
    // sqrt(2) (:347) and 1 / sqrt(4N), sqrt(4N) (:125, :145); the DCT itself runs in float32 (:326-344)
    for (double& v : t.fold) v = bf16_round(v);
    for (double& v : t.unfold) v = bf16_round(v);
    const double r2 = bf16_round(std::sqrt(2.0)) / std::sqrt(2.0);
    scale_fwd *= r2 * bf16_round(1.0 / std::sqrt(4.0 * n)) * std::sqrt(4.0 * n);
    scale_inv *= r2 * bf16_round(std::sqrt(4.0 * n)) / std::sqrt(4.0 * n);
  }
  std::vector<float4> fold(h), unfold(h);
  std::vector<float2> tw(h), twf(h), twi(h), roots(h);
  for (int p = 0; p < h; ++p) {
    fold[p] = make_float4((float)t.fold[4 * p], (float)t.fold[4 * p + 1], (float)t.fold[4 * p + 2], (float)t.fold[4 * p + 3]);
    unfold[p] = make_float4((float)t.unfold[4 * p], (float)t.unfold[4 * p + 1], (float)t.unfold[4 * p + 2], (float)t.unfold[4 * p + 3]);
    const double ang = -pi * (p + 0.125) / n;
    tw[p] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    twf[p] = make_float2((float)(std::cos(ang) * scale_fwd), (float)(std::sin(ang) * scale_fwd));
    twi[p] = make_float2((float)(std::cos(ang) * scale_inv), (float)(std::sin(ang) * scale_inv));
    const double ra = -2.0 * pi * p / h;
    roots[p] = make_float2((float)std::cos(ra), (float)std::sin(ra));
  }
  // tile-kernel tables (mdct_tile_kernels.cu); variant 0 reads / writes position p (or 2k) first, variant 1 the
  // mirrored position N-1-p (or N-1-2k) first, so the coefficient pairs are swapped
  std::vector<float4> pre_fwd(static_cast<size_t>(4) * h), post_fwd(static_cast<size_t>(2) * h);
  std::vector<float4> pre_inv(static_cast<size_t>(2) * h), post_inv(static_cast<size_t>(2) * h);
  for (int m = 0; m < h; ++m) {
    const double ang = -pi * (m + 0.125) / n, c = std::cos(ang), s = std::sin(ang);
    const bool low = m < h / 2;
    const int p = low ? h - 1 - 2 * m : 2 * m - h;
    double re[4], im[4];
    if (p >= 0 && p < h) {
      const double a0 = t.fold[4 * p], a1 = t.fold[4 * p + 1], a2 = t.fold[4 * p + 2], a3 = t.fold[4 * p + 3];
      if (low) {   // z = alpha + i beta
        re[0] = a0 * c; re[1] = a1 * c; re[2] = -a2 * s; re[3] = -a3 * s;
        im[0] = a0 * s; im[1] = a1 * s; im[2] = a2 * c; im[3] = a3 * c;
      } else {     // z = beta + i alpha
        re[0] = -a0 * s; re[1] = -a1 * s; re[2] = a2 * c; re[3] = a3 * c;
        im[0] = a0 * c; im[1] = a1 * c; im[2] = a2 * s; im[3] = a3 * s;
      }
    } else {
      for (int i = 0; i < 4; ++i) re[i] = im[i] = 0.0;
    }
    // layout [variant][Re / Im][n]: the eight lanes of a quarter-warp read one contiguous 128-byte line
    pre_fwd[(0 * 2 + 0) * h + m] = make_float4((float)re[0], (float)re[1], (float)re[2], (float)re[3]);
    pre_fwd[(0 * 2 + 1) * h + m] = make_float4((float)im[0], (float)im[1], (float)im[2], (float)im[3]);
    pre_fwd[(1 * 2 + 0) * h + m] = make_float4((float)re[1], (float)re[0], (float)re[3], (float)re[2]);
    pre_fwd[(1 * 2 + 1) * h + m] = make_float4((float)im[1], (float)im[0], (float)im[3], (float)im[2]);
    // first store S1 = vx c0 + vy c1, second S2 = vx c2 + vy c3; variant 0: S1 = Re D -> [2k], S2 = -Im D -> [N-1-2k]
    const double fx = c * scale_fwd, fy = s * scale_fwd, ix = c * scale_inv, iy = s * scale_inv;
    post_fwd[0 * h + m] = make_float4((float)fx, (float)-fy, (float)-fy, (float)-fx);
    post_fwd[1 * h + m] = make_float4((float)-fy, (float)-fx, (float)fx, (float)-fy);
    post_inv[0 * h + m] = make_float4((float)ix, (float)-iy, (float)-iy, (float)-ix);
    post_inv[1 * h + m] = make_float4((float)-iy, (float)-ix, (float)ix, (float)-iy);
    // loads L1, L2 (variant 0: Y[2n], Y[N-1-2n]):  Re = L1 k0 + L2 k1,  Im = L1 k2 + L2 k3
    pre_inv[0 * h + m] = make_float4((float)c, (float)-s, (float)s, (float)c);
    pre_inv[1 * h + m] = make_float4((float)-s, (float)c, (float)c, (float)s);
  }
  std::vector<float2> tw_pass1, tw_pass2;
  {
    int r0 = 1, r1 = 1, r2 = 1;
    if (ac::tile_fft_radices(n, &r0, &r1, &r2)) {
      auto build = [&](int radix, int ns, std::vector<float2>& out) {
        const int per = h / radix;                    // butterflies of the pass
        out.resize(static_cast<size_t>(radix - 1) * per);
        for (int r = 1; r < radix; ++r)
          for (int j = 0; j < per; ++j) {
            const double ang = -2.0 * pi * r * (j % ns) / (static_cast<double>(ns) * radix);
            out[static_cast<size_t>(r - 1) * per + j] = make_float2((float)std::cos(ang), (float)std::sin(ang));
          }
      };
      if (r1 > 1) build(r1, r0, tw_pass1);
      if (r2 > 1) build(r2, r0 * r1, tw_pass2);
    }
  }
  // the O(N^2) kernels' table: generic N, and power-of-two N whose any-channel tile does not fit in shared memory
  std::vector<float> cos_table(static_cast<size_t>(8) * n);
  for (int m = 0; m < 8 * n; ++m) cos_table[m] = (float)std::cos(pi * m / (4.0 * n));
  plan->tb.n = n;
  plan->tb.scale_fwd = (float)scale_fwd;
  plan->tb.scale_inv = (float)scale_inv;
  if ((err = upload(fold, &plan->tb.fold, plan->owned)) != cudaSuccess ||
      (err = upload(unfold, &plan->tb.unfold, plan->owned)) != cudaSuccess ||
      (err = upload(tw, &plan->tb.tw_pre, plan->owned)) != cudaSuccess ||
      (err = upload(twf, &plan->tb.tw_post_fwd, plan->owned)) != cudaSuccess ||
      (err = upload(twi, &plan->tb.tw_post_inv, plan->owned)) != cudaSuccess ||
      (err = upload(roots, &plan->tb.roots, plan->owned)) != cudaSuccess ||
      (err = upload(cos_table, &plan->tb.cos_table, plan->owned)) != cudaSuccess ||
      (err = upload(pre_fwd, &plan->tb.pre_fwd, plan->owned)) != cudaSuccess ||
      (err = upload(post_fwd, &plan->tb.post_fwd, plan->owned)) != cudaSuccess ||
      (err = upload(pre_inv, &plan->tb.pre_inv, plan->owned)) != cudaSuccess ||
      (err = upload(post_inv, &plan->tb.post_inv, plan->owned)) != cudaSuccess ||
      (err = upload(tw_pass1, &plan->tb.tw_pass1, plan->owned)) != cudaSuccess ||
      (err = upload(tw_pass2, &plan->tb.tw_pass2, plan->owned)) != cudaSuccess) {
    free_all(plan->owned);
    delete plan;
    return cuda_fail(err, "uploading MDCT tables");
  }
  // float64 compute dtype (f64_kernels.cu): the unrounded tables
  {
    std::vector<double> cos64(static_cast<size_t>(8) * n);
    for (int m = 0; m < 8 * n; ++m) cos64[m] = std::cos(pi * m / (4.0 * n));
    plan->tb64.n = n;
    plan->tb64.scale_fwd = scale_fwd;
    plan->tb64.scale_inv = scale_inv;
    if ((err = upload(t.fold, &plan->tb64.fold, plan->owned)) != cudaSuccess ||
        (err = upload(t.unfold, &plan->tb64.unfold, plan->owned)) != cudaSuccess ||
        (err = upload(cos64, &plan->tb64.cos_table, plan->owned)) != cudaSuccess) {
      free_all(plan->owned);
      delete plan;
      return cuda_fail(err, "uploading the float64 MDCT tables");
    }
  }
  *out = plan;
  return AC_OK;
}

int ac_mdct_plan_destroy(ac_mdct_plan* plan) {
  if (plan == nullptr) return AC_OK;
  free_all(plan->owned);
  delete plan;
  return AC_OK;
}

// plans hold device tables: a call with another device current would hand the kernels foreign pointers
static int check_sizes_and_device(bool have_plan, int plan_device, int64_t batches, int64_t len, int channels) {
  if (!have_plan) return fail(AC_ERR_INVALID, "plan is null");
  if (batches < 0 || len < 0 || channels < 1) return fail(AC_ERR_INVALID, "negative size or channels < 1");
  int dev = -1;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return cuda_fail(err, "cudaGetDevice");
  if (dev != plan_device)
    return fail(AC_ERR_INVALID, "plan was created on cuda:%d but the current device is cuda:%d", plan_device, dev);
  return AC_OK;
}
#define check_common(plan, batches, len, channels) \
  check_sizes_and_device((plan) != nullptr, (plan) != nullptr ? (plan)->device : 0, batches, len, channels)

int ac_mdct_forward_f32(const ac_mdct_plan* plan, const float* x, float* y, int64_t batches, int64_t samples,
                        int channels, void* stream) {
  if (int rc = check_common(plan, batches, samples, channels)) return rc;
  const int n = plan->tb.n;
  if (samples % n != 0)
    return fail(AC_ERR_INVALID, "samples_n (%lld) must be a multiple of filters_n (%d)", (long long)samples, n);
  const int64_t blocks = samples / n;
  if (blocks + 1 > 2147483647LL / 2) return fail(AC_ERR_INVALID, "too many blocks per batch row");
  if (batches == 0) return AC_OK;
  if (x == nullptr && samples > 0) return fail(AC_ERR_INVALID, "x is null");
  if (y == nullptr) return fail(AC_ERR_INVALID, "y is null");
  if (!aligned16(x) || !aligned16(y)) return fail(AC_ERR_INVALID, "x and y must be 16-byte aligned");
  cudaError_t err = ac::mdct_forward(plan->tb, x, y, batches, blocks, channels, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "mdct_forward launch");
}

static int inverse_common(const ac_mdct_plan* plan, const float* y, const int32_t* q, const float* thr, float* x,
                          int64_t batches, int64_t blocks, int channels, void* stream) {
  if (int rc = check_common(plan, batches, blocks, channels)) return rc;
  if (blocks + 1 > 2147483647LL / 2) return fail(AC_ERR_INVALID, "too many blocks per batch row");
  if (batches == 0) return AC_OK;
  if (x == nullptr) return fail(AC_ERR_INVALID, "x is null");
  if (blocks > 0) {
    if (q == nullptr && y == nullptr) return fail(AC_ERR_INVALID, "y is null");
    if (q != nullptr && thr == nullptr) return fail(AC_ERR_INVALID, "thr is null");
  }
  if (!aligned16(x) || !aligned16(y) || !aligned16(q) || !aligned16(thr))
    return fail(AC_ERR_INVALID, "all tensors must be 16-byte aligned");
  cudaError_t err = ac::mdct_inverse(plan->tb, y, q, thr, x, batches, blocks, channels, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "mdct_inverse launch");
}

int ac_mdct_inverse_f32(const ac_mdct_plan* plan, const float* y, float* x, int64_t batches, int64_t blocks,
                        int channels, void* stream) {
  return inverse_common(plan, y, nullptr, nullptr, x, batches, blocks, channels, stream);
}

int ac_mdct_inverse_dequant_f32(const ac_mdct_plan* plan, const int32_t* q, const float* thr, float* x,
                                int64_t batches, int64_t blocks, int channels, void* stream) {
  if (q == nullptr && blocks > 0 && batches > 0) return fail(AC_ERR_INVALID, "q is null");
  return inverse_common(plan, nullptr, q, thr, x, batches, blocks, channels, stream);
}

int ac_mdct_inverse_dequant_compact_f32(const ac_mdct_plan* plan, const ac_pa_plan* pa_plan, const int32_t* q,
                                        const float* bark_thr, float thr_scale, float* x, int64_t batches, int64_t blocks,
                                        int channels, void* stream) {
  if (int rc = check_common(plan, batches, blocks, channels)) return rc;
  if (pa_plan == nullptr) return fail(AC_ERR_INVALID, "psychoacoustic plan is null");
  if (pa_plan->device != plan->device || pa_plan->tb.n != plan->tb.n)
    return fail(AC_ERR_INVALID, "the two plans must share the device and filter_bands_n");
  if (!(thr_scale > 0.f)) return fail(AC_ERR_INVALID, "thr_scale must be positive");
  if (blocks + 1 > 2147483647LL / 2) return fail(AC_ERR_INVALID, "too many blocks per batch row");
  if (batches == 0) return AC_OK;
  if (x == nullptr) return fail(AC_ERR_INVALID, "x is null");
  if (blocks > 0 && (q == nullptr || bark_thr == nullptr)) return fail(AC_ERR_INVALID, "null tensor");
  if (!aligned16(x) || !aligned16(q) || !aligned16(bark_thr)) return fail(AC_ERR_INVALID, "all tensors must be 16-byte aligned");
  const ac::PaDeviceTables& pt = pa_plan->tb;
  if (pt.nb != 64 || !pt.tile_ok || pt.filt4 == nullptr || !(channels == 1 || channels == 2) ||
      !(plan->tb.n == 256 || plan->tb.n == 512 || plan->tb.n == 1024))
    return fail(AC_ERR_UNSUPPORTED, "the fused compact decoder is built for bark_bands_n == 64, <= 3 bands per filter, "
                                    "filters_n 256 / 512 / 1024 and 1 or 2 channels (use ac_pa_expand_threshold_f32)");
  cudaError_t err = ac::mdct_inverse_compact_tile(plan->tb, q, bark_thr, pt.filt4, pt.eps * (thr_scale * thr_scale), x,
                                                  batches, blocks, channels, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "mdct_inverse_dequant_compact launch");
}

// ------------------------------------------------------------------------------------- psychoacoustics
int ac_pa_plan_create(double sample_rate, int filter_bands_n, int bark_bands_n, double alpha, ac_pa_plan** out) {
  return ac_pa_plan_create_ex(sample_rate, filter_bands_n, bark_bands_n, alpha, AC_DTYPE_F32, out);
}

int ac_pa_plan_create_ex(double sample_rate, int filter_bands_n, int bark_bands_n, double alpha_in, int compute_dtype,
                         ac_pa_plan** out) {
  if (out == nullptr) return fail(AC_ERR_INVALID, "out is null");
  *out = nullptr;
  if (compute_dtype != AC_DTYPE_F32 && compute_dtype != AC_DTYPE_BF16)
    return fail(AC_ERR_INVALID, "compute_dtype must be AC_DTYPE_F32 or AC_DTYPE_BF16 (float64 runs on any plan)");
  const bool bf16 = compute_dtype == AC_DTYPE_BF16;
  // compute_dtype=tf.bfloat16: python scalars take the tensor's dtype, so alpha and 1 / alpha enter tf.pow as bfloat16
  // values (psychoacoustic.py:197, 206, 208) - and 1 / alpha is rounded on its own, it is not the inverse of the rounded alpha
  const double alpha = bf16 ? bf16_round(alpha_in) : alpha_in;
  const double inv_alpha = bf16 ? bf16_round(1. / alpha_in) : 1. / alpha_in;
  if (!(sample_rate > 0) || filter_bands_n < 1 || bark_bands_n < 1 || !(alpha > 0))
    return fail(AC_ERR_INVALID, "invalid psychoacoustic parameters (sample_rate %g, filter_bands_n %d, bark_bands_n %d, alpha %g)",
                sample_rate, filter_bands_n, bark_bands_n, alpha);
  if (filter_bands_n > 16384 || bark_bands_n > 1024)
    return fail(AC_ERR_UNSUPPORTED, "filter_bands_n <= 16384 and bark_bands_n <= 1024 are built");
  ac_pa_plan* plan = new (std::nothrow) ac_pa_plan();
  if (plan == nullptr) return fail(AC_ERR_ALLOC, "out of host memory");
  cudaError_t err = cudaGetDevice(&plan->device);
  if (err != cudaSuccess) {
    delete plan;
    return cuda_fail(err, "cudaGetDevice (no CUDA device? this library has no CPU path)");
  }
  plan->host = ac::build_pa_tables(sample_rate, filter_bands_n, bark_bands_n, alpha_in);
  plan->compute_dtype = compute_dtype;
  if (bf16) {     // W, W_inv, the quiet threshold and the spreading matrix are cast to the compute dtype (:65-69),
    ac::PaTables& h = plan->host;   // tf.linspace runs in it (:187-189)
    for (double& v : h.w) v = bf16_round(v);
    for (double& v : h.w_inv) v = bf16_round(v);
    for (double& v : h.quiet) v = bf16_round(v);
    for (double& v : h.spread_fn) v = bf16_round(v);
    for (float& v : h.lin) v = static_cast<float>(bf16_round(v));
    for (float& v : h.band_w) v = static_cast<float>(bf16_round(v));
    for (float& v : h.filt_w) v = static_cast<float>(bf16_round(v));
  }
  const ac::PaTables& t = plan->host;
  ac::PaDeviceTables& d = plan->tb;
  d.n = t.n;
  d.nb = t.nb;
  d.alpha = static_cast<float>(alpha);
  d.neg_alpha = static_cast<float>(-alpha);
  d.inv_alpha = static_cast<float>(inv_alpha);
  d.eps = bf16 ? static_cast<float>(bf16_round(1e-14)) : 1e-14f;
  d.max_band_cnt = t.max_band_cnt;
  d.max_filt_cnt = t.max_filt_cnt;
  d.gain_log2 = static_cast<float>(-alpha * std::log2(10.0) / 10.0);
  // ---- tile-kernel tables: chunks of chunk_k filters; a chunk's band sums are cut into steps of four filters
  //      and the steps dealt to the four warps of a CTA in contiguous, equally long runs
  d.chunk_k = 128;
  d.n_chunks = (t.n + d.chunk_k - 1) / d.chunk_k;
  d.tile_ok = (t.nb == 64 && t.max_filt_cnt <= 3) ? 1 : 0;
  std::vector<int4> band_desc;
  std::vector<float> band_w4;
  std::vector<int32_t> desc_start(static_cast<size_t>(d.n_chunks) * 5 + 1, 0);
  for (int c = 0; c < d.n_chunks; ++c) {
    const int kc0 = c * d.chunk_k, kc1 = std::min(t.n, kc0 + d.chunk_k);
    const size_t first_desc = band_desc.size();
    std::vector<double> cost;
    for (int i = 0; i < t.nb; ++i) {
      const int ka = std::max(t.band_k0[i], kc0), kb = std::min(t.band_k0[i] + t.band_cnt[i], kc1);
      if (kb <= ka) continue;
      const int steps = (kb - ka + 3) / 4;
      const bool final = t.band_k0[i] + t.band_cnt[i] <= kc1;
      for (int st = 0; st < steps; ++st) {
        int4 ds;
        ds.x = ka - kc0 + 4 * st;                        // row of T
        ds.y = static_cast<int>(band_w4.size());         // four weights
        ds.z = i | (t.band_k0[i] < kc0 ? 0x100 : 0) | (final ? 0x200 : 0) | (st == 0 ? 0x400 : 0) |
               (st == steps - 1 ? 0x800 : 0);
        ds.w = 0;
        for (int u = 0; u < 4; ++u) {
          const int k = ka + 4 * st + u;
          band_w4.push_back(k < kb ? t.band_w[t.band_ptr[i] + (k - t.band_k0[i])] : 0.f);
        }
        band_desc.push_back(ds);
        cost.push_back(14.0 + (st == steps - 1 ? (final ? 30.0 : 6.0) : 0.0));   // ~instructions per lane
      }
    }
    // contiguous runs of about equal cost; a band's steps stay with one warp (the accumulator is a register)
    double total = 0;
    for (double v : cost) total += v;
    double run = 0;
    int w = 0;
    desc_start[static_cast<size_t>(c) * 5] = static_cast<int32_t>(first_desc);
    for (size_t sidx = 0; sidx < cost.size(); ++sidx) {
      run += cost[sidx];
      const bool band_end = (band_desc[first_desc + sidx].z & 0x800) != 0;
      while (band_end && w < 3 && run >= total * (w + 1) / 4.0)
        desc_start[static_cast<size_t>(c) * 5 + (++w)] = static_cast<int32_t>(first_desc + sidx + 1);
    }
    while (w < 4) desc_start[static_cast<size_t>(c) * 5 + (++w)] = static_cast<int32_t>(band_desc.size());
  }
  desc_start[static_cast<size_t>(d.n_chunks) * 5] = static_cast<int32_t>(band_desc.size());
  d.n_desc = static_cast<int>(band_desc.size());
  d.n_band_w4 = static_cast<int>(band_w4.size());
  // ---- tensor-core tile kernel: job list, tonality ranges and step weights (build_mma_jobs above)
  std::vector<float> mma_w4;
  if (build_mma_jobs(t, plan->jobs, mma_w4, &d.mma_chunk_k, &d.mma_n_chunks, &d.n_jobs)) d.jobs_host = &plan->jobs;
  d.n_mma_w4 = static_cast<int>(mma_w4.size());
  auto pow_table = [](float a) {
    std::vector<float2> tab(256);
    for (int e = 0; e < 256; ++e) {
      if (e == 255) {
        tab[e] = make_float2(INFINITY, 0.f);
        continue;
      }
      const double ae = static_cast<double>(a) * (e - 127);
      const double nr = std::nearbyint(ae);
      tab[e] = make_float2(static_cast<float>(std::ldexp(1.0, static_cast<int>(nr))), static_cast<float>(ae - nr));
    }
    return tab;
  };
  const std::vector<float2> pow_alpha = pow_table(d.alpha), pow_inv_alpha = pow_table(d.inv_alpha);
  // (acc 10^(-alpha offset / 10))^(1 / alpha): the offset's factor is alpha (1 / alpha) = 1, except with both rounded to bfloat16
  d.offset_log2 = static_cast<float>(-std::log2(10.0) / 10.0 * (bf16 ? alpha * inv_alpha : 1.0));
  d.pow_c1 = static_cast<float>(alpha - 0.5);
  d.pow_c2 = static_cast<float>(inv_alpha - 2.);
  // |c| lg2(x) keeps ~2^-24 / |c| of headroom against fp32 rounding of lg2: split powers for alpha in [0.45, 0.65]
  d.pow_split = (alpha >= 0.45 && alpha <= 0.65 && std::getenv("AC_PA_POW_TABLE") == nullptr) ? 1 : 0;
  {
    // max(eps, masking)^(1/alpha) <= eps when alpha <= 1; the result is then raised to the quiet threshold anyway
    float quiet_min = INFINITY;
    for (double qv : t.quiet) quiet_min = std::min(quiet_min, static_cast<float>(qv));
    d.clamp_needed = (alpha <= 1.0 && quiet_min >= d.eps) ? 0 : 1;
  }
  std::vector<float4> filt4(t.n, make_float4(0.f, 0.f, 0.f, 0.f));
  if (d.tile_ok) {
    for (int k = 0; k < t.n; ++k) {
      const int b0 = std::max(0, std::min(t.filt_b0[k], t.nb - 3));
      float w3[3] = {0.f, 0.f, 0.f};
      for (int s = 0; s < t.filt_cnt[k]; ++s) w3[t.filt_b0[k] + s - b0] = t.filt_w[t.filt_ptr[k] + s];
      float b0f;
      std::memcpy(&b0f, &b0, sizeof(float));
      filt4[k] = make_float4(w3[0], w3[1], w3[2], b0f);
      if (k / 32 < 64)
        for (int sl = 0; sl < 3; ++sl)
          if (w3[sl] != 0.f) d.filt_mask[k / 32] |= static_cast<uint8_t>(1u << sl);
    }
    float floor_min = INFINITY;      // the quiet threshold's share of every filter's intensity threshold, as the kernel sums it
    for (int k = 0; k < t.n; ++k) {
      const int b0 = __float_as_int_host(filt4[k].w);
      float v = 0.f;
      const float w3[3] = {filt4[k].x, filt4[k].y, filt4[k].z};
      for (int sl = 0; sl < 3; ++sl) v += w3[sl] * static_cast<float>(t.quiet[b0 + sl]);
      floor_min = std::min(floor_min, v);
    }
    d.thr_clamp_needed = (floor_min >= 4.0f * d.eps) ? 0 : 1;
  }
  std::vector<float> quiet(t.quiet.begin(), t.quiet.end()), spread(t.spread_fn.begin(), t.spread_fn.end());
  if ((err = upload(t.band_k0, &d.band_k0, plan->owned)) != cudaSuccess ||
      (err = upload(t.band_cnt, &d.band_cnt, plan->owned)) != cudaSuccess ||
      (err = upload(t.band_ptr, &d.band_ptr, plan->owned)) != cudaSuccess ||
      (err = upload(t.band_w, &d.band_w, plan->owned)) != cudaSuccess ||
      (err = upload(t.filt_b0, &d.filt_b0, plan->owned)) != cudaSuccess ||
      (err = upload(t.filt_cnt, &d.filt_cnt, plan->owned)) != cudaSuccess ||
      (err = upload(t.filt_ptr, &d.filt_ptr, plan->owned)) != cudaSuccess ||
      (err = upload(t.filt_w, &d.filt_w, plan->owned)) != cudaSuccess ||
      (err = upload(quiet, &d.quiet, plan->owned)) != cudaSuccess ||
      (err = upload(spread, &d.spread_fn, plan->owned)) != cudaSuccess ||
      (err = upload(t.lin, &d.lin, plan->owned)) != cudaSuccess ||
      (err = upload(band_desc, &d.band_desc, plan->owned)) != cudaSuccess ||
      (err = upload(band_w4, &d.band_w4, plan->owned)) != cudaSuccess ||
      (err = upload(desc_start, &d.desc_start, plan->owned)) != cudaSuccess ||
      (err = upload(filt4, &d.filt4, plan->owned)) != cudaSuccess ||
      (err = upload(mma_w4, &d.mma_w4, plan->owned)) != cudaSuccess ||
      (err = upload(pow_alpha, &d.pow_alpha, plan->owned)) != cudaSuccess ||
      (err = upload(pow_inv_alpha, &d.pow_inv_alpha, plan->owned)) != cudaSuccess) {
    free_all(plan->owned);
    delete plan;
    return cuda_fail(err, "uploading psychoacoustic tables");
  }
  {
    // ticket counters of the masking kernel's tile scheduler (zero between launches)
    const std::vector<unsigned> zeros(static_cast<size_t>(2) * ac::kPaSchedSlots, 0u);
    const unsigned* sched = nullptr;
    if ((err = upload(zeros, &sched, plan->owned)) != cudaSuccess) {
      free_all(plan->owned);
      delete plan;
      return cuda_fail(err, "allocating the tile scheduler counters");
    }
    d.sched = const_cast<unsigned*>(sched);
  }
  // float64 compute dtype (f64_kernels.cu): unrounded weights; the index arrays are shared with the fp32 plan
  {
    ac::PaDeviceTables64& e = plan->tb64;
    e.n = t.n;
    e.nb = t.nb;
    e.alpha = alpha_in;
    e.inv_alpha = 1. / alpha_in;
    e.eps = 1e-14;
    e.band_k0 = d.band_k0;
    e.band_cnt = d.band_cnt;
    e.band_ptr = d.band_ptr;
    e.filt_b0 = d.filt_b0;
    e.filt_cnt = d.filt_cnt;
    e.filt_ptr = d.filt_ptr;
    std::vector<double> band_w64, filt_w64, lin64(t.nb, 0.0);
    for (int i = 0; i < t.nb; ++i)
      for (int k = t.band_k0[i]; k < t.band_k0[i] + t.band_cnt[i]; ++k) band_w64.push_back(t.w[static_cast<size_t>(k) * t.nb + i]);
    for (int k = 0; k < t.n; ++k)
      for (int i = t.filt_b0[k]; i < t.filt_b0[k] + t.filt_cnt[k]; ++i) filt_w64.push_back(t.w_inv[static_cast<size_t>(i) * t.n + k]);
    if (t.nb > 1) {                                        // tf.linspace(0, max_bark, nb) in float64 (:187-189)
      const double delta = t.max_bark / (t.nb - 1);
      for (int j = 0; j < t.nb; ++j) lin64[j] = j * delta;
      lin64[t.nb - 1] = t.max_bark;
    }
    if ((err = upload(band_w64, &e.band_w, plan->owned)) != cudaSuccess ||
        (err = upload(filt_w64, &e.filt_w, plan->owned)) != cudaSuccess ||
        (err = upload(t.quiet, &e.quiet, plan->owned)) != cudaSuccess ||
        (err = upload(t.spread_fn, &e.spread_fn, plan->owned)) != cudaSuccess ||
        (err = upload(lin64, &e.lin, plan->owned)) != cudaSuccess) {
      free_all(plan->owned);
      delete plan;
      return cuda_fail(err, "uploading the float64 psychoacoustic tables");
    }
  }
  *out = plan;
  return AC_OK;
}

int ac_pa_plan_destroy(ac_pa_plan* plan) {
  if (plan == nullptr) return AC_OK;
  free_all(plan->owned);
  delete plan;
  return AC_OK;
}

int ac_pa_tonality_f32(const ac_pa_plan* plan, const float* y, float* ton, int64_t batches, int64_t blocks,
                       int channels, void* stream) {
  if (int rc = check_common(plan, batches, blocks, channels)) return rc;
  if (batches * blocks == 0) return AC_OK;
  if (y == nullptr || ton == nullptr) return fail(AC_ERR_INVALID, "null tensor");
  if (!aligned16(y)) return fail(AC_ERR_INVALID, "y must be 16-byte aligned");
  cudaError_t err = ac::pa_tonality(plan->tb, y, ton, batches * blocks, channels, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "pa_tonality launch");
}

int ac_pa_threshold_f32(const ac_pa_plan* plan, const float* y, const float* ton, float drown, float* thr,
                        int64_t batches, int64_t blocks, int channels, void* stream) {
  if (int rc = check_common(plan, batches, blocks, channels)) return rc;
  if (batches * blocks == 0) return AC_OK;
  if (y == nullptr || thr == nullptr) return fail(AC_ERR_INVALID, "null tensor");
  cudaError_t err = ac::pa_threshold(plan->tb, y, ton, drown, 1.0f, thr, nullptr, batches * blocks, channels,
                                     static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "pa_threshold launch");
}

int ac_pa_encode_f32(const ac_pa_plan* plan, const float* y, float drown, float thr_scale, float* thr_out, int32_t* q,
                     int64_t batches, int64_t blocks, int channels, void* stream) {
  if (int rc = check_common(plan, batches, blocks, channels)) return rc;
  if (!(thr_scale > 0.f)) return fail(AC_ERR_INVALID, "thr_scale must be positive");
  if (batches * blocks == 0) return AC_OK;
  if (y == nullptr || q == nullptr) return fail(AC_ERR_INVALID, "null tensor");
  cudaError_t err = ac::pa_threshold(plan->tb, y, nullptr, drown, thr_scale, thr_out, q, batches * blocks, channels,
                                     static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "pa_encode launch");
}

int ac_pa_encode_compact_f32(const ac_pa_plan* plan, const float* y, float drown, float thr_scale, float* bark_thr,
                             int32_t* q, int64_t batches, int64_t blocks, int channels, void* stream) {
  if (int rc = check_common(plan, batches, blocks, channels)) return rc;
  if (!(thr_scale > 0.f)) return fail(AC_ERR_INVALID, "thr_scale must be positive");
  if (batches * blocks == 0) return AC_OK;
  if (y == nullptr || q == nullptr || bark_thr == nullptr) return fail(AC_ERR_INVALID, "null tensor");
  cudaError_t err = ac::pa_encode_compact(plan->tb, y, drown, thr_scale, bark_thr, q, batches * blocks, channels,
                                          static_cast<cudaStream_t>(stream));
  if (err == cudaErrorNotSupported)
    return fail(AC_ERR_UNSUPPORTED, "compact side information needs bark_bands_n == 64, <= 3 bands per filter, 1/2/4 channels");
  if (err == cudaErrorMisalignedAddress) return fail(AC_ERR_INVALID, "compact side information needs 16-byte aligned tensors");
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "pa_encode_compact launch");
}

int ac_pa_expand_threshold_f32(const ac_pa_plan* plan, const float* bark_thr, float thr_scale, float* thr, int64_t batches,
                               int64_t blocks, int channels, void* stream) {
  if (int rc = check_common(plan, batches, blocks, channels)) return rc;
  if (!(thr_scale > 0.f)) return fail(AC_ERR_INVALID, "thr_scale must be positive");
  if (batches * blocks == 0) return AC_OK;
  if (bark_thr == nullptr || thr == nullptr) return fail(AC_ERR_INVALID, "null tensor");
  cudaError_t err = ac::pa_expand_threshold(plan->tb, bark_thr, thr_scale, thr, batches * blocks, channels,
                                            static_cast<cudaStream_t>(stream));
  if (err == cudaErrorNotSupported)
    return fail(AC_ERR_UNSUPPORTED, "compact side information needs bark_bands_n == 64, <= 3 bands per filter, 1/2/4 channels");
  if (err == cudaErrorMisalignedAddress) return fail(AC_ERR_INVALID, "compact side information needs 16-byte aligned tensors");
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "pa_expand_threshold launch");
}

// ----------------------------------------------------------------------------------- single-pass encoder
int64_t ac_codec_encode_workspace_bytes(const ac_mdct_plan* mdct, const ac_pa_plan* pa, int64_t batches, int64_t samples,
                                        int channels) {
  if (mdct == nullptr || pa == nullptr || batches < 0 || samples < 0 || channels < 1) return -1;
  const int n = mdct->tb.n;
  if (samples % n != 0) return -1;
  if (ac::pa_encode_fused_supported(pa->tb, mdct->tb, channels)) return 0;
  return batches * (samples / n + 1) * n * channels * static_cast<int64_t>(sizeof(float));
}

int ac_codec_encode_f32(const ac_mdct_plan* mdct, const ac_pa_plan* pa, const float* x, float drown, float thr_scale,
                        float* step_out, float* bark_thr_out, int32_t* q, int64_t batches, int64_t samples, int channels,
                        void* workspace, void* stream) {
  if (int rc = check_common(mdct, batches, samples, channels)) return rc;
  if (int rc = check_common(pa, batches, samples, channels)) return rc;
  const int n = mdct->tb.n;
  if (pa->tb.n != n) return fail(AC_ERR_INVALID, "the two plans have different filters_n (%d, %d)", n, pa->tb.n);
  if (samples % n != 0)
    return fail(AC_ERR_INVALID, "samples_n (%lld) must be a multiple of filters_n (%d)", (long long)samples, n);
  if (!(thr_scale > 0.f)) return fail(AC_ERR_INVALID, "thr_scale must be positive");
  const int64_t blocks = samples / n;
  if (blocks + 1 > 2147483647LL / 2) return fail(AC_ERR_INVALID, "too many blocks per batch row");
  if (batches == 0) return AC_OK;
  if ((x == nullptr && samples > 0) || q == nullptr) return fail(AC_ERR_INVALID, "null tensor");
  if (!aligned16(x) || !aligned16(q) || !aligned16(step_out) || !aligned16(bark_thr_out) || !aligned16(workspace))
    return fail(AC_ERR_INVALID, "all tensors must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (ac::pa_encode_fused_supported(pa->tb, mdct->tb, channels)) {
    cudaError_t err = ac::pa_encode_fused(pa->tb, mdct->tb, x, drown, thr_scale, step_out, bark_thr_out, q, batches, blocks,
                                          channels, st);
    return err == cudaSuccess ? AC_OK : cuda_fail(err, "fused encoder launch");
  }
  // other shapes: transform into the caller's workspace, then the masking / quantising kernel
  if (workspace == nullptr) return fail(AC_ERR_INVALID, "this shape needs a workspace (ac_codec_encode_workspace_bytes)");
  float* y = static_cast<float*>(workspace);
  cudaError_t err = ac::mdct_forward(mdct->tb, x, y, batches, blocks, channels, st);
  if (err != cudaSuccess) return cuda_fail(err, "mdct_forward launch");
  const int64_t rows = batches * (blocks + 1);
  if (bark_thr_out != nullptr) {
    if (!ac::pa_mma_tile_supported(pa->tb, channels))
      return fail(AC_ERR_UNSUPPORTED, "compact side information needs bark_bands_n == 64, <= 3 bands per filter, 1/2/4 channels");
    ac::PaDeviceTables with_out = pa->tb;
    with_out.bark_out = bark_thr_out;
    err = ac::pa_threshold_mma_tile(with_out, y, nullptr, static_cast<float>(1.0 - static_cast<double>(drown)), thr_scale,
                                    step_out, q, rows, channels, st);
  } else {
    err = ac::pa_threshold(pa->tb, y, nullptr, drown, thr_scale, step_out, q, rows, channels, st);
  }
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "pa_encode launch");
}

int ac_pa_add_noise_f32(const float* y, const float* thr, float* out, int64_t n, uint64_t seed, void* stream) {
  if (n < 0) return fail(AC_ERR_INVALID, "negative size");
  if (n == 0) return AC_OK;
  if (y == nullptr || thr == nullptr || out == nullptr) return fail(AC_ERR_INVALID, "null tensor");
  cudaError_t err = ac::add_noise(y, thr, out, n, seed, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "add_noise launch");
}

int ac_quantize_f32(const float* y, const float* thr, int32_t* q, int64_t n, void* stream) {
  if (n < 0) return fail(AC_ERR_INVALID, "negative size");
  if (n == 0) return AC_OK;
  if (y == nullptr || thr == nullptr || q == nullptr) return fail(AC_ERR_INVALID, "null tensor");
  cudaError_t err = ac::quantize(y, thr, q, n, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "quantize launch");
}

int ac_dequantize_f32(const int32_t* q, const float* thr, float* y, int64_t n, void* stream) {
  if (n < 0) return fail(AC_ERR_INVALID, "negative size");
  if (n == 0) return AC_OK;
  if (y == nullptr || thr == nullptr || q == nullptr) return fail(AC_ERR_INVALID, "null tensor");
  cudaError_t err = ac::dequantize(q, thr, y, n, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "dequantize launch");
}

// --------------------------------------------------------------------------------- host-buffer streaming
// A ring of kRingSlots chunk-sized device buffers for x and x_hat: the device footprint is a few chunks whatever the
// size of the host batch (cfg5 streams through one or two GPUs), the H2D stream runs up to kRingSlots chunks ahead of
// the kernels and the D2H stream trails them.  Amplitudes, steps and integers need one chunk-sized set because the
// kernels of all chunks run in order on one stream.
struct ac_codec_pipeline {
  static constexpr int kRingSlots = 4;
  const ac_mdct_plan* mdct = nullptr;
  const ac_pa_plan* pa = nullptr;
  int64_t chunk_clips = 0, samples = 0;
  int channels = 0, device = 0;
  cudaStream_t h2d = nullptr, run = nullptr, d2h = nullptr;
  cudaEvent_t entry = nullptr, done = nullptr, stats_ready = nullptr;
  float *y = nullptr, *step = nullptr;        // one chunk
  int32_t* q = nullptr;
  float *x_ring = nullptr, *xhat_ring = nullptr;   // kRingSlots chunks each
  // per slot: x has landed (h2d), the forward MDCT has read x (run), x_hat is complete (run), x_hat is on the host (d2h)
  cudaEvent_t x_ready[kRingSlots] = {}, x_free[kRingSlots] = {}, out_ready[kRingSlots] = {}, out_free[kRingSlots] = {};
  unsigned long long* stats_dev = nullptr;   // [3] coefficients, non-zeros, fixed-point sum of log2(2|q|+1)
};

int ac_codec_pipeline_destroy(ac_codec_pipeline* p) {
  if (p == nullptr) return AC_OK;
  cudaFree(p->y);
  cudaFree(p->step);
  cudaFree(p->q);
  cudaFree(p->x_ring);
  cudaFree(p->xhat_ring);
  cudaFree(p->stats_dev);
  for (int i = 0; i < ac_codec_pipeline::kRingSlots; ++i)
    for (cudaEvent_t e : {p->x_ready[i], p->x_free[i], p->out_ready[i], p->out_free[i]})
      if (e) cudaEventDestroy(e);
  if (p->entry) cudaEventDestroy(p->entry);
  if (p->done) cudaEventDestroy(p->done);
  if (p->stats_ready) cudaEventDestroy(p->stats_ready);
  if (p->h2d) cudaStreamDestroy(p->h2d);
  if (p->run) cudaStreamDestroy(p->run);
  if (p->d2h) cudaStreamDestroy(p->d2h);
  delete p;
  return AC_OK;
}

int ac_codec_pipeline_create(const ac_mdct_plan* mdct, const ac_pa_plan* pa, int64_t chunk_clips, int64_t samples,
                             int channels, ac_codec_pipeline** out) {
  if (out == nullptr) return fail(AC_ERR_INVALID, "out is null");
  *out = nullptr;
  if (mdct == nullptr || pa == nullptr) return fail(AC_ERR_INVALID, "plan is null");
  if (chunk_clips < 1 || samples < 0 || channels < 1) return fail(AC_ERR_INVALID, "invalid pipeline shape");
  const int n = mdct->tb.n;
  if (pa->tb.n != n) return fail(AC_ERR_INVALID, "filter_bands_n (%d) != filters_n (%d)", pa->tb.n, n);
  if (samples % n != 0)
    return fail(AC_ERR_INVALID, "samples_n (%lld) must be a multiple of filters_n (%d)", (long long)samples, n);
  ac_codec_pipeline* p = new (std::nothrow) ac_codec_pipeline();
  if (p == nullptr) return fail(AC_ERR_ALLOC, "out of host memory");
  p->mdct = mdct;
  p->pa = pa;
  p->chunk_clips = chunk_clips;
  p->samples = samples;
  p->channels = channels;
  const size_t frames = static_cast<size_t>(samples / n + 1);
  const size_t amp_elems = std::max<size_t>(static_cast<size_t>(chunk_clips) * frames * n * channels, 4);
  const size_t in_elems = std::max<size_t>(static_cast<size_t>(chunk_clips) * samples * channels, 4);
  const size_t out_elems = static_cast<size_t>(chunk_clips) * (frames + 1) * n * channels;
  cudaError_t err = cudaGetDevice(&p->device);
  auto ok = [&](cudaError_t e) {
    if (err == cudaSuccess) err = e;
    return err == cudaSuccess;
  };
  ok(cudaStreamCreateWithFlags(&p->h2d, cudaStreamNonBlocking));
  ok(cudaStreamCreateWithFlags(&p->run, cudaStreamNonBlocking));
  ok(cudaStreamCreateWithFlags(&p->d2h, cudaStreamNonBlocking));
  ok(cudaEventCreateWithFlags(&p->entry, cudaEventDisableTiming));
  ok(cudaEventCreateWithFlags(&p->done, cudaEventDisableTiming));
  ok(cudaEventCreateWithFlags(&p->stats_ready, cudaEventDisableTiming));
  for (int i = 0; i < ac_codec_pipeline::kRingSlots; ++i) {
    ok(cudaEventCreateWithFlags(&p->x_ready[i], cudaEventDisableTiming));
    ok(cudaEventCreateWithFlags(&p->x_free[i], cudaEventDisableTiming));
    ok(cudaEventCreateWithFlags(&p->out_ready[i], cudaEventDisableTiming));
    ok(cudaEventCreateWithFlags(&p->out_free[i], cudaEventDisableTiming));
  }
  ok(cudaMalloc(reinterpret_cast<void**>(&p->stats_dev), 3 * sizeof(unsigned long long)));
  ok(cudaMalloc(reinterpret_cast<void**>(&p->y), amp_elems * sizeof(float)));
  ok(cudaMalloc(reinterpret_cast<void**>(&p->step), amp_elems * sizeof(float)));
  ok(cudaMalloc(reinterpret_cast<void**>(&p->q), amp_elems * sizeof(int32_t)));
  ok(cudaMalloc(reinterpret_cast<void**>(&p->x_ring), ac_codec_pipeline::kRingSlots * in_elems * sizeof(float)));
  ok(cudaMalloc(reinterpret_cast<void**>(&p->xhat_ring), ac_codec_pipeline::kRingSlots * out_elems * sizeof(float)));
  if (err != cudaSuccess) {
    ac_codec_pipeline_destroy(p);
    return cuda_fail(err, "creating the streaming pipeline");
  }
  *out = p;
  return AC_OK;
}

int ac_codec_roundtrip_host_f32(ac_codec_pipeline* p, const float* x_host, float* xhat_host, int64_t batches, float drown,
                                float thr_scale, double* stats, void* stream) {
  if (p == nullptr) return fail(AC_ERR_INVALID, "pipeline is null");
  if (batches < 0) return fail(AC_ERR_INVALID, "negative batch");
  if (!(thr_scale > 0.f)) return fail(AC_ERR_INVALID, "thr_scale must be positive");
  if (batches > 0 && (x_host == nullptr || xhat_host == nullptr)) return fail(AC_ERR_INVALID, "null host buffer");
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev != p->device)
    return fail(AC_ERR_INVALID, "pipeline was created on cuda:%d but the current device is cuda:%d", p->device, dev);
  constexpr int K = ac_codec_pipeline::kRingSlots;
  const int n = p->mdct->tb.n, c = p->channels;
  const int64_t s = p->samples, blocks = s / n, frames = blocks + 1;
  const size_t in_clip = static_cast<size_t>(s) * c, out_clip = static_cast<size_t>(frames + 1) * n * c;
  const size_t in_slot = std::max<size_t>(static_cast<size_t>(p->chunk_clips) * in_clip, 4);
  const size_t out_slot = static_cast<size_t>(p->chunk_clips) * out_clip;
  cudaError_t err = cudaSuccess;
  auto ok = [&](cudaError_t e) {
    if (err == cudaSuccess) err = e;
    return err == cudaSuccess;
  };
  // chunk schedule: the D2H copy of a chunk can only start when its H2D copy and its kernels are over, so the D2H stream
  // trails the H2D stream by one chunk, idles whenever the chunks grow, and the call ends one (last) chunk after the last
  // H2D copy.  A linear ramp (1, 2, ..., 8) / 8 of chunk_clips at the start (the first D2H starts after one clip, every growth step
  // costs one clip of idle D2H), full chunks in the middle, a ramp down in steps of two at the end: measured best of
  // fifteen schedules on cfg2 (tools/e2e_schedules.py: 5.25 ms against 5.45 for the doubling ramp, 5.45 for uniform chunks)
  std::vector<int64_t> chunk_of;
  {
    const int64_t S = p->chunk_clips, step = std::max<int64_t>(1, S / 8);     // eight ramp steps whatever the clip length
    std::vector<int64_t> up, down;
    int64_t up_sum = 0, down_sum = 0;
    for (int64_t v = step; v < S; v += step) {
      up.push_back(v);
      up_sum += v;
    }
    for (int64_t v = S - 2 * step; v >= step; v -= 2 * step) {
      down.push_back(v);
      down_sum += v;
    }
    int64_t left = batches;
    if (left <= up_sum + down_sum) {        // a short batch: as much of the ramp up as fits, the rest in one piece
      for (int64_t v : up) {
        if (left <= 0) break;
        const int64_t take = std::min(v, left);
        chunk_of.push_back(take);
        left -= take;
      }
      while (left > 0) {
        const int64_t take = std::min(S, left);
        chunk_of.push_back(take);
        left -= take;
      }
    } else {
      chunk_of = up;
      left -= up_sum + down_sum;
      while (left > 0) {
        const int64_t take = std::min(S, left);
        chunk_of.push_back(take);
        left -= take;
      }
      chunk_of.insert(chunk_of.end(), down.begin(), down.end());
    }
    // AC_PIPE_SCHEDULE="c0,c1,..." (development): explicit chunk sizes, the last one repeated, each at most chunk_clips
    if (const char* e = std::getenv("AC_PIPE_SCHEDULE")) {
      std::vector<int64_t> forced;
      int64_t rest = batches, last = 1;
      const char* q = e;
      while (rest > 0) {
        if (*q) {
          char* end = nullptr;
          const long v = std::strtol(q, &end, 10);
          if (end == q) break;
          last = std::max<int64_t>(1, std::min<int64_t>(v, p->chunk_clips));
          q = (*end == ',') ? end + 1 : end;
        }
        const int64_t take = std::min(last, rest);
        forced.push_back(take);
        rest -= take;
      }
      if (rest == 0 && !forced.empty()) chunk_of = forced;
    }
  }
  const size_t n_chunks = chunk_of.size();
  cudaStream_t caller = static_cast<cudaStream_t>(stream);
  ok(cudaEventRecord(p->entry, caller));
  ok(cudaStreamWaitEvent(p->h2d, p->entry, 0));
  ok(cudaStreamWaitEvent(p->run, p->entry, 0));
  ok(cudaStreamWaitEvent(p->d2h, p->entry, 0));
  if (stats != nullptr) ok(cudaMemsetAsync(p->stats_dev, 0, 3 * sizeof(unsigned long long), p->run));
  // Per chunk k (slot k mod K), enqueued in this order so that every event is recorded before it is waited for:
  //   h2d: wait until the forward MDCT of chunk k - K has read the slot's x      -> copy x      -> x_ready
  //   run: wait for x_ready and until x_hat of chunk k - K has left the slot     -> the kernels -> x_free, out_ready
  //   d2h: wait for out_ready                                                    -> copy x_hat   -> out_free
  // The host thread runs ahead of the GPU, so the copy engines see up to K chunks of work queued behind the kernels.
  // (AC_PIPE_HOST_GATE gates the re-use of a ring slot with cudaEventSynchronize on the host instead of the two stream
  // waits: measured equal on B200, tools/e2e_schedules.py.)
  // AC_PIPE_TRACE=1 (development): timing events around every stage of every chunk, printed relative to the entry
  std::vector<cudaEvent_t> tr;
  const bool trace = std::getenv("AC_PIPE_TRACE") != nullptr;
  const bool stream_gate = std::getenv("AC_PIPE_HOST_GATE") == nullptr;     // development: A/B of the two gates
  auto mark = [&](cudaStream_t st) {
    if (!trace) return;
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    tr.push_back(e);
  };
  mark(p->h2d);
  size_t k = 0;
  for (int64_t i = 0; k < n_chunks && err == cudaSuccess; i += chunk_of[k], ++k) {
    const int64_t cb = chunk_of[k];
    const int slot = static_cast<int>(k % K);
    float* xs = p->x_ring + slot * in_slot;
    float* xh = p->xhat_ring + slot * out_slot;
    if (k >= K) {
      if (stream_gate) {
        ok(cudaStreamWaitEvent(p->h2d, p->x_free[slot], 0));
        ok(cudaStreamWaitEvent(p->run, p->out_free[slot], 0));
      } else {
        ok(cudaEventSynchronize(p->x_free[slot]));
        ok(cudaEventSynchronize(p->out_free[slot]));
      }
    }
    mark(p->h2d);
    if (in_clip > 0)
      ok(cudaMemcpyAsync(xs, x_host + i * in_clip, cb * in_clip * sizeof(float), cudaMemcpyHostToDevice, p->h2d));
    mark(p->h2d);
    ok(cudaEventRecord(p->x_ready[slot], p->h2d));
    ok(cudaStreamWaitEvent(p->run, p->x_ready[slot], 0));
    mark(p->run);
    ok(ac::mdct_forward(p->mdct->tb, xs, p->y, cb, blocks, c, p->run));
    ok(cudaEventRecord(p->x_free[slot], p->run));
    ok(ac::pa_threshold(p->pa->tb, p->y, nullptr, drown, thr_scale, p->step, p->q, cb * frames, c, p->run));
    if (stats != nullptr) ok(ac::codec_stats(p->q, cb * frames * n * c, p->stats_dev, p->run));
    ok(ac::mdct_inverse(p->mdct->tb, nullptr, p->q, p->step, xh, cb, frames, c, p->run));
    mark(p->run);
    ok(cudaEventRecord(p->out_ready[slot], p->run));
    ok(cudaStreamWaitEvent(p->d2h, p->out_ready[slot], 0));
    mark(p->d2h);
    ok(cudaMemcpyAsync(xhat_host + i * out_clip, xh, cb * out_clip * sizeof(float), cudaMemcpyDeviceToHost, p->d2h));
    mark(p->d2h);
    ok(cudaEventRecord(p->out_free[slot], p->d2h));
  }
  unsigned long long host_stats[3] = {0, 0, 0};
  if (stats != nullptr) {
    ok(cudaEventRecord(p->stats_ready, p->run));      // also orders the memset before the copy when there are no chunks
    ok(cudaStreamWaitEvent(p->d2h, p->stats_ready, 0));
    ok(cudaMemcpyAsync(host_stats, p->stats_dev, sizeof(host_stats), cudaMemcpyDeviceToHost, p->d2h));
  }
  ok(cudaEventRecord(p->done, p->d2h));
  ok(cudaStreamWaitEvent(caller, p->done, 0));
  ok(cudaEventSynchronize(p->done));          // the result is host memory: hand it back complete
  if (trace) {
    cudaDeviceSynchronize();
    std::fprintf(stderr, "chunk clips | h2d start end | kernels start end | d2h start end   (us after the first H2D was enqueued)\n");
    for (size_t j = 0; j < n_chunks && 1 + 6 * j + 5 < tr.size(); ++j) {
      float t[6];
      for (int u = 0; u < 6; ++u) cudaEventElapsedTime(&t[u], tr[0], tr[1 + 6 * j + u]);
      std::fprintf(stderr, "%3zu %3lld | %8.1f %8.1f | %8.1f %8.1f | %8.1f %8.1f\n", j, (long long)chunk_of[j], 1e3 * t[0], 1e3 * t[1],
                   1e3 * t[2], 1e3 * t[3], 1e3 * t[4], 1e3 * t[5]);
    }
    for (cudaEvent_t e : tr) cudaEventDestroy(e);
  }
  if (err != cudaSuccess) return cuda_fail(err, "streaming round trip");
  if (stats != nullptr) {
    stats[0] = static_cast<double>(host_stats[0]);
    stats[1] = static_cast<double>(host_stats[1]);
    stats[2] = static_cast<double>(host_stats[2]) / 65536.0;
  }
  return AC_OK;
}

// ------------------------------------------------------------------------------------- bitstream statistics
int ac_codec_stats_i32(const int32_t* q, int64_t n, uint64_t* stats_dev, void* stream) {
  if (n < 0) return fail(AC_ERR_INVALID, "negative size");
  if (stats_dev == nullptr || (n > 0 && q == nullptr)) return fail(AC_ERR_INVALID, "null tensor");
  static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "64-bit counters");
  cudaError_t err = ac::codec_stats(q, n, reinterpret_cast<unsigned long long*>(stats_dev), static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "codec_stats launch");
}

// ------------------------------------------------------------------------------------ float64 compute dtype
int ac_mdct_forward_f64(const ac_mdct_plan* plan, const double* x, double* y, int64_t batches, int64_t samples,
                        int channels, void* stream) {
  if (int rc = check_common(plan, batches, samples, channels)) return rc;
  const int n = plan->tb64.n;
  if (samples % n != 0)
    return fail(AC_ERR_INVALID, "samples_n (%lld) must be a multiple of filters_n (%d)", (long long)samples, n);
  if (batches == 0) return AC_OK;
  if ((x == nullptr && samples > 0) || y == nullptr) return fail(AC_ERR_INVALID, "null tensor");
  cudaError_t err = ac::mdct_forward_f64(plan->tb64, x, y, batches, samples / n, channels, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "mdct_forward_f64 launch");
}

int ac_mdct_inverse_f64(const ac_mdct_plan* plan, const double* y, double* x, int64_t batches, int64_t blocks,
                        int channels, void* stream) {
  if (int rc = check_common(plan, batches, blocks, channels)) return rc;
  if (batches == 0) return AC_OK;
  if ((y == nullptr && blocks > 0) || x == nullptr) return fail(AC_ERR_INVALID, "null tensor");
  cudaError_t err = ac::mdct_inverse_f64(plan->tb64, y, x, batches, blocks, channels, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "mdct_inverse_f64 launch");
}

int ac_pa_tonality_f64(const ac_pa_plan* plan, const double* y, double* ton, int64_t batches, int64_t blocks,
                       int channels, void* stream) {
  if (int rc = check_common(plan, batches, blocks, channels)) return rc;
  if (batches * blocks == 0) return AC_OK;
  if (y == nullptr || ton == nullptr) return fail(AC_ERR_INVALID, "null tensor");
  cudaError_t err = ac::pa_tonality_f64(plan->tb64, y, ton, batches * blocks, channels, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "pa_tonality_f64 launch");
}

int ac_pa_threshold_f64(const ac_pa_plan* plan, const double* y, const double* ton, double drown, double* thr,
                        int64_t batches, int64_t blocks, int channels, void* stream) {
  if (int rc = check_common(plan, batches, blocks, channels)) return rc;
  if (batches * blocks == 0) return AC_OK;
  if (y == nullptr || thr == nullptr) return fail(AC_ERR_INVALID, "null tensor");
  cudaError_t err = ac::pa_threshold_f64(plan->tb64, y, ton, drown, thr, batches * blocks, channels,
                                         static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "pa_threshold_f64 launch");
}

int ac_quantize_f64(const double* y, const double* thr, int32_t* q, int64_t n, void* stream) {
  if (n < 0) return fail(AC_ERR_INVALID, "negative size");
  if (n > 0 && (y == nullptr || thr == nullptr || q == nullptr)) return fail(AC_ERR_INVALID, "null tensor");
  cudaError_t err = ac::quantize_f64(y, thr, q, n, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "quantize_f64 launch");
}

int ac_dequantize_f64(const int32_t* q, const double* thr, double* y, int64_t n, void* stream) {
  if (n < 0) return fail(AC_ERR_INVALID, "negative size");
  if (n > 0 && (y == nullptr || thr == nullptr || q == nullptr)) return fail(AC_ERR_INVALID, "null tensor");
  cudaError_t err = ac::dequantize_f64(q, thr, y, n, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "dequantize_f64 launch");
}

// ------------------------------------------------------------------------------------------- backward pass
int ac_pa_tonality_backward_f32(const ac_pa_plan* plan, const float* y, const float* grad_ton, float* grad_y,
                                int64_t batches, int64_t blocks, int channels, void* stream) {
  if (int rc = check_common(plan, batches, blocks, channels)) return rc;
  if (batches * blocks == 0) return AC_OK;
  if (y == nullptr || grad_ton == nullptr || grad_y == nullptr) return fail(AC_ERR_INVALID, "null tensor");
  cudaError_t err = ac::pa_tonality_backward(plan->tb, y, grad_ton, grad_y, batches * blocks, channels,
                                             static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "pa_tonality_backward launch");
}

int ac_pa_threshold_backward_f32(const ac_pa_plan* plan, const float* y, const float* ton, float drown, const float* grad_thr,
                                 float* grad_y, float* grad_ton, int64_t batches, int64_t blocks, int channels, void* stream) {
  if (int rc = check_common(plan, batches, blocks, channels)) return rc;
  if (batches * blocks == 0) return AC_OK;
  if (y == nullptr || ton == nullptr || grad_thr == nullptr || grad_y == nullptr) return fail(AC_ERR_INVALID, "null tensor");
  cudaError_t err = ac::pa_threshold_backward(plan->tb, y, ton, drown, grad_thr, grad_y, grad_ton, batches * blocks, channels,
                                              static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "pa_threshold_backward launch");
}

// ----------------------------------------------------------------------------------- entropy-coded bitstream
static int entropy_args(int64_t rows, int64_t row_len) {
  if (rows < 0 || row_len < 16 || row_len % 16 != 0 || row_len > (1 << 24))
    return fail(AC_ERR_INVALID, "rows >= 0 and row_len a multiple of 16 in [16, 2^24] expected (got %lld rows of %lld)",
                (long long)rows, (long long)row_len);
  return AC_OK;
}

int ac_entropy_plan_i32(const int32_t* q, int64_t rows, int64_t row_len, int64_t* offsets, void* stream) {
  if (int rc = entropy_args(rows, row_len)) return rc;
  if (offsets == nullptr || (q == nullptr && rows > 0)) return fail(AC_ERR_INVALID, "null tensor");
  cudaError_t err = ac::entropy_plan(q, rows, static_cast<int>(row_len), offsets, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "entropy_plan launch");
}

int ac_entropy_encode_i32(const int32_t* q, int64_t rows, int64_t row_len, const int64_t* offsets, uint8_t* bytes, void* stream) {
  if (int rc = entropy_args(rows, row_len)) return rc;
  if (rows == 0) return AC_OK;
  if (q == nullptr || offsets == nullptr || bytes == nullptr) return fail(AC_ERR_INVALID, "null tensor");
  if (reinterpret_cast<uintptr_t>(bytes) & 3) return fail(AC_ERR_INVALID, "the stream buffer must be 4-byte aligned");
  cudaError_t err = ac::entropy_encode(q, rows, static_cast<int>(row_len), offsets, bytes, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "entropy_encode launch");
}

int ac_entropy_decode_i32(const uint8_t* bytes, const int64_t* offsets, int64_t rows, int64_t row_len, int32_t* q, void* stream) {
  if (int rc = entropy_args(rows, row_len)) return rc;
  if (rows == 0) return AC_OK;
  if (q == nullptr || offsets == nullptr || bytes == nullptr) return fail(AC_ERR_INVALID, "null tensor");
  if (reinterpret_cast<uintptr_t>(bytes) & 3) return fail(AC_ERR_INVALID, "the stream buffer must be 4-byte aligned");
  cudaError_t err = ac::entropy_decode(bytes, offsets, rows, static_cast<int>(row_len), q, static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "entropy_decode launch");
}

// ------------------------------------------------------------------------------- bfloat16 compute dtype
// Tensors are bfloat16 at the boundary, the plan's tables and constants are bfloat16 values (ac_*_plan_create_ex with
// AC_DTYPE_BF16), the kernels run in float32 on float32 copies held in the caller's workspace: the reference's rule for
// the DCT (mdctransformer.py:326-344) applied to the whole path.  workspace: ac_bf16_workspace_bytes(in, out) bytes.
static inline int64_t pad4(int64_t n) { return (n + 3) & ~static_cast<int64_t>(3); }

int64_t ac_bf16_workspace_bytes(int64_t in_elems, int64_t out_elems) {
  if (in_elems < 0 || out_elems < 0) return -1;
  return 4 * (pad4(in_elems) + pad4(out_elems));
}

#define AC_REQUIRE_BF16(plan)                                                                                       \
  if ((plan)->compute_dtype != AC_DTYPE_BF16)                                                                       \
    return fail(AC_ERR_INVALID, "the plan holds float32 tables: create it with ac_*_plan_create_ex(..., AC_DTYPE_BF16)")

int ac_mdct_forward_bf16(const ac_mdct_plan* plan, const void* x, void* y, int64_t batches, int64_t samples, int channels,
                         void* workspace, void* stream) {
  if (int rc = check_common(plan, batches, samples, channels)) return rc;
  AC_REQUIRE_BF16(plan);
  const int n = plan->tb.n;
  if (samples % n != 0)
    return fail(AC_ERR_INVALID, "samples_n (%lld) must be a multiple of filters_n (%d)", (long long)samples, n);
  const int64_t blocks = samples / n, in = batches * samples * channels, out = batches * (blocks + 1) * n * channels;
  if (blocks + 1 > 2147483647LL / 2) return fail(AC_ERR_INVALID, "too many blocks per batch row");
  if (batches == 0) return AC_OK;
  if ((x == nullptr && samples > 0) || y == nullptr || workspace == nullptr) return fail(AC_ERR_INVALID, "null tensor");
  if (!aligned16(workspace)) return fail(AC_ERR_INVALID, "workspace must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* xf = static_cast<float*>(workspace);
  float* yf = xf + pad4(in);
  cudaError_t err = ac::bf16_to_f32(x, xf, in, st);
  if (err == cudaSuccess) err = ac::mdct_forward(plan->tb, xf, yf, batches, blocks, channels, st);
  if (err == cudaSuccess) err = ac::f32_to_bf16(yf, y, out, st);
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "mdct_forward (bfloat16) launch");
}

int ac_mdct_inverse_bf16(const ac_mdct_plan* plan, const void* y, void* x, int64_t batches, int64_t blocks, int channels,
                         void* workspace, void* stream) {
  if (int rc = check_common(plan, batches, blocks, channels)) return rc;
  AC_REQUIRE_BF16(plan);
  const int n = plan->tb.n;
  const int64_t in = batches * blocks * n * channels, out = batches * (blocks + 1) * n * channels;
  if (blocks + 1 > 2147483647LL / 2) return fail(AC_ERR_INVALID, "too many blocks per batch row");
  if (batches == 0) return AC_OK;
  if ((y == nullptr && blocks > 0) || x == nullptr || workspace == nullptr) return fail(AC_ERR_INVALID, "null tensor");
  if (!aligned16(workspace)) return fail(AC_ERR_INVALID, "workspace must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* yf = static_cast<float*>(workspace);
  float* xf = yf + pad4(in);
  cudaError_t err = ac::bf16_to_f32(y, yf, in, st);
  if (err == cudaSuccess) err = ac::mdct_inverse(plan->tb, yf, nullptr, nullptr, xf, batches, blocks, channels, st);
  if (err == cudaSuccess) err = ac::f32_to_bf16(xf, x, out, st);
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "mdct_inverse (bfloat16) launch");
}

int ac_pa_tonality_bf16(const ac_pa_plan* plan, const void* y, void* ton, int64_t batches, int64_t blocks, int channels,
                        void* workspace, void* stream) {
  if (int rc = check_common(plan, batches, blocks, channels)) return rc;
  AC_REQUIRE_BF16(plan);
  const int64_t rows = batches * blocks, in = rows * plan->tb.n * channels, out = rows * channels;
  if (rows == 0) return AC_OK;
  if (y == nullptr || ton == nullptr || workspace == nullptr) return fail(AC_ERR_INVALID, "null tensor");
  if (!aligned16(workspace)) return fail(AC_ERR_INVALID, "workspace must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* yf = static_cast<float*>(workspace);
  float* tf = yf + pad4(in);
  cudaError_t err = ac::bf16_to_f32(y, yf, in, st);
  if (err == cudaSuccess) err = ac::pa_tonality(plan->tb, yf, tf, rows, channels, st);
  if (err == cudaSuccess) err = ac::f32_to_bf16(tf, ton, out, st);
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "pa_tonality (bfloat16) launch");
}

int ac_pa_threshold_bf16(const ac_pa_plan* plan, const void* y, const void* ton, float drown, void* thr, int64_t batches,
                         int64_t blocks, int channels, void* workspace, void* stream) {
  if (int rc = check_common(plan, batches, blocks, channels)) return rc;
  AC_REQUIRE_BF16(plan);
  const int64_t rows = batches * blocks, amp = rows * plan->tb.n * channels, tn = ton != nullptr ? rows * channels : 0;
  if (rows == 0) return AC_OK;
  if (y == nullptr || thr == nullptr || workspace == nullptr) return fail(AC_ERR_INVALID, "null tensor");
  if (!aligned16(workspace)) return fail(AC_ERR_INVALID, "workspace must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // workspace: [amplitudes | tonality] as the inputs, [threshold] as the output:
  // ac_bf16_workspace_bytes(amplitudes + 4 * ceil(rows * channels / 4), amplitudes)
  float* yf = static_cast<float*>(workspace);
  float* tf = yf + pad4(amp);
  float* hf = tf + pad4(tn);
  cudaError_t err = ac::bf16_to_f32(y, yf, amp, st);
  if (err == cudaSuccess && ton != nullptr) err = ac::bf16_to_f32(ton, tf, tn, st);
  if (err == cudaSuccess)
    err = ac::pa_threshold(plan->tb, yf, ton != nullptr ? tf : nullptr, drown, 1.0f, hf, nullptr, rows, channels, st);
  if (err == cudaSuccess) err = ac::f32_to_bf16(hf, thr, amp, st);
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "pa_threshold (bfloat16) launch");
}

// ----------------------------------------------------------------------------------------- dB utilities
int ac_pa_amplitude_to_db_f32(const ac_pa_plan* plan, const float* a, float* out, int64_t n, int normalised, void* stream) {
  if (plan == nullptr) return fail(AC_ERR_INVALID, "plan is null");
  if (n < 0) return fail(AC_ERR_INVALID, "negative size");
  if (n == 0) return AC_OK;
  if (a == nullptr || out == nullptr) return fail(AC_ERR_INVALID, "null tensor");
  cudaError_t err = ac::amplitude_to_db_f32(a, out, n, plan->tb.eps, 120.f, plan->host.db_min, normalised != 0,
                                            static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "amplitude_to_dB launch");
}

int ac_pa_amplitude_to_db_f64(const ac_pa_plan* plan, const double* a, double* out, int64_t n, int normalised, void* stream) {
  if (plan == nullptr) return fail(AC_ERR_INVALID, "plan is null");
  if (n < 0) return fail(AC_ERR_INVALID, "negative size");
  if (n == 0) return AC_OK;
  if (a == nullptr || out == nullptr) return fail(AC_ERR_INVALID, "null tensor");
  cudaError_t err = ac::amplitude_to_db_f64(a, out, n, 1e-14, 120., 10. * std::log(1e-14) / std::log(10.) + 120., normalised != 0,
                                            static_cast<cudaStream_t>(stream));
  return err == cudaSuccess ? AC_OK : cuda_fail(err, "amplitude_to_dB launch");
}

// ---------------------------------------------------------------------------------------------- DLPack
int ac_mdct_forward_dl(const ac_mdct_plan* plan, struct DLManagedTensor* x, struct DLManagedTensor* y, void* stream) {
  if (plan == nullptr) return fail(AC_ERR_INVALID, "plan is null");
  void *xd, *yd;
  const int64_t *xs, *ys;
  if (int rc = unwrap_dl(x, "x", 3, 2, &xd, &xs)) return rc;
  if (int rc = unwrap_dl(y, "y", 4, 2, &yd, &ys)) return rc;
  const int n = plan->tb.n;
  if (xs[1] % n != 0)
    return fail(AC_ERR_INVALID, "samples_n (%lld) must be a multiple of filters_n (%d)", (long long)xs[1], n);
  if (ys[0] != xs[0] || ys[1] != xs[1] / n + 1 || ys[2] != n || ys[3] != xs[2])
    return fail(AC_ERR_INVALID, "y must have shape [batches, samples / filters_n + 1, filters_n, channels]");
  return ac_mdct_forward_f32(plan, static_cast<const float*>(xd), static_cast<float*>(yd), xs[0], xs[1], (int)xs[2], stream);
}

int ac_mdct_inverse_dl(const ac_mdct_plan* plan, struct DLManagedTensor* y, struct DLManagedTensor* x, void* stream) {
  if (plan == nullptr) return fail(AC_ERR_INVALID, "plan is null");
  void *xd, *yd;
  const int64_t *xs, *ys;
  if (int rc = unwrap_dl(y, "mdct_amplitudes", 4, 2, &yd, &ys)) return rc;
  if (int rc = unwrap_dl(x, "x", 3, 2, &xd, &xs)) return rc;
  const int n = plan->tb.n;
  if (ys[2] != n) return fail(AC_ERR_INVALID, "mdct_amplitudes.shape[2] (%lld) != filters_n (%d)", (long long)ys[2], n);
  if (xs[0] != ys[0] || xs[1] != (ys[1] + 1) * n || xs[2] != ys[3])
    return fail(AC_ERR_INVALID, "x must have shape [batches, (blocks + 1) * filters_n, channels]");
  return ac_mdct_inverse_f32(plan, static_cast<const float*>(yd), static_cast<float*>(xd), ys[0], ys[1], (int)ys[3], stream);
}

int ac_pa_tonality_dl(const ac_pa_plan* plan, struct DLManagedTensor* y, struct DLManagedTensor* ton, void* stream) {
  if (plan == nullptr) return fail(AC_ERR_INVALID, "plan is null");
  void *yd, *td;
  const int64_t *ys, *ts;
  if (int rc = unwrap_dl(y, "mdct_amplitudes", 4, 2, &yd, &ys)) return rc;
  if (int rc = unwrap_dl(ton, "tonality", 4, 2, &td, &ts)) return rc;
  if (ys[2] != plan->tb.n) return fail(AC_ERR_INVALID, "mdct_amplitudes.shape[2] != filter_bands_n");
  if (ts[0] != ys[0] || ts[1] != ys[1] || ts[2] != 1 || ts[3] != ys[3])
    return fail(AC_ERR_INVALID, "tonality must have shape [batches, blocks, 1, channels]");
  return ac_pa_tonality_f32(plan, static_cast<const float*>(yd), static_cast<float*>(td), ys[0], ys[1], (int)ys[3], stream);
}

int ac_pa_threshold_dl(const ac_pa_plan* plan, struct DLManagedTensor* y, struct DLManagedTensor* ton_or_null, float drown,
                       struct DLManagedTensor* thr, void* stream) {
  if (plan == nullptr) return fail(AC_ERR_INVALID, "plan is null");
  void *yd, *td = nullptr, *hd;
  const int64_t *ys, *ts, *hs;
  if (int rc = unwrap_dl(y, "mdct_amplitudes", 4, 2, &yd, &ys)) return rc;
  if (int rc = unwrap_dl(thr, "threshold", 4, 2, &hd, &hs)) return rc;
  if (ys[2] != plan->tb.n) return fail(AC_ERR_INVALID, "mdct_amplitudes.shape[2] != filter_bands_n");
  for (int d = 0; d < 4; ++d)
    if (hs[d] != ys[d]) return fail(AC_ERR_INVALID, "threshold must have the shape of mdct_amplitudes");
  if (ton_or_null != nullptr) {
    if (int rc = unwrap_dl(ton_or_null, "tonality_per_block", 4, 2, &td, &ts)) return rc;
    if (ts[0] != ys[0] || ts[1] != ys[1] || ts[2] != 1 || ts[3] != ys[3])
      return fail(AC_ERR_INVALID, "tonality_per_block must have shape [batches, blocks, 1, channels]");
  }
  return ac_pa_threshold_f32(plan, static_cast<const float*>(yd), static_cast<const float*>(td), drown,
                             static_cast<float*>(hd), ys[0], ys[1], (int)ys[3], stream);
}

}  // extern "C"
