// Host-side table construction; see tables.h.  Reference line numbers are into
// /root/reference/audiocodec/{mdctransformer,psychoacoustic}.py.
#include "tables.h"

#include <algorithm>
#include <cmath>

namespace ac {

namespace {
constexpr double kPi = 3.14159265358979323846;

template <typename T>
void window_and_folds(int n, int window_type, MdctTables& t) {
  const int h = n / 2;
  std::vector<T> w(n + h);
  for (int k = 0; k < n + h; ++k) {
    const T pos = static_cast<T>(k + 0.5);                        // tf.range(0.5, 3N/2 + 0.5)   (:202,207)
    if (window_type == 1) {
      w[k] = std::sin(static_cast<T>(kPi / (2 * n)) * pos);      // sine window                  (:201-203)
    } else if (window_type == 2) {
      const T s = std::sin(static_cast<T>(kPi / (2. * n)) * pos);
      w[k] = std::sin(static_cast<T>(kPi / 2.) * (s * s));       // vorbis power-complementary   (:205-208)
    } else {
      w[k] = static_cast<T>(1);                                   // "no modified window"         (:211)
    }
  }
  t.window.assign(w.begin(), w.end());
  t.fold.resize(4 * h);
  t.unfold.resize(4 * h);
  for (int p = 0; p < h; ++p) {
    // consistency rule with its cancellation, g = (1 - w[N+p] w[N-1-p]) / w[p]                    (:219-221)
    const T g = (static_cast<T>(1) - w[n + p] * w[n - 1 - p]) / w[p];
    const T a0 = w[p];           // F[p,     h-1-p]  upper-left anti-diagonal                      (:214)
    const T a1 = w[n - 1 - p];   // F[N-1-p, h-1-p]  lower-left diagonal                           (:215)
    const T a2 = w[n + p];       // F[p,     h+p]    upper-right diagonal                          (:216)
    const T a3 = -g;             // F[N-1-p, h+p]    lower-right anti-diagonal                     (:226)
    t.fold[4 * p + 0] = a0;
    t.fold[4 * p + 1] = a1;
    t.fold[4 * p + 2] = a2;
    t.fold[4 * p + 3] = a3;
    // inv(F) block-wise (:185): [x_p, x_q] = [u_a, u_b] * inv([[a0, a2], [a1, a3]])
    const T det = a0 * a3 - a2 * a1;
    t.unfold[4 * p + 0] = a3 / det;    // Finv[h-1-p, p]
    t.unfold[4 * p + 1] = -a1 / det;   // Finv[h+p,   p]
    t.unfold[4 * p + 2] = -a2 / det;   // Finv[h-1-p, N-1-p]
    t.unfold[4 * p + 3] = a0 / det;    // Finv[h+p,   N-1-p]
  }
}
}  // namespace

MdctTables build_mdct_tables(int n, int window_type, bool precompute_f32) {
  MdctTables t;
  t.n = n;
  if (precompute_f32) {
    window_and_folds<float>(n, window_type, t);
  } else {
    window_and_folds<double>(n, window_type, t);
  }
  return t;
}

PaTables build_pa_tables(double sample_rate, int n, int nb, double alpha) {
  PaTables t;
  t.n = n;
  t.nb = nb;
  t.alpha = alpha;
  t.sample_rate = sample_rate;
  t.max_frequency = sample_rate / 2.0;                                   // (:61)
  t.max_bark = 6. * std::asinh(t.max_frequency / 600.);                  // (:62, :335)
  t.bark_band_width = t.max_bark / nb;                                   // (:63)
  // _dB_MIN = amplitude_to_dB(eps) evaluated in the fp32 compute dtype   (:58, :83-84)
  {
    const float eps = 1e-14f;
    t.db_min = 10.f * std::log(std::max(eps, eps * eps)) / std::log(10.f) + 120.f;
  }
  auto bark2freq = [](double bark) { return 600. * std::sinh(bark / 6.); };   // (:339)

  // ---- W / W_inv: fractional overlap of filter band k with bark band i    (:281-299)
  const double fw = t.max_frequency / n;
  t.w.assign(static_cast<size_t>(n) * nb, 0.0);
  t.w_inv.assign(static_cast<size_t>(nb) * n, 0.0);
  for (int i = 0; i < nb; ++i) {
    const double bark_low = t.bark_band_width * i;
    const double lo = bark2freq(bark_low);
    const double hi = bark2freq(bark_low + t.bark_band_width);
    for (int k = 0; k < n; ++k) {
      const double f_lo = fw * k;
      const double f_hi = f_lo + fw;
      const double lo_c = std::min(std::max(lo, f_lo), f_hi);
      const double hi_c = std::min(std::max(hi, f_lo), f_hi);
      const double overlap = hi_c - lo_c;
      t.w[static_cast<size_t>(k) * nb + i] = overlap / fw;
      t.w_inv[static_cast<size_t>(i) * n + k] = overlap / (hi - lo);
    }
  }

  // ---- threshold in quiet, Zoelzer (9.3)                                   (:240-253)
  t.quiet.resize(nb);
  for (int i = 0; i < nb; ++i) {
    const double mid = t.bark_band_width * i + t.bark_band_width / 2.;
    const double khz = bark2freq(mid) / 1000.;
    double db = 3.64 * std::pow(khz, -0.8) - 6.5 * std::exp(-0.6 * std::pow(khz - 3.3, 2.)) + 1e-3 * std::pow(khz, 4.);
    db = std::min(std::max(db, static_cast<double>(t.db_min)), 120.);
    t.quiet[i] = std::pow(10.0, (db - 120.) / 10);
  }

  // ---- spreading prototype on linspace(-max_bark, max_bark, 2 nb)          (:219-223)
  t.spread_fn.resize(2 * nb);
  const double step = (t.max_bark + t.max_bark) / (2 * nb - 1);
  for (int s = 0; s < 2 * nb; ++s) {
    const double z = (s == 2 * nb - 1) ? t.max_bark : -t.max_bark + s * step;
    const double f = 15.81 + 7.5 * (z + 0.474) - 17.5 * std::sqrt(1 + std::pow(z + 0.474, 2));
    t.spread_fn[s] = std::pow(10.0, alpha * f / 10.0);
  }

  // ---- tf.linspace(0, max_bark, nb) in the compute dtype                    (:187-189)
  t.lin.assign(nb, 0.f);
  if (nb > 1) {
    const float stop = static_cast<float>(t.max_bark);
    const float delta = stop / static_cast<float>(nb - 1);
    for (int j = 0; j < nb; ++j) t.lin[j] = static_cast<float>(j) * delta;
    t.lin[nb - 1] = stop;
  }

  // ---- staircase sparsity: per band the filter range, per filter the band range
  t.band_k0.assign(nb, 0);
  t.band_cnt.assign(nb, 0);
  t.band_ptr.assign(nb + 1, 0);
  for (int i = 0; i < nb; ++i) {
    int first = -1, last = -1;
    for (int k = 0; k < n; ++k) {
      if (static_cast<float>(t.w[static_cast<size_t>(k) * nb + i]) != 0.f) {
        if (first < 0) first = k;
        last = k;
      }
    }
    t.band_k0[i] = first < 0 ? 0 : first;
    t.band_cnt[i] = first < 0 ? 0 : last - first + 1;
    t.band_ptr[i + 1] = t.band_ptr[i] + t.band_cnt[i];
    for (int k = t.band_k0[i]; k < t.band_k0[i] + t.band_cnt[i]; ++k)
      t.band_w.push_back(static_cast<float>(t.w[static_cast<size_t>(k) * nb + i]));
    t.max_band_cnt = std::max(t.max_band_cnt, t.band_cnt[i]);
  }
  t.filt_b0.assign(n, 0);
  t.filt_cnt.assign(n, 0);
  t.filt_ptr.assign(n + 1, 0);
  for (int k = 0; k < n; ++k) {
    int first = -1, last = -1;
    for (int i = 0; i < nb; ++i) {
      if (static_cast<float>(t.w_inv[static_cast<size_t>(i) * n + k]) != 0.f) {
        if (first < 0) first = i;
        last = i;
      }
    }
    t.filt_b0[k] = first < 0 ? 0 : first;
    t.filt_cnt[k] = first < 0 ? 0 : last - first + 1;
    t.filt_ptr[k + 1] = t.filt_ptr[k] + t.filt_cnt[k];
    for (int i = t.filt_b0[k]; i < t.filt_b0[k] + t.filt_cnt[k]; ++i)
      t.filt_w.push_back(static_cast<float>(t.w_inv[static_cast<size_t>(i) * n + k]));
    t.max_filt_cnt = std::max(t.max_filt_cnt, t.filt_cnt[k]);
  }
  return t;
}

}  // namespace ac
