"""CPU oracle: NumPy restatement of the audiocodec encode/decode hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product path (audiocodec_b200/) never does and fails loudly
when its CUDA library is missing.

What it restates (all file:line into /root/reference/audiocodec/):
  * MDCT analysis / synthesis filter bank   mdctransformer.py:13-59, 61-125, 127-153, 155-368
  * psychoacoustic model                    psychoacoustic.py:14-69, 71-148, 169-339
  * quantize / dequantize                   NO reference symbol exists (SURVEY.md finding 2).  The spec
    is build-defined from add_noise (psychoacoustic.py:150-167, "masking_threshold = 6*sigma"): a uniform
    mid-tread quantiser with step = threshold, q = rint(A / thr) (round-half-even == tf.round),
    A_hat = q * thr.  PARITY UNPINNED for this step (nothing in the reference to pin it to).

Pinning: the unmodified reference sources, executed under oracle/tf_shim (a NumPy implementation of the
TF ops they call), pass the reference's own seven unit tests including the golden vector at
tests/test_mdctransformer.py:51-52, and generated the fixtures in tests/golden/ (make_golden.py).  This
restatement is checked against those fixtures and against the reference's tests re-expressed in
tests/test_oracle.py.  Real TensorFlow was never available (not installable offline), so the TF kernels
themselves (Eigen conv, FFT-based DCT, einsum) are represented by their documented semantics.

Three arithmetic modes, selected by `compute_dtype`:
  * "bfloat16"  - every op rounds its result to bfloat16 (ml_dtypes), contractions / reductions accumulate in float32,
    the DCT runs in float32 (mdctransformer.py:326-344).  UNPINNED: TensorFlow's own bfloat16 kernels were never run.
  * np.float64  - "truth": every op in double (what the reference computes with compute_dtype=tf.float64)
  * np.float32  - "fp32-faithful": tables built in float64 then cast (mdctransformer.py:58-59,
    psychoacoustic.py:65-69), data-path arithmetic in float32 like the reference's default graph.
"""

import math

import numpy as np
import scipy.fft

__all__ = ["MDCTransformer", "PsychoacousticModel", "quantize", "dequantize", "synthetic_audio"]


def _bfloat16():
  import ml_dtypes          # numpy ufuncs on ml_dtypes.bfloat16 compute in float32 and round: TF's per-op rounding
  return ml_dtypes.bfloat16


def _as_np_dtype(dtype):
  if isinstance(dtype, str) and dtype in ("bfloat16", "bf16"):
    return _bfloat16()
  return np.dtype(dtype).type


def _is_bf16(ct):
  return np.dtype(ct).name == "bfloat16"


def _einsum(ct, spec, *ops):
  """np.einsum in the compute dtype; bfloat16 operands are contracted with float32 accumulation and rounded once, which is
  how TF's CPU kernels (Eigen) run a bfloat16 contraction or reduction."""
  if _is_bf16(ct):
    return np.einsum(spec, *[np.asarray(o, dtype=np.float32) for o in ops]).astype(ct)
  return np.einsum(spec, *ops).astype(ct)


# ================================================================================================ MDCT
class MDCTransformer:
  """Restates mdctransformer.py:12-368.  Layouts: x [B, S, C] <-> Y [B, S/N + 1, N, C]."""

  def __init__(self, filters_n=1024, window_type='vorbis', compute_dtype=np.float32, precompute_dtype=np.float64):
    assert (filters_n % 2) == 0, "number of filters used in mdct transformation needs to be even"  # :26
    self.filters_n = int(filters_n)
    self.window_type = window_type
    self.compute_dtype = _as_np_dtype(compute_dtype)
    pd = _as_np_dtype(precompute_dtype)
    self.H = self._analysis_taps(pd).astype(self.compute_dtype)       # :58
    self.H_inv = self._synthesis_taps(pd).astype(self.compute_dtype)  # :59

  # ---- tables --------------------------------------------------------------------------------------
  def window_samples(self, pd=np.float64):
    """The 3N/2 window points the reference generates (mdctransformer.py:199-211)."""
    n = self.filters_n
    # tf.range(0.5, 3N//2 + 0.5) is a float32 range of exact half-integers, then cast (:202-203, :207-208)
    t = np.arange(0.5, (3 * n) // 2 + 0.5, dtype=np.float32).astype(pd)
    kind = self.window_type.lower()  # window_type=None raises AttributeError here, as in the reference (:199)
    if kind == 'sine':
      return np.sin(math.pi / (2 * n) * t)
    if kind == 'vorbis':
      return np.sin(math.pi / 2. * np.sin(math.pi / (2. * n) * t) ** 2)
    return np.ones(n + n // 2, dtype=pd)

  def fold_matrix(self, pd=np.float64):
    """Diamond-shaped fold matrix F [N, N] (mdctransformer.py:213-229)."""
    n, h = self.filters_n, self.filters_n // 2
    w = self.window_samples(pd)
    f = np.zeros((n, n), dtype=pd)
    r = np.arange(h)
    f[r, h - 1 - r] = w[r]                 # upper-left: anti-diagonal of w[0:h]           (:214)
    f[h + r, r] = w[h + r]                 # lower-left: diagonal of w[h:N]                (:215)
    f[r, h + r] = w[n + r]                 # upper-right: diagonal of w[N:3N/2]            (:216)
    # consistency rule, incl. its cancellation: g[i] = (1 - w[N+i] w[N-1-i]) / w[i]        (:219-221)
    g = (np.ones(h, dtype=pd) - w[n:n + h] * w[n - 1 - r]) / w[r]
    # reverse -> diag -> column-reverse -> negate                                         (:219,226)
    f[n - 1 - r, h + r] = -g
    return f

  def _analysis_taps(self, pd):
    """H [2, N, N]: tap 0 acts on the current block, tap 1 on the previous one (:155-174, :231-242)."""
    n, h = self.filters_n, self.filters_n // 2
    f = self.fold_matrix(pd)
    upper = np.concatenate([np.zeros(h, dtype=pd), np.ones(h, dtype=pd)])   # delay matrix, z^0 part (:238)
    lower = np.concatenate([np.ones(h, dtype=pd), np.zeros(h, dtype=pd)])   # z^-1 part              (:240)
    delay = np.stack([np.diag(upper), np.diag(lower)], axis=0)              # [2, N, N]
    taps = _polymatmul(f[:, None, :], delay)                                # [N, 2, N]              (:172)
    return np.transpose(taps, (1, 0, 2))

  def _synthesis_taps(self, pd):
    """H_inv [2, N, N] from inv(F) and the causal inverse delay (:176-190, :244-255)."""
    n, h = self.filters_n, self.filters_n // 2
    f_inv = np.linalg.inv(self.fold_matrix(pd))                              # (:185)
    now = np.concatenate([np.ones(h, dtype=pd), np.zeros(h, dtype=pd)])      # (:251)
    late = np.concatenate([np.zeros(h, dtype=pd), np.ones(h, dtype=pd)])     # (:253)
    delay_inv = np.stack([np.diag(now), np.diag(late)], axis=1)              # [N, 2, N]
    taps = _polymatmul(delay_inv, f_inv[None, :, :])                         # [N, 2, N]             (:188)
    return np.transpose(taps, (1, 0, 2))

  # ---- data path -----------------------------------------------------------------------------------
  def _dct4(self, u):
    """Orthonormal DCT-IV along the last axis, y_k = sqrt(2/N) sum_n u_n cos(pi/N (n+1/2)(k+1/2)) (:314).

    The reference reaches it through a zero-interleaved DCT-III (:333-347); the transform computed is the
    same.  SciPy evaluates it in the dtype of `u` (float32 stays float32).
    """
    if _is_bf16(u.dtype):
      # up-cast to float32 (:326-330), DCT-III of the zero-interleaved signal = DCT-IV / sqrt(2), down-cast (:340-344),
      # then the bfloat16 constant sqrt(2) (:347)
      y = (scipy.fft.dct(u.astype(np.float32), type=4, norm='ortho', axis=-1) / np.float32(math.sqrt(2.0))).astype(u.dtype)
      return (np.sqrt(u.dtype.type(2.)) * y).astype(u.dtype)
    return scipy.fft.dct(u, type=4, norm='ortho', axis=-1).astype(u.dtype)

  def transform(self, x):
    """x [B, S, C] -> Y [B, S/N + 1, N, C]  (mdctransformer.py:61-125)."""
    x = np.asarray(x)
    if x.dtype != np.dtype(self.compute_dtype):
      raise TypeError("input dtype must equal compute_dtype (mdctransformer.py:22-23)")
    b, s, c = x.shape
    n = self.filters_n
    if s % n != 0:
      raise ValueError("samples_n must be a multiple of filters_n (mdctransformer.py:287)")
    blocks = np.transpose(x, (0, 2, 1)).reshape(b * c, s // n, n)            # (:292-295)
    frames = self._dct4(_polymatmul(blocks, self.H))                         # [BC, M+1, N]  (:118)
    frames = frames.reshape(b, c, frames.shape[1], n).transpose(0, 2, 3, 1)  # (:121-122)
    scale = self.compute_dtype(1.) / np.sqrt(self.compute_dtype(4.) * self.compute_dtype(n))  # (:125)
    return (scale * frames).astype(self.compute_dtype)

  def inverse_transform(self, mdct_amplitudes):
    """Y [B, M', N, C] -> x_hat [B, (M'+1) N, C]  (mdctransformer.py:127-153)."""
    y = np.asarray(mdct_amplitudes)
    if y.dtype != np.dtype(self.compute_dtype):
      raise TypeError("input dtype must equal compute_dtype")
    b, m, n, c = y.shape
    assert n == self.filters_n
    frames = np.transpose(y, (0, 3, 1, 2)).reshape(b * c, m, n)              # (:141-142)
    rescaled = np.sqrt(self.compute_dtype(4.) * self.compute_dtype(n)) * frames   # (:145)
    blocks = _polymatmul(self._dct4(rescaled.astype(self.compute_dtype)), self.H_inv)  # (:148)
    out = blocks.reshape(b, c, -1).transpose(0, 2, 1)                        # (:306-307)
    return np.ascontiguousarray(out).astype(self.compute_dtype)


def _polymatmul(a, f):
  """C[b, n, k] = sum_m sum_q A[b, m, q] F[n - m, q, k]  (full polynomial product, mdctransformer.py:349-368).

  A's 2nd axis and F's 1st axis hold polynomial coefficients in z^-1; the result has
  deg(A) + deg(F) + 1 coefficients (the reference pads A with deg(F) zero blocks on both sides and runs a
  VALID correlation with the flipped F, which is this sum).
  """
  blocks = a.shape[1]
  taps = f.shape[0]
  if _is_bf16(a.dtype):      # one convolution op: float32 accumulation over both taps, one rounding
    out = np.zeros((a.shape[0], blocks + taps - 1, f.shape[2]), dtype=np.float32)
    for t in range(taps):
      out[:, t:t + blocks, :] += np.matmul(a.astype(np.float32), f[t].astype(np.float32))
    return out.astype(a.dtype)
  out = np.zeros((a.shape[0], blocks + taps - 1, f.shape[2]), dtype=np.result_type(a, f))
  for t in range(taps):
    out[:, t:t + blocks, :] += np.matmul(a, f[t])
  return out


# ====================================================================================== psychoacoustics
class PsychoacousticModel:
  """Restates psychoacoustic.py:13-339.  Amplitude layout [B, M, N, C]; tonality [B, M, 1, C]."""

  def __init__(self, sample_rate, filter_bands_n=1024, bark_bands_n=64, alpha=0.6,
               compute_dtype=np.float32, precompute_dtype=np.float64):
    self.alpha = alpha
    self.sample_rate = sample_rate
    self.bark_bands_n = int(bark_bands_n)
    self.filter_bands_n = int(filter_bands_n)
    ct = _as_np_dtype(compute_dtype)
    if ct not in (np.float64, np.float32) and not _is_bf16(ct):
      raise TypeError("compute_dtype should be float64, float32 or bfloat16 (:42-43)")
    self.compute_dtype = ct
    pd = _as_np_dtype(precompute_dtype)

    self._dB_MAX = ct(120.)                                                   # (:52)
    self._INTENSITY_EPS = ct(1e-14)                                           # (:56)
    self._dB_MIN = self.amplitude_to_dB(self._INTENSITY_EPS)                  # (:58) == -20 dB

    self.max_frequency = pd(self.sample_rate) / pd(2.0)                       # (:61)
    self.max_bark = self.freq2bark(self.max_frequency)                        # (:62)
    self.bark_band_width = self.max_bark / self.bark_bands_n                  # (:63)

    w, w_inv = self._bark_freq_mapping(pd)
    self.W = w.astype(ct)                                                     # (:66)
    self.W_inv = w_inv.astype(ct)                                             # (:67)
    self.quiet_threshold_intensity = self._quiet_threshold_intensity_in_bark(pd).astype(ct)   # (:68)
    self.spreading_matrix = self._spreading_matrix_in_bark().astype(ct)       # (:69)

  # ---- utilities -----------------------------------------------------------------------------------
  def amplitude_to_dB(self, mdct_amplitude):
    """psychoacoustic.py:71-85."""
    ct = self.compute_dtype
    a = np.asarray(mdct_amplitude, dtype=ct)
    return ct(10.) * np.log(np.maximum(self._INTENSITY_EPS, a ** ct(2.0))) / np.log(ct(10.)) + self._dB_MAX

  def amplitude_to_dB_norm(self, mdct_amplitude):
    """psychoacoustic.py:87-100."""
    return (self.amplitude_to_dB(mdct_amplitude) - self._dB_MIN) / (self._dB_MAX - self._dB_MIN)

  @staticmethod
  def freq2bark(frequencies):
    """psychoacoustic.py:333-335."""
    return 6. * np.arcsinh(frequencies / 600.)

  @staticmethod
  def bark2freq(bark_band):
    """psychoacoustic.py:337-339."""
    return 600. * np.sinh(bark_band / 6.)

  # ---- tables --------------------------------------------------------------------------------------
  def _bark_freq_mapping(self, pd):
    """W [N, nb], W_inv [nb, N]: fractional interval overlaps (psychoacoustic.py:257-299)."""
    n, nb = self.filter_bands_n, self.bark_bands_n
    filter_band_width = self.max_frequency / n                                # (:281)
    band = np.arange(nb, dtype=pd).reshape(1, nb)
    filt = np.arange(n, dtype=pd).reshape(n, 1)
    bark_low = self.bark_band_width * band                                    # (:284)
    lo_hz = np.broadcast_to(self.bark2freq(bark_low), (n, nb))                # (:285)
    hi_hz = np.broadcast_to(self.bark2freq(bark_low + self.bark_band_width), (n, nb))   # (:286)
    f_lo = filter_band_width * filt                                           # (:288)
    lo_clip = np.minimum(np.maximum(lo_hz, f_lo), f_lo + filter_band_width)   # (:289)
    hi_clip = np.minimum(np.maximum(hi_hz, f_lo), f_lo + filter_band_width)   # (:290)
    overlap = hi_clip - lo_clip                                               # (:292)
    w = overlap / filter_band_width
    w_inv_t = overlap / (hi_hz - lo_hz)                                       # (:293)
    return w.astype(pd), np.transpose(w_inv_t).astype(pd)

  def _quiet_threshold_intensity_in_bark(self, pd):
    """Zoelzer (9.3) at the bark-band mid frequencies, [1, 1, nb, 1] (psychoacoustic.py:232-255)."""
    nb = self.bark_bands_n
    mid_bark = self.bark_band_width * np.arange(nb, dtype=pd) + self.bark_band_width / 2.   # (:240)
    khz = self.bark2freq(mid_bark) / 1000.                                    # (:241)
    db = (3.64 * np.power(khz, -0.8)
          - 6.5 * np.exp(-0.6 * np.power(khz - 3.3, 2.))
          + 1e-3 * np.power(khz, 4.))                                         # (:246-248)
    db = np.minimum(np.maximum(db, pd(self._dB_MIN)), pd(self._dB_MAX))       # (:245,249)
    intensity = np.power(pd(10.0), (db - pd(self._dB_MAX)) / 10)              # (:252-253)
    return intensity.reshape(1, 1, nb, 1)

  def _spreading_matrix_in_bark(self):
    """Toeplitz spreading matrix S[i (masker), j (maskee)] (psychoacoustic.py:212-230); always float64."""
    nb = self.bark_bands_n
    mb = np.float64(self.max_bark)
    # tf.linspace(-max_bark, max_bark, 2 nb): start + i * (stop - start) / (2 nb - 1)     (:220)
    z = -mb + np.arange(2 * nb, dtype=np.float64) * ((mb + mb) / np.float64(2 * nb - 1))
    z[-1] = mb
    proto_db = 15.81 + 7.5 * (z + 0.474) - 17.5 * np.sqrt(1 + np.power(z + 0.474, 2))     # (:219)
    proto = np.power(np.float64(10.0), self.alpha * proto_db / 10.0)          # (:223)
    return np.stack([proto[nb - row:2 * nb - row] for row in range(nb)], axis=0)   # (:227-228)

  # ---- data path -----------------------------------------------------------------------------------
  def tonality(self, mdct_amplitudes):
    """Spectral-flatness tonality, [B, M, N, C] -> [B, M, 1, C] (psychoacoustic.py:102-120)."""
    ct = self.compute_dtype
    a = self._check(mdct_amplitudes)
    intensity = np.power(a, ct(2))                                            # (:113)
    acc = np.float32 if _is_bf16(ct) else ct          # a bfloat16 reduction accumulates in float32 and rounds once
    log_gm = np.mean(np.log(np.maximum(self._INTENSITY_EPS, intensity)), axis=2, keepdims=True, dtype=acc).astype(ct)
    am = np.mean(intensity, axis=2, keepdims=True, dtype=acc).astype(ct) + self._INTENSITY_EPS
    sfm = ct(10.) * np.log(np.exp(log_gm) / am) / ct(math.log(10.0))          # (:114-116)
    return np.minimum(sfm / ct(-60.), ct(1.0)).astype(ct)                     # (:118)

  def global_masking_threshold(self, mdct_amplitudes, tonality_per_block, drown=0.0):
    """[B, M, N, C], [B, M, 1, C] -> threshold amplitude [B, M, N, C] (psychoacoustic.py:122-148)."""
    masking = self._masking_intensity_in_bark(mdct_amplitudes, tonality_per_block, drown)
    total = np.maximum(masking, self.quiet_threshold_intensity)               # (:144)
    return self._bark_intensity_to_freq_ampl(total)                           # (:146)

  def add_noise(self, mdct_amplitudes, masking_threshold, rng=None):
    """A + thr * N(0, 1/6)  (psychoacoustic.py:150-167)."""
    ct = self.compute_dtype
    rng = np.random.default_rng() if rng is None else rng
    a = self._check(mdct_amplitudes)
    noise = masking_threshold * (rng.standard_normal(size=a.shape) / 6.).astype(ct)
    return a + noise

  def masking_offset_factor(self, tonality_per_block, drown=0.0):
    """10^(-alpha * offset[j] / 10), [B, M, nb, C] (psychoacoustic.py:185-191,197)."""
    ct = self.compute_dtype
    ton = np.asarray(tonality_per_block, dtype=ct)
    nb = self.bark_bands_n
    # tf.linspace(0, max_bark, nb) evaluated in the compute dtype (:187-189)
    stop = ct(self.max_bark)
    lin = (np.arange(nb).astype(ct) * (stop / ct(nb - 1))).astype(ct) if nb > 1 else np.zeros(1, ct)
    if nb > 1:
      lin[-1] = stop
    offset = ct(1. - drown) * (_einsum(ct, 'nbic,j->nbjc', ton, lin) + ct(9.) * ton + ct(5.5))   # (:185-191)
    return np.power(ct(10.0), ct(-self.alpha) * offset / ct(10.0)).astype(ct)                  # (:197)

  def _masking_intensity_in_bark(self, mdct_amplitudes, tonality_per_block, drown=0.0):
    """psychoacoustic.py:169-210 without materialising the [B, M, nb, nb, C] tensor.

    masking_matrix[n,b,i,j,c] = S[i,j] * gain[n,b,j,c] (:195-197), so
    sum_i P[i] * masking_matrix[i,j] == gain[j] * sum_i P[i] * S[i,j].  (`materialise=True` in
    `_masking_intensity_dense` keeps the reference's literal form for cross-checks.)
    """
    ct = self.compute_dtype
    gain = self.masking_offset_factor(tonality_per_block, drown)
    bark = self._to_bark_intensity(mdct_amplitudes)                                        # (:204)
    p = np.power(np.maximum(self._INTENSITY_EPS, bark), ct(self.alpha))                    # (:206)
    spread = _einsum(ct, 'nbic,ij->nbjc', p, self.spreading_matrix) * gain        # (:205-207)
    return np.power(np.maximum(self._INTENSITY_EPS, spread), ct(1. / self.alpha)).astype(ct)   # (:208)

  def _masking_intensity_dense(self, mdct_amplitudes, tonality_per_block, drown=0.0):
    """Literal form of psychoacoustic.py:195-208 (materialised 5-D masking matrix); small inputs only."""
    ct = self.compute_dtype
    gain = self.masking_offset_factor(tonality_per_block, drown)
    masking_matrix = _einsum(ct, 'ij,nbjc->nbijc', self.spreading_matrix, gain)              # (:195)
    bark = self._to_bark_intensity(mdct_amplitudes)
    p = np.power(np.maximum(self._INTENSITY_EPS, bark), ct(self.alpha))
    spread = _einsum(ct, 'nbic,nbijc->nbjc', p, masking_matrix)                   # (:205)
    return np.power(np.maximum(self._INTENSITY_EPS, spread), ct(1. / self.alpha)).astype(ct)

  def _to_bark_intensity(self, mdct_amplitudes):
    """I_bark = A^2 . W  (psychoacoustic.py:301-315)."""
    ct = self.compute_dtype
    a = self._check(mdct_amplitudes)
    return _einsum(ct, 'nbic,ij->nbjc', np.power(a, ct(2)), self.W)               # (:312-313)

  def _bark_intensity_to_freq_ampl(self, bark_intensity):
    """sqrt(max(eps, I_bark . W_inv))  (psychoacoustic.py:317-331)."""
    ct = self.compute_dtype
    spread = _einsum(ct, 'nbic,ij->nbjc', bark_intensity, self.W_inv)             # (:330)
    return np.power(np.maximum(self._INTENSITY_EPS, spread), ct(0.5)).astype(ct)           # (:331)

  def _check(self, a):
    a = np.asarray(a)
    if a.dtype != np.dtype(self.compute_dtype):
      raise TypeError("input dtype must equal compute_dtype (psychoacoustic.py:30)")
    return a


# =========================================================================================== quantiser
def quantize(mdct_amplitudes, masking_threshold):
  """q = rint(A / thr), int32.  Build-defined (no reference symbol); arithmetic in the input dtype."""
  a = np.asarray(mdct_amplitudes)
  thr = np.asarray(masking_threshold, dtype=a.dtype)
  return np.rint(a / thr).astype(np.int32)


def dequantize(q, masking_threshold):
  """A_hat = q * thr in the threshold's dtype."""
  thr = np.asarray(masking_threshold)
  return (np.asarray(q).astype(thr.dtype) * thr).astype(thr.dtype)


# ====================================================================================== synthetic input
def synthetic_audio(batch, samples, channels, sample_rate, first_clip=0, dtype=np.float32):
  """SURVEY.md 8(d) workload: 0.5 sin(2 pi f_b s / sr + phi_c) + 0.05 N(0,1), clipped to [-1, 1].

  Clip `b` is seeded with 1234 + global clip index so shards of a batch are reproducible on any rank.
  """
  out = np.empty((batch, samples, channels), dtype=dtype)
  s = np.arange(samples, dtype=np.float64)
  for b in range(batch):
    rng = np.random.default_rng(1234 + first_clip + b)
    f = 110.0 * 2.0 ** (6.0 * rng.random())          # log-uniform in [110, 7040] Hz
    noise = rng.standard_normal(size=(samples, channels))
    for c in range(channels):
      clip = 0.5 * np.sin(2.0 * np.pi * f * s / sample_rate + c * np.pi / 3.0) + 0.05 * noise[:, c]
      out[b, :, c] = np.clip(clip, -1.0, 1.0).astype(dtype)
  return out
