"""PsychoacousticModel: drop-in for /root/reference/audiocodec/psychoacoustic.py (class at :13-339).

Same constructor keywords, attributes, method names and tensor layouts as the reference; tonality, the
global masking threshold and the (build-defined) quantiser run as sm_100a kernels behind the C ABI.
"""

import ctypes
import math

import numpy as np
import torch

from . import _capi, autograd
from ._tensors import torch_dtype, adopt, bf16_workspace, normalise_compute_dtype, normalise_precompute_dtype, stream_ptr


class PsychoacousticModel:
  def __init__(self, sample_rate, filter_bands_n=1024, bark_bands_n=64, alpha=0.6,
               compute_dtype='float32', precompute_dtype='float64'):
    """Same arguments as the reference (psychoacoustic.py:14-15).

    :raises TypeError: compute_dtype outside {float64, float32, bfloat16} (:42-43)

    float32 is the tuned path, float64 runs functional kernels, bfloat16 runs the float32 kernels on bfloat16 tensors
    with W, W_inv, the quiet threshold, the spreading matrix, eps, alpha and 1 / alpha cast to bfloat16 as in the
    reference (:56, 65-69, 197, 206-208); the quantiser and add_noise are float32 / float64 only.
    """
    self.alpha = alpha
    self.sample_rate = sample_rate
    self.bark_bands_n = int(bark_bands_n)
    self.filter_bands_n = int(filter_bands_n)
    self.compute_dtype = normalise_compute_dtype(compute_dtype, "PsychoacousticModel")
    self._dtype = torch_dtype(self.compute_dtype)
    self._f64 = self.compute_dtype == "float64"
    self._bf16 = self.compute_dtype == "bfloat16"
    if normalise_precompute_dtype(precompute_dtype) != "float64":
      raise NotImplementedError("PsychoacousticModel tables are precomputed in float64")

    self._dB_MAX = 120.0                     # (:52)
    self._INTENSITY_EPS = 1e-14              # (:56)

    n, nb = self.filter_bands_n, self.bark_bands_n
    w = np.empty((n, nb), dtype=np.float32)
    w_inv = np.empty((nb, n), dtype=np.float32)
    quiet = np.empty(nb, dtype=np.float32)
    spreading = np.empty((nb, nb), dtype=np.float32)
    scalars = np.empty(4, dtype=np.float64)
    fp = ctypes.POINTER(ctypes.c_float)
    _capi.check(_capi.lib().ac_pa_tables_host(
      float(sample_rate), n, nb, float(alpha), w.ctypes.data_as(fp), w_inv.ctypes.data_as(fp),
      quiet.ctypes.data_as(fp), spreading.ctypes.data_as(fp), scalars.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
    self.max_frequency, self.max_bark, self.bark_band_width, self._dB_MIN = (float(v) for v in scalars)
    self.W = torch.from_numpy(w)                                                   # [N, nb]      (:66)
    self.W_inv = torch.from_numpy(w_inv)                                           # [nb, N]      (:67)
    self.quiet_threshold_intensity = torch.from_numpy(quiet).reshape(1, 1, nb, 1)  # (:68)
    self.spreading_matrix = torch.from_numpy(spreading)                            # [nb, nb]     (:69)
    if self._bf16:       # tf.cast(..., compute_dtype) (:66-69)
      self.W, self.W_inv = self.W.to(torch.bfloat16), self.W_inv.to(torch.bfloat16)
      self.quiet_threshold_intensity = self.quiet_threshold_intensity.to(torch.bfloat16)
      self.spreading_matrix = self.spreading_matrix.to(torch.bfloat16)
    self._plans = {}

  # ---- plans ------------------------------------------------------------------------------------------
  def _plan(self, device):
    index = device.index if device.index is not None else torch.cuda.current_device()
    plan = self._plans.get(index)
    if plan is None:
      handle = ctypes.c_void_p()
      with torch.cuda.device(index):
        _capi.check(_capi.lib().ac_pa_plan_create_ex(float(self.sample_rate), self.filter_bands_n, self.bark_bands_n,
                                                     float(self.alpha),
                                                     _capi.DTYPE_BF16 if self._bf16 else _capi.DTYPE_F32,
                                                     ctypes.byref(handle)))
      plan = self._plans[index] = handle
    return plan

  def __del__(self):
    for handle in getattr(self, "_plans", {}).values():
      try:
        _capi.lib().ac_pa_plan_destroy(handle)
      except Exception:
        pass

  def _check_amplitudes(self, a, name="mdct_amplitudes"):
    if a.dim() != 4 or a.shape[2] != self.filter_bands_n:
      raise ValueError(f"{name} must be [batches_n, blocks_n, {self.filter_bands_n}, channels_n], got {tuple(a.shape)}")

  # ---- utilities ---------------------------------------------------------------------------------------
  def _to_db(self, mdct_amplitude, normalised):
    """One element-wise kernel (ac_pa_amplitude_to_db_f32 / _f64) for device tensors; python scalars and host tensors
    (the reference calls it on a constant in __init__, :58) take the same formula on the host."""
    if isinstance(mdct_amplitude, torch.Tensor) and mdct_amplitude.is_cuda or \
        (not isinstance(mdct_amplitude, (torch.Tensor, np.ndarray, float, int)) and hasattr(mdct_amplitude, "__dlpack__")):
      a, back = adopt(mdct_amplitude, "mdct_amplitude", dtype=self._dtype)
      work = a.float() if self._bf16 else a             # bfloat16: the float32 kernel between two casts
      out = torch.empty_like(work)
      with torch.cuda.device(a.device):
        fn = _capi.lib().ac_pa_amplitude_to_db_f64 if self._f64 else _capi.lib().ac_pa_amplitude_to_db_f32
        _capi.check(fn(self._plan(a.device), work.data_ptr(), out.data_ptr(), work.numel(), int(normalised),
                       stream_ptr(a.device)))
      return back(out.to(self._dtype))
    a = torch.as_tensor(mdct_amplitude, dtype=torch.float64 if self._f64 else torch.float32)
    db = 10. * torch.log(torch.clamp_min(a ** 2.0, self._INTENSITY_EPS)) / math.log(10.) + self._dB_MAX
    return (db - self._dB_MIN) / (self._dB_MAX - self._dB_MIN) if normalised else db

  def amplitude_to_dB(self, mdct_amplitude):
    """psychoacoustic.py:71-85: 10 log10(max(eps, a^2)) + 120, in [_dB_MIN, _dB_MAX]."""
    return self._to_db(mdct_amplitude, False)

  def amplitude_to_dB_norm(self, mdct_amplitude):
    """psychoacoustic.py:87-100: the same on the normalised scale [0, 1]."""
    return self._to_db(mdct_amplitude, True)

  @staticmethod
  def freq2bark(frequencies):
    """Empirical Bark scale (psychoacoustic.py:333-335)."""
    return 6. * np.arcsinh(np.asarray(frequencies, dtype=np.float64) / 600.)

  @staticmethod
  def bark2freq(bark_band):
    """Empirical Bark scale (psychoacoustic.py:337-339)."""
    return 600. * np.sinh(np.asarray(bark_band, dtype=np.float64) / 6.)

  # ---- data path --------------------------------------------------------------------------------------
  def tonality(self, mdct_amplitudes):
    """Spectral-flatness tonality, 0 (noise) .. 1 (tonal) (psychoacoustic.py:102-120).

    :param mdct_amplitudes: [batches_n, blocks_n, filter_bands_n, channels_n], float32, CUDA
    :return:                [batches_n, blocks_n, 1, channels_n]
    """
    if autograd.wants_grad(mdct_amplitudes) and self.compute_dtype == "float32":
      return autograd.tonality(self, mdct_amplitudes)      # differentiable layer (the reference's @tf.function, :102)
    a, back = adopt(mdct_amplitudes, "mdct_amplitudes", dtype=self._dtype)
    self._check_amplitudes(a)
    b, m, _, c = a.shape
    ton = torch.empty((b, m, 1, c), dtype=self._dtype, device=a.device)
    with torch.cuda.device(a.device):
      if self._bf16:
        work = bf16_workspace(_capi.lib(), a.numel(), ton.numel(), a.device)
        _capi.check(_capi.lib().ac_pa_tonality_bf16(self._plan(a.device), a.data_ptr(), ton.data_ptr(), b, m, c,
                                                    work.data_ptr(), stream_ptr(a.device)))
      else:
        tonality = _capi.lib().ac_pa_tonality_f64 if self._f64 else _capi.lib().ac_pa_tonality_f32
        _capi.check(tonality(self._plan(a.device), a.data_ptr(), ton.data_ptr(), b, m, c, stream_ptr(a.device)))
    return back(ton)

  def global_masking_threshold(self, mdct_amplitudes, tonality_per_block, drown=0.0):
    """Masking threshold amplitude per filter band (psychoacoustic.py:122-148).

    :param tonality_per_block: [batches_n, blocks_n, 1, channels_n]; pass None to fuse the tonality of
                               mdct_amplitudes into the same kernel (extension over the reference)
    :param drown:              0..1, python float
    :return:                   [batches_n, blocks_n, filter_bands_n, channels_n], never below 1e-7
    """
    if tonality_per_block is not None and self.compute_dtype == "float32" and \
        autograd.wants_grad(mdct_amplitudes, tonality_per_block):
      return autograd.global_masking_threshold(self, mdct_amplitudes, tonality_per_block, drown)
    a, back = adopt(mdct_amplitudes, "mdct_amplitudes", dtype=self._dtype)
    self._check_amplitudes(a)
    b, m, _, c = a.shape
    ton_ptr = None
    if tonality_per_block is not None:
      ton, _ = adopt(tonality_per_block, "tonality_per_block", dtype=self._dtype)
      if tuple(ton.shape) != (b, m, 1, c):
        raise ValueError(f"tonality_per_block must be [{b}, {m}, 1, {c}], got {tuple(ton.shape)}")
      ton_ptr = ton.data_ptr()
    thr = torch.empty_like(a)
    with torch.cuda.device(a.device):
      if self._bf16:
        work = bf16_workspace(_capi.lib(), a.numel() + ((b * m * c + 3) & ~3), a.numel(), a.device)
        _capi.check(_capi.lib().ac_pa_threshold_bf16(self._plan(a.device), a.data_ptr(), ton_ptr, float(drown),
                                                     thr.data_ptr(), b, m, c, work.data_ptr(), stream_ptr(a.device)))
      else:
        threshold = _capi.lib().ac_pa_threshold_f64 if self._f64 else _capi.lib().ac_pa_threshold_f32
        _capi.check(threshold(self._plan(a.device), a.data_ptr(), ton_ptr, float(drown), thr.data_ptr(), b, m, c,
                              stream_ptr(a.device)))
    return back(thr)

  def add_noise(self, mdct_amplitudes, masking_threshold, seed=None):
    """mdct_amplitudes + masking_threshold * N(0, 1/6) (psychoacoustic.py:150-167), Philox counter RNG."""
    if self._f64 or self._bf16:
      raise NotImplementedError("add_noise is built for float32 only")
    a, back = adopt(mdct_amplitudes, "mdct_amplitudes")
    thr, _ = adopt(masking_threshold, "masking_threshold")
    if a.shape != thr.shape:
      raise ValueError("masking_threshold must have the shape of mdct_amplitudes")
    if seed is None:
      seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
      _capi.check(_capi.lib().ac_pa_add_noise_f32(a.data_ptr(), thr.data_ptr(), out.data_ptr(), a.numel(),
                                                  int(seed), stream_ptr(a.device)))
    return back(out)

  # ---- quantiser (build-defined: the reference has none) ---------------------------
  def quantize(self, mdct_amplitudes, masking_threshold):
    """q = rint(A / thr), int32 (IEEE divide, round-half-even)."""
    if self._bf16:
      raise NotImplementedError("the quantiser (build-defined, no reference symbol) is built for float32 and float64")
    a, _ = adopt(mdct_amplitudes, "mdct_amplitudes", dtype=self._dtype)
    thr, _ = adopt(masking_threshold, "masking_threshold", dtype=self._dtype)
    if a.shape != thr.shape:
      raise ValueError("masking_threshold must have the shape of mdct_amplitudes")
    q = torch.empty(a.shape, dtype=torch.int32, device=a.device)
    with torch.cuda.device(a.device):
      quantize = _capi.lib().ac_quantize_f64 if self._f64 else _capi.lib().ac_quantize_f32
      _capi.check(quantize(a.data_ptr(), thr.data_ptr(), q.data_ptr(), a.numel(), stream_ptr(a.device)))
    return q

  def dequantize(self, q, masking_threshold):
    """A_hat = q * thr."""
    if self._bf16:
      raise NotImplementedError("the quantiser (build-defined, no reference symbol) is built for float32 and float64")
    q, _ = adopt(q, "q", dtype=torch.int32)
    thr, back = adopt(masking_threshold, "masking_threshold", dtype=self._dtype)
    if q.shape != thr.shape:
      raise ValueError("masking_threshold must have the shape of q")
    out = torch.empty_like(thr)
    with torch.cuda.device(q.device):
      dequantize = _capi.lib().ac_dequantize_f64 if self._f64 else _capi.lib().ac_dequantize_f32
      _capi.check(dequantize(q.data_ptr(), thr.data_ptr(), out.data_ptr(), q.numel(), stream_ptr(q.device)))
    return back(out)

  def encode(self, mdct_amplitudes, drown=0.0, thr_scale=1.0, return_threshold=True):
    """Encoder fusion: tonality -> threshold -> q = rint(A / (thr_scale * thr)) in one pass over A.

    :return: (q int32, step float32) with step = thr_scale * threshold, or q alone.
    """
    if self._bf16:
      raise NotImplementedError("the quantiser (build-defined, no reference symbol) is built for float32 and float64")
    if self._f64:                             # no fused float64 kernel: threshold (internal tonality), then quantise
      if float(thr_scale) != 1.0:
        raise NotImplementedError("thr_scale != 1 is built for float32 only")
      thr = self.global_masking_threshold(mdct_amplitudes, None, drown=drown)
      q = self.quantize(mdct_amplitudes, thr)
      return (q, thr) if return_threshold else q
    a, _ = adopt(mdct_amplitudes, "mdct_amplitudes")
    self._check_amplitudes(a)
    b, m, _, c = a.shape
    q = torch.empty(a.shape, dtype=torch.int32, device=a.device)
    thr = torch.empty_like(a) if return_threshold else None
    with torch.cuda.device(a.device):
      _capi.check(_capi.lib().ac_pa_encode_f32(self._plan(a.device), a.data_ptr(), float(drown), float(thr_scale),
                                               thr.data_ptr() if thr is not None else None, q.data_ptr(), b, m, c,
                                               stream_ptr(a.device)))
    return (q, thr) if return_threshold else q

  # ---- compact side information (decoder-side mapping: psychoacoustic.py:330-331) ----
  def encode_compact(self, mdct_amplitudes, drown=0.0, thr_scale=1.0):
    """Like encode(), but the side information is the 64 bark-domain thresholds of every (frame, channel) instead of
    one step per coefficient: N / 64 times fewer floats to store next to q.

    :return: (q int32 [B, M, N, C], bark_thr float32 [B, M, 64, C]); expand_threshold(bark_thr, thr_scale) is
             bit-identical to the step of encode().
    """
    if self._f64:
      raise NotImplementedError("compact side information is built for float32 only")
    a, _ = adopt(mdct_amplitudes, "mdct_amplitudes")
    self._check_amplitudes(a)
    b, m, _, c = a.shape
    q = torch.empty(a.shape, dtype=torch.int32, device=a.device)
    bark = torch.empty((b, m, 64, c), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
      _capi.check(_capi.lib().ac_pa_encode_compact_f32(self._plan(a.device), a.data_ptr(), float(drown), float(thr_scale),
                                                       bark.data_ptr(), q.data_ptr(), b, m, c, stream_ptr(a.device)))
    return q, bark

  def expand_threshold(self, bark_thr, thr_scale=1.0):
    """Decoder side of encode_compact: step [B, M, N, C] = sqrt(bark_thr W_inv) (psychoacoustic.py:330-331). thr_scale
    must be the encoder's (it only enters through the eps floor of the mapping)."""
    if self._f64:
      raise NotImplementedError("compact side information is built for float32 only")
    g, back = adopt(bark_thr, "bark_thr")
    if g.dim() != 4 or g.shape[2] != 64:
      raise ValueError("bark_thr must be [batches, blocks, 64, channels]")
    b, m, _, c = g.shape
    thr = torch.empty((b, m, self.filter_bands_n, c), dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
      _capi.check(_capi.lib().ac_pa_expand_threshold_f32(self._plan(g.device), g.data_ptr(), float(thr_scale),
                                                         thr.data_ptr(), b, m, c, stream_ptr(g.device)))
    return back(thr)

  # ---- bitstream statistics and rate loop (no reference symbol) ---------------
  def bit_estimate(self, q):
    """{coefficients, nonzero, bits}: bits = sum log2(2|q| + 1), accumulated on the device in 16.16 fixed point."""
    q, _ = adopt(q, "q", dtype=torch.int32)
    stats = torch.zeros(3, dtype=torch.int64, device=q.device)
    with torch.cuda.device(q.device):
      _capi.check(_capi.lib().ac_codec_stats_i32(q.data_ptr(), q.numel(), stats.data_ptr(), stream_ptr(q.device)))
    n, nonzero, bits = (int(v) for v in stats.cpu())
    return {"coefficients": n, "nonzero": nonzero, "bits": bits / 65536.0}

  def encode_at_bitrate(self, mdct_amplitudes, bits_per_coefficient, drown=0.0, iterations=14):
    """Rate loop: the quantiser step is the masking threshold times ONE scalar ("fixed bitrate");
    the scalar is found by bisection on log2(scale) so that the bit estimate meets the target from below.

    :return: (q, step, thr_scale); bit_estimate(q)["bits"] <= bits_per_coefficient * q.numel()
    """
    if self.compute_dtype != "float32":
      raise NotImplementedError("encode_at_bitrate: the rate loop is built on the float32 fused encoder only")
    target = float(bits_per_coefficient) * mdct_amplitudes.numel()
    lo, hi = -10.0, 10.0                      # log2(scale): bits fall monotonically as the scale grows

    def bits_at(log2_scale):
      return self.bit_estimate(self.encode(mdct_amplitudes, drown=drown, thr_scale=2.0 ** log2_scale,
                                           return_threshold=False))["bits"]

    # both ends of the bracket first: widen it (up to 2^+-40) when the target lies outside, refuse what no scale reaches
    while bits_at(hi) > target:
      if hi >= 40.0:
        raise ValueError(f"encode_at_bitrate: {bits_per_coefficient} bits per coefficient is not reachable "
                         f"(even thr_scale = 2^{hi:.0f} needs more)")
      lo, hi = hi, hi + 10.0
    while lo > -40.0 and bits_at(lo) <= target:
      lo, hi = lo - 10.0, lo                  # the finest bracket end already fits: move towards finer steps
    for _ in range(int(iterations)):
      mid = 0.5 * (lo + hi)
      q = self.encode(mdct_amplitudes, drown=drown, thr_scale=2.0 ** mid, return_threshold=False)
      if self.bit_estimate(q)["bits"] > target:
        lo = mid
      else:
        hi = mid
    scale = 2.0 ** hi
    q, step = self.encode(mdct_amplitudes, drown=drown, thr_scale=scale)
    return q, step, scale
