"""MDCTransformer: drop-in for /root/reference/audiocodec/mdctransformer.py (class at :12-368).

Same constructor keywords, attributes (filters_n, window_type, H, H_inv), method names, tensor layouts and
error behaviour; the work is done by hand-written sm_100a kernels behind the C ABI
(include/audiocodec_b200.h), never by framework ops and never on the CPU.
"""

import ctypes

import numpy as np
import torch

from . import _capi, autograd
from ._tensors import adopt, bf16_workspace, normalise_compute_dtype, normalise_precompute_dtype, stream_ptr, torch_dtype


class MDCTransformer:
  def __init__(self, filters_n=1024, window_type='vorbis', compute_dtype='float32', precompute_dtype='float64'):
    """Same arguments as the reference (mdctransformer.py:13-14).

    :param filters_n:        number of filter bands (needs to be even; AssertionError otherwise, :26)
    :param window_type:      'sine', 'vorbis' (default); any other string selects the rectangular window (:199-211)
    :param compute_dtype:    dtype of inputs and outputs (tf / torch / numpy dtype or string): float32 (the tuned
                             tile kernels), float64 (functional kernels) or bfloat16 (H / H_inv and the scale constants
                             cast to bfloat16 as in the reference, float32 arithmetic in between - the reference's own
                             rule around the DCT, :326-344 - one rounding to bfloat16 at the output)
    :param precompute_dtype: float64 (default) or float32 for the window tables (:58-59)
    """
    assert (filters_n % 2) == 0, "number of filters used in mdct transformation needs to be even"
    self.filters_n = int(filters_n)
    self.window_type = window_type
    self.compute_dtype = normalise_compute_dtype(compute_dtype, "MDCTransformer")
    self._dtype = torch_dtype(self.compute_dtype)
    self._sfx = {"float64": "f64", "bfloat16": "bf16"}.get(self.compute_dtype, "f32")
    self._bf16 = self.compute_dtype == "bfloat16"
    self._precompute_f32 = int(normalise_precompute_dtype(precompute_dtype) == "float32")
    kind = window_type.lower()   # window_type=None fails here exactly like the reference (:199)
    self._window_code = {"sine": _capi.WINDOW_SINE, "vorbis": _capi.WINDOW_VORBIS}.get(kind, _capi.WINDOW_ONES)
    h = self.filters_n // 2
    fold = np.empty(4 * h, dtype=np.float64)
    unfold = np.empty(4 * h, dtype=np.float64)
    _capi.check(_capi.lib().ac_mdct_tables_host(
      self.filters_n, self._window_code, self._precompute_f32,
      fold.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), unfold.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
    self._fold = fold.reshape(h, 4)
    self._unfold = unfold.reshape(h, 4)
    self._plans = {}     # cuda device index -> ac_mdct_plan*
    self._dense = None

  # ---- reference attributes ---------------------------------------------------------------------------
  def _dense_taps(self):
    """Dense [2, N, N] H / H_inv, only materialised when a caller reads the attributes (:58-59)."""
    if self._dense is None:
      n, h = self.filters_n, self.filters_n // 2
      p = np.arange(h)
      H = np.zeros((2, n, n))
      H[1, p, h - 1 - p] = self._fold[:, 0]
      H[1, n - 1 - p, h - 1 - p] = self._fold[:, 1]
      H[0, p, h + p] = self._fold[:, 2]
      H[0, n - 1 - p, h + p] = self._fold[:, 3]
      H_inv = np.zeros((2, n, n))
      H_inv[0, h - 1 - p, p] = self._unfold[:, 0]
      H_inv[1, h + p, p] = self._unfold[:, 1]
      H_inv[0, h - 1 - p, n - 1 - p] = self._unfold[:, 2]
      H_inv[1, h + p, n - 1 - p] = self._unfold[:, 3]
      self._dense = (torch.from_numpy(H.astype(np.float32)), torch.from_numpy(H_inv.astype(np.float32)))
      if self._bf16:       # tf.cast(H, bfloat16) (:58-59)
        self._dense = tuple(t.to(torch.bfloat16) for t in self._dense)
    return self._dense

  @property
  def H(self):
    return self._dense_taps()[0]

  @property
  def H_inv(self):
    return self._dense_taps()[1]

  # ---- plans ------------------------------------------------------------------------------------------
  def _plan(self, device):
    index = device.index if device.index is not None else torch.cuda.current_device()
    plan = self._plans.get(index)
    if plan is None:
      handle = ctypes.c_void_p()
      with torch.cuda.device(index):
        _capi.check(_capi.lib().ac_mdct_plan_create_ex(self.filters_n, self._window_code, self._precompute_f32,
                                                       _capi.DTYPE_BF16 if self._bf16 else _capi.DTYPE_F32,
                                                       ctypes.byref(handle)))
      plan = self._plans[index] = handle
    return plan

  def __del__(self):
    for handle in getattr(self, "_plans", {}).values():
      try:
        _capi.lib().ac_mdct_plan_destroy(handle)
      except Exception:   # interpreter shutdown
        pass

  # ---- data path --------------------------------------------------------------------------------------
  def transform(self, x):
    """MDCT analysis filter bank (mdctransformer.py:61-125).

    :param x: signal in -1..1, [batches_n, samples_n, channels_n], float32, on a CUDA device
    :return:  [batches_n, blocks_n + 1, filters_n, channels_n] with samples_n = blocks_n * filters_n
    :raises ValueError: samples_n is not a multiple of filters_n (the reference raises InvalidArgumentError, :287)
    """
    if autograd.wants_grad(x) and self.compute_dtype == "float32":
      return autograd.transform(self, x)          # differentiable layer (the reference's @tf.function, :61)
    x, back = adopt(x, "x", dtype=self._dtype)
    if x.dim() != 3:
      raise ValueError(f"x must be [batches_n, samples_n, channels_n], got shape {tuple(x.shape)}")
    b, s, c = x.shape
    if s % self.filters_n != 0:
      raise ValueError(f"samples_n ({s}) must be a multiple of filters_n ({self.filters_n})")
    y = torch.empty((b, s // self.filters_n + 1, self.filters_n, c), dtype=self._dtype, device=x.device)
    with torch.cuda.device(x.device):
      forward = getattr(_capi.lib(), "ac_mdct_forward_" + self._sfx)
      if self._bf16:
        work = bf16_workspace(_capi.lib(), x.numel(), y.numel(), x.device)
        _capi.check(forward(self._plan(x.device), x.data_ptr(), y.data_ptr(), b, s, c, work.data_ptr(), stream_ptr(x.device)))
      else:
        _capi.check(forward(self._plan(x.device), x.data_ptr(), y.data_ptr(), b, s, c, stream_ptr(x.device)))
    return back(y)

  def inverse_transform(self, mdct_amplitudes):
    """MDCT synthesis filter bank with TDAC overlap-add (mdctransformer.py:127-153).

    :param mdct_amplitudes: [batches_n, blocks_n, filters_n, channels_n], float32, CUDA
    :return:                [batches_n, (blocks_n + 1) * filters_n, channels_n]
    """
    if autograd.wants_grad(mdct_amplitudes) and self.compute_dtype == "float32":
      return autograd.inverse_transform(self, mdct_amplitudes)
    y, back = adopt(mdct_amplitudes, "mdct_amplitudes", dtype=self._dtype)
    if y.dim() != 4 or y.shape[2] != self.filters_n:
      raise ValueError(f"mdct_amplitudes must be [batches_n, blocks_n, {self.filters_n}, channels_n], got {tuple(y.shape)}")
    b, m, n, c = y.shape
    x = torch.empty((b, (m + 1) * n, c), dtype=self._dtype, device=y.device)
    with torch.cuda.device(y.device):
      inverse = getattr(_capi.lib(), "ac_mdct_inverse_" + self._sfx)
      if self._bf16:
        work = bf16_workspace(_capi.lib(), y.numel(), x.numel(), y.device)
        _capi.check(inverse(self._plan(y.device), y.data_ptr(), x.data_ptr(), b, m, c, work.data_ptr(), stream_ptr(y.device)))
      else:
        _capi.check(inverse(self._plan(y.device), y.data_ptr(), x.data_ptr(), b, m, c, stream_ptr(y.device)))
    return back(x)

  def inverse_transform_compact(self, q, bark_thr, psychoacoustic, thr_scale=1.0):
    """Decoder fusion on the compact side information of PsychoacousticModel.encode_compact: the quantiser steps are
    rebuilt from the 64 bark-domain thresholds per (frame, channel) inside the dequantising inverse kernel. Bit-identical
    to inverse_transform_dequantized(q, psychoacoustic.expand_threshold(bark_thr, thr_scale)).

    :param q:        int32 [batches_n, blocks_n, filters_n, channels_n]
    :param bark_thr: float32 [batches_n, blocks_n, 64, channels_n]
    """
    q, _ = adopt(q, "q", dtype=torch.int32)
    g, back = adopt(bark_thr, "bark_thr")
    if q.dim() != 4 or q.shape[2] != self.filters_n or g.dim() != 4 or \
        g.shape != (q.shape[0], q.shape[1], 64, q.shape[3]):
      raise ValueError("q must be [batches_n, blocks_n, filters_n, channels_n] and bark_thr [batches_n, blocks_n, 64, channels_n]")
    if self.compute_dtype == "float64":
      raise NotImplementedError("compact side information is built for float32 only")
    b, m, n, c = q.shape
    x = torch.empty((b, (m + 1) * n, c), dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
      _capi.check(_capi.lib().ac_mdct_inverse_dequant_compact_f32(
          self._plan(q.device), psychoacoustic._plan(q.device), q.data_ptr(), g.data_ptr(), float(thr_scale), x.data_ptr(),
          b, m, c, stream_ptr(q.device)))
    return back(x)

  def inverse_transform_dequantized(self, q, masking_threshold):
    """Decoder fusion: inverse_transform(q * masking_threshold) in one kernel (no reference symbol).

    :param q:                 int32 quantised amplitudes [batches_n, blocks_n, filters_n, channels_n]
    :param masking_threshold: float32 quantiser step, same shape
    """
    q, _ = adopt(q, "q", dtype=torch.int32)
    thr, back = adopt(masking_threshold, "masking_threshold", dtype=self._dtype)
    if q.dim() != 4 or q.shape[2] != self.filters_n or q.shape != thr.shape:
      raise ValueError("q and masking_threshold must both be [batches_n, blocks_n, filters_n, channels_n]")
    b, m, n, c = q.shape
    if self.compute_dtype == "float64":       # two kernels: dequantise, then the float64 inverse transform
      amplitudes = torch.empty_like(thr)
      with torch.cuda.device(q.device):
        _capi.check(_capi.lib().ac_dequantize_f64(q.data_ptr(), thr.data_ptr(), amplitudes.data_ptr(), q.numel(),
                                                  stream_ptr(q.device)))
      return back(self.inverse_transform(amplitudes))
    x = torch.empty((b, (m + 1) * n, c), dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
      _capi.check(_capi.lib().ac_mdct_inverse_dequant_f32(self._plan(q.device), q.data_ptr(), thr.data_ptr(),
                                                          x.data_ptr(), b, m, c, stream_ptr(q.device)))
    return back(x)
