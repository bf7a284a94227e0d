"""AudioCodec: the encode -> decode chain the benchmark times, assembled from the two reference classes.

encode: x -> transform -> (tonality, global_masking_threshold, quantise)  -> (q, step)
decode: (q, step) -> inverse_transform(q * step)                           -> x_hat
(the reference stops at the masking threshold / add_noise; the quantiser is build-defined)
"""

import ctypes

import torch

from . import _capi
from ._tensors import adopt, stream_ptr
from .mdctransformer import MDCTransformer
from .psychoacoustic import PsychoacousticModel


class AudioCodec:
  def __init__(self, sample_rate, filters_n=1024, bark_bands_n=64, alpha=0.6, window_type='vorbis'):
    self.sample_rate = sample_rate
    self.filters_n = int(filters_n)
    self.mdct = MDCTransformer(filters_n, window_type=window_type)
    self.psychoacoustic = PsychoacousticModel(sample_rate, filter_bands_n=filters_n, bark_bands_n=bark_bands_n, alpha=alpha)
    self._pipes = {}

  def encode(self, x, drown=0.0, thr_scale=1.0, compact=False, return_threshold=True):
    """x [B, S, C] -> (q int32 [B, S/N + 1, N, C], step float32 same shape).

    One call of the library's single-pass encoder (ac_codec_encode_f32): for stereo signals with filters_n = 256 the
    forward MDCT, the masking model and the quantiser run in ONE kernel and the amplitudes never reach global memory;
    other shapes run transform + encode through a scratch tensor.  Bit-identical to
    psychoacoustic.encode(mdct.transform(x)) either way.  compact=True returns the bark-domain thresholds
    [B, S/N + 1, 64, C] instead of the steps (decode with decode_compact); return_threshold=False returns q alone.
    """
    xt, back = adopt(x, "x")
    if xt.dim() != 3:
      raise ValueError("x must be [batches_n, samples_n, channels_n]")
    b, s, c = xt.shape
    n = self.filters_n
    if s % n != 0:
      raise ValueError(f"samples_n ({s}) must be a multiple of filters_n ({n})")
    frames = s // n + 1
    dev = xt.device
    lib = _capi.lib()
    with torch.cuda.device(dev):
      mplan, pplan = self.mdct._plan(dev), self.psychoacoustic._plan(dev)
      q = torch.empty((b, frames, n, c), dtype=torch.int32, device=dev)
      side = None
      if compact:
        side = torch.empty((b, frames, 64, c), dtype=torch.float32, device=dev)
      elif return_threshold:
        side = torch.empty((b, frames, n, c), dtype=torch.float32, device=dev)
      need = lib.ac_codec_encode_workspace_bytes(mplan, pplan, b, s, c)
      if need < 0:
        raise ValueError("invalid shape for the encoder")
      work = torch.empty(need // 4, dtype=torch.float32, device=dev) if need > 0 else None
      _capi.check(lib.ac_codec_encode_f32(
        mplan, pplan, xt.data_ptr(), float(drown), float(thr_scale),
        side.data_ptr() if (side is not None and not compact) else None,
        side.data_ptr() if compact else None, q.data_ptr(), b, s, c,
        work.data_ptr() if work is not None else None, stream_ptr(dev)))
    if side is None:
      return back(q)
    return back(q), back(side)

  def decode_compact(self, q, bark_thr, thr_scale=1.0):
    """(q, bark-domain thresholds of encode(compact=True)) -> x_hat; the steps are rebuilt inside the inverse MDCT."""
    return self.mdct.inverse_transform_compact(q, bark_thr, self.psychoacoustic, thr_scale=thr_scale)

  def decode(self, q, step):
    """(q, step) -> x_hat [B, S + 2 N, C]; x_hat[:, N:-N] reconstructs x (one-block delay, mdctransformer.py:156)."""
    return self.mdct.inverse_transform_dequantized(q, step)

  def roundtrip(self, x, drown=0.0, thr_scale=1.0):
    q, step = self.encode(x, drown=drown, thr_scale=thr_scale)
    return self.decode(q, step), q, step

  # ---- host-buffer streaming ----------------------------------------------------------------------------
  def roundtrip_host(self, x_host, out_host=None, drown=0.0, thr_scale=1.0, device=None, chunk_clips=None,
                     return_stats=False):
    """encode + decode of clips that live in HOST memory: x_host [B, S, C] -> out_host [B, S + 2N, C].

    The batch is cut into chunks of clips that flow through three streams (H2D copy, kernels, D2H copy)
    over a ring of device buffers inside the C library (ac_codec_roundtrip_host_f32), so both PCIe
    directions and the kernels overlap.  Pinned host tensors give asynchronous copies; pageable ones work
    but serialise.  Returns out_host once it is complete (and, with return_stats, the shard's
    [coefficients, non-zero integers, sum log2(2|q|+1)]).
    """
    if x_host.is_cuda or x_host.dtype != torch.float32 or x_host.dim() != 3:
      raise TypeError("x_host must be a float32 host tensor [batches_n, samples_n, channels_n]")
    x_host = x_host.contiguous()
    b, s, c = x_host.shape
    n = self.filters_n
    if s % n != 0:
      raise ValueError(f"samples_n ({s}) must be a multiple of filters_n ({n})")
    frames = s // n + 1
    if out_host is None:
      out_host = torch.empty((b, (frames + 1) * n, c), dtype=torch.float32, pin_memory=True)
    elif (tuple(out_host.shape) != (b, (frames + 1) * n, c) or out_host.dtype != torch.float32 or out_host.is_cuda
          or not out_host.is_contiguous()):
      raise ValueError("out_host must be a contiguous float32 host tensor [batches_n, samples_n + 2 filters_n, channels_n]")
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if chunk_clips is None:   # ~30 MB of input per chunk measured best on B200 / PCIe 5 (tools/e2e_chunks.py)
      chunk_clips = max(1, min(max(b, 1), (30 << 20) // max(1, 4 * s * c)))
    stats = (ctypes.c_double * 3)() if return_stats else None
    with torch.cuda.device(device):
      pipe = self._pipe(device, int(chunk_clips), s, c)
      _capi.check(_capi.lib().ac_codec_roundtrip_host_f32(
        pipe, x_host.data_ptr(), out_host.data_ptr(), b, float(drown), float(thr_scale), stats,
        torch.cuda.current_stream(device).cuda_stream))
    if return_stats:
      return out_host, torch.tensor(list(stats), dtype=torch.float64)
    return out_host

  def _pipe(self, device, chunk_clips, s, c):
    key = (device.index, chunk_clips, s, c)
    pipe = self._pipes.get(key)
    if pipe is None:
      self._drop_pipes()       # one resident workspace: a new shape replaces the old ring
      handle = ctypes.c_void_p()
      _capi.check(_capi.lib().ac_codec_pipeline_create(self.mdct._plan(device), self.psychoacoustic._plan(device),
                                                       chunk_clips, s, c, ctypes.byref(handle)))
      pipe = self._pipes[key] = handle
    return pipe

  def _drop_pipes(self):
    for handle in getattr(self, "_pipes", {}).values():
      try:
        _capi.lib().ac_codec_pipeline_destroy(handle)
      except Exception:   # interpreter shutdown
        pass
    self._pipes = {}

  def __del__(self):
    self._drop_pipes()

  # ---- entropy-coded bitstream (no reference symbol) -----------------------------------
  def pack(self, q):
    """q int32 [B, F, N, C] -> (stream uint8 [bytes + 4], offsets int64 [B F + 1]): adaptive Golomb-Rice, one
    independent, 4-byte aligned byte range per frame (ac_entropy_plan_i32 + ac_entropy_encode_i32).  offsets[-1] is the
    size of the stream in bytes; the four bytes behind it are read-ahead slack of the decoder."""
    q, _ = adopt(q, "q", dtype=torch.int32)
    if q.dim() != 4:
      raise ValueError("q must be [batches_n, blocks_n, filters_n, channels_n]")
    rows, row_len = q.shape[0] * q.shape[1], q.shape[2] * q.shape[3]
    lib = _capi.lib()
    with torch.cuda.device(q.device):
      offsets = torch.empty(rows + 1, dtype=torch.int64, device=q.device)
      _capi.check(lib.ac_entropy_plan_i32(q.data_ptr(), rows, row_len, offsets.data_ptr(), stream_ptr(q.device)))
      total = int(offsets[-1].item())          # the one host read: the size of the allocation
      stream = torch.zeros(total + 4, dtype=torch.uint8, device=q.device)
      _capi.check(lib.ac_entropy_encode_i32(q.data_ptr(), rows, row_len, offsets.data_ptr(), stream.data_ptr(),
                                            stream_ptr(q.device)))
    return stream, offsets

  def unpack(self, stream, offsets, shape):
    """(stream, offsets) of pack -> q int32 of `shape` [B, F, N, C]."""
    b, f, n, c = (int(v) for v in shape)
    rows, row_len = b * f, n * c
    if offsets.numel() != rows + 1 or offsets.dtype != torch.int64 or stream.dtype != torch.uint8:
      raise ValueError("offsets must be int64 [rows + 1] and stream uint8")
    q = torch.empty((b, f, n, c), dtype=torch.int32, device=stream.device)
    with torch.cuda.device(stream.device):
      _capi.check(_capi.lib().ac_entropy_decode_i32(stream.data_ptr(), offsets.data_ptr(), rows, row_len, q.data_ptr(),
                                                    stream_ptr(stream.device)))
    return q

  def stream_bytes(self, q):
    """Size in bytes of the entropy-coded stream pack(q) would write (ac_entropy_plan_i32 alone: sizes and offsets)."""
    q, _ = adopt(q, "q", dtype=torch.int32)
    rows, row_len = q.shape[0] * q.shape[1], q.shape[2] * q.shape[3]
    if row_len % 16 != 0 or rows == 0:
      return 0
    with torch.cuda.device(q.device):
      offsets = torch.empty(rows + 1, dtype=torch.int64, device=q.device)
      _capi.check(_capi.lib().ac_entropy_plan_i32(q.data_ptr(), rows, row_len, offsets.data_ptr(), stream_ptr(q.device)))
    return int(offsets[-1].item())

  def stats(self, q):
    """Per-shard bitstream statistics gathered across ranks at the end of a job: a device tensor
    [coefficients, non-zero integers, sum log2(2|q|+1), bytes of the entropy-coded stream], counted by the library's
    own kernels (ac_codec_stats_i32, ac_entropy_plan_i32)."""
    est = self.psychoacoustic.bit_estimate(q)
    return torch.tensor([float(est["coefficients"]), float(est["nonzero"]), est["bits"], float(self.stream_bytes(q))],
                        dtype=torch.float64, device=q.device)
