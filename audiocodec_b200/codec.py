"""AudioCodec: the encode -> decode chain the benchmark times, assembled from the two reference classes.

encode: x -> transform -> (tonality, global_masking_threshold, quantise)  -> (q, step)
decode: (q, step) -> inverse_transform(q * step)                           -> x_hat
(the reference stops at the masking threshold / add_noise; the quantiser is build-defined)
"""

import torch

from . import _capi
from .mdctransformer import MDCTransformer
from .psychoacoustic import PsychoacousticModel


class AudioCodec:
  def __init__(self, sample_rate, filters_n=1024, bark_bands_n=64, alpha=0.6, window_type='vorbis'):
    self.sample_rate = sample_rate
    self.filters_n = int(filters_n)
    self.mdct = MDCTransformer(filters_n, window_type=window_type)
    self.psychoacoustic = PsychoacousticModel(sample_rate, filter_bands_n=filters_n, bark_bands_n=bark_bands_n, alpha=alpha)
    self._pipes = {}

  def encode(self, x, drown=0.0, thr_scale=1.0):
    """x [B, S, C] -> (q int32 [B, S/N + 1, N, C], step float32 same shape)."""
    return self.psychoacoustic.encode(self.mdct.transform(x), drown=drown, thr_scale=thr_scale)

  def decode(self, q, step):
    """(q, step) -> x_hat [B, S + 2 N, C]; x_hat[:, N:-N] reconstructs x (one-block delay, mdctransformer.py:156)."""
    return self.mdct.inverse_transform_dequantized(q, step)

  def roundtrip(self, x, drown=0.0, thr_scale=1.0):
    q, step = self.encode(x, drown=drown, thr_scale=thr_scale)
    return self.decode(q, step), q, step

  # ---- host-buffer streaming ----------------------------------------------------------------------------
  def roundtrip_host(self, x_host, out_host=None, drown=0.0, thr_scale=1.0, device=None, chunk_clips=None):
    """encode + decode of clips that live in HOST memory: x_host [B, S, C] -> out_host [B, S + 2N, C].

    The batch is cut into chunks of clips that flow through three streams (H2D copy, kernels, D2H copy)
    over a ring of device buffers, so both PCIe directions and the kernels overlap.  Pinned host tensors
    give asynchronous copies; pageable ones work but serialise.  Returns out_host once it is complete.
    """
    if x_host.is_cuda or x_host.dtype != torch.float32 or x_host.dim() != 3:
      raise TypeError("x_host must be a float32 host tensor [batches_n, samples_n, channels_n]")
    b, s, c = x_host.shape
    n = self.filters_n
    if s % n != 0:
      raise ValueError(f"samples_n ({s}) must be a multiple of filters_n ({n})")
    frames = s // n + 1
    if out_host is None:
      out_host = torch.empty((b, (frames + 1) * n, c), dtype=torch.float32, pin_memory=True)
    elif tuple(out_host.shape) != (b, (frames + 1) * n, c) or out_host.dtype != torch.float32 or out_host.is_cuda:
      raise ValueError("out_host must be a float32 host tensor [batches_n, samples_n + 2 filters_n, channels_n]")
    if b == 0:
      return out_host
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if chunk_clips is None:   # ~32 MB of input per chunk: long enough copies, short pipeline fill
      chunk_clips = max(1, min(b, (32 << 20) // (4 * s * c)))
    pipe = self._pipe(device, chunk_clips, s, c)
    lib = _capi.lib()
    mplan, pplan = self.mdct._plan(device), self.psychoacoustic._plan(device)
    entry = torch.cuda.current_stream(device)
    for st in (pipe["h2d"], pipe["run"], pipe["d2h"]):
      st.wait_stream(entry)
    slots = pipe["slots"]
    with torch.cuda.device(device):
      for k, i in enumerate(range(0, b, chunk_clips)):
        j = min(b, i + chunk_clips)
        cb = j - i
        slot = slots[k % len(slots)]
        with torch.cuda.stream(pipe["h2d"]):
          pipe["h2d"].wait_event(slot["x_free"])
          slot["x"][:cb].copy_(x_host[i:j], non_blocking=True)
          slot["x_ready"].record(pipe["h2d"])
        run = pipe["run"]
        run.wait_event(slot["x_ready"])
        sp = run.cuda_stream
        _capi.check(lib.ac_mdct_forward_f32(mplan, slot["x"].data_ptr(), slot["y"].data_ptr(), cb, s, c, sp))
        slot["x_free"].record(run)
        _capi.check(lib.ac_pa_encode_f32(pplan, slot["y"].data_ptr(), float(drown), float(thr_scale),
                                         slot["step"].data_ptr(), slot["q"].data_ptr(), cb, frames, c, sp))
        run.wait_event(slot["out_free"])
        _capi.check(lib.ac_mdct_inverse_dequant_f32(mplan, slot["q"].data_ptr(), slot["step"].data_ptr(),
                                                    slot["xhat"].data_ptr(), cb, frames, c, sp))
        slot["out_ready"].record(run)
        with torch.cuda.stream(pipe["d2h"]):
          pipe["d2h"].wait_event(slot["out_ready"])
          out_host[i:j].copy_(slot["xhat"][:cb], non_blocking=True)
          slot["out_free"].record(pipe["d2h"])
    done = torch.cuda.Event()
    done.record(pipe["d2h"])
    entry.wait_event(done)
    done.synchronize()     # the result is host memory: hand it back complete
    return out_host

  def _pipe(self, device, chunk_clips, s, c):
    key = (device.index, chunk_clips, s, c)
    pipe = self._pipes.get(key)
    if pipe is None:
      n = self.filters_n
      frames = s // n + 1
      with torch.cuda.device(device):
        slots = []
        for _ in range(3):
          slot = {
            "x": torch.empty((chunk_clips, s, c), dtype=torch.float32, device=device),
            "y": torch.empty((chunk_clips, frames, n, c), dtype=torch.float32, device=device),
            "q": torch.empty((chunk_clips, frames, n, c), dtype=torch.int32, device=device),
            "step": torch.empty((chunk_clips, frames, n, c), dtype=torch.float32, device=device),
            "xhat": torch.empty((chunk_clips, (frames + 1) * n, c), dtype=torch.float32, device=device),
            "x_ready": torch.cuda.Event(), "x_free": torch.cuda.Event(),
            "out_ready": torch.cuda.Event(), "out_free": torch.cuda.Event(),
          }
          slots.append(slot)
        pipe = {"h2d": torch.cuda.Stream(device), "run": torch.cuda.Stream(device), "d2h": torch.cuda.Stream(device),
                "slots": slots}
      self._pipes.clear()      # one resident workspace: a new shape replaces the old ring
      self._pipes[key] = pipe
    return pipe

  @staticmethod
  def stats(q):
    """Per-shard bitstream statistics gathered across ranks at the end of a job."""
    qa = q.abs().to(torch.float32)
    return torch.stack([torch.tensor(float(q.numel()), device=q.device),
                        (qa > 0).sum().to(torch.float32),
                        torch.log2(2.0 * qa + 1.0).sum()])
