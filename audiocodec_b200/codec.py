"""AudioCodec: the encode -> decode chain the benchmark times, assembled from the two reference classes.

encode: x -> transform -> (tonality, global_masking_threshold, quantise)  -> (q, step)
decode: (q, step) -> inverse_transform(q * step)                           -> x_hat
(the reference stops at the masking threshold / add_noise; the quantiser is build-defined)
"""

import torch

from .mdctransformer import MDCTransformer
from .psychoacoustic import PsychoacousticModel


class AudioCodec:
  def __init__(self, sample_rate, filters_n=1024, bark_bands_n=64, alpha=0.6, window_type='vorbis'):
    self.sample_rate = sample_rate
    self.filters_n = int(filters_n)
    self.mdct = MDCTransformer(filters_n, window_type=window_type)
    self.psychoacoustic = PsychoacousticModel(sample_rate, filter_bands_n=filters_n, bark_bands_n=bark_bands_n, alpha=alpha)

  def encode(self, x, drown=0.0, thr_scale=1.0):
    """x [B, S, C] -> (q int32 [B, S/N + 1, N, C], step float32 same shape)."""
    return self.psychoacoustic.encode(self.mdct.transform(x), drown=drown, thr_scale=thr_scale)

  def decode(self, q, step):
    """(q, step) -> x_hat [B, S + 2 N, C]; x_hat[:, N:-N] reconstructs x (one-block delay, mdctransformer.py:156)."""
    return self.mdct.inverse_transform_dequantized(q, step)

  def roundtrip(self, x, drown=0.0, thr_scale=1.0):
    q, step = self.encode(x, drown=drown, thr_scale=thr_scale)
    return self.decode(q, step), q, step

  @staticmethod
  def stats(q):
    """Per-shard bitstream statistics gathered across ranks at the end of a job."""
    qa = q.abs().to(torch.float32)
    return torch.stack([torch.tensor(float(q.numel()), device=q.device),
                        (qa > 0).sum().to(torch.float32),
                        torch.log2(2.0 * qa + 1.0).sum()])
