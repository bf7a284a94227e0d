#!/usr/bin/env python
"""Generate tests/golden/reference_vectors.npz by executing the UNMODIFIED reference sources.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

TensorFlow cannot be installed here, so the reference modules are imported with oracle/tf_shim on
sys.path: a NumPy implementation of the TF ops they call.  The reference's own unit tests pass under that
shim (tests/test_reference_under_shim.py), including its golden vector, which is what pins the shim.
Every array stored here is an input we made up (seeded) or an output of the reference code on it.
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get("AUDIOCODEC_REFERENCE", "/root/reference")

sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
sys.path.insert(0, REFERENCE)

import tensorflow as tf  # noqa: E402  (the shim)
from audiocodec.mdctransformer import MDCTransformer  # noqa: E402  (the reference)
from audiocodec.psychoacoustic import PsychoacousticModel  # noqa: E402  (the reference)


def sine_wav(amplitude, frequency, sample_rate, duration_sec):
  t = np.arange(0, sample_rate * duration_sec, dtype=np.float32)
  return (amplitude * np.sin(2.0 * np.pi * frequency * t / sample_rate)).reshape(1, -1, 1)


def main():
  rng = np.random.default_rng(20261018)
  out = {}

  # ---- MDCT tables: dense H / H_inv for small N, every window type -------------------------------
  for n, window in [(8, 'vorbis'), (16, 'sine'), (12, 'ones'), (64, 'vorbis')]:
    m = MDCTransformer(n, window_type=window, compute_dtype=tf.float64)
    out[f"H_{n}_{window}"] = np.asarray(m.H)
    out[f"Hinv_{n}_{window}"] = np.asarray(m.H_inv)

  # ---- MDCT data path --------------------------------------------------------------------------------
  cases = [("kat64", 64, 'vorbis', sine_wav(0.8, 4, 64, 4.)[:, :256].astype(np.float32)),
           ("sine256", 256, 'vorbis', sine_wav(0.8, 880, 16000, 1.)[:, :256 * 62].astype(np.float32)),
           ("rand64_c2", 64, 'vorbis', rng.standard_normal((3, 64 * 6, 2)).astype(np.float32)),
           ("rand256_sine_c2", 256, 'sine', rng.uniform(-1, 1, (2, 256 * 9, 2)).astype(np.float32)),
           ("rand1024_c1", 1024, 'vorbis', rng.uniform(-1, 1, (1, 1024 * 4, 1)).astype(np.float32)),
           ("rand12_ones_c3", 12, 'ones', rng.uniform(-1, 1, (2, 12 * 5, 3)).astype(np.float32))]
  for name, n, window, x in cases:
    out[f"mdct_{name}_x"] = x
    for tag, dt in (("f32", tf.float32), ("f64", tf.float64)):
      m = MDCTransformer(n, window_type=window, compute_dtype=dt)
      xin = x.astype(dt)
      y = m.transform(xin)
      out[f"mdct_{name}_{tag}_y"] = np.asarray(y)
      out[f"mdct_{name}_{tag}_xhat"] = np.asarray(m.inverse_transform(y))

  # ---- psychoacoustic tables -------------------------------------------------------------------------
  for sr, n, nb, alpha in [(32768, 64, 64, 0.6), (44100, 256, 64, 0.6), (48000, 1024, 64, 0.6),
                           (16000, 128, 24, 0.8)]:
    pa = PsychoacousticModel(sr, filter_bands_n=n, bark_bands_n=nb, alpha=alpha, compute_dtype=tf.float64)
    key = f"pa_{sr}_{n}_{nb}"
    out[f"{key}_W"] = np.asarray(pa.W)
    out[f"{key}_Winv"] = np.asarray(pa.W_inv)
    out[f"{key}_quiet"] = np.asarray(pa.quiet_threshold_intensity)
    out[f"{key}_S"] = np.asarray(pa.spreading_matrix)
    out[f"{key}_scalars"] = np.asarray([float(pa.max_bark), float(pa.bark_band_width), float(pa._dB_MIN)])

  # ---- psychoacoustic data path: tonality, threshold (drown 0 and 0.35), on MDCT output of tone+noise --
  def tone_noise(b, s, c, sr):
    t = np.arange(s, dtype=np.float64)[None, :, None]
    f = np.asarray([220., 1000., 3520., 9000.])[:b, None, None]
    ph = (np.arange(c) * np.pi / 3.)[None, None, :]
    x = 0.5 * np.sin(2 * np.pi * f * t / sr + ph) + 0.05 * rng.standard_normal((b, s, c))
    return np.clip(x, -1, 1).astype(np.float32)

  pa_cases = [("n256", 44100, 256, 64, 0.6, tone_noise(3, 256 * 12, 2, 44100)),
              ("n1024", 48000, 1024, 64, 0.6, tone_noise(2, 1024 * 5, 1, 48000)),
              ("n64", 32768, 64, 64, 0.6, tone_noise(2, 64 * 10, 2, 32768)),
              ("n128_nb24", 16000, 128, 24, 0.8, tone_noise(2, 128 * 7, 1, 16000))]
  for name, sr, n, nb, alpha, x in pa_cases:
    out[f"pa_{name}_x"] = x
    for tag, dt in (("f32", tf.float32), ("f64", tf.float64)):
      mdct = MDCTransformer(n, compute_dtype=dt)
      pa = PsychoacousticModel(sr, filter_bands_n=n, bark_bands_n=nb, alpha=alpha, compute_dtype=dt)
      y = mdct.transform(x.astype(dt))
      # also feed silence and a digital-silence frame: all four eps clamps get exercised
      y = np.concatenate([y, np.zeros_like(y[:, :1])], axis=1)
      ton = pa.tonality(y)
      out[f"pa_{name}_{tag}_y"] = np.asarray(y)
      out[f"pa_{name}_{tag}_ton"] = np.asarray(ton)
      out[f"pa_{name}_{tag}_thr"] = np.asarray(pa.global_masking_threshold(y, ton))
      out[f"pa_{name}_{tag}_thr_drown"] = np.asarray(pa.global_masking_threshold(y, ton, drown=0.35))
      if name == "n64":
        out[f"pa_{name}_{tag}_dB"] = np.asarray(pa.amplitude_to_dB(y))
        out[f"pa_{name}_{tag}_dBnorm"] = np.asarray(pa.amplitude_to_dB_norm(y))

  path = os.path.join(HERE, "reference_vectors.npz")
  np.savez_compressed(path, **out)
  print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
  main()
