"""The oracle (oracle/audiocodec_oracle.py) against the reference's golden vectors and tests.

CPU only.  Pins: (1) the reference's own unit tests re-expressed (file:line cited per test),
(2) the fixtures in tests/golden/ = outputs of the unmodified reference sources under oracle/tf_shim.
"""

import numpy as np
import pytest

from oracle import audiocodec_oracle as oracle
from conftest import rms

# golden vector, /root/reference/audiocodec/tests/test_mdctransformer.py:51-52
KAT64 = np.asarray([-0.000412722176, 0.000430465181, 0.000789350364, -0.000867388735, -0.00275337417,
                    0.0132110268, 0.0193885863, 0.156005412, -0.233544752, -0.0129148215])
# comment-embedded H example for filters_n = 8, /root/reference/audiocodec/mdctransformer.py:37-56
H8_TAP0 = {(0, 4): 0.99988616, (1, 5): 0.9912527, (2, 6): 0.93969655, (3, 7): 0.80674446,
           (4, 7): -0.5909005, (5, 6): -0.3420093, (6, 5): -0.13197729, (7, 4): -0.01509063}
H8_TAP1 = {(0, 3): 0.01509063, (1, 2): 0.13197729, (2, 1): 0.3420093, (3, 0): 0.5909005,
           (4, 0): 0.80674446, (5, 1): 0.93969655, (6, 2): 0.9912527, (7, 3): 0.99988616}


def sine_wav(amplitude, frequency, sample_rate, duration_sec):
  t = np.arange(0, sample_rate * duration_sec, dtype=np.float32)
  return (amplitude * np.sin(2.0 * np.pi * frequency * t / sample_rate)).astype(np.float32).reshape(1, -1, 1)


# ------------------------------------------------------------------------------ reference unit tests
def test_inverse_identity():
  """test_mdctransformer.py:19-37."""
  n = 256
  mdct = oracle.MDCTransformer(n)
  x = sine_wav(0.8, 880, 16000, 1.)
  x = x[:, 0:n * (x.shape[1] // n)]
  back = mdct.inverse_transform(mdct.transform(x))
  assert back.shape == (1, x.shape[1] + 2 * n, 1)
  assert np.max(np.abs(x - back[:, n:-n])) < 1e-5


@pytest.mark.parametrize("precompute,tol", [(np.float32, 1e-7), (np.float64, 1e-6)])
def test_mdct_calculation(precompute, tol):
  """test_mdctransformer.py:39-54; two-sided, and tighter with the float32 precompute that made the vector."""
  mdct = oracle.MDCTransformer(64, precompute_dtype=precompute)
  x = sine_wav(0.8, 4, 64, 4.)[:, :256]
  y = mdct.transform(x)
  assert np.max(np.abs(y[0, 1, :10, 0] - KAT64)) < tol


def test_mdct_shape():
  """test_mdctransformer.py:56-75."""
  rng = np.random.default_rng(0)
  y = oracle.MDCTransformer(64).transform(rng.standard_normal((128, 640, 2)).astype(np.float32))
  assert y.shape == (128, 11, 64, 2) and y.dtype == np.float32


def test_h_example_n8():
  """mdctransformer.py:37-56."""
  h = oracle.MDCTransformer(8).H
  for tap, entries in ((0, H8_TAP0), (1, H8_TAP1)):
    dense = np.zeros((8, 8))
    for (r, c), v in entries.items():
      dense[r, c] = v
    assert np.max(np.abs(h[tap] - dense)) < 5e-8


def test_energy_conservation():
  """test_psychoacoustic.py:14-30."""
  pa = oracle.PsychoacousticModel(sample_rate=32768, filter_bands_n=64)
  assert np.sum(np.abs(pa.W.sum(axis=1) - 1.0)) < 1e-6
  assert np.sum(np.abs(pa.W_inv.sum(axis=1) - 1.0)) < 1e-6


def test_tonality_tone():
  """test_psychoacoustic.py:32-42."""
  y = oracle.MDCTransformer(64).transform(sine_wav(0.8, 4, 64, 5.))
  ton = oracle.PsychoacousticModel(sample_rate=64, filter_bands_n=64).tonality(y)
  assert ton[0, 1, 0, 0] == 1.0


def test_tonality_noise():
  """test_psychoacoustic.py:44-65."""
  rng = np.random.default_rng(1)
  y = oracle.MDCTransformer(64).transform(rng.uniform(-1, 1, (10, 640, 2)).astype(np.float32))
  ton = oracle.PsychoacousticModel(sample_rate=64, filter_bands_n=64).tonality(y)
  assert ton.shape == (10, 11, 1, 2)
  assert np.mean(ton[0, 1:-1]) < 0.1


# ----------------------------------------------------------------------------------- golden fixtures
@pytest.mark.parametrize("n,window", [(8, 'vorbis'), (16, 'sine'), (12, 'ones'), (64, 'vorbis')])
def test_tables_mdct(golden, n, window):
  m = oracle.MDCTransformer(n, window_type=window, compute_dtype=np.float64)
  assert np.max(np.abs(m.H - golden[f"H_{n}_{window}"])) == 0.0
  assert np.max(np.abs(m.H_inv - golden[f"Hinv_{n}_{window}"])) < 1e-15


MDCT_CASES = [("kat64", 64, 'vorbis'), ("sine256", 256, 'vorbis'), ("rand64_c2", 64, 'vorbis'),
              ("rand256_sine_c2", 256, 'sine'), ("rand1024_c1", 1024, 'vorbis'), ("rand12_ones_c3", 12, 'ones')]


@pytest.mark.parametrize("name,n,window", MDCT_CASES)
def test_mdct_vs_golden(golden, name, n, window):
  x = golden[f"mdct_{name}_x"]
  for tag, dt, tol in (("f64", np.float64, 1e-13), ("f32", np.float32, 2e-6)):
    m = oracle.MDCTransformer(n, window_type=window, compute_dtype=dt)
    y = m.transform(x.astype(dt))
    y_ref = golden[f"mdct_{name}_{tag}_y"]
    assert y.shape == y_ref.shape and y.dtype == y_ref.dtype
    assert np.max(np.abs(y - y_ref)) <= tol * max(1.0, rms(x))
    back = m.inverse_transform(y_ref)
    back_ref = golden[f"mdct_{name}_{tag}_xhat"]
    assert back.shape == back_ref.shape
    assert np.max(np.abs(back - back_ref)) <= 10 * tol * max(1.0, rms(x))


PA_TABLES = [(32768, 64, 64, 0.6), (44100, 256, 64, 0.6), (48000, 1024, 64, 0.6), (16000, 128, 24, 0.8)]


@pytest.mark.parametrize("sr,n,nb,alpha", PA_TABLES)
def test_tables_pa(golden, sr, n, nb, alpha):
  pa = oracle.PsychoacousticModel(sr, n, nb, alpha, compute_dtype=np.float64)
  key = f"pa_{sr}_{n}_{nb}"
  assert np.max(np.abs(pa.W - golden[f"{key}_W"])) < 1e-15
  assert np.max(np.abs(pa.W_inv - golden[f"{key}_Winv"])) < 1e-15
  np.testing.assert_allclose(pa.quiet_threshold_intensity, golden[f"{key}_quiet"], rtol=1e-13)
  np.testing.assert_allclose(pa.spreading_matrix, golden[f"{key}_S"], rtol=1e-12)
  np.testing.assert_allclose([pa.max_bark, pa.bark_band_width, pa._dB_MIN], golden[f"{key}_scalars"], rtol=1e-14)


PA_CASES = [("n256", 44100, 256, 64, 0.6), ("n1024", 48000, 1024, 64, 0.6), ("n64", 32768, 64, 64, 0.6),
            ("n128_nb24", 16000, 128, 24, 0.8)]


@pytest.mark.parametrize("name,sr,n,nb,alpha", PA_CASES)
def test_pa_vs_golden(golden, name, sr, n, nb, alpha):
  for tag, dt, rtol in (("f64", np.float64, 1e-10), ("f32", np.float32, 2e-4)):
    pa = oracle.PsychoacousticModel(sr, n, nb, alpha, compute_dtype=dt)
    y = golden[f"pa_{name}_{tag}_y"]
    ton = pa.tonality(y)
    ton_ref = golden[f"pa_{name}_{tag}_ton"]
    assert ton.shape == ton_ref.shape
    assert np.max(np.abs(ton - ton_ref)) < (1e-12 if dt is np.float64 else 5e-6)  # fp32: summation-order noise
    for key, drown in (("thr", 0.0), ("thr_drown", 0.35)):
      thr = pa.global_masking_threshold(y, ton_ref, drown=drown)
      thr_ref = golden[f"pa_{name}_{tag}_{key}"]
      assert thr.shape == thr_ref.shape and thr.dtype == thr_ref.dtype
      np.testing.assert_allclose(thr, thr_ref, rtol=rtol)
      assert thr.min() >= 1e-7 * (1 - 1e-6)


def test_pa_dense_form_matches(golden):
  """The factored spreading (gain[j] * sum_i P[i] S[i,j]) equals the reference's literal 5-D form."""
  pa = oracle.PsychoacousticModel(32768, 64, compute_dtype=np.float64)
  y, ton = golden["pa_n64_f64_y"], golden["pa_n64_f64_ton"]
  a = pa._masking_intensity_in_bark(y, ton, 0.2)
  b = pa._masking_intensity_dense(y, ton, 0.2)
  np.testing.assert_allclose(a, b, rtol=1e-12)


def test_pa_db_utilities(golden):
  for tag, dt in (("f64", np.float64), ("f32", np.float32)):
    pa = oracle.PsychoacousticModel(32768, 64, compute_dtype=dt)
    y = golden[f"pa_n64_{tag}_y"]
    np.testing.assert_allclose(pa.amplitude_to_dB(y), golden[f"pa_n64_{tag}_dB"], rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(pa.amplitude_to_dB_norm(y), golden[f"pa_n64_{tag}_dBnorm"], rtol=1e-5, atol=1e-6)


def test_fp32_oracle_close_to_truth(golden):
  """The fp32-faithful oracle sits well inside the 1e-5 * RMS budget of the float64 truth."""
  x = golden["mdct_rand256_sine_c2_x"]
  y32 = oracle.MDCTransformer(256, 'sine', np.float32).transform(x)
  y64 = oracle.MDCTransformer(256, 'sine', np.float64).transform(x.astype(np.float64))
  assert np.max(np.abs(y32 - y64)) < 1e-6 * rms(x)


def test_error_behaviour():
  with pytest.raises(AssertionError):
    oracle.MDCTransformer(7)
  with pytest.raises(ValueError):
    oracle.MDCTransformer(8).transform(np.zeros((1, 20, 1), np.float32))
  with pytest.raises(TypeError):
    oracle.MDCTransformer(8).transform(np.zeros((1, 16, 1), np.float64))


def test_quantizer_spec():
  a = np.asarray([0.5, 1.5, 2.5, -0.5, -1.5, 0.26, -3.49], np.float32)
  thr = np.ones_like(a)
  q = oracle.quantize(a, thr)
  assert q.dtype == np.int32
  assert q.tolist() == [0, 2, 2, 0, -2, 0, -3]          # round-half-to-even, like tf.round
  assert oracle.dequantize(q, 0.5 * thr).tolist() == [0, 1, 1, 0, -1, 0, -1.5]


def test_entropy_restatement_round_trip():
  """oracle/entropy_oracle.py (the CPU restatement of the build-defined bitstream): lossless, sizes consistent."""
  from oracle import entropy_oracle as eo
  rng = np.random.default_rng(11)
  q = np.rint(rng.standard_normal((9, 96)) * rng.choice([0, 0.3, 2, 40, 3000], (9, 1))).astype(np.int32)
  q[0] = 0
  q[1, 5], q[2, 7] = 2 ** 31 - 1, -2 ** 31
  stream, offsets = eo.encode(q)
  assert np.array_equal(eo.decode(stream, offsets, 9, 96), q)
  assert np.array_equal(np.diff(offsets), eo.row_sizes(q)) and np.all(np.diff(offsets) % 4 == 0)
  assert offsets[1] - offsets[0] == 4 * ((6 * 5 + 31) // 32)            # an all-zero row: six 5-bit headers
