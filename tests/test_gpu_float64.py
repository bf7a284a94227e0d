"""GPU parity of the float64 compute dtype with the oracle's float64 "truth" mode.

The reference accepts compute_dtype=tf.float64 for both classes (mdctransformer.py:13-23, psychoacoustic.py:42-44);
tolerances here are those of double arithmetic with a different summation order.
"""

import numpy as np
import pytest
import torch

import audiocodec_b200
from oracle import audiocodec_oracle as oracle

pytestmark = pytest.mark.gpu


def cuda(a):
  return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("n,window,b,blocks,c", [(256, "vorbis", 2, 9, 2), (64, "sine", 3, 5, 1), (100, "vorbis", 1, 4, 3),
                                                  (1024, "vorbis", 1, 3, 2), (16, "rect", 2, 7, 1)])
def test_mdct_float64(n, window, b, blocks, c):
  rng = np.random.default_rng(n + c)
  x = rng.uniform(-1, 1, (b, blocks * n, c))
  ref = oracle.MDCTransformer(n, window_type=window, compute_dtype=np.float64)
  m = audiocodec_b200.MDCTransformer(n, window_type=window, compute_dtype="float64")
  y = m.transform(cuda(x))
  assert y.dtype == torch.float64 and tuple(y.shape) == (b, blocks + 1, n, c)
  y_ref = ref.transform(x)
  assert np.max(np.abs(y.cpu().numpy() - y_ref)) < 1e-13
  xhat = m.inverse_transform(y)
  assert tuple(xhat.shape) == (b, (blocks + 2) * n, c)
  np.testing.assert_allclose(xhat.cpu().numpy(), ref.inverse_transform(y_ref), atol=1e-12)
  assert np.max(np.abs(xhat.cpu().numpy()[:, n:-n] - x)) < 1e-12           # perfect reconstruction (:35-37)
  with pytest.raises(TypeError):
    m.transform(cuda(x.astype(np.float32)))                               # no implicit casting (:22-23)


@pytest.mark.parametrize("sr,n,nb,alpha,b,m,c", [(44100, 256, 64, 0.6, 2, 11, 2), (48000, 1024, 64, 0.6, 1, 5, 1),
                                                   (22050, 128, 24, 0.8, 2, 6, 3), (44100, 100, 40, 0.6, 1, 4, 1)])
def test_psychoacoustic_float64(sr, n, nb, alpha, b, m, c):
  rng = np.random.default_rng(n + nb)
  y = rng.standard_normal((b, m, n, c)) * 10.0 ** rng.uniform(-6, 0, (b, m, n, c))
  y[0, 0] = 0.0
  ref = oracle.PsychoacousticModel(sr, n, nb, alpha, compute_dtype=np.float64)
  pa = audiocodec_b200.PsychoacousticModel(sr, n, nb, alpha, compute_dtype="float64")
  ton = pa.tonality(cuda(y))
  ton_ref = ref.tonality(y)
  assert ton.dtype == torch.float64
  assert np.max(np.abs(ton.cpu().numpy() - ton_ref)) < 1e-12
  for drown in (0.0, 0.5):
    thr = pa.global_masking_threshold(cuda(y), ton, drown=drown)
    thr_ref = ref.global_masking_threshold(y, ton_ref, drown=drown)
    np.testing.assert_allclose(thr.cpu().numpy(), thr_ref, rtol=1e-10)
  fused = pa.global_masking_threshold(cuda(y), None)
  np.testing.assert_allclose(fused.cpu().numpy(), ref.global_masking_threshold(y, ton_ref), rtol=1e-10)
  thr = cuda(thr_ref)
  q = pa.quantize(cuda(y), thr)
  assert q.dtype == torch.int32 and np.array_equal(q.cpu().numpy(), oracle.quantize(y, thr_ref))
  assert np.array_equal(pa.dequantize(q, thr).cpu().numpy(), oracle.dequantize(q.cpu().numpy(), thr_ref))
  q2, step = pa.encode(cuda(y))
  assert torch.equal(q2, pa.quantize(cuda(y), step))


def test_float64_round_trip_chain():
  sr, n, c = 44100, 256, 2
  x = oracle.synthetic_audio(2, 20 * n, c, sr).astype(np.float64)
  mdct = audiocodec_b200.MDCTransformer(n, compute_dtype="float64")
  pa = audiocodec_b200.PsychoacousticModel(sr, n, compute_dtype="float64")
  y = mdct.transform(cuda(x))
  q, step = pa.encode(y)
  xhat = mdct.inverse_transform_dequantized(q, step)
  ref_m = oracle.MDCTransformer(n, compute_dtype=np.float64)
  ref_p = oracle.PsychoacousticModel(sr, n, compute_dtype=np.float64)
  y_ref = ref_m.transform(x)
  thr_ref = ref_p.global_masking_threshold(y_ref, ref_p.tonality(y_ref))
  q_ref = oracle.quantize(y_ref, thr_ref)
  diff = np.abs(q.cpu().numpy().astype(np.int64) - q_ref)
  assert diff.max() <= 1 and np.mean(diff == 0) >= 0.9999
  xhat_ref = ref_m.inverse_transform(oracle.dequantize(q_ref, thr_ref))
  assert np.sqrt(np.mean((xhat.cpu().numpy() - xhat_ref) ** 2)) < 1e-3 * np.sqrt(np.mean(x ** 2))
