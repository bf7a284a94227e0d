"""compute_dtype = bfloat16 (psychoacoustic.py:42-44, mdctransformer.py:326-344) and the dB
utilities as device ops (psychoacoustic.py:71-100).

bfloat16 here means: bfloat16 tensors at the boundary, the reference's tables and constants cast to bfloat16, float32
arithmetic in between (the reference's own rule around the DCT) and one rounding at the output.  Two checks:
  * against that definition restated with the float32 oracle on bfloat16-rounded tables: equal up to one bfloat16 ulp;
  * against the oracle's bfloat16 mode, which rounds after every op as TensorFlow would (UNPINNED - TensorFlow's
    bfloat16 kernels were never run): within a few bfloat16 ulps of the largest value.
"""

import math

import ml_dtypes
import numpy as np
import pytest
import torch

import audiocodec_b200
from oracle import audiocodec_oracle as oracle

pytestmark = pytest.mark.gpu
BF = ml_dtypes.bfloat16


def to_bf16_cuda(a):
  return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda().to(torch.bfloat16)


def f32(t):
  return t.float().cpu().numpy()


def rbf(a):
  return np.asarray(a, dtype=np.float32).astype(BF).astype(np.float32)


@pytest.mark.parametrize("n,window,b,blocks,c", [(256, 'vorbis', 2, 40, 2), (1024, 'sine', 1, 9, 2), (64, 'vorbis', 3, 17, 1),
                                                 (100, 'vorbis', 1, 6, 1)])
def test_mdct_bfloat16(n, window, b, blocks, c):
  rng = np.random.default_rng(n)
  x = rbf(rng.uniform(-1, 1, (b, blocks * n, c)))
  mdct = audiocodec_b200.MDCTransformer(n, window_type=window, compute_dtype='bfloat16')
  y = mdct.transform(to_bf16_cuda(x))
  assert y.dtype == torch.bfloat16 and tuple(y.shape) == (b, blocks + 1, n, c)
  # the definition: float32 arithmetic on bfloat16-rounded H, bfloat16 constants sqrt(2) and 1 / sqrt(4N), one rounding
  ref = oracle.MDCTransformer(n, window_type=window, compute_dtype=np.float32)
  ref.H, ref.H_inv = rbf(ref.H), rbf(ref.H_inv)
  s = 1.0 / math.sqrt(4.0 * n)
  k_fwd = float(rbf(math.sqrt(2.0))) / math.sqrt(2.0) * float(rbf(s)) / s
  y_def = ref.transform(x).astype(np.float64) * k_fwd
  d = np.abs(f32(y) - rbf(y_def))
  assert np.all(d <= 2.0 ** -7 * np.abs(y_def) + 1e-7)            # at most one bfloat16 ulp (a tie broken by fp32 rounding)
  assert np.mean(d == 0) > 0.98
  # TensorFlow-style per-op rounding
  y_tf = oracle.MDCTransformer(n, window_type=window, compute_dtype='bfloat16').transform(x.astype(BF)).astype(np.float32)
  assert np.max(np.abs(f32(y) - y_tf)) <= 0.02 * np.max(np.abs(y_tf))
  # inverse
  back = mdct.inverse_transform(y)
  assert back.dtype == torch.bfloat16 and tuple(back.shape) == (b, (blocks + 2) * n, c)
  k_inv = float(rbf(math.sqrt(2.0))) / math.sqrt(2.0) * float(rbf(1.0 / s)) * s
  back_def = ref.inverse_transform(f32(y)).astype(np.float64) * k_inv
  d = np.abs(f32(back) - rbf(back_def))
  assert np.all(d <= 2.0 ** -7 * np.abs(back_def) + 1e-6)
  assert np.max(np.abs(f32(back)[:, n:-n] - x)) < 0.03            # round trip at bfloat16 resolution
  with pytest.raises(TypeError):
    mdct.transform(torch.zeros(1, n, 1, device="cuda"))           # float32 input on a bfloat16 model: no implicit cast


@pytest.mark.parametrize("sr,n,c", [(44100, 256, 2), (48000, 1024, 1)])
def test_psychoacoustic_bfloat16(sr, n, c):
  x = oracle.synthetic_audio(2, 24 * n, c, sr)
  y32 = oracle.MDCTransformer(n).transform(x)
  y = rbf(y32)
  pa = audiocodec_b200.PsychoacousticModel(sr, n, compute_dtype='bfloat16')
  pa32 = audiocodec_b200.PsychoacousticModel(sr, n)
  yb = to_bf16_cuda(y)
  ton = pa.tonality(yb)
  assert ton.dtype == torch.bfloat16 and tuple(ton.shape) == (2, 25, 1, c)
  ton32 = pa32.tonality(yb.float())
  assert np.max(np.abs(f32(ton) - rbf(f32(ton32)))) <= 2.0 ** -7    # eps is the only bfloat16 constant in the tonality
  ref = oracle.PsychoacousticModel(sr, n, compute_dtype='bfloat16')
  ton_tf = ref.tonality(y.astype(BF))
  assert np.max(np.abs(f32(ton) - ton_tf.astype(np.float32))) <= 0.03
  thr = pa.global_masking_threshold(yb, ton)
  assert thr.dtype == torch.bfloat16 and thr.shape == yb.shape
  thr_tf = ref.global_masking_threshold(y.astype(BF), ton_tf).astype(np.float32)
  ratio = f32(thr) / thr_tf
  assert 0.85 < ratio.min() and ratio.max() < 1.15 and abs(np.median(ratio) - 1) < 0.02
  # against the float32 model: alpha = 0.6 -> 0.6016, 1 / alpha = 1.6667 -> 1.6641 in bfloat16 move the threshold by a few %
  ratio32 = f32(thr) / f32(pa32.global_masking_threshold(yb.float(), ton.float()))
  assert 0.8 < ratio32.min() and ratio32.max() < 1.25
  fused = pa.global_masking_threshold(yb, None)                    # internal tonality
  assert np.max(np.abs(f32(fused) / f32(thr) - 1)) < 0.02
  with pytest.raises(NotImplementedError):
    pa.encode(yb)


def test_db_utilities_are_device_ops():
  from audiocodec_b200 import _capi
  rng = np.random.default_rng(1)
  a = (rng.standard_normal(100003) * 10.0 ** rng.uniform(-9, 0, 100003)).astype(np.float32)
  a[:3] = [0.0, 1.0, -1.0]
  pa = audiocodec_b200.PsychoacousticModel(44100, 64)
  ref = oracle.PsychoacousticModel(44100, 64)
  launches0 = _capi.lib().ac_kernel_launch_count()
  db = pa.amplitude_to_dB(torch.from_numpy(a).cuda())
  dbn = pa.amplitude_to_dB_norm(torch.from_numpy(a).cuda())
  assert _capi.lib().ac_kernel_launch_count() - launches0 == 2     # one kernel each
  np.testing.assert_allclose(db.cpu().numpy(), ref.amplitude_to_dB(a), rtol=2e-6, atol=2e-5)
  np.testing.assert_allclose(dbn.cpu().numpy(), ref.amplitude_to_dB_norm(a), rtol=2e-6, atol=1e-6)
  assert db.min().item() >= -20.0 - 1e-4 and abs(db[1].item() - 120.0) < 1e-4 and abs(dbn[0].item()) < 1e-6
  pa64 = audiocodec_b200.PsychoacousticModel(44100, 64, compute_dtype='float64')
  ref64 = oracle.PsychoacousticModel(44100, 64, compute_dtype=np.float64)
  a64 = a.astype(np.float64)
  np.testing.assert_allclose(pa64.amplitude_to_dB(torch.from_numpy(a64).cuda()).cpu().numpy(), ref64.amplitude_to_dB(a64),
                             rtol=1e-13, atol=1e-11)
  np.testing.assert_allclose(pa64.amplitude_to_dB_norm(torch.from_numpy(a64).cuda()).cpu().numpy(),
                             ref64.amplitude_to_dB_norm(a64), rtol=1e-13, atol=1e-13)
  pab = audiocodec_b200.PsychoacousticModel(44100, 64, compute_dtype='bfloat16')
  dbb = pab.amplitude_to_dB(to_bf16_cuda(a))
  assert dbb.dtype == torch.bfloat16
  np.testing.assert_allclose(f32(dbb), ref.amplitude_to_dB(rbf(a)), rtol=2.0 ** -7, atol=0.05)
