"""GPU parity tests of the psychoacoustic kernels and the quantiser against the oracle and reference fixtures.

Tolerances (BASELINE.json north_star): masking threshold within 1e-5 relative to signal RMS; quantised
integers bit-exact given the reference's threshold; end to end >= 99.99 % identical, remainder +-1.
"""

import numpy as np
import pytest
import torch

import audiocodec_b200
from audiocodec_b200 import _capi
from oracle import audiocodec_oracle as oracle
from conftest import rms

pytestmark = pytest.mark.gpu


def cuda(a):
  return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def sine_wav(amplitude, frequency, sample_rate, duration_sec):
  t = np.arange(0, sample_rate * duration_sec, dtype=np.float32)
  return (amplitude * np.sin(2.0 * np.pi * frequency * t / sample_rate)).astype(np.float32).reshape(1, -1, 1)


# ---- the reference's own tests, on the CUDA path -----------------------------------------------------------
def test_tonality_tone():
  """audiocodec/tests/test_psychoacoustic.py:32-42."""
  y = audiocodec_b200.MDCTransformer(64).transform(cuda(sine_wav(0.8, 4, 64, 5.)))
  ton = audiocodec_b200.PsychoacousticModel(sample_rate=64, filter_bands_n=64).tonality(y)
  assert ton[0, 1].item() == 1.0


def test_tonality_noise():
  """audiocodec/tests/test_psychoacoustic.py:44-65."""
  x = torch.rand(10, 640, 2, device="cuda") * 2 - 1
  y = audiocodec_b200.MDCTransformer(64).transform(x)
  ton = audiocodec_b200.PsychoacousticModel(sample_rate=64, filter_bands_n=64).tonality(y)
  assert tuple(ton.shape) == (10, 11, 1, 2)
  assert ton[0, 1:-1].mean().item() < 0.1


# ---- reference fixtures ------------------------------------------------------------------------------------
PA_CASES = [("n256", 44100, 256, 64, 0.6), ("n1024", 48000, 1024, 64, 0.6), ("n64", 32768, 64, 64, 0.6),
            ("n128_nb24", 16000, 128, 24, 0.8)]


@pytest.mark.parametrize("name,sr,n,nb,alpha", PA_CASES)
def test_against_reference_fixture(golden, name, sr, n, nb, alpha):
  pa = audiocodec_b200.PsychoacousticModel(sr, n, nb, alpha)
  y = golden[f"pa_{name}_f32_y"]
  y64 = golden[f"pa_{name}_f64_y"]
  assert np.max(np.abs(y - y64)) < 1e-6          # same amplitudes up to fp32 rounding
  signal_rms = rms(golden[f"pa_{name}_x"])
  ton = pa.tonality(cuda(y)).cpu().numpy()
  ton_ref = golden[f"pa_{name}_f64_ton"]
  assert ton.shape == ton_ref.shape
  assert np.max(np.abs(ton - ton_ref)) < 2e-5
  for key, drown in (("thr", 0.0), ("thr_drown", 0.35)):
    thr_ref = golden[f"pa_{name}_f64_{key}"]
    thr = pa.global_masking_threshold(cuda(y), cuda(golden[f"pa_{name}_f32_ton"]), drown=drown).cpu().numpy()
    assert thr.shape == thr_ref.shape
    assert np.max(np.abs(thr - thr_ref)) <= 1e-5 * signal_rms        # the north-star tolerance
    np.testing.assert_allclose(thr, thr_ref, rtol=2e-4)              # and tight relative to thr itself
    assert thr.min() >= 1e-7 * (1 - 1e-6)
    fused = pa.global_masking_threshold(cuda(y), None, drown=drown).cpu().numpy()   # internal tonality
    np.testing.assert_allclose(fused, thr_ref, rtol=2e-4)


@pytest.mark.parametrize("sr,n,nb,alpha,b,m,c", [(44100, 256, 64, 0.6, 3, 17, 2), (48000, 1024, 64, 0.6, 2, 9, 2),
                                                   (22050, 512, 48, 0.5, 2, 5, 1), (44100, 2048, 64, 0.6, 1, 3, 2),
                                                   (8000, 32, 16, 1.0, 2, 6, 3), (44100, 100, 40, 0.6, 2, 4, 1),
                                                   # tile kernel: mono / four channels, ragged last tile, odd N
                                                   (44100, 256, 64, 0.6, 3, 23, 1), (44100, 256, 64, 0.6, 2, 5, 4),
                                                   (48000, 1024, 64, 0.6, 1, 7, 4), (44100, 320, 64, 0.6, 2, 9, 2),
                                                   (44100, 4096, 64, 0.6, 1, 2, 1),
                                                   # tile kernel with the exponent tables (alpha away from 1/2) and with
                                                   # the clamp before ^(1/alpha) (alpha > 1)
                                                   # (mono N = 1024 with alpha != 0.6 and drown = 1 sits at the edge of the
                                                   # absolute criterion for every fp32 kernel here: thr is ~20 x rms there)
                                                   (44100, 256, 64, 0.8, 2, 37, 2), (44100, 256, 64, 1.2, 2, 21, 2)])
def test_against_oracle_random_spectra(sr, n, nb, alpha, b, m, c):
  rng = np.random.default_rng(n + nb)
  # spectra with a large dynamic range, including exact zeros
  y = (rng.standard_normal((b, m, n, c)) * 10.0 ** rng.uniform(-6, 0, (b, m, n, c))).astype(np.float32)
  y[0, 0] = 0.0
  y[-1, -1, ::3] = 0.0
  ref = oracle.PsychoacousticModel(sr, n, nb, alpha, compute_dtype=np.float64)
  pa = audiocodec_b200.PsychoacousticModel(sr, n, nb, alpha)
  ton = pa.tonality(cuda(y))
  ton_ref = ref.tonality(y.astype(np.float64))
  assert np.max(np.abs(ton.cpu().numpy() - ton_ref)) < 2e-5
  for drown in (0.0, 1.0):
    thr = pa.global_masking_threshold(cuda(y), ton, drown=drown).cpu().numpy()
    thr_ref = ref.global_masking_threshold(y.astype(np.float64), ton_ref, drown=drown)
    np.testing.assert_allclose(thr, thr_ref, rtol=3e-4)
    assert np.max(np.abs(thr - thr_ref)) <= 1e-5 * max(rms(y), 1e-3)


# ---- quantiser ----------------------------------------------------------------------------------------------
def test_quantizer_bit_exact_given_reference_threshold(golden):
  for name, sr, n, nb, alpha in PA_CASES:
    y = golden[f"pa_{name}_f32_y"]
    thr = golden[f"pa_{name}_f32_thr"]
    pa = audiocodec_b200.PsychoacousticModel(sr, n, nb, alpha)
    q = pa.quantize(cuda(y), cuda(thr))
    q_ref = oracle.quantize(y, thr)
    assert q.dtype == torch.int32
    assert np.array_equal(q.cpu().numpy(), q_ref)
    deq = pa.dequantize(q, cuda(thr)).cpu().numpy()
    assert np.array_equal(deq, oracle.dequantize(q_ref, thr))


def test_quantizer_half_way_cases():
  a = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, 0.26, -3.49, 7.5], device="cuda")
  pa = audiocodec_b200.PsychoacousticModel(44100, 8)
  q = pa.quantize(a.reshape(1, 1, 8, 1), torch.ones(1, 1, 8, 1, device="cuda"))
  assert q.flatten().tolist() == [0, 2, 2, 0, -2, 0, -3, 8]


@pytest.mark.parametrize("sr,n,c", [(44100, 256, 2), (48000, 1024, 2), (44100, 256, 1)])
def test_end_to_end_quantised_integers(sr, n, c):
  """x -> q on the GPU vs the fp32-faithful oracle: >= 99.99 % identical, the rest +-1."""
  b, blocks = 4, 40
  x = oracle.synthetic_audio(b, blocks * n, c, sr)
  codec = audiocodec_b200.AudioCodec(sr, filters_n=n)
  q, step = codec.encode(cuda(x))
  mdct = oracle.MDCTransformer(n, compute_dtype=np.float32)
  pa = oracle.PsychoacousticModel(sr, n, compute_dtype=np.float32)
  y_ref = mdct.transform(x)
  thr_ref = pa.global_masking_threshold(y_ref, pa.tonality(y_ref))
  q_ref = oracle.quantize(y_ref, thr_ref)
  diff = np.abs(q.cpu().numpy().astype(np.int64) - q_ref)
  assert diff.max() <= 1
  assert np.mean(diff == 0) >= 0.9999, np.mean(diff == 0)
  assert np.max(np.abs(step.cpu().numpy() - thr_ref)) <= 1e-5 * rms(x)
  # and the decoder: same reconstruction error as the reference chain, inside the north-star tolerance
  xhat = codec.decode(q, step).cpu().numpy()[:, n:-n]
  xhat_ref = mdct.inverse_transform(oracle.dequantize(q_ref, thr_ref))[:, n:-n]
  err, err_ref = rms(xhat - x), rms(xhat_ref - x)
  assert abs(err - err_ref) <= 1e-3 * err_ref


@pytest.mark.parametrize("channels", [1, 2, 4, 3])
def test_encode_matches_unfused_chain(channels):
  x = cuda(oracle.synthetic_audio(3, 256 * 30, channels, 44100))
  codec = audiocodec_b200.AudioCodec(44100, filters_n=256)
  y = codec.mdct.transform(x)
  pa = codec.psychoacoustic
  thr = pa.global_masking_threshold(y, pa.tonality(y))
  q, step = pa.encode(y)
  np.testing.assert_allclose(step.cpu().numpy(), thr.cpu().numpy(), rtol=1e-5)
  assert torch.equal(q, pa.quantize(y, step))
  q2, step2 = pa.encode(y, thr_scale=2.0)
  np.testing.assert_allclose(step2.cpu().numpy(), 2 * step.cpu().numpy(), rtol=1e-6)
  assert torch.equal(pa.encode(y, return_threshold=False), q)


def test_full_size_threshold_properties_cfg4_slice():
  """Size-independent properties on a large batch: thr >= 1e-7, scale covariance, drown monotonicity."""
  sr, n = 44100, 256
  x = torch.rand(64, n * 431, 1, device="cuda") - 0.5
  y = audiocodec_b200.MDCTransformer(n).transform(x)
  pa = audiocodec_b200.PsychoacousticModel(sr, n)
  thr = pa.global_masking_threshold(y, None)
  assert thr.min().item() >= 1e-7 * (1 - 1e-6) and torch.isfinite(thr).all()
  louder = pa.global_masking_threshold(2 * y, None)
  assert (louder >= thr * (1 - 1e-5)).all()           # more signal never lowers the threshold
  drowned = pa.global_masking_threshold(y, None, drown=1.0)
  assert (drowned >= thr * (1 - 1e-5)).all()          # drown=1 -> offset 0 -> maximal masking (:185)
  ton = pa.tonality(y)
  assert ton.max().item() <= 1.0


def test_add_noise_statistics():
  pa = audiocodec_b200.PsychoacousticModel(44100, 256)
  a = torch.zeros(8, 64, 256, 2, device="cuda")
  thr = torch.full_like(a, 0.6)
  noisy = pa.add_noise(a, thr, seed=42)
  assert abs(noisy.mean().item()) < 1e-3
  assert abs(noisy.std().item() - 0.1) < 1e-3          # sigma = thr / 6  (psychoacoustic.py:154-156)
  assert torch.equal(noisy, pa.add_noise(a, thr, seed=42))
  assert not torch.equal(noisy, pa.add_noise(a, thr, seed=43))
  z = (noisy / 0.1).flatten()
  assert abs((z ** 4).mean().item() - 3.0) < 0.05      # gaussian kurtosis


def test_dlpack_entry_points(golden):
  y = cuda(golden["pa_n256_f32_y"])
  pa = audiocodec_b200.PsychoacousticModel(44100, 256)
  ton_ptr = pa.tonality(y)
  thr_ptr = pa.global_masking_threshold(y, ton_ptr, drown=0.1)
  ton, thr = torch.empty_like(ton_ptr), torch.empty_like(thr_ptr)
  cy, ct, ch = y.__dlpack__(), ton.__dlpack__(), thr.__dlpack__()
  stream = torch.cuda.current_stream().cuda_stream
  plan = pa._plan(y.device)
  _capi.check(_capi.lib().ac_pa_tonality_dl(plan, _capi.dl_pointer(cy), _capi.dl_pointer(ct), stream))
  _capi.check(_capi.lib().ac_pa_threshold_dl(plan, _capi.dl_pointer(cy), _capi.dl_pointer(ct), 0.1, _capi.dl_pointer(ch), stream))
  assert torch.equal(ton, ton_ptr) and torch.equal(thr, thr_ptr)
  _capi.check(_capi.lib().ac_pa_threshold_dl(plan, _capi.dl_pointer(cy), None, 0.1, _capi.dl_pointer(ch), stream))
  np.testing.assert_allclose(thr.cpu().numpy(), thr_ptr.cpu().numpy(), rtol=1e-5)


def test_host_streaming_round_trip_matches_device_path():
  """AudioCodec.roundtrip_host (C pipeline: H2D, kernels, D2H over three streams) == encode + decode on the device."""
  sr, n, c, b = 44100, 256, 2, 11
  x = oracle.synthetic_audio(b, 256 * 50, c, sr)
  codec = audiocodec_b200.AudioCodec(sr, filters_n=n)
  xh = torch.from_numpy(x).pin_memory()
  q, step = codec.encode(torch.from_numpy(x).cuda())
  ref = codec.decode(q, step).cpu()
  for chunk in (1, 3, 4, 16):
    out, stats = codec.roundtrip_host(xh, chunk_clips=chunk, return_stats=True)
    assert torch.equal(out, ref)
    assert stats[0].item() == q.numel()
    assert stats[1].item() == (q != 0).sum().item()
    bits = torch.log2(2.0 * q.abs().double() + 1.0).sum().item()
    assert abs(stats[2].item() - bits) <= 1e-4 * bits
  out2 = codec.roundtrip_host(torch.from_numpy(x))             # pageable host memory also works
  assert torch.equal(out2, ref)
  with pytest.raises(TypeError):
    codec.roundtrip_host(torch.from_numpy(x).cuda())
  with pytest.raises(ValueError):
    codec.roundtrip_host(torch.zeros(1, 100, 2))
  assert tuple(codec.roundtrip_host(torch.zeros(0, 512, 2)).shape) == (0, 1024, 2)


def test_quantizer_odd_sizes_take_the_scalar_kernel():
  rng = np.random.default_rng(3)
  pa = audiocodec_b200.PsychoacousticModel(44100, 8)
  for shape in ((1, 1, 7, 1), (2, 3, 5, 3), (1, 2, 8, 2)):
    y = rng.standard_normal(shape).astype(np.float32)
    thr = (rng.random(shape).astype(np.float32) * 0.1 + 1e-3)
    q = pa.quantize(cuda(y), cuda(thr))
    assert np.array_equal(q.cpu().numpy(), oracle.quantize(y, thr))
    assert np.array_equal(pa.dequantize(q, cuda(thr)).cpu().numpy(), oracle.dequantize(q.cpu().numpy(), thr))


# ---- tensor-core tile kernel against the first-generation (fp32 FMA) tile kernel and the generic kernel ------
@pytest.mark.parametrize("sr,n,c,b,m", [(44100, 256, 2, 3, 37), (44100, 256, 1, 2, 70), (44100, 256, 4, 2, 11),
                                        (48000, 1024, 2, 2, 19), (44100, 64, 2, 2, 33), (44100, 128, 1, 1, 40),
                                        (44100, 512, 2, 1, 18), (44100, 320, 2, 2, 9), (48000, 2048, 1, 1, 5)])
def test_mma_tile_kernel_against_other_kernels(sr, n, c, b, m, monkeypatch):
  rng = np.random.default_rng(7 * n + c)
  y = cuda((rng.standard_normal((b, m, n, c)) * 10.0 ** rng.uniform(-5, 0, (b, m, n, c))).astype(np.float32))
  pa = audiocodec_b200.PsychoacousticModel(sr, n)
  out = {}
  for kernel in ("mma", "fma", "generic"):
    monkeypatch.setenv("AC_PA_KERNEL", kernel)
    q, step = pa.encode(y, thr_scale=1.5, drown=0.25)
    thr = pa.global_masking_threshold(y, pa.tonality(y), drown=0.25)
    torch.cuda.synchronize()
    out[kernel] = (q.cpu().numpy(), step.cpu().numpy(), thr.cpu().numpy())
    assert torch.equal(q, pa.quantize(y, step))             # the fused division is the IEEE one
  monkeypatch.delenv("AC_PA_KERNEL")
  # the spreading product on tcgen05 / tensor memory (the default where its operand tiles fit) against mma.sync
  monkeypatch.setenv("AC_PA_MMA", "sync")
  q, step = pa.encode(y, thr_scale=1.5, drown=0.25)
  thr = pa.global_masking_threshold(y, pa.tonality(y), drown=0.25)
  out["sync"] = (q.cpu().numpy(), step.cpu().numpy(), thr.cpu().numpy())
  monkeypatch.delenv("AC_PA_MMA")
  for other in ("fma", "generic", "sync"):
    np.testing.assert_allclose(out["mma"][1], out[other][1], rtol=5e-6)
    np.testing.assert_allclose(out["mma"][2], out[other][2], rtol=5e-6)
    diff = np.abs(out["mma"][0].astype(np.int64) - out[other][0])
    assert diff.max() <= 1 and np.mean(diff == 0) >= 0.999
  np.testing.assert_allclose(out["mma"][1], 1.5 * out["mma"][2], rtol=2e-6)


def test_full_size_encode_cfg2_kernels_agree(monkeypatch):
  """The bench workload itself (cfg2: 64 stereo clips x 10 s, 220 544 items, every CTA of the persistent grid walks
  several tiles from the end of the tensor, last tile ragged): the tensor-core kernel, the first-generation tile kernel
  and the stand-alone quantiser agree on the whole tensor; compared on the device."""
  sr, n = 44100, 256
  gen = torch.Generator(device="cuda").manual_seed(2)
  x = (torch.rand(64, 440832, 2, device="cuda", generator=gen) - 0.5) * \
      torch.logspace(-4, 0, 64, device="cuda").view(64, 1, 1)
  y = audiocodec_b200.MDCTransformer(n).transform(x)
  del x
  pa = audiocodec_b200.PsychoacousticModel(sr, n)
  q, step = pa.encode(y)
  assert torch.isfinite(step).all() and step.min().item() >= 1e-7 * (1 - 1e-6)
  assert torch.equal(q, pa.quantize(y, step))               # fused division = IEEE division, every coefficient
  q_again, step_again = pa.encode(y)
  assert torch.equal(q, q_again) and torch.equal(step, step_again)     # no run-to-run variation (no atomics)
  monkeypatch.setenv("AC_PA_KERNEL", "fma")
  q_fma, step_fma = pa.encode(y)
  monkeypatch.delenv("AC_PA_KERNEL")
  rel = ((step - step_fma).abs() / step_fma).max().item()
  assert rel <= 5e-6
  dq = (q - q_fma).abs()
  assert dq.max().item() <= 1 and (dq != 0).float().mean().item() <= 1e-3
  del q_fma, step_fma, dq
  monkeypatch.setenv("AC_PA_MMA", "sync")                   # mma.sync product against the tcgen05 product (the default)
  q_sync, step_sync = pa.encode(y)
  monkeypatch.delenv("AC_PA_MMA")
  assert ((step - step_sync).abs() / step_sync).max().item() <= 5e-6
  dq = (q - q_sync).abs()
  assert dq.max().item() <= 1 and (dq != 0).float().mean().item() <= 1e-3
  del q_sync, step_sync, dq
  # decode: the reconstruction error of every coefficient is at most half a quantiser step (+ fp32 rounding of y / step
  # and q * step)
  err = (pa.dequantize(q, step) - y).abs()
  assert (err <= 0.5 * step + 2e-7 * y.abs()).all()


# ---- compact side information: q + 64 bark-domain thresholds per (frame, channel), SURVEY.md 8f row 2 ------
@pytest.mark.parametrize("sr,n,c,b,m", [(44100, 256, 2, 3, 37), (44100, 256, 1, 2, 70), (44100, 256, 4, 2, 11),
                                        (48000, 1024, 2, 2, 19), (44100, 512, 2, 1, 18), (44100, 320, 2, 2, 9),
                                        (48000, 2048, 1, 1, 5), (44100, 256, 2, 64, 431)])
def test_compact_side_information_rebuilds_the_step_bit_for_bit(sr, n, c, b, m):
  rng = np.random.default_rng(3 * n + c)
  y = cuda((rng.standard_normal((b, m, n, c)) * 10.0 ** rng.uniform(-5, 0, (b, m, 1, c))).astype(np.float32))
  pa = audiocodec_b200.PsychoacousticModel(sr, n)
  q, step = pa.encode(y, thr_scale=1.5, drown=0.25)
  qc, bark = pa.encode_compact(y, thr_scale=1.5, drown=0.25)
  assert bark.shape == (b, m, 64, c) and torch.isfinite(bark).all() and bark.min().item() > 0
  assert torch.equal(q, qc)
  rebuilt = pa.expand_threshold(bark, thr_scale=1.5)
  assert torch.equal(rebuilt, step)                         # decoder and encoder use the same step, bit for bit
  assert torch.equal(pa.dequantize(qc, rebuilt), pa.dequantize(q, step))
  if c <= 2 and n in (256, 512, 1024):                      # fused decoder: expansion inside the inverse MDCT
    mdct = audiocodec_b200.MDCTransformer(n)
    x_ref = mdct.inverse_transform_dequantized(q, step)
    assert torch.equal(mdct.inverse_transform_compact(qc, bark, pa, thr_scale=1.5), x_ref)
  elif c <= 2:
    with pytest.raises(NotImplementedError):
      audiocodec_b200.MDCTransformer(n).inverse_transform_compact(qc, bark, pa, thr_scale=1.5)


def test_compact_side_information_unsupported_configurations():
  pa = audiocodec_b200.PsychoacousticModel(44100, 256, bark_bands_n=48)
  y = torch.zeros(1, 4, 256, 2, device="cuda")
  with pytest.raises(NotImplementedError):
    pa.encode_compact(y)
  pa3 = audiocodec_b200.PsychoacousticModel(44100, 256)
  with pytest.raises(NotImplementedError):
    pa3.encode_compact(torch.zeros(1, 4, 256, 3, device="cuda"))    # three channels: no tile kernel
  with pytest.raises(NotImplementedError):                          # N = 64: filters wider than three bark bands
    audiocodec_b200.PsychoacousticModel(44100, 64).encode_compact(torch.zeros(1, 4, 64, 2, device="cuda"))
  q, bark = pa3.encode_compact(torch.zeros(0, 4, 256, 2, device="cuda"))
  assert q.shape == (0, 4, 256, 2) and bark.shape == (0, 4, 64, 2)


def test_unaligned_views_take_the_generic_kernel():
  """A 4-byte-aligned view (odd float offset into a larger buffer) must not reach the vectorised tile kernels."""
  n, c = 256, 2
  pa = audiocodec_b200.PsychoacousticModel(44100, n)
  rng = np.random.default_rng(5)
  y = cuda(rng.standard_normal((2, 9, n, c)).astype(np.float32))
  buf = torch.empty(y.numel() + 1, device="cuda")
  view = buf[1:].view(y.shape)
  view.copy_(y)
  q_ref, step_ref = pa.encode(y)
  q, step = pa.encode(view)
  np.testing.assert_allclose(step.cpu().numpy(), step_ref.cpu().numpy(), rtol=5e-6)
  assert (q - q_ref).abs().max().item() <= 1


def test_non_finite_amplitudes_stay_local():
  """inf / NaN coefficients poison their own (frame, channel) item only and never fault (exponent-table lookups stay
  inside the shared-memory allocation whatever the bit pattern)."""
  n, c = 256, 2
  pa = audiocodec_b200.PsychoacousticModel(44100, n)
  rng = np.random.default_rng(11)
  y = rng.standard_normal((2, 70, n, c)).astype(np.float32)
  bad = y.copy()
  bad[0, 3, 17, 0] = np.inf
  bad[1, 40, 200, 1] = np.nan
  bad[1, 41, 5, 0] = -np.inf
  q_ref, step_ref = pa.encode(cuda(y))
  q, step = pa.encode(cuda(bad))
  torch.cuda.synchronize()
  keep = np.ones((2, 70, c), dtype=bool)
  keep[0, 3, 0] = keep[1, 40, 1] = keep[1, 41, 0] = False
  step, step_ref = step.cpu().numpy(), step_ref.cpu().numpy()
  q, q_ref = q.cpu().numpy(), q_ref.cpu().numpy()
  for b in range(2):
    for f in range(70):
      for ch in range(c):
        if keep[b, f, ch]:
          assert np.array_equal(step[b, f, :, ch], step_ref[b, f, :, ch])
          assert np.array_equal(q[b, f, :, ch], q_ref[b, f, :, ch])


def test_bit_estimate_and_rate_loop():
  """Device-side bitstream statistics equal the definition; the rate loop meets its target from below."""
  sr, n, c = 44100, 256, 2
  x = cuda(oracle.synthetic_audio(4, 40 * n, c, sr))
  codec = audiocodec_b200.AudioCodec(sr, filters_n=n)
  pa = codec.psychoacoustic
  y = codec.mdct.transform(x)
  q, _ = pa.encode(y)
  est = pa.bit_estimate(q)
  qa = q.abs().double()
  assert est["coefficients"] == q.numel() and est["nonzero"] == int((qa > 0).sum().item())
  assert abs(est["bits"] - torch.log2(2 * qa + 1).sum().item()) <= 1e-4 * est["bits"]
  last = None
  for bpc in (2.0, 1.0, 0.5):
    qr, step, scale = pa.encode_at_bitrate(y, bpc)
    bits = pa.bit_estimate(qr)["bits"]
    assert bits <= bpc * q.numel() and bits >= 0.98 * bpc * q.numel()
    assert torch.equal(qr, pa.quantize(y, step))
    assert last is None or scale > last          # fewer bits need a coarser step
    last = scale


def test_host_streaming_ring_is_bounded_and_reused():
  """The pipeline stages x / x_hat in a ring of four chunk buffers: many more chunks than slots give the same bits as
  the device path, and the device footprint does not grow with the host batch (capacity for config 5)."""
  sr, n, c = 44100, 256, 2
  codec = audiocodec_b200.AudioCodec(sr, filters_n=n)
  x = oracle.synthetic_audio(37, 256 * 20, c, sr)
  q, step = codec.encode(torch.from_numpy(x).cuda())
  ref = codec.decode(q, step).cpu()
  xh = torch.from_numpy(x).pin_memory()
  for chunk in (1, 2):                                  # 37 / 22 chunks through 4 slots
    assert torch.equal(codec.roundtrip_host(xh, chunk_clips=chunk), ref)
    assert torch.equal(codec.roundtrip_host(xh, chunk_clips=chunk), ref)      # the ring's events are re-armed
  del q, step, ref
  # 192 clips x 30 s (2 GB each way on the host) through 2-clip chunks: the library's device memory stays at a few chunks
  s = (sr * 30 // n) * n
  big = torch.empty(192, s, c, dtype=torch.float32).uniform_(-0.5, 0.5)
  torch.cuda.synchronize()
  torch.cuda.empty_cache()
  free0, _ = torch.cuda.mem_get_info()
  out = codec.roundtrip_host(big, chunk_clips=2)
  free1, _ = torch.cuda.mem_get_info()
  clip_bytes = 4 * s * c
  assert free0 - free1 < 4 * 2 * clip_bytes * 2 + 3 * 2 * clip_bytes + (256 << 20)    # ring (x, x_hat) + Y, q, step + slack
  assert free0 - free1 < big.numel() * 4 // 4                                         # far below the batch itself
  probe = [0, 95, 191]
  dev = torch.stack([big[i] for i in probe]).cuda()
  qd, sd = codec.encode(dev)
  assert torch.equal(out[probe], codec.decode(qd, sd).cpu())


def test_host_streaming_long_schedule():
  """More than 128 chunks: the ramp of the chunk schedule (1, 1, 2, 4, ... clips) must saturate at chunk_clips - it once
  kept doubling into an int64 overflow, which produced empty chunks and an invalid launch configuration."""
  sr, n, c = 44100, 256, 2
  codec = audiocodec_b200.AudioCodec(sr, filters_n=n)
  x = torch.from_numpy(oracle.synthetic_audio(150, 256 * 8, c, sr))
  q, step = codec.encode(x.cuda())
  ref = codec.decode(q, step).cpu()
  for chunk in (1, 2):                                  # 150 / 76 chunks
    assert torch.equal(codec.roundtrip_host(x.pin_memory(), chunk_clips=chunk), ref)
  out, stats = codec.roundtrip_host(x.pin_memory(), chunk_clips=1, return_stats=True)
  assert torch.equal(out, ref) and int(stats[0]) == q.numel() and int(stats[1]) == int((q != 0).sum())
