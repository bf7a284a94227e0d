"""CPU-only checks of the host side: the C library loads, exports the whole header, and its float64 table
builders agree with the oracle (and through it with the reference's golden fixtures).  No kernel runs here."""

import ctypes
import os
import re

import numpy as np
import pytest
import torch

import audiocodec_b200
from audiocodec_b200 import _capi
from oracle import audiocodec_oracle as oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
  header = open(os.path.join(ROOT, "include", "audiocodec_b200.h")).read()
  header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
  declared = set(re.findall(r"\b(ac_[a-z0-9_]+)\s*\(", header))
  assert len(declared) >= 20
  handle = ctypes.CDLL(_capi.library_path())
  for name in sorted(declared):
    assert hasattr(handle, name), f"{name} declared in the header but not exported"
  assert declared == set(_capi.SIGNATURES), "ctypes signature table out of sync with the header"
  assert _capi.lib().ac_abi_version() == 1


@pytest.mark.parametrize("n,window", [(8, 'vorbis'), (16, 'sine'), (12, 'ones'), (64, 'vorbis'), (256, 'vorbis'),
                                      (1024, 'sine')])
def test_mdct_tables_match_oracle(n, window):
  ours = audiocodec_b200.MDCTransformer(n, window_type=window)
  ref = oracle.MDCTransformer(n, window_type=window, compute_dtype=np.float32)
  assert ours.H.shape == (2, n, n) and ours.H.dtype == torch.float32
  assert np.max(np.abs(ours.H.numpy() - ref.H)) <= 1e-7
  assert np.max(np.abs(ours.H_inv.numpy() - ref.H_inv)) <= 2e-7
  assert np.count_nonzero(ours.H.numpy()) <= 2 * n


def test_mdct_tables_match_reference_fixture(golden):
  for n, window in [(8, 'vorbis'), (16, 'sine'), (12, 'ones'), (64, 'vorbis')]:
    ours = audiocodec_b200.MDCTransformer(n, window_type=window)
    assert np.max(np.abs(ours.H.numpy() - golden[f"H_{n}_{window}"])) <= 6e-8
    assert np.max(np.abs(ours.H_inv.numpy() - golden[f"Hinv_{n}_{window}"])) <= 6e-8


def test_mdct_float32_precompute_variant():
  a = audiocodec_b200.MDCTransformer(64, precompute_dtype='float32')
  b = oracle.MDCTransformer(64, precompute_dtype=np.float32)
  # float32 sin() implementations differ by an ulp and the consistency rule (mdctransformer.py:219-221)
  # amplifies that by 1 / w[0]; agreement is to a few float32 ulps, not bit-exact
  assert np.max(np.abs(a.H.numpy() - b.H)) <= 5e-6


@pytest.mark.parametrize("sr,n,nb,alpha", [(32768, 64, 64, 0.6), (44100, 256, 64, 0.6), (48000, 1024, 64, 0.6),
                                           (16000, 128, 24, 0.8)])
def test_pa_tables_match_oracle_and_fixture(golden, sr, n, nb, alpha):
  ours = audiocodec_b200.PsychoacousticModel(sr, n, nb, alpha)
  ref = oracle.PsychoacousticModel(sr, n, nb, alpha, compute_dtype=np.float32)
  assert np.array_equal(ours.W.numpy(), ref.W)
  assert np.array_equal(ours.W_inv.numpy(), ref.W_inv)
  np.testing.assert_allclose(ours.quiet_threshold_intensity.numpy(), ref.quiet_threshold_intensity, rtol=2e-7)
  np.testing.assert_allclose(ours.spreading_matrix.numpy(), ref.spreading_matrix, rtol=2e-7)
  key = f"pa_{sr}_{n}_{nb}"
  np.testing.assert_allclose(ours.spreading_matrix.numpy(), golden[f"{key}_S"], rtol=2e-7)
  np.testing.assert_allclose(ours.W.numpy(), golden[f"{key}_W"], atol=6e-8)
  assert abs(ours.max_bark - golden[f"{key}_scalars"][0]) < 1e-12
  assert abs(ours._dB_MIN - golden[f"{key}_scalars"][2]) < 1e-5
  # reference unit tests test_psychoacoustic.py:14-30
  assert float(torch.sum(torch.abs(ours.W.sum(dim=1) - 1.0))) < 1e-5
  assert float(torch.sum(torch.abs(ours.W_inv.sum(dim=1) - 1.0))) < 1e-5


def test_constructor_errors_match_reference():
  with pytest.raises(AssertionError):
    audiocodec_b200.MDCTransformer(7)                          # mdctransformer.py:26
  with pytest.raises(AttributeError):
    audiocodec_b200.MDCTransformer(8, window_type=None)        # mdctransformer.py:199 (.lower() on None)
  with pytest.raises(TypeError):
    audiocodec_b200.PsychoacousticModel(44100, compute_dtype='float16')   # psychoacoustic.py:42-43
  bf = audiocodec_b200.PsychoacousticModel(44100, 64, compute_dtype='bfloat16')      # accepted, as in the reference
  assert bf.compute_dtype == "bfloat16" and bf.W.dtype == torch.bfloat16 and bf.spreading_matrix.dtype == torch.bfloat16
  assert audiocodec_b200.MDCTransformer(8, compute_dtype=torch.bfloat16).H.dtype == torch.bfloat16
  with pytest.raises(TypeError):
    audiocodec_b200.MDCTransformer(8, compute_dtype='int32')
  assert audiocodec_b200.MDCTransformer(8, compute_dtype=torch.float32).compute_dtype == "float32"
  assert audiocodec_b200.MDCTransformer(8, compute_dtype=np.float32).compute_dtype == "float32"
  assert audiocodec_b200.MDCTransformer(8, compute_dtype=torch.float64).compute_dtype == "float64"
  assert audiocodec_b200.PsychoacousticModel(44100, 8, compute_dtype="float64").compute_dtype == "float64"


def test_no_cpu_path():
  """Host tensors are refused: the product has no CPU fallback."""
  m = audiocodec_b200.MDCTransformer(8)
  with pytest.raises(RuntimeError, match="no CPU path"):
    m.transform(torch.zeros(1, 16, 1))
  with pytest.raises(TypeError):
    m.transform(torch.zeros(1, 16, 1, dtype=torch.float64))
  pa = audiocodec_b200.PsychoacousticModel(32768, 8)
  with pytest.raises(RuntimeError, match="no CPU path"):
    pa.tonality(torch.zeros(1, 2, 8, 1))


def test_utilities():
  pa = audiocodec_b200.PsychoacousticModel(44100, 64)
  ref = oracle.PsychoacousticModel(44100, 64)
  a = np.asarray([0., 1e-9, 1e-3, 0.5, 1.0], np.float32)
  np.testing.assert_allclose(pa.amplitude_to_dB(torch.from_numpy(a)).numpy(), ref.amplitude_to_dB(a), rtol=1e-6)
  np.testing.assert_allclose(pa.amplitude_to_dB_norm(torch.from_numpy(a)).numpy(), ref.amplitude_to_dB_norm(a),
                             rtol=1e-5, atol=1e-6)
  assert abs(float(pa.bark2freq(pa.freq2bark(1234.5))) - 1234.5) < 1e-9


def test_product_does_not_import_oracle():
  """The oracle is test infrastructure; nothing under audiocodec_b200/ may reference it."""
  pkg = os.path.join(ROOT, "audiocodec_b200")
  for dirpath, _, files in os.walk(pkg):
    for name in files:
      if name.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
        text = open(os.path.join(dirpath, name)).read()
        assert "oracle" not in text.replace("the oracle", "").replace("The oracle", ""), name


def test_pa_mma_index_maps_emulation():
  """NumPy emulation of the tensor-core tile kernel's fragment addressing (tools/emulate_pa_mma.py): the swizzled P
  layout, the mma.m16n8k8 fragments and the rotating Toeplitz B fragments compute P^T S, bank-conflict free."""
  import importlib.util
  import pathlib
  path = pathlib.Path(__file__).resolve().parents[1] / "tools" / "emulate_pa_mma.py"
  spec = importlib.util.spec_from_file_location("emulate_pa_mma", path)
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  mod.main()


@pytest.mark.parametrize("sr,n", [(44100, 256), (48000, 1024), (44100, 320), (22050, 512), (44100, 64), (48000, 2048)])
def test_mma_job_list_reproduces_w(sr, n):
  """The job list of the tensor-core masking kernel (ac_pa_mma_jobs_host) is a lossless re-arrangement of W: every
  non-zero W[k][band] appears exactly once at the right chunk row, flags mark bands that cross chunks, the warps'
  job ranges and tonality ranges tile each chunk."""
  import ctypes
  from audiocodec_b200 import _capi
  nb = 64
  jobs = np.zeros(4 * 112, dtype=np.int32)
  start = np.zeros(9 * 32 + 1, dtype=np.int32)
  ton = np.zeros(9 * 32 + 2, dtype=np.int16)
  weights = np.zeros(8 * (n + 4 * nb), dtype=np.float32)
  counts = np.zeros(5, dtype=np.int32)
  _capi.check(_capi.lib().ac_pa_mma_jobs_host(float(sr), n, nb, 0.6, jobs.ctypes.data, start.ctypes.data, ton.ctypes.data,
                                              weights.ctypes.data, counts.ctypes.data))
  chunk, n_chunks, n_jobs, n_w, fits = (int(v) for v in counts)
  assert fits == 1 and chunk in (32, 64) and n_chunks == (n + chunk - 1) // chunk and n_w % 4 == 0
  w = np.empty((n, nb), dtype=np.float32)
  fp = ctypes.POINTER(ctypes.c_float)
  _capi.check(_capi.lib().ac_pa_tables_host(float(sr), n, nb, 0.6, w.ctypes.data_as(fp), None, None, None, None))
  rec = np.zeros_like(w)
  seen_band = set()
  for c in range(n_chunks):
    kc0, kcn = c * chunk, min(chunk, n - c * chunk)
    assert start[9 * c] <= start[9 * c + 8] and list(start[9 * c:9 * c + 9]) == sorted(start[9 * c:9 * c + 9])
    if c + 1 < n_chunks:
      assert start[9 * c + 8] == start[9 * (c + 1)]
    t = ton[9 * c:9 * c + 9]
    assert t[0] == 0 and t[8] == kcn and list(t) == sorted(t)
    for j in range(start[9 * c], start[9 * c + 8]):
      x, y, steps, wd = (int(v) for v in jobs[4 * j:4 * j + 4])
      so = (wd & 0xffffffff) >> 18                             # row of P in the layout of the mma.sync product
      band, swz = so // 256, (so % 256) // 4
      assert swz == (band & 3) << 3 and x % 264 == 0 and y % 32 == 0 and steps >= 1
      kp = band ^ 4                                            # the same row in the tcgen05 operand layout (low 16 bits)
      assert (wd & 0xffff) == (kp >> 2) * 512 + (kp & 3) * 128 + ((kp & 3) << 5)
      row0 = x // 264
      assert row0 + 4 * steps <= kcn + 3                       # padded steps stay inside the three zero rows
      assert bool(wd & 0x10000) == (band in seen_band)         # an earlier chunk holds the first part of the band
      for s in range(4 * steps):
        wt = weights[y // 8 + s]
        if wt != 0.0:
          assert row0 + s < kcn and rec[kc0 + row0 + s, band] == 0.0
          rec[kc0 + row0 + s, band] = wt
      nz = np.nonzero(w[:, band])[0]
      assert bool(wd & 0x20000) == (nz[-1] < kc0 + kcn)        # complete when its last filter is in this chunk
      seen_band.add(band)
  assert start[9 * n_chunks] == n_jobs and np.array_equal(rec, w)
