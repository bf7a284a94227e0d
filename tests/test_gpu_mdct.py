"""GPU parity tests of the MDCT kernels (through the C ABI) against the oracle and the reference fixtures.

Tolerance (BASELINE.json north_star): MDCT / IMDCT coefficients within 1e-5 relative to signal RMS.
"""

import numpy as np
import pytest
import torch

import audiocodec_b200
from audiocodec_b200 import _capi
from oracle import audiocodec_oracle as oracle
from conftest import rms

pytestmark = pytest.mark.gpu
TOL = 1e-5   # x signal RMS


def cuda(a):
  return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def sine_wav(amplitude, frequency, sample_rate, duration_sec):
  t = np.arange(0, sample_rate * duration_sec, dtype=np.float32)
  return (amplitude * np.sin(2.0 * np.pi * frequency * t / sample_rate)).astype(np.float32).reshape(1, -1, 1)


# ---- the reference's own tests, on the CUDA path -----------------------------------------------------------
def test_inverse_identity():
  """audiocodec/tests/test_mdctransformer.py:19-37."""
  n = 256
  mdct = audiocodec_b200.MDCTransformer(n)
  x = sine_wav(0.8, 880, 16000, 1.)
  x = x[:, 0:n * (x.shape[1] // n)]
  back = mdct.inverse_transform(mdct.transform(cuda(x))).cpu().numpy()
  assert back.shape == (1, x.shape[1] + 2 * n, 1)
  assert np.max(np.abs(x - back[:, n:-n])) < 1e-5


def test_mdct_calculation():
  """audiocodec/tests/test_mdctransformer.py:39-54 (two-sided here)."""
  kat = np.asarray([-0.000412722176, 0.000430465181, 0.000789350364, -0.000867388735, -0.00275337417,
                    0.0132110268, 0.0193885863, 0.156005412, -0.233544752, -0.0129148215])
  y = audiocodec_b200.MDCTransformer(64).transform(cuda(sine_wav(0.8, 4, 64, 4.)[:, :256])).cpu().numpy()
  assert np.max(np.abs(y[0, 1, :10, 0] - kat)) < 1e-6
  y32 = audiocodec_b200.MDCTransformer(64, precompute_dtype='float32').transform(cuda(sine_wav(0.8, 4, 64, 4.)[:, :256]))
  assert np.max(np.abs(y32.cpu().numpy()[0, 1, :10, 0] - kat)) < 2e-7


def test_mdct_shape():
  """audiocodec/tests/test_mdctransformer.py:56-75."""
  y = audiocodec_b200.MDCTransformer(64).transform(torch.randn(128, 640, 2, device="cuda"))
  assert tuple(y.shape) == (128, 11, 64, 2) and y.dtype == torch.float32 and y.is_cuda


# ---- reference fixtures ------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,n,window", [("kat64", 64, 'vorbis'), ("sine256", 256, 'vorbis'),
                                           ("rand64_c2", 64, 'vorbis'), ("rand256_sine_c2", 256, 'sine'),
                                           ("rand1024_c1", 1024, 'vorbis'), ("rand12_ones_c3", 12, 'ones')])
def test_against_reference_fixture(golden, name, n, window):
  x = golden[f"mdct_{name}_x"]
  mdct = audiocodec_b200.MDCTransformer(n, window_type=window)
  y = mdct.transform(cuda(x)).cpu().numpy()
  y_ref = golden[f"mdct_{name}_f64_y"]
  assert y.shape == y_ref.shape
  assert np.max(np.abs(y - y_ref)) <= TOL * rms(x)
  back = mdct.inverse_transform(cuda(golden[f"mdct_{name}_f32_y"])).cpu().numpy()
  back_ref = golden[f"mdct_{name}_f64_xhat"]
  assert back.shape == back_ref.shape
  assert np.max(np.abs(back - back_ref)) <= TOL * rms(x)


# ---- oracle on seeded inputs: every fast-path size, generic sizes, ragged shapes ---------------------------
CASES = [  # n, window, B, blocks, C
  (16, 'vorbis', 3, 5, 1), (32, 'sine', 2, 7, 2), (64, 'vorbis', 5, 33, 2), (128, 'vorbis', 2, 19, 3),
  (256, 'vorbis', 3, 40, 2), (256, 'vorbis', 2, 127, 1), (256, 'sine', 1, 128, 2), (256, 'vorbis', 1, 129, 2),
  (256, 'vorbis', 2, 300, 1), (512, 'vorbis', 2, 21, 2), (1024, 'vorbis', 2, 37, 2), (1024, 'sine', 1, 9, 1),
  (2048, 'vorbis', 1, 11, 2), (4096, 'vorbis', 2, 5, 1), (4096, 'vorbis', 1, 3, 2),
  (12, 'ones', 2, 5, 3), (48, 'vorbis', 2, 9, 2), (100, 'sine', 1, 6, 1), (6, 'vorbis', 2, 4, 1),
  (256, 'vorbis', 2, 1, 2), (256, 'vorbis', 2, 0, 2), (64, 'vorbis', 1, 3, 5), (1024, 'vorbis', 1, 2, 12),
  # tile kernels: rectangular window, tile-edge block counts (16 frames per stereo tile, 32 per mono tile), N = 512 mono
  (256, 'ones', 2, 9, 2), (256, 'vorbis', 1, 15, 2), (256, 'vorbis', 1, 16, 2), (256, 'vorbis', 1, 17, 2),
  (256, 'vorbis', 1, 31, 1), (256, 'vorbis', 1, 32, 1), (512, 'sine', 2, 23, 1), (128, 'vorbis', 3, 50, 2),
  (1024, 'vorbis', 1, 16, 2), (1024, 'vorbis', 2, 7, 1),
]


@pytest.mark.parametrize("n,window,b,blocks,c", CASES)
def test_against_oracle(n, window, b, blocks, c):
  rng = np.random.default_rng(n * 1000 + blocks * 10 + c)
  x = rng.uniform(-1, 1, (b, blocks * n, c)).astype(np.float32)
  ref = oracle.MDCTransformer(n, window_type=window, compute_dtype=np.float64)
  mdct = audiocodec_b200.MDCTransformer(n, window_type=window)
  y = mdct.transform(cuda(x))
  assert tuple(y.shape) == (b, blocks + 1, n, c)
  y_ref = ref.transform(x.astype(np.float64))
  scale = max(rms(x), 1e-3)
  assert np.max(np.abs(y.cpu().numpy() - y_ref), initial=0.0) <= TOL * scale
  back = mdct.inverse_transform(y)
  assert tuple(back.shape) == (b, (blocks + 2) * n, c)
  back_ref = ref.inverse_transform(y_ref)
  assert np.max(np.abs(back.cpu().numpy() - back_ref), initial=0.0) <= TOL * scale
  # inverse on its own input (not a round trip): arbitrary coefficients
  coefs = rng.standard_normal((b, blocks + 3, n, c)).astype(np.float32)
  inv = mdct.inverse_transform(cuda(coefs)).cpu().numpy()
  inv_ref = ref.inverse_transform(coefs.astype(np.float64))
  assert np.max(np.abs(inv - inv_ref)) <= TOL * max(rms(inv_ref), 1e-3)


def test_empty_batch():
  mdct = audiocodec_b200.MDCTransformer(64)
  y = mdct.transform(torch.zeros(0, 128, 2, device="cuda"))
  assert tuple(y.shape) == (0, 3, 64, 2)
  assert tuple(mdct.inverse_transform(y).shape) == (0, 256, 2)


def test_errors_on_device():
  mdct = audiocodec_b200.MDCTransformer(64)
  with pytest.raises(ValueError):
    mdct.transform(torch.zeros(1, 100, 1, device="cuda"))            # samples_n % filters_n != 0 (:287)
  with pytest.raises(TypeError):
    mdct.transform(torch.zeros(1, 128, 1, device="cuda", dtype=torch.float64))   # no implicit cast (:22-23)
  with pytest.raises(ValueError):
    mdct.inverse_transform(torch.zeros(1, 3, 32, 1, device="cuda"))


def test_non_contiguous_and_misaligned_inputs():
  rng = np.random.default_rng(5)
  x = rng.uniform(-1, 1, (2, 64 * 6 + 1, 2)).astype(np.float32)
  mdct = audiocodec_b200.MDCTransformer(64)
  xc = cuda(x)
  y0 = mdct.transform(xc[:, 1:, :].contiguous())
  y1 = mdct.transform(xc[:, 1:, :])            # a strided view with a 8-byte offset
  assert torch.equal(y0, y1)


# ---- full-size properties (cfg2: 64 stereo clips x 10 s @ 44.1 kHz, N = 256) -------------------------------
def test_full_size_round_trip_cfg2():
  n, b, c = 256, 64, 2
  s = (441000 // n) * n
  g = torch.Generator(device="cuda").manual_seed(7)
  x = (torch.rand(b, s, c, device="cuda", generator=g) * 2 - 1)
  mdct = audiocodec_b200.MDCTransformer(n)
  y = mdct.transform(x)
  assert tuple(y.shape) == (b, s // n + 1, n, c)
  back = mdct.inverse_transform(y)
  err = (back[:, n:-n] - x).abs().max().item()
  assert err < 1e-5, err
  # the fold with a power-complementary window and the DCT-IV are orthogonal, the transform then scales by
  # 1 / sqrt(4N) (mdctransformer.py:125): energy is conserved up to the factor 4N
  e_x = x.double().pow(2).sum().item()
  e_y = y.double().pow(2).sum().item() * 4 * n
  assert abs(e_y / e_x - 1) < 1e-4
  # linearity
  y2 = mdct.transform(0.5 * x)
  assert (y2 - 0.5 * y).abs().max().item() < 1e-6


def test_long_window_round_trip_cfg3_slice():
  n, b, c = 1024, 8, 2
  s = (48000 * 30 // n) * n
  x = torch.rand(b, s, c, device="cuda") * 2 - 1
  mdct = audiocodec_b200.MDCTransformer(n)
  back = mdct.inverse_transform(mdct.transform(x))
  assert (back[:, n:-n] - x).abs().max().item() < 1e-5


# ---- decoder fusion and DLPack entry points -----------------------------------------------------------------
@pytest.mark.parametrize("n,c", [(256, 2), (1024, 1), (48, 2)])
def test_inverse_dequant_matches_unfused(n, c):
  g = torch.Generator(device="cuda").manual_seed(3)
  q = torch.randint(-7, 8, (3, 21, n, c), device="cuda", dtype=torch.int32, generator=g)
  thr = torch.rand(3, 21, n, c, device="cuda", generator=g) * 0.01 + 1e-4
  mdct = audiocodec_b200.MDCTransformer(n)
  fused = mdct.inverse_transform_dequantized(q, thr)
  plain = mdct.inverse_transform(q.float() * thr)
  assert torch.equal(fused, plain)


def test_dlpack_entry_points():
  rng = np.random.default_rng(11)
  x = cuda(rng.uniform(-1, 1, (2, 256 * 5, 2)).astype(np.float32))
  mdct = audiocodec_b200.MDCTransformer(256)
  y_ptr = mdct.transform(x)
  y = torch.empty_like(y_ptr)
  cx, cy = x.__dlpack__(), y.__dlpack__()
  stream = torch.cuda.current_stream().cuda_stream
  _capi.check(_capi.lib().ac_mdct_forward_dl(mdct._plan(x.device), _capi.dl_pointer(cx), _capi.dl_pointer(cy), stream))
  assert torch.equal(y, y_ptr)
  back = torch.empty(2, 256 * 7, 2, device="cuda")
  cb = back.__dlpack__()
  _capi.check(_capi.lib().ac_mdct_inverse_dl(mdct._plan(x.device), _capi.dl_pointer(cy), _capi.dl_pointer(cb), stream))
  assert torch.equal(back, mdct.inverse_transform(y_ptr))
  # validation: wrong dtype / shape are refused with the reference's error class
  bad = torch.zeros(2, 256 * 5, 2, device="cuda", dtype=torch.float64).__dlpack__()
  with pytest.raises(ValueError, match="float32"):
    _capi.check(_capi.lib().ac_mdct_forward_dl(mdct._plan(x.device), _capi.dl_pointer(bad), _capi.dl_pointer(cy), stream))
  short = torch.zeros(2, 5, 256, 2, device="cuda").__dlpack__()
  with pytest.raises(ValueError, match="shape"):
    _capi.check(_capi.lib().ac_mdct_forward_dl(mdct._plan(x.device), _capi.dl_pointer(cx), _capi.dl_pointer(short), stream))
  host = torch.zeros(2, 256 * 5, 2).__dlpack__()
  with pytest.raises(ValueError, match="not on a CUDA device"):
    _capi.check(_capi.lib().ac_mdct_forward_dl(mdct._plan(x.device), _capi.dl_pointer(host), _capi.dl_pointer(cy), stream))


def test_foreign_dlpack_tensor_is_adopted():
  """Any object with __dlpack__ (a TensorFlow eager tensor in production) is viewed zero-copy."""

  class Foreign:
    def __init__(self, t):
      self._t = t

    def __dlpack__(self, stream=None):
      return self._t.__dlpack__()

    def __dlpack_device__(self):
      return self._t.__dlpack_device__()

  x = torch.rand(1, 512, 1, device="cuda") - 0.5
  mdct = audiocodec_b200.MDCTransformer(256)
  assert torch.equal(mdct.transform(Foreign(x)), mdct.transform(x))


def test_float32_precompute_on_the_tile_path():
  rng = np.random.default_rng(17)
  x = rng.uniform(-1, 1, (2, 256 * 20, 2)).astype(np.float32)
  ref = oracle.MDCTransformer(256, compute_dtype=np.float64, precompute_dtype=np.float32)
  mdct = audiocodec_b200.MDCTransformer(256, precompute_dtype='float32')
  y = mdct.transform(cuda(x))
  y_ref = ref.transform(x.astype(np.float64))
  # float32 window chains differ by an ulp between sin() implementations and the consistency rule amplifies it
  # (see test_mdct_float32_precompute_variant): a looser, still tight, bound
  assert np.max(np.abs(y.cpu().numpy() - y_ref)) <= 5e-6
  back = mdct.inverse_transform(y).cpu().numpy()
  assert np.max(np.abs(back[:, 256:-256] - x)) < 2e-5


def test_legacy_path_matches_tile_path():
  """The any-channel-count kernels (AC_MDCT_LEGACY=1 in a fresh process) and the tile kernels agree to fp32 rounding."""
  import os, subprocess, sys, tempfile
  rng = np.random.default_rng(23)
  x = rng.uniform(-1, 1, (3, 256 * 37, 2)).astype(np.float32)
  mdct = audiocodec_b200.MDCTransformer(256)
  y = mdct.transform(cuda(x)).cpu().numpy()
  with tempfile.TemporaryDirectory() as tmp:
    np.save(os.path.join(tmp, "x.npy"), x)
    code = ("import numpy as np, torch, audiocodec_b200, sys\n"
            "x = torch.from_numpy(np.load(sys.argv[1] + '/x.npy')).cuda()\n"
            "m = audiocodec_b200.MDCTransformer(256)\n"
            "y = m.transform(x)\n"
            "np.save(sys.argv[1] + '/y.npy', y.cpu().numpy()); np.save(sys.argv[1] + '/b.npy', m.inverse_transform(y).cpu().numpy())\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, AC_MDCT_LEGACY="1", PYTHONPATH=root)
    proc = subprocess.run([sys.executable, "-c", code, tmp], env=env, capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stderr[-1500:]
    y_legacy, back_legacy = np.load(os.path.join(tmp, "y.npy")), np.load(os.path.join(tmp, "b.npy"))
  assert np.max(np.abs(y - y_legacy)) < 2e-7
  back = mdct.inverse_transform(cuda(y)).cpu().numpy()
  assert np.max(np.abs(back - back_legacy)) < 2e-6


def test_kernels_really_launched():
  before = _capi.lib().ac_kernel_launch_count()
  audiocodec_b200.MDCTransformer(256).transform(torch.zeros(1, 512, 1, device="cuda"))
  assert _capi.lib().ac_kernel_launch_count() == before + 1


# ---- repeated-run determinism on ragged tiles (stand-in for racecheck: compute-sanitizer is closed on the pool) -----
@pytest.mark.parametrize("n,b,blocks,c", [(256, 3, 37, 2), (256, 2, 129, 1), (1024, 2, 19, 2), (512, 2, 23, 1),
                                          (256, 5, 1, 2), (64, 3, 50, 2)])
def test_repeated_runs_are_bit_identical(n, b, blocks, c):
  """The in-place shared-memory rows of the tile kernels (input block -> FFT scratch -> output frame, overlap-add of
  adjacent frames) are reused without atomics: a missing barrier would show up as run-to-run differences.  Twenty runs
  of K1, the plain K2 and the dequantising K2 on ragged tile shapes, interleaved with a kernel that perturbs timing."""
  rng = np.random.default_rng(n + blocks)
  x = cuda(rng.uniform(-1, 1, (b, blocks * n, c)).astype(np.float32))
  mdct = audiocodec_b200.MDCTransformer(n)
  y0 = mdct.transform(x)
  back0 = mdct.inverse_transform(y0)
  q = torch.randint(-7, 8, tuple(y0.shape), device="cuda", dtype=torch.int32)
  step = torch.rand(tuple(y0.shape), device="cuda") + 0.01
  deq0 = mdct.inverse_transform_dequantized(q, step)
  junk = torch.empty(1 << 22, device="cuda")
  for i in range(20):
    if i % 3 == 0:
      junk.normal_()                                            # different co-residency / cache state per run
    assert torch.equal(mdct.transform(x), y0)
    assert torch.equal(mdct.inverse_transform(y0), back0)
    assert torch.equal(mdct.inverse_transform_dequantized(q, step), deq0)


# ---- sizes behind the tuned paths ------------------------------------------------------------------
@pytest.mark.parametrize("n,window,b,blocks,c", [(4100, 'sine', 1, 3, 1), (5000, 'vorbis', 1, 2, 2),
                                                 (4096, 'vorbis', 1, 2, 6), (2048, 'vorbis', 1, 3, 12)])
def test_generic_kernels_large_n_and_wide_tiles(n, window, b, blocks, c):
  """Generic O(N^2) kernels with more than 48 KB of dynamic shared memory (filters_n > 4096) and power-of-two filters_n
  whose any-channel tile does not fit in shared memory (falls through to the generic kernels).  The reference here is
  the library's float64 path (itself pinned to the oracle at 1e-13 in test_gpu_float64.py): the oracle's dense
  [2, N, N] tables take a minute to build at these sizes."""
  rng = np.random.default_rng(n + c)
  x = rng.uniform(-1, 1, (b, blocks * n, c)).astype(np.float32)
  ref = audiocodec_b200.MDCTransformer(n, window_type=window, compute_dtype='float64')
  mdct = audiocodec_b200.MDCTransformer(n, window_type=window)
  y = mdct.transform(cuda(x))
  y_ref = ref.transform(cuda(x.astype(np.float64)))
  assert (y.double() - y_ref).abs().max().item() <= 2 * TOL * rms(x)        # O(N^2) fp32 sums: twice the FFT tolerance
  back = mdct.inverse_transform(y)
  assert (back.double() - ref.inverse_transform(y_ref)).abs().max().item() <= 2 * TOL * rms(x)
  assert (back[:, n:-n] - cuda(x)).abs().max().item() <= 1e-4
  q = torch.randint(-3, 4, tuple(y.shape), device="cuda", dtype=torch.int32)
  step = torch.full(tuple(y.shape), 0.01, device="cuda")
  deq = mdct.inverse_transform_dequantized(q, step)
  deq_ref = ref.inverse_transform(q.double() * 0.01)
  assert (deq.double() - deq_ref).abs().max().item() <= 2 * TOL * max(rms(deq_ref.cpu().numpy()), 1e-3)
