"""pytest configuration: markers, paths, shared fixtures."""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
  """A bare `pytest` on a box without a CUDA device skips the gpu tests instead of failing in CUDA initialisation."""
  try:
    import torch
    have_gpu = torch.cuda.is_available()
  except ImportError:
    have_gpu = False
  if have_gpu:
    return
  skip = pytest.mark.skip(reason="needs a CUDA device (audiocodec_b200 has no CPU path)")
  for item in items:
    if "gpu" in item.keywords:
      item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
  """Outputs of the unmodified reference (run under oracle/tf_shim by tests/golden/make_golden.py)."""
  with np.load(GOLDEN) as z:
    return {k: z[k] for k in z.files}


def rms(a):
  a = np.asarray(a, dtype=np.float64)
  if a.size == 0:
    return 0.0
  return float(np.sqrt(np.mean(a * a)))
