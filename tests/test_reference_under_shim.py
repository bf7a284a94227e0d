"""Runs the reference's OWN unit tests against its unmodified sources under oracle/tf_shim.

Build-container only: skipped wherever /root/reference is absent (e.g. the GPU box).  This is what pins
the NumPy TF shim (and through it tests/golden/reference_vectors.npz) to the reference's golden vector
(audiocodec/tests/test_mdctransformer.py:51-52) and its six other tests.
"""

import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("AUDIOCODEC_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "audiocodec")), reason="reference tree not present")
def test_reference_unit_tests_pass_under_shim():
  env = dict(os.environ)
  env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "oracle", "tf_shim"), REFERENCE])
  env["PYTHONDONTWRITEBYTECODE"] = "1"
  proc = subprocess.run(
    [sys.executable, "-W", "ignore", "-m", "unittest", "audiocodec.tests.test_mdctransformer",
     "audiocodec.tests.test_psychoacoustic"],
    cwd="/tmp", env=env, capture_output=True, text=True, timeout=300)
  assert proc.returncode == 0, proc.stderr[-2000:]
  assert "Ran 7 tests" in proc.stderr and "OK" in proc.stderr
