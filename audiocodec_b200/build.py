"""Builds audiocodec_b200/lib/libaudiocodec_b200.so with nvcc for sm_100a (in-tree, so it ships with gpurun).

    python -m audiocodec_b200.build [--force] [--verbose]
"""

import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIBNAME = "libaudiocodec_b200.so"
SOURCES = ["capi.cu", "mdct_kernels.cu", "mdct_tile_kernels.cu", "psycho_kernels.cu", "psycho_mma_kernels.cu", "f64_kernels.cu", "elementwise_kernels.cu", "entropy_kernels.cu", "backward_kernels.cu", "tables.cpp"]
HEADERS = ["kernels.h", "tables.h", "fft_core.cuh", "async_copy.cuh", "mdct_tile_core.cuh", os.path.join("..", "..", "include", "audiocodec_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--shared"]


def lib_path():
  return os.path.join(LIBDIR, LIBNAME)


def _nvcc():
  return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _fingerprint():
  h = hashlib.sha256()
  for name in SOURCES + HEADERS:
    with open(os.path.join(CSRC, name), "rb") as f:
      h.update(name.encode())
      h.update(f.read())
  h.update(" ".join(NVCC_FLAGS).encode())
  return h.hexdigest()


def build(force=False, verbose=False):
  """Compiles the library if its sources changed; returns the path of the .so."""
  os.makedirs(LIBDIR, exist_ok=True)
  stamp = os.path.join(LIBDIR, LIBNAME + ".sha256")
  fp = _fingerprint()
  if not force and os.path.exists(lib_path()) and os.path.exists(stamp):
    with open(stamp) as f:
      if f.read().strip() == fp:
        return lib_path()
  cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        ["-I", os.path.join(PKG, "..", "include"), "-o", lib_path()] + [os.path.join(CSRC, s) for s in SOURCES]
  proc = subprocess.run(cmd, capture_output=True, text=True)
  if proc.returncode != 0:
    sys.stderr.write(proc.stdout + proc.stderr)
    raise RuntimeError("nvcc failed building " + LIBNAME)
  if verbose:
    sys.stderr.write(proc.stdout + proc.stderr)
  with open(stamp, "w") as f:
    f.write(fp)
  return lib_path()


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
