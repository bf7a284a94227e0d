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
  # one nvcc per source, in parallel (the masking kernels alone take a minute), then one link step
  import concurrent.futures
  import tempfile
  compile_flags = [f for f in NVCC_FLAGS if f != "--shared"] + (["-Xptxas", "-v"] if verbose else [])
  log = []
  with tempfile.TemporaryDirectory(prefix="audiocodec_b200_build_") as tmp:
    def compile_one(src):
      obj = os.path.join(tmp, os.path.splitext(src)[0] + ".o")
      proc = subprocess.run([_nvcc()] + compile_flags + ["-I", os.path.join(PKG, "..", "include"), "-c", "-o", obj,
                             os.path.join(CSRC, src)], capture_output=True, text=True)
      return src, obj, proc
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
      results = list(pool.map(compile_one, SOURCES))
    for src, obj, proc in results:
      log.append(proc.stdout + proc.stderr)
      if proc.returncode != 0:
        sys.stderr.write("".join(log))
        raise RuntimeError("nvcc failed compiling " + src)
    link = subprocess.run([_nvcc()] + NVCC_FLAGS + ["-o", lib_path()] + [obj for _, obj, _ in results],
                          capture_output=True, text=True)
    log.append(link.stdout + link.stderr)
    if link.returncode != 0:
      sys.stderr.write("".join(log))
      raise RuntimeError("nvcc failed linking " + LIBNAME)
  if verbose:
    sys.stderr.write("".join(log))
  with open(stamp, "w") as f:
    f.write(fp)
  return lib_path()


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
