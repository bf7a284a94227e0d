"""Development aid: phase timeline of the masking kernel (which CTAs are in which phase when).

Builds a copy of the library with -DAC_PA_TRACE (global-timer stamps at the phase boundaries of every tile), runs one
launch on cfg2 and stores the stamps in gpurun_out/k3_trace.npy: [cta][16][4] uint64 (slot 0 = SM id; slots 1.. = tiles:
start of the chunk loop, of the MMA phase, of phase D, end of phase D).
"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audiocodec_b200 import build as _build

out = os.path.join(ROOT, "tools", "_build", "libaudiocodec_b200_trace.so")
if "--build" in sys.argv or not os.path.exists(out):
  os.makedirs(os.path.dirname(out), exist_ok=True)
  cmd = [_build._nvcc()] + _build.NVCC_FLAGS + ["-DAC_PA_TRACE", "-I", os.path.join(ROOT, "include"), "-o", out] + \
        [os.path.join(_build.CSRC, s) for s in _build.SOURCES]
  subprocess.run(cmd, check=True)
  if "--build" in sys.argv:
    sys.exit(0)
os.environ["AUDIOCODEC_B200_LIB"] = out
import numpy as np
import torch
import audiocodec_b200
from audiocodec_b200 import _capi
import bench

b, c, sr, s, n = bench.workload_shape("cfg2")
dev = torch.device("cuda")
x = bench.device_synthetic_audio(torch, b, s, c, sr, 0, dev)
codec = audiocodec_b200.AudioCodec(sr, filters_n=n)
y = codec.mdct.transform(x)
pa = codec.psychoacoustic
for _ in range(3):
  pa.encode(y)
trace = torch.zeros(148 * 4 * 16 * 4, dtype=torch.int64, device=dev)
lib = _capi.lib()
lib.ac_debug_set_pa_trace.argtypes = [ctypes.c_void_p]
assert lib.ac_debug_set_pa_trace(trace.data_ptr()) == 0
torch.cuda.synchronize()
if os.environ.get("TRACE_FUSED"):
  codec.encode(x)          # the single-pass encoder: stamp 0 of a tile follows its bulk-copy wait and forward MDCT
else:
  pa.encode(y)
torch.cuda.synchronize()
np.save(os.path.join(ROOT, "gpurun_out", "k3_trace" + os.environ.get("TRACE_TAG", "") + ".npy"),
        trace.cpu().numpy().reshape(-1, 16, 4))
print("trace written")
