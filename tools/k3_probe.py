"""Development aid: where the fixed part of the masking kernel's time goes (cfg2, CUDA events).

Times ac_pa_encode_f32 (a) back to back between two events and (b) with an event pair around every launch (what
bench.py's per-kernel numbers see), for a list of AC_PA_ABLATE / AC_PA_CTAS settings, next to a trivial kernel.
"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audiocodec_b200
from audiocodec_b200 import _capi
import bench

b, c, sr, s, n = bench.workload_shape(os.environ.get("PROBE_WORKLOAD", "cfg2"))
dev = torch.device("cuda")
x = bench.device_synthetic_audio(torch, b, s, c, sr, 0, dev)
codec = audiocodec_b200.AudioCodec(sr, filters_n=n)
mdct, pa = codec.mdct, codec.psychoacoustic
y = mdct.transform(x)
frames = y.shape[1]
q = torch.empty(y.shape, dtype=torch.int32, device=dev)
step = torch.empty_like(y)
xhat = torch.empty(b, (frames + 1) * n, c, device=dev)
lib = _capi.lib()
pplan, mplan = pa._plan(dev), mdct._plan(dev)
sp = torch.cuda.current_stream().cuda_stream
tiny = torch.zeros(32, device=dev)


def k3():
  _capi.check(lib.ac_pa_encode_f32(pplan, y.data_ptr(), 0.0, 1.0, step.data_ptr(), q.data_ptr(), b, frames, c, sp))


def k1():
  _capi.check(lib.ac_mdct_forward_f32(mplan, x.data_ptr(), y.data_ptr(), b, s, c, sp))


def k2():
  _capi.check(lib.ac_mdct_inverse_dequant_f32(mplan, q.data_ptr(), step.data_ptr(), xhat.data_ptr(), b, frames, c, sp))


def back_to_back(fn, reps=20):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / reps * 1e3


def per_launch(fn, reps=20):
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
  torch.cuda.synchronize()
  ev[0].record()
  for i in range(reps):
    fn()
    ev[i + 1].record()
  torch.cuda.synchronize()
  return statistics.fmean(ev[i].elapsed_time(ev[i + 1]) for i in range(reps)) * 1e3


def chain():
  k1(); k3(); k2()


print(f"trivial kernel: back to back {back_to_back(lambda: tiny.add_(1.0)):.2f} us, per launch {per_launch(lambda: tiny.add_(1.0)):.2f} us")
for label, env in [("full", {}), ("ablate all (63)", {"AC_PA_ABLATE": "63"}), ("ablate all, 1 CTA/SM", {"AC_PA_ABLATE": "63", "AC_PA_CTAS": "1"}),
                   ("skip A (3+32)", {"AC_PA_ABLATE": "35"}), ("skip B (12)", {"AC_PA_ABLATE": "12"}), ("skip D (16)", {"AC_PA_ABLATE": "16"}),
                   ("skip A,B (47)", {"AC_PA_ABLATE": "47"})] + [(k, dict(kv.split("=") for kv in v.split(","))) for k, v in
                                                                  (e.split(":") for e in os.environ.get("PROBE_EXTRA", "").split(";") if e)]:
  for k in ("AC_PA_ABLATE", "AC_PA_CTAS", "AC_PA_MMA"):
    os.environ.pop(k, None)
  os.environ.update(env)
  print(f"K3 {label:24s}: back to back {back_to_back(k3):7.2f} us, per launch {per_launch(k3):7.2f} us")
for k in ("AC_PA_ABLATE", "AC_PA_CTAS", "AC_PA_MMA"):
  os.environ.pop(k, None)
print(f"K1: back to back {back_to_back(k1):.2f} us, per launch {per_launch(k1):.2f} us")
print(f"K2: back to back {back_to_back(k2):.2f} us, per launch {per_launch(k2):.2f} us")
print(f"chain K1 K3 K2: back to back {back_to_back(chain):.2f} us, with an event after each step {per_launch(chain):.2f} us")
if c == 2 and n == 256:
  def kf():
    _capi.check(lib.ac_codec_encode_f32(mplan, pplan, x.data_ptr(), 0.0, 1.0, step.data_ptr(), None, q.data_ptr(), b, s, c, None, sp))
  def kfq():
    _capi.check(lib.ac_codec_encode_f32(mplan, pplan, x.data_ptr(), 0.0, 1.0, None, None, q.data_ptr(), b, s, c, None, sp))
  print(f"fused encoder (x -> q, step): back to back {back_to_back(kf):.2f} us, per launch {per_launch(kf):.2f} us")
  print(f"fused encoder (x -> q only): back to back {back_to_back(kfq):.2f} us")
  def chain2():
    kf(); k2()
  print(f"chain fused K2: back to back {back_to_back(chain2):.2f} us, with an event after each step {per_launch(chain2):.2f} us")
