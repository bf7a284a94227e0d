"""Development aid: AudioCodec.roundtrip_host on cfg2 (pinned buffers) for explicit chunk schedules (AC_PIPE_SCHEDULE)
and for the two ways of gating the re-use of a ring slot (AC_PIPE_HOST_GATE), interleaved in one process."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audiocodec_b200

b, s, c, sr, n = 64, (441000 // 256) * 256, 2, 44100, 256
codec = audiocodec_b200.AudioCodec(sr, filters_n=n)
x = (torch.rand(b, s, c) - 0.5).pin_memory()
out = torch.empty(b, s + 2 * n, c).pin_memory()


def timeit(reps=10):
  codec.roundtrip_host(x, out, chunk_clips=8); torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(reps):
    codec.roundtrip_host(x, out, chunk_clips=8)
  torch.cuda.synchronize()
  return (time.perf_counter() - t0) / reps * 1e3


schedules = {"doubling": "1,1,2,4,8,8,8,8,8,8,4,2,1,1", "linear": "1,2,3,4,5,6,7,8,8,8,6,4,2", "linear from 2": "2,3,4,5,6,7,8,8,8,7,4,2",
             "uniform 8": "8"}
res = {}
for rep in range(4):
  for name, sch in schedules.items():
    for gate in ("host", "stream"):
      os.environ["AC_PIPE_SCHEDULE"] = sch
      if gate == "host":
        os.environ["AC_PIPE_HOST_GATE"] = "1"
      else:
        os.environ.pop("AC_PIPE_HOST_GATE", None)
      res.setdefault((name, gate), []).append(timeit())
for k, v in res.items():
  print(f"{k[0]:14s} {k[1]:6s} gate: " + " ".join(f"{t:.2f}" for t in v) + f"   min {min(v):.2f} ms")
