"""Development aid: AudioCodec.roundtrip_host on cfg2 (pinned buffers) for explicit chunk schedules (AC_PIPE_SCHEDULE)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audiocodec_b200

b, s, c, sr, n = 64, (441000 // 256) * 256, 2, 44100, 256
codec = audiocodec_b200.AudioCodec(sr, filters_n=n)
x = (torch.rand(b, s, c) - 0.5).pin_memory()
out = torch.empty(b, s + 2 * n, c).pin_memory()


def timeit(reps=8):
  codec.roundtrip_host(x, out, chunk_clips=16); torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(reps):
    codec.roundtrip_host(x, out, chunk_clips=16)
  torch.cuda.synchronize()
  return (time.perf_counter() - t0) / reps * 1e3


schedules = ["1,1,2,4,8,8,8,8,8,8,4,2,1,1", "2,2,4,4", "1,2,4,4", "1,2,3,4,5,6,7,8,8,8,6,4,2", "16", "6", "8", "1,1,2,4,8", "1,2,4,8,8,8,8,8,8,8,4",
             "1,2,4,6,6,6,6,6,6,6,6,6,2,1", "2,4,8,8,8,8,8,8,6,3,1", "1,1,2,4,8,8,8,8,8,8,5,2,1", "3,5,8,8,8,8,8,8,5,3", "1,1,2,4,8,16,16,8,4,2,1,1",
             "1,1,2,4,5,5,5,5,5,5,5,5,5,5,5,2"]
for rep in range(2):
  for sch in schedules:
    os.environ["AC_PIPE_SCHEDULE"] = sch
    print(f"{sch:45s} {timeit():.2f} ms")
