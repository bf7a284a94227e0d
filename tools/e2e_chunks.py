"""Development aid: end-to-end (host buffers) time of AudioCodec.roundtrip_host for several chunk sizes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audiocodec_b200

b, s, c, sr, n = 64, (441000 // 256) * 256, 2, 44100, 256
codec = audiocodec_b200.AudioCodec(sr, filters_n=n)
x = (torch.rand(b, s, c) - 0.5).pin_memory()
out = torch.empty(b, s + 2 * n, c).pin_memory()
xd = torch.empty(b, s, c, device="cuda")
od = torch.empty(b, s + 2 * n, c, device="cuda")
def timeit(f, reps=5):
  f(); torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(reps): f()
  torch.cuda.synchronize()
  return (time.perf_counter() - t0) / reps * 1e3
print("h2d only  %.2f ms" % timeit(lambda: xd.copy_(x, non_blocking=True)))
print("d2h only  %.2f ms" % timeit(lambda: out.copy_(od, non_blocking=True)))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
  with torch.cuda.stream(s1): xd.copy_(x, non_blocking=True)
  with torch.cuda.stream(s2): out.copy_(od, non_blocking=True)
print("h2d + d2h concurrently  %.2f ms" % timeit(both))
for cc in (1, 2, 4, 8, 16, 32):
  print("chunk_clips %2d: %.2f ms" % (cc, timeit(lambda: codec.roundtrip_host(x, out, chunk_clips=cc))))
