"""Development aid: per-chunk timeline of AudioCodec.roundtrip_host on cfg2 (AC_PIPE_TRACE=1: the C pipeline prints,
for every chunk, when its H2D copy, its kernels and its D2H copy started and ended)."""
import os, sys, time
os.environ["AC_PIPE_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audiocodec_b200

b, s, c, sr, n = 64, (441000 // 256) * 256, 2, 44100, 256
codec = audiocodec_b200.AudioCodec(sr, filters_n=n)
x = (torch.rand(b, s, c) - 0.5).pin_memory()
out = torch.empty(b, s + 2 * n, c).pin_memory()
cc = int(sys.argv[1]) if len(sys.argv) > 1 else None
for _ in range(3):
  t0 = time.perf_counter()
  codec.roundtrip_host(x, out, chunk_clips=cc)
  print("call: %.2f ms" % (1e3 * (time.perf_counter() - t0)), file=sys.stderr)
