"""Development aid: what costs the copy engines time between two copy operations?  226 MB each way (cfg2), pinned host
memory, 28 MB pieces on one stream per direction, both directions at once:
  plain        copies only
  record       an event (timing disabled) recorded behind every copy
  record+wait  ... and a third stream that waits for every H2D event and runs a small kernel (the pipeline's shape)
  chained      ... and every D2H copy waits for that kernel's event (H2D -> kernel -> D2H, as roundtrip_host does)"""
import sys, os, time
import torch

nbytes = 64 * 440832 * 2 * 4
x = torch.empty(nbytes // 4).pin_memory(); out = torch.empty(nbytes // 4).pin_memory()
xd = torch.empty(nbytes // 4, device="cuda"); od = torch.empty(nbytes // 4, device="cuda")
h2d, d2h, run = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
small = torch.zeros(1 << 16, device="cuda")


def once(piece, mode):
  for i in range(0, x.numel(), piece):
    with torch.cuda.stream(h2d):
      xd[i:i + piece].copy_(x[i:i + piece], non_blocking=True)
      if mode != "plain":
        e = torch.cuda.Event(); e.record(h2d)
    if mode in ("record+wait", "chained"):
      with torch.cuda.stream(run):
        run.wait_event(e)
        small.add_(1.0)
        e2 = torch.cuda.Event(); e2.record(run)
    with torch.cuda.stream(d2h):
      if mode == "chained":
        d2h.wait_event(e2)
      out[i:i + piece].copy_(od[i:i + piece], non_blocking=True)
      if mode != "plain":
        e3 = torch.cuda.Event(); e3.record(d2h)


for piece_mb in (7, 28):
  piece = piece_mb * (1 << 20) // 4
  res = []
  for mode in ("plain", "record", "record+wait", "chained"):
    once(piece, mode); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
      once(piece, mode)
    torch.cuda.synchronize()
    res.append(f"{mode} {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms")
  print(f"pieces of {piece_mb} MB: " + ", ".join(res))
