"""Development aid: floor of a chunked H2D -> (nothing) -> D2H pipeline on two streams (no kernels)."""
import time, torch
b, s, c = 64, (441000 // 256) * 256, 2
x = torch.empty(b, s, c).pin_memory(); out = torch.empty(b, s + 512, c).pin_memory()
xd = torch.empty(b, s, c, device="cuda"); od = torch.empty(b, s + 512, c, device="cuda")
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
def run(chunk, dep):
  evs = []
  for i in range(0, b, chunk):
    with torch.cuda.stream(sa):
      xd[i:i + chunk].copy_(x[i:i + chunk], non_blocking=True)
      e = torch.cuda.Event(); e.record(sa); evs.append(e)
  for k, i in enumerate(range(0, b, chunk)):
    with torch.cuda.stream(sb):
      if dep: sb.wait_event(evs[k])
      out[i:i + chunk].copy_(od[i:i + chunk], non_blocking=True)
  torch.cuda.synchronize()
for dep in (False, True):
  for chunk in (1, 2, 4, 8, 16, 64):
    run(chunk, dep); t0 = time.perf_counter()
    for _ in range(5): run(chunk, dep)
    print("dependent" if dep else "independent", "chunk", chunk, "%.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
