"""Development aid: the cfg2 chain (forward MDCT -> masking + quantiser -> dequantising inverse MDCT) back to back
between two CUDA events, for A/B runs of environment switches that are read when the library loads (AC_PDL)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

chain = bench.Chain(torch, os.environ.get("PROBE_WORKLOAD", "cfg2"), torch.device("cuda"))
for _ in range(5):
  chain.step_once()
torch.cuda.synchronize()
best = []
for rep in range(3):
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(40):
    chain.step_once()
  e1.record()
  torch.cuda.synchronize()
  best.append(e0.elapsed_time(e1) / 40 * 1e3)
print(f"AC_PDL={os.environ.get('AC_PDL', 'default')} AC_PA_MMA={os.environ.get('AC_PA_MMA', 'default')}: chain " +
      " / ".join(f"{v:.2f}" for v in best) + " us per step")
