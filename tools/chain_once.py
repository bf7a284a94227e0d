"""Development aid: a short program for ncu captures - warm-up, then one launch each of the three kernels of the bench
chain and of the fused encoder on the bench workload (PROBE_WORKLOAD, default cfg2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

chain = bench.Chain(torch, os.environ.get("PROBE_WORKLOAD", "cfg2"), torch.device("cuda"))
for _ in range(3):
  chain.step_once()
  if chain.c == 2 and chain.n == 256:
    chain.codec.encode(chain.x)
torch.cuda.synchronize()
print("ok", float(chain.xhat.abs().mean()))
