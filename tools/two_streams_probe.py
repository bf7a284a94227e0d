"""Development aid: two independent batches (two sets of buffers) of the cfg2 chain in flight on two streams against the
same 2 K steps on one stream: does kernel-level concurrency fill the tails of the persistent kernels?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda")
name = os.environ.get("PROBE_WORKLOAD", "cfg2")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
with torch.cuda.stream(s1):
  a = bench.Chain(torch, name, dev)
with torch.cuda.stream(s2):
  b = bench.Chain(torch, name, dev, first_clip=a.b)
torch.cuda.synchronize()
K = 40


def run(chains):
  for c in chains:
    c.step_once()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for i in range(K):
    chains[i % len(chains)].step_once()
  torch.cuda.synchronize()
  return (time.perf_counter() - t0) / K * 1e6


for rep in range(3):
  print(f"one stream: {run([a]):.2f} us per step;  two streams, alternating batches: {run([a, b]):.2f} us per step")
