"""world_size-2 gloo tests of the multi-GPU host logic (sharding by clip range, the one gather of statistics).

The data path has no collective, so what N > 1 adds is exactly this: which clips a rank owns, that the shards
tile the batch, and that the gathered statistics are the sum over shards of what one rank would have computed
alone.  The per-shard "work" here is the CPU oracle (test infrastructure) - the CUDA path is covered by -m gpu.
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audiocodec_b200 import sharding
from oracle import audiocodec_oracle as oracle


def test_shard_range_tiles_the_batch():
  for total in (0, 1, 7, 64, 8192, 8193):
    for world in (1, 2, 3, 8):
      ranges = [sharding.shard_range(total, world, r) for r in range(world)]
      assert ranges[0][0] == 0 and ranges[-1][1] == total
      assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
      sizes = [b - a for a, b in ranges]
      assert max(sizes) - min(sizes) <= 1
  with pytest.raises(ValueError):
    sharding.shard_range(10, 2, 2)


def _shard_stats(first, last, n, sr, blocks):
  """[clips, coefficients, non-zero q, bit estimate] of clips [first, last) through the oracle chain."""
  if last <= first:
    return torch.zeros(4, dtype=torch.float64)
  x = oracle.synthetic_audio(last - first, blocks * n, 2, sr, first_clip=first)
  mdct, pa = oracle.MDCTransformer(n), oracle.PsychoacousticModel(sr, n)
  y = mdct.transform(x)
  q = oracle.quantize(y, pa.global_masking_threshold(y, pa.tonality(y)))
  qa = np.abs(q).astype(np.float64)
  return torch.tensor([last - first, q.size, float((qa > 0).sum()), float(np.log2(2 * qa + 1).sum())], dtype=torch.float64)


def _worker(rank, world, port, total, out):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    first, last = sharding.shard_range(total, world, rank)
    stats = _shard_stats(first, last, 64, 16000, 6)
    gathered = sharding.gather_stats(stats)
    times = sharding.max_over_ranks(torch.tensor([float(rank + 1), 5.0 - rank]))
    if rank == 0:
      out.put((gathered.numpy(), times.numpy()))
    dist.barrier()
  finally:
    dist.destroy_process_group()


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


@pytest.mark.timeout(300)
def test_two_rank_gather_matches_single_rank():
  total, world = 5, 2
  ctx = mp.get_context("spawn")
  out = ctx.Queue()
  procs = [ctx.Process(target=_worker, args=(r, world, (port := _free_port()) if r == 0 else port, total, out)) for r in range(world)]
  for p in procs:
    p.start()
  gathered, times = out.get(timeout=240)
  for p in procs:
    p.join(60)
    assert p.exitcode == 0
  assert gathered.shape == (2, 4)
  assert gathered[:, 0].tolist() == [3.0, 2.0]                     # clips per rank
  single = _shard_stats(0, total, 64, 16000, 6).numpy()
  np.testing.assert_allclose(gathered.sum(0), single, rtol=1e-12)  # shards are independent: sums match exactly
  assert times.tolist() == [2.0, 5.0]                              # element-wise max over ranks


def test_gather_without_process_group_is_identity():
  s = torch.tensor([1.0, 2.0, 3.0])
  assert torch.equal(sharding.gather_stats(s), s.unsqueeze(0))
  assert torch.equal(sharding.max_over_ranks(s.clone()), s)
