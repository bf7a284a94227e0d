"""Batch sharding across the GPUs of one node.

Every (clip, channel) sequence is independent (mdctransformer.py:295 folds channels into the batch; frames
only couple to their neighbours inside one sequence), so a batch shards over ranks by contiguous clip
ranges with NO collective on the data path.  The only exchange of a job is one all_gather of a small
per-rank statistics vector (coefficients, non-zero integers, bit estimate, ...) after the work is done:
NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""

import torch
import torch.distributed as dist


def shard_range(total_clips, world_size, rank):
  """Contiguous [first, last) clip range of `rank`; sizes differ by at most one clip, ranks may be empty."""
  if world_size < 1 or not (0 <= rank < world_size):
    raise ValueError(f"rank {rank} outside world of size {world_size}")
  base, extra = divmod(int(total_clips), world_size)
  first = rank * base + min(rank, extra)
  return first, first + base + (1 if rank < extra else 0)


def gather_stats(stats, group=None):
  """all_gather of a 1-D statistics tensor: returns [world_size, len(stats)] on every rank (rank order)."""
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
    return stats.unsqueeze(0).clone()
  out = [torch.empty_like(stats) for _ in range(dist.get_world_size(group))]
  dist.all_gather(out, stats.contiguous(), group=group)
  return torch.stack(out)


def max_over_ranks(values, group=None):
  """Element-wise maximum over ranks of a 1-D tensor of timings (the job takes as long as its slowest rank)."""
  if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
    dist.all_reduce(values, op=dist.ReduceOp.MAX, group=group)
  return values
