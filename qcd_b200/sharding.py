"""Multi-GPU layout: the env batch is sharded by contiguous index range, one process per GPU.

Envs never exchange data (each owns its state; the target is replicated), so the data path has NO
collective.  The only exchange is the episode-statistics vector (QCD_STATS_LEN float64 = 96 B,
Monitor fields of monitor.py:37-48 + termination histogram of algorithm.py:112-114) summed with one
all-reduce (NCCL over NVLink on GPUs; gloo in the CPU tests), and an optional all-gather of
per-env rewards for a single learner.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from ._lib import STAT_NAMES


def shard_range(num_envs: int, rank: int, world_size: int) -> tuple[int, int]:
  """[lo, hi) of the global env indices owned by `rank`; sizes differ by at most one."""
  if not 0 <= rank < world_size:
    raise ValueError(f"rank {rank} outside world of {world_size}")
  base, extra = divmod(num_envs, world_size)
  lo = rank * base + min(rank, extra)
  return lo, lo + base + (1 if rank < extra else 0)


def all_reduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
  """Sum the per-rank statistics vector in place over the process group; returns it."""
  if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
  return stats


def gather_rewards(local: torch.Tensor, num_envs: int, group=None) -> torch.Tensor:
  """All-gather the per-env reward shards into the global [num_envs] vector (index order)."""
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
    return local
  world, rank = dist.get_world_size(group), dist.get_rank(group)
  sizes = [shard_range(num_envs, r, world) for r in range(world)]
  assert local.numel() == sizes[rank][1] - sizes[rank][0]
  width = max(hi - lo for lo, hi in sizes)
  padded = torch.zeros(width, dtype=local.dtype, device=local.device)
  padded[:local.numel()] = local
  parts = [torch.empty_like(padded) for _ in range(world)]
  dist.all_gather(parts, padded, group=group)
  return torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, sizes)])


def summarize(stats: torch.Tensor) -> dict:
  """Means over finished episodes, the quantities the reference logs (algorithm.py:100-114)."""
  v = dict(zip(STAT_NAMES, stats.detach().cpu().tolist()))
  n = max(v['episodes'], 1.0)
  return {'episodes': v['episodes'], 'steps': v['steps'], 'bad_actions': v['bad_actions'],
          'return': v['sum_r'] / n, 'length': v['sum_l'] / n, 'depth': v['sum_d'] / n,
          'operations': v['sum_o'] / n, 'qubits': v['sum_q'] / n, 'metric': v['sum_m'] / n,
          'cost': v['sum_c'] / n, 'done': v['n_done'], 'depth_truncated': v['n_depth']}
