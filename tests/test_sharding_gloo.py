"""Multi-GPU host logic on CPU: world_size-2 gloo process group (127.0.0.1 rendezvous).

The data path has no collective (envs are sharded by index range); the only exchanges are the
all-reduce of the episode-statistics vector and the optional all-gather of per-env rewards."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qcd_b200.sharding import all_reduce_stats, gather_rewards, shard_range, summarize


def test_shard_range_partitions_exactly():
  for n, w in [(65536, 8), (10, 4), (7, 8), (1048576, 3)]:
    parts = [shard_range(n, r, w) for r in range(w)]
    assert parts[0][0] == 0 and parts[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    sizes = [hi - lo for lo, hi in parts]
    assert max(sizes) - min(sizes) <= 1
  with pytest.raises(ValueError): shard_range(8, 2, 2)


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


def _worker(rank, world, port, n_envs, q):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    lo, hi = shard_range(n_envs, rank, world)
    # per-rank statistics as the step kernel would accumulate them for its shard
    rng = np.random.default_rng(rank)
    local = np.zeros(12)
    local[0] = hi - lo                       # one finished episode per env
    local[1] = rng.random(hi - lo).sum()
    local[10] = 5 * (hi - lo)
    stats = torch.tensor(local)
    all_reduce_stats(stats)
    rewards = torch.arange(lo, hi, dtype=torch.float64) * 0.5
    full = gather_rewards(rewards, n_envs)
    q.put((rank, local.tolist(), stats.tolist(), full.tolist()))
  finally:
    dist.destroy_process_group()


def test_stats_allreduce_and_reward_gather_world2():
  world, n_envs = 2, 11
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, world, port, n_envs, q)) for r in range(world)]
  for p in procs: p.start()
  results = sorted(q.get(timeout=120) for _ in range(world))
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  want = np.sum([r[1] for r in results], axis=0)
  for _, _, reduced, full in results:
    np.testing.assert_allclose(reduced, want)
    np.testing.assert_allclose(full, np.arange(n_envs) * 0.5)
  s = summarize(torch.tensor(results[0][2]))
  assert s['episodes'] == n_envs and s['steps'] == 5 * n_envs
  np.testing.assert_allclose(s['return'], want[1] / n_envs)


def test_single_process_is_identity():
  stats = torch.arange(12, dtype=torch.float64)
  assert torch.equal(all_reduce_stats(stats.clone()), stats)
  r = torch.ones(5, dtype=torch.float64)
  assert gather_rewards(r, 5) is r
