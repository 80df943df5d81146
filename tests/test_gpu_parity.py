"""Parity of the CUDA path (through the C-ABI, via qcd_b200.batch.EnvBatch) with the oracle.

Bar (north_star): depth / operations / used_wires / terminated / truncated / status bit-exact;
state, metric, cost, reward within 1e-12 (complex128); float32 observations equal to float32(oracle)
up to the 1e-12 state difference (<= 1 ulp)."""
import zlib

import numpy as np
import pytest
import torch

from oracle.batch_oracle import OracleBatch
from oracle.qcd_oracle import make_target
from tests.helpers import CONFIGS, uniform_box_tape

pytestmark = pytest.mark.gpu
TOL = 1e-12
OBS_TOL = 1.2e-7      # one float32 ulp at |x| <= 1


def _cmp(gpu, ora, final_mask=None, what=""):
  g = lambda t: t.cpu().numpy()
  for name in ('depth', 'operations', 'used_wires', 'terminated', 'truncated', 'status', 'meta',
               'ep_length', 'episode_length'):
    a, b = getattr(gpu, name), getattr(ora, name)
    if a is None: continue
    if name == 'episode_length' and final_mask is not None:
      np.testing.assert_array_equal(g(a)[final_mask], b[final_mask], err_msg=f"{what}:{name}")
    else:
      np.testing.assert_array_equal(g(a), b, err_msg=f"{what}:{name}")
  for name in ('reward', 'metric', 'cost', 'last', 'ep_return'):
    a, b = getattr(gpu, name), getattr(ora, name)
    if a is None: continue
    np.testing.assert_allclose(g(a), b, atol=TOL, rtol=0, err_msg=f"{what}:{name}")
  np.testing.assert_allclose(g(gpu.state), ora.state, atol=TOL, rtol=0, err_msg=f"{what}:state")
  if gpu.obs is not None:
    np.testing.assert_allclose(g(gpu.obs), ora.obs, atol=OBS_TOL, rtol=0, err_msg=f"{what}:obs")
  if gpu.final_obs is not None and final_mask is not None and final_mask.any():
    np.testing.assert_allclose(g(gpu.final_obs)[final_mask], ora.final_obs[final_mask], atol=OBS_TOL, rtol=0)
    np.testing.assert_allclose(g(gpu.episode_return)[final_mask], ora.episode_return[final_mask], atol=TOL, rtol=0)


def _pair(N, eta, delta, objective, target, **kw):
  from qcd_b200.batch import EnvBatch
  return EnvBatch(N, eta, delta, objective, target, **kw), OracleBatch(N, eta, delta, objective, target, **kw)


@pytest.mark.parametrize("objective,eta,delta", CONFIGS + [('SP-random', 9, 24), ('SP-random', 12, 36), ('UC-random', 6, 24), ('UC-random', 5, 20)])
@pytest.mark.parametrize("sparse,punish,act_dtype", [(True, True, np.float32), (False, True, np.float64), (False, False, np.float32)])
def test_step_matches_oracle(cuda_device, objective, eta, delta, sparse, punish, act_dtype):
  rng = np.random.default_rng(zlib.crc32(repr((objective, eta, sparse, punish)).encode()))
  big = (4 ** eta if 'UC' in objective else 2 ** eta) >= 512
  N, T = (37, 24) if big else (301, 60)
  target = make_target(objective, eta, seed=5)
  gpu, ora = _pair(N, eta, delta, objective, target, punish=punish, sparse=sparse, autoreset=True, final_obs=True)
  _cmp(gpu, ora, what="reset")
  tape = uniform_box_tape(rng, eta, T, N, dtype=act_dtype, no_terminate=bool(rng.random() < 0.5) and not big)
  dev_tape = torch.as_tensor(tape, device=cuda_device)
  dones = 0
  for t in range(T):
    gpu.step(dev_tape[t]); ora.step(tape[t])
    fin = (ora.terminated | ora.truncated).astype(bool)
    dones += int(fin.sum())
    _cmp(gpu, ora, final_mask=fin, what=f"t={t}")
  assert dones > 0
  assert gpu.stats[0].item() == dones == ora.stats[0]
  np.testing.assert_allclose(gpu.stats.cpu().numpy(), ora.stats, rtol=1e-11, atol=1e-9)   # fused K3


@pytest.mark.parametrize("objective,eta,delta", [('SP-ghz5', 5, 20), ('UC-toffoli', 3, 30), ('SP-random', 8, 20)])
def test_no_autoreset_no_obs_and_per_env_targets(cuda_device, objective, eta, delta):
  rng = np.random.default_rng(9)
  N, T = 64, 70
  S = 4 ** eta if 'UC' in objective else 2 ** eta
  targets = rng.standard_normal((N, S)) + 1j * rng.standard_normal((N, S))
  targets /= np.linalg.norm(targets, axis=1, keepdims=True)
  gpu, ora = _pair(N, eta, delta, objective, targets, autoreset=False, obs_mode='state')
  assert gpu.per_env_target
  tape = uniform_box_tape(rng, eta, T, N, no_terminate=True)
  for t in range(T):
    gpu.step(torch.as_tensor(tape[t], device=cuda_device)); ora.step(tape[t])
    _cmp(gpu, ora, what=f"t={t}")
  assert ora.truncated.any()          # stepped past the depth limit without a reset: still identical
  gpu2, ora2 = _pair(N, eta, delta, objective, targets[0], obs_mode='none')
  for t in range(5):
    gpu2.step(torch.as_tensor(tape[t], device=cuda_device)); ora2.step(tape[t])
    _cmp(gpu2, ora2, what=f"noobs t={t}")


def test_invalid_actions_status_and_masked_reset(cuda_device):
  target = make_target('SP-ghz3', 3)
  gpu, ora = _pair(5, 3, 15, 'SP-ghz3', target)
  ok = np.array([[1, 0, 0, 1.0]] * 5)
  bad = np.array([[0, 3, 0, 0], [0, 0, -0.5, 0], [3, 0, 0, 0], [float('nan'), 0, 0, 0], [1, 1, 0, 0.3]])
  for a in (ok, bad, ok):
    gpu.step(torch.as_tensor(a, device=cuda_device)); ora.step(a)
    _cmp(gpu, ora)
  assert gpu.status.cpu().tolist() == [0, 0, 0, 0, 0]
  mask = np.array([1, 0, 1, 0, 0], np.uint8)
  gpu.reset(torch.as_tensor(mask)); ora.reset(mask)
  _cmp(gpu, ora, what="masked reset")


@pytest.mark.parametrize("objective,eta,delta", CONFIGS + [('SP-random', 12, 36), ('UC-random', 6, 24), ('SP-random', 11, 30),
                                                          ('SP-random', 10, 30), ('SP-random', 9, 27), ('UC-random', 5, 20)])
@pytest.mark.parametrize("sparse", [True, False])
def test_replay_equals_repeated_steps(cuda_device, objective, eta, delta, sparse):
  """qcd_replay(T) must leave every buffer exactly as T qcd_step launches do (same arithmetic,
  state kept in shared memory), and its per-step records must equal the per-step outputs."""
  from qcd_b200.batch import EnvBatch
  rng = np.random.default_rng(zlib.crc32(repr((objective, eta, sparse, 'replay')).encode()))
  big = (4 ** eta if 'UC' in objective else 2 ** eta) >= 512
  N, T = (19, 20) if big else (157, 48)
  target = make_target(objective, eta, seed=3)
  kw = dict(punish=True, sparse=sparse, autoreset=True)
  a = EnvBatch(N, eta, delta, objective, target, **kw)
  b = EnvBatch(N, eta, delta, objective, target, **kw)
  tape = torch.as_tensor(uniform_box_tape(rng, eta, T, N), device=cuda_device)
  rec = {k: torch.zeros((T, N), dtype=dt, device=cuda_device) for k, dt in
         [('reward', torch.float64), ('metric', torch.float64), ('cost', torch.float64), ('depth', torch.int32),
          ('terminated', torch.uint8), ('truncated', torch.uint8), ('status', torch.int8),
          ('operations', torch.int32), ('used_wires', torch.int32)]}
  b.replay(tape, rec)
  # amplitudes and integers are bit-identical (same explicitly rounded gate arithmetic in every
  # kernel); metric-derived floats may differ in the last bits (different reduction trees).
  exact = ('depth', 'terminated', 'truncated', 'status', 'operations', 'used_wires')
  def same(x, y, name):
    if name in exact or not x.dtype.is_floating_point:
      assert torch.equal(x, y), name
    else:
      assert torch.allclose(x, y, atol=1e-13, rtol=0, equal_nan=True), name
  for t in range(T):
    a.step(tape[t])
    for k in rec:
      same(rec[k][t], getattr(a, k), f"{k} t={t}")
  assert torch.equal(a.state, b.state) and torch.equal(a.obs, b.obs)
  for name in ('meta', 'last', 'ep_return', 'ep_length', 'reward', 'metric', 'cost', 'depth',
               'operations', 'used_wires', 'terminated', 'truncated', 'status', 'episode_return', 'episode_length'):
    same(getattr(a, name), getattr(b, name), name)
  # sparse replay without a metric record only evaluates the metric where it is needed
  if sparse:
    c = EnvBatch(N, eta, delta, objective, target, **kw)
    rec2 = {'reward': torch.zeros((T, N), dtype=torch.float64, device=cuda_device)}
    c.replay(tape, rec2)
    assert torch.allclose(rec2['reward'], rec['reward'], atol=1e-13, rtol=0) and torch.equal(c.state, a.state)


@pytest.mark.parametrize("fused", [True, False])
def test_stats_reduction(cuda_device, fused):
  from qcd_b200.batch import EnvBatch
  rng = np.random.default_rng(2)
  N, T, eta, delta = 1000, 40, 3, 15
  target = make_target('SP-ghz3', eta)
  gpu = EnvBatch(N, eta, delta, 'SP-ghz3', target, fused_stats=fused)
  ora = OracleBatch(N, eta, delta, 'SP-ghz3', target)
  tape = uniform_box_tape(rng, eta, T, N)
  want = np.zeros(12)
  for t in range(T):
    gpu.step(torch.as_tensor(tape[t], device=cuda_device)); gpu.reduce_stats(); ora.step(tape[t])
    fin = (ora.terminated | ora.truncated).astype(bool)
    want += [fin.sum(), ora.episode_return[fin].sum(), ora.episode_length[fin].sum(), ora.depth[fin].sum(),
             ora.operations[fin].sum(), ora.used_wires[fin].sum(), ora.metric[fin].sum(), ora.cost[fin].sum(),
             (fin & ~ora.truncated.astype(bool)).sum(), ora.truncated.sum(), N, 0]
  np.testing.assert_allclose(gpu.stats.cpu().numpy(), want, rtol=1e-12)
  np.testing.assert_allclose(ora.stats, want, rtol=1e-12)


def test_large_batch_indexing(cuda_device):
  """> 4 GiB of state: 2^30 amplitudes (262 144 envs x eta=12).  Env i replays the tape of env
  i mod 256, so every block of 256 envs must be bit-identical to the first one, which is checked
  against the oracle.  Guards the 64-bit index arithmetic of the step / reset kernels."""
  from qcd_b200.batch import EnvBatch
  free, _ = torch.cuda.mem_get_info(cuda_device)
  N, eta, delta, M, T = 262144, 12, 36, 256, 6
  if free < 45e9:
    pytest.skip("needs ~35 GB of device memory")
  rng = np.random.default_rng(11)
  target = make_target('SP-random', eta, seed=1)
  small = uniform_box_tape(rng, eta, T, M)
  gpu = EnvBatch(N, eta, delta, 'SP-random', target, obs_mode='state')
  ora = OracleBatch(M, eta, delta, 'SP-random', target, obs_mode='state')
  for t in range(T):
    a = torch.as_tensor(small[t], device=cuda_device).repeat(N // M, 1)
    gpu.step(a); ora.step(small[t])
  st = gpu.state.view(N // M, M, -1)
  assert torch.equal(st, st[:1].expand_as(st))
  ob = gpu.obs.view(N // M, M, -1)
  assert torch.equal(ob, ob[:1].expand_as(ob))
  for name in ('reward', 'depth', 'terminated', 'truncated'):
    v = getattr(gpu, name).view(N // M, M)
    assert torch.equal(v, v[:1].expand_as(v)), name
  np.testing.assert_allclose(gpu.state[-M:].cpu().numpy(), ora.state, atol=TOL, rtol=0)
  np.testing.assert_allclose(gpu.reward[-M:].cpu().numpy(), ora.reward, atol=TOL, rtol=0)
  np.testing.assert_array_equal(gpu.depth[-M:].cpu().numpy(), ora.depth)
  assert gpu.stats[10].item() == N * T


@pytest.mark.parametrize("objective,eta,delta,N", [('SP-random', 9, 27, 9), ('UC-random', 5, 20, 5), ('SP-ghz5', 5, 20, 70)])
def test_replay_long_tape_crosses_decode_chunks(cuda_device, objective, eta, delta, N):
  """T = 300 > 256: the CTA replay kernel decodes the tape in chunks of 256 steps (fusion never spans a
  chunk boundary); the warp replay kernel prefetches the next action every step."""
  from qcd_b200.batch import EnvBatch
  rng = np.random.default_rng(21)
  T = 300
  target = make_target(objective, eta, seed=9)
  a = EnvBatch(N, eta, delta, objective, target)
  b = EnvBatch(N, eta, delta, objective, target)
  tape = torch.as_tensor(uniform_box_tape(rng, eta, T, N), device=cuda_device)
  rec = {'reward': torch.zeros((T, N), dtype=torch.float64, device=cuda_device),
         'depth': torch.zeros((T, N), dtype=torch.int32, device=cuda_device)}
  b.replay(tape, rec)
  for t in range(T):
    a.step(tape[t])
    assert torch.equal(rec['depth'][t], a.depth), t
    assert torch.allclose(rec['reward'][t], a.reward, atol=1e-13, rtol=0), t
  assert torch.equal(a.state, b.state) and torch.equal(a.meta, b.meta) and torch.equal(a.obs, b.obs)
  assert torch.equal(a.ep_length, b.ep_length)
  assert torch.allclose(a.stats, b.stats, rtol=1e-12, atol=1e-9)


def test_empty_batch_is_a_noop(cuda_device):
  from qcd_b200.batch import EnvBatch
  b = EnvBatch(0, 3, 15, 'SP-ghz3', make_target('SP-ghz3', 3))
  b.step(torch.zeros((0, 4), dtype=torch.float32, device=cuda_device))
  b.replay(torch.zeros((5, 0, 4), dtype=torch.float32, device=cuda_device))
  b.reset(); b.reduce_stats()
  torch.cuda.synchronize()
  assert b.state.shape == (0, 8) and float(b.stats.sum()) == 0


@pytest.mark.parametrize("objective,eta", [('SP-random', 2), ('UC-random', 2)])
def test_maximum_depth_255_counter_boundary(cuda_device, objective, eta):
  """max_depth = 255 is the largest the uint8 depth stack supports: episodes must truncate exactly
  there (and reset), identically to the oracle, over several such episodes."""
  rng = np.random.default_rng(3)
  N, T, delta = 48, 900, 255
  target = make_target(objective, eta, seed=1)
  gpu, ora = _pair(N, eta, delta, objective, target, obs_mode='state')
  tape = uniform_box_tape(rng, eta, T, N, no_terminate=True)
  tape[..., 2] = tape[..., 1]                       # single-qubit gates only: depth grows by <= 1 per step
  dev_tape = torch.as_tensor(tape, device=cuda_device)
  seen = 0
  for t in range(T):
    gpu.step(dev_tape[t]); ora.step(tape[t])
    if ora.truncated.any():
      seen += int(ora.truncated.sum())
      assert (ora.depth[ora.truncated.astype(bool)] == 255).all()
      np.testing.assert_array_equal(gpu.truncated.cpu().numpy(), ora.truncated)
      np.testing.assert_array_equal(gpu.depth.cpu().numpy(), ora.depth)
      np.testing.assert_allclose(gpu.reward.cpu().numpy(), ora.reward, atol=TOL, rtol=0)
  assert seen >= N
  _cmp(gpu, ora, what="end")


@pytest.mark.parametrize("objective,eta,delta", [('SP-ghz5', 5, 20), ('UC-toffoli', 3, 30), ('SP-random', 10, 30)])
def test_replay_with_per_env_targets(cuda_device, objective, eta, delta):
  """Per-env targets take the general (first-generation) kernels in both step and replay mode."""
  from qcd_b200.batch import EnvBatch
  rng = np.random.default_rng(6)
  N, T = 23, 30
  S = 4 ** eta if 'UC' in objective else 2 ** eta
  targets = rng.standard_normal((N, S)) + 1j * rng.standard_normal((N, S))
  targets /= np.linalg.norm(targets, axis=1, keepdims=True)
  a = EnvBatch(N, eta, delta, objective, targets)
  b = EnvBatch(N, eta, delta, objective, targets)
  ora = OracleBatch(N, eta, delta, objective, targets)
  tape_np = uniform_box_tape(rng, eta, T, N)
  tape = torch.as_tensor(tape_np, device=cuda_device)
  b.replay(tape)
  for t in range(T):
    a.step(tape[t]); ora.step(tape_np[t])
  assert torch.equal(a.state, b.state) and torch.equal(a.meta, b.meta)
  assert torch.allclose(a.reward, b.reward, atol=1e-13, rtol=0)
  np.testing.assert_allclose(b.state.cpu().numpy(), ora.state, atol=TOL, rtol=0)
  np.testing.assert_allclose(b.reward.cpu().numpy(), ora.reward, atol=TOL, rtol=0)
  np.testing.assert_array_equal(b.depth.cpu().numpy(), ora.depth)
