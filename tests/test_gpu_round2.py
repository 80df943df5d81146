"""Round-2 parity tests: the holes VERDICT r01 named.

* actions produced by a CUDA kernel launched right before every `qcd_step` (programmatic dependent
  launch must not read them before `griddepcontrol.wait`), eager and inside a CUDA graph;
* BASELINE.json configs[3] and configs[4] at their FULL sizes against the oracle on a strided sample;
* the terminal observation of `qcd_replay`, per-env targets through `step_host`.
"""
import zlib

import numpy as np
import pytest
import torch

from oracle.batch_oracle import OracleBatch
from oracle.qcd_oracle import make_target
from tests.helpers import uniform_box_tape

pytestmark = pytest.mark.gpu
TOL = 1e-12
OBS_TOL = 1.2e-7


def _check(gpu, ora, sel=None, what="", skip=()):
  """ints bit-exact, floats <= 1e-12, float32 observations <= 1 ulp; `sel` = env indices of `gpu` that `ora` holds."""
  pick = (lambda t: t.cpu().numpy()) if sel is None else (lambda t: t[sel].cpu().numpy())
  for name in ('depth', 'operations', 'used_wires', 'terminated', 'truncated', 'status', 'meta', 'ep_length'):
    if name in skip: continue
    np.testing.assert_array_equal(pick(getattr(gpu, name)), getattr(ora, name), err_msg=f"{what}:{name}")
  for name in ('reward', 'metric', 'cost', 'ep_return', 'state'):
    if name in skip: continue
    np.testing.assert_allclose(pick(getattr(gpu, name)), getattr(ora, name), atol=TOL, rtol=0, err_msg=f"{what}:{name}")
  if gpu.obs is not None:
    np.testing.assert_allclose(pick(gpu.obs), ora.obs, atol=OBS_TOL, rtol=0, err_msg=f"{what}:obs")


@pytest.mark.parametrize("objective,eta,delta", [('SP-ghz5', 5, 20), ('UC-toffoli', 3, 30), ('SP-random', 9, 27)])
@pytest.mark.parametrize("graphed", [False, True])
def test_actions_written_by_a_kernel_right_before_every_step(cuda_device, objective, eta, delta, graphed):
  """A policy stand-in (an elementwise torch kernel) overwrites ONE action buffer immediately before every
  qcd_step on the same stream; with programmatic dependent launch on, a step that read its action before
  `griddepcontrol.wait` would see the previous step's values and diverge from the oracle."""
  from qcd_b200.batch import EnvBatch
  S = 4 ** eta if 'UC' in objective else 2 ** eta
  N, T = (65536, 200) if S <= 64 else (2048, 60)
  rng = np.random.default_rng(17)
  target = make_target(objective, eta, seed=2)
  tape = uniform_box_tape(rng, eta, T, N)
  tape[1::2, :, 0] = np.where(tape[1::2, :, 0] >= 2, 1.5, tape[1::2, :, 0])   # odd steps never Terminate: a stale
  dtape = torch.as_tensor(tape, device=cuda_device)                          # action changes depth / state
  gpu = EnvBatch(N, eta, delta, objective, target, device=cuda_device)
  ora = OracleBatch(N, eta, delta, objective, target)
  act = torch.zeros((N, 4), dtype=torch.float32, device=cuda_device)
  one = torch.ones((), dtype=torch.float32, device=cuda_device)
  checkpoints = sorted({0, 1, 2, T // 2, T - 1})
  stream = torch.cuda.Stream(cuda_device)

  def produce_and_step(t):
    torch.mul(dtape[t], one, out=act)        # a kernel writes the action buffer ...
    gpu.step(act)                            # ... and the env step follows on the same stream

  done = 0
  with torch.cuda.stream(stream):
    if not graphed:
      for t in range(T):
        produce_and_step(t)
        if t in checkpoints:
          stream.synchronize()
          while done <= t:
            ora.step(tape[done]); done += 1
          _check(gpu, ora, what=f"eager t={t}")
    else:
      produce_and_step(0)                    # lazy module loading outside of the capture
      stream.synchronize()
      gpu.reset(); gpu.zero_stats()
      g = torch.cuda.CUDAGraph()
      with torch.cuda.graph(g, stream=stream):
        for t in range(T):
          produce_and_step(t)
      gpu.reset(); gpu.zero_stats()
      g.replay()
      stream.synchronize()
      for t in range(T): ora.step(tape[t])
      _check(gpu, ora, what="graph")
  assert gpu.stats[10].item() == N * T


def test_config4_full_scale_sp_random12_one_million_envs(cuda_device):
  """BASELINE.json configs[3] at full size on ONE GPU: SP-random eta=12, 1 Mi envs (64 GiB of states + 32 GiB of
  state-half observations), punish=True; every 1024th env is checked against the oracle."""
  from qcd_b200.batch import EnvBatch
  free, _ = torch.cuda.mem_get_info(cuda_device)
  if free < 112e9:
    pytest.skip("needs ~100 GB of device memory")
  N, eta, delta, T, stride = 1 << 20, 12, 36, 5, 1024
  rng = np.random.default_rng(4)
  target = make_target('SP-random', eta, seed=42)
  gpu = EnvBatch(N, eta, delta, 'SP-random', target, obs_mode='state', punish=True, device=cuda_device)
  sel = torch.arange(0, N, stride, device=cuda_device)
  ora = OracleBatch(N // stride, eta, delta, 'SP-random', target, obs_mode='state', punish=True)
  for t in range(T):
    a = uniform_box_tape(rng, eta, 1, N)[0]
    if t == T - 2: a[:, 0] = 2.5                    # everybody terminates once: metric + reset at full size
    gpu.step(torch.as_tensor(a, device=cuda_device)); ora.step(a[::stride])
    _check(gpu, ora, sel=sel, what=f"C4 t={t}")
  st = gpu.stats.cpu().numpy()
  assert st[10] == N * T and st[0] >= N
  # size-independent property: every state is normalised (unitary evolution of e0)
  nrm = torch.linalg.vector_norm(gpu.state, dim=1)
  assert float((nrm - 1).abs().max()) < 1e-12
  del gpu
  torch.cuda.empty_cache()


def test_config5_full_scale_uc_random6_replay(cuda_device):
  """BASELINE.json configs[4] at full size: UC-random eta=6 (64x64 unitaries), 262 144 envs, one qcd_replay of
  T=64 steps; every 256th env against the oracle stepping the same tape."""
  from qcd_b200.batch import EnvBatch
  free, _ = torch.cuda.mem_get_info(cuda_device)
  if free < 40e9:
    pytest.skip("needs ~30 GB of device memory")
  N, eta, delta, T, stride = 262144, 6, 24, 64, 256
  rng = np.random.default_rng(8)
  target = make_target('UC-random', eta, seed=42)
  gpu = EnvBatch(N, eta, delta, 'UC-random', target, obs_mode='state', device=cuda_device)
  ora = OracleBatch(N // stride, eta, delta, 'UC-random', target, obs_mode='state')
  tape = uniform_box_tape(rng, eta, T, N)
  tape[:, 5 * stride] = uniform_box_tape(rng, eta, T, 1, no_terminate=True)[:, 0]   # one env runs into the depth limit
  rec = {'reward': torch.zeros((T, N), dtype=torch.float64, device=cuda_device),
         'depth': torch.zeros((T, N), dtype=torch.int32, device=cuda_device),
         'terminated': torch.zeros((T, N), dtype=torch.uint8, device=cuda_device),
         'truncated': torch.zeros((T, N), dtype=torch.uint8, device=cuda_device)}
  gpu.replay(torch.as_tensor(tape, device=cuda_device), rec)
  sel = torch.arange(0, N, stride, device=cuda_device)
  for t in range(T):
    ora.step(tape[t, ::stride])
    np.testing.assert_allclose(rec['reward'][t, sel].cpu().numpy(), ora.reward, atol=TOL, rtol=0, err_msg=f"t={t}")
    np.testing.assert_array_equal(rec['depth'][t, sel].cpu().numpy(), ora.depth, err_msg=f"t={t}")
    np.testing.assert_array_equal(rec['terminated'][t, sel].cpu().numpy(), ora.terminated)
    np.testing.assert_array_equal(rec['truncated'][t, sel].cpu().numpy(), ora.truncated)
  assert rec['truncated'][:, 5 * stride].any()
  _check(gpu, ora, sel=sel, what="C5 end")
  assert gpu.stats[10].item() == N * T
  # size-independent property: every resident V is unitary  (V^H V = I), checked on a random 4096-env subset
  idx = torch.as_tensor(rng.choice(N, 4096, replace=False), device=cuda_device)
  V = gpu.state_view()[idx]
  eye = torch.eye(64, dtype=torch.complex128, device=cuda_device)
  assert float((V.conj().transpose(1, 2) @ V - eye).abs().max()) < 1e-12
  del gpu
  torch.cuda.empty_cache()


@pytest.mark.parametrize("objective,eta,delta", [('SP-ghz5', 5, 20), ('UC-toffoli', 3, 30), ('SP-random', 10, 30), ('UC-random', 6, 24)])
def test_replay_leaves_terminal_observations_like_steps(cuda_device, objective, eta, delta):
  """final_obs after qcd_replay(T) == final_obs after T qcd_step calls (written at every reset of the tape)."""
  from qcd_b200.batch import EnvBatch
  rng = np.random.default_rng(31)
  S = 4 ** eta if 'UC' in objective else 2 ** eta
  N, T = (11, 24) if S >= 512 else (130, 40)
  target = make_target(objective, eta, seed=3)
  a = EnvBatch(N, eta, delta, objective, target, final_obs=True)
  b = EnvBatch(N, eta, delta, objective, target, final_obs=True)
  ora = OracleBatch(N, eta, delta, objective, target, final_obs=True)
  tape_np = uniform_box_tape(rng, eta, T, N)
  tape = torch.as_tensor(tape_np, device=cuda_device)
  b.replay(tape)
  seen = np.zeros(N, bool)
  for t in range(T):
    a.step(tape[t]); ora.step(tape_np[t])
    seen |= (ora.terminated | ora.truncated).astype(bool)
  assert seen.any()
  assert torch.equal(a.final_obs, b.final_obs) and torch.equal(a.obs, b.obs)
  np.testing.assert_allclose(b.final_obs.cpu().numpy()[seen], ora.final_obs[seen], atol=OBS_TOL, rtol=0)


def test_step_host_with_per_env_targets_returns_each_envs_target_half(cuda_device):
  from qcd_b200 import CircuitDesignerVectorEnv
  rng = np.random.default_rng(5)
  N, eta, delta = 64, 3, 9
  targets = rng.standard_normal((N, 8)) + 1j * rng.standard_normal((N, 8))
  targets /= np.linalg.norm(targets, axis=1, keepdims=True)
  venv = CircuitDesignerVectorEnv(N, eta, delta, 'SP-random', target=targets)
  ora = OracleBatch(N, eta, delta, 'SP-random', targets)
  venv.reset()
  for t in range(6):
    a = uniform_box_tape(rng, eta, 1, N)[0]
    obs, r, term, trunc = venv.step_host(a)
    ora.step(a)
    np.testing.assert_allclose(obs, ora.obs, atol=OBS_TOL, rtol=0)          # full rows: [state | own target]
    np.testing.assert_allclose(r, ora.reward, atol=TOL, rtol=0)


def test_fold_stats_sums_and_clears_the_replicas(cuda_device):
  from qcd_b200 import _lib
  from qcd_b200.batch import EnvBatch
  rng = np.random.default_rng(2)
  N, eta, delta = 4096, 3, 15
  gpu = EnvBatch(N, eta, delta, 'SP-ghz3', make_target('SP-ghz3', eta))
  tape = torch.as_tensor(uniform_box_tape(rng, eta, 12, N), device=cuda_device)
  for t in range(12): gpu.step(tape[t])
  want = gpu.stats.clone()
  out = gpu.fold_stats(zero=True)
  torch.cuda.synchronize()
  assert torch.allclose(out[:_lib.STATS_LEN], want, rtol=1e-13, atol=0) and float(out[_lib.STATS_LEN:].abs().sum()) == 0
  assert float(gpu._stats_raw.abs().sum()) == 0 and want[10].item() == 12 * N


@pytest.mark.parametrize("groups,zero_copy", [(2, True), (3, True), (2, False)])
def test_pipelined_host_groups_match_oracle(cuda_device, groups, zero_copy):
  """step_host_async / step_host_wait: env groups stepped on their own streams; every group's results are
  those of the oracle, whatever the interleaving, and the compact observation rows are the state halves.
  zero_copy: the kernel writes reward / terminated / truncated into the pinned host arrays itself (the device
  tensors of those three are then not the step's outputs); otherwise they are copied like the observations."""
  from qcd_b200 import CircuitDesignerVectorEnv
  rng = np.random.default_rng(12)
  N, eta, delta, T = 1000, 5, 20, 30
  venv = CircuitDesignerVectorEnv(N, eta, delta, 'SP-ghz5', obs_mode='state', zero_copy_scalars=zero_copy)
  ora = OracleBatch(N, eta, delta, 'SP-ghz5', venv.target, obs_mode='state')
  venv.reset()
  ranges = venv.host_groups(groups)
  assert ranges[0][0] == 0 and ranges[-1][1] == N
  tape = uniform_box_tape(rng, eta, T, N)
  got = []
  for t in range(T):
    for g, (lo, hi) in enumerate(ranges):
      if t > 0:
        got.append((t - 1, lo, hi, [np.array(x) for x in venv.step_host_wait(g)[:4]]))
      venv.step_host_async(g, tape[t, lo:hi])
  with pytest.raises(RuntimeError): venv.step_host_async(0, tape[0, :ranges[0][1]])     # previous step not collected
  for g, (lo, hi) in enumerate(ranges):
    got.append((T - 1, lo, hi, [np.array(x) for x in venv.step_host_wait(g)[:4]]))
  got.sort(key=lambda r: (r[0], r[1]))
  it = iter(got)
  for t in range(T):
    ora.step(tape[t])
    for _ in ranges:
      tt, lo, hi, (obs, r, term, trunc) = next(it)
      assert tt == t
      np.testing.assert_allclose(obs, ora.obs[lo:hi], atol=OBS_TOL, rtol=0)
      np.testing.assert_allclose(r, ora.reward[lo:hi], atol=TOL, rtol=0)
      np.testing.assert_array_equal(term, ora.terminated[lo:hi].astype(bool))
      np.testing.assert_array_equal(trunc, ora.truncated[lo:hi].astype(bool))
  _check(venv.batch, ora, what="groups end", skip=('reward', 'terminated', 'truncated') if zero_copy else ())
  tobs = venv.target_obs
  assert tobs.shape == (64,) and tobs.dtype == np.float32
  assert venv.episode_stats()['steps'] == N * T


def test_single_env_zero_copy_outputs_are_fresh_arrays(cuda_device):
  """The single-env path hands out copies (like the reference: fresh numpy arrays per step), not views of the
  mapped buffers that the next step overwrites."""
  import qcd_gym
  from qcd_b200 import gym_compat as gym
  qcd_gym.register_envs()
  env = gym.make("CircuitDesigner-v0", max_qubits=2, max_depth=12, objective='SP-bell')
  o0, _ = env.reset()
  o1 = env.step([1, 0, 0, np.pi / 2])[0]
  o2 = env.step([0, 0, 0, np.pi / 2])[0]
  assert o0[0] == 1 and not np.array_equal(o1, o2) and not np.array_equal(o0, o1)
  np.testing.assert_allclose(o1[:8], [np.sqrt(.5), 0, 0, 0, 0, -np.sqrt(.5), 0, 0], atol=1e-7)


def test_device_episode_window_matches_a_host_deque(cuda_device):
  """episode_window=100 on the device == collections.deque(maxlen=100) of Monitor records fed in env order
  (SB3 ep_info_buffer; algorithm.py:49,89-91,100)."""
  from collections import deque
  from qcd_b200 import CircuitDesignerVectorEnv
  rng = np.random.default_rng(3)
  N, eta, delta, T = 257, 3, 15, 40
  venv = CircuitDesignerVectorEnv(N, eta, delta, 'SP-ghz3', episode_window=100)
  ora = OracleBatch(N, eta, delta, 'SP-ghz3', venv.target)
  venv.reset()
  ref = deque(maxlen=100)
  assert np.isnan(venv.rolling_means()['r'])
  for t in range(T):
    a = uniform_box_tape(rng, eta, 1, N, no_terminate=(t % 3 != 0))[0]
    if t == 5: a[:, 0] = 2.5                       # 257 episodes finish in one step: only the last 100 may land
    venv.step(torch.as_tensor(a, device=cuda_device)); ora.step(a)
    for i in np.nonzero(ora.terminated | ora.truncated)[0]:
      ref.append({'r': ora.episode_return[i], 'l': int(ora.episode_length[i]), 'd': int(ora.depth[i]), 'o': int(ora.operations[i]),
                  'q': int(ora.used_wires[i]), 'm': ora.metric[i], 'c': ora.cost[i], 'env': int(i)})
    if t in (2, 5, 6, T - 1):
      got = venv.window.episodes()
      assert len(got) == len(ref)
      for g, r in zip(got, ref):
        assert (g['env'], g['l'], g['d'], g['o'], g['q']) == (r['env'], r['l'], r['d'], r['o'], r['q'])
        np.testing.assert_allclose([g['r'], g['m'], g['c']], [r['r'], r['m'], r['c']], atol=TOL, rtol=0)
      means = venv.rolling_means()
      for k in 'rldoqmc':
        np.testing.assert_allclose(means[k], np.mean([e[k] for e in ref]), atol=1e-11, rtol=0)
  assert int(venv.window.count.item()) == int(ora.stats[0])


def test_env_state_checkpoint_round_trip(cuda_device):
  """state_dict() / load_state_dict(): a restored batch continues bit-identically."""
  from qcd_b200.batch import EnvBatch
  rng = np.random.default_rng(14)
  N, eta, delta = 300, 5, 20
  target = make_target('SP-ghz5', eta)
  a = EnvBatch(N, eta, delta, 'SP-ghz5', target, sparse=False, final_obs=True)
  tape = torch.as_tensor(uniform_box_tape(rng, eta, 30, N), device=cuda_device)
  for t in range(12): a.step(tape[t])
  sd = a.state_dict()
  b = EnvBatch(N, eta, delta, 'SP-ghz5', target, sparse=False, final_obs=True)
  b.load_state_dict(sd)
  for t in range(12, 30):
    a.step(tape[t]); b.step(tape[t])
  for k in EnvBatch._STATE_KEYS:
    assert torch.equal(getattr(a, k), getattr(b, k)), k
  with pytest.raises(ValueError):
    EnvBatch(N, eta, delta + 1, 'SP-ghz5', target, sparse=False, final_obs=True).load_state_dict(sd)


def test_make_vec_builds_the_batched_env(cuda_device):
  import qcd_gym
  from qcd_b200 import gym_compat as gym
  from qcd_b200.vector_env import CircuitDesignerVectorEnv
  qcd_gym.register_envs()
  venv = gym.make_vec("CircuitDesigner-v0", num_envs=64, max_qubits=2, max_depth=12, objective='SP-bell')
  assert isinstance(venv, CircuitDesignerVectorEnv) and venv.num_envs == 64 and venv.spec.id == "CircuitDesigner-v0"
  obs, info = venv.reset()
  assert obs.shape == (64, 16) and venv.single_action_space.shape == (4,)
  a = torch.tensor([[1, 0, 0, np.pi / 2]] * 64, dtype=torch.float32, device=cuda_device)
  obs, r, term, trunc, info = venv.step(a)
  assert not bool(term.any()) and int(info['depth'].sum()) == 64


def test_sb3_vec_env_rolling_window_and_zero_copy(cuda_device):
  from qcd_b200.sb3_vec_env import CircuitDesignerSB3VecEnv
  rng = np.random.default_rng(2)
  n, eta, delta = 4, 2, 12
  vec = CircuitDesignerSB3VecEnv(n, eta, delta, 'SP-bell', seed=1, stats_window_size=10)
  vec.reset()
  eps = []
  for t in range(200):
    obs, rew, dones, infos = vec.step(uniform_box_tape(rng, eta, 1, n)[0])
    assert obs.shape == (n, 16) and rew.dtype == np.float32
    eps += [i['episode'] for i in infos if 'episode' in i]
  assert len(eps) > 20 and list(vec.ep_info_buffer) == eps[-10:]
  np.testing.assert_allclose(vec.rolling_means()['m'], np.mean([e['m'] for e in eps[-10:]]))
  assert sum(len(e.termination_reasons) for e in vec.envs) == len(eps)


@pytest.mark.parametrize("objective,eta,delta", [('SP-bell', 2, 12), ('SP-ghz5', 5, 20), ('UC-toffoli', 3, 30), ('SP-random', 7, 21), ('UC-hadamard', 1, 9)])
def test_step_outputs_do_not_depend_on_the_terminal_observation_buffer(cuda_device, objective, eta, delta):
  """A batch of full tiles (n_envs % 32 == 0) with and without a terminal-observation buffer: bit-identical states and
  outputs (the final_obs writes are a side channel), both equal to the oracle.  (A step-kernel instantiation compiled
  without the tail / null-pointer predicates for this shape was measured 2.5 % SLOWER on SP-ghz5 and dropped.)"""
  from qcd_b200.batch import EnvBatch
  rng = np.random.default_rng(77)
  N, T = 96, 50
  target = make_target(objective, eta, seed=6)
  fast = EnvBatch(N, eta, delta, objective, target, sparse=False)
  gen = EnvBatch(N, eta, delta, objective, target, sparse=False, final_obs=True)
  ora = OracleBatch(N, eta, delta, objective, target, sparse=False)
  tape = uniform_box_tape(rng, eta, T, N)
  dt = torch.as_tensor(tape, device=cuda_device)
  for t in range(T):
    fast.step(dt[t]); gen.step(dt[t]); ora.step(tape[t])
    for k in ('state', 'obs', 'meta', 'reward', 'metric', 'cost', 'depth', 'operations', 'used_wires', 'terminated',
              'truncated', 'status', 'ep_return', 'ep_length', 'last', 'episode_return'):
      assert torch.equal(getattr(fast, k), getattr(gen, k)), (t, k)
  _check(fast, ora, what="full tiles")
  assert torch.equal(fast.stats, gen.stats)


@pytest.mark.parametrize("objective,eta,delta", [('SP-random', 9, 27), ('UC-random', 5, 20), ('SP-ghz5', 5, 20)])
def test_replay_counts_invalid_actions_like_steps(cuda_device, objective, eta, delta):
  """Tapes with out-of-range wires / gates: status, untouched state and the steps / bad_actions statistics of a replay
  equal those of repeated steps and of the oracle."""
  from qcd_b200.batch import EnvBatch
  rng = np.random.default_rng(41)
  N, T = 9, 40
  target = make_target(objective, eta, seed=1)
  tape_np = uniform_box_tape(rng, eta, T, N)
  bad = rng.random((T, N)) < 0.3
  tape_np[..., 1] = np.where(bad, eta + 0.5, tape_np[..., 1])          # wire out of range(eta)
  tape_np[3, :, 0] = 3.2                                               # unknown gate
  tape = torch.as_tensor(tape_np, device=cuda_device)
  a = EnvBatch(N, eta, delta, objective, target); b = EnvBatch(N, eta, delta, objective, target)
  ora = OracleBatch(N, eta, delta, objective, target)
  rec = {'status': torch.zeros((T, N), dtype=torch.int8, device=cuda_device)}
  b.replay(tape, rec)
  for t in range(T):
    a.step(tape[t]); ora.step(tape_np[t])
    np.testing.assert_array_equal(rec['status'][t].cpu().numpy(), ora.status)
  assert torch.equal(a.state, b.state) and torch.equal(a.meta, b.meta)
  sa, sb = a.stats.cpu().numpy(), b.stats.cpu().numpy()
  assert sb[11] == sa[11] == ora.stats[11] > 0 and sb[10] == sa[10] == ora.stats[10]
  np.testing.assert_allclose(sb, sa, rtol=1e-12, atol=1e-9)
  _check(b, ora, what="replay with invalid actions")


@pytest.mark.parametrize("objective,eta,delta", [('UC-random', 5, 20), ('SP-random', 9, 27), ('UC-random', 6, 24)])
@pytest.mark.parametrize("mode", ["no_autoreset", "carried_return", "dense_long"])
def test_replay_finalize_paths_equal_repeated_steps(cuda_device, objective, eta, delta, mode):
  """replay_reg_kernel finalizes the metric steps of a chunk one per thread; what depends on earlier steps of the
  episode goes through a sequential pass first.  The cases the other replay tests do not reach: no auto-reset (the
  return keeps accumulating after an episode ended), a sparse-reward batch that carries a non-zero return into the
  launch, and a dense reward over a tape longer than one 128-step chunk.  Everything must equal T qcd_step calls."""
  from qcd_b200.batch import EnvBatch
  rng = np.random.default_rng(zlib.crc32(repr((objective, eta, mode)).encode()))
  N, T = 7, (300 if mode == "dense_long" else 40)
  kw = dict(punish=True, sparse=mode != "dense_long", autoreset=mode != "no_autoreset")
  target = make_target(objective, eta, seed=5)
  a = EnvBatch(N, eta, delta, objective, target, **kw)
  b = EnvBatch(N, eta, delta, objective, target, **kw)
  if mode == "carried_return":
    for x in (a, b):
      x.ep_return.fill_(0.375); x.ep_length.fill_(3)
  tape = torch.as_tensor(uniform_box_tape(rng, eta, T, N), device=cuda_device)
  rec = {k: torch.zeros((T, N), dtype=dt, device=cuda_device) for k, dt in
         [('reward', torch.float64), ('metric', torch.float64), ('cost', torch.float64), ('depth', torch.int32),
          ('terminated', torch.uint8), ('truncated', torch.uint8)]}
  b.replay(tape, rec)
  for t in range(T):
    a.step(tape[t])
    for k in rec:
      x, y = rec[k][t], getattr(a, k)
      if x.dtype.is_floating_point:
        assert torch.allclose(x, y, atol=1e-13, rtol=0, equal_nan=True), (k, t)
      else:
        assert torch.equal(x, y), (k, t)
  assert torch.equal(a.state, b.state) and torch.equal(a.meta, b.meta) and torch.equal(a.obs, b.obs)
  assert torch.equal(a.ep_length, b.ep_length) and torch.equal(a.episode_length, b.episode_length)
  assert torch.allclose(a.ep_return, b.ep_return, atol=1e-12, rtol=0)
  assert torch.allclose(a.episode_return, b.episode_return, atol=1e-12, rtol=0)
  assert torch.allclose(a.last, b.last, atol=1e-13, rtol=0)
  sa, sb = a.fold_stats().cpu().numpy(), b.fold_stats().cpu().numpy()
  np.testing.assert_allclose(sb, sa, rtol=1e-12, atol=1e-9)
  assert sa[0] > 0                                     # episodes did finish
