"""The reference's own tests (test/bell.py, ghz.py, hadamard.py, toffoli.py) run through the same
public boundary — gym.make("CircuitDesigner-v0", ...) — on the CUDA path, plus the single-env API
against the oracle step by step, and the vectorised env's host-buffer call."""
import numpy as np
import pytest
import torch

from oracle.qcd_oracle import OracleCircuitDesigner, sample_actions
from tests.test_oracle_kats import KATS

pytestmark = pytest.mark.gpu


def _make(**kw):
  import qcd_gym
  from qcd_b200 import gym_compat as gym
  qcd_gym.register_envs()
  return gym.make("CircuitDesigner-v0", **kw)


@pytest.mark.parametrize("name", list(KATS))
def test_reference_kats_through_gym_make(cuda_device, name):
  kw, seq = KATS[name]
  env = _make(**kw)
  env.reset()
  for a in seq[:-1]: env.step(a)
  reward = env.step(seq[-1])[1]
  np.testing.assert_almost_equal(reward, 1)


def test_step_before_reset_raises(cuda_device):
  from qcd_b200.gym_compat import ResetNeeded
  env = _make(max_qubits=2, max_depth=12, objective='SP-bell')
  with pytest.raises(ResetNeeded): env.step([0, 0, 0, 0.1])


@pytest.mark.parametrize("objective,eta,delta", [('SP-bell', 2, 12), ('SP-random', 3, 9), ('UC-toffoli', 3, 12), ('UC-random', 2, 8)])
@pytest.mark.parametrize("sparse,punish", [(True, True), (False, False)])
def test_single_env_matches_oracle(cuda_device, objective, eta, delta, sparse, punish):
  env = _make(max_qubits=eta, max_depth=delta, objective=objective, sparse=sparse, punish=punish, seed=7)
  ora = OracleCircuitDesigner(eta, delta, objective, sparse=sparse, punish=punish, seed=7)
  assert env.unwrapped.name == ora.name and env.unwrapped.max_steps == ora.max_steps
  np.testing.assert_array_equal(env.unwrapped.target, ora.target)
  rng = np.random.default_rng(0)
  obs, info = env.reset(); o2, i2 = ora.reset()
  assert obs.dtype == np.float32 and obs.shape == env.observation_space.shape == o2.shape
  np.testing.assert_array_equal(obs, o2); assert info == i2
  episodes = 0
  for t in range(120):
    a = sample_actions(rng, eta)
    obs, r, term, trunc, info = env.step(a); o2, r2, t2, u2, i2 = ora.step(a)
    assert (term, trunc) == (t2, u2) and isinstance(term, bool) and isinstance(r, float)
    assert set(info) == set(i2)
    for k in ('depth', 'operations', 'used_wires'): assert info[k] == i2[k]
    if 'termination_reason' in i2: assert info['termination_reason'] == i2['termination_reason']
    np.testing.assert_allclose([r, info['metric'], info['cost']], [r2, i2['metric'], i2['cost']], atol=1e-12, rtol=0)
    np.testing.assert_allclose(obs, o2, atol=1.2e-7, rtol=0)
    if term or trunc:
      episodes += 1
      env.reset(); ora.reset()
  assert episodes > 3


def test_single_env_errors_like_reference(cuda_device):
  env = _make(max_qubits=3, max_depth=9, objective='SP-ghz3'); env.reset()
  for bad in ([0, 3, 0, 0], [0, 0, -0.5, 0], [3, 0, 0, 0]):
    with pytest.raises(AssertionError): env.step(bad)
  with pytest.raises(ValueError): _make(max_qubits=4, max_depth=9, objective='SP-ghz3')   # size mismatch (env.py:109)
  with pytest.raises(AssertionError): _make(max_qubits=2, max_depth=9, objective='UC-toffoli')
  with pytest.raises(AssertionError): _make(max_qubits=2, max_depth=9, objective='XX-foo')


def test_monitor_episode_info(cuda_device):
  from qcd_gym.wrappers.monitor import Monitor
  kw, seq = KATS['bell']
  env = Monitor(_make(**kw))
  with pytest.raises(RuntimeError): env.step(seq[0])
  env.reset()
  for a in seq: out = env.step(a)
  ep = out[4]['episode']
  assert ep['l'] == 5 and ep['d'] == 4 and ep['o'] == 4 and ep['q'] == 2 and ep['c'] == 0
  np.testing.assert_almost_equal(ep['r'], 1); np.testing.assert_almost_equal(ep['m'], 1)
  assert env.termination_reasons == ['DONE'] and 'termination_reason' not in out[4]
  with pytest.raises(RuntimeError): env.step(seq[0])


def test_render_text(cuda_device):
  kw, seq = KATS['bell']
  env = _make(**kw, render_mode='text'); env.reset()
  for a in seq[:4]: env.step(a)
  txt = env.render()
  assert txt.count('\n') == 1 and 'Rx(1.57)' in txt and '■' in txt and 'X' in txt


def test_render_image_is_an_rgb_array(cuda_device):
  kw, seq = KATS['bell']
  env = _make(**kw, render_mode='image'); env.reset()
  empty = env.render()
  for a in seq[:4]: env.step(a)
  img = env.render()
  assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 3 and img.shape[0] == 2 * 40 + 16
  assert img.shape[1] > empty.shape[1] and (img != 255).any()          # four gate columns were drawn


def test_vector_env_host_path_and_semantics(cuda_device):
  from qcd_b200 import CircuitDesignerVectorEnv
  from oracle.batch_oracle import OracleBatch
  N, eta, delta = 512, 5, 20
  venv = CircuitDesignerVectorEnv(N, eta, delta, 'SP-ghz5', final_obs=True)
  ora = OracleBatch(N, eta, delta, 'SP-ghz5', venv.target, final_obs=True)
  obs, info = venv.reset()
  assert obs.shape == (N, 128) and obs.dtype == torch.float32
  np.testing.assert_array_equal(obs.cpu().numpy(), ora.obs)
  rng = np.random.default_rng(1)
  for t in range(25):
    a = sample_actions(rng, eta, N)
    if t % 2:
      obs_h, r, term, trunc = venv.step_host(a)
      obs_np = obs_h
    else:
      obs, r, term, trunc, info = venv.step(torch.as_tensor(a, device=cuda_device))
      assert term.dtype == torch.bool and r.dtype == torch.float64
      obs_np, r, term, trunc = obs.cpu().numpy(), r.cpu().numpy(), term.cpu().numpy(), trunc.cpu().numpy()
    ora.step(a)
    np.testing.assert_allclose(obs_np, ora.obs, atol=1.2e-7, rtol=0)
    np.testing.assert_allclose(r, ora.reward, atol=1e-12, rtol=0)
    np.testing.assert_array_equal(term, ora.terminated.astype(bool)); np.testing.assert_array_equal(trunc, ora.truncated.astype(bool))
    venv.accumulate_stats()
  st = venv.episode_stats()
  assert st['steps'] == 25 * N and st['episodes'] > 0 and st['episodes'] == st['n_done'] + st['n_depth']
  h2d, d2h = venv.host_bytes_per_step()
  assert h2d == N * 16 and d2h == N * 10 + N * 64 * 4


def test_vector_env_raise_on_invalid(cuda_device):
  from qcd_b200 import CircuitDesignerVectorEnv
  venv = CircuitDesignerVectorEnv(4, 2, 12, 'SP-bell', raise_on_invalid=True); venv.reset()
  a = torch.tensor([[0, 0, 0, 0.1]] * 3 + [[0, 2, 0, 0.1]], device=cuda_device)
  with pytest.raises(AssertionError): venv.step(a)


def test_smoke_without_programmatic_dependent_launch(cuda_device):
  """QCD_PDL=0 (plain stream-ordered launches) must give the same results: run smoke() in a fresh
  process with the switch off."""
  import os, subprocess, sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  env = dict(os.environ, QCD_PDL="0")
  res = subprocess.run([sys.executable, os.path.join(root, "__graft_entry__.py"), "smoke"], capture_output=True,
                       text=True, timeout=300, cwd=root, env=env)
  assert res.returncode == 0, res.stderr[-1500:]
  assert res.stdout.count("match the oracle") == 3


def test_step_on_side_stream_and_inside_cuda_graph(cuda_device):
  """The C-ABI calls are asynchronous on the caller's stream and capturable: stepping on a side stream
  and replaying a captured graph of steps must give exactly what eager stepping gives."""
  from qcd_b200.batch import EnvBatch
  from qcd_b200.targets import make_target
  N, eta, delta, T = 3000, 5, 20, 12
  target = make_target('SP-ghz5', eta)
  rng = np.random.default_rng(4)
  tape = torch.as_tensor(np.stack([sample_actions(rng, eta, N) for _ in range(T)]), device=cuda_device)
  eager = EnvBatch(N, eta, delta, 'SP-ghz5', target, device=cuda_device)
  for t in range(T): eager.step(tape[t])
  side = EnvBatch(N, eta, delta, 'SP-ghz5', target, device=cuda_device)
  graphed = EnvBatch(N, eta, delta, 'SP-ghz5', target, device=cuda_device)
  torch.cuda.synchronize()
  s = torch.cuda.Stream(cuda_device)
  with torch.cuda.stream(s):
    for t in range(T): side.step(tape[t])
    g = torch.cuda.CUDAGraph()
    graphed.step(tape[0]); graphed.reset(); s.synchronize()      # warm-up outside capture
    with torch.cuda.graph(g, stream=s):
      for t in range(T): graphed.step(tape[t])
    graphed.reset(); graphed.zero_stats()
    g.replay()
  s.synchronize()
  for other in (side, graphed):
    for name in ('state', 'meta', 'obs', 'reward', 'depth', 'terminated', 'truncated', 'ep_return', 'ep_length'):
      assert torch.equal(getattr(eager, name), getattr(other, name)), name
    assert torch.allclose(eager.stats, other.stats, rtol=1e-12, atol=1e-9)
