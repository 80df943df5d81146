"""SURVEY §8f next rows on the CUDA path: SB3-style VecEnv facade, device-side random baseline,
whole-circuit fitness for the GA baseline (replay kernel)."""
import numpy as np
import pytest
import torch

from oracle.qcd_oracle import OracleCircuitDesigner, circuit_depth, make_target, sample_actions, simulate_operator, simulate_statevector

pytestmark = pytest.mark.gpu


def test_sb3_vec_env_facade_matches_monitor_semantics(cuda_device):
  from qcd_b200.sb3_vec_env import CircuitDesignerSB3VecEnv
  n, eta, delta = 4, 3, 15
  vec = CircuitDesignerSB3VecEnv(n, eta, delta, 'SP-ghz3', seed=3)
  oras = [OracleCircuitDesigner(eta, delta, 'SP-ghz3', seed=3) for _ in range(n)]      # same seed for all (algorithm.py:25-27)
  assert vec.envs[0].unwrapped.name == oras[0].name
  obs = vec.reset()
  assert obs.shape == (n, 32) and obs.dtype == np.float32
  for o in oras: o.reset()
  rng = np.random.default_rng(5)
  ep_r = np.zeros(n); ep_l = np.zeros(n, int); episodes = 0
  for t in range(80):
    a = sample_actions(rng, eta, n)
    obs, rew, dones, infos = vec.step(a)
    for i, o in enumerate(oras):
      o2, r2, term, trunc, info = o.step(a[i])
      ep_r[i] += r2; ep_l[i] += 1
      assert dones[i] == (term or trunc) and abs(rew[i] - r2) < 1e-6
      assert infos[i]['depth'] == info['depth'] and infos[i]['TimeLimit.truncated'] == (trunc and not term)
      if term or trunc:
        episodes += 1
        np.testing.assert_allclose(infos[i]['terminal_observation'], o2, atol=1.2e-7)
        ep = infos[i]['episode']
        assert ep['l'] == ep_l[i] and abs(ep['r'] - ep_r[i]) < 1e-6 and ep['d'] == info['depth'] and ep['q'] == info['used_wires']
        np.testing.assert_allclose([ep['m'], ep['c']], [info['metric'], info['cost']], atol=1e-12)
        assert vec.envs[i].termination_reasons[-1] == info['termination_reason']
        ep_r[i] = 0; ep_l[i] = 0
        o2, _ = o.reset()
      np.testing.assert_allclose(obs[i], o2, atol=1.2e-7)
  assert episodes > 10


def test_random_baseline_statistics(cuda_device):
  from qcd_b200.baselines import random_baseline
  res = random_baseline('SP-bell-q2-d12', seeds=4, episodes=400)
  assert set(res) == {'Return', 'Depth', 'Qubits', 'Metric', 'Cost'} and res['Return'].shape == (4,)
  # CPU replica of plot/metrics.py:161-176 with the oracle (different random stream, same distribution)
  env = OracleCircuitDesigner(2, 12, 'SP-bell', seed=42)
  rng = np.random.default_rng(0)
  m, d = [], []
  for _ in range(1500):
    env.reset(); done = False
    while not done:
      _, r, term, trunc, info = env.step(sample_actions(rng, 2))
      done = term or trunc
    m.append(info['metric']); d.append(info['depth'])
  assert abs(res['Metric'].mean() - np.mean(m)) < 0.03 and abs(res['Depth'].mean() - np.mean(d)) < 0.2
  assert np.all(res['Return'] <= res['Metric'] + 1e-12)


def test_circuit_fitness_matches_whole_circuit_evaluation(cuda_device):
  from qcd_b200.baselines import circuit_fitness
  rng = np.random.default_rng(8)
  for objective, eta, delta in (('SP-ghz3', 3, 15), ('UC-toffoli', 3, 30), ('SP-random', 9, 27)):
    target = make_target(objective, eta, seed=4)
    circuits = []
    for _ in range(40):
      gates = []
      for _ in range(rng.integers(0, 12)):
        kind = ['p', 'rx', 'cp', 'cx'][rng.integers(4)]
        w = int(rng.integers(eta)); c = int((w + 1 + rng.integers(eta - 1)) % eta)
        gates.append((kind, w, c, float(rng.uniform(-np.pi, np.pi)) if kind != 'cx' else None))
      circuits.append(gates)
    got = circuit_fitness(objective, eta, delta, circuits, target=target)
    for gates, f in zip(circuits, got):
      ops = [(k, [w] if k in ('p', 'rx') else [c, w], th) for k, w, c, th in gates]
      if 'SP' in objective:
        R = abs(np.vdot(simulate_statevector(ops, eta), target)) ** 2              # evo/run.py:17-20
      else:
        R = 1 - 2 * np.arctan(np.linalg.norm(target - simulate_operator(ops, eta))) / np.pi   # evo/run.py:22-25
      C = max(0, circuit_depth(ops, eta) - delta / 3) / (delta / 2 * 3)           # evo/run.py:30
      assert abs(f - (R - C)) < 1e-12
