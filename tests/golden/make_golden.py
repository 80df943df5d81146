"""Regenerates tests/golden/*.npz.

The reference (qcd_gym/env.py on qiskit 1.0.2) cannot be imported in this image, so these vectors
are produced by oracle/qcd_oracle.py (the einsum restatement pinned on the reference's four KATs) —
they are REGRESSION fixtures for the unpinned part of the semantics (observations, depth, costs,
dense rewards), not independent ground truth.  `kats.npz` holds the reference's own test tapes
(test/bell.py, ghz.py, hadamard.py, toffoli.py) with the per-step values the oracle produces and the
terminal reward the reference asserts (== 1 to 7 decimals).

  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.qcd_oracle import OracleCircuitDesigner, make_target, sample_actions  # noqa: E402
from tests.test_oracle_kats import KATS  # noqa: E402

CASES = [('SP-bell', 2, 12, True, True), ('SP-ghz3', 3, 15, True, True), ('SP-ghz5', 5, 20, True, True),
         ('SP-ghz5', 5, 20, False, True), ('SP-random', 4, 12, False, False), ('UC-hadamard', 2, 9, True, True),
         ('UC-toffoli', 3, 30, True, True), ('UC-toffoli', 3, 30, False, True), ('UC-random', 2, 8, True, False)]


def rollout(env, tape):
  """Steps `env` through tape[T,4] with reset-on-done; returns per-step arrays."""
  T = len(tape)
  out = {k: [] for k in ('obs', 'reward', 'metric', 'cost', 'depth', 'operations', 'used_wires', 'terminated', 'truncated')}
  env.reset()
  for t in range(T):
    obs, r, term, trunc, info = env.step(tape[t])
    for k, v in zip(out, (obs, r, info['metric'], info['cost'], info['depth'], info['operations'], info['used_wires'], term, trunc)):
      out[k].append(v)
    if term or trunc:
      env.reset()
  return {k: np.asarray(v) for k, v in out.items()}


def main():
  rng = np.random.default_rng(20240229)
  for objective, eta, delta, sparse, punish in CASES:
    N, T = 3, 32
    target = make_target(objective, eta, seed=17)
    tape = np.stack([sample_actions(rng, eta, N) for _ in range(T)])          # [T,N,4] float32
    per_env = [rollout(OracleCircuitDesigner(eta, delta, objective, punish=punish, sparse=sparse, target=target), tape[:, i])
               for i in range(N)]
    arrays = {k: np.stack([p[k] for p in per_env], axis=1) for k in per_env[0]}   # [T,N,...]
    name = f"{objective}_q{eta}_d{delta}_{'sparse' if sparse else 'dense'}_{'punish' if punish else 'nopunish'}.npz"
    np.savez_compressed(os.path.join(HERE, name), tape=tape, target=target, objective=objective, eta=eta, delta=delta,
                        sparse=sparse, punish=punish, **arrays)
    print(name, {k: v.shape for k, v in arrays.items()}['obs'])
  kat = {}
  for name, (kw, seq) in KATS.items():
    env = OracleCircuitDesigner(**kw)
    r = rollout(env, np.asarray(seq, dtype=np.float64))
    kat[f"{name}_tape"] = np.asarray(seq, dtype=np.float64)
    for k, v in r.items():
      kat[f"{name}_{k}"] = v
    kat[f"{name}_kwargs"] = np.array([kw['max_qubits'], kw['max_depth']])
    kat[f"{name}_objective"] = kw['objective']
  np.savez_compressed(os.path.join(HERE, "kats.npz"), **kat)
  print("kats.npz", sorted(k for k in kat if k.endswith('_reward')))


if __name__ == "__main__":
  main()
