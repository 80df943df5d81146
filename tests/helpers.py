"""Shared helpers for the parity tests: action tapes drawn like the reference's sampler."""
import numpy as np

from oracle.qcd_oracle import action_bounds


def uniform_box_tape(rng, n_qubits, T, N, dtype=np.float32, no_terminate=False):
  """[T,N,4] actions ~ Box.sample() (plot/metrics.py:174 `env.action_space.sample()`)."""
  low, high = action_bounds(n_qubits)
  high = high.copy()
  if no_terminate:
    high[0] = np.float32(2 - 1e-5)
  return rng.uniform(low=low, high=high, size=(T, N, 4)).astype(np.float32).astype(dtype)


CONFIGS = [  # (objective, eta, delta)  — BASELINE.json configs + the reference's own test configs
    ('SP-bell', 2, 12), ('SP-ghz3', 3, 15), ('SP-ghz5', 5, 20), ('SP-random', 1, 6), ('SP-random', 4, 12),
    ('SP-random', 7, 20), ('UC-hadamard', 1, 9), ('UC-hadamard', 2, 9), ('UC-toffoli', 3, 30), ('UC-random', 4, 16),
]
