"""Targets of CircuitDesigner._target (/root/reference/qcd_gym/env.py:52-70) as complex128 numpy arrays,
and the task-string grammar of algorithm/algorithm.py:20-23.  Host-side, runs once per env batch."""
from __future__ import annotations

import re

import numpy as np


def parse_task(task: str) -> dict:
  """'SP-ghz3-q3-d15' -> gym.make kwargs (algorithm/algorithm.py:20-23 `named`)."""
  max_qubits = int(re.search(r'-q(\d+)', task).group(1)); task = re.sub(r'-q(\d+)', '', task)
  max_depth = int(re.search(r'-d(\d+)', task).group(1)); task = re.sub(r'-d(\d+)', '', task)
  return {'id': 'CircuitDesigner-v0', 'max_qubits': max_qubits, 'max_depth': max_depth, 'objective': task}


def _on_low_qubits(gate: np.ndarray, n_qubits: int) -> np.ndarray:
  """Operator of a circuit with `gate` on qubits 0..k-1 (little-endian): I (x) gate."""
  k = int(np.log2(gate.shape[0]))
  return np.kron(np.eye(2 ** (n_qubits - k), dtype=np.complex128), gate.astype(np.complex128))


def make_target(objective: str, n_qubits: int, seed=None) -> np.ndarray:
  """complex128 [D] for 'SP-*', [D,D] for 'UC-*' (env.py:52-70)."""
  task, name = objective.split('-')
  dim = 2 ** n_qubits
  if task == 'SP':
    if name == 'random':        # qiskit.quantum_info.random_statevector(dim, seed)
      rng = seed if isinstance(seed, np.random.Generator) else np.random.default_rng(seed)
      vec = rng.standard_normal(dim).astype(complex)
      vec += 1j * rng.standard_normal(dim)
      vec /= np.linalg.norm(vec)
      return vec
    if name == 'bell':
      return np.array([1 / np.sqrt(2), 0, 0, 1 / np.sqrt(2)], dtype=np.complex128)
    if 'ghz' in name:
      m = int(name[-1])
      assert 2 <= m <= n_qubits, f"GHZ entangled state must have at least 2 and at most {n_qubits} qubits."
      vec = np.zeros(2 ** m, dtype=np.complex128)
      vec[0] = vec[-1] = 1 / np.sqrt(2)
      return vec
  if task == 'UC':
    if name == 'random':        # qiskit.quantum_info.random_unitary(dim, seed)
      from scipy import stats
      rng = seed if isinstance(seed, np.random.Generator) else np.random.default_rng(seed)
      return np.asarray(stats.unitary_group.rvs(dim, random_state=rng), dtype=np.complex128).reshape(dim, dim)
    if name == 'hadamard':
      return _on_low_qubits(np.array([[1, 1], [1, -1]]) / np.sqrt(2), n_qubits)
    if name == 'toffoli':
      assert n_qubits >= 3, "to build Toffoli gate you need at least three wires/qubits."
      ccx = np.eye(8)
      ccx[[3, 7]] = ccx[[7, 3]]      # controls q0,q1 (bits 0,1), target q2 (bit 2)
      return _on_low_qubits(ccx, n_qubits)
  assert False, f'{task}-{name} not defined.'
