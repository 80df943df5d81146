"""CPU oracle for the CircuitDesigner hot path (TEST INFRASTRUCTURE, not product code).

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import this module.  Nothing under `qcd_b200/` imports it.

What it restates
----------------
* `/root/reference/qcd_gym/env.py` (whole file): action decode (env.py:90-100), the
  from-scratch re-simulation of the circuit on every step (env.py:80-87), truncation
  (env.py:129), fidelity / arctan-Frobenius metric and depth cost (env.py:106-114),
  sparse / punish shaping (env.py:134-135), reset (env.py:117-121), targets (env.py:52-70).
* The arithmetic env.py reaches in the un-vendored third-party dependency
  `qiskit==1.0.2` (pinned at /root/reference/setup.py:14), restated from that release's
  published algorithm:
    - `quantum_info.Statevector.from_instruction` -> `_evolve_instruction` -> `_evolve_operator`
      (start at e0; per gate: reshape to a rank-n tensor, TRANSPOSE the gate's axes to the front,
      `np.dot(gate, .)`, transpose back; little-endian qargs)
                                                   -> `simulate_statevector`
    - `quantum_info.Operator._einsum_matmul`        -> `_einsum_matmul`
    - `quantum_info.Operator(QuantumCircuit)` / `_append_instruction` / `compose`
      (start at I, LEFT multiplication gate @ V through `_einsum_matmul`)
                                                   -> `simulate_operator`
    - `circuit.library.{PhaseGate,RXGate,CPhaseGate,CXGate,HGate,CCXGate}.__array__`
                                                   -> `gate_matrix`, `_H`, `_CCX`
    - `QuantumCircuit.depth / count_ops`, `circuit_to_dag(qc).idle_wires()`
                                                   -> `circuit_depth`, `len(ops)`, `used_wires`
    - `quantum_info.random_statevector / random_unitary` -> `make_target`
  and `gymnasium==0.29` `spaces.Box` bounds/sampling -> `action_bounds`, `sample_actions`.

Parity status
-------------
qiskit and gymnasium are NOT installable in this image (no wheel, no network), so the
oracle cannot be run against the reference itself.  It is pinned against every golden
vector the reference's own tests hold for this path (test/bell.py:8-18, test/ghz.py:8-16,
test/hadamard.py:7-21, test/toffoli.py:8-30: terminal reward == 1 to 7 decimals), see
`tests/test_oracle_kats.py`.  Everything those four KATs do not constrain (observation
layout, depth/ops/used_wires values, truncation, cost > 0, non-sparse deltas, random
targets, error behaviour) follows the qiskit-1.0.2 / gymnasium-0.29 rules listed in
SURVEY.md §8c and is **parity unpinned**.
"""
from __future__ import annotations

import cmath
import math

import numpy as np

GATES = 3  # env.py:11

# gate kinds used in the op list: ('p', [w], theta), ('rx', [w], theta),
# ('cp', [c, w], theta), ('cx', [c, w], None)


# --------------------------------------------------------------------------- gate matrices
def gate_matrix(kind: str, theta) -> np.ndarray:
  """qiskit 1.0.2 `__array__` of the four gates env.py:95-98 constructs.

  PhaseGate:  [[1,0],[0,exp(1j*lam)]] with cmath.exp;   RXGate: cos/sin of theta/2 via math;
  CPhaseGate: diag(1,1,1,exp(1j*theta));                CXGate: control = qargs[0] (LSB).
  """
  if kind == 'p':
    lam = float(theta)
    return np.array([[1, 0], [0, cmath.exp(1j * lam)]], dtype=complex)
  if kind == 'rx':
    cos = math.cos(float(theta) / 2)
    sin = math.sin(float(theta) / 2)
    return np.array([[cos, -1j * sin], [-1j * sin, cos]], dtype=complex)
  if kind == 'cp':
    eith = cmath.exp(1j * float(theta))
    return np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, eith]], dtype=complex)
  if kind == 'cx':
    return np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=complex)
  raise ValueError(kind)


_H = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)   # HGate.__array__
_CCX = np.eye(8, dtype=complex)                                # CCXGate: controls q0,q1 target q2
_CCX[[3, 7]] = _CCX[[7, 3]]


# --------------------------------------------------------------------------- qiskit einsum core
def _einsum_matmul(tensor, mat, indices, shift=0, right_mul=False):
  """qiskit 1.0.2 `Operator._einsum_matmul`: contract `mat` into axes `indices` of `tensor`."""
  rank = tensor.ndim
  rank_mat = mat.ndim
  if rank_mat % 2 != 0:
    raise ValueError("Contracted matrix must have an even number of indices.")
  indices_tensor = list(range(rank))
  for j, index in enumerate(indices):
    indices_tensor[index + shift] = rank + j
  mat_contract = list(reversed(range(rank, rank + len(indices))))
  mat_free = [index + shift for index in reversed(indices)]
  indices_mat = mat_contract + mat_free if right_mul else mat_free + mat_contract
  return np.einsum(tensor, indices_tensor, mat, indices_mat)


def simulate_statevector(ops, n: int) -> np.ndarray:
  """`Statevector.from_instruction(qc)` (env.py:84,109): e0, then per gate `Statevector._evolve_operator`
  of qiskit 1.0.2 - NOT an einsum: the state is reshaped to (2,)*n, the gate's axes are transposed to the
  front (`indices = [n-1-q for q in reversed(qargs)]`), contracted with `np.dot(gate, .)` as a
  (2^k, 2^(n-k)) matrix, and transposed back."""
  psi = np.zeros(2 ** n, dtype=complex)
  psi[0] = 1.0
  for kind, qargs, theta in ops:
    g = gate_matrix(kind, theta)
    indices = [n - 1 - q for q in reversed(qargs)]
    axes = indices + [i for i in range(n) if i not in indices]
    axes_inv = np.argsort(axes).tolist()
    contract_dim = g.shape[1]
    data = np.reshape(np.transpose(np.reshape(psi, (2,) * n), axes), (contract_dim, 2 ** n // contract_dim))
    data = np.reshape(np.dot(g, data), (2,) * n)
    psi = np.reshape(np.transpose(data, axes_inv), 2 ** n)
  return psi


def simulate_operator(ops, n: int) -> np.ndarray:
  """`Operator(qc)`: identity, then per gate V <- G_full @ V (compose, front=False) (env.py:83,111)."""
  dim = 2 ** n
  v = np.eye(dim, dtype=complex)
  for kind, qargs, theta in ops:
    g = gate_matrix(kind, theta)
    k = len(qargs)
    tensor = np.reshape(v, (2,) * (2 * n))
    mat = np.reshape(g, (2,) * (2 * k))
    indices = [n - 1 - q for q in qargs]      # num_indices = n output dims, shift = 0
    v = np.reshape(_einsum_matmul(tensor, mat, indices, 0, False), (dim, dim))
  return v


def embed_operator(g: np.ndarray, qargs, n: int) -> np.ndarray:
  """Operator of a circuit holding the single gate `g` on `qargs` (targets, env.py:63-69)."""
  dim = 2 ** n
  k = len(qargs)
  tensor = np.reshape(np.eye(dim, dtype=complex), (2,) * (2 * n))
  mat = np.reshape(g, (2,) * (2 * k))
  indices = [n - 1 - q for q in qargs]
  return np.reshape(_einsum_matmul(tensor, mat, indices, 0, False), (dim, dim))


# --------------------------------------------------------------------------- circuit bookkeeping
def circuit_depth(ops, n: int) -> int:
  """`QuantumCircuit.depth()`: per-qubit stack, every instruction lands at max(level)+1."""
  stack = [0] * n
  for _, qargs, _ in ops:
    level = max(stack[q] + 1 for q in qargs)
    for q in qargs:
      stack[q] = level
  return max(stack) if stack else 0


def used_wires(ops, n: int) -> int:
  """η − |idle_wires(circuit_to_dag(qc))| (env.py:77)."""
  touched = set()
  for _, qargs, _ in ops:
    touched.update(int(q) for q in qargs)
  return len(touched)


# --------------------------------------------------------------------------- targets
def make_target(objective: str, n: int, seed=None) -> np.ndarray:
  """env.py:52-70.  Returns complex128 [D] (SP) or [D,D] (UC)."""
  task, target = objective.split('-')
  dim = 2 ** n
  if task == 'SP':
    if target == 'random':      # qiskit random_statevector
      rng = seed if isinstance(seed, np.random.Generator) else np.random.default_rng(seed)
      vec = rng.standard_normal(dim).astype(complex)
      vec += 1j * rng.standard_normal(dim)
      vec /= np.linalg.norm(vec)
      return vec
    if target == 'bell':
      return np.array([1 / np.sqrt(2), 0, 0, 1 / np.sqrt(2)], dtype=np.complex128)
    if 'ghz' in target:
      m = int(target[-1])
      assert 2 <= m <= n, f"GHZ entangled state must have at least 2 and at most {n} qubits."
      vec = np.zeros(shape=(2 ** m,), dtype=np.complex128)
      vec[0] = vec[-1] = 1 / np.sqrt(2)
      return vec
  if task == 'UC':
    if target == 'random':      # qiskit random_unitary -> scipy unitary_group
      from scipy import stats
      rng = seed if isinstance(seed, np.random.Generator) else np.random.default_rng(seed)
      return np.asarray(stats.unitary_group.rvs(dim, random_state=rng), dtype=complex).reshape(dim, dim)
    if target == 'hadamard':
      return embed_operator(_H, [0], n)
    if target == 'toffoli':
      assert n >= 3, "to build Toffoli gate you need at least three wires/qubits."
      return embed_operator(_CCX, [0, 1, 2], n)
  assert False, f'{task}-{target} not defined.'


def flat(data: np.ndarray) -> np.ndarray:
  """env.py:10 — concat(real, imag) along axis 0, float32 (RNE), C-order flatten."""
  return np.concatenate([data.real, data.imag]).astype(np.float32).flatten()


# --------------------------------------------------------------------------- action space
def action_bounds(n: int):
  """gymnasium 0.29 Box(low, high) as built at env.py:46-49; bounds are stored as float32."""
  m = 1e-5
  low = np.array([0, 0, 0, -np.pi]).astype(np.float32)
  high = np.array([GATES - m, n - m, n - m, np.pi]).astype(np.float32)
  return low, high


def sample_actions(rng: np.random.Generator, n: int, size=None) -> np.ndarray:
  """Box.sample() for a bounded Box: uniform(low, high) in float64, cast to float32."""
  low, high = action_bounds(n)
  shape = (4,) if size is None else (size, 4)
  return rng.uniform(low=low, high=high, size=shape).astype(np.float32)


# --------------------------------------------------------------------------- the env
class OracleCircuitDesigner:
  """Restatement of `CircuitDesigner` (env.py:13-141) on the einsum core above.

  Follows the reference's per-step ALGORITHM (re-simulate the whole op list for the
  observation, again for the reward, walk it for depth), so timing it is timing the
  reference's CPU path minus qiskit object overhead.
  """

  def __init__(self, max_qubits: int, max_depth: int, objective: str,
               punish=True, sparse=True, seed=None, target=None):
    self.qubits, self.depth = max_qubits, max_depth
    self.max_steps = max_depth * max_qubits * 2                       # env.py:37
    self.punish, self.sparse, self.objective = punish, sparse, objective
    self.name = f"{objective}|{max_qubits}-{max_depth}"
    self.target = make_target(objective, max_qubits, seed) if target is None \
        else np.asarray(target, dtype=complex)
    self._ops = []
    self.last_reward = 0
    self.last_cost = 0

  # -- circuit -> numbers ----------------------------------------------------------------
  def _simulate(self):
    if 'UC' in self.objective:
      return simulate_operator(self._ops, self.qubits)
    return simulate_statevector(self._ops, self.qubits)

  @property
  def _state(self):                                                   # env.py:80-87
    state = flat(self._simulate())
    observation = np.concatenate([state, flat(self.target)])
    info = {'depth': circuit_depth(self._ops, self.qubits), 'operations': len(self._ops),
            'used_wires': used_wires(self._ops, self.qubits)}
    return observation, info

  def _operation(self, action):                                       # env.py:90-100
    gate, wire, cntrl, theta = action
    gate, wire, cntrl = np.floor([gate, wire, cntrl]).astype(int)
    assert wire in range(self.qubits) and cntrl in range(self.qubits), f"{action}"
    if wire == cntrl and gate == 0: return ('p', [wire], theta)
    if wire == cntrl and gate == 1: return ('rx', [wire], theta)
    if gate == 0: return ('cp', [cntrl, wire], theta)
    if gate == 1: return ('cx', [cntrl, wire], None)
    if gate == 2: return None
    assert False, f'Unhandled Action on gate {gate}'

  @property
  def _reward(self):                                                  # env.py:106-114
    if 'SP' in self.objective:
      reward = abs(np.vdot(self._simulate(), self.target)) ** 2
    if 'UC' in self.objective:
      reward = 1 - 2 * np.arctan(np.linalg.norm(self.target - self._simulate())) / np.pi
    cost = (max(0, circuit_depth(self._ops, self.qubits) - self.depth / 3)) / (self.depth / 2 * 3)
    if not self.sparse:                                               # env.py:102-104
      delta = (reward - self.last_reward, cost - self.last_cost)
      self.last_reward, self.last_cost = reward, cost                 # the undifferenced values are kept
      reward, cost = delta
    return reward, cost

  def reset(self, seed=None, options=None):                           # env.py:117-121
    if not self.sparse:
      self.last_reward = 0
      self.last_cost = 0
    self._ops = []
    return self._state

  def step(self, action):                                             # env.py:124-136
    operation = self._operation(action)
    terminated = operation is None
    if not terminated:
      self._ops.append(operation)
    state, info = self._state
    truncated = circuit_depth(self._ops, self.qubits) >= self.depth or len(self._ops) >= self.max_steps
    if terminated: info['termination_reason'] = 'DONE'
    if truncated: info['termination_reason'] = 'DEPTH'
    reward, cost = self._reward
    info = {**info, 'metric': reward, 'cost': cost}
    if self.sparse and not (terminated or truncated): reward, cost = 0, 0
    if self.punish: reward -= cost
    return state, reward, terminated, truncated, info

  # -- extras used by the tests only -----------------------------------------------------
  def raw_state(self) -> np.ndarray:
    """complex128 state/unitary of the current circuit (what the GPU keeps resident)."""
    return self._simulate()
