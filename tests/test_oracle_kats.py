"""Pin the oracle on every golden vector the reference's own tests hold for the hot path
(/root/reference/test/{bell,ghz,hadamard,toffoli}.py) and on the hand-derivable vectors of
SURVEY.md §8c."""
import numpy as np
import pytest

from oracle.qcd_oracle import OracleCircuitDesigner, gate_matrix, simulate_operator

PI = np.pi
H = lambda q: [[0, q, q, PI / 2], [1, q, q, PI / 2], [0, q, q, PI / 2]]       # P RX P = Hadamard

KATS = {
    'bell': (dict(max_qubits=2, max_depth=12, objective='SP-bell'), H(0) + [[1, 1, 0, PI], [2, 0, 0, 0]]),
    'ghz': (dict(max_qubits=3, max_depth=15, objective='SP-ghz3'),
            H(0) + [[1, 1, 0, PI], [1, 2, 1, PI], [2, 0, 0, 0]]),
    'hadamard1': (dict(max_qubits=1, max_depth=9, objective='UC-hadamard'), H(0) + [[2, 0, 0, 0]]),
    'hadamard2': (dict(max_qubits=2, max_depth=9, objective='UC-hadamard'), H(0) + [[2, 0, 0, 0]]),
    'toffoli': (dict(max_qubits=3, max_depth=63, objective='UC-toffoli'),
                H(2) + [[0, 2, 1, PI / 2]] + H(2) + [[1, 1, 0, PI / 2]] + H(2) + [[0, 2, 1, -PI / 2]] + H(2)
                + [[1, 1, 0, PI / 2]] + H(2) + [[0, 2, 0, PI / 2]] + H(2) + [[2, 0, 0, 0]]),
}


@pytest.mark.parametrize("name", list(KATS))
def test_reference_kat_terminal_reward(name):
  kw, seq = KATS[name]
  env = OracleCircuitDesigner(**kw)
  env.reset()
  for a in seq[:-1]:
    _, r, term, trunc, _ = env.step(a)
    assert r == 0 and not term and not trunc        # sparse: nothing until the episode ends
  _, reward, term, trunc, info = env.step(seq[-1])
  np.testing.assert_almost_equal(reward, 1)          # the reference's own assertion
  assert term and not trunc and info['termination_reason'] == 'DONE' and info['cost'] == 0


def test_bell_tape_hand_derived():
  env = OracleCircuitDesigner(2, 12, 'SP-bell')
  obs, info = env.reset()
  assert obs.dtype == np.float32 and obs.shape == (16,)
  assert info == {'depth': 0, 'operations': 0, 'used_wires': 0}
  s = 1 / np.sqrt(2)
  want_states = [[1, 0, 0, 0], [s, -1j * s, 0, 0], [s, s, 0, 0], [s, 0, 0, s]]
  want_metric = [0.5, 0.25, 0.25, 1.0]
  want_used = [1, 1, 1, 2]
  _, seq = KATS['bell']
  for t, a in enumerate(seq[:4]):
    obs, r, term, trunc, info = env.step(a)
    np.testing.assert_allclose(env.raw_state(), want_states[t], atol=1e-15)
    np.testing.assert_allclose(info['metric'], want_metric[t], atol=1e-15)
    assert (info['depth'], info['operations'], info['used_wires']) == (t + 1, t + 1, want_used[t])
    np.testing.assert_array_equal(obs[:8], np.concatenate([env.raw_state().real, env.raw_state().imag]).astype(np.float32))
    np.testing.assert_array_equal(obs[8:], np.array([s, 0, 0, s, 0, 0, 0, 0], np.float32))


def test_toffoli_depth_boundary_and_cost_formula():
  kw, seq = KATS['toffoli']
  env = OracleCircuitDesigner(**kw); env.reset()
  for a in seq: out = env.step(a)
  assert out[4]['depth'] == 21 and out[4]['operations'] == 23 and out[4]['cost'] == 0   # 21 - 63/3 == 0
  env = OracleCircuitDesigner(2, 12, 'SP-bell'); env.reset()
  for t in range(12): out = env.step([1, 0, 0, 0.1])
  assert out[3] and out[4]['depth'] == 12 and out[4]['termination_reason'] == 'DEPTH'
  assert out[4]['cost'] == 8 / 18
  assert out[1] == out[4]['metric'] - out[4]['cost']


def test_h_identity_entrywise():
  v = simulate_operator([('p', [0], PI / 2), ('rx', [0], PI / 2), ('p', [0], PI / 2)], 1)
  np.testing.assert_allclose(v, np.array([[1, 1], [1, -1]]) / np.sqrt(2), atol=1e-15)


def test_cx_roles_little_endian():
  """CX control=cntrl target=wire; index = q1q0.  Swapping the roles breaks bell (fidelity 0.5)."""
  env = OracleCircuitDesigner(2, 12, 'SP-bell'); env.reset()
  for a in H(0) + [[1, 0, 1, PI]]: out = env.step(a)      # control q1 (|0>) -> no flip
  np.testing.assert_allclose(out[4]['metric'], 0.25, atol=1e-15)
  m = gate_matrix('cx', None)
  assert m[3, 1] == 1 and m[1, 3] == 1 and m[0, 0] == 1 and m[2, 2] == 1


def test_decode_errors_and_precedence():
  env = OracleCircuitDesigner(3, 9, 'SP-ghz3'); env.reset()
  for bad in ([0, 3, 0, 0], [0, 0, -0.5, 0], [2, 0, 3.2, 0], [3, 0, 0, 0], [-0.1, 1, 1, 0]):
    with pytest.raises(AssertionError): env.step(bad)
  out = env.step([2.99, 2.99, 0.01, 1.0])                 # floor -> Terminate regardless of wires
  assert out[2] and out[4]['operations'] == 0


def test_dense_reward_deltas_and_no_punish():
  env = OracleCircuitDesigner(2, 12, 'SP-bell', sparse=False, punish=False); env.reset()
  _, seq = KATS['bell']
  total = 0
  for a in seq: total += env.step(a)[1]
  np.testing.assert_allclose(total, 1.0, atol=1e-14)      # telescoping sum of metric deltas


def _dense(kind, qargs, theta, n):
  """Full 2^n x 2^n matrix of one gate built WITHOUT the einsum core: explicit sum over basis states
  with qiskit's documented little-endian convention (qubit k = bit k of the basis index; the gate's
  own matrix is indexed with qargs[0] as its least-significant bit)."""
  g = gate_matrix(kind, theta)
  dim = 2 ** n
  full = np.zeros((dim, dim), complex)
  for col in range(dim):
    sub_in = sum(((col >> q) & 1) << k for k, q in enumerate(qargs))
    rest = col & ~sum(1 << q for q in qargs)
    for sub_out in range(2 ** len(qargs)):
      row = rest | sum(((sub_out >> k) & 1) << q for k, q in enumerate(qargs))
      full[row, col] += g[sub_out, sub_in]
  return full


def test_einsum_core_equals_dense_little_endian_products():
  from oracle.qcd_oracle import simulate_statevector
  rng = np.random.default_rng(12)
  for n in (1, 2, 3, 4):
    ops = []
    for _ in range(12):
      kind = ['p', 'rx', 'cp', 'cx'][rng.integers(4)]
      w = int(rng.integers(n))
      if kind in ('cp', 'cx'):
        if n == 1: continue
        c = int((w + 1 + rng.integers(n - 1)) % n)
        ops.append((kind, [c, w], float(rng.uniform(-PI, PI)) if kind == 'cp' else None))
      else:
        ops.append((kind, [w], float(rng.uniform(-PI, PI))))
    U = np.eye(2 ** n, dtype=complex)
    for kind, qargs, theta in ops:
      U = _dense(kind, qargs, theta, n) @ U                      # left multiplication, circuit order
    np.testing.assert_allclose(simulate_operator(ops, n), U, atol=1e-14)
    np.testing.assert_allclose(simulate_statevector(ops, n), U[:, 0], atol=1e-14)
