"""Property tests of the oracles (hypothesis): invariants the reference's semantics imply, checked on
generated action sequences — unit norm / unitarity, metric range, depth monotonicity inside an episode,
the telescoping sum of dense rewards, and agreement of the C batch oracle with the einsum oracle."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle.batch_oracle import OracleBatch
from oracle.qcd_oracle import OracleCircuitDesigner, make_target

f32 = lambda lo, hi: st.floats(min_value=float(np.float32(lo)), max_value=float(np.float32(hi)), width=32, allow_nan=False)


def actions(eta, n):
  one = st.tuples(f32(0, 2.9999), f32(0, eta - 1e-3), f32(0, eta - 1e-3), f32(-3.1415925, 3.1415925))
  return st.lists(one, min_size=1, max_size=n)


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 4).flatmap(lambda eta: st.tuples(st.just(eta), actions(eta, 20))), st.booleans())
def test_state_invariants_and_c_oracle_agreement(case, uc):
  eta, seq = case
  eta = min(eta, 3) if uc else eta
  objective = 'UC-random' if uc else 'SP-random'
  delta = 9
  target = make_target(objective, eta, seed=3)
  env = OracleCircuitDesigner(eta, delta, objective, sparse=False, punish=False, target=target)
  cb = OracleBatch(1, eta, delta, objective, target, sparse=False, punish=False, autoreset=False)
  env.reset()
  total, last_depth = 0.0, 0
  for a in seq:
    a = np.asarray([min(a[0], 2.9999), min(a[1], eta - 1e-3), min(a[2], eta - 1e-3), a[3]], np.float32)
    obs, r, term, trunc, info = env.step(a)
    cb.step(a[None])
    s = env.raw_state()
    if uc:
      np.testing.assert_allclose(s.conj().T @ s, np.eye(2 ** eta), atol=1e-13)
    else:
      np.testing.assert_allclose(np.vdot(s, s).real, 1.0, atol=1e-13)
    assert info['depth'] >= last_depth and info['operations'] >= info['depth'] >= (1 if info['operations'] else 0)
    last_depth = info['depth']
    total += r
    assert np.all(np.abs(obs) <= 1 + 1e-6) and obs.dtype == np.float32
    np.testing.assert_allclose(cb.state[0], s.reshape(-1), atol=5e-15)
    assert (cb.depth[0], cb.operations[0], cb.used_wires[0]) == (info['depth'], info['operations'], info['used_wires'])
    assert (bool(cb.terminated[0]), bool(cb.truncated[0])) == (term, trunc)
    np.testing.assert_allclose(cb.reward[0], r, atol=5e-15)
  # dense rewards telescope to the last raw metric (env.py:102-104,113)
  raw = abs(np.vdot(env.raw_state(), target)) ** 2 if not uc else \
      1 - 2 * np.arctan(np.linalg.norm(target - env.raw_state())) / np.pi
  np.testing.assert_allclose(total, raw, atol=1e-12)
