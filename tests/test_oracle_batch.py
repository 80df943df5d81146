"""The C restatement of the batched step (oracle/qcd_oracle_batch.c) against the einsum restatement
of the reference (oracle/qcd_oracle.py) on random action tapes drawn like Box.sample()."""
import zlib

import numpy as np
import pytest

from oracle.batch_oracle import OracleBatch
from oracle.qcd_oracle import OracleCircuitDesigner, make_target
from tests.helpers import CONFIGS, uniform_box_tape


@pytest.mark.parametrize("objective,eta,delta", CONFIGS)
@pytest.mark.parametrize("sparse,punish", [(True, True), (False, True), (False, False)])
def test_batch_c_matches_einsum_oracle(objective, eta, delta, sparse, punish):
  rng = np.random.default_rng(zlib.crc32(repr((objective, eta, sparse, punish)).encode()))
  N, T = 6, 40
  target = make_target(objective, eta, seed=11)
  batch = OracleBatch(N, eta, delta, objective, target, punish=punish, sparse=sparse, autoreset=True, final_obs=True)
  envs = [OracleCircuitDesigner(eta, delta, objective, punish=punish, sparse=sparse, target=target) for _ in range(N)]
  obs0 = [e.reset()[0] for e in envs]
  np.testing.assert_array_equal(batch.obs, np.stack(obs0))
  tape = uniform_box_tape(rng, eta, T, N, no_terminate=(rng.random() < 0.5))
  ep_ret = np.zeros(N); ep_len = np.zeros(N, int)
  n_done = 0
  for t in range(T):
    batch.step(tape[t])
    for i, env in enumerate(envs):
      obs, r, term, trunc, info = env.step(tape[t, i])
      ep_ret[i] += r; ep_len[i] += 1
      assert batch.status[i] == 0
      assert (bool(batch.terminated[i]), bool(batch.truncated[i])) == (term, trunc)
      assert (batch.depth[i], batch.operations[i], batch.used_wires[i]) == (info['depth'], info['operations'], info['used_wires'])
      np.testing.assert_allclose(batch.metric[i], info['metric'], atol=2e-15, rtol=0)
      assert batch.cost[i] == info['cost']
      np.testing.assert_allclose(batch.reward[i], r, atol=2e-15, rtol=0)
      if term or trunc:
        n_done += 1
        np.testing.assert_allclose(batch.final_obs[i], obs, atol=2e-7, rtol=0)
        np.testing.assert_allclose(batch.episode_return[i], ep_ret[i], atol=1e-14)
        assert batch.episode_length[i] == ep_len[i]
        ep_ret[i] = 0; ep_len[i] = 0
        obs, _ = env.reset()
      np.testing.assert_allclose(batch.obs[i], obs, atol=2e-7, rtol=0)
      np.testing.assert_allclose(batch.state[i], env.raw_state().reshape(-1), atol=5e-15, rtol=0)
  assert n_done > 0


def test_invalid_actions_leave_env_untouched():
  target = make_target('SP-ghz3', 3)
  b = OracleBatch(4, 3, 15, 'SP-ghz3', target)
  b.step(np.array([[1, 0, 0, 1.0]] * 4))
  before = b.state.copy(); meta = b.meta.copy()
  b.step(np.array([[0, 3, 0, 0], [0, 0, -0.5, 0], [3, 0, 0, 0], [float('nan'), 0, 0, 0]]))
  assert list(b.status) == [1, 1, 2, 2]
  np.testing.assert_array_equal(b.state, before); np.testing.assert_array_equal(b.meta, meta)
  assert not b.terminated.any() and not b.truncated.any() and (b.reward == 0).all()


def test_c_oracle_depth_255_boundary_matches_einsum_oracle():
  """uint8 depth counters of the batched layout: truncation at max_depth = 255 happens on the same
  step as in the (unbounded-counter) einsum oracle."""
  eta, delta = 1, 255
  target = make_target('SP-random', eta, seed=2)
  env = OracleCircuitDesigner(eta, delta, 'SP-random', target=target)
  cb = OracleBatch(1, eta, delta, 'SP-random', target, autoreset=False)
  env.reset()
  a = np.array([1.0, 0.0, 0.0, 0.37], np.float32)
  for t in range(255):
    obs, r, term, trunc, info = env.step(a)
    cb.step(a[None])
    assert bool(cb.truncated[0]) == trunc == (t == 254)
    assert cb.depth[0] == info['depth'] == t + 1
  np.testing.assert_allclose(cb.reward[0], r, atol=1e-13)
  assert info['cost'] == cb.cost[0] == (255 - 255 / 3) / (255 / 2 * 3)
