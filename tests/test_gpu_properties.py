"""Size-independent properties at BASELINE.json's full batch sizes, and adversarial action decoding.

At 65 536 envs the oracle only checks a slice; the whole batch is checked through invariants the domain
offers: unit norm of every statevector, unitarity of every composed operator, metric range, episode
accounting of the fused statistics, and replay == repeated steps."""
import numpy as np
import pytest
import torch

from oracle.batch_oracle import OracleBatch
from oracle.qcd_oracle import make_target
from tests.helpers import uniform_box_tape

pytestmark = pytest.mark.gpu


def _run(dev, objective, eta, delta, N, T, seed):
  from qcd_b200.batch import EnvBatch
  rng = np.random.default_rng(seed)
  target = make_target(objective, eta, seed=2)
  tape = uniform_box_tape(rng, eta, T, N)
  gpu = EnvBatch(N, eta, delta, objective, target, device=dev)
  M = 1024
  ora = OracleBatch(M, eta, delta, objective, target)
  dones = 0
  dtape = torch.as_tensor(tape, device=dev)
  for t in range(T):
    gpu.step(dtape[t]); ora.step(tape[t, :M])
    dones += int((gpu.terminated.bool() | gpu.truncated.bool()).sum())
    assert float(gpu.metric.min()) >= -1e-12 and float(gpu.metric.max()) <= 1 + 1e-12
  np.testing.assert_allclose(gpu.state[:M].cpu().numpy(), ora.state, atol=1e-12, rtol=0)
  np.testing.assert_allclose(gpu.reward[:M].cpu().numpy(), ora.reward, atol=1e-12, rtol=0)
  np.testing.assert_array_equal(gpu.depth[:M].cpu().numpy(), ora.depth)
  st = gpu.stats.cpu().numpy()
  assert st[0] == dones and st[10] == N * T and st[8] + st[9] == st[0] and st[11] == 0
  return gpu, dtape, target


def test_sp_ghz5_full_batch_invariants(cuda_device):
  from qcd_b200.batch import EnvBatch
  N, T = 65536, 40
  gpu, dtape, target = _run(cuda_device, 'SP-ghz5', 5, 20, N, T, seed=1)
  norms = (gpu.state.real ** 2 + gpu.state.imag ** 2).sum(1)
  assert float((norms - 1).abs().max()) < 1e-12                       # gates are unitary, resets are e0
  obs = gpu.obs[:, :64]
  assert torch.equal(obs, torch.cat([gpu.state.real, gpu.state.imag], 1).to(torch.float32))   # env.py:10
  assert torch.equal(gpu.obs[:, 64:], gpu.obs[:1, 64:].expand(N, 64))                         # constant target half
  rep = EnvBatch(N, 5, 20, 'SP-ghz5', target, device=cuda_device)
  rep.replay(dtape)
  assert torch.equal(rep.state, gpu.state) and torch.equal(rep.meta, gpu.meta) and torch.equal(rep.obs, gpu.obs)
  assert torch.allclose(rep.reward, gpu.reward, atol=1e-13, rtol=0)


def test_uc_toffoli_full_batch_unitarity(cuda_device):
  N, T = 65536, 45
  gpu, _, _ = _run(cuda_device, 'UC-toffoli', 3, 30, N, T, seed=3)
  V = gpu.state_view()
  eye = torch.eye(8, dtype=torch.complex128, device=cuda_device)
  err = (V.conj().transpose(1, 2) @ V - eye).abs().amax()
  assert float(err) < 1e-12


@pytest.mark.parametrize("objective,eta,delta", [('SP-random', 12, 36), ('UC-random', 6, 24)])
def test_large_state_invariants(cuda_device, objective, eta, delta):
  from qcd_b200.batch import EnvBatch
  N, T = 2048, 30
  rng = np.random.default_rng(4)
  target = make_target(objective, eta, seed=2)
  tape = torch.as_tensor(uniform_box_tape(rng, eta, T, N, no_terminate=True), device=cuda_device)
  a = EnvBatch(N, eta, delta, objective, target, device=cuda_device)
  b = EnvBatch(N, eta, delta, objective, target, device=cuda_device)
  for t in range(T): a.step(tape[t])
  b.replay(tape)
  assert torch.equal(a.state, b.state) and torch.equal(a.meta, b.meta)
  if 'SP' in objective:
    n = (a.state.real ** 2 + a.state.imag ** 2).sum(1)
    assert float((n - 1).abs().max()) < 1e-12
  else:
    V = a.state_view()[:64]
    err = (V.conj().transpose(1, 2) @ V - torch.eye(64, dtype=torch.complex128, device=cuda_device)).abs().amax()
    assert float(err) < 1e-12


def test_adversarial_action_decoding_is_bit_exact(cuda_device):
  """floor() decode on the edges: just below / at integers, signed zero, subnormals, huge, inf, nan."""
  from qcd_b200.batch import EnvBatch
  eta, delta = 5, 20
  f32 = np.float32
  edge = [0.0, -0.0, 0.99999994, 1.0, 1.0000001, 1.9999999, 2.0, 2.9999998, 2.99999, 3.0, 4.9999995, 4.99999, 5.0,
          -1e-45, 1e-45, -1.0, 1e9, -1e9, float('inf'), float('-inf'), float('nan'), 0.5, 3.5]
  grid = np.array([[g, w, c, th] for g in edge for w in edge for c in (0.0, 4.99999, 5.0, -0.0, float('nan'), 1.5)
                   for th in (0.3,)], dtype=f32)
  N = len(grid)
  target = make_target('SP-ghz5', eta)
  for dt in (np.float32, np.float64):
    gpu = EnvBatch(N, eta, delta, 'SP-ghz5', target, device=cuda_device)
    ora = OracleBatch(N, eta, delta, 'SP-ghz5', target)
    a = grid.astype(dt)
    gpu.step(torch.as_tensor(a, device=cuda_device)); ora.step(a)
    for name in ('status', 'terminated', 'truncated', 'depth', 'operations', 'used_wires', 'meta'):
      np.testing.assert_array_equal(getattr(gpu, name).cpu().numpy(), getattr(ora, name), err_msg=name)
    np.testing.assert_allclose(gpu.state.cpu().numpy(), ora.state, atol=1e-15, rtol=0)
    assert set(np.unique(ora.status)) == {0, 1, 2}
