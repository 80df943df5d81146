"""Smallest case that runs every kernel family once (for compute-sanitizer)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from qcd_b200.batch import EnvBatch
from qcd_b200.targets import make_target
dev = torch.device('cuda', 0)
rng = np.random.default_rng(0)
for objective, eta, delta, N, T in (('SP-ghz5', 5, 20, 70, 6), ('UC-toffoli', 3, 30, 40, 6), ('SP-random', 10, 30, 3, 12),
                                    ('UC-random', 5, 20, 3, 12), ('SP-random', 12, 36, 2, 12), ('UC-random', 6, 24, 2, 12)):
  low = np.array([0, 0, 0, -np.pi]); high = np.array([3 - 1e-5, eta - 1e-5, eta - 1e-5, np.pi])
  tape = torch.as_tensor(rng.uniform(low, high, size=(T, N, 4)).astype(np.float32), device=dev)
  a = EnvBatch(N, eta, delta, objective, make_target(objective, eta, seed=1), final_obs=True, device=dev)
  b = EnvBatch(N, eta, delta, objective, make_target(objective, eta, seed=1), final_obs=True, device=dev)
  for t in range(T): a.step(tape[t])
  b.replay(tape)
  out = a.fold_stats(zero=True)
  torch.cuda.synchronize()
  assert torch.equal(a.state, b.state), objective
  print(objective, eta, 'ok', int(out[10].item()))
