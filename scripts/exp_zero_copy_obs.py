"""Experiment: ALL per-step outputs (observation rows included) written by the step kernel into mapped pinned host
memory (EnvBatch(host_outputs=True)): no copy-engine transfer at all.  Compared with the e2e path of bench.py."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from qcd_b200.batch import EnvBatch
from qcd_b200.targets import make_target
dev = torch.device('cuda', 0)
N, eta, delta = 65536, 5, 20
rng = np.random.default_rng(0)
low = np.array([0, 0, 0, -np.pi]); high = np.array([3 - 1e-5, eta - 1e-5, eta - 1e-5, np.pi])
tape = torch.as_tensor(rng.uniform(low, high, size=(16, N, 4)).astype(np.float32)).pin_memory()
for groups in (1, 2, 4):
  n = N // groups
  bs = [EnvBatch(n, eta, delta, 'SP-ghz5', make_target('SP-ghz5', eta), obs_mode='state', host_outputs=True, device=dev)
        for _ in range(groups)]
  streams = [torch.cuda.Stream(dev) for _ in range(groups)]
  acts = [torch.zeros((n, 4), dtype=torch.float32, device=dev) for _ in range(groups)]
  for b in bs: b.reset()
  torch.cuda.synchronize()
  def loop(steps):
    for i in range(steps):
      for g, b in enumerate(bs):
        streams[g].synchronize()                      # previous results of this group are on the host
        with torch.cuda.stream(streams[g]):
          acts[g].copy_(tape[i % 16, g * n:(g + 1) * n], non_blocking=True)
          b.step(acts[g])
    for s in streams: s.synchronize()
  loop(3)
  t0 = time.perf_counter(); K = 20; loop(K); dt = time.perf_counter() - t0
  byts = N * (2 * 32 * 4 + 10)
  print(f"groups={groups}: {N * K / dt / 1e6:.1f} M env-steps/s, {dt / K * 1e3:.4f} ms/step, {byts * K / dt / 1e9:.1f} GB/s over PCIe (kernel stores)")
