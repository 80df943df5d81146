"""Single-env steps/s through gym.make (BASELINE.json configs[0]) next to the numpy oracle port."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import qcd_gym
from qcd_b200 import gym_compat as gym
from oracle.qcd_oracle import OracleCircuitDesigner, sample_actions

qcd_gym.register_envs()
for objective, eta, delta in (('SP-bell', 2, 12), ('UC-toffoli', 3, 30), ('SP-ghz5', 5, 20)):
  rng = np.random.default_rng(0)
  acts = [sample_actions(rng, eta) for _ in range(4096)]
  for name, env in (('gpu', gym.make("CircuitDesigner-v0", max_qubits=eta, max_depth=delta, objective=objective)),
                    ('oracle', OracleCircuitDesigner(eta, delta, objective))):
    env.reset()
    for warm in (True, False):
      n = 2000 if warm else 20000
      t0 = time.perf_counter()
      for i in range(n):
        _, _, term, trunc, _ = env.step(acts[i & 4095])
        if term or trunc: env.reset()
      dt = time.perf_counter() - t0
    print(f"{objective} {name}: {n / dt:,.0f} steps/s ({dt / n * 1e6:.1f} us/step)", flush=True)

# breakdown for SP-bell: the C call alone (launch + kernel + sync) vs the whole Python step
import ctypes
env = gym.make("CircuitDesigner-v0", max_qubits=2, max_depth=12, objective='SP-bell')
env.reset()
u = env.unwrapped
act = (ctypes.c_double * 4)(1.0, 0.0, 0.0, 0.5)
for n in (2000, 20000):
  t0 = time.perf_counter()
  for i in range(n):
    u._batch.step1_sync(act, u._sptr)
  dt = time.perf_counter() - t0
print(f"SP-bell qcd_step1_sync alone: {dt / n * 1e6:.1f} us/call", flush=True)
