"""Steps/s of the single-env gymnasium API (reference config: SP-bell, max_qubits=2, max_depth=12)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qcd_gym
from qcd_b200 import gym_compat as gym
from oracle.qcd_oracle import OracleCircuitDesigner, sample_actions
qcd_gym.register_envs()
for obj, q, d in (('SP-bell', 2, 12), ('UC-toffoli', 3, 30), ('SP-random', 10, 30)):
  env = gym.make("CircuitDesigner-v0", max_qubits=q, max_depth=d, objective=obj, seed=1)
  ora = OracleCircuitDesigner(q, d, obj, seed=1)
  rng = np.random.default_rng(0)
  acts = sample_actions(rng, q, 4000)
  for name, e in (('gpu', env), ('oracle-port', ora)):
    e.reset(); n = 0; t0 = time.perf_counter()
    for a in acts[:(4000 if name == 'gpu' or q < 8 else 300)]:
      _, _, term, trunc, _ = e.step(a); n += 1
      if term or trunc: e.reset()
    dt = time.perf_counter() - t0
    print(f"{obj} q{q}: {name} {n / dt:.0f} steps/s")
