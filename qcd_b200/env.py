"""CircuitDesigner — the reference's single-environment gymnasium API on the B200 batch engine.

Drop-in for /root/reference/qcd_gym/env.py:13-141: same constructor kwargs, attributes, spaces,
`reset` / `step` / `render` signatures, observation layout, info keys and exceptions.  The state
lives on the GPU as a vector env of one env (autoreset off); every `step` is one `qcd_step_host`
call (H2D action, step kernel, D2H of the observation and the scalars) and one synchronisation.
"""
from __future__ import annotations

import numpy as np

from .gym_compat import Box, Env, np_random
from .targets import make_target
from .vector_env import GATES, CircuitDesignerVectorEnv, action_space_for


class CircuitDesigner(Env):
  """ Quantum Circuit Environment: build a quantum circuit gate-by-gate for a desired objective.

  Attributes (env.py:16-20): qubits, depth, objective, punish; plus max_steps, sparse, target, name.
  """

  metadata = {"render_modes": ["image", "text"], "render_fps": 30}

  def __init__(self, max_qubits: int, max_depth: int, objective: str,
               punish=True, sparse=True, seed=None, render_mode=None, device=None):
    super().__init__()
    if seed is not None: self._np_random, seed = np_random(seed)                  # env.py:32
    self.render_mode = render_mode; self.name = f"{objective}|{max_qubits}-{max_depth}"
    self.qubits, self.depth = max_qubits, max_depth
    self.max_steps = max_depth * max_qubits * 2                                   # env.py:37
    self.punish = punish; self.sparse = sparse
    self.objective = objective
    self.target = make_target(objective, max_qubits, seed)                        # env.py:40
    self._venv = CircuitDesignerVectorEnv(1, max_qubits, max_depth, objective, punish=punish, sparse=sparse,
                                          target=self.target, device=device, autoreset=False)
    self._batch = self._venv.batch
    self._ops = []                         # host mirror of the circuit, used by render() only
    b = self._batch
    self.observation_space = Box(low=-1.0, high=+1.0, shape=(4 * b.S,))           # env.py:44
    self.action_space = action_space_for(self.qubits)                             # env.py:46-49

  # ------------------------------------------------------------------------------------------
  def _operation(self, action):
    """env.py:90-100: host-side replica of the decode, for the exceptions and for render()."""
    gate, wire, cntrl, theta = action
    gate, wire, cntrl = np.floor([gate, wire, cntrl]).astype(int)
    assert wire in range(self.qubits) and cntrl in range(self.qubits), f"{action}"
    if wire == cntrl and gate == 0: return ('p', [wire], float(theta))
    if wire == cntrl and gate == 1: return ('rx', [wire], float(theta))
    if gate == 0: return ('cp', [cntrl, wire], float(theta))
    if gate == 1: return ('cx', [cntrl, wire], None)
    if gate == 2: return None
    assert False, f'Unhandled Action on gate {gate}'

  def reset(self, seed=None, options=None):
    super().reset(seed=seed)
    obs, _ = self._venv.reset()
    self._ops = []
    return obs[0].cpu().numpy(), {'depth': 0, 'operations': 0, 'used_wires': 0}

  def step(self, action):
    operation = self._operation(action)              # raises AssertionError like env.py:94,100
    if operation is not None: self._ops.append(operation)
    # one C-ABI call (H2D action, step kernel, D2H of the observation and every scalar) + one sync
    obs, reward, term, trunc, inf = self._venv.step_host(np.asarray(action, dtype=np.float64).reshape(1, 4),
                                                        copy_info=True)
    terminated, truncated = bool(term[0]), bool(trunc[0])
    info = {'depth': int(inf['depth'][0]), 'operations': int(inf['operations'][0]),
            'used_wires': int(inf['used_wires'][0])}
    if terminated: info['termination_reason'] = 'DONE'
    if truncated: info['termination_reason'] = 'DEPTH'                            # env.py:130-131
    info = {**info, 'metric': float(inf['metric'][0]), 'cost': float(inf['cost'][0])}
    return obs[0].copy(), float(reward[0]), terminated, truncated, info

  # ------------------------------------------------------------------------------------------
  def render(self):
    """Text drawing of the circuit built so far (the reference delegates to qiskit's drawer,
    env.py:139-141; 'image' needs matplotlib+qiskit and is not available here)."""
    if self.render_mode is None: return None
    if self.render_mode != 'text':
      raise NotImplementedError("render_mode='image' needs qiskit's matplotlib drawer")
    rows = [[f"q_{q}: "] for q in range(self.qubits)]
    for kind, qargs, theta in self._ops:
      label = {'p': 'P', 'rx': 'Rx', 'cp': 'P', 'cx': 'X'}[kind] + ('' if theta is None else f"({theta:.3g})")
      width = len(label) + 2
      tgt = qargs[-1]
      ctl = qargs[0] if len(qargs) == 2 else None
      lo, hi = (min(qargs), max(qargs))
      for q in range(self.qubits):
        if q == tgt: cell = label.center(width, '─')
        elif q == ctl: cell = '■'.center(width, '─')
        elif lo < q < hi: cell = '│'.center(width, '─')
        else: cell = '─' * width
        rows[q].append(cell)
    return '\n'.join(''.join(r) for r in rows)

  @property
  def state(self) -> np.ndarray:
    """complex128 state vector / unitary currently resident on the device."""
    return self._batch.state_view()[0].cpu().numpy()
