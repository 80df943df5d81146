"""CircuitDesigner — the reference's single-environment gymnasium API on the B200 batch engine.

Drop-in for /root/reference/qcd_gym/env.py:13-141: same constructor kwargs, attributes, spaces,
`reset` / `step` / `render` signatures, observation layout, info keys and exceptions.  The state
lives on the GPU as a batch of one env (autoreset off).  The action buffer, the observation row and every
per-step output live in pinned host memory that UVA maps into the device address space: the action
travels to the step kernel as a kernel argument and the kernel writes its results over PCIe itself, so a
`step` is ONE C call (`qcd_step1_sync`: launch + stream synchronisation) with no copy on either side.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np

from .gym_compat import Box, Env, np_random
from .targets import make_target
from .vector_env import action_space_for


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
    import torch
    from .batch import EnvBatch
    self._batch = b = EnvBatch(1, max_qubits, max_depth, objective, self.target, punish=punish, sparse=sparse,
                               autoreset=False, device=device, host_outputs=True, fused_stats=False)
    self._stream = torch.cuda.Stream(b.device)
    self._sptr = self._stream.cuda_stream
    self._act = (ctypes.c_double * 4)()                 # the action travels as a kernel argument
    self._torch = torch
    # the mapped outputs, grouped by dtype (valid after the synchronisation inside qcd_step1_sync):
    # [reward, metric, cost, ep_r], [depth, operations, used_wires, ep_l], [terminated, truncated, status]
    self._f, self._i, self._u = b.host_blocks
    self._obs_np = b.obs.numpy()[0]
    self._ops = []                         # host mirror of the circuit, used by render() only
    self.observation_space = Box(low=-1.0, high=+1.0, shape=(4 * b.S,))           # env.py:44
    self.action_space = action_space_for(self.qubits)                             # env.py:46-49

  # ------------------------------------------------------------------------------------------
  def _operation(self, action):
    """env.py:90-100: host-side replica of the decode, for the exceptions and for render()."""
    gate, wire, cntrl, theta = action
    try:                                   # np.floor([...]).astype(int) of the reference, on Python floats
      gate, wire, cntrl = math.floor(gate), math.floor(wire), math.floor(cntrl)
    except (ValueError, OverflowError):    # nan / inf: the reference's int cast gives INT_MIN -> AssertionError
      raise AssertionError(f"{action}") from None
    assert 0 <= wire < self.qubits and 0 <= cntrl < self.qubits, f"{action}"
    if wire == cntrl and gate == 0: return ('p', [wire], float(theta))
    if wire == cntrl and gate == 1: return ('rx', [wire], float(theta))
    if gate == 0: return ('cp', [cntrl, wire], float(theta))
    if gate == 1: return ('cx', [cntrl, wire], None)
    if gate == 2: return None
    assert False, f'Unhandled Action on gate {gate}'

  def reset(self, seed=None, options=None):
    super().reset(seed=seed)
    with self._torch.cuda.stream(self._stream):
      self._batch.reset()
    self._stream.synchronize()
    self._ops = []
    return self._obs_np.copy(), {'depth': 0, 'operations': 0, 'used_wires': 0}

  def step(self, action):
    operation = self._operation(action)              # raises AssertionError like env.py:94,100
    if operation is not None: self._ops.append(operation)
    # ONE C call: launch (the action is a kernel argument, the kernel writes the mapped results) + synchronise
    self._act[:] = action
    self._batch.step1_sync(self._act, self._sptr)
    reward, metric, cost, _ = self._f.tolist()
    depth, operations, used_wires, _ = self._i.tolist()
    terminated, truncated, _ = self._u.tolist()
    terminated, truncated = terminated != 0, truncated != 0
    info = {'depth': depth, 'operations': operations, 'used_wires': used_wires}
    if terminated: info['termination_reason'] = 'DONE'
    if truncated: info['termination_reason'] = 'DEPTH'                            # env.py:130-131
    info['metric'] = metric; info['cost'] = cost
    return self._obs_np.copy(), reward, terminated, truncated, info

  # ------------------------------------------------------------------------------------------
  def render(self):
    """Drawing of the circuit built so far.  The reference delegates to qiskit's drawer (`qc.draw(mode)`,
    env.py:139-141: 'text' -> an ASCII drawing, 'image' -> a matplotlib figure); qiskit is not available here, so
    'text' is drawn by `_draw_text` and 'image' is an RGB array (H, W, 3) uint8 drawn with PIL (`_draw_image`)."""
    if self.render_mode is None: return None
    if self.render_mode == 'text': return self._draw_text()
    if self.render_mode == 'image': return self._draw_image()
    raise NotImplementedError(f"render_mode={self.render_mode!r}")

  @staticmethod
  def _label(kind, theta):
    return {'p': 'P', 'rx': 'Rx', 'cp': 'P', 'cx': 'X'}[kind] + ('' if theta is None else f"({theta:.3g})")

  def _draw_text(self) -> str:
    rows = [[f"q_{q}: "] for q in range(self.qubits)]
    for kind, qargs, theta in self._ops:
      label = self._label(kind, theta)
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

  def _draw_image(self) -> np.ndarray:
    """One column per gate, one horizontal wire per qubit; boxes on targets, dots on controls."""
    from PIL import Image, ImageDraw
    col, row, left, box = 64, 40, 48, (52, 24)
    W, H = left + col * max(1, len(self._ops)) + 16, row * self.qubits + 16
    img = Image.new('RGB', (W, H), 'white')
    d = ImageDraw.Draw(img)
    y = lambda q: 8 + row * q + row // 2
    for q in range(self.qubits):
      d.text((4, y(q) - 6), f"q{q}", fill='black')
      d.line([(left - 16, y(q)), (W - 8, y(q))], fill='black', width=1)
    for i, (kind, qargs, theta) in enumerate(self._ops):
      x = left + col * i + col // 2
      tgt = qargs[-1]
      if len(qargs) == 2:
        d.line([(x, y(qargs[0])), (x, y(tgt))], fill=(60, 60, 200), width=2)
        d.ellipse([x - 4, y(qargs[0]) - 4, x + 4, y(qargs[0]) + 4], fill=(60, 60, 200))
      d.rectangle([x - box[0] // 2, y(tgt) - box[1] // 2, x + box[0] // 2, y(tgt) + box[1] // 2],
                  fill=(225, 235, 255), outline=(60, 60, 200))
      d.text((x - box[0] // 2 + 3, y(tgt) - 6), self._label(kind, theta)[:9], fill='black')
    return np.asarray(img, dtype=np.uint8)

  @property
  def state(self) -> np.ndarray:
    """complex128 state vector / unitary currently resident on the device."""
    return self._batch.state_view()[0].cpu().numpy()
