"""SB3-style VecEnv facade over the batched engine.

The reference trains with `DummyVecEnv([Monitor(gym.make(...))] * n_envs)`
(/root/reference/algorithm/algorithm.py:16-27) and reads, per env, `info['episode']` (monitor.py:37-48),
`info['terminal_observation']`, `info['TimeLimit.truncated']`, `envs[i].unwrapped.name`
(algorithm.py:154) and `envs[i].termination_reasons` (algorithm.py:113).  This class offers that
surface (numpy in / numpy out, list-of-dict infos) without importing stable-baselines3, backed by one
`qcd_step` launch per step.  Meant for the reference's small n_envs; large batches should use
`CircuitDesignerVectorEnv` directly and keep everything on the device.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from .vector_env import CircuitDesignerVectorEnv


class _EnvView:
  """What `vec.envs[i]` must offer to the reference's trainer."""

  def __init__(self, parent, index):
    self._parent, self._index = parent, index
    self.termination_reasons = []

  @property
  def unwrapped(self):
    return self

  @property
  def name(self):
    return self._parent.venv.name

  @property
  def target(self):
    return self._parent.venv.target


class CircuitDesignerSB3VecEnv:
  def __init__(self, num_envs: int, max_qubits: int, max_depth: int, objective: str, punish=True, sparse=True,
               seed=None, device=None):
    self.venv = CircuitDesignerVectorEnv(num_envs, max_qubits, max_depth, objective, punish=punish, sparse=sparse,
                                         seed=seed, device=device, final_obs=True)
    self.num_envs = num_envs
    self.observation_space = self.venv.single_observation_space
    self.action_space = self.venv.single_action_space
    self.envs = [_EnvView(self, i) for i in range(num_envs)]
    self.t_start = time.time()
    self._actions = None
    self.render_mode = None

  def reset(self):
    obs, _ = self.venv.reset()
    return obs.cpu().numpy()

  def step_async(self, actions):
    self._actions = torch.as_tensor(np.asarray(actions, dtype=np.float32), device=self.venv.device)

  def step_wait(self):
    obs, reward, term, trunc, info = self.venv.step(self._actions)
    obs = obs.cpu().numpy(); reward = reward.cpu().numpy().astype(np.float32)
    term = term.cpu().numpy(); trunc = trunc.cpu().numpy()
    dones = term | trunc
    host = {k: info[k].cpu().numpy() for k in ('depth', 'operations', 'used_wires', 'metric', 'cost')}
    infos = []
    fin = info['final_observation'].cpu().numpy() if dones.any() else None
    ep_r = info['episode']['r'].cpu().numpy(); ep_l = info['episode']['l'].cpu().numpy()
    for i in range(self.num_envs):
      d = {k: (int(v[i]) if v.dtype.kind == 'i' else float(v[i])) for k, v in host.items()}
      d['TimeLimit.truncated'] = bool(trunc[i] and not term[i])
      if dones[i]:
        self.envs[i].termination_reasons.append('DEPTH' if trunc[i] else 'DONE')       # env.py:130-131
        d['terminal_observation'] = fin[i]
        d['episode'] = {'r': round(float(ep_r[i]), 6), 'l': int(ep_l[i]), 't': round(time.time() - self.t_start, 6),
                        'd': d['depth'], 'o': d['operations'], 'q': d['used_wires'], 'm': d['metric'], 'c': d['cost']}
      infos.append(d)
    return obs, reward, dones, infos

  def step(self, actions):
    self.step_async(actions)
    return self.step_wait()

  # -- the rest of the VecEnv protocol SB3 touches
  def close(self): pass

  def seed(self, seed=None): return [seed] * self.num_envs

  def get_attr(self, name, indices=None):
    idx = range(self.num_envs) if indices is None else indices
    return [getattr(self.envs[i], name) for i in idx]

  def set_attr(self, name, value, indices=None):
    for i in (range(self.num_envs) if indices is None else indices): setattr(self.envs[i], name, value)

  def env_method(self, name, *args, indices=None, **kwargs):
    return [getattr(self.envs[i], name)(*args, **kwargs) for i in (range(self.num_envs) if indices is None else indices)]

  def env_is_wrapped(self, wrapper_class, indices=None):
    return [False] * self.num_envs
