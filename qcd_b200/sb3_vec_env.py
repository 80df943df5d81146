"""SB3 VecEnv over the batched engine.

The reference trains with `DummyVecEnv([Monitor(gym.make(...))] * n_envs)`
(/root/reference/algorithm/algorithm.py:16-27) and reads, per env, `info['episode']` (monitor.py:37-48),
`info['terminal_observation']`, `info['TimeLimit.truncated']`, `envs[i].unwrapped.name`
(algorithm.py:154) and `envs[i].termination_reasons` (algorithm.py:113); SB3 keeps the last 100
`info['episode']` records in `ep_info_buffer` (algorithm.py:49,89-91,100).  This class offers that surface
(numpy in / numpy out, list-of-dict infos).  It IS a `stable_baselines3.common.vec_env.VecEnv` when SB3 is
importable (so `BaseAlgorithm._wrap_env` takes it as it is); otherwise it is a plain class with the same methods.

One `qcd_step_sync` call per step: the action array, the observation rows, the terminal observations and every
scalar output live in pinned host memory mapped into the device address space, so `step_wait` issues no copy and
waits once.  Meant for the reference's small n_envs; large batches should use `CircuitDesignerVectorEnv` and keep
everything on the device (its `episode_window=100` is the device-side `ep_info_buffer`).
"""
from __future__ import annotations

import time
from collections import deque

import numpy as np
import torch

from . import _lib
from .batch import EnvBatch
from .gym_compat import Box, np_random
from .targets import make_target
from .vector_env import action_space_for

try:                                    # pragma: no cover - not available in this image
  from stable_baselines3.common.vec_env import VecEnv as _VecEnvBase
  HAVE_SB3 = True
except Exception:
  _VecEnvBase = object
  HAVE_SB3 = False


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
    return self._parent.name

  @property
  def target(self):
    return self._parent.target


class CircuitDesignerSB3VecEnv(_VecEnvBase):
  def __init__(self, num_envs: int, max_qubits: int, max_depth: int, objective: str, punish=True, sparse=True,
               seed=None, device=None, stats_window_size: int = 100):
    if seed is not None:
      _, seed = np_random(seed)
    self.target = make_target(objective, max_qubits, seed)       # same seed, same target for every env (algorithm.py:25-27)
    self.name = f"{objective}|{max_qubits}-{max_depth}"
    self.batch = b = EnvBatch(num_envs, max_qubits, max_depth, objective, self.target, punish=punish, sparse=sparse,
                              autoreset=True, final_obs=True, device=device, host_outputs=True)
    observation_space = Box(low=-1.0, high=+1.0, shape=(4 * b.S,))
    action_space = action_space_for(max_qubits)
    if HAVE_SB3:                        # pragma: no cover
      super().__init__(num_envs, observation_space, action_space)
    else:
      self.num_envs, self.observation_space, self.action_space = num_envs, observation_space, action_space
      self.render_mode = None
    self.envs = [_EnvView(self, i) for i in range(num_envs)]
    self.t_start = time.time()
    self.ep_info_buffer = deque(maxlen=stats_window_size)        # what SB3 keeps per algorithm (algorithm.py:49)
    self._stream = torch.cuda.Stream(b.device)
    self._act = torch.zeros((num_envs, 4), dtype=torch.float32, pin_memory=True)   # mapped: read by the kernel
    self._act_np = self._act.numpy()
    self._np = {k: getattr(b, k).numpy() for k in ('reward', 'metric', 'cost', 'depth', 'operations', 'used_wires',
                                                   'terminated', 'truncated', 'episode_return', 'episode_length')}
    self._obs_np, self._fin_np = b.obs.numpy(), b.final_obs.numpy()

  def reset(self):
    with torch.cuda.stream(self._stream):
      self.batch.reset()
    self._stream.synchronize()
    return self._obs_np.copy()

  def step_async(self, actions):
    self._act_np[...] = np.asarray(actions, dtype=np.float32).reshape(self.num_envs, 4)

  def step_wait(self):
    b, n = self.batch, self._np
    b.step_sync(self._act.data_ptr(), _lib.ACT_F32, self._stream.cuda_stream)      # launch + ONE wait, no copies
    term, trunc = n['terminated'] != 0, n['truncated'] != 0
    dones = term | trunc
    reward = n['reward'].astype(np.float32)
    infos = []
    now = round(time.time() - self.t_start, 6)
    for i in range(self.num_envs):
      d = {'depth': int(n['depth'][i]), 'operations': int(n['operations'][i]), 'used_wires': int(n['used_wires'][i]),
           'metric': float(n['metric'][i]), 'cost': float(n['cost'][i]),
           'TimeLimit.truncated': bool(trunc[i] and not term[i])}
      if dones[i]:
        self.envs[i].termination_reasons.append('DEPTH' if trunc[i] else 'DONE')       # env.py:130-131
        d['terminal_observation'] = self._fin_np[i].copy()
        d['episode'] = {'r': round(float(n['episode_return'][i]), 6), 'l': int(n['episode_length'][i]), 't': now,
                        'd': d['depth'], 'o': d['operations'], 'q': d['used_wires'], 'm': d['metric'], 'c': d['cost']}
        self.ep_info_buffer.append(d['episode'])
      infos.append(d)
    return self._obs_np.copy(), reward, dones, infos

  def step(self, actions):
    self.step_async(actions)
    return self.step_wait()

  def rolling_means(self) -> dict:
    """safe_mean over the last `stats_window_size` episodes of r,l,d,o,q,m,c (algorithm.py:89-91,100)."""
    if not self.ep_info_buffer:
      return {k: float('nan') for k in 'rldoqmc'}
    return {k: float(np.mean([ep[k] for ep in self.ep_info_buffer])) for k in 'rldoqmc'}

  # -- the rest of the VecEnv protocol SB3 touches
  def close(self): pass

  def seed(self, seed=None): return [seed] * self.num_envs

  def get_attr(self, attr_name, indices=None):
    return [getattr(self.envs[i], attr_name) for i in self._indices(indices)]

  def set_attr(self, attr_name, value, indices=None):
    for i in self._indices(indices): setattr(self.envs[i], attr_name, value)

  def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
    return [getattr(self.envs[i], method_name)(*method_args, **method_kwargs) for i in self._indices(indices)]

  def env_is_wrapped(self, wrapper_class, indices=None):
    return [False for _ in self._indices(indices)]

  def _indices(self, indices):
    if indices is None: return range(self.num_envs)
    if isinstance(indices, int): return [indices]
    return indices
