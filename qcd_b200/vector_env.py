"""CircuitDesignerVectorEnv — N reference environments stepped by one kernel launch.

Replaces the SB3 `DummyVecEnv([Monitor(gym.make(...))] * n)` loop of
/root/reference/algorithm/algorithm.py:25-27 (sequential Python, one qiskit re-simulation per env
per step) behind the same constructor kwargs as `CircuitDesigner` (env.py:29-30).

Semantics: gymnasium-0.29 VectorEnv same-step auto-reset.  When env i finishes, `reward[i]`,
`terminated[i]`, `truncated[i]` and the info tensors describe the finished step, `obs[i]` is already
the first observation of the next episode, `info['final_observation']` (if enabled) holds the
terminal one, and `info['episode']` carries Monitor's r/l (monitor.py:37-38); d/o/q/m/c of
monitor.py:44-48 are the info tensors depth/operations/used_wires/metric/cost at the same index.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .batch import EnvBatch
from .gym_compat import Box, VectorEnvBase, np_random
from .targets import make_target

GATES = 3


def action_space_for(qubits: int) -> Box:
  """env.py:46-49 — Box([0,0,0,-pi], [3-m, eta-m, eta-m, pi]) in float32."""
  m = 1e-5
  return Box(np.array([0, 0, 0, -np.pi]), np.array([GATES - m, qubits - m, qubits - m, np.pi]))


class CircuitDesignerVectorEnv(VectorEnvBase):
  """A `gymnasium.vector.VectorEnv` (when gymnasium is importable) whose envs live in HBM; it is the
  `vector_entry_point` of `CircuitDesigner-v0` (registration.py).  Observations, rewards and flags are torch
  CUDA tensors; `step_host*` are the numpy-in / numpy-out variants."""
  metadata = {"render_modes": [], "autoreset_mode": "same_step"}
  spec = None
  render_mode = None
  closed = False

  def __init__(self, num_envs: int, max_qubits: int, max_depth: int, objective: str,
               punish=True, sparse=True, seed=None, target=None, device=None,
               obs_mode: str = 'full', final_obs: bool = False, autoreset: bool = True,
               raise_on_invalid: bool = False, episode_window: int = 0, render_mode=None,
               zero_copy_scalars: bool = False):
    self.num_envs = int(num_envs)
    # host path (step_host*): True = the step kernel writes the requested scalar results straight into the pinned host
    # arrays (see _host_desc) and only the observation rows are copied; the DEVICE tensors of those scalars
    # (batch.reward, .terminated, .truncated, ...) are then not outputs of such a step, so leave it off when device
    # code (accumulate_stats, an episode window) consumes them after a host step
    self.zero_copy_scalars = bool(zero_copy_scalars)
    self.qubits, self.depth = max_qubits, max_depth
    self.max_steps = max_depth * max_qubits * 2
    self.punish, self.sparse, self.objective = punish, sparse, objective
    self.name = f"{objective}|{max_qubits}-{max_depth}"
    self.raise_on_invalid = raise_on_invalid
    self._np_random = None
    if seed is not None:
      self._np_random, seed = np_random(seed)
    # env.py:40 — one target for the lifetime of the env; DummyVecEnv gives every env the same seed
    # and therefore the same target (algorithm.py:25-27), which is the `shared` layout here.
    self.target = make_target(objective, max_qubits, seed) if target is None else target
    self.batch = EnvBatch(self.num_envs, max_qubits, max_depth, objective, self.target, punish=punish,
                          sparse=sparse, autoreset=autoreset, obs_mode=obs_mode, final_obs=final_obs,
                          device=device)
    self.device = self.batch.device
    S = self.batch.S
    self.single_observation_space = Box(low=-1.0, high=+1.0, shape=(4 * S,))
    self.single_action_space = action_space_for(max_qubits)
    self.observation_space = Box(low=-1.0, high=+1.0, shape=(self.num_envs, self.batch.obs_stride))
    self.action_space = Box(np.tile(self.single_action_space.low, (self.num_envs, 1)),
                            np.tile(self.single_action_space.high, (self.num_envs, 1)))
    self._host = None
    self._groups = None
    # device-side rolling window of the last `episode_window` episodes (SB3 ep_info_buffer; algorithm.py:49,100)
    self.window = None
    if episode_window:
      from .batch import EpisodeWindow
      self.window = EpisodeWindow(self.batch, episode_window)

  # ------------------------------------------------------------------------------------------
  def _info(self) -> dict:
    b = self.batch
    info = {'depth': b.depth, 'operations': b.operations, 'used_wires': b.used_wires,
            'metric': b.metric, 'cost': b.cost, 'status': b.status}
    if b.episode_return is not None:
      info['episode'] = {'r': b.episode_return, 'l': b.episode_length}
    if b.final_obs is not None:
      info['final_observation'] = b.final_obs
    return info

  def reset(self, *, seed=None, options=None):
    if seed is not None:
      self._np_random, _ = np_random(seed)
    self.batch.reset()
    b = self.batch
    return b.obs, {'depth': b.depth, 'operations': b.operations, 'used_wires': b.used_wires}

  def step(self, actions: torch.Tensor):
    """actions: float32/float64 [N,4] tensor on the env's device.  Returns persistent device
    tensors (no allocation, no synchronisation): obs, reward f64, terminated, truncated, info."""
    b = self.batch
    b.step(actions)
    if self.window is not None:
      self.window.push()
    if self.raise_on_invalid and bool((b.status != 0).any()):
      bad = int(torch.nonzero(b.status)[0])
      raise AssertionError(f"env {bad}: invalid action {actions[bad].tolist()} (status {int(b.status[bad])})")
    return b.obs, b.reward, b.terminated.view(torch.bool), b.truncated.view(torch.bool), self._info()

  def replay(self, tape: torch.Tensor, record: dict | None = None):
    """Apply a whole action tape [T,N,4] in one launch (states stay in shared memory)."""
    b = self.batch
    b.replay(tape, record)
    return b.obs, b.reward, b.terminated.view(torch.bool), b.truncated.view(torch.bool), self._info()

  # ------------------------------------------------------------------------------------------
  def _host_buffers(self, action_dtype):
    if self._host is None or self._host['actions'].dtype != action_dtype:
      N, b = self.num_envs, self.batch
      pin = dict(pin_memory=True)
      host = {
          'actions': torch.zeros((N, 4), dtype=action_dtype, **pin),
          'actions_dev': torch.zeros((N, 4), dtype=action_dtype, device=self.device),
          'reward': torch.zeros(N, dtype=torch.float64, **pin),
          'metric': torch.zeros(N, dtype=torch.float64, **pin),
          'cost': torch.zeros(N, dtype=torch.float64, **pin),
          'depth': torch.zeros(N, dtype=torch.int32, **pin),
          'operations': torch.zeros(N, dtype=torch.int32, **pin),
          'used_wires': torch.zeros(N, dtype=torch.int32, **pin),
          'terminated': torch.zeros(N, dtype=torch.uint8, **pin),
          'truncated': torch.zeros(N, dtype=torch.uint8, **pin),
          'status': torch.zeros(N, dtype=torch.int8, **pin),
          'obs': None,
      }
      if b.obs is not None:
        # same row width as on the device: 'full' rows carry the constant target half (pitched copy of the
        # state half), 'state' rows are compact and move in ONE contiguous transfer
        host['obs'] = torch.zeros((N, b.obs_stride), dtype=torch.float32, **pin)
        if b.obs_mode == 'full':
          host['obs'][:, 2 * b.S:] = b.obs[:, 2 * b.S:].cpu()      # per row: targets may differ per env
      host['np'] = {k: host[k].numpy() for k in ('actions', 'reward', 'metric', 'cost', 'depth', 'operations',
                                                  'used_wires', 'terminated', 'truncated', 'status')}
      host['np']['terminated'] = host['np']['terminated'].view(np.bool_)
      host['np']['truncated'] = host['np']['truncated'].view(np.bool_)
      host['np']['obs'] = None if host['obs'] is None else host['obs'].numpy()
      self._host = host
      self._groups = None
    return self._host

  _SCALARS = ('reward', 'metric', 'cost', 'depth', 'operations', 'used_wires', 'terminated', 'truncated', 'status')

  def _host_desc(self, h, lo, hi, copy_info):
    """`qcd_batch_t` of the env range [lo, hi) for the HOST path: the requested per-step scalar outputs point at
    the rows [lo, hi) of the pinned host arrays (UVA maps them into the device address space), so the step kernel
    writes them over PCIe itself and only the observation rows need the copy engine (one transfer instead of
    four to ten per group and step).  The device-side tensors of those scalars are NOT written by such a step."""
    key = ('desc', lo, hi, bool(copy_info))
    if key not in h:
      c = self.batch.sub_batch(lo, hi)
      if self.zero_copy_scalars:
        for k in (self._SCALARS if copy_info else ('reward', 'terminated', 'truncated')):
          setattr(c, k, h[k].data_ptr() + lo * h[k].element_size())
      h[key] = c
    return h[key]

  def _host_out(self, h, lo, copy_obs, copy_info):
    """qcd_host_out_t whose destinations are the rows [lo, ...) of the pinned host arrays."""
    b = self.batch
    names = () if self.zero_copy_scalars else (self._SCALARS if copy_info else ('reward', 'terminated', 'truncated'))
    out = _lib.QcdHostOut(obs_stride=b.obs_stride)
    for k in names:
      setattr(out, k, h[k].data_ptr() + lo * h[k].element_size())
    if copy_obs and h['obs'] is not None:
      out.obs = h['obs'].data_ptr() + lo * b.obs_stride * 4
    return out

  def step_host(self, actions, copy_obs: bool = True, copy_info: bool = False):
    """The end-to-end call with HOST buffers: numpy actions in, numpy results out.

    H2D copy of the actions, the step kernel and the D2H copies of reward / terminated / truncated
    (and the state half of every observation row; with copy_info also metric, cost, depth,
    operations, used_wires, status) are issued on the current stream by `qcd_step_host`, followed by
    one stream synchronisation.  Returned arrays are views of pinned host buffers that the next call
    overwrites.  With obs_mode='state' the observation rows are the compact state halves [N, 2*S]
    (one contiguous transfer; `target_obs` is the constant other half of env.py:85), with 'full' they are
    the reference rows [N, 4*S].  Returns (obs, reward, terminated, truncated[, info])."""
    a = np.asarray(actions)
    dt = torch.float64 if a.dtype == np.float64 else torch.float32
    h = self._host_buffers(dt)
    h['np']['actions'][...] = a.reshape(self.num_envs, 4)
    import ctypes as C
    b = self.batch
    out = self._host_out(h, 0, copy_obs, copy_info)
    desc = self._host_desc(h, 0, self.num_envs, copy_info)
    stream = torch.cuda.current_stream(self.device)
    with torch.cuda.device(self.device):
      _lib.check(b.lib.qcd_step_host(
          C.byref(desc), h['actions'].data_ptr(), h['actions_dev'].data_ptr(),
          _lib.ACT_F64 if dt == torch.float64 else _lib.ACT_F32, out, stream.cuda_stream), "qcd_step_host")
    stream.synchronize()
    n = h['np']
    res = (n['obs'] if (copy_obs and n['obs'] is not None) else None, n['reward'], n['terminated'], n['truncated'])
    if copy_info:
      res += ({k: n[k] for k in ('metric', 'cost', 'depth', 'operations', 'used_wires', 'status')},)
    return res

  # ------------------------------------------------------------------------------------------
  # Pipelined host path: the envs are split into `groups` contiguous groups, each stepped on its own stream.
  # While group g's results travel to the host (D2H engine), group g+1's actions go up (H2D engine) and its
  # kernel runs; the caller only ever supplies actions for a group AFTER it has received that group's previous
  # results, so the RL data dependence (action_{t+1} depends on obs_t) is respected per group.
  def host_groups(self, groups: int = 2, action_dtype=np.float32):
    """[(lo, hi)] env ranges of the groups (creates streams / descriptors on first use)."""
    dt = torch.float64 if np.dtype(action_dtype) == np.float64 else torch.float32
    h = self._host_buffers(dt)
    if self._groups is None or len(self._groups) != groups:
      from .sharding import shard_range
      self._groups = []
      for g in range(groups):
        lo, hi = shard_range(self.num_envs, g, groups)
        self._groups.append({'lo': lo, 'hi': hi,
                             'stream': torch.cuda.Stream(self.device), 'event': torch.cuda.Event(),
                             'pending': False, 'args': None})
    return [(g['lo'], g['hi']) for g in self._groups]

  def step_host_async(self, group: int, actions, copy_obs: bool = True, copy_info: bool = False) -> None:
    """Submit one step of env group `group` (actions: [hi-lo, 4] host array) without waiting."""
    import ctypes as C
    a = np.asarray(actions)
    dt = torch.float64 if a.dtype == np.float64 else torch.float32
    h = self._host_buffers(dt)
    if self._groups is None:
      self.host_groups(2, a.dtype)
    g = self._groups[group]
    if g['pending']:
      raise RuntimeError(f"group {group}: step_host_wait must collect the previous step first")
    lo, hi = g['lo'], g['hi']
    h['np']['actions'][lo:hi] = a.reshape(hi - lo, 4)
    out = self._host_out(h, lo, copy_obs, copy_info)
    esz = h['actions'].element_size() * 4
    with torch.cuda.device(self.device):
      _lib.check(self.batch.lib.qcd_step_host(
          C.byref(self._host_desc(h, lo, hi, copy_info)), h['actions'].data_ptr() + lo * esz,
          h['actions_dev'].data_ptr() + lo * esz,
          _lib.ACT_F64 if dt == torch.float64 else _lib.ACT_F32, out, g['stream'].cuda_stream), "qcd_step_host")
    g['event'].record(g['stream'])
    g['pending'], g['args'] = True, (copy_obs, copy_info)

  def step_host_wait(self, group: int):
    """Wait for the submitted step of `group`; returns views of rows [lo, hi) of the pinned host arrays:
    (obs, reward, terminated, truncated[, info])."""
    g = self._groups[group]
    if not g['pending']:
      raise RuntimeError(f"group {group}: nothing submitted")
    g['event'].synchronize()
    g['pending'] = False
    copy_obs, copy_info = g['args']
    n, lo, hi = self._host['np'], g['lo'], g['hi']
    res = (n['obs'][lo:hi] if (copy_obs and n['obs'] is not None) else None, n['reward'][lo:hi],
           n['terminated'][lo:hi], n['truncated'][lo:hi])
    if copy_info:
      res += ({k: n[k][lo:hi] for k in ('metric', 'cost', 'depth', 'operations', 'used_wires', 'status')},)
    return res

  @property
  def target_obs(self) -> np.ndarray:
    """float32 [2*S] (or [N, 2*S] with per-env targets): flat(target), the constant half of env.py:85."""
    t = self.batch.target
    return torch.cat([t.real, t.imag], dim=-1).to(torch.float32).cpu().numpy()

  def host_bytes_per_step(self, action_dtype=np.float32, copy_obs: bool = True):
    """(h2d, d2h) bytes `step_host` moves per call."""
    N, b = self.num_envs, self.batch
    h2d = N * 4 * np.dtype(action_dtype).itemsize
    d2h = N * (8 + 1 + 1) + (N * 2 * b.S * 4 if (copy_obs and b.obs is not None) else 0)
    return h2d, d2h

  # ------------------------------------------------------------------------------------------
  def accumulate_stats(self) -> None:
    """Fold the last step's finished episodes into the running statistics vector (device)."""
    self.batch.reduce_stats()

  def episode_stats(self, reset: bool = False) -> dict:
    """Running sums as a dict (synchronises)."""
    vec = self.batch.stats.cpu().numpy().copy()
    if reset:
      self.batch.zero_stats()
    return dict(zip(_lib.STAT_NAMES, vec.tolist()))

  def rolling_means(self) -> dict:
    """Means of r,l,d,o,q,m,c over the last `episode_window` finished episodes (the reference's `return-100`,
    `length-100` and progress-bar numbers: algorithm.py:49,89-91,100).  Synchronises."""
    if self.window is None:
      raise RuntimeError("construct the env with episode_window=100")
    return self.window.means()

  def close(self, **kwargs):
    self.closed = True

  def close_extras(self, **kwargs):
    pass
