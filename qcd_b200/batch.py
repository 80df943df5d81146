"""EnvBatch: N CircuitDesigner environments resident in HBM + the launches of libqcd_b200.so.

Owns every device buffer named in `qcd_batch_t` (include/qcd_b200.h) as a torch tensor, so the
kernels never allocate; torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import QcdBatch, QcdTapeOut, check


def objective_kind(objective: str) -> int:
  """'SP-…' -> OBJ_SP, 'UC-…' -> OBJ_UC (the reference tests `'UC' in objective`, env.py:82-84)."""
  if 'UC' in objective:
    return _lib.OBJ_UC
  if 'SP' in objective:
    return _lib.OBJ_SP
  raise AssertionError(f'{objective} not defined.')


def _ptr(t):
  return None if t is None else t.data_ptr()


class EnvBatch:
  """Device-resident state of `num_envs` environments with a common (eta, delta, objective).

  target: complex128 numpy/torch array, [S-shaped] shared or [N, S-shaped] per env.
  obs_mode: 'full'  -> obs rows are the reference observation [state | target] (4*S floats; the target
                       half is written once here and never touched by the kernels),
            'state' -> rows hold only the state half (2*S floats), 'none' -> no observations.
  host_outputs: the per-step outputs and the observation rows live in PINNED HOST memory, which UVA maps
            into the device address space: the kernels write them over PCIe and the host reads them after a
            stream synchronisation, with no copy call (the single-env path; sensible for a few envs only).
  actions_stable: promise that every action tensor passed to `step` was completely written before the
            previous kernel of the stream was launched (pre-generated tapes): QCD_F_ACTIONS_STABLE lets
            the step kernel decode the action while the previous step drains.  Leave it off when a
            policy / sampler kernel writes the actions right before `step`.
  """

  def __init__(self, num_envs: int, max_qubits: int, max_depth: int, objective: str, target,
               punish: bool = True, sparse: bool = True, autoreset: bool = True, obs_mode: str = 'full',
               final_obs: bool = False, track_episodes: bool = True, fused_stats: bool = True, device=None,
               stats_buffer: torch.Tensor | None = None, actions_stable: bool = False,
               host_outputs: bool = False):
    if not torch.cuda.is_available():
      raise RuntimeError("qcd_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    self.lib = _lib.load_library()
    self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    self.kind = objective_kind(objective)
    if not 1 <= max_qubits <= _lib.MAX_QUBITS[self.kind]:
      raise ValueError(f"max_qubits={max_qubits} outside 1..{_lib.MAX_QUBITS[self.kind]} for {objective}")
    if not 1 <= max_depth <= _lib.MAX_DEPTH:
      raise ValueError(f"max_depth={max_depth} outside 1..{_lib.MAX_DEPTH}")
    if obs_mode not in ('full', 'state', 'none'):
      raise ValueError(obs_mode)
    self.num_envs, self.qubits, self.max_depth, self.objective = int(num_envs), int(max_qubits), int(max_depth), objective
    self.punish, self.sparse, self.autoreset, self.obs_mode = bool(punish), bool(sparse), bool(autoreset), obs_mode
    self.dim = 2 ** self.qubits
    self.S = self.dim if self.kind == _lib.OBJ_SP else self.dim * self.dim
    N, S, dev = self.num_envs, self.S, self.device

    tgt = torch.as_tensor(np.asarray(target) if not torch.is_tensor(target) else target).to(torch.complex128)
    if tgt.numel() == S:
      self.per_env_target = False
      tgt = tgt.reshape(S)
    elif tgt.numel() == N * S:
      self.per_env_target = True
      tgt = tgt.reshape(N, S)
    else:                     # e.g. SP-ghz3 on 4 qubits: the reference fails inside np.vdot (env.py:109)
      raise ValueError(f"target has {tgt.numel()} amplitudes, the state has {S}")
    self.target = tgt.contiguous().to(dev)

    f64, i32, u8 = torch.float64, torch.int32, torch.uint8
    self.host_outputs = bool(host_outputs)
    odev = dict(device='cpu', pin_memory=True) if host_outputs else dict(device=dev)     # where outputs live
    self.state = torch.zeros((N, S), dtype=torch.complex128, device=dev)
    self.meta = torch.zeros((N, 16), dtype=u8, device=dev)
    self.last = torch.zeros((N, 2), dtype=f64, device=dev)
    self.ep_return = torch.zeros(N, dtype=f64, device=dev) if track_episodes else None
    self.ep_length = torch.zeros(N, dtype=i32, device=dev) if track_episodes else None
    if host_outputs:
      # ONE pinned allocation, carved into the output arrays grouped by dtype: the host reads
      # [reward|metric|cost], [depth|operations|used_wires|episode_length], [terminated|truncated|status]
      # as three contiguous blocks (`host_blocks`)
      blob = torch.zeros(N * (4 * 8 + 4 * 4 + 3), dtype=u8, pin_memory=True)
      f = blob[:N * 32].view(f64); i = blob[N * 32:N * 48].view(i32); u = blob[N * 48:]
      self.reward, self.metric, self.cost, ep_r = f[:N], f[N:2 * N], f[2 * N:3 * N], f[3 * N:]
      self.depth, self.operations, self.used_wires, ep_l = i[:N], i[N:2 * N], i[2 * N:3 * N], i[3 * N:]
      self.terminated, self.truncated, self.status = u[:N], u[N:2 * N], u[2 * N:].view(torch.int8)
      self.episode_return = ep_r if track_episodes else None
      self.episode_length = ep_l if track_episodes else None
      self.host_blocks = (f.numpy(), i.numpy(), u.numpy())
      self._blob = blob
    else:
      self.episode_return = torch.zeros(N, dtype=f64, device=dev) if track_episodes else None
      self.episode_length = torch.zeros(N, dtype=i32, device=dev) if track_episodes else None
      self.reward = torch.zeros(N, dtype=f64, device=dev)
      self.metric = torch.zeros(N, dtype=f64, device=dev)
      self.cost = torch.zeros(N, dtype=f64, device=dev)
      self.depth = torch.zeros(N, dtype=i32, device=dev)
      self.operations = torch.zeros(N, dtype=i32, device=dev)
      self.used_wires = torch.zeros(N, dtype=i32, device=dev)
      self.terminated = torch.zeros(N, dtype=u8, device=dev)
      self.truncated = torch.zeros(N, dtype=u8, device=dev)
      self.status = torch.zeros(N, dtype=torch.int8, device=dev)
    if stats_buffer is None:
      self._stats_raw = torch.zeros((_lib.STATS_COPIES, _lib.STATS_STRIDE), dtype=f64, device=dev)
    else:                       # several batches of one job may accumulate into a common vector
      if tuple(stats_buffer.shape) != (_lib.STATS_COPIES, _lib.STATS_STRIDE) or stats_buffer.dtype != f64 \
          or stats_buffer.device != dev or not stats_buffer.is_contiguous():
        raise ValueError("stats_buffer must be a contiguous float64 [STATS_COPIES, STATS_STRIDE] tensor on the env device")
      self._stats_raw = stats_buffer

    self.obs = self.final_obs = None
    self.obs_stride = 2 * S
    if obs_mode != 'none':
      self.obs_stride = 4 * S if obs_mode == 'full' else 2 * S
      self.obs = torch.zeros((N, self.obs_stride), dtype=torch.float32, **odev)
      if final_obs:
        self.final_obs = torch.zeros((N, self.obs_stride), dtype=torch.float32, **odev)
      if obs_mode == 'full':                      # constant half: flat(target) (env.py:85)
        tflat = torch.cat([self.target.real, self.target.imag], dim=-1).to(torch.float32)
        if host_outputs: tflat = tflat.cpu()
        self.obs[:, 2 * S:] = tflat
        if self.final_obs is not None:
          self.final_obs[:, 2 * S:] = tflat

    flags = (_lib.F_PUNISH if self.punish else 0) | (_lib.F_SPARSE if self.sparse else 0) | \
            (_lib.F_AUTORESET if self.autoreset else 0) | (_lib.F_TARGET_PER_ENV if self.per_env_target else 0) | \
            (_lib.F_ACTIONS_STABLE if actions_stable else 0)
    d = self.max_depth
    self.c = QcdBatch(
        n_envs=N, n_qubits=self.qubits, objective=self.kind, max_depth=d, flags=flags, reserved0=0,
        depth_free=d / 3, cost_denom=d / 2 * 3,         # env.py:112, evaluated in Python float64
        obs_stride=self.obs_stride,
        state=_ptr(self.state), target=_ptr(self.target), meta=_ptr(self.meta), last=_ptr(self.last),
        ep_return=_ptr(self.ep_return), ep_length=_ptr(self.ep_length),
        obs=_ptr(self.obs), final_obs=_ptr(self.final_obs),
        reward=_ptr(self.reward), metric=_ptr(self.metric), cost=_ptr(self.cost),
        depth=_ptr(self.depth), operations=_ptr(self.operations), used_wires=_ptr(self.used_wires),
        terminated=_ptr(self.terminated), truncated=_ptr(self.truncated), status=_ptr(self.status),
        episode_return=_ptr(self.episode_return), episode_length=_ptr(self.episode_length),
        stats=_ptr(self._stats_raw) if fused_stats else None)
    self.fused_stats = bool(fused_stats)
    self._cref = C.byref(self.c)
    self.reset()

  # ------------------------------------------------------------------------------------------
  def _stream(self):
    return torch.cuda.current_stream(self.device).cuda_stream

  def _action_arg(self, actions: torch.Tensor, lead: tuple):
    if not torch.is_tensor(actions):
      raise TypeError("actions must be a torch tensor on the env's device (use step_host for host arrays)")
    if actions.device != self.device:
      raise ValueError(f"actions live on {actions.device}, the envs on {self.device}")
    if tuple(actions.shape) != lead + (self.num_envs, 4):
      raise ValueError(f"actions must have shape {lead + (self.num_envs, 4)}, got {tuple(actions.shape)}")
    if actions.dtype == torch.float32:
      code = _lib.ACT_F32
    elif actions.dtype == torch.float64:
      code = _lib.ACT_F64
    else:
      raise TypeError(f"actions must be float32 or float64, got {actions.dtype}")
    if not actions.is_contiguous():
      actions = actions.contiguous()
    return actions, code

  def reset(self, mask: torch.Tensor | None = None) -> None:
    """env.py:117-121 for all envs, or those with mask != 0 (uint8 [N] on device)."""
    if mask is not None:
      mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
    with torch.cuda.device(self.device):
      check(self.lib.qcd_reset(self._cref, _ptr(mask), self._stream()), "qcd_reset")

  def step(self, actions: torch.Tensor) -> None:
    """One env.step for every env (env.py:124-136); results land in the output tensors."""
    actions, code = self._action_arg(actions, ())
    with torch.cuda.device(self.device):
      check(self.lib.qcd_step(self._cref, actions.data_ptr(), code, self._stream()), "qcd_step")

  def step_sync(self, actions_ptr: int, code: int, stream_ptr: int) -> None:
    """qcd_step + one stream synchronisation in ONE C call; `actions_ptr` may be mapped pinned host memory
    (action in, results out over PCIe without a copy call)."""
    check(self.lib.qcd_step_sync(self._cref, actions_ptr, code, stream_ptr), "qcd_step_sync")

  def step1_sync(self, action4, stream_ptr: int) -> None:
    """env.step of a batch of ONE env: `action4` is a ctypes array of 4 doubles that travels as a kernel
    argument; returns after the stream synchronisation (outputs in mapped host memory are then valid)."""
    check(self.lib.qcd_step1_sync(self._cref, action4, stream_ptr), "qcd_step1_sync")

  def sub_batch(self, lo: int, hi: int) -> QcdBatch:
    """`qcd_batch_t` of the env range [lo, hi): the same buffers with every per-env pointer advanced by `lo`
    rows (a group of envs that is stepped on its own stream; statistics go to the common vector)."""
    if not 0 <= lo < hi <= self.num_envs:
      raise ValueError((lo, hi))
    c = QcdBatch()
    C.pointer(c)[0] = self.c
    c.n_envs = hi - lo
    def adv(name, t, row_elems=None):
      if t is None: return
      row = t.stride(0) if row_elems is None else row_elems
      setattr(c, name, t.data_ptr() + lo * row * t.element_size())
    adv('state', self.state); adv('meta', self.meta); adv('last', self.last)
    adv('ep_return', self.ep_return); adv('ep_length', self.ep_length)
    adv('obs', self.obs); adv('final_obs', self.final_obs)
    for name in ('reward', 'metric', 'cost', 'depth', 'operations', 'used_wires', 'terminated', 'truncated', 'status',
                 'episode_return', 'episode_length'):
      adv(name, getattr(self, name))
    if self.per_env_target:
      adv('target', self.target)
    return c

  def replay(self, tape: torch.Tensor, record: dict | None = None) -> None:
    """T fused steps with the states resident in shared memory.  tape: [T,N,4] on device.
    record: optional dict of preallocated [T,N] tensors keyed by reward/metric/cost/depth/
    terminated/truncated/status/operations/used_wires."""
    tape, code = self._action_arg(tape, (int(tape.shape[0]),))
    out = QcdTapeOut()
    want = {'reward': torch.float64, 'metric': torch.float64, 'cost': torch.float64, 'depth': torch.int32,
            'terminated': torch.uint8, 'truncated': torch.uint8, 'status': torch.int8, 'operations': torch.int32,
            'used_wires': torch.int32}
    for key, t in (record or {}).items():
      if key not in want:
        raise KeyError(key)
      if t.dtype != want[key] or tuple(t.shape) != (tape.shape[0], self.num_envs) or not t.is_contiguous() \
          or t.device != self.device:
        raise ValueError(f"record[{key!r}] must be a contiguous {want[key]} [T,N] tensor on {self.device}")
      setattr(out, key, t.data_ptr())
    with torch.cuda.device(self.device):
      check(self.lib.qcd_replay(self._cref, tape.data_ptr(), code, int(tape.shape[0]), C.byref(out),
                                self._stream()), "qcd_replay")

  def reduce_stats(self) -> None:
    """Add the finished-episode statistics of the last step to `self.stats` (stand-alone launch;
    a no-op when the step kernel already accumulates them, `fused_stats=True`)."""
    if self.fused_stats:
      return
    with torch.cuda.device(self.device):
      check(self.lib.qcd_reduce_stats(self._cref, self._stats_raw.data_ptr(), self._stream()), "qcd_reduce_stats")

  def push_episodes(self, ring: torch.Tensor, count: torch.Tensor) -> None:
    """Append the episodes that finished in the last step to the device ring (see `EpisodeWindow`)."""
    with torch.cuda.device(self.device):
      check(self.lib.qcd_push_episodes(self._cref, ring.data_ptr(), count.data_ptr(), int(ring.shape[0]), self._stream()),
            "qcd_push_episodes")

  @property
  def stats(self) -> torch.Tensor:
    """float64 [QCD_STATS_LEN] running episode statistics (sum over the atomic-spreading copies)."""
    return self._stats_raw.sum(0)[:_lib.STATS_LEN]

  def zero_stats(self) -> None:
    self._stats_raw.zero_()

  def fold_stats(self, out: torch.Tensor | None = None, zero: bool = False) -> torch.Tensor:
    """Sum the atomic-spreading replicas into `out` (float64 [STATS_STRIDE], device) with ONE small kernel on
    the current stream (capturable in a CUDA graph; no torch reduction), optionally clearing them: the
    vector a rank hands to the statistics all-reduce at the end of a reporting interval."""
    if out is None:
      out = torch.empty(_lib.STATS_STRIDE, dtype=torch.float64, device=self.device)
    with torch.cuda.device(self.device):
      check(self.lib.qcd_fold_stats(self._stats_raw.data_ptr(), out.data_ptr(), int(zero), self._stream()),
            "qcd_fold_stats")
    return out

  # ------------------------------------------------------------------------------------------
  # Checkpoint / resume of the environment state: everything a later `step` reads.
  _STATE_KEYS = ('state', 'meta', 'last', 'ep_return', 'ep_length', 'episode_return', 'episode_length', 'reward',
                 'metric', 'cost', 'depth', 'operations', 'used_wires', 'terminated', 'truncated', 'status', 'obs',
                 'final_obs', '_stats_raw')

  def state_dict(self) -> dict:
    """Host copies of every persistent buffer + the configuration they belong to (synchronises)."""
    cfg = dict(num_envs=self.num_envs, max_qubits=self.qubits, max_depth=self.max_depth, objective=self.objective,
               punish=self.punish, sparse=self.sparse, autoreset=self.autoreset, obs_mode=self.obs_mode)
    bufs = {k: getattr(self, k).detach().cpu().clone() for k in self._STATE_KEYS if getattr(self, k) is not None}
    return {'config': cfg, 'target': self.target.cpu().clone(), 'buffers': bufs}

  def load_state_dict(self, sd: dict) -> None:
    """Restore a `state_dict()` of a batch with the same configuration (shapes are checked)."""
    mine = self.state_dict()['config']
    if sd['config'] != mine:
      raise ValueError(f"checkpoint is for {sd['config']}, this batch is {mine}")
    if sd['target'].shape != self.target.shape or not torch.equal(sd['target'], self.target.cpu()):
      raise ValueError("checkpoint was taken with another target")
    for k, v in sd['buffers'].items():
      dst = getattr(self, k)
      if dst is None or dst.shape != v.shape or dst.dtype != v.dtype:
        raise ValueError(f"buffer {k} does not match")
      dst.copy_(v)

  def state_view(self) -> torch.Tensor:
    """complex128 [N,D] (SP) or [N,D,D] (UC) view of the resident states."""
    return self.state if self.kind == _lib.OBJ_SP else self.state.view(self.num_envs, self.dim, self.dim)

  def algorithmic_bytes_per_step(self) -> int:
    """32*S state r/w + 8*S float32 observation + 96 B scalars."""
    return 32 * self.S + (8 * self.S if self.obs is not None else 0) + 96


class EpisodeWindow:
  """The most recent `window` finished episodes of a batch, on the device.

  What SB3 keeps as `ep_info_buffer = deque(maxlen=100)` of Monitor's info["episode"] records and the reference's
  trainer averages for its progress bar and the `return-100` / `length-100` curves
  (/root/reference/algorithm/algorithm.py:49,89-91,100; monitor.py:37-48).  `push()` after every step is one small
  launch, no synchronisation; `means()` / `episodes()` read the ring back."""

  def __init__(self, batch: EnvBatch, window: int = 100):
    if batch.host_outputs:
      raise ValueError("EpisodeWindow reads device buffers; a host_outputs batch keeps its episode records on the host")
    self.batch, self.window = batch, int(window)
    self.ring = torch.zeros((self.window, _lib.EP_FIELDS), dtype=torch.float64, device=batch.device)
    self.count = torch.zeros(1, dtype=torch.int64, device=batch.device)

  def push(self) -> None:
    self.batch.push_episodes(self.ring, self.count)

  def reset(self) -> None:
    self.ring.zero_(); self.count.zero_()

  def episodes(self) -> list:
    """Oldest -> newest list of {'r','l','d','o','q','m','c','env'} dicts (synchronises)."""
    n = int(self.count.item())
    rows = self.ring.cpu().numpy()
    k = min(n, self.window)
    order = [(n - k + i) % self.window for i in range(k)]
    out = []
    for j in order:
      d = dict(zip(_lib.EP_NAMES, rows[j].tolist()))
      for key in ('l', 'd', 'o', 'q', 'env'):
        d[key] = int(d[key])
      out.append(d)
    return out

  def means(self) -> dict:
    """np.mean([ep[k] for ep in ep_info_buffer]) for k in r,l,d,o,q,m,c (synchronises); NaN while empty."""
    n = min(int(self.count.item()), self.window)
    if n == 0:
      return {k: float('nan') for k in _lib.EP_NAMES[:7]}
    m = self.ring[:n].mean(0).cpu().tolist() if n < self.window else self.ring.mean(0).cpu().tolist()
    return dict(zip(_lib.EP_NAMES[:7], m[:7]))
