"""Episode monitor for the single-env API.

Behavioural contract taken from the reference wrapper (qcd_gym/wrappers/monitor.py:29-52): while an
episode runs it records observations, actions and rewards; when `step` reports terminated or truncated
it moves `termination_reason` out of `info` into its own history and attaches

    info["episode"] = {r, l, t, history, d, o, q, m, c}

(return rounded to 6 digits, length, wall time since construction, the recorded trajectory, and the
final depth / operations / used wires / metric / cost); stepping again before `reset` raises
RuntimeError.  The vectorised env keeps the same r/l accumulators on the device inside the step kernel.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np

from .gym_compat import Wrapper

_QMETRICS = {'d': 'depth', 'o': 'operations', 'q': 'used_wires', 'm': 'metric', 'c': 'cost'}


@dataclass
class _Trajectory:
  states: list = field(default_factory=list)
  actions: list = field(default_factory=list)
  rewards: list = field(default_factory=list)

  def snapshot(self) -> dict:
    return {'states': list(self.states), 'actions': list(self.actions), 'rewards': list(self.rewards)}


class Monitor(Wrapper):
  def __init__(self, env, record_video: bool = False):
    super().__init__(env)
    self.t_start = time.time()
    self.record_video = record_video
    self.needs_reset = True
    self._traj = _Trajectory()
    self._frames: list = []
    self._log = {'returns': [], 'lengths': [], 'times': [], 'reasons': []}
    self._total_steps = 0

  # -- trajectory of the running episode, under the reference's attribute names
  states = property(lambda self: self._traj.states)
  actions = property(lambda self: self._traj.actions)
  rewards = property(lambda self: self._traj.rewards)

  def _grab_frame(self):
    if self.record_video:
      self._frames.append(self.render())

  def reset(self, **kwargs):
    observation, info = self.env.reset(**kwargs)
    self._traj = _Trajectory(states=[observation])
    self.needs_reset = False
    self._grab_frame()
    return observation, info

  def step(self, action):
    if self.needs_reset:
      raise RuntimeError("Tried to step environment that needs reset")
    observation, reward, terminated, truncated, info = self.env.step(action)
    traj = self._traj
    traj.states.append(observation); traj.actions.append(action); traj.rewards.append(float(reward))
    self._grab_frame()
    self._total_steps += 1
    if terminated or truncated:
      self._close_episode(info)
    return observation, reward, terminated, truncated, info

  def _close_episode(self, info: dict) -> None:
    elapsed = time.time() - self.t_start
    total, length = sum(self._traj.rewards), len(self._traj.rewards)
    log = self._log
    log['returns'].append(total); log['lengths'].append(length); log['times'].append(elapsed)
    log['reasons'].append(info.pop('termination_reason'))
    episode = {'r': round(total, 6), 'l': length, 't': round(elapsed, 6), 'history': self._traj.snapshot()}
    episode.update({short: info[key] for short, key in _QMETRICS.items()})
    info['episode'] = episode
    self.needs_reset = True

  # -- video helpers (text mode only, like the reference's write_video)
  def get_video(self, reset: bool = True):
    frames = np.array(self._frames)
    if reset:
      self._frames = []
    return frames

  def write_video(self, writer, label, step):
    if self.render_mode == 'text':
      writer.add_text(label, str(self.env.render()), step)

  # -- accumulated history
  @property
  def total_steps(self) -> int:
    return self._total_steps

  @property
  def episode_returns(self) -> list:
    return self._log['returns']

  @property
  def episode_lengths(self) -> list:
    return self._log['lengths']

  @property
  def episode_times(self) -> list:
    return self._log['times']

  @property
  def termination_reasons(self) -> list:
    return self._log['reasons']
