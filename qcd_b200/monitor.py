"""Monitor wrapper for the single-env API (reference: qcd_gym/wrappers/monitor.py:6-76).

Records episode return / length / wall time and the Q-metrics d,o,q,m,c into info['episode'] when
an episode ends; refuses to step an env that needs a reset.  The vectorised env does the same
bookkeeping on the device (ep_return / ep_length accumulators inside the step kernel)."""
from __future__ import annotations

import time

import numpy as np

from .gym_compat import Wrapper


class Monitor(Wrapper):
  def __init__(self, env, record_video=False):
    super().__init__(env)
    self.t_start = time.time()
    self.record_video = record_video; self._frame_buffer = []
    self.states, self.actions, self.rewards = [], [], []
    self._episode_returns, self._termination_reasons = [], []
    self._episode_lengths, self._episode_times = [], []
    self._total_steps = 0; self.needs_reset = True

  def _history(self):
    return {key: list(getattr(self, key)) for key in ('states', 'actions', 'rewards')}

  def reset(self, **kwargs):
    self.needs_reset = False
    state, info = self.env.reset(**kwargs)
    self.rewards = []; self.states = [state]; self.actions = []
    if self.record_video: self._frame_buffer.append(self.render())
    return state, info

  def step(self, action):
    if self.needs_reset: raise RuntimeError("Tried to step environment that needs reset")
    state, reward, terminated, truncated, info = self.env.step(action)
    self.states.append(state); self.actions.append(action); self.rewards.append(float(reward))
    if self.record_video: self._frame_buffer.append(self.render())
    if terminated or truncated:
      self.needs_reset = True
      ep_rew, ep_len, now = sum(self.rewards), len(self.rewards), time.time() - self.t_start
      self._episode_returns.append(ep_rew); self._episode_lengths.append(ep_len); self._episode_times.append(now)
      self._termination_reasons.append(info.pop('termination_reason'))
      info["episode"] = {"r": round(ep_rew, 6), "l": ep_len, "t": round(now, 6), 'history': self._history(),
                         'd': info["depth"], 'o': info["operations"], 'q': info["used_wires"],
                         'm': info["metric"], 'c': info["cost"]}
    self._total_steps += 1
    return state, reward, terminated, truncated, info

  def get_video(self, reset=True):
    frames = list(self._frame_buffer)
    if reset: self._frame_buffer = []
    return np.array(frames)

  def write_video(self, writer, label, step):
    if self.render_mode == 'text': writer.add_text(label, str(self.env.render()), step)

  total_steps = property(lambda self: self._total_steps)
  episode_returns = property(lambda self: self._episode_returns)
  termination_reasons = property(lambda self: self._termination_reasons)
  episode_lengths = property(lambda self: self._episode_lengths)
  episode_times = property(lambda self: self._episode_times)
