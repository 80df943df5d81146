"""gymnasium facade.

The reference plugs into `gymnasium==0.29` (/root/reference/setup.py:14,21).  When gymnasium is
importable it is used unchanged; this image does not ship it, so the few pieces the hot path's
boundary touches are restated here: `spaces.Box` (bounds stored in the Box dtype, `sample`,
`contains`, `seed`), `Env`, `Wrapper`, `utils.seeding.np_random`, and the `register` / `make`
registry with the `OrderEnforcing` guard `gym.make` adds.
"""
from __future__ import annotations

import importlib

import numpy as np

try:                                    # pragma: no cover - not available in this image
  import gymnasium as _gym
  HAVE_GYMNASIUM = True
except Exception:                       # ModuleNotFoundError, or a broken install
  _gym = None
  HAVE_GYMNASIUM = False


if HAVE_GYMNASIUM:                      # pragma: no cover
  Box = _gym.spaces.Box
  Env = _gym.Env
  Wrapper = _gym.Wrapper
  registry = _gym.envs.registration.registry
  register = _gym.envs.registration.register
  make = _gym.make
  make_vec = getattr(_gym, "make_vec", None)
  np_random = _gym.utils.seeding.np_random
  ResetNeeded = _gym.error.ResetNeeded
  try:
    VectorEnvBase = _gym.vector.VectorEnv
  except Exception:
    VectorEnvBase = object
else:
  VectorEnvBase = object                 # gymnasium.vector.VectorEnv when gymnasium is importable


  class ResetNeeded(Exception):
    """gymnasium.error.ResetNeeded"""

  def np_random(seed=None):
    """gymnasium.utils.seeding.np_random: (Generator(PCG64(SeedSequence(seed))), seed)."""
    if seed is not None and not (isinstance(seed, (int, np.integer)) and 0 <= seed):
      raise ValueError(f"Seed must be a python integer, actual type: {type(seed)}")
    seed_seq = np.random.SeedSequence(seed)
    np_seed = seed_seq.entropy
    return np.random.Generator(np.random.PCG64(seed_seq)), np_seed

  class Box:
    """gymnasium.spaces.Box restricted to bounded (or one-sided) real intervals."""

    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
      self.dtype = np.dtype(dtype)
      if shape is None:
        shape = np.shape(low) if not np.isscalar(low) else np.shape(high)
      self._shape = tuple(int(s) for s in shape)
      self.low = np.broadcast_to(np.asarray(low, dtype=np.float64), self._shape).astype(self.dtype)
      self.high = np.broadcast_to(np.asarray(high, dtype=np.float64), self._shape).astype(self.dtype)
      self._np_random = None
      if seed is not None:
        self.seed(seed)

    @property
    def shape(self):
      return self._shape

    @property
    def np_random(self):
      if self._np_random is None:
        self.seed()
      return self._np_random

    def seed(self, seed=None):
      self._np_random, seed = np_random(seed)
      return [seed]

    def sample(self):
      """Bounded dims: uniform(low, high) in float64, cast to the Box dtype (gymnasium 0.29)."""
      return self.np_random.uniform(low=self.low, high=self.high, size=self._shape).astype(self.dtype)

    def contains(self, x) -> bool:
      x = np.asarray(x)
      return bool(x.shape == self._shape and np.all(x >= self.low) and np.all(x <= self.high))

    __contains__ = contains

    def __repr__(self):
      return f"Box({self.low.min()}, {self.high.max()}, {self._shape}, {self.dtype})"

  class Env:
    """gymnasium.Env: spaces, metadata, `np_random` reseeded by reset(seed=...)."""
    metadata = {"render_modes": []}
    render_mode = None
    spec = None
    observation_space = None
    action_space = None
    _np_random = None

    @property
    def np_random(self):
      if self._np_random is None:
        self._np_random, _ = np_random()
      return self._np_random

    @np_random.setter
    def np_random(self, value):
      self._np_random = value

    @property
    def unwrapped(self):
      return self

    def reset(self, *, seed=None, options=None):
      if seed is not None:
        self._np_random, _ = np_random(seed)

    def step(self, action):
      raise NotImplementedError

    def render(self):
      raise NotImplementedError

    def close(self):
      pass

  class Wrapper(Env):
    """gymnasium.Wrapper: forwards everything to `env`."""

    def __init__(self, env):
      self.env = env

    def __getattr__(self, name):
      if name.startswith('_'):
        raise AttributeError(name)
      return getattr(self.env, name)

    @property
    def unwrapped(self):
      return self.env.unwrapped

    @property
    def observation_space(self):
      return self.env.observation_space

    @property
    def action_space(self):
      return self.env.action_space

    @property
    def metadata(self):
      return self.env.metadata

    @property
    def render_mode(self):
      return self.env.render_mode

    def reset(self, **kwargs):
      return self.env.reset(**kwargs)

    def step(self, action):
      return self.env.step(action)

    def render(self):
      return self.env.render()

    def close(self):
      return self.env.close()

  class OrderEnforcing(Wrapper):
    """The guard gym.make wraps around every env: step()/render() before reset() raise."""

    def __init__(self, env):
      super().__init__(env)
      self._has_reset = False

    def step(self, action):
      if not self._has_reset:
        raise ResetNeeded("Cannot call env.step() before calling env.reset()")
      return self.env.step(action)

    def reset(self, **kwargs):
      self._has_reset = True
      return self.env.reset(**kwargs)

  class _Spec:
    def __init__(self, id, entry_point, kwargs, vector_entry_point=None):
      self.id, self.entry_point, self.kwargs = id, entry_point, dict(kwargs or {})
      self.vector_entry_point = vector_entry_point

  registry: dict = {}

  def register(id, entry_point=None, vector_entry_point=None, **kwargs):
    registry[id] = _Spec(id, entry_point, kwargs.get('kwargs'), vector_entry_point)

  def _load(entry):
    if isinstance(entry, str):
      mod, attr = entry.split(':')
      entry = getattr(importlib.import_module(mod), attr)
    return entry

  def make_vec(id, num_envs=1, **kwargs):
    """gymnasium.make_vec for specs that carry a vector_entry_point."""
    if id not in registry:
      raise KeyError(f"Environment `{id}` is not registered")
    spec = registry[id]
    if spec.vector_entry_point is None:
      raise KeyError(f"Environment `{id}` has no vector_entry_point")
    env = _load(spec.vector_entry_point)(num_envs=num_envs, **{**spec.kwargs, **kwargs})
    env.spec = spec
    return env

  def make(id, **kwargs):
    if id not in registry:
      raise KeyError(f"Environment `{id}` is not registered")
    spec = registry[id]
    entry = spec.entry_point
    if isinstance(entry, str):
      mod, attr = entry.split(':')
      entry = getattr(importlib.import_module(mod), attr)
    env = entry(**{**spec.kwargs, **kwargs})
    env.spec = spec
    return OrderEnforcing(env)
