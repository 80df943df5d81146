"""Plugin boundary: `CircuitDesigner-v0` in the gymnasium registry.

Reference: entry-point group `gymnasium.envs`, `__root__ = qcd_gym.__init__:register_envs`
(/root/reference/setup.py:21) -> `register(id="CircuitDesigner-v0", entry_point="qcd_gym.env:CircuitDesigner")`
(/root/reference/qcd_gym/__init__.py:3-5).  Same id, same kwargs; the entry point is this package's env, and the
registration also names the batched env as `vector_entry_point`, so that `gymnasium.make_vec("CircuitDesigner-v0",
num_envs=N, ...)` builds N envs resident in HBM instead of N Python objects (gymnasium releases without that
argument get the plain registration).  This repo's setup.py declares the same `gymnasium.envs` entry point, so an
installed copy is discovered by `gymnasium.make` without an explicit import.
"""
import inspect

from .gym_compat import register, registry

ENV_ID = "CircuitDesigner-v0"
ENTRY_POINT = "qcd_b200.env:CircuitDesigner"
VECTOR_ENTRY_POINT = "qcd_b200.vector_env:CircuitDesignerVectorEnv"


def register_envs():
  if ENV_ID in registry: return
  kwargs = {}
  try:
    if "vector_entry_point" in inspect.signature(register).parameters:
      kwargs["vector_entry_point"] = VECTOR_ENTRY_POINT
  except (TypeError, ValueError):
    pass
  register(id=ENV_ID, entry_point=ENTRY_POINT, **kwargs)
