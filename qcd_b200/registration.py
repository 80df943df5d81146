"""Plugin boundary: `CircuitDesigner-v0` in the gymnasium registry.

Reference: entry-point group `gymnasium.envs`, `__root__ = qcd_gym.__init__:register_envs`
(/root/reference/setup.py:21) -> `register(id="CircuitDesigner-v0", entry_point="qcd_gym.env:CircuitDesigner")`
(/root/reference/qcd_gym/__init__.py:3-5).  Same id, same kwargs; the entry point is this package's env.
"""
from .gym_compat import register, registry

ENV_ID = "CircuitDesigner-v0"


def register_envs():
  if ENV_ID in registry: return
  register(id=ENV_ID, entry_point="qcd_b200.env:CircuitDesigner")
