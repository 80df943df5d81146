"""qcd_b200 — B200-native batched simulator behind the `CircuitDesigner-v0` gymnasium API.

Public surface (mirrors /root/reference/qcd_gym):
  register_envs()                      qcd_gym/__init__.py:3-5
  CircuitDesigner                      qcd_gym/env.py:13-141 (single env, numpy in/out)
  CircuitDesignerVectorEnv             N envs resident in HBM, one launch per step
  Monitor                              qcd_gym/wrappers/monitor.py:6-76
"""
from .registration import register_envs, ENV_ID
from .targets import make_target, parse_task

__all__ = ["register_envs", "ENV_ID", "make_target", "parse_task", "CircuitDesigner",
           "CircuitDesignerVectorEnv", "EnvBatch", "Monitor"]

__version__ = "0.1.0"


def __getattr__(name):          # torch / the CUDA library are only touched when an env is requested
  if name == "CircuitDesigner":
    from .env import CircuitDesigner
    return CircuitDesigner
  if name == "CircuitDesignerVectorEnv":
    from .vector_env import CircuitDesignerVectorEnv
    return CircuitDesignerVectorEnv
  if name == "EnvBatch":
    from .batch import EnvBatch
    return EnvBatch
  if name == "Monitor":
    from .monitor import Monitor
    return Monitor
  raise AttributeError(name)
