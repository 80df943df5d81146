"""numpy/ctypes front-end of oracle/qcd_oracle_batch.c (TEST INFRASTRUCTURE, see that file's header).

`OracleBatch` mirrors the buffer set of `qcd_b200.batch.EnvBatch` with numpy arrays, so a parity test
can compare the two buffer by buffer after every step.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle_batch.so")
SRC = os.path.join(HERE, "qcd_oracle_batch.c")


def build(force: bool = False) -> str:
  if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
    subprocess.run(["make", "-C", HERE, "-B", "liboracle_batch.so"], check=True, capture_output=True)
  return LIB


class _Batch(C.Structure):
  _fields_ = [
      ("n_envs", C.c_int32), ("n_qubits", C.c_int32), ("objective", C.c_int32), ("max_depth", C.c_int32),
      ("flags", C.c_uint32), ("reserved0", C.c_int32),
      ("depth_free", C.c_double), ("cost_denom", C.c_double), ("obs_stride", C.c_int64),
      ("state", C.c_void_p), ("target", C.c_void_p), ("meta", C.c_void_p), ("last", C.c_void_p),
      ("ep_return", C.c_void_p), ("ep_length", C.c_void_p),
      ("obs", C.c_void_p), ("final_obs", C.c_void_p),
      ("reward", C.c_void_p), ("metric", C.c_void_p), ("cost", C.c_void_p),
      ("depth", C.c_void_p), ("operations", C.c_void_p), ("used_wires", C.c_void_p),
      ("terminated", C.c_void_p), ("truncated", C.c_void_p), ("status", C.c_void_p),
      ("episode_return", C.c_void_p), ("episode_length", C.c_void_p), ("stats", C.c_void_p),
  ]


_lib = None


def _load():
  global _lib
  if _lib is None:
    _lib = C.CDLL(build())
    _lib.oracle_reset.argtypes = [C.POINTER(_Batch), C.c_void_p]
    _lib.oracle_step.argtypes = [C.POINTER(_Batch), C.c_void_p]
  return _lib


class OracleBatch:
  def __init__(self, num_envs, max_qubits, max_depth, objective, target, punish=True, sparse=True,
               autoreset=True, obs_mode='full', final_obs=False):
    self.lib = _load()
    uc = 'UC' in objective
    self.num_envs, self.qubits, self.max_depth = num_envs, max_qubits, max_depth
    self.S = S = 2 ** (2 * max_qubits if uc else max_qubits)
    N = num_envs
    tgt = np.ascontiguousarray(np.asarray(target, dtype=np.complex128))
    self.per_env_target = tgt.size == N * S and N > 1
    assert tgt.size in (S, N * S)
    self.target = tgt.reshape(-1)
    self.state = np.zeros((N, S), np.complex128)
    self.meta = np.zeros((N, 16), np.uint8)
    self.last = np.zeros((N, 2), np.float64)
    self.ep_return = np.zeros(N, np.float64); self.ep_length = np.zeros(N, np.int32)
    self.episode_return = np.zeros(N, np.float64); self.episode_length = np.zeros(N, np.int32)
    self.reward = np.zeros(N, np.float64); self.metric = np.zeros(N, np.float64); self.cost = np.zeros(N, np.float64)
    self.depth = np.zeros(N, np.int32); self.operations = np.zeros(N, np.int32); self.used_wires = np.zeros(N, np.int32)
    self.terminated = np.zeros(N, np.uint8); self.truncated = np.zeros(N, np.uint8); self.status = np.zeros(N, np.int8)
    self.stats = np.zeros(12, np.float64)
    self.obs = self.final_obs = None
    self.obs_stride = 2 * S
    if obs_mode != 'none':
      self.obs_stride = 4 * S if obs_mode == 'full' else 2 * S
      self.obs = np.zeros((N, self.obs_stride), np.float32)
      if final_obs: self.final_obs = np.zeros((N, self.obs_stride), np.float32)
      if obs_mode == 'full':
        t2 = tgt.reshape(-1, S)
        tflat = np.concatenate([t2.real, t2.imag], axis=-1).astype(np.float32)
        self.obs[:, 2 * S:] = tflat
        if final_obs: self.final_obs[:, 2 * S:] = tflat
    p = lambda a: None if a is None else a.ctypes.data
    flags = (1 if punish else 0) | (2 if sparse else 0) | (4 if autoreset else 0) | (8 if self.per_env_target else 0)
    self.c = _Batch(N, max_qubits, 1 if uc else 0, max_depth, flags, 0, max_depth / 3, max_depth / 2 * 3,
                    self.obs_stride, p(self.state), p(self.target), p(self.meta), p(self.last),
                    p(self.ep_return), p(self.ep_length), p(self.obs), p(self.final_obs),
                    p(self.reward), p(self.metric), p(self.cost), p(self.depth), p(self.operations),
                    p(self.used_wires), p(self.terminated), p(self.truncated), p(self.status),
                    p(self.episode_return), p(self.episode_length), p(self.stats))
    self.reset()

  def reset(self, mask=None):
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    self.lib.oracle_reset(C.byref(self.c), None if m is None else m.ctypes.data)

  def step(self, actions):
    a = np.ascontiguousarray(np.asarray(actions, dtype=np.float64).reshape(self.num_envs, 4))
    self.lib.oracle_step(C.byref(self.c), a.ctypes.data)
