"""ctypes binding of include/qcd_b200.h.  There is no CPU fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH

ABI_VERSION = 6
STATS_LEN = 12
STATS_COPIES = 64
STATS_STRIDE = 16
STAT_NAMES = ("episodes", "sum_r", "sum_l", "sum_d", "sum_o", "sum_q", "sum_m", "sum_c",
              "n_done", "n_depth", "steps", "bad_actions")

F_PUNISH, F_SPARSE, F_AUTORESET, F_TARGET_PER_ENV, F_ACTIONS_STABLE = 1, 2, 4, 8, 16
OBJ_SP, OBJ_UC = 0, 1
ACT_F64, ACT_F32 = 0, 1
ST_OK, ST_BAD_WIRE, ST_BAD_GATE = 0, 1, 2
MAX_QUBITS = {OBJ_SP: 12, OBJ_UC: 6}
MAX_DEPTH = 255

ERRORS = {-1: "QCD_E_NULL (required pointer is NULL)", -2: "QCD_E_RANGE (n_envs/n_qubits/max_depth out of range)",
          -3: "QCD_E_OBJ (unknown objective or dtype)", -4: "QCD_E_STRIDE (obs_stride < 2*S)"}


class QcdBatch(C.Structure):
  """Mirror of `qcd_batch_t` (include/qcd_b200.h)."""
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


class QcdTapeOut(C.Structure):
  """Mirror of `qcd_tape_out_t`."""
  _fields_ = [("reward", C.c_void_p), ("metric", C.c_void_p), ("cost", C.c_void_p), ("depth", C.c_void_p),
              ("terminated", C.c_void_p), ("truncated", C.c_void_p), ("status", C.c_void_p),
              ("operations", C.c_void_p), ("used_wires", C.c_void_p)]


class QcdHostOut(C.Structure):
  """Mirror of `qcd_host_out_t` (host destinations of qcd_step_host)."""
  _fields_ = [("reward", C.c_void_p), ("metric", C.c_void_p), ("cost", C.c_void_p), ("depth", C.c_void_p),
              ("operations", C.c_void_p), ("used_wires", C.c_void_p), ("terminated", C.c_void_p),
              ("truncated", C.c_void_p), ("status", C.c_void_p), ("obs", C.c_void_p), ("obs_stride", C.c_int64)]


EP_FIELDS = 8
EP_NAMES = ("r", "l", "d", "o", "q", "m", "c", "env")
EXPORTS = ("qcd_abi_version", "qcd_reset", "qcd_step", "qcd_replay", "qcd_reduce_stats", "qcd_fold_stats", "qcd_push_episodes",
           "qcd_step_sync", "qcd_step1_sync", "qcd_step_host")

_lib = None


def load_library(path: str | None = None) -> C.CDLL:
  """dlopen libqcd_b200.so and declare the prototypes.  Raises if it has not been built."""
  global _lib
  if _lib is not None and path is None:
    return _lib
  # QCD_B200_LIB: a library built elsewhere from the same header (tuning variants of scripts/variants.sh); it is
  # loaded as is: the ABI check below still applies, the staleness check cannot
  cache = path is None
  path = path or os.environ.get("QCD_B200_LIB") or None
  p = path or LIB_PATH
  if path is None:
    from .build import build_library, is_stale
    if is_stale():                        # missing, or built from other sources than the tree holds
      try:                                # compile the real thing (needs nvcc) ...
        build_library()
      except Exception as exc:            # ... and fail loudly otherwise: there is nothing to fall back to
        raise RuntimeError(
            f"{p} is missing or stale and could not be rebuilt ({exc}). Build it with "
            "`python -m qcd_b200.build` (or __graft_entry__.build()). qcd_b200 has no CPU fallback.") from exc
  if not os.path.exists(p):
    raise RuntimeError(f"{p} is missing. qcd_b200 has no CPU fallback.")
  lib = C.CDLL(p)
  lib.qcd_abi_version.restype = C.c_int
  lib.qcd_abi_version.argtypes = []
  lib.qcd_reset.restype = C.c_int
  lib.qcd_reset.argtypes = [C.POINTER(QcdBatch), C.c_void_p, C.c_void_p]
  lib.qcd_step.restype = C.c_int
  lib.qcd_step.argtypes = [C.POINTER(QcdBatch), C.c_void_p, C.c_int, C.c_void_p]
  lib.qcd_replay.restype = C.c_int
  lib.qcd_replay.argtypes = [C.POINTER(QcdBatch), C.c_void_p, C.c_int, C.c_int32, C.POINTER(QcdTapeOut), C.c_void_p]
  lib.qcd_reduce_stats.restype = C.c_int
  lib.qcd_reduce_stats.argtypes = [C.POINTER(QcdBatch), C.c_void_p, C.c_void_p]
  lib.qcd_fold_stats.restype = C.c_int
  lib.qcd_fold_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
  lib.qcd_push_episodes.restype = C.c_int
  lib.qcd_push_episodes.argtypes = [C.POINTER(QcdBatch), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
  lib.qcd_step_sync.restype = C.c_int
  lib.qcd_step_sync.argtypes = [C.POINTER(QcdBatch), C.c_void_p, C.c_int, C.c_void_p]
  lib.qcd_step1_sync.restype = C.c_int
  lib.qcd_step1_sync.argtypes = [C.POINTER(QcdBatch), C.POINTER(C.c_double), C.c_void_p]
  lib.qcd_step_host.restype = C.c_int
  lib.qcd_step_host.argtypes = [C.POINTER(QcdBatch), C.c_void_p, C.c_void_p, C.c_int, C.POINTER(QcdHostOut), C.c_void_p]
  if lib.qcd_abi_version() != ABI_VERSION:
    raise RuntimeError(f"libqcd_b200.so ABI {lib.qcd_abi_version()} != binding ABI {ABI_VERSION}: rebuild")
  if cache:
    _lib = lib
  return lib


def check(rc: int, what: str) -> None:
  if rc == 0:
    return
  if rc < 0:
    raise ValueError(f"{what}: {ERRORS.get(rc, rc)}")
  raise RuntimeError(f"{what}: CUDA error {rc}")
