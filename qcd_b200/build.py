"""Build recipe for libqcd_b200.so (sm_100a only, in-tree so that the .so travels with the repo)."""
from __future__ import annotations

import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "qcd_b200", "csrc")
LIB_PATH = os.path.join(ROOT, "qcd_b200", "libqcd_b200.so")
SOURCES = [os.path.join(CSRC, "qcd_kernels.cu")]
HEADERS = [os.path.join(ROOT, "include", "qcd_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "-I", os.path.join(ROOT, "include"),
]


def find_nvcc() -> str:
  nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
  if not os.path.exists(nvcc):
    raise RuntimeError("nvcc not found: libqcd_b200.so cannot be built (no CPU fallback exists)")
  return nvcc


def is_stale() -> bool:
  if not os.path.exists(LIB_PATH):
    return True
  built = os.path.getmtime(LIB_PATH)
  return any(os.path.getmtime(p) > built for p in SOURCES + HEADERS)


def build_library(force: bool = False, verbose: bool = False) -> str:
  """Compile csrc/*.cu -> qcd_b200/libqcd_b200.so.  Returns the library path."""
  if not force and not is_stale():
    return LIB_PATH
  extra = os.environ.get("QCD_NVCC_EXTRA", "").split()      # tuning knobs, e.g. -DQCD_REG_MINB=3
  cmd = [find_nvcc(), *NVCC_FLAGS, *extra, "-o", LIB_PATH, *SOURCES]
  if verbose:
    cmd.insert(1, "-Xptxas=-v")
  res = subprocess.run(cmd, capture_output=True, text=True)
  if res.returncode != 0:
    raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
  if verbose:
    print(res.stderr)
  return LIB_PATH


if __name__ == "__main__":
  print(build_library(force=True, verbose=True))
