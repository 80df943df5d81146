"""Build recipe for libqcd_b200.so (sm_100a only, in-tree so that the .so travels with the repo)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "qcd_b200", "csrc")
LIB_PATH = os.path.join(ROOT, "qcd_b200", "libqcd_b200.so")
STAMP_PATH = LIB_PATH + ".srchash"        # content hash of the sources the library was built from
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


def source_hash() -> str:
  """sha256 over the sources, the header and the base flags (mtimes do not survive a copy of the tree to
  another box; contents do).  QCD_NVCC_EXTRA is deliberately not part of it: a library built with tuning
  knobs stays valid for processes that do not carry the variable."""
  h = hashlib.sha256()
  for p in SOURCES + HEADERS:
    with open(p, "rb") as f:
      h.update(f.read())
  h.update(" ".join(NVCC_FLAGS[:-1]).encode())
  return h.hexdigest()


def is_stale() -> bool:
  """True when libqcd_b200.so is missing or was built from other sources than the ones in the tree."""
  if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP_PATH):
    return True
  with open(STAMP_PATH) as f:
    return f.read().strip() != source_hash()


def build_library(force: bool = False, verbose: bool = False) -> str:
  """Compile csrc/*.cu -> qcd_b200/libqcd_b200.so.  Returns the library path."""
  if not force and not is_stale():
    return LIB_PATH
  extra = os.environ.get("QCD_NVCC_EXTRA", "").split()      # tuning knobs, e.g. -DQCD_WARP_CHUNK=4
  cmd = [find_nvcc(), *NVCC_FLAGS, *extra, "-o", LIB_PATH, *SOURCES]
  if verbose:
    cmd.insert(1, "-Xptxas=-v")
  res = subprocess.run(cmd, capture_output=True, text=True)
  if res.returncode != 0:
    raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
  with open(STAMP_PATH, "w") as f:
    f.write(source_hash() + "\n")
  if verbose:
    print(res.stderr)
  return LIB_PATH


if __name__ == "__main__":
  print(build_library(force=True, verbose=True))
