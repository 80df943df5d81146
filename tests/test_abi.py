"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/qcd_b200.h declares, the ctypes mirror of qcd_batch_t has the C layout, argument validation
returns the documented error codes (no kernel is launched), and the gymnasium plugin registers."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from qcd_b200 import _lib
from qcd_b200.build import LIB_PATH, build_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
  build_library()
  return _lib.load_library()


def test_header_symbols_exported(lib):
  header = open(os.path.join(ROOT, "include", "qcd_b200.h")).read()
  declared = re.findall(r"^int (qcd_\w+)\(", header, flags=re.M)
  assert sorted(declared) == sorted(_lib.EXPORTS)
  for name in declared:
    assert hasattr(lib, name), name
  assert lib.qcd_abi_version() == _lib.ABI_VERSION


def test_struct_layout_matches_c():
  # 6 int32 + 2 double + 1 int64 + 20 pointers, natural alignment, no padding surprises
  assert C.sizeof(_lib.QcdBatch) == 6 * 4 + 2 * 8 + 8 + 20 * 8
  assert _lib.QcdBatch.state.offset == 48 and _lib.QcdBatch.obs_stride.offset == 40
  assert C.sizeof(_lib.QcdTapeOut) == 9 * 8


def test_argument_validation_without_gpu(lib):
  b = _lib.QcdBatch()
  assert lib.qcd_step(None, None, 0, None) == -1
  b.n_envs, b.n_qubits, b.objective, b.max_depth = 4, 3, 7, 10
  assert lib.qcd_step(C.byref(b), None, 0, None) == -3            # unknown objective
  b.objective = _lib.OBJ_UC; b.n_qubits = 7
  assert lib.qcd_step(C.byref(b), None, 0, None) == -2            # UC supports eta <= 6
  b.n_qubits = 3; b.max_depth = 300
  assert lib.qcd_step(C.byref(b), None, 0, None) == -2            # uint8 depth counters
  b.max_depth = 30
  assert lib.qcd_step(C.byref(b), None, 0, None) == -1            # NULL buffers
  assert lib.qcd_reset(C.byref(b), None, None) == -1
  assert lib.qcd_replay(C.byref(b), None, 0, 0, None, None) == -1
  assert lib.qcd_step(C.byref(b), None, 5, None) == -3            # unknown action dtype


def test_plugin_registration_and_alias_package():
  import qcd_gym
  from qcd_b200.gym_compat import registry
  qcd_gym.register_envs(); qcd_gym.register_envs()
  assert "CircuitDesigner-v0" in registry
  spec = registry["CircuitDesigner-v0"]
  assert getattr(spec, "entry_point") == "qcd_b200.env:CircuitDesigner"


def test_product_does_not_import_oracle():
  pkg = os.path.join(ROOT, "qcd_b200")
  for fn in os.listdir(pkg):
    if fn.endswith(".py"):
      assert "oracle" not in open(os.path.join(pkg, fn)).read().replace("no CPU fallback", ""), fn


def test_missing_cuda_fails_loudly():
  import torch
  if torch.cuda.is_available():
    pytest.skip("CUDA present")
  from qcd_b200.batch import EnvBatch
  with pytest.raises(RuntimeError, match="no CPU fallback"):
    EnvBatch(2, 2, 12, 'SP-bell', np.array([1, 0, 0, 1]) / np.sqrt(2))


def test_action_space_bounds_match_reference():
  from qcd_b200.vector_env import action_space_for
  from oracle.qcd_oracle import action_bounds
  for eta in (1, 2, 5, 12):
    sp = action_space_for(eta); low, high = action_bounds(eta)
    assert sp.low.dtype == np.float32
    np.testing.assert_array_equal(sp.low, low); np.testing.assert_array_equal(sp.high, high)
    sp.seed(3); a = sp.sample()
    assert a.dtype == np.float32 and sp.contains(a)
    assert np.floor(a[0]) in (0, 1, 2) and 0 <= np.floor(a[1]) < eta


def test_task_string_grammar():
  from qcd_b200 import parse_task
  assert parse_task('SP-ghz3-q3-d15') == {'id': 'CircuitDesigner-v0', 'max_qubits': 3, 'max_depth': 15, 'objective': 'SP-ghz3'}
  assert parse_task('UC-random-q6-d24')['objective'] == 'UC-random'
