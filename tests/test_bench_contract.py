"""bench.py contract checks that run without a GPU: the reference arm prints ONE JSON line with the
driver's keys, and the native arm refuses to run without CUDA instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=300):
  return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                        timeout=timeout, cwd=ROOT)


def test_reference_arm_json_line():
  res = _run("--impl", "reference", "--steps", "3", "--warmup", "3", "--ref-envs-per-core", "4")
  assert res.returncode == 0, res.stderr[-2000:]
  lines = [l for l in res.stdout.splitlines() if l.strip()]
  assert len(lines) == 1
  d = json.loads(lines[0])
  assert d["impl"] == "reference" and d["metric"] == "batched env-steps/sec" and d["unit"] == "env-steps/s"
  assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["steps"] == 3 and d["warmup"] == 3
  assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == os.cpu_count()
  assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
  assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["gpu_launches"] == 0
  assert "SP-ghz5" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_silently():
  env = dict(os.environ, RANK="1", WORLD_SIZE="2")
  res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
  assert res.returncode == 0 and res.stdout.strip() == ""


def test_native_arm_needs_cuda():
  import torch
  if torch.cuda.is_available():
    pytest.skip("CUDA present")
  res = _run("--steps", "3", "--warmup", "3", "--no-cpu")
  assert res.returncode != 0 and "no CPU fallback" in (res.stderr + res.stdout)
