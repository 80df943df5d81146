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


@pytest.mark.gpu
def test_native_arm_json_line_on_gpu(cuda_device):
  """The default workload with a short run: ONE JSON line carrying the contract's keys, a measured
  roofline against MEASURED_PEAKS.json, the CPU baseline, the end-to-end number and sane clocks."""
  res = _run("--steps", "40", "--warmup", "5", "--cpu-seconds", "1", "--no-sweep", "--e2e-steps", "10", timeout=900)
  assert res.returncode == 0, res.stderr[-2000:]
  lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
  assert len(lines) == 1
  d = json.loads(lines[0])
  for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
    assert key in d, key
  assert d["n_gpus"] == 1 and d["steps"] == 40 and d["scaling"] == "weak"
  assert d["gpu_launches"] == 40 + 1          # the timed steps + the statistics fold that closes the interval
  assert d["vs_baseline"] is None and d["data"] == "synthetic" and "SP-ghz5" in d["config"]["workload"]
  assert "model" not in d["config"] and d["config"]["cuda_graph"] is True
  r = d["roofline"]
  assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
  assert 0.2 < r["frac"] < 1.3 and r["traffic"] is None or r["traffic"] > 0
  assert d["value"] > 1e9 and abs(d["value"] - 65536 / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
  e = d["e2e"]
  assert 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] == 65536 * 16 and e["d2h_bytes_per_step"] > 65536 * 10
  c = d["cpu_baseline"]
  assert c["kind"] == "port" and c["cores"] == 1 and 0 < c["value"] < 1e6
  # the other BASELINE.json configs ride in the same line, each with its own roofline
  cf = d["configs"]
  assert set(cf) == {"uc-toffoli", "sp-random12-1Mi", "uc-random6-replay", "sp-bell-single-env"}
  assert cf["uc-toffoli"]["roofline"]["bound"] == "hbm" and 0.2 < cf["uc-toffoli"]["roofline"]["frac"] < 1.3
  if "skipped" not in cf["sp-random12-1Mi"]:
    assert cf["sp-random12-1Mi"]["scaling"] == "strong" and cf["sp-random12-1Mi"]["envs_total"] == 1 << 20
    assert 0.5 < cf["sp-random12-1Mi"]["roofline"]["frac"] < 1.3
  rp = cf["uc-random6-replay"]
  assert rp["roofline"]["bound"] == "fp64" and rp["roofline"]["unit"] == "TFLOP/s" and 0 < rp["roofline"]["frac"] < 1
  assert rp["envs_total"] == 262144 and rp["value"] > 1e8
  assert cf["sp-bell-single-env"]["value"] > cf["sp-bell-single-env"]["cpu_baseline"]["value"]
  assert d["clocks"]["sm_max_mhz"] and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_bench_cost_models():
  """The two analytic models bench.py reports beside the measured numbers: algorithmic FP64 flops of a tape (SURVEY 8d:
  RX 6S, P 3S, CP 1.5S, CX 0, 8S per metric evaluation) and the 32-byte-sector model of the bytes a step really moves."""
  import numpy as np
  sys.path.insert(0, ROOT)
  import bench
  S = 64
  #           P(q0)            RX(q1)           CP(c=0,w=2)      CX(c=1,w=0)      Terminate
  tape = np.array([[[0.2, 0.5, 0.9, 1.0]], [[1.1, 1.2, 1.7, 1.0]], [[0.3, 2.1, 0.4, 1.0]], [[1.9, 0.2, 1.5, 1.0]],
                   [[2.5, 0.0, 0.0, 0.0]]], dtype=np.float32)
  flops = bench.tape_flops_per_env_step(tape, 3, S, metric_evals_per_env_step=0.2)
  assert abs(flops - ((3 + 6 + 1.5 + 0 + 0) * S / 5 + 8 * S * 0.2)) < 1e-9
  # SP, eta=3, S=8: state read 128 + obs 64 + 150 scalars + state writes: P on bit 0 -> all sectors (128), RX all
  # (128), CP with bit 0 involved -> half (64), CX with control bit 1 -> half (64), Terminate + reset -> all (128)
  got = bench.modelled_dram_bytes_per_env_step(tape, 3, 8, False, 64)
  assert abs(got - (128 + 64 + 150 + (128 + 128 + 64 + 64 + 128) / 5)) < 1e-9
  assert got < 32 * 8 + 8 * 8 + 96 + 150          # below the algorithmic figure + scalars: stores are skipped
