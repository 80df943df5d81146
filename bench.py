#!/usr/bin/env python
"""bench.py — batched env-steps/s of the CircuitDesigner hot path on B200, next to the CPU path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload sp-ghz5]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE launch of the fused step kernel over one batch of `envs` environments (one
`env.step()` of every env: decode, gate, depth, metric, cost, reward, float32 observation,
auto-reset, episode statistics).  Default workload = BASELINE.json configs[1]: SP-ghz5, 65 536 envs
per GPU, max_depth 20, actions drawn like the reference's own sampler (Box.sample(),
plot/metrics.py:174).  Weak scaling: every rank owns `envs` envs; the only collective is one
all-reduce of the 96-byte episode-statistics vector at the end of the timed region.

L2 policy: the batch (57 MB at the default) fits the 126 MB L2, so the bench holds R independent
replicas of the batch (total footprint >= 3x L2) and steps them round-robin: every launch finds its
inputs in HBM, not in L2 ("inputs larger than L2").

One JSON line on stdout (rank 0).  `value` = device-timed whole-job throughput with inputs resident
in HBM; `e2e` = the same metric through CircuitDesignerVectorEnv.step_host (pinned host actions in,
host reward/terminated/truncated/observations out, every step); `roofline` = algorithmic bytes per
launch / average launch time vs the measured HBM copy peak; `cpu_baseline` = the numpy oracle port of
the reference's per-step algorithm timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (objective, eta, delta, envs per GPU, mode)
    'sp-bell': ('SP-bell', 2, 12, 65536, 'step'),
    'sp-ghz5': ('SP-ghz5', 5, 20, 65536, 'step'),                 # BASELINE.json configs[1] (headline)
    'uc-toffoli': ('UC-toffoli', 3, 30, 65536, 'step'),           # configs[2]
    'sp-random12': ('SP-random', 12, 36, 16384, 'step'),          # configs[3] shape (N scaled to --envs)
    'uc-random6': ('UC-random', 6, 24, 16384, 'step'),            # configs[4] shape, step kernel
    'sp-ghz3': ('SP-ghz3', 3, 15, 1048576, 'step'),
    'sp-random1': ('SP-random', 1, 6, 1048576, 'step'),
    'sp-random4': ('SP-random', 4, 12, 1048576, 'step'),
    'uc-hadamard1': ('UC-hadamard', 1, 9, 1048576, 'step'),
    'uc-hadamard2': ('UC-hadamard', 2, 9, 1048576, 'step'),
    'sp-random6': ('SP-random', 6, 18, 524288, 'step'),
    'sp-random7': ('SP-random', 7, 21, 524288, 'step'),
    'sp-random8': ('SP-random', 8, 24, 262144, 'step'),
    'sp-random9': ('SP-random', 9, 27, 131072, 'step'),
    'sp-random10': ('SP-random', 10, 30, 65536, 'step'),
    'sp-random11': ('SP-random', 11, 33, 32768, 'step'),
    'uc-random4': ('UC-random', 4, 16, 262144, 'step'),
    'uc-random5': ('UC-random', 5, 20, 65536, 'step'),
    'uc-random6-replay': ('UC-random', 6, 24, 4096, 'replay'),    # configs[4]: replay kernel, T=64
    'sp-random12-replay': ('SP-random', 12, 36, 4096, 'replay'),
    'sp-ghz5-replay': ('SP-ghz5', 5, 20, 65536, 'replay'),
}
METRIC = "batched env-steps/sec"
UNIT = "env-steps/s"
L2_BYTES = 126e6
TAPE_LEN = 16
REPLAY_T = 64


# ------------------------------------------------------------------------------------------------
def load_peaks():
  path = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(path):
    try:
      return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
      pass
  return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
  """NVML sampler thread: SM clock + throttle reasons DURING the timed region."""

  def __init__(self, index: int, period_s: float = 0.01):
    self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
    self._stop = threading.Event()
    self._thread = None
    self.period = period_s
    try:
      import pynvml
      pynvml.nvmlInit()
      self.nv = pynvml
      self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
      self.ok = True
    except Exception as exc:            # no NVML: report that instead of inventing numbers
      self.err = repr(exc)

  def _names(self, mask: int):
    nv = self.nv
    table = {'sw_power_cap': 'nvmlClocksThrottleReasonSwPowerCap', 'hw_slowdown': 'nvmlClocksThrottleReasonHwSlowdown',
             'hw_thermal_slowdown': 'nvmlClocksThrottleReasonHwThermalSlowdown',
             'sw_thermal_slowdown': 'nvmlClocksThrottleReasonSwThermalSlowdown',
             'hw_power_brake': 'nvmlClocksThrottleReasonHwPowerBrakeSlowdown',
             'sync_boost': 'nvmlClocksThrottleReasonSyncBoost',
             'app_clocks': 'nvmlClocksThrottleReasonApplicationsClocksSetting'}
    return {k for k, attr in table.items() if mask & getattr(nv, attr, 0)}

  def sample(self):
    if not self.ok:
      return
    try:
      self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
      self.reasons |= self._names(int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)))
    except Exception:
      pass

  def _run(self):
    while not self._stop.is_set():
      self.sample()
      self._stop.wait(self.period)

  def start(self):
    if self.ok:
      self._thread = threading.Thread(target=self._run, daemon=True)
      self._thread.start()

  def stop(self):
    self.sample()
    self._stop.set()
    if self._thread is not None:
      self._thread.join()

  def summary(self):
    if not self.ok or not self.samples:
      return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
    return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
def make_tape(n_qubits: int, T: int, N: int, seed: int, no_terminate: bool = False):
  """[T,N,4] float32 actions ~ Box.sample() of env.py:46-49 (uniform in the float32 bounds).
  no_terminate: gate index drawn from [0,2) so that episodes run into the depth truncation."""
  m = 1e-5
  low = np.array([0, 0, 0, -np.pi]).astype(np.float32)
  high = np.array([(2 if no_terminate else 3) - m, n_qubits - m, n_qubits - m, np.pi]).astype(np.float32)
  rng = np.random.default_rng(seed)
  return rng.uniform(low=low, high=high, size=(T, N, 4)).astype(np.float32)


# ------------------------------------------------------------------------------------------------
def cpu_port_steps_per_s(objective, eta, delta, tape, budget_s: float, max_envs: int = 64):
  """Time the numpy oracle port (reference per-step algorithm: re-simulate the circuit for the
  observation and again for the reward, walk it for depth) on ONE core for ~budget_s seconds."""
  from oracle.qcd_oracle import OracleCircuitDesigner, make_target
  target = make_target(objective, eta, seed=42)
  T, N = tape.shape[0], min(tape.shape[1], max_envs)
  envs = [OracleCircuitDesigner(eta, delta, objective, target=target) for _ in range(N)]
  for e in envs:
    e.reset()
  steps, t0, t = 0, time.perf_counter(), 0
  while True:
    for i, e in enumerate(envs):
      _, _, term, trunc, _ = e.step(tape[t % T, i])
      if term or trunc:
        e.reset()
    steps += N
    t += 1
    if time.perf_counter() - t0 >= budget_s:
      break
  dt = time.perf_counter() - t0
  return steps / dt, steps, dt


def _ref_worker(args):
  objective, eta, delta, tape, n_steps = args
  from oracle.qcd_oracle import OracleCircuitDesigner, make_target
  target = make_target(objective, eta, seed=42)
  N = tape.shape[1]
  envs = [OracleCircuitDesigner(eta, delta, objective, target=target) for _ in range(N)]
  for e in envs:
    e.reset()
  # state persists in the worker between calls through this closure-free global
  global _REF_ENVS
  _REF_ENVS = envs
  return N


def _ref_step(args):
  tape_row, = args
  for i, e in enumerate(_REF_ENVS):
    _, _, term, trunc, _ = e.step(tape_row[i])
    if term or trunc:
      e.reset()
  return len(_REF_ENVS)


def run_reference(args):
  """--impl reference: the reference's CPU implementation of the path (oracle port; qiskit itself is
  not installable here) on all host cores.  One step = one env.step() of a bounded sample of the
  workload's envs, spread over one worker process per core."""
  import multiprocessing as mp
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return 0
  objective, eta, delta, envs, mode = WORKLOADS[args.workload]
  cores = os.cpu_count() or 1
  per_worker = args.ref_envs_per_core
  tape = make_tape(eta, TAPE_LEN, per_worker * cores, seed=1234)
  ctx = mp.get_context("fork")
  # one single-process pool per core so that each worker keeps its own envs between steps
  pools = [ctx.Pool(1) for _ in range(cores)]
  try:
    for w, p in enumerate(pools):
      p.apply(_ref_worker, ((objective, eta, delta, tape[:, w * per_worker:(w + 1) * per_worker], 0),))

    def one_step(t):
      res = [p.apply_async(_ref_step, ((tape[t % TAPE_LEN, w * per_worker:(w + 1) * per_worker],),))
             for w, p in enumerate(pools)]
      return sum(r.get() for r in res)

    for t in range(args.warmup):
      one_step(t)
    t0 = time.perf_counter()
    total = 0
    for t in range(args.steps):
      total += one_step(args.warmup + t)
    dt = time.perf_counter() - t0
  finally:
    for p in pools:
      p.terminate()
  value = total / dt
  sample = (f"{per_worker * cores} of {envs} envs per step ({per_worker} per core), same Box.sample() "
            f"action distribution, numpy restatement of env.py on qiskit-1.0.2 semantics")
  line = {
      "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
      "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex128",
      "data": "synthetic",
      "config": {"workload": f"{objective} eta={eta} max_depth={delta} envs={envs} (reference arm: "
                             f"{per_worker * cores}-env sample per step)", "actions": "uniform-box"},
      "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                       "per_core": value / cores,
                       "per_core_note": "one process alone reaches ~14 k env-steps/s on this workload (the `cpu_baseline` "
                                        "of the native arm); with one worker per logical core the per-core rate roughly "
                                        "halves (SMT siblings, one IPC round trip per step), so this arm understates the "
                                        "CPU by up to 2x"},
      "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
      "gpu_launches": 0,
      "note": "qiskit==1.0.2/gymnasium==0.29 are not installable in this image; the timed code is "
              "oracle/qcd_oracle.py, which follows the reference's per-step algorithm",
  }
  print(json.dumps(line), flush=True)
  return 0



# ------------------------------------------------------------------------------------------------
FP64_PEAK_FILE = os.path.join(ROOT, "MEASURED_PEAKS_FP64.json")       # committed copy of the gpurun measurement


def load_fp64_peak():
  """Measured DFMA peak of this pool's B200 (scripts/microbench_peaks.cu run under gpurun, committed)."""
  try:
    return float(json.load(open(FP64_PEAK_FILE))["fp64_dfma_tflops"]), "measured (MEASURED_PEAKS_FP64.json: scripts/microbench_peaks.cu on this pool's B200)"
  except Exception:
    return 37.1, "fallback (64 DFMA/clk/SM x 148 SMs x 1.965 GHz)"


def tape_flops_per_env_step(tape_np, eta, S, metric_evals_per_env_step):
  """Algorithmic FP64 flops per env-step of a tape: RX 6S, P 3S, CP 1.5S, CX 0 per gate,
  8S per metric evaluation."""
  g, w, c = (np.floor(tape_np[..., i]).astype(np.int64) for i in range(3))
  same = w == c
  per_gate = np.where(g == 0, np.where(same, 3.0, 1.5), np.where(g == 1, np.where(same, 6.0, 0.0), 0.0)) * S
  return float(per_gate.mean() + 8.0 * S * metric_evals_per_env_step)


def run_configs(args, torch, dist, dev, rank, world, stream):
  """Short runs of BASELINE.json configs[0], [2], [3], [4]; each entry carries its own roofline and CPU port
  rate.  configs[3] and [4] name a TOTAL number of envs, so they are strong-scaled (total / world per GPU)."""
  from qcd_b200.batch import EnvBatch
  from qcd_b200.targets import make_target
  peak, peak_src = load_peaks()
  out = {}

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize(dev)

  def timed_ms(launch, k, w):
    with torch.cuda.stream(stream):
      launch(0)                                   # (lazy loading / allocation before the barrier)
      stream.synchronize()
      barrier()
      for i in range(w):                          # warm-up right before the timed launches (see run_native)
        launch(i)
      stream.synchronize()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record(stream)
      for i in range(k):
        launch(w + i)
      e1.record(stream)
      stream.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

  def cpu(objective, eta, delta, tape_np, seconds):
    if args.no_cpu or rank != 0:
      return None
    v, n, dt = cpu_port_steps_per_s(objective, eta, delta, tape_np, seconds, max_envs=16)
    return {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"first 16 envs, {n} env-steps in {dt:.1f} s, same action distribution (numpy restatement of env.py)"}

  # ---- configs[2]: UC-toffoli, 65 536 envs per GPU, step kernel (weak scaling like the headline)
  objective, eta, delta, N, _ = WORKLOADS['uc-toffoli']
  target = make_target(objective, eta, seed=42)
  S = 4 ** eta
  Rr = 4                                                     # 4 x 108 MB of replicas > L2
  bs = [EnvBatch(N, eta, delta, objective, target, device=dev) for _ in range(Rr)]
  tape_np = make_tape(eta, TAPE_LEN, N, seed=4321 + rank)
  tape = torch.as_tensor(tape_np, device=dev)
  k3 = 400
  ms = timed_ms(lambda i: bs[i % Rr].step(tape[i % TAPE_LEN]), k3, 40)
  b_env = bs[0].algorithmic_bytes_per_step()
  ach = b_env * N / (ms / k3 * 1e-3) / 1e9
  out["uc-toffoli"] = {
      "baseline_config": "configs[2]: UC-toffoli 3-qubit unitary composition, 65536 envs, max_depth=30",
      "value": N * k3 * world / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "scaling": "weak", "steps": k3,
      "ms_per_step": ms / k3, "envs_per_gpu": N, "l2": f"{Rr} batch replicas round-robin (eager launches, PDL)",
      "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                   "kernel": "qcd::step_warp_kernel<log2S=6,UC>", "algorithmic_bytes_per_env_step": b_env,
                   "modelled_dram_bytes_per_env_step": modelled_dram_bytes_per_env_step(tape_np, eta, S, True, 8 * S),
                   "peak_source": peak_src},
      "cpu_baseline": cpu(objective, eta, delta, tape_np, 3.0)}
  del bs, tape
  torch.cuda.empty_cache()

  # ---- configs[3]: SP-random eta=12, 1 Mi envs IN TOTAL sharded over the GPUs, punish=True, state-half observations
  objective, eta, delta = 'SP-random', 12, 36
  total = 1 << 20
  N = total // world
  S = 2 ** eta
  free, _ = torch.cuda.mem_get_info(dev)
  need = N * (16 * S + 8 * S + 256)
  if free < need * 1.05:
    out["sp-random12-1Mi"] = {"skipped": f"needs {need / 2**30:.0f} GiB per GPU, {free / 2**30:.0f} GiB free"}
  else:
    target = make_target(objective, eta, seed=42)
    bt = EnvBatch(N, eta, delta, objective, target, punish=True, obs_mode='state', device=dev)
    tape_np = make_tape(eta, 2, N, seed=777 + rank)
    tape = torch.as_tensor(tape_np, device=dev)
    k4 = 6
    ms = timed_ms(lambda i: bt.step(tape[i % 2]), k4, 3)
    b_env = bt.algorithmic_bytes_per_step()
    ach = b_env * N / (ms / k4 * 1e-3) / 1e9
    out["sp-random12-1Mi"] = {
        "baseline_config": "configs[3]: SP-random max_qubits=12, 1M envs sharded over 1/2/4/8 B200, punish=True",
        "value": total * k4 / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "scaling": "strong", "steps": k4,
        "ms_per_step": ms / k4, "envs_total": total, "envs_per_gpu": N, "max_depth": delta, "observations": "state",
        "resident_gib_per_gpu": need / 2**30, "l2": "inputs (64 GiB / world of states) far larger than L2",
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                     "kernel": "qcd::qcd_kernel<log2S=12,SP> (pair-streaming step)", "algorithmic_bytes_per_env_step": b_env,
                     "modelled_dram_bytes_per_env_step": modelled_dram_bytes_per_env_step(tape_np, eta, S, False, 8 * S),
                     "frac_of_modelled_traffic": modelled_dram_bytes_per_env_step(tape_np, eta, S, False, 8 * S) / b_env * ach / peak,
                     "peak_source": peak_src,
                     "note": "algorithmic bytes count a full state write; stores of amplitudes a gate leaves untouched "
                             "are skipped (ncu r01: real DRAM traffic = 0.83 x algorithmic), so frac can exceed 1"},
        "cpu_baseline": cpu(objective, eta, delta, tape_np[:, :16], 3.0)}
    del bt, tape
    torch.cuda.empty_cache()

  # ---- configs[4]: UC-random eta=6, 262 144 envs IN TOTAL, replay kernel, T = 64
  objective, eta, delta = 'UC-random', 6, 24
  total = 262144
  N = total // world
  S = 4 ** eta
  target = make_target(objective, eta, seed=42)
  br = EnvBatch(N, eta, delta, objective, target, obs_mode='state', device=dev)
  tape_np = make_tape(eta, 2 * REPLAY_T, N, seed=999 + rank)
  tape = torch.as_tensor(tape_np, device=dev)
  k5 = 4
  ms = timed_ms(lambda i: br.replay(tape[(i % 2) * REPLAY_T:(i % 2 + 1) * REPLAY_T]), k5, 2)
  # metric evaluations of the last launch: finished episodes + the last step of envs that did not finish on it
  br.zero_stats()
  with torch.cuda.stream(stream):
    br.replay(tape[:REPLAY_T])
    stream.synchronize()
  episodes = float(br.stats[0].item())
  done_last = float(((br.terminated != 0) | (br.truncated != 0)).sum().item())
  evals = (episodes + N - done_last) / (N * REPLAY_T)
  flops = tape_flops_per_env_step(tape_np[:REPLAY_T], eta, S, evals)
  rate = total * k5 * REPLAY_T / (ms * 1e-3)
  fpeak, fsrc = load_fp64_peak()
  ach = rate / world * flops / 1e12
  out["uc-random6-replay"] = {
      "baseline_config": "configs[4]: UC-random max_qubits=6 (64x64 unitary), 262144 envs, sequence-replay kernel",
      "value": rate, "unit": UNIT, "n_gpus": world, "scaling": "strong", "steps": k5, "ms_per_step": ms / k5,
      "envs_total": total, "envs_per_gpu": N, "steps_per_launch": REPLAY_T, "max_depth": delta,
      "actions": "uniform-box (mean episode = 3 steps: a metric evaluation on every third step)",
      "roofline": {"bound": "fp64", "achieved": ach, "peak": fpeak, "unit": "TFLOP/s", "frac": ach / fpeak, "traffic": None,
                   "kernel": "qcd::replay_reg_kernel<log2S=12,UC>", "algorithmic_flops_per_env_step": flops,
                   "metric_evaluations_per_env_step": evals, "peak_source": fsrc,
                   "note": "state register-resident for the whole tape: HBM traffic is one state in/out per launch (3 % of "
                           "the copy peak); FP64 is the only roofline left.  ncu (profiles/r02_summary.md): FP64 pipe 19 % "
                           "active, issue slots 51 %, shared-memory/shuffle pipe 49 % - bound by per-step latency, the gate "
                           "dispatch and the shared-memory traffic of the metric (64 KiB of target per evaluation), not by "
                           "DFMA throughput"},
      "cpu_baseline": cpu(objective, eta, delta, tape_np[:, :16], 3.0)}
  del br, tape
  torch.cuda.empty_cache()

  # ---- configs[0]: SP-bell, max_qubits=2, max_depth=12, ONE env through gym.make, random actions, reset on done
  if rank == 0:
    import qcd_gym
    from qcd_b200 import gym_compat as gym
    from oracle.qcd_oracle import OracleCircuitDesigner, sample_actions
    qcd_gym.register_envs()
    rng = np.random.default_rng(0)
    acts = [sample_actions(rng, 2) for _ in range(4096)]
    res = {}
    for name, env in (("gpu", gym.make("CircuitDesigner-v0", max_qubits=2, max_depth=12, objective='SP-bell')),
                      ("cpu_port", OracleCircuitDesigner(2, 12, 'SP-bell'))):
      if name == "cpu_port" and args.no_cpu:
        continue
      env.reset()
      for n in (2000, 20000):                                 # warm-up, then timed
        t0 = time.perf_counter()
        for i in range(n):
          _, _, term, trunc, _ = env.step(acts[i & 4095])
          if term or trunc:
            env.reset()
        dt = time.perf_counter() - t0
      res[name] = n / dt
    out["sp-bell-single-env"] = {
        "baseline_config": "configs[0]: SP-bell, max_qubits=2, max_depth=12, single env, random actions",
        "value": res["gpu"], "unit": "env-steps/s", "n_gpus": 1, "us_per_step": 1e6 / res["gpu"], "steps": 20000,
        "api": "gym.make('CircuitDesigner-v0').step(action): numpy in / numpy out, one qcd_step1_sync call per step "
               "(action as kernel argument, results written by the kernel into mapped pinned host memory)",
        "cpu_baseline": None if "cpu_port" not in res else
        {"value": res["cpu_port"], "unit": "env-steps/s", "cores": 1, "kind": "port",
         "sample": "the same 20000 steps through oracle/qcd_oracle.py OracleCircuitDesigner"}}
  barrier()
  return out


def modelled_dram_bytes_per_env_step(tape_np, eta, S, uc, obs_bytes):
  """Bytes one env-step really moves to/from DRAM under a 32-byte-sector model of the step kernels, which only store
  the amplitudes a gate changed: state read 16S, state write = the sectors holding changed amplitudes (RX all, P / CX
  half or all depending on whether the bit is bit 0 of the flat index, CP a quarter or half, Terminate + auto-reset
  all), the float32 state half of the observation, and 150 B of scalars.  A model, not a measurement (truncation
  resets are not counted); the ncu figures are in profiles/."""
  g, w, c = (np.floor(tape_np[..., i]).astype(np.int64) for i in range(3))
  shift = eta if uc else 0
  wb, cb = w + shift, c + shift
  same = w == c
  full, half, quarter = 16.0 * S, 8.0 * S, 4.0 * S
  p_w = np.where(wb == 0, full, half)
  cp_w = np.where(np.minimum(wb, cb) == 0, half, quarter)
  cx_w = np.where(cb == 0, full, half)
  state_w = np.where(g == 2, full, np.where(g == 0, np.where(same, p_w, cp_w), np.where(same, full, cx_w)))
  return float(16.0 * S + state_w.mean() + obs_bytes + 150.0)


# ------------------------------------------------------------------------------------------------
def run_native(args):
  import torch
  import torch.distributed as dist
  from qcd_b200.batch import EnvBatch
  from qcd_b200.sharding import all_reduce_stats, summarize
  from qcd_b200.targets import make_target
  from qcd_b200.vector_env import CircuitDesignerVectorEnv

  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  if not torch.cuda.is_available():
    raise SystemExit("bench.py needs a CUDA device: qcd_b200 has no CPU fallback")
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  if world != args.gpus and rank == 0:
    print(f"# note: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

  objective, eta, delta, envs, mode = WORKLOADS[args.workload]
  N = args.envs or envs
  target = make_target(objective, eta, seed=42)
  K, W = args.steps, args.warmup
  obs_mode = args.obs

  # ---- R replicas of the batch so that consecutive launches never find their data in L2
  from qcd_b200 import _lib as qlib
  shared_stats = torch.zeros((qlib.STATS_COPIES, qlib.STATS_STRIDE), dtype=torch.float64, device=dev)
  probe = EnvBatch(N, eta, delta, objective, target, obs_mode=obs_mode, final_obs=args.final_obs,
                   fused_stats=not args.no_stats, device=dev, stats_buffer=shared_stats,
                   actions_stable=args.actions_stable)
  S = probe.S
  bytes_step_env = probe.algorithmic_bytes_per_step()
  footprint = N * (16 * S + (8 * S if obs_mode != 'none' else 0) + 96)
  R = max(2, int(np.ceil(3 * L2_BYTES / footprint))) if not args.no_rotate else 1
  free, _ = torch.cuda.mem_get_info(dev)
  per_replica = N * (16 * S + (4 * probe.obs_stride if obs_mode != 'none' else 0) * (2 if args.final_obs else 1) + 200)
  R = max(1, min(R, int(0.8 * free // per_replica)))
  batches = [probe] + [EnvBatch(N, eta, delta, objective, target, obs_mode=obs_mode, final_obs=args.final_obs,
                                fused_stats=not args.no_stats, device=dev, stats_buffer=shared_stats,
                                actions_stable=args.actions_stable)
                       for _ in range(R - 1)]
  tape_np = make_tape(eta, TAPE_LEN if mode == 'step' else REPLAY_T * 2, N, seed=1234 + rank,
                      no_terminate=args.no_terminate)
  tape = torch.as_tensor(tape_np, device=dev)

  if mode == 'step':
    steps_per_launch = 1

    def launch(i):
      batches[i % R].step(tape[i % TAPE_LEN])
  else:
    steps_per_launch = REPLAY_T

    def launch(i):
      batches[i % R].replay(tape[(i % 2) * REPLAY_T:(i % 2 + 1) * REPLAY_T])

  # ---- CUDA graphs of consecutive launches (launch-bound otherwise: ~20 us kernels).  One graph of G
  # launches (a whole number of replica/tape periods) plus, for step counts that are not a multiple
  # of G, one graph per remainder, so that EVERY timed launch is replayed from a graph.
  G = int(np.lcm(R, TAPE_LEN)) if mode == 'step' else R * 2
  stream = torch.cuda.Stream(dev)
  side = torch.cuda.Stream(dev)
  stat_buf = [torch.zeros(qlib.STATS_STRIDE, dtype=torch.float64, device=dev) for _ in range(2)]
  graphs = {}

  def capture(n, close):
    """n consecutive launches; close=True appends the statistics fold that ends a reporting interval."""
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
      for i in range(n):
        launch(i)
      if close:
        probe.fold_stats(out=stat_buf[1], zero=True)
    return g

  def plan(k):
    """[(graph length, closes the interval)] covering k launches; the last graph also folds the statistics."""
    lens = [G] * (k // G) + ([k % G] if k % G else [])
    return [(n, i == len(lens) - 1) for i, n in enumerate(lens)]

  with torch.cuda.stream(stream):
    for i in range(max(W, 3)):
      launch(i)
    probe.fold_stats(out=stat_buf[1], zero=True)     # (lazy module loading outside of any capture)
    stream.synchronize()
    if not args.no_graph:
      for key in set(plan(K)) | set(plan(W)):
        graphs[key] = capture(*key)
        graphs[key].replay()
      stream.synchronize()
  graph = next(iter(graphs.values())) if graphs else None

  def run_steps(k):
    """k launches + the closing statistics fold, every kernel replayed from a CUDA graph; returns the graph lengths."""
    if not graphs:
      for i in range(k):
        launch(i)
      probe.fold_stats(out=stat_buf[1], zero=True)
      return [0] * k
    for key in plan(k):
      graphs[key].replay()
    return [n for n, _ in plan(k)]

  # ---- the one collective of the path: the episode statistics of a reporting interval (96 B per rank), exchanged
  # over NCCL/NVLink on a SIDE stream, off the stepping stream's critical path (the reference only logs these
  # numbers, algorithm.py:99-114).  Every interval ends with one small fold kernel (the 64 atomic-spreading
  # replicas -> one vector, replicas cleared), captured as the last node of the interval's last graph.  Interval
  # 0 = the warm-up, interval 1 = the K timed steps.  By default both vectors are all-reduced on the side stream
  # right after the timed region; with --exchange overlap interval 0's all-reduce is in flight DURING the timed
  # region (what a long-running trainer gets).  Measured on 8 GPUs: the concurrent NCCL kernel costs 2-3 of the 8
  # ranks ~35 us once (bimodal per-rank times 0.372 / 0.408 ms), i.e. nothing for an interval of thousands of steps
  # but 9 % of a 20-step region, which is why it is not the default of the 20-step protocol.
  # Order: [setup, captures, NCCL warm-up] -> barrier + synchronize -> W untimed warm-up steps -> synchronize ->
  # K timed steps -> synchronize + barrier.  The warm-up runs AFTER the barrier so that every rank enters the
  # timed region straight from its warm-up (a GPU that idles at a barrier for milliseconds waiting for slower
  # ranks starts its next kernels measurably slower; with 20 timed steps = 0.37 ms that decides the number).
  clocks = ClockSampler(local)
  with torch.cuda.stream(stream):
    with torch.cuda.stream(side):                 # warm NCCL (communicator setup is lazy)
      all_reduce_stats(torch.zeros(qlib.STATS_STRIDE, dtype=torch.float64, device=dev))
    side.synchronize()
    stream.synchronize()
    clocks.start()                                # polling before, during and after the timed region
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize(dev)
    run_steps(W)                                  # the W untimed warm-up steps (interval 0, closed by its fold)
    stat_buf[0].copy_(stat_buf[1])                # interval 0's vector
    stream.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.exchange == "overlap":
      with torch.cuda.stream(side):
        all_reduce_stats(stat_buf[0])             # in flight on the side stream during the timed region
    ev0.record(stream)
    graphs_used = run_steps(K)                    # K step launches + 1 fold, nothing else
    ev1.record(stream)
    with torch.cuda.stream(side):
      side.wait_event(ev1)
      if args.exchange != "overlap":
        all_reduce_stats(stat_buf[0])
      all_reduce_stats(stat_buf[1])
    stream.synchronize()
    side.synchronize()
    clocks.stop()
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize(dev)
  stats = stat_buf[1]
  elapsed_ms = ev0.elapsed_time(ev1)
  t_all = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
  rank_ms = [elapsed_ms]
  if world > 1:
    every = [torch.zeros_like(t_all) for _ in range(world)]
    dist.all_gather(every, t_all)
    rank_ms = [float(t.item()) for t in every]
    dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
  elapsed_ms = float(t_all.item())
  env_steps = N * K * steps_per_launch * world
  value = env_steps / (elapsed_ms * 1e-3)

  # ---- a device-resident loop as an RL trainer would run it: a (stand-in) policy writes new actions on the
  # GPU every step (torch.rand + affine map to the Box bounds, 3 small torch kernels), then the env steps;
  # no graph, no host data.  Sits between `value` (tape replayed from a graph) and `e2e` (host buffers).
  device_loop = None
  if mode == 'step' and not args.no_sweep:
    lowv = torch.tensor([0.0, 0.0, 0.0, -np.pi], device=dev, dtype=torch.float32)
    span = torch.tensor([3 - 1e-5, eta - 1e-5, eta - 1e-5, 2 * np.pi], device=dev, dtype=torch.float32)
    act = torch.empty((N, 4), device=dev, dtype=torch.float32)
    Kd = min(K, 500)
    with torch.cuda.stream(stream):
      for i in range(10):
        torch.rand((N, 4), out=act); act.mul_(span).add_(lowv); batches[i % R].step(act)
      stream.synchronize()
      if world > 1:
        dist.barrier()
      d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      d0.record(stream)
      for i in range(Kd):
        torch.rand((N, 4), out=act); act.mul_(span).add_(lowv); batches[i % R].step(act)
      d1.record(stream)
      stream.synchronize()
    td = torch.tensor([d0.elapsed_time(d1)], dtype=torch.float64, device=dev)
    if world > 1:
      dist.all_reduce(td, op=dist.ReduceOp.MAX)
    device_loop = {"value": N * Kd * world / (float(td.item()) * 1e-3), "unit": UNIT, "steps": Kd,
                   "ms_per_step": float(td.item()) / Kd,
                   "what": "torch.rand policy stand-in (3 torch kernels) + qcd_step per iteration, python-launched"}

  # ---- secondary point: the same kernel at 1 Mi envs per GPU (steady state, launch ramp amortised)
  sweep = None
  if mode == 'step' and not args.no_sweep and args.envs == 0 and S <= 64:
    del batches, probe
    torch.cuda.empty_cache()
    Nb = 1 << 20
    big = [EnvBatch(Nb, eta, delta, objective, target, obs_mode=obs_mode, final_obs=args.final_obs,
                    fused_stats=not args.no_stats, device=dev) for _ in range(2)]
    tape_b = torch.as_tensor(make_tape(eta, 4, Nb, seed=99 + rank, no_terminate=args.no_terminate), device=dev)
    Kb = 120
    with torch.cuda.stream(stream):
      for i in range(8):
        big[i % 2].step(tape_b[i % 4])
      stream.synchronize()
      if world > 1:
        dist.barrier()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record(stream)
      for i in range(Kb):
        big[i % 2].step(tape_b[i % 4])
      e1.record(stream)
      stream.synchronize()
    tb = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
      dist.all_reduce(tb, op=dist.ReduceOp.MAX)
    ms_b = float(tb.item()) / Kb
    sweep = {"envs_per_gpu": Nb, "steps": Kb, "ms_per_step": ms_b, "value": Nb * world / (ms_b * 1e-3),
             "roofline_frac": bytes_step_env * Nb / (ms_b * 1e-3) / 1e9 / load_peaks()[0],
             "l2": "2 batch replicas alternated, 1.8 GB footprint"}
    del big, tape_b
    torch.cuda.empty_cache()

  # ---- end to end through the public API with HOST buffers (step workloads only)
  e2e = None
  host_obs_bytes = N * 2 * S * 4 if obs_mode != 'none' and not args.e2e_no_obs else 0
  if mode == 'step' and (host_obs_bytes > 8e9 or args.e2e_steps <= 0):
    e2e = {"value": None, "unit": UNIT, "skipped": "pinned host observation buffer would be "
           f"{host_obs_bytes / 1e9:.0f} GB" if host_obs_bytes > 8e9 else "--e2e-steps 0"}
  elif mode == 'step':
    # The reference-facing call with HOST buffers, every step: pinned host actions -> H2D -> step kernel ->
    # D2H of reward / terminated / truncated and of the float32 state half of every observation row (the
    # target half of env.py:85 is constant: `venv.target_obs`) -> host.  The envs are stepped as G groups
    # on G streams (step_host_async / step_host_wait): a group's next actions are only submitted after its
    # previous results have arrived on the host, so the RL data dependence holds per group while group g's
    # D2H overlaps group g+1's H2D and kernel.  G = 1 is the plain synchronous step_host.
    e2e_obs = 'none' if (obs_mode == 'none' or args.e2e_no_obs) else args.e2e_obs
    venv = CircuitDesignerVectorEnv(N, eta, delta, objective, target=target, device=dev, obs_mode=e2e_obs,
                                    zero_copy_scalars=not args.e2e_copy_scalars)
    venv.reset()
    Ke = max(3, min(K, args.e2e_steps))
    G = max(1, args.e2e_groups)
    ranges = venv.host_groups(G, np.float32)
    consumed = 0.0

    def e2e_loop(n_steps):
      nonlocal consumed
      if G == 1:
        for i in range(n_steps):
          out = venv.step_host(tape_np[i % TAPE_LEN], copy_obs=e2e_obs != 'none')
          consumed += float(out[1][0])
        return
      for i in range(n_steps):
        for g, (lo, hi) in enumerate(ranges):
          if i > 0:
            out = venv.step_host_wait(g)                 # results of step i-1 of this group are on the host
            consumed += float(out[1][0])
          venv.step_host_async(g, tape_np[i % TAPE_LEN, lo:hi], copy_obs=e2e_obs != 'none')
      for g in range(G):
        out = venv.step_host_wait(g)
        consumed += float(out[1][0])

    e2e_loop(3)
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    e2e_loop(Ke)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
      dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    h2d, d2h = venv.host_bytes_per_step(np.float32, copy_obs=e2e_obs != 'none')
    e2e = {"value": N * Ke * world / float(t_e.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "steps": Ke, "ms_per_step": float(t_e.item()) / Ke * 1e3,
           "d2h_gbs_per_gpu": d2h * Ke / float(t_e.item()) / 1e9,
           "api": ("CircuitDesignerVectorEnv.step_host_async/step_host_wait -> qcd_step_host, %d env groups on %d "
                   "streams (pinned host buffers, one event wait per group and step)" % (G, G)) if G > 1 else
                  "CircuitDesignerVectorEnv.step_host -> qcd_step_host (pinned host buffers, sync per step)",
           "groups": G, "obs_copied": e2e_obs != 'none',
           "scalars": ("reward / terminated / truncated written by the step kernel straight into the pinned host arrays "
                       "(zero_copy_scalars=True; the bytes still cross PCIe and are counted in d2h_bytes_per_step)")
                      if not args.e2e_copy_scalars else "reward / terminated / truncated copied D2H",
           "obs_rows": {"state": "compact state half [N, 2S] float32, one contiguous D2H", "full": "reference rows "
                        "[N, 4S], pitched D2H of the state half", "none": "no observations"}[e2e_obs]}
    del venv

  # ---- the other BASELINE.json configs, short runs in the same job (driver-visible under `configs`)
  configs = None
  if args.workload == 'sp-ghz5' and not args.no_configs:
    try: del batches, probe, tape
    except NameError: pass
    torch.cuda.empty_cache()
    configs = run_configs(args, torch, dist, dev, rank, world, stream)

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return 0

  # ---- roofline of the dominant kernel (the only kernel in the timed region)
  peak, peak_src = load_peaks()
  launch_ms = elapsed_ms / K
  if mode == 'step':
    alg_bytes = bytes_step_env * N
  else:   # replay: state in + out once per launch, actions per step, scalars once
    alg_bytes = N * (32 * S + (8 * S if obs_mode != 'none' else 0) + 96 + 16 * REPLAY_T)
  achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
  modelled = modelled_dram_bytes_per_env_step(tape_np, eta, S, 'UC' in objective, 8 * S if obs_mode != 'none' else 0)
  # DRAM bytes cannot be measured outside the profiler: `traffic` is the steady-state ncu figure of profiles/traffic.json
  # (same workload, same env count, replicas rotating as here), else null
  traffic, traffic_src = None, None
  try:
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")) as f:
      rec = json.load(f).get(args.workload)
    if mode == 'step' and rec and rec["envs"] == N and not args.no_rotate:
      traffic = rec["bytes"]
      traffic_src = ("ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, mean of %d launches in steady state "
                     "(profiles/traffic.json, scripts/r3_traffic.sh)" % rec["launches"])
  except (OSError, ValueError, KeyError):
    pass
  roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
              "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": "qcd::%s<log2S=%d,%s>" % (
                  ("step_warp_kernel" if S <= 128 else "qcd_kernel(pair-streaming step)") if mode == 'step' else
                  ("replay_reg_kernel" if S >= 512 else "replay_warp_kernel"),
                  int(np.log2(S)), "UC" if 'UC' in objective else "SP"),
              "algorithmic_bytes_per_env_step": bytes_step_env, "launch_us": launch_ms * 1e3,
              "modelled_dram_bytes_per_env_step": modelled if mode == 'step' else None,
              "frac_of_modelled_traffic": (modelled * N / (launch_ms * 1e-3) / 1e9 / peak) if mode == 'step' else None,
              "note": ("`achieved` uses SURVEY 8(d)'s algorithmic bytes, which count a full state write; the kernels skip "
                       "stores of amplitudes a gate leaves untouched, so frac can read above 1 at large S.  "
                       "`modelled_dram_bytes_per_env_step` is what a 32-byte-sector model says really moves") if mode == 'step'
              else ("replay keeps the state in shared memory for the whole tape: HBM traffic is one state in/out "
                    "per launch, the kernel is instruction-issue bound (ncu: profiles/r01_summary.md), so the "
                    "HBM fraction is low by construction")}

  # ---- CPU baseline beside it (rank 0, one core, bounded sample)
  cpu = None
  if not args.no_cpu:
    v, n_steps, dt = cpu_port_steps_per_s(objective, eta, delta, tape_np, args.cpu_seconds)
    cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
           "sample": f"first {min(N, 64)} envs of the batch, {n_steps} env-steps in {dt:.1f} s, same action tape; "
                     f"numpy restatement of env.py (re-simulates the circuit twice per step like the reference); "
                     f"host has {os.cpu_count()} cores"}

  line = {
      "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
      "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
      "dtype": "complex128", "data": "synthetic",
      "config": {"workload": f"{objective} eta={eta} max_depth={delta} envs_per_gpu={N}", "mode": mode,
                 "envs_total": N * world, "actions": ("no-terminate" if args.no_terminate else "uniform-box") +
                 " float32 (Box.sample distribution), tape of " f"{TAPE_LEN if mode == 'step' else 2 * REPLAY_T} steps reused", "punish": True, "sparse": True, "autoreset": True,
                 "observations": obs_mode, "final_obs": bool(args.final_obs), "fused_stats": not args.no_stats,
                 "l2": f"{R} batch replicas stepped round-robin, footprint {R * footprint / 1e6:.0f} MB > L2 "
                       f"126 MB" if R > 1 else "single batch (no rotation)",
                 "cuda_graph": graph is not None, "graph_len": graphs_used if graph is not None else 0,
                 "steps_per_launch": steps_per_launch, "parallelism": f"env-sharded x{world}, stats all-reduce only",
                 "stats_exchange": args.exchange, "rank_ms": [round(t, 5) for t in rank_ms]},
      "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": K + 1, "roofline": roofline, "cpu_baseline": cpu,
      "episode_stats": summarize(stats[:qlib.STATS_LEN]), "sweep": sweep, "device_loop": device_loop,
      "configs": configs,
  }
  print(json.dumps(line), flush=True)
  if world > 1:
    dist.destroy_process_group()
  return 0


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=4000)
  ap.add_argument("--warmup", type=int, default=200)
  ap.add_argument("--impl", default="native", choices=["native", "reference"])
  ap.add_argument("--workload", default="sp-ghz5", choices=sorted(WORKLOADS))
  ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the workload's)")
  ap.add_argument("--obs", default="full", choices=["full", "state", "none"])
  ap.add_argument("--final-obs", action="store_true", help="also write terminal observations of finished envs")
  ap.add_argument("--no-graph", action="store_true")
  ap.add_argument("--exchange", default="after", choices=["overlap", "after"],
                  help="statistics all-reduce of the warm-up interval: after the timed region (default), or in flight on the "
                       "side stream during it (measured on 8 GPUs: the NCCL kernel then costs some ranks ~35 us once)")
  ap.add_argument("--actions-stable", action="store_true",
                  help="QCD_F_ACTIONS_STABLE: the tape is pre-generated, decode before griddepcontrol.wait")
  ap.add_argument("--no-rotate", action="store_true", help="single batch replica (L2-resident at small sizes)")
  ap.add_argument("--no-terminate", action="store_true", help="action tape without Terminate (episodes end by depth)")
  ap.add_argument("--no-sweep", action="store_true", help="skip the secondary 1 Mi-env measurement")
  ap.add_argument("--no-cpu", action="store_true")
  ap.add_argument("--no-configs", action="store_true", help="skip the short runs of the other BASELINE.json configs")
  ap.add_argument("--no-stats", action="store_true", help="do not accumulate episode statistics in the step kernel")
  ap.add_argument("--cpu-seconds", type=float, default=10.0)
  ap.add_argument("--e2e-steps", type=int, default=200)
  ap.add_argument("--e2e-copy-scalars", action="store_true",
                  help="e2e: D2H copies of reward / terminated / truncated instead of the kernel writing them into the "
                       "pinned host arrays itself (zero_copy_scalars=False)")
  ap.add_argument("--e2e-no-obs", action="store_true", help="e2e without the observation D2H copy")
  ap.add_argument("--e2e-groups", type=int, default=2, help="env groups in flight on the host path (1 = synchronous)")
  ap.add_argument("--e2e-obs", default="state", choices=["state", "full"],
                  help="host observation rows: compact state half (contiguous copy) or full reference rows (pitched)")
  ap.add_argument("--ref-envs-per-core", type=int, default=64)
  args = ap.parse_args()
  if args.warmup < 3:
    args.warmup = 3
  if args.impl == "reference":
    return run_reference(args)
  return run_native(args)


if __name__ == "__main__":
  sys.exit(main())
