"""Device-side drivers of the two non-RL baselines that call the hot path.

random_baseline : /root/reference/plot/metrics.py:161-176 — per seed, 100 episodes of
                  `env.step(env.action_space.sample())`, mean of the Monitor fields r,d,q,m,c: one replay launch.
circuit_fitness : /root/reference/baselines/evo/run.py:17-33 — fitness of whole circuits of
                  cx/rx/cp/p gates: fidelity | similarity minus the depth cost, evaluated with the
                  replay kernel (one launch for the whole population).
"""
from __future__ import annotations

import numpy as np
import torch

from .batch import EnvBatch
from .targets import make_target, parse_task
from .vector_env import action_space_for


def sample_actions(n: int, qubits: int, generator: torch.Generator, device) -> torch.Tensor:
  """[n,4] float32 ~ Box.sample(): uniform in the float32 bounds of env.py:46-49, on the device."""
  sp = action_space_for(qubits)
  low = torch.as_tensor(sp.low, device=device, dtype=torch.float64)
  high = torch.as_tensor(sp.high, device=device, dtype=torch.float64)
  u = torch.rand((n, 4), generator=generator, device=device, dtype=torch.float64)
  return (low + (high - low) * u).to(torch.float32)       # float32(high) < the next integer, so floor() stays in range


def random_baseline(task: str, seeds: int = 8, episodes: int = 100, env_seed: int = 42, device=None) -> dict:
  """Means of Return/Depth/Qubits/Metric/Cost per seed for `task` = '<obj>-q<eta>-d<delta>'.

  Every (seed, episode) pair is its own env: the action tapes of all seeds*episodes envs are drawn on the
  device and ONE `qcd_replay` launch (auto-reset off, per-step records) plays every episode to its end — no
  Python loop over steps, no synchronisation per step; the Monitor fields of each env are then read at its first
  finished step.  Episodes are independent, so this is the same experiment as the reference's sequential loop (the
  action streams differ from numpy's, the distribution does not).  T = max_depth*max_qubits + 1 steps suffice: every
  max_qubits gates raise the depth by at least one, and depth >= max_depth truncates (env.py:129)."""
  kw = parse_task(task)
  eta, delta, objective = kw['max_qubits'], kw['max_depth'], kw['objective']
  dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
  n = seeds * episodes
  T = delta * eta + 1                                        # depth grows by >= 1 per eta gates: truncation by then
  batch = EnvBatch(n, eta, delta, objective, make_target(objective, eta, env_seed), autoreset=False, obs_mode='none',
                   track_episodes=True, fused_stats=False, device=dev)
  gens = [torch.Generator(device=dev).manual_seed(s + 1) for s in range(seeds)]
  tape = torch.stack([torch.cat([sample_actions(episodes, eta, g, dev) for g in gens]) for _ in range(T)])   # [T,n,4]
  f64, i32, u8 = torch.float64, torch.int32, torch.uint8
  rec = {k: torch.zeros((T, n), dtype=dt, device=dev) for k, dt in
         (('reward', f64), ('metric', f64), ('cost', f64), ('depth', i32), ('used_wires', i32), ('terminated', u8), ('truncated', u8))}
  batch.replay(tape, rec)
  done = (rec['terminated'] | rec['truncated']).bool()                                  # [T,n]
  first = torch.argmax(done.to(torch.int8), dim=0)                                      # first finished step of each env
  assert bool(done.any(0).all()), "an episode outlived max_depth*max_qubits+1 steps"
  # Monitor's return = sum of the rewards up to and including the finishing step (monitor.py:35-37)
  upto = torch.arange(T, device=dev)[:, None] <= first[None, :]
  ret = (rec['reward'] * upto).sum(0)
  pick = lambda t: t.gather(0, first[None, :])[0].to(f64)
  out = {'Return': ret, 'Depth': pick(rec['depth']), 'Qubits': pick(rec['used_wires']), 'Metric': pick(rec['metric']),
         'Cost': pick(rec['cost'])}
  return {k: v.view(seeds, episodes).mean(1).cpu().numpy() for k, v in out.items()}


GATE_CODES = {'p': 0, 'cp': 0, 'rx': 1, 'cx': 1}              # env.py:95-98: gate index of each kind


def encode_circuit(gates, max_len: int) -> np.ndarray:
  """[(kind, wire, control, theta)] -> [max_len+1, 4] float64 actions, padded with Terminate (gate 2)."""
  tape = np.zeros((max_len + 1, 4), np.float64)
  tape[:, 0] = 2
  for t, (kind, wire, cntrl, theta) in enumerate(gates):
    single = kind in ('p', 'rx')
    tape[t] = (GATE_CODES[kind], wire, wire if single else cntrl, 0.0 if theta is None else theta)
  return tape


def gates_from_individual(solution, n_qubits: int | None = None, n_gates_per_qubit: int | None = None) -> list:
  """Gate list of a GA individual, in the order the reference builds its QuantumCircuit from the gate grid.

  `solution` is the individual's `n_qubits x n_gates_per_qubit` grid (or an object with `.solution`, `.n_qubits`,
  `.n_gates_per_qubit`) of gate objects with `.name` in {'p','rx','cp','cx'}, `.parameters` and `.affected_qubits`
  (cells may be None).  Mirrors `transform_to_executable_circuit`
  (/root/reference/baselines/evo/individual.py:158-201): columns outer, qubits inner; a gate is emitted where its
  cell is first met and the cells of the other qubits it touches in that column are skipped; the qiskit call is
  `getattr(qc, name)(*parameters, *affected_qubits)`, i.e. `cx(control, target)`, `cp(theta, control, target)`,
  `rx(theta, q)`, `p(theta, q)`.  Returns [(kind, wire, control, theta)] for `circuit_fitness` / `encode_circuit`."""
  if hasattr(solution, 'solution'):
    n_qubits = solution.n_qubits if n_qubits is None else n_qubits
    n_gates_per_qubit = solution.n_gates_per_qubit if n_gates_per_qubit is None else n_gates_per_qubit
    solution = solution.solution
  n_qubits = len(solution) if n_qubits is None else n_qubits
  n_gates_per_qubit = (len(solution[0]) if n_qubits else 0) if n_gates_per_qubit is None else n_gates_per_qubit
  handled = [[False] * n_gates_per_qubit for _ in range(n_qubits)]
  gates = []
  for col in range(n_gates_per_qubit):
    for q in range(n_qubits):
      gate = solution[q][col]
      if gate is None or handled[q][col]:
        continue
      qubits = [int(x) for x in gate.affected_qubits]
      params = list(gate.parameters or [])
      if gate.name in ('p', 'rx'):
        gates.append((gate.name, qubits[0], qubits[0], float(params[0])))
      elif gate.name == 'cp':
        gates.append(('cp', qubits[1], qubits[0], float(params[0])))      # cp(theta, control, target)
      elif gate.name == 'cx':
        gates.append(('cx', qubits[1], qubits[0], None))                  # cx(control, target)
      else:
        raise ValueError(f"gate {gate.name!r} is not in the reference's gate set ['cx', 'rx', 'cp', 'p']")
      if len(qubits) > 1:
        for other in qubits:
          if other != q:
            handled[other][col] = True
  return gates


def circuit_fitness(objective: str, max_qubits: int, max_depth: int, circuits, target=None, seed=None,
                    penalize: bool = True, device=None) -> np.ndarray:
  """fitness(individual) of evo/run.py:25-31 for a population: R - penalize*C of each whole circuit.

  circuits: list of gate lists [(kind in {'p','rx','cp','cx'}, wire, control, theta)] (see
  `gates_from_individual` for GA individuals).  One replay launch
  applies every circuit from |0..0> / I and terminates it; the reward of the Terminate step is
  metric - cost (env.py:134-135 with sparse=True, punish=penalize)."""
  dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
  n = len(circuits)
  max_len = max(1, max(len(c) for c in circuits))
  tgt = make_target(objective, max_qubits, seed) if target is None else target
  batch = EnvBatch(n, max_qubits, max_depth, objective, tgt, punish=penalize, sparse=True, autoreset=False,
                   obs_mode='none', track_episodes=False, fused_stats=False, device=dev)
  tape = np.stack([encode_circuit(c, max_len) for c in circuits], axis=1)             # [T,N,4]
  rec = {'reward': torch.zeros((max_len + 1, n), dtype=torch.float64, device=dev)}
  batch.replay(torch.as_tensor(tape, device=dev), rec)
  idx = torch.as_tensor([len(c) for c in circuits], device=dev)
  return rec['reward'][idx, torch.arange(n, device=dev)].cpu().numpy()
