"""The drop-in boundary as an INSTALLED plugin (CPU-only checks).

* `pip install -e . --no-deps` exposes the reference's entry point (`gymnasium.envs: __root__ =
  qcd_gym.__init__:register_envs`, /root/reference/setup.py:21) and loading it registers CircuitDesigner-v0
  with a vector entry point;
* with gymnasium / stable-baselines3 importable (minimal stand-ins of their class hierarchy here: neither is
  installable in this image) the batched envs ARE `gymnasium.vector.VectorEnv` / SB3 `VecEnv` subclasses and the
  registration goes through gymnasium's own `register`.
"""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _py(code, cwd, env=None, timeout=300):
  e = dict(os.environ)
  e.pop("PYTHONPATH", None)
  e.update(env or {})
  return subprocess.run([sys.executable, "-c", textwrap.dedent(code)], capture_output=True, text=True, cwd=cwd, env=e,
                        timeout=timeout)


def test_pip_editable_install_exposes_the_gymnasium_entry_point(tmp_path):
  prefix = tmp_path / "prefix"
  res = subprocess.run([sys.executable, "-m", "pip", "install", "--no-deps", "--no-build-isolation", "--no-index", "-q",
                        "-e", ROOT, "--prefix", str(prefix)], capture_output=True, text=True, timeout=600)
  assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
  site_dirs = [str(p.parent) for p in prefix.rglob("*.dist-info")]
  assert site_dirs, "no dist-info written"
  code = f"""
    import site, sys
    site.addsitedir({site_dirs[0]!r})            # processes the editable .pth like a real site-packages
    from importlib.metadata import entry_points
    eps = entry_points(group="gymnasium.envs")
    ep = [e for e in eps if e.name == "__root__" and e.value == "qcd_gym.__init__:register_envs"]
    assert len(ep) == 1, list(eps)
    ep[0].load()()                               # what gymnasium does with plugin entry points
    import qcd_b200
    from qcd_b200.gym_compat import registry
    spec = registry["CircuitDesigner-v0"]
    assert spec.entry_point == "qcd_b200.env:CircuitDesigner"
    assert spec.vector_entry_point == "qcd_b200.vector_env:CircuitDesignerVectorEnv"
    assert qcd_b200.__file__.startswith({ROOT!r})
    print("OK")
  """
  out = _py(code, cwd=str(tmp_path))
  assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr


STUBS = {
    "gymnasium/__init__.py": """
        from . import spaces, vector, error, utils
        from .core import Env, Wrapper
        from .envs.registration import make, make_vec, register, registry
        __version__ = "0.29.0-stub"
    """,
    "gymnasium/core.py": """
        class Env:
          metadata = {}; render_mode = None; spec = None
          @property
          def unwrapped(self): return self
          def reset(self, *, seed=None, options=None): pass
        class Wrapper(Env):
          def __init__(self, env): self.env = env
    """,
    "gymnasium/spaces.py": """
        import numpy as np
        class Box:
          def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            self.low = np.asarray(low, dtype=dtype); self.high = np.asarray(high, dtype=dtype)
            self.shape = tuple(shape) if shape is not None else self.low.shape; self.dtype = np.dtype(dtype)
    """,
    "gymnasium/vector.py": """
        class VectorEnv:
          '''stand-in for gymnasium.vector.VectorEnv'''
          closed = False
    """,
    "gymnasium/error.py": "class ResetNeeded(Exception): pass\n",
    "gymnasium/utils/__init__.py": "from . import seeding\n",
    "gymnasium/utils/seeding.py": """
        import numpy as np
        def np_random(seed=None):
          ss = np.random.SeedSequence(seed)
          return np.random.Generator(np.random.PCG64(ss)), ss.entropy
    """,
    "gymnasium/envs/__init__.py": "from . import registration\n",
    "gymnasium/envs/registration.py": """
        registry = {}
        calls = []
        class EnvSpec:
          def __init__(self, **kw): self.__dict__.update(kw)
        def register(id, entry_point=None, reward_threshold=None, nondeterministic=False, max_episode_steps=None,
                     order_enforce=True, autoreset=False, disable_env_checker=False, apply_api_compatibility=False,
                     additional_wrappers=(), vector_entry_point=None, **kwargs):
          calls.append((id, entry_point, vector_entry_point))
          registry[id] = EnvSpec(id=id, entry_point=entry_point, vector_entry_point=vector_entry_point, kwargs=kwargs)
        def make(id, **kw): raise NotImplementedError
        def make_vec(id, num_envs=1, **kw): raise NotImplementedError
    """,
    "stable_baselines3/__init__.py": "",
    "stable_baselines3/common/__init__.py": "",
    "stable_baselines3/common/vec_env/__init__.py": """
        class VecEnv:
          '''stand-in for stable_baselines3.common.vec_env.VecEnv'''
          def __init__(self, num_envs, observation_space, action_space):
            self.num_envs, self.observation_space, self.action_space = num_envs, observation_space, action_space
            self.render_mode = None
    """,
}


def test_with_gymnasium_and_sb3_importable_the_batched_envs_are_real_subclasses(tmp_path):
  for rel, src in STUBS.items():
    f = tmp_path / rel
    f.parent.mkdir(parents=True, exist_ok=True)
    f.write_text(textwrap.dedent(src))
  code = """
    import gymnasium, stable_baselines3.common.vec_env as sb3v
    import qcd_gym
    from qcd_b200 import gym_compat
    assert gym_compat.HAVE_GYMNASIUM and gym_compat.Env is gymnasium.Env and gym_compat.Box is gymnasium.spaces.Box
    qcd_gym.register_envs(); qcd_gym.register_envs()
    from gymnasium.envs.registration import calls, registry
    assert calls == [("CircuitDesigner-v0", "qcd_b200.env:CircuitDesigner", "qcd_b200.vector_env:CircuitDesignerVectorEnv")], calls
    from qcd_b200.vector_env import CircuitDesignerVectorEnv
    from qcd_b200.sb3_vec_env import CircuitDesignerSB3VecEnv, HAVE_SB3
    from qcd_b200.env import CircuitDesigner
    from qcd_b200.monitor import Monitor
    assert issubclass(CircuitDesignerVectorEnv, gymnasium.vector.VectorEnv)
    assert HAVE_SB3 and issubclass(CircuitDesignerSB3VecEnv, sb3v.VecEnv)
    assert issubclass(CircuitDesigner, gymnasium.Env) and issubclass(Monitor, gymnasium.Wrapper)
    print("OK")
  """
  out = _py(code, cwd=str(tmp_path), env={"PYTHONPATH": os.pathsep.join([str(tmp_path), ROOT])})
  assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr


class _Gate:
  def __init__(self, name, qubits, params=()):
    self.name, self.affected_qubits, self.parameters = name, list(qubits), list(params)


def test_ga_individual_grid_maps_to_the_reference_circuit_order():
  """individual.py:158-201: columns outer, qubits inner, two-qubit gates emitted once, cx(control, target)."""
  import numpy as np
  from oracle.qcd_oracle import simulate_statevector
  from qcd_b200.baselines import encode_circuit, gates_from_individual
  cx = _Gate('cx', [2, 0]); cp = _Gate('cp', [1, 2], [0.7])
  grid = [[_Gate('rx', [0], [0.3]), cx,   None],                     # qubit 0
          [None,                    None, cp],                       # qubit 1
          [_Gate('p', [2], [1.1]),  cx,   cp]]                       # qubit 2
  gates = gates_from_individual(grid)
  assert gates == [('rx', 0, 0, 0.3), ('p', 2, 2, 1.1), ('cx', 0, 2, None), ('cp', 2, 1, 0.7)]

  class Ind: solution, n_qubits, n_gates_per_qubit = grid, 3, 3
  assert gates_from_individual(Ind()) == gates
  tape = encode_circuit(gates, 5)
  assert tape.shape == (6, 4) and tape[:, 0].tolist() == [1, 0, 1, 0, 2, 2] and tape[2, 1:3].tolist() == [0, 2]
  # the tape's action encoding decodes back to the same qiskit-ordered operations (env.py:95-98)
  ops = [(k, [w] if k in ('p', 'rx') else [c, w], th) for k, w, c, th in gates]
  psi = simulate_statevector(ops, 3)
  assert abs(np.vdot(psi, psi) - 1) < 1e-12
