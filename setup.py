"""Packaging of the B200-native drop-in for qcd-gym's hot path.

Same plugin declaration as the reference (/root/reference/setup.py:21): the entry-point group `gymnasium.envs`,
name `__root__`, target `qcd_gym.__init__:register_envs`, so that `gymnasium.make("CircuitDesigner-v0", ...)`
finds the env of an installed copy without an explicit import.  `qcd_gym` here is the alias package of
`qcd_b200`; the CUDA library `qcd_b200/libqcd_b200.so` is compiled for sm_100a by `qcd_b200.build` (at
install time when nvcc is available, otherwise on first use - there is no CPU fallback).
"""
from pathlib import Path

from setuptools import find_packages, setup
from setuptools.command.build_py import build_py


class build_py_with_cuda(build_py):
  def run(self):
    try:                                  # a regular (non-editable) install ships the built library
      import importlib.util
      spec = importlib.util.spec_from_file_location("qcd_b200_build", Path(__file__).parent / "qcd_b200" / "build.py")
      mod = importlib.util.module_from_spec(spec)
      spec.loader.exec_module(mod)
      mod.build_library()
    except Exception as exc:              # no nvcc on the build host: the library is built on first use
      print(f"warning: libqcd_b200.so not built now ({exc})")
    super().run()


setup(
  name="qcd-b200", version="0.2.0",
  description="B200-native batched simulator behind qcd-gym's CircuitDesigner-v0 gymnasium API",
  long_description=(Path(__file__).parent / "README.md").read_text(), long_description_content_type="text/markdown",
  packages=find_packages(include=['qcd_b200', 'qcd_gym', 'qcd_gym.wrappers']),
  package_data={'qcd_b200': ['libqcd_b200.so', 'libqcd_b200.so.srchash', 'csrc/*.cu']},
  data_files=[('include', ['include/qcd_b200.h'])],
  install_requires=["numpy", "torch"],
  extras_require={"gym": ["gymnasium>=0.29"], "train": ["stable_baselines3>=2.0.0"], "tests": ["pytest", "hypothesis"]},
  python_requires=">=3.9",
  cmdclass={'build_py': build_py_with_cuda},
  entry_points={"gymnasium.envs": ["__root__ = qcd_gym.__init__:register_envs"]},
)
