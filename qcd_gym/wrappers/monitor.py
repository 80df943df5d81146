"""`qcd_gym.wrappers.monitor:Monitor` — the reference's import path, served by qcd_b200."""
from qcd_b200.monitor import Monitor  # noqa: F401
