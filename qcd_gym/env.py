"""`qcd_gym.env:CircuitDesigner` — the reference's entry-point path, served by qcd_b200."""
from qcd_b200.env import CircuitDesigner, GATES  # noqa: F401
