"""Import-compatible alias of the reference package name (`qcd_gym`, /root/reference/qcd_gym/__init__.py):
`import qcd_gym; qcd_gym.register_envs()` keeps working, backed by the B200 engine in `qcd_b200`."""
from qcd_b200.registration import register_envs, ENV_ID  # noqa: F401
