#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_15_tests.log 2>&1
echo "gpu tests rc=$?"; tail -6 gpurun_out/r2_15_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python - <<'PY'
import time, torch
from qcd_b200.baselines import random_baseline
for task in ('SP-bell-q2-d12', 'SP-ghz3-q3-d15', 'UC-toffoli-q3-d63'):
    random_baseline(task, seeds=2, episodes=10)
    torch.cuda.synchronize(); t0=time.perf_counter()
    r = random_baseline(task)
    torch.cuda.synchronize(); print(task, 'random baseline (8 seeds x 100 episodes) %.1f ms' % ((time.perf_counter()-t0)*1e3), {k: round(float(v.mean()),4) for k,v in r.items()})
PY
