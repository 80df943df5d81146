#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_11_tests.log 2>&1
echo "gpu tests rc=$?"; tail -15 gpurun_out/r2_11_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
