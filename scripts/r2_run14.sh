#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_properties.py -x -q -k "fast_shape or step_matches or actions_written or golden or adversarial or invariants" > gpurun_out/r2_14_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r2_14_tests.log
for wl in sp-ghz5 uc-toffoli sp-bell; do
  for steps in 20 2000; do
    python bench.py --workload $wl --no-cpu --no-configs --no-sweep --e2e-steps 0 --steps $steps --warmup 5 > /tmp/o.json 2>/tmp/e.log || tail -3 /tmp/e.log
    python -c "import json;d=json.loads([l for l in open('/tmp/o.json') if l.startswith('{')][0]);print('$wl steps=$steps', '%.3e'%d['value'], 'frac=%.3f'%d['roofline']['frac'], 'us=%.2f'%d['roofline']['launch_us'])"
  done
done
