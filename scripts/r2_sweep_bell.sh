#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for v in "16" "21" "24" "28" "32"; do
  QCD_NVCC_EXTRA="-DQCD_WARP_STEP_MINB=$v" python -m qcd_b200.build 2>&1 | grep -A3 "step_warp_kernelILi2ELb0" | grep -E "spill|registers" | tr '\n' ' '; echo
  for wl in sp-bell uc-hadamard1 sp-ghz3; do
  python bench.py --workload $wl --envs 65536 --no-cpu --no-configs --no-sweep --e2e-steps 0 --steps 2000 --warmup 20 > /tmp/o.json 2>/tmp/e.log || tail -3 /tmp/e.log
  python -c "import json;d=json.loads([l for l in open('/tmp/o.json') if l.startswith('{')][0]);print('MINB=$v $wl', '%.3e'%d['value'], 'frac=%.3f'%d['roofline']['frac'], 'us=%.2f'%d['roofline']['launch_us'])"
  done
done
