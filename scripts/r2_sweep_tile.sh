#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for v in "24 19 8" "24 20 8" "24 18 8" "24 16 8" "16 28 4" "16 24 8" "16 20 8" "32 16 8"; do
  set -- $v
  QCD_NVCC_EXTRA="-DQCD_WARP_STEP_TILE=$1 -DQCD_WARP_STEP_MINB=$2 -DQCD_WARP_PF_V1=$3" python -m qcd_b200.build 2>&1 | grep -A3 "step_warp_kernelILi5ELb0" | grep -E "spill|registers" | sed 's/ptxas info    : //' | tr '\n' ' '; echo
  python bench.py --workload sp-ghz5 --no-cpu --no-configs --no-sweep --e2e-steps 0 --steps 2000 --warmup 20 > /tmp/o.json 2>/tmp/e.log || tail -3 /tmp/e.log
  python -c "import json;d=json.loads([l for l in open('/tmp/o.json') if l.startswith('{')][0]);print('TE=$1 MINB=$2 PF=$3', '%.3e'%d['value'], 'frac=%.3f'%d['roofline']['frac'], 'us=%.2f'%d['roofline']['launch_us'])"
done
