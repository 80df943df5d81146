#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
run() {
  QCD_NVCC_EXTRA="$1" python -m qcd_b200.build 2>&1 | grep -A3 "step_warp_kernelILi5ELb0" | grep -E "spill|registers" | sed 's/ptxas info    : //' | tr '\n' ' '; echo
  python bench.py --workload sp-ghz5 --no-cpu --no-configs --no-sweep --e2e-steps 0 --steps 2000 --warmup 20 > /tmp/o.json 2>/tmp/e.log || tail -3 /tmp/e.log
  python -c "import json;d=json.loads([l for l in open('/tmp/o.json') if l.startswith('{')][0]);print('$1', '%.3e'%d['value'], 'frac=%.3f'%d['roofline']['frac'], 'us=%.2f'%d['roofline']['launch_us'])"
}
run "-DQCD_WARP_LOG2L_S32=3 -DQCD_WARP_BRANCHLESS_MAX_V=4 -DQCD_WARP_PF_V4=2 -DQCD_WARP_STEP_MINB=14"
run "-DQCD_WARP_LOG2L_S32=3 -DQCD_WARP_BRANCHLESS_MAX_V=4 -DQCD_WARP_PF_V4=1 -DQCD_WARP_STEP_MINB=14"
run "-DQCD_WARP_LOG2L_S32=4 -DQCD_WARP_PF_V2=4 -DQCD_WARP_STEP_MINB=14"
run "-DQCD_WARP_STEP_MINB=14"
run "-DQCD_WARP_STEP_MINB=14 -DQCD_WARP_PF_V1=16 -DQCD_WARP_CHUNK=16"
