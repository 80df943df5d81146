#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
show() { python - "$@" <<'PY'
import json,sys
tag=sys.argv[1]
for line in open(sys.argv[2]):
    if line.startswith('{'):
        d=json.loads(line); print(tag, d['config']['workload'].split()[0], "%.3e"%d['value'], "frac=%.3f"%d['roofline']['frac'], "us=%.2f"%d['roofline']['launch_us'])
PY
}
for v in "20 8" "24 4" "28 2" "16 8"; do
  set -- $v
  QCD_NVCC_EXTRA="-DQCD_WARP_STEP_MINB=$1 -DQCD_WARP_PF_V1=$2" python -m qcd_b200.build > /dev/null 2>&1
  for steps in 20 2000; do
    python bench.py --workload sp-ghz5 --no-cpu --no-configs --no-sweep --e2e-steps 0 --steps $steps --warmup 5 > /tmp/o.json 2>/tmp/e.log || tail -3 /tmp/e.log
    show "MINB=$1,PF=$2,steps=$steps" /tmp/o.json
  done
done
