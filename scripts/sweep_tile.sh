B="python bench.py --no-cpu --e2e-steps 3 --steps 2000 --warmup 100"
show() { python - "$@" <<'PY'
import json,sys
tag=sys.argv[1]
for line in open(sys.argv[2]):
    if line.startswith('{'):
        d=json.loads(line); print(tag, d['config']['workload'].split()[0], d['config']['envs_total'], "%.3e"%d['value'], "frac=%.3f"%d['roofline']['frac'], "us=%.1f"%d['roofline']['launch_us'])
PY
}
for v in "4 16 1 4" "4 16 0 4" "4 8 1 4" "4 8 0 4" "4 32 1 3" "8 16 1 2" "2 16 1 8" "4 16 1 5"; do
  set -- $v
  QCD_NVCC_EXTRA="-DQCD_WARP_STEP_WARPS=$1 -DQCD_WARP_STEP_TILE=$2 -DQCD_WARP_STEP_STAGE=$3 -DQCD_WARP_STEP_MINB=$4" python -m qcd_b200.build > /dev/null 2>&1
  timeout 120 python __graft_entry__.py smoke 2>&1 | grep -E "smoke|Error|error" | head -3
  for w in sp-ghz5 uc-toffoli sp-bell; do
    timeout 120 $B --workload $w > /tmp/o.json 2>/tmp/e.log || tail -3 /tmp/e.log
    show "W=$1,TE=$2,STAGE=$3,MINB=$4" /tmp/o.json
  done
  timeout 120 $B --workload sp-ghz5 --envs 1048576 --steps 300 --warmup 20 > /tmp/o.json 2>/tmp/e.log; show "W=$1,TE=$2,STAGE=$3,MINB=$4" /tmp/o.json
done
