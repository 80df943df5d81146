set -e
B="python bench.py --no-cpu --e2e-steps 3 --steps 2000 --warmup 100"
run() { python - "$@" <<'PY'
import json,sys
tag=sys.argv[1]
for line in open(sys.argv[2]):
    if line.startswith('{'):
        d=json.loads(line); print(tag, d['config']['workload'].split()[0], "%.3e"%d['value'], "frac=%.3f"%d['roofline']['frac'], "us=%.1f"%d['roofline']['launch_us'])
PY
}
for v in "256 4" "256 3" "256 2" "128 8" "128 6" "128 4"; do
  set -- $v
  QCD_NVCC_EXTRA="-DQCD_REG_STEP_THREADS=$1 -DQCD_REG_MINB=$2" python -m qcd_b200.build > /dev/null 2>&1
  for w in sp-ghz5 uc-toffoli sp-bell; do
    $B --workload $w > /tmp/o.json 2>/tmp/e.log || tail -3 /tmp/e.log
    run "NT=$1,MINB=$2" /tmp/o.json
  done
  $B --workload sp-ghz5 --envs 1048576 --steps 300 --warmup 20 > /tmp/o.json 2>/tmp/e.log; run "NT=$1,MINB=$2,N=1Mi" /tmp/o.json
done
