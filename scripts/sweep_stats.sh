B="python bench.py --no-cpu --e2e-steps 3 --steps 2000 --warmup 100"
show() { python - "$@" <<'PY'
import json,sys
tag=sys.argv[1]
for line in open(sys.argv[2]):
    if line.startswith('{'):
        d=json.loads(line); print(tag, d['config']['workload'].split()[0], "%.3e"%d['value'], "frac=%.3f"%d['roofline']['frac'], "us=%.1f"%d['roofline']['launch_us'])
PY
}
for v in "256 4" "256 3" "128 8"; do
  set -- $v
  QCD_NVCC_EXTRA="-DQCD_REG_STEP_THREADS=$1 -DQCD_REG_MINB=$2" python -m qcd_b200.build > /dev/null 2>&1
  for w in sp-ghz5 uc-toffoli sp-bell; do
    $B --workload $w --no-stats > /tmp/o.json 2>/tmp/e.log || tail -3 /tmp/e.log
    show "nostats NT=$1,MINB=$2" /tmp/o.json
  done
  $B --workload sp-ghz5 --no-stats --envs 1048576 --steps 300 --warmup 20 > /tmp/o.json 2>/tmp/e.log; show "nostats NT=$1,MINB=$2,N=1Mi" /tmp/o.json
done
QCD_NVCC_EXTRA="-DQCD_REG_STEP_MAX_LOG2S=0" python -m qcd_b200.build > /dev/null 2>&1
for w in sp-ghz5 uc-toffoli sp-bell; do $B --workload $w --no-stats > /tmp/o.json 2>/tmp/e.log; show "nostats pair-streaming" /tmp/o.json; done
