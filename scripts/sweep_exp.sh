B="python bench.py --no-cpu --e2e-steps 3 --steps 2000 --warmup 100 --no-stats"
show() { python - "$@" <<'PY'
import json,sys
tag=sys.argv[1]
for line in open(sys.argv[2]):
    if line.startswith('{'):
        d=json.loads(line); print(tag, d['config']['workload'].split()[0], d['config']['envs_total'], "%.3e"%d['value'], "frac=%.3f"%d['roofline']['frac'], "us=%.1f"%d['roofline']['launch_us'])
PY
}
for x in "" "-DQCD_EXP_NOSINCOS"; do
  QCD_NVCC_EXTRA="$x" python -m qcd_b200.build > /dev/null 2>&1
  $B --workload sp-ghz5 > /tmp/o.json 2>/tmp/e.log; show "[$x]" /tmp/o.json
  $B --workload sp-ghz5 --obs none > /tmp/o.json 2>/tmp/e.log; show "[$x] obs=none" /tmp/o.json
  $B --workload sp-ghz5 --envs 1048576 --steps 300 --warmup 20 > /tmp/o.json 2>/tmp/e.log; show "[$x]" /tmp/o.json
  $B --workload sp-ghz5 --envs 262144 --steps 1000 --warmup 20 > /tmp/o.json 2>/tmp/e.log; show "[$x]" /tmp/o.json
  $B --workload sp-ghz5 --no-rotate > /tmp/o.json 2>/tmp/e.log; show "[$x] L2-resident" /tmp/o.json
done
