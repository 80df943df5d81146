# usage: quick.sh "<nvcc extra>" [workloads...]
X="$1"; shift
B="python bench.py --no-cpu --e2e-steps 3 --steps 2000 --warmup 100"
QCD_NVCC_EXTRA="$X" python -m qcd_b200.build > /dev/null 2>&1 || { echo build failed; exit 1; }
timeout 120 python __graft_entry__.py smoke 2>&1 | grep -E "smoke|Error|error|assert" | head -3
for w in "$@"; do
  timeout 120 $B --workload $w > /tmp/o.json 2>/tmp/e.log || tail -3 /tmp/e.log
  python - "[$X]" /tmp/o.json <<'PY'
import json,sys
for line in open(sys.argv[2]):
    if line.startswith('{'):
        d=json.loads(line); print(sys.argv[1], d['config']['workload'].split()[0], d['config']['envs_total'], "%.3e"%d['value'], "frac=%.3f"%d['roofline']['frac'], "us=%.1f"%d['roofline']['launch_us'])
PY
done
timeout 120 $B --workload sp-ghz5 --envs 1048576 --steps 300 --warmup 20 > /tmp/o.json 2>/tmp/e.log
python - "[$X]" /tmp/o.json <<'PY'
import json,sys
for line in open(sys.argv[2]):
    if line.startswith('{'):
        d=json.loads(line); print(sys.argv[1], d['config']['workload'].split()[0], d['config']['envs_total'], "%.3e"%d['value'], "frac=%.3f"%d['roofline']['frac'], "us=%.1f"%d['roofline']['launch_us'])
PY
