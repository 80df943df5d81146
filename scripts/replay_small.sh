B="python bench.py --no-cpu --steps 40 --warmup 4"
for w in sp-ghz5-replay; do
  for x in "" "--no-terminate"; do
  timeout 300 $B --workload $w $x > /tmp/o.json 2>/tmp/e.log || tail -3 /tmp/e.log
  python - "$1" /tmp/o.json <<'PY'
import json,sys
for line in open(sys.argv[2]):
    if line.startswith('{'):
        d=json.loads(line); print(sys.argv[1], d['config']['workload'], d['config']['actions'][:12], "%.3e"%d['value'], "ms/launch=%.3f"%d['ms_per_step'])
PY
  done
done
