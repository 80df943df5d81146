B="python bench.py --no-cpu --e2e-steps 3 --steps 2000 --warmup 100"
show() { python - "$@" <<'PY'
import json,sys
for line in open(sys.argv[2]):
    if line.startswith('{'):
        d=json.loads(line); print(sys.argv[1], d['config']['workload'].split()[0], d['config']['envs_total'], "%.3e"%d['value'], "frac=%.3f"%d['roofline']['frac'], "us=%.2f"%d['roofline']['launch_us'])
PY
}
for pdl in 1 0; do
  for extra in "" "--no-graph"; do
    for w in sp-ghz5 uc-toffoli sp-bell; do
      QCD_PDL=$pdl timeout 120 $B --workload $w $extra > /tmp/o.json 2>/tmp/e.log || tail -3 /tmp/e.log
      show "PDL=$pdl $extra" /tmp/o.json
    done
  done
  QCD_PDL=$pdl timeout 120 $B --workload sp-ghz5 --envs 1048576 --steps 300 --warmup 20 > /tmp/o.json 2>/tmp/e.log; show "PDL=$pdl" /tmp/o.json
done
