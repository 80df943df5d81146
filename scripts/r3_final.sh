#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r3_final_tests.log 2>&1
echo "gpu tests rc=$?"; tail -4 gpurun_out/r3_final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r3_final_ref.json 2> gpurun_out/r3_final_ref.err
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r3_final_bench.json 2> gpurun_out/r3_final_bench.err
tail -3 gpurun_out/r3_final_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r3_final_bench.json') if l.startswith('{')][0])
r=json.loads([l for l in open('gpurun_out/r3_final_ref.json') if l.startswith('{')][0])
print('value %.4e frac %.3f modelled %.3f ms %.5f e2e %.4e ref %.4e e2e/ref %.0f' % (d['value'], d['roofline']['frac'], d['roofline']['frac_of_modelled_traffic'], d['ms_per_step'], d['e2e']['value'], r['value'], d['e2e']['value']/r['value']))
print('graph_len', d['config']['graph_len'], 'gpu_launches', d['gpu_launches'], 'clocks', d['clocks'])
for k,v in d['configs'].items(): print(k, '%.4e' % v['value'], (v.get('roofline') or {}).get('frac'), (v.get('cpu_baseline') or {}).get('value'))
PY
