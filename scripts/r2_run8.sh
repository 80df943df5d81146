#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
./scripts/microbench_peaks > gpurun_out/r2_fp64_smem_peaks_v2.json 2> gpurun_out/r2_peaks.err; cat gpurun_out/r2_fp64_smem_peaks_v2.json
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_8_bench_default.json 2> gpurun_out/r2_8_bench_default.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_8_bench_default.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_8_bench_default.json') if l.startswith('{')][0])
print('value',d['value'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'], 'graph_len', d['config']['graph_len'])
for k,v in (d.get('configs') or {}).items():
    print(k, v.get('value'), (v.get('roofline') or {}).get('frac'), (v.get('cpu_baseline') or {}).get('value'), v.get('skipped'))
PY
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2_8_bench_ref.json 2> gpurun_out/r2_8_bench_ref.err
tail -3 gpurun_out/r2_8_bench_ref.err; cut -c1-300 gpurun_out/r2_8_bench_ref.json
