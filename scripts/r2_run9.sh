#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_9_bench_1gpu.json 2> gpurun_out/r2_9_bench_1gpu.err
  python scripts/single_env_rate.py 2>&1 | tee gpurun_out/r2_9_single_env.txt
else
  python bench.py --gpus 1 --steps 20 --warmup 5 --no-configs --no-cpu > gpurun_out/r2_9_bench_1gpu.json 2> gpurun_out/r2_9_bench_1gpu.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_9_bench_${N}gpu.json 2> gpurun_out/r2_9_bench_${N}gpu.err
  echo "rc=$?"; tail -5 gpurun_out/r2_9_bench_${N}gpu.err
fi
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_9_bench_*gpu.json')):
    ls=[l for l in open(f) if l.startswith('{')]
    if not ls: print(f,'no line'); continue
    d=json.loads(ls[0])
    print(f, 'value',d['value'],'frac',d['roofline']['frac'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
    for k,v in (d.get('configs') or {}).items():
        print('  ',k, v.get('value'), (v.get('roofline') or {}).get('frac'), v.get('skipped'))
PY
