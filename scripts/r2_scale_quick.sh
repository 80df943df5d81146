#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
F="--steps 20 --warmup 5 --no-cpu --no-configs --no-sweep --e2e-steps 0"
for rep in 1 2; do
python bench.py --gpus 1 $F > gpurun_out/r2_sq_1_$rep.json 2> gpurun_out/r2_sq_1.err
for N in 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$rep bench.py --gpus $N $F > gpurun_out/r2_sq_${N}_$rep.json 2> gpurun_out/r2_sq_$N.err
done
done
python - <<PY
import json
for rep in (1,2):
  base=None
  for N in (1,8):
    ls=[l for l in open(f'gpurun_out/r2_sq_{N}_{rep}.json') if l.startswith('{')]
    if not ls: print(N,'no line'); continue
    d=json.loads(ls[0])
    if base is None: base=d['value']
    print(rep, N, 'value %.3e eff %.3f frac %.3f ms %.5f' % (d['value'], d['value']/base/N, d['roofline']['frac'], d['ms_per_step']), d['config']['rank_ms'])
PY
