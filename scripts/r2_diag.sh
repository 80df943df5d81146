#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-4}
i=0
for v in "--exchange overlap --sampler start" "--exchange after --sampler start" "--exchange overlap --sampler early" "--exchange after --sampler early"; do
  i=$((i+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$i bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-configs --no-sweep --e2e-steps 0 $v > gpurun_out/r2_diag_$i.json 2> gpurun_out/r2_diag_$i.err
  python - <<PY
import json
ls=[l for l in open('gpurun_out/r2_diag_$i.json') if l.startswith('{')]
d=json.loads(ls[0]); print("$v", 'ms/step %.5f'%d['ms_per_step'], d['config']['rank_ms'], d['clocks'])
PY
done
