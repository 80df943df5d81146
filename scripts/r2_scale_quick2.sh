#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
F="--steps 20 --warmup 5 --no-cpu --no-configs --no-sweep --e2e-steps 0"
i=0
for v in "--exchange after" "--exchange overlap" "--exchange after" "--exchange overlap"; do
i=$((i+1))
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$i bench.py --gpus 8 $F $v > gpurun_out/r2_sq2_$i.json 2> gpurun_out/r2_sq2_$i.err
python - <<PY
import json
ls=[l for l in open('gpurun_out/r2_sq2_$i.json') if l.startswith('{')]
d=json.loads(ls[0]); print("$v", 'ms %.5f' % d['ms_per_step'], d['config']['rank_ms'])
PY
done
