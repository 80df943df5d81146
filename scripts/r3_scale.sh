#!/bin/bash
# bench.py under the driver's protocol on N GPUs of one box (final build): scripts/r3_scale.sh N
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu > gpurun_out/r3_scale_${N}gpu.json 2> gpurun_out/r3_scale_${N}gpu.err
echo rc=$?; tail -2 gpurun_out/r3_scale_${N}gpu.err
python - $N <<'PY'
import json, sys
d=json.loads([l for l in open(f"gpurun_out/r3_scale_{sys.argv[1]}gpu.json") if l.startswith("{")][0])
print("value %.4e frac %.3f ms %.5f e2e %.4e" % (d["value"], d["roofline"]["frac"], d["ms_per_step"], d["e2e"]["value"]))
for k,v in d["configs"].items(): print(k, "%.4e" % v["value"], (v.get("roofline") or {}).get("frac"))
PY
