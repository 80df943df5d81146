#!/bin/bash
# 1 -> 8 GPU scaling under the driver's protocol (--steps 20 --warmup 5), one box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_scale_${N}gpu.json 2> gpurun_out/r2_scale_${N}gpu.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_scale_${N}gpu.json 2> gpurun_out/r2_scale_${N}gpu.err
  fi
  echo "N=$N rc=$?"
done
python - <<PY
import json
base=None
for N in (1,2,4,8):
    ls=[l for l in open(f'gpurun_out/r2_scale_{N}gpu.json') if l.startswith('{')]
    if not ls: print(N,'no line'); continue
    d=json.loads(ls[0])
    if base is None: base=(d['value'], d['e2e']['value'])
    print(N, 'value %.3e eff %.3f frac %.3f ms %.5f | e2e %.3e eff %.3f' % (d['value'], d['value']/base[0]/N, d['roofline']['frac'], d['ms_per_step'], d['e2e']['value'], d['e2e']['value']/base[1]/N))
    for k,v in (d.get('configs') or {}).items():
        print('    ',k, '%.4e'%v['value'] if v.get('value') else v, (v.get('roofline') or {}).get('frac'))
PY
nvidia-smi topo -m > gpurun_out/r2_scale_topo.txt 2>&1; lscpu | head -20 > gpurun_out/r2_scale_lscpu.txt; numactl -H >> gpurun_out/r2_scale_lscpu.txt 2>&1
