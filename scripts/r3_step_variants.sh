#!/bin/bash
# compare prebuilt library variants (scripts/variants.sh) on the STEP workloads: parity tests first, then the bench
# usage: scripts/r3_step_variants.sh "<pytest -k expr>" name1 name2 ...
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
KEXPR=$1; shift
for v in "$@"; do
  export QCD_B200_LIB=$PWD/build/variants/$v.so
  if [ -n "$KEXPR" ]; then
    timeout 600 python -m pytest tests -x -q -m gpu -k "$KEXPR" > gpurun_out/r3_${v}_tests.log 2>&1
    echo "$v tests rc=$? $(tail -1 gpurun_out/r3_${v}_tests.log)"
  fi
  for w in ${WORKLOADS:-sp-ghz5 uc-toffoli sp-bell}; do
    for steps in 20 2000; do
      timeout 300 python bench.py --workload $w --no-cpu --no-configs --no-sweep --e2e-steps 0 --steps $steps --warmup 5 > /tmp/o.json 2>/tmp/e.log || tail -3 /tmp/e.log
      python -c "import json;d=json.loads([l for l in open('/tmp/o.json') if l.startswith('{')][0]);print('$v $w steps=$steps', '%.4e'%d['value'], 'frac=%.3f'%d['roofline']['frac'], 'us=%.2f'%d['roofline']['launch_us'])"
    done
  done
done
