#!/bin/bash
# compare prebuilt library variants (scripts/variants.sh) on the replay workloads: parity tests first, then the bench
# usage: scripts/r3_variants.sh "<pytest -k expr>" name1 name2 ...
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
KEXPR=$1; shift
for v in "$@"; do
  export QCD_B200_LIB=$PWD/build/variants/$v.so
  if [ -n "$KEXPR" ]; then
    timeout 600 python -m pytest tests -x -q -m gpu -k "$KEXPR" > gpurun_out/r3_${v}_tests.log 2>&1
    echo "$v tests rc=$? $(tail -1 gpurun_out/r3_${v}_tests.log)"
  fi
  for w in "uc-random6-replay" "uc-random6-replay --no-terminate" "sp-random12-replay" "uc-random6-replay --envs 4096"; do
    timeout 300 python bench.py --no-cpu --no-configs --steps ${STEPS:-6} --warmup 2 --workload $w > /tmp/o.json 2>/tmp/e.log || tail -3 /tmp/e.log
    python - "$v" /tmp/o.json "$w" <<'PY'
import json,sys
for line in open(sys.argv[2]):
    if line.startswith('{'):
        d=json.loads(line); print(sys.argv[1], sys.argv[3], "%.4e"%d['value'], "ms/launch=%.3f"%d['ms_per_step'])
PY
  done
done
