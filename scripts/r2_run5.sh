#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_next_rows.py tests/test_gpu_api.py -x -q -k "replay or config5 or fitness or terminal or per_env or single or kats or monitor" > gpurun_out/r2_16_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2_16_tests.log
for wl in uc-random6-replay sp-random12-replay; do
  for extra in "" "--no-terminate"; do
    python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu $extra > gpurun_out/r2_16_bench_${wl}${extra}.json 2>>gpurun_out/r2_16_bench.err
    python -c "import json;d=json.load(open('gpurun_out/r2_16_bench_${wl}${extra}.json'));print('${wl} ${extra}', d['value'], d['ms_per_step'])"
  done
done
python bench.py --workload uc-random6-replay --steps 6 --warmup 3 --no-cpu --no-graph > gpurun_out/r2_16_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:replay_reg -s 3 -c 1 -o gpurun_out/r2_prof_replay_reg_uc6_v8 \
  python bench.py --workload uc-random6-replay --steps 6 --warmup 3 --no-cpu --no-graph > gpurun_out/r2_16_ncu.log 2>&1
echo "ncu rc=$?"
