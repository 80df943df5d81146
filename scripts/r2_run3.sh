#!/bin/bash
# round 2, GPU call 3: replay kernel after the register-copy fix; single-env zero-copy; pipelined e2e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_next_rows.py tests/test_gpu_api.py -x -q -k "replay or config5 or fitness or terminal or single or pipelined or kats or vector_env or monitor or per_env" > gpurun_out/r2_3_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r2_3_tests.log
for wl in uc-random6-replay sp-random12-replay; do
  for extra in "" "--no-terminate"; do
    python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu $extra > gpurun_out/r2_3_bench_${wl}${extra}.json 2>>gpurun_out/r2_3_bench.err
    python -c "import json;d=json.load(open('gpurun_out/r2_3_bench_${wl}${extra}.json'));print('${wl} ${extra}', d['value'], d['ms_per_step'])"
  done
done
python scripts/single_env_rate.py 2>&1 | tee gpurun_out/r2_3_single_env.txt
for g in 1 2 3; do for ob in state full; do
  python bench.py --steps 200 --warmup 20 --no-cpu --no-sweep --e2e-groups $g --e2e-obs $ob > gpurun_out/r2_3_e2e_g${g}_${ob}.json 2>>gpurun_out/r2_3_bench.err
  python -c "import json;d=json.load(open('gpurun_out/r2_3_e2e_g${g}_${ob}.json'));print('e2e groups=$g obs=$ob', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['d2h_gbs_per_gpu'])"
done; done
python bench.py --workload uc-random6-replay --steps 6 --warmup 3 --no-cpu --no-graph > gpurun_out/r2_3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:replay_reg -s 3 -c 1 -o gpurun_out/r2_prof_replay_reg_uc6_v2 \
  python bench.py --workload uc-random6-replay --steps 6 --warmup 3 --no-cpu --no-graph > gpurun_out/r2_3_ncu.log 2>&1
echo "ncu rc=$?"
