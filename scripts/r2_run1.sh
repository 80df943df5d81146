#!/bin/bash
# round 2, GPU call 1: new parity tests, the whole GPU suite, PDL action-load cost, measured FP64/smem peaks
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_1_smi.txt
./scripts/microbench_peaks > gpurun_out/r2_fp64_smem_peaks.json 2> gpurun_out/r2_peaks.err
cat gpurun_out/r2_fp64_smem_peaks.json
timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -k "not terminal_observations" > gpurun_out/r2_1_tests_new.log 2>&1
echo "new tests rc=$?"; tail -5 gpurun_out/r2_1_tests_new.log
timeout 1200 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_round2.py > gpurun_out/r2_1_tests_all.log 2>&1
echo "all tests rc=$?"; tail -3 gpurun_out/r2_1_tests_all.log
for extra in "" "--actions-stable"; do
  python bench.py --steps 2000 --warmup 100 --no-cpu --no-sweep --e2e-steps 0 $extra > gpurun_out/r2_1_bench_ghz5${extra}.json 2>gpurun_out/r2_1_bench.err
  python -c "import json;d=json.load(open('gpurun_out/r2_1_bench_ghz5${extra}.json'));print('ghz5 ${extra}', d['value'], d['roofline']['frac'])"
  python bench.py --steps 20 --warmup 5 --no-cpu --no-sweep --e2e-steps 0 $extra > gpurun_out/r2_1_bench20_ghz5${extra}.json 2>>gpurun_out/r2_1_bench.err
  python -c "import json;d=json.load(open('gpurun_out/r2_1_bench20_ghz5${extra}.json'));print('ghz5 20 steps ${extra}', d['value'], d['roofline']['frac'])"
  python bench.py --workload uc-toffoli --steps 2000 --warmup 100 --no-cpu --no-sweep --e2e-steps 0 $extra > gpurun_out/r2_1_bench_toffoli${extra}.json 2>>gpurun_out/r2_1_bench.err
  python -c "import json;d=json.load(open('gpurun_out/r2_1_bench_toffoli${extra}.json'));print('toffoli ${extra}', d['value'], d['roofline']['frac'])"
done
