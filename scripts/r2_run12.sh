#!/bin/bash
# profiles for round 2: launch list of the bench's timed path + full capture of the step kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 48 --warmup 8 --no-cpu --no-configs --no-sweep --e2e-steps 3 --no-graph"
$CMD > gpurun_out/r2_12_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 80 --csv --log-file gpurun_out/r02_launches_sp-ghz5.csv $CMD > gpurun_out/r2_12_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:step_warp -s 20 -c 1 -o gpurun_out/r2_prof_step_warp_ghz5 $CMD > gpurun_out/r2_12_ncu2.log 2>&1
echo "full capture rc=$?"
