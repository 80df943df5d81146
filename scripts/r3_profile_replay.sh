#!/bin/bash
# full ncu capture of replay_reg_kernel (UC eta=6, uniform-box tape, 4096 envs, T=64): plain run first, then the profiler
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r3}
CMD="python bench.py --workload uc-random6-replay --steps 6 --warmup 3 --no-graph --no-cpu --no-configs"
$CMD > gpurun_out/${TAG}_replay_plain.log 2>&1 || { tail -5 gpurun_out/${TAG}_replay_plain.log; exit 1; }
grep '^{' gpurun_out/${TAG}_replay_plain.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:replay_reg -s 4 -c 1 -o gpurun_out/${TAG}_prof_replay_reg_uc6 -f $CMD > gpurun_out/${TAG}_replay_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/${TAG}_replay_ncu.log
