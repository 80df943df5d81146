#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python scripts/host_d2h_ceiling.py 2>/dev/null | tee gpurun_out/r02_host_d2h_ceiling.jsonl
for N in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N scripts/host_d2h_ceiling.py 2>/dev/null | grep '^{' | tee -a gpurun_out/r02_host_d2h_ceiling.jsonl
done
