#!/bin/bash
# Steady-state DRAM traffic per launch of the step kernels: single-pass ncu metrics (no kernel replay), caches NOT
# flushed between launches, the bench rotating its batch replicas as in the timed run - so the write-back L2's
# evictions of earlier launches are counted where they happen.  Plain run first.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for w in sp-ghz5 uc-toffoli; do
  CMD="python bench.py --workload $w --steps 48 --warmup 8 --no-cpu --no-configs --no-sweep --e2e-steps 0 --no-graph"
  $CMD > gpurun_out/r3_traffic_${w}_plain.log 2>&1 || { tail -3 gpurun_out/r3_traffic_${w}_plain.log; continue; }
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none \
      -k regex:step_warp -s 16 -c 40 --csv --log-file gpurun_out/r3_traffic_${w}.csv $CMD > gpurun_out/r3_traffic_${w}_ncu.log 2>&1
  echo "$w ncu rc=$?"
  python - gpurun_out/r3_traffic_${w}.csv <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; im, iv, iu = hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
acc = collections.defaultdict(list)
for r in rows[1:]:
  v = float(r[iv].replace(',', ''))
  u = r[iu]
  if u == 'Mbyte': v *= 1e6
  elif u == 'Kbyte': v *= 1e3
  elif u == 'Gbyte': v *= 1e9
  acc[r[im]].append(v)
for k, v in acc.items(): print(k, 'launches', len(v), 'mean %.4g' % (sum(v) / len(v)), 'min %.4g max %.4g' % (min(v), max(v)))
PY
done
