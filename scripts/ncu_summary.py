"""Summarise an ncu raw+source CSV pair: key metrics and the hot SASS regions (by executed instructions)."""
import csv, sys
raw, src, units_per_launch = sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct',
        'sm__inst_executed_pipe_fp64.avg.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__average_warps_issue_stalled',
        'sm__warps_active.avg.pct', 'sm__inst_executed_pipe_alu.avg.pct', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'sass__inst_executed_local', 'dram__throughput.avg.pct',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct', 'launch__occupancy_limit', 'launch__grid_size', 'launch__block_size', 'lts__t_sector_hit_rate.pct']
for h, u, v in zip(hdr, units, vals):
  if any((h + ' ').startswith(w) or w in h for w in want) and 'Not Issued' not in h and '.max' not in h and '.min' not in h and '.sum.p' not in h:
    try:
      if 'stalled' in h and float(v) < 0.3: continue
    except ValueError: pass
    print(f"{h} [{u}] = {v}")
rows = list(csv.reader(open(src)))
hdr = rows[1]
ia, isrc = hdr.index('Instructions Executed'), hdr.index('Source')
data = []
for r in rows[2:]:
  try: data.append((r[isrc], int(r[ia])))
  except Exception: pass
tot = sum(d[1] for d in data)
print(f"SASS instructions {len(data)}, warp-instructions executed {tot} = {tot / units_per_launch:.1f} per unit")
prev, start, out = None, 0, []
for i, (sr, n) in enumerate(data):
  if prev is None or abs(n - prev) > 0.02 * max(n, prev, 1):
    if prev is not None: out.append((start, i - 1, prev, i - start))
    start, prev = i, n
out.append((start, len(data) - 1, prev, len(data) - start))
for s, e, n, c in out:
  if n * c > 0.008 * tot:
    print(f"{s:5d}-{e:5d} exec/inst={n:10d} ninstr={c:4d} share={n * c / tot * 100:5.1f}%  {data[s][0][:60]}")
