X="$1"; shift
QCD_NVCC_EXTRA="$X" python -m qcd_b200.build > /dev/null 2>&1 || { echo build failed; exit 1; }
timeout 120 python __graft_entry__.py smoke 2>&1 | grep -E "Error|error|assert" | head -3
for w in "$@"; do timeout 200 python bench.py --no-cpu --e2e-steps 3 --steps 400 --warmup 20 --workload $w 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$X]', d[\"config\"][\"workload\"], \"%.3e\"%d[\"value\"], \"frac=%.3f\"%d[\"roofline\"][\"frac\"], d[\"roofline\"][\"kernel\"][:30])"; done
