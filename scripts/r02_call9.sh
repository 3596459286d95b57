#!/usr/bin/env bash
# Round-2 GPU call 9: re-pivot-on-demand (stop + continue), aids, adaptive; full suite; C2 bench.
set -u
mkdir -p gpurun_out
timeout 300 python scripts/dbg_ring.py 2>&1 | tail -30
echo "== new tests"; timeout 900 python -m pytest tests -m gpu -x -q -s -k "adaptive or grid_kernel or aids or outlier or long_ring" 2>&1 | tail -15
echo "== gpu suite"; ( time timeout 1500 python -m pytest tests -m gpu -q ) 2>&1 | tail -12
echo "== C2 bench"
for hflag in 1 0; do
  S21_PIVOT_HEALTH=$hflag timeout 300 python bench.py --steps 20 --warmup 5 --extras 0 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('health=$hflag', 'ms_per_step', round(d['ms_per_step'],5), 'kernel_ms', round(d['roofline']['kernel_ms'],5), 'e2e_ms', round(d['e2e']['ms_per_step'],5))"
done 2>&1 | tee gpurun_out/r02h_health_cost.txt
