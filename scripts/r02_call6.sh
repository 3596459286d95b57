#!/usr/bin/env bash
# Round-2 GPU call 6: adaptive-transient tests, full suite, cost of the pivot-health test in the team kernel.
set -u
mkdir -p gpurun_out
echo "== new tests"; timeout 600 python -m pytest tests -m gpu -x -q -s -k "adaptive or grid_kernel" 2>&1 | tail -12
echo "== gpu suite"; ( time timeout 1200 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6
echo "== C2 bench with / without the pivot-health test in the generated kernel"
for hflag in 1 0; do
  S21_PIVOT_HEALTH=$hflag timeout 300 python bench.py --steps 20 --warmup 5 --extras 0 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('health=$hflag', 'ms_per_step', round(d['ms_per_step'],5), 'kernel_ms', round(d['roofline']['kernel_ms'],5), 'e2e_ms', round(d['e2e']['ms_per_step'],5))"
done 2>&1 | tee gpurun_out/r02f_health_cost.txt
