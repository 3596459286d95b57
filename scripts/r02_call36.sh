#!/usr/bin/env bash
# does the 10 Hz nvidia-smi poll of bench.py perturb the C4 launches? bench --config c4 with / without it, twice each; then 2-GPU bench
set -u
mkdir -p gpurun_out
{
for k in 1 2; do for smp in 1 0; do
  S21_BENCH_SAMPLE_EXTRAS=$smp timeout 600 python bench.py --config c4 --extras 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d.get('detail',{}); print('run $k sampler=$smp default ms', d.get('ms_per_step'), 'rcp ms', (r.get('rcp_division') or {}).get('ms_per_transient'), 'clocks', d.get('clocks'))
"; done; done
} > gpurun_out/r02D_c4_sampler.txt 2>&1
cat gpurun_out/r02D_c4_sampler.txt
