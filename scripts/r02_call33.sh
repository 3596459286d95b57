#!/usr/bin/env bash
# full GPU suite + C3 (512-thread CTAs) + bench
set -u
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "^E  |passed|failed|fast division|^tests/test_gpu.py:[0-9]+|FAILED" | cut -c1-500 | head -40 ) 2>&1 | tail -45
{
for th in 512; do for rings in 2000; do
  echo "--- $rings rings, $th threads per CTA"; S21_GRID_THREADS=$th S21_PLAN_INFO=1 timeout 600 python scripts/run_c3.py $rings 5 2e-10 2>&1 | grep -E "s21 grid|second run" | tail -2 | cut -c1-420
done; done
} > gpurun_out/r02A_c3_phases.txt 2>&1
cat gpurun_out/r02A_c3_phases.txt
timeout 900 python bench.py > gpurun_out/r02A_bench.json 2> gpurun_out/r02A_bench.err; tail -c 300 gpurun_out/r02A_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02A_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])
for k, v in d['configs'].items():
    print(k, v.get('value'), v.get('unit'), v.get('ms_per_transient'), v.get('ms_per_timepoint'), (v.get('rcp_division') or {}).get('ms_per_transient'))
PY
