#!/usr/bin/env bash
# opt-in rcp-division kernel inside the library (S21_B4_FAST=1), C3 after the residual lists, bench line
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "fast_division or grid_kernel or sanitizer or c4_bsim4" -s 2>&1 | grep -E "^E  |passed|failed|fast division|^tests/test_gpu.py:[0-9]+" | cut -c1-500 | head -40
{
for rings in 400 2000; do
  echo "--- $rings rings"; S21_PLAN_INFO=1 timeout 600 python scripts/run_c3.py $rings 5 2e-10 2>&1 | grep -E "s21 grid|rings=|second run" | cut -c1-420
done
} > gpurun_out/r02x_c3_phases.txt 2>&1
cat gpurun_out/r02x_c3_phases.txt
timeout 900 python bench.py > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err; tail -c 300 gpurun_out/r02x_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02x_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])
for k, v in d['configs'].items():
    print(k, v.get('value'), v.get('unit'), v.get('ms_per_transient'), v.get('ms_per_timepoint'), v.get('rcp_division'))
PY
