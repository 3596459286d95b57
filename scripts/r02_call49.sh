#!/usr/bin/env bash
# full GPU suite (OP repair fix, transient hand-back / resume, tightened c4x / fast-division tests) + smoke + bench
set -u
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |passed|failed|^tests/test_gpu.py:[0-9]+|FAILED" | cut -c1-500 | head -40 ) 2>&1 | tail -45
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02N_bench.json 2> gpurun_out/r02N_bench.err; tail -c 300 gpurun_out/r02N_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02N_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])
for k, v in d['configs'].items():
    print(k, v.get('value'), v.get('unit'), v.get('ms_per_transient'), v.get('ms_per_timepoint'), 'e2e', (v.get('e2e') or {}).get('ms'), (v.get('rcp_division') or {}).get('ms_per_transient'))
PY
