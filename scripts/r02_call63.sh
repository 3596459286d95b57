#!/usr/bin/env bash
# committed state in HBM (automatic choice): C1 sweep timing, its test, transient tests, sanitizer, then the whole suite + bench
set -u
mkdir -p gpurun_out
timeout 600 python scripts/sweep_tran.py 4 8192 2>&1 | cut -c1-170
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |passed|failed|^tests/test_gpu.py:[0-9]+|FAILED" | cut -c1-500 | head -30 ) 2>&1 | tail -34
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02AB_bench.json 2> gpurun_out/r02AB_bench.err; echo "rc=$?"; tail -c 300 gpurun_out/r02AB_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02AB_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'])
print('tran', d.get('tran', {}).get('value'), d.get('tran', {}).get('ms_per_transient'))
for k, v in d['configs'].items():
    print(k, v.get('value'), v.get('unit'), v.get('ms_per_transient'), v.get('ms_per_timepoint'), 'e2e', (v.get('e2e') or {}).get('ms'), (v.get('rcp_division') or {}).get('ms_per_transient'))
PY
