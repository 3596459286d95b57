#!/usr/bin/env bash
# team kernel with the compact staging area: timing (C1 sweep, C2), then the whole suite + smoke + bench
set -u
mkdir -p gpurun_out
{
timeout 600 python scripts/sweep_tran.py 4 8192,4096,16384 2>&1 | cut -c1-170
echo "--- committed state forced into HBM on top (S21_TEAM_SOPG=1)"; S21_TEAM_SOPG=1 timeout 600 python scripts/sweep_tran.py 4 8192 2>&1 | cut -c1-170
timeout 600 python scripts/sweep_batch.py 2>&1 | tail -12 | cut -c1-200
} > gpurun_out/r02X_team_compact.txt 2>&1
cat gpurun_out/r02X_team_compact.txt
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |passed|failed|^tests/test_gpu.py:[0-9]+|FAILED" | cut -c1-500 | head -30 ) 2>&1 | tail -34
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02X_bench.json 2> gpurun_out/r02X_bench.err; echo "rc=$?"; tail -c 300 gpurun_out/r02X_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02X_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'])
for k, v in d['configs'].items():
    print(k, v.get('value'), v.get('unit'), v.get('ms_per_transient'), v.get('ms_per_timepoint'), 'e2e', (v.get('e2e') or {}).get('ms'), (v.get('rcp_division') or {}).get('ms_per_transient'))
PY
