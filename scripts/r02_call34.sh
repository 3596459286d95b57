#!/usr/bin/env bash
# waveforms / AC results transposed on the device: tests that read them + bench (e2e of C1 / C5)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "ac or tran or golden or sanitizer or sweep or adaptive or time_varying" 2>&1 | grep -E "^E  |passed|failed|^tests/test_gpu.py:[0-9]+|FAILED" | cut -c1-400 | head -30
timeout 900 python bench.py > gpurun_out/r02B_bench.json 2> gpurun_out/r02B_bench.err; tail -c 300 gpurun_out/r02B_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02B_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])
for k, v in d['configs'].items():
    print(k, v.get('value'), v.get('unit'), v.get('ms_per_transient'), v.get('ms_per_timepoint'), 'e2e', (v.get('e2e') or {}).get('ms'), (v.get('rcp_division') or {}).get('ms_per_transient'))
PY
