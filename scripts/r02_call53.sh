#!/usr/bin/env bash
# bench.py with the NVML clock sampler; sanitizer tests with the extended target
set -u
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02R_bench.json 2> gpurun_out/r02R_bench.err; echo "rc=$?"; tail -c 300 gpurun_out/r02R_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02R_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'])
for k, v in d['configs'].items():
    print(k, v.get('value'), v.get('unit'), v.get('ms_per_transient'), v.get('ms_per_timepoint'), 'e2e', (v.get('e2e') or {}).get('ms'), (v.get('rcp_division') or {}).get('ms_per_transient'))
PY
timeout 900 python -m pytest tests -m gpu -q -k "sanitizer" 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-400 | head
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
