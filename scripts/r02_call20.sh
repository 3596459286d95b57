#!/usr/bin/env bash
# Round-2 GPU call 20/21: suite and bench after the in-place re-pivot (resume launches).
set -u
mkdir -p gpurun_out
echo "== gpu suite"; ( time timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "^E  |passed|failed|FAILED|long ring|pivot repair" | cut -c1-400 | head -60 )
echo "== bench"; ( time timeout 900 python bench.py > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err ); echo "rc=$?"; tail -3 gpurun_out/r02r_bench.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02r_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','scaling','n_gpus')}, 'kernel_ms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['setup'])
    for k,v in (d.get('configs') or {}).items(): print(k, v.get('value'), v.get('unit'), 'ms', v.get('ms_per_transient') or v.get('ms_per_sweep') or v.get('ms_per_timepoint'), 'e2e', (v.get('e2e') or {}).get('value'), (v.get('e2e') or {}).get('ms'), 'first', v.get('first_call_wall_s'))
except Exception as e: print('parse failed', e)
PY
