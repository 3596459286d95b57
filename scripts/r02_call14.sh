#!/usr/bin/env bash
# Round-2 GPU call 14: re-entry baseline — full GPU suite, smoke, bench (N = 1) with its launch list.
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "== gpu suite"; ( time timeout 1500 python -m pytest tests -m gpu -q ) 2>&1 | tail -15
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench"; ( time timeout 900 python bench.py > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err ); echo "rc=$?"; tail -3 gpurun_out/r02k_bench.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02k_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','scaling','n_gpus')}, 'kernel_ms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
    for k,v in (d.get('configs') or {}).items(): print(k, v.get('value'), v.get('unit'), 'ms', v.get('ms_per_transient') or v.get('ms_per_sweep') or v.get('ms_per_timepoint'), 'e2e', (v.get('e2e') or {}).get('value'))
except Exception as e: print('parse failed', e)
PY
echo "== launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02k_launches.csv python bench.py --steps 5 --warmup 3 --extras 0 > gpurun_out/r02k_launches.log 2>&1; echo "rc=$?"
