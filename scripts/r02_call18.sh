#!/usr/bin/env bash
# Round-2 GPU call 18: host-rows path (tests, e2e trace on / off), C3 first-call time, full suite, bench.
set -u
mkdir -p gpurun_out
echo "== focused tests"; timeout 1500 python -m pytest tests -m gpu -q -x -k "team or view or packed or sweep or sanitizer or c4x or monte_carlo or ragged or variants_bit" 2>&1 | tail -6
echo "== e2e trace (host rows on / off)"
for hr in 1 0 1 0; do echo "--- S21_HOST_ROWS=$hr"; S21_HOST_ROWS=$hr S21_TRACE_E2E=1 timeout 300 python scripts/e2e_trace.py 2>&1 | tail -3; done | tee gpurun_out/r02o_e2e_trace.txt
for hr in 1 0; do echo "--- S21_HOST_ROWS=$hr (no trace)"; S21_HOST_ROWS=$hr timeout 300 python scripts/e2e_trace.py 2>&1 | tail -1; done | tee -a gpurun_out/r02o_e2e_trace.txt
echo "== C3 400 rings, phases"
S21_PLAN_INFO=1 timeout 900 python scripts/run_c3.py 400 5 2e-10 > gpurun_out/r02o_c3_phases.txt 2>&1; grep -E "s21 tran|rings=|second" gpurun_out/r02o_c3_phases.txt | cut -c1-260
echo "== gpu suite"; ( time timeout 1800 python -m pytest tests -m gpu -q ) 2>&1 | tail -8
echo "== bench"; ( time timeout 900 python bench.py > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err ); echo "rc=$?"; tail -3 gpurun_out/r02o_bench.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02o_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','scaling','n_gpus')}, 'kernel_ms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
    for k,v in (d.get('configs') or {}).items(): print(k, v.get('value'), v.get('unit'), 'ms', v.get('ms_per_transient') or v.get('ms_per_sweep') or v.get('ms_per_timepoint'), 'e2e', (v.get('e2e') or {}).get('value'), 'first', v.get('first_call_wall_s'))
except Exception as e: print('parse failed', e)
PY
