#!/usr/bin/env bash
# Round-2 GPU call 5: suite, C3 with tolerance-mode level schedules (400 and 2000 rings), bench dry run.
set -u
mkdir -p gpurun_out
echo "== gpu tests"; ( time timeout 1200 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -8
echo "== C3"
{
for exact in 1 0; do
  echo "--- S21_PLAN_EXACT=$exact, 400 rings (N=2803), 20 points"
  S21_PLAN_EXACT=$exact S21_PLAN_INFO=1 timeout 600 python scripts/run_c3.py 400 5 2e-10 2>&1 | grep -E "s21 plan\] mode=1 N=|rings=|second run"
done
echo "--- tolerance mode, 2000 rings (N=14003, 20 000 transistors), 20 points"
S21_PLAN_INFO=1 timeout 900 python scripts/run_c3.py 2000 5 2e-10 2>&1 | grep -E "s21 plan\]|rings=|second run|ring 0"
} 2>&1 | tee gpurun_out/r02e_c3.txt
echo "== bench dry run"
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err ); echo "rc=$?"; tail -5 gpurun_out/r02e_bench.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02e_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','scaling')}, 'kernel_ms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
    print('roofline', {k:v for k,v in d['roofline'].items() if k in ('bound','achieved','peak','frac')}, d['roofline']['hbm_algorithmic']['frac'])
    for k,v in (d.get('configs') or {}).items(): print(k, v.get('value'), v.get('unit'), 'ms', v.get('ms_per_transient') or v.get('ms_per_sweep') or v.get('ms_per_timepoint'), 'e2e', (v.get('e2e') or {}).get('value'), 'cpu', (v.get('cpu_baseline') or {}).get('value'), 'roof', (v.get('roofline') or {}).get('frac'), v.get('plan'))
    print('skipped', d.get('skipped'))
except Exception as e: print('parse failed', e)
PY
