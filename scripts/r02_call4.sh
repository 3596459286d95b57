#!/usr/bin/env bash
# Round-2 GPU call 4: full GPU suite with the new tests, ncu of the default C2 kernel (warp-private team kernel), bench.py dry run.
set -u
mkdir -p gpurun_out
echo "== gpu tests"; ( time timeout 1200 python -m pytest tests -m gpu -x -q -s -k "outlier or referee or high_gain" ) 2>&1 | tail -12
( time timeout 1200 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -8
echo "== ncu C2 kernel (one launch of the bench's kernel)"
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:k_jit --launch-skip 6 --launch-count 1 \
  -o gpurun_out/r02d_c2 -f python bench.py --steps 2 --warmup 3 --extras 0 > gpurun_out/r02d_c2_ncu.log 2>&1; tail -2 gpurun_out/r02d_c2_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02d_launches.csv \
  python bench.py --steps 2 --warmup 3 --extras 0 > gpurun_out/r02d_launches.log 2>&1; tail -3 gpurun_out/r02d_launches.csv | cut -c1-300
echo "== bench dry run"
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err ); echo "rc=$?"; tail -5 gpurun_out/r02d_bench.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02d_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','scaling')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'setup', d.get('setup'))
    print('roofline', {k:v for k,v in d['roofline'].items() if k in ('bound','achieved','peak','frac')}, d['roofline']['hbm_algorithmic']['frac'])
    for k,v in (d.get('configs') or {}).items(): print(k, v.get('value'), v.get('unit'), 'ms', v.get('ms_per_transient') or v.get('ms_per_sweep') or v.get('ms_per_timepoint'), 'e2e', (v.get('e2e') or {}).get('value'), 'cpu', (v.get('cpu_baseline') or {}).get('value'), 'roof', (v.get('roofline') or {}).get('frac'))
    print('skipped', d.get('skipped'))
except Exception as e: print('parse failed', e)
PY
