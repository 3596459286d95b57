#!/usr/bin/env bash
# Round-2 GPU call 23: time-varying sources + Bsim4 cards on the wire (f2), c4x, sanitizer, then the full suite.
set -u
mkdir -p gpurun_out
echo "== f2 / c4x tests"; timeout 1500 python -m pytest tests -m gpu -q -x -k "time_varying or on_the_wire or c4x" 2>&1 | tail -12
echo "== gpu suite"; ( time timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 )
echo "== bench (C2 only, check for regressions from the Env change)"; timeout 600 python bench.py --extras 0 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', round(d['ms_per_step'],5), 'kernel_ms', round(d['roofline']['kernel_ms'],5), 'e2e_ms', round(d['e2e']['ms_per_step'],5), 'tran', d.get('tran',{}).get('ms_per_transient'))"
