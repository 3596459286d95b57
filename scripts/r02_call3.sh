#!/usr/bin/env bash
# Round-2 GPU call 3: Bsim4 evaluation variants on C4 (direct parameter blocks, out-of-line division / math), bench.py dry run.
set -u
mkdir -p gpurun_out
run_c4() { # label, env..., B
  echo "--- $1"; shift
  env "$@" timeout 300 python scripts/run_c4.py 2>&1 | grep -E "^rep 1|Error|error" 
}
{
echo "C4 (21-stage Bsim4 ring, 100 points) device ms by build variant"
for B in 2048 256; do
  echo "== B=$B"
  for v in "generic-par:S21_PAR_DIRECT=0" "direct-par:S21_PAR_DIRECT=1" "direct+nidiv:S21_LIB=$PWD/spice21_b200/variants/libspice21cu_nidiv.so" "direct+nidiv+nimath:S21_LIB=$PWD/spice21_b200/variants/libspice21cu_nimath.so"; do
    echo "--- ${v%%:*}"; env "${v#*:}" timeout 300 python scripts/run_c4.py $B 21 100 2>&1 | grep -E "^rep 1|rror"
  done
done
} 2>&1 | tee gpurun_out/r02c_c4_variants.txt
echo "== bit-exactness of the variants (bsim4 tests)"
for lib in nidiv nimath; do S21_LIB=$PWD/spice21_b200/variants/libspice21cu_$lib.so timeout 600 python -m pytest tests/test_gpu.py -m gpu -x -q -k "bsim4 or golden" 2>&1 | tail -2; done
echo "== gpu tests (default lib)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== bench dry run"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "rc=$?"; tail -5 gpurun_out/r02c_bench.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02c_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','scaling')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'setup', d.get('setup'))
    for k,v in (d.get('configs') or {}).items(): print(k, v.get('value'), v.get('unit'), 'ms', v.get('ms_per_transient') or v.get('ms_per_sweep') or v.get('ms_per_timepoint'), 'e2e', v.get('e2e'), 'cpu', (v.get('cpu_baseline') or {}).get('value'), 'roof', (v.get('roofline') or {}).get('frac'))
    print('skipped', d.get('skipped'))
except Exception as e: print('parse failed', e)
PY
