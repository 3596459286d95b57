#!/usr/bin/env bash
# Round-2 GPU call 22: full suite, smoke, final N=1 bench, launch list + ncu of the C2 kernel and of the C4 transient kernel, C4x timing.
set -u
mkdir -p gpurun_out
echo "== gpu suite"; ( time timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 )
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; ( time timeout 900 python bench.py > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err ); echo "rc=$?"; tail -3 gpurun_out/r02s_bench.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02s_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','scaling','n_gpus')}, 'kernel_ms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
    for k,v in (d.get('configs') or {}).items(): print(k, v.get('value'), v.get('unit'), 'ms', v.get('ms_per_transient') or v.get('ms_per_sweep') or v.get('ms_per_timepoint'), 'e2e', (v.get('e2e') or {}).get('value'), (v.get('e2e') or {}).get('ms'))
except Exception as e: print('parse failed', e)
PY
echo "== reference arm"; ( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02s_bench_ref.json 2> gpurun_out/r02s_bench_ref.err ); tail -c 600 gpurun_out/r02s_bench_ref.json
echo "== launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02s_launches.csv python bench.py --steps 5 --warmup 3 --extras 0 > gpurun_out/r02s_launches.log 2>&1; echo "rc=$?"
echo "== ncu C2 kernel (device-resident step and the end-to-end step with rows written to the host)"
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name regex:k_jit --launch-skip 6 --launch-count 1 -o gpurun_out/r02s_c2 -f python bench.py --steps 2 --warmup 3 --extras 0 > gpurun_out/r02s_c2_ncu.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --kernel-name regex:k_jit --launch-skip 14 --launch-count 1 -o gpurun_out/r02s_c2_e2e -f python bench.py --steps 2 --warmup 3 --extras 0 > gpurun_out/r02s_c2_e2e_ncu.log 2>&1; echo "rc=$?"
echo "== ncu C4 transient kernel (2048 instances, 5 points)"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_coop -c 1 -s 1 -f -o gpurun_out/r02s_c4_2048 python scripts/run_c4.py 2048 21 5 > gpurun_out/r02s_c4_ncu.log 2>&1; echo "rc=$?"
timeout 900 ncu --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:k_coop -c 1 -s 1 --csv --log-file gpurun_out/r02s_c4_fp64ops.csv python scripts/run_c4.py 2048 21 5 > gpurun_out/r02s_c4_fp64ops.log 2>&1; echo "rc=$?"; grep -E "loads|rep 1" gpurun_out/r02s_c4_fp64ops.log | tail -2
echo "== C4x: 41 stages, 2 ICs, rbodymod = rgatemod = 1 (N = 375), 2048 and 256 instances, 50 points"
{
timeout 900 python scripts/run_c4x.py 41 20 2048 50 1 8 2>&1 | grep -E "^rep 1|oracle|max|plan|rror"
timeout 900 python scripts/run_c4x.py 41 20 256 50 1 0 2>&1 | grep -E "^rep 1|plan|rror"
timeout 900 python scripts/run_c4x.py 41 20 2048 50 0 8 2>&1 | grep -E "^rep 1|oracle|max|plan|rror"
} 2>&1 | tee gpurun_out/r02s_c4x.txt
