#!/usr/bin/env bash
# Round-2 GPU call 2: new tests (sweep, cache, full-size C2), ncu of the C4 (Bsim4) transient kernel, C3 baseline.
set -u
mkdir -p gpurun_out
echo "== new tests"; ( time timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -8
echo "== ncu C4 B=2048 (source-level)"
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name regex:k_coop --launch-skip 1 --launch-count 1 \
  -o gpurun_out/r02b_c4_2048 -f python scripts/run_c4.py 2048 21 5 > gpurun_out/r02b_c4_2048.log 2>&1; tail -3 gpurun_out/r02b_c4_2048.log
echo "== ncu C4 B=256"
timeout 600 ncu --set full --clock-control none --kernel-name regex:k_coop --launch-skip 1 --launch-count 1 \
  -o gpurun_out/r02b_c4_256 -f python scripts/run_c4.py 256 21 5 > gpurun_out/r02b_c4_256.log 2>&1; tail -3 gpurun_out/r02b_c4_256.log
echo "== fp64 op counts C4 B=2048"
timeout 600 ncu --clock-control none --kernel-name regex:k_coop --launch-skip 1 --launch-count 1 --csv \
  --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum \
  --log-file gpurun_out/r02b_c4_fp64ops.csv python scripts/run_c4.py 2048 21 5 > gpurun_out/r02b_c4_fp64ops.log 2>&1; tail -12 gpurun_out/r02b_c4_fp64ops.csv
echo "== C3 400 rings"
S21_PLAN_INFO=1 timeout 600 python scripts/run_c3.py 400 5 2e-10 2>&1 | tail -12 | tee gpurun_out/r02b_c3_400.txt
ls -la gpurun_out/
