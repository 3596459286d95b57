#!/usr/bin/env bash
# Round-2 GPU call 17: where C3's first call spends 170 s; e2e trace of the C2 step; new tests.
set -u
mkdir -p gpurun_out
echo "== C3 400 rings, phases"
S21_PLAN_INFO=1 timeout 900 python scripts/run_c3.py 400 5 2e-10 > gpurun_out/r02n_c3_phases.txt 2>&1; grep -E "s21 tran|s21 plan|rings=|second" gpurun_out/r02n_c3_phases.txt | cut -c1-260
echo "== e2e trace"
S21_TRACE_E2E=1 timeout 300 python scripts/e2e_trace.py > gpurun_out/r02n_e2e_trace.txt 2>&1; tail -8 gpurun_out/r02n_e2e_trace.txt
echo "== new tests"; timeout 1500 python -m pytest tests -m gpu -q -x -k "invariants or c4x or override_specs or outlier or long_ring or aids" 2>&1 | tail -8
