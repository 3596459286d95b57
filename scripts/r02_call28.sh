#!/usr/bin/env bash
# C3 with the residual over the rows of A proper; C4: rcp division + FMA contraction variant; what the two failing tests of the rcp build see
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "grid_kernel or sanitizer" 2>&1 | tail -3 | cut -c1-300
{
for rings in 400 2000; do
  echo "--- $rings rings"; S21_PLAN_INFO=1 timeout 600 python scripts/run_c3.py $rings 5 2e-10 2>&1 | grep -E "s21 grid|rings=|second run" | cut -c1-420
done
echo "--- 100 rings against the oracle"; timeout 600 python scripts/run_c3.py 100 5 1e-10 oracle 2>&1 | grep -E "oracle|rings="
} > gpurun_out/r02w_c3_phases.txt 2>&1
cat gpurun_out/r02w_c3_phases.txt
{
for v in rcpdiv rcpfma; do for B in 2048 256; do
  echo "--- B=$B $v"; S21_LIB=spice21_b200/variants/libspice21cu_$v.so timeout 600 python scripts/run_c4.py $B 21 100 8 2>&1 | grep -E "^rep 1|rror|oracle"
done; done
} > gpurun_out/r02w_c4_fast.txt 2>&1
cat gpurun_out/r02w_c4_fast.txt | cut -c1-300
for v in rcpdiv rcpfma; do
echo "== bsim4 tests on the $v build"; S21_LIB=spice21_b200/variants/libspice21cu_$v.so timeout 900 python -m pytest tests -m gpu -q -k "bsim4 or c4 or golden" 2>&1 | grep -E "^E  |passed|failed|^tests/test_gpu.py:[0-9]+" | cut -c1-400 | head -40
done
