#!/usr/bin/env bash
# C3 phase breakdown of the grid-wide kernel (tolerance mode, 1 / 2 CTAs per SM); C4: a / b as a * rcp(b) (S21_B4_RCPDIV build)
set -u
mkdir -p gpurun_out
{
for per in 1 2; do for rings in 400 2000; do
  echo "--- $rings rings, S21_GRID_CTAS=$per"; S21_GRID_CTAS=$per S21_PLAN_INFO=1 timeout 600 python scripts/run_c3.py $rings 5 2e-10 2>&1 | grep -E "s21 grid|rings=|second run" | cut -c1-420
done; done
} > gpurun_out/r02v_c3_phases.txt 2>&1
cat gpurun_out/r02v_c3_phases.txt
V=spice21_b200/variants/libspice21cu_rcpdiv.so
{
for B in 2048 256; do
  echo "--- B=$B default"; timeout 600 python scripts/run_c4.py $B 21 100 2>&1 | grep -E "^rep 1|rror"
  echo "--- B=$B rcpdiv"; S21_LIB=$V timeout 600 python scripts/run_c4.py $B 21 100 4 2>&1 | grep -E "^rep 1|rror|oracle"
done
} > gpurun_out/r02v_c4_rcpdiv.txt 2>&1
cat gpurun_out/r02v_c4_rcpdiv.txt | cut -c1-300
echo "== bsim4 tests on the rcpdiv build"; S21_LIB=$V timeout 900 python -m pytest tests -m gpu -q -k "bsim4 or c4 or golden" 2>&1 | tail -15 | cut -c1-300
