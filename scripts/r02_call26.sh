#!/usr/bin/env bash
# grid-wide kernel, tolerance mode: long sums by the warp / the grid (C3)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "grid_kernel or sanitizer or c3 or large" 2>&1 | tail -5 | cut -c1-300
{
echo "--- 400 rings, tolerance mode"; S21_PLAN_INFO=1 timeout 300 python scripts/run_c3.py 400 5 2e-10 2>&1 | grep -v "symbolic\]"
echo "--- 2000 rings (N=14003), tolerance mode"; S21_PLAN_INFO=1 timeout 600 python scripts/run_c3.py 2000 5 2e-10 2>&1 | grep -v "symbolic\]"
echo "--- 100 rings against the oracle"; timeout 600 python scripts/run_c3.py 100 5 1e-10 oracle 2>&1 | grep -v "symbolic\]"
} > gpurun_out/r02u_c3.txt 2>&1
cat gpurun_out/r02u_c3.txt | cut -c1-400
