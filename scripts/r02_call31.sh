#!/usr/bin/env bash
# C3 per-level timing of the grid-wide kernel; C4 default kernel re-measured (184 vs 232 ms between two builds / boxes)
set -u
mkdir -p gpurun_out
{
for rings in 2000 400; do
  echo "--- $rings rings"; S21_PLAN_INFO=1 timeout 600 python scripts/run_c3.py $rings 5 2e-10 2>&1 | grep -E "s21 grid|second run" | tail -6 | cut -c1-700
done
} > gpurun_out/r02z_c3_levels.txt 2>&1
cat gpurun_out/r02z_c3_levels.txt
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv,noheader
for k in 1 2; do
  echo "--- B=2048 default (run $k)"; timeout 600 python scripts/run_c4.py 2048 21 100 2>&1 | grep -E "^rep|rror"
  echo "--- B=2048 S21_B4_FAST=1 (run $k)"; S21_B4_FAST=1 timeout 600 python scripts/run_c4.py 2048 21 100 2>&1 | grep -E "^rep|rror"
done
echo "--- B=256 default"; timeout 600 python scripts/run_c4.py 256 21 100 2>&1 | grep -E "^rep 1|rror"
echo "--- B=256 S21_B4_FAST=1"; S21_B4_FAST=1 timeout 600 python scripts/run_c4.py 256 21 100 2>&1 | grep -E "^rep 1|rror"
} > gpurun_out/r02z_c4_repeat.txt 2>&1
cat gpurun_out/r02z_c4_repeat.txt | cut -c1-250
