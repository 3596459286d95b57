#!/usr/bin/env bash
# AC sweeps below the 16 384-point switch to the thread-per-point kernel: which kernel wins at 12 500 / 6250 / 2048 points
set -u
mkdir -p gpurun_out
{
for n in 12500 6250 2048; do
  echo "--- $n points default"; timeout 300 python scripts/run_c5.py $n 2>&1 | grep -E "^rep 2" | cut -c1-220
  echo "--- $n points S21_KERNEL=direct"; S21_KERNEL=direct timeout 300 python scripts/run_c5.py $n 2>&1 | grep -E "^rep 2" | cut -c1-220
done
} > gpurun_out/r02AD_c5_small.txt 2>&1
cat gpurun_out/r02AD_c5_small.txt
