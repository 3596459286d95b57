#!/usr/bin/env bash
# C4 at the longest ring the reference's Newton loop converges: 41 stages, 2 ICs, plain cards (N = 47), 2048 instances x 100 points
set -u
mkdir -p gpurun_out
{
for B in 2048 256; do
echo "--- B=$B 41 stages default"; RUNC4_IC_EVERY=20 timeout 600 python scripts/run_c4.py $B 41 100 16 2>&1 | grep -E "^rep|rror|oracle|stats" | cut -c1-300
echo "--- B=$B 41 stages S21_B4_FAST=1"; S21_B4_FAST=1 RUNC4_IC_EVERY=20 timeout 600 python scripts/run_c4.py $B 41 100 16 2>&1 | grep -E "^rep|rror|oracle" | cut -c1-200
done
} > gpurun_out/r02I_c4_41.txt 2>&1
cat gpurun_out/r02I_c4_41.txt
