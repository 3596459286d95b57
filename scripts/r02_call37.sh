#!/usr/bin/env bash
# why bench.py times the C4 launch at 218 ms and run_c4.py at 180 ms: what bench does around the batch, one thing at a time;
# and the L1 / shared-memory carveout of the spill-heavy kernel
set -u
mkdir -p gpurun_out
{
echo "--- plain"; timeout 300 python scripts/run_c4.py 2048 21 100 2>&1 | grep -E "^rep 1|rror" | cut -c1-120
echo "--- torch stream"; RUNC4_TORCH=1 timeout 300 python scripts/run_c4.py 2048 21 100 2>&1 | grep -E "^rep 1|rror" | cut -c1-120
echo "--- torch stream + 256 MB buffer"; RUNC4_TORCH=2 timeout 300 python scripts/run_c4.py 2048 21 100 2>&1 | grep -E "^rep 1|rror" | cut -c1-120
echo "--- forced parameter upload"; RUNC4_FORCE=1 timeout 300 python scripts/run_c4.py 2048 21 100 2>&1 | grep -E "^rep 1|rror" | cut -c1-120
echo "--- 4 reps"; RUNC4_REPS=4 timeout 300 python scripts/run_c4.py 2048 21 100 2>&1 | grep -E "^rep|rror" | cut -c1-120
for cv in 0 25 50 100; do
echo "--- S21_COOP_CARVEOUT=$cv"; S21_COOP_CARVEOUT=$cv timeout 300 python scripts/run_c4.py 2048 21 100 2>&1 | grep -E "^rep 1|rror" | cut -c1-120
echo "--- S21_COOP_CARVEOUT=$cv S21_B4_FAST=1"; S21_B4_FAST=1 S21_COOP_CARVEOUT=$cv timeout 300 python scripts/run_c4.py 2048 21 100 2>&1 | grep -E "^rep 1|rror" | cut -c1-120
done
} > gpurun_out/r02E_c4_modes.txt 2>&1
cat gpurun_out/r02E_c4_modes.txt
