#!/usr/bin/env bash
# team kernel with the committed device state in its HBM column (two CTAs per SM on the C1-circuit sweep): timing, parity, sanitizer
set -u
mkdir -p gpurun_out
{
for sg in 1 0; do
  echo "--- S21_TEAM_SOPG=$sg"; S21_TEAM_SOPG=$sg timeout 600 python scripts/sweep_tran.py 4 8192,4096,16384 2>&1 | cut -c1-200
done
echo "--- default (automatic choice)"; timeout 600 python scripts/sweep_tran.py 4 8192,2048 2>&1 | cut -c1-200
} > gpurun_out/r02U_c1_sopg.txt 2>&1
cat gpurun_out/r02U_c1_sopg.txt
timeout 900 python -m pytest tests -m gpu -q -k "golden or tran or team or sanitizer or time_varying or adaptive" 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-300 | head
