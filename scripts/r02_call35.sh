#!/usr/bin/env bash
# C4: two timing modes seen across runs (184 / 232 ms): separate processes, with and without a padded instance stride
set -u
mkdir -p gpurun_out
{
for k in 1 2 3; do for pad in 0 1; do for fast in 0 1; do
  echo "--- run $k S21_STRIDE_PAD=$pad S21_B4_FAST=$fast"; S21_STRIDE_PAD=$pad S21_B4_FAST=$fast timeout 300 python scripts/run_c4.py 2048 21 100 2>&1 | grep -E "^rep|rror" | cut -c1-140
done; done; done
echo "--- bench-like order: C1 sweep first, then C4 (in one process)"
for pad in 0 1; do S21_STRIDE_PAD=$pad timeout 600 python bench.py --config c4 --extras 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pad=$pad', d.get('value'), d.get('ms_per_step'), (d.get('rcp_division') or d.get('config',{})).__class__)
"; done
} > gpurun_out/r02C_c4_modes.txt 2>&1
cat gpurun_out/r02C_c4_modes.txt
