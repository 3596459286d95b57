#!/usr/bin/env bash
# cooperative kernel: CTA-blocked staging area against the batch-strided one (C4, default and reciprocal-division kernels)
set -u
mkdir -p gpurun_out
{
for k in 1 2; do for blk in 1 0; do for fast in 0 1; do
  echo "--- run $k S21_COOP_STAGE_BLOCKED=$blk S21_B4_FAST=$fast"; S21_COOP_STAGE_BLOCKED=$blk S21_B4_FAST=$fast timeout 300 python scripts/run_c4.py 2048 21 100 2>&1 | grep -E "^rep 1|rror" | cut -c1-140
done; done; done
for blk in 1 0; do
  echo "--- 256 MiB buffer allocated before the first transient, S21_COOP_STAGE_BLOCKED=$blk"; S21_COOP_STAGE_BLOCKED=$blk RUNC4_TORCH=2 timeout 300 python scripts/run_c4.py 2048 21 100 2>&1 | grep -E "^rep 1|rror" | cut -c1-140
  echo "--- B=256 S21_COOP_STAGE_BLOCKED=$blk"; S21_COOP_STAGE_BLOCKED=$blk timeout 300 python scripts/run_c4.py 256 21 100 2>&1 | grep -E "^rep 1|rror" | cut -c1-140
  echo "--- C4x N=375 B=2048 S21_COOP_STAGE_BLOCKED=$blk"; S21_COOP_STAGE_BLOCKED=$blk timeout 600 python scripts/run_c4x.py 2>&1 | grep -E "device_ms|rror" | head -3 | cut -c1-200
done
} > gpurun_out/r02P_c4_stage.txt 2>&1
cat gpurun_out/r02P_c4_stage.txt
timeout 900 python -m pytest tests -m gpu -q -k "bsim4 or c4 or coop or variants or sanitizer or ac_" 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-300 | head
