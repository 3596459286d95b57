#!/usr/bin/env bash
# level-scheduled per-thread AC kernel: tests, C5 timing against the plain thread kernel
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "ac_ or sanitizer" 2>&1 | grep -E "^E  |passed|failed|^tests/test_gpu.py:[0-9]+|FAILED" | cut -c1-400 | head -30
{
echo "--- default (direct-levels)"; timeout 600 python scripts/run_c5.py 2>&1 | grep -E "^rep|rror" | cut -c1-300
echo "--- S21_AC_LEVELS=0 (plain thread kernel)"; S21_AC_LEVELS=0 timeout 600 python scripts/run_c5.py 2>&1 | grep -E "^rep|rror" | cut -c1-300
} > gpurun_out/r02L_c5.txt 2>&1
cat gpurun_out/r02L_c5.txt
