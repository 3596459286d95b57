#!/usr/bin/env bash
# direct kernel: four independent operations in flight per pivot step — bit-identity tests, AC tests, C5 timing
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "ac or variants or bit_identical or direct or known or dcop_matches or repivot or adaptive" 2>&1 | grep -E "^E  |passed|failed|^tests/test_gpu.py:[0-9]+|FAILED" | cut -c1-400 | head -30
{
timeout 600 python scripts/run_c5.py 2>&1 | grep -E "^rep|rror" | cut -c1-300
} > gpurun_out/r02K_c5.txt 2>&1
cat gpurun_out/r02K_c5.txt
