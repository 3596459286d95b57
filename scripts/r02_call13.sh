#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
{
timeout 600 python scripts/run_c4x.py 41 0 256 50 0 2 2>&1 | grep -E "^rep 1|oracle|max|plan|rror"
timeout 600 python scripts/run_c4x.py 101 26 256 50 0 2 2>&1 | grep -E "^rep 1|oracle|max|plan|rror"
timeout 900 python scripts/run_c4x.py 101 26 256 20 1 1 2>&1 | grep -E "^rep|oracle|max|plan|rror"
} 2>&1 | tee gpurun_out/r02j_c4x.txt
echo "== aids test"; timeout 600 python -m pytest tests -m gpu -q -x -k "aids" 2>&1 | tail -3
