#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -k "time_varying or c4x or grid_kernel" 2>&1 | grep -E "^E  |passed|failed|FAILED|^tests/test_gpu.py:[0-9]+" | cut -c1-600 | head -60
