#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== failing tests, verbose"; timeout 900 python -m pytest tests -m gpu -q -s -k "adaptive or aids or outlier or long_ring or golden" 2>&1 | grep -E "^E  |passed|failed|Error|assert|pivot repair|grid kernel" | cut -c1-300 | head -60
echo "== gpu suite"; ( time timeout 1500 python -m pytest tests -m gpu -q ) 2>&1 | tail -12
