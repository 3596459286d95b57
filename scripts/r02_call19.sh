#!/usr/bin/env bash
# Round-2 GPU call 19: zero-copy parameter reads (S21_HOST_PARAMS) on / off, c4x test.
set -u
mkdir -p gpurun_out
echo "== c4x test"; timeout 900 python -m pytest tests -m gpu -q -x -k "c4x" 2>&1 | tail -4
echo "== e2e: S21_HOST_PARAMS on / off"
for hp in 1 0 1 0; do echo "--- S21_HOST_PARAMS=$hp"; S21_HOST_PARAMS=$hp timeout 300 python scripts/e2e_trace.py 2>&1 | tail -1; done | tee gpurun_out/r02p_host_params.txt
for hp in 1 0; do echo "--- S21_HOST_PARAMS=$hp (trace)"; S21_HOST_PARAMS=$hp S21_TRACE_E2E=1 timeout 300 python scripts/e2e_trace.py 2>&1 | tail -3; done | tee -a gpurun_out/r02p_host_params.txt
echo "== correctness with S21_HOST_PARAMS=1"; S21_HOST_PARAMS=1 timeout 900 python -m pytest tests -m gpu -q -x -k "view or monte_carlo" 2>&1 | tail -3
