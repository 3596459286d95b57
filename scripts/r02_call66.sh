#!/usr/bin/env bash
set -u
timeout 600 python -m pytest tests -m gpu -q -k "ac_ or sweep_single" 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-300 | head
timeout 120 python scripts/run_c5.py 12500 2>&1 | grep -E "^rep 2" | cut -c1-200
