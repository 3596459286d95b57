#!/usr/bin/env bash
# last check of HEAD after the AC kernel switch moved: whole GPU suite + smoke
set -u
( time timeout 400 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |passed|failed|FAILED" | cut -c1-300 | head -20 ) 2>&1 | tail -24
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
