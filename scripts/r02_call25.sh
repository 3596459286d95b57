#!/usr/bin/env bash
# full GPU suite at HEAD + smoke + both bench arms (validation after the f2 commit)
set -u
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=25 ) > gpurun_out/r02t_pytest.log 2>&1
tail -40 gpurun_out/r02t_pytest.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err; tail -c 600 gpurun_out/r02t_bench.json
