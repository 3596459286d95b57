#!/usr/bin/env bash
# Round-2 GPU call 1: FP64 peak, the experiments round 1 left unmeasured (warp-private team kernel, BSIM4 divisions through
# scalar.h), and the full GPU suite. Every step under its own timeout.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 120 scripts/micro/fp64_peak | tee gpurun_out/fp64_peak.json
echo "== accept (warp-private)"; S21_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu.py -m gpu -x -q -k "warp_private" 2>&1 | tail -5
echo "== full suite"; ( time timeout 900 python -m pytest tests -m gpu -x -q --durations=8 ) 2>&1 | tail -16
echo "== wp sweep"
for wp in 0 1; do for gi in 32 16; do
  echo "WP=$wp GI=$gi"; S21_TEAM_WP=$wp S21_TEAM_GI=$gi timeout 120 python scripts/sweep_batch.py jitteam:2,jitteam:4,jitteam:8 1024,2048,4096,8192,65536
done; done 2>&1 | tee gpurun_out/r02_wp_sweep.txt
for wp in 0 1; do echo "WP=$wp"; S21_TEAM_WP=$wp timeout 300 python scripts/sweep_tran.py 4,2 8192; done 2>&1 | tee gpurun_out/r02_wp_tran.txt
S21_TEAM_WP=1 timeout 300 python bench.py > gpurun_out/r02_bench_wp.json 2> gpurun_out/r02_bench_wp.err; cut -c1-400 gpurun_out/r02_bench_wp.json
echo "== sdiv"
timeout 600 python scripts/run_c4.py 2048 21 100 2>&1 | tail -4 | tee gpurun_out/r02_c4_default.txt
S21_LIB=$PWD/spice21_b200/libspice21cu_sdiv.so timeout 600 python scripts/run_c4.py 2048 21 100 2>&1 | tail -4 | tee gpurun_out/r02_c4_sdiv.txt
timeout 300 python scripts/run_c4.py 256 21 100 2>&1 | tail -4 | tee gpurun_out/r02_c4_default_256.txt
S21_LIB=$PWD/spice21_b200/libspice21cu_sdiv.so timeout 900 python -m pytest tests/test_gpu.py -m gpu -x -q -k "bsim4" 2>&1 | tail -3
