#!/usr/bin/env bash
# ncu evidence for this session's kernels: launch list of the bench command, --set full of the grid-wide kernel (C3 transient, 400 rings)
# and of the reciprocal-division cooperative kernel (C4 operating point launch), C2 kernel re-captured
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02G_launches.csv python bench.py --steps 2 --warmup 3 --extras 0 > gpurun_out/r02G_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_grid -c 1 -s 1 -f -o gpurun_out/r02G_c3_grid python scripts/run_c3.py 400 5 1e-10 > gpurun_out/r02G_c3_ncu.log 2>&1; echo "c3 rc=$?"
S21_B4_FAST=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_coop -c 1 -s 1 -f -o gpurun_out/r02G_c4_rcp python scripts/run_c4.py 2048 21 5 > gpurun_out/r02G_c4_ncu.log 2>&1; echo "c4 rcp rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_jit -c 1 -s 3 -f -o gpurun_out/r02G_c2 python bench.py --steps 2 --warmup 3 --extras 0 > gpurun_out/r02G_c2_ncu.log 2>&1; echo "c2 rc=$?"
for r in r02G_c3_grid r02G_c4_rcp r02G_c2; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
  ls -la gpurun_out/$r.ncu-rep gpurun_out/$r.raw.csv
done
