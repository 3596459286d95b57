#!/usr/bin/env bash
# the AC thread-per-point kernel under ncu: how many bytes it really moves (C5, 100 000 points)
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_ac -c 1 -s 1 -f -o gpurun_out/r02M_c5_kac python scripts/run_c5.py > gpurun_out/r02M_c5_ncu.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/r02M_c5_kac.ncu-rep --page raw --csv > gpurun_out/r02M_c5_kac.raw.csv 2>/dev/null; ls -la gpurun_out/r02M*
timeout 600 python scripts/run_c5.py 2>&1 | grep -E "^rep" | cut -c1-250
