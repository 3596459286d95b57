#!/usr/bin/env bash
# C1-circuit transient sweep: lanes per instance of the warp-private team kernel (8 and 16 were never measured with it)
set -u
mkdir -p gpurun_out
timeout 600 python scripts/sweep_tran.py 16,8,4,2 8192 > gpurun_out/r02S_c1_lpi.txt 2>&1
timeout 600 python scripts/sweep_tran.py 8,4 2048 >> gpurun_out/r02S_c1_lpi.txt 2>&1
cat gpurun_out/r02S_c1_lpi.txt | cut -c1-200
