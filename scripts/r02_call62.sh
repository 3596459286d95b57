#!/usr/bin/env bash
# ncu of the final default C4 cooperative kernel (CTA-blocked staging area) and of the grid-wide kernel at the full C3 size
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_coop -c 1 -s 1 -f -o gpurun_out/r02AA_c4_default python scripts/run_c4.py 2048 21 5 > gpurun_out/r02AA_c4_ncu.log 2>&1; echo "c4 rc=$?"
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:k_grid -c 1 -s 1 -f -o gpurun_out/r02AA_c3_full python scripts/run_c3.py 2000 5 1e-10 > gpurun_out/r02AA_c3_ncu.log 2>&1; echo "c3 rc=$?"
for r in r02AA_c4_default r02AA_c3_full; do ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null; ls -la gpurun_out/$r.ncu-rep; done
