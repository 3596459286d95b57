#!/usr/bin/env bash
# Round-2 GPU call 16: memcheck after the gather-table fix, C4 with arena core in shared memory x mixed workspace, C3 repivot cost.
set -u
mkdir -p gpurun_out
echo "== sanitizer tests"; timeout 1500 python -m pytest tests -m gpu -q -x -k "sanitizer" 2>&1 | tail -5
echo "== C4: arena core copy / mixed workspace"
{
for B in 2048 256; do for cfgs in "0 0" "1 0" "1 1" "0 0" "1 0" "1 1"; do
  set -- $cfgs
  echo "--- B=$B S21_COOP_ARENA=$1 S21_COOP_MIXED=$2"; S21_COOP_ARENA=$1 S21_COOP_MIXED=$2 timeout 600 python scripts/run_c4.py $B 21 100 2>&1 | grep -E "^rep 1|rror"
done; done
} 2>&1 | tee gpurun_out/r02m_c4_arena_mixed.txt
echo "== bsim4 / kernel-variant tests"; timeout 900 python -m pytest tests -m gpu -q -x -k "bsim4 or c4 or golden or variants or grid" 2>&1 | tail -3
echo "== C3 (400 rings): weak-pivot multiplier of the tolerance mode"
{
for wm in 1.000001e3 1e9; do
  echo "--- S21_GRID_WEAK_MULT=$wm"; S21_GRID_WEAK_MULT=$wm S21_PLAN_INFO=1 timeout 900 python scripts/run_c3.py 400 5 2e-10 2>&1 | grep -E "rings=|second run|ring 0" | cut -c1-400
done
} 2>&1 | tee gpurun_out/r02m_c3_weak.txt
echo "== ncu C4 tran kernel (5 points)"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_coop -c 1 -s 3 -f -o gpurun_out/r02m_c4_2048 python scripts/run_c4.py 2048 21 5 > gpurun_out/r02m_c4_ncu.log 2>&1; echo "rc=$?"
