#!/usr/bin/env bash
# Round-2 GPU call 15: memcheck details, C4 with the mixed workspace (on / off), ncu of the new C4 kernel, e2e breakdown.
set -u
mkdir -p gpurun_out
echo "== memcheck (team part)"
timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 6 python tests/sanitize_target.py team > gpurun_out/r02l_memcheck_team.txt 2>&1; echo "rc=$?"; grep -c "Invalid" gpurun_out/r02l_memcheck_team.txt; head -60 gpurun_out/r02l_memcheck_team.txt | cut -c1-220
for part in bsim4 ac adaptive grid; do
  timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 3 python tests/sanitize_target.py $part > gpurun_out/r02l_memcheck_$part.txt 2>&1; echo "$part rc=$? $(grep -c Invalid gpurun_out/r02l_memcheck_$part.txt) $(tail -1 gpurun_out/r02l_memcheck_$part.txt)"
done
echo "== C4 mixed workspace"
{
for B in 2048 256; do for mx in 0 1 0 1; do
  echo "--- B=$B S21_COOP_MIXED=$mx"; S21_COOP_MIXED=$mx timeout 600 python scripts/run_c4.py $B 21 100 2>&1 | grep -E "^rep 1|rror"
done; done
} 2>&1 | tee gpurun_out/r02l_c4_mixed.txt
echo "== bsim4 tests"; timeout 900 python -m pytest tests -m gpu -q -x -k "bsim4 or c4 or golden" 2>&1 | tail -3
echo "== ncu C4 (mixed, 5 points)"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_coop -c 1 -s 1 -f -o gpurun_out/r02l_c4_2048 python scripts/run_c4.py 2048 21 5 > gpurun_out/r02l_c4_ncu.log 2>&1; echo "rc=$?"
echo "== e2e breakdown"; timeout 300 python scripts/e2e_breakdown.py 2>&1 | tee gpurun_out/r02l_e2e_breakdown.txt
