#!/usr/bin/env bash
# GPU capture plan for what round 1 left unmeasured (run pieces through gpurun; every step has its own `timeout`, so a
# hang costs seconds, not the budget — the 8-GPU hang of round 1 cost 84 GPU-minutes).
#   gpurun --timeout 900 -- 'bash scripts/r02_gpu_plan.sh accept'        # 1 GPU
#   gpurun --timeout 900 -- 'bash scripts/r02_gpu_plan.sh wp'            # 1 GPU
#   gpurun --timeout 1500 -- 'bash scripts/r02_gpu_plan.sh sdiv'         # 1 GPU (builds a second library, ~2 min)
#   gpurun --timeout 1200 -- 'bash scripts/r02_gpu_plan.sh cache'        # 1 GPU
#   gpurun --gpus 2 --timeout 300 -- 'bash scripts/r02_gpu_plan.sh scale 2'   # then 4, then 8: ~1 min of box time each
set -u
mkdir -p gpurun_out
case "${1:-}" in
  accept)  # acceptance test of the warp-private team kernel + the full GPU suite
    S21_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu.py -m gpu -x -q -k "warp_private" 2>&1 | tail -5
    timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ;;
  wp)      # warp-private loop vs the CTA-wide evaluation phase, CTA sizes, lanes per instance
    for wp in 0 1; do for gi in 32 16; do
      echo "WP=$wp GI=$gi"; S21_TEAM_WP=$wp S21_TEAM_GI=$gi timeout 120 python scripts/sweep_batch.py jitteam:2,jitteam:4 4736,8192,16384,65536
    done; done 2>&1 | tee gpurun_out/r02_wp_sweep.txt
    for wp in 0 1; do echo "WP=$wp"; S21_TEAM_WP=$wp timeout 300 python scripts/sweep_tran.py 4,2 8192; done 2>&1 | tee gpurun_out/r02_wp_tran.txt
    S21_TEAM_WP=1 timeout 300 python bench.py > gpurun_out/r02_bench_wp.json 2> gpurun_out/r02_bench_wp.err; cut -c1-400 gpurun_out/r02_bench_wp.json ;;
  sdiv)    # BSIM4 divisions through scalar.h: same bits (variants test) and C4 timing, against the default build
    timeout 900 make -C spice21_b200/csrc -j8 B4_SDIV=1 OUT=/tmp/libspice21cu_sdiv.so > gpurun_out/r02_sdiv_build.log 2>&1 || { tail -5 gpurun_out/r02_sdiv_build.log; exit 1; }
    timeout 600 python scripts/run_c4.py 2048 21 100 2>&1 | tail -6 | tee gpurun_out/r02_c4_default.txt
    S21_LIB=/tmp/libspice21cu_sdiv.so timeout 600 python scripts/run_c4.py 2048 21 100 2>&1 | tail -6 | tee gpurun_out/r02_c4_sdiv.txt
    S21_LIB=/tmp/libspice21cu_sdiv.so timeout 900 python -m pytest tests/test_gpu.py -m gpu -x -q -k "bsim4" 2>&1 | tail -3 ;;
  cache)   # transient plan kept when the probe values repeat (S21_PLAN_CACHE=1): same results, second run without the symbolic phase
    S21_PLAN_CACHE=1 timeout 900 python -m pytest tests/test_gpu.py -m gpu -x -q -k "tran or golden or variants or grid" 2>&1 | tail -3
    for c in 0 1; do echo "PLAN_CACHE=$c"; S21_PLAN_CACHE=$c S21_PLAN_INFO=1 timeout 600 python scripts/run_c3.py 400 5 2e-10 2>&1 | grep -E "second run|symbolic phase|rings=" ; done | tee gpurun_out/r02_plan_cache.txt ;;
  scale)   # multi-GPU bench, one N per call; 120 s NCCL timeout inside bench.py, 240 s here
    n="${2:?number of GPUs}"
    timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus "$n" --steps 20 --warmup 5 > "gpurun_out/r02_bench_n$n.json" 2> "gpurun_out/r02_bench_n$n.err"
    echo "rc=$?"; cut -c1-300 "gpurun_out/r02_bench_n$n.json"; tail -3 "gpurun_out/r02_bench_n$n.err" ;;
  *) echo "usage: $0 accept | wp | sdiv | cache | scale N"; exit 2 ;;
esac
