#!/usr/bin/env bash
# Round-2 GPU call 12 (2 GPUs): the in-library sweep over both devices, bench.py under torchrun at N = 2 (strong + weak).
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "== sweep / aids / adaptive tests on 2 GPUs"; timeout 600 python -m pytest tests -m gpu -q -x -k "sweep_single_process or aids or adaptive_rc" 2>&1 | tail -5
echo "== bench N=2 (strong)"
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
  bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02i_bench_n2.json 2> gpurun_out/r02i_bench_n2.err ); echo "rc=$?"; tail -4 gpurun_out/r02i_bench_n2.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02i_bench_n2.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','scaling','n_gpus')}, 'kernel_ms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'weak', d.get('weak'))
    for k,v in (d.get('configs') or {}).items(): print(k, v.get('value'), v.get('unit'), 'ms', v.get('ms_per_transient') or v.get('ms_per_sweep') or v.get('ms_per_timepoint'), 'e2e', (v.get('e2e') or {}).get('value'))
    print('skipped', d.get('skipped'), 'limiter', d.get('limiter'))
except Exception as e: print('parse failed', e)
PY
