#!/usr/bin/env bash
# bench.py on 8 GPUs under torchrun (the driver's launch): the small-shard paths (C2 2048 per GPU, C5 25 000 points per GPU)
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02AC_bench_n8.json 2> gpurun_out/r02AC_bench_n8.err; echo "rc=$?"; tail -c 300 gpurun_out/r02AC_bench_n8.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02AC_bench_n8.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['scaling'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], (d.get('weak') or {}).get('value'), d.get('limiter', '')[:80])
for k, v in (d.get('configs') or {}).items():
    print(k, v.get('value'), v.get('unit'), v.get('ms_per_transient'), v.get('ms_per_timepoint'), 'e2e', (v.get('e2e') or {}).get('ms'), (v.get('rcp_division') or {}).get('ms_per_transient'))
PY
