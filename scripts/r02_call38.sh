#!/usr/bin/env bash
# bench.py on 2 GPUs under torchrun (the driver's launch), strong scaling default; the reference arm
set -u
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02F_bench_n2.json 2> gpurun_out/r02F_bench_n2.err; echo "rc=$?"; tail -c 400 gpurun_out/r02F_bench_n2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02F_bench_n2.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['scaling'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], (d.get('weak') or {}).get('value'))
for k, v in (d.get('configs') or {}).items():
    print(k, v.get('value'), v.get('unit'), v.get('ms_per_transient'), v.get('ms_per_timepoint'), 'e2e', (v.get('e2e') or {}).get('ms'), (v.get('rcp_division') or {}).get('ms_per_transient'))
PY
timeout 900 python bench.py > gpurun_out/r02F_bench.json 2> gpurun_out/r02F_bench.err; tail -c 200 gpurun_out/r02F_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02F_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])
for k, v in d['configs'].items():
    print(k, v.get('value'), v.get('unit'), v.get('ms_per_transient'), v.get('ms_per_timepoint'), 'e2e', (v.get('e2e') or {}).get('ms'), (v.get('rcp_division') or {}).get('ms_per_transient'))
PY
