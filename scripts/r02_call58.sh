#!/usr/bin/env bash
# ncu of the C1-circuit transient sweep kernel (team kernel, 8192 instances x 200 points): issue slots and FP64 pipe
set -u
mkdir -p gpurun_out
cat > /tmp/c1_one.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import circuits as cc, spice21_b200 as s21
B = 8192
ro = cc.cmos_ro3(cc.add_mos1_defaults)
b = s21.Batch(ro.to_s21().elaborate(ic={"1": 0.0}), B)
b.override("V:v1:dc", np.linspace(0.9, 1.1, B))
for rep in range(2):
    b.reset()
    t, w, st, it = b.tran(1e-11, 200e-11, save=np.array([0, 1, 2], dtype=np.int32))
print("iters", int(it.sum()), "device_ms", b.stats()["device_ms"], "kernel", b.kernel_name())
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_jit -c 1 -s 3 -f -o gpurun_out/r02W_c1_tran python /tmp/c1_one.py > gpurun_out/r02W_c1_ncu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r02W_c1_ncu.log
ncu -i gpurun_out/r02W_c1_tran.ncu-rep --page raw --csv > gpurun_out/r02W_c1_tran.raw.csv 2>/dev/null; ls -la gpurun_out/r02W*
timeout 900 python bench.py --extras 1 > gpurun_out/r02W_bench.json 2> gpurun_out/r02W_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02W_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])
print(json.dumps(d['configs']['c1']['roofline'])[:700])
PY
