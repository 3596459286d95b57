#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
cat > /tmp/c4_41.py <<'PY'
import os, sys, time, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import circuits as cc, spice21_b200 as s21
ck, ic = cc.bsim4_ring(41, ic_every=20)
for B in (2048,):
    ovr = cc.c4_sweep(B)
    b = s21.Batch(ck.to_s21().elaborate(ic=ic), B)
    for k, v in ovr.items():
        b.override(k, v)
    for rep in range(2):
        b.reset()
        t0 = time.time()
        t, w, st, it = b.tran(1e-10, 100e-10, save=[1, 2, 3])
        wall = time.time() - t0
    vals, cnt = np.unique(st, return_counts=True)
    bad = np.unique(ovr["V:vsup:dc"][st != 0])
    print(f"B={B} status counts {dict(zip(vals.tolist(), cnt.tolist()))} failing supplies {bad.tolist()} iters {int(it.sum())} device_ms {b.stats()['device_ms']:.2f} wall {wall:.3f} launches {b.stats()['launches']} setup {b.setup_stats()}")
PY
{
echo "--- default"; timeout 600 python /tmp/c4_41.py 2>&1 | tail -2 | cut -c1-600
echo "--- S21_B4_FAST=1"; S21_B4_FAST=1 timeout 600 python /tmp/c4_41.py 2>&1 | tail -2 | cut -c1-600
} > gpurun_out/r02J_c4_41.txt 2>&1
cat gpurun_out/r02J_c4_41.txt
