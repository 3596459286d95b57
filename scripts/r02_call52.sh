#!/usr/bin/env bash
# where the C1-sweep transient's end-to-end time goes (19 ms against a 9.3 ms kernel), with and without the probe-keyed plan cache
set -u
mkdir -p gpurun_out
cat > /tmp/c1_e2e.py <<'PY'
import os, sys, time, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import circuits as cc, spice21_b200 as s21
B = 8192
ro = cc.cmos_ro3(cc.add_mos1_defaults)
b = s21.Batch(ro.to_s21().elaborate(ic={"1": 0.0}), B)
b.override("V:v1:dc", np.linspace(0.9, 1.1, B))
save = np.array([0, 1, 2], dtype=np.int32)
walls = []
for rep in range(6):
    if rep == 5:
        os.environ["S21_PLAN_INFO"] = "1"
    t0 = time.perf_counter()
    b.reset()
    t, w, st, it = b.tran(1e-11, 200e-11, save=save)
    walls.append(time.perf_counter() - t0)
print("PLAN_CACHE", os.environ.get("S21_PLAN_CACHE", "0"), "wall ms per call", [round(1e3 * x, 2) for x in walls], "device_ms", round(b.stats()["device_ms"], 3), "ok", bool(np.all(st == 0)), "checksum", float(w.sum()))
PY
{
for pc in 0 1; do S21_PLAN_CACHE=$pc timeout 300 python /tmp/c1_e2e.py 2>&1 | grep -v "symbolic\]\|s21 plan\]" | cut -c1-300; done
} > gpurun_out/r02Q_c1_e2e.txt 2>&1
cat gpurun_out/r02Q_c1_e2e.txt
