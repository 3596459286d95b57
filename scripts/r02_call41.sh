#!/usr/bin/env bash
set -u
cat > /tmp/probe_c4x.py <<'PY'
import os, sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import circuits as cc, spice21_b200 as s21
B, npts, tstep = 16, 48, 1e-10
sup = np.linspace(0.8, 1.2, 64)[[0, 8, 12, 16, 17, 24, 30, 33, 37, 40, 47, 50, 54, 60, 62, 63]]
ck, ic = cc.bsim4_ring(41, ic_every=20)
b = s21.Batch(ck.to_s21().elaborate(ic=ic), B)
b.override("V:vsup:dc", sup)
x, st0, it0 = b.dcop()
print("dcop status", st0.tolist(), "iters", it0.tolist())
b.reset()
t, w, st, it = b.tran(tstep, 4 * tstep)
print("tran(4 points) status", st.tolist(), "kernel", b.kernel_name(), "first NaN point per instance", [int(np.argmax(~np.isfinite(w[i]).all(axis=1))) if not np.isfinite(w[i]).all() else -1 for i in range(B)])
for k in (2, 6):
    b1 = s21.Batch(ck.to_s21().elaborate(ic=ic), 1)
    b1.override("V:vsup:dc", sup[k:k+1])
    t, w1, st1, it1 = b1.tran(tstep, 4 * tstep)
    print("instance", k, "alone: status", st1.tolist(), "iters", it1.tolist(), "kernel", b1.kernel_name())
PY
S21_PLAN_INFO=1 timeout 600 python /tmp/probe_c4x.py 2>&1 | grep -v "symbolic\]\|s21 plan\]" | cut -c1-400 | tail -60
