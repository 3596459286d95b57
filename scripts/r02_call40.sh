#!/usr/bin/env bash
# transient re-pivot + resume (cooperative kernel): the tests whose instances used to end with Singular Matrix inside the time loop
set -u
mkdir -p gpurun_out
cat > /tmp/probe_c4x.py <<'PY'
import os, sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import circuits as cc, spice21_b200 as s21
from oracle import pyoracle as po
B, npts, tstep = 16, 48, 1e-10
sup = np.linspace(0.8, 1.2, 64)[[0, 8, 12, 16, 17, 24, 30, 33, 37, 40, 47, 50, 54, 60, 62, 63]]
for name, (ck, ic) in (("41-stage plain", cc.bsim4_ring(41, ic_every=20)), ("41-stage rbodymod=rgatemod=1", cc.bsim4_ring(41, ic_every=20, rbodymod=1, rgatemod=1))):
    for rep in ("1", "0"):
        os.environ["S21_TRAN_REPIVOT"] = rep
        b = s21.Batch(ck.to_s21().elaborate(ic=ic), B)
        b.override("V:vsup:dc", sup)
        t, w, st, it = b.tran(tstep, npts * tstep)
        ss = b.setup_stats()
        print(name, "S21_TRAN_REPIVOT=" + rep, "status", st.tolist(), "nan", np.any(~np.isfinite(w), axis=(1, 2)).astype(int).tolist(), "repaired", ss.get("repaired_instances"), "weak", ss.get("weak_pivot_instances"), "device_ms", round(b.stats()["device_ms"], 2), flush=True)
    o = po.Circuit(ck.to_text()).batch(1, B, overrides={"V:vsup:dc": sup}, tstep=tstep, tstop=npts * tstep, ic=ic, nthreads=8)
    ok = (st == 0) & (o["status"] == 0) & ~np.any(~np.isfinite(w), axis=(1, 2))
    print(name, "oracle status", o["status"].tolist(), "max |gpu - oracle| where both ok:", float(np.max(np.abs(w[ok] - o["x"][ok]))) if ok.any() else None, "iters", it[ok].tolist(), o["iters"][ok].tolist(), flush=True)
PY
timeout 900 python /tmp/probe_c4x.py 2>&1 | cut -c1-700
timeout 900 python -m pytest tests -m gpu -q -s -k "fast_division or c4x or c4_ or bsim4 or sanitizer or golden" 2>&1 | grep -E "^E  |passed|failed|fast division|^tests/test_gpu.py:[0-9]+|FAILED" | cut -c1-500 | head -40
