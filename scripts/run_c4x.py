"""C4 at SURVEY's size: longer BSIM4 rings (41 stages single IC; 101 stages with ICs every 26 stages), optionally with
rbodymod = rgatemod = 1 (4 internal nodes per device, N = 913). Timing on the GPU and parity of a few instances with the oracle.
usage: python scripts/run_c4x.py <n_stages> <ic_every> <B> <n_points> <extra: 0 | 1 (rbodymod=1, rgatemod=1)> [oracle_instances]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import circuits as cc  # noqa: E402
import spice21_b200 as s21  # noqa: E402

ns, ice, B, npts, extra = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
n_or = int(sys.argv[6]) if len(sys.argv) > 6 else 0
ovr_cards = {"rbodymod": 1, "rgatemod": 1} if extra else {}
tstep = 1e-10
ck, ic = cc.bsim4_ring(ns, ic_every=ice, **ovr_cards)
ovr = cc.c4_sweep(B)
t0 = time.time()
c = ck.to_s21().elaborate(ic=ic)
save = [c.names.index("s1"), c.names.index(f"s{ns // 2}"), c.names.index("vdd")]
b = s21.Batch(c, B)
for k, v in ovr.items():
    b.override(k, v)
for rep in range(2):
    b.reset()
    t1 = time.time()
    t, w, st, it = b.tran(tstep, npts * tstep, save=save)
    wall = time.time() - t1
    stt = b.stats()
    print(f"rep {rep}: stages={ns} ic_every={ice} extra={extra} N={c.n_vars} devices={c.n_devices} B={B} points={len(t) - 1} ok={int(np.sum(st == 0))}/{B} "
          f"iters={int(it.sum())} device_ms={stt['device_ms']:.1f} wall_s={wall:.2f} iters/s={it.sum() / (stt['device_ms'] * 1e-3):.3e} "
          f"nnz_lu={stt['nnz_lu']} kernel={b.kernel_name()} setup={b.setup_stats()}", flush=True)
print("plan", b.plan_info())
if n_or:
    from oracle import pyoracle as po
    sub = {k: v[:n_or] for k, v in ovr.items()}
    t2 = time.time()
    o = po.Circuit(ck.to_text()).batch(1, n_or, overrides=sub, tstep=tstep, tstop=npts * tstep, ic=ic, nthreads=16)
    print(f"oracle {n_or} instances: wall {time.time() - t2:.1f} s, ok {int(np.sum(o['status'] == 0))}, iters {o['iters'].tolist()[:4]}")
    ow = o["x"][:, :, save]
    print("max |gpu - oracle| =", float(np.nanmax(np.abs(w[:n_or] - ow))), " gpu iters", it[:n_or].tolist()[:4])
