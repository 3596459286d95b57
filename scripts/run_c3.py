"""Config C3 (large single circuit: many 5-stage Mos1 ring oscillators on one supply node) on the GPU.
usage: python scripts/run_c3.py <n_rings> <n_stages> <tstop> [oracle]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import circuits as cc  # noqa: E402
import spice21_b200 as s21  # noqa: E402

nr, ns, tstop = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
ck, ic = cc.inverter_array(nr, ns)
t0 = time.time()
c = ck.to_s21().elaborate(ic=ic)
t1 = time.time()
b = s21.Batch(c, 1)
save = np.array([c.names.index(n) for n in ("vdd", "vddi", "vsup", "r0s0", "r0s1", f"r{nr - 1}s0", f"r{nr - 1}s1")], dtype=np.int32)
t, w, st, it = b.tran(1e-11, tstop, save=save)
t2 = time.time()
stats = b.stats()
print(f"rings={nr} stages={ns} N={c.n_vars} devices={c.n_devices} elaborate={t1 - t0:.2f}s tran(wall incl. symbolic)={t2 - t1:.2f}s "
      f"points={len(t)} status={st[0]} newton_iters={int(it[0])} stats={stats}")
t3 = time.time()
t, w2, st, it = b.tran(1e-11, tstop, save=save)
print(f"second run wall={time.time() - t3:.2f}s device_ms={b.stats()['device_ms']:.2f}")
print("ring 0 vs last ring max |diff|:", float(np.max(np.abs(w[0, :, 3:5] - w[0, :, 5:7]))), " v(r0s1) tail:", w[0, -3:, 4])
if len(sys.argv) > 4:
    from oracle import pyoracle as po
    o = po.Circuit(ck.to_text()).tran(1e-11, tstop, ic=ic)
    ref = o.data[:, save]
    print("oracle solves", o.solves, "max |gpu - oracle| =", float(np.max(np.abs(w[0] - ref))), "oracle solve_s", round(o.seconds, 2))
