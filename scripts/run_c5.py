"""Config C5 on the GPU: AC analysis of the RC ladder + Mos1 op-amp, 100 000 frequency points as one batch of complex
f64 sparse solves (tests/circuits.py::rc_opamp). usage: python scripts/run_c5.py [n_points=100000] [n_sections=64]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import circuits as cc  # noqa: E402
import spice21_b200 as s21  # noqa: E402

npts = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
nsec = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ck = cc.rc_opamp(nsec)
c = ck.to_s21().elaborate()
f = s21.ac_freqs(1, 10**10, npts - 1)
b = s21.Batch(c, 1)
for rep in range(3):
    t0 = time.time()
    x, st, it = b.ac(f)
    wall = time.time() - t0
    s = b.stats()
    print(f"rep {rep}: points={len(f)} N={c.n_vars} nnz_lu={s['nnz_lu']} kernel={b.kernel_name()} ok={int(np.sum(st == 0))} solves={int(it.sum())} "
          f"device_ms={s['device_ms']:.3f} wall_s={wall:.3f} points/s(device)={len(f) / s['device_ms'] * 1e3:.3e} "
          f"complex LU+solve/s(device)={it.sum() / s['device_ms'] * 1e3:.3e}", flush=True)
