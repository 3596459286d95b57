"""C1 circuit (Mos1 CMOS ring oscillator) as a supply sweep: transient device time per team-kernel shape.
usage: python scripts/sweep_tran.py [lpis=8,4,2] [sizes=8192]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import circuits as cc  # noqa: E402
import spice21_b200 as s21  # noqa: E402

lpis = (sys.argv[1] if len(sys.argv) > 1 else "8,4,2").split(",")
sizes = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "8192").split(",")]
os.environ["S21_KERNEL"] = "jitteam"
for B in sizes:
    ref = None
    for lpi in lpis:
        os.environ["S21_TEAM_LPI"] = lpi
        ro = cc.cmos_ro3(cc.add_mos1_defaults)
        b = s21.Batch(ro.to_s21().elaborate(ic={"1": 0.0}), B)
        b.override("V:v1:dc", np.linspace(0.9, 1.1, B))
        best = 1e9
        for rep in range(2):
            b.reset()
            t, w, st, it = b.tran(1e-11, 2e-9, save=np.array([0, 1, 2], dtype=np.int32))
            best = min(best, b.stats()["device_ms"])
        if ref is None:
            ref = w
        print(f"B={B:7d} lpi={lpi:>2s} kernel={b.kernel_name()} device_ms={best:8.3f} points/s={B * (len(t) - 1) / best * 1e3:.3e} "
              f"iters/s={it.sum() / best * 1e3:.3e} ok={int(np.sum(st == 0))} same_bits={bool(np.array_equal(w, ref))}", flush=True)
