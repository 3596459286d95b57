"""C2 (Mos1 diff-pair Monte-Carlo dcop) at several batch sizes, per kernel variant: device ms and Newton iterations/s.
usage: python scripts/sweep_batch.py [kernels=hybrid,jit] [sizes=8192,65536,262144]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import circuits as cc  # noqa: E402
import spice21_b200 as s21  # noqa: E402

kernels = (sys.argv[1] if len(sys.argv) > 1 else "hybrid,jit").split(",")
sizes = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "8192,65536,262144").split(",")]
for B in sizes:
    ref = None
    for k in kernels:
        kk, _, lpi = k.partition(":")  # "jitteam:16" = team kernel with 16 lanes per instance
        os.environ["S21_KERNEL"] = kk
        os.environ.pop("S21_TEAM_LPI", None)
        if lpi:
            os.environ["S21_TEAM_LPI"] = lpi
        ck, ovr = cc.diffpair(), cc.diffpair_mc(B)
        b = s21.Batch(ck.to_s21().elaborate(), B)
        for key, v in ovr.items():
            b.override(key, v)
        best = 1e9
        for rep in range(5):
            b.reset()
            x, st, it = b.dcop()
            best = min(best, b.stats()["device_ms"])
        if ref is None:
            ref = x
        print(f"B={B:7d} kernel={k:10s} device_ms={best:8.3f} iters={int(it.sum()):9d} iters/s={it.sum() / best * 1e3:.3e} "
              f"ok={int(np.sum(st == 0))} same_bits={bool(np.array_equal(x, ref))}", flush=True)
