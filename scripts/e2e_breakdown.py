"""Where the end-to-end time of one C2 step goes (host wall clock per C-ABI call, median of 200)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import circuits as cc  # noqa: E402
import spice21_b200 as s21  # noqa: E402

B = 8192
ck, ovr = cc.diffpair(), cc.diffpair_mc(B)
b = s21.Batch(ck.to_s21().elaborate(), B)
for k, v in ovr.items():
    b.override(k, v)
b.dcop()
T = {"sync_params": [], "reset": [], "dcop_device": [], "read": [], "total": []}
for rep in range(200):
    t0 = time.perf_counter(); b.sync_params(True)
    t1 = time.perf_counter(); b.reset()
    t2 = time.perf_counter(); b.dcop_device()
    t3 = time.perf_counter(); x, st, it = b.read()
    t4 = time.perf_counter()
    for k, v in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0)):
        T[k].append(v)
for k, v in T.items():
    print(f"{k:12s} median {np.median(v) * 1e6:8.1f} us   min {np.min(v) * 1e6:8.1f} us")
print("device_ms", b.stats()["device_ms"])
