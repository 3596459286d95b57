"""One C2 step end to end, as bench.py's e2e loop runs it (sync_params(force) + reset + dcop_view), with the library's
S21_TRACE_E2E device/host time stamps printed for the last repetitions. usage: S21_TRACE_E2E=1 python scripts/e2e_trace.py [B]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import circuits as cc  # noqa: E402
import spice21_b200 as s21  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ck, ovr = cc.diffpair(), cc.diffpair_mc(B)
b = s21.Batch(ck.to_s21().elaborate(), B)
for k, v in ovr.items():
    b.override(k, v)
b.dcop()
ts = []
for rep in range(60):
    t0 = time.perf_counter()
    x, st, it, nb = b.step_dcop_view(upload=True, reset=True)
    ts.append(time.perf_counter() - t0)
print(f"python wall per step: median {np.median(ts) * 1e6:.1f} us, min {np.min(ts) * 1e6:.1f} us; kernel {b.kernel_name()} device_ms {b.stats()['device_ms']:.4f}")
