"""Per-phase cycle breakdown of the hybrid kernel on the bench workload. Needs the instrumented library built with
`make -C spice21_b200/csrc PROFILE=1 OUT=<path>`; pass its path in S21_PROF_LIB."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import spice21_b200 as s21  # noqa: E402
s21.LIB_PATH = os.environ["S21_PROF_LIB"]
import circuits as cc  # noqa: E402

B = 8192
ck, ovr = cc.diffpair(), cc.diffpair_mc(B)
b = s21.Batch(ck.to_s21().elaborate(), B)
for k, v in ovr.items():
    b.override(k, v)
b.dcop()
out = (C.c_ulonglong * 16)()
s21.lib().s21_debug_phase_cycles(out, 1)
b.reset()
x, st, it = b.dcop()
s21.lib().s21_debug_phase_cycles(out, 0)
names = ["eval", "barrier after eval", "assemble", "residual+decide", "LU", "fwd+bwd", "update", "barrier end"]
n_cta = B // 32
for w in range(2):
    tot = sum(out[8 * w + k] for k in range(8))
    print(f"warp {w} (evaluates {'Mos1' if w < 2 else 'passive'}): total {tot / n_cta:.0f} cycles per CTA, {tot / n_cta / 20:.0f} per iteration")
    for k in range(8):
        print(f"   {names[k]:20s} {out[8 * w + k] / n_cta:10.0f} cycles/CTA  {100.0 * out[8 * w + k] / tot:5.1f}%")
print("iters", int(it.sum()), "stats", b.stats())
