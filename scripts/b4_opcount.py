"""Executed divisions / exp / log / sqrt per Bsim4 evaluation, counted on the host build of the shared evaluation headers
(oracle/Makefile: liboracle_count.so, -DS21_B4_COUNT), and how many of the divisions fall outside the range in which
the hardware division fast path is exact. Input to the C4 kernel work."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle_count.so"], check=True, capture_output=True)
os.environ["ORC_LIB"] = os.path.join(ROOT, "oracle", "liboracle_count.so")
import circuits as cc  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

L = po.lib()
L.orc_b4_counts.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]


def counts(reset=True):
    out = (C.c_ulonglong * 6)()
    L.orc_b4_counts(out, 1 if reset else 0)
    return list(out)


def report(name):
    ev, dv, sp, ex, lg, sq = counts()
    if ev == 0:
        print(f"{name}: no evaluations")
        return
    print(f"{name}: {ev} evaluations; per evaluation: div {dv / ev:.1f} (outside the exact fast-path range: {sp} in total), "
          f"exp {ex / ev:.1f}, log {lg / ev:.1f}, sqrt {sq / ev:.1f}")


if __name__ == "__main__":
    ck, ic = cc.bsim4_ring(21)
    counts()
    po.Circuit(ck.to_text()).tran(1e-10, 2e-9, ic=ic)
    report("C4 ring (21 stages, default cards), OP + 20 points")
    ck, ic = cc.bsim4_ring(7, cards="ptm65")
    try:
        po.Circuit(ck.to_text()).tran(1e-11, 2e-10, ic=ic)
    except Exception as e:  # the short-channel cards fail to converge at some supplies
        print("ptm65:", e)
    report("7-stage ring, PTM-65 cards")
    ck, ic = cc.bsim4_ring(5, rbodymod=1, rgatemod=1, igcmod=1, igbmod=1)
    try:
        po.Circuit(ck.to_text()).tran(1e-10, 1e-9, ic=ic)
    except Exception as e:
        print("selectors:", e)
    report("5-stage ring, rbodymod = rgatemod = igcmod = igbmod = 1")
