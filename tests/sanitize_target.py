"""Workload that tests/test_gpu.py::test_compute_sanitizer_clean runs UNDER compute-sanitizer (memcheck / racecheck):
one small pass through every kernel family — the run-time specialised team kernel (dcop + transient, ragged batch so
that padding lanes exist), the thread-per-instance specialised kernel, hybrid, cooperative (shared-memory and HBM
workspace, the latter with Bsim4 — default and reciprocal-division builds, plus a hand-back / resume inside the time loop),
direct (dcop, tran, adaptive tran, AC), grid-wide, the probe and pack kernels (results, waveforms, AC rows).
Sizes are tiny: the sanitizer slows kernels by one to two orders of magnitude. Prints SANITIZE_TARGET_OK at the end."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import circuits as cc  # noqa: E402
import spice21_b200 as s21  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"


def dcop_mc(B, kernel):
    if kernel:
        os.environ["S21_KERNEL"] = kernel
    else:
        os.environ.pop("S21_KERNEL", None)
    dp = cc.diffpair()
    b = s21.Batch(dp.to_s21().elaborate(), B)
    for k, v in cc.diffpair_mc(B).items():
        b.override(k, v)
    x, st, it = b.dcop()
    assert np.all(st == 0), (kernel, st)
    return x


def tran_ro(B, kernel, points=12):
    if kernel:
        os.environ["S21_KERNEL"] = kernel
    else:
        os.environ.pop("S21_KERNEL", None)
    ck = cc.cmos_ro3(cc.add_mos1_defaults)
    b = s21.Batch(ck.to_s21().elaborate(ic={"1": 0.0}), B)
    if B > 1:
        b.override("V:v1:dc", np.linspace(0.9, 1.1, B))
    t, w, st, it = b.tran(1e-11, points * 1e-11)
    assert np.all(st == 0), (kernel, st)
    return w


if which in ("all", "team"):
    ref = dcop_mc(37, "direct")                      # ragged: 37 instances, padding lanes in the last warp
    for kern in (None, "jitteam", "jit", "hybrid", "coop"):
        x = dcop_mc(37, kern)
        assert np.array_equal(x, ref), kern
    wref = tran_ro(5, "direct")
    for kern in (None, "hybrid", "coop"):
        assert np.array_equal(tran_ro(5, kern), wref), kern
    os.environ["S21_TEAM_SOPG"] = "1"  # team kernel with the committed device state left in its HBM column (host/jit_team.hpp)
    assert np.array_equal(tran_ro(5, "jitteam"), wref)
    os.environ.pop("S21_TEAM_SOPG")
    os.environ.pop("S21_KERNEL", None)

if which in ("all", "bsim4"):
    rb = cc.cmos_ro3(cc.add_bsim4_defaults)
    b = s21.Batch(rb.to_s21().elaborate(ic={"1": 0.0}), 3)
    b.override("V:v1:dc", np.array([0.9, 1.0, 1.1]))
    t, w, st, it = b.tran(1e-10, 4e-10)
    assert np.all(st == 0), st
    # the opt-in reciprocal-division build of the same kernel, with every instance handed back once inside the time loop
    # and continued by a resume launch (SolveCtl::tran_stop; S21_TRAN_INJECT is the test hook)
    os.environ["S21_B4_FAST"] = "1"
    os.environ["S21_TRAN_INJECT"] = "2"
    b = s21.Batch(rb.to_s21().elaborate(ic={"1": 0.0}), 3)
    b.override("V:v1:dc", np.array([0.9, 1.0, 1.1]))
    t, w2, st, it = b.tran(1e-10, 4e-10)
    os.environ.pop("S21_B4_FAST"); os.environ.pop("S21_TRAN_INJECT")
    assert np.all(st == 0) and b.kernel_name() == "coop-rcp" and np.max(np.abs(w2 - w)) < 1e-6, st

if which in ("all", "ac"):
    c = cc.Ckt().V("vin", "inp", cc.GND, 1.0, acm=1.0).R("r1", "inp", "out", 1e-3).C("c1", "out", cc.GND, 1e-9)
    xa, st, it = s21.Batch(c.to_s21().elaborate(), 1).ac(np.logspace(1, 8, 40))
    assert np.all(st == 0)
    ro = cc.rc_opamp(4)
    xa, st, it = s21.Batch(ro.to_s21().elaborate(), 1).ac(np.logspace(1, 8, 9))
    assert np.all(st == 0)

if which in ("all", "adaptive"):
    c = cc.Ckt().V("vin", "inp", cc.GND, 1.0).R("r1", "inp", "out", 1e-3).C("c1", "out", cc.GND, 1e-9)
    b = s21.Batch(c.to_s21().elaborate(ic={"out": 0.0}), 2)
    r = b.tran_adaptive(1e-7, 5e-6)
    assert np.all(r[2] == 0)

if which in ("all", "grid"):
    ck, ic = cc.inverter_array(20, 5)
    b = s21.Batch(ck.to_s21().elaborate(ic=ic), 1)
    t, w, st, it = b.tran(1e-11, 3e-11)
    assert st[0] == 0 and b.kernel_name() == "grid", (st, b.kernel_name())

print("SANITIZE_TARGET_OK")
