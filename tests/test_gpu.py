"""Parity tests proper: the CUDA path (through the C ABI) against the oracle on the same inputs, against the
reference's golden waveforms, and — at BASELINE.json's full batch size — through size-independent properties.

Tolerances (BASELINE.json north_star): dcop node voltages / branch currents within 1e-9 relative; tran and ac within
SPICE reltol = 1e-3 / vntol = 1e-6 — in practice both agree to ~1e-12, and the tests assert the tighter figure so a
regression in operation order is caught. Integer structures (stamp map, pivot order, fill) are exact.
"""
import os

import numpy as np
import pytest

import circuits as cc
from circuits import GND, Ckt
from conftest import golden

pytestmark = pytest.mark.gpu


def rel_err(a, b, floor=1e-6):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))


def close(a, b, rtol=1e-9, atol=1e-11):
    """north_star dcop tolerance (1e-9 relative) with an absolute floor for branch currents that are numerically zero."""
    worst = float(np.max(np.abs(a - b) / (rtol * np.abs(b) + atol)))
    assert worst <= 1.0, f"worst error is {worst:.3g} x the tolerance"
    return True


# ------------------------------------------------------------------------------------------------ transient
RO = [
    (cc.cmos_ro3, cc.add_mos1_defaults, "test_mos1_cmos_ro_tran", 1e-11, 1e-8, True),
    (cc.cmos_ro3, cc.add_mos0_defaults, "test_mos0_cmos_ro_tran", 1e-15, 1e-12, False),
    (cc.nmos_ro3, cc.add_mos1_defaults, "test_mos1_nmos_ro_tran", 1e-11, 1e-8, True),
    (cc.pmos_ro3, cc.add_mos1_defaults, "test_mos1_pmos_ro_tran", 1e-11, 1e-8, True),
    (cc.cmos_ro3, cc.add_bsim4_defaults, "test_bsim4_cmos_ro_tran", 1e-10, 3e-7, True),   # tests.rs:948-964
    (cc.nmos_ro3, cc.add_bsim4_defaults, "test_bsim4_nmos_ro_tran", 1e-9, 1e-6, True),    # tests.rs:1360-1376
    (cc.pmos_ro3, cc.add_bsim4_defaults, "test_bsim4_pmos_ro_tran", 1e-11, 1e-8, True),   # tests.rs:1053-1069
]


@pytest.mark.parametrize("builder,defaults,fixture,tstep,tstop,protoable", RO, ids=[r[2] for r in RO])
def test_golden_waveforms(s21, oracle, builder, defaults, fixture, tstep, tstop, protoable):
    """BASELINE config 1 and its siblings: the reference's golden transients (abs tol 1e-6, tests.rs:788)."""
    g = golden(fixture)
    ck = builder(defaults)
    c = ck.to_s21(via_proto=protoable).elaborate(ic={"1": 0.0})
    t, wave, status, iters = s21.Batch(c, 1).tran(tstep, tstop)
    assert status[0] == 0
    assert np.array_equal(t, g["time"])
    for k, name in enumerate(c.names):
        assert np.max(np.abs(wave[0, :, k] - g[name])) <= 1e-6, name
    o = oracle.Circuit(ck.to_text()).tran(tstep, tstop, ic={"1": 0.0})
    assert np.max(np.abs(wave[0] - o.data)) <= 1e-9
    # iteration-count parity with the reference algorithm (hard thresholds; allow a handful of flips)
    assert abs(int(iters[0]) - o.solves) <= max(3, o.solves // 500), (int(iters[0]), o.solves)


def test_tran_bytes_api_matches_golden(s21):
    """Tran::call_bytes drop-in (proto.rs:79-97): protobuf in, protobuf out, incl. the "time" key."""
    from spice21_b200 import protos as P
    g = golden("test_mos1_cmos_ro_tran")
    res = s21.tran(cc.cmos_ro3(cc.add_mos1_defaults).to_proto(), args=P.TranOptions(tstep=1e-11, tstop=1e-8, ic={"1": 0.0}))
    assert set(res.keys()) == set(g.keys())
    for k in g:
        assert np.max(np.abs(np.array(res[k]) - g[k])) <= 1e-6, k


def test_tran_rc_step(s21, oracle):  # tests.rs:695-717
    ck = Ckt().V("v1", "inp", GND, 1.0).R("r1", "inp", "out", 1e-3).C("c1", "out", GND, 1e-9)
    c = ck.to_s21().elaborate(ic={"out": 0.0})
    t, w, st, _ = s21.Batch(c, 1).tran(10e-9, 10e-6)
    inp, out = w[0, :, c.names.index("inp")], w[0, :, c.names.index("out")]
    assert st[0] == 0 and np.all(inp == 1.0)
    assert abs(out[0]) < 1e-3 and abs(out[-1] - 1.0) < 1e-3 and np.all(np.diff(out) > 0)
    o = oracle.Circuit(ck.to_text()).tran(10e-9, 10e-6, ic={"out": 0.0})
    assert np.max(np.abs(w[0] - o.data)) <= 1e-12


def test_tran_batched_sweep(s21, oracle):
    """A VDD sweep of the ring oscillator: every instance must match its own oracle run."""
    B = 8
    ck = cc.cmos_ro3(cc.add_mos1_defaults)
    vdd = np.linspace(0.8, 1.2, B)
    b = s21.Batch(ck.to_s21().elaborate(ic={"1": 0.0}), B)
    b.override("V:v1:dc", vdd)
    t, w, st, it = b.tran(1e-11, 2e-9)
    o = oracle.Circuit(ck.to_text()).batch(1, B, overrides={"V:v1:dc": vdd}, tstep=1e-11, tstop=2e-9, ic={"1": 0.0})
    assert np.all(st == 0) and np.all(o["status"] == 0)
    assert np.max(np.abs(w - o["x"])) <= 1e-9


# ------------------------------------------------------------------------------------------------ dcop
def _both_dcop(s21, oracle, ck, via_proto=False, **kw):
    c = ck.to_s21(via_proto=via_proto).elaborate(**kw)
    x, st, it = s21.Batch(c, 1).dcop()
    o = oracle.Circuit(ck.to_text()).dcop(**{k: v for k, v in kw.items() if k == "opts"})
    assert c.names == o.names
    return dict(zip(c.names, x[0])), x[0], st[0], int(it[0]), o


def test_dcop_known_answers(s21, oracle):
    # tests.rs:30-44 (exact), 66-83 (exact), 670-692 (exact vectors)
    v, x, st, it, o = _both_dcop(s21, oracle, Ckt(signals=["vdd"]).I("i1", "vdd", GND, 1e-3).R("r1", "vdd", GND, 1e-3), via_proto=True)
    assert st == 0 and v["vdd"] == 1.0
    ck = Ckt(signals=["vdd", "div"]).V("v1", "vdd", GND, 1.0).R("r1", "vdd", "div", 2e-3).R("r2", GND, "div", 2e-3)
    v, x, st, it, o = _both_dcop(s21, oracle, ck, via_proto=True)
    assert v["vdd"] == 1.0 and v["div"] == 0.5 and v["v1"] == -1e-3
    v, x, st, it, o = _both_dcop(s21, oracle, Ckt().R("r1", "1", "0", 1e-3).C("c1", "1", GND, 1e-9).V("v1", "0", GND, 1.0))
    assert x.tolist() == [1.0, 1.0, 0.0]
    v, x, st, it, o = _both_dcop(s21, oracle, Ckt().C("c1", "i", "o", 1e-9).R("r1", "o", GND, 1e-3).V("v1", "i", GND, 1.0))
    assert x.tolist() == [1.0, 0.0, 0.0]
    v, x, st, it, o = _both_dcop(s21, oracle, Ckt().R("r1", "a", GND, 1e-3))
    assert x.tolist() == [0.0] and it == 0


DCOP = [
    ("mos0_nchar", lambda: cc.add_mos0_defaults(Ckt(signals=["g", "d"])).M("m", "nmos", "default", d="d", g="g", s=GND, b=GND)
        .V("v1", "g", GND, 1.0).V("v2", "d", GND, 1.0)),
    ("mos0_diode_swapped", lambda: cc.add_mos0_defaults(Ckt()).I("i1", "0", GND, -5e-3).M("m", "pmos", "", d=GND, g="0", s="0", b=GND)),
    ("mos1_op", lambda: cc.add_mos1_defaults(Ckt()).M("m", "default", "default", d="0", g="0", s=GND, b=GND).V("v1", "0", GND, 1.0)),
    ("mos1_inv", lambda: cc.cmos_inv(cc.add_mos1_defaults)),
    ("mos1_ro3", lambda: cc.cmos_ro3(cc.add_mos1_defaults)),
    ("diffpair", cc.diffpair),
    ("diode_v", lambda: cc.add_diode_defaults(Ckt(signals=["p"])).D("dd", "p", GND, "default", "default").V("vin", "p", GND, 0.7)),
    ("diode_i", lambda: cc.add_diode_defaults(Ckt(signals=["p"])).D("dd", "p", GND, "default", "default").I("i1", "p", GND, 5.7e-3)),
    ("diode_rs_bv", lambda: Ckt(signals=["p"]).define("diodemodel", "default", rs=10.0, bv=5.0, cj0=1e-12).define("diodeinst", "default", area=2.0)
        .D("dd", "p", GND, "default", "default").V("vin", "p", GND, -6.0)),
    ("mos1_rd_rs", lambda: Ckt().define("mos1model", "n", 0, rd=10.0, rs=5.0, vt0=0.4, kp=1e-4).define("mos1inst", "i")
        .M("m", "n", "i", d="d", g="g", s=GND, b=GND).V("vg", "g", GND, 1.0).V("vd", "d", GND, 1.0)),
]


@pytest.mark.parametrize("name,build", DCOP, ids=[d[0] for d in DCOP])
def test_dcop_matches_oracle(s21, oracle, name, build):
    v, x, st, it, o = _both_dcop(s21, oracle, build())
    assert st == 0
    assert rel_err(x, o.data[0], floor=1e-9) <= 1e-9
    assert it == o.solves


def test_dcop_bytes_api(s21):
    """Op::call_bytes drop-in: the reference's Python-binding test (spice21py/tests/test_spice21.py:21-29)."""
    c = s21.circuit([s21.Resistor(p="1", g=1e-3), s21.Capacitor(p="1"), s21.Isrc(p="1", dc=1e-3)])
    res = s21.dcop(s21.protos.Op(ckt=c))
    assert isinstance(res, dict) and res["1"] == 1.0
    res = s21.tran(c)  # test_spice21.py:42-50: default TranOptions -> the t=0 point only
    assert res["1"] == [1.0]


def test_dcop_options(s21, oracle):
    ck = cc.add_diode_defaults(Ckt(signals=["p"])).D("dd", "p", GND, "default", "default").V("vin", "p", GND, 0.65)
    opts = {"temp": 350.0, "gmin": 1e-10}
    v, x, st, it, o = _both_dcop(s21, oracle, ck, opts=opts)
    assert st == 0 and rel_err(x, o.data[0], floor=1e-9) <= 1e-9


def test_pivot_order_from_device_values(s21, oracle):
    """The symbolic phase runs on values probed on the GPU: the resulting order/fill equals the reference's."""
    for ck, ic in ((cc.cmos_ro3(cc.add_mos1_defaults), {"1": 0.0}), (cc.diffpair(), None)):
        o = oracle.Circuit(ck.to_text()).structure(ic=ic)
        b = s21.Batch(ck.to_s21().elaborate(ic=ic), 1)
        b.dcop()
        p = b.pivot_order()
        assert np.array_equal(p["row_i2e"], o["row_i2e"]) and np.array_equal(p["col_i2e"], o["col_i2e"])
        assert set(zip(p["lu_row"].tolist(), p["lu_col"].tolist(), p["lu_fill"].tolist())) == \
            set(zip(o["lu_row"].tolist(), o["lu_col"].tolist(), o["lu_fill"].tolist()))


def test_monte_carlo_dcop_matches_oracle(s21, oracle):
    """BASELINE config 2 at a size the oracle finishes in well under a second."""
    B = 512
    ck, ovr = cc.diffpair(), cc.diffpair_mc(512)
    o = oracle.Circuit(ck.to_text()).batch(0, B, overrides=ovr, nthreads=4)
    b = s21.Batch(ck.to_s21().elaborate(), B)
    for k, v in ovr.items():
        b.override(k, v)
    x, st, it = b.dcop()
    assert np.all(st == 0) and np.all(o["status"] == 0)
    assert rel_err(x, o["x"], floor=1e-9) <= 1e-9
    assert np.mean(it == o["iters"]) >= 0.99
    # warm restart: a second dcop from the converged point needs no linear solve (analysis.rs:188-194)
    x2, st2, it2 = b.dcop()
    assert np.all(st2 == 0) and np.array_equal(it2, it) and np.array_equal(x2, x)
    b.reset()
    x3, st3, it3 = b.dcop()
    assert np.array_equal(x3, x) and np.array_equal(it3, it)  # deterministic


def test_monte_carlo_dcop_full_size_properties(s21):
    """BASELINE config 2 at full size (8192 instances): size-independent properties of the solution."""
    B = 8192
    ck, ovr = cc.diffpair(), cc.diffpair_mc(8192)
    c = ck.to_s21().elaborate()
    b = s21.Batch(c, B)
    for k, v in ovr.items():
        b.override(k, v)
    x, st, it = b.dcop()
    assert np.all(st == 0)
    n = {name: k for k, name in enumerate(c.names)}
    # forced nodes are exact; KCL at the supply: i(vdd) = -(g1 (vdd - on) + g2 (vdd - op)); tail current splits
    assert np.all(x[:, n["vdd"]] == 1.8) and np.all(x[:, n["inp"]] == 0.9)
    i_sup = ovr["R:r1:g"] * (1.8 - x[:, n["on"]]) + ovr["R:r2:g"] * (1.8 - x[:, n["op"]])
    assert np.max(np.abs(i_sup + x[:, n["vsup"]])) < 1e-11  # branch current of the supply source
    assert np.max(np.abs(i_sup - 20e-6)) < 1e-9  # all of the tail current comes from the supply
    assert np.all(it >= 2) and np.all(it <= 100)
    # permutation equivariance: solving a shuffled batch gives the shuffled solution, bit for bit
    perm = np.random.default_rng(0).permutation(B)
    b2 = s21.Batch(c, B)
    for k, v in ovr.items():
        b2.override(k, v[perm])
    x2, st2, it2 = b2.dcop()
    assert np.array_equal(x2, x[perm]) and np.array_equal(it2, it[perm])


def test_monte_carlo_dcop_full_size_matches_oracle(s21, oracle):
    """BASELINE config 2 at full size, compared outright: x within 1e-9 relative and the Newton iteration count of every one
    of the 8192 instances against the CPU restatement (which re-pivots every iteration; 15 ms of host time)."""
    B = 8192
    ck, ovr = cc.diffpair(), cc.diffpair_mc(B)
    o = oracle.Circuit(ck.to_text()).batch(0, B, overrides=ovr, nthreads=8)
    b = s21.Batch(ck.to_s21().elaborate(), B)
    for k, v in ovr.items():
        b.override(k, v)
    x, st, it = b.dcop()
    assert np.all(st == 0) and np.all(o["status"] == 0)
    assert rel_err(x, o["x"], floor=1e-9) <= 1e-9
    assert np.array_equal(it, o["iters"])


def test_sweep_single_process_matches_batch(s21):
    """s21_sweep_* (one process, one host thread + stream per GPU) over every visible device: the gathered x / status / iters
    of a dcop, a transient and an AC sweep equal one Batch's, bit for bit; shard bounds follow s21_sweep_partition."""
    G = s21.cuda_device_count()
    B = 1000  # ragged: not a multiple of the device count times 32
    dp, ovr = cc.diffpair(), cc.diffpair_mc(B)
    c = dp.to_s21().elaborate()
    ref = s21.Batch(c, B)
    sw = s21.Sweep(c, B, n_devices=G)
    assert sw.n_devices == G
    for g, (dev, first, count) in enumerate(sw.shards()):
        assert (first, count) == s21.sweep_partition(B, G, g) and dev == g
    for k, v in ovr.items():
        ref.override(k, v)
        sw.override(k, v)
    x, st, it = ref.dcop()
    xs, sts, its = sw.dcop()
    assert np.array_equal(x, xs) and np.array_equal(st, sts) and np.array_equal(it, its)
    sw.reset()
    assert sw.sync_params(force_upload=True) > 0
    xv, stv, itv = sw.dcop_view()
    assert np.array_equal(x, xv) and np.array_equal(st, stv) and np.array_equal(it, itv)
    assert sw.stats()["iters"] == int(it.sum())
    # transient: the C1 ring as a supply sweep
    ro = cc.cmos_ro3(cc.add_mos1_defaults).to_s21().elaborate(ic={"1": 0.0})
    Bt = 70
    vs = np.linspace(0.9, 1.1, Bt)
    rb, rs = s21.Batch(ro, Bt), s21.Sweep(ro, Bt, n_devices=G)
    rb.override("V:v1:dc", vs)
    rs.override("V:v1:dc", vs)
    t, w, st, it = rb.tran(1e-11, 3e-10)
    t2, w2, st2, it2 = rs.tran(1e-11, 3e-10)
    assert np.array_equal(t, t2) and np.array_equal(w, w2) and np.array_equal(st, st2) and np.array_equal(it, it2)
    # AC: the frequency axis is what shards
    oc = cc.rc_opamp(8).to_s21().elaborate()
    f = s21.ac_freqs(1, 10**9, 500)
    xa, sa, ia = s21.Batch(oc, 1).ac(f)
    xb, sb, ib = s21.Sweep(oc, 1, n_devices=G).ac(f)
    assert np.all(sa == 0) and np.array_equal(sa, sb) and np.array_equal(ia, ib)
    assert np.max(np.abs(xa - xb)) <= 1e-12 * max(1.0, float(np.max(np.abs(xa))))  # each shard pivots on its own first point


def test_packed_device_results(s21):
    """s21_batch_packed_device: the device-resident result block a multi-rank job hands to NCCL equals what dcop returns."""
    torch = pytest.importorskip("torch")
    B = 96
    b = s21.Batch(cc.diffpair().to_s21().elaborate(), B)
    for k, v in cc.diffpair_mc(B).items():
        b.override(k, v)
    x, st, it = b.dcop()
    ptr, words = b.packed_device()

    class Dev:
        __cuda_array_interface__ = {"shape": (words,), "typestr": "<f8", "data": (ptr, False), "version": 2}

    torch.cuda.synchronize()
    host = torch.as_tensor(Dev(), device="cuda").cpu().numpy()
    assert np.array_equal(host[: B * b.N].reshape(B, b.N), x)
    tail = host[B * b.N:].view(np.int32)
    assert np.array_equal(tail[:B], st) and np.array_equal(tail[B:2 * B], it)


def _outlier_circuit():
    """A stage whose gate node `x` hangs on the bias source through a resistor gx: the diagonal (x, x) is gx, the column
    below it carries gm. With gx = 1 S the diagonal is a fine pivot; with gx = 1e-15 S it is 11 orders below gm, and the
    reference's threshold test (|d| >= 1e-3 column max, mod.rs:735-783) refuses it."""
    ck = Ckt().define("mos1model", "m", 0, **cc.C2_MODEL).define("mos1inst", "wl", **cc.C2_INST)
    ck.V("vd", "vdd", GND, 1.8).V("vg", "g0", GND, 0.9).R("rg", "g0", "x", 1.0).R("rl", "vdd", "d", 5e-5).R("rs", "s", GND, 2e-4)
    ck.M("m1", "m", "wl", d="d", g="x", s="s", b=GND)
    return ck


def test_outlier_instance_zero_pivot_order_is_repaired(s21, oracle, monkeypatch):
    """The frozen pivot order comes from instance 0. Here instance 0 is the outlier (gx = 1 S) and every other instance has
    gx = 1e-15 S, for which that order divides by a pivot far below the reference's 1e-3 threshold. The kernels stop those
    instances before the update (internal status 9), the host continues them with a symbolic phase of their own at the
    iterate where they stopped, and results AND iteration counts agree with the CPU restatement (which re-pivots every
    iteration) for every instance. With the repair switched off the instances report Pivot Search Fail instead of
    carrying on with a pivot the reference would have refused."""
    B = 40
    gx = np.full(B, 1e-15)
    gx[0] = 1.0
    gx[17] = 1.0
    ck = _outlier_circuit()
    o = oracle.Circuit(ck.to_text()).batch(0, B, overrides={"R:rg:g": gx})
    assert np.all(o["status"] == 0)
    c = ck.to_s21().elaborate()
    b = s21.Batch(c, B)
    b.override("R:rg:g", gx)
    x, st, it = b.dcop()
    ss = b.setup_stats()
    print("pivot repair stats:", ss, b.kernel_name())
    assert np.all(st == 0)
    assert rel_err(x, o["x"], floor=1e-9) <= 1e-9
    assert np.array_equal(it, o["iters"])
    assert ss["weak_pivot_instances"] == B - 2 and ss["repaired_instances"] == B - 2
    # the view path (what bench.py's end-to-end loop uses) repairs too
    b.reset()
    xv, stv, itv = b.dcop_view()
    assert np.array_equal(xv, x) and np.all(stv == 0) and np.array_equal(itv, it)
    monkeypatch.setenv("S21_PIVOT_REPAIR", "0")
    b2 = s21.Batch(c, B)
    b2.override("R:rg:g", gx)
    x2, st2, it2 = b2.dcop()
    s2 = b2.setup_stats()
    want = np.full(B, s21.S21_PIVOT_SEARCH_FAIL)
    want[[0, 17]] = 0
    assert np.array_equal(st2, want) and s2["repaired_instances"] == 0
    # an ordinary Monte-Carlo batch stops nobody
    monkeypatch.delenv("S21_PIVOT_REPAIR")
    b3 = s21.Batch(cc.diffpair().to_s21().elaborate(), 256)
    for k, v in cc.diffpair_mc(256).items():
        b3.override(k, v)
    b3.dcop()
    assert b3.setup_stats()["weak_pivot_instances"] == 0


def test_long_ring_needs_repivoting_matches_oracle(s21, oracle):
    """A 21-stage Mos1 ring released from one IC: its operating point is a switching wave that travels down the chain, so the
    matrix changes character from one Newton iteration to the next (devices leave cut-off one stage at a time) and the pivot
    order frozen at x = 0 meets pivots many orders of magnitude below their columns — with it alone the solve ends in NaNs.
    The kernels stop at the first factorisation the reference would have ordered differently, the host takes a new order at
    that iterate and the stopped instances are continued IN PLACE (resume launch): 110 iterations for OP + transient exactly
    as in the reference, and the waveforms agree with the oracle."""
    ck, ic = cc.inverter_array(1, 21)
    o = oracle.Circuit(ck.to_text()).tran(1e-11, 2e-10, ic=ic)
    c = ck.to_s21().elaborate(ic=ic)
    b = s21.Batch(c, 3)
    x, st, it = b.dcop()
    assert np.all(st == 0) and np.all(np.isfinite(x))
    assert np.max(np.abs(x[0] - o.data[0])) <= 1e-9 and b.setup_stats()["repaired_instances"] >= 1
    # a warm second solve (no reset) converges at once: the continued solve left x, device state and counters consistent
    x2, st2, it2 = b.dcop()
    assert np.all(st2 == 0) and np.all(it2 - it <= 1) and np.max(np.abs(x2 - x)) < 1e-9
    t, w, stt, itt = s21.Batch(ck.to_s21().elaborate(ic=ic), 2).tran(1e-11, 2e-10)
    assert np.all(stt == 0) and np.array_equal(t, o.axis)
    assert np.max(np.abs(w[0] - o.data)) <= 1e-8 and np.array_equal(w[0], w[1])
    assert int(itt[0]) == o.solves


def test_repivot_resume_leaves_other_instances_untouched(s21, oracle):
    """The in-place continuation (SolveCtl::resume) only runs the stopped instances: in a batch that mixes the outlier circuit's
    two kinds of instance, the ones that converged in the first launch keep their x, status and iteration count bit for bit
    through every repair round, and the repaired ones match the oracle; batch sizes that leave ragged warps / CTAs."""
    ck = _outlier_circuit()
    for B in (5, 40, 257):
        gx = np.full(B, 1e-15)
        gx[::3] = 1.0                                   # instance 0 is an outlier, and so is every third one
        o = oracle.Circuit(ck.to_text()).batch(0, B, overrides={"R:rg:g": gx})
        b = s21.Batch(ck.to_s21().elaborate(), B)
        b.override("R:rg:g", gx)
        x, st, it = b.dcop()
        ss = b.setup_stats()
        assert np.all(st == 0) and np.array_equal(it, o["iters"]) and rel_err(x, o["x"], floor=1e-9) <= 1e-9
        assert ss["repaired_instances"] == int(np.sum(gx < 1.0))
        ref = s21.Batch(ck.to_s21().elaborate(), B)      # all-outlier batch: nobody stops, no repair
        ref.override("R:rg:g", np.ones(B))
        xr, str_, itr = ref.dcop()
        assert ref.setup_stats()["repaired_instances"] == 0
        assert np.array_equal(x[::3], xr[::3]) and np.array_equal(it[::3], itr[::3])


def test_convergence_aids_source_and_gmin_stepping(s21, oracle):
    """SURVEY §8 f4 (opt-in; the reference has no continuation). A 61-stage ring released from ONE initial condition needs
    more than the reference's 100 Newton iterations for its operating point (the logic levels propagate ~1.4 iterations
    per stage under the 1 V step limit): "Convergence Failed" in the reference, in the oracle and here by default. With
    source stepping the failed instances are continued at 0.1 ... 1.0 of the supply, warm-started, and settle; a plain
    solve from the rescued point then needs no further iteration, and the transient that follows the rescued OP runs."""
    ck, ic = cc.inverter_array(1, 61)
    with pytest.raises(oracle.OracleError):
        oracle.Circuit(ck.to_text()).tran(1e-11, 2e-11, ic=ic)
    c = ck.to_s21().elaborate(ic=ic)
    B = 5
    vs = np.linspace(0.95, 1.05, B)
    b = s21.Batch(c, B)
    b.override("V:vsup:dc", vs)
    x0, st0, it0 = b.dcop()
    failed = st0 == s21.S21_CONVERGENCE_FAILED     # the higher supplies need more than 100 iterations (at 1.0 V the oracle fails too)
    assert failed.sum() >= 2 and failed[B // 2] and np.all(it0[failed] == 100) and np.all(st0[~failed] == 0)
    for flags in ({"source_stepping": True}, {"gmin_stepping": True, "source_stepping": True}):
        b = s21.Batch(c, B)
        b.override("V:vsup:dc", vs)
        b.set_aids(**flags)
        x, st, it = b.dcop()
        assert np.all(st == 0), st
        assert b.setup_stats()["aided_instances"] == failed.sum()
        assert np.array_equal(x[~failed], x0[~failed]) and np.array_equal(it[~failed], it0[~failed])   # the others keep their result
        n = {name: k for k, name in enumerate(c.names)}
        lv = x[:, [n[f"r0s{k}"] for k in range(61)]]
        assert np.all(np.abs(lv[:, 0]) < 1e-2)                              # the IC (a 1 S resistor to 0 V) holds stage 0 low
        assert np.all(lv[:, 1:20:2] > 0.9 * vs[:, None]) and np.all(lv[:, 2:20:2] < 0.1)   # alternating logic levels behind it
        x2, st2, it2 = b.dcop()                                              # warm restart from the rescued point
        assert np.all(st2 == 0) and np.all(it2 - it <= 1) and np.max(np.abs(x2 - x)) < 1e-9
    t, w, stt, itt = b.tran(1e-11, 2e-10, save=[n["r0s1"], n["r0s30"]])
    assert np.all(stt == 0) and np.all(np.isfinite(w))


def test_override_specs_are_validated(s21):
    """An override that can never take effect is refused when it is given (S21_ERROR), and leaves the batch usable: unknown
    kind, a name no device carries, a parameter the kind does not have, opt:gmin (Options.gmin is one value per solve)."""
    B = 4
    b = s21.Batch(cc.diffpair().to_s21().elaborate(), B)
    v = np.ones(B)
    for spec in ("resistor:r1:g", "R:nosuch:g", "R:r1:r", "V:vdd:ac", "opt::gmin", "mos1model:nosuch:vt0", "bsim4inst:i:w", "R:r1"):
        with pytest.raises(s21.Spice21Error):
            b.override(spec, v)
    x0, st0, _ = b.dcop()
    b.override("R:r1:g", np.full(B, 5e-5) * np.array([1.0, 1.1, 0.9, 1.0]))
    b.reset()
    x1, st1, _ = b.dcop()
    assert np.all(st0 == 0) and np.all(st1 == 0) and np.array_equal(x1[0], x0[0]) and not np.array_equal(x1[1], x0[1])


def test_per_instance_failure_is_contained(s21, oracle):
    """One non-converging Monte-Carlo sample must not take the batch down (per-instance status vector)."""
    B = 64
    ck = cc.diffpair()
    b = s21.Batch(ck.to_s21().elaborate(), B)
    v = np.full(B, 0.9)
    v[7] = 500.0  # 1 V per iteration step limit (analysis.rs:197-203): 100 iterations cannot reach 500 V
    b.override("V:vinp:dc", v)
    x, st, it = b.dcop()
    assert st[7] == s21.S21_CONVERGENCE_FAILED and it[7] == 100 and np.all(np.delete(st, 7) == 0)
    o = oracle.Circuit(ck.to_text()).batch(0, B, overrides={"V:vinp:dc": v})
    assert np.array_equal(o["status"], st) and np.array_equal(o["iters"], it)


def test_singular_matrix_status(s21, oracle):
    # two voltage sources in parallel: the last pivot position has no element -> "Singular Matrix" (mod.rs:969-972)
    ck = Ckt().V("v1", "a", GND, 1.0).V("v2", "a", GND, 1.0)
    with pytest.raises(oracle.OracleError) as e:
        oracle.Circuit(ck.to_text()).dcop()
    x, st, it = s21.Batch(ck.to_s21().elaborate(), 1).dcop()
    assert st[0] == e.value.status == s21.S21_SINGULAR_MATRIX
    with pytest.raises(s21.Spice21Error) as e2:  # and through the bytes API it is the reference's error string
        s21.dcop(ck.to_proto())
    assert e2.value.desc == "Singular Matrix"
    # a numerically singular matrix is NOT an error in the reference (0/0 in back-substitution): NaNs, status OK
    ck = Ckt().R("r1", "a", "b", 1e-3).C("c1", "b", GND, 1e-9)
    o = oracle.Circuit(ck.to_text()).dcop()
    x, st, it = s21.Batch(ck.to_s21().elaborate(), 1).dcop()
    assert st[0] == 0 and np.array_equal(np.isnan(x[0]), np.isnan(o.data[0]))


# ------------------------------------------------------------------------------------------------ time-varying sources (f2)
def _pulse(t, v1, v2, td, tr, tf, pw, per):
    tt = t - td
    if tt < 0.0:
        return v1
    if per > 0.0:
        tt -= per * np.floor(tt / per)
    if tt < tr:
        return v1 + (v2 - v1) * (tt / tr)
    if tt < tr + pw:
        return v2
    if tt < tr + pw + tf:
        return v2 + (v1 - v2) * ((tt - tr - pw) / tf)
    return v1


@pytest.mark.parametrize("kernel", [None, "direct", "coop", "hybrid", "jit", "jitteam"])
def test_time_varying_sources_rc_against_recurrence(s21, monkeypatch, kernel):
    """SURVEY §8 f2 (an extension: the reference's Vsrc is DC / acm only, spice21.proto:29-35): PULSE and SIN voltage sources,
    evaluated on the device at every time point. Referee: the Backward-Euler recurrence of an RC low-pass written out in
    numpy on the reference's time axis (t starts at tstep and accumulates tstep, analysis.rs:552-569),
    v_k = (v_{k-1} + a u(t_k)) / (1 + a), a = g dt / C — independent of the product and of the oracle. Every kernel family,
    builder API and protobuf wire (Vsrc extension fields 6 / 7), a batch whose instances differ in R."""
    if kernel:
        monkeypatch.setenv("S21_KERNEL", kernel)
    g, cap, tstep, tstop = 1e-3, 1e-9, 2e-8, 4e-6          # RC = 1 us
    pulse = [0.2, 1.0, 1e-7, 5e-8, 1e-7, 4e-7, 1e-6]
    sine = [0.5, 0.4, 2e6, 2e-7, 3e5]
    B = 5
    gs = g * np.linspace(0.8, 1.2, B)
    for via_proto in (False, True):
        for kind, w, dc in (("pulse", pulse, pulse[0]), ("sin", sine, sine[0])):
            ck = Ckt().V("vin", "a", GND, dc, wave=(kind, w)).R("r1", "a", "b", g).C("c1", "b", GND, cap)
            c = ck.to_s21(via_proto=via_proto).elaborate()
            b = s21.Batch(c, B)
            b.override("R:r1:g", gs)
            t, wave, st, it = b.tran(tstep, tstop)
            assert np.all(st == 0)
            if kernel in ("jit", "jitteam"):
                assert b.kernel_name() in ("jit-thread", "jit-team")
            ia, ib = c.names.index("a"), c.names.index("b")
            if kind == "pulse":
                u = np.array([dc] + [_pulse(tk, *w) for tk in t[1:]])
            else:
                tt = t - w[3]
                u = np.where(tt < 0.0, w[0], w[0] + w[1] * np.exp(-tt * w[4]) * np.sin(2.0 * np.pi * w[2] * tt))
                u[0] = dc
            assert np.max(np.abs(wave[:, :, ia] - u[None, :])) <= 1e-12
            for i in range(B):
                a = gs[i] * tstep / cap
                v = np.empty(len(t))
                v[0] = dc
                for k in range(1, len(t)):
                    v[k] = (v[k - 1] + a * u[k]) / (1.0 + a)
                assert np.max(np.abs(wave[i, :, ib] - v)) <= 1e-9, (kind, via_proto, i)
            assert np.ptp(wave[0, :, ib]) > (0.2 if kind == "pulse" else 0.03)          # the output really moves (2 MHz sine behind RC = 1 us)


def test_time_varying_source_drives_mos1_inverter_and_adaptive_steps(s21):
    """A Mos1 CMOS inverter driven by a PULSE: the output switches with the input (fixed-step transient), and the adaptive
    transient follows the same waveform with far fewer solves on the flat parts, taking small steps at the edges."""
    ck = cc.add_mos1_defaults(Ckt())
    ck.V("vdd", "vdd", GND, 1.0).V("vin", "inp", GND, 0.0, wave=("pulse", [0.0, 1.0, 2e-10, 1e-10, 1e-10, 8e-10, 2e-9]))
    ck.M("mp", "pmos", "default", d="out", g="inp", s="vdd", b="vdd").M("mn", "nmos", "default", d="out", g="inp", s=GND, b=GND)
    ck.C("cl", "out", GND, 1e-14)
    c = ck.to_s21().elaborate()
    io, ii = c.names.index("out"), c.names.index("inp")
    t, w, st, it = s21.Batch(c, 2).tran(1e-11, 4e-9, save=[ii, io])
    assert np.all(st == 0) and np.array_equal(w[0], w[1])
    hi_in = w[0, :, 0] > 0.99
    lo_in = w[0, :, 0] < 0.01
    assert hi_in.sum() > 100 and lo_in.sum() > 100
    out = w[0, :, 1]
    k_fall = int(np.argmax(hi_in))                       # first point with the input high
    k_rise = k_fall + int(np.argmax(~hi_in[k_fall:]))    # first point after it with the input leaving the high level
    # the output starts high, is pulled down while the input is high (the 1 um devices discharge 10 fF slowly) and recovers after
    assert out[0] > 0.95 and out[k_rise] < out[k_fall] - 0.4 and np.all(np.diff(out[k_fall + 5:k_rise]) < 0.0)
    assert out[-1] > out[k_rise + 20] and np.max(out[k_rise + 20:]) > out[k_rise] + 0.2
    ta, wa, sta, ita, acc, rej = s21.Batch(c, 1).tran_adaptive(1e-11, 4e-9, save=[ii, io], hmax=2e-10)
    assert sta[0] == 0 and np.max(np.abs(wa[0, :, 0] - w[0, : len(ta), 0])) < 0.05   # the print grid samples the same input
    assert np.max(np.abs(wa[0, :, 1] - w[0, : len(ta), 1])) < 0.1 and acc[0] < len(t)


# ------------------------------------------------------------------------------------------------ adaptive transient (f1)
def test_tran_adaptive_rc_step(s21):
    """LTE-controlled adaptive Backward Euler on the device (s21_batch_tran_adaptive) on an RC step with an exact answer:
    v(t) = 1 - exp(-t / RC). The step size grows as the exponential flattens (far fewer accepted steps than print points),
    the error stays within the tolerance the LTE test enforces, and a tighter trtol buys accuracy with more steps."""
    ck = Ckt().V("v1", "a", GND, 1.0).R("r1", "a", "b", 1e-3).C("c1", "b", GND, 1e-9)   # RC = 1 us
    c = ck.to_s21().elaborate(ic={"b": 0.0})
    tstep, tstop = 2e-8, 1e-5
    b = s21.Batch(c, 3)
    t, w, st, it, acc, rej = b.tran_adaptive(tstep, tstop, save=[c.names.index("b")], hmax=50 * tstep)
    assert b.kernel_name() == "direct-adaptive" and np.all(st == 0) and len(t) in (500, 501)
    exact = 1.0 - np.exp(-t / 1e-6)
    err = np.max(np.abs(w[0, :, 0] - exact))
    assert err < 3e-2, err   # trtol 7 x reltol 1e-3 per step, accumulated over the rise
    assert 10 < acc[0] < 200 and np.all(acc == acc[0]) and np.all(w[1] == w[0])      # identical instances take identical steps
    t2, w2, st2, it2, acc2, rej2 = s21.Batch(c, 1).tran_adaptive(tstep, tstop, save=[c.names.index("b")], hmax=50 * tstep, trtol=0.1)
    err2 = np.max(np.abs(w2[0, :, 0] - exact))
    assert st2[0] == 0 and err2 < 0.3 * err and acc2[0] > acc[0]
    # fixed-step BE at the print step for scale: the adaptive run with default trtol is in the same error class with fewer solves
    tf, wf, stf, itf = s21.Batch(c, 1).tran(tstep, tstop, save=[c.names.index("b")])
    assert np.max(np.abs(wf[0, :, 0] - exact)) < 3e-2 and it[0] < itf[0]


def test_tran_adaptive_ring_oscillator_sweep(s21):
    """The reference's Mos1 ring oscillator as a supply sweep: every instance runs its own time axis. Against a fixed-step
    run four times finer than the print grid, the adaptive waveforms stay within a few per cent of the swing over the first
    oscillation periods; instances differ in their step counts; rejected steps occur (the edges) and are retried."""
    ro = cc.cmos_ro3(cc.add_mos1_defaults)
    c = ro.to_s21().elaborate(ic={"1": 0.0})
    B = 33
    vs = np.linspace(0.9, 1.1, B)
    tstep, tstop = 1e-11, 1.5e-9
    b = s21.Batch(c, B)
    b.override("V:v1:dc", vs)
    t, w, st, it, acc, rej = b.tran_adaptive(tstep, tstop, save=[0, 1, 2], trtol=1.0, hmax=2 * tstep)
    assert np.all(st == 0) and w.shape == (B, len(t), 3) and np.all(np.isfinite(w))
    bf = s21.Batch(ro.to_s21().elaborate(ic={"1": 0.0}), B)
    bf.override("V:v1:dc", vs)
    tf, wf, stf, itf = bf.tran(tstep / 4, tstop, save=[0, 1, 2])
    ref = wf[:, ::4, :][:, : len(t), :]
    assert np.allclose(tf[::4][: len(t)], t, rtol=0, atol=1e-15)
    assert np.max(np.abs(w - ref)) < 0.06, float(np.max(np.abs(w - ref)))
    assert np.ptp(w[:, :, 0]) > 0.5                      # it oscillates
    assert len(set(acc.tolist())) > 1 and np.all(acc > 20) and np.sum(rej) > 0


# ------------------------------------------------------------------------------------------------ ac
def test_ac_rc_lowpass(s21, oracle):  # spice21py/tests/test_spice21.py:53-70
    ck = Ckt().R("r1", "inp", "out", 1e-3).C("c1", "out", GND, 1e-9).V("vi", "inp", GND, 1e-3, acm=1.0)
    res = s21.ac(ck.to_proto(), args=s21.protos.AcOptions(fstart=1, fstop=10**9, npts=90))
    f = s21.ac_freqs(1, 10**9, 90)
    assert len(res["inp"]) == 91 and all(v == complex(1.0, 0.0) for v in res["inp"])
    out = np.array(res["out"])
    assert np.all(np.diff(np.abs(out)) <= 0)
    assert np.max(np.abs(out - 1.0 / (1.0 + 2j * np.pi * f * 1e-9 / 1e-3))) < 1e-9
    o = oracle.Circuit(ck.to_text()).ac(fstart=1, fstop=10**9, npts=90)
    assert np.max(np.abs(out - o.get("out"))) < 1e-12


def test_ac_mos1_common_source(s21, oracle):  # tests.rs:1296-1325 (the reference asserts only "does not error")
    ck = cc.add_mos1_defaults(Ckt()).C("c1", "d", GND, 1e-9).M("m", "default", "default", d="d", g="g", s=GND, b=GND)
    ck.V("v1", "vdd", GND, 1.0).V("vg", "g", GND, 0.7, acm=1.0)
    c = ck.to_s21().elaborate()
    f = s21.ac_freqs(1, 10**6, 30)
    x, st, it = s21.Batch(c, 1).ac(f)
    o = oracle.Circuit(ck.to_text()).ac(fstart=1, fstop=10**6, npts=30)
    assert np.all(st == 0) and np.array_equal(f, o.axis)
    assert np.max(np.abs(x - o.data)) <= 1e-9 * max(1.0, np.max(np.abs(o.data)))
    x0, _, _ = s21.Batch(c, 1).ac(np.array([0.0]))  # AcOptions::default -> a single point at f = 0 (tests.rs:1246)
    assert x0.shape == (1, c.n_vars)


@pytest.mark.parametrize("pmos", [False, True], ids=["nmos", "pmos"])
def test_ac_mos1_second_referee(s21, pmos):
    """Row a18: Mos1::load_ac on the GPU against tests/mos1_referee.py, a numpy transcription of mos.rs written independently
    of oracle/ (the reference pins no AC value): DC operating point, iteration count and the complex response over 7 decades."""
    from test_oracle import _cs_amp
    ck, dense = _cs_amp(pmos)
    c = ck.to_s21().elaborate()
    x, iters, ops = dense.dcop()
    xg, st, it = s21.Batch(c, 1).dcop()
    assert st[0] == 0 and it[0] == iters
    for n in dense.names:
        assert abs(xg[0, c.names.index(n)] - x[dense.ix[n]]) <= 1e-9 * max(1.0, abs(x[dense.ix[n]])), n
    freqs = np.array([1e3, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10])
    xa = dense.ac(ops, freqs)
    xga, sta, _ = s21.Batch(c, 1).ac(freqs)
    assert np.all(sta == 0)
    for n in dense.names:
        ref = xa[:, dense.ix[n]]
        assert np.all(np.abs(xga[:, c.names.index(n)] - ref) <= 1e-9 * np.maximum(1.0, np.abs(ref))), n


def test_ac_high_gain_needs_no_step_limit(s21, monkeypatch):
    """A response above ~19 is out of reach of the reference's Newton shell from a cold start (1.0 step limit x 20
    iterations, analysis.rs:253-303; the reference gets there only by warm-starting along the sweep). The batched sweep
    solves each point directly (SolveCtl::ac_direct): gain ~ 100 comes back exact; S21_AC_NEWTON=1 restores the shell,
    which reports Convergence Failed for those points."""
    import mos1_referee as mr
    model = mr.resolve_model(0, **cc.C2_MODEL)
    ip = mr.derive(model, mr.resolve_inst(**cc.C2_INST))
    ck = Ckt().define("mos1model", "m", 0, **cc.C2_MODEL).define("mos1inst", "wl", **cc.C2_INST)
    ck.V("vd", "vdd", GND, 3.0).V("vg", "g", GND, 0.53, acm=1.0).R("rl", "vdd", "d", 5e-7).C("cl", "d", GND, 1e-13)
    ck.M("m1", "m", "wl", d="d", g="g", s=GND, b=GND)
    dense = mr.Dense([("V", "vd", "vdd", "", 3.0, 0.0), ("V", "vg", "g", "", 0.53, 1.0), ("R", "vdd", "d", 5e-7), ("C", "d", "", 1e-13),
                      ("M", model, ip, "d", "g", "", "")])
    x, iters, ops = dense.dcop()
    freqs = np.array([1.0, 1e3, 1e5, 1e7, 1e9])
    ref = dense.ac(ops, freqs, direct=True)
    assert abs(ref[0, dense.ix["d"]]) > 50.0
    c = ck.to_s21().elaborate()
    xa, st, it = s21.Batch(c, 1).ac(freqs)
    assert np.all(st == 0) and np.all(it == 1)
    for n in dense.names:
        r = ref[:, dense.ix[n]]
        assert np.all(np.abs(xa[:, c.names.index(n)] - r) <= 1e-9 * np.maximum(1.0, np.abs(r))), n
    monkeypatch.setenv("S21_AC_NEWTON", "1")
    xb, stb, itb = s21.Batch(c, 1).ac(freqs)
    assert stb[0] == s21.S21_CONVERGENCE_FAILED and stb[-1] == 0 and itb[-1] == 2  # low-frequency gain unreachable; 1 GHz point is small


def test_ac_unsupported(s21):
    c = Ckt().R("r1", "a", GND, 1e-3).I("i1", "a", GND, 1e-3).to_s21().elaborate()
    with pytest.raises(s21.Spice21Error) as e:
        s21.Batch(c, 1).ac(np.array([1.0, 10.0]))
    assert e.value.status == s21.S21_UNSUPPORTED


# ------------------------------------------------------------------------------------------------ config C5 / ragged batches
def test_ac_rc_opamp_matches_oracle(s21, oracle):
    """BASELINE config 5 (RC ladder + Mos1 op-amp) at a size the oracle finishes in a fraction of a second."""
    ck = cc.rc_opamp(64)
    c = ck.to_s21().elaborate()
    f = s21.ac_freqs(1, 10**10, 999)
    x, st, it = s21.Batch(c, 1).ac(f)
    o = oracle.Circuit(ck.to_text()).ac(fstart=1, fstop=10**10, npts=999)
    assert np.all(st == 0) and np.array_equal(f, o.axis) and c.names == o.names
    # SPICE reltol / vntol (north_star tolerance for ac): |dx| <= 1e-3 |x| + 1e-6. The reference warm-starts every point
    # from the previous frequency and accepts it untouched while |res| < 1e-9 A (analysis.rs:271-279), so on high-impedance
    # nodes its own answer lags by up to ~1e-3 V; the batched solve starts every point cold and is exact.
    assert np.all(np.abs(x - o.data) <= 1e-3 * np.abs(o.data) + 1e-6)
    # exactness: single-point sweeps make the reference start cold too -> agreement to round-off
    for fq in (1, 10, 1000, 10**5, 10**7, 10**9):
        o1 = oracle.Circuit(ck.to_text()).ac(fstart=fq, fstop=fq, npts=1)
        x1, st1, _ = s21.Batch(c, 1).ac(np.array([float(fq)]))
        assert st1[0] == 0 and o1.data.shape[0] == 1
        assert np.max(np.abs(x1[0] - o1.data[0])) <= 1e-12 * max(1.0, np.max(np.abs(o1.data[0])))


def test_ac_rc_opamp_full_sweep_properties(s21, oracle):
    """BASELINE config 5 at full size: 100 000 frequency points (end-inclusive sweep, analysis.rs:791-819)."""
    ck = cc.rc_opamp(64)
    c = ck.to_s21().elaborate()
    f = s21.ac_freqs(1, 10**10, 99999)
    assert len(f) == 100000
    b = s21.Batch(c, 1)
    x, st, it = b.ac(f)
    n = {name: k for k, name in enumerate(c.names)}
    assert np.all(st == 0) and np.all(it >= 1) and np.all(it <= 3)  # linear problem: 1-2 solves per point
    assert np.all(x[:, n["l0"]] == 1.0 + 0j)                          # the driven node is exact
    out = np.abs(x[:, n["out"]])
    assert abs(out[0] - 1.0) < 1e-3 and np.max(out) < 1.2 and out[-1] < 1e-12  # unity-gain buffer behind a low-pass ladder
    lad = np.abs(x[:, [n[f"l{k}"] for k in range(65)]])
    assert np.all(np.diff(lad, axis=1) <= 1e-12)                     # magnitude decays monotonically along the ladder
    # the first points of the long sweep are the same frequencies the reference would visit: compare with the oracle
    o = oracle.Circuit(ck.to_text()).ac(fstart=1, fstop=10**10, npts=99999, max_points=1500)
    assert np.array_equal(o.axis, f[:1500]) and np.all(np.abs(x[:1500] - o.data) <= 1e-3 * np.abs(o.data) + 1e-6)
    # every 10th point of the sweep (10 000 points) against the oracle's cold-start solve at the same frequency
    # (oracle ac_at = what a one-point reference sweep computes there): round-off agreement
    oa = oracle.Circuit(ck.to_text()).ac_at(f[::10])
    assert oa.data.shape == (10000, c.n_vars)
    assert np.max(np.abs(x[::10] - oa.data) / np.maximum(1.0, np.abs(oa.data))) <= 1e-9
    # frequency points are independent: a shuffled batch gives the shuffled answer bit for bit
    # (2048 points: below the switch to the thread-per-point kernel, so this also compares two kernel families)
    perm = np.random.default_rng(5).permutation(2048)
    x2, st2, _ = b.ac(f[:2048][perm])
    assert b.kernel_name() != "direct" and np.array_equal(x2, x[:2048][perm])


@pytest.mark.parametrize("B", [1, 31, 33, 257])
def test_ragged_batch_sizes(s21, oracle, B):
    """Batch sizes that do not fill a warp / a CTA."""
    ck, ovr = cc.diffpair(), cc.diffpair_mc(B)
    o = oracle.Circuit(ck.to_text()).batch(0, B, overrides=ovr)
    b = s21.Batch(ck.to_s21().elaborate(), B)
    for k, v in ovr.items():
        b.override(k, v)
    x, st, it = b.dcop()
    assert np.all(st == 0) and rel_err(x, o["x"], floor=1e-9) <= 1e-9 and np.array_equal(it, o["iters"])


@pytest.mark.parametrize("kernel", ["direct", "coop", "hybrid", "jit", "jitteam", "jitteam:2", "jitteam:4", "jitteam:8", "jitteam:10", "jitteam:16"])
def test_kernel_variants_bit_identical(s21, kernel, monkeypatch):
    """The Newton kernels (one thread per instance, CTA-cooperative, hybrid, and the two run-time specialised shapes)
    perform the same operations in the same order per value: identical bits, identical iteration counts — dcop and
    transient. B = 83 leaves a ragged last CTA / warp in every kernel."""
    B = 83
    ck, ovr = cc.diffpair(), cc.diffpair_mc(B)

    def run(k):
        k, _, lpi = k.partition(":")
        monkeypatch.setenv("S21_KERNEL", k)
        if lpi:
            monkeypatch.setenv("S21_TEAM_LPI", lpi)
        b = s21.Batch(ck.to_s21().elaborate(), B)
        for key, v in ovr.items():
            b.override(key, v)
        x, st, it = b.dcop()
        ro = cc.cmos_ro3(cc.add_mos1_defaults)
        b2 = s21.Batch(ro.to_s21().elaborate(ic={"1": 0.0}), 3)
        t, w, st2, it2 = b2.tran(1e-11, 5e-10)
        names = (b.kernel_name(), b2.kernel_name())
        return (x, st, it, w, it2), names

    (ref, _), (got, names) = run("direct"), run(kernel)
    for a, b_ in zip(ref, got):
        assert np.array_equal(a, b_)
    if kernel.startswith("jit"):  # the forced specialised shape really ran (the others choose by circuit and batch size)
        want = "jit-thread" if kernel == "jit" else "jit-team"
        assert names == (want, want)


def test_device_division_is_exact(s21):
    """csrc/scalar.h splits f64 division into a shareable reciprocal and a quotient step (plus an exact shortcut for zero
    numerators). It must be the IEEE quotient, bit for bit, on every operand class: 3 x 2^25 quotients against `a / b`."""
    bad, first = s21.selftest_div(1 << 25, seed=20261017)
    assert bad == 0, f"a={first[0]!r} b={first[1]!r} ours={first[2]!r} ieee={first[3]!r}"


def test_team_kernel_devices_and_limits(s21, monkeypatch):
    """The team-shaped specialised kernel on every device type it accepts (R, C, I, V, Diode, Mos0, Mos1; dcop and
    transient), bit for bit against the direct kernel; and its refusal of circuits it is not generated for."""
    def circuits():
        d = cc.add_diode_defaults(Ckt(signals=["a", "b"])).V("v", "a", GND, 0.75).R("r", "a", "b", 1e-2).D("d1", "b", GND, "default", "default")
        d.C("c1", "b", GND, 1e-12).I("i1", GND, "b", 1e-4)
        yield "diode", d, None, (1e-10, 2e-9)  # its transient hits the 100-iteration cap, as in the reference: same status, NaN rows
        yield "mos0", cc.nmos_ro3(cc.add_mos0_defaults), {"1": 0.0}, (1e-11, 3e-10)
        yield "mos1", cc.pmos_ro3(cc.add_mos1_defaults), {"1": 0.0}, (1e-11, 3e-10)

    for name, ck, ic, (tstep, tstop) in circuits():
        out = {}
        for k in ("direct", "jitteam"):
            monkeypatch.setenv("S21_KERNEL", k)
            bd = s21.Batch(ck.to_s21().elaborate(), 5)
            x, st, it = bd.dcop()
            bt = s21.Batch(ck.to_s21().elaborate(ic=ic) if ic else ck.to_s21().elaborate(), 5)
            t, w, st2, it2 = bt.tran(tstep, tstop)
            out[k] = (x, st, it, w, st2, it2)
            if k == "jitteam":
                assert bd.kernel_name() == "jit-team" and bt.kernel_name() == "jit-team", name
        for a, b_ in zip(out["direct"], out["jitteam"]):
            assert np.array_equal(a, b_, equal_nan=True), name
    monkeypatch.setenv("S21_KERNEL", "jitteam")
    with pytest.raises(s21.Spice21Error):  # N = 68 > 16 rows
        s21.Batch(cc.rc_opamp(8).to_s21().elaborate(), 2).dcop()


def test_dcop_view_matches_dcop(s21):
    """s21_batch_dcop_view hands out the library's pinned result buffer instead of copying into caller arrays: same values,
    read-only, and refreshed by the next solve."""
    B = 100
    b = s21.Batch(cc.diffpair().to_s21().elaborate(), B)
    for key, v in cc.diffpair_mc(B).items():
        b.override(key, v)
    x, st, it = b.dcop()
    b.reset()
    xv, stv, itv = b.dcop_view()
    assert xv.shape == (B, b.N) and not xv.flags.writeable
    assert np.array_equal(x, xv) and np.array_equal(st, stv) and np.array_equal(it, itv)
    b.reset()
    _, st2, it2 = b.dcop_view(want_x=False)
    assert np.array_equal(st, st2) and np.array_equal(it, it2)
    # the one-call sweep step of bench.py's end-to-end loop (forced parameter upload + cold start + solve + view): the team
    # kernel writes the result rows into the pinned host buffer itself; same values, and a warm second step converges at once
    xs, sts, its, h2d = b.step_dcop_view(upload=True, reset=True)
    assert b.kernel_name() == "jit-team" and h2d > 0
    assert np.array_equal(x, xs) and np.array_equal(st, sts) and np.array_equal(it, its)
    xw, stw, itw, h2w = b.step_dcop_view(upload=False, reset=False)
    assert h2w == 0 and np.all(stw == 0) and np.all(itw - it <= 1) and np.max(np.abs(xw - x)) < 1e-9
    # a parameter change between steps is picked up by the step's upload
    b.override("R:r1:g", cc.diffpair_mc(B)["R:r1:g"] * 1.05)
    xc, stc, itc, _ = b.step_dcop_view(upload=True, reset=True)
    b2 = s21.Batch(cc.diffpair().to_s21().elaborate(), B)
    for key, v in cc.diffpair_mc(B).items():
        b2.override(key, v if key != "R:r1:g" else v * 1.05)
    x2, st2b, it2b = b2.dcop()
    assert np.array_equal(np.array(xc), x2) and np.array_equal(np.array(itc), it2b)


def _team_cases(s21, monkeypatch, variants):
    """Direct kernel vs the team kernel generated under each (S21_TEAM_FAST, S21_TEAM_WP, S21_TEAM_GI) variant: C2 dcop at a
    ragged batch size, a circuit whose L entries leave the fast quotient's domain (forces the exact redo), and — when asked —
    the C1 ring as a transient supply sweep."""
    tiny = Ckt(signals=["a", "b", "c"]).V("v", "a", GND, 1.0).R("r1", "a", "b", 1e-3).R("rt", "b", "c", 1e-300).R("r2", "c", GND, 1e-3)
    tiny.R("r3", "b", GND, 2e-3)
    B = 83
    dp, ovr = cc.diffpair(), cc.diffpair_mc(B)
    ro = cc.cmos_ro3(cc.add_mos1_defaults)

    def run(kernel, fast="1", wp=None, gi=None, with_tran=False):
        monkeypatch.setenv("S21_KERNEL", kernel)
        monkeypatch.setenv("S21_TEAM_FAST", fast)
        for key, val in (("S21_TEAM_WP", wp), ("S21_TEAM_GI", gi)):
            if val is None:
                monkeypatch.delenv(key, raising=False)
            else:
                monkeypatch.setenv(key, val)
        b = s21.Batch(dp.to_s21().elaborate(), B)
        for key, v in ovr.items():
            b.override(key, v)
        bt = s21.Batch(tiny.to_s21().elaborate(), 5)
        out, names = b.dcop() + bt.dcop(), [b.kernel_name(), bt.kernel_name()]
        if with_tran:
            br = s21.Batch(ro.to_s21().elaborate(ic={"1": 0.0}), 37)
            br.override("V:v1:dc", np.linspace(0.9, 1.1, 37))
            t, w, st2, it2 = br.tran(1e-11, 4e-10)
            out, names = out + (w, st2, it2), names + [br.kernel_name()]
        if wp == "1":  # the kernel-written result rows, through the zero-copy read the bench's end-to-end loop uses
            b.reset()
            xv, stv, itv = b.dcop_view()
            assert np.array_equal(xv, out[0]) and np.array_equal(stv, out[1]) and np.array_equal(itv, out[2])
        return out, names

    with_tran = any(v[3] for v in variants)
    ref, _ = run("direct", with_tran=with_tran)
    assert np.all(ref[1] == 0) and np.all(ref[4] == 0)
    for fast, wp, gi, tr in variants:
        got, names = run("jitteam", fast, wp, gi, tr)
        assert all(n == "jit-team" for n in names)
        for a, b_ in zip(ref, got):
            assert np.array_equal(a, b_), (fast, wp, gi)


def test_team_kernel_fast_and_exact_text_agree(s21, monkeypatch):
    """The team kernel's linear algebra is generated twice (host/jit_team.hpp): a branch-free text whose divisions defer
    their exceptions, and the exact text it falls back to. Both, and the fallback itself — forced here by a 1e-300 S
    resistor, whose L entries lie below the fast quotient's domain — give the direct kernel's bits."""
    _team_cases(s21, monkeypatch, [("1", None, None, False), ("0", None, None, False)])


def test_team_kernel_warp_private(s21, monkeypatch):
    """The warp-private team kernel (default since round 2; S21_TEAM_WP=0 = the CTA-wide evaluation phase of round 1): no
    block barriers, shared evaluation text for same-type devices, CTAs down to one warp, result rows written by the kernel.
    Same bits as the direct kernel for dcop, the forced exact redo and a transient, with both texts and several CTA sizes."""
    _team_cases(s21, monkeypatch, [("1", "1", None, True), ("0", "1", None, True), ("1", "1", "16", True), ("1", "1", "8", True),
                                   ("1", "0", "16", True), ("1", "0", None, True)])


def test_strong_scaling_shard_shape_same_bits(s21, monkeypatch):
    """Shards of <= 2048 instances take the 8-lane / 16-instance shape of the team kernel (host/jit_team.hpp::team_lpi):
    same bits as the 2-lane shape a full-size batch uses."""
    B = 1024
    dp, ovr = cc.diffpair(), cc.diffpair_mc(B)

    def run(lpi):
        if lpi:
            monkeypatch.setenv("S21_TEAM_LPI", lpi)
        else:
            monkeypatch.delenv("S21_TEAM_LPI", raising=False)
        b = s21.Batch(dp.to_s21().elaborate(), B)
        for k, v in ovr.items():
            b.override(k, v)
        return b.dcop(), b.kernel_name()

    (x8, st8, it8), name = run(None)
    (x2, st2, it2), _ = run("2")
    assert name == "jit-team" and np.all(st8 == 0)
    assert np.array_equal(x8, x2) and np.array_equal(it8, it2)


# ------------------------------------------------------------------------------------------------ Bsim4
def _bsim4_amp(nsel=None, psel=None, inst=None):
    """A CMOS stage with a resistive load and caps, Bsim4 cards given through the C ABI (the wire format carries none)."""
    c = Ckt().define("bsim4model", "n", 0, **(nsel or {})).define("bsim4model", "p", 1, **(psel or nsel or {}))
    c.define("bsim4inst", "i", **(inst or {"l": 1e-6, "w": 4e-6}))
    c.M("mp", "p", "i", d="out", g="inp", s="vdd", b="vdd").M("mn", "n", "i", d="out", g="inp", s=GND, b=GND)
    c.V("vi", "src", GND, 0.45).R("rg", "src", "inp", 1e-4).V("vd", "vdd", GND, 1.0).R("rl", "out", GND, 1e-6).C("cl", "out", GND, 1e-14)
    return c


BSIM4_IC = {"inp": 0.9, "out": 0.6}  # released at t = 0: gate and drain both move, so every charge term is exercised


def test_bsim4_known_answers(s21):  # bsim4/tests.rs:57-160
    c = cc.add_bsim4_defaults(Ckt()).M("bsim4", "default", "default", d="gd", g="gd", s=GND, b=GND)
    c.V("v1", "gd", GND, 1.0).R("r1", "gd", GND, 1e-10)
    v = s21.dcop(c.to_proto())
    assert v["gd"] == 1.0 and abs(abs(v["v1"]) - 150e-6) < 1e-6
    c = cc.add_bsim4_defaults(Ckt()).M("bsim4", "pmos", "default", d="gd", g="gd", s=GND, b=GND)
    c.V("v1", "gd", GND, -1.0).R("r1", "gd", GND, 1e-10)
    v = s21.dcop(c.to_proto())
    assert v["gd"] == -1.0 and abs(abs(v["v1"]) - 57e-6) < 1e-6
    c = cc.add_bsim4_defaults(Ckt())
    c.M("p", "pmos", "default", d="d", g="inp", s="vdd", b="vdd").M("n", "nmos", "default", d="d", g="inp", s=GND, b=GND)
    c.V("vinp", "inp", GND, 0.0).V("vvdd", "vdd", GND, 1.0)
    v = s21.dcop(c.to_proto())
    assert v["vdd"] == 1.0 and v["inp"] == 0.0 and v["d"] > 0.95 and abs(v["vinp"]) < 1e-6 and abs(v["vvdd"]) < 1e-6


BSIM4_VARIANTS = [
    {}, {"rgatemod": 1}, {"rgatemod": 2}, {"rgatemod": 3}, {"rbodymod": 1}, {"rdsmod": 1}, {"trnqsmod": 1},
    {"rgatemod": 3, "rbodymod": 1, "rdsmod": 1, "trnqsmod": 1},
    {"capmod": 0}, {"capmod": 1}, {"capmod": 0, "xpart": 1.0}, {"capmod": 1, "xpart": 0.5}, {"xpart": 1.0}, {"cvchargemod": 1},
    {"mobmod": 1}, {"mobmod": 2}, {"mobmod": 3}, {"mobmod": 4}, {"mobmod": 5}, {"mobmod": 6},
    {"igcmod": 1, "igbmod": 1}, {"igcmod": 2, "igbmod": 1, "pigcd": 1.0}, {"gidlmod": 1}, {"diomod": 0}, {"diomod": 2},
    {"tempmod": 2, "tnom": 40.0}, {"mtrlmod": 1}, {"vtl": 2.0e5}, {"lambda": 2.0e-5}, {"dvtp0": 1e-8, "dvtp4": 0.5, "dvtp2": 0.1},
    {"jss": 1e-4, "jsd": 1e-4, "cjs": 5e-4, "cjd": 5e-4, "cjsws": 5e-10, "cjswd": 5e-10, "jtss": 1e-4, "jtsd": 1e-4},
]


@pytest.mark.parametrize("sel", BSIM4_VARIANTS, ids=["_".join(f"{k}{v}" for k, v in s_.items()) or "default" for s_ in BSIM4_VARIANTS])
def test_bsim4_variants_match_oracle(s21, oracle, sel):
    """Every selector branch of the model: dcop and a short transient on the GPU against the oracle.

    What this does and does not prove: the oracle compiles the SAME Bsim4 equation headers for the CPU (oracle/bsim4.hpp
    includes spice21_b200/csrc/bsim4/*.hpp — stated there and in DESIGN.md §2), so this test compares one source compiled
    twice. It catches CUDA-compilation, kernel, assembly and solver faults on every selector branch; it cannot catch a
    transcription fault in the equations themselves. The equations are pinned to the reference only where the reference
    holds numbers — its three Bsim4 golden transients and three known answers, all on the default card path
    (test_golden_waveforms, test_bsim4_known_answers) — and, independently of any transcription, by the conservation and
    DC invariants of test_bsim4_terminal_current_invariants below."""
    ck = _bsim4_amp(sel)
    o = oracle.Circuit(ck.to_text())
    c = ck.to_s21().elaborate()
    b = s21.Batch(c, 1)
    x, status, iters = b.dcop()
    od = o.dcop()
    assert status[0] == 0
    # both sides stop on the reference's criterion (|dx| < 1e-3, |res| < 1e-12 A); with 1e3 S series conductances next to
    # 1e-6 S loads (rdsmod / rbodymod networks) last-bit libm differences are amplified to ~1e-9 relative
    assert close(x[0], od.data[0], rtol=1e-8)
    ot = oracle.Circuit(ck.to_text()).tran(2e-11, 2e-9, ic=BSIM4_IC)
    t, wave, status, _ = s21.Batch(ck.to_s21().elaborate(ic=BSIM4_IC), 1).tran(2e-11, 2e-9)
    assert status[0] == 0 and wave.shape[1] == ot.data.shape[0]
    assert np.max(np.abs(wave[0] - ot.data)) <= 1e-9


def _bsim4_four_terminal(sel, typ, bias):
    c = Ckt().define("bsim4model", "m", typ, **sel).define("bsim4inst", "i", l=1e-6, w=4e-6)
    c.M("m1", "m", "i", d="d", g="g", s="s", b="b")
    for n, v in zip("dgsb", bias):
        c.V("v" + n, n, GND, v)
    return c


def test_bsim4_terminal_current_invariants(s21):
    """A check of the Bsim4 stamps that needs no second transcription of the equations: one device driven by four voltage
    sources, every selector variant, NMOS and PMOS, forward and source/drain-swapped bias.
    (1) Charge conservation of the stamp: the four source currents sum to zero (up to the gmin leakage, ~1e-12 A) — a
        conductance or current term stamped on a wrong node, with a wrong sign or without its partner breaks it.
    (2) DC invariants of the network options: a gate resistance (rgatemod 1-3) carries no DC current, so every terminal
        current equals the rgatemod = 0 value; a body network (rbodymod = 1) only carries junction leakage (1e-9 relative);
        source/drain series resistances (rdsmod = 1, ~1e-3 ohm-scale here) move the drain current by < 1e-5 relative.
    (3) Thin-oxide gate tunnelling (igcmod / igbmod, toxe = 1.2 nm): a measurable gate current appears, flows into the gate
        for an NMOS with vg > vs, vd, and conservation still holds."""
    def currents(sel, typ, bias):
        c = _bsim4_four_terminal(sel, typ, bias).to_s21().elaborate()
        x, st, it = s21.Batch(c, 1).dcop()
        assert st[0] == 0, (sel, typ, bias)
        return np.array([x[0, c.names.index("v" + n)] for n in "dgsb"])

    for typ, bias in ((0, (0.7, 0.9, 0.05, -0.1)), (1, (-0.7, -0.9, -0.05, 0.1)), (0, (0.05, 0.9, 0.7, -0.1))):
        base = currents({}, typ, bias)
        assert abs(base[0]) > 1e-5 and abs(base.sum()) < 1e-11
        for sel in BSIM4_VARIANTS:
            cur = currents(sel, typ, bias)
            assert abs(cur.sum()) < 1e-11, (sel, typ, cur)
            if set(sel) <= {"rgatemod", "trnqsmod"}:
                assert np.max(np.abs(cur - base)) <= 1e-12 * np.max(np.abs(base)) + 1e-15, (sel, cur, base)
            elif set(sel) == {"rbodymod"}:
                assert np.max(np.abs(cur - base)) <= 1e-8 * np.max(np.abs(base)) + 1e-11, (sel, cur, base)  # + gmin leakage of the network
            elif "rdsmod" in sel:
                assert np.max(np.abs(cur - base)) <= 1e-5 * np.max(np.abs(base)), (sel, cur, base)
    thin = {"igcmod": 1, "igbmod": 1, "toxe": 1.2e-9, "toxp": 1.2e-9, "toxm": 1.2e-9}
    cur = currents(thin, 0, (0.7, 0.9, 0.05, -0.1))
    off = currents(dict(thin, igcmod=0, igbmod=0), 0, (0.7, 0.9, 0.05, -0.1))
    assert abs(cur.sum()) < 1e-11 and abs(off[1]) < 1e-15
    assert -cur[1] > 1e-10    # current flows from the source vg INTO the gate: the branch current of vg is negative


def test_bsim4_model_cards_on_the_wire(s21):
    """SURVEY §8 f2, second half: the reference's `Bsim4Model` message carries only `mos_type` (bsim4.proto:45-50, its 876 model
    fields are commented out), so a full card cannot reach `dcop` / `tran` through the bytes API. The product reads them from an
    extension field (901: repeated name / value, skipped by the reference's decoder): a circuit with non-default cards sent as
    protobuf bytes gives the same bits as the same circuit built through s21_ckt_define, through the structured path and
    through s21_op_bytes."""
    sel = {"mobmod": 1, "rgatemod": 1, "rbodymod": 1, "igcmod": 1, "igbmod": 1, "toxe": 2.0e-9, "toxp": 2.0e-9, "toxm": 2.0e-9, "vth0": 0.35}
    psel = dict(sel, vth0=-0.35)
    ck = _bsim4_amp(sel, psel)
    x_api, st_api, _ = s21.Batch(ck.to_s21().elaborate(), 1).dcop()
    c_wire = ck.to_s21(via_proto=True).elaborate()
    x_wire, st_wire, _ = s21.Batch(c_wire, 1).dcop()
    assert st_api[0] == 0 and st_wire[0] == 0 and np.array_equal(x_api, x_wire)
    plain = _bsim4_amp()
    x_plain, _, _ = s21.Batch(plain.to_s21().elaborate(), 1).dcop()
    assert x_plain.shape != x_api.shape or not np.array_equal(x_plain, x_api)   # the cards really arrived (internal nodes, other currents)
    res = s21.dcop(ck.to_proto())                                                  # bytes in, OpResult out
    names = c_wire.names
    assert abs(res["out"] - x_wire[0, names.index("out")]) == 0.0 and len(res) == len(names)


def test_bsim4_instance_sweep_matches_oracle(s21, oracle):
    """Per-instance Bsim4 cards (config C4's sweep axis): width / length / threshold shift differ per instance."""
    B = 96
    rng = np.random.default_rng(4)
    ck = _bsim4_amp()
    w = 4e-6 * (1.0 + 0.2 * rng.standard_normal(B).clip(-2, 2))
    l = 1e-6 * (1.0 + 0.1 * rng.standard_normal(B).clip(-2, 2))
    dv = 0.03 * rng.standard_normal(B)
    ovr = {"bsim4inst:i:w": w, "bsim4inst:i:l": l, "bsim4inst:i:delvto": dv}
    b = s21.Batch(ck.to_s21().elaborate(), B)
    for k, v in ovr.items():
        b.override(k, v)
    x, status, iters = b.dcop()
    o = oracle.Circuit(ck.to_text()).batch(0, B, overrides=ovr, nthreads=4)
    assert np.all(status == 0) and np.all(o["status"] == 0)
    assert close(x, o["x"].reshape(x.shape))
    out = b.ckt.names.index("out")
    assert len(np.unique(np.round(x[:, out], 9))) > B // 2  # the sweep really differs per instance


@pytest.mark.parametrize("kernel", ["direct", "coop", "hybrid"])
def test_bsim4_kernel_variants_bit_identical(s21, kernel, monkeypatch):
    ck = cc.cmos_ro3(cc.add_bsim4_defaults)
    ref = s21.Batch(ck.to_s21().elaborate(ic={"1": 0.0}), 3).tran(1e-10, 2e-8)
    monkeypatch.setenv("S21_KERNEL", kernel)
    got = s21.Batch(ck.to_s21().elaborate(ic={"1": 0.0}), 3).tran(1e-10, 2e-8)
    assert np.array_equal(ref[1], got[1]) and np.array_equal(ref[3], got[3])


def test_c4_bsim4_ring_sweep_matches_oracle(s21, oracle):
    """Config C4 (BASELINE.json configs[3]) at a size the oracle finishes in seconds: 21-stage BSIM4 ring, VDD sweep."""
    B, npts, tstep = 24, 60, 1e-10
    ck, ic = cc.bsim4_ring(21)
    ovr = {k: v[::85][:B] for k, v in cc.c4_sweep(2048).items()}  # 24 instances spread over the VDD axis
    c = ck.to_s21().elaborate(ic=ic)
    b = s21.Batch(c, B)
    for k, v in ovr.items():
        b.override(k, v)
    t, wave, status, iters = b.tran(tstep, npts * tstep)
    o = oracle.Circuit(ck.to_text()).batch(1, B, overrides=ovr, tstep=tstep, tstop=npts * tstep, ic=ic, nthreads=8)
    assert np.all(status == 0) and np.all(o["status"] == 0)
    assert wave.shape == o["x"].shape
    # free-running oscillators amplify last-bit differences along the waveform; the reference's own golden tolerance is 1e-6
    assert np.max(np.abs(wave - o["x"])) <= 1e-7
    vdd = c.names.index("vdd")
    assert np.allclose(wave[:, -1, vdd], ovr["V:vsup:dc"], rtol=0, atol=1e-12)  # every instance really ran at its own supply


def test_bsim4_fast_division_within_tolerance(s21, oracle, monkeypatch):
    """S21_B4_FAST=1 (kernels/coop_fast.cu): the cooperative kernel with the Bsim4 evaluation's divisions as a * rcp(b) —
    quotients within ~1.5 ulp instead of correctly rounded. Not bit-identical to the default kernels (which is why it is
    opt-in); it must stay inside BASELINE.json's bounds — dcop 1e-9 relative, transient SPICE reltol 1e-3 / vntol 1e-6 —
    against the oracle, on config C4's sweep (24 instances x 60 points), on a dcop of the Bsim4 amplifier and on the
    reference's own Bsim4 ring-oscillator golden waveform (reference tolerance 1e-6, tests.rs:948-964)."""
    monkeypatch.setenv("S21_B4_FAST", "1")
    monkeypatch.setenv("S21_KERNEL", "coop")
    # dcop
    ck = _bsim4_amp()
    b = s21.Batch(ck.to_s21().elaborate(), 4)
    x, status, _ = b.dcop()
    assert b.kernel_name() == "coop-rcp" and np.all(status == 0)
    od = oracle.Circuit(ck.to_text()).dcop()
    assert close(x[0], od.data[0], rtol=1e-9)
    # the reference's golden transient
    g = golden("test_bsim4_cmos_ro_tran")
    rb = cc.cmos_ro3(cc.add_bsim4_defaults)
    cb = rb.to_s21().elaborate(ic={"1": 0.0})
    bb = s21.Batch(cb, 1)
    tb, wb, stb, _ = bb.tran(1e-10, 3e-7)
    assert bb.kernel_name() == "coop-rcp" and stb[0] == 0 and len(tb) == len(g["time"])
    for k, name in enumerate(cb.names):
        assert np.max(np.abs(wb[0, :, k] - g[name])) <= 1e-6, name
    monkeypatch.delenv("S21_KERNEL")
    # C4 sweep
    B, npts, tstep = 24, 60, 1e-10
    ck, ic = cc.bsim4_ring(21)
    ovr = {k: v[::85][:B] for k, v in cc.c4_sweep(2048).items()}
    bf = s21.Batch(ck.to_s21().elaborate(ic=ic), B)
    for k, v in ovr.items():
        bf.override(k, v)
    t, wave, status, iters = bf.tran(tstep, npts * tstep)
    assert bf.kernel_name() == "coop-rcp"
    monkeypatch.delenv("S21_B4_FAST")
    bd = s21.Batch(ck.to_s21().elaborate(ic=ic), B)
    for k, v in ovr.items():
        bd.override(k, v)
    td, wd, sd, itd = bd.tran(tstep, npts * tstep)
    assert bd.kernel_name() == "coop"
    o = oracle.Circuit(ck.to_text()).batch(1, B, overrides=ovr, tstep=tstep, tstop=npts * tstep, ic=ic, nthreads=8)
    okp = status == 0
    print(f"fast division: status {status.tolist()}, max |fast - default| = {float(np.max(np.abs(wave[okp] - wd[okp]))):.3e}, "
          f"max |fast - oracle| = {float(np.max(np.abs(wave[okp] - o['x'][okp]))):.3e}, Newton iterations "
          f"{int(iters[okp].sum())} vs {int(itd[okp].sum())}")
    # (before the OP repair kept instances that are singular under ANOTHER instance's pivot order open, instance 17 of this
    # sweep ended its operating point with Singular Matrix under this kernel's last bits — resolve_repivot, host/batch.hpp)
    ok = status == 0
    assert np.all(o["status"] == 0) and np.all(sd == 0) and np.all(ok)
    tol = 1e-6 + 1e-3 * np.abs(o["x"][ok])
    assert np.all(np.abs(wave[ok] - o["x"][ok]) <= tol)
    assert np.max(np.abs(wave[ok] - o["x"][ok])) <= 1e-6  # in fact far inside vntol on this circuit


def test_c4_ptm65_statuses_match_oracle(s21, oracle):
    """SURVEY's literal C4 cards (PTM 65 nm): under the reference's Newton loop some supply voltages do not converge.
    The GPU path must report the same per-instance outcome as the oracle, and the same waveforms where it converges."""
    B, npts, tstep = 17, 30, 1e-11
    ck, ic = cc.bsim4_ring(7, cards="ptm65", l=65e-9, wn=200e-9, wp=400e-9)
    ovr = {"V:vsup:dc": np.linspace(0.8, 1.2, B)}
    c = ck.to_s21().elaborate(ic=ic)
    b = s21.Batch(c, B)
    for k, v in ovr.items():
        b.override(k, v)
    t, wave, status, iters = b.tran(tstep, npts * tstep)
    o = oracle.Circuit(ck.to_text()).batch(1, B, overrides=ovr, tstep=tstep, tstop=npts * tstep, ic=ic, nthreads=8)
    assert np.array_equal(status != 0, o["status"] != 0)
    assert 0 < int(np.sum(status == 0)) < B  # the case is only interesting while it mixes outcomes
    ok = status == 0
    assert np.max(np.abs(wave[ok] - o["x"][ok])) <= 1e-9


def test_c4x_41_stage_ring_with_internal_nodes_matches_oracle(s21, oracle):
    """Config C4 towards SURVEY's size. Under the reference's own Newton loop (100 iterations, no continuation) a single-IC
    Bsim4 ring stops converging beyond ~31 stages and the 101-stage ring fails with any IC spacing (checked with the
    oracle, DESIGN.md "C4"); 41 stages with a second IC at stage 20 is the longest ring that converges over the whole
    0.8-1.2 V sweep. With rbodymod = rgatemod = 1 every device gets its internal gate and body nodes
    (bsim4ports.rs:60-88): 82 devices, N = 373 — the cooperative kernel's large-circuit regime (HBM-resident staging,
    hundreds of dependency levels). On those cards half of the chosen supply voltages fail inside the transient under the
    reference's loop (a Newton iteration that cycles until the 100-iteration cap): where both sides converge the waveforms
    and iteration counts must agree; a cycling iteration is sensitive to the last bits (the frozen pivot order rounds
    differently from the reference's per-iteration order), so single instances may fall on the other side of the cap —
    the outcome must agree for at least 14 of the 16 (measured: all 16), and no instance may return a wrong waveform as converged.
    The plain-card 41-stage ring (N = 47) converges at every supply in the reference, and here."""
    ck, ic = cc.bsim4_ring(41, ic_every=20, rbodymod=1, rgatemod=1)
    c = ck.to_s21().elaborate(ic=ic)
    assert c.n_vars == 375 and sorted(ic) == ["s0", "s20"]   # 373 circuit unknowns + the second IC's source branch pair
    B, npts, tstep = 16, 20, 1e-10
    ovr = {"V:vsup:dc": np.linspace(0.8, 1.2, 64)[[0, 8, 12, 16, 17, 24, 30, 33, 37, 40, 47, 50, 54, 60, 62, 63]]}  # 8 of them fail in the oracle
    b = s21.Batch(c, B)
    for k, v in ovr.items():
        b.override(k, v)
    save = [c.names.index(n) for n in ("s1", "s20", "s21", "s40", "vdd")]
    t, wave, status, iters = b.tran(tstep, npts * tstep, save=save)
    o = oracle.Circuit(ck.to_text()).batch(1, B, overrides=ovr, tstep=tstep, tstop=npts * tstep, ic=ic, nthreads=8)
    assert b.kernel_name() == "coop" and b.plan_info()["n"] == 375
    # Inside the device-resident time loop nobody re-pivots; when the frozen order meets a vanishing pivot there, x turns NaN
    # and — exactly as in the reference, whose tolerance test lets NaN through (analysis.rs:331-345; pinned for dcop by
    # test_singular_matrix_status) — the step counts as converged: such an instance comes back with NaN waveforms, which is
    # how a caller tells. It is counted as failed here.
    gpu_failed = (status != 0) | np.any(~np.isfinite(wave), axis=(1, 2))
    same = gpu_failed == (o["status"] != 0)
    assert int(np.sum(same)) >= 14, (status, gpu_failed, o["status"])
    ok = ~gpu_failed & (o["status"] == 0)
    assert 4 <= int(np.sum(ok)) < B
    assert np.max(np.abs(wave[ok] - o["x"][ok][:, :, save])) <= 1e-7
    # ~190 Newton iterations per instance. Inside the device-resident time loop nobody re-pivots (the frozen order of the first
    # transient iteration is used throughout, weak pivots included), so the Newton path of a time point differs from the
    # reference's in its inexact steps: same waveforms, iteration counts within 15 %
    assert np.all(np.abs(iters[ok] - o["iters"][ok]) <= 0.15 * o["iters"][ok]), (iters[ok], o["iters"][ok])
    only_gpu = ~gpu_failed & (o["status"] != 0)   # converged here, capped in the oracle: the result must still be a solution
    assert np.all(np.abs(wave[only_gpu][:, :, -1] - ovr["V:vsup:dc"][only_gpu, None]) < 1e-9)
    # The plain-card 41-stage ring (N = 47) converges at every supply in the reference, and here. (Until the OP repair kept
    # instances that are singular under ANOTHER instance's pivot order open — resolve_repivot, host/batch.hpp — 7 of these 16
    # supplies ended their operating point with Singular Matrix after 2 iterations inside the batch, although every one of
    # them converges alone; the earlier reading "the frozen order fails inside the time loop" was wrong.)
    ck2, ic2 = cc.bsim4_ring(41, ic_every=20)
    c2 = ck2.to_s21().elaborate(ic=ic2)
    b2 = s21.Batch(c2, B)
    b2.override("V:vsup:dc", ovr["V:vsup:dc"])
    t2, w2, st2, it2 = b2.tran(tstep, npts * tstep)
    o2 = oracle.Circuit(ck2.to_text()).batch(1, B, overrides=ovr, tstep=tstep, tstop=npts * tstep, ic=ic2, nthreads=8)
    assert np.all(o2["status"] == 0) and np.all(st2 == 0), (st2, o2["status"])
    assert np.max(np.abs(w2 - o2["x"])) <= 1e-8
    assert np.all(np.abs(it2 - o2["iters"]) <= 0.1 * o2["iters"]), (it2, o2["iters"])


def test_team_kernel_committed_state_in_hbm(s21, monkeypatch):
    """Team kernel, transients of batches with more CTAs than SMs (host/jit_team.hpp SOPG): the committed device state stays in
    its HBM column (read through L2) instead of shared memory, which takes a CTA of the C1-circuit sweep from 139 KB to
    114 KB — two CTAs per SM, one wave for 8192 instances instead of two. Same arithmetic: forced on a small ragged batch
    (S21_TEAM_SOPG=1), waveforms, iteration counts and the committed state a second transient starts from must equal the
    shared-memory version's bit for bit."""
    ck = cc.cmos_ro3(cc.add_mos1_defaults)
    sup = np.linspace(0.9, 1.1, 70)

    def run():
        b = s21.Batch(ck.to_s21().elaborate(ic={"1": 0.0}), 70)
        b.override("V:v1:dc", sup)
        t, w, st, it = b.tran(1e-11, 60e-11)
        x2, st2, it2 = b.dcop()  # warm dcop from the transient's end state: reads x and the device state the kernel left in HBM
        return w, st, it, x2, b.kernel_name()
    monkeypatch.setenv("S21_KERNEL", "jitteam")
    monkeypatch.setenv("S21_TEAM_SOPG", "0")
    w0, st0, it0, x0, kn = run()
    monkeypatch.setenv("S21_TEAM_SOPG", "1")
    w1, st1, it1, x1, _ = run()
    assert kn == "jit-team" and np.all(st0 == 0) and np.all(st1 == 0)
    assert np.array_equal(w0, w1) and np.array_equal(it0, it1) and np.array_equal(x0, x1)


def test_transient_hand_back_and_resume(s21, monkeypatch):
    """Re-pivoting inside a transient (cooperative kernel; SolveCtl::tran_stop, Batch::resolve_tran_stops): an instance whose
    frozen pivot order meets an exactly zero pivot — or a non-finite step — inside the time loop goes back to its last
    accepted time point, comes back to the host with the time point recorded, gets a pivot order taken at that iterate and
    is continued by a resume launch from its own time point on. No circuit of the suite drives a pivot to zero inside its
    time loop, so the path is exercised by injection (S21_TRAN_INJECT=7: every instance is handed back once, in the second
    Newton iteration of time point 7 — x and the in-flight device state already moved): the continued waveforms must equal
    the uninterrupted run's (same matrix, same first-iteration order: bit for bit), and the hand-backs must show in the counters."""
    ck, ic = cc.bsim4_ring(21)
    ovr = {k: v[::85][:24] for k, v in cc.c4_sweep(2048).items()}

    def run():
        b = s21.Batch(ck.to_s21().elaborate(ic=ic), 24)
        for k, v in ovr.items():
            b.override(k, v)
        t, w, st, it = b.tran(1e-10, 40e-10)
        return w, st, it, b.setup_stats(), b.kernel_name()
    w0, st0, it0, ss0, kn = run()
    monkeypatch.setenv("S21_TRAN_INJECT", "7")
    w1, st1, it1, ss1, _ = run()
    assert kn == "coop" and np.all(st0 == 0) and np.all(st1 == 0)
    assert ss1["repaired_instances"] == ss0["repaired_instances"] + 24
    assert np.array_equal(w0, w1)
    assert np.all(it1 >= it0) and np.all(it1 - it0 <= 3)  # the abandoned attempt's iterations are counted


def test_grid_kernel_single_large_circuit(s21, oracle, monkeypatch):
    """Config C3's shape at a size the oracle can follow: one circuit, N > 96 -> the grid-wide kernel (all SMs on one
    instance, grid barriers between dependency levels). With S21_PLAN_EXACT=1 (every value sees its updates in the
    reference's order) it is bit-identical to the single-CTA cooperative kernel; the default tolerance-mode schedule
    (same-target updates applied atomically, host/symbolic.hpp build_levels) stays within 1e-9 of it, with far fewer
    levels; both agree with the oracle."""
    ck, ic = cc.inverter_array(20, 5)
    c = ck.to_s21().elaborate(ic=ic)
    assert c.n_vars > 96
    t, w, st, it = s21.Batch(c, 1).tran(1e-11, 1e-10)                       # default: tolerance mode
    monkeypatch.setenv("S21_PLAN_EXACT", "1")
    bx = s21.Batch(ck.to_s21().elaborate(ic=ic), 1)
    tx, wx, stx, itx = bx.tran(1e-11, 1e-10)
    assert bx.kernel_name() == "grid"
    monkeypatch.setenv("S21_KERNEL", "coop")
    t2, w2, st2, it2 = s21.Batch(ck.to_s21().elaborate(ic=ic), 1).tran(1e-11, 1e-10)
    monkeypatch.delenv("S21_KERNEL")
    monkeypatch.delenv("S21_PLAN_EXACT")
    assert st[0] == 0 and st2[0] == 0 and stx[0] == 0
    assert np.array_equal(wx, w2) and np.array_equal(itx, it2)
    # tolerance mode reorders the sums of same-target updates: values move in their last bits, and because the reference's
    # residual test (|res| <= 1e-12 A) sits at the round-off level of these currents, an occasional Newton iteration more
    # or less follows; the waveforms stay far inside SPICE's vntol
    diff = float(np.max(np.abs(w - wx)))
    print(f"grid kernel, tolerance vs exact level schedule: max |dv| = {diff:.3e}, Newton iterations {int(it[0])} vs {int(itx[0])}")
    # (tolerance mode also never stops for a weak pivot — an inexact step instead of a host round trip — which on this ring
    # array takes FEWER iterations than the exact mode's re-ordered ones: 42 against 59)
    assert diff <= 1e-7 and abs(int(it[0]) - int(itx[0])) <= 0.4 * int(itx[0])
    o = oracle.Circuit(ck.to_text()).tran(1e-11, 1e-10, ic=ic)
    # the tolerance mode is not bit-reproducible (atomic sums): every run stops its Newton iterations on the reference's
    # criterion (|dx| < 1e-3, |res| < 1e-12 A) along a slightly different path — measured 3e-10 ... 2.5e-8 against the oracle
    # over this round's runs; the bound sits 10x inside SPICE's vntol (1e-6)
    assert np.max(np.abs(w[0] - o.data)) <= 1e-7 and np.max(np.abs(wx[0] - o.data)) <= 1e-7
    x, sd, _ = s21.Batch(cc.inverter_array(20, 5)[0].to_s21().elaborate(), 1).dcop()
    assert sd[0] in (0, 1)  # without the IC the 100-iteration cap may hit, as in the reference; the launch must not hang


# ------------------------------------------------------------------------------------------------ hygiene
@pytest.mark.parametrize("tool,parts", [("memcheck", "all"), ("racecheck", "team"), ("racecheck", "bsim4"), ("synccheck", "team")])
def test_compute_sanitizer_clean(s21, tool, parts):
    """tests/sanitize_target.py — one small pass through every kernel family (specialised team / thread kernels with a
    ragged batch, hybrid, cooperative with shared-memory and HBM workspace, direct dcop / tran / adaptive / AC, grid-wide,
    probe, pack) — under compute-sanitizer: memcheck over everything, racecheck and synccheck over the kernels that
    communicate through shared memory, warp shuffles with lane masks and the deferred-exception redo paths. Skipped when
    the tool is not installed on the box."""
    import shutil
    import subprocess
    import sys
    exe = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(exe):
        pytest.skip("compute-sanitizer not installed")
    target = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sanitize_target.py")
    env = dict(os.environ)
    env.pop("S21_KERNEL", None)
    p = subprocess.run([exe, "--tool", tool, "--error-exitcode", "86", "--print-limit", "5", sys.executable, target, parts],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500, env=env)
    tail = "\n".join(p.stdout.splitlines()[-40:])
    assert "SANITIZE_TARGET_OK" in p.stdout, tail
    assert p.returncode == 0, tail
    assert "ERROR SUMMARY: 0 errors" in p.stdout or "RACECHECK SUMMARY: 0 hazards" in p.stdout, tail
