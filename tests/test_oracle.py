"""The oracle (oracle/, CPU restatement of the reference) against everything the reference's own tests pin for the
hot path: sparse21 unit vectors, dcop known answers, and the golden transient waveforms."""
import numpy as np
import pytest

import circuits as cc
from circuits import GND, Ckt
from conftest import golden


def test_sparse21_unit_vectors(oracle):
    # spice21/src/sparse21/mod.rs:1165-1526 restated in oracle_capi.cpp::orc_sparse21_selftest
    rc, msg = oracle.sparse21_selftest()
    assert rc == 0, msg


@pytest.mark.parametrize("builder,defaults,fixture,tstep,tstop", [
    (cc.cmos_ro3, cc.add_mos1_defaults, "test_mos1_cmos_ro_tran", 1e-11, 1e-8),   # tests.rs:932-947  (BASELINE config 1)
    (cc.cmos_ro3, cc.add_mos0_defaults, "test_mos0_cmos_ro_tran", 1e-15, 1e-12),  # tests.rs:773-790
    (cc.nmos_ro3, cc.add_mos1_defaults, "test_mos1_nmos_ro_tran", 1e-11, 1e-8),   # tests.rs:993-1008
    (cc.pmos_ro3, cc.add_mos1_defaults, "test_mos1_pmos_ro_tran", 1e-11, 1e-8),   # tests.rs:1036-1051
    (cc.cmos_ro3, cc.add_bsim4_defaults, "test_bsim4_cmos_ro_tran", 1e-10, 3e-7),  # tests.rs:948-964
    (cc.nmos_ro3, cc.add_bsim4_defaults, "test_bsim4_nmos_ro_tran", 1e-9, 1e-6),   # tests.rs:1360-1376
    (cc.pmos_ro3, cc.add_bsim4_defaults, "test_bsim4_pmos_ro_tran", 1e-11, 1e-8),  # tests.rs:1053-1069
])
def test_golden_waveforms(oracle, builder, defaults, fixture, tstep, tstop):
    g = golden(fixture)
    r = oracle.Circuit(builder(defaults).to_text()).tran(tstep, tstop, ic={"1": 0.0})
    assert r.data.shape[0] == len(g["time"])  # the float-accumulated time loop decides the point count
    assert np.array_equal(r.axis, g["time"])
    assert set(r.names) | {"time"} == set(g.keys())
    for k, name in enumerate(r.names):
        # reference tolerance is 1e-6 abs (tests.rs:788); the restatement is in fact bit-identical on glibc
        assert np.max(np.abs(r.data[:, k] - g[name])) <= 1e-6, name


def _dcop(oracle, ckt, **kw):
    r = oracle.Circuit(ckt.to_text()).dcop(**kw)
    return dict(zip(r.names, r.data[0])), r


def test_bsim4_nmos_dcop1(oracle):  # bsim4/tests.rs:57-87
    c = cc.add_bsim4_defaults(Ckt()).M("bsim4", "default", "default", d="gd", g="gd", s=GND, b=GND)
    c.V("v1", "gd", GND, 1.0).R("r1", "gd", GND, 1e-10)
    v, _ = _dcop(oracle, c)
    assert v["gd"] == 1.0
    assert abs(abs(v["v1"]) - 150e-6) < 1e-6


def test_bsim4_pmos_dcop1(oracle):  # bsim4/tests.rs:88-117
    c = cc.add_bsim4_defaults(Ckt()).M("bsim4", "pmos", "default", d="gd", g="gd", s=GND, b=GND)
    c.V("v1", "gd", GND, -1.0).R("r1", "gd", GND, 1e-10)
    v, _ = _dcop(oracle, c)
    assert v["gd"] == -1.0
    assert abs(abs(v["v1"]) - 57e-6) < 1e-6


def test_bsim4_inv_dcop(oracle):  # bsim4/tests.rs:118-160
    c = cc.add_bsim4_defaults(Ckt())
    c.M("p", "pmos", "default", d="d", g="inp", s="vdd", b="vdd").M("n", "nmos", "default", d="d", g="inp", s=GND, b=GND)
    c.V("vinp", "inp", GND, 0.0).V("vvdd", "vdd", GND, 1.0)
    v, _ = _dcop(oracle, c)
    assert v["vdd"] == 1.0 and v["inp"] == 0.0 and v["d"] > 0.95
    assert abs(v["vinp"]) < 1e-6 and abs(v["vvdd"]) < 1e-6


def test_dcop_r_only(oracle):  # tests.rs:22-27
    v, _ = _dcop(oracle, Ckt().R("r1", "a", GND, 1e-3))
    assert v == {"a": 0.0}


def test_dcop_i_r(oracle):  # tests.rs:30-44
    v, _ = _dcop(oracle, Ckt(signals=["vdd"]).I("i1", "vdd", GND, 1e-3).R("r1", "vdd", GND, 1e-3))
    assert v["vdd"] == 1.0


def test_dcop_v_r_r_exact(oracle):  # tests.rs:66-83 — exact equality in the reference
    c = Ckt(signals=["vdd", "div"]).V("v1", "vdd", GND, 1.0).R("r1", "vdd", "div", 2e-3).R("r2", GND, "div", 2e-3)
    v, _ = _dcop(oracle, c)
    assert v["vdd"] == 1.0 and v["div"] == 0.5 and v["v1"] == -1e-3


def test_dcop_diode_bias_consistency(oracle):  # tests.rs:87-133
    c = cc.add_diode_defaults(Ckt(signals=["p"])).D("dd", "p", GND, "default", "default").V("vin", "p", GND, 0.70)
    v, _ = _dcop(oracle, c)
    i = abs(v["vin"])
    assert 1e-3 < i < 100e-3
    c2 = cc.add_diode_defaults(Ckt(signals=["p"])).D("dd", "p", GND, "default", "default").I("i1", "p", GND, i)
    v2, _ = _dcop(oracle, c2)
    assert abs(v2["p"] - 0.70) < 1e-3


@pytest.mark.parametrize("model,vdc,isign", [("nmos", 1.0, -1), ("pmos", -1.0, +1)])
def test_dcop_mos0_char(oracle, model, vdc, isign):  # tests.rs:137-180
    c = cc.add_mos0_defaults(Ckt(signals=["g", "d"])).M("m", model, "default", d="d", g="g", s=GND, b=GND)
    c.V("v1", "g", GND, vdc).V("v2", "d", GND, vdc)
    v, _ = _dcop(oracle, c)
    assert v["g"] == vdc and v["d"] == vdc and v["v1"] == 0.0
    assert abs(v["v2"] - isign * 14.1e-3) < 1e-4


@pytest.mark.parametrize("model,idc,d,s,expect", [
    ("nmos", 5e-3, "0", GND, 0.697), ("nmos", 5e-3, GND, "0", 0.697),      # tests.rs:183-204, 237-258
    ("pmos", -5e-3, "0", GND, -0.697), ("pmos", -5e-3, GND, "0", -0.697),  # tests.rs:261-281, 315-335
])
def test_dcop_mos0_diode_connected(oracle, model, idc, d, s, expect):
    c = cc.add_mos0_defaults(Ckt()).I("i1", "0", GND, idc).M("m", model, "", d=d, g="0", s=s, b=GND)
    v, _ = _dcop(oracle, c)
    assert abs(v["0"] - expect) < 1e-3


def test_dcop_mos0_series_inverters(oracle):  # tests.rs:558-666
    c = cc.add_mos0_defaults(Ckt())
    for k in range(5):
        c.R("r1", str(k), GND, 1e-9)
    for k in range(4):
        c.M(f"p{k + 1}", "pmos", "", d=str(k + 1), g=str(k), s="0", b="0")
        c.M(f"n{k + 1}", "nmos", "", d=str(k + 1), g=str(k), s=GND, b=GND)
    c.V("v1", "0", GND, 1.0)
    _, r = _dcop(oracle, c)
    x = r.data[0]
    assert x[0] == 1.0 and abs(x[1]) < 1e-3 and abs(x[2] - 1.0) < 1e-3 and abs(x[3]) < 1e-3 and abs(x[4] - 1.0) < 1e-3 and abs(x[5]) < 1e-6


def test_dcop_rc_exact(oracle):  # tests.rs:670-692
    _, r = _dcop(oracle, Ckt().R("r1", "1", "0", 1e-3).C("c1", "1", GND, 1e-9).V("v1", "0", GND, 1.0))
    assert r.data[0].tolist() == [1.0, 1.0, 0.0]
    _, r = _dcop(oracle, Ckt().C("c1", "i", "o", 1e-9).R("r1", "o", GND, 1e-3).V("v1", "i", GND, 1.0))
    assert r.data[0].tolist() == [1.0, 0.0, 0.0]


def test_tran_rc_step(oracle):  # tests.rs:695-717
    c = Ckt().V("v1", "inp", GND, 1.0).R("r1", "inp", "out", 1e-3).C("c1", "out", GND, 1e-9)
    r = oracle.Circuit(c.to_text()).tran(10e-9, 10e-6, ic={"out": 0.0})
    inp, out = r.get("inp"), r.get("out")
    assert np.all(inp == 1.0)
    assert abs(out[0]) < 1e-3 and abs(out[-1] - 1.0) < 1e-3 and np.all(np.diff(out) > 0)


def test_mos1_op_and_inverter(oracle):  # tests.rs:792-815, 851-866, 915-929
    c = cc.add_mos1_defaults(Ckt()).M("m", "default", "default", d="0", g="0", s=GND, b=GND).V("v1", "0", GND, 1.0)
    _, r = _dcop(oracle, c)
    assert r.data[0][0] == 1.0 and -1e-3 < r.data[0][1] < 0.0
    v, _ = _dcop(oracle, cc.cmos_inv(cc.add_mos1_defaults))
    assert v["vdd"] == 1.0 and v["vss"] == 0.0 and v["inp"] == 0.0 and abs(v["out"] - 1.0) < 1e-3
    assert all(abs(v[k]) < 1e-6 for k in ("v1", "v2", "v3"))
    v, _ = _dcop(oracle, cc.cmos_ro3(cc.add_mos1_defaults))
    assert v["vdd"] == 1.0 and all(0.45 < v[k] < 0.55 for k in "123")


def test_hier_elaboration_counts(oracle):  # tests.rs:1380-1409: 11 comps / 5 vars after flattening
    c = Ckt()
    m = c.module("good_luck", ["inp", "out", "vss"])
    m.R("r1", "inp", "out", 0.001).C("r2", "out", "vss", 0.001)
    c.R("r0", "inp", GND, 0.001).R("rt", "out", "out2", 0.001).C("ct", "out2", GND, 0.001).C("ct", "out3", GND, 0.001).C("ct", "out4", GND, 0.001)
    c.X("x1", "good_luck", inp="inp", out="out", vss=GND).X("x2", "good_luck", inp="out2", out="out3", vss=GND)
    c.X("x3", "good_luck", inp="out3", out="out4", vss=GND)
    st = oracle.Circuit(c.to_text()).structure()
    assert st["n_vars"] == 5 and len(st["comp_kinds"]) == 11


def test_ac_rc_lowpass(oracle):  # spice21py/tests/test_spice21.py:53-70 + analytic check (AC values are unpinned by the reference)
    c = Ckt().R("r1", "inp", "out", 1e-3).C("c1", "out", GND, 1e-9).V("vi", "inp", GND, 1e-3, acm=1.0)
    r = oracle.Circuit(c.to_text()).ac(fstart=1, fstop=10**9, npts=90)
    assert r.data.shape[0] == 91  # end-inclusive sweep: npts + 1 points (analysis.rs:797-819)
    assert np.all(r.get("inp") == 1.0 + 0j)
    h = 1.0 / (1.0 + 2j * np.pi * r.axis * 1e-9 / 1e-3)
    assert np.max(np.abs(r.get("out") - h)) < 1e-9
    assert np.all(np.diff(np.abs(r.get("out"))) <= 0)


def test_ac_unsupported_device(oracle):  # comps/mod.rs:86-88: Isrc has no load_ac
    c = Ckt().R("r1", "a", GND, 1e-3).I("i1", "a", GND, 1e-3)
    with pytest.raises(oracle.OracleError) as e:
        oracle.Circuit(c.to_text()).ac(fstart=1, fstop=10, npts=2)
    assert e.value.status == 6


# ------------------------------------------------------------------------------------------------ second referee (row a18)
def _cs_amp(pmos=False):
    """A common-source stage with a resistive load and a load capacitor, C2's Mos1 model (tox given: non-zero Meyer
    capacitances, overlaps, body effect through a source degeneration resistor)."""
    import mos1_referee as mr
    sign = -1.0 if pmos else 1.0
    ck = Ckt().define("mos1model", "m", 1 if pmos else 0, **dict(cc.C2_MODEL, vt0=sign * cc.C2_MODEL["vt0"])).define("mos1inst", "wl", **cc.C2_INST)
    ck.V("vd", "vdd", GND, sign * 1.8).V("vg", "g", GND, sign * 0.9, acm=1.0)
    ck.R("rl", "vdd", "d", 5e-5).R("rs", "s", GND, 2e-4).C("cl", "d", GND, 1e-13)
    ck.M("m1", "m", "wl", d="d", g="g", s="s", b=GND)
    model = mr.resolve_model(1 if pmos else 0, **dict(cc.C2_MODEL, vt0=sign * cc.C2_MODEL["vt0"]))
    ip = mr.derive(model, mr.resolve_inst(**cc.C2_INST))
    dense = mr.Dense([("V", "vd", "vdd", "", sign * 1.8, 0.0), ("V", "vg", "g", "", sign * 0.9, 1.0), ("R", "vdd", "d", 5e-5),
                      ("R", "s", "", 2e-4), ("C", "d", "", 1e-13), ("M", model, ip, "d", "g", "s", "")])
    return ck, dense


@pytest.mark.parametrize("pmos", [False, True], ids=["nmos", "pmos"])
def test_mos1_second_referee_agrees_with_oracle(oracle, pmos):
    """tests/mos1_referee.py (numpy, written from mos.rs independently of oracle/) against the C++ restatement: the DC
    operating point, its iteration count, and the AC response of Mos1::load_ac incl. its duplicated (G,dr) stamp."""
    ck, dense = _cs_amp(pmos)
    x, iters, ops = dense.dcop()
    o = oracle.Circuit(ck.to_text()).dcop()
    om = dict(zip(o.names, o.data[0]))
    assert abs(ops[0]["ids"]) > 1e-5 and ops[0]["gmbs"] > 0.0  # the stage is biased on, with body effect (source degeneration)
    for n in dense.names:
        assert abs(x[dense.ix[n]] - om[n]) <= 1e-9 * max(1.0, abs(om[n])), n
    assert iters == o.solves
    freqs = [1e3, 1e6, 1e8, 1e9, 1e10]
    xa = dense.ac(ops, freqs)
    for k, f in enumerate(freqs):
        oa = oracle.Circuit(ck.to_text()).ac(fstart=int(f), fstop=int(f), npts=1)
        for n in dense.names:
            ref = oa.get(n)[0]
            assert abs(xa[k, dense.ix[n]] - ref) <= 1e-9 * max(1.0, abs(ref)), (f, n)
    if not pmos:
        assert abs(xa[0, dense.ix["d"]]) > 1.0 and abs(xa[-1, dense.ix["d"]]) < abs(xa[0, dense.ix["d"]])  # gain, then roll-off


# ------------------------------------------------------------------------------------------------ Bsim4 invariants
BSIM4_TOPOLOGY_VARIANTS = [
    {}, {"rgatemod": 1}, {"rgatemod": 2}, {"rgatemod": 3}, {"rbodymod": 1}, {"rdsmod": 1}, {"trnqsmod": 1},
    {"rgatemod": 3, "rbodymod": 1, "rdsmod": 1, "trnqsmod": 1}, {"rgatemod": 1, "rbodymod": 1, "igcmod": 1, "igbmod": 1},
    {"capmod": 0}, {"mobmod": 1}, {"mobmod": 4}, {"igcmod": 1, "igbmod": 1}, {"gidlmod": 1}, {"diomod": 0}, {"diomod": 2},
]


def test_bsim4_terminal_current_invariants(oracle):
    """The Bsim4 equations exist once in this repository (oracle/bsim4.hpp compiles the product's headers), so the reference
    pins them only through its goldens and known answers, all on the default card. These checks need no transcription to
    compare against: for one device driven by four voltage sources (NMOS and PMOS, forward and swapped bias)
    (1) the four source currents sum to zero up to the gmin leakage — the stamp conserves charge on every topology the
        selectors create (rgatemod 1-3, rbodymod, rdsmod, trnqsmod, and the PTM-65 combination rgatemod = rbodymod = igcmod = 1);
    (2) a gate resistance carries no DC current (currents equal to rgatemod = 0 to round-off), a body network only junction
        leakage (1e-8 relative), series resistances of this size move the drain current by < 1e-5 relative;
    (3) with a 1.2 nm oxide, igcmod / igbmod produce a gate current of the right sign, and conservation still holds.
    The same test runs on the GPU path (tests/test_gpu.py)."""
    def currents(sel, typ, bias):
        c = Ckt().define("bsim4model", "m", typ, **sel).define("bsim4inst", "i", l=1e-6, w=4e-6)
        c.M("m1", "m", "i", d="d", g="g", s="s", b="b")
        for n, v in zip("dgsb", bias):
            c.V("v" + n, n, GND, v)
        m = oracle.Circuit(c.to_text()).dcop().as_map()
        return np.array([float(np.ravel(m["v" + n])[0]) for n in "dgsb"])

    for typ, bias in ((0, (0.7, 0.9, 0.05, -0.1)), (1, (-0.7, -0.9, -0.05, 0.1)), (0, (0.05, 0.9, 0.7, -0.1))):
        base = currents({}, typ, bias)
        assert abs(base[0]) > 1e-5 and abs(base.sum()) < 1e-11
        for sel in BSIM4_TOPOLOGY_VARIANTS:
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
    assert -cur[1] > 1e-10  # current flows from the source vg INTO the gate: the branch current of vg is negative
