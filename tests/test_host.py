"""Host logic of the product, no GPU needed: the C-ABI library loads and exports every declared symbol, the protobuf
decoder and the builder produce the reference's variable numbering and stamp map, and the symbolic phase picks the
reference's pivot order and fill pattern integer for integer (checked against the oracle)."""
import os
import re
import sys

import numpy as np
import pytest

import circuits as cc
from circuits import GND, Ckt


def test_abi_symbols_match_header(s21):
    import os
    hdr = open(os.path.join(os.path.dirname(s21.__file__), "..", "include", "spice21cu.h")).read()
    declared = sorted(set(re.findall(r"\b(s21_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(s21.ABI_SYMBOLS)
    for sym in declared:
        assert hasattr(s21.lib(), sym), sym


def test_rust_ffi_declares_every_symbol(s21):
    """bindings/rust/spice21cu-sys (sources only; no cargo in this image) stays in step with the C header."""
    import os
    import re
    src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "bindings", "rust", "spice21cu-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"pub fn (s21_\w+)\(", src))
    assert declared == set(s21.ABI_SYMBOLS)


def test_no_cpu_fallback(s21):
    """Without a CUDA device the compute entry points must fail loudly (never fall back to a CPU path)."""
    if s21.cuda_device_count() > 0:
        pytest.skip("a GPU is present")
    c = Ckt(signals=["vdd"]).I("i1", "vdd", GND, 1e-3).R("r1", "vdd", GND, 1e-3).to_s21().elaborate()
    with pytest.raises(s21.Spice21Error) as e:
        s21.Batch(c, 4)
    assert e.value.status == s21.S21_CUDA_ERROR
    with pytest.raises(s21.Spice21Error) as e:
        s21.dcop(Ckt(signals=["vdd"]).I("i1", "vdd", GND, 1e-3).R("r1", "vdd", GND, 1e-3).to_proto())
    assert e.value.status == s21.S21_CUDA_ERROR


def test_sweep_partition_and_no_gpu(s21):
    """s21_sweep_partition (contiguous blocks of ceil(B / G)) covers [0, B) exactly once for every
    (B, G); without a CUDA device a sweep fails as loudly as a batch."""
    for B in (1, 3, 31, 1000, 2048, 8192, 100000):
        for G in (1, 2, 3, 4, 8):
            blocks = [s21.sweep_partition(B, G, g) for g in range(G)]
            per = -(-B // G)
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == B
            for g in range(G):
                assert blocks[g][0] == min(B, g * per) and blocks[g][1] <= per
                if g:
                    assert blocks[g][0] == blocks[g - 1][0] + blocks[g - 1][1]
    assert [s21.sweep_partition(8192, 8, g)[1] for g in range(8)] == [1024] * 8
    assert [s21.sweep_partition(2048, 8, g)[1] for g in range(8)] == [256] * 8
    with pytest.raises(s21.Spice21Error):
        s21.sweep_partition(10, 2, 2)
    if s21.cuda_device_count() == 0:
        c = Ckt(signals=["vdd"]).I("i1", "vdd", GND, 1e-3).R("r1", "vdd", GND, 1e-3).to_s21().elaborate()
        with pytest.raises(s21.Spice21Error) as e:
            s21.Sweep(c, 64, n_devices=2)
        assert e.value.status == s21.S21_CUDA_ERROR


def test_cubin_disk_cache(s21, tmp_path, monkeypatch):
    """The run-time specialised kernels' cubins are cached on disk keyed on the generated source (host/jit.hpp): a second
    compilation of the same source is a file read, a different source is a different file, a corrupt file is a miss."""
    import time
    monkeypatch.setenv("S21_CACHE_DIR", str(tmp_path))
    src = 'extern "C" __global__ void k_jit(double* x) { x[threadIdx.x] = x[threadIdx.x] * 3.0 + 1.0; }\n// %d\n' % os.getpid()
    t0 = time.time(); s21.jit_check(src); t1 = time.time(); s21.jit_check(src); t2 = time.time()
    files = sorted(os.listdir(tmp_path))
    assert len(files) == 1 and files[0].endswith(".cubin")
    blob = open(os.path.join(tmp_path, files[0]), "rb").read()
    assert src.encode() in blob and b"\x7fELF" in blob
    assert (t2 - t1) < (t1 - t0) or (t2 - t1) < 0.05
    s21.jit_check(src + "// other\n")
    assert len(os.listdir(tmp_path)) == 2
    open(os.path.join(tmp_path, files[0]), "wb").write(blob[: len(blob) // 2])  # truncated file: recompiled and replaced
    monkeypatch.setenv("S21_JIT_NO_MEMCACHE", "1")
    s21.jit_check(src)
    assert open(os.path.join(tmp_path, files[0]), "rb").read() == blob


def test_product_does_not_touch_oracle():
    import os
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "spice21_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".h", ".cu", ".cuh", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle/" not in txt and "pyoracle" not in txt and "liboracle" not in txt, os.path.join(dirpath, f)


CASES = [
    ("cmos_ro3_mos1", lambda: cc.cmos_ro3(cc.add_mos1_defaults), {"1": 0.0}, True),
    ("cmos_ro3_mos0", lambda: cc.cmos_ro3(cc.add_mos0_defaults), {"1": 0.0}, False),
    ("nmos_ro3", lambda: cc.nmos_ro3(cc.add_mos1_defaults), {"1": 0.0}, True),
    ("pmos_ro3", lambda: cc.pmos_ro3(cc.add_mos1_defaults), {"1": 0.0}, True),
    ("cmos_inv", lambda: cc.cmos_inv(cc.add_mos1_defaults), None, True),
    ("diffpair", cc.diffpair, None, True),
    ("diode", lambda: cc.add_diode_defaults(Ckt(signals=["p"])).D("dd", "p", GND, "default", "default").V("vin", "p", GND, 0.7), None, True),
    ("diode_rs", lambda: Ckt(signals=["p"]).define("diodemodel", "default", rs=10.0, bv=5.0, cj0=1e-12).define("diodeinst", "default", area=2.0)
        .D("dd", "p", GND, "default", "default").V("vin", "p", GND, 0.7), None, True),
    ("mos1_rd_rs", lambda: Ckt().define("mos1model", "n", 0, rd=10.0, rs=5.0, vt0=0.4).define("mos1inst", "i")
        .M("m", "n", "i", d="d", g="g", s=GND, b=GND).V("vg", "g", GND, 1.0).V("vd", "d", GND, 1.0), None, True),
    ("cmos_ro3_bsim4", lambda: cc.cmos_ro3(cc.add_bsim4_defaults), {"1": 0.0}, True),
    ("nmos_ro3_bsim4", lambda: cc.nmos_ro3(cc.add_bsim4_defaults), {"1": 0.0}, True),
] + [
    # every optional Bsim4 sub-network: internal nodes (bsim4ports.rs:24-112) and matrix pointers (bsim4solver.rs:28-115)
    ("bsim4_" + "_".join(f"{k}{v}" for k, v in sel.items()),
     (lambda sel=sel: Ckt().define("bsim4model", "n", 0, **sel).define("bsim4model", "p", 1, **sel).define("bsim4inst", "i", l=1e-6, w=2e-6)
      .M("mp", "p", "i", d="out", g="inp", s="vdd", b="vdd").M("mn", "n", "i", d="out", g="inp", s=GND, b=GND)
      .V("vi", "inp", GND, 0.4).V("vd", "vdd", GND, 1.0).R("rl", "out", GND, 1e-6)), None, False)
    for sel in ({"rgatemod": 1}, {"rgatemod": 2}, {"rgatemod": 3}, {"rbodymod": 1}, {"rdsmod": 1}, {"trnqsmod": 1},
                {"rgatemod": 3, "rbodymod": 1, "rdsmod": 1, "trnqsmod": 1})
]


@pytest.mark.parametrize("name,build,ic,protoable", CASES, ids=[c[0] for c in CASES])
def test_numbering_and_stamp_map(s21, oracle, name, build, ic, protoable):
    """Variable order (elab.rs:244-265) and element ids (Matrix::make in component order) are bit-exact."""
    ck = build()
    o = oracle.Circuit(ck.to_text()).structure(ic=ic)
    for via in ([False, True] if protoable else [False]):
        c = ck.to_s21(via_proto=via).elaborate(ic=ic)
        sm = c.stamp_map()
        assert c.names == o["names"]
        assert np.array_equal(sm["elem_row"], o["elem_row"]) and np.array_equal(sm["elem_col"], o["elem_col"])
        assert np.array_equal(sm["dev_off"], o["comp_off"]) and np.array_equal(sm["dev_elems"], o["comp_matps"])


def _same_plan(a, b):
    if a["status"] != b["status"]:
        return False
    if a["status"] != 0:
        return True
    return (np.array_equal(a["row_i2e"], b["row_i2e"]) and np.array_equal(a["col_i2e"], b["col_i2e"]) and
            set(zip(a["lu_row"].tolist(), a["lu_col"].tolist(), a["lu_fill"].tolist())) ==
            set(zip(b["lu_row"].tolist(), b["lu_col"].tolist(), b["lu_fill"].tolist())))


@pytest.mark.parametrize("name,build,ic,protoable", CASES, ids=[c[0] for c in CASES])
def test_pivot_order_on_circuit_matrices(s21, oracle, name, build, ic, protoable):
    """Symbolic phase == the reference's first factorisation (Markowitz order + fill), on the first-iteration matrix."""
    o = oracle.Circuit(build().to_text()).structure(ic=ic)
    sy = s21.symbolic(o["n_vars"], o["elem_row"], o["elem_col"], o["a0"])
    if len(o["row_i2e"]) == 0:
        assert sy["status"] != 0
        return
    assert sy["status"] == 0
    assert np.array_equal(sy["row_i2e"], o["row_i2e"]) and np.array_equal(sy["col_i2e"], o["col_i2e"])
    assert set(zip(sy["lu_row"].tolist(), sy["lu_col"].tolist(), sy["lu_fill"].tolist())) == \
        set(zip(o["lu_row"].tolist(), o["lu_col"].tolist(), o["lu_fill"].tolist()))


def test_pivot_order_random_matrices(s21, oracle):
    """Random sparse matrices incl. missing diagonals, structural zeros, ties and complex values."""
    rng = np.random.default_rng(21)
    for trial in range(400):
        n = int(rng.integers(2, 40))
        mask = rng.random((n, n)) < rng.uniform(0.05, 0.5)
        if rng.random() < 0.7:
            mask |= np.eye(n, dtype=bool)
        if rng.random() < 0.3:
            mask[int(rng.integers(0, n)), int(rng.integers(0, n))] = False
        mask[n - 1, int(rng.integers(0, n))] = True  # the reference sizes the matrix by its elements
        mask[int(rng.integers(0, n)), n - 1] = True
        r, c = np.nonzero(mask)
        perm = rng.permutation(len(r))
        r, c = r[perm], c[perm]
        cplx = rng.random() < 0.3
        v = rng.standard_normal(len(r)) * 10.0 ** rng.integers(-6, 3, len(r))
        if cplx:
            v = v + 1j * rng.standard_normal(len(r))
        if rng.random() < 0.3:
            v = np.round(v.real) + (1j * np.round(v.imag) if cplx else 0)
        assert _same_plan(oracle.lu_order(n, r, c, v), s21.symbolic(n, r, c, v)), trial


def test_pivot_order_tie_heavy_matrices(s21, oracle):
    """Larger matrices whose values come from a handful of magnitudes: many exact ties in the column maxima and in the
    Markowitz products, columns several cache blocks long (host/symbolic.hpp keeps column maxima per block of 16 list
    positions), dense rows/columns that every elimination step touches, like the supply node of config C3."""
    rng = np.random.default_rng(2026)
    for trial in range(60):
        n = int(rng.integers(40, 140))
        mask = rng.random((n, n)) < rng.uniform(0.02, 0.12)
        mask |= np.eye(n, dtype=bool)
        for _ in range(int(rng.integers(0, 3))):  # a few (nearly) full rows / columns
            k = int(rng.integers(0, n))
            mask[k, :] |= rng.random(n) < 0.9
            mask[:, k] |= rng.random(n) < 0.9
        r, c = np.nonzero(mask)
        perm = rng.permutation(len(r))
        r, c = r[perm], c[perm]
        v = rng.choice([1.0, -1.0, 2.0, -2.0, 0.5, 1e-4], size=len(r))
        v[r == c] *= rng.choice([1.0, 4.0, 1e-3], size=int(np.sum(r == c)))
        if trial % 5 == 4:
            v = v + 1j * rng.choice([0.0, 1.0, -1.0], size=len(r))
        assert _same_plan(oracle.lu_order(n, r, c, v), s21.symbolic(n, r, c, v)), trial


def test_bsim4_divisions_stay_routed():
    """Every division of the BSIM4 evaluation goes through B4_DIV (bsim4/bsim4_eval.hpp) so that the device build can
    route them through csrc/scalar.h with one definition. The rewriting tool must keep C++'s grouping (self-test on random
    expressions: identical values), and a dry run over the headers must find nothing left to rewrite."""
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import b4_route_divisions as tool
    tool.selftest(trials=400, seed=11)
    d = os.path.join(ROOT, "spice21_b200", "csrc", "bsim4")
    for f in tool.FILES:
        out, n = tool.rewrite(open(os.path.join(d, f)).read())
        assert n == 0, f"{f}: {n} plain divisions; run scripts/b4_route_divisions.py --apply"
    assert tool.rewrite("x = -a * b / c / (d + 1.0) * e;")[0] == "x = B4_DIV(B4_DIV(-a * b, c), (d + 1.0)) * e;"
    assert tool.rewrite("y = *p * (1.0 + t) / q;")[0] == "y = B4_DIV(*p * (1.0 + t), q);"
    assert tool.rewrite("z = a + f(b / c, 2.0) / s.m[1];")[0] == "z = a + B4_DIV(f(B4_DIV(b, c), 2.0), s.m[1]);"


def test_proto_decode_errors(s21):
    with pytest.raises(s21.Spice21Error) as e:
        s21.Circuit(b"\x0a\xff\xff\xff\xff\x0f")  # truncated length-delimited field
    assert e.value.status == s21.S21_DECODE_ERROR
    with pytest.raises(s21.Spice21Error) as e:  # "No Circuit Provided" (proto.rs:62)
        s21._dcop(b"")
    assert "No Circuit Provided" in e.value.desc


def test_invalid_circuits(s21):
    c = cc.add_mos1_defaults(Ckt()).M("m", "nosuchmodel", "default", d="d", g="g", s=GND, b=GND).to_s21()
    with pytest.raises(s21.Spice21Error) as e:  # elab.rs:167 panics "Model not defined"
        c.elaborate()
    assert e.value.status == s21.S21_INVALID_CIRCUIT and "Model not defined" in e.value.desc
    c = Ckt()
    c.module("m", ["a"]).R("r", "a", "undeclared", 1.0)
    c.R("r0", "x", GND, 1.0).X("x1", "m", a="x")
    with pytest.raises(s21.Spice21Error) as e:  # elab.rs:49: nodes inside modules must pre-exist
        c.to_s21().elaborate()
    assert e.value.status == s21.S21_INVALID_CIRCUIT


def test_time_and_frequency_axes(s21, oracle):
    """Point counts are decided by the reference's floating-point accumulation (analysis.rs:553-570, 791-819)."""
    assert s21.lib().s21_tran_num_points(1e-11, 1e-8) == 1001
    assert s21.lib().s21_tran_num_points(1e-15, 1e-12) == 1000
    assert s21.lib().s21_tran_num_points(1e-10, 3e-7) == 3000
    assert s21.lib().s21_tran_num_points(0.0, 0.0) == 1
    f = s21.ac_freqs(1, 10**9, 90)
    c = Ckt().R("r1", "inp", "out", 1e-3).C("c1", "out", GND, 1e-9).V("vi", "inp", GND, 1e-3, acm=1.0)
    assert np.array_equal(f, oracle.Circuit(c.to_text()).ac(fstart=1, fstop=10**9, npts=90).axis)
    assert len(s21.ac_freqs(1, 10**10, 99999)) == 100000
    assert len(s21.ac_freqs(0, 0, 0)) == 1


@pytest.mark.parametrize("shape", [0, 1], ids=["thread", "team"])
def test_specialised_kernel_sources_compile(s21, oracle, shape):
    """The run-time specialised Newton kernels (host/jit.hpp, host/jit_team.hpp) are generated from the plan and compiled
    with NVRTC on first use. Generation and compilation need no GPU: check here that the sources generated for an OP and
    a transient plan compile for sm_100a, and that the team generator refuses what it is not written for."""
    for ck, ic, mode in ((cc.diffpair(), None, 0), (cc.cmos_ro3(cc.add_mos1_defaults), {"1": 0.0}, 1)):
        o = oracle.Circuit(ck.to_text()).structure(ic=ic)
        c = ck.to_s21().elaborate(ic=ic) if ic else ck.to_s21().elaborate()
        src, smem = c.jit_source(o["a0"], mode=mode, shape=shape)
        assert "k_jit" in src and 0 < smem <= 227 * 1024
        try:
            s21.jit_check(src)
        except s21.Spice21Error as e:
            if "libnvrtc" in str(e):
                pytest.skip("NVRTC is not installed here")
            raise
    if shape == 1:
        ck = cc.rc_opamp(8)  # N = 28 rows: beyond the register-resident linear algebra
        o = oracle.Circuit(ck.to_text()).structure()
        with pytest.raises(s21.Spice21Error):
            ck.to_s21().elaborate().jit_source(o["a0"], mode=0, shape=1)


def test_wire_extensions_decode_like_the_builder(s21):
    """The two wire-format extensions of SURVEY §8 f2 — Vsrc fields 6 / 7 (time-varying source) and Bsim4Model field 901 (model
    card parameters by name) — decode into the same elaborated circuit as the C-ABI builder calls: same variables, same stamp
    map (the card's rgatemod = 1 adds the internal gate node), and a message without them still decodes as before."""
    def build(wave, card):
        c = Ckt().define("bsim4model", "n", 0, **card).define("bsim4inst", "i", l=1e-6, w=4e-6)
        c.M("mn", "n", "i", d="d", g="g", s=GND, b=GND).V("vd", "d", GND, 1.0).V("vg", "g", GND, 0.8, wave=wave)
        return c

    full = build(("pulse", [0.0, 1.0, 1e-9, 1e-10, 1e-10, 1e-9, 4e-9]), {"mobmod": 1, "rgatemod": 1, "toxe": 2e-9, "toxp": 2e-9})
    a, b = full.to_s21().elaborate(), full.to_s21(via_proto=True).elaborate()
    assert a.names == b.names and a.n_vars == 5
    ma, mb = a.stamp_map(), b.stamp_map()
    assert all(np.array_equal(x, y) for x, y in zip(ma, mb))
    plain = build(None, {})
    p1, p2 = plain.to_s21().elaborate(), plain.to_s21(via_proto=True).elaborate()
    assert p1.names == p2.names and p1.n_vars == 4
    raw = full.to_proto().SerializeToString()
    assert b"rgatemod" in raw and len(raw) > len(plain.to_proto().SerializeToString())
