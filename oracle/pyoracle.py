"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes driver for ``oracle/liboracle.so`` (the CPU restatement of the reference solver). Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may import this module;
the product package ``spice21_b200`` never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile liboracle.so with the committed Makefile (g++ only; seconds)."""
    so = os.path.join(_HERE, "liboracle.so")
    if force and os.path.exists(so):
        os.remove(so)
    subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.environ.get("ORC_LIB") or os.path.join(_HERE, "liboracle.so")  # ORC_LIB: an instrumented build (oracle/Makefile)
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_ckt_parse.restype = C.c_void_p
        L.orc_ckt_parse.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.orc_ckt_free.argtypes = [C.c_void_p]
        for f in ("orc_res_nsig", "orc_res_npts", "orc_res_width", "orc_st_nvars"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = C.c_int
        for f in ("orc_res_names", "orc_st_names", "orc_st_comp_kinds"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = C.c_char_p
        for f in ("orc_res_data", "orc_res_axis", "orc_res_stats"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_void_p]
        L.orc_res_free.argtypes = [C.c_void_p]
        L.orc_st_free.argtypes = [C.c_void_p]
        L.orc_st_vec.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_st_vec.restype = C.c_int
        L.orc_sparse21_selftest.argtypes = [C.c_char_p, C.c_int]
        L.orc_sparse21_selftest.restype = C.c_int
        _LIB = L
    return _LIB


class OracleError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"[{status}] {msg}")
        self.status = status
        self.desc = msg


def _opts5(opts):
    o = np.full(5, np.nan)
    if opts:
        for k, name in enumerate(("temp", "tnom", "gmin", "iabstol", "reltol")):
            if name in opts and opts[name] is not None:
                o[k] = opts[name]
    return o


def _strs(names):
    arr = (C.c_char_p * max(1, len(names)))()
    for k, s in enumerate(names):
        arr[k] = (s if s != "" else "~").encode()
    return arr


class Result:
    def __init__(self, handle):
        L = lib()
        self.names = L.orc_res_names(handle).decode().split("\n") if L.orc_res_nsig(handle) else []
        nsig, npts, w = L.orc_res_nsig(handle), L.orc_res_npts(handle), L.orc_res_width(handle)
        data = np.zeros(nsig * npts * w)
        L.orc_res_data(handle, data.ctypes.data_as(C.c_void_p))
        self.data = data.reshape(npts, nsig) if w == 1 else data.reshape(npts, nsig, 2).view(np.complex128).reshape(npts, nsig)
        self.axis = np.zeros(npts)
        if npts and (w == 2 or npts > 1):
            L.orc_res_axis(handle, self.axis.ctypes.data_as(C.c_void_p))
        st = np.zeros(4)
        L.orc_res_stats(handle, st.ctypes.data_as(C.c_void_p))
        self.loads, self.solves, self.factorizations, self.seconds = int(st[0]), int(st[1]), int(st[2]), float(st[3])
        L.orc_res_free(handle)

    def get(self, name):
        return self.data[:, self.names.index(name)]

    def as_map(self):
        return {n: self.data[:, k] for k, n in enumerate(self.names)}


class Circuit:
    """A parsed oracle circuit (text netlist, see oracle_capi.cpp header)."""

    def __init__(self, text):
        err = C.create_string_buffer(512)
        self.h = lib().orc_ckt_parse(text.encode(), err, 512)
        if not self.h:
            raise OracleError(5, err.value.decode())
        self.text = text

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_ckt_free(self.h)
            self.h = None

    def _check(self, st, err):
        if st != 0:
            raise OracleError(st, err.value.decode())

    def dcop(self, opts=None):
        out, err = C.c_void_p(), C.create_string_buffer(512)
        o = _opts5(opts)
        lib().orc_run_op.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_char_p, C.c_int]
        self._check(lib().orc_run_op(self.h, o.ctypes.data_as(C.c_void_p), C.byref(out), err, 512), err)
        return Result(out)

    def tran(self, tstep, tstop, ic=None, opts=None, max_points=0):
        out, err = C.c_void_p(), C.create_string_buffer(512)
        o = _opts5(opts)
        ic = ic or {}
        nodes = _strs([str(k) for k in ic])
        vals = np.array([float(v) for v in ic.values()] + [0.0])
        f = lib().orc_run_tran
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_long,
                      C.POINTER(C.c_void_p), C.c_char_p, C.c_int]
        self._check(f(self.h, o.ctypes.data_as(C.c_void_p), tstep, tstop, len(ic), nodes, vals.ctypes.data_as(C.c_void_p),
                      max_points, C.byref(out), err, 512), err)
        return Result(out)

    def ac(self, fstart=0, fstop=0, npts=0, opts=None, max_points=0):
        out, err = C.c_void_p(), C.create_string_buffer(512)
        o = _opts5(opts)
        f = lib().orc_run_ac
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_ulonglong, C.c_ulonglong, C.c_long, C.POINTER(C.c_void_p), C.c_char_p, C.c_int]
        self._check(f(self.h, o.ctypes.data_as(C.c_void_p), int(fstart), int(fstop), int(npts), int(max_points), C.byref(out), err, 512), err)
        return Result(out)

    def ac_at(self, freqs, opts=None):
        """Oracle-only: cold-start AC solves at the given frequencies (what a one-point sweep computes at each of them)."""
        out, err = C.c_void_p(), C.create_string_buffer(512)
        o = _opts5(opts)
        fr = np.ascontiguousarray(freqs, dtype=np.float64)
        f = lib().orc_run_ac_at
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.POINTER(C.c_void_p), C.c_char_p, C.c_int]
        self._check(f(self.h, o.ctypes.data_as(C.c_void_p), fr.ctypes.data_as(C.c_void_p), len(fr), C.byref(out), err, 512), err)
        return Result(out)

    def structure(self, ic=None, opts=None):
        """Variable numbering, element creation order (stamp map) and first-factorisation pivot order / fill."""
        out, err = C.c_void_p(), C.create_string_buffer(512)
        o = _opts5(opts)
        ic = ic or {}
        nodes = _strs([str(k) for k in ic])
        vals = np.array([float(v) for v in ic.values()] + [0.0])
        f = lib().orc_structure
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_char_p, C.c_int]
        self._check(f(self.h, o.ctypes.data_as(C.c_void_p), len(ic), nodes, vals.ctypes.data_as(C.c_void_p), C.byref(out), err, 512), err)
        L = lib()
        st = {"n_vars": L.orc_st_nvars(out), "names": L.orc_st_names(out).decode().split("\n"),
              "comp_kinds": L.orc_st_comp_kinds(out).decode().split("\n")}
        for which, key in enumerate(("elem_row", "elem_col", "comp_off", "comp_matps", "row_i2e", "col_i2e", "lu_row", "lu_col", "lu_fill")):
            n = L.orc_st_vec(out, which, None, 0)
            v = np.zeros(max(n, 1), dtype=np.int32)
            L.orc_st_vec(out, which, v.ctypes.data_as(C.c_void_p), n)
            st[key] = v[:n]
        L.orc_st_a0.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_st_a0.restype = C.c_int
        n = L.orc_st_a0(out, None, 0)
        a0 = np.zeros(max(n, 1))
        L.orc_st_a0(out, a0.ctypes.data_as(C.c_void_p), n)
        st["a0"] = a0[:n]
        L.orc_st_free(out)
        return st

    def batch(self, kind, B, overrides=None, opts=None, tstep=0.0, tstop=0.0, ic=None, max_points=0, nthreads=1, want_x=True):
        """B independent instances with per-instance overrides {"kind:name:param": array[B]}.

        Returns dict(x, iters, status, seconds, n_vars, n_pts). ``seconds`` covers Solver::solve only
        (solvers are built beforehand, untimed)."""
        overrides = overrides or {}
        specs = (C.c_char_p * max(1, len(overrides)))()
        for k, s in enumerate(overrides):
            specs[k] = s.encode()
        vals = np.ascontiguousarray(np.array([np.asarray(v, dtype=np.float64) for v in overrides.values()]).reshape(len(overrides), B)) \
            if overrides else np.zeros((1, B))
        ic = ic or {}
        nodes = _strs([str(k) for k in ic])
        icv = np.array([float(v) for v in ic.values()] + [0.0])
        o = _opts5(opts)
        # discover sizes with a dry structure call
        st = self.structure(ic=ic if kind == 1 else None, opts=opts)
        N = st["n_vars"]
        if kind == 1:
            t, T = tstep, 1
            while t < tstop and (not max_points or T - 1 < max_points):
                T += 1
                t += tstep
        else:
            T = 1
        x = np.zeros((B, T, N)) if want_x else None
        iters = np.zeros(B, dtype=np.int64)
        status = np.zeros(B, dtype=np.int32)
        secs, nv, npt = C.c_double(), C.c_int(), C.c_int()
        err = C.create_string_buffer(512)
        f = lib().orc_batch_run
        f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int,
                      C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double),
                      C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.c_int]
        stt = f(self.h, kind, o.ctypes.data_as(C.c_void_p), B, len(overrides), specs, vals.ctypes.data_as(C.c_void_p), tstep, tstop,
                len(ic), nodes, icv.ctypes.data_as(C.c_void_p), max_points, nthreads,
                x.ctypes.data_as(C.c_void_p) if want_x else None, iters.ctypes.data_as(C.c_void_p),
                status.ctypes.data_as(C.c_void_p), C.byref(secs), C.byref(nv), C.byref(npt), err, 512)
        self._check(stt, err)
        return {"x": x[:, 0, :] if (want_x and kind == 0) else x, "iters": iters, "status": status, "seconds": secs.value,
                "n_vars": nv.value, "n_pts": npt.value, "names": st["names"]}


def lu_order(n, rows, cols, vals):
    """Pivot order + L+U pattern the restated sparse21 produces for an arbitrary matrix (real or complex vals)."""
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    vals = np.asarray(vals)
    width = 2 if np.iscomplexobj(vals) else 1
    v = np.ascontiguousarray(vals, dtype=np.complex128 if width == 2 else np.float64)
    cap = n * n + len(rows) + 8
    row_i2e, col_i2e = np.arange(n, dtype=np.int32), np.arange(n, dtype=np.int32)
    lr, lc, lf = (np.zeros(cap, dtype=np.int32) for _ in range(3))
    nnz = C.c_int()
    f = lib().orc_lu_order
    f.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.POINTER(C.c_int)]
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    st = f(n, len(rows), p(rows), p(cols), p(v), width, p(row_i2e), p(col_i2e), p(lr), p(lc), p(lf), cap, C.byref(nnz))
    k = nnz.value
    return {"status": st, "row_i2e": row_i2e, "col_i2e": col_i2e, "lu_row": lr[:k], "lu_col": lc[:k], "lu_fill": lf[:k]}


def sparse21_selftest():
    msg = C.create_string_buffer(256)
    rc = lib().orc_sparse21_selftest(msg, 256)
    return rc, msg.value.decode()
