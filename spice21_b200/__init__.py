"""spice21_b200 — Python host side over ``libspice21cu.so`` (C ABI in ``include/spice21cu.h``).

Two layers, both thin:

* the reference's own Python surface (``spice21py/spice21py/__init__.py``: ``dcop``, ``tran``, ``ac``, ``circuit``,
  ``add`` — protobuf messages in, dicts out), routed through ``s21_{op,tran,ac}_bytes``;
* the structured/batched surface (``Circuit`` builder → ``elaborate`` → ``Batch``) that a host Solver uses to run the
  Newton loop for thousands of instances on one GPU.

There is no CPU implementation behind any of this: if ``libspice21cu.so`` is missing the import fails, and if no CUDA
device is visible every solve raises ``Spice21Error`` with status ``S21_CUDA_ERROR``.
"""
import ctypes as C
import os

import numpy as np

from . import protos
from .protos import (Capacitor, Resistor, Diode, Mos, Isrc, Vsrc, MosType, TranOptions, Bsim4Model, Bsim4InstParams)  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("S21_LIB") or os.path.join(_HERE, "libspice21cu.so")  # S21_LIB: an experimental build (e.g. make B4_SDIV=1 OUT=...)

S21_OK, S21_CONVERGENCE_FAILED, S21_SINGULAR_MATRIX, S21_PIVOT_SEARCH_FAIL, S21_DECODE_ERROR, S21_INVALID_CIRCUIT, \
    S21_UNSUPPORTED, S21_CUDA_ERROR, S21_OTHER = range(9)

# every symbol include/spice21cu.h declares (tests check the .so exports all of them)
ABI_SYMBOLS = [
    "s21_last_error", "s21_free", "s21_cuda_device_count", "s21_op_bytes", "s21_tran_bytes", "s21_ac_bytes", "s21_ckt_from_proto",
    "s21_ckt_new", "s21_ckt_destroy", "s21_ckt_signal", "s21_ckt_add_r", "s21_ckt_add_c", "s21_ckt_add_i", "s21_ckt_add_v",
    "s21_ckt_add_d", "s21_ckt_add_v_wave", "s21_ckt_add_mos", "s21_ckt_add_x", "s21_ckt_def_module", "s21_ckt_define", "s21_ckt_elaborate",
    "s21_ckt_num_vars", "s21_ckt_var_name", "s21_ckt_var_kind", "s21_ckt_num_devices", "s21_ckt_stamp_map", "s21_batch_create",
    "s21_batch_destroy", "s21_batch_set_stream", "s21_batch_override", "s21_batch_sync_params", "s21_batch_reset", "s21_batch_dcop",
    "s21_batch_dcop_device", "s21_batch_read", "s21_tran_num_points", "s21_batch_tran", "s21_ac_freqs", "s21_batch_ac",
    "s21_batch_pivot_order", "s21_batch_dcop_view", "s21_batch_stats", "s21_batch_kernel_name", "s21_jit_source", "s21_jit_check", "s21_selftest_div", "s21_symbolic",
    "s21_batch_setup_stats", "s21_batch_plan_info", "s21_batch_tran_adaptive", "s21_batch_set_aids", "s21_batch_packed_device", "s21_batch_wave_device", "s21_batch_step_dcop_view", "s21_sweep_partition", "s21_sweep_create", "s21_sweep_destroy", "s21_sweep_num_devices", "s21_sweep_shard",
    "s21_sweep_override", "s21_sweep_sync_params", "s21_sweep_reset", "s21_sweep_dcop", "s21_sweep_dcop_view", "s21_sweep_tran", "s21_sweep_ac",
    "s21_sweep_stats",
]


class Spice21Error(RuntimeError):
    """SpError{desc} of the reference (spice21/src/spresult.rs:8-12) plus the S21_* status code."""

    def __init__(self, status, desc):
        super().__init__(f"[{status}] {desc}")
        self.status = status
        self.desc = desc


_lib = None


def lib():
    """The loaded C-ABI library. Raises if it has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build the CUDA extension first (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.s21_last_error.restype = C.c_char_p
        L.s21_ckt_var_name.restype = C.c_char_p
        L.s21_ckt_var_name.argtypes = [C.c_void_p, C.c_int32]
        L.s21_tran_num_points.restype = C.c_int64
        L.s21_tran_num_points.argtypes = [C.c_double, C.c_double]
        L.s21_ac_freqs.restype = C.c_int64
        L.s21_ac_freqs.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_size_t]
        L.s21_free.argtypes = [C.c_void_p]
        for f in ("s21_op_bytes", "s21_tran_bytes", "s21_ac_bytes"):
            getattr(L, f).argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.s21_ckt_from_proto.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.s21_ckt_new.argtypes = [C.POINTER(C.c_void_p)]
        L.s21_ckt_destroy.argtypes = [C.c_void_p]
        L.s21_ckt_signal.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        for f in ("s21_ckt_add_r", "s21_ckt_add_c", "s21_ckt_add_i"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_double]
        L.s21_ckt_add_v.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_double, C.c_double]
        L.s21_ckt_add_d.argtypes = [C.c_void_p] + [C.c_char_p] * 6
        L.s21_ckt_add_v_wave.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_double, C.c_double, C.c_int32, C.c_size_t, C.c_void_p]
        L.s21_ckt_add_mos.argtypes = [C.c_void_p] + [C.c_char_p] * 8
        L.s21_ckt_add_x.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.s21_ckt_def_module.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p]
        L.s21_ckt_define.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int32, C.c_size_t, C.c_void_p, C.c_void_p]
        L.s21_ckt_elaborate.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        for f in ("s21_ckt_num_vars", "s21_ckt_num_devices"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.s21_ckt_var_kind.argtypes = [C.c_void_p, C.c_int32]
        L.s21_ckt_stamp_map.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.s21_batch_create.argtypes = [C.c_void_p, C.c_int32, C.c_size_t, C.POINTER(C.c_void_p)]
        L.s21_batch_destroy.argtypes = [C.c_void_p]
        L.s21_batch_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.s21_batch_override.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.s21_batch_sync_params.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_size_t)]
        L.s21_batch_reset.argtypes = [C.c_void_p]
        L.s21_batch_dcop.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.s21_batch_dcop_device.argtypes = [C.c_void_p]
        L.s21_batch_dcop_view.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.s21_batch_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.s21_batch_step_dcop_view.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.s21_batch_tran.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.s21_batch_ac.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        L.s21_batch_pivot_order.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.s21_batch_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.s21_batch_kernel_name.argtypes = [C.c_void_p]
        L.s21_batch_kernel_name.restype = C.c_char_p
        L.s21_jit_source.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        L.s21_jit_check.argtypes = [C.c_char_p, C.c_size_t]
        L.s21_selftest_div.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.s21_batch_setup_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.s21_batch_plan_info.argtypes = [C.c_void_p, C.c_void_p]
        L.s21_batch_set_aids.argtypes = [C.c_void_p, C.c_int32]
        L.s21_batch_tran_adaptive.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_size_t] + [C.c_void_p] * 6
        L.s21_batch_packed_device.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.s21_batch_wave_device.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.s21_sweep_partition.argtypes = [C.c_size_t, C.c_int32, C.c_int32, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.s21_sweep_create.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.s21_sweep_destroy.argtypes = [C.c_void_p]
        L.s21_sweep_num_devices.argtypes = [C.c_void_p]
        L.s21_sweep_shard.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.s21_sweep_override.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.s21_sweep_sync_params.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_size_t)]
        L.s21_sweep_reset.argtypes = [C.c_void_p]
        L.s21_sweep_dcop.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.s21_sweep_dcop_view.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.s21_sweep_tran.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.s21_sweep_ac.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        L.s21_sweep_stats.argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _check(status):
    if status != S21_OK:
        raise Spice21Error(status, lib().s21_last_error().decode(errors="replace"))


def cuda_device_count():
    return lib().s21_cuda_device_count()


def health():
    """spice21py.health() (spice21py/src/lib.rs:28-31)."""
    lib()
    return "alive"


# ---------------------------------------------------------------------------------------------- bytes in / bytes out
def _call_bytes(fn, enc):
    out, n = C.c_void_p(), C.c_size_t()
    _check(getattr(lib(), fn)(enc, len(enc), C.byref(out), C.byref(n)))
    try:
        return C.string_at(out, n.value)
    finally:
        lib().s21_free(out)


def _dcop(enc):  # spice21py/src/lib.rs:50-53
    return _call_bytes("s21_op_bytes", enc)


def _tran(enc):  # spice21py/src/lib.rs:56-59
    return _call_bytes("s21_tran_bytes", enc)


def _ac(enc):  # spice21py/src/lib.rs:62-65
    return _call_bytes("s21_ac_bytes", enc)


def _wrap(arg, cls, ckt, **kw):
    if arg is None:
        return cls(ckt=ckt, **{k: v for k, v in kw.items() if v is not None})
    if isinstance(arg, protos.Circuit):
        return cls(ckt=arg, **{k: v for k, v in kw.items() if v is not None})
    if isinstance(arg, cls):
        return arg
    raise TypeError


def dcop(arg=None, *, ckt=None, opts=None):
    """DC operating point: {signal name: value} (spice21py/__init__.py:16-43)."""
    rv = protos.OpResult()
    rv.ParseFromString(_dcop(_wrap(arg, protos.Op, ckt, opts=opts).SerializeToString()))
    return dict(rv.vals)


def tran(arg=None, *, ckt=None, opts=None, args=None):
    """Transient: {signal name: [values]} incl. "time" (spice21py/__init__.py:46-74)."""
    rv = protos.TranResult()
    rv.ParseFromString(_tran(_wrap(arg, protos.Tran, ckt, opts=opts, args=args).SerializeToString()))
    return {name: list(arr.vals) for name, arr in dict(rv.vals).items()}


def ac(arg=None, *, ckt=None, opts=None, args=None):
    """AC: {signal name: [complex]} (spice21py/__init__.py:77-109)."""
    rv = protos.AcResult()
    rv.ParseFromString(_ac(_wrap(arg, protos.Ac, ckt, opts=opts, args=args).SerializeToString()))
    return {name: [complex(c.re, c.im) for c in arr.vals] for name, arr in dict(rv.vals).items()}


def add(ckt, comp):
    """spice21py.add (spice21py/__init__.py:112-146)."""
    from .protos import Instance
    if isinstance(comp, Instance):
        return ckt.comps.append(comp)
    for cls, key in ((Diode, "d"), (Capacitor, "c"), (Resistor, "r"), (Mos, "m"), (Isrc, "i"), (Vsrc, "v"), (protos.ModuleInstance, "x")):
        if isinstance(comp, cls):
            return ckt.comps.append(Instance(**{key: comp}))
    for cls, key in ((protos.Bsim4Model, "bsim4model"), (protos.Bsim4InstParams, "bsim4inst"), (protos.Mos1Model, "mos1model"),
                     (protos.Mos1InstParams, "mos1inst"), (protos.DiodeModel, "diodemodel"), (protos.DiodeInstParams, "diodeinst"),
                     (protos.Module, "module")):
        if isinstance(comp, cls):
            return ckt.defs.append(protos.Def(**{key: comp}))
    raise TypeError(f"Invalid Circuit Component {type(comp)}")


def circuit(*args):
    """spice21py.circuit (spice21py/__init__.py:149-161)."""
    _args = args[0] if len(args) == 1 and isinstance(args[0], (list, tuple)) else args
    c = protos.Circuit()
    for a in _args:
        add(c, a)
    return c


# ---------------------------------------------------------------------------------------------- structured surface
def _b(s):
    return (s or "").encode()


def _strarr(strs):
    arr = (C.c_char_p * max(1, len(strs)))()
    for k, s in enumerate(strs):
        arr[k] = _b(s)
    return arr


class Circuit:
    """``s21_ckt``: Ckt::new / Ckt::from_proto + builder calls, then ``elaborate`` (Solver::new + Tran::ic)."""

    def __init__(self, proto=None):
        self.h = C.c_void_p()
        if proto is None:
            _check(lib().s21_ckt_new(C.byref(self.h)))
        else:
            enc = proto if isinstance(proto, (bytes, bytearray)) else proto.SerializeToString()
            _check(lib().s21_ckt_from_proto(bytes(enc), len(enc), C.byref(self.h)))
        self.names = None

    def __del__(self):
        if getattr(self, "h", None):
            lib().s21_ckt_destroy(self.h)
            self.h = None

    def signal(self, name, module=None):
        _check(lib().s21_ckt_signal(self.h, _b(module) if module else None, _b(name)))

    def r(self, name, p, n, g, module=None):
        _check(lib().s21_ckt_add_r(self.h, _b(module) if module else None, _b(name), _b(p), _b(n), g))

    def c(self, name, p, n, c, module=None):
        _check(lib().s21_ckt_add_c(self.h, _b(module) if module else None, _b(name), _b(p), _b(n), c))

    def i(self, name, p, n, dc, module=None):
        _check(lib().s21_ckt_add_i(self.h, _b(module) if module else None, _b(name), _b(p), _b(n), dc))

    def v_wave(self, name, p, n, dc, kind, params, acm=0.0, module=None):
        """Time-varying voltage source (extension, s21_ckt_add_v_wave): kind "pulse" (v1 v2 td tr tf pw per) or "sin"
        (vo va freq td theta); `dc` is the operating-point value."""
        w = np.ascontiguousarray(params, dtype=np.float64)
        _check(lib().s21_ckt_add_v_wave(self.h, _b(module) if module else None, _b(name), _b(p), _b(n), dc, acm, {"pulse": 1, "sin": 2}[kind], len(w),
                                        w.ctypes.data_as(C.c_void_p)))
        return self

    def v(self, name, p, n, dc, acm=0.0, module=None):
        _check(lib().s21_ckt_add_v(self.h, _b(module) if module else None, _b(name), _b(p), _b(n), dc, acm))

    def d(self, name, p, n, model, params, module=None):
        _check(lib().s21_ckt_add_d(self.h, _b(module) if module else None, _b(name), _b(p), _b(n), _b(model), _b(params)))

    def mos(self, name, model, params, d, g, s, b, module=None):
        _check(lib().s21_ckt_add_mos(self.h, _b(module) if module else None, _b(name), _b(model), _b(params), _b(d), _b(g), _b(s), _b(b)))

    def x(self, name, module_name, ports, module=None):
        keys, vals = _strarr(list(ports.keys())), _strarr(list(ports.values()))
        _check(lib().s21_ckt_add_x(self.h, _b(module) if module else None, _b(name), _b(module_name), len(ports), keys, vals))

    def def_module(self, name, ports):
        _check(lib().s21_ckt_def_module(self.h, _b(name), len(ports), _strarr(list(ports))))

    def define(self, kind, name, mos_type=0, **params):
        keys = _strarr(list(params.keys()))
        vals = np.array([float(v) for v in params.values()] + [0.0])
        _check(lib().s21_ckt_define(self.h, _b(kind), _b(name), int(mos_type), len(params), keys, vals.ctypes.data_as(C.c_void_p)))

    def elaborate(self, opts=None, ic=None):
        o = np.full(5, np.nan)
        for k, key in enumerate(("temp", "tnom", "gmin", "iabstol", "reltol")):
            if opts and opts.get(key) is not None:
                o[k] = opts[key]
        ic = ic or {}
        nodes = _strarr([str(k) for k in ic])
        vals = np.array([float(v) for v in ic.values()] + [0.0])
        _check(lib().s21_ckt_elaborate(self.h, o.ctypes.data_as(C.c_void_p), len(ic), nodes, vals.ctypes.data_as(C.c_void_p)))
        n = lib().s21_ckt_num_vars(self.h)
        self.names = [lib().s21_ckt_var_name(self.h, k).decode() for k in range(n)]
        return self

    @property
    def n_vars(self):
        return lib().s21_ckt_num_vars(self.h)

    @property
    def n_devices(self):
        return lib().s21_ckt_num_devices(self.h)

    def var_kinds(self):
        return [lib().s21_ckt_var_kind(self.h, k) for k in range(self.n_vars)]

    def stamp_map(self):
        er, ec, do, de = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        ne = C.c_size_t()
        _check(lib().s21_ckt_stamp_map(self.h, C.byref(er), C.byref(ec), C.byref(ne), C.byref(do), C.byref(de)))
        nd = self.n_devices
        as_np = lambda p, n: np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(max(n, 1),))[:n].copy()
        dev_off = as_np(do, nd + 1)
        return {"elem_row": as_np(er, ne.value), "elem_col": as_np(ec, ne.value), "dev_off": dev_off,
                "dev_elems": as_np(de, int(dev_off[-1]) if nd >= 0 else 0)}


    def jit_source(self, vals, mode=0, shape=1):
        """CUDA source of the run-time specialised Newton kernel for the plan these first-iteration matrix values give
        (s21_jit_source; host only). Returns (source, dynamic shared-memory bytes)."""
        v = np.ascontiguousarray(vals, dtype=np.float64)
        out, n, smem = C.c_void_p(), C.c_size_t(), C.c_size_t()
        _check(lib().s21_jit_source(self.h, mode, shape, v.ctypes.data_as(C.c_void_p), v.size, C.byref(out), C.byref(n), C.byref(smem)))
        try:
            return C.string_at(out, n.value).decode(), smem.value
        finally:
            lib().s21_free(out)


def selftest_div(n=1 << 24, seed=1):
    """(mismatches, [a, b, ours, ieee]) of the device division self-test (s21_selftest_div); needs a GPU."""
    bad, first = C.c_uint64(), np.zeros(4)
    _check(lib().s21_selftest_div(n, seed, C.byref(bad), first.ctypes.data_as(C.c_void_p)))
    return bad.value, first


def jit_check(source):
    """Compile a generated kernel source for sm_100a with NVRTC (s21_jit_check; no GPU needed)."""
    b = source.encode()
    _check(lib().s21_jit_check(b, len(b)))


class Batch:
    """``s21_batch``: B instances of one elaborated circuit resident on one GPU."""

    def __init__(self, ckt, B=1, device=0):
        self.ckt = ckt
        self.B = B
        self.N = ckt.n_vars
        self.h = C.c_void_p()
        self._views = None  # numpy views of the pinned result buffer, rebuilt only when the library moves it
        _check(lib().s21_batch_create(ckt.h, device, B, C.byref(self.h)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().s21_batch_destroy(self.h)
            self.h = None

    def set_stream(self, cuda_stream_ptr):
        _check(lib().s21_batch_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def override(self, spec, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        assert v.shape == (self.B,)
        _check(lib().s21_batch_override(self.h, _b(spec), v.ctypes.data_as(C.c_void_p)))

    def sync_params(self, force_upload=False):
        n = C.c_size_t()
        _check(lib().s21_batch_sync_params(self.h, 1 if force_upload else 0, C.byref(n)))
        return n.value

    def reset(self):
        _check(lib().s21_batch_reset(self.h))

    def set_aids(self, gmin_stepping=False, source_stepping=False):
        """Opt-in convergence aids for dcop (s21_batch_set_aids); the reference has none."""
        _check(lib().s21_batch_set_aids(self.h, (1 if gmin_stepping else 0) | (2 if source_stepping else 0)))

    def dcop(self):
        x = np.zeros((self.B, self.N))
        status = np.zeros(self.B, dtype=np.int32)
        iters = np.zeros(self.B, dtype=np.int32)
        _check(lib().s21_batch_dcop(self.h, x.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p), iters.ctypes.data_as(C.c_void_p)))
        return x, status, iters

    def dcop_view(self, want_x=True):
        """dcop with the results left in the library's pinned staging buffer (s21_batch_dcop_view): the returned arrays are
        read-only views, valid until the next solve or read on this batch — no host-side allocation or copy per call."""
        px, ps, pi = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(lib().s21_batch_dcop_view(self.h, C.byref(px) if want_x else None, C.byref(ps), C.byref(pi)))

        def view(p, ctype, shape):
            a = np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), shape=shape)
            a.flags.writeable = False
            return a

        x = view(px, C.c_double, (self.B, self.N)) if want_x else None
        return x, view(ps, C.c_int32, (self.B,)), view(pi, C.c_int32, (self.B,))

    def step_dcop_view(self, upload=True, reset=True):
        """One sweep step in one call (s21_batch_step_dcop_view): forced H2D of the parameter pool, cold start, dcop, results
        as read-only views into the library's pinned buffer. Returns (x, status, iters, h2d_bytes)."""
        px, ps, pi, nb = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_size_t()
        _check(lib().s21_batch_step_dcop_view(self.h, (1 if upload else 0) | (2 if reset else 0), C.byref(px), C.byref(ps), C.byref(pi), C.byref(nb)))
        if self._views is None or self._views[0] != (px.value, ps.value, pi.value):
            def view(p, ctype, shape):
                a = np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), shape=shape)
                a.flags.writeable = False
                return a
            self._views = ((px.value, ps.value, pi.value), view(px, C.c_double, (self.B, self.N)), view(ps, C.c_int32, (self.B,)), view(pi, C.c_int32, (self.B,)))
        return self._views[1], self._views[2], self._views[3], nb.value

    def dcop_device(self):
        _check(lib().s21_batch_dcop_device(self.h))

    def packed_device(self):
        """(device pointer, f64 words) of the last solve's results packed in HBM (s21_batch_packed_device): x rows [B][N], then
        status / iters / loads as int32. For handing to a collective without a host round trip."""
        p, n = C.c_void_p(), C.c_size_t()
        _check(lib().s21_batch_packed_device(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def wave_device(self):
        """(device pointer, T, n_save, stride) of the last transient's waveforms in HBM, [T][n_save][stride] f64."""
        p, T, ns, st = C.c_void_p(), C.c_size_t(), C.c_size_t(), C.c_size_t()
        _check(lib().s21_batch_wave_device(self.h, C.byref(p), C.byref(T), C.byref(ns), C.byref(st)))
        return p.value, T.value, ns.value, st.value

    def read(self, want_x=True):
        x = np.zeros((self.B, self.N)) if want_x else None
        status = np.zeros(self.B, dtype=np.int32)
        iters = np.zeros(self.B, dtype=np.int32)
        _check(lib().s21_batch_read(self.h, x.ctypes.data_as(C.c_void_p) if want_x else None, status.ctypes.data_as(C.c_void_p),
                                    iters.ctypes.data_as(C.c_void_p)))
        return x, status, iters

    def tran(self, tstep, tstop, save=None, want_wave=True, out=None):
        """want_wave=False leaves the waveforms in HBM (wave_device) and returns None for them. out: a C-contiguous float64
        array of shape [B][T][n_save] to receive the waveforms (a caller that repeats a sweep reuses its buffer: a fresh
        40 MB array costs more in first-touch page faults than the copy that fills it)."""
        T = lib().s21_tran_num_points(tstep, tstop)
        save = np.arange(self.N, dtype=np.int32) if save is None else np.ascontiguousarray(save, dtype=np.int32)
        time = np.zeros(T)
        if want_wave and out is not None:
            if out.shape != (self.B, T, len(save)) or out.dtype != np.float64 or not out.flags["C_CONTIGUOUS"]:
                raise ValueError(f"out must be a C-contiguous float64 array of shape {(self.B, T, len(save))}")
            wave = out
        else:
            wave = np.empty((self.B, T, len(save))) if want_wave else None  # every entry is written by the library
        status = np.zeros(self.B, dtype=np.int32)
        iters = np.zeros(self.B, dtype=np.int64)
        _check(lib().s21_batch_tran(self.h, tstep, tstop, save.ctypes.data_as(C.c_void_p), len(save), time.ctypes.data_as(C.c_void_p),
                                    wave.ctypes.data_as(C.c_void_p) if want_wave else None, status.ctypes.data_as(C.c_void_p),
                                    iters.ctypes.data_as(C.c_void_p)))
        return time, wave, status, iters

    def tran_adaptive(self, tstep, tstop, save=None, h0=0.0, hmin=0.0, hmax=0.0, trtol=0.0, reltol=0.0, vntol=0.0):
        """LTE-controlled adaptive-step transient on the print grid k * tstep (s21_batch_tran_adaptive). Returns
        (time, wave[B][T][n_save], status, iters, accepted_steps, rejected_steps); 0 = default for every control."""
        T = lib().s21_tran_num_points(tstep, tstop)
        save = np.arange(self.N, dtype=np.int32) if save is None else np.ascontiguousarray(save, dtype=np.int32)
        ctl = np.array([h0, hmin, hmax, trtol, reltol, vntol, 0.0])
        time, wave = np.zeros(T), np.zeros((self.B, T, len(save)))
        status, iters = np.zeros(self.B, dtype=np.int32), np.zeros(self.B, dtype=np.int64)
        acc, rej = np.zeros(self.B, dtype=np.int32), np.zeros(self.B, dtype=np.int32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        _check(lib().s21_batch_tran_adaptive(self.h, tstep, tstop, p(ctl), p(save), len(save), p(time), p(wave), p(status), p(iters), p(acc), p(rej)))
        return time, wave, status, iters, acc, rej

    def ac(self, freqs, out=None):
        """out: a C-contiguous complex128 array of shape [F][N] to receive the results (see tran)."""
        f = np.ascontiguousarray(freqs, dtype=np.float64)
        if out is not None:
            if out.shape != (len(f), self.N) or out.dtype != np.complex128 or not out.flags["C_CONTIGUOUS"]:
                raise ValueError(f"out must be a C-contiguous complex128 array of shape {(len(f), self.N)}")
            x = out.view(np.float64).reshape(len(f), self.N, 2)
        else:
            x = np.empty((len(f), self.N, 2))  # every entry is written by the library
        status = np.zeros(len(f), dtype=np.int32)
        iters = np.zeros(len(f), dtype=np.int32)
        _check(lib().s21_batch_ac(self.h, f.ctypes.data_as(C.c_void_p), len(f), x.ctypes.data_as(C.c_void_p),
                                  status.ctypes.data_as(C.c_void_p), iters.ctypes.data_as(C.c_void_p)))
        return x.view(np.complex128).reshape(len(f), self.N), status, iters

    def pivot_order(self):
        p = [C.c_void_p() for _ in range(5)]
        n, nnz = C.c_size_t(), C.c_size_t()
        _check(lib().s21_batch_pivot_order(self.h, C.byref(p[0]), C.byref(p[1]), C.byref(n), C.byref(p[2]), C.byref(p[3]), C.byref(p[4]),
                                           C.byref(nnz)))
        as_np = lambda q, m: np.ctypeslib.as_array(C.cast(q, C.POINTER(C.c_int32)), shape=(max(m, 1),))[:m].copy()
        return {"row_i2e": as_np(p[0], n.value), "col_i2e": as_np(p[1], n.value), "lu_row": as_np(p[2], nnz.value),
                "lu_col": as_np(p[3], nnz.value), "lu_fill": as_np(p[4], nnz.value)}

    def stats(self):
        s = np.zeros(8)
        _check(lib().s21_batch_stats(self.h, s.ctypes.data_as(C.c_void_p)))
        return {"launches": int(s[0]), "device_ms": float(s[1]), "iters": int(s[2]), "loads": int(s[3]), "nnz_a": int(s[4]),
                "nnz_lu": int(s[5]), "n": int(s[6]), "stamps": int(s[7])}

    def kernel_name(self):
        """Which Newton kernel the last solve ran on (s21_batch_kernel_name)."""
        return lib().s21_batch_kernel_name(self.h).decode()

    def plan_info(self):
        """Shape of the numeric plan of the last solve (s21_batch_plan_info): sizes, operation and level counts."""
        v = np.zeros(8, dtype=np.int64)
        _check(lib().s21_batch_plan_info(self.h, v.ctypes.data_as(C.c_void_p)))
        return {"n": int(v[0]), "nnz_lu": int(v[1]), "lu_ops": int(v[2]), "lu_levels": int(v[3]), "fw_ops": int(v[4]), "fw_levels": int(v[5]),
                "bw_levels": int(v[6]), "relaxed": bool(v[7])}

    def setup_stats(self):
        """Setup cost behind the solves (s21_batch_setup_stats): host symbolic seconds of this batch, NVRTC seconds / runs and
        cubin-cache hits of the process."""
        s = np.zeros(8)
        _check(lib().s21_batch_setup_stats(self.h, s.ctypes.data_as(C.c_void_p)))
        return {"symbolic_s": float(s[0]), "nvrtc_s": float(s[1]), "nvrtc_runs": int(s[2]), "disk_hits": int(s[3]), "mem_hits": int(s[4]),
                "weak_pivot_instances": int(s[5]), "repaired_instances": int(s[6]), "aided_instances": int(s[7])}


def sweep_partition(B, n_devices, g):
    """(first, count) of shard g: contiguous blocks of ceil(B / n_devices) instances (s21_sweep_partition; host only)."""
    f, c = C.c_size_t(), C.c_size_t()
    _check(lib().s21_sweep_partition(B, n_devices, g, C.byref(f), C.byref(c)))
    return f.value, c.value


class Sweep:
    """``s21_sweep``: B instances of one elaborated circuit split over several GPUs of this process (one host thread and
    stream per GPU inside the library); same calls as ``Batch`` with arrays covering the whole sweep."""

    def __init__(self, ckt, B, devices=None, n_devices=0):
        self.ckt, self.B, self.N = ckt, B, ckt.n_vars
        self.h = C.c_void_p()
        if devices is not None:
            d = np.ascontiguousarray(devices, dtype=np.int32)
            _check(lib().s21_sweep_create(ckt.h, len(d), d.ctypes.data_as(C.c_void_p), B, C.byref(self.h)))
        else:
            _check(lib().s21_sweep_create(ckt.h, n_devices, None, B, C.byref(self.h)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().s21_sweep_destroy(self.h)
            self.h = None

    @property
    def n_devices(self):
        return lib().s21_sweep_num_devices(self.h)

    def shards(self):
        out = []
        for g in range(self.n_devices):
            d, f, c = C.c_int32(), C.c_size_t(), C.c_size_t()
            _check(lib().s21_sweep_shard(self.h, g, C.byref(d), C.byref(f), C.byref(c)))
            out.append((d.value, f.value, c.value))
        return out

    def override(self, spec, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        assert v.shape == (self.B,)
        _check(lib().s21_sweep_override(self.h, _b(spec), v.ctypes.data_as(C.c_void_p)))

    def sync_params(self, force_upload=False):
        n = C.c_size_t()
        _check(lib().s21_sweep_sync_params(self.h, 1 if force_upload else 0, C.byref(n)))
        return n.value

    def reset(self):
        _check(lib().s21_sweep_reset(self.h))

    def dcop(self):
        x = np.zeros((self.B, self.N))
        status = np.zeros(self.B, dtype=np.int32)
        iters = np.zeros(self.B, dtype=np.int32)
        _check(lib().s21_sweep_dcop(self.h, x.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p), iters.ctypes.data_as(C.c_void_p)))
        return x, status, iters

    def dcop_view(self):
        px, ps, pi = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(lib().s21_sweep_dcop_view(self.h, C.byref(px), C.byref(ps), C.byref(pi)))

        def view(p, ctype, shape):
            a = np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), shape=shape)
            a.flags.writeable = False
            return a

        return view(px, C.c_double, (self.B, self.N)), view(ps, C.c_int32, (self.B,)), view(pi, C.c_int32, (self.B,))

    def tran(self, tstep, tstop, save=None):
        T = lib().s21_tran_num_points(tstep, tstop)
        save = np.arange(self.N, dtype=np.int32) if save is None else np.ascontiguousarray(save, dtype=np.int32)
        time = np.zeros(T)
        wave = np.zeros((self.B, T, len(save)))
        status = np.zeros(self.B, dtype=np.int32)
        iters = np.zeros(self.B, dtype=np.int64)
        _check(lib().s21_sweep_tran(self.h, tstep, tstop, save.ctypes.data_as(C.c_void_p), len(save), time.ctypes.data_as(C.c_void_p),
                                    wave.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p), iters.ctypes.data_as(C.c_void_p)))
        return time, wave, status, iters

    def ac(self, freqs):
        f = np.ascontiguousarray(freqs, dtype=np.float64)
        x = np.zeros((len(f), self.N, 2))
        status = np.zeros(len(f), dtype=np.int32)
        iters = np.zeros(len(f), dtype=np.int32)
        _check(lib().s21_sweep_ac(self.h, f.ctypes.data_as(C.c_void_p), len(f), x.ctypes.data_as(C.c_void_p),
                                  status.ctypes.data_as(C.c_void_p), iters.ctypes.data_as(C.c_void_p)))
        return x.view(np.complex128).reshape(len(f), self.N), status, iters

    def stats(self):
        s = np.zeros(8)
        _check(lib().s21_sweep_stats(self.h, s.ctypes.data_as(C.c_void_p)))
        return {"launches": int(s[0]), "device_ms": float(s[1]), "iters": int(s[2]), "loads": int(s[3]), "nnz_a": int(s[4]),
                "nnz_lu": int(s[5]), "n": int(s[6]), "stamps": int(s[7])}


def symbolic(n, rows, cols, vals):
    """Host-only symbolic phase on an arbitrary COO matrix (s21_symbolic). Returns status + pivot order + L+U pattern."""
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    vals = np.asarray(vals)
    width = 2 if np.iscomplexobj(vals) else 1
    v = np.ascontiguousarray(vals, dtype=np.complex128 if width == 2 else np.float64)
    cap = n * n + len(rows) + 8
    row_i2e, col_i2e = np.zeros(max(n, 1), dtype=np.int32), np.zeros(max(n, 1), dtype=np.int32)
    lr, lc, lf = (np.zeros(cap, dtype=np.int32) for _ in range(3))
    nnz = C.c_size_t()
    f = lib().s21_symbolic
    f.argtypes = [C.c_int32, C.c_size_t] + [C.c_void_p] * 3 + [C.c_int32] + [C.c_void_p] * 5 + [C.c_size_t, C.POINTER(C.c_size_t)]
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    st = f(n, len(rows), p(rows), p(cols), p(v), width, p(row_i2e), p(col_i2e), p(lr), p(lc), p(lf), cap, C.byref(nnz))
    k = nnz.value
    return {"status": st, "row_i2e": row_i2e[:n], "col_i2e": col_i2e[:n], "lu_row": lr[:k], "lu_col": lc[:k], "lu_fill": lf[:k]}


def ac_freqs(fstart, fstop, npts):
    n = lib().s21_ac_freqs(int(fstart), int(fstop), int(npts), None, 0)
    f = np.zeros(n)
    lib().s21_ac_freqs(int(fstart), int(fstop), int(npts), f.ctypes.data_as(C.c_void_p), n)
    return f
