"""Protobuf message classes for Spice21's wire format, built at import time from programmatic descriptors.

The reference ships ``spice21/protos/{spice21,mos,bsim4}.proto`` and generates Python classes with protoc
(``spice21py/spice21py/protos``). This image has no protoc, so the same schema (package ``spice21``, identical
message names, field names and field numbers) is declared here through ``descriptor_pb2``. Bytes produced by these
classes are what ``libspice21cu.so``'s ``s21_*_bytes`` entry points consume, and vice versa.
"""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
from google.protobuf import wrappers_pb2  # noqa: F401  (registers google/protobuf/wrappers.proto in the default pool)

_F = descriptor_pb2.FieldDescriptorProto
_DV = ".google.protobuf.DoubleValue"
_UV = ".google.protobuf.UInt64Value"
_IV = ".google.protobuf.Int64Value"


def _msg(fd, name):
    m = fd.message_type.add()
    m.name = name
    return m


def _field(m, name, number, ftype, type_name=None, repeated=False, oneof=None):
    f = m.field.add()
    f.name = name
    f.number = number
    f.type = ftype
    f.label = _F.LABEL_REPEATED if repeated else _F.LABEL_OPTIONAL
    if type_name:
        f.type_name = type_name
    if oneof is not None:
        f.oneof_index = oneof
    return f


def _map(m, name, number, vtype, vtype_name=None):
    """map<string, V> = a repeated nested FooEntry{key=1, value=2} with map_entry set."""
    entry_name = "".join(p.capitalize() for p in name.split("_")) + "Entry"
    e = m.nested_type.add()
    e.name = entry_name
    e.options.map_entry = True
    _field(e, "key", 1, _F.TYPE_STRING)
    _field(e, "value", 2, vtype, vtype_name)
    _field(m, name, number, _F.TYPE_MESSAGE, ".spice21." + m.name + "." + entry_name, repeated=True)


def _two_term(fd, name, valname, extra=()):
    m = _msg(fd, name)
    _field(m, "name", 1, _F.TYPE_STRING)
    _field(m, "p", 2, _F.TYPE_STRING)
    _field(m, "n", 3, _F.TYPE_STRING)
    _field(m, valname, 4, _F.TYPE_DOUBLE)
    for k, (n, num) in enumerate(extra):
        _field(m, n, num, _F.TYPE_DOUBLE)
    return m


def _wrapped(m, fields, wrapper=_DV):
    for name, num in fields:
        _field(m, name, num, _F.TYPE_MESSAGE, wrapper)


def _build():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "spice21_b200/spice21.proto"
    fd.package = "spice21"
    fd.syntax = "proto3"
    fd.dependency.append("google/protobuf/wrappers.proto")

    # ---- mos.proto
    en = fd.enum_type.add()
    en.name = "MosType"
    for k, v in (("NMOS", 0), ("PMOS", 1)):
        ev = en.value.add()
        ev.name, ev.number = k, v
    m = _msg(fd, "MosPorts")
    for k, n in enumerate(("d", "g", "s", "b")):
        _field(m, n, k + 1, _F.TYPE_STRING)
    m = _msg(fd, "Mos")
    _field(m, "name", 1, _F.TYPE_STRING)
    _field(m, "model", 2, _F.TYPE_STRING)
    _field(m, "params", 3, _F.TYPE_STRING)
    _field(m, "ports", 4, _F.TYPE_MESSAGE, ".spice21.MosPorts")
    m = _msg(fd, "Mos1InstParams")
    _field(m, "name", 1, _F.TYPE_STRING)
    _wrapped(m, [("m", 2), ("l", 3), ("w", 4), ("a_d", 5), ("a_s", 6), ("pd", 7), ("ps", 8), ("nrd", 9), ("nrs", 10), ("temp", 11)])
    m = _msg(fd, "Mos1Model")
    _field(m, "name", 1, _F.TYPE_STRING)
    _field(m, "mos_type", 2, _F.TYPE_ENUM, ".spice21.MosType")
    _wrapped(m, [("vt0", 3), ("kp", 4), ("gamma", 5), ("phi", 6), ("lambda", 7), ("rd", 8), ("rs", 9), ("cbd", 10), ("cbs", 11),
                 ("is", 12), ("pb", 13), ("cgso", 14), ("cgdo", 15), ("cgbo", 16), ("rsh", 17), ("cj", 18), ("mj", 19), ("cjsw", 20),
                 ("mjsw", 21), ("js", 22), ("tox", 23), ("ld", 24), ("u0", 25), ("fc", 26), ("nsub", 27), ("nss", 29), ("tnom", 30),
                 ("kf", 31), ("af", 32)])
    _field(m, "tpg", 28, _F.TYPE_MESSAGE, _IV)

    # ---- bsim4.proto (the model message carries only mos_type and name on the wire)
    m = _msg(fd, "Bsim4InstParams")
    _wrapped(m, [("l", 1), ("w", 2), ("nf", 3), ("sa", 4), ("sb", 5), ("sd", 6), ("sca", 7), ("scb", 8), ("scc", 9), ("sc", 10),
                 ("ad", 11), ("as", 12), ("pd", 13), ("ps", 14), ("nrd", 15), ("nrs", 16)])
    _wrapped(m, [("min", 17), ("rgeomod", 18)], _UV)
    _wrapped(m, [("rbdb", 19), ("rbsb", 20), ("rbpb", 21), ("rbps", 22), ("rbpd", 23), ("delvto", 24), ("xgw", 25), ("ngcon", 26)])
    _wrapped(m, [("trnqsmod", 27), ("acnqsmod", 28), ("rbodymod", 29), ("rgatemod", 30), ("geomod", 31)], _UV)
    _field(m, "name", 40, _F.TYPE_STRING)
    m = _msg(fd, "Bsim4ModelParam")  # extension: one model-card parameter by its Bsim4ModelSpecs field name
    _field(m, "name", 1, _F.TYPE_STRING)
    _field(m, "value", 2, _F.TYPE_DOUBLE)
    m = _msg(fd, "Bsim4Model")
    _field(m, "mos_type", 1, _F.TYPE_ENUM, ".spice21.MosType")
    _field(m, "name", 900, _F.TYPE_STRING)
    # extension field (the reference's message stops at `name`; its 876 model fields are commented out, bsim4.proto:51-929)
    _field(m, "params", 901, _F.TYPE_MESSAGE, ".spice21.Bsim4ModelParam", repeated=True)

    # ---- spice21.proto
    _two_term(fd, "Resistor", "g")
    _two_term(fd, "Capacitor", "c")
    _two_term(fd, "Isrc", "dc")
    m = _two_term(fd, "Vsrc", "dc", extra=(("acm", 5),))
    # extension fields (not in the reference's spice21.proto:29-35; its decoder skips unknown fields): a time-varying source,
    # wave_kind 1 = PULSE(v1 v2 td tr tf pw per), 2 = SIN(vo va freq td theta) — include/spice21cu.h s21_ckt_add_v_wave
    _field(m, "wave_kind", 6, _F.TYPE_INT32)
    _field(m, "wave", 7, _F.TYPE_DOUBLE, repeated=True)
    m = _msg(fd, "TwoTerms")
    _field(m, "p", 2, _F.TYPE_STRING)
    _field(m, "n", 3, _F.TYPE_STRING)
    m = _msg(fd, "DiodeModel")
    _field(m, "name", 1, _F.TYPE_STRING)
    _wrapped(m, [("tnom", 2), ("is", 3), ("n", 4), ("tt", 5), ("vj", 6), ("m", 7), ("eg", 8), ("xti", 9), ("kf", 10), ("af", 11),
                 ("fc", 12), ("bv", 13), ("ibv", 14), ("rs", 15), ("cj0", 16)])
    m = _msg(fd, "DiodeInstParams")
    _field(m, "name", 1, _F.TYPE_STRING)
    _field(m, "model", 2, _F.TYPE_STRING)
    _wrapped(m, [("area", 4), ("temp", 5)])
    m = _msg(fd, "Diode")
    for k, n in enumerate(("name", "p", "n", "model", "params")):
        _field(m, n, k + 1, _F.TYPE_STRING)
    m = _msg(fd, "Instance")
    m.oneof_decl.add().name = "comp"
    for n, num, t in (("r", 1, "Resistor"), ("c", 2, "Capacitor"), ("m", 3, "Mos"), ("i", 4, "Isrc"), ("v", 5, "Vsrc"), ("d", 6, "Diode"),
                      ("x", 7, "ModuleInstance")):
        _field(m, n, num, _F.TYPE_MESSAGE, ".spice21." + t, oneof=0)
    m = _msg(fd, "Module")
    _field(m, "name", 1, _F.TYPE_STRING)
    _field(m, "ports", 2, _F.TYPE_STRING, repeated=True)
    _field(m, "signals", 4, _F.TYPE_STRING, repeated=True)
    _field(m, "comps", 5, _F.TYPE_MESSAGE, ".spice21.Instance", repeated=True)
    _map(m, "params", 9, _F.TYPE_DOUBLE)
    m = _msg(fd, "ModuleInstance")
    _field(m, "name", 1, _F.TYPE_STRING)
    _field(m, "module", 2, _F.TYPE_STRING)
    _map(m, "ports", 3, _F.TYPE_STRING)
    _map(m, "params", 4, _F.TYPE_DOUBLE)
    m = _msg(fd, "Def")
    m.oneof_decl.add().name = "defines"
    for n, num, t in (("module", 1, "Module"), ("diodemodel", 2, "DiodeModel"), ("diodeinst", 3, "DiodeInstParams"),
                      ("bsim4model", 4, "Bsim4Model"), ("bsim4inst", 5, "Bsim4InstParams"), ("mos1model", 6, "Mos1Model"),
                      ("mos1inst", 7, "Mos1InstParams")):
        _field(m, n, num, _F.TYPE_MESSAGE, ".spice21." + t, oneof=0)
    m = _msg(fd, "Defs")
    _field(m, "defs", 1, _F.TYPE_MESSAGE, ".spice21.Def", repeated=True)
    m = _msg(fd, "Circuit")
    _field(m, "name", 1, _F.TYPE_STRING)
    _field(m, "signals", 2, _F.TYPE_STRING, repeated=True)
    _field(m, "defs", 3, _F.TYPE_MESSAGE, ".spice21.Def", repeated=True)
    _field(m, "comps", 4, _F.TYPE_MESSAGE, ".spice21.Instance", repeated=True)
    m = _msg(fd, "SimOptions")
    _wrapped(m, [("temp", 1), ("tnom", 2), ("gmin", 3), ("iabstol", 4), ("reltol", 5)])
    m = _msg(fd, "Op")
    _field(m, "ckt", 1, _F.TYPE_MESSAGE, ".spice21.Circuit")
    _field(m, "opts", 2, _F.TYPE_MESSAGE, ".spice21.SimOptions")
    m = _msg(fd, "OpResult")
    _map(m, "vals", 1, _F.TYPE_DOUBLE)
    m = _msg(fd, "TranOptions")
    _field(m, "tstop", 1, _F.TYPE_DOUBLE)
    _field(m, "tstep", 2, _F.TYPE_DOUBLE)
    _map(m, "ic", 3, _F.TYPE_DOUBLE)
    m = _msg(fd, "Tran")
    _field(m, "ckt", 1, _F.TYPE_MESSAGE, ".spice21.Circuit")
    _field(m, "opts", 2, _F.TYPE_MESSAGE, ".spice21.SimOptions")
    _field(m, "args", 3, _F.TYPE_MESSAGE, ".spice21.TranOptions")
    m = _msg(fd, "DoubleArray")
    _field(m, "vals", 1, _F.TYPE_DOUBLE, repeated=True)
    m = _msg(fd, "TranResult")
    _field(m, "time", 1, _F.TYPE_MESSAGE, ".spice21.DoubleArray")
    _map(m, "vals", 2, _F.TYPE_MESSAGE, ".spice21.DoubleArray")
    m = _msg(fd, "ComplexNum")
    _field(m, "re", 1, _F.TYPE_DOUBLE)
    _field(m, "im", 2, _F.TYPE_DOUBLE)
    m = _msg(fd, "ComplexArray")
    _field(m, "vals", 1, _F.TYPE_MESSAGE, ".spice21.ComplexNum", repeated=True)
    m = _msg(fd, "AcOptions")
    _field(m, "fstart", 1, _F.TYPE_UINT64)
    _field(m, "fstop", 2, _F.TYPE_UINT64)
    _field(m, "npts", 3, _F.TYPE_UINT64)
    m = _msg(fd, "Ac")
    _field(m, "ckt", 1, _F.TYPE_MESSAGE, ".spice21.Circuit")
    _field(m, "opts", 2, _F.TYPE_MESSAGE, ".spice21.SimOptions")
    _field(m, "args", 3, _F.TYPE_MESSAGE, ".spice21.AcOptions")
    m = _msg(fd, "AcResult")
    _field(m, "freq", 1, _F.TYPE_MESSAGE, ".spice21.DoubleArray")
    _map(m, "vals", 2, _F.TYPE_MESSAGE, ".spice21.ComplexArray")
    return fd


_pool = descriptor_pool.Default()
_file = _pool.Add(_build()) if hasattr(_pool, "Add") else None
if _file is None:
    _pool.AddSerializedFile(_build().SerializeToString())
_NAMES = ["MosPorts", "Mos", "Mos1InstParams", "Mos1Model", "Bsim4InstParams", "Bsim4ModelParam", "Bsim4Model", "Resistor", "Capacitor", "Isrc", "Vsrc",
          "TwoTerms", "DiodeModel", "DiodeInstParams", "Diode", "Instance", "Module", "ModuleInstance", "Def", "Defs", "Circuit",
          "SimOptions", "Op", "OpResult", "TranOptions", "Tran", "DoubleArray", "TranResult", "ComplexNum", "ComplexArray", "AcOptions",
          "Ac", "AcResult"]
for _n in _NAMES:
    globals()[_n] = message_factory.GetMessageClass(_pool.FindMessageTypeByName("spice21." + _n))
MosType = _pool.FindEnumTypeByName("spice21.MosType")
NMOS, PMOS = 0, 1
__all__ = _NAMES + ["MosType", "NMOS", "PMOS"]
