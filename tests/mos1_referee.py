"""Second, independent referee for Mos1 DC / AC ("guard the restated oracle with an independent second
transcription of Mos1::op_stamp / load_ac"): a plain-Python/numpy transcription of

  * Mos1Model::resolve              spice21/src/comps/mos.rs:140-237   (the `tox` given / no `nsub` track)
  * Mos1InstanceParams::resolve     mos.rs:271-290
  * Mos1InternalParams::derive      mos.rs:320-476
  * Mos1::op_stamp (OP analysis)    mos.rs:649-893
  * Mos1::load_ac                   mos.rs:914-968   (with its quirks: `(G,dr)` pushed twice, intrinsic Meyer half-caps only)
  * Solver::<f64>::solve / Solver::<Complex>::solve   analysis.rs:169-210, 253-303 on DENSE numpy matrices

written from the Rust text, sharing nothing with oracle/ (C++) or the CUDA path. Test infrastructure only: small dense
circuits (a handful of nodes), pure-Python loops. The reference itself pins no AC value (its AC tests assert nothing
numeric), so agreement of three independent transcriptions — this one, oracle/, the GPU kernels — is what row a18 rests on.
"""
import math

import numpy as np

KB = 1.3806226e-23        # comps/mod.rs:24-37
Q = 1.6021918e-19
KB_OVER_Q = KB / Q
TEMP_REF = 300.15
KELVIN_TO_C = 273.15
SIO2_PERMITTIVITY = 3.9 * 8.854214871e-12


def resolve_model(mos_type=0, **s):
    """Mos1Model::resolve (mos.rs:140-237); `nsub` is not supported here (asserted)."""
    assert "nsub" not in s
    tnom = s["tnom"] + KELVIN_TO_C if "tnom" in s else TEMP_REF
    m = {"p": -1.0 if mos_type == 1 else 1.0, "tnom": tnom, "cox_per_area": 0.0,
         "vt0": s.get("vt0", 0.0), "kp": s.get("kp", 2.0e-5), "phi": s.get("phi", 0.6), "gamma": s.get("gamma", 0.0)}
    if "tox" in s:
        m["cox_per_area"] = SIO2_PERMITTIVITY / s["tox"]
        if "kp" not in s:
            m["kp"] = s.get("u0", 600.0) * m["cox_per_area"] * 1e-4
    for k, d in (("lambda", 0.0), ("pb", 0.8), ("cbd", 0.0), ("cbs", 0.0), ("cgso", 0.0), ("cgdo", 0.0), ("cgbo", 0.0), ("cj", 0.0),
                 ("cjsw", 0.0), ("mj", 0.5), ("mjsw", 0.5), ("is", 1.0e-14), ("js", 1.0e-8), ("ld", 0.0), ("fc", 0.5)):
        m[k] = s.get(k, d)
    for k in ("rd", "rs", "rsh"):
        m[k] = s.get(k)
    return m


def resolve_inst(**s):
    """Mos1InstanceParams::resolve (mos.rs:271-290)."""
    d = {"l": 1e-6, "w": 1e-6, "a_d": 1e-12, "a_s": 1e-12, "pd": 1e-6, "ps": 1e-6, "nrd": 1.0, "nrs": 1.0}
    d.update(s)
    return d


def derive(model, inst, temp=TEMP_REF):
    """Mos1InternalParams::derive (mos.rs:320-476): what op_stamp / load_ac read (junction capacitance terms omitted: the AC
    stamp never uses them and an OP analysis does not integrate them)."""
    tnom = model["tnom"]
    fact1 = tnom / TEMP_REF
    vtnom = tnom * KB_OVER_Q
    kt1 = KB * tnom
    egfet1 = 1.16 - (7.02e-4 * tnom ** 2) / (tnom + 1108.0)
    arg1 = -egfet1 / 2.0 / kt1 + 1.1150877 / (KB * 2.0 * TEMP_REF)
    pbfact1 = -2.0 * vtnom * (1.5 * math.log(fact1) + Q * arg1)
    kt = temp * KB
    vtherm = temp * KB_OVER_Q
    temp_ratio = temp / tnom
    fact2 = temp / TEMP_REF
    egfet = 1.16 - (7.02e-4 * temp ** 2) / (temp + 1108.0)
    arg = -egfet / 2.0 / kt + 1.1150877 / (KB * 2.0 * TEMP_REF)
    pbfact = -2.0 * vtherm * (1.5 * math.log(fact2) + Q * arg)
    leff = inst["l"] - 2.0 * model["ld"]
    p = model["p"]
    phio = (model["phi"] - pbfact1) / fact1
    phi_t = fact2 * phio + pbfact
    vbi_t = model["vt0"] - p * (model["gamma"] * math.sqrt(model["phi"])) + 0.5 * (egfet1 - egfet) + p * 0.5 * (phi_t - model["phi"])
    vt0_t = vbi_t + p * model["gamma"] * math.sqrt(phi_t)
    isat_t = model["is"] * math.exp(-egfet / vtherm + egfet1 / vtnom)
    jsat_t = model["js"] * math.exp(-egfet / vtherm + egfet1 / vtnom)
    use_default_isat = jsat_t == 0.0 or inst["a_d"] == 0.0 or inst["a_s"] == 0.0
    isat_d = isat_t if use_default_isat else jsat_t * inst["a_d"]
    isat_s = isat_t if use_default_isat else jsat_t * inst["a_s"]

    def gr(r, n):
        if model[r] is not None:
            return 0.0 if model[r] <= 0.0 else 1.0 / model[r]
        if model["rsh"] is not None:
            return 0.0 if model["rsh"] <= 0.0 else 1.0 / model["rsh"] / inst[n]
        return 0.0

    kp_t = model["kp"] / temp_ratio * math.sqrt(temp_ratio)
    return {"vt0_t": vt0_t, "phi_t": phi_t, "vtherm": vtherm, "beta": kp_t * inst["w"] / leff, "cox": model["cox_per_area"] * leff * inst["w"],
            "isat_d": isat_d, "isat_s": isat_s, "grs": gr("rs", "nrs"), "grd": gr("rd", "nrd")}


def op_stamp(model, ip, vd, vg, vs, vb, gmin=1e-12):
    """Mos1::op_stamp for an OP analysis (mos.rs:649-893) on a device whose dp/sp alias d/s (no rd/rs/rsh).
    Returns (op point dict, G stamps [(row, col, value)] over terminals "d","g","s","b", rhs stamps [(terminal, value)])."""
    p = model["p"]
    reversed_ = p * (vd - vs) < 0.0
    vd_, vs_ = (vs, vd) if reversed_ else (vd, vs)
    vgs = p * (vg - vs_)
    vds = p * (vd_ - vs_)
    vsb = p * (vs_ - vb)
    vdb = p * (vd_ - vb)
    von = ip["vt0_t"] + model["gamma"] * (math.sqrt(ip["phi_t"] + vsb) - math.sqrt(ip["phi_t"])) if vsb > 0.0 else ip["vt0_t"]
    vov = vgs - von
    vdsat = max(vov, 0.0)
    ids = gm = gds = gmbs = 0.0
    lam, beta = model["lambda"], ip["beta"]
    if vov > 0.0:
        if vds >= vov:
            ids = beta / 2.0 * vov ** 2 * (1.0 + lam * vds)
            gm = beta * vov * (1.0 + lam * vds)
            gds = lam * beta / 2.0 * vov ** 2
        else:
            ids = beta * (vov * vds - vds ** 2 / 2.0) * (1.0 + lam * vds)
            gm = beta * vds * (1.0 + lam * vds)
            gds = beta * ((vov - vds) * (1.0 + lam * vds) + lam * ((vov * vds) - vds ** 2 / 2.0))
        gmbs = gm * model["gamma"] / 2.0 / math.sqrt(ip["phi_t"] + vsb) if ip["phi_t"] + vsb > 0.0 else 0.0
    vt = ip["vtherm"]
    isat_bs, isat_bd = (ip["isat_s"], ip["isat_d"]) if not reversed_ else (ip["isat_d"], ip["isat_s"])
    ibs = isat_bs * (math.exp(-vsb / vt) - 1.0)
    gbs = (isat_bs / vt) * math.exp(-vsb / vt) + gmin
    ibs_rhs = ibs + vsb * gbs
    ibd = isat_bd * (math.exp(-vdb / vt) - 1.0)
    gbd = (isat_bd / vt) * math.exp(-vdb / vt) + gmin
    ibd_rhs = ibd + vdb * gbd
    cox, phi = ip["cox"], ip["phi_t"]
    if vov <= -phi:
        cgb1, cgs1, cgd1 = cox / 2.0, 0.0, 0.0
    elif vov <= -phi / 2.0:
        cgb1, cgs1, cgd1 = -vov * cox / (2.0 * phi), 0.0, 0.0
    elif vov <= 0.0:
        cgb1, cgs1, cgd1 = -vov * cox / (2.0 * phi), vov * cox / (1.5 * phi) + cox / 3.0, 0.0
    elif vdsat <= vds:
        cgs1, cgd1, cgb1 = cox / 3.0, 0.0, 0.0
    else:
        vddif, vddif1 = 2.0 * vdsat - vds, vdsat - vds
        vddif2 = vddif * vddif
        cgd1 = cox * (1.0 - vdsat * vdsat / vddif2) / 3.0
        cgs1 = cox * (1.0 - vddif1 * vddif1 / vddif2) / 3.0
        cgb1 = 0.0
    irhs = ids - gm * vgs - gds * vds
    sr, dr = ("s", "d") if not reversed_ else ("d", "s")
    grd, grs = ip["grd"], ip["grs"]
    assert grd == 0.0 and grs == 0.0
    g = [(dr, dr, gds + grd + gbd), (sr, sr, gm + gds + grs + gbs + gmbs), (dr, sr, -gm - gds - gmbs), (sr, dr, -gds), (dr, "g", gm),
         (sr, "g", -gm), ("b", "b", gbd + gbs), ("b", dr, -gbd), ("b", sr, -gbs), (dr, "b", -gbd + gmbs), (sr, "b", -gbs - gmbs)]
    b = [(dr, p * (-irhs + ibd_rhs)), (sr, p * (irhs + ibs_rhs)), ("b", -p * (ibd_rhs + ibs_rhs))]
    op = {"gm": gm, "gds": gds, "gmbs": gmbs, "gbs": gbs, "gbd": gbd, "cgs": cgs1, "cgd": cgd1, "cgb": cgb1, "reversed": reversed_, "ids": ids}
    return op, g, b


def ac_stamp(op, omega):
    """Mos1::load_ac (mos.rs:914-968), dp/sp aliasing d/s, grd = grs = 0. `(G,dr)` appears twice, as in the reference."""
    gm, gds, gmbs, gbs, gbd = op["gm"], op["gds"], op["gmbs"], op["gbs"], op["gbd"]
    gcgs, gcgd, gcgb = omega * op["cgs"], omega * op["cgd"], omega * op["cgb"]
    sr, dr = ("s", "d") if not op["reversed"] else ("d", "s")
    return [(dr, dr, complex(gds + gbd, gcgd)), (sr, sr, complex(gm + gds + gbs + gmbs, gcgs)), (dr, sr, complex(-gm - gds - gmbs, 0.0)),
            (sr, dr, complex(-gds, 0.0)), (dr, "g", complex(gm, -gcgd)), (sr, "g", complex(-gm, -gcgs)), ("g", "g", complex(0.0, gcgd + gcgs + gcgb)),
            ("b", "b", complex(gbd + gbs, gcgb)), ("g", "b", complex(0.0, -gcgb)), ("g", dr, complex(0.0, -gcgd)), ("g", sr, complex(0.0, -gcgs)),
            ("b", "g", complex(0.0, -gcgb)), ("g", dr, complex(0.0, -gcgd)), ("b", dr, complex(-gbd, 0.0)), ("b", sr, complex(-gbs, 0.0)),
            (dr, "b", complex(-gbd + gmbs, 0.0)), (sr, "b", complex(-gbs - gmbs, 0.0))]


class Dense:
    """A small circuit on dense matrices. Devices: ("R", p, n, g) ("C", p, n, c) ("V", name, p, n, dc, acm)
    ("M", model, intparams, d, g, s, b). Node "" is ground; unknowns = nodes in first-encounter order, then V branch currents
    named after the source."""

    def __init__(self, devices):
        self.devices = devices
        self.names = []
        for d in devices:
            nodes = {"R": d[1:3], "C": d[1:3], "V": d[2:4], "M": d[3:7]}[d[0]]
            for n in nodes:
                if n and n not in self.names:
                    self.names.append(n)
        for d in devices:
            if d[0] == "V":
                self.names.append(d[1])
        self.ix = {n: k for k, n in enumerate(self.names)}

    def _i(self, n):
        return self.ix[n] if n else None

    def _add(self, A, r, c, v):
        if r is not None and c is not None:
            A[r, c] += v

    def load(self, x):
        """Solver::update for an OP analysis: (A, rhs, [Mos1 op points])."""
        N = len(self.names)
        A, b, ops = np.zeros((N, N)), np.zeros(N), []
        volt = lambda n: x[self.ix[n]] if n else 0.0
        for d in self.devices:
            if d[0] == "R":
                p, n, g = self._i(d[1]), self._i(d[2]), d[3]
                for r, c, v in ((p, p, g), (p, n, -g), (n, p, -g), (n, n, g)):
                    self._add(A, r, c, v)
            elif d[0] == "V":
                i, p, n = self.ix[d[1]], self._i(d[2]), self._i(d[3])
                for r, c, v in ((p, i, 1.0), (i, p, 1.0), (n, i, -1.0), (i, n, -1.0)):
                    self._add(A, r, c, v)
                b[i] += d[4]
            elif d[0] == "M":
                t = {"d": d[3], "g": d[4], "s": d[5], "b": d[6]}
                op, g, rhs = op_stamp(d[1], d[2], volt(t["d"]), volt(t["g"]), volt(t["s"]), volt(t["b"]))
                ops.append(op)
                for r, c, v in g:
                    self._add(A, self._i(t[r]), self._i(t[c]), v)
                for r, v in rhs:
                    if self._i(t[r]) is not None:
                        b[self._i(t[r])] += v
        return A, b, ops

    def dcop(self):
        """Solver::<f64>::solve (analysis.rs:169-210): returns (x, iterations that reached the linear solve, Mos1 op points)."""
        x, dx = np.zeros(len(self.names)), np.zeros(len(self.names))
        for k in range(100):
            A, b, ops = self.load(x)
            res = b - A @ x
            if np.all(np.abs(dx) <= 1e-3) and np.all(np.abs(res) <= 1e-12):
                return x, k, ops
            dx = np.linalg.solve(A, res)
            m = np.max(np.abs(dx))
            if m > 1.0:
                dx = dx * 1.0 / m
            x = x + dx
        raise RuntimeError("Convergence Failed")

    def ac(self, ops, freqs, direct=False):
        """Solver::<Complex>::solve per frequency (analysis.rs:253-303), each point from x = 0: x[F][N] complex.
        direct=True: the plain linear solve x = A^-1 b (what the Newton shell converges to when its 1.0 step limit lets it)."""
        N = len(self.names)
        out = np.zeros((len(freqs), N), dtype=complex)
        for fi, f in enumerate(freqs):
            om = 2.0 * math.pi * f
            A, b, k = np.zeros((N, N), dtype=complex), np.zeros(N, dtype=complex), 0
            for d in self.devices:
                if d[0] == "R":
                    p, n, g = self._i(d[1]), self._i(d[2]), d[3]
                    for r, c, v in ((p, p, g), (p, n, -g), (n, p, -g), (n, n, g)):
                        self._add(A, r, c, v)
                elif d[0] == "C":
                    p, n, y = self._i(d[1]), self._i(d[2]), complex(0.0, om * d[3])
                    for r, c, v in ((p, p, y), (p, n, -y), (n, p, -y), (n, n, y)):
                        self._add(A, r, c, v)
                elif d[0] == "V":
                    i, p, n = self.ix[d[1]], self._i(d[2]), self._i(d[3])
                    for r, c, v in ((p, i, 1.0), (i, p, 1.0), (n, i, -1.0), (i, n, -1.0)):
                        self._add(A, r, c, v)
                    b[i] += d[5]
                elif d[0] == "M":
                    t = {"d": d[3], "g": d[4], "s": d[5], "b": d[6]}
                    for r, c, v in ac_stamp(ops[k], om):
                        self._add(A, self._i(t[r]), self._i(t[c]), v)
                    k += 1
            if direct:
                out[fi] = np.linalg.solve(A, b)
                continue
            x, dx = np.zeros(N, dtype=complex), np.zeros(N, dtype=complex)
            for _ in range(20):
                res = b - A @ x
                if np.all(np.abs(dx) < 1e-3) and np.all(np.abs(res) < 1e-9):
                    break
                dx = np.linalg.solve(A, res)
                m = np.max(np.abs(dx))
                if m > 1.0:
                    dx = dx * 1.0 / m
                x = x + dx
            out[fi] = x
        return out
