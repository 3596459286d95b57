// ORACLE — TEST INFRASTRUCTURE ONLY (see num.hpp header).
//
// CPU restatement of the reference's devices:
//   spice21/src/comps/mod.rs:95-346    Vsrc, Capacitor, Resistor, Isrc
//   spice21/src/comps/diode.rs:52-389  DiodeModel, DiodeIntParams::derive, Diode::{limit,load}, Diode0
//   spice21/src/comps/mos.rs:140-1099  Mos1Model::resolve, Mos1InstanceParams::resolve,
//                                      Mos1InternalParams::derive, MosJunction::qc, Mos1::{op_stamp,load,load_ac}, Mos0
// Reference quirks are reproduced on purpose and flagged with "QUIRK".
#pragma once
#include <map>

#include "analysis.hpp"

namespace orc {

#define ORC_CME_BOTH                                                        \
  void create_matrix_elems(Matrix<double>& m) override { cme(m); }          \
  void create_matrix_elems(Matrix<Cplx>& m) override { cme(m); }

// ---------------------------------------------------------------- Vsrc  (comps/mod.rs:95-150)
struct Vsrc : Component {
  double v, acm;
  VarIndex p, n, ivar;
  Eindex pi = -1, ip = -1, ni = -1, in_ = -1;
  Vsrc(double vdc, double acm_, VarIndex p_, VarIndex n_, VarIndex ivar_) : v(vdc), acm(acm_), p(p_), n(n_), ivar(ivar_) {}
  void update(double val) override { v = val; }
  template <class T> void cme(Matrix<T>& mat) {
    pi = make_matrix_elem(mat, p, ivar);
    ip = make_matrix_elem(mat, ivar, p);
    ni = make_matrix_elem(mat, n, ivar);
    in_ = make_matrix_elem(mat, ivar, n);
  }
  ORC_CME_BOTH
  Stamps<double> load(const Variables<double>&, const AnalysisInfo&, const Options&) override {
    Stamps<double> s;
    s.g = {{pi, 1.0}, {ip, 1.0}, {ni, -1.0}, {in_, -1.0}};
    s.b = {{ivar, v}};
    return s;
  }
  Stamps<Cplx> load_ac(const Variables<Cplx>&, const AnalysisInfo&, const Options&) override {
    Stamps<Cplx> s;
    s.g = {{pi, Cplx(1.0, 0.0)}, {ip, Cplx(1.0, 0.0)}, {ni, Cplx(-1.0, 0.0)}, {in_, Cplx(-1.0, 0.0)}};
    s.b = {{ivar, Cplx(acm, 0.0)}};
    return s;
  }
  void matps_list(std::vector<Eindex>& o) const override { o = {pi, ip, ni, in_}; }
  const char* kind_name() const override { return "V"; }
};

// ---------------------------------------------------------------- Capacitor  (comps/mod.rs:152-239)
struct Capacitor : Component {
  double c;
  VarIndex p, n;
  Eindex pp = -1, nn = -1, pn = -1, np = -1;
  struct OpPoint { double v = 0, q = 0, i = 0; } op, guess;
  Capacitor(double c_, VarIndex p_, VarIndex n_) : c(c_), p(p_), n(n_) {}
  double q(double v) const { return c * v; }
  double dq_dv(double) const { return c; }
  template <class T> void cme(Matrix<T>& mat) {
    pp = make_matrix_elem(mat, p, p);
    pn = make_matrix_elem(mat, p, n);
    np = make_matrix_elem(mat, n, p);
    nn = make_matrix_elem(mat, n, n);
  }
  ORC_CME_BOTH
  void commit() override { op = guess; }
  Stamps<double> load(const Variables<double>& guess_, const AnalysisInfo& an, const Options&) override {
    double vd = guess_.get(p) - guess_.get(n);
    double q_ = q(vd);
    Stamps<double> s;
    if (an.kind == AnalysisInfo::OP) {
      guess = {vd, q_, 0.0};
      return s;
    } else if (an.kind == AnalysisInfo::TRAN) {
      double g, i, rhs;
      an.tran->integrate(q_ - op.q, dq_dv(vd), vd, op.i, &g, &i, &rhs);
      guess = {vd, q_, i};
      s.g = {{pp, g}, {nn, g}, {pn, -g}, {np, -g}};
      s.b = {{p, -rhs}, {n, rhs}};
      return s;
    }
    throw Panic("HOW WE GET HERE?!?");
  }
  Stamps<Cplx> load_ac(const Variables<Cplx>&, const AnalysisInfo& an, const Options&) override {
    if (an.kind != AnalysisInfo::AC) throw Panic("Invalid AC AnalysisInfo");
    double c_ = dq_dv(0.0);
    double om = an.ac->omega;
    Stamps<Cplx> s;
    s.g = {{pp, Cplx(0.0, om * c_)}, {nn, Cplx(0.0, om * c_)}, {pn, Cplx(0.0, -om * c_)}, {np, Cplx(0.0, -om * c_)}};
    return s;
  }
  void matps_list(std::vector<Eindex>& o) const override { o = {pp, pn, np, nn}; }
  const char* kind_name() const override { return "C"; }
};

// ---------------------------------------------------------------- Resistor  (comps/mod.rs:271-323)
struct Resistor : Component {
  double g;
  VarIndex terms[2];
  Eindex matps[2][2] = {{-1, -1}, {-1, -1}};
  Resistor(double g_, VarIndex p, VarIndex n) : g(g_) { terms[0] = p; terms[1] = n; }
  void update(double val) override { g = val; }
  template <class T> void cme(Matrix<T>& mat) {
    for (int l = 0; l < 2; l++)
      for (int r = 0; r < 2; r++) matps[l][r] = make_matrix_elem(mat, terms[l], terms[r]);
  }
  ORC_CME_BOTH
  Stamps<double> load(const Variables<double>&, const AnalysisInfo&, const Options&) override {
    Stamps<double> s;
    s.g = {{matps[0][0], g}, {matps[1][1], g}, {matps[0][1], -g}, {matps[1][0], -g}};
    return s;
  }
  Stamps<Cplx> load_ac(const Variables<Cplx>&, const AnalysisInfo&, const Options&) override {
    Stamps<Cplx> s;
    s.g = {{matps[0][0], Cplx(g, 0.0)}, {matps[1][1], Cplx(g, 0.0)}, {matps[0][1], Cplx(-g, 0.0)}, {matps[1][0], Cplx(-g, 0.0)}};
    return s;
  }
  void matps_list(std::vector<Eindex>& o) const override { o = {matps[0][0], matps[0][1], matps[1][0], matps[1][1]}; }
  const char* kind_name() const override { return "R"; }
};

// ---------------------------------------------------------------- Isrc  (comps/mod.rs:325-346)
struct Isrc : Component {
  double i;
  VarIndex p, n;
  Isrc(double i_, VarIndex p_, VarIndex n_) : i(i_), p(p_), n(n_) {}
  void update(double) override {}
  template <class T> void cme(Matrix<T>&) {}
  ORC_CME_BOTH
  Stamps<double> load(const Variables<double>&, const AnalysisInfo&, const Options&) override {
    Stamps<double> s;
    s.b = {{p, i}, {n, -i}};
    return s;
  }
  const char* kind_name() const override { return "I"; }
};

// ---------------------------------------------------------------- Diode  (comps/diode.rs)
struct DiodeModel {  // diode.rs:17-70
  double tnom = 300.15, is = 1e-14, n = 1.0, tt = 0.0, vj = 1.0, m = 0.5, eg = 1.11, xti = 3.0, kf = 0.0, af = 1.0,
         fc = 0.5, bv = 0.0, ibv = 1e-3, rs = 0.0, cj0 = 0.0;
  bool has_rs() const { return rs != 0.0; }
  bool has_bv() const { return bv != 0.0; }
  static DiodeModel from(const std::map<std::string, double>& specs) {
    DiodeModel m;
    auto g = [&](const char* k, double d) { auto it = specs.find(k); return it == specs.end() ? d : it->second; };
    m.tnom = g("tnom", 300.15); m.is = g("is", 1e-14); m.n = g("n", 1.0); m.tt = g("tt", 0.0); m.vj = g("vj", 1.0);
    m.m = g("m", 0.5); m.eg = g("eg", 1.11); m.xti = g("xti", 3.0); m.kf = g("kf", 0.0); m.af = g("af", 1.0);
    m.fc = g("fc", 0.5); m.bv = g("bv", 0.0); m.ibv = g("ibv", 1e-3); m.rs = g("rs", 0.0); m.cj0 = g("cj0", 0.0);
    return m;
  }
};
struct DiodeInstParams { OptF area, temp; };  // spice21.proto:63-68
struct DiodeIntParams {                        // diode.rs:130-212
  double vt, vte, vcrit, isat, gspr, cz, cz2, dep_threshold, f1, f2, f3, bv;
  static DiodeIntParams derive(const DiodeModel& model, const DiodeInstParams& inst, const Options& opts) {
    using namespace consts;
    double tnom = model.tnom;
    double temp = inst.temp.some ? inst.temp.v : opts.temp;
    double area = inst.area.some ? inst.area.v : 1.0;
    double gs = model.has_rs() ? 1.0 / model.rs : 0.0;
    double vt = KB_OVER_Q * temp;
    double vtnom = KB_OVER_Q * tnom;
    double fact2 = temp / TEMP_REF;
    double egfet = 1.16 - (7.02e-4 * temp * temp) / (temp + 1108.0);
    double arg = -egfet / (2.0 * KB * temp) + 1.1150877 / (2.0 * KB * TEMP_REF);
    double pbfact = -2.0 * vt * (1.5 * std::log(fact2) + Q * arg);
    double egfet1 = 1.16 - (7.02e-4 * tnom) / (tnom + 1108.0);  // QUIRK :162 tnom, not tnom^2
    double arg1 = -egfet1 / (KB * 2.0 * tnom) + 1.1150877 / (2.0 * KB * TEMP_REF);
    double fact1 = tnom / TEMP_REF;
    double pbfact1 = -2.0 * vtnom * (1.5 * std::log(fact1) + Q * arg1);
    double pbo = (model.vj - pbfact1) / fact1;
    double gmaold = (model.vj - pbo) / pbo;
    double cjunc = model.cj0 / (1.0 + model.m * (400e-6 * (tnom - TEMP_REF) - gmaold));
    double vjunc = pbfact + fact2 * pbo;
    double gmanew = (vjunc - pbo) / pbo;
    cjunc *= 1.0 + model.m * (400e-6 * (temp - TEMP_REF) - gmanew);
    (void)cjunc;
    double isat = model.is * std::exp(((temp / tnom) - 1.0) * model.eg / model.n * vt + model.xti / model.n * std::log(temp / tnom));
    double xfc = 1.0 - std::log(model.fc);
    double f1 = vjunc * (1.0 - std::exp(1.0 - model.m * xfc)) / (1.0 - model.m);
    double dep_threshold = model.fc * model.vj;
    double vte = model.n * vt;
    double vcrit = vte * (vte / std::sqrt(2.0) / isat);  // QUIRK :179 no ln()
    double bv = model.bv;
    if (model.has_bv()) {
      double ibv = model.ibv;
      for (int i = 0; i < 25; i++) bv = model.bv - vt * std::log(ibv / isat + 1.0 - bv / vt);
    }
    double f2 = std::exp(xfc * (1.0 + model.m));
    double f3 = 1.0 - model.fc * (1.0 + model.m);
    double gspr = gs * area;
    double cz = model.cj0 * area;
    double cz2 = cz / f2;
    return DiodeIntParams{vt, vte, vcrit, isat, gspr, cz, cz2, dep_threshold, f1, f2, f3, bv};
  }
};
struct DiodeOpPoint { double vd = 0, id = 0, gd = 0, cd = 0, charge = 0, capcur = 0, p = 0; };

struct Diode : Component {
  VarIndex p = -1, n = -1, r = -1;
  std::shared_ptr<DiodeModel> model;
  std::shared_ptr<DiodeIntParams> intp;
  Eindex pp = -1, pr = -1, rp = -1, rr = -1, nr = -1, rn = -1, nn = -1;
  DiodeOpPoint op, guess;
  double limit(double vd, bool has_past, double past) const {  // diode.rs:228-245
    double vnew = vd;
    double vold = has_past ? past : guess.vd;
    const DiodeIntParams& ip = *intp;
    if (vnew <= ip.vcrit || std::fabs(vnew - vold) <= 2.0 * ip.vte) return vnew;
    if (vold > 0.0) {
      double arg = 1.0 + (vnew - vold) / ip.vte;
      if (arg > 0.0) return vold + ip.vte * std::log(arg);
      return ip.vcrit;
    }
    return ip.vte * std::log(vnew / ip.vte);
  }
  template <class T> void cme(Matrix<T>& mat) {  // :248-256
    pp = make_matrix_elem(mat, p, p);
    pr = make_matrix_elem(mat, p, r);
    rp = make_matrix_elem(mat, r, p);
    rr = make_matrix_elem(mat, r, r);
    nr = make_matrix_elem(mat, n, r);
    rn = make_matrix_elem(mat, r, n);
    nn = make_matrix_elem(mat, n, n);
  }
  ORC_CME_BOTH
  void commit() override { op = guess; }
  Stamps<double> load(const Variables<double>& g_, const AnalysisInfo& an, const Options& opts) override {  // :279-355
    const DiodeModel& m = *model;
    const DiodeIntParams& ip = *intp;
    double gmin = opts.gmin;
    double vd = g_.get(r) - g_.get(n);
    if (m.has_bv() && vd < std::fmin(10.0 * ip.vte - ip.bv, 0.0)) {
      double vtemp = limit(-ip.bv, true, ip.bv - guess.vd);
      vd = vtemp - ip.bv;
    } else {
      vd = limit(vd, false, 0.0);
    }
    double id, gd;
    if (!m.has_bv() || vd >= -ip.bv) {
      double e = std::exp(vd / ip.vte);
      id = ip.isat * (e - 1.0) + gmin * vd;
      gd = ip.isat * e / ip.vte + gmin;
    } else {
      double e = std::exp((vd - ip.bv) / ip.vte);
      id = -ip.isat * e + gmin * vd;
      gd = ip.isat * e / ip.vte + gmin;
    }
    double qd, cd;
    if (vd < ip.dep_threshold) {
      double a = 1.0 - vd / m.vj;
      double s = -m.m * std::log(a);
      qd = m.tt * m.vj * ip.cz * (1.0 - a * s) / (1.0 - m.m);
      cd = m.tt * gd + ip.cz * s;
    } else {
      qd = m.tt * id + ip.cz * ip.f1 +
           ip.cz2 * (ip.f3 * (vd - ip.dep_threshold) + m.m / 2.0 / m.vj * (vd * vd - ip.dep_threshold * ip.dep_threshold));
      cd = m.tt + ip.cz2 * ip.f3 + m.m * vd / m.vj;
    }
    double gc = 0.0, ic = 0.0, rh = 0.0;
    if (an.kind == AnalysisInfo::TRAN) an.tran->integrate(qd - op.charge, cd, vd, op.capcur, &gc, &ic, &rh);
    id += ic;
    gd += gc;
    guess = DiodeOpPoint{vd, id, gd, cd, qd, ic, vd * id};
    double irhs = id - vd * gd;
    Stamps<double> s;
    s.g = {{nn, gd}, {rn, -gd}, {nr, -gd}, {rr, gd + ip.gspr}, {pp, ip.gspr}, {pr, -ip.gspr}, {rp, -ip.gspr}};
    s.b = {{r, -irhs}, {n, irhs}};
    return s;
  }
  void matps_list(std::vector<Eindex>& o) const override { o = {pp, pr, rp, rr, nr, rn, nn}; }
  const char* kind_name() const override { return "D"; }
};

// Diode0 (diode.rs:359-389) — never constructed by elaboration in the reference; kept for completeness.
struct Diode0 : Component {
  double isat = 0.0, vt = 0.0;
  VarIndex p = -1, n = -1;
  Eindex pp = -1, nn = -1, pn = -1, np = -1;
  template <class T> void cme(Matrix<T>& mat) {
    pp = make_matrix_elem(mat, p, p);
    pn = make_matrix_elem(mat, p, n);
    np = make_matrix_elem(mat, n, p);
    nn = make_matrix_elem(mat, n, n);
  }
  ORC_CME_BOTH
  Stamps<double> load(const Variables<double>& g_, const AnalysisInfo&, const Options&) override {
    double vp = g_.get(p), vn = g_.get(n);
    double vd = std::fmin(std::fmax(vp - vn, -1.5), 1.5);
    double i = isat * (std::exp(vd / vt) - 1.0);
    double gd = (isat / vt) * std::exp(vd / vt);
    double irhs = i - vd * gd;
    Stamps<double> s;
    s.g = {{pp, gd}, {nn, gd}, {pn, -gd}, {np, -gd}};
    s.b = {{p, -irhs}, {n, irhs}};
    return s;
  }
  const char* kind_name() const override { return "D0"; }
};

// ---------------------------------------------------------------- MOS common
enum class MosType { NMOS = 0, PMOS = 1 };
inline double mos_p(MosType t) { return t == MosType::PMOS ? -1.0 : 1.0; }  // mos.rs:99-104

// Optional-valued spec bag: stands in for the prost `Option<f64>` fields of proto::Mos1Model etc.
struct Specs {
  std::map<std::string, double> d;
  OptF get(const char* k) const { auto it = d.find(k); return it == d.end() ? OptF() : OptF(it->second); }
};

// ---------------------------------------------------------------- Mos1 model / params (mos.rs:107-476)
struct Mos1Model {
  MosType mos_type = MosType::NMOS;
  double vt0, kp, gamma, cox_per_area, phi, lambda, cbd, cbs, is, pb, cgso, cgdo, cgbo, cj, mj, cjsw, mjsw, js, tox, ld, fc,
      tnom, kf, af;
  OptF rd, rs, rsh;
  double p() const { return mos_p(mos_type); }
  static Mos1Model resolve(const Specs& specs, int mos_type_i, bool has_tpg, long tpg) {  // :140-237
    using namespace consts;
    Mos1Model m;
    m.mos_type = mos_type_i == 1 ? MosType::PMOS : MosType::NMOS;
    OptF s_tnom = specs.get("tnom");
    double tnom = s_tnom.some ? s_tnom.v + KELVIN_TO_C : TEMP_REF;
    double fact1 = tnom / TEMP_REF;
    double vtnom = tnom * KB_OVER_Q;
    double kt1 = KB * tnom;
    double egfet1 = 1.16 - (7.02e-4 * (tnom * tnom)) / (tnom + 1108.0);
    double arg1 = -egfet1 / 2.0 / kt1 + 1.1150877 / (KB * 2.0 * TEMP_REF);
    double pbfact1 = -2.0 * vtnom * (1.5 * std::log(fact1) + Q * arg1);
    (void)pbfact1;
    double cox_per_area = 0.0;
    double vt0 = specs.get("vt0").or_(0.0);
    double kp = specs.get("kp").or_(2.0e-5);
    double phi = specs.get("phi").or_(0.6);
    double gamma = specs.get("gamma").or_(0.0);
    OptF tox = specs.get("tox");
    if (tox.some) {
      cox_per_area = SIO2_PERMITTIVITY / tox.v;
      if (!specs.get("kp").some) {
        double u0 = specs.get("u0").or_(600.0);
        kp = u0 * cox_per_area * 1e-4;
      }
      OptF nsub = specs.get("nsub");
      if (nsub.some) {
        if (nsub.v * 1e6 <= 1.45e16) throw Panic("Invalid Mos1 Substrate Doping nsub < ni (1.45e16)");
        if (!specs.get("phi").some) {
          phi = 2.0 * vtnom * std::log(nsub.v * 1e6 / 1.45e16);
          phi = std::fmax(phi, 0.1);
        }
        double fermis = m.p() * 0.5 * phi;
        double wkfng = 3.2;
        double gate_type = 1.0;
        if (has_tpg) {
          if (tpg > 1 || tpg < -1) throw Panic("Invalid Mos1 tps");
          gate_type = (double)tpg;
        }
        if (gate_type != 0.0) {
          double fermig = m.p() * gate_type * 0.5 * egfet1;
          wkfng = 3.25 + 0.5 * egfet1 - fermig;
        }
        if (!specs.get("gamma").some) gamma = std::sqrt(2.0 * 11.70 * 8.854214871e-12 * Q * nsub.v * 1e6) / cox_per_area;
        if (!specs.get("vt0").some) {
          double nss = specs.get("nss").or_(0.0);
          double wkfngs = wkfng - (3.25 + 0.5 * egfet1 + fermis);
          double vfb = wkfngs - nss * 1e4 * Q / cox_per_area;
          vt0 = vfb + m.p() * (gamma * std::sqrt(phi) + phi);
        }
      }
    }
    m.vt0 = vt0; m.kp = kp; m.cox_per_area = cox_per_area; m.gamma = gamma; m.phi = phi; m.tnom = tnom;
    m.lambda = specs.get("lambda").or_(0.0);
    m.pb = specs.get("pb").or_(0.8);
    m.cbd = specs.get("cbd").or_(0.0);
    m.cbs = specs.get("cbs").or_(0.0);
    m.cgso = specs.get("cgso").or_(0.0);
    m.cgdo = specs.get("cgdo").or_(0.0);
    m.cgbo = specs.get("cgbo").or_(0.0);
    m.cj = specs.get("cj").or_(0.0);
    m.cjsw = specs.get("cjsw").or_(0.0);
    m.mj = specs.get("mj").or_(0.5);
    m.mjsw = specs.get("mjsw").or_(0.5);
    m.is = specs.get("is").or_(1.0e-14);
    m.js = specs.get("js").or_(1.0e-8);
    m.tox = specs.get("tox").or_(1.0e-7);
    m.ld = specs.get("ld").or_(0.0);
    m.fc = specs.get("fc").or_(0.5);
    m.kf = specs.get("kf").or_(0.0);
    m.af = specs.get("af").or_(1.0);
    m.rd = specs.get("rd");
    m.rs = specs.get("rs");
    m.rsh = specs.get("rsh");
    return m;
  }
};
struct Mos1InstanceParams {  // mos.rs:251-290
  double m, l, w, a_d, a_s, pd, ps, nrd, nrs;
  OptF temp;
  static Mos1InstanceParams resolve(const Specs& s) {
    Mos1InstanceParams p;
    p.m = s.get("m").or_(0.0);  // QUIRK: m defaults to 0.0 and is unused
    p.l = s.get("l").or_(1e-6);
    p.w = s.get("w").or_(1e-6);
    p.a_d = s.get("a_d").or_(1e-12);
    p.a_s = s.get("a_s").or_(1e-12);
    p.pd = s.get("pd").or_(1e-6);
    p.ps = s.get("ps").or_(1e-6);
    p.nrd = s.get("nrd").or_(1.0);
    p.nrs = s.get("nrs").or_(1.0);
    p.temp = s.get("temp");
    return p;
  }
};
struct MosJunction {  // mos.rs:489-521
  double area = 0, isat = 0, depletion_threshold = 0, bulkpot_t = 0, vcrit = 0, czb = 0, czbsw = 0, f2 = 0, f3 = 0, f4 = 0;
  void qc(double v, const Mos1Model& model, double* q, double* c) const {
    if (czb == 0.0 && czbsw == 0.0) { *q = 0.0; *c = 0.0; return; }
    if (v < depletion_threshold) {
      double arg = 1.0 - v / bulkpot_t;
      double sarg = std::exp(-model.mj * std::log(arg));
      double sargsw = std::exp(-model.mjsw * std::log(arg));
      *q = bulkpot_t * (czb * (1.0 - arg * sarg) / (1.0 - model.mj) + czbsw * (1.0 - arg * sargsw) / (1.0 - model.mjsw));
      *c = czb * sarg + czbsw * sargsw;
    } else {
      *q = f4 + v * (f2 + v * f3 / 2.0);
      *c = f2 + v * f3;
    }
  }
};
struct Mos1InternalParams {  // mos.rs:301-476
  double temp, vtherm, vt0_t, kp_t, phi_t, beta, cox, cgs_ov, cgd_ov, cgb_ov, leff, grd, grs;
  MosJunction drain_junc, source_junc;
  static Mos1InternalParams derive(const Mos1Model& model, const Mos1InstanceParams& inst, const Options& opts) {
    using namespace consts;
    if (inst.temp.some) throw Panic("Mos1 Instance Temperatures Are Not Supported");
    double temp = opts.temp;
    double fact1 = model.tnom / TEMP_REF;
    double vtnom = model.tnom * KB_OVER_Q;
    double kt1 = KB * model.tnom;
    double egfet1 = 1.16 - (7.02e-4 * (model.tnom * model.tnom)) / (model.tnom + 1108.0);
    double arg1 = -egfet1 / 2.0 / kt1 + 1.1150877 / (KB * 2.0 * TEMP_REF);
    double pbfact1 = -2.0 * vtnom * (1.5 * std::log(fact1) + Q * arg1);
    double kt = temp * KB;
    double vtherm = temp * KB_OVER_Q;
    double temp_ratio = temp / model.tnom;
    double fact2 = temp / TEMP_REF;
    double egfet = 1.16 - (7.02e-4 * (temp * temp)) / (temp + 1108.0);
    double arg = -egfet / 2.0 / kt + 1.1150877 / (KB * 2.0 * TEMP_REF);
    double pbfact = -2.0 * vtherm * (1.5 * std::log(fact2) + Q * arg);
    double leff = inst.l - 2.0 * model.ld;
    if (leff < 0.0) throw Panic("Mos1 Effective Length < 0");
    double phio = (model.phi - pbfact1) / fact1;
    double phi_t = fact2 * phio + pbfact;
    double vbi_t = model.vt0 - model.p() * (model.gamma * std::sqrt(model.phi)) + 0.5 * (egfet1 - egfet) +
                   model.p() * 0.5 * (phi_t - model.phi);
    double vt0_t = vbi_t + model.p() * model.gamma * std::sqrt(phi_t);
    double isat_t = model.is * std::exp(-egfet / vtherm + egfet1 / vtnom);
    double jsat_t = model.js * std::exp(-egfet / vtherm + egfet1 / vtnom);
    double pbo = (model.pb - pbfact1) / fact1;
    double gmaold = (model.pb - pbo) / pbo;
    double capfact = 1.0 / (1.0 + model.mj * (4e-4 * (model.tnom - TEMP_REF) - gmaold));
    double cbd_t = model.cbd * capfact;
    double cbs_t = model.cbs * capfact;
    double cj_t = model.cj * capfact;
    capfact = 1.0 / (1.0 + model.mjsw * (4e-4 * (model.tnom - TEMP_REF) - gmaold));
    double cjsw_t = model.cjsw * capfact;
    double bulkpot_t = fact2 * pbo + pbfact;
    double gmanew = (bulkpot_t - pbo) / pbo;
    capfact = 1.0 / (1.0 + model.mj * (4e-4 * (temp - TEMP_REF) - gmanew));
    cbd_t *= capfact;
    cbs_t *= capfact;
    cj_t *= capfact;
    capfact = 1.0 / (1.0 + model.mjsw * (4e-4 * (temp - TEMP_REF) - gmanew));
    cjsw_t *= capfact;
    double depletion_threshold = model.fc * bulkpot_t;
    double arg_ = 1.0 - model.fc;
    double sarg = std::exp((-model.mj) * std::log(arg_));
    double sargsw = std::exp((-model.mjsw) * std::log(arg_));
    bool use_default_isat = jsat_t == 0.0 || inst.a_d == 0.0 || inst.a_s == 0.0;
    auto junc_new = [&](double area, double perim, bool is_drain) {
      MosJunction j;
      double isat = use_default_isat ? isat_t : jsat_t * area;
      double vcrit = vtherm * std::log(vtherm / (SQRT2 * isat));
      double czb;
      if (is_drain) czb = (model.cbd == 0.0) ? cj_t * area : cbd_t;
      else czb = (model.cbs == 0.0) ? cj_t * area : cbs_t;
      double czbsw = cjsw_t * perim;
      double f2 = czb * (1.0 - model.fc * (1.0 + model.mj)) * sarg / arg_ + czbsw * (1.0 - model.fc * (1.0 + model.mjsw)) * sargsw / arg_;
      double f3 = czb * model.mj * sarg / arg_ / bulkpot_t + czbsw * model.mjsw * sargsw / arg_ / bulkpot_t;
      double f4 = czb * bulkpot_t * (1.0 - arg_ * sarg) / (1.0 - model.mj) + czbsw * bulkpot_t * (1.0 - arg_ * sargsw) / (1.0 - model.mjsw) -
                  f3 / 2.0 * (depletion_threshold * depletion_threshold) - depletion_threshold * f2;
      j.area = area; j.isat = isat; j.depletion_threshold = depletion_threshold; j.bulkpot_t = bulkpot_t; j.vcrit = vcrit;
      j.czb = czb; j.czbsw = czbsw; j.f2 = f2; j.f3 = f3; j.f4 = f4;
      return j;
    };
    Mos1InternalParams ip;
    ip.drain_junc = junc_new(inst.a_d, inst.pd, true);
    ip.source_junc = junc_new(inst.a_s, inst.ps, false);
    double grs;
    if (model.rs.some) grs = model.rs.v <= 0.0 ? 0.0 : 1.0 / model.rs.v;
    else if (model.rsh.some) grs = model.rsh.v <= 0.0 ? 0.0 : 1.0 / model.rsh.v / inst.nrs;
    else grs = 0.0;
    double grd;
    if (model.rd.some) grd = model.rd.v <= 0.0 ? 0.0 : 1.0 / model.rd.v;
    else if (model.rsh.some) grd = model.rsh.v <= 0.0 ? 0.0 : 1.0 / model.rsh.v / inst.nrd;
    else grd = 0.0;
    double kp_t = model.kp / temp_ratio * std::sqrt(temp_ratio);
    ip.vt0_t = vt0_t; ip.kp_t = kp_t; ip.temp = temp; ip.vtherm = vtherm; ip.leff = leff;
    ip.cox = model.cox_per_area * leff * inst.w;
    ip.beta = kp_t * inst.w / leff;
    ip.phi_t = phi_t;
    ip.cgs_ov = inst.w * model.cgso;
    ip.cgd_ov = inst.w * model.cgdo;
    ip.cgb_ov = leff * model.cgbo;
    ip.grs = grs; ip.grd = grd;
    return ip;
  }
};

struct Mos1TranState { ChargeInteg gs, gd, gb, bs, bd; };
struct Mos1OpPoint {  // mos.rs:526-546
  double ids = 0, vgs = 0, vds = 0, vgd = 0, vgb = 0, vdb = 0, vsb = 0, gm = 0, gds = 0, gmbs = 0, gbs = 0, gbd = 0, cgs = 0, cgd = 0,
         cgb = 0, cbs = 0, cbd = 0;
  bool reversed = false;
  Mos1TranState tr;
};
enum Mos1Var { M1_D = 0, M1_G = 1, M1_S = 2, M1_B = 3, M1_DP = 4, M1_SP = 5 };  // mos.rs:570-577

struct Mos1 : Component {
  std::shared_ptr<Mos1Model> model;
  std::shared_ptr<Mos1InternalParams> intparams;
  VarIndex ports[6] = {-1, -1, -1, -1, -1, -1};  // indexed by Mos1Var
  Mos1OpPoint op, guess;
  Eindex matps[6][6];
  Mos1() { for (auto& r : matps) for (auto& e : r) e = -1; }

  // mos.rs:649-893
  Stamps<double> op_stamp(const double v[6], const AnalysisInfo& an, const Options& opts, Mos1OpPoint* out) const {
    const Mos1Model& model_ = *model;
    const Mos1InternalParams& intp = *intparams;
    double gmin = opts.gmin;
    double p = model_.p();
    bool reversed = p * (v[M1_D] - v[M1_S]) < 0.0;
    double vd = reversed ? v[M1_S] : v[M1_D];
    double vs = reversed ? v[M1_D] : v[M1_S];
    double vgs = p * (v[M1_G] - vs);
    double vgd = p * (v[M1_G] - vd);
    double vds = p * (vd - vs);
    double vgb = p * (v[M1_G] - v[M1_B]);
    double vsb = p * (vs - v[M1_B]);
    double vdb = p * (vd - v[M1_B]);
    double von = vsb > 0.0 ? intp.vt0_t + model_.gamma * (std::sqrt(intp.phi_t + vsb) - std::sqrt(intp.phi_t)) : intp.vt0_t;
    double vov = vgs - von;
    double vdsat = std::fmax(vov, 0.0);
    double ids = 0.0, gm = 0.0, gds = 0.0, gmbs = 0.0;
    if (vov > 0.0) {
      if (vds >= vov) {
        ids = intp.beta / 2.0 * (vov * vov) * (1.0 + model_.lambda * vds);
        gm = intp.beta * vov * (1.0 + model_.lambda * vds);
        gds = model_.lambda * intp.beta / 2.0 * (vov * vov);
      } else {
        ids = intp.beta * (vov * vds - (vds * vds) / 2.0) * (1.0 + model_.lambda * vds);
        gm = intp.beta * vds * (1.0 + model_.lambda * vds);
        gds = intp.beta * ((vov - vds) * (1.0 + model_.lambda * vds) + model_.lambda * ((vov * vds) - (vds * vds) / 2.0));
      }
      gmbs = (intp.phi_t + vsb > 0.0) ? gm * model_.gamma / 2.0 / std::sqrt(intp.phi_t + vsb) : 0.0;
    }
    double vtherm = intp.vtherm;
    const MosJunction& bs_junc = !reversed ? intp.source_junc : intp.drain_junc;
    const MosJunction& bd_junc = !reversed ? intp.drain_junc : intp.source_junc;
    double ibs = bs_junc.isat * (std::exp(-vsb / vtherm) - 1.0);
    double gbs = (bs_junc.isat / vtherm) * std::exp(-vsb / vtherm) + gmin;
    double ibs_rhs = ibs + vsb * gbs;
    double ibd = bd_junc.isat * (std::exp(-vdb / vtherm) - 1.0);
    double gbd = (bd_junc.isat / vtherm) * std::exp(-vdb / vtherm) + gmin;
    double ibd_rhs = ibd + vdb * gbd;
    double cox = intp.cox;
    double cgs1, cgd1, cgb1;
    if (vov <= -intp.phi_t) {
      cgb1 = cox / 2.0; cgs1 = 0.0; cgd1 = 0.0;
    } else if (vov <= -intp.phi_t / 2.0) {
      cgb1 = -vov * cox / (2.0 * intp.phi_t); cgs1 = 0.0; cgd1 = 0.0;
    } else if (vov <= 0.0) {
      cgb1 = -vov * cox / (2.0 * intp.phi_t);
      cgs1 = vov * cox / (1.5 * intp.phi_t) + cox / 3.0;
      cgd1 = 0.0;
    } else if (vdsat <= vds) {
      cgs1 = cox / 3.0; cgd1 = 0.0; cgb1 = 0.0;
    } else {
      double vddif = 2.0 * vdsat - vds;
      double vddif1 = vdsat - vds;
      double vddif2 = vddif * vddif;
      cgd1 = cox * (1.0 - vdsat * vdsat / vddif2) / 3.0;
      cgs1 = cox * (1.0 - vddif1 * vddif1 / vddif2) / 3.0;
      cgb1 = 0.0;
    }
    double cgs2 = (op.cgs == 0.0) ? cgs1 : (reversed == op.reversed ? op.cgs : op.cgd);
    double cgs = cgs1 + cgs2 + intp.cgs_ov;
    double cgd = cgd1 + intp.cgd_ov + (reversed == op.reversed ? op.cgd : op.cgs);
    double cgb = cgb1 + intp.cgb_ov + op.cgb;
    double qbs_, cbs, qbd_, cbd;
    bs_junc.qc(-vsb, model_, &qbs_, &cbs);
    bd_junc.qc(-vdb, model_, &qbd_, &cbd);
    Mos1TranState tr;
    if (an.kind == AnalysisInfo::TRAN) {
      const TranState& st = *an.tran;
      {
        double dqgs = (reversed == op.reversed) ? (vgs - op.vgs) * cgs : (vgs - op.vgd) * cgs;
        double ip = (reversed == op.reversed) ? op.tr.gs.i : op.tr.gd.i;
        tr.gs = integq(st, dqgs, cgs, vgs, ip);
      }
      {
        double dqgd = (reversed == op.reversed) ? (vgd - op.vgd) * cgd : (vgd - op.vgs) * cgd;
        double ip = (reversed == op.reversed) ? op.tr.gd.i : op.tr.gs.i;
        tr.gs = integq(st, dqgd, cgd, vgd, ip);  // QUIRK mos.rs:801 — assigned to tr.gs again; tr.gd stays 0
      }
      {
        double dqgb = (vgb - op.vgb) * cgb;
        tr.gb = integq(st, dqgb, cgb, vgb, op.tr.gb.i);
      }
      {
        double dqbs = (reversed == op.reversed) ? (-vsb + op.vsb) * cbs : (-vsb + op.vdb) * cbs;
        double dqbd = (reversed == op.reversed) ? (-vdb + op.vdb) * cbd : (-vdb + op.vsb) * cbd;
        double isp = (reversed == op.reversed) ? op.tr.gs.i : op.tr.gd.i;
        double idp = (reversed == op.reversed) ? op.tr.gd.i : op.tr.gs.i;
        tr.bs = integq(st, dqbs, cbs, -vsb, isp);
        tr.bd = integq(st, dqbd, cbd, -vdb, idp);
      }
    }
    double irhs = ids - gm * vgs - gds * vds;
    Mos1Var sr, sx, dr, dx;
    if (!reversed) { sr = M1_SP; sx = M1_S; dr = M1_DP; dx = M1_D; }
    else { sr = M1_DP; sx = M1_D; dr = M1_SP; dx = M1_S; }
    double grd = intp.grd, grs = intp.grs;
    const int G = M1_G, B = M1_B;
    Stamps<double> s;
    s.g = {
        {matps[dr][dr], gds + grd + gbd + tr.gd.g},
        {matps[sr][sr], gm + gds + grs + gbs + gmbs + tr.gs.g},
        {matps[dr][sr], -gm - gds - gmbs},
        {matps[sr][dr], -gds},
        {matps[dr][G], gm - tr.gd.g},
        {matps[sr][G], -gm - tr.gs.g},
        {matps[G][G], (tr.gd.g + tr.gs.g + tr.gb.g)},
        {matps[B][B], (gbd + gbs + tr.gb.g)},
        {matps[G][B], -tr.gb.g},
        {matps[G][dr], -tr.gd.g},
        {matps[G][sr], -tr.gs.g},
        {matps[B][G], -tr.gb.g},
        {matps[B][dr], -gbd},
        {matps[B][sr], -gbs},
        {matps[dr][B], -gbd + gmbs},
        {matps[sr][B], -gbs - gmbs},
        {matps[dx][dr], -grd},
        {matps[dr][dx], -grd},
        {matps[dx][dx], grd},
        {matps[sx][sr], -grs},
        {matps[sr][sx], -grs},
        {matps[sx][sx], grs},
    };
    s.b = {
        {ports[dr], p * (-irhs + ibd_rhs + tr.gd.rhs)},
        {ports[sr], p * (irhs + ibs_rhs + tr.gs.rhs)},
        {ports[G], -p * (tr.gs.rhs + tr.gb.rhs + tr.gd.rhs)},
        {ports[B], -p * (ibd_rhs + ibs_rhs - tr.gb.rhs)},
    };
    Mos1OpPoint g;
    g.ids = ids; g.vgs = vgs; g.vds = vds; g.vgd = vgd; g.vgb = vgb; g.vdb = vdb; g.vsb = vsb; g.gm = gm; g.gds = gds;
    g.gmbs = gmbs; g.gbs = gbs; g.gbd = gbd; g.reversed = reversed; g.cgs = cgs1; g.cgd = cgd1; g.cgb = cgb1;
    g.cbs = cbs; g.cbd = cbd; g.tr = tr;
    *out = g;
    return s;
  }
  template <class T> void cme(Matrix<T>& mat) {  // mos.rs:896-903
    const Mos1Var order[6] = {M1_G, M1_D, M1_S, M1_B, M1_DP, M1_SP};
    for (Mos1Var t1 : order)
      for (Mos1Var t2 : order) matps[t1][t2] = make_matrix_elem(mat, ports[t1], ports[t2]);
  }
  ORC_CME_BOTH
  void commit() override { op = guess; }
  Stamps<double> load(const Variables<double>& vars, const AnalysisInfo& an, const Options& opts) override {  // :908-913
    double v[6];
    for (int k = 0; k < 6; k++) v[k] = vars.get(ports[k]);
    Mos1OpPoint g;
    Stamps<double> s = op_stamp(v, an, opts, &g);
    guess = g;
    return s;
  }
  Stamps<Cplx> load_ac(const Variables<Cplx>&, const AnalysisInfo& an, const Options&) override {  // :914-968
    const Mos1InternalParams& intp = *intparams;
    if (an.kind != AnalysisInfo::AC) throw Panic("Invalid AC AnalysisInfo");
    double omega = an.ac->omega;
    double gm = op.gm, gds = op.gds, gmbs = op.gmbs, gbs = op.gbs, gbd = op.gbd;
    double gcgs = omega * op.cgs;
    double gcgd = omega * op.cgd;
    double gcgb = omega * op.cgb;
    Mos1Var sr, sx, dr, dx;
    if (!op.reversed) { sr = M1_SP; sx = M1_S; dr = M1_DP; dx = M1_D; }
    else { sr = M1_DP; sx = M1_D; dr = M1_SP; dx = M1_S; }
    const int G = M1_G, B = M1_B;
    Stamps<Cplx> s;
    s.g = {
        {matps[dr][dr], Cplx(gds + intp.grd + gbd, gcgd)},
        {matps[sr][sr], Cplx(gm + gds + intp.grs + gbs + gmbs, gcgs)},
        {matps[dr][sr], Cplx(-gm - gds - gmbs, 0.0)},
        {matps[sr][dr], Cplx(-gds, 0.0)},
        {matps[dr][G], Cplx(gm, -gcgd)},
        {matps[sr][G], Cplx(-gm, -gcgs)},
        {matps[G][G], Cplx(0.0, gcgd + gcgs + gcgb)},
        {matps[B][B], Cplx(gbd + gbs, gcgb)},
        {matps[G][B], Cplx(0.0, -gcgb)},
        {matps[G][dr], Cplx(0.0, -gcgd)},
        {matps[G][sr], Cplx(0.0, -gcgs)},
        {matps[B][G], Cplx(0.0, -gcgb)},
        {matps[G][dr], Cplx(0.0, -gcgd)},  // QUIRK mos.rs:954 — (G,dr) pushed twice
        {matps[B][dr], Cplx(-gbd, 0.0)},
        {matps[B][sr], Cplx(-gbs, 0.0)},
        {matps[dr][B], Cplx(-gbd + gmbs, 0.0)},
        {matps[sr][B], Cplx(-gbs - gmbs, 0.0)},
        {matps[dx][dr], Cplx(-intp.grd, 0.0)},
        {matps[dr][dx], Cplx(-intp.grd, 0.0)},
        {matps[dx][dx], Cplx(intp.grd, 0.0)},
        {matps[sx][sr], Cplx(-intp.grs, 0.0)},
        {matps[sr][sx], Cplx(-intp.grs, 0.0)},
        {matps[sx][sx], Cplx(intp.grs, 0.0)},
    };
    return s;
  }
  void matps_list(std::vector<Eindex>& o) const override {
    const Mos1Var order[6] = {M1_G, M1_D, M1_S, M1_B, M1_DP, M1_SP};
    o.clear();
    for (Mos1Var t1 : order) for (Mos1Var t2 : order) o.push_back(matps[t1][t2]);
  }
  const char* kind_name() const override { return "M1"; }
};

// ---------------------------------------------------------------- Mos0 (mos.rs:1008-1099)
enum MosTerm { MT_D = 0, MT_G = 1, MT_S = 2, MT_B = 3 };
struct Mos0 : Component {
  MosType mos_type;
  double vth = 0.25, beta = 50e-3, lam = 3e-3;
  VarIndex ports[4];  // d g s b
  Eindex matps[4][4];
  Mos0(const VarIndex p[4], MosType t) : mos_type(t) {
    for (int k = 0; k < 4; k++) ports[k] = p[k];
    for (auto& r : matps) for (auto& e : r) e = -1;
  }
  template <class T> void cme(Matrix<T>& mat) {
    const int pr[6][2] = {{MT_D, MT_D}, {MT_S, MT_S}, {MT_D, MT_S}, {MT_S, MT_D}, {MT_D, MT_G}, {MT_S, MT_G}};
    for (auto& t : pr) matps[t[0]][t[1]] = make_matrix_elem(mat, ports[t[0]], ports[t[1]]);
  }
  ORC_CME_BOTH
  Stamps<double> load(const Variables<double>& guess, const AnalysisInfo&, const Options& opts) override {
    double gmin = opts.gmin;
    double vg = guess.get(ports[MT_G]);
    double vd = guess.get(ports[MT_D]);
    double vs = guess.get(ports[MT_S]);
    double p = mos_p(mos_type);
    double vds1 = p * (vd - vs);
    bool reversed = vds1 < 0.0;
    double vgs = reversed ? p * (vg - vd) : p * (vg - vs);
    double vds = reversed ? -vds1 : vds1;
    double vov = vgs - vth;
    double ids = 0.0, gm = 0.0, gds = 0.0;
    if (vov > 0.0) {
      if (vds >= vov) {
        ids = beta / 2.0 * (vov * vov) * (1.0 + lam * vds);
        gm = beta * vov * (1.0 + lam * vds);
        gds = lam * beta / 2.0 * (vov * vov);
      } else {
        ids = beta * (vov * vds - (vds * vds) / 2.0) * (1.0 + lam * vds);
        gm = beta * vds * (1.0 + lam * vds);
        gds = beta * ((vov - vds) * (1.0 + lam * vds) + lam * ((vov * vds) - (vds * vds) / 2.0));
      }
    }
    int sr = reversed ? MT_D : MT_S, dr = reversed ? MT_S : MT_D;
    double irhs = ids - gm * vgs - gds * vds;
    Stamps<double> s;
    s.g = {{matps[dr][dr], gds + gmin}, {matps[sr][sr], (gm + gds + gmin)}, {matps[dr][sr], -(gm + gds + gmin)},
           {matps[sr][dr], -gds - gmin}, {matps[dr][MT_G], gm}, {matps[sr][MT_G], -gm}};
    s.b = {{ports[dr], -p * irhs}, {ports[sr], p * irhs}};
    return s;
  }
  void matps_list(std::vector<Eindex>& o) const override {
    o = {matps[MT_D][MT_D], matps[MT_S][MT_S], matps[MT_D][MT_S], matps[MT_S][MT_D], matps[MT_D][MT_G], matps[MT_S][MT_G]};
  }
  const char* kind_name() const override { return "M0"; }
};

}  // namespace orc
