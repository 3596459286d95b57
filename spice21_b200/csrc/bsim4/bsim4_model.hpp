// BSIM4 model cards on the host: parameter defaulting / clamping and the temperature-derived model quantities that the
// per-instance precompute (bsim4_size.hpp) and the device evaluation (bsim4_eval.cuh) read.
//
// Behaviour follows spice21/src/comps/bsim4/model/vals.rs:9-1376 (`resolve`) and bsim4derive.rs:7-202 (`derive`),
// including the reference's simplifications: the simulator temperature is fixed at 300.15 K (bsim4derive.rs:38), tnom
// given in Celsius, toxe must equal toxp + dtox exactly (vals.rs:1209-1211). The ~850 parameter names and literal
// defaults are a table (bsim4_model_table.inc) expanded here; everything conditional is written out below.
// Host only; compiled with -ffp-contract=off.
#pragma once
#include <cmath>
#include <map>
#include <set>
#include <stdexcept>
#include <string>

namespace s21 {
namespace b4 {

// comps/mod.rs:24-37 (the BSIM4 code uses EPS0 / EPSSI / Q / KB_OVER_Q / VT_REF / SQRT2)
constexpr double KB = 1.3806226e-23;
constexpr double QE = 1.6021918e-19;
constexpr double KB_OVER_Q = KB / QE;
constexpr double TEMP_REF = 273.15 + 27.0;
constexpr double VT_REF = KB * TEMP_REF / QE;
constexpr double SQRT2 = 1.4142135624;
constexpr double EPS0 = 8.85418e-12;
constexpr double EPSSI = 1.03594e-10;
constexpr double PI = 3.14159265358979323846264338327950288;

struct ModelError : std::runtime_error {
  explicit ModelError(const std::string& m) : std::runtime_error(m) {}
};

// Every model parameter as a double (selectors included), plus the "given" information downstream code asks for.
struct Model {
  int mos_type = 0;  // 0 NMOS, 1 PMOS
  double tnom = 300.15;
#define B4P(name, dflt) double name = 0.0;
#define B4A(name, other) double name = 0.0;
#define B4T(name, n, p) double name = 0.0;
#include "bsim4_model_table.inc"
#undef B4P
#undef B4A
#undef B4T
  double ua = 0, uc = 0, uc1 = 0, cf = 0, cgso = 0, cgdo = 0, cgbo = 0;
  double aigs = 0, aigd = 0, bigs = 0, bigd = 0, cigs = 0, cigd = 0, aigsd = 0, bigsd = 0, cigsd = 0;
  int bodymode = 1;
  std::set<std::string> given;
  bool has(const char* k) const { return given.count(k) != 0; }
  double p() const { return mos_type == 1 ? -1.0 : 1.0; }
};

inline Model resolve_model(int mos_type, const std::map<std::string, double>& specs) {
  Model v;
  v.mos_type = mos_type == 1 ? 1 : 0;
  for (auto& kv : specs) v.given.insert(kv.first);
  auto sp = [&](const char* k, double* out) {
    auto it = specs.find(k);
    if (it == specs.end()) return false;
    *out = it->second;
    return true;
  };
  double t;
  v.tnom = sp("tnom", &t) ? t + 273.15 : 300.15;
  const bool nmos = v.mos_type == 0;
  // table-driven defaults, in the reference's order (an alias reads a value resolved earlier)
#define B4P(name, dflt) if (!sp(#name, &v.name)) v.name = dflt;
#define B4A(name, other) if (!sp(#name, &v.name)) v.name = v.other;
#define B4T(name, n, p) if (!sp(#name, &v.name)) v.name = nmos ? (n) : (p);
#include "bsim4_model_table.inc"
#undef B4P
#undef B4A
#undef B4T
  // selector range checks happen right after each selector in the reference; none of the later defaults reads a selector
  // before its check except ua / uc / uc1 (mobmod), resolved below with the checked value
  auto sel = [](double& x, double hi, double dflt) { if (x > hi) x = dflt; };
  sel(v.mobmod, 6, 0); sel(v.diomod, 2, 1); sel(v.capmod, 2, 2); sel(v.rdsmod, 1, 0); sel(v.rbodymod, 2, 0); sel(v.rgatemod, 3, 0);
  sel(v.permod, 1, 1); sel(v.fnoimod, 1, 1); sel(v.tnoimod, 2, 0); sel(v.trnqsmod, 1, 0); sel(v.acnqsmod, 1, 0); sel(v.mtrlmod, 1, 0);
  sel(v.mtrlcompatmod, 1, 0); sel(v.igcmod, 2, 0); sel(v.igbmod, 1, 0); sel(v.tempmod, 3, 0); sel(v.wpemod, 1, 0);
  if (!sp("ua", &v.ua)) v.ua = v.mobmod == 2 ? 1.0e-15 : 1.0e-9;
  if (!sp("uc", &v.uc)) v.uc = v.mobmod == 1 ? -0.0465 : -0.0465e-9;
  if (!sp("uc1", &v.uc1)) v.uc1 = v.mobmod == 1 ? -0.056 : -0.056e-9;
  // gate-tunnelling S/D groups: a combined "xigsd" overrides both sides (vals.rs "aigsd"/"bigsd"/"cigsd" blocks)
  const double a0 = nmos ? 1.36e-2 : 9.80e-3, b0 = nmos ? 1.71e-3 : 7.59e-4, c0 = nmos ? 0.075 : 0.03;
  if (sp("aigsd", &t)) { v.aigs = t; v.aigd = t; } else { v.aigsd = a0; if (!sp("aigs", &v.aigs)) v.aigs = a0; if (!sp("aigd", &v.aigd)) v.aigd = a0; }
  if (sp("bigsd", &t)) { v.bigs = t; v.bigd = t; } else { v.bigsd = b0; if (!sp("bigs", &v.bigs)) v.bigs = b0; if (!sp("bigd", &v.bigd)) v.bigd = b0; }
  if (sp("cigsd", &t)) { v.cigs = t; v.cigd = t; } else { v.cigsd = c0; if (!sp("cigs", &v.cigs)) v.cigs = c0; if (!sp("cigd", &v.cigd)) v.cigd = c0; }
  const bool a_sep = !v.has("aigsd") && (v.has("aigs") || v.has("aigd"));
  const bool b_sep = !v.has("bigsd") && (v.has("bigs") || v.has("bigd"));
  const bool c_sep = !v.has("cigsd") && (v.has("cigs") || v.has("cigd"));
  if (!a_sep) { v.laigs = v.laigd = v.laigsd; v.waigs = v.waigd = v.waigsd; v.paigs = v.paigd = v.paigsd; }
  if (!b_sep) { v.lbigs = v.lbigd = v.lbigsd; v.wbigs = v.wbigd = v.wbigsd; v.pbigs = v.pbigd = v.pbigsd; }
  if (!c_sep) { v.lcigs = v.lcigd = v.lcigsd; v.wcigs = v.wcigd = v.wcigsd; v.pcigs = v.pcigd = v.pcigsd; }
  if (!sp("cf", &v.cf)) v.cf = 2.0 * v.epsrox * EPS0 / PI * std::log(1.0 + 0.4e-6 / v.toxe);
  if (v.toxe != v.toxp + v.dtox) throw ModelError("Invalid toxe, toxp and dtox params");
  const double coxe = v.epsrox * EPS0 / v.toxe;
  if (!sp("cgso", &v.cgso)) v.cgso = (v.has("dlc") && v.dlc > 0.0) ? v.dlc * coxe - v.cgsl : 0.6 * v.xj * coxe;
  if (!sp("cgdo", &v.cgdo)) v.cgdo = (v.has("dlc") && v.dlc > 0.0) ? v.dlc * coxe - v.cgdl : 0.6 * v.xj * coxe;
  if (!sp("cgbo", &v.cgbo)) v.cgbo = 2.0 * v.dwc * coxe;
  // range limiting
  auto floor_to = [](double& x, double lo) { if (x < lo) x = lo; };
  floor_to(v.pbs, 0.1); floor_to(v.pbsws, 0.1); floor_to(v.pbswgs, 0.1); floor_to(v.pbd, 0.1); floor_to(v.pbswd, 0.1); floor_to(v.pbswgd, 0.1);
  auto nonpos_to_zero = [](double& x) { if (x <= 0.0) x = 0.0; };
  nonpos_to_zero(v.ijthdfwd); nonpos_to_zero(v.ijthsfwd); nonpos_to_zero(v.ijthdrev); nonpos_to_zero(v.ijthsrev);
  if (v.diomod == 2 || v.diomod == 0) { nonpos_to_zero(v.xjbvd); nonpos_to_zero(v.xjbvs); }
  nonpos_to_zero(v.bvd); nonpos_to_zero(v.bvs);
  floor_to(v.jtweff, 0.0); floor_to(v.cjsws, 0.0); floor_to(v.cjswd, 0.0); floor_to(v.wlod, 0.0);
  if (!v.has("rbps0") || !v.has("rbpd0")) v.bodymode = 1;
  else if ((!v.has("rbsbx0") && !v.has("rbsby0")) || (!v.has("rbdbx0") && !v.has("rbdby0"))) v.bodymode = 3;
  else v.bodymode = 5;
  return v;
}

// bsim4derive.rs:7-202 — quantities shared by every instance of a model
struct ModelDerived {
  double coxp = 0, Eg0 = 0, vtm = 0, vtm0 = 0, coxe = 0, vcrit = 0, factor1 = 0, PhiBS = 0, PhiBSWS = 0, PhiBSWGS = 0;
  double SjctTempSatCurDensity = 0, SjctSidewallTempSatCurDensity = 0, SjctGateSidewallTempSatCurDensity = 0;
  double PhiBD = 0, PhiBSWD = 0, PhiBSWGD = 0;
  double DjctTempSatCurDensity = 0, DjctSidewallTempSatCurDensity = 0, DjctGateSidewallTempSatCurDensity = 0;
  double SunitAreaTempJctCap = 0, DunitAreaTempJctCap = 0, SunitLengthSidewallTempJctCap = 0, DunitLengthSidewallTempJctCap = 0;
  double SunitLengthGateSidewallTempJctCap = 0, DunitLengthGateSidewallTempJctCap = 0;
  double njtsstemp = 0, njtsswstemp = 0, njtsswgstemp = 0, njtsdtemp = 0, njtsswdtemp = 0, njtsswgdtemp = 0;
  double TempRatio = 0, epssub = 0, ni = 0;
  double Nvtms = 0, Nvtmd = 0, Nvtmrss = 0, Nvtmrssws = 0, Nvtmrsswgs = 0, Nvtmrsd = 0, Nvtmrsswd = 0, Nvtmrsswgd = 0;
};

constexpr double B4_TEMP = 300.15;  // the reference ignores the simulator temperature for BSIM4 (bsim4derive.rs:38)

inline ModelDerived derive_model(const Model& m) {
  ModelDerived d;
  d.epssub = m.mtrlmod != 0 ? EPS0 * m.epsrsub : EPSSI;
  d.coxp = (m.mtrlmod == 0 || m.mtrlcompatmod != 0) ? m.epsrox * EPS0 / m.toxp : 0.0;
  d.coxe = m.epsrox * EPS0 / m.toxe;
  d.vcrit = VT_REF * std::log(VT_REF / (SQRT2 * 1.0e-14));
  d.factor1 = std::sqrt(d.epssub / (m.epsrox * EPS0) * m.toxe);
  const double Temp = B4_TEMP, Tnom = m.tnom;
  d.TempRatio = Temp / Tnom;
  const double Vtm0 = KB_OVER_Q * Tnom;
  d.vtm0 = Vtm0;
  double Eg0, ni, Eg;
  if (m.mtrlmod == 0) {
    Eg0 = 1.16 - 7.02e-4 * Tnom * Tnom / (Tnom + 1108.0);
    ni = 1.45e10 * (Tnom / 300.15) * std::sqrt(Tnom / 300.15) * std::exp(21.5565981 - Eg0 / (2.0 * Vtm0));
  } else {
    Eg0 = m.bg0sub - m.tbgasub * Tnom * Tnom / (Tnom + m.tbgbsub);
    const double T0 = m.bg0sub - m.tbgasub * 90090.0225 / (300.15 + m.tbgbsub);
    ni = m.ni0sub * (Tnom / 300.15) * std::sqrt(Tnom / 300.15) * std::exp((T0 - Eg0) / (2.0 * Vtm0));
  }
  d.Eg0 = Eg0;
  d.vtm = KB_OVER_Q * Temp;
  d.ni = ni;
  Eg = m.mtrlmod == 0 ? 1.16 - 7.02e-4 * Temp * Temp / (Temp + 1108.0) : m.bg0sub - m.tbgasub * Temp * Temp / (Temp + m.tbgbsub);
  if (Temp != Tnom) {
    const double T0 = Eg0 / Vtm0 - Eg / d.vtm;
    const double T1 = std::log(Temp / Tnom);
    double T3 = std::exp((T0 + m.xtis * T1) / m.njs);
    d.SjctTempSatCurDensity = m.jss * T3;
    d.SjctSidewallTempSatCurDensity = m.jsws * T3;
    d.SjctGateSidewallTempSatCurDensity = m.jswgs * T3;
    T3 = std::exp((T0 + m.xtid * T1) / m.njd);
    d.DjctTempSatCurDensity = m.jsd * T3;
    d.DjctSidewallTempSatCurDensity = m.jswd * T3;
    d.DjctGateSidewallTempSatCurDensity = m.jswgd * T3;
  } else {
    d.SjctTempSatCurDensity = m.jss; d.SjctSidewallTempSatCurDensity = m.jsws; d.SjctGateSidewallTempSatCurDensity = m.jswgs;
    d.DjctTempSatCurDensity = m.jsd; d.DjctSidewallTempSatCurDensity = m.jswd; d.DjctGateSidewallTempSatCurDensity = m.jswgd;
  }
  for (double* x : {&d.SjctTempSatCurDensity, &d.SjctSidewallTempSatCurDensity, &d.SjctGateSidewallTempSatCurDensity,
                    &d.DjctTempSatCurDensity, &d.DjctSidewallTempSatCurDensity, &d.DjctGateSidewallTempSatCurDensity})
    if (*x < 0.0) *x = 0.0;
  // junction capacitance temperature dependence (a negative factor leaves the field at its zero default)
  const double delTemp = Temp - m.tnom;
  double T0 = m.tcj * delTemp;
  if (T0 >= -1.0) { d.SunitAreaTempJctCap = m.cjs * (1.0 + T0); d.DunitAreaTempJctCap = m.cjd * (1.0 + T0); }
  T0 = m.tcjsw * delTemp;
  if (T0 >= -1.0) { d.SunitLengthSidewallTempJctCap = m.cjsws * (1.0 + T0); d.DunitLengthSidewallTempJctCap = m.cjswd * (1.0 + T0); }
  T0 = m.tcjswg * delTemp;
  if (T0 >= -1.0) { d.SunitLengthGateSidewallTempJctCap = m.cjswgs * (1.0 + T0); d.DunitLengthGateSidewallTempJctCap = m.cjswgd * (1.0 + T0); }
  d.PhiBS = m.pbs - m.tpb * delTemp;       if (d.PhiBS < 0.01) d.PhiBS = 0.01;
  d.PhiBD = m.pbd - m.tpb * delTemp;       if (d.PhiBD < 0.01) d.PhiBD = 0.01;
  d.PhiBSWS = m.pbsws - m.tpbsw * delTemp; if (d.PhiBSWS <= 0.01) d.PhiBSWS = 0.01;
  d.PhiBSWD = m.pbswd - m.tpbsw * delTemp; if (d.PhiBSWD <= 0.01) d.PhiBSWD = 0.01;
  d.PhiBSWGS = m.pbswgs - m.tpbswg * delTemp; if (d.PhiBSWGS <= 0.01) d.PhiBSWGS = 0.01;
  d.PhiBSWGD = m.pbswgd - m.tpbswg * delTemp; if (d.PhiBSWGD <= 0.01) d.PhiBSWGD = 0.01;
  T0 = d.TempRatio - 1.0;
  d.njtsstemp = m.njts * (1.0 + m.tnjts * T0);
  d.njtsswstemp = m.njtssw * (1.0 + m.tnjtssw * T0);
  d.njtsswgstemp = m.njtsswg * (1.0 + m.tnjtsswg * T0);
  d.njtsdtemp = m.njtsd * (1.0 + m.tnjtsd * T0);
  d.njtsswdtemp = m.njtsswd * (1.0 + m.tnjtsswd * T0);
  d.njtsswgdtemp = m.njtsswgd * (1.0 + m.tnjtsswgd * T0);
  d.Nvtms = d.vtm * m.njs;
  d.Nvtmd = d.vtm * m.njd;
  d.Nvtmrssws = d.vtm0 * d.njtsswstemp;
  d.Nvtmrsswgs = d.vtm0 * d.njtsswgstemp;
  d.Nvtmrss = d.vtm0 * d.njtsstemp;
  d.Nvtmrsswd = d.vtm0 * d.njtsswdtemp;
  d.Nvtmrsswgd = d.vtm0 * d.njtsswgdtemp;
  d.Nvtmrsd = d.vtm0 * d.njtsdtemp;
  return d;
}

}  // namespace b4
}  // namespace s21
