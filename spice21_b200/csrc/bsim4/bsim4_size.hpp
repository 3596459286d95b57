// BSIM4 per-instance precompute on the host: geometry/binning, temperature scaling, stress and well-proximity shifts,
// parasitic resistances/areas and the junction-diode linearisation points. Everything here runs once per (model,
// instance-parameter set); the device evaluation (bsim4_eval.hpp) only reads the result.
//
// Behaviour follows spice21/src/comps/bsim4/bsim4inst.rs:9-1378 (`from`) and its geometry helpers (:1380-1749),
// including the reference's quirks, which are kept because they change results:
//   * the drain end resistance is computed with the SOURCE flag (bsim4inst.rs:1098-1109),
//   * the internal finger resistance is discarded for geomod < 9 (a shadowing `Rint = 0`, :1642),
//   * when mtrlmod && !mtrlcompatmod the EOT iteration feeds toxp but coxp is still taken from the model card (:1340-1341).
// Host only; compiled with -ffp-contract=off.
#pragma once
#include <algorithm>
#include <cmath>
#include <optional>

#include "bsim4_model.hpp"

namespace s21 {
namespace b4 {

constexpr double EXPL_THRESHOLD = 100.0, EXP_THRESHOLD = 34.0, MAX_EXP = 5.834617425e14, MIN_EXP = 1.713908431e-15;
constexpr double MAX_EXPL = 2.688117142e+43, MIN_EXPL = 3.720075976e-44;
constexpr double STRESS_DELTA = 1e-9;  // bsim4inst.rs:6

inline double dexpb(double a) {  // bsim4/mod.rs:46-54
  if (a > EXP_THRESHOLD) return MAX_EXP * (1.0 + a - EXP_THRESHOLD);
  if (a < -EXP_THRESHOLD) return MIN_EXP;
  return std::exp(a);
}

// Instance card (bsim4/inst: Bsim4InstSpecs); an unset optional takes the default in size_and_instance().
struct InstSpec {
  std::optional<double> l, w, nf, sa, sb, sd, sca, scb, scc, sc, ad, as, pd, ps, nrd, nrs, delvto, min, rgeomod;
  std::optional<double> rbdb, rbsb, rbpb, rbps, rbpd, xgw, ngcon;
};

struct SizeDep {
#define B4S(name) double name = 0.0;
#define B4I(name)
#include "bsim4_fields.inc"
#undef B4S
#undef B4I
};
struct Internal {
#define B4S(name)
#define B4I(name) double name = 0.0;
#include "bsim4_fields.inc"
#undef B4S
#undef B4I
};

// ---- geometry helpers (bsim4inst.rs:1380-1749)
inline double dio_ijth_vjm(double Nvtm, double Ijth, double Isb, double XExpBV) {  // :1380-1385
  const double Tc = XExpBV, Tb = 1.0 + Ijth / Isb - Tc;
  return Nvtm * std::log(0.5 * (Tb + std::sqrt(Tb * Tb + 4.0 * Tc)));
}
struct FingerDiff { double intD, endD, intS, endS; };
inline FingerDiff finger_diffusions(double nf, int minSD) {  // :1489-1509
  const long NF = (long)nf;
  if (NF % 2 != 0) {
    const double nint = 2.0 * std::max((nf - 1.0) / 2.0, 0.0);
    return {nint, 1.0, nint, 1.0};
  }
  const double inner = 2.0 * std::max(nf / 2.0 - 1.0, 0.0);
  if (minSD == 1) return {inner, 2.0, nf, 0.0};
  return {nf, 0.0, inner, 2.0};
}
struct PAeff { double Ps, Pd, As, Ad; };
inline PAeff pa_eff_geo(double nf, int geo, int minSD, double Weffcj, double DMCG, double DMCI, double DMDG) {  // :1387-1487
  FingerDiff n{0, 0, 0, 0};
  if (geo < 9) n = finger_diffusions(nf, minSD);
  const double t = DMCG + DMCI;
  const double p_iso = t + t + Weffcj, p_sha = DMCG + DMCG, p_mer = DMDG + DMDG;
  const double a_iso = t * Weffcj, a_sha = DMCG * Weffcj, a_mer = DMDG * Weffcj;
  // per side: which kind of end diffusion it has (0 isolated, 1 shared, 2 merged); interior ones are always shared
  static const int src_kind[9] = {0, 0, 1, 1, 0, 1, 2, 2, 2}, drn_kind[9] = {0, 1, 0, 1, 2, 2, 0, 1, 2};
  auto side = [&](int kind, double nEnd, double nInt, double* P, double* A) {
    if (kind == 1) { *P = (nEnd + nInt) * p_sha; *A = (nEnd + nInt) * a_sha; return; }
    const double pe = kind == 0 ? p_iso : p_mer, ae = kind == 0 ? a_iso : a_mer;
    *P = nEnd * pe + nInt * p_sha;
    *A = nEnd * ae + nInt * a_sha;
  };
  PAeff r{0, 0, 0, 0};
  if (geo >= 0 && geo <= 8) {
    side(src_kind[geo], n.endS, n.intS, &r.Ps, &r.As);
    side(drn_kind[geo], n.endD, n.intD, &r.Pd, &r.Ad);
  } else if (geo == 9) {  // geo 9 / 10 only arise for an even finger count
    r.Ps = p_iso + (nf - 1.0) * p_sha; r.Pd = nf * p_sha; r.As = a_iso + (nf - 1.0) * a_sha; r.Ad = nf * a_sha;
  } else if (geo == 10) {
    r.Ps = nf * p_sha; r.Pd = p_iso + (nf - 1.0) * p_sha; r.As = nf * a_sha; r.Ad = a_iso + (nf - 1.0) * a_sha;
  }
  return r;
}
// End-diffusion resistance; `wide` selects the DMCG-long contact formula, else the Weffcj/(k*n*d) one. (:1511-1613)
inline double rds_end(double Weffcj, double Rsh, double DMCG, double DMCI, double nuEnd, int rgeo, int is_source, bool isolated) {
  bool wide, point;
  if (is_source) { wide = rgeo == 1 || rgeo == 2 || rgeo == 5; point = rgeo == 3 || rgeo == 4 || rgeo == 6; }
  else           { wide = rgeo == 1 || rgeo == 3 || rgeo == 7; point = rgeo == 2 || rgeo == 4 || rgeo == 8; }
  if (wide) return nuEnd == 0.0 ? 0.0 : Rsh * DMCG / (Weffcj * nuEnd);
  if (point) {
    const double d = isolated ? DMCG + DMCI : DMCG;
    if (nuEnd == 0.0 || d == 0.0) return 0.0;
    return Rsh * Weffcj / ((isolated ? 3.0 : 6.0) * nuEnd * d);
  }
  return 0.0;
}
inline double rds_eff_geo(double nf, int geo, int rgeo, int minSD, double Weffcj, double Rsh, double DMCG, double DMCI, double DMDG,
                          int is_source) {  // :1615-1749
  FingerDiff n{0, 0, 0, 0};
  if (geo < 9) n = finger_diffusions(nf, minSD);
  const double nuEnd = is_source ? n.endS : n.endD;
  double Rint = 0.0, Rend = 0.0;  // the interior-finger resistance of geo < 9 never reaches the result in the reference
  static const int src_kind[9] = {0, 0, 1, 1, 0, 1, 2, 2, 2}, drn_kind[9] = {0, 1, 0, 1, 2, 2, 0, 1, 2};
  if (geo >= 0 && geo <= 8) {
    const int kind = is_source ? src_kind[geo] : drn_kind[geo];
    if (kind == 2) {
      // merged end: geo 4, 6, 8 have a single merged diffusion; 5 and 7 divide by the end count
      const bool divided = (geo == 5 && !is_source) || (geo == 7 && is_source);
      Rend = divided ? Rsh * DMDG / (Weffcj * nuEnd) : Rsh * DMDG / Weffcj;
    } else {
      Rend = rds_end(Weffcj, Rsh, DMCG, DMCI, nuEnd, rgeo, is_source, kind == 0);
    }
  } else if (geo == 9 || geo == 10) {
    const bool end_side = (geo == 9) == (is_source != 0);
    if (end_side) {
      Rend = 0.5 * Rsh * DMCG / Weffcj;
      Rint = nf == 2.0 ? 0.0 : Rsh * DMCG / (Weffcj * (nf - 2.0));
    } else {
      Rend = 0.0;
      Rint = Rsh * DMCG / (Weffcj * nf);
    }
  }
  if (Rint <= 0.0) return Rend;
  if (Rend <= 0.0) return Rint;
  return Rint * Rend / (Rint + Rend);
}

// The short-channel exponential roll-off used all over the threshold model: e^x / ((e^x - 1)^2 + 2 e^x MIN_EXP).
inline double sce_theta(double x) {
  if (x < EXP_THRESHOLD) {
    const double e = std::exp(x), em1 = e - 1.0;
    return e / (em1 * em1 + 2.0 * e * MIN_EXP);
  }
  return 1.0 / (MAX_EXP - 2.0);
}

struct SizeAndInstance { SizeDep s; Internal i; };

inline SizeAndInstance size_and_instance(const Model& m, const ModelDerived& d, const InstSpec& in) {
  SizeAndInstance out;
  SizeDep& s = out.s;
  Internal& p = out.i;
  const double Temp = B4_TEMP, delTemp = Temp - m.tnom;
  const double tp = m.p();

  // ---- instance defaults (:106-146)
  p.l = in.l.value_or(5.0e-6); p.w = in.w.value_or(5.0e-6); p.nf = in.nf.value_or(1.0);
  p.sa = in.sa.value_or(0.0); p.sb = in.sb.value_or(0.0); p.sd = in.sd.value_or(2.0 * m.dmcg); p.sc = in.sc.value_or(0.0);
  p.ad = in.ad.value_or(0.0); p.as = in.as.value_or(0.0); p.pd = in.pd.value_or(0.0); p.ps = in.ps.value_or(0.0);
  p.nrd = in.nrd.value_or(1.0); p.nrs = in.nrs.value_or(1.0); p.delvto = in.delvto.value_or(0.0);
  const int minSD = in.min ? (int)*in.min : 0;
  const int rgeomod = in.rgeomod ? (int)*in.rgeomod : 0;
  p.min = minSD; p.rgeomod = rgeomod;
  p.rbdb = in.rbdb.value_or(m.rbdb); p.rbsb = in.rbsb.value_or(m.rbsb); p.rbpb = in.rbpb.value_or(m.rbpb);
  p.rbps = in.rbps.value_or(m.rbps); p.rbpd = in.rbpd.value_or(m.rbpd);
  p.xgw = in.xgw.value_or(m.xgw); p.ngcon = in.ngcon.value_or(m.ngcon);
  p.trnqsmod = m.trnqsmod; p.acnqsmod = m.acnqsmod; p.rbodymod = m.rbodymod; p.rgatemod = m.rgatemod;

  const double Ldrn = p.l, Wdrn = p.w / p.nf;
  const double Lnew = p.l + m.xl, Wnew = p.w / p.nf + m.xw;

  // ---- effective geometry (:181-223)
  s.Length = p.l; s.Width = p.w; s.NFinger = p.nf;
  {
    const double T0 = std::pow(Lnew, m.lln), T1 = std::pow(Wnew, m.lwn);
    s.dl = m.lint + (m.ll / T0 + m.lw / T1 + m.lwl / (T0 * T1));
    s.dlc = m.dlc + (m.llc / T0 + m.lwc / T1 + m.lwlc / (T0 * T1));
    const double T2 = std::pow(Lnew, m.wln), T3 = std::pow(Wnew, m.wwn);
    s.dw = m.wint + (m.wl / T2 + m.ww / T3 + m.wwl / (T2 * T3));
    const double tmp2 = m.wlc / T2 + m.wwc / T3 + m.wwlc / (T2 * T3);
    s.dwc = m.dwc + tmp2;
    s.dwj = m.dwj + tmp2;
  }
  s.leff = Lnew - 2.0 * s.dl;
  if (s.leff <= 0.0) throw ModelError("BSIM4: Effective channel length <= 0");
  s.weff = Wnew - 2.0 * s.dw;
  if (s.weff <= 0.0) throw ModelError("BSIM4: Effective channel width <= 0");
  s.leffCV = Lnew - 2.0 * s.dlc;
  if (s.leffCV <= 0.0) throw ModelError("BSIM4: Effective channel length for C-V <= 0");
  s.weffCV = Wnew - 2.0 * s.dwc;
  if (s.weffCV <= 0.0) throw ModelError("BSIM4: Effective channel width for C-V <= 0");
  s.weffCJ = Wnew - 2.0 * s.dwj;
  if (s.weffCJ <= 0.0) throw ModelError("BSIM4: Effective channel width for S/D junctions <= 0");

  // ---- binning (:225-391)
  double Inv_L, Inv_W, Inv_LW;
  if (m.binunit == 1) { Inv_L = 1.0e-6 / s.leff; Inv_W = 1.0e-6 / s.weff; Inv_LW = 1.0e-12 / (s.leff * s.weff); }
  else                { Inv_L = 1.0 / s.leff;    Inv_W = 1.0 / s.weff;    Inv_LW = 1.0 / (s.leff * s.weff); }
#define B4BIN(dst, base, lt, wt, pt) s.dst = m.base + m.lt * Inv_L + m.wt * Inv_W + m.pt * Inv_LW;
#include "bsim4_binned_table.inc"
#undef B4BIN
  s.abulkCVfactor = 1.0 + std::pow(s.clc / s.leffCV, s.cle);

  // ---- temperature scaling of mobility / saturation velocity / series resistance (:394-469)
  const double dT = d.TempRatio - 1.0;
  const double PowWeffWr = std::pow(s.weffCJ * 1.0e6, s.wr) * p.nf;
  double Rdw = 0.0, Rdwmin = 0.0, Rsw = 0.0, Rswmin = 0.0;
  s.ucs = s.ucs * std::pow(d.TempRatio, s.ucste);
  if (m.tempmod == 0) {
    s.ua = s.ua + s.ua1 * dT; s.ub = s.ub + s.ub1 * dT; s.uc = s.uc + s.uc1 * dT; s.ud = s.ud + s.ud1 * dT;
    s.vsattemp = s.vsat - s.at * dT;
    const double T10 = s.prt * dT;
    if (m.rdsmod != 0) { Rdw = s.rdw + T10; Rdwmin = m.rdwmin + T10; Rsw = s.rsw + T10; Rswmin = m.rswmin + T10; }
    s.rds0 = (s.rdsw + T10) * p.nf / PowWeffWr;
    s.rdswmin = (m.rdswmin + T10) * p.nf / PowWeffWr;
  } else {
    if (m.tempmod == 3) {
      s.ua = s.ua * std::pow(d.TempRatio, s.ua1); s.ub = s.ub * std::pow(d.TempRatio, s.ub1);
      s.uc = s.uc * std::pow(d.TempRatio, s.uc1); s.ud = s.ud * std::pow(d.TempRatio, s.ud1);
    } else {
      s.ua = s.ua * (1.0 + s.ua1 * delTemp); s.ub = s.ub * (1.0 + s.ub1 * delTemp);
      s.uc = s.uc * (1.0 + s.uc1 * delTemp); s.ud = s.ud * (1.0 + s.ud1 * delTemp);
    }
    s.vsattemp = s.vsat * (1.0 - s.at * delTemp);
    const double T10 = 1.0 + s.prt * delTemp;
    if (m.rdsmod != 0) { Rdw = s.rdw * T10; Rdwmin = m.rdwmin * T10; Rsw = s.rsw * T10; Rswmin = m.rswmin * T10; }
    s.rds0 = s.rdsw * T10 * p.nf / PowWeffWr;
    s.rdswmin = m.rdswmin * T10 * p.nf / PowWeffWr;
  }
  if (Rdw < 0.0) Rdw = 0.0;
  if (Rdwmin < 0.0) Rdwmin = 0.0;
  if (Rsw < 0.0) Rsw = 0.0;
  if (Rswmin < 0.0) Rswmin = 0.0;
  s.rd0 = Rdw / PowWeffWr; s.rdwmin = Rdwmin / PowWeffWr; s.rs0 = Rsw / PowWeffWr; s.rswmin = Rswmin / PowWeffWr;

  if (s.u0 > 1.0) s.u0 = s.u0 / 1.0e4;
  s.u0temp = s.u0 * (1.0 - s.up * std::exp(-s.leff / s.lp)) * std::pow(d.TempRatio, s.ute);
  if (s.eu < 0.0) s.eu = 0.0;
  if (s.ucs < 0.0) s.ucs = 0.0;

  s.vfbsdoff = s.vfbsdoff * (1.0 + s.tvfbsdoff * delTemp);
  s.voff = s.voff * (1.0 + s.tvoff * delTemp);
  s.nfactor = s.nfactor + s.tnfactor * delTemp / m.tnom;
  s.voffcv = s.voffcv * (1.0 + s.tvoffcv * delTemp);
  s.eta0 = s.eta0 + s.teta0 * delTemp / m.tnom;

  if (m.has("vtl") && m.vtl > 0.0) {  // source-end velocity limit (:495-503)
    s.lc = m.lc < 0.0 ? 0.0 : m.lc;
    const double T0 = s.leff / (s.xn * s.leff + s.lc);
    s.tfactor = (1.0 - T0) / (1.0 + T0);
  }

  s.cgdo = (m.cgdo + s.cf) * s.weffCV;
  s.cgso = (m.cgso + s.cf) * s.weffCV;
  s.cgbo = m.cgbo * s.leffCV * p.nf;

  // ---- electrostatics (:509-576)
  if (!m.has("ndep") && m.has("gamma1")) {
    const double T0 = s.gamma1 * d.coxe;
    s.ndep = 3.01248e22 * T0 * T0;
  }
  s.phi = d.vtm0 * std::log(s.ndep / d.ni) + s.phin + 0.4;
  s.sqrtPhi = std::sqrt(s.phi);
  s.phis3 = s.sqrtPhi * s.phi;
  s.Xdep0 = std::sqrt(2.0 * d.epssub / (QE * s.ndep * 1.0e6)) * s.sqrtPhi;
  s.sqrtXdep0 = std::sqrt(s.Xdep0);
  s.litl = m.mtrlmod == 0 ? std::sqrt(3.0 * 3.9 / m.epsrox * s.xj * m.toxe) : std::sqrt(m.epsrsub / m.epsrox * s.xj * m.toxe);
  s.vbi = d.vtm0 * std::log(s.nsd * s.ndep / (d.ni * d.ni));
  if (m.mtrlmod == 0) {
    s.vfbsd = s.ngate > 0.0 ? d.vtm0 * std::log(s.ngate / s.nsd) : 0.0;
  } else {
    double T0 = d.vtm0 * std::log(s.nsd / d.ni);
    const double T1 = 0.5 * d.Eg0;
    if (T0 > T1) T0 = T1;
    s.vfbsd = m.phig - (m.easub + T1 - tp * T0);
  }
  s.cdep0 = std::sqrt(QE * d.epssub * s.ndep * 1.0e6 / 2.0 / s.phi);
  s.ToxRatio = std::exp(s.ntox * std::log(m.toxref / m.toxe)) / m.toxe / m.toxe;
  s.ToxRatioEdge = std::exp(s.ntox * std::log(m.toxref / (m.toxe * s.poxedge))) / m.toxe / m.toxe / s.poxedge / s.poxedge;
  s.Aechvb = m.mos_type == 0 ? 4.97232e-7 : 3.42537e-7;
  s.Bechvb = m.mos_type == 0 ? 7.45669e11 : 1.16645e12;
  s.AechvbEdgeS = s.Aechvb * s.weff * m.dlcig * s.ToxRatioEdge;
  s.AechvbEdgeD = s.Aechvb * s.weff * m.dlcigd * s.ToxRatioEdge;
  s.BechvbEdge = -s.Bechvb * m.toxe * s.poxedge;
  s.Aechvb *= s.weff * s.leff * s.ToxRatio;
  s.Bechvb *= -m.toxe;
  s.mstar = 0.5 + std::atan(s.minv) / PI;
  s.mstarcv = 0.5 + std::atan(s.minvcv) / PI;
  s.voffcbn = s.voff + m.voffl / s.leff;
  s.voffcbncv = s.voffcv + m.voffcvl / s.leff;
  s.ldeb = std::sqrt(d.epssub * d.vtm0 / (QE * s.ndep * 1.0e6)) / 3.0;
  s.acde *= std::pow(s.ndep / 2.0e16, -0.25);

  // ---- body-effect coefficients, flat band and threshold (:578-647)
  const bool k1g = m.has("k1"), k2g = m.has("k2");
  s.k1 = k1g ? m.k1 + m.lk1 * Inv_L + m.wk1 * Inv_W + m.pk1 * Inv_LW : 0.53;
  s.k2 = k2g ? m.k2 + m.lk2 * Inv_L + m.wk2 * Inv_W + m.pk2 * Inv_LW : -0.0186;
  if (!k1g && !k2g) {
    if (!m.has("vbx")) s.vbx = s.phi - 7.7348e-4 * s.ndep * s.xt * s.xt;
    if (s.vbx > 0.0) s.vbx = -s.vbx;
    if (s.vbm > 0.0) s.vbm = -s.vbm;
    if (!m.has("gamma1")) s.gamma1 = 5.753e-12 * std::sqrt(s.ndep) / d.coxe;
    if (!m.has("gamma2")) s.gamma2 = 5.753e-12 * std::sqrt(s.nsub) / d.coxe;
    const double T0 = s.gamma1 - s.gamma2;
    const double T1 = std::sqrt(s.phi - s.vbx) - s.sqrtPhi;
    const double T2 = std::sqrt(s.phi * (s.phi - s.vbm)) - s.phi;
    s.k2 = T0 * T1 / (2.0 * T2 + s.vbm);
    s.k1 = s.gamma2 - 2.0 * s.k2 * std::sqrt(s.phi - s.vbm);
  }
  if (!m.has("vfb")) {
    if (m.has("vth0")) {
      s.vfb = tp * s.vth0 - s.phi - s.k1 * s.sqrtPhi;
    } else if (m.mtrlmod != 0 && m.has("phig") && m.has("nsub")) {
      double T0 = d.vtm0 * std::log(s.nsub / d.ni);
      const double T1 = 0.5 * d.Eg0;
      if (T0 > T1) T0 = T1;
      s.vfb = m.phig - (m.easub + T1 + tp * T0);
    } else {
      s.vfb = -1.0;
    }
  }
  if (!m.has("vth0")) s.vth0 = tp * (s.vfb + s.phi + s.k1 * s.sqrtPhi);
  s.k1ox = s.k1 * m.toxe / m.toxm;

  // ---- short-channel / DIBL prefactors (:649-713)
  {
    const double lt0 = std::sqrt(d.epssub / (m.epsrox * EPS0) * m.toxe * s.Xdep0);
    s.theta0vb0 = sce_theta(s.dsub * s.leff / lt0);
    s.thetaRout = s.pdibl1 * sce_theta(s.drout * s.leff / lt0) + s.pdibl2;
    const double vbi_phi = s.vbi - s.phi;
    const double lt1 = d.factor1 * std::sqrt(s.Xdep0);
    const double T8 = (s.dvt0w * sce_theta(s.dvt1w * s.weff * s.leff / lt1)) * vbi_phi;
    const double T9 = s.dvt0 * sce_theta(s.dvt1 * s.leff / lt1) * vbi_phi;
    const double T4 = m.toxe * s.phi / (s.weff + s.w0);
    const double T0 = std::sqrt(1.0 + s.lpe0 / s.leff);
    double T3 = 0.0;
    if (m.tempmod == 1 || m.tempmod == 0) T3 = (s.kt1 + s.kt1l / s.leff) * (d.TempRatio - 1.0);
    if (m.tempmod == 2 || m.tempmod == 3) T3 = -s.kt1 * (d.TempRatio - 1.0);
    const double T5 = s.k1ox * (T0 - 1.0) * s.sqrtPhi + T3;
    s.vfbzbfactor = -T8 - T9 + s.k3 * T4 + T5 - s.phi - s.k1 * s.sqrtPhi;
  }

  // ---- layout-dependent stress: size part (:715-734)
  const double W_tmp = Wnew + m.wlod;
  {
    double T0 = std::pow(Lnew, m.llodku0), T1 = std::pow(W_tmp, m.wlodku0);
    s.ku0 = 1.0 + (m.lku0 / T0 + m.wku0 / T1 + m.pku0 / (T0 * T1));
    T0 = std::pow(Lnew, m.llodvth); T1 = std::pow(W_tmp, m.wlodvth);
    s.kvth0 = 1.0 + (m.lkvth0 / T0 + m.wkvth0 / T1 + m.pkvth0 / (T0 * T1));
    s.kvth0 = std::sqrt(s.kvth0 * s.kvth0 + STRESS_DELTA);
    s.ku0temp = s.ku0 * (1.0 + m.tku0 * (d.TempRatio - 1.0)) + STRESS_DELTA;
    const double Inv_saref = 1.0 / (m.saref + 0.5 * Ldrn), Inv_sbref = 1.0 / (m.sbref + 0.5 * Ldrn);
    s.inv_od_ref = Inv_saref + Inv_sbref;
    s.rho_ref = m.ku0 / s.ku0temp * s.inv_od_ref;
  }

  if (m.mobmod == 3) {  // Vgsteff at threshold for the high-k mobility model (:738-776)
    const double Theta0 = sce_theta(s.dvt1 * s.leff / (d.factor1 * s.sqrtXdep0));
    const double tmp3 = (s.nfactor * (d.epssub / s.Xdep0) + s.cdsc * Theta0 + s.cit) / d.coxe;
    const double n0 = tmp3 >= -0.5 ? 1.0 + tmp3 : (1.0 + 3.0 * tmp3) * (1.0 / (3.0 + 8.0 * tmp3));
    const double T0 = n0 * d.vtm, T2 = s.voffcbn / T0;
    double T3;
    if (T2 < -EXP_THRESHOLD) T3 = d.coxe * MIN_EXP / s.cdep0;
    else if (T2 > EXP_THRESHOLD) T3 = d.coxe * MAX_EXP / s.cdep0;
    else T3 = std::exp(T2) * d.coxe / s.cdep0;
    s.VgsteffVth = T0 * std::log(2.0) / (s.mstar + T3 * n0);
  }
  s.dvtp2factor = s.dvtp5 + s.dvtp2 * dexpb(-s.dvtp3 * std::log(s.leff));

  // ---- layout-dependent stress: instance part (:785-827)
  if (p.sa > 0.0 && p.sb > 0.0 && (p.nf == 1.0 || (p.nf > 1.0 && p.sd > 0.0))) {
    double Inv_sa = 0.0, Inv_sb = 0.0;
    const double kvsat = std::min(std::max(m.kvsat, -1.0), 1.0);
    const long nfi = (long)p.nf;
    for (long i = 0; i < nfi; i++) {
      Inv_sa += 1.0 / p.nf / (p.sa + 0.5 * Ldrn + (double)i * (p.sd + Ldrn));
      Inv_sb += 1.0 / p.nf / (p.sb + 0.5 * Ldrn + (double)i * (p.sd + Ldrn));
    }
    const double Inv_ODeff = Inv_sa + Inv_sb;
    const double rho = m.ku0 / s.ku0temp * Inv_ODeff;
    p.u0temp = s.u0temp * ((1.0 + rho) / (1.0 + s.rho_ref));
    p.vsattemp = s.vsattemp * ((1.0 + kvsat * rho) / (1.0 + kvsat * s.rho_ref));
    const double OD_offset = Inv_ODeff - s.inv_od_ref;
    p.vth0 = s.vth0 + m.kvth0 / s.kvth0 * OD_offset;
    p.eta0 = s.eta0 + m.steta0 / std::pow(s.kvth0, m.lodeta0) * OD_offset;
    p.k2 = s.k2 + m.stk2 / std::pow(s.kvth0, m.lodk2) * OD_offset;
  } else {
    p.u0temp = s.u0temp; p.vth0 = s.vth0; p.vsattemp = s.vsattemp; p.eta0 = s.eta0; p.k2 = s.k2;
  }

  // ---- well proximity (:830-877)
  if (m.wpemod != 0) {
    p.sca = in.sca.value_or(0.0); p.scb = in.scb.value_or(0.0); p.scc = in.scc.value_or(0.0);
    if (!in.sca && !in.scb && !in.scc && in.sc && p.sc > 0.0) {
      const double T1 = p.sc + Wdrn, T2 = 1.0 / m.scref;
      p.sca = m.scref * m.scref / (p.sc * T1);
      p.scb = ((0.1 * p.sc + 0.01 * m.scref) * std::exp(-10.0 * p.sc * T2) - (0.1 * T1 + 0.01 * m.scref) * std::exp(-10.0 * T1 * T2)) / Wdrn;
      p.scc = ((0.05 * p.sc + 0.0025 * m.scref) * std::exp(-20.0 * p.sc * T2) - (0.05 * T1 + 0.0025 * m.scref) * std::exp(-20.0 * T1 * T2)) / Wdrn;
    }
    if (p.sca < 0.0) p.sca = 0.0;
    if (p.scb < 0.0) p.scb = 0.0;
    if (p.scc < 0.0) p.scc = 0.0;
    if (p.sc < 0.0) p.sc = 0.0;
    const double sceff = p.sca + m.web * p.scb + m.wec * p.scc;
    p.vth0 += s.kvth0we * sceff;
    p.k2 += s.k2we * sceff;
    double T3 = 1.0 + s.ku0we * sceff;
    if (T3 <= 0.0) T3 = 0.0;
    p.u0temp *= T3;
  }

  // ---- threshold shift and the quantities hanging off vth0 / k2 (:880-913)
  p.vth0 += p.delvto;
  p.vfb = s.vfb + tp * p.delvto;
  {
    const double T3 = tp * p.vth0 - p.vfb - s.phi;
    p.vtfbphi1 = std::max(tp > 0.0 ? T3 + T3 : 2.5 * T3, 0.0);
    p.vtfbphi2 = std::max(4.0 * T3, 0.0);
  }
  if (p.k2 < 0.0) {
    const double T0 = 0.5 * s.k1 / p.k2;
    p.vbsc = std::min(std::max(0.9 * (s.phi - T0 * T0), -30.0), -3.0);
  } else {
    p.vbsc = -30.0;
  }
  if (p.vbsc > s.vbm) p.vbsc = s.vbm;
  p.k2ox = p.k2 * m.toxe / m.toxm;
  p.vfbzb = s.vfbzbfactor + tp * p.vth0;

  // ---- body resistance network (:919-1003)
  {
    const double lnl = std::log(s.leff * 1.0e6), lnw = std::log(s.weff * 1.0e6), lnnf = std::log(p.nf);
    auto scaled = [&](double r0, double el, double ew, double enf) { return r0 * std::exp(el * lnl + ew * lnw + enf * lnnf); };
    auto par = [](double a, double b) { return a * b / (a + b); };
    auto cond = [&](double r) { return r < 1.0e-3 ? 1.0e3 : m.gbmin + 1.0 / r; };
    if (m.rbodymod == 2) {
      if (m.bodymode == 5) {
        p.rbsb = par(scaled(m.rbsbx0, m.rbsdbxl, m.rbsdbxw, m.rbsdbxnf), scaled(m.rbsby0, m.rbsdbyl, m.rbsdbyw, m.rbsdbynf));
        p.rbdb = par(scaled(m.rbdbx0, m.rbsdbxl, m.rbsdbxw, m.rbsdbxnf), scaled(m.rbdby0, m.rbsdbyl, m.rbsdbyw, m.rbsdbynf));
      }
      if (m.bodymode == 3 || m.bodymode == 5) {
        p.rbps = scaled(m.rbps0, m.rbpsl, m.rbpsw, m.rbpsnf);
        p.rbpd = scaled(m.rbpd0, m.rbpdl, m.rbpdw, m.rbpdnf);
      }
      p.rbpb = par(scaled(m.rbpbx0, m.rbpbxl, m.rbpbxw, m.rbpbxnf), scaled(m.rbpby0, m.rbpbyl, m.rbpbyw, m.rbpbynf));
    }
    if (m.rbodymod == 1 || (m.rbodymod == 2 && m.bodymode == 5)) {
      p.grbdb = cond(p.rbdb); p.grbpb = cond(p.rbpb); p.grbps = cond(p.rbps); p.grbsb = cond(p.rbsb); p.grbpd = cond(p.rbpd);
    }
    if (m.rbodymod == 2 && m.bodymode == 3) {
      p.grbdb = m.gbmin; p.grbsb = m.gbmin;
      p.grbpb = cond(p.rbpb); p.grbps = cond(p.rbps); p.grbpd = cond(p.rbpd);
    }
    if (m.rbodymod == 2 && m.bodymode == 1) {
      p.grbdb = m.gbmin; p.grbsb = m.gbmin; p.grbps = 1.0e3; p.grbpd = 1.0e3;
      p.grbpb = cond(p.rbpb);
    }
  }

  // ---- geometry-dependent parasitics (:1009-1118)
  p.grgeltd = m.rshg * (p.xgw + s.weffCJ / 3.0 / p.ngcon) / (p.ngcon * p.nf * (Lnew - m.xgl));
  p.grgeltd = p.grgeltd > 0.0 ? 1.0 / p.grgeltd : 1.0e3;
  const double DMCGeff = m.dmcg - m.dmcgt, DMCIeff = m.dmci, DMDGeff = m.dmdg - m.dmcgt;
  const int geomod = (int)m.geomod;
  const PAeff pa = pa_eff_geo(p.nf, geomod, minSD, s.weffCJ, DMCGeff, DMCIeff, DMDGeff);
  auto perimeter = [&](const std::optional<double>& given, double calc) {
    double v = calc;
    if (given) v = *given < 0.0 ? 0.0 : (m.permod == 0 ? *given : *given - s.weffCJ * p.nf);
    return v < 0.0 ? 0.0 : v;
  };
  p.Pseff = perimeter(in.ps, pa.Ps);
  p.Pdeff = perimeter(in.pd, pa.Pd);
  p.Aseff = std::max(in.as.value_or(pa.As), 0.0);
  p.Adeff = std::max(in.ad.value_or(pa.Ad), 0.0);
  auto end_conductance = [&](const std::optional<double>& squares) {
    double r = 0.0;
    if (squares) r = m.rsh * *squares;
    else if (rgeomod > 0) r = rds_eff_geo(p.nf, geomod, rgeomod, minSD, s.weffCJ, m.rsh, DMCGeff, DMCIeff, DMDGeff, 1);
    return r > 0.0 ? 1.0 / r : 1.0e3;
  };
  p.sourceConductance = end_conductance(in.nrs);
  p.drainConductance = end_conductance(in.nrd);

  // ---- junction diode saturation currents and linearisation points (:1120-1218, :1355-1373)
  auto sat_current = [&](double A, double P, double jA, double jP, double jG) {
    return (A <= 0.0 && P <= 0.0) ? 0.0 : A * jA + P * jP + s.weffCJ * p.nf * jG;
  };
  struct Junction { double *XExpBV, *vjmFwd, *vjmRev, *IVjmFwd, *IVjmRev, *slpFwd, *slpRev; };
  auto junction = [&](double Isat, double Nvtm, double bv, double xjbv, double ijthfwd, double ijthrev, Junction j) {
    if (!(Isat > 0.0)) return;
    if (m.diomod == 0) {
      *j.XExpBV = (bv / Nvtm) > EXP_THRESHOLD ? xjbv * MIN_EXP : xjbv * std::exp(-bv / Nvtm);
    } else if (m.diomod == 1) {
      *j.vjmFwd = dio_ijth_vjm(Nvtm, ijthfwd, Isat, 0.0);
      *j.IVjmFwd = Isat * std::exp(*j.vjmFwd / Nvtm);
    } else if (m.diomod == 2) {
      if ((bv / Nvtm) > EXP_THRESHOLD) {
        *j.XExpBV = xjbv * MIN_EXP;
      } else {
        *j.XExpBV = std::exp(-bv / Nvtm);
        *j.XExpBV *= xjbv;
      }
      *j.vjmFwd = dio_ijth_vjm(Nvtm, ijthfwd, Isat, *j.XExpBV);
      const double T0 = std::exp(*j.vjmFwd / Nvtm);
      *j.IVjmFwd = Isat * (T0 - *j.XExpBV / T0 + *j.XExpBV - 1.0);
      *j.slpFwd = Isat * (T0 + *j.XExpBV / T0) / Nvtm;
      double T2 = ijthrev / Isat;
      if (T2 < 1.0) T2 = 10.0;
      *j.vjmRev = -bv - Nvtm * std::log((T2 - 1.0) / xjbv);
      const double T1 = xjbv * std::exp(-(bv + *j.vjmRev) / Nvtm);
      *j.IVjmRev = Isat * (1.0 + T1);
      *j.slpRev = -Isat * T1 / Nvtm;
    }
  };
  p.SourceSatCurrent = sat_current(p.Aseff, p.Pseff, d.SjctTempSatCurDensity, d.SjctSidewallTempSatCurDensity, d.SjctGateSidewallTempSatCurDensity);
  p.DrainSatCurrent = sat_current(p.Adeff, p.Pdeff, d.DjctTempSatCurDensity, d.DjctSidewallTempSatCurDensity, d.DjctGateSidewallTempSatCurDensity);
  junction(p.SourceSatCurrent, d.vtm * m.njs, m.bvs, m.xjbvs, m.ijthsfwd, m.ijthsrev,
           Junction{&p.XExpBVS, &p.vjsmFwd, &p.vjsmRev, &p.IVjsmFwd, &p.IVjsmRev, &p.SslpFwd, &p.SslpRev});
  junction(p.DrainSatCurrent, d.vtm * m.njd, m.bvd, m.xjbvd, m.ijthdfwd, m.ijthdrev,
           Junction{&p.XExpBVD, &p.vjdmFwd, &p.vjdmRev, &p.IVjdmFwd, &p.IVjdmRev, &p.DslpFwd, &p.DslpRev});

  // ---- trap-assisted tunnelling saturation currents (:1220-1242)
  {
    const double T7 = d.Eg0 / d.vtm * (d.TempRatio - 1.0);
    const double T10 = s.weffCJ * p.nf;
    const double T11 = std::sqrt(m.jtweff / s.weffCJ) + 1.0;
    p.SjctTempRevSatCur = dexpb(m.xtss * T7) * p.Aseff * m.jtss;
    p.DjctTempRevSatCur = dexpb(m.xtsd * T7) * p.Adeff * m.jtsd;
    p.SswTempRevSatCur = dexpb(m.xtssws * T7) * p.Pseff * m.jtssws;
    p.DswTempRevSatCur = dexpb(m.xtsswd * T7) * p.Pdeff * m.jtsswd;
    p.SswgTempRevSatCur = dexpb(m.xtsswgs * T7) * T10 * T11 * m.jtsswgs;
    p.DswgTempRevSatCur = dexpb(m.xtsswgd * T7) * T10 * T11 * m.jtsswgd;
  }

  // ---- electrical oxide thickness from EOT for non-SiO2 stacks (:1244-1345)
  if (m.mtrlmod != 0 && m.mtrlcompatmod == 0) {
    const double Vtm0eot = KB_OVER_Q * m.tempeot, Vtmeot = Vtm0eot;
    const double vbieot = Vtm0eot * std::log(s.nsd * s.ndep / (d.ni * d.ni));
    const double phieot = Vtm0eot * std::log(s.ndep / d.ni) + s.phin + 0.4;
    const double vddeot = tp * m.vddeot;
    double Vgs_eff = vddeot;
    {
      const double tmp2 = p.vfb + phieot, T0 = m.epsrgate * EPS0;
      if (s.ngate > 1.0e18 && s.ngate < 1.0e25 && vddeot > tmp2 && T0 != 0.0) {
        const double T1 = 1.0e6 * QE * T0 * s.ngate / (d.coxe * d.coxe);
        const double T8 = vddeot - tmp2;
        const double T4 = std::sqrt(1.0 + 2.0 * T8 / T1);
        const double T2 = 2.0 * T8 / (T4 + 1.0);
        const double T3 = 0.5 * T2 * T2 / T1;
        const double T7 = 1.12 - T3 - 0.05;
        const double T6 = std::sqrt(T7 * T7 + 0.224);
        Vgs_eff = vddeot - (1.12 - 0.5 * (T7 + T6));
      }
    }
    const double V0 = vbieot - phieot;
    const double lt1 = d.factor1 * s.sqrtXdep0;
    const double Theta0 = sce_theta(s.dvt1 * m.leffeot / lt1);
    const double Delt_vth = s.dvt0 * Theta0 * V0;
    const double T2w = s.dvt0w * sce_theta(s.dvt1w * m.weffeot * m.leffeot / lt1) * V0;
    const double TempRatioeot = m.tempeot / m.tnom - 1.0;
    const double T1 = s.k1ox * (std::sqrt(1.0 + s.lpe0 / m.leffeot) - 1.0) * std::sqrt(phieot) + (s.kt1 + s.kt1l / m.leffeot) * TempRatioeot;
    const double Vth_NarrowW = m.toxe * phieot / (m.weffeot + s.w0);
    const double Lpe_Vb = std::sqrt(1.0 + s.lpeb / m.leffeot);
    double Vth = tp * p.vth0 + (s.k1ox - s.k1) * std::sqrt(phieot) * Lpe_Vb - Delt_vth - T2w + s.k3 * Vth_NarrowW + T1;
    const double tmp3 = (s.nfactor * (d.epssub / s.Xdep0) + s.cdsc * Theta0 + s.cit) / d.coxe;
    const double n = tmp3 >= -0.5 ? 1.0 + tmp3 : (1.0 + 3.0 * tmp3) * (1.0 / (3.0 + 8.0 * tmp3));
    if (s.dvtp0 > 0.0) {
      const double t_ = m.leffeot + s.dvtp0 * 2.0;
      Vth -= n * (m.tempmod < 2 ? Vtmeot : Vtm0eot) * std::log(m.leffeot / t_);
    }
    const double Vgsteff = Vgs_eff - Vth;
    const double T3 = tp * p.vth0 - p.vfb - phieot;
    const double vtfbphi2eot = std::max(4.0 * T3, 0.0);
    double toxpf = m.toxe;
    for (int it = 0; it < 4; it++) {
      const double T0 = (Vgsteff + vtfbphi2eot) / (2.0e8 * toxpf);
      const double Tcen = m.ados * 1.9e-9 / (1.0 + std::exp(m.bdos * 0.7 * std::log(T0)));
      toxpf = m.toxe - m.epsrox / m.epsrsub * Tcen;
    }
    p.toxp = toxpf;
    p.coxp = m.epsrox * EPS0 / m.toxp;
  } else {
    p.toxp = m.toxp;
    p.coxp = d.coxp;
  }
  return out;
}

}  // namespace b4
}  // namespace s21
