// Once-per-(model, instance-params, options) host precompute that fills the device parameter tables read by the
// CUDA device-evaluation kernels. The arithmetic is the reference's:
//   Mos1:  Mos1Model::resolve (spice21/src/comps/mos.rs:140-237), Mos1InstanceParams::resolve (:271-290),
//          Mos1InternalParams::derive (:320-476)
//   Diode: DiodeModel::from (spice21/src/comps/diode.rs:52-70), DiodeIntParams::derive (:146-212)
// but the output is a flat double[] in the layout of device_layout.h instead of nested structs.
// Compiled with -ffp-contract=off: Rust never fuses a*b+c.
#pragma once
#include <cmath>

#include "../device_layout.h"
#include "circuit.hpp"

namespace s21 {

struct SimOptions {  // analysis.rs:642-693 (settable subset)
  double temp = 300.15, tnom = 300.15, gmin = 1e-12, iabstol = 1e-12, reltol = 1e-3;
};

namespace phys {  // comps/mod.rs:24-37
constexpr double KB = 1.3806226e-23;
constexpr double Q = 1.6021918e-19;
constexpr double KB_OVER_Q = KB / Q;
constexpr double KELVIN_TO_C = 273.15;
constexpr double TEMP_REF = KELVIN_TO_C + 27.0;
constexpr double SIO2_PERMITTIVITY = 3.9 * 8.854214871e-12;
constexpr double SQRT2 = 1.4142135624;
}  // namespace phys

// Silicon band gap fit used throughout the SPICE temperature code
inline double egfet_of(double t) { return 1.16 - (7.02e-4 * (t * t)) / (t + 1108.0); }

struct Mos1Derived {
  bool has_dp = false, has_sp = false;  // internal drain/source prime nodes (mos.rs:599-610)
  double par[M1P_N];
};

inline Mos1Derived mos1_derive(const MosModelSpec& ms, const ParamBag& is, const SimOptions& opts) {
  using namespace phys;
  const ParamBag& s = ms.p;
  const double pol = ms.mos_type == 1 ? -1.0 : 1.0;
  // ---- model resolution
  const double tnom = s.has("tnom") ? s.get("tnom", 0.0) + KELVIN_TO_C : TEMP_REF;
  const double vtnom = tnom * KB_OVER_Q;
  const double egfet1 = egfet_of(tnom);
  double cox_per_area = 0.0;
  double vt0 = s.get("vt0", 0.0), kp = s.get("kp", 2.0e-5), phi = s.get("phi", 0.6), gamma = s.get("gamma", 0.0);
  if (s.has("tox")) {
    cox_per_area = SIO2_PERMITTIVITY / s.get("tox", 0.0);
    if (!s.has("kp")) kp = s.get("u0", 600.0) * cox_per_area * 1e-4;
    if (s.has("nsub")) {
      const double nsub = s.get("nsub", 0.0);
      if (nsub * 1e6 <= 1.45e16) throw S21Error(ST_INVALID, "Invalid Mos1 Substrate Doping nsub < ni (1.45e16)");
      if (!s.has("phi")) phi = std::fmax(2.0 * vtnom * std::log(nsub * 1e6 / 1.45e16), 0.1);
      const double fermis = pol * 0.5 * phi;
      double wkfng = 3.2;
      double gate_type = 1.0;
      if (ms.has_tpg) {
        if (ms.tpg > 1 || ms.tpg < -1) throw S21Error(ST_INVALID, "Invalid Mos1 tps");
        gate_type = (double)ms.tpg;
      }
      if (gate_type != 0.0) wkfng = 3.25 + 0.5 * egfet1 - pol * gate_type * 0.5 * egfet1;
      if (!s.has("gamma")) gamma = std::sqrt(2.0 * 11.70 * 8.854214871e-12 * Q * nsub * 1e6) / cox_per_area;
      if (!s.has("vt0")) {
        const double wkfngs = wkfng - (3.25 + 0.5 * egfet1 + fermis);
        const double vfb = wkfngs - s.get("nss", 0.0) * 1e4 * Q / cox_per_area;
        vt0 = vfb + pol * (gamma * std::sqrt(phi) + phi);
      }
    }
  }
  const double lambda = s.get("lambda", 0.0), pb = s.get("pb", 0.8), cbd = s.get("cbd", 0.0), cbs = s.get("cbs", 0.0);
  const double cj = s.get("cj", 0.0), cjsw = s.get("cjsw", 0.0), mj = s.get("mj", 0.5), mjsw = s.get("mjsw", 0.5);
  const double is_ = s.get("is", 1.0e-14), js = s.get("js", 1.0e-8), ld = s.get("ld", 0.0), fc = s.get("fc", 0.5);
  // ---- instance params
  if (is.has("temp")) throw S21Error(ST_INVALID, "Mos1 Instance Temperatures Are Not Supported");
  const double l = is.get("l", 1e-6), w = is.get("w", 1e-6), a_d = is.get("a_d", 1e-12), a_s = is.get("a_s", 1e-12);
  const double pd = is.get("pd", 1e-6), ps = is.get("ps", 1e-6), nrd = is.get("nrd", 1.0), nrs = is.get("nrs", 1.0);
  // ---- temperature derivation
  const double temp = opts.temp;
  const double fact1 = tnom / TEMP_REF;
  const double kt1 = KB * tnom;
  const double arg1 = -egfet1 / 2.0 / kt1 + 1.1150877 / (KB * 2.0 * TEMP_REF);
  const double pbfact1 = -2.0 * vtnom * (1.5 * std::log(fact1) + Q * arg1);
  const double kt = temp * KB;
  const double vtherm = temp * KB_OVER_Q;
  const double temp_ratio = temp / tnom;
  const double fact2 = temp / TEMP_REF;
  const double egfet = egfet_of(temp);
  const double arg = -egfet / 2.0 / kt + 1.1150877 / (KB * 2.0 * TEMP_REF);
  const double pbfact = -2.0 * vtherm * (1.5 * std::log(fact2) + Q * arg);
  const double leff = l - 2.0 * ld;
  if (leff < 0.0) throw S21Error(ST_INVALID, "Mos1 Effective Length < 0");
  const double phio = (phi - pbfact1) / fact1;
  const double phi_t = fact2 * phio + pbfact;
  const double vbi_t = vt0 - pol * (gamma * std::sqrt(phi)) + 0.5 * (egfet1 - egfet) + pol * 0.5 * (phi_t - phi);
  const double vt0_t = vbi_t + pol * gamma * std::sqrt(phi_t);
  const double sat_scale = std::exp(-egfet / vtherm + egfet1 / vtnom);
  const double isat_t = is_ * sat_scale, jsat_t = js * sat_scale;
  const double pbo = (pb - pbfact1) / fact1;
  const double gmaold = (pb - pbo) / pbo;
  const double capfact_nom = 1.0 / (1.0 + mj * (4e-4 * (tnom - TEMP_REF) - gmaold));
  const double capfact_nom_sw = 1.0 / (1.0 + mjsw * (4e-4 * (tnom - TEMP_REF) - gmaold));
  const double bulkpot_t = fact2 * pbo + pbfact;
  const double gmanew = (bulkpot_t - pbo) / pbo;
  const double capfact_t = 1.0 / (1.0 + mj * (4e-4 * (temp - TEMP_REF) - gmanew));
  const double capfact_t_sw = 1.0 / (1.0 + mjsw * (4e-4 * (temp - TEMP_REF) - gmanew));
  const double cbd_t = cbd * capfact_nom * capfact_t;
  const double cbs_t = cbs * capfact_nom * capfact_t;
  const double cj_t = cj * capfact_nom * capfact_t;
  const double cjsw_t = cjsw * capfact_nom_sw * capfact_t_sw;
  const double dep_th = fc * bulkpot_t;
  const double one_m_fc = 1.0 - fc;
  const double sarg = std::exp((-mj) * std::log(one_m_fc));
  const double sargsw = std::exp((-mjsw) * std::log(one_m_fc));
  const bool default_isat = jsat_t == 0.0 || a_d == 0.0 || a_s == 0.0;

  Mos1Derived out;
  out.has_dp = s.has("rd") || s.has("rsh");
  out.has_sp = s.has("rs") || s.has("rsh");
  double* P = out.par;
  auto junction = [&](double* J, double area, double perim, double c_fixed, double c_fixed_t) {
    const double isat = default_isat ? isat_t : jsat_t * area;
    const double czb = (c_fixed == 0.0) ? cj_t * area : c_fixed_t;
    const double czbsw = cjsw_t * perim;
    const double f2 = czb * (1.0 - fc * (1.0 + mj)) * sarg / one_m_fc + czbsw * (1.0 - fc * (1.0 + mjsw)) * sargsw / one_m_fc;
    const double f3 = czb * mj * sarg / one_m_fc / bulkpot_t + czbsw * mjsw * sargsw / one_m_fc / bulkpot_t;
    const double f4 = czb * bulkpot_t * (1.0 - one_m_fc * sarg) / (1.0 - mj) + czbsw * bulkpot_t * (1.0 - one_m_fc * sargsw) / (1.0 - mjsw) -
                      f3 / 2.0 * (dep_th * dep_th) - dep_th * f2;
    J[MJ_ISAT] = isat; J[MJ_CZB] = czb; J[MJ_CZBSW] = czbsw; J[MJ_BULKPOT] = bulkpot_t; J[MJ_DEPTH] = dep_th;
    J[MJ_F2] = f2; J[MJ_F3] = f3; J[MJ_F4] = f4;
  };
  junction(P + M1P_DJ, a_d, pd, cbd, cbd_t);
  junction(P + M1P_SJ, a_s, ps, cbs, cbs_t);
  auto ohmic = [&](const char* r_key, double nsq) {
    if (s.has(r_key)) { const double r = s.get(r_key, 0.0); return r <= 0.0 ? 0.0 : 1.0 / r; }
    if (s.has("rsh")) { const double rsh = s.get("rsh", 0.0); return rsh <= 0.0 ? 0.0 : 1.0 / rsh / nsq; }
    return 0.0;
  };
  const double kp_t = kp / temp_ratio * std::sqrt(temp_ratio);
  P[M1P_P] = pol; P[M1P_VT0T] = vt0_t; P[M1P_PHIT] = phi_t; P[M1P_GAMMA] = gamma; P[M1P_BETA] = kp_t * w / leff;
  P[M1P_LAMBDA] = lambda; P[M1P_VTHERM] = vtherm; P[M1P_COX] = cox_per_area * leff * w; P[M1P_CGSOV] = w * s.get("cgso", 0.0);
  P[M1P_CGDOV] = w * s.get("cgdo", 0.0); P[M1P_CGBOV] = leff * s.get("cgbo", 0.0); P[M1P_GRD] = ohmic("rd", nrd);
  P[M1P_GRS] = ohmic("rs", nrs); P[M1P_MJ] = mj; P[M1P_MJSW] = mjsw;
  return out;
}

struct DiodeDerived {
  bool has_r = false;  // internal "r" node when rs != 0 (diode.rs:103-110)
  double par[DP_N];
};

inline DiodeDerived diode_derive(const ParamBag& ms, const ParamBag& is, const SimOptions& opts) {
  using namespace phys;
  const double tnom = ms.get("tnom", 300.15), is_ = ms.get("is", 1e-14), n = ms.get("n", 1.0), tt = ms.get("tt", 0.0);
  const double vj = ms.get("vj", 1.0), m = ms.get("m", 0.5), eg = ms.get("eg", 1.11), xti = ms.get("xti", 3.0);
  const double fc = ms.get("fc", 0.5), bv0 = ms.get("bv", 0.0), ibv = ms.get("ibv", 1e-3), rs = ms.get("rs", 0.0), cj0 = ms.get("cj0", 0.0);
  const double temp = is.has("temp") ? is.get("temp", 0.0) : opts.temp;
  const double area = is.has("area") ? is.get("area", 0.0) : 1.0;
  const double gs = rs != 0.0 ? 1.0 / rs : 0.0;
  const double vt = KB_OVER_Q * temp;
  const double vtnom = KB_OVER_Q * tnom;
  const double fact2 = temp / TEMP_REF;
  const double egfet = 1.16 - (7.02e-4 * temp * temp) / (temp + 1108.0);
  const double arg = -egfet / (2.0 * KB * temp) + 1.1150877 / (2.0 * KB * TEMP_REF);
  const double pbfact = -2.0 * vt * (1.5 * std::log(fact2) + Q * arg);
  const double egfet1 = 1.16 - (7.02e-4 * tnom) / (tnom + 1108.0);  // reference uses tnom, not tnom^2 (diode.rs:162)
  const double arg1 = -egfet1 / (KB * 2.0 * tnom) + 1.1150877 / (2.0 * KB * TEMP_REF);
  const double fact1 = tnom / TEMP_REF;
  const double pbfact1 = -2.0 * vtnom * (1.5 * std::log(fact1) + Q * arg1);
  const double pbo = (vj - pbfact1) / fact1;
  const double vjunc = pbfact + fact2 * pbo;
  const double isat = is_ * std::exp(((temp / tnom) - 1.0) * eg / n * vt + xti / n * std::log(temp / tnom));
  const double xfc = 1.0 - std::log(fc);
  const double f1 = vjunc * (1.0 - std::exp(1.0 - m * xfc)) / (1.0 - m);
  const double vte = n * vt;
  const double vcrit = vte * (vte / std::sqrt(2.0) / isat);  // reference has no ln() here (diode.rs:179)
  double bv = bv0;
  if (bv0 != 0.0)
    for (int k = 0; k < 25; k++) bv = bv0 - vt * std::log(ibv / isat + 1.0 - bv / vt);
  const double f2 = std::exp(xfc * (1.0 + m));
  const double cz = cj0 * area;
  DiodeDerived d;
  d.has_r = rs != 0.0;
  double* P = d.par;
  P[DP_VTE] = vte; P[DP_VCRIT] = vcrit; P[DP_ISAT] = isat; P[DP_GSPR] = gs * area; P[DP_CZ] = cz; P[DP_CZ2] = cz / f2;
  P[DP_DEPTH] = fc * vj; P[DP_F1] = f1; P[DP_F3] = 1.0 - fc * (1.0 + m); P[DP_BV] = bv; P[DP_HASBV] = bv0 != 0.0 ? 1.0 : 0.0;
  P[DP_TT] = tt; P[DP_VJ] = vj; P[DP_M] = m;
  return d;
}

}  // namespace s21
