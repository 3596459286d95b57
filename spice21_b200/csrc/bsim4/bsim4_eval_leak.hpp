// BSIM4 evaluation, phase 4: leakage — gate-induced drain/source leakage and gate oxide tunnelling, followed by the
// finger-count scaling of every DC quantity. Follows bsim4solver.rs:1987-2640 (see bsim4_eval.hpp header).
#pragma once

namespace s21 {
namespace b4e {

// One GIDL-type current I(vde, vg_eff, vb) for an edge; `legacy` selects the gidlmod = 0 formulation.
struct B4EdgeLeakPar { double a, b, c, e, f, k, r; };
B4_HD void b4_edge_leak(bool legacy, int mtrlmod, const B4EdgeLeakPar& p, double T0, double vfbsd, double weffCJ, double vde, double vg_eff,
                        double dvg_eff_dvg, double vb, double* I, double* Gd, double* Gg, double* Gb) {
  double T1, T2, T3, T4, T5, T6, T7, T8;
  if (legacy) {
    T1 = mtrlmod == 0 ? B4_DIV((vde - vg_eff - p.e), T0) : B4_DIV((vde - vg_eff - p.e + vfbsd), T0);
    if (p.a <= 0.0 || p.b <= 0.0 || T1 <= 0.0 || p.c <= 0.0 || vb > 0.0) {
      *I = 0.0; *Gd = 0.0; *Gg = 0.0; *Gb = 0.0;
      return;
    }
    const double dT1_dVd = B4_DIV(1.0, T0);
    const double dT1_dVg = -dvg_eff_dvg * dT1_dVd;
    T2 = B4_DIV(p.b, T1);
    if (T2 < 100.0) {
      *I = p.a * weffCJ * T1 * exp(-T2);
      T3 = B4_DIV(*I * (1.0 + T2), T1);
      *Gd = T3 * dT1_dVd;
      *Gg = T3 * dT1_dVg;
    } else {
      *I = p.a * weffCJ * 3.720075976e-44;
      *Gd = *I * dT1_dVd;
      *Gg = *I * dT1_dVg;
      *I *= T1;
    }
    T4 = vb * vb;
    T5 = -vb * T4;
    T6 = p.c + T5;
    T7 = B4_DIV(T5, T6);
    T8 = B4_DIV(B4_DIV(3.0 * p.c * T4, T6), T6);
    *Gd = *Gd * T7 + *I * T8;
    *Gg = *Gg * T7;
    *Gb = -*I * T8;
    *I *= T7;
    return;
  }
  T1 = mtrlmod == 0 ? B4_DIV((vde - p.r * vg_eff - p.e), T0) : B4_DIV((vde - p.r * vg_eff - p.e + vfbsd), T0);
  if (p.a <= 0.0 || p.b <= 0.0 || T1 <= 0.0 || p.c < 0.0) {
    *I = 0.0; *Gd = 0.0; *Gg = 0.0; *Gb = 0.0;
    return;
  }
  const double dT1_dVd = B4_DIV(1.0, T0);
  const double dT1_dVg = -p.r * dT1_dVd * dvg_eff_dvg;
  T2 = B4_DIV(p.b, T1);
  if (T2 < B4C_EXPL_THRESHOLD) {
    *I = weffCJ * p.a * T1 * exp(-T2);
    T3 = B4_DIV(*I, T1) * (T2 + 1.0);
    *Gd = T3 * dT1_dVd;
    *Gg = T3 * dT1_dVg;
  } else {
    T3 = weffCJ * p.a * B4C_MIN_EXPL;
    *I = T3 * T1;
    *Gd = T3 * dT1_dVd;
    *Gg = T3 * dT1_dVg;
  }
  T4 = vb - p.f;
  T5 = T4 == 0.0 ? B4C_EXPL_THRESHOLD : B4_DIV(p.k, T4);
  if (T5 < B4C_EXPL_THRESHOLD) {
    T6 = exp(T5);
    *Gb = B4_DIV(-*I * T6 * T5, T4);
  } else {
    T6 = B4C_MAX_EXPL;
    *Gb = 0.0;
  }
  *Gd *= T6;
  *Gg *= T6;
  *I *= T6;
}

// exp() of a tunnelling exponent clamped to [MIN_EXP, MAX_EXP]; the un-clamped branch has derivative y * s1 * s2
B4_HD void b4_clamped_exp(double x, double s1, double s2, double* y, double* dy) {
  if (x > B4C_EXP_THRESHOLD) { *y = B4C_MAX_EXP; *dy = 0.0; }
  else if (x < -B4C_EXP_THRESHOLD) { *y = B4C_MIN_EXP; *dy = 0.0; }
  else { *y = exp(x); *dy = *y * s1 * s2; }
}

// Gate-tunnelling bookkeeping that the charge model also reads (Vfb is overwritten there for capmod 0).
struct B4Tunnel { double Vfb, Voxacc, dVoxacc_dVg, dVoxacc_dVb, Voxdepinv, dVoxdepinv_dVg, dVoxdepinv_dVd, dVoxdepinv_dVb; };

template <class E> B4_HD void b4_leakage(E& e, const B4Bias& v, B4Op& o, B4Chan& c, B4Tunnel& t) {
  const int mtrlmod = (int)M_(mtrlmod), igcmod = (int)M_(igcmod), igbmod = (int)M_(igbmod), tempmod = (int)M_(tempmod);
  const double toxe = M_(toxe), tp = M_(type_sign);
  double T0, T1, T2, T3, T4, T5, T6, T7, T8, T9, T10, T11, T12, T13, T14;
  double dT2_dVg, dT2_dVd, dT2_dVb, dT6_dVg, dT6_dVd, dT6_dVb, dT7_dVg, dT7_dVd, dT7_dVb, dT8_dVg, dT8_dVd, dT8_dVb;
  double dT9_dVg, dT9_dVd, dT9_dVb, dT10_dVg, dT10_dVd, dT10_dVb;

  // ---- GIDL / GISL (:1987-2187)
  T0 = mtrlmod == 0 ? 3.0 * toxe : B4_DIV(M_(epsrsub) * toxe, M_(epsrox));
  {
    const bool legacy = (int)M_(gidlmod) == 0;
    const double vfbsd = S_(vfbsd), weffCJ = S_(weffCJ);
    const B4EdgeLeakPar dp{S_(agidl), S_(bgidl), S_(cgidl), S_(egidl), S_(fgidl), S_(kgidl), S_(rgidl)};
    const B4EdgeLeakPar sp{S_(agisl), S_(bgisl), S_(cgisl), S_(egisl), S_(fgisl), S_(kgisl), S_(rgisl)};
    b4_edge_leak(legacy, mtrlmod, dp, T0, vfbsd, weffCJ, v.vds, c.vgs_eff, c.dvgs_eff_dvg, v.vbd, &o.Igidl, &o.ggidld, &o.ggidlg, &o.ggidlb);
    b4_edge_leak(legacy, mtrlmod, sp, T0, vfbsd, weffCJ, -v.vds, c.vgd_eff, c.dvgd_eff_dvg, v.vbs, &o.Igisl, &o.ggisls, &o.ggislg, &o.ggislb);
  }

  // ---- oxide voltages in accumulation and depletion/inversion (:2189-2237)
  const double Vgs_eff = c.Vgs_eff, dVgs_eff_dVg = c.dVgs_eff_dVg, Vbseff = c.Vbseff, dVbseff_dVb = c.dVbseff_dVb;
  const double Vgsteff = c.Vgsteff, dVgsteff_dVg = c.dVgsteff_dVg, dVgsteff_dVd = c.dVgsteff_dVd, dVgsteff_dVb = c.dVgsteff_dVb;
  const double Vdseff = c.Vdseff, dVdseff_dVg = c.dVdseff_dVg, dVdseff_dVd = c.dVdseff_dVd, dVdseff_dVb = c.dVdseff_dVb;
  t.Vfb = 0.0; t.Voxacc = 0.0; t.dVoxacc_dVg = 0.0; t.dVoxacc_dVb = 0.0;
  t.Voxdepinv = 0.0; t.dVoxdepinv_dVg = 0.0; t.dVoxdepinv_dVd = 0.0; t.dVoxdepinv_dVb = 0.0;
  if (igcmod != 0 || igbmod != 0) {
    const double Vfb = I_(vfbzb);
    t.Vfb = Vfb;
    const double V3 = Vfb - Vgs_eff + Vbseff - B4C_DELTA_3;
    T0 = Vfb <= 0.0 ? sqrt(V3 * V3 - 4.0 * B4C_DELTA_3 * Vfb) : sqrt(V3 * V3 + 4.0 * B4C_DELTA_3 * Vfb);
    T1 = 0.5 * (1.0 + B4_DIV(V3, T0));
    const double Vfbeff = Vfb - 0.5 * (V3 + T0);
    const double dVfbeff_dVg = T1 * dVgs_eff_dVg;
    const double dVfbeff_dVb = -T1;
    t.Voxacc = Vfb - Vfbeff;
    t.dVoxacc_dVg = -dVfbeff_dVg;
    t.dVoxacc_dVb = -dVfbeff_dVb;
    if (t.Voxacc < 0.0) { t.Voxacc = 0.0; t.dVoxacc_dVg = 0.0; t.dVoxacc_dVb = 0.0; }
    const double k1ox = S_(k1ox);
    T0 = 0.5 * k1ox;
    T3 = Vgs_eff - Vfbeff - Vbseff - Vgsteff;
    if (k1ox == 0.0) {
      t.Voxdepinv = 0.0; t.dVoxdepinv_dVg = 0.0; t.dVoxdepinv_dVd = 0.0; t.dVoxdepinv_dVb = 0.0;
    } else if (T3 < 0.0) {
      t.Voxdepinv = -T3;
      t.dVoxdepinv_dVg = -dVgs_eff_dVg + dVfbeff_dVg + dVgsteff_dVg;
      t.dVoxdepinv_dVd = dVgsteff_dVd;
      t.dVoxdepinv_dVb = dVfbeff_dVb + 1.0 + dVgsteff_dVb;
    } else {
      T1 = sqrt(T0 * T0 + T3);
      T2 = B4_DIV(T0, T1);
      t.Voxdepinv = k1ox * (T1 - T0);
      t.dVoxdepinv_dVg = T2 * (dVgs_eff_dVg - dVfbeff_dVg - dVgsteff_dVg);
      t.dVoxdepinv_dVd = -T2 * dVgsteff_dVd;
      t.dVoxdepinv_dVb = -T2 * (dVfbeff_dVb + 1.0 + dVgsteff_dVb);
    }
    t.Voxdepinv += Vgsteff;
    t.dVoxdepinv_dVg += dVgsteff_dVg;
    t.dVoxdepinv_dVd += dVgsteff_dVd;
    t.dVoxdepinv_dVb += dVgsteff_dVb;
  }
  const double Voxdepinv = t.Voxdepinv, dVoxdepinv_dVg = t.dVoxdepinv_dVg, dVoxdepinv_dVd = t.dVoxdepinv_dVd, dVoxdepinv_dVb = t.dVoxdepinv_dVb;
  const double vt_tun = tempmod < 2 ? c.Vtm : c.Vtm0;
  double Vaux = 0.0, dVaux_dVg = 0.0, dVaux_dVd = 0.0, dVaux_dVb = 0.0;

  // ---- gate-to-channel and gate-to-S/D-overlap tunnelling (:2245-2473)
  if (igcmod != 0) {
    T0 = vt_tun * S_(nigc);
    double VxNVt;
    if (igcmod == 1) {
      VxNVt = B4_DIV((Vgs_eff - tp * I_(vth0)), T0);
      if (VxNVt > B4C_EXP_THRESHOLD) { Vaux = Vgs_eff - tp * I_(vth0); dVaux_dVg = dVgs_eff_dVg; dVaux_dVd = 0.0; dVaux_dVb = 0.0; }
    } else {
      VxNVt = B4_DIV((Vgs_eff - o.von), T0);
      if (VxNVt > B4C_EXP_THRESHOLD) { Vaux = Vgs_eff - o.von; dVaux_dVg = dVgs_eff_dVg; dVaux_dVd = -c.dVth_dVd; dVaux_dVb = -c.dVth_dVb; }
    }
    if (VxNVt < -B4C_EXP_THRESHOLD) {
      Vaux = T0 * log(1.0 + B4C_MIN_EXP);
      dVaux_dVg = 0.0; dVaux_dVd = 0.0; dVaux_dVb = 0.0;
    } else if (VxNVt >= -B4C_EXP_THRESHOLD && VxNVt <= B4C_EXP_THRESHOLD) {
      const double ExpVxNVt = exp(VxNVt);
      Vaux = T0 * log(1.0 + ExpVxNVt);
      dVaux_dVg = B4_DIV(ExpVxNVt, (1.0 + ExpVxNVt));
      if (igcmod == 1) { dVaux_dVd = 0.0; dVaux_dVb = 0.0; }
      else { dVaux_dVd = -dVaux_dVg * c.dVth_dVd; dVaux_dVb = -dVaux_dVg * c.dVth_dVb; }
      dVaux_dVg *= dVgs_eff_dVg;
    }
    T2 = Vgs_eff * Vaux;
    dT2_dVg = dVgs_eff_dVg * Vaux + Vgs_eff * dVaux_dVg;
    dT2_dVd = Vgs_eff * dVaux_dVd;
    dT2_dVb = Vgs_eff * dVaux_dVb;
    T11 = S_(Aechvb);
    T12 = S_(Bechvb);
    T3 = S_(aigc) * S_(cigc) - S_(bigc);
    T4 = S_(bigc) * S_(cigc);
    T5 = T12 * (S_(aigc) + T3 * Voxdepinv - T4 * Voxdepinv * Voxdepinv);
    b4_clamped_exp(T5, T12, T3 - 2.0 * T4 * Voxdepinv, &T6, &dT6_dVg);
    dT6_dVd = dT6_dVg * dVoxdepinv_dVd;
    dT6_dVb = dT6_dVg * dVoxdepinv_dVb;
    dT6_dVg *= dVoxdepinv_dVg;
    const double Igc = T11 * T2 * T6;
    const double dIgc_dVg = T11 * (T2 * dT6_dVg + T6 * dT2_dVg);
    const double dIgc_dVd = T11 * (T2 * dT6_dVd + T6 * dT2_dVd);
    const double dIgc_dVb = T11 * (T2 * dT6_dVb + T6 * dT2_dVb);

    double Pigcd, dPigcd_dVg, dPigcd_dVd, dPigcd_dVb;
    if (M_(pigcd) == 0.0) {
      Pigcd = S_(pigcd); dPigcd_dVg = 0.0; dPigcd_dVd = 0.0; dPigcd_dVb = 0.0;
    } else {
      T11 = -S_(Bechvb);
      T12 = Vgsteff + 1.0e-20;
      T13 = B4_DIV(B4_DIV(T11, T12), T12);
      T14 = B4_DIV(-T13, T12);
      Pigcd = T13 * (1.0 - B4_DIV(0.5 * Vdseff, T12));
      dPigcd_dVg = T14 * (2.0 + 0.5 * (dVdseff_dVg - B4_DIV(3.0 * Vdseff, T12)));
      dPigcd_dVd = 0.5 * T14 * dVdseff_dVd;
      dPigcd_dVb = 0.5 * T14 * dVdseff_dVb;
    }
    T7 = -Pigcd * Vdseff;
    dT7_dVg = -Vdseff * dPigcd_dVg - Pigcd * dVdseff_dVg;
    dT7_dVd = -Vdseff * dPigcd_dVd - Pigcd * dVdseff_dVd + dT7_dVg * dVgsteff_dVd;
    dT7_dVb = -Vdseff * dPigcd_dVb - Pigcd * dVdseff_dVb + dT7_dVg * dVgsteff_dVb;
    dT7_dVg *= dVgsteff_dVg;
    T8 = T7 * T7 + 2.0e-4;
    dT8_dVg = 2.0 * T7;
    dT8_dVd = dT8_dVg * dT7_dVd;
    dT8_dVb = dT8_dVg * dT7_dVb;
    dT8_dVg *= dT7_dVg;
    if (T7 > B4C_EXP_THRESHOLD) { T9 = B4C_MAX_EXP; dT9_dVg = 0.0; dT9_dVd = 0.0; dT9_dVb = 0.0; }
    else if (T7 < -B4C_EXP_THRESHOLD) { T9 = B4C_MIN_EXP; dT9_dVg = 0.0; dT9_dVd = 0.0; dT9_dVb = 0.0; }
    else { T9 = exp(T7); dT9_dVg = T9 * dT7_dVg; dT9_dVd = T9 * dT7_dVd; dT9_dVb = T9 * dT7_dVb; }
    T1 = T9 - 1.0 + 1.0e-4;
    T10 = B4_DIV((T1 - T7), T8);
    dT10_dVg = B4_DIV((dT9_dVg - dT7_dVg - T10 * dT8_dVg), T8);
    dT10_dVd = B4_DIV((dT9_dVd - dT7_dVd - T10 * dT8_dVd), T8);
    dT10_dVb = B4_DIV((dT9_dVb - dT7_dVb - T10 * dT8_dVb), T8);
    o.Igcs = Igc * T10;
    o.gIgcsg = dIgc_dVg * T10 + Igc * dT10_dVg;
    o.gIgcsd = dIgc_dVd * T10 + Igc * dT10_dVd;
    o.gIgcsb = (dIgc_dVb * T10 + Igc * dT10_dVb) * dVbseff_dVb;
    T1 = T9 - 1.0 - 1.0e-4;
    T10 = B4_DIV((T7 * T9 - T1), T8);
    dT10_dVg = B4_DIV((dT7_dVg * T9 + (T7 - 1.0) * dT9_dVg - T10 * dT8_dVg), T8);
    dT10_dVd = B4_DIV((dT7_dVd * T9 + (T7 - 1.0) * dT9_dVd - T10 * dT8_dVd), T8);
    dT10_dVb = B4_DIV((dT7_dVb * T9 + (T7 - 1.0) * dT9_dVb - T10 * dT8_dVb), T8);
    o.Igcd = Igc * T10;
    o.gIgcdg = dIgc_dVg * T10 + Igc * dT10_dVg;
    o.gIgcdd = dIgc_dVd * T10 + Igc * dT10_dVd;
    o.gIgcdb = (dIgc_dVb * T10 + Igc * dT10_dVb) * dVbseff_dVb;

    // overlap regions: the same direct-tunnelling form in the edge oxide, symmetric about the S/D flat band
    const double vfbsd_tot = S_(vfbsd) + S_(vfbsdoff);
    T12 = S_(BechvbEdge);
    {
      T0 = v.vgs - vfbsd_tot;
      const double vgs_eff = sqrt(T0 * T0 + 1.0e-4);
      const double dvgs_eff_dvg = B4_DIV(T0, vgs_eff);
      T2 = v.vgs * vgs_eff;
      dT2_dVg = v.vgs * dvgs_eff_dvg + vgs_eff;
      T11 = S_(AechvbEdgeS);
      T3 = S_(aigs) * S_(cigs) - S_(bigs);
      T4 = S_(bigs) * S_(cigs);
      T5 = T12 * (S_(aigs) + T3 * vgs_eff - T4 * vgs_eff * vgs_eff);
      if (T5 > B4C_EXP_THRESHOLD) { T6 = B4C_MAX_EXP; dT6_dVg = 0.0; }
      else if (T5 < -B4C_EXP_THRESHOLD) { T6 = B4C_MIN_EXP; dT6_dVg = 0.0; }
      else { T6 = exp(T5); dT6_dVg = T6 * T12 * (T3 - 2.0 * T4 * vgs_eff) * dvgs_eff_dvg; }
      o.Igs = T11 * T2 * T6;
      o.gIgsg = T11 * (T2 * dT6_dVg + T6 * dT2_dVg);
      o.gIgss = -o.gIgsg;
    }
    {
      T0 = v.vgd - vfbsd_tot;
      const double vgd_eff = sqrt(T0 * T0 + 1.0e-4);
      const double dvgd_eff_dvg = B4_DIV(T0, vgd_eff);
      T2 = v.vgd * vgd_eff;
      dT2_dVg = v.vgd * dvgd_eff_dvg + vgd_eff;
      T11 = S_(AechvbEdgeD);
      T3 = S_(aigd) * S_(cigd) - S_(bigd);
      T4 = S_(bigd) * S_(cigd);
      T5 = T12 * (S_(aigd) + T3 * vgd_eff - T4 * vgd_eff * vgd_eff);
      if (T5 > B4C_EXP_THRESHOLD) { T6 = B4C_MAX_EXP; dT6_dVg = 0.0; }
      else if (T5 < -B4C_EXP_THRESHOLD) { T6 = B4C_MIN_EXP; dT6_dVg = 0.0; }
      else { T6 = exp(T5); dT6_dVg = T6 * T12 * (T3 - 2.0 * T4 * vgd_eff) * dvgd_eff_dvg; }
      o.Igd = T11 * T2 * T6;
      o.gIgdg = T11 * (T2 * dT6_dVg + T6 * dT2_dVg);
      o.gIgdd = -o.gIgdg;
    }
  } else {
    o.Igcs = 0.0; o.gIgcsg = 0.0; o.gIgcsd = 0.0; o.gIgcsb = 0.0;
    o.Igcd = 0.0; o.gIgcdg = 0.0; o.gIgcdd = 0.0; o.gIgcdb = 0.0;
    o.Igs = 0.0; o.gIgsg = 0.0; o.gIgss = 0.0; o.Igd = 0.0; o.gIgdg = 0.0; o.gIgdd = 0.0;
  }

  // ---- gate-to-body tunnelling: accumulation (ECB) and inversion (EVB) components (:2475-2589)
  if (igbmod != 0) {
    const double Vfb = t.Vfb, Voxacc = t.Voxacc;
    T0 = vt_tun * S_(nigbacc);
    T1 = -Vgs_eff + Vbseff + Vfb;
    double VxNVt = B4_DIV(T1, T0);
    if (VxNVt > B4C_EXP_THRESHOLD) { Vaux = T1; dVaux_dVg = -dVgs_eff_dVg; dVaux_dVb = 1.0; }
    else if (VxNVt < -B4C_EXP_THRESHOLD) { Vaux = T0 * log(1.0 + B4C_MIN_EXP); dVaux_dVg = 0.0; dVaux_dVb = 0.0; }
    else {
      const double ExpVxNVt = exp(VxNVt);
      Vaux = T0 * log(1.0 + ExpVxNVt);
      dVaux_dVb = B4_DIV(ExpVxNVt, (1.0 + ExpVxNVt));
      dVaux_dVg = -dVaux_dVb * dVgs_eff_dVg;
    }
    T2 = (Vgs_eff - Vbseff) * Vaux;
    dT2_dVg = dVgs_eff_dVg * Vaux + (Vgs_eff - Vbseff) * dVaux_dVg;
    dT2_dVb = -Vaux + (Vgs_eff - Vbseff) * dVaux_dVb;
    T11 = 4.97232e-7 * S_(weff) * S_(leff) * S_(ToxRatio);
    T12 = -7.45669e11 * toxe;
    T3 = S_(aigbacc) * S_(cigbacc) - S_(bigbacc);
    T4 = S_(bigbacc) * S_(cigbacc);
    T5 = T12 * (S_(aigbacc) + T3 * Voxacc - T4 * Voxacc * Voxacc);
    b4_clamped_exp(T5, T12, T3 - 2.0 * T4 * Voxacc, &T6, &dT6_dVg);
    dT6_dVb = dT6_dVg * t.dVoxacc_dVb;
    dT6_dVg *= t.dVoxacc_dVg;
    const double Igbacc = T11 * T2 * T6;
    const double dIgbacc_dVg = T11 * (T2 * dT6_dVg + T6 * dT2_dVg);
    const double dIgbacc_dVb = T11 * (T2 * dT6_dVb + T6 * dT2_dVb);

    T0 = vt_tun * S_(nigbinv);
    T1 = Voxdepinv - S_(eigbinv);
    VxNVt = B4_DIV(T1, T0);
    if (VxNVt > B4C_EXP_THRESHOLD) {
      Vaux = T1; dVaux_dVg = dVoxdepinv_dVg; dVaux_dVd = dVoxdepinv_dVd; dVaux_dVb = dVoxdepinv_dVb;
    } else if (VxNVt < -B4C_EXP_THRESHOLD) {
      Vaux = T0 * log(1.0 + B4C_MIN_EXP); dVaux_dVg = 0.0; dVaux_dVd = 0.0; dVaux_dVb = 0.0;
    } else {
      const double ExpVxNVt = exp(VxNVt);
      Vaux = T0 * log(1.0 + ExpVxNVt);
      dVaux_dVg = B4_DIV(ExpVxNVt, (1.0 + ExpVxNVt));
      dVaux_dVd = dVaux_dVg * dVoxdepinv_dVd;
      dVaux_dVb = dVaux_dVg * dVoxdepinv_dVb;
      dVaux_dVg *= dVoxdepinv_dVg;
    }
    T2 = (Vgs_eff - Vbseff) * Vaux;
    dT2_dVg = dVgs_eff_dVg * Vaux + (Vgs_eff - Vbseff) * dVaux_dVg;
    dT2_dVd = (Vgs_eff - Vbseff) * dVaux_dVd;
    dT2_dVb = -Vaux + (Vgs_eff - Vbseff) * dVaux_dVb;
    T11 *= 0.75610;
    T12 *= 1.31724;
    T3 = S_(aigbinv) * S_(cigbinv) - S_(bigbinv);
    T4 = S_(bigbinv) * S_(cigbinv);
    T5 = T12 * (S_(aigbinv) + T3 * Voxdepinv - T4 * Voxdepinv * Voxdepinv);
    b4_clamped_exp(T5, T12, T3 - 2.0 * T4 * Voxdepinv, &T6, &dT6_dVg);
    dT6_dVd = dT6_dVg * dVoxdepinv_dVd;
    dT6_dVb = dT6_dVg * dVoxdepinv_dVb;
    dT6_dVg *= dVoxdepinv_dVg;
    const double Igbinv = T11 * T2 * T6;
    const double dIgbinv_dVg = T11 * (T2 * dT6_dVg + T6 * dT2_dVg);
    const double dIgbinv_dVd = T11 * (T2 * dT6_dVd + T6 * dT2_dVd);
    const double dIgbinv_dVb = T11 * (T2 * dT6_dVb + T6 * dT2_dVb);
    o.Igb = Igbinv + Igbacc;
    o.gIgbg = dIgbinv_dVg + dIgbacc_dVg;
    o.gIgbd = dIgbinv_dVd;
    o.gIgbb = (dIgbinv_dVb + dIgbacc_dVb) * dVbseff_dVb;
  } else {
    o.Igb = 0.0; o.gIgbg = 0.0; o.gIgbd = 0.0; o.gIgbs = 0.0; o.gIgbb = 0.0;
  }

  // ---- multi-finger scaling of the DC quantities (:2591-2640)
  const double nf = I_(nf);
  double cdrain = c.cdrain;
  if (nf != 1.0) {
    cdrain *= nf;
    o.gds *= nf; o.gm *= nf; o.gmbs *= nf;
    o.gbbs *= nf; o.gbgs *= nf; o.gbds *= nf; o.csub *= nf;
    o.Igidl *= nf; o.ggidld *= nf; o.ggidlg *= nf; o.ggidlb *= nf;
    o.Igisl *= nf; o.ggisls *= nf; o.ggislg *= nf; o.ggislb *= nf;
    o.Igcs *= nf; o.gIgcsg *= nf; o.gIgcsd *= nf; o.gIgcsb *= nf;
    o.Igcd *= nf; o.gIgcdg *= nf; o.gIgcdd *= nf; o.gIgcdb *= nf;
    o.Igs *= nf; o.gIgsg *= nf; o.gIgss *= nf; o.Igd *= nf; o.gIgdg *= nf; o.gIgdd *= nf;
    o.Igb *= nf; o.gIgbg *= nf; o.gIgbd *= nf; o.gIgbb *= nf;
  }
  o.gIgbs = -(o.gIgbg + o.gIgbd + o.gIgbb);
  o.gIgcss = -(o.gIgcsg + o.gIgcsd + o.gIgcsb);
  o.gIgcds = -(o.gIgcdg + o.gIgcdd + o.gIgcdb);
  o.cd = cdrain;
}

}  // namespace b4e
}  // namespace s21

#include "bsim4_eval_charge.hpp"
