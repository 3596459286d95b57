// BSIM4 evaluation, phase 3: channel DC — threshold, effective gate drive, mobility, saturation, output-resistance
// terms, drain current with its three derivatives, impact-ionisation substrate current, the intrinsic-gate input
// resistance and the bias-dependent S/D resistances. Follows bsim4solver.rs:779-1989 (see bsim4_eval.hpp header).
#pragma once

namespace s21 {
namespace b4e {

// Channel quantities later phases need (gate tunnelling, charge model).
struct B4Chan {
  double Vds, Vgs, Vbs, Vdb;
  double Vbseff, dVbseff_dVb, Phis, sqrtPhis, dsqrtPhis_dVb, Xdep, dXdep_dVb;
  double Vth, dVth_dVb, dVth_dVd, n, dn_dVb, dn_dVd;
  double vgs_eff, vgd_eff, dvgs_eff_dvg, dvgd_eff_dvg, Vgs_eff, dVgs_eff_dVg, Vgst;
  double Vgsteff, dVgsteff_dVg, dVgsteff_dVd, dVgsteff_dVb;
  double Weff, Abulk, Abulk0, dAbulk0_dVb, dAbulk_dVg, dAbulk_dVb, AbovVgst2Vtm;
  double ueff, EsatL, Vdsat, dVdsat_dVg, dVdsat_dVd, dVdsat_dVb;
  double Vdseff, dVdseff_dVg, dVdseff_dVd, dVdseff_dVb;
  double Coxeff, dCoxeff_dVg, Tcen, dTcen_dVg, thetavth, Vtm, Vtm0, Leff, Lpe_Vb, Vth_NarrowW;
  double cdrain;
};

// soft-limited roll-off factor: e^x/((e^x-1)^2 + 2 e^x MIN_EXP) and its derivative w.r.t. Vb through the length `lt`
B4_HD void b4_sce_rolloff(double x, double dlt_dVb, double lt, double* theta, double* dtheta_dVb) {
  if (x < B4C_EXP_THRESHOLD) {
    const double T1 = exp(x);
    const double T2 = T1 - 1.0;
    const double T3 = T2 * T2;
    const double T4 = T3 + 2.0 * T1 * B4C_MIN_EXP;
    *theta = B4_DIV(T1, T4);
    const double dT1_dVb = B4_DIV(-x * T1 * dlt_dVb, lt);
    *dtheta_dVb = B4_DIV(B4_DIV(dT1_dVb * (T4 - 2.0 * T1 * (T2 + B4C_MIN_EXP)), T4), T4);
  } else {
    *theta = B4_DIV(1.0, (B4C_MAX_EXP - 2.0));
    *dtheta_dVb = 0.0;
  }
}

template <class E> B4_HD void b4_channel_dc(E& e, const B4Bias& v, B4Op& o, B4Chan& c) {
  const double tp = M_(type_sign);
  const int mobmod = (int)M_(mobmod), rdsmod = (int)M_(rdsmod), rgatemod = (int)M_(rgatemod), tempmod = (int)M_(tempmod);
  const int mtrlmod = (int)M_(mtrlmod), mtrlcompatmod = (int)M_(mtrlcompatmod);
  double T0, T1, T2, T3, T4, T5, T6, T7, T8, T9, T10, T11, T12, T13, T14;
  double dT0_dVg, dT0_dVd, dT0_dVb, dT1_dVg, dT1_dVd, dT1_dVb, dT2_dVg, dT2_dVd, dT2_dVb, dT3_dVg, dT3_dVd, dT3_dVb;
  double dT5_dVg, dT5_dVd, dT5_dVb, dT6_dVg, dT6_dVd, dT6_dVb, dT8_dVg, dT8_dVd, dT8_dVb, dT9_dVg, dT9_dVd, dT9_dVb;
  double dT10_dVg, dT10_dVd, dT10_dVb, dT4_dVd;
  double tmp, tmp1, tmp2, tmp3, tmp4;

  double Vds, Vgs, Vbs, Vdb;
  if (v.vds >= 0.0) {
    o.mode = 1;
    Vds = v.vds; Vgs = v.vgs; Vbs = v.vbs; Vdb = v.vds - v.vbs;
  } else {
    o.mode = -1;
    Vds = -v.vds; Vgs = v.vgd; Vbs = v.vbd; Vdb = -v.vbs;
  }
  c.Vds = Vds; c.Vgs = Vgs; c.Vbs = Vbs; c.Vdb = Vdb;

  const double toxe = M_(toxe), epssub = D_(epssub), coxe = D_(coxe);
  const double phi = S_(phi), sqrtPhi = S_(sqrtPhi), vbsc = I_(vbsc);

  // ---- effective body bias, clamped to [vbsc, 0.95 phi] (:797-817)
  double Vbseff, dVbseff_dVb;
  T0 = Vbs - vbsc - 0.001;
  T1 = sqrt(T0 * T0 - 0.004 * vbsc);
  if (T0 >= 0.0) {
    Vbseff = vbsc + 0.5 * (T0 + T1);
    dVbseff_dVb = 0.5 * (1.0 + B4_DIV(T0, T1));
  } else {
    T2 = B4_DIV(-0.002, (T1 - T0));
    Vbseff = vbsc * (1.0 + T2);
    dVbseff_dVb = B4_DIV(T2 * vbsc, T1);
  }
  T9 = 0.95 * phi;
  T0 = T9 - Vbseff - 0.001;
  T1 = sqrt(T0 * T0 + 0.004 * T9);
  Vbseff = T9 - 0.5 * (T0 + T1);
  dVbseff_dVb *= 0.5 * (1.0 + B4_DIV(T0, T1));
  const double Phis = phi - Vbseff;
  const double sqrtPhis = sqrt(Phis);
  const double dsqrtPhis_dVb = B4_DIV(-0.5, sqrtPhis);
  const double Xdep = B4_DIV(S_(Xdep0) * sqrtPhis, sqrtPhi);
  const double dXdep_dVb = (B4_DIV(S_(Xdep0), sqrtPhi)) * dsqrtPhis_dVb;
  const double Leff = S_(leff), Vtm = D_(vtm), Vtm0 = D_(vtm0);

  // ---- threshold voltage (:826-976)
  T3 = sqrt(Xdep);
  const double V0 = S_(vbi) - phi;
  const double factor1 = D_(factor1);
  T0 = S_(dvt2) * Vbseff;
  if (T0 >= -0.5) { T1 = 1.0 + T0; T2 = S_(dvt2); }
  else { T4 = B4_DIV(1.0, (3.0 + 8.0 * T0)); T1 = (1.0 + 3.0 * T0) * T4; T2 = S_(dvt2) * T4 * T4; }
  const double lt1 = factor1 * T3 * T1;
  const double dlt1_dVb = factor1 * (B4_DIV(0.5, T3) * T1 * dXdep_dVb + T3 * T2);
  T0 = S_(dvt2w) * Vbseff;
  if (T0 >= -0.5) { T1 = 1.0 + T0; T2 = S_(dvt2w); }
  else { T4 = B4_DIV(1.0, (3.0 + 8.0 * T0)); T1 = (1.0 + 3.0 * T0) * T4; T2 = S_(dvt2w) * T4 * T4; }
  const double ltw = factor1 * T3 * T1;
  const double dltw_dVb = factor1 * (B4_DIV(0.5, T3) * T1 * dXdep_dVb + T3 * T2);

  double Theta0, dTheta0_dVb;
  b4_sce_rolloff(B4_DIV(S_(dvt1) * Leff, lt1), dlt1_dVb, lt1, &Theta0, &dTheta0_dVb);
  const double thetavth = S_(dvt0) * Theta0;
  const double Delt_vth = thetavth * V0;
  const double dDelt_vth_dVb = S_(dvt0) * dTheta0_dVb * V0;
  b4_sce_rolloff(B4_DIV(S_(dvt1w) * S_(weff) * Leff, ltw), dltw_dVb, ltw, &T5, &dT5_dVb);
  T0 = S_(dvt0w) * T5;
  T2 = T0 * V0;
  dT2_dVb = S_(dvt0w) * dT5_dVb * V0;

  const double TempRatio = D_(TempRatio);
  T0 = sqrt(1.0 + B4_DIV(S_(lpe0), Leff));
  T1 = S_(k1ox) * (T0 - 1.0) * sqrtPhi + (S_(kt1) + B4_DIV(S_(kt1l), Leff) + S_(kt2) * Vbseff) * (TempRatio - 1.0);
  const double Vth_NarrowW = B4_DIV(toxe * phi, (S_(weff) + S_(w0)));

  T3 = I_(eta0) + S_(etab) * Vbseff;
  if (T3 < 1.0e-4) { T9 = B4_DIV(1.0, (3.0 - 2.0e4 * T3)); T3 = (2.0e-4 - T3) * T9; T4 = T9 * T9; }
  else T4 = 1.0;
  const double dDIBL_Sft_dVd = T3 * S_(theta0vb0);
  const double DIBL_Sft = dDIBL_Sft_dVd * Vds;
  const double Lpe_Vb = sqrt(1.0 + B4_DIV(S_(lpeb), Leff));

  double Vth = tp * I_(vth0) + (S_(k1ox) * sqrtPhis - S_(k1) * sqrtPhi) * Lpe_Vb - I_(k2ox) * Vbseff - Delt_vth - T2
               + (S_(k3) + S_(k3b) * Vbseff) * Vth_NarrowW + T1 - DIBL_Sft;
  double dVth_dVb = Lpe_Vb * S_(k1ox) * dsqrtPhis_dVb - I_(k2ox) - dDelt_vth_dVb - dT2_dVb + S_(k3b) * Vth_NarrowW
                    - S_(etab) * Vds * S_(theta0vb0) * T4 + S_(kt2) * (TempRatio - 1.0);
  double dVth_dVd = -dDIBL_Sft_dVd;

  // subthreshold swing factor n
  double n, dn_dVb, dn_dVd;
  tmp1 = B4_DIV(epssub, Xdep);
  tmp2 = S_(nfactor) * tmp1;
  tmp3 = S_(cdsc) + S_(cdscb) * Vbseff + S_(cdscd) * Vds;
  tmp4 = B4_DIV((tmp2 + tmp3 * Theta0 + S_(cit)), coxe);
  if (tmp4 >= -0.5) {
    n = 1.0 + tmp4;
    dn_dVb = B4_DIV((B4_DIV(-tmp2, Xdep) * dXdep_dVb + tmp3 * dTheta0_dVb + S_(cdscb) * Theta0), coxe);
    dn_dVd = B4_DIV(S_(cdscd) * Theta0, coxe);
  } else {
    T0 = B4_DIV(1.0, (3.0 + 8.0 * tmp4));
    n = (1.0 + 3.0 * tmp4) * T0;
    T0 *= T0;
    dn_dVb = B4_DIV((B4_DIV(-tmp2, Xdep) * dXdep_dVb + tmp3 * dTheta0_dVb + S_(cdscb) * Theta0), coxe) * T0;
    dn_dVd = B4_DIV(S_(cdscd) * Theta0, coxe) * T0;
  }

  if (S_(dvtp0) > 0.0) {  // pocket-implant (DITS) threshold shift
    T0 = -S_(dvtp1) * Vds;
    if (T0 < -B4C_EXP_THRESHOLD) { T2 = B4C_MIN_EXP; dT2_dVd = 0.0; }
    else { T2 = exp(T0); dT2_dVd = -S_(dvtp1) * T2; }
    T3 = Leff + S_(dvtp0) * (1.0 + T2);
    dT3_dVd = S_(dvtp0) * dT2_dVd;
    if (tempmod < 2) { T4 = Vtm * log(B4_DIV(Leff, T3)); dT4_dVd = B4_DIV(-Vtm * dT3_dVd, T3); }
    else { T4 = Vtm0 * log(B4_DIV(Leff, T3)); dT4_dVd = B4_DIV(-Vtm0 * dT3_dVd, T3); }
    const double dDITS_Sft_dVd = dn_dVd * T4 + n * dT4_dVd;
    const double dDITS_Sft_dVb = T4 * dn_dVb;
    Vth -= n * T4;
    dVth_dVd -= dDITS_Sft_dVd;
    dVth_dVb -= dDITS_Sft_dVb;
  }
  if (!(S_(dvtp4) == 0.0 || S_(dvtp2factor) == 0.0)) {  // second DITS term (BSIM4.7)
    T1 = 2.0 * S_(dvtp4) * Vds;
    T0 = b4_dexpb(T1);
    T10 = b4_dexpc(T1);
    const double DITS_Sft2 = B4_DIV(S_(dvtp2factor) * (T0 - 1.0), (T0 + 1.0));
    const double dDITS_Sft2_dVd = B4_DIV(S_(dvtp2factor) * S_(dvtp4) * 4.0 * T10, ((T0 + 1.0) * (T0 + 1.0)));
    Vth -= DITS_Sft2;
    dVth_dVd -= dDITS_Sft2_dVd;
  }
  o.von = Vth;

  // ---- poly depletion and effective gate overdrive (:978-1062)
  T0 = I_(vfb) + phi;
  T1 = mtrlmod == 0 ? B4C_EPSSI : M_(epsrgate) * B4C_EPS0;
  double vgs_eff, dvgs_eff_dvg, vgd_eff, dvgd_eff_dvg;
  b4_poly_depletion(T0, S_(ngate), T1, coxe, v.vgs, &vgs_eff, &dvgs_eff_dvg);
  b4_poly_depletion(T0, S_(ngate), T1, coxe, v.vgd, &vgd_eff, &dvgd_eff_dvg);
  const double Vgs_eff = o.mode > 0 ? vgs_eff : vgd_eff;
  const double dVgs_eff_dVg = o.mode > 0 ? dvgs_eff_dvg : dvgd_eff_dvg;
  const double Vgst = Vgs_eff - Vth;
  const double mstar = S_(mstar);

  T0 = n * Vtm;
  T1 = mstar * Vgst;
  T2 = B4_DIV(T1, T0);
  if (T2 > B4C_EXP_THRESHOLD) {
    T10 = T1;
    dT10_dVg = mstar * dVgs_eff_dVg;
    dT10_dVd = -dVth_dVd * mstar;
    dT10_dVb = -dVth_dVb * mstar;
  } else if (T2 < -B4C_EXP_THRESHOLD) {
    T10 = Vtm * log(1.0 + B4C_MIN_EXP);
    dT10_dVg = 0.0;
    dT10_dVd = T10 * dn_dVd;
    dT10_dVb = T10 * dn_dVb;
    T10 *= n;
  } else {
    const double ExpVgst = exp(T2);
    T3 = Vtm * log(1.0 + ExpVgst);
    T10 = n * T3;
    dT10_dVg = B4_DIV(mstar * ExpVgst, (1.0 + ExpVgst));
    dT10_dVb = T3 * dn_dVb - dT10_dVg * (dVth_dVb + B4_DIV(Vgst * dn_dVb, n));
    dT10_dVd = T3 * dn_dVd - dT10_dVg * (dVth_dVd + B4_DIV(Vgst * dn_dVd, n));
    dT10_dVg *= dVgs_eff_dVg;
  }
  T1 = S_(voffcbn) - (1.0 - mstar) * Vgst;
  T2 = B4_DIV(T1, T0);
  if (T2 < -B4C_EXP_THRESHOLD) {
    T3 = B4_DIV(coxe * B4C_MIN_EXP, S_(cdep0));
    T9 = mstar + T3 * n;
    dT9_dVg = 0.0; dT9_dVd = dn_dVd * T3; dT9_dVb = dn_dVb * T3;
  } else if (T2 > B4C_EXP_THRESHOLD) {
    T3 = B4_DIV(coxe * B4C_MAX_EXP, S_(cdep0));
    T9 = mstar + T3 * n;
    dT9_dVg = 0.0; dT9_dVd = dn_dVd * T3; dT9_dVb = dn_dVb * T3;
  } else {
    const double ExpVgst = exp(T2);
    T3 = B4_DIV(coxe, S_(cdep0));
    T4 = T3 * ExpVgst;
    T5 = B4_DIV(T1 * T4, T0);
    T9 = mstar + n * T4;
    dT9_dVg = B4_DIV(T3 * (mstar - 1.0) * ExpVgst, Vtm);
    dT9_dVb = T4 * dn_dVb - dT9_dVg * dVth_dVb - T5 * dn_dVb;
    dT9_dVd = T4 * dn_dVd - dT9_dVg * dVth_dVd - T5 * dn_dVd;
    dT9_dVg *= dVgs_eff_dVg;
  }
  const double Vgsteff = B4_DIV(T10, T9);
  T11 = T9 * T9;
  const double dVgsteff_dVg = B4_DIV((T9 * dT10_dVg - T10 * dT9_dVg), T11);
  const double dVgsteff_dVd = B4_DIV((T9 * dT10_dVd - T10 * dT9_dVd), T11);
  const double dVgsteff_dVb = B4_DIV((T9 * dT10_dVb - T10 * dT9_dVb), T11);

  // ---- effective width; the reference never enables the internal bias-dependent Rds (:1064-1104)
  T9 = sqrtPhis - sqrtPhi;
  double Weff = S_(weff) - 2.0 * (S_(dwg) * Vgsteff + S_(dwb) * T9);
  double dWeff_dVg = -2.0 * S_(dwg);
  double dWeff_dVb = -2.0 * S_(dwb) * dsqrtPhis_dVb;
  if (Weff < 2.0e-8) {
    T0 = B4_DIV(1.0, (6.0e-8 - 2.0 * Weff));
    Weff = 2.0e-8 * (4.0e-8 - Weff) * T0;
    T0 *= T0 * 4.0e-16;
    dWeff_dVg *= T0;
    dWeff_dVb *= T0;
  }
  const double Rds = 0.0, dRds_dVg = 0.0, dRds_dVb = 0.0;  // `rdsmod > 1` cannot hold: the selector is clamped to {0, 1}

  // ---- bulk charge factor (:1106-1159)
  T9 = B4_DIV(0.5 * S_(k1ox) * Lpe_Vb, sqrtPhis);
  T1 = T9 + I_(k2ox) - S_(k3b) * Vth_NarrowW;
  dT1_dVb = B4_DIV(-T9, sqrtPhis) * dsqrtPhis_dVb;
  T9 = sqrt(S_(xj) * Xdep);
  tmp1 = Leff + 2.0 * T9;
  T5 = B4_DIV(Leff, tmp1);
  tmp2 = S_(a0) * T5;
  tmp3 = S_(weff) + S_(b1);
  tmp4 = B4_DIV(S_(b0), tmp3);
  T2 = tmp2 + tmp4;
  dT2_dVb = B4_DIV(B4_DIV(-T9, tmp1), Xdep) * dXdep_dVb;
  T6 = T5 * T5;
  T7 = T5 * T6;
  double Abulk0 = 1.0 + T1 * T2;
  double dAbulk0_dVb = T1 * tmp2 * dT2_dVb + T2 * dT1_dVb;
  T8 = S_(ags) * S_(a0) * T7;
  double dAbulk_dVg = -T1 * T8;
  double Abulk = Abulk0 + dAbulk_dVg * Vgsteff;
  double dAbulk_dVb = dAbulk0_dVb - T8 * Vgsteff * (dT1_dVb + 3.0 * T1 * dT2_dVb);
  if (Abulk0 < 0.1) {
    T9 = B4_DIV(1.0, (3.0 - 20.0 * Abulk0));
    Abulk0 = (0.2 - Abulk0) * T9;
    dAbulk0_dVb *= T9 * T9;
  }
  if (Abulk < 0.1) {
    T9 = B4_DIV(1.0, (3.0 - 20.0 * Abulk));
    Abulk = (0.2 - Abulk) * T9;
    T10 = T9 * T9;
    dAbulk_dVb *= T10;
    dAbulk_dVg *= T10;
  }
  T2 = S_(keta) * Vbseff;
  if (T2 >= -0.9) { T0 = B4_DIV(1.0, (1.0 + T2)); dT0_dVb = -S_(keta) * T0 * T0; }
  else { T1 = B4_DIV(1.0, (0.8 + T2)); T0 = (17.0 + 20.0 * T2) * T1; dT0_dVb = -S_(keta) * T1 * T1; }
  dAbulk_dVg *= T0;
  dAbulk_dVb = dAbulk_dVb * T0 + Abulk * dT0_dVb;
  dAbulk0_dVb = dAbulk0_dVb * T0 + Abulk0 * dT0_dVb;
  Abulk *= T0;
  Abulk0 *= T0;

  // ---- mobility (:1161-1311)
  double Denomi, dDenomi_dVg, dDenomi_dVd, dDenomi_dVb;
  T14 = (mtrlmod != 0 && mtrlcompatmod == 0) ? 2.0 * tp * (M_(phig) - M_(easub) - 0.5 * D_(Eg0) + 0.45) : 0.0;
  const double ua = S_(ua), ub = S_(ub), uc = S_(uc), ud = S_(ud);
  if (mobmod == 0 || mobmod == 1 || mobmod == 4 || mobmod == 5) {
    // vertical-field mobility: Vth-referenced (0, 1) or (Vth - Vfb - phi)-referenced (4, 5); uc multiplies (1, 5) or adds (0, 4)
    const bool vth_ref = mobmod < 2, mult = mobmod == 1 || mobmod == 5;
    const double Vx = vth_ref ? Vth : I_(vtfbphi1);
    T0 = vth_ref ? Vgsteff + Vth + Vth - T14 : Vgsteff + I_(vtfbphi1) - T14;
    T2 = mult ? 1.0 + uc * Vbseff : ua + uc * Vbseff;
    T3 = B4_DIV(T0, toxe);
    T4 = T3 * (ua + ub * T3);  // used by the multiplicative forms only
    T12 = sqrt(Vx * Vx + 0.0001);
    T9 = B4_DIV(1.0, (Vgsteff + 2.0 * T12));
    T10 = T9 * toxe;
    T8 = ud * T10 * T10 * Vx;
    T6 = T8 * Vx;
    T5 = mult ? T4 * T2 + T6 : T3 * (T2 + ub * T3) + T6;
    T7 = -2.0 * T6 * T9;
    dDenomi_dVg = mult ? B4_DIV((ua + 2.0 * ub * T3) * T2, toxe) : B4_DIV((T2 + 2.0 * ub * T3), toxe);
    if (vth_ref) {
      T11 = B4_DIV(T7 * Vth, T12);
      T13 = 2.0 * (dDenomi_dVg + T11 + T8);
      dDenomi_dVd = T13 * dVth_dVd;
      dDenomi_dVb = T13 * dVth_dVb + (mult ? uc * T4 : uc * T3);
    } else {
      dDenomi_dVd = 0.0;
      dDenomi_dVb = mult ? uc * T4 : uc * T3;
    }
    dDenomi_dVg += T7;
  } else if (mobmod == 2 || mobmod == 6) {
    T0 = B4_DIV((Vgsteff + I_(vtfbphi1)), toxe);
    T1 = exp(S_(eu) * log(T0));
    dT1_dVg = B4_DIV(B4_DIV(T1 * S_(eu), T0), toxe);
    T2 = ua + uc * Vbseff;
    const double Vx = mobmod == 2 ? Vth : I_(vtfbphi1);
    T12 = sqrt(Vx * Vx + 0.0001);
    T9 = B4_DIV(1.0, (Vgsteff + 2.0 * T12));
    T10 = T9 * toxe;
    T8 = ud * T10 * T10 * Vx;
    T6 = T8 * Vx;
    T5 = T1 * T2 + T6;
    T7 = -2.0 * T6 * T9;
    dDenomi_dVg = T2 * dT1_dVg + T7;
    if (mobmod == 2) {
      T11 = B4_DIV(T7 * Vth, T12);
      T13 = 2.0 * (T11 + T8);
      dDenomi_dVd = T13 * dVth_dVd;
      dDenomi_dVb = T13 * dVth_dVb + T1 * uc;
    } else {
      dDenomi_dVd = 0.0;
      dDenomi_dVb = T1 * uc;
    }
  } else {  // mobmod 3: universal + Coulomb scattering for high-k stacks
    T0 = B4_DIV(B4_DIV((Vgsteff + I_(vtfbphi1)) * 1.0e-8, toxe), 6.0);
    T1 = exp(S_(eu) * log(T0));
    dT1_dVg = B4_DIV(B4_DIV(B4_DIV(T1 * S_(eu) * 1.0e-8, T0), toxe), 6.0);
    T2 = ua + uc * Vbseff;
    const double VgsteffVth = S_(VgsteffVth);
    T10 = exp(S_(ucs) * log(0.5 + B4_DIV(0.5 * Vgsteff, VgsteffVth)));
    T11 = B4_DIV(ud, T10);
    const double dT11_dVg = B4_DIV(B4_DIV(-0.5 * S_(ucs) * T11, (0.5 + B4_DIV(0.5 * Vgsteff, VgsteffVth))), VgsteffVth);
    dDenomi_dVg = T2 * dT1_dVg + dT11_dVg;
    dDenomi_dVd = 0.0;
    dDenomi_dVb = T1 * uc;
    T5 = T1 * T2 + T11;
  }
  if (T5 >= -0.8) {
    Denomi = 1.0 + T5;
  } else {
    T9 = B4_DIV(1.0, (7.0 + 10.0 * T5));
    Denomi = (0.6 + T5) * T9;
    T9 *= T9;
    dDenomi_dVg *= T9; dDenomi_dVd *= T9; dDenomi_dVb *= T9;
  }
  const double ueff = B4_DIV(I_(u0temp), Denomi);
  T9 = B4_DIV(-ueff, Denomi);
  const double dueff_dVg = T9 * dDenomi_dVg, dueff_dVd = T9 * dDenomi_dVd, dueff_dVb = T9 * dDenomi_dVb;

  // ---- saturation voltage (:1313-1400)
  const double vsattemp = I_(vsattemp);
  const double WVCox = Weff * vsattemp * coxe;
  const double WVCoxRds = WVCox * Rds;
  double Esat = B4_DIV(2.0 * vsattemp, ueff);
  double EsatL = Esat * Leff;
  T0 = B4_DIV(-EsatL, ueff);
  double dEsatL_dVg = T0 * dueff_dVg, dEsatL_dVd = T0 * dueff_dVd, dEsatL_dVb = T0 * dueff_dVb;

  double Lambda, dLambda_dVg;
  const double a1 = S_(a1), a2 = S_(a2);
  if (a1 == 0.0) {
    Lambda = a2;
    dLambda_dVg = 0.0;
  } else if (a1 > 0.0) {
    T0 = 1.0 - a2;
    T1 = T0 - a1 * Vgsteff - 0.0001;
    T2 = sqrt(T1 * T1 + 0.0004 * T0);
    Lambda = a2 + T0 - 0.5 * (T1 + T2);
    dLambda_dVg = 0.5 * a1 * (1.0 + B4_DIV(T1, T2));
  } else {
    T1 = a2 + a1 * Vgsteff - 0.0001;
    T2 = sqrt(T1 * T1 + 0.0004 * a2);
    Lambda = 0.5 * (T1 + T2);
    dLambda_dVg = 0.5 * a1 * (1.0 + B4_DIV(T1, T2));
  }

  const double Vgst2Vtm = Vgsteff + 2.0 * Vtm;
  if (Rds > 0.0) { tmp2 = B4_DIV(dRds_dVg, Rds) + B4_DIV(dWeff_dVg, Weff); tmp3 = B4_DIV(dRds_dVb, Rds) + B4_DIV(dWeff_dVb, Weff); }
  else { tmp2 = B4_DIV(dWeff_dVg, Weff); tmp3 = B4_DIV(dWeff_dVb, Weff); }
  double Vdsat, dVdsat_dVg, dVdsat_dVd, dVdsat_dVb;
  if (Rds == 0.0 && Lambda == 1.0) {
    T0 = B4_DIV(1.0, (Abulk * EsatL + Vgst2Vtm));
    tmp1 = 0.0;
    T1 = T0 * T0;
    T2 = Vgst2Vtm * T0;
    T3 = EsatL * Vgst2Vtm;
    Vdsat = T3 * T0;
    dT0_dVg = -(Abulk * dEsatL_dVg + EsatL * dAbulk_dVg + 1.0) * T1;
    dT0_dVd = -(Abulk * dEsatL_dVd) * T1;
    dT0_dVb = -(Abulk * dEsatL_dVb + dAbulk_dVb * EsatL) * T1;
    dVdsat_dVg = T3 * dT0_dVg + T2 * dEsatL_dVg + EsatL * T0;
    dVdsat_dVd = T3 * dT0_dVd + T2 * dEsatL_dVd;
    dVdsat_dVb = T3 * dT0_dVb + T2 * dEsatL_dVb;
  } else {
    tmp1 = B4_DIV(dLambda_dVg, (Lambda * Lambda));
    T9 = Abulk * WVCoxRds;
    T8 = Abulk * T9;
    T7 = Vgst2Vtm * T9;
    T6 = Vgst2Vtm * WVCoxRds;
    T0 = 2.0 * Abulk * (T9 - 1.0 + B4_DIV(1.0, Lambda));
    dT0_dVg = 2.0 * (T8 * tmp2 - Abulk * tmp1 + (2.0 * T9 + B4_DIV(1.0, Lambda) - 1.0) * dAbulk_dVg);
    dT0_dVb = 2.0 * (T8 * (B4_DIV(2.0, Abulk) * dAbulk_dVb + tmp3) + (B4_DIV(1.0, Lambda) - 1.0) * dAbulk_dVb);
    dT0_dVd = 0.0;
    T1 = Vgst2Vtm * (B4_DIV(2.0, Lambda) - 1.0) + Abulk * EsatL + 3.0 * T7;
    dT1_dVg = (B4_DIV(2.0, Lambda) - 1.0) - 2.0 * Vgst2Vtm * tmp1 + Abulk * dEsatL_dVg + EsatL * dAbulk_dVg + 3.0 * (T9 + T7 * tmp2 + T6 * dAbulk_dVg);
    dT1_dVb = Abulk * dEsatL_dVb + EsatL * dAbulk_dVb + 3.0 * (T6 * dAbulk_dVb + T7 * tmp3);
    dT1_dVd = Abulk * dEsatL_dVd;
    T2 = Vgst2Vtm * (EsatL + 2.0 * T6);
    dT2_dVg = EsatL + Vgst2Vtm * dEsatL_dVg + T6 * (4.0 + 2.0 * Vgst2Vtm * tmp2);
    dT2_dVb = Vgst2Vtm * (dEsatL_dVb + 2.0 * T6 * tmp3);
    dT2_dVd = Vgst2Vtm * dEsatL_dVd;
    T3 = sqrt(T1 * T1 - 2.0 * T0 * T2);
    Vdsat = B4_DIV((T1 - T3), T0);
    dT3_dVg = B4_DIV((T1 * dT1_dVg - 2.0 * (T0 * dT2_dVg + T2 * dT0_dVg)), T3);
    dT3_dVd = B4_DIV((T1 * dT1_dVd - 2.0 * (T0 * dT2_dVd + T2 * dT0_dVd)), T3);
    dT3_dVb = B4_DIV((T1 * dT1_dVb - 2.0 * (T0 * dT2_dVb + T2 * dT0_dVb)), T3);
    (void)dT3_dVg; (void)dT3_dVd; (void)dT3_dVb;
    dVdsat_dVg = B4_DIV((dT1_dVg - B4_DIV((T1 * dT1_dVg - dT0_dVg * T2 - T0 * dT2_dVg), T3) - Vdsat * dT0_dVg), T0);
    dVdsat_dVb = B4_DIV((dT1_dVb - B4_DIV((T1 * dT1_dVb - dT0_dVb * T2 - T0 * dT2_dVb), T3) - Vdsat * dT0_dVb), T0);
    dVdsat_dVd = B4_DIV((dT1_dVd - B4_DIV((T1 * dT1_dVd - T0 * dT2_dVd), T3)), T0);
  }

  // ---- effective Vds, smoothly limited to Vdsat (:1402-1441)
  const double delta = S_(delta);
  T1 = Vdsat - Vds - delta;
  dT1_dVg = dVdsat_dVg; dT1_dVd = dVdsat_dVd - 1.0; dT1_dVb = dVdsat_dVb;
  T2 = sqrt(T1 * T1 + 4.0 * delta * Vdsat);
  T0 = B4_DIV(T1, T2);
  T9 = 2.0 * delta;
  T3 = B4_DIV(T9, T2);
  dT2_dVg = T0 * dT1_dVg + T3 * dVdsat_dVg;
  dT2_dVd = T0 * dT1_dVd + T3 * dVdsat_dVd;
  dT2_dVb = T0 * dT1_dVb + T3 * dVdsat_dVb;
  double Vdseff, dVdseff_dVg, dVdseff_dVd, dVdseff_dVb;
  if (T1 >= 0.0) {
    Vdseff = Vdsat - 0.5 * (T1 + T2);
    dVdseff_dVg = dVdsat_dVg - 0.5 * (dT1_dVg + dT2_dVg);
    dVdseff_dVd = dVdsat_dVd - 0.5 * (dT1_dVd + dT2_dVd);
    dVdseff_dVb = dVdsat_dVb - 0.5 * (dT1_dVb + dT2_dVb);
  } else {
    T4 = B4_DIV(T9, (T2 - T1));
    T5 = 1.0 - T4;
    T6 = B4_DIV(Vdsat * T4, (T2 - T1));
    Vdseff = Vdsat * T5;
    dVdseff_dVg = dVdsat_dVg * T5 + T6 * (dT2_dVg - dT1_dVg);
    dVdseff_dVd = dVdsat_dVd * T5 + T6 * (dT2_dVd - dT1_dVd);
    dVdseff_dVb = dVdsat_dVb * T5 + T6 * (dT2_dVb - dT1_dVb);
  }
  if (Vds == 0.0) { Vdseff = 0.0; dVdseff_dVg = 0.0; dVdseff_dVb = 0.0; }
  if (Vdseff > Vds) Vdseff = Vds;
  const double diffVds = Vds - Vdseff;

  // ---- velocity overshoot (:1443-1484)
  if (M_(lambda) > 0.0) {
    T1 = Leff * ueff;
    T2 = B4_DIV(S_(lambda), T1);
    T3 = B4_DIV(-T2, T1) * Leff;
    dT2_dVd = T3 * dueff_dVd; dT2_dVg = T3 * dueff_dVg; dT2_dVb = T3 * dueff_dVb;
    T5 = B4_DIV(1.0, (Esat * S_(litl)));
    T4 = B4_DIV(-T5, EsatL);
    dT5_dVg = dEsatL_dVg * T4; dT5_dVd = dEsatL_dVd * T4; dT5_dVb = dEsatL_dVb * T4;
    T6 = 1.0 + diffVds * T5;
    dT6_dVg = dT5_dVg * diffVds - dVdseff_dVg * T5;
    dT6_dVd = dT5_dVd * diffVds + (1.0 - dVdseff_dVd) * T5;
    dT6_dVb = dT5_dVb * diffVds - dVdseff_dVb * T5;
    T7 = B4_DIV(2.0, (T6 * T6 + 1.0));
    T8 = 1.0 - T7;
    T9 = T6 * T7 * T7;
    dT8_dVg = T9 * dT6_dVg; dT8_dVd = T9 * dT6_dVd; dT8_dVb = T9 * dT6_dVb;
    T10 = 1.0 + T2 * T8;
    dT10_dVg = dT2_dVg * T8 + T2 * dT8_dVg;
    dT10_dVd = dT2_dVd * T8 + T2 * dT8_dVd;
    dT10_dVb = dT2_dVb * T8 + T2 * dT8_dVb;
    if (T10 == 1.0) { dT10_dVg = 0.0; dT10_dVd = 0.0; dT10_dVb = 0.0; }
    dEsatL_dVg *= T10; dEsatL_dVg += EsatL * dT10_dVg;
    dEsatL_dVd *= T10; dEsatL_dVd += EsatL * dT10_dVd;
    dEsatL_dVb *= T10; dEsatL_dVb += EsatL * dT10_dVb;
    EsatL *= T10;
    Esat = B4_DIV(EsatL, Leff);
  }

  // ---- Early voltage at saturation (:1486-1506)
  tmp4 = 1.0 - B4_DIV(0.5 * Abulk * Vdsat, Vgst2Vtm);
  T9 = WVCoxRds * Vgsteff;
  T8 = B4_DIV(T9, Vgst2Vtm);
  T0 = EsatL + Vdsat + 2.0 * T9 * tmp4;
  T7 = 2.0 * WVCoxRds * tmp4;
  dT0_dVg = dEsatL_dVg + dVdsat_dVg + T7 * (1.0 + tmp2 * Vgsteff) - T8 * (Abulk * dVdsat_dVg - B4_DIV(Abulk * Vdsat, Vgst2Vtm) + Vdsat * dAbulk_dVg);
  dT0_dVb = dEsatL_dVb + dVdsat_dVb + T7 * tmp3 * Vgsteff - T8 * (dAbulk_dVb * Vdsat + Abulk * dVdsat_dVb);
  dT0_dVd = dEsatL_dVd + dVdsat_dVd - T8 * Abulk * dVdsat_dVd;
  T9 = WVCoxRds * Abulk;
  T1 = B4_DIV(2.0, Lambda) - 1.0 + T9;
  dT1_dVg = -2.0 * tmp1 + WVCoxRds * (Abulk * tmp2 + dAbulk_dVg);
  dT1_dVb = dAbulk_dVb * WVCoxRds + T9 * tmp3;
  const double Vasat = B4_DIV(T0, T1);
  const double dVasat_dVg = B4_DIV((dT0_dVg - Vasat * dT1_dVg), T1);
  const double dVasat_dVb = B4_DIV((dT0_dVb - Vasat * dT1_dVb), T1);
  const double dVasat_dVd = B4_DIV(dT0_dVd, T1);

  // ---- linear-region current with the inversion-layer centroid correction (:1508-1559)
  tmp1 = I_(vtfbphi2);
  tmp2 = 2.0e8 * I_(toxp);
  dT0_dVg = B4_DIV(1.0, tmp2);
  T0 = (Vgsteff + tmp1) * dT0_dVg;
  tmp3 = exp(M_(bdos) * 0.7 * log(T0));
  T1 = 1.0 + tmp3;
  T2 = B4_DIV(M_(bdos) * 0.7 * tmp3, T0);
  const double Tcen = B4_DIV(M_(ados) * 1.9e-9, T1);
  const double dTcen_dVg = B4_DIV(-Tcen * T2 * dT0_dVg, T1);
  const double coxp = I_(coxp);
  const double Coxeff = B4_DIV(epssub * coxp, (epssub + coxp * Tcen));
  const double dCoxeff_dVg = B4_DIV(-Coxeff * Coxeff * dTcen_dVg, epssub);
  const double CoxeffWovL = B4_DIV(Coxeff * Weff, Leff);
  const double beta = ueff * CoxeffWovL;
  T3 = B4_DIV(ueff, Leff);
  const double dbeta_dVg = CoxeffWovL * dueff_dVg + T3 * (Weff * dCoxeff_dVg + Coxeff * dWeff_dVg);
  const double dbeta_dVd = CoxeffWovL * dueff_dVd;
  const double dbeta_dVb = CoxeffWovL * dueff_dVb + T3 * Coxeff * dWeff_dVb;

  const double AbovVgst2Vtm = B4_DIV(Abulk, Vgst2Vtm);
  T0 = 1.0 - 0.5 * Vdseff * AbovVgst2Vtm;
  dT0_dVg = B4_DIV(-0.5 * (Abulk * dVdseff_dVg - B4_DIV(Abulk * Vdseff, Vgst2Vtm) + Vdseff * dAbulk_dVg), Vgst2Vtm);
  dT0_dVd = B4_DIV(-0.5 * Abulk * dVdseff_dVd, Vgst2Vtm);
  dT0_dVb = B4_DIV(-0.5 * (Abulk * dVdseff_dVb + dAbulk_dVb * Vdseff), Vgst2Vtm);
  const double fgche1 = Vgsteff * T0;
  const double dfgche1_dVg = Vgsteff * dT0_dVg + T0, dfgche1_dVd = Vgsteff * dT0_dVd, dfgche1_dVb = Vgsteff * dT0_dVb;
  T9 = B4_DIV(Vdseff, EsatL);
  const double fgche2 = 1.0 + T9;
  const double dfgche2_dVg = B4_DIV((dVdseff_dVg - T9 * dEsatL_dVg), EsatL);
  const double dfgche2_dVd = B4_DIV((dVdseff_dVd - T9 * dEsatL_dVd), EsatL);
  const double dfgche2_dVb = B4_DIV((dVdseff_dVb - T9 * dEsatL_dVb), EsatL);
  const double gche = B4_DIV(beta * fgche1, fgche2);
  const double dgche_dVg = B4_DIV((beta * dfgche1_dVg + fgche1 * dbeta_dVg - gche * dfgche2_dVg), fgche2);
  const double dgche_dVd = B4_DIV((beta * dfgche1_dVd + fgche1 * dbeta_dVd - gche * dfgche2_dVd), fgche2);
  const double dgche_dVb = B4_DIV((beta * dfgche1_dVb + fgche1 * dbeta_dVb - gche * dfgche2_dVb), fgche2);
  T0 = 1.0 + gche * Rds;
  const double Idl = B4_DIV(gche, T0);
  T1 = B4_DIV((1.0 - Idl * Rds), T0);
  T2 = Idl * Idl;
  const double dIdl_dVg = T1 * dgche_dVg - T2 * dRds_dVg;
  const double dIdl_dVd = T1 * dgche_dVd;
  const double dIdl_dVb = T1 * dgche_dVb - T2 * dRds_dVb;

  // ---- output resistance: pocket degradation, CLM, DIBL, DITS, SCBE (:1561-1732)
  double FP, dFP_dVg;
  if (S_(fprout) <= 0.0) { FP = 1.0; dFP_dVg = 0.0; }
  else {
    T9 = B4_DIV(S_(fprout) * sqrt(Leff), Vgst2Vtm);
    FP = B4_DIV(1.0, (1.0 + T9));
    dFP_dVg = B4_DIV(FP * FP * T9, Vgst2Vtm);
  }
  T8 = B4_DIV(S_(pvag), EsatL);
  T9 = T8 * Vgsteff;
  double PvagTerm, dPvagTerm_dVg, dPvagTerm_dVd, dPvagTerm_dVb;
  if (T9 > -0.9) {
    PvagTerm = 1.0 + T9;
    dPvagTerm_dVg = T8 * (1.0 - B4_DIV(Vgsteff * dEsatL_dVg, EsatL));
    dPvagTerm_dVb = B4_DIV(-T9 * dEsatL_dVb, EsatL);
    dPvagTerm_dVd = B4_DIV(-T9 * dEsatL_dVd, EsatL);
  } else {
    T4 = B4_DIV(1.0, (17.0 + 20.0 * T9));
    PvagTerm = (0.8 + T9) * T4;
    T4 *= T4;
    dPvagTerm_dVg = T8 * (1.0 - B4_DIV(Vgsteff * dEsatL_dVg, EsatL)) * T4;
    T9 *= B4_DIV(T4, EsatL);
    dPvagTerm_dVb = -T9 * dEsatL_dVb;
    dPvagTerm_dVd = -T9 * dEsatL_dVd;
  }
  double Cclm, dCclm_dVg, dCclm_dVd, dCclm_dVb, VACLM, dVACLM_dVg, dVACLM_dVd, dVACLM_dVb;
  if (S_(pclm) > B4C_MIN_EXP && diffVds > 1.0e-10) {
    T0 = 1.0 + Rds * Idl;
    dT0_dVg = dRds_dVg * Idl + Rds * dIdl_dVg;
    dT0_dVd = Rds * dIdl_dVd;
    dT0_dVb = dRds_dVb * Idl + Rds * dIdl_dVb;
    T2 = B4_DIV(Vdsat, Esat);
    T1 = Leff + T2;
    dT1_dVg = B4_DIV((dVdsat_dVg - B4_DIV(T2 * dEsatL_dVg, Leff)), Esat);
    dT1_dVd = B4_DIV((dVdsat_dVd - B4_DIV(T2 * dEsatL_dVd, Leff)), Esat);
    dT1_dVb = B4_DIV((dVdsat_dVb - B4_DIV(T2 * dEsatL_dVb, Leff)), Esat);
    Cclm = B4_DIV(FP * PvagTerm * T0 * T1, (S_(pclm) * S_(litl)));
    dCclm_dVg = Cclm * (B4_DIV(dFP_dVg, FP) + B4_DIV(dPvagTerm_dVg, PvagTerm) + B4_DIV(dT0_dVg, T0) + B4_DIV(dT1_dVg, T1));
    dCclm_dVb = Cclm * (B4_DIV(dPvagTerm_dVb, PvagTerm) + B4_DIV(dT0_dVb, T0) + B4_DIV(dT1_dVb, T1));
    dCclm_dVd = Cclm * (B4_DIV(dPvagTerm_dVd, PvagTerm) + B4_DIV(dT0_dVd, T0) + B4_DIV(dT1_dVd, T1));
    VACLM = Cclm * diffVds;
    dVACLM_dVg = dCclm_dVg * diffVds - dVdseff_dVg * Cclm;
    dVACLM_dVb = dCclm_dVb * diffVds - dVdseff_dVb * Cclm;
    dVACLM_dVd = dCclm_dVd * diffVds + (1.0 - dVdseff_dVd) * Cclm;
  } else {
    VACLM = B4C_MAX_EXP; Cclm = B4C_MAX_EXP;
    dVACLM_dVd = 0.0; dVACLM_dVg = 0.0; dVACLM_dVb = 0.0;
    dCclm_dVd = 0.0; dCclm_dVg = 0.0; dCclm_dVb = 0.0;
  }
  double VADIBL, dVADIBL_dVg, dVADIBL_dVd, dVADIBL_dVb;
  if (S_(thetaRout) > B4C_MIN_EXP) {
    T8 = Abulk * Vdsat;
    T0 = Vgst2Vtm * T8;
    dT0_dVg = Vgst2Vtm * Abulk * dVdsat_dVg + T8 + Vgst2Vtm * Vdsat * dAbulk_dVg;
    dT0_dVb = Vgst2Vtm * (dAbulk_dVb * Vdsat + Abulk * dVdsat_dVb);
    dT0_dVd = Vgst2Vtm * Abulk * dVdsat_dVd;
    T1 = Vgst2Vtm + T8;
    dT1_dVg = 1.0 + Abulk * dVdsat_dVg + Vdsat * dAbulk_dVg;
    dT1_dVb = Abulk * dVdsat_dVb + dAbulk_dVb * Vdsat;
    dT1_dVd = Abulk * dVdsat_dVd;
    T9 = T1 * T1;
    T2 = S_(thetaRout);
    VADIBL = B4_DIV((Vgst2Vtm - B4_DIV(T0, T1)), T2);
    dVADIBL_dVg = B4_DIV((1.0 - B4_DIV(dT0_dVg, T1) + B4_DIV(T0 * dT1_dVg, T9)), T2);
    dVADIBL_dVb = B4_DIV((B4_DIV(-dT0_dVb, T1) + B4_DIV(T0 * dT1_dVb, T9)), T2);
    dVADIBL_dVd = B4_DIV((B4_DIV(-dT0_dVd, T1) + B4_DIV(T0 * dT1_dVd, T9)), T2);
    T7 = S_(pdiblb) * Vbseff;
    if (T7 >= -0.9) {
      T3 = B4_DIV(1.0, (1.0 + T7));
      VADIBL *= T3;
      dVADIBL_dVg *= T3;
      dVADIBL_dVb = (dVADIBL_dVb - VADIBL * S_(pdiblb)) * T3;
      dVADIBL_dVd *= T3;
    } else {
      T4 = B4_DIV(1.0, (0.8 + T7));
      T3 = (17.0 + 20.0 * T7) * T4;
      dVADIBL_dVg *= T3;
      dVADIBL_dVb = dVADIBL_dVb * T3 - VADIBL * S_(pdiblb) * T4 * T4;
      dVADIBL_dVd *= T3;
      VADIBL *= T3;
    }
    dVADIBL_dVg = dVADIBL_dVg * PvagTerm + VADIBL * dPvagTerm_dVg;
    dVADIBL_dVb = dVADIBL_dVb * PvagTerm + VADIBL * dPvagTerm_dVb;
    dVADIBL_dVd = dVADIBL_dVd * PvagTerm + VADIBL * dPvagTerm_dVd;
    VADIBL *= PvagTerm;
  } else {
    VADIBL = B4C_MAX_EXP;
    dVADIBL_dVd = 0.0; dVADIBL_dVg = 0.0; dVADIBL_dVb = 0.0;
  }
  const double Va = Vasat + VACLM;
  const double dVa_dVg = dVasat_dVg + dVACLM_dVg, dVa_dVb = dVasat_dVb + dVACLM_dVb, dVa_dVd = dVasat_dVd + dVACLM_dVd;

  double VADITS, dVADITS_dVg, dVADITS_dVd;
  T0 = S_(pditsd) * Vds;
  if (T0 > B4C_EXP_THRESHOLD) { T1 = B4C_MAX_EXP; dT1_dVd = 0.0; }
  else { T1 = exp(T0); dT1_dVd = T1 * S_(pditsd); }
  if (S_(pdits) > B4C_MIN_EXP) {
    T2 = 1.0 + M_(pditsl) * Leff;
    VADITS = B4_DIV((1.0 + T2 * T1), S_(pdits));
    dVADITS_dVg = VADITS * dFP_dVg;
    dVADITS_dVd = B4_DIV(FP * T2 * dT1_dVd, S_(pdits));
    VADITS *= FP;
  } else {
    VADITS = B4C_MAX_EXP; dVADITS_dVg = 0.0; dVADITS_dVd = 0.0;
  }
  double VASCBE = B4C_MAX_EXP, dVASCBE_dVg = 0.0, dVASCBE_dVd = 0.0, dVASCBE_dVb = 0.0;
  if (S_(pscbe2) > 0.0 && S_(pscbe1) >= 0.0) {
    if (diffVds > B4_DIV(S_(pscbe1) * S_(litl), B4C_EXP_THRESHOLD)) {
      T0 = B4_DIV(S_(pscbe1) * S_(litl), diffVds);
      VASCBE = B4_DIV(Leff * exp(T0), S_(pscbe2));
      T1 = B4_DIV(T0 * VASCBE, diffVds);
      dVASCBE_dVg = T1 * dVdseff_dVg;
      dVASCBE_dVd = -T1 * (1.0 - dVdseff_dVd);
      dVASCBE_dVb = T1 * dVdseff_dVb;
    } else {
      VASCBE = B4_DIV(B4C_MAX_EXP * Leff, S_(pscbe2));
    }
  }

  // ---- assemble Ids/Vdseff with DIBL, DITS and CLM (:1734-1764)
  T9 = B4_DIV(diffVds, VADIBL);
  T0 = 1.0 + T9;
  double Idsa = Idl * T0;
  double dIdsa_dVg = T0 * dIdl_dVg - B4_DIV(Idl * (dVdseff_dVg + T9 * dVADIBL_dVg), VADIBL);
  double dIdsa_dVd = T0 * dIdl_dVd + B4_DIV(Idl * (1.0 - dVdseff_dVd - T9 * dVADIBL_dVd), VADIBL);
  double dIdsa_dVb = T0 * dIdl_dVb - B4_DIV(Idl * (dVdseff_dVb + T9 * dVADIBL_dVb), VADIBL);
  T9 = B4_DIV(diffVds, VADITS);
  T0 = 1.0 + T9;
  dIdsa_dVg = T0 * dIdsa_dVg - B4_DIV(Idsa * (dVdseff_dVg + T9 * dVADITS_dVg), VADITS);
  dIdsa_dVd = T0 * dIdsa_dVd + B4_DIV(Idsa * (1.0 - dVdseff_dVd - T9 * dVADITS_dVd), VADITS);
  dIdsa_dVb = T0 * dIdsa_dVb - B4_DIV(Idsa * dVdseff_dVb, VADITS);
  Idsa *= T0;
  T0 = log(B4_DIV(Va, Vasat));
  dT0_dVg = B4_DIV(dVa_dVg, Va) - B4_DIV(dVasat_dVg, Vasat);
  dT0_dVb = B4_DIV(dVa_dVb, Va) - B4_DIV(dVasat_dVb, Vasat);
  dT0_dVd = B4_DIV(dVa_dVd, Va) - B4_DIV(dVasat_dVd, Vasat);
  T1 = B4_DIV(T0, Cclm);
  T9 = 1.0 + T1;
  dT9_dVg = B4_DIV((dT0_dVg - T1 * dCclm_dVg), Cclm);
  dT9_dVb = B4_DIV((dT0_dVb - T1 * dCclm_dVb), Cclm);
  dT9_dVd = B4_DIV((dT0_dVd - T1 * dCclm_dVd), Cclm);
  dIdsa_dVg = dIdsa_dVg * T9 + Idsa * dT9_dVg;
  dIdsa_dVb = dIdsa_dVb * T9 + Idsa * dT9_dVb;
  dIdsa_dVd = dIdsa_dVd * T9 + Idsa * dT9_dVd;
  Idsa *= T9;

  // ---- impact-ionisation substrate current (:1766-1803)
  double Isub, Gbd, Gbb, Gbg;
  tmp = S_(alpha0) + S_(alpha1) * Leff;
  if (tmp <= 0.0 || S_(beta0) <= 0.0) {
    Isub = 0.0; Gbd = 0.0; Gbb = 0.0; Gbg = 0.0;
  } else {
    T2 = B4_DIV(tmp, Leff);
    if (diffVds > B4_DIV(S_(beta0), B4C_EXP_THRESHOLD)) {
      T0 = B4_DIV(-S_(beta0), diffVds);
      T1 = T2 * diffVds * exp(T0);
      T3 = B4_DIV(T1, diffVds) * (T0 - 1.0);
      dT1_dVg = T3 * dVdseff_dVg;
      dT1_dVd = T3 * (dVdseff_dVd - 1.0);
      dT1_dVb = T3 * dVdseff_dVb;
    } else {
      T3 = T2 * B4C_MIN_EXP;
      T1 = T3 * diffVds;
      dT1_dVg = -T3 * dVdseff_dVg;
      dT1_dVd = T3 * (1.0 - dVdseff_dVd);
      dT1_dVb = -T3 * dVdseff_dVb;
    }
    T4 = Idsa * Vdseff;
    Isub = T1 * T4;
    Gbg = T1 * (dIdsa_dVg * Vdseff + Idsa * dVdseff_dVg) + T4 * dT1_dVg;
    Gbd = T1 * (dIdsa_dVd * Vdseff + Idsa * dVdseff_dVd) + T4 * dT1_dVd;
    Gbb = T1 * (dIdsa_dVb * Vdseff + Idsa * dVdseff_dVb) + T4 * dT1_dVb;
    Gbd += Gbg * dVgsteff_dVd;
    Gbb += Gbg * dVgsteff_dVb;
    Gbg *= dVgsteff_dVg;
    Gbb *= dVbseff_dVb;
  }
  o.csub = Isub; o.gbbs = Gbb; o.gbgs = Gbg; o.gbds = Gbd;

  // ---- SCBE, chain rule to terminal voltages, drain current (:1805-1868)
  T9 = B4_DIV(diffVds, VASCBE);
  T0 = 1.0 + T9;
  const double Ids = Idsa * T0;
  double Gm = T0 * dIdsa_dVg - B4_DIV(Idsa * (dVdseff_dVg + T9 * dVASCBE_dVg), VASCBE);
  double Gds = T0 * dIdsa_dVd + B4_DIV(Idsa * (1.0 - dVdseff_dVd - T9 * dVASCBE_dVd), VASCBE);
  double Gmb = T0 * dIdsa_dVb - B4_DIV(Idsa * (dVdseff_dVb + T9 * dVASCBE_dVb), VASCBE);
  tmp1 = Gds + Gm * dVgsteff_dVd;
  tmp2 = Gmb + Gm * dVgsteff_dVb;
  tmp3 = Gm;
  Gm = (Ids * dVdseff_dVg + Vdseff * tmp3) * dVgsteff_dVg;
  Gds = Ids * (dVdseff_dVd + dVdseff_dVg * dVgsteff_dVd) + Vdseff * tmp1;
  Gmb = (Ids * (dVdseff_dVb + dVdseff_dVg * dVgsteff_dVb) + Vdseff * tmp2) * dVbseff_dVb;
  double cdrain = Ids * Vdseff;

  if (M_(vtl_given) != 0.0 && M_(vtl) > 0.0) {  // source-end velocity limit
    T12 = B4_DIV(B4_DIV(1.0, Leff), CoxeffWovL);
    T11 = B4_DIV(T12, Vgsteff);
    T10 = B4_DIV(-T11, Vgsteff);
    const double vs = cdrain * T11;
    const double dvs_dVg = Gm * T11 + cdrain * T10 * dVgsteff_dVg;
    const double dvs_dVd = Gds * T11 + cdrain * T10 * dVgsteff_dVd;
    const double dvs_dVb = Gmb * T11 + cdrain * T10 * dVgsteff_dVb;
    T0 = 6.0;
    T1 = B4_DIV(vs, (S_(vtl) * S_(tfactor)));
    if (T1 > 0.0) {
      T2 = 1.0 + exp(T0 * log(T1));
      T3 = B4_DIV((T2 - 1.0) * T0, vs);
      const double Fsevl = B4_DIV(1.0, exp(B4_DIV(log(T2), T0)));
      T4 = B4_DIV(B4_DIV(-1.0, T0) * Fsevl, T2);
      const double dFsevl_dVg = T4 * (T3 * dvs_dVg), dFsevl_dVd = T4 * (T3 * dvs_dVd), dFsevl_dVb = T4 * (T3 * dvs_dVb);
      Gm *= Fsevl;  Gm += cdrain * dFsevl_dVg;
      Gmb *= Fsevl; Gmb += cdrain * dFsevl_dVb;
      Gds *= Fsevl; Gds += cdrain * dFsevl_dVd;
      cdrain *= Fsevl;
    }
  }
  o.gds = Gds; o.gm = Gm; o.gmbs = Gmb;

  // ---- intrinsic-gate input resistance (:1870-1900)
  o.gcrg = 0.0; o.gcrgd = 0.0; o.gcrgg = 0.0; o.gcrgs = 0.0; o.gcrgb = 0.0;
  if (rgatemod > 1 || M_(trnqsmod) != 0.0 || I_(acnqsmod) != 0.0) {
    T9 = S_(xrcrg2) * Vtm;
    T0 = T9 * beta;
    dT0_dVd = (dbeta_dVd + dbeta_dVg * dVgsteff_dVd) * T9;
    dT0_dVb = (dbeta_dVb + dbeta_dVg * dVgsteff_dVb) * T9;
    dT0_dVg = dbeta_dVg * T9;
    o.gcrg = S_(xrcrg1) * (T0 + Ids);
    o.gcrgd = S_(xrcrg1) * (dT0_dVd + tmp1);
    o.gcrgb = S_(xrcrg1) * (dT0_dVb + tmp2) * dVbseff_dVb;
    o.gcrgg = S_(xrcrg1) * (dT0_dVg + tmp3) * dVgsteff_dVg;
    const double nf = I_(nf);
    if (nf != 1.0) { o.gcrg *= nf; o.gcrgg *= nf; o.gcrgd *= nf; o.gcrgb *= nf; }
    if (rgatemod == 2) {
      const double grgeltd = I_(grgeltd);
      T10 = grgeltd * grgeltd;
      T11 = grgeltd + o.gcrg;
      o.gcrg = B4_DIV(grgeltd * o.gcrg, T11);
      T12 = B4_DIV(B4_DIV(T10, T11), T11);
      o.gcrgg *= T12; o.gcrgd *= T12; o.gcrgb *= T12;
    }
    o.gcrgs = -(o.gcrgg + o.gcrgd + o.gcrgb);
  }

  // ---- bias-dependent external source/drain resistance (:1902-1989)
  if (rdsmod != 0) {
    const double prwg = S_(prwg), prwb = S_(prwb), vfbsd = S_(vfbsd);
    auto end_res = [&](double vg, double vb, double r0, double rmin, double gend, double* gtot, double* dg_dvg, double* dg_dvb) {
      double t0 = vg - vfbsd;
      double t1 = sqrt(t0 * t0 + 1.0e-4);
      const double vg_eff = 0.5 * (t0 + t1);
      const double dvg_eff = B4_DIV(vg_eff, t1);
      t0 = 1.0 + prwg * vg_eff;
      const double dt0_dvg = B4_DIV(B4_DIV(-prwg, t0), t0) * dvg_eff;
      t1 = -prwb * vb;
      const double dt1_dvb = -prwb;
      const double t2 = B4_DIV(1.0, t0) + t1;
      const double t3 = t2 + sqrt(t2 * t2 + 0.01);
      const double dt3_dvg = B4_DIV(t3, (t3 - t2));
      const double dt3_dvb = dt3_dvg * dt1_dvb;
      const double dt3_dvg2 = dt3_dvg * dt0_dvg;
      const double t4 = r0 * 0.5;
      const double R = rmin + t3 * t4;
      const double dR_dvg = t4 * dt3_dvg2, dR_dvb = t4 * dt3_dvb;
      t0 = 1.0 + gend * R;
      *gtot = B4_DIV(gend, t0);
      t0 = -*gtot * *gtot;
      *dg_dvg = t0 * dR_dvg;
      *dg_dvb = t0 * dR_dvb;
    };
    double dgstot_dvg, dgstot_dvb, dgdtot_dvg, dgdtot_dvb;
    end_res(v.vgs, v.vbs, S_(rs0), S_(rswmin), I_(sourceConductance), &o.gstot, &dgstot_dvg, &dgstot_dvb);
    const double dgstot_dvd = 0.0;
    const double dgstot_dvs = -(dgstot_dvg + dgstot_dvb + dgstot_dvd);
    end_res(v.vgd, v.vbd, S_(rd0), S_(rdwmin), I_(drainConductance), &o.gdtot, &dgdtot_dvg, &dgdtot_dvb);
    const double dgdtot_dvs = 0.0;
    const double dgdtot_dvd = -(dgdtot_dvg + dgdtot_dvb + dgdtot_dvs);
    o.gstotd = v.vses * dgstot_dvd; o.gstotg = v.vses * dgstot_dvg; o.gstots = v.vses * dgstot_dvs; o.gstotb = v.vses * dgstot_dvb;
    T2 = v.vdes - v.vds;
    o.gdtotd = T2 * dgdtot_dvd; o.gdtotg = T2 * dgdtot_dvg; o.gdtots = T2 * dgdtot_dvs; o.gdtotb = T2 * dgdtot_dvb;
  } else {
    o.gstot = 0.0; o.gstotd = 0.0; o.gstotg = 0.0; o.gstots = 0.0; o.gstotb = 0.0;
    o.gdtot = 0.0; o.gdtotd = 0.0; o.gdtotg = 0.0; o.gdtots = 0.0; o.gdtotb = 0.0;
  }

  c.Vbseff = Vbseff; c.dVbseff_dVb = dVbseff_dVb; c.Phis = Phis; c.sqrtPhis = sqrtPhis; c.dsqrtPhis_dVb = dsqrtPhis_dVb;
  c.Xdep = Xdep; c.dXdep_dVb = dXdep_dVb;
  c.Vth = Vth; c.dVth_dVb = dVth_dVb; c.dVth_dVd = dVth_dVd; c.n = n; c.dn_dVb = dn_dVb; c.dn_dVd = dn_dVd;
  c.vgs_eff = vgs_eff; c.vgd_eff = vgd_eff; c.dvgs_eff_dvg = dvgs_eff_dvg; c.dvgd_eff_dvg = dvgd_eff_dvg;
  c.Vgs_eff = Vgs_eff; c.dVgs_eff_dVg = dVgs_eff_dVg; c.Vgst = Vgst;
  c.Vgsteff = Vgsteff; c.dVgsteff_dVg = dVgsteff_dVg; c.dVgsteff_dVd = dVgsteff_dVd; c.dVgsteff_dVb = dVgsteff_dVb;
  c.Weff = Weff; c.Abulk = Abulk; c.Abulk0 = Abulk0; c.dAbulk0_dVb = dAbulk0_dVb; c.dAbulk_dVg = dAbulk_dVg; c.dAbulk_dVb = dAbulk_dVb;
  c.AbovVgst2Vtm = AbovVgst2Vtm;
  c.ueff = ueff; c.EsatL = EsatL; c.Vdsat = Vdsat; c.dVdsat_dVg = dVdsat_dVg; c.dVdsat_dVd = dVdsat_dVd; c.dVdsat_dVb = dVdsat_dVb;
  c.Vdseff = Vdseff; c.dVdseff_dVg = dVdseff_dVg; c.dVdseff_dVd = dVdseff_dVd; c.dVdseff_dVb = dVdseff_dVb;
  c.Coxeff = Coxeff; c.dCoxeff_dVg = dCoxeff_dVg; c.Tcen = Tcen; c.dTcen_dVg = dTcen_dVg; c.thetavth = thetavth;
  c.Vtm = Vtm; c.Vtm0 = Vtm0; c.Leff = Leff; c.Lpe_Vb = Lpe_Vb; c.Vth_NarrowW = Vth_NarrowW;
  c.cdrain = cdrain;
  (void)T13; (void)T14; (void)Esat;
}

}  // namespace b4e
}  // namespace s21

#include "bsim4_eval_leak.hpp"
