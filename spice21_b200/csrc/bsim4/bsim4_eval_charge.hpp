// BSIM4 evaluation, phase 5: terminal charges and their capacitance matrix (intrinsic charge model for capmod 0/1/2,
// S/D junction depletion charge, gate overlap charge) and the partition of the total charge onto the terminals.
// Follows bsim4solver.rs:2642-3746 (see bsim4_eval.hpp header).
#pragma once

namespace s21 {
namespace b4e {

// Smoothed (Vgst - voff)/(n vt) -> Vgsteff used by the charge model when cvchargemod == 1 (same construction as the
// I-V one, with the C-V specific mstarcv / voffcbncv). Kept separate: the I-V version keeps extra state.
struct B4VgsteffCV { double v, dVg, dVd, dVb; };

// depletion charge and capacitance of one junction component (bottom, sidewall or gate-edge sidewall), reverse bias
B4_HD void b4_jct_component(double cz, double phi, double mj, double vj, bool first, double* q, double* cap) {
  if (cz > 0.0) {
    const double arg = 1.0 - B4_DIV(vj, phi);
    const double sarg = mj == 0.5 ? B4_DIV(1.0, sqrt(arg)) : exp(-mj * log(arg));
    if (first) { *q = B4_DIV(phi * cz * (1.0 - arg * sarg), (1.0 - mj)); *cap = cz * sarg; }
    else { *q += B4_DIV(phi * cz * (1.0 - arg * sarg), (1.0 - mj)); *cap += cz * sarg; }
  } else if (first) {
    *q = 0.0;
    *cap = 0.0;
  }
}
B4_HD void b4_junction_charge(double vj, double cz, double czsw, double czswg, double phi, double phisw, double phiswg, double mj, double mjsw,
                              double mjswg, double* q, double* cap) {
  if (vj == 0.0) {
    *q = 0.0;
    *cap = cz + czsw + czswg;
  } else if (vj < 0.0) {
    b4_jct_component(cz, phi, mj, vj, true, q, cap);
    b4_jct_component(czsw, phisw, mjsw, vj, false, q, cap);
    b4_jct_component(czswg, phiswg, mjswg, vj, false, q, cap);
  } else {
    const double T0 = cz + czsw + czswg;
    const double T1 = vj * (B4_DIV(cz * mj, phi) + B4_DIV(czsw * mjsw, phisw) + B4_DIV(czswg * mjswg, phiswg));
    *q = vj * (T0 + 0.5 * T1);
    *cap = T0 + T1;
  }
}
// bias-dependent overlap capacitance and charge of one gate edge (capmod != 0)
B4_HD void b4_overlap(double vg, double cov, double weffCV, double cl, double ckappa, double* c, double* q) {
  const double T0 = vg + B4C_DELTA_1;
  const double T1 = sqrt(T0 * T0 + 4.0 * B4C_DELTA_1);
  const double T2 = 0.5 * (T0 - T1);
  const double T3 = weffCV * cl;
  const double T4 = sqrt(1.0 - B4_DIV(4.0 * T2, ckappa));
  *c = cov + T3 - T3 * (1.0 - B4_DIV(1.0, T4)) * (0.5 - B4_DIV(0.5 * T0, T1));
  *q = (cov + T3) * vg - T3 * (T2 + 0.5 * ckappa * (T4 - 1.0));
}

template <class E> B4_HD void b4_charge(E& e, const B4Bias& v, B4Op& o, const B4Chan& c, const B4Tunnel& tun) {
  const int capmod = (int)M_(capmod), trnqsmod = (int)M_(trnqsmod), rgatemod = (int)M_(rgatemod), rbodymod = (int)M_(rbodymod);
  const double xpart = M_(xpart), nf = I_(nf), coxe = D_(coxe), phi = S_(phi), k1ox = S_(k1ox);
  const double Vds = c.Vds, Vbs = c.Vbs, Vgs_eff = c.Vgs_eff, dVgs_eff_dVg = c.dVgs_eff_dVg;
  const double Vbseff = c.Vbseff, dVbseff_dVb = c.dVbseff_dVb, Phis = c.Phis, sqrtPhis = c.sqrtPhis, dsqrtPhis_dVb = c.dsqrtPhis_dVb;
  const double dPhis_dVb = -1.0;
  const double Vtm = c.Vtm, epssub = D_(epssub);
  const double Abulk0 = c.Abulk0, dAbulk0_dVb = c.dAbulk0_dVb, abulkCVfactor = S_(abulkCVfactor);
  double T0, T1, T2, T3, T4, T5, T6, T7, T8, T9, T10, T11, T12, tmp, tmp1;
  double dT0_dVg, dT0_dVb, dT1_dVg, dT1_dVd, dT1_dVb;
  double qgate = 0.0, qbulk = 0.0, qdrn = 0.0, qsrc = 0.0, qgmid = 0.0;
  (void)tun;

  o.qgate = 0.0; o.qbulk = 0.0; o.qdrn = 0.0;
  o.qchqs = 0.0; o.qcheq = 0.0; o.cqgb = 0.0; o.cqdb = 0.0; o.cqsb = 0.0; o.cqbb = 0.0;
  if (xpart < 0.0) {
    o.cggb = 0.0; o.cgsb = 0.0; o.cgdb = 0.0; o.cdgb = 0.0; o.cdsb = 0.0; o.cddb = 0.0; o.cbgb = 0.0; o.cbsb = 0.0; o.cbdb = 0.0;
  } else {
    const double CoxWL = coxe * S_(weffCV) * S_(leffCV) * nf;
    if (capmod == 0) {
      // ---- piecewise long-channel charge model (:2699-2961)
      double VbseffCV, dVbseffCV_dVb;
      if (Vbseff < 0.0) { VbseffCV = Vbs; dVbseffCV_dVb = 1.0; }
      else { VbseffCV = phi - Phis; dVbseffCV_dVb = -dPhis_dVb * dVbseff_dVb; }
      const double Vfb = S_(vfbcv);
      const double Vth = Vfb + phi + k1ox * sqrtPhis;
      const double Vgst = Vgs_eff - Vth;
      const double dVth_dVb = k1ox * dsqrtPhis_dVb * dVbseff_dVb;
      const double Arg1 = Vgs_eff - VbseffCV - Vfb;
      if (Arg1 <= 0.0) {  // accumulation
        qgate = CoxWL * Arg1;
        qbulk = -qgate;
        qdrn = 0.0;
        o.cggb = CoxWL * dVgs_eff_dVg;
        o.cgdb = 0.0;
        o.cgsb = CoxWL * (dVbseffCV_dVb - dVgs_eff_dVg);
        o.cdgb = 0.0; o.cddb = 0.0; o.cdsb = 0.0;
        o.cbgb = -CoxWL * dVgs_eff_dVg;
        o.cbdb = 0.0;
        o.cbsb = -o.cgsb;
      } else if (Vgst <= 0.0) {  // depletion
        T1 = 0.5 * k1ox;
        T2 = sqrt(T1 * T1 + Arg1);
        qgate = CoxWL * k1ox * (T2 - T1);
        qbulk = -qgate;
        qdrn = 0.0;
        T0 = B4_DIV(CoxWL * T1, T2);
        o.cggb = T0 * dVgs_eff_dVg;
        o.cgdb = 0.0;
        o.cgsb = T0 * (dVbseffCV_dVb - dVgs_eff_dVg);
        o.cdgb = 0.0; o.cddb = 0.0; o.cdsb = 0.0;
        o.cbgb = -o.cggb;
        o.cbdb = 0.0;
        o.cbsb = -o.cgsb;
      } else {  // inversion
        const double One_Third_CoxWL = B4_DIV(CoxWL, 3.0);
        const double Two_Third_CoxWL = 2.0 * One_Third_CoxWL;
        const double AbulkCV = Abulk0 * abulkCVfactor;
        const double dAbulkCV_dVb = abulkCVfactor * dAbulk0_dVb * dVbseff_dVb;
        const double dVdsat_dVg = B4_DIV(1.0, AbulkCV);
        const double Vdsat = Vgst * dVdsat_dVg;
        const double dVdsat_dVb = -(Vdsat * dAbulkCV_dVb + dVth_dVb) * dVdsat_dVg;
        const bool saturated = xpart > 0.5 ? Vdsat <= Vds : Vds >= Vdsat;
        if (saturated) {
          T1 = B4_DIV(Vdsat, 3.0);
          qgate = CoxWL * (Vgs_eff - Vfb - phi - T1);
          T2 = -Two_Third_CoxWL * Vgst;
          qbulk = -(qgate + T2);
          o.cggb = One_Third_CoxWL * (3.0 - dVdsat_dVg) * dVgs_eff_dVg;
          if (xpart > 0.5) qdrn = 0.0;
          else if (xpart < 0.5) qdrn = 0.4 * T2;
          else qdrn = 0.5 * T2;
          T2 = -One_Third_CoxWL * dVdsat_dVb;
          o.cgsb = -(o.cggb + T2);
          o.cgdb = 0.0;
          if (xpart > 0.5) {
            o.cdgb = 0.0; o.cddb = 0.0; o.cdsb = 0.0;
          } else if (xpart < 0.5) {
            T3 = 0.4 * Two_Third_CoxWL;
            o.cdgb = -T3 * dVgs_eff_dVg;
            o.cddb = 0.0;
            T4 = T3 * dVth_dVb;
            o.cdsb = -(T4 + o.cdgb);
          } else {
            o.cdgb = -One_Third_CoxWL * dVgs_eff_dVg;
            o.cddb = 0.0;
            T4 = One_Third_CoxWL * dVth_dVb;
            o.cdsb = -(T4 + o.cdgb);
          }
          o.cbgb = -(o.cggb - Two_Third_CoxWL * dVgs_eff_dVg);
          T3 = -(T2 + Two_Third_CoxWL * dVth_dVb);
          o.cbsb = -(o.cbgb + T3);
          o.cbdb = 0.0;
        } else {  // linear region
          const double Alphaz = B4_DIV(Vgst, Vdsat);
          T1 = 2.0 * Vdsat - Vds;
          T2 = B4_DIV(Vds, (3.0 * T1));
          T3 = T2 * Vds;
          T9 = 0.25 * CoxWL;
          T4 = T9 * Alphaz;
          qgate = CoxWL * (Vgs_eff - Vfb - phi - 0.5 * (Vds - T3));
          T5 = B4_DIV(T3, T1);
          o.cggb = CoxWL * (1.0 - T5 * dVdsat_dVg) * dVgs_eff_dVg;
          o.cgdb = CoxWL * (T2 - 0.5 + 0.5 * T5);
          if (xpart > 0.5) {  // 0/100
            T7 = 2.0 * Vds - T1 - 3.0 * T3;
            T8 = T3 - T1 - 2.0 * Vds;
            T10 = T4 * T8;
            qdrn = T4 * T7;
            qbulk = -(qgate + qdrn + T10);
            T11 = -CoxWL * T5 * dVdsat_dVb;
            o.cgsb = -(o.cggb + T11 + o.cgdb);
            T6 = B4_DIV(1.0, Vdsat);
            const double dAlphaz_dVg = T6 * (1.0 - Alphaz * dVdsat_dVg);
            const double dAlphaz_dVb = -T6 * (dVth_dVb + Alphaz * dVdsat_dVb);
            T7 = T9 * T7;
            T8 = T9 * T8;
            T9 = 2.0 * T4 * (1.0 - 3.0 * T5);
            o.cdgb = (T7 * dAlphaz_dVg - T9 * dVdsat_dVg) * dVgs_eff_dVg;
            T12 = T7 * dAlphaz_dVb - T9 * dVdsat_dVb;
            o.cddb = T4 * (3.0 - 6.0 * T2 - 3.0 * T5);
            o.cdsb = -(o.cdgb + T12 + o.cddb);
            T9 = 2.0 * T4 * (1.0 + T5);
            T10 = (T8 * dAlphaz_dVg - T9 * dVdsat_dVg) * dVgs_eff_dVg;
            T11 = T8 * dAlphaz_dVb - T9 * dVdsat_dVb;
            T12 = T4 * (2.0 * T2 + T5 - 1.0);
            T0 = -(T10 + T11 + T12);
            o.cbgb = -(o.cggb + o.cdgb + T10);
            o.cbdb = -(o.cgdb + o.cddb + T12);
            o.cbsb = -(o.cgsb + o.cdsb + T0);
          } else if (xpart < 0.5) {  // 40/60
            tmp = -CoxWL * T5 * dVdsat_dVb;
            o.cgsb = -(o.cggb + o.cgdb + tmp);
            T6 = B4_DIV(1.0, Vdsat);
            const double dAlphaz_dVg = T6 * (1.0 - Alphaz * dVdsat_dVg);
            const double dAlphaz_dVb = -T6 * (dVth_dVb + Alphaz * dVdsat_dVb);
            T6 = 8.0 * Vdsat * Vdsat - 6.0 * Vdsat * Vds + 1.2 * Vds * Vds;
            T8 = B4_DIV(T2, T1);
            T7 = Vds - T1 - T8 * T6;
            qdrn = T4 * T7;
            T7 *= T9;
            tmp = B4_DIV(T8, T1);
            tmp1 = T4 * (2.0 - 4.0 * tmp * T6 + T8 * (16.0 * Vdsat - 6.0 * Vds));
            o.cdgb = (T7 * dAlphaz_dVg - tmp1 * dVdsat_dVg) * dVgs_eff_dVg;
            T10 = T7 * dAlphaz_dVb - tmp1 * dVdsat_dVb;
            o.cddb = T4 * (2.0 - (B4_DIV(1.0, (3.0 * T1 * T1)) + 2.0 * tmp) * T6 + T8 * (6.0 * Vdsat - 2.4 * Vds));
            o.cdsb = -(o.cdgb + T10 + o.cddb);
            T7 = 2.0 * (T1 + T3);
            qbulk = -(qgate - T4 * T7);
            T7 *= T9;
            T0 = 4.0 * T4 * (1.0 - T5);
            T12 = (-T7 * dAlphaz_dVg - T0 * dVdsat_dVg) * dVgs_eff_dVg - o.cdgb;
            T11 = -T7 * dAlphaz_dVb - T10 - T0 * dVdsat_dVb;
            T10 = -4.0 * T4 * (T2 - 0.5 + 0.5 * T5) - o.cddb;
            tmp = -(T10 + T11 + T12);
            o.cbgb = -(o.cggb + o.cdgb + T12);
            o.cbdb = -(o.cgdb + o.cddb + T10);
            o.cbsb = -(o.cgsb + o.cdsb + tmp);
          } else {  // 50/50
            tmp = -CoxWL * T5 * dVdsat_dVb;
            o.cgsb = -(o.cggb + o.cgdb + tmp);
            T6 = B4_DIV(1.0, Vdsat);
            const double dAlphaz_dVg = T6 * (1.0 - Alphaz * dVdsat_dVg);
            const double dAlphaz_dVb = -T6 * (dVth_dVb + Alphaz * dVdsat_dVb);
            T7 = T1 + T3;
            qdrn = -T4 * T7;
            qbulk = -(qgate + qdrn + qdrn);
            T7 *= T9;
            T0 = T4 * (2.0 * T5 - 2.0);
            o.cdgb = (T0 * dVdsat_dVg - T7 * dAlphaz_dVg) * dVgs_eff_dVg;
            T12 = T0 * dVdsat_dVb - T7 * dAlphaz_dVb;
            o.cddb = T4 * (1.0 - 2.0 * T2 - T5);
            o.cdsb = -(o.cdgb + T12 + o.cddb);
            o.cbgb = -(o.cggb + 2.0 * o.cdgb);
            o.cbdb = -(o.cgdb + 2.0 * o.cddb);
            o.cbsb = -(o.cgsb + 2.0 * o.cdsb);
          }
        }
      }
    } else {
      // ---- capmod 1 / 2: single-piece charge model on a C-V specific Vgsteff (:2962-3462)
      double VbseffCV, dVbseffCV_dVb;
      if (Vbseff < 0.0) { VbseffCV = Vbseff; dVbseffCV_dVb = 1.0; }
      else { VbseffCV = phi - Phis; dVbseffCV_dVb = -dPhis_dVb; }
      const double Vgst = c.Vgst, n = c.n, dn_dVd = c.dn_dVd, dn_dVb = c.dn_dVb, dVth_dVd = c.dVth_dVd, dVth_dVb = c.dVth_dVb;
      double Vgsteff, dVgsteff_dVg, dVgsteff_dVd, dVgsteff_dVb;
      if ((int)M_(cvchargemod) == 0) {
        const double noff = n * S_(noff);
        const double dnoff_dVd = S_(noff) * dn_dVd, dnoff_dVb = S_(noff) * dn_dVb;
        T0 = Vtm * noff;
        const double voffcv = S_(voffcv);
        const double VgstNVt = B4_DIV((Vgst - voffcv), T0);
        if (VgstNVt > B4C_EXP_THRESHOLD) {
          Vgsteff = Vgst - voffcv;
          dVgsteff_dVg = dVgs_eff_dVg; dVgsteff_dVd = -dVth_dVd; dVgsteff_dVb = -dVth_dVb;
        } else if (VgstNVt < -B4C_EXP_THRESHOLD) {
          Vgsteff = T0 * log(1.0 + B4C_MIN_EXP);
          dVgsteff_dVg = 0.0;
          dVgsteff_dVd = B4_DIV(Vgsteff, noff);
          dVgsteff_dVb = dVgsteff_dVd * dnoff_dVb;
          dVgsteff_dVd *= dnoff_dVd;
        } else {
          const double ExpVgst = exp(VgstNVt);
          Vgsteff = T0 * log(1.0 + ExpVgst);
          dVgsteff_dVg = B4_DIV(ExpVgst, (1.0 + ExpVgst));
          dVgsteff_dVd = -dVgsteff_dVg * (dVth_dVd + B4_DIV((Vgst - voffcv), noff) * dnoff_dVd) + B4_DIV(Vgsteff, noff) * dnoff_dVd;
          dVgsteff_dVb = -dVgsteff_dVg * (dVth_dVb + B4_DIV((Vgst - voffcv), noff) * dnoff_dVb) + B4_DIV(Vgsteff, noff) * dnoff_dVb;
          dVgsteff_dVg *= dVgs_eff_dVg;
        }
      } else {
        const double mstarcv = S_(mstarcv);
        double dT10_dVg, dT10_dVd, dT10_dVb, dT9_dVg, dT9_dVd, dT9_dVb;
        T0 = n * Vtm;
        T1 = mstarcv * Vgst;
        T2 = B4_DIV(T1, T0);
        if (T2 > B4C_EXP_THRESHOLD) {
          T10 = T1;
          dT10_dVg = mstarcv * dVgs_eff_dVg; dT10_dVd = -dVth_dVd * mstarcv; dT10_dVb = -dVth_dVb * mstarcv;
        } else if (T2 < -B4C_EXP_THRESHOLD) {
          T10 = Vtm * log(1.0 + B4C_MIN_EXP);
          dT10_dVg = 0.0; dT10_dVd = T10 * dn_dVd; dT10_dVb = T10 * dn_dVb;
          T10 *= n;
        } else {
          const double ExpVgst = exp(T2);
          T3 = Vtm * log(1.0 + ExpVgst);
          T10 = n * T3;
          dT10_dVg = B4_DIV(mstarcv * ExpVgst, (1.0 + ExpVgst));
          dT10_dVb = T3 * dn_dVb - dT10_dVg * (dVth_dVb + B4_DIV(Vgst * dn_dVb, n));
          dT10_dVd = T3 * dn_dVd - dT10_dVg * (dVth_dVd + B4_DIV(Vgst * dn_dVd, n));
          dT10_dVg *= dVgs_eff_dVg;
        }
        T1 = S_(voffcbncv) - (1.0 - mstarcv) * Vgst;
        T2 = B4_DIV(T1, T0);
        if (T2 < -B4C_EXP_THRESHOLD) {
          T3 = B4_DIV(coxe * B4C_MIN_EXP, S_(cdep0));
          T9 = mstarcv + T3 * n;
          dT9_dVg = 0.0; dT9_dVd = dn_dVd * T3; dT9_dVb = dn_dVb * T3;
        } else if (T2 > B4C_EXP_THRESHOLD) {
          T3 = B4_DIV(coxe * B4C_MAX_EXP, S_(cdep0));
          T9 = mstarcv + T3 * n;
          dT9_dVg = 0.0; dT9_dVd = dn_dVd * T3; dT9_dVb = dn_dVb * T3;
        } else {
          const double ExpVgst = exp(T2);
          T3 = B4_DIV(coxe, S_(cdep0));
          T4 = T3 * ExpVgst;
          T5 = B4_DIV(T1 * T4, T0);
          T9 = mstarcv + n * T4;
          dT9_dVg = B4_DIV(T3 * (mstarcv - 1.0) * ExpVgst, Vtm);
          dT9_dVb = T4 * dn_dVb - dT9_dVg * dVth_dVb - T5 * dn_dVb;
          dT9_dVd = T4 * dn_dVd - dT9_dVg * dVth_dVd - T5 * dn_dVd;
          dT9_dVg *= dVgs_eff_dVg;
        }
        Vgsteff = B4_DIV(T10, T9);
        T11 = T9 * T9;
        dVgsteff_dVg = B4_DIV((T9 * dT10_dVg - T10 * dT9_dVg), T11);
        dVgsteff_dVd = B4_DIV((T9 * dT10_dVd - T10 * dT9_dVd), T11);
        dVgsteff_dVb = B4_DIV((T9 * dT10_dVb - T10 * dT9_dVb), T11);
      }

      // effective flat band (accumulation charge) — common to capmod 1 and 2
      const double vfbzb = I_(vfbzb);
      const double V3 = vfbzb - Vgs_eff + VbseffCV - B4C_DELTA_3;
      T0 = vfbzb <= 0.0 ? sqrt(V3 * V3 - 4.0 * B4C_DELTA_3 * vfbzb) : sqrt(V3 * V3 + 4.0 * B4C_DELTA_3 * vfbzb);
      T1 = 0.5 * (1.0 + B4_DIV(V3, T0));
      const double Vfbeff = vfbzb - 0.5 * (V3 + T0);
      const double dVfbeff_dVg = T1 * dVgs_eff_dVg;
      const double dVfbeff_dVb = -T1 * dVbseffCV_dVb;
      const double AbulkCV = Abulk0 * abulkCVfactor;
      const double dAbulkCV_dVb = abulkCVfactor * dAbulk0_dVb;
      double Qac0, dQac0_dVg, dQac0_dVb, Qsub0, dQsub0_dVg, dQsub0_dVd, dQsub0_dVb;
      double Cgg1, Cgd1, Cgb1, Cbg1, Cbd1, Cbb1, Csg, Csd, Csb, Cgg, Cgd, Cgb, Cbg, Cbd, Cbb;
      double VdseffCV, dVdseffCV_dVg, dVdseffCV_dVd, dVdseffCV_dVb;

      if (capmod == 1) {
        Qac0 = CoxWL * (Vfbeff - vfbzb);
        dQac0_dVg = CoxWL * dVfbeff_dVg;
        dQac0_dVb = CoxWL * dVfbeff_dVb;
        T0 = 0.5 * k1ox;
        T3 = Vgs_eff - Vfbeff - VbseffCV - Vgsteff;
        if (k1ox == 0.0) { T1 = 0.0; T2 = 0.0; }
        else if (T3 < 0.0) { T1 = T0 + B4_DIV(T3, k1ox); T2 = CoxWL; }
        else { T1 = sqrt(T0 * T0 + T3); T2 = B4_DIV(CoxWL * T0, T1); }
        Qsub0 = CoxWL * k1ox * (T1 - T0);
        dQsub0_dVg = T2 * (dVgs_eff_dVg - dVfbeff_dVg - dVgsteff_dVg);
        dQsub0_dVd = -T2 * dVgsteff_dVd;
        dQsub0_dVb = -T2 * (dVfbeff_dVb + dVbseffCV_dVb + dVgsteff_dVb);

        const double VdsatCV = B4_DIV(Vgsteff, AbulkCV);
        T0 = VdsatCV - Vds - B4C_DELTA_4;
        dT0_dVg = B4_DIV(1.0, AbulkCV);
        dT0_dVb = B4_DIV(-VdsatCV * dAbulkCV_dVb, AbulkCV);
        T1 = sqrt(T0 * T0 + 4.0 * B4C_DELTA_4 * VdsatCV);
        dT1_dVg = B4_DIV((T0 + B4C_DELTA_4 + B4C_DELTA_4), T1);
        dT1_dVd = B4_DIV(-T0, T1);
        dT1_dVb = dT1_dVg * dT0_dVb;
        dT1_dVg *= dT0_dVg;
        if (T0 >= 0.0) {
          VdseffCV = VdsatCV - 0.5 * (T0 + T1);
          dVdseffCV_dVg = 0.5 * (dT0_dVg - dT1_dVg);
          dVdseffCV_dVd = 0.5 * (1.0 - dT1_dVd);
          dVdseffCV_dVb = 0.5 * (dT0_dVb - dT1_dVb);
        } else {
          T3 = B4_DIV((B4C_DELTA_4 + B4C_DELTA_4), (T1 - T0));
          T4 = 1.0 - T3;
          T5 = B4_DIV(VdsatCV * T3, (T1 - T0));
          VdseffCV = VdsatCV * T4;
          dVdseffCV_dVg = dT0_dVg * T4 + T5 * (dT1_dVg - dT0_dVg);
          dVdseffCV_dVd = T5 * (dT1_dVd + 1.0);
          dVdseffCV_dVb = dT0_dVb * (T4 - T5) + T5 * dT1_dVb;
        }
        if (Vds == 0.0) { VdseffCV = 0.0; dVdseffCV_dVg = 0.0; dVdseffCV_dVb = 0.0; }

        T0 = AbulkCV * VdseffCV;
        T1 = 12.0 * (Vgsteff - 0.5 * T0 + 1.0e-20);
        T2 = B4_DIV(VdseffCV, T1);
        T3 = T0 * T2;
        T4 = (1.0 - 12.0 * T2 * T2 * AbulkCV);
        T5 = (B4_DIV(6.0 * T0 * (4.0 * Vgsteff - T0), (T1 * T1)) - 0.5);
        T6 = 12.0 * T2 * T2 * Vgsteff;
        qgate = CoxWL * (Vgsteff - 0.5 * VdseffCV + T3);
        Cgg1 = CoxWL * (T4 + T5 * dVdseffCV_dVg);
        Cgd1 = CoxWL * T5 * dVdseffCV_dVd + Cgg1 * dVgsteff_dVd;
        Cgb1 = CoxWL * (T5 * dVdseffCV_dVb + T6 * dAbulkCV_dVb) + Cgg1 * dVgsteff_dVb;
        Cgg1 *= dVgsteff_dVg;
        T7 = 1.0 - AbulkCV;
        qbulk = CoxWL * T7 * (0.5 * VdseffCV - T3);
        T4 = -T7 * (T4 - 1.0);
        T5 = -T7 * T5;
        T6 = -(T7 * T6 + (0.5 * VdseffCV - T3));
        Cbg1 = CoxWL * (T4 + T5 * dVdseffCV_dVg);
        Cbd1 = CoxWL * T5 * dVdseffCV_dVd + Cbg1 * dVgsteff_dVd;
        Cbb1 = CoxWL * (T5 * dVdseffCV_dVb + T6 * dAbulkCV_dVb) + Cbg1 * dVgsteff_dVb;
        Cbg1 *= dVgsteff_dVg;
        if (xpart > 0.5) {
          T1 = T1 + T1;
          qsrc = -CoxWL * (0.5 * Vgsteff + 0.25 * T0 - B4_DIV(T0 * T0, T1));
          T7 = B4_DIV((4.0 * Vgsteff - T0), (T1 * T1));
          T4 = -(0.5 + B4_DIV(24.0 * T0 * T0, (T1 * T1)));
          T5 = -(0.25 * AbulkCV - 12.0 * AbulkCV * T0 * T7);
          T6 = -(0.25 * VdseffCV - 12.0 * T0 * VdseffCV * T7);
          Csg = CoxWL * (T4 + T5 * dVdseffCV_dVg);
          Csd = CoxWL * T5 * dVdseffCV_dVd + Csg * dVgsteff_dVd;
          Csb = CoxWL * (T5 * dVdseffCV_dVb + T6 * dAbulkCV_dVb) + Csg * dVgsteff_dVb;
          Csg *= dVgsteff_dVg;
        } else if (xpart < 0.5) {
          T1 = B4_DIV(T1, 12.0);
          T2 = B4_DIV(0.5 * CoxWL, (T1 * T1));
          T3 = Vgsteff * (B4_DIV(2.0 * T0 * T0, 3.0) + Vgsteff * (Vgsteff - B4_DIV(4.0 * T0, 3.0))) - B4_DIV(2.0 * T0 * T0 * T0, 15.0);
          qsrc = -T2 * T3;
          T7 = B4_DIV(4.0, 3.0) * Vgsteff * (Vgsteff - T0) + 0.4 * T0 * T0;
          T4 = B4_DIV(-2.0 * qsrc, T1) - T2 * (Vgsteff * (3.0 * Vgsteff - B4_DIV(8.0 * T0, 3.0)) + B4_DIV(2.0 * T0 * T0, 3.0));
          T5 = (B4_DIV(qsrc, T1) + T2 * T7) * AbulkCV;
          T6 = (B4_DIV(qsrc, T1) * VdseffCV + T2 * T7 * VdseffCV);
          Csg = (T4 + T5 * dVdseffCV_dVg);
          Csd = T5 * dVdseffCV_dVd + Csg * dVgsteff_dVd;
          Csb = (T5 * dVdseffCV_dVb + T6 * dAbulkCV_dVb) + Csg * dVgsteff_dVb;
          Csg *= dVgsteff_dVg;
        } else {
          qsrc = -0.5 * (qgate + qbulk);
          Csg = -0.5 * (Cgg1 + Cbg1);
          Csb = -0.5 * (Cgb1 + Cbb1);
          Csd = -0.5 * (Cgd1 + Cbd1);
        }
        qgate += Qac0 + Qsub0;
        qbulk -= (Qac0 + Qsub0);
        qdrn = -(qgate + qbulk + qsrc);
        Cgg = dQac0_dVg + dQsub0_dVg + Cgg1;
        Cgd = dQsub0_dVd + Cgd1;
        Cgb = dQac0_dVb + dQsub0_dVb + Cgb1;
        Cbg = Cbg1 - dQac0_dVg - dQsub0_dVg;
        Cbd = Cbd1 - dQsub0_dVd;
        Cbb = Cbb1 - dQac0_dVb - dQsub0_dVb;
      } else {
        // charge-thickness model: finite inversion/accumulation layer thickness lowers the effective oxide capacitance
        const double Cox = I_(coxp);
        double Tox = 1.0e8 * I_(toxp);
        T0 = B4_DIV((Vgs_eff - VbseffCV - vfbzb), Tox);
        dT0_dVg = B4_DIV(dVgs_eff_dVg, Tox);
        dT0_dVb = B4_DIV(-dVbseffCV_dVb, Tox);
        const double ldeb = S_(ldeb), acde = S_(acde);
        tmp = T0 * acde;
        double Tcen, dTcen_dVg, dTcen_dVb;
        if (-B4C_EXP_THRESHOLD < tmp && tmp < B4C_EXP_THRESHOLD) {
          Tcen = ldeb * exp(tmp);
          dTcen_dVg = acde * Tcen;
          dTcen_dVb = dTcen_dVg * dT0_dVb;
          dTcen_dVg *= dT0_dVg;
        } else if (tmp <= -B4C_EXP_THRESHOLD) {
          Tcen = ldeb * B4C_MIN_EXP; dTcen_dVg = 0.0; dTcen_dVb = 0.0;
        } else {
          Tcen = ldeb * B4C_MAX_EXP; dTcen_dVg = 0.0; dTcen_dVb = 0.0;
        }
        const double LINK = 1.0e-3 * I_(toxp);
        const double V3c = ldeb - Tcen - LINK;
        const double V4c = sqrt(V3c * V3c + 4.0 * LINK * ldeb);
        Tcen = ldeb - 0.5 * (V3c + V4c);
        T1 = 0.5 * (1.0 + B4_DIV(V3c, V4c));
        dTcen_dVg *= T1;
        dTcen_dVb *= T1;
        double Ccen = B4_DIV(epssub, Tcen);
        T2 = B4_DIV(Cox, (Cox + Ccen));
        double Coxeff = T2 * Ccen;
        T3 = B4_DIV(-Ccen, Tcen);
        double dCoxeff_base = T2 * T2 * T3;
        double dCoxeff_dVb = dCoxeff_base * dTcen_dVb;
        double dCoxeff_dVg = dCoxeff_base * dTcen_dVg;
        double CoxWLcen = B4_DIV(CoxWL * Coxeff, coxe);
        Qac0 = CoxWLcen * (Vfbeff - vfbzb);
        double QovCox = B4_DIV(Qac0, Coxeff);
        dQac0_dVg = CoxWLcen * dVfbeff_dVg + QovCox * dCoxeff_dVg;
        dQac0_dVb = CoxWLcen * dVfbeff_dVb + QovCox * dCoxeff_dVb;
        T0 = 0.5 * k1ox;
        T3 = Vgs_eff - Vfbeff - VbseffCV - Vgsteff;
        if (k1ox == 0.0) { T1 = 0.0; T2 = 0.0; }
        else if (T3 < 0.0) { T1 = T0 + B4_DIV(T3, k1ox); T2 = CoxWLcen; }
        else { T1 = sqrt(T0 * T0 + T3); T2 = B4_DIV(CoxWLcen * T0, T1); }
        Qsub0 = CoxWLcen * k1ox * (T1 - T0);
        QovCox = B4_DIV(Qsub0, Coxeff);
        dQsub0_dVg = T2 * (dVgs_eff_dVg - dVfbeff_dVg - dVgsteff_dVg) + QovCox * dCoxeff_dVg;
        dQsub0_dVd = -T2 * dVgsteff_dVd;
        dQsub0_dVb = -T2 * (dVfbeff_dVb + dVbseffCV_dVb + dVgsteff_dVb) + QovCox * dCoxeff_dVb;

        // gate-bias dependent surface potential increment
        double Denomi;
        if (k1ox <= 0.0) { Denomi = 0.25 * S_(moin) * Vtm; T0 = 0.5 * S_(sqrtPhi); }
        else { Denomi = S_(moin) * Vtm * k1ox * k1ox; T0 = k1ox * S_(sqrtPhi); }
        T1 = 2.0 * T0 + Vgsteff;
        const double DeltaPhi = Vtm * log(1.0 + B4_DIV(T1 * Vgsteff, Denomi));
        const double dDeltaPhi_dVg = B4_DIV(2.0 * Vtm * (T1 - T0), (Denomi + T1 * Vgsteff));
        T0 = Vgsteff - DeltaPhi - 0.001;
        dT0_dVg = 1.0 - dDeltaPhi_dVg;
        T1 = sqrt(T0 * T0 + Vgsteff * 0.004);
        const double VgDP = 0.5 * (T0 + T1);
        const double dVgDP_dVg = 0.5 * (dT0_dVg + B4_DIV((T0 * dT0_dVg + 0.002), T1));

        // inversion-layer centroid
        Tox += Tox;
        T0 = B4_DIV((Vgsteff + I_(vtfbphi2)), Tox);
        tmp = exp(M_(bdos) * 0.7 * log(T0));
        T1 = 1.0 + tmp;
        T2 = B4_DIV(M_(bdos) * 0.7 * tmp, (T0 * Tox));
        Tcen = B4_DIV(M_(ados) * 1.9e-9, T1);
        dTcen_dVg = B4_DIV(-Tcen * T2, T1);
        const double dTcen_dVd = dTcen_dVg * dVgsteff_dVd;
        dTcen_dVb = dTcen_dVg * dVgsteff_dVb;
        dTcen_dVg *= dVgsteff_dVg;
        Ccen = B4_DIV(epssub, Tcen);
        T0 = B4_DIV(Cox, (Cox + Ccen));
        Coxeff = T0 * Ccen;
        T1 = B4_DIV(-Ccen, Tcen);
        dCoxeff_base = T0 * T0 * T1;
        const double dCoxeff_dVd = dCoxeff_base * dTcen_dVd;
        dCoxeff_dVb = dCoxeff_base * dTcen_dVb;
        dCoxeff_dVg = dCoxeff_base * dTcen_dVg;
        CoxWLcen = B4_DIV(CoxWL * Coxeff, coxe);

        const double VdsatCV = B4_DIV(VgDP, AbulkCV);
        T0 = VdsatCV - Vds - B4C_DELTA_4;
        dT0_dVg = B4_DIV(dVgDP_dVg, AbulkCV);
        dT0_dVb = B4_DIV(-VdsatCV * dAbulkCV_dVb, AbulkCV);
        T1 = sqrt(T0 * T0 + 4.0 * B4C_DELTA_4 * VdsatCV);
        dT1_dVg = B4_DIV((T0 + B4C_DELTA_4 + B4C_DELTA_4), T1);
        dT1_dVd = B4_DIV(-T0, T1);
        dT1_dVb = dT1_dVg * dT0_dVb;
        dT1_dVg *= dT0_dVg;
        if (T0 >= 0.0) {
          VdseffCV = VdsatCV - 0.5 * (T0 + T1);
          dVdseffCV_dVg = 0.5 * (dT0_dVg - dT1_dVg);
          dVdseffCV_dVd = 0.5 * (1.0 - dT1_dVd);
          dVdseffCV_dVb = 0.5 * (dT0_dVb - dT1_dVb);
        } else {
          T3 = B4_DIV((B4C_DELTA_4 + B4C_DELTA_4), (T1 - T0));
          T4 = 1.0 - T3;
          T5 = B4_DIV(VdsatCV * T3, (T1 - T0));
          VdseffCV = VdsatCV * T4;
          dVdseffCV_dVg = dT0_dVg * T4 + T5 * (dT1_dVg - dT0_dVg);
          dVdseffCV_dVd = T5 * (dT1_dVd + 1.0);
          dVdseffCV_dVb = dT0_dVb * (T4 - T5) + T5 * dT1_dVb;
        }
        if (Vds == 0.0) { VdseffCV = 0.0; dVdseffCV_dVg = 0.0; dVdseffCV_dVb = 0.0; }

        T0 = AbulkCV * VdseffCV;
        T1 = VgDP;
        T2 = 12.0 * (T1 - 0.5 * T0 + 1.0e-20);
        T3 = B4_DIV(T0, T2);
        T4 = 1.0 - 12.0 * T3 * T3;
        T5 = AbulkCV * (B4_DIV(6.0 * T0 * (4.0 * T1 - T0), (T2 * T2)) - 0.5);
        T6 = B4_DIV(T5 * VdseffCV, AbulkCV);
        qgate = CoxWLcen * (T1 - T0 * (0.5 - T3));
        QovCox = B4_DIV(qgate, Coxeff);
        Cgg1 = CoxWLcen * (T4 * dVgDP_dVg + T5 * dVdseffCV_dVg);
        Cgd1 = CoxWLcen * T5 * dVdseffCV_dVd + Cgg1 * dVgsteff_dVd + QovCox * dCoxeff_dVd;
        Cgb1 = CoxWLcen * (T5 * dVdseffCV_dVb + T6 * dAbulkCV_dVb) + Cgg1 * dVgsteff_dVb + QovCox * dCoxeff_dVb;
        Cgg1 = Cgg1 * dVgsteff_dVg + QovCox * dCoxeff_dVg;
        T7 = 1.0 - AbulkCV;
        T8 = T2 * T2;
        T9 = B4_DIV(12.0 * T7 * T0 * T0, (T8 * AbulkCV));
        T10 = T9 * dVgDP_dVg;
        T11 = B4_DIV(-T7 * T5, AbulkCV);
        T12 = -(B4_DIV(T9 * T1, AbulkCV) + VdseffCV * (0.5 - B4_DIV(T0, T2)));
        qbulk = CoxWLcen * T7 * (0.5 * VdseffCV - B4_DIV(T0 * VdseffCV, T2));
        QovCox = B4_DIV(qbulk, Coxeff);
        Cbg1 = CoxWLcen * (T10 + T11 * dVdseffCV_dVg);
        Cbd1 = CoxWLcen * T11 * dVdseffCV_dVd + Cbg1 * dVgsteff_dVd + QovCox * dCoxeff_dVd;
        Cbb1 = CoxWLcen * (T11 * dVdseffCV_dVb + T12 * dAbulkCV_dVb) + Cbg1 * dVgsteff_dVb + QovCox * dCoxeff_dVb;
        Cbg1 = Cbg1 * dVgsteff_dVg + QovCox * dCoxeff_dVg;
        if (xpart > 0.5) {
          qsrc = -CoxWLcen * (B4_DIV(T1, 2.0) + B4_DIV(T0, 4.0) - B4_DIV(0.5 * T0 * T0, T2));
          QovCox = B4_DIV(qsrc, Coxeff);
          T2 += T2;
          T3 = T2 * T2;
          T7 = -(0.25 - B4_DIV(12.0 * T0 * (4.0 * T1 - T0), T3));
          T4 = -(0.5 + B4_DIV(24.0 * T0 * T0, T3)) * dVgDP_dVg;
          T5 = T7 * AbulkCV;
          T6 = T7 * VdseffCV;
          Csg = CoxWLcen * (T4 + T5 * dVdseffCV_dVg);
          Csd = CoxWLcen * T5 * dVdseffCV_dVd + Csg * dVgsteff_dVd + QovCox * dCoxeff_dVd;
          Csb = CoxWLcen * (T5 * dVdseffCV_dVb + T6 * dAbulkCV_dVb) + Csg * dVgsteff_dVb + QovCox * dCoxeff_dVb;
          Csg = Csg * dVgsteff_dVg + QovCox * dCoxeff_dVg;
        } else if (xpart < 0.5) {
          T2 = B4_DIV(T2, 12.0);
          T3 = B4_DIV(0.5 * CoxWLcen, (T2 * T2));
          T4 = T1 * (B4_DIV(2.0 * T0 * T0, 3.0) + T1 * (T1 - B4_DIV(4.0 * T0, 3.0))) - B4_DIV(2.0 * T0 * T0 * T0, 15.0);
          qsrc = -T3 * T4;
          QovCox = B4_DIV(qsrc, Coxeff);
          T8 = B4_DIV(4.0, 3.0) * T1 * (T1 - T0) + 0.4 * T0 * T0;
          T5 = B4_DIV(-2.0 * qsrc, T2) - T3 * (T1 * (3.0 * T1 - B4_DIV(8.0 * T0, 3.0)) + B4_DIV(2.0 * T0 * T0, 3.0));
          T6 = AbulkCV * (B4_DIV(qsrc, T2) + T3 * T8);
          T7 = B4_DIV(T6 * VdseffCV, AbulkCV);
          Csg = T5 * dVgDP_dVg + T6 * dVdseffCV_dVg;
          Csd = Csg * dVgsteff_dVd + T6 * dVdseffCV_dVd + QovCox * dCoxeff_dVd;
          Csb = Csg * dVgsteff_dVb + T6 * dVdseffCV_dVb + T7 * dAbulkCV_dVb + QovCox * dCoxeff_dVb;
          Csg = Csg * dVgsteff_dVg + QovCox * dCoxeff_dVg;
        } else {
          qsrc = -0.5 * qgate;
          Csg = -0.5 * Cgg1;
          Csd = -0.5 * Cgd1;
          Csb = -0.5 * Cgb1;
        }
        qgate += Qac0 + Qsub0 - qbulk;
        qbulk -= (Qac0 + Qsub0);
        qdrn = -(qgate + qbulk + qsrc);
        Cbg = Cbg1 - dQac0_dVg - dQsub0_dVg;
        Cbd = Cbd1 - dQsub0_dVd;
        Cbb = Cbb1 - dQac0_dVb - dQsub0_dVb;
        Cgg = Cgg1 - Cbg;
        Cgd = Cgd1 - Cbd;
        Cgb = Cgb1 - Cbb;
      }
      Cgb *= dVbseff_dVb;
      Cbb *= dVbseff_dVb;
      Csb *= dVbseff_dVb;
      o.cggb = Cgg;
      o.cgsb = -(Cgg + Cgd + Cgb);
      o.cgdb = Cgd;
      o.cdgb = -(Cgg + Cbg + Csg);
      o.cdsb = (Cgg + Cgd + Cgb + Cbg + Cbd + Cbb + Csg + Csd + Csb);
      o.cddb = -(Cgd + Cbd + Csd);
      o.cbgb = Cbg;
      o.cbsb = -(Cbg + Cbd + Cbb);
      o.cbdb = Cbd;
    }
    o.qgate = qgate; o.qbulk = qbulk; o.qdrn = qdrn;
    // channel charge seen by the NQS relaxation network (:3477-3495)
    if (trnqsmod != 0 || I_(acnqsmod) != 0.0) {
      const double qcheq = -(qbulk + qgate);
      o.qchqs = qcheq;
      o.cqgb = -(o.cggb + o.cbgb);
      o.cqdb = -(o.cgdb + o.cbdb);
      o.cqsb = -(o.cgsb + o.cbsb);
      o.cqbb = -(o.cqgb + o.cqdb + o.cqsb);
      o.qcheq = qcheq;
    }
  }

  // ---- S/D junction depletion charge (:3498-3614)
  {
    const double weffCJ_nf = S_(weffCJ);
    const double czbd = D_(DunitAreaTempJctCap) * I_(Adeff), czbs = D_(SunitAreaTempJctCap) * I_(Aseff);
    const double czbdsw = D_(DunitLengthSidewallTempJctCap) * I_(Pdeff);
    const double czbdswg = D_(DunitLengthGateSidewallTempJctCap) * weffCJ_nf * nf;
    const double czbssw = D_(SunitLengthSidewallTempJctCap) * I_(Pseff);
    const double czbsswg = D_(SunitLengthGateSidewallTempJctCap) * weffCJ_nf * nf;
    b4_junction_charge(v.vbs_jct, czbs, czbssw, czbsswg, D_(PhiBS), D_(PhiBSWS), D_(PhiBSWGS), M_(mjs), M_(mjsws), M_(mjswgs), &o.qbs, &o.capbs);
    b4_junction_charge(v.vbd_jct, czbd, czbdsw, czbdswg, D_(PhiBD), D_(PhiBSWD), D_(PhiBSWGD), M_(mjd), M_(mjswd), M_(mjswgd), &o.qbd, &o.capbd);
  }

  // ---- NQS time constant (:3629-3635)
  if (trnqsmod != 0) {
    const double CoxWL = coxe * S_(weffCV) * nf * S_(leffCV);
    T1 = B4_DIV(o.gcrg, CoxWL);
    o.gtau = T1 * 1.0e-9;
  } else {
    o.gtau = 0.0;
  }

  // ---- gate overlap capacitance/charge (:3637-3672)
  const double vgdx = rgatemod == 3 ? v.vgmd : v.vgd, vgsx = rgatemod == 3 ? v.vgms : v.vgs;
  double cgdo, qgdo, cgso, qgso;
  if (capmod == 0) {
    cgdo = S_(cgdo); qgdo = S_(cgdo) * vgdx;
    cgso = S_(cgso); qgso = S_(cgso) * vgsx;
  } else {
    b4_overlap(vgdx, S_(cgdo), S_(weffCV), S_(cgdl), S_(ckappad), &cgdo, &qgdo);
    b4_overlap(vgsx, S_(cgso), S_(weffCV), S_(cgsl), S_(ckappas), &cgso, &qgso);
  }
  if (nf != 1.0) { cgdo *= nf; cgso *= nf; qgdo *= nf; qgso *= nf; }
  o.cgdo = cgdo; o.qgdo = qgdo; o.cgso = cgso; o.qgso = qgso;

  // ---- distribute the overlap charge onto the terminals (:3679-3746)
  const double cgbo = S_(cgbo);
  if (o.mode > 0) {
    if (trnqsmod == 0) {
      qdrn -= qgdo;
      if (rgatemod == 3) {
        const double qgmb = cgbo * v.vgmb;
        qgmid = qgdo + qgso + qgmb;
        qbulk -= qgmb;
        qsrc = -(qgate + qgmid + qbulk + qdrn);
      }
    } else if (rgatemod == 3) {
      const double qgmb = cgbo * v.vgmb;
      qgmid = qgdo + qgso + qgmb;
      qgate = 0.0; qbulk = -qgmb; qdrn = -qgdo;
      qsrc = -(qgmid + qbulk + qdrn);
    } else {
      const double qgb = cgbo * v.vgb;
      qgate = qgdo + qgso + qgb;
      qbulk = -qgb; qdrn = -qgdo;
      qsrc = -(qgate + qbulk + qdrn);
    }
  } else {
    if (trnqsmod == 0) {
      qsrc = qdrn - qgso;
      if (rgatemod == 3) {
        const double qgmb = cgbo * v.vgmb;
        qgmid = qgdo + qgso + qgmb;
        qbulk -= qgmb;
        qdrn = -(qgate + qgmid + qbulk + qsrc);
      } else {
        const double qgb = cgbo * v.vgb;
        qgate += qgdo + qgso + qgb;
        qbulk -= qgb;
        qdrn = -(qgate + qbulk + qsrc);
      }
    } else if (rgatemod == 3) {
      const double qgmb = cgbo * v.vgmb;
      qgmid = qgdo + qgso + qgmb;
      qgate = 0.0; qbulk = -qgmb; qdrn = -qgdo; qsrc = -qgso;
    } else {
      const double qgb = cgbo * v.vgb;
      qgate = qgdo + qgso + qgb;
      qbulk = -qgb; qdrn = -qgdo; qsrc = -qgso;
    }
  }
  (void)qsrc;
  o.qg = qgate;
  o.qd = qdrn - o.qbd;
  o.qgmid = rgatemod == 3 ? qgmid : 0.0;
  o.qb = rbodymod != 0 ? qbulk + o.qbd + o.qbs : qbulk;
}

}  // namespace b4e
}  // namespace s21

#include "bsim4_eval_stamp.hpp"
