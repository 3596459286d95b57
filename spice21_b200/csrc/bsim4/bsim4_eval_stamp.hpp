// BSIM4 evaluation, phases 6 and 7: the transient companion model (capacitances -> conductances and equivalent
// currents, Backward Euler baked in as in the reference) and the MNA stamp. Ends with `load_bsim4`, the device's
// Component::load. Follows tran.rs:11-549 and stamp.rs:8-572 (see bsim4_eval.hpp header).
#pragma once

namespace s21 {
namespace b4e {

// Dynamic (dq/dt) part of the stamp. All zero in an operating-point solve.
struct B4Dyn {
  double gcdgb, gcddb, gcdsb, gcdbb, gcsgb, gcsdb, gcssb, gcsbb, gcggb, gcgdb, gcgsb, gcgbb, gcbdb, gcbgb, gcbsb, gcbbb;
  double gcgmgmb, gcgmdb, gcgmsb, gcgmbb, gcdgmb, gcsgmb, gcbgmb, gcsbsb;
  double gqdef, gcqgb, gcqdb, gcqsb, gcqbb, ggtg, ggtd, ggtb, ggts;
  double dxpart, sxpart, ddxpart_dVd, ddxpart_dVg, ddxpart_dVb, ddxpart_dVs, dsxpart_dVd, dsxpart_dVg, dsxpart_dVb, dsxpart_dVs;
  double ceqqg, ceqqd, ceqqb, ceqqjs, ceqqjd, ceqqgmid, cqdef, cqcheq;
};

template <class E> B4_HD void b4_tran_caps(E& e, const B4Bias& v, const B4Op& o, B4Dyn& y) {
  const int trnqsmod = (int)M_(trnqsmod), rgatemod = (int)M_(rgatemod), rbodymod = (int)M_(rbodymod);
  const double nqs_scaling_factor = 1.0e-9;
  const double ag0 = B4_DIV(1.0, e.dt);
  const double cgdo = o.cgdo, cgso = o.cgso, cgbo = S_(cgbo);
  const bool fwd = o.mode > 0;
  double gcdbdb = 0.0;  // used for the equivalent currents below, but never carried into the stamp (tran.rs:497-549)
  y.gqdef = 0.0; y.gcqgb = 0.0; y.gcqdb = 0.0; y.gcqsb = 0.0; y.gcqbb = 0.0;
  y.ggtg = 0.0; y.ggtd = 0.0; y.ggtb = 0.0; y.ggts = 0.0;
  y.ddxpart_dVd = 0.0; y.ddxpart_dVg = 0.0; y.ddxpart_dVb = 0.0; y.ddxpart_dVs = 0.0;
  y.dsxpart_dVd = 0.0; y.dsxpart_dVg = 0.0; y.dsxpart_dVb = 0.0; y.dsxpart_dVs = 0.0;
  y.gcgmgmb = 0.0; y.gcgmdb = 0.0; y.gcgmsb = 0.0; y.gcgmbb = 0.0; y.gcdgmb = 0.0; y.gcsgmb = 0.0; y.gcbgmb = 0.0;
  y.ceqqjs = 0.0; y.ceqqjd = 0.0; y.ceqqgmid = 0.0; y.cqdef = 0.0; y.cqcheq = 0.0;

  if (rgatemod == 3) {  // the mid-gate node carries the overlap capacitances in every variant
    y.gcgmgmb = (cgdo + cgso + cgbo) * ag0;
    y.gcgmdb = -cgdo * ag0;
    y.gcgmsb = -cgso * ag0;
    y.gcgmbb = -cgbo * ag0;
    y.gcdgmb = y.gcgmdb;
    y.gcsgmb = y.gcgmsb;
    y.gcbgmb = y.gcgmbb;
  }
  if (trnqsmod == 0) {
    // quasi-static: intrinsic capacitances, with drain/source columns exchanged in reverse mode
    const double cgd_ = fwd ? o.cgdb : o.cgsb, cgs_ = fwd ? o.cgsb : o.cgdb;
    if (rgatemod == 3) {
      y.gcggb = o.cggb * ag0;
      y.gcgdb = cgd_ * ag0;
      y.gcgsb = cgs_ * ag0;
      y.gcgbb = -(y.gcggb + y.gcgdb + y.gcgsb);
      if (fwd) { y.gcdgb = o.cdgb * ag0; y.gcsgb = -(o.cggb + o.cbgb + o.cdgb) * ag0; }
      else { y.gcdgb = -(o.cggb + o.cbgb + o.cdgb) * ag0; y.gcsgb = o.cdgb * ag0; }
      y.gcbgb = o.cbgb * ag0;
    } else {
      y.gcggb = (o.cggb + cgdo + cgso + cgbo) * ag0;
      y.gcgdb = (cgd_ - cgdo) * ag0;
      y.gcgsb = (cgs_ - cgso) * ag0;
      y.gcgbb = -(y.gcggb + y.gcgdb + y.gcgsb);
      if (fwd) { y.gcdgb = (o.cdgb - cgdo) * ag0; y.gcsgb = -(o.cggb + o.cbgb + o.cdgb + cgso) * ag0; }
      else { y.gcdgb = -(o.cggb + o.cbgb + o.cdgb + cgdo) * ag0; y.gcsgb = (o.cdgb - cgso) * ag0; }
      y.gcbgb = (o.cbgb - cgbo) * ag0;
    }
    if (fwd) {
      y.gcddb = (o.cddb + o.capbd + cgdo) * ag0;
      y.gcdsb = o.cdsb * ag0;
      y.gcsdb = -(o.cgdb + o.cbdb + o.cddb) * ag0;
      y.gcssb = (o.capbs + cgso - (o.cgsb + o.cbsb + o.cdsb)) * ag0;
      if (rbodymod == 0) {
        y.gcdbb = -(y.gcdgb + y.gcddb + y.gcdsb + y.gcdgmb);
        y.gcsbb = -(y.gcsgb + y.gcsdb + y.gcssb + y.gcsgmb);
        y.gcbdb = (o.cbdb - o.capbd) * ag0;
        y.gcbsb = (o.cbsb - o.capbs) * ag0;
        gcdbdb = 0.0; y.gcsbsb = 0.0;
      } else {
        y.gcdbb = -(o.cddb + o.cdgb + o.cdsb) * ag0;
        y.gcsbb = -(y.gcsgb + y.gcsdb + y.gcssb + y.gcsgmb) + o.capbs * ag0;
        y.gcbdb = o.cbdb * ag0;
        y.gcbsb = o.cbsb * ag0;
        gcdbdb = -o.capbd * ag0; y.gcsbsb = -o.capbs * ag0;
      }
      y.gcbbb = -(y.gcbdb + y.gcbgb + y.gcbsb + y.gcbgmb);
      y.sxpart = 0.6; y.dxpart = 0.4;
    } else {
      y.gcddb = (o.capbd + cgdo - (o.cgsb + o.cbsb + o.cdsb)) * ag0;
      y.gcdsb = -(o.cgdb + o.cbdb + o.cddb) * ag0;
      y.gcsdb = o.cdsb * ag0;
      y.gcssb = (o.cddb + o.capbs + cgso) * ag0;
      if (rbodymod == 0) {
        y.gcdbb = -(y.gcdgb + y.gcddb + y.gcdsb + y.gcdgmb);
        y.gcsbb = -(y.gcsgb + y.gcsdb + y.gcssb + y.gcsgmb);
        y.gcbdb = (o.cbsb - o.capbd) * ag0;
        y.gcbsb = (o.cbdb - o.capbs) * ag0;
        gcdbdb = 0.0; y.gcsbsb = 0.0;
      } else {
        y.gcdbb = -(y.gcdgb + y.gcddb + y.gcdsb + y.gcdgmb) + o.capbd * ag0;
        y.gcsbb = -(o.cddb + o.cdgb + o.cdsb) * ag0;
        y.gcbdb = o.cbsb * ag0;
        y.gcbsb = o.cbdb * ag0;
        gcdbdb = -o.capbd * ag0; y.gcsbsb = -o.capbs * ag0;
      }
      y.gcbbb = -(y.gcbgb + y.gcbdb + y.gcbsb + y.gcbgmb);
      y.sxpart = 0.4; y.dxpart = 0.6;
    }
  } else {
    // non-quasi-static: the channel charge lives on the internal q node; terminals only see overlap + junction caps
    const double qcheq = o.qchqs;
    const double CoxWL = D_(coxe) * S_(weffCV) * I_(nf) * S_(leffCV);
    const double T0 = B4_DIV(v.qdef * nqs_scaling_factor, CoxWL);
    y.ggtg = T0 * o.gcrgg;
    y.ggtb = T0 * o.gcrgb;
    if (fwd) { y.ggtd = T0 * o.gcrgd; y.ggts = T0 * o.gcrgs; }
    else { y.ggts = T0 * o.gcrgd; y.ggtd = T0 * o.gcrgs; }
    y.gqdef = nqs_scaling_factor * ag0;
    y.gcqgb = o.cqgb * ag0;
    y.gcqdb = (fwd ? o.cqdb : o.cqsb) * ag0;
    y.gcqsb = (fwd ? o.cqsb : o.cqdb) * ag0;
    y.gcqbb = o.cqbb * ag0;
    const double xpart = M_(xpart);
    // partition of the channel charge between the terminal that plays "drain" (p_) and the other one
    double p_, dp_dVd, dp_dVg, dp_dVs, dp_dVb;
    if (fabs(qcheq) <= 1.0e-5 * CoxWL) {
      p_ = xpart < 0.5 ? 0.4 : (xpart > 0.5 ? 0.0 : 0.5);
      dp_dVd = 0.0; dp_dVg = 0.0; dp_dVb = 0.0; dp_dVs = 0.0;
    } else {
      p_ = B4_DIV(o.qdrn, qcheq);
      const double Cdd = o.cddb;
      const double Csd = -(o.cgdb + o.cddb + o.cbdb);
      dp_dVd = B4_DIV((Cdd - p_ * (Cdd + Csd)), qcheq);
      const double Cdg = o.cdgb;
      const double Csg = -(o.cggb + o.cdgb + o.cbgb);
      dp_dVg = B4_DIV((Cdg - p_ * (Cdg + Csg)), qcheq);
      const double Cds = o.cdsb;
      const double Css = fwd ? -(o.cgsb + o.cdsb + o.cbsb) : -(o.cgsb + o.cdsb + o.cbs);  // reverse mode reads the junction current (tran.rs:358)
      dp_dVs = B4_DIV((Cds - p_ * (Cds + Css)), qcheq);
      dp_dVb = -(dp_dVd + dp_dVg + dp_dVs);
    }
    if (fwd) {
      y.dxpart = p_; y.ddxpart_dVd = dp_dVd; y.ddxpart_dVg = dp_dVg; y.ddxpart_dVs = dp_dVs; y.ddxpart_dVb = dp_dVb;
      y.sxpart = 1.0 - y.dxpart;
      y.dsxpart_dVd = -y.ddxpart_dVd; y.dsxpart_dVg = -y.ddxpart_dVg; y.dsxpart_dVs = -y.ddxpart_dVs;
      y.dsxpart_dVb = -(y.dsxpart_dVd + y.dsxpart_dVg + y.dsxpart_dVs);
    } else {
      // roles exchanged: what was computed against the "drain" columns belongs to the source terminal
      y.sxpart = p_; y.dsxpart_dVs = dp_dVd; y.dsxpart_dVg = dp_dVg; y.dsxpart_dVd = dp_dVs; y.dsxpart_dVb = dp_dVb;
      if (fabs(qcheq) <= 1.0e-5 * CoxWL) { y.dsxpart_dVb = 0.0; }
      else y.dsxpart_dVb = -(y.dsxpart_dVd + y.dsxpart_dVg + y.dsxpart_dVs);
      y.dxpart = 1.0 - y.sxpart;
      y.ddxpart_dVd = -y.dsxpart_dVd; y.ddxpart_dVg = -y.dsxpart_dVg; y.ddxpart_dVs = -y.dsxpart_dVs;
      y.ddxpart_dVb = -(y.ddxpart_dVd + y.ddxpart_dVg + y.ddxpart_dVs);
    }
    if (rgatemod == 3) {
      y.gcdgb = 0.0; y.gcsgb = 0.0; y.gcbgb = 0.0; y.gcggb = 0.0; y.gcgdb = 0.0; y.gcgsb = 0.0; y.gcgbb = 0.0;
    } else {
      y.gcggb = (cgdo + cgso + cgbo) * ag0;
      y.gcgdb = -cgdo * ag0;
      y.gcgsb = -cgso * ag0;
      y.gcgbb = -cgbo * ag0;
      y.gcdgb = y.gcgdb; y.gcsgb = y.gcgsb; y.gcbgb = y.gcgbb;
    }
    y.gcddb = (o.capbd + cgdo) * ag0;
    y.gcdsb = 0.0;
    y.gcsdb = 0.0;
    y.gcssb = (o.capbs + cgso) * ag0;
    if (rbodymod == 0) {
      y.gcdbb = -(y.gcdgb + y.gcddb + y.gcdgmb);
      y.gcsbb = -(y.gcsgb + y.gcssb + y.gcsgmb);
      y.gcbdb = -o.capbd * ag0;
      y.gcbsb = -o.capbs * ag0;
      gcdbdb = 0.0; y.gcsbsb = 0.0;
    } else {
      y.gcdbb = 0.0; y.gcsbb = 0.0; y.gcbdb = 0.0; y.gcbsb = 0.0;
      gcdbdb = -o.capbd * ag0; y.gcsbsb = -o.capbs * ag0;
    }
    y.gcbbb = -(y.gcbdb + y.gcbgb + y.gcbsb + y.gcbgmb);
  }

  // terminal currents i = dq/dt by Backward Euler against the committed charges (tran.rs:449-470)
  const double dt = e.dt;
  const double cqb = B4_DIV((o.qb - e.op(B4S_QB)), dt);
  const double cqg = B4_DIV((o.qg - e.op(B4S_QG)), dt);
  const double cqd = B4_DIV((o.qd - e.op(B4S_QD)), dt);
  double cqcdump = 0.0, cqgmid = 0.0, cqbs = 0.0, cqbd = 0.0;
  if (trnqsmod != 0) cqcdump = B4_DIV((o.qdef_dump - e.op(B4S_QCDUMP)), dt);
  if (rgatemod == 3) cqgmid = B4_DIV((o.qgmid - e.op(B4S_QGMID)), dt);
  if (rbodymod != 0) {
    cqbs = B4_DIV((o.qbs - e.op(B4S_QBS)), dt);
    cqbd = B4_DIV((o.qbd - e.op(B4S_QBD)), dt);
  }
  const double vgb = v.vgb, vbd = v.vbd, vbs = v.vbs, vgmb = v.vgmb;
  y.ceqqg = cqg - y.gcggb * vgb + y.gcgdb * vbd + y.gcgsb * vbs;
  y.ceqqd = cqd - y.gcdgb * vgb - y.gcdgmb * vgmb + (y.gcddb + gcdbdb) * vbd - gcdbdb * v.vbd_jct + y.gcdsb * vbs;
  y.ceqqb = cqb - y.gcbgb * vgb - y.gcbgmb * vgmb + y.gcbdb * vbd + y.gcbsb * vbs;
  if (rgatemod == 3) y.ceqqgmid = cqgmid + y.gcgmdb * vbd + y.gcgmsb * vbs - y.gcgmgmb * vgmb;
  if (rbodymod != 0) {
    y.ceqqjs = cqbs + y.gcsbsb * v.vbs_jct;
    y.ceqqjd = cqbd + gcdbdb * v.vbd_jct;
  }
  if (trnqsmod != 0) {
    const double cqcheq_i = B4_DIV((o.qcheq - e.op(B4S_QCHEQ)), dt);
    const double T0 = y.ggtg * vgb - y.ggtd * vbd - y.ggts * vbs;
    y.ceqqg += T0;
    const double T1 = v.qdef * o.gtau;
    y.ceqqd -= y.dxpart * T0 + T1 * (y.ddxpart_dVg * vgb - y.ddxpart_dVd * vbd - y.ddxpart_dVs * vbs);
    y.cqdef = cqcdump - y.gqdef * v.qdef;
    y.cqcheq = cqcheq_i - (y.gcqgb * vgb - y.gcqdb * vbd - y.gcqsb * vbs) + T0;
  }
}

B4_HD void b4_dyn_zero(B4Dyn& y) {
  y.gcdgb = 0.0; y.gcddb = 0.0; y.gcdsb = 0.0; y.gcdbb = 0.0; y.gcsgb = 0.0; y.gcsdb = 0.0; y.gcssb = 0.0; y.gcsbb = 0.0;
  y.gcggb = 0.0; y.gcgdb = 0.0; y.gcgsb = 0.0; y.gcgbb = 0.0; y.gcbdb = 0.0; y.gcbgb = 0.0; y.gcbsb = 0.0; y.gcbbb = 0.0;
  y.gcgmgmb = 0.0; y.gcgmdb = 0.0; y.gcgmsb = 0.0; y.gcgmbb = 0.0; y.gcdgmb = 0.0; y.gcsgmb = 0.0; y.gcbgmb = 0.0; y.gcsbsb = 0.0;
  y.gqdef = 0.0; y.gcqgb = 0.0; y.gcqdb = 0.0; y.gcqsb = 0.0; y.gcqbb = 0.0; y.ggtg = 0.0; y.ggtd = 0.0; y.ggtb = 0.0; y.ggts = 0.0;
  y.dxpart = 0.0; y.sxpart = 0.0;
  y.ddxpart_dVd = 0.0; y.ddxpart_dVg = 0.0; y.ddxpart_dVb = 0.0; y.ddxpart_dVs = 0.0;
  y.dsxpart_dVd = 0.0; y.dsxpart_dVg = 0.0; y.dsxpart_dVb = 0.0; y.dsxpart_dVs = 0.0;
  y.ceqqg = 0.0; y.ceqqd = 0.0; y.ceqqb = 0.0; y.ceqqjs = 0.0; y.ceqqjd = 0.0; y.ceqqgmid = 0.0; y.cqdef = 0.0; y.cqcheq = 0.0;
}

// ------------------------------------------------------------------------------------------------------------------
// Phase 7: MNA stamp (stamp.rs:8-572). Every push goes to its own itab slot (bsim4_layout.h).
template <class E> B4_HD void b4_stamp(E& e, const B4Bias& v, const B4Op& o, const B4Dyn& y) {
  const double tp = M_(type_sign);
  const int igcmod = (int)M_(igcmod), igbmod = (int)M_(igbmod), rgatemod = (int)M_(rgatemod), rdsmod = (int)M_(rdsmod);
  const int rbodymod = (int)M_(rbodymod), trnqsmod = (int)M_(trnqsmod);
  const double vds = v.vds, vgs = v.vgs, vbs = v.vbs, vgd = v.vgd, vbd = v.vbd;
  double Gm, Gmbs, FwdSum, RevSum, ceqdrn, ceqbd, ceqbs;
  double gbbdp, gbbsp, gbdpg, gbdpdp, gbdpb, gbdpsp, gbspg, gbspdp, gbspb, gbspsp;
  double gIstotg = 0.0, gIstotd = 0.0, gIstots = 0.0, gIstotb = 0.0, Istoteq = 0.0;
  double gIdtotg = 0.0, gIdtotd = 0.0, gIdtots = 0.0, gIdtotb = 0.0, Idtoteq = 0.0;
  double gIbtotg = 0.0, gIbtotd = 0.0, gIbtots = 0.0, gIbtotb = 0.0, Ibtoteq = 0.0;
  double gIgtotg = 0.0, gIgtotd = 0.0, gIgtots = 0.0, gIgtotb = 0.0, Igtoteq = 0.0;
  double ceqgcrg = 0.0, gcrg = 0.0, gcrgd = 0.0, gcrgg = 0.0, gcrgs = 0.0, gcrgb = 0.0;

  if (o.mode >= 0) {
    Gm = o.gm; Gmbs = o.gmbs;
    FwdSum = Gm + Gmbs; RevSum = 0.0;
    ceqdrn = tp * (o.cd - o.gds * vds - Gm * vgs - Gmbs * vbs);
    ceqbd = tp * (o.csub + o.Igidl - (o.gbds + o.ggidld) * vds - (o.gbgs + o.ggidlg) * vgs - (o.gbbs + o.ggidlb) * vbs);
    ceqbs = tp * (o.Igisl + o.ggisls * vds - o.ggislg * vgd - o.ggislb * vbd);
    gbbdp = -(o.gbds);
    gbbsp = o.gbds + o.gbgs + o.gbbs;
    gbdpg = o.gbgs; gbdpdp = o.gbds; gbdpb = o.gbbs;
    gbdpsp = -(gbdpg + gbdpdp + gbdpb);
    gbspg = 0.0; gbspdp = 0.0; gbspb = 0.0; gbspsp = 0.0;
    if (igcmod != 0) {
      gIstotg = o.gIgsg + o.gIgcsg;
      gIstotd = o.gIgcsd;
      gIstots = o.gIgss + o.gIgcss;
      gIstotb = o.gIgcsb;
      Istoteq = tp * (o.Igs + o.Igcs - gIstotg * vgs - o.gIgcsd * vds - o.gIgcsb * vbs);
      gIdtotg = o.gIgdg + o.gIgcdg;
      gIdtotd = o.gIgdd + o.gIgcdd;
      gIdtots = o.gIgcds;
      gIdtotb = o.gIgcdb;
      Idtoteq = tp * (o.Igd + o.Igcd - o.gIgdg * vgd - o.gIgcdg * vgs - o.gIgcdd * vds - o.gIgcdb * vbs);
    }
    if (igbmod != 0) {
      gIbtotg = o.gIgbg; gIbtotd = o.gIgbd; gIbtots = o.gIgbs; gIbtotb = o.gIgbb;
      Ibtoteq = tp * (o.Igb - o.gIgbg * vgs - o.gIgbd * vds - o.gIgbb * vbs);
    }
    if (rgatemod > 1) {
      const double tmp = rgatemod == 2 ? v.vges - vgs : v.vgms - vgs;
      gcrgd = o.gcrgd * tmp; gcrgg = o.gcrgg * tmp; gcrgs = o.gcrgs * tmp; gcrgb = o.gcrgb * tmp;
      ceqgcrg = -(gcrgd * vds + gcrgg * vgs + gcrgb * vbs);
      gcrgg -= o.gcrg;
      gcrg = o.gcrg;
    }
  } else {
    Gm = -o.gm; Gmbs = -o.gmbs;
    FwdSum = 0.0; RevSum = -(Gm + Gmbs);
    ceqdrn = -tp * (o.cd + o.gds * vds + Gm * vgd + Gmbs * vbd);
    ceqbs = tp * (o.csub + o.Igisl + (o.gbds + o.ggisls) * vds - (o.gbgs + o.ggislg) * vgd - (o.gbbs + o.ggislb) * vbd);
    ceqbd = tp * (o.Igidl - o.ggidld * vds - o.ggidlg * vgs - o.ggidlb * vbs);
    gbbsp = -(o.gbds);
    gbbdp = o.gbds + o.gbgs + o.gbbs;
    gbdpg = 0.0; gbdpsp = 0.0; gbdpb = 0.0; gbdpdp = 0.0;
    gbspg = o.gbgs; gbspsp = o.gbds; gbspb = o.gbbs;
    gbspdp = -(gbspg + gbspsp + gbspb);
    if (igcmod != 0) {
      gIstotg = o.gIgsg + o.gIgcdg;
      gIstotd = o.gIgcds;
      gIstots = o.gIgss + o.gIgcdd;
      gIstotb = o.gIgcdb;
      Istoteq = tp * (o.Igs + o.Igcd - o.gIgsg * vgs - o.gIgcdg * vgd + o.gIgcdd * vds - o.gIgcdb * vbd);
      gIdtotg = o.gIgdg + o.gIgcsg;
      gIdtotd = o.gIgdd + o.gIgcss;
      gIdtots = o.gIgcsd;
      gIdtotb = o.gIgcsb;
      Idtoteq = tp * (o.Igd + o.Igcs - (o.gIgdg + o.gIgcsg) * vgd + o.gIgcsd * vds - o.gIgcsb * vbd);
    }
    if (igbmod != 0) {
      gIbtotg = o.gIgbg; gIbtotd = o.gIgbs; gIbtots = o.gIgbd; gIbtotb = o.gIgbb;
      Ibtoteq = tp * (o.Igb - o.gIgbg * vgd + o.gIgbd * vds - o.gIgbb * vbd);
    }
    if (rgatemod > 1) {
      const double tmp = rgatemod == 2 ? v.vges - vgs : v.vgms - vgs;
      gcrgd = o.gcrgs * tmp; gcrgg = o.gcrgg * tmp; gcrgs = o.gcrgd * tmp; gcrgb = o.gcrgb * tmp;
      ceqgcrg = -(gcrgg * vgd - gcrgs * vds + gcrgb * vbd);
      gcrgg -= o.gcrg;
      gcrg = o.gcrg;
    }
  }
  if (igcmod != 0 || igbmod != 0) {
    gIgtotg = gIstotg + gIdtotg + gIbtotg;
    gIgtotd = gIstotd + gIdtotd + gIbtotd;
    gIgtots = gIstots + gIdtots + gIbtots;
    gIgtotb = gIstotb + gIdtotb + gIbtotb;
    Igtoteq = Istoteq + Idtoteq + Ibtoteq;
  }

  double gstot = 0.0, gstotd = 0.0, gstotg = 0.0, gstots = 0.0, gstotb = 0.0, ceqgstot = 0.0;
  double gdtot = 0.0, gdtotd = 0.0, gdtotg = 0.0, gdtots = 0.0, gdtotb = 0.0, ceqgdtot = 0.0;
  if (rdsmod == 1) {
    ceqgstot = tp * (o.gstotd * vds + o.gstotg * vgs + o.gstotb * vbs);
    gstot = o.gstot; gstotd = o.gstotd; gstotg = o.gstotg; gstots = o.gstots - gstot; gstotb = o.gstotb;
    ceqgdtot = -tp * (o.gdtotd * vds + o.gdtotg * vgs + o.gdtotb * vbs);
    gdtot = o.gdtot; gdtotd = o.gdtotd - gdtot; gdtotg = o.gdtotg; gdtots = o.gdtots; gdtotb = o.gdtotb;
  }

  double ceqjs = o.cbs - o.gbs * v.vbs_jct, ceqjd = o.cbd - o.gbd * v.vbd_jct;
  double ceqqg = y.ceqqg, ceqqd = y.ceqqd, ceqqb = y.ceqqb, cqdef = y.cqdef, cqcheq = y.cqcheq;
  double ceqqjs = y.ceqqjs, ceqqjd = y.ceqqjd, ceqqgmid = y.ceqqgmid;
  if (tp < 0.0) {
    ceqjs = -ceqjs; ceqjd = -ceqjd;
    ceqqg = -ceqqg; ceqqd = -ceqqd; ceqqb = -ceqqb; ceqgcrg = -ceqgcrg;
    if (trnqsmod != 0) { cqdef = -cqdef; cqcheq = -cqcheq; }
    if (rbodymod != 0) { ceqqjs = -ceqqjs; ceqqjd = -ceqqjd; }
    if (rgatemod == 3) ceqqgmid = -ceqqgmid;
  }

  // ---- right-hand side
  e.add_b_at(B4B_DP, (ceqjd - ceqbd + ceqgdtot - ceqdrn - ceqqd + Idtoteq));
  e.add_b_at(B4B_GP, -(ceqqg - ceqgcrg + Igtoteq));
  if (rgatemod == 2) e.add_b_at(B4B_GX, -ceqgcrg);
  else if (rgatemod == 3) e.add_b_at(B4B_GX, -(ceqqgmid + ceqgcrg));
  if (rbodymod == 0) {
    e.add_b_at(B4B_BP, (ceqbd + ceqbs - ceqjd - ceqjs - ceqqb + Ibtoteq));
    e.add_b_at(B4B_SP, (ceqdrn - ceqbs + ceqjs + ceqqg + ceqqb + ceqqd + ceqqgmid - ceqgstot + Istoteq));
  } else {
    e.add_b_at(B4B_DB, -(ceqjd + ceqqjd));
    e.add_b_at(B4B_BP, (ceqbd + ceqbs - ceqqb + Ibtoteq));
    e.add_b_at(B4B_SB, -(ceqjs + ceqqjs));
    e.add_b_at(B4B_SP, (ceqdrn - ceqbs + ceqjs + ceqqd + ceqqg + ceqqb + ceqqjd + ceqqjs + ceqqgmid - ceqgstot + Istoteq));
  }
  if (rdsmod != 0) {
    e.add_b_at(B4B_D, -ceqgdtot);
    e.add_b_at(B4B_S, ceqgstot);
  }
  if (trnqsmod != 0) e.add_b_at(B4B_Q, cqcheq - cqdef);

  // ---- Jacobian
  const double gjbd = rbodymod != 0 ? o.gbd : 0.0, gjbs = rbodymod != 0 ? o.gbs : 0.0;
  const double gdpr = rdsmod != 0 ? I_(drainConductance) : 0.0, gspr = rdsmod != 0 ? I_(sourceConductance) : 0.0;
  if (rgatemod == 1) {
    const double geltd = I_(grgeltd);
    e.add_g_at(B4G_GEge, geltd);
    e.add_g_at(B4G_GPge, -(geltd));
    e.add_g_at(B4G_GEgp, -(geltd));
    e.add_g_at(B4G_GPgp, y.gcggb + geltd - y.ggtg + gIgtotg);
    e.add_g_at(B4G_GPdp, y.gcgdb - y.ggtd + gIgtotd);
    e.add_g_at(B4G_GPsp, y.gcgsb - y.ggts + gIgtots);
    e.add_g_at(B4G_GPbp, y.gcgbb - y.ggtb + gIgtotb);
  } else if (rgatemod == 2) {
    e.add_g_at(B4G_GEge, gcrg);
    e.add_g_at(B4G_GEgp, gcrgg);
    e.add_g_at(B4G_GEdp, gcrgd);
    e.add_g_at(B4G_GEsp, gcrgs);
    e.add_g_at(B4G_GEbp, gcrgb);
    e.add_g_at(B4G_GPge, -(gcrg));
    e.add_g_at(B4G_GPgp, y.gcggb - gcrgg - y.ggtg + gIgtotg);
    e.add_g_at(B4G_GPdp, y.gcgdb - gcrgd - y.ggtd + gIgtotd);
    e.add_g_at(B4G_GPsp, y.gcgsb - gcrgs - y.ggts + gIgtots);
    e.add_g_at(B4G_GPbp, y.gcgbb - gcrgb - y.ggtb + gIgtotb);
  } else if (rgatemod == 3) {
    const double geltd = I_(grgeltd);
    e.add_g_at(B4G_GEge, geltd);
    e.add_g_at(B4G_GEgm, -(geltd));
    e.add_g_at(B4G_GMge, -(geltd));
    e.add_g_at(B4G_GMgm, geltd + gcrg + y.gcgmgmb);
    e.add_g_at(B4G_GMdp, gcrgd + y.gcgmdb);
    e.add_g_at(B4G_GMgp, gcrgg);
    e.add_g_at(B4G_GMsp, gcrgs + y.gcgmsb);
    e.add_g_at(B4G_GMbp, gcrgb + y.gcgmbb);
    e.add_g_at(B4G_DPgm, y.gcdgmb);
    e.add_g_at(B4G_GPgm, -(gcrg));
    e.add_g_at(B4G_SPgm, y.gcsgmb);
    e.add_g_at(B4G_BPgm, y.gcbgmb);
    e.add_g_at(B4G_GPgp, y.gcggb - gcrgg - y.ggtg + gIgtotg);
    e.add_g_at(B4G_GPdp, y.gcgdb - gcrgd - y.ggtd + gIgtotd);
    e.add_g_at(B4G_GPsp, y.gcgsb - gcrgs - y.ggts + gIgtots);
    e.add_g_at(B4G_GPbp, y.gcgbb - gcrgb - y.ggtb + gIgtotb);
  } else {
    e.add_g_at(B4G_GPgp, y.gcggb - y.ggtg + gIgtotg);
    e.add_g_at(B4G_GPdp, y.gcgdb - y.ggtd + gIgtotd);
    e.add_g_at(B4G_GPsp, y.gcgsb - y.ggts + gIgtots);
    e.add_g_at(B4G_GPbp, y.gcgbb - y.ggtb + gIgtotb);
  }
  if (rdsmod != 0) {
    e.add_g_at(B4G_Dgp, gdtotg);
    e.add_g_at(B4G_Dsp, gdtots);
    e.add_g_at(B4G_Dbp, gdtotb);
    e.add_g_at(B4G_Sdp, gstotd);
    e.add_g_at(B4G_Sgp, gstotg);
    e.add_g_at(B4G_Sbp, gstotb);
  }
  {
    const double tmp1 = v.qdef * o.gtau;
    const double gcdbdb = 0.0;  // never carried over from the transient model in the reference
    e.add_g_at(B4G_DPdp, gdpr + o.gds + o.gbd + tmp1 * y.ddxpart_dVd - gdtotd + RevSum + y.gcddb + gbdpdp + y.dxpart * y.ggtd - gIdtotd);
    e.add_g_at(B4G_DPd, -(gdpr + gdtot));
    e.add_g_at(B4G_DPgp, Gm + y.gcdgb - gdtotg + gbdpg - gIdtotg + y.dxpart * y.ggtg + tmp1 * y.ddxpart_dVg);
    e.add_g_at(B4G_DPsp, -(o.gds + gdtots - y.dxpart * y.ggts + gIdtots - tmp1 * y.ddxpart_dVs + FwdSum - y.gcdsb - gbdpsp));
    e.add_g_at(B4G_DPbp, -(gjbd + gdtotb - Gmbs - y.gcdbb - gbdpb + gIdtotb - tmp1 * y.ddxpart_dVb - y.dxpart * y.ggtb));
    e.add_g_at(B4G_Ddp, -(gdpr - gdtotd));
    e.add_g_at(B4G_Dd, gdpr + gdtot);
    e.add_g_at(B4G_SPdp, -(o.gds + gstotd + RevSum - y.gcsdb - gbspdp - tmp1 * y.dsxpart_dVd - y.sxpart * y.ggtd + gIstotd));
    e.add_g_at(B4G_SPgp, y.gcsgb - Gm - gstotg + gbspg + y.sxpart * y.ggtg + tmp1 * y.dsxpart_dVg - gIstotg);
    e.add_g_at(B4G_SPsp, gspr + o.gds + o.gbs + tmp1 * y.dsxpart_dVs - gstots + FwdSum + y.gcssb + gbspsp + y.sxpart * y.ggts - gIstots);
    e.add_g_at(B4G_SPs, -(gspr + gstot));
    e.add_g_at(B4G_SPbp, -(gjbs + gstotb + Gmbs - y.gcsbb - gbspb - y.sxpart * y.ggtb - tmp1 * y.dsxpart_dVb + gIstotb));
    e.add_g_at(B4G_Ssp, -(gspr - gstots));
    e.add_g_at(B4G_Ss, gspr + gstot);
    e.add_g_at(B4G_BPdp, y.gcbdb - gjbd + gbbdp - gIbtotd);
    e.add_g_at(B4G_BPgp, y.gcbgb - o.gbgs - gIbtotg);
    e.add_g_at(B4G_BPsp, y.gcbsb - gjbs + gbbsp - gIbtots);
    e.add_g_at(B4G_BPbp, gjbd + gjbs + y.gcbbb - o.gbbs - gIbtotb);

    const double ggidld = o.ggidld, ggidlg = o.ggidlg, ggidlb = o.ggidlb, ggislg = o.ggislg, ggisls = o.ggisls, ggislb = o.ggislb;
    e.add_g_at(B4G_L_DPdp, ggidld);
    e.add_g_at(B4G_L_DPgp, ggidlg);
    e.add_g_at(B4G_L_DPsp, -(ggidlg + ggidld + ggidlb));
    e.add_g_at(B4G_L_DPbp, ggidlb);
    e.add_g_at(B4G_L_BPdp, -(ggidld));
    e.add_g_at(B4G_L_BPgp, -(ggidlg));
    e.add_g_at(B4G_L_BPsp, (ggidlg + ggidld + ggidlb));
    e.add_g_at(B4G_L_BPbp, -(ggidlb));
    e.add_g_at(B4G_S_SPdp, -(ggisls + ggislg + ggislb));
    e.add_g_at(B4G_S_SPgp, ggislg);
    e.add_g_at(B4G_S_SPsp, ggisls);
    e.add_g_at(B4G_S_SPbp, ggislb);
    e.add_g_at(B4G_S_BPdp, (ggislg + ggisls + ggislb));
    e.add_g_at(B4G_S_BPgp, -(ggislg));
    e.add_g_at(B4G_S_BPsp, -(ggisls));
    e.add_g_at(B4G_S_BPbp, -(ggislb));

    if (rbodymod != 0) {
      const double grbpd = I_(grbpd), grbdb = I_(grbdb), grbpb = I_(grbpb), grbps = I_(grbps), grbsb = I_(grbsb);
      e.add_g_at(B4G_DPdb, gcdbdb - o.gbd);
      e.add_g_at(B4G_SPsb, -(o.gbs - y.gcsbsb));
      e.add_g_at(B4G_DBdp, gcdbdb - o.gbd);
      e.add_g_at(B4G_DBdb, o.gbd - gcdbdb + grbpd + grbdb);
      e.add_g_at(B4G_DBbp, -(grbpd));
      e.add_g_at(B4G_DBb, -(grbdb));
      e.add_g_at(B4G_BPdb, -(grbpd));
      e.add_g_at(B4G_BPb, -(grbpb));
      e.add_g_at(B4G_BPsb, -(grbps));
      e.add_g_at(B4G_R_BPbp, grbpd + grbps + grbpb);
      e.add_g_at(B4G_SBsp, y.gcsbsb - o.gbs);
      e.add_g_at(B4G_SBbp, -(grbps));
      e.add_g_at(B4G_SBb, -(grbsb));
      e.add_g_at(B4G_SBsb, o.gbs - y.gcsbsb + grbps + grbsb);
      e.add_g_at(B4G_Bdb, -(grbdb));
      e.add_g_at(B4G_Bbp, -(grbpb));
      e.add_g_at(B4G_Bsb, -(grbsb));
      e.add_g_at(B4G_Bb, grbsb + grbdb + grbpb);
    }
  }
  if (trnqsmod != 0) {
    e.add_g_at(B4G_Qq, y.gqdef + o.gtau);
    e.add_g_at(B4G_Qgp, y.ggtg - y.gcqgb);
    e.add_g_at(B4G_Qdp, y.ggtd - y.gcqdb);
    e.add_g_at(B4G_Qsp, y.ggts - y.gcqsb);
    e.add_g_at(B4G_Qbp, y.ggtb - y.gcqbb);
    e.add_g_at(B4G_DPq, y.dxpart * o.gtau);
    e.add_g_at(B4G_SPq, y.sxpart * o.gtau);
    e.add_g_at(B4G_GPq, -(o.gtau));
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Component::load for a BSIM4 instance (bsim4solver.rs:134-144, 3761-3767)
template <class E> B4_HD void load_bsim4(E& e) {
  B4Bias v;
  B4Op o;
  B4Chan c;
  B4Tunnel t;
  B4Dyn y;
  b4_limit_bias(e, v);
  b4_junction_dc(e, v, o);
  b4_channel_dc(e, v, o, c);
  b4_leakage(e, v, o, c, t);
  b4_charge(e, v, o, c, t);
  o.qdef_dump = v.qdef * 1.0e-9;
  if (e.mode == AN_TRAN) b4_tran_caps(e, v, o, y);
  else b4_dyn_zero(y);
  b4_stamp(e, v, o, y);
  // the new in-flight operating point (what `self.guess = newop` keeps and later evaluations read back)
  e.set_guess(B4S_VGS, v.vgs); e.set_guess(B4S_VDS, v.vds); e.set_guess(B4S_VBS, v.vbs);
  e.set_guess(B4S_VGES, v.vges); e.set_guess(B4S_VGMS, v.vgms); e.set_guess(B4S_VDBS, v.vdbs); e.set_guess(B4S_VSBS, v.vsbs);
  e.set_guess(B4S_VSES, v.vses); e.set_guess(B4S_VDES, v.vdes); e.set_guess(B4S_VBD, v.vbd); e.set_guess(B4S_VDBD, v.vdbd);
  e.set_guess(B4S_VON, o.von);
  e.set_guess(B4S_QB, o.qb); e.set_guess(B4S_QG, o.qg); e.set_guess(B4S_QD, o.qd); e.set_guess(B4S_QGMID, o.qgmid);
  e.set_guess(B4S_QBS, o.qbs); e.set_guess(B4S_QBD, o.qbd);
  e.set_guess(B4S_QCDUMP, e.mode == AN_TRAN && M_(trnqsmod) != 0.0 ? o.qdef_dump : 0.0);
  e.set_guess(B4S_QCHEQ, o.qcheq);
}

#undef M_
#undef D_
#undef S_
#undef I_

}  // namespace b4e
}  // namespace s21
