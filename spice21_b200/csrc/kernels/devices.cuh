// Device-evaluation ("load") functions: the CUDA side of the reference's `Component::load` / `load_ac` plugin surface
// (spice21/src/comps/mod.rs:74-93). One function per device type; each reads its terminal voltages, its parameter
// columns and its committed state through an `Env`, and pushes its G and b stamps in exactly the reference's push
// order, so that per-element accumulation order matches `Solver::update` (analysis.rs:153-168).
// Arithmetic is written expression-for-expression after the reference and compiled with -fmad=false: Rust never
// contracts a*b+c, and the oracle is built with -ffp-contract=off.
//
// Env concept (see newton.cu for the per-instance implementation):
//   int    node(k)            k-th itab entry (variable index or element handle)
//   double par(k)             k-th parameter (shared or per-instance column)
//   double volt(var)          real solution entry, 0 for ground (-1)
//   double op(k), guess(k)    committed / in-flight state slot k of this device;  set_guess(k, v)
//   void   add_g_at(pos, v)   push a matrix stamp for the element handle stored at itab position `pos`
//   void   add_b_at(pos, v)   push an RHS stamp for the variable stored at itab position `pos`
//   void   add_g_dup(pos, dup, v)  second push to the same element within one load (own staging slot `dup`)
// A direct Env adds straight into A / rhs (one thread owns the instance); a staged Env writes private per-device
// staging slots that a later gather phase sums in the reference's order (devices evaluated in parallel).
//   int mode; double dt, gmin, omega
#pragma once
#include "../device_layout.h"
#include "../scalar.h"

namespace s21 {

// Math hooks. An Env may bring its own division / sqrt / exp (m_div, m_sqrt, m_exp): the run-time specialised team
// kernel (host/jit_team.hpp) uses that to evaluate Mos1 with the branch-free forms of scalar.h and a deferred
// exception flag. Every other Env gets the exact s_div and the library calls — the same bits either way.
template <class Env> __device__ __forceinline__ auto e_div_(Env& e, double a, double b, int) -> decltype(e.m_div(a, b)) { return e.m_div(a, b); }
template <class Env> __device__ __forceinline__ double e_div_(Env&, double a, double b, long) { return s_div(a, b); }
template <class Env> __device__ __forceinline__ auto e_sqrt_(Env& e, double a, int) -> decltype(e.m_sqrt(a)) { return e.m_sqrt(a); }
template <class Env> __device__ __forceinline__ double e_sqrt_(Env&, double a, long) { return sqrt(a); }
template <class Env> __device__ __forceinline__ auto e_exp_(Env& e, double a, int) -> decltype(e.m_exp(a)) { return e.m_exp(a); }
template <class Env> __device__ __forceinline__ double e_exp_(Env&, double a, long) { return exp(a); }
#define E_DIV(a, b) e_div_(e, (a), (b), 0)
#define E_SQRT(a) e_sqrt_(e, (a), 0)
#define E_EXP(a) e_exp_(e, (a), 0)

// TranState::integrate, Backward Euler (analysis.rs:420-428): g = dq_dv/dt, i = dq/dt, rhs = i - g*vguess
struct Integ { double g, i, rhs; };
template <class Env> __device__ __forceinline__ Integ integrate_be_e(Env& e, double dt, double dq, double dq_dv, double vguess) {
  Integ r;
  r.g = E_DIV(dq_dv, dt);
  r.i = E_DIV(dq, dt);
  r.rhs = r.i - r.g * vguess;
  return r;
}
__device__ __forceinline__ Integ integrate_be(double dt, double dq, double dq_dv, double vguess) {
  Integ r;
  r.g = s_div(dq_dv, dt);  // s_div: exact, with a shortcut for the (frequent) zero numerator
  r.i = s_div(dq, dt);
  r.rhs = r.i - r.g * vguess;
  return r;
}

// ---------------------------------------------------------------- Resistor (comps/mod.rs:299-310)
template <class Env> __device__ __forceinline__ void load_resistor(Env& e) {
  const double g = e.par(e.mode == AN_OP ? RP_G_OP : RP_G_TRAN);
  e.add_g_at(R_EPP, g);
  e.add_g_at(R_ENN, g);
  e.add_g_at(R_EPN, -g);
  e.add_g_at(R_ENP, -g);
}
// ---------------------------------------------------------------- Capacitor (comps/mod.rs:200-222)
template <class Env> __device__ __forceinline__ void load_capacitor(Env& e) {
  const double vd = e.volt(e.node(R_P)) - e.volt(e.node(R_N));
  const double c = e.par(CP_C);
  const double q = c * vd;
  e.set_guess(CS_Q, q);
  if (e.mode == AN_OP) return;  // no stamps in OP; only records guess{v, q}
  const Integ k = integrate_be(e.dt, q - e.op(CS_Q), c, vd);
  e.add_g_at(R_EPP, k.g);
  e.add_g_at(R_ENN, k.g);
  e.add_g_at(R_EPN, -k.g);
  e.add_g_at(R_ENP, -k.g);
  e.add_b_at(R_P, -k.rhs);
  e.add_b_at(R_N, k.rhs);
}
// ---------------------------------------------------------------- Isrc (comps/mod.rs:340-345)
template <class Env> __device__ __forceinline__ void load_isrc(Env& e) {
  const double i = e.par(IP_I);
  e.add_b_at(I_P, i);
  e.add_b_at(I_N, -i);
}
// ---------------------------------------------------------------- Vsrc (comps/mod.rs:133-138)
// Time-varying source value at time t (extension, device_layout.h VP_WKIND): SPICE's PULSE and SIN.
template <class Env> __device__ __noinline__ double source_wave(Env& e, int kind, double t) {
  const double w0 = e.par(VP_W0), w1 = e.par(VP_W1), w2 = e.par(VP_W2), w3 = e.par(VP_W3), w4 = e.par(VP_W4), w5 = e.par(VP_W5), w6 = e.par(VP_W6);
  if (kind == SRC_PULSE) {  // v1 v2 td tr tf pw per
    double tt = t - w2;
    if (tt < 0.0) return w0;
    if (w6 > 0.0) tt -= w6 * floor(tt / w6);
    if (tt < w3) return w0 + (w1 - w0) * (tt / w3);
    if (tt < w3 + w5) return w1;
    if (tt < w3 + w5 + w4) return w1 + (w0 - w1) * ((tt - w3 - w5) / w4);
    return w0;
  }
  // SIN: vo va freq td theta
  const double tt = t - w3;
  if (tt < 0.0) return w0;
  return w0 + w1 * exp(-tt * w4) * sin(6.283185307179586476925286766559 * w2 * tt);
}
template <class Env> __device__ __forceinline__ void load_vsrc(Env& e) {
  e.add_g_at(V_EPI, 1.0);
  e.add_g_at(V_EIP, 1.0);
  e.add_g_at(V_ENI, -1.0);
  e.add_g_at(V_EIN, -1.0);
  double v = e.par(e.mode == AN_OP ? VP_V_OP : VP_V_TRAN);
  if (e.mode == AN_TRAN) {
    const int kind = (int)e.par(VP_WKIND);
    if (kind != SRC_NONE) v = source_wave(e, kind, e.time);
  }
  e.add_b_at(V_I, v);
}
// ---------------------------------------------------------------- Diode (comps/diode.rs:228-245, 279-355)
template <class Env> __device__ __forceinline__ double diode_limit(Env& e, double vd, double vold) {
  const double vcrit = e.par(DP_VCRIT), vte = e.par(DP_VTE);
  const double vnew = vd;
  if (vnew <= vcrit || fabs(vnew - vold) <= 2.0 * vte) return vnew;
  if (vold > 0.0) {
    const double arg = 1.0 + s_div(vnew - vold, vte);
    if (arg > 0.0) return vold + vte * log(arg);
    return vcrit;
  }
  return vte * log(vnew / vte);
}
template <class Env> __device__ __forceinline__ void load_diode(Env& e) {
  const double vte = e.par(DP_VTE), isat = e.par(DP_ISAT), gspr = e.par(DP_GSPR), bv = e.par(DP_BV);
  const bool has_bv = e.par(DP_HASBV) != 0.0;
  const double tt = e.par(DP_TT), vj = e.par(DP_VJ), m = e.par(DP_M), cz = e.par(DP_CZ);
  const double gmin = e.gmin;
  const int r = e.node(D_R), n = e.node(D_N);
  double vd = e.volt(r) - e.volt(n);
  const double vprev = e.guess(DS_VD);  // iteration-carried: the previous load's (limited) vd
  if (has_bv && vd < fmin(10.0 * vte - bv, 0.0)) {
    const double vtemp = diode_limit(e, -bv, bv - vprev);
    vd = vtemp - bv;
  } else {
    vd = diode_limit(e, vd, vprev);
  }
  double id, gd;
  if (!has_bv || vd >= -bv) {
    const double ex = exp(s_div(vd, vte));
    id = isat * (ex - 1.0) + gmin * vd;
    gd = isat * ex / vte + gmin;
  } else {
    const double ex = exp(s_div(vd - bv, vte));
    id = -isat * ex + gmin * vd;
    gd = isat * ex / vte + gmin;
  }
  double qd, cd;
  const double dep = e.par(DP_DEPTH);
  if (vd < dep) {
    const double a = 1.0 - s_div(vd, vj);
    const double s = -m * log(a);
    qd = tt * vj * cz * (1.0 - a * s) / (1.0 - m);
    cd = tt * gd + cz * s;
  } else {
    const double cz2 = e.par(DP_CZ2), f3 = e.par(DP_F3);
    qd = tt * id + cz * e.par(DP_F1) + cz2 * (f3 * (vd - dep) + m / 2.0 / vj * (vd * vd - dep * dep));
    cd = tt + cz2 * f3 + m * vd / vj;
  }
  double gc = 0.0, ic = 0.0;
  if (e.mode == AN_TRAN) {
    const Integ k = integrate_be(e.dt, qd - e.op(DS_CHARGE), cd, vd);
    gc = k.g;
    ic = k.i;
  }
  id += ic;
  gd += gc;
  e.set_guess(DS_VD, vd);
  e.set_guess(DS_CHARGE, qd);
  const double irhs = id - vd * gd;
  e.add_g_at(D_ENN, gd);
  e.add_g_at(D_ERN, -gd);
  e.add_g_at(D_ENR, -gd);
  e.add_g_at(D_ERR, gd + gspr);
  e.add_g_at(D_EPP, gspr);
  e.add_g_at(D_EPR, -gspr);
  e.add_g_at(D_ERP, -gspr);
  e.add_b_at(D_R, -irhs);
  e.add_b_at(D_N, irhs);
}
// ---------------------------------------------------------------- Mos0 (comps/mos.rs:1051-1099)
template <class Env> __device__ __forceinline__ void load_mos0(Env& e) {
  const double vth = 0.25, beta = 50e-3, lam = 3e-3;  // Mos0Params::default (mos.rs:1014-1023)
  const double gmin = e.gmin;
  const double vg = e.volt(e.node(M0_G)), vd = e.volt(e.node(M0_D)), vs = e.volt(e.node(M0_S));
  const double p = e.par(M0P_P);
  const double vds1 = p * (vd - vs);
  const bool reversed = vds1 < 0.0;
  const double vgs = reversed ? p * (vg - vd) : p * (vg - vs);
  const double vds = reversed ? -vds1 : vds1;
  const double vov = vgs - vth;
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
  const double irhs = ids - gm * vgs - gds * vds;
  // (sr, dr) = (S, D) or swapped
  const int e_drdr = reversed ? M0_ESS : M0_EDD, e_srsr = reversed ? M0_EDD : M0_ESS;
  const int e_drsr = reversed ? M0_ESD : M0_EDS, e_srdr = reversed ? M0_EDS : M0_ESD;
  const int e_drg = reversed ? M0_ESG : M0_EDG, e_srg = reversed ? M0_EDG : M0_ESG;
  e.add_g_at(e_drdr, gds + gmin);
  e.add_g_at(e_srsr, (gm + gds + gmin));
  e.add_g_at(e_drsr, -(gm + gds + gmin));
  e.add_g_at(e_srdr, -gds - gmin);
  e.add_g_at(e_drg, gm);
  e.add_g_at(e_srg, -gm);
  e.add_b_at(reversed ? M0_S : M0_D, -p * irhs);
  e.add_b_at(reversed ? M0_D : M0_S, p * irhs);
}
// ---------------------------------------------------------------- Mos1 (comps/mos.rs:504-521, 649-893)
// MosJunction::qc (mos.rs:504-521) feeds only tr.bs/tr.bd and op.cbs/op.cbd, none of which is ever stamped or read back
// by the reference (mos.rs:825-826, 929-930), so the junction capacitance is not evaluated here.
template <class Env> __device__ __forceinline__ void load_mos1(Env& e) {
  const double gmin = e.gmin;
  const double p = e.par(M1P_P);
  const double v_d = e.volt(e.node(M1_D)), v_g = e.volt(e.node(M1_G)), v_s = e.volt(e.node(M1_S)), v_b = e.volt(e.node(M1_B));
  const bool reversed = p * (v_d - v_s) < 0.0;
  const double vd = reversed ? v_s : v_d, vs = reversed ? v_d : v_s;
  const double vgs = p * (v_g - vs);
  const double vgd = p * (v_g - vd);
  const double vds = p * (vd - vs);
  const double vgb = p * (v_g - v_b);
  const double vsb = p * (vs - v_b);
  const double vdb = p * (vd - v_b);
  const double vt0_t = e.par(M1P_VT0T), phi_t = e.par(M1P_PHIT), gamma = e.par(M1P_GAMMA), beta = e.par(M1P_BETA);
  const double lambda = e.par(M1P_LAMBDA);
  const double von = vsb > 0.0 ? vt0_t + gamma * (E_SQRT(phi_t + vsb) - E_SQRT(phi_t)) : vt0_t;
  const double vov = vgs - von;
  const double vdsat = fmax(vov, 0.0);
  double ids = 0.0, gm = 0.0, gds = 0.0, gmbs = 0.0;
  if (vov > 0.0) {
    if (vds >= vov) {
      ids = beta / 2.0 * (vov * vov) * (1.0 + lambda * vds);
      gm = beta * vov * (1.0 + lambda * vds);
      gds = lambda * beta / 2.0 * (vov * vov);
    } else {
      ids = beta * (vov * vds - (vds * vds) / 2.0) * (1.0 + lambda * vds);
      gm = beta * vds * (1.0 + lambda * vds);
      gds = beta * ((vov - vds) * (1.0 + lambda * vds) + lambda * ((vov * vds) - (vds * vds) / 2.0));
    }
    gmbs = (phi_t + vsb > 0.0) ? E_DIV(gm * gamma / 2.0, E_SQRT(phi_t + vsb)) : 0.0;
  }
  // bulk junction diodes; the junction parameter blocks swap with the channel direction (mos.rs:710-714)
  const double vtherm = e.par(M1P_VTHERM);
  const int bs_j = reversed ? M1P_DJ : M1P_SJ, bd_j = reversed ? M1P_SJ : M1P_DJ;
  const double bs_isat = e.par(bs_j + MJ_ISAT), bd_isat = e.par(bd_j + MJ_ISAT);
  const double ebs = E_EXP(E_DIV(-vsb, vtherm));
  const double ibs = bs_isat * (ebs - 1.0);
  const double gbs = E_DIV(bs_isat, vtherm) * ebs + gmin;
  const double ibs_rhs = ibs + vsb * gbs;
  const double ebd = E_EXP(E_DIV(-vdb, vtherm));
  const double ibd = bd_isat * (ebd - 1.0);
  const double gbd = E_DIV(bd_isat, vtherm) * ebd + gmin;
  const double ibd_rhs = ibd + vdb * gbd;
  // Meyer gate capacitances (mos.rs:725-752)
  const double cox = e.par(M1P_COX);
  double cgs1, cgd1, cgb1;
  if (vov <= -phi_t) {
    cgb1 = cox / 2.0; cgs1 = 0.0; cgd1 = 0.0;
  } else if (vov <= -phi_t / 2.0) {
    cgb1 = E_DIV(-vov * cox, 2.0 * phi_t); cgs1 = 0.0; cgd1 = 0.0;
  } else if (vov <= 0.0) {
    cgb1 = E_DIV(-vov * cox, 2.0 * phi_t);
    cgs1 = E_DIV(vov * cox, 1.5 * phi_t) + E_DIV(cox, 3.0);
    cgd1 = 0.0;
  } else if (vdsat <= vds) {
    cgs1 = E_DIV(cox, 3.0); cgd1 = 0.0; cgb1 = 0.0;
  } else {
    const double vddif = 2.0 * vdsat - vds;
    const double vddif1 = vdsat - vds;
    const double vddif2 = vddif * vddif;
    cgd1 = E_DIV(cox * (1.0 - E_DIV(vdsat * vdsat, vddif2)), 3.0);
    cgs1 = E_DIV(cox * (1.0 - E_DIV(vddif1 * vddif1, vddif2)), 3.0);
    cgb1 = 0.0;
  }
  // history averaging against the committed point (mos.rs:757-767)
  const double op_cgs = e.op(M1S_CGS), op_cgd = e.op(M1S_CGD), op_cgb = e.op(M1S_CGB);
  const bool same_dir = reversed == (e.op(M1S_REV) != 0.0);
  const double cgs2 = (op_cgs == 0.0) ? cgs1 : (same_dir ? op_cgs : op_cgd);
  const double cgs = cgs1 + cgs2 + e.par(M1P_CGSOV);
  const double cgd = cgd1 + e.par(M1P_CGDOV) + (same_dir ? op_cgd : op_cgs);
  const double cgb = cgb1 + e.par(M1P_CGBOV) + op_cgb;
  Integ tr_gs = {0.0, 0.0, 0.0}, tr_gb = {0.0, 0.0, 0.0};
  const Integ tr_gd = {0.0, 0.0, 0.0};  // never assigned in the reference (mos.rs:801 writes tr.gs a second time)
  if (e.mode == AN_TRAN) {
    const double dqgs = same_dir ? (vgs - e.op(M1S_VGS)) * cgs : (vgs - e.op(M1S_VGD)) * cgs;
    tr_gs = integrate_be_e(e, e.dt, dqgs, cgs, vgs);
    const double dqgd = same_dir ? (vgd - e.op(M1S_VGD)) * cgd : (vgd - e.op(M1S_VGS)) * cgd;
    tr_gs = integrate_be_e(e, e.dt, dqgd, cgd, vgd);  // overwrites the gate-source result, as the reference does
    const double dqgb = (vgb - e.op(M1S_VGB)) * cgb;
    tr_gb = integrate_be_e(e, e.dt, dqgb, cgb, vgb);
    // tr.bs / tr.bd (mos.rs:808-827) are computed by the reference but never stamped nor read back: skipped.
  }
  const double irhs = ids - gm * vgs - gds * vds;
  const int sr = reversed ? M1_DP : M1_SP, sx = reversed ? M1_D : M1_S, dr = reversed ? M1_SP : M1_DP, dx = reversed ? M1_S : M1_D;
  const double grd = e.par(M1P_GRD), grs = e.par(M1P_GRS);
#define M1E(a, b) (M1_E0 + (a) * 6 + (b))
  e.add_g_at(M1E(dr, dr), gds + grd + gbd + tr_gd.g);
  e.add_g_at(M1E(sr, sr), gm + gds + grs + gbs + gmbs + tr_gs.g);
  e.add_g_at(M1E(dr, sr), -gm - gds - gmbs);
  e.add_g_at(M1E(sr, dr), -gds);
  e.add_g_at(M1E(dr, M1_G), gm - tr_gd.g);
  e.add_g_at(M1E(sr, M1_G), -gm - tr_gs.g);
  e.add_g_at(M1E(M1_G, M1_G), (tr_gd.g + tr_gs.g + tr_gb.g));
  e.add_g_at(M1E(M1_B, M1_B), (gbd + gbs + tr_gb.g));
  e.add_g_at(M1E(M1_G, M1_B), -tr_gb.g);
  e.add_g_at(M1E(M1_G, dr), -tr_gd.g);
  e.add_g_at(M1E(M1_G, sr), -tr_gs.g);
  e.add_g_at(M1E(M1_B, M1_G), -tr_gb.g);
  e.add_g_at(M1E(M1_B, dr), -gbd);
  e.add_g_at(M1E(M1_B, sr), -gbs);
  e.add_g_at(M1E(dr, M1_B), -gbd + gmbs);
  e.add_g_at(M1E(sr, M1_B), -gbs - gmbs);
  e.add_g_at(M1E(dx, dr), -grd);
  e.add_g_at(M1E(dr, dx), -grd);
  e.add_g_at(M1E(dx, dx), grd);
  e.add_g_at(M1E(sx, sr), -grs);
  e.add_g_at(M1E(sr, sx), -grs);
  e.add_g_at(M1E(sx, sx), grs);
  e.add_b_at(dr, p * (-irhs + ibd_rhs + tr_gd.rhs));
  e.add_b_at(sr, p * (irhs + ibs_rhs + tr_gs.rhs));
  e.add_b_at(M1_G, -p * (tr_gs.rhs + tr_gb.rhs + tr_gd.rhs));
  e.add_b_at(M1_B, -p * (ibd_rhs + ibs_rhs - tr_gb.rhs));
  e.set_guess(M1S_VGS, vgs); e.set_guess(M1S_VGD, vgd); e.set_guess(M1S_VGB, vgb); e.set_guess(M1S_VSB, vsb);
  e.set_guess(M1S_VDB, vdb); e.set_guess(M1S_CGS, cgs1); e.set_guess(M1S_CGD, cgd1); e.set_guess(M1S_CGB, cgb1);
  e.set_guess(M1S_REV, reversed ? 1.0 : 0.0);
  e.set_guess(M1S_GM, gm); e.set_guess(M1S_GDS, gds); e.set_guess(M1S_GMBS, gmbs); e.set_guess(M1S_GBS, gbs); e.set_guess(M1S_GBD, gbd);
}

// ================================================================ AC (complex) loads
// Resistor / Capacitor / Vsrc load_ac (comps/mod.rs:139-149, 223-238, 311-322); Mos1::load_ac (mos.rs:914-968).
template <class Env> __device__ __forceinline__ void load_ac_resistor(Env& e) {
  const double g = e.par(RP_G_TRAN);
  e.add_g_at(R_EPP, mk(g, 0.0));
  e.add_g_at(R_ENN, mk(g, 0.0));
  e.add_g_at(R_EPN, mk(-g, 0.0));
  e.add_g_at(R_ENP, mk(-g, 0.0));
}
template <class Env> __device__ __forceinline__ void load_ac_capacitor(Env& e) {
  const double c = e.par(CP_C);
  const double w = e.omega;
  e.add_g_at(R_EPP, mk(0.0, w * c));
  e.add_g_at(R_ENN, mk(0.0, w * c));
  e.add_g_at(R_EPN, mk(0.0, -w * c));
  e.add_g_at(R_ENP, mk(0.0, -w * c));
}
template <class Env> __device__ __forceinline__ void load_ac_vsrc(Env& e) {
  e.add_g_at(V_EPI, mk(1.0, 0.0));
  e.add_g_at(V_EIP, mk(1.0, 0.0));
  e.add_g_at(V_ENI, mk(-1.0, 0.0));
  e.add_g_at(V_EIN, mk(-1.0, 0.0));
  e.add_b_at(V_I, mk(e.par(VP_ACM), 0.0));
}
template <class Env> __device__ __forceinline__ void load_ac_mos1(Env& e) {
  const double omega = e.omega;
  const double gm = e.op(M1S_GM), gds = e.op(M1S_GDS), gmbs = e.op(M1S_GMBS), gbs = e.op(M1S_GBS), gbd = e.op(M1S_GBD);
  const double gcgs = omega * e.op(M1S_CGS);
  const double gcgd = omega * e.op(M1S_CGD);
  const double gcgb = omega * e.op(M1S_CGB);
  const bool reversed = e.op(M1S_REV) != 0.0;
  const int sr = reversed ? M1_DP : M1_SP, sx = reversed ? M1_D : M1_S, dr = reversed ? M1_SP : M1_DP, dx = reversed ? M1_S : M1_D;
  const double grd = e.par(M1P_GRD), grs = e.par(M1P_GRS);
  e.add_g_at(M1E(dr, dr), mk(gds + grd + gbd, gcgd));
  e.add_g_at(M1E(sr, sr), mk(gm + gds + grs + gbs + gmbs, gcgs));
  e.add_g_at(M1E(dr, sr), mk(-gm - gds - gmbs, 0.0));
  e.add_g_at(M1E(sr, dr), mk(-gds, 0.0));
  e.add_g_at(M1E(dr, M1_G), mk(gm, -gcgd));
  e.add_g_at(M1E(sr, M1_G), mk(-gm, -gcgs));
  e.add_g_at(M1E(M1_G, M1_G), mk(0.0, gcgd + gcgs + gcgb));
  e.add_g_at(M1E(M1_B, M1_B), mk(gbd + gbs, gcgb));
  e.add_g_at(M1E(M1_G, M1_B), mk(0.0, -gcgb));
  e.add_g_at(M1E(M1_G, dr), mk(0.0, -gcgd));
  e.add_g_at(M1E(M1_G, sr), mk(0.0, -gcgs));
  e.add_g_at(M1E(M1_B, M1_G), mk(0.0, -gcgb));
  e.add_g_dup(M1E(M1_G, dr), M1_DUP_GDR, mk(0.0, -gcgd));  // pushed twice by the reference (mos.rs:951 and :954)
  e.add_g_at(M1E(M1_B, dr), mk(-gbd, 0.0));
  e.add_g_at(M1E(M1_B, sr), mk(-gbs, 0.0));
  e.add_g_at(M1E(dr, M1_B), mk(-gbd + gmbs, 0.0));
  e.add_g_at(M1E(sr, M1_B), mk(-gbs - gmbs, 0.0));
  e.add_g_at(M1E(dx, dr), mk(-grd, 0.0));
  e.add_g_at(M1E(dr, dx), mk(-grd, 0.0));
  e.add_g_at(M1E(dx, dx), mk(grd, 0.0));
  e.add_g_at(M1E(sx, sr), mk(-grs, 0.0));
  e.add_g_at(M1E(sr, sx), mk(-grs, 0.0));
  e.add_g_at(M1E(sx, sx), mk(grs, 0.0));
}
#undef M1E

}  // namespace s21

// ---------------------------------------------------------------- Bsim4 (bsim4/bsim4_eval.hpp; the reference has no load_ac for it)
#ifndef S21_JIT  // the run-time specialised kernel (host/jit.hpp) never meets a Bsim4 device
#include "../bsim4/bsim4_eval.hpp"
namespace s21 {
// Not inlined: the model is ~3000 lines of straight-line arithmetic, shared by every kernel variant that can meet one.
template <class Env> __device__ __noinline__ void load_bsim4(Env& e) { b4e::load_bsim4(e); }
}  // namespace s21
#endif  // S21_JIT
