// BSIM4 device evaluation: one Newton-iteration "load" of a BSIM4 MOSFET, written once for both
// the CUDA kernels (through kernels/devices.cuh) and host C++.
//
// What it computes is the Berkeley BSIM4.8 model as the reference restates it in
//   spice21/src/comps/bsim4/bsim4solver.rs:145-3754 (`op`), :3771-3873 (limiters, poly depletion),
//   spice21/src/comps/bsim4/tran.rs:11-549 (`tran_op`) and spice21/src/comps/bsim4/stamp.rs:8-572 (`stamp`).
// The model equations are evaluated expression-for-expression in the reference's operation order (no FMA contraction:
// nvcc -fmad=false, g++ -ffp-contract=off) so results agree to the last bits of libm. Reference deviations from the
// Berkeley code are kept because they change results, e.g. the bias-dependent Rds is only computed for rdsmod > 1,
// i.e. never (bsim4solver.rs:1082), several capacitance terms are not carried into the stamp (tran.rs:497-549) and the
// drain/source junction use the un-limited sbNode/dbNode voltages when rbodymod == 0 (bsim4solver.rs:531-532).
//
// Organisation (ours): the evaluation is a pipeline of phases over two plain structs — `B4Bias` (limited terminal
// voltages) and `B4Op` (everything the stamp needs) — with the parameter block read through the Env:
//   b4_limit_bias -> b4_junction_dc -> b4_channel_dc -> b4_leakage -> b4_charge -> b4_tran_caps -> b4_stamp
// Env concept: kernels/devices.cuh. Parameters come from e.par(B4F_*) (bsim4_layout.h).
#pragma once
#include <math.h>

#include "../device_layout.h"
#include "bsim4_layout.h"

#if defined(__CUDACC__)
#define B4_HD __host__ __device__ __forceinline__
#else
#define B4_HD inline
#endif

namespace s21 {
namespace b4e {

// Every f64 division of the evaluation goes through this one macro (scripts/b4_route_divisions.py rewrote the ~450
// `X / Y` sites, keeping C++'s grouping: B4_DIV(whole multiplicative chain to the left, next unary expression)). The
// default is the plain quotient. Two device-side variants were measured on C4 (2048 instances x 100 points, B200):
//   * through the split exact division of csrc/scalar.h, inlined: 252 ms against 219 ms — slower, removed
//     (profiles/r02a_c4_sdiv.txt);
//   * -DS21_B4_NIDIV: ONE out-of-line copy of the compiler's division per kernel (a call per site): the evaluation's
//     text shrinks by the ~460 inlined fast-path/slow-path sequences (ncu: 20 % of the stall samples of the C4 kernel
//     are instruction-cache misses, its text is 39.6 k instructions);
//   * -DS21_B4_NIMATH: the same for exp / log / sqrt.
#if defined(__CUDACC__) && defined(S21_B4_NIDIV)
static __device__ __noinline__ double b4_ddiv_ool(double a, double b) { return __ddiv_rn(a, b); }  // IEEE quotient
#endif
//   * -DS21_B4_RCPDIV (kernels/coop_fast.cu, opt-in at run time with S21_B4_FAST=1): a / b as a * rcp(b), rcp = the hardware
//     seed (MUFU.RCP64H) plus one cubic Newton step (3 DFMA) — 8 instructions and no control flow against the 16 + an
//     out-of-line slow path of the IEEE quotient — and, because the reciprocal is a pure function of b, the compiler shares
//     it between all divisions by one denominator (806 MUFU in the kernel text against 1258). The quotient is within
//     ~1.5 ulp of a / b instead of correctly rounded: inside the 1e-9 (dcop) / reltol (tran) parity bounds, but no longer
//     the reference's bits, so it is not the default; the host build and the oracle keep the IEEE quotient. Counted on
//     the host (scripts/b4_opcount.py): 235 divisions are executed per evaluation, none with an operand or result outside
//     the normal range on the exact trajectory — a Newton iterate that overshoots can still produce an infinite or zero
//     denominator (exp overflow), where the Newton step gives NaN: then the seed itself (0 for +-inf, +-inf for +-0) is
//     the IEEE answer and is selected without a branch.
#if defined(__CUDACC__) && defined(S21_B4_RCPDIV)
__device__ __forceinline__ double b4_rcp(double b) {
  double r0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(b));  // not volatile: identical reciprocals are merged
  double e = __fma_rn(-b, r0, 1.0);
  const bool fin = e == e;  // NaN <=> b is +-0 / subnormal (seed +-inf), +-inf (seed 0) or NaN
  e = __fma_rn(e, e, e);
  const double r = __fma_rn(r0, e, r0);
  return fin ? r : r0;
}
#endif
#if defined(__CUDA_ARCH__) && defined(S21_B4_NIDIV)
#define B4_DIV(a, b) ::s21::b4e::b4_ddiv_ool((double)(a), (double)(b))
#elif defined(__CUDA_ARCH__) && defined(S21_B4_RCPDIV)
#define B4_DIV(a, b) ((double)(a) * ::s21::b4e::b4_rcp((double)(b)))
#elif !defined(B4_DIV)  // a host build may bring its own (an instrumented test build counts the executed divisions)
#define B4_DIV(a, b) ((a) / (b))
#endif
#if defined(__CUDACC__) && defined(S21_B4_NIMATH)
// unqualified exp / log / sqrt inside this namespace resolve to these (inner scope hides ::exp): one copy per kernel
static __host__ __device__ __noinline__ double exp(double x) { return ::exp(x); }
static __host__ __device__ __noinline__ double log(double x) { return ::log(x); }
static __host__ __device__ __noinline__ double sqrt(double x) { return ::sqrt(x); }
#endif

// bsim4/mod.rs:35-63 and comps/consts
#define B4C_EXPL_THRESHOLD 100.0
#define B4C_EXP_THRESHOLD 34.0
#define B4C_MAX_EXP 5.834617425e14
#define B4C_MIN_EXP 1.713908431e-15
#define B4C_MAX_EXPL 2.688117142e+43
#define B4C_MIN_EXPL 3.720075976e-44
#define B4C_DELTA_1 0.02
#define B4C_DELTA_2 0.02
#define B4C_DELTA_3 0.02
#define B4C_DELTA_4 0.02
#define B4C_Q 1.6021918e-19
#define B4C_KB 1.3806226e-23
#define B4C_KB_OVER_Q (B4C_KB / B4C_Q)
#define B4C_VT_REF (B4C_KB * (273.15 + 27.0) / B4C_Q)
#define B4C_EPS0 8.85418e-12
#define B4C_EPSSI 1.03594e-10
#define B4C_PI 3.14159265358979323846264338327950288

B4_HD double b4_dexpb(double a) {
  if (a > B4C_EXP_THRESHOLD) return B4C_MAX_EXP * (1.0 + a - B4C_EXP_THRESHOLD);
  if (a < -B4C_EXP_THRESHOLD) return B4C_MIN_EXP;
  return exp(a);
}
B4_HD double b4_dexpc(double a) {
  if (a > B4C_EXP_THRESHOLD) return B4C_MAX_EXP;
  if (a < -B4C_EXP_THRESHOLD) return 0.0;
  return exp(a);
}
B4_HD double b4_max(double a, double b) { return a > b ? a : b; }  // cmath MAX / f64::max on non-NaN operands
B4_HD double b4_min(double a, double b) { return a < b ? a : b; }

// ---- inter-iteration limiters (bsim4solver.rs:3789-3873)
B4_HD double b4_limvds(double vnew, double vold) {
  if (vold >= 3.5) {
    if (vnew > vold) return b4_min(vnew, 3.0 * vold + 2.0);
    if (vnew < 3.5) return b4_max(vnew, 2.0);
    return vnew;
  }
  return vnew > vold ? b4_min(vnew, 4.0) : b4_max(vnew, -0.5);
}
B4_HD double b4_fetlim(double vnew, double vold, double vto) {
  const double vtsthi = fabs(2.0 * (vold - vto)) + 2.0;
  const double vtstlo = B4_DIV(vtsthi, 2.0) + 2.0;
  const double vtox = vto + 3.5;
  const double delv = vnew - vold;
  if (vold >= vto) {
    if (vold >= vtox) {                       // well on
      if (delv <= 0.0) {
        if (vnew >= vtox) { if (-delv > vtstlo) return vold - vtstlo; }
        else return b4_max(vnew, vto + 2.0);
      } else if (delv >= vtsthi) {
        return vold + vtsthi;
      }
    } else {                                  // around threshold
      return delv <= 0.0 ? b4_max(vnew, vto - 0.5) : b4_min(vnew, vto + 4.0);
    }
  } else {                                    // off
    if (delv <= 0.0) { if (-delv > vtsthi) return vold - vtsthi; }
    else {
      const double vtemp = vto + 0.5;
      if (vnew <= vtemp) { if (delv > vtstlo) return vold + vtstlo; }
      else return vtemp;
    }
  }
  return vnew;
}
B4_HD double b4_pnjlim(double vnew, double vold, double vt, double vcrit) {
  if (vnew > vcrit && fabs(vnew - vold) > (vt + vt)) {
    if (vold > 0.0) {
      const double arg = 1.0 + B4_DIV((vnew - vold), vt);
      return arg > 0.0 ? vold + vt * log(arg) : vcrit;
    }
    return vt * log(B4_DIV(vnew, vt));
  }
  return vnew;
}
// poly-gate depletion (bsim4solver.rs:3771-3787): effective gate voltage and its derivative
B4_HD void b4_poly_depletion(double phi, double ngate, double epsgate, double coxe, double Vgs, double* Vgs_eff, double* dVgs_eff_dVg) {
  if (ngate > 1.0e18 && ngate < 1.0e25 && Vgs > phi && epsgate != 0.0) {
    const double T1 = B4_DIV(1.0e6 * B4C_Q * epsgate * ngate, (coxe * coxe));
    const double T8 = Vgs - phi;
    const double T4 = sqrt(1.0 + B4_DIV(2.0 * T8, T1));
    const double T2 = B4_DIV(2.0 * T8, (T4 + 1.0));
    const double T3 = B4_DIV(0.5 * T2 * T2, T1);
    const double T7 = 1.12 - T3 - 0.05;
    const double T6 = sqrt(T7 * T7 + 0.224);
    const double T5 = 1.12 - 0.5 * (T7 + T6);
    *Vgs_eff = Vgs - T5;
    *dVgs_eff_dVg = 1.0 - (0.5 - B4_DIV(0.5, T4)) * (1.0 + B4_DIV(T7, T6));
  } else {
    *Vgs_eff = Vgs;
    *dVgs_eff_dVg = 1.0;
  }
}

#define M_(f) e.par(B4F_m_##f)
#define D_(f) e.par(B4F_d_##f)
#define S_(f) e.par(B4F_s_##f)
#define I_(f) e.par(B4F_i_##f)

// Limited terminal voltages of one evaluation, in the device's own polarity.
struct B4Bias {
  double vds, vgs, vbs, vges, vgms, vdbs, vsbs, vses, vdes, qdef;
  double vbd, vgd, vgb, vged, vgmd, vgmb, vdbd, vbs_jct, vbd_jct;
};

// Everything the stamp and the state update read (the part of Bsim4OpPoint that is live, bsim4/mod.rs:217-438).
struct B4Op {
  int mode;
  double von, gbs, cbs, gbd, cbd;
  double cd, gds, gm, gmbs, csub, gbds, gbgs, gbbs;
  double Igidl, ggidld, ggidlg, ggidlb, Igisl, ggisls, ggislg, ggislb;
  double Igs, gIgsg, gIgss, Igd, gIgdg, gIgdd, Igcs, gIgcsg, gIgcsd, gIgcsb, Igcd, gIgcdg, gIgcdd, gIgcdb, gIgcss, gIgcds;
  double Igb, gIgbg, gIgbd, gIgbb, gIgbs;
  double gcrg, gcrgd, gcrgg, gcrgs, gcrgb;
  double gstot, gstotd, gstotg, gstots, gstotb, gdtot, gdtotd, gdtotg, gdtots, gdtotb;
  double qgate, qbulk, qdrn, qsrc, qgmid, qbs, qbd, qchqs, qcheq, qdef_dump;
  double cggb, cgdb, cgsb, cbgb, cbdb, cbsb, cdgb, cddb, cdsb, capbd, capbs, cgdo, cgso, qgdo, qgso;
  double cqgb, cqdb, cqsb, cqbb, gtau;
  double qg, qd, qb;
};

// ------------------------------------------------------------------------------------------------------------------
// Phase 1: terminal voltages, limited against the previous iteration (bsim4solver.rs:415-532)
template <class E> B4_HD void b4_limit_bias(E& e, B4Bias& v) {
  const double tp = M_(type_sign);
  const int rgatemod = (int)M_(rgatemod), rdsmod = (int)M_(rdsmod), rbodymod = (int)M_(rbodymod);
  const double vsp = e.volt(e.node(B4N_SP));
  double vds = tp * (e.volt(e.node(B4N_DP)) - vsp);
  double vgs = tp * (e.volt(e.node(B4N_GP)) - vsp);
  double vbs = tp * (e.volt(e.node(B4N_BP)) - vsp);
  double vges = tp * (e.volt(e.node(B4N_GE)) - vsp);
  double vgms = tp * (e.volt(e.node(B4N_GM)) - vsp);
  double vdbs = tp * (e.volt(e.node(B4N_DB)) - vsp);
  double vsbs = tp * (e.volt(e.node(B4N_SB)) - vsp);
  double vses = tp * (e.volt(e.node(B4N_S)) - vsp);
  double vdes = tp * (e.volt(e.node(B4N_D)) - vsp);
  v.qdef = tp * e.volt(e.node(B4N_Q));

  const double o_vgs = e.guess(B4S_VGS), o_vds = e.guess(B4S_VDS), o_vbs = e.guess(B4S_VBS), o_vges = e.guess(B4S_VGES);
  const double o_vgms = e.guess(B4S_VGMS), o_vdbs = e.guess(B4S_VDBS), o_vsbs = e.guess(B4S_VSBS), o_vses = e.guess(B4S_VSES);
  const double o_vdes = e.guess(B4S_VDES), o_vbd = e.guess(B4S_VBD), o_vdbd = e.guess(B4S_VDBD), von = e.guess(B4S_VON);
  const double vgdo = o_vgs - o_vds, vgedo = o_vges - o_vds, vgmdo = o_vgms - o_vds;

  double vbd = vbs - vds, vdbd = vdbs - vds, vgd = vgs - vds, vged = vges - vds, vgmd = vgms - vds;

  if (o_vds >= 0.0) {
    vgs = b4_fetlim(vgs, o_vgs, von);
    vds = vgs - vgd;
    vds = b4_limvds(vds, o_vds);
    vgd = vgs - vds;
    if (rgatemod == 3) {
      vges = b4_fetlim(vges, o_vges, von);
      vgms = b4_fetlim(vgms, o_vgms, von);
      vged = vges - vds;
      vgmd = vgms - vds;
    } else if (rgatemod == 1 || rgatemod == 2) {
      vges = b4_fetlim(vges, o_vges, von);
      vged = vges - vds;
    }
    if (rdsmod != 0) {
      vdes = b4_limvds(vdes, o_vdes);
      vses = -b4_limvds(-vses, -o_vses);
    }
  } else {
    vgd = b4_fetlim(vgd, vgdo, von);
    vds = vgs - vgd;
    vds = -b4_limvds(-vds, -o_vds);
    vgs = vgd + vds;
    if (rgatemod == 3) {
      vged = b4_fetlim(vged, vgedo, von);
      vges = vged + vds;
      vgmd = b4_fetlim(vgmd, vgmdo, von);
      vgms = vgmd + vds;
    }
    if (rgatemod == 1 || rgatemod == 2) {
      vged = b4_fetlim(vged, vgedo, von);
      vges = vged + vds;
    }
    if (rdsmod != 0) {
      vdes = -b4_limvds(-vdes, -o_vdes);
      vses = b4_limvds(vses, o_vses);
    }
  }

  const double vcrit = D_(vcrit);
  if (vds >= 0.0) {
    vbs = b4_pnjlim(vbs, o_vbs, B4C_VT_REF, vcrit);
    vbd = vbs - vds;
    if (rbodymod != 0) {
      vdbs = b4_pnjlim(vdbs, o_vdbs, B4C_VT_REF, vcrit);
      vdbd = vdbs - vds;
      vsbs = b4_pnjlim(vsbs, o_vsbs, B4C_VT_REF, vcrit);
    }
  } else {
    vbd = b4_pnjlim(vbd, o_vbd, B4C_VT_REF, vcrit);
    vbs = vbd + vds;
    if (rbodymod != 0) {
      vdbd = b4_pnjlim(vdbd, o_vdbd, B4C_VT_REF, vcrit);
      vdbs = vdbd + vds;
      const double vsbdo = o_vsbs - o_vds;
      const double vsbd = b4_pnjlim(vsbs - vds, vsbdo, B4C_VT_REF, vcrit);
      vsbs = vsbd + vds;
    }
  }
  (void)vged; (void)vgmd; (void)vdbd;

  v.vds = vds; v.vgs = vgs; v.vbs = vbs; v.vges = vges; v.vgms = vgms; v.vdbs = vdbs; v.vsbs = vsbs; v.vses = vses; v.vdes = vdes;
  v.vbd = vbs - vds;
  v.vgd = vgs - vds;
  v.vgb = vgs - vbs;
  v.vged = vges - vds;
  v.vgmd = vgms - vds;
  v.vgmb = vgms - vbs;
  v.vdbd = vdbs - vds;
  v.vbs_jct = rbodymod != 0 ? vbs : vsbs;
  v.vbd_jct = rbodymod != 0 ? v.vbd : v.vdbd;
}

// ------------------------------------------------------------------------------------------------------------------
// Phase 2: source/drain junction diode DC current and conductance, plus trap-assisted tunnelling (bsim4solver.rs:534-775)
struct B4JunctionSide {  // parameters of one junction, gathered by the caller
  double Isat, Nvtm, xjbv, bv, XExpBV, vjmFwd, vjmRev, IVjmFwd, IVjmRev, slpFwd, slpRev;
};
B4_HD void b4_junction_iv(int diomod, const B4JunctionSide& j, double vj, double gmin, double* g, double* c) {
  if (j.Isat <= 0.0) {
    *g = gmin;
    *c = *g * vj;
    return;
  }
  if (diomod == 0) {
    const double ev = exp(B4_DIV(vj, j.Nvtm));
    const double T1 = j.xjbv * exp(B4_DIV(-(j.bv + vj), j.Nvtm));
    *g = B4_DIV(j.Isat * (ev + T1), j.Nvtm) + gmin;
    *c = j.Isat * (ev + j.XExpBV - T1 - 1.0) + gmin * vj;
  } else if (diomod == 1) {
    const double T2 = B4_DIV(vj, j.Nvtm);
    if (T2 < -B4C_EXP_THRESHOLD) {
      *g = gmin;
      *c = j.Isat * (B4C_MIN_EXP - 1.0) + gmin * vj;
    } else if (vj <= j.vjmFwd) {
      const double ev = exp(T2);
      *g = B4_DIV(j.Isat * ev, j.Nvtm) + gmin;
      *c = j.Isat * (ev - 1.0) + gmin * vj;
    } else {
      const double T0 = B4_DIV(j.IVjmFwd, j.Nvtm);
      *g = T0 + gmin;
      *c = j.IVjmFwd - j.Isat + T0 * (vj - j.vjmFwd) + gmin * vj;
    }
  } else {  // diomod 2: resistive both in forward and in breakdown
    if (!(vj < j.vjmRev) && !(vj <= j.vjmFwd)) {
      *g = j.slpFwd + gmin;
      *c = j.IVjmFwd + j.slpFwd * (vj - j.vjmFwd) + gmin * vj;
      return;
    }
    const double T0 = B4_DIV(vj, j.Nvtm);
    double ev, dev;
    if (T0 < -B4C_EXP_THRESHOLD) { ev = B4C_MIN_EXP; dev = 0.0; }
    else { ev = exp(T0); dev = B4_DIV(ev, j.Nvtm); }
    if (vj < j.vjmRev) {
      const double T1 = ev - 1.0;
      const double T2 = j.IVjmRev + j.slpRev * (vj - j.vjmRev);
      *g = dev * T2 + T1 * j.slpRev + gmin;
      *c = T1 * T2 + gmin * vj;
    } else {
      const double T1 = B4_DIV((j.bv + vj), j.Nvtm);
      double T2, T3;
      if (T1 > B4C_EXP_THRESHOLD) { T2 = B4C_MIN_EXP; T3 = 0.0; }
      else { T2 = exp(-T1); T3 = B4_DIV(-T2, j.Nvtm); }
      *g = j.Isat * (dev - j.xjbv * T3) + gmin;
      *c = j.Isat * (ev + j.XExpBV - 1.0 - j.xjbv * T2) + gmin * vj;
    }
  }
}
// one trap-assisted-tunnelling term: value and d/dVb of the bias factor
B4_HD void b4_tat_factor(double vts, double Nvtmr, double vj, double* f, double* df) {
  if ((vts - vj) < (vts * 1e-3)) {
    const double T9 = 1.0e3;
    const double T0 = B4_DIV(-vj, Nvtmr) * T9;
    *f = b4_dexpb(T0);
    *df = B4_DIV(b4_dexpc(T0), Nvtmr) * T9;
  } else {
    const double T9 = B4_DIV(1.0, (vts - vj));
    const double T0 = B4_DIV(-vj, Nvtmr) * vts * T9;
    const double dT0 = B4_DIV(vts, Nvtmr) * (T9 + vj * T9 * T9);
    *f = b4_dexpb(T0);
    *df = b4_dexpc(T0) * dT0;
  }
}
template <class E> B4_HD void b4_junction_dc(E& e, const B4Bias& v, B4Op& o) {
  const int diomod = (int)M_(diomod);
  const double gmin = e.gmin;
  B4JunctionSide s{I_(SourceSatCurrent), D_(Nvtms), M_(xjbvs), M_(bvs), I_(XExpBVS), I_(vjsmFwd), I_(vjsmRev), I_(IVjsmFwd), I_(IVjsmRev), I_(SslpFwd), I_(SslpRev)};
  b4_junction_iv(diomod, s, v.vbs_jct, gmin, &o.gbs, &o.cbs);
  B4JunctionSide d{I_(DrainSatCurrent), D_(Nvtmd), M_(xjbvd), M_(bvd), I_(XExpBVD), I_(vjdmFwd), I_(vjdmRev), I_(IVjdmFwd), I_(IVjdmRev), I_(DslpFwd), I_(DslpRev)};
  b4_junction_iv(diomod, d, v.vbd_jct, gmin, &o.gbd, &o.cbd);

  double T1, T2, T3, T4, T5, T6, dT1, dT2, dT3, dT4, dT5, dT6;
  b4_tat_factor(M_(vtss), D_(Nvtmrss), v.vbs_jct, &T1, &dT1);
  b4_tat_factor(M_(vtsd), D_(Nvtmrsd), v.vbd_jct, &T2, &dT2);
  b4_tat_factor(M_(vtssws), D_(Nvtmrssws), v.vbs_jct, &T3, &dT3);
  b4_tat_factor(M_(vtsswd), D_(Nvtmrsswd), v.vbd_jct, &T4, &dT4);
  b4_tat_factor(M_(vtsswgs), D_(Nvtmrsswgs), v.vbs_jct, &T5, &dT5);
  b4_tat_factor(M_(vtsswgd), D_(Nvtmrsswgd), v.vbd_jct, &T6, &dT6);
  o.gbs += I_(SjctTempRevSatCur) * dT1 + I_(SswTempRevSatCur) * dT3 + I_(SswgTempRevSatCur) * dT5;
  o.cbs -= I_(SjctTempRevSatCur) * (T1 - 1.0) + I_(SswTempRevSatCur) * (T3 - 1.0) + I_(SswgTempRevSatCur) * (T5 - 1.0);
  o.gbd += I_(DjctTempRevSatCur) * dT2 + I_(DswTempRevSatCur) * dT4 + I_(DswgTempRevSatCur) * dT6;
  o.cbd -= I_(DjctTempRevSatCur) * (T2 - 1.0) + I_(DswTempRevSatCur) * (T4 - 1.0) + I_(DswgTempRevSatCur) * (T6 - 1.0);
}

}  // namespace b4e
}  // namespace s21

#include "bsim4_eval_channel.hpp"
