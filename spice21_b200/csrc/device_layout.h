// Structure-of-arrays layouts shared by the host table builders and the CUDA kernels.
//
// Every device instance owns, in three flat tables that are shared by the whole batch:
//   itab  (int32): its variable indices (-1 = ground) followed by its matrix-element handles (-1 = dropped),
//   ptab  (int32): one "param code" per parameter: (offset << 1) | per_instance into the value pool `pval`,
//                  value = pval[offset + per_instance * instance_id]  (shared params broadcast, MC columns coalesce),
//   state slots  : committed (`op`) and in-flight (`guess`) device state, one column per instance.
// Device type codes follow the reference's `enum ComponentSolver` (spice21/src/comps/mod.rs:60-72).
#pragma once

namespace s21 {

enum DevType { DT_R = 0, DT_C = 1, DT_I = 2, DT_V = 3, DT_DIODE = 4, DT_MOS0 = 5, DT_MOS1 = 6, DT_BSIM4 = 7, DT_NTYPES = 8 };

enum AnMode { AN_OP = 0, AN_TRAN = 1, AN_AC = 2 };

// ---- Resistor (comps/mod.rs:271-323): itab [p, n, e_pp, e_pn, e_np, e_nn]; params: g in OP, g in TRAN/AC.
// The two values differ only for the initial-condition resistor of Tran::ic (analysis.rs:517, 547-549).
enum { R_P = 0, R_N, R_EPP, R_EPN, R_ENP, R_ENN, R_NI };
enum { RP_G_OP = 0, RP_G_TRAN, RP_N };
// ---- Capacitor (comps/mod.rs:152-239): same itab; param c; state q.
enum { CP_C = 0, CP_N };
enum { CS_Q = 0, CS_N };
// ---- Isrc (comps/mod.rs:325-346): itab [p, n]; param i.
enum { I_P = 0, I_N, I_NI };
enum { IP_I = 0, IP_N };
// ---- Vsrc (comps/mod.rs:95-150): itab [p, n, i, e_pi, e_ip, e_ni, e_in]; params v (OP), v (TRAN), acm.
enum { V_P = 0, V_N, V_I, V_EPI, V_EIP, V_ENI, V_EIN, V_NI };
// VP_WKIND / VP_W0..6: time-varying source (an extension — the reference's Vsrc is DC / acm only,
// spice21.proto:29-35). kind 0 = none (the transient value is VP_V_TRAN, as in the reference), 1 = PULSE(v1 v2 td tr tf pw per),
// 2 = SIN(vo va freq td theta); SPICE's definitions. The OP always uses VP_V_OP (the `dc` field).
enum { VP_V_OP = 0, VP_V_TRAN, VP_ACM, VP_WKIND, VP_W0, VP_W1, VP_W2, VP_W3, VP_W4, VP_W5, VP_W6, VP_N };
enum { SRC_NONE = 0, SRC_PULSE = 1, SRC_SIN = 2 };
// ---- Diode (comps/diode.rs:217-355): itab [p, n, r, e_pp, e_pr, e_rp, e_rr, e_nr, e_rn, e_nn].
enum { D_P = 0, D_N, D_R, D_EPP, D_EPR, D_ERP, D_ERR, D_ENR, D_ERN, D_ENN, D_NI };
enum { DP_VTE = 0, DP_VCRIT, DP_ISAT, DP_GSPR, DP_CZ, DP_CZ2, DP_DEPTH, DP_F1, DP_F3, DP_BV, DP_HASBV, DP_TT, DP_VJ, DP_M, DP_N };
enum { DS_VD = 0, DS_CHARGE, DS_N };
// ---- Mos0 (comps/mos.rs:1026-1099): itab [d, g, s, e_dd, e_ss, e_ds, e_sd, e_dg, e_sg]; param p (+1/-1).
enum { M0_D = 0, M0_G, M0_S, M0_EDD, M0_ESS, M0_EDS, M0_ESD, M0_EDG, M0_ESG, M0_NI };
enum { M0P_P = 0, M0P_N };
// ---- Mos1 (comps/mos.rs:625-969): itab [D,G,S,B,DP,SP, e[6][6]] with e indexed by the reference's Mos1Var values.
enum { M1_D = 0, M1_G = 1, M1_S = 2, M1_B = 3, M1_DP = 4, M1_SP = 5, M1_E0 = 6, M1_NI = 6 + 36,
       M1_DUP_GDR = M1_NI,  // staging slot of load_ac's second (G,dr) push (mos.rs:954)
       M1_NSTAGE = M1_NI + 1 };
enum {
  M1P_P = 0, M1P_VT0T, M1P_PHIT, M1P_GAMMA, M1P_BETA, M1P_LAMBDA, M1P_VTHERM, M1P_COX, M1P_CGSOV, M1P_CGDOV, M1P_CGBOV, M1P_GRD,
  M1P_GRS, M1P_MJ, M1P_MJSW,
  M1P_SJ = 15,  // source junction block, then drain junction block (8 each)
  M1P_DJ = 23,
  M1P_N = 31
};
enum { MJ_ISAT = 0, MJ_CZB, MJ_CZBSW, MJ_BULKPOT, MJ_DEPTH, MJ_F2, MJ_F3, MJ_F4, MJ_N };
// Committed/in-flight Mos1 state: only what op_stamp reads back (mos.rs:757-826) plus what load_ac reads (:924-930).
enum { M1S_VGS = 0, M1S_VGD, M1S_VGB, M1S_VSB, M1S_VDB, M1S_CGS, M1S_CGD, M1S_CGB, M1S_REV, M1S_GM, M1S_GDS, M1S_GMBS, M1S_GBS, M1S_GBD, M1S_N };

// Per-type table sizes: ints, params, state slots, stamp pushes (G + b) per load.
struct DevShape { int n_itab, n_par, n_state, n_stamps; };

}  // namespace s21
