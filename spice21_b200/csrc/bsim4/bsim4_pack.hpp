// Host: flatten (model card, model-derived, size-dependent, per-instance) into the parameter block the device
// evaluation indexes with B4F_* (bsim4_layout.h), and describe a device's flavour (which internal nodes / stamp slots
// exist: bsim4ports.rs:24-112, bsim4solver.rs:28-115).
#pragma once
#include <vector>

#include "bsim4_layout.h"
#include "bsim4_size.hpp"

namespace s21 {
namespace b4 {

inline void pack_params(const Model& m, const ModelDerived& d, const SizeDep& s, const Internal& i, double* out) {
  // model fields: two synthetic entries, the rest straight from the card
  out[B4F_m_type_sign] = m.p();
  out[B4F_m_vtl_given] = m.has("vtl") ? 1.0 : 0.0;
#define X(n) out[B4F_m_##n] = m.n;
  B4_MODEL_CARD_FIELDS(X)
#undef X
#define X(n) out[B4F_d_##n] = d.n;
  B4_DERIVED_EVAL_FIELDS(X)
#undef X
#define B4S(n) out[B4F_s_##n] = s.n;
#define B4I(n) out[B4F_i_##n] = i.n;
#include "bsim4_fields.inc"
#undef B4S
#undef B4I
}

// Instance card from a name -> value bag (bsim4.proto:6-43 field names; unknown keys are an error).
inline InstSpec inst_from_map(const std::map<std::string, double>& kv) {
  InstSpec in;
  for (auto& e : kv) {
    const std::string& k = e.first;
    const double x = e.second;
    if (k == "l") in.l = x; else if (k == "w") in.w = x; else if (k == "nf") in.nf = x; else if (k == "sa") in.sa = x;
    else if (k == "sb") in.sb = x; else if (k == "sd") in.sd = x; else if (k == "sca") in.sca = x; else if (k == "scb") in.scb = x;
    else if (k == "scc") in.scc = x; else if (k == "sc") in.sc = x; else if (k == "ad") in.ad = x; else if (k == "as") in.as = x;
    else if (k == "pd") in.pd = x; else if (k == "ps") in.ps = x; else if (k == "nrd") in.nrd = x; else if (k == "nrs") in.nrs = x;
    else if (k == "delvto") in.delvto = x; else if (k == "min") in.min = x; else if (k == "rgeomod") in.rgeomod = x;
    else if (k == "rbdb") in.rbdb = x; else if (k == "rbsb") in.rbsb = x; else if (k == "rbpb") in.rbpb = x; else if (k == "rbps") in.rbps = x;
    else if (k == "rbpd") in.rbpd = x; else if (k == "xgw") in.xgw = x; else if (k == "ngcon") in.ngcon = x;
    else if (k == "trnqsmod" || k == "acnqsmod" || k == "rbodymod" || k == "rgatemod" || k == "geomod") {
      // accepted on the wire, ignored by the reference: the model card's selectors win (bsim4inst.rs:134-146)
    } else throw ModelError("unknown Bsim4 instance parameter: " + k);
  }
  return in;
}

// Which optional parts of the device exist.
struct Flavor {
  int rgatemod = 0, rdsmod = 0, rbodymod = 0, trnqsmod = 0;
  bool drain_source_prime = false;  // separate dNodePrime / sNodePrime (rdsmod != 0 or tnoimod == 1)
};
inline Flavor flavor_of(const Model& m, const Internal& i) {
  Flavor f;
  f.rgatemod = (int)i.rgatemod; f.rdsmod = (int)m.rdsmod; f.rbodymod = (int)i.rbodymod; f.trnqsmod = (int)i.trnqsmod;
  f.drain_source_prime = m.rdsmod != 0 || m.tnoimod == 1;
  return f;
}

// One (model card, instance card) pair reduced to what a device needs: the parameter block and the flavour.
struct Derived { std::vector<double> par; Flavor flavor; };
inline Derived derive_device(int mos_type, const std::map<std::string, double>& model_kv, const std::map<std::string, double>& inst_kv) {
  const Model m = resolve_model(mos_type, model_kv);
  const ModelDerived d = derive_model(m);
  const SizeAndInstance si = size_and_instance(m, d, inst_from_map(inst_kv));
  Derived r;
  r.par.assign(B4F_COUNT, 0.0);
  pack_params(m, d, si.s, si.i, r.par.data());
  r.flavor = flavor_of(m, si.i);
  return r;
}

// Matrix elements in the reference's creation order (bsim4solver.rs:28-115): (row node, col node) as B4Node positions.
// Elements of absent blocks are skipped, exactly as the reference skips their creation.
struct ElemSpec { int slot_key, row, col; };
enum MatP {
  MP_DPbp, MP_GPbp, MP_SPbp, MP_BPdp, MP_BPgp, MP_BPsp, MP_BPbp, MP_Dd, MP_GPgp, MP_Ss, MP_DPdp, MP_SPsp, MP_Ddp, MP_GPdp, MP_GPsp, MP_Ssp,
  MP_DPsp, MP_DPd, MP_DPgp, MP_SPgp, MP_SPs, MP_SPdp, MP_Qq, MP_Qbp, MP_Qdp, MP_Qsp, MP_Qgp, MP_DPq, MP_SPq, MP_GPq,
  MP_GEge, MP_GEgp, MP_GPge, MP_GEdp, MP_GEsp, MP_GEbp, MP_GMdp, MP_GMgp, MP_GMgm, MP_GMge, MP_GMsp, MP_GMbp, MP_DPgm, MP_GPgm, MP_GEgm,
  MP_SPgm, MP_BPgm,
  MP_DPdb, MP_SPsb, MP_DBdp, MP_DBdb, MP_DBbp, MP_DBb, MP_BPdb, MP_BPb, MP_BPsb, MP_SBsp, MP_SBbp, MP_SBb, MP_SBsb, MP_Bdb, MP_Bbp, MP_Bsb, MP_Bb,
  MP_Dgp, MP_Dsp, MP_Dbp, MP_Sdp, MP_Sgp, MP_Sbp, MP_COUNT
};
inline std::vector<ElemSpec> matrix_pointers(const Flavor& f) {
  std::vector<ElemSpec> v;
  auto E = [&](int key, int r, int c) { v.push_back({key, r, c}); };
  const int D = B4N_D, DP = B4N_DP, S = B4N_S, SP = B4N_SP, GE = B4N_GE, GP = B4N_GP, GM = B4N_GM, B = B4N_B, BP = B4N_BP, DB = B4N_DB,
            SB = B4N_SB, Q = B4N_Q;
  E(MP_DPbp, DP, BP); E(MP_GPbp, GP, BP); E(MP_SPbp, SP, BP);
  E(MP_BPdp, BP, DP); E(MP_BPgp, BP, GP); E(MP_BPsp, BP, SP); E(MP_BPbp, BP, BP);
  E(MP_Dd, D, D); E(MP_GPgp, GP, GP); E(MP_Ss, S, S); E(MP_DPdp, DP, DP); E(MP_SPsp, SP, SP); E(MP_Ddp, D, DP); E(MP_GPdp, GP, DP);
  E(MP_GPsp, GP, SP); E(MP_Ssp, S, SP); E(MP_DPsp, DP, SP); E(MP_DPd, DP, D); E(MP_DPgp, DP, GP); E(MP_SPgp, SP, GP); E(MP_SPs, SP, S);
  E(MP_SPdp, SP, DP);
  E(MP_Qq, Q, Q); E(MP_Qbp, Q, BP); E(MP_Qdp, Q, DP); E(MP_Qsp, Q, SP); E(MP_Qgp, Q, GP); E(MP_DPq, DP, Q); E(MP_SPq, SP, Q); E(MP_GPq, GP, Q);
  if (f.rgatemod != 0) {
    E(MP_GEge, GE, GE); E(MP_GEgp, GE, GP); E(MP_GPge, GP, GE); E(MP_GEdp, GE, DP); E(MP_GEsp, GE, SP); E(MP_GEbp, GE, BP);
    E(MP_GMdp, GM, DP); E(MP_GMgp, GM, GP); E(MP_GMgm, GM, GM); E(MP_GMge, GM, GE); E(MP_GMsp, GM, SP); E(MP_GMbp, GM, BP);
    E(MP_DPgm, DP, GM); E(MP_GPgm, GP, GM); E(MP_GEgm, GE, GM); E(MP_SPgm, SP, GM); E(MP_BPgm, BP, GM);
  }
  if (f.rbodymod == 1 || f.rbodymod == 2) {
    E(MP_DPdb, DP, DB); E(MP_SPsb, SP, SB);
    E(MP_DBdp, DB, DP); E(MP_DBdb, DB, DB); E(MP_DBbp, DB, BP); E(MP_DBb, DB, B);
    E(MP_BPdb, BP, DB); E(MP_BPb, BP, B); E(MP_BPsb, BP, SB);
    E(MP_SBsp, SB, SP); E(MP_SBbp, SB, BP); E(MP_SBb, SB, B); E(MP_SBsb, SB, SB);
    E(MP_Bdb, B, DB); E(MP_Bbp, B, BP); E(MP_Bsb, B, SB); E(MP_Bb, B, B);
  }
  if (f.rdsmod != 0) {
    E(MP_Dgp, D, GP); E(MP_Dsp, D, SP); E(MP_Dbp, D, BP); E(MP_Sdp, S, DP); E(MP_Sgp, S, GP); E(MP_Sbp, S, BP);
  }
  return v;
}

// The G pushes of one load, in stamp order (stamp.rs:381-567): (itab slot, matrix pointer it lands on).
struct PushSpec { int slot, matp; };
inline std::vector<PushSpec> g_push_sequence(const Flavor& f) {
  std::vector<PushSpec> v;
  auto P = [&](int slot, int mp) { v.push_back({slot, mp}); };
  if (f.rgatemod == 1) {
    P(B4G_GEge, MP_GEge); P(B4G_GPge, MP_GPge); P(B4G_GEgp, MP_GEgp);
  } else if (f.rgatemod == 2) {
    P(B4G_GEge, MP_GEge); P(B4G_GEgp, MP_GEgp); P(B4G_GEdp, MP_GEdp); P(B4G_GEsp, MP_GEsp); P(B4G_GEbp, MP_GEbp); P(B4G_GPge, MP_GPge);
  } else if (f.rgatemod == 3) {
    P(B4G_GEge, MP_GEge); P(B4G_GEgm, MP_GEgm); P(B4G_GMge, MP_GMge); P(B4G_GMgm, MP_GMgm);
    P(B4G_GMdp, MP_GMdp); P(B4G_GMgp, MP_GMgp); P(B4G_GMsp, MP_GMsp); P(B4G_GMbp, MP_GMbp);
    P(B4G_DPgm, MP_DPgm); P(B4G_GPgm, MP_GPgm); P(B4G_SPgm, MP_SPgm); P(B4G_BPgm, MP_BPgm);
  }
  P(B4G_GPgp, MP_GPgp); P(B4G_GPdp, MP_GPdp); P(B4G_GPsp, MP_GPsp); P(B4G_GPbp, MP_GPbp);
  if (f.rdsmod != 0) {
    P(B4G_Dgp, MP_Dgp); P(B4G_Dsp, MP_Dsp); P(B4G_Dbp, MP_Dbp); P(B4G_Sdp, MP_Sdp); P(B4G_Sgp, MP_Sgp); P(B4G_Sbp, MP_Sbp);
  }
  P(B4G_DPdp, MP_DPdp); P(B4G_DPd, MP_DPd); P(B4G_DPgp, MP_DPgp); P(B4G_DPsp, MP_DPsp); P(B4G_DPbp, MP_DPbp);
  P(B4G_Ddp, MP_Ddp); P(B4G_Dd, MP_Dd);
  P(B4G_SPdp, MP_SPdp); P(B4G_SPgp, MP_SPgp); P(B4G_SPsp, MP_SPsp); P(B4G_SPs, MP_SPs); P(B4G_SPbp, MP_SPbp);
  P(B4G_Ssp, MP_Ssp); P(B4G_Ss, MP_Ss);
  P(B4G_BPdp, MP_BPdp); P(B4G_BPgp, MP_BPgp); P(B4G_BPsp, MP_BPsp); P(B4G_BPbp, MP_BPbp);
  P(B4G_L_DPdp, MP_DPdp); P(B4G_L_DPgp, MP_DPgp); P(B4G_L_DPsp, MP_DPsp); P(B4G_L_DPbp, MP_DPbp);
  P(B4G_L_BPdp, MP_BPdp); P(B4G_L_BPgp, MP_BPgp); P(B4G_L_BPsp, MP_BPsp); P(B4G_L_BPbp, MP_BPbp);
  P(B4G_S_SPdp, MP_SPdp); P(B4G_S_SPgp, MP_SPgp); P(B4G_S_SPsp, MP_SPsp); P(B4G_S_SPbp, MP_SPbp);
  P(B4G_S_BPdp, MP_BPdp); P(B4G_S_BPgp, MP_BPgp); P(B4G_S_BPsp, MP_BPsp); P(B4G_S_BPbp, MP_BPbp);
  if (f.rbodymod != 0) {
    P(B4G_DPdb, MP_DPdb); P(B4G_SPsb, MP_SPsb);
    P(B4G_DBdp, MP_DBdp); P(B4G_DBdb, MP_DBdb); P(B4G_DBbp, MP_DBbp); P(B4G_DBb, MP_DBb);
    P(B4G_BPdb, MP_BPdb); P(B4G_BPb, MP_BPb); P(B4G_BPsb, MP_BPsb); P(B4G_R_BPbp, MP_BPbp);
    P(B4G_SBsp, MP_SBsp); P(B4G_SBbp, MP_SBbp); P(B4G_SBb, MP_SBb); P(B4G_SBsb, MP_SBsb);
    P(B4G_Bdb, MP_Bdb); P(B4G_Bbp, MP_Bbp); P(B4G_Bsb, MP_Bsb); P(B4G_Bb, MP_Bb);
  }
  if (f.trnqsmod != 0) {
    P(B4G_Qq, MP_Qq); P(B4G_Qgp, MP_Qgp); P(B4G_Qdp, MP_Qdp); P(B4G_Qsp, MP_Qsp); P(B4G_Qbp, MP_Qbp);
    P(B4G_DPq, MP_DPq); P(B4G_SPq, MP_SPq); P(B4G_GPq, MP_GPq);
  }
  return v;
}
// The RHS pushes of one load, in stamp order (stamp.rs:338-368): (itab slot, node position whose variable it adds to).
inline std::vector<PushSpec> b_push_sequence(const Flavor& f) {
  std::vector<PushSpec> v;
  auto P = [&](int slot, int node) { v.push_back({slot, node}); };
  P(B4B_DP, B4N_DP); P(B4B_GP, B4N_GP);
  if (f.rgatemod == 2) P(B4B_GX, B4N_GE);
  else if (f.rgatemod == 3) P(B4B_GX, B4N_GM);
  if (f.rbodymod == 0) { P(B4B_BP, B4N_BP); P(B4B_SP, B4N_SP); }
  else { P(B4B_DB, B4N_DB); P(B4B_BP, B4N_BP); P(B4B_SB, B4N_SB); P(B4B_SP, B4N_SP); }
  if (f.rdsmod != 0) { P(B4B_D, B4N_D); P(B4B_S, B4N_S); }
  if (f.trnqsmod != 0) P(B4B_Q, B4N_Q);
  return v;
}

}  // namespace b4
}  // namespace s21
