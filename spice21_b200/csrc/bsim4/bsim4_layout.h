// BSIM4 device tables: terminal / stamp positions in the device's itab, state slots, and the flat per-device parameter
// block the evaluation reads (bsim4_eval.hpp). Shared by host C++ and CUDA.
//
// itab layout: [12 node variables][4 + 6 RHS push slots][G push slots]. Every push of bsim4/stamp.rs has its OWN slot
// (the reference pushes several values to one matrix element in a single load: DPdp/DPd/Ddp/Dd coincide when there is no
// drain resistance, and the GIDL/GISL block re-pushes elements of the main block). A slot that the device's flavour
// (rgatemod / rdsmod / rbodymod / trnqsmod) never pushes holds handle -1. The slots a default-flavour device uses come
// first so that its itab (and staging area) stays short.
#pragma once

namespace s21 {

// node variables (bsim4ports.rs:10-22)
enum B4Node { B4N_D = 0, B4N_DP, B4N_S, B4N_SP, B4N_GE, B4N_GP, B4N_GM, B4N_B, B4N_BP, B4N_DB, B4N_SB, B4N_Q, B4N_COUNT };

// RHS push slots (stamp.rs:338-368); B4B_GX is gNodeExt (rgatemod 2) or gNodeMid (rgatemod 3)
enum B4RhsSlot { B4B_DP = B4N_COUNT, B4B_GP, B4B_BP, B4B_SP, B4B_BASE_END,
                 B4G_BASE = B4B_BASE_END };

// G push slots. Base block = what rgatemod 0 / rdsmod 0 / rbodymod 0 / trnqsmod 0 pushes (stamp.rs:421-425, 436-501, 503-529).
enum B4GSlot {
  B4G_GPgp = B4G_BASE, B4G_GPdp, B4G_GPsp, B4G_GPbp,
  B4G_DPdp, B4G_DPd, B4G_DPgp, B4G_DPsp, B4G_DPbp, B4G_Ddp, B4G_Dd,
  B4G_SPdp, B4G_SPgp, B4G_SPsp, B4G_SPs, B4G_SPbp, B4G_Ssp, B4G_Ss,
  B4G_BPdp, B4G_BPgp, B4G_BPsp, B4G_BPbp,
  B4G_L_DPdp, B4G_L_DPgp, B4G_L_DPsp, B4G_L_DPbp, B4G_L_BPdp, B4G_L_BPgp, B4G_L_BPsp, B4G_L_BPbp,   // GIDL
  B4G_S_SPdp, B4G_S_SPgp, B4G_S_SPsp, B4G_S_SPbp, B4G_S_BPdp, B4G_S_BPgp, B4G_S_BPsp, B4G_S_BPbp,   // GISL
  B4_BASE_END,
  // optional RHS slots
  B4B_GX = B4_BASE_END, B4B_DB, B4B_SB, B4B_D, B4B_S, B4B_Q,
  // gate resistance network (stamp.rs:381-420)
  B4G_GEge, B4G_GPge, B4G_GEgp, B4G_GEdp, B4G_GEsp, B4G_GEbp,
  B4G_GEgm, B4G_GMge, B4G_GMgm, B4G_GMdp, B4G_GMgp, B4G_GMsp, B4G_GMbp, B4G_DPgm, B4G_GPgm, B4G_SPgm, B4G_BPgm,
  // bias-dependent S/D resistance (stamp.rs:427-434)
  B4G_Dgp, B4G_Dsp, B4G_Dbp, B4G_Sdp, B4G_Sgp, B4G_Sbp,
  // body resistance network (stamp.rs:531-555)
  B4G_DPdb, B4G_SPsb, B4G_DBdp, B4G_DBdb, B4G_DBbp, B4G_DBb, B4G_BPdb, B4G_BPb, B4G_BPsb, B4G_R_BPbp,
  B4G_SBsp, B4G_SBbp, B4G_SBb, B4G_SBsb, B4G_Bdb, B4G_Bbp, B4G_Bsb, B4G_Bb,
  // NQS charge node (stamp.rs:557-567)
  B4G_Qq, B4G_Qgp, B4G_Qdp, B4G_Qsp, B4G_Qbp, B4G_DPq, B4G_SPq, B4G_GPq,
  B4_ITAB_MAX
};

// State slots: the previous iteration's limited biases (Bsim4OpPoint fields read back through `self.guess`,
// bsim4solver.rs:426-520) and the charges whose committed values feed dq/dt (tran.rs:449-470, 494).
enum B4State {
  B4S_VGS = 0, B4S_VDS, B4S_VBS, B4S_VGES, B4S_VGMS, B4S_VDBS, B4S_VSBS, B4S_VSES, B4S_VDES, B4S_VBD, B4S_VDBD, B4S_VON,
  B4S_QB, B4S_QG, B4S_QD, B4S_QGMID, B4S_QBS, B4S_QBD, B4S_QCDUMP, B4S_QCHEQ,
  B4S_COUNT
};

// Flat parameter block: model card entries the evaluation reads (m_), model-derived (d_), size-dependent (s_) and
// per-instance (i_) precomputed values. Index = position in this enum; all doubles.
#define B4_MODEL_CARD_FIELDS(X) \
  X(mobmod) X(diomod) X(capmod) X(rdsmod) X(rbodymod) X(rgatemod) X(trnqsmod) X(mtrlmod) X(mtrlcompatmod) \
  X(igcmod) X(igbmod) X(tempmod) X(gidlmod) X(cvchargemod) \
  X(epsrox) X(toxe) X(epsrgate) X(phig) X(easub) X(xjbvs) X(xjbvd) X(bvs) X(bvd) \
  X(vtss) X(vtsd) X(vtssws) X(vtsswd) X(vtsswgs) X(vtsswgd) X(lambda) X(vtl) X(pditsl) \
  X(xpart) X(ados) X(bdos) X(epsrsub) X(mjs) X(mjd) X(mjsws) X(mjswd) X(mjswgs) X(mjswgd) X(pigcd)
// type_sign: +1 NMOS / -1 PMOS; vtl_given: 1 when the card set vtl (bsim4solver.rs:1826)
#define B4_MODEL_EVAL_FIELDS(X) X(type_sign) X(vtl_given) B4_MODEL_CARD_FIELDS(X)
#define B4_DERIVED_EVAL_FIELDS(X) \
  X(coxp) X(Eg0) X(vtm) X(vtm0) X(coxe) X(vcrit) X(factor1) X(PhiBS) X(PhiBSWS) X(PhiBSWGS) X(PhiBD) X(PhiBSWD) X(PhiBSWGD) \
  X(SunitAreaTempJctCap) X(DunitAreaTempJctCap) X(SunitLengthSidewallTempJctCap) X(DunitLengthSidewallTempJctCap) \
  X(SunitLengthGateSidewallTempJctCap) X(DunitLengthGateSidewallTempJctCap) X(TempRatio) X(epssub) X(ni) \
  X(Nvtms) X(Nvtmd) X(Nvtmrss) X(Nvtmrssws) X(Nvtmrsswgs) X(Nvtmrsd) X(Nvtmrsswd) X(Nvtmrsswgd)

enum B4Field {
#define X(n) B4F_m_##n,
  B4_MODEL_EVAL_FIELDS(X)
#undef X
#define X(n) B4F_d_##n,
  B4_DERIVED_EVAL_FIELDS(X)
#undef X
#define B4S(n) B4F_s_##n,
#define B4I(n) B4F_i_##n,
#include "bsim4_fields.inc"
#undef B4S
#undef B4I
  B4F_COUNT
};

}  // namespace s21
