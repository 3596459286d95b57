// Staged assembly tables for the cooperative kernel (kernels/coop.cu). Devices are evaluated in parallel, each writing
// its stamps to private staging slots (one per itab position); a gather phase then forms every matrix element / RHS row
// by summing its staging slots in the order the reference's `Solver::update` would have added them
// (spice21/src/analysis.rs:153-168): component order, then push order inside the device.
// For devices whose drain/source roles swap at run time (Mos0, Mos1) the push order of the non-reversed orientation is
// used; the two orientations only disagree about the relative order of stamps that land on the SAME element, which
// happens only for a device whose drain and source are the same node.
#pragma once
#include <algorithm>

#include "flatten.hpp"
#include "symbolic.hpp"

namespace s21 {

struct StageInfo {
  std::vector<int> stage_off;   // per device
  std::vector<int> eval_order;  // device ids sorted by type (stable)
  int n_stage = 0;
};

inline StageInfo make_stage_info(const FlatCkt& flat) {
  StageInfo si;
  for (const FlatDev& d : flat.devs) {
    si.stage_off.push_back(si.n_stage);
    si.n_stage += d.type == DT_MOS1 ? M1_NSTAGE : d.n_itab;
  }
  si.eval_order.resize(flat.devs.size());
  for (size_t k = 0; k < flat.devs.size(); k++) si.eval_order[k] = (int)k;
  // heaviest device types first, so that the warps that get them start early
  auto weight = [](int t) { return t == DT_BSIM4 ? 0 : t == DT_MOS1 ? 1 : t == DT_DIODE ? 2 : t == DT_MOS0 ? 3 : 4 + t; };
  std::stable_sort(si.eval_order.begin(), si.eval_order.end(),
                   [&](int a, int b) { return weight(flat.devs[(size_t)a].type) < weight(flat.devs[(size_t)b].type); });
  return si;
}

// Canonical push lists. A G entry is (itab position of the element handle, staging position); they differ only for
// Mos1's duplicated AC stamp. A b entry is the itab position of the variable.
struct PushLists {
  std::vector<std::pair<int, int>> g;
  std::vector<int> b;
};
inline PushLists push_lists(const FlatDev& dev, int mode) {
  PushLists L;
  const int type = dev.type;
  auto G = [&](int pos) { L.g.push_back({pos, pos}); };
  auto m1 = [](int a, int b) { return M1_E0 + a * 6 + b; };
  switch (type) {
    case DT_R:  // comps/mod.rs:301-308, 313-320
      G(R_EPP); G(R_ENN); G(R_EPN); G(R_ENP);
      break;
    case DT_C:  // comps/mod.rs:205-237: nothing in OP
      if (mode != AN_OP) { G(R_EPP); G(R_ENN); G(R_EPN); G(R_ENP); }
      if (mode == AN_TRAN) L.b = {R_P, R_N};
      break;
    case DT_I:  // comps/mod.rs:341-344
      L.b = {I_P, I_N};
      break;
    case DT_V:  // comps/mod.rs:134-137, 140-148
      G(V_EPI); G(V_EIP); G(V_ENI); G(V_EIN);
      L.b = {V_I};
      break;
    case DT_DIODE:  // diode.rs:342-353
      G(D_ENN); G(D_ERN); G(D_ENR); G(D_ERR); G(D_EPP); G(D_EPR); G(D_ERP);
      L.b = {D_R, D_N};
      break;
    case DT_MOS0:  // mos.rs:1088-1098 with (sr, dr) = (S, D)
      G(M0_EDD); G(M0_ESS); G(M0_EDS); G(M0_ESD); G(M0_EDG); G(M0_ESG);
      L.b = {M0_D, M0_S};
      break;
    case DT_MOS1: {
      const int dr = M1_DP, sr = M1_SP, dx = M1_D, sx = M1_S, g = M1_G, b = M1_B;
      if (mode != AN_AC) {  // mos.rs:838-868
        const int pairs[22][2] = {{dr, dr}, {sr, sr}, {dr, sr}, {sr, dr}, {dr, g}, {sr, g}, {g, g}, {b, b}, {g, b}, {g, dr}, {g, sr},
                                  {b, g}, {b, dr}, {b, sr}, {dr, b}, {sr, b}, {dx, dr}, {dr, dx}, {dx, dx}, {sx, sr}, {sr, sx}, {sx, sx}};
        for (auto& pr : pairs) G(m1(pr[0], pr[1]));
        L.b = {dr, sr, g, b};
      } else {  // mos.rs:940-966: (G,dr) is pushed twice
        const int pairs[23][2] = {{dr, dr}, {sr, sr}, {dr, sr}, {sr, dr}, {dr, g}, {sr, g}, {g, g}, {b, b}, {g, b}, {g, dr}, {g, sr},
                                  {b, g}, {g, dr}, {b, dr}, {b, sr}, {dr, b}, {sr, b}, {dx, dr}, {dr, dx}, {dx, dx}, {sx, sr}, {sr, sx}, {sx, sx}};
        for (int k = 0; k < 23; k++) {
          if (k == 12) L.g.push_back({m1(g, dr), M1_DUP_GDR});
          else G(m1(pairs[k][0], pairs[k][1]));
        }
      }
      break;
    }
    case DT_BSIM4:  // stamp.rs:338-567: same pushes in OP and TRAN; each has its own slot (bsim4_layout.h)
      for (int slot : dev.push_g) G(slot);
      L.b = dev.push_b;
      break;
    default: break;
  }
  return L;
}

// Fill P.asm_off / P.asm_src / P.n_stage for analysis `mode`. `itab` holds L+U slots as element handles.
inline void build_gather(const FlatCkt& flat, const StageInfo& si, int mode, const std::vector<int>& itab, Plan& P) {
  const int nt = P.nnzLU + P.N;
  std::vector<std::vector<int>> src((size_t)nt);
  for (size_t k = 0; k < flat.devs.size(); k++) {
    const FlatDev& d = flat.devs[k];
    const PushLists L = push_lists(d, mode);
    const int* t = itab.data() + d.itab_off;
    for (auto& ge : L.g) {
      const int h = t[ge.first];
      if (h >= 0) src[(size_t)h].push_back(si.stage_off[k] + ge.second);
    }
    for (int pos : L.b) {
      const int v = t[pos];
      if (v >= 0) src[(size_t)P.nnzLU + (size_t)v].push_back(si.stage_off[k] + pos);
    }
  }
  P.asm_off.assign((size_t)nt + 1, 0);
  P.asm_src.clear();
  for (int q = 0; q < nt; q++) {
    P.asm_src.insert(P.asm_src.end(), src[(size_t)q].begin(), src[(size_t)q].end());
    P.asm_off[(size_t)q + 1] = (int)P.asm_src.size();
  }
  P.n_stage = si.n_stage;
}

}  // namespace s21
