// Elaboration for the GPU path: hierarchy flattening, variable numbering, element creation (the stamp map) and the
// per-device SoA tables. Restates the *behaviour* of
//   spice21/src/elab.rs:35-265           (first-encounter variable numbering, namespaces, auto-noding at top level)
//   spice21/src/analysis.rs:308-330      (Solver::new: create_matrix_elems per component, in component order)
//   spice21/src/analysis.rs:510-525      (Tran::ic: two extra variables + a 1 S resistor + a source per initial condition)
//   spice21/src/sparse21/mod.rs:265-270  (Matrix::make: get-or-create => element id == creation order)
// The output is flat integer/double tables (device_layout.h), not a vector of polymorphic solver objects.
#pragma once
#include <unordered_map>

#include "../bsim4/bsim4_pack.hpp"
#include "params.hpp"

namespace s21 {

struct FlatDev {
  int type = 0;
  std::string path;              // flattened instance path ("x1.p"); the key for per-instance overrides
  int itab_off = 0, par_off = 0, state_off = 0;
  int n_itab = 0, n_par = 0, n_state = 0;
  std::vector<int> created;      // element handles in create_matrix_elems order (-1 = ground): stamp-map export
  std::string model, params;     // definition names (Mos1 / Diode / Bsim4), for re-derivation under overrides
  bool is_ic = false;
  std::vector<int> push_g, push_b;  // Bsim4 only: itab slots of this flavour's G / RHS pushes, in push order (bsim4_pack.hpp)
};

struct FlatCkt {
  SimOptions opts;
  std::vector<std::string> var_names;
  std::vector<int> var_kinds;            // 0 V, 1 I (analysis.rs:34-38)
  std::vector<FlatDev> devs;
  std::vector<int> itab;                 // nodes + element ids
  std::vector<double> par;               // shared (instance-independent) parameter values
  std::vector<int> elem_row, elem_col;   // element id -> original (row, col)
  int n_state = 0;
  int n_stamps = 0;                      // total G + b pushes per load sweep (bookkeeping for the roofline record)
  std::vector<int> dev_off_export, dev_elems_export;
  int n_vars() const { return (int)var_names.size(); }
  int n_elems() const { return (int)elem_row.size(); }
};

class Flattener {
 public:
  Flattener(const CktSpec& spec, const SimOptions& opts) : spec_(spec) { out_.opts = opts; }

  FlatCkt run(const std::vector<std::pair<std::string, double>>& ics) {
    Namespace top;
    top[""] = -1;
    for (const std::string& s : spec_.signals) declare_signal(s, top);
    for (const CompSpec& c : spec_.comps) instance(c, top, /*autonode=*/true);
    for (auto& ic : ics) initial_condition(ic.first, ic.second);
    out_.dev_off_export.push_back(0);
    for (const FlatDev& d : out_.devs) {
      out_.dev_elems_export.insert(out_.dev_elems_export.end(), d.created.begin(), d.created.end());
      out_.dev_off_export.push_back((int)out_.dev_elems_export.size());
    }
    return std::move(out_);
  }

 private:
  typedef std::map<std::string, int> Namespace;  // node name -> variable index (-1 = ground)
  const CktSpec& spec_;
  FlatCkt out_;
  std::vector<std::string> path_;
  std::unordered_map<std::string, int> first_var_;     // name -> first variable carrying it
  std::unordered_map<uint64_t, int> elem_at_;          // (row, col) -> element id

  std::string joined(const std::string& leaf = std::string()) const {
    std::string s;
    for (const std::string& p : path_) { if (!s.empty()) s += "."; s += p; }
    if (!leaf.empty()) { if (!s.empty()) s += "."; s += leaf; }
    return s;
  }
  int new_var(const std::string& name, int kind) {  // Variables::add — no duplicate check (analysis.rs:67-73)
    out_.var_names.push_back(name);
    out_.var_kinds.push_back(kind);
    int idx = (int)out_.var_names.size() - 1;
    first_var_.emplace(name, idx);
    return idx;
  }
  int find_or_create(const std::string& name) {  // analysis.rs:92-111
    if (name.empty()) return -1;
    auto it = first_var_.find(name);
    if (it != first_var_.end()) return it->second;
    return new_var(name, 0);
  }
  void declare_signal(const std::string& s, Namespace& ns) { ns[s] = new_var(joined(s), 0); }  // elab.rs:209-216
  int lookup(const Namespace& ns, const std::string& node) const {
    auto it = ns.find(node);
    if (it == ns.end()) throw S21Error(ST_INVALID, "node '" + node + "' is not declared in this scope");  // elab.rs:49, :91-92, :198
    return it->second;
  }
  // elab.rs:35-52. With auto-noding (top level only) an unseen name becomes a new variable on first encounter.
  int node(const std::string& n, bool autonode, Namespace& ns) {
    if (!autonode) return lookup(ns, n);
    if (n.empty()) return -1;
    int v = find_or_create(joined(n));
    ns[n] = v;
    return v;
  }
  int element(int row, int col) {  // make_matrix_elem (comps/mod.rs:349-354) over Matrix::make
    if (row < 0 || col < 0) return -1;
    uint64_t key = ((uint64_t)(uint32_t)row << 32) | (uint32_t)col;
    auto it = elem_at_.find(key);
    if (it != elem_at_.end()) return it->second;
    int id = (int)out_.elem_row.size();
    out_.elem_row.push_back(row);
    out_.elem_col.push_back(col);
    elem_at_.emplace(key, id);
    return id;
  }
  FlatDev& begin_dev(int type, const std::string& path, int n_itab, int n_par, int n_state, int n_stamps) {
    FlatDev d;
    d.type = type;
    d.path = path;
    d.itab_off = (int)out_.itab.size();
    d.par_off = (int)out_.par.size();
    d.state_off = out_.n_state;
    d.n_itab = n_itab; d.n_par = n_par; d.n_state = n_state;
    out_.itab.resize(out_.itab.size() + (size_t)n_itab, -1);
    out_.par.resize(out_.par.size() + (size_t)n_par, 0.0);
    out_.n_state += n_state;
    out_.n_stamps += n_stamps;
    out_.devs.push_back(d);
    return out_.devs.back();
  }
  int* it(const FlatDev& d) { return out_.itab.data() + d.itab_off; }
  double* pp(const FlatDev& d) { return out_.par.data() + d.par_off; }

  // Two-terminal pattern shared by Resistor and Capacitor: (p,p) (p,n) (n,p) (n,n)  (comps/mod.rs:190-195, 291-298)
  void two_term_elems(FlatDev& d, int p, int n) {
    int* t = it(d);
    t[R_P] = p; t[R_N] = n;
    t[R_EPP] = element(p, p); t[R_EPN] = element(p, n); t[R_ENP] = element(n, p); t[R_ENN] = element(n, n);
    d.created = {t[R_EPP], t[R_EPN], t[R_ENP], t[R_ENN]};
  }
  void add_resistor(const std::string& path, int p, int n, double g_op, double g_tran, bool is_ic) {
    FlatDev& d = begin_dev(DT_R, path, R_NI, RP_N, 0, 4);
    two_term_elems(d, p, n);
    pp(d)[RP_G_OP] = g_op; pp(d)[RP_G_TRAN] = g_tran;
    d.is_ic = is_ic;
  }
  void add_vsrc(const std::string& path, int p, int n, int ivar, double v_op, double v_tran, double acm, bool is_ic, int wave_kind = 0,
                const double* wave = nullptr) {
    FlatDev& d = begin_dev(DT_V, path, V_NI, VP_N, 0, 5);
    int* t = it(d);
    t[V_P] = p; t[V_N] = n; t[V_I] = ivar;
    t[V_EPI] = element(p, ivar); t[V_EIP] = element(ivar, p); t[V_ENI] = element(n, ivar); t[V_EIN] = element(ivar, n);  // comps/mod.rs:127-132
    d.created = {t[V_EPI], t[V_EIP], t[V_ENI], t[V_EIN]};
    pp(d)[VP_V_OP] = v_op; pp(d)[VP_V_TRAN] = v_tran; pp(d)[VP_ACM] = acm;
    pp(d)[VP_WKIND] = (double)wave_kind;
    for (int k = 0; k < 7; k++) pp(d)[VP_W0 + k] = (wave_kind && wave) ? wave[k] : 0.0;
    d.is_ic = is_ic;
  }

  void instance(const CompSpec& c, Namespace& ns, bool autonode) {  // elab.rs:55-84
    switch (c.kind) {
      case CK_R: {
        int p = node(c.p, autonode, ns), n = node(c.n, autonode, ns);
        add_resistor(joined(c.name), p, n, c.val, c.val, false);
        break;
      }
      case CK_C: {
        int p = node(c.p, autonode, ns), n = node(c.n, autonode, ns);
        FlatDev& d = begin_dev(DT_C, joined(c.name), R_NI, CP_N, CS_N, 6);
        two_term_elems(d, p, n);
        pp(d)[CP_C] = c.val;
        break;
      }
      case CK_I: {
        int p = node(c.p, autonode, ns), n = node(c.n, autonode, ns);
        FlatDev& d = begin_dev(DT_I, joined(c.name), I_NI, IP_N, 0, 2);
        it(d)[I_P] = p; it(d)[I_N] = n;
        pp(d)[IP_I] = c.val;
        break;
      }
      case CK_V: {  // elab.rs:112-127: nodes first, then the branch-current variable named after the instance path
        bool top = path_.empty();
        int p = node(c.p, top, ns), n = node(c.n, top, ns);
        int ivar = new_var(joined(c.name), 1);
        add_vsrc(joined(c.name), p, n, ivar, c.val, c.val, c.acm, false, c.wave_kind, c.wave);
        break;
      }
      case CK_D: diode(c, ns); break;
      case CK_MOS: mosfet(c, ns); break;
      case CK_X: module_instance(c, ns); break;
    }
  }
  void diode(const CompSpec& c, Namespace& ns) {  // elab.rs:85-111 — never auto-nodes
    int p = lookup(ns, c.p), n = lookup(ns, c.n);
    auto mi = spec_.diode_models.find(c.model);
    auto ii = spec_.diode_insts.find(c.params);
    if (mi == spec_.diode_models.end() || ii == spec_.diode_insts.end()) throw S21Error(ST_INVALID, "Parameters not defined: " + c.params);
    DiodeDerived dd = diode_derive(mi->second, ii->second, out_.opts);
    std::string path = joined(c.name);
    int r = dd.has_r ? new_var(path + ".r", 0) : p;
    FlatDev& d = begin_dev(DT_DIODE, path, D_NI, DP_N, DS_N, 9);
    int* t = it(d);
    t[D_P] = p; t[D_N] = n; t[D_R] = r;
    t[D_EPP] = element(p, p); t[D_EPR] = element(p, r); t[D_ERP] = element(r, p); t[D_ERR] = element(r, r);  // diode.rs:248-256
    t[D_ENR] = element(n, r); t[D_ERN] = element(r, n); t[D_ENN] = element(n, n);
    d.created = {t[D_EPP], t[D_EPR], t[D_ERP], t[D_ERR], t[D_ENR], t[D_ERN], t[D_ENN]};
    for (int k = 0; k < DP_N; k++) pp(d)[k] = dd.par[k];
    d.model = c.model; d.params = c.params;
  }
  void mosfet(const CompSpec& c, Namespace& ns) {  // elab.rs:128-173
    bool top = path_.empty();
    int d_ = node(c.d, top, ns), g_ = node(c.g, top, ns), s_ = node(c.s, top, ns), b_ = node(c.b, top, ns);
    std::string path = joined(c.name);
    if (spec_.bsim4_models.count(c.model)) {
      bsim4(c, path, d_, g_, s_, b_);
    } else if (spec_.mos1_models.count(c.model)) {
      auto ii = spec_.mos1_insts.find(c.params);
      if (ii == spec_.mos1_insts.end()) throw S21Error(ST_INVALID, "Parameters not defined: " + c.params);
      Mos1Derived md = mos1_derive(spec_.mos1_models.at(c.model), ii->second, out_.opts);
      int dp = md.has_dp ? new_var(path + ".dp", 0) : d_;  // mos.rs:599-610: dp first, then sp
      int sp = md.has_sp ? new_var(path + ".sp", 0) : s_;
      FlatDev& d = begin_dev(DT_MOS1, path, M1_NI, M1P_N, M1S_N, 26);
      int* t = it(d);
      t[M1_D] = d_; t[M1_G] = g_; t[M1_S] = s_; t[M1_B] = b_; t[M1_DP] = dp; t[M1_SP] = sp;
      static const int order[6] = {M1_G, M1_D, M1_S, M1_B, M1_DP, M1_SP};  // mos.rs:896-903
      for (int a : order)
        for (int b : order) {
          int e = element(t[a], t[b]);
          t[M1_E0 + a * 6 + b] = e;
          d.created.push_back(e);
        }
      for (int k = 0; k < M1P_N; k++) pp(d)[k] = md.par[k];
      d.model = c.model; d.params = c.params;
    } else if (spec_.mos0.count(c.model)) {
      FlatDev& d = begin_dev(DT_MOS0, path, M0_NI, M0P_N, 0, 8);
      int* t = it(d);
      t[M0_D] = d_; t[M0_G] = g_; t[M0_S] = s_;
      t[M0_EDD] = element(d_, d_); t[M0_ESS] = element(s_, s_); t[M0_EDS] = element(d_, s_);  // mos.rs:1044-1050
      t[M0_ESD] = element(s_, d_); t[M0_EDG] = element(d_, g_); t[M0_ESG] = element(s_, g_);
      d.created = {t[M0_EDD], t[M0_ESS], t[M0_EDS], t[M0_ESD], t[M0_EDG], t[M0_ESG]};
      pp(d)[M0P_P] = spec_.mos0.at(c.model) == 1 ? -1.0 : 1.0;
    } else {
      throw S21Error(ST_INVALID, "Model not defined: " + c.model);
    }
  }
  // Bsim4 (elab.rs:143-146): internal variables per bsim4ports.rs:24-112, elements in the order of bsim4solver.rs:28-115
  void bsim4(const CompSpec& c, const std::string& path, int d_, int g_, int s_, int b_) {
    auto ii = spec_.bsim4_insts.find(c.params);
    if (ii == spec_.bsim4_insts.end()) throw S21Error(ST_INVALID, "Parameters not defined: " + c.params);
    const MosModelSpec& mm = spec_.bsim4_models.at(c.model);
    b4::Derived bd;
    try {
      bd = b4::derive_device(mm.mos_type, mm.p.kv, ii->second.kv);
    } catch (const b4::ModelError& e) {
      throw S21Error(ST_INVALID, e.what());
    }
    const b4::Flavor& f = bd.flavor;
    int nodev[B4N_COUNT];
    nodev[B4N_D] = d_; nodev[B4N_S] = s_; nodev[B4N_GE] = g_; nodev[B4N_B] = b_;
    nodev[B4N_DP] = f.drain_source_prime ? new_var(path + ".drain", 0) : d_;
    nodev[B4N_SP] = f.drain_source_prime ? new_var(path + ".source", 0) : s_;
    nodev[B4N_GP] = f.rgatemod > 0 ? new_var(path + ".gate", 0) : g_;
    nodev[B4N_GM] = f.rgatemod == 3 ? new_var(path + ".midgate", 0) : g_;
    if (f.rbodymod == 1 || f.rbodymod == 2) {
      nodev[B4N_DB] = new_var(path + ".dbody", 0);
      nodev[B4N_BP] = new_var(path + ".body", 0);
      nodev[B4N_SB] = new_var(path + ".sbody", 0);
    } else {
      nodev[B4N_DB] = b_; nodev[B4N_BP] = b_; nodev[B4N_SB] = b_;
    }
    nodev[B4N_Q] = f.trnqsmod != 0 ? new_var(path + ".charge", 2) : -1;
    const std::vector<b4::PushSpec> pg = b4::g_push_sequence(f), pb = b4::b_push_sequence(f);
    int n_itab = B4N_COUNT;
    for (auto& p : pg) n_itab = std::max(n_itab, p.slot + 1);
    for (auto& p : pb) n_itab = std::max(n_itab, p.slot + 1);
    FlatDev& d = begin_dev(DT_BSIM4, path, n_itab, B4F_COUNT, B4S_COUNT, (int)(pg.size() + pb.size()));
    int* t = it(d);
    for (int k = 0; k < B4N_COUNT; k++) t[k] = nodev[k];
    int mp_elem[b4::MP_COUNT];
    for (int& e : mp_elem) e = -1;
    for (const b4::ElemSpec& e : b4::matrix_pointers(f)) {
      mp_elem[e.slot_key] = element(nodev[e.row], nodev[e.col]);
      d.created.push_back(mp_elem[e.slot_key]);
    }
    for (auto& p : pg) { t[p.slot] = mp_elem[p.matp]; d.push_g.push_back(p.slot); }
    for (auto& p : pb) { t[p.slot] = nodev[p.matp]; d.push_b.push_back(p.slot); }
    for (int k = 0; k < B4F_COUNT; k++) pp(d)[k] = bd.par[(size_t)k];
    d.model = c.model; d.params = c.params;
  }
  void module_instance(const CompSpec& c, Namespace& ns) {  // elab.rs:182-239
    auto mi = spec_.modules.find(c.module);
    if (mi == spec_.modules.end()) throw S21Error(ST_INVALID, "ModuleDef not found: " + c.module);
    Namespace inner;
    for (auto& kv : c.ports) inner[kv.first] = lookup(ns, kv.second);
    path_.push_back(c.name);
    if (path_.size() > 1024) throw S21Error(ST_INVALID, "Elaboration Error: Too deep a hierarchy (for now)!");
    for (const std::string& s : mi->second.signals) declare_signal(s, inner);
    for (const CompSpec& sub : mi->second.comps) instance(sub, inner, /*autonode=*/false);
    path_.pop_back();
  }
  void initial_condition(const std::string& n, double val) {  // analysis.rs:510-525
    int fnode = new_var("." + n + ".vic", 0);
    int ivar = new_var("." + n + ".iic", 1);
    int target = find_or_create(n);
    add_resistor("." + n + ".ric", fnode, target, 1.0, 1e-9, true);  // g = 1 during OP, 1e-9 afterwards (:517, :547-549)
    add_vsrc("." + n + ".vic", fnode, -1, ivar, val, 0.0, 0.0, true);  // forced to `val` during OP, 0 V afterwards (:521, :544-546)
  }
};

}  // namespace s21
