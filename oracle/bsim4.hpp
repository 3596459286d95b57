// ORACLE — TEST INFRASTRUCTURE ONLY (see num.hpp header).
//
// BSIM4 for the CPU oracle.
//
// READ THIS BEFORE TRUSTING IT: unlike every other oracle component, the BSIM4 *device equations* here are NOT an
// independent restatement. The model is ~6000 lines of arithmetic; it exists once in this repository, as the
// host/device headers under spice21_b200/csrc/bsim4/ (each citing the reference file:line it follows), and this file
// compiles those same headers for the CPU. What IS independent, and what this oracle therefore checks, is everything
// around the device: variable creation order (bsim4ports.rs:24-112), matrix-element creation order
// (bsim4solver.rs:28-115), the Stamps push order (stamp.rs:338-567), commit semantics (bsim4solver.rs:3765-3767), and
// the reference's sparse21 / Newton / transient code (sparse21.hpp, solver.hpp), against which the product's frozen
// -pivot GPU solver is compared. The shared equations are pinned instead by the reference's own artefacts:
//   * its three BSIM4 golden waveforms (tests/golden/test_bsim4_{cmos,nmos,pmos}_ro_tran.npz, from spice21/src/tests.rs
//     :948-1374), reproduced by this oracle in tests/test_oracle.py;
//   * its known answers (bsim4/tests.rs:57-118: diode-connected NMOS 150 uA +- 1 uA, PMOS 57 uA +- 1 uA, inverter).
#pragma once
#include <map>
#include <memory>

#ifdef S21_B4_COUNT
#include "bsim4_count.hpp"  // instrumented build: defines B4_DIV and the counting exp / log / sqrt before the evaluation headers
#endif
#include "../spice21_b200/csrc/bsim4/bsim4_eval.hpp"
#include "../spice21_b200/csrc/bsim4/bsim4_pack.hpp"
#include "circuit.hpp"

namespace orc {

// Bsim4Cache (bsim4/cache.rs:50-75): model and instance cards by name, derived pairs cached on first use.
struct Bsim4Depot {
  std::map<std::string, std::pair<int, Specs>> models;
  std::map<std::string, Specs> insts;
  std::map<std::pair<std::string, std::string>, std::shared_ptr<s21::b4::Derived>> cache;
  std::shared_ptr<s21::b4::Derived> get(const std::string& model, const std::string& inst) {
    auto key = std::make_pair(model, inst);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    auto mi = models.find(model);
    auto ii = insts.find(inst);
    if (mi == models.end() || ii == insts.end()) return nullptr;
    std::shared_ptr<s21::b4::Derived> d;
    try {
      d = std::make_shared<s21::b4::Derived>(s21::b4::derive_device(mi->second.first, mi->second.second.d, ii->second.d));
    } catch (const s21::b4::ModelError& e) {
      throw Panic(e.what());
    }
    cache[key] = d;
    return d;
  }
};

struct Bsim4 : Component {
  VarIndex ports[s21::B4N_COUNT];
  std::shared_ptr<s21::b4::Derived> prm;
  std::vector<s21::b4::ElemSpec> mps;   // matrix pointers in creation order
  Eindex matps[s21::b4::MP_COUNT];
  std::vector<Eindex> created;
  double guess[s21::B4S_COUNT], op[s21::B4S_COUNT];
  int slot_elem[s21::B4_ITAB_MAX];      // G push slot -> Eindex, RHS push slot -> VarIndex, node slot -> VarIndex

  Bsim4() {
    for (auto& g : guess) g = 0.0;
    for (auto& o : op) o = 0.0;
    for (auto& m : matps) m = -1;
    for (auto& s : slot_elem) s = -1;
  }
  const char* kind_name() const override { return "Bsim4"; }
  template <class T> void create(Matrix<T>& mat) {  // bsim4solver.rs:28-115
    created.clear();
    for (const auto& s : mps) {
      matps[s.slot_key] = make_matrix_elem(mat, ports[s.row], ports[s.col]);
      created.push_back(matps[s.slot_key]);
    }
    for (int k = 0; k < s21::B4N_COUNT; k++) slot_elem[k] = ports[k];
    for (const auto& p : s21::b4::g_push_sequence(prm->flavor)) slot_elem[p.slot] = matps[p.matp];
    for (const auto& p : s21::b4::b_push_sequence(prm->flavor)) slot_elem[p.slot] = ports[p.matp];
  }
  void create_matrix_elems(Matrix<double>& mat) override { create(mat); }
  void create_matrix_elems(Matrix<Cplx>& mat) override { create(mat); }
  void matps_list(std::vector<Eindex>& out) const override { out.insert(out.end(), created.begin(), created.end()); }
  void commit() override { for (int k = 0; k < s21::B4S_COUNT; k++) op[k] = guess[k]; }

  // Env over the oracle's own Variables / Stamps (concept: spice21_b200/csrc/kernels/devices.cuh)
  struct Env {
    Bsim4* dev;
    const Variables<double>* vars;
    Stamps<double>* out;
    int mode;
    double dt, gmin, omega;
    int node(int k) const { return dev->slot_elem[k]; }
    double par(int k) const { return dev->prm->par[(size_t)k]; }
    double volt(int var) const { return vars->get(var); }
    double op(int k) const { return dev->op[k]; }
    double guess(int k) const { return dev->guess[k]; }
    void set_guess(int k, double v) { dev->guess[k] = v; }
    void add_g_at(int pos, double v) { out->g.push_back({dev->slot_elem[pos], v}); }
    void add_b_at(int pos, double v) { out->b.push_back({dev->slot_elem[pos], v}); }
  };
  Stamps<double> load(const Variables<double>& vars, const AnalysisInfo& an, const Options& opts) override {  // bsim4solver.rs:134-144
    Stamps<double> st;
    Env e{this, &vars, &st, an.kind == AnalysisInfo::TRAN ? (int)s21::AN_TRAN : (int)s21::AN_OP, an.tran ? an.tran->dt : 0.0, opts.gmin, 0.0};
#ifdef S21_B4_COUNT
    s21::b4e::b4_counts().evals++;
#endif
    s21::b4e::load_bsim4(e);
    return st;
  }
};

// Bsim4Ports::from (bsim4ports.rs:24-112): internal variables are created in this order, named "<path>.<what>".
template <class T>
inline void bsim4_make_ports(Bsim4& b, const std::string& path, const VarIndex terms[4], Variables<T>& vars) {
  using namespace s21;
  const b4::Flavor& f = b.prm->flavor;
  const VarIndex d = terms[0], g = terms[1], s = terms[2], bb = terms[3];
  b.ports[B4N_D] = d; b.ports[B4N_S] = s; b.ports[B4N_GE] = g; b.ports[B4N_B] = bb;
  b.ports[B4N_DP] = f.drain_source_prime ? vars.add(path + ".drain", VarKind::V) : d;
  b.ports[B4N_SP] = f.drain_source_prime ? vars.add(path + ".source", VarKind::V) : s;
  b.ports[B4N_GP] = f.rgatemod > 0 ? vars.add(path + ".gate", VarKind::V) : g;
  b.ports[B4N_GM] = f.rgatemod == 3 ? vars.add(path + ".midgate", VarKind::V) : g;
  if (f.rbodymod == 1 || f.rbodymod == 2) {
    b.ports[B4N_DB] = vars.add(path + ".dbody", VarKind::V);
    b.ports[B4N_BP] = vars.add(path + ".body", VarKind::V);
    b.ports[B4N_SB] = vars.add(path + ".sbody", VarKind::V);
  } else {
    b.ports[B4N_DB] = bb; b.ports[B4N_BP] = bb; b.ports[B4N_SB] = bb;
  }
  b.ports[B4N_Q] = f.trnqsmod != 0 ? vars.add(path + ".charge", VarKind::Q) : -1;
}

inline std::shared_ptr<Component> bsim4_elaborate(Defs& defs, const std::string& model, const std::string& params, const std::string& path,
                                                  const VarIndex ports[4], void* vars_any, bool cplx, const Options&) {
  auto depot = std::static_pointer_cast<Bsim4Depot>(defs.bsim4);
  if (!depot || !depot->models.count(model)) return nullptr;
  auto prm = depot->get(model, params);
  if (!prm) throw Panic("called `Option::unwrap()` on a `None` value");  // elab.rs:144
  auto b = std::make_shared<Bsim4>();
  b->prm = prm;
  b->mps = s21::b4::matrix_pointers(prm->flavor);
  if (cplx) bsim4_make_ports(*b, path, ports, *(Variables<Cplx>*)vars_any);
  else bsim4_make_ports(*b, path, ports, *(Variables<double>*)vars_any);
  return b;
}

inline void bsim4_install_defs(Defs& defs, const std::map<std::string, std::pair<int, Specs>>& models, const std::map<std::string, Specs>& insts) {
  auto depot = std::make_shared<Bsim4Depot>();
  depot->models = models;
  depot->insts = insts;
  defs.bsim4 = depot;
  bsim4_elab_hook() = &bsim4_elaborate;
}

}  // namespace orc
