// ORACLE — TEST INFRASTRUCTURE ONLY (see num.hpp header).
//
// CPU restatement of the reference's solver core:
//   spice21/src/analysis.rs:23-31   Stamps
//   spice21/src/analysis.rs:44-123  Variables
//   spice21/src/analysis.rs:389-449 AnalysisInfo, TranState::integrate (Backward Euler; TRAP panics)
//   spice21/src/analysis.rs:642-693 Options
//   spice21/src/comps/mod.rs:74-93  trait Component (the device plugin surface)
#pragma once
#include <cmath>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "num.hpp"
#include "sparse21.hpp"

namespace orc {

// comps/mod.rs:24-37
namespace consts {
static const double PI = 3.14159265358979323846264338327950288;
static const double KB = 1.3806226e-23;
static const double Q = 1.6021918e-19;
static const double KB_OVER_Q = KB / Q;
static const double KELVIN_TO_C = 273.15;
static const double TEMP_REF = KELVIN_TO_C + 27.0;
static const double VT_REF = KB * TEMP_REF / Q;
static const double SIO2_PERMITTIVITY = 3.9 * 8.854214871e-12;
static const double SQRT2 = 1.4142135624;
static const double EPS0 = 8.85418e-12;
static const double EPSSI = 1.03594e-10;
}  // namespace consts

// Option<f64> stand-in
struct OptF {
  bool some = false;
  double v = 0.0;
  OptF() {}
  OptF(double x) : some(true), v(x) {}
  double or_(double d) const { return some ? v : d; }
};

// analysis.rs:642-693. Only temp/tnom/gmin/iabstol/reltol are settable (proto SimOptions).
struct Options {
  double temp = 300.15, tnom = 300.15, gmin = 1e-12, iabstol = 1e-12, reltol = 1e-3;
};

typedef int VarIndex;  // -1 == None (ground)
typedef int Eindex;    // -1 == None

enum class VarKind { V = 0, I, Q };

// analysis.rs:23-31
template <class T>
struct Stamps {
  std::vector<std::pair<Eindex, T>> g;
  std::vector<std::pair<VarIndex, T>> b;
};

// analysis.rs:44-123
template <class T>
struct Variables {
  std::vector<VarKind> kinds;
  std::vector<T> values;
  std::vector<std::string> names;
  // lookup accelerator with the same semantics as `names.iter().position(..)` (first match wins)
  std::unordered_map<std::string, int> first_index;

  template <class U>
  static Variables from(const Variables<U>& o) {  // :59-65
    Variables v;
    v.kinds = o.kinds;
    v.names = o.names;
    v.values.assign(o.values.size(), zero<T>());
    v.first_index = o.first_index;
    return v;
  }
  VarIndex add(const std::string& name, VarKind kind) {  // :67-73
    kinds.push_back(kind);
    names.push_back(name);
    values.push_back(zero<T>());
    int idx = (int)kinds.size() - 1;
    if (!first_index.count(name)) first_index[name] = idx;
    return idx;
  }
  VarIndex addv(const std::string& name) { return add(name, VarKind::V); }
  VarIndex addi(const std::string& name) { return add(name, VarKind::I); }
  VarIndex find(const std::string& name) const {  // :83-89
    auto it = first_index.find(name);
    return it == first_index.end() ? -1 : it->second;
  }
  // :92-111. NodeRef is represented by its to_string(): "" == Gnd, Num(n) == n.to_string().
  VarIndex find_or_create(const std::string& node) {
    if (node.empty()) return -1;
    VarIndex i = find(node);
    if (i >= 0) return i;
    return add(node, VarKind::V);
  }
  T get(VarIndex i) const { return i < 0 ? zero<T>() : values[(size_t)i]; }  // :114-119
  size_t len() const { return kinds.size(); }
};

// analysis.rs:395-449
struct TranState {
  double t = 0.0, dt = 0.0;
  std::vector<size_t> vic, ric;
  // Backward Euler only (NumericalIntegration::BE default; TRAP panics :429-435)
  void integrate(double dq, double dq_dv, double vguess, double /*_ip*/, double* g, double* i, double* rhs) const {
    double dt_ = dt;
    *g = dq_dv / dt_;
    *i = dq / dt_;
    *rhs = *i - *g * vguess;
  }
};
struct ChargeInteg { double g = 0.0, i = 0.0, rhs = 0.0; };
inline ChargeInteg integq(const TranState& s, double dq, double dq_dv, double vguess, double ip) {
  ChargeInteg c;
  s.integrate(dq, dq_dv, vguess, ip, &c.g, &c.i, &c.rhs);
  return c;
}
struct AcState { double omega = 0.0; };

// analysis.rs:389-393
struct AnalysisInfo {
  enum Kind { OP, TRAN, AC } kind = OP;
  const TranState* tran = nullptr;
  const AcState* ac = nullptr;
};

// comps/mod.rs:74-93 — the device plugin surface
struct Component {
  virtual ~Component() {}
  virtual void commit() {}
  virtual void update(double /*val*/) {}
  virtual Stamps<Cplx> load_ac(const Variables<Cplx>&, const AnalysisInfo&, const Options&) {
    throw Panic("AC Not Implemented For This Component!");
  }
  virtual Stamps<double> load(const Variables<double>&, const AnalysisInfo&, const Options&) = 0;
  virtual void create_matrix_elems(Matrix<double>& mat) = 0;
  virtual void create_matrix_elems(Matrix<Cplx>& mat) = 0;
  // instrumentation (not in the reference): the Eindex handles in create order, for stamp-map export
  virtual void matps_list(std::vector<Eindex>& /*out*/) const {}
  virtual const char* kind_name() const = 0;
};

// comps/mod.rs:349-354
template <class T>
inline Eindex make_matrix_elem(Matrix<T>& mat, VarIndex row, VarIndex col) {
  if (row >= 0 && col >= 0) return mat.make((size_t)row, (size_t)col);
  return -1;
}

}  // namespace orc
