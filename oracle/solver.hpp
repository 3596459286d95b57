// ORACLE — TEST INFRASTRUCTURE ONLY (see num.hpp header).
//
// CPU restatement of the reference's Newton solver and analyses:
//   spice21/src/analysis.rs:139-346  Solver<NumT>: update / solve (real: <=100 iters, complex: <=20) / converged
//   spice21/src/analysis.rs:383-388  dcop
//   spice21/src/analysis.rs:489-573  Tran::{new, ic, solve}  (fixed-step Backward Euler)
//   spice21/src/analysis.rs:761-832  ac (log sweep, end-inclusive; file side effects omitted)
#pragma once
#include <chrono>

#include "circuit.hpp"

namespace orc {

struct SolveStats {  // instrumentation only
  uint64_t loads = 0;     // passes through update()
  uint64_t solves = 0;    // passes that reached mat.solve()  == "Newton iterations" of the metric
  double seconds = 0.0;   // wall time spent inside Solver::solve
};

template <class T>
struct Solver {  // analysis.rs:139-147
  std::vector<std::shared_ptr<Component>> comps;
  Variables<T> vars;
  Matrix<T> mat;
  std::vector<T> rhs;
  Defs defs;
  Options opts;
  SolveStats stats;

  // analysis.rs:308-330
  static Solver make(const Ckt& ckt, const Options& opts) {
    Elaborator<T> e = elaborate<T>(ckt, opts);
    Solver s;
    s.defs = e.defs;
    s.comps = e.comps;
    s.vars = e.vars;
    s.opts = e.opts;
    for (auto& c : s.comps) c->create_matrix_elems(s.mat);
    return s;
  }
  bool converged(const std::vector<T>& dx, const std::vector<T>& res) const {  // :331-345
    for (auto& e : dx) if (absv(e) > opts.reltol) return false;
    for (auto& e : res) if (absv(e) > opts.iabstol) return false;
    return true;
  }
};

// Real update + solve (analysis.rs:151-211)
inline void update(Solver<double>& s, const AnalysisInfo& an) {
  for (auto& comp : s.comps) {
    Stamps<double> u = comp->load(s.vars, an, s.opts);
    for (auto& g : u.g) if (g.first >= 0) s.mat.update(g.first, g.second);
    for (auto& b : u.b) if (b.first >= 0) s.rhs.at((size_t)b.first) += b.second;
  }
  s.stats.loads++;
}
inline std::vector<double> solve(Solver<double>& s, const AnalysisInfo& an) {
  auto t0 = std::chrono::steady_clock::now();
  struct Timer {
    Solver<double>& s; std::chrono::steady_clock::time_point t0;
    ~Timer() { s.stats.seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
  } timer{s, t0};
  std::vector<double> dx(s.vars.len(), 0.0);
  for (int k = 0; k < 100; k++) {
    s.mat.reset();
    s.rhs.assign(s.vars.len(), 0.0);
    update(s, an);
    std::vector<double> res = s.mat.res(s.vars.values, s.rhs);
    if (s.converged(dx, res)) {
      for (auto& c : s.comps) c->commit();
      return s.vars.values;
    }
    dx = s.mat.solve(res);
    s.stats.solves++;
    double max_step = 1000e-3;
    double max_abs = 0.0;
    for (double v : dx) if (std::fabs(v) > max_abs) max_abs = std::fabs(v);
    if (max_abs > max_step) {
      for (size_t r = 0; r < dx.size(); r++) dx[r] = dx[r] * max_step / max_abs;
    }
    for (size_t r = 0; r < s.vars.len(); r++) s.vars.values[r] += dx.at(r);
  }
  throw SpError("Convergence Failed");
}

// Complex solver (analysis.rs:215-304)
inline Solver<Cplx> to_complex(Solver<double>& re) {  // :218-234
  Solver<Cplx> op;
  op.comps = re.comps;
  op.vars = Variables<Cplx>::from(re.vars);
  op.defs = re.defs;
  op.opts = re.opts;
  for (auto& c : op.comps) c->create_matrix_elems(op.mat);
  return op;
}
inline void update(Solver<Cplx>& s, const AnalysisInfo& an) {
  for (auto& comp : s.comps) {
    Stamps<Cplx> u = comp->load_ac(s.vars, an, s.opts);
    for (auto& g : u.g) if (g.first >= 0) s.mat.update(g.first, g.second);
    for (auto& b : u.b) if (b.first >= 0) s.rhs.at((size_t)b.first) = s.rhs.at((size_t)b.first) + b.second;
  }
  s.stats.loads++;
}
inline std::vector<Cplx> solve(Solver<Cplx>& s, const AnalysisInfo& an) {
  auto t0 = std::chrono::steady_clock::now();
  struct Timer {
    Solver<Cplx>& s; std::chrono::steady_clock::time_point t0;
    ~Timer() { s.stats.seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
  } timer{s, t0};
  std::vector<Cplx> dx(s.vars.len(), Cplx());
  for (int k = 0; k < 20; k++) {
    s.mat.reset();
    s.rhs.assign(s.vars.len(), Cplx());
    update(s, an);
    std::vector<Cplx> res = s.mat.res(s.vars.values, s.rhs);
    bool vtol = true, itol = true;
    for (auto& v : dx) if (!(absv(v) < 1e-3)) vtol = false;
    for (auto& v : res) if (!(absv(v) < 1e-9)) itol = false;
    if (vtol && itol) {
      for (auto& c : s.comps) c->commit();
      return s.vars.values;
    }
    dx = s.mat.solve(res);
    s.stats.solves++;
    double max_step = 1.0;
    double max_abs = 0.0;
    for (auto& v : dx) if (absv(v) > max_abs) max_abs = absv(v);
    if (max_abs > max_step) {
      for (size_t r = 0; r < dx.size(); r++) dx[r] = dx[r] * max_step / max_abs;
    }
    for (size_t r = 0; r < s.vars.len(); r++) s.vars.values[r] = s.vars.values[r] + dx.at(r);
  }
  throw SpError("Convergence Failed");
}

// ------------------------------------------------------------------ results
struct OpResult { std::vector<std::string> names; std::vector<double> values; };
struct TranResult {
  std::vector<std::string> signals;
  std::vector<double> time;
  std::vector<std::vector<double>> data;  // [timepoint][signal]
};
struct AcResult {
  std::vector<std::string> signals;
  std::vector<double> freq;
  std::vector<std::vector<Cplx>> data;  // [freq][signal]
};

// analysis.rs:383-388
inline OpResult dcop(Solver<double>& s) {
  AnalysisInfo an;
  an.kind = AnalysisInfo::OP;
  solve(s, an);
  return OpResult{s.vars.names, s.vars.values};
}

struct TranOptions {  // analysis.rs:451-456
  double tstep = 0.0, tstop = 0.0;
  std::vector<std::pair<std::string, double>> ic;
};

struct Tran {  // analysis.rs:489-574
  Solver<double> solver;
  TranState state;
  TranOptions opts;
  static Tran make(const Ckt& ckt, const Options& o, const TranOptions& args) {
    Tran t;
    t.solver = Solver<double>::make(ckt, o);
    t.opts = args;
    for (auto& nv : args.ic) t.ic(nv.first, nv.second);
    return t;
  }
  void ic(const std::string& n, double val) {  // :510-525
    VarIndex fnode = solver.vars.add("." + n + ".vic", VarKind::V);
    VarIndex ivar = solver.vars.add("." + n + ".iic", VarKind::I);
    auto r = std::make_shared<Resistor>(1.0, fnode, solver.vars.find_or_create(n));
    r->create_matrix_elems(solver.mat);
    solver.comps.push_back(r);
    state.ric.push_back(solver.comps.size() - 1);
    auto v = std::make_shared<Vsrc>(val, 0.0, fnode, -1, ivar);
    v->create_matrix_elems(solver.mat);
    solver.comps.push_back(v);
    state.vic.push_back(solver.comps.size() - 1);
  }
  // max_points: oracle-only bound so the CPU baseline can time a prefix of a long run (0 = unbounded)
  TranResult solve_(size_t max_points = 0) {  // :526-573
    TranResult results;
    results.signals = solver.vars.names;
    AnalysisInfo op;
    op.kind = AnalysisInfo::OP;
    std::vector<double> tdata = solve(solver, op);
    results.time.push_back(state.t);
    results.data.push_back(tdata);
    for (size_t c : state.vic) solver.comps[c]->update(0.0);
    for (size_t c : state.ric) solver.comps[c]->update(1e-9);
    size_t tpoint = 0;
    size_t max_tpoints = (size_t)1e9;
    state.t = opts.tstep;
    state.dt = opts.tstep;
    while (state.t < opts.tstop && tpoint < max_tpoints) {
      if (max_points && tpoint >= max_points) break;
      AnalysisInfo an;
      an.kind = AnalysisInfo::TRAN;
      an.tran = &state;
      tdata = solve(solver, an);
      results.time.push_back(state.t);
      results.data.push_back(tdata);
      tpoint += 1;
      state.t += opts.tstep;
    }
    return results;
  }
};

struct AcOptions { uint64_t fstart = 0, fstop = 0, npts = 0; };  // analysis.rs:701-713

// analysis.rs:761-832 (stream.ac.json / ac.json side effects omitted)
// max_points: oracle-only bound so tests can check a prefix of a long sweep (0 = unbounded)
inline AcResult ac(const Ckt& ckt, const Options& opts, const AcOptions& args, SolveStats* op_stats = nullptr,
                   SolveStats* ac_stats = nullptr, size_t max_points = 0) {
  Solver<double> re = Solver<double>::make(ckt, opts);
  AnalysisInfo op;
  op.kind = AnalysisInfo::OP;
  solve(re, op);
  if (op_stats) *op_stats = re.stats;
  Solver<Cplx> solver = to_complex(re);
  AcState state;
  AcResult results;
  results.signals = solver.vars.names;
  double f = (double)args.fstart;
  double fstop = (double)args.fstop;
  double fstep = std::pow(10.0, std::log10(fstop / f) / (double)args.npts);
  while (f <= fstop) {
    state.omega = 2.0 * consts::PI * f;
    AnalysisInfo an;
    an.kind = AnalysisInfo::AC;
    an.ac = &state;
    std::vector<Cplx> fsoln = solve(solver, an);
    results.freq.push_back(f);
    results.data.push_back(fsoln);
    if (f == fstop) break;
    if (max_points && results.freq.size() >= max_points) break;
    f = std::fmin(f * fstep, fstop);
  }
  if (ac_stats) *ac_stats = solver.stats;
  return results;
}

// ORACLE-ONLY extension (no counterpart in the reference): the body of the sweep loop above (analysis.rs:797-819) at
// caller-given frequencies, every point started cold — solver.vars zeroed as Variables::from leaves them (analysis.rs:59-65)
// — which is exactly what a one-point sweep `ac(fstart = fstop = f)` computes, without repeating elaboration and the OP.
// Lets the tests compare a strided subset of a long batched sweep (config C5) point by point.
inline AcResult ac_at(const Ckt& ckt, const Options& opts, const std::vector<double>& freqs, SolveStats* ac_stats = nullptr) {
  Solver<double> re = Solver<double>::make(ckt, opts);
  AnalysisInfo op;
  op.kind = AnalysisInfo::OP;
  solve(re, op);
  Solver<Cplx> solver = to_complex(re);
  AcState state;
  AcResult results;
  results.signals = solver.vars.names;
  for (double f : freqs) {
    for (auto& v : solver.vars.values) v = Cplx();
    state.omega = 2.0 * consts::PI * f;
    AnalysisInfo an;
    an.kind = AnalysisInfo::AC;
    an.ac = &state;
    std::vector<Cplx> fsoln = solve(solver, an);
    results.freq.push_back(f);
    results.data.push_back(fsoln);
  }
  if (ac_stats) *ac_stats = solver.stats;
  return results;
}

}  // namespace orc
