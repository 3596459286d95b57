// ORACLE — TEST INFRASTRUCTURE ONLY (see num.hpp header).
//
// C API over the CPU restatement so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can
// drive it through ctypes. Circuits arrive as a small line-oriented text netlist (tests/circuits.py writes the
// same circuit as protobuf bytes for the product and as this text for the oracle), which mirrors what
// Ckt::from_proto does (spice21/src/circuit.rs:281-330): models are *resolved* while the Ckt is built.
//
//   signal <name>
//   mos0 <model> <0|1>                      (defs.mos0, tests.rs:1443-1447)
//   mos1model <name> <0|1> [tpg=<i>] [k=v]  (Mos1Model::resolve)
//   mos1inst <name> [k=v]...                (Mos1InstanceParams::resolve)
//   diodemodel <name> [k=v]...   diodeinst <name> [area=v] [temp=v]
//   bsim4model <name> <0|1> [k=v]...  bsim4inst <name> [k=v]...
//   module <name> <port>... / msignal <name> / endmodule     (component lines in between belong to the module)
//   R|C|I <name> <p> <n> <val>     V <name> <p> <n> <dc> <acm>
//   D <name> <p> <n> <model> <params>     M <name> <model> <params> <d> <g> <s> <b>
//   X <name> <module> <port>=<node>...
// `~` is the empty string (ground).
#include <atomic>
#include <cstring>
#include <sstream>
#include <thread>

#include "solver.hpp"
#ifdef ORC_WITH_BSIM4
#include "bsim4.hpp"
#endif

using namespace orc;

namespace {

struct Mos1ModelSpec { Specs specs; int mos_type = 0; bool has_tpg = false; long tpg = 0; };
struct CktSpec {
  std::vector<std::string> signals;
  std::vector<CompDesc> comps;
  std::map<std::string, std::shared_ptr<ModuleDef>> modules;
  std::map<std::string, MosType> mos0;
  std::map<std::string, Mos1ModelSpec> mos1_models;
  std::map<std::string, Specs> mos1_insts;
  std::map<std::string, Specs> diode_models;
  std::map<std::string, Specs> diode_insts;
  std::map<std::string, std::pair<int, Specs>> bsim4_models;
  std::map<std::string, Specs> bsim4_insts;
};

std::string tok_node(const std::string& t) { return t == "~" ? std::string() : t; }

void parse_kv(std::istringstream& ls, Specs& s, Mos1ModelSpec* mm = nullptr) {
  std::string t;
  while (ls >> t) {
    size_t eq = t.find('=');
    if (eq == std::string::npos) throw SpError("bad k=v token: " + t);
    std::string k = t.substr(0, eq), v = t.substr(eq + 1);
    if (mm && k == "tpg") { mm->has_tpg = true; mm->tpg = std::strtol(v.c_str(), nullptr, 10); continue; }
    s.d[k] = std::strtod(v.c_str(), nullptr);
  }
}

bool parse_comp(const std::string& kw, std::istringstream& ls, CompDesc* c) {
  std::string a, b;
  if (kw == "R" || kw == "C" || kw == "I") {
    c->kind = kw == "R" ? CompDesc::R : kw == "C" ? CompDesc::C : CompDesc::I;
    std::string v;
    ls >> c->name >> a >> b >> v;
    c->p = tok_node(a); c->n = tok_node(b);
    c->val = std::strtod(v.c_str(), nullptr);
    return true;
  }
  if (kw == "V") {
    c->kind = CompDesc::V;
    std::string v, m;
    ls >> c->name >> a >> b >> v >> m;
    c->p = tok_node(a); c->n = tok_node(b);
    c->val = std::strtod(v.c_str(), nullptr);
    c->acm = std::strtod(m.c_str(), nullptr);
    return true;
  }
  if (kw == "D") {
    c->kind = CompDesc::D;
    ls >> c->name >> a >> b >> c->model >> c->params;
    c->p = tok_node(a); c->n = tok_node(b);
    c->model = tok_node(c->model); c->params = tok_node(c->params);
    return true;
  }
  if (kw == "M") {
    c->kind = CompDesc::MOS;
    std::string d, g, s, bb;
    ls >> c->name >> c->model >> c->params >> d >> g >> s >> bb;
    c->model = tok_node(c->model); c->params = tok_node(c->params);
    c->d = tok_node(d); c->g = tok_node(g); c->s = tok_node(s); c->b = tok_node(bb);
    return true;
  }
  if (kw == "X") {
    c->kind = CompDesc::MODULE;
    ls >> c->name >> c->module;
    std::string t;
    while (ls >> t) {
      size_t eq = t.find('=');
      if (eq == std::string::npos) throw SpError("bad port token: " + t);
      c->ports.push_back({t.substr(0, eq), tok_node(t.substr(eq + 1))});
    }
    return true;
  }
  return false;
}

CktSpec parse_text(const char* text) {
  CktSpec cs;
  std::istringstream in(text);
  std::string line;
  std::shared_ptr<ModuleDef> cur;
  while (std::getline(in, line)) {
    std::istringstream ls(line);
    std::string kw;
    if (!(ls >> kw) || kw[0] == '#') continue;
    if (kw == "signal") { std::string n; ls >> n; cs.signals.push_back(n); }
    else if (kw == "mos0") { std::string n; int t; ls >> n >> t; cs.mos0[n] = t == 1 ? MosType::PMOS : MosType::NMOS; }
    else if (kw == "mos1model") { std::string n; Mos1ModelSpec m; ls >> n >> m.mos_type; parse_kv(ls, m.specs, &m); cs.mos1_models[n] = m; }
    else if (kw == "mos1inst") { std::string n; Specs s; ls >> n; n = tok_node(n); parse_kv(ls, s); cs.mos1_insts[n] = s; }
    else if (kw == "diodemodel") { std::string n; Specs s; ls >> n; parse_kv(ls, s); cs.diode_models[n] = s; }
    else if (kw == "diodeinst") { std::string n; Specs s; ls >> n; parse_kv(ls, s); cs.diode_insts[n] = s; }
    else if (kw == "bsim4model") { std::string n; int t; Specs s; ls >> n >> t; parse_kv(ls, s); cs.bsim4_models[n] = {t, s}; }
    else if (kw == "bsim4inst") { std::string n; Specs s; ls >> n; n = tok_node(n); parse_kv(ls, s); cs.bsim4_insts[n] = s; }
    else if (kw == "module") {
      cur = std::make_shared<ModuleDef>();
      ls >> cur->name;
      std::string p;
      while (ls >> p) cur->ports.push_back(p);
    } else if (kw == "msignal") { std::string n; ls >> n; if (!cur) throw SpError("msignal outside module"); cur->signals.push_back(n); }
    else if (kw == "endmodule") { cs.modules[cur->name] = cur; cur.reset(); }
    else {
      CompDesc c;
      if (!parse_comp(kw, ls, &c)) throw SpError("unknown netlist keyword: " + kw);
      if (cur) cur->comps.push_back(c); else cs.comps.push_back(c);
    }
  }
  return cs;
}

// One per-instance override: "<kind>:<name>:<param>" with kind in
//   mos1model, mos1inst, diodemodel, bsim4model, bsim4inst, R, C, I, V (param g|c|dc), opt (name ignored; temp|tnom|gmin|...)
struct Override { std::string kind, name, param; };
Override parse_override(const char* s) {
  std::string t(s);
  size_t a = t.find(':'), b = t.find(':', a + 1);
  if (a == std::string::npos || b == std::string::npos) throw SpError("bad override spec: " + t);
  return Override{t.substr(0, a), t.substr(a + 1, b - a - 1), t.substr(b + 1)};
}

// Ckt::from_proto equivalent (circuit.rs:281-330): resolve models, collect comps.
Ckt build_ckt(const CktSpec& cs0, const std::vector<Override>& ovr, const double* vals /*[n_over]*/, Options* opts) {
  CktSpec cs = cs0;
  for (size_t k = 0; k < ovr.size(); k++) {
    const Override& o = ovr[k];
    double v = vals[k];
    if (o.kind == "mos1model") cs.mos1_models.at(o.name).specs.d[o.param] = v;
    else if (o.kind == "mos1inst") cs.mos1_insts.at(o.name).d[o.param] = v;
    else if (o.kind == "diodemodel") cs.diode_models.at(o.name).d[o.param] = v;
    else if (o.kind == "bsim4model") cs.bsim4_models.at(o.name).second.d[o.param] = v;
    else if (o.kind == "bsim4inst") cs.bsim4_insts.at(o.name).d[o.param] = v;
    else if (o.kind == "opt") {
      if (o.param == "temp") opts->temp = v;
      else if (o.param == "tnom") opts->tnom = v;
      else if (o.param == "gmin") opts->gmin = v;
      else if (o.param == "iabstol") opts->iabstol = v;
      else if (o.param == "reltol") opts->reltol = v;
      else throw SpError("bad opt override");
    } else {
      bool found = false;
      for (auto& c : cs.comps) {
        if (c.name != o.name) continue;
        if ((o.kind == "R" && c.kind == CompDesc::R) || (o.kind == "C" && c.kind == CompDesc::C) ||
            (o.kind == "I" && c.kind == CompDesc::I) || (o.kind == "V" && c.kind == CompDesc::V)) {
          if (o.kind == "V" && o.param == "acm") c.acm = v; else c.val = v;
          found = true;
        }
      }
      if (!found) throw SpError("override target not found: " + o.kind + ":" + o.name);
    }
  }
  Ckt ckt;
  ckt.signals = cs.signals;
  ckt.comps = cs.comps;
  ckt.defs.modules = cs.modules;
  ckt.defs.mos0 = cs.mos0;
  for (auto& kv : cs.mos1_models)
    ckt.defs.mos1_models[kv.first] =
        std::make_shared<Mos1Model>(Mos1Model::resolve(kv.second.specs, kv.second.mos_type, kv.second.has_tpg, kv.second.tpg));
  for (auto& kv : cs.mos1_insts) ckt.defs.mos1_insts[kv.first] = std::make_shared<Mos1InstanceParams>(Mos1InstanceParams::resolve(kv.second));
  for (auto& kv : cs.diode_models) ckt.defs.diode_models[kv.first] = std::make_shared<DiodeModel>(DiodeModel::from(kv.second.d));
  for (auto& kv : cs.diode_insts) {
    auto p = std::make_shared<DiodeInstParams>();
    p->area = kv.second.get("area");
    p->temp = kv.second.get("temp");
    ckt.defs.diode_insts[kv.first] = p;
  }
#ifdef ORC_WITH_BSIM4
  bsim4_install_defs(ckt.defs, cs.bsim4_models, cs.bsim4_insts);
#else
  if (!cs.bsim4_models.empty()) throw SpError("oracle built without BSIM4");
#endif
  return ckt;
}

Options make_opts(const double* o5) {
  Options o;
  if (o5) {
    if (!std::isnan(o5[0])) o.temp = o5[0];
    if (!std::isnan(o5[1])) o.tnom = o5[1];
    if (!std::isnan(o5[2])) o.gmin = o5[2];
    if (!std::isnan(o5[3])) o.iabstol = o5[3];
    if (!std::isnan(o5[4])) o.reltol = o5[4];
  }
  return o;
}

struct Result {
  std::string names;   // '\n'-joined
  int nsig = 0, npts = 0, width = 1;  // width 2 for complex
  std::vector<double> axis, data;
  uint64_t loads = 0, solves = 0, factorizations = 0;
  double seconds = 0.0;
};
void set_names(Result* r, const std::vector<std::string>& names) {
  r->nsig = (int)names.size();
  for (size_t k = 0; k < names.size(); k++) { if (k) r->names += "\n"; r->names += names[k]; }
}

struct Structure {
  int n_vars = 0;
  std::string names;
  std::vector<int> elem_row, elem_col;        // creation order, original coordinates
  std::vector<int> comp_off, comp_matps;      // CSR over components: Eindex handles in create order (-1 = ground)
  std::string comp_kinds;                     // '\n'-joined
  // after the first factorisation (first Newton iteration from x = 0):
  std::vector<int> row_i2e, col_i2e;          // pivot order
  std::vector<int> lu_row, lu_col, lu_fill;   // every element incl. fill-ins, internal coords + fill flag, creation order
  std::vector<double> a0;                     // assembled values of the first load sweep, by element id
};

int fail(int code, const std::exception& e, char* err, int errlen) {
  if (err && errlen > 0) { std::strncpy(err, e.what(), (size_t)errlen - 1); err[errlen - 1] = 0; }
  return code;
}
// status codes shared with include/spice21cu.h
enum { ST_OK = 0, ST_CONV = 1, ST_SINGULAR = 2, ST_PIVOT = 3, ST_DECODE = 4, ST_INVALID = 5, ST_UNSUPPORTED = 6, ST_OTHER = 8 };
int classify(const std::exception& e) {
  std::string w = e.what();
  if (w == "Convergence Failed") return ST_CONV;
  if (w == "Singular Matrix") return ST_SINGULAR;
  if (w.rfind("Assert Neq Failed", 0) == 0) return ST_SINGULAR;  // zero pivot, reported through the assert helper (mod.rs:872)
  if (w == "Pivot Search Fail") return ST_PIVOT;
  if (w.find("AC Not Implemented") != std::string::npos) return ST_UNSUPPORTED;
  if (dynamic_cast<const Panic*>(&e)) return ST_INVALID;
  return ST_OTHER;
}

}  // namespace

extern "C" {

void* orc_ckt_parse(const char* text, char* err, int errlen) {
  try {
    return new CktSpec(parse_text(text));
  } catch (const std::exception& e) { fail(0, e, err, errlen); return nullptr; }
}
void orc_ckt_free(void* c) { delete (CktSpec*)c; }

int orc_run_op(void* ckt, const double* opts5, void** out, char* err, int errlen) {
  try {
    Options o = make_opts(opts5);
    Ckt c = build_ckt(*(CktSpec*)ckt, {}, nullptr, &o);
    Solver<double> s = Solver<double>::make(c, o);
    OpResult r = dcop(s);
    auto* res = new Result();
    set_names(res, r.names);
    res->npts = 1;
    res->data = r.values;
    res->loads = s.stats.loads; res->solves = s.stats.solves; res->factorizations = s.mat.n_factorizations; res->seconds = s.stats.seconds;
    *out = res;
    return ST_OK;
  } catch (const std::exception& e) { return fail(classify(e), e, err, errlen); }
}

int orc_run_tran(void* ckt, const double* opts5, double tstep, double tstop, int n_ic, const char** ic_nodes, const double* ic_vals,
                 long max_points, void** out, char* err, int errlen) {
  try {
    Options o = make_opts(opts5);
    Ckt c = build_ckt(*(CktSpec*)ckt, {}, nullptr, &o);
    TranOptions to;
    to.tstep = tstep; to.tstop = tstop;
    for (int k = 0; k < n_ic; k++) to.ic.push_back({tok_node(ic_nodes[k]), ic_vals[k]});
    Tran t = Tran::make(c, o, to);
    TranResult r = t.solve_((size_t)max_points);
    auto* res = new Result();
    set_names(res, r.signals);
    res->npts = (int)r.time.size();
    res->axis = r.time;
    for (auto& row : r.data) res->data.insert(res->data.end(), row.begin(), row.end());
    res->loads = t.solver.stats.loads; res->solves = t.solver.stats.solves; res->factorizations = t.solver.mat.n_factorizations;
    res->seconds = t.solver.stats.seconds;
    *out = res;
    return ST_OK;
  } catch (const std::exception& e) { return fail(classify(e), e, err, errlen); }
}

int orc_run_ac(void* ckt, const double* opts5, unsigned long long fstart, unsigned long long fstop, unsigned long long npts, long max_points,
               void** out, char* err, int errlen) {
  try {
    Options o = make_opts(opts5);
    Ckt c = build_ckt(*(CktSpec*)ckt, {}, nullptr, &o);
    AcOptions a;
    a.fstart = fstart; a.fstop = fstop; a.npts = npts;
    SolveStats st_op, st_ac;
    AcResult r = ac(c, o, a, &st_op, &st_ac, (size_t)max_points);
    auto* res = new Result();
    set_names(res, r.signals);
    res->npts = (int)r.freq.size();
    res->width = 2;
    res->axis = r.freq;
    for (auto& row : r.data) for (auto& z : row) { res->data.push_back(z.re); res->data.push_back(z.im); }
    res->loads = st_ac.loads; res->solves = st_ac.solves; res->seconds = st_ac.seconds;
    *out = res;
    return ST_OK;
  } catch (const std::exception& e) { return fail(classify(e), e, err, errlen); }
}

int orc_run_ac_at(void* ckt, const double* opts5, const double* freqs, long n, void** out, char* err, int errlen) {
  try {
    Options o = make_opts(opts5);
    Ckt c = build_ckt(*(CktSpec*)ckt, {}, nullptr, &o);
    SolveStats st_ac;
    AcResult r = ac_at(c, o, std::vector<double>(freqs, freqs + n), &st_ac);
    auto* res = new Result();
    set_names(res, r.signals);
    res->npts = (int)r.freq.size();
    res->width = 2;
    res->axis = r.freq;
    for (auto& row : r.data) for (auto& z : row) { res->data.push_back(z.re); res->data.push_back(z.im); }
    res->loads = st_ac.loads; res->solves = st_ac.solves; res->seconds = st_ac.seconds;
    *out = res;
    return ST_OK;
  } catch (const std::exception& e) { return fail(classify(e), e, err, errlen); }
}

int orc_res_nsig(void* r) { return ((Result*)r)->nsig; }
int orc_res_npts(void* r) { return ((Result*)r)->npts; }
int orc_res_width(void* r) { return ((Result*)r)->width; }
const char* orc_res_names(void* r) { return ((Result*)r)->names.c_str(); }
void orc_res_data(void* r, double* out) { auto* x = (Result*)r; std::memcpy(out, x->data.data(), x->data.size() * sizeof(double)); }
void orc_res_axis(void* r, double* out) { auto* x = (Result*)r; std::memcpy(out, x->axis.data(), x->axis.size() * sizeof(double)); }
void orc_res_stats(void* r, double* out4) {
  auto* x = (Result*)r;
  out4[0] = (double)x->loads; out4[1] = (double)x->solves; out4[2] = (double)x->factorizations; out4[3] = x->seconds;
}
void orc_res_free(void* r) { delete (Result*)r; }

// Structure export: variable numbering, element creation order (== stamp map), and the pivot order / fill pattern
// produced by the reference algorithm on the first Newton iteration of the OP solve (x = 0).
int orc_structure(void* ckt, const double* opts5, int n_ic, const char** ic_nodes, const double* ic_vals, void** out, char* err, int errlen) {
  try {
    Options o = make_opts(opts5);
    Ckt c = build_ckt(*(CktSpec*)ckt, {}, nullptr, &o);
    TranOptions to;
    for (int k = 0; k < n_ic; k++) to.ic.push_back({tok_node(ic_nodes[k]), ic_vals[k]});
    Tran t = Tran::make(c, o, to);
    Solver<double>& s = t.solver;
    auto* st = new Structure();
    st->n_vars = (int)s.vars.len();
    for (size_t k = 0; k < s.vars.names.size(); k++) { if (k) st->names += "\n"; st->names += s.vars.names[k]; }
    for (auto& e : s.mat.elements) { st->elem_row.push_back((int)e.orig_row); st->elem_col.push_back((int)e.orig_col); }
    st->comp_off.push_back(0);
    for (size_t k = 0; k < s.comps.size(); k++) {
      std::vector<Eindex> m;
      s.comps[k]->matps_list(m);
      st->comp_matps.insert(st->comp_matps.end(), m.begin(), m.end());
      st->comp_off.push_back((int)st->comp_matps.size());
      if (k) st->comp_kinds += "\n";
      st->comp_kinds += s.comps[k]->kind_name();
    }
    // first Newton iteration, exactly as Solver::solve does it
    AnalysisInfo an;
    an.kind = AnalysisInfo::OP;
    s.mat.reset();
    s.rhs.assign(s.vars.len(), 0.0);
    update(s, an);
    std::vector<double> res = s.mat.res(s.vars.values, s.rhs);
    for (auto& e : s.mat.elements) st->a0.push_back(e.val);
    try {
      s.mat.solve(res);
      for (size_t k = 0; k < s.mat.axes[ROWS].mapping.i2e.size(); k++) st->row_i2e.push_back((int)s.mat.axes[ROWS].mapping.i2e[k]);
      for (size_t k = 0; k < s.mat.axes[COLS].mapping.i2e.size(); k++) st->col_i2e.push_back((int)s.mat.axes[COLS].mapping.i2e[k]);
      for (auto& e : s.mat.elements) { st->lu_row.push_back((int)e.row); st->lu_col.push_back((int)e.col); st->lu_fill.push_back(e.fillin ? 1 : 0); }
    } catch (const SpError&) {
      // singular at x = 0: structure still exported, pivot arrays left empty
    }
    *out = st;
    return ST_OK;
  } catch (const std::exception& e) { return fail(classify(e), e, err, errlen); }
}
int orc_st_nvars(void* s) { return ((Structure*)s)->n_vars; }
const char* orc_st_names(void* s) { return ((Structure*)s)->names.c_str(); }
const char* orc_st_comp_kinds(void* s) { return ((Structure*)s)->comp_kinds.c_str(); }
static int copy_ivec(const std::vector<int>& v, int* out, int cap) {
  if (out) for (size_t k = 0; k < v.size() && (int)k < cap; k++) out[k] = v[k];
  return (int)v.size();
}
int orc_st_vec(void* s, int which, int* out, int cap) {
  auto* st = (Structure*)s;
  switch (which) {
    case 0: return copy_ivec(st->elem_row, out, cap);
    case 1: return copy_ivec(st->elem_col, out, cap);
    case 2: return copy_ivec(st->comp_off, out, cap);
    case 3: return copy_ivec(st->comp_matps, out, cap);
    case 4: return copy_ivec(st->row_i2e, out, cap);
    case 5: return copy_ivec(st->col_i2e, out, cap);
    case 6: return copy_ivec(st->lu_row, out, cap);
    case 7: return copy_ivec(st->lu_col, out, cap);
    case 8: return copy_ivec(st->lu_fill, out, cap);
  }
  return -1;
}
int orc_st_a0(void* s, double* out, int cap) {
  auto* st = (Structure*)s;
  if (out) for (size_t k = 0; k < st->a0.size() && (int)k < cap; k++) out[k] = st->a0[k];
  return (int)st->a0.size();
}
void orc_st_free(void* s) { delete (Structure*)s; }

#if defined(ORC_WITH_BSIM4) && defined(S21_B4_COUNT)
// instrumented build only (make liboracle_count.so; scripts/b4_opcount.py): executed operations of the Bsim4 evaluation
void orc_b4_counts(unsigned long long* out6, int reset) {
  auto& c = s21::b4e::b4_counts();
  out6[0] = c.evals; out6[1] = c.div; out6[2] = c.div_special; out6[3] = c.exp; out6[4] = c.log; out6[5] = c.sqrt;
  if (reset) c = {0, 0, 0, 0, 0, 0};
}
#endif

// Factorise an arbitrary matrix with the restated sparse21 and report the pivot order and the L+U pattern
// (internal coordinates, creation order). width 1 = real vals[nnz], 2 = complex vals[nnz][2].
// Returns the status of lu_factorize (0 OK / 2 singular / 3 pivot fail); outputs sized by the caller:
// row_i2e[n], col_i2e[n], lu_row/lu_col/lu_fill[cap]; *nnz_lu receives the element count incl. fill-ins.
}  // extern "C"
template <class T>
static int lu_order_impl(int n, int nnz, const int* rows, const int* cols, const T* vals, int* row_i2e, int* col_i2e, int* lu_row,
                         int* lu_col, int* lu_fill, int cap, int* nnz_lu) {
  Matrix<T> m;
  for (int k = 0; k < nnz; k++) m.add_element((size_t)rows[k], (size_t)cols[k], vals[k]);
  int st = ST_OK;
  try {
    m.lu_factorize();
  } catch (const std::exception& e) { st = classify(e); }
  if (m.axes[ROWS].has_mapping)
    for (int k = 0; k < n && k < (int)m.axes[ROWS].mapping.i2e.size(); k++) row_i2e[k] = (int)m.axes[ROWS].mapping.i2e[(size_t)k];
  if (m.axes[COLS].has_mapping)
    for (int k = 0; k < n && k < (int)m.axes[COLS].mapping.i2e.size(); k++) col_i2e[k] = (int)m.axes[COLS].mapping.i2e[(size_t)k];
  *nnz_lu = (int)m.elements.size();
  for (int k = 0; k < (int)m.elements.size() && k < cap; k++) {
    lu_row[k] = (int)m.elements[(size_t)k].row; lu_col[k] = (int)m.elements[(size_t)k].col; lu_fill[k] = m.elements[(size_t)k].fillin ? 1 : 0;
  }
  if (st == ST_OK && (int)m.diag.size() == n && n > 0 && m.diag[(size_t)n - 1] < 0) st = ST_SINGULAR;  // surfaces in solve() (mod.rs:969-972)
  return st;
}
extern "C" {
int orc_lu_order(int n, int nnz, const int* rows, const int* cols, const double* vals, int width, int* row_i2e, int* col_i2e, int* lu_row,
                 int* lu_col, int* lu_fill, int cap, int* nnz_lu) {
  if (width == 1) return lu_order_impl<double>(n, nnz, rows, cols, vals, row_i2e, col_i2e, lu_row, lu_col, lu_fill, cap, nnz_lu);
  std::vector<Cplx> z((size_t)nnz);
  for (int k = 0; k < nnz; k++) z[(size_t)k] = Cplx(vals[2 * k], vals[2 * k + 1]);
  return lu_order_impl<Cplx>(n, nnz, rows, cols, z.data(), row_i2e, col_i2e, lu_row, lu_col, lu_fill, cap, nnz_lu);
}

// Batched runs for the CPU baseline: B independent instances of one circuit with per-instance overrides.
// kind 0 = dcop, 1 = tran. Solvers are built first (untimed); then `nthreads` std::threads pull instances off an
// atomic counter and run Solver::solve; *seconds_out is the wall time of that parallel region only.
// x_out: dcop [B][N]; tran [B][T][N] when non-null (T = number of time points incl. t=0), iters_out [B] = solves.
int orc_batch_run(void* ckt, int kind, const double* opts5, int B, int n_over, const char** over_specs, const double* over_vals /*[n_over][B]*/,
                  double tstep, double tstop, int n_ic, const char** ic_nodes, const double* ic_vals, long max_points, int nthreads,
                  double* x_out, long long* iters_out, int* status_out, double* seconds_out, int* n_vars_out, int* n_pts_out, char* err,
                  int errlen) {
  try {
    const CktSpec& cs = *(CktSpec*)ckt;
    std::vector<Override> ovr;
    for (int k = 0; k < n_over; k++) ovr.push_back(parse_override(over_specs[k]));
    std::vector<std::unique_ptr<Tran>> runs((size_t)B);
    std::atomic<int> next(0);
    std::atomic<int> build_fail(0);
    std::string build_err;
    auto nt = (size_t)std::max(1, nthreads);
    {
      std::vector<std::thread> th;
      for (size_t t = 0; t < nt; t++)
        th.emplace_back([&]() {
          for (;;) {
            int i = next.fetch_add(1);
            if (i >= B) break;
            try {
              Options o = make_opts(opts5);
              std::vector<double> v((size_t)n_over);
              for (int k = 0; k < n_over; k++) v[(size_t)k] = over_vals[(size_t)k * (size_t)B + (size_t)i];
              Ckt c = build_ckt(cs, ovr, v.data(), &o);
              TranOptions to;
              to.tstep = tstep; to.tstop = tstop;
              if (kind == 1) for (int k = 0; k < n_ic; k++) to.ic.push_back({tok_node(ic_nodes[k]), ic_vals[k]});
              runs[(size_t)i].reset(new Tran(Tran::make(c, o, to)));
            } catch (const std::exception& e) {
              if (build_fail.fetch_add(1) == 0) build_err = e.what();
            }
          }
        });
      for (auto& t : th) t.join();
    }
    if (build_fail.load()) throw Panic("batch build failed: " + build_err);
    int N = (int)runs[0]->solver.vars.len();
    if (n_vars_out) *n_vars_out = N;
    std::vector<int> npts((size_t)B, 1);
    next = 0;
    auto t0 = std::chrono::steady_clock::now();
    {
      std::vector<std::thread> th;
      for (size_t t = 0; t < nt; t++)
        th.emplace_back([&]() {
          for (;;) {
            int i = next.fetch_add(1);
            if (i >= B) break;
            Tran& r = *runs[(size_t)i];
            int st = ST_OK;
            try {
              if (kind == 0) {
                OpResult o = dcop(r.solver);
                if (x_out) std::memcpy(x_out + (size_t)i * (size_t)N, o.values.data(), sizeof(double) * (size_t)N);
              } else {
                TranResult tr = r.solve_((size_t)max_points);
                npts[(size_t)i] = (int)tr.time.size();
                if (x_out) {
                  size_t T = tr.time.size();
                  for (size_t p = 0; p < T; p++)
                    std::memcpy(x_out + ((size_t)i * T + p) * (size_t)N, tr.data[p].data(), sizeof(double) * (size_t)N);
                }
              }
            } catch (const std::exception& e) { st = classify(e); }
            if (status_out) status_out[i] = st;
            if (iters_out) iters_out[i] = (long long)r.solver.stats.solves;
          }
        });
      for (auto& t : th) t.join();
    }
    if (seconds_out) *seconds_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (n_pts_out) *n_pts_out = npts[0];
    return ST_OK;
  } catch (const std::exception& e) { return fail(classify(e), e, err, errlen); }
}

// ------------------------------------------------------------------ sparse21 known-answer tests
// Restates the reference's own unit tests (sparse21/mod.rs:1165-1526). Returns 0 when all pass, else the
// 1-based number of the first failing check; `msg` names it.
static int kat_fail(int n, const char* what, char* msg, int len) {
  if (msg && len > 0) { std::snprintf(msg, (size_t)len, "check %d failed: %s", n, what); }
  return n;
}
int orc_sparse21_selftest(char* msg, int len) {
  int n = 0;
#define CHECK(cond) do { n++; if (!(cond)) return kat_fail(n, #cond, msg, len); } while (0)
  try {
    auto getv = [](Matrix<double>& m, size_t r, size_t c) { double v = NAN; m.get(r, c, &v); return v; };
    {  // test_add_element :1188
      Matrix<double> m;
      m.add_element(0, 0, 1.0);
      CHECK(m.num_rows() == 1 && m.num_cols() == 1 && m.diag.size() == 1);
      m.add_element(100, 100, 1.0);
      CHECK(m.num_rows() == 101 && m.num_cols() == 101 && m.diag.size() == 101);
    }
    for (size_t k = 1; k < 10; k++) {  // test_identity :1214
      Matrix<double> ik = Matrix<double>::identity(k);
      CHECK(ik.num_rows() == k && ik.num_cols() == k && ik.elements.size() == k);
      CHECK(ik.checkups().empty());
      for (size_t v = 0; v < k; v++) CHECK(ik.hdr(ROWS, v) == ik.hdr(COLS, v) && ik.hdr(ROWS, v) >= 0);
    }
    {  // test_swap_rows0 :1240
      Matrix<double> m;
      m.add_element(0, 0, 11.0); m.add_element(7, 0, 22.0); m.add_element(0, 7, 33.0); m.add_element(7, 7, 44.0);
      CHECK(m.checkups().empty());
      m.state = MatrixState::FACTORING;
      m.swap(ROWS, 0, 7);
      CHECK(m.checkups().empty());
      CHECK(getv(m, 7, 0) == 11.0 && getv(m, 0, 0) == 22.0 && getv(m, 7, 7) == 33.0 && getv(m, 0, 7) == 44.0);
    }
    {  // test_swap_rows1 :1266
      Matrix<double> m;
      m.add_element(0, 0, 11.1); m.add_element(2, 2, 22.2);
      CHECK(m.checkups().empty());
      m.state = MatrixState::FACTORING;
      m.swap(ROWS, 0, 2);
      CHECK(m.checkups().empty());
      double tmp;
      CHECK(getv(m, 2, 0) == 11.1 && getv(m, 0, 2) == 22.2 && !m.get(1, 1, &tmp));
    }
    {  // test_swap_rows2 :1288
      Matrix<double> m;
      double v = 1.0;
      for (size_t r = 0; r < 3; r++) for (size_t c = 0; c < 3; c++) m.add_element(r, c, v++);
      CHECK(m.checkups().empty());
      m.state = MatrixState::FACTORING;
      m.swap(ROWS, 0, 2);
      CHECK(m.checkups().empty());
      CHECK(getv(m, 0, 0) == 7.0 && getv(m, 2, 0) == 1.0);
    }
    {  // test_swap_rows3 :1313
      Matrix<double> m;
      m.add_element(1, 0, 71.0); m.add_element(2, 0, -11.0); m.add_element(2, 2, 99.0);
      CHECK(m.checkups().empty());
      m.state = MatrixState::FACTORING;
      m.swap(ROWS, 0, 2);
      CHECK(m.checkups().empty());
      CHECK(getv(m, 1, 0) == 71.0 && getv(m, 0, 0) == -11.0 && getv(m, 0, 2) == 99.0);
    }
    {  // test_swap_rows4 :1335
      Matrix<double> m;
      for (size_t r = 0; r < 3; r++) for (size_t c = 0; c < 3; c++) if (r != 0 || c != 1) m.add_element(r, c, (double)((r + 1) * (c + 1)));
      CHECK(m.checkups().empty());
      m.state = MatrixState::FACTORING;
      m.swap(ROWS, 0, 1);
      CHECK(m.checkups().empty());
    }
    {  // test_row_mappings :1357
      Matrix<double> m = Matrix<double>::identity(4);
      m.state = MatrixState::FACTORING;
      m.axes[ROWS].setup_factoring();
      m.swap(ROWS, 0, 3);
      CHECK(m.checkups().empty());
      CHECK((m.axes[ROWS].mapping.e2i == std::vector<size_t>{3, 1, 2, 0}));
      CHECK((m.axes[ROWS].mapping.i2e == std::vector<size_t>{3, 1, 2, 0}));
      m.swap(ROWS, 0, 2);
      CHECK(m.checkups().empty());
      CHECK((m.axes[ROWS].mapping.e2i == std::vector<size_t>{3, 1, 0, 2}));
      CHECK((m.axes[ROWS].mapping.i2e == std::vector<size_t>{2, 1, 3, 0}));
    }
    {  // test_lu_id3 :1378
      Matrix<double> m = Matrix<double>::identity(3);
      m.lu_factorize();
      CHECK(m.checkups().empty());
      CHECK(getv(m, 0, 0) == 1.0 && getv(m, 1, 1) == 1.0 && getv(m, 2, 2) == 1.0);
    }
    {  // test_lu_lower :1390
      Matrix<double> m;
      m.add_element(0, 0, 1.0); m.add_element(1, 0, 1.0); m.add_element(2, 0, 1.0);
      m.add_element(1, 1, 1.0); m.add_element(2, 1, 1.0); m.add_element(2, 2, 1.0);
      m.lu_factorize();
      CHECK(m.checkups().empty());
      CHECK(getv(m, 0, 0) == 1.0 && getv(m, 1, 0) == 1.0 && getv(m, 2, 0) == 1.0 && getv(m, 1, 1) == 1.0 && getv(m, 2, 1) == 1.0 &&
            getv(m, 2, 2) == 1.0);
    }
    {  // test_lu :1422
      Matrix<double> m = Matrix<double>::from_entries({{2, 2, -1.0}, {2, 1, 5.0}, {2, 0, 2.0}, {1, 2, 5.0}, {1, 1, 2.0}, {0, 2, 1.0}, {0, 1, 1.0}, {0, 0, 1.0}});
      CHECK(m.checkups().empty());
      m.lu_factorize();
      CHECK(m.checkups().empty());
    }
    {  // test_solve :1450
      Matrix<double> m = Matrix<double>::from_entries({{0, 0, 1.0}, {0, 1, 1.0}, {0, 2, 1.0}, {1, 1, 2.0}, {1, 2, 5.0}, {2, 0, 2.0}, {2, 1, 5.0}, {2, 2, -1.0}});
      m.lu_factorize();
      CHECK(m.checkups().empty());
      std::vector<double> soln = m.solve({6.0, -4.0, 27.0});
      const double correct[3] = {5.0, 3.0, -2.0};
      for (int k = 0; k < 3; k++) CHECK(std::fabs(soln[(size_t)k] - correct[k]) < 1e-9);
    }
    {  // test_solve_id3 :1474
      Matrix<double> m = Matrix<double>::identity(3);
      CHECK((m.solve({11.1, 30.3, 99.9}) == std::vector<double>{11.1, 30.3, 99.9}));
    }
    for (size_t s = 1; s < 10; s++) {  // test_solve_identity :1482
      Matrix<double> m = Matrix<double>::identity(s);
      std::vector<double> rhs;
      for (size_t e = 0; e < s; e++) rhs.push_back((double)e);
      CHECK(m.solve(rhs) == rhs);
    }
    {  // test_solve_complex_id2 :1508
      Matrix<Cplx> m = Matrix<Cplx>::from_entries({{0, 0, Cplx(1, 0)}, {1, 1, Cplx(1, 0)}});
      std::vector<Cplx> s = m.solve({Cplx(0, 1), Cplx(0, 1)});
      CHECK(s[0] == Cplx(0, 1) && s[1] == Cplx(0, 1));
    }
    {  // test_solve_complex :1516
      Matrix<Cplx> m = Matrix<Cplx>::from_entries({{0, 0, Cplx(1, 0)}, {1, 0, Cplx(-1, 0)}, {0, 1, Cplx(-1, 0)}, {1, 1, Cplx(1, 1)}});
      std::vector<Cplx> s = m.solve({Cplx(1, 0), Cplx(0, 0)});
      CHECK(s[0] == Cplx(1.0, -1.0) && s[1] == Cplx(0.0, -1.0));
    }
  } catch (const std::exception& e) {
    return kat_fail(n + 1, e.what(), msg, len);
  }
#undef CHECK
  if (msg && len > 0) std::snprintf(msg, (size_t)len, "%d checks passed", n);
  return 0;
}

}  // extern "C"
