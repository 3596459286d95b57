// extern "C" surface of libspice21cu.so — see include/spice21cu.h for the contract and the reference interfaces
// each entry point replaces.
#include "../../include/spice21cu.h"

#include <cstdlib>
#include <cstring>

#include "host/batch.hpp"
#include "host/sweep.hpp"

using namespace s21;

struct s21_ckt {
  CktSpec spec;
  bool elaborated = false;
  FlatCkt flat;
  std::vector<std::pair<std::string, double>> ics;
};
struct s21_batch {
  std::unique_ptr<Batch> b;
  const s21_ckt* ckt = nullptr;
};
struct s21_sweep {
  std::unique_ptr<Sweep> s;
  const s21_ckt* ckt = nullptr;
};

namespace {
thread_local std::string g_last_error;

int32_t fail(const std::exception& e) {
  g_last_error = e.what();
  if (auto* s = dynamic_cast<const S21Error*>(&e)) return s->code;
  if (dynamic_cast<const DecodeError*>(&e)) return S21_DECODE_ERROR;
  return S21_OTHER;
}
#define S21_TRY try {
#define S21_CATCH \
  }               \
  catch (const std::exception& e) { return fail(e); }

SimOptions to_opts(const s21_options* o) {
  SimOptions r;
  if (o) {
    if (!std::isnan(o->temp)) r.temp = o->temp;
    if (!std::isnan(o->tnom)) r.tnom = o->tnom;
    if (!std::isnan(o->gmin)) r.gmin = o->gmin;
    if (!std::isnan(o->iabstol)) r.iabstol = o->iabstol;
    if (!std::isnan(o->reltol)) r.reltol = o->reltol;
  }
  return r;
}
std::vector<CompSpec>& comp_list(s21_ckt* c, const char* module) {
  if (!module || !*module) return c->spec.comps;
  auto it = c->spec.modules.find(module);
  if (it == c->spec.modules.end()) throw S21Error(ST_INVALID, std::string("ModuleDef not found: ") + module);
  return it->second.comps;
}
const char* nz(const char* s) { return s ? s : ""; }

// Tran::solve's time axis (analysis.rs:551-570): t = tstep; while t < tstop { push(t); t += tstep }
std::vector<double> tran_times(double tstep, double tstop) {
  std::vector<double> t;
  t.push_back(0.0);
  if (!(tstep > 0.0)) return t;  // a zero step would never terminate in the reference
  double x = tstep;
  while (x < tstop && t.size() < (size_t)1e9) { t.push_back(x); x += tstep; }
  return t;
}
std::vector<double> ac_freqs(uint64_t fstart, uint64_t fstop_, uint64_t npts) {  // analysis.rs:791-819
  std::vector<double> out;
  double f = (double)fstart;
  const double fstop = (double)fstop_;
  const double fstep = std::pow(10.0, std::log10(fstop / f) / (double)npts);
  while (f <= fstop) {
    out.push_back(f);
    if (f == fstop) break;
    f = std::fmin(f * fstep, fstop);
    if (out.size() > (size_t)1e8) break;
  }
  return out;
}

struct Decoded {
  CktSpec ckt;
  bool has_ckt = false;
  pb::SimOptionsPb opts;
  bool has_opts = false;
  // TranOptions (spice21.proto:155-159) / AcOptions (:192-196)
  double tstop = 0.0, tstep = 0.0;
  std::vector<std::pair<std::string, double>> ic;
  uint64_t fstart = 0, fstop = 0, npts = 0;
};
// Op / Tran / Ac all share {ckt = 1, opts = 2, args = 3}
Decoded decode_sim(const uint8_t* d, size_t n, int kind /*0 op, 1 tran, 2 ac*/) {
  Decoded r;
  PbReader rd(d, n);
  uint32_t f, w;
  while (rd.next(&f, &w)) {
    if (f == 1 && w == 2) { r.ckt = pb::circuit(rd.sub()); r.has_ckt = true; }
    else if (f == 2 && w == 2) { r.opts = pb::sim_options(rd.sub()); r.has_opts = true; }
    else if (f == 3 && w == 2 && kind == 1) {
      PbReader a = rd.sub();
      uint32_t af, aw;
      while (a.next(&af, &aw)) {
        if (af == 1 && aw == 1) r.tstop = a.fixed64_double();
        else if (af == 2 && aw == 1) r.tstep = a.fixed64_double();
        else if (af == 3 && aw == 2) {
          PbReader e = a.sub();
          std::string k;
          double v = 0.0;
          uint32_t ef, ew;
          while (e.next(&ef, &ew)) {
            if (ef == 1 && ew == 2) k = e.str();
            else if (ef == 2 && ew == 1) v = e.fixed64_double();
            else e.skip(ew);
          }
          r.ic.push_back({k, v});
        } else a.skip(aw);
      }
    } else if (f == 3 && w == 2 && kind == 2) {
      PbReader a = rd.sub();
      uint32_t af, aw;
      while (a.next(&af, &aw)) {
        if (aw == 0 && af == 1) r.fstart = a.varint();
        else if (aw == 0 && af == 2) r.fstop = a.varint();
        else if (aw == 0 && af == 3) r.npts = a.varint();
        else a.skip(aw);
      }
    } else rd.skip(w);
  }
  if (!r.has_ckt) throw S21Error(ST_OTHER, "No Circuit Provided");  // proto.rs:62
  return r;
}
int32_t give_bytes(const PbWriter& w, uint8_t** out, size_t* out_n) {
  *out = (uint8_t*)std::malloc(std::max<size_t>(w.buf.size(), 1));
  if (!*out) { g_last_error = "out of memory"; return S21_OTHER; }
  std::memcpy(*out, w.buf.data(), w.buf.size());
  *out_n = w.buf.size();
  return S21_OK;
}
s21_options opts_of(const Decoded& d) {
  s21_options o;
  o.temp = o.tnom = o.gmin = o.iabstol = o.reltol = NAN;
  if (d.has_opts) { o.temp = d.opts.v[0]; o.tnom = d.opts.v[1]; o.gmin = d.opts.v[2]; o.iabstol = d.opts.v[3]; o.reltol = d.opts.v[4]; }
  return o;
}
void elaborate_into(s21_ckt* c, const s21_options* o, const std::vector<std::pair<std::string, double>>& ics) {
  c->ics = ics;
  Flattener f(c->spec, to_opts(o));
  c->flat = f.run(ics);
  c->elaborated = true;
}
}  // namespace

extern "C" {

const char* s21_last_error(void) { return g_last_error.c_str(); }
void s21_free(uint8_t* p) { std::free(p); }
int32_t s21_cuda_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// ------------------------------------------------------------------------------------------------ circuits
int32_t s21_ckt_from_proto(const uint8_t* circuit, size_t n, s21_ckt** out) {
  S21_TRY
  auto* c = new s21_ckt();
  try { c->spec = pb::circuit(PbReader(circuit, n)); } catch (...) { delete c; throw; }
  *out = c;
  return S21_OK;
  S21_CATCH
}
int32_t s21_ckt_new(s21_ckt** out) { *out = new s21_ckt(); return S21_OK; }
void s21_ckt_destroy(s21_ckt* c) { delete c; }

int32_t s21_ckt_signal(s21_ckt* c, const char* module, const char* name) {
  S21_TRY
  if (!module || !*module) c->spec.signals.push_back(nz(name));
  else {
    auto it = c->spec.modules.find(module);
    if (it == c->spec.modules.end()) throw S21Error(ST_INVALID, std::string("ModuleDef not found: ") + module);
    it->second.signals.push_back(nz(name));
  }
  return S21_OK;
  S21_CATCH
}
static int32_t add_two_term(s21_ckt* c, const char* module, CompKind k, const char* name, const char* p, const char* n, double v, double acm) {
  S21_TRY
  CompSpec s;
  s.kind = k; s.name = nz(name); s.p = nz(p); s.n = nz(n); s.val = v; s.acm = acm;
  comp_list(c, module).push_back(s);
  return S21_OK;
  S21_CATCH
}
int32_t s21_ckt_add_r(s21_ckt* c, const char* m, const char* name, const char* p, const char* n, double g) { return add_two_term(c, m, CK_R, name, p, n, g, 0.0); }
int32_t s21_ckt_add_c(s21_ckt* c, const char* m, const char* name, const char* p, const char* n, double cap) { return add_two_term(c, m, CK_C, name, p, n, cap, 0.0); }
int32_t s21_ckt_add_i(s21_ckt* c, const char* m, const char* name, const char* p, const char* n, double dc) { return add_two_term(c, m, CK_I, name, p, n, dc, 0.0); }
int32_t s21_ckt_add_v(s21_ckt* c, const char* m, const char* name, const char* p, const char* n, double dc, double acm) { return add_two_term(c, m, CK_V, name, p, n, dc, acm); }
int32_t s21_ckt_add_v_wave(s21_ckt* c, const char* module, const char* name, const char* p, const char* n, double dc, double acm, int32_t kind,
                           size_t n_params, const double* params) {
  S21_TRY
  if (kind != SRC_PULSE && kind != SRC_SIN) throw S21Error(ST_INVALID, "source wave kind must be 1 (PULSE) or 2 (SIN)");
  if (n_params > 7 || (n_params && !params)) throw S21Error(ST_INVALID, "a source wave takes at most 7 parameters");
  CompSpec s;
  s.kind = CK_V; s.name = nz(name); s.p = nz(p); s.n = nz(n); s.val = dc; s.acm = acm; s.wave_kind = kind;
  for (size_t k = 0; k < n_params; k++) s.wave[k] = params[k];
  if (kind == SRC_PULSE && (!(s.wave[3] > 0.0) || !(s.wave[4] > 0.0))) throw S21Error(ST_INVALID, "PULSE needs rise and fall times > 0");
  comp_list(c, module).push_back(s);
  return S21_OK;
  S21_CATCH
}
int32_t s21_ckt_add_d(s21_ckt* c, const char* module, const char* name, const char* p, const char* n, const char* model, const char* params) {
  S21_TRY
  CompSpec s;
  s.kind = CK_D; s.name = nz(name); s.p = nz(p); s.n = nz(n); s.model = nz(model); s.params = nz(params);
  comp_list(c, module).push_back(s);
  return S21_OK;
  S21_CATCH
}
int32_t s21_ckt_add_mos(s21_ckt* c, const char* module, const char* name, const char* model, const char* params, const char* d, const char* g,
                        const char* s_, const char* b) {
  S21_TRY
  CompSpec s;
  s.kind = CK_MOS; s.name = nz(name); s.model = nz(model); s.params = nz(params); s.d = nz(d); s.g = nz(g); s.s = nz(s_); s.b = nz(b);
  comp_list(c, module).push_back(s);
  return S21_OK;
  S21_CATCH
}
int32_t s21_ckt_add_x(s21_ckt* c, const char* module, const char* name, const char* module_name, size_t n_ports, const char* const* port_names,
                      const char* const* port_nodes) {
  S21_TRY
  CompSpec s;
  s.kind = CK_X; s.name = nz(name); s.module = nz(module_name);
  for (size_t k = 0; k < n_ports; k++) s.ports.push_back({nz(port_names[k]), nz(port_nodes[k])});
  comp_list(c, module).push_back(s);
  return S21_OK;
  S21_CATCH
}
int32_t s21_ckt_def_module(s21_ckt* c, const char* name, size_t n_ports, const char* const* ports) {
  S21_TRY
  ModuleSpec m;
  m.name = nz(name);
  for (size_t k = 0; k < n_ports; k++) m.ports.push_back(nz(ports[k]));
  c->spec.modules[m.name] = m;
  return S21_OK;
  S21_CATCH
}
int32_t s21_ckt_define(s21_ckt* c, const char* kind_, const char* name_, int32_t mos_type, size_t n, const char* const* keys, const double* vals) {
  S21_TRY
  std::string kind = nz(kind_), name = nz(name_);
  ParamBag bag;
  MosModelSpec mm;
  mm.mos_type = mos_type;
  for (size_t k = 0; k < n; k++) {
    std::string key = nz(keys[k]);
    if (key == "tpg" && (kind == "mos1model")) { mm.has_tpg = true; mm.tpg = (long)vals[k]; continue; }
    bag.kv[key] = vals[k];
  }
  mm.p = bag;
  if (kind == "mos0") c->spec.mos0[name] = mos_type;
  else if (kind == "mos1model") c->spec.mos1_models[name] = mm;
  else if (kind == "mos1inst") c->spec.mos1_insts[name] = bag;
  else if (kind == "diodemodel") c->spec.diode_models[name] = bag;
  else if (kind == "diodeinst") c->spec.diode_insts[name] = bag;
  else if (kind == "bsim4model") c->spec.bsim4_models[name] = mm;
  else if (kind == "bsim4inst") c->spec.bsim4_insts[name] = bag;
  else throw S21Error(ST_OTHER, "unknown definition kind: " + kind);
  return S21_OK;
  S21_CATCH
}

int32_t s21_ckt_elaborate(s21_ckt* c, const s21_options* opts, size_t n_ic, const char* const* ic_nodes, const double* ic_vals) {
  S21_TRY
  std::vector<std::pair<std::string, double>> ics;
  for (size_t k = 0; k < n_ic; k++) ics.push_back({nz(ic_nodes[k]), ic_vals[k]});
  elaborate_into(c, opts, ics);
  return S21_OK;
  S21_CATCH
}
int32_t s21_ckt_num_vars(const s21_ckt* c) { return c->elaborated ? c->flat.n_vars() : -1; }
const char* s21_ckt_var_name(const s21_ckt* c, int32_t i) {
  if (!c->elaborated || i < 0 || i >= c->flat.n_vars()) return nullptr;
  return c->flat.var_names[(size_t)i].c_str();
}
int32_t s21_ckt_var_kind(const s21_ckt* c, int32_t i) {
  if (!c->elaborated || i < 0 || i >= c->flat.n_vars()) return -1;
  return c->flat.var_kinds[(size_t)i];
}
int32_t s21_ckt_num_devices(const s21_ckt* c) { return c->elaborated ? (int32_t)c->flat.devs.size() : -1; }
int32_t s21_ckt_stamp_map(const s21_ckt* c, const int32_t** elem_row, const int32_t** elem_col, size_t* n_elem, const int32_t** dev_off,
                          const int32_t** dev_elems) {
  S21_TRY
  if (!c->elaborated) throw S21Error(ST_OTHER, "circuit is not elaborated");
  if (elem_row) *elem_row = c->flat.elem_row.data();
  if (elem_col) *elem_col = c->flat.elem_col.data();
  if (n_elem) *n_elem = c->flat.elem_row.size();
  if (dev_off) *dev_off = c->flat.dev_off_export.data();
  if (dev_elems) *dev_elems = c->flat.dev_elems_export.data();
  return S21_OK;
  S21_CATCH
}

// ------------------------------------------------------------------------------------------------ batches
int32_t s21_batch_create(const s21_ckt* c, int32_t cuda_device, size_t B, s21_batch** out) {
  S21_TRY
  if (!c->elaborated) throw S21Error(ST_OTHER, "circuit is not elaborated");
  auto* b = new s21_batch();
  try { b->b.reset(new Batch(c->spec, c->flat, cuda_device, B)); } catch (...) { delete b; throw; }
  b->ckt = c;
  *out = b;
  return S21_OK;
  S21_CATCH
}
void s21_batch_destroy(s21_batch* b) { delete b; }
int32_t s21_batch_set_stream(s21_batch* b, void* s) { b->b->set_stream(s); return S21_OK; }
int32_t s21_batch_override(s21_batch* b, const char* spec, const double* values) {
  S21_TRY
  b->b->add_override(nz(spec), values);
  return S21_OK;
  S21_CATCH
}
int32_t s21_batch_sync_params(s21_batch* b, int32_t force_upload, size_t* h2d_bytes) {
  S21_TRY
  b->b->sync_params(force_upload != 0);
  if (h2d_bytes) *h2d_bytes = b->b->last_h2d_bytes();
  return S21_OK;
  S21_CATCH
}
int32_t s21_batch_set_aids(s21_batch* b, int32_t flags) {
  S21_TRY
  b->b->set_aids(flags);
  return S21_OK;
  S21_CATCH
}
int32_t s21_batch_reset(s21_batch* b) {
  S21_TRY
  b->b->reset();
  return S21_OK;
  S21_CATCH
}
int32_t s21_batch_dcop_device(s21_batch* b) {
  S21_TRY
  b->b->dcop_device();
  return S21_OK;
  S21_CATCH
}
int32_t s21_batch_read(s21_batch* b, double* x, int32_t* status, int32_t* iters) {
  S21_TRY
  b->b->read(x, status, iters);
  return S21_OK;
  S21_CATCH
}
int32_t s21_batch_dcop(s21_batch* b, double* x, int32_t* status, int32_t* iters) {
  S21_TRY
  b->b->dcop_device(x != nullptr);
  b->b->read(x, status, iters);
  return S21_OK;
  S21_CATCH
}
int32_t s21_batch_dcop_view(s21_batch* b, const double** x, const int32_t** status, const int32_t** iters) {
  S21_TRY
  b->b->dcop_device(x != nullptr);
  b->b->read_view(x != nullptr, x, status, iters);
  return S21_OK;
  S21_CATCH
}
int32_t s21_batch_step_dcop_view(s21_batch* b, int32_t flags, const double** x, const int32_t** status, const int32_t** iters,
                                 size_t* h2d_bytes) {
  S21_TRY
  if (flags & 2) b->b->reset();
  b->b->dcop_device(x != nullptr, (flags & 1) != 0);
  if (h2d_bytes) *h2d_bytes = (flags & 1) ? b->b->last_h2d_bytes() : 0;
  b->b->read_view(x != nullptr, x, status, iters);
  return S21_OK;
  S21_CATCH
}
int32_t s21_batch_packed_device(s21_batch* b, const double** dev_ptr, size_t* n_words) {
  S21_TRY
  b->b->packed_device(dev_ptr, n_words);
  return S21_OK;
  S21_CATCH
}
int32_t s21_batch_wave_device(const s21_batch* b, const double** dev_ptr, size_t* T, size_t* n_save, size_t* stride) {
  S21_TRY
  b->b->wave_device(dev_ptr, T, n_save, stride);
  return S21_OK;
  S21_CATCH
}
int64_t s21_tran_num_points(double tstep, double tstop) { return (int64_t)tran_times(tstep, tstop).size(); }
int32_t s21_batch_tran(s21_batch* b, double tstep, double tstop, const int32_t* save_vars, size_t n_save, double* time, double* wave,
                       int32_t* status, int64_t* iters) {
  S21_TRY
  std::vector<double> t = tran_times(tstep, tstop);
  if (time) std::memcpy(time, t.data(), t.size() * sizeof(double));
  b->b->tran(tstep, (int)t.size(), save_vars, n_save, wave, status, iters);
  return S21_OK;
  S21_CATCH
}
int32_t s21_batch_tran_adaptive(s21_batch* b, double tstep, double tstop, const double* ctl7, const int32_t* save_vars, size_t n_save, double* time,
                                double* wave, int32_t* status, int64_t* iters, int32_t* accepted, int32_t* rejected) {
  S21_TRY
  std::vector<double> t = tran_times(tstep, tstop);
  if (time) std::memcpy(time, t.data(), t.size() * sizeof(double));
  b->b->tran_adaptive(tstep, (int)t.size(), ctl7, save_vars, n_save, wave, status, iters, accepted, rejected);
  return S21_OK;
  S21_CATCH
}
int64_t s21_ac_freqs(uint64_t fstart, uint64_t fstop, uint64_t npts, double* freqs, size_t cap) {
  std::vector<double> f = ac_freqs(fstart, fstop, npts);
  if (freqs) for (size_t k = 0; k < f.size() && k < cap; k++) freqs[k] = f[k];
  return (int64_t)f.size();
}
int32_t s21_batch_ac(s21_batch* b, const double* freqs, size_t F, double* x, int32_t* status, int32_t* iters) {
  S21_TRY
  b->b->ac(freqs, F, x, status, iters);
  return S21_OK;
  S21_CATCH
}
int32_t s21_batch_pivot_order(const s21_batch* b, const int32_t** row_i2e, const int32_t** col_i2e, size_t* n, const int32_t** lu_row,
                              const int32_t** lu_col, const int32_t** lu_is_fill, size_t* nnz_lu) {
  S21_TRY
  const Plan* p = b->b->last_plan();
  if (!p) throw S21Error(ST_OTHER, "no solve has run on this batch yet");
  if (row_i2e) *row_i2e = p->row_i2e.data();
  if (col_i2e) *col_i2e = p->col_i2e.data();
  if (n) *n = (size_t)p->N;
  if (lu_row) *lu_row = p->lu_row.data();
  if (lu_col) *lu_col = p->lu_col.data();
  if (lu_is_fill) *lu_is_fill = p->lu_fill.data();
  if (nnz_lu) *nnz_lu = (size_t)p->nnzLU;
  return S21_OK;
  S21_CATCH
}
int32_t s21_batch_stats(const s21_batch* b, double* out8) {
  S21_TRY
  b->b->stats(out8);
  return S21_OK;
  S21_CATCH
}

const char* s21_batch_kernel_name(const s21_batch* b) { return b && b->b ? b->b->kernel_name() : ""; }

int32_t s21_batch_plan_info(const s21_batch* b, int64_t* out8) {
  S21_TRY
  const Plan* p = b->b->last_plan();
  if (!p) throw S21Error(ST_OTHER, "no solve has run on this batch yet");
  auto levels = [](const std::vector<int>& off) { return off.empty() ? (int64_t)0 : (int64_t)off.size() - 1; };
  out8[0] = p->N; out8[1] = p->nnzLU; out8[2] = (int64_t)p->lu_t.size(); out8[3] = levels(p->lu_lvl_off);
  out8[4] = (int64_t)p->fw_k.size(); out8[5] = levels(p->fw_lvl_off); out8[6] = levels(p->bw_lvl_off); out8[7] = p->relaxed ? 1 : 0;
  return S21_OK;
  S21_CATCH
}
int32_t s21_batch_setup_stats(const s21_batch* b, double* out8) {
  S21_TRY
  const jit::CacheStats& cs = jit::cache_stats();
  out8[0] = b && b->b ? b->b->symbolic_seconds() : 0.0;
  out8[1] = cs.nvrtc_seconds; out8[2] = (double)cs.nvrtc_runs; out8[3] = (double)cs.disk_hits; out8[4] = (double)cs.mem_hits;
  out8[5] = b && b->b ? (double)b->b->weak_seen() : 0.0;
  out8[6] = b && b->b ? (double)b->b->repaired() : 0.0;
  out8[7] = b && b->b ? (double)b->b->aided() : 0.0;
  return S21_OK;
  S21_CATCH
}

// ------------------------------------------------------------------------------------------------ multi-GPU sweeps
int32_t s21_sweep_partition(size_t B, int32_t n_devices, int32_t g, size_t* first, size_t* count) {
  S21_TRY
  if (n_devices <= 0 || g < 0 || g >= n_devices) throw S21Error(ST_OTHER, "s21_sweep_partition: 0 <= g < n_devices required");
  size_t f = 0, c = 0;
  sweep_partition(B, n_devices, g, &f, &c);
  if (first) *first = f;
  if (count) *count = c;
  return S21_OK;
  S21_CATCH
}
int32_t s21_sweep_create(const s21_ckt* c, int32_t n_devices, const int32_t* devices, size_t B, s21_sweep** out) {
  S21_TRY
  if (!c->elaborated) throw S21Error(ST_OTHER, "circuit is not elaborated");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    throw S21Error(ST_CUDA, "no CUDA device available: libspice21cu has no CPU fallback (" + std::string(cudaGetErrorString(e)) + ")");
  if (n_devices <= 0) n_devices = count;  // all visible devices
  std::vector<int> devs;
  for (int g = 0; g < n_devices; g++) {
    const int d = devices ? devices[g] : g;
    if (d < 0 || d >= count) throw S21Error(ST_CUDA, "invalid CUDA device index in s21_sweep_create");
    devs.push_back(d);
  }
  auto* s = new s21_sweep();
  try { s->s.reset(new Sweep(c->spec, c->flat, devs, B)); } catch (...) { delete s; throw; }
  s->ckt = c;
  *out = s;
  return S21_OK;
  S21_CATCH
}
void s21_sweep_destroy(s21_sweep* s) { delete s; }
int32_t s21_sweep_num_devices(const s21_sweep* s) { return s && s->s ? s->s->n_devices() : 0; }
int32_t s21_sweep_shard(const s21_sweep* s, int32_t g, int32_t* cuda_device, size_t* first, size_t* count) {
  S21_TRY
  if (g < 0 || g >= s->s->n_devices()) throw S21Error(ST_OTHER, "shard index out of range");
  int d = 0;
  size_t f = 0, c = 0;
  s->s->shard_range(g, &d, &f, &c);
  if (cuda_device) *cuda_device = d;
  if (first) *first = f;
  if (count) *count = c;
  return S21_OK;
  S21_CATCH
}
int32_t s21_sweep_override(s21_sweep* s, const char* spec, const double* values) {
  S21_TRY
  s->s->add_override(nz(spec), values);
  return S21_OK;
  S21_CATCH
}
int32_t s21_sweep_sync_params(s21_sweep* s, int32_t force_upload, size_t* h2d_bytes) {
  S21_TRY
  const size_t n = s->s->sync_params(force_upload != 0);
  if (h2d_bytes) *h2d_bytes = n;
  return S21_OK;
  S21_CATCH
}
int32_t s21_sweep_reset(s21_sweep* s) {
  S21_TRY
  s->s->reset();
  return S21_OK;
  S21_CATCH
}
int32_t s21_sweep_dcop(s21_sweep* s, double* x, int32_t* status, int32_t* iters) {
  S21_TRY
  s->s->dcop(x, status, iters);
  return S21_OK;
  S21_CATCH
}
int32_t s21_sweep_dcop_view(s21_sweep* s, const double** x, const int32_t** status, const int32_t** iters) {
  S21_TRY
  s->s->dcop_view(x, status, iters);
  return S21_OK;
  S21_CATCH
}
int32_t s21_sweep_tran(s21_sweep* s, double tstep, double tstop, const int32_t* save_vars, size_t n_save, double* time, double* wave,
                       int32_t* status, int64_t* iters) {
  S21_TRY
  std::vector<double> t = tran_times(tstep, tstop);
  if (time) std::memcpy(time, t.data(), t.size() * sizeof(double));
  s->s->tran(tstep, (int)t.size(), save_vars, n_save, wave, status, iters);
  return S21_OK;
  S21_CATCH
}
int32_t s21_sweep_ac(s21_sweep* s, const double* freqs, size_t F, double* x, int32_t* status, int32_t* iters) {
  S21_TRY
  s->s->ac(freqs, F, x, status, iters);
  return S21_OK;
  S21_CATCH
}
int32_t s21_sweep_stats(const s21_sweep* s, double* out8) {
  S21_TRY
  s->s->stats(out8);
  return S21_OK;
  S21_CATCH
}

int32_t s21_jit_source(const s21_ckt* c, int32_t mode, int32_t shape, const double* vals, size_t n_vals, uint8_t** out, size_t* out_n,
                       size_t* smem_bytes) {
  S21_TRY
  if (!c->elaborated) throw S21Error(ST_OTHER, "circuit is not elaborated");
  if (n_vals != c->flat.elem_row.size()) throw S21Error(ST_INVALID, "s21_jit_source: one value per matrix element expected");
  size_t smem = 0;
  const std::string src = debug_jit_source(c->flat, mode, shape, vals, &smem);
  *out = (uint8_t*)std::malloc(src.size() + 1);
  std::memcpy(*out, src.c_str(), src.size() + 1);
  if (out_n) *out_n = src.size();
  if (smem_bytes) *smem_bytes = smem;
  return S21_OK;
  S21_CATCH
}
int32_t s21_jit_check(const uint8_t* src, size_t n) {
  S21_TRY
  std::vector<char> cubin;
  std::string err;
  if (!jit::cubin_for(std::string((const char*)src, n), &cubin, &err)) throw S21Error(ST_OTHER, err);
  return S21_OK;
  S21_CATCH
}

int32_t s21_selftest_div(uint64_t n, uint64_t seed, uint64_t* mismatches, double* first4) {
  S21_TRY
  unsigned long long bad = 0;
  double f4[4] = {0, 0, 0, 0};
  const int rc = selftest_div(n, seed, &bad, f4);
  if (rc) throw S21Error(ST_CUDA, std::string("selftest_div: ") + cudaGetErrorString((cudaError_t)rc));
  if (mismatches) *mismatches = bad;
  if (first4) std::memcpy(first4, f4, sizeof f4);
  return S21_OK;
  S21_CATCH
}

int32_t s21_symbolic(int32_t n, size_t nnz, const int32_t* rows, const int32_t* cols, const double* vals, int32_t width, int32_t* row_i2e,
                     int32_t* col_i2e, int32_t* lu_row, int32_t* lu_col, int32_t* lu_is_fill, size_t cap, size_t* nnz_lu) {
  S21_TRY
  std::vector<int> r(rows, rows + nnz), c(cols, cols + nnz);
  Plan p;
  if (width == 2) {
    std::vector<cplx> z(nnz);
    for (size_t k = 0; k < nnz; k++) z[k] = mk(vals[2 * k], vals[2 * k + 1]);
    p = build_plan<cplx>(n, r, c, z.data());
  } else {
    p = build_plan<double>(n, r, c, vals);
  }
  for (int k = 0; k < n; k++) { row_i2e[k] = p.row_i2e[(size_t)k]; col_i2e[k] = p.col_i2e[(size_t)k]; }
  for (size_t k = 0; k < (size_t)p.nnzLU && k < cap; k++) { lu_row[k] = p.lu_row[k]; lu_col[k] = p.lu_col[k]; lu_is_fill[k] = p.lu_fill[k]; }
  *nnz_lu = (size_t)p.nnzLU;
  g_last_error = p.status == ST_OK ? "" : Batch::status_text(p.status);
  return p.status;
  S21_CATCH
}

// ------------------------------------------------------------------------------------------------ bytes API
// decode -> elaborate -> one-instance batch on device 0 -> encode (CallableProto::call_bytes, proto.rs:40-44)
int32_t s21_op_bytes(const uint8_t* op, size_t n, uint8_t** out, size_t* out_n) {
  S21_TRY
  Decoded d = decode_sim(op, n, 0);
  s21_ckt c;
  c.spec = d.ckt;
  s21_options o = opts_of(d);
  elaborate_into(&c, &o, {});
  Batch b(c.spec, c.flat, 0, 1);
  const int N = c.flat.n_vars();
  std::vector<double> x((size_t)N);
  int32_t st = 0;
  b.dcop_device();
  b.read(x.data(), &st, nullptr);
  if (st != S21_OK) throw S21Error(st, Batch::status_text(st));
  PbWriter w;  // OpResult { map<string,double> vals = 1 }  (spice21.proto:150-152)
  for (int k = 0; k < N; k++) {
    PbWriter e;
    e.f_string(1, c.flat.var_names[(size_t)k]);
    e.f_double(2, x[(size_t)k]);
    w.f_msg(1, e);
  }
  return give_bytes(w, out, out_n);
  S21_CATCH
}
int32_t s21_tran_bytes(const uint8_t* tran, size_t n, uint8_t** out, size_t* out_n) {
  S21_TRY
  Decoded d = decode_sim(tran, n, 1);
  s21_ckt c;
  c.spec = d.ckt;
  s21_options o = opts_of(d);
  elaborate_into(&c, &o, d.ic);
  Batch b(c.spec, c.flat, 0, 1);
  const int N = c.flat.n_vars();
  std::vector<double> t = tran_times(d.tstep, d.tstop);
  std::vector<int32_t> save((size_t)N);
  for (int k = 0; k < N; k++) save[(size_t)k] = k;
  std::vector<double> wave(t.size() * (size_t)N);
  int32_t st = 0;
  b.tran(d.tstep, (int)t.size(), save.data(), (size_t)N, wave.data(), &st, nullptr);
  if (st != S21_OK) throw S21Error(st, Batch::status_text(st));
  PbWriter w;  // TranResult { DoubleArray time = 1; map<string,DoubleArray> vals = 2 }  (spice21.proto:175-178)
  {
    PbWriter a;
    a.f_packed_doubles(1, t.data(), t.size());
    w.f_msg(1, a);
  }
  auto put = [&](const std::string& name, const std::vector<double>& v) {
    PbWriter a, e;
    a.f_packed_doubles(1, v.data(), v.size());
    e.f_string(1, name);
    e.f_msg(2, a);
    w.f_msg(2, e);
  };
  put("time", t);  // TranResult::end inserts "time" into the signal map too (analysis.rs:605)
  std::vector<double> col(t.size());
  for (int k = 0; k < N; k++) {
    for (size_t p = 0; p < t.size(); p++) col[p] = wave[p * (size_t)N + (size_t)k];
    put(c.flat.var_names[(size_t)k], col);
  }
  return give_bytes(w, out, out_n);
  S21_CATCH
}
int32_t s21_ac_bytes(const uint8_t* ac, size_t n, uint8_t** out, size_t* out_n) {
  S21_TRY
  Decoded d = decode_sim(ac, n, 2);
  s21_ckt c;
  c.spec = d.ckt;
  s21_options o = opts_of(d);
  elaborate_into(&c, &o, {});
  Batch b(c.spec, c.flat, 0, 1);
  const int N = c.flat.n_vars();
  std::vector<double> f = ac_freqs(d.fstart, d.fstop, d.npts);
  std::vector<double> x(f.size() * (size_t)N * 2);
  std::vector<int32_t> st(f.size());
  b.ac(f.data(), f.size(), x.data(), st.data(), nullptr);
  for (int32_t s : st) if (s != S21_OK) throw S21Error(s, Batch::status_text(s));
  PbWriter w;  // AcResult { DoubleArray freq = 1; map<string,ComplexArray> vals = 2 }  (spice21.proto:203-206)
  {
    PbWriter a;
    a.f_packed_doubles(1, f.data(), f.size());
    w.f_msg(1, a);
  }
  for (int k = 0; k < N; k++) {
    PbWriter arr;
    for (size_t p = 0; p < f.size(); p++) {
      PbWriter z;
      z.f_double(1, x[(p * (size_t)N + (size_t)k) * 2]);
      z.f_double(2, x[(p * (size_t)N + (size_t)k) * 2 + 1]);
      arr.f_msg(1, z);
    }
    PbWriter e;
    e.f_string(1, c.flat.var_names[(size_t)k]);
    e.f_msg(2, arr);
    w.f_msg(2, e);
  }
  return give_bytes(w, out, out_n);
  S21_CATCH
}

}  // extern "C"
