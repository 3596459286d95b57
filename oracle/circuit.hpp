// ORACLE — TEST INFRASTRUCTURE ONLY (see num.hpp header).
//
// CPU restatement of the reference's middle end:
//   spice21/src/circuit.rs:26-331  NodeRef, Comp, Ckt (node refs are carried as their to_string(): "" == Gnd)
//   spice21/src/defs.rs:44-135     ModuleDefs, ModelInstanceCache (get() derives + caches per (inst, model)), Defs
//   spice21/src/elab.rs:22-265     Elaborator: hierarchy flattening + VARIABLE NUMBERING (first-encounter order)
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "comps.hpp"

namespace orc {

struct CompDesc;
struct ModuleDef {  // proto Module (spice21.proto:95-101)
  std::string name;
  std::vector<std::string> ports, signals;
  std::vector<CompDesc> comps;
};
// circuit.rs:109-117 `enum Comp`
struct CompDesc {
  enum Kind { V, I, R, C, D, MOS, MODULE } kind;
  std::string name;
  std::string p, n;                 // two-terminal devices
  double val = 0.0, acm = 0.0;      // vdc / dc / g / c
  std::string model, params;        // D, MOS
  std::string d, g, s, b;           // MOS ports
  std::string module;               // MODULE
  std::vector<std::pair<std::string, std::string>> ports;  // MODULE port connections
};

struct Mos1CacheEntry {
  std::shared_ptr<Mos1Model> model;
  std::shared_ptr<Mos1InstanceParams> inst;
  std::shared_ptr<Mos1InternalParams> intp;
};
struct DiodeCacheEntry {
  std::shared_ptr<DiodeModel> model;
  std::shared_ptr<DiodeInstParams> inst;
  std::shared_ptr<DiodeIntParams> intp;
};

// defs.rs:129-135
struct Defs {
  std::map<std::string, std::shared_ptr<ModuleDef>> modules;
  std::map<std::string, MosType> mos0;
  std::map<std::string, std::shared_ptr<Mos1Model>> mos1_models;
  std::map<std::string, std::shared_ptr<Mos1InstanceParams>> mos1_insts;
  std::map<std::pair<std::string, std::string>, Mos1CacheEntry> mos1_cache;
  std::map<std::string, std::shared_ptr<DiodeModel>> diode_models;
  std::map<std::string, std::shared_ptr<DiodeInstParams>> diode_insts;
  std::map<std::pair<std::string, std::string>, DiodeCacheEntry> diode_cache;
  // bsim4 depots are added by bsim4.hpp through this opaque hook
  std::shared_ptr<void> bsim4;

  // defs.rs:97-114 ModelInstanceCache::get(inst, model, opts)
  bool mos1_get(const std::string& inst, const std::string& model, const Options& opts, Mos1CacheEntry* out) {
    auto key = std::make_pair(inst, model);
    auto it = mos1_cache.find(key);
    if (it != mos1_cache.end()) { *out = it->second; return true; }
    auto ii = mos1_insts.find(inst);
    if (ii == mos1_insts.end()) return false;
    auto mi = mos1_models.find(model);
    if (mi == mos1_models.end()) return false;
    Mos1CacheEntry e;
    e.model = mi->second;
    e.inst = ii->second;
    e.intp = std::make_shared<Mos1InternalParams>(Mos1InternalParams::derive(*mi->second, *ii->second, opts));
    mos1_cache[key] = e;
    *out = e;
    return true;
  }
  bool diode_get(const std::string& inst, const std::string& model, const Options& opts, DiodeCacheEntry* out) {
    auto key = std::make_pair(inst, model);
    auto it = diode_cache.find(key);
    if (it != diode_cache.end()) { *out = it->second; return true; }
    auto ii = diode_insts.find(inst);
    if (ii == diode_insts.end()) return false;
    auto mi = diode_models.find(model);
    if (mi == diode_models.end()) return false;
    DiodeCacheEntry e;
    e.model = mi->second;
    e.inst = ii->second;
    e.intp = std::make_shared<DiodeIntParams>(DiodeIntParams::derive(*mi->second, *ii->second, opts));
    diode_cache[key] = e;
    *out = e;
    return true;
  }
};

// circuit.rs:222-227
struct Ckt {
  std::string name;
  std::vector<std::string> signals;
  std::vector<CompDesc> comps;
  Defs defs;
};

// Hook for BSIM4 elaboration (set by bsim4.hpp); returns nullptr if `model` is not a BSIM4 model.
typedef std::shared_ptr<Component> (*Bsim4ElabFn)(Defs& defs, const std::string& model, const std::string& params,
                                                  const std::string& path, const VarIndex ports[4], void* vars_any, bool cplx,
                                                  const Options& opts);
inline Bsim4ElabFn& bsim4_elab_hook() { static Bsim4ElabFn f = nullptr; return f; }

typedef std::map<std::string, VarIndex> Namespace;  // HashMap<String, Option<VarIndex>>; -1 == None

// elab.rs:22-240
template <class T>
struct Elaborator {
  std::vector<std::shared_ptr<Component>> comps;
  Variables<T> vars;
  Defs defs;
  std::vector<std::string> path;
  Options opts;

  std::string pathstr() const {  // :175-177
    std::string s;
    for (size_t k = 0; k < path.size(); k++) { if (k) s += "."; s += path[k]; }
    return s;
  }
  bool on_top() const { return path.empty(); }  // :179-181
  static VarIndex ns_get(const Namespace& ns, const std::string& key, const char* what) {
    auto it = ns.find(key);
    if (it == ns.end()) throw Panic(std::string(what) + ": node not in namespace: '" + key + "'");
    return it->second;
  }
  VarIndex node_var(const std::string& node, bool autonode, Namespace& ns) {  // :35-52
    if (autonode) {
      if (node.empty()) return -1;
      path.push_back(node);
      std::string pathname = pathstr();
      VarIndex var = vars.find_or_create(pathname);
      ns[node] = var;
      path.pop_back();
      return var;
    }
    return ns_get(ns, node, "!!!");
  }
  void elaborate_instance(const CompDesc& inst, Namespace& ns, bool autonode) {  // :55-84
    switch (inst.kind) {
      case CompDesc::R: {
        VarIndex pvar = node_var(inst.p, autonode, ns);
        VarIndex nvar = node_var(inst.n, autonode, ns);
        comps.push_back(std::make_shared<Resistor>(inst.val, pvar, nvar));
        break;
      }
      case CompDesc::C: {
        VarIndex pvar = node_var(inst.p, autonode, ns);
        VarIndex nvar = node_var(inst.n, autonode, ns);
        comps.push_back(std::make_shared<Capacitor>(inst.val, pvar, nvar));
        break;
      }
      case CompDesc::I: {
        VarIndex pvar = node_var(inst.p, autonode, ns);
        VarIndex nvar = node_var(inst.n, autonode, ns);
        comps.push_back(std::make_shared<Isrc>(inst.val, pvar, nvar));
        break;
      }
      case CompDesc::V: elaborate_vsrc(inst, ns); break;
      case CompDesc::D: elaborate_diode(inst, ns); break;
      case CompDesc::MOS: elaborate_mos(inst, ns); break;
      case CompDesc::MODULE: elaborate_module_inst(inst, ns); break;
    }
  }
  void elaborate_diode(const CompDesc& d, Namespace& ns) {  // :85-111 (no auto-noding)
    VarIndex pvar = ns_get(ns, d.p, "diode");
    VarIndex nvar = ns_get(ns, d.n, "diode");
    path.push_back(d.name);
    DiodeCacheEntry e;
    if (!defs.diode_get(d.params, d.model, opts, &e)) throw Panic("Parameters not defined: " + d.params);
    auto dd = std::make_shared<Diode>();
    // DiodePorts::from (diode.rs:101-110)
    VarIndex r = e.model->has_rs() ? vars.add(pathstr() + ".r", VarKind::V) : pvar;
    dd->p = pvar; dd->n = nvar; dd->r = r;
    dd->model = e.model;
    dd->intp = e.intp;
    path.pop_back();
    comps.push_back(dd);
  }
  void elaborate_vsrc(const CompDesc& vi, Namespace& ns) {  // :112-127
    VarIndex pvar = node_var(vi.p, on_top(), ns);
    VarIndex nvar = node_var(vi.n, on_top(), ns);
    path.push_back(vi.name);
    VarIndex ivar = vars.addi(pathstr());
    path.pop_back();
    comps.push_back(std::make_shared<Vsrc>(vi.val, vi.acm, pvar, nvar, ivar));
  }
  void elaborate_mos(const CompDesc& m, Namespace& ns) {  // :128-173
    VarIndex ports[4];
    ports[0] = node_var(m.d, on_top(), ns);
    ports[1] = node_var(m.g, on_top(), ns);
    ports[2] = node_var(m.s, on_top(), ns);
    ports[3] = node_var(m.b, on_top(), ns);
    path.push_back(m.name);
    std::shared_ptr<Component> c;
    if (bsim4_elab_hook()) c = bsim4_elab_hook()(defs, m.model, m.params, pathstr(), ports, (void*)&vars, sizeof(T) != sizeof(double), opts);
    if (c) {
      // Bsim4 (checked first, elab.rs:143-146)
    } else if (defs.mos1_models.count(m.model)) {
      Mos1CacheEntry e;
      if (!defs.mos1_get(m.params, m.model, opts, &e)) throw Panic("Parameters not defined: " + m.params);
      auto mm = std::make_shared<Mos1>();
      // Mos1Vars::from (mos.rs:593-619): dp first, then sp
      VarIndex dp = (e.model->rd.some || e.model->rsh.some) ? vars.add(pathstr() + ".dp", VarKind::V) : ports[0];
      VarIndex sp = (e.model->rs.some || e.model->rsh.some) ? vars.add(pathstr() + ".sp", VarKind::V) : ports[2];
      mm->ports[M1_D] = ports[0]; mm->ports[M1_G] = ports[1]; mm->ports[M1_S] = ports[2]; mm->ports[M1_B] = ports[3];
      mm->ports[M1_DP] = dp; mm->ports[M1_SP] = sp;
      mm->model = e.model;
      mm->intparams = e.intp;
      c = mm;
    } else if (defs.mos0.count(m.model)) {
      c = std::make_shared<Mos0>(ports, defs.mos0[m.model]);
    } else {
      throw Panic("Model not defined: " + m.model);
    }
    comps.push_back(c);
    path.pop_back();
  }
  void elaborate_module_inst(const CompDesc& m, Namespace& ns) {  // :182-207
    auto it = defs.modules.find(m.module);
    if (it == defs.modules.end()) throw Panic("ModuleDef not found: " + m.module);
    std::shared_ptr<ModuleDef> mdef = it->second;
    Namespace inst_ns;
    for (auto& kv : m.ports) inst_ns[kv.first] = ns_get(ns, kv.second, "module port");
    path.push_back(m.name);
    if (path.size() > 1024) throw Panic("Elaboration Error: Too deep a hierarchy (for now)!");
    elaborate_module(*mdef, inst_ns);
    path.pop_back();
  }
  void elaborate_signal(const std::string& signame, Namespace& ns) {  // :209-216
    path.push_back(signame);
    VarIndex var = vars.addv(pathstr());
    ns[signame] = var;
    path.pop_back();
  }
  void elaborate_module(const ModuleDef& m, Namespace& ns) {  // :218-239
    for (auto& s : m.signals) elaborate_signal(s, ns);
    for (auto& inst : m.comps) elaborate_instance(inst, ns, false);
  }
};

// elab.rs:244-265
template <class T>
inline Elaborator<T> elaborate(const Ckt& ckt, const Options& opts) {
  Elaborator<T> e;
  e.defs = ckt.defs;
  e.opts = opts;
  Namespace ns;
  ns[""] = -1;
  for (auto& s : ckt.signals) e.elaborate_signal(s, ns);
  for (auto& inst : ckt.comps) e.elaborate_instance(inst, ns, true);
  return e;
}

}  // namespace orc
