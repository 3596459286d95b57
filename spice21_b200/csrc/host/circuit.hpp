// Host-side circuit description: what Ckt::from_proto (spice21/src/circuit.rs:281-330) produces, kept as flat
// "spec" tables so that per-instance overrides (Monte-Carlo / sweeps) can re-run the parameter derivations.
#pragma once
#include <cmath>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "pbwire.hpp"

namespace s21 {

// Error carrying one of the S21_* status codes of include/spice21cu.h
struct S21Error : std::runtime_error {
  int code;
  S21Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
enum { ST_OK = 0, ST_CONV = 1, ST_SINGULAR = 2, ST_PIVOT = 3, ST_DECODE = 4, ST_INVALID = 5, ST_UNSUPPORTED = 6, ST_CUDA = 7, ST_OTHER = 8 };

// Optional-valued parameter set (the prost `Option<f64>` fields of the *Model / *InstParams messages).
struct ParamBag {
  std::map<std::string, double> kv;
  bool has(const std::string& k) const { return kv.count(k) != 0; }
  double get(const std::string& k, double dflt) const {
    auto it = kv.find(k);
    return it == kv.end() ? dflt : it->second;
  }
};

enum CompKind { CK_R, CK_C, CK_I, CK_V, CK_D, CK_MOS, CK_X };

struct CompSpec {
  CompKind kind = CK_R;
  std::string name;
  std::string p, n;            // two-terminal devices ("" = ground)
  double val = 0.0, acm = 0.0; // g | c | dc ; acm for V
  int wave_kind = 0;           // V only (extension, device_layout.h SRC_*): 0 none, 1 PULSE, 2 SIN
  double wave[7] = {0, 0, 0, 0, 0, 0, 0};
  std::string model, params;   // D, MOS
  std::string d, g, s, b;      // MOS ports
  std::string module;          // X
  std::vector<std::pair<std::string, std::string>> ports;  // X: port -> node
};
struct ModuleSpec {
  std::string name;
  std::vector<std::string> ports, signals;
  std::vector<CompSpec> comps;
};
struct MosModelSpec {
  int mos_type = 0;  // 0 NMOS, 1 PMOS
  bool has_tpg = false;
  long tpg = 0;
  ParamBag p;
};

struct CktSpec {
  std::string name;
  std::vector<std::string> signals;
  std::vector<CompSpec> comps;
  std::map<std::string, ModuleSpec> modules;
  std::map<std::string, int> mos0;  // model name -> mos_type
  std::map<std::string, MosModelSpec> mos1_models;
  std::map<std::string, ParamBag> mos1_insts;
  std::map<std::string, ParamBag> diode_models;
  std::map<std::string, ParamBag> diode_insts;
  std::map<std::string, MosModelSpec> bsim4_models;
  std::map<std::string, ParamBag> bsim4_insts;
};

// ---------------------------------------------------------------------------------------------------- decoding
namespace pb {

inline void two_term(PbReader r, CompSpec* c, bool has_acm) {  // Resistor/Capacitor/Isrc/Vsrc (spice21.proto:8-36)
  uint32_t f, w;
  int nwave = 0;
  while (r.next(&f, &w)) {
    if (f == 1 && w == 2) c->name = r.str();
    else if (f == 2 && w == 2) c->p = r.str();
    else if (f == 3 && w == 2) c->n = r.str();
    else if (f == 4 && w == 1) c->val = r.fixed64_double();
    else if (f == 5 && w == 1 && has_acm) c->acm = r.fixed64_double();
    // extension fields of Vsrc (not in the reference's spice21.proto:29-35, which is DC / acm only): 6 = wave kind (varint),
    // 7 = wave parameters (repeated double, packed or not)
    else if (f == 6 && w == 0 && has_acm) c->wave_kind = (int)r.varint();
    else if (f == 7 && w == 1 && has_acm) { const double v = r.fixed64_double(); if (nwave < 7) c->wave[nwave++] = v; }
    else if (f == 7 && w == 2 && has_acm) { PbReader pr = r.sub(); while (!pr.done()) { const double v = pr.fixed64_double(); if (nwave < 7) c->wave[nwave++] = v; } }
    else r.skip(w);
  }
}
inline void diode(PbReader r, CompSpec* c) {  // spice21.proto:70-76
  uint32_t f, w;
  while (r.next(&f, &w)) {
    if (w != 2) { r.skip(w); continue; }
    switch (f) {
      case 1: c->name = r.str(); break;
      case 2: c->p = r.str(); break;
      case 3: c->n = r.str(); break;
      case 4: c->model = r.str(); break;
      case 5: c->params = r.str(); break;
      default: r.skip(w);
    }
  }
}
inline void mos(PbReader r, CompSpec* c) {  // mos.proto:15-20, 8-13
  uint32_t f, w;
  while (r.next(&f, &w)) {
    if (w != 2) { r.skip(w); continue; }
    switch (f) {
      case 1: c->name = r.str(); break;
      case 2: c->model = r.str(); break;
      case 3: c->params = r.str(); break;
      case 4: {
        PbReader pr = r.sub();
        uint32_t pf, pw;
        while (pr.next(&pf, &pw)) {
          if (pw != 2) { pr.skip(pw); continue; }
          switch (pf) {
            case 1: c->d = pr.str(); break;
            case 2: c->g = pr.str(); break;
            case 3: c->s = pr.str(); break;
            case 4: c->b = pr.str(); break;
            default: pr.skip(pw);
          }
        }
        break;
      }
      default: r.skip(w);
    }
  }
}
inline void module_instance(PbReader r, CompSpec* c) {  // spice21.proto:103-108
  uint32_t f, w;
  while (r.next(&f, &w)) {
    if (f == 1 && w == 2) c->name = r.str();
    else if (f == 2 && w == 2) c->module = r.str();
    else if (f == 3 && w == 2) {  // map<string,string> entry {key=1, value=2}
      PbReader e = r.sub();
      std::string k, v;
      uint32_t ef, ew;
      while (e.next(&ef, &ew)) {
        if (ef == 1 && ew == 2) k = e.str();
        else if (ef == 2 && ew == 2) v = e.str();
        else e.skip(ew);
      }
      c->ports.push_back({k, v});
    } else r.skip(w);
  }
}
// Instance (spice21.proto:80-90). Returns false when the oneof is empty ("Invalid Component", circuit.rs:326).
inline bool instance(PbReader r, CompSpec* c) {
  uint32_t f, w;
  bool got = false;
  while (r.next(&f, &w)) {
    if (w != 2 || f < 1 || f > 7) { r.skip(w); continue; }
    PbReader s = r.sub();
    *c = CompSpec();  // last one wins, as for any protobuf oneof
    got = true;
    switch (f) {
      case 1: c->kind = CK_R; two_term(s, c, false); break;
      case 2: c->kind = CK_C; two_term(s, c, false); break;
      case 3: c->kind = CK_MOS; mos(s, c); break;
      case 4: c->kind = CK_I; two_term(s, c, false); break;
      case 5: c->kind = CK_V; two_term(s, c, true); break;
      case 6: c->kind = CK_D; diode(s, c); break;
      case 7: c->kind = CK_X; module_instance(s, c); break;
    }
  }
  return got;
}
inline ModuleSpec module_def(PbReader r) {  // spice21.proto:92-101
  ModuleSpec m;
  uint32_t f, w;
  while (r.next(&f, &w)) {
    if (f == 1 && w == 2) m.name = r.str();
    else if (f == 2 && w == 2) m.ports.push_back(r.str());
    else if (f == 4 && w == 2) m.signals.push_back(r.str());
    else if (f == 5 && w == 2) {
      CompSpec c;
      if (!instance(r.sub(), &c)) throw S21Error(ST_INVALID, "Invalid Comp!!!");  // elab.rs:235
      m.comps.push_back(c);
    } else r.skip(w);
  }
  return m;
}
// A message whose fields are `name` (string) plus DoubleValue wrappers; `names[f]` gives the key of field f.
inline std::string wrapped_params(PbReader r, uint32_t name_field, const std::map<uint32_t, const char*>& names, ParamBag* bag,
                                  MosModelSpec* mm = nullptr, uint32_t type_field = 0, uint32_t tpg_field = 0) {
  std::string name;
  uint32_t f, w;
  while (r.next(&f, &w)) {
    if (f == name_field && w == 2) name = r.str();
    else if (mm && f == type_field && w == 0) mm->mos_type = (int)r.varint();
    else if (mm && tpg_field && f == tpg_field && w == 2) { mm->has_tpg = true; mm->tpg = (long)read_int_value(r.sub()); }
    else if (w == 2 && names.count(f)) bag->kv[names.at(f)] = read_double_value(r.sub());
    else r.skip(w);
  }
  return name;
}
inline const std::map<uint32_t, const char*>& mos1_model_fields() {  // mos.proto:43-76
  static const std::map<uint32_t, const char*> m = {
      {3, "vt0"}, {4, "kp"}, {5, "gamma"}, {6, "phi"}, {7, "lambda"}, {8, "rd"}, {9, "rs"}, {10, "cbd"}, {11, "cbs"}, {12, "is"},
      {13, "pb"}, {14, "cgso"}, {15, "cgdo"}, {16, "cgbo"}, {17, "rsh"}, {18, "cj"}, {19, "mj"}, {20, "cjsw"}, {21, "mjsw"},
      {22, "js"}, {23, "tox"}, {24, "ld"}, {25, "u0"}, {26, "fc"}, {27, "nsub"}, {29, "nss"}, {30, "tnom"}, {31, "kf"}, {32, "af"}};
  return m;
}
inline const std::map<uint32_t, const char*>& mos1_inst_fields() {  // mos.proto:22-41
  static const std::map<uint32_t, const char*> m = {{2, "m"}, {3, "l"}, {4, "w"}, {5, "a_d"}, {6, "a_s"}, {7, "pd"},
                                                    {8, "ps"}, {9, "nrd"}, {10, "nrs"}, {11, "temp"}};
  return m;
}
inline const std::map<uint32_t, const char*>& diode_model_fields() {  // spice21.proto:44-61
  static const std::map<uint32_t, const char*> m = {{2, "tnom"}, {3, "is"}, {4, "n"}, {5, "tt"}, {6, "vj"}, {7, "m"}, {8, "eg"}, {9, "xti"},
                                                    {10, "kf"}, {11, "af"}, {12, "fc"}, {13, "bv"}, {14, "ibv"}, {15, "rs"}, {16, "cj0"}};
  return m;
}
inline const std::map<uint32_t, const char*>& diode_inst_fields() {  // spice21.proto:63-68
  static const std::map<uint32_t, const char*> m = {{4, "area"}, {5, "temp"}};
  return m;
}
inline const std::map<uint32_t, const char*>& bsim4_inst_fields() {  // bsim4.proto:6-43 (DoubleValue and UInt64Value wrappers)
  static const std::map<uint32_t, const char*> m = {
      {1, "l"}, {2, "w"}, {3, "nf"}, {4, "sa"}, {5, "sb"}, {6, "sd"}, {7, "sca"}, {8, "scb"}, {9, "scc"}, {10, "sc"}, {11, "ad"},
      {12, "as"}, {13, "pd"}, {14, "ps"}, {15, "nrd"}, {16, "nrs"}, {19, "rbdb"}, {20, "rbsb"}, {21, "rbpb"}, {22, "rbps"},
      {23, "rbpd"}, {24, "delvto"}, {25, "xgw"}, {26, "ngcon"}};
  return m;
}
inline const std::map<uint32_t, const char*>& bsim4_inst_int_fields() {
  static const std::map<uint32_t, const char*> m = {{17, "min"}, {18, "rgeomod"}, {27, "trnqsmod"}, {28, "acnqsmod"},
                                                    {29, "rbodymod"}, {30, "rgatemod"}, {31, "geomod"}};
  return m;
}

// Def (spice21.proto:114-124) -> CktSpec tables
inline void def(PbReader r, CktSpec* cs) {
  uint32_t f, w;
  bool got = false;
  while (r.next(&f, &w)) {
    if (w != 2 || f < 1 || f > 7) { r.skip(w); continue; }
    PbReader s = r.sub();
    got = true;
    switch (f) {
      case 1: { ModuleSpec m = module_def(s); cs->modules[m.name] = m; break; }
      case 2: { ParamBag b; std::string n = wrapped_params(s, 1, diode_model_fields(), &b); cs->diode_models[n] = b; break; }
      case 3: { ParamBag b; std::string n = wrapped_params(s, 1, diode_inst_fields(), &b); cs->diode_insts[n] = b; break; }
      case 4: {  // Bsim4Model: only mos_type (=1) and name (=900) exist on the reference's wire (bsim4.proto:45-50, 930)
        // EXTENSION: field 901, repeated { string name = 1; double value = 2 } — the model card's parameters by
        // their `Bsim4ModelSpecs` field names (the 876 commented-out fields of bsim4.proto, carried as a list instead of one
        // numbered field each). The reference's decoder skips it; here it fills the same bag s21_ckt_define("bsim4model") does.
        MosModelSpec m;
        std::string n;
        uint32_t ff, ww;
        while (s.next(&ff, &ww)) {
          if (ff == 900 && ww == 2) n = s.str();
          else if (ff == 1 && ww == 0) m.mos_type = (int)s.varint();
          else if (ff == 901 && ww == 2) {
            PbReader q = s.sub();
            std::string key;
            double val = 0.0;
            uint32_t f3, w3;
            while (q.next(&f3, &w3)) {
              if (f3 == 1 && w3 == 2) key = q.str();
              else if (f3 == 2 && w3 == 1) val = q.fixed64_double();
              else q.skip(w3);
            }
            if (!key.empty()) m.p.kv[key] = val;
          } else s.skip(ww);
        }
        cs->bsim4_models[n] = m;
        break;
      }
      case 5: {
        ParamBag b;
        std::string name;
        uint32_t ff, ww;
        while (s.next(&ff, &ww)) {
          if (ff == 40 && ww == 2) name = s.str();
          else if (ww == 2 && bsim4_inst_fields().count(ff)) b.kv[bsim4_inst_fields().at(ff)] = read_double_value(s.sub());
          else if (ww == 2 && bsim4_inst_int_fields().count(ff)) b.kv[bsim4_inst_int_fields().at(ff)] = (double)read_int_value(s.sub());
          else s.skip(ww);
        }
        cs->bsim4_insts[name] = b;
        break;
      }
      case 6: { MosModelSpec m; std::string n = wrapped_params(s, 1, mos1_model_fields(), &m.p, &m, 2, 28); cs->mos1_models[n] = m; break; }
      case 7: { ParamBag b; std::string n = wrapped_params(s, 1, mos1_inst_fields(), &b); cs->mos1_insts[n] = b; break; }
    }
  }
  if (!got) throw S21Error(ST_INVALID, "called `Option::unwrap()` on a `None` value (Def.defines)");  // circuit.rs:293
}

// Circuit (spice21.proto:128-133)
inline CktSpec circuit(PbReader r) {
  CktSpec cs;
  uint32_t f, w;
  while (r.next(&f, &w)) {
    if (f == 1 && w == 2) cs.name = r.str();
    else if (f == 2 && w == 2) cs.signals.push_back(r.str());
    else if (f == 3 && w == 2) def(r.sub(), &cs);
    else if (f == 4 && w == 2) {
      CompSpec c;
      if (!instance(r.sub(), &c)) throw S21Error(ST_OTHER, "Invalid Component");  // circuit.rs:326
      cs.comps.push_back(c);
    } else r.skip(w);
  }
  return cs;
}

struct SimOptionsPb { double v[5]; };  // NaN = not given (spice21.proto:136-142)
inline SimOptionsPb sim_options(PbReader r) {
  SimOptionsPb o;
  for (double& x : o.v) x = NAN;
  uint32_t f, w;
  while (r.next(&f, &w)) {
    if (w == 2 && f >= 1 && f <= 5) o.v[f - 1] = read_double_value(r.sub());
    else r.skip(w);
  }
  return o;
}

}  // namespace pb
}  // namespace s21
