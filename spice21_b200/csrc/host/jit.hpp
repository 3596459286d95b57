// Circuit-specialised Newton kernel for small real circuits (dcop / tran), compiled at run time with NVRTC.
//
// Why: for circuits like the bench workload (N = 9, 8 devices) the generic kernels spend ~80x more instructions than
// the arithmetic needs — they interpret index tables (stamp positions, L+U slots, level schedules) and synchronise
// lanes that co-operate on one instance. Here the symbolic phase's result is compiled INTO the kernel: one thread owns
// one instance, its workspace column sits in shared memory (conflict-free: entry k of thread t at W[k*TPB + t]), every
// stamp position / L+U slot / update triple is a literal, devices are evaluated by the very same `load_*` functions
// (kernels/devices.cuh, included by the generated source), and there is no synchronisation at all. Operation order is
// that of the direct kernel (kernels/newton.cu::newton_solve), which follows the reference (analysis.rs:169-210,
// sparse21/mod.rs:298-327, 865-979), so results are bit-identical to the other kernels.
//
// libnvrtc / libcuda are loaded lazily with dlopen: the library must still load (and say "no CUDA device") on machines
// without a driver. If NVRTC is unavailable the caller falls back to the hybrid kernel.
#pragma once
#include <dlfcn.h>

#include <cstdio>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "flatten.hpp"
#include "symbolic.hpp"

namespace s21 {
namespace jit {

// ---------------------------------------------------------------------------------------------- dynamic driver / NVRTC
struct Api {
  void *h_nvrtc = nullptr, *h_cuda = nullptr;
  int (*nvrtcCreateProgram)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*nvrtcCompileProgram)(void*, int, const char* const*) = nullptr;
  int (*nvrtcGetProgramLogSize)(void*, size_t*) = nullptr;
  int (*nvrtcGetProgramLog)(void*, char*) = nullptr;
  int (*nvrtcGetCUBINSize)(void*, size_t*) = nullptr;
  int (*nvrtcGetCUBIN)(void*, char*) = nullptr;
  int (*nvrtcDestroyProgram)(void**) = nullptr;
  int (*cuModuleLoadData)(void**, const void*) = nullptr;
  int (*cuModuleGetFunction)(void**, void*, const char*) = nullptr;
  int (*cuModuleUnload)(void*) = nullptr;
  int (*cuFuncSetAttribute)(void*, int, int) = nullptr;
  int (*cuLaunchKernel)(void*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**, void**) = nullptr;
  bool ok = false;
  std::string why;
};
inline Api& api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* n : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
      a.h_nvrtc = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (a.h_nvrtc) break;
    }
    for (const char* n : {"libcuda.so.1", "libcuda.so"}) {
      a.h_cuda = dlopen(n, RTLD_NOW | RTLD_LOCAL);
      if (a.h_cuda) break;
    }
    if (!a.h_nvrtc) { a.why = "libnvrtc not found"; return; }
    if (!a.h_cuda) { a.why = "libcuda not found"; return; }
    bool all = true;
    auto sym = [&](void* h, const char* name, auto& fp) {
      void* p = dlsym(h, name);
      if (!p) { all = false; a.why = std::string("missing symbol ") + name; }
      fp = reinterpret_cast<typename std::remove_reference<decltype(fp)>::type>(p);
    };
    sym(a.h_nvrtc, "nvrtcCreateProgram", a.nvrtcCreateProgram);
    sym(a.h_nvrtc, "nvrtcCompileProgram", a.nvrtcCompileProgram);
    sym(a.h_nvrtc, "nvrtcGetProgramLogSize", a.nvrtcGetProgramLogSize);
    sym(a.h_nvrtc, "nvrtcGetProgramLog", a.nvrtcGetProgramLog);
    sym(a.h_nvrtc, "nvrtcGetCUBINSize", a.nvrtcGetCUBINSize);
    sym(a.h_nvrtc, "nvrtcGetCUBIN", a.nvrtcGetCUBIN);
    sym(a.h_nvrtc, "nvrtcDestroyProgram", a.nvrtcDestroyProgram);
    sym(a.h_cuda, "cuModuleLoadData", a.cuModuleLoadData);
    sym(a.h_cuda, "cuModuleGetFunction", a.cuModuleGetFunction);
    sym(a.h_cuda, "cuModuleUnload", a.cuModuleUnload);
    sym(a.h_cuda, "cuFuncSetAttribute", a.cuFuncSetAttribute);
    sym(a.h_cuda, "cuLaunchKernel", a.cuLaunchKernel);
    a.ok = all;
  });
  return a;
}

// directory holding this library's CUDA sources (…/spice21_b200/csrc): the generated source #includes kernels/devices.cuh
inline std::string csrc_dir() {
  if (const char* e = std::getenv("S21_JIT_CSRC")) return e;
  Dl_info info;
  if (dladdr((void*)&csrc_dir, &info) && info.dli_fname) {
    std::string p = info.dli_fname;
    size_t s = p.find_last_of('/');
    return (s == std::string::npos ? std::string(".") : p.substr(0, s)) + "/csrc";
  }
  return "csrc";
}

struct Kernel {
  void* module = nullptr;
  void* fn = nullptr;
  int tpb = 128;
  int inst_per_cta = 128;
  size_t smem = 0;
  bool tran = false;
};

constexpr int TPB = 128;

// A circuit is eligible when every device has a real load function the generated source can inline cheaply and the
// workspace column (x, rhs, c, lu) of 128 threads fits in shared memory.
inline bool eligible(const FlatCkt& flat, const Plan& P, size_t max_smem) {
  for (const FlatDev& d : flat.devs)
    if (d.type == DT_BSIM4) return false;
  const size_t words = (size_t)P.nnzLU + 3 * (size_t)P.N;
  if (words * 8 * TPB > max_smem) return false;
  size_t ops = P.upd_t.size() + P.l_slot.size() + P.colidx.size();
  return ops <= 4000 && flat.devs.size() <= 256;
}

inline std::string source(const FlatCkt& flat, const Plan& P, const std::vector<int>& itab, const std::vector<int>& pcode, bool tran, int tpb = TPB) {
  std::ostringstream o;
  const int N = P.N, NNZ = P.nnzLU;
  const int OFF_X = 0, OFF_RHS = N, OFF_C = 2 * N, OFF_LU = 3 * N;
  o << "#define S21_JIT 1\n#include \"kernels/devices.cuh\"\nnamespace s21 {\n";
  o << "#define TPB " << tpb << "\n";
  o << "struct JBase {\n  const double* pval; size_t pinst; double* sop; double* sguess; size_t sstride; double* W;\n"
       "  int mode; double dt, gmin, omega;\n"
       "  __device__ __forceinline__ double volt(int var) const { return var < 0 ? 0.0 : W[(" << OFF_X << " + var) * TPB]; }\n};\n";
  for (size_t k = 0; k < flat.devs.size(); k++) {
    const FlatDev& d = flat.devs[k];
    o << "struct E" << k << " : JBase {\n";
    o << "  __device__ __forceinline__ int node(int k) const { switch (k) {";
    for (int j = 0; j < d.n_itab; j++) o << " case " << j << ": return " << itab[(size_t)d.itab_off + (size_t)j] << ";";
    o << " default: return -1; } }\n";
    o << "  __device__ __forceinline__ double par(int k) const { switch (k) {";
    for (int j = 0; j < d.n_par; j++) {
      const int c = pcode[(size_t)d.par_off + (size_t)j];
      o << " case " << j << ": return pval[" << (c >> 1) << ((c & 1) ? " + pinst" : "") << "];";
    }
    o << " default: return 0.0; } }\n";
    o << "  __device__ __forceinline__ double op(int k) const { return sop[(size_t)(" << d.state_off << " + k) * sstride]; }\n";
    o << "  __device__ __forceinline__ double guess(int k) const { return sguess[(size_t)(" << d.state_off << " + k) * sstride]; }\n";
    o << "  __device__ __forceinline__ void set_guess(int k, double v) { sguess[(size_t)(" << d.state_off << " + k) * sstride] = v; }\n";
    o << "  __device__ __forceinline__ void add_g_at(int pos, double v) { const int h = node(pos); if (h >= 0) { double* a = W + (" << OFF_LU
      << " + h) * TPB; *a = s_add(*a, v); } }\n";
    o << "  __device__ __forceinline__ void add_b_at(int pos, double v) { const int h = node(pos); if (h >= 0) { double* a = W + (" << OFF_RHS
      << " + h) * TPB; *a = s_add(*a, v); } }\n";
    o << "  __device__ __forceinline__ void add_g_dup(int pos, int, double v) { add_g_at(pos, v); }\n};\n";
  }
  auto W = [&](int off, int k) { return "W[" + std::to_string((off + k)) + " * TPB]"; };
  // ---- one Newton solve (newton.cu::newton_solve); returns the S21 status
  o << "__device__ __forceinline__ int solve_one(JBase base, int n_state, double reltol, double iabstol, int* n_solves, int* n_loads) {\n"
       "  double* W = base.W;\n  bool dx_ok = true;\n  for (int iter = 0; iter < 100; iter++) {\n";
  for (int k = 0; k < NNZ; k++) o << "    " << W(OFF_LU, k) << " = 0.0;\n";
  for (int k = 0; k < N; k++) o << "    " << W(OFF_RHS, k) << " = 0.0;\n";
  for (size_t k = 0; k < flat.devs.size(); k++) {
    const char* fn = nullptr;
    switch (flat.devs[k].type) {
      case DT_R: fn = "load_resistor"; break;
      case DT_C: fn = "load_capacitor"; break;
      case DT_I: fn = "load_isrc"; break;
      case DT_V: fn = "load_vsrc"; break;
      case DT_DIODE: fn = "load_diode"; break;
      case DT_MOS0: fn = "load_mos0"; break;
      case DT_MOS1: fn = "load_mos1"; break;
      default: fn = nullptr;
    }
    if (!fn) { o << "    return 6;\n"; continue; }
    o << "    { E" << k << " e; static_cast<JBase&>(e) = base; " << fn << "(e); }\n";
  }
  o << "    *n_loads += 1;\n    bool res_ok = true;\n";
  for (int r = 0; r < N; r++) {
    o << "    { double acc = 0.0;\n";
    for (int s = P.rowptr[(size_t)r]; s < P.rowptr[(size_t)r + 1]; s++)
      o << "      acc = s_add(acc, s_mul(" << W(OFF_LU, s) << ", " << W(OFF_X, P.col_i2e[(size_t)P.colidx[(size_t)s]]) << "));\n";
    o << "      const double rv = s_sub(" << W(OFF_RHS, P.row_i2e[(size_t)r]) << ", acc); " << W(OFF_C, r)
      << " = rv; res_ok = res_ok && !(s_abs(rv) > iabstol); }\n";
  }
  o << "    if (dx_ok && res_ok) {\n"
       "      for (int k = 0; k < n_state; k++) base.sop[(size_t)k * base.sstride] = base.sguess[(size_t)k * base.sstride];\n"
       "      return 0;\n    }\n";
  for (int k = 0; k + 1 < N; k++) {
    o << "    { const double piv = " << W(OFF_LU, P.diag_slot[(size_t)k]) << "; if (piv == 0.0) return 2;\n";
    for (int j = P.l_off[(size_t)k]; j < P.l_off[(size_t)k + 1]; j++)
      o << "      " << W(OFF_LU, P.l_slot[(size_t)j]) << " = s_div(" << W(OFF_LU, P.l_slot[(size_t)j]) << ", piv);\n";
    o << "    }\n";
    for (int j = P.upd_off[(size_t)k]; j < P.upd_off[(size_t)k + 1]; j++)
      o << "    " << W(OFF_LU, P.upd_t[(size_t)j]) << " = s_sub(" << W(OFF_LU, P.upd_t[(size_t)j]) << ", s_mul(" << W(OFF_LU, P.upd_u[(size_t)j])
        << ", " << W(OFF_LU, P.upd_l[(size_t)j]) << "));\n";
  }
  for (int k = 0; k < N; k++) {
    if (P.l_off[(size_t)k] == P.l_off[(size_t)k + 1]) continue;
    o << "    { const double ck = " << W(OFF_C, k) << "; if (!(ck == 0.0)) {\n";
    for (int j = P.l_off[(size_t)k]; j < P.l_off[(size_t)k + 1]; j++)
      o << "      " << W(OFF_C, P.l_row[(size_t)j]) << " = s_sub(" << W(OFF_C, P.l_row[(size_t)j]) << ", s_mul(ck, " << W(OFF_LU, P.l_slot[(size_t)j])
        << "));\n";
    o << "    } }\n";
  }
  for (int k = N - 1; k >= 0; k--) {
    const int ds = P.diag_slot[(size_t)k], en = P.rowptr[(size_t)k + 1];
    o << "    { double ck = " << W(OFF_C, k) << ";\n";
    for (int s = ds + 1; s < en; s++) o << "      ck = s_sub(ck, s_mul(" << W(OFF_C, P.colidx[(size_t)s]) << ", " << W(OFF_LU, s) << "));\n";
    o << "      " << W(OFF_C, k) << " = s_div(ck, " << W(OFF_LU, ds) << "); }\n";
  }
  o << "    *n_solves += 1;\n    double max_abs = 0.0;\n";
  for (int k = 0; k < N; k++) o << "    { const double a = s_abs(" << W(OFF_C, P.col_e2i[(size_t)k]) << "); if (a > max_abs) max_abs = a; }\n";
  o << "    const bool limit = max_abs > 1.0;\n    dx_ok = true;\n";
  for (int k = 0; k < N; k++)
    o << "    { double dxk = " << W(OFF_C, P.col_e2i[(size_t)k]) << "; if (limit) dxk = s_scale(dxk, 1.0, max_abs); " << W(OFF_X, k) << " = s_add("
      << W(OFF_X, k) << ", dxk); dx_ok = dx_ok && !(s_abs(dxk) > reltol); }\n";
  o << "  }\n  return 1;\n}\n";
  // ---- kernel: dcop, or the OP-committed fixed-step transient loop (newton.cu::k_dcop / k_tran)
  o << "extern \"C\" __global__ void __launch_bounds__(TPB) k_jit(const double* __restrict__ pval, double* gx, double* st_op, double* st_guess,\n"
       "    int* status, int* iters, int* loads, size_t stride, size_t st_stride, int B, int n_state, int mode, double gmin, double dt,\n"
       "    double reltol, double iabstol, int cold, int T_points, int n_save, const int* __restrict__ save_vars, double* wave) {\n"
       "  extern __shared__ double sm[];\n"
       "  const size_t inst = (size_t)blockIdx.x * TPB + threadIdx.x;\n  if (inst >= (size_t)B) return;\n"
       "  double* W = sm + threadIdx.x;\n"
       "  JBase base; base.pval = pval; base.pinst = inst; base.sop = st_op + inst; base.sguess = st_guess + inst; base.sstride = st_stride;\n"
       "  base.W = W; base.mode = mode; base.dt = dt; base.gmin = gmin; base.omega = 0.0;\n";
  o << "  if (cold) {\n    for (int k = 0; k < " << N << "; k++) W[k * TPB] = 0.0;\n"
       "    for (int k = 0; k < n_state; k++) { base.sop[(size_t)k * st_stride] = 0.0; base.sguess[(size_t)k * st_stride] = 0.0; }\n"
       "  } else {\n    for (int k = 0; k < " << N << "; k++) W[k * TPB] = gx[(size_t)k * stride + inst];\n  }\n";
  o << "  int ns = 0, nl = 0;\n";
  if (!tran) {
    o << "  const int st = solve_one(base, n_state, reltol, iabstol, &ns, &nl);\n"
         "  for (int k = 0; k < " << N << "; k++) gx[(size_t)k * stride + inst] = W[k * TPB];\n"
         "  status[inst] = st;\n  iters[inst] = (cold ? 0 : iters[inst]) + ns;\n  loads[inst] = (cold ? 0 : loads[inst]) + nl;\n}\n";
  } else {
    o << "  int st = status[inst];\n"
         "  for (int s = 0; s < n_save; s++) wave[(size_t)s * stride + inst] = W[save_vars[s] * TPB];\n"
         "  for (int tp = 1; tp < T_points; tp++) {\n"
         "    if (st == 0) st = solve_one(base, n_state, reltol, iabstol, &ns, &nl);\n"
         "    for (int s = 0; s < n_save; s++)\n"
         "      wave[((size_t)tp * n_save + s) * stride + inst] = st == 0 ? W[save_vars[s] * TPB] : __longlong_as_double(0x7ff8000000000000LL);\n"
         "  }\n"
         "  for (int k = 0; k < " << N << "; k++) gx[(size_t)k * stride + inst] = W[k * TPB];\n"
         "  status[inst] = st;\n  iters[inst] += ns;\n  loads[inst] += nl;\n}\n";
  }
  o << "}  // namespace s21\n";
  return o.str();
}

// Compile (or fetch from the per-process cache) the kernel for `src`. Returns false with `err` set when NVRTC is not usable.
inline bool compile(const std::string& src, bool tran, int tpb, size_t smem, Kernel* out, std::string* err) {
  Api& a = api();
  if (!a.ok) { *err = a.why; return false; }
  static std::mutex mu;
  static std::map<std::string, Kernel> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(src);
  if (it != cache.end()) { *out = it->second; return true; }
  void* prog = nullptr;
  if (a.nvrtcCreateProgram(&prog, src.c_str(), "s21_jit.cu", 0, nullptr, nullptr) != 0) { *err = "nvrtcCreateProgram failed"; return false; }
  const std::string inc = "-I" + csrc_dir();
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-fmad=false", "-lineinfo", inc.c_str()};
  const int rc = a.nvrtcCompileProgram(prog, 5, opts);
  if (rc != 0) {
    size_t n = 0;
    a.nvrtcGetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) a.nvrtcGetProgramLog(prog, &log[0]);
    a.nvrtcDestroyProgram(&prog);
    *err = "NVRTC compile failed: " + log.substr(0, 2000);
    return false;
  }
  size_t n = 0;
  a.nvrtcGetCUBINSize(prog, &n);
  std::vector<char> cubin(n);
  a.nvrtcGetCUBIN(prog, cubin.data());
  a.nvrtcDestroyProgram(&prog);
  Kernel k;
  k.tran = tran;
  k.tpb = tpb;
  k.smem = smem;
  if (a.cuModuleLoadData(&k.module, cubin.data()) != 0) { *err = "cuModuleLoadData failed"; return false; }
  if (a.cuModuleGetFunction(&k.fn, k.module, "k_jit") != 0) { *err = "cuModuleGetFunction failed"; return false; }
  if (a.cuFuncSetAttribute(k.fn, 8 /* CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES */, (int)k.smem) != 0) {
    *err = "cuFuncSetAttribute(max dynamic smem) failed";
    return false;
  }
  cache[src] = k;
  *out = k;
  return true;
}

}  // namespace jit
}  // namespace s21
