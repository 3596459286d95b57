// Circuit-specialised TEAM kernel for small real circuits (dcop / tran), compiled at run time with NVRTC.
//
// Why a second specialised shape: one thread per instance (host/jit.hpp) has the fewest instructions but the longest
// dependent chain, so it only pays once there are enough warps to overlap chains (~12 k instances). At the BASELINE
// batch size (8192) the generic hybrid kernel (kernels/hybrid.cu) was the fastest, but it INTERPRETS the plan: every
// gather, every L+U operation and every substitution step loads its indices from tables and round-trips its values
// through shared memory (ncu: ~9 cycles per instruction per warp; assembly + residual + LU + substitution + update are
// 80 % of an iteration). This generator keeps the hybrid kernel's thread mapping —
//   * device evaluation: lane = instance, warp = device (full lanes on the expensive Mos1 evaluation),
//   * linear algebra: a team of 8 lanes per instance, 4 instances per warp, warp-synchronous,
// — and compiles the plan INTO the linear algebra: team member j owns pivoted rows j, j+8 (N <= 16) and keeps them in
// REGISTERS as named scalars (one per column of the union pattern); pivot rows, substitution values and x travel
// between members with warp shuffles; the L+U pattern, the level structure and the update triples are literals and
// lane masks. Only the staged assembly still reads a table (the gather lists differ per lane). Two block barriers per
// Newton iteration, as in the hybrid kernel.
//
// Arithmetic per value is in the reference's order (analysis.rs:169-210; sparse21/mod.rs:298-327 residual, 865-919
// elimination, 947-979 substitution): per element the stamps are summed in component/push order, Schur updates arrive
// in ascending pivot order, the back-substitution row sum runs over ascending columns — results are bit-identical to
// the other kernels (tests/test_gpu.py::test_kernel_variants_bit_identical).
#pragma once
#include <cstring>

#include "jit.hpp"
#include "staging.hpp"

namespace s21 {
namespace jit {

constexpr int TM_P = 36;     // padded instance stride of the shared-memory columns (as kernels/hybrid.cu)
constexpr int TM_GI = 32;    // instances per CTA
constexpr int TM_MAXN = 16;  // rows per instance the register-resident linear algebra is generated for
// Lanes per instance in the linear-algebra phase: 8 (4 instances per warp, 8 warps per CTA; rows beyond 8 become a
// second register set), 10 (3 instances per warp) or 16 (2 instances per warp, 16 warps per CTA; one row per member).
inline bool team_wp(bool tran, size_t B);
inline int team_lpi(int N, int n_heavy = 0, size_t B = 0, bool tran = false) {
  if (const char* e = std::getenv("S21_TEAM_LPI")) {
    const int v = std::atoi(e);
    if (v == 1 || v == 2 || v == 4 || v == 8 || v == 10 || v == 16) return v;
  }
  // Fewer lanes per instance = more independent rows per lane (instruction-level parallelism inside the one dependent
  // chain a Newton iteration is) and fewer warps per SM for the same batch. Measured on B200 (profiles/r01t_team_shapes.txt),
  // C2 dcop (N = 9, 2 Mos1) at 8192 instances: 16 lanes 0.160 ms, 10 lanes 0.123, 8 lanes 0.116, 4 lanes 0.100,
  // 2 lanes 0.085, 1 lane 0.106; C1 transient (N = 7, 6 Mos1, 200 points): 8 lanes 13.3 ms, 4 lanes 12.1, 2 lanes 12.7,
  // 1 lane 15.1 — the evaluation phase wants a warp per expensive device, so device-heavy circuits keep 4 lanes.
  // (2 lanes were measured up to N = 9, i.e. five register sets per lane; beyond ten rows stay with 4 lanes until measured)
  // Small shards of a strong-scaled sweep (<= 2048 instances per GPU) leave most SMs with one CTA: there the warp-private
  // loop with 8 lanes per instance and 16-instance CTAs is the shortest chain (profiles/r02a_wp_sweep.txt: 0.075 ms against
  // 0.078-0.081 at 1024 / 2048 instances of C2).
  if (!tran && B > 0 && B <= 2048 && team_wp(tran, B) && n_heavy <= 2 && N <= 10) return 8;
  return (n_heavy <= 2 && N <= 10) ? 2 : 4;
}
inline int team_heavy_devices(const FlatCkt& flat) {
  int n = 0;
  for (const FlatDev& d : flat.devs) n += (d.type == DT_MOS1 || d.type == DT_DIODE || d.type == DT_MOS0) ? 1 : 0;
  return n;
}
// S21_TEAM_PROFILE=1 adds clock64() probes at the phase boundaries of warp 0 (evaluates the heaviest device) and of the
// last warp (idle during evaluation) of CTA 0 and prints the per-phase cycle sums when the kernel ends (diagnostic only).
inline bool team_profile() { const char* e = std::getenv("S21_TEAM_PROFILE"); return e && std::atoi(e) != 0; }
// S21_TEAM_FAST=0 generates the exact (branching) linear-algebra text only; default is the branch-free text with the exact
// one as its redo path. Both give the same bits (tests/test_gpu.py::test_team_kernel_fast_and_exact_text_agree).
// Instances per CTA: a full warp of instances in the evaluation phase (30 with 3 instances per warp). Smaller CTAs that
// balance the SMs better (8192 instances: 293 CTAs of 28 -> 56 per SM instead of 64) measured the same 0.115 ms: the
// launch is one dependent chain per CTA, not a throughput problem. S21_TEAM_GI overrides (experiments).
inline int team_gi(size_t B, int n_sm, int lpi, bool tran = false) {
  const int ipw = 32 / lpi, gmax = TM_GI / ipw * ipw;
  if (const char* e = std::getenv("S21_TEAM_GI")) { const int v = std::atoi(e); if (v >= ipw && v <= gmax && v % ipw == 0) return v; }
  (void)n_sm;
  if (!tran && B > 0 && B <= 2048 && lpi == 8 && team_wp(tran, B)) return 16;  // see team_lpi
  return gmax;
}
// Warp-private loop (S21_TEAM_WP=1 / 0 forces it on / off for both analyses): a warp evaluates the devices of its own
// instances (lane = (instance, device group)), so an iteration needs no block barrier and same-type devices share one
// evaluation text; the CTA may then be a single warp (the shared-memory stride follows the instances per CTA), and the
// kernel leaves the host's result layout behind (no packing kernel before the D2H copy). Measured (profiles/r02a_wp_*,
// r02d_*): the C1-circuit transient x 8192 (6 Mos1, 4 lanes) goes from 12.08 to 8.98 ms, so transients take it by
// default; the C2 dcop x 8192 does not gain (0.083 ms either way back to back, and with the L2 flushed between steps the
// result rows it writes cost more than the packing kernel they replace: 0.0948 against 0.0887 ms), so a dcop of a
// full-size batch keeps the CTA-wide evaluation phase. Small shards (<= 2048 instances) take it for the 8-lane shape.
inline bool team_wp_env(int* forced) {
  const char* e = std::getenv("S21_TEAM_WP");
  if (!e) return false;
  *forced = std::atoi(e) != 0 ? 1 : 0;
  return true;
}
inline bool team_wp(bool tran = true, size_t B = 0) {
  int f = 0;
  if (team_wp_env(&f)) return f != 0;
  return tran || (B > 0 && B <= 2048);
}
inline bool team_fast() { const char* e = std::getenv("S21_TEAM_FAST"); return !e || std::atoi(e) != 0; }

struct TeamGather {
  std::vector<int> table;  // [steps][lpi] staging offsets (slot * TM_P) or the zero row
  int steps = 0;
};

inline size_t team_smem_bytes(const FlatCkt& flat, const Plan& P, int gather_steps) {
  const size_t ints = (size_t)TM_GI + (size_t)gather_steps * 16;
  const size_t ctrl = (ints * 4 + 15) / 16 * 16;
  return ctrl + 8 * (size_t)TM_P * ((size_t)P.N + (size_t)P.n_stage + 1 + 2 * (size_t)std::max(flat.n_state, 1));
}

inline bool team_eligible(const FlatCkt& flat, const Plan& P, size_t max_smem) {
  if (P.status != ST_OK || P.N > TM_MAXN || P.N < 2) return false;
  if (flat.devs.size() > 64) return false;
  for (const FlatDev& d : flat.devs)
    if (d.type == DT_BSIM4) return false;
  // the gather table is bounded by the stamp count
  return team_smem_bytes(flat, P, (int)P.asm_src.size() + 2 * TM_MAXN) <= max_smem;
}

inline std::string team_source(const FlatCkt& flat, const Plan& P, const StageInfo& si, const std::vector<int>& itab,
                               const std::vector<int>& pcode, bool tran, int TM_LPI, size_t* smem_out, int n_sm = 148, int GI = TM_GI, int* tpb_out = nullptr,
                               size_t B = 0) {
  std::ostringstream o;
  const bool XP = team_wp(tran, B);
  const int PSV = XP ? GI + 4 : TM_P;  // padded instance stride of the shared-memory columns (36 for a full CTA, as kernels/hybrid.cu)
  const int N = P.N, NST = P.n_stage, NSTATE = std::max(flat.n_state, 1), Q = (N + TM_LPI - 1) / TM_LPI;
  // Committed device state in HBM / L2 instead of shared memory (transients only). The state of a time loop lives on chip
  // for the whole launch; on the C1-circuit sweep that is 139 KB per 32-instance CTA — ONE CTA per SM, so 8192 instances =
  // 256 CTAs ran as two waves on 148 SMs (ncu: profiles/r02T_c1_full.txt; 2048 instances take half the time of 8192).
  // The committed copy (`op`) is read a few times per device evaluation and written once per accepted time point; leaving
  // it in its HBM column (read through L2: __ldcg, coherent with the warp's own stores) takes the CTA under half an SM's
  // shared memory — two CTAs per SM, one wave. Taken only when that is what it buys: a transient, a batch of more CTAs
  // than SMs, and a footprint that crosses the half-SM line with it. S21_TEAM_SOPG=0 / 1 forces it off / on.
  bool SOPG = false;  // decided below, once the size of the gather table (part of the footprint) is known
  const int IPW = 32 / TM_LPI;        // instances per warp in the linear-algebra phase
  const int NW_LA = GI / IPW;      // warps of the linear-algebra phase
  // S21_TEAM_NW adds warps that only take part in the evaluation phase and sit out the linear algebra (ri >= GI).
  int NW = NW_LA;  // measured: extra evaluation warps cost more at the barriers than they save (C2 0.086 -> 0.091 / 0.100 ms with 4 / 8)
  if (const char* e = std::getenv("S21_TEAM_NW")) { const int v = std::atoi(e); if (v >= NW_LA && v <= 16) NW = v; }
  if (tpb_out) *tpb_out = NW * 32;
  const bool POW2 = (TM_LPI & (TM_LPI - 1)) == 0;  // otherwise the last lanes of a warp belong to no instance (j >= TM_LPI)
  const unsigned FULLSET = (1u << TM_LPI) - 1u;
  unsigned IMASK = 0u;  // lanes of instance 0 of a warp
  for (int jj = 0; jj < TM_LPI; jj++) IMASK |= 1u << (IPW * jj);
  auto qof = [&](int r) { return r / TM_LPI; };
  auto jof = [&](int r) { return r % TM_LPI; };
  // ---- pattern bookkeeping in pivoted coordinates
  std::vector<std::vector<int>> slot((size_t)N, std::vector<int>((size_t)N, -1));
  for (int r = 0; r < N; r++)
    for (int s = P.rowptr[(size_t)r]; s < P.rowptr[(size_t)r + 1]; s++) slot[(size_t)r][(size_t)P.colidx[(size_t)s]] = s;
  // M[q][c]: members of set q whose row has an entry in column c
  std::vector<std::vector<unsigned>> M((size_t)Q, std::vector<unsigned>((size_t)N, 0u));
  for (int r = 0; r < N; r++)
    for (int c = 0; c < N; c++)
      if (slot[(size_t)r][(size_t)c] >= 0) M[(size_t)qof(r)][(size_t)c] |= 1u << jof(r);
  // LM[q][k]: members of set q whose row is below the diagonal in column k (the L entries of pivot k)
  std::vector<std::vector<unsigned>> LM((size_t)Q, std::vector<unsigned>((size_t)N, 0u));
  for (int k = 0; k < N; k++)
    for (int jx = P.l_off[(size_t)k]; jx < P.l_off[(size_t)k + 1]; jx++) {
      const int r = P.l_row[(size_t)jx];
      LM[(size_t)qof(r)][(size_t)k] |= 1u << jof(r);
    }
  auto A = [&](int q, int c) { return "a" + std::to_string(q) + "_" + std::to_string(c); };
  auto mask_test = [&](unsigned m) {
    char buf[32];
    std::snprintf(buf, sizeof buf, "0x%04xu", m);
    return std::string("(") + buf + " & jbit)";
  };
  // valid members of set q (rows < N)
  auto valid_mask = [&](int q) { const int n = std::min(TM_LPI, N - q * TM_LPI); return n >= TM_LPI ? FULLSET : ((1u << n) - 1u); };

  // ---- gather table: step-major, 8 members per step
  TeamGather G;
  const int zero_off = NST * PSV;
  std::ostringstream gath;  // generated gather code
  auto emit_gather = [&](const std::string& var, const std::vector<int>** lists) {
    size_t maxlen = 0;
    for (int j = 0; j < TM_LPI; j++)
      if (lists[j]) maxlen = std::max(maxlen, lists[j]->size());
    gath << "        double " << var << " = 0.0;\n";
    for (size_t p = 0; p < maxlen; p++) {
      for (int j = 0; j < TM_LPI; j++) G.table.push_back(lists[j] && p < lists[j]->size() ? (*lists[j])[p] * PSV : zero_off);
      gath << "        " << var << " = s_add(" << var << ", Sri[gtj[" << G.steps * TM_LPI << "]]);\n";
      G.steps++;
    }
  };
  std::vector<std::vector<int>> src((size_t)P.nnzLU + (size_t)N);
  for (size_t t = 0; t + 1 < P.asm_off.size(); t++)
    src[t].assign(P.asm_src.begin() + P.asm_off[t], P.asm_src.begin() + P.asm_off[t + 1]);
  for (int q = 0; q < Q; q++) {
    for (int c = 0; c < N; c++) {
      if (!M[(size_t)q][(size_t)c]) continue;
      const std::vector<int>* lists[16];
      for (int j = 0; j < TM_LPI; j++) {
        const int r = q * TM_LPI + j;
        lists[j] = (r < N && slot[(size_t)r][(size_t)c] >= 0) ? &src[(size_t)slot[(size_t)r][(size_t)c]] : nullptr;
      }
      emit_gather(A(q, c), lists);
    }
    const std::vector<int>* lists[16];
    for (int j = 0; j < TM_LPI; j++) {
      const int r = q * TM_LPI + j;
      lists[j] = r < N ? &src[(size_t)P.nnzLU + (size_t)P.row_i2e[(size_t)r]] : nullptr;
    }
    emit_gather("b" + std::to_string(q), lists);
  }

  {
    const size_t half_sm = (233472 - 2 * 1024) / 2;  // sm_100a: 228 KB per SM, 1 KB reserved per CTA
    const size_t slots_full = (size_t)N + (size_t)NST + 1 + 2 * (size_t)NSTATE, slots_red = slots_full - (size_t)NSTATE;
    const size_t ctrl = (((size_t)GI + G.table.size() + (size_t)TM_LPI) * 4 + 15) / 16 * 16;  // as ctrl_bytes below (the table gets one more row)
    const bool two_waves = B > 0 && (B + (size_t)GI - 1) / (size_t)GI > (size_t)std::max(n_sm, 1);
    SOPG = tran && two_waves && ctrl + 8 * (size_t)PSV * slots_full > half_sm && ctrl + 8 * (size_t)PSV * slots_red <= half_sm;
    if (const char* e = std::getenv("S21_TEAM_SOPG")) SOPG = tran && std::atoi(e) != 0;
  }
  // ---- source
  const bool prof = team_profile();
  const bool fast_la = team_fast();
  o << "#define S21_JIT 1\n#include \"kernels/devices.cuh\"\n";
  if (prof)
    o << "extern \"C\" int printf(const char*, ...);\n"
         "#define PH(n) if (lane == 0 && blockIdx.x == 0 && (warp == 0 || warp == " << NW - 1 << ")) { const long long t_ = clock64(); "
         "prof_s[(warp ? 16 : 0) + (n)] += t_ - t_last; t_last = t_; }\n";
  else
    o << "#define PH(n)\n";
  o << "namespace s21 {\n";
  o << "#define PS " << PSV << "\n#define FULLM 0xffffffffu\n#define BC(v, jj) __shfl_sync(FULLM, (v), base + " << IPW << " * (jj))\n";
  // idle lanes (j >= lanes per instance) read one row further: the row must be part of the emitted array (it used to be
  // appended after GT_G had been written out, so the copy into shared memory read TM_LPI ints past its end — found by
  // compute-sanitizer memcheck, tests/test_gpu.py::test_compute_sanitizer_clean)
  for (int jj = 0; jj < TM_LPI; jj++) G.table.push_back(zero_off);
  o << "__device__ const int GT_G[" << std::max<size_t>(G.table.size(), 1) << "] = {";
  for (size_t k = 0; k < G.table.size(); k++) o << (k ? "," : "") << G.table[k];
  if (G.table.empty()) o << "0";
  o << "};\n";
  o << "__device__ const int XO_G[" << Q * TM_LPI + 8 << "] = {";
  for (int k = 0; k < Q * TM_LPI + 8; k++) o << (k ? "," : "") << (k < N ? P.col_i2e[(size_t)k] * PSV : 0);
  o << "};\n";
  o << "struct JBase {\n  const double* pval; size_t pinst; double* sop; double* sguess; const double* X; double* S;" << (SOPG ? " size_t sops;" : "") << "\n"
       "  int mode" << (XP ? ", j" : "") << "; double dt, gmin, omega, time;\n"
       "  __device__ __forceinline__ double volt(int var) const { return var < 0 ? 0.0 : X[var * PS]; }\n};\n";
  // the Env base of the branch-free evaluation (kernels/devices.cuh math hooks): fast paths only, exceptions deferred
  o << "struct JFast : JBase {\n  bool dbad;\n"
       "  __device__ __forceinline__ double m_div(double a, double b) { return s_div_rf(a, b, s_rcp(b), dbad); }\n"
       "  __device__ __forceinline__ double m_sqrt(double a) { return s_sqrt_f(a, dbad); }\n"
       "  __device__ __forceinline__ double m_exp(double a) { return s_exp_f(a, dbad); }\n};\n";
  for (size_t k = 0; k < flat.devs.size(); k++) {
    const FlatDev& d = flat.devs[k];
    const int sto = si.stage_off[k];
    o << "template <class Base> struct E" << k << " : Base {\n  using Base::pval; using Base::pinst; using Base::sop; using Base::sguess; using Base::S;\n";
    o << "  __device__ __forceinline__ int node(int k) const { switch (k) {";
    for (int j = 0; j < d.n_itab; j++) o << " case " << j << ": return " << itab[(size_t)d.itab_off + (size_t)j] << ";";
    o << " default: return -1; } }\n";
    o << "  __device__ __forceinline__ double par(int k) const { switch (k) {";
    for (int j = 0; j < d.n_par; j++) {
      const int c = pcode[(size_t)d.par_off + (size_t)j];
      o << " case " << j << ": return __ldg(pval + " << (c >> 1) << ((c & 1) ? " + pinst" : "") << ");";
    }
    o << " default: return 0.0; } }\n";
    if (SOPG) o << "  using Base::sops;\n  __device__ __forceinline__ double op(int k) const { return __ldcg(sop + (size_t)(" << d.state_off << " + k) * sops); }\n";
    else o << "  __device__ __forceinline__ double op(int k) const { return sop[(" << d.state_off << " + k) * PS]; }\n";
    o << "  __device__ __forceinline__ double guess(int k) const { return sguess[(" << d.state_off << " + k) * PS]; }\n";
    o << "  __device__ __forceinline__ void set_guess(int k, double v) { sguess[(" << d.state_off << " + k) * PS] = v; }\n";
    o << "  __device__ __forceinline__ void add_g_at(int pos, double v) { S[(" << sto << " + pos) * PS] = v; }\n";
    o << "  __device__ __forceinline__ void add_b_at(int pos, double v) { S[(" << sto << " + pos) * PS] = v; }\n";
    o << "  __device__ __forceinline__ void add_g_dup(int, int dup, double v) { S[(" << sto << " + dup) * PS] = v; }\n};\n";
  }
  // ---- warp-private evaluation: slots of up to TM_LPI same-type devices, member j of a team evaluates slot[s][j]
  const bool WP = XP && POW2 && NW == NW_LA;
  std::vector<std::vector<int>> slots;
  if (WP) {
    std::vector<bool> used(flat.devs.size(), false);
    for (size_t a = 0; a < si.eval_order.size(); a++) {
      const int d0 = si.eval_order[a];
      if (used[(size_t)d0]) continue;
      std::vector<int> slot;
      for (size_t b2 = a; b2 < si.eval_order.size() && (int)slot.size() < TM_LPI; b2++) {
        const int d1 = si.eval_order[b2];
        if (!used[(size_t)d1] && flat.devs[(size_t)d1].type == flat.devs[(size_t)d0].type && flat.devs[(size_t)d1].n_itab == flat.devs[(size_t)d0].n_itab &&
            flat.devs[(size_t)d1].n_par == flat.devs[(size_t)d0].n_par) { slot.push_back(d1); used[(size_t)d1] = true; }
      }
      slots.push_back(slot);
    }
    auto selj = [&](const std::vector<long long>& v) {  // value of member j's device: a literal when all agree
      bool same = true;
      for (size_t k = 1; k < v.size(); k++) same = same && v[k] == v[0];
      if (same) return std::to_string(v[0]);
      std::string r;
      for (size_t k = 0; k + 1 < v.size(); k++) r += "(j == " + std::to_string(k) + " ? " + std::to_string(v[k]) + " : ";
      r += std::to_string(v.back());
      for (size_t k = 0; k + 1 < v.size(); k++) r += ")";
      return r;
    };
    for (size_t sidx = 0; sidx < slots.size(); sidx++) {
      const std::vector<int>& sl = slots[sidx];
      auto vals = [&](auto f) { std::vector<long long> v; for (int d : sl) v.push_back((long long)f(flat.devs[(size_t)d], d)); return v; };
      const FlatDev& d0 = flat.devs[(size_t)sl[0]];
      o << "template <class Base> struct P" << sidx << " : Base {\n  using Base::pval; using Base::pinst; using Base::sop; using Base::sguess; using Base::S; using Base::j;\n";
      o << "  __device__ __forceinline__ int node(int k) const { switch (k) {";
      for (int k = 0; k < d0.n_itab; k++)
        o << " case " << k << ": return " << selj(vals([&](const FlatDev& d, int) { return itab[(size_t)d.itab_off + (size_t)k]; })) << ";";
      o << " default: return -1; } }\n";
      o << "  __device__ __forceinline__ double par(int k) const { switch (k) {";
      for (int k = 0; k < d0.n_par; k++) {
        const std::string off = selj(vals([&](const FlatDev& d, int) { return pcode[(size_t)d.par_off + (size_t)k] >> 1; }));
        const std::vector<long long> fl = vals([&](const FlatDev& d, int) { return pcode[(size_t)d.par_off + (size_t)k] & 1; });
        bool any = false, all = true;
        for (long long f : fl) { any = any || f; all = all && f; }
        o << " case " << k << ": return __ldg(pval + " << off << (all ? " + pinst" : any ? " + (" + selj(fl) + " ? pinst : (size_t)0)" : "") << ");";
      }
      o << " default: return 0.0; } }\n";
      const std::string so = selj(vals([&](const FlatDev& d, int) { return d.state_off; }));
      const std::string st = selj(vals([&](const FlatDev&, int d) { return si.stage_off[(size_t)d]; }));
      if (SOPG) o << "  using Base::sops;\n  __device__ __forceinline__ double op(int k) const { return __ldcg(sop + (size_t)(" << so << " + k) * sops); }\n";
      else o << "  __device__ __forceinline__ double op(int k) const { return sop[(" << so << " + k) * PS]; }\n";
      o << "  __device__ __forceinline__ double guess(int k) const { return sguess[(" << so << " + k) * PS]; }\n";
      o << "  __device__ __forceinline__ void set_guess(int k, double v) { sguess[(" << so << " + k) * PS] = v; }\n";
      o << "  __device__ __forceinline__ void add_g_at(int pos, double v) { S[(" << st << " + pos) * PS] = v; }\n";
      o << "  __device__ __forceinline__ void add_b_at(int pos, double v) { S[(" << st << " + pos) * PS] = v; }\n";
      o << "  __device__ __forceinline__ void add_g_dup(int, int dup, double v) { S[(" << st << " + dup) * PS] = v; }\n};\n";
    }
  }
  const size_t n_gt = G.table.size();
  const size_t ctrl_ints = (size_t)GI + n_gt;
  const size_t ctrl_bytes = (ctrl_ints * 4 + 15) / 16 * 16;
  *smem_out = ctrl_bytes + 8 * (size_t)PSV * ((size_t)N + (size_t)NST + 1 + (SOPG ? 1 : 2) * (size_t)NSTATE);

  o << "extern \"C\" __global__ void __launch_bounds__(" << NW * 32 << ", " << 2 << ") k_jit(const double* __restrict__ pval, double* gx, double* st_op, double* st_guess,\n"
       "    int* status, int* iters, int* loads, size_t stride, size_t st_stride, int B, int n_state_arg, int mode, double gmin, double dt,\n"
       "    double reltol, double iabstol, int cold, int T_points, int n_save, const int* __restrict__ save_vars, double* wave, double* rows, int max_iter) {\n"
       "  extern __shared__ __align__(16) unsigned char smem_raw[];\n"
       "  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;\n"
       "  const int i0 = blockIdx.x * " << GI << ";\n"
       "  const int ni = min(" << GI << ", B - i0);\n"
       "  const int ei = lane, base = lane % " << IPW << ", j = lane / " << IPW << ", ri = warp * " << IPW << " + base;\n"
       "  const unsigned imask = " << IMASK << "u << base, jbit = 1u << j;\n"
       "  int* act_s = (int*)smem_raw;\n  int* gt = act_s + " << GI << ";\n"
       "  double* X = (double*)(smem_raw + " << ctrl_bytes << ");\n"
       "  double* S = X + " << N * PSV << ";\n"
       << (SOPG ? "  double* sop = st_op + i0;  // committed state stays in its HBM column: entry k of instance e at sop[k * st_stride + e]\n"
                : "  double* sop = S + " + std::to_string((NST + 1) * PSV) + ";\n")
       << "  double* sguess = " << (SOPG ? "S + " + std::to_string((NST + 1) * PSV) : "sop + " + std::to_string(NSTATE * PSV)) << ";\n"
       "  for (int k = tid; k < " << n_gt << "; k += " << NW * 32 << ") gt[k] = GT_G[k];\n"
       "  if (tid < PS) S[" << zero_off << " + tid] = 0.0;\n"
       "  const bool evalid = ei < ni, rvalid = ri < ni" << (POW2 ? "" : " && j < " + std::to_string(TM_LPI)) << ", rin = ri < " << GI << ";\n"
       "  " << (XP ? "if (ei < PS) " : "") << "for (int k = warp; k < " << N << "; k += " << NW << ") X[k * PS + ei] = (evalid && !cold) ? gx[(size_t)k * stride + i0 + ei] : 0.0;\n"
       "  " << (XP ? "if (ei < PS) " : "") << "for (int k = warp; k < " << flat.n_state << "; k += " << NW << ") {\n"
       "    const size_t src = (size_t)k * st_stride + (size_t)i0 + (size_t)(evalid ? ei : 0);\n"
       << (SOPG ? "    if (cold && evalid) st_op[src] = 0.0;\n" : "    sop[k * PS + ei] = cold ? 0.0 : st_op[src];\n")
       << "    sguess[k * PS + ei] = cold ? 0.0 : st_guess[src];\n  }\n"
       "  int r_stat = " << (tran ? "rvalid ? status[i0 + ri] : 0" : "0") << ";\n"
       "  int r_wk = 0;\n"
       "  int r_nsol = 0, r_nld = 0;\n"
       "  __syncthreads();\n";
  if (tran)
    o << "  if (rvalid) for (int s = j; s < n_save; s += " << TM_LPI << ") wave[(size_t)s * stride + i0 + ri] = X[save_vars[s] * PS + ri];\n";
  o << "  const int* gtj = gt + j;\n  const double* Sri = S + ri;\n";
  for (int q = 0; q < Q; q++) {
    o << "  const int xo" << q << " = XO_G[" << q * TM_LPI << " + j];\n";
    o << "  const bool v" << q << " = " << (valid_mask(q) == FULLSET && POW2 ? std::string("true") : "j < " + std::to_string(std::min(TM_LPI, N - q * TM_LPI))) << ";\n";
  }
  const char* ebi = WP ? "ri" : "ei";  // the instance a lane evaluates devices for
  o << "  JBase eb; eb.pval = pval; eb.pinst = (size_t)i0 + (size_t)" << ebi << "; eb.sop = sop + " << ebi << "; eb.sguess = sguess + " << ebi
    << "; eb.X = X + " << ebi << "; eb.S = S + " << ebi << ";" << (XP ? " eb.j = j;" : "") << (SOPG ? " eb.sops = st_stride;" : "") << "\n"
       "  eb.mode = " << (tran ? "AN_TRAN" : "AN_OP") << "; eb.dt = dt; eb.gmin = gmin; eb.omega = 0.0; eb.time = " << (tran ? "dt" : "0.0") << ";\n";  // literal: the other mode's code is dropped
  if (prof) o << "  __shared__ long long prof_s[32];\n  if (tid < 32) prof_s[tid] = 0;\n  __syncthreads();\n  long long t_last = clock64();\n";
  o << "  const int n_points = " << (tran ? "T_points" : "2") << ";\n"
       "  for (int tp = 1; tp < n_points; tp++" << (tran ? ", eb.time += dt" : "") << ") {\n"
       "    bool r_act = rvalid && r_stat == 0;\n    bool r_dxok = true;\n"
       << (WP ? "    __syncwarp();\n" : "    if (j == 0 && rin) act_s[ri] = r_act ? 1 : 0;\n    __syncthreads();\n");
  for (int q = 0; q < Q; q++) o << "    double xp" << q << " = (v" << q << " && rin) ? X[xo" << q << " + ri] : 0.0;\n";
  auto load_fn = [&](int dev) -> const char* {
    switch (flat.devs[(size_t)dev].type) {
      case DT_R: return "load_resistor";
      case DT_C: return "load_capacitor";
      case DT_I: return "load_isrc";
      case DT_V: return "load_vsrc";
      case DT_DIODE: return "load_diode";
      case DT_MOS0: return "load_mos0";
      case DT_MOS1: return "load_mos1";
      default: return nullptr;
    }
  };
  o << "    for (int iter = 0; iter < max_iter; iter++) {\n      PH(0)\n";
  if (WP) {
    // ---- device evaluation, warp-private: member j of instance `ri` evaluates device j of each slot
    for (size_t sidx = 0; sidx < slots.size(); sidx++) {
      const std::vector<int>& sl = slots[sidx];
      const char* fn = load_fn(sl[0]);
      if (!fn) continue;
      const std::string cond = (int)sl.size() == TM_LPI ? "r_act" : "r_act && j < " + std::to_string(sl.size());
      // Mos1 reads no in-flight state, so a flagged evaluation can simply be repeated on the exact path
      if (fast_la && flat.devs[(size_t)sl[0]].type == DT_MOS1)
        o << "      if (" << cond << ") { P" << sidx << "<JFast> e; static_cast<JBase&>(e) = eb; e.dbad = false; " << fn << "(e);\n"
             "        if (e.dbad) { P" << sidx << "<JBase> x; static_cast<JBase&>(x) = eb; " << fn << "(x); } }\n";
      else
        o << "      if (" << cond << ") { P" << sidx << "<JBase> e; static_cast<JBase&>(e) = eb; " << fn << "(e); }\n";
    }
    o << "      PH(1)\n      __syncwarp();\n      PH(2)\n";
  } else {
    o << "      if (ei < " << GI << " && act_s[ei]) {\n        switch (warp) {\n";
    // ---- device evaluation: eval_order position w, w+NW, ... on warp w
    for (int w = 0; w < NW && w < (int)si.eval_order.size(); w++) {
      o << "          case " << w << ": {\n";
      for (size_t item = (size_t)w; item < si.eval_order.size(); item += (size_t)NW) {
        const int dev = si.eval_order[item];
        const char* fn = load_fn(dev);
        if (fn && fast_la && flat.devs[(size_t)dev].type == DT_MOS1)
          o << "            { E" << dev << "<JFast> e; static_cast<JBase&>(e) = eb; e.dbad = false; " << fn << "(e);\n"
               "              if (e.dbad) { E" << dev << "<JBase> x; static_cast<JBase&>(x) = eb; " << fn << "(x); } }\n";
        else if (fn)
          o << "            { E" << dev << "<JBase> e; static_cast<JBase&>(e) = eb; " << fn << "(e); }\n";
      }
      o << "          } break;\n";
    }
    o << "        }\n      }\n      PH(1)\n      __syncthreads();\n      PH(2)\n";
  }
  // ---- linear algebra, warp-synchronous
  // residual in pivoted row order; x by pivoted column comes from the member that owns it
  auto emit_residual = [&](std::ostream& o) {
    for (int q = 0; q < Q; q++) o << "        c" << q << " = 0.0;\n";
    for (int c = 0; c < N; c++) {
      bool any = false;
      for (int q = 0; q < Q; q++) any = any || M[(size_t)q][(size_t)c];
      if (!any) continue;
      o << "        { const double xc = BC(xp" << qof(c) << ", " << jof(c) << ");\n";
      for (int q = 0; q < Q; q++) {
        const unsigned m = M[(size_t)q][(size_t)c];
        if (!m) continue;
        o << "          " << (m == FULLSET ? std::string("") : "if " + mask_test(m) + " ") << "c" << q << " = s_add(c" << q << ", s_mul(" << A(q, c)
          << ", xc));\n";
      }
      o << "        }\n";
    }
    for (int q = 0; q < Q; q++) o << "        c" << q << " = s_sub(b" << q << ", c" << q << ");\n";
  };
  // numeric LU on the frozen pattern + both substitutions. `fast` = branch-free text: every division is the fast path
  // of scalar.h with its exception deferred into `dbad`, every lane mask and every zero test a select, so that the whole
  // solve is one basic block (independent pivots, rows and columns overlap); the exact text is kept for the redo.
  auto emit_solve = [&](std::ostream& o, bool fast) {
    auto divide = [&](const std::string& dst, const std::string& num, const std::string& den, const std::string& rcp, const std::string& bok,
                      const std::string& cond) {
      if (fast)
        o << "            { bool bd = false; const double t = s_div_rp(" << num << ", " << den << ", " << rcp << ", " << bok << ", bd); if (" << cond << ") { "
          << dst << " = t; dbad = dbad || bd; } }\n";
      else
        o << "            if (" << cond << ") " << dst << " = s_div_r(" << num << ", " << den << ", " << rcp << ");\n";
    };
    for (int k = 0; k + 1 < N; k++) {
      const int qk = qof(k), jk = jof(k);
      o << "          { const double piv = BC(" << A(qk, k) << ", " << jk << "); sing = sing || (piv == 0.0);\n";
      bool anyL = false;
      for (int q = 0; q < Q; q++) anyL = anyL || LM[(size_t)q][(size_t)k];
      if (anyL) {
        o << "            const double rp = s_rcp(piv);\n";  // one reciprocal per pivot, shared by the column's entries (scalar.h)
        if (fast) o << "            const bool pok = s_div_bok(piv);\n";
        // pivot health (kernels/newton.cu): |pivot| < 1e-3 x an entry below it <=> a multiplier above 1e3 — tested on the
        // quotient that is computed anyway, as a running integer maximum of the high words (FP64 compares with a live
        // predicate cost 6-13 % of the C2 kernel, profiles/r02f_health_cost.txt). S21_PIVOT_HEALTH=0 leaves the test out
        // of the generated text (to measure what it costs; the interpreted kernels always test).
        static const bool health = [] { const char* e = std::getenv("S21_PIVOT_HEALTH"); return !e || std::atoi(e) != 0; }();
        for (int q = 0; q < Q; q++)
          if (LM[(size_t)q][(size_t)k]) {
            divide(A(q, k), A(q, k), "piv", "rp", "pok", mask_test(LM[(size_t)q][(size_t)k]));
            if (health && !tran && P.piv_checked[(size_t)k])  // (dcop text only: nobody re-pivots inside a device-resident time loop)  // (threshold-checked pivots only) integer max of the multipliers' high words (monotonic in |l|): ALU pipe, no predicate kept alive
              o << "            if (" << mask_test(LM[(size_t)q][(size_t)k]) << ") r_wk = max(r_wk, __double2hiint(" << A(q, k) << ") & 0x7fffffff);\n";
          }
        for (int s = P.diag_slot[(size_t)k] + 1; s < P.rowptr[(size_t)k + 1]; s++) {
          const int c = P.colidx[(size_t)s];
          o << "            { const double u = BC(" << A(qk, c) << ", " << jk << ");\n";
          for (int q = 0; q < Q; q++)
            if (LM[(size_t)q][(size_t)k]) {
              const std::string upd = "s_sub(" + A(q, c) + ", s_mul(u, " + A(q, k) + "))";
              o << "              if " << mask_test(LM[(size_t)q][(size_t)k]) << " " << A(q, c) << " = " << upd << ";\n";
            }
          o << "            }\n";
        }
      }
      o << "          }\n";
    }
    o << "          PH(5)\n";
    // forward substitution
    for (int k = 0; k < N; k++) {
      bool anyL = false;
      for (int q = 0; q < Q; q++) anyL = anyL || LM[(size_t)q][(size_t)k];
      if (!anyL) continue;
      o << "          { const double ck = BC(c" << qof(k) << ", " << jof(k) << ");\n";
      if (fast) {
        o << "            const bool nz = !(ck == 0.0);\n";
        for (int q = 0; q < Q; q++)
          if (LM[(size_t)q][(size_t)k])
            o << "            if (nz && " << mask_test(LM[(size_t)q][(size_t)k]) << ") c" << q << " = s_sub(c" << q << ", s_mul(ck, " << A(q, k) << "));\n";
        o << "          }\n";
      } else {
        o << "            if (!(ck == 0.0)) {\n";
        for (int q = 0; q < Q; q++)
          if (LM[(size_t)q][(size_t)k])
            o << "              if " << mask_test(LM[(size_t)q][(size_t)k]) << " c" << q << " = s_sub(c" << q << ", s_mul(ck, " << A(q, k) << "));\n";
        o << "            }\n          }\n";
      }
    }
    o << "          PH(6)\n";
    // backward substitution: the owner of row k forms its sum over ascending columns, then broadcasts the result
    std::vector<bool> need_bc((size_t)N, false);
    for (int r = 0; r < N; r++)
      for (int s = P.diag_slot[(size_t)r] + 1; s < P.rowptr[(size_t)r + 1]; s++) need_bc[(size_t)P.colidx[(size_t)s]] = true;
    // reciprocals of the diagonal, all rows of a set at once (off the substitution's dependent chain)
    for (int q = 0; q < Q; q++) {
      o << "          double dg" << q << " = 1.0;\n";
      for (int jj = 0; jj < TM_LPI && q * TM_LPI + jj < N; jj++)
        o << "          if (j == " << jj << ") dg" << q << " = " << A(q, q * TM_LPI + jj) << ";\n";
      o << "          const double rd" << q << " = s_rcp(dg" << q << ");\n";
      if (fast) o << "          const bool dok" << q << " = s_div_bok(dg" << q << ");\n";
    }
    for (int k = N - 1; k >= 0; k--) {
      const int qk = qof(k), jk = jof(k);
      o << "          { double ck = c" << qk << ";\n";
      for (int s = P.diag_slot[(size_t)k] + 1; s < P.rowptr[(size_t)k + 1]; s++) {
        const int c = P.colidx[(size_t)s];
        o << "            ck = s_sub(ck, s_mul(cb" << c << ", " << A(qk, c) << "));\n";
      }
      divide("c" + std::to_string(qk), "ck", "dg" + std::to_string(qk), "rd" + std::to_string(qk), "dok" + std::to_string(qk), "j == " + std::to_string(jk));
      o << "          }\n";
      if (need_bc[(size_t)k]) o << "          const double cb" << k << " = BC(c" << qk << ", " << jk << ");\n";
    }
  };
  o << "      if (__any_sync(FULLM, r_act)) {\n";
  o << gath.str() << "        PH(3)\n";
  o << "        double c0";
  for (int q = 1; q < Q; q++) o << ", c" << q;
  o << ";\n";
  emit_residual(o);
  o << "        bool bad = false;\n";
  for (int q = 0; q < Q; q++) o << "        bad = bad || (v" << q << " && s_abs(c" << q << ") > iabstol);\n";
  o << "        const bool resok = (__ballot_sync(FULLM, bad) & imask) == 0;\n"
       "        if (r_act) {\n          r_nld += 1;\n          if (r_dxok && resok) {\n"
       "            for (int k = j; k < " << flat.n_state << "; k += " << TM_LPI << ") sop[" << (SOPG ? "(size_t)k * st_stride" : "k * PS") << " + ri] = sguess[k * PS + ri];\n"
       "            r_act = false;\n          }\n        }\n";
  o << "        PH(4)\n        if (__any_sync(FULLM, r_act)) {\n          bool sing = false;\n";
  if (fast_la) {
    o << "          bool dbad = false;\n          {\n";
    emit_solve(o, true);
    o << "          }\n";
    // deferred exceptions: some quotient of this warp left the fast path's domain -> the exact text, from the stamps
    // (only of instances still iterating: finished and padding instances compute on stale or arbitrary data)
    o << "          if (__any_sync(FULLM, dbad && r_act)) {\n          sing = false;\n          r_wk = 0;\n";  // the fast text's quotients were not valid
    std::string g2 = gath.str();
    o << g2;
    emit_residual(o);
    emit_solve(o, false);
    o << "          }\n";
  } else {
    emit_solve(o, false);
  }
  // max |dx| over the team, global step limit, update
  o << "          PH(7)\n          double m = 0.0;\n";
  for (int q = 0; q < Q; q++) o << "          if (v" << q << ") m = fmax(m, s_abs(c" << q << "));\n";
  if (POW2) {
    for (int off = IPW; off < 32; off *= 2) o << "          m = fmax(m, __shfl_xor_sync(FULLM, m, " << off << "));\n";
  } else {  // reduce down to member 0, then broadcast
    for (int off = 8; off >= 1; off /= 2)
      o << "          m = fmax(m, __shfl_sync(FULLM, m, j + " << off << " < " << TM_LPI << " ? lane + " << IPW * off << " : lane));\n";
    o << "          m = BC(m, 0);\n";
  }
  // pivot health: a multiplier of this factorisation >= 1000 on any lane of the instance -> stop before the update, the host
  // re-pivots at this iterate (kernels/newton.cu, host/batch.hpp resolve_repivot)
  {  // a multiplier above SolveCtl::weak_mult (high words compare like the magnitudes; the bound's own high word is the literal)
    const double wm = pivot_weak_mult();
    long long bits; std::memcpy(&bits, &wm, sizeof bits);
    char lit[24]; std::snprintf(lit, sizeof lit, "0x%08x", (unsigned)((unsigned long long)bits >> 32));
    o << "          const bool wk = (__ballot_sync(FULLM, r_wk > " << lit << ") & imask) != 0;\n          r_wk = 0;\n";
  }
  o << "          bool baddx = false;\n          const double rm = s_rcp(m);\n          if (r_act && !sing && !wk) {\n";
  for (int q = 0; q < Q; q++)
    o << "            if (v" << q << ") { double dxk = c" << q << "; if (m > 1.0) dxk = s_div_r(s_mul(dxk, 1.0), m, rm); xp" << q << " = s_add(xp" << q
      << ", dxk); X[xo" << q << " + ri] = xp" << q << "; baddx = baddx || (s_abs(dxk) > reltol); }\n";
  o << "          }\n"
       "          r_dxok = (__ballot_sync(FULLM, baddx) & imask) == 0;\n"
       "          if (r_act) {\n            if (sing) { r_act = false; r_stat = 2; }\n            else if (wk) { r_act = false; r_stat = 9; }\n            else r_nsol += 1;\n          }\n"
       "        }\n      }\n"
       << (WP ? "      PH(8)\n      __syncwarp();\n      const bool any_ = __any_sync(FULLM, r_act);\n      PH(9)\n      if (!any_) break;\n"
              : "      if (j == 0 && rin) act_s[ri] = r_act ? 1 : 0;\n      PH(8)\n"
                "      const int any_ = __syncthreads_or(r_act);\n      PH(9)\n      if (!any_) break;\n") <<
       "    }\n"
       "    if (r_act) { r_stat = 1; r_act = false; }\n";
  if (tran)
    o << "    if (rvalid) {\n      const bool good = r_stat == 0;\n      for (int s = j; s < n_save; s += " << TM_LPI << ")\n"
         "        wave[((size_t)tp * n_save + s) * stride + i0 + ri] = good ? X[save_vars[s] * PS + ri] : __longlong_as_double(0x7ff8000000000000LL);\n"
         "    }\n";
  o << "  }\n  __syncthreads();\n"
       "  if (evalid) {\n"
       "    for (int k = warp; k < " << N << "; k += " << NW << ") gx[(size_t)k * stride + i0 + ei] = X[k * PS + ei];\n"
       "    for (int k = warp; k < " << flat.n_state << "; k += " << NW << ") {\n"
       "      const size_t dst = (size_t)k * st_stride + i0 + ei;\n"
       << (SOPG ? "" : "      st_op[dst] = sop[k * PS + ei];\n") << "      st_guess[dst] = sguess[k * PS + ei];\n    }\n  }\n"
       "  if (rvalid && j == 0) {\n"
       "    status[i0 + ri] = r_stat;\n"
       "    iters[i0 + ri] = (cold ? 0 : iters[i0 + ri]) + r_nsol;\n"
       "    loads[i0 + ri] = (cold ? 0 : loads[i0 + ri]) + r_nld;\n  }\n";
  // The host's result layout ([instance][variable] rows, then status / iters / loads as int32 — k_pack_out in kernels/newton.cu),
  // written by the kernel itself when `rows` is given, so that a read needs no packing kernel. A CTA's rows are one contiguous
  // block: they leave shared memory as flat, fully coalesced 8-byte stores (the destination may be mapped pinned HOST memory —
  // the result then crosses PCIe while other CTAs still iterate, and a read is only a stream synchronise).
  o << "  if (rows) {\n    __syncthreads();\n"
       "    if (rvalid && j == 0) { int* tail = (int*)(rows + (size_t)B * " << N << "); tail[i0 + ri] = r_stat; tail[(size_t)B + i0 + ri] = iters[i0 + ri];"
       " tail[2 * (size_t)B + i0 + ri] = loads[i0 + ri]; }\n"
       "    double* rb = rows + (size_t)i0 * " << N << ";\n"
       "    for (int f = tid; f < ni * " << N << "; f += " << NW * 32 << ") rb[f] = X[(f % " << N << ") * PS + f / " << N << "];\n  }\n";
  if (prof)
    o << "  __syncthreads();\n  if (tid == 0 && blockIdx.x == 0) {\n"
         "    const char* nm[10] = {\"loop/other\", \"eval\", \"barrier-after-eval\", \"gather\", \"residual+conv\", \"LU\", \"forward\", \"backward\", \"limit+update\", \"end-barrier\"};\n"
         "    for (int k = 0; k < 10; k++) printf(\"[team profile] %-20s warp0 %10lld   last warp %10lld cycles\\n\", nm[k], prof_s[k], prof_s[16 + k]);\n  }\n";
  o << "}\n";
  o << "}  // namespace s21\n";
  return o.str();
}

}  // namespace jit
}  // namespace s21
