// The batched Newton loop on the GPU: one thread per circuit instance, the whole loop (and for transient the whole
// time loop) inside ONE kernel launch — device evaluation, MNA assembly, residual, convergence test, sparse LU
// refactorisation, forward/back substitution, step limiting and the solution update never leave the device.
//
// Replaces, per instance:  Solver::<f64>::solve      spice21/src/analysis.rs:169-210
//                          Solver::<Complex>::solve  spice21/src/analysis.rs:253-303
//                          Solver::converged         spice21/src/analysis.rs:331-345
//                          Matrix::{reset,update,res,lu_factorize(numeric part),solve}  spice21/src/sparse21/mod.rs:272-327, 865-991
//                          Tran::solve time loop     spice21/src/analysis.rs:553-570
//
// Data layout: every per-instance quantity is a column of a structure-of-arrays table, a[k*stride + instance], so a
// warp (32 consecutive instances) touches 256 contiguous bytes per access; all index tables (device tables, stamp
// handles, the LU plan) are shared by the batch and are read with warp-uniform addresses (one L1/L2 sector per warp).
// Within an instance every floating-point operation happens in the reference's order (assembly in component/push
// order, Schur updates in ascending pivot order, substitution sums in list order) and the file is compiled with
// -fmad=false, so results differ from the CPU restatement only through libm's exp/log.
#include <cuda_runtime.h>

#include <type_traits>

#include "devices.cuh"
#include "engine.hpp"

namespace s21 {

enum { ST_OK_ = 0, ST_CONV_ = 1, ST_SINGULAR_ = 2, ST_REPIVOT_ = 9 };  // 9: engine.hpp ST_REPIVOT_CODE

// ------------------------------------------------------------------------------------------------ per-thread Env
template <class T>
struct Env {
  // program (warp-uniform)
  const int* it;
  const int* pc;
  const double* pval;
  // state columns of this device
  double* sop;
  double* sguess;
  size_t sstride, sinst, pinst;
  // workspace columns of this instance
  const T* x;
  T* lu;
  T* rhs;
  size_t stride, w;
  int mode;
  double dt, gmin, omega;
  double time;  // transient: the time point being solved (time-varying sources)

  __device__ __forceinline__ int node(int k) const { return __ldg(it + k); }
  __device__ __forceinline__ double par(int k) const {
    const int c = __ldg(pc + k);
    return pval[(size_t)(c >> 1) + (size_t)(c & 1) * pinst];
  }
  __device__ __forceinline__ double volt(int var) const;
  __device__ __forceinline__ double op(int k) const { return sop[(size_t)k * sstride + sinst]; }
  __device__ __forceinline__ double guess(int k) const { return sguess[(size_t)k * sstride + sinst]; }
  __device__ __forceinline__ void set_guess(int k, double v) { sguess[(size_t)k * sstride + sinst] = v; }
  __device__ __forceinline__ void add_g(int h, T v) {
    if (h >= 0) { T* a = lu + (size_t)h * stride + w; *a = s_add(*a, v); }
  }
  __device__ __forceinline__ void add_b(int var, T v) {
    if (var >= 0) { T* a = rhs + (size_t)var * stride + w; *a = s_add(*a, v); }
  }
  // direct sinks: one thread owns the instance, so stamps go straight into A / rhs in push order
  __device__ __forceinline__ void add_g_at(int pos, T v) { add_g(node(pos), v); }
  __device__ __forceinline__ void add_b_at(int pos, T v) { add_b(node(pos), v); }
  __device__ __forceinline__ void add_g_dup(int pos, int, T v) { add_g(node(pos), v); }
};
template <> __device__ __forceinline__ double Env<double>::volt(int var) const { return var < 0 ? 0.0 : x[(size_t)var * stride + w]; }
template <> __device__ __forceinline__ double Env<cplx>::volt(int) const { return 0.0; }  // load_ac never reads the guess

// One pass of Solver::update (analysis.rs:153-168 / 237-252): every device, in component order.
// Returns false when a device has no load function for this analysis (the reference panics: comps/mod.rs:86-88).
template <bool B4>
__device__ __forceinline__ bool load_sweep(const DevTables& d, Env<double>& e, double* st_op, double* st_guess) {
  for (int k = 0; k < d.n_dev; k++) {
    e.it = d.itab + __ldg(d.itab_off + k);
    e.pc = d.pcode + __ldg(d.par_off + k);
    const size_t so = (size_t)__ldg(d.state_off + k) * e.sstride;
    e.sop = st_op + so;
    e.sguess = st_guess + so;
    switch (__ldg(d.type + k)) {
      case DT_R: load_resistor(e); break;
      case DT_C: load_capacitor(e); break;
      case DT_I: load_isrc(e); break;
      case DT_V: load_vsrc(e); break;
      case DT_DIODE: load_diode(e); break;
      case DT_MOS0: load_mos0(e); break;
      case DT_MOS1: load_mos1(e); break;
      case DT_BSIM4: if constexpr (B4) load_bsim4(e); else return false; break;
      default: return false;
    }
  }
  return true;
}
template <bool B4>
__device__ __forceinline__ bool load_sweep(const DevTables& d, Env<cplx>& e, double* st_op, double* st_guess) {
  for (int k = 0; k < d.n_dev; k++) {
    e.it = d.itab + __ldg(d.itab_off + k);
    e.pc = d.pcode + __ldg(d.par_off + k);
    const size_t so = (size_t)__ldg(d.state_off + k) * e.sstride;
    e.sop = st_op + so;
    e.sguess = st_guess + so;
    switch (__ldg(d.type + k)) {
      case DT_R: load_ac_resistor(e); break;
      case DT_C: load_ac_capacitor(e); break;
      case DT_V: load_ac_vsrc(e); break;
      case DT_MOS1: load_ac_mos1(e); break;
      default: return false;  // Isrc, Diode, Mos0, Bsim4 have no load_ac in the reference
    }
  }
  return true;
}

template <class T> struct Tol;
template <> struct Tol<double> {  // analysis.rs:331-345: fail on  > tol
  static __device__ __forceinline__ bool ok(double a, double tol) { return !(a > tol); }
  static const int max_iter = 100;  // analysis.rs:173
};
template <> struct Tol<cplx> {  // analysis.rs:271-272: pass on  < tol
  static __device__ __forceinline__ bool ok(double a, double tol) { return a < tol; }
  static const int max_iter = 20;  // analysis.rs:258
};

// Commit every device: op <- guess (analysis.rs:190-192; Capacitor/Diode/Mos1::commit)
__device__ __forceinline__ void commit_state(const DevTables& d, double* st_op, const double* st_guess, size_t sstride, size_t sinst) {
  for (int k = 0; k < d.n_state; k++) st_op[(size_t)k * sstride + sinst] = st_guess[(size_t)k * sstride + sinst];
}

// The Newton shell for one instance. Returns an S21 status; *n_solves / *n_loads are incremented.
template <class T, bool B4>
__device__ int newton_solve(const DevTables& d, const PlanTables& p, const WorkTables<T>& wk, const SolveCtl& ctl, size_t inst,
                            double omega, double vtol, double itol, bool do_commit, int* n_solves, int* n_loads) {
  const size_t S = wk.stride;
  T* x = wk.x + inst;
  T* rhs = wk.rhs + inst;
  T* c = wk.c + inst;
  T* lu = wk.lu + inst;
  const size_t pinst = inst * ctl.par_inst_stride;
  Env<T> e;
  e.pval = d.pval;
  e.sstride = wk.st_stride; e.sinst = pinst; e.pinst = pinst;
  e.x = wk.x; e.lu = wk.lu; e.rhs = wk.rhs; e.stride = S; e.w = inst;
  e.mode = ctl.mode; e.dt = ctl.dt; e.gmin = ctl.gmin; e.omega = omega; e.time = ctl.time;
  const int N = p.N;
  bool dx_ok = true;  // dx = 0 before the first iteration
  const int max_it = std::is_same<T, double>::value ? min(Tol<T>::max_iter, ctl.max_iter) : Tol<T>::max_iter;
  for (int iter = 0; iter < max_it; iter++) {
    // Matrix::reset + fresh rhs (analysis.rs:178-179)
    for (int k = 0; k < p.nnz; k++) lu[(size_t)k * S] = Scalar<T>::zero();
    for (int k = 0; k < N; k++) rhs[(size_t)k * S] = Scalar<T>::zero();
    if (!load_sweep<B4>(d, e, wk.st_op, wk.st_guess)) return 6;  // S21_UNSUPPORTED
    *n_loads += 1;
    // Matrix::res (sparse21/mod.rs:298-327), produced directly in internal row order: c[k] = res[row_i2e[k]]
    bool res_ok = true;
    for (int r = 0; r < N; r++) {
      T acc = Scalar<T>::zero();
      const int b = __ldg(p.rowptr + r), en = __ldg(p.rowptr + r + 1);
      int s = b;
      for (; s + 4 <= en; s += 4) {  // operands of four entries in flight, the sum in list order (see the LU loop below)
        const int c0 = __ldg(p.col_i2e + __ldg(p.colidx + s)), c1 = __ldg(p.col_i2e + __ldg(p.colidx + s + 1));
        const int c2 = __ldg(p.col_i2e + __ldg(p.colidx + s + 2)), c3 = __ldg(p.col_i2e + __ldg(p.colidx + s + 3));
        const T a0 = lu[(size_t)s * S], a1 = lu[(size_t)(s + 1) * S], a2 = lu[(size_t)(s + 2) * S], a3 = lu[(size_t)(s + 3) * S];
        const T x0 = x[(size_t)c0 * S], x1 = x[(size_t)c1 * S], x2 = x[(size_t)c2 * S], x3 = x[(size_t)c3 * S];
        acc = s_add(s_add(s_add(s_add(acc, s_mul(a0, x0)), s_mul(a1, x1)), s_mul(a2, x2)), s_mul(a3, x3));
      }
      for (; s < en; s++) {
        const int col = __ldg(p.colidx + s);
        acc = s_add(acc, s_mul(lu[(size_t)s * S], x[(size_t)__ldg(p.col_i2e + col) * S]));
      }
      const T rv = s_sub(rhs[(size_t)__ldg(p.row_i2e + r) * S], acc);
      c[(size_t)r * S] = rv;
      res_ok = res_ok && Tol<T>::ok(s_abs(rv), itol);
    }
    if (dx_ok && res_ok) {
      if (do_commit) commit_state(d, wk.st_op, wk.st_guess, wk.st_stride, pinst);
      return ST_OK_;
    }
    // ---- numeric LU on the frozen pattern (row_col_elim, sparse21/mod.rs:865-919), ascending pivot index
    bool weak = false;
    for (int k = 0; k + 1 < N; k++) {
      const T piv = lu[(size_t)__ldg(p.diag_slot + k) * S];
      if (s_is_zero(piv)) return ST_SINGULAR_;
      const int lb = __ldg(p.l_off + k), le = __ldg(p.l_off + k + 1);
      const double pth = __ldg(p.piv_chk + k) ? s_abs(piv) * ctl.weak_mult : INFINITY;  // pivots the fallback searches chose carry no threshold
      // One thread walks its instance's whole factorisation against an HBM / L2-resident workspace: every operation is a
      // chain index -> operand -> store, and the compiler cannot overlap two of them (a store to lu[] may alias the next
      // load). Within one pivot step the operations are independent — the entries of the L column, and the update targets
      // (i, j), are all distinct, the operands (pivot row, divided column) are not written — so four are issued together:
      // indices, then operands, then the arithmetic and the stores. Same values, same bits; four times the loads in flight
      // (config C5, 100 000 AC points on this kernel: the one workload whose bytes really move).
      int j = lb;
      for (; j + 4 <= le; j += 4) {
        T* a0 = lu + (size_t)__ldg(p.l_slot + j) * S; T* a1 = lu + (size_t)__ldg(p.l_slot + j + 1) * S;
        T* a2 = lu + (size_t)__ldg(p.l_slot + j + 2) * S; T* a3 = lu + (size_t)__ldg(p.l_slot + j + 3) * S;
        const T v0 = *a0, v1 = *a1, v2 = *a2, v3 = *a3;
        if (pth < s_abs(v0) || pth < s_abs(v1) || pth < s_abs(v2) || pth < s_abs(v3)) weak = true;
        *a0 = s_div(v0, piv); *a1 = s_div(v1, piv); *a2 = s_div(v2, piv); *a3 = s_div(v3, piv);
      }
      for (; j < le; j++) {
        T* a = lu + (size_t)__ldg(p.l_slot + j) * S;
        // pivot health: the reference, which searches its pivots anew in every factorisation, would not have taken this
        // diagonal (|d| < 1e-3 * column max, mod.rs:735-783)
        if (pth < s_abs(*a)) weak = true;
        *a = s_div(*a, piv);
      }
      const int ub = __ldg(p.upd_off + k), ue = __ldg(p.upd_off + k + 1);
      j = ub;
      for (; j + 4 <= ue; j += 4) {
        T* t0 = lu + (size_t)__ldg(p.upd_t + j) * S; T* t1 = lu + (size_t)__ldg(p.upd_t + j + 1) * S;
        T* t2 = lu + (size_t)__ldg(p.upd_t + j + 2) * S; T* t3 = lu + (size_t)__ldg(p.upd_t + j + 3) * S;
        const T u0 = lu[(size_t)__ldg(p.upd_u + j) * S], u1 = lu[(size_t)__ldg(p.upd_u + j + 1) * S];
        const T u2 = lu[(size_t)__ldg(p.upd_u + j + 2) * S], u3 = lu[(size_t)__ldg(p.upd_u + j + 3) * S];
        const T l0 = lu[(size_t)__ldg(p.upd_l + j) * S], l1 = lu[(size_t)__ldg(p.upd_l + j + 1) * S];
        const T l2 = lu[(size_t)__ldg(p.upd_l + j + 2) * S], l3 = lu[(size_t)__ldg(p.upd_l + j + 3) * S];
        const T w0 = *t0, w1 = *t1, w2 = *t2, w3 = *t3;
        *t0 = s_sub(w0, s_mul(u0, l0)); *t1 = s_sub(w1, s_mul(u1, l1)); *t2 = s_sub(w2, s_mul(u2, l2)); *t3 = s_sub(w3, s_mul(u3, l3));
      }
      for (; j < ue; j++) {
        T* t = lu + (size_t)__ldg(p.upd_t + j) * S;
        const T v = s_mul(lu[(size_t)__ldg(p.upd_u + j) * S], lu[(size_t)__ldg(p.upd_l + j) * S]);
        *t = s_sub(*t, v);
      }
    }
    // The frozen order no longer suits this instance's matrix: stop BEFORE the update (x stays at the current iterate) and
    // let the host take a new order from this iterate and continue (host/batch.hpp resolve_repivot).
    if (weak && ctl.stop_on_weak) return ST_REPIVOT_;
    // ---- forward substitution (sparse21/mod.rs:947-964)
    for (int k = 0; k < N; k++) {
      const T ck = c[(size_t)k * S];
      if (s_is_zero(ck)) continue;
      const int lb = __ldg(p.l_off + k), le = __ldg(p.l_off + k + 1);
      int j = lb;
      for (; j + 4 <= le; j += 4) {  // the rows of one L column are distinct
        T* t0 = c + (size_t)__ldg(p.l_row + j) * S; T* t1 = c + (size_t)__ldg(p.l_row + j + 1) * S;
        T* t2 = c + (size_t)__ldg(p.l_row + j + 2) * S; T* t3 = c + (size_t)__ldg(p.l_row + j + 3) * S;
        const T l0 = lu[(size_t)__ldg(p.l_slot + j) * S], l1 = lu[(size_t)__ldg(p.l_slot + j + 1) * S];
        const T l2 = lu[(size_t)__ldg(p.l_slot + j + 2) * S], l3 = lu[(size_t)__ldg(p.l_slot + j + 3) * S];
        const T w0 = *t0, w1 = *t1, w2 = *t2, w3 = *t3;
        *t0 = s_sub(w0, s_mul(ck, l0)); *t1 = s_sub(w1, s_mul(ck, l1)); *t2 = s_sub(w2, s_mul(ck, l2)); *t3 = s_sub(w3, s_mul(ck, l3));
      }
      for (; j < le; j++) {
        T* t = c + (size_t)__ldg(p.l_row + j) * S;
        *t = s_sub(*t, s_mul(ck, lu[(size_t)__ldg(p.l_slot + j) * S]));
      }
    }
    // ---- backward substitution (sparse21/mod.rs:967-979)
    for (int k = N - 1; k >= 0; k--) {
      const int ds = __ldg(p.diag_slot + k);
      const int en = __ldg(p.rowptr + k + 1);
      T ck = c[(size_t)k * S];
      int s = ds + 1;
      for (; s + 4 <= en; s += 4) {  // operands of four entries in flight, the subtractions in list order
        const T c0 = c[(size_t)__ldg(p.colidx + s) * S], c1 = c[(size_t)__ldg(p.colidx + s + 1) * S];
        const T c2 = c[(size_t)__ldg(p.colidx + s + 2) * S], c3 = c[(size_t)__ldg(p.colidx + s + 3) * S];
        const T a0 = lu[(size_t)s * S], a1 = lu[(size_t)(s + 1) * S], a2 = lu[(size_t)(s + 2) * S], a3 = lu[(size_t)(s + 3) * S];
        ck = s_sub(s_sub(s_sub(s_sub(ck, s_mul(c0, a0)), s_mul(c1, a1)), s_mul(c2, a2)), s_mul(c3, a3));
      }
      for (; s < en; s++) ck = s_sub(ck, s_mul(c[(size_t)__ldg(p.colidx + s) * S], lu[(size_t)s * S]));
      c[(size_t)k * S] = s_div(ck, lu[(size_t)ds * S]);  // no zero check here in the reference either (mod.rs:978)
    }
    *n_solves += 1;
    // ---- global step limit + update (analysis.rs:197-207 / 283-293); dx[e] = c[col_e2i[e]]
    double max_abs = 0.0;
    for (int k = 0; k < N; k++) {
      const double a = s_abs(c[(size_t)__ldg(p.col_e2i + k) * S]);
      if (a > max_abs) max_abs = a;
    }
    const bool direct = !std::is_same<T, double>::value && ctl.ac_direct;  // linear AC system: one solve is the answer
    const bool limit = max_abs > 1.0 && !direct;
    dx_ok = true;
    for (int k = 0; k < N; k++) {
      T dxk = c[(size_t)__ldg(p.col_e2i + k) * S];
      if (limit) dxk = s_scale(dxk, 1.0, max_abs);
      x[(size_t)k * S] = s_add(x[(size_t)k * S], dxk);
      dx_ok = dx_ok && Tol<T>::ok(s_abs(dxk), vtol);
    }
    if (direct) return ST_OK_;
  }
  return ST_CONV_;
}

// ------------------------------------------------------------------------------------------------ kernels
template <bool B4>
__global__ void __launch_bounds__(128) k_dcop(DevTables d, PlanTables p, WorkTables<double> w, NewtonOut o, SolveCtl ctl) {
  const size_t inst = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= (size_t)ctl.B) return;
  int ns = 0, nl = 0;
  if (ctl.resume) {  // continue the stopped instances only, each with what is left of its own iteration budget
    if (o.status[inst] != ST_REPIVOT_) return;
    const int left = ctl.max_iter - (o.iters[inst] - (o.iters_base ? o.iters_base[inst] : 0));
    if (left <= 0) { o.status[inst] = ST_CONV_; return; }
    ctl.max_iter = left;
  }
  const int st = newton_solve<double, B4>(d, p, w, ctl, inst, 0.0, ctl.reltol, ctl.iabstol, true, &ns, &nl);
  o.status[inst] = st;
  o.iters[inst] += ns;
  o.loads[inst] += nl;
}

template <bool B4>
__global__ void __launch_bounds__(128) k_tran(DevTables d, PlanTables p, WorkTables<double> w, NewtonOut o, SolveCtl ctl, int T,
                                             const int* save_vars, int n_save, double* wave) {
  const size_t inst = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= (size_t)ctl.B) return;
  const size_t B = w.stride;  // wave is [T][n_save][stride], instance fastest, like every other per-instance table
  int st = o.status[inst];  // status of the OP solve
  for (int s = 0; s < n_save; s++) wave[(size_t)s * B + inst] = w.x[(size_t)__ldg(save_vars + s) * w.stride + inst];
  int ns = 0, nl = 0;
  ctl.time = ctl.dt;  // analysis.rs:552-569: t starts at tstep and accumulates tstep (the host's time axis is built the same way)
  for (int tp = 1; tp < T; tp++, ctl.time += ctl.dt) {
    if (st == ST_OK_) st = newton_solve<double, B4>(d, p, w, ctl, inst, 0.0, ctl.reltol, ctl.iabstol, true, &ns, &nl);
    for (int s = 0; s < n_save; s++) {
      const double v = st == ST_OK_ ? w.x[(size_t)__ldg(save_vars + s) * w.stride + inst] : __longlong_as_double(0x7ff8000000000000LL);
      wave[((size_t)tp * n_save + s) * B + inst] = v;
    }
  }
  o.status[inst] = st;
  o.iters[inst] += ns;
  o.loads[inst] += nl;
}

// ---- adaptive transient (opt-in, the fixed-step kernels above are the parity path) ------------------
// One thread per instance runs its OWN time axis: Backward Euler with a local-truncation-error estimate per step, step
// rejection and step-size control on the device, no host round trip. The reference has none of this (fixed step only,
// analysis.rs:553-570; `trtol` / `chgtol` are dead fields, :650-656).
//   LTE estimate: the BE solution against the linear extrapolation x_p through the two previous accepted points;
//     x_BE - x_p = x''/2 * h (2h + h1)  =>  LTE = x''/2 * h^2 = (x_BE - x_p) * h / (2h + h1)
//   accepted when  max_i |LTE_i| / (trtol * (reltol_lte * max(|x_i|, |x_i,prev|) + vntol))  <= 1,
//   next step h * clamp(0.9 / sqrt(ratio), 0.1, 2), at most hmax; a Newton failure retries with h / 8; below hmin the
//   instance reports Convergence Failed.
// Output is on the caller's print grid t_k = k * tstep (the fixed-step axis), by linear interpolation between accepted
// points — every instance fills the same [T][n_save] block whatever steps it took.
struct AdaptCtl {
  double tstep, h0, hmin, hmax, trtol, reltol, vntol;
  int T;                  // print points incl. t = 0
  double *x1, *xs;        // [N][stride] previous accepted solution / solution at the start of the attempt
  double *st_save;        // [n_state][st_stride] committed device state at the start of the attempt
  int32_t *accepted, *rejected;  // [B]
};
template <bool B4>
__global__ void __launch_bounds__(128) k_tran_adaptive(DevTables d, PlanTables p, WorkTables<double> w, NewtonOut o, SolveCtl ctl, AdaptCtl a,
                                                      const int* save_vars, int n_save, double* wave) {
  const size_t inst = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= (size_t)ctl.B) return;
  const size_t S = w.stride, SS = w.st_stride;
  const int N = p.N;
  int st = o.status[inst];
  double* x = w.x + inst;
  double* x1 = a.x1 + inst;
  double* xs = a.xs + inst;
  for (int s = 0; s < n_save; s++) wave[(size_t)s * S + inst] = x[(size_t)__ldg(save_vars + s) * S];
  for (int k = 0; k < N; k++) x1[(size_t)k * S] = x[(size_t)k * S];
  int ns = 0, nl = 0, nacc = 0, nrej = 0, kp = 1;
  double t = 0.0, h = a.h0, h1 = 0.0;
  while (kp < a.T && st == ST_OK_) {
    // start of an attempt: keep x_n and the committed device state
    for (int k = 0; k < N; k++) xs[(size_t)k * S] = x[(size_t)k * S];
    for (int k = 0; k < d.n_state; k++) a.st_save[(size_t)k * SS + inst] = w.st_op[(size_t)k * SS + inst];
    SolveCtl c = ctl;
    c.dt = h;
    c.time = t + h;
    int rc = newton_solve<double, B4>(d, p, w, c, inst, 0.0, ctl.reltol, ctl.iabstol, true, &ns, &nl);  // a weak frozen pivot counts as a failed attempt
    double ratio = 0.0;
    if (rc == ST_OK_ && nacc >= 1) {
      const double coef = h / (2.0 * h + h1);
      for (int k = 0; k < N; k++) {
        const double xn = xs[(size_t)k * S], xo = x1[(size_t)k * S], xb = x[(size_t)k * S];
        const double xp = xn + (xn - xo) * (h / h1);
        const double lte = fabs(xb - xp) * coef;
        const double tol = a.trtol * (a.reltol * fmax(fabs(xb), fabs(xn)) + a.vntol);
        ratio = fmax(ratio, lte / tol);
      }
    }
    if (rc != ST_OK_ || ratio > 1.0) {  // reject: back to x_n and its device state, smaller step
      for (int k = 0; k < N; k++) x[(size_t)k * S] = xs[(size_t)k * S];
      for (int k = 0; k < d.n_state; k++) {  // after a commit op == guess: both go back (the limiters read the in-flight copy)
        const double v = a.st_save[(size_t)k * SS + inst];
        w.st_op[(size_t)k * SS + inst] = v;
        w.st_guess[(size_t)k * SS + inst] = v;
      }
      nrej++;
      h = rc != ST_OK_ ? h * 0.125 : h * fmax(0.1, 0.9 / sqrt(ratio));
      if (h < a.hmin) st = rc == ST_SINGULAR_ ? ST_SINGULAR_ : rc == ST_REPIVOT_ ? ST_REPIVOT_ : ST_CONV_;
      continue;
    }
    // accept: emit the print points passed, shift the history
    const double tn = t + h;
    while (kp < a.T && (double)kp * a.tstep <= tn * (1.0 + 1e-12)) {
      const double f = fmin(1.0, fmax(0.0, ((double)kp * a.tstep - t) / h));
      for (int s = 0; s < n_save; s++) {
        const size_t v = (size_t)__ldg(save_vars + s) * S;
        wave[((size_t)kp * n_save + s) * S + inst] = xs[v] + (x[v] - xs[v]) * f;
      }
      kp++;
    }
    for (int k = 0; k < N; k++) x1[(size_t)k * S] = xs[(size_t)k * S];
    h1 = h;
    t = tn;
    nacc++;
    h = fmin(a.hmax, h * fmin(2.0, fmax(0.1, 0.9 / sqrt(fmax(ratio, 1e-4)))));
  }
  for (; kp < a.T; kp++)  // a failed instance: NaN for the points it never reached (as the fixed-step kernels do)
    for (int s = 0; s < n_save; s++) wave[((size_t)kp * n_save + s) * S + inst] = __longlong_as_double(0x7ff8000000000000LL);
  o.status[inst] = st;
  o.iters[inst] += ns;
  o.loads[inst] += nl;
  a.accepted[inst] = nacc;
  a.rejected[inst] = nrej;
}

__global__ void __launch_bounds__(128) k_ac(DevTables d, PlanTables p, WorkTables<cplx> w, NewtonOut o, SolveCtl ctl) {
  const size_t inst = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= (size_t)ctl.B) return;
  int ns = 0, nl = 0;
  // hard-coded complex tolerances (analysis.rs:271-272); no commit is observable in AC (load_ac reads only `op`)
  const int st = newton_solve<cplx, false>(d, p, w, ctl, inst, ctl.omega[inst], 1e-3, 1e-9, false, &ns, &nl);
  o.status[inst] = st;
  o.iters[inst] += ns;
  o.loads[inst] += nl;
}

// Probe: the first load sweep of one instance with RAW element ids as handles, assembled into that instance's own
// `lu` column (the caller sizes it for max(n_elems, nnz)) and copied out densely. Feeds the host symbolic phase.
template <class T>
__global__ void k_probe(DevTables d, WorkTables<T> w, SolveCtl ctl, int n_elems, int N, int inst_, T* out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const size_t inst = (size_t)inst_;
  const size_t pinst = inst * ctl.par_inst_stride;
  Env<T> e;
  e.pval = d.pval;
  e.sstride = w.st_stride; e.sinst = pinst; e.pinst = pinst;
  e.x = w.x; e.lu = w.lu; e.rhs = w.rhs; e.stride = w.stride; e.w = inst;
  e.mode = ctl.mode; e.dt = ctl.dt; e.gmin = ctl.gmin; e.omega = ctl.omega ? ctl.omega[inst] : 0.0;
  e.time = ctl.mode == AN_TRAN ? ctl.dt : 0.0;  // the transient plan is taken at the first time point
  for (int k = 0; k < n_elems; k++) w.lu[(size_t)k * w.stride + inst] = Scalar<T>::zero();
  for (int k = 0; k < N; k++) w.rhs[(size_t)k * w.stride + inst] = Scalar<T>::zero();
  load_sweep<true>(d, e, w.st_op, w.st_guess);
  for (int k = 0; k < n_elems; k++) out[k] = w.lu[(size_t)k * w.stride + inst];
}

static inline int grid_for(int B, int block) { return (B + block - 1) / block; }

int launch_dcop(const DevTables& d, const PlanTables& p, const WorkTables<double>& w, const NewtonOut& o, const SolveCtl& c, void* stream) {
  if (c.has_bsim4) k_dcop<true><<<grid_for(c.B, 128), 128, 0, (cudaStream_t)stream>>>(d, p, w, o, c);
  else k_dcop<false><<<grid_for(c.B, 128), 128, 0, (cudaStream_t)stream>>>(d, p, w, o, c);
  return (int)cudaGetLastError();
}
int launch_tran(const DevTables& d, const PlanTables& p, const WorkTables<double>& w, const NewtonOut& o, const SolveCtl& c, int T,
                const int* save_vars, int n_save, double* wave, void* stream) {
  if (c.has_bsim4) k_tran<true><<<grid_for(c.B, 128), 128, 0, (cudaStream_t)stream>>>(d, p, w, o, c, T, save_vars, n_save, wave);
  else k_tran<false><<<grid_for(c.B, 128), 128, 0, (cudaStream_t)stream>>>(d, p, w, o, c, T, save_vars, n_save, wave);
  return (int)cudaGetLastError();
}
int launch_tran_adaptive(const DevTables& d, const PlanTables& p, const WorkTables<double>& w, const NewtonOut& o, const SolveCtl& c,
                         const AdaptiveArgs& g, const int* save_vars, int n_save, double* wave, void* stream) {
  AdaptCtl a;
  a.tstep = g.tstep; a.h0 = g.h0; a.hmin = g.hmin; a.hmax = g.hmax; a.trtol = g.trtol; a.reltol = g.reltol; a.vntol = g.vntol; a.T = g.T;
  a.x1 = g.x1; a.xs = g.xs; a.st_save = g.st_save; a.accepted = g.accepted; a.rejected = g.rejected;
  if (c.has_bsim4) k_tran_adaptive<true><<<grid_for(c.B, 128), 128, 0, (cudaStream_t)stream>>>(d, p, w, o, c, a, save_vars, n_save, wave);
  else k_tran_adaptive<false><<<grid_for(c.B, 128), 128, 0, (cudaStream_t)stream>>>(d, p, w, o, c, a, save_vars, n_save, wave);
  return (int)cudaGetLastError();
}
int launch_ac(const DevTables& d, const PlanTables& p, const WorkTables<cplx>& w, const NewtonOut& o, const SolveCtl& c, void* stream) {
  k_ac<<<grid_for(c.B, 128), 128, 0, (cudaStream_t)stream>>>(d, p, w, o, c);
  return (int)cudaGetLastError();
}
// Result packing on the device, so that the host needs ONE contiguous copy and no strided pass:
//   out = [ x as row-major [instance][variable] : B*N f64 ][ status : B i32 ][ iters : B i32 ][ loads : B i32 ]
__global__ void k_pack_out(const double* __restrict__ x, const int32_t* __restrict__ status, const int32_t* __restrict__ iters,
                           const int32_t* __restrict__ loads, double* __restrict__ out, int N, size_t stride, int B) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B) return;
  for (int k = 0; k < N; k++) out[i * (size_t)N + (size_t)k] = x[(size_t)k * stride + i];
  int32_t* tail = reinterpret_cast<int32_t*>(out + (size_t)B * (size_t)N);
  tail[i] = status[i];
  tail[(size_t)B + i] = iters[i];
  tail[2 * (size_t)B + i] = loads[i];
}
int launch_pack_out(const double* x, const int32_t* status, const int32_t* iters, const int32_t* loads, double* out, int N, size_t stride, int B,
                    void* stream) {
  k_pack_out<<<grid_for(B, 128), 128, 0, (cudaStream_t)stream>>>(x, status, iters, loads, out, N, stride, B);
  return (int)cudaGetLastError();
}
// Waveforms leave the time loops as [time point][saved variable][instance] (instance fastest: coalesced stores from every
// kernel family); the caller's layout is [instance][time point][saved variable]. A tiled transpose of the [M][stride]
// matrix, M = T * n_save, so that the host needs one contiguous copy instead of a strided pass over 40 MB (C1 sweep:
// 8192 instances x 201 points x 3 signals took ~20 ms on the host).
template <class T>
__global__ void k_pack_wave(const T* __restrict__ w, T* __restrict__ out, size_t M, size_t stride, int B) {
  __shared__ T tile[32][33];
  const size_t m0 = (size_t)blockIdx.y * 32, i0 = (size_t)blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const size_t m = m0 + r, i = i0 + threadIdx.x;
    tile[r][threadIdx.x] = (m < M && i < (size_t)B) ? w[m * stride + i] : Scalar<T>::zero();
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const size_t i = i0 + r, m = m0 + threadIdx.x;
    if (i < (size_t)B && m < M) out[i * M + m] = tile[threadIdx.x][r];
  }
}
template <class T>
static int pack_rows(const T* w, T* out, size_t M, size_t stride, int B, void* stream) {
  if (M == 0 || B <= 0) return 0;
  const size_t gy = (M + 31) / 32;
  if (gy > 65535) return (int)cudaErrorInvalidConfiguration;  // the caller falls back to the host transpose
  dim3 grid((unsigned)((B + 31) / 32), (unsigned)gy), block(32, 8);
  k_pack_wave<T><<<grid, block, 0, (cudaStream_t)stream>>>(w, out, M, stride, B);
  return (int)cudaGetLastError();
}
int launch_pack_wave(const double* wave, double* out, size_t M, size_t stride, int B, void* stream) { return pack_rows<double>(wave, out, M, stride, B, stream); }
// the same for an AC sweep: x [variable][frequency point] (complex) -> out [frequency point][variable]
int launch_pack_ac(const cplx* x, cplx* out, size_t N, size_t stride, int F, void* stream) { return pack_rows<cplx>(x, out, N, stride, F, stream); }
int launch_probe_real(const DevTables& d, const WorkTables<double>& w, const SolveCtl& c, int n_elems, int N, int inst, double* out, void* stream) {
  k_probe<double><<<1, 32, 0, (cudaStream_t)stream>>>(d, w, c, n_elems, N, inst, out);
  return (int)cudaGetLastError();
}
int launch_probe_cplx(const DevTables& d, const WorkTables<cplx>& w, const SolveCtl& c, int n_elems, int N, int inst, cplx* out, void* stream) {
  k_probe<cplx><<<1, 32, 0, (cudaStream_t)stream>>>(d, w, c, n_elems, N, inst, out);
  return (int)cudaGetLastError();
}

// ---- self-test of the split division (scalar.h): s_div / s_rcp + s_div_r against the compiler's `a / b`, bit for bit
__device__ __forceinline__ unsigned long long st_mix(unsigned long long& x) {
  x += 0x9E3779B97F4A7C15ull;
  unsigned long long z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double st_draw(unsigned long long& x, int kind) {
  const unsigned long long r = st_mix(x);
  const double u = (double)(r >> 11) * (1.0 / 9007199254740992.0);
  switch (kind & 7) {
    case 0: return __longlong_as_double((long long)r);                                     // any bit pattern (NaN, inf, denormals)
    case 1: return (u - 0.5) * 4.0;                                                        // order 1
    case 2: return (u - 0.5) * exp((double)((long long)(st_mix(x) % 1400) - 700));         // whole normal range
    case 3: return __longlong_as_double((long long)(r & 0x800fffffffffffffull));           // denormals and zeros
    case 4: return (r & 1) ? 0.0 : -0.0;
    case 5: return __longlong_as_double((long long)((r & 0x800fffffffffffffull) | 0x7fe0000000000000ull));  // near overflow
    case 6: return __longlong_as_double((long long)((r & 0x800fffffffffffffull) | 0x0360000000000000ull));  // at the fast-path edge
    default: return (double)((long long)(r % 2001) - 1000) * 1e-3;                         // small grid incl. exact zero
  }
}
__global__ void k_selftest_div(unsigned long long n_per_thread, unsigned long long seed, unsigned long long* bad, double* first_bad) {
  unsigned long long x = seed + 0x632BE59BD9B4E019ull * (unsigned long long)(blockIdx.x * blockDim.x + threadIdx.x + 1);
  for (unsigned long long i = 0; i < n_per_thread; i++) {
    const int ka = (int)(st_mix(x) & 7), kb = (int)(st_mix(x) & 7);
    const double a = st_draw(x, ka), a2 = st_draw(x, ka), b = st_draw(x, kb);
    const double r = s_rcp(b);
    bool flagged = false;  // s_div_rf: whenever it does not raise its flag the value must be the IEEE quotient
    const double qf = s_div_rf(a2, b, r, flagged);
    bool fs = false, fe = false;  // likewise the branch-free sqrt / exp against the library calls
    const double x_e = ka == 7 ? a * 700.0 : a;
    const double sq = s_sqrt_f(a2, fs), ex = s_exp_f(x_e, fe);
    bool fp = false;
    const double qp = s_div_rp(a2, b, r, s_div_bok(b), fp);
    const double q[7] = {s_div(a, b), s_div_r(a2, b, r), s_scale(a, 1.0, b), flagged ? a2 / b : qf, fs ? sqrt(a2) : sq, fe ? exp(x_e) : ex, fp ? a2 / b : qp};
    const double w[7] = {a / b, a2 / b, a * 1.0 / b, a2 / b, sqrt(a2), exp(x_e), a2 / b};
    for (int k = 0; k < 7; k++) {
      const bool same = (q[k] != q[k] && w[k] != w[k]) || __double_as_longlong(q[k]) == __double_as_longlong(w[k]);
      if (!same && atomicAdd(bad, 1ull) == 0ull) { first_bad[0] = (k == 1 || k == 3 || k == 4 || k == 6) ? a2 : (k == 5 ? x_e : a); if (k >= 4) first_bad[1] = (double)k; first_bad[1] = b; first_bad[2] = q[k]; first_bad[3] = w[k]; }
    }
  }
}
int selftest_div(unsigned long long n, unsigned long long seed, unsigned long long* mismatches, double* first4) {
  unsigned long long* d_bad = nullptr;
  double* d_first = nullptr;
  cudaError_t e = cudaMalloc(&d_bad, sizeof(unsigned long long));
  if (e != cudaSuccess) return (int)e;
  e = cudaMalloc(&d_first, 4 * sizeof(double));
  if (e != cudaSuccess) { cudaFree(d_bad); return (int)e; }
  cudaMemset(d_bad, 0, sizeof(unsigned long long));
  cudaMemset(d_first, 0, 4 * sizeof(double));
  const int threads = 256, blocks = 592;
  const unsigned long long per = (n + (unsigned long long)threads * blocks - 1) / ((unsigned long long)threads * blocks);
  k_selftest_div<<<blocks, threads>>>(per, seed, d_bad, d_first);
  e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(mismatches, d_bad, sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(first4, d_first, 4 * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d_bad);
  cudaFree(d_first);
  return (int)e;
}

}  // namespace s21
