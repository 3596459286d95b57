// Host-visible interface of the CUDA engine (kernels/newton.cu). Plain structs of device pointers; no CUDA headers
// needed by the includer.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>

#include "../scalar.h"

namespace s21 {

// Device tables shared by every instance of a batch (device_layout.h). `itab` holds element handles that index the
// value array the kernel assembles into: raw element ids for a probe, L+U slots for a planned solve.
struct DevTables {
  int n_dev;
  const int* type;
  const int* itab_off;
  const int* par_off;
  const int* state_off;
  const int* itab;
  const int* pcode;     // (offset << 1) | per_instance
  const double* pval;   // parameter value pool
  int n_state;
  // [n_dev] 1 when every parameter of the device is a shared value (no per-instance column): its block is then the
  // contiguous run pval[par_off .. par_off + n_par) and can be read with ONE load at a literal offset instead of the
  // code lookup + address arithmetic + load (ncu, C4: 37 % of the executed instructions of a Bsim4 evaluation were that
  // arithmetic). nullptr = no device is direct.
  const int* par_direct = nullptr;
};

// Static plan of the sparse LU (host/symbolic.hpp), in internal (pivoted) coordinates.
struct PlanTables {
  int N, nnz;
  const int *row_i2e, *col_i2e, *col_e2i;
  const int *rowptr, *colidx, *diag_slot;
  const int *l_off, *l_slot, *l_row;
  const int* piv_chk;  // [N] pivot k passed the reference's threshold test when it was chosen: the pivot-health test applies to it
  const int *upd_off, *upd_t, *upd_u, *upd_l;
};

// Per-instance workspace, structure-of-arrays with the instance index fastest: entry k of instance i is a[k*stride+i].
template <class T>
struct WorkTables {
  T* x;     // [N]   solution / Newton guess (persists across solves: warm start, analysis.rs:139-147)
  T* rhs;   // [N]
  T* c;     // [N]   residual -> forward/back substitution -> dx, internal order
  T* lu;    // [nnz] assembled matrix, factorised in place
  size_t stride;
  double* st_op;     // [n_state] committed device state (Component::commit, comps/mod.rs:77)
  double* st_guess;  // [n_state] in-flight device state
  size_t st_stride;
};

// Kernel-internal status: a frozen pivot fell below the reference's 1e-3 threshold against an entry under it (the
// reference re-runs its pivot search in every factorisation, sparse21/mod.rs:929-932, 735-783, and would have chosen
// otherwise). The instance stops before the update, x at the current iterate; the host takes a new pivot order from that
// iterate and continues the solve (host/batch.hpp resolve_repivot). Callers never see this code.
enum { ST_REPIVOT_CODE = 9 };

struct NewtonOut {
  int32_t* status;    // [B] S21_* code
  // resume launches (SolveCtl::resume) only: [B] value of iters[] when the solve being continued started (nullptr = zeros),
  // so that an instance's budget is what is left of ITS 100 iterations
  const int32_t* iters_base = nullptr;
  int32_t* iters;     // [B] iterations that reached the linear solve (accumulated)
  int32_t* loads;     // [B] device-load sweeps (accumulated)
};

struct SolveCtl {
  int B;             // instances (threads)
  int mode;          // AnMode
  double gmin, dt;
  double time = 0.0;       // direct kernels: the time point being solved (they carry it here; the others keep it in their time loop)
  double reltol, iabstol;  // real solve: |dx| <= reltol (absolute!) and |res| <= iabstol (analysis.rs:331-345)
  const double* omega;     // [B] AC only
  size_t par_inst_stride;  // 1: parameter/state instance == workspace instance; 0: all columns use instance 0 (AC sweep)
  // AC only. 1 (default): the small-signal system is linear, so each frequency point is ONE factorisation + solve from
  // x = 0 with no step limit (x = A^-1 b). 0: the reference's Newton shell (analysis.rs:253-303: 1.0 step limit, <= 20
  // iterations, a second factorisation to confirm) — kept for iteration-count parity; it cannot reach |x| > ~19 from a cold start.
  int ac_direct = 0;
  // 1: stop an instance at a weak frozen pivot (ST_REPIVOT_CODE) — set where the host can re-pivot and continue (dcop, the OP
  // that opens a transient / AC sweep, AC points); 0: carry on with the frozen order (inside a device-resident time loop there
  // is nobody to hand the instance to; round 1's behaviour)
  int stop_on_weak = 0;
  int max_iter = 100;      // real solves: iteration budget of this launch (analysis.rs:173 caps a solve at 100; a solve continued
                           // after a re-pivot gets what is left)
  // A frozen pivot is "weak" when |pivot| * weak_mult < |an entry below it|, i.e. when a multiplier of the elimination exceeds
  // weak_mult. 1e3 (+ a margin that keeps a pivot AT the host's threshold unflagged) is the reference's own acceptance test
  // (sparse21/mod.rs:735-783): every factorisation the reference — which searches its pivots anew each time — would have
  // ordered differently is re-ordered here too, so iteration counts and the outcome of the 100-iteration cap follow the
  // reference's. S21_PIVOT_MULT=<bound> loosens it (experiments): measured with 1e10 the answers of converging instances
  // stay within tolerance (the convergence test is on the true residual) but iteration counts drift by up to 15 %, instances
  // converge that the reference caps, and a long Mos1 chain returns NaNs as "converged" (the reference's tolerance test lets
  // NaN through, analysis.rs:331-345) — profiles/r02q_*; the bound therefore stays at the reference's figure.
  double weak_mult = 1.000001e3;
  // 1 (dcop kernels with plan tables only): continue a solve in place. Only instances whose status is ST_REPIVOT_CODE run — warm,
  // from the iterate where they stopped, against the pivot order of THIS launch's plan, with the iterations their solve has left
  // (max_iter - (iters - iters_base)); every other instance keeps its x, device state, status and counters untouched.
  int resume = 0;
  // Cooperative transient kernel only. 1: an instance whose frozen pivot order fails INSIDE the time loop (an exactly zero pivot,
  // or a non-finite step from a vanishing one) is not ended with Singular Matrix: it goes back to its last accepted time
  // point and stops with ST_REPIVOT_CODE and the time point recorded, and the host continues it — pivot order taken at that
  // iterate, a resume launch (resume = 1) that runs only those instances from their own time point on.
  int tran_stop = 0;
  int tran_inject_tp = 0;  // test hook (S21_TRAN_INJECT): every instance is handed back once, in the second iteration of this time point
  int relaxed = 0;         // the plan's level schedules are in tolerance mode (host/symbolic.hpp build_levels): apply updates atomically
  int has_bsim4 = 0;       // selects the kernel build that links the Bsim4 evaluation (kept out of the others: register pressure)
};

// Host-side: the multiplier bound of SolveCtl::weak_mult; read at every solve setup / kernel generation (the generated text
// carries the figure, so the cubin cache keys on it).
inline double pivot_weak_mult() {
  if (const char* e = std::getenv("S21_PIVOT_MULT")) { const double v = std::atof(e); if (v >= 1.0) return v; }
  return 1.000001e3;
}

// Extra shared tables of the cooperative kernel (kernels/coop.cu): staged assembly + level schedules (host/symbolic.hpp).
struct CoopTables {
  int n_stage;               // staging slots (sum over devices of their itab length, +1 per Mos1)
  const int* stage_off;      // [n_dev] first staging slot of each device
  const int* eval_order;     // [n_dev] device ids sorted by type (keeps warps convergent when a warp spans devices)
  const int *asm_off, *asm_src;  // [nnz + N + 1], gather lists
  int n_lu_lvl, n_fw_lvl, n_bw_lvl;
  const int *lu_lvl_off, *lu_t, *lu_u, *lu_l;
  const int *fw_lvl_off, *fw_k, *fw_row, *fw_slot;
  const int *bw_lvl_off, *bw_row;
  // tolerance-mode plans only (grid-wide kernel; nullptr otherwise): per pivoted row, those of its L+U slots that receive
  // stamps — the entries of A proper. The residual then skips the fill-in slots (exact zeros before the factorisation:
  // 4.0 M of the 4.07 M slots of config C3)
  const int *res_off = nullptr, *res_slot = nullptr, *res_x = nullptr;  // res_x: the variable each of those entries multiplies
};

// Launch geometry of the cooperative kernel: a CTA of `threads` threads (a multiple of `gi`) owns `gi` consecutive
// instances and keeps their whole workspace in `smem_bytes` of shared memory (0 = workspace stays in HBM, read through
// L1). All shared index tables (DevTables ints, PlanTables, CoopTables) must live in ONE device allocation, the arena;
// with arena_in_smem the kernel TMA-copies it into shared memory and rebases the table pointers.
struct CoopCfg {
  int gi, threads;
  size_t smem_bytes;
  const int* arena;
  size_t arena_bytes;  // multiple of 16
  bool arena_in_smem;
  // cooperative kernel: > 0 = copy only this prefix of the arena (everything but the parameter-code table, which is last and,
  // for Bsim4 circuits, larger than all the other tables together) — multiple of 16
  size_t arena_core_bytes = 0;
  bool cold = false;   // hybrid kernel only: start from x = 0 / zero device state instead of loading them (a folded reset)
  // cooperative kernel: smem_bytes covers x, rhs, residual and the L+U values only (coop_mixed_bytes); the stamp staging area
  // (`stage` must be given) and the device state stay in HBM/L2
  bool mixed = false;
  // cooperative transient kernel with SolveCtl::tran_stop / resume (CoopArgs::tp_stop / x_acc)
  int* tp_stop = nullptr;
  double* x_acc = nullptr;
};
// Shared-memory budget of one CTA: control words + (optional) arena copy + workspace.
size_t coop_ctrl_bytes(int gi);
size_t coop_work_bytes(int N, int nnz, int n_stage, int n_state, int gi, int scalar_width);
size_t coop_mixed_bytes(int N, int nnz, int gi, int scalar_width);
int coop_max_smem_optin(int device);

// All launchers enqueue on `stream` (a cudaStream_t) and return a cudaError_t as int.
// Cooperative variants: same contracts as launch_dcop / launch_tran / launch_ac, bit-identical results.
int launch_coop_dcop(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<double>& w, double* stage,
                     const NewtonOut& o, const SolveCtl& c, const CoopCfg& cfg, void* stream);
int launch_coop_tran(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<double>& w, double* stage,
                     const NewtonOut& o, const SolveCtl& c, const CoopCfg& cfg, int T, const int* save_vars, int n_save, double* wave,
                     void* stream);
int launch_coop_ac(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<cplx>& w, cplx* stage,
                   const NewtonOut& o, const SolveCtl& c, const CoopCfg& cfg, void* stream);
// kernels/coop_fast.cu (Bsim4 circuits only, S21_B4_FAST=1): divisions of the Bsim4 evaluation as a * rcp(b); same contracts,
// results within round-off of the default kernels instead of bit-identical
int launch_coop_dcop_fast(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<double>& w, double* stage,
                          const NewtonOut& o, const SolveCtl& c, const CoopCfg& cfg, void* stream);
int launch_coop_tran_fast(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<double>& w, double* stage,
                          const NewtonOut& o, const SolveCtl& c, const CoopCfg& cfg, int T, const int* save_vars, int n_save, double* wave,
                          void* stream);
// Hybrid variants (kernels/hybrid.cu) for small circuits: 32 instances per 256-thread CTA, workspace + arena always in
// shared memory (cfg.gi / cfg.threads are ignored; cfg.smem_bytes = hybrid_work_bytes(...)). Bit-identical results.
size_t hybrid_smem_bytes(int N, int nnz, int n_stage, int n_state, size_t arena_bytes, int scalar_width);
size_t hybrid_work_bytes(int N, int nnz, int n_stage, int n_state, int scalar_width);
int launch_hybrid_dcop(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<double>& w, const NewtonOut& o,
                       const SolveCtl& c, const CoopCfg& cfg, void* stream);
int launch_hybrid_tran(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<double>& w, const NewtonOut& o,
                       const SolveCtl& c, const CoopCfg& cfg, int T, const int* save_vars, int n_save, double* wave, void* stream);
int launch_hybrid_ac(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<cplx>& w, const NewtonOut& o,
                     const SolveCtl& c, const CoopCfg& cfg, void* stream);
// Grid-wide variant (kernels/grid.cu) for ONE large circuit (B = 1, instance stride 1, workspace and staging in HBM/L2):
// a cooperative launch with grid barriers between phases / dependency levels. `gc` is a device-resident control block.
struct GridCtl {
  static constexpr int kHugeCap = 64;
  int stat, nsol, nld, dxok, act, resok, sing, convnow, weak;
  int n_huge;             // tolerance mode: gather lists too long for one warp, summed in chunks by the whole grid (grid.cu)
  unsigned long long maxabs;
  int huge[kHugeCap];
  // wall clock per phase of the launch in ns (thread 0, %globaltimer): 0 evaluation, 1 assembly, 2 residual + decision, 3 LU,
  // 4 forward, 5 backward substitution, 6 step limit + update
  static constexpr int kPhases = 7;
  unsigned long long phase_ns[kPhases];
  // the same per dependency level (the first kLevels of each schedule): LU, forward, backward
  static constexpr int kLevels = 24;
  unsigned long long level_ns[3][kLevels];
};
int launch_grid_dcop(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<double>& w, double* stage, const NewtonOut& o,
                     const SolveCtl& c, GridCtl* gc, void* stream);
int launch_grid_tran(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<double>& w, double* stage, const NewtonOut& o,
                     const SolveCtl& c, GridCtl* gc, int T, const int* save_vars, int n_save, double* wave, void* stream);
int launch_dcop(const DevTables& d, const PlanTables& p, const WorkTables<double>& w, const NewtonOut& o, const SolveCtl& c, void* stream);
// OP must already be solved and committed; runs points 1..T-1 of Tran::solve. wave: [T][n_save][w.stride] device (point 0 written too).
int launch_tran(const DevTables& d, const PlanTables& p, const WorkTables<double>& w, const NewtonOut& o, const SolveCtl& c, int T,
                const int* save_vars, int n_save, double* wave, void* stream);
// Adaptive-step transient (kernels/newton.cu::k_tran_adaptive; opt-in): per-instance step control by local truncation
// error on the device. Scratch arrays are device memory sized like the workspace.
struct AdaptiveArgs {
  double tstep, h0, hmin, hmax, trtol, reltol, vntol;
  int T;
  double *x1, *xs, *st_save;
  int32_t *accepted, *rejected;
};
int launch_tran_adaptive(const DevTables& d, const PlanTables& p, const WorkTables<double>& w, const NewtonOut& o, const SolveCtl& c,
                         const AdaptiveArgs& g, const int* save_vars, int n_save, double* wave, void* stream);
int launch_ac(const DevTables& d, const PlanTables& p, const WorkTables<cplx>& w, const NewtonOut& o, const SolveCtl& c, void* stream);
// Result packing on the device: out = [x as [instance][variable], B*N f64][status B i32][iters B i32][loads B i32].
int launch_pack_out(const double* x, const int32_t* status, const int32_t* iters, const int32_t* loads, double* out, int N, size_t stride, int B,
                    void* stream);
// Waveform packing on the device: wave [M = T * n_save][stride] (instance fastest) -> out [B][M], the caller's layout.
int launch_pack_wave(const double* wave, double* out, size_t M, size_t stride, int B, void* stream);
int launch_pack_ac(const cplx* x, cplx* out, size_t N, size_t stride, int F, void* stream);  // x [N][stride] -> out [F][N]
// First load sweep of instance `inst` in `mode`, assembled by raw element id into out[n_elems] (device memory).
int launch_probe_real(const DevTables& d, const WorkTables<double>& w, const SolveCtl& c, int n_elems, int N, int inst, double* out, void* stream);
int launch_probe_cplx(const DevTables& d, const WorkTables<cplx>& w, const SolveCtl& c, int n_elems, int N, int inst, cplx* out, void* stream);

// Device self-test of scalar.h's split division against the compiler's `a / b` on ~n random / edge-case pairs (three
// quotients per pair); *mismatches must come back 0. first4 = {a, b, ours, compiler's} of the first mismatch.
int selftest_div(unsigned long long n, unsigned long long seed, unsigned long long* mismatches, double* first4);

}  // namespace s21
