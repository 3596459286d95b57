// Cooperative Newton kernel: a CTA owns `gi` consecutive circuit instances and ALL of its threads work on them, one
// phase at a time, with the instances' whole workspace (x, rhs, residual, L+U values, stamp staging, device state)
// resident in shared memory for the entire Newton loop — and, for transient, for the entire time loop. The shared
// index tables of the circuit (device tables, stamp handles, gather lists, LU plan and level schedules — one packed
// "arena") are brought into shared memory once per CTA with a TMA bulk copy (cp.async.bulk + mbarrier).
//
// Why: with one thread per instance (newton.cu) the 8192-instance benchmark batch is only 256 warps on 148 SMs and
// every warp walks ~9 k dependent instructions per iteration, 45 % of them 64-bit address arithmetic
// (profiles/r01a_*: 6 % warps active, 0.03 % DRAM). Here the work of an iteration is spread over (item, instance) pairs,
// instance fastest, every thread keeping one fixed instance column:
//     eval      item = device            (devices evaluated in parallel into private staging slots)
//     assemble  item = L+U slot | rhs row (gather of staging slots in the reference's accumulation order)
//     residual  item = row
//     LU        item = operation of the current dependency level (host/symbolic.hpp build_levels)
//     forward / backward substitution: item = operation / row of the current level
//     update    item = variable
// With gi = 32 a warp is one item x 32 instances (convergent, conflict-free smem columns); with smaller gi a warp spans
// several items, which buys more CTAs when the batch is small. Every floating-point operation on any single value
// happens in the same order as in newton.cu (and hence as in the reference): results are bit-identical.
//
// Replaces the same reference functions as newton.cu (analysis.rs:153-210, 253-303, 331-345, 553-570;
// sparse21/mod.rs:272-327, 865-991).
#include <cstdlib>

#include "coop_common.cuh"

namespace s21 {

using namespace coopk;

namespace {

// WS: where the per-instance workspace lives. 0 = all of it in HBM/L2; 1 = all of it in shared memory; 2 = mixed — x, rhs,
// residual and the L+U values in shared memory (the linear algebra's barriers then wait on on-chip latencies), the stamp
// staging area and the device state in HBM/L2 (Bsim4 circuits: one staging slot per stamp is 20 KB per instance).
template <class T, int KIND, int WS, bool B4>
__global__ void __launch_bounds__(B4 ? 320 : 256, B4 ? 1 : 2) k_coop(DevTables d, PlanTables p, CoopTables ct, WorkTables<T> g, T* gstage, NewtonOut o,
                                                SolveCtl ctl, CoopArgs a) {
  typedef typename std::conditional<WS != 0, unsigned, size_t>::type I;    // x, rhs, c, lu
  typedef typename std::conditional<WS == 1, unsigned, size_t>::type IS;   // stamp staging, device state
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int gi = a.gi;
  const int li = tid % gi;              // this thread's instance column, fixed for the whole kernel (nt % gi == 0)
  const int item0 = tid / gi;           // first item of this thread in every phase
  const int istep = nt / gi;            // items handled per sweep of the CTA
  const int i0 = blockIdx.x * gi;
  const int ni = min(gi, ctl.B - i0);
  const int N = p.N, nnz = p.nnz;
  constexpr bool real_kind = KIND != K_AC;

  // ---- carve shared memory: [control words][arena copy][workspace]
  double* maxabs = (double*)smem_raw;
  int* act = (int*)(maxabs + gi);
  int* resok = act + gi;
  int* dxok = resok + gi;
  int* sing = dxok + gi;
  int* stat = sing + gi;
  int* nsol = stat + gi;
  int* nld = nsol + gi;
  int* convnow = nld + gi;
  int* weak = convnow + gi;  // pivot health: a frozen pivot the reference's threshold test would have refused (mod.rs:735-783)
  int* left = weak + gi;     // resume: iterations this instance's solve has left; < 0 = not a stopped instance, leave it alone
  int* start = left + gi;    // transient: first time point this instance takes part in (1; a resumed instance: where it stopped)
  int* stopat = start + gi;  // transient: time point at which this launch stopped the instance for the host to re-pivot (0 = not)
  uint64_t* mbar = (uint64_t*)(((size_t)(stopat + gi) + 7) & ~(size_t)7);  // 12 int arrays: 8-byte alignment is not automatic
  size_t off = ((size_t)((unsigned char*)(mbar + 1) - smem_raw) + 15) / 16 * 16;
  if (a.arena_bytes > 0) {
    int* sa = (int*)(smem_raw + off);
    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(mbar, (uint32_t)a.arena_bytes);
      tma_load_1d(sa, a.arena, (uint32_t)a.arena_bytes, mbar);
    }
    off += (size_t)a.arena_bytes;
    // Rebase every table pointer onto the shared-memory copy. The new pointer must be DERIVED FROM `sa`: nvcc assumes
    // pointers that come from kernel arguments address global memory and would emit ld.global for them.
#define RB(ptr) ptr = sa + ((ptr) - a.arena)
    RB(d.type); RB(d.itab_off); RB(d.par_off); RB(d.state_off); RB(d.itab); if (!a.pcode_global) RB(d.pcode); if (d.par_direct) RB(d.par_direct);
    RB(p.row_i2e); RB(p.col_i2e); RB(p.col_e2i); RB(p.rowptr); RB(p.colidx); RB(p.diag_slot);
    RB(ct.stage_off); RB(ct.eval_order); RB(ct.asm_off); RB(ct.asm_src);
    RB(ct.lu_lvl_off); RB(ct.lu_t); RB(ct.lu_u); RB(ct.lu_l);
    RB(ct.fw_lvl_off); RB(ct.fw_k); RB(ct.fw_row); RB(ct.fw_slot); RB(ct.bw_lvl_off); RB(ct.bw_row);
#undef RB
  }
  // ---- workspace views: entry k of this thread's instance is a[k * ws + col]
  T *x, *rhs, *c, *lu, *S;
  double *sop, *sguess;
  I ws, col;
  IS wS, colS, ss, scol;
  if constexpr (WS != 0) {
    x = (T*)(smem_raw + off); off += sizeof(T) * (size_t)N * gi;
    rhs = (T*)(smem_raw + off); off += sizeof(T) * (size_t)N * gi;
    c = (T*)(smem_raw + off); off += sizeof(T) * (size_t)N * gi;
    lu = (T*)(smem_raw + off); off += sizeof(T) * (size_t)nnz * gi;
    ws = (I)gi; col = (I)li;
  } else {
    x = g.x; rhs = g.rhs; c = g.c; lu = g.lu;
    ws = (I)g.stride; col = (I)i0 + (I)li;
  }
  if constexpr (WS == 1) {
    S = (T*)(smem_raw + off); off += sizeof(T) * (size_t)ct.n_stage * gi;
    sop = (double*)(smem_raw + off); off += sizeof(double) * (size_t)d.n_state * gi;
    sguess = (double*)(smem_raw + off);
    wS = (IS)gi; colS = (IS)li; ss = (IS)gi; scol = (IS)li;
  } else {
    // The staging area is scratch of this launch: each CTA gets ONE contiguous block ([slot][instance of the CTA], 20 KB x gi on
    // config C4) instead of columns of the batch-wide [slot][stride] table, whose entries of one instance lie a whole batch
    // stride (16 KB at 2048 instances) apart. S21_COOP_STAGE_BLOCKED=0 (CoopArgs::stage_blocked) keeps the strided layout.
    if (a.stage_blocked) { S = gstage + (size_t)blockIdx.x * (size_t)ct.n_stage * (size_t)gi; wS = (IS)gi; colS = (IS)li; }
    else { S = gstage; wS = (IS)g.stride; colS = (IS)i0 + (IS)li; }
    sop = g.st_op; sguess = g.st_guess;
    ss = (IS)g.st_stride; scol = ((IS)i0 + (IS)li) * (IS)ctl.par_inst_stride;
  }
  const bool valid = li < ni;

  // ---- prologue
  if (tid < gi) {
    stat[tid] = (tid < ni && KIND == K_TRAN) ? o.status[i0 + tid] : 0;
    left[tid] = 1 << 30;
    if (KIND == K_DCOP && ctl.resume) {  // continue the stopped instances only (SolveCtl::resume)
      const int s0 = tid < ni ? o.status[i0 + tid] : 0;
      if (tid < ni && s0 == CST_REPIVOT) {
        const int l = ctl.max_iter - (o.iters[i0 + tid] - (o.iters_base ? o.iters_base[i0 + tid] : 0));
        stat[tid] = l > 0 ? CST_OK : CST_CONV;
        left[tid] = l > 0 ? l : -1;
      } else {
        stat[tid] = s0;
        left[tid] = -1;
      }
    }
    start[tid] = 1; stopat[tid] = 0;
    if (KIND == K_TRAN && ctl.resume) {  // continue, from the time point where they stopped, the instances a transient launch handed back
      const int s0 = tid < ni ? o.status[i0 + tid] : 0;
      if (tid < ni && s0 == CST_REPIVOT) { stat[tid] = CST_OK; start[tid] = a.tp_stop[i0 + tid]; }
      else { stat[tid] = s0; left[tid] = -1; }
    }
    weak[tid] = 0;
    nsol[tid] = 0; nld[tid] = 0; convnow[tid] = 0; dxok[tid] = 1; act[tid] = 0;
  }
  if constexpr (WS != 0) {
    for (int k = item0; k < N; k += istep) x[(I)k * ws + col] = valid ? g.x[(size_t)k * g.stride + i0 + li] : Scalar<T>::zero();
  }
  if constexpr (WS == 1) {
    for (int k = item0; k < d.n_state; k += istep) {
      const size_t src = (size_t)k * g.st_stride + ((size_t)i0 + (size_t)(valid ? li : 0)) * ctl.par_inst_stride;
      sop[(IS)k * ss + scol] = g.st_op[src];
      sguess[(IS)k * ss + scol] = g.st_guess[src];
    }
  }
  if (a.arena_bytes > 0) mbar_wait(mbar, 0);
  __syncthreads();
  if constexpr (KIND == K_TRAN) {
    if (!ctl.resume)
      for (int s = item0; s < a.n_save; s += istep)
        if (valid) a.wave[(size_t)s * g.stride + i0 + li] = x[(I)a.save_vars[s] * ws + col];
  }
  const double vtol = real_kind ? ctl.reltol : 1e-3, itol = real_kind ? ctl.iabstol : 1e-9;  // analysis.rs:271-272, 331-345
  const int n_points = KIND == K_TRAN ? a.T_points : 2;
  const size_t pinst = ((size_t)i0 + (size_t)li) * ctl.par_inst_stride;
  double omega = 0.0;
  if constexpr (KIND == K_AC) omega = valid ? ctl.omega[i0 + li] : 0.0;

  double tnow = KIND == K_TRAN ? ctl.dt : 0.0;  // analysis.rs:552-569: t starts at tstep and accumulates tstep
  for (int tp = 1; tp < n_points; tp++, tnow += ctl.dt) {
    if (tid < gi) { act[tid] = (tid < ni && stat[tid] == CST_OK && left[tid] > 0 && tp >= start[tid]) ? 1 : 0; dxok[tid] = 1; }
    __syncthreads();
    if constexpr (KIND == K_TRAN && real_kind) {
      if (ctl.resume && !__syncthreads_or(tid < gi && act[tid])) continue;  // nothing of this CTA has reached its time point yet
      // the last accepted point: what a stopped instance goes back to (SolveCtl::tran_stop)
      if (ctl.tran_stop && act[li] != 0)
        for (int k = item0; k < N; k += istep) a.x_acc[(size_t)k * g.stride + i0 + li] = x[(I)k * ws + col];
    }
    const int max_it = real_kind ? min(TolC<T>::max_iter, ctl.max_iter) : TolC<T>::max_iter;
    for (int iter = 0; iter < max_it; iter++) {
      const bool on = act[li] != 0;  // stable until the decision phase (which is fenced by barriers on both sides)
      // ---- P1: device evaluation (Solver::update, analysis.rs:153-168), devices in parallel
      if (on) {
        for (int item = item0; item < d.n_dev; item += istep) {
          const int dev = ct.eval_order[item];
          EnvS<T, I, IS> e;
          e.it = d.itab + d.itab_off[dev];
          e.pc = d.pcode + d.par_off[dev];
          e.pval = d.pval;
          e.pinst = pinst;
          const IS so = (IS)d.state_off[dev] * ss + scol;
          e.sop = sop + so; e.sguess = sguess + so; e.sstride = ss;
          e.x = x + col; e.xstride = ws;
          e.S = S + (IS)ct.stage_off[dev] * wS + colS; e.Sstride = wS;
          e.mode = ctl.mode; e.dt = ctl.dt; e.gmin = ctl.gmin; e.omega = omega; e.time = tnow;
          load_one<T, B4>(d.type[dev], e, (d.par_direct && d.par_direct[dev]) ? d.pval + d.par_off[dev] : nullptr);
        }
      }
      if (tid < gi) { resok[tid] = 1; sing[tid] = 0; weak[tid] = 0; maxabs[tid] = 0.0; }
      __syncthreads();
      // ---- P2: assembly — gather staging slots per L+U slot / rhs row in the reference's accumulation order
      if (on) {
        for (int t = item0; t < nnz + N; t += istep) {
          T acc = Scalar<T>::zero();
          int q = ct.asm_off[t];
          const int qe = ct.asm_off[t + 1];
          // four loads in flight, the additions in list order (when the staging area is in HBM/L2 the chain of dependent
          // load -> add pairs was the whole cost of this phase)
          for (; q + 4 <= qe; q += 4) {
            const T v0 = S[(IS)ct.asm_src[q] * wS + colS], v1 = S[(IS)ct.asm_src[q + 1] * wS + colS];
            const T v2 = S[(IS)ct.asm_src[q + 2] * wS + colS], v3 = S[(IS)ct.asm_src[q + 3] * wS + colS];
            acc = s_add(s_add(s_add(s_add(acc, v0), v1), v2), v3);
          }
          for (; q < qe; q++) acc = s_add(acc, S[(IS)ct.asm_src[q] * wS + colS]);
          if (t < nnz) lu[(I)t * ws + col] = acc;
          else rhs[(I)(t - nnz) * ws + col] = acc;
        }
      }
      __syncthreads();
      // ---- P3: residual in pivoted row order (Matrix::res, sparse21/mod.rs:298-327)
      if (on) {
        for (int r = item0; r < N; r += istep) {
          T acc = Scalar<T>::zero();
          for (int s = p.rowptr[r]; s < p.rowptr[r + 1]; s++)
            acc = s_add(acc, s_mul(lu[(I)s * ws + col], x[(I)p.col_i2e[p.colidx[s]] * ws + col]));
          const T rv = s_sub(rhs[(I)p.row_i2e[r] * ws + col], acc);
          c[(I)r * ws + col] = rv;
          if (!TolC<T>::ok(s_abs(rv), itol)) resok[li] = 0;
        }
      }
      __syncthreads();
      // ---- convergence decision (Solver::converged, analysis.rs:331-345)
      if (tid < gi) {
        int cn = 0;
        if (act[tid]) {
          nld[tid] += 1;
          if (dxok[tid] && resok[tid]) { act[tid] = 0; cn = 1; }
        }
        convnow[tid] = cn;
        dxok[tid] = 1;
      }
      __syncthreads();
      if (real_kind && convnow[li]) {  // Component::commit on convergence: op <- guess
        for (int k = item0; k < d.n_state; k += istep) sop[(IS)k * ss + scol] = sguess[(IS)k * ss + scol];
      }
      if (!__syncthreads_or(tid < gi && act[tid])) break;
      const bool go = act[li] != 0;
      // ---- numeric LU on the frozen pattern, one barrier per dependency level (row_col_elim, mod.rs:865-919)
      for (int q = 0; q < ct.n_lu_lvl; q++) {
        if (go) {
          const int b = ct.lu_lvl_off[q], e_ = ct.lu_lvl_off[q + 1];
          for (int op = b + item0; op < e_; op += istep) {
            const int l = ct.lu_l[op];
            T* t = lu + (I)ct.lu_t[op] * ws + col;
            const T u = lu[(I)ct.lu_u[op] * ws + col];
            if (l < 0) {
              if (l == -1 && ctl.stop_on_weak && s_abs(u) * ctl.weak_mult < s_abs(*t)) weak[li] = 1;  // -2: a pivot the reference chose without threshold
              *t = s_div(*t, u);
            } else {
              *t = s_sub(*t, s_mul(u, lu[(I)l * ws + col]));
            }
          }
        }
        __syncthreads();
      }
      // ---- forward substitution (mod.rs:947-964); c is already in pivoted row order
      for (int q = 0; q < ct.n_fw_lvl; q++) {
        if (go) {
          const int b = ct.fw_lvl_off[q], e_ = ct.fw_lvl_off[q + 1];
          for (int op = b + item0; op < e_; op += istep) {
            const T ck = c[(I)ct.fw_k[op] * ws + col];
            if (s_is_zero(ck)) continue;
            T* t = c + (I)ct.fw_row[op] * ws + col;
            *t = s_sub(*t, s_mul(ck, lu[(I)ct.fw_slot[op] * ws + col]));
          }
        }
        __syncthreads();
      }
      // ---- backward substitution (mod.rs:967-979): each row's sum runs in list order inside one thread
      for (int q = 0; q < ct.n_bw_lvl; q++) {
        if (go) {
          const int b = ct.bw_lvl_off[q], e_ = ct.bw_lvl_off[q + 1];
          for (int r = b + item0; r < e_; r += istep) {
            const int k = ct.bw_row[r];
            const int ds = p.diag_slot[k];
            T ck = c[(I)k * ws + col];
            for (int s = ds + 1; s < p.rowptr[k + 1]; s++) ck = s_sub(ck, s_mul(c[(I)p.colidx[s] * ws + col], lu[(I)s * ws + col]));
            c[(I)k * ws + col] = s_div(ck, lu[(I)ds * ws + col]);
          }
        }
        __syncthreads();
      }
      // ---- zero-pivot check (assert(pivot_val).ne(0), mod.rs:871-872) and max |dx| (analysis.rs:198)
      if (go) {
        double m = 0.0;
        for (int k = item0; k < N; k += istep) {
          if (k + 1 < N && s_is_zero(lu[(I)p.diag_slot[k] * ws + col])) sing[li] = 1;
          const double v = s_abs(c[(I)p.col_e2i[k] * ws + col]);
          if (v > m) m = v;
          // a vanishing (not exactly zero) frozen pivot shows as a non-finite step: where the host can re-pivot, hand it back too
          if (KIND == K_TRAN && ctl.tran_stop && !(v <= 1.7976931348623157e308)) weak[li] = 1;
        }
        if (KIND == K_TRAN && ctl.tran_stop && !ctl.resume && ctl.tran_inject_tp == tp && iter == 1) weak[li] = 1;  // test hook
        if (m > 0.0) atomicMax((unsigned long long*)&maxabs[li], (unsigned long long)__double_as_longlong(m));
      }
      __syncthreads();
      // ---- global step limit and update (analysis.rs:197-207 / 283-293)
      if (go && !sing[li] && !weak[li]) {
        const double m = maxabs[li];
        for (int k = item0; k < N; k += istep) {
          T dxk = c[(I)p.col_e2i[k] * ws + col];
          if (m > 1.0 && !(KIND == K_AC && ctl.ac_direct)) dxk = s_scale(dxk, 1.0, m);
          T* xv = x + (I)k * ws + col;
          *xv = s_add(*xv, dxk);
          if (!TolC<T>::ok(s_abs(dxk), vtol)) dxok[li] = 0;
        }
      }
      __syncthreads();
      if (tid < gi && act[tid]) {
        if (KIND == K_TRAN && ctl.tran_stop && (sing[tid] || weak[tid])) {
          // inside the time loop: back to the last accepted point (below), the host takes a pivot order there and resumes
          act[tid] = 0; stat[tid] = CST_REPIVOT; stopat[tid] = tp; a.tp_stop[i0 + tid] = tp;
        }
        else if (sing[tid]) { act[tid] = 0; stat[tid] = CST_SINGULAR; }
        else if (weak[tid]) { act[tid] = 0; stat[tid] = CST_REPIVOT; }  // x untouched: the host re-pivots at this iterate and continues
        else {
          nsol[tid] += 1;
          if (KIND == K_AC && ctl.ac_direct) act[tid] = 0;  // linear system: x = A^-1 b is the answer (engine.hpp SolveCtl::ac_direct)
          if (nsol[tid] >= left[tid]) { act[tid] = 0; stat[tid] = CST_CONV; }  // resume: its own 100 iterations are used up
        }
      }
      __syncthreads();
    }
    if (tid < gi && act[tid]) { stat[tid] = CST_CONV; act[tid] = 0; }  // "Convergence Failed" (analysis.rs:209, 302)
    __syncthreads();
    if constexpr (KIND == K_TRAN) {
      if (real_kind && ctl.tran_stop && stopat[li] == tp && tp > 0) {  // stopped at this point: x and the in-flight state as the point began
        for (int k = item0; k < N; k += istep) x[(I)k * ws + col] = a.x_acc[(size_t)k * g.stride + i0 + li];
        for (int k = item0; k < d.n_state; k += istep) sguess[(IS)k * ss + scol] = sop[(IS)k * ss + scol];
      }
      if (valid && left[li] > 0 && tp >= start[li]) {  // an instance this launch leaves alone keeps its waveform
        const bool good = stat[li] == CST_OK;
        for (int s = item0; s < a.n_save; s += istep)
          a.wave[((size_t)tp * a.n_save + s) * g.stride + i0 + li] =
              good ? x[(I)a.save_vars[s] * ws + col] : __longlong_as_double(0x7ff8000000000000LL);
      }
    }
  }

  // ---- epilogue: results back to HBM
  if constexpr (WS != 0) {
    if (valid) {
      for (int k = item0; k < N; k += istep) g.x[(size_t)k * g.stride + i0 + li] = x[(I)k * ws + col];
      if constexpr (WS == 1) {
        if (real_kind)
          for (int k = item0; k < d.n_state; k += istep) {
            const size_t dst = (size_t)k * g.st_stride + i0 + li;
            g.st_op[dst] = sop[(IS)k * ss + scol];
            g.st_guess[dst] = sguess[(IS)k * ss + scol];
          }
      }
    }
  }
  if (tid < ni) {
    o.status[i0 + tid] = stat[tid];  // resume: an instance that was left alone has its old code in stat[] and zeros in the counters
    o.iters[i0 + tid] += nsol[tid];
    o.loads[i0 + tid] += nld[tid];
  }
}

size_t ctrl_bytes(int gi) { return ((8 + 12 * 4) * (size_t)gi + 8 + 8 + 15) / 16 * 16; }

template <class T, int KIND, bool B4>
int launch_b4(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<T>& w, T* stage, const NewtonOut& o,
           const SolveCtl& c, const CoopCfg& cfg, int T_points, const int* save_vars, int n_save, double* wave, void* stream) {
  CoopArgs a;
  a.lg_gi = 0; a.gi = cfg.gi; a.cold = 0; a.T_points = T_points; a.n_save = n_save; a.save_vars = save_vars; a.wave = wave;
  a.tp_stop = cfg.tp_stop; a.x_acc = cfg.x_acc;
  {
    const char* sb = std::getenv("S21_COOP_STAGE_BLOCKED");
    a.stage_blocked = (sb && std::atoi(sb) == 0) ? 0 : 1;
  }
  a.arena = cfg.arena;
  a.pcode_global = cfg.arena_in_smem && cfg.arena_core_bytes > 0 ? 1 : 0;
  a.arena_bytes = cfg.arena_in_smem ? (int)(a.pcode_global ? cfg.arena_core_bytes : cfg.arena_bytes) : 0;
  const size_t smem = ctrl_bytes(cfg.gi) + (size_t)a.arena_bytes + cfg.smem_bytes;
  const int grid = (c.B + cfg.gi - 1) / cfg.gi;
  cudaError_t e;
  if (cfg.smem_bytes > 0 && cfg.mixed) {
    auto kern = k_coop<T, KIND, 2, B4>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    if (const char* cv = std::getenv("S21_COOP_CARVEOUT"))  // measurement: shared-memory share of the L1 / shared array, in percent
      cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, std::atoi(cv));
    kern<<<grid, cfg.threads, smem, (cudaStream_t)stream>>>(d, p, ct, w, stage, o, c, a);
  } else if (cfg.smem_bytes > 0) {
    auto kern = k_coop<T, KIND, 1, B4>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, cfg.threads, smem, (cudaStream_t)stream>>>(d, p, ct, w, stage, o, c, a);
  } else {
    auto kern = k_coop<T, KIND, 0, B4>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, cfg.threads, smem, (cudaStream_t)stream>>>(d, p, ct, w, stage, o, c, a);
  }
  return (int)cudaGetLastError();
}

template <class T, int KIND>
int launch(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<T>& w, T* stage, const NewtonOut& o,
           const SolveCtl& c, const CoopCfg& cfg, int T_points, const int* save_vars, int n_save, double* wave, void* stream) {
  if constexpr (KIND != K_AC) {
    if (c.has_bsim4) return launch_b4<T, KIND, true>(d, p, ct, w, stage, o, c, cfg, T_points, save_vars, n_save, wave, stream);
  }
#ifdef S21_COOP_FAST
  return (int)cudaErrorInvalidValue;  // coop_fast.cu holds the Bsim4 kernels only
#else
  return launch_b4<T, KIND, false>(d, p, ct, w, stage, o, c, cfg, T_points, save_vars, n_save, wave, stream);
#endif
}

}  // namespace

#ifdef S21_COOP_FAST
// kernels/coop_fast.cu: the same Bsim4 kernels with the evaluation's divisions as a * rcp(b) (bsim4/bsim4_eval.hpp)
int launch_coop_dcop_fast(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<double>& w, double* stage,
                          const NewtonOut& o, const SolveCtl& c, const CoopCfg& cfg, void* stream) {
  return launch<double, K_DCOP>(d, p, ct, w, stage, o, c, cfg, 2, nullptr, 0, nullptr, stream);
}
int launch_coop_tran_fast(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<double>& w, double* stage,
                          const NewtonOut& o, const SolveCtl& c, const CoopCfg& cfg, int T, const int* save_vars, int n_save, double* wave,
                          void* stream) {
  return launch<double, K_TRAN>(d, p, ct, w, stage, o, c, cfg, T, save_vars, n_save, wave, stream);
}
#else
size_t coop_ctrl_bytes(int gi) { return ctrl_bytes(gi); }
size_t coop_mixed_bytes(int N, int nnz, int gi, int scalar_width) { return 8 * (size_t)scalar_width * (size_t)gi * (3 * (size_t)N + (size_t)nnz); }
size_t coop_work_bytes(int N, int nnz, int n_stage, int n_state, int gi, int scalar_width) {
  const size_t ts = 8 * (size_t)scalar_width;
  return ts * (size_t)gi * (3 * (size_t)N + (size_t)nnz + (size_t)n_stage) + 8 * (size_t)gi * 2 * (size_t)n_state;
}
int coop_max_smem_optin(int device) {
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) return 0;
  return v;
}

int launch_coop_dcop(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<double>& w, double* stage,
                     const NewtonOut& o, const SolveCtl& c, const CoopCfg& cfg, void* stream) {
  return launch<double, K_DCOP>(d, p, ct, w, stage, o, c, cfg, 2, nullptr, 0, nullptr, stream);
}
int launch_coop_tran(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<double>& w, double* stage,
                     const NewtonOut& o, const SolveCtl& c, const CoopCfg& cfg, int T, const int* save_vars, int n_save, double* wave,
                     void* stream) {
  return launch<double, K_TRAN>(d, p, ct, w, stage, o, c, cfg, T, save_vars, n_save, wave, stream);
}
int launch_coop_ac(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<cplx>& w, cplx* stage,
                   const NewtonOut& o, const SolveCtl& c, const CoopCfg& cfg, void* stream) {
  return launch<cplx, K_AC>(d, p, ct, w, stage, o, c, cfg, 2, nullptr, 0, nullptr, stream);
}
#endif  // S21_COOP_FAST

}  // namespace s21
