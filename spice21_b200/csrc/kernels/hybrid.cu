// Hybrid Newton kernel for small circuits (the benchmark regime: thousands of instances of an N ~ 10 circuit).
//
// A CTA of 8 warps owns 32 consecutive instances whose whole workspace lives in shared memory (as in coop.cu), but the
// two halves of a Newton iteration use different thread mappings:
//   * device evaluation: lane = instance, warp = device  -> every lane of a warp runs the same device model on 32
//     instances (no divergence, full lane utilisation for the expensive Mos1 evaluation);
//   * everything else (assembly gather, residual, convergence test, level-scheduled LU, forward/back substitution,
//     step limit, update): each WARP owns 4 instances and spends 8 lanes on each, so the dependent chain of tiny
//     phases is ordered by __syncwarp() only. Convergence and step-limit reductions are warp ballots / shuffles.
// Only two block-wide barriers remain per Newton iteration (eval -> rest, rest -> next eval), against ~35 in coop.cu
// where every dependency level of the LU is a CTA barrier (profiles/r01c_*: 4.0 barrier-stalled warps per issue).
//
// Shared-memory columns are padded to 36 doubles per slot: with (4 instances x 8 slots) per warp access the 32 lanes
// fall into 2 wavefronts, the minimum for 32 x 8 bytes; the lane = instance accesses of the eval phase are contiguous.
// Arithmetic order per value is unchanged, so results are bit-identical to newton.cu / coop.cu.
#include "coop_common.cuh"

namespace s21 {

using namespace coopk;

#ifdef S21_PHASE_PROFILE
// Debug build only (make PROFILE=1): cycles spent by warp 0 / warp 1 of every CTA in each phase, summed over CTAs.
__device__ unsigned long long s21_phase_cycles[16];
#define PH_T(var) const long long var = clock64()
#define PH_ADD(slot, t0, t1) if (lane == 0 && warp < 2) atomicAdd(&s21_phase_cycles[(slot) + 8 * warp], (unsigned long long)((t1) - (t0)))
#else
#define PH_T(var)
#define PH_ADD(slot, t0, t1)
#endif

namespace {

constexpr int HY_GI = 32;    // instances per CTA
constexpr int HY_P = 36;     // padded column stride (elements)
constexpr int HY_WARPS = 8;
constexpr int HY_LPI = 8;    // lanes per instance in the per-warp phases
constexpr unsigned FULL = 0xffffffffu;

template <class T, int KIND, bool B4>
__global__ void __launch_bounds__(256, 2) k_hyb(DevTables d, PlanTables p, CoopTables ct, WorkTables<T> g, NewtonOut o, SolveCtl ctl,
                                               CoopArgs a) {
  typedef unsigned I;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int i0 = blockIdx.x * HY_GI;
  const int ni = min(HY_GI, ctl.B - i0);
  const int N = p.N, nnz = p.nnz;
  constexpr bool real_kind = KIND != K_AC;
  // eval mapping: lane = instance, warp = first device
  const int ei = lane;
  // per-warp mapping: 4 instances x 8 workers, instance fastest
  const int ri = warp * 4 + (lane & 3);
  const int q = lane >> 2;
  const unsigned imask = 0x11111111u << (lane & 3);  // the 8 lanes that share this thread's instance

  // ---- shared memory: [act flags][mbarrier][arena copy][workspace]
  int* act_s = (int*)smem_raw;
  uint64_t* mbar = (uint64_t*)(act_s + HY_GI);
  size_t off = ((size_t)((unsigned char*)(mbar + 1) - smem_raw) + 15) / 16 * 16;
  {
    int* sa = (int*)(smem_raw + off);
    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(mbar, (uint32_t)a.arena_bytes);
      tma_load_1d(sa, a.arena, (uint32_t)a.arena_bytes, mbar);
    }
    off += (size_t)a.arena_bytes;
#define RB(ptr) ptr = sa + ((ptr) - a.arena)  // derived from the shared base on purpose (see coop.cu)
    RB(d.type); RB(d.itab_off); RB(d.par_off); RB(d.state_off); RB(d.itab); RB(d.pcode); if (d.par_direct) RB(d.par_direct);
    RB(p.row_i2e); RB(p.col_i2e); RB(p.col_e2i); RB(p.rowptr); RB(p.colidx); RB(p.diag_slot);
    RB(ct.stage_off); RB(ct.eval_order); RB(ct.asm_off); RB(ct.asm_src);
    RB(ct.lu_lvl_off); RB(ct.lu_t); RB(ct.lu_u); RB(ct.lu_l);
    RB(ct.fw_lvl_off); RB(ct.fw_k); RB(ct.fw_row); RB(ct.fw_slot); RB(ct.bw_lvl_off); RB(ct.bw_row);
#undef RB
  }
  T* x = (T*)(smem_raw + off); off += sizeof(T) * (size_t)N * HY_P;
  T* rhs = (T*)(smem_raw + off); off += sizeof(T) * (size_t)N * HY_P;
  T* c = (T*)(smem_raw + off); off += sizeof(T) * (size_t)N * HY_P;
  T* lu = (T*)(smem_raw + off); off += sizeof(T) * (size_t)nnz * HY_P;
  T* S = (T*)(smem_raw + off); off += sizeof(T) * (size_t)ct.n_stage * HY_P;
  double* sop = (double*)(smem_raw + off); off += sizeof(double) * (size_t)d.n_state * HY_P;
  double* sguess = (double*)(smem_raw + off);

  // ---- prologue: x and device state into shared memory (lane = instance: coalesced)
  const bool evalid = ei < ni;
  const bool cold = a.cold != 0;  // Batch::reset() folded into this launch: a fresh Solver starts from zeros
  for (int k = warp; k < N; k += HY_WARPS)
    x[(I)k * HY_P + ei] = (evalid && !cold) ? g.x[(size_t)k * g.stride + i0 + ei] : Scalar<T>::zero();
  for (int k = warp; k < d.n_state; k += HY_WARPS) {
    const size_t src = (size_t)k * g.st_stride + ((size_t)i0 + (size_t)(evalid ? ei : 0)) * ctl.par_inst_stride;
    sop[(I)k * HY_P + ei] = cold ? 0.0 : g.st_op[src];
    sguess[(I)k * HY_P + ei] = cold ? 0.0 : g.st_guess[src];
  }
  // per-instance control state lives in registers, replicated over the 8 lanes of the instance
  const bool rvalid = ri < ni;
  int r_stat = (rvalid && KIND == K_TRAN) ? o.status[i0 + ri] : 0;
  int r_left = 1 << 30;      // resume: iterations this instance's solve has left
  bool r_skip = false;       // resume: not a stopped instance — nothing of it is touched
  if (KIND == K_DCOP && ctl.resume) {
    const int s0 = rvalid ? o.status[i0 + ri] : 0;
    r_skip = !rvalid || s0 != CST_REPIVOT;
    r_stat = r_skip ? s0 : CST_OK;
    if (!r_skip) {
      r_left = ctl.max_iter - (o.iters[i0 + ri] - (o.iters_base ? o.iters_base[i0 + ri] : 0));
      if (r_left <= 0) { r_stat = CST_CONV; r_skip = true; }
    }
  }
  bool r_weak = false;  // pivot health (see newton.cu): set by the division steps of the current factorisation
  int r_nsol = 0, r_nld = 0;
  mbar_wait(mbar, 0);
  __syncthreads();
  if constexpr (KIND == K_TRAN) {
    if (rvalid)
      for (int s = q; s < a.n_save; s += HY_LPI) a.wave[(size_t)s * g.stride + i0 + ri] = x[(I)a.save_vars[s] * HY_P + ri];
  }
  const double vtol = real_kind ? ctl.reltol : 1e-3, itol = real_kind ? ctl.iabstol : 1e-9;  // analysis.rs:271-272, 331-345
  const int n_points = KIND == K_TRAN ? a.T_points : 2;
  const size_t pinst = ((size_t)i0 + (size_t)ei) * ctl.par_inst_stride;
  double omega = 0.0;
  if constexpr (KIND == K_AC) omega = evalid ? ctl.omega[i0 + ei] : 0.0;

  double tnow = KIND == K_TRAN ? ctl.dt : 0.0;  // analysis.rs:552-569: t starts at tstep and accumulates tstep
  for (int tp = 1; tp < n_points; tp++, tnow += ctl.dt) {
    bool r_act = rvalid && r_stat == CST_OK && !r_skip;
    bool r_dxok = true;
    if (q == 0) act_s[ri] = r_act ? 1 : 0;
    __syncthreads();
    const int max_it = real_kind ? min(TolC<T>::max_iter, ctl.max_iter) : TolC<T>::max_iter;
    for (int iter = 0; iter < max_it; iter++) {
      // ================= device evaluation: lane = instance, warp = device (Solver::update, analysis.rs:153-168)
      PH_T(t0);
      if (act_s[ei]) {
        for (int item = warp; item < d.n_dev; item += HY_WARPS) {
          const int dev = ct.eval_order[item];
          EnvS<T, I> e;
          e.it = d.itab + d.itab_off[dev];
          e.pc = d.pcode + d.par_off[dev];
          e.pval = d.pval;
          e.pinst = pinst;
          const I so = (I)d.state_off[dev] * HY_P + ei;
          e.sop = sop + so; e.sguess = sguess + so; e.sstride = HY_P;
          e.x = x + ei; e.xstride = HY_P; e.Sstride = HY_P;
          e.S = S + (I)ct.stage_off[dev] * HY_P + ei;
          e.mode = ctl.mode; e.dt = ctl.dt; e.gmin = ctl.gmin; e.omega = omega; e.time = tnow;
          load_one<T, B4>(d.type[dev], e, (d.par_direct && d.par_direct[dev]) ? d.pval + d.par_off[dev] : nullptr);
        }
      }
      PH_T(t1);
      __syncthreads();
      PH_T(t2);
      PH_ADD(0, t0, t1);
      PH_ADD(1, t1, t2);
      // ================= the rest of the iteration: this warp's 4 instances, 8 lanes each, warp-synchronous
      // ---- assembly: gather staging slots in the reference's accumulation order
      if (r_act) {
        for (int t = q; t < nnz + N; t += HY_LPI) {
          T acc = Scalar<T>::zero();
          for (int j = ct.asm_off[t]; j < ct.asm_off[t + 1]; j++) acc = s_add(acc, S[(I)ct.asm_src[j] * HY_P + ri]);
          if (t < nnz) lu[(I)t * HY_P + ri] = acc;
          else rhs[(I)(t - nnz) * HY_P + ri] = acc;
        }
      }
      __syncwarp();
      PH_T(t3);
      PH_ADD(2, t2, t3);
      // ---- residual in pivoted row order (Matrix::res, sparse21/mod.rs:298-327) + KCL test
      bool bad = false;
      if (r_act) {
        for (int r = q; r < N; r += HY_LPI) {
          T acc = Scalar<T>::zero();
          for (int s = p.rowptr[r]; s < p.rowptr[r + 1]; s++)
            acc = s_add(acc, s_mul(lu[(I)s * HY_P + ri], x[(I)p.col_i2e[p.colidx[s]] * HY_P + ri]));
          const T rv = s_sub(rhs[(I)p.row_i2e[r] * HY_P + ri], acc);
          c[(I)r * HY_P + ri] = rv;
          bad = bad || !TolC<T>::ok(s_abs(rv), itol);
        }
      }
      const bool resok = (__ballot_sync(FULL, bad) & imask) == 0;
      // ---- convergence decision (Solver::converged, analysis.rs:331-345) and commit (op <- guess)
      if (r_act) {
        r_nld += 1;
        if (r_dxok && resok) {
          if (real_kind)
            for (int k = q; k < d.n_state; k += HY_LPI) sop[(I)k * HY_P + ri] = sguess[(I)k * HY_P + ri];
          r_act = false;
        }
      }
      PH_T(t4);
      PH_ADD(3, t3, t4);
      // ---- numeric LU on the frozen pattern (row_col_elim, mod.rs:865-919): one __syncwarp per dependency level
      for (int lv = 0; lv < ct.n_lu_lvl; lv++) {
        if (r_act) {
          const int e_ = ct.lu_lvl_off[lv + 1];
          for (int op = ct.lu_lvl_off[lv] + q; op < e_; op += HY_LPI) {
            const int l = ct.lu_l[op];
            T* t = lu + (I)ct.lu_t[op] * HY_P + ri;
            const T u = lu[(I)ct.lu_u[op] * HY_P + ri];
            if (l < 0) { r_weak = r_weak || (l == -1 && ctl.stop_on_weak && s_abs(u) * ctl.weak_mult < s_abs(*t)); *t = s_div(*t, u); }
            else *t = s_sub(*t, s_mul(u, lu[(I)l * HY_P + ri]));
          }
        }
        __syncwarp();
      }
      PH_T(t5);
      PH_ADD(4, t4, t5);
      // ---- forward substitution (mod.rs:947-964)
      for (int lv = 0; lv < ct.n_fw_lvl; lv++) {
        if (r_act) {
          const int e_ = ct.fw_lvl_off[lv + 1];
          for (int op = ct.fw_lvl_off[lv] + q; op < e_; op += HY_LPI) {
            const T ck = c[(I)ct.fw_k[op] * HY_P + ri];
            if (s_is_zero(ck)) continue;
            T* t = c + (I)ct.fw_row[op] * HY_P + ri;
            *t = s_sub(*t, s_mul(ck, lu[(I)ct.fw_slot[op] * HY_P + ri]));
          }
        }
        __syncwarp();
      }
      // ---- backward substitution (mod.rs:967-979)
      for (int lv = 0; lv < ct.n_bw_lvl; lv++) {
        if (r_act) {
          const int e_ = ct.bw_lvl_off[lv + 1];
          for (int r = ct.bw_lvl_off[lv] + q; r < e_; r += HY_LPI) {
            const int k = ct.bw_row[r];
            const int ds = p.diag_slot[k];
            T ck = c[(I)k * HY_P + ri];
            for (int s = ds + 1; s < p.rowptr[k + 1]; s++) ck = s_sub(ck, s_mul(c[(I)p.colidx[s] * HY_P + ri], lu[(I)s * HY_P + ri]));
            c[(I)k * HY_P + ri] = s_div(ck, lu[(I)ds * HY_P + ri]);
          }
        }
        __syncwarp();
      }
      PH_T(t6);
      PH_ADD(5, t5, t6);
      // ---- zero-pivot check (mod.rs:871-872), max |dx| (analysis.rs:198): shuffles over the instance's 8 lanes
      bool zp = false;
      double m = 0.0;
      if (r_act) {
        for (int k = q; k < N; k += HY_LPI) {
          if (k + 1 < N && s_is_zero(lu[(I)p.diag_slot[k] * HY_P + ri])) zp = true;
          const double v = s_abs(c[(I)p.col_e2i[k] * HY_P + ri]);
          if (v > m) m = v;
        }
      }
      const bool sing = (__ballot_sync(FULL, zp) & imask) != 0;
      const bool wk = (__ballot_sync(FULL, r_weak) & imask) != 0;  // some lane of the instance divided by a weak pivot
      r_weak = false;
      m = fmax(m, __shfl_xor_sync(FULL, m, 4));
      m = fmax(m, __shfl_xor_sync(FULL, m, 8));
      m = fmax(m, __shfl_xor_sync(FULL, m, 16));
      // ---- global step limit and update (analysis.rs:197-207 / 283-293)
      bool baddx = false;
      if (r_act && !sing && !wk) {
        for (int k = q; k < N; k += HY_LPI) {
          T dxk = c[(I)p.col_e2i[k] * HY_P + ri];
          if (m > 1.0 && !(KIND == K_AC && ctl.ac_direct)) dxk = s_scale(dxk, 1.0, m);
          T* xv = x + (I)k * HY_P + ri;
          *xv = s_add(*xv, dxk);
          baddx = baddx || !TolC<T>::ok(s_abs(dxk), vtol);
        }
      }
      r_dxok = (__ballot_sync(FULL, baddx) & imask) == 0;
      if (r_act) {
        if (sing) { r_act = false; r_stat = CST_SINGULAR; }
        else if (wk) { r_act = false; r_stat = CST_REPIVOT; }  // x untouched: the host re-pivots at this iterate and continues
        else {
          r_nsol += 1;
          if (KIND == K_AC && ctl.ac_direct) r_act = false;  // linear system: one solve is the answer
          if (r_nsol >= r_left) { r_act = false; r_stat = CST_CONV; }  // resume: this instance's own 100 iterations are used up
        }
      }
      if (q == 0) act_s[ri] = r_act ? 1 : 0;
      PH_T(t7);
      PH_ADD(6, t6, t7);
      const int any_ = __syncthreads_or(r_act);
      PH_T(t8);
      PH_ADD(7, t7, t8);
      if (!any_) break;
    }
    if (r_act) { r_stat = CST_CONV; r_act = false; }  // "Convergence Failed" (analysis.rs:209, 302)
    if constexpr (KIND == K_TRAN) {
      __syncwarp();
      if (rvalid) {
        const bool good = r_stat == CST_OK;
        for (int s = q; s < a.n_save; s += HY_LPI)
          a.wave[((size_t)tp * a.n_save + s) * g.stride + i0 + ri] =
              good ? x[(I)a.save_vars[s] * HY_P + ri] : __longlong_as_double(0x7ff8000000000000LL);
      }
    }
  }
  __syncthreads();
  // ---- epilogue: results back to HBM (lane = instance: coalesced)
  if (evalid) {
    for (int k = warp; k < N; k += HY_WARPS) g.x[(size_t)k * g.stride + i0 + ei] = x[(I)k * HY_P + ei];
    if (real_kind)
      for (int k = warp; k < d.n_state; k += HY_WARPS) {
        const size_t dst = (size_t)k * g.st_stride + i0 + ei;
        g.st_op[dst] = sop[(I)k * HY_P + ei];
        g.st_guess[dst] = sguess[(I)k * HY_P + ei];
      }
  }
  if (rvalid && q == 0) {
    o.status[i0 + ri] = r_stat;
    o.iters[i0 + ri] = (cold ? 0 : o.iters[i0 + ri]) + r_nsol;
    o.loads[i0 + ri] = (cold ? 0 : o.loads[i0 + ri]) + r_nld;
  }
}

size_t hyb_ctrl_bytes() { return ((size_t)HY_GI * 4 + 8 + 15) / 16 * 16; }

template <class T, int KIND, bool B4>
int launch_b4(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<T>& w, const NewtonOut& o, const SolveCtl& c,
           const CoopCfg& cfg, int T_points, const int* save_vars, int n_save, double* wave, void* stream) {
  CoopArgs a;
  a.lg_gi = 5; a.gi = HY_GI; a.cold = cfg.cold ? 1 : 0; a.T_points = T_points; a.n_save = n_save; a.save_vars = save_vars; a.wave = wave;
  a.arena = cfg.arena; a.arena_bytes = (int)cfg.arena_bytes; a.pcode_global = 0;
  const size_t smem = hyb_ctrl_bytes() + cfg.arena_bytes + cfg.smem_bytes;
  const int grid = (c.B + HY_GI - 1) / HY_GI;
  auto kern = k_hyb<T, KIND, B4>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<grid, HY_WARPS * 32, smem, (cudaStream_t)stream>>>(d, p, ct, w, o, c, a);
  return (int)cudaGetLastError();
}

template <class T, int KIND>
int launch(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<T>& w, const NewtonOut& o, const SolveCtl& c,
           const CoopCfg& cfg, int T_points, const int* save_vars, int n_save, double* wave, void* stream) {
  if constexpr (KIND != K_AC) {
    if (c.has_bsim4) return launch_b4<T, KIND, true>(d, p, ct, w, o, c, cfg, T_points, save_vars, n_save, wave, stream);
  }
  return launch_b4<T, KIND, false>(d, p, ct, w, o, c, cfg, T_points, save_vars, n_save, wave, stream);
}

}  // namespace

#ifdef S21_PHASE_PROFILE
extern "C" int s21_debug_phase_cycles(unsigned long long* out16, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out16, s21_phase_cycles, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(s21_phase_cycles, z, sizeof(z)); }
  return 0;
}
#endif

size_t hybrid_smem_bytes(int N, int nnz, int n_stage, int n_state, size_t arena_bytes, int scalar_width) {
  const size_t ts = 8 * (size_t)scalar_width;
  return hyb_ctrl_bytes() + arena_bytes + ts * HY_P * (3 * (size_t)N + (size_t)nnz + (size_t)n_stage) + 8 * (size_t)HY_P * 2 * (size_t)n_state;
}
size_t hybrid_work_bytes(int N, int nnz, int n_stage, int n_state, int scalar_width) {
  return hybrid_smem_bytes(N, nnz, n_stage, n_state, 0, scalar_width) - hyb_ctrl_bytes();
}

int launch_hybrid_dcop(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<double>& w, const NewtonOut& o,
                       const SolveCtl& c, const CoopCfg& cfg, void* stream) {
  return launch<double, K_DCOP>(d, p, ct, w, o, c, cfg, 2, nullptr, 0, nullptr, stream);
}
int launch_hybrid_tran(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<double>& w, const NewtonOut& o,
                       const SolveCtl& c, const CoopCfg& cfg, int T, const int* save_vars, int n_save, double* wave, void* stream) {
  return launch<double, K_TRAN>(d, p, ct, w, o, c, cfg, T, save_vars, n_save, wave, stream);
}
int launch_hybrid_ac(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<cplx>& w, const NewtonOut& o,
                     const SolveCtl& c, const CoopCfg& cfg, void* stream) {
  return launch<cplx, K_AC>(d, p, ct, w, o, c, cfg, 2, nullptr, 0, nullptr, stream);
}

}  // namespace s21
