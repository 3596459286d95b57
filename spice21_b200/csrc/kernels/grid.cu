// Grid-wide Newton kernel for ONE large circuit (config C3: N ~ 14 k, 34 k devices, 4 M L+U slots).
//
// With a single instance there is no batch axis to spread over the GPU: the cooperative kernel (coop.cu) would put the
// whole circuit on one CTA and leave 147 SMs idle. Here the same phases — parallel device evaluation into private
// staging slots, gather in the reference's accumulation order, residual, level-scheduled LU on the frozen pattern,
// level-scheduled substitutions, step limit, update — run over ALL threads of a cooperative launch (one persistent
// grid sized to the occupancy limit), with grid-wide barriers where the cooperative kernel has CTA barriers. The
// workspace lives in HBM/L2 with instance stride 1 (dense). Control words (convergence flags, max |dx|, counters) are a
// small block in global memory. Every floating-point operation on any single value happens in the same order as in the
// other kernels: results are bit-identical to them.
//
// Replaces the same reference functions as newton.cu (analysis.rs:153-210, 331-345, 553-570; sparse21/mod.rs:272-327,
// 865-991), for B = 1.
#include <cooperative_groups.h>

#include "coop_common.cuh"

namespace s21 {

using namespace coopk;
namespace cg = cooperative_groups;

namespace {

template <int KIND, bool B4>
__global__ void __launch_bounds__(256, B4 ? 1 : 2) k_grid(DevTables d, PlanTables p, CoopTables ct, WorkTables<double> g, double* S, NewtonOut o,
                                                            SolveCtl ctl, GridCtl* gc_, int T_points, int n_save, const int* save_vars, double* wave) {
  volatile GridCtl* gc = gc_;  // control words change between grid barriers: never keep them in registers
  cg::grid_group grid = cg::this_grid();
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nt = (size_t)gridDim.x * blockDim.x;
  const int N = p.N, nnz = p.nnz;
  double* x = g.x; double* rhs = g.rhs; double* c = g.c; double* lu = g.lu;
  double* sop = g.st_op; double* sguess = g.st_guess;
  const double vtol = ctl.reltol, itol = ctl.iabstol;

  // resume (SolveCtl::resume; the host only launches it for an instance it found stopped): what is left of the solve's budget
  const int used = (KIND == K_DCOP && ctl.resume) ? o.iters[0] - (o.iters_base ? o.iters_base[0] : 0) : 0;  // written at the very end only
  if (tid == 0) {
    gc->stat = KIND == K_TRAN ? o.status[0] : (used >= ctl.max_iter ? CST_CONV : 0);
    gc->nsol = 0; gc->nld = 0; gc->dxok = 1; gc->act = 0; gc->resok = 1; gc->sing = 0; gc->weak = 0; gc->maxabs = 0ull;
  }
  if constexpr (KIND == K_TRAN) {
    for (size_t s = tid; s < (size_t)n_save; s += nt) wave[s] = x[save_vars[s]];
  }
  grid.sync();
  const int n_points = KIND == K_TRAN ? T_points : 2;
  double tnow = KIND == K_TRAN ? ctl.dt : 0.0;  // analysis.rs:552-569: t starts at tstep and accumulates tstep
  for (int tp = 1; tp < n_points; tp++, tnow += ctl.dt) {
    if (tid == 0) { gc->act = gc->stat == CST_OK ? 1 : 0; gc->dxok = 1; }
    grid.sync();
    const int max_it = min(TolC<double>::max_iter, ctl.max_iter) - used;
    for (int iter = 0; iter < max_it; iter++) {
      if (!gc->act) break;  // uniform: written before the last barrier
      // ---- P1: device evaluation, devices in parallel
      for (size_t item = tid; item < (size_t)d.n_dev; item += nt) {
        const int dev = ct.eval_order[item];
        EnvS<double, size_t> e;
        e.it = d.itab + d.itab_off[dev];
        e.pc = d.pcode + d.par_off[dev];
        e.pval = d.pval;
        e.pinst = 0;
        const size_t so = (size_t)d.state_off[dev];
        e.sop = sop + so; e.sguess = sguess + so; e.sstride = 1;
        e.x = x; e.xstride = 1; e.Sstride = 1;
        e.S = S + (size_t)ct.stage_off[dev];
        e.mode = ctl.mode; e.dt = ctl.dt; e.gmin = ctl.gmin; e.omega = 0.0; e.time = tnow;
        load_one<double, B4>(d.type[dev], e, (d.par_direct && d.par_direct[dev]) ? d.pval + d.par_off[dev] : nullptr);
      }
      if (tid == 0) { gc->resok = 1; gc->sing = 0; gc->weak = 0; gc->maxabs = 0ull; }
      grid.sync();
      // ---- P2: assembly (fill-in slots have empty lists and come out as exact zeros)
      for (size_t t = tid; t < (size_t)(nnz + N); t += nt) {
        double acc = 0.0;
        for (int q = ct.asm_off[t]; q < ct.asm_off[t + 1]; q++) acc = s_add(acc, S[ct.asm_src[q]]);
        if (t < (size_t)nnz) lu[t] = acc;
        else rhs[t - (size_t)nnz] = acc;
      }
      grid.sync();
      // ---- P3: residual in pivoted row order
      {
        bool ok = true;
        for (size_t r = tid; r < (size_t)N; r += nt) {
          double acc = 0.0;
          for (int s = p.rowptr[r]; s < p.rowptr[r + 1]; s++) acc = s_add(acc, s_mul(lu[s], x[p.col_i2e[p.colidx[s]]]));
          const double rv = s_sub(rhs[p.row_i2e[r]], acc);
          c[r] = rv;
          if (!TolC<double>::ok(s_abs(rv), itol)) ok = false;
        }
        if (!ok) gc->resok = 0;
      }
      grid.sync();
      // ---- convergence decision
      if (tid == 0) {
        gc->nld += 1;
        gc->convnow = (gc->dxok && gc->resok) ? 1 : 0;
        if (gc->convnow) gc->act = 0;
        gc->dxok = 1;
      }
      grid.sync();
      if (gc->convnow) {
        for (size_t k = tid; k < (size_t)d.n_state; k += nt) sop[k] = sguess[k];  // Component::commit
        break;
      }
      // ---- numeric LU, one barrier per dependency level
      for (int q = 0; q < ct.n_lu_lvl; q++) {
        const size_t b = (size_t)ct.lu_lvl_off[q], e_ = (size_t)ct.lu_lvl_off[q + 1];
        for (size_t op = b + tid; op < e_; op += nt) {
          const int l = ct.lu_l[op];
          double* t = lu + ct.lu_t[op];
          const double u = lu[ct.lu_u[op]];
          if (l < 0) {
            if (l == -1 && ctl.stop_on_weak && s_abs(u) * ctl.weak_mult < s_abs(*t)) gc->weak = 1;  // pivot health (newton.cu)
            *t = s_div(*t, u);
          } else if (ctl.relaxed) atomicAdd(t, -s_mul(u, lu[l]));  // several updates of one level may share the target
          else *t = s_sub(*t, s_mul(u, lu[l]));
        }
        grid.sync();
      }
      // ---- forward substitution
      for (int q = 0; q < ct.n_fw_lvl; q++) {
        const size_t b = (size_t)ct.fw_lvl_off[q], e_ = (size_t)ct.fw_lvl_off[q + 1];
        for (size_t op = b + tid; op < e_; op += nt) {
          const double ck = c[ct.fw_k[op]];
          if (s_is_zero(ck)) continue;
          double* t = c + ct.fw_row[op];
          if (ctl.relaxed) atomicAdd(t, -s_mul(ck, lu[ct.fw_slot[op]]));
          else *t = s_sub(*t, s_mul(ck, lu[ct.fw_slot[op]]));
        }
        grid.sync();
      }
      // ---- backward substitution
      for (int q = 0; q < ct.n_bw_lvl; q++) {
        const size_t b = (size_t)ct.bw_lvl_off[q], e_ = (size_t)ct.bw_lvl_off[q + 1];
        for (size_t r = b + tid; r < e_; r += nt) {
          const int k = ct.bw_row[r];
          const int ds = p.diag_slot[k];
          double ck = c[k];
          for (int s = ds + 1; s < p.rowptr[k + 1]; s++) ck = s_sub(ck, s_mul(c[p.colidx[s]], lu[s]));
          c[k] = s_div(ck, lu[ds]);
        }
        grid.sync();
      }
      // ---- zero-pivot check and max |dx|
      {
        double m = 0.0;
        bool sg = false;
        for (size_t k = tid; k < (size_t)N; k += nt) {
          if (k + 1 < (size_t)N && s_is_zero(lu[p.diag_slot[k]])) sg = true;
          const double v = s_abs(c[p.col_e2i[k]]);
          if (v > m) m = v;
        }
        if (sg) gc->sing = 1;
        if (m > 0.0) atomicMax(&gc_->maxabs, (unsigned long long)__double_as_longlong(m));
      }
      grid.sync();
      // ---- global step limit and update
      if (!gc->sing && !gc->weak) {
        const double m = __longlong_as_double((long long)gc->maxabs);
        bool ok = true;
        for (size_t k = tid; k < (size_t)N; k += nt) {
          double dxk = c[p.col_e2i[k]];
          if (m > 1.0) dxk = s_scale(dxk, 1.0, m);
          x[k] = s_add(x[k], dxk);
          if (!TolC<double>::ok(s_abs(dxk), vtol)) ok = false;
        }
        if (!ok) gc->dxok = 0;
      }
      grid.sync();
      if (tid == 0) {
        if (gc->sing) { gc->act = 0; gc->stat = CST_SINGULAR; }
        else if (gc->weak) { gc->act = 0; gc->stat = CST_REPIVOT; }  // x untouched: the host re-pivots at this iterate and continues
        else {
          gc->nsol += 1;
          if (iter + 1 == max_it) { gc->act = 0; gc->stat = CST_CONV; }
        }
      }
      grid.sync();
    }
    grid.sync();
    if constexpr (KIND == K_TRAN) {
      const bool good = gc->stat == CST_OK;
      for (size_t s = tid; s < (size_t)n_save; s += nt)
        wave[(size_t)tp * n_save + s] = good ? x[save_vars[s]] : __longlong_as_double(0x7ff8000000000000LL);
    }
  }
  grid.sync();
  if (tid == 0) {
    o.status[0] = gc->stat;
    o.iters[0] += gc->nsol;
    o.loads[0] += gc->nld;
  }
}

template <int KIND, bool B4>
int launch_k(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<double>& w, double* stage, const NewtonOut& o,
             const SolveCtl& c, GridCtl* gc, int T, const int* save_vars, int n_save, double* wave, void* stream) {
  auto kern = k_grid<KIND, B4>;
  int dev = 0, sms = 0, per_sm = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0);
  if (e != cudaSuccess) return (int)e;
  if (per_sm < 1) return (int)cudaErrorLaunchOutOfResources;
  // barriers cost more with more CTAs and the widest phases (4 M gathers) are still short: one CTA per SM is enough
  dim3 grid((unsigned)sms), block(256);
  DevTables d_ = d; PlanTables p_ = p; CoopTables ct_ = ct; WorkTables<double> w_ = w; NewtonOut o_ = o; SolveCtl c_ = c;
  void* args[] = {&d_, &p_, &ct_, &w_, &stage, &o_, &c_, &gc, &T, &n_save, &save_vars, &wave};
  e = cudaLaunchCooperativeKernel((const void*)kern, grid, block, args, 0, (cudaStream_t)stream);
  return (int)e;
}

}  // namespace

int launch_grid_dcop(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<double>& w, double* stage, const NewtonOut& o,
                     const SolveCtl& c, GridCtl* gc, void* stream) {
  return c.has_bsim4 ? launch_k<K_DCOP, true>(d, p, ct, w, stage, o, c, gc, 2, nullptr, 0, nullptr, stream)
                     : launch_k<K_DCOP, false>(d, p, ct, w, stage, o, c, gc, 2, nullptr, 0, nullptr, stream);
}
int launch_grid_tran(const DevTables& d, const PlanTables& p, const CoopTables& ct, const WorkTables<double>& w, double* stage, const NewtonOut& o,
                     const SolveCtl& c, GridCtl* gc, int T, const int* save_vars, int n_save, double* wave, void* stream) {
  return c.has_bsim4 ? launch_k<K_TRAN, true>(d, p, ct, w, stage, o, c, gc, T, save_vars, n_save, wave, stream)
                     : launch_k<K_TRAN, false>(d, p, ct, w, stage, o, c, gc, T, save_vars, n_save, wave, stream);
}

}  // namespace s21
