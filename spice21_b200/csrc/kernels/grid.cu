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

#include <algorithm>
#include <cstdlib>

#include "coop_common.cuh"

namespace s21 {

using namespace coopk;
namespace cg = cooperative_groups;

namespace {

// ---- tolerance ("relaxed") mode helpers. A single large circuit has a few VERY long sums — on config C3 the supply
// node's diagonal collects ~40 k stamps (every PMOS of 2000 rings), and 2000 rows of the L+U pattern have ~2000 entries
// each — and one thread walking such a list while the other 37 887 wait at the next grid barrier was most of an
// iteration (18.8 ms per time point at full size). In tolerance mode the order of a sum is free (the level schedule
// already applies same-target updates atomically), so: a warp owns 32 consecutive items, short ones stay one-thread
// sums in list order, long ones are summed by the whole warp (strided, four accumulators per lane, shuffle tree), and the
// few huge gather lists are cut into chunks spread over the grid. The exact mode keeps the one-thread sums.
constexpr int kLongList = 48;     // beyond this a list is summed by its warp
constexpr int kHugeList = 4096;   // beyond this a gather list is cut into chunks for the whole grid
constexpr int kHugeChunk = 1024;

__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// per-phase wall clock of the launch, kept by thread 0 (it leaves every grid barrier with everybody else): GridCtl::phase_ns,
// printed by the host under S21_PLAN_INFO
#define GRID_LEVEL(w, q) do { if (tid == 0 && (q) < GridCtl::kLevels) { const unsigned long long t_ = now_ns(); gc_->level_ns[w][q] += t_ - t_lv; t_lv = t_; } } while (0)
#define GRID_PHASE(k) do { if (tid == 0) { const unsigned long long t_ = now_ns(); gc_->phase_ns[k] += t_ - t_ph; t_ph = t_; } } while (0)

__device__ __forceinline__ double warp_total(double a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  return a;
}
// sum of f(q) over q in [b0, b1), all 32 lanes of the warp calling with the same bounds
template <class F>
__device__ __forceinline__ double warp_sum(int b0, int b1, int lane, F f) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int q = b0 + lane;
  for (; q + 96 < b1; q += 128) {  // four chains of dependent loads in flight per lane (eight measured slower: backward 11.7 -> 18.2 ms on C3)
    const double v0 = f(q), v1 = f(q + 32), v2 = f(q + 64), v3 = f(q + 96);
    a0 += v0; a1 += v1; a2 += v2; a3 += v3;
  }
  for (; q < b1; q += 32) a0 += f(q);
  return warp_total((a0 + a1) + (a2 + a3));
}

template <int KIND, bool B4>
__global__ void __launch_bounds__(B4 ? 256 : 512, 1) k_grid(DevTables d, PlanTables p, CoopTables ct, WorkTables<double> g, double* S, NewtonOut o,
                                                            SolveCtl ctl, GridCtl* gc_, int T_points, int n_save, const int* save_vars, double* wave) {
  volatile GridCtl* gc = gc_;  // control words change between grid barriers: never keep them in registers
  cg::grid_group grid = cg::this_grid();
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nt = (size_t)gridDim.x * blockDim.x;
  const int N = p.N, nnz = p.nnz;
  double* x = g.x; double* rhs = g.rhs; double* c = g.c; double* lu = g.lu;
  double* sop = g.st_op; double* sguess = g.st_guess;
  const double vtol = ctl.reltol, itol = ctl.iabstol;
  const int lane = threadIdx.x & 31;
  const size_t gw = tid >> 5, nw = nt >> 5;  // blockDim is a multiple of 32: warps never straddle CTAs
  const bool relaxed = ctl.relaxed != 0;     // kernel argument: uniform over the grid

  // resume (SolveCtl::resume; the host only launches it for an instance it found stopped): what is left of the solve's budget
  const int used = (KIND == K_DCOP && ctl.resume) ? o.iters[0] - (o.iters_base ? o.iters_base[0] : 0) : 0;  // written at the very end only
  if (tid == 0) {
    gc->stat = KIND == K_TRAN ? o.status[0] : (used >= ctl.max_iter ? CST_CONV : 0);
    gc->nsol = 0; gc->nld = 0; gc->dxok = 1; gc->act = 0; gc->resok = 1; gc->sing = 0; gc->weak = 0; gc->maxabs = 0ull;
    for (int k = 0; k < GridCtl::kPhases; k++) gc_->phase_ns[k] = 0ull;
    for (int w = 0; w < 3; w++) for (int k = 0; k < GridCtl::kLevels; k++) gc_->level_ns[w][k] = 0ull;
  }
  unsigned long long t_ph = now_ns(), t_lv = t_ph;
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
      if (tid == 0) { gc->resok = 1; gc->sing = 0; gc->weak = 0; gc->maxabs = 0ull; gc->n_huge = 0; }
      grid.sync();
      GRID_PHASE(0);
      // ---- P2: assembly (fill-in slots have empty lists and come out as exact zeros)
      if (!relaxed) {
        for (size_t t = tid; t < (size_t)(nnz + N); t += nt) {
          double acc = 0.0;
          for (int q = ct.asm_off[t]; q < ct.asm_off[t + 1]; q++) acc = s_add(acc, S[ct.asm_src[q]]);
          if (t < (size_t)nnz) lu[t] = acc;
          else rhs[t - (size_t)nnz] = acc;
        }
        grid.sync();
      } else {
        const size_t total = (size_t)nnz + (size_t)N;
        for (size_t base = gw * 32; base < total; base += nw * 32) {  // warp-uniform trip count
          const size_t t = base + lane;
          int q0 = 0, qe = 0;
          if (t < total) { q0 = ct.asm_off[t]; qe = ct.asm_off[t + 1]; }
          double* dst = t < (size_t)nnz ? lu + t : rhs + (t - (size_t)nnz);
          int kind = (qe - q0) <= kLongList ? 0 : ((qe - q0) <= kHugeList ? 1 : 2);
          if (kind == 2) {  // left to the whole grid (below); a full work list falls back to the warp
            const int slot = atomicAdd(&gc_->n_huge, 1);
            if (slot < GridCtl::kHugeCap) { gc_->huge[slot] = (int)t; *dst = 0.0; }
            else kind = 1;
          }
          if (kind == 0 && t < total) {
            double acc = 0.0;
            for (int q = q0; q < qe; q++) acc = s_add(acc, S[ct.asm_src[q]]);
            *dst = acc;
          }
          unsigned m = __ballot_sync(0xffffffffu, kind == 1);
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const int b0 = __shfl_sync(0xffffffffu, q0, src), b1 = __shfl_sync(0xffffffffu, qe, src);
            const double acc = warp_sum(b0, b1, lane, [&](int q) { return S[ct.asm_src[q]]; });
            if (lane == src) *dst = acc;
          }
        }
        grid.sync();
        const int nh = min((int)gc->n_huge, (int)GridCtl::kHugeCap);  // uniform: written before the barrier
        if (nh > 0) {
          for (int h = 0; h < nh; h++) {
            const int t = gc->huge[h];
            const int q0 = ct.asm_off[t], qe = ct.asm_off[t + 1];
            double* dst = t < nnz ? lu + t : rhs + (t - nnz);
            const size_t nchunk = (size_t)((qe - q0) + kHugeChunk - 1) / kHugeChunk;
            for (size_t ch = gw; ch < nchunk; ch += nw) {
              const int b0 = q0 + (int)ch * kHugeChunk, b1 = min(qe, b0 + kHugeChunk);
              const double acc = warp_sum(b0, b1, lane, [&](int q) { return S[ct.asm_src[q]]; });
              if (lane == 0) atomicAdd(dst, acc);
            }
          }
          grid.sync();
        }
      }
      GRID_PHASE(1);
      // ---- P3: residual in pivoted row order
      if (!relaxed) {
        bool ok = true;
        for (size_t r = tid; r < (size_t)N; r += nt) {
          double acc = 0.0;
          for (int s = p.rowptr[r]; s < p.rowptr[r + 1]; s++) acc = s_add(acc, s_mul(lu[s], x[p.col_i2e[p.colidx[s]]]));
          const double rv = s_sub(rhs[p.row_i2e[r]], acc);
          c[r] = rv;
          if (!TolC<double>::ok(s_abs(rv), itol)) ok = false;
        }
        if (!ok) gc->resok = 0;
      } else {
        // fill-in slots (empty gather list) hold exact zeros at this point: the rows of A proper (CoopTables::res_off) skip them
        bool ok = true;
        for (size_t base = gw * 32; base < (size_t)N; base += nw * 32) {
          const size_t r = base + lane;
          int s0 = 0, se = 0;
          if (r < (size_t)N) { s0 = ct.res_off[r]; se = ct.res_off[r + 1]; }
          const bool lng = (se - s0) > kLongList;
          if (!lng && r < (size_t)N) {
            double acc = 0.0;
            for (int q = s0; q < se; q++) acc = s_add(acc, s_mul(lu[ct.res_slot[q]], x[ct.res_x[q]]));
            const double rv = s_sub(rhs[p.row_i2e[r]], acc);
            c[r] = rv;
            if (!TolC<double>::ok(s_abs(rv), itol)) ok = false;
          }
          unsigned m = __ballot_sync(0xffffffffu, lng);
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const int b0 = __shfl_sync(0xffffffffu, s0, src), b1 = __shfl_sync(0xffffffffu, se, src);
            const double acc = warp_sum(b0, b1, lane, [&](int q) { return s_mul(lu[ct.res_slot[q]], x[ct.res_x[q]]); });
            if (lane == src) {
              const double rv = s_sub(rhs[p.row_i2e[r]], acc);
              c[r] = rv;
              if (!TolC<double>::ok(s_abs(rv), itol)) ok = false;
            }
          }
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
        GRID_PHASE(2);
        break;
      }
      GRID_PHASE(2);
      // ---- numeric LU, one barrier per dependency level
      if (tid == 0) t_lv = now_ns();
      for (int q = 0; q < ct.n_lu_lvl; q++) {
        const size_t b = (size_t)ct.lu_lvl_off[q], e_ = (size_t)ct.lu_lvl_off[q + 1];
        if (!relaxed) {
          for (size_t op = b + tid; op < e_; op += nt) {
            const int l = ct.lu_l[op];
            double* t = lu + ct.lu_t[op];
            const double u = lu[ct.lu_u[op]];
            if (l < 0) {
              if (l == -1 && ctl.stop_on_weak && s_abs(u) * ctl.weak_mult < s_abs(*t)) gc->weak = 1;  // pivot health (newton.cu)
              *t = s_div(*t, u);
            } else *t = s_sub(*t, s_mul(u, lu[l]));
          }
        } else {
          // the operations of one level are independent apart from shared targets (atomic): four of them in flight per thread
          // — index loads, then operand loads, then the updates — instead of one dependent chain of L2 latencies per operation
          for (size_t op0 = b + (tid - lane); op0 < e_; op0 += 4 * nt) {  // warp-uniform trip count: the body shuffles
            const size_t op = op0 + lane;
            int li[4], ti[4], ui[4];
            double uv[4], lv[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const size_t o_ = op + (size_t)j * nt;
              const bool in = o_ < e_;
              li[j] = in ? ct.lu_l[o_] : -3;  // -3: nothing to do
              ti[j] = in ? ct.lu_t[o_] : 0;
              ui[j] = in ? ct.lu_u[o_] : 0;
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
              uv[j] = li[j] > -3 ? lu[ui[j]] : 0.0;
              lv[j] = li[j] >= 0 ? lu[li[j]] : (li[j] > -3 ? lu[ti[j]] : 0.0);  // a division reads its own target instead
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
              // Several updates of one level may share the target (atomic). The host sorted the level by target, so a shared
              // target is a run of consecutive lanes: the run is summed inside the warp (segmented shuffle reduction) and only
              // its first lane issues the atomic — the 2000 diagonal entries of C3's dense block receive ~1000 updates each
              // in one level, which as single atomics queued on one L2 address each (5.4 of the LU phase's 10.5 ms).
              const bool upd = li[j] >= 0;
              const int key = upd ? ti[j] : -1 - lane;  // never equal between lanes unless both are updates of one target
              double v = upd ? -s_mul(uv[j], lv[j]) : 0.0;
              const int knext = __shfl_down_sync(0xffffffffu, key, 1);
              if (__any_sync(0xffffffffu, lane < 31 && knext == key)) {
#pragma unroll
                for (int dlt = 1; dlt < 32; dlt <<= 1) {
                  const double ov = __shfl_down_sync(0xffffffffu, v, dlt);
                  const int ok_ = __shfl_down_sync(0xffffffffu, key, dlt);
                  if (lane + dlt < 32 && ok_ == key) v += ov;
                }
                const int kprev = __shfl_up_sync(0xffffffffu, key, 1);
                if (upd && (lane == 0 || kprev != key)) atomicAdd(lu + ti[j], v);
              } else if (upd) atomicAdd(lu + ti[j], v);
              if (!upd && li[j] > -3) {
                if (li[j] == -1 && ctl.stop_on_weak && s_abs(uv[j]) * ctl.weak_mult < s_abs(lv[j])) gc->weak = 1;
                lu[ti[j]] = s_div(lv[j], uv[j]);
              }
            }
          }
        }
        grid.sync();
        GRID_LEVEL(0, q);
      }
      GRID_PHASE(3);
      // ---- forward substitution
      for (int q = 0; q < ct.n_fw_lvl; q++) {
        const size_t b = (size_t)ct.fw_lvl_off[q], e_ = (size_t)ct.fw_lvl_off[q + 1];
        if (!relaxed) {
          for (size_t op = b + tid; op < e_; op += nt) {
            const double ck = c[ct.fw_k[op]];
            if (s_is_zero(ck)) continue;
            double* t = c + ct.fw_row[op];
            *t = s_sub(*t, s_mul(ck, lu[ct.fw_slot[op]]));
          }
        } else {
          for (size_t op = b + tid; op < e_; op += 4 * nt) {
            int ki[4], ri[4], si[4];
            double cv[4], lv[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const size_t o_ = op + (size_t)j * nt;
              const bool in = o_ < e_;
              ki[j] = in ? ct.fw_k[o_] : -1;
              ri[j] = in ? ct.fw_row[o_] : 0;
              si[j] = in ? ct.fw_slot[o_] : 0;
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
              cv[j] = ki[j] >= 0 ? c[ki[j]] : 0.0;
              lv[j] = ki[j] >= 0 ? lu[si[j]] : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 4; j++)
              if (ki[j] >= 0 && !s_is_zero(cv[j])) atomicAdd(c + ri[j], -s_mul(cv[j], lv[j]));
          }
        }
        grid.sync();
        GRID_LEVEL(1, q);
      }
      GRID_PHASE(4);
      // ---- backward substitution
      for (int q = 0; q < ct.n_bw_lvl; q++) {
        const size_t b = (size_t)ct.bw_lvl_off[q], e_ = (size_t)ct.bw_lvl_off[q + 1];
        if (!relaxed) {
          for (size_t r = b + tid; r < e_; r += nt) {
            const int k = ct.bw_row[r];
            const int ds = p.diag_slot[k];
            double ck = c[k];
            for (int s = ds + 1; s < p.rowptr[k + 1]; s++) ck = s_sub(ck, s_mul(c[p.colidx[s]], lu[s]));
            c[k] = s_div(ck, lu[ds]);
          }
        } else {
          // Long rows come in blocks here (C3: one level holds the 2000 rows of a dense triangle, 2 M entries), so a warp
          // owning 32 consecutive rows would walk 32 long rows while 2300 warps idle (measured: 13.5 of 14.8 ms). Two passes
          // over the level instead, on disjoint rows: a thread per short row, then a warp per long row, warp-strided.
          for (size_t r = b + tid; r < e_; r += nt) {
            const int k = ct.bw_row[r];
            const int ds = p.diag_slot[k], se = p.rowptr[k + 1];
            if (se - (ds + 1) > kLongList) continue;
            double ck = c[k];
            for (int s = ds + 1; s < se; s++) ck = s_sub(ck, s_mul(c[p.colidx[s]], lu[s]));
            c[k] = s_div(ck, lu[ds]);
          }
          for (size_t r = b + gw; r < e_; r += nw) {  // warp-uniform row
            const int k = ct.bw_row[r];
            const int ds = p.diag_slot[k], se = p.rowptr[k + 1];
            if (se - (ds + 1) <= kLongList) continue;
            const double acc = warp_sum(ds + 1, se, lane, [&](int s) { return s_mul(c[p.colidx[s]], lu[s]); });
            if (lane == 0) c[k] = s_div(s_sub(c[k], acc), lu[ds]);
          }
        }
        grid.sync();
        GRID_LEVEL(2, q);
      }
      GRID_PHASE(5);
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
      GRID_PHASE(6);
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
  // One CTA per SM: a grid barrier costs with the number of CTAs, not of threads. Measured on C3 with two 256-thread CTAs per
  // SM (profiles/r02v_c3_phases.txt): the latency-bound phases (assembly, LU, forward) gain from the second set of warps,
  // the barrier-bound ones lose as much — hence ONE CTA of 512 threads for circuits without Bsim4 devices (256 with: the
  // evaluation needs the registers). S21_GRID_THREADS / S21_GRID_CTAS override both for measurements.
  int threads = B4 ? 256 : 512;
  if (const char* v = std::getenv("S21_GRID_THREADS")) threads = std::max(64, std::min(B4 ? 256 : 512, std::atoi(v) / 32 * 32));
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0);
  if (e != cudaSuccess) return (int)e;
  if (per_sm < 1) return (int)cudaErrorLaunchOutOfResources;
  int per = 1;
  if (const char* v = std::getenv("S21_GRID_CTAS")) per = std::max(1, std::min(per_sm, std::atoi(v)));
  dim3 grid((unsigned)(sms * per)), block((unsigned)threads);
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
