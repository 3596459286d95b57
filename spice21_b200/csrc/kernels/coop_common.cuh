// Shared pieces of the cooperative kernels (coop.cu, hybrid.cu): TMA bulk-copy helpers, the staged Env that device
// functions write their stamps through, the device dispatch, tolerances and the launch-argument block.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "devices.cuh"
#include "engine.hpp"

namespace s21 {
namespace coopk {

enum { K_DCOP = 0, K_TRAN = 1, K_AC = 2 };
enum { CST_OK = 0, CST_CONV = 1, CST_SINGULAR = 2, CST_REPIVOT = 9 };  // 9 = engine.hpp ST_REPIVOT_CODE

// ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier (sm_90+; SASS: UBLKCP + SYNCS)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  }
}

template <class T, class I, class IS = I>
struct EnvS {  // staged Env (see devices.cuh): stamps go to this device's private staging slots
  const int* it;
  const int* pc;
  const double* pval;
  size_t pinst;
  double* sop;      // already offset to (state_off, instance); element k at sop[k * sstride]
  double* sguess;
  IS sstride;
  const T* x;       // already offset to the instance column; variable v at x[v * xstride]
  T* S;             // already offset to (stage_off, instance); position p at S[p * Sstride]
  I xstride;
  IS Sstride;       // the staging area may live elsewhere than x (cooperative kernel, mixed workspace: x on chip, staging in HBM)
  int mode;
  double dt, gmin, omega;
  double time;      // transient: the time point being solved (time-varying sources)
  __device__ __forceinline__ int node(int k) const { return it[k]; }
  __device__ __forceinline__ double par(int k) const {
    const int c = pc[k];
    return __ldg(pval + (size_t)(c >> 1) + (size_t)(c & 1) * pinst);
  }
  __device__ __forceinline__ double volt(int var) const {
    if constexpr (std::is_same<T, double>::value) return var < 0 ? 0.0 : x[(I)var * xstride];
    else return 0.0;  // load_ac never reads the guess
  }
  __device__ __forceinline__ double op(int k) const { return sop[(IS)k * sstride]; }
  __device__ __forceinline__ double guess(int k) const { return sguess[(IS)k * sstride]; }
  __device__ __forceinline__ void set_guess(int k, double v) { sguess[(IS)k * sstride] = v; }
  __device__ __forceinline__ void add_g_at(int pos, T v) { S[(IS)pos * Sstride] = v; }
  __device__ __forceinline__ void add_b_at(int pos, T v) { S[(IS)pos * Sstride] = v; }
  __device__ __forceinline__ void add_g_dup(int, int dup, T v) { S[(IS)dup * Sstride] = v; }
};

// Env of a device whose parameter block is all shared values: par(k) is one load at a literal offset from the block.
template <class T, class I, class IS = I>
struct EnvSD : EnvS<T, I, IS> {
  const double* pblk;
  __device__ __forceinline__ double par(int k) const { return __ldg(pblk + k); }
};

template <class T, bool B4, class E> __device__ __forceinline__ void load_one(int type, E& e, const double* pblk = nullptr) {
  if constexpr (std::is_same<T, double>::value) {
    switch (type) {
      case DT_R: load_resistor(e); break;
      case DT_C: load_capacitor(e); break;
      case DT_I: load_isrc(e); break;
      case DT_V: load_vsrc(e); break;
      case DT_DIODE: load_diode(e); break;
      case DT_MOS0: load_mos0(e); break;
      case DT_MOS1: load_mos1(e); break;
      case DT_BSIM4:
        if constexpr (B4) {
          if (pblk) {
            EnvSD<T, decltype(e.xstride), decltype(e.Sstride)> ed;
            static_cast<E&>(ed) = e;
            ed.pblk = pblk;
            load_bsim4(ed);
          } else {
            load_bsim4(e);
          }
        }
        break;
      default: break;
    }
  } else {
    switch (type) {
      case DT_R: load_ac_resistor(e); break;
      case DT_C: load_ac_capacitor(e); break;
      case DT_V: load_ac_vsrc(e); break;
      case DT_MOS1: load_ac_mos1(e); break;
      default: break;  // the host refuses AC for devices without load_ac before launching
    }
  }
}

template <class T> struct TolC;
template <> struct TolC<double> {
  static __device__ __forceinline__ bool ok(double a, double tol) { return !(a > tol); }
  static const int max_iter = 100;
};
template <> struct TolC<cplx> {
  static __device__ __forceinline__ bool ok(double a, double tol) { return a < tol; }
  static const int max_iter = 20;
};

struct CoopArgs {
  int lg_gi;             // hybrid kernel (always 32 instances per CTA)
  int gi;                // cooperative kernel: instances per CTA, any value >= 1 (blockDim.x is a multiple of it)
  int cold;              // 1: a pending reset is folded into this launch: x and device state start at zero, counters restart
  int T_points, n_save;
  const int* save_vars;
  double* wave;
  const int* arena;      // packed index tables in HBM (all table pointers point into it)
  int arena_bytes;       // > 0: copy the arena into shared memory with TMA and rebase the pointers
  int pcode_global;      // cooperative kernel: the copy stops short of the parameter-code table (last in the arena), which stays in HBM
  // cooperative transient kernel with SolveCtl::tran_stop / resume: [B] time point at which an instance was handed back (and
  // resumes), [N][stride] x of every instance at its last accepted time point
  int* tp_stop = nullptr;
  double* x_acc = nullptr;
  int stage_blocked = 0;  // cooperative kernel, staging in HBM: one contiguous [slot][gi] block per CTA (needs n_stage x (grid x gi) entries)
};


}  // namespace coopk
}  // namespace s21
