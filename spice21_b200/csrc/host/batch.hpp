// Batch runtime: owns the device tables, the per-instance workspace in HBM, the LU plans and the CUDA stream of one
// batch of circuit instances, and drives the kernels of kernels/newton.cu. Host work here is setup only (tables,
// the once-per-mode symbolic phase, per-instance parameter derivation); the Newton loop itself never returns to the
// host. There is no CPU path: every solve needs a CUDA device.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>

#include "../kernels/engine.hpp"
#include "flatten.hpp"
#include "jit.hpp"
#include "jit_team.hpp"
#include "staging.hpp"
#include "symbolic.hpp"

namespace s21 {

inline void cuda_check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw S21Error(ST_CUDA, std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
}
#define S21_CUDA(x) ::s21::cuda_check((x), #x)

template <class T>
struct DBuf {  // device buffer
  T* p = nullptr;
  size_t n = 0;
  DBuf() {}
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  ~DBuf() { release(); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  void alloc(size_t count) {
    if (count <= n && p) return;
    release();
    S21_CUDA(cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(T)));
    n = count;
  }
  void upload(const std::vector<T>& h, cudaStream_t s) {
    alloc(h.size());
    if (!h.empty()) S21_CUDA(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  }
};
template <class T>
struct PinnedBuf {  // page-locked host buffer
  T* p = nullptr;
  size_t n = 0;
  PinnedBuf() {}
  PinnedBuf(const PinnedBuf&) = delete;
  PinnedBuf& operator=(const PinnedBuf&) = delete;
  ~PinnedBuf() { if (p) cudaFreeHost(p); }
  void alloc(size_t count) {
    if (count <= n && p) return;
    if (p) cudaFreeHost(p);
    p = nullptr;
    S21_CUDA(cudaMallocHost((void**)&p, std::max<size_t>(count, 1) * sizeof(T)));
    n = count;
  }
};

// Device index table with element handles rewritten from raw element ids to L+U slots of plan `P`.
inline std::vector<int> itab_with_slots(const FlatCkt& flat, const Plan& P) {
  std::vector<int> itab = flat.itab;
  for (const FlatDev& d : flat.devs) {
    int first_elem = 0;
    switch (d.type) {
      case DT_R: case DT_C: first_elem = R_EPP; break;
      case DT_I: first_elem = I_NI; break;
      case DT_V: first_elem = V_EPI; break;
      case DT_DIODE: first_elem = D_EPP; break;
      case DT_MOS0: first_elem = M0_EDD; break;
      case DT_MOS1: first_elem = M1_E0; break;
      default: first_elem = d.n_itab;
    }
    for (int k = first_elem; k < d.n_itab; k++) {
      int& h = itab[(size_t)d.itab_off + (size_t)k];
      if (h >= 0) h = P.elem_slot[(size_t)h];
    }
    for (int k : d.push_g) {  // Bsim4: element handles are interleaved with variable slots
      int& h = itab[(size_t)d.itab_off + (size_t)k];
      if (h >= 0) h = P.elem_slot[(size_t)h];
    }
  }
  return itab;
}

// Diagnostic, host-only: the CUDA source the run-time specialised kernels would be compiled from, for the plan that the
// given first-iteration matrix values produce (all parameters taken as shared). shape 0 = one thread per instance
// (host/jit.hpp), 1 = team (host/jit_team.hpp). Used by the CPU test-suite to check that the generators' output compiles.
inline std::string debug_jit_source(const FlatCkt& flat, int mode, int shape, const double* vals, size_t* smem_out) {
  Plan P = build_plan<double>(flat.n_vars(), flat.elem_row, flat.elem_col, vals);
  if (P.status != ST_OK) throw S21Error(P.status, "the symbolic phase fails on these values");
  const StageInfo si = make_stage_info(flat);
  const std::vector<int> itab = itab_with_slots(flat, P);
  build_gather(flat, si, mode, itab, P);
  int n_par = 0;
  for (const FlatDev& d : flat.devs) n_par = std::max(n_par, d.par_off + d.n_par);
  std::vector<int> pcode((size_t)n_par);
  for (int j = 0; j < n_par; j++) pcode[(size_t)j] = j << 1;
  if (shape == 0) {
    *smem_out = ((size_t)P.nnzLU + 3 * (size_t)P.N) * 8 * (size_t)jit::TPB;
    return jit::source(flat, P, itab, pcode, mode == AN_TRAN);
  }
  if (!jit::team_eligible(flat, P, (size_t)227 * 1024)) throw S21Error(ST_UNSUPPORTED, "circuit not eligible for the team kernel");
  const int lpi = jit::team_lpi(P.N, jit::team_heavy_devices(flat));
  size_t Bdbg = 0;  // S21_JIT_SOURCE_B: the batch size the generator's shape decisions are taken for (host-only debugging)
  if (const char* e = std::getenv("S21_JIT_SOURCE_B")) Bdbg = (size_t)std::atoll(e);
  return jit::team_source(flat, P, si, itab, pcode, mode == AN_TRAN, lpi, smem_out, 148, jit::team_gi(Bdbg, 148, lpi, mode == AN_TRAN), nullptr, Bdbg);
}

struct PlanDevice {
  Plan host;
  bool valid = false;
  std::vector<int> host_itab;         // itab with element handles translated to L+U slots
  jit::Kernel jit_dcop, jit_tran;     // circuit-specialised kernels (host/jit.hpp), compiled on first use
  bool jit_tried_dcop = false, jit_tried_tran = false;
  jit::Kernel jitt_dcop, jitt_tran;   // team-shaped specialised kernels (host/jit_team.hpp)
  bool jitt_tried_dcop = false, jitt_tried_tran = false;
  DBuf<int> row_i2e, col_i2e, col_e2i, rowptr, colidx, diag_slot, l_off, l_slot, l_row, piv_chk, upd_off, upd_t, upd_u, upd_l, itab;
  // cooperative kernel: every shared index table packed into one allocation ("arena") so that a CTA can bring all of
  // them into shared memory with a single TMA bulk copy. Offsets are in ints, each table 16-byte aligned.
  DBuf<int> arena;
  DBuf<int> res_off, res_slot, res_x;  // CoopTables::res_off / res_slot (tolerance-mode plans; outside the arena: the grid-wide kernel never copies it)
  size_t arena_bytes = 0, arena_core_bytes = 0;  // core = all tables but the parameter codes (packed last)
  struct ArenaOffsets {
    size_t type, itab_off, par_off, state_off, itab, pcode, par_direct, row_i2e, col_i2e, col_e2i, rowptr, colidx, diag_slot, stage_off, eval_order,
        asm_off, asm_src, lu_lvl_off, lu_t, lu_u, lu_l, fw_lvl_off, fw_k, fw_row, fw_slot, bw_lvl_off, bw_row;
  } ao;
  DevTables coop_dev(int n_dev, const double* pval, int n_state) const {
    DevTables d;
    d.n_dev = n_dev; d.n_state = n_state; d.pval = pval;
    d.type = arena.p + ao.type; d.itab_off = arena.p + ao.itab_off; d.par_off = arena.p + ao.par_off; d.state_off = arena.p + ao.state_off;
    d.itab = arena.p + ao.itab; d.pcode = arena.p + ao.pcode; d.par_direct = arena.p + ao.par_direct;
    return d;
  }
  PlanTables coop_plan() const {
    PlanTables t;
    t.N = host.N; t.nnz = host.nnzLU;
    t.row_i2e = arena.p + ao.row_i2e; t.col_i2e = arena.p + ao.col_i2e; t.col_e2i = arena.p + ao.col_e2i;
    t.rowptr = arena.p + ao.rowptr; t.colidx = arena.p + ao.colidx; t.diag_slot = arena.p + ao.diag_slot;
    t.l_off = t.l_slot = t.l_row = t.piv_chk = t.upd_off = t.upd_t = t.upd_u = t.upd_l = nullptr;
    return t;
  }
  CoopTables coop() const {
    CoopTables c;
    c.n_stage = host.n_stage; c.stage_off = arena.p + ao.stage_off; c.eval_order = arena.p + ao.eval_order;
    c.asm_off = arena.p + ao.asm_off; c.asm_src = arena.p + ao.asm_src;
    c.n_lu_lvl = (int)host.lu_lvl_off.size() - 1; c.n_fw_lvl = (int)host.fw_lvl_off.size() - 1; c.n_bw_lvl = (int)host.bw_lvl_off.size() - 1;
    c.lu_lvl_off = arena.p + ao.lu_lvl_off; c.lu_t = arena.p + ao.lu_t; c.lu_u = arena.p + ao.lu_u; c.lu_l = arena.p + ao.lu_l;
    c.fw_lvl_off = arena.p + ao.fw_lvl_off; c.fw_k = arena.p + ao.fw_k; c.fw_row = arena.p + ao.fw_row; c.fw_slot = arena.p + ao.fw_slot;
    c.bw_lvl_off = arena.p + ao.bw_lvl_off; c.bw_row = arena.p + ao.bw_row;
    if (host.relaxed) { c.res_off = res_off.p; c.res_slot = res_slot.p; c.res_x = res_x.p; }
    return c;
  }
  PlanTables tables() const {
    PlanTables t;
    t.N = host.N; t.nnz = host.nnzLU;
    t.row_i2e = row_i2e.p; t.col_i2e = col_i2e.p; t.col_e2i = col_e2i.p;
    t.rowptr = rowptr.p; t.colidx = colidx.p; t.diag_slot = diag_slot.p;
    t.l_off = l_off.p; t.l_slot = l_slot.p; t.l_row = l_row.p; t.piv_chk = piv_chk.p;
    t.upd_off = upd_off.p; t.upd_t = upd_t.p; t.upd_u = upd_u.p; t.upd_l = upd_l.p;
    return t;
  }
};

struct Override {
  std::string kind, name, param;
  std::vector<double> values;  // [B]
};

class Batch {
 public:
  Batch(const CktSpec& spec, const FlatCkt& flat, int device, size_t B) : spec_(spec), flat_(flat), device_(device), B_(B) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
      throw S21Error(ST_CUDA, "no CUDA device available: libspice21cu has no CPU fallback (" + std::string(cudaGetErrorString(e)) + ")");
    if (device < 0 || device >= count) throw S21Error(ST_CUDA, "invalid CUDA device index");
    if (B == 0) throw S21Error(ST_OTHER, "batch size must be positive");
    S21_CUDA(cudaSetDevice(device_));
    S21_CUDA(cudaStreamCreateWithFlags(&own_stream_, cudaStreamNonBlocking));
    stream_ = own_stream_;
    S21_CUDA(cudaEventCreate(&ev0_));
    S21_CUDA(cudaEventCreate(&ev1_));
    const int N = flat_.n_vars();
    // Instance stride of every per-instance table. Rounded to a warp for batches; a single large circuit (C3) keeps its
    // tables dense instead — with stride 32 its 4 M L+U values would be spread over 1 GB and miss L2 on every access.
    Bs_ = (B_ == 1 && N > 96) ? 1 : (B_ + 31) / 32 * 32;
    // (a stride padded to an odd multiple of 32 instances, against power-of-two strides between table entries, was measured on
    // C4: default kernel 180 -> 199 ms, reciprocal-division kernel unchanged — not done; profiles/r02C_c4_modes.txt)
    // shared device tables
    std::vector<int> type, ioff, poff, soff;
    for (const FlatDev& d : flat_.devs) { type.push_back(d.type); ioff.push_back(d.itab_off); poff.push_back(d.par_off); soff.push_back(d.state_off); }
    poff_eff_ = poff;  // rebuild_param_pool() re-points devices with identical parameter blocks at one block
    d_type_.upload(type, stream_); d_ioff_.upload(ioff, stream_); d_poff_.upload(poff, stream_); d_soff_.upload(soff, stream_);
    d_itab_raw_.upload(flat_.itab, stream_);
    si_ = make_stage_info(flat_);
    d_stage_off_.upload(si_.stage_off, stream_);
    d_eval_order_.upload(si_.eval_order, stream_);
    if (const char* k = std::getenv("S21_KERNEL")) {
      const std::string v(k);
      use_coop_ = v != "direct";
      allow_hybrid_ = v != "coop" && v != "direct";
      allow_jit_ = v == "jit" || v == "jitteam";
      jit_forced_ = allow_jit_;
      jit_team_forced_ = v == "jitteam";
      ac_kernel_forced_ = true;
    }
    if (const char* f = std::getenv("S21_B4_FAST")) b4_fast_ = std::atoi(f) != 0;  // kernels/coop_fast.cu
    max_smem_ = (size_t)coop_max_smem_optin(device_);
    if (cudaDeviceGetAttribute(&n_sm_, cudaDevAttrMultiProcessorCount, device_) != cudaSuccess || n_sm_ <= 0) n_sm_ = 148;
    // workspace
    x_.alloc((size_t)N * Bs_); rhs_.alloc((size_t)N * Bs_); c_.alloc((size_t)N * Bs_);
    st_op_.alloc((size_t)std::max(flat_.n_state, 1) * Bs_); st_guess_.alloc((size_t)std::max(flat_.n_state, 1) * Bs_);
    status_.alloc(Bs_); iters_.alloc(Bs_); loads_.alloc(Bs_);
    ensure_lu_rows((size_t)flat_.n_elems());
    params_dirty_ = true;
    reset();
    materialize_reset();
    S21_CUDA(cudaStreamSynchronize(stream_));  // the caller may move the batch to a stream of its own right away
  }
  ~Batch() {
    cudaSetDevice(device_);
    cudaStreamSynchronize(stream_);
    if (ev0_) cudaEventDestroy(ev0_);
    if (ev1_) cudaEventDestroy(ev1_);
    if (own_stream_) cudaStreamDestroy(own_stream_);
  }
  size_t B() const { return B_; }
  int N() const { return flat_.n_vars(); }
  void set_stream(void* s) { stream_ = s ? (cudaStream_t)s : own_stream_; }

  void add_override(const std::string& spec, const double* values) {
    Override o;
    size_t a = spec.find(':'), b = a == std::string::npos ? a : spec.find(':', a + 1);
    if (a == std::string::npos || b == std::string::npos) throw S21Error(ST_OTHER, "bad override spec: " + spec);
    o.kind = spec.substr(0, a); o.name = spec.substr(a + 1, b - a - 1); o.param = spec.substr(b + 1);
    validate_override(o, spec);  // an override that could never take effect is refused here, not at the next solve
    o.values.assign(values, values + B_);
    for (Override& x : overrides_)
      if (x.kind == o.kind && x.name == o.name && x.param == o.param) { x.values = o.values; params_dirty_ = true; rebuild_ = true; return; }
    overrides_.push_back(o);
    params_dirty_ = true;
    rebuild_ = true;
  }

  // kind must be known, name must match a device (card name for the model / instance kinds, flattened path for R / C / I / V),
  // param must be one that kind has: otherwise the override would silently do nothing — or, for an unknown kind, fail every
  // later solve of this batch with no way to remove it.
  void validate_override(const Override& o, const std::string& spec) const {
    auto bad = [&](const std::string& why) { throw S21Error(ST_OTHER, "bad override \"" + spec + "\": " + why); };
    if (o.kind == "opt") {
      if (o.param != "temp") bad("opt takes temp only (Options.gmin is one value per solve: use s21_options)");
      return;
    }
    int type = -1;
    bool by_model = false, by_inst = false;
    if (o.kind == "mos1model") { type = DT_MOS1; by_model = true; }
    else if (o.kind == "mos1inst") { type = DT_MOS1; by_inst = true; }
    else if (o.kind == "diodemodel") { type = DT_DIODE; by_model = true; }
    else if (o.kind == "diodeinst") { type = DT_DIODE; by_inst = true; }
    else if (o.kind == "bsim4model") { type = DT_BSIM4; by_model = true; }
    else if (o.kind == "bsim4inst") { type = DT_BSIM4; by_inst = true; }
    else if (o.kind == "R") type = DT_R;
    else if (o.kind == "C") type = DT_C;
    else if (o.kind == "I") type = DT_I;
    else if (o.kind == "V") type = DT_V;
    else bad("unknown kind (mos1model | mos1inst | diodemodel | diodeinst | bsim4model | bsim4inst | R | C | I | V | opt)");
    bool hit = false;
    for (const FlatDev& d : flat_.devs) {
      if (d.is_ic || d.type != type) continue;
      if (by_model ? d.model == o.name : by_inst ? d.params == o.name : d.path == o.name) { hit = true; break; }
    }
    if (!hit) bad("no device of that kind is named \"" + o.name + "\"");
    const char* want = type == DT_R ? "g" : type == DT_C ? "c" : type == DT_I ? "dc" : nullptr;
    if (want && o.param != want) bad(std::string("param must be ") + want);
    if (type == DT_V && o.param != "dc" && o.param != "acm") bad("param must be dc or acm");
    if (o.param.empty()) bad("empty parameter name");
  }

  // A fresh Solver: x = 0, op = guess = Default (all zeros), counters cleared. Lazy: the hybrid dcop kernel folds the
  // reset into its prologue (it then never reads the old x / state from HBM); every other consumer materialises it.
  void reset() { reset_pending_ = true; }
  void materialize_reset() {
    if (!reset_pending_) return;
    reset_pending_ = false;
    rows_fresh_ = false; host_rows_fresh_ = false;
    S21_CUDA(cudaSetDevice(device_));
    S21_CUDA(cudaMemsetAsync(x_.p, 0, x_.n * sizeof(double), stream_));
    S21_CUDA(cudaMemsetAsync(st_op_.p, 0, st_op_.n * sizeof(double), stream_));
    S21_CUDA(cudaMemsetAsync(st_guess_.p, 0, st_guess_.n * sizeof(double), stream_));
    S21_CUDA(cudaMemsetAsync(status_.p, 0, status_.n * sizeof(int32_t), stream_));
    S21_CUDA(cudaMemsetAsync(iters_.p, 0, iters_.n * sizeof(int32_t), stream_));
    S21_CUDA(cudaMemsetAsync(loads_.p, 0, loads_.n * sizeof(int32_t), stream_));
  }

  // (Re)build the parameter pool on the host (derivations per instance where overridden) and upload it.
  void sync_params(bool force_upload) {
    S21_CUDA(cudaSetDevice(device_));
    if (rebuild_ || pcode_h_.empty()) { rebuild_param_pool(); rebuild_ = false; params_dirty_ = true; codes_dirty_ = true; }
    if (params_dirty_ || force_upload) {
      h2d_bytes_ = 0;
      if (codes_dirty_) {  // the index tables only change with the pool's layout; a forced upload re-sends the values alone
        d_pcode_.upload(pcode_h_, stream_);
        d_poff_.upload(poff_eff_, stream_);
        d_pdirect_.upload(pdirect_h_, stream_);
        codes_dirty_ = false;
        h2d_bytes_ += (pcode_h_.size() + poff_eff_.size()) * sizeof(int);
      }
      d_pval_.alloc(pval_n_);
      trace_mark(0);
      S21_CUDA(cudaMemcpyAsync(d_pval_.p, pval_h_.p, pval_n_ * sizeof(double), cudaMemcpyHostToDevice, stream_));
      trace_mark(1);
      params_dirty_ = false;
      h2d_bytes_ += pval_n_ * sizeof(double);
    }
  }
  size_t last_h2d_bytes() const { return h2d_bytes_; }

  // ---- dcop -----------------------------------------------------------------------------------------------
  // dcop_device(true): the caller is about to read the results on the host (dcop / dcop_view) — a specialised team kernel then
  // writes its result rows straight into the pinned host buffer (mapped: the same pointer is valid on the device), the
  // packing kernel and the D2H copy of a read fall away. S21_HOST_ROWS=0 keeps the rows in HBM (pack + one D2H copy).
  static bool host_rows_enabled() { static const bool on = [] { const char* e = std::getenv("S21_HOST_ROWS"); return !e || std::atoi(e) != 0; }(); return on; }
  // force_upload: this step's parameter pool crosses PCIe again (the per-step input transfer of an end-to-end run) — a
  // cudaMemcpyAsync from the pinned pool. Letting the kernel read the pool from the mapped host copy instead was measured and
  // dropped: the values are re-read every Newton iteration and host memory is not held in L1/L2 across them (kernel 0.099 ->
  // 0.180 ms, profiles/r02p_host_params.txt).
  void dcop_device(bool read_follows = false, bool force_upload = false) {
    S21_CUDA(cudaSetDevice(device_));
    want_host_rows_ = read_follows && host_rows_enabled() && !ext_x_;
    sync_params(false);
    launches_ = 0;
    ensure_plan(op_plan_, AN_OP, 0.0);
    if (force_upload) sync_params(true);
    S21_CUDA(cudaEventRecord(ev0_, stream_)); ev_pair_ = false;
    run_op();
    S21_CUDA(cudaEventRecord(ev1_, stream_)); ev_pair_ = true;
    last_plan_ = &op_plan_;
  }
  // Results of the last solve in the library's own pinned buffer: x rows [instance][variable] + status / iteration counts.
  // Packed on the device, ONE contiguous D2H copy, no host-side copy; the pointers stay valid until the next read.
  void read_view(bool want_x, const double** x, const int32_t** status, const int32_t** iters) {
    S21_CUDA(cudaSetDevice(device_));
    materialize_reset();
    const int N = flat_.n_vars();
    int32_t *hs, *hi, *hl;
    if (want_x && host_rows_fresh_ && !ext_x_) {  // the kernel wrote the rows into hx_ itself: only wait for it
      trace_mark(2); trace_mark(3);
      S21_CUDA(cudaStreamSynchronize(stream_));
      trace_report();
      hs = reinterpret_cast<int32_t*>(hx_.p + (size_t)N * B_);
      hi = hs + B_;
      hl = hi + B_;
    } else if (want_x) {
      const size_t words = rows_words();
      d_rows_.alloc(words);
      if (!ext_x_) hx_.alloc(words);
      if (!rows_fresh_) {
        int rc = launch_pack_out(x_.p, status_.p, iters_.p, loads_.p, d_rows_.p, N, Bs_, (int)B_, stream_);
        launches_++;
        if (rc) throw S21Error(ST_CUDA, std::string("k_pack_out launch failed: ") + cudaGetErrorString((cudaError_t)rc));
      }
      if (ext_x_) {  // a multi-GPU sweep owns one pinned buffer for all shards: x rows land in this shard's slice of it
        S21_CUDA(cudaMemcpyAsync(ext_x_, d_rows_.p, (size_t)N * B_ * sizeof(double), cudaMemcpyDeviceToHost, stream_));
        S21_CUDA(cudaMemcpyAsync(ext_tail_, d_rows_.p + (size_t)N * B_, 3 * B_ * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
        S21_CUDA(cudaStreamSynchronize(stream_));
        hs = ext_tail_;
      } else {
        trace_mark(2);
        S21_CUDA(cudaMemcpyAsync(hx_.p, d_rows_.p, words * sizeof(double), cudaMemcpyDeviceToHost, stream_));
        trace_mark(3);
        S21_CUDA(cudaStreamSynchronize(stream_));
        trace_report();
        hs = reinterpret_cast<int32_t*>(hx_.p + (size_t)N * B_);
      }
      hi = hs + B_;
      hl = hi + B_;
    } else {
      hstatus_.alloc(Bs_); hiters_.alloc(Bs_); hloads_.alloc(Bs_);
      S21_CUDA(cudaMemcpyAsync(hstatus_.p, status_.p, Bs_ * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
      S21_CUDA(cudaMemcpyAsync(hiters_.p, iters_.p, Bs_ * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
      S21_CUDA(cudaMemcpyAsync(hloads_.p, loads_.p, Bs_ * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
      S21_CUDA(cudaStreamSynchronize(stream_));
      hs = hstatus_.p; hi = hiters_.p; hl = hloads_.p;
    }
    // Pivot health. A kernel stops an instance with ST_REPIVOT_CODE when, in one of its factorisations, a frozen pivot is
    // smaller than 1e-3 x an entry below it, i.e. when the reference — which re-runs its Markowitz search in every
    // factorisation (sparse21/mod.rs:929-932, 735-783) — would not have taken that pivot; S21_SINGULAR_MATRIX from a kernel
    // is an exactly zero frozen pivot. Either way the order taken from instance 0's first iteration does not suit this
    // instance's matrix at its current iterate. After a dcop those instances are continued with an order of their own
    // (resolve_repivot); elsewhere (inside a transient's time loop, an AC sweep) the code is reported as Pivot Search Fail.
    std::vector<size_t> flagged;
    for (size_t i = 0; i < B_; i++) {
      const int32_t st = hs[i];
      if (st == ST_REPIVOT_CODE || (st == ST_SINGULAR && last_plan_ && last_plan_->host.status == ST_OK && !no_resolve_))
        flagged.push_back(i);
    }
    if (!no_resolve_) weak_seen_ += (long long)flagged.size();
    if (!flagged.empty() && want_x && last_plan_ == &op_plan_ && pivot_repair_enabled() && !no_resolve_)
      resolve_repivot(flagged, ext_x_ ? ext_x_ : hx_.p, hs, hi, hl);
    if (!no_resolve_)
      for (size_t i = 0; i < B_; i++)
        if (hs[i] == ST_REPIVOT_CODE) hs[i] = ST_PIVOT;  // not continued (repair off / not a dcop): the reference's "Pivot Search Fail"
    if (aids_ && want_x && last_plan_ == &op_plan_ && repair_depth_ == 0) {
      std::vector<size_t> failed;
      for (size_t i = 0; i < B_; i++) if (hs[i] == ST_CONV) failed.push_back(i);
      if (!failed.empty()) aid_dcop(failed, ext_x_ ? ext_x_ : hx_.p, hs, hi, hl);
    }
    sum_iters_ = 0; sum_loads_ = 0;
    for (size_t i = 0; i < B_; i++) {
      sum_iters_ += hi[i];
      sum_loads_ += hl[i];
    }
    if (x) *x = want_x ? (ext_x_ ? ext_x_ : hx_.p) : nullptr;
    if (status) *status = hs;
    if (iters) *iters = hi;
    float ms = 0.f;
    // only a closed pair is queried (a read between two launches — the OP rescue inside tran / ac — has none to report)
    if (ev_pair_ && cudaEventElapsedTime(&ms, ev0_, ev1_) == cudaSuccess) last_ms_ = ms;
    else if (ev_pair_) (void)cudaGetLastError();
  }
  // ---- convergence aids (second half; opt-in — the reference has neither: `src_factor` and `diag_gmin` are
  // dead fields, analysis.rs:659-660, so with the aids off a failing instance reports Convergence Failed exactly as the
  // reference does). Instances that failed a dcop are gathered into a batch of their own and continued:
  //   bit 0, gmin stepping: the junction gmin of every device (Options.gmin, read by Mos1 / Diode / Bsim4) starts at 1e-2 S
  //     and falls a decade per solve down to its target, every solve warm-started from the previous one;
  //   bit 1, source stepping: every independent source is scaled by 0.1, 0.2, ... 1.0, warm-started likewise — each solve
  //     has the reference's 100 iterations and 1 V step limit, so a chain that needs more than 100 iterations to settle
  //     (a long ring released from one IC) gets there in stages.
  // x and the committed device state of the rescued instances replace the failed ones (host rows and HBM columns).
  void set_aids(int flags) { aids_ = flags; }
  // The operating point that opens a transient / AC sweep: instances a kernel stopped for a new pivot order — and, with the
  // aids on, instances whose OP failed — are continued before the analysis goes on (their x and device-state columns are
  // replaced in HBM, their status updated). Costs one D2H of the status vector when nothing needs doing.
  void rescue_op() {
    if (no_resolve_) return;
    hstatus_.alloc(Bs_);
    S21_CUDA(cudaMemcpyAsync(hstatus_.p, status_.p, Bs_ * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
    S21_CUDA(cudaStreamSynchronize(stream_));
    bool need = false;
    for (size_t i = 0; i < B_ && !need; i++) {
      const int32_t st = hstatus_.p[i];
      need = st == ST_REPIVOT_CODE || (st == ST_SINGULAR && op_plan_.host.status == ST_OK) || (aids_ && st == ST_CONV);
    }
    if (!need) return;
    last_plan_ = &op_plan_;
    const double* hx; const int32_t *hs, *hi;
    read_view(true, &hx, &hs, &hi);  // continues the stopped / failed instances; their columns in HBM are replaced
  }
  long long aided() const { return aided_; }
  void copy_instance_from(const Batch& sub, size_t k, size_t i) {
    const int N = flat_.n_vars();
    S21_CUDA(cudaMemcpy2DAsync(x_.p + i, Bs_ * sizeof(double), sub.x_.p + k, sub.Bs_ * sizeof(double), sizeof(double), (size_t)N,
                               cudaMemcpyDeviceToDevice, stream_));
    if (flat_.n_state > 0) {
      S21_CUDA(cudaMemcpy2DAsync(st_op_.p + i, Bs_ * sizeof(double), sub.st_op_.p + k, sub.Bs_ * sizeof(double), sizeof(double),
                                 (size_t)flat_.n_state, cudaMemcpyDeviceToDevice, stream_));
      S21_CUDA(cudaMemcpy2DAsync(st_guess_.p + i, Bs_ * sizeof(double), sub.st_guess_.p + k, sub.Bs_ * sizeof(double), sizeof(double),
                                 (size_t)flat_.n_state, cudaMemcpyDeviceToDevice, stream_));
    }
  }
  void aid_dcop(const std::vector<size_t>& failed, double* hx, int32_t* hs, int32_t* hi, int32_t* hl) {
    const int N = flat_.n_vars();
    std::vector<size_t> todo = failed;
    auto make_sub = [&](const std::vector<size_t>& list, double src_factor) {
      std::unique_ptr<Batch> sub(new Batch(spec_, flat_, device_, list.size()));
      sub->repair_depth_ = 1;  // no nested aids; pivot repair stays available
      sub->allow_jit_ = false;
      sub->set_stream(stream_);
      std::vector<double> vals(list.size());
      for (const Override& o : overrides_) {
        const bool is_src = (o.kind == "V" || o.kind == "I") && o.param != "acm";
        for (size_t k = 0; k < list.size(); k++) vals[k] = o.values[list[k]] * (is_src ? src_factor : 1.0);
        sub->add_override(o.kind + ":" + o.name + ":" + o.param, vals.data());
      }
      if (src_factor != 1.0)  // sources without a per-instance override: scale their shared value
        for (const FlatDev& d : flat_.devs) {
          if ((d.type != DT_V && d.type != DT_I) || d.is_ic) continue;
          bool has = false;
          for (const Override& o : overrides_) has = has || ((o.kind == (d.type == DT_V ? "V" : "I")) && o.name == d.path && o.param != "acm");
          if (has) continue;
          const double base = flat_.par[(size_t)d.par_off + (size_t)(d.type == DT_V ? VP_V_OP : IP_I)];
          std::fill(vals.begin(), vals.end(), base * src_factor);
          sub->add_override(std::string(d.type == DT_V ? "V:" : "I:") + d.path + ":dc", vals.data());
        }
      return sub;
    };
    auto harvest = [&](Batch& sub, const std::vector<size_t>& list, const int32_t* ss, const int64_t* it_sum, const int64_t* ld_sum,
                       const double* sx) {
      std::vector<size_t> still;
      for (size_t k = 0; k < list.size(); k++) {
        const size_t i = list[k];
        if (ss[k] != ST_OK) { still.push_back(i); continue; }
        std::memcpy(hx + i * (size_t)N, sx + k * (size_t)N, (size_t)N * sizeof(double));
        hs[i] = ST_OK;
        hi[i] += (int32_t)it_sum[k];
        hl[i] += (int32_t)ld_sum[k];
        copy_instance_from(sub, k, i);
        S21_CUDA(cudaMemcpyAsync(status_.p + i, hs + i, sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
        S21_CUDA(cudaMemcpyAsync(iters_.p + i, hi + i, sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
        S21_CUDA(cudaMemcpyAsync(loads_.p + i, hl + i, sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
        aided_++;
      }
      S21_CUDA(cudaStreamSynchronize(stream_));
      return still;
    };
    // run a ladder of warm-started solves on `sub`; per-instance iteration totals accumulate over the rungs
    auto ladder = [&](Batch& sub, size_t n, int rungs, const std::function<void(int)>& set_rung, std::vector<int32_t>* st_out,
                      std::vector<int64_t>* it_sum, std::vector<int64_t>* ld_sum, std::vector<double>* x_out) {
      it_sum->assign(n, 0); ld_sum->assign(n, 0);
      for (int r = 0; r < rungs; r++) {
        set_rung(r);
        sub.dcop_device();  // warm: no reset between rungs
        const double* sx; const int32_t *ss, *si;
        sub.read_view(true, &sx, &ss, &si);
        const int32_t* sl = si + n;
        for (size_t k = 0; k < n; k++) { (*it_sum)[k] += si[k]; (*ld_sum)[k] += sl[k]; }
        if (r == rungs - 1) { st_out->assign(ss, ss + n); x_out->assign(sx, sx + n * (size_t)N); }
      }
    };
    std::vector<int32_t> st;
    std::vector<int64_t> its, lds;
    std::vector<double> xs;
    if ((aids_ & 1) && !todo.empty()) {
      std::unique_ptr<Batch> sub = make_sub(todo, 1.0);
      const double target = flat_.opts.gmin;
      std::vector<double> gs;
      for (double g = 1e-2; g > target * 1.0001; g *= 0.1) gs.push_back(g);
      gs.push_back(target);
      ladder(*sub, todo.size(), (int)gs.size(), [&](int r) { sub->gmin_eff_ = gs[(size_t)r]; }, &st, &its, &lds, &xs);
      todo = harvest(*sub, todo, st.data(), its.data(), lds.data(), xs.data());
    }
    if ((aids_ & 2) && !todo.empty()) {
      // the source values live in the parameter pool: one sub-batch per rung would lose the warm start, so the rungs
      // rewrite the overrides of ONE sub-batch (which re-derives its pool and symbolic phase; x and device state stay)
      std::unique_ptr<Batch> sub = make_sub(todo, 0.1);
      const std::vector<size_t> list = todo;
      ladder(*sub, list.size(), 10,
             [&](int r) {
               if (r == 0) return;
               std::unique_ptr<Batch> next = make_sub(list, 0.1 * (double)(r + 1));
               sub->overrides_ = next->overrides_;
               sub->params_dirty_ = true; sub->rebuild_ = true;
             },
             &st, &its, &lds, &xs);
      todo = harvest(*sub, list, st.data(), its.data(), lds.data(), xs.data());
    }
    rows_fresh_ = false; host_rows_fresh_ = false;
  }
  static bool tran_stop_enabled() { const char* e = std::getenv("S21_TRAN_REPIVOT"); return !e || std::atoi(e) != 0; }
  // Re-pivoting INSIDE a transient (cooperative kernel). The time loop runs on the pivot order frozen at its first iteration;
  // the reference takes a new order in every factorisation (sparse21/mod.rs:930-932). An instance whose frozen order meets an
  // exactly zero pivot — or a vanishing one: a non-finite step — at some time point goes back to its last accepted point and
  // comes back with ST_REPIVOT_CODE and that time point (SolveCtl::tran_stop). Round by round, as for a dcop: the first such
  // instance's load sweep at that point is probed, the reference's Markowitz search orders that matrix, and the kernel is
  // relaunched in resume mode — only those instances run, each from its own time point to the end, against the new order;
  // whichever stops again goes into the next round. S21_TRAN_REPIVOT=0 restores the earlier behaviour (Singular Matrix).
  void resolve_tran_stops(double tstep, int T, size_t n_save) {
    hstatus_.alloc(Bs_);
    S21_CUDA(cudaMemcpyAsync(hstatus_.p, status_.p, B_ * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
    S21_CUDA(cudaStreamSynchronize(stream_));
    std::vector<size_t> open;
    for (size_t i = 0; i < B_; i++) if (hstatus_.p[i] == ST_REPIVOT_CODE) open.push_back(i);
    if (open.empty()) return;
    weak_seen_ += (long long)open.size();
    const int N = flat_.n_vars();
    const size_t budget = std::max<size_t>(10000000, (size_t)4000 * (size_t)N);
    auto give_up = [&](size_t i, int32_t st) {
      S21_CUDA(cudaMemcpyAsync(status_.p + i, &st, sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
      S21_CUDA(cudaStreamSynchronize(stream_));
    };
    const bool info = std::getenv("S21_PLAN_INFO") != nullptr;
    for (int round = 0; round < 64 && !open.empty(); round++) {
      repair_plan_.valid = false;
      ensure_plan(repair_plan_, AN_TRAN, tstep, open[0], budget);
      if (info) {
        int tp0 = -1;
        cudaMemcpy(&tp0, tp_stop_.p + open[0], sizeof(int), cudaMemcpyDeviceToHost);
        std::fprintf(stderr, "[s21 tran] re-pivot round %d: %zu instances handed back, first = %zu at time point %d, plan status %d\n", round,
                     open.size(), open[0], tp0, repair_plan_.host.status);
      }
      if (repair_plan_.host.status != ST_OK) {  // the reference's own factorisation fails on this matrix
        give_up(open[0], repair_plan_.host.status);
        open.erase(open.begin());
        continue;
      }
      SolveCtl ctl = make_ctl(AN_TRAN, tstep, &repair_plan_);
      ctl.tran_stop = 1; ctl.resume = 1;
      CoopCfg cfg = coop_cfg(repair_plan_, B_, 1);
      cfg.tp_stop = tp_stop_.p; cfg.x_acc = x_acc_.p;
      const bool fast = b4_fast_ && ctl.has_bsim4;
      int rc = (fast ? launch_coop_tran_fast : launch_coop_tran)(coop_dev(repair_plan_), repair_plan_.coop_plan(), repair_plan_.coop(), work(),
                                                                 stage_for(cfg, repair_plan_.host), out(), ctl, cfg, T, d_save_.p, (int)n_save,
                                                                 d_wave_.p, stream_);
      launches_++;
      if (rc) throw S21Error(ST_CUDA, std::string("tran resume launch failed: ") + cudaGetErrorString((cudaError_t)rc));
      S21_CUDA(cudaMemcpyAsync(hstatus_.p, status_.p, B_ * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
      S21_CUDA(cudaStreamSynchronize(stream_));
      repivot_rounds_++;
      std::vector<size_t> next;
      for (size_t i : open) {
        if (hstatus_.p[i] == ST_REPIVOT_CODE) next.push_back(i);
        else repaired_++;
      }
      open.swap(next);
    }
    for (size_t i : open) give_up(i, ST_SINGULAR);  // 64 rounds without settling
    rows_fresh_ = false; host_rows_fresh_ = false;
  }
  static bool pivot_stop_enabled() { const char* e = std::getenv("S21_PIVOT_HEALTH"); return !e || std::atoi(e) != 0; }
  static bool pivot_repair_enabled() { const char* e = std::getenv("S21_PIVOT_REPAIR"); return !e || std::atoi(e) != 0; }
  // Per-instance re-pivoting (first half): continue, IN PLACE, the instances a kernel stopped with
  // ST_REPIVOT_CODE (or with an exactly zero frozen pivot). Round by round: the first open instance's load sweep AT ITS
  // CURRENT ITERATE is probed on the GPU, the reference's Markowitz search gives a pivot order for that matrix (so its next
  // factorisation passes the reference's threshold test by construction and every round advances at least that instance by
  // one Newton iteration), and a plan-table kernel is relaunched in resume mode (SolveCtl::resume): only the stopped
  // instances run — warm, with what is left of their own 100 iterations —, every other column of the batch is untouched.
  // Instances the new order suits run on to convergence, the others stop again and go into the next round. No sub-batches,
  // no column copies, no allocation per round: a round costs a probe, a small symbolic phase, one table upload and one
  // launch (the first implementation gathered the open instances into a batch of their own per round: C1-circuit supply
  // sweep 295 ms, C4 429 ms end to end, against 33 / 212 ms without the repair).
  void resolve_repivot(const std::vector<size_t>& flagged, double* hx, int32_t* hs, int32_t* hi, int32_t* hl) {
    const int N = flat_.n_vars();
    std::vector<size_t> open = flagged;
    bool rewrite = false;
    for (size_t i : open) if (hs[i] != ST_REPIVOT_CODE) { hs[i] = ST_REPIVOT_CODE; rewrite = true; }  // zero frozen pivot: continue it as well
    hstatus_.alloc(Bs_); hiters_.alloc(Bs_); hloads_.alloc(Bs_);
    if (rewrite) {
      std::memcpy(hstatus_.p, hs, B_ * sizeof(int32_t));
      S21_CUDA(cudaMemcpyAsync(status_.p, hstatus_.p, B_ * sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
    }
    const size_t budget = std::max<size_t>(10000000, (size_t)4000 * (size_t)N);
    std::vector<int> sing_retries_(B_, 0);
    for (int round = 0; round < 256 && !open.empty(); round++) {
      repair_plan_.valid = false;
      ensure_plan(repair_plan_, AN_OP, 0.0, open[0], budget);
      if (repair_plan_.host.status != ST_OK) {  // the reference's own factorisation fails on this matrix (or the attempt ran over budget)
        const int32_t st = repair_plan_.host.status;
        hs[open[0]] = st;
        S21_CUDA(cudaMemcpyAsync(status_.p + open[0], hs + open[0], sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
        S21_CUDA(cudaStreamSynchronize(stream_));
        open.erase(open.begin());
        continue;
      }
      launch_plan_dcop(repair_plan_, true);
      S21_CUDA(cudaMemcpyAsync(hstatus_.p, status_.p, B_ * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
      S21_CUDA(cudaStreamSynchronize(stream_));
      repivot_rounds_++;
      // An order taken at open[0]'s matrix can hold an exactly zero pivot for ANOTHER instance (its devices sit in other
      // regions): the kernel reports Singular Matrix with x untouched. That instance is not singular under an order of its
      // own — it stays open (measured on the plain 41-stage Bsim4 ring: 7 of 16 supplies ended their OP with Singular Matrix
      // after 2 iterations inside the batch, every one of them converges alone). Truly singular matrices end when they head
      // the list: the host's own factorisation then says so; the 256 rounds bound the loop. (A cap of 8 such retries per
      // instance was too low for a 2048-instance batch with ~60 rounds: 160 instances of the 41-stage ring sweep were dropped.)
      std::vector<size_t> next;
      bool rewrite_st = false;
      for (size_t i : open) {
        hs[i] = hstatus_.p[i];
        if (hs[i] == ST_SINGULAR && sing_retries_[i] < 200) { sing_retries_[i]++; hs[i] = ST_REPIVOT_CODE; hstatus_.p[i] = ST_REPIVOT_CODE; rewrite_st = true; }
        if (hs[i] == ST_REPIVOT_CODE) next.push_back(i);
        else repaired_++;
      }
      if (rewrite_st) S21_CUDA(cudaMemcpyAsync(status_.p, hstatus_.p, B_ * sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
      open.swap(next);
    }
    for (size_t i : open) {  // 256 rounds without settling
      hs[i] = ST_PIVOT;
      S21_CUDA(cudaMemcpyAsync(status_.p + i, hs + i, sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
    }
    // the continued instances' rows, counters and codes: pack everything again (one kernel, one copy)
    d_rows_.alloc(rows_words());
    int rc = launch_pack_out(x_.p, status_.p, iters_.p, loads_.p, d_rows_.p, N, Bs_, (int)B_, stream_);
    launches_++;
    if (rc) throw S21Error(ST_CUDA, std::string("k_pack_out launch failed: ") + cudaGetErrorString((cudaError_t)rc));
    PinnedBuf<double>& stage = repack_;
    stage.alloc(rows_words());
    S21_CUDA(cudaMemcpyAsync(stage.p, d_rows_.p, rows_words() * sizeof(double), cudaMemcpyDeviceToHost, stream_));
    S21_CUDA(cudaStreamSynchronize(stream_));
    const int32_t* ts = reinterpret_cast<const int32_t*>(stage.p + (size_t)N * B_);
    for (size_t i : flagged) {
      std::memcpy(hx + i * (size_t)N, stage.p + i * (size_t)N, (size_t)N * sizeof(double));
      hs[i] = ts[i]; hi[i] = ts[B_ + i]; hl[i] = ts[2 * B_ + i];
    }
    rows_fresh_ = true; host_rows_fresh_ = false;
  }
  long long weak_seen() const { return weak_seen_; }
  long long repaired() const { return repaired_; }
  // Device-resident results of the last solve in the host's layout ([x rows B*N f64][status B][iters B][loads B] i32, packed by
  // k_pack_out): what a multi-process job hands to its collective (NCCL gather of per-instance solutions and convergence
  // flags) without a detour through host memory. Valid until the next solve on this batch.
  void packed_device(const double** dptr, size_t* words) {
    S21_CUDA(cudaSetDevice(device_));
    materialize_reset();
    d_rows_.alloc(rows_words());
    if (!rows_fresh_) {
      int rc = launch_pack_out(x_.p, status_.p, iters_.p, loads_.p, d_rows_.p, flat_.n_vars(), Bs_, (int)B_, stream_);
      launches_++;
      if (rc) throw S21Error(ST_CUDA, std::string("k_pack_out launch failed: ") + cudaGetErrorString((cudaError_t)rc));
      rows_fresh_ = true;
    }
    *dptr = d_rows_.p;
    *words = rows_words();
  }
  // Waveforms of the last transient as the kernel left them in HBM: [T][n_save][stride] f64, instance fastest.
  void wave_device(const double** dptr, size_t* T, size_t* n_save, size_t* stride) const {
    *dptr = d_wave_.p; *T = wave_T_; *n_save = wave_ns_; *stride = Bs_;
  }
  // Results of read_view go to caller-provided pinned memory instead of the batch's own buffer: x rows [B][N] to x_dst,
  // status / iters / loads (3 x B int32) to tail_dst. Used by the multi-GPU sweep (host/sweep.hpp); nullptr restores.
  void set_result_target(double* x_dst, int32_t* tail_dst) { ext_x_ = x_dst; ext_tail_ = tail_dst; }
  int device() const { return device_; }
  double symbolic_seconds() const { return symbolic_s_; }
  size_t rows_words() const { return (size_t)flat_.n_vars() * B_ + (3 * B_ * sizeof(int32_t) + 7) / 8; }
  void read(double* x, int32_t* status, int32_t* iters) {
    const double* hx = nullptr;
    const int32_t *hs = nullptr, *hi = nullptr;
    read_view(x != nullptr, &hx, &hs, &hi);
    if (x) std::memcpy(x, hx, (size_t)flat_.n_vars() * B_ * sizeof(double));
    if (status) std::memcpy(status, hs, B_ * sizeof(int32_t));
    if (iters) std::memcpy(iters, hi, B_ * sizeof(int32_t));
  }

  // ---- tran -----------------------------------------------------------------------------------------------
  void tran(double tstep, int T, const int32_t* save_vars, size_t n_save, double* wave, int32_t* status, int64_t* iters) {
    S21_CUDA(cudaSetDevice(device_));
    const bool info = std::getenv("S21_PLAN_INFO") != nullptr;
    auto t_ph = std::chrono::steady_clock::now();
    auto phase = [&](const char* what) {  // S21_PLAN_INFO: host wall clock per phase of the call (synchronises: diagnostics only)
      if (!info) return;
      cudaStreamSynchronize(stream_);
      const auto now = std::chrono::steady_clock::now();
      std::fprintf(stderr, "[s21 tran] %-28s %9.3f ms (weak %lld, repaired %lld, re-pivot rounds %lld)\n", what,
                   std::chrono::duration<double>(now - t_ph).count() * 1e3, weak_seen_, repaired_, repivot_rounds_);
      t_ph = now;
    };
    materialize_reset();
    sync_params(false);
    launches_ = 0;
    phase("reset + parameters");
    ensure_plan(op_plan_, AN_OP, 0.0);
    phase("OP plan (probe + symbolic)");
    S21_CUDA(cudaEventRecord(ev0_, stream_)); ev_pair_ = false;
    run_op();
    phase("OP solve");
    rescue_op();
    phase("OP rescue");
    // The matrix changes character after the OP (capacitor companions appear, IC resistors are released):
    // take the pivot order again from the first transient iteration.
    tran_plan_.valid = false;
    rows_fresh_ = false; host_rows_fresh_ = false;  // the transient moves x on
    ensure_plan(tran_plan_, AN_TRAN, tstep);
    phase("TRAN plan (probe + symbolic)");
    std::vector<int> sv(save_vars, save_vars + n_save);
    d_save_.upload(sv, stream_);
    S21_CUDA(cudaEventRecord(ev0_, stream_)); ev_pair_ = false;  // time the transient kernel itself: the symbolic phase above is host work
    d_wave_.alloc((size_t)T * n_save * Bs_);
    wave_T_ = (size_t)T; wave_ns_ = n_save;
    DevTables dt = dev_tables(tran_plan_.itab.p);
    SolveCtl ctl = make_ctl(AN_TRAN, tstep);
    int rc = 0;
    bool coop_tran_stop = false;
    if (tran_plan_.host.status != ST_OK) {
      throw S21Error(tran_plan_.host.status, status_text(tran_plan_.host.status));
    } else if (const jit::Kernel* jk = jit_kernel(tran_plan_, true)) {
      last_kernel_ = jk->team ? "jit-team" : "jit-thread";
      rc = launch_jit(*jk, AN_TRAN, tstep, false, T, d_save_.p, (int)n_save, d_wave_.p);
    } else if (CoopCfg hcfg; use_coop_ && use_hybrid(tran_plan_, 1, &hcfg)) {
      last_kernel_ = "hybrid";
      rc = launch_hybrid_tran(coop_dev(tran_plan_), tran_plan_.coop_plan(), tran_plan_.coop(), work(), out(), ctl, hcfg, T, d_save_.p,
                              (int)n_save, d_wave_.p, stream_);
    } else if (use_coop_ && use_grid()) {
      CoopCfg cfg = coop_cfg(tran_plan_, B_, 1);
      cfg.smem_bytes = 0; cfg.mixed = false;
      d_gctl_.alloc(1);
      last_kernel_ = "grid";
      rc = launch_grid_tran(coop_dev(tran_plan_), tran_plan_.coop_plan(), tran_plan_.coop(), work(), stage_for(cfg, tran_plan_.host), out(), ctl,
                            d_gctl_.p, T, d_save_.p, (int)n_save, d_wave_.p, stream_);
    } else if (use_coop_) {
      CoopCfg cfg = coop_cfg(tran_plan_, B_, 1);
      const bool fast = b4_fast_ && ctl.has_bsim4;
      last_kernel_ = fast ? "coop-rcp" : "coop";
      coop_tran_stop = tran_stop_enabled();
      if (coop_tran_stop) {  // instances whose frozen order fails inside the time loop come back for a pivot order of their own
        ctl.tran_stop = 1;
        if (const char* inj = std::getenv("S21_TRAN_INJECT")) ctl.tran_inject_tp = std::atoi(inj);
        tp_stop_.alloc(Bs_); x_acc_.alloc((size_t)flat_.n_vars() * Bs_);
        cfg.tp_stop = tp_stop_.p; cfg.x_acc = x_acc_.p;
      }
      rc = (fast ? launch_coop_tran_fast : launch_coop_tran)(coop_dev(tran_plan_), tran_plan_.coop_plan(), tran_plan_.coop(), work(),
                                                             stage_for(cfg, tran_plan_.host), out(), ctl, cfg, T, d_save_.p, (int)n_save,
                                                             d_wave_.p, stream_);
    } else {
      last_kernel_ = "direct";
      rc = launch_tran(dt, tran_plan_.tables(), work(), out(), ctl, T, d_save_.p, (int)n_save, d_wave_.p, stream_);
    }
    launches_++;
    if (rc) throw S21Error(ST_CUDA, std::string("tran kernel launch failed: ") + cudaGetErrorString((cudaError_t)rc));
    S21_CUDA(cudaEventRecord(ev1_, stream_)); ev_pair_ = true;
    phase("time loop kernel");
    if (info && last_kernel_ == "grid") grid_phase_report();
    last_plan_ = &tran_plan_;
    if (coop_tran_stop) {
      resolve_tran_stops(tstep, T, n_save);
      phase("time loop: re-pivot + resume");
    }
    const bool packed = wave && wave_fetch_begin(T, n_save);
    std::vector<int32_t> it32(B_);
    read(nullptr, status, it32.data());
    if (iters) for (size_t i = 0; i < B_; i++) iters[i] = it32[i];
    if (wave) wave_fetch_end(packed, T, n_save, wave);
  }
  // Waveforms to the caller's [instance][time point][saved variable] layout: transposed on the device (k_pack_wave), one
  // contiguous D2H into pinned memory, one memcpy. Falls back to the strided host pass when the launch is refused.
  bool wave_fetch_begin(int T, size_t n_save) {
    const size_t M = (size_t)T * n_save;
    bool packed = false;
    if (!std::getenv("S21_WAVE_HOST_TRANSPOSE")) {
      d_wave_rows_.alloc(B_ * M);
      packed = launch_pack_wave(d_wave_.p, d_wave_rows_.p, M, Bs_, (int)B_, stream_) == 0;
      if (packed) launches_++;
    }
    if (packed) {
      hwave_.alloc(B_ * M);
      S21_CUDA(cudaMemcpyAsync(hwave_.p, d_wave_rows_.p, B_ * M * sizeof(double), cudaMemcpyDeviceToHost, stream_));
    } else {
      hwave_.alloc(M * Bs_);
      S21_CUDA(cudaMemcpyAsync(hwave_.p, d_wave_.p, M * Bs_ * sizeof(double), cudaMemcpyDeviceToHost, stream_));
    }
    return packed;
  }
  void wave_fetch_end(bool packed, int T, size_t n_save, double* wave) {  // after a stream synchronise (read())
    const size_t M = (size_t)T * n_save;
    if (packed) { std::memcpy(wave, hwave_.p, B_ * M * sizeof(double)); return; }
    for (size_t i = 0; i < B_; i++)
      for (size_t m = 0; m < M; m++) wave[i * M + m] = hwave_.p[m * Bs_ + i];
  }

  // ---- adaptive transient (opt-in) -------------------------------------------------------------------
  // OP, IC release, then per-instance LTE-controlled Backward Euler on the device (kernels/newton.cu::k_tran_adaptive),
  // waveforms on the print grid k * tstep. ctl7 = {h0, hmin, hmax, trtol, reltol, vntol, reserved}; a non-positive / NaN
  // entry takes its default (tstep / 16, tstep * 1e-9, 4 * tstep, 7, 1e-3, 1e-6 — SPICE's trtol / reltol / vntol).
  void tran_adaptive(double tstep, int T, const double* ctl7, const int32_t* save_vars, size_t n_save, double* wave, int32_t* status,
                     int64_t* iters, int32_t* accepted, int32_t* rejected) {
    S21_CUDA(cudaSetDevice(device_));
    materialize_reset();
    sync_params(false);
    launches_ = 0;
    ensure_plan(op_plan_, AN_OP, 0.0);
    run_op();
    rescue_op();
    auto pick = [&](int k, double dflt) { return (ctl7 && ctl7[k] > 0.0) ? ctl7[k] : dflt; };
    AdaptiveArgs g;
    g.tstep = tstep; g.T = T;
    g.h0 = pick(0, tstep / 16.0); g.hmin = pick(1, tstep * 1e-9); g.hmax = pick(2, 4.0 * tstep);
    g.trtol = pick(3, 7.0); g.reltol = pick(4, 1e-3); g.vntol = pick(5, 1e-6);
    // the frozen pivot order is taken at the first step size; the companion conductances C/h move with h, so the pivot-health
    // flag (bit 8 of the status word) is what tells whether the order stayed adequate
    tran_plan_.valid = false;
    rows_fresh_ = false; host_rows_fresh_ = false;
    ensure_plan(tran_plan_, AN_TRAN, g.h0);
    if (tran_plan_.host.status != ST_OK) throw S21Error(tran_plan_.host.status, status_text(tran_plan_.host.status));
    std::vector<int> sv(save_vars, save_vars + n_save);
    d_save_.upload(sv, stream_);
    const int N = flat_.n_vars();
    d_wave_.alloc((size_t)T * n_save * Bs_);
    wave_T_ = (size_t)T; wave_ns_ = n_save;
    ad_x1_.alloc((size_t)N * Bs_); ad_xs_.alloc((size_t)N * Bs_); ad_st_.alloc((size_t)std::max(flat_.n_state, 1) * Bs_);
    ad_acc_.alloc(Bs_); ad_rej_.alloc(Bs_);
    g.x1 = ad_x1_.p; g.xs = ad_xs_.p; g.st_save = ad_st_.p; g.accepted = ad_acc_.p; g.rejected = ad_rej_.p;
    S21_CUDA(cudaEventRecord(ev0_, stream_)); ev_pair_ = false;
    last_kernel_ = "direct-adaptive";
    int rc = launch_tran_adaptive(dev_tables(tran_plan_.itab.p), tran_plan_.tables(), work(), out(), make_ctl(AN_TRAN, g.h0), g, d_save_.p, (int)n_save,
                                  d_wave_.p, stream_);
    launches_++;
    if (rc) throw S21Error(ST_CUDA, std::string("adaptive tran kernel launch failed: ") + cudaGetErrorString((cudaError_t)rc));
    S21_CUDA(cudaEventRecord(ev1_, stream_)); ev_pair_ = true;
    last_plan_ = &tran_plan_;
    const bool packed = wave && wave_fetch_begin(T, n_save);
    std::vector<int32_t> acc(Bs_), rej(Bs_), it32(B_);
    S21_CUDA(cudaMemcpyAsync(acc.data(), ad_acc_.p, Bs_ * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
    S21_CUDA(cudaMemcpyAsync(rej.data(), ad_rej_.p, Bs_ * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
    read(nullptr, status, it32.data());
    for (size_t i = 0; i < B_; i++) {
      if (iters) iters[i] = it32[i];
      if (accepted) accepted[i] = acc[i];
      if (rejected) rejected[i] = rej[i];
    }
    if (wave) wave_fetch_end(packed, T, n_save, wave);
  }

  // ---- ac: frequency points are the batch axis of circuit instance 0 ------------------------------------------
  void ac(const double* freqs, size_t F, double* x_out, int32_t* status, int32_t* iters) {
    S21_CUDA(cudaSetDevice(device_));
    materialize_reset();
    for (const FlatDev& d : flat_.devs)
      if (d.type != DT_R && d.type != DT_C && d.type != DT_V && d.type != DT_MOS1)
        throw S21Error(ST_UNSUPPORTED, "AC Not Implemented For This Component!");  // comps/mod.rs:86-88
    sync_params(false);
    launches_ = 0;
    ensure_plan(op_plan_, AN_OP, 0.0);
    S21_CUDA(cudaEventRecord(ev0_, stream_)); ev_pair_ = false;
    run_op();
    rescue_op();
    {  // the reference stops at the OP error (analysis.rs:775)
      int32_t st0 = 0;
      S21_CUDA(cudaMemcpyAsync(&st0, status_.p, sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
      S21_CUDA(cudaStreamSynchronize(stream_));
      if (st0 != ST_OK) throw S21Error(st0, status_text(st0));
    }
    const size_t Fs = (F + 31) / 32 * 32;
    const int N = flat_.n_vars();
    const double PI = 3.14159265358979323846264338327950288;
    std::vector<double> om(Fs, 0.0);
    for (size_t k = 0; k < F; k++) om[k] = 2.0 * PI * freqs[k];  // analysis.rs:799
    d_omega_.upload(om, stream_);
    ac_status_.alloc(Fs); ac_iters_.alloc(Fs); ac_loads_.alloc(Fs);
    S21_CUDA(cudaMemsetAsync(ac_status_.p, 0, Fs * sizeof(int32_t), stream_));
    S21_CUDA(cudaMemsetAsync(ac_iters_.p, 0, Fs * sizeof(int32_t), stream_));
    S21_CUDA(cudaMemsetAsync(ac_loads_.p, 0, Fs * sizeof(int32_t), stream_));
    zx_.alloc((size_t)N * Fs); zrhs_.alloc((size_t)N * Fs); zc_.alloc((size_t)N * Fs);
    S21_CUDA(cudaMemsetAsync(zx_.p, 0, zx_.n * sizeof(cplx), stream_));  // Variables::from zeroes the values (analysis.rs:59-65)
    zlu_.alloc((size_t)flat_.n_elems() * Fs);
    WorkTables<cplx> w;
    w.x = zx_.p; w.rhs = zrhs_.p; w.c = zc_.p; w.lu = zlu_.p; w.stride = Fs;
    w.st_op = st_op_.p; w.st_guess = st_guess_.p; w.st_stride = Bs_;
    SolveCtl ctl = make_ctl(AN_AC, 0.0);
    ctl.B = (int)F; ctl.omega = d_omega_.p; ctl.par_inst_stride = 0;
    // Direct solve per frequency point by default (engine.hpp SolveCtl::ac_direct); S21_AC_NEWTON=1 keeps the reference's shell.
    const char* ac_newton = std::getenv("S21_AC_NEWTON");
    ctl.ac_direct = (ac_newton && std::atoi(ac_newton) != 0) ? 0 : 1;
    // symbolic phase on the first frequency point
    {
      DevTables dt = dev_tables(d_itab_raw_.p);
      DBuf<cplx> probe;
      probe.alloc((size_t)flat_.n_elems());
      int rc = launch_probe_cplx(dt, w, ctl, flat_.n_elems(), N, 0, probe.p, stream_);
      launches_++;
      if (rc) throw S21Error(ST_CUDA, "k_probe launch failed");
      std::vector<cplx> vals((size_t)flat_.n_elems());
      S21_CUDA(cudaMemcpyAsync(vals.data(), probe.p, vals.size() * sizeof(cplx), cudaMemcpyDeviceToHost, stream_));
      S21_CUDA(cudaStreamSynchronize(stream_));
      ac_plan_.host = build_plan<cplx>(N, flat_.elem_row, flat_.elem_col, vals.data());
      upload_plan(ac_plan_, AN_AC);
    }
    if (ac_plan_.host.status != ST_OK) throw S21Error(ac_plan_.host.status, status_text(ac_plan_.host.status));
    if ((size_t)ac_plan_.host.nnzLU > (size_t)flat_.n_elems()) { zlu_.alloc((size_t)ac_plan_.host.nnzLU * Fs); w.lu = zlu_.p; }
    NewtonOut o;
    o.status = ac_status_.p; o.iters = ac_iters_.p; o.loads = ac_loads_.p;
    DevTables dt = dev_tables(ac_plan_.itab.p);
    int rc;
    CoopCfg hcfg;
    // time the sweep kernel itself, as tran() does for its time loop: the operating point, the probe and the host symbolic phase
    // above were inside this pair until ncu showed the kernel at 1.07 ms of the 1.8 ms reported for C5 (profiles/r02M_c5_full.txt)
    S21_CUDA(cudaEventRecord(ev0_, stream_)); ev_pair_ = false;
    // A long sweep is a huge batch of independent points: one thread per point (kernels/newton.cu::k_ac, HBM-resident
    // workspace, instance-fastest layout = coalesced) fills the GPU by itself and skips every barrier of the co-operative
    // kernels. Measured on C5 (N = 73, nnzLU = 365, 100 000 points): 3.8 ms against 29.8 ms (profiles/r01p_c5.txt).
    // Where the switch sits (profiles/r02AD_c5_small.txt, C5 circuit): 2049 points 0.37 ms co-operative against 0.44 ms thread
    // per point, 6251 points 1.01 against 0.43 ms, 12 500 points (an 8-GPU shard of the 100 000) 2.00 against 0.56 ms.
    const bool ac_direct = !ac_kernel_forced_ && F >= (size_t)4096;
    if (use_coop_ && !ac_direct && use_hybrid(ac_plan_, 2, &hcfg)) {
      last_kernel_ = "hybrid";
      rc = launch_hybrid_ac(coop_dev(ac_plan_), ac_plan_.coop_plan(), ac_plan_.coop(), w, o, ctl, hcfg, stream_);
    } else if (use_coop_ && !ac_direct) {
      CoopCfg cfg = coop_cfg(ac_plan_, F, 2);
      cplx* stage = nullptr;
      if (cfg.smem_bytes == 0 || cfg.mixed) { zstage_.alloc((size_t)ac_plan_.host.n_stage * (Fs + 32)); stage = zstage_.p; }
      last_kernel_ = "coop";
      rc = launch_coop_ac(coop_dev(ac_plan_), ac_plan_.coop_plan(), ac_plan_.coop(), w, stage, o, ctl, cfg, stream_);
    } else {
      // (a level-scheduled variant of the thread-per-point kernel — staged assembly, four independent operations of a
      // dependency level in flight — was built on the reading that this kernel is a chain of dependent round trips; it is
      // not: bit-identical, and SLOWER, 2.19-2.28 ms against 1.73-1.77 ms on C5, by about what its extra staging traffic costs
      // at ~5 TB/s. The kernel is bandwidth-bound in the bytes it really moves; removed. profiles/r02L_c5.txt)
      last_kernel_ = "direct";
      rc = launch_ac(dt, ac_plan_.tables(), w, o, ctl, stream_);
    }
    launches_++;
    if (rc) throw S21Error(ST_CUDA, std::string("ac kernel launch failed: ") + cudaGetErrorString((cudaError_t)rc));
    S21_CUDA(cudaEventRecord(ev1_, stream_)); ev_pair_ = true;
    last_plan_ = &ac_plan_;
    // x leaves the kernels as [variable][frequency point]; the caller's layout is [frequency point][variable]: transposed on
    // the device and copied straight into the caller's buffer (100 000 points x 73 variables = 117 MB: the strided host
    // pass over a pageable staging vector was ~60 of the sweep's 79 ms end to end)
    bool x_packed = false;
    std::vector<cplx> hx;
    if (x_out) {
      if (!std::getenv("S21_WAVE_HOST_TRANSPOSE")) {
        zrows_.alloc(F * (size_t)N);
        x_packed = launch_pack_ac(zx_.p, zrows_.p, (size_t)N, Fs, (int)F, stream_) == 0;
        if (x_packed) launches_++;
      }
      if (x_packed) S21_CUDA(cudaMemcpyAsync(x_out, zrows_.p, F * (size_t)N * sizeof(cplx), cudaMemcpyDeviceToHost, stream_));
      else {
        hx.resize((size_t)N * Fs);
        S21_CUDA(cudaMemcpyAsync(hx.data(), zx_.p, hx.size() * sizeof(cplx), cudaMemcpyDeviceToHost, stream_));
      }
    }
    std::vector<int32_t> hs(Fs), hi(Fs), hl(Fs);
    S21_CUDA(cudaMemcpyAsync(hs.data(), ac_status_.p, Fs * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
    S21_CUDA(cudaMemcpyAsync(hi.data(), ac_iters_.p, Fs * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
    S21_CUDA(cudaMemcpyAsync(hl.data(), ac_loads_.p, Fs * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
    S21_CUDA(cudaStreamSynchronize(stream_));
    sum_iters_ = 0; sum_loads_ = 0;
    for (size_t f = 0; f < F; f++) {
      if (status) status[f] = hs[f];
      if (iters) iters[f] = hi[f];
      sum_iters_ += hi[f]; sum_loads_ += hl[f];
      if (x_out && !x_packed)
        for (int k = 0; k < N; k++) {
          x_out[(f * (size_t)N + (size_t)k) * 2 + 0] = hx[(size_t)k * Fs + f].re;
          x_out[(f * (size_t)N + (size_t)k) * 2 + 1] = hx[(size_t)k * Fs + f].im;
        }
    }
    float ms = 0.f;
    // only a closed pair is queried (a read between two launches — the OP rescue inside tran / ac — has none to report)
    if (ev_pair_ && cudaEventElapsedTime(&ms, ev0_, ev1_) == cudaSuccess) last_ms_ = ms;
    else if (ev_pair_) (void)cudaGetLastError();
    // Frequency points the order taken from the first point does not suit (a kernel stopped them with ST_REPIVOT_CODE:
    // jwC entries grow over the sweep) are solved again as a sweep of their own, whose order comes from ITS first point.
    std::vector<size_t> again;
    for (size_t f = 0; f < F; f++) if (hs[f] == ST_REPIVOT_CODE) again.push_back(f);
    weak_seen_ += (long long)again.size();
    if (!again.empty() && again.size() < F + (size_t)(again[0] != 0) && ac_depth_ < 16) {
      std::vector<double> fs(again.size()), xs(again.size() * (size_t)N * 2);
      std::vector<int32_t> ss(again.size()), is(again.size());
      for (size_t k = 0; k < again.size(); k++) fs[k] = freqs[again[k]];
      const float ms_keep = last_ms_;
      const long long it_keep = sum_iters_, ld_keep = sum_loads_;
      ac_depth_++;
      ac(fs.data(), fs.size(), xs.data(), ss.data(), is.data());
      ac_depth_--;
      last_ms_ += ms_keep; sum_iters_ += it_keep; sum_loads_ += ld_keep;
      for (size_t k = 0; k < again.size(); k++) {
        const size_t f = again[k];
        if (status) status[f] = ss[k];
        if (iters) iters[f] = is[k];
        if (x_out) std::memcpy(x_out + f * (size_t)N * 2, xs.data() + k * (size_t)N * 2, (size_t)N * 2 * sizeof(double));
        repaired_++;
      }
    } else if (status) {
      for (size_t f : again) status[f] = ST_PIVOT;
    }
  }

  const Plan* last_plan() const { return last_plan_ ? &last_plan_->host : nullptr; }
  const char* kernel_name() const { return last_kernel_; }
  void stats(double* out8) const {
    out8[0] = launches_; out8[1] = last_ms_; out8[2] = (double)sum_iters_; out8[3] = (double)sum_loads_;
    out8[4] = flat_.n_elems(); out8[5] = last_plan_ ? last_plan_->host.nnzLU : 0; out8[6] = flat_.n_vars(); out8[7] = flat_.n_stamps;
  }
  static const char* status_text(int st) {
    switch (st) {
      case ST_CONV: return "Convergence Failed";
      case ST_SINGULAR: return "Singular Matrix";
      case ST_PIVOT: return "Pivot Search Fail";
      case ST_UNSUPPORTED: return "AC Not Implemented For This Component!";
      default: return "Error";
    }
  }

 private:
  CktSpec spec_;
  FlatCkt flat_;
  int device_;
  size_t B_, Bs_ = 0;
  cudaStream_t own_stream_ = nullptr, stream_ = nullptr;
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
  bool ev_pair_ = false;  // ev0_ and ev1_ bracket a finished launch sequence
  // S21_TRACE_E2E=1 (diagnostics): device time stamps around the H2D copy of the parameter pool, the solve and the D2H copy
  // of the result rows, and the host clock from the first mark to the return of the synchronise, printed per read.
  cudaEvent_t tr_ev_[4] = {nullptr, nullptr, nullptr, nullptr};
  std::chrono::steady_clock::time_point tr_t0_;
  static bool trace_on() { static const bool on = [] { const char* e = std::getenv("S21_TRACE_E2E"); return e && std::atoi(e) != 0; }(); return on; }
  void trace_mark(int k) {
    if (!trace_on()) return;
    if (!tr_ev_[k]) cudaEventCreate(&tr_ev_[k]);
    if (k == 0) tr_t0_ = std::chrono::steady_clock::now();
    cudaEventRecord(tr_ev_[k], stream_);
  }
  void trace_report() {
    if (!trace_on() || !tr_ev_[0] || !tr_ev_[3] || !ev_pair_) return;
    const double host_us = std::chrono::duration<double>(std::chrono::steady_clock::now() - tr_t0_).count() * 1e6;
    float h2d = 0, gap1 = 0, kern = 0, gap2 = 0, d2h = 0;
    cudaEventElapsedTime(&h2d, tr_ev_[0], tr_ev_[1]); cudaEventElapsedTime(&gap1, tr_ev_[1], ev0_); cudaEventElapsedTime(&kern, ev0_, ev1_);
    cudaEventElapsedTime(&gap2, ev1_, tr_ev_[2]); cudaEventElapsedTime(&d2h, tr_ev_[2], tr_ev_[3]);
    std::fprintf(stderr, "[s21 e2e] h2d %.1f us, gap %.1f, solve %.1f, gap(+pack) %.1f, d2h %.1f; host first mark -> sync return %.1f us\n", h2d * 1e3,
                 gap1 * 1e3, kern * 1e3, gap2 * 1e3, d2h * 1e3, host_us);
    (void)cudaGetLastError();
  }
  DBuf<int> d_type_, d_ioff_, d_poff_, d_soff_, d_itab_raw_, d_pcode_, d_pdirect_, d_save_;
  DBuf<double> d_pval_, x_, rhs_, c_, lu_, st_op_, st_guess_, d_wave_, d_omega_, d_rows_;
  DBuf<GridCtl> d_gctl_;
  void grid_phase_report() {  // S21_PLAN_INFO: where the grid-wide kernel's launch went (GridCtl::phase_ns)
    GridCtl h;
    if (cudaMemcpy(&h, d_gctl_.p, sizeof(GridCtl), cudaMemcpyDeviceToHost) != cudaSuccess) return;
    static const char* names[GridCtl::kPhases] = {"evaluation", "assembly", "residual+decision", "LU", "forward", "backward", "limit+update"};
    double tot = 0.0;
    for (int k = 0; k < GridCtl::kPhases; k++) tot += (double)h.phase_ns[k];
    std::fprintf(stderr, "[s21 grid] phases of the launch (ms):");
    for (int k = 0; k < GridCtl::kPhases; k++) std::fprintf(stderr, " %s %.3f", names[k], (double)h.phase_ns[k] * 1e-6);
    std::fprintf(stderr, " | total %.3f, %d loads, huge gather lists %d\n", tot * 1e-6, h.nld, h.n_huge);
    const Plan& P = tran_plan_.host;
    const std::vector<int>* offs[3] = {&P.lu_lvl_off, &P.fw_lvl_off, &P.bw_lvl_off};
    static const char* what[3] = {"LU", "forward", "backward"};
    for (int w = 0; w < 3; w++) {
      std::fprintf(stderr, "[s21 grid] %s levels (items: ms):", what[w]);
      for (size_t q = 0; q + 1 < offs[w]->size() && q < (size_t)GridCtl::kLevels; q++)
        std::fprintf(stderr, " %d: %.3f", (*offs[w])[q + 1] - (*offs[w])[q], (double)h.level_ns[w][q] * 1e-6);
      std::fprintf(stderr, "\n");
    }
    {  // backward: entries per level (the rows' U parts)
      std::fprintf(stderr, "[s21 grid] backward entries per level:");
      for (size_t q = 0; q + 1 < P.bw_lvl_off.size(); q++) {
        long long n = 0; int mx = 0;
        for (int r = P.bw_lvl_off[q]; r < P.bw_lvl_off[q + 1]; r++) {
          const int k = P.bw_row[(size_t)r];
          const int len = P.rowptr[(size_t)k + 1] - P.diag_slot[(size_t)k] - 1;
          n += len; mx = std::max(mx, len);
        }
        std::fprintf(stderr, " %lld (max row %d)", n, mx);
      }
      std::fprintf(stderr, "\n");
    }
  }
  DBuf<double> ad_x1_, ad_xs_, ad_st_;   // adaptive transient scratch
  DBuf<int32_t> ad_acc_, ad_rej_;
  double symbolic_s_ = 0.0;  // host time spent in build_plan (diagnostics)
  DBuf<cplx> zx_, zrhs_, zc_, zlu_;
  DBuf<int32_t> status_, iters_, loads_, ac_status_, ac_iters_, ac_loads_;
  PinnedBuf<double> pval_h_, hx_, hwave_;
  DBuf<double> d_wave_rows_;  // waveforms in the caller's layout (k_pack_wave)
  DBuf<int> tp_stop_;         // cooperative transient kernel: time point at which an instance was handed back (resolve_tran_stops)
  DBuf<double> x_acc_;        // ... and every instance's x at its last accepted time point
  double* ext_x_ = nullptr;       // set_result_target
  size_t wave_T_ = 0, wave_ns_ = 0;
  int repair_depth_ = 0;
  bool no_resolve_ = false;
  int ac_depth_ = 0;
  int max_iter_ = 100;           // iteration budget of the next real solves (a solve continued after a re-pivot gets the rest)
  int aids_ = 0;                 // set_aids: bit 0 gmin stepping, bit 1 source stepping (both off: the reference has neither)
  double gmin_eff_ = -1.0;       // >= 0: overrides Options.gmin for the next solves (gmin stepping)
  long long aided_ = 0;
  long long weak_seen_ = 0, repaired_ = 0, repivot_rounds_ = 0;
  int32_t* ext_tail_ = nullptr;
  PinnedBuf<int32_t> hstatus_, hiters_, hloads_;
  std::vector<int> pcode_h_, poff_eff_, pdirect_h_;
  size_t pval_n_ = 0, h2d_bytes_ = 0;
  std::vector<Override> overrides_;
  bool params_dirty_ = true, rebuild_ = true, codes_dirty_ = true, rows_fresh_ = false;
  bool host_rows_fresh_ = false, want_host_rows_ = false;  // result rows of the last solve already in hx_ (written by the kernel) host_rows_fresh_ = false;
  PlanDevice op_plan_, tran_plan_, ac_plan_;
  PlanDevice repair_plan_;          // resolve_repivot: the order taken at a stopped instance's iterate
  DBuf<int32_t> iters_base_;        // iters_ at the start of the current solve (only kept for warm starts)
  bool iters_base_valid_ = false;
  PinnedBuf<double> repack_;        // resolve_repivot: staging of the re-packed result rows
  const PlanDevice* last_plan_ = nullptr;
  size_t lu_rows_ = 0;
  int launches_ = 0;
  float last_ms_ = 0.f;
  const char* last_kernel_ = "";
  long long sum_iters_ = 0, sum_loads_ = 0;

  StageInfo si_;
  DBuf<int> d_stage_off_, d_eval_order_;
  DBuf<double> d_stage_;
  DBuf<cplx> zstage_;
  DBuf<cplx> zrows_;  // AC results in the caller's layout (launch_pack_ac)
  bool b4_fast_ = false;  // S21_B4_FAST=1: Bsim4 batches on the cooperative kernel run kernels/coop_fast.cu
  bool use_coop_ = true, allow_hybrid_ = true, allow_jit_ = true, jit_forced_ = false, jit_team_forced_ = false, ac_kernel_forced_ = false;
  std::string jit_error_;  // why the specialised kernel is not in use (empty when it is, or was never wanted)
  bool reset_pending_ = false;
  size_t max_smem_ = 0;
  int n_sm_ = 148;

  // Launch geometry of the cooperative kernel: the largest instance group per CTA that still leaves >= 2 CTAs per SM
  // (148 SMs) and whose workspace fits in shared memory; HBM-resident workspace when even one instance does not fit.
  DevTables coop_dev(const PlanDevice& pd) const { return pd.coop_dev((int)flat_.devs.size(), d_pval_.p, flat_.n_state); }
  // Circuit-specialised kernel (host/jit.hpp) for this plan, or nullptr: not wanted, not eligible, or NVRTC unavailable.
  const jit::Kernel* jit_kernel(PlanDevice& pd, bool tran) {
    if (!allow_jit_ || pd.host.status != ST_OK) return nullptr;
    // Two specialised shapes. One thread per instance (host/jit.hpp) has the fewest instructions but the longest
    // dependent chain: it wins once there are enough warps to overlap chains (measured on C2, profiles/r01g_*: 2.5 G
    // iters/s at 1 M instances, but 0.284 ms at 8192 where every warp sits alone on its scheduler). Below that the
    // team shape (host/jit_team.hpp: 8 lanes per instance, rows in registers) is used. S21_KERNEL=jit / jitteam force one.
    // Since the team kernel runs with 2-4 lanes per instance it is ahead of the thread kernel at every batch size measured
    // (C2: 0.084 vs 0.212 ms at 8192, 4.36 vs 3.24 G iters/s at 1 M; profiles/r01t_team_shapes.txt), so the thread kernel
    // is only picked for circuits the team generator does not take.
    const bool thread_ok = !jit_team_forced_ && jit::eligible(flat_, pd.host, max_smem_) &&
                           (jit_forced_ || (B_ >= jit_thread_min_b() && !jit::team_eligible(flat_, pd.host, max_smem_)));
    if (!thread_ok) return jit_team_kernel(pd, tran);
    jit::Kernel& k = tran ? pd.jit_tran : pd.jit_dcop;
    bool& tried = tran ? pd.jit_tried_tran : pd.jit_tried_dcop;
    if (!tried) {
      tried = true;
      std::string err;
      const int tpb = B_ >= (size_t)65536 ? 128 : 64;  // smaller CTAs spread a mid-sized batch over all SMs
      const size_t smem = ((size_t)pd.host.nnzLU + 3 * (size_t)pd.host.N) * 8 * (size_t)tpb;
      const std::string src = jit::source(flat_, pd.host, pd.host_itab, pcode_h_, tran, tpb);
      if (const char* dump = std::getenv("S21_JIT_DUMP")) {
        if (FILE* f = std::fopen(dump, "w")) { std::fwrite(src.data(), 1, src.size(), f); std::fclose(f); }
      }
      if (jit::compile(src, tran, tpb, smem, device_, &k, &err)) {
        k.inst_per_cta = tpb;
      } else {
        k = jit::Kernel();
        jit_error_ = err;
        if (jit_forced_) throw S21Error(ST_CUDA, "S21_KERNEL=jit requested but unavailable: " + err);
      }
    }
    return k.fn ? &k : nullptr;
  }
  static size_t jit_thread_min_b() {
    if (const char* e = std::getenv("S21_JIT_THREAD_MIN_B")) return (size_t)std::atoll(e);
    return 12288;
  }
  const jit::Kernel* jit_team_kernel(PlanDevice& pd, bool tran) {
    if (jit_forced_ && !jit_team_forced_) return nullptr;
    if (!jit::team_eligible(flat_, pd.host, max_smem_)) {
      if (jit_team_forced_) throw S21Error(ST_CUDA, "S21_KERNEL=jitteam requested but the circuit is not eligible (N <= 16, no Bsim4)");
      return nullptr;
    }
    jit::Kernel& k = tran ? pd.jitt_tran : pd.jitt_dcop;
    bool& tried = tran ? pd.jitt_tried_tran : pd.jitt_tried_dcop;
    if (!tried) {
      tried = true;
      std::string err;
      size_t smem = 0;
      const int lpi = jit::team_lpi(pd.host.N, jit::team_heavy_devices(flat_), B_, tran);
      const int gi = jit::team_gi(B_, n_sm_, lpi, tran);
      int tpb = 0;
      const std::string src = jit::team_source(flat_, pd.host, si_, pd.host_itab, pcode_h_, tran, lpi, &smem, n_sm_, gi, &tpb, B_);
      if (const char* dump = std::getenv("S21_JIT_DUMP")) {
        if (FILE* f = std::fopen(dump, "w")) { std::fwrite(src.data(), 1, src.size(), f); std::fclose(f); }
      }
      if (smem <= max_smem_ && jit::compile(src, tran, tpb, smem, device_, &k, &err)) {
        k.inst_per_cta = gi;
        k.team = true;
      } else {
        k = jit::Kernel();
        jit_error_ = err.empty() ? "team kernel: shared memory footprint too large" : err;
        if (jit_team_forced_) throw S21Error(ST_CUDA, "S21_KERNEL=jitteam requested but unavailable: " + jit_error_);
      }
    }
    return k.fn ? &k : nullptr;
  }
  int launch_jit(const jit::Kernel& k, int mode, double dt, bool cold, int T, const int* save_vars, int n_save, double* wave, double* rows = nullptr) {
    const double* pval = d_pval_.p;
    double *gx = x_.p, *sop = st_op_.p, *sg = st_guess_.p;
    int *st = status_.p, *it = iters_.p, *ld = loads_.p;
    size_t stride = Bs_, st_stride = Bs_;
    int B = (int)B_, n_state = flat_.n_state, md = mode, cold_i = cold ? 1 : 0, Tp = T, ns = n_save;
    double gmin = gmin_eff_ >= 0.0 ? gmin_eff_ : flat_.opts.gmin, dtv = dt, reltol = flat_.opts.reltol, iabstol = flat_.opts.iabstol;
    int max_iter = max_iter_;
    void* args[] = {&pval, &gx, &sop, &sg, &st, &it, &ld, &stride, &st_stride, &B, &n_state, &md, &gmin, &dtv, &reltol, &iabstol,
                    &cold_i, &Tp, &ns, &save_vars, &wave, &rows, &max_iter};  // `rows`: written by the warp-private team kernel only
    const unsigned grid = (unsigned)((B_ + (size_t)k.inst_per_cta - 1) / (size_t)k.inst_per_cta);
    return jit::api().cuLaunchKernel(k.fn, grid, 1, 1, (unsigned)k.tpb, 1, 1, (unsigned)k.smem, (void*)stream_, args, nullptr);
  }
  // One large circuit (dense instance stride): all SMs work on it through the grid-wide kernel. S21_KERNEL=coop keeps it on
  // the single-CTA cooperative kernel (the bit-identity tests compare the two).
  bool use_grid() const { return Bs_ == 1 && allow_hybrid_; }
  CoopCfg coop_cfg(const PlanDevice& pd, size_t n_inst, int width) const {
    const Plan& P = pd.host;
    CoopCfg cfg;
    cfg.arena = pd.arena.p;
    cfg.arena_bytes = pd.arena_bytes;
    auto total = [&](int gi, bool with_arena) {
      return coop_ctrl_bytes(gi) + (with_arena ? pd.arena_bytes : 0) + coop_work_bytes(P.N, P.nnzLU, P.n_stage, flat_.n_state, gi, width);
    };
    int gi = 32;
    if (const char* e = std::getenv("S21_COOP_GI")) gi = std::max(1, std::min(32, std::atoi(e)));
    else while (gi > 1 && (n_inst + (size_t)gi - 1) / (size_t)gi < 2 * 148) gi >>= 1;
    int lg = 0;
    while ((1 << lg) < gi) lg++;
    gi = 1 << lg;
    // Bsim4 circuits are bound by device evaluation (thousands of f64 operations per device and iteration), not by the
    // linear algebra: what matters is resident warps during the evaluation sweep. Their per-instance footprint (one
    // staging slot per stamp) would leave one small CTA per SM if it had to sit in shared memory, so the workspace stays
    // in HBM/L2 and the instance group is kept small (<= 8 instances, >= 64 CTAs): measured on C4, 2.8x faster at 2048
    // instances and 1.5x at 256 (profiles/r01f_*).
    int n_b4 = 0;
    for (const FlatDev& d : flat_.devs) n_b4 += d.type == DT_BSIM4;
    const char* force_global = std::getenv("S21_COOP_GLOBAL");  // 1 / 0 overrides the choice
    const bool global_ws = force_global ? std::atoi(force_global) != 0 : n_b4 > 0;
    if (n_b4 > 0) {
      // One CTA per SM (the Bsim4 build of the kernel is compiled for 320 threads x 204 registers: the evaluation spills
      // at 128), every SM the same number of instances, and evaluation rounds of equal size: gi = ceil(B / SMs), devices
      // split evenly over ceil(n_b4 * gi / 320) rounds. C4 (2048 instances, 42 Bsim4 devices): 147 CTAs x 14 instances,
      // 2 rounds of 21 devices x 14 instances = 294 threads.
      if (const char* e = std::getenv("S21_COOP_GI")) gi = std::max(1, std::min(32, std::atoi(e)));
      else gi = (int)std::max<size_t>(1, std::min<size_t>(16, (n_inst + (size_t)n_sm_ - 1) / (size_t)n_sm_));
    }
    while (!global_ws && gi > 1 && total(gi, false) > max_smem_) gi >>= 1;
    cfg.gi = gi;
    const bool work_fits = !global_ws && total(gi, false) <= max_smem_;
    cfg.smem_bytes = work_fits ? coop_work_bytes(P.N, P.nnzLU, P.n_stage, flat_.n_state, gi, width) : 0;
    // the arena rides along in shared memory when it is small next to what the CTA already uses (occupancy first)
    cfg.arena_in_smem = pd.arena_bytes <= 48 * 1024 && (work_fits ? total(gi, true) : coop_ctrl_bytes(gi) + pd.arena_bytes) <= max_smem_;
    size_t arena_copy = cfg.arena_in_smem ? pd.arena_bytes : 0;
    // A Bsim4 device has ~380 parameter codes: 42 devices make the code table 64 KB while every other table together is half
    // of that, and a device whose block is read directly (par_direct) never looks its codes up. When the whole arena is too
    // large to ride along, its core (device tables, gather lists, level schedules) still does — the Bsim4 build runs one
    // CTA per SM, so up to 96 KB cost no occupancy; the level loops then read their indices from shared memory.
    if (!cfg.arena_in_smem && n_b4 > 0 && !work_fits && pd.arena_core_bytes <= 96 * 1024 && coop_ctrl_bytes(gi) + pd.arena_core_bytes <= max_smem_) {
      cfg.arena_in_smem = true;
      cfg.arena_core_bytes = pd.arena_core_bytes;
      arena_copy = pd.arena_core_bytes;
    }
    if (const char* e = std::getenv("S21_COOP_ARENA")) { if (std::atoi(e) == 0) { cfg.arena_in_smem = false; cfg.arena_core_bytes = 0; arena_copy = 0; } }
    // Mixed workspace: when the whole footprint stays in HBM/L2 (Bsim4: one staging slot per stamp), x / rhs / residual and
    // the L+U values still go to shared memory if they fit — the per-level barriers of the factorisation and of the
    // substitutions then wait on shared-memory latencies instead of L2 round trips (S21_COOP_MIXED=0: all in HBM/L2).
    if (!work_fits) {
      const char* mx = std::getenv("S21_COOP_MIXED");
      const size_t mb = coop_mixed_bytes(P.N, P.nnzLU, gi, width);
      if ((!mx || std::atoi(mx) != 0) && coop_ctrl_bytes(gi) + arena_copy + mb <= max_smem_) {
        cfg.smem_bytes = mb;
        cfg.mixed = true;
      }
    }
    const size_t widest = (size_t)std::max(P.nnzLU + P.N, (int)flat_.devs.size()) * (size_t)gi;
    cfg.threads = widest <= 64 ? 64 : widest <= 1024 ? 128 : 256;
    if (const char* e = std::getenv("S21_COOP_THREADS")) cfg.threads = std::max(32, std::min(256, std::atoi(e) / 32 * 32));
    if (n_b4 > 0) {
      const int max_items = std::max(1, 320 / gi);                         // devices one sweep of the CTA can take
      const int rounds = (n_b4 + max_items - 1) / max_items;
      cfg.threads = gi * ((n_b4 + rounds - 1) / rounds);
      if (const char* e = std::getenv("S21_COOP_THREADS")) cfg.threads = std::max(gi, std::min(320, std::atoi(e)) / gi * gi);
    }
    cfg.threads = std::max(cfg.threads, gi);
    cfg.threads = cfg.threads / gi * gi;  // every thread keeps one instance column: blockDim.x must be a multiple of gi
    return cfg;
  }
  // Small circuits take the hybrid kernel: its whole footprint must leave room for >= 2 CTAs per SM.
  bool use_hybrid(const PlanDevice& pd, int width, CoopCfg* cfg) const {
    if (!allow_hybrid_) return false;
    const Plan& P = pd.host;
    const size_t need = hybrid_smem_bytes(P.N, P.nnzLU, P.n_stage, flat_.n_state, pd.arena_bytes, width);
    if (need > max_smem_ / 2 || P.N > 96) return false;
    cfg->gi = 32; cfg->threads = 256;
    cfg->arena = pd.arena.p; cfg->arena_bytes = pd.arena_bytes; cfg->arena_in_smem = true;
    cfg->smem_bytes = hybrid_work_bytes(P.N, P.nnzLU, P.n_stage, flat_.n_state, width);
    return true;
  }
  double* stage_for(const CoopCfg& cfg, const Plan& P) {
    if (cfg.smem_bytes && !cfg.mixed) return nullptr;
    d_stage_.alloc((size_t)P.n_stage * (Bs_ + 32));  // + 32: the cooperative kernel's CTA-blocked layout rounds the batch up to grid x gi
    return d_stage_.p;
  }

  void ensure_lu_rows(size_t rows) {
    if (rows <= lu_rows_) return;
    lu_.alloc(rows * Bs_);
    lu_rows_ = rows;
  }
  DevTables dev_tables(const int* itab) const {
    DevTables d;
    d.n_dev = (int)flat_.devs.size();
    d.type = d_type_.p; d.itab_off = d_ioff_.p; d.par_off = d_poff_.p; d.state_off = d_soff_.p;
    d.itab = itab; d.pcode = d_pcode_.p; d.pval = d_pval_.p; d.n_state = flat_.n_state;
    d.par_direct = d_pdirect_.p;
    return d;
  }
  WorkTables<double> work() const {
    WorkTables<double> w;
    w.x = x_.p; w.rhs = rhs_.p; w.c = c_.p; w.lu = lu_.p; w.stride = Bs_;
    w.st_op = st_op_.p; w.st_guess = st_guess_.p; w.st_stride = Bs_;
    return w;
  }
  NewtonOut out() const {
    NewtonOut o;
    o.status = status_.p; o.iters = iters_.p; o.loads = loads_.p;
    return o;
  }
  SolveCtl make_ctl(int mode, double dt, const PlanDevice* plan = nullptr) const {
    SolveCtl c;
    c.B = (int)B_; c.mode = mode; c.gmin = gmin_eff_ >= 0.0 ? gmin_eff_ : flat_.opts.gmin; c.dt = dt;
    c.reltol = flat_.opts.reltol; c.iabstol = flat_.opts.iabstol; c.omega = nullptr; c.par_inst_stride = 1;
    c.has_bsim4 = 0;
    for (const FlatDev& d : flat_.devs) if (d.type == DT_BSIM4) { c.has_bsim4 = 1; break; }
    c.max_iter = max_iter_;
    c.stop_on_weak = (mode == AN_OP || mode == AN_AC) && pivot_stop_enabled() ? 1 : 0;
    c.relaxed = (plan ? *plan : mode == AN_OP ? op_plan_ : mode == AN_TRAN ? tran_plan_ : ac_plan_).host.relaxed ? 1 : 0;
    c.weak_mult = pivot_weak_mult();
    if (c.relaxed) {
      c.weak_mult = INFINITY;  // |pivot| * inf < |entry| is never true (and NaN for a zero pivot, which the singular test catches)
      if (const char* e = std::getenv("S21_GRID_WEAK_MULT")) { const double v = std::atof(e); if (v >= 1.0) c.weak_mult = v; }
    }
    return c;
  }
  // One dcop launch on a plan-table kernel (hybrid / grid-wide / cooperative / direct, by circuit shape) against plan `pd`.
  // resume: SolveCtl::resume — continue the stopped instances in place with this plan's pivot order.
  int launch_plan_dcop(PlanDevice& pd, bool resume) {
    SolveCtl ctl = make_ctl(AN_OP, 0.0, &pd);
    ctl.resume = resume ? 1 : 0;
    NewtonOut o = out();
    o.iters_base = iters_base_valid_ ? iters_base_.p : nullptr;
    CoopCfg hcfg;
    int rc;
    if (use_coop_ && use_hybrid(pd, 1, &hcfg)) {
      hcfg.cold = !resume && reset_pending_;
      if (!resume) reset_pending_ = false;
      if (!resume) last_kernel_ = "hybrid";
      rc = launch_hybrid_dcop(coop_dev(pd), pd.coop_plan(), pd.coop(), work(), o, ctl, hcfg, stream_);
    } else if (use_coop_ && use_grid()) {
      materialize_reset();
      CoopCfg cfg = coop_cfg(pd, B_, 1);
      cfg.smem_bytes = 0; cfg.mixed = false;
      d_gctl_.alloc(1);
      if (!resume) last_kernel_ = "grid";
      rc = launch_grid_dcop(coop_dev(pd), pd.coop_plan(), pd.coop(), work(), stage_for(cfg, pd.host), o, ctl, d_gctl_.p, stream_);
    } else if (use_coop_) {
      materialize_reset();
      CoopCfg cfg = coop_cfg(pd, B_, 1);
      const bool fast = b4_fast_ && ctl.has_bsim4;
      if (!resume) last_kernel_ = fast ? "coop-rcp" : "coop";
      rc = (fast ? launch_coop_dcop_fast : launch_coop_dcop)(coop_dev(pd), pd.coop_plan(), pd.coop(), work(), stage_for(cfg, pd.host), o, ctl, cfg,
                                                             stream_);
    } else {
      materialize_reset();
      if (!resume) last_kernel_ = "direct";
      rc = launch_dcop(dev_tables(pd.itab.p), pd.tables(), work(), o, ctl, stream_);
    }
    launches_++;
    if (rc) throw S21Error(ST_CUDA, std::string("dcop kernel launch failed: ") + cudaGetErrorString((cudaError_t)rc));
    return rc;
  }
  void run_op() {
    if (op_plan_.host.status != ST_OK) {  // the reference fails in its first factorisation: every instance reports it
      materialize_reset();
      std::vector<int32_t> st(Bs_, op_plan_.host.status);
      S21_CUDA(cudaMemcpyAsync(status_.p, st.data(), Bs_ * sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
      S21_CUDA(cudaStreamSynchronize(stream_));
      return;
    }
    int rc;
    rows_fresh_ = false; host_rows_fresh_ = false;
    // iteration counts accumulate over warm solves: a continued solve (resolve_repivot) needs to know where THIS one started
    iters_base_valid_ = false;
    if (!reset_pending_) {
      iters_base_.alloc(Bs_);
      S21_CUDA(cudaMemcpyAsync(iters_base_.p, iters_.p, Bs_ * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream_));
      iters_base_valid_ = true;
    }
    if (const jit::Kernel* jk = jit_kernel(op_plan_, false)) {
      const bool cold = reset_pending_;
      reset_pending_ = false;
      last_kernel_ = jk->team ? "jit-team" : "jit-thread";
      double* rows = nullptr;
      bool to_host = false;
      if (jk->team && want_host_rows_) {  // result rows straight into the pinned host buffer (see dcop_device)
        hx_.alloc(rows_words());
        rows = hx_.p;
        to_host = true;
      } else if (jk->team && jit::team_wp(false, B_)) {  // the warp-private team kernel leaves the host's result layout behind in HBM
        d_rows_.alloc(rows_words());
        rows = d_rows_.p;
      }
      rc = launch_jit(*jk, AN_OP, 0.0, cold, 2, nullptr, 0, nullptr, rows);
      rows_fresh_ = rc == 0 && rows != nullptr && !to_host;
      host_rows_fresh_ = rc == 0 && to_host;
    } else {
      rc = launch_plan_dcop(op_plan_, false);
      launches_--;  // counted below
    }
    launches_++;
    if (rc) throw S21Error(ST_CUDA, std::string("dcop kernel launch failed: ") + cudaGetErrorString((cudaError_t)rc));
  }
  // Symbolic phase for one analysis mode: probe one instance's load sweep at its current iterate on the GPU (instance 0 at the
  // start of a solve; the first stopped instance when a solve is continued with a new order), pivot on the host.
  // update_budget > 0: the attempt is abandoned beyond that many Schur updates (host/symbolic.hpp build_plan).
  void ensure_plan(PlanDevice& pd, int mode, double dt, size_t inst = 0, size_t update_budget = 0) {
    if (pd.valid) return;
    materialize_reset();
    const int N = flat_.n_vars();
    DevTables dtab = dev_tables(d_itab_raw_.p);
    DBuf<double> probe;
    probe.alloc((size_t)flat_.n_elems());
    // The probe is one more load sweep of instance 0 on the live workspace; devices with iteration-carried state (Diode
    // limiting, Bsim4 limiting and von) write their in-flight copy during a load, which would make instance 0's next real
    // iteration limit against the probe instead of its own previous iteration. Its guess column is kept and put back.
    DBuf<double> keep;
    const size_t ns_rows = (size_t)flat_.n_state;
    if (ns_rows) {
      keep.alloc(ns_rows);
      S21_CUDA(cudaMemcpy2DAsync(keep.p, sizeof(double), st_guess_.p + inst, Bs_ * sizeof(double), sizeof(double), ns_rows, cudaMemcpyDeviceToDevice, stream_));
    }
    int rc = launch_probe_real(dtab, work(), make_ctl(mode, dt), flat_.n_elems(), N, (int)inst, probe.p, stream_);
    launches_++;
    if (rc) throw S21Error(ST_CUDA, std::string("k_probe launch failed: ") + cudaGetErrorString((cudaError_t)rc));
    if (ns_rows)
      S21_CUDA(cudaMemcpy2DAsync(st_guess_.p + inst, Bs_ * sizeof(double), keep.p, sizeof(double), sizeof(double), ns_rows, cudaMemcpyDeviceToDevice, stream_));
    std::vector<double> vals((size_t)flat_.n_elems());
    S21_CUDA(cudaMemcpyAsync(vals.data(), probe.p, vals.size() * sizeof(double), cudaMemcpyDeviceToHost, stream_));
    S21_CUDA(cudaStreamSynchronize(stream_));
    // (A plan cache keyed on these probe values — skip the derivation when an unchanged batch probes the same bits — was
    // measured and removed: 0.07 ms of a 19 ms call on the C1 sweep, and it cannot hit on C3, whose tolerance-mode operating
    // point is not bit-reproducible. profiles/r02Q_c1_e2e.txt)
    const auto t_sym0 = std::chrono::steady_clock::now();
    // One large circuit on the grid-wide kernel: tolerance-mode level schedules (host/symbolic.hpp build_levels) unless
    // S21_PLAN_EXACT=1 asks for the bit-identical chains.
    const char* exact = std::getenv("S21_PLAN_EXACT");
    const bool relaxed = use_coop_ && use_grid() && !(exact && std::atoi(exact) != 0);
    pd.host = build_plan<double>(N, flat_.elem_row, flat_.elem_col, vals.data(), relaxed, update_budget);
    const auto t_sym1 = std::chrono::steady_clock::now();
    upload_plan(pd, mode);
    symbolic_s_ += std::chrono::duration<double>(t_sym1 - t_sym0).count();
    if (std::getenv("S21_PLAN_INFO"))
      std::fprintf(stderr, "[s21 plan] mode=%d host symbolic phase %.3f s, tables+upload %.3f s\n", mode,
                   std::chrono::duration<double>(t_sym1 - t_sym0).count(),
                   std::chrono::duration<double>(std::chrono::steady_clock::now() - t_sym1).count());
    ensure_lu_rows((size_t)pd.host.nnzLU);
  }
  void upload_plan(PlanDevice& pd, int mode) {
    Plan& P = pd.host;
    pd.row_i2e.upload(P.row_i2e, stream_); pd.col_i2e.upload(P.col_i2e, stream_); pd.col_e2i.upload(P.col_e2i, stream_);
    pd.rowptr.upload(P.rowptr, stream_); pd.colidx.upload(P.colidx, stream_); pd.diag_slot.upload(P.diag_slot, stream_);
    pd.l_off.upload(P.l_off, stream_); pd.l_slot.upload(P.l_slot, stream_); pd.l_row.upload(P.l_row, stream_); pd.piv_chk.upload(P.piv_checked, stream_);
    pd.upd_off.upload(P.upd_off, stream_); pd.upd_t.upload(P.upd_t, stream_); pd.upd_u.upload(P.upd_u, stream_); pd.upd_l.upload(P.upd_l, stream_);
    std::vector<int> itab = itab_with_slots(flat_, P);
    pd.itab.upload(itab, stream_);
    pd.host_itab = itab;
    if (std::getenv("S21_PLAN_INFO"))
      std::fprintf(stderr, "[s21 plan] mode=%d N=%d nnzLU=%d lu_ops=%zu lu_levels=%zu fw_ops=%zu fw_levels=%zu bw_levels=%zu\n", mode, P.N,
                   P.nnzLU, P.lu_t.size(), P.lu_lvl_off.empty() ? 0 : P.lu_lvl_off.size() - 1, P.fw_k.size(),
                   P.fw_lvl_off.empty() ? 0 : P.fw_lvl_off.size() - 1, P.bw_lvl_off.empty() ? 0 : P.bw_lvl_off.size() - 1);
    pd.jit_dcop = jit::Kernel(); pd.jit_tran = jit::Kernel(); pd.jit_tried_dcop = pd.jit_tried_tran = false;
    pd.jitt_dcop = jit::Kernel(); pd.jitt_tran = jit::Kernel(); pd.jitt_tried_dcop = pd.jitt_tried_tran = false;
    if (P.status == ST_OK) {
      build_gather(flat_, si_, mode, itab, P);
      std::vector<int> A;
      auto put = [&](const std::vector<int>& v) {
        const size_t at = A.size();
        A.insert(A.end(), v.begin(), v.end());
        while (A.size() % 4) A.push_back(0);
        return at;
      };
      std::vector<int> type, ioff, soff;
      const std::vector<int>& poff = poff_eff_;
      for (const FlatDev& d : flat_.devs) { type.push_back(d.type); ioff.push_back(d.itab_off); soff.push_back(d.state_off); }
      auto& o = pd.ao;
      o.type = put(type); o.itab_off = put(ioff); o.par_off = put(poff); o.state_off = put(soff); o.itab = put(itab); o.par_direct = put(pdirect_h_);
      o.row_i2e = put(P.row_i2e); o.col_i2e = put(P.col_i2e); o.col_e2i = put(P.col_e2i); o.rowptr = put(P.rowptr); o.colidx = put(P.colidx);
      o.diag_slot = put(P.diag_slot); o.stage_off = put(si_.stage_off); o.eval_order = put(si_.eval_order);
      o.asm_off = put(P.asm_off); o.asm_src = put(P.asm_src);
      o.lu_lvl_off = put(P.lu_lvl_off); o.lu_t = put(P.lu_t); o.lu_u = put(P.lu_u); o.lu_l = put(P.lu_l);
      o.fw_lvl_off = put(P.fw_lvl_off); o.fw_k = put(P.fw_k); o.fw_row = put(P.fw_row); o.fw_slot = put(P.fw_slot);
      o.bw_lvl_off = put(P.bw_lvl_off); o.bw_row = put(P.bw_row);
      pd.arena_core_bytes = A.size() * sizeof(int);
      o.pcode = put(pcode_h_);
      pd.arena_bytes = A.size() * sizeof(int);
      pd.arena.upload(A, stream_);
      if (P.relaxed) {  // rows of A proper in pivoted order (slots with a non-empty gather list), for the grid-wide kernel's residual
        std::vector<int> ro((size_t)P.N + 1, 0), rs, rx;
        for (int r = 0; r < P.N; r++) {
          for (int sl = P.rowptr[(size_t)r]; sl < P.rowptr[(size_t)r + 1]; sl++)
            if (P.asm_off[(size_t)sl + 1] > P.asm_off[(size_t)sl]) { rs.push_back(sl); rx.push_back(P.col_i2e[(size_t)P.colidx[(size_t)sl]]); }
          ro[(size_t)r + 1] = (int)rs.size();
        }
        if (rs.empty()) { rs.push_back(0); rx.push_back(0); }
        pd.res_off.upload(ro, stream_); pd.res_slot.upload(rs, stream_); pd.res_x.upload(rx, stream_);
      }
    }
    S21_CUDA(cudaStreamSynchronize(stream_));  // host vectors above are temporaries
    pd.valid = true;
  }

  // Parameter pool: shared values first (one per (device, param)), then one B-column per parameter that varies.
  void rebuild_param_pool() {
    const size_t n_shared = flat_.par.size();
    std::vector<double> shared = flat_.par;
    pcode_h_.assign(n_shared, 0);
    for (size_t j = 0; j < n_shared; j++) pcode_h_[j] = (int)(j << 1);
    std::vector<std::vector<double>> columns;  // each [B]
    std::vector<size_t> column_target;         // index into the shared table
    auto set_column = [&](size_t j, const std::vector<double>& col) {
      bool same = true;
      for (size_t i = 1; i < B_ && same; i++) same = col[i] == col[0];
      if (same) { shared[j] = col[0]; return; }
      columns.push_back(col);
      column_target.push_back(j);
    };
    if (!overrides_.empty()) {
      for (const FlatDev& d : flat_.devs) {
        if (d.is_ic) continue;
        std::vector<const Override*> mine;
        for (const Override& o : overrides_) {
          bool hit = false;
          if (o.kind == "opt") hit = (d.type == DT_MOS1 || d.type == DT_DIODE) && o.param == "temp";
          else if (o.kind == "mos1model") hit = d.type == DT_MOS1 && d.model == o.name;
          else if (o.kind == "mos1inst") hit = d.type == DT_MOS1 && d.params == o.name;
          else if (o.kind == "diodemodel") hit = d.type == DT_DIODE && d.model == o.name;
          else if (o.kind == "diodeinst") hit = d.type == DT_DIODE && d.params == o.name;
          else if (o.kind == "bsim4model") hit = d.type == DT_BSIM4 && d.model == o.name;
          else if (o.kind == "bsim4inst") hit = d.type == DT_BSIM4 && d.params == o.name;
          else if (o.kind == "R") hit = d.type == DT_R && d.path == o.name;
          else if (o.kind == "C") hit = d.type == DT_C && d.path == o.name;
          else if (o.kind == "I") hit = d.type == DT_I && d.path == o.name;
          else if (o.kind == "V") hit = d.type == DT_V && d.path == o.name;
          else throw S21Error(ST_OTHER, "unknown override kind: " + o.kind);
          if (hit) mine.push_back(&o);
        }
        if (mine.empty()) continue;
        const size_t base = (size_t)d.par_off;
        if (d.type == DT_R || d.type == DT_C || d.type == DT_I || d.type == DT_V) {
          for (const Override* o : mine) {
            if (d.type == DT_R) { set_column(base + RP_G_OP, o->values); set_column(base + RP_G_TRAN, o->values); }
            else if (d.type == DT_C) set_column(base + CP_C, o->values);
            else if (d.type == DT_I) set_column(base + IP_I, o->values);
            else if (o->param == "acm") set_column(base + VP_ACM, o->values);
            else { set_column(base + VP_V_OP, o->values); set_column(base + VP_V_TRAN, o->values); }
          }
          continue;
        }
        // Mos1 / Diode / Bsim4: re-run the reference derivation per instance
        const int np = d.n_par;
        std::vector<std::vector<double>> cols((size_t)np, std::vector<double>(B_));
        unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        size_t nt = std::min<size_t>(hw, std::max<size_t>(1, B_ / 256));
        std::vector<std::thread> th;
        std::vector<std::string> errs(nt);
        for (size_t t = 0; t < nt; t++)
          th.emplace_back([&, t]() {
            try {
              MosModelSpec mm;
              ParamBag mbag, ibag;
              if (d.type == DT_MOS1) { mm = spec_.mos1_models.at(d.model); ibag = spec_.mos1_insts.at(d.params); }
              else if (d.type == DT_BSIM4) { mm = spec_.bsim4_models.at(d.model); ibag = spec_.bsim4_insts.at(d.params); }
              else { mbag = spec_.diode_models.at(d.model); ibag = spec_.diode_insts.at(d.params); }
              for (size_t i = t; i < B_; i += nt) {
                SimOptions o = flat_.opts;
                for (const Override* ov : mine) {
                  const double v = ov->values[i];
                  if (ov->kind == "opt") o.temp = v;
                  else if (ov->kind == "mos1model" || ov->kind == "bsim4model") mm.p.kv[ov->param] = v;
                  else if (ov->kind == "diodemodel") mbag.kv[ov->param] = v;
                  else ibag.kv[ov->param] = v;
                }
                if (d.type == DT_MOS1) {
                  Mos1Derived r = mos1_derive(mm, ibag, o);
                  for (int k = 0; k < np; k++) cols[(size_t)k][i] = r.par[k];
                } else if (d.type == DT_BSIM4) {
                  b4::Derived r = b4::derive_device(mm.mos_type, mm.p.kv, ibag.kv);
                  if (b4::g_push_sequence(r.flavor).size() != d.push_g.size() || b4::b_push_sequence(r.flavor).size() != d.push_b.size())
                    throw std::runtime_error("a Bsim4 override may not change the device topology (rgatemod/rdsmod/rbodymod/trnqsmod)");
                  for (int k = 0; k < np; k++) cols[(size_t)k][i] = r.par[(size_t)k];
                } else {
                  DiodeDerived r = diode_derive(mbag, ibag, o);
                  for (int k = 0; k < np; k++) cols[(size_t)k][i] = r.par[k];
                }
              }
            } catch (const std::exception& e) { errs[t] = e.what(); }
          });
        for (auto& x : th) x.join();
        for (auto& e : errs) if (!e.empty()) throw S21Error(ST_INVALID, e);
        for (int k = 0; k < np; k++) set_column(base + (size_t)k, cols[(size_t)k]);
      }
    }
    pval_n_ = n_shared + columns.size() * Bs_;
    pval_h_.alloc(pval_n_);
    std::memcpy(pval_h_.p, shared.data(), n_shared * sizeof(double));
    for (size_t c = 0; c < columns.size(); c++) {
      double* dst = pval_h_.p + n_shared + c * Bs_;
      std::memcpy(dst, columns[c].data(), B_ * sizeof(double));
      for (size_t i = B_; i < Bs_; i++) dst[i] = columns[c][0];
      pcode_h_[column_target[c]] = (int)(((n_shared + c * Bs_) << 1) | 1);
    }
    // Devices whose parameter blocks are identical (same model card, same size, nothing per-instance) share ONE block:
    // a Bsim4 block is ~5 KB of values plus as many codes, and 42 private copies (config C4) overflow L1, so that every
    // parameter read of the evaluation went to L2 (ncu r01f: long-scoreboard = 36 % of the stalls inside load_bsim4).
    poff_eff_.resize(flat_.devs.size());
    {
      std::vector<size_t> firsts;  // device indices that own a distinct block
      for (size_t k = 0; k < flat_.devs.size(); k++) {
        const FlatDev& d = flat_.devs[k];
        poff_eff_[k] = d.par_off;
        if (d.n_par < 16) continue;  // only the big blocks matter
        bool all_shared = true;
        for (int j = 0; j < d.n_par && all_shared; j++) all_shared = (pcode_h_[(size_t)d.par_off + (size_t)j] & 1) == 0;
        if (!all_shared) continue;
        bool found = false;
        for (size_t f : firsts) {
          const FlatDev& e = flat_.devs[f];
          if (e.type != d.type || e.n_par != d.n_par) continue;
          if (std::memcmp(shared.data() + e.par_off, shared.data() + d.par_off, (size_t)d.n_par * sizeof(double)) == 0) {
            poff_eff_[k] = e.par_off;
            found = true;
            break;
          }
        }
        if (!found) firsts.push_back(k);
      }
    }
    // devices whose whole block is shared values: the kernels read it directly (engine.hpp DevTables::par_direct)
    pdirect_h_.assign(flat_.devs.size(), 0);
    static const bool allow_direct = [] { const char* e = std::getenv("S21_PAR_DIRECT"); return !e || std::atoi(e) != 0; }();
    for (size_t k = 0; k < flat_.devs.size() && allow_direct; k++) {
      const FlatDev& d = flat_.devs[k];
      if (d.type != DT_BSIM4) continue;  // only the Bsim4 evaluation has a direct-block instantiation
      bool all_shared = true;
      for (int j = 0; j < d.n_par && all_shared; j++) all_shared = pcode_h_[(size_t)poff_eff_[k] + (size_t)j] == (int)(((size_t)poff_eff_[k] + (size_t)j) << 1);
      pdirect_h_[k] = all_shared ? 1 : 0;
    }
    // a parameter change invalidates the frozen pivot orders (and the tables packed with them)
    op_plan_.valid = false; tran_plan_.valid = false; ac_plan_.valid = false;
  }
};

}  // namespace s21
