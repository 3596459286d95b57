/* spice21cu.h — C ABI of libspice21cu.so: the B200-native Newton loop behind Spice21's own entry points.
 *
 * Every entry point names the reference interface it replaces (paths relative to the Spice21 tree).
 * Plain pointers and sizes only; no C++ or torch types cross this boundary. All functions return an
 * int32 status (S21_OK == 0) unless stated; s21_last_error() gives the SpError.desc-style text of the last
 * failure on the calling thread. Handles are not thread-safe; distinct handles may be driven from distinct
 * host threads, each bound to one CUDA device / stream. There is no CPU fallback: without a CUDA device the
 * compute entry points fail with S21_CUDA_ERROR.
 */
#ifndef SPICE21CU_H
#define SPICE21CU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes. 1-3 are the reference's SpError strings (spice21/src/analysis.rs:209,302;
 * spice21/src/sparse21/mod.rs:651,665); 5 stands for the reference's panics on malformed circuits
 * (spice21/src/elab.rs:49,97,151,167,188); 6 for the `load_ac` trait default (spice21/src/comps/mod.rs:86-88). */
enum {
  S21_OK = 0,
  S21_CONVERGENCE_FAILED = 1,
  S21_SINGULAR_MATRIX = 2,
  S21_PIVOT_SEARCH_FAIL = 3,
  S21_DECODE_ERROR = 4,
  S21_INVALID_CIRCUIT = 5,
  S21_UNSUPPORTED = 6,
  S21_CUDA_ERROR = 7,
  S21_OTHER = 8
};

/* spice21/src/analysis.rs:642-693 `Options`: the five fields settable through proto SimOptions
 * (spice21/protos/spice21.proto:136-142). NaN means "not given" (reference default applies). */
typedef struct s21_options {
  double temp, tnom, gmin, iabstol, reltol;
} s21_options;

typedef struct s21_ckt s21_ckt;     /* one circuit: definitions + instances, then its elaborated structure */
typedef struct s21_batch s21_batch; /* B independent instances of one elaborated circuit, resident on one GPU */

const char* s21_last_error(void);
void s21_free(uint8_t* p);
/* CUDA device count visible to the library (0 when there is no driver / no GPU). Never fails. */
int32_t s21_cuda_device_count(void);

/* ---- bytes in / bytes out: drop-in for CallableProto::call_bytes (spice21/src/proto.rs:40-44) ----------
 * `Op`, `Tran`, `Ac` messages in (spice21.proto:145,162,197), `OpResult`, `TranResult`, `AcResult` out
 * (spice21.proto:150,175,203). These are what spice21py `_dcop/_tran/_ac` (spice21py/src/lib.rs:41-67) and
 * spice21js (spice21js/native/src/lib.rs:34-66) bind. *out is malloc'ed; release with s21_free. */
int32_t s21_op_bytes(const uint8_t* op, size_t n, uint8_t** out, size_t* out_n);
int32_t s21_tran_bytes(const uint8_t* tran, size_t n, uint8_t** out, size_t* out_n);
int32_t s21_ac_bytes(const uint8_t* ac, size_t n, uint8_t** out, size_t* out_n);

/* ---- circuit construction: Ckt::decode / Ckt::from_proto (spice21/src/circuit.rs:248-330) ---------------- */
int32_t s21_ckt_from_proto(const uint8_t* circuit, size_t n, s21_ckt** out); /* a `Circuit` message */
int32_t s21_ckt_new(s21_ckt** out);                                           /* Ckt::new (circuit.rs:230) */
void s21_ckt_destroy(s21_ckt* c);
/* Builder calls: the Rust-side `Ckt::add` / `Comp::{r,c,idc,vdc}` / `Mosi` / `DiodeI` / `ModuleI` forms
 * (circuit.rs:64-147, 259). `module` is NULL for a top-level instance, else the ModuleDef being filled.
 * An empty string "" is ground (circuit.rs:51-58). */
int32_t s21_ckt_signal(s21_ckt* c, const char* module, const char* name);
int32_t s21_ckt_add_r(s21_ckt* c, const char* module, const char* name, const char* p, const char* n, double g);
int32_t s21_ckt_add_c(s21_ckt* c, const char* module, const char* name, const char* p, const char* n, double cap);
int32_t s21_ckt_add_i(s21_ckt* c, const char* module, const char* name, const char* p, const char* n, double dc);
int32_t s21_ckt_add_v(s21_ckt* c, const char* module, const char* name, const char* p, const char* n, double dc, double acm);
/* EXTENSION (the reference's Vsrc is DC / acm only, spice21.proto:29-35, comps/mod.rs:95-150): a voltage
 * source whose transient value follows SPICE's PULSE (kind 1: v1 v2 td tr tf pw per; per = 0: a single pulse) or SIN (kind 2:
 * vo va freq td theta). The operating point uses `dc` (give it the wave's value at t = 0). On the wire: Vsrc fields 6 (kind)
 * and 7 (repeated double), ignored by the reference's decoder. Evaluated on the device at every time point (fixed-step and
 * adaptive transients). */
int32_t s21_ckt_add_v_wave(s21_ckt* c, const char* module, const char* name, const char* p, const char* n, double dc, double acm,
                           int32_t kind, size_t n_params, const double* params);
int32_t s21_ckt_add_d(s21_ckt* c, const char* module, const char* name, const char* p, const char* n, const char* model,
                      const char* params);
int32_t s21_ckt_add_mos(s21_ckt* c, const char* module, const char* name, const char* model, const char* params, const char* d,
                        const char* g, const char* s, const char* b);
int32_t s21_ckt_add_x(s21_ckt* c, const char* module, const char* name, const char* module_name, size_t n_ports,
                      const char* const* port_names, const char* const* port_nodes);
int32_t s21_ckt_def_module(s21_ckt* c, const char* name, size_t n_ports, const char* const* ports);
/* Definitions (defs.rs:129-135). Parameters are (key, value) pairs; absent keys take the reference defaults.
 * kind: "mos0" (keys ignored; tests.rs:1443-1447), "mos1model" (Mos1Model::resolve, mos.rs:140-237; key "tpg" is
 * the integer gate type), "mos1inst" (mos.rs:271-290), "diodemodel" (diode.rs:52-70), "diodeinst",
 * "bsim4model", "bsim4inst". mos_type: 0 NMOS, 1 PMOS (ignored where meaningless). */
int32_t s21_ckt_define(s21_ckt* c, const char* kind, const char* name, int32_t mos_type, size_t n, const char* const* keys,
                       const double* vals);

/* ---- elaboration: Solver::new + Tran::ic (analysis.rs:308-330, 510-525; elab.rs:244-265) --------------------
 * Flattens the hierarchy, numbers the variables in the reference's first-encounter order, calls every device's
 * create_matrix_elems in component order (which fixes element ids == the stamp map), and appends the two
 * initial-condition devices per `ic` entry. Must be called once before batches are created. */
int32_t s21_ckt_elaborate(s21_ckt* c, const s21_options* opts, size_t n_ic, const char* const* ic_nodes, const double* ic_vals);
int32_t s21_ckt_num_vars(const s21_ckt* c);
const char* s21_ckt_var_name(const s21_ckt* c, int32_t i); /* valid until the circuit is destroyed */
int32_t s21_ckt_var_kind(const s21_ckt* c, int32_t i);     /* analysis.rs:34-38: 0 V, 1 I, 2 Q */
int32_t s21_ckt_num_devices(const s21_ckt* c);
/* Stamp map (sparse21/mod.rs:265 `make`, comps/mod.rs:349 `make_matrix_elem`): element id -> (row, col) in
 * creation order, and per device the element handles in create order (-1 = dropped ground entry).
 * dev_off has num_devices+1 entries. Pointers stay valid until the circuit is destroyed. */
int32_t s21_ckt_stamp_map(const s21_ckt* c, const int32_t** elem_row, const int32_t** elem_col, size_t* n_elem,
                          const int32_t** dev_off, const int32_t** dev_elems);

/* ---- batches: the Newton loop (analysis.rs:169-210, 253-303) over B instances on one GPU ----------------- */
int32_t s21_batch_create(const s21_ckt* c, int32_t cuda_device, size_t B, s21_batch** out);
void s21_batch_destroy(s21_batch* b);
/* Launch on a caller-owned stream (a cudaStream_t passed as void*); NULL = the library's own stream. */
int32_t s21_batch_set_stream(s21_batch* b, void* cuda_stream);
/* Per-instance values for one reference-level parameter: spec is "<kind>:<name>:<param>" with kind one of
 * mos1model | mos1inst | diodemodel | diodeinst | bsim4model | bsim4inst | R | C | I | V | opt  (the model / inst kinds
 * take the card's name, R/C/I/V the flattened instance path with param g | c | dc | dc or acm; opt takes temp with an
 * empty name — Options.gmin is one value per solve, s21_options). The host re-runs the reference's
 * once-per-(model,inst) derivation (mos.rs:320-476, diode.rs:146-212, bsim4inst.rs:9-1378) per instance and keeps only
 * the columns that actually vary. values[B] is host memory. Takes effect at the next solve. An unknown kind, a name that
 * matches no device, or a param the kind does not have is refused with S21_ERROR (nothing is recorded). */
int32_t s21_batch_override(s21_batch* b, const char* spec, const double* values);
/* Rebuild (if overrides changed) and upload the parameter pool from pinned host memory to HBM; with force_upload != 0
 * the host->device copy is repeated even when nothing changed (the per-step input transfer of an end-to-end run).
 * *h2d_bytes (may be NULL) receives the bytes copied. The solves call this implicitly with force_upload = 0. */
int32_t s21_batch_sync_params(s21_batch* b, int32_t force_upload, size_t* h2d_bytes);
/* Convergence aids for s21_batch_dcop / _dcop_view (opt-in, default 0 = the reference's behaviour:
 * its `src_factor` / `diag_gmin` options are dead fields, analysis.rs:659-660, and a solve that exceeds 100 iterations is
 * "Convergence Failed"). flags bit 0: gmin stepping (junction gmin from 1e-2 S down a decade per warm-started solve to
 * Options.gmin); bit 1: source stepping (all independent sources at 0.1, 0.2 ... 1.0 of their value, warm-started).
 * Only the instances that failed are continued, in a batch of their own; the others keep their result. */
int32_t s21_batch_set_aids(s21_batch* b, int32_t flags);
/* Zero the solution guess and all device state (a fresh Solver: x = 0, op = guess = default). */
int32_t s21_batch_reset(s21_batch* b);

/* dcop (analysis.rs:383-388) for every instance. Host outputs (any may be NULL): x[B][N] row-major,
 * status[B] (S21_* per instance), iters[B] = Newton iterations that reached the linear solve. */
int32_t s21_batch_dcop(s21_batch* b, double* x, int32_t* status, int32_t* iters);
/* Same solve with everything left resident in HBM (the timed kernel of bench.py); read back with s21_batch_read. */
int32_t s21_batch_dcop_device(s21_batch* b);
int32_t s21_batch_read(s21_batch* b, double* x, int32_t* status, int32_t* iters);
/* s21_batch_dcop without the last host-side copy: the results stay in the library's pinned staging buffer and *x
 * ([B][N] row-major; pass NULL to skip it), *status and *iters point into it. Valid until the next solve / read on this
 * batch; the caller must not free or write them. (The reference returns an owned Vec, analysis.rs:383-388; a caller that
 * needs ownership copies, which is what s21_batch_dcop does.) */
int32_t s21_batch_dcop_view(s21_batch* b, const double** x, const int32_t** status, const int32_t** iters);
/* One Monte-Carlo / sweep step in ONE call: [flags bit 0] s21_batch_sync_params(force_upload = 1) — the host->device copy of
 * the parameter pool from pinned memory —, [bit 1] s21_batch_reset (cold start), then s21_batch_dcop_view. What a host loop
 * that re-draws its parameters every step calls (the reference's equivalent is a fresh `dcop(ckt, opts)` per sample,
 * analysis.rs:383-388). With a specialised team kernel the result rows are written by the kernel straight into the pinned
 * host buffer (no packing kernel, no separate D2H copy). *h2d_bytes (may be NULL) = bytes uploaded. */
int32_t s21_batch_step_dcop_view(s21_batch* b, int32_t flags, const double** x, const int32_t** status, const int32_t** iters,
                                 size_t* h2d_bytes);
/* Results of the last solve left in HBM in the host's layout — [x as [B][N] f64][status B i32][iters B i32][loads B i32],
 * *n_words f64 words in all — for a caller that hands them to a collective (bench.py: NCCL gather of per-instance solutions
 * and convergence flags across ranks) instead of copying them to the host first. Device pointer, owned
 * by the batch, valid until its next solve; work is enqueued on the batch's stream. */
int32_t s21_batch_packed_device(s21_batch* b, const double** dev_ptr, size_t* n_words);
/* Waveforms of the last s21_batch_tran as the kernel left them in HBM: [T][n_save][stride] f64, instance fastest. */
int32_t s21_batch_wave_device(const s21_batch* b, const double** dev_ptr, size_t* T, size_t* n_save, size_t* stride);
/* Tran::solve (analysis.rs:526-573): OP at t=0, IC release, then fixed-step Backward Euler while t < tstop.
 * n_points_out = number of time points incl. t=0 (decided by the reference's floating-point `t += tstep`).
 * wave[B][T][n_save] and time[T] are host buffers sized by s21_tran_num_points; iters[B] counts all solves.
 * The whole time loop runs on the device against the pivot order taken at its first iteration (the reference takes one per
 * factorisation, sparse21/mod.rs:930-932). On the cooperative kernel (Bsim4 circuits, larger circuits) an instance whose
 * frozen order fails inside the loop — an exactly zero pivot, a non-finite step — is not ended with Singular Matrix: it is
 * handed back at its last accepted time point and continued with an order taken there (S21_TRAN_REPIVOT=0 turns that off);
 * the other kernel families report Singular Matrix for such an instance. An instance that fails has NaN from that time
 * point on in wave[] and its code in status[]. */
int64_t s21_tran_num_points(double tstep, double tstop);
int32_t s21_batch_tran(s21_batch* b, double tstep, double tstop, const int32_t* save_vars, size_t n_save, double* time,
                       double* wave, int32_t* status, int64_t* iters);
/* Adaptive-step transient (opt-in; SURVEY section 8 f1 — the reference integrates with a fixed step only, analysis.rs:553-570,
 * and never reads its own `trtol` / `chgtol` options, :650-656): OP at t = 0, IC release, then every instance advances on
 * its own time axis with Backward Euler, a local-truncation-error estimate per step (the BE solution against the linear
 * extrapolation of the two previous accepted points), step rejection and step-size control — all on the device. Results
 * are returned on the SAME print grid as s21_batch_tran (t_k = k * tstep), by linear interpolation between accepted
 * points. ctl7 (may be NULL) = {h0 first step, hmin, hmax, trtol, reltol, vntol, reserved}; entries <= 0 take the
 * defaults tstep/16, tstep*1e-9, 4*tstep, 7, 1e-3, 1e-6. accepted / rejected [B] (may be NULL) count each instance's
 * steps. The fixed-step entry point above stays the parity path. */
int32_t s21_batch_tran_adaptive(s21_batch* b, double tstep, double tstop, const double* ctl7, const int32_t* save_vars, size_t n_save,
                                double* time, double* wave, int32_t* status, int64_t* iters, int32_t* accepted, int32_t* rejected);
/* ac (analysis.rs:761-832) with the frequency points as the batch axis of ONE circuit instance (B is ignored;
 * the DC operating point is solved first). freqs[F] in Hz as produced by s21_ac_freqs; x[F][N][2] = (re, im). */
int64_t s21_ac_freqs(uint64_t fstart, uint64_t fstop, uint64_t npts, double* freqs, size_t cap);
int32_t s21_batch_ac(s21_batch* b, const double* freqs, size_t F, double* x, int32_t* status, int32_t* iters);

/* Symbolic phase as used by the last solve (sparse21/mod.rs:647-919 run once on the first Newton iteration's
 * values): internal->external row and column orders, and the L+U pattern incl. fill-ins in internal coordinates.
 * Pointers stay valid until the next solve on this batch. */
int32_t s21_batch_pivot_order(const s21_batch* b, const int32_t** row_i2e, const int32_t** col_i2e, size_t* n, const int32_t** lu_row,
                              const int32_t** lu_col, const int32_t** lu_is_fill, size_t* nnz_lu);
/* Counters of the last solve: [0] kernel launches, [1] device milliseconds (CUDA events on the launch stream),
 * [2] sum of Newton iterations, [3] sum of device-load sweeps, [4] nnz(A), [5] nnz(L+U), [6] N, [7] stamp slots. */
int32_t s21_batch_stats(const s21_batch* b, double* out8);
/* Name of the Newton kernel the last solve was dispatched to ("hybrid", "jit-team", "jit-thread", "coop", "grid",
 * "direct"); valid until the next solve. The choice depends on circuit size, device content and batch size. */
const char* s21_batch_kernel_name(const s21_batch* b);

/* Shape of the numeric plan the last solve ran on (host/symbolic.hpp): [0] N, [1] nnz(L+U), [2] LU operations,
 * [3] LU dependency levels, [4] forward-substitution operations, [5] forward levels, [6] backward levels, [7] 1 when the
 * level schedules are in tolerance mode (same-target updates share a level and are applied atomically; the default for one
 * large circuit on the grid-wide kernel, S21_PLAN_EXACT=1 restores the reference's per-value order). The level counts are
 * the barriers per Newton iteration of the level-scheduled kernels. */
int32_t s21_batch_plan_info(const s21_batch* b, int64_t* out8);
/* Setup cost behind the solves of this process (none of it is per Newton iteration): [0] host seconds this batch spent in
 * the symbolic phase (Markowitz order + fill + level schedules), [1] seconds spent inside NVRTC by the whole process,
 * [2] NVRTC compilations, [3] kernels taken from the on-disk cubin cache ($S21_CACHE_DIR, default ~/.cache/spice21cu;
 * "off" disables), [4] from the in-process cache; [5] instances whose solves met a weak frozen pivot (|pivot| < 1e-3 x an
 * entry below it: the reference, which re-pivots every iteration (sparse21/mod.rs:735-783), would have chosen otherwise) or
 * an exactly zero one, [6] instances re-solved with a symbolic phase of their own because of it (dcop; S21_PIVOT_REPAIR=0
 * disables); [7] instances rescued by the convergence aids (s21_batch_set_aids). */
int32_t s21_batch_setup_stats(const s21_batch* b, double* out8);

/* ---- multi-GPU sweeps: ONE process, one host thread + CUDA stream per GPU ---------------------
 * The B instances of a sweep are split into contiguous blocks of ceil(B / n_devices) per device (s21_sweep_partition is
 * that rule, host-only); every device holds its own copy of the shared structure. Instances never exchange data during
 * a solve; at the end every device copies its block of results into its slice of one pinned host buffer, concurrently.
 * The reference has no multi-device path (a Solver is single-threaded, analysis.rs:308-388); the calls mirror the
 * s21_batch_* ones, with [B]-sized host arrays covering the whole sweep. devices == NULL means CUDA devices
 * 0..n_devices-1; n_devices <= 0 means all visible devices. A sweep handle is driven from one host thread. */
typedef struct s21_sweep s21_sweep;
int32_t s21_sweep_partition(size_t B, int32_t n_devices, int32_t g, size_t* first, size_t* count);
int32_t s21_sweep_create(const s21_ckt* c, int32_t n_devices, const int32_t* devices, size_t B, s21_sweep** out);
void s21_sweep_destroy(s21_sweep* s);
int32_t s21_sweep_num_devices(const s21_sweep* s);
int32_t s21_sweep_shard(const s21_sweep* s, int32_t g, int32_t* cuda_device, size_t* first, size_t* count);
int32_t s21_sweep_override(s21_sweep* s, const char* spec, const double* values /* [B] */);
int32_t s21_sweep_sync_params(s21_sweep* s, int32_t force_upload, size_t* h2d_bytes /* sum over devices */);
int32_t s21_sweep_reset(s21_sweep* s);
int32_t s21_sweep_dcop(s21_sweep* s, double* x /* [B][N] */, int32_t* status /* [B] */, int32_t* iters /* [B] */);
/* as s21_batch_dcop_view: pointers into the sweep's pinned gather buffer, valid until the next solve on this sweep */
int32_t s21_sweep_dcop_view(s21_sweep* s, const double** x, const int32_t** status, const int32_t** iters);
int32_t s21_sweep_tran(s21_sweep* s, double tstep, double tstop, const int32_t* save_vars, size_t n_save, double* time,
                       double* wave /* [B][T][n_save] */, int32_t* status, int64_t* iters);
/* the frequency axis sharded over the devices (blocks of ceil(F / n_devices) points), instance 0 of the circuit */
int32_t s21_sweep_ac(s21_sweep* s, const double* freqs, size_t F, double* x /* [F][N][2] */, int32_t* status, int32_t* iters);
/* [0] kernel launches (sum over devices), [1] device milliseconds (max over devices), [2] Newton iterations (sum),
 * [3] load sweeps (sum), [4] nnz(A), [5] nnz(L+U), [6] N, [7] stamp slots */
int32_t s21_sweep_stats(const s21_sweep* s, double* out8);

/* Diagnostics of the run-time specialised kernels; neither needs a GPU. s21_jit_source returns the CUDA
 * source generated for the plan that the given first-iteration matrix values (one per element of s21_ckt_stamp_map)
 * produce in analysis `mode` (0 OP, 1 TRAN), shape 0 = one thread per instance, 1 = team of 8 lanes per instance;
 * *out is malloc'ed (s21_free). s21_jit_check compiles a source for sm_100a with NVRTC and reports the log on error. */
int32_t s21_jit_source(const s21_ckt* c, int32_t mode, int32_t shape, const double* vals, size_t n_vals, uint8_t** out, size_t* out_n,
                       size_t* smem_bytes);
int32_t s21_jit_check(const uint8_t* src, size_t n);
/* GPU self-test: the kernels' f64 division (csrc/scalar.h: shared reciprocal + exact zero-numerator shortcut) against the
 * compiler's IEEE `a / b` on ~n random and edge-case operand pairs, bit for bit. *mismatches must be 0;
 * first4 = {a, b, ours, IEEE} of the first mismatch. */
int32_t s21_selftest_div(uint64_t n, uint64_t seed, uint64_t* mismatches, double* first4);

/* Host-only: run the symbolic phase on an arbitrary matrix (COO, values real when width == 1, interleaved (re, im)
 * when width == 2) and report what s21_batch_pivot_order would. Needs no GPU; used to check pivot-order parity with
 * the reference algorithm (sparse21/mod.rs:647-919) integer for integer. Returns the status the reference's
 * lu_factorize/solve would (S21_OK, S21_SINGULAR_MATRIX, S21_PIVOT_SEARCH_FAIL). Output arrays are caller-sized:
 * row_i2e[n], col_i2e[n], lu_*[cap]; *nnz_lu receives the L+U entry count. */
int32_t s21_symbolic(int32_t n, size_t nnz, const int32_t* rows, const int32_t* cols, const double* vals, int32_t width, int32_t* row_i2e,
                     int32_t* col_i2e, int32_t* lu_row, int32_t* lu_col, int32_t* lu_is_fill, size_t cap, size_t* nnz_lu);

#ifdef __cplusplus
}
#endif
#endif /* SPICE21CU_H */
