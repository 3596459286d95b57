//! Raw FFI to `libspice21cu.so`. One declaration per entry point of `include/spice21cu.h` (generated from it by
//! scripts/gen_rust_ffi.py; the header cites the reference interface each entry replaces). Status codes: 0 = S21_OK, see
//! the header.
#![allow(non_camel_case_types, non_snake_case)]
use std::os::raw::{c_char, c_void};

#[repr(C)]
pub struct s21_ckt {
    _private: [u8; 0],
}
#[repr(C)]
pub struct s21_batch {
    _private: [u8; 0],
}
#[repr(C)]
pub struct s21_sweep {
    _private: [u8; 0],
}
/// `spice21::analysis::Options` (analysis.rs:348-381)
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct s21_options {
    pub temp: f64,
    pub tnom: f64,
    pub gmin: f64,
    pub iabstol: f64,
    pub reltol: f64,
}

extern "C" {
    pub fn s21_last_error() -> *const c_char;
    pub fn s21_free(p: *mut u8);
    pub fn s21_cuda_device_count() -> i32;
    pub fn s21_op_bytes(op: *const u8, n: usize, out: *mut *mut u8, out_n: *mut usize) -> i32;
    pub fn s21_tran_bytes(tran: *const u8, n: usize, out: *mut *mut u8, out_n: *mut usize) -> i32;
    pub fn s21_ac_bytes(ac: *const u8, n: usize, out: *mut *mut u8, out_n: *mut usize) -> i32;
    pub fn s21_ckt_from_proto(circuit: *const u8, n: usize, out: *mut *mut s21_ckt) -> i32;
    pub fn s21_ckt_new(out: *mut *mut s21_ckt) -> i32;
    pub fn s21_ckt_destroy(c: *mut s21_ckt);
    pub fn s21_ckt_signal(c: *mut s21_ckt, module: *const c_char, name: *const c_char) -> i32;
    pub fn s21_ckt_add_r(c: *mut s21_ckt, module: *const c_char, name: *const c_char, p: *const c_char, n: *const c_char, g: f64) -> i32;
    pub fn s21_ckt_add_c(c: *mut s21_ckt, module: *const c_char, name: *const c_char, p: *const c_char, n: *const c_char, cap: f64) -> i32;
    pub fn s21_ckt_add_i(c: *mut s21_ckt, module: *const c_char, name: *const c_char, p: *const c_char, n: *const c_char, dc: f64) -> i32;
    pub fn s21_ckt_add_v(c: *mut s21_ckt, module: *const c_char, name: *const c_char, p: *const c_char, n: *const c_char, dc: f64, acm: f64) -> i32;
    pub fn s21_ckt_add_v_wave(c: *mut s21_ckt, module: *const c_char, name: *const c_char, p: *const c_char, n: *const c_char, dc: f64, acm: f64, kind: i32, n_params: usize, params: *const f64) -> i32;
    pub fn s21_ckt_add_d(c: *mut s21_ckt, module: *const c_char, name: *const c_char, p: *const c_char, n: *const c_char, model: *const c_char, params: *const c_char) -> i32;
    pub fn s21_ckt_add_mos(c: *mut s21_ckt, module: *const c_char, name: *const c_char, model: *const c_char, params: *const c_char, d: *const c_char, g: *const c_char, s: *const c_char, b: *const c_char) -> i32;
    pub fn s21_ckt_add_x(c: *mut s21_ckt, module: *const c_char, name: *const c_char, module_name: *const c_char, n_ports: usize, port_names: *const *const c_char, port_nodes: *const *const c_char) -> i32;
    pub fn s21_ckt_def_module(c: *mut s21_ckt, name: *const c_char, n_ports: usize, ports: *const *const c_char) -> i32;
    pub fn s21_ckt_define(c: *mut s21_ckt, kind: *const c_char, name: *const c_char, mos_type: i32, n: usize, keys: *const *const c_char, vals: *const f64) -> i32;
    pub fn s21_ckt_elaborate(c: *mut s21_ckt, opts: *const s21_options, n_ic: usize, ic_nodes: *const *const c_char, ic_vals: *const f64) -> i32;
    pub fn s21_ckt_num_vars(c: *const s21_ckt) -> i32;
    pub fn s21_ckt_var_name(c: *const s21_ckt, i: i32) -> *const c_char;
    pub fn s21_ckt_var_kind(c: *const s21_ckt, i: i32) -> i32;
    pub fn s21_ckt_num_devices(c: *const s21_ckt) -> i32;
    pub fn s21_ckt_stamp_map(c: *const s21_ckt, elem_row: *mut *const i32, elem_col: *mut *const i32, n_elem: *mut usize, dev_off: *mut *const i32, dev_elems: *mut *const i32) -> i32;
    pub fn s21_batch_create(c: *const s21_ckt, cuda_device: i32, B: usize, out: *mut *mut s21_batch) -> i32;
    pub fn s21_batch_destroy(b: *mut s21_batch);
    pub fn s21_batch_set_stream(b: *mut s21_batch, cuda_stream: *mut c_void) -> i32;
    pub fn s21_batch_override(b: *mut s21_batch, spec: *const c_char, values: *const f64) -> i32;
    pub fn s21_batch_sync_params(b: *mut s21_batch, force_upload: i32, h2d_bytes: *mut usize) -> i32;
    pub fn s21_batch_set_aids(b: *mut s21_batch, flags: i32) -> i32;
    pub fn s21_batch_reset(b: *mut s21_batch) -> i32;
    pub fn s21_batch_dcop(b: *mut s21_batch, x: *mut f64, status: *mut i32, iters: *mut i32) -> i32;
    pub fn s21_batch_dcop_device(b: *mut s21_batch) -> i32;
    pub fn s21_batch_read(b: *mut s21_batch, x: *mut f64, status: *mut i32, iters: *mut i32) -> i32;
    pub fn s21_batch_dcop_view(b: *mut s21_batch, x: *mut *const f64, status: *mut *const i32, iters: *mut *const i32) -> i32;
    pub fn s21_batch_step_dcop_view(b: *mut s21_batch, flags: i32, x: *mut *const f64, status: *mut *const i32, iters: *mut *const i32, h2d_bytes: *mut usize) -> i32;
    pub fn s21_batch_packed_device(b: *mut s21_batch, dev_ptr: *mut *const f64, n_words: *mut usize) -> i32;
    pub fn s21_batch_wave_device(b: *const s21_batch, dev_ptr: *mut *const f64, T: *mut usize, n_save: *mut usize, stride: *mut usize) -> i32;
    pub fn s21_tran_num_points(tstep: f64, tstop: f64) -> i64;
    pub fn s21_batch_tran(b: *mut s21_batch, tstep: f64, tstop: f64, save_vars: *const i32, n_save: usize, time: *mut f64, wave: *mut f64, status: *mut i32, iters: *mut i64) -> i32;
    pub fn s21_batch_tran_adaptive(b: *mut s21_batch, tstep: f64, tstop: f64, ctl7: *const f64, save_vars: *const i32, n_save: usize, time: *mut f64, wave: *mut f64, status: *mut i32, iters: *mut i64, accepted: *mut i32, rejected: *mut i32) -> i32;
    pub fn s21_ac_freqs(fstart: u64, fstop: u64, npts: u64, freqs: *mut f64, cap: usize) -> i64;
    pub fn s21_batch_ac(b: *mut s21_batch, freqs: *const f64, F: usize, x: *mut f64, status: *mut i32, iters: *mut i32) -> i32;
    pub fn s21_batch_pivot_order(b: *const s21_batch, row_i2e: *mut *const i32, col_i2e: *mut *const i32, n: *mut usize, lu_row: *mut *const i32, lu_col: *mut *const i32, lu_is_fill: *mut *const i32, nnz_lu: *mut usize) -> i32;
    pub fn s21_batch_stats(b: *const s21_batch, out8: *mut f64) -> i32;
    pub fn s21_batch_kernel_name(b: *const s21_batch) -> *const c_char;
    pub fn s21_batch_plan_info(b: *const s21_batch, out8: *mut i64) -> i32;
    pub fn s21_batch_setup_stats(b: *const s21_batch, out8: *mut f64) -> i32;
    pub fn s21_sweep_partition(B: usize, n_devices: i32, g: i32, first: *mut usize, count: *mut usize) -> i32;
    pub fn s21_sweep_create(c: *const s21_ckt, n_devices: i32, devices: *const i32, B: usize, out: *mut *mut s21_sweep) -> i32;
    pub fn s21_sweep_destroy(s: *mut s21_sweep);
    pub fn s21_sweep_num_devices(s: *const s21_sweep) -> i32;
    pub fn s21_sweep_shard(s: *const s21_sweep, g: i32, cuda_device: *mut i32, first: *mut usize, count: *mut usize) -> i32;
    pub fn s21_sweep_override(s: *mut s21_sweep, spec: *const c_char, values: *const f64) -> i32;
    pub fn s21_sweep_sync_params(s: *mut s21_sweep, force_upload: i32, h2d_bytes: *mut usize) -> i32;
    pub fn s21_sweep_reset(s: *mut s21_sweep) -> i32;
    pub fn s21_sweep_dcop(s: *mut s21_sweep, x: *mut f64, status: *mut i32, iters: *mut i32) -> i32;
    pub fn s21_sweep_dcop_view(s: *mut s21_sweep, x: *mut *const f64, status: *mut *const i32, iters: *mut *const i32) -> i32;
    pub fn s21_sweep_tran(s: *mut s21_sweep, tstep: f64, tstop: f64, save_vars: *const i32, n_save: usize, time: *mut f64, wave: *mut f64, status: *mut i32, iters: *mut i64) -> i32;
    pub fn s21_sweep_ac(s: *mut s21_sweep, freqs: *const f64, F: usize, x: *mut f64, status: *mut i32, iters: *mut i32) -> i32;
    pub fn s21_sweep_stats(s: *const s21_sweep, out8: *mut f64) -> i32;
    pub fn s21_jit_source(c: *const s21_ckt, mode: i32, shape: i32, vals: *const f64, n_vals: usize, out: *mut *mut u8, out_n: *mut usize, smem_bytes: *mut usize) -> i32;
    pub fn s21_jit_check(src: *const u8, n: usize) -> i32;
    pub fn s21_selftest_div(n: u64, seed: u64, mismatches: *mut u64, first4: *mut f64) -> i32;
    pub fn s21_symbolic(n: i32, nnz: usize, rows: *const i32, cols: *const i32, vals: *const f64, width: i32, row_i2e: *mut i32, col_i2e: *mut i32, lu_row: *mut i32, lu_col: *mut i32, lu_is_fill: *mut i32, cap: usize, nnz_lu: *mut usize) -> i32;
}
