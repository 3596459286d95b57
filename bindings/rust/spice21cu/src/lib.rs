//! Safe wrapper over `spice21cu-sys`.
//!
//! * [`dcop_bytes`], [`tran_bytes`], [`ac_bytes`] are drop-ins for `CallableProto::call_bytes`
//!   (spice21/src/proto.rs:40-44): an encoded `Op` / `Tran` / `Ac` message in, an encoded result message out.
//! * [`Circuit`] / [`Batch`] expose the instance-batched Newton loop (Monte-Carlo, parameter sweeps, AC points).
use spice21cu_sys as sys;
use std::ffi::{CStr, CString};
use std::os::raw::c_char;
use std::ptr;

/// Mirrors `spice21::SpError`: the message is the reference's own error string where one exists.
#[derive(Debug, Clone)]
pub struct SpError {
    pub status: i32,
    pub desc: String,
}
pub type SpResult<T> = Result<T, SpError>;

fn check(status: i32) -> SpResult<()> {
    if status == 0 {
        return Ok(());
    }
    let desc = unsafe { CStr::from_ptr(sys::s21_last_error()) }.to_string_lossy().into_owned();
    Err(SpError { status, desc })
}

fn call_bytes(f: unsafe extern "C" fn(*const u8, usize, *mut *mut u8, *mut usize) -> i32, msg: &[u8]) -> SpResult<Vec<u8>> {
    let mut out: *mut u8 = ptr::null_mut();
    let mut n: usize = 0;
    check(unsafe { f(msg.as_ptr(), msg.len(), &mut out, &mut n) })?;
    let v = unsafe { std::slice::from_raw_parts(out, n) }.to_vec();
    unsafe { sys::s21_free(out) };
    Ok(v)
}
/// `Op::call_bytes` (proto.rs:53-77)
pub fn dcop_bytes(op: &[u8]) -> SpResult<Vec<u8>> {
    call_bytes(sys::s21_op_bytes, op)
}
/// `Tran::call_bytes` (proto.rs:79-112)
pub fn tran_bytes(tran: &[u8]) -> SpResult<Vec<u8>> {
    call_bytes(sys::s21_tran_bytes, tran)
}
/// `Ac::call_bytes` (proto.rs:114-145)
pub fn ac_bytes(ac: &[u8]) -> SpResult<Vec<u8>> {
    call_bytes(sys::s21_ac_bytes, ac)
}

/// An elaborated circuit (`Ckt::from_proto` + `Solver::new`).
pub struct Circuit {
    h: *mut sys::s21_ckt,
}
impl Circuit {
    pub fn from_proto(circuit: &[u8], opts: Option<sys::s21_options>, ic: &[(&str, f64)]) -> SpResult<Self> {
        let mut h = ptr::null_mut();
        check(unsafe { sys::s21_ckt_from_proto(circuit.as_ptr(), circuit.len(), &mut h) })?;
        let c = Circuit { h };
        let names: Vec<CString> = ic.iter().map(|(n, _)| CString::new(*n).unwrap()).collect();
        let ptrs: Vec<*const c_char> = names.iter().map(|s| s.as_ptr()).collect();
        let vals: Vec<f64> = ic.iter().map(|(_, v)| *v).collect();
        let o = opts.as_ref().map(|o| o as *const _).unwrap_or(ptr::null());
        check(unsafe { sys::s21_ckt_elaborate(c.h, o, ic.len(), ptrs.as_ptr(), vals.as_ptr()) })?;
        Ok(c)
    }
    pub fn num_vars(&self) -> usize {
        unsafe { sys::s21_ckt_num_vars(self.h) as usize }
    }
    pub fn var_name(&self, i: usize) -> String {
        unsafe { CStr::from_ptr(sys::s21_ckt_var_name(self.h, i as i32)) }.to_string_lossy().into_owned()
    }
}
impl Drop for Circuit {
    fn drop(&mut self) {
        unsafe { sys::s21_ckt_destroy(self.h) }
    }
}

/// `b` independent instances of one circuit, resident on one GPU.
pub struct Batch<'c> {
    h: *mut sys::s21_batch,
    ckt: &'c Circuit,
    b: usize,
}
pub struct DcopResult {
    /// `[instance][variable]`, row-major
    pub x: Vec<f64>,
    pub status: Vec<i32>,
    pub iters: Vec<i32>,
}
impl<'c> Batch<'c> {
    pub fn new(ckt: &'c Circuit, device: i32, b: usize) -> SpResult<Self> {
        let mut h = ptr::null_mut();
        check(unsafe { sys::s21_batch_create(ckt.h, device, b, &mut h) })?;
        Ok(Batch { h, ckt, b })
    }
    /// Per-instance override, e.g. `"mos1inst:default:w"`, `"bsim4inst:n:delvto"`, `"V:vsup:dc"`, `"opt:_:temp"`.
    pub fn set(&mut self, spec: &str, values: &[f64]) -> SpResult<()> {
        assert_eq!(values.len(), self.b);
        let s = CString::new(spec).unwrap();
        check(unsafe { sys::s21_batch_override(self.h, s.as_ptr(), values.as_ptr()) })
    }
    /// Batched `dcop` (analysis.rs:383-388): one launch, all instances.
    pub fn dcop(&mut self) -> SpResult<DcopResult> {
        let n = self.ckt.num_vars();
        let mut r = DcopResult { x: vec![0.0; n * self.b], status: vec![0; self.b], iters: vec![0; self.b] };
        check(unsafe { sys::s21_batch_dcop(self.h, r.x.as_mut_ptr(), r.status.as_mut_ptr(), r.iters.as_mut_ptr()) })?;
        Ok(r)
    }
}
/// Borrowed results of [`Batch::dcop_view`]: slices into the library's pinned staging buffer.
pub struct DcopView<'b> {
    pub x: &'b [f64],
    pub status: &'b [i32],
    pub iters: &'b [i32],
}
impl<'c> Batch<'c> {
    /// `dcop` without the final host-side copy (`s21_batch_dcop_view`); the borrow ends before the next solve.
    pub fn dcop_view(&mut self) -> SpResult<DcopView<'_>> {
        let n = self.ckt.num_vars();
        let (mut x, mut st, mut it) = (ptr::null(), ptr::null(), ptr::null());
        check(unsafe { sys::s21_batch_dcop_view(self.h, &mut x, &mut st, &mut it) })?;
        Ok(unsafe {
            DcopView {
                x: std::slice::from_raw_parts(x, n * self.b),
                status: std::slice::from_raw_parts(st, self.b),
                iters: std::slice::from_raw_parts(it, self.b),
            }
        })
    }
}
impl<'c> Batch<'c> {
    /// One sweep step in one call (`s21_batch_step_dcop_view`): forced upload of the parameter pool, cold start, dcop;
    /// the result rows are written by the kernel into the library's pinned buffer. Returns the view and the bytes uploaded.
    pub fn step_dcop_view(&mut self, upload: bool, reset: bool) -> SpResult<(DcopView<'_>, usize)> {
        let n = self.ckt.num_vars();
        let (mut x, mut st, mut it) = (ptr::null(), ptr::null(), ptr::null());
        let mut h2d: usize = 0;
        let flags = (upload as i32) | ((reset as i32) << 1);
        check(unsafe { sys::s21_batch_step_dcop_view(self.h, flags, &mut x, &mut st, &mut it, &mut h2d) })?;
        Ok((
            unsafe {
                DcopView {
                    x: std::slice::from_raw_parts(x, n * self.b),
                    status: std::slice::from_raw_parts(st, self.b),
                    iters: std::slice::from_raw_parts(it, self.b),
                }
            },
            h2d,
        ))
    }
}
impl<'c> Drop for Batch<'c> {
    fn drop(&mut self) {
        unsafe { sys::s21_batch_destroy(self.h) }
    }
}

/// `b` instances split over several GPUs of THIS process (`s21_sweep_*`: one host thread and stream per GPU inside the
/// library, contiguous blocks of `ceil(b / n_devices)` instances, results gathered into one pinned host buffer).
pub struct Sweep<'c> {
    h: *mut sys::s21_sweep,
    ckt: &'c Circuit,
    b: usize,
}
impl<'c> Sweep<'c> {
    /// `n_devices <= 0` takes every visible CUDA device.
    pub fn new(ckt: &'c Circuit, n_devices: i32, b: usize) -> SpResult<Self> {
        let mut h = ptr::null_mut();
        check(unsafe { sys::s21_sweep_create(ckt.h, n_devices, ptr::null(), b, &mut h) })?;
        Ok(Sweep { h, ckt, b })
    }
    pub fn num_devices(&self) -> usize {
        unsafe { sys::s21_sweep_num_devices(self.h) as usize }
    }
    /// `(first, count)` of shard `g` of `n_devices`: the partition rule, usable without a GPU.
    pub fn partition(b: usize, n_devices: i32, g: i32) -> SpResult<(usize, usize)> {
        let (mut first, mut count) = (0usize, 0usize);
        check(unsafe { sys::s21_sweep_partition(b, n_devices, g, &mut first, &mut count) })?;
        Ok((first, count))
    }
    pub fn set(&mut self, spec: &str, values: &[f64]) -> SpResult<()> {
        assert_eq!(values.len(), self.b);
        let s = CString::new(spec).unwrap();
        check(unsafe { sys::s21_sweep_override(self.h, s.as_ptr(), values.as_ptr()) })
    }
    pub fn reset(&mut self) -> SpResult<()> {
        check(unsafe { sys::s21_sweep_reset(self.h) })
    }
    pub fn dcop(&mut self) -> SpResult<DcopResult> {
        let n = self.ckt.num_vars();
        let mut r = DcopResult { x: vec![0.0; n * self.b], status: vec![0; self.b], iters: vec![0; self.b] };
        check(unsafe { sys::s21_sweep_dcop(self.h, r.x.as_mut_ptr(), r.status.as_mut_ptr(), r.iters.as_mut_ptr()) })?;
        Ok(r)
    }
    pub fn dcop_view(&mut self) -> SpResult<DcopView<'_>> {
        let n = self.ckt.num_vars();
        let (mut x, mut st, mut it) = (ptr::null(), ptr::null(), ptr::null());
        check(unsafe { sys::s21_sweep_dcop_view(self.h, &mut x, &mut st, &mut it) })?;
        Ok(unsafe {
            DcopView {
                x: std::slice::from_raw_parts(x, n * self.b),
                status: std::slice::from_raw_parts(st, self.b),
                iters: std::slice::from_raw_parts(it, self.b),
            }
        })
    }
    /// `Tran::solve` for every instance; `wave` is `[instance][timepoint][saved variable]`.
    pub fn tran(&mut self, tstep: f64, tstop: f64, save: &[i32]) -> SpResult<(Vec<f64>, Vec<f64>, Vec<i32>, Vec<i64>)> {
        let t = unsafe { sys::s21_tran_num_points(tstep, tstop) } as usize;
        let (mut time, mut wave) = (vec![0.0; t], vec![0.0; self.b * t * save.len()]);
        let (mut status, mut iters) = (vec![0i32; self.b], vec![0i64; self.b]);
        check(unsafe {
            sys::s21_sweep_tran(self.h, tstep, tstop, save.as_ptr(), save.len(), time.as_mut_ptr(), wave.as_mut_ptr(), status.as_mut_ptr(), iters.as_mut_ptr())
        })?;
        Ok((time, wave, status, iters))
    }
}
impl<'c> Drop for Sweep<'c> {
    fn drop(&mut self) {
        unsafe { sys::s21_sweep_destroy(self.h) }
    }
}
