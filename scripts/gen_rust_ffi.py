"""Regenerates bindings/rust/spice21cu-sys/src/lib.rs from include/spice21cu.h (one `pub fn` per declared entry point), so
that the Rust side a maintainer links cannot drift from the C header; tests/test_host.py checks the sync.
usage: python scripts/gen_rust_ffi.py"""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = {"void": "c_void", "char": "c_char", "double": "f64", "float": "f32", "int32_t": "i32", "int64_t": "i64", "uint8_t": "u8",
        "uint64_t": "u64", "size_t": "usize", "s21_ckt": "s21_ckt", "s21_batch": "s21_batch", "s21_sweep": "s21_sweep",
        "s21_options": "s21_options"}


def rust_type(ctype):
    """C declarator type (no name) -> Rust. Handles const-qualified pointers of any depth."""
    toks = re.findall(r"const|\*|\w+", ctype)
    base, i, base_const = None, 0, False
    while i < len(toks) and toks[i] != "*":
        if toks[i] == "const":
            base_const = True
        else:
            base = toks[i]
        i += 1
    t, pointee_const = BASE[base], base_const
    while i < len(toks):
        assert toks[i] == "*"
        t = ("*const " if pointee_const else "*mut ") + t
        pointee_const = False
        i += 1
        if i < len(toks) and toks[i] == "const":  # `* const`: the pointer itself is const -> the next level's pointee is const
            pointee_const = True
            i += 1
    return t


def main():
    hdr = open(os.path.join(ROOT, "include", "spice21cu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    decls = re.findall(r"^\s*((?:const\s+)?\w+\s*\**)\s*(s21_\w+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.M | re.S)
    lines = []
    for ret, name, params in decls:
        ps = []
        params = " ".join(params.split())
        if params and params != "void":
            for p in params.split(","):
                m = re.match(r"^(.*?)(\w+)$", p.strip())
                ctype, pname = m.group(1).strip(), m.group(2)
                if pname in ("type", "fn", "ref", "in", "move"):
                    pname += "_"
                ps.append(f"{pname}: {rust_type(ctype)}")
        r = ret.strip()
        rr = "" if r == "void" else f" -> {rust_type(r)}"
        lines.append(f"    pub fn {name}({', '.join(ps)}){rr};")
    out = '''//! Raw FFI to `libspice21cu.so`. One declaration per entry point of `include/spice21cu.h` (generated from it by
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
''' + "\n".join(lines) + "\n}\n"
    path = os.path.join(ROOT, "bindings", "rust", "spice21cu-sys", "src", "lib.rs")
    open(path, "w").write(out)
    print(f"{len(lines)} entry points -> {path}")


if __name__ == "__main__":
    main()
