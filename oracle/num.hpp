// ORACLE — TEST INFRASTRUCTURE ONLY.
// CPU restatement of the reference's numeric base. Nothing in the product path
// (spice21_b200/) may include, link or call this. Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference leg use it, as the checker.
//
// Follows: spice21/src/spnum.rs:17-31 (Abs::absv: |x| for f64, norm() for Complex<f64>)
// and the `num` 0.3 crate (`num-complex`), which is NOT vendored in /root/reference
// (spice21/Cargo.toml:23 `num = "0.3.0"`, no Cargo.lock). Its published algorithm is restated:
//   (a+bi)*(c+di) = (ac-bd) + (ad+bc)i
//   (a+bi)/(c+di) = ((ac+bd) + (bc-ad)i) / (c^2+d^2)      (no scaling, no Smith's method)
//   norm()        = hypot(re, im)
//   z*s, z/s      = componentwise for real s
#pragma once
#include <cmath>
#include <stdexcept>
#include <string>

namespace orc {

struct SpError : std::runtime_error {   // spresult.rs:8-12  SpError{desc}
  explicit SpError(const std::string& d) : std::runtime_error(d) {}
};
struct Panic : std::runtime_error {     // Rust panic!/unwrap()/assert! sites
  explicit Panic(const std::string& d) : std::runtime_error(d) {}
};

struct Cplx {
  double re, im;
  Cplx() : re(0.0), im(0.0) {}
  Cplx(double r, double i) : re(r), im(i) {}
};
inline Cplx operator+(Cplx a, Cplx b) { return Cplx(a.re + b.re, a.im + b.im); }
inline Cplx operator-(Cplx a, Cplx b) { return Cplx(a.re - b.re, a.im - b.im); }
inline Cplx operator*(Cplx a, Cplx b) { return Cplx(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
inline Cplx operator/(Cplx a, Cplx b) {
  double n = b.re * b.re + b.im * b.im;
  double re = a.re * b.re + a.im * b.im;
  double im = a.im * b.re - a.re * b.im;
  return Cplx(re / n, im / n);
}
inline Cplx operator*(Cplx a, double s) { return Cplx(a.re * s, a.im * s); }
inline Cplx operator/(Cplx a, double s) { return Cplx(a.re / s, a.im / s); }
inline bool operator==(Cplx a, Cplx b) { return a.re == b.re && a.im == b.im; }
inline bool operator!=(Cplx a, Cplx b) { return !(a == b); }

// Abs::absv (spnum.rs:17-31)
inline double absv(double x) { return std::fabs(x); }
inline double absv(Cplx x) { return std::hypot(x.re, x.im); }

template <class T> inline T zero();
template <> inline double zero<double>() { return 0.0; }
template <> inline Cplx zero<Cplx>() { return Cplx(0.0, 0.0); }

}  // namespace orc
