// TEST INFRASTRUCTURE. Host-only instrumentation of the Bsim4 evaluation (Makefile: liboracle_count.so, -DS21_B4_COUNT;
// scripts/b4_opcount.py): how many divisions / exp / log / sqrt one evaluation executes, and how many of the divisions
// leave the range in which a reciprocal-based device division is valid (zero, subnormal, huge or non-finite operands /
// results). Never part of the product build.
#pragma once
#include <math.h>

namespace s21 {
namespace b4e {

struct B4Counts { unsigned long long evals, div, div_special, exp, log, sqrt; };
inline B4Counts& b4_counts() { static B4Counts c = {0, 0, 0, 0, 0, 0}; return c; }
inline double b4_count_div(double a, double b) {
  B4Counts& c = b4_counts();
  c.div++;
  const double q = a / b, ab = fabs(b), aq = fabs(q), aa = fabs(a);
  if (!(ab >= 1e-290 && ab <= 1e290) || !(aa <= 1e290) || (aa != 0.0 && aa < 1e-290) || !(aq <= 1e290) || (aq != 0.0 && aq < 1e-290)) c.div_special++;
  return q;
}
inline double exp(double x) { b4_counts().exp++; return ::exp(x); }
inline double log(double x) { b4_counts().log++; return ::log(x); }
inline double sqrt(double x) { b4_counts().sqrt++; return ::sqrt(x); }

}  // namespace b4e
}  // namespace s21
// the evaluation headers route every division through B4_DIV and call exp / log / sqrt unqualified inside s21::b4e
#define B4_DIV(a, b) ::s21::b4e::b4_count_div((double)(a), (double)(b))
