// Regression tool for host/symbolic.hpp: prints size + FNV hash of EVERY vector of the symbolic Plan (permutations,
// L+U pattern, op lists, level schedules) for a COO matrix read from a binary file (int32 n, nnz, width; int32 rows[nnz];
// int32 cols[nnz]; f64 vals[nnz * width]). Build it twice — against the working tree and against an older symbolic.hpp
// (git show <rev>:spice21_b200/csrc/host/symbolic.hpp into a copy of csrc/) — and diff the outputs: that is how the
// round-1 rewrite of the Markowitz search (blocked column maxima, diagonal arrays, flat hash) was shown to leave the
// complete plan unchanged on the C3 matrix and 40 random real / complex / tie-heavy matrices.
//   g++ -O2 -std=c++17 -ffp-contract=off -Ispice21_b200/csrc -Iinclude -DSYMBOLIC_HPP='"host/symbolic.hpp"' -o plan_dump scripts/plan_dump.cpp
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>
#include "host/circuit.hpp"
#include SYMBOLIC_HPP
using namespace s21;
template <class V> static void dump(FILE* f, const char* name, const V& v) {
  std::fprintf(f, "%s %zu", name, v.size());
  unsigned long long h = 1469598103934665603ull;
  for (auto x : v) { h ^= (unsigned long long)(long long)x; h *= 1099511628211ull; }
  std::fprintf(f, " %016llx\n", h);
}
int main(int argc, char** argv) {
  FILE* in = std::fopen(argv[1], "rb");
  int hdr[3];
  if (std::fread(hdr, 4, 3, in) != 3) return 2;
  const int n = hdr[0], nnz = hdr[1], width = hdr[2];
  std::vector<int> r((size_t)nnz), c((size_t)nnz);
  std::vector<double> v((size_t)nnz * (size_t)width);
  if (std::fread(r.data(), 4, (size_t)nnz, in) != (size_t)nnz || std::fread(c.data(), 4, (size_t)nnz, in) != (size_t)nnz ||
      std::fread(v.data(), 8, v.size(), in) != v.size()) return 2;
  Plan P;
  if (width == 2) {
    std::vector<cplx> z((size_t)nnz);
    for (int k = 0; k < nnz; k++) z[(size_t)k] = mk(v[2 * (size_t)k], v[2 * (size_t)k + 1]);
    P = build_plan<cplx>(n, r, c, z.data());
  } else {
    P = build_plan<double>(n, r, c, v.data());
  }
  FILE* f = stdout;
  std::fprintf(f, "status %d N %d nnzA %d nnzLU %d n_stage %d\n", P.status, P.N, P.nnzA, P.nnzLU, P.n_stage);
#define D(x) dump(f, #x, P.x)
  D(row_i2e); D(row_e2i); D(col_i2e); D(col_e2i); D(rowptr); D(colidx); D(diag_slot); D(lu_row); D(lu_col); D(lu_fill); D(elem_slot);
  D(l_off); D(l_slot); D(l_row); D(upd_off); D(upd_t); D(upd_u); D(upd_l); D(lu_lvl_off); D(lu_t); D(lu_u); D(lu_l);
  D(fw_lvl_off); D(fw_k); D(fw_row); D(fw_slot); D(bw_lvl_off); D(bw_row);
  return 0;
}
