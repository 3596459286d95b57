// Symbolic phase of the sparse LU, run ONCE per analysis mode on the host with the values of the first Newton
// iteration. It reproduces the reference solver's decisions exactly —
//   spice21/src/sparse21/mod.rs:647-674  lu_factorize (N-1 pivot steps; singular checks)
//   spice21/src/sparse21/mod.rs:735-783  markowitz_search_diagonal (rel. threshold 1e-3, ties_mult 5, early exit)
//   spice21/src/sparse21/mod.rs:785-834  markowitz_search_submatrix (as written: only column n is examined)
//   spice21/src/sparse21/mod.rs:837-863  find_max
//   spice21/src/sparse21/mod.rs:865-919  row_col_elim (fill-in creation, Markowitz count maintenance)
// — but works on permutation vectors over externally-indexed rows/columns instead of physically swapping
// orthogonal linked lists. "First in list order" tie-breaks of the reference become "smallest current internal
// index". The product is a static plan: pivot order, L+U pattern with fill, and flat op lists for the numeric
// kernels (refactorisation, forward/back substitution, residual SpMV), all in final internal coordinates.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#include "../scalar.h"
#include "circuit.hpp"

namespace s21 {

struct Plan {
  int status = ST_OK;  // ST_SINGULAR / ST_PIVOT when the reference would have returned that SpError on this matrix
  bool over_budget = false;  // build_plan gave up at its update budget (status ST_PIVOT): not the reference's verdict
  int N = 0, nnzA = 0, nnzLU = 0;
  std::vector<int> row_i2e, row_e2i, col_i2e, col_e2i;
  // L+U pattern: CSR over internal rows, ascending internal column. Slot = position in this order.
  std::vector<int> rowptr, colidx, diag_slot;
  std::vector<int> lu_row, lu_col, lu_fill;  // per slot (export / tests)
  std::vector<int> elem_slot;                // original element id -> slot
  // Numeric refactorisation, pivot k = 0..N-2: L_k = slots below the diagonal in column k (with their rows),
  // and one (target, u, l) triple per Schur update. Offsets have N entries + 1.
  std::vector<int> l_off, l_slot, l_row;
  // [N] 1 when pivot k was chosen by the reference's diagonal search, i.e. passed its 1e-3 threshold against its column; the
  // fallback searches (mod.rs:785-863) apply no threshold, so the kernels' pivot-health test skips those pivots.
  std::vector<int> piv_checked;
  std::vector<int> upd_off, upd_t, upd_u, upd_l;
  // Level schedule for the cooperative kernel: the same operations, grouped so that every operation of a level is
  // independent of the others in it (one barrier per level). Within any single value the operations keep the
  // reference's order (ascending pivot index), so the arithmetic is unchanged.
  //   LU op j: l < 0 ? lu[t] /= lu[u] : lu[t] -= lu[u] * lu[l]
  std::vector<int> lu_lvl_off, lu_t, lu_u, lu_l;
  //   forward op j: c[row] -= c[k] * lu[slot]   (skipped when c[k] == 0, sparse21/mod.rs:949-951)
  std::vector<int> fw_lvl_off, fw_k, fw_row, fw_slot;
  //   backward level: rows whose c[k] = (c[k] - sum_j U[k,j] c[j]) / U[k,k] can be formed together
  std::vector<int> bw_lvl_off, bw_row;
  // Staged assembly (filled per analysis mode by the batch runtime): target t in [0, nnzLU) is an L+U slot, target
  // nnzLU + v is rhs[v]; asm_src lists the staging slots to sum, in the reference's accumulation order.
  std::vector<int> asm_off, asm_src;
  int n_stage = 0;
  // Level schedules built in "tolerance" mode (build_levels): Schur updates / forward-substitution updates that hit the
  // same target no longer wait for each other, so several may sit in one level and the kernel must apply them with
  // atomic adds (their order, hence the last bits of the sum, is then not fixed).
  bool relaxed = false;
};

// Group operations by dependency level. `lev[j]` >= 1 for every op; returns offsets of the stable level sort.
inline std::vector<int> level_sort(const std::vector<int>& lev, std::vector<int>* order) {
  int nl = 0;
  for (int l : lev) nl = std::max(nl, l);
  std::vector<int> cnt((size_t)nl + 2, 0);  // after the prefix sum: cnt[l] = number of ops with level < l
  for (int l : lev) cnt[(size_t)l + 1] += 1;
  for (int l = 1; l <= nl + 1; l++) cnt[(size_t)l] += cnt[(size_t)l - 1];
  order->assign(lev.size(), 0);
  std::vector<int> pos(cnt.begin(), cnt.end());
  for (size_t j = 0; j < lev.size(); j++) (*order)[(size_t)pos[(size_t)lev[j]]++] = (int)j;
  std::vector<int> offs;
  for (int l = 1; l <= nl + 1; l++) offs.push_back(cnt[(size_t)l]);
  return offs;  // offs[q] = first op of the q-th level (q = 0 .. nl), offs[nl] = total
}

// relaxed = false: every value sees its updates in the reference's order (ascending pivot index), so results are
// bit-identical to the one-thread kernel — but all updates into one target form a chain. On config C3 (2000 rings on one
// supply node) the 10 000 contributions to the supply node's diagonal serialise the whole factorisation: 10 k LU levels
// and 12 k forward levels, one grid barrier each. relaxed = true ("tolerance" mode, SPICE reltol is 1e-3): an update
// waits only for its two factors; updates into one target commute, land in the same level and are applied atomically.
// The level count falls to the depth of the elimination DAG.
inline void build_levels(Plan& P, bool relaxed = false) {
  const int N = P.N;
  P.relaxed = relaxed;
  // ---- LU
  {
    std::vector<int> ready((size_t)P.nnzLU, 0), t, u, l, lev;
    for (int k = 0; k + 1 < N; k++) {
      const int piv = P.diag_slot[(size_t)k];
      if (piv < 0) continue;
      for (int j = P.l_off[(size_t)k]; j < P.l_off[(size_t)k + 1]; j++) {
        const int ls = P.l_slot[(size_t)j];
        const int lv = std::max(ready[(size_t)ls], ready[(size_t)piv]) + 1;
        t.push_back(ls); u.push_back(piv); l.push_back(P.piv_checked[(size_t)k] ? -1 : -2); lev.push_back(lv);  // -1: division by a threshold-checked pivot, -2: unchecked
        ready[(size_t)ls] = lv;
      }
      for (int j = P.upd_off[(size_t)k]; j < P.upd_off[(size_t)k + 1]; j++) {
        const int tt = P.upd_t[(size_t)j], uu = P.upd_u[(size_t)j], ll = P.upd_l[(size_t)j];
        const int lv = std::max(relaxed ? 0 : ready[(size_t)tt], std::max(ready[(size_t)uu], ready[(size_t)ll])) + 1;
        t.push_back(tt); u.push_back(uu); l.push_back(ll); lev.push_back(lv);
        ready[(size_t)tt] = std::max(ready[(size_t)tt], lv);
      }
    }
    std::vector<int> order;
    P.lu_lvl_off = level_sort(lev, &order);
    if (relaxed)  // updates of one level that share a target sit next to each other: the kernel sums each run inside a warp
      for (size_t q = 0; q + 1 < P.lu_lvl_off.size(); q++)  // and issues ONE atomic per run (grid.cu) instead of a chain on one address
        std::stable_sort(order.begin() + P.lu_lvl_off[q], order.begin() + P.lu_lvl_off[q + 1],
                         [&](int a, int b) { return t[(size_t)a] < t[(size_t)b]; });
    for (int j : order) { P.lu_t.push_back(t[(size_t)j]); P.lu_u.push_back(u[(size_t)j]); P.lu_l.push_back(l[(size_t)j]); }
  }
  // ---- forward substitution
  {
    std::vector<int> ready((size_t)N, 0), kk, rr, ss, lev;
    for (int k = 0; k < N; k++)
      for (int j = P.l_off[(size_t)k]; j < P.l_off[(size_t)k + 1]; j++) {
        const int row = P.l_row[(size_t)j];
        const int lv = std::max(ready[(size_t)k], relaxed ? 0 : ready[(size_t)row]) + 1;
        kk.push_back(k); rr.push_back(row); ss.push_back(P.l_slot[(size_t)j]); lev.push_back(lv);
        ready[(size_t)row] = std::max(ready[(size_t)row], lv);
      }
    std::vector<int> order;
    P.fw_lvl_off = level_sort(lev, &order);
    for (int j : order) { P.fw_k.push_back(kk[(size_t)j]); P.fw_row.push_back(rr[(size_t)j]); P.fw_slot.push_back(ss[(size_t)j]); }
  }
  // ---- backward substitution
  {
    std::vector<int> lvl((size_t)N, 1), rows, lev;
    for (int k = N - 1; k >= 0; k--) {
      const int ds = P.diag_slot[(size_t)k];
      if (ds < 0) continue;
      int lv = 1;
      for (int s = ds + 1; s < P.rowptr[(size_t)k + 1]; s++) lv = std::max(lv, lvl[(size_t)P.colidx[(size_t)s]] + 1);
      lvl[(size_t)k] = lv;
      rows.push_back(k); lev.push_back(lv);
    }
    std::vector<int> order;
    P.bw_lvl_off = level_sort(lev, &order);
    for (int j : order) P.bw_row.push_back(rows[(size_t)j]);
  }
}

namespace detail {
template <class T>
struct SymEntry {
  int r, c;  // external coordinates
  T val;
  bool fill;
};
inline uint64_t rc_key(int r, int c) { return ((uint64_t)(uint32_t)r << 32) | (uint32_t)c; }
// (row, col) -> entry id: open addressing with linear probing (the elimination step looks up one coordinate per Schur
// update and inserts one per fill-in; std::unordered_map spent most of that step in allocation and pointer chasing).
class CoordMap {
 public:
  explicit CoordMap(size_t expect) {
    size_t cap = 64;
    while (cap < expect * 2) cap <<= 1;
    keys_.assign(cap, kEmpty);
    vals_.assign(cap, -1);
  }
  int find(uint64_t key) const {
    const size_t mask = keys_.size() - 1;
    for (size_t h = hash(key) & mask;; h = (h + 1) & mask) {
      if (keys_[h] == key) return vals_[h];
      if (keys_[h] == kEmpty) return -1;
    }
  }
  void insert(uint64_t key, int val) {  // key must not be present
    if ((n_ + 1) * 2 > keys_.size()) grow();
    put(key, val);
    n_++;
  }

 private:
  static constexpr uint64_t kEmpty = ~0ull;  // (row, col) = (-1, -1) never occurs
  static size_t hash(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33;
    return (size_t)k;
  }
  void put(uint64_t key, int val) {
    const size_t mask = keys_.size() - 1;
    size_t h = hash(key) & mask;
    while (keys_[h] != kEmpty) h = (h + 1) & mask;
    keys_[h] = key;
    vals_[h] = val;
  }
  void grow() {
    std::vector<uint64_t> ok;
    std::vector<int> ov;
    ok.swap(keys_);
    ov.swap(vals_);
    keys_.assign(ok.size() * 2, kEmpty);
    vals_.assign(ok.size() * 2, -1);
    for (size_t i = 0; i < ok.size(); i++)
      if (ok[i] != kEmpty) put(ok[i], ov[i]);
  }
  std::vector<uint64_t> keys_;
  std::vector<int> vals_;
  size_t n_ = 0;
};
}  // namespace detail

// vals[e] = assembled value of element e after the first device-load sweep.
// max_updates > 0: give up (status ST_PIVOT) once the elimination has applied that many Schur updates — used for re-pivot
// attempts at an intermediate Newton iterate, where the reference's value-driven Markowitz order can pick a dense row early
// (a supply node with thousands of devices) and fill the whole matrix: measured on 400 rings (N = 2803), 1.15 G updates and
// 2.4 M L+U entries against 20 k / 15 k for the order taken at x = 0, i.e. 143 s of host time for one Newton step.
template <class T>
Plan build_plan(int N, const std::vector<int>& elem_row, const std::vector<int>& elem_col, const T* vals, bool relaxed = false, size_t max_updates = 0) {
  using detail::SymEntry;
  using detail::rc_key;
  Plan P;
  P.N = N;
  P.nnzA = (int)elem_row.size();
  std::vector<SymEntry<T>> E;
  E.reserve(elem_row.size() * 2);
  std::vector<std::vector<int>> in_row((size_t)N), in_col((size_t)N);
  detail::CoordMap at(elem_row.size() * 2);
  for (size_t e = 0; e < elem_row.size(); e++) {
    E.push_back({elem_row[e], elem_col[e], vals[e], false});
    in_row[(size_t)elem_row[e]].push_back((int)e);
    in_col[(size_t)elem_col[e]].push_back((int)e);
    if (at.find(rc_key(elem_row[e], elem_col[e])) < 0) at.insert(rc_key(elem_row[e], elem_col[e]), (int)e);
  }
  P.row_i2e.resize((size_t)N); P.row_e2i.resize((size_t)N); P.col_i2e.resize((size_t)N); P.col_e2i.resize((size_t)N);
  for (int k = 0; k < N; k++) P.row_i2e[(size_t)k] = P.row_e2i[(size_t)k] = P.col_i2e[(size_t)k] = P.col_e2i[(size_t)k] = k;
  auto lookup = [&](int r, int c) { return at.find(rc_key(r, c)); };
  // diag[k] = id of the entry at internal (k, k), or -1 (the reference keeps the same array, mod.rs:225). Maintained under
  // the two swaps of a step and under fill-in creation, so the diagonal search reads an array instead of hashing N
  // coordinates per pivot (the search was 2/3 of the symbolic phase of config C3).
  // dabs[k] = |value| of that entry, kept current by the elimination step: the search then touches only small arrays
  // (the entry table is tens of MB for C3 and was read at random for every candidate).
  std::vector<int> diag((size_t)N, -1);
  std::vector<double> dabs((size_t)N, 0.0);
  auto set_diag = [&](int k, int id) { diag[(size_t)k] = id; dabs[(size_t)k] = id < 0 ? 0.0 : s_abs(E[(size_t)id].val); };
  for (int k = 0; k < N; k++) set_diag(k, lookup(k, k));

  // mod.rs:649-658 — an empty row or column is "Singular Matrix"
  for (int k = 0; k < N; k++)
    if (in_row[(size_t)k].empty() || in_col[(size_t)k].empty()) { P.status = ST_SINGULAR; }
  std::vector<long> mrow((size_t)N), mcol((size_t)N);  // Markowitz counts, keyed by external row / column
  for (int k = 0; k < N; k++) { mrow[(size_t)k] = (long)in_row[(size_t)k].size(); mcol[(size_t)k] = (long)in_col[(size_t)k].size(); }

  struct Step { std::vector<int> L, U; };  // entry ids, in the order the reference walks them
  std::vector<Step> steps((size_t)std::max(N - 1, 0));

  auto swap_int = [](std::vector<int>& i2e, std::vector<int>& e2i, int x, int y) {
    std::swap(i2e[(size_t)x], i2e[(size_t)y]);
    e2i[(size_t)i2e[(size_t)x]] = x;
    e2i[(size_t)i2e[(size_t)y]] = y;
  };
  // max |val| among the entries of external column c whose current internal row is >= n; ties -> smallest internal row
  // (max_after_loc walks the column in ascending internal row and replaces its best only on a strictly larger value).
  // Cached per column, and inside a column per block of BS list positions: an elimination step marks exactly the blocks
  // whose answer can have changed — entries it updated or created, the entries of the pivot row (they leave the active
  // part) and of the row that the row swap moved (its internal index, hence its rank in a tie, changed). A long column
  // that every step touches in a few places (the shared supply node of config C3: ~N entries, N steps) then costs a few
  // blocks plus one pass over the block winners per step instead of a full walk — that walk was most of the symbolic
  // phase of C3 after the diagonal search itself had been reduced to array reads.
  constexpr int BS = 16;
  size_t n_recomp = 0, n_visit = 0;  // statistics for S21_PLAN_INFO
  std::vector<int> epos(E.size());  // position of an entry in its column's list
  {
    std::vector<int> fillc((size_t)N, 0);
    for (size_t e = 0; e < E.size(); e++) epos[e] = fillc[(size_t)E[e].c]++;
  }
  std::vector<std::vector<int>> blk_best((size_t)N), blk_ir((size_t)N);  // winner of a block, its internal row and |value|
  std::vector<std::vector<double>> blk_val((size_t)N);
  std::vector<std::vector<char>> blk_dirty((size_t)N);
  for (int c = 0; c < N; c++) {
    const size_t nb = (in_col[(size_t)c].size() + BS - 1) / BS;
    blk_best[(size_t)c].assign(nb, -1);
    blk_ir[(size_t)c].assign(nb, 0);
    blk_val[(size_t)c].assign(nb, 0.0);
    blk_dirty[(size_t)c].assign(nb, 1);
  }
  std::vector<int> cmax_id((size_t)N, -2);  // -2 = not cached, -1 = no active entry
  std::vector<double> cmax_abs((size_t)N, 0.0);
  auto touch = [&](int id) {
    const int c = E[(size_t)id].c;
    blk_dirty[(size_t)c][(size_t)(epos[(size_t)id] / BS)] = 1;
    cmax_id[(size_t)c] = -2;
  };
  auto col_max_from = [&](int c, int n) {
    if (cmax_id[(size_t)c] != -2) return cmax_id[(size_t)c];
    n_recomp++;
    const std::vector<int>& lst = in_col[(size_t)c];
    int best = -1, best_ir = 0;
    double best_val = 0.0;
    for (size_t blk = 0; blk < blk_best[(size_t)c].size(); blk++) {
      int bb = -1, bb_ir = 0;
      double bb_val = 0.0;
      if (blk_dirty[(size_t)c][blk]) {
        const size_t hi = std::min(lst.size(), (blk + 1) * (size_t)BS);
        for (size_t p = blk * (size_t)BS; p < hi; p++) {
          n_visit++;
          const int id = lst[p];
          const int ir = P.row_e2i[(size_t)E[(size_t)id].r];
          if (ir < n) continue;
          const double a = s_abs(E[(size_t)id].val);
          if (bb < 0 || a > bb_val || (a == bb_val && ir < bb_ir)) { bb = id; bb_val = a; bb_ir = ir; }
        }
        blk_best[(size_t)c][blk] = bb;
        blk_ir[(size_t)c][blk] = bb_ir;
        blk_val[(size_t)c][blk] = bb_val;
        blk_dirty[(size_t)c][blk] = 0;
      } else {
        bb = blk_best[(size_t)c][blk];
        bb_ir = blk_ir[(size_t)c][blk];
        bb_val = blk_val[(size_t)c][blk];
      }
      if (bb >= 0 && (best < 0 || bb_val > best_val || (bb_val == best_val && bb_ir < best_ir))) { best = bb; best_val = bb_val; best_ir = bb_ir; }
    }
    cmax_id[(size_t)c] = best;
    cmax_abs[(size_t)c] = best_val;
    return best;
  };

  const bool info = std::getenv("S21_PLAN_INFO") != nullptr;
  using clk = std::chrono::steady_clock;
  double t_search = 0.0, t_elim = 0.0;
  size_t n_upd = 0, n_cand = 0;
  const auto t_begin = clk::now();
  P.piv_checked.assign((size_t)std::max(N, 1), 0);
  for (int n = 0; n + 1 < N && P.status == ST_OK; n++) {
    int pivot = -1;
    const auto t_s0 = clk::now();
    {  // ---- markowitz_search_diagonal
      long best_mark = -1;  // -1 stands for usize::MAX
      double best_ratio = 0.0;
      long num_ties = 0;
      bool done = false;
      for (int k = n; k < N && !done; k++) {
        n_cand++;
        int d = diag[(size_t)k];
        if (d < 0) continue;
        const int ec = P.col_i2e[(size_t)k];
        int mx = col_max_from(ec, n);
        if (mx < 0) continue;
        const double mabs = cmax_abs[(size_t)ec], da = dabs[(size_t)k];
        double threshold = 1e-3 * mabs + 0.0;
        if (da < threshold) continue;
        long mr = mrow[(size_t)P.row_i2e[(size_t)k]], mc = mcol[(size_t)ec];
        if (!(mr > 0 && mc > 0)) throw S21Error(ST_OTHER, "markowitz count underflow");
        long mark = (mr - 1) * (mc - 1);
        // |d / m|: for real values the quotient of the magnitudes, bit for bit; complex values take the num-style quotient
        auto ratio_of = [&]() {
          if (Scalar<T>::width == 1) return s_div(da, mabs);
          return s_abs(s_div(E[(size_t)d].val, E[(size_t)mx].val));
        };
        if (best_mark < 0 || mark < best_mark) {
          num_ties = 0;
          pivot = d;
          best_mark = mark;
          best_ratio = ratio_of();
        } else if (mark == best_mark) {
          num_ties += 1;
          double ratio = ratio_of();
          if (ratio > best_ratio) { pivot = d; best_ratio = ratio; }
          if (num_ties >= best_mark * 5) done = true;
        }
      }
    }
    P.piv_checked[(size_t)n] = pivot >= 0 ? 1 : 0;
    if (pivot < 0) {  // ---- markowitz_search_submatrix: column n only
      std::vector<std::pair<int, int>> cand;  // (internal row, id)
      for (int id : in_col[(size_t)P.col_i2e[(size_t)n]]) {
        int ir = P.row_e2i[(size_t)E[(size_t)id].r];
        if (ir >= n) cand.push_back({ir, id});
      }
      std::sort(cand.begin(), cand.end());
      if (!cand.empty()) {
        int mx = cand[0].second;
        double mxv = s_abs(E[(size_t)mx].val);
        for (auto& c : cand) { double a = s_abs(E[(size_t)c.second].val); if (a > mxv) { mx = c.second; mxv = a; } }
        long best_mark = -1;
        double best_ratio = 0.0;
        for (auto& c : cand) {
          int id = c.second;
          long mr = mrow[(size_t)E[(size_t)id].r], mc = mcol[(size_t)E[(size_t)id].c];
          if (!(mr > 0 && mc > 0)) throw S21Error(ST_OTHER, "markowitz count underflow");
          long mark = (mr - 1) * (mc - 1);
          double ratio = s_abs(s_div(E[(size_t)id].val, E[(size_t)mx].val));
          if (best_mark < 0 || mark < best_mark) { pivot = id; best_mark = mark; best_ratio = ratio; }
          else if (mark == best_mark && ratio > best_ratio) { pivot = id; best_ratio = ratio; }
        }
      }
    }
    if (pivot < 0) {  // ---- find_max over the active submatrix, column-major in internal order
      double max_val = 0.0;
      for (int k = n; k < N; k++) {
        std::vector<std::pair<int, int>> cand;
        for (int id : in_col[(size_t)P.col_i2e[(size_t)k]]) {
          int ir = P.row_e2i[(size_t)E[(size_t)id].r];
          if (ir >= n) cand.push_back({ir, id});
        }
        std::sort(cand.begin(), cand.end());
        for (auto& c : cand) { double a = s_abs(E[(size_t)c.second].val); if (a > max_val) { pivot = c.second; max_val = a; } }
      }
    }
    if (pivot < 0) { P.status = ST_PIVOT; break; }
    const auto t_s1 = clk::now();
    t_search += std::chrono::duration<double>(t_s1 - t_s0).count();

    // ---- swap the pivot to (n, n)
    {
      const int xr = P.row_e2i[(size_t)E[(size_t)pivot].r], xc = P.col_e2i[(size_t)E[(size_t)pivot].c];
      for (int id : in_row[(size_t)P.row_i2e[(size_t)xr]]) touch(id);  // the pivot row leaves the active part
      if (xr != n)  // the row at n moves to xr: it held the smallest active index, so it won every tie it was part of —
        for (int id : in_row[(size_t)P.row_i2e[(size_t)n]]) {  // only where it is the recorded winner can the answer change
          const int c = E[(size_t)id].c;
          if (blk_best[(size_t)c][(size_t)(epos[(size_t)id] / BS)] == id || cmax_id[(size_t)c] == id) touch(id);
        }
      swap_int(P.row_i2e, P.row_e2i, xr, n);
      swap_int(P.col_i2e, P.col_e2i, xc, n);
      for (int k : {xr, xc, n}) set_diag(k, lookup(P.row_i2e[(size_t)k], P.col_i2e[(size_t)k]));
    }

    // ---- row_col_elim
    const int pr = E[(size_t)pivot].r, pc = E[(size_t)pivot].c;
    T pivot_val = E[(size_t)pivot].val;
    if (s_is_zero(pivot_val)) { P.status = ST_SINGULAR; break; }
    Step& st = steps[(size_t)n];
    {
      std::vector<std::pair<int, int>> ls, us;
      for (int id : in_col[(size_t)pc]) { int ir = P.row_e2i[(size_t)E[(size_t)id].r]; if (ir > n) ls.push_back({ir, id}); }
      for (int id : in_row[(size_t)pr]) { int ic = P.col_e2i[(size_t)E[(size_t)id].c]; if (ic > n) us.push_back({ic, id}); }
      std::sort(ls.begin(), ls.end());
      std::sort(us.begin(), us.end());
      for (auto& x : ls) st.L.push_back(x.second);
      for (auto& x : us) st.U.push_back(x.second);
    }
    for (int l : st.L) E[(size_t)l].val = s_div(E[(size_t)l].val, pivot_val);
    for (int u : st.U) {
      const int uc = E[(size_t)u].c;
      for (int l : st.L) {
        const int lr = E[(size_t)l].r;
        int t = lookup(lr, uc);
        if (t < 0) {  // fill-in
          t = (int)E.size();
          E.push_back({lr, uc, Scalar<T>::zero(), true});
          in_row[(size_t)lr].push_back(t);
          epos.push_back((int)in_col[(size_t)uc].size());
          in_col[(size_t)uc].push_back(t);
          if (blk_best[(size_t)uc].size() * (size_t)BS < in_col[(size_t)uc].size()) {
            blk_best[(size_t)uc].push_back(-1); blk_ir[(size_t)uc].push_back(0); blk_val[(size_t)uc].push_back(0.0); blk_dirty[(size_t)uc].push_back(1);
          }
          at.insert(rc_key(lr, uc), t);
          if (P.row_e2i[(size_t)lr] == P.col_e2i[(size_t)uc]) diag[(size_t)P.row_e2i[(size_t)lr]] = t;
          mrow[(size_t)lr] += 1;
          mcol[(size_t)uc] += 1;
        }
        E[(size_t)t].val = s_sub(E[(size_t)t].val, s_mul(E[(size_t)u].val, E[(size_t)l].val));
        touch(t);
        if (P.row_e2i[(size_t)lr] == P.col_e2i[(size_t)uc]) dabs[(size_t)P.row_e2i[(size_t)lr]] = s_abs(E[(size_t)t].val);
      }
      mcol[(size_t)uc] -= 1;
    }
    mrow[(size_t)pr] -= 1;
    mcol[(size_t)pc] -= 1;
    for (int l : st.L) mrow[(size_t)E[(size_t)l].r] -= 1;
    n_upd += st.L.size() * st.U.size();
    if (max_updates && n_upd > max_updates) { P.status = ST_PIVOT; P.over_budget = true; break; }
    t_elim += std::chrono::duration<double>(clk::now() - t_s1).count();
  }
  const auto t_fact = clk::now();

  // ---- freeze: slots in final internal coordinates
  P.nnzLU = (int)E.size();
  auto irow = [&](int id) { return P.row_e2i[(size_t)E[(size_t)id].r]; };
  auto icol = [&](int id) { return P.col_e2i[(size_t)E[(size_t)id].c]; };
  std::vector<int> order((size_t)P.nnzLU);
  {  // sort by (internal row, internal column): the keys are materialised once, the sort then runs on contiguous memory
    std::vector<std::pair<uint64_t, int>> keyed((size_t)P.nnzLU);
    for (int k = 0; k < P.nnzLU; k++) keyed[(size_t)k] = {detail::rc_key(irow(k), icol(k)), k};
    std::sort(keyed.begin(), keyed.end());
    for (int k = 0; k < P.nnzLU; k++) order[(size_t)k] = keyed[(size_t)k].second;
  }
  std::vector<int> slot_of((size_t)P.nnzLU);
  P.rowptr.assign((size_t)N + 1, 0);
  P.colidx.resize((size_t)P.nnzLU);
  P.lu_row.resize((size_t)P.nnzLU); P.lu_col.resize((size_t)P.nnzLU); P.lu_fill.resize((size_t)P.nnzLU);
  P.diag_slot.assign((size_t)N, -1);
  for (int s = 0; s < P.nnzLU; s++) {
    int id = order[(size_t)s];
    slot_of[(size_t)id] = s;
    int r = irow(id), c = icol(id);
    P.rowptr[(size_t)r + 1] += 1;
    P.colidx[(size_t)s] = c;
    P.lu_row[(size_t)s] = r; P.lu_col[(size_t)s] = c; P.lu_fill[(size_t)s] = E[(size_t)id].fill ? 1 : 0;
    if (r == c) P.diag_slot[(size_t)r] = s;
  }
  for (int r = 0; r < N; r++) P.rowptr[(size_t)r + 1] += P.rowptr[(size_t)r];
  P.elem_slot.resize(elem_row.size());
  for (size_t e = 0; e < elem_row.size(); e++) P.elem_slot[e] = slot_of[e];
  if (P.status == ST_OK)
    for (int k = 0; k < N; k++)
      if (P.diag_slot[(size_t)k] < 0) P.status = ST_SINGULAR;  // mod.rs:955-958, 969-972

  P.l_off.assign((size_t)N + 1, 0);
  P.upd_off.assign((size_t)N + 1, 0);
  for (int n = 0; n < N; n++) {
    if (n + 1 < N && P.status == ST_OK) {
      const Step& st = steps[(size_t)n];
      // L in ascending FINAL internal row (the order forward substitution and the kernels walk)
      std::vector<std::pair<int, int>> ls;
      for (int l : st.L) ls.push_back({irow(l), slot_of[(size_t)l]});
      std::sort(ls.begin(), ls.end());
      for (auto& x : ls) { P.l_row.push_back(x.first); P.l_slot.push_back(x.second); }
      for (int u : st.U)
        for (int l : st.L) {
          int t = lookup(E[(size_t)l].r, E[(size_t)u].c);
          P.upd_t.push_back(slot_of[(size_t)t]);
          P.upd_u.push_back(slot_of[(size_t)u]);
          P.upd_l.push_back(slot_of[(size_t)l]);
        }
    }
    P.l_off[(size_t)n + 1] = (int)P.l_slot.size();
    P.upd_off[(size_t)n + 1] = (int)P.upd_t.size();
  }
  const auto t_frozen = clk::now();
  if (P.status == ST_OK) build_levels(P, relaxed);
  if (info)
    std::fprintf(stderr, "[s21 symbolic] N=%d nnzLU=%d candidates=%zu updates=%zu | search %.2f s, elimination %.2f s, freeze+op lists %.2f s, levels %.2f s\n", N,
                 P.nnzLU, n_cand, n_upd, t_search, t_elim, std::chrono::duration<double>(t_frozen - t_fact).count(),
                 std::chrono::duration<double>(clk::now() - t_frozen).count());
  if (info) std::fprintf(stderr, "[s21 symbolic] column maxima recomputed %zu times, %zu entries visited\n", n_recomp, n_visit);
  (void)t_begin;
  return P;
}

}  // namespace s21
