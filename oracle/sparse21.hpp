// ORACLE — TEST INFRASTRUCTURE ONLY (see num.hpp header).
//
// CPU restatement of the reference's sparse matrix + Markowitz LU:
//   spice21/src/sparse21/mod.rs:69-1012  (Element, AxisMapping, AxisData, Matrix<T>)
// Orthogonal linked lists, element handles (Eindex = creation order), in-place right-looking LU with
// Markowitz pivoting re-run on every factorisation, persistent row/col mappings, fill-ins kept forever.
// Every function below cites the reference lines it follows. `-1` plays the role of `None`.
#pragma once
#include <cstddef>
#include <cstdint>
#include <utility>
#include <vector>

#include "num.hpp"

namespace orc {

enum Axis { ROWS = 0, COLS = 1 };
inline Axis other(Axis a) { return a == ROWS ? COLS : ROWS; }

enum class MatrixState { CREATED, FACTORING, FACTORED, RESET };  // mod.rs:61-66

template <class T>
struct Element {  // mod.rs:69-78
  int index;
  size_t row, col;
  T val;
  bool fillin;
  size_t orig_row, orig_col;
  int next_in_row = -1, next_in_col = -1;
  size_t loc(Axis ax) const { return ax == ROWS ? row : col; }
  void set_loc(Axis ax, size_t to) { (ax == ROWS ? row : col) = to; }
  int next(Axis ax) const { return ax == ROWS ? next_in_row : next_in_col; }
  void set_next(Axis ax, int e) { (ax == ROWS ? next_in_row : next_in_col) = e; }
};

struct AxisMapping {  // mod.rs:126-162
  std::vector<size_t> e2i, i2e;
  std::vector<std::pair<size_t, size_t>> history;
  explicit AxisMapping(size_t size) {
    for (size_t k = 0; k < size; k++) { e2i.push_back(k); i2e.push_back(k); }
  }
  void swap_int(size_t x, size_t y) {  // :153-161
    size_t tmp = i2e[x];
    i2e[x] = i2e[y];
    i2e[y] = tmp;
    e2i[i2e[x]] = x;
    e2i[i2e[y]] = y;
    history.push_back({x, y});
  }
};

struct AxisData {  // mod.rs:164-205
  std::vector<int> hdrs;
  std::vector<size_t> qtys, markowitz;
  bool has_mapping = false;
  AxisMapping mapping{0};
  void grow(size_t to) {  // :180-190
    if (to <= hdrs.size()) return;
    size_t by = to - hdrs.size();
    for (size_t k = 0; k < by; k++) { hdrs.push_back(-1); qtys.push_back(0); markowitz.push_back(0); }
  }
  void setup_factoring() {  // :191-196
    markowitz = qtys;
    if (!has_mapping) { mapping = AxisMapping(hdrs.size()); has_mapping = true; }
  }
  void swap(size_t x, size_t y) {  // :197-204
    std::swap(hdrs[x], hdrs[y]);
    std::swap(qtys[x], qtys[y]);
    std::swap(markowitz[x], markowitz[y]);
    if (has_mapping) mapping.swap_int(x, y);
  }
};

// mod.rs:213-217
struct MarkowitzConfig { double rel_threshold, abs_threshold; size_t ties_mult; };
static const MarkowitzConfig MARKOWITZ_CONFIG = {1e-3, 0.0, 5};

template <class T>
struct Matrix {  // mod.rs:220-228
  MatrixState state = MatrixState::CREATED;
  std::vector<Element<T>> elements;
  AxisData axes[2];
  std::vector<int> diag;
  std::vector<int> fillins;
  // instrumentation (not in the reference): counters for the CPU baseline report
  uint64_t n_factorizations = 0;

  Element<T>& el(int i) { return elements[(size_t)i]; }
  const Element<T>& el(int i) const { return elements[(size_t)i]; }
  int hdr(Axis ax, size_t loc) const { return axes[ax].hdrs[loc]; }   // :1000
  void set_hdr(Axis ax, size_t loc, int ei) { axes[ax].hdrs[loc] = ei; }
  size_t num_rows() const { return axes[ROWS].hdrs.size(); }
  size_t num_cols() const { return axes[COLS].hdrs.size(); }

  static Matrix from_entries(const std::vector<std::tuple<size_t, size_t, T>>& entries) {  // :245-251
    Matrix m;
    for (auto& e : entries) m.add_element(std::get<0>(e), std::get<1>(e), std::get<2>(e));
    return m;
  }
  static Matrix identity(size_t n) {  // :1016-1022
    Matrix m;
    for (size_t k = 0; k < n; k++) m.add_element(k, k, T(1.0));
    return m;
  }
  void add_element(size_t row, size_t col, T val) { _add_element(row, col, val, false); }  // :253
  int make(size_t row, size_t col) {  // :265-270
    int ei = get_elem(row, col);
    if (ei >= 0) return ei;
    return _add_element(row, col, zero<T>(), false);
  }
  void reset() {  // :272-277
    for (auto& e : elements) e.val = zero<T>();
    state = MatrixState::RESET;
  }
  void update(int ei, T val) {  // :279-282
    T tmp = el(ei).val + val;
    el(ei).val = tmp;
  }
  std::vector<T> vecmul(const std::vector<T>& x) const {  // :284-297
    if (x.size() != num_cols()) throw SpError("Invalid Dimensions");
    std::vector<T> y(num_rows(), zero<T>());
    for (size_t row = 0; row < num_rows(); row++) {
      int ep = hdr(ROWS, row);
      while (ep >= 0) {
        y[row] = y[row] + el(ep).val * x[el(ep).col];
        ep = el(ep).next_in_row;
      }
    }
    return y;
  }
  std::vector<T> res(const std::vector<T>& x, const std::vector<T>& rhs) const {  // :298-327
    std::vector<T> xi(num_cols(), zero<T>());
    if (axes[COLS].has_mapping) {
      for (size_t k = 0; k < xi.size(); k++) xi[k] = x.at(axes[COLS].mapping.i2e[k]);
    } else {
      for (size_t k = 0; k < xi.size(); k++) xi[k] = x.at(k);
    }
    std::vector<T> ri = vecmul(xi);
    std::vector<T> r(ri.size(), zero<T>());
    if (axes[ROWS].has_mapping) {
      for (size_t k = 0; k < xi.size(); k++) r.at(k) = rhs.at(k) - ri.at(axes[ROWS].mapping.e2i[k]);
    } else {
      for (size_t k = 0; k < xi.size(); k++) r.at(k) = rhs.at(k) - ri.at(k);
    }
    return r;
  }
  void insert(Element<T>& e) {  // :328-361
    bool expanded = false;
    if (e.row + 1 > num_rows()) { axes[ROWS].grow(e.row + 1); expanded = true; }
    if (e.col + 1 > num_cols()) { axes[COLS].grow(e.col + 1); expanded = true; }
    if (expanded) {
      size_t new_diag_len = std::min(num_rows(), num_cols());
      size_t add = new_diag_len - diag.size();
      for (size_t k = 0; k < add; k++) diag.push_back(-1);
    }
    insert_axis(COLS, e);
    insert_axis(ROWS, e);
    axes[ROWS].qtys[e.row] += 1;
    axes[COLS].qtys[e.col] += 1;
    if (state == MatrixState::FACTORING) {
      axes[ROWS].markowitz[e.row] += 1;
      axes[COLS].markowitz[e.col] += 1;
    }
    if (e.row == e.col) diag.at(e.row) = e.index;
    if (e.fillin) fillins.push_back(e.index);
  }
  void insert_axis(Axis ax, Element<T>& e) {  // :362-391
    int head_ptr = axes[ax].hdrs[e.loc(ax)];
    if (head_ptr < 0) { set_hdr(ax, e.loc(ax), e.index); return; }
    Axis off_ax = other(ax);
    if (el(head_ptr).loc(off_ax) > e.loc(off_ax)) {
      e.set_next(ax, head_ptr);
      set_hdr(ax, e.loc(ax), e.index);
      return;
    }
    int prev = head_ptr;
    while (el(prev).next(ax) >= 0) {
      int next = el(prev).next(ax);
      if (el(next).loc(off_ax) >= e.loc(off_ax)) break;
      prev = next;
    }
    e.set_next(ax, el(prev).next(ax));
    el(prev).set_next(ax, e.index);
  }
  int add_fillin(size_t row, size_t col) { return _add_element(row, col, zero<T>(), true); }  // :392
  int _add_element(size_t row, size_t col, T val, bool fillin) {  // :395-402
    Element<T> e;
    e.index = (int)elements.size();
    e.row = row; e.col = col; e.val = val; e.fillin = fillin;
    e.orig_row = row; e.orig_col = col;
    insert(e);
    elements.push_back(e);
    return e.index;
  }
  int get_elem(size_t row, size_t col) const {  // :404-428
    if (row >= num_rows()) return -1;
    if (col >= num_cols()) return -1;
    if (row == col) return diag[row];
    int ep = hdr(ROWS, row);
    while (ep >= 0) {
      const Element<T>& e = el(ep);
      if (e.col == col) return ep;
      else if (e.col > col) return -1;
      ep = e.next_in_row;
    }
    return -1;
  }
  bool get(size_t row, size_t col, T* out) const {  // :430-435
    int ei = get_elem(row, col);
    if (ei < 0) return false;
    *out = el(ei).val;
    return true;
  }
  void move_element(Axis ax, int idx, size_t to) {  // :436-499
    size_t loc = el(idx).loc(ax);
    if (loc == to) return;
    Axis off_ax = other(ax);
    size_t y = el(idx).loc(off_ax);
    if (loc < to) {
      int br = before_loc(off_ax, y, to, idx);
      if (br < 0) throw Panic("ERROR");
      if (br != idx) {
        int be = prev(off_ax, idx, -1);
        int nxt = el(idx).next(off_ax);
        if (be < 0) set_hdr(off_ax, y, nxt);
        else el(be).set_next(off_ax, nxt);
        int brn = el(br).next(off_ax);
        el(idx).set_next(off_ax, brn);
        el(br).set_next(off_ax, idx);
      }
    } else {
      int br = before_loc(off_ax, y, to, -1);
      int be = prev(off_ax, idx, -1);
      if (br != be) {
        if (be >= 0) {
          int nxt = el(idx).next(off_ax);
          el(be).set_next(off_ax, nxt);
        }
        if (br < 0) {
          int first = hdr(off_ax, y);
          el(idx).set_next(off_ax, first);
          axes[off_ax].hdrs[y] = idx;
        } else if (br != idx) {
          int nxt = el(br).next(off_ax);
          el(idx).set_next(off_ax, nxt);
          el(br).set_next(off_ax, idx);
        }
      }
    }
    el(idx).set_loc(ax, to);
    if (loc == y) diag.at(loc) = -1;
    else if (to == y) diag.at(to) = idx;
  }
  void exchange_elements(Axis ax, int ix, int iy) {  // :500-551
    Axis off_ax = other(ax);
    size_t off_loc = el(ix).loc(off_ax);
    int bx = prev(off_ax, ix, -1);
    int by = prev(off_ax, iy, ix);
    if (by < 0) throw Panic("ERROR!");
    size_t locx = el(ix).loc(ax);
    size_t locy = el(iy).loc(ax);
    el(iy).set_loc(ax, locx);
    el(ix).set_loc(ax, locy);
    if (bx < 0) set_hdr(off_ax, off_loc, iy);
    else el(bx).set_next(off_ax, iy);
    if (by == ix) {
      int tmp = el(iy).next(off_ax);
      el(iy).set_next(off_ax, ix);
      el(ix).set_next(off_ax, tmp);
    } else {
      int xnxt = el(ix).next(off_ax);
      int ynxt = el(iy).next(off_ax);
      el(iy).set_next(off_ax, xnxt);
      el(ix).set_next(off_ax, ynxt);
      el(by).set_next(off_ax, ix);
    }
    if (locx == off_loc) diag.at(off_loc) = iy;
    else if (locy == off_loc) diag.at(off_loc) = ix;
  }
  int prev(Axis ax, int idx, int hint) const {  // :552-575
    int p = hint >= 0 ? hint : hdr(ax, el(idx).loc(ax));
    if (p < 0) return -1;
    if (p == idx) return -1;
    int pi = p;
    while (el(pi).next(ax) >= 0) {
      int nxt = el(pi).next(ax);
      if (nxt == idx) break;
      pi = nxt;
    }
    return pi;
  }
  int before_loc(Axis ax, size_t loc, size_t before, int hint) const {  // :576-598
    int p = hint >= 0 ? hint : hdr(ax, loc);
    Axis off_ax = other(ax);
    if (p < 0) return -1;
    if (el(p).loc(off_ax) >= before) return -1;
    int pi = p;
    while (el(pi).next(ax) >= 0) {
      int nxt = el(pi).next(ax);
      if (el(nxt).loc(off_ax) >= before) break;
      pi = nxt;
    }
    return pi;
  }
  void swap(Axis ax, size_t a, size_t b) {  // :599-643
    if (a == b) return;
    size_t x = std::min(a, b), y = std::max(a, b);
    int ix = axes[ax].hdrs[x];
    int iy = axes[ax].hdrs[y];
    Axis off_ax = other(ax);
    for (;;) {
      if (ix >= 0 && iy >= 0) {
        size_t ox = el(ix).loc(off_ax), oy = el(iy).loc(off_ax);
        if (ox < oy) {
          int ex = ix;
          move_element(ax, ex, y);
          ix = el(ex).next(ax);
        } else if (oy < ox) {
          int ey = iy;
          move_element(ax, ey, x);
          iy = el(ey).next(ax);
        } else {
          int ex = ix, ey = iy;
          exchange_elements(ax, ex, ey);
          ix = el(ex).next(ax);
          iy = el(ey).next(ax);
        }
      } else if (ix < 0 && iy >= 0) {
        int ey = iy;
        move_element(ax, ey, x);
        iy = el(ey).next(ax);
      } else if (ix >= 0 && iy < 0) {
        int ex = ix;
        move_element(ax, ex, y);
        ix = el(ex).next(ax);
      } else {
        break;
      }
    }
    axes[ax].swap(x, y);
  }
  void lu_factorize() {  // :647-674
    if (!(diag.size() > 0)) throw SpError("Assertion Failed");
    for (size_t k = 0; k < axes[ROWS].hdrs.size(); k++)
      if (hdr(ROWS, k) < 0) throw SpError("Singular Matrix");
    for (size_t k = 0; k < axes[COLS].hdrs.size(); k++)
      if (hdr(COLS, k) < 0) throw SpError("Singular Matrix");
    state = MatrixState::FACTORING;
    axes[ROWS].setup_factoring();
    axes[COLS].setup_factoring();
    n_factorizations++;
    for (size_t n = 0; n + 1 < diag.size(); n++) {
      int pivot = search_for_pivot(n);
      if (pivot < 0) throw SpError("Pivot Search Fail");
      swap(ROWS, el(pivot).row, n);
      swap(COLS, el(pivot).col, n);
      row_col_elim(pivot, n);
    }
    state = MatrixState::FACTORED;
  }
  int search_for_pivot(size_t n) const {  // :676-686
    int ei = markowitz_search_diagonal(n);
    if (ei >= 0) return ei;
    ei = markowitz_search_submatrix(n);
    if (ei >= 0) return ei;
    return find_max(n);
  }
  int max_after(Axis ax, int after) const {  // :688-702
    int best = after;
    double best_val = absv(el(after).val);
    int e = el(after).next(ax);
    while (e >= 0) {
      double val = absv(el(e).val);
      if (val > best_val) { best = e; best_val = val; }
      e = el(e).next(ax);
    }
    return best;
  }
  int max_after_loc(Axis ax, size_t in_loc, size_t after_loc) const {  // :704-724
    int e = axes[ax].hdrs[in_loc];
    Axis off_ax = other(ax);
    while (e >= 0) {
      if (el(e).loc(off_ax) >= after_loc) break;
      e = el(e).next(ax);
    }
    if (e < 0) return -1;
    int best = e;
    double best_val = absv(el(best).val);
    while (e >= 0) {
      double val = absv(el(e).val);
      if (val > best_val) { best = e; best_val = val; }
      e = el(e).next(ax);
    }
    return best;
  }
  size_t markowitz_product(int ei) const {  // :726-733
    const Element<T>& e = el(ei);
    size_t mr = axes[ROWS].markowitz[e.row];
    size_t mc = axes[COLS].markowitz[e.col];
    if (!(mr > 0)) throw Panic("assertion failed: mr > 0");
    if (!(mc > 0)) throw Panic("assertion failed: mc > 0");
    return (mr - 1) * (mc - 1);
  }
  int markowitz_search_diagonal(size_t n) const {  // :735-783
    int best_elem = -1;
    size_t best_mark = SIZE_MAX;
    double best_ratio = 0.0;
    size_t num_ties = 0;
    for (size_t k = n; k < diag.size(); k++) {
      int d = diag[k];
      if (d < 0) continue;
      int max_in_col = max_after_loc(COLS, k, n);
      if (max_in_col < 0) continue;
      double threshold = MARKOWITZ_CONFIG.rel_threshold * absv(el(max_in_col).val) + MARKOWITZ_CONFIG.abs_threshold;
      if (absv(el(d).val) < threshold) continue;
      size_t mark = markowitz_product(d);
      if (mark < best_mark) {
        num_ties = 0;
        best_elem = diag[k];
        best_mark = mark;
        best_ratio = absv(el(d).val / el(max_in_col).val);
      } else if (mark == best_mark) {
        num_ties += 1;
        double ratio = absv(el(d).val / el(max_in_col).val);
        if (ratio > best_ratio) {
          best_elem = diag[k];
          best_mark = mark;
          best_ratio = ratio;
        }
        if (num_ties >= best_mark * MARKOWITZ_CONFIG.ties_mult) return best_elem;
      }
    }
    return best_elem;
  }
  int markowitz_search_submatrix(size_t n) const {  // :785-834  (as written: only column n is ever examined)
    int best_elem = -1;
    size_t best_mark = SIZE_MAX;
    double best_ratio = 0.0;
    for (size_t _k = n; _k < axes[COLS].hdrs.size(); _k++) {
      int e = hdr(COLS, n);
      while (e >= 0) {
        if (el(e).row >= n) break;
        e = el(e).next_in_col;
      }
      if (e < 0) continue;
      int max_in_col = max_after(COLS, e);
      while (e >= 0) {
        int ei = e;
        size_t mark = markowitz_product(ei);
        if (mark < best_mark) {
          best_elem = e;
          best_mark = mark;
          best_ratio = absv(el(ei).val / el(max_in_col).val);
        } else if (mark == best_mark) {
          double ratio = absv(el(ei).val / el(max_in_col).val);
          if (ratio > best_ratio) {
            best_elem = e;
            best_mark = mark;
            best_ratio = ratio;
          }
        }
        e = el(ei).next_in_col;
      }
    }
    return best_elem;
  }
  int find_max(size_t n) const {  // :837-863
    int max_elem = -1;
    double max_val = 0.0;
    for (size_t k = n; k < axes[COLS].hdrs.size(); k++) {
      int ep = hdr(COLS, k);
      while (ep >= 0) {
        if (el(ep).row >= n) break;
        ep = el(ep).next_in_col;
      }
      while (ep >= 0) {
        double val = absv(el(ep).val);
        if (val > max_val) { max_elem = ep; max_val = val; }
        ep = el(ep).next_in_col;
      }
    }
    return max_elem;
  }
  void row_col_elim(int pivot, size_t n) {  // :865-919
    int de = diag[n];
    if (de < 0) throw SpError("Singular Matrix");
    if (de != pivot) throw SpError("Assertion Failed");
    T pivot_val = el(pivot).val;
    if (pivot_val == zero<T>()) throw SpError("Assert Neq Failed: zero pivot");  // assert(pivot_val).ne(T::zero())? (assert.rs:38-44)
    int plower = el(pivot).next_in_col;
    while (plower >= 0) {
      el(plower).val = el(plower).val / pivot_val;
      plower = el(plower).next_in_col;
    }
    int pupper = el(pivot).next_in_row;
    while (pupper >= 0) {
      int pue = pupper;
      size_t pupper_col = el(pue).col;
      plower = el(pivot).next_in_col;
      int psub = el(pue).next_in_col;
      while (plower >= 0) {
        int ple = plower;
        while (psub >= 0) {
          if (el(psub).row >= el(ple).row) break;
          psub = el(psub).next_in_col;
        }
        int pse;
        if (psub < 0) pse = add_fillin(el(ple).row, pupper_col);
        else if (el(psub).row > el(ple).row) pse = add_fillin(el(ple).row, pupper_col);
        else pse = psub;
        T v = el(pue).val * el(ple).val;
        el(pse).val = el(pse).val - v;
        psub = el(pse).next_in_col;
        plower = el(ple).next_in_col;
      }
      axes[COLS].markowitz[pupper_col] -= 1;
      pupper = el(pue).next_in_row;
    }
    axes[ROWS].markowitz[n] -= 1;
    axes[COLS].markowitz[n] -= 1;
    plower = el(pivot).next_in_col;
    while (plower >= 0) {
      size_t plower_row = el(plower).row;
      axes[ROWS].markowitz[plower_row] -= 1;
      plower = el(plower).next_in_col;
    }
  }
  std::vector<T> solve(const std::vector<T>& rhs) {  // :929-991
    if (state != MatrixState::FACTORED) lu_factorize();
    std::vector<T> c(rhs.size(), zero<T>());
    if (axes[ROWS].has_mapping) {
      for (size_t k = 0; k < c.size(); k++) c[k] = rhs.at(axes[ROWS].mapping.i2e.at(k));
    } else {
      throw SpError("Missing Row Mapping");
    }
    for (size_t k = 0; k < diag.size(); k++) {
      if (c.at(k) == zero<T>()) continue;
      int di = diag[k];
      if (di < 0) throw SpError("Singular Matrix");
      int e = el(di).next_in_col;
      while (e >= 0) {
        c.at(el(e).row) = c.at(el(e).row) - c[k] * el(e).val;
        e = el(e).next_in_col;
      }
    }
    for (size_t kk = diag.size(); kk-- > 0;) {
      size_t k = kk;
      int di = diag[k];
      if (di < 0) throw SpError("Singular Matrix");
      int ep = el(di).next_in_row;
      while (ep >= 0) {
        c.at(k) = c.at(k) - c.at(el(ep).col) * el(ep).val;
        ep = el(ep).next_in_row;
      }
      c.at(k) = c.at(k) / el(di).val;
    }
    std::vector<T> soln(c.size(), zero<T>());
    if (axes[COLS].has_mapping) {
      for (size_t k = 0; k < c.size(); k++) soln[k] = c.at(axes[COLS].mapping.e2i.at(k));
    } else {
      throw SpError("Missing Column Mapping");
    }
    return soln;
  }
  std::vector<std::vector<T>> to_dense() const {  // :993-999
    std::vector<std::vector<T>> r(num_rows(), std::vector<T>(num_cols(), zero<T>()));
    for (auto& e : elements) r[e.row][e.col] = e.val;
    return r;
  }
  // Port of the tests' `checkups()` invariant checker (mod.rs:1088-1135). Returns "" when all hold.
  std::string checkups() const {
    size_t next_in_rows = 0, next_in_cols = 0;
    for (auto& e : elements) {
      if (e.next_in_row >= 0) {
        next_in_rows++;
        if (!(el(e.next_in_row).col > e.col)) return "row order";
        if (el(e.next_in_row).row != e.row) return "row membership";
      }
      if (e.next_in_col >= 0) {
        next_in_cols++;
        if (!(el(e.next_in_col).row > e.row)) return "col order";
        if (el(e.next_in_col).col != e.col) return "col membership";
      }
    }
    size_t row_hdrs = 0, col_hdrs = 0;
    for (int h : axes[ROWS].hdrs) if (h >= 0) row_hdrs++;
    for (int h : axes[COLS].hdrs) if (h >= 0) col_hdrs++;
    if (next_in_rows + row_hdrs != elements.size()) return "row count";
    if (next_in_cols + col_hdrs != elements.size()) return "col count";
    for (size_t k = 0; k < diag.size(); k++) {
      int d = diag[k];
      if (d >= 0 && !(el(d).row == k && el(d).col == k)) return "diag";
      if (d < 0 && get_elem_slow(k, k) >= 0) return "diag missing";
    }
    for (int ax = 0; ax < 2; ax++) {
      for (size_t k = 0; k < axes[ax].hdrs.size(); k++) {
        size_t cnt = 0;
        int e = axes[ax].hdrs[k];
        while (e >= 0) { cnt++; if (el(e).loc((Axis)ax) != k) return "loc"; e = el(e).next((Axis)ax); }
        if (cnt != axes[ax].qtys[k]) return "qtys";
      }
    }
    return "";
  }
  int get_elem_slow(size_t row, size_t col) const {
    for (auto& e : elements) if (e.row == row && e.col == col) return e.index;
    return -1;
  }
};

}  // namespace orc
