// armadillo_min.hpp -- TEST INFRASTRUCTURE.  A minimal, eager, header-only stand-in for the subset of
// Armadillo that the reference's src/updated-FABLE-functions.cpp uses, so that the reference's OWN source
// file can be compiled where it lies (oracle/Makefile target _ref/libfable_ref.so) and run against the
// restated oracle.  It is NOT Armadillo: expression semantics are restated from Armadillo's documentation
// (column-major storage; sum(X,0) = column sums as a row, sum(X,1) = row sums as a column; `/` and `%`
// between matrices are element-wise; each_row()/each_col() broadcast in place; trimatu_ind enumerates the
// upper triangle in column-major order; fill::randn fills in memory order; quantile = Hyndman-Fan type 5).
// Random numbers come from caller-installed hooks (injected draws), in the order the reference consumes them.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <functional>
#include <initializer_list>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

namespace arma {

typedef unsigned long long uword;

namespace fill {
struct zeros_t {}; struct ones_t {}; struct randn_t {};
static const zeros_t zeros = zeros_t();
static const ones_t ones = ones_t();
static const randn_t randn = randn_t();
}  // namespace fill

// hooks installed by the bridge: standard normals / Gamma(shape a, scale b) variates, one at a time
struct rng_hooks {
  static std::function<double()>& normal() { static std::function<double()> f; return f; }
  static std::function<double(double, double)>& gamma() { static std::function<double(double, double)> f; return f; }
};

struct SizeMat { uword n_rows, n_cols; };

template <typename T> class Mat;
template <typename T> struct RowView;
template <typename T> struct DiagView;
template <typename T> struct EachRow;
template <typename T> struct EachCol;

template <typename T>
class Mat {
 public:
  uword n_rows = 0, n_cols = 0, n_elem = 0;
  std::vector<T> mem;   // column-major

  Mat() {}
  Mat(uword r, uword c) { init(r, c, T(0)); }
  Mat(uword r, uword c, fill::zeros_t) { init(r, c, T(0)); }
  Mat(uword r, uword c, fill::ones_t) { init(r, c, T(1)); }
  Mat(uword r, uword c, fill::randn_t) {
    init(r, c, T(0));
    for (uword i = 0; i < n_elem; ++i) mem[i] = rng_hooks::normal()();   // memory (column-major) order
  }
  // column-vector forms (arma::vec v(n, fill::...))
  explicit Mat(uword n) { init(n, 1, T(0)); }
  Mat(uword n, fill::zeros_t) { init(n, 1, T(0)); }
  Mat(uword n, fill::ones_t) { init(n, 1, T(1)); }
  Mat(std::initializer_list<T> l) { init(l.size(), 1, T(0)); std::copy(l.begin(), l.end(), mem.begin()); }
  Mat(const RowView<T>& v);
  Mat(const DiagView<T>& v);

  void init(uword r, uword c, T val) { n_rows = r; n_cols = c; n_elem = r * c; mem.assign(n_elem, val); }
  void check(bool ok, const char* what) const { if (!ok) throw std::logic_error(std::string("arma shim: ") + what); }

  T& operator()(uword i) { check(i < n_elem, "index out of bounds"); return mem[i]; }
  const T& operator()(uword i) const { check(i < n_elem, "index out of bounds"); return mem[i]; }
  T& operator()(uword i, uword j) { check(i < n_rows && j < n_cols, "index out of bounds"); return mem[i + n_rows * j]; }
  const T& operator()(uword i, uword j) const { check(i < n_rows && j < n_cols, "index out of bounds"); return mem[i + n_rows * j]; }
  // gather by linear indices: B(uvec)
  Mat<T> operator()(const Mat<uword>& idx) const {
    Mat<T> out(idx.n_elem, 1);
    for (uword i = 0; i < idx.n_elem; ++i) { check(idx.mem[i] < n_elem, "index out of bounds"); out.mem[i] = mem[idx.mem[i]]; }
    return out;
  }

  Mat<T> t() const {
    Mat<T> out(n_cols, n_rows);
    for (uword j = 0; j < n_cols; ++j) for (uword i = 0; i < n_rows; ++i) out.mem[j + n_cols * i] = mem[i + n_rows * j];
    return out;
  }
  Mat<T> cols(uword a, uword b) const {
    check(a <= b && b < n_cols, "cols(): indices out of bounds");
    Mat<T> out(n_rows, b - a + 1);
    std::copy(mem.begin() + n_rows * a, mem.begin() + n_rows * (b + 1), out.mem.begin());
    return out;
  }
  Mat<T> col(uword j) const { return cols(j, j); }
  Mat<T> subvec(uword a, uword b) const {
    check(a <= b && b < n_elem, "subvec(): indices out of bounds");
    Mat<T> out(b - a + 1, 1);
    if (n_rows == 1) out.init(1, b - a + 1, T(0));
    std::copy(mem.begin() + a, mem.begin() + b + 1, out.mem.begin());
    return out;
  }
  RowView<T> row(uword i) { check(i < n_rows, "row(): index out of bounds"); return RowView<T>{this, i}; }
  DiagView<T> diag() { return DiagView<T>{this}; }
  EachRow<T> each_row() { return EachRow<T>{this}; }
  EachCol<T> each_col() { return EachCol<T>{this}; }
  uword index_min() const {
    check(n_elem > 0, "index_min(): empty");
    return static_cast<uword>(std::min_element(mem.begin(), mem.end()) - mem.begin());   // first minimum
  }
};

template <typename T> struct RowView {
  Mat<T>* m; uword i;
  RowView& operator=(const Mat<T>& r) {
    m->check(r.n_elem == m->n_cols, "row() assignment: size mismatch");
    for (uword j = 0; j < m->n_cols; ++j) (*m)(i, j) = r.mem[j];
    return *this;
  }
};
template <typename T> struct DiagView {
  Mat<T>* m;
  uword len() const { return std::min(m->n_rows, m->n_cols); }
  DiagView& operator=(const Mat<T>& v) {
    m->check(v.n_elem == len(), "diag() assignment: size mismatch");
    for (uword i = 0; i < len(); ++i) (*m)(i, i) = v.mem[i];
    return *this;
  }
};
template <typename T> Mat<T>::Mat(const RowView<T>& v) { init(1, v.m->n_cols, T(0)); for (uword j = 0; j < n_cols; ++j) mem[j] = (*v.m)(v.i, j); }
template <typename T> Mat<T>::Mat(const DiagView<T>& v) { init(v.len(), 1, T(0)); for (uword i = 0; i < n_elem; ++i) mem[i] = (*v.m)(i, i); }
template <typename T> Mat<T> operator/(const DiagView<T>& d, double s) { Mat<T> v(d); for (auto& x : v.mem) x /= s; return v; }

template <typename T> struct EachRow {
  Mat<T>* m;
  void operator%=(const Mat<T>& r) {
    m->check(r.n_elem == m->n_cols, "each_row(): size mismatch");
    for (uword j = 0; j < m->n_cols; ++j) for (uword i = 0; i < m->n_rows; ++i) (*m)(i, j) *= r.mem[j];
  }
  void operator/=(const Mat<T>& r) {
    m->check(r.n_elem == m->n_cols, "each_row(): size mismatch");
    for (uword j = 0; j < m->n_cols; ++j) for (uword i = 0; i < m->n_rows; ++i) (*m)(i, j) /= r.mem[j];
  }
};
template <typename T> struct EachCol {
  Mat<T>* m;
  void operator%=(const Mat<T>& c) {
    m->check(c.n_elem == m->n_rows, "each_col(): size mismatch");
    for (uword j = 0; j < m->n_cols; ++j) for (uword i = 0; i < m->n_rows; ++i) (*m)(i, j) *= c.mem[i];
  }
};

typedef Mat<double> mat;
typedef Mat<double> vec;
typedef Mat<double> rowvec;
typedef Mat<uword> uvec;

// ---- element-wise and scalar arithmetic -------------------------------------------------------
template <typename T, typename F> Mat<T> zip(const Mat<T>& a, const Mat<T>& b, F f, const char* what) {
  a.check(a.n_rows == b.n_rows && a.n_cols == b.n_cols, what);
  Mat<T> out(a.n_rows, a.n_cols);
  for (uword i = 0; i < a.n_elem; ++i) out.mem[i] = f(a.mem[i], b.mem[i]);
  return out;
}
template <typename T, typename F> Mat<T> map(const Mat<T>& a, F f) {
  Mat<T> out(a.n_rows, a.n_cols);
  for (uword i = 0; i < a.n_elem; ++i) out.mem[i] = f(a.mem[i]);
  return out;
}
inline mat operator+(const mat& a, const mat& b) { return zip(a, b, [](double x, double y) { return x + y; }, "addition: size mismatch"); }
inline mat operator-(const mat& a, const mat& b) { return zip(a, b, [](double x, double y) { return x - y; }, "subtraction: size mismatch"); }
inline mat operator%(const mat& a, const mat& b) { return zip(a, b, [](double x, double y) { return x * y; }, "element-wise multiplication: size mismatch"); }
inline mat operator/(const mat& a, const mat& b) { return zip(a, b, [](double x, double y) { return x / y; }, "element-wise division: size mismatch"); }
inline mat operator/(const mat& a, double s) { return map(a, [s](double x) { return x / s; }); }
inline mat operator/(double s, const mat& a) { return map(a, [s](double x) { return s / x; }); }
inline mat operator*(double s, const mat& a) { return map(a, [s](double x) { return s * x; }); }
inline mat operator*(const mat& a, double s) { return map(a, [s](double x) { return x * s; }); }
inline mat operator-(const mat& a) { return map(a, [](double x) { return -x; }); }
inline uvec operator-(const uvec& a, int s) { uvec o(a.n_rows, a.n_cols); for (uword i = 0; i < a.n_elem; ++i) o.mem[i] = a.mem[i] - static_cast<uword>(s); return o; }

// matrix product (column-major, axpy order: vectorises)
inline mat operator*(const mat& a, const mat& b) {
  a.check(a.n_cols == b.n_rows, "matrix multiplication: incompatible dimensions");
  mat c(a.n_rows, b.n_cols);
  for (uword j = 0; j < b.n_cols; ++j)
    for (uword l = 0; l < a.n_cols; ++l) {
      const double blj = b.mem[l + b.n_rows * j];
      const double* al = &a.mem[a.n_rows * l];
      double* cj = &c.mem[c.n_rows * j];
      for (uword i = 0; i < a.n_rows; ++i) cj[i] += al[i] * blj;
    }
  return c;
}

inline mat square(const mat& a) { return map(a, [](double x) { return x * x; }); }
inline mat sqrt(const mat& a) { return map(a, [](double x) { return std::sqrt(x); }); }
inline mat log(const mat& a) { return map(a, [](double x) { return std::log(x); }); }
using std::sqrt;
using std::log;

inline double accu(const mat& a) {   // Armadillo's accu: two running sums over the linear memory
  double s1 = 0.0, s2 = 0.0;
  uword i = 0;
  for (; i + 1 < a.n_elem; i += 2) { s1 += a.mem[i]; s2 += a.mem[i + 1]; }
  if (i < a.n_elem) s1 += a.mem[i];
  return s1 + s2;
}
inline double mean(const mat& v) { v.check(v.n_rows == 1 || v.n_cols == 1, "mean(): vector expected"); return accu(v) / static_cast<double>(v.n_elem); }
inline double as_scalar(const mat& a) { a.check(a.n_elem == 1, "as_scalar(): not 1x1"); return a.mem[0]; }

inline mat sum(const mat& a, int dim) {
  if (dim == 0) {
    mat out(1, a.n_cols);
    for (uword j = 0; j < a.n_cols; ++j) { double s = 0.0; for (uword i = 0; i < a.n_rows; ++i) s += a.mem[i + a.n_rows * j]; out.mem[j] = s; }
    return out;
  }
  mat out(a.n_rows, 1);
  for (uword j = 0; j < a.n_cols; ++j) for (uword i = 0; i < a.n_rows; ++i) out.mem[i] += a.mem[i + a.n_rows * j];
  return out;
}

inline mat diagmat(const mat& v) {
  mat out(v.n_elem, v.n_elem);
  for (uword i = 0; i < v.n_elem; ++i) out.mem[i + v.n_elem * i] = v.mem[i];
  return out;
}
inline mat vectorise(const mat& a) { mat out(a.n_elem, 1); out.mem = a.mem; return out; }
inline SizeMat size(const mat& a) { return SizeMat{a.n_rows, a.n_cols}; }
inline uvec trimatu_ind(const SizeMat& s) {
  std::vector<uword> v;
  for (uword j = 0; j < s.n_cols; ++j) for (uword i = 0; i <= j && i < s.n_rows; ++i) v.push_back(i + s.n_rows * j);
  uvec out(v.size(), 1);
  out.mem = v;
  return out;
}
inline mat regspace(double a, double d, double b) {
  std::vector<double> v;
  for (double x = a; x <= b + 1e-12; x += d) v.push_back(x);
  mat out(v.size(), 1);
  out.mem = v;
  return out;
}

struct distr_param { double a, b; distr_param(double a_, double b_) : a(a_), b(b_) {} };
inline mat randg(uword n, const distr_param& dp) {   // Gamma(shape a, scale b), drawn one by one in order
  mat out(n, 1);
  for (uword i = 0; i < n; ++i) out.mem[i] = rng_hooks::gamma()(dp.a, dp.b);
  return out;
}

// quantile(V, P): continuous sample quantile, Hyndman-Fan type 5
inline mat quantile(const mat& v, const mat& P) {
  std::vector<double> y = v.mem;
  std::sort(y.begin(), y.end());
  const double N = static_cast<double>(y.size());
  mat out(P.n_elem, 1);
  for (uword q = 0; q < P.n_elem; ++q) {
    const double p = P.mem[q];
    double val;
    if (p < 0.5 / N) val = p < 0.0 ? -std::numeric_limits<double>::infinity() : y.front();
    else if (p > (N - 0.5) / N) val = p > 1.0 ? std::numeric_limits<double>::infinity() : y.back();
    else {
      const uword k = static_cast<uword>(std::floor(N * p + 0.5));
      const double Pk = (static_cast<double>(k) - 0.5) / N, w = (p - Pk) * N;
      val = k >= y.size() ? y.back() : (1.0 - w) * y[k - 1] + w * y[k];   // p == P_max exactly: w == 0, the maximum
    }
    out.mem[q] = val;
  }
  return out;
}

// CPPsvd is exported by the reference but never called on the hot path: not provided by the shim
inline bool svd(mat&, vec&, mat&, const mat&, const char* = "dc") { throw std::logic_error("arma shim: svd() is not provided"); }
inline bool svd_econ(mat&, vec&, mat&, const mat&, const char* = "both", const char* = "dc") { throw std::logic_error("arma shim: svd_econ() is not provided"); }

}  // namespace arma
