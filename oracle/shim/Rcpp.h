// Rcpp.h (shim) -- TEST INFRASTRUCTURE.  The few Rcpp names the reference's translation unit touches:
// Rcpp::List (create / operator[]), Rcpp::Named, Rcpp::Rcout, Rcpp::checkUserInterrupt.
#pragma once
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <utility>

#include "armadillo_min.hpp"

namespace Rcpp {

struct Value {
  int kind = 0;   // 0 none, 1 matrix/vector, 2 double, 3 int
  arma::mat m;
  double d = 0.0;
  int i = 0;
  Value() {}
  Value(const arma::mat& x) : kind(1), m(x) {}
  Value(double x) : kind(2), d(x) {}
  Value(int x) : kind(3), i(x) {}
  operator arma::mat() const { return m; }
  operator double() const { return kind == 3 ? static_cast<double>(i) : d; }
  operator int() const { return kind == 2 ? static_cast<int>(d) : i; }
};

struct NamedValue { std::string name; Value value; };
struct Named {
  std::string name;
  explicit Named(const char* n) : name(n) {}
  template <typename T> NamedValue operator=(const T& v) const { return NamedValue{name, Value(v)}; }
};

class List {
 public:
  std::map<std::string, Value> items;
  static void add(List&) {}
  template <typename... Rest> static void add(List& l, const NamedValue& nv, const Rest&... rest) {
    l.items[nv.name] = nv.value;
    add(l, rest...);
  }
  template <typename... Args> static List create(const Args&... args) { List l; add(l, args...); return l; }
  const Value& operator[](const char* name) const {
    auto it = items.find(name);
    if (it == items.end()) throw std::logic_error(std::string("Rcpp shim: no list element named ") + name);
    return it->second;
  }
};

struct NullStream : std::ostream {
  struct Buf : std::streambuf { int overflow(int c) override { return c; } } buf;
  NullStream() : std::ostream(&buf) {}
};
static NullStream Rcout;
inline void checkUserInterrupt() {}

}  // namespace Rcpp
