// Declaration-only stand-in for <Rcpp.h>: TEST INFRASTRUCTURE.  R and Rcpp are not installed in the build image, so the
// package's native layer (rpkg/src/*.cpp) cannot be built here; this header lets `g++ -fsyntax-only` parse and TYPE-CHECK
// it -- every call into the C-ABI of include/fable_b200.h (argument counts and types), every list element, every
// registered arity -- which is what can go wrong in pure marshalling code.  Nothing here has a body worth running.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <initializer_list>
#include <string>
#include <vector>

typedef struct SEXPREC* SEXP;
typedef std::ptrdiff_t R_xlen_t;
typedef void* (*DL_FUNC)();
struct DllInfo;
struct R_CallMethodDef { const char* name; DL_FUNC fun; int numArgs; };
#ifndef FALSE
#define FALSE 0
#endif
extern "C" {
int R_registerRoutines(DllInfo*, const void*, const R_CallMethodDef*, const void*, const void*);
int R_useDynamicSymbols(DllInfo*, int);
double unif_rand(void);
}

namespace Rcpp {

class NumericVector {
 public:
  NumericVector();
  explicit NumericVector(R_xlen_t n);
  NumericVector(SEXP);
  double* begin();
  const double* begin() const;
  R_xlen_t size() const;
  static NumericVector create(double a, double b);
};
class IntegerVector {
 public:
  IntegerVector();
  IntegerVector(SEXP);
  const int* begin() const;
  const int* end() const;
  R_xlen_t size() const;
};
class NumericMatrix {
 public:
  NumericMatrix();
  NumericMatrix(int nrow, int ncol);
  NumericMatrix(SEXP);
  double* begin();
  const double* begin() const;
  int nrow() const;
  int ncol() const;
  double& operator()(int i, int j);
  double operator()(int i, int j) const;
};
class RObject {
 public:
  RObject();
  template <typename T> RObject& operator=(const T&);
  operator SEXP() const;
};
class List;
struct NamedArg { };
struct Named {
  explicit Named(const char*);
  template <typename T> NamedArg operator=(const T&) const;
};
struct ListProxy {
  operator NumericMatrix() const;
  operator NumericVector() const;
  operator int() const;
  operator double() const;
};
class List {
 public:
  List();
  List(SEXP);
  template <typename... A> static List create(const A&...);
  ListProxy operator[](const char*) const;
};
struct RNGScope { RNGScope(); ~RNGScope(); };
[[noreturn]] void stop(const char* fmt, ...);
void checkUserInterrupt();
NumericVector rgamma(int n, double shape, double scale);
template <typename T> SEXP wrap(const T&);
template <typename T> T as(SEXP);

}  // namespace Rcpp

#define BEGIN_RCPP try {
#define END_RCPP } catch (...) { } return static_cast<SEXP>(nullptr);
#define RcppExport extern "C"
