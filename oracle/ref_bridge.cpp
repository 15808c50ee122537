// ref_bridge.cpp -- TEST INFRASTRUCTURE.  Compiles the REFERENCE'S OWN translation unit
// (/root/reference/src/updated-FABLE-functions.cpp, included from where it lies, never copied) against the
// minimal Armadillo/Rcpp stand-ins in oracle/shim/ and exposes its exported functions through a C ABI, so
// tests can pin the restated oracle against the reference's own function bodies: formulas, index
// conventions, control flow, list layout and random-draw consumption order are the reference's; only the
// linear-algebra primitives underneath are the shim's (see oracle/shim/armadillo_min.hpp).
// Output: oracle/_ref/libfable_ref.so (git-ignored; travels to the GPU box with the snapshot).
#include FABLE_REFERENCE_SOURCE

#include <cstdint>
#include <cstring>

namespace {
arma::mat to_mat(const double* p, int64_t r, int64_t c) {
  arma::mat m(static_cast<arma::uword>(r), static_cast<arma::uword>(c));
  std::memcpy(m.mem.data(), p, sizeof(double) * r * c);
  return m;
}
void out(double* dst, const arma::mat& m) { if (dst) std::memcpy(dst, m.mem.data(), sizeof(double) * m.n_elem); }

// injected draws, consumed strictly in call order
const double* g_gam = nullptr; const double* g_z = nullptr;
int64_t g_ngam = 0, g_nz = 0, g_igam = 0, g_iz = 0;
void install(const double* gam, int64_t ngam, const double* z, int64_t nz) {
  g_gam = gam; g_ngam = ngam; g_igam = 0; g_z = z; g_nz = nz; g_iz = 0;
  arma::rng_hooks::normal() = []() -> double {
    if (g_iz >= g_nz) throw std::logic_error("bridge: normal draws exhausted");
    return g_z[g_iz++];
  };
  arma::rng_hooks::gamma() = [](double, double scale) -> double {
    if (g_igam >= g_ngam) throw std::logic_error("bridge: gamma draws exhausted");
    return g_gam[g_igam++] * scale;      // injected Gamma(shape, 1) variates times the requested scale
  };
}
thread_local std::string g_err;
template <typename F> int guard(F f) {
  try { f(); return 0; } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
}  // namespace

extern "C" {
const char* ref_last_error() { return g_err.c_str(); }

int ref_criterion(int kInd, const double* Y, int64_t n, int64_t p, const double* U, const double* V, const double* d,
                  int64_t r, double* o) {
  return guard([&] { *o = CPPCriterionFunction(kInd, to_mat(Y, n, p), to_mat(U, n, r), to_mat(V, p, r), to_mat(d, r, 1)); });
}
int ref_bisection(int lower, int upper, const double* Y, int64_t n, int64_t p, const double* U, const double* V,
                  const double* d, int64_t r, double* lohi) {
  return guard([&] { out(lohi, CPPBisectionRecursion(lower, upper, to_mat(Y, n, p), to_mat(U, n, r), to_mat(V, p, r), to_mat(d, r, 1))); });
}
int ref_rank(const double* Y, int64_t n, int64_t p, const double* U, const double* V, const double* d, int64_t r,
             int kMax, int* k) {
  return guard([&] { *k = CPPRankEstimator(to_mat(Y, n, p), to_mat(U, n, r), to_mat(V, p, r), to_mat(d, r, 1), kMax); });
}
int ref_hyper(const double* Y, int64_t n, int64_t p, const double* U, const double* V, const double* d, int64_t r,
              int kEst, double* sig, double* G, double* G0, double* tau) {
  return guard([&] {
    Rcpp::List l = FABLEHyperParameters(to_mat(Y, n, p), to_mat(U, n, r), to_mat(V, p, r), to_mat(d, r, 1), kEst);
    out(sig, l["SigmaSqEstimate"]); out(G, l["G"]); out(G0, l["G0"]); *tau = l["tauSqEstimate"];
  });
}
int ref_cov_correct(const double* sig, const double* G, int64_t p, double* Bupper) {
  return guard([&] { out(Bupper, CPPcov_correct_matrix(to_mat(sig, p, 1), to_mat(G, p, p))); });
}
int ref_post_mean(const double* Y, int64_t n, int64_t p, double gamma0, double delta0sq, const double* U,
                  const double* V, const double* d, int64_t r, int kMax, double* Cov, int* kEst) {
  return guard([&] {
    Rcpp::List l = CPPFABLEPostMean(to_mat(Y, n, p), gamma0, delta0sq, to_mat(U, n, r), to_mat(V, p, r), to_mat(d, r, 1), kMax);
    out(Cov, l["CovPostMean"]); *kEst = l["kEst"];
  });
}
// gam: MC*p Gamma(shape = (gamma0+n)/2, scale 1) variates, z: MC*p*k standard normals, both in the order
// the reference's loop consumes them (cpp:599-618)
int ref_sampler(const double* Y, int64_t n, int64_t p, double gamma0, double delta0sq, int MC, const double* U,
                const double* V, const double* d, int64_t r, int kEst, double varInflation, const double* gam,
                const double* z, double* L, double* S, double* postmean, double* plugin, double* tau, double* G,
                int64_t* consumed) {
  return guard([&] {
    install(gam, static_cast<int64_t>(MC) * p, z, static_cast<int64_t>(MC) * p * kEst);
    Rcpp::List l = CPPFABLESampler(to_mat(Y, n, p), gamma0, delta0sq, MC, to_mat(U, n, r), to_mat(V, p, r), to_mat(d, r, 1), kEst,
                                   varInflation);
    out(L, l["LambdaSamples"]); out(S, l["SigmaSqSamples"]); out(postmean, l["SigmaSqEstimatePostMean"]);
    out(plugin, l["SigmaSqEstimatePlugIn"]); *tau = l["tauSqEstimate"]; out(G, l["G"]);
    if (consumed) { consumed[0] = g_igam; consumed[1] = g_iz; }
  });
}
// which: 0 = CPPCCFABLEPostProcessing, 1 = CCFABLEPostProcessingSubmatrix, 2 = ..._Optimized
int ref_post_processing(int which, const double* L, const double* S, int64_t MC, int64_t p, int k, double alpha,
                        const int64_t* sel1, int64_t nsel, double* mean, double* lower, double* upper) {
  return guard([&] {
    Rcpp::List in = Rcpp::List::create(Rcpp::Named("LambdaSamples") = to_mat(L, MC, p * k),
                                       Rcpp::Named("SigmaSqSamples") = to_mat(S, MC, p), Rcpp::Named("EstimatedRank") = k);
    arma::uvec sel(static_cast<arma::uword>(nsel), 1);
    for (int64_t i = 0; i < nsel; ++i) sel.mem[i] = static_cast<arma::uword>(sel1[i]);
    Rcpp::List o = which == 0 ? CPPCCFABLEPostProcessing(in, alpha)
                 : which == 1 ? CCFABLEPostProcessingSubmatrix(in, alpha, sel)
                              : CCFABLEPostProcessingSubmatrix_Optimized(in, alpha, sel);
    out(mean, o["PostMeanMatrix"]); out(lower, o["LowerQuantileMatrix"]); out(upper, o["UpperQuantileMatrix"]);
  });
}
int ref_symmetrize(const double* A, int64_t n, double* o) { return guard([&] { out(o, SymmetrizeMatrix(to_mat(A, n, n))); }); }
int ref_loglik(const double* Y, int64_t n, int64_t p, const double* M, const double* Lam, int k, const double* sig, double* o) {
  return guard([&] { *o = CPPLogLikelihoodEval(to_mat(Y, n, p), to_mat(M, n, k), to_mat(Lam, p, k), to_mat(sig, p, 1)); });
}
}  // extern "C"
