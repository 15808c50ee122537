// Rcpp host layer of the B200 drop-in: replaces the hot exports of the reference's
// src/updated-FABLE-functions.cpp with pure marshalling into the C-ABI of libfable_b200.so
// (include/fable_b200.h).  Signatures, argument order and returned list element names are the
// reference's (file:line cited per function), so Rcpp::compileAttributes() regenerates the SAME
// R/RcppExports.R and src/RcppExports.cpp symbols (`_FABLE_<name>`, same arity).
//
// No Armadillo: inputs are R matrices (column-major doubles) passed as raw pointers; outputs are R
// vectors allocated here and filled by the library with one device-to-host copy.
// NOT COMPILED in the build image (no R / Rcpp there): kept thin enough to review by eye.
#include <Rcpp.h>
#include "fable_b200.h"

namespace {

// One context per R session, created on first use.  The Philox key is drawn from R's RNG inside the
// RNGScope that Rcpp wraps around every export (reference: src/RcppExports.cpp:159), so set.seed()
// makes the native draws reproducible.
fable_ctx* ctx_handle = nullptr;

fable_ctx* ctx() {
  if (!ctx_handle) {
    const char* dev = std::getenv("FABLE_B200_DEVICE");
    int rc = fable_ctx_create(dev ? std::atoi(dev) : 0, 0, &ctx_handle);
    if (rc != FABLE_OK) Rcpp::stop("libfable_b200: %s", fable_last_error(nullptr));
    fable_ctx_set_interrupt(ctx_handle, [](void*) -> int {
      try { Rcpp::checkUserInterrupt(); } catch (...) { return 1; }   // cpp:699 pattern
      return 0;
    }, nullptr);
  }
  return ctx_handle;
}

void check(int rc) {
  if (rc != FABLE_OK) Rcpp::stop("libfable_b200 (status %d): %s", rc, fable_last_error(ctx_handle));
}

void reseed() {
  const uint64_t hi = static_cast<uint64_t>(unif_rand() * 4294967296.0);
  const uint64_t lo = static_cast<uint64_t>(unif_rand() * 4294967296.0);
  check(fable_ctx_set_seed(ctx(), (hi << 32) | lo));
}

struct Svd {
  const double *Y, *U, *V, *d;
  int64_t n, p, r;
  Svd(const Rcpp::NumericMatrix& Ym, const Rcpp::NumericMatrix& Um, const Rcpp::NumericMatrix& Vm,
      const Rcpp::NumericVector& dv)
      : Y(Ym.begin()), U(Um.begin()), V(Vm.begin()), d(dv.begin()), n(Ym.nrow()), p(Ym.ncol()), r(dv.size()) {
    if (Um.nrow() != n || Um.ncol() < r) Rcpp::stop("U_Y must be n x length(svalsY)");
    if (Vm.nrow() != p || Vm.ncol() < r) Rcpp::stop("V_Y must be p x length(svalsY)");
  }
};

}  // namespace

// reference cpp:95-125
// [[Rcpp::export]]
double CPPLogLikelihoodEval(Rcpp::NumericMatrix Y, Rcpp::NumericMatrix M, Rcpp::NumericMatrix Lambda, Rcpp::NumericVector SigmaSq) {
  double out = 0.0;
  check(fable_log_likelihood_eval(ctx(), Y.begin(), Y.nrow(), Y.ncol(), M.begin(), Lambda.begin(), M.ncol(), SigmaSq.begin(), &out));
  return out;
}

// reference cpp:127-213
// [[Rcpp::export]]
double CPPCriterionFunction(int kInd, Rcpp::NumericMatrix Y, Rcpp::NumericMatrix U_Y, Rcpp::NumericMatrix V_Y,
                            Rcpp::NumericVector svalsY) {
  Svd s(Y, U_Y, V_Y, svalsY);
  double out = 0.0;
  check(fable_criterion(ctx(), kInd, s.Y, s.n, s.p, s.U, s.V, s.d, s.r, &out));
  return out;
}

// reference cpp:215-271 (returns the window as a length-2 double vector, :263-267)
// [[Rcpp::export]]
Rcpp::NumericVector CPPBisectionRecursion(int lower, int upper, Rcpp::NumericMatrix Y, Rcpp::NumericMatrix U_Y,
                                          Rcpp::NumericMatrix V_Y, Rcpp::NumericVector svalsY) {
  Svd s(Y, U_Y, V_Y, svalsY);
  int lo = 0, hi = 0;
  check(fable_bisection(ctx(), lower, upper, s.Y, s.n, s.p, s.U, s.V, s.d, s.r, &lo, &hi));
  return Rcpp::NumericVector::create(lo, hi);
}

// reference cpp:281-341
// [[Rcpp::export]]
int CPPRankEstimator(Rcpp::NumericMatrix Y, Rcpp::NumericMatrix U_Y, Rcpp::NumericMatrix V_Y,
                     Rcpp::NumericVector svalsY, int kMax) {
  Svd s(Y, U_Y, V_Y, svalsY);
  int k = 0;
  check(fable_rank_estimator(ctx(), s.Y, s.n, s.p, s.U, s.V, s.d, s.r, kMax, &k));
  return k;
}

// reference cpp:343-399 (list element names :393-397)
// [[Rcpp::export]]
Rcpp::List FABLEHyperParameters(Rcpp::NumericMatrix Y, Rcpp::NumericMatrix U_Y, Rcpp::NumericMatrix V_Y,
                                Rcpp::NumericVector svalsY, int kEst) {
  Svd s(Y, U_Y, V_Y, svalsY);
  if (kEst < 1 || kEst > s.r) Rcpp::stop("kEst must be in 1..length(svalsY)");
  Rcpp::NumericVector sig(s.p);
  Rcpp::NumericMatrix G0(s.p, kEst), G(s.p, s.p);
  double tau = 0.0;
  check(fable_hyper_parameters(ctx(), s.Y, s.n, s.p, s.U, s.V, s.d, s.r, kEst, sig.begin(), G0.begin(), &tau, G.begin()));
  return Rcpp::List::create(Rcpp::Named("SigmaSqEstimate") = sig, Rcpp::Named("G") = G,
                            Rcpp::Named("EstimatedRank") = kEst, Rcpp::Named("G0") = G0,
                            Rcpp::Named("tauSqEstimate") = tau);
}

// reference cpp:401-428: upper-triangle VECTOR in trimatu_ind (column-major) order
// [[Rcpp::export]]
Rcpp::NumericVector CPPcov_correct_matrix(Rcpp::NumericVector sigsq_hat, Rcpp::NumericMatrix llprime_hat) {
  const int64_t p = sigsq_hat.size();
  if (llprime_hat.nrow() != p || llprime_hat.ncol() != p) Rcpp::stop("llprime_hat must be p x p");
  Rcpp::NumericVector out(static_cast<R_xlen_t>(p) * (p + 1) / 2);
  check(fable_cov_correct_matrix(ctx(), sigsq_hat.begin(), llprime_hat.begin(), p, out.begin()));
  return out;
}

// NEW (not in the reference): the scalar R/FABLE_wrapper.R:41-44 derives from cov_correct_matrix(),
// computed from the factor G0 without forming B.  variant 0 = wrapper formula, 1 = script formula.
// [[Rcpp::export]]
double FABLEVarInflation(Rcpp::NumericVector sigsq_hat, Rcpp::NumericMatrix G0, int variant = 0) {
  double out = 0.0;
  check(fable_var_inflation(ctx(), sigsq_hat.begin(), G0.begin(), G0.nrow(), G0.ncol(), variant, &out));
  return out;
}

// reference cpp:440-519 (list :516-517)
// [[Rcpp::export]]
Rcpp::List CPPFABLEPostMean(Rcpp::NumericMatrix Y, double gamma0, double delta0sq, Rcpp::NumericMatrix U_Y,
                            Rcpp::NumericMatrix V_Y, Rcpp::NumericVector svalsY, int kMax) {
  Svd s(Y, U_Y, V_Y, svalsY);
  Rcpp::NumericMatrix Cov(s.p, s.p);
  int k = 0;
  check(fable_post_mean(ctx(), s.Y, s.n, s.p, gamma0, delta0sq, s.U, s.V, s.d, s.r, kMax, Cov.begin(), &k, nullptr, nullptr));
  return Rcpp::List::create(Rcpp::Named("CovPostMean") = Cov, Rcpp::Named("kEst") = k);
}

// reference cpp:532-629 (list :620-627)
// [[Rcpp::export]]
Rcpp::List CPPFABLESampler(Rcpp::NumericMatrix Y, double gamma0, double delta0sq, int MC, Rcpp::NumericMatrix U_Y,
                           Rcpp::NumericMatrix V_Y, Rcpp::NumericVector svalsY, int kEst, double varInflation) {
  Svd s(Y, U_Y, V_Y, svalsY);
  if (MC < 1) Rcpp::stop("MC must be >= 1");
  if (kEst < 1 || kEst > s.r) Rcpp::stop("kEst must be in 1..length(svalsY)");
  reseed();
  Rcpp::NumericMatrix L(MC, static_cast<int>(kEst * s.p)), S(MC, static_cast<int>(s.p)), G(s.p, s.p);
  Rcpp::NumericVector pm(s.p), plug(s.p);
  double tau = 0.0;
  check(fable_sampler(ctx(), s.Y, s.n, s.p, gamma0, delta0sq, MC, s.U, s.V, s.d, s.r, kEst, varInflation,
                      /*gam_inj=*/nullptr, /*z_inj=*/nullptr, /*m0=*/0, L.begin(), S.begin(), pm.begin(), plug.begin(),
                      &tau, G.begin(), /*psi_sum=*/nullptr, /*psi_sumsq=*/nullptr));
  return Rcpp::List::create(Rcpp::Named("LambdaSamples") = L, Rcpp::Named("SigmaSqSamples") = S,
                            Rcpp::Named("SigmaSqEstimatePostMean") = pm, Rcpp::Named("SigmaSqEstimatePlugIn") = plug,
                            Rcpp::Named("EstimatedRank") = kEst, Rcpp::Named("VarianceInflation") = varInflation,
                            Rcpp::Named("tauSqEstimate") = tau, Rcpp::Named("G") = G);
}

// NEW: sampler + entry-wise posterior mean / second moment of Psi in one pass (the fusion of
// CPPFABLESampler with the mean of CPPCCFABLEPostProcessing, cpp:631-713, that never stores Psi draws)
// [[Rcpp::export]]
Rcpp::List FABLESamplerMoments(Rcpp::NumericMatrix Y, double gamma0, double delta0sq, int MC, Rcpp::NumericMatrix U_Y,
                               Rcpp::NumericMatrix V_Y, Rcpp::NumericVector svalsY, int kEst, double varInflation) {
  Svd s(Y, U_Y, V_Y, svalsY);
  reseed();
  Rcpp::NumericMatrix s1(s.p, s.p), s2(s.p, s.p);
  check(fable_sampler(ctx(), s.Y, s.n, s.p, gamma0, delta0sq, MC, s.U, s.V, s.d, s.r, kEst, varInflation, nullptr, nullptr,
                      0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, s1.begin(), s2.begin()));
  return Rcpp::List::create(Rcpp::Named("PsiSum") = s1, Rcpp::Named("PsiSumSq") = s2, Rcpp::Named("MC") = MC);
}

// reference cpp:631-713 and cpp:716-818 / 821-893 (f1): mean + (alpha/2, 1-alpha/2) sample quantiles
static Rcpp::List post_processing(Rcpp::List FABLEOutput, double alpha, const int64_t* sel, int64_t S) {
  Rcpp::NumericMatrix L = FABLEOutput["LambdaSamples"], Sg = FABLEOutput["SigmaSqSamples"];
  const int k = FABLEOutput["EstimatedRank"];
  const int64_t MC = Sg.nrow(), p = Sg.ncol();
  if (!sel) S = p;
  Rcpp::NumericMatrix M(S, S), Lo(S, S), Up(S, S);
  check(fable_post_processing(ctx(), L.begin(), Sg.begin(), MC, p, k, alpha, sel, S, M.begin(), Lo.begin(), Up.begin()));
  return Rcpp::List::create(Rcpp::Named("PostMeanMatrix") = M, Rcpp::Named("LowerQuantileMatrix") = Lo,
                            Rcpp::Named("UpperQuantileMatrix") = Up);
}

// [[Rcpp::export]]
Rcpp::List CPPCCFABLEPostProcessing(Rcpp::List FABLEOutput, double alpha) {
  return post_processing(FABLEOutput, alpha, nullptr, 0);
}

// [[Rcpp::export]]
Rcpp::List CCFABLEPostProcessingSubmatrix(Rcpp::List FABLEOutput, double alpha, Rcpp::IntegerVector SelectedIndices) {
  std::vector<int64_t> sel(SelectedIndices.begin(), SelectedIndices.end());      // 1-based, as in R (cpp:721)
  return post_processing(FABLEOutput, alpha, sel.data(), static_cast<int64_t>(sel.size()));
}

// [[Rcpp::export]]
Rcpp::List CCFABLEPostProcessingSubmatrix_Optimized(Rcpp::List FABLEOutput, double alpha, Rcpp::IntegerVector SelectedIndices) {
  return CCFABLEPostProcessingSubmatrix(FABLEOutput, alpha, SelectedIndices);
}

// reference cpp:40-74 (exported, unused by the wrappers).  flag 1 / 2: arma::svd, the full decomposition (U n x n, V p x p,
// :52-58); 3 / anything else: svd_econ (:60-66).  The LAPACK driver the flag also selected ("std" / "dc") has no counterpart.
// [[Rcpp::export]]
Rcpp::List CPPsvd(Rcpp::NumericMatrix X, int flag) {
  const int64_t n = X.nrow(), p = X.ncol(), r = n < p ? n : p;
  const bool full = flag == 1 || flag == 2;
  Rcpp::NumericMatrix U(n, full ? n : r), V(p, full ? p : r);
  Rcpp::NumericVector d(r);
  check((full ? fable_svd_full : fable_svd)(ctx(), X.begin(), n, p, U.begin(), d.begin(), V.begin()));
  return Rcpp::List::create(Rcpp::Named("U") = U, Rcpp::Named("d") = d, Rcpp::Named("V") = V);
}

// reference cpp:17-26: n Gamma(shape a, rate b) draws.  Not on the device path (the sampler draws its gammas on the GPU
// from the Philox stream); kept because exportPattern makes it public.  R's own generator, honouring set.seed().
// [[Rcpp::export]]
Rcpp::NumericVector rGamma(int n, double a, double b) { return Rcpp::rgamma(n, a, 1.0 / b); }

// reference cpp:28-37: A + A' with the diagonal halved.  Host loop: O(p^2) copies, no arithmetic worth a kernel.
// [[Rcpp::export]]
Rcpp::NumericMatrix SymmetrizeMatrix(Rcpp::NumericMatrix A) {
  const int n = A.nrow();
  if (A.ncol() != n) Rcpp::stop("SymmetrizeMatrix: A must be square");
  Rcpp::NumericMatrix out(n, n);
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) out(i, j) = i == j ? (A(i, i) + A(i, i)) / 2.0 : A(i, j) + A(j, i);
  return out;
}

// NEW: FABLEPosteriorSampler / FABLEPosteriorMean as pipelines with Y uploaded once (a11).  The R wrappers
// may call these instead of the export sequence; the returned lists have the wrappers' element names.
// [[Rcpp::export]]
Rcpp::List FABLEPosteriorSamplerB200(Rcpp::NumericMatrix Y, double gamma0, double delta0sq, double maxProp, int MC) {
  const int64_t n = Y.nrow(), p = Y.ncol(), r = n < p ? n : p;
  reseed();
  Rcpp::NumericMatrix U(n, r), V(p, r), hG(p, p), G(p, p);
  Rcpp::NumericVector d(r), hsig(p), pm(p), plug(p);
  int kMax = 0, k = 0;
  double htau = 0.0, vi = 0.0, tau = 0.0;
  // the resident row posterior is released on every path out of this function -- an R allocation failure or an
  // interrupt below would otherwise leak its device buffers for the rest of the session
  struct PostGuard {
    fable_posterior* p = nullptr;
    ~PostGuard() { fable_posterior_destroy(p); }
  } guard;
  check(fable_wrapper_sampler_begin(ctx(), Y.begin(), n, p, gamma0, delta0sq, maxProp, U.begin(), d.begin(), V.begin(), &kMax,
                                    &k, hsig.begin(), /*hypG0 (fetched below, once k is known)=*/nullptr, &htau, hG.begin(), &vi, &guard.p));
  Rcpp::NumericMatrix hG0(static_cast<int>(p), k);                                   // sized now that k is known
  check(fable_posterior_hyper_G0(guard.p, hG0.begin()));
  Rcpp::NumericMatrix L(MC, static_cast<int>(k * p)), S(MC, static_cast<int>(p));
  check(fable_posterior_sampler_outputs(guard.p, MC, 0, nullptr, nullptr, L.begin(), S.begin(), pm.begin(), plug.begin(),
                                        &tau, G.begin(), nullptr, nullptr));
  Rcpp::List samples = Rcpp::List::create(Rcpp::Named("LambdaSamples") = L, Rcpp::Named("SigmaSqSamples") = S,
                                          Rcpp::Named("SigmaSqEstimatePostMean") = pm, Rcpp::Named("SigmaSqEstimatePlugIn") = plug,
                                          Rcpp::Named("EstimatedRank") = k, Rcpp::Named("VarianceInflation") = vi,
                                          Rcpp::Named("tauSqEstimate") = tau, Rcpp::Named("G") = G);
  Rcpp::List hyp = Rcpp::List::create(Rcpp::Named("SigmaSqEstimate") = hsig, Rcpp::Named("G") = hG, Rcpp::Named("EstimatedRank") = k,
                                      Rcpp::Named("G0") = hG0, Rcpp::Named("tauSqEstimate") = htau);
  Rcpp::List svdY = Rcpp::List::create(Rcpp::Named("d") = d, Rcpp::Named("u") = U, Rcpp::Named("v") = V);
  return Rcpp::List::create(Rcpp::Named("CCFABLESamples") = samples, Rcpp::Named("FABLEHyperParameters") = hyp,
                            Rcpp::Named("svdY") = svdY, Rcpp::Named("estRank") = k, Rcpp::Named("varInflation") = vi);
}

// [[Rcpp::export]]
Rcpp::List FABLEPosteriorMeanB200(Rcpp::NumericMatrix Y, double gamma0, double delta0sq, double maxProp) {
  const int64_t n = Y.nrow(), p = Y.ncol(), r = n < p ? n : p;
  Rcpp::NumericMatrix U(n, r), V(p, r), Cov(p, p);
  Rcpp::NumericVector d(r);
  int k = 0, kMax = 0;
  check(fable_wrapper_posterior_mean(ctx(), Y.begin(), n, p, gamma0, delta0sq, maxProp, Cov.begin(), &k, &kMax, U.begin(),
                                     d.begin(), V.begin()));
  Rcpp::List svdY = Rcpp::List::create(Rcpp::Named("d") = d, Rcpp::Named("u") = U, Rcpp::Named("v") = V);
  return Rcpp::List::create(Rcpp::Named("FABLEPostMean") = Cov, Rcpp::Named("svdY") = svdY, Rcpp::Named("estRank") = k);
}

// NEW: drop-in for base-R svd(Y) as the wrappers use it (R/FABLE_wrapper.R:23-26, 90-93): list(d, u, v)
// [[Rcpp::export]]
Rcpp::List FABLEsvd(Rcpp::NumericMatrix Y) {
  const int64_t n = Y.nrow(), p = Y.ncol(), r = n < p ? n : p;
  Rcpp::NumericMatrix U(n, r), V(p, r);
  Rcpp::NumericVector d(r);
  check(fable_svd(ctx(), Y.begin(), n, p, U.begin(), d.begin(), V.begin()));
  return Rcpp::List::create(Rcpp::Named("d") = d, Rcpp::Named("u") = U, Rcpp::Named("v") = V);
}

// f3 -- the pure-R prototype path (R/FABLE_code.R).  The R-level functions RankEstimator(Ymat, svdmod, kMax) (:32-186) and
// CCFABLE_DirectSampler(Y, gamma0, delta0sq, MC) (:212-298) keep their signatures in rpkg/R/FABLE_code.R and call these.
// [[Rcpp::export]]
int RankEstimatorB200(Rcpp::NumericMatrix Ymat, Rcpp::NumericMatrix U_Y, Rcpp::NumericMatrix V_Y, Rcpp::NumericVector svalsY, int kMax) {
  Svd s(Ymat, U_Y, V_Y, svalsY);
  int k = 0;
  check(fable_rank_estimator_proto(ctx(), s.Y, s.n, s.p, s.U, s.V, s.d, s.r, kMax, &k));
  return k;
}

// returns the (MC, p, p) array CovSampleStor as a numeric vector with a dim attribute set on the R side
// [[Rcpp::export]]
Rcpp::NumericVector CCFABLEDirectSamplerB200(Rcpp::NumericMatrix Y, double gamma0, double delta0sq, int MC) {
  const int64_t n = Y.nrow(), p = Y.ncol();
  if (MC < 1) Rcpp::stop("MC must be >= 1");
  reseed();
  struct PostGuard {
    fable_posterior* p = nullptr;
    ~PostGuard() { fable_posterior_destroy(p); }
  } guard;
  int kMax = 0, k = 0;
  check(fable_direct_sampler_begin(ctx(), Y.begin(), n, p, gamma0, delta0sq, &kMax, &k, &guard.p));
  Rcpp::NumericVector out(static_cast<R_xlen_t>(MC) * p * p);
  check(fable_direct_sampler_draws(guard.p, MC, 0, nullptr, nullptr, out.begin()));
  return out;
}
