# The two functions of the reference's pure-R prototype layer (R/FABLE_code.R) that this port serves from the device, with
# their signatures and return values unchanged:
#   RankEstimator(Ymat, svdmod, kMax)                     reference :32-186  (its own criterion: sigma^2 = resid / (n - k),
#                                                         log-likelihood and penalty scaled by 1 / (n p))
#   CCFABLE_DirectSampler(Y, gamma0, delta0sq, MC)        reference :212-298 (entry-wise coverage-corrected draws of Psi,
#                                                         an array of dim c(MC, p, p))
# The wrapper FABLEPosteriorSampler no longer needs cov_correct_matrix() (reference :189-206; see FABLE_wrapper.R); it
# stays available through CPPcov_correct_matrix / FABLEVarInflation.  The remaining prototypes of that file
# (FABLEPostmean, FABLESampler, CCFABLEPostProcessing) are superseded by the C++ exports in the reference itself and are
# not carried over.

RankEstimator <- function(Ymat, svdmod, kMax) {
  RankEstimatorB200(as.matrix(Ymat), svdmod$u, svdmod$v, svdmod$d, as.integer(kMax))
}

CCFABLE_DirectSampler <- function(Y, gamma0 = 1, delta0sq = 1, MC = 1000) {
  Y <- as.matrix(Y)
  p <- ncol(Y)
  out <- CCFABLEDirectSamplerB200(Y, gamma0, delta0sq, as.integer(MC))
  dim(out) <- c(MC, p, p)
  out
}
