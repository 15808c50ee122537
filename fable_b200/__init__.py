"""fable_b200 -- B200 (sm_100a) drop-in for the estimation + sampling hot path of the FABLE R package.

The product is `libfable_b200.so` (C-ABI in `include/fable_b200.h`).  This Python package is the
host-side mirror of the reference's R interface (`R/FABLE_wrapper.R`, `R/RcppExports.R`) over that
ABI: same function names, argument meaning, return-list element names and error behaviour, so the
parity tests read like calls into the R package.  There is no CPU fallback: importing works
anywhere, but every call raises `FableError` when the library or a B200 is missing.
"""
from .api import (  # noqa: F401
    FableError, Context, MultiContext, Posterior,
    FABLEPosteriorMean, FABLEPosteriorSampler,
    CPPRankEstimator, CPPCriterionFunction, CPPBisectionRecursion, CPPLogLikelihoodEval,
    FABLEHyperParameters, CPPcov_correct_matrix, cov_correct_matrix, var_inflation,
    CPPFABLEPostMean, CPPFABLESampler, CPPsvd, svd, kmax_rule,
    CPPCCFABLEPostProcessing, CCFABLEPostProcessingSubmatrix, CCFABLEPostProcessingSubmatrix_Optimized,
    default_context, library_path, row_posterior,
    RankEstimator, RankEstimatorCriterion, CCFABLE_DirectSampler,
)
from . import synth  # noqa: F401

__all__ = [
    "FableError", "Context", "MultiContext", "Posterior", "FABLEPosteriorMean", "FABLEPosteriorSampler",
    "CPPRankEstimator", "CPPCriterionFunction", "CPPBisectionRecursion", "CPPLogLikelihoodEval", "FABLEHyperParameters",
    "CPPcov_correct_matrix", "cov_correct_matrix", "var_inflation", "CPPFABLEPostMean",
    "CPPFABLESampler", "CPPsvd", "svd", "kmax_rule", "CPPCCFABLEPostProcessing",
    "CCFABLEPostProcessingSubmatrix", "CCFABLEPostProcessingSubmatrix_Optimized", "default_context", "library_path", "row_posterior", "synth",
    "RankEstimator", "RankEstimatorCriterion", "CCFABLE_DirectSampler",
]
