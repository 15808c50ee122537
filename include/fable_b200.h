/* fable_b200.h -- C-ABI of libfable_b200.so: the B200 (sm_100a) drop-in for the estimation and
 * sampling hot path of the FABLE R package (shounakch/FABLE).
 *
 * Every entry point replaces one function the reference reaches through Rcpp's `.Call` shims
 * (reference `src/RcppExports.cpp:96-172`, registered at `:212-228`).  The reference side binds
 * these from its Rcpp translation unit (see INTEGRATION.md for the stub a maintainer adds).
 *
 * Conventions
 *   - plain C, `extern "C"`; no torch / Rcpp / Armadillo types cross this boundary.
 *   - every matrix is FP64 COLUMN-MAJOR (R / Armadillo memory order), leading dimension = rows.
 *   - `fable_*` functions take HOST pointers, never retain them past return, and write results
 *     straight into caller-allocated host memory (R vectors): one D2H copy, no intermediates.
 *   - `fable_dev_*` functions take DEVICE pointers (resident-in-HBM path used by bench.py, the
 *     multi-GPU drivers and the host layer itself).
 *   - return value: 0 = FABLE_OK, otherwise an error code; `fable_last_error(ctx)` returns the
 *     message.  The library never calls exit/abort and never calls back into R.
 *   - there is NO CPU fallback: every function fails with FABLE_ERR_CUDA when no sm_100 device
 *     is usable.
 *   - a context is bound to one device and is not thread-safe; use one per calling thread/GPU.
 */
#ifndef FABLE_B200_H
#define FABLE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FABLE_B200_ABI_VERSION 2

typedef enum {
  FABLE_OK = 0,
  FABLE_ERR_INVALID = 1,   /* bad dimension / rank / NULL pointer / non-finite scalar   */
  FABLE_ERR_CUDA = 2,      /* CUDA runtime error or no usable sm_100 device              */
  FABLE_ERR_NOMEM = 3,     /* device or pinned-host allocation failed                    */
  FABLE_ERR_NUMERIC = 4,   /* eigensolver did not converge / non-positive variance        */
  FABLE_ERR_INTERRUPT = 5  /* the caller's interrupt callback asked to stop              */
} fable_status;

typedef struct fable_ctx fable_ctx;             /* one device, one stream set, scratch pools */
typedef struct fable_posterior fable_posterior; /* device-resident row posterior + draw state  */

/* ---- context ------------------------------------------------------------------------------ */
int fable_abi_version(void);
/* device: CUDA ordinal.  seed: key of the counter-based Philox4x32-10 stream ("fable-philox-v1",
 * DESIGN.md); the Rcpp layer derives it from R's RNG inside RNGScope so set.seed() is honoured
 * (reference RNGScope: src/RcppExports.cpp:159). */
int fable_ctx_create(int device, uint64_t seed, fable_ctx** out);
void fable_ctx_destroy(fable_ctx* ctx);
const char* fable_last_error(const fable_ctx* ctx); /* ctx may be NULL: last create() failure */
int fable_ctx_set_seed(fable_ctx* ctx, uint64_t seed);
/* optional: polled between draw batches; non-zero return aborts with FABLE_ERR_INTERRUPT
 * (reference: Rcpp::checkUserInterrupt, updated-FABLE-functions.cpp:699). */
int fable_ctx_set_interrupt(fable_ctx* ctx, int (*poll)(void*), void* user);
/* number of kernels this context has launched since creation (bench.py "gpu_launches"). */
int64_t fable_ctx_launch_count(const fable_ctx* ctx);
/* Work buffers released by the entry points are kept in a per-device block cache (bounded by FABLE_B200_CACHE_GB,
 * default 40, flushed automatically when an allocation fails, emptied when the context is destroyed) so that
 * repeated calls do not pay cudaMalloc / cudaFree of multi-GB blocks.  This returns the idle blocks to the driver
 * now, e.g. before another library needs the memory.  (No reference counterpart: R's gc() plays this role.) */
int fable_ctx_trim_cache(fable_ctx* ctx);
/* ---- column-sharded estimation (multi-GPU: one context per GPU, each holding a block of the variables) ----
 * The reference has no distributed code; this is how its estimation path (svd -> kMax -> rank search -> row
 * posterior) shards when p is large: rank g passes only ITS columns of Y (n x p_local) and its rows of V, and the
 * library completes every sum over variables -- the Gram Y Y' of the SVD, the column norms and the sign convention
 * of V, sum_j log sigma_j^2 of the criterion (cpp:199), the tauSq mean (cpp:383/483/577) -- through `fn`, which
 * must sum `count` host doubles element-wise over all ranks IN PLACE with the same bits on every rank (an NCCL /
 * MPI all-reduce; return 0 on success).  `p_total` = all variables (used wherever the reference's formula has p),
 * `col_offset` = index of this rank's first column.  U, d, kMax, the criterion, the bisection and kEst come out
 * identical on every rank; V, G0, gnd, sigma^2 are the local rows.  p_total = 0 switches the shard off.
 * With a shard set, fable_svd (n <= p_total), fable_criterion, fable_bisection, fable_rank_estimator,
 * fable_hyper_parameters (G = NULL) and fable_row_posterior work on the local block; entry points that need every
 * pair of variables on one device (dense p x p outputs, the covariance pass) return FABLE_ERR_INVALID: gather the
 * O(p k) row posterior and build the sampler with fable_posterior_create_from_rows. */
typedef int (*fable_allreduce_sum_fn)(void* user, double* vals, int64_t count);
int fable_ctx_set_column_shard(fable_ctx* ctx, int64_t p_total, int64_t col_offset, int rank, int world,
                               fable_allreduce_sum_fn fn, void* user);
/* Optional, after fable_ctx_set_column_shard: sums of DEVICE-resident doubles -- the n x n Gram of the sharded SVD, the
 * one O(n^2) exchange of the path (200 MB at n = 5000) -- go through `fn` instead of being bounced through the host
 * callback: fn must sum `count` doubles at the device pointer over all ranks in place, ordered after the work already
 * queued on `cuda_stream` (the context's stream, a cudaStream_t) -- e.g. ncclAllReduce on that stream.  NULL: off. */
typedef int (*fable_allreduce_sum_dev_fn)(void* user, double* dev_vals, int64_t count, void* cuda_stream);
int fable_ctx_set_column_shard_dev(fable_ctx* ctx, fable_allreduce_sum_dev_fn fn, void* user);
int fable_ctx_synchronize(fable_ctx* ctx);
/* the context's main stream as a cudaStream_t cast to void* (for CUDA-event timing). */
void* fable_ctx_stream(fable_ctx* ctx);
/* Per-launch CUDA-event timing of the covariance kernel on the context's stream.  enable != 0
 * starts (and clears) the record; fable_ctx_kernel_time synchronises and returns the summed device
 * time in milliseconds and the number of launches recorded since it was enabled.  The time includes the kernels
 * that fold the launches' partial sums into the accumulators (they are part of the same pass), the count does not. */
int fable_ctx_kernel_timing(fable_ctx* ctx, int enable);
int fable_ctx_kernel_time(fable_ctx* ctx, double* total_ms, int64_t* launches);
/* Stage split of the single-upload wrapper pipelines (fable_wrapper_posterior_mean / fable_wrapper_sampler_begin =
 * R/FABLE_wrapper.R:80-128 / 12-54).  enable != 0 clears the record and makes the pipelines drain the stream at every
 * stage boundary (host clock); fable_ctx_stage_times copies the milliseconds of the last pipeline call:
 * [0] upload of Y, [1] Gram + eigensolver + back-multiplication (R: svd(Y)), [2] U, d, V to the host + kMax rule,
 * [3] rank search (cpp:281-341), [4] row posterior / hyper-parameters / varInflation, [5] p x p results to the host;
 * entries past count are not written, entries 6, 7 are zero.  (No reference counterpart: R's runTime is one number.) */
int fable_ctx_stage_timing(fable_ctx* ctx, int enable);
int fable_ctx_stage_times(const fable_ctx* ctx, double* ms, int count);
/* Diagnostics: register-resident issue rate of the FP64 tensor path (mma.sync.m8n8k4.f64 -> DMMA.8x8x4) on this
 * device, in TFLOP/s -- the denominator bench.py divides the Gram step by (no reference counterpart). */
int fable_ctx_measure_fp64_peak(fable_ctx* ctx, double* dmma_tflops);
/* slots > 0: fable_sampler's Psi pass also MATERIALISES every draw Psi_m (dense p x p) into a device
 * ring of `slots` matrices owned by the context (batches of min(slots, 32) draws, draw b of a batch in slot b), next to the sum / sumsq
 * accumulators -- the block-wise covariance formation named by the north star, left resident for
 * device-side consumers (fable_ctx_psi_ring).  slots == 0 (default): accumulate only. */
int fable_ctx_set_psi_ring(fable_ctx* ctx, int slots);
int fable_ctx_psi_ring(fable_ctx* ctx, double** dev_ring, int* slots);

/* ---- a1/a2: thin SVD of Y and the kMax rule ---------------------------------------------------
 * Replaces base-R `svd(Y)` inside the wrappers (reference R/FABLE_wrapper.R:23-26, 90-93) and the
 * exported-but-unused CPPsvd (updated-FABLE-functions.cpp:40-74).
 * Y n x p.  r = min(n,p).  U n x r, d r (descending), V p x r.  Gram (DMMA) + Jacobi eigensolve
 * + back-multiplication.  Signs: each V column has its largest-|entry| positive. */
int fable_svd(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, double* U, double* d, double* V);
/* CPPsvd(X, flag) with flag 1 / 2 (:52-58, arma::svd = the FULL decomposition): U n x n, d r, V p x p; flags 3 / 4 (svd_econ,
 * :60-66) are fable_svd.  `flag` also chose the LAPACK driver there ("std" / "dc"); there is one algorithm here. */
int fable_svd_full(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, double* U, double* d, double* V);
/* min{ i : cumsum(d)_i / sum(d) >= maxProp }, 1-based (R/FABLE_wrapper.R:27, 94). Host-only. */
int fable_kmax(const double* d, int64_t r, double maxProp, int* kMax);

/* ---- f4: CPPLogLikelihoodEval (updated-FABLE-functions.cpp:95-125; exported, unused by the wrappers) ----
 * -n sum_j log sqrt(SigmaSq_j) - 0.5 sum_ij ((Y - M Lambda')_ij / sqrt(SigmaSq_j))^2.  M n x k, Lambda p x k. */
int fable_log_likelihood_eval(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, const double* M,
                              const double* Lambda, int k, const double* SigmaSq, double* out);

/* ---- a3: CPPCriterionFunction (updated-FABLE-functions.cpp:127-213) ------------------------- */
int fable_criterion(fable_ctx* ctx, int kInd, const double* Y, int64_t n, int64_t p,
                    const double* U, const double* V, const double* d, int64_t r, double* out);
/* ---- a4: CPPBisectionRecursion (:215-271) -> (lower, upper) ---------------------------------- */
int fable_bisection(fable_ctx* ctx, int lower, int upper, const double* Y, int64_t n, int64_t p,
                    const double* U, const double* V, const double* d, int64_t r, int* lo, int* hi);
/* ---- a5: CPPRankEstimator (:281-341) --------------------------------------------------------- */
int fable_rank_estimator(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, const double* U,
                         const double* V, const double* d, int64_t r, int kMax, int* kEst);

/* ---- a6: FABLEHyperParameters (:343-399), UNSHRUNK G0 = YtU/sqrt(n) ---------------------------
 * sig p, G0 p x k, tausq scalar, G p x p (NULL to skip the dense p x p). */
int fable_hyper_parameters(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, const double* U,
                           const double* V, const double* d, int64_t r, int kEst, double* sig,
                           double* G0, double* tausq, double* G);

/* ---- a7: coverage-correction ---------------------------------------------------------------------
 * fable_cov_correct_matrix = CPPcov_correct_matrix (:401-428): upper-triangle vector, column-major
 * trimatu order, length p(p+1)/2, from sig (p) and a dense G (p x p).
 * fable_var_inflation: the scalar the wrappers derive from it, computed as a fused reduction that
 * never stores B.  G is given by its factor G0 (p x k, G = G0 G0').  variant 0 = wrapper formula
 * (sum over the FULL matrix / (p(p+1)/2))^2 (R/FABLE_wrapper.R:41-44); variant 1 = script formula
 * mean(upper triangle)^2 (extras/replicationCodes/runtimeScript.R:75-78). */
int fable_cov_correct_matrix(fable_ctx* ctx, const double* sig, const double* G, int64_t p,
                             double* Bupper);
int fable_cov_correct_full(fable_ctx* ctx, const double* sig, const double* G, int64_t p,
                           double* B); /* R twin cov_correct_matrix (R/FABLE_code.R:189-206) */
int fable_var_inflation(fable_ctx* ctx, const double* sig, const double* G0, int64_t p, int k,
                        int variant, double* out);

/* ---- a8: CPPFABLEPostMean (:440-519) ----------------------------------------------------------
 * Cov p x p (= CovPostMean).  Optional extra outputs (NULL to skip) expose the internals the
 * reference keeps private, for parity checks: G0s p x k (shrunk Lambda mean), SigmaHat p. */
int fable_post_mean(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, double gamma0,
                    double delta0sq, const double* U, const double* V, const double* d, int64_t r,
                    int kMax, double* Cov, int* kEst, double* G0s, double* SigmaHat);

/* ---- a9: CPPFABLESampler (:532-629) -------------------------------------------------------------
 * Outputs (any may be NULL to skip): LambdaSamples MC x (k p) [element (m, j*k+l) at m+MC*(j*k+l)],
 * SigmaSqSamples MC x p, SigmaPostMean p, SigmaPlugIn p, tausq, G p x p (= G0 G0', shrunk G0).
 * Draw source: gam_inj / z_inj non-NULL -> INJECTED draws in the reference's consumption order
 * (gam[m*p+j]; z[m*p*k + l*p + j]); both NULL -> native Philox stream, draw index m0+m.
 * psi_sum / psi_sumsq (p x p each, NULL to skip): entry-wise sum_m Psi_m and sum_m Psi_m^2 of
 * Psi_m = Lambda_m Lambda_m' + diag(sigma2_m) (entry formula :672-684; the mean is :688). */
int fable_sampler(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, double gamma0,
                  double delta0sq, int64_t MC, const double* U, const double* V, const double* d,
                  int64_t r, int kEst, double varInflation, const double* gam_inj,
                  const double* z_inj, int64_t m0, double* LambdaSamples, double* SigmaSqSamples,
                  double* SigmaPostMean, double* SigmaPlugIn, double* tausq, double* G,
                  double* psi_sum, double* psi_sumsq);

/* ---- a11: the two R wrappers as single calls -----------------------------------------------------------
 * FABLEPosteriorMean (R/FABLE_wrapper.R:80-128) and FABLEPosteriorSampler (:12-68) run svd(Y) -> kMax rule ->
 * CPP* exports, passing Y, U, V back and forth.  These entry points run the same sequence with Y uploaded
 * once and every intermediate resident in HBM; the Rcpp layer unpacks the outputs into the wrappers'
 * unchanged return lists.  U n x r, d r, V p x r (r = min(n,p)) are the `svdY` element; any output pointer
 * may be NULL to skip it (kEst is required).  varInflation follows the wrapper's formula (:41-44).
 * hyp* = FABLEHyperParameters' list (unshrunk G0; hypG0, if not NULL, must hold p x min(n,p) doubles, the first
 * kEst columns are written -- or pass NULL and call fable_posterior_hyper_G0 once kEst is known). */
int fable_wrapper_posterior_mean(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, double gamma0,
                                 double delta0sq, double maxProp, double* Cov, int* kEst, int* kMax,
                                 double* U, double* d, double* V);
/* FABLEPosteriorSampler in two steps, because R must size LambdaSamples (MC x k p) after the rank is
 * known: begin = wrapper:20-44 (svd, kMax, rank, FABLEHyperParameters, varInflation) and leaves the row
 * posterior of CPPFABLESampler (:553-597) resident; fable_posterior_sampler_outputs then produces
 * everything CPPFABLESampler returns (:599-627) -- same outputs as fable_sampler -- from that handle. */
int fable_wrapper_sampler_begin(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, double gamma0,
                                double delta0sq, double maxProp, double* U, double* d, double* V, int* kMax,
                                int* kEst, double* hypSigma, double* hypG0, double* hypTau, double* hypG,
                                double* varInflation, fable_posterior** post);
int fable_posterior_sampler_outputs(fable_posterior* post, int64_t MC, int64_t m0, const double* gam_inj,
                                    const double* z_inj, double* LambdaSamples, double* SigmaSqSamples,
                                    double* SigmaPostMean, double* SigmaPlugIn, double* tausq, double* G,
                                    double* psi_sum, double* psi_sumsq);
int fable_posterior_rank(fable_posterior* post, int* k);
/* FABLEHyperParameters$G0 (unshrunk, p x kEst) of a posterior made by fable_wrapper_sampler_begin: lets the caller size
 * the matrix after the rank is known (hypG0 = NULL in sampler_begin) instead of passing a p x min(n,p) scratch. */
int fable_posterior_hyper_G0(fable_posterior* post, double* hypG0);

/* ---- f1: CPPCCFABLEPostProcessing (:631-713) / CCFABLEPostProcessingSubmatrix(_Optimized) (:716-893) ---
 * Entry-wise posterior mean and (alpha/2, 1-alpha/2) sample quantiles (Armadillo quantile, type 5) of
 * Psi over the stored draws, on the variables `selected` (1-based R indices, length S; NULL = all p).
 * LambdaSamples MC x (k p), SigmaSqSamples MC x p as CPPFABLESampler returns them.  Outputs S x S. */
int fable_post_processing(fable_ctx* ctx, const double* LambdaSamples, const double* SigmaSqSamples,
                          int64_t MC, int64_t p, int k, double alpha, const int64_t* selected, int64_t S,
                          double* PostMean, double* Lower, double* Upper);

/* The same summaries from RESIDENT draws: draws m0 .. m0+MC-1 of a posterior object are generated on the device (native
 * stream, or injected gam / z in the reference's consumption order) and every entry of Psi_m is formed block-wise on
 * the fly -- no MC x k p sample matrix crosses PCIe, nothing of size p^2 MC is stored.  MC <= 65535, k <= 32.
 * Outputs S x S (selected = NULL: all p variables).  Both entry points use the tile-resident selection of
 * post_tiles.cu (exact order statistics by radix selection over recomputed values) wherever these limits hold. */
int fable_posterior_post_processing(fable_posterior* post, int64_t MC, int64_t m0, double alpha, const int64_t* selected,
                                    int64_t S, const double* gam_inj, const double* z_inj, double* PostMean,
                                    double* Lower, double* Upper);

/* ---- f3: the pure-R prototype path CCFABLE_DirectSampler (R/FABLE_code.R:212-298) ------------------------------
 * The prototype picks the rank with its own RankEstimator (R:32-186): the same bisection and window as the C++ one over
 * a different criterion -- sigma_j^2 = resid_j / (n - k) (:87), log-likelihood and penalty scaled by 1 / (n p) (:106-114).
 * fable_criterion_proto / fable_rank_estimator_proto are that criterion and rule at the CPP* boundary (U, V, d given).
 * fable_direct_sampler_begin = R:214-263 (svd, kMax at 0.95, that rank rule, the shrunk row posterior, a_n without
 * variance inflation, the entry-wise correction B = cov_correct_matrix(sigma^2, G) on); fable_direct_sampler_draws = the
 * MC loop (:267-294): CovSampleStor, an (MC, p, p) array with the FIRST index fastest (element (m, u, v) at
 * m + MC (u + p v)), draws m0 .. m0+MC-1 of the native stream or injected gam / z.  MC p^2 doubles must stay below
 * 32 GB (the prototype's own comments send larger problems to the sampler + post-processing route, :209-211). */
int fable_criterion_proto(fable_ctx* ctx, int kInd, const double* Y, int64_t n, int64_t p, const double* U,
                          const double* V, const double* d, int64_t r, double* out);
int fable_rank_estimator_proto(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, const double* U, const double* V,
                               const double* d, int64_t r, int kMax, int* kEst);
int fable_direct_sampler_begin(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, double gamma0, double delta0sq,
                               int* kMax, int* kEst, fable_posterior** post);
int fable_direct_sampler_draws(fable_posterior* post, int64_t MC, int64_t m0, const double* gam_inj, const double* z_inj,
                               double* CovSampleStor);

/* ---- resident-in-HBM posterior object (what fable_sampler is built from) ---------------------- */
/* Uploads Y, U_k, V_k, d_k, runs the row-posterior kernels, keeps G0 / gnd / a_n on the device. */
int fable_posterior_create(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, double gamma0,
                           double delta0sq, const double* U, const double* V, const double* d,
                           int64_t r, int kEst, double varInflation, fable_posterior** out);
/* The per-variable conjugate posterior by itself: cpp:371-391 (shrunk = 0: G0 = YtU / sqrt(n), the
 * FABLEHyperParameters form) or cpp:470-499 = 564-592 (shrunk = 1: G0 = sqrt(n)/(n + 1/tauSq) YtU, gnd, the
 * CPPFABLEPostMean / CPPFABLESampler form).  Outputs: G0 p x k, gnd p, sig = sigma^2 plug-in (cpp:379), sigmahat =
 * gnd / (gamma_n - 2) (cpp:509), tauSq.  Under a column shard: the local rows, tauSq global. */
int fable_row_posterior(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, double gamma0, double delta0sq,
                        const double* U, const double* V, const double* d, int64_t r, int k, int shrunk, double* G0,
                        double* gnd, double* sig, double* sigmahat, double* tausq);
/* The same resident object from an already computed (e.g. gathered from column shards) row posterior: G0 p x k
 * (shrunk form), gnd, sig, sigmahat p, tauSq.  Y is not needed: a_n = sqrt(varInflation) / sqrt(n + 1/tauSq). */
int fable_posterior_create_from_rows(fable_ctx* ctx, int64_t n, int64_t p, int k, double gamma0, double delta0sq,
                                     const double* G0, const double* gnd, const double* sig, const double* sigmahat,
                                     double tausq, double varInflation, fable_posterior** out);
void fable_posterior_destroy(fable_posterior* post);
/* copy-out of the row posterior (NULL to skip): G0s p x k, gnd p, sig p, scalars[4] =
 * {tausq, a_n, gamma_n, w = n + 1/tausq}. */
int fable_posterior_get(fable_posterior* post, double* G0s, double* gnd, double* sig, double* scalars);
/* One batch of B draws m0..m0+B-1 entirely on the device:
 *   sampler kernel -> Lambda_m, sigma2_m (working layout) -> covariance kernel.
 * dev_psi: device buffer of `psi_slots` >= B dense p x p matrices (NULL = do not materialise); draw m
 * lands in slot m - m0 (tiles are processed batch-major, so every draw of a launch needs its own slot).
 * accumulate != 0 adds the batch into the posterior's own sum / sumsq accumulators (compact upper tiles, see
 * "tile layout" below).  gam/z: injected DEVICE buffers or NULL. */
int fable_posterior_draw_batch(fable_posterior* post, int64_t m0, int B, double* dev_psi,
                               int psi_slots, int accumulate, const double* dev_gam,
                               const double* dev_z);
/* The same batch with the draws materialised COMPACTLY for the posterior's tile slice (multi-GPU partition B: a rank
 * stores only what it owns).  dev_psi_tiles: `psi_slots` >= B slots of 2 * tiles * 4096 doubles, tiles =
 * fable_tile_count(p, t_begin, t_end); in a slot, the slice's tile number c (compact order) is stored as the pair
 * (block (I, J) of Psi_m at 2c, its mirror image block (J, I) at 2c + 1), each a column-major 64 x 64 block -- the
 * rank's share of the dense symmetric matrix, 8 p^2 / world bytes per draw.  Works for any p (edge tiles are padded
 * with zeros).  fable_posterior_tiles_to_dense(.., paired = 1, ..) expands a slot. */
int fable_posterior_draw_batch_tiles(fable_posterior* post, int64_t m0, int B, double* dev_psi_tiles,
                                     int psi_slots, int accumulate, const double* dev_gam,
                                     const double* dev_z);
/* f3: CCFABLE_DirectSampler (R/FABLE_code.R:212-298).  on != 0: the draws materialised by draw_batch become
 *   Psi_m = G + B o (Lambda_m Lambda_m' - G) + diag(sigma2_m),  B = cov_correct_matrix(sig_hat, G)  (R:251, 282-290),
 * the entry-wise coverage correction of the R prototype, with G = G0 G0' (shrunk G0) and B formed per tile
 * on the fly (never stored).  The prototype uses a_n = 1/sqrt(n + 1/tau^2): create the posterior with
 * varInflation = 1.  Accumulators, when requested, sum the corrected draws. */
int fable_posterior_set_entrywise_correction(fable_posterior* post, int on);
/* ---- tile layout and multi-GPU partition B (tiles) ----------------------------------------------------------
 * The covariance pass cuts Psi into 64 x 64 tiles and computes the tiles (I, J), I <= J -- the reference's pair loop
 * j1 <= j2 (updated-FABLE-functions.cpp:650, 662) -- enumerated in 16 x 16 super-blocks (super-blocks over their own
 * upper triangle, column-major; tiles column-major inside).  The enumeration has fable_tile_slots(p) slots, padding
 * slots (below the diagonal of diagonal super-blocks, past the edge) included; storage is COMPACT: valid tiles only, in
 * enumeration order, 4096 doubles (a column-major 64 x 64 block) each.  A slice [t_begin, t_end) of the enumeration
 * therefore owns a contiguous range of compact tiles (fable_tile_count of them): the sum / sumsq accumulators of a
 * posterior hold exactly the upper tiles of its slice.
 * fable_posterior_set_tile_range restricts draw_batch / draw_batch_tiles to a slice (t_end == 0: everything) and
 * drops the accumulators of the previous slice.  Ranks owning disjoint slices and processing the same draws produce
 * disjoint parts of every Psi_m and of the accumulators: no collective is needed (the reference's answer to large p
 * is likewise to work on sub-matrices, :716-818).  fable_tile_partition gives rank `rank` of `world` a slice with an
 * equal share of the valid tiles; fable_tile_locate the slot and compact index (counted from the start of the
 * enumeration) of tile (I, J). */
int fable_tile_slots(int64_t p, int64_t* slots);
int fable_tile_count(int64_t p, int64_t t_begin, int64_t t_end, int64_t* tiles);
int fable_tile_partition(int64_t p, int rank, int world, int64_t* t_begin, int64_t* t_end);
int fable_tile_locate(int64_t p, int64_t I, int64_t J, int64_t* slot, int64_t* compact_index);
int fable_posterior_tile_slots(fable_posterior* post, int64_t* slots);
int fable_posterior_set_tile_range(fable_posterior* post, int64_t t_begin, int64_t t_end);
/* compact tiles of the posterior's slice -> columns [col0, col0 + ncols) of the dense symmetric p x p matrix (DEVICE
 * pointers; dev_out is p x ncols with leading dimension p; col0 a multiple of 64).  paired = 0: upper tiles only (the
 * accumulators), the lower triangle is the transpose; paired = 1: (direct, mirrored) pairs as draw_batch_tiles writes
 * them.  Blocks outside the slice are written as zeros, so the panels of the ranks of a partition add up. */
int fable_posterior_tiles_to_dense(fable_posterior* post, const double* dev_tiles, int paired, int64_t col0,
                                   int64_t ncols, double* dev_out);
/* reference-layout samples for draws m0..m0+MC-1 into HOST arrays with leading dimension ldm
 * (>= MC; rows m-m0) -- NULL to skip either. */
int fable_posterior_samples(fable_posterior* post, int64_t m0, int64_t MC, int64_t ldm,
                            double* LambdaSamples, double* SigmaSqSamples,
                            const double* gam_inj, const double* z_inj);
/* The draw loop of CPPFABLESampler (:599-618) for draws m0 .. m0 + MC - 1: the reference-layout samples go to HOST
 * arrays with leading dimension ldm (rows 0 .. MC-1; NULL to skip), and with psi_pass != 0 the covariance pass adds
 * the same draws into the device-resident accumulators (no reset, nothing p x p crosses PCIe).  This is what one
 * rank of multi-GPU partition A runs for its share of the MC axis; the accumulators are then summed on the device
 * (fable_posterior_accum_dev + NCCL) and fetched once by whoever needs them.  fable_posterior_sampler_outputs is
 * accum_reset + this + accum_fetch + G. */
int fable_posterior_sample_range(fable_posterior* post, int64_t MC, int64_t m0, int64_t ldm, const double* gam_inj,
                                 const double* z_inj, double* LambdaSamples, double* SigmaSqSamples, int psi_pass);
/* the elements of CPPFABLESampler's list that do not depend on the draws (:622-627); any pointer may be NULL */
int fable_posterior_fixed_outputs(fable_posterior* post, double* SigmaPostMean, double* SigmaPlugIn, double* tausq,
                                  double* G);
/* draw_batch restricted to the tiles [t_begin, t_end) of the posterior's slice (same storage, fewer tiles per launch), and
 * the dense blocks of whole super-blocks (256 consecutive slots = 16 x 16 tiles; block and mirror image) of the
 * accumulators into the caller's dense p x p host matrices, everything else left untouched: the pieces
 * fable_sampler's tail is made of -- the last draws run chunk-major so that finished chunks cross PCIe while the
 * following ones compute.  fetch_superblocks runs on the stream the context currently uses and returns when the host
 * matrices hold the blocks. */
int fable_posterior_draw_batch_range(fable_posterior* post, int64_t m0, int B, double* dev_psi, int psi_slots,
                                     int accumulate, const double* dev_gam, const double* dev_z, int64_t t_begin,
                                     int64_t t_end);
int fable_posterior_accum_fetch_superblocks(fable_posterior* post, int64_t sb_begin, int64_t sb_end, double* sum,
                                            double* sumsq);
/* accumulators: device pointers to the compact upper tiles of the slice (count doubles each = tiles * 4096; the
 * one NCCL sum of multi-GPU partition A -- draws sharded over ranks -- reduces these buffers as they are), reset,
 * fetch as dense symmetric p x p host matrices (zeros outside the slice), fetch as raw compact tiles.
 * (Materialising batches keep their sums in a few partial buffers first -- no DRAM reads inside the write-bound
 * covariance kernel -- which every one of these calls folds into the accumulators before it looks at them.) */
int fable_posterior_accum_dev(fable_posterior* post, double** dev_sum, double** dev_sumsq, int64_t* count);
int fable_posterior_accum_reset(fable_posterior* post);
int fable_posterior_accum_fetch(fable_posterior* post, double* sum, double* sumsq);            /* host out */
int fable_posterior_accum_fetch_tiles(fable_posterior* post, double* sum_tiles, double* sumsq_tiles);

/* ---- multi-GPU behind the boundary: one process (the R session), several devices of one box ------------------------
 * The reference has no multi-device code; its MC loop (:599-618) is embarrassingly parallel over draws.
 * fable_multi_sampler is CPPFABLESampler (:532-629; same arguments and outputs as fable_sampler) on the draw partition
 * (SURVEY 8(e) A): device g draws a contiguous share of the MC axis with the stream one device would use (the union
 * of the rows is bit-identical to fable_sampler), writes its rows straight into the caller's LambdaSamples /
 * SigmaSqSamples, and the Psi sum / sumsq accumulators are summed by a hand-written kernel over NVLink peer memory
 * (device g reduces slice g of the compact tiles over all devices into device 0's buffers; no library collective)
 * before device 0 alone downloads the dense p x p sums.  devices = NULL: ordinals 0 .. ndev-1.  Needs peer access
 * between the devices (NVLink / NVSwitch); there is no host-staged fallback. */
typedef struct fable_multi fable_multi;
int fable_multi_create(int ndev, const int* devices, uint64_t seed, fable_multi** out);
void fable_multi_destroy(fable_multi* m);
const char* fable_multi_last_error(const fable_multi* m);   /* m may be NULL: last create() failure */
int fable_multi_devices(const fable_multi* m);
fable_ctx* fable_multi_ctx(fable_multi* m, int g);          /* the per-device contexts (owned by m) */
int fable_multi_sampler(fable_multi* m, const double* Y, int64_t n, int64_t p, double gamma0, double delta0sq,
                        int64_t MC, const double* U, const double* V, const double* d, int64_t r, int kEst,
                        double varInflation, const double* gam_inj, const double* z_inj, int64_t m0,
                        double* LambdaSamples, double* SigmaSqSamples, double* SigmaPostMean, double* SigmaPlugIn,
                        double* tausq, double* G, double* psi_sum, double* psi_sumsq);

/* ---- device-pointer primitives ------------------------------------------------------------------ */
/* out[u + p*v] = sum_l A[u + lda*l] A[v + lda*l] + (u==v ? s[u] : 0), dense symmetric p x p
 * (serves G = G0 G0' :391/:491/:627 and CovPostMean :512).  s may be NULL. */
int fable_dev_rankk_cov(fable_ctx* ctx, const double* dA, int64_t lda, const double* ds, int64_t p,
                        int k, double* dOut);
/* C (M x N, ldc) = alpha * op(A) op(B)':  A is M x K, B is N x K.  a_kfast / b_kfast: the K index
 * is the contiguous one of that operand (i.e. it is stored K x M with leading dimension ld).
 * FP64 DMMA tensor-core GEMM (Gram YY' / Y'Y, V = Y'U D^-1, rank-k reconstruction). */
int fable_dev_gemm(fable_ctx* ctx, int a_kfast, int b_kfast, int64_t M, int64_t N, int64_t K,
                   double alpha, const double* dA, int64_t lda, const double* dB, int64_t ldb,
                   double* dC, int64_t ldc);
/* C (M x M dense, both triangles) = alpha * A A', A: M x K stored K-slow (column-major M x K, lda) or K-fast
 * (a_kfast: K x M with leading dimension lda).  Only the upper tile triangle is computed on the FP64
 * tensor cores (SYRK: M(M+1)K flops), then mirrored: the Gram step Y Y' / Y'Y of fable_svd. */
int fable_dev_syrk(fable_ctx* ctx, int a_kfast, int64_t M, int64_t K, double alpha, const double* dA,
                   int64_t lda, double* dC);
/* symmetric eigen-decomposition of the m x m matrix dG (destroyed): eigenvalues descending in
 * dW (m), eigenvectors in the columns of dQ (m x m). */
int fable_dev_syevj(fable_ctx* ctx, double* dG, int64_t m, double* dW, double* dQ, int* sweeps);
/* Row pass of the conjugate posterior / rank criterion on resident inputs (cpp:371-379 == :470-479 == :564-573,
 * and the residual of CPPCriterionFunction :161-166): for every variable j, with c = YtU (p x k, ldc),
 *   resid[j] = sum_i (Y[i,j] - sum_l U[i,l] c[j,l])^2      (NOT divided by n)
 *   s[j]     = sum_i Y[i,j]^2                              (s may be NULL)
 * One streaming read of Y (8 n p bytes): the HBM-read-bound kernel of SURVEY 8(d) "Row-posterior". */
int fable_dev_rowstats(fable_ctx* ctx, const double* dY, int64_t n, int64_t p, const double* dU,
                       const double* dC, int64_t ldc, int k, double* ds, double* dresid);

#ifdef __cplusplus
}
#endif
#endif /* FABLE_B200_H */
