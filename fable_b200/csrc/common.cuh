// Shared host/device plumbing of libfable_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/fable_b200.h"

struct fable_ctx {
  int device = 0;
  int sms = 148;
  uint64_t seed = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  std::string err;
  int64_t launches = 0;
  int (*poll)(void*) = nullptr;
  void* poll_user = nullptr;
  // optional per-kernel CUDA-event timing of the covariance kernel (bench.py roofline)
  bool timing = false;
  // optional stage split of the wrapper pipelines (fable_ctx_stage_timing): host clock, the stream drained at every mark
  const char* prefault_lo[4] = {nullptr, nullptr, nullptr, nullptr};   // pageable result ranges being touched by HostPrefault (api.cu)
  const char* prefault_hi[4] = {nullptr, nullptr, nullptr, nullptr};
  bool stage_timing = false;
  double stage_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::vector<cudaEvent_t> ev_pool;     // pairs: [2i] before, [2i+1] after the launch
  size_t ev_used = 0;
  std::vector<cudaEvent_t> ev_extra;    // pairs around work done on behalf of those launches (folding their partial sums): counted in the total time, not as launches
  // optional device ring of materialised Psi draws kept by fable_sampler (fable_ctx_set_psi_ring)
  double* gemm_ws = nullptr;    // grow-only split-K workspace of the DMMA GEMM
  size_t gemm_ws_bytes = 0;
  void* stage[2] = {nullptr, nullptr};          // pinned staging for D2H into pageable caller memory
  cudaEvent_t stage_ev[2] = {nullptr, nullptr};
  int ring_slots = 0;
  double* ring = nullptr;
  size_t ring_bytes = 0;
  // column shard (fable_ctx_set_column_shard): this context holds columns [shard_off, shard_off + p_loc) of a
  // p_total-column problem; sums over columns go through the caller's all-reduce
  int64_t shard_p = 0, shard_off = 0;
  int shard_rank = 0, shard_world = 1;
  int (*shard_fn)(void*, double*, int64_t) = nullptr;
  void* shard_user = nullptr;
  int (*shard_dev_fn)(void*, double*, int64_t, void*) = nullptr;    // optional: sums of DEVICE doubles (the n x n Gram)
  void* shard_dev_user = nullptr;

  int fail(int code, const std::string& msg) {
    err = msg;
    return code;
  }
};

#define FB_CUDA(ctx, expr)                                                                       \
  do {                                                                                           \
    cudaError_t e__ = (expr);                                                                    \
    if (e__ != cudaSuccess) {                                                                    \
      (void)cudaGetLastError();                                                                  \
      return (ctx)->fail(e__ == cudaErrorMemoryAllocation ? FABLE_ERR_NOMEM : FABLE_ERR_CUDA,    \
                         std::string(#expr) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + \
                             ":" + std::to_string(__LINE__) + ")");                              \
    }                                                                                            \
  } while (0)

#define FB_TRY(expr)               \
  do {                             \
    int rc__ = (expr);             \
    if (rc__ != FABLE_OK) return rc__; \
  } while (0)

#define FB_REQUIRE(ctx, cond, msg)                                   \
  do {                                                               \
    if (!(cond)) return (ctx)->fail(FABLE_ERR_INVALID, std::string(msg)); \
  } while (0)

// every kernel launch goes through this so bench.py can report "gpu_launches".
#define FB_LAUNCHED(ctx)                 \
  do {                                   \
    ++(ctx)->launches;                   \
    FB_CUDA(ctx, cudaGetLastError());    \
  } while (0)

// Device block cache behind DevBuf (api.cu): the entry points allocate the same multi-GB work buffers on every
// call (accumulators, sample layouts, residuals), and cudaMalloc / cudaFree of such blocks map and unmap device
// memory synchronously -- tens to hundreds of milliseconds per call, with outliers.  Released blocks are therefore
// kept per device and handed out again to requests of the same rounded size; the cache is bounded
// (FABLE_B200_CACHE_GB, default 40), flushed when an allocation fails, and emptied by fb_mem_trim().
cudaError_t fb_mem_acquire(void** ptr, size_t bytes);
void fb_mem_release(void* ptr);
void fb_mem_trim(int device);      // device < 0: every device
void fb_mem_register_stream(int device, cudaStream_t s, bool add);   // streams a release records its events on

// RAII device buffer of doubles (or bytes via raw()).
struct DevBuf {
  void* ptr = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : ptr(o.ptr), bytes(o.bytes) { o.ptr = nullptr; o.bytes = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); ptr = o.ptr; bytes = o.bytes; o.ptr = nullptr; o.bytes = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (ptr) fb_mem_release(ptr);
    ptr = nullptr; bytes = 0;
  }
  cudaError_t alloc(size_t nbytes) {
    release();
    if (nbytes == 0) nbytes = 16;
    cudaError_t e = fb_mem_acquire(&ptr, nbytes);
    if (e == cudaSuccess) bytes = nbytes; else ptr = nullptr;
    return e;
  }
  double* d() const { return static_cast<double*>(ptr); }
};

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }

// ---- internal device-level entry points shared between translation units -------------------
// rows.cu
// d_c = YtU (p x k, ldc); d_s (may be NULL) = column sums of squares; d_resid = RAW residual sums.
int fb_rowstats(fable_ctx* ctx, const double* dY, int64_t n, int64_t p, int64_t ldy, const double* dU,
                int64_t ldu, const double* d_c, int64_t ldc, int k, double* d_s, double* d_resid);
int fb_colsumsq(fable_ctx* ctx, const double* dY, int64_t n, int64_t p, int64_t ldy, double* d_s);
int fb_cq(fable_ctx* ctx, const double* dV, int64_t ldv, const double* dd, int64_t p, int k,
          double* d_c, int64_t ldc, double* d_q);
// reduce.cu
int fb_sum_log_scaled(fable_ctx* ctx, const double* d_x, int64_t p, double scale, double* host_out);
// all-rank criterion pieces: host_out[k'-1] = sum_j log((s_j - sum_{l <= k'} (d_l V_jl)^2) / n), k' = 1 .. kMax;
// max |T - V diag(d)| / max |V diag(d)| over p x k; max |A - I| over k x k
int fb_prefix_sumlog(fable_ctx* ctx, const double* dV, int64_t ldv, const double* dd, const double* ds, int64_t p, int kMax,
                     int64_t n, double* host_out);
int fb_svd_consistency(fable_ctx* ctx, const double* dT, int64_t ldt, const double* dV, int64_t ldv, const double* dd, int64_t p,
                       int k, double* max_diff, double* max_abs);
int fb_identity_err(fable_ctx* ctx, const double* dA, int64_t k, double* max_diff);
int fb_resid_over_scaled_self_sum(fable_ctx* ctx, const double* d_resid, int64_t p, double inv, double* host_sum);   // sum_j r_j / (r_j * inv)
int fb_loglik_terms(fable_ctx* ctx, const double* d_resid, const double* d_sig, int64_t p, double* sum_log_sd,
                    double* sum_resid_over_sig);
int fb_q_over_sig_sum(fable_ctx* ctx, const double* d_q, const double* d_resid, int64_t n, int64_t p, double* host_sum);
int fb_tausq(fable_ctx* ctx, const double* d_q, const double* d_resid, int64_t n, int64_t p, int k,
             double* host_tausq);
int fb_finish_posterior(fable_ctx* ctx, const double* d_c, int64_t ldc, const double* d_s,
                        const double* d_q, const double* d_resid, int64_t n, int64_t p, int k,
                        double gamma0, double delta0sq, double tausq, int shrunk, double* d_G0,
                        int64_t ldg, double* d_gnd, double* d_sig, double* d_sigmahat);
int fb_gdiag(fable_ctx* ctx, const double* d_G0, int64_t ldg, int64_t p, int k, double* d_gdiag);
int fb_var_inflation(fable_ctx* ctx, const double* d_sig, const double* d_G0, int64_t ldg, int64_t p,
                     int k, double* host_sum_upper, double* host_sum_diag);
int fb_cov_correct_dense(fable_ctx* ctx, const double* d_sig, const double* d_G, int64_t p,
                         double* d_out, int upper_vector);
int fb_sum_partials(fable_ctx* ctx, const double* d_part, int64_t nparts, int64_t len, int64_t stride,
                    double alpha, double* d_out);
// gemm.cu
int fb_gemm(fable_ctx* ctx, int a_kfast, int b_kfast, int64_t M, int64_t N, int64_t K, double alpha,
            const double* dA, int64_t lda, const double* dB, int64_t ldb, double* dC, int64_t ldc);
int fb_syrk(fable_ctx* ctx, int a_kfast, int64_t M, int64_t K, double alpha, const double* dA, int64_t lda,
            double* dC, int64_t ldc);
int fb_dmma_peak(fable_ctx* ctx, double* tflops);     // measured DMMA.8x8x4 issue peak of this device
// residual column sums of squares: d_resid[j] = sum_i (Y[i,j] - sum_l U[i,l] C[j,l])^2, l < k.
int fb_gemm_resid(fable_ctx* ctx, int64_t n, int64_t p, int k, const double* dY, int64_t ldy,
                  const double* dU, int64_t ldu, const double* dC, int64_t ldc, double* d_resid);
// eig.cu
// eigenvalues at the noise floor (<= ~16 m eps lambda_1: exactly singular G) come back as 0 with their vectors replaced by
// an orthonormal completion; *deflated (may be NULL) = their number
int fb_syevj(fable_ctx* ctx, double* dG, int64_t m, double* dW, double* dQ, int* sweeps, int64_t* deflated = nullptr);
// columns [done, total) of Q (rows x total, ld) <- unit vectors orthogonal to columns [0, done) and to each other
int fb_complete_orthonormal(fable_ctx* ctx, double* dQ, int64_t rows, int64_t ld, int64_t done, int64_t total);
// svd.cu: thin SVD of the device matrix dY (n x p, ld n): dU n x r, dd r, dV p x r (dense lds).
int fb_svd(fable_ctx* ctx, const double* dY, int64_t n, int64_t p, double* dU, double* dd, double* dV);
// column-wise pieces of the back-multiplication, for the column-sharded SVD: sums of squares, (max |x|, first row
// index of it, the signed value there) and x[:, l] *= scale[l]
int fb_col_sumsq(fable_ctx* ctx, const double* dT, int64_t rows, int64_t ld, int64_t cols, double* d_out);
int fb_col_absmax(fable_ctx* ctx, const double* dT, int64_t rows, int64_t ld, int64_t cols, double* d_max, double* d_idx,
                  double* d_val);
int fb_col_scale(fable_ctx* ctx, double* dT, int64_t rows, int64_t ld, int64_t cols, const double* d_scale);
// psi.cu
struct PsiArgs {
  const double* Lw;    // working-layout draws: Lw[(b*KP + l)*pp + j]
  const double* s2w;   // sigma^2 draws: s2w[b*pp + j]   (may be NULL: no diagonal term)
  int64_t pp;          // padded row count (multiple of the tile edge)
  int64_t p;
  int k, KP;           // rank and its zero-padded size (multiple of 4)
  int B;               // draws in this batch
  double* psi;         // materialised output (NULL: none): dense p x p per slot, or (out_compact) the tiles of the slice
  int psi_slots;
  int slot0;           // draw b goes to slot slot0 + b (< psi_slots)
  bool out_compact = false;   // psi holds, per slot, the slice's tiles compactly: direct tile c at 2c, mirrored at 2c + 1
  double* acc_sum;     // accumulators (NULL: none): COMPACT upper tiles of the slice (tiles.cuh), 4096 doubles per tile
  double* acc_sq;
  bool acc_fresh = false;     // materialising kernels: do not load the running tiles -- store this batch's own sums (the
                              // pointers then name a PARTIAL buffer that fb_fold_partials adds into the accumulators later)
  const double* cc_gdiag = nullptr;   // entry-wise coverage correction (f3): diag(G) and sig_hat (length p);
  const double* cc_sig = nullptr;     // Lw then carries G0 in slot 0 ahead of the B draws
  int64_t store_begin = 0, store_end = 0; // slice of the tile enumeration the compact buffers cover (store_end == 0: all tiles)
  int64_t tile_begin = 0, tile_end = 0;   // tiles this launch processes, inside the storage slice (tile_end == 0: the whole slice)
  int tma_mode = 0;    // set by fb_psi (materialised output): 1 = 4-D whole-tile TMA boxes, 2 = 3-D sub-boxes, 0 = plain stores
  int nT = 0;          // set by fb_psi: tiles per side
  int64_t cbase = 0;   // set by fb_psi: compact index of the slice's first tile
};
int fb_psi(fable_ctx* ctx, const PsiArgs& a);
// acc[i] += sum_{s < np} part[s * stride + i], i in [begin, end): the batches' partial sums folded into an accumulator
int fb_fold_partials(fable_ctx* ctx, double* d_acc, const double* d_part, int np, int64_t stride, int64_t begin, int64_t end);
// compact tiles of the slice [t_lo, t_hi) (t_hi == 0: all) -> columns [col0, col0 + ncols) of the dense symmetric p x p
// matrix (d_out: p x ncols, ld p; col0 a multiple of 64); blocks outside the slice become zeros.  paired: the buffer
// holds (direct, mirrored) tile pairs (materialised draws) instead of upper tiles only (accumulators).
int fb_tiles_to_dense(fable_ctx* ctx, const double* d_tiles, int64_t p, int64_t t_lo, int64_t t_hi, int paired, int64_t col0,
                      int64_t ncols, double* d_out);
// one super-block (16 x 16 tiles, index sb of the enumeration) of compact UPPER tiles (buffer starts at compact index
// cbase) -> its dense block(s): d_direct = rows [1024 sI, ..) x columns [1024 sJ, ..) (nr x nc, ld nr) and, for an
// off-diagonal super-block, d_mirror = the transposed block (nc x nr, ld nc); a diagonal super-block gives the full
// symmetric block in d_direct alone.
// (ld_direct / ld_mirror > 0: leading dimensions of the two targets, e.g. p when they point into a dense p x p matrix)
int fb_superblock_to_dense(fable_ctx* ctx, const double* d_tiles, int64_t cbase, int64_t p, int64_t sb, double* d_direct,
                           double* d_mirror, int64_t ld_direct = 0, int64_t ld_mirror = 0);
// a batch of B materialised draws (slot stride p2) -> rows m_off .. m_off + B - 1 of an (ldm, p, p) array, first index fastest
int fb_slots_to_marray(fable_ctx* ctx, const double* d_slots, int B, int64_t p2, int64_t ldm, int64_t m_off, double* d_out);
// dst (k rows of pp, zero-padded) <- src (k rows of lds, p valid)
int fb_pad_rows(fable_ctx* ctx, const double* src, int64_t lds, int64_t p, int64_t pp, int k, int KP, double* dst);
// post.cu: Lsel [S][k][MC], Ssel [S][MC], idx [S] (0-based selected variables); outputs S x S column-major
int fb_post_processing(fable_ctx* ctx, const double* dLsel, const double* dSsel, const int64_t* d_idx, int64_t MC, int k,
                       int64_t S, double alpha, double* d_mean, double* d_lower, double* d_upper);
// post_tiles.cu: tile-resident mean + exact type-5 quantiles from draws in the WORKING layout (d_Lw [MC][KP][pp], d_s2w [MC][pp],
// rows >= S zero, pp a multiple of 64); d_idx [S] = the variable behind every position (diagonal term where two are equal)
bool fb_post_tiles_supported(int64_t MC, int k);
int fb_post_tiles(fable_ctx* ctx, const double* d_Lw, const double* d_s2w, const int64_t* d_idx, int64_t MC, int KP, int64_t S,
                  int64_t pp, double alpha, double* d_mean, double* d_lower, double* d_upper, int* unresolved);
int fb_ref_to_work(fable_ctx* ctx, const double* dLsel, const double* dSsel, int64_t S, int k, int64_t MC, int64_t pp, int KP,
                   double* d_Lw, double* d_s2w);
int fb_gather_work_rows(fable_ctx* ctx, const double* d_Lw, const double* d_s2w, int64_t pp, const int64_t* d_idx, int64_t S,
                        int64_t pps, int KP, int64_t MC, double* d_Lsel, double* d_s2sel);
// sampler.cu
struct SampleArgs {
  const double* G0;    // p x k shrunk loadings mean, leading dimension ldg
  int64_t ldg;
  const double* gnd;   // p
  double a_n, shape;   // shape = gamma_n / 2
  int64_t p;
  int k;
  uint64_t seed;
  const double* gam;   // injected (device) or NULL; element [(m - m_inj0)*p + j]
  const double* z;     // injected (device) or NULL; element [(m - m_inj0)*p*k + l*p + j]
  int64_t m_inj0;
};
int fb_sample_work(fable_ctx* ctx, const SampleArgs& a, int64_t m0, int B, int64_t pp, int KP,
                   double* Lw, double* s2w);
int fb_sample_ref(fable_ctx* ctx, const SampleArgs& a, int64_t m0, int64_t MC, int64_t j0,
                  int64_t nj, int64_t ldm, double* dLambda, double* dSigma);
