// C-ABI of libfable_b200.so (include/fable_b200.h): the host layer that sits where the reference's
// Rcpp exports sit (src/RcppExports.cpp:96-172) and drives the sm_100a kernels.
// Control flow that the reference runs on the host (bisection, final window, argmin:
// src/updated-FABLE-functions.cpp:215-341) is replayed here verbatim on criterion values computed
// on the device.  No CPU arithmetic fallback exists for any matrix-sized quantity.
#include <cmath>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <map>
#include <new>
#include <thread>

#include <sys/mman.h>
#include <sys/resource.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "common.cuh"
#include "tiles.cuh"

// ---- device block cache (see common.cuh) ----------------------------------------------------------
#include <condition_variable>
#include <functional>
#include <mutex>
#include <unordered_map>
namespace {
struct MemCache {
  std::mutex mu;
  struct Block { int device; size_t bytes; };
  std::unordered_map<void*, Block> live;                         // handed out
  std::multimap<std::pair<int, size_t>, void*> idle;             // (device, rounded bytes) -> block
  // A released block may still be in use by work queued on the library's streams.  Instead of draining the device at
  // every release (a full barrier, also across the copy stream the release was meant to overlap), the release records
  // one event per registered stream of the device and the NEXT owner waits for those -- by then they have normally
  // completed long ago.
  std::unordered_map<void*, std::vector<cudaEvent_t>> pending;   // idle block -> events to wait for before reuse
  std::multimap<int, cudaStream_t> streams;                      // device -> streams of the live contexts
  std::vector<cudaEvent_t> ev_pool;
  size_t idle_bytes = 0;
  cudaEvent_t get_event() {
    if (!ev_pool.empty()) { cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    return e;
  }
  void drop_pending(void* ptr, bool wait) {
    auto it = pending.find(ptr);
    if (it == pending.end()) return;
    for (cudaEvent_t e : it->second) {
      if (wait) cudaEventSynchronize(e);
      ev_pool.push_back(e);
    }
    pending.erase(it);
  }
  size_t cap() {
    static size_t c = [] {
      const char* e = getenv("FABLE_B200_CACHE_GB");
      const double gb = e ? atof(e) : 40.0;
      return static_cast<size_t>((gb < 0 ? 0 : gb) * 1073741824.0);
    }();
    return c;
  }
  // caller holds mu; frees idle blocks of `device` (all devices if < 0); returns to the caller's device
  void trim_locked(int device) {
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto it = idle.begin(); it != idle.end();) {
      if (device < 0 || it->first.first == device) {
        cudaSetDevice(it->first.first);
        drop_pending(it->second, false);       // cudaFree waits for the device itself
        cudaFree(it->second);
        idle_bytes -= it->first.second;
        it = idle.erase(it);
      } else {
        ++it;
      }
    }
    cudaSetDevice(cur);
  }
};
MemCache& mem_cache() { static MemCache* m = new MemCache; return *m; }   // leaked on purpose: outlives static destructors
size_t mem_round(size_t b) { return b <= (1u << 20) ? (b + 511) / 512 * 512 : (b + (2u << 20) - 1) / (2u << 20) * (2u << 20); }
}  // namespace

cudaError_t fb_mem_acquire(void** ptr, size_t bytes) {
  MemCache& m = mem_cache();
  const size_t rb = mem_round(bytes);
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(m.mu);
  auto it = m.idle.find({dev, rb});
  if (it != m.idle.end()) {
    *ptr = it->second;
    m.idle.erase(it);
    m.idle_bytes -= rb;
    m.drop_pending(*ptr, true);                // whatever the previous owner had queued on the block has finished
  } else {
    e = cudaMalloc(ptr, rb);
    if (e == cudaErrorMemoryAllocation) {      // the cache may be what is holding the memory: flush and retry once
      (void)cudaGetLastError();
      m.trim_locked(dev);
      e = cudaMalloc(ptr, rb);
    }
    if (e != cudaSuccess) { *ptr = nullptr; return e; }
  }
  m.live[*ptr] = {dev, rb};
  return cudaSuccess;
}

void fb_mem_release(void* ptr) {
  if (!ptr) return;
  MemCache& m = mem_cache();
  std::lock_guard<std::mutex> lk(m.mu);
  auto it = m.live.find(ptr);
  if (it == m.live.end()) { cudaFree(ptr); return; }
  const MemCache::Block b = it->second;
  m.live.erase(it);
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur != b.device) cudaSetDevice(b.device);
  if (b.bytes > m.cap()) {
    cudaFree(ptr);                             // (waits for the device)
  } else {
    // the next owner may use the block on another stream: it waits for these events, not the releaser for the device
    std::vector<cudaEvent_t> evs;
    bool ok = true;
    auto range = m.streams.equal_range(b.device);
    for (auto st = range.first; st != range.second && ok; ++st) {
      cudaEvent_t e = m.get_event();
      if (!e || cudaEventRecord(e, st->second) != cudaSuccess) { (void)cudaGetLastError(); if (e) m.ev_pool.push_back(e); ok = false; break; }
      evs.push_back(e);
    }
    if (!ok) { for (cudaEvent_t e : evs) m.ev_pool.push_back(e); evs.clear(); cudaDeviceSynchronize(); }
    if (!evs.empty()) m.pending[ptr] = std::move(evs);
    while (m.idle_bytes + b.bytes > m.cap() && !m.idle.empty()) {      // make room: drop the largest idle blocks first
      auto big = std::prev(m.idle.end());
      cudaSetDevice(big->first.first);
      m.drop_pending(big->second, false);
      cudaFree(big->second);
      m.idle_bytes -= big->first.second;
      m.idle.erase(big);
    }
    cudaSetDevice(b.device);
    m.idle.insert({{b.device, b.bytes}, ptr});
    m.idle_bytes += b.bytes;
  }
  if (cur != b.device) cudaSetDevice(cur);
}

void fb_mem_trim(int device) {
  MemCache& m = mem_cache();
  std::lock_guard<std::mutex> lk(m.mu);
  m.trim_locked(device);
}

void fb_mem_register_stream(int device, cudaStream_t s, bool add) {
  MemCache& m = mem_cache();
  std::lock_guard<std::mutex> lk(m.mu);
  if (add) { m.streams.insert({device, s}); return; }
  auto range = m.streams.equal_range(device);
  for (auto it = range.first; it != range.second; ++it)
    if (it->second == s) { m.streams.erase(it); break; }
}

namespace {

constexpr int KSMALL = 32;   // ranks up to this use the streaming row kernel; above: DMMA residual GEMM

thread_local std::string g_create_error;

struct Inputs {
  DevBuf Y, U, V, d;
  int64_t n = 0, p = 0, r = 0;
  int kcols = 0;   // number of leading singular triplets resident
  bool own_svd = false;   // (U, V, d) were produced by fb_svd from this Y: V diag(d) = Y'U and U'U = I by construction
};

// Column shard (fable_ctx_set_column_shard): Y, V and every per-variable vector hold the columns this context
// owns; p_total replaces p wherever the reference's formula means "all variables", and sums over variables are
// completed by the caller's all-reduce.
inline bool shard_on(const fable_ctx* ctx) { return ctx->shard_p > 0; }
inline int64_t glob_p(const fable_ctx* ctx, int64_t p_local) { return shard_on(ctx) ? ctx->shard_p : p_local; }
int shard_sum(fable_ctx* ctx, double* vals, int64_t count) {
  if (!shard_on(ctx)) return FABLE_OK;
  if (ctx->shard_fn(ctx->shard_user, vals, count) != 0) return ctx->fail(FABLE_ERR_INVALID, "column shard: the all-reduce callback failed");
  return FABLE_OK;
}
// sum of `count` DEVICE doubles over the ranks, in place: through the device-pointer callback when the caller gave one
// (NCCL on the context's stream: the n x n Gram never leaves HBM), else bounced through the host callback
int shard_sum_dev(fable_ctx* ctx, double* dptr, int64_t count) {
  if (!shard_on(ctx)) return FABLE_OK;
  if (ctx->shard_dev_fn) {
    if (ctx->shard_dev_fn(ctx->shard_dev_user, dptr, count, static_cast<void*>(ctx->stream)) != 0)
      return ctx->fail(FABLE_ERR_INVALID, "column shard: the device all-reduce callback failed");
    return FABLE_OK;
  }
  std::vector<double> h(static_cast<size_t>(count));
  FB_CUDA(ctx, cudaMemcpyAsync(h.data(), dptr, h.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  FB_TRY(shard_sum(ctx, h.data(), count));
  FB_CUDA(ctx, cudaMemcpyAsync(dptr, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // h goes out of scope
  return FABLE_OK;
}
#define FB_NO_SHARD(ctx, what)                                                                                          \
  FB_REQUIRE(ctx, !shard_on(ctx), what " needs all variables on one device: not available under a column shard "        \
                                       "(gather the row posterior and use fable_posterior_create_from_rows)")

int check_dims(fable_ctx* ctx, const void* Y, int64_t n, int64_t p, const void* U, const void* V, const void* d,
               int64_t r) {
  FB_REQUIRE(ctx, Y && U && V && d, "NULL input matrix");
  FB_REQUIRE(ctx, n >= 1 && p >= 1, "Y must have at least one row and one column");
  FB_REQUIRE(ctx, !shard_on(ctx) || ctx->shard_off + p <= ctx->shard_p, "column shard: offset + local columns exceed p_total");
  const int64_t pg = glob_p(ctx, p);
  FB_REQUIRE(ctx, r >= 1 && r <= (n < pg ? n : pg), "length(svalsY) must be in 1..min(n,p)");
  return FABLE_OK;
}

int upload_bytes(fable_ctx* ctx, void* dptr, const void* h, size_t bytes);
int copy_threads();
int download_2d(fable_ctx* ctx, double* h, int64_t hld, const double* dptr, int64_t rows, int64_t ncols);

int upload(fable_ctx* ctx, DevBuf& b, const double* h, size_t count) {
  FB_CUDA(ctx, b.alloc(count * sizeof(double)));
  return upload_bytes(ctx, b.ptr, h, count * sizeof(double));
}

// U is n x r and V is p x r column-major, so their first kcols columns are one contiguous block.
int upload_inputs(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, const double* U, const double* V,
                  const double* d, int64_t r, int kcols, Inputs& in) {
  FB_TRY(check_dims(ctx, Y, n, p, U, V, d, r));
  FB_REQUIRE(ctx, kcols >= 1 && kcols <= r, "rank must be in 1..length(svalsY)");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  in.n = n; in.p = p; in.r = r; in.kcols = kcols;
  FB_TRY(upload(ctx, in.Y, Y, static_cast<size_t>(n) * p));
  FB_TRY(upload(ctx, in.U, U, static_cast<size_t>(n) * kcols));
  FB_TRY(upload(ctx, in.V, V, static_cast<size_t>(p) * kcols));
  FB_TRY(upload(ctx, in.d, d, static_cast<size_t>(kcols)));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FABLE_OK;
}

// scratch reused across criterion calls
struct RowScratch {
  DevBuf c, q, s, resid;
  int kcap = 0;
  int ensure(fable_ctx* ctx, int64_t p, int k) {
    if (k > kcap) { FB_CUDA(ctx, c.alloc(static_cast<size_t>(p) * k * sizeof(double))); kcap = k; }
    if (!q.ptr) {
      FB_CUDA(ctx, q.alloc(p * sizeof(double)));
      FB_CUDA(ctx, s.alloc(p * sizeof(double)));
      FB_CUDA(ctx, resid.alloc(p * sizeof(double)));
    }
    return FABLE_OK;
  }
};

// c = V_k .* d_k, q, raw residual sums (and s when want_s) for rank k.
int row_pass(fable_ctx* ctx, const Inputs& in, int k, RowScratch& sc, bool want_s) {
  FB_TRY(sc.ensure(ctx, in.p, k));
  FB_TRY(fb_cq(ctx, in.V.d(), in.p, in.d.d(), in.p, k, sc.c.d(), in.p, sc.q.d()));
  static const int ksmall = getenv("FABLE_B200_KSMALL") ? atoi(getenv("FABLE_B200_KSMALL")) : KSMALL;
  if (k <= ksmall) {
    FB_TRY(fb_rowstats(ctx, in.Y.d(), in.n, in.p, in.n, in.U.d(), in.n, sc.c.d(), in.p, k,
                       want_s ? sc.s.d() : nullptr, sc.resid.d()));
  } else {
    FB_TRY(fb_gemm_resid(ctx, in.n, in.p, k, in.Y.d(), in.n, in.U.d(), in.n, sc.c.d(), in.p, sc.resid.d()));
    if (want_s) FB_TRY(fb_colsumsq(ctx, in.Y.d(), in.n, in.p, in.n, sc.s.d()));
  }
  return FABLE_OK;
}

// The criterion of the pure-R prototype RankEstimator (R/FABLE_code.R:57-116) from the residual column sums of squares r_j
// and their logs: sigma_j^2 = r_j / (n - k) (:87), LogLik = (-n sum_j log sd_j - 0.5 sum_ij (resid_ij / sd_j)^2) / (n p)
// (LogLikelihoodEvaluator :12-27), criterion = -2 LogLik + k max(n,p) log(min(n,p)) / (n p) (:113-114).
// sumlog = sum_j log(r_j / (n - k)), ros = sum_j r_j / sigma_j^2.
double proto_criterion(double dn, double dp, int k, double sumlog, double ros) {
  const double term1 = -dn * (0.5 * sumlog), term2 = -0.5 * ros;
  const double ll = (term1 + term2) / (dn * dp);
  const double maxint = dn > dp ? dn : dp, minint = dn < dp ? dn : dp;
  return (-2.0 * ll) + (static_cast<double>(k) * maxint * std::log(minint) / (dn * dp));
}

// CPPCriterionFunction (cpp:127-213) on resident inputs; variant 1: the R prototype's criterion (above).
int criterion_dev(fable_ctx* ctx, const Inputs& in, int k, RowScratch& sc, double* out, int variant = 0) {
  FB_REQUIRE(ctx, k >= 1 && k <= in.kcols, "criterion: kInd out of range");   // arma: cols(0,k-1) throws
  FB_TRY(row_pass(ctx, in, k, sc, false));
  double sumlog = 0.0;
  const double dn = static_cast<double>(in.n), dp = static_cast<double>(glob_p(ctx, in.p));
  if (variant == 1) {
    FB_REQUIRE(ctx, in.n > k, "prototype criterion: needs k < n (sigma^2 = resid / (n - k))");
    const double inv = 1.0 / (dn - static_cast<double>(k));
    double both[2] = {0.0, 0.0};
    FB_TRY(fb_sum_log_scaled(ctx, sc.resid.d(), in.p, inv, &both[0]));
    FB_TRY(fb_resid_over_scaled_self_sum(ctx, sc.resid.d(), in.p, inv, &both[1]));
    FB_TRY(shard_sum(ctx, both, 2));
    if (!std::isfinite(both[0]) || !std::isfinite(both[1])) return ctx->fail(FABLE_ERR_NUMERIC, "criterion: non-positive residual variance");
    *out = proto_criterion(dn, dp, k, both[0], both[1]);
    return FABLE_OK;
  }
  FB_TRY(fb_sum_log_scaled(ctx, sc.resid.d(), in.p, 1.0 / dn, &sumlog));
  FB_TRY(shard_sum(ctx, &sumlog, 1));                                            // sum over ALL variables (cpp:199)
  if (!std::isfinite(sumlog)) return ctx->fail(FABLE_ERR_NUMERIC, "criterion: non-positive residual variance");
  const double LL = (-dn * dp / 2.0) - (dn * sumlog / 2.0);                     // cpp:199
  const double maxint = dn > dp ? dn : dp, minint = dn < dp ? dn : dp;          // cpp:206-207
  *out = (-2.0 * LL) + (static_cast<double>(k) * maxint * std::log(minint));    // cpp:209
  return FABLE_OK;
}

// Criterion values for the rank search.  The reference evaluates CPPCriterionFunction from scratch at every visited rank
// (cpp:233-247, 321-330: ~25 calls, each an n x p x k' product and an n x p residual pass).  Here the search is STEERED by
// the all-rank form of SURVEY A.3 -- when (U, V, d) is a consistent SVD of Y, every sum_j log sigma_j^2(k') comes out of
// ONE pass over V -- and DECIDED by the reference's own arithmetic wherever it matters: two criteria closer than 1e-9
// relative are re-evaluated with the explicit residual before they are compared, so the bisection takes the same
// branches and the window arg-min picks the same rank.  Consistency is checked, not assumed: inputs at the
// CPP* boundary are the caller's (Y'U against V diag(d) and U'U against I, one n x p x kMax product); if the check
// fails, or any shortcut value is not finite, every evaluation is the explicit one.
struct CritCache {
  fable_ctx* ctx; const Inputs& in; RowScratch sc; std::map<int, double> memo;
  std::vector<double> fast;      // fast[k-1]: criterion from the all-rank pass (empty: not available)
  int variant = 0;               // 0: CPPCriterionFunction (cpp:127-213); 1: the R prototype's criterion (R/FABLE_code.R:57-116)
  CritCache(fable_ctx* c, const Inputs& i, int v = 0) : ctx(c), in(i), variant(v) {}
  int eval(int k, double* v) {                    // the reference's arithmetic (explicit residual)
    auto it = memo.find(k);
    if (it != memo.end()) { *v = it->second; return FABLE_OK; }   // same inputs -> same value (cpp:233 re-evaluates)
    FB_TRY(criterion_dev(ctx, in, k, sc, v, variant));
    memo[k] = *v;
    return FABLE_OK;
  }
  int prepare(int kMax) {
    const bool on = !(getenv("FABLE_B200_RANK_FAST") && atoi(getenv("FABLE_B200_RANK_FAST")) == 0);   // 0: every evaluation explicit (A/B, tests)
    fast.clear();
    if (!on || kMax < 3 || kMax > in.kcols) return FABLE_OK;
    const int64_t n = in.n, p = in.p;
    int bad = 0;
    if (!in.own_svd) {
      DevBuf T, E;
      FB_CUDA(ctx, T.alloc(static_cast<size_t>(p) * kMax * sizeof(double)));
      FB_TRY(fb_gemm(ctx, 1, 1, p, kMax, n, 1.0, in.Y.d(), n, in.U.d(), n, T.d(), p));                 // T = Y'U_kMax
      double md = 0.0, ma = 0.0;
      FB_TRY(fb_svd_consistency(ctx, T.d(), p, in.V.d(), p, in.d.d(), p, kMax, &md, &ma));
      FB_CUDA(ctx, E.alloc(static_cast<size_t>(kMax) * kMax * sizeof(double)));
      FB_TRY(fb_gemm(ctx, 1, 1, kMax, kMax, n, 1.0, in.U.d(), n, in.U.d(), n, E.d(), kMax));           // U'U
      double me = 0.0;
      FB_TRY(fb_identity_err(ctx, E.d(), kMax, &me));
      if (!(md <= 1e-11 * ma) || !(me <= 1e-11)) bad = 1;
    }
    std::vector<double> sl(kMax, 0.0);
    if (!bad) {
      FB_TRY(sc.ensure(ctx, p, 1));
      FB_TRY(fb_colsumsq(ctx, in.Y.d(), n, p, n, sc.s.d()));
      FB_TRY(fb_prefix_sumlog(ctx, in.V.d(), p, in.d.d(), sc.s.d(), p, kMax, n, sl.data()));
    }
    if (shard_on(ctx)) {            // sums over ALL variables; a failed check on any rank switches every rank to the explicit path
      std::vector<double> buf(sl);
      buf.push_back(static_cast<double>(bad));
      FB_TRY(shard_sum(ctx, buf.data(), static_cast<int64_t>(buf.size())));
      bad = buf.back() != 0.0;
      std::copy(buf.begin(), buf.end() - 1, sl.begin());
    }
    if (bad) return FABLE_OK;
    const double dn = static_cast<double>(n), dp = static_cast<double>(glob_p(ctx, p));
    const double maxint = dn > dp ? dn : dp, minint = dn < dp ? dn : dp;
    for (int k = 1; k <= kMax; ++k) {
      if (!std::isfinite(sl[k - 1]) || (variant == 1 && k >= n)) { fast.clear(); return FABLE_OK; }
      if (variant == 1) {      // sum_j log(r_j / (n-k)) = sum_j log(r_j / n) + p log(n / (n-k));  sum_j r_j / sigma_j^2 = p (n-k)
        const double dk = static_cast<double>(k);
        fast.push_back(proto_criterion(dn, dp, k, sl[k - 1] + dp * std::log(dn / (dn - dk)), dp * (dn - dk)));
        continue;
      }
      const double LL = (-dn * dp / 2.0) - (dn * sl[k - 1] / 2.0);                                      // cpp:199
      fast.push_back((-2.0 * LL) + (static_cast<double>(k) * maxint * std::log(minint)));               // cpp:209
    }
    return FABLE_OK;
  }
  // sign of crit(a) - crit(b) as the reference's arithmetic gives it: -1, 0, +1
  int compare(int a, int b, int* sign) {
    if (a == b) { *sign = 0; return FABLE_OK; }
    double va, vb;
    const bool have = !fast.empty() && a >= 1 && b >= 1 && a <= static_cast<int>(fast.size()) && b <= static_cast<int>(fast.size());
    if (have && !memo.count(a) && !memo.count(b)) {
      va = fast[a - 1]; vb = fast[b - 1];
      const double mag = std::fabs(va) > std::fabs(vb) ? std::fabs(va) : std::fabs(vb);
      if (std::fabs(va - vb) > 1e-9 * mag) { *sign = va < vb ? -1 : 1; return FABLE_OK; }
    }
    if (have) {
      // one of them already explicit, or too close to call: explicit values for both unless the shortcut is decisive
      va = memo.count(a) ? memo[a] : fast[a - 1];
      vb = memo.count(b) ? memo[b] : fast[b - 1];
      const double mag = std::fabs(va) > std::fabs(vb) ? std::fabs(va) : std::fabs(vb);
      if (std::fabs(va - vb) > 1e-9 * mag) { *sign = va < vb ? -1 : 1; return FABLE_OK; }
    }
    FB_TRY(eval(a, &va));
    FB_TRY(eval(b, &vb));
    *sign = va < vb ? -1 : (va > vb ? 1 : 0);
    return FABLE_OK;
  }
};

// CPPBisectionRecursion (cpp:215-271), iterative form of the tail recursion.
int bisection(CritCache& cc, int lower, int upper, int* lo, int* hi) {
  while (upper - lower > 5) {                                    // cpp:223
    const int mid = (lower + upper) / 2;                         // cpp:225
    int sg;
    FB_TRY(cc.compare(lower, mid, &sg));                         // cpp:233-235
    if (sg < 0) { upper = mid; continue; }                       // cpp:237-239: crit(lower) < crit(mid)
    const int tq = (mid + upper) / 2;                            // cpp:243
    FB_TRY(cc.compare(mid, tq, &sg));                            // cpp:247
    if (sg <= 0) upper = mid; else lower = mid;                  // cpp:249-255: crit(mid) <= crit(tq)
  }
  *lo = lower; *hi = upper;
  return FABLE_OK;
}

// CPPRankEstimator (cpp:281-341); variant 1: the R prototype RankEstimator (R/FABLE_code.R:32-186: the same bisection and
// window, :118-184, over its own criterion)
int rank_dev(fable_ctx* ctx, const Inputs& in, int kMax, int* kEst, int variant = 0) {
  FB_REQUIRE(ctx, kMax >= 1 && kMax <= in.kcols, "kMax must be in 1..length(svalsY)");
  CritCache cc(ctx, in, variant);
  FB_TRY(cc.prepare(kMax));
  int lo, hi;
  FB_TRY(bisection(cc, 1, kMax, &lo, &hi));                      // cpp:314
  int best = lo;
  for (int k = lo + 1; k <= hi; ++k) {                           // cpp:321-330
    int sg;
    FB_TRY(cc.compare(k, best, &sg));
    if (sg < 0) best = k;                                        // index_min = first minimum (cpp:337)
  }
  *kEst = best;
  return FABLE_OK;
}

struct RowPosterior {
  DevBuf G0, gnd, sig, sigmahat;
  double tausq = 0, w = 0, gamma_n = 0;
};

// A.2 of SURVEY.md: cpp:371-391 (shrunk=0) or cpp:470-499 / 564-592 (shrunk=1)
int row_posterior(fable_ctx* ctx, const Inputs& in, int k, double gamma0, double delta0sq, int shrunk,
                  RowPosterior& rp) {
  RowScratch sc;
  FB_TRY(row_pass(ctx, in, k, sc, true));
  double qsum = 0.0;
  FB_TRY(fb_q_over_sig_sum(ctx, sc.q.d(), sc.resid.d(), in.n, in.p, &qsum));
  FB_TRY(shard_sum(ctx, &qsum, 1));
  // mean_j(q_j / sig_j) / (n k) over all variables (cpp:383, :483, :577)
  rp.tausq = (qsum / static_cast<double>(glob_p(ctx, in.p))) / (static_cast<double>(in.n) * static_cast<double>(k));
  if (!(rp.tausq > 0.0) || !std::isfinite(rp.tausq))
    return ctx->fail(FABLE_ERR_NUMERIC, "row posterior: tauSq estimate is not positive and finite");
  rp.w = static_cast<double>(in.n) + 1.0 / rp.tausq;
  rp.gamma_n = gamma0 + static_cast<double>(in.n);
  FB_CUDA(ctx, rp.G0.alloc(static_cast<size_t>(in.p) * k * sizeof(double)));
  FB_CUDA(ctx, rp.gnd.alloc(in.p * sizeof(double)));
  FB_CUDA(ctx, rp.sig.alloc(in.p * sizeof(double)));
  FB_CUDA(ctx, rp.sigmahat.alloc(in.p * sizeof(double)));
  FB_TRY(fb_finish_posterior(ctx, sc.c.d(), in.p, sc.s.d(), sc.q.d(), sc.resid.d(), in.n, in.p, k, gamma0,
                             delta0sq, rp.tausq, shrunk, rp.G0.d(), in.p, rp.gnd.d(), rp.sig.d(),
                             rp.sigmahat.d()));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // scratch is released on return
  return FABLE_OK;
}

// Device -> host.  Results land in caller memory, which under R is PAGEABLE: a plain cudaMemcpy then runs
// at ~6 GB/s (the driver stages through one pinned buffer on one thread).  Large copies into pageable
// memory are therefore pipelined by hand: 64 MB chunks go D2H into two pinned staging buffers on the
// context's stream while host threads copy the previous chunk out in parallel.  Pinned destinations
// (bench.py's e2e buffers) and small copies take the plain asynchronous path.
constexpr size_t kStageBytes = size_t(64) << 20;
constexpr size_t kFastCopyMin = size_t(16) << 20;

// Host copy workers: a small persistent pool (spawning threads per copy costs ~50 us each, which matters when a p x p result
// leaves as hundreds of 8 MB blocks).  run(n, fn) executes fn(0) .. fn(n-1) on the workers and the calling thread.
class CopyPool {
 public:
  static CopyPool& get() { static CopyPool* p = new CopyPool; return *p; }     // leaked on purpose (no static-destruction order)
  void run(int njobs, const std::function<void(int)>& fn) {
    if (njobs <= 1) { if (njobs == 1) fn(0); return; }
    std::unique_lock<std::mutex> call(call_mu_);                 // one parallel section at a time
    ensure(njobs - 1);
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn; njobs_ = njobs; next_ = 0; pending_ = njobs; ++gen_;
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return pending_ == 0; });
    fn_ = nullptr;
  }
 private:
  void ensure(int n) {
    while (static_cast<int>(workers_.size()) < n && workers_.size() < 64)
      workers_.emplace_back([this] {
        uint64_t seen = 0;
        for (;;) {
          {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return gen_ != seen; });
            seen = gen_;
          }
          work();
        }
      }), workers_.back().detach();
  }
  void work() {
    for (;;) {
      int j;
      const std::function<void(int)>* f;
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (!fn_ || next_ >= njobs_) return;
        j = next_++; f = fn_;
      }
      (*f)(j);
      std::lock_guard<std::mutex> lk(mu_);
      if (--pending_ == 0) done_cv_.notify_all();
    }
  }
  std::mutex mu_, call_mu_;
  std::condition_variable cv_, done_cv_;
  std::vector<std::thread> workers_;
  const std::function<void(int)>* fn_ = nullptr;
  int njobs_ = 0, next_ = 0, pending_ = 0;
  uint64_t gen_ = 0;
};

void parallel_memcpy(char* dst, const char* src, size_t bytes, int nthreads) {
  if (nthreads <= 1 || bytes < (size_t(4) << 20)) { std::memcpy(dst, src, bytes); return; }
  const size_t per = ((bytes + nthreads - 1) / nthreads + 4095) & ~size_t(4095);
  const int njobs = static_cast<int>((bytes + per - 1) / per);
  CopyPool::get().run(njobs, [=](int t) {
    const size_t off = per * t;
    const size_t len = bytes - off < per ? bytes - off : per;
    std::memcpy(dst + off, src + off, len);
  });
}

// A freshly allocated result vector is untouched virtual memory: the copy-out pays one page fault (and one kernel page
// clear) per 4 KB.  R's allocVector gets such blocks from malloc/mmap without any huge-page advice (numpy asks for it
// itself), so on systems whose transparent-huge-page mode is "madvise" the library asks for the 2 MB-aligned interior
// of a large pageable destination: a hint without semantic effect, ignored where THP is off.  Measured (tools/
// bench_d2h.py, 3.2 GB into a fresh anonymous mmap): 146-173 ms -> 131-133 ms.  FABLE_B200_THP=0: no hint.
void hint_huge_pages(char* dst, size_t bytes) {
  static const bool on = !(getenv("FABLE_B200_THP") && atoi(getenv("FABLE_B200_THP")) == 0);
  if (!on) return;
  const uintptr_t two_mb = uintptr_t(2) << 20;
  const uintptr_t lo = (reinterpret_cast<uintptr_t>(dst) + two_mb - 1) & ~(two_mb - 1);
  const uintptr_t hi = (reinterpret_cast<uintptr_t>(dst) + bytes) & ~(two_mb - 1);
  if (hi > lo) (void)madvise(reinterpret_cast<void*>(lo), hi - lo, MADV_HUGEPAGE);
}

bool is_pageable(const void* h) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, h) != cudaSuccess) { (void)cudaGetLastError(); return true; }
  return at.type == cudaMemoryTypeUnregistered;
}

// First touch of a large pageable result while the device still computes.  A freshly allocated p x p result is untouched
// virtual memory and the copy-out pays one page fault (and one page clear) per 4 KB: 109 of the 146 ms of FABLEPosteriorMean at
// C3 were the 3.2 GB CovPostMean crossing into such memory at 29 GB/s, PCIe alone would need 60.  The pages are therefore
// touched -- one zero byte per 4 KB, by plain page faults from a few host threads, each on its own slice -- from the moment
// the entry point knows the destination, i.e. during the upload, the SVD and the rank search, and -- being atomic adds of
// zero -- on through the result copy itself.  (The earlier experiment used madvise(MADV_POPULATE_WRITE), which serialises on the
// address-space lock against the copy threads and doubled the wall time; plain faults do not.)  FABLE_B200_PREFAULT=0: off.
class HostPrefault {
 public:
  HostPrefault() = default;
  HostPrefault(const HostPrefault&) = delete;
  HostPrefault& operator=(const HostPrefault&) = delete;
  ~HostPrefault() { join(); }
  // (a copy into a range that an entry point is already touching does not start a second set of threads)
  void start(fable_ctx* ctx, void* dst, size_t bytes, size_t min_bytes = size_t(256) << 20) {
    const char* e = getenv("FABLE_B200_PREFAULT");
    if (!dst || bytes < min_bytes || (e && atoi(e) == 0) || !is_pageable(dst)) return;
    {
      const char* lo_ = static_cast<const char*>(dst);
      for (int i = 0; i < 4; ++i)
        if (ctx->prefault_lo[i] && lo_ < ctx->prefault_hi[i] && ctx->prefault_lo[i] < lo_ + bytes) return;
      for (int i = 0; i < 4 && !owner_; ++i)
        if (!ctx->prefault_lo[i]) { ctx->prefault_lo[i] = lo_; ctx->prefault_hi[i] = lo_ + bytes; owner_ = ctx; slot_ = i; }
    }
    int nt = copy_threads();
    if (nt <= 0) return;
    if (nt > 16) nt = 16;
    hint_huge_pages(static_cast<char*>(dst), bytes);
    char* base = static_cast<char*>(dst);
    const size_t per = ((bytes + nt - 1) / nt + 4095) & ~size_t(4095);
    for (int t = 0; t < nt; ++t) {
      const size_t lo = per * t, hi = lo + per < bytes ? lo + per : bytes;
      if (lo >= hi) break;
      try {
      threads_.emplace_back([base, lo, hi] {
        // lowest priority: the calling thread drives the eigensolver's sweeps (a launch and a 4-byte read-back per sweep)
        // and must not queue behind 16 page-faulting threads (measured: SVD stage 19.7 -> 30-35 ms without this)
        (void)setpriority(PRIO_PROCESS, static_cast<id_t>(syscall(SYS_gettid)), 19);
#if defined(__x86_64__)
        // an atomic add of zero faults the page in and leaves its content alone even if the result copy is already
        // writing it: the copy does not have to wait for these threads (they are joined when the entry point returns)
        for (size_t off = lo; off < hi; off += 4096) asm volatile("lock addb $0, (%0)" ::"r"(base + off) : "memory", "cc");
        asm volatile("lock addb $0, (%0)" ::"r"(base + hi - 1) : "memory", "cc");
#else
        volatile char* q = base;
        for (size_t off = lo; off < hi; off += 4096) q[off] = 0;
        q[hi - 1] = 0;
#endif
      });
      } catch (...) { break; }           // no thread to spare: the copy faults its pages itself, as before
    }
  }
  void join() {
    for (std::thread& t : threads_) t.join();
    threads_.clear();
    if (owner_) { owner_->prefault_lo[slot_] = nullptr; owner_->prefault_hi[slot_] = nullptr; owner_ = nullptr; }
  }
  // before the first byte of the result is written: nothing to wait for where the touch is an atomic read-modify-write
  void before_result() {
#if !defined(__x86_64__)
    join();
#endif
  }
 private:
  std::vector<std::thread> threads_;
  fable_ctx* owner_ = nullptr;
  int slot_ = 0;
};


int download(fable_ctx* ctx, double* h, const double* dptr, size_t count) {
  if (!h) return FABLE_OK;
  const size_t bytes = count * sizeof(double);
  if (bytes < kFastCopyMin || !is_pageable(h)) {
    FB_CUDA(ctx, cudaMemcpyAsync(h, dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return FABLE_OK;
  }
  for (int s = 0; s < 2; ++s)
    if (!ctx->stage[s]) {
      FB_CUDA(ctx, cudaHostAlloc(&ctx->stage[s], kStageBytes, cudaHostAllocDefault));
      FB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->stage_ev[s], cudaEventDisableTiming));
    }
  const int nthreads = copy_threads();
  if (nthreads <= 0) {                 // 0: leave the copy to the driver
    FB_CUDA(ctx, cudaMemcpyAsync(h, dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return FABLE_OK;
  }
  const char* src = reinterpret_cast<const char*>(dptr);
  char* dst = reinterpret_cast<char*>(h);
  hint_huge_pages(dst, bytes);
  // (touching the pages from extra threads only once the copy starts does not pay: 103 vs 112 ms for 3.2 GB; HostPrefault is
  // started by the entry points that have device work to hide it behind)
  const size_t nchunks = (bytes + kStageBytes - 1) / kStageBytes;
  // the staging buffers are shared with the uploads, which may have run on the context's other stream: whatever last
  // used them must have finished before a transfer of this stream writes into them
  for (int s = 0; s < 2; ++s) FB_CUDA(ctx, cudaEventSynchronize(ctx->stage_ev[s]));
  auto issue = [&](size_t i) -> cudaError_t {
    const size_t off = i * kStageBytes, len = bytes - off < kStageBytes ? bytes - off : kStageBytes;
    cudaError_t e = cudaMemcpyAsync(ctx->stage[i & 1], src + off, len, cudaMemcpyDeviceToHost, ctx->stream);
    if (e != cudaSuccess) return e;
    return cudaEventRecord(ctx->stage_ev[i & 1], ctx->stream);
  };
  FB_CUDA(ctx, issue(0));
  for (size_t i = 0; i < nchunks; ++i) {
    if (i + 1 < nchunks) FB_CUDA(ctx, issue(i + 1));           // next chunk crosses PCIe while this one is copied out
    FB_CUDA(ctx, cudaEventSynchronize(ctx->stage_ev[i & 1]));
    const size_t off = i * kStageBytes, len = bytes - off < kStageBytes ? bytes - off : kStageBytes;
    parallel_memcpy(dst + off, static_cast<const char*>(ctx->stage[i & 1]), len, nthreads);
  }
  return FABLE_OK;
}

// Strided device -> host: `ncols` columns of `rows` doubles, contiguous on the device, to host columns `hld` doubles
// apart (a block of a column-major matrix).  Pinned destinations take one 2-D copy; pageable ones go through the pinned
// staging and are scattered by the host threads (a 2-D copy into pageable memory is a synchronous staged copy per row).
int download_2d(fable_ctx* ctx, double* h, int64_t hld, const double* dptr, int64_t rows, int64_t ncols) {
  if (!h || rows <= 0 || ncols <= 0) return FABLE_OK;
  if (hld == rows) return download(ctx, h, dptr, static_cast<size_t>(rows) * ncols);
  const int nthreads = copy_threads();
  if (!is_pageable(h) || nthreads <= 0) {
    FB_CUDA(ctx, cudaMemcpy2DAsync(h, static_cast<size_t>(hld) * sizeof(double), dptr, static_cast<size_t>(rows) * sizeof(double),
                                   static_cast<size_t>(rows) * sizeof(double), static_cast<size_t>(ncols), cudaMemcpyDeviceToHost, ctx->stream));
    return FABLE_OK;
  }
  for (int s = 0; s < 2; ++s)
    if (!ctx->stage[s]) {
      FB_CUDA(ctx, cudaHostAlloc(&ctx->stage[s], kStageBytes, cudaHostAllocDefault));
      FB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->stage_ev[s], cudaEventDisableTiming));
    }
  const int64_t cols_per = static_cast<int64_t>(kStageBytes / (static_cast<size_t>(rows) * sizeof(double)));
  FB_REQUIRE(ctx, cols_per >= 1, "download_2d: a column exceeds the staging buffer");
  for (int s = 0; s < 2; ++s) FB_CUDA(ctx, cudaEventSynchronize(ctx->stage_ev[s]));
  const int64_t nchunks = ceil_div(ncols, cols_per);
  auto issue = [&](int64_t i) -> cudaError_t {
    const int64_t c0 = i * cols_per, nc = ncols - c0 < cols_per ? ncols - c0 : cols_per;
    cudaError_t e = cudaMemcpyAsync(ctx->stage[i & 1], dptr + c0 * rows, static_cast<size_t>(nc) * rows * sizeof(double),
                                    cudaMemcpyDeviceToHost, ctx->stream);
    if (e != cudaSuccess) return e;
    return cudaEventRecord(ctx->stage_ev[i & 1], ctx->stream);
  };
  FB_CUDA(ctx, issue(0));
  for (int64_t i = 0; i < nchunks; ++i) {
    if (i + 1 < nchunks) FB_CUDA(ctx, issue(i + 1));
    FB_CUDA(ctx, cudaEventSynchronize(ctx->stage_ev[i & 1]));
    const int64_t c0 = i * cols_per, nc = ncols - c0 < cols_per ? ncols - c0 : cols_per;
    const double* src = static_cast<const double*>(ctx->stage[i & 1]);
    int nt = static_cast<int>((static_cast<size_t>(nc) * rows * sizeof(double)) >> 20);      // >= 1 MB per worker
    nt = nt < 1 ? 1 : (nt > nthreads ? nthreads : nt);
    if (nt > nc) nt = static_cast<int>(nc);
    CopyPool::get().run(nt, [=](int t) {
      for (int64_t c = t; c < nc; c += nt) std::memcpy(h + (c0 + c) * hld, src + c * rows, static_cast<size_t>(rows) * sizeof(double));
    });
  }
  return FABLE_OK;
}

// Host -> device, the mirror image: R hands over PAGEABLE matrices (Y is 160 MB at C3, V_kMax another 150 MB), which
// a plain cudaMemcpyAsync stages through the driver's single bounce buffer.  Large pageable sources are copied by
// several host threads into the two pinned staging buffers and sent from there, chunk i + 1 being gathered while
// chunk i crosses PCIe.
int copy_threads() {
  unsigned hw = std::thread::hardware_concurrency();
  // measured on the 16-thread host of the B200 box, 3.2 GB p x p result into fresh pageable memory: 4 threads 15 GB/s,
  // 8 threads 25 GB/s, 16 threads 32 GB/s (the copy-out and its first-touch page faults are the limit, not PCIe)
  int nthreads = hw >= 16 ? 16 : (hw >= 4 ? static_cast<int>(hw / 2) : 1);
  if (const char* e = getenv("FABLE_B200_COPY_THREADS")) nthreads = atoi(e);
  return nthreads;
}

int upload_bytes(fable_ctx* ctx, void* dptr, const void* h, size_t bytes) {
  const int nthreads = copy_threads();
  if (bytes < kFastCopyMin || nthreads <= 0 || !is_pageable(h)) {
    FB_CUDA(ctx, cudaMemcpyAsync(dptr, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return FABLE_OK;
  }
  for (int s = 0; s < 2; ++s)
    if (!ctx->stage[s]) {
      FB_CUDA(ctx, cudaHostAlloc(&ctx->stage[s], kStageBytes, cudaHostAllocDefault));
      FB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->stage_ev[s], cudaEventDisableTiming));
    }
  const char* src = static_cast<const char*>(h);
  char* dst = static_cast<char*>(dptr);
  const size_t nchunks = (bytes + kStageBytes - 1) / kStageBytes;
  for (size_t i = 0; i < nchunks; ++i) {
    const size_t off = i * kStageBytes, len = bytes - off < kStageBytes ? bytes - off : kStageBytes;
    FB_CUDA(ctx, cudaEventSynchronize(ctx->stage_ev[i & 1]));        // the transfer that last used this buffer is done
    parallel_memcpy(static_cast<char*>(ctx->stage[i & 1]), src + off, len, nthreads);
    FB_CUDA(ctx, cudaMemcpyAsync(dst + off, ctx->stage[i & 1], len, cudaMemcpyHostToDevice, ctx->stream));
    FB_CUDA(ctx, cudaEventRecord(ctx->stage_ev[i & 1], ctx->stream));
  }
  return FABLE_OK;
}

// dense p x p  A A' + diag(s) into a host buffer, via a device p x p
int rankk_cov_to_host(fable_ctx* ctx, const double* dA, int64_t lda, const double* ds, int64_t p, int k,
                      double* hOut) {
  DevBuf out;
  FB_CUDA(ctx, out.alloc(static_cast<size_t>(p) * p * sizeof(double)));
  FB_TRY(fable_dev_rankk_cov(ctx, dA, lda, ds, p, k, out.d()));
  FB_TRY(download(ctx, hOut, out.d(), static_cast<size_t>(p) * p));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FABLE_OK;
}

}  // namespace

struct fable_posterior {
  fable_ctx* ctx = nullptr;
  int64_t n = 0, p = 0, pp = 0;
  int k = 0, KP = 0;
  double gamma0 = 0, delta0sq = 0, varInflation = 0, a_n = 0;
  RowPosterior rp;
  DevBuf Lw, s2w;
  int Bcap = 0;
  DevBuf acc_sum, acc_sq;       // compact upper tiles of the slice [tile_begin, tile_end) (tiles.cuh): acc_tiles x 4096 doubles
  int64_t acc_tiles = 0;
  // Partial sums of the materialising batches (see ensure_partials): slot s holds (sum, sumsq) of one batch, folded into
  // the accumulators every part_slots batches and before anyone looks at them.
  DevBuf part;
  int part_slots = -1;          // -1: not decided yet; 0: off (classic read-modify-write inside the covariance kernel)
  int part_pending = 0;
  DevBuf hyp_G0;                // unshrunk G0 of FABLEHyperParameters (cpp:390), kept by fable_wrapper_sampler_begin
  DevBuf sb_stage;              // dense blocks of super-blocks on their way to the host (accum_fetch_superblocks)
  cudaEvent_t sb_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  ~fable_posterior() { for (cudaEvent_t e : sb_ev) if (e) cudaEventDestroy(e); }
  DevBuf cc_gdiag;              // entry-wise coverage correction (f3): diag(G0 G0'); NULL = off
  bool cc = false;
  int64_t tile_begin = 0, tile_end = 0;   // tile_end == 0: every tile
  int64_t slice_tiles() const {
    const int nT = static_cast<int>(fbt::n_tiles(p));
    const int64_t hi = tile_end > 0 ? tile_end : fbt::n_slots(p);
    return fbt::cidx(hi, nT) - fbt::cidx(tile_begin, nT);
  }
};

// row posterior (shrunk G0, cpp:564-597) from inputs already resident in HBM
static int posterior_from_inputs(fable_ctx* ctx, const Inputs& in, double gamma0, double delta0sq, int kEst,
                                 double varInflation, fable_posterior** out) {
  FB_REQUIRE(ctx, varInflation >= 0.0 && std::isfinite(varInflation), "varInflation must be finite and >= 0");
  FB_REQUIRE(ctx, gamma0 + static_cast<double>(in.n) > 2.0, "gamma0 + n must exceed 2");
  fable_posterior* post = new (std::nothrow) fable_posterior();
  if (!post) return ctx->fail(FABLE_ERR_NOMEM, "out of host memory");
  post->ctx = ctx; post->n = in.n; post->p = in.p; post->pp = round_up(in.p, 64); post->k = kEst;
  post->KP = static_cast<int>(round_up(kEst, 4));
  post->gamma0 = gamma0; post->delta0sq = delta0sq; post->varInflation = varInflation;
  const int rc = row_posterior(ctx, in, kEst, gamma0, delta0sq, /*shrunk=*/1, post->rp);
  if (rc != FABLE_OK) { delete post; return rc; }
  post->a_n = std::sqrt(varInflation) / std::sqrt(post->rp.w);                 // cpp:597
  *out = post;
  return FABLE_OK;
}

static int sampler_outputs(fable_posterior* post, int64_t MC, const double* gam_inj, const double* z_inj, int64_t m0,
                           double* LambdaSamples, double* SigmaSqSamples, double* SigmaPostMean, double* SigmaPlugIn,
                           double* tausq, double* G, double* psi_sum, double* psi_sumsq);

// Thin SVD of a column-sharded Y (SURVEY 8(e) "Gram/SVD"; n <= p_total, so the Gram side is n): every rank forms the
// n x n Gram of its own columns on the FP64 tensor cores, the Grams are summed by the caller's all-reduce (the only
// O(n^2) exchange), the eigensolver runs replicated on bit-identical input -- so every rank holds the same U -- and
// V = Y'U D^-1 is formed for the local columns only.  d and the sign convention (largest |entry| of each V column
// positive, first index on ties) need the whole column: one all-reduce of r partial sums of squares and one of the
// per-rank (max |v|, global row, value) triples.
static int svd_sharded(fable_ctx* ctx, const double* dY, int64_t n, int64_t p_loc, double* dU, double* dd, double* dV) {
  const int64_t r = n;
  DevBuf G, W, tmp;
  FB_CUDA(ctx, G.alloc(static_cast<size_t>(n) * n * sizeof(double)));
  FB_CUDA(ctx, W.alloc(static_cast<size_t>(n) * sizeof(double)));
  FB_CUDA(ctx, tmp.alloc(static_cast<size_t>(3) * r * sizeof(double)));
  FB_TRY(fb_syrk(ctx, 0, n, p_loc, 1.0, dY, n, G.d(), n));                          // local Y Y'
  FB_TRY(shard_sum_dev(ctx, G.d(), n * n));                                          // the only O(n^2) exchange
  int sweeps = 0;
  int64_t defl = 0;      // triplets deflated at the eigensolver's noise floor (rank-deficient Y): d = 0, U completed, V zero
  FB_TRY(fb_syevj(ctx, G.d(), n, W.d(), dU, &sweeps, &defl));                       // replicated, identical bits
  FB_TRY(fb_gemm(ctx, 1, 1, p_loc, r, n, 1.0, dY, n, dU, n, dV, p_loc));           // T = Y_loc' U
  // d_l = || T[:, l] || over all ranks
  FB_TRY(fb_col_sumsq(ctx, dV, p_loc, p_loc, r, tmp.d()));
  std::vector<double> ss(r);
  FB_CUDA(ctx, cudaMemcpyAsync(ss.data(), tmp.d(), r * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  FB_TRY(shard_sum(ctx, ss.data(), r));
  // sign convention: gather every rank's candidate by summing into disjoint slots
  FB_TRY(fb_col_absmax(ctx, dV, p_loc, p_loc, r, tmp.d(), tmp.d() + r, tmp.d() + 2 * r));
  std::vector<double> loc(static_cast<size_t>(3) * r), all(static_cast<size_t>(3) * r * ctx->shard_world, 0.0);
  FB_CUDA(ctx, cudaMemcpyAsync(loc.data(), tmp.d(), loc.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int64_t l = 0; l < r; ++l) {
    double* slot = all.data() + (static_cast<size_t>(ctx->shard_rank) * r + l) * 3;
    slot[0] = loc[l]; slot[1] = loc[r + l] + static_cast<double>(ctx->shard_off); slot[2] = loc[2 * r + l];
  }
  FB_TRY(shard_sum(ctx, all.data(), static_cast<int64_t>(all.size())));
  std::vector<double> scale(static_cast<size_t>(2) * r);
  for (int64_t l = 0; l < r; ++l) {
    double bm = -1.0, bi = 0.0, bv = 0.0;
    for (int g = 0; g < ctx->shard_world; ++g) {
      const double* slot = all.data() + (static_cast<size_t>(g) * r + l) * 3;
      if (slot[0] > bm || (slot[0] == bm && slot[1] < bi)) { bm = slot[0]; bi = slot[1]; bv = slot[2]; }
    }
    const double sgn = bv < 0.0 ? -1.0 : 1.0;
    const double nv = l >= r - defl ? 0.0 : std::sqrt(ss[l]);     // (a shard cannot complete V: the column is left zero)
    ss[l] = nv;
    scale[l] = nv > 0.0 ? sgn / nv : 0.0;     // V column: normalise and orient
    scale[r + l] = sgn;                       // U column: orient
  }
  FB_CUDA(ctx, cudaMemcpyAsync(tmp.d(), scale.data(), scale.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  FB_TRY(fb_col_scale(ctx, dV, p_loc, p_loc, r, tmp.d()));
  FB_TRY(fb_col_scale(ctx, dU, n, n, r, tmp.d() + r));
  FB_CUDA(ctx, cudaMemcpyAsync(dd, ss.data(), r * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // host vectors go out of scope
  return FABLE_OK;
}

extern "C" {

int fable_abi_version(void) { return FABLE_B200_ABI_VERSION; }

int fable_ctx_create(int device, uint64_t seed, fable_ctx** out) {
  if (!out) { g_create_error = "fable_ctx_create: NULL out pointer"; return FABLE_ERR_INVALID; }
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    (void)cudaGetLastError();
    g_create_error = std::string("no CUDA device available (libfable_b200 has no CPU fallback): ") +
                     (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return FABLE_ERR_CUDA;
  }
  if (device < 0 || device >= count) { g_create_error = "fable_ctx_create: device ordinal out of range"; return FABLE_ERR_INVALID; }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return FABLE_ERR_CUDA; }
  if (prop.major != 10) {
    g_create_error = std::string("device '") + prop.name + "' is compute capability " + std::to_string(prop.major) +
                     "." + std::to_string(prop.minor) + "; libfable_b200 carries sm_100a code only";
    return FABLE_ERR_CUDA;
  }
  fable_ctx* ctx = new (std::nothrow) fable_ctx();
  if (!ctx) { g_create_error = "out of host memory"; return FABLE_ERR_NOMEM; }
  ctx->device = device; ctx->seed = seed; ctx->sms = prop.multiProcessorCount;
  // the copy stream outranks the main stream: its small kernels (sample layouts, dense blocks on their way to the host) are
  // dispatched as soon as an SM has room instead of queueing behind the covariance kernel's backlog of CTAs
  int prio_lo = 0, prio_hi = 0;
  cudaSetDevice(device);
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaStreamCreateWithPriority(&ctx->copy_stream, cudaStreamNonBlocking, prio_hi)) != cudaSuccess) {
    g_create_error = cudaGetErrorString(e);
    delete ctx;
    return FABLE_ERR_CUDA;
  }
  fb_mem_register_stream(device, ctx->stream, true);
  fb_mem_register_stream(device, ctx->copy_stream, true);
  *out = ctx;
  return FABLE_OK;
}

void fable_ctx_destroy(fable_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();              // blocks released from now on carry no events of these streams
  fb_mem_register_stream(ctx->device, ctx->stream, false);
  fb_mem_register_stream(ctx->device, ctx->copy_stream, false);
  for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->ev_extra) cudaEventDestroy(e);
  if (ctx->ring) cudaFree(ctx->ring);
  for (int s = 0; s < 2; ++s) {
    if (ctx->stage[s]) cudaFreeHost(ctx->stage[s]);
    if (ctx->stage_ev[s]) cudaEventDestroy(ctx->stage_ev[s]);
  }
  if (ctx->gemm_ws) cudaFree(ctx->gemm_ws);
  fb_mem_trim(ctx->device);          // the block cache does not outlive the context that filled it
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
}

const char* fable_last_error(const fable_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
int fable_ctx_set_seed(fable_ctx* ctx, uint64_t seed) { if (!ctx) return FABLE_ERR_INVALID; ctx->seed = seed; return FABLE_OK; }
int fable_ctx_set_interrupt(fable_ctx* ctx, int (*poll)(void*), void* user) {
  if (!ctx) return FABLE_ERR_INVALID;
  ctx->poll = poll; ctx->poll_user = user;
  return FABLE_OK;
}
int64_t fable_ctx_launch_count(const fable_ctx* ctx) { return ctx ? ctx->launches : -1; }
int fable_ctx_trim_cache(fable_ctx* ctx) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  FB_CUDA(ctx, cudaDeviceSynchronize());
  fb_mem_trim(ctx->device);
  return FABLE_OK;
}
int fable_ctx_set_column_shard(fable_ctx* ctx, int64_t p_total, int64_t col_offset, int rank, int world,
                               fable_allreduce_sum_fn fn, void* user) {
  if (!ctx) return FABLE_ERR_INVALID;
  if (p_total == 0) {          // off
    ctx->shard_p = 0; ctx->shard_off = 0; ctx->shard_rank = 0; ctx->shard_world = 1; ctx->shard_fn = nullptr; ctx->shard_user = nullptr;
    ctx->shard_dev_fn = nullptr; ctx->shard_dev_user = nullptr;
    return FABLE_OK;
  }
  FB_REQUIRE(ctx, fn, "column shard: NULL all-reduce callback");
  FB_REQUIRE(ctx, world >= 1 && rank >= 0 && rank < world, "column shard: need 0 <= rank < world");
  FB_REQUIRE(ctx, p_total >= 1 && col_offset >= 0 && col_offset <= p_total, "column shard: need 0 <= col_offset <= p_total");
  ctx->shard_p = p_total; ctx->shard_off = col_offset; ctx->shard_rank = rank; ctx->shard_world = world;
  ctx->shard_fn = fn; ctx->shard_user = user;
  return FABLE_OK;
}

int fable_ctx_set_column_shard_dev(fable_ctx* ctx, fable_allreduce_sum_dev_fn fn, void* user) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_REQUIRE(ctx, shard_on(ctx) || !fn, "column shard: set the shard (fable_ctx_set_column_shard) before its device all-reduce");
  ctx->shard_dev_fn = fn; ctx->shard_dev_user = user;
  return FABLE_OK;
}

int fable_ctx_synchronize(fable_ctx* ctx) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FABLE_OK;
}
void* fable_ctx_stream(fable_ctx* ctx) { return ctx ? static_cast<void*>(ctx->stream) : nullptr; }

int fable_ctx_kernel_timing(fable_ctx* ctx, int enable) {
  if (!ctx) return FABLE_ERR_INVALID;
  ctx->timing = enable != 0;
  ctx->ev_used = 0;
  for (cudaEvent_t e : ctx->ev_extra) cudaEventDestroy(e);
  ctx->ev_extra.clear();
  return FABLE_OK;
}

int fable_ctx_stage_timing(fable_ctx* ctx, int enable) {
  if (!ctx) return FABLE_ERR_INVALID;
  ctx->stage_timing = enable != 0;
  for (double& x : ctx->stage_ms) x = 0.0;
  return FABLE_OK;
}

int fable_ctx_stage_times(const fable_ctx* ctx, double* ms, int count) {
  if (!ctx || !ms || count < 0) return FABLE_ERR_INVALID;
  for (int i = 0; i < count; ++i) ms[i] = i < 8 ? ctx->stage_ms[i] : 0.0;
  return FABLE_OK;
}

int fable_ctx_kernel_time(fable_ctx* ctx, double* total_ms, int64_t* launches) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  double tot = 0.0;
  for (size_t i = 0; i + 1 < ctx->ev_used; i += 2) {
    float ms = 0.f;
    FB_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_pool[i], ctx->ev_pool[i + 1]));
    tot += ms;
  }
  for (size_t i = 0; i + 1 < ctx->ev_extra.size(); i += 2) {
    float ms = 0.f;
    FB_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_extra[i], ctx->ev_extra[i + 1]));
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = static_cast<int64_t>(ctx->ev_used / 2);
  return FABLE_OK;
}

int fable_ctx_measure_fp64_peak(fable_ctx* ctx, double* dmma_tflops) {
  if (!ctx || !dmma_tflops) return FABLE_ERR_INVALID;
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  return fb_dmma_peak(ctx, dmma_tflops);
}

int fable_ctx_set_psi_ring(fable_ctx* ctx, int slots) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_REQUIRE(ctx, slots >= 0, "psi ring: slots must be >= 0");
  ctx->ring_slots = slots;
  if (slots == 0 && ctx->ring) {
    FB_CUDA(ctx, cudaSetDevice(ctx->device));
    FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->ring); ctx->ring = nullptr; ctx->ring_bytes = 0;
  }
  return FABLE_OK;
}

int fable_ctx_psi_ring(fable_ctx* ctx, double** dev_ring, int* slots) {
  if (!ctx) return FABLE_ERR_INVALID;
  if (dev_ring) *dev_ring = ctx->ring;
  if (slots) *slots = ctx->ring ? ctx->ring_slots : 0;
  return FABLE_OK;
}

// ---- a1 / a2 -------------------------------------------------------------------------------------
int fable_svd(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, double* U, double* d, double* V) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_REQUIRE(ctx, Y && U && d && V, "svd: NULL pointer");
  FB_REQUIRE(ctx, n >= 1 && p >= 1, "svd: Y must have at least one row and one column");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  FB_REQUIRE(ctx, !shard_on(ctx) || (n <= ctx->shard_p && ctx->shard_off + p <= ctx->shard_p),
             "svd under a column shard: needs n <= p_total and offset + local columns <= p_total");
  const int64_t r = shard_on(ctx) ? n : (n < p ? n : p);
  DevBuf dY, dU, dd, dV;
  FB_TRY(upload(ctx, dY, Y, static_cast<size_t>(n) * p));
  FB_CUDA(ctx, dU.alloc(static_cast<size_t>(n) * r * sizeof(double)));
  FB_CUDA(ctx, dV.alloc(static_cast<size_t>(p) * r * sizeof(double)));
  FB_CUDA(ctx, dd.alloc(static_cast<size_t>(r) * sizeof(double)));
  if (shard_on(ctx)) FB_TRY(svd_sharded(ctx, dY.d(), n, p, dU.d(), dd.d(), dV.d()));
  else FB_TRY(fb_svd(ctx, dY.d(), n, p, dU.d(), dd.d(), dV.d()));
  FB_TRY(download(ctx, U, dU.d(), static_cast<size_t>(n) * r));
  FB_TRY(download(ctx, d, dd.d(), static_cast<size_t>(r)));
  FB_TRY(download(ctx, V, dV.d(), static_cast<size_t>(p) * r));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FABLE_OK;
}

// CPPsvd with flag 1 / 2 (cpp:52-58: arma::svd, the FULL decomposition): U n x n, d r, V p x p.  The thin factors come from
// fable_svd's path; the remaining columns of the longer side complete an orthonormal basis (any such completion is a
// valid full SVD: those columns meet zero singular values only).
int fable_svd_full(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, double* U, double* d, double* V) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_NO_SHARD(ctx, "the full SVD");
  FB_REQUIRE(ctx, Y && U && d && V, "svd: NULL pointer");
  FB_REQUIRE(ctx, n >= 1 && p >= 1, "svd: Y must have at least one row and one column");
  const int64_t r = n < p ? n : p, big = n < p ? p : n;
  FB_REQUIRE(ctx, big - r <= 8192, "full SVD: more than 8192 basis vectors to complete; use the economy form (flag 3 / 4)");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  DevBuf dY, dU, dd, dV;
  FB_TRY(upload(ctx, dY, Y, static_cast<size_t>(n) * p));
  FB_CUDA(ctx, dU.alloc(static_cast<size_t>(n) * n * sizeof(double)));
  FB_CUDA(ctx, dV.alloc(static_cast<size_t>(p) * p * sizeof(double)));
  FB_CUDA(ctx, dd.alloc(static_cast<size_t>(r) * sizeof(double)));
  FB_TRY(fb_svd(ctx, dY.d(), n, p, dU.d(), dd.d(), dV.d()));              // first r columns of each (dense leading dimensions)
  FB_TRY(fb_complete_orthonormal(ctx, dU.d(), n, n, r, n));
  FB_TRY(fb_complete_orthonormal(ctx, dV.d(), p, p, r, p));
  FB_TRY(download(ctx, U, dU.d(), static_cast<size_t>(n) * n));
  FB_TRY(download(ctx, d, dd.d(), static_cast<size_t>(r)));
  FB_TRY(download(ctx, V, dV.d(), static_cast<size_t>(p) * p));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FABLE_OK;
}

int fable_kmax(const double* d, int64_t r, double maxProp, int* kMax) {
  if (!d || !kMax || r < 1) return FABLE_ERR_INVALID;
  // R accumulates sum() and cumsum() sequentially in long double and rounds each result to double.
  long double acc = 0.0L;
  for (int64_t i = 0; i < r; ++i) acc += d[i];               // sum(svalsY)
  const double total = static_cast<double>(acc);
  acc = 0.0L;
  for (int64_t i = 0; i < r; ++i) {                          // cumsum(svalsY)/sum >= maxProp
    acc += d[i];
    if (static_cast<double>(acc) / total >= maxProp) { *kMax = static_cast<int>(i + 1); return FABLE_OK; }
  }
  return FABLE_ERR_INVALID;                                  // R: min(integer(0)) = Inf + warning
}

// ---- f4: CPPLogLikelihoodEval (cpp:95-125) ----------------------------------------------------------
int fable_log_likelihood_eval(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, const double* M, const double* Lambda,
                              int k, const double* SigmaSq, double* out) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_NO_SHARD(ctx, "CPPLogLikelihoodEval");
  FB_REQUIRE(ctx, Y && M && Lambda && SigmaSq && out, "log_likelihood: NULL pointer");
  FB_REQUIRE(ctx, n >= 1 && p >= 1 && k >= 1, "log_likelihood: need n, p, k >= 1");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  DevBuf dY, dM, dL, dS, dR;
  FB_TRY(upload(ctx, dY, Y, static_cast<size_t>(n) * p));
  FB_TRY(upload(ctx, dM, M, static_cast<size_t>(n) * k));
  FB_TRY(upload(ctx, dL, Lambda, static_cast<size_t>(p) * k));
  FB_TRY(upload(ctx, dS, SigmaSq, static_cast<size_t>(p)));
  FB_CUDA(ctx, dR.alloc(static_cast<size_t>(p) * sizeof(double)));
  // residual column sums of squares of Y - M Lambda' (cpp:113): the row-pass kernels with U := M, c := Lambda
  if (k <= KSMALL) FB_TRY(fb_rowstats(ctx, dY.d(), n, p, n, dM.d(), n, dL.d(), p, k, nullptr, dR.d()));
  else FB_TRY(fb_gemm_resid(ctx, n, p, k, dY.d(), n, dM.d(), n, dL.d(), p, dR.d()));
  double sum_log_sd = 0.0, sum_ros = 0.0;
  FB_TRY(fb_loglik_terms(ctx, dR.d(), dS.d(), p, &sum_log_sd, &sum_ros));
  *out = (-static_cast<double>(n) * sum_log_sd) + (-0.5 * sum_ros);                 // cpp:116-120
  return FABLE_OK;
}

// ---- a3 / a4 / a5 --------------------------------------------------------------------------------
int fable_criterion(fable_ctx* ctx, int kInd, const double* Y, int64_t n, int64_t p, const double* U,
                    const double* V, const double* d, int64_t r, double* out) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_REQUIRE(ctx, out, "criterion: NULL output");
  FB_REQUIRE(ctx, kInd >= 1 && kInd <= r, "criterion: kInd must be in 1..length(svalsY)");
  Inputs in;
  FB_TRY(upload_inputs(ctx, Y, n, p, U, V, d, r, kInd, in));
  RowScratch sc;
  return criterion_dev(ctx, in, kInd, sc, out);
}

int fable_bisection(fable_ctx* ctx, int lower, int upper, const double* Y, int64_t n, int64_t p,
                    const double* U, const double* V, const double* d, int64_t r, int* lo, int* hi) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_REQUIRE(ctx, lo && hi, "bisection: NULL output");
  FB_REQUIRE(ctx, lower >= 1 && upper >= lower && upper <= r, "bisection: need 1 <= lower <= upper <= length(svalsY)");
  Inputs in;
  FB_TRY(upload_inputs(ctx, Y, n, p, U, V, d, r, upper, in));
  CritCache cc(ctx, in);
  FB_TRY(cc.prepare(upper));
  return bisection(cc, lower, upper, lo, hi);
}

int fable_rank_estimator(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, const double* U,
                         const double* V, const double* d, int64_t r, int kMax, int* kEst) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_REQUIRE(ctx, kEst, "rank estimator: NULL output");
  FB_REQUIRE(ctx, kMax >= 1 && kMax <= r, "kMax must be in 1..length(svalsY)");
  Inputs in;
  FB_TRY(upload_inputs(ctx, Y, n, p, U, V, d, r, kMax, in));
  return rank_dev(ctx, in, kMax, kEst);
}

// the criterion / rank rule of the pure-R prototype (R/FABLE_code.R:32-186), used by CCFABLE_DirectSampler (:221)
int fable_criterion_proto(fable_ctx* ctx, int kInd, const double* Y, int64_t n, int64_t p, const double* U, const double* V,
                          const double* d, int64_t r, double* out) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_REQUIRE(ctx, out, "criterion: NULL output");
  FB_REQUIRE(ctx, kInd >= 1 && kInd <= r, "criterion: kInd must be in 1..length(svalsY)");
  Inputs in;
  FB_TRY(upload_inputs(ctx, Y, n, p, U, V, d, r, kInd, in));
  RowScratch sc;
  return criterion_dev(ctx, in, kInd, sc, out, /*variant=*/1);
}

int fable_rank_estimator_proto(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, const double* U, const double* V,
                               const double* d, int64_t r, int kMax, int* kEst) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_REQUIRE(ctx, kEst, "rank estimator: NULL output");
  FB_REQUIRE(ctx, kMax >= 1 && kMax <= r, "kMax must be in 1..length(svalsY)");
  Inputs in;
  FB_TRY(upload_inputs(ctx, Y, n, p, U, V, d, r, kMax, in));
  return rank_dev(ctx, in, kMax, kEst, /*variant=*/1);
}

// ---- a6 ------------------------------------------------------------------------------------------
int fable_hyper_parameters(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, const double* U,
                           const double* V, const double* d, int64_t r, int kEst, double* sig, double* G0,
                           double* tausq, double* G) {
  if (!ctx) return FABLE_ERR_INVALID;
  if (G) FB_NO_SHARD(ctx, "the dense G = G0 G0' of FABLEHyperParameters");
  FB_REQUIRE(ctx, kEst >= 1 && kEst <= r, "kEst must be in 1..length(svalsY)");
  Inputs in;
  FB_TRY(upload_inputs(ctx, Y, n, p, U, V, d, r, kEst, in));
  RowPosterior rp;
  FB_TRY(row_posterior(ctx, in, kEst, 1.0, 1.0, /*shrunk=*/0, rp));
  FB_TRY(download(ctx, sig, rp.sig.d(), p));
  FB_TRY(download(ctx, G0, rp.G0.d(), static_cast<size_t>(p) * kEst));
  if (tausq) *tausq = rp.tausq;
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (G) FB_TRY(rankk_cov_to_host(ctx, rp.G0.d(), p, nullptr, p, kEst, G));   // cpp:391
  return FABLE_OK;
}

// ---- a7 ------------------------------------------------------------------------------------------
static int cov_correct_host(fable_ctx* ctx, const double* sig, const double* G, int64_t p, double* out, int upper) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_REQUIRE(ctx, sig && G && out && p >= 1, "cov_correct: bad arguments");
  FB_REQUIRE(ctx, p <= 65535, "cov_correct (dense form): p too large; use fable_var_inflation");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  DevBuf ds, dG, dO;
  FB_TRY(upload(ctx, ds, sig, p));
  FB_TRY(upload(ctx, dG, G, static_cast<size_t>(p) * p));
  const size_t outn = upper ? static_cast<size_t>(p) * (p + 1) / 2 : static_cast<size_t>(p) * p;
  FB_CUDA(ctx, dO.alloc(outn * sizeof(double)));
  FB_TRY(fb_cov_correct_dense(ctx, ds.d(), dG.d(), p, dO.d(), upper));
  FB_TRY(download(ctx, out, dO.d(), outn));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FABLE_OK;
}
int fable_cov_correct_matrix(fable_ctx* ctx, const double* sig, const double* G, int64_t p, double* Bupper) {
  return cov_correct_host(ctx, sig, G, p, Bupper, 1);
}
int fable_cov_correct_full(fable_ctx* ctx, const double* sig, const double* G, int64_t p, double* B) {
  return cov_correct_host(ctx, sig, G, p, B, 0);
}

static int var_inflation_from_sums(double su, double sd, int64_t p, int variant, double* out) {
  const double tri = static_cast<double>(p) * (static_cast<double>(p) + 1.0) / 2.0;
  const double total = variant == 0 ? (2.0 * su - sd) : su;   // full-matrix sum (wrapper:44) vs upper triangle (script)
  const double m = total / tri;
  *out = m * m;
  return FABLE_OK;
}

int fable_var_inflation(fable_ctx* ctx, const double* sig, const double* G0, int64_t p, int k, int variant,
                        double* out) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_REQUIRE(ctx, sig && G0 && out && p >= 1 && k >= 1, "var_inflation: bad arguments");
  FB_REQUIRE(ctx, variant == 0 || variant == 1, "var_inflation: variant must be 0 (wrapper) or 1 (script)");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  DevBuf ds, dG0;
  FB_TRY(upload(ctx, ds, sig, p));
  FB_TRY(upload(ctx, dG0, G0, static_cast<size_t>(p) * k));
  double su, sd;
  FB_TRY(fb_var_inflation(ctx, ds.d(), dG0.d(), p, p, k, &su, &sd));
  return var_inflation_from_sums(su, sd, p, variant, out);
}

// ---- a8 ------------------------------------------------------------------------------------------
int fable_post_mean(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, double gamma0, double delta0sq,
                    const double* U, const double* V, const double* d, int64_t r, int kMax, double* Cov,
                    int* kEst, double* G0s, double* SigmaHat) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_NO_SHARD(ctx, "CPPFABLEPostMean (dense CovPostMean)");
  FB_REQUIRE(ctx, kEst, "post_mean: NULL kEst");
  FB_REQUIRE(ctx, kMax >= 1 && kMax <= r, "kMax must be in 1..length(svalsY)");
  Inputs in;
  FB_TRY(upload_inputs(ctx, Y, n, p, U, V, d, r, kMax, in));
  int k = 1;
  FB_TRY(rank_dev(ctx, in, kMax, &k));                                           // cpp:457
  RowPosterior rp;
  FB_TRY(row_posterior(ctx, in, k, gamma0, delta0sq, /*shrunk=*/1, rp));
  *kEst = k;
  FB_TRY(download(ctx, G0s, rp.G0.d(), static_cast<size_t>(p) * k));
  FB_TRY(download(ctx, SigmaHat, rp.sigmahat.d(), p));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (Cov) FB_TRY(rankk_cov_to_host(ctx, rp.G0.d(), p, rp.sigmahat.d(), p, k, Cov));   // cpp:491, 509-512
  return FABLE_OK;
}

// ---- posterior object ----------------------------------------------------------------------------
int fable_posterior_create(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, double gamma0,
                           double delta0sq, const double* U, const double* V, const double* d, int64_t r,
                           int kEst, double varInflation, fable_posterior** out) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_NO_SHARD(ctx, "the covariance pass");
  FB_REQUIRE(ctx, out, "posterior_create: NULL out");
  *out = nullptr;
  FB_REQUIRE(ctx, kEst >= 1 && kEst <= r, "kEst must be in 1..length(svalsY)");
  Inputs in;
  FB_TRY(upload_inputs(ctx, Y, n, p, U, V, d, r, kEst, in));
  return posterior_from_inputs(ctx, in, gamma0, delta0sq, kEst, varInflation, out);
}

// The per-variable conjugate posterior of A.2 (cpp:371-391 unshrunk / cpp:470-499 = 564-592 shrunk) as its own entry
// point: under a column shard every rank computes the rows of its own variables (tauSq is completed by the
// all-reduce), the ranks exchange these O(p k) numbers, and each builds its sampler from all rows.
int fable_row_posterior(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, double gamma0, double delta0sq,
                        const double* U, const double* V, const double* d, int64_t r, int k, int shrunk, double* G0,
                        double* gnd, double* sig, double* sigmahat, double* tausq) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_REQUIRE(ctx, k >= 1 && k <= r, "k must be in 1..length(svalsY)");
  FB_REQUIRE(ctx, G0 && gnd && sig && sigmahat, "row posterior: NULL output");
  FB_REQUIRE(ctx, gamma0 + static_cast<double>(n) > 2.0, "gamma0 + n must exceed 2");
  Inputs in;
  FB_TRY(upload_inputs(ctx, Y, n, p, U, V, d, r, k, in));
  RowPosterior rp;
  FB_TRY(row_posterior(ctx, in, k, gamma0, delta0sq, shrunk ? 1 : 0, rp));
  FB_TRY(download(ctx, G0, rp.G0.d(), static_cast<size_t>(p) * k));
  FB_TRY(download(ctx, gnd, rp.gnd.d(), p));
  FB_TRY(download(ctx, sig, rp.sig.d(), p));
  FB_TRY(download(ctx, sigmahat, rp.sigmahat.d(), p));
  if (tausq) *tausq = rp.tausq;
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FABLE_OK;
}

int fable_posterior_create_from_rows(fable_ctx* ctx, int64_t n, int64_t p, int k, double gamma0, double delta0sq,
                                     const double* G0, const double* gnd, const double* sig, const double* sigmahat,
                                     double tausq, double varInflation, fable_posterior** out) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_REQUIRE(ctx, out, "posterior_create_from_rows: NULL out");
  *out = nullptr;
  FB_REQUIRE(ctx, G0 && gnd && sig && sigmahat, "posterior_create_from_rows: NULL input");
  FB_REQUIRE(ctx, n >= 1 && p >= 1 && k >= 1, "posterior_create_from_rows: n, p, k must be positive");
  FB_REQUIRE(ctx, tausq > 0.0 && std::isfinite(tausq), "tauSq must be positive and finite");
  FB_REQUIRE(ctx, varInflation >= 0.0 && std::isfinite(varInflation), "varInflation must be finite and >= 0");
  FB_REQUIRE(ctx, gamma0 + static_cast<double>(n) > 2.0, "gamma0 + n must exceed 2");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  fable_posterior* post = new (std::nothrow) fable_posterior();
  if (!post) return ctx->fail(FABLE_ERR_NOMEM, "out of host memory");
  struct Guard { fable_posterior* p; ~Guard() { delete p; } } guard{post};
  post->ctx = ctx; post->n = n; post->p = p; post->pp = round_up(p, 64); post->k = k;
  post->KP = static_cast<int>(round_up(k, 4));
  post->gamma0 = gamma0; post->delta0sq = delta0sq; post->varInflation = varInflation;
  post->rp.tausq = tausq;
  post->rp.w = static_cast<double>(n) + 1.0 / tausq;                          // cpp:584
  post->rp.gamma_n = gamma0 + static_cast<double>(n);                          // cpp:586
  FB_TRY(upload(ctx, post->rp.G0, G0, static_cast<size_t>(p) * k));
  FB_TRY(upload(ctx, post->rp.gnd, gnd, p));
  FB_TRY(upload(ctx, post->rp.sig, sig, p));
  FB_TRY(upload(ctx, post->rp.sigmahat, sigmahat, p));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  post->a_n = std::sqrt(varInflation) / std::sqrt(post->rp.w);                 // cpp:597
  guard.p = nullptr;
  *out = post;
  return FABLE_OK;
}

void fable_posterior_destroy(fable_posterior* post) {
  if (!post) return;
  cudaSetDevice(post->ctx->device);
  delete post;
}

int fable_posterior_get(fable_posterior* post, double* G0s, double* gnd, double* sig, double* scalars) {
  if (!post) return FABLE_ERR_INVALID;
  fable_ctx* ctx = post->ctx;
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  FB_TRY(download(ctx, G0s, post->rp.G0.d(), static_cast<size_t>(post->p) * post->k));
  FB_TRY(download(ctx, gnd, post->rp.gnd.d(), post->p));
  FB_TRY(download(ctx, sig, post->rp.sig.d(), post->p));
  if (scalars) { scalars[0] = post->rp.tausq; scalars[1] = post->a_n; scalars[2] = post->rp.gamma_n; scalars[3] = post->rp.w; }
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FABLE_OK;
}

static SampleArgs sample_args(const fable_posterior* post, const double* dgam, const double* dz, int64_t m_inj0) {
  SampleArgs a;
  a.G0 = post->rp.G0.d(); a.ldg = post->p; a.gnd = post->rp.gnd.d(); a.a_n = post->a_n;
  a.shape = 0.5 * post->rp.gamma_n;                                            // cpp:604
  a.p = post->p; a.k = post->k; a.seed = post->ctx->seed; a.gam = dgam; a.z = dz; a.m_inj0 = m_inj0;
  return a;
}

// The sum / sumsq accumulators hold the upper 64 x 64 tiles of the posterior's tile slice only, tile-major and compact
// (tiles.cuh): p(p+1)/2-ish doubles instead of p^2, a rank of the tile partition holds just its share, and a slice of
// the enumeration is one contiguous range (the payload of the draw partition's NCCL sum needs no packing).
static int ensure_accum(fable_posterior* post) {
  fable_ctx* ctx = post->ctx;
  if (post->acc_sum.ptr) return FABLE_OK;
  post->acc_tiles = post->slice_tiles();
  const size_t bytes = static_cast<size_t>(post->acc_tiles) * fbt::TILE_DOUBLES * sizeof(double);
  FB_CUDA(ctx, post->acc_sum.alloc(bytes));
  FB_CUDA(ctx, post->acc_sq.alloc(bytes));
  FB_CUDA(ctx, cudaMemsetAsync(post->acc_sum.ptr, 0, bytes, ctx->stream));
  FB_CUDA(ctx, cudaMemsetAsync(post->acc_sq.ptr, 0, bytes, ctx->stream));
  return FABLE_OK;
}

static int ensure_partials(fable_posterior* post);
static int fold_range(fable_posterior* post, int64_t c_lo, int64_t c_hi, bool consume);
static int fold_partials(fable_posterior* post) { return fold_range(post, 0, -1, true); }

static int draw_batch_impl(fable_posterior* post, int64_t m0, int B, double* dev_psi, int psi_slots, bool compact,
                           int accumulate, const double* dev_gam, const double* dev_z, int64_t launch_begin = 0,
                           int64_t launch_end = 0, int partial_slot = -1) {
  if (!post) return FABLE_ERR_INVALID;
  fable_ctx* ctx = post->ctx;
  FB_REQUIRE(ctx, B >= 1 && m0 >= 0, "draw_batch: need B >= 1 and m0 >= 0");
  FB_REQUIRE(ctx, dev_psi || accumulate, "draw_batch: nothing to do");
  // one CTA owns a tile for the whole batch (tile-major order), so every draw of the launch needs
  // its own slot: a shorter ring would be overwritten before anyone could read it
  FB_REQUIRE(ctx, !dev_psi || psi_slots >= B, "draw_batch: psi_slots must be >= B (one p x p slot per draw of the batch)");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (B > post->Bcap) {
    FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // one extra leading slot: G0 as the pseudo-draw of the entry-wise correction (unused otherwise)
    FB_CUDA(ctx, post->Lw.alloc(static_cast<size_t>(B + 1) * post->KP * post->pp * sizeof(double)));
    FB_CUDA(ctx, post->s2w.alloc(static_cast<size_t>(B) * post->pp * sizeof(double)));
    post->Bcap = B;
  }
  if (accumulate) FB_TRY(ensure_accum(post));
  double* draws = post->Lw.d() + static_cast<size_t>(post->KP) * post->pp;      // slots 1..B
  FB_TRY(fb_sample_work(ctx, sample_args(post, dev_gam, dev_z, m0), m0, B, post->pp, post->KP, draws, post->s2w.d()));
  PsiArgs a;
  if (post->cc) {
    FB_REQUIRE(ctx, dev_psi, "draw_batch: the entry-wise correction applies to materialised draws");
    FB_TRY(fb_pad_rows(ctx, post->rp.G0.d(), post->p, post->p, post->pp, post->k, post->KP, post->Lw.d()));   // slot 0 = G0
    a.cc_gdiag = post->cc_gdiag.d(); a.cc_sig = post->rp.sig.d();
  }
  a.Lw = post->cc ? post->Lw.d() : draws; a.s2w = post->s2w.d(); a.pp = post->pp; a.p = post->p; a.k = post->k; a.KP = post->KP;
  a.B = B; a.psi = dev_psi; a.psi_slots = psi_slots; a.slot0 = 0; a.out_compact = compact;
  a.acc_sum = accumulate ? post->acc_sum.d() : nullptr; a.acc_sq = accumulate ? post->acc_sq.d() : nullptr;
  a.store_begin = post->tile_begin; a.store_end = post->tile_end;
  a.tile_begin = launch_begin; a.tile_end = launch_end;
  // materialising + accumulating batch over the whole slice: its sums go to the next partial buffer (ensure_partials);
  // partial_slot >= 0: the caller (chunk-major tail) names the slot and folds itself
  bool to_partial = false;
  if (accumulate && dev_psi && !post->cc) {
    FB_TRY(ensure_partials(post));
    if (post->part_slots > 0 && (partial_slot >= 0 || launch_end == 0)) {
      const int slot = partial_slot >= 0 ? partial_slot : post->part_pending;
      const int64_t cnt = post->acc_tiles * fbt::TILE_DOUBLES;
      a.acc_sum = post->part.d() + static_cast<int64_t>(slot) * 2 * cnt;
      a.acc_sq = a.acc_sum + cnt;
      a.acc_fresh = true;
      to_partial = partial_slot < 0;
    } else if (post->part_pending > 0) {
      FB_TRY(fold_partials(post));                 // a classic launch adds onto the accumulators: they must be current
    }
  } else if (accumulate && post->part_pending > 0 && post->cc) {
    FB_TRY(fold_partials(post));
  }
  FB_TRY(fb_psi(ctx, a));
  if (to_partial && ++post->part_pending == post->part_slots) FB_TRY(fold_partials(post));
  return FABLE_OK;
}

// The covariance kernel can add a batch into the HBM accumulators itself (load the running tiles, store old + batch), but
// DRAM reads mixed into its saturated write stream cost ~10 x their streaming price (profiles/r01_psi_lab.md: 3.2 GB of
// accumulator reads = 2.0 ms of a 23 ms launch at C3).  Instead every materialising batch STORES its own sums into the
// next of a few partial buffers (pure writes, at streaming cost) and a small read-mostly kernel folds them into the
// accumulators every part_slots batches -- and before anything reads, resets or re-slices them.  Sums are folded in batch
// order: deterministic; they differ from the classic path in the last bits only (association).  Off (classic) when the
// buffers would not fit: FABLE_B200_ACC_PARTIALS = slots (default 6, 0 = off), at most ~1/8 of the device memory.
static int ensure_partials(fable_posterior* post) {
  if (post->part_slots >= 0) return FABLE_OK;
  post->part_slots = 0;
  const char* e = getenv("FABLE_B200_ACC_PARTIALS");
  int want = e ? atoi(e) : 6;
  // fewer than four buffers do not pay: the fold re-reads the accumulators once per fold (two buffers at C4 on two GPUs:
  // 430 draws/s against 487 with the classic in-kernel update)
  const int least = e ? 2 : 4;
  if (want < least) return FABLE_OK;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { (void)cudaGetLastError(); return FABLE_OK; }
  const size_t per = static_cast<size_t>(2) * post->acc_tiles * fbt::TILE_DOUBLES * sizeof(double);
  if (per == 0) return FABLE_OK;
  while (want >= least && static_cast<size_t>(want) * per > total_b / 8) --want;
  if (want < least || static_cast<size_t>(want) * per + (size_t(4) << 30) > free_b) return FABLE_OK;
  if (post->part.alloc(static_cast<size_t>(want) * per) != cudaSuccess) { (void)cudaGetLastError(); return FABLE_OK; }
  post->part_slots = want;
  return FABLE_OK;
}

// fold the pending partial sums of tiles [c_lo, c_hi) (compact indices inside the slice; default: all) into the accumulators
static int fold_range(fable_posterior* post, int64_t c_lo, int64_t c_hi, bool consume) {
  if (post->part_pending <= 0) return FABLE_OK;
  fable_ctx* ctx = post->ctx;
  const int64_t cnt = post->acc_tiles * fbt::TILE_DOUBLES;
  if (c_hi < 0) c_hi = post->acc_tiles;
  const int64_t b = c_lo * fbt::TILE_DOUBLES, e = c_hi * fbt::TILE_DOUBLES;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  if (ctx->timing) {            // the fold is part of the covariance pass: its time goes on the same bill (fable_ctx_kernel_time)
    FB_CUDA(ctx, cudaEventCreate(&t0));
    FB_CUDA(ctx, cudaEventCreate(&t1));
    ctx->ev_extra.push_back(t0); ctx->ev_extra.push_back(t1);
    FB_CUDA(ctx, cudaEventRecord(t0, ctx->stream));
  }
  FB_TRY(fb_fold_partials(ctx, post->acc_sum.d(), post->part.d(), post->part_pending, 2 * cnt, b, e));
  FB_TRY(fb_fold_partials(ctx, post->acc_sq.d(), post->part.d() + cnt, post->part_pending, 2 * cnt, b, e));
  if (t1) FB_CUDA(ctx, cudaEventRecord(t1, ctx->stream));
  if (consume) post->part_pending = 0;
  return FABLE_OK;
}

int fable_posterior_draw_batch_range(fable_posterior* post, int64_t m0, int B, double* dev_psi, int psi_slots, int accumulate,
                                     const double* dev_gam, const double* dev_z, int64_t t_begin, int64_t t_end) {
  if (!post) return FABLE_ERR_INVALID;
  FB_REQUIRE(post->ctx, t_end > t_begin && t_begin >= 0, "draw_batch_range: empty launch range");
  return draw_batch_impl(post, m0, B, dev_psi, psi_slots, false, accumulate, dev_gam, dev_z, t_begin, t_end);
}

int fable_posterior_draw_batch(fable_posterior* post, int64_t m0, int B, double* dev_psi, int psi_slots,
                               int accumulate, const double* dev_gam, const double* dev_z) {
  return draw_batch_impl(post, m0, B, dev_psi, psi_slots, false, accumulate, dev_gam, dev_z);
}

int fable_posterior_draw_batch_tiles(fable_posterior* post, int64_t m0, int B, double* dev_psi_tiles, int psi_slots,
                                     int accumulate, const double* dev_gam, const double* dev_z) {
  return draw_batch_impl(post, m0, B, dev_psi_tiles, psi_slots, true, accumulate, dev_gam, dev_z);
}

int fable_posterior_set_entrywise_correction(fable_posterior* post, int on) {
  if (!post) return FABLE_ERR_INVALID;
  fable_ctx* ctx = post->ctx;
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (on && !post->cc_gdiag.ptr) {
    FB_CUDA(ctx, post->cc_gdiag.alloc(static_cast<size_t>(post->p) * sizeof(double)));
    FB_TRY(fb_gdiag(ctx, post->rp.G0.d(), post->p, post->p, post->k, post->cc_gdiag.d()));     // diag(G), R:193
  }
  post->cc = on != 0;
  return FABLE_OK;
}

int fable_posterior_tile_slots(fable_posterior* post, int64_t* slots) {
  if (!post || !slots) return FABLE_ERR_INVALID;
  *slots = fbt::n_slots(post->p);
  return FABLE_OK;
}

int fable_tile_slots(int64_t p, int64_t* slots) {
  if (!slots || p < 1) return FABLE_ERR_INVALID;
  *slots = fbt::n_slots(p);
  return FABLE_OK;
}

int fable_tile_count(int64_t p, int64_t t_begin, int64_t t_end, int64_t* tiles) {
  if (!tiles || p < 1) return FABLE_ERR_INVALID;
  const int64_t total = fbt::n_slots(p);
  if (t_end == 0 && t_begin == 0) t_end = total;
  if (t_begin < 0 || t_begin > t_end || t_end > total) return FABLE_ERR_INVALID;
  const int nT = static_cast<int>(fbt::n_tiles(p));
  *tiles = fbt::cidx(t_end, nT) - fbt::cidx(t_begin, nT);
  return FABLE_OK;
}

// slice g of `world` slices of the enumeration with (almost) equal numbers of VALID tiles: boundary g is the first slot
// t with cidx(t) >= g * valid / world
int fable_tile_partition(int64_t p, int rank, int world, int64_t* t_begin, int64_t* t_end) {
  if (!t_begin || !t_end || p < 1 || world < 1 || rank < 0 || rank >= world) return FABLE_ERR_INVALID;
  const int nT = static_cast<int>(fbt::n_tiles(p));
  const int64_t total = fbt::n_slots(p), valid = fbt::n_valid(p);
  auto boundary = [&](int g) -> int64_t {
    if (g <= 0) return 0;
    if (g >= world) return total;
    const int64_t want = valid / world * g + (valid % world) * g / world;
    int64_t lo = 0, hi = total;                        // smallest t with cidx(t) >= want
    while (lo < hi) {
      const int64_t mid = (lo + hi) / 2;
      if (fbt::cidx(mid, nT) >= want) hi = mid; else lo = mid + 1;
    }
    return lo;
  };
  *t_begin = boundary(rank); *t_end = boundary(rank + 1);
  return FABLE_OK;
}

int fable_tile_locate(int64_t p, int64_t I, int64_t J, int64_t* slot, int64_t* compact_index) {
  const int64_t nT = p >= 1 ? fbt::n_tiles(p) : 0;
  if (I < 0 || I > J || J >= nT) return FABLE_ERR_INVALID;
  const int64_t t = fbt::slot_of(static_cast<int>(I), static_cast<int>(J));
  if (slot) *slot = t;
  if (compact_index) *compact_index = fbt::cidx(t, static_cast<int>(nT));
  return FABLE_OK;
}

int fable_posterior_set_tile_range(fable_posterior* post, int64_t t_begin, int64_t t_end) {
  if (!post) return FABLE_ERR_INVALID;
  fable_ctx* ctx = post->ctx;
  FB_REQUIRE(ctx, t_begin >= 0 && t_begin <= t_end && t_end <= fbt::n_slots(post->p), "tile range out of bounds");
  if (t_begin == post->tile_begin && t_end == post->tile_end) return FABLE_OK;
  if (post->acc_sum.ptr) {              // the accumulators cover one slice: a new slice starts from zero
    FB_CUDA(ctx, cudaSetDevice(ctx->device));
    FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    post->acc_sum.release(); post->acc_sq.release(); post->acc_tiles = 0;
    post->part.release(); post->part_slots = -1; post->part_pending = 0;
  }
  post->tile_begin = t_begin; post->tile_end = t_end;
  return FABLE_OK;
}

int fable_posterior_tiles_to_dense(fable_posterior* post, const double* dev_tiles, int paired, int64_t col0, int64_t ncols,
                                   double* dev_out) {
  if (!post) return FABLE_ERR_INVALID;
  fable_ctx* ctx = post->ctx;
  FB_REQUIRE(ctx, dev_tiles && dev_out, "tiles_to_dense: NULL pointer");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  return fb_tiles_to_dense(ctx, dev_tiles, post->p, post->tile_begin, post->tile_end, paired, col0, ncols, dev_out);
}

int fable_posterior_samples(fable_posterior* post, int64_t m0, int64_t MC, int64_t ldm, double* LambdaSamples,
                            double* SigmaSqSamples, const double* gam_inj, const double* z_inj) {
  if (!post) return FABLE_ERR_INVALID;
  fable_ctx* ctx = post->ctx;
  FB_REQUIRE(ctx, MC >= 1 && ldm >= MC && m0 >= 0, "samples: need MC >= 1, ldm >= MC, m0 >= 0");
  FB_REQUIRE(ctx, (gam_inj == nullptr) == (z_inj == nullptr), "injected draws need both gam and z");
  if (!LambdaSamples && !SigmaSqSamples) return FABLE_OK;
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t p = post->p; const int k = post->k;
  DevBuf dgam, dz;
  if (gam_inj) {
    FB_TRY(upload(ctx, dgam, gam_inj, static_cast<size_t>(MC) * p));
    FB_TRY(upload(ctx, dz, z_inj, static_cast<size_t>(MC) * p * k));
  }
  const SampleArgs sa = sample_args(post, gam_inj ? dgam.d() : nullptr, gam_inj ? dz.d() : nullptr, m0);
  // chunk over rows j so that the device staging stays below ~4 GiB; a row chunk is a contiguous
  // block of columns of the MC x (k p) host matrix.
  const int64_t budget = (int64_t(1) << 32) / 8;
  int64_t nj_max = budget / (MC * (k + 1));
  if (nj_max < 1) nj_max = 1;
  if (nj_max > p) nj_max = p;
  DevBuf dL, dS;
  if (LambdaSamples) FB_CUDA(ctx, dL.alloc(static_cast<size_t>(MC) * nj_max * k * sizeof(double)));
  if (SigmaSqSamples) FB_CUDA(ctx, dS.alloc(static_cast<size_t>(MC) * nj_max * sizeof(double)));
  for (int64_t j0 = 0; j0 < p; j0 += nj_max) {
    const int64_t nj = p - j0 < nj_max ? p - j0 : nj_max;
    FB_TRY(fb_sample_ref(ctx, sa, m0, MC, j0, nj, MC, LambdaSamples ? dL.d() : nullptr,
                         SigmaSqSamples ? dS.d() : nullptr));
    if (LambdaSamples) {
      if (ldm == MC) FB_TRY(download(ctx, LambdaSamples + ldm * (j0 * k), dL.d(), static_cast<size_t>(MC) * nj * k));
      else FB_CUDA(ctx, cudaMemcpy2DAsync(LambdaSamples + ldm * (j0 * k), ldm * sizeof(double), dL.d(), MC * sizeof(double),
                                          MC * sizeof(double), nj * k, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (SigmaSqSamples) {
      if (ldm == MC) FB_TRY(download(ctx, SigmaSqSamples + ldm * j0, dS.d(), static_cast<size_t>(MC) * nj));
      else FB_CUDA(ctx, cudaMemcpy2DAsync(SigmaSqSamples + ldm * j0, ldm * sizeof(double), dS.d(), MC * sizeof(double),
                                          MC * sizeof(double), nj, cudaMemcpyDeviceToHost, ctx->stream));
    }
    FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->poll && ctx->poll(ctx->poll_user)) return ctx->fail(FABLE_ERR_INTERRUPT, "interrupted");
  }
  return FABLE_OK;
}

int fable_posterior_accum_dev(fable_posterior* post, double** dev_sum, double** dev_sumsq, int64_t* count) {
  if (!post) return FABLE_ERR_INVALID;
  FB_CUDA(post->ctx, cudaSetDevice(post->ctx->device));
  FB_TRY(ensure_accum(post));
  FB_TRY(fold_partials(post));
  if (dev_sum) *dev_sum = post->acc_sum.d();
  if (dev_sumsq) *dev_sumsq = post->acc_sq.d();
  if (count) *count = post->acc_tiles * fbt::TILE_DOUBLES;
  return FABLE_OK;
}

int fable_posterior_accum_reset(fable_posterior* post) {
  if (!post) return FABLE_ERR_INVALID;
  fable_ctx* ctx = post->ctx;
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  post->part_pending = 0;               // partial sums not folded yet belong to what is being discarded
  if (!post->acc_sum.ptr) return ensure_accum(post);
  const size_t bytes = static_cast<size_t>(post->acc_tiles) * fbt::TILE_DOUBLES * sizeof(double);
  FB_CUDA(ctx, cudaMemsetAsync(post->acc_sum.ptr, 0, bytes, ctx->stream));
  FB_CUDA(ctx, cudaMemsetAsync(post->acc_sq.ptr, 0, bytes, ctx->stream));
  return FABLE_OK;
}

// Dense symmetric p x p copies of the accumulators in caller (host) memory: the compact upper tiles are expanded on
// the device into column panels (both triangles; blocks outside the posterior's tile slice are zeros, so the results
// of the ranks of a tile partition add up) and each panel -- a contiguous run of columns of the dense matrix -- goes
// out as one copy.  No dense p x p buffer exists on the device.
int fable_posterior_accum_fetch(fable_posterior* post, double* sum, double* sumsq) {
  if (!post) return FABLE_ERR_INVALID;
  fable_ctx* ctx = post->ctx;
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  FB_TRY(ensure_accum(post));
  FB_TRY(fold_partials(post));
  const int64_t p = post->p;
  int64_t W = (int64_t(256) << 20) / (8 * p) / fbt::TILE * fbt::TILE;      // ~256 MB panels
  if (W < fbt::TILE) W = fbt::TILE;
  if (W > post->pp) W = post->pp;
  DevBuf panel[2];
  for (auto& b : panel) FB_CUDA(ctx, b.alloc(static_cast<size_t>(p) * W * sizeof(double)));
  int turn = 0;
  for (int which = 0; which < 2; ++which) {
    double* h = which ? sumsq : sum;
    if (!h) continue;
    const double* tiles = which ? post->acc_sq.d() : post->acc_sum.d();
    for (int64_t c0 = 0; c0 < p; c0 += W, turn ^= 1) {
      const int64_t nc = p - c0 < W ? p - c0 : W;
      // (two panels: a pageable destination makes download() return while its last chunk is still being copied out
      // of the pinned staging, never while the device buffer is still being read)
      FB_TRY(fb_tiles_to_dense(ctx, tiles, p, post->tile_begin, post->tile_end, 0, c0, nc, panel[turn].d()));
      FB_TRY(download(ctx, h + c0 * p, panel[turn].d(), static_cast<size_t>(p) * nc));
    }
  }
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FABLE_OK;
}

// The dense blocks of the super-blocks [sb_begin, sb_end) of the enumeration (16 x 16 tiles = up to 1024 x 1024 entries
// each, both the block and its mirror image) into the caller's dense p x p host matrices; everything else is left
// untouched.  A range of super-blocks is a contiguous range of accumulator tiles, so the tail of a sampling run can
// be processed chunk by chunk and every finished chunk can cross PCIe while the next one still computes (sampler
// below).  Runs on the context's CURRENT stream; returns when the host matrices hold the blocks.
int fable_posterior_accum_fetch_superblocks(fable_posterior* post, int64_t sb_begin, int64_t sb_end, double* sum, double* sumsq) {
  if (!post) return FABLE_ERR_INVALID;
  fable_ctx* ctx = post->ctx;
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  FB_TRY(ensure_accum(post));
  const int64_t p = post->p, slots = fbt::n_slots(p), per_sb = fbt::SB * fbt::SB;
  // (pending partial sums are the caller's business here: the sampler's tail folds chunk by chunk on the main stream and
  // calls this on the copy stream)
  const int64_t s_lo = post->tile_begin, s_hi = post->tile_end > 0 ? post->tile_end : slots;
  FB_REQUIRE(ctx, sb_begin >= 0 && sb_begin <= sb_end && sb_end * per_sb <= slots, "fetch_superblocks: range out of bounds");
  FB_REQUIRE(ctx, sb_begin * per_sb >= s_lo && sb_end * per_sb <= s_hi, "fetch_superblocks: range outside the posterior's tile slice");
  const int nT = static_cast<int>(fbt::n_tiles(p));
  const int64_t cbase = fbt::cidx(s_lo, nT);
  const int64_t span = static_cast<int64_t>(fbt::SB) * fbt::TILE, blk = span * span;
  // staging: NSET sets of (direct, mirror) blocks per accumulator, reused round-robin behind events; owned by the
  // posterior (a block taken from the device cache here would first wait for everything queued on the main stream)
  constexpr int NSET = 4;
  DevBuf& stage = post->sb_stage;
  if (!stage.ptr) {
    FB_CUDA(ctx, stage.alloc(static_cast<size_t>(NSET) * 4 * blk * sizeof(double)));
    for (auto& e : post->sb_ev) FB_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  cudaEvent_t* ev = post->sb_ev;
  int rc = FABLE_OK;
  double* zero_copy[2] = {nullptr, nullptr};
  if (getenv("FABLE_B200_TAIL_ZEROCOPY") && atoi(getenv("FABLE_B200_TAIL_ZEROCOPY")) != 0)
    for (int which = 0; which < 2; ++which) {
      double* h = which ? sumsq : sum;
      void* dp = nullptr;
      if (h && !is_pageable(h) && cudaHostGetDevicePointer(&dp, h, 0) == cudaSuccess) zero_copy[which] = static_cast<double*>(dp);
      else (void)cudaGetLastError();
    }
  for (int64_t sb = sb_begin; sb < sb_end && rc == FABLE_OK; ++sb) {
    int sI, sJ;
    fbt::tri_of(sb, sI, sJ);
    const int64_t r0 = sI * span, c0 = sJ * span;
    if (c0 >= p) continue;
    const int64_t nr = p - r0 < span ? p - r0 : span, nc = p - c0 < span ? p - c0 : span;
    const int set = static_cast<int>((sb - sb_begin) % NSET);
    if (sb - sb_begin >= NSET) FB_CUDA(ctx, cudaEventSynchronize(ev[set]));       // its previous copies have left
    for (int which = 0; which < 2 && rc == FABLE_OK; ++which) {
      double* h = which ? sumsq : sum;
      if (!h) continue;
      if (zero_copy[which]) {
        // pinned, device-mapped destination: the gather kernel writes the block and its mirror image straight into
        // the caller's matrix over PCIe (coalesced 512-byte runs) -- no staging, no 2-D copies of 8 KB rows
        rc = fb_superblock_to_dense(ctx, which ? post->acc_sq.d() : post->acc_sum.d(), cbase, p, sb, zero_copy[which] + r0 + p * c0,
                                    zero_copy[which] + c0 + p * r0, p, p);
        continue;
      }
      double* d_dir = stage.d() + (static_cast<size_t>(set) * 4 + which * 2) * blk;
      double* d_mir = d_dir + blk;
      rc = fb_superblock_to_dense(ctx, which ? post->acc_sq.d() : post->acc_sum.d(), cbase, p, sb, d_dir, d_mir);
      if (rc == FABLE_OK) rc = download_2d(ctx, h + r0 + p * c0, p, d_dir, nr, nc);
      if (rc == FABLE_OK && sI != sJ) rc = download_2d(ctx, h + c0 + p * r0, p, d_mir, nc, nr);
    }
    if (rc == FABLE_OK) FB_CUDA(ctx, cudaEventRecord(ev[set], ctx->stream));
  }
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return rc;
}

// the compact tiles as they are (acc_tiles x 4096 doubles each): the raw share of a rank of the tile partition
int fable_posterior_accum_fetch_tiles(fable_posterior* post, double* sum_tiles, double* sumsq_tiles) {
  if (!post) return FABLE_ERR_INVALID;
  fable_ctx* ctx = post->ctx;
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  FB_TRY(ensure_accum(post));
  FB_TRY(fold_partials(post));
  const size_t cnt = static_cast<size_t>(post->acc_tiles) * fbt::TILE_DOUBLES;
  FB_TRY(download(ctx, sum_tiles, post->acc_sum.d(), cnt));
  FB_TRY(download(ctx, sumsq_tiles, post->acc_sq.d(), cnt));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FABLE_OK;
}

// ---- a9 ------------------------------------------------------------------------------------------
int fable_sampler(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, double gamma0, double delta0sq,
                  int64_t MC, const double* U, const double* V, const double* d, int64_t r, int kEst,
                  double varInflation, const double* gam_inj, const double* z_inj, int64_t m0,
                  double* LambdaSamples, double* SigmaSqSamples, double* SigmaPostMean, double* SigmaPlugIn,
                  double* tausq, double* G, double* psi_sum, double* psi_sumsq) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_NO_SHARD(ctx, "CPPFABLESampler");
  FB_REQUIRE(ctx, MC >= 1, "MC must be >= 1");
  FB_REQUIRE(ctx, (gam_inj == nullptr) == (z_inj == nullptr), "injected draws need both gam and z");
  FB_REQUIRE(ctx, (psi_sum == nullptr) == (psi_sumsq == nullptr), "psi_sum and psi_sumsq go together");
  fable_posterior* post = nullptr;
  FB_TRY(fable_posterior_create(ctx, Y, n, p, gamma0, delta0sq, U, V, d, r, kEst, varInflation, &post));
  struct Guard { fable_posterior* p; ~Guard() { fable_posterior_destroy(p); } } guard{post};
  return sampler_outputs(post, MC, gam_inj, z_inj, m0, LambdaSamples, SigmaSqSamples, SigmaPostMean, SigmaPlugIn, tausq, G,
                         psi_sum, psi_sumsq);
}

}  // extern "C"

// The draw loop of CPPFABLESampler (cpp:599-618) for draws m0 .. m0 + MC - 1 from a resident row posterior: the
// reference-layout samples go to the host (rows 0 .. MC-1 of arrays with leading dimension ldm), and with psi_pass the
// covariance pass adds the same draws into the posterior's device-resident sum / sumsq accumulators (and materialises
// them into the context's ring when one is set) -- nothing p x p crosses PCIe here, so the ranks of a draw partition can
// sum their accumulators on the device before anyone fetches them.
//
// fetch_sum / fetch_sq (single-device callers that want the dense sums on the host at the end): the last ~256 draws are
// then processed CHUNK-major instead of batch-major -- all of their batches for the first group of super-blocks, then for
// the next ... -- so that the accumulator tiles of a finished group are final while the following groups still compute,
// and their dense blocks cross PCIe (copy stream) behind that work instead of after it.
static int sample_range(fable_posterior* post, int64_t MC, int64_t m0, int64_t ldm, const double* gam_inj, const double* z_inj,
                        double* LambdaSamples, double* SigmaSqSamples, bool psi_pass, double* fetch_sum = nullptr,
                        double* fetch_sq = nullptr) {
  fable_ctx* ctx = post->ctx;
  const int64_t p = post->p;
  const int kEst = post->k;
  // The Psi pass (draw loop on the main stream) is enqueued first; the reference-layout samples are
  // then generated and copied out on the copy stream, so their D2H traffic (8 p (k+1) bytes per draw)
  // overlaps the covariance kernels.  Both read the same immutable row posterior.
  DevBuf dgam, dz;
  // pageable sample matrices (what R hands over: MC x kp and MC x p doubles, 1.8 GB at C3 with MC = 1000) and dense sums are
  // touched by background threads from here on; the copies below find most of their pages mapped
  HostPrefault touch_l, touch_s, touch_sum, touch_sq;
  if (ldm == MC) {                    // (a row range of a larger matrix is shared with other callers: left alone)
    if (LambdaSamples) touch_l.start(ctx, LambdaSamples, static_cast<size_t>(MC) * p * kEst * sizeof(double));
    if (SigmaSqSamples) touch_s.start(ctx, SigmaSqSamples, static_cast<size_t>(MC) * p * sizeof(double));
  }
  if (fetch_sum) touch_sum.start(ctx, fetch_sum, static_cast<size_t>(p) * p * sizeof(double));
  if (fetch_sq) touch_sq.start(ctx, fetch_sq, static_cast<size_t>(p) * p * sizeof(double));
  struct TailChunk { cudaEvent_t ev; int64_t sb0, sb1; };
  struct TailEvents : std::vector<TailChunk> {
    ~TailEvents() { for (auto& t : *this) cudaEventDestroy(t.ev); }
  } tail_events;
  if (psi_pass) {
    if (gam_inj) {
      FB_TRY(upload(ctx, dgam, gam_inj, static_cast<size_t>(MC) * p));
      FB_TRY(upload(ctx, dz, z_inj, static_cast<size_t>(MC) * p * kEst));
    }
    FB_TRY(ensure_accum(post));
    if (ctx->ring_slots > 0) {
      const size_t need = static_cast<size_t>(ctx->ring_slots) * p * p * sizeof(double);
      if (need != ctx->ring_bytes) {
        FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->ring) { cudaFree(ctx->ring); ctx->ring = nullptr; ctx->ring_bytes = 0; }
        if (cudaMalloc(reinterpret_cast<void**>(&ctx->ring), need) != cudaSuccess) {
          (void)cudaGetLastError();
          fb_mem_trim(ctx->device);                  // idle cached blocks may be what is in the way
          FB_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->ring), need));
        }
        ctx->ring_bytes = need;
      }
    }
    const int B = ctx->ring_slots > 0 ? (ctx->ring_slots < 64 ? ctx->ring_slots : 64) : 32;
    const int64_t nbatch = ceil_div(MC, B);
    // tail: the batches of the last ~256 draws, run chunk-major when the caller fetches the sums (FABLE_B200_TAIL_DRAWS=0: off)
    int64_t tail = 0;
    // (pinned destinations only: into pageable memory the block-wise scatter by host threads is slower than the plain
    // panel fetch -- measured 1543 vs 1612 draws/s end to end at c3)
    if (fetch_sum && post->tile_end == 0 && !is_pageable(fetch_sum) && !is_pageable(fetch_sq)) {
      const char* e = getenv("FABLE_B200_TAIL_DRAWS");
      const int64_t td = e ? atoll(e) : 256;
      tail = td <= 0 ? 0 : ceil_div(td, B);
      if (ctx->ring_slots > 0) {                       // materialising run: the tail's partial sums need one slot per batch
        FB_TRY(ensure_partials(post));
        if (post->part_slots > 0 && tail > post->part_slots) tail = post->part_slots;
      }
      if (tail > nbatch) tail = nbatch;
      // the fetch staging is allocated before anything is queued (the call drains the stream)
      if (tail > 0 && !post->sb_stage.ptr) FB_TRY(fable_posterior_accum_fetch_superblocks(post, 0, 0, fetch_sum, fetch_sq));
    }
    auto batch = [&](int64_t bi, int64_t lb, int64_t le, int partial_slot) -> int {
      const int64_t b0 = bi * B;
      const int nb = static_cast<int>(MC - b0 < B ? MC - b0 : B);
      return draw_batch_impl(post, m0 + b0, nb, ctx->ring_slots > 0 ? ctx->ring : nullptr, ctx->ring_slots, false, 1,
                             gam_inj ? dgam.d() + b0 * p : nullptr, gam_inj ? dz.d() + b0 * p * kEst : nullptr, lb, le, partial_slot);
    };
    for (int64_t bi = 0; bi < nbatch - tail; ++bi) {
      FB_TRY(batch(bi, 0, 0, -1));
      if (ctx->poll && ctx->poll(ctx->poll_user)) {
        cudaStreamSynchronize(ctx->stream);
        return ctx->fail(FABLE_ERR_INTERRUPT, "interrupted");
      }
    }
    if (tail > 0) {
      // the tail batches of a chunk store their sums into partial slots 0 .. tail-1 (when the posterior uses partial
      // sums: the same association as the batch-major path, so the results do not depend on the tail) and are folded
      // into the chunk's accumulator tiles before the chunk's event
      FB_TRY(fold_partials(post));
      const bool parts = post->part_slots >= tail && ctx->ring_slots > 0;
      const int nT = static_cast<int>(fbt::n_tiles(p));
      const int64_t nsb = fbt::n_slots(p) / (fbt::SB * fbt::SB);
      const int64_t nchunk = nsb < 12 ? nsb : 12;
      for (int64_t c = 0; c < nchunk; ++c) {
        const int64_t sb0 = nsb * c / nchunk, sb1 = nsb * (c + 1) / nchunk;
        const int64_t t0 = sb0 * fbt::SB * fbt::SB, t1 = sb1 * fbt::SB * fbt::SB;
        for (int64_t bi = nbatch - tail; bi < nbatch; ++bi) FB_TRY(batch(bi, t0, t1, parts ? static_cast<int>(bi - (nbatch - tail)) : -1));
        if (parts) {
          post->part_pending = static_cast<int>(tail);
          FB_TRY(fold_range(post, fbt::cidx(t0, nT), fbt::cidx(t1, nT), true));
        }
        cudaEvent_t ev;
        FB_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        tail_events.push_back({ev, sb0, sb1});
        FB_CUDA(ctx, cudaEventRecord(ev, ctx->stream));
      }
    }
  }
  {
    struct StreamSwap {      // run the sample kernels + copies on the copy stream
      fable_ctx* c; cudaStream_t main;
      explicit StreamSwap(fable_ctx* cx) : c(cx), main(cx->stream) { c->stream = c->copy_stream; }
      ~StreamSwap() { c->stream = main; }
    } swap(ctx);
    int rc = fable_posterior_samples(post, m0, MC, ldm, LambdaSamples, SigmaSqSamples, gam_inj, z_inj);
    // finished groups of super-blocks leave on the copy stream while the main stream works on the following ones
    for (size_t c = 0; c < tail_events.size() && rc == FABLE_OK; ++c) {
      const cudaError_t e = cudaEventSynchronize(tail_events[c].ev);
      if (e != cudaSuccess) { rc = ctx->fail(FABLE_ERR_CUDA, std::string("tail event: ") + cudaGetErrorString(e)); break; }
      rc = fable_posterior_accum_fetch_superblocks(post, tail_events[c].sb0, tail_events[c].sb1, fetch_sum, fetch_sq);
    }
    if (rc != FABLE_OK) { cudaStreamSynchronize(swap.main); return rc; }
  }
  if (psi_pass) FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // dgam / dz go out of scope
  if (psi_pass && fetch_sum && tail_events.empty()) FB_TRY(fable_posterior_accum_fetch(post, fetch_sum, fetch_sq));
  return FABLE_OK;
}

// everything CPPFABLESampler returns (cpp:599-627) from a resident row posterior
static int sampler_outputs(fable_posterior* post, int64_t MC, const double* gam_inj, const double* z_inj, int64_t m0,
                           double* LambdaSamples, double* SigmaSqSamples, double* SigmaPostMean, double* SigmaPlugIn,
                           double* tausq, double* G, double* psi_sum, double* psi_sumsq) {
  fable_ctx* ctx = post->ctx;
  HostPrefault touch_g;               // G (cpp:627) leaves last: its pages are touched while the draws are made and copied
  if (G) touch_g.start(ctx, G, static_cast<size_t>(post->p) * post->p * sizeof(double));
  FB_TRY(fable_posterior_fixed_outputs(post, SigmaPostMean, SigmaPlugIn, tausq, nullptr));
  if (psi_sum) FB_TRY(fable_posterior_accum_reset(post));
  FB_TRY(sample_range(post, MC, m0, MC, gam_inj, z_inj, LambdaSamples, SigmaSqSamples, psi_sum != nullptr, psi_sum, psi_sumsq));
  if (G) FB_TRY(rankk_cov_to_host(ctx, post->rp.G0.d(), post->p, nullptr, post->p, post->k, G));   // cpp:627
  return FABLE_OK;
}

// the elements of CPPFABLESampler's list that do not depend on the draws (cpp:622-627)
extern "C" int fable_posterior_fixed_outputs(fable_posterior* post, double* SigmaPostMean, double* SigmaPlugIn, double* tausq,
                                             double* G) {
  if (!post) return FABLE_ERR_INVALID;
  fable_ctx* ctx = post->ctx;
  const int64_t p = post->p;
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  FB_TRY(download(ctx, SigmaPostMean, post->rp.sigmahat.d(), p));              // cpp:622
  FB_TRY(download(ctx, SigmaPlugIn, post->rp.sig.d(), p));                     // cpp:623
  if (tausq) *tausq = post->rp.tausq;                                          // cpp:626
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (G) FB_TRY(rankk_cov_to_host(ctx, post->rp.G0.d(), p, nullptr, p, post->k, G));   // cpp:627
  return FABLE_OK;
}

extern "C" int fable_posterior_sample_range(fable_posterior* post, int64_t MC, int64_t m0, int64_t ldm, const double* gam_inj,
                                            const double* z_inj, double* LambdaSamples, double* SigmaSqSamples, int psi_pass) {
  if (!post) return FABLE_ERR_INVALID;
  fable_ctx* ctx = post->ctx;
  FB_REQUIRE(ctx, MC >= 1 && m0 >= 0, "sample_range: need MC >= 1 and m0 >= 0");
  FB_REQUIRE(ctx, (gam_inj == nullptr) == (z_inj == nullptr), "injected draws need both gam and z");
  FB_REQUIRE(ctx, (!LambdaSamples && !SigmaSqSamples) || ldm >= MC, "sample_range: ldm must be >= MC");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  return sample_range(post, MC, m0, ldm < MC ? MC : ldm, gam_inj, z_inj, LambdaSamples, SigmaSqSamples, psi_pass != 0);
}

// ---- a11: the two R wrappers as single calls (Y crosses PCIe once, nothing round-trips) -------------
namespace {
struct Pipeline {
  Inputs in;          // Y plus the FULL thin SVD, resident
  std::vector<double> d;
  int kMax = 0;
};

// R/FABLE_wrapper.R:20-27 == :87-94: as.matrix, svd(Y), kMax rule
// Stage marks of the wrapper pipelines (only when fable_ctx_stage_timing is on: every mark drains the stream).
// Slots: 0 upload of Y, 1 Gram + eigensolver + back-multiplication, 2 U / d / V to the host + kMax rule, 3 rank search,
// 4 row posterior (+ hyper-parameters, varInflation), 5 p x p results to the host.
struct StageClock {
  fable_ctx* ctx;
  std::chrono::steady_clock::time_point t0;
  explicit StageClock(fable_ctx* c) : ctx(c) {
    if (ctx->stage_timing) { cudaStreamSynchronize(ctx->stream); t0 = std::chrono::steady_clock::now(); }
  }
  void mark(int slot) {
    if (!ctx->stage_timing) return;
    cudaStreamSynchronize(ctx->stream);
    const auto t1 = std::chrono::steady_clock::now();
    ctx->stage_ms[slot] += std::chrono::duration<double, std::milli>(t1 - t0).count();
    t0 = t1;
  }
};

int pipeline_svd(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, double maxProp, Pipeline& pl, double* U,
                 double* dOut, double* V) {
  FB_REQUIRE(ctx, Y && n >= 1 && p >= 1, "Y must be a non-empty matrix");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ctx->stage_timing) for (double& x : ctx->stage_ms) x = 0.0;
  StageClock clk(ctx);
  HostPrefault touch_v;               // svdY$v (p x r: 160 MB at C3) is known now and written after the eigensolver
  if (V) touch_v.start(ctx, V, static_cast<size_t>(p) * (n < p ? n : p) * sizeof(double), size_t(64) << 20);
  const int64_t r = n < p ? n : p;
  Inputs& in = pl.in;
  in.n = n; in.p = p; in.r = r; in.kcols = static_cast<int>(r);
  FB_TRY(upload(ctx, in.Y, Y, static_cast<size_t>(n) * p));
  clk.mark(0);
  FB_CUDA(ctx, in.U.alloc(static_cast<size_t>(n) * r * sizeof(double)));
  FB_CUDA(ctx, in.V.alloc(static_cast<size_t>(p) * r * sizeof(double)));
  FB_CUDA(ctx, in.d.alloc(static_cast<size_t>(r) * sizeof(double)));
  FB_TRY(fb_svd(ctx, in.Y.d(), n, p, in.U.d(), in.d.d(), in.V.d()));
  in.own_svd = n <= p;        // n <= p: V = Y'U / d column by column and U is the eigensolver's orthogonal factor
  clk.mark(1);
  pl.d.resize(r);
  FB_TRY(download(ctx, pl.d.data(), in.d.d(), r));
  FB_TRY(download(ctx, U, in.U.d(), static_cast<size_t>(n) * r));
  FB_TRY(download(ctx, V, in.V.d(), static_cast<size_t>(p) * r));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (dOut) std::memcpy(dOut, pl.d.data(), r * sizeof(double));
  if (fable_kmax(pl.d.data(), r, maxProp, &pl.kMax) != FABLE_OK)
    return ctx->fail(FABLE_ERR_INVALID, "kMax rule: no index reaches maxProp");       // R: min(integer(0)) -> Inf
  clk.mark(2);
  return FABLE_OK;
}
}  // namespace

extern "C" {

int fable_wrapper_posterior_mean(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, double gamma0, double delta0sq,
                                 double maxProp, double* Cov, int* kEst, int* kMaxOut, double* U, double* d, double* V) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_NO_SHARD(ctx, "FABLEPosteriorMean");
  FB_REQUIRE(ctx, kEst, "posterior_mean: NULL kEst");
  HostPrefault touch;
  if (Cov && p >= 1) touch.start(ctx, Cov, static_cast<size_t>(p) * p * sizeof(double));
  Pipeline pl;
  FB_TRY(pipeline_svd(ctx, Y, n, p, maxProp, pl, U, d, V));                      // wrapper:87-94
  if (kMaxOut) *kMaxOut = pl.kMax;
  int k = 1;
  StageClock clk(ctx);
  FB_TRY(rank_dev(ctx, pl.in, pl.kMax, &k));                                     // cpp:457
  clk.mark(3);
  RowPosterior rp;
  FB_TRY(row_posterior(ctx, pl.in, k, gamma0, delta0sq, /*shrunk=*/1, rp));
  clk.mark(4);
  *kEst = k;
  touch.before_result();
  if (Cov) FB_TRY(rankk_cov_to_host(ctx, rp.G0.d(), p, rp.sigmahat.d(), p, k, Cov));   // cpp:491, 509-512
  clk.mark(5);
  return FABLE_OK;
}

// wrapper:20-44 on the device: SVD, kMax, rank search, hyper-parameters, varInflation; leaves the (shrunk) row
// posterior of wrapper:46-54 resident in *out
int fable_wrapper_sampler_begin(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, double gamma0, double delta0sq,
                                double maxProp, double* U, double* d, double* V, int* kMaxOut, int* kEstOut,
                                double* hypSigma, double* hypG0, double* hypTau, double* hypG, double* varInflationOut,
                                fable_posterior** out) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_NO_SHARD(ctx, "FABLEPosteriorSampler");
  FB_REQUIRE(ctx, out, "sampler_begin: NULL out");
  *out = nullptr;
  HostPrefault touch;
  if (hypG && p >= 1) touch.start(ctx, hypG, static_cast<size_t>(p) * p * sizeof(double));
  Pipeline pl;
  FB_TRY(pipeline_svd(ctx, Y, n, p, maxProp, pl, U, d, V));                      // wrapper:20-27
  if (kMaxOut) *kMaxOut = pl.kMax;
  int k = 1;
  StageClock clk(ctx);
  FB_TRY(rank_dev(ctx, pl.in, pl.kMax, &k));                                     // wrapper:29-33
  clk.mark(3);
  if (kEstOut) *kEstOut = k;
  double vi = 0.0;
  RowPosterior hyp;                                                              // wrapper:35-39 (unshrunk G0, cpp:390)
  {
    FB_TRY(row_posterior(ctx, pl.in, k, 1.0, 1.0, /*shrunk=*/0, hyp));
    FB_TRY(download(ctx, hypSigma, hyp.sig.d(), p));
    FB_TRY(download(ctx, hypG0, hyp.G0.d(), static_cast<size_t>(p) * k));
    if (hypTau) *hypTau = hyp.tausq;
    double su, sd;                                                               // wrapper:41-44, B never stored
    FB_TRY(fb_var_inflation(ctx, hyp.sig.d(), hyp.G0.d(), p, p, k, &su, &sd));
    FB_TRY(var_inflation_from_sums(su, sd, p, /*variant=*/0, &vi));
    touch.before_result();
    if (hypG) FB_TRY(rankk_cov_to_host(ctx, hyp.G0.d(), p, nullptr, p, k, hypG));     // cpp:391
    FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  if (varInflationOut) *varInflationOut = vi;
  FB_TRY(posterior_from_inputs(ctx, pl.in, gamma0, delta0sq, k, vi, out));       // wrapper:46-54, cpp:553-597
  clk.mark(4);
  (*out)->hyp_G0 = std::move(hyp.G0);      // R sizes FABLEHyperParameters$G0 (p x kEst) once it knows kEst: fetched later
  return FABLE_OK;
}

int fable_posterior_hyper_G0(fable_posterior* post, double* hypG0) {
  if (!post) return FABLE_ERR_INVALID;
  fable_ctx* ctx = post->ctx;
  FB_REQUIRE(ctx, hypG0, "hyper_G0: NULL output");
  FB_REQUIRE(ctx, post->hyp_G0.ptr, "hyper_G0: this posterior was not created by fable_wrapper_sampler_begin");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  FB_TRY(download(ctx, hypG0, post->hyp_G0.d(), static_cast<size_t>(post->p) * post->k));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FABLE_OK;
}

// ---- f3: CCFABLE_DirectSampler(Y, gamma0, delta0sq, MC) (R/FABLE_code.R:212-298) as a callable -----------------------
// begin = R:214-263: svd(Y), kMax at 0.95 (:219), the prototype's own RankEstimator (:221), the row posterior (shrunk G0,
// :236-257 = cpp:564-592) with a_n = 1 / sqrt(n + 1/tau^2) (:263: no variance inflation) and the entry-wise correction
// B = cov_correct_matrix(sigma^2, G) (:251) switched on.  draws = the MC loop (:267-294) into the (MC, p, p) array.
int fable_direct_sampler_begin(fable_ctx* ctx, const double* Y, int64_t n, int64_t p, double gamma0, double delta0sq, int* kMaxOut,
                               int* kEstOut, fable_posterior** out) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_NO_SHARD(ctx, "CCFABLE_DirectSampler");
  FB_REQUIRE(ctx, out, "direct_sampler_begin: NULL out");
  *out = nullptr;
  Pipeline pl;
  FB_TRY(pipeline_svd(ctx, Y, n, p, /*maxProp=*/0.95, pl, nullptr, nullptr, nullptr));          // R:216-219
  if (kMaxOut) *kMaxOut = pl.kMax;
  int k = 1;
  FB_TRY(rank_dev(ctx, pl.in, pl.kMax, &k, /*variant=*/1));                                       // R:221
  if (kEstOut) *kEstOut = k;
  FB_TRY(posterior_from_inputs(ctx, pl.in, gamma0, delta0sq, k, /*varInflation=*/1.0, out));     // R:236-263
  const int rc = fable_posterior_set_entrywise_correction(*out, 1);                              // R:251, 285
  if (rc != FABLE_OK) { fable_posterior_destroy(*out); *out = nullptr; }
  return rc;
}

int fable_direct_sampler_draws(fable_posterior* post, int64_t MC, int64_t m0, const double* gam_inj, const double* z_inj,
                               double* CovSampleStor) {
  if (!post) return FABLE_ERR_INVALID;
  fable_ctx* ctx = post->ctx;
  FB_REQUIRE(ctx, MC >= 1 && m0 >= 0 && CovSampleStor, "direct_sampler_draws: need MC >= 1, m0 >= 0 and an output array");
  FB_REQUIRE(ctx, (gam_inj == nullptr) == (z_inj == nullptr), "injected draws need both gam and z");
  const int64_t p = post->p, p2 = p * p;
  const int k = post->k;
  FB_REQUIRE(ctx, static_cast<double>(MC) * static_cast<double>(p2) * 8.0 <= 32e9,
             "CCFABLE_DirectSampler stores MC x p x p doubles (R/FABLE_code.R:262): above 32 GB use FABLEPosteriorSampler + "
             "post-processing instead, as the reference's own comments advise (:209-211)");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  DevBuf dgam, dz, ring, outb;
  if (gam_inj) {
    FB_TRY(upload(ctx, dgam, gam_inj, static_cast<size_t>(MC) * p));
    FB_TRY(upload(ctx, dz, z_inj, static_cast<size_t>(MC) * p * k));
  }
  int64_t B = (int64_t(2) << 30) / (8 * p2);                     // ~2 GB of dense draws per batch
  B = B < 1 ? 1 : (B > 64 ? 64 : B);
  if (B > MC) B = MC;
  FB_CUDA(ctx, ring.alloc(static_cast<size_t>(B) * p2 * sizeof(double)));
  FB_CUDA(ctx, outb.alloc(static_cast<size_t>(MC) * p2 * sizeof(double)));
  for (int64_t b0 = 0; b0 < MC; b0 += B) {
    const int nb = static_cast<int>(MC - b0 < B ? MC - b0 : B);
    FB_TRY(fable_posterior_draw_batch(post, m0 + b0, nb, ring.d(), static_cast<int>(B), 0, gam_inj ? dgam.d() + b0 * p : nullptr,
                                      gam_inj ? dz.d() + b0 * p * k : nullptr));
    FB_TRY(fb_slots_to_marray(ctx, ring.d(), nb, p2, MC, b0, outb.d()));                      // CovSampleStor[m, , ] = PsiSample (R:290)
    if (ctx->poll && ctx->poll(ctx->poll_user)) {
      cudaStreamSynchronize(ctx->stream);
      return ctx->fail(FABLE_ERR_INTERRUPT, "interrupted");
    }
  }
  FB_TRY(download(ctx, CovSampleStor, outb.d(), static_cast<size_t>(MC) * p2));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FABLE_OK;
}

int fable_posterior_sampler_outputs(fable_posterior* post, int64_t MC, int64_t m0, const double* gam_inj,
                                    const double* z_inj, double* LambdaSamples, double* SigmaSqSamples,
                                    double* SigmaPostMean, double* SigmaPlugIn, double* tausq, double* G, double* psi_sum,
                                    double* psi_sumsq) {
  if (!post) return FABLE_ERR_INVALID;
  fable_ctx* ctx = post->ctx;
  FB_REQUIRE(ctx, MC >= 1, "MC must be >= 1");
  FB_REQUIRE(ctx, (gam_inj == nullptr) == (z_inj == nullptr), "injected draws need both gam and z");
  FB_REQUIRE(ctx, (psi_sum == nullptr) == (psi_sumsq == nullptr), "psi_sum and psi_sumsq go together");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  return sampler_outputs(post, MC, gam_inj, z_inj, m0, LambdaSamples, SigmaSqSamples, SigmaPostMean, SigmaPlugIn, tausq, G,
                         psi_sum, psi_sumsq);
}

int fable_posterior_rank(fable_posterior* post, int* k) {
  if (!post || !k) return FABLE_ERR_INVALID;
  *k = post->k;
  return FABLE_OK;
}

// ---- f1: post-processing of the stored draws -------------------------------------------------------
int fable_post_processing(fable_ctx* ctx, const double* LambdaSamples, const double* SigmaSqSamples, int64_t MC,
                          int64_t p, int k, double alpha, const int64_t* selected, int64_t S, double* PostMean,
                          double* Lower, double* Upper) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_REQUIRE(ctx, LambdaSamples && SigmaSqSamples && PostMean && Lower && Upper, "post_processing: NULL pointer");
  FB_REQUIRE(ctx, MC >= 1 && p >= 1 && k >= 1, "post_processing: need MC, p, k >= 1");
  if (!selected) S = p;
  FB_REQUIRE(ctx, S >= 1, "post_processing: empty selection");
  std::vector<int64_t> idx(S);
  for (int64_t i = 0; i < S; ++i) {
    idx[i] = selected ? selected[i] - 1 : i;                          // cpp:721: 1-based R indices
    FB_REQUIRE(ctx, idx[i] >= 0 && idx[i] < p, "post_processing: SelectedIndices out of range");   // arma bounds error
  }
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  DevBuf dL, dS, dI, dM, dLo, dUp;
  FB_CUDA(ctx, dL.alloc(static_cast<size_t>(S) * k * MC * sizeof(double)));
  FB_CUDA(ctx, dS.alloc(static_cast<size_t>(S) * MC * sizeof(double)));
  FB_CUDA(ctx, dI.alloc(static_cast<size_t>(S) * sizeof(int64_t)));
  for (auto* b : {&dM, &dLo, &dUp}) FB_CUDA(ctx, b->alloc(static_cast<size_t>(S) * S * sizeof(double)));
  if (!selected) {
    FB_TRY(upload_bytes(ctx, dL.ptr, LambdaSamples, static_cast<size_t>(p) * k * MC * sizeof(double)));
    FB_TRY(upload_bytes(ctx, dS.ptr, SigmaSqSamples, static_cast<size_t>(p) * MC * sizeof(double)));
  } else {
    for (int64_t i = 0; i < S; ++i) {   // the k factor columns of one variable are contiguous in the reference layout
      FB_CUDA(ctx, cudaMemcpyAsync(dL.d() + i * k * MC, LambdaSamples + idx[i] * k * MC, static_cast<size_t>(k) * MC * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      FB_CUDA(ctx, cudaMemcpyAsync(dS.d() + i * MC, SigmaSqSamples + idx[i] * MC, static_cast<size_t>(MC) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
  }
  FB_CUDA(ctx, cudaMemcpyAsync(dI.ptr, idx.data(), static_cast<size_t>(S) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
  // tile-resident selection (post_tiles.cu) wherever its shape limits allow; the per-pair kernels of post.cu otherwise, and
  // for the (theoretical) entries the selection reports as undecidable
  const bool tiles_on = !(getenv("FABLE_B200_POST_TILES") && atoi(getenv("FABLE_B200_POST_TILES")) == 0);   // 0: per-pair kernels (A/B, tests)
  int unresolved = -1;
  if (tiles_on && fb_post_tiles_supported(MC, k)) {
    const int64_t pps = round_up(S, 64);
    const int KP = static_cast<int>(round_up(k, 4));
    DevBuf Lw, s2w;
    FB_CUDA(ctx, Lw.alloc(static_cast<size_t>(MC) * KP * pps * sizeof(double)));
    FB_CUDA(ctx, s2w.alloc(static_cast<size_t>(MC) * pps * sizeof(double)));
    FB_TRY(fb_ref_to_work(ctx, dL.d(), dS.d(), S, k, MC, pps, KP, Lw.d(), s2w.d()));
    FB_TRY(fb_post_tiles(ctx, Lw.d(), s2w.d(), static_cast<const int64_t*>(dI.ptr), MC, KP, S, pps, alpha, dM.d(), dLo.d(), dUp.d(),
                         &unresolved));
  }
  if (unresolved != 0)
    FB_TRY(fb_post_processing(ctx, dL.d(), dS.d(), static_cast<const int64_t*>(dI.ptr), MC, k, S, alpha, dM.d(), dLo.d(), dUp.d()));
  FB_TRY(download(ctx, PostMean, dM.d(), static_cast<size_t>(S) * S));
  FB_TRY(download(ctx, Lower, dLo.d(), static_cast<size_t>(S) * S));
  FB_TRY(download(ctx, Upper, dUp.d(), static_cast<size_t>(S) * S));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FABLE_OK;
}

// f1 without the 16 GB round trip: the draws never leave the device.  Draws m0 .. m0 + MC - 1 of the resident posterior are
// generated in the working layout (native stream, or injected draws), and the tile-resident selection forms every entry of
// Psi_m on the fly -- the consumer of the block-wise covariance formation that the reference reaches by storing MC x k p
// samples and looping over the pairs (cpp:631-713).
int fable_posterior_post_processing(fable_posterior* post, int64_t MC, int64_t m0, double alpha, const int64_t* selected, int64_t S,
                                    const double* gam_inj, const double* z_inj, double* PostMean, double* Lower, double* Upper) {
  if (!post) return FABLE_ERR_INVALID;
  fable_ctx* ctx = post->ctx;
  FB_REQUIRE(ctx, PostMean && Lower && Upper, "post_processing: NULL output");
  FB_REQUIRE(ctx, MC >= 1 && m0 >= 0, "post_processing: need MC >= 1 and m0 >= 0");
  FB_REQUIRE(ctx, (gam_inj == nullptr) == (z_inj == nullptr), "injected draws need both gam and z");
  FB_REQUIRE(ctx, fb_post_tiles_supported(MC, post->k), "post_processing on resident draws: needs MC <= 65535 and k <= 32 "
                                                        "(use fable_post_processing on stored samples beyond that)");
  const int64_t p = post->p, pp = post->pp;
  const int k = post->k, KP = post->KP;
  if (!selected) S = p;
  FB_REQUIRE(ctx, S >= 1, "post_processing: empty selection");
  std::vector<int64_t> idx(S);
  for (int64_t i = 0; i < S; ++i) {
    idx[i] = selected ? selected[i] - 1 : i;                          // 1-based R indices (cpp:721)
    FB_REQUIRE(ctx, idx[i] >= 0 && idx[i] < p, "post_processing: SelectedIndices out of range");
  }
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  DevBuf dgam, dz, Lw, s2w, Lsel, s2sel, dI, dM, dLo, dUp;
  if (gam_inj) {
    FB_TRY(upload(ctx, dgam, gam_inj, static_cast<size_t>(MC) * p));
    FB_TRY(upload(ctx, dz, z_inj, static_cast<size_t>(MC) * p * k));
  }
  FB_CUDA(ctx, Lw.alloc(static_cast<size_t>(MC) * KP * pp * sizeof(double)));
  FB_CUDA(ctx, s2w.alloc(static_cast<size_t>(MC) * pp * sizeof(double)));
  const SampleArgs sa = sample_args(post, gam_inj ? dgam.d() : nullptr, gam_inj ? dz.d() : nullptr, m0);
  for (int64_t b0 = 0; b0 < MC; b0 += 32768) {
    const int nb = static_cast<int>(MC - b0 < 32768 ? MC - b0 : 32768);
    FB_TRY(fb_sample_work(ctx, sa, m0 + b0, nb, pp, KP, Lw.d() + b0 * KP * pp, s2w.d() + b0 * pp));
  }
  FB_CUDA(ctx, dI.alloc(static_cast<size_t>(S) * sizeof(int64_t)));
  FB_CUDA(ctx, cudaMemcpyAsync(dI.ptr, idx.data(), static_cast<size_t>(S) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
  const double* useL = Lw.d(); const double* useS = s2w.d();
  int64_t pps = pp;
  if (selected) {
    pps = round_up(S, 64);
    FB_CUDA(ctx, Lsel.alloc(static_cast<size_t>(MC) * KP * pps * sizeof(double)));
    FB_CUDA(ctx, s2sel.alloc(static_cast<size_t>(MC) * pps * sizeof(double)));
    FB_TRY(fb_gather_work_rows(ctx, Lw.d(), s2w.d(), pp, static_cast<const int64_t*>(dI.ptr), S, pps, KP, MC, Lsel.d(), s2sel.d()));
    useL = Lsel.d(); useS = s2sel.d();
  }
  for (auto* b : {&dM, &dLo, &dUp}) FB_CUDA(ctx, b->alloc(static_cast<size_t>(S) * S * sizeof(double)));
  int unresolved = 0;
  FB_TRY(fb_post_tiles(ctx, useL, useS, static_cast<const int64_t*>(dI.ptr), MC, KP, S, pps, alpha, dM.d(), dLo.d(), dUp.d(), &unresolved));
  if (unresolved != 0)
    return ctx->fail(FABLE_ERR_NUMERIC, "post_processing: " + std::to_string(unresolved) + " entries have more than 16 distinct draws "
                                        "closer than 2^-52 of their range; use fable_post_processing on stored samples");
  FB_TRY(download(ctx, PostMean, dM.d(), static_cast<size_t>(S) * S));
  FB_TRY(download(ctx, Lower, dLo.d(), static_cast<size_t>(S) * S));
  FB_TRY(download(ctx, Upper, dUp.d(), static_cast<size_t>(S) * S));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // idx and the device buffers go out of scope
  return FABLE_OK;
}

// ---- device-pointer primitives --------------------------------------------------------------------
int fable_dev_rankk_cov(fable_ctx* ctx, const double* dA, int64_t lda, const double* ds, int64_t p, int k,
                        double* dOut) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_REQUIRE(ctx, dA && dOut && p >= 1 && k >= 1 && lda >= p, "rankk_cov: bad arguments");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  // the covariance kernel reads its operand through 16-byte cp.async from rows zero-padded to a
  // multiple of 64: stage a padded copy (p x k doubles, negligible next to the p x p output)
  const int64_t pp = round_up(p, 64);
  const int KP = static_cast<int>(round_up(k, 4));
  double* padded = nullptr;
  FB_CUDA(ctx, cudaMallocAsync(reinterpret_cast<void**>(&padded), static_cast<size_t>(pp) * KP * sizeof(double), ctx->stream));
  int rc = fb_pad_rows(ctx, dA, lda, p, pp, k, KP, padded);
  if (rc == FABLE_OK) {
    PsiArgs a;
    a.Lw = padded; a.s2w = ds; a.pp = pp; a.p = p; a.k = k; a.KP = KP; a.B = 1;
    a.psi = dOut; a.psi_slots = 1; a.slot0 = 0; a.acc_sum = nullptr; a.acc_sq = nullptr;
    rc = fb_psi(ctx, a);
  }
  cudaFreeAsync(padded, ctx->stream);
  return rc;
}

int fable_dev_gemm(fable_ctx* ctx, int a_kfast, int b_kfast, int64_t M, int64_t N, int64_t K, double alpha,
                   const double* dA, int64_t lda, const double* dB, int64_t ldb, double* dC, int64_t ldc) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_REQUIRE(ctx, dA && dB && dC, "gemm: NULL pointer");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  return fb_gemm(ctx, a_kfast != 0, b_kfast != 0, M, N, K, alpha, dA, lda, dB, ldb, dC, ldc);
}

int fable_dev_syrk(fable_ctx* ctx, int a_kfast, int64_t M, int64_t K, double alpha, const double* dA, int64_t lda,
                   double* dC) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_REQUIRE(ctx, dA && dC, "syrk: NULL pointer");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  return fb_syrk(ctx, a_kfast != 0, M, K, alpha, dA, lda, dC, M);
}

int fable_dev_rowstats(fable_ctx* ctx, const double* dY, int64_t n, int64_t p, const double* dU, const double* dC,
                       int64_t ldc, int k, double* ds, double* dresid) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_REQUIRE(ctx, dY && dU && dC && dresid && n >= 1 && p >= 1 && ldc >= p, "rowstats: bad arguments");
  FB_REQUIRE(ctx, k >= 1 && k <= KSMALL, "rowstats: rank outside the streaming kernel's range");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  return fb_rowstats(ctx, dY, n, p, n, dU, n, dC, ldc, k, ds, dresid);
}

int fable_dev_syevj(fable_ctx* ctx, double* dG, int64_t m, double* dW, double* dQ, int* sweeps) {
  if (!ctx) return FABLE_ERR_INVALID;
  FB_REQUIRE(ctx, dG && dW && dQ, "syevj: NULL pointer");
  FB_CUDA(ctx, cudaSetDevice(ctx->device));
  return fb_syevj(ctx, dG, m, dW, dQ, sweeps);
}

}  // extern "C"
