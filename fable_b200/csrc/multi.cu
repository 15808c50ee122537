// Multi-GPU behind the C-ABI: one process (the R session), several B200s of one box.
//
// The reference has no multi-device code; its MC loop (src/updated-FABLE-functions.cpp:599-618) is
// embarrassingly parallel over draws.  fable_multi_sampler is CPPFABLESampler (cpp:532-629) on the draw partition
// (SURVEY 8(e) partition A): device g gets a contiguous share of the MC axis, draws it with the same counter-based stream
// one device would use (the union of the rows is bit-identical), writes its rows of LambdaSamples / SigmaSqSamples
// straight into the caller's matrices, and keeps its Psi sum / sumsq accumulators in HBM.  The one exchange step is done by
// a hand-written kernel over NVLink peer memory, no library collective: the compact accumulator tiles are cut into one
// slice per device, device g sums slice g over ALL devices' buffers (peer loads) and stores the result into device 0's
// buffer (peer store) -- a reduce-scatter and a gather in one pass, every NVLink direction busy at once.  Device 0
// alone then expands and downloads the dense p x p sums.
#include <new>
#include <thread>

#include "common.cuh"

namespace {

constexpr int kMaxDev = 16;

struct PeerPtrs { const double* src[kMaxDev]; };

// dst[i] = sum_g src[g][i] for i in [begin, end): src[g] are the devices' buffers (peer-mapped), dst is device 0's.
// Each element is read and written by exactly one thread of one device, so dst may alias src[0].
__global__ void __launch_bounds__(256) k_peer_sum(PeerPtrs p, int ndev, double* __restrict__ dst, int64_t begin, int64_t end) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 2;
  for (int64_t i = begin + (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 2; i < end; i += stride) {
    if (i + 1 < end) {
      double2 acc = *reinterpret_cast<const double2*>(p.src[0] + i);
      for (int g = 1; g < ndev; ++g) {
        const double2 x = *reinterpret_cast<const double2*>(p.src[g] + i);
        acc.x += x.x; acc.y += x.y;
      }
      *reinterpret_cast<double2*>(dst + i) = acc;
    } else {
      double acc = p.src[0][i];
      for (int g = 1; g < ndev; ++g) acc += p.src[g][i];
      dst[i] = acc;
    }
  }
}

}  // namespace

struct fable_multi {
  std::vector<fable_ctx*> ctx;
  std::string err;
  int fail(int code, const std::string& m) { err = m; return code; }
};

static thread_local std::string g_multi_error;

extern "C" {

int fable_multi_create(int ndev, const int* devices, uint64_t seed, fable_multi** out) {
  if (!out) { g_multi_error = "fable_multi_create: NULL out pointer"; return FABLE_ERR_INVALID; }
  *out = nullptr;
  if (ndev < 1 || ndev > kMaxDev) { g_multi_error = "fable_multi_create: need 1 <= ndev <= 16"; return FABLE_ERR_INVALID; }
  fable_multi* m = new (std::nothrow) fable_multi();
  if (!m) { g_multi_error = "out of host memory"; return FABLE_ERR_NOMEM; }
  for (int g = 0; g < ndev; ++g) {
    fable_ctx* c = nullptr;
    const int rc = fable_ctx_create(devices ? devices[g] : g, seed, &c);
    if (rc != FABLE_OK) {
      g_multi_error = std::string("device ") + std::to_string(devices ? devices[g] : g) + ": " + fable_last_error(nullptr);
      for (fable_ctx* x : m->ctx) fable_ctx_destroy(x);
      delete m;
      return rc;
    }
    m->ctx.push_back(c);
  }
  // every device reads every other device's accumulators and writes device 0's: peer access all round
  for (int a = 0; a < ndev; ++a)
    for (int b = 0; b < ndev; ++b) {
      if (a == b || m->ctx[a]->device == m->ctx[b]->device) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, m->ctx[a]->device, m->ctx[b]->device);
      cudaError_t e = cudaSuccess;
      if (can) {
        cudaSetDevice(m->ctx[a]->device);
        e = cudaDeviceEnablePeerAccess(m->ctx[b]->device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); e = cudaSuccess; }
      }
      if (!can || e != cudaSuccess) {
        g_multi_error = "no peer access between devices " + std::to_string(m->ctx[a]->device) + " and " +
                        std::to_string(m->ctx[b]->device) + " (the accumulator sum runs over NVLink peer memory; there is no host fallback)";
        (void)cudaGetLastError();
        for (fable_ctx* x : m->ctx) fable_ctx_destroy(x);
        delete m;
        return FABLE_ERR_CUDA;
      }
    }
  *out = m;
  return FABLE_OK;
}

void fable_multi_destroy(fable_multi* m) {
  if (!m) return;
  for (fable_ctx* c : m->ctx) fable_ctx_destroy(c);
  delete m;
}

const char* fable_multi_last_error(const fable_multi* m) { return m ? m->err.c_str() : g_multi_error.c_str(); }
int fable_multi_devices(const fable_multi* m) { return m ? static_cast<int>(m->ctx.size()) : 0; }
fable_ctx* fable_multi_ctx(fable_multi* m, int g) { return (m && g >= 0 && g < static_cast<int>(m->ctx.size())) ? m->ctx[g] : nullptr; }

int fable_multi_sampler(fable_multi* m, const double* Y, int64_t n, int64_t p, double gamma0, double delta0sq, int64_t MC,
                        const double* U, const double* V, const double* d, int64_t r, int kEst, double varInflation,
                        const double* gam_inj, const double* z_inj, int64_t m0, double* LambdaSamples, double* SigmaSqSamples,
                        double* SigmaPostMean, double* SigmaPlugIn, double* tausq, double* G, double* psi_sum, double* psi_sumsq) {
  if (!m) return FABLE_ERR_INVALID;
  if (MC < 1) return m->fail(FABLE_ERR_INVALID, "MC must be >= 1");
  if ((gam_inj == nullptr) != (z_inj == nullptr)) return m->fail(FABLE_ERR_INVALID, "injected draws need both gam and z");
  if ((psi_sum == nullptr) != (psi_sumsq == nullptr)) return m->fail(FABLE_ERR_INVALID, "psi_sum and psi_sumsq go together");
  const int ndev = static_cast<int>(m->ctx.size());
  std::vector<fable_posterior*> post(ndev, nullptr);
  std::vector<int> rc(ndev, FABLE_OK);
  struct Guard {
    std::vector<fable_posterior*>& p;
    ~Guard() { for (fable_posterior* x : p) fable_posterior_destroy(x); }
  } guard{post};
  auto share = [&](int g, int64_t& lo, int64_t& hi) {
    const int64_t base = MC / ndev, rem = MC % ndev;
    lo = g * base + (g < rem ? g : rem);
    hi = lo + base + (g < rem ? 1 : 0);
  };
  // one host thread per device: upload the inputs, row posterior, this device's share of the draws (its rows of the
  // caller's sample matrices: leading dimension MC), covariance pass into its own accumulators
  {
    std::vector<std::thread> th;
    for (int g = 0; g < ndev; ++g)
      th.emplace_back([&, g] {
        fable_ctx* c = m->ctx[g];
        int64_t lo, hi;
        share(g, lo, hi);
        rc[g] = fable_posterior_create(c, Y, n, p, gamma0, delta0sq, U, V, d, r, kEst, varInflation, &post[g]);
        if (rc[g] != FABLE_OK) return;
        if (psi_sum) rc[g] = fable_posterior_accum_reset(post[g]);
        if (rc[g] != FABLE_OK || hi <= lo) return;
        rc[g] = fable_posterior_sample_range(post[g], hi - lo, m0 + lo, MC, gam_inj ? gam_inj + lo * p : nullptr,
                                             z_inj ? z_inj + lo * p * kEst : nullptr,
                                             LambdaSamples ? LambdaSamples + lo : nullptr,
                                             SigmaSqSamples ? SigmaSqSamples + lo : nullptr, psi_sum != nullptr);
        if (rc[g] == FABLE_OK) rc[g] = fable_ctx_synchronize(c);
      });
    for (auto& t : th) t.join();
  }
  for (int g = 0; g < ndev; ++g)
    if (rc[g] != FABLE_OK) return m->fail(rc[g], std::string("device ") + std::to_string(m->ctx[g]->device) + ": " + fable_last_error(m->ctx[g]));
  fable_ctx* c0 = m->ctx[0];
  if (psi_sum) {
    // the exchange step over NVLink peer memory: device g sums slice g of the compact accumulator tiles over all
    // devices and stores it into device 0's buffers
    PeerPtrs ps{}, pq{};
    int64_t count = 0;
    double *dst_s = nullptr, *dst_q = nullptr;
    for (int g = 0; g < ndev; ++g) {
      double *a = nullptr, *b = nullptr;
      int64_t cnt = 0;
      const int e = fable_posterior_accum_dev(post[g], &a, &b, &cnt);
      if (e != FABLE_OK) return m->fail(e, fable_last_error(m->ctx[g]));
      ps.src[g] = a; pq.src[g] = b;
      if (g == 0) { dst_s = a; dst_q = b; count = cnt; }
    }
    if (ndev > 1) {
      for (int g = 0; g < ndev; ++g) {
        fable_ctx* c = m->ctx[g];
        const int64_t per = (count / ndev + 1) & ~int64_t(1);            // even slice starts keep the 16-byte loads aligned
        const int64_t b0 = per * g < count ? per * g : count, b1 = (g == ndev - 1 || per * (g + 1) > count) ? count : per * (g + 1);
        if (b1 <= b0) continue;
        if (cudaSetDevice(c->device) != cudaSuccess) return m->fail(FABLE_ERR_CUDA, "cudaSetDevice failed");
        const int grid = c->sms * 8;
        k_peer_sum<<<grid, 256, 0, c->stream>>>(ps, ndev, dst_s, b0, b1);
        k_peer_sum<<<grid, 256, 0, c->stream>>>(pq, ndev, dst_q, b0, b1);
        c->launches += 2;
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return m->fail(FABLE_ERR_CUDA, std::string("peer sum launch: ") + cudaGetErrorString(e));
      }
      for (int g = 0; g < ndev; ++g) {
        const int e = fable_ctx_synchronize(m->ctx[g]);
        if (e != FABLE_OK) return m->fail(e, fable_last_error(m->ctx[g]));
      }
    }
    const int e = fable_posterior_accum_fetch(post[0], psi_sum, psi_sumsq);
    if (e != FABLE_OK) return m->fail(e, fable_last_error(c0));
  }
  // the outputs that do not depend on the draws come from device 0 (cpp:620-627): MC = 0 draws, no Psi pass
  {
    const int e = fable_posterior_fixed_outputs(post[0], SigmaPostMean, SigmaPlugIn, tausq, G);
    if (e != FABLE_OK) return m->fail(e, fable_last_error(c0));
  }
  return FABLE_OK;
}

}  // extern "C"
