// Small deterministic reductions and element-wise finishing steps of the row posterior.
// All reductions are two-stage with a FIXED grid and shuffle trees, so results are bit-stable
// run to run (no floating-point atomics).
#include "common.cuh"

namespace {

constexpr int RB = 256;      // threads per reduction block
constexpr int RGRID = 296;   // 2 x 148 SMs

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sh[RB / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = (threadIdx.x < RB / 32) ? sh[threadIdx.x] : 0.0;
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;  // valid in thread 0
}

struct LogScaled {
  const double* x; double scale;
  __device__ double operator()(int64_t j) const { return log(x[j] * scale); }
};
// CPPLogLikelihoodEval (cpp:95-125): Term1 = -n sum_j log(sqrt(SigmaSq_j)); Term2 = -0.5 sum_j resid_j / SigmaSq_j
struct LogSd {
  const double* sig;
  __device__ double operator()(int64_t j) const { return log(sqrt(sig[j])); }
};
struct ResidOverSig {
  const double* resid; const double* sig;
  __device__ double operator()(int64_t j) const { return resid[j] / sig[j]; }
};
// R prototype criterion (R/FABLE_code.R:20-22 with sigma^2 = resid / (n - k)): (resid_j / sd_j)^2 summed over a column
struct ResidOverScaledSelf {
  const double* resid; double inv;
  __device__ double operator()(int64_t j) const { return resid[j] / (resid[j] * inv); }
};
struct QOverSig {
  const double* q; const double* resid; double inv_n;
  __device__ double operator()(int64_t j) const { return q[j] / (resid[j] * inv_n); }
};

template <typename F>
__global__ void __launch_bounds__(RB) k_reduce1(F f, int64_t len, double* __restrict__ part) {
  double acc = 0.0;
  for (int64_t j = blockIdx.x * static_cast<int64_t>(RB) + threadIdx.x; j < len;
       j += static_cast<int64_t>(gridDim.x) * RB)
    acc += f(j);
  acc = block_sum(acc);
  if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(RB) k_reduce2(const double* __restrict__ part, int nparts,
                                                double* __restrict__ out) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < nparts; i += RB) acc += part[i];
  acc = block_sum(acc);
  if (threadIdx.x == 0) out[0] = acc;
}

template <typename F>
int reduce_to_host(fable_ctx* ctx, F f, int64_t len, double* host_out) {
  DevBuf part;
  FB_CUDA(ctx, part.alloc((RGRID + 1) * sizeof(double)));
  k_reduce1<<<RGRID, RB, 0, ctx->stream>>>(f, len, part.d());
  FB_LAUNCHED(ctx);
  k_reduce2<<<1, RB, 0, ctx->stream>>>(part.d(), RGRID, part.d() + RGRID);
  FB_LAUNCHED(ctx);
  FB_CUDA(ctx, cudaMemcpyAsync(host_out, part.d() + RGRID, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FABLE_OK;
}

// ---- all-rank criterion from one pass ----------------------------------------------------------
// When (U, V, d) is a consistent SVD of Y -- Y'U = V diag(d) and U'U = I -- the residual column sums of squares of the
// rank-k' fit (cpp:161-166) are s_j - sum_{l <= k'} (d_l V_jl)^2 for EVERY k' at once.  k_prefix_sumlog forms
// sumlog[k'-1] = sum_j log((s_j - q_j(k')) / n) for k' = 1 .. kMax: one thread per variable walks the ranks with a
// running q_j; the sums over variables are two-stage (per-CTA partials in shared memory, then a fixed-order sum over
// the CTAs), so the result is bit-stable.  The host uses these values to steer the bisection (api.cu) and falls back
// on the explicit residual wherever two criteria are too close to call.
constexpr int PL_CHUNK = 32;     // ranks per shared-memory round
__global__ void __launch_bounds__(RB)
k_prefix_sumlog(const double* __restrict__ V, int64_t ldv, const double* __restrict__ d, const double* __restrict__ s, int64_t p,
                int kMax, double inv_n, double* __restrict__ part /* [gridDim.x][kMax] */) {
  __shared__ double sh[RB / 32][PL_CHUNK];
  const int64_t j = blockIdx.x * static_cast<int64_t>(RB) + threadIdx.x;
  const bool in = j < p;
  const double sj = in ? s[j] : 0.0;
  double q = 0.0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int l0 = 0; l0 < kMax; l0 += PL_CHUNK) {
    const int nl = kMax - l0 < PL_CHUNK ? kMax - l0 : PL_CHUNK;
    for (int l = 0; l < nl; ++l) {
      double v = 0.0;
      if (in) {
        const double c = V[j + ldv * (l0 + l)] * d[l0 + l];
        q = fma(c, c, q);
        v = log((sj - q) * inv_n);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) sh[warp][l] = v;
    }
    __syncthreads();
    if (threadIdx.x < nl) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < RB / 32; ++w) t += sh[w][threadIdx.x];
      part[static_cast<int64_t>(blockIdx.x) * kMax + l0 + threadIdx.x] = t;
    }
    __syncthreads();
  }
}
// out[l] = sum over the CTAs of part[cta][l], in CTA order
__global__ void __launch_bounds__(RB) k_prefix_sum_parts(const double* __restrict__ part, int64_t nparts, int kMax, double* __restrict__ out) {
  const int l = blockIdx.x * RB + threadIdx.x;
  if (l >= kMax) return;
  double t = 0.0;
  for (int64_t c = 0; c < nparts; ++c) t += part[c * kMax + l];
  out[l] = t;
}

// max_jl |T_jl - d_l V_jl| and max_jl |d_l V_jl|: is Y'U = V diag(d)?
struct AbsDiffTV {
  const double* T; int64_t ldt; const double* V; int64_t ldv; const double* d; int64_t p;
  __device__ double2 operator()(int64_t idx) const {
    const int64_t l = idx / p, j = idx - l * p;
    const double c = V[j + ldv * l] * d[l];
    return make_double2(fabs(T[j + ldt * l] - c), fabs(c));
  }
};
// max |A - I| of a k x k matrix
struct AbsDiffEye {
  const double* A; int64_t k;
  __device__ double2 operator()(int64_t idx) const {
    const int64_t c = idx / k, r = idx - c * k;
    return make_double2(fabs(A[idx] - (r == c ? 1.0 : 0.0)), 1.0);
  }
};
template <typename F>
__global__ void __launch_bounds__(RB) k_max2(F f, int64_t len, double* __restrict__ part) {
  __shared__ double sh[2][RB / 32];
  double a = 0.0, b = 0.0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(RB) + threadIdx.x; i < len; i += static_cast<int64_t>(gridDim.x) * RB) {
    const double2 v = f(i);
    a = fmax(a, v.x); b = fmax(b, v.y);       // (a NaN difference must not vanish: checked below)
    if (v.x != v.x) a = INFINITY;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a = fmax(a, __shfl_xor_sync(0xffffffffu, a, o)); b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o)); }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = a; sh[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < RB / 32; ++w) { a = fmax(a, sh[0][w]); b = fmax(b, sh[1][w]); }
    part[2 * blockIdx.x] = a; part[2 * blockIdx.x + 1] = b;
  }
}
template <typename F>
int max2_to_host(fable_ctx* ctx, F f, int64_t len, double* m0, double* m1) {
  DevBuf part;
  FB_CUDA(ctx, part.alloc(2 * RGRID * sizeof(double)));
  k_max2<<<RGRID, RB, 0, ctx->stream>>>(f, len, part.d());
  FB_LAUNCHED(ctx);
  double h[2 * RGRID];
  FB_CUDA(ctx, cudaMemcpyAsync(h, part.d(), sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *m0 = 0.0; *m1 = 0.0;
  for (int i = 0; i < RGRID; ++i) { *m0 = h[2 * i] > *m0 ? h[2 * i] : *m0; *m1 = h[2 * i + 1] > *m1 ? h[2 * i + 1] : *m1; }
  return FABLE_OK;
}

// sig = resid/n ; G0 = coef * c ; gnd = g0d0 + s - imp*q ; sigmahat = gnd/(gamma_n-2)
__global__ void __launch_bounds__(256)
k_finish(const double* __restrict__ c, int64_t ldc, const double* __restrict__ s, const double* __restrict__ q,
         const double* __restrict__ resid, int64_t p, int k, double inv_n, double coef, double g0d0,
         double imp, double inv_gm2, double* __restrict__ G0, int64_t ldg, double* __restrict__ gnd,
         double* __restrict__ sig, double* __restrict__ sigmahat) {
  const int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (j >= p) return;
  if (sig) sig[j] = resid[j] * inv_n;
  if (G0) for (int l = 0; l < k; ++l) G0[j + ldg * l] = coef * c[j + ldc * l];
  if (gnd || sigmahat) {
    const double g = g0d0 + s[j] - imp * q[j];
    if (gnd) gnd[j] = g;
    if (sigmahat) sigmahat[j] = g * inv_gm2;
  }
}

// ---- coverage-correction sums without storing B ------------------------------------------
// B_uv = sqrt(1 + (g_uu g_vv + g_uv^2)/(g_uu s_v + s_u g_vv)), diagonal ratio halved
// (R/FABLE_code.R:196-202; cpp:413-421) with g = G0 G0' formed on the fly from the p x k factor.
constexpr int VT = 64;
__global__ void __launch_bounds__(256)
k_varinfl(const double* __restrict__ sig, const double* __restrict__ G0, int64_t ldg, int64_t p, int k,
          const double* __restrict__ gdiag, double* __restrict__ part_upper, double* __restrict__ part_diag) {
  __shared__ double As[16][VT], Bs[16][VT];
  // upper-triangle tile enumeration (column-major)
  const int64_t t = blockIdx.x;
  int64_t Jl = static_cast<int64_t>((sqrt(8.0 * static_cast<double>(t) + 1.0) - 1.0) * 0.5);
  while (Jl * (Jl + 1) / 2 > t) --Jl;
  while ((Jl + 1) * (Jl + 2) / 2 <= t) ++Jl;
  const int64_t Il = t - Jl * (Jl + 1) / 2;
  const int64_t u0 = Il * VT, v0 = Jl * VT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double c[2][8];
#pragma unroll
  for (int iu = 0; iu < 2; ++iu)
#pragma unroll
    for (int iv = 0; iv < 8; ++iv) c[iu][iv] = 0.0;
  for (int l0 = 0; l0 < k; l0 += 16) {
    const int kc = min(16, k - l0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < kc * VT; idx += 256) {
      const int l = idx >> 6, r = idx & 63;
      As[l][r] = (u0 + r < p) ? G0[u0 + r + ldg * (l0 + l)] : 0.0;
      Bs[l][r] = (v0 + r < p) ? G0[v0 + r + ldg * (l0 + l)] : 0.0;
    }
    __syncthreads();
    for (int l = 0; l < kc; ++l) {
      const double a0 = As[l][lane], a1 = As[l][lane + 32];
#pragma unroll
      for (int iv = 0; iv < 8; ++iv) {
        const double b = Bs[l][warp * 8 + iv];
        c[0][iv] = fma(a0, b, c[0][iv]);
        c[1][iv] = fma(a1, b, c[1][iv]);
      }
    }
  }
  double su = 0.0, sd = 0.0;
#pragma unroll
  for (int iu = 0; iu < 2; ++iu) {
    const int64_t u = u0 + lane + 32 * iu;
    if (u >= p) continue;
    const double gu = gdiag[u], s_u = sig[u];
#pragma unroll
    for (int iv = 0; iv < 8; ++iv) {
      const int64_t v = v0 + warp * 8 + iv;
      if (v >= p || u > v) continue;
      const double gv = gdiag[v], s_v = sig[v];
      double b = (gu * gv + c[iu][iv] * c[iu][iv]) / (gu * s_v + s_u * gv);
      if (u == v) b = b / 2.0;
      b = sqrt(1.0 + b);
      su += b;
      if (u == v) sd += b;
    }
  }
  su = block_sum(su);
  __syncthreads();
  sd = block_sum(sd);
  if (threadIdx.x == 0) { part_upper[t] = su; part_diag[t] = sd; }
}

__global__ void __launch_bounds__(256)
k_gdiag(const double* __restrict__ G0, int64_t ldg, int64_t p, int k, double* __restrict__ gdiag) {
  const int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (j >= p) return;
  double acc = 0.0;
  for (int l = 0; l < k; ++l) { const double x = G0[j + ldg * l]; acc = fma(x, x, acc); }
  gdiag[j] = acc;
}

struct Ident {
  const double* x;
  __device__ double operator()(int64_t j) const { return x[j]; }
};

// dense-G forms (CPPcov_correct_matrix cpp:401-428 / R cov_correct_matrix R/FABLE_code.R:189-206)
__global__ void __launch_bounds__(256)
k_covcorrect_dense(const double* __restrict__ sig, const double* __restrict__ G, int64_t p,
                   double* __restrict__ out, int upper_vector) {
  const int64_t u = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t v = blockIdx.y;
  if (u >= p || (upper_vector && u > v)) return;
  const double gu = G[u + p * u], gv = G[v + p * v], guv = G[u + p * v];
  double b = (gu * gv + guv * guv) / (gu * sig[v] + sig[u] * gv);
  if (u == v) b = b / 2.0;
  b = sqrt(1.0 + b);
  if (upper_vector) out[v * (v + 1) / 2 + u] = b; else out[u + p * v] = b;
}

__global__ void __launch_bounds__(256)
k_sum_partials(const double* __restrict__ part, int64_t nparts, int64_t len, int64_t stride, double alpha,
               double* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= len) return;
  double acc = 0.0;
  for (int64_t s = 0; s < nparts; ++s) acc += part[s * stride + i];
  out[i] = alpha * acc;
}

}  // namespace

int fb_prefix_sumlog(fable_ctx* ctx, const double* dV, int64_t ldv, const double* dd, const double* ds, int64_t p, int kMax,
                     int64_t n, double* host_out) {
  const int64_t nparts = ceil_div(p, RB);
  DevBuf part, out;
  FB_CUDA(ctx, part.alloc(static_cast<size_t>(nparts) * kMax * sizeof(double)));
  FB_CUDA(ctx, out.alloc(static_cast<size_t>(kMax) * sizeof(double)));
  k_prefix_sumlog<<<static_cast<unsigned>(nparts), RB, 0, ctx->stream>>>(dV, ldv, dd, ds, p, kMax, 1.0 / static_cast<double>(n), part.d());
  FB_LAUNCHED(ctx);
  k_prefix_sum_parts<<<static_cast<unsigned>(ceil_div(kMax, RB)), RB, 0, ctx->stream>>>(part.d(), nparts, kMax, out.d());
  FB_LAUNCHED(ctx);
  FB_CUDA(ctx, cudaMemcpyAsync(host_out, out.d(), static_cast<size_t>(kMax) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FABLE_OK;
}

int fb_svd_consistency(fable_ctx* ctx, const double* dT, int64_t ldt, const double* dV, int64_t ldv, const double* dd, int64_t p,
                       int k, double* max_diff, double* max_abs) {
  return max2_to_host(ctx, AbsDiffTV{dT, ldt, dV, ldv, dd, p}, p * static_cast<int64_t>(k), max_diff, max_abs);
}

int fb_identity_err(fable_ctx* ctx, const double* dA, int64_t k, double* max_diff) {
  double dummy;
  return max2_to_host(ctx, AbsDiffEye{dA, k}, k * k, max_diff, &dummy);
}

int fb_sum_log_scaled(fable_ctx* ctx, const double* d_x, int64_t p, double scale, double* host_out) {
  return reduce_to_host(ctx, LogScaled{d_x, scale}, p, host_out);
}

int fb_resid_over_scaled_self_sum(fable_ctx* ctx, const double* d_resid, int64_t p, double inv, double* host_sum) {
  return reduce_to_host(ctx, ResidOverScaledSelf{d_resid, inv}, p, host_sum);
}

int fb_loglik_terms(fable_ctx* ctx, const double* d_resid, const double* d_sig, int64_t p, double* sum_log_sd,
                    double* sum_resid_over_sig) {
  FB_TRY(reduce_to_host(ctx, LogSd{d_sig}, p, sum_log_sd));
  return reduce_to_host(ctx, ResidOverSig{d_resid, d_sig}, p, sum_resid_over_sig);
}

// sum_j q_j / sig_j over the columns held here (the numerator of tauSq; all-reduced under a column shard)
int fb_q_over_sig_sum(fable_ctx* ctx, const double* d_q, const double* d_resid, int64_t n, int64_t p, double* host_sum) {
  return reduce_to_host(ctx, QOverSig{d_q, d_resid, 1.0 / static_cast<double>(n)}, p, host_sum);
}

int fb_tausq(fable_ctx* ctx, const double* d_q, const double* d_resid, int64_t n, int64_t p, int k,
             double* host_tausq) {
  double sum = 0.0;
  FB_TRY(fb_q_over_sig_sum(ctx, d_q, d_resid, n, p, &sum));
  // mean_j(q_j / sig_j) / (n k)   (cpp:383, :483, :577)
  *host_tausq = (sum / static_cast<double>(p)) / (static_cast<double>(n) * static_cast<double>(k));
  return FABLE_OK;
}

int fb_finish_posterior(fable_ctx* ctx, const double* d_c, int64_t ldc, const double* d_s, const double* d_q,
                        const double* d_resid, int64_t n, int64_t p, int k, double gamma0, double delta0sq,
                        double tausq, int shrunk, double* d_G0, int64_t ldg, double* d_gnd, double* d_sig,
                        double* d_sigmahat) {
  const double dn = static_cast<double>(n);
  const double w = dn + (1.0 / tausq);
  const double coef = shrunk ? (sqrt(dn) / w) : (1.0 / sqrt(dn));   // cpp:490/:584 vs cpp:390
  const double imp = dn / w;                                        // cpp:497/:590
  const double gamma_n = gamma0 + dn;                               // cpp:493/:586
  k_finish<<<static_cast<unsigned>(ceil_div(p, 256)), 256, 0, ctx->stream>>>(
      d_c, ldc, d_s, d_q, d_resid, p, k, 1.0 / dn, coef, gamma0 * delta0sq, imp, 1.0 / (gamma_n - 2.0), d_G0,
      ldg, d_gnd, d_sig, d_sigmahat);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}

int fb_gdiag(fable_ctx* ctx, const double* d_G0, int64_t ldg, int64_t p, int k, double* d_gdiag) {
  k_gdiag<<<static_cast<unsigned>(ceil_div(p, 256)), 256, 0, ctx->stream>>>(d_G0, ldg, p, k, d_gdiag);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}

int fb_var_inflation(fable_ctx* ctx, const double* d_sig, const double* d_G0, int64_t ldg, int64_t p, int k,
                     double* host_sum_upper, double* host_sum_diag) {
  const int64_t nT = ceil_div(p, VT), tiles = nT * (nT + 1) / 2;
  DevBuf gd, pu, pd;
  FB_CUDA(ctx, gd.alloc(p * sizeof(double)));
  FB_CUDA(ctx, pu.alloc(tiles * sizeof(double)));
  FB_CUDA(ctx, pd.alloc(tiles * sizeof(double)));
  k_gdiag<<<static_cast<unsigned>(ceil_div(p, 256)), 256, 0, ctx->stream>>>(d_G0, ldg, p, k, gd.d());
  FB_LAUNCHED(ctx);
  k_varinfl<<<static_cast<unsigned>(tiles), 256, 0, ctx->stream>>>(d_sig, d_G0, ldg, p, k, gd.d(), pu.d(), pd.d());
  FB_LAUNCHED(ctx);
  FB_TRY(reduce_to_host(ctx, Ident{pu.d()}, tiles, host_sum_upper));
  FB_TRY(reduce_to_host(ctx, Ident{pd.d()}, tiles, host_sum_diag));
  return FABLE_OK;
}

int fb_cov_correct_dense(fable_ctx* ctx, const double* d_sig, const double* d_G, int64_t p, double* d_out,
                         int upper_vector) {
  dim3 grid(static_cast<unsigned>(ceil_div(p, 256)), static_cast<unsigned>(p));
  k_covcorrect_dense<<<grid, 256, 0, ctx->stream>>>(d_sig, d_G, p, d_out, upper_vector);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}

int fb_sum_partials(fable_ctx* ctx, const double* d_part, int64_t nparts, int64_t len, int64_t stride,
                    double alpha, double* d_out) {
  k_sum_partials<<<static_cast<unsigned>(ceil_div(len, 256)), 256, 0, ctx->stream>>>(d_part, nparts, len, stride,
                                                                                     alpha, d_out);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}
