// Symmetric eigensolver for the m x m Gram matrix (m = min(n,p)): one-sided (Hestenes) Jacobi.
// Together with the DMMA Gram and the back-multiplication this replaces LAPACK dgesdd behind
// base-R svd(Y) (reference R/FABLE_wrapper.R:23, 90); the wrappers need ALL r singular values
// (kMax rule, :27/:94) and return all r vectors (:61/:122), so this is a full decomposition.
//
// Right rotations Q orthogonalise the columns of A = G Q.  For symmetric positive semi-definite
// G the limit has columns lambda_i q_i: eigenvalue = column norm, eigenvector = normalised column,
// no separate accumulation of Q.  Each round applies m/2 disjoint rotations (round-robin
// "tournament" ordering), one CTA per column pair; a sweep is m-1 rounds.
#include <algorithm>
#include <cmath>
#include <numeric>

#include "common.cuh"

namespace {

constexpr int JT = 128;

__device__ __forceinline__ void block_sum3(double& a, double& b, double& c) {
  __shared__ double sh[3][JT / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { sh[0][w] = a; sh[1][w] = b; sh[2][w] = c; }
  __syncthreads();
  a = 0.0; b = 0.0; c = 0.0;
#pragma unroll
  for (int i = 0; i < JT / 32; ++i) { a += sh[0][i]; b += sh[1][i]; c += sh[2][i]; }
}

__global__ void __launch_bounds__(JT)
k_jacobi_round(double* __restrict__ A, int64_t m, int64_t ld, int round, int N, double tol, int* __restrict__ nrot) {
  const int i = blockIdx.x;
  int ca, cb;
  if (i == 0) { ca = N - 1; cb = round; }
  else { ca = (round + i) % (N - 1); cb = (round - i + (N - 1)) % (N - 1); }
  if (ca >= m || cb >= m) return;   // padding player when m is odd
  if (ca > cb) { const int tmp = ca; ca = cb; cb = tmp; }
  double* x = A + ld * ca;
  double* y = A + ld * cb;
  double alpha = 0.0, beta = 0.0, gamma = 0.0;
  for (int64_t r = threadIdx.x; r < m; r += JT) {
    const double xv = x[r], yv = y[r];
    alpha = fma(xv, xv, alpha); beta = fma(yv, yv, beta); gamma = fma(xv, yv, gamma);
  }
  block_sum3(alpha, beta, gamma);
  if (gamma == 0.0 || fabs(gamma) <= tol * sqrt(alpha) * sqrt(beta)) return;
  if (threadIdx.x == 0) atomicAdd(nrot, 1);
  const double zeta = (beta - alpha) / (2.0 * gamma);
  const double tt = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double cs = 1.0 / sqrt(1.0 + tt * tt), sn = cs * tt;
  for (int64_t r = threadIdx.x; r < m; r += JT) {
    const double xv = x[r], yv = y[r];
    x[r] = cs * xv - sn * yv;
    y[r] = sn * xv + cs * yv;
  }
}

__global__ void __launch_bounds__(JT)
k_colnorm(const double* __restrict__ A, int64_t m, int64_t ld, double* __restrict__ nrm) {
  const double* x = A + ld * blockIdx.x;
  double a = 0.0, b = 0.0, c = 0.0;
  for (int64_t r = threadIdx.x; r < m; r += JT) a = fma(x[r], x[r], a);
  block_sum3(a, b, c);
  if (threadIdx.x == 0) nrm[blockIdx.x] = sqrt(a);
}

// Q[:, i] = A[:, perm[i]] / nrm[perm[i]]; W[i] = nrm[perm[i]]
__global__ void __launch_bounds__(JT)
k_gather_normalise(const double* __restrict__ A, int64_t m, int64_t ld, const int* __restrict__ perm,
                   const double* __restrict__ nrm, double* __restrict__ Q, double* __restrict__ W) {
  const int src = perm[blockIdx.x];
  const double nv = nrm[src];
  const double inv = nv > 0.0 ? 1.0 / nv : 0.0;
  for (int64_t r = threadIdx.x; r < m; r += JT) Q[r + m * blockIdx.x] = A[r + ld * src] * inv;
  if (threadIdx.x == 0) W[blockIdx.x] = nv;
}

}  // namespace

int fb_syevj(fable_ctx* ctx, double* dG, int64_t m, double* dW, double* dQ, int* sweeps_out) {
  FB_REQUIRE(ctx, m > 0 && m < (1 << 30), "syevj: bad dimension");
  const int N = static_cast<int>(m + (m & 1));
  const double tol = std::sqrt(static_cast<double>(m)) * 2.220446049250313e-16;
  DevBuf cnt, nrm, perm;
  FB_CUDA(ctx, cnt.alloc(sizeof(int)));
  FB_CUDA(ctx, nrm.alloc(m * sizeof(double)));
  FB_CUDA(ctx, perm.alloc(m * sizeof(int)));
  int sweeps = 0;
  const int max_sweeps = 40;
  if (m > 1) {
    for (; sweeps < max_sweeps; ++sweeps) {
      FB_CUDA(ctx, cudaMemsetAsync(cnt.ptr, 0, sizeof(int), ctx->stream));
      for (int r = 0; r < N - 1; ++r) {
        k_jacobi_round<<<N / 2, JT, 0, ctx->stream>>>(dG, m, m, r, N, tol, static_cast<int*>(cnt.ptr));
        ++ctx->launches;
      }
      FB_CUDA(ctx, cudaGetLastError());
      int nrot = 0;
      FB_CUDA(ctx, cudaMemcpyAsync(&nrot, cnt.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      if (nrot == 0) { ++sweeps; break; }
    }
    if (sweeps >= max_sweeps) return ctx->fail(FABLE_ERR_NUMERIC, "syevj: Jacobi sweeps did not converge");
  }
  k_colnorm<<<static_cast<unsigned>(m), JT, 0, ctx->stream>>>(dG, m, m, nrm.d());
  FB_LAUNCHED(ctx);
  std::vector<double> h(m);
  FB_CUDA(ctx, cudaMemcpyAsync(h.data(), nrm.d(), m * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  std::vector<int> order(m);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h[a] > h[b]; });
  FB_CUDA(ctx, cudaMemcpyAsync(perm.ptr, order.data(), m * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  k_gather_normalise<<<static_cast<unsigned>(m), JT, 0, ctx->stream>>>(dG, m, m, static_cast<const int*>(perm.ptr),
                                                                      nrm.d(), dQ, dW);
  FB_LAUNCHED(ctx);
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (sweeps_out) *sweeps_out = sweeps;
  return FABLE_OK;
}
