// Entry-wise posterior summaries of Psi over the Monte-Carlo draws on a selected index set:
// CCFABLEPostProcessingSubmatrix (reference src/updated-FABLE-functions.cpp:716-818) and, with every
// index selected, CPPCCFABLEPostProcessing (:631-713).  For each pair a <= b of selected variables:
//   v[m] = sum_l Lam_m[ja,l] Lam_m[jb,l]  (+ sigma2_m[ja] when ja == jb, an INDEX comparison: :770-772)
//   mean = accu(v)/MC (:782),  lower / upper = Armadillo quantile(v, alpha/2), quantile(v, 1-alpha/2)
//   (:786-787; Hyndman-Fan type 5), mirrored by SymmetrizeMatrix (:805-807).
// One CTA per pair: the MC values live in shared memory, are sorted there (bitonic network, padded with
// +inf to a power of two) and the two order statistics are interpolated exactly as the reference does.
// The samples arrive in the reference's own layout (MC contiguous per (variable, factor) column), so the
// loads along m are coalesced.  Bound: shared-memory sort, O(MC log^2 MC) per pair.
#include <cmath>

#include "common.cuh"

namespace {

constexpr int PT = 256;

__device__ __forceinline__ void pair_of(int64_t t, int& a, int& b) {
  int64_t j = static_cast<int64_t>((sqrt(8.0 * static_cast<double>(t) + 1.0) - 1.0) * 0.5);
  while (j * (j + 1) / 2 > t) --j;
  while ((j + 1) * (j + 2) / 2 <= t) ++j;
  b = static_cast<int>(j);
  a = static_cast<int>(t - j * (j + 1) / 2);
}

// Armadillo op_quantile (type 5): P_min = 0.5/N, P_max = (N-0.5)/N, h = N P + 0.5
__device__ double quantile5(const double* x, int N, double P) {
  const double dn = static_cast<double>(N);
  if (P < 0.5 / dn) return P < 0.0 ? -INFINITY : x[0];
  if (P > (dn - 0.5) / dn) return P > 1.0 ? INFINITY : x[N - 1];
  const double h = dn * P + 0.5;
  const int kk = static_cast<int>(floor(h));
  const double w = h - static_cast<double>(kk);
  return (1.0 - w) * x[kk - 1] + w * x[kk];
}

// Lsel: [S][k][MC], Ssel: [S][MC], idx: the selected (0-based) variable of each position
__global__ void __launch_bounds__(PT)
k_post(const double* __restrict__ Lsel, const double* __restrict__ Ssel, const int64_t* __restrict__ idx, int64_t MC,
       int k, int64_t S, int N2, double alpha, int64_t pair_begin, double* __restrict__ mean,
       double* __restrict__ lower, double* __restrict__ upper) {
  extern __shared__ double v[];
  __shared__ double red[PT / 32];
  int a, b;
  pair_of(pair_begin + blockIdx.x, a, b);
  const double* La = Lsel + static_cast<int64_t>(a) * k * MC;
  const double* Lb = Lsel + static_cast<int64_t>(b) * k * MC;
  const bool same = idx[a] == idx[b];
  double part = 0.0;
  for (int64_t m = threadIdx.x; m < N2; m += PT) {
    double x = INFINITY;
    if (m < MC) {
      x = La[m] * Lb[m];
      for (int l = 1; l < k; ++l) x += La[static_cast<int64_t>(l) * MC + m] * Lb[static_cast<int64_t>(l) * MC + m];
      if (same) x += Ssel[static_cast<int64_t>(a) * MC + m];
      part += x;
    }
    v[m] = x;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  for (int k2 = 2; k2 <= N2; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < N2; i += PT) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const double x = v[i], y = v[ixj];
          const bool asc = (i & k2) == 0;
          if ((x > y) == asc) { v[i] = y; v[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < PT / 32; ++i) tot += red[i];
    const double mu = tot / static_cast<double>(MC);
    const double lo = quantile5(v, static_cast<int>(MC), alpha / 2.0);
    const double hi = quantile5(v, static_cast<int>(MC), 1.0 - (alpha / 2.0));
    mean[a + S * b] = mu; mean[b + S * a] = mu;
    lower[a + S * b] = lo; lower[b + S * a] = lo;
    upper[a + S * b] = hi; upper[b + S * a] = hi;
  }
}

}  // namespace

int fb_post_processing(fable_ctx* ctx, const double* dLsel, const double* dSsel, const int64_t* d_idx, int64_t MC, int k,
                       int64_t S, double alpha, double* d_mean, double* d_lower, double* d_upper) {
  int N2 = 1;
  while (N2 < MC) N2 <<= 1;
  const size_t smem = static_cast<size_t>(N2) * sizeof(double);
  FB_REQUIRE(ctx, smem <= 200 * 1024, "post-processing: MC too large for the in-shared-memory selection (max 16384 draws)");
  FB_CUDA(ctx, cudaFuncSetAttribute(k_post, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int64_t pairs = S * (S + 1) / 2;
  const int64_t chunk = 1 << 20;
  for (int64_t t0 = 0; t0 < pairs; t0 += chunk) {
    const unsigned g = static_cast<unsigned>(pairs - t0 < chunk ? pairs - t0 : chunk);
    k_post<<<g, PT, smem, ctx->stream>>>(dLsel, dSsel, d_idx, MC, k, S, N2, alpha, t0, d_mean, d_lower, d_upper);
    FB_LAUNCHED(ctx);
    if (ctx->poll) {
      FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      if (ctx->poll(ctx->poll_user)) return ctx->fail(FABLE_ERR_INTERRUPT, "interrupted");   // cpp:793 pattern
    }
  }
  return FABLE_OK;
}
