// Thin SVD of Y through the Gram matrix: replaces base-R svd(Y) inside the wrappers
// (reference R/FABLE_wrapper.R:23-26, 90-93; the unused CPPsvd at
// src/updated-FABLE-functions.cpp:40-74 has the same contract).
//   n <= p :  G = Y Y' (n x n, DMMA, contraction over the long axis p), G = U L U',
//             T = Y'U (p x r),  d_l = ||T[:,l]||,  V = T / d
//   p <  n :  G = Y'Y (p x p),  G = V L V',  T = Y V (n x r),  d_l = ||T[:,l]||,  U = T / d
// d is taken from the back-multiplied column norms (a Rayleigh quotient: second-order accurate
// in the eigenvector error) rather than sqrt(lambda), and the back-multiplied side has exactly
// unit columns.  Sign convention: the largest-|entry| of each V column is positive.
// Rank-deficient Y (the reference expects column-centred data, cpp:296, which makes one singular value exactly 0
// when n <= p): the eigensolver deflates Gram eigenvalues at its noise floor (sigma <~ 4 sqrt(m eps) sigma_1 cannot be
// resolved through the Gram matrix in any case) -- those triplets get d = 0 exactly and singular vectors that complete
// the others to orthonormal bases on both sides, as LAPACK's do.
#include "common.cuh"

namespace {

constexpr int ST = 256;

__device__ __forceinline__ double block_sum1(double v) {
  __shared__ double sh[ST / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < ST / 32; ++i) t += sh[i];
  return t;
}

// one CTA per column: d[l] = ||T[:,l]||, T[:,l] /= d[l]
__global__ void __launch_bounds__(ST) k_norm_scale(double* __restrict__ T, int64_t rows, int64_t ld,
                                                   double* __restrict__ d) {
  double* x = T + ld * blockIdx.x;
  double a = 0.0;
  for (int64_t r = threadIdx.x; r < rows; r += ST) a = fma(x[r], x[r], a);
  a = block_sum1(a);
  const double nv = sqrt(a), inv = nv > 0.0 ? 1.0 / nv : 0.0;
  for (int64_t r = threadIdx.x; r < rows; r += ST) x[r] *= inv;
  if (threadIdx.x == 0) d[blockIdx.x] = nv;
}

// one CTA per singular triplet: flip so that the largest-|entry| of V[:,l] is positive
__global__ void __launch_bounds__(ST) k_sign_canon(double* __restrict__ U, int64_t n, double* __restrict__ V,
                                                   int64_t p) {
  __shared__ double bestv[ST];
  __shared__ long long besti[ST];
  double* v = V + p * blockIdx.x;
  double bv = -1.0; long long bi = 0;
  for (int64_t r = threadIdx.x; r < p; r += ST) {
    const double a = fabs(v[r]);
    if (a > bv) { bv = a; bi = r; }     // first maximum within the thread's strided set
  }
  bestv[threadIdx.x] = bv; besti[threadIdx.x] = bi;
  __syncthreads();
  for (int s = ST / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      const double ov = bestv[threadIdx.x + s]; const long long oi = besti[threadIdx.x + s];
      if (ov > bestv[threadIdx.x] || (ov == bestv[threadIdx.x] && oi < besti[threadIdx.x])) {
        bestv[threadIdx.x] = ov; besti[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  const bool flip = v[besti[0]] < 0.0;
  __syncthreads();
  if (!flip) return;
  for (int64_t r = threadIdx.x; r < p; r += ST) v[r] = -v[r];
  double* u = U + n * blockIdx.x;
  for (int64_t r = threadIdx.x; r < n; r += ST) u[r] = -u[r];
}

// ---- column-wise pieces for the column-sharded SVD (api.cu svd_sharded) ------------------------------------
__global__ void __launch_bounds__(ST) k_col_sumsq(const double* __restrict__ T, int64_t rows, int64_t ld, double* __restrict__ out) {
  const double* x = T + ld * blockIdx.x;
  double a = 0.0;
  for (int64_t r = threadIdx.x; r < rows; r += ST) a = fma(x[r], x[r], a);
  a = block_sum1(a);
  if (threadIdx.x == 0) out[blockIdx.x] = a;
}

__global__ void __launch_bounds__(ST) k_col_absmax(const double* __restrict__ T, int64_t rows, int64_t ld, double* __restrict__ omax,
                                                    double* __restrict__ oidx, double* __restrict__ oval) {
  __shared__ double bestv[ST];
  __shared__ long long besti[ST];
  const double* v = T + ld * blockIdx.x;
  double bv = -1.0; long long bi = 0;
  for (int64_t r = threadIdx.x; r < rows; r += ST) {
    const double a = fabs(v[r]);
    if (a > bv) { bv = a; bi = r; }
  }
  bestv[threadIdx.x] = bv; besti[threadIdx.x] = bi;
  __syncthreads();
  for (int s = ST / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      const double ov = bestv[threadIdx.x + s]; const long long oi = besti[threadIdx.x + s];
      if (ov > bestv[threadIdx.x] || (ov == bestv[threadIdx.x] && oi < besti[threadIdx.x])) {
        bestv[threadIdx.x] = ov; besti[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    omax[blockIdx.x] = rows > 0 ? bestv[0] : -1.0;
    oidx[blockIdx.x] = static_cast<double>(besti[0]);
    oval[blockIdx.x] = rows > 0 ? v[besti[0]] : 0.0;
  }
}

__global__ void __launch_bounds__(ST) k_col_scale(double* __restrict__ T, int64_t rows, int64_t ld, const double* __restrict__ scale) {
  double* x = T + ld * blockIdx.x;
  const double sc = scale[blockIdx.x];
  for (int64_t r = threadIdx.x; r < rows; r += ST) x[r] *= sc;
}

}  // namespace

int fb_col_sumsq(fable_ctx* ctx, const double* dT, int64_t rows, int64_t ld, int64_t cols, double* d_out) {
  k_col_sumsq<<<static_cast<unsigned>(cols), ST, 0, ctx->stream>>>(dT, rows, ld, d_out);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}
int fb_col_absmax(fable_ctx* ctx, const double* dT, int64_t rows, int64_t ld, int64_t cols, double* d_max, double* d_idx,
                  double* d_val) {
  k_col_absmax<<<static_cast<unsigned>(cols), ST, 0, ctx->stream>>>(dT, rows, ld, d_max, d_idx, d_val);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}
int fb_col_scale(fable_ctx* ctx, double* dT, int64_t rows, int64_t ld, int64_t cols, const double* d_scale) {
  k_col_scale<<<static_cast<unsigned>(cols), ST, 0, ctx->stream>>>(dT, rows, ld, d_scale);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}

int fb_svd(fable_ctx* ctx, const double* dY, int64_t n, int64_t p, double* dU, double* dd, double* dV) {
  const int64_t r = n < p ? n : p;
  DevBuf G, W;
  FB_CUDA(ctx, G.alloc(static_cast<size_t>(r) * r * sizeof(double)));
  FB_CUDA(ctx, W.alloc(static_cast<size_t>(r) * sizeof(double)));
  int sweeps = 0;
  int64_t defl = 0;
  if (n <= p) {
    FB_TRY(fb_syrk(ctx, 0, n, p, 1.0, dY, n, G.d(), n));                        // Y Y' (upper tiles + mirror)
    FB_TRY(fb_syevj(ctx, G.d(), n, W.d(), dU, &sweeps, &defl));                 // U (n x n)
    FB_TRY(fb_gemm(ctx, 1, 1, p, r, n, 1.0, dY, n, dU, n, dV, p));             // T = Y'U
    k_norm_scale<<<static_cast<unsigned>(r), ST, 0, ctx->stream>>>(dV, p, p, dd);
    FB_LAUNCHED(ctx);
    if (defl > 0) {                        // Y'u = 0 for the deflated u: d = 0, V completed
      FB_CUDA(ctx, cudaMemsetAsync(dd + (r - defl), 0, static_cast<size_t>(defl) * sizeof(double), ctx->stream));
      FB_TRY(fb_complete_orthonormal(ctx, dV, p, p, r - defl, r));
    }
  } else {
    FB_TRY(fb_syrk(ctx, 1, p, n, 1.0, dY, n, G.d(), p));                        // Y'Y (upper tiles + mirror)
    FB_TRY(fb_syevj(ctx, G.d(), p, W.d(), dV, &sweeps, &defl));                 // V (p x p)
    FB_TRY(fb_gemm(ctx, 0, 1, n, r, p, 1.0, dY, n, dV, p, dU, n));             // T = Y V
    k_norm_scale<<<static_cast<unsigned>(r), ST, 0, ctx->stream>>>(dU, n, n, dd);
    FB_LAUNCHED(ctx);
    if (defl > 0) {
      FB_CUDA(ctx, cudaMemsetAsync(dd + (r - defl), 0, static_cast<size_t>(defl) * sizeof(double), ctx->stream));
      FB_TRY(fb_complete_orthonormal(ctx, dU, n, n, r - defl, r));
    }
  }
  k_sign_canon<<<static_cast<unsigned>(r), ST, 0, ctx->stream>>>(dU, n, dV, p);
  FB_LAUNCHED(ctx);
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FABLE_OK;
}
