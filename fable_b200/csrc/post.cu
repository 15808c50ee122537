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
// More than 16384 draws (the sort's power-of-two padding no longer fits in shared memory): k_post_select keeps the
// MC values as order-preserving 64-bit keys (shared memory up to 25600 draws, a per-CTA global scratch beyond) and
// finds the four order statistics by radix selection (8-bit digits, shared-memory histograms): O(MC) per pass,
// 2 x 8 passes + at most 2 minimum passes per pair, no draw-count limit.
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
  // P == P_max exactly (alpha MC == 1 for the upper level): kk == N, w == 0 -- the value is the maximum; x[N] is the
  // +inf padding of the sort (or past the buffer when N is a power of two) and 0 * inf would be NaN
  if (kk >= N) return x[N - 1];
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

// ---- radix selection for large MC -----------------------------------------------------------------------
__device__ __forceinline__ unsigned long long key_of(double x) {
  const unsigned long long u = static_cast<unsigned long long>(__double_as_longlong(x));
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);        // ascending keys <=> ascending doubles
}
__device__ __forceinline__ double value_of(unsigned long long kx) {
  const unsigned long long u = (kx >> 63) ? (kx & 0x7fffffffffffffffull) : ~kx;
  return __longlong_as_double(static_cast<long long>(u));
}

struct SelectShared {
  unsigned int hist[256];
  unsigned long long prefix;
  unsigned long long red[PT / 32];
  int r, dup;
};

// key of rank r (0-based, ascending) among keys[0..N); *rem = its index among its duplicates, *dup = their number
__device__ unsigned long long block_select(const unsigned long long* keys, int N, int r, SelectShared& sh, int* rem, int* dup) {
  unsigned long long prefix = 0, mask = 0;
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += PT) sh.hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += PT) {
      const unsigned long long kx = keys[i];
      if ((kx & mask) == prefix) atomicAdd(&sh.hist[(kx >> shift) & 255ull], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int cum = 0, bin = 0;
      for (; bin < 255; ++bin) {
        const int h = static_cast<int>(sh.hist[bin]);
        if (r < cum + h) break;
        cum += h;
      }
      sh.prefix = prefix | (static_cast<unsigned long long>(bin) << shift);
      sh.r = r - cum;
      sh.dup = static_cast<int>(sh.hist[bin]);
    }
    __syncthreads();
    prefix = sh.prefix; r = sh.r;
    mask |= 0xffull << shift;
    if (shift == 0) { *rem = sh.r; *dup = sh.dup; }
    __syncthreads();
  }
  return prefix;
}

// smallest key strictly above `kx` (kx itself if there is none)
__device__ unsigned long long block_next_above(const unsigned long long* keys, int N, unsigned long long kx, SelectShared& sh) {
  unsigned long long best = ~0ull;
  bool found = false;
  for (int i = threadIdx.x; i < N; i += PT) {
    const unsigned long long y = keys[i];
    if (y > kx && (!found || y < best)) { best = y; found = true; }
  }
  if (!found) best = ~0ull;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other < best ? other : best;
  }
  if ((threadIdx.x & 31) == 0) sh.red[threadIdx.x >> 5] = best;
  __syncthreads();
  unsigned long long out = sh.red[0];
  for (int i = 1; i < PT / 32; ++i) out = sh.red[i] < out ? sh.red[i] : out;
  __syncthreads();
  // ~0ull is not a key of a finite double (it would be -NaN): "none above"
  return out == ~0ull ? kx : out;
}

// quantile5 on unsorted keys: the two neighbouring order statistics by selection
__device__ double block_quantile5(const unsigned long long* keys, int N, double P, SelectShared& sh) {
  const double dn = static_cast<double>(N);
  int rem, dup;
  if (P < 0.5 / dn) return P < 0.0 ? -INFINITY : value_of(block_select(keys, N, 0, sh, &rem, &dup));
  if (P > (dn - 0.5) / dn) return P > 1.0 ? INFINITY : value_of(block_select(keys, N, N - 1, sh, &rem, &dup));
  const double h = dn * P + 0.5;
  const int kk = static_cast<int>(floor(h));
  const double w = h - static_cast<double>(kk);
  const unsigned long long k1 = block_select(keys, N, kk - 1, sh, &rem, &dup);
  unsigned long long k2 = k1;
  if (kk < N && rem + 1 >= dup) k2 = block_next_above(keys, N, k1, sh);
  return (1.0 - w) * value_of(k1) + w * value_of(k2);
}

__global__ void __launch_bounds__(PT)
k_post_select(const double* __restrict__ Lsel, const double* __restrict__ Ssel, const int64_t* __restrict__ idx, int64_t MC,
              int k, int64_t S, double alpha, int64_t pair_begin, int64_t pair_end, unsigned long long* __restrict__ scratch,
              double* __restrict__ mean, double* __restrict__ lower, double* __restrict__ upper) {
  extern __shared__ unsigned long long skeys[];
  __shared__ SelectShared sh;
  __shared__ double red[PT / 32];
  unsigned long long* keys = scratch ? scratch + static_cast<int64_t>(blockIdx.x) * MC : skeys;
  const int N = static_cast<int>(MC);
  for (int64_t pr = pair_begin + blockIdx.x; pr < pair_end; pr += gridDim.x) {
    int a, b;
    pair_of(pr, a, b);
    const double* La = Lsel + static_cast<int64_t>(a) * k * MC;
    const double* Lb = Lsel + static_cast<int64_t>(b) * k * MC;
    const bool same = idx[a] == idx[b];
    double part = 0.0;
    for (int m = threadIdx.x; m < N; m += PT) {
      double x = La[m] * Lb[m];
      for (int l = 1; l < k; ++l) x += La[static_cast<int64_t>(l) * MC + m] * Lb[static_cast<int64_t>(l) * MC + m];
      if (same) x += Ssel[static_cast<int64_t>(a) * MC + m];
      part += x;
      keys[m] = key_of(x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();                 // keys (shared or this CTA's global scratch) and red visible to the block
    const double lo = block_quantile5(keys, N, alpha / 2.0, sh);
    const double hi = block_quantile5(keys, N, 1.0 - (alpha / 2.0), sh);
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int i = 0; i < PT / 32; ++i) tot += red[i];
      const double mu = tot / static_cast<double>(MC);
      mean[a + S * b] = mu; mean[b + S * a] = mu;
      lower[a + S * b] = lo; lower[b + S * a] = lo;
      upper[a + S * b] = hi; upper[b + S * a] = hi;
    }
    __syncthreads();                 // red / keys are rewritten by the next pair
  }
}

constexpr int64_t SORT_MAX_MC = 16384;       // bitonic path: padded power of two in shared memory
constexpr int64_t SELECT_SMEM_MC = 25600;    // selection path: keys in shared memory up to here (200 KB)

}  // namespace

int fb_post_processing(fable_ctx* ctx, const double* dLsel, const double* dSsel, const int64_t* d_idx, int64_t MC, int k,
                       int64_t S, double alpha, double* d_mean, double* d_lower, double* d_upper) {
  const int64_t pairs = S * (S + 1) / 2;
  if (MC > SORT_MAX_MC) {
    FB_REQUIRE(ctx, MC < (1ll << 31), "post-processing: more than 2^31 - 1 draws");
    const bool in_smem = MC <= SELECT_SMEM_MC;
    const size_t smem = in_smem ? static_cast<size_t>(MC) * sizeof(unsigned long long) : 0;
    FB_CUDA(ctx, cudaFuncSetAttribute(k_post_select, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int64_t grid = static_cast<int64_t>(ctx->sms) * (in_smem ? 1 : 4);
    if (grid > pairs) grid = pairs;
    DevBuf scratch;
    if (!in_smem) FB_CUDA(ctx, scratch.alloc(static_cast<size_t>(grid) * MC * sizeof(unsigned long long)));
    // one launch for everything, or slices of 64 pairs per CTA when the caller wants to be polled (cpp:793 pattern)
    const int64_t slice = ctx->poll ? grid * 64 : pairs;
    for (int64_t t0 = 0; t0 < pairs; t0 += slice) {
      const int64_t t1 = t0 + slice < pairs ? t0 + slice : pairs;
      k_post_select<<<static_cast<unsigned>(grid), PT, smem, ctx->stream>>>(
          dLsel, dSsel, d_idx, MC, k, S, alpha, t0, t1, in_smem ? nullptr : static_cast<unsigned long long*>(scratch.ptr), d_mean,
          d_lower, d_upper);
      FB_LAUNCHED(ctx);
      FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // also: the scratch must outlive the kernel
      if (ctx->poll && ctx->poll(ctx->poll_user)) return ctx->fail(FABLE_ERR_INTERRUPT, "interrupted");
    }
    return FABLE_OK;
  }
  int N2 = 1;
  while (N2 < MC) N2 <<= 1;
  const size_t smem = static_cast<size_t>(N2) * sizeof(double);
  FB_CUDA(ctx, cudaFuncSetAttribute(k_post, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
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
