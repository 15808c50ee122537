// Row-posterior statistics: one streaming pass over Y (column-major n x p, column j contiguous).
// Reference: src/updated-FABLE-functions.cpp:371-383 == :470-483 == :564-577
//   YtU = V_k .* d_k                       -> c_jl, q_j = sum_l c_jl^2
//   sigsq_hat = colsum((Y - U_k D_k V_k')^2)/n   -> resid_j (explicit residual, never s_j - q_j)
//   sum(square(Y),0)                       -> s_j            (:499, :592)
// p independent columns: warp-per-column-group, lanes stride the n rows (coalesced), U_k staged
// in shared memory.  HBM-read bound: 8 n p bytes.  Ranks above KSMALL go through the DMMA
// residual GEMM (gemm.cu) instead.
#include "common.cuh"

namespace {

constexpr int CW = 4;          // columns per warp
constexpr int WARPS = 8;
constexpr int COLS_PER_BLOCK = CW * WARPS;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256)
k_cq(const double* __restrict__ V, int64_t ldv, const double* __restrict__ d, int64_t p, int k,
     double* __restrict__ c, int64_t ldc, double* __restrict__ q) {
  const int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (j >= p) return;
  double acc = 0.0;
  for (int l = 0; l < k; ++l) {
    const double x = V[j + ldv * l] * d[l];
    c[j + ldc * l] = x;
    acc = fma(x, x, acc);
  }
  if (q) q[j] = acc;
}

__global__ void __launch_bounds__(256)
k_colsumsq(const double* __restrict__ Y, int64_t n, int64_t p, int64_t ldy, double* __restrict__ s) {
  const int lane = threadIdx.x & 31;
  const int64_t j = blockIdx.x * static_cast<int64_t>(WARPS) + (threadIdx.x >> 5);
  if (j >= p) return;
  const double* y = Y + ldy * j;
  double acc = 0.0;
  for (int64_t i = lane; i < n; i += 32) acc = fma(y[i], y[i], acc);
  acc = warp_sum(acc);
  if (lane == 0) s[j] = acc;
}

// dynamic smem: Us[k][NI] (U chunk, row index contiguous) then cs[WARPS][k][CW] (c of the warp's columns)
// Each lane owns RW = 4 rows of a 128-row slab (rows lane, lane+32, lane+64, lane+96: every load is a
// fully coalesced 256-byte request) for the warp's CW = 4 columns: 16 independent Y loads are in flight per
// lane and slab, and each rank step costs 4 U loads + 2 broadcast c loads (16 bytes) for 16 FMAs.
constexpr int RW = 4;
__global__ void __launch_bounds__(256)
k_rowstats(const double* __restrict__ Y, int64_t n, int64_t p, int64_t ldy, const double* __restrict__ U,
           int64_t ldu, const double* __restrict__ c, int64_t ldc, int k, int NI,
           double* __restrict__ s_out, double* __restrict__ resid_out) {
  extern __shared__ __align__(16) double smem[];
  double* Us = smem;
  double* cs = smem + static_cast<size_t>(k) * NI;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t jb = blockIdx.x * static_cast<int64_t>(COLS_PER_BLOCK);
  for (int idx = threadIdx.x; idx < COLS_PER_BLOCK * k; idx += 256) {
    const int w = idx / (k * CW), rem = idx - w * k * CW, l = rem / CW, t = rem - l * CW;
    const int64_t j = jb + w * CW + t;
    cs[idx] = (j < p) ? c[j + ldc * l] : 0.0;
  }
  double ar[CW], as[CW];
#pragma unroll
  for (int t = 0; t < CW; ++t) { ar[t] = 0.0; as[t] = 0.0; }
  const double* cw = cs + static_cast<size_t>(warp) * k * CW;
  const int64_t j0 = jb + warp * CW;
  for (int64_t i0 = 0; i0 < n; i0 += NI) {
    const int ni = static_cast<int>(n - i0 < NI ? n - i0 : NI);
    __syncthreads();
    for (int idx = threadIdx.x; idx < k * NI; idx += 256) {
      const int l = idx / NI, i = idx - l * NI;
      Us[idx] = (i < ni) ? U[i0 + i + ldu * l] : 0.0;
    }
    __syncthreads();
    for (int ib = 0; ib < ni; ib += 32 * RW) {
      double y[RW][CW], f[RW][CW];
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        const int i = ib + lane + 32 * r;
#pragma unroll
        for (int t = 0; t < CW; ++t) {
          y[r][t] = (i < ni && j0 + t < p) ? __ldcs(Y + i0 + i + ldy * (j0 + t)) : 0.0;    // Y is read once: streaming load
          f[r][t] = 0.0;
        }
      }
      for (int l = 0; l < k; ++l) {
        const double2 c01 = *reinterpret_cast<const double2*>(cw + l * CW), c23 = *reinterpret_cast<const double2*>(cw + l * CW + 2);
        const double cv[CW] = {c01.x, c01.y, c23.x, c23.y};
#pragma unroll
        for (int r = 0; r < RW; ++r) {
          const double u = Us[l * NI + ib + lane + 32 * r];     // rows past ni are zero-filled
#pragma unroll
          for (int t = 0; t < CW; ++t) f[r][t] = fma(u, cv[t], f[r][t]);
        }
      }
#pragma unroll
      for (int r = 0; r < RW; ++r)
#pragma unroll
        for (int t = 0; t < CW; ++t) {
          const double d = y[r][t] - f[r][t];               // rows / columns out of range contribute 0 - 0
          ar[t] = fma(d, d, ar[t]);
          as[t] = fma(y[r][t], y[r][t], as[t]);
        }
    }
  }
#pragma unroll
  for (int t = 0; t < CW; ++t) {
    const double r = warp_sum(ar[t]), s = warp_sum(as[t]);
    const int64_t j = j0 + t;
    if (lane == 0 && j < p) {
      resid_out[j] = r;
      if (s_out) s_out[j] = s;
    }
  }
}

}  // namespace

int fb_cq(fable_ctx* ctx, const double* dV, int64_t ldv, const double* dd, int64_t p, int k, double* d_c,
          int64_t ldc, double* d_q) {
  k_cq<<<static_cast<unsigned>(ceil_div(p, 256)), 256, 0, ctx->stream>>>(dV, ldv, dd, p, k, d_c, ldc, d_q);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}

int fb_colsumsq(fable_ctx* ctx, const double* dY, int64_t n, int64_t p, int64_t ldy, double* d_s) {
  k_colsumsq<<<static_cast<unsigned>(ceil_div(p, WARPS)), 256, 0, ctx->stream>>>(dY, n, p, ldy, d_s);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}

// d_c: p x k (ldc) = YtU; outputs s (may be NULL), resid (sum of squared residuals, NOT divided by n)
int fb_rowstats(fable_ctx* ctx, const double* dY, int64_t n, int64_t p, int64_t ldy, const double* dU,
                int64_t ldu, const double* d_c, int64_t ldc, int k, double* d_s, double* d_resid) {
  const size_t budget = 96 * 1024 / sizeof(double);
  int64_t NI = (static_cast<int64_t>(budget) - COLS_PER_BLOCK * k) / k;
  NI = NI / 128 * 128;                         // whole 128-row slabs (zero-filled past n)
  FB_REQUIRE(ctx, NI >= 128, "rowstats: rank too large for the small-k kernel");
  if (NI > round_up(n, 128)) NI = round_up(n, 128);
  const size_t smem = (static_cast<size_t>(k) * NI + static_cast<size_t>(COLS_PER_BLOCK) * k) * sizeof(double);
  static bool attr_dev[64] = {};
  bool& attr_set = attr_dev[ctx->device & 63];
  if (!attr_set) {
    FB_CUDA(ctx, cudaFuncSetAttribute(k_rowstats, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_set = true;
  }
  k_rowstats<<<static_cast<unsigned>(ceil_div(p, COLS_PER_BLOCK)), 256, smem, ctx->stream>>>(
      dY, n, p, ldy, dU, ldu, d_c, ldc, k, static_cast<int>(NI), d_s, d_resid);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}
