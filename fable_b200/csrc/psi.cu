// Covariance kernel: Psi_m = Lambda_m Lambda_m' + diag(sigma2_m), dense symmetric p x p FP64,
// for a batch of draws; optionally materialised to HBM and/or folded into entry-wise
// sum / sum-of-squares accumulators.
//
// Reference arithmetic: entry formula src/updated-FABLE-functions.cpp:672-684 (per pair j1<=j2:
// sum_l Lam[j1,l]*Lam[j2,l], + sigma2[j1] on the diagonal), mean :688, SymmetrizeMatrix :28-37;
// the same kernel with one "draw" (G0, SigmaHat) forms G = G0 G0' (:391, :491, :627) and
// CovPostMean = G + diag(.) (:512).
//
// Roofline: k is small (10-20), so each 8-byte output costs only 2k flops: the kernel is bound by
// HBM WRITE bandwidth (8 p^2 bytes per draw).  Only upper-triangle tiles are computed (halves the
// FP64 work to ~25% of the DFMA pipe at full write bandwidth); the mirrored tile is transposed
// through shared memory so that BOTH stores are full 256-byte coalesced runs, written once with
// streaming (evict-first) stores.
//
// Tile = 64 x 64, 256 threads: lane -> row u (two rows, u = lane + 32*iu), warp -> 8 columns v.
// Accumulators for the batch stay in registers across the draw loop (one read-modify-write of the
// HBM accumulators per batch: 16*S*p^2/(2B) bytes per draw, upper tiles only).
#include "common.cuh"

namespace {

constexpr int TILE = 64;
constexpr int KC = 12;            // rank chunk staged in shared memory (static smem < 48 KB)
constexpr int TSTRIDE = TILE + 1; // odd stride: conflict-free transposed 8-byte accesses

__device__ __forceinline__ void tile_of(int64_t t, int& I, int& J) {
  // column-major enumeration of the upper triangle: t = J(J+1)/2 + I, 0 <= I <= J
  int64_t j = static_cast<int64_t>((sqrt(8.0 * static_cast<double>(t) + 1.0) - 1.0) * 0.5);
  while (j * (j + 1) / 2 > t) --j;
  while ((j + 1) * (j + 2) / 2 <= t) ++j;
  J = static_cast<int>(j);
  I = static_cast<int>(t - j * (j + 1) / 2);
}

template <bool MAT, bool ACC>
__global__ void __launch_bounds__(256, ACC ? 2 : 3)
k_psi(PsiArgs a, int64_t tile_begin) {
  __shared__ double As[KC][TILE];
  __shared__ double Bs[KC][TILE];
  __shared__ double T[MAT ? TILE * TSTRIDE : 1];

  int I, J;
  tile_of(tile_begin + blockIdx.x, I, J);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t u0 = static_cast<int64_t>(I) * TILE, v0 = static_cast<int64_t>(J) * TILE;
  const int64_t p = a.p;
  const bool diag = (I == J);

  double acc_s[2][8], acc_q[2][8];
  if (ACC) {
#pragma unroll
    for (int iu = 0; iu < 2; ++iu)
#pragma unroll
      for (int iv = 0; iv < 8; ++iv) { acc_s[iu][iv] = 0.0; acc_q[iu][iv] = 0.0; }
  }

  for (int b = 0; b < a.B; ++b) {
    double c[2][8];
#pragma unroll
    for (int iu = 0; iu < 2; ++iu)
#pragma unroll
      for (int iv = 0; iv < 8; ++iv) c[iu][iv] = 0.0;

    const double* Lb = a.Lw + static_cast<int64_t>(b) * a.KP * a.pp;
    for (int l0 = 0; l0 < a.k; l0 += KC) {
      const int kc = min(KC, a.k - l0);
      __syncthreads();  // previous chunk / previous draw's mirror reads are done
      // stage kc x 64 rows of both operands (coalesced: 64 contiguous doubles per l)
      for (int idx = threadIdx.x; idx < kc * TILE; idx += 256) {
        const int l = idx >> 6, r = idx & 63;
        const double* src = Lb + static_cast<int64_t>(l0 + l) * a.pp;
        As[l][r] = (u0 + r < p) ? src[u0 + r] : 0.0;
        Bs[l][r] = (v0 + r < p) ? src[v0 + r] : 0.0;
      }
      __syncthreads();
#pragma unroll 4
      for (int l = 0; l < kc; ++l) {
        const double a0 = As[l][lane], a1 = As[l][lane + 32];
        const double2* bp = reinterpret_cast<const double2*>(&Bs[l][warp * 8]);
        const double2 b01 = bp[0], b23 = bp[1], b45 = bp[2], b67 = bp[3];
        const double bv[8] = {b01.x, b01.y, b23.x, b23.y, b45.x, b45.y, b67.x, b67.y};
#pragma unroll
        for (int iv = 0; iv < 8; ++iv) {
          c[0][iv] = fma(a0, bv[iv], c[0][iv]);
          c[1][iv] = fma(a1, bv[iv], c[1][iv]);
        }
      }
    }
    if (diag && a.s2w) {
#pragma unroll
      for (int iu = 0; iu < 2; ++iu)
#pragma unroll
        for (int iv = 0; iv < 8; ++iv)
          if (lane + 32 * iu == warp * 8 + iv) {
            const int64_t u = u0 + lane + 32 * iu;
            if (u < p) c[iu][iv] += a.s2w[static_cast<int64_t>(b) * a.pp + u];
          }
    }
    if (ACC) {
#pragma unroll
      for (int iu = 0; iu < 2; ++iu)
#pragma unroll
        for (int iv = 0; iv < 8; ++iv) {
          acc_s[iu][iv] += c[iu][iv];
          acc_q[iu][iv] = fma(c[iu][iv], c[iu][iv], acc_q[iu][iv]);
        }
    }
    if (MAT) {
      double* out = a.psi + static_cast<int64_t>((a.slot0 + b) % a.psi_slots) * p * p;
      // direct tile: column v, rows u -- 256-byte coalesced runs
#pragma unroll
      for (int iv = 0; iv < 8; ++iv) {
        const int64_t v = v0 + warp * 8 + iv;
        if (v < p) {
#pragma unroll
          for (int iu = 0; iu < 2; ++iu) {
            const int64_t u = u0 + lane + 32 * iu;
            if (u < p) __stcs(out + v * p + u, c[iu][iv]);
          }
        }
      }
      if (!diag) {
        // mirrored tile through shared memory: T[u][v], then columns u, rows v coalesced
#pragma unroll
        for (int iu = 0; iu < 2; ++iu)
#pragma unroll
          for (int iv = 0; iv < 8; ++iv) T[(lane + 32 * iu) * TSTRIDE + warp * 8 + iv] = c[iu][iv];
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int uc = warp * 8 + i;
          const int64_t u = u0 + uc;
          if (u < p) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int64_t v = v0 + lane + 32 * h;
              if (v < p) __stcs(out + u * p + v, T[uc * TSTRIDE + lane + 32 * h]);
            }
          }
        }
      }
    }
  }

  if (ACC) {
#pragma unroll
    for (int iv = 0; iv < 8; ++iv) {
      const int64_t v = v0 + warp * 8 + iv;
      if (v < p) {
#pragma unroll
        for (int iu = 0; iu < 2; ++iu) {
          const int64_t u = u0 + lane + 32 * iu;
          if (u < p) {
            a.acc_sum[v * p + u] += acc_s[iu][iv];
            a.acc_sq[v * p + u] += acc_q[iu][iv];
          }
        }
      }
    }
  }
}

// copy the strict upper triangle (element (u,v), u<v, at u + p*v) onto the lower one.
__global__ void __launch_bounds__(256) k_mirror_upper(double* __restrict__ A, int64_t p) {
  __shared__ double t[32][33];
  const int I = blockIdx.x, J = blockIdx.y;
  if (I > J) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int64_t u = static_cast<int64_t>(I) * 32 + tx, v = static_cast<int64_t>(J) * 32 + r;
    t[r][tx] = (u < p && v < p) ? A[u + p * v] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    // write element (row v, col u) = A[v + p*u] for v = J*32+tx, u = I*32+r, where u < v
    const int64_t v = static_cast<int64_t>(J) * 32 + tx, u = static_cast<int64_t>(I) * 32 + r;
    if (u < p && v < p && u < v) A[v + p * u] = t[tx][r];
  }
}

}  // namespace

int fb_psi(fable_ctx* ctx, const PsiArgs& a) {
  FB_REQUIRE(ctx, a.p > 0 && a.k > 0 && a.B > 0, "psi: empty problem");
  FB_REQUIRE(ctx, (a.acc_sum == nullptr) == (a.acc_sq == nullptr), "psi: need both accumulators");
  const bool mat = a.psi != nullptr, acc = a.acc_sum != nullptr;
  FB_REQUIRE(ctx, mat || acc, "psi: nothing to do");
  FB_REQUIRE(ctx, !mat || a.psi_slots > 0, "psi: psi_slots must be positive");
  const int64_t nT = ceil_div(a.p, TILE);
  const int64_t tiles = nT * (nT + 1) / 2;
  const int64_t chunk = 1 << 30;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (ctx->timing) {
    while (ctx->ev_pool.size() < ctx->ev_used + 2) {
      cudaEvent_t e;
      FB_CUDA(ctx, cudaEventCreate(&e));
      ctx->ev_pool.push_back(e);
    }
    ev0 = ctx->ev_pool[ctx->ev_used]; ev1 = ctx->ev_pool[ctx->ev_used + 1];
    ctx->ev_used += 2;
    FB_CUDA(ctx, cudaEventRecord(ev0, ctx->stream));
  }
  for (int64_t t0 = 0; t0 < tiles; t0 += chunk) {
    const unsigned g = static_cast<unsigned>(tiles - t0 < chunk ? tiles - t0 : chunk);
    if (mat && acc) k_psi<true, true><<<g, 256, 0, ctx->stream>>>(a, t0);
    else if (mat) k_psi<true, false><<<g, 256, 0, ctx->stream>>>(a, t0);
    else k_psi<false, true><<<g, 256, 0, ctx->stream>>>(a, t0);
    FB_LAUNCHED(ctx);
  }
  if (ev1) FB_CUDA(ctx, cudaEventRecord(ev1, ctx->stream));
  return FABLE_OK;
}

int fb_mirror_upper(fable_ctx* ctx, double* dA, int64_t p) {
  const unsigned nT = static_cast<unsigned>(ceil_div(p, 32));
  k_mirror_upper<<<dim3(nT, nT), 256, 0, ctx->stream>>>(dA, p);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}
