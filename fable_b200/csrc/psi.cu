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
// HBM WRITE bandwidth (8 p^2 bytes per draw).  Design for that bound:
//   * only upper-triangle 64 x 64 tiles are computed (FP64 work ~25% of the DFMA pipe at full write
//     bandwidth); one CTA owns a tile for the whole batch, so the sum / sumsq accumulators live in
//     registers across the draw loop (one read-modify-write of the HBM accumulators per batch);
//   * the rank-k operand rows of draw b+1 are prefetched with cp.async (LDGSTS) into a second
//     shared-memory buffer while draw b is computed;
//   * a finished tile is staged in shared memory in BOTH orientations (direct [v][u] and mirrored
//     [u][v], padded row stride 66 so the transposed 16-byte stores are conflict-free) and leaves
//     the SM as 128 asynchronous 512-byte bulk stores (cp.async.bulk.global.shared::cta -> UBLKCP),
//     so HBM writes drain while the warps already compute the next draw; every global run is a
//     full, aligned 512-byte line group written exactly once.
// Tile = 64 x 64, 256 threads: lane -> row u (two rows, u = lane + 32*iu), warp -> 8 columns v.
#include "common.cuh"

namespace {

constexpr int TILE = 64;
constexpr int KC = 16;             // rank chunk per pipeline step
constexpr int MSTR = TILE + 2;     // mirrored staging row stride (doubles): 66 = 2 mod 16 -> conflict-free STS.128
constexpr int OP_DOUBLES = 2 * 2 * KC * TILE;                      // [buf][operand][l][row]
constexpr size_t OP_BYTES = OP_DOUBLES * sizeof(double);           // 32 KB
constexpr size_t STAGE_BYTES = (TILE * TILE + TILE * MSTR) * sizeof(double);   // 32 KB + 33 KB

__device__ __forceinline__ void tile_of(int64_t t, int& I, int& J) {
  // column-major enumeration of the upper triangle: t = J(J+1)/2 + I, 0 <= I <= J
  int64_t j = static_cast<int64_t>((sqrt(8.0 * static_cast<double>(t) + 1.0) - 1.0) * 0.5);
  while (j * (j + 1) / 2 > t) --j;
  while ((j + 1) * (j + 2) / 2 <= t) ++j;
  J = static_cast<int>(j);
  I = static_cast<int>(t - j * (j + 1) / 2);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// operand rows are re-read by every tile of a block row/column: keep them in L2 (evict_last)
__device__ __forceinline__ void cp_async16(void* dst, const void* src, uint64_t pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// the materialised draws are write-once streams: evict_first keeps them from displacing the operands
__device__ __forceinline__ void bulk_store(double* gdst, const double* ssrc, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
               "r"(smem_u32(ssrc)), "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Lw must be zero-padded to pp rows (pp a multiple of TILE) and 16-byte aligned.
template <bool MAT, bool ACC>
__global__ void __launch_bounds__(256, 2)
k_psi(PsiArgs a, int64_t tile_begin) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* op = reinterpret_cast<double*>(smem_raw);
  double* D = reinterpret_cast<double*>(smem_raw + OP_BYTES);   // direct  [v][u], stride TILE
  double* M = D + TILE * TILE;                                  // mirrored [u][v], stride MSTR

  int I, J;
  tile_of(tile_begin + blockIdx.x, I, J);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t u0 = static_cast<int64_t>(I) * TILE, v0 = static_cast<int64_t>(J) * TILE;
  const int64_t p = a.p;
  const bool diag = (I == J);
  const bool bulk = MAT && ((p & 1) == 0);       // 16-byte aligned runs of even length
  const uint32_t run_u = static_cast<uint32_t>((p - u0 < TILE ? p - u0 : TILE) * 8);   // bytes of one direct run
  const uint32_t run_v = static_cast<uint32_t>((p - v0 < TILE ? p - v0 : TILE) * 8);   // bytes of one mirrored run
  const int nch = (a.k + KC - 1) / KC;
  const int nsteps = a.B * nch;
  const uint64_t pol_keep = policy_evict_last(), pol_stream = policy_evict_first();

  auto prefetch = [&](int s, int buf) {
    const int b = s / nch, l0 = (s - b * nch) * KC;
    const int kc = min(KC, a.k - l0);
    const double* base = a.Lw + (static_cast<int64_t>(b) * a.KP + l0) * a.pp;
    double* dst = op + buf * (2 * KC * TILE);
    for (int idx = tid; idx < 2 * kc * (TILE / 2); idx += 256) {
      const int o = idx / (kc * (TILE / 2)), rem = idx - o * kc * (TILE / 2);
      const int l = rem >> 5, ch = rem & 31;
      cp_async16(dst + (o * KC + l) * TILE + 2 * ch, base + static_cast<int64_t>(l) * a.pp + (o ? v0 : u0) + 2 * ch, pol_keep);
    }
    cp_async_commit();
  };

  double acc_s[2][8], acc_q[2][8];
  if (ACC) {
#pragma unroll
    for (int iu = 0; iu < 2; ++iu)
#pragma unroll
      for (int iv = 0; iv < 8; ++iv) { acc_s[iu][iv] = 0.0; acc_q[iu][iv] = 0.0; }
  }
  double c[2][8];
#pragma unroll
  for (int iu = 0; iu < 2; ++iu)
#pragma unroll
    for (int iv = 0; iv < 8; ++iv) c[iu][iv] = 0.0;

  prefetch(0, 0);
  for (int s = 0; s < nsteps; ++s) {
    const int buf = s & 1;
    const int b = s / nch, ch = s - b * nch;
    const int kc = min(KC, a.k - ch * KC);
    cp_async_wait_all();
    __syncthreads();                     // operands of step s landed; buffer buf^1 is no longer read
    if (s + 1 < nsteps) prefetch(s + 1, buf ^ 1);

    const double* As = op + buf * (2 * KC * TILE);
    const double* Bs = As + KC * TILE;
#pragma unroll 2
    for (int l = 0; l < kc; ++l) {
      const double a0 = As[l * TILE + lane], a1 = As[l * TILE + lane + 32];
      const double2* bp = reinterpret_cast<const double2*>(Bs + l * TILE + warp * 8);
      const double2 b01 = bp[0], b23 = bp[1], b45 = bp[2], b67 = bp[3];
      const double bv[8] = {b01.x, b01.y, b23.x, b23.y, b45.x, b45.y, b67.x, b67.y};
#pragma unroll
      for (int iv = 0; iv < 8; ++iv) {
        c[0][iv] = fma(a0, bv[iv], c[0][iv]);
        c[1][iv] = fma(a1, bv[iv], c[1][iv]);
      }
    }
    if (ch + 1 < nch) continue;          // more rank chunks of this draw to come

    // ---- draw b of this tile is complete in registers ----------------------------------------
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
      if (bulk && tid < 2 * TILE) bulk_wait_read();   // the previous draw's bulk stores have read the staging
      __syncthreads();
#pragma unroll
      for (int iu = 0; iu < 2; ++iu) {
#pragma unroll
        for (int iv = 0; iv < 8; ++iv) D[(warp * 8 + iv) * TILE + lane + 32 * iu] = c[iu][iv];
        if (!diag) {
          double2* mp = reinterpret_cast<double2*>(M + (lane + 32 * iu) * MSTR + warp * 8);
          mp[0] = make_double2(c[iu][0], c[iu][1]); mp[1] = make_double2(c[iu][2], c[iu][3]);
          mp[2] = make_double2(c[iu][4], c[iu][5]); mp[3] = make_double2(c[iu][6], c[iu][7]);
        }
      }
      if (bulk) {
        fence_async_smem();              // generic-proxy smem writes -> visible to the async (bulk copy) proxy
        __syncthreads();
        if (tid < TILE) {
          if (v0 + tid < p)
            bulk_store(out + (v0 + tid) * p + u0, D + tid * TILE, run_u, pol_stream);
          bulk_commit();
        } else if (tid < 2 * TILE) {
          const int cidx = tid - TILE;
          if (!diag && u0 + cidx < p)
            bulk_store(out + (u0 + cidx) * p + v0, M + cidx * MSTR, run_v, pol_stream);
          bulk_commit();
        }
      } else {
        __syncthreads();
        // odd p: runs are not 16-byte aligned -> plain coalesced stores from the staging
        for (int cidx = warp; cidx < TILE; cidx += 8) {
          if (v0 + cidx < p)
            for (int h = 0; h < 2; ++h)
              if (u0 + lane + 32 * h < p) __stcs(out + (v0 + cidx) * p + u0 + lane + 32 * h, D[cidx * TILE + lane + 32 * h]);
          if (!diag && u0 + cidx < p)
            for (int h = 0; h < 2; ++h)
              if (v0 + lane + 32 * h < p) __stcs(out + (u0 + cidx) * p + v0 + lane + 32 * h, M[cidx * MSTR + lane + 32 * h]);
        }
        __syncthreads();
      }
    }
#pragma unroll
    for (int iu = 0; iu < 2; ++iu)
#pragma unroll
      for (int iv = 0; iv < 8; ++iv) c[iu][iv] = 0.0;
  }
  if (bulk && tid < 2 * TILE) bulk_wait_read();       // staging must stay valid until the last stores have read it

  if (ACC) {
    // one read-modify-write of the HBM accumulators per batch; loads are batched ahead of the adds
    // (c[] is free here) so the 16 independent requests per pass are all in flight together
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      double* __restrict__ dst = pass == 0 ? a.acc_sum : a.acc_sq;
#pragma unroll
      for (int iv = 0; iv < 8; ++iv) {
        const int64_t v = v0 + warp * 8 + iv;
#pragma unroll
        for (int iu = 0; iu < 2; ++iu) {
          const int64_t u = u0 + lane + 32 * iu;
          c[iu][iv] = (v < p && u < p) ? dst[v * p + u] : 0.0;
        }
      }
#pragma unroll
      for (int iv = 0; iv < 8; ++iv) {
        const int64_t v = v0 + warp * 8 + iv;
#pragma unroll
        for (int iu = 0; iu < 2; ++iu) {
          const int64_t u = u0 + lane + 32 * iu;
          if (v < p && u < p) dst[v * p + u] = c[iu][iv] + (pass == 0 ? acc_s[iu][iv] : acc_q[iu][iv]);
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

// dst[(l*pp) + j] = j < p ? src[l*lds + j] : 0   (zero-padded, 16-byte aligned operand layout)
__global__ void __launch_bounds__(256) k_pad_rows(const double* __restrict__ src, int64_t lds, int64_t p, int64_t pp,
                                                  int k, double* __restrict__ dst) {
  const int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (j >= pp) return;
  for (int l = 0; l < k; ++l) dst[static_cast<int64_t>(l) * pp + j] = (j < p) ? src[static_cast<int64_t>(l) * lds + j] : 0.0;
}

template <bool MAT, bool ACC>
int launch_psi(fable_ctx* ctx, const PsiArgs& a, unsigned grid, int64_t t0) {
  constexpr size_t smem = OP_BYTES + (MAT ? STAGE_BYTES : 0);
  static bool attr = false;
  if (!attr) {
    FB_CUDA(ctx, cudaFuncSetAttribute(k_psi<MAT, ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr = true;
  }
  k_psi<MAT, ACC><<<grid, 256, smem, ctx->stream>>>(a, t0);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}

}  // namespace

int fb_psi(fable_ctx* ctx, const PsiArgs& a) {
  FB_REQUIRE(ctx, a.p > 0 && a.k > 0 && a.B > 0, "psi: empty problem");
  FB_REQUIRE(ctx, (a.acc_sum == nullptr) == (a.acc_sq == nullptr), "psi: need both accumulators");
  const bool mat = a.psi != nullptr, acc = a.acc_sum != nullptr;
  FB_REQUIRE(ctx, mat || acc, "psi: nothing to do");
  FB_REQUIRE(ctx, !mat || a.psi_slots > 0, "psi: psi_slots must be positive");
  FB_REQUIRE(ctx, a.pp % TILE == 0 && a.pp >= a.p && (reinterpret_cast<uintptr_t>(a.Lw) & 15) == 0,
             "psi: operand rows must be zero-padded to a multiple of 64 and 16-byte aligned");
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
    if (mat && acc) FB_TRY((launch_psi<true, true>(ctx, a, g, t0)));
    else if (mat) FB_TRY((launch_psi<true, false>(ctx, a, g, t0)));
    else FB_TRY((launch_psi<false, true>(ctx, a, g, t0)));
  }
  if (ev1) FB_CUDA(ctx, cudaEventRecord(ev1, ctx->stream));
  return FABLE_OK;
}

int fb_pad_rows(fable_ctx* ctx, const double* src, int64_t lds, int64_t p, int64_t pp, int k, double* dst) {
  k_pad_rows<<<static_cast<unsigned>(ceil_div(pp, 256)), 256, 0, ctx->stream>>>(src, lds, p, pp, k, dst);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}

int fb_mirror_upper(fable_ctx* ctx, double* dA, int64_t p) {
  const unsigned nT = static_cast<unsigned>(ceil_div(p, 32));
  k_mirror_upper<<<dim3(nT, nT), 256, 0, ctx->stream>>>(dA, p);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}
