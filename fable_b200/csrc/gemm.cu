// FP64 tensor-core GEMM (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4 on sm_100a; tcgen05 has no f64
// kind) used for every dense contraction of the path:
//   * Gram step  Y Y' (n <= p) or Y'Y (p < n)   -- replaces the dgesdd inside base-R svd(Y)
//                                                   (reference R/FABLE_wrapper.R:23, 90)
//   * V = Y'U D^-1 / U = Y V D^-1 back-multiplication
//   * the rank-k' reconstruction + residual of CPPCriterionFunction
//     (src/updated-FABLE-functions.cpp:161-166): fused epilogue, U D V' is never stored.
//
// C (M x N) = alpha * A B',  A: M x K, B: N x K.  Each operand is either "K-slow" (column-major
// M x K: element (m,kk) at m + ld*kk) or "K-fast" (element (m,kk) at kk + ld*m).
// CTA tile 128 x 128 x 16, 8 warps as 2 (M) x 4 (N), warp tile 64 x 32 = 8 x 4 DMMA tiles.
// The Gram Y Y' of the K-slow layout has its own TMA-staged kernel (k_syrk_tma, below); k_gemm stages through registers:
// Shared tiles are [kk][m] with row stride 132 doubles (132 mod 16 = 4): the 16 lanes of a
// half-warp (g = 0..3, t = 0..3) read bank 4t+g -> conflict-free 8-byte fragment loads.
#include <cuda.h>   // CUtensorMap types only; the encoder is fetched through cudaGetDriverEntryPoint

#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int LDS_ = BM + 4;
constexpr int STAGE_DOUBLES = BK * LDS_;

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <bool KFAST>
__device__ __forceinline__ void load_tile_regs(const double* __restrict__ X, int64_t ld, int64_t rows,
                                               int64_t row0, int64_t k0, int64_t kend, double (&r)[8]) {
  if (!KFAST) {
    const int m = threadIdx.x & 127, kb = threadIdx.x >> 7;
    const int64_t gm = row0 + m;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t gk = k0 + kb + 2 * i;
      r[i] = (gm < rows && gk < kend) ? X[gm + ld * gk] : 0.0;
    }
  } else {
    const int kk = threadIdx.x & 15, mb = threadIdx.x >> 4;
    const int64_t gk = k0 + kk;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t gm = row0 + mb + 16 * i;
      r[i] = (gm < rows && gk < kend) ? X[gk + ld * gm] : 0.0;
    }
  }
}

template <bool KFAST>
__device__ __forceinline__ void store_tile_smem(double* __restrict__ S, const double (&r)[8]) {
  if (!KFAST) {
    const int m = threadIdx.x & 127, kb = threadIdx.x >> 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) S[(kb + 2 * i) * LDS_ + m] = r[i];
  } else {
    const int kk = threadIdx.x & 15, mb = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) S[kk * LDS_ + mb + 16 * i] = r[i];
  }
}

// RESID = false: C[z] = alpha * partial product (z = split index, partial stride M*N when split)
// RESID = true : out[blockIdx.x * N + n] = sum_{m in tile} (Yres[m + ldy*n] - acc)^2
// SYRK: B == A and only the tiles with block row <= block column are computed (blockIdx.x enumerates
// the upper triangle of the tile grid column-major); the caller mirrors the result.
template <bool AKF, bool BKF, bool RESID, bool SYRK = false>
__global__ void __launch_bounds__(256)
k_gemm(int64_t M, int64_t N, int64_t K, int64_t kper, double alpha, const double* __restrict__ A, int64_t lda,
       const double* __restrict__ B, int64_t ldb, double* __restrict__ C, int64_t ldc, int64_t split_stride,
       const double* __restrict__ Yres, int64_t ldy) {
  extern __shared__ double smem[];
  double* As = smem;                      // [2][BK][LDS_]
  double* Bs = smem + 2 * STAGE_DOUBLES;  // [2][BK][LDS_]

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 1, wn = warp >> 1;
  int64_t bi = blockIdx.x, bj = blockIdx.y;
  if (SYRK) {
    const int64_t t = blockIdx.x;
    int64_t j = static_cast<int64_t>((sqrt(8.0 * static_cast<double>(t) + 1.0) - 1.0) * 0.5);
    while (j * (j + 1) / 2 > t) --j;
    while ((j + 1) * (j + 2) / 2 <= t) ++j;
    bj = j; bi = t - j * (j + 1) / 2;
  }
  const int64_t m0 = bi * BM, n0 = bj * BN;
  const int64_t kbeg = static_cast<int64_t>(blockIdx.z) * kper;
  const int64_t kend = (kbeg + kper < K) ? kbeg + kper : K;

  double c[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { c[i][j][0] = 0.0; c[i][j][1] = 0.0; }

  double ra[8], rb[8];
  load_tile_regs<AKF>(A, lda, M, m0, kbeg, kend, ra);
  load_tile_regs<BKF>(B, ldb, N, n0, kbeg, kend, rb);
  store_tile_smem<AKF>(As, ra);
  store_tile_smem<BKF>(Bs, rb);
  __syncthreads();

  int buf = 0;
  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = (k0 + BK < kend);
    if (more) {
      load_tile_regs<AKF>(A, lda, M, m0, k0 + BK, kend, ra);
      load_tile_regs<BKF>(B, ldb, N, n0, k0 + BK, kend, rb);
    }
    const double* as = As + buf * STAGE_DOUBLES + wm * 64 + g;
    const double* bs = Bs + buf * STAGE_DOUBLES + wn * 32 + g;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ++ks) {
      double af[8], bf[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) af[i] = as[(ks * 4 + t) * LDS_ + i * 8];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = bs[(ks * 4 + t) * LDS_ + j * 8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(c[i][j], af[i], bf[j]);
    }
    if (more) {
      store_tile_smem<AKF>(As + (buf ^ 1) * STAGE_DOUBLES, ra);
      store_tile_smem<BKF>(Bs + (buf ^ 1) * STAGE_DOUBLES, rb);
      __syncthreads();
      buf ^= 1;
    }
  }

  if (!RESID) {
    double* Cz = C + static_cast<int64_t>(blockIdx.z) * split_stride;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int64_t n = n0 + wn * 32 + j * 8 + 2 * t + e;
        if (n >= N) continue;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t m = m0 + wm * 64 + i * 8 + g;
          if (m < M) Cz[m + ldc * n] = alpha * c[i][j][e];
        }
      }
    }
  } else {
    __syncthreads();          // everyone is done with the operand tiles: reuse smem for the reduce
    double* red = smem;       // [2 (wm)][BN]
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int nl = wn * 32 + j * 8 + 2 * t + e;
        const int64_t n = n0 + nl;
        double acc = 0.0;
        if (n < N) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int64_t m = m0 + wm * 64 + i * 8 + g;
            if (m < M) {
              const double r = Yres[m + ldy * n] - c[i][j][e];
              acc = fma(r, r, acc);
            }
          }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 8);
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        if (g == 0) red[wm * BN + nl] = acc;
      }
    }
    __syncthreads();
    if (threadIdx.x < BN) {
      const int64_t n = n0 + threadIdx.x;
      if (n < N) C[static_cast<int64_t>(blockIdx.x) * N + n] = red[threadIdx.x] + red[BN + threadIdx.x];
    }
  }
}

constexpr size_t kSmemBytes = 4 * STAGE_DOUBLES * sizeof(double);

// ---- Gram step with TMA-staged tiles -------------------------------------------------------------------
// C (upper tiles) = alpha * A A' for a K-slow A (column-major M x K: the Gram Y Y' of an n x p data matrix, n <= p,
// contraction over the long axis).  Same CTA tile, warp layout and accumulator mapping as k_gemm, but the operand
// tiles are brought in by the TMA unit instead of through registers:
//   * A is described to TMA as a 2-D tensor (m, k); a 128 x 16 operand tile is eight {16 m, 16 k} boxes with the
//     128-byte swizzle, so the shared image of a box is 16 rows (k) of 128 bytes (16 m) with 16-byte chunk c of row k
//     stored at c ^ (k & 7).  Rows past M and columns past K are zero-filled by the hardware (no edge code).
//   * 4 stages of (A tile, B tile) = 128 KB; full[] mbarriers carry the transaction bytes, empty[] mbarriers are
//     arrived on once per warp, so the warps never meet at a CTA barrier inside the K loop; one thread issues the
//     loads three k-tiles ahead.
//   * dense 128-byte rows would make the four k's of a DMMA hit the same banks; the swizzle only separates rows
//     that differ in k & 6, so DMMA number ks of a k-tile takes k = {0,2,4,6} + (ks & 1) + 8 (ks >> 1) instead of
//     {4ks .. 4ks+3} (the sum over a k-tile does not care) and every half-warp fragment load touches 16 distinct
//     8-byte slots of one 128-byte row set: conflict-free.
constexpr int TS = 4;                                        // pipeline stages
constexpr int TBOX = 16 * 16;                                // doubles per {16, 16} box
constexpr int TTILE = (BM / 16) * TBOX;                      // one 128 x 16 operand tile: 2048 doubles = 16 KB
constexpr size_t kTmaSmemBytes = static_cast<size_t>(TS) * 2 * TTILE * sizeof(double) + 1024 + 2 * TS * sizeof(uint64_t);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred P1;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_box(const CUtensorMap* tmap, int m, int k, double* sdst, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(sdst)),
               "l"(reinterpret_cast<uint64_t>(tmap)), "r"(m), "r"(k), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(256)
k_syrk_tma(const __grid_constant__ CUtensorMap tmap, int64_t M, int64_t K, int64_t kper, double alpha, double* __restrict__ C,
           int64_t ldc, int64_t split_stride) {
  extern __shared__ unsigned char smem_raw_[];
  unsigned char* base = smem_raw_ + ((1024u - (smem_u32(smem_raw_) & 1023u)) & 1023u);   // swizzle atoms: 1 KB aligned
  double* tiles = reinterpret_cast<double*>(base);                                        // [TS][2][TTILE]
  uint64_t* full = reinterpret_cast<uint64_t*>(base + static_cast<size_t>(TS) * 2 * TTILE * sizeof(double));
  uint64_t* empty = full + TS;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 1, wn = warp >> 1;
  int64_t bi, bj;
  {
    const int64_t tt = blockIdx.x;
    int64_t j = static_cast<int64_t>((sqrt(8.0 * static_cast<double>(tt) + 1.0) - 1.0) * 0.5);
    while (j * (j + 1) / 2 > tt) --j;
    while ((j + 1) * (j + 2) / 2 <= tt) ++j;
    bj = j; bi = tt - j * (j + 1) / 2;
  }
  const bool diag = (bi == bj);                       // B tile == A tile: fetched once
  const int m0 = static_cast<int>(bi * BM), n0 = static_cast<int>(bj * BN);
  const int64_t kbeg = static_cast<int64_t>(blockIdx.z) * kper;
  const int64_t kend = (kbeg + kper < K) ? kbeg + kper : K;
  const int nk = static_cast<int>((kend - kbeg + BK - 1) / BK);
  const uint32_t stage_bytes = (diag ? 1u : 2u) * TTILE * sizeof(double);

  if (tid == 0) {
    for (int s = 0; s < TS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto produce = [&](int it) {          // thread 0 only
    const int s = it % TS;
    if (it >= TS) mbar_wait(&empty[s], ((it / TS) - 1) & 1);
    mbar_expect_tx(&full[s], stage_bytes);
    const int k0 = static_cast<int>(kbeg) + it * BK;
    double* a = tiles + static_cast<size_t>(s) * 2 * TTILE;
#pragma unroll
    for (int q = 0; q < BM / 16; ++q) tma_load_box(&tmap, m0 + q * 16, k0, a + q * TBOX, &full[s]);
    if (!diag) {
#pragma unroll
      for (int q = 0; q < BN / 16; ++q) tma_load_box(&tmap, n0 + q * 16, k0, a + TTILE + q * TBOX, &full[s]);
    }
  };
  if (tid == 0)
    for (int it = 0; it < TS - 1 && it < nk; ++it) produce(it);

  double c[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { c[i][j][0] = 0.0; c[i][j][1] = 0.0; }

  // fragment offsets inside a tile: element (m, k) at (m >> 4) * TBOX + k * 16 + ((((m & 15) >> 1) ^ (k & 7)) << 1) + (m & 1)
  int aoff[8], boff[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) aoff[i] = (wm * 4 + (i >> 1)) * TBOX + (g & 1);
#pragma unroll
  for (int j = 0; j < 4; ++j) boff[j] = (wn * 2 + (j >> 1)) * TBOX + (g & 1);
  const int cha = g >> 1;               // chunk of m within its 16-block: (i & 1) * 4 + (g >> 1)

  for (int it = 0; it < nk; ++it) {
    if (tid == 0 && it + TS - 1 < nk) produce(it + TS - 1);
    const int s = it % TS;
    mbar_wait(&full[s], (it / TS) & 1);
    const double* as = tiles + static_cast<size_t>(s) * 2 * TTILE;
    const double* bs = diag ? as : as + TTILE;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ++ks) {
      const int kk = 2 * t + (ks & 1) + 8 * (ks >> 1);      // this lane's k of DMMA number ks
      const int x = kk & 7, krow = kk * 16;
      double af[8], bf[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) af[i] = as[aoff[i] + krow + (((((i & 1) << 2) | cha) ^ x) << 1)];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = bs[boff[j] + krow + (((((j & 1) << 2) | cha) ^ x) << 1)];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(c[i][j], af[i], bf[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  double* Cz = C + static_cast<int64_t>(blockIdx.z) * split_stride;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int64_t n = n0 + wn * 32 + j * 8 + 2 * t + e;
      if (n >= M) continue;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + wm * 64 + i * 8 + g;
        if (m < M) Cz[m + ldc * n] = alpha * c[i][j][e];
      }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D map of a column-major M x K matrix (element (m, k) at m + lda k), {16, 16} boxes, 128-byte swizzle, zero fill
int make_operand_map(fable_ctx* ctx, const double* A, int64_t M, int64_t K, int64_t lda, CUtensorMap* out) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    FB_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return ctx->fail(FABLE_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(M), static_cast<cuuint64_t>(K)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(lda) * 8};
  const cuuint32_t box[2] = {16, 16}, estr[2] = {1, 1};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(A), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return ctx->fail(FABLE_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string(r));
  return FABLE_OK;
}

template <bool AKF, bool BKF, bool RESID, bool SYRK = false>
int launch(fable_ctx* ctx, dim3 grid, int64_t M, int64_t N, int64_t K, int64_t kper, double alpha, const double* A,
           int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t split_stride,
           const double* Yres, int64_t ldy) {
  static bool attr_dev[64] = {};
  bool& attr = attr_dev[ctx->device & 63];
  if (!attr) {
    FB_CUDA(ctx, cudaFuncSetAttribute(k_gemm<AKF, BKF, RESID, SYRK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(kSmemBytes)));
    attr = true;
  }
  k_gemm<AKF, BKF, RESID, SYRK><<<grid, 256, kSmemBytes, ctx->stream>>>(M, N, K, kper, alpha, A, lda, B, ldb, C, ldc,
                                                                 split_stride, Yres, ldy);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}

}  // namespace

namespace {
// Split K so that tiles x splits fills whole waves of the SMs (one 128 x 128 CTA per SM at 248 registers):
// the smallest number of waves whose last wave is >= 90% full wins (36 tiles: 4 splits = 144 CTAs = one
// wave, not 9 splits = 324 CTAs = 2.2 waves), subject to >= 4 k-tiles per split.
int64_t choose_splits(const fable_ctx* ctx, int64_t tiles, int64_t K, int64_t* kper) {
  const int64_t ktiles = ceil_div(K, BK), sms = ctx->sms;
  int64_t splits = 1;
  if (tiles < sms) {
    const int64_t cap = ktiles / 4 > 0 ? (ktiles / 4 < 64 ? ktiles / 4 : 64) : 1;
    double best_eff = 0.0;
    for (int64_t w = 1; w <= 3; ++w) {
      int64_t s = w * sms / tiles;
      if (s > cap) s = cap;
      if (s < 1) continue;
      const double eff = static_cast<double>(tiles * s) / static_cast<double>(ceil_div(tiles * s, sms) * sms);
      if (eff > best_eff + 1e-9) { best_eff = eff; splits = s; }
      if (eff >= 0.9) break;
    }
  } else if (tiles < 16 * sms && ktiles >= 64) {
    // a few waves with a poorly filled last one (820 tiles on 148 SMs: 5.5 waves = 92 %): splitting K by a small
    // factor evens the tail out (x 3: 16.6 waves = 98 %) for one extra pass over the partial sums
    double best_eff = static_cast<double>(tiles) / static_cast<double>(ceil_div(tiles, sms) * sms);
    if (best_eff < 0.97)
      for (int64_t s2 = 2; s2 <= 6 && ktiles / s2 >= 32; ++s2) {
        const double eff = static_cast<double>(tiles * s2) / static_cast<double>(ceil_div(tiles * s2, sms) * sms);
        if (eff > best_eff + 0.02) { best_eff = eff; splits = s2; }
        if (eff >= 0.97) break;
      }
  }
  if (const char* e = getenv("FABLE_B200_SPLITS")) splits = atoi(e);
  *kper = round_up(ceil_div(K, splits), BK);
  return ceil_div(K, *kper);
}
// grow-only split-K workspace owned by the context (no allocation or synchronisation per call)
int workspace(fable_ctx* ctx, size_t bytes, double** out) {
  if (bytes > ctx->gemm_ws_bytes) {
    FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->gemm_ws) cudaFree(ctx->gemm_ws);
    ctx->gemm_ws = nullptr; ctx->gemm_ws_bytes = 0;
    FB_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->gemm_ws), bytes));
    ctx->gemm_ws_bytes = bytes;
  }
  *out = ctx->gemm_ws;
  return FABLE_OK;
}
}  // namespace

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

static int fb_mirror_upper(fable_ctx* ctx, double* dA, int64_t p) {
  const unsigned nT = static_cast<unsigned>(ceil_div(p, 32));
  k_mirror_upper<<<dim3(nT, nT), 256, 0, ctx->stream>>>(dA, p);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}

// C (M x M, upper triangle + mirrored) = alpha * A A',  A: M x K (K-slow or K-fast)
int fb_syrk(fable_ctx* ctx, int a_kfast, int64_t M, int64_t K, double alpha, const double* dA, int64_t lda,
            double* dC, int64_t ldc) {
  FB_REQUIRE(ctx, M > 0 && K > 0 && ldc == M, "syrk: need M, K > 0 and a dense output");
  const int64_t gm = ceil_div(M, BM), tiles = gm * (gm + 1) / 2;
  int64_t kper = 0;
  const int64_t splits = choose_splits(ctx, tiles, K, &kper);
  dim3 grid(static_cast<unsigned>(tiles), 1, static_cast<unsigned>(splits));
  double* out = dC;
  int64_t stride = 0;
  double a_eff = alpha;
  if (splits > 1) {
    FB_TRY(workspace(ctx, static_cast<size_t>(splits) * M * M * sizeof(double), &out));
    stride = M * M; a_eff = 1.0;
    // lower tiles of the partial buffers are never written: zero them once so the sum is defined
    FB_CUDA(ctx, cudaMemsetAsync(out, 0, static_cast<size_t>(splits) * M * M * sizeof(double), ctx->stream));
  }
  // TMA-staged kernel for the K-slow Gram (rows 16-byte aligned: even leading dimension); FABLE_B200_SYRK_TMA=0
  // forces the register-staged kernel (A/B measurements)
  static const bool tma_ok = !(getenv("FABLE_B200_SYRK_TMA") && atoi(getenv("FABLE_B200_SYRK_TMA")) == 0);
  if (!a_kfast && tma_ok && (lda & 1) == 0 && (reinterpret_cast<uintptr_t>(dA) & 15) == 0 && M < (1ll << 31) && K < (1ll << 31)) {
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    FB_TRY(make_operand_map(ctx, dA, M, K, lda, &tmap));
    static bool attr_dev[64] = {};
    bool& attr = attr_dev[ctx->device & 63];
    if (!attr) {
      FB_CUDA(ctx, cudaFuncSetAttribute(k_syrk_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kTmaSmemBytes)));
      attr = true;
    }
    k_syrk_tma<<<grid, 256, kTmaSmemBytes, ctx->stream>>>(tmap, M, K, kper, a_eff, out, M, stride);
    FB_LAUNCHED(ctx);
  } else if (a_kfast) FB_TRY((launch<true, true, false, true>(ctx, grid, M, M, K, kper, a_eff, dA, lda, dA, lda, out, M, stride, nullptr, 0)));
  else FB_TRY((launch<false, false, false, true>(ctx, grid, M, M, K, kper, a_eff, dA, lda, dA, lda, out, M, stride, nullptr, 0)));
  if (splits > 1) FB_TRY(fb_sum_partials(ctx, out, splits, M * M, M * M, alpha, dC));
  return fb_mirror_upper(ctx, dC, M);
}

// Register-resident issue peak of the FP64 tensor path (mma.sync.m8n8k4.f64 -> DMMA.8x8x4): the denominator of the Gram
// step's roofline, measured on the device the bench runs on (MEASURED_PEAKS.json has no FP64 entry).  8 independent
// accumulator pairs per warp, 8 warps per CTA, 8 CTAs per SM; 512 flops per warp-level MMA.
__global__ void __launch_bounds__(256) k_dmma_peak(double* out, int iters, double a0, double b0) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
  const double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma(c[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int fb_dmma_peak(fable_ctx* ctx, double* tflops) {
  const int ctas = ctx->sms * 8, iters = 4096;
  DevBuf out;
  FB_CUDA(ctx, out.alloc(static_cast<size_t>(ctas) * 256 * sizeof(double)));
  cudaEvent_t e0, e1;
  FB_CUDA(ctx, cudaEventCreate(&e0));
  FB_CUDA(ctx, cudaEventCreate(&e1));
  float best = 0.f;
  for (int rep = 0; rep < 6; ++rep) {                 // first launches warm the clocks up; best of the rest
    cudaEventRecord(e0, ctx->stream);
    k_dmma_peak<<<ctas, 256, 0, ctx->stream>>>(out.d(), iters, 1.0, 1e-3);
    cudaEventRecord(e1, ctx->stream);
    ++ctx->launches;
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) { cudaEventDestroy(e0); cudaEventDestroy(e1); FB_CUDA(ctx, e); }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep >= 2 && (best == 0.f || ms < best)) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *tflops = static_cast<double>(ctas) * 8 /*warps*/ * iters * 8 /*mma*/ * 512.0 / (best * 1e-3) / 1e12;
  return FABLE_OK;
}

int fb_gemm(fable_ctx* ctx, int a_kfast, int b_kfast, int64_t M, int64_t N, int64_t K, double alpha,
            const double* dA, int64_t lda, const double* dB, int64_t ldb, double* dC, int64_t ldc) {
  FB_REQUIRE(ctx, M > 0 && N > 0 && K > 0, "gemm: empty problem");
  const int64_t gm = ceil_div(M, BM), gn = ceil_div(N, BN);
  FB_REQUIRE(ctx, gn <= 65535, "gemm: N too large for one launch");
  int64_t kper = 0;
  const int64_t splits = choose_splits(ctx, gm * gn, K, &kper);
  dim3 grid(static_cast<unsigned>(gm), static_cast<unsigned>(gn), static_cast<unsigned>(splits));
  double* out = dC;
  int64_t out_ld = ldc, stride = 0;
  double a_eff = alpha;
  if (splits > 1) {
    FB_TRY(workspace(ctx, static_cast<size_t>(splits) * M * N * sizeof(double), &out));
    out_ld = M; stride = M * N; a_eff = 1.0;
  }
#define FB_GEMM_CASE(AK, BKF_)                                                                                \
  if (a_kfast == AK && b_kfast == BKF_)                                                                       \
    FB_TRY((launch<AK, BKF_, false>(ctx, grid, M, N, K, kper, a_eff, dA, lda, dB, ldb, out, out_ld, stride,   \
                                    nullptr, 0)));
  FB_GEMM_CASE(0, 0) FB_GEMM_CASE(0, 1) FB_GEMM_CASE(1, 0) FB_GEMM_CASE(1, 1)
#undef FB_GEMM_CASE
  if (splits > 1) {
    if (ldc == M) {
      FB_TRY(fb_sum_partials(ctx, out, splits, M * N, M * N, alpha, dC));
    } else {
      for (int64_t n = 0; n < N; ++n)  // rare: strided output
        FB_TRY(fb_sum_partials(ctx, out + n * M, splits, M, M * N, alpha, dC + n * ldc));
    }
  }
  return FABLE_OK;
}

int fb_gemm_resid(fable_ctx* ctx, int64_t n, int64_t p, int k, const double* dY, int64_t ldy, const double* dU,
                  int64_t ldu, const double* dC, int64_t ldc, double* d_resid) {
  const int64_t gm = ceil_div(n, BM), gn = ceil_div(p, BN);
  FB_REQUIRE(ctx, gn <= 65535, "gemm_resid: p too large for one launch");
  double* part = nullptr;
  FB_TRY(workspace(ctx, static_cast<size_t>(gm) * p * sizeof(double), &part));
  dim3 grid(static_cast<unsigned>(gm), static_cast<unsigned>(gn), 1);
  FB_TRY((launch<false, false, true>(ctx, grid, n, p, k, round_up(k, BK), 1.0, dU, ldu, dC, ldc, part, 0, 0,
                                     dY, ldy)));
  return fb_sum_partials(ctx, part, gm, p, p, 1.0, d_resid);
}
