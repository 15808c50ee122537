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
    for (int idx0 = threadIdx.x; idx0 < k * NI; idx0 += 256 * 8) {     // 8 independent loads in flight per thread
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int idx = idx0 + u * 256, l = idx / NI, i = idx - l * NI;
        v[u] = (idx < k * NI && i < ni) ? __ldg(U + i0 + i + ldu * l) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (idx0 + u * 256 < k * NI) Us[idx0 + u * 256] = v[u];
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


// ---------------------------------------------------------------------------------------------------
// k_rowstats_mma<KS, IPW>: the same pass with the fit on the FP64 tensor cores (n even, n <= 32 * 16 * IPW rows with
// KS * IPW <= 6; everything else takes k_rowstats above).
//
//   D' = Y_tile' + C_tile (-U_item)'          m8n8k4: M = 8 columns of Y, N = 8 rows of Y, K = 4 ranks
//
// In that orientation the accumulator fragment of lane (g, t) is column g, rows (2t, 2t + 1) of an 8 x 8 block:
// two doubles that are ADJACENT in column-major Y, so an accumulator is initialised from one 16-byte piece of Y
// (a warp instruction covers 8 columns x 64 contiguous bytes).  A = c (8 columns x 4 ranks), B = -U (4 ranks x
// 8 rows).
//
// Work split: one CTA of 16 warps per SM.  A tile is 8 columns x all n rows, an item 32 rows of a tile (2 KB of Y).
// G = min(16, items) warps share a tile, the CTA works on 16 / G tiles at a time, warp r of a group owns items r,
// r + G, ... (at most IPW of them) of EVERY tile its group visits -- so the -U fragments of its items never change
// and live in registers (KS * IPW * 4 doubles per lane, loaded once): no shared memory is spent on U_k, all of it
// goes to the prefetch ring.
// Prefetch: every lane copies ITS OWN accumulator fragments (16-byte cp.async, zero fill past n / p) and c fragments
// (8-byte cp.async) of the following items into a per-warp ring of S stages (S = what 227 KB allow, 5 at k = 10) laid
// out in fragment order -- a lane reads back exactly the slots it copied, 512 contiguous bytes per LDS.128, no
// cross-lane hand-over, no barrier -- and waits with cp.async.wait_group S - 1, which leaves the S - 1 newer items
// in flight: 128 KB per SM outstanding, re-issued item by item, so the memory pipe never drains.  (A ring of
// REGISTER buffers cannot do that here: the DMMAs take four of the six scoreboards, all Y loads land on the one
// that is left, and waiting for the oldest item waits for every load behind it -- measured: depth 3 in the source
// ran like depth 0.  And with U_k in shared memory only 3 stages fit: 64 KB in flight = latency-bound at 3.5 TB/s.)
// The per-column partial sums of a group's warps meet once per tile in a double-buffered shared array behind a
// NAMED barrier of that group only -- groups drift freely against each other -- and are added in warp order:
// bit-stable.
#ifdef ROWS_LAB_TIMING
__device__ int rows_lab_count = 0;
#endif
#ifndef ROWS_ISSUE_LATE
#define ROWS_ISSUE_LATE 1
#endif
constexpr int MW = 16;                 // warps per CTA
constexpr int ITEM = 32;               // rows per item (4 DMMA column blocks of the transposed product)
constexpr int MAXS = 8;                // ring stages per warp, at most

__device__ __forceinline__ void dmma_884(double2& c, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c.x), "+d"(c.y) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp16_zfill(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp8_zfill(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
template <int KS>
struct RowsMmaCfg {
  static constexpr int kStage = 2048 + KS * 256;                  // 4 x 512 B of Y fragments, KS x 256 B of c fragments
  static constexpr int kRed = 2 * MW * 16 * 8;
  static constexpr int kFit = (227 * 1024 - kRed) / (MW * kStage);
  static constexpr int S = kFit > MAXS ? MAXS : kFit;              // ring stages per warp (5 at k = 10)
};

// The loop is written for few instructions per item (the first version spent 275 per 2 KB item -- 64-bit index
// arithmetic and bounds predicates on every copy -- and was bound by the serial latency of that stream with four
// warps per scheduler, not by memory): per-lane source pointers advance by constants, validity is one flag per
// tile plus a 32-bit row compare, interior items (a warp-uniform test) are copied without any predicate, predicated-off
// copies of the ragged edges read 0 bytes from a valid address, the ring depth is a compile-time constant.
template <int KS, int IPW>
__global__ void __launch_bounds__(MW * 32, 1)
k_rowstats_mma(const double* __restrict__ Y, int n, int p, int64_t ldy, const double* __restrict__ U,
               int64_t ldu, const double* __restrict__ c, int64_t ldc, int k, int nitems, int G, int nsuper,
               double* __restrict__ s_out, double* __restrict__ resid_out) {
  constexpr int STB = RowsMmaCfg<KS>::kStage, S = RowsMmaCfg<KS>::S;
  extern __shared__ __align__(16) double smem[];
  double* red = smem;                                              // [2][MW][16]
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
  const int TC = MW / G, ts = w / G, wr = w - ts * G;              // tile slot of the warp's group, rank in the group
  if (ts >= TC) return;                                            // 16 % G warps have no group (no CTA barrier below)
  const uint32_t ring = static_cast<uint32_t>(__cvta_generic_to_shared(red + 2 * MW * 16)) + w * (S * STB) + lane * 16;
  const uint32_t ring_a = ring - lane * 8 + 2048;                  // c fragments: 8 bytes per lane
  const int sstep = gridDim.x, dcol = 8 * TC * sstep;

  // load cursor: S - 1 items ahead of the work cursor
  int lst = blockIdx.x, lit = wr, lstage = 0;
  int lcol = (lst * TC + ts) * 8 + g;
  bool lok = lst < nsuper && lcol < p;
  const double* ytile = Y + ldy * lcol + 2 * t;                    // never dereferenced while !lok
  const int64_t ytile_step = ldy * dcol;
  const double* cl = c + lcol;
  bool lfull = lst < nsuper && lcol - g + 8 <= p;                  // warp-uniform: all 8 columns of the tile exist
  auto issue = [&]() {
    const uint32_t sb = lstage * STB;
    const int r0 = lit * ITEM + 2 * t;
    const double* yp = ytile + lit * ITEM;
    const bool first = lit == wr;                        // first item of a tile: its c fragments ride along
    if (lfull && lit * ITEM + ITEM <= n) {               // interior item (warp-uniform): no predicates at all
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) cp16_zfill(ring + sb + nb * 512, yp + nb * 8, 16);
      if (first) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const int l = ks * 4 + t;
          cp8_zfill(ring_a + sb + ks * 256, cl + ldc * (l < k ? l : k - 1), l < k ? 8 : 0);
        }
      }
    } else {                                             // ragged rows / columns / past the last tile: predicated-off
#pragma unroll                                           // copies read 0 bytes from a valid address
      for (int nb = 0; nb < 4; ++nb) {
        const bool ok = lok && r0 + nb * 8 < n;
        cp16_zfill(ring + sb + nb * 512, ok ? yp + nb * 8 : Y, ok ? 16 : 0);
      }
      if (first) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const bool ok = lok && ks * 4 + t < k;
          cp8_zfill(ring_a + sb + ks * 256, ok ? cl + ldc * (ks * 4 + t) : c, ok ? 8 : 0);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    lstage = lstage == S - 1 ? 0 : lstage + 1;
    lit += G;
    if (lit >= nitems) {
      lit = wr; lst += sstep; lcol += dcol; ytile += ytile_step; cl += dcol;
      lok = lst < nsuper && lcol < p;
      lfull = lst < nsuper && lcol - g + 8 <= p;
    }
  };
#pragma unroll
  for (int d = 0; d < S - 1; ++d) issue();

  // -U fragments of the warp's own items: B[k = t][n = g] = -U[row g of the block][rank ks * 4 + t]
  double ub[IPW][KS][4];
#pragma unroll
  for (int q = 0; q < IPW; ++q)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) {
        const int row = (wr + q * G) * ITEM + nb * 8 + g, l = ks * 4 + t;
        ub[q][ks][nb] = (row < n && l < k) ? __ldg(U + row + ldu * l) : 0.0;
      }
#pragma unroll
  for (int q = 0; q < IPW; ++q)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) ub[q][ks][nb] = -ub[q][ks][nb];      // after ALL loads are in flight

  int buf = 0, stage = 0;
  const int bar_id = 1 + ts, bar_threads = G * 32;
#ifdef ROWS_LAB_TIMING
  long long lab_wait = 0, lab_math = 0, lab_bar = 0, lab_issue = 0;
  const long long lab_t0 = clock64();
#endif
  for (int st = blockIdx.x; st < nsuper; st += sstep) {
    double ar0 = 0.0, ar1 = 0.0, as0 = 0.0, as1 = 0.0;
    double a[KS];
#pragma unroll
    for (int q = 0; q < IPW; ++q) {
      if (q == 0 || wr + q * G < nitems) {               // warp-uniform
#ifdef ROWS_LAB_TIMING
        const long long lab_a = clock64();
#endif
#if !ROWS_ISSUE_LATE
        issue();                                         // item + S - 1 into the stage consumed by the previous step
#endif
#ifdef ROWS_LAB_TIMING
        const long long lab_b = clock64();
#endif
        asm volatile("cp.async.wait_group %0;" ::"n"(ROWS_ISSUE_LATE ? S - 2 : S - 1) : "memory");
#ifdef ROWS_LAB_TIMING
        const long long lab_c = clock64();
        lab_issue += lab_b - lab_a; lab_wait += lab_c - lab_b;
#endif
        double2 acc[4];
        const uint32_t sb = stage * STB;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb)
          asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(acc[nb].x), "=d"(acc[nb].y) : "r"(ring + sb + nb * 512) : "memory");
        if (q == 0) {
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a[ks]) : "r"(ring_a + sb + ks * 256) : "memory");
        }
        stage = stage == S - 1 ? 0 : stage + 1;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) { as0 = fma(acc[nb].x, acc[nb].x, as0); as1 = fma(acc[nb].y, acc[nb].y, as1); }
#ifndef ROWS_LAB_NOMATH               // switch-off lab (-DROWS_LAB_NOMATH): the stream, the ring and the sums without the DMMAs
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
          for (int nb = 0; nb < 4; ++nb) dmma_884(acc[nb], a[ks], ub[q][ks][nb]);
#else
        acc[0].x += a[0] * ub[q][0][0];
#endif
#if ROWS_ISSUE_LATE
        // the copies of item + S - 1 (into the stage read one step earlier) go out BEHIND the DMMAs: the four LDGSTS wait for
        // their turn in the memory pipe (300-560 cycles per item in the cycle counters) while the tensor pipe works through
        // the twelve DMMAs this warp has just queued, instead of in front of them
        issue();
#endif
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) { ar0 = fma(acc[nb].x, acc[nb].x, ar0); ar1 = fma(acc[nb].y, acc[nb].y, ar1); }
#ifdef ROWS_LAB_TIMING
        lab_math += clock64() - lab_c;
#endif
      }
    }
    // the warp's part of this tile is complete
    double r = ar0 + ar1, s = as0 + as1;
    r += __shfl_xor_sync(0xffffffffu, r, 1); s += __shfl_xor_sync(0xffffffffu, s, 1);
    r += __shfl_xor_sync(0xffffffffu, r, 2); s += __shfl_xor_sync(0xffffffffu, s, 2);
    const int col = (st * TC + ts) * 8;
    if (G == 1) {
      if (t == 0 && col + g < p) {
        resid_out[col + g] = r;
        if (s_out) s_out[col + g] = s;
      }
    } else {
      if (t == 0) {
        red[(buf * MW + w) * 16 + g] = r;
        red[(buf * MW + w) * 16 + 8 + g] = s;
      }
#ifdef ROWS_LAB_TIMING
      const long long lab_d = clock64();
#endif
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_threads) : "memory");
#ifdef ROWS_LAB_TIMING
      lab_bar += clock64() - lab_d;
#endif
      if (wr == 0 && lane < 16) {
        // the group's partial sums, fetched together and added as a fixed tree: the sixteen dependent LDS -> DADD steps of a
        // plain loop (~700 cycles, on the critical path of every tile: the other warps meet this one at the next barrier)
        // become one shared-memory latency and four additions
        double v[MW];
#pragma unroll
        for (int w2 = 0; w2 < MW; ++w2) v[w2] = w2 < G ? red[(buf * MW + w + w2) * 16 + lane] : 0.0;
#pragma unroll
        for (int h = 1; h < MW; h <<= 1)
#pragma unroll
          for (int w2 = 0; w2 + h < MW; w2 += 2 * h) v[w2] += v[w2 + h];
        const double sum = v[0];
        if (col + (lane & 7) < p) {
          if (lane < 8) resid_out[col + lane] = sum;
          else if (s_out) s_out[col + lane - 8] = sum;
        }
      }
      buf ^= 1;
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#ifdef ROWS_LAB_TIMING
  if ((blockIdx.x == 0 || blockIdx.x == 77) && lane == 0 && (w == 0 || w == 5 || w == 15) && atomicAdd(&rows_lab_count, 1) / 6 == 50)
    printf("ROWS_LAB cta %3d warp %2d: issue %lld wait %lld math %lld barrier %lld total %lld cycles\n", blockIdx.x, w, lab_issue, lab_wait, lab_math,
           lab_bar, clock64() - lab_t0);
#endif
}

// ---------------------------------------------------------------------------------------------------
// k_rowstats_tma<KS, IPW>: the same fit, the stream of Y moved by the TMA engine instead of by the computing warps.
// Cycle counters in k_rowstats_mma (-DROWS_LAB_TIMING, C3) show what bounds it: a warp never waits for its data (190 of
// 64 000 cycles in cp.async.wait_group) -- it spends 10-19 k cycles ISSUING its own copies (four LDGSTS per item behind the
// memory pipe's back-pressure) and 22-34 k in the math section, one after the other, so neither the memory pipe nor the
// FP64 pipe (44 % busy) is kept full by four warps per scheduler.  Here the TMA engine owns the stream: a stage is a whole
// tile group -- 8 (16 / G) columns x n rows, CONTIGUOUS in column-major Y (needs ldy == n) -- and arrives by ONE
// cp.async.bulk (global -> shared, completion counted in bytes on an mbarrier) issued by one lane of warp 0 two stages ahead
// (a first version with one copy per column and 512-row chunk spent 850 cycles per stage issuing them: a bulk copy costs
// its issuing thread ~100 cycles); the 16 warps wait on the stage's "full" barrier, read their fragments with one LDS.128
// each, release the stage on its "empty" barrier after their last read and otherwise do nothing but math.  Rows past n of
// a warp's last item belong to the next column and are masked by the reader; columns past p are not copied and their sums
// are not stored.  c fragments: plain loads, one tile ahead.
__device__ __forceinline__ uint32_t rows_smem_u32(const void* p_) { return static_cast<uint32_t>(__cvta_generic_to_shared(p_)); }
__device__ __forceinline__ void rows_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rows_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void rows_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rows_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rows_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(rows_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void rows_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred P1;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(rows_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void rows_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(rows_smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(rows_smem_u32(bar))
               : "memory");
}
constexpr int TMA_HDR = 2 * MW * 16 * 8 + 128;      // red[2][MW][16] + full[8] + empty[8]
constexpr int TMA_SLACK = 256;                      // masked reads past the last column of the last slot stay inside the allocation

template <int KS, int IPW>
__global__ void __launch_bounds__(MW * 32, 1)
k_rowstats_tma(const double* __restrict__ Y, int n, int p, const double* __restrict__ U, int64_t ldu, const double* __restrict__ c,
               int64_t ldc, int k, int nitems, int G, int nsuper, int S, int stage_bytes, double* __restrict__ s_out,
               double* __restrict__ resid_out) {
  extern __shared__ __align__(128) unsigned char rows_smem[];
  double* red = reinterpret_cast<double*>(rows_smem);                             // [2][MW][16]
  uint64_t* full = reinterpret_cast<uint64_t*>(rows_smem + 2 * MW * 16 * 8);
  uint64_t* empty = full + MAXS;
  unsigned char* stages = rows_smem + TMA_HDR;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int TC = MW / G, stage_cols = 8 * TC;
  if (tid == 0) {
    for (int i = 0; i < S; ++i) { rows_mbar_init(full + i, 1); rows_mbar_init(empty + i, MW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // ---- the stream: stage j (tile group blockIdx.x + j gridDim.x) lives in slot j % S and may be refilled once all 16
  // warps have released stage j - S; issued S - 1 stages ahead of the fit ----
  int pj = 0;
  auto produce = [&]() {
    const int64_t pst = blockIdx.x + static_cast<int64_t>(pj) * gridDim.x;
    if (pst >= nsuper) return;
    const int slot = pj % S;
    const int64_t col0 = pst * stage_cols;
    const int ncol = p - col0 < stage_cols ? static_cast<int>(p - col0) : stage_cols;
    if (lane == 0) {
      if (pj >= S) rows_mbar_wait(empty + slot, ((pj / S) - 1) & 1);
      const uint32_t bytes = static_cast<uint32_t>(ncol) * n * 8;
      rows_mbar_expect_tx(full + slot, bytes);
      rows_bulk_load(stages + static_cast<size_t>(slot) * stage_bytes, Y + col0 * n, bytes, full + slot);
    }
    ++pj;
  };
  if (w == 0)
    for (int i = 0; i < S - 1; ++i) produce();
  // ---- the fit ---------------------------------------------------------------------------------------
  const int g = lane >> 2, t = lane & 3;
  const int ts = w / G, wr = w - ts * G;
  const bool grouped = ts < TC;                              // 16 % G warps have no group: they only pass the barriers
  double ub[IPW][KS][4];
#pragma unroll
  for (int q = 0; q < IPW; ++q)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) {
        const int row = (wr + q * G) * ITEM + nb * 8 + g, l = ks * 4 + t;
        ub[q][ks][nb] = (grouped && row < n && l < k) ? __ldg(U + row + ldu * l) : 0.0;
      }
#pragma unroll
  for (int q = 0; q < IPW; ++q)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) ub[q][ks][nb] = -ub[q][ks][nb];
  auto load_c = [&](int64_t st, double (&a)[KS]) {
    const int64_t col = (st * TC + ts) * 8 + g;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int l = ks * 4 + t;
      a[ks] = (grouped && st < nsuper && col < p && l < k) ? __ldg(c + col + ldc * l) : 0.0;
    }
  };
  double a_next[KS];
  load_c(blockIdx.x, a_next);
  int buf = 0, use = 0;
  const int bar_id = 1 + ts, bar_threads = G * 32;
#ifdef ROWS_LAB_TIMING
  long long lab_prod = 0, lab_wait = 0, lab_math = 0, lab_bar = 0;
  const long long lab_t0 = clock64();
#endif
  const uint32_t lane_off = static_cast<uint32_t>(((ts * 8 + g) * n + wr * ITEM + 2 * t) * 8);
  const uint32_t stage0 = rows_smem_u32(stages);
  for (int64_t st = blockIdx.x; st < nsuper; st += gridDim.x, ++use) {
    double a[KS];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) a[ks] = a_next[ks];
    load_c(st + gridDim.x, a_next);
    double ar0 = 0.0, ar1 = 0.0, as0 = 0.0, as1 = 0.0;
    const int slot = use % S;
#ifdef ROWS_LAB_TIMING
    const long long lab_a = clock64();
#endif
    if (w == 0) produce();                                   // stage use + S - 1 into the slot released during stage use - 1
#ifdef ROWS_LAB_TIMING
    const long long lab_b = clock64();
#endif
    rows_mbar_wait(full + slot, (use / S) & 1);
#ifdef ROWS_LAB_TIMING
    const long long lab_c = clock64();
    lab_prod += lab_b - lab_a; lab_wait += lab_c - lab_b;
#endif
    const uint32_t src0 = stage0 + slot * stage_bytes + lane_off;
#pragma unroll
    for (int q = 0; q < IPW; ++q) {
      const int it = wr + q * G;
      double2 acc[4];
      const bool mine = grouped && it < nitems;              // warp-uniform
      if (mine) {
#pragma unroll
        for (int nb = 0; nb < 4; ++nb)
          asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(acc[nb].x), "=d"(acc[nb].y) : "r"(src0 + (q * G * ITEM + nb * 8) * 8) : "memory");
      }
      if (q == IPW - 1) {                                    // last read of this stage
        __syncwarp();
        if (lane == 0) rows_mbar_arrive(empty + slot);
      }
      if (mine) {
        if (it * ITEM + ITEM > n) {                          // last item of a column: rows past n belong to the next column
#pragma unroll
          for (int nb = 0; nb < 4; ++nb)
            if (it * ITEM + nb * 8 + 2 * t >= n) acc[nb] = make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) { as0 = fma(acc[nb].x, acc[nb].x, as0); as1 = fma(acc[nb].y, acc[nb].y, as1); }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
          for (int nb = 0; nb < 4; ++nb) dmma_884(acc[nb], a[ks], ub[q][ks][nb]);
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) { ar0 = fma(acc[nb].x, acc[nb].x, ar0); ar1 = fma(acc[nb].y, acc[nb].y, ar1); }
      }
    }
#ifdef ROWS_LAB_TIMING
    lab_math += clock64() - lab_c;
#endif
    if (!grouped) continue;
    double r = ar0 + ar1, s = as0 + as1;
    r += __shfl_xor_sync(0xffffffffu, r, 1); s += __shfl_xor_sync(0xffffffffu, s, 1);
    r += __shfl_xor_sync(0xffffffffu, r, 2); s += __shfl_xor_sync(0xffffffffu, s, 2);
    const int64_t col = (st * TC + ts) * 8;
    if (G == 1) {
      if (t == 0 && col + g < p) {
        resid_out[col + g] = r;
        if (s_out) s_out[col + g] = s;
      }
    } else {
      if (t == 0) {
        red[(buf * MW + w) * 16 + g] = r;
        red[(buf * MW + w) * 16 + 8 + g] = s;
      }
#ifdef ROWS_LAB_TIMING
      const long long lab_d = clock64();
#endif
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_threads) : "memory");
#ifdef ROWS_LAB_TIMING
      lab_bar += clock64() - lab_d;
#endif
      if (wr == 0 && lane < 16) {
        // the group's partial sums, fetched together and added as a fixed tree: the sixteen dependent LDS -> DADD steps of a
        // plain loop (~700 cycles, on the critical path of every tile: the other warps meet this one at the next barrier)
        // become one shared-memory latency and four additions
        double v[MW];
#pragma unroll
        for (int w2 = 0; w2 < MW; ++w2) v[w2] = w2 < G ? red[(buf * MW + w + w2) * 16 + lane] : 0.0;
#pragma unroll
        for (int h = 1; h < MW; h <<= 1)
#pragma unroll
          for (int w2 = 0; w2 + h < MW; w2 += 2 * h) v[w2] += v[w2 + h];
        const double sum = v[0];
        if (col + (lane & 7) < p) {
          if (lane < 8) resid_out[col + lane] = sum;
          else if (s_out) s_out[col + lane - 8] = sum;
        }
      }
      buf ^= 1;
    }
  }
#ifdef ROWS_LAB_TIMING
  if ((blockIdx.x == 0 || blockIdx.x == 77) && lane == 0 && (w == 0 || w == 5 || w == 15) && atomicAdd(&rows_lab_count, 1) / 6 == 50)
    printf("ROWS_LAB(tma) cta %3d warp %2d: produce %lld wait %lld math %lld barrier %lld total %lld cycles\n", blockIdx.x, w, lab_prod, lab_wait,
           lab_math, lab_bar, clock64() - lab_t0);
#endif
}

template <int KS, int IPW>
int launch_rowstats_tma(fable_ctx* ctx, const double* dY, int64_t n, int64_t p, int64_t ldy, const double* dU, int64_t ldu,
                        const double* d_c, int64_t ldc, int k, int nitems, double* d_s, double* d_resid, bool* launched) {
  const int G = nitems < MW ? nitems : MW;
  const int64_t stage_bytes = round_up(static_cast<int64_t>(8) * (MW / G) * n * 8, 128);
  int64_t S = (227 * 1024 - TMA_HDR - TMA_SLACK) / stage_bytes;
  if (S > MAXS) S = MAXS;
  *launched = false;
  if (S < 3 || ldy != n) return FABLE_OK;                    // (the tile must be contiguous; longer columns: k_rowstats_mma)
  const size_t smem = TMA_HDR + TMA_SLACK + static_cast<size_t>(S) * stage_bytes;
  static bool attr_dev[64] = {};
  bool& attr_set = attr_dev[ctx->device & 63];
  if (!attr_set) {
    FB_CUDA(ctx, cudaFuncSetAttribute(k_rowstats_tma<KS, IPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int64_t nsuper = ceil_div(ceil_div(p, 8), MW / G);
  const unsigned grid = static_cast<unsigned>(nsuper < ctx->sms ? nsuper : ctx->sms);
  k_rowstats_tma<KS, IPW><<<grid, MW * 32, smem, ctx->stream>>>(dY, static_cast<int>(n), static_cast<int>(p), dU, ldu, d_c, ldc, k, nitems, G,
                                                                static_cast<int>(nsuper), static_cast<int>(S), static_cast<int>(stage_bytes), d_s, d_resid);
  FB_LAUNCHED(ctx);
  *launched = true;
  return FABLE_OK;
}

template <int KS, int IPW>
int launch_rowstats_mma(fable_ctx* ctx, const double* dY, int64_t n, int64_t p, int64_t ldy, const double* dU, int64_t ldu,
                        const double* d_c, int64_t ldc, int k, int nitems, double* d_s, double* d_resid) {
  using Cfg = RowsMmaCfg<KS>;
  const size_t smem = Cfg::kRed + static_cast<size_t>(Cfg::S) * MW * Cfg::kStage;
  static bool attr_dev[64] = {};
  bool& attr_set = attr_dev[ctx->device & 63];
  if (!attr_set) {
    FB_CUDA(ctx, cudaFuncSetAttribute(k_rowstats_mma<KS, IPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_set = true;
  }
  const int G = nitems < MW ? nitems : MW;
  const int64_t nsuper = ceil_div(ceil_div(p, 8), MW / G);
  const unsigned grid = static_cast<unsigned>(nsuper < ctx->sms ? nsuper : ctx->sms);
  k_rowstats_mma<KS, IPW><<<grid, MW * 32, smem, ctx->stream>>>(
      dY, static_cast<int>(n), static_cast<int>(p), ldy, dU, ldu, d_c, ldc, k, nitems, G, static_cast<int>(nsuper), d_s, d_resid);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
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
  // tensor-core pass: needs 16-byte aligned row pairs and the warps' -U fragments in registers (KS * IPW <= 6)
  const char* env_mma = getenv("FABLE_B200_ROWS_MMA");       // 0: force the DFMA kernel (A/B runs, tests)
  const int use_mma = env_mma ? atoi(env_mma) : 1;
  if (use_mma && n % 2 == 0 && ldy % 2 == 0 && reinterpret_cast<uintptr_t>(dY) % 16 == 0 && n < (1 << 30) && p < (1 << 27)) {
    const int KS = (k + 3) / 4;
    const int64_t nitems = ceil_div(n, ITEM);
    const int64_t IPW = ceil_div(nitems, MW);
    const int ni = static_cast<int>(nitems);
    // TMA-streamed form where it measured faster: ranks up to 4 (C3 shape at k = 4: 28.9 vs 30.8 us = 0.85 of HBM); at
    // k = 10 the two are level (39.7 vs 38.9 us), at n = 500, k = 20 and n = 200 the cp.async ring wins (74 vs 82, 29 vs 33 us).
    // FABLE_B200_ROWS_TMA=1 / 0 forces / forbids it.
    const char* env_tma = getenv("FABLE_B200_ROWS_TMA");
    const bool use_tma = env_tma ? atoi(env_tma) != 0 : k <= 4;
#define FB_ROWS_MMA(KS_, IPW_) \
    if (KS == KS_ && IPW == IPW_) { \
      if (use_tma) { \
        bool launched = false; \
        FB_TRY((launch_rowstats_tma<KS_, IPW_>(ctx, dY, n, p, ldy, dU, ldu, d_c, ldc, k, ni, d_s, d_resid, &launched))); \
        if (launched) return FABLE_OK; \
      } \
      return launch_rowstats_mma<KS_, IPW_>(ctx, dY, n, p, ldy, dU, ldu, d_c, ldc, k, ni, d_s, d_resid); \
    }
    FB_ROWS_MMA(1, 1) FB_ROWS_MMA(2, 1) FB_ROWS_MMA(3, 1) FB_ROWS_MMA(4, 1) FB_ROWS_MMA(5, 1) FB_ROWS_MMA(6, 1)
    FB_ROWS_MMA(1, 2) FB_ROWS_MMA(2, 2) FB_ROWS_MMA(3, 2)
    FB_ROWS_MMA(1, 3) FB_ROWS_MMA(2, 3)
    FB_ROWS_MMA(1, 4) FB_ROWS_MMA(1, 5) FB_ROWS_MMA(1, 6)
#undef FB_ROWS_MMA
  }
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
