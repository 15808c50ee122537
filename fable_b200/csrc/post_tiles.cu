// Tile-resident posterior summaries of Psi: entry-wise mean and exact (alpha/2, 1 - alpha/2) sample quantiles over the
// Monte-Carlo draws for whole 64 x 32 blocks of entries at a time -- CPPCCFABLEPostProcessing (reference
// src/updated-FABLE-functions.cpp:631-713) and CCFABLEPostProcessingSubmatrix(_Optimized) (:716-893) at full-matrix scale.
//
// The reference loops over the p(p+1)/2 pairs and, per pair, forms the MC values sum_l Lam_m[j1,l] Lam_m[j2,l]
// (+ sigma2_m[j1] when the two indices are equal, :672-684), takes their mean (:688) and two Armadillo `quantile`s
// (:692-693).  Storing the values is out of the question at scale (p^2 MC / 2 doubles), so the values are RECOMPUTED from
// the draws -- they cost 2k flops each on the FP64 tensor cores -- and the order statistics are found by an exact
// most-significant-digit radix selection over several passes:
//   * a CTA of 16 warps owns 64 x 32 entries (4 per thread, in the DMMA accumulator layout of the covariance kernel); per pass it
//     streams its 96 rows of every draw from HBM/L2 (cp.async, double-buffered), forms the block with DMMA and folds the
//     values into per-entry state that never leaves the SM;
//   * pass A: sum (-> mean), minimum and maximum of every entry.  They define a per-entry order-preserving 52-bit key
//     K(x) = mantissa of fma(x - min, (2^52 - 2) / (max - min), 2^52): monotone in x, exact ties stay ties;
//   * digit passes: 4 bits of the key per pass, most significant first, both quantile levels at once, counted into
//     THREAD-PRIVATE 16-bin histograms in shared memory (no atomics: an entry belongs to one thread); after a pass every
//     (entry, level) narrows to the bin holding its target rank, until at most CAP values remain (MC = 5000: two or three
//     passes);
//   * collect pass: the remaining candidates of every (entry, level) go to a small private list, the smallest value above
//     the bin is tracked on the side; the list is sorted and the two neighbouring order statistics are interpolated
//     exactly as Armadillo's type-5 quantile does.
// Exactness: the selection only ever compares keys, which are a monotone function of the values, and the final
// candidates are compared as doubles; the one case it cannot decide -- more than CAP DISTINCT values sharing all 52 key
// bits -- is counted in `unresolved` and handled by the caller (per-pair kernel of post.cu).
#include <cmath>

#include "common.cuh"

namespace {

constexpr int TILE = 64, TV = 32;
constexpr int OSTR = TILE + 4, VSTR = TV + 4;      // operand row strides (4 mod 16: conflict-free DMMA fragment loads)
constexpr int PT = 512;                            // 16 warps: 4 along u (16 rows each) x 4 along v (8 columns each)
constexpr int NE = 4;                              // entries per thread: 2 (i: 8-row DMMA tiles) x 2 (h: adjacent columns)
constexpr int NB = 16;                             // bins per digit (4 bits)
constexpr int NDIG = 13;                           // 13 x 4 = 52 key bits
constexpr int CAP = 16;                            // candidates kept per (entry, level)
constexpr int KPMAX = 32;

struct Level {            // one quantile level, resolved on the host (Armadillo op_quantile, type 5)
  int mode;               // 0: interpolate x[r], x[r+1] with weight w; 1: -inf; 2: +inf
  int r;                  // 0-based rank of the lower neighbour
  double w;               // weight of the upper neighbour (0: the lower one alone)
};

struct PostTileArgs {
  const double* Lw;       // [MC][KP][pp]
  const double* s2w;      // [MC][pp]
  const int64_t* idx;     // [S] variable of every position (diagonal term when idx[a] == idx[b], cpp:770-772)
  int64_t pp, S, MC;
  int KP;
  Level lev[2];
  double* mean; double* lower; double* upper;   // S x S column-major
  double* scratch;        // [gridDim.x][PT][2][NE][CAP]
  int* unresolved;
  int64_t unit_begin, unit_end;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void tri_of(int64_t t, int& I, int& J) {
  int64_t j = static_cast<int64_t>((sqrt(8.0 * static_cast<double>(t) + 1.0) - 1.0) * 0.5);
  while (j * (j + 1) / 2 > t) --j;
  while ((j + 1) * (j + 2) / 2 <= t) ++j;
  J = static_cast<int>(j);
  I = static_cast<int>(t - j * (j + 1) / 2);
}

// order-preserving 52-bit key relative to the entry's own range
__device__ __forceinline__ unsigned long long key52(double x, double lo, double sc) {
  const double t = fma(x - lo, sc, 4503599627370496.0);            // in [2^52, 2^53): the mantissa is the integer part
  return static_cast<unsigned long long>(__double_as_longlong(t)) & 0xFFFFFFFFFFFFFull;
}

__global__ void __launch_bounds__(PT, 1) k_post_tiles(PostTileArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  const int KP = a.KP;
  const int opd = KP * (OSTR + VSTR) + TILE;                        // doubles of one operand buffer: u rows, v rows, sigma2 of the u rows
  double* op = reinterpret_cast<double*>(smem_dyn);
  unsigned short* hist = reinterpret_cast<unsigned short*>(op + 2 * opd);      // [2][NE][NB][PT]
  __shared__ long long s_iu[TILE], s_iv[TV];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wu = warp & 3, wv = warp >> 2;                          // 4 warps along u (16 rows), 4 along v (8 columns)
  const int nchunks = KP * (TILE / 2) + KP * (TV / 2) + TILE / 2;   // 16-byte pieces per draw
  double* myscratch = a.scratch + (static_cast<int64_t>(blockIdx.x) * PT + tid) * (2 * NE * CAP);

  for (int64_t unit = a.unit_begin + blockIdx.x; unit < a.unit_end; unit += gridDim.x) {
    int I, J;
    tri_of(unit >> 1, I, J);
    const int64_t u0 = static_cast<int64_t>(I) * TILE, v0 = static_cast<int64_t>(J) * TILE + (unit & 1) * TV;
    if (v0 >= a.S) continue;                                          // second half of an edge tile (uniform)
    __syncthreads();                                                  // the previous unit is done with s_iu / s_iv / op
    if (tid < TILE) s_iu[tid] = u0 + tid < a.S ? a.idx[u0 + tid] : -1 - tid;
    if (tid < TV) s_iv[tid] = v0 + tid < a.S ? a.idx[v0 + tid] : -1000 - tid;
    __syncthreads();
    // entry e = (i, j, h): row ul = wu*16 + i*8 + g, column vl = wv*16 + j*8 + 2t + h
    unsigned same = 0, valid = 0;
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      const int ul = wu * 16 + (e >> 1) * 8 + g, vl = wv * 8 + 2 * t + (e & 1);
      if (u0 + ul < a.S && v0 + vl < a.S) valid |= 1u << e;
      if (s_iu[ul] == s_iv[vl]) same |= 1u << e;
    }

    auto prefetch = [&](int64_t m, int buf) {
      const double* Lm = a.Lw + m * KP * a.pp;
      double* dst = op + buf * opd;
      for (int c = tid; c < nchunks; c += PT) {
        if (c < KP * (TILE / 2)) {
          const int l = c / (TILE / 2), x = c - l * (TILE / 2);
          cp16(dst + l * OSTR + 2 * x, Lm + static_cast<int64_t>(l) * a.pp + u0 + 2 * x);
        } else if (c < KP * (TILE / 2) + KP * (TV / 2)) {
          const int c2 = c - KP * (TILE / 2), l = c2 / (TV / 2), x = c2 - l * (TV / 2);
          cp16(dst + KP * OSTR + l * VSTR + 2 * x, Lm + static_cast<int64_t>(l) * a.pp + v0 + 2 * x);
        } else {
          const int x = c - KP * (TILE / 2) - KP * (TV / 2);
          cp16(dst + KP * (OSTR + VSTR) + 2 * x, a.s2w + m * a.pp + u0 + 2 * x);
        }
      }
      cp_commit();
    };
    // one pass over the draws: f(m, x[NE]) is called with the NE values of draw m
    auto stream = [&](auto&& f) {
      prefetch(0, 0);
      cp_wait_all();
      __syncthreads();
      for (int64_t m = 0; m < a.MC; ++m) {
        const int buf = static_cast<int>(m & 1);
        if (m + 1 < a.MC) prefetch(m + 1, buf ^ 1);
        const double* As = op + buf * opd + wu * 16 + g + t * OSTR;
        const double* Bs = op + buf * opd + KP * OSTR + wv * 8 + g + t * VSTR;
        const double* s2 = op + buf * opd + KP * (OSTR + VSTR);
        double c[2][2];
#pragma unroll
        for (int i = 0; i < 2; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
        for (int ks = 0; ks < KP / 4; ++ks) {
          double af[2];
#pragma unroll
          for (int i = 0; i < 2; ++i) af[i] = As[ks * 4 * OSTR + i * 8];
          const double bf = Bs[ks * 4 * VSTR];
#pragma unroll
          for (int i = 0; i < 2; ++i) dmma(c[i], af[i], bf);
        }
        double x[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          x[e] = c[e >> 1][e & 1];
          if (same & (1u << e)) x[e] += s2[wu * 16 + (e >> 1) * 8 + g];          // + sigma2_m[j1] (cpp:676-684)
        }
        f(m, x);
        cp_wait_all();
        __syncthreads();
      }
    };

    // ---- pass A: sum, minimum, maximum ------------------------------------------------------------------
    double sum[NE], lo[NE], sc[NE];
    {
      double hi[NE];
#pragma unroll
      for (int e = 0; e < NE; ++e) { sum[e] = 0.0; lo[e] = INFINITY; hi[e] = -INFINITY; }
      stream([&](int64_t, const double (&x)[NE]) {
#pragma unroll
        for (int e = 0; e < NE; ++e) { sum[e] += x[e]; lo[e] = fmin(lo[e], x[e]); hi[e] = fmax(hi[e], x[e]); }
      });
#pragma unroll
      for (int e = 0; e < NE; ++e) sc[e] = hi[e] > lo[e] ? 4503599627370494.0 / (hi[e] - lo[e]) : 0.0;
    }

    // ---- selection state per (level q, entry e) ------------------------------------------------------------------
    unsigned long long prefix[2][NE];
    int below[2][NE];
    unsigned done = 0, depth_bits_lo = 0, depth_bits_hi = 0;       // done: bit q*NE+e; depth: 4 bits per (q, e) packed in two words
    auto get_depth = [&](int q, int e) -> int { return ((q ? depth_bits_hi : depth_bits_lo) >> (4 * e)) & 15; };
    auto set_depth = [&](int q, int e, int d) {
      unsigned& wref = q ? depth_bits_hi : depth_bits_lo;
      wref = (wref & ~(15u << (4 * e))) | (static_cast<unsigned>(d) << (4 * e));
    };
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        prefix[q][e] = 0; below[q][e] = 0;
        if (a.lev[q].mode != 0 || a.MC <= CAP || sc[e] == 0.0 || !(valid & (1u << e))) done |= 1u << (q * NE + e);
      }

    for (int d = 0; d < NDIG; ++d) {
      if (__syncthreads_and(done == ((1u << (2 * NE)) - 1u))) break;
      for (int i = tid; i < 2 * NE * NB * PT; i += PT) hist[i] = 0;          // (a thread clears its own column: i % PT == tid)
      const int sh_pref = 52 - 4 * d, sh_dig = 48 - 4 * d;
      stream([&](int64_t, const double (&x)[NE]) {
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          if ((done >> e & 1u) && (done >> (NE + e) & 1u)) continue;
          const unsigned long long K = key52(x[e], lo[e], sc[e]);
          const unsigned long long pk = d == 0 ? 0ull : (K >> sh_pref);
          const int bin = static_cast<int>((K >> sh_dig) & 15ull);
#pragma unroll
          for (int q = 0; q < 2; ++q)
            if (!(done >> (q * NE + e) & 1u) && pk == prefix[q][e]) ++hist[((q * NE + e) * NB + bin) * PT + tid];
        }
      });
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          if (done >> (q * NE + e) & 1u) continue;
          const int want = a.lev[q].r - below[q][e];
          int cum = 0, b = 0;
          for (; b < NB - 1; ++b) {
            const int h = hist[((q * NE + e) * NB + b) * PT + tid];
            if (want < cum + h) break;
            cum += h;
          }
          prefix[q][e] = (prefix[q][e] << 4) | static_cast<unsigned long long>(b);
          below[q][e] += cum;
          set_depth(q, e, d + 1);
          if (hist[((q * NE + e) * NB + b) * PT + tid] <= CAP || d + 1 == NDIG) done |= 1u << (q * NE + e);
        }
    }

    // ---- collect pass: candidates of every (level, entry), smallest value above them ----------------------------------
    int cnt[2][NE];
    double above[2][NE];
    unsigned neq = 0;              // bit q*NE+e: a candidate beyond the first CAP differs from the first one
    bool any_interp = false;
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        cnt[q][e] = 0; above[q][e] = INFINITY;
        if (a.lev[q].mode == 0 && (valid & (1u << e)) && sc[e] != 0.0) any_interp = true;
      }
    if (__syncthreads_or(any_interp)) {
      stream([&](int64_t, const double (&x)[NE]) {
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          if (sc[e] == 0.0 || !(valid & (1u << e))) continue;
          const unsigned long long K = key52(x[e], lo[e], sc[e]);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (a.lev[q].mode != 0) continue;
            const int dp = get_depth(q, e);
            const unsigned long long pk = dp == 0 ? 0ull : (K >> (52 - 4 * dp));
            if (pk == prefix[q][e]) {
              if (cnt[q][e] < CAP) myscratch[(q * NE + e) * CAP + cnt[q][e]] = x[e];
              else if (x[e] != myscratch[(q * NE + e) * CAP]) neq |= 1u << (q * NE + e);
              ++cnt[q][e];
            } else if (pk > prefix[q][e]) {
              above[q][e] = fmin(above[q][e], x[e]);
            }
          }
        }
      });
    }

    // ---- finish: sort the candidates, interpolate (Armadillo quantile, type 5), write both triangles ---------------------
#pragma unroll
    for (int e = 0; e < NE; ++e) {
      if (!(valid & (1u << e))) continue;
      const int ul = wu * 16 + (e >> 1) * 8 + g, vl = wv * 8 + 2 * t + (e & 1);
      const int64_t ra = u0 + ul, rb = v0 + vl;
      double res[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const Level L = a.lev[q];
        double val;
        if (L.mode == 1) val = -INFINITY;
        else if (L.mode == 2) val = INFINITY;
        else if (sc[e] == 0.0) val = lo[e];                            // every draw gave the same value
        else {
          const int rr = L.r - below[q][e];
          double xr, xn;
          if (cnt[q][e] <= CAP) {
            double* li = myscratch + (q * NE + e) * CAP;
            const int nn = cnt[q][e];
            for (int i2 = 1; i2 < nn; ++i2) {                          // insertion sort of <= CAP values
              const double key = li[i2];
              int j2 = i2 - 1;
              while (j2 >= 0 && li[j2] > key) { li[j2 + 1] = li[j2]; --j2; }
              li[j2 + 1] = key;
            }
            xr = li[rr];
            xn = rr + 1 < nn ? li[rr + 1] : above[q][e];
          } else {
            const double* li = myscratch + (q * NE + e) * CAP;
            bool alleq = !(neq >> (q * NE + e) & 1u);
            for (int i2 = 1; i2 < CAP; ++i2) alleq = alleq && li[i2] == li[0];
            if (alleq) {                                               // a run of exact ties longer than CAP
              xr = li[0];
              xn = rr + 1 < cnt[q][e] ? li[0] : above[q][e];
            } else {                                                   // > CAP distinct values share all 52 key bits
              atomicAdd(a.unresolved, 1);
              xr = xn = __longlong_as_double(0x7ff8000000000000ll);
            }
          }
          val = L.w == 0.0 ? xr : (1.0 - L.w) * xr + L.w * xn;
        }
        res[q] = val;
      }
      const double mu = sum[e] / static_cast<double>(a.MC);
      a.mean[ra + a.S * rb] = mu; a.mean[rb + a.S * ra] = mu;
      a.lower[ra + a.S * rb] = res[0]; a.lower[rb + a.S * ra] = res[0];
      a.upper[ra + a.S * rb] = res[1]; a.upper[rb + a.S * ra] = res[1];
    }
  }
}

// reference layout [a][l][m] (MC contiguous per (selected variable, factor): what fable_post_processing stages) ->
// working layout Lw[(m*KP + l)*pp + a], s2w[m*pp + a], zero rows / ranks in the padding
__global__ void __launch_bounds__(256) k_ref_to_work(const double* __restrict__ Lsel, const double* __restrict__ Ssel, int64_t S,
                                                     int k, int64_t MC, int64_t pp, int KP, double* __restrict__ Lw,
                                                     double* __restrict__ s2w) {
  __shared__ double tle[32][33];
  // block: 32 positions (a) x 32 draws (m) of one rank row l (blockIdx.z = l, KP = the sigma2 plane)
  const int l = blockIdx.z;
  const int64_t a0 = static_cast<int64_t>(blockIdx.x) * 32, m0 = static_cast<int64_t>(blockIdx.y) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const bool sig = l == KP;
  for (int r = ty; r < 32; r += 8) {             // read: m contiguous
    const int64_t aa = a0 + r, m = m0 + tx;
    double v = 0.0;
    if (aa < S && m < MC && (sig || l < k)) v = sig ? Ssel[aa * MC + m] : Lsel[(aa * k + l) * MC + m];
    tle[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {             // write: a contiguous
    const int64_t m = m0 + r, aa = a0 + tx;
    if (m < MC && aa < pp) {
      if (sig) s2w[m * pp + aa] = tle[tx][r];
      else Lw[(m * KP + l) * pp + aa] = tle[tx][r];
    }
  }
}

// rows idx[a] of the full working layout -> compact working layout of the selection
__global__ void __launch_bounds__(256) k_gather_work_rows(const double* __restrict__ Lw, const double* __restrict__ s2w, int64_t pp,
                                                          const int64_t* __restrict__ idx, int64_t S, int64_t pps, int KP,
                                                          double* __restrict__ Lsel, double* __restrict__ s2sel) {
  const int64_t a = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t m = blockIdx.y;
  if (a >= pps) return;
  const int64_t j = a < S ? idx[a] : -1;
  for (int l = 0; l < KP; ++l) Lsel[(m * KP + l) * pps + a] = j >= 0 ? Lw[(m * KP + l) * pp + j] : 0.0;
  s2sel[m * pps + a] = j >= 0 ? s2w[m * pp + j] : 0.0;
}

Level make_level(double P, int64_t N) {           // Armadillo op_quantile (type 5): P_min = 0.5/N, P_max = (N-0.5)/N, h = N P + 0.5
  Level L;
  const double dn = static_cast<double>(N);
  if (P < 0.5 / dn) { L.mode = P < 0.0 ? 1 : 0; L.r = 0; L.w = 0.0; return L; }
  if (P > (dn - 0.5) / dn) { L.mode = P > 1.0 ? 2 : 0; L.r = static_cast<int>(N - 1); L.w = 0.0; return L; }
  const double h = dn * P + 0.5;
  const int kk = static_cast<int>(std::floor(h));
  L.mode = 0; L.r = kk - 1; L.w = h - static_cast<double>(kk);
  if (kk >= N) { L.r = static_cast<int>(N - 1); L.w = 0.0; }      // P == P_max exactly: the maximum
  return L;
}

}  // namespace

bool fb_post_tiles_supported(int64_t MC, int k) { return MC >= 1 && MC <= 65535 && k >= 1 && round_up(k, 4) <= KPMAX; }

// d_Lw [MC][KP][pp], d_s2w [MC][pp] (rows >= S zero), d_idx [S]; outputs S x S.  *unresolved = entries the selection could
// not decide (see the header comment; 0 in practice).
int fb_post_tiles(fable_ctx* ctx, const double* d_Lw, const double* d_s2w, const int64_t* d_idx, int64_t MC, int KP, int64_t S,
                  int64_t pp, double alpha, double* d_mean, double* d_lower, double* d_upper, int* unresolved) {
  FB_REQUIRE(ctx, KP % 4 == 0 && KP <= KPMAX && MC >= 1 && MC <= 65535 && pp % TILE == 0 && pp >= S, "post_tiles: unsupported shape");
  const int64_t nT = ceil_div(S, TILE), units = nT * (nT + 1);     // 2 halves per upper tile
  const size_t smem = static_cast<size_t>(2) * (KP * (OSTR + VSTR) + TILE) * sizeof(double) + static_cast<size_t>(2) * NE * NB * PT * sizeof(unsigned short);
  FB_CUDA(ctx, cudaFuncSetAttribute(k_post_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int64_t grid = ctx->sms;
  if (grid > units) grid = units;
  DevBuf scratch, flag;
  FB_CUDA(ctx, scratch.alloc(static_cast<size_t>(grid) * PT * 2 * NE * CAP * sizeof(double)));
  FB_CUDA(ctx, flag.alloc(sizeof(int)));
  FB_CUDA(ctx, cudaMemsetAsync(flag.ptr, 0, sizeof(int), ctx->stream));
  PostTileArgs a;
  a.Lw = d_Lw; a.s2w = d_s2w; a.idx = d_idx; a.pp = pp; a.S = S; a.MC = MC; a.KP = KP;
  a.lev[0] = make_level(alpha / 2.0, MC); a.lev[1] = make_level(1.0 - (alpha / 2.0), MC);
  a.mean = d_mean; a.lower = d_lower; a.upper = d_upper;
  a.scratch = scratch.d(); a.unresolved = static_cast<int*>(flag.ptr);
  // one launch, or slices when the caller wants to be polled between them (cpp:697-701 pattern)
  const int64_t slice = ctx->poll ? grid * 4 : units;
  for (int64_t u0 = 0; u0 < units; u0 += slice) {
    a.unit_begin = u0; a.unit_end = u0 + slice < units ? u0 + slice : units;
    k_post_tiles<<<static_cast<unsigned>(grid), PT, smem, ctx->stream>>>(a);
    FB_LAUNCHED(ctx);
    if (ctx->poll) {
      FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      if (ctx->poll(ctx->poll_user)) return ctx->fail(FABLE_ERR_INTERRUPT, "interrupted");
    }
  }
  int h = 0;
  FB_CUDA(ctx, cudaMemcpyAsync(&h, flag.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (unresolved) *unresolved = h;
  return FABLE_OK;
}

int fb_ref_to_work(fable_ctx* ctx, const double* dLsel, const double* dSsel, int64_t S, int k, int64_t MC, int64_t pp, int KP,
                   double* d_Lw, double* d_s2w) {
  dim3 grid(static_cast<unsigned>(ceil_div(pp, 32)), static_cast<unsigned>(ceil_div(MC, 32)), static_cast<unsigned>(KP + 1));
  k_ref_to_work<<<grid, 256, 0, ctx->stream>>>(dLsel, dSsel, S, k, MC, pp, KP, d_Lw, d_s2w);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}

int fb_gather_work_rows(fable_ctx* ctx, const double* d_Lw, const double* d_s2w, int64_t pp, const int64_t* d_idx, int64_t S,
                        int64_t pps, int KP, int64_t MC, double* d_Lsel, double* d_s2sel) {
  dim3 grid(static_cast<unsigned>(ceil_div(pps, 256)), static_cast<unsigned>(MC));
  k_gather_work_rows<<<grid, 256, 0, ctx->stream>>>(d_Lw, d_s2w, pp, d_idx, S, pps, KP, d_Lsel, d_s2sel);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}
