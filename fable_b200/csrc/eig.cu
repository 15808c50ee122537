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
#include <cstdlib>
#include <cmath>
#include <numeric>

#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

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
k_jacobi_round(double* __restrict__ A, int64_t m, int64_t ld, int round, int N, double tol, double floor2, int* __restrict__ nrot) {
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
  // a column below the noise floor of the Gram matrix (numerically zero: singular G, e.g. column-centred Y) is left alone:
  // its inner products are rounding noise, so the relative test would never settle
  if (gamma == 0.0 || fabs(gamma) <= tol * sqrt(alpha) * sqrt(beta) || alpha <= floor2 || beta <= floor2) return;
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


// ---------------------------------------------------------------------------------------------------
// Block one-sided Jacobi (m > kBlockMin).  Columns are grouped in blocks of BW = 32; a round pairs the
// blocks (round-robin) and for every pair (I, J) of a round
//   1. k_bj_gram : H = W'W for the m x 64 panel W = [A_I A_J]   (FP64 tensor cores, split over the rows)
//   2. k_bj_eig  : one two-sided cyclic Jacobi sweep on the 64 x 64 matrix H in shared memory -> R
//                  (inexact inner solve: more inner sweeps cost 0.9 ms per pair and do not reduce the
//                  number of outer sweeps; measured)
//                  Rounds 1 .. NB-2 of a sweep rotate only the 32 x 32 cross pairs (bj_solve's symmetric form: H as its
//                  upper triangle, the rotations recorded) and
//  2b. k_bj_replay rebuilds R from the 16 KB of rotations on 8 CTAs per pair;
//   3. k_bj_apply: W <- W R         (FP64 tensor cores, in place, one CTA per 64 rows)
// so a round streams the matrix through HBM once per 32 columns instead of once per column: m/32 - 1
// rounds per sweep instead of m - 1.  Where every pair's cluster of 8 CTAs is resident (m <= 960) the whole round is ONE
// kernel, k_bj_round, with the panel slabs in shared memory.  A sweep is replayed as a CUDA graph.  The working matrix is padded with zero rows (to a multiple of 16)
// and zero columns (to an even number of blocks); zero columns are never rotated.
constexpr int BW = 32, PW = 2 * BW;
constexpr int GKT = 16, GS = GKT + 4;      // Gram stage: 16 rows, column stride 20 (= 4 mod 16: conflict-free fragments)
constexpr int RSTR = PW + 4;

// Programmatic dependent launch (the four kernels of a round are a strict chain): a kernel lets its successor be scheduled
// as soon as all of its own CTAs are resident, and waits for its predecessor's completion (and memory flush) before it
// touches global memory -- the successor's launch latency and CTA start-up overlap the predecessor's tail.  Without the
// launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_chain() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ void bj_pair(int i, int round, int NB, int& bi, int& bj) {
  int ca, cb;
  if (i == 0) { ca = NB - 1; cb = round; }
  else { ca = (round + i) % (NB - 1); cb = (round - i + (NB - 1)) % (NB - 1); }
  bi = ca < cb ? ca : cb; bj = ca < cb ? cb : ca;
}
__device__ __forceinline__ int64_t bj_col(int a, int bi, int bj) { return a < BW ? static_cast<int64_t>(bi) * BW + a : static_cast<int64_t>(bj) * BW + (a - BW); }

__device__ __forceinline__ void dmma4(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst))), "l"(src) : "memory");
}

// H_part[(pair * nsplit + split)][64 x 64 column-major] = sum over the split's rows of W[i,a] W[i,b] -- the 8 x 8 tiles on
// and above the diagonal are formed (36 of 64: the FP64 tensor pipe bounds this kernel, 112 us of a 276 us round at m = 5000)
// and stored twice (mirrored reads in k_bj_eig were uncoalesced: n = 1000 16 -> 21 ms).  Warp q owns tile rows q and 7 - q (8 - q and q + 1 tiles: nine per warp).
__global__ void __launch_bounds__(128)
k_bj_gram(const double* __restrict__ A, int64_t ld, int64_t rows_per_split, int64_t mrows, int round, int NB,
          double* __restrict__ Hpart) {
  __shared__ __align__(16) double T[2][PW * GS];
  pdl_chain();
  int bi, bj;
  bj_pair(blockIdx.x, round, NB, bi, bj);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int64_t r0 = blockIdx.y * rows_per_split;
  const int64_t r1 = r0 + rows_per_split < mrows ? r0 + rows_per_split : mrows;
  double c[9][2];
#pragma unroll
  for (int i = 0; i < 9; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
  // stage = 16 rows x 64 columns = 512 chunks of 16 bytes (2 rows); 4 per thread
  auto load = [&](int64_t i0, int buf) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = tid + 128 * q, a = idx >> 3, ch = idx & 7;
      cp16(&T[buf][a * GS + 2 * ch], A + bj_col(a, bi, bj) * ld + i0 + 2 * ch);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load(r0, 0);
  int buf = 0;
  const int rowA = warp, rowB = 7 - warp;                   // tile rows; columns rowA .. 7 and rowB .. 7
  for (int64_t i0 = r0; i0 < r1; i0 += GKT, buf ^= 1) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (i0 + GKT < r1) load(i0 + GKT, buf ^ 1);
    const double* Ts = T[buf];
#pragma unroll
    for (int ks = 0; ks < GKT / 4; ++ks) {
      const double afA = Ts[(rowA * 8 + g) * GS + ks * 4 + t], afB = Ts[(rowB * 8 + g) * GS + ks * 4 + t];
      // slot i of the warp: i < 8 - warp -> tile (rowA, rowA + i), else tile (rowB, rowB + i - (8 - warp))
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const bool first = i < 8 - warp;                    // warp-uniform
        const int tj = first ? rowA + i : rowB + i - (8 - warp);
        dmma4(c[i], first ? afA : afB, Ts[(tj * 8 + g) * GS + ks * 4 + t]);
      }
    }
  }
  double* H = Hpart + (static_cast<int64_t>(blockIdx.x) * gridDim.y + blockIdx.y) * (PW * PW);
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const bool first = i < 8 - warp;
    const int ti = first ? rowA : rowB, tj = first ? rowA + i : rowB + i - (8 - warp);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      H[(ti * 8 + g) + PW * (tj * 8 + 2 * t + e)] = c[i][e];
      if (ti != tj) H[(tj * 8 + 2 * t + e) + PW * (ti * 8 + g)] = c[i][e];      // mirror image: the reader stays coalesced
    }
  }
}

// two-sided cyclic Jacobi on H (64 x 64, shared memory), rotations accumulated in R; one CTA per pair.
// This kernel is 85-90 % of a round (ncu: gram 10 us, apply 10 us, first version of this one 187 us), i.e. of the
// whole SVD at n <= 1000, and it is pure latency: 63 dependent steps.  Per step: one warp computes the 32 rotations
// (one sqrt, one division, one rsqrt: t = 2 h_pq / (d + sign(d) sqrt(d^2 + 4 h_pq^2)), d = h_qq - h_pp, c = rsqrt(1 +
// t^2) -- the textbook zeta form costs two sqrt and three divisions in sequence), then every 2 x 2 block {p_i, q_i} x
// {p_j, q_j} of H is rotated from both sides by the one thread that owns it (no barrier between the column and the
// row update, half the shared-memory traffic) while R <- R J proceeds pair by pair: two barriers per step, 1024
// threads so that each owns one block and two R pairs (512 threads: +12 %).
// cross_only: only the 32 x 32 pairs (column of block I, column of block J) are rotated, 32 steps instead of 63.  The
// pairs inside a block need one visit per outer sweep, not one per round: round 0 of every sweep runs the full 63-step
// ordering, the other rounds the cross pairs, so an outer sweep is an exact cyclic sweep over all element pairs.  Same
// 14-15 outer sweeps, 40 % less time (measured at n = 500, 1000, 2000).
// REPLAY (cross_only launches with one inner sweep): the CTA does not carry R at all.  A step is bound by the shared-memory
// pipe (per warp 21 wavefronts for the H block + 18 for its two R row pairs: 1250 of the ~2000 cycles of a step on the one SM
// a pair occupies), and R <- R J is row-wise independent work that only needs the rotations: warp 0 writes every rotation
// it derives to global memory as well (32 steps x 32 pairs x (c, s) = 16 KB per block pair) and k_bj_replay rebuilds R from
// them on 8 CTAs per pair, one warp per row of R.  mode[pair]: 0 = nothing to rotate (R = I), 1 = rotations recorded,
// 2 = R written here (full ordering / legacy path).
constexpr int EIGT = 1024;
#ifdef BJ_LAB
__device__ int bj_lab_count = 0, bj_lab_count2 = 0;
#endif
// The inner solve proper; called by all EIGT threads of a CTA with H (64 x 64, row stride 65) at dsm and, unless REPLAY &&
// cross_only, R = I behind it.  Returns (uniformly) the mode: 0 = nothing to rotate (R = I), 1 = cross rotations recorded in
// rots_pair (global or shared), 2 = R complete in shared memory behind H.
template <bool REPLAY>
__device__ __forceinline__ int bj_solve(double* dsm, double tol, double floor2, int inner, int cross_only, int* __restrict__ nrot,
                                        double2* rots_pair, int iso) {
  constexpr int S = PW + 1;
  double* H = dsm;
  double* R = dsm + PW * S;
  __shared__ double cs[BW], sn[BW];
  __shared__ double2 rot2[2][BW];
  __shared__ int pp[BW], qq[BW];
  __shared__ int any_rot, need;
  const int tid = threadIdx.x;
  const int nsteps = cross_only ? BW : PW - 1;
  const double tol2 = tol * tol;
#ifdef BJ_LAB
  const long long lab_s0 = clock64();
#endif
  if (tid == 0) { any_rot = 0; need = 0; }
  __syncthreads();
  // does this pair need work at all?  (outer convergence test: |h_ab| > tol sqrt(h_aa h_bb))
  for (int idx = tid; idx < PW * PW; idx += EIGT) {
    const int a = idx & (PW - 1), b = idx >> 6;
    if (a < b) {
      const double h = H[a * S + b];
      if (h != 0.0 && h * h > tol2 * H[a * S + a] * H[b * S + b] && H[a * S + a] > floor2 && H[b * S + b] > floor2) need = 1;
    }
  }
  __syncthreads();
  if constexpr (REPLAY) {
    if (cross_only && !(iso & 2)) {
      // Symmetric form of the pipelined cross steps below (which it replaces when the rotations are recorded).  Measured
      // with clock64 in the form below (tools: -DBJ_LAB): a step took 1470 cycles, the deriving warp 0 was busy for 1360 of
      // them and every applying warp for 920 -- the 31 warps' 650 shared-memory wavefronts per step queue in front of warp
      // 0's own twelve loads, and their FP64 instructions in front of its dependent chain (alone on its scheduler: 810).
      // So (i) H is kept as its upper triangle only: entry (r, c) lives at [min][max], a block (row pair a, column pair b)
      // with a <= b is rotated from both sides by one thread and its mirror image is never formed -- 528 blocks on 17 warps
      // instead of 1024 on 31, 24 instead of 21 wavefronts per warp; (ii) those 17 warps are the ones that do NOT share a
      // scheduler with warp 0 (warp ids 1-3, 5-7, .. 21-22), the other warps leave; (iii) the step barrier is a named
      // barrier of the 18 warps.  A diagonal block needs no special case: its (q, p) address is its (p, q) address.
      if (!need) return 0;
      if (tid == 0) atomicAdd(nrot, 1);
      const int lane = tid & 31, w = tid >> 5;
      const bool applies = (w & 3) != 0 && w <= 22;
      if (w == 0 || applies) {
      const int v = w - 1 - (w >> 2);                    // 0 .. 16 over the applying warps
      int a, b;
      if (v < 16) {
        if (lane >= v) { a = v; b = lane; } else { a = 31 - v; b = 32 - v + lane; }
      } else {
        a = b = 16 + (lane & 15);
      }
      const bool has_block = applies && (v < 16 || lane < 16);
      auto rotation = [&](double hpq, double hpp, double hqq, double2& cs_) {
        cs_ = make_double2(1.0, 0.0);
        if (hpq != 0.0 && hpq * hpq > tol2 * hpp * hqq && hpp > floor2 && hqq > floor2) {
          const double d = hqq - hpp, h2 = 2.0 * hpq, r2 = fma(d, d, h2 * h2);
          const double tt = h2 * (1.0 / (d + copysign(r2 * rsqrt(r2), d)));
          const double c_ = rsqrt(fma(tt, tt, 1.0));
          cs_ = make_double2(c_, c_ * tt);
        }
      };
      if (w == 0) {
        const int q = BW + lane;
        double2 r0;
        rotation(H[lane * S + q], H[lane * S + lane], H[q * S + q], r0);
        rot2[0][lane] = r0;
        rots_pair[lane] = r0;
      }
      asm volatile("bar.sync 1, 576;" ::: "memory");
#ifdef BJ_LAB
      if (blockIdx.x == 0 && tid == 0 && bj_lab_count2 == 300) printf("BJ_LAB(solve) need check + first rotation %lld cycles\n", clock64() - lab_s0);
#endif
      int cur = 0, nxt = 2 * PW * S;
      int o_a = a, o_b = b, o_l = lane;                  // (index + step) % 32
#ifdef BJ_LAB
      long long lab_busy = 0, lab_t0 = clock64();
#endif
      for (int step = 0; step < BW; ++step) {
        const double2* rr = rot2[step & 1];
        const double* Hc = dsm + cur;
        double* Hn = dsm + nxt;
#ifdef BJ_LAB
        const long long lab_s = clock64();
#endif
        if (w == 0) {
          if (step + 1 < BW) {
            const int i = lane, i1 = (i + 1) & (BW - 1), qi = BW + o_l, qn = BW + ((o_l + 1) & (BW - 1));
            const int ilo = i < i1 ? i : i1, ihi = i < i1 ? i1 : i, qlo = qi < qn ? qi : qn, qhi = qi < qn ? qn : qi;
            const double2 r1 = rr[i], r2 = rr[i1];
            const double h_ii = Hc[i * S + i], h_iqi = Hc[i * S + qi], h_qiqi = Hc[qi * S + qi];
            const double h_11 = Hc[i1 * S + i1], h_1qn = Hc[i1 * S + qn], h_qnqn = Hc[qn * S + qn];
            const double h_i1 = Hc[ilo * S + ihi], h_iqn = Hc[i * S + qn], h_1qi = Hc[i1 * S + qi], h_qiqn = Hc[qlo * S + qhi];
            const double ap = r1.x * h_ii - r1.y * h_iqi, aq = r1.x * h_iqi - r1.y * h_qiqi;
            const double hii = r1.x * ap - r1.y * aq;                                             // H'[i][i]
            const double bp1 = r2.y * h_11 + r2.x * h_1qn, bq1 = r2.y * h_1qn + r2.x * h_qnqn;
            const double hqq = r2.y * bp1 + r2.x * bq1;                                           // H'[qn][qn]
            const double bp = r2.y * h_i1 + r2.x * h_iqn, bq = r2.y * h_1qi + r2.x * h_qiqn;
            const double hpq = r1.x * bp - r1.y * bq;                                             // H'[i][qn]
            double2 rn;
            rotation(hpq, hii, hqq, rn);
            rot2[(step + 1) & 1][i] = rn;
            rots_pair[(step + 1) * BW + i] = rn;
          }
        } else if (has_block) {
          const double2 r1 = rr[a], r2 = rr[b];
          const int lo = o_a < o_b ? o_a : o_b, hi = o_a < o_b ? o_b : o_a;
          const int A_pp = a * S + b, A_pq = a * S + BW + o_b, A_qp = b * S + BW + o_a, A_qq = (BW + lo) * S + BW + hi;
          const double hpp_ = Hc[A_pp], hpq_ = Hc[A_pq], hqp_ = Hc[A_qp], hqq_ = Hc[A_qq];
          const double ap = r2.x * hpp_ - r2.y * hpq_, bp = r2.y * hpp_ + r2.x * hpq_;
          const double aq = r2.x * hqp_ - r2.y * hqq_, bq = r2.y * hqp_ + r2.x * hqq_;
          Hn[A_pp] = r1.x * ap - r1.y * aq; Hn[A_qp] = r1.y * ap + r1.x * aq;
          Hn[A_pq] = r1.x * bp - r1.y * bq; Hn[A_qq] = r1.y * bp + r1.x * bq;
        }
#ifdef BJ_LAB
        lab_busy += clock64() - lab_s;
#endif
        asm volatile("bar.sync 1, 576;" ::: "memory");
        const int tswap = cur; cur = nxt; nxt = tswap;
        o_a = (o_a + 1) & (BW - 1); o_b = (o_b + 1) & (BW - 1); o_l = (o_l + 1) & (BW - 1);
      }
#ifdef BJ_LAB
      if (blockIdx.x == 0 && lane == 0 && (w == 0 || w == 1 || w == 5 || w == 22) && atomicAdd(&bj_lab_count, 1) / 4 == 300)
        printf("BJ_LAB(sym) warp %2d: busy %lld cycles of %lld for 32 steps\n", w, lab_busy, clock64() - lab_t0);
#endif
      }
      __syncthreads();
      return 1;
    }
  }
  if (need && cross_only) {
    // Cross pairs, software-pipelined: step s rotates (i, 32 + (i + s) % 32), i = 0 .. 31.  While warps 1-31 apply step s
    // (H_next = J' H_cur J into the OTHER buffer, R <- R J in place), warp 0 derives the rotations of step s + 1: the three
    // entries it needs of the not-yet-written H_next are formed from H_cur and the current rotations with the block
    // formulas of the update itself.  One barrier per step.  The kernel is bound by instruction issue (ncu: 1350 warp
    // instructions per scheduler and step in the first form), so every thread keeps ONE fixed block (warp w: row pair
    // w - 1, column pair = lane; warp 31 also row pair 31) and fixed R entries, with addresses advanced by increments.
    if (tid == 0) atomicAdd(nrot, 1);
    const int lane = tid & 31, w = tid >> 5;
    int cur = 0, nxt = 2 * PW * S;                       // offsets of H_cur / H_next in dsm (H at 0, second H behind R)
    auto rotation = [&](double hpq, double hpp, double hqq, double2& cs_) {
      cs_ = make_double2(1.0, 0.0);
      if (hpq != 0.0 && hpq * hpq > tol2 * hpp * hqq && hpp > floor2 && hqq > floor2) {
        const double d = hqq - hpp, h2 = 2.0 * hpq, r2 = fma(d, d, h2 * h2);
        const double tt = h2 * (1.0 / (d + copysign(r2 * rsqrt(r2), d)));
        const double c_ = rsqrt(fma(tt, tt, 1.0));
        cs_ = make_double2(c_, c_ * tt);
      }
    };
    auto block = [&](const double* Hc, double* Hn, int p1S, int q1S, int p2, int q2, double2 r1, double2 r2) {
      const double hpp_ = Hc[p1S + p2], hpq_ = Hc[p1S + q2], hqp_ = Hc[q1S + p2], hqq_ = Hc[q1S + q2];
      const double ap = r2.x * hpp_ - r2.y * hpq_, bp = r2.y * hpp_ + r2.x * hpq_;
      const double aq = r2.x * hqp_ - r2.y * hqq_, bq = r2.y * hqp_ + r2.x * hqq_;
      Hn[p1S + p2] = r1.x * ap - r1.y * aq; Hn[q1S + p2] = r1.y * ap + r1.x * aq;
      Hn[p1S + q2] = r1.x * bp - r1.y * bq; Hn[q1S + q2] = r1.y * bp + r1.x * bq;
    };
    for (int sweep = 0; sweep < inner; ++sweep) {
      if (tid == 0) any_rot = 0;
      if (tid < BW) {
        const double* Hc = dsm + cur;
        const int q = BW + tid;
        double2 r0;
        rotation(Hc[tid * S + q], Hc[tid * S + tid], Hc[q * S + q], r0);
        rot2[0][tid] = r0;
        if (REPLAY) rots_pair[tid] = r0;
        if (r0.y != 0.0) any_rot = 1;
      }
      __syncthreads();
      // rotating column offsets: o = (index + step) % 32
      int o_lane = lane, o_w = (w - 1) & (BW - 1), o_r0 = (tid - BW) >> 6, o_r1 = ((tid - BW) >> 6) + 16;
#ifdef BJ_LAB
      long long lab_busy = 0, lab_t0 = clock64();
#endif
      for (int step = 0; step < BW; ++step) {
        const double2* rr = rot2[step & 1];
        const double* Hc = dsm + cur;
        double* Hn = dsm + nxt;
#ifdef BJ_LAB
        const long long lab_s = clock64();
#endif
        if (w == 0) {
          if (step + 1 < BW) {
            const int i = lane, i1 = (i + 1) & (BW - 1), qi = BW + o_lane, qn = BW + ((o_lane + 1) & (BW - 1));
            const double2 r1 = rr[i], r2 = rr[i1];
            const double ap = r1.x * Hc[i * S + i] - r1.y * Hc[i * S + qi], aq = r1.x * Hc[qi * S + i] - r1.y * Hc[qi * S + qi];
            const double hii = r1.x * ap - r1.y * aq;                                             // H'[i][i]
            const double bp1 = r2.y * Hc[i1 * S + i1] + r2.x * Hc[i1 * S + qn], bq1 = r2.y * Hc[qn * S + i1] + r2.x * Hc[qn * S + qn];
            const double hqq = r2.y * bp1 + r2.x * bq1;                                           // H'[qn][qn]
            const double bp = r2.y * Hc[i * S + i1] + r2.x * Hc[i * S + qn], bq = r2.y * Hc[qi * S + i1] + r2.x * Hc[qi * S + qn];
            const double hpq = r1.x * bp - r1.y * bq;                                             // H'[i][qn]
            double2 rn;
            rotation(hpq, hii, hqq, rn);
            rot2[(step + 1) & 1][i] = rn;
            if (REPLAY) rots_pair[(step + 1) * BW + i] = rn;
            if (rn.y != 0.0) any_rot = 1;
          }
        } else if (iso && (w & 3) == 0) {
          // iso: warps 4, 8, .. 28 share the scheduler (and its FP64 pipe) of the deriving warp 0 and stay idle; their row
          // pairs are taken by the next warp
        } else {
          const int q2 = BW + o_lane;
          const double2 r2 = rr[lane];
          block(Hc, Hn, (w - 1) * S, (BW + o_w) * S, lane, q2, rr[w - 1], r2);
          if (iso && (w & 3) == 1 && w > 1) block(Hc, Hn, (w - 2) * S, (BW + ((o_w - 1) & (BW - 1))) * S, lane, q2, rr[w - 2], r2);
          if (w == BW - 1) block(Hc, Hn, (BW - 1) * S, (BW + ((o_w + 1) & (BW - 1))) * S, lane, q2, rr[BW - 1], r2);
          // R <- R J: thread u = tid - 32 owns row i = u % 64 of the pairs u / 64 and u / 64 + 16 (u < 992: pairs 15 and 31
          // have rows 32 .. 63 left over)
          if (!REPLAY) {
          const int i = (tid - BW) & (PW - 1);
          {
            const int r = (tid - BW) >> 6;
            const double2 cs_ = rr[r];
            if (cs_.y != 0.0) {
              const int q = BW + ((o_r0 + step) & (BW - 1));
              const double rp = R[i * S + r], rq = R[i * S + q];
              R[i * S + r] = cs_.x * rp - cs_.y * rq; R[i * S + q] = cs_.y * rp + cs_.x * rq;
            }
          }
          {
            const int r = ((tid - BW) >> 6) + 16;
            if (r < BW) {
              const double2 cs_ = rr[r];
              if (cs_.y != 0.0) {
                const int q = BW + ((o_r1 + step) & (BW - 1));
                const double rp = R[i * S + r], rq = R[i * S + q];
                R[i * S + r] = cs_.x * rp - cs_.y * rq; R[i * S + q] = cs_.y * rp + cs_.x * rq;
              }
            }
          }
          if (w <= 2) {                                   // left over: rows 32 .. 63 of pairs 15 (warp 1) and 31 (warp 2)
            const int i2 = BW + lane, r = w == 1 ? 15 : BW - 1;
            const double2 cs_ = rr[r];
            if (cs_.y != 0.0) {
              const int q = BW + ((r + step) & (BW - 1));
              const double rp = R[i2 * S + r], rq = R[i2 * S + q];
              R[i2 * S + r] = cs_.x * rp - cs_.y * rq; R[i2 * S + q] = cs_.y * rp + cs_.x * rq;
            }
          }
          }
        }
#ifdef BJ_LAB
        lab_busy += clock64() - lab_s;
#endif
        __syncthreads();
        const int tswap = cur; cur = nxt; nxt = tswap;
        o_lane = (o_lane + 1) & (BW - 1); o_w = (o_w + 1) & (BW - 1);
      }
#ifdef BJ_LAB
      if (blockIdx.x == 0 && lane == 0 && (w == 0 || w == 1 || w == 5 || w == 31) && atomicAdd(&bj_lab_count, 1) / 4 == 300)
        printf("BJ_LAB warp %2d: busy %lld cycles of %lld for 32 steps (iso %d)\n", w, lab_busy, clock64() - lab_t0, iso);
#endif
      if (!any_rot) break;
    }
  } else if (need) {
    if (tid == 0) atomicAdd(nrot, 1);
    for (int sweep = 0; sweep < inner; ++sweep) {
      if (tid == 0) any_rot = 0;
      __syncthreads();
      for (int step = 0; step < nsteps; ++step) {
        if (tid < BW) {
          int a, b;
          if (cross_only) { a = tid; b = BW + ((tid + step) & (BW - 1)); }       // column of block I x column of block J
          else if (tid == 0) { a = PW - 1; b = step; }
          else { a = (step + tid) % (PW - 1); b = (step - tid + (PW - 1)) % (PW - 1); }
          const int p = a < b ? a : b, q = a < b ? b : a;
          const double hpq = H[p * S + q], hpp = H[p * S + p], hqq = H[q * S + q];
          double c_ = 1.0, s_ = 0.0;
          if (hpq != 0.0 && hpq * hpq > tol2 * hpp * hqq && hpp > floor2 && hqq > floor2) {
            const double d = hqq - hpp, h2 = 2.0 * hpq, r2 = fma(d, d, h2 * h2);
            const double tt = h2 * (1.0 / (d + copysign(r2 * rsqrt(r2), d)));      // rsqrt, reciprocal, rsqrt: no general division
            c_ = rsqrt(fma(tt, tt, 1.0)); s_ = c_ * tt;
            any_rot = 1;
          }
          cs[tid] = c_; sn[tid] = s_; pp[tid] = p; qq[tid] = q;
        }
        __syncthreads();
        // H <- J' H J, block by block
        for (int blk = tid; blk < BW * BW; blk += EIGT) {
          const int ri = blk >> 5, rj = blk & (BW - 1);
          const double c1 = cs[ri], s1 = sn[ri], c2 = cs[rj], s2 = sn[rj];
          if (s1 != 0.0 || s2 != 0.0) {
            const int p1 = pp[ri], q1 = qq[ri], p2 = pp[rj], q2 = qq[rj];
            const double hpp_ = H[p1 * S + p2], hpq_ = H[p1 * S + q2], hqp_ = H[q1 * S + p2], hqq_ = H[q1 * S + q2];
            const double ap = c2 * hpp_ - s2 * hpq_, bp = s2 * hpp_ + c2 * hpq_;     // columns, row p1
            const double aq = c2 * hqp_ - s2 * hqq_, bq = s2 * hqp_ + c2 * hqq_;     // columns, row q1
            H[p1 * S + p2] = c1 * ap - s1 * aq; H[q1 * S + p2] = s1 * ap + c1 * aq;  // rows
            H[p1 * S + q2] = c1 * bp - s1 * bq; H[q1 * S + q2] = s1 * bp + c1 * bq;
          }
        }
        // R <- R J
        for (int idx = tid; idx < BW * PW; idx += EIGT) {
          const int r = idx >> 6, i = idx & (PW - 1);
          const double c_ = cs[r], s_ = sn[r];
          if (s_ != 0.0) {
            const int p = pp[r], q = qq[r];
            const double rp = R[i * S + p], rq = R[i * S + q];
            R[i * S + p] = c_ * rp - s_ * rq; R[i * S + q] = s_ * rp + c_ * rq;
          }
        }
        __syncthreads();
      }
      if (!any_rot) break;
    }
  }
  return (REPLAY && cross_only) ? (need ? 1 : 0) : 2;
}

template <bool REPLAY>
__global__ void __launch_bounds__(EIGT)
k_bj_eig(const double* __restrict__ Hpart, int nsplit, double tol, double floor2, int inner, int cross_only, double* __restrict__ Rout,
         int* __restrict__ nrot, double2* __restrict__ rots, int* __restrict__ mode, int iso) {
  constexpr int S = PW + 1;
  extern __shared__ double dsm[];
  double* H = dsm;
  double* R = dsm + PW * S;
  const int tid = threadIdx.x;
  const double* src = Hpart + static_cast<int64_t>(blockIdx.x) * nsplit * (PW * PW);
  pdl_chain();
#ifdef BJ_LAB
  const long long lab_k0 = clock64();
#endif
  {
    // partial Grams of the row splits, added in split order; the loads of a thread's four entries go out together,
    // eight splits at a time (one load per iteration = nsplit x 4 serial L2 round trips, 20 us of a 50 us kernel)
    // every thread owns two pairs of adjacent rows (16-byte loads: half as many requests as one entry per load -- the fetch is
    // bound by the number of loads in flight, not by bytes); the symmetric cross steps read H on and above the diagonal only,
    // so the 8 x 8 tiles below it are not fetched there
    constexpr int NP = PW * PW / (2 * EIGT);
    double2 v[NP];
    const bool upper_only = REPLAY && cross_only && !(iso & 2);
    bool want[NP];
#pragma unroll
    for (int e = 0; e < NP; ++e) {
      v[e] = make_double2(0.0, 0.0);
      const int idx = 2 * (tid + e * EIGT);
      want[e] = !upper_only || ((idx & (PW - 1)) >> 3) <= (idx >> 9);
    }
    for (int s0 = 0; s0 < nsplit; s0 += 8) {
      double2 x[8][NP];
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int e = 0; e < NP; ++e)
          x[u][e] = (s0 + u < nsplit && want[e])
                        ? __ldg(reinterpret_cast<const double2*>(src + static_cast<int64_t>(s0 + u) * (PW * PW)) + tid + e * EIGT)
                        : make_double2(0.0, 0.0);
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int e = 0; e < NP; ++e) { v[e].x += x[u][e].x; v[e].y += x[u][e].y; }
    }
#pragma unroll
    for (int e = 0; e < NP; ++e) {
      const int idx = 2 * (tid + e * EIGT), a = idx & (PW - 1), b = idx >> 6;
      H[a * S + b] = v[e].x;
      H[(a + 1) * S + b] = v[e].y;
      if (!(REPLAY && cross_only)) {
        R[a * S + b] = (a == b) ? 1.0 : 0.0;
        R[(a + 1) * S + b] = (a + 1 == b) ? 1.0 : 0.0;
      }
    }
  }
#ifdef BJ_LAB
  __syncthreads();
  const long long lab_k1 = clock64();
#endif
  const int md = bj_solve<REPLAY>(dsm, tol, floor2, inner, cross_only, nrot,
                                  REPLAY ? rots + static_cast<int64_t>(blockIdx.x) * (BW * BW) : nullptr, iso);
#ifdef BJ_LAB
  if (blockIdx.x == 0 && tid == 0 && REPLAY && atomicAdd(&bj_lab_count2, 1) == 300)
    printf("BJ_LAB(kernel) fetch of %d partial Grams %lld cycles, solve (need check, first rotation, 32 steps) %lld cycles\n", nsplit, lab_k1 - lab_k0,
           clock64() - lab_k1);
#endif
  if (tid == 0 && mode) mode[blockIdx.x] = md;
  if (md != 2) return;                                     // k_bj_replay writes R
  double* Rg = Rout + static_cast<int64_t>(blockIdx.x) * (PW * PW);
  for (int idx = tid; idx < PW * PW; idx += EIGT) {
    const int a = idx & (PW - 1), b = idx >> 6;
    Rg[a + PW * b] = R[a * S + b];          // column-major 64 x 64
  }
}

// R of a block pair from the recorded cross rotations: one warp per row of R, lane i holds the entries of column i (p) and
// of column 32 + (i + step) % 32 (q) -- the q entries move one lane per step, so step s rotates (i, 32 + (i + s) % 32) as
// k_bj_eig does.  32 dependent steps of two FMAs and a shuffle: ~1 us, on 8 CTAs per pair instead of inside the one CTA
// whose shared-memory pipe bounds the round.  grid (pairs, 8), 256 threads.
__global__ void __launch_bounds__(256)
k_bj_replay(const double2* __restrict__ rots, const int* __restrict__ mode, double* __restrict__ Rall) {
  __shared__ double2 rs[BW * BW];                        // the pair's 32 x 32 rotations: one coalesced fetch, then LDS per step
  pdl_chain();
  const int md = mode[blockIdx.x];                       // (straight from global the 32 loads of a lane were serialised: 8.6 us)
  if (md == 2) return;
  const int lane = threadIdx.x & 31, row = blockIdx.y * 8 + (threadIdx.x >> 5);
  double p = row == lane ? 1.0 : 0.0, q = row == BW + lane ? 1.0 : 0.0;
  if (md == 1) {
    const double2* r = rots + static_cast<int64_t>(blockIdx.x) * (BW * BW);
    double2 v[BW * BW / 256];
#pragma unroll
    for (int e = 0; e < BW * BW / 256; ++e) v[e] = r[threadIdx.x + e * 256];
#pragma unroll
    for (int e = 0; e < BW * BW / 256; ++e) rs[threadIdx.x + e * 256] = v[e];
    __syncthreads();
    double2 cs = rs[lane];
#pragma unroll
    for (int s = 0; s < BW; ++s) {
      const double2 nx = rs[((s + 1) & (BW - 1)) * BW + lane];      // next step's rotation while this one is applied
      const double np = cs.x * p - cs.y * q, nq = cs.y * p + cs.x * q;
      p = np;
      q = __shfl_sync(0xffffffffu, nq, (lane + 1) & 31);
      cs = nx;
    }
  }
  double* Rg = Rall + static_cast<int64_t>(blockIdx.x) * (PW * PW);
  Rg[row + PW * lane] = p;
  Rg[row + PW * (BW + lane)] = q;
}

// W (rows i0 .. i0 + RB - 1 of the pair's 64 columns) <- W R, in place; RB / 16 warps, each a 32 x 32 piece.  RB = 64 (four
// warps, 66 KB of shared memory: three CTAs per SM, and 256 instead of 128 CTAs at m = 1000, where 128 left twenty SMs idle)
// is the default; RB = 128 is the round-1 shape (FABLE_B200_BJ_APPLY_ROWS=128).
template <int RB>
__global__ void __launch_bounds__(RB * 2)
k_bj_apply(double* __restrict__ A, int64_t ld, int round, int NB, const double* __restrict__ Rall) {
  constexpr int NT = RB * 2, AS = RB + 4;               // column stride of the A tile: 4 mod 16 (conflict-free fragments)
  extern __shared__ __align__(16) double sm[];
  double* TA = sm;                    // [64 cols a][AS]  (RB rows each)
  double* TR = sm + PW * AS;          // [64 cols c][RSTR]  (64 entries a each): R[a][c]
  pdl_chain();
  int bi, bj;
  bj_pair(blockIdx.x, round, NB, bi, bj);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int64_t i0 = static_cast<int64_t>(blockIdx.y) * RB;
  const double* R = Rall + static_cast<int64_t>(blockIdx.x) * (PW * PW);
  for (int idx = tid; idx < PW * (RB / 2); idx += NT) {    // A tile: 64 columns x RB / 2 chunks of 2 rows
    const int a = idx / (RB / 2), ch = idx % (RB / 2);
    if (i0 + 2 * ch < ld) cp16(TA + a * AS + 2 * ch, A + bj_col(a, bi, bj) * ld + i0 + 2 * ch);
  }
  for (int idx = tid; idx < PW * 32; idx += NT) {          // R: 64 columns x 32 chunks
    const int cc = idx >> 5, ch = idx & 31;
    cp16(TR + cc * RSTR + 2 * ch, R + cc * PW + 2 * ch);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const int wr = (warp % (RB / 32)) * 32, wc = (warp / (RB / 32)) * 32;
  double c[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { c[i][j][0] = 0.0; c[i][j][1] = 0.0; }
#pragma unroll 4
  for (int ks = 0; ks < PW / 4; ++ks) {
    double af[4], bf[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) af[i] = TA[(ks * 4 + t) * AS + wr + i * 8 + g];        // A[m = row][k = a]
#pragma unroll
    for (int j = 0; j < 4; ++j) bf[j] = TR[(wc + j * 8 + g) * RSTR + ks * 4 + t];      // B[k = a][n = c] = R[a][c]
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma4(c[i][j], af[i], bf[j]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int64_t row = i0 + wr + i * 8 + g;
        if (row < ld) A[bj_col(wc + j * 8 + 2 * t + e, bi, bj) * ld + row] = c[i][j][e];
      }
}

// ---------------------------------------------------------------------------------------------------
// k_bj_round: one round of the block Jacobi as ONE kernel (m <= 8 * 240 rows) -- a thread-block cluster of CL = 8 CTAs per
// block pair.  CTA r keeps rows [r rps, (r + 1) rps) of the pair's 64-column panel in shared memory for the whole round:
//   1. panel slab -> shared memory (cp.async), partial Gram of the slab (DMMA) -> global scratch
//   2. cluster barrier; every CTA adds one eighth of the 64 x 64 entries over the eight partial Grams, in CTA order
//   3. cluster barrier; CTA 0 runs the inner solve (bj_solve: rotations recorded, or R for the full ordering of round 0)
//   4. cluster barrier; every CTA rebuilds R from the 16 KB of recorded rotations (one warp per two rows of R) or fetches
//      it, and updates its slab in place: W <- W R (DMMA) -- the panel is read once and written once per round.
// The cluster is used for what it guarantees -- the eight CTAs are co-resident and meet at a hardware barrier with
// release / acquire semantics; the exchanged data (32 KB partial Grams, rotations) goes through L2, not through
// distributed shared memory (17-21 B per cycle and CTA, measured in B300_MICROARCH.md: slower than L2 for a one-to-seven
// broadcast).  Replaces four dependent launches per round (k_bj_gram, k_bj_eig, k_bj_replay, k_bj_apply).
constexpr int CL = 8;
__global__ void __launch_bounds__(EIGT, 1)
k_bj_round(double* __restrict__ A, int64_t ld, int rps, int round, int NB, double tol, double floor2, int inner, int cross_only,
           double* __restrict__ Hwork, double2* __restrict__ rots, double* __restrict__ Rall, int* __restrict__ mode,
           int* __restrict__ nrot, int iso) {
  constexpr int S = PW + 1;
  extern __shared__ __align__(16) double dsm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int pair = blockIdx.x / CL, r = static_cast<int>(cluster.block_rank());
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
  int bi, bj;
  bj_pair(pair, round, NB, bi, bj);
#ifdef BJ_LAB
  const long long lab_t0 = clock64();
#endif
  double* T = dsm + 3 * PW * S;                              // slab: 64 columns x rps rows, column stride TS = 4 mod 16
  const int TS = rps + 4;
  const int64_t row0 = static_cast<int64_t>(r) * rps;
  {
    const int cpc = rps >> 1;                                // 16-byte chunks per column
    for (int idx = tid; idx < PW * cpc; idx += EIGT) {
      const int a = idx / cpc, ch = idx - a * cpc;
      double* dst = T + a * TS + 2 * ch;
      if (row0 + 2 * ch < ld) cp16(dst, A + bj_col(a, bi, bj) * ld + row0 + 2 * ch);
      else { dst[0] = 0.0; dst[1] = 0.0; }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
#ifdef BJ_LAB
  const long long lab_t1 = clock64();
#endif
  double* Hp = Hwork + static_cast<int64_t>(pair) * (CL + 1) * (PW * PW);
  {
    // partial Gram, upper 8 x 8 tiles only (36 of 64; the reduction below mirrors them): tile q -> warp q % 32, so every
    // scheduler gets nine tiles; four K interleaves per tile keep the FP64 tensor pipe fed from one warp
    double* P = Hp + static_cast<int64_t>(r) * (PW * PW);
    for (int q = w; q < 36; q += EIGT / 32) {
      int ti = 0, rem = q;
      while (rem >= 8 - ti) { rem -= 8 - ti; ++ti; }
      const int a0 = ti * 8, b0 = (ti + rem) * 8;
      double c[4][2] = {};
      const double* Ta = T + (a0 + g) * TS + t;
      const double* Tb = T + (b0 + g) * TS + t;
      for (int k0 = 0; k0 < rps; k0 += 16) {
#pragma unroll
        for (int u = 0; u < 4; ++u) dmma4(c[u], Ta[k0 + 4 * u], Tb[k0 + 4 * u]);
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) P[(a0 + g) + PW * (b0 + 2 * t + e)] = (c[0][e] + c[1][e]) + (c[2][e] + c[3][e]);
    }
  }
#ifdef BJ_LAB
  const long long lab_t2 = clock64();
#endif
  cluster.sync();
#ifdef BJ_LAB
  const long long lab_t3 = clock64();
#endif
  if (tid < PW * PW / CL) {
    const int idx = r * (PW * PW / CL) + tid;
    const int ea = idx & (PW - 1), eb = idx >> 6;            // entry (ea, eb); tiles below the diagonal were not formed
    const int src = (ea >> 3) > (eb >> 3) ? eb + PW * ea : idx;
    double x[CL];
#pragma unroll
    for (int u = 0; u < CL; ++u) x[u] = __ldcg(Hp + u * (PW * PW) + src);
    double sum = x[0];
#pragma unroll
    for (int u = 1; u < CL; ++u) sum += x[u];
    Hp[CL * (PW * PW) + idx] = sum;
  }
  cluster.sync();
#ifdef BJ_LAB
  const long long lab_t4 = clock64();
#endif
  if (r == 0) {
    double* H = dsm;
    double* R = dsm + PW * S;
    const double* Hred = Hp + CL * (PW * PW);
    double v[PW * PW / EIGT];
#pragma unroll
    for (int e = 0; e < PW * PW / EIGT; ++e) v[e] = __ldcg(Hred + tid + e * EIGT);
#pragma unroll
    for (int e = 0; e < PW * PW / EIGT; ++e) {
      const int idx = tid + e * EIGT, a = idx & (PW - 1), b = idx >> 6;
      H[a * S + b] = v[e];
      if (!cross_only) R[a * S + b] = (a == b) ? 1.0 : 0.0;
    }
    const int md = bj_solve<true>(dsm, tol, floor2, inner, cross_only, nrot, rots + static_cast<int64_t>(pair) * (BW * BW), iso);
    if (md == 2) {
      double* Rg = Rall + static_cast<int64_t>(pair) * (PW * PW);
      for (int idx = tid; idx < PW * PW; idx += EIGT) {
        const int a = idx & (PW - 1), b = idx >> 6;
        Rg[a + PW * b] = R[a * S + b];
      }
    }
    if (tid == 0) mode[pair] = md;
  }
#ifdef BJ_LAB
  const long long lab_t5 = clock64();
#endif
  cluster.sync();
#ifdef BJ_LAB
  const long long lab_t6 = clock64();
#endif
  const int md = __ldcg(mode + pair);
  if (md == 0) return;                                       // R = I: the panel stays as it is
  double* Rt = dsm;                                          // Rt[c * RSTR + a] = R[a][c] (the solve buffers are dead now)
  if (md == 2) {
    const double* Rg = Rall + static_cast<int64_t>(pair) * (PW * PW);
    for (int idx = tid; idx < PW * PW; idx += EIGT) Rt[(idx >> 6) * RSTR + (idx & (PW - 1))] = __ldcg(Rg + idx);
  } else {
    // every CTA of the cluster replays eight rows of R (one warp each; lane i holds columns i (p) and 32 + (i + step) % 32
    // (q), as k_bj_replay), the cluster meets once more and everybody fetches the 64 x 64 result
    double2* rs = reinterpret_cast<double2*>(dsm + 2 * PW * S);
    rs[tid] = __ldcg(rots + static_cast<int64_t>(pair) * (BW * BW) + tid);       // BW * BW == EIGT
    __syncthreads();
    double* Rg = Rall + static_cast<int64_t>(pair) * (PW * PW);
    if (w < PW / CL) {
      const int row = r * (PW / CL) + w;
      double p_ = row == lane ? 1.0 : 0.0, q_ = row == BW + lane ? 1.0 : 0.0;
      double2 cs_ = rs[lane];
#pragma unroll 8
      for (int s = 0; s < BW; ++s) {
        const double2 nx = rs[((s + 1) & (BW - 1)) * BW + lane];
        const double np = cs_.x * p_ - cs_.y * q_, nq = cs_.y * p_ + cs_.x * q_;
        p_ = np;
        q_ = __shfl_sync(0xffffffffu, nq, (lane + 1) & 31);
        cs_ = nx;
      }
      Rg[row + PW * lane] = p_;
      Rg[row + PW * (BW + lane)] = q_;
    }
    cluster.sync();
    for (int idx = tid; idx < PW * PW; idx += EIGT) Rt[(idx >> 6) * RSTR + (idx & (PW - 1))] = __ldcg(Rg + idx);
  }
  __syncthreads();
#ifdef BJ_LAB
  const long long lab_t7 = clock64();
#endif
  {
    const int wc = (w & 1) * 32;
    for (int rt = w >> 1; rt < (rps >> 3); rt += EIGT / 64) {
      double c[4][2] = {};
      const double* Ta = T + t * TS + rt * 8 + g;
      const double* Rb = Rt + (wc + g) * RSTR + t;
#pragma unroll 4
      for (int ks = 0; ks < PW / 4; ++ks) {
        const double af = Ta[ks * 4 * TS];
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma4(c[j], af, Rb[j * 8 * RSTR + ks * 4]);
      }
      const int64_t row = row0 + rt * 8 + g;
      if (row < ld) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) A[bj_col(wc + j * 8 + 2 * t + e, bi, bj) * ld + row] = c[j][e];
      }
    }
  }
#ifdef BJ_LAB
  if (pair == 0 && tid == 0 && (r == 0 || r == 3) && atomicAdd(&bj_lab_count, 1) / 6 == 200)
    printf("BJ_LAB(round) cta %d: load %lld gram %lld sync1 %lld reduce+sync2 %lld solve %lld sync3 %lld buildR %lld apply %lld total %lld\n", r,
           lab_t1 - lab_t0, lab_t2 - lab_t1, lab_t3 - lab_t2, lab_t4 - lab_t3, lab_t5 - lab_t4, lab_t6 - lab_t5, lab_t7 - lab_t6,
           clock64() - lab_t7, clock64() - lab_t0);
#endif
}

__global__ void __launch_bounds__(256) k_bj_copy_in(const double* __restrict__ G, int64_t m, double* __restrict__ W, int64_t ld, int64_t ncols) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x, j = blockIdx.y;
  if (i >= ld || j >= ncols) return;
  W[i + ld * j] = (i < m && j < m) ? G[i + m * j] : 0.0;
}

// ---- completion of an orthonormal basis (deflated eigen / singular vectors) ---------------------------------
// lev[r] = sum_{i < done} Q[r, i]^2: e_r keeps 1 - lev[r] of its length after projection on the complement
__global__ void __launch_bounds__(256) k_leverage(const double* __restrict__ Q, int64_t rows, int64_t ld, int64_t done, double* __restrict__ lev) {
  const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (r >= rows) return;
  double a = 0.0;
  for (int64_t i = 0; i < done; ++i) { const double x = Q[r + ld * i]; a = fma(x, x, a); }
  lev[r] = a;
}
__global__ void __launch_bounds__(256) k_unit_vector(double* __restrict__ v, int64_t rows, int64_t j) {
  const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (r < rows) v[r] = r == j ? 1.0 : 0.0;
}
// coef[i] = Q[:, i] . v   (one CTA per column)
__global__ void __launch_bounds__(JT) k_proj_coef(const double* __restrict__ Q, int64_t rows, int64_t ld, const double* __restrict__ v,
                                                  double* __restrict__ coef) {
  const double* x = Q + ld * blockIdx.x;
  double a = 0.0, b = 0.0, c = 0.0;
  for (int64_t r = threadIdx.x; r < rows; r += JT) a = fma(x[r], v[r], a);
  block_sum3(a, b, c);
  if (threadIdx.x == 0) coef[blockIdx.x] = a;
}
// v -= Q[:, :done] coef
__global__ void __launch_bounds__(256) k_proj_sub(const double* __restrict__ Q, int64_t rows, int64_t ld, int64_t done,
                                                  const double* __restrict__ coef, double* __restrict__ v) {
  const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (r >= rows) return;
  double a = v[r];
  for (int64_t i = 0; i < done; ++i) a = fma(-coef[i], Q[r + ld * i], a);
  v[r] = a;
}
__global__ void __launch_bounds__(JT) k_unit_norm(double* __restrict__ v, int64_t rows) {
  double a = 0.0, b = 0.0, c = 0.0;
  for (int64_t r = threadIdx.x; r < rows; r += JT) a = fma(v[r], v[r], a);
  block_sum3(a, b, c);
  const double inv = a > 0.0 ? 1.0 / sqrt(a) : 0.0;
  for (int64_t r = threadIdx.x; r < rows; r += JT) v[r] *= inv;
}

}  // namespace

// Columns [done, total) of Q (rows x total, ld) are replaced by unit vectors orthogonal to columns [0, done) and to each
// other: column c starts from the coordinate vector that is furthest from the span of the columns before it and is
// projected three times (classical Gram-Schmidt with re-orthogonalisation).  Used for the singular / eigen vectors of
// numerically zero values, which the Gram route leaves undetermined (LAPACK returns an orthonormal basis there).
int fb_complete_orthonormal(fable_ctx* ctx, double* dQ, int64_t rows, int64_t ld, int64_t done, int64_t total) {
  if (done >= total) return FABLE_OK;
  FB_REQUIRE(ctx, total <= rows, "orthonormal completion: more columns than rows");
  DevBuf lev, coef;
  FB_CUDA(ctx, lev.alloc(static_cast<size_t>(rows) * sizeof(double)));
  FB_CUDA(ctx, coef.alloc(static_cast<size_t>(total) * sizeof(double)));
  std::vector<double> h(rows);
  const unsigned gr = static_cast<unsigned>(ceil_div(rows, 256));
  for (int64_t c = done; c < total; ++c) {
    double* v = dQ + ld * c;
    k_leverage<<<gr, 256, 0, ctx->stream>>>(dQ, rows, ld, c, lev.d());
    FB_LAUNCHED(ctx);
    FB_CUDA(ctx, cudaMemcpyAsync(h.data(), lev.d(), rows * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int64_t j = 0;
    for (int64_t r = 1; r < rows; ++r) if (h[r] < h[j]) j = r;
    k_unit_vector<<<gr, 256, 0, ctx->stream>>>(v, rows, j);
    FB_LAUNCHED(ctx);
    if (c > 0)
      for (int pass = 0; pass < 3; ++pass) {
        k_proj_coef<<<static_cast<unsigned>(c), JT, 0, ctx->stream>>>(dQ, rows, ld, v, coef.d());
        k_proj_sub<<<gr, 256, 0, ctx->stream>>>(dQ, rows, ld, c, coef.d(), v);
        k_unit_norm<<<1, JT, 0, ctx->stream>>>(v, rows);
        ctx->launches += 3;
      }
    FB_CUDA(ctx, cudaGetLastError());
  }
  return FABLE_OK;
}

namespace {
// <<<>>> with the programmatic-stream-serialization attribute (pdl != 0): see pdl_chain()
template <typename... KArgs, typename... Args>
cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
}  // namespace

int fb_syevj(fable_ctx* ctx, double* dG, int64_t m, double* dW, double* dQ, int* sweeps_out, int64_t* deflated_out) {
  FB_REQUIRE(ctx, m > 0 && m < (1 << 30), "syevj: bad dimension");
  const int N = static_cast<int>(m + (m & 1));
  const double eps = 2.220446049250313e-16;
  const double tol = std::sqrt(static_cast<double>(m)) * eps * (getenv("FABLE_B200_BJ_TOL") ? atof(getenv("FABLE_B200_BJ_TOL")) : 1.0);
  DevBuf cnt, nrm, perm;
  FB_CUDA(ctx, cnt.alloc(sizeof(int)));
  FB_CUDA(ctx, nrm.alloc(m * sizeof(double)));
  FB_CUDA(ctx, perm.alloc(m * sizeof(int)));
  std::vector<double> h(m);
  // Noise floor: the working columns converge to lambda_i q_i, so a column whose norm is below ~ m eps lambda_1 belongs
  // to a numerically zero eigenvalue (the Gram matrix of column-centred data is exactly singular, reference cpp:296).  Such
  // columns are not rotated any more (their inner products are rounding noise) and their eigenpairs are deflated below.
  // lambda_1 is bounded by the largest column norm of G from below up to sqrt(m) and from above by it: use that norm.
  k_colnorm<<<static_cast<unsigned>(m), JT, 0, ctx->stream>>>(dG, m, m, nrm.d());
  FB_LAUNCHED(ctx);
  FB_CUDA(ctx, cudaMemcpyAsync(h.data(), nrm.d(), m * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  double gmax = 0.0;
  for (double x : h) gmax = x > gmax ? x : gmax;
  const double floor1 = 4.0 * static_cast<double>(m) * eps * gmax, floor2 = floor1 * floor1;
  int sweeps = 0;
  const int max_sweeps = 40;
  const int64_t kBlockMin = getenv("FABLE_B200_BJ_MIN") ? atoll(getenv("FABLE_B200_BJ_MIN")) : 16;      // block path from m = 17 on: 2x the scalar rounds even at m = 64 (measured)
  DevBuf Wbuf;
  double* Acols = dG;          // matrix whose columns end up as lambda_i q_i, leading dimension lda
  int64_t lda = m;
  if (m > kBlockMin) {
    const int NB = static_cast<int>(ceil_div(m, BW) + (ceil_div(m, BW) & 1));     // even number of blocks
    const int64_t ld = round_up(m, 16), ncols = static_cast<int64_t>(NB) * BW;
    const int npairs = NB / 2;
    const int inner = getenv("FABLE_B200_BJ_INNER") ? atoi(getenv("FABLE_B200_BJ_INNER")) : 1;
    const int cross = getenv("FABLE_B200_BJ_CROSS") ? atoi(getenv("FABLE_B200_BJ_CROSS")) : 1;
    const int64_t splitf = getenv("FABLE_B200_BJ_SPLITF") ? atoll(getenv("FABLE_B200_BJ_SPLITF")) : 8;   // CTAs of the panel Gram per SM: its 16-row k loop is latency-bound (n = 5000: 856 -> 722 ms)
    int64_t nsplit = ceil_div(splitf * static_cast<int64_t>(ctx->sms), npairs);
    if (nsplit > ld / 64) nsplit = ld / 64 > 0 ? ld / 64 : 1;
    const int64_t rows_per_split = round_up(ceil_div(ld, nsplit), GKT);
    nsplit = ceil_div(ld, rows_per_split);
    // cross rounds record their rotations and k_bj_replay rebuilds R (FABLE_B200_BJ_REPLAY=0: R carried inside k_bj_eig)
    const bool replay = cross && inner == 1 && !(getenv("FABLE_B200_BJ_REPLAY") && atoi(getenv("FABLE_B200_BJ_REPLAY")) == 0);
    const int iso = getenv("FABLE_B200_BJ_ISO") ? atoi(getenv("FABLE_B200_BJ_ISO")) : 0;
    // programmatic dependent launch of the chain: off by default -- measured n = 1000 24.9 vs 16.6 ms, n = 2000 68.8 vs 61.2 ms
    // (the time of plain launches without the graph: the early-resident successors do not pay for themselves here)
    const int pdl = getenv("FABLE_B200_BJ_PDL") ? atoi(getenv("FABLE_B200_BJ_PDL")) : 0;
    DevBuf Hpart, Rall, Rots, Mode;
    FB_CUDA(ctx, Rots.alloc(static_cast<size_t>(npairs) * BW * BW * sizeof(double2)));
    FB_CUDA(ctx, Mode.alloc(static_cast<size_t>(npairs) * sizeof(int)));
    FB_CUDA(ctx, Wbuf.alloc(static_cast<size_t>(ld) * ncols * sizeof(double)));
    FB_CUDA(ctx, Hpart.alloc(static_cast<size_t>(npairs) * nsplit * PW * PW * sizeof(double)));
    FB_CUDA(ctx, Rall.alloc(static_cast<size_t>(npairs) * PW * PW * sizeof(double)));
    k_bj_copy_in<<<dim3(static_cast<unsigned>(ceil_div(ld, 256)), static_cast<unsigned>(ncols)), 256, 0, ctx->stream>>>(
        dG, m, Wbuf.d(), ld, ncols);
    FB_LAUNCHED(ctx);
    const int apply_rows = getenv("FABLE_B200_BJ_APPLY_ROWS") && atoi(getenv("FABLE_B200_BJ_APPLY_ROWS")) == 128 ? 128 : 64;
    const size_t apply_smem = static_cast<size_t>(PW * (apply_rows + 4) + PW * RSTR) * sizeof(double);
    const size_t eig_smem = static_cast<size_t>(3 * PW * (PW + 1)) * sizeof(double);       // H, R, second H of the pipelined cross steps
    static bool attr_dev[64] = {};
    bool& attr = attr_dev[ctx->device & 63];
    if (!attr) {
      FB_CUDA(ctx, cudaFuncSetAttribute(k_bj_apply<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>((PW * (64 + 4) + PW * RSTR) * sizeof(double))));
      FB_CUDA(ctx, cudaFuncSetAttribute(k_bj_apply<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>((PW * (128 + 4) + PW * RSTR) * sizeof(double))));
      FB_CUDA(ctx, cudaFuncSetAttribute(k_bj_eig<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(eig_smem)));
      FB_CUDA(ctx, cudaFuncSetAttribute(k_bj_eig<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(eig_smem)));
      attr = true;
    }
    // One sweep = NB - 1 rounds x (panel Gram, inner solve, panel update): up to ~470 small launches at m = 5000, identical
    // from sweep to sweep.  The sweep is captured once into a CUDA graph and replayed, which removes the per-launch host
    // cost and the gaps between the dependent kernels (FABLE_B200_BJ_GRAPH=0: plain launches).
    // Fused round (k_bj_round, a cluster of CL CTAs per pair) when the slabs fit: rps rows per CTA, CL * rps >= ld
    const int rps = static_cast<int>(round_up(ceil_div(ld, CL), 16));
    const size_t round_smem = (static_cast<size_t>(3) * PW * (PW + 1) + static_cast<size_t>(PW) * (rps + 4)) * sizeof(double);
    const int want_fused = getenv("FABLE_B200_BJ_FUSED") ? atoi(getenv("FABLE_B200_BJ_FUSED")) : 1;      // 2: also with a second wave
    bool fused = replay && rps <= 240 && want_fused != 0;
    DevBuf Hwork;
    cudaLaunchConfig_t rcfg = {};
    cudaLaunchAttribute rattr[1];
    if (fused) {
      rcfg.gridDim = dim3(static_cast<unsigned>(npairs) * CL);
      rcfg.blockDim = dim3(EIGT);
      rcfg.dynamicSmemBytes = round_smem;
      rcfg.stream = ctx->stream;
      rattr[0].id = cudaLaunchAttributeClusterDimension;
      rattr[0].val.clusterDim.x = CL; rattr[0].val.clusterDim.y = 1; rattr[0].val.clusterDim.z = 1;
      rcfg.attrs = rattr; rcfg.numAttrs = 1;
      int nclusters = 0;
      if (cudaFuncSetAttribute(k_bj_round, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(round_smem)) != cudaSuccess ||
          cudaOccupancyMaxActiveClusters(&nclusters, k_bj_round, &rcfg) != cudaSuccess || (nclusters < npairs && want_fused != 2) || nclusters < 1) {
        // (a second wave of clusters doubles the round: B200 holds 15 clusters of 8 such CTAs, m = 1000 has 16 pairs)
        (void)cudaGetLastError();
        fused = false;
      }
    }
    if (fused) FB_CUDA(ctx, Hwork.alloc(static_cast<size_t>(npairs) * (CL + 1) * PW * PW * sizeof(double)));
    auto enqueue_sweep = [&]() -> cudaError_t {
      cudaError_t e = cudaMemsetAsync(cnt.ptr, 0, sizeof(int), ctx->stream);
      if (e != cudaSuccess) return e;
      if (fused) {
        for (int r = 0; r < NB - 1; ++r) {
          e = cudaLaunchKernelEx(&rcfg, k_bj_round, Wbuf.d(), ld, rps, r, NB, tol, floor2, inner, (cross && r > 0) ? 1 : 0, Hwork.d(),
                                 static_cast<double2*>(Rots.ptr), Rall.d(), static_cast<int*>(Mode.ptr), static_cast<int*>(cnt.ptr), iso);
          if (e != cudaSuccess) return e;
        }
        return cudaGetLastError();
      }
      for (int r = 0; r < NB - 1; ++r) {
        const int cross_only = (cross && r > 0) ? 1 : 0;
        const dim3 ggrid(npairs, static_cast<unsigned>(nsplit)), agrid(npairs, static_cast<unsigned>(ceil_div(ld, apply_rows)));
        // (the first kernel of a sweep follows a memset: plain dependency)
        e = launch_chain(k_bj_gram, ggrid, dim3(128), 0, ctx->stream, pdl && r > 0, Wbuf.d(), ld, rows_per_split, ld, r, NB, Hpart.d());
        if (e != cudaSuccess) return e;
        if (replay && cross_only) {
          e = launch_chain(k_bj_eig<true>, dim3(npairs), dim3(EIGT), eig_smem, ctx->stream, pdl, Hpart.d(), static_cast<int>(nsplit), tol, floor2, inner, 1,
                           Rall.d(), static_cast<int*>(cnt.ptr), static_cast<double2*>(Rots.ptr), static_cast<int*>(Mode.ptr), iso);
          if (e != cudaSuccess) return e;
          e = launch_chain(k_bj_replay, dim3(npairs, PW / 8), dim3(256), 0, ctx->stream, pdl, static_cast<const double2*>(Rots.ptr),
                           static_cast<const int*>(Mode.ptr), Rall.d());
        } else {
          e = launch_chain(k_bj_eig<false>, dim3(npairs), dim3(EIGT), eig_smem, ctx->stream, pdl, Hpart.d(), static_cast<int>(nsplit), tol, floor2, inner,
                           cross_only, Rall.d(), static_cast<int*>(cnt.ptr), static_cast<double2*>(nullptr), static_cast<int*>(nullptr), iso);
        }
        if (e != cudaSuccess) return e;
        e = apply_rows == 64
                ? launch_chain(k_bj_apply<64>, agrid, dim3(128), apply_smem, ctx->stream, pdl, Wbuf.d(), ld, r, NB, static_cast<const double*>(Rall.d()))
                : launch_chain(k_bj_apply<128>, agrid, dim3(256), apply_smem, ctx->stream, pdl, Wbuf.d(), ld, r, NB, static_cast<const double*>(Rall.d()));
        if (e != cudaSuccess) return e;
      }
      return cudaGetLastError();
    };
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    const bool want_graph = !(getenv("FABLE_B200_BJ_GRAPH") && atoi(getenv("FABLE_B200_BJ_GRAPH")) == 0);
    if (want_graph && cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      const cudaError_t e1 = enqueue_sweep();
      const cudaError_t e2 = cudaStreamEndCapture(ctx->stream, &graph);
      if (e1 != cudaSuccess || e2 != cudaSuccess || cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) {
        (void)cudaGetLastError();
        if (gexec) { cudaGraphExecDestroy(gexec); gexec = nullptr; }
      }
    } else {
      (void)cudaGetLastError();
    }
    struct GraphGuard {
      cudaGraph_t g; cudaGraphExec_t x;
      ~GraphGuard() { if (x) cudaGraphExecDestroy(x); if (g) cudaGraphDestroy(g); }
    } graph_guard{graph, gexec};
    for (; sweeps < max_sweeps; ++sweeps) {
      if (gexec) FB_CUDA(ctx, cudaGraphLaunch(gexec, ctx->stream));
      else FB_CUDA(ctx, enqueue_sweep());
      ctx->launches += fused ? (NB - 1) : (replay ? 4 : 3) * (NB - 1) - (replay ? 1 : 0);
      int nrot = 0;
      FB_CUDA(ctx, cudaMemcpyAsync(&nrot, cnt.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      if (nrot == 0) { ++sweeps; break; }
    }
    if (sweeps >= max_sweeps) return ctx->fail(FABLE_ERR_NUMERIC, "syevj: block Jacobi sweeps did not converge");
    Acols = Wbuf.d(); lda = ld;
  } else if (m > 1) {
    for (; sweeps < max_sweeps; ++sweeps) {
      FB_CUDA(ctx, cudaMemsetAsync(cnt.ptr, 0, sizeof(int), ctx->stream));
      for (int r = 0; r < N - 1; ++r) {
        k_jacobi_round<<<N / 2, JT, 0, ctx->stream>>>(dG, m, m, r, N, tol, floor2, static_cast<int*>(cnt.ptr));
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
  k_colnorm<<<static_cast<unsigned>(m), JT, 0, ctx->stream>>>(Acols, m, lda, nrm.d());
  FB_LAUNCHED(ctx);
  FB_CUDA(ctx, cudaMemcpyAsync(h.data(), nrm.d(), m * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  std::vector<int> order(m);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h[a] > h[b]; });
  FB_CUDA(ctx, cudaMemcpyAsync(perm.ptr, order.data(), m * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  k_gather_normalise<<<static_cast<unsigned>(m), JT, 0, ctx->stream>>>(Acols, m, lda, static_cast<const int*>(perm.ptr),
                                                                      nrm.d(), dQ, dW);
  FB_LAUNCHED(ctx);
  // Deflation: eigenvalues at the noise floor are set to exactly 0 and their vectors -- normalised rounding noise at this
  // point -- are replaced by an orthonormal completion of the others.
  int64_t good = m;
  const double top = h[order[0]];
  while (good > 0 && h[order[good - 1]] <= 4.0 * floor1 + 16.0 * static_cast<double>(m) * eps * top) --good;
  if (good < m) {
    FB_CUDA(ctx, cudaMemsetAsync(dW + good, 0, static_cast<size_t>(m - good) * sizeof(double), ctx->stream));
    FB_TRY(fb_complete_orthonormal(ctx, dQ, m, m, good, m));
  }
  FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (sweeps_out) *sweeps_out = sweeps;
  if (deflated_out) *deflated_out = m - good;
  return FABLE_OK;
}
