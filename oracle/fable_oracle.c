/* CPU oracle, C twin -- TEST INFRASTRUCTURE ONLY (see oracle/fable_oracle.py header).
 *
 * Plain C restatement of the loops the reference runs in C++ (shounakch/FABLE,
 * src/updated-FABLE-functions.cpp; cited per function), used
 *   (1) as the checker for the CUDA path at sizes the numpy oracle is too slow for,
 *   (2) as the timed CPU baseline ("port", all host threads via OpenMP) in bench.py.
 * The product library (libfable_b200.so) never links or calls this file.
 *
 * Pinning: see oracle/fable_oracle.py (this twin is checked against it, and it against the reference's
 * own translation unit compiled with the stand-ins of oracle/shim/).
 *
 * It also holds the CPU statement of the library's OWN random-number specification
 * ("fable-philox-v1"): the reference's R/Armadillo stream cannot be reproduced, so
 * native-RNG parity is GPU-vs-this-file on the same counters, and reference parity is by
 * injected draws.
 *
 * Build: gcc -O3 -march=x86-64-v3 -fopenmp -fPIC -shared -o oracle/_build/liboracle.so oracle/fable_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11), the counter-based generator named by the north star.
 * ---------------------------------------------------------------------------------------- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* fable-philox-v1: counter = (row j, draw m low, slot, draw m high), key = 64-bit seed.
 * Uniform in (0,1): 53 bits of (w_hi:w_lo), +0.5 ulp.  Normal pair: Box-Muller on (u1,u2). */
static inline double u53(uint32_t hi, uint32_t lo) {
  uint64_t x = (((uint64_t)hi << 32) | lo) >> 11;
  return ((double)x + 0.5) * (1.0 / 9007199254740992.0);
}
#define ORC_SLOT_GAMMA 0x10000u
static inline void rng_block(uint64_t seed, uint64_t m, uint32_t j, uint32_t slot, uint32_t w[4]) {
  uint32_t ctr[4] = {j, (uint32_t)m, slot, (uint32_t)(m >> 32)};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  orc_philox4x32_10(ctr, key, w);
}
static inline void normal_pair(uint64_t seed, uint64_t m, uint32_t j, uint32_t slot, double* z0, double* z1) {
  uint32_t w[4];
  rng_block(seed, m, j, slot, w);
  double u1 = u53(w[0], w[1]), u2 = u53(w[2], w[3]);
  double r = sqrt(-2.0 * log(u1)), th = 6.283185307179586476925 * u2;
  *z0 = r * cos(th);
  *z1 = r * sin(th);
}
/* Marsaglia-Tsang (2000) for shape a >= 1, scale 1.  Attempt t uses slots GAMMA+2t (normal)
 * and GAMMA+2t+1 (uniform). */
static double gamma_mt(uint64_t seed, uint64_t m, uint32_t j, double a) {
  const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (uint32_t t = 0;; ++t) {
    double x, unused;
    normal_pair(seed, m, j, ORC_SLOT_GAMMA + 2u * t, &x, &unused);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    uint32_t w[4];
    rng_block(seed, m, j, ORC_SLOT_GAMMA + 2u * t + 1u, w);
    double u = u53(w[0], w[1]);
    double x2 = x * x;
    if (u < 1.0 - 0.0331 * x2 * x2) return d * v;
    if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) return d * v;
  }
}

/* Fill the INJECTED-layout buffers from fable-philox-v1 for draws
 * m0 .. m0+MC-1:  gam[(m-m0)*p + j],  z[(m-m0)*p*k + l*p + j]. */
void orc_rng_fill(uint64_t seed, int64_t m0, int64_t MC, int64_t p, int k, double shape,
                  double* gam, double* z) {
#pragma omp parallel for schedule(static)
  for (int64_t mm = 0; mm < MC; ++mm) {
    uint64_t m = (uint64_t)(m0 + mm);
    for (int64_t j = 0; j < p; ++j) {
      if (gam) gam[mm * p + j] = gamma_mt(seed, m, (uint32_t)j, shape);
      if (z)
        for (int l = 0; l < k; l += 2) {
          double z0, z1;
          normal_pair(seed, m, (uint32_t)j, (uint32_t)(l >> 1), &z0, &z1);
          z[mm * p * k + (int64_t)l * p + j] = z0;
          if (l + 1 < k) z[mm * p * k + (int64_t)(l + 1) * p + j] = z1;
        }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * CPPFABLESampler Monte-Carlo loop (cpp:599-618) with injected draws.
 * G0 p x k column-major; LambdaSamples MC x (k p) column-major, element (m, j*k+l) (cpp:616);
 * SigmaSqSamples MC x p column-major (cpp:606).
 * ---------------------------------------------------------------------------------------- */
void orc_sampler(int64_t MC, int64_t p, int k, const double* G0, const double* gnd, double a_n,
                 const double* gam, const double* z, double* LambdaSamples, double* SigmaSqSamples) {
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < p; ++j) {
    for (int64_t m = 0; m < MC; ++m) {
      double s2 = (0.5 * gnd[j]) / gam[m * p + j];          /* cpp:604 */
      SigmaSqSamples[m + MC * j] = s2;                       /* cpp:606 */
      double sd = sqrt(s2);                                  /* cpp:613 */
      for (int l = 0; l < k; ++l) {
        double zz = z[m * p * k + (int64_t)l * p + j] * sd;  /* cpp:613 */
        LambdaSamples[m + MC * (j * k + l)] = G0[j + p * l] + a_n * zz;   /* cpp:615-616 */
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * Entry-wise summaries of Psi_m = Lam_m Lam_m' + diag(sigma2_m) over the MC draws, following
 * the pair loop of CPPCCFABLEPostProcessing (cpp:650-688): for each j1 <= j2 form the MC-vector
 * sum_l Lam[.,j1,l]*Lam[.,j2,l] (cpp:672) (+ sigma2[.,j1] on the diagonal, cpp:676-678), then
 * reduce.  Outputs the raw sum (mean = sum/MC, cpp:688) and the raw sum of squares, mirrored to
 * the full symmetric p x p (SymmetrizeMatrix, cpp:705).  Quantiles are the f1 "next" row.
 * ---------------------------------------------------------------------------------------- */
void orc_psi_moments(int64_t MC, int64_t p, int k, const double* LambdaSamples,
                     const double* SigmaSqSamples, double* sum, double* sumsq) {
#pragma omp parallel
  {
    double* v = (double*)malloc(sizeof(double) * (size_t)MC);
#pragma omp for schedule(dynamic, 4)
    for (int64_t j1 = 0; j1 < p; ++j1) {
      const double* A = LambdaSamples + MC * (j1 * k);
      for (int64_t j2 = j1; j2 < p; ++j2) {
        const double* B = LambdaSamples + MC * (j2 * k);
        if (j2 == j1) {
          const double* S = SigmaSqSamples + MC * j1;
          for (int64_t m = 0; m < MC; ++m) v[m] = S[m];
        } else {
          for (int64_t m = 0; m < MC; ++m) v[m] = 0.0;
        }
        for (int l = 0; l < k; ++l) {
          const double* a = A + MC * l;
          const double* b = B + MC * l;
#pragma omp simd
          for (int64_t m = 0; m < MC; ++m) v[m] += a[m] * b[m];
        }
        double s1 = 0.0, s2 = 0.0;
#pragma omp simd reduction(+ : s1, s2)
        for (int64_t m = 0; m < MC; ++m) { s1 += v[m]; s2 += v[m] * v[m]; }
        sum[j1 + p * j2] = s1; sum[j2 + p * j1] = s1;
        sumsq[j1 + p * j2] = s2; sumsq[j2 + p * j1] = s2;
      }
    }
    free(v);
  }
}

/* One materialised draw Psi_m (full symmetric p x p, column-major) from the sampler outputs
 * (entry formula cpp:672-684). */
void orc_psi_draw(int64_t MC, int64_t p, int k, int64_t m, const double* LambdaSamples,
                  const double* SigmaSqSamples, double* psi) {
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t j1 = 0; j1 < p; ++j1)
    for (int64_t j2 = j1; j2 < p; ++j2) {
      double s = 0.0;
      for (int l = 0; l < k; ++l)
        s += LambdaSamples[m + MC * (j1 * k + l)] * LambdaSamples[m + MC * (j2 * k + l)];
      if (j1 == j2) s += SigmaSqSamples[m + MC * j1];
      psi[j1 + p * j2] = s; psi[j2 + p * j1] = s;
    }
}

/* ------------------------------------------------------------------------------------------
 * cov_correct_matrix reductions (R/FABLE_code.R:189-206; cpp:401-428) without storing B:
 * returns sum over the FULL matrix (wrapper formula, R/FABLE_wrapper.R:44) and over the upper
 * triangle incl. diagonal (script formula, runtimeScript.R:78).  G is p x p column-major.
 * ---------------------------------------------------------------------------------------- */
void orc_cov_correct_sums(int64_t p, const double* sig, const double* G, double* sum_full,
                          double* sum_upper) {
  double sf = 0.0, su = 0.0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : sf, su)
  for (int64_t v = 0; v < p; ++v) {
    double gv = G[v + p * v];
    for (int64_t u = 0; u <= v; ++u) {
      double gu = G[u + p * u], guv = G[u + p * v];
      double num = gu * gv + guv * guv;            /* R:196 */
      double den = gu * sig[v] + sig[u] * gv;      /* R:197 */
      double b = num / den;                        /* R:199 */
      if (u == v) b = b / 2.0;                     /* R:200 */
      b = sqrt(1.0 + b);                           /* R:202 */
      su += b;
      sf += (u == v) ? b : 2.0 * b;
    }
  }
  *sum_full = sf; *sum_upper = su;
}

/* Explicit residual column variances (cpp:161-166) in plain loops (no BLAS). */
void orc_resid_sigsq(int64_t n, int64_t p, int k, const double* Y, const double* U, int64_t ldu,
                     const double* V, int64_t ldv, const double* d, double* sig) {
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < p; ++j) {
    double acc = 0.0;
    for (int64_t i = 0; i < n; ++i) {
      double f = 0.0;
      for (int l = 0; l < k; ++l) f += U[i + ldu * l] * d[l] * V[j + ldv * l];
      double r = Y[i + n * j] - f;
      acc += r * r;
    }
    sig[j] = acc / (double)n;
  }
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the timed CPU arm sets its thread count itself */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n >= 1) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
