// "fable-philox-v1": the library's own counter-based random stream.
// Philox4x32-10 keyed by the 64-bit context seed; counter = (row j, draw m low, slot, draw m high),
// so every variate is a pure function of (seed, draw, row, slot): independent of batch size, GPU
// count and tile schedule.  Replaces arma::randg / fill::randn of the reference sampler loop
// (updated-FABLE-functions.cpp:604, 610), whose R/Armadillo stream is not reproducible outside R.
#pragma once
#include <cstdint>

namespace fbrng {

constexpr uint32_t kSlotGamma = 0x10000u;

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// uniform in (0,1): top 53 bits of hi:lo, centred.
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
  const uint64_t x = ((static_cast<uint64_t>(hi) << 32) | lo) >> 11;
  return (static_cast<double>(x) + 0.5) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ void block(uint64_t seed, uint64_t m, uint32_t j, uint32_t slot,
                                      uint32_t (&w)[4]) {
  philox4x32_10(j, static_cast<uint32_t>(m), slot, static_cast<uint32_t>(m >> 32),
                static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), w);
}

// Box-Muller pair from one Philox block.
__device__ __forceinline__ void normal_pair(uint64_t seed, uint64_t m, uint32_t j, uint32_t slot,
                                            double& z0, double& z1) {
  uint32_t w[4];
  block(seed, m, j, slot, w);
  const double u1 = u53(w[0], w[1]), u2 = u53(w[2], w[3]);
  const double r = sqrt(-2.0 * log(u1));
  double s, c;
  sincos(6.283185307179586476925 * u2, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

// Marsaglia-Tsang (2000) Gamma(shape a >= 1, scale 1); attempt t consumes slots kSlotGamma+2t
// (normal) and kSlotGamma+2t+1 (uniform).
__device__ __forceinline__ double gamma_mt(uint64_t seed, uint64_t m, uint32_t j, double a) {
  const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (uint32_t t = 0;; ++t) {
    double x, unused;
    normal_pair(seed, m, j, kSlotGamma + 2u * t, x, unused);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    uint32_t w[4];
    block(seed, m, j, kSlotGamma + 2u * t + 1u, w);
    const double u = u53(w[0], w[1]);
    const double x2 = x * x;
    if (u < 1.0 - 0.0331 * x2 * x2) return d * v;
    if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) return d * v;
  }
}

}  // namespace fbrng
