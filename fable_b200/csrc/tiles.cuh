// Tile enumeration of the covariance pass (host + device).
//
// The p x p symmetric output is cut into 64 x 64 tiles; only tiles (I, J) with I <= J are computed (the reference's
// pair loop j1 <= j2, src/updated-FABLE-functions.cpp:650, 662) and mirrored.  Tiles are enumerated in super-blocks of
// 16 x 16 tiles (the super-blocks over their own upper triangle, column-major; tiles column-major inside), so that
// the CTAs resident at any time share 2 x 16 operand row blocks.  The enumeration index t runs over ALL slots of the
// visited super-blocks: slots below the diagonal of a diagonal super-block or past the matrix edge are padding.
//
// Storage is COMPACT: valid tiles only, in enumeration order.  cidx(t) = number of valid tiles with enumeration index
// < t, so a tile's compact index is cidx(t) and a slice [t0, t1) of the enumeration (one rank of the tile partition,
// SURVEY 8(e) partition B) owns the contiguous compact range [cidx(t0), cidx(t1)).  A compact buffer is a column-major
// matrix of 64 rows x (64 x tiles) columns: tile c is columns [64c, 64c + 64), 32 KB contiguous.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define FB_HD __host__ __device__ __forceinline__
#else
#define FB_HD inline
#endif

namespace fbt {

constexpr int TILE = 64;
#ifndef FB_SB
#define FB_SB 16                      // lab builds only: -DFB_SB=32 (every translation unit), see DESIGN.md 3.1
#endif
constexpr int SB = FB_SB;
constexpr int TILE_DOUBLES = TILE * TILE;

// column-major enumeration of an upper triangle: t = J(J+1)/2 + I, 0 <= I <= J
FB_HD void tri_of(int64_t t, int& I, int& J) {
#ifdef __CUDA_ARCH__
  int64_t j = static_cast<int64_t>((sqrt(8.0 * static_cast<double>(t) + 1.0) - 1.0) * 0.5);
#else
  int64_t j = static_cast<int64_t>((__builtin_sqrt(8.0 * static_cast<double>(t) + 1.0) - 1.0) * 0.5);
#endif
  while (j * (j + 1) / 2 > t) --j;
  while ((j + 1) * (j + 2) / 2 <= t) ++j;
  J = static_cast<int>(j);
  I = static_cast<int>(t - j * (j + 1) / 2);
}

FB_HD int64_t n_tiles(int64_t p) { return (p + TILE - 1) / TILE; }
// length of the enumeration (padding slots included)
FB_HD int64_t n_slots(int64_t p) {
  const int64_t nT = n_tiles(p), nSB = (nT + SB - 1) / SB;
  return nSB * (nSB + 1) / 2 * SB * SB;
}
FB_HD int64_t n_valid(int64_t p) {
  const int64_t nT = n_tiles(p);
  return nT * (nT + 1) / 2;
}

// valid tiles in super-block columns 0 .. sJ-1 = tiles (I <= J) with J < min(16 sJ, nT)
FB_HD int64_t valid_before_sbcol(int64_t sJ, int nT) {
  const int64_t m = sJ * SB < nT ? sJ * SB : nT;
  return m * (m + 1) / 2;
}

// number of valid tiles with enumeration index < t  (0 <= t <= n_slots)
FB_HD int64_t cidx(int64_t t, int nT) {
  const int64_t s = t / (SB * SB);
  const int r = static_cast<int>(t % (SB * SB));
  int sI, sJ;
  tri_of(s, sI, sJ);
  const int rem = nT - sJ * SB;
  const int wJ = rem <= 0 ? 0 : (rem < SB ? rem : SB);         // rem <= 0 only for t == n_slots
  int64_t c = valid_before_sbcol(sJ, nT) + static_cast<int64_t>(sI) * SB * wJ;
  const int j = r / SB, i = r % SB;
  const int jj = j < wJ ? j : wJ;
  if (sI < sJ) c += jj * SB + (j < wJ ? i : 0);
  else c += jj * (jj + 1) / 2 + (j < wJ ? (i < j + 1 ? i : j + 1) : 0);
  return c;
}

// tile (I, J) of enumeration slot t; false for padding slots
FB_HD bool tile_of(int64_t t, int nT, int& I, int& J) {
  int sI, sJ;
  tri_of(t / (SB * SB), sI, sJ);
  const int r = static_cast<int>(t % (SB * SB));
  I = sI * SB + (r % SB);
  J = sJ * SB + (r / SB);
  return I <= J && J < nT;
}

// enumeration slot of the valid tile (I, J), I <= J
FB_HD int64_t slot_of(int I, int J) {
  const int64_t sI = I / SB, sJ = J / SB;
  return (sJ * (sJ + 1) / 2 + sI) * (SB * SB) + (J % SB) * SB + (I % SB);
}

}  // namespace fbt
