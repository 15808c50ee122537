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
//   * a finished tile is staged in shared memory in BOTH orientations (direct and mirrored, as
//     128-byte-swizzled TMA boxes) and leaves the SM as two asynchronous TMA tensor stores
//     (cp.async.bulk.tensor -> UTMASTG), so HBM writes drain while the warps already compute the
//     next draw; every global run is a full, aligned 512-byte line group written exactly once;
//   * the HBM accumulators are read into the register accumulators when the CTA starts (latency hidden
//     behind the first draw) and written back once per batch by TMA tensor stores.
// Tile slot = 64 x 64; a CTA (256 threads = 8 warps) owns 64 rows x TV columns of it: TV = 64 (2 x 4 warps of
// 32 x 16, two CTAs per SM) or TV = 32 (4 x 2 warps of 16 x 16, two CTAs per slot, three per SM); FP64 tensor-core
// tiles (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4; tcgen05 has no f64 kind), rank padded to a multiple
// of 4 with zero rows.  DMMA runs at the DFMA rate on sm_100a but needs 6 fragment loads per 8 MMAs
// (2048 FMAs) instead of 6 loads per 16 DFMAs (512 FMAs), which frees the issue slots and shared-memory
// bandwidth that the staging and the asynchronous stores need.
#include <cuda.h>   // CUtensorMap types only; the encoder is fetched through cudaGetDriverEntryPoint

#include <cstring>

#include "common.cuh"
#include "tiles.cuh"

// PSI_LAB: diagnostic builds only (tools/psi_lab.sh, profiles/r01_psi_lab.md): 1 skips the DMMA loop, 2 the operand
// prefetch, 4 the staging stores, 16 the accumulator loads, 32 makes every draw read the operands of draw 0 (L2-resident) -- results are then wrong by construction; never defined
// in the product build.
#ifndef PSI_LAB
#define PSI_LAB 0
#endif

namespace {

constexpr int TILE = fbt::TILE;
constexpr int KC_DEFAULT = 16;     // rank chunk per pipeline step (multiple of 4); 24 for 16 < k <= 24 (one chunk, no mid-draw barrier)
constexpr int OSTR = TILE + 4;     // operand row stride: 68 = 4 mod 16 -> conflict-free DMMA fragment loads
constexpr int BOXW = 16;           // TMA box row: 16 doubles = 128 bytes = one swizzle row
constexpr int BOX_DOUBLES = BOXW * TILE;   // 16 x 64 sub-box, 8 KB; 4 sub-boxes per orientation
// A CTA owns TILE rows (u) x TV columns (v) of one 64 x 64 tile slot: TV = 64 (one CTA per slot, two CTAs per
// SM) or TV = 32 (two CTAs per slot, three per SM -- more independent store streams per SM, see fb_psi).
template <int TV, int KC = KC_DEFAULT> struct Geo {
  static constexpr int VSTR = TV + 4;                                  // v-operand row stride (4 mod 16, as OSTR)
  static constexpr int BUF_DOUBLES = KC * (OSTR + VSTR);               // one operand buffer: [u rows][v rows] x KC
  static constexpr size_t OP_BYTES = 2 * BUF_DOUBLES * sizeof(double); // double-buffered: 34 KB / 26 KB
  static constexpr int STAGE_DOUBLES = TILE * TV;                      // one orientation: 32 KB / 16 KB
  static constexpr size_t SMEM_BYTES = OP_BYTES + 2 * STAGE_DOUBLES * sizeof(double) + 1024;   // + 1 KB alignment slack
  static constexpr int NWU = TV == 64 ? 2 : 4;                         // warps along u; 8 / NWU along v (16 columns each)
  static constexpr int WR = TILE / NWU;                                // warp rows: 32 / 16
  static constexpr int NI = WR / 8;                                    // 8-row DMMA tiles per warp: 4 / 2
  static constexpr int CTAS = TV == 64 ? 2 : 3;                        // resident CTAs per SM
};

// Staging layouts = the shared-memory image of 128-byte-swizzled TMA boxes: sub-box q holds 16
// consecutive rows-of-the-matrix (the contiguous index) for 64 columns; inside a sub-box, row r is 8
// chunks of 16 bytes and chunk c sits at position c ^ (r & 7).
//   direct   orientation: sub-box q = u in [16q, 16q+16), row = v
//   mirrored orientation: sub-box q = v in [16q, 16q+16), row = u
// With the DMMA accumulator layout (u = g, v = 2t+e per 8 x 8 tile) the direct 8-byte stores are
// bank-conflict-free and the mirrored 16-byte stores are 2-way conflicted (dense 512-byte rows
// would be 4-way / 2-way).
template <int TV = TILE>
__device__ __forceinline__ int d_off(int u, int v) {
  return (u >> 4) * (BOXW * TV) + v * BOXW + ((((u & 15) >> 1) ^ (v & 7)) << 1) + (u & 1);
}
__device__ __forceinline__ int m_off(int u, int v) {
  return (v >> 4) * BOX_DOUBLES + u * BOXW + ((((v & 15) >> 1) ^ (u & 7)) << 1) + (v & 1);
}

constexpr int SB = fbt::SB;        // tile order: super-blocks of SB x SB tiles (tiles.cuh)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// PSI_OP_POLICY (lab builds): L2 policy of the operand rows -- 0 evict_last (product), 1 evict_normal.  With k = 20 and 47
// draws per launch the operands (376 MB) exceed the L2, unlike at c3 (77 MB); measured: c4 share 22.45 vs 22.37 ms, c3 22.24 vs
// 21.91 ms -- evict_last stays.
#ifndef PSI_OP_POLICY
#define PSI_OP_POLICY 0
#endif
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
#if PSI_OP_POLICY == 1
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
#else
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
#endif
  return pol;
}
// operand rows are re-read by every tile of a block row/column: keep them in L2 (evict_last)
__device__ __forceinline__ void cp_async16(void* dst, const void* src, uint64_t pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// (the materialised draws are write-once streams: their stores carry evict_first so they do not displace the operands)
// The p x p x slots output is described to TMA as a 4-D tensor (r_lo = 16 contiguous rows, column,
// r_hi = row block of 16, slot) so that ONE store moves a whole 64 x 64 tile as a {16, 64, 4, 1} box whose
// shared-memory image is the four swizzled sub-boxes above (TMA clips the box at the matrix edge).
__device__ __forceinline__ void tma_store_tile(const CUtensorMap* tmap, int col, int rowblk, int slot, const double* ssrc,
                                               uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2, %3, %4}], [%5], %6;" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(0), "r"(col), "r"(rowblk), "r"(slot), "r"(smem_u32(ssrc)), "l"(pol)
               : "memory");
}
// Fallback when p is even but not a multiple of 16: the plain 3-D map (row, column, slot) and one
// {16, 64, 1} store per swizzled sub-box (four per orientation).
__device__ __forceinline__ void tma_store_subbox(const CUtensorMap* tmap, int row, int col, int slot, const double* ssrc,
                                                 uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2, %3}], [%4], %5;" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(row), "r"(col), "r"(slot), "r"(smem_u32(ssrc)), "l"(pol)
               : "memory");
}
// 64 x 64 box added (FP64, in L2) into a p x p accumulator: each entry is touched by exactly one CTA
// per launch, so the result is deterministic
__device__ __forceinline__ void tma_reduce_add_tile(const CUtensorMap* tmap, int col, int rowblk, const double* ssrc) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(0), "r"(col), "r"(rowblk), "r"(0), "r"(smem_u32(ssrc))
               : "memory");
}
// accumulator tiles come in by TMA tensor loads (UTMALDG) signalled on an mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred P1;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_tile(const CUtensorMap* tmap, int col, int rowblk, double* sdst, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                   smem_u32(sdst)),
               "l"(reinterpret_cast<uint64_t>(tmap)), "r"(0), "r"(col), "r"(rowblk), "r"(0), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// Lw: working-layout draws, rows zero-padded to pp (a multiple of TILE), rank zero-padded to KP (a
// multiple of 4), 16-byte aligned.
//
// CC (f3, R/FABLE_code.R:212-298 CCFABLE_DirectSampler): the entry-wise coverage correction
//   Psi_m = G + B o (Lambda_m Lambda_m' - G) + diag(sigma2_m),  B = cov_correct_matrix(sig_hat, G)  (R:251, 282-290).
// Slot 0 of Lw then holds G0 as a pseudo-draw (b = -1): its tile product is G, from which B is formed once
// per CTA (R:196-202) and both stay in registers (64 more: one CTA per SM).
template <bool MAT, bool ACC, bool CC, int TV, int KC = KC_DEFAULT>
__global__ void __launch_bounds__(256, CC ? 1 : Geo<TV>::CTAS)
k_psi(PsiArgs a, int64_t tile_begin, const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_m,
      const __grid_constant__ CUtensorMap tmap_sum, const __grid_constant__ CUtensorMap tmap_sq) {
  using G_ = Geo<TV, KC>;
  constexpr int NI = G_::NI, WR = G_::WR, VSTR = G_::VSTR, BUFD = G_::BUF_DOUBLES, HALVES = TILE / TV;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);   // swizzle atoms: 1 KB aligned
  double* op = reinterpret_cast<double*>(smem_raw);
  double* D = reinterpret_cast<double*>(smem_raw + G_::OP_BYTES);   // direct   staging (swizzled sub-boxes)
  double* M = D + G_::STAGE_DOUBLES;                                // mirrored staging (swizzled sub-boxes)

  int I, J;
  const int64_t tslot = tile_begin + blockIdx.x / HALVES;
  if (!fbt::tile_of(tslot, a.nT, I, J)) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wu = warp % G_::NWU, wv = warp / G_::NWU;
  const int vh = (blockIdx.x % HALVES) * TV;     // column offset of this CTA's part inside the tile
  const int64_t u0 = static_cast<int64_t>(I) * TILE, v0 = static_cast<int64_t>(J) * TILE + vh;
  const int64_t p = a.p;
  if (v0 >= p) return;                           // second half of an edge tile
  const bool diag = (I == J);
  const bool bulk = MAT && a.tma_mode != 0;      // output tensor maps exist (1: 4-D whole-tile boxes, 2: 3-D sub-boxes)
  // TMA coordinates (column, 16-row block).  Accumulators are always COMPACT (tiles.cuh): tile c = columns [64c, 64c+64)
  // of a 64-row matrix.  The materialised draws are either dense p x p matrices (direct tile at (u0, v0), mirrored at
  // (v0, u0)) or, under a tile partition, compact per rank: direct tile at 2c, mirrored at 2c + 1.
  const int cl = static_cast<int>(fbt::cidx(tslot, a.nT) - a.cbase);
  const int acol = cl * TILE + vh;
  const int dcol = a.out_compact ? 2 * cl * TILE + vh : static_cast<int>(v0);
  const int drb = a.out_compact ? 0 : static_cast<int>(u0 >> 4);
  const int mcol = a.out_compact ? (2 * cl + 1) * TILE : static_cast<int>(u0);
  const int mrb = a.out_compact ? (vh >> 4) : static_cast<int>(v0 >> 4);
  const int K4 = a.KP;
  const int nch = (K4 + KC - 1) / KC;
  const int nsteps = (a.B + (CC ? 1 : 0)) * nch;
  const uint64_t pol_keep = policy_evict_last(), pol_stream = policy_evict_first();
  double gq[CC ? NI : 1][2][2], bq[CC ? NI : 1][2][2];    // G tile and B tile of the entry-wise correction

  // threads 0..127 stage the u-operand, 128..255 the v-operand: chunk column = tid & 31 (16 bytes),
  // rank rows (tid >> 5) & 3, +4, +8, +12
  const int pf_o = tid >> 7, pf_ch = tid & 31, pf_l0 = (tid >> 5) & 3;
  const int64_t pf_col = (pf_o ? v0 : u0) + 2 * pf_ch;
  auto prefetch = [&](int s, int buf) {
    const int b = (PSI_LAB & 32) ? 0 : s / nch, l0 = (s - s / nch * nch) * KC;
    const int kc = nch == 1 ? a.k : min(KC, K4 - l0);      // single chunk: the zero pad rows are never fetched
    if (TV == TILE) {
      const double* src = a.Lw + (static_cast<int64_t>(b) * a.KP + l0 + pf_l0) * a.pp + pf_col;
      double* dst = op + buf * BUFD + (pf_o * KC + pf_l0) * OSTR + 2 * pf_ch;
#pragma unroll
      for (int l = 0; l < KC; l += 4)
        if (pf_l0 + l < kc) cp_async16(dst + l * OSTR, src + static_cast<int64_t>(l) * a.pp, pol_keep);
    } else {
      // 64 u-rows: chunk tid & 31 of rank rows tid >> 5 and + 8; TV v-rows: chunk tid & 15 of rank row tid >> 4
      const double* base = a.Lw + (static_cast<int64_t>(b) * a.KP + l0) * a.pp;
      double* dst = op + buf * BUFD;
#pragma unroll
      for (int h = 0; h < KC / 8; ++h) {
        const int l = (tid >> 5) + 8 * h;
        if (l < kc) cp_async16(dst + l * OSTR + 2 * (tid & 31), base + static_cast<int64_t>(l) * a.pp + u0 + 2 * (tid & 31), pol_keep);
      }
#pragma unroll
      for (int h = 0; h < (KC + 15) / 16; ++h) {
        const int l = (tid >> 4) + 16 * h;
        if (l < kc) cp_async16(dst + KC * OSTR + l * VSTR + 2 * (tid & 15), base + static_cast<int64_t>(l) * a.pp + v0 + 2 * (tid & 15), pol_keep);
      }
    }
    cp_async_commit();
  };

  // Materialising kernels with whole-box tensor maps: the running accumulator tiles are fetched by two TMA tensor
  // loads into the (still unused) staging buffers and folded into the registers when draw 0 is accumulated.
  __shared__ uint64_t acc_bar;
  const bool acc_tma = MAT && ACC && !a.acc_fresh && !(PSI_LAB & 16);
  if (acc_tma) {
    if (tid == 0) mbar_init(&acc_bar, 1);
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(&acc_bar, 2 * G_::STAGE_DOUBLES * sizeof(double));
      tma_load_tile(&tmap_sum, acol, 0, D, &acc_bar);
      tma_load_tile(&tmap_sq, acol, 0, M, &acc_bar);
    }
  }

  // fragment coordinates: c[i][j][e] is entry (u, v) = (wu*WR + i*8 + g, wv*16 + j*8 + 2t + e) of the tile
  double acc_s[NI][2][2], acc_q[NI][2][2];
  double c[NI][2][2];
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        c[i][j][e] = 0.0;
        // the running HBM accumulators (materialising kernels) arrive by TMA and are first needed when draw 0 is folded
        // in; the batch ends with a plain store (no read-modify-write, no L2 atomics).  The accumulate-only kernel is
        // FP64-bound and keeps zero-init + one TMA reduce-add per batch.
        if (ACC) { acc_s[i][j][e] = 0.0; acc_q[i][j][e] = 0.0; }
      }

  if (nch == 1 && a.k < K4) {
    // rank rows k..KP-1 are zero padding: clear them once in both operand buffers instead of re-reading zeros from
    // L2 / HBM for every draw (17 % of the operand traffic at k = 10)
    const int nz = (K4 - a.k) * (OSTR + VSTR);
    for (int idx = tid; idx < 2 * nz; idx += 256) {
      const int buf = idx / nz, r = idx - buf * nz;
      const int lu = r / OSTR;                       // first the (K4 - k) u-rows, then the v-rows
      double* base = op + buf * BUFD;
      if (lu < K4 - a.k) base[(a.k + lu) * OSTR + (r - lu * OSTR)] = 0.0;
      else {
        const int rv = r - (K4 - a.k) * OSTR, lv = rv / VSTR;
        base[KC * OSTR + (a.k + lv) * VSTR + (rv - lv * VSTR)] = 0.0;
      }
    }
  }
  prefetch(0, 0);
  cp_async_wait_all();
  __syncthreads();
  for (int s = 0; s < nsteps; ++s) {
    const int buf = s & 1;
    const int bs = s / nch, ch = s - bs * nch;
    const int b = bs - (CC ? 1 : 0);                   // CC: step group 0 is the G0 pseudo-draw
    const int kc = min(KC, K4 - ch * KC);
    if (s + 1 < nsteps && !(PSI_LAB & 2)) prefetch(s + 1, buf ^ 1);      // buf^1 was last read before the previous barrier

    const double* As = op + buf * BUFD + wu * WR + g + t * OSTR;
    const double* Bs = op + buf * BUFD + KC * OSTR + wv * 16 + g + t * VSTR;
#pragma unroll
    for (int ks = 0; ks < KC / 4; ++ks) {
      if (ks * 4 < kc && !(PSI_LAB & 1)) {
        double af[NI], bf[2];
#pragma unroll
        for (int i = 0; i < NI; ++i) af[i] = As[ks * 4 * OSTR + i * 8];
#pragma unroll
        for (int j = 0; j < 2; ++j) bf[j] = Bs[ks * 4 * VSTR + j * 8];
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j) dmma(c[i][j], af[i], bf[j]);
      }
    }
    if (ch + 1 < nch) {                  // more rank chunks of this draw to come
      cp_async_wait_all();
      __syncthreads();
      continue;
    }

    if (CC) {
      if (b < 0) {
        // c holds the G tile: keep it, form B_uv = sqrt(1 + (g_uu g_vv + g_uv^2) / (g_uu s_v + s_u g_vv)), the
        // ratio halved on the diagonal (R/FABLE_code.R:196-202)
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int64_t u = u0 + wu * WR + i * 8 + g, v = v0 + wv * 16 + j * 8 + 2 * t + e;
              double bb = 0.0;
              if (u < p && v < p) {
                const double gu = a.cc_gdiag[u], gv = a.cc_gdiag[v], su = a.cc_sig[u], sv = a.cc_sig[v];
                // products rounded before they are added (no FMA contraction): B_uv == B_vu bit for bit, as in R's
                // element-wise matrix arithmetic, so diagonal tiles stay exactly symmetric
                bb = __ddiv_rn(__dadd_rn(__dmul_rn(gu, gv), __dmul_rn(c[i][j][e], c[i][j][e])), __dadd_rn(__dmul_rn(gu, sv), __dmul_rn(su, gv)));
                if (u == v) bb = bb / 2.0;
                bb = sqrt(1.0 + bb);
              }
              gq[i][j][e] = c[i][j][e]; bq[i][j][e] = bb; c[i][j][e] = 0.0;
            }
        cp_async_wait_all();
        __syncthreads();
        continue;
      }
#pragma unroll
      for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) c[i][j][e] = gq[i][j][e] + bq[i][j][e] * (c[i][j][e] - gq[i][j][e]);   // R:285
    }

    // ---- draw b of this tile is complete in registers ----------------------------------------
    if (diag && a.s2w) {
#pragma unroll
      for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int64_t u = u0 + wu * WR + i * 8 + g, v = v0 + wv * 16 + j * 8 + 2 * t + e;
            if (u == v && u < p) c[i][j][e] += a.s2w[static_cast<int64_t>(b) * a.pp + u];
          }
    }
    if (ACC) {
      if (acc_tma && b == 0) {
        mbar_wait(&acc_bar, 0);
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int ul = wu * WR + i * 8 + g, vl = wv * 16 + j * 8 + 2 * t + e;
              acc_s[i][j][e] = D[d_off<TV>(ul, vl)];
              acc_q[i][j][e] = M[d_off<TV>(ul, vl)];
            }
      }
#pragma unroll
      for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            acc_s[i][j][e] += c[i][j][e];
            acc_q[i][j][e] = fma(c[i][j][e], c[i][j][e], acc_q[i][j][e]);
          }
    }
    if (MAT) {
      if (bulk && tid == 0) bulk_wait_read();         // the previous draw's TMA stores have read the staging
      __syncthreads();
      if (!(PSI_LAB & 4))
#pragma unroll
      for (int i = 0; i < NI; ++i) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int ul = wu * WR + i * 8 + g, vl = wv * 16 + j * 8 + 2 * t;
          D[d_off<TV>(ul, vl)] = c[i][j][0];
          D[d_off<TV>(ul, vl + 1)] = c[i][j][1];
        }
        if (!diag) {
          // mirrored 16-byte stores: a quarter-warp is rows u = g, g + 1 with the four column pairs t; stored tile by tile
          // both rows hit the same 64-byte half of their 128-byte swizzle row (the XOR with u & 7 = g flips only bit 0
          // between them): 2-way conflicts on every STS.128, 498 M excess wavefronts per launch in ncu.  Odd rows therefore
          // store their j = 1 tile while even rows store j = 0 and vice versa: the halves alternate, no conflict.
          const int ul = wu * WR + i * 8 + g;
          const bool odd = g & 1;
          const double2 t0 = make_double2(c[i][0][0], c[i][0][1]), t1 = make_double2(c[i][1][0], c[i][1][1]);
          *reinterpret_cast<double2*>(M + m_off(ul, wv * 16 + (odd ? 8 : 0) + 2 * t)) = odd ? t1 : t0;
          *reinterpret_cast<double2*>(M + m_off(ul, wv * 16 + (odd ? 0 : 8) + 2 * t)) = odd ? t0 : t1;
        }
      }
      if (bulk) fence_async_smem();      // generic-proxy smem writes -> visible to the async (bulk copy) proxy
    }
    cp_async_wait_all();
    __syncthreads();                     // staging complete; operands of step s+1 landed for every thread
    if (MAT) {
      if (bulk) {
        // two TMA tensor stores per draw (UTMASTG): the direct tile (columns v0.., row blocks u0/16..)
        // and the mirrored tile (columns u0.., row blocks v0/16..) of slot b.  Issued by one thread, so
        // back-pressure from HBM never blocks the warps at issue: it is absorbed at bulk_wait_read,
        // after the next draw has been computed.
        if (tid == 0) {
          if (a.tma_mode == 1) {
            tma_store_tile(&tmap, dcol, drb, a.slot0 + b, D, pol_stream);
            if (!diag) tma_store_tile(&tmap_m, mcol, mrb, a.slot0 + b, M, pol_stream);
          } else {
            for (int q = 0; q < TILE / BOXW; ++q) {
              if (u0 + q * BOXW < p)
                tma_store_subbox(&tmap, static_cast<int>(u0) + q * BOXW, static_cast<int>(v0), a.slot0 + b, D + q * BOX_DOUBLES, pol_stream);
              if (!diag && v0 + q * BOXW < p)
                tma_store_subbox(&tmap, static_cast<int>(v0) + q * BOXW, static_cast<int>(u0), a.slot0 + b, M + q * BOX_DOUBLES, pol_stream);
            }
          }
          bulk_commit();
        }
      } else {
        // odd p (rows not 16-byte aligned, no tensor map) -> plain coalesced stores from the staging
        // (the barrier that opens the next draw's staging keeps it intact until these reads are done)
        double* out = a.psi + static_cast<int64_t>(a.slot0 + b) * p * p;
        for (int cidx = warp; cidx < TILE; cidx += 8) {
          if (v0 + cidx < p)
            for (int h = 0; h < 2; ++h)
              if (u0 + lane + 32 * h < p) __stcs(out + (v0 + cidx) * p + u0 + lane + 32 * h, D[d_off<TV>(lane + 32 * h, cidx)]);
          if (!diag && u0 + cidx < p)
            for (int h = 0; h < 2; ++h)
              if (v0 + lane + 32 * h < p) __stcs(out + (u0 + cidx) * p + v0 + lane + 32 * h, M[m_off(cidx, lane + 32 * h)]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) { c[i][j][0] = 0.0; c[i][j][1] = 0.0; }
  }
  if (bulk && tid == 0) bulk_wait_read();             // staging must stay valid until the last stores have read it

  if (ACC) {
    // one update of the HBM accumulators per batch: the register accumulators (old value + batch) are staged as two
    // swizzled tiles and leave as plain TMA tensor stores into the compact tile-major buffers (any p: a tile's 32 KB
    // are always whole and aligned; rows / columns past p hold zeros)
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int ul = wu * WR + i * 8 + g, vl = wv * 16 + j * 8 + 2 * t;
        D[d_off<TV>(ul, vl)] = acc_s[i][j][0];
        D[d_off<TV>(ul, vl + 1)] = acc_s[i][j][1];
        M[d_off<TV>(ul, vl)] = acc_q[i][j][0];      // both accumulators use the direct orientation
        M[d_off<TV>(ul, vl + 1)] = acc_q[i][j][1];
      }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      if (MAT) {
        tma_store_tile(&tmap_sum, acol, 0, 0, D, pol_stream);
        tma_store_tile(&tmap_sq, acol, 0, 0, M, pol_stream);
      } else {
        tma_reduce_add_tile(&tmap_sum, acol, 0, D);
        tma_reduce_add_tile(&tmap_sq, acol, 0, M);
      }
      bulk_commit();
      bulk_wait_read();
    }
  }
}

// Compact tiles -> dense column panel.  out is the p x ncols panel (leading dimension p) of columns [col0, col0 + ncols)
// of the dense symmetric matrix; the CTA of dense block (R, C) reads tile (min, max) of the compact buffer -- transposed
// when R > C (`paired` = 0: accumulators, upper tiles only) or from the stored mirrored copy (paired = 1: materialised
// draws, direct tile at 2c, mirrored at 2c + 1).  Blocks whose tile lies outside the slice [t_lo, t_hi) are written as
// zeros, so the panels of the ranks of a tile partition add up to the full matrix.
__global__ void __launch_bounds__(256) k_tiles_to_dense(const double* __restrict__ tiles, int nT, int64_t t_lo, int64_t t_hi,
                                                        int64_t cbase, int paired, int64_t p, int64_t col0, int64_t ncols,
                                                        double* __restrict__ out) {
  __shared__ double sm[TILE][TILE + 1];
  const int R = blockIdx.x;                                        // dense row block
  const int64_t cb0 = col0 + static_cast<int64_t>(blockIdx.y) * TILE;  // first dense column of this block
  const int C = static_cast<int>(cb0 / TILE);                     // col0 is a multiple of 64
  const int I = R < C ? R : C, J = R < C ? C : R;
  const int64_t ts = fbt::slot_of(I, J);
  const bool own = ts >= t_lo && ts < t_hi;
  const bool lower = R > C;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;          // 64 x 4
  const double* src = nullptr;
  if (own) {
    const int64_t c = fbt::cidx(ts, nT) - cbase;
    src = tiles + (paired ? 2 * c + (lower ? 1 : 0) : c) * fbt::TILE_DOUBLES;
  }
  const bool transpose = lower && !paired;
  if (own && transpose) {
    for (int cc = ty; cc < TILE; cc += 4) sm[cc][tx] = src[cc * TILE + tx];      // sm[v][u] = tile(u, v)
    __syncthreads();
  }
  for (int cc = ty; cc < TILE; cc += 4) {
    const int64_t col = cb0 + cc, row = static_cast<int64_t>(R) * TILE + tx;
    if (col >= col0 + ncols || col >= p || row >= p) continue;
    double v = 0.0;
    if (own) v = transpose ? sm[tx][cc] : src[cc * TILE + tx];      // dense (row, col) = tile(col - .., row - ..) transposed
    out[(col - col0) * p + row] = v;
  }
}

// One super-block of compact upper tiles -> dense block(s); grid (16, 16) = (i, j) tile of the super-block.
__global__ void __launch_bounds__(256) k_superblock_to_dense(const double* __restrict__ tiles, int nT, int64_t cbase, int64_t p, int sI,
                                                             int sJ, int64_t nr, int64_t nc, double* __restrict__ direct, int64_t ldd,
                                                             double* __restrict__ mirror, int64_t ldm) {
  __shared__ double sm[TILE][TILE + 1];
  const int i = blockIdx.x, j = blockIdx.y;
  const int I = sI * SB + i, J = sJ * SB + j;
  if (I > J || J >= nT) return;
  const double* src = tiles + (fbt::cidx(fbt::slot_of(I, J), nT) - cbase) * fbt::TILE_DOUBLES;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const bool diag_sb = sI == sJ;
  for (int cc = ty; cc < TILE; cc += 4) {
    const double v = src[cc * TILE + tx];                          // tile(u = tx, v = cc)
    sm[cc][tx] = v;
    const int64_t r = static_cast<int64_t>(i) * TILE + tx, c = static_cast<int64_t>(j) * TILE + cc;
    if (r < nr && c < nc) direct[r + ldd * c] = v;
  }
  if (diag_sb && I == J) return;                                   // (a diagonal tile is its own mirror image)
  __syncthreads();
  double* out = diag_sb ? direct : mirror;                         // diagonal super-block: lower part of the same block
  const int64_t ld = diag_sb ? ldd : ldm;
  for (int cc = ty; cc < TILE; cc += 4) {
    // element (row = 64 j + tx, column = 64 i + cc) of the transposed block = tile(u = cc, v = tx)
    const int64_t r = static_cast<int64_t>(j) * TILE + tx, c = static_cast<int64_t>(i) * TILE + cc;
    if (r < (diag_sb ? nr : nc) && c < (diag_sb ? nc : nr)) out[r + ld * c] = sm[tx][cc];
  }
}

// dst[(l*pp) + j] = (j < p && l < k) ? src[l*lds + j] : 0, l < KP   (zero-padded operand layout)
__global__ void __launch_bounds__(256) k_pad_rows(const double* __restrict__ src, int64_t lds, int64_t p, int64_t pp,
                                                  int k, int KP, double* __restrict__ dst) {
  const int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (j >= pp) return;
  for (int l = 0; l < KP; ++l)
    dst[static_cast<int64_t>(l) * pp + j] = (j < p && l < k) ? src[static_cast<int64_t>(l) * lds + j] : 0.0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// `slots` column-major FP64 matrices of rows x cols (slot stride rows * cols), 128-byte swizzle.
//   rows % 16 == 0: 4-D (r_lo 16, column, r_hi rows/16, slot), box {16, box_cols, box_rblk, 1}: {64, 4} whole 64 x 64 tiles,
//                   {32, 4} / {64, 2} the direct / mirrored image of a 64 x 32 half tile
//   rows %  2 == 0: 3-D (row, column, slot),                    box {16, 64, 1} = one sub-box per store
// Dense draws: rows = cols = p.  Compact tile buffers (tiles.cuh): rows = 64, cols = 64 x tiles.
int make_output_map(fable_ctx* ctx, double* base, int64_t rows, int64_t cols, int slots, CUtensorMap* out, int box_cols = TILE,
                    int box_rblk = TILE / BOXW) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    FB_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return ctx->fail(FABLE_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t ur = static_cast<cuuint64_t>(rows), uc = static_cast<cuuint64_t>(cols);
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r;
  if (rows % BOXW == 0) {
    const cuuint64_t dims[4] = {BOXW, uc, ur / BOXW, static_cast<cuuint64_t>(slots)};
    const cuuint64_t strides[3] = {ur * 8, BOXW * 8, ur * uc * 8};
    const cuuint32_t box[4] = {BOXW, static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rblk), 1};
    r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const cuuint64_t dims[3] = {ur, uc, static_cast<cuuint64_t>(slots)};
    const cuuint64_t strides[2] = {ur * 8, ur * uc * 8};
    const cuuint32_t box[3] = {BOXW, TILE, 1};
    r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) return ctx->fail(FABLE_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string(r));
  return FABLE_OK;
}

struct PsiMaps { CUtensorMap out, out_m, sum, sq; };

template <bool MAT, bool ACC, bool CC = false, int TV = TILE, int KC = KC_DEFAULT>
int launch_psi(fable_ctx* ctx, const PsiArgs& a, unsigned slots, int64_t t0, const PsiMaps& m) {
  constexpr size_t smem = Geo<TV, KC>::SMEM_BYTES;
  static bool attr_dev[64] = {};
  bool& attr = attr_dev[ctx->device & 63];
  if (!attr) {
    FB_CUDA(ctx, cudaFuncSetAttribute(k_psi<MAT, ACC, CC, TV, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr = true;
  }
  k_psi<MAT, ACC, CC, TV, KC><<<slots * (TILE / TV), 256, smem, ctx->stream>>>(a, t0, m.out, m.out_m, m.sum, m.sq);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}

// TV = 32 (half tiles, three CTAs per SM) is used for the materialising kernels whenever whole-tile 4-D stores
// exist (dense p % 16 == 0, or compact output): the batch is bound by how evenly the SMs keep HBM's write queues fed,
// and three staggered store streams per SM do that better than two (tools/store_lab.cu lab 2).  FABLE_B200_PSI_TV=64
// forces the whole-tile kernel (A/B measurements).
int psi_tv(const PsiArgs& a, bool mat) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("FABLE_B200_PSI_TV");
    forced = e ? atoi(e) : 0;
  }
  if (!mat || a.cc_gdiag || a.tma_mode != 1) return TILE;
  return forced == TILE ? TILE : 32;
}

}  // namespace

int fb_psi(fable_ctx* ctx, const PsiArgs& a_in) {
  PsiArgs a = a_in;
  a.nT = static_cast<int>(fbt::n_tiles(a.p));
  a.tma_mode = (a.out_compact || a.p % BOXW == 0) ? 1 : ((a.p & 1) == 0 ? 2 : 0);
  FB_REQUIRE(ctx, a.p > 0 && a.k > 0 && a.B > 0, "psi: empty problem");
  FB_REQUIRE(ctx, (a.acc_sum == nullptr) == (a.acc_sq == nullptr), "psi: need both accumulators");
  const bool mat = a.psi != nullptr, acc = a.acc_sum != nullptr;
  FB_REQUIRE(ctx, mat || acc, "psi: nothing to do");
  FB_REQUIRE(ctx, (a.cc_gdiag == nullptr) == (a.cc_sig == nullptr), "psi: entry-wise correction needs gdiag and sig");
  FB_REQUIRE(ctx, !a.cc_gdiag || mat, "psi: the entry-wise correction is defined for materialised draws");
  FB_REQUIRE(ctx, !mat || a.psi_slots > 0, "psi: psi_slots must be positive");
  FB_REQUIRE(ctx, a.pp % TILE == 0 && a.pp >= a.p && (reinterpret_cast<uintptr_t>(a.Lw) & 15) == 0,
             "psi: operand rows must be zero-padded to a multiple of 64 and 16-byte aligned");
  FB_REQUIRE(ctx, a.KP % 4 == 0 && a.KP >= a.k, "psi: operand rank must be zero-padded to a multiple of 4");
  FB_REQUIRE(ctx, !mat || a.psi_slots >= a.slot0 + a.B, "psi: one output slot per draw of the batch");
  const int64_t tiles = fbt::n_slots(a.p);    // super-block order, padding slots exit at once
  // storage slice of the tile enumeration (multi-GPU partition B: every rank processes all draws for its own tiles; no
  // collective): the compact buffers start at its first tile.  The launch may cover a part of it only (the tail of
  // fable_sampler runs chunk by chunk so that finished accumulator tiles can leave while the rest still computes).
  int64_t s_lo = 0, s_hi = tiles;
  if (a.store_end > 0) {
    FB_REQUIRE(ctx, a.store_begin >= 0 && a.store_begin <= a.store_end && a.store_end <= tiles, "psi: tile slice out of bounds");
    s_lo = a.store_begin; s_hi = a.store_end;
  }
  int64_t t_lo = s_lo, t_hi = s_hi;
  if (a.tile_end > 0) {
    FB_REQUIRE(ctx, a.tile_begin >= s_lo && a.tile_begin <= a.tile_end && a.tile_end <= s_hi, "psi: launch range outside the tile slice");
    t_lo = a.tile_begin; t_hi = a.tile_end;
  }
  a.cbase = fbt::cidx(s_lo, a.nT);
  const int64_t ntile = fbt::cidx(s_hi, a.nT) - a.cbase;          // valid tiles of the storage slice
  if (ntile == 0 || fbt::cidx(t_hi, a.nT) == fbt::cidx(t_lo, a.nT)) return FABLE_OK;
  FB_REQUIRE(ctx, ntile < (int64_t(1) << 24), "psi: more than 2^24 tiles in one slice");
  const int64_t chunk = 1 << 29;
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
  PsiMaps m;
  memset(&m, 0, sizeof(m));
  const int tv = psi_tv(a, mat);
  if (mat && a.tma_mode != 0) {
    FB_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(a.psi) & 15) == 0, "psi: the output buffer must be 16-byte aligned");
    const int64_t rows = a.out_compact ? TILE : a.p, cols = a.out_compact ? 2 * ntile * TILE : a.p;
    FB_TRY(make_output_map(ctx, a.psi, rows, cols, a.psi_slots, &m.out, tv, TILE / BOXW));
    FB_TRY(make_output_map(ctx, a.psi, rows, cols, a.psi_slots, &m.out_m, TILE, tv / BOXW));
  }
  if (acc) {
    FB_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(a.acc_sum) | reinterpret_cast<uintptr_t>(a.acc_sq)) & 15) == 0,
               "psi: the accumulators must be 16-byte aligned");
    FB_TRY(make_output_map(ctx, a.acc_sum, TILE, ntile * TILE, 1, &m.sum, tv, TILE / BOXW));
    FB_TRY(make_output_map(ctx, a.acc_sq, TILE, ntile * TILE, 1, &m.sq, tv, TILE / BOXW));
  }
  static const bool kc24_ok = !(getenv("FABLE_B200_PSI_KC24") && atoi(getenv("FABLE_B200_PSI_KC24")) == 0);
  const bool k24 = kc24_ok && a.KP > 16 && a.KP <= 24;
  for (int64_t t0 = t_lo; t0 < t_hi; t0 += chunk) {
    const unsigned g = static_cast<unsigned>(t_hi - t0 < chunk ? t_hi - t0 : chunk);
    if (a.cc_gdiag && mat && acc) FB_TRY((launch_psi<true, true, true>(ctx, a, g, t0, m)));
    else if (a.cc_gdiag && mat) FB_TRY((launch_psi<true, false, true>(ctx, a, g, t0, m)));
    // 16 < k <= 24 (BASELINE configs[3]: k = 20): one 24-row rank chunk instead of 16 + 8 with a barrier in mid-draw
    else if (mat && acc && tv == 32 && k24) FB_TRY((launch_psi<true, true, false, 32, 24>(ctx, a, g, t0, m)));
    else if (mat && tv == 32 && k24) FB_TRY((launch_psi<true, false, false, 32, 24>(ctx, a, g, t0, m)));
    else if (mat && acc && tv == 32) FB_TRY((launch_psi<true, true, false, 32>(ctx, a, g, t0, m)));
    else if (mat && tv == 32) FB_TRY((launch_psi<true, false, false, 32>(ctx, a, g, t0, m)));
    else if (mat && acc) FB_TRY((launch_psi<true, true>(ctx, a, g, t0, m)));
    else if (mat) FB_TRY((launch_psi<true, false>(ctx, a, g, t0, m)));
    else FB_TRY((launch_psi<false, true>(ctx, a, g, t0, m)));
  }
  if (ev1) FB_CUDA(ctx, cudaEventRecord(ev1, ctx->stream));
  return FABLE_OK;
}

// acc[i] += sum_s part[s][i]: a read stream of np + 1 buffers and one write stream, each element touched by one thread
__global__ void __launch_bounds__(256) k_fold_partials(double* __restrict__ acc, const double* __restrict__ part, int np, int64_t stride,
                                                       int64_t begin, int64_t end) {
  const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x * 2;
  for (int64_t i = begin + (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 2; i < end; i += step) {
    double2 a = *reinterpret_cast<const double2*>(acc + i);
    for (int s2 = 0; s2 < np; ++s2) {
      const double2 x = __ldcs(reinterpret_cast<const double2*>(part + s2 * stride + i));
      a.x += x.x; a.y += x.y;
    }
    *reinterpret_cast<double2*>(acc + i) = a;
  }
}

int fb_fold_partials(fable_ctx* ctx, double* d_acc, const double* d_part, int np, int64_t stride, int64_t begin, int64_t end) {
  if (np <= 0 || end <= begin) return FABLE_OK;
  FB_REQUIRE(ctx, begin % 2 == 0 && end % 2 == 0 && stride % 2 == 0, "fold_partials: ranges must be even");
  k_fold_partials<<<ctx->sms * 8, 256, 0, ctx->stream>>>(d_acc, d_part, np, stride, begin, end);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}

// out[e * ldm + m_off + b] = slots[b * p2 + e], b < B: a batch of materialised draws into an (MC, p, p) R array (first index
// fastest: CovSampleStor[m, , ] of R/FABLE_code.R:262, 290).  32 entries x 32 draws per block through shared memory.
__global__ void __launch_bounds__(256) k_slots_to_marray(const double* __restrict__ slots, int B, int64_t p2, int64_t ldm, int64_t m_off,
                                                         double* __restrict__ out) {
  __shared__ double t[32][33];
  const int64_t e0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int b0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int b = b0 + r;
    t[r][tx] = (b < B && e0 + tx < p2) ? slots[static_cast<int64_t>(b) * p2 + e0 + tx] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int64_t e = e0 + r;
    const int b = b0 + tx;
    if (e < p2 && b < B) out[e * ldm + m_off + b] = t[tx][r];
  }
}

int fb_slots_to_marray(fable_ctx* ctx, const double* d_slots, int B, int64_t p2, int64_t ldm, int64_t m_off, double* d_out) {
  dim3 grid(static_cast<unsigned>(ceil_div(p2, 32)), static_cast<unsigned>(ceil_div(B, 32)));
  k_slots_to_marray<<<grid, 256, 0, ctx->stream>>>(d_slots, B, p2, ldm, m_off, d_out);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}

int fb_pad_rows(fable_ctx* ctx, const double* src, int64_t lds, int64_t p, int64_t pp, int k, int KP, double* dst) {
  k_pad_rows<<<static_cast<unsigned>(ceil_div(pp, 256)), 256, 0, ctx->stream>>>(src, lds, p, pp, k, KP, dst);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}

int fb_superblock_to_dense(fable_ctx* ctx, const double* d_tiles, int64_t cbase, int64_t p, int64_t sb, double* d_direct,
                           double* d_mirror, int64_t ld_direct, int64_t ld_mirror) {
  const int nT = static_cast<int>(fbt::n_tiles(p));
  int sI, sJ;
  fbt::tri_of(sb, sI, sJ);
  const int64_t span = static_cast<int64_t>(SB) * TILE;
  const int64_t r0 = sI * span, c0 = sJ * span;
  FB_REQUIRE(ctx, c0 < p, "superblock_to_dense: super-block outside the matrix");
  const int64_t nr = p - r0 < span ? p - r0 : span, nc = p - c0 < span ? p - c0 : span;
  k_superblock_to_dense<<<dim3(SB, SB), 256, 0, ctx->stream>>>(d_tiles, nT, cbase, p, sI, sJ, nr, nc, d_direct, ld_direct > 0 ? ld_direct : nr,
                                                                d_mirror, ld_mirror > 0 ? ld_mirror : nc);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}

int fb_tiles_to_dense(fable_ctx* ctx, const double* d_tiles, int64_t p, int64_t t_lo, int64_t t_hi, int paired, int64_t col0,
                      int64_t ncols, double* d_out) {
  FB_REQUIRE(ctx, col0 % TILE == 0 && col0 >= 0 && ncols >= 1 && col0 + ncols <= p, "tiles_to_dense: bad column panel");
  const int nT = static_cast<int>(fbt::n_tiles(p));
  if (t_hi <= 0) { t_lo = 0; t_hi = fbt::n_slots(p); }
  dim3 grid(static_cast<unsigned>(nT), static_cast<unsigned>(ceil_div(ncols, TILE)));
  k_tiles_to_dense<<<grid, 256, 0, ctx->stream>>>(d_tiles, nT, t_lo, t_hi, fbt::cidx(t_lo, nT), paired, p, col0, ncols, d_out);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}
