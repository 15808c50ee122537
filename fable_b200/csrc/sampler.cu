// Monte-Carlo draw kernels: the loop body of CPPFABLESampler
// (reference src/updated-FABLE-functions.cpp:599-618), one thread per (draw m, row j).
//   sigma2_mj = (gnd_j / 2) / Gamma_mj            (:604)
//   Lambda_m[j,l] = G0[j,l] + a_n * sqrt(sigma2_mj) * Z_m[j,l]   (:613-615)
// Two output layouts from the same variates:
//   k_sample_work : working layout for the covariance kernel, Lw[(b*KP + l)*pp + j] (row j
//                   contiguous per (draw, l): coalesced tile loads), s2w[b*pp + j]
//   k_sample_ref  : the reference's return layout, LambdaSamples[m + ldm*(j*k+l)] (:616),
//                   SigmaSqSamples[m + ldm*j] (:606); lanes run along m so stores coalesce.
#include "common.cuh"
#include "philox.cuh"

namespace {

struct DrawSrc {
  uint64_t seed;
  const double* gam;
  const double* z;
  int64_t m_inj0, p;
  int k;
  double shape;
};

__device__ __forceinline__ double draw_gamma(const DrawSrc& s, int64_t m, int64_t j) {
  if (s.gam) return s.gam[(m - s.m_inj0) * s.p + j];
  return fbrng::gamma_mt(s.seed, static_cast<uint64_t>(m), static_cast<uint32_t>(j), s.shape);
}

// normals for (m, j), pair index h -> l = 2h, 2h+1
__device__ __forceinline__ void draw_normal_pair(const DrawSrc& s, int64_t m, int64_t j, int h,
                                                 double& z0, double& z1) {
  if (s.z) {
    const double* base = s.z + (m - s.m_inj0) * s.p * s.k + j;
    z0 = base[static_cast<int64_t>(2 * h) * s.p];
    z1 = (2 * h + 1 < s.k) ? base[static_cast<int64_t>(2 * h + 1) * s.p] : 0.0;
  } else {
    fbrng::normal_pair(s.seed, static_cast<uint64_t>(m), static_cast<uint32_t>(j),
                       static_cast<uint32_t>(h), z0, z1);
  }
}

__global__ void __launch_bounds__(256)
k_sample_work(DrawSrc src, const double* __restrict__ G0, int64_t ldg, const double* __restrict__ gnd,
              double a_n, int64_t m0, int B, int64_t pp, int KP, double* __restrict__ Lw,
              double* __restrict__ s2w) {
  const int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int b = blockIdx.y;
  if (j >= pp || b >= B) return;
  const int64_t m = m0 + b;
  const int k = src.k;
  double* Lrow = Lw + (static_cast<int64_t>(b) * KP) * pp + j;
  if (j >= src.p) {  // zero padding rows so the covariance tiles need no bounds checks on loads
    s2w[static_cast<int64_t>(b) * pp + j] = 0.0;
    for (int l = 0; l < KP; ++l) Lrow[static_cast<int64_t>(l) * pp] = 0.0;
    return;
  }
  const double s2 = (0.5 * gnd[j]) / draw_gamma(src, m, j);
  s2w[static_cast<int64_t>(b) * pp + j] = s2;
  const double sd = sqrt(s2);
  for (int h = 0; 2 * h < k; ++h) {
    double z0, z1;
    draw_normal_pair(src, m, j, h, z0, z1);
    const int l = 2 * h;
    Lrow[static_cast<int64_t>(l) * pp] = G0[j + ldg * l] + a_n * (z0 * sd);
    if (l + 1 < k) Lrow[static_cast<int64_t>(l + 1) * pp] = G0[j + ldg * (l + 1)] + a_n * (z1 * sd);
  }
  for (int l = k; l < KP; ++l) Lrow[static_cast<int64_t>(l) * pp] = 0.0;
}

// lanes along m: thread (m, j); block = 128 draws x 1.. rows handled by blockIdx.y
__global__ void __launch_bounds__(128)
k_sample_ref(DrawSrc src, const double* __restrict__ G0, int64_t ldg, const double* __restrict__ gnd,
             double a_n, int64_t m0, int64_t MC, int64_t j0, int64_t nj, int64_t ldm,
             double* __restrict__ Lam, double* __restrict__ Sig) {
  const int64_t mi = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (mi >= MC) return;
  const int64_t m = m0 + mi;
  const int k = src.k;
  for (int64_t jj = blockIdx.y; jj < nj; jj += gridDim.y) {
    const int64_t j = j0 + jj;
    const double s2 = (0.5 * gnd[j]) / draw_gamma(src, m, j);
    if (Sig) Sig[mi + ldm * jj] = s2;
    if (Lam) {
      const double sd = sqrt(s2);
      for (int h = 0; 2 * h < k; ++h) {
        double z0, z1;
        draw_normal_pair(src, m, j, h, z0, z1);
        const int l = 2 * h;
        Lam[mi + ldm * (jj * k + l)] = G0[j + ldg * l] + a_n * (z0 * sd);
        if (l + 1 < k) Lam[mi + ldm * (jj * k + l + 1)] = G0[j + ldg * (l + 1)] + a_n * (z1 * sd);
      }
    }
  }
}

DrawSrc make_src(const SampleArgs& a) {
  DrawSrc s;
  s.seed = a.seed; s.gam = a.gam; s.z = a.z; s.m_inj0 = a.m_inj0; s.p = a.p; s.k = a.k; s.shape = a.shape;
  return s;
}

}  // namespace

int fb_sample_work(fable_ctx* ctx, const SampleArgs& a, int64_t m0, int B, int64_t pp, int KP,
                   double* Lw, double* s2w) {
  FB_REQUIRE(ctx, (a.gam == nullptr) == (a.z == nullptr), "injected draws need both gam and z");
  dim3 grid(static_cast<unsigned>(ceil_div(pp, 256)), static_cast<unsigned>(B));
  k_sample_work<<<grid, 256, 0, ctx->stream>>>(make_src(a), a.G0, a.ldg, a.gnd, a.a_n, m0, B, pp, KP, Lw, s2w);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}

int fb_sample_ref(fable_ctx* ctx, const SampleArgs& a, int64_t m0, int64_t MC, int64_t j0, int64_t nj,
                  int64_t ldm, double* dLambda, double* dSigma) {
  FB_REQUIRE(ctx, (a.gam == nullptr) == (a.z == nullptr), "injected draws need both gam and z");
  if (MC <= 0 || nj <= 0) return FABLE_OK;
  dim3 grid(static_cast<unsigned>(ceil_div(MC, 128)), static_cast<unsigned>(nj < 32768 ? nj : 32768));
  k_sample_ref<<<grid, 128, 0, ctx->stream>>>(make_src(a), a.G0, a.ldg, a.gnd, a.a_n, m0, MC, j0, nj, ldm,
                                              dLambda, dSigma);
  FB_LAUNCHED(ctx);
  return FABLE_OK;
}
