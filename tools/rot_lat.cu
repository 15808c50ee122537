// Latency lab for the inner Jacobi step (eig.cu k_bj_eig): how many cycles does ONE warp need for the dependent chain
// "three entries from block formulas -> rotation (rsqrt, reciprocal, rsqrt)", alone on the SM and beside 31 warps that run
// apply-like FP64 work; plus the plain DFMA / rsqrt / reciprocal / BAR.SYNC latencies it is made of.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/rot_lat tools/rot_lat.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void rotation(double hpq, double hpp, double hqq, double tol2, double floor2, double2& cs_) {
  cs_ = make_double2(1.0, 0.0);
  if (hpq != 0.0 && hpq * hpq > tol2 * hpp * hqq && hpp > floor2 && hqq > floor2) {
    const double d = hqq - hpp, h2 = 2.0 * hpq, r2 = fma(d, d, h2 * h2);
    const double tt = h2 * (1.0 / (d + copysign(r2 * rsqrt(r2), d)));
    const double c_ = rsqrt(fma(tt, tt, 1.0));
    cs_ = make_double2(c_, c_ * tt);
  }
}

// mode 0: dependent DFMA chain; 1: rsqrt chain; 2: reciprocal chain; 3: rotation chain (output feeds the next input);
// 4: rotation chain + __syncthreads per iteration; 5: __syncthreads only
__global__ void __launch_bounds__(1024) k_lab(int mode, int iters, int busy, double seed, double* out, long long* cyc) {
  const int w = threadIdx.x >> 5;
  double x = seed + threadIdx.x * 1e-3, y = 0.3, z = 1.7;
  __shared__ double sh[1024];
  sh[threadIdx.x] = x;
  __syncthreads();
  long long t0 = clock64();
  if (w == 0) {
    for (int i = 0; i < iters; ++i) {
      if (mode == 0) { x = fma(x, 0.999999, 1e-9); }
      else if (mode == 1) { x = rsqrt(x) + 1.0; }
      else if (mode == 2) { x = 1.0 / x + 1.0; }
      else {
        double2 r;
        rotation(y * x * 0.1, x, z + x, 1e-30, 1e-300, r);
        x = fma(r.y, 0.25, 1.0 + r.x);                      // next input depends on (c, s)
        if (mode == 4) __syncthreads();
      }
      if (mode == 5) __syncthreads();
    }
  } else if (busy) {
    // apply-like work: 16 FP64 ops + 4 LDS + 4 STS per iteration and thread
    double a = x, b = y, c = z, d = 0.5;
    for (int i = 0; i < iters; ++i) {
      const double h0 = sh[threadIdx.x], h1 = sh[(threadIdx.x + 33) & 1023], h2 = sh[(threadIdx.x + 65) & 1023], h3 = sh[(threadIdx.x + 97) & 1023];
      const double ap = a * h0 - b * h1, bp = b * h0 + a * h1, aq = a * h2 - b * h3, bq = b * h2 + a * h3;
      const double n0 = c * ap - d * aq, n1 = d * ap + c * aq, n2 = c * bp - d * bq, n3 = d * bp + c * bq;
      if (mode >= 4) __syncthreads();
      sh[threadIdx.x] = n0 * 1e-3 + 1.0; sh[(threadIdx.x + 33) & 1023] = n1 * 1e-3 + 1.0;
      sh[(threadIdx.x + 65) & 1023] = n2 * 1e-3 + 1.0; sh[(threadIdx.x + 97) & 1023] = n3 * 1e-3 + 1.0;
      if (mode >= 4) { /* one barrier per step like the kernel */ }
    }
    x = a + sh[threadIdx.x];
  } else if (mode >= 4) {
    for (int i = 0; i < iters; ++i) __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = x;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * sizeof(double)); cudaMalloc(&cyc, sizeof(long long));
  const char* names[] = {"dependent DFMA", "rsqrt() + add", "1/x + add", "rotation chain", "rotation chain + barrier", "barrier only"};
  const int iters = 2000;
  for (int threads : {32, 1024})
    for (int busy = 0; busy < 2; ++busy) {
      if (threads == 32 && busy) continue;
      for (int mode = 0; mode < 6; ++mode) {
        for (int rep = 0; rep < 2; ++rep) k_lab<<<1, threads>>>(mode, iters, busy, 1.5, out, cyc);
        long long h = 0;
        cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("threads %4d busy %d  %-26s %8.1f cycles / iteration\n", threads, busy, names[mode], double(h) / iters);
      }
    }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
