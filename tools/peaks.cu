// Step-0 microbenchmark: the denominators the FABLE path is divided by.
//   (a) FP64 DMMA.8x8x4 register-resident issue peak   (b) FP64 DFMA peak
//   (c) HBM write-only / read-only / copy GB/s with the store flavours the covariance kernel can use
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/bin/peaks tools/peaks.cu
// Prints one JSON object on stdout. Not part of the product path.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a0, double b0) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a0, double b0) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = i * 1e-3;
  double a = a0 + threadIdx.x * 1e-12, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- HBM streams: grid-stride over 16-byte elements --------------------------------------
__global__ void __launch_bounds__(256) k_write_plain(double2* __restrict__ dst, size_t n16, double v) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  double2 x = make_double2(v, v + 1.0);
  for (; i < n16; i += stride) dst[i] = x;
}
__global__ void __launch_bounds__(256) k_write_cs(double2* __restrict__ dst, size_t n16, double v) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n16; i += stride)
    asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" :: "l"(dst + i), "d"(v), "d"(v + 1.0) : "memory");
}
__global__ void __launch_bounds__(256) k_write_na(double2* __restrict__ dst, size_t n16, double v) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n16; i += stride)
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" :: "l"(dst + i), "d"(v), "d"(v + 1.0) : "memory");
}
// 32-byte stores (sm_100 has st.global.v4.f64 / .b256? keep to v2 x2 unrolled)
__global__ void __launch_bounds__(256) k_write_x4(double2* __restrict__ dst, size_t n16, double v) {
  // each thread writes 4 consecutive-in-warp-stride 16B elements per iteration (4 stores in flight)
  size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nthreads = (size_t)gridDim.x * blockDim.x;
  double2 x = make_double2(v, v + 1.0);
  size_t i = tid;
  for (; i + 3 * nthreads < n16; i += 4 * nthreads) {
    dst[i] = x; dst[i + nthreads] = x; dst[i + 2 * nthreads] = x; dst[i + 3 * nthreads] = x;
  }
  for (; i < n16; i += nthreads) dst[i] = x;
}
__global__ void __launch_bounds__(256) k_read(const double2* __restrict__ src, size_t n16, double* out) {
  size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nthreads = (size_t)gridDim.x * blockDim.x;
  double s = 0;
  size_t i = tid;
  for (; i + 3 * nthreads < n16; i += 4 * nthreads) {
    double2 a = __ldg(src + i), b = __ldg(src + i + nthreads), c = __ldg(src + i + 2 * nthreads), d = __ldg(src + i + 3 * nthreads);
    s += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
  }
  for (; i < n16; i += nthreads) { double2 a = __ldg(src + i); s += a.x + a.y; }
  if (s == 123.456) out[0] = s;
}
__global__ void __launch_bounds__(256) k_copy(const double2* __restrict__ src, double2* __restrict__ dst, size_t n16) {
  size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nthreads = (size_t)gridDim.x * blockDim.x;
  size_t i = tid;
  for (; i + 3 * nthreads < n16; i += 4 * nthreads) {
    double2 a = __ldg(src + i), b = __ldg(src + i + nthreads), c = __ldg(src + i + 2 * nthreads), d = __ldg(src + i + 3 * nthreads);
    dst[i] = a; dst[i + nthreads] = b; dst[i + 2 * nthreads] = c; dst[i + 3 * nthreads] = d;
  }
  for (; i < n16; i += nthreads) dst[i] = __ldg(src + i);
}

// TMA-engine store: fill a smem tile once, then stream it out with cp.async.bulk shared->global.
// CHUNK bytes per bulk copy; each CTA keeps up to 8 bulk groups in flight.
template <int CHUNK>
__global__ void __launch_bounds__(128) k_write_bulk(char* __restrict__ dst, size_t nbytes) {
  extern __shared__ __align__(128) char smem[];
  for (int i = threadIdx.x; i < CHUNK / 8; i += blockDim.x) reinterpret_cast<double*>(smem)[i] = 1.0 + i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    size_t nchunks = nbytes / CHUNK;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   :: "l"(dst + c * CHUNK), "r"(s), "r"(CHUNK) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

template <typename F>
static double time_ms(F&& launch, int warm, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < warm; ++i) launch();
  CK(cudaDeviceSynchronize());
  std::vector<float> t;
  for (int i = 0; i < reps; ++i) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); t.push_back(ms);
  }
  CK(cudaGetLastError());
  std::sort(t.begin(), t.end());
  return t[t.size() / 2];
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\"", prop.name, sms, prop.major, prop.minor);

  double* scratch; CK(cudaMalloc(&scratch, (size_t)sms * 16 * 256 * sizeof(double)));
  // ---- FP64 pipes ----
  {
    const int iters = 4096;
    for (int occ : {1, 2, 4, 8}) {
      int grid = sms * occ;
      double ms = time_ms([&] { k_dmma<8><<<grid, 256>>>(scratch, iters, 1.0000001, 0.9999999); }, 2, 5);
      double flops = (double)grid * 8 /*warps*/ * iters * 8 /*NACC*/ * 512.0;
      printf(", \"dmma884_tflops_occ%d\": %.3f", occ, flops / ms * 1e-9);
    }
    for (int occ : {1, 2, 4, 8}) {
      int grid = sms * occ;
      double ms = time_ms([&] { k_dfma<16><<<grid, 256>>>(scratch, iters, 1.0000001, 1e-9); }, 2, 5);
      double flops = (double)grid * 256 * iters * 16 * 2.0;
      printf(", \"dfma_tflops_occ%d\": %.3f", occ, flops / ms * 1e-9);
    }
  }
  // ---- HBM ----
  {
    const size_t nbytes = (size_t)8 << 30;  // 8 GiB >> 126 MB L2
    char *a, *b; CK(cudaMalloc(&a, nbytes)); CK(cudaMalloc(&b, nbytes));
    CK(cudaMemset(a, 0, nbytes)); CK(cudaMemset(b, 0, nbytes));
    size_t n16 = nbytes / 16;
    double gb = nbytes * 1e-9;
    for (int mult : {4, 8, 16, 32}) {
      int grid = sms * mult;
      double ms;
      ms = time_ms([&] { k_write_plain<<<grid, 256>>>((double2*)a, n16, 1.0); }, 2, 7);
      printf(", \"write_plain_gbs_x%d\": %.1f", mult, gb / ms * 1e3);
      ms = time_ms([&] { k_write_cs<<<grid, 256>>>((double2*)a, n16, 1.0); }, 2, 7);
      printf(", \"write_cs_gbs_x%d\": %.1f", mult, gb / ms * 1e3);
      ms = time_ms([&] { k_write_na<<<grid, 256>>>((double2*)a, n16, 1.0); }, 2, 7);
      printf(", \"write_na_gbs_x%d\": %.1f", mult, gb / ms * 1e3);
      ms = time_ms([&] { k_write_x4<<<grid, 256>>>((double2*)a, n16, 1.0); }, 2, 7);
      printf(", \"write_x4_gbs_x%d\": %.1f", mult, gb / ms * 1e3);
      ms = time_ms([&] { k_read<<<grid, 256>>>((const double2*)a, n16, scratch); }, 2, 7);
      printf(", \"read_gbs_x%d\": %.1f", mult, gb / ms * 1e3);
      ms = time_ms([&] { k_copy<<<grid, 256>>>((const double2*)a, (double2*)b, n16); }, 2, 7);
      printf(", \"copy_rw_gbs_x%d\": %.1f", mult, 2 * gb / ms * 1e3);
    }
    {
      double ms = time_ms([&] { CK(cudaMemsetAsync(a, 1, nbytes)); }, 2, 7);
      printf(", \"memset_gbs\": %.1f", gb / ms * 1e3);
      ms = time_ms([&] { CK(cudaMemcpyAsync(b, a, nbytes, cudaMemcpyDeviceToDevice)); }, 2, 7);
      printf(", \"memcpy_d2d_rw_gbs\": %.1f", 2 * gb / ms * 1e3);
    }
    // bulk (TMA-engine) stores
    {
      CK(cudaFuncSetAttribute(k_write_bulk<16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
      CK(cudaFuncSetAttribute(k_write_bulk<65536>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
      for (int mult : {1, 2, 4}) {
        int grid = sms * mult;
        double ms = time_ms([&] { k_write_bulk<4096><<<grid, 128, 4096>>>(a, nbytes); }, 2, 7);
        printf(", \"write_bulk4k_gbs_x%d\": %.1f", mult, gb / ms * 1e3);
        ms = time_ms([&] { k_write_bulk<16384><<<grid, 128, 16384>>>(a, nbytes); }, 2, 7);
        printf(", \"write_bulk16k_gbs_x%d\": %.1f", mult, gb / ms * 1e3);
        ms = time_ms([&] { k_write_bulk<65536><<<grid, 128, 65536>>>(a, nbytes); }, 2, 7);
        printf(", \"write_bulk64k_gbs_x%d\": %.1f", mult, gb / ms * 1e3);
      }
      // 1 KB copies = one 128-row tile column of the covariance output
      for (int mult : {2, 4, 8}) {
        int grid = sms * mult;
        double ms = time_ms([&] { k_write_bulk<1024><<<grid, 128, 1024>>>(a, nbytes); }, 2, 7);
        printf(", \"write_bulk1k_gbs_x%d\": %.1f", mult, gb / ms * 1e3);
      }
    }
    CK(cudaFree(a)); CK(cudaFree(b));
  }
  printf("}\n");
  return 0;
}
