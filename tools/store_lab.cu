// store_lab: HBM write bandwidth of the access patterns the covariance kernel can choose from.
// Pure stores (no arithmetic): a dense symmetric p x p FP64 matrix per "draw", written as square
// T x T tiles of the upper triangle plus their mirrored tiles, B draws per CTA into a ring of slots.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/store_lab tools/store_lab.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void tri_of(int64_t t, int& I, int& J) {
  int64_t j = (int64_t)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while (j * (j + 1) / 2 > t) --j;
  while ((j + 1) * (j + 2) / 2 <= t) ++j;
  J = (int)j; I = (int)(t - j * (j + 1) / 2);
}

// order 0: column-major upper triangle.  order S>0: super-blocks of S x S tiles, the super-blocks
// enumerated over their own upper triangle, tiles inside row-major; tiles below the diagonal skipped
// (CTA exits).
__device__ __forceinline__ bool tile_of(int64_t t, int nT, int S, int& I, int& J) {
  if (S == 0) { tri_of(t, I, J); return true; }
  const int64_t per = (int64_t)S * S;
  int sI, sJ;
  tri_of(t / per, sI, sJ);
  const int r = (int)(t % per);
  I = sI * S + r % S; J = sJ * S + r / S;
  return I <= J && J < nT;
}

template <int T, int KIND>
__global__ void __launch_bounds__(256) k_tiles(double* __restrict__ out, int64_t p, int nT, int S, int B, int slots,
                                               int mirror) {
  int I, J;
  if (!tile_of(blockIdx.x, nT, S, I, J)) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t u0 = (int64_t)I * T, v0 = (int64_t)J * T;
  for (int b = 0; b < B; ++b) {
    double* o = out + (int64_t)(b % slots) * p * p;
    const double val = (double)(b + I + J);
    // direct tile: columns v0..v0+T-1, rows u0.. (runs of T doubles)
    for (int c = warp; c < T; c += 8) {
      const int64_t v = v0 + c;
      if (v >= p) break;
      for (int r = lane; r < T; r += 32) {
        const int64_t u = u0 + r;
        if (u < p) {
          if (KIND == 0) o[v * p + u] = val;
          else __stcs(o + v * p + u, val);
        }
      }
    }
    if (mirror && I != J) {
      for (int c = warp; c < T; c += 8) {
        const int64_t u = u0 + c;
        if (u >= p) break;
        for (int r = lane; r < T; r += 32) {
          const int64_t v = v0 + r;
          if (v < p) {
            if (KIND == 0) o[u * p + v] = val;
            else __stcs(o + u * p + v, val);
          }
        }
      }
    }
  }
}

// bulk (TMA) stores from shared memory: T x T tile, direct + mirrored runs of T doubles each, issued
// either by lane 0 of each of the 8 warps (2T/8 runs per warp) or spread over 2T threads.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int T, int SPREAD>
__global__ void __launch_bounds__(256) k_bulk(double* __restrict__ out, int64_t p, int nT, int B) {
  extern __shared__ __align__(128) double sm[];   // [2][T][T+2]
  int I, J;
  tri_of(blockIdx.x, I, J);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t u0 = (int64_t)I * T, v0 = (int64_t)J * T;
  for (int i = threadIdx.x; i < 2 * T * (T + 2); i += 256) sm[i] = (double)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const uint32_t ru = (uint32_t)((p - u0 < T ? p - u0 : T) * 8), rv = (uint32_t)((p - v0 < T ? p - v0 : T) * 8);
  for (int b = 0; b < B; ++b) {
    double* o = out + (int64_t)b * p * p;
    if (SPREAD) {
      if (threadIdx.x < 2 * T && (T == 64 || threadIdx.x < 256)) {
        for (int r = threadIdx.x; r < 2 * T; r += 256) {
          const int c = r % T; const bool mir = r >= T;
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          if (!mir) { if (v0 + c < p) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(o + (v0 + c) * p + u0), "r"(smem_u32(sm + c * (T + 2))), "r"(ru) : "memory"); }
          else if (I != J && u0 + c < p) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(o + (u0 + c) * p + v0), "r"(smem_u32(sm + (T + c) * (T + 2))), "r"(rv) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    } else if (lane == 0) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      for (int r = warp * (2 * T / 8); r < (warp + 1) * (2 * T / 8); ++r) {
        const int c = r % T; const bool mir = r >= T;
        if (!mir) { if (v0 + c < p) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(o + (v0 + c) * p + u0), "r"(smem_u32(sm + c * (T + 2))), "r"(ru) : "memory"); }
        else if (I != J && u0 + c < p) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(o + (u0 + c) * p + v0), "r"(smem_u32(sm + (T + c) * (T + 2))), "r"(rv) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <int T, int SPREAD>
float run_bulk(double* d, int64_t p, int B) {
  const int nT = (int)((p + T - 1) / T);
  const int64_t grid = (int64_t)nT * (nT + 1) / 2;
  const size_t smem = sizeof(double) * 2 * T * (T + 2);
  CK(cudaFuncSetAttribute(k_bulk<T, SPREAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_bulk<T, SPREAD><<<(unsigned)grid, 256, smem>>>(d, p, nT, 2);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k_bulk<T, SPREAD><<<(unsigned)grid, 256, smem>>>(d, p, nT, B);
  CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms;
}

// contiguous reference: grid-stride, one draw = p*p doubles
__global__ void __launch_bounds__(256) k_contig(double* __restrict__ out, int64_t n, int B, int slots) {
  for (int b = 0; b < B; ++b) {
    double* o = out + (int64_t)(b % slots) * n;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) __stcs(o + i, (double)b);
  }
}

template <int T, int KIND>
float run_tiles(double* d, int64_t p, int S, int B, int slots, int mirror) {
  const int nT = (int)((p + T - 1) / T);
  int64_t grid;
  if (S == 0) grid = (int64_t)nT * (nT + 1) / 2;
  else { const int64_t nS = (nT + S - 1) / S; grid = nS * (nS + 1) / 2 * S * S; }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_tiles<T, KIND><<<(unsigned)grid, 256>>>(d, p, nT, S, 2, slots, mirror);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k_tiles<T, KIND><<<(unsigned)grid, 256>>>(d, p, nT, S, B, slots, mirror);
  CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms;
}

int lab2_main(int64_t p, int B);
int main(int argc, char** argv) {
  const int64_t p = argc > 1 ? atoll(argv[1]) : 20000;
  const int B = argc > 2 ? atoi(argv[2]) : 8;
  if (argc > 3) return lab2_main(p, B);
  const int slots = B;   // one slot per draw: a shorter ring is absorbed by the 126 MB L2 (tile-major order)
  double* d; CK(cudaMalloc(&d, sizeof(double) * p * p * slots));
  CK(cudaMemset(d, 0, sizeof(double) * p * p * slots));
  const double gb = 8.0 * p * p * B / 1e9;
  {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_contig<<<148 * 8, 256>>>(d, p * p, 1, slots); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); k_contig<<<148 * 8, 256>>>(d, p * p, B, slots); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("contig                                  %8.1f GB/s\n", gb / (ms * 1e-3));
  }
  printf("bulk T=64  (512 B runs) lane0/warp       %8.1f GB/s\n", gb / (run_bulk<64, 0>(d, p, B) * 1e-3));
  printf("bulk T=64  (512 B runs) spread            %8.1f GB/s\n", gb / (run_bulk<64, 1>(d, p, B) * 1e-3));
  printf("bulk T=128 (1 KB runs)  lane0/warp        %8.1f GB/s\n", gb / (run_bulk<128, 0>(d, p, B) * 1e-3));
  const int orders[] = {0, 16};
  for (int mirror = 0; mirror < 2; ++mirror)
    for (int S : orders) {
      // full-matrix bytes when mirror=1; upper triangle only (about half) when mirror=0
      const double g = mirror ? gb : gb * 0.5;
      printf("T=64  kind=cs  order=%2d mirror=%d        %8.1f GB/s\n", S, mirror, g / (run_tiles<64, 1>(d, p, S, B, slots, mirror) * 1e-3));
      printf("T=128 kind=cs  order=%2d mirror=%d        %8.1f GB/s\n", S, mirror, g / (run_tiles<128, 1>(d, p, S, B, slots, mirror) * 1e-3));
      printf("T=256 kind=cs  order=%2d mirror=%d        %8.1f GB/s\n", S, mirror, g / (run_tiles<256, 1>(d, p, S, B, slots, mirror) * 1e-3));
      printf("T=64  kind=st  order=%2d mirror=%d        %8.1f GB/s\n", S, mirror, g / (run_tiles<64, 0>(d, p, S, B, slots, mirror) * 1e-3));
    }
  return 0;
}

// ---- lab 2: stage + TMA tensor store structure of the covariance kernel, without the math -------------
// Per draw and CTA: wait until the previous stores have read the staging, rewrite the staging (both
// orientations), fence, barrier, one thread issues the two tensor stores.  TU x TV tile, NT threads,
// `spin` clocks of dummy work per draw standing in for the math.  Answers: what does this structure reach
// with 2 agents per SM (64 x 64, 256 threads) vs 4 agents per SM (64 x 32, 128 threads)?
#include <cuda.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CUtensorMap make_map(double* base, int64_t p, int slots, int box_cols, int box_rblk) {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  CUtensorMap m;
  const cuuint64_t up = (cuuint64_t)p;
  const cuuint64_t dims[4] = {16, up, up / 16, (cuuint64_t)slots};
  const cuuint64_t strides[3] = {up * 8, 16 * 8, up * up * 8};
  const cuuint32_t box[4] = {16, (cuuint32_t)box_cols, (cuuint32_t)box_rblk, 1}, es[4] = {1, 1, 1, 1};
  CUresult r = ((EncodeTiledFn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}
template <int TU, int TV, int NT>
__global__ void __launch_bounds__(NT) k_stage(const __grid_constant__ CUtensorMap mapD, const __grid_constant__ CUtensorMap mapM,
                                              int nTu, int B, int spin, double* sink) {
  extern __shared__ unsigned char raw[];
  unsigned char* al = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  double* D = reinterpret_cast<double*>(al);
  double* M = D + TU * TV;
  // tile (I over TU-row blocks, J over TV-column blocks), upper part only: I*TU <= J*TV + TV - 1
  const int J = blockIdx.y, I = blockIdx.x;
  if ((int64_t)I * TU > (int64_t)J * TV + TV - 1) return;
  const bool diag = ((int64_t)I * TU + TU > (int64_t)J * TV);
  const int tid = threadIdx.x;
  double acc = tid;
  for (int b = 0; b < B; ++b) {
    if (spin > 0) { const long long t0 = clock64(); while (clock64() - t0 < spin) acc = acc * 1.0000001 + 1e-9; }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();
    for (int i = tid; i < TU * TV; i += NT) { D[i] = acc + i; M[i] = acc - i; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"((uint64_t)&mapD), "r"(0),
                   "r"(J * TV), "r"(I * TU / 16), "r"(b), "r"(smem_u32(D)) : "memory");
      if (!diag)
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"((uint64_t)&mapM), "r"(0),
                     "r"(I * TU), "r"(J * TV / 16), "r"(b), "r"(smem_u32(M)) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  if (acc == 12345.678) sink[0] = acc;
}
template <int TU, int TV, int NT>
void run_stage(double* d, int64_t p, int B, int spin, size_t smem_force, const char* label) {
  CUtensorMap mD = make_map(d, p, B, TV, TU / 16), mM = make_map(d, p, B, TU, TV / 16);
  size_t smem = 2 * TU * TV * 8 + 1024;
  if (smem_force > smem) smem = smem_force;
  CK(cudaFuncSetAttribute(k_stage<TU, TV, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((p + TU - 1) / TU), (unsigned)((p + TV - 1) / TV));
  double* sink; CK(cudaMalloc(&sink, 8));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_stage<TU, TV, NT><<<grid, NT, smem>>>(mD, mM, 0, 2, spin, sink); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k_stage<TU, TV, NT><<<grid, NT, smem>>>(mD, mM, 0, B, spin, sink);
  CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("%-44s spin=%5d  %8.1f GB/s\n", label, spin, 8.0 * p * p * B / 1e9 / (ms * 1e-3));
}
int lab2_main(int64_t p, int B) {
  double* d; CK(cudaMalloc(&d, sizeof(double) * p * p * B));
  for (int spin : {0, 2500, 3500, 4000, 4500, 5000}) {
    run_stage<64, 64, 256>(d, p, B, spin, 100 * 1024, "stage+TMA 64x64, 256 thr, 2 CTAs/SM");
    run_stage<64, 64, 256>(d, p, B, spin, 70 * 1024, "stage+TMA 64x64, 256 thr, 3 CTAs/SM");
    run_stage<64, 32, 128>(d, p, B, spin, 50 * 1024, "stage+TMA 64x32, 128 thr, 4 CTAs/SM");
    run_stage<64, 32, 128>(d, p, B, spin, 35 * 1024, "stage+TMA 64x32, 128 thr, 6 CTAs/SM");
  }
  return 0;
}
