// Micro-benchmark: cost of random 8-byte gathers from a large array as a function of the L2 fetch
// granularity (cudaLimitMaxL2FetchGranularity) and of the load's PTX cache hints.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_gran gather_gran.cu && ./gather_gran
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__device__ __forceinline__ long long ld8(const long long* p) {
  long long v;
  if (MODE == 0) v = __ldg(p);
  else if (MODE == 1) asm volatile("ld.global.nc.L2::64B.b64 %0, [%1];" : "=l"(v) : "l"(p));
  else if (MODE == 2) asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(v) : "l"(p));
  else if (MODE == 3) asm volatile("ld.global.cv.b64 %0, [%1];" : "=l"(v) : "l"(p));
  else asm volatile("ld.global.nc.L1::no_allocate.L2::64B.b64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

template <int MODE>
__global__ void gather(const long long* __restrict__ a, const uint32_t* __restrict__ idx, int64_t n, long long* out) {
  long long s = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s += ld8<MODE>(a + idx[i]);
  if (s == 0x1234567) *out = s;
}

__global__ void make_idx(uint32_t* idx, int64_t n, uint32_t stride_rows) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u;
    idx[i] = (uint32_t)(i * stride_rows + (h >> 27));  // ascending, ~1 survivor per stride_rows rows, jittered
  }
}

template <int MODE>
float run(const long long* a, const uint32_t* idx, int64_t n, long long* out) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  gather<MODE><<<148 * 8, 256>>>(a, idx, n, out);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) gather<MODE><<<148 * 8, 256>>>(a, idx, n, out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main() {
  const int64_t rows = 600000000, n = rows / 55;  // ~1.8 % survivors
  long long* a;
  uint32_t* idx;
  long long* out;
  cudaMalloc(&a, rows * 8);
  cudaMemset(a, 1, rows * 8);
  cudaMalloc(&idx, n * 4);
  cudaMalloc(&out, 8);
  make_idx<<<1024, 256>>>(idx, n, 55);
  cudaDeviceSynchronize();
  size_t lim = 0;
  cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity);
  printf("default cudaLimitMaxL2FetchGranularity = %zu\n", lim);
  for (size_t g : {(size_t)0, (size_t)32, (size_t)64, (size_t)128}) {
    if (g) {
      cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g);
      cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity);
      printf("set %zu -> %s, now %zu\n", g, cudaGetErrorString(e), lim);
    }
    printf("  ldg            %.4f ms\n", run<0>(a, idx, n, out));
    printf("  nc.L2::64B     %.4f ms\n", run<1>(a, idx, n, out));
    printf("  nc.no_allocate %.4f ms\n", run<2>(a, idx, n, out));
    printf("  cv             %.4f ms\n", run<3>(a, idx, n, out));
    printf("  nc.noalloc.64B %.4f ms\n", run<4>(a, idx, n, out));
  }
  printf("survivors %lld (%.1f MB at 32 B, %.1f MB at 64 B, %.1f MB at 128 B per survivor) + index %.1f MB\n", (long long)n,
         n * 32e-6, n * 64e-6, n * 128e-6, n * 4e-6);
  return 0;
}
