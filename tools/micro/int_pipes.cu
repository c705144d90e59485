// Micro-benchmark: issue rate of the integer instructions the decode + lookup kernel is made of,
// alone and mixed, to see which of them share the ALU pipe and which run on the FMA pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipes int_pipes.cu && ./int_pipes
// Prints thread-instructions per clock per SM (128 = one warp instruction per scheduler per clock).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CH 8
template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t m, uint32_t s, int iters) {
  uint32_t a[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) a[i] = threadIdx.x * 17 + i + m;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      uint32_t x = a[i];
      if (OP == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(x) : "r"(m), "r"(s));
      if (OP == 1) asm volatile("shr.u32 %0, %0, 3;" : "+r"(x));
      if (OP == 2) asm volatile("shf.r.wrap.b32 %0, %0, %0, %1;" : "+r"(x) : "r"(s));
      if (OP == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(s));
      if (OP == 4) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x) : "r"(m));
      if (OP == 5) asm volatile("shl.b32 %0, %0, 3;" : "+r"(x));
      if (OP == 6) asm volatile("prmt.b32 %0, %0, %1, 0x3715;" : "+r"(x) : "r"(m));
      if (OP == 7) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(x) : "r"(m), "r"(s));
                     asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(s)); }
      if (OP == 8) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(x) : "r"(m), "r"(s));
                     asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x) : "r"(m)); }
      if (OP == 9) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(x) : "r"(m), "r"(s));
                     asm volatile("shf.r.wrap.b32 %0, %0, %0, %1;" : "+r"(x) : "r"(s)); }
      if (OP == 10) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(s));
                      asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x) : "r"(m)); }
      if (OP == 11) { uint32_t lo, hi; asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(*(unsigned long long*)&a[i & ~1]) : "r"(x), "r"(m)); x = a[i]; }
      if (OP == 12) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(x) : "r"(m), "r"(s));
                      asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(x) : "r"(m), "r"(s));
                      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(s)); }
      a[i] = x;
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) r ^= a[i];
  if (r == 0x12345) *out = r;
}

template <int OP>
void run(const char* name, int per_iter, uint32_t* out) {
  const int iters = 4096, grid = 148 * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP><<<grid, 256>>>(out, 0x9e3779b1u, 5, iters);
  cudaEventRecord(e0);
  k<OP><<<grid, 256>>>(out, 0x9e3779b1u, 5, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double ops = (double)grid * 256 * iters * CH * per_iter;
  printf("%-28s %8.3f ms  %7.1f thread-inst/clk/SM (at %d MHz nominal)\n", name, ms, ops / (ms * 1e-3) / (clk * 1e3) / 148, clk / 1000);
}

int main() {
  uint32_t* out; cudaMalloc(&out, 4);
  run<0>("LOP3", 1, out);
  run<1>("SHF.R const", 1, out);
  run<2>("SHF.R.W variable (rotate)", 1, out);
  run<3>("IMAD", 1, out);
  run<4>("IMAD.HI.U32", 1, out);
  run<5>("SHL const", 1, out);
  run<6>("PRMT", 1, out);
  run<7>("LOP3 + IMAD", 2, out);
  run<8>("LOP3 + IMAD.HI", 2, out);
  run<9>("LOP3 + SHF", 2, out);
  run<10>("IMAD + IMAD.HI", 2, out);
  run<11>("IMAD.WIDE", 1, out);
  run<12>("2 LOP3 + IMAD", 3, out);
  return 0;
}
