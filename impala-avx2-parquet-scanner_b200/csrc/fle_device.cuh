// Device-side building blocks shared by the FLE kernels (sm_100a).
//   * mbarrier + TMA 1-D bulk copy wrappers (cp.async.bulk, SASS UBLKCP)
//   * shared-memory plane loaders
//   * the 32x32 bit-matrix transpose that turns bit planes into values and back
#pragma once
#ifdef __CUDACC_RTC__
#ifndef FLE_RTC_TYPES
#define FLE_RTC_TYPES
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef int int32_t;
typedef unsigned long long uint64_t;
typedef long long int64_t;
#endif
#else
#include <cstdint>
#endif

#ifdef __CUDACC__
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#endif
#define FLE_HD __host__ __device__ __forceinline__
#else
#define FLE_HD inline
#endif

namespace fle {

// ------------------------------------------------------------------------------------------
// 32 x 32 bit-matrix transpose, fully unrolled so that it lives in registers and so that
// rows known to be zero at compile time (planes >= bit width) fold away.
//
// Convention (checked on the host in tests/test_host_math.py through fle_selftest):
// after transpose32(A):   A'[k] bit i  ==  A[31 - i] bit (31 - k).
// Unpack: load plane i's 32-bit half into A[31 - i]; A'[k] is then the value of row k of
// that half (row k lives at bit 31-k of the half-word, fle-encoding.h:426-430).
// Pack:   load the value of row k into A[k]; A'[31 - i] is plane i's half-word
// (fle-encoding.h:8338-8340).  The network is its own inverse.
// ------------------------------------------------------------------------------------------
template <int J, uint32_t M>
FLE_HD void transpose32_stage(uint32_t (&A)[32]) {
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    if ((k & J) == 0) {
#ifdef __CUDA_ARCH__
      // the 16- and 8-bit stages move whole bytes: one PRMT per output word instead of
      // shift/xor/and/xor/shift/xor
      if (J == 16) {
        const uint32_t x = A[k], y = A[k + J];
        A[k] = __byte_perm(x, y, 0x3276);      // (x & 0xFFFF0000) | (y >> 16)
        A[k + J] = __byte_perm(x, y, 0x1054);  // (y & 0x0000FFFF) | (x << 16)
        continue;
      }
      if (J == 8) {
        const uint32_t x = A[k], y = A[k + J];
        A[k] = __byte_perm(x, y, 0x3715);      // (x & 0xFF00FF00) | ((y >> 8) & 0x00FF00FF)
        A[k + J] = __byte_perm(x, y, 0x2604);  // (y & 0x00FF00FF) | ((x & 0x00FF00FF) << 8)
        continue;
      }
      if (true) {
        // (a & ~M) | (b & M) as ONE LOP3 per output word: written in C the compiler spends a
        // LOP3 per constant (M and ~M), and the xor form below needs three
        const uint32_t x = A[k], y = A[k + J];
        asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(A[k]) : "r"(x), "r"(y >> J), "n"(M));
        asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(A[k + J]) : "r"(x << J), "r"(y), "n"(M));
        continue;
      }
#endif
      uint32_t t = (A[k] ^ (A[k + J] >> J)) & M;
      A[k] ^= t;
      A[k + J] ^= (t << J);
    }
  }
}

FLE_HD void transpose32(uint32_t (&A)[32]) {
  transpose32_stage<16, 0x0000FFFFu>(A);
  transpose32_stage<8, 0x00FF00FFu>(A);
  transpose32_stage<4, 0x0F0F0F0Fu>(A);
  transpose32_stage<2, 0x33333333u>(A);
  transpose32_stage<1, 0x55555555u>(A);
}

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------
// mbarrier / TMA bulk copy (PTX ISA: mbarrier, cp.async.bulk)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init() {
  // make the inits visible to the async proxy (TMA) before the first bulk copy
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

// Bounded wait: a protocol bug must trap (and fail the launch) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}

// global -> shared bulk copy, completion signalled on an mbarrier in this CTA.
// dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(uint32_t dst_smem, const void* src, uint32_t bytes,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// shared -> global bulk store (bulk async-group completion).
__device__ __forceinline__ void tma_store_1d(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
               "r"(src_smem), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Ampere-style asynchronous 16-byte copies global -> shared (LDGSTS): arbitrary (swizzled)
// destinations, no staging registers; bytes past src_bytes are written as zero.
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint64_t lds64(uint32_t addr) {
  uint64_t v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void lds128(uint32_t addr, uint64_t& a, uint64_t& b) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
__device__ __forceinline__ void sts64(uint32_t addr, uint64_t v) {
  asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void lds128u(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c,
                                        uint32_t& d) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
               : "r"(addr));
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c,
                                        uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c),
               "r"(d)
               : "memory");
}
// (a & ~M) | (b & M) as ONE LOP3 (written in C the compiler spends a LOP3 per constant, M and ~M)
template <uint32_t M>
__device__ __forceinline__ uint32_t bitsel(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(r) : "r"(a), "r"(b), "n"(M));
  return r;
}
// 8 x 8 bit transposes of the four byte lanes of S[0..7] at once (three butterfly stages, 12 instructions per
// stage):  S'[w] byte b, bit j == S[7 - j] byte b, bit 7 - w.  With plane i's 32-row half (row r at bit 31 - r)
// in S[7 - i]:  S'[w] byte b = low 8 code bits of the half's row 24 - 8 b + w.
template <int J, uint32_t M>
__device__ __forceinline__ void swar8_stage(uint32_t (&S)[8]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if ((k & J) == 0) {
      const uint32_t a = S[k], b = S[k + J];
      S[k] = bitsel<M>(a, b >> J);
      S[k + J] = bitsel<M>(a << J, b);
    }
  }
}
__device__ __forceinline__ void swar8_transpose(uint32_t (&S)[8]) {
  swar8_stage<4, 0x0F0F0F0Fu>(S);
  swar8_stage<2, 0x33333333u>(S);
  swar8_stage<1, 0x55555555u>(S);
}

// streaming (read-once) global loads/stores: keep L1 clean
__device__ __forceinline__ uint64_t ldg_stream64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
// random 8-byte gathers (survivor materialisation): by default an L2 miss fetches 128 bytes from
// DRAM here; the 64-byte hint halves the traffic of a sparse gather (measured, tools/micro/
// gather_gran.cu: 10.9 M random loads out of 4.8 GB take 0.215 ms plain, 0.167 ms with the hint;
// cudaLimitMaxL2FetchGranularity has no effect)
__device__ __forceinline__ uint64_t ldg_gather64(const void* p) {
  uint64_t v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::64B.b64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ldg_stream128(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream128(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
#endif  // __CUDACC__

}  // namespace fle
