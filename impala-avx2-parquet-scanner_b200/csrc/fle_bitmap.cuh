// Device helpers shared by the bitmap kernels (fle_bitmap.cu) and the fused evaluate + deposit kernels
// (fle_eval1_deposit.cu): block reductions, tail masks, bit deposit.
#pragma once
#include "fle_device.cuh"

namespace fle {

constexpr int kBmThreads = 256;
constexpr int kBmWordsPerThread = 8;
constexpr int kBmWordsPerBlock = kBmThreads * kBmWordsPerThread;

__device__ __forceinline__ uint64_t tail_mask_word(int64_t w, int64_t nbits) {
  const int64_t left = nbits - w * 64;
  if (left >= 64) return ~0ull;
  if (left <= 0) return 0ull;
  return (1ull << left) - 1ull;
}

template <int NT>
__device__ __forceinline__ uint64_t block_sum(uint64_t v, uint64_t* warp_buf) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) warp_buf[threadIdx.x >> 5] = v;
  __syncthreads();
  uint64_t t = 0;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) t += warp_buf[w];
  __syncthreads();
  return t;
}

// bit deposit: the i-th set bit of mask (from the LSB) receives bit i of src
__device__ __forceinline__ uint64_t deposit_bits(uint64_t src, uint64_t mask) {
  if (mask == ~0ull) return src;
  const int ones = __popcll(mask);
  if (ones >= 32) {
    // few holes: open a zero gap at every clear position of mask, lowest first
    uint64_t holes = ~mask;
    uint64_t r = src;
    while (holes) {
      const int z = __ffsll((long long)holes) - 1;
      const uint64_t low = (1ull << z) - 1ull;
      r = (r & low) | ((r & ~low) << 1);
      holes &= holes - 1;
    }
    return r & mask;
  }
  uint64_t r = 0, m = mask;
  int i = 0;
  while (m) {
    const int p = __ffsll((long long)m) - 1;
    r |= ((src >> i) & 1ull) << p;
    m &= m - 1;
    ++i;
  }
  return r;
}

// not-null word w of a column whose def levels are 1-bit FLE planes (max_def_level == 1): the plane word of group w,
// bit-reversed (row k sits at bit 63 - k), rows past the end cleared
__device__ __forceinline__ uint64_t def_word(const uint64_t* __restrict__ def_planes, int64_t w, int64_t nwords, int64_t n_root) {
  return w < nwords ? (__brevll(def_planes[w]) & tail_mask_word(w, n_root)) : 0ull;
}

}  // namespace fle
