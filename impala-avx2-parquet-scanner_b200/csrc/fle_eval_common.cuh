// Pieces shared by the fused-eval kernels (fle_eval.cu: generic program kernels,
// fle_eval1.cu: single-pass kernels specialised per bit width).
#pragma once
#include "fle_device.cuh"
#include "fle_params.h"

namespace fle {

// plane order -> row order (fle-encoding.h:7989-7998 does this with six swap steps), then
// clear the bits of rows past the end (the reference pushes only num_rows bits, :8000-8006).
__device__ __forceinline__ uint64_t finish_word(uint64_t planes_order, int64_t g, int64_t nrows) {
  uint64_t r = __brevll(planes_order);
  const int64_t rows_left = nrows - g * 64;
  if (rows_left < 64) r &= (rows_left <= 0) ? 0ull : ((1ull << rows_left) - 1ull);
  return r;
}

// block-wide reduction of the popcounts (and, for fused aggregates, of a second 64-bit sum kept in
// partial[gridDim.x ...]) + "last block sums", so the launch needs no memset/atomics on the
// result slots.
template <int NT, bool WITH_SUM>
__device__ __forceinline__ void publish_totals(const EvalParams& P, uint64_t cnt, uint64_t sum) {
  __shared__ uint64_t warp_cnt[NT / 32];
  __shared__ uint64_t warp_sum[WITH_SUM ? NT / 32 : 1];
  __shared__ bool is_last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (WITH_SUM) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  }
  if ((threadIdx.x & 31) == 0) {
    warp_cnt[threadIdx.x >> 5] = cnt;
    if (WITH_SUM) warp_sum[threadIdx.x >> 5] = sum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t t = 0, u = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) {
      t += warp_cnt[w];
      if (WITH_SUM) u += warp_sum[w];
    }
    P.partial[blockIdx.x] = t;
    if (WITH_SUM) P.partial[gridDim.x + blockIdx.x] = u;
    __threadfence();
    const unsigned int prev = atomicAdd(P.ticket, 1u);
    is_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    uint64_t t = 0, u = 0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += NT) {
      t += *((volatile uint64_t*)&P.partial[i]);
      if (WITH_SUM) u += *((volatile uint64_t*)&P.partial[gridDim.x + i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      t += __shfl_xor_sync(0xffffffffu, t, o);
      if (WITH_SUM) u += __shfl_xor_sync(0xffffffffu, u, o);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
      warp_cnt[threadIdx.x >> 5] = t;
      if (WITH_SUM) warp_sum[threadIdx.x >> 5] = u;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint64_t tot = 0, tot2 = 0;
#pragma unroll
      for (int w = 0; w < NT / 32; ++w) {
        tot += warp_cnt[w];
        if (WITH_SUM) tot2 += warp_sum[w];
      }
      if (P.count_out) *P.count_out = tot;
      if (WITH_SUM && P.agg_sum_out) *P.agg_sum_out = (int64_t)tot2;
      *P.ticket = 0;  // self-reset for the next launch on this stream
      if (P.tile_counter) *P.tile_counter = 0;
    }
  }
}

template <int NT>
__device__ __forceinline__ void publish_count(const EvalParams& P, uint64_t cnt) {
  publish_totals<NT, false>(P, cnt, 0);
}

__device__ __forceinline__ void load_in_lits(const EvalParams& P, uint32_t* lits_smem) {
  for (int i = threadIdx.x; i < P.n_in_lits; i += blockDim.x) lits_smem[i] = P.in_lits[i];
}


// TMA producer: one tile = T groups of every column -> stage `stage`.  Each column's tile is
// split into 8 segments whose shared-memory images are skewed by 16 bytes (seg_stride), which
// makes the per-thread strided 64/128-bit loads bank-conflict free for every bit width.
__device__ __forceinline__ void issue_tile(const EvalParams& P, uint32_t bar0, uint32_t stage0, int stage,
                                           int64_t tile, int T) {
  const uint32_t bar = bar0 + stage * 8;
  uint32_t total = 0;
  for (int c = 0; c < P.ncols; ++c) total += (uint32_t)T * P.bw[c] * 8u;
  mbar_arrive_expect_tx(bar, total);
  for (int c = 0; c < P.ncols; ++c) {
    const uint32_t seg_bytes = (uint32_t)(T >> 3) * P.bw[c] * 8u;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(P.planes[c]) + (uint64_t)tile * T * P.bw[c] * 8u;
    const uint32_t dst = stage0 + stage * P.stage_bytes + P.col_off[c];
#pragma unroll
    for (int s = 0; s < 8; ++s)
      tma_load_1d(dst + s * P.seg_stride[c], src + (uint64_t)s * seg_bytes, seg_bytes, bar);
  }
}

}  // namespace fle
