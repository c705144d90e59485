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

// ------------------------------------------------------------------------------------------
// Cross-GPU sum of the two totals, run by the first warp of the LAST CTA of a launch (the one that
// owns the totals) -- the all-reduce the scan ends with, fused into the scan kernel:
// lane r posts {a, b} into entry [seq % kGroupSlots][my rank] of rank r's mailbox with peer stores
// (NVLink / NVSwitch; data, system-wide fence, then the sequence number as the release flag), then
// waits for rank r's entry of the same sequence number in its OWN mailbox and the warp sums the
// entries.  A rank posts launch seq+1 only after it has read every entry of launch seq (stream
// order), and launch seq+2 only after every peer has posted seq+1, i.e. after every peer is done
// reading seq: two slots would do, kGroupSlots = 4.  A peer that does not show up within
// grp_timeout_ns is counted in grp_state[1] (the host checks it) and its entry taken as 0.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t group_ld_acquire(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void group_st_release(uint64_t* p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t group_now_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void group_allreduce_warp(uint64_t* const* mail, unsigned long long* state, int world,
                                                     int rank, uint64_t timeout_ns, uint64_t& a, uint64_t& b) {
  const int lane = threadIdx.x & 31;
  const uint64_t seq = *((volatile unsigned long long*)state) + 1ull;
  const uint32_t slot = (uint32_t)(seq % (uint64_t)kGroupSlots);
  uint64_t va = 0, vb = 0;
  bool lost = false;
  if (lane < world) {
    uint64_t* dst = mail[lane] + ((size_t)slot * FLE_MAX_GROUP + rank) * kGroupEntryWords;
    *((volatile uint64_t*)(dst + 1)) = a;
    *((volatile uint64_t*)(dst + 2)) = b;
    __threadfence_system();
    group_st_release(dst, seq);
    const uint64_t* src = mail[rank] + ((size_t)slot * FLE_MAX_GROUP + lane) * kGroupEntryWords;
    const uint64_t t0 = group_now_ns();
    while (group_ld_acquire(src) != seq) {
      if (group_now_ns() - t0 > timeout_ns) {
        lost = true;
        break;
      }
    }
    if (!lost) {
      va = *((const volatile uint64_t*)(src + 1));
      vb = *((const volatile uint64_t*)(src + 2));
    }
  }
  const unsigned nlost = __popc(__ballot_sync(0xffffffffu, lost));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    va += __shfl_xor_sync(0xffffffffu, va, o);
    vb += __shfl_xor_sync(0xffffffffu, vb, o);
  }
  a = va;
  b = vb;
  if (lane == 0) {
    state[0] = seq;
    if (nlost) state[1] += nlost;
  }
}

// block-wide reduction of the popcounts (and, for fused aggregates, of a second 64-bit sum kept in
// partial[gridDim.x ...]) + "last block sums", so the launch needs no memset/atomics on the
// result slots.  With a group attached (P.grp_world > 1) the last block also sums the totals over
// the GPUs of the group before writing them.
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
    if (threadIdx.x < 32) {
      uint64_t tot = 0, tot2 = 0;
#pragma unroll
      for (int w = 0; w < NT / 32; ++w) {
        tot += warp_cnt[w];
        if (WITH_SUM) tot2 += warp_sum[w];
      }
      if (P.grp_world > 1)
        group_allreduce_warp(P.grp_mail, P.grp_state, P.grp_world, P.grp_rank, P.grp_timeout_ns, tot, tot2);
      if (threadIdx.x == 0) {
        if (P.count_out) *P.count_out = tot;
        if (WITH_SUM && P.agg_sum_out) *P.agg_sum_out = (int64_t)tot2;
        *P.ticket = 0;  // self-reset for the next launch on this stream
        if (P.tile_counter) *P.tile_counter = 0;
      }
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
