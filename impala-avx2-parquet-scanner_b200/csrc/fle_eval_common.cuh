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
// Cross-GPU sum of the two totals, run by the LAST CTA of a launch (the one that owns the totals) --
// the all-reduce the scan ends with, fused into the scan kernel.
//   post(seq):    lane r writes the entry [seq % kGroupSlots][my rank] of rank r's mailbox with peer stores
//                 (NVLink / NVSwitch).  An entry is four 64-bit words {seq:32 | a.lo, seq:32 | a.hi,
//                 seq:32 | b.lo, seq:32 | b.hi}: every word carries its own sequence tag and 64-bit stores
//                 are single-copy atomic, so the words need no ordering among themselves -- no fence, no
//                 separate flag, the stores are fire-and-forget;
//   collect(seq): lane r polls rank r's entry in the OWN mailbox until all four tags read seq; the warp sums.
// Immediate mode: post(seq), collect(seq), write the totals -- results are ordered on the stream like any
// kernel output, at the price of one NVLink round trip and of running the GPUs in lockstep.
// Deferred mode (context option "group_deferred"): post(seq) only; the NEXT exchange of the context (or
// fle_group_flush) collects seq -- a second warp does it while the first one posts, and by then the peers'
// entries have long arrived -- and writes the totals where call seq wanted them.  The exchange then adds
// no dependent memory round trip to the kernel's tail and the GPUs may drift apart by one query.
// Sequence numbers and the pending call's pointers are kept by the HOST (every member issues the same
// sequence of group calls) and travel as kernel parameters: the tail reads no device-side state.
// Slot reuse: a rank posts seq only after its previous launch has collected seq-2 (deferred) or seq-1
// (immediate) from every peer, and a peer posts its seq-1 only after its launch seq-2 is complete, i.e.
// after it has consumed everything up to seq-3: entries seq-3 .. seq can be live, kGroupSlots = 8 > 4.
// A peer that does not show up within grp_timeout_ns is counted in *grp_lost (the host checks it) and its
// entry taken as 0.
// ------------------------------------------------------------------------------------------
struct GroupView {
  uint64_t* const* mail;
  unsigned long long* lost;
  int world, rank;
  uint64_t timeout_ns;
};

__device__ __forceinline__ uint64_t group_now_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void group_post(const GroupView& g, uint32_t seq, uint64_t a, uint64_t b) {
  const int lane = threadIdx.x & 31;
  if (lane < g.world) {
    uint64_t* dst = g.mail[lane] + ((size_t)(seq % (uint32_t)kGroupSlots) * FLE_MAX_GROUP + g.rank) * kGroupEntryWords;
    const uint64_t tag = (uint64_t)seq << 32;
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(tag | (a & 0xffffffffull)), "l"(tag | (a >> 32))
                 : "memory");
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst + 2), "l"(tag | (b & 0xffffffffull)), "l"(tag | (b >> 32))
                 : "memory");
  }
}

__device__ __forceinline__ void group_collect(const GroupView& g, uint32_t seq, uint64_t& a, uint64_t& b) {
  const int lane = threadIdx.x & 31;
  uint64_t va = 0, vb = 0;
  bool lost = false;
  if (lane < g.world) {
    const uint64_t* src = g.mail[g.rank] + ((size_t)(seq % (uint32_t)kGroupSlots) * FLE_MAX_GROUP + lane) * kGroupEntryWords;
    uint64_t w0, w1, w2, w3;
    uint64_t t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
      asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
      asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w2), "=l"(w3) : "l"(src + 2) : "memory");
      if ((uint32_t)(w0 >> 32) == seq && (uint32_t)(w1 >> 32) == seq && (uint32_t)(w2 >> 32) == seq && (uint32_t)(w3 >> 32) == seq) break;
      if ((spin & 63u) == 63u) {
        const uint64_t now = group_now_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > g.timeout_ns) {
          lost = true;
          break;
        }
      }
    }
    if (!lost) {
      va = (w0 & 0xffffffffull) | (w1 << 32);
      vb = (w2 & 0xffffffffull) | (w3 << 32);
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
  if (lane == 0 && nlost) atomicAdd(g.lost, (unsigned long long)nlost);
}

// collects what an earlier deferred exchange left pending and writes it where that call wanted it
// (warp-collective)
__device__ __forceinline__ void group_collect_pending(const GroupView& g, uint32_t pend_seq, uint64_t* pend_count,
                                                      uint64_t* pend_sum) {
  if (pend_seq == 0) return;
  uint64_t pa, pb;
  group_collect(g, pend_seq, pa, pb);
  if ((threadIdx.x & 31) == 0) {
    if (pend_count) *pend_count = pa;
    if (pend_sum) *pend_sum = pb;
  }
}

// Two consecutive plane words at ANY byte address, from 32-bit loads and a funnel shift (skewed columns,
// EvalParams::plane_skew).  Up to 3 bytes before the first and 4 after the last word are touched: inside the
// staged superset for shared memory, inside the column buffer's slack for global memory.
__device__ __forceinline__ void lds_pair_unaligned(uint32_t addr, bool second, uint64_t& w0, uint64_t& w1) {
  const uint32_t b = addr & ~3u, sh = (addr & 3u) * 8u;
  const uint32_t v0 = lds32(b), v1 = lds32(b + 4), v2 = lds32(b + 8);
  w0 = ((uint64_t)__funnelshift_r(v1, v2, sh) << 32) | __funnelshift_r(v0, v1, sh);
  w1 = 0;
  if (second) {
    const uint32_t v3 = lds32(b + 12), v4 = lds32(b + 16);
    w1 = ((uint64_t)__funnelshift_r(v3, v4, sh) << 32) | __funnelshift_r(v2, v3, sh);
  }
}
// All BW plane words of a group at ANY byte address of shared memory: 128-bit loads of the 16-byte chunks around
// it (the same conflict-free access pattern as the aligned layout: the eight lanes of a quarter warp sit in eight
// segments whose images are 3 chunks apart modulo 8), then the words are moved down by the thread's word offset
// with two rounds of selects and funnel-shifted by its byte offset.  32-bit loads would be simpler and are 8-way
// bank-conflicted with this layout (0.54 of roofline at bw 12).
template <int BW>
__device__ __forceinline__ void lds_planes_unaligned(uint32_t addr, uint64_t (&w)[BW]) {
  constexpr int NCH = (BW * 8 + 19 + 15) / 16;  // chunks that hold bytes [addr & 15, (addr & 15) + BW * 8 + 4)
  const uint32_t a16 = addr & ~15u, e = (addr >> 2) & 3u, sh = (addr & 3u) * 8u;
  uint32_t V[NCH * 4 + 3];
#pragma unroll
  for (int c = 0; c < NCH; ++c) lds128u(a16 + 16 * c, V[4 * c], V[4 * c + 1], V[4 * c + 2], V[4 * c + 3]);
  V[NCH * 4] = V[NCH * 4 + 1] = V[NCH * 4 + 2] = 0u;
#pragma unroll
  for (int j = 0; j < NCH * 4; ++j) V[j] = (e & 1u) ? V[j + 1] : V[j];
#pragma unroll
  for (int j = 0; j < NCH * 4; ++j) V[j] = (e & 2u) ? V[j + 2] : V[j];
#pragma unroll
  for (int i = 0; i < BW; ++i)
    w[i] = ((uint64_t)__funnelshift_r(V[2 * i + 1], V[2 * i + 2], sh) << 32) | __funnelshift_r(V[2 * i], V[2 * i + 1], sh);
}

__device__ __forceinline__ void ldg_pair_unaligned(const void* p, bool second, uint64_t& w0, uint64_t& w1) {
  const unsigned long long a = reinterpret_cast<unsigned long long>(p);
  const uint32_t* b = reinterpret_cast<const uint32_t*>(a & ~3ull);
  const uint32_t sh = (uint32_t)(a & 3ull) * 8u;
  const uint32_t v0 = __ldg(b), v1 = __ldg(b + 1), v2 = __ldg(b + 2);
  w0 = ((uint64_t)__funnelshift_r(v1, v2, sh) << 32) | __funnelshift_r(v0, v1, sh);
  w1 = 0;
  if (second) {
    const uint32_t v3 = __ldg(b + 3), v4 = __ldg(b + 4);
    w1 = ((uint64_t)__funnelshift_r(v3, v4, sh) << 32) | __funnelshift_r(v2, v3, sh);
  }
}

// the optional second operand of the single-IN-list kernels (EvalParams::comb_*)
__device__ __forceinline__ uint64_t comb_load(const EvalParams& P, int64_t g) {
  return P.comb_op ? *((const volatile uint64_t*)(P.comb_bm + g)) : 0ull;
}
__device__ __forceinline__ uint64_t comb_apply(const EvalParams& P, uint64_t r, uint64_t other) {
  if (P.comb_op == FOLD_AND) return r & other;
  if (P.comb_op == FOLD_OR) return r | other;
  return r;
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
    const bool grp = P.grp_world > 1;
    const GroupView gv{P.grp_mail, P.grp_lost, P.grp_world, P.grp_rank, P.grp_timeout_ns};
    if (threadIdx.x < 32) {
      uint64_t tot = 0, tot2 = 0;
#pragma unroll
      for (int w = 0; w < NT / 32; ++w) {
        tot += warp_cnt[w];
        if (WITH_SUM) tot2 += warp_sum[w];
      }
      bool write_now = true;
      if (grp) {
        group_post(gv, P.grp_seq, tot, tot2);
        if (P.grp_deferred) write_now = false;
        else group_collect(gv, P.grp_seq, tot, tot2);
      }
      if (threadIdx.x == 0) {
        if (write_now && P.count_out) *P.count_out = tot;
        if (write_now && WITH_SUM && P.agg_sum_out) *P.agg_sum_out = (int64_t)tot2;
        *P.ticket = 0;  // self-reset for the next launch on this stream
        if (P.tile_counter) *P.tile_counter = 0;
      }
    } else if (threadIdx.x < 64 && grp) {
      // the totals an earlier deferred call left pending, while the first warp posts
      group_collect_pending(gv, P.grp_pend_seq, P.grp_pend_count, P.grp_pend_sum);
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
  // (a skewed column stages 16 more bytes per segment: the aligned superset of the segment's byte range)
  for (int c = 0; c < P.ncols; ++c) total += (uint32_t)T * P.bw[c] * 8u + (P.plane_skew[c] ? 128u : 0u);
  mbar_arrive_expect_tx(bar, total);
  for (int c = 0; c < P.ncols; ++c) {
    const uint32_t seg_bytes = (uint32_t)(T >> 3) * P.bw[c] * 8u;
    const uint32_t copy_bytes = seg_bytes + (P.plane_skew[c] ? 16u : 0u);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(P.planes[c]) + (uint64_t)tile * T * P.bw[c] * 8u;
    const uint32_t dst = stage0 + stage * P.stage_bytes + P.col_off[c];
#pragma unroll
    for (int s = 0; s < 8; ++s)
      tma_load_1d(dst + s * P.seg_stride[c], src + (uint64_t)s * seg_bytes, copy_bytes, bar);
  }
}

}  // namespace fle
