// Fused-eval kernel specialised per predicate PROGRAM (shape) at run time.
//
// The reference evaluates a conjunct tree by interpreting SimplePredicate nodes leaf by leaf
// (simple-predicates.h:113-163 -> HdfsParquetScanner::Eq..In -> FleDecoder::Eq..In); Impala's
// own answer to interpretation overhead is per-query code generation.  Same idea here: the
// generic kernels of fle_eval.cu interpret the lowered pass list per 64-row group (~500
// instructions for a three-column program), this file is the SAME tile pipeline with the
// per-group function `jit_eval_group` written out by the host for one program shape
// (fle_jit.cu: column bit widths, chain kinds, slot/fold structure) -- the bare plane algebra,
// every literal mask a constant-bank operand of its LOP3, no loops, no slot memory.
// Literals stay kernel parameters (JitArgs), so all queries of one shape share one kernel.
//
// Compiled by NVRTC with "fle_jit_program.h" generated per shape (tests/test_jit_codegen.py
// compiles a set of shapes on the CPU-only box, so syntax errors surface without a GPU).
#pragma once
#include "fle_eval_common.cuh"

namespace fle {

// one plane word into one '<' / '<=' borrow chain (both 32-row halves)
__device__ __forceinline__ void jit_lt(uint32_t& al, uint32_t& ah, uint64_t w, uint32_t mk) {
  const uint32_t wl = (uint32_t)w, wh = (uint32_t)(w >> 32);
  al = (~wl & mk) | (~(wl ^ mk) & al);
  ah = (~wh & mk) | (~(wh ^ mk) & ah);
}
// one plane word into one equality chain
__device__ __forceinline__ void jit_eq(uint32_t& al, uint32_t& ah, uint64_t w, uint32_t mk) {
  const uint32_t wl = (uint32_t)w, wh = (uint32_t)(w >> 32);
  al = al & ~(wl ^ mk);
  ah = ah & ~(wh ^ mk);
}
__device__ __forceinline__ uint64_t jit_res(uint32_t al, uint32_t ah, uint32_t nm) {
  return ((uint64_t)(ah ^ nm) << 32) | (al ^ nm);
}

// a group's planes in a TMA stage (per-column shared addresses) ...
template <int NCOLS>
struct JitSmemGroup {
  uint32_t a[NCOLS];
  // every plane word of column C's group (skewed columns: whole-chunk loads + funnel shift)
  template <int C, int BW>
  __device__ __forceinline__ void load_all(uint64_t (&w)[BW]) const {
    lds_planes_unaligned<BW>(a[C], w);
  }
  template <int C, int BW>
  __device__ __forceinline__ void load_pair(int i, uint64_t& w0, uint64_t& w1) const {
#ifdef FLE_JIT_SKEW
    lds_pair_unaligned(a[C] + i * 8, i + 1 < BW, w0, w1);  // (generated code for skewed columns uses load_all)
#else
    if constexpr (BW % 2 == 0) {
      lds128(a[C] + i * 8, w0, w1);
    } else {
      w0 = lds64(a[C] + i * 8);
      w1 = (i + 1 < BW) ? lds64(a[C] + (i + 1) * 8) : 0ull;
    }
#endif
  }
};
// ... or in global memory (ragged remainder of the row range)
template <int NCOLS>
struct JitGlobalGroup {
  const uint64_t* p[NCOLS];
  template <int C, int BW>
  __device__ __forceinline__ void load_all(uint64_t (&w)[BW]) const {
#pragma unroll
    for (int i = 0; i < BW; i += 2) {
      uint64_t a_, b_;
      ldg_pair_unaligned(reinterpret_cast<const uint8_t*>(p[C]) + i * 8, i + 1 < BW, a_, b_);
      w[i] = a_;
      if (i + 1 < BW) w[i + 1 < BW ? i + 1 : i] = b_;
    }
  }
  template <int C, int BW>
  __device__ __forceinline__ void load_pair(int i, uint64_t& w0, uint64_t& w1) const {
#ifdef FLE_JIT_SKEW
    ldg_pair_unaligned(reinterpret_cast<const uint8_t*>(p[C]) + i * 8, i + 1 < BW, w0, w1);
#else
    w0 = __ldg(p[C] + i);
    w1 = (i + 1 < BW) ? __ldg(p[C] + i + 1) : 0ull;
#endif
  }
};

// OR of equality chains over a literal list of run-time length (IN lists too long to unroll),
// four literals per sweep; the column's planes are already in registers
template <int BW>
__device__ __forceinline__ uint64_t jit_in_list(const uint64_t (&w)[BW], const uint32_t* lits, int lcnt) {
  uint64_t any = 0;
  for (int j = 0; j < lcnt; j += 4) {
    uint32_t lit[4], al[4], ah[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      lit[k] = lits[min(j + k, lcnt - 1)];
      al[k] = ah[k] = ~0u;
    }
#pragma unroll
    for (int i = 0; i < BW; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) jit_eq(al[k], ah[k], w[i], 0u - ((lit[k] >> i) & 1u));
    }
    any |= ((uint64_t)((ah[0] | ah[1]) | (ah[2] | ah[3])) << 32) | ((al[0] | al[1]) | (al[2] | al[3]));
  }
  return any;
}

}  // namespace fle

// Generated per program shape:
//   constexpr int kJitNCols;  struct JitArgs;
//   jit_stage_addr(P, base, seg, jrow, a[])   shared address of the thread's first group per column
//   jit_advance(a[])                          ... to its next group (NT/8 groups further)
//   jit_global_addr(P, g, p[])                plane pointers of group g
//   jit_eval_group(P, A, ld, lits_smem)       plane-order result word of one group
//   [FLE_JIT_AGG]  kJitAggCol, kJitAggBw, jit_agg_code_smem(addr, sh), jit_agg_code_global(ptr, sh):
//                  code of the row at plane bit sh of a group of the aggregate's column
#include "fle_jit_program.h"

#ifdef FLE_JIT_AGG
namespace fle {
// Fused aggregate: acc += plain[row] * dict[code of the row in the aggregate's dictionary column]
// for every selected row.  A warp holds 32 groups' selection words; their survivors are walked in
// FLATTENED order -- survivor t goes to lane t % 32 -- so that every lane keeps an independent
// random load of plain[] in flight however the survivors are spread over the groups (one thread
// looping over its own group's survivors serialises dependent loads and made the fused kernel
// slower than two separate ones).  Each lane lists its own survivors as (owner lane, row) entries
// at its exclusive prefix position in a per-warp shared array; lane t then simply reads entry t.
// The code is picked out of the column's planes, which are still in the stage the predicate has
// just read.  The load of plain[] is DEFERRED: the first two rounds park the price (still in
// flight) and the dictionary value in registers and multiply them at the next call -- after another
// group's predicate work -- so the load latency overlaps the scan instead of stalling the warp
// (all warps of a CTA reach this point together, tile after tile).  Must be called by all 32 lanes.
constexpr uint32_t kJitAggSlots = 128;  // survivors per 32 groups on the fast path (selectivity <~ 6 %)

struct JitAggScratch {  // per warp, in shared memory
  uint32_t prefix[33];
  uint32_t addr[32];
  uint64_t word[32];
  int64_t group[32];
  uint16_t slot[kJitAggSlots];
};

struct JitAggPending {
  int64_t price[2], d[2];
};

__device__ __forceinline__ int jit_nth_set_bit64(uint64_t x, int n) {  // position of the n-th (0-based) set bit
  const uint32_t lo = (uint32_t)x;
  const int clo = __popc(lo);
  if (n < clo) return (int)__fns(lo, 0, n + 1);
  return 32 + (int)__fns((uint32_t)(x >> 32), 0, n - clo + 1);
}

template <bool STAGED>
__device__ __forceinline__ uint32_t jit_agg_code_of(const EvalParams& P, const JitAggScratch& sc, int j, int64_t gg, int k) {
  return STAGED ? jit_agg_code_smem(sc.addr[j], 63 - k)
                : jit_agg_code_global(P.planes[kJitAggCol] + gg * kJitAggBw, 63 - k);
}

template <bool STAGED>
__device__ __forceinline__ void jit_agg_warp(const EvalParams& P, JitAggScratch& sc, uint64_t r, int64_t g, uint32_t a,
                                             int64_t& acc, JitAggPending& pend) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t mine = __popcll(r);
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (uint32_t)o) incl += y;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  if (total == 0) return;
  __syncwarp();
  sc.group[lane] = g;
  if (STAGED) sc.addr[lane] = a;
  if (total <= kJitAggSlots) {
    uint32_t pos = incl - mine;
    for (uint64_t m = r; m; m &= m - 1) sc.slot[pos++] = (uint16_t)((lane << 8) | (uint32_t)(__ffsll((long long)m) - 1));
    __syncwarp();
    acc += pend.price[0] * pend.d[0] + pend.price[1] * pend.d[1];  // parked by the previous call
    pend.price[0] = pend.price[1] = 0;
    pend.d[0] = pend.d[1] = 0;
#pragma unroll
    for (int round = 0; round < 2; ++round) {  // loads land straight in the parked registers
      const uint32_t t = lane + 32u * round;
      if (t < total) {
        const uint32_t e = sc.slot[t];
        const int j = (int)(e >> 8), k = (int)(e & 63u);
        const int64_t gg = sc.group[j];
        pend.price[round] = (int64_t)ldg_gather64(P.agg_plain + gg * 64 + k);
        const uint32_t code = jit_agg_code_of<STAGED>(P, sc, j, gg, k);
        pend.d[round] = code < (uint32_t)P.agg_ndict ? __ldg(P.agg_dict + code) : 0;
      }
    }
    for (uint32_t t = lane + 64; t < total; t += 32) {  // further rounds: consumed at once
      const uint32_t e = sc.slot[t];
      const int j = (int)(e >> 8), k = (int)(e & 63u);
      const int64_t gg = sc.group[j];
      const int64_t price = (int64_t)ldg_gather64(P.agg_plain + gg * 64 + k);
      const uint32_t code = jit_agg_code_of<STAGED>(P, sc, j, gg, k);
      acc += price * (code < (uint32_t)P.agg_ndict ? __ldg(P.agg_dict + code) : 0);
    }
    return;
  }
  // dense selections: owner lane by binary search over the prefix sums, row by n-th set bit
  sc.prefix[lane] = incl - mine;  // exclusive
  if (lane == 31) sc.prefix[32] = total;
  sc.word[lane] = r;
  __syncwarp();
  for (uint32_t t = lane; t < total; t += 32) {
    int j = 0;  // prefix[j] <= t < prefix[j + 1]
#pragma unroll
    for (int step = 16; step > 0; step >>= 1)
      if (sc.prefix[j + step] <= t) j += step;
    const int k = jit_nth_set_bit64(sc.word[j], (int)(t - sc.prefix[j]));
    const int64_t gg = sc.group[j];
    const int64_t price = (int64_t)ldg_gather64(P.agg_plain + gg * 64 + k);
    const uint32_t code = jit_agg_code_of<STAGED>(P, sc, j, gg, k);
    acc += price * (code < (uint32_t)P.agg_ndict ? __ldg(P.agg_dict + code) : 0);
  }
}
}  // namespace fle
#endif

extern "C" __global__ void __launch_bounds__(256)
fle_jit_kernel(const __grid_constant__ fle::EvalParams P, const __grid_constant__ fle::JitArgs A) {
  using namespace fle;
  constexpr int NT = 256;
  extern __shared__ __align__(128) uint8_t smem[];
  uint32_t* lits_smem = reinterpret_cast<uint32_t*>(smem + P.smem_lits_off);
  const uint32_t bar0 = smem_u32(smem + P.smem_bar_off);
  const uint32_t stage0 = smem_u32(smem + P.smem_stage_off);
  const int gpt = P.gpt;
  const int T = NT * gpt;
  const int seg_groups = T >> 3;
  const int nst = P.nstages;  // 0: unstaged launch, every group takes the global-memory loop below
#ifdef FLE_JIT_SKEW
  const int64_t ntiles = nst > 0 ? P.staged_groups / T : 0;
#else
  const int64_t ntiles = nst > 0 ? P.ngroups / T : 0;
#endif

  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) mbar_init(bar0 + s * 8, 1);
    mbar_fence_init();
  }
  load_in_lits(P, lits_smem);
  __syncthreads();
  // dynamic tile scheduler, as in eval1_kernel (fle_eval1.cu): the producer thread claims tiles from a
  // global counter one iteration ahead and hands the tile id to the consumers through shared memory
  __shared__ int64_t stage_tile[8];  // tile held by each stage, -1 = no more tiles
  unsigned int next_claim = 0;
  auto arm = [&](int s, unsigned int claimed) {  // producer thread only
    const int64_t tile = (int64_t)claimed;
    if (tile < ntiles) {
      stage_tile[s] = tile;
      issue_tile(P, bar0, stage0, s, tile, T);
    } else {
      stage_tile[s] = -1;
      mbar_arrive(bar0 + s * 8);  // completes the phase without data
    }
  };
  if (threadIdx.x == 0 && nst > 0) {
    for (int s = 0; s < nst - 1; ++s) arm(s, atomicAdd(P.tile_counter, 1u));
    next_claim = atomicAdd(P.tile_counter, 1u);
  }

  const uint32_t seg = threadIdx.x & 7, jrow = threadIdx.x >> 3;
  const int64_t thread_g = (int64_t)seg * seg_groups + jrow;
  uint64_t cnt = 0;
#ifdef FLE_JIT_AGG
  int64_t acc = 0;
  __shared__ JitAggScratch agg_scratch[NT / 32];
  JitAggScratch& sc = agg_scratch[threadIdx.x >> 5];
  JitAggPending pend[FLE_JIT_GPT] = {};
#endif
  int stage = 0;
  uint32_t parity = 0;
  while (nst > 0) {
    __syncthreads();  // everyone is done with the stage consumed in the previous iteration
    if (threadIdx.x == 0) {
      int st2 = stage + nst - 1;
      if (st2 >= nst) st2 -= nst;
      arm(st2, next_claim);
      next_claim = atomicAdd(P.tile_counter, 1u);
    }
    mbar_wait(bar0 + stage * 8, parity);
    const int64_t tile = stage_tile[stage];
    if (tile < 0) break;  // (every later stage holds -1 as well: nothing is in flight)
    JitSmemGroup<kJitNCols> ld;
    jit_stage_addr(P, stage0 + stage * P.stage_bytes, seg, jrow, ld.a);
    int64_t g = tile * T + thread_g;
#ifdef FLE_JIT_AGG
    // fixed trip count (the host launches aggregate kernels with gpt == FLE_JIT_GPT): the survivors
    // parked by group q of this tile are consumed by group q of the NEXT tile, a whole tile period
    // (~2 us) later, which covers the latency of their random loads
#pragma unroll
    for (int q = 0; q < FLE_JIT_GPT; ++q) {
      const uint64_t r = finish_word(jit_eval_group(P, A, ld, lits_smem), g, P.nrows);
      if (P.bitmap) P.bitmap[g] = r;
      cnt += __popcll(r);
      jit_agg_warp<true>(P, sc, r, g, ld.a[kJitAggCol], acc, pend[q]);
      jit_advance(ld.a);
      g += NT / 8;
    }
#else
#pragma unroll FLE_JIT_UNROLL
    for (int q = 0; q < gpt; ++q) {
      const uint64_t r = finish_word(jit_eval_group(P, A, ld, lits_smem), g, P.nrows);
      P.bitmap[g] = r;
      cnt += __popcll(r);
      jit_advance(ld.a);
      g += NT / 8;
    }
#endif
    if (++stage == nst) {
      stage = 0;
      parity ^= 1u;
    }
  }
  // ragged remainder (< T groups), planes straight from global memory
#ifdef FLE_JIT_AGG
  // (warp-uniform trip count: the aggregate's warp-wide walk needs all 32 lanes)
  for (int64_t g = ntiles * T + (int64_t)blockIdx.x * NT + threadIdx.x; g - (threadIdx.x & 31) < P.ngroups;
       g += (int64_t)gridDim.x * NT) {
    uint64_t r = 0;
    if (g < P.ngroups) {
      JitGlobalGroup<kJitNCols> ld;
      jit_global_addr(P, g, ld.p);
      r = finish_word(jit_eval_group(P, A, ld, lits_smem), g, P.nrows);
      if (P.bitmap) P.bitmap[g] = r;
      cnt += __popcll(r);
    }
    jit_agg_warp<false>(P, sc, r, g, 0u, acc, pend[0]);
  }
#else
  for (int64_t g = ntiles * T + (int64_t)blockIdx.x * NT + threadIdx.x; g < P.ngroups;
       g += (int64_t)gridDim.x * NT) {
    JitGlobalGroup<kJitNCols> ld;
    jit_global_addr(P, g, ld.p);
    const uint64_t r = finish_word(jit_eval_group(P, A, ld, lits_smem), g, P.nrows);
    P.bitmap[g] = r;
    cnt += __popcll(r);
  }
#endif
#ifdef FLE_JIT_AGG
#pragma unroll
  for (int q = 0; q < FLE_JIT_GPT; ++q) acc += pend[q].price[0] * pend[q].d[0] + pend[q].price[1] * pend[q].d[1];
  publish_totals<NT, true>(P, cnt, (uint64_t)acc);
#else
  publish_count<NT>(P, cnt);
#endif
}
