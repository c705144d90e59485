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
  template <int C, int BW>
  __device__ __forceinline__ void load_pair(int i, uint64_t& w0, uint64_t& w1) const {
    if constexpr (BW % 2 == 0) {
      lds128(a[C] + i * 8, w0, w1);
    } else {
      w0 = lds64(a[C] + i * 8);
      w1 = (i + 1 < BW) ? lds64(a[C] + (i + 1) * 8) : 0ull;
    }
  }
};
// ... or in global memory (ragged remainder of the row range)
template <int NCOLS>
struct JitGlobalGroup {
  const uint64_t* p[NCOLS];
  template <int C, int BW>
  __device__ __forceinline__ void load_pair(int i, uint64_t& w0, uint64_t& w1) const {
    w0 = __ldg(p[C] + i);
    w1 = (i + 1 < BW) ? __ldg(p[C] + i + 1) : 0ull;
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
#include "fle_jit_program.h"

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
  const int64_t ntiles = nst > 0 ? P.ngroups / T : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) mbar_init(bar0 + s * 8, 1);
    mbar_fence_init();
  }
  load_in_lits(P, lits_smem);
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int s = 0; s < nst - 1; ++s) {
      const int64_t tile = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
      if (tile < ntiles) issue_tile(P, bar0, stage0, s, tile, T);
    }
  }

  const uint32_t seg = threadIdx.x & 7, jrow = threadIdx.x >> 3;
  const int64_t thread_g = (int64_t)seg * seg_groups + jrow;
  uint64_t cnt = 0;
  int stage = 0;
  uint32_t parity = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();  // everyone is done with the stage consumed in the previous iteration
    if (threadIdx.x == 0) {
      const int64_t nxt = tile + (int64_t)(nst - 1) * gridDim.x;
      int st2 = stage + nst - 1;
      if (st2 >= nst) st2 -= nst;
      if (nxt < ntiles) issue_tile(P, bar0, stage0, st2, nxt, T);
    }
    mbar_wait(bar0 + stage * 8, parity);
    JitSmemGroup<kJitNCols> ld;
    jit_stage_addr(P, stage0 + stage * P.stage_bytes, seg, jrow, ld.a);
    int64_t g = tile * T + thread_g;
    for (int q = 0; q < gpt; ++q) {
      const uint64_t r = finish_word(jit_eval_group(P, A, ld, lits_smem), g, P.nrows);
      P.bitmap[g] = r;
      cnt += __popcll(r);
      jit_advance(ld.a);
      g += NT / 8;
    }
    if (++stage == nst) {
      stage = 0;
      parity ^= 1u;
    }
  }
  // ragged remainder (< T groups), planes straight from global memory
  for (int64_t g = ntiles * T + (int64_t)blockIdx.x * NT + threadIdx.x; g < P.ngroups;
       g += (int64_t)gridDim.x * NT) {
    JitGlobalGroup<kJitNCols> ld;
    jit_global_addr(P, g, ld.p);
    const uint64_t r = finish_word(jit_eval_group(P, A, ld, lits_smem), g, P.nrows);
    P.bitmap[g] = r;
    cnt += __popcll(r);
  }
  publish_count<NT>(P, cnt);
}
