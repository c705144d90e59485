// Single-pass fused-eval kernels, one specialisation per <kind, chains, bit width>.
//
// Programs whose leaves all sit on ONE column and fold with ONE operator -- a single compare,
// BETWEEN (= And(Ge, Le)), NOT BETWEEN, a short conjunction or disjunction of
// compares, Eq -- lower to exactly one pass (fle_internal.h).  For those the whole program is
// known at compile time except the literals, so the kernel is the bare plane algebra:
// per 64-row group  BW/2 x LDS.128  +  2*K*BW x LOP3  +  ~20 instructions of epilogue
// (merge, BREV, tail mask, store, POPC).  The literal masks live in registers for the whole
// kernel.  Same TMA/mbarrier tile pipeline and same skewed shared-memory layout as the generic
// kernel in fle_eval.cu; replaces FleDecoder::Eq/Lt/Le/Gt/Ge (fle-encoding.h:7962-8234) and
// the And/OrOperate `&=`/`|=` on top of them (simple-predicates.h:145-163).
#include <algorithm>

#include "fle_eval1.cuh"
#include "fle_internal.h"

namespace fle {

template <bool IS_LT, int K, int BW>
__global__ void __launch_bounds__(256) eval1_kernel(const __grid_constant__ EvalParams P) {
  eval1_body<PassRegs<IS_LT, K, BW>, BW>(P);
}

template <int BW>
__global__ void __launch_bounds__(256) eval_in_mux_kernel(const __grid_constant__ EvalParams P) {
  eval1_body<InMuxRegs<BW>, BW>(P);
}

// ------------------------------------------------------------------------------------------
// Narrow columns (bw 1, 2, 4): a group is 8 / 16 / 32 bytes, so a warp reading consecutive
// groups straight from global memory with 128-bit loads IS coalesced -- no shared-memory
// staging, no per-tile barrier, no mbarrier wait.  At these widths the staged kernel spends
// its time on exactly that per-tile overhead (a 600 M-row bw 1 pass is 150 MB, 23 us at the
// HBM peak) and on 8-byte bitmap stores scattered over the 8 tile segments; here a thread
// owns `kGroups` adjacent groups of each of its U units, all U loads are issued before the
// first use, and bw 1 stores its two result words with one 128-bit store.
// ------------------------------------------------------------------------------------------
template <bool IS_LT, int K, int BW>
__global__ void __launch_bounds__(256) eval1_narrow_kernel(const __grid_constant__ EvalParams P) {
  static_assert(BW == 1 || BW == 2 || BW == 4, "narrow kernel: one unit = 16 or 32 contiguous bytes");
  constexpr int NT = 256, U = kNarrowUnits;
  constexpr int kGroups = BW == 1 ? 2 : 1;   // groups per unit
  constexpr int kVecs = BW == 4 ? 2 : 1;     // 16-byte loads per unit
  PassRegs<IS_LT, K, BW> pr;
  pr.init(P.passes[0]);
  const uint4* src = reinterpret_cast<const uint4*>(P.planes[0]);
  const int64_t nunits = P.ngroups / kGroups;
  const int64_t ntiles = nunits / (NT * U);
  uint64_t cnt = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t u0 = tile * (NT * U) + threadIdx.x;
    uint4 v[U][kVecs];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < kVecs; ++j) v[u][j] = ldg_stream128(src + (u0 + u * NT) * kVecs + j);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t unit = u0 + u * NT;
      if constexpr (BW == 1) {
        RegGroup<1> g0{{((uint64_t)v[u][0].y << 32) | v[u][0].x}}, g1{{((uint64_t)v[u][0].w << 32) | v[u][0].z}};
        const uint64_t r0 = finish_word(pr.run(g0), unit * 2, P.nrows);
        const uint64_t r1 = finish_word(pr.run(g1), unit * 2 + 1, P.nrows);
        stg_stream128(P.bitmap + unit * 2, make_uint4((uint32_t)r0, (uint32_t)(r0 >> 32), (uint32_t)r1, (uint32_t)(r1 >> 32)));
        cnt += __popcll(r0) + __popcll(r1);
      } else {
        RegGroup<BW> g;
#pragma unroll
        for (int j = 0; j < kVecs; ++j) {
          g.w[2 * j] = ((uint64_t)v[u][j].y << 32) | v[u][j].x;
          g.w[2 * j + 1] = ((uint64_t)v[u][j].w << 32) | v[u][j].z;
        }
        const uint64_t r = finish_word(pr.run(g), unit, P.nrows);
        P.bitmap[unit] = r;
        cnt += __popcll(r);
      }
    }
  }
  // ragged remainder (< NT * U units, and the odd last group of bw 1): one group per thread
  for (int64_t g = ntiles * (NT * U) * kGroups + (int64_t)blockIdx.x * NT + threadIdx.x; g < P.ngroups;
       g += (int64_t)gridDim.x * NT) {
    const uint64_t r = finish_word(pr.run(GlobalGroup{P.planes[0] + g * BW}), g, P.nrows);
    P.bitmap[g] = r;
    cnt += __popcll(r);
  }
  publish_count<NT>(P, cnt);
}

// ------------------------------------------------------------------------------------------
// host side: kernel table
// ------------------------------------------------------------------------------------------
namespace {

#define FLE_BW_CASES(X) \
  X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) \
  X(17) X(18) X(19) X(20) X(21) X(22) X(23) X(24) X(25) X(26) X(27) X(28) X(29) X(30) X(31) X(32)

template <bool IS_LT, int K>
const void* eval1_for_bw(int bw) {
  switch (bw) {
#define X(B) case B: return (const void*)eval1_kernel<IS_LT, K, B>;
    FLE_BW_CASES(X)
#undef X
  }
  return nullptr;
}

}  // namespace

template <bool IS_LT, int K>
const void* eval1_narrow_for_bw(int bw) {
  if (bw == 1) return (const void*)eval1_narrow_kernel<IS_LT, K, 1>;
  if (bw == 2) return (const void*)eval1_narrow_kernel<IS_LT, K, 2>;
  if (bw == 4) return (const void*)eval1_narrow_kernel<IS_LT, K, 4>;
  return nullptr;
}

const void* eval1_narrow_kernel_for(int type, int k, int bw) {
  if (type == PASS_LT) {
    if (k == 1) return eval1_narrow_for_bw<true, 1>(bw);
    if (k == 2) return eval1_narrow_for_bw<true, 2>(bw);
    if (k == 4) return eval1_narrow_for_bw<true, 4>(bw);
  } else if (type == PASS_EQ) {
    if (k == 1) return eval1_narrow_for_bw<false, 1>(bw);
    if (k == 4) return eval1_narrow_for_bw<false, 4>(bw);
  }
  return nullptr;
}

const void* eval_in_mux_kernel_for(int bw) {
  switch (bw) {
#define X(B) case B: return (const void*)eval_in_mux_kernel<B>;
    X(1) X(2) X(3) X(4) X(5) X(6) X(7)
#undef X
  }
  return nullptr;
}

const void* eval1_kernel_for(int type, int k, int bw) {
  if (type == PASS_LT) {
    if (k == 1) return eval1_for_bw<true, 1>(bw);
    if (k == 2) return eval1_for_bw<true, 2>(bw);
    if (k == 4) return eval1_for_bw<true, 4>(bw);
  } else if (type == PASS_EQ) {
    if (k == 1) return eval1_for_bw<false, 1>(bw);
    if (k == 4) return eval1_for_bw<false, 4>(bw);
  }
  return nullptr;
}

}  // namespace fle
