// IN-list evaluation by decode + membership lookup, for dictionary-coded columns (bw <= 16).
//
// The reference evaluates `x IN (c1..cL)` as an OR over L equality chains and re-reads every
// plane once per literal (fle-encoding.h:8279-8290): O(L * bw) boolean ops per 64 rows.  That is
// what the generic kernels do too (4 literals per plane sweep) and it is ALU-bound beyond a
// dozen literals (profiles/: L = 256 at 2 % of the HBM roofline).  For bw <= 16 -- every column
// the scanner can reach, since a dictionary holds at most 40000 entries (dict-encoding.h:157) --
// this kernel instead
//   1. builds a 2^bw-bit membership table of the literals in shared memory (<= 8 KB); for
//      bw <= 13 the table is replicated per lane (word w of lane l at w * 128 + l * 4 bytes, <= 32 KB)
//      so that the 64 data-dependent lookups of a group never conflict on a bank -- with one shared
//      copy 32 random words hit ~3.5 words per bank and the LSU pipe, not HBM, bounded the kernel,
//   2. transposes each group's planes into its 64 codes in registers (one 32x32 butterfly, both
//      32-row halves at once, as in the unpack kernel),
//   3. looks every code up in the table and assembles the 64-bit selection word directly in
//      row order,
// which costs the same for any L.  Same TMA/mbarrier tile pipeline and skewed shared layout as
// fle_eval1.cu.  Results are bit-identical to the OR-of-equalities by construction (set
// membership of the same bw-bit codes).
#include <algorithm>

#include "fle_eval_common.cuh"

namespace fle {

// lane-private table copies (bank = lane) while they fit: 2^13 bits x 32 lanes = 32 KB
template <int BW>
struct InTable {
  static constexpr bool kPerLane = BW <= 13;
  static constexpr int kWords = BW <= 5 ? 1 : (1 << (BW - 5));
  static constexpr int kShift = kPerLane ? 7 : 2;  // byte stride between consecutive table words
};

template <int BW, class L>
__device__ __forceinline__ uint64_t in_lookup_group(const L& ld, uint32_t table) {
  uint64_t w[BW];
#pragma unroll
  for (int i = 0; i < BW; i += 2) {
    uint64_t a, b;
    ld.template load_pair<BW>(i, a, b);
    w[i] = a;
    if (i + 1 < BW) w[i + 1] = b;
  }
  // A'[k] = code(row k) | code(row 32 + k) << 16   (see fle_device.cuh / fle_codec.cu)
  uint32_t A[32];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    A[31 - i] = i < BW ? (uint32_t)(w[i < BW ? i : 0] >> 32) : 0u;
    A[15 - i] = i < BW ? (uint32_t)(w[i < BW ? i : 0]) : 0u;
  }
  transpose32(A);
  uint32_t lo = 0, hi = 0;  // rows 0..31 / 32..63, row order (bit r%32)
#pragma unroll
  for (int k = 31; k >= 0; --k) {
    const uint32_t c0 = A[k] & 0xFFFFu, c1 = A[k] >> 16;
    const uint32_t t0 = lds32(table + ((c0 >> 5) << InTable<BW>::kShift));
    const uint32_t t1 = lds32(table + ((c1 >> 5) << InTable<BW>::kShift));
    lo = (lo << 1) | (__funnelshift_r(t0, 0u, c0) & 1u);  // funnel shift uses c & 31
    hi = (hi << 1) | (__funnelshift_r(t1, 0u, c1) & 1u);
  }
  return ((uint64_t)hi << 32) | lo;
}

struct InSmemGroup {
  uint32_t base;
  template <int BW>
  __device__ __forceinline__ void load_pair(int i, uint64_t& w0, uint64_t& w1) const {
    if constexpr (BW % 2 == 0) {
      lds128(base + i * 8, w0, w1);
    } else {
      w0 = lds64(base + i * 8);
      w1 = (i + 1 < BW) ? lds64(base + (i + 1) * 8) : 0ull;
    }
  }
};

struct InGlobalGroup {
  const uint64_t* base;
  template <int BW>
  __device__ __forceinline__ void load_pair(int i, uint64_t& w0, uint64_t& w1) const {
    w0 = __ldg(base + i);
    w1 = (i + 1 < BW) ? __ldg(base + i + 1) : 0ull;
  }
};

__device__ __forceinline__ uint64_t finish_row_order(uint64_t r, uint64_t negmask, int64_t g, int64_t nrows) {
  r ^= negmask;
  const int64_t rows_left = nrows - g * 64;
  if (rows_left < 64) r &= (rows_left <= 0) ? 0ull : ((1ull << rows_left) - 1ull);
  return r;
}

template <int BW>
__global__ void __launch_bounds__(256) eval_in_kernel(const __grid_constant__ EvalParams P) {
  constexpr int NT = 256;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t bar0 = smem_u32(smem + P.smem_bar_off);
  const uint32_t stage0 = smem_u32(smem + P.smem_stage_off);
  // [lane-private copies: kWords x 32 words (bw <= 13 only)][one shared copy: kWords words]
  using IT = InTable<BW>;
  uint32_t* rep_ptr = reinterpret_cast<uint32_t*>(smem + P.smem_masks_off);
  uint32_t* table_ptr = rep_ptr + (IT::kPerLane ? IT::kWords * 32 : 0);
  const uint32_t table = IT::kPerLane ? smem_u32(rep_ptr) + (threadIdx.x & 31u) * 4u : smem_u32(table_ptr);
  const int gpt = P.gpt;
  const int T = NT * gpt;
  const int seg_groups = T >> 3;
  const int64_t ntiles = P.ngroups / T;
  const int nst = P.nstages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) mbar_init(bar0 + s * 8, 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < IT::kWords; i += NT) table_ptr[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int s = 0; s < nst - 1; ++s) {
      const int64_t tile = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
      if (tile < ntiles) issue_tile(P, bar0, stage0, s, tile, T);
    }
  }
  const PassDesc& pd = P.passes[0];
  const int off = (int)pd.lit[0], cnt = (int)pd.lit[1];
  for (int i = threadIdx.x; i < cnt; i += NT) {
    const uint32_t c = P.in_lits[off + i];  // already masked to BW bits by the host
    atomicOr(&table_ptr[c >> 5], 1u << (c & 31));
  }
  __syncthreads();
  if (IT::kPerLane) {
    for (int i = threadIdx.x; i < IT::kWords * 32; i += NT) rep_ptr[i] = table_ptr[i >> 5];
    __syncthreads();
  }
  const uint64_t negmask = pd.neg[0] ? ~0ull : 0ull;

  const uint32_t seg = threadIdx.x & 7, jrow = threadIdx.x >> 3;
  const uint32_t thread_off = P.col_off[0] + seg * P.seg_stride[0] + jrow * (uint32_t)(BW * 8);
  constexpr uint32_t kGroupStride = (uint32_t)(NT / 8) * BW * 8u;
  const int64_t thread_g = (int64_t)seg * seg_groups + jrow;
  uint64_t cnt_bits = 0;

  int stage = 0;
  uint32_t parity = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const int64_t nxt = tile + (int64_t)(nst - 1) * gridDim.x;
      int st2 = stage + nst - 1;
      if (st2 >= nst) st2 -= nst;
      if (nxt < ntiles) issue_tile(P, bar0, stage0, st2, nxt, T);
    }
    mbar_wait(bar0 + stage * 8, parity);
    uint32_t a = stage0 + stage * P.stage_bytes + thread_off;
    int64_t g = tile * T + thread_g;
    for (int q = 0; q < gpt; ++q) {
      const uint64_t r = finish_row_order(in_lookup_group<BW>(InSmemGroup{a}, table), negmask, g, P.nrows);
      P.bitmap[g] = r;
      cnt_bits += __popcll(r);
      a += kGroupStride;
      g += NT / 8;
    }
    if (++stage == nst) {
      stage = 0;
      parity ^= 1u;
    }
  }
  for (int64_t g = ntiles * T + (int64_t)blockIdx.x * NT + threadIdx.x; g < P.ngroups;
       g += (int64_t)gridDim.x * NT) {
    const uint64_t r = finish_row_order(in_lookup_group<BW>(InGlobalGroup{P.planes[0] + g * BW}, table), negmask, g, P.nrows);
    P.bitmap[g] = r;
    cnt_bits += __popcll(r);
  }
  publish_count<NT>(P, cnt_bits);
}

const void* eval_in_kernel_for(int bw) {
  switch (bw) {
#define X(B) case B: return (const void*)eval_in_kernel<B>;
    X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)
#undef X
  }
  return nullptr;
}

}  // namespace fle
