// Building blocks of the single-pass kernels (fle_eval1.cu; their skew-capable twins in fle_eval1_skew.cu):
// plane loaders, the per-group plane algebra (PassRegs: compare chains, InMuxRegs: IN lists on narrow columns)
// and the tile loop around them.
#pragma once
#include <type_traits>

#include "fle_eval_common.cuh"

namespace fle {

struct SmemGroup {
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

struct GlobalGroup {
  const uint64_t* base;
  template <int BW>
  __device__ __forceinline__ void load_pair(int i, uint64_t& w0, uint64_t& w1) const {
    w0 = __ldg(base + i);
    w1 = (i + 1 < BW) ? __ldg(base + i + 1) : 0ull;
  }
};

// a group's planes already in registers
template <int BW>
struct RegGroup {
  uint64_t w[BW];
  template <int B>
  __device__ __forceinline__ void load_pair(int i, uint64_t& w0, uint64_t& w1) const {
    w0 = w[i];
    w1 = (i + 1 < BW) ? w[i + 1 < BW ? i + 1 : i] : 0ull;
  }
};

// the same loaders for a column at an arbitrary byte address (EvalParams::plane_skew)
struct GlobalGroupSkew {
  const uint8_t* base;
  template <int BW>
  __device__ __forceinline__ void load_pair(int i, uint64_t& w0, uint64_t& w1) const {
    ldg_pair_unaligned(base + i * 8, i + 1 < BW, w0, w1);
  }
};

template <bool IS_LT, int K, int BW>
struct PassRegs {
  static constexpr bool kComb = false;  // (compare kernels never fold a second bitmap in: keep their inner loop bare)
  static constexpr bool kRegMasks = (K * BW <= 64);
  uint32_t m[kRegMasks ? K * BW : 1];
  uint32_t lit[K], nm[K], ini[K];
  int mop;

  __device__ __forceinline__ void init(const PassDesc& pd) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      lit[k] = pd.lit[k];
      nm[k] = pd.neg[k] ? ~0u : 0u;
      ini[k] = IS_LT ? (pd.init[k] ? ~0u : 0u) : ~0u;
      if constexpr (kRegMasks) {
#pragma unroll
        for (int i = 0; i < BW; ++i) m[k * BW + i] = 0u - ((lit[k] >> i) & 1u);
      }
    }
    mop = pd.fold[K - 1];
  }

  // plane-order result word of one group (bit 63-r = row r)
  template <class L>
  __device__ __forceinline__ uint64_t run(const EvalParams&, const L& ld) const {
    return run(ld);
  }
  template <class L>
  __device__ __forceinline__ uint64_t run(const L& ld) const {
    uint32_t al[K], ah[K];
#pragma unroll
    for (int k = 0; k < K; ++k) al[k] = ah[k] = ini[k];
#pragma unroll
    for (int i = 0; i < BW; i += 2) {
      uint64_t w0, w1;
      ld.template load_pair<BW>(i, w0, w1);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (i + h < BW) {
          const uint64_t w = h ? w1 : w0;
          const uint32_t wl = (uint32_t)w, wh = (uint32_t)(w >> 32);
#pragma unroll
          for (int k = 0; k < K; ++k) {
            uint32_t mk;
            if constexpr (kRegMasks) mk = m[k * BW + i + h];
            else mk = 0u - ((lit[k] >> (i + h)) & 1u);
            if (IS_LT) {
              al[k] = (~wl & mk) | (~(wl ^ mk) & al[k]);
              ah[k] = (~wh & mk) | (~(wh ^ mk) & ah[k]);
            } else {
              al[k] = al[k] & ~(wl ^ mk);
              ah[k] = ah[k] & ~(wh ^ mk);
            }
          }
        }
      }
    }
    uint32_t rl = al[0] ^ nm[0], rh = ah[0] ^ nm[0];
#pragma unroll
    for (int k = 1; k < K; ++k) {
      const uint32_t xl = al[k] ^ nm[k], xh = ah[k] ^ nm[k];
      if (mop == FOLD_AND) { rl &= xl; rh &= xh; } else { rl |= xl; rh |= xh; }
    }
    return ((uint64_t)rh << 32) | rl;
  }
};

// Tiles are handed out dynamically: the producer thread claims the next tile from a global counter
// (one iteration before it is used, so the atomic's round trip to L2 is not waited for) and passes
// the tile id to the consumers through shared memory.  With a static split a CTA that starts late
// or shares its SM (an NCCL reduction of the previous query's count running beside the scan)
// delays the whole kernel by its full share: 8.5 us per 150 us step on two GPUs; with claimed tiles
// such a CTA only costs the tiles it did not get.  Claiming in batches of consecutive tiles was
// slower (bw 12: 1.04 -> 1.00 of roofline): neighbouring CTAs on neighbouring tiles is what the
// memory system likes.
// ------------------------------------------------------------------------------------------
// IN list on a narrow column, in the plane domain.  The reference ORs one equality chain per literal
// (fle-encoding.h:8279-8290: L * bw operations per 64 rows); the lookup kernels of fle_eval_in.cu
// transpose the planes into codes and look every row up (~9 instructions per ROW whatever the width).
// For bw <= 7 the membership function itself is small: a truth table of 2^bw bits.  Shannon expansion
// over the planes, per 32 rows:
//   minterms of the low T = min(bw, 3) planes            2^T LOP3 (immediate truth tables)
//   one leaf per value of the high planes:
//       leaf = OR_j (minterm_j & member[leaf * 2^T + j])   2^bw LOP3, the member masks are kernel
//                                                          parameters = constant-bank operands
//   multiplexer tree over the high planes                 2^(bw - T) - 1 LOP3 (x ? b : a)
// i.e. about 2^bw + 2^(bw-3) + 8 instructions per 32 rows, independent of the list length: bw 6 158 per
// group against 463 for the lookup kernel, bw 3 16.  Bit-identical to the OR of equalities: the same
// boolean function of the row's bits.
// LOP3 issues at 64 lanes per clock per SM, half the scheduler rate, and with every term on that pipe bw 6 / 7
// ran at 0.55 / 0.41 of the HBM roofline, ALU-bound.  The minterms of a leaf are DISJOINT bit sets, so their OR
// is also their SUM, and `minterm & member` with member in {0, ~0} is `minterm * member` with member in {0, 1}:
// every other term is therefore an IMAD on the FMA pipe (also 64 lanes per clock, independent of the ALU pipe;
// tools/micro/int_pipes.cu: LOP3 + IMAD together 122 lanes per clock).
// ------------------------------------------------------------------------------------------
template <int BW>
struct InMuxRegs {
  static_assert(BW >= 1 && BW <= 7, "plane-domain IN lists: bw <= 7");
  static constexpr bool kComb = true;   // EvalParams::comb_*: long list AND / OR the other operands' bitmap
  static constexpr int T = BW < 3 ? BW : 3;
  uint32_t nm;
  __device__ __forceinline__ void init(const PassDesc& pd) { nm = pd.neg[0] ? ~0u : 0u; }

  __device__ __forceinline__ static uint32_t half(const EvalParams& P, const uint32_t (&x)[BW]) {
    uint32_t m[1 << T];
    if constexpr (T == 1) {
      m[0] = ~x[0];
      m[1] = x[0];
    } else if constexpr (T == 2) {
      m[0] = ~x[0] & ~x[1];
      m[1] = x[0] & ~x[1];
      m[2] = ~x[0] & x[1];
      m[3] = x[0] & x[1];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t a = (j & 1) ? x[0] : ~x[0], b = (j & 2) ? x[1] : ~x[1], c = (j & 4) ? x[2] : ~x[2];
        m[j] = a & b & c;  // one LOP3: the complements fold into its truth table
      }
    }
    constexpr int kLeaves = 1 << (BW - T);
    uint32_t node[kLeaves];
#pragma unroll
    for (int k = 0; k < kLeaves; ++k) {
      uint32_t acc = m[0] & P.in_masks[k * (1 << T)];
#pragma unroll
      for (int j = 1; j < (1 << T); ++j) {
        if (in_mux_term_on_fma(BW, j)) acc = m[j] * P.in_masks[k * (1 << T) + j] + acc;  // IMAD: disjoint bits, 0 / 1 factor
        else acc |= m[j] & P.in_masks[k * (1 << T) + j];
      }
      node[k] = acc;
    }
#pragma unroll
    for (int i = T; i < BW; ++i) {
#pragma unroll
      for (int k = 0; k < (kLeaves >> (i - T + 1)); ++k) node[k] = (node[2 * k] & ~x[i]) | (node[2 * k + 1] & x[i]);
    }
    return node[0];
  }

  template <class L>
  __device__ __forceinline__ uint64_t run(const EvalParams& P, const L& ld) const {
    uint32_t xl[BW], xh[BW];
#pragma unroll
    for (int i = 0; i < BW; i += 2) {
      uint64_t w0, w1;
      ld.template load_pair<BW>(i, w0, w1);
      xl[i] = (uint32_t)w0;
      xh[i] = (uint32_t)(w0 >> 32);
      if (i + 1 < BW) {
        xl[i + 1 < BW ? i + 1 : i] = (uint32_t)w1;
        xh[i + 1 < BW ? i + 1 : i] = (uint32_t)(w1 >> 32);
      }
    }
    const uint32_t rl = half(P, xl) ^ nm, rh = half(P, xh) ^ nm;
    return ((uint64_t)rh << 32) | rl;
  }
};

template <class PR, int BW, bool SKEW = false>
__device__ __forceinline__ void eval1_body(const EvalParams& P) {
  constexpr int NT = 256;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ int64_t stage_tile[8];  // tile held by each stage, -1 = no more tiles
  const uint32_t bar0 = smem_u32(smem + P.smem_bar_off);
  const uint32_t stage0 = smem_u32(smem + P.smem_stage_off);
  const int gpt = P.gpt;
  const int T = NT * gpt;
  const int seg_groups = T >> 3;
  const int64_t ntiles = (SKEW ? P.staged_groups : P.ngroups) / T;
  const int nst = P.nstages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) mbar_init(bar0 + s * 8, 1);
    mbar_fence_init();
  }
  __syncthreads();
  // (the claim for the NEXT iteration is issued one iteration early, so the round trip of the
  //  atomic to L2 -- about one tile's worth of time -- is never waited for)
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
  if (threadIdx.x == 0) {
    for (int s = 0; s < nst - 1; ++s) arm(s, atomicAdd(P.tile_counter, 1u));
    next_claim = atomicAdd(P.tile_counter, 1u);
  }

  PR pr;
  pr.init(P.passes[0]);

  const uint32_t seg = threadIdx.x & 7, jrow = threadIdx.x >> 3;
  // shared address of the thread's first group inside a stage, and the step to its next one
  const uint32_t thread_off = P.col_off[0] + seg * P.seg_stride[0] + jrow * (uint32_t)(BW * 8) + (SKEW ? P.plane_skew[0] : 0u);
  constexpr uint32_t kGroupStride = (uint32_t)(NT / 8) * BW * 8u;
  // bitmap word of the thread's first group inside a tile
  const int64_t thread_g = (int64_t)seg * seg_groups + jrow;
  uint64_t cnt = 0;

  int stage = 0;
  uint32_t parity = 0;
  while (true) {
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
    uint32_t a = stage0 + stage * P.stage_bytes + thread_off;
    int64_t g = tile * T + thread_g;
    // two groups interleaved (their dependency chains are short of bandwidth-bound on their own:
    // bw 4 0.91 -> 0.95, bw 12 1.02 -> 1.04)
#pragma unroll 2
    for (int q = 0; q < gpt; ++q) {
      uint64_t other = 0;
      if constexpr (PR::kComb) other = comb_load(P, g);
      uint64_t r;
      if constexpr (SKEW) {
        RegGroup<BW> rg;
        lds_planes_unaligned<BW>(a, rg.w);
        r = finish_word(pr.run(P, rg), g, P.nrows);
      } else {
        r = finish_word(pr.run(P, SmemGroup{a}), g, P.nrows);
      }
      if constexpr (PR::kComb) r = comb_apply(P, r, other);
      P.bitmap[g] = r;
      cnt += __popcll(r);
      a += kGroupStride;
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
    uint64_t other = 0;
    if constexpr (PR::kComb) other = comb_load(P, g);
    uint64_t r;
    if constexpr (SKEW)
      r = finish_word(pr.run(P, GlobalGroupSkew{reinterpret_cast<const uint8_t*>(P.planes[0]) + P.plane_skew[0] + g * (BW * 8)}), g, P.nrows);
    else
      r = finish_word(pr.run(P, GlobalGroup{P.planes[0] + g * BW}), g, P.nrows);
    if constexpr (PR::kComb) r = comb_apply(P, r, other);
    P.bitmap[g] = r;
    cnt += __popcll(r);
  }
  publish_count<NT>(P, cnt);
}

}  // namespace fle
