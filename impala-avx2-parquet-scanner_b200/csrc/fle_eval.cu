// Fused encoded-domain predicate evaluation on FLE (bit-sliced) column chunks.
//
// Replaces, in one pass over HBM:
//   FleDecoder::Eq/Lt/Le/Gt/Ge/In     /root/reference/fle-encoding.h:7962-8313
//   And/OrOperate::GetBitset          /root/reference/simple-predicates.h:145-163
//   conjunct `&=` loop                /root/reference/hdfs-parquet-scanner.cc:1857-1862
//
// Work unit: one 64-row group = sum(bw_c) plane words in, one bitmap word out.
// One thread owns one group at a time and walks its planes LSB -> MSB with 64-bit boolean
// algebra (2 x LOP3 per plane per comparison), so no value is ever decoded.
//
// eval_tma_kernel (default for programs with several passes): persistent CTAs; every tile of
// T = 256*gpt groups is staged global -> shared by TMA bulk copies (cp.async.bulk, one per column
// per segment) into an nstages-deep mbarrier ring, so ~50-150 KB per SM are in flight with no
// register cost.  The tile of each column is split into 8 segments skewed by 16 B so that the
// per-thread (strided by bw*8 bytes) 64/128-bit shared loads are bank-conflict free for every
// bit width.
// eval_direct_kernel: same algebra, planes read straight from global memory; used for
// unaligned inputs, tiles that do not fit in shared memory, the ragged last tile's groups, and
// as an on-device cross-check (eval_variant = 0).
// Single-pass programs and single IN lists have their own kernels (fle_eval1.cu, fle_eval_in.cu).
#include <algorithm>
#include <cstring>
#include <functional>
#include <mutex>
#include <set>
#include <utility>
#include <vector>

#include "fle_device.cuh"
#include "fle_internal.h"
#include "fle_eval_common.cuh"

namespace fle {

// ------------------------------------------------------------------------------------------
// plane loaders
// ------------------------------------------------------------------------------------------
struct SmemPlanes {
  uint32_t base;  // shared address of plane 0 of this thread's group
  __device__ __forceinline__ uint64_t load1(int i) const { return lds64(base + i * 8); }
  __device__ __forceinline__ void load2(int i, uint64_t& a, uint64_t& b) const {
    lds128(base + i * 8, a, b);
  }
};

struct GlobalPlanes {
  const uint64_t* base;
  __device__ __forceinline__ uint64_t load1(int i) const { return __ldg(base + i); }
  __device__ __forceinline__ void load2(int i, uint64_t& a, uint64_t& b) const {
    a = __ldg(base + i);
    b = __ldg(base + i + 1);
  }
};

// ------------------------------------------------------------------------------------------
// chain steps
// ------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void lt_step(uint64_t w, int i, const uint32_t (&lit)[4],
                                        uint64_t (&acc)[K]) {
  const uint32_t wl = (uint32_t)w, wh = (uint32_t)(w >> 32);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const uint32_t m = 0u - ((lit[k] >> i) & 1u);  // warp-uniform
    uint32_t al = (uint32_t)acc[k], ah = (uint32_t)(acc[k] >> 32);
    al = (~wl & m) | (~(wl ^ m) & al);
    ah = (~wh & m) | (~(wh ^ m) & ah);
    acc[k] = ((uint64_t)ah << 32) | al;
  }
}

template <int K>
__device__ __forceinline__ void eq_step(uint64_t w, int i, const uint32_t (&lit)[4],
                                        uint64_t (&acc)[K]) {
  const uint32_t wl = (uint32_t)w, wh = (uint32_t)(w >> 32);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const uint32_t m = 0u - ((lit[k] >> i) & 1u);
    uint32_t al = (uint32_t)acc[k], ah = (uint32_t)(acc[k] >> 32);
    al = al & ~(wl ^ m);
    ah = ah & ~(wh ^ m);
    acc[k] = ((uint64_t)ah << 32) | al;
  }
}

template <bool IS_LT, int K, class L>
__device__ __forceinline__ void run_chains(const L& ld, int bw, const uint32_t (&lit)[4],
                                           uint64_t (&acc)[K]) {
  if (bw & 1) {
#pragma unroll 4
    for (int i = 0; i < bw; ++i) {
      const uint64_t w = ld.load1(i);
      if (IS_LT) lt_step<K>(w, i, lit, acc); else eq_step<K>(w, i, lit, acc);
    }
  } else {
#pragma unroll 2
    for (int i = 0; i < bw; i += 2) {
      uint64_t w0, w1;
      ld.load2(i, w0, w1);
      if (IS_LT) {
        lt_step<K>(w0, i, lit, acc);
        lt_step<K>(w1, i + 1, lit, acc);
      } else {
        eq_step<K>(w0, i, lit, acc);
        eq_step<K>(w1, i + 1, lit, acc);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// slot machine: slot 0 in a register, slots >= 1 in shared memory (one word per thread)
// ------------------------------------------------------------------------------------------
struct Slots {
  uint64_t s0;
  uint32_t smem;    // shared address of this thread's word of slot 1
  uint32_t stride;  // bytes between slots
  __device__ __forceinline__ uint64_t get(int s) const {
    return s == 0 ? s0 : lds64(smem + (s - 1) * stride);
  }
  __device__ __forceinline__ void put(int s, uint64_t v) {
    if (s == 0) s0 = v; else sts64(smem + (s - 1) * stride, v);
  }
  __device__ __forceinline__ void fold(int s, int op, uint64_t v) {
    if (op == FOLD_SET) {
      put(s, v);
    } else {
      const uint64_t cur = get(s);
      put(s, op == FOLD_AND ? (cur & v) : (cur | v));
    }
  }
};

template <bool IS_LT, int K, class L>
__device__ __forceinline__ void do_pass(const PassDesc& pd, const L& ld, int bw, Slots& sl) {
  uint64_t acc[K];
  uint32_t lit[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) lit[k] = pd.lit[k];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = IS_LT ? (pd.init[k] ? ~0ull : 0ull) : ~0ull;
  run_chains<IS_LT, K, L>(ld, bw, lit, acc);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const uint64_t r = pd.neg[k] ? ~acc[k] : acc[k];
    sl.fold(pd.slot[k], pd.fold[k], r);
  }
}

template <class L>
__device__ __forceinline__ void do_pass_in(const PassDesc& pd, const L& ld, int bw,
                                           const uint32_t* lits_smem, Slots& sl) {
  const int off = (int)pd.lit[0], cnt = (int)pd.lit[1];
  uint64_t any = 0;
  for (int j = 0; j < cnt; j += 4) {
    uint32_t lit[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) lit[k] = lits_smem[off + min(j + k, cnt - 1)];  // pad by repeat
    uint64_t acc[4] = {~0ull, ~0ull, ~0ull, ~0ull};
    run_chains<false, 4, L>(ld, bw, lit, acc);
    any |= (acc[0] | acc[1]) | (acc[2] | acc[3]);
  }
  sl.fold(pd.slot[0], pd.fold[0], pd.neg[0] ? ~any : any);
}

// Evaluate the whole lowered program for one group.  LoaderOf(col) yields the plane loader
// of this thread's group in column col.  Returns the plane-order result word (bit 63-k = row k).
template <class L, class LoaderOf>
__device__ __forceinline__ uint64_t eval_group(const EvalParams& P, const uint32_t* lits_smem,
                                               Slots& sl, LoaderOf loader_of) {
  sl.s0 = P.slot0_init;
  for (int p = 0; p < P.npasses; ++p) {
    const PassDesc& pd = P.passes[p];
    if (pd.type == PASS_CONST) {
      sl.fold(pd.slot[0], pd.fold[0], pd.init[0] ? ~0ull : 0ull);
      continue;
    }
    const int bw = (int)P.bw[pd.col];
    const L ld = loader_of(pd.col);
    if (pd.type == PASS_LT) {
      switch (pd.k) {
        case 1: do_pass<true, 1, L>(pd, ld, bw, sl); break;
        case 2: do_pass<true, 2, L>(pd, ld, bw, sl); break;
        case 3: do_pass<true, 3, L>(pd, ld, bw, sl); break;
        default: do_pass<true, 4, L>(pd, ld, bw, sl); break;
      }
    } else if (pd.type == PASS_EQ) {
      switch (pd.k) {
        case 1: do_pass<false, 1, L>(pd, ld, bw, sl); break;
        case 2: do_pass<false, 2, L>(pd, ld, bw, sl); break;
        case 3: do_pass<false, 3, L>(pd, ld, bw, sl); break;
        default: do_pass<false, 4, L>(pd, ld, bw, sl); break;
      }
    } else {
      do_pass_in<L>(pd, ld, bw, lits_smem, sl);
    }
  }
  for (int c = 0; c < P.ncombs; ++c) {
    const CombDesc cd = P.combs[c];
    const uint64_t a = sl.get(cd.a), b = sl.get(cd.b);
    sl.put(cd.dst, cd.op == FOLD_AND ? (a & b) : (a | b));
  }
  return sl.s0;
}

// ------------------------------------------------------------------------------------------
// Kernel 0: direct global loads
// ------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT) eval_direct_kernel(const __grid_constant__ EvalParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint32_t* lits_smem = reinterpret_cast<uint32_t*>(smem + P.smem_lits_off);
  load_in_lits(P, lits_smem);
  __syncthreads();
  Slots sl;
  sl.smem = smem_u32(smem + P.smem_slots_off) + threadIdx.x * 8;
  sl.stride = NT * 8;
  uint64_t cnt = 0;
  for (int64_t g = (int64_t)blockIdx.x * NT + threadIdx.x; g < P.ngroups;
       g += (int64_t)gridDim.x * NT) {
    auto loader_of = [&](int col) {
      return GlobalPlanes{P.planes[col] + g * (int64_t)P.bw[col]};
    };
    const uint64_t r = finish_word(eval_group<GlobalPlanes>(P, lits_smem, sl, loader_of), g, P.nrows);
    P.bitmap[g] = r;
    cnt += __popcll(r);
  }
  publish_count<NT>(P, cnt);
}

// ------------------------------------------------------------------------------------------
// Kernel 1: TMA-staged persistent pipeline for generic programs
//
// Tiles of T = 256 * gpt groups are staged by TMA exactly as in fle_eval1.cu.  A thread then
// evaluates its group pass by pass with small bodies templated on <kind, K chains> only: the
// bit width is a run-time loop over plane QUADS -- two LDS.128 fetch four planes, one LDS.128
// per chain fetches the four literal masks (precomputed once per CTA into shared memory),
// 8*K LOP3 do the work -- plus a remainder loop for bw % 4 planes.  Slot 0 (the only slot of
// flat AND / OR programs) lives in a register, further slots of mixed trees in per-thread shared
// words.  Keeping the bodies small (no per-bit-width unrolling here; the single-pass kernels of
// fle_eval1.cu have that) keeps the kernel's register allocation and instruction cache tight.
// ------------------------------------------------------------------------------------------
template <bool IS_LT, int K>
__device__ __forceinline__ void quad_step(const uint64_t (&w)[4], const uint32_t (&m)[K][4], uint32_t (&al)[K],
                                          uint32_t (&ah)[K]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t wl = (uint32_t)w[j], wh = (uint32_t)(w[j] >> 32);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const uint32_t mk = m[k][j];
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

// One pass on one group: planes at shared address `a` (8-byte aligned; 16-byte aligned when bw is
// even), masks at mask_addr ([4 chains][32 planes] u32).  r[k] = plane-order result of chain k.
template <bool IS_LT, int K>
__device__ __forceinline__ void staged_pass(uint32_t a, int bw, uint32_t mask_addr, uint32_t inits4, uint32_t negs4,
                                            uint64_t (&r)[K]) {
  uint32_t al[K], ah[K];
#pragma unroll
  for (int k = 0; k < K; ++k) al[k] = ah[k] = IS_LT ? (0u - ((inits4 >> (8 * k)) & 1u)) : ~0u;
  int i = 0;
  const bool even = (bw & 1) == 0;
  for (; i + 4 <= bw; i += 4) {
    uint64_t w[4];
    if (even) {
      lds128(a + i * 8, w[0], w[1]);
      lds128(a + i * 8 + 16, w[2], w[3]);
    } else {
      w[0] = lds64(a + i * 8);
      w[1] = lds64(a + i * 8 + 8);
      w[2] = lds64(a + i * 8 + 16);
      w[3] = lds64(a + i * 8 + 24);
    }
    uint32_t m[K][4];
#pragma unroll
    for (int k = 0; k < K; ++k) lds128u(mask_addr + k * 128 + i * 4, m[k][0], m[k][1], m[k][2], m[k][3]);
    quad_step<IS_LT, K>(w, m, al, ah);
  }
  for (; i < bw; ++i) {
    const uint64_t w = lds64(a + i * 8);
    const uint32_t wl = (uint32_t)w, wh = (uint32_t)(w >> 32);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const uint32_t mk = lds32(mask_addr + k * 128 + i * 4);
      if (IS_LT) {
        al[k] = (~wl & mk) | (~(wl ^ mk) & al[k]);
        ah[k] = (~wh & mk) | (~(wh ^ mk) & ah[k]);
      } else {
        al[k] = al[k] & ~(wl ^ mk);
        ah[k] = ah[k] & ~(wh ^ mk);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const uint32_t nm = 0u - ((negs4 >> (8 * k)) & 1u);
    r[k] = ((uint64_t)(ah[k] ^ nm) << 32) | (al[k] ^ nm);
  }
}

// fold K chain results into their slots (slot 0 = register s0, others = per-thread shared words)
template <int K>
__device__ __forceinline__ void staged_fold(const uint64_t (&r)[K], uint32_t slots4, uint32_t folds4, uint32_t merge,
                                            Slots& sl) {
  if (K > 1 && merge) {
    const int mop = (int)((folds4 >> (8 * (K - 1))) & 0xffu);
    uint64_t v = r[0];
#pragma unroll
    for (int k = 1; k < K; ++k) v = (mop == FOLD_AND) ? (v & r[k]) : (v | r[k]);
    sl.fold((int)(slots4 & 0xffu), (int)(folds4 & 0xffu), v);
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) sl.fold((int)((slots4 >> (8 * k)) & 0xffu), (int)((folds4 >> (8 * k)) & 0xffu), r[k]);
  }
}

template <int NT>
__global__ void __launch_bounds__(NT, 768 / NT) eval_tma_kernel(const __grid_constant__ EvalParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint32_t* lits_smem = reinterpret_cast<uint32_t*>(smem + P.smem_lits_off);
  uint32_t* masks_smem = reinterpret_cast<uint32_t*>(smem + P.smem_masks_off);
  const uint32_t bar0 = smem_u32(smem + P.smem_bar_off);
  const uint32_t stage0 = smem_u32(smem + P.smem_stage_off);
  const uint32_t masks0 = smem_u32(masks_smem);
  const int gpt = P.gpt;
  const int T = NT * gpt;            // groups per tile
  const int seg_groups = T >> 3;     // groups per skewed segment
  const int64_t ntiles = P.ngroups / T;  // full tiles; the ragged rest goes the direct way
  const int nst = P.nstages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) mbar_init(bar0 + s * 8, 1);
    mbar_fence_init();
  }
  load_in_lits(P, lits_smem);
  // literal masks: entry [p][k][i] = all-ones iff bit i of chain k's literal of pass p is set
  for (int e = threadIdx.x; e < P.npasses * 128; e += NT) {
    const int p = e >> 7, k = (e >> 5) & 3, i = e & 31;
    masks_smem[e] = 0u - ((P.passes[p].lit[k] >> i) & 1u);
  }
  __syncthreads();

  // producer side (thread 0): stage `stage` <- tile `tile`
  auto issue = [&](int stage, int64_t tile) { issue_tile(P, bar0, stage0, stage, tile, T); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < nst - 1; ++s) {
      const int64_t tile = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
      if (tile < ntiles) issue(s, tile);
    }
  }

  const uint32_t seg = threadIdx.x & 7, jrow = threadIdx.x >> 3;
  Slots sl;
  sl.smem = smem_u32(smem + P.smem_slots_off) + threadIdx.x * 8u;
  sl.stride = NT * 8;
  uint64_t cnt = 0;

  int stage = 0;
  uint32_t parity = 0;
  for (int64_t n = 0;; ++n) {
    const int64_t tile = (int64_t)blockIdx.x + n * gridDim.x;
    if (tile >= ntiles) break;
    // every thread is done reading the stage consumed in iteration n-1 -> safe to refill it
    __syncthreads();
    if (threadIdx.x == 0) {
      const int64_t nxt = tile + (int64_t)(nst - 1) * gridDim.x;
      int st2 = stage + nst - 1;
      if (st2 >= nst) st2 -= nst;
      if (nxt < ntiles) issue(st2, nxt);
    }
    mbar_wait(bar0 + stage * 8, parity);
    const uint32_t sbase = stage0 + stage * P.stage_bytes;

    for (int q = 0; q < gpt; ++q) {
      const uint32_t jq = (uint32_t)q * (NT / 8) + jrow;  // group index inside the thread's segment
      sl.s0 = P.slot0_init;
      // passes: warp-uniform control flow; the pass record is read as 32-bit constant-bank words
      for (int p = 0; p < P.npasses; ++p) {
        const uint32_t* pw = reinterpret_cast<const uint32_t*>(&P.passes[p]);
        const uint32_t w0 = pw[0], slots4 = pw[1], folds4 = pw[2], negs4 = pw[3], inits4 = pw[4];
        const uint32_t type = w0 & 0xffu, col = (w0 >> 8) & 0xffu, k = (w0 >> 16) & 0xffu, merge = w0 >> 24;
        if (type == PASS_CONST) {
          sl.fold((int)(slots4 & 0xffu), (int)(folds4 & 0xffu), (inits4 & 1u) ? ~0ull : 0ull);
          continue;
        }
        const int bw = (int)P.bw[col];
        const uint32_t a = sbase + P.col_off[col] + seg * P.seg_stride[col] + jq * (uint32_t)(bw * 8);
        const uint32_t mask_addr = masks0 + (uint32_t)p * 512u;
        if (type == PASS_LT) {
          if (k == 1) { uint64_t r[1]; staged_pass<true, 1>(a, bw, mask_addr, inits4, negs4, r); staged_fold<1>(r, slots4, folds4, merge, sl); }
          else if (k == 2) { uint64_t r[2]; staged_pass<true, 2>(a, bw, mask_addr, inits4, negs4, r); staged_fold<2>(r, slots4, folds4, merge, sl); }
          else { uint64_t r[4]; staged_pass<true, 4>(a, bw, mask_addr, inits4, negs4, r); staged_fold<4>(r, slots4, folds4, merge, sl); }
        } else if (type == PASS_EQ) {
          if (k == 1) { uint64_t r[1]; staged_pass<false, 1>(a, bw, mask_addr, inits4, negs4, r); staged_fold<1>(r, slots4, folds4, merge, sl); }
          else { uint64_t r[4]; staged_pass<false, 4>(a, bw, mask_addr, inits4, negs4, r); staged_fold<4>(r, slots4, folds4, merge, sl); }
        } else {  // PASS_IN: OR of equality chains, four literals per sweep (fle-encoding.h:8282-8290)
          const uint32_t* lits = lits_smem + pw[5];
          const int lcnt = (int)pw[6];
          uint64_t any = 0;
          for (int j = 0; j < lcnt; j += 4) {
            uint32_t lit[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) lit[kk] = lits[min(j + kk, lcnt - 1)];
            uint32_t al[4], ah[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) al[kk] = ah[kk] = ~0u;
            for (int i = 0; i < bw; ++i) {
              const uint64_t w = lds64(a + i * 8);
              const uint32_t wl = (uint32_t)w, wh = (uint32_t)(w >> 32);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint32_t mk = 0u - ((lit[kk] >> i) & 1u);
                al[kk] = al[kk] & ~(wl ^ mk);
                ah[kk] = ah[kk] & ~(wh ^ mk);
              }
            }
            any |= ((uint64_t)((ah[0] | ah[1]) | (ah[2] | ah[3])) << 32) | ((al[0] | al[1]) | (al[2] | al[3]));
          }
          sl.fold((int)(slots4 & 0xffu), (int)(folds4 & 0xffu), (negs4 & 1u) ? ~any : any);
        }
      }
      for (int c = 0; c < P.ncombs; ++c) {
        const CombDesc cd = P.combs[c];
        const uint64_t x = sl.get(cd.a), y = sl.get(cd.b);
        sl.put(cd.dst, cd.op == FOLD_AND ? (x & y) : (x | y));
      }
      const int64_t g = tile * T + (int64_t)seg * seg_groups + jq;
      const uint64_t r = finish_word(sl.s0, g, P.nrows);
      P.bitmap[g] = r;
      cnt += __popcll(r);
    }
    if (++stage == nst) {
      stage = 0;
      parity ^= 1u;
    }
  }

  // ragged remainder (< T groups): spread over all CTAs, planes straight from global memory
  {
    const int64_t g0 = ntiles * T;
    for (int64_t g = g0 + (int64_t)blockIdx.x * NT + threadIdx.x; g < P.ngroups;
         g += (int64_t)gridDim.x * NT) {
      auto loader_of = [&](int col) {
        return GlobalPlanes{P.planes[col] + g * (int64_t)P.bw[col]};
      };
      const uint64_t r = finish_word(eval_group<GlobalPlanes>(P, lits_smem, sl, loader_of), g, P.nrows);
      P.bitmap[g] = r;
      cnt += __popcll(r);
    }
  }
  publish_count<NT>(P, cnt);
}

// ------------------------------------------------------------------------------------------
// host: lowering
// ------------------------------------------------------------------------------------------
namespace {
struct Chain {
  uint8_t type, col, slot, fold, neg, init;
  uint32_t lit;
  uint32_t in_off, in_cnt;
  int order;
};
}  // namespace

int lower_program(const fle_col_desc* cols, int ncols, const fle_pred_node* nodes, int nnodes,
                  LoweredProgram* out, std::string* err, bool mask_literals) {
  auto fail = [&](int code, const char* m) {
    if (err) *err = m;
    return code;
  };
  if (!cols || !nodes || !out) return fail(FLE_E_INVALID, "null argument");
  if (ncols < 1 || nnodes < 1) return fail(FLE_E_INVALID, "empty program");
  if (ncols > FLE_MAX_COLS) return fail(FLE_E_UNSUPPORTED, "too many columns");
  if (nnodes > FLE_MAX_NODES) return fail(FLE_E_UNSUPPORTED, "too many predicate nodes");
  for (int c = 0; c < ncols; ++c)
    if (cols[c].bit_width < 1 || cols[c].bit_width > 32)
      return fail(FLE_E_INVALID, "bit width outside 1..32");

  // validate the postfix program and find out whether it is one flat AND / OR
  int depth = 0, nleaves = 0;
  bool has_and = false, has_or = false;
  for (int i = 0; i < nnodes; ++i) {
    const fle_pred_node& nd = nodes[i];
    if (nd.op >= FLE_OP_EQ && nd.op <= FLE_OP_IN) {
      if (nd.col < 0 || nd.col >= ncols) return fail(FLE_E_INVALID, "leaf column out of range");
      if (nd.op == FLE_OP_IN) {
        if (nd.ncodes < 0 || nd.ncodes > FLE_MAX_IN_CODES || (nd.ncodes > 0 && !nd.codes))
          return fail(FLE_E_INVALID, "bad IN list");
      }
      ++depth;
      ++nleaves;
    } else if (nd.op == FLE_OP_ALL || nd.op == FLE_OP_NONE) {
      ++depth;
      ++nleaves;
    } else if (nd.op == FLE_OP_AND || nd.op == FLE_OP_OR) {
      if (depth < 2) return fail(FLE_E_INVALID, "malformed postfix program");
      --depth;
      (nd.op == FLE_OP_AND ? has_and : has_or) = true;
    } else {
      return fail(FLE_E_INVALID, "unknown opcode");
    }
  }
  if (depth != 1) return fail(FLE_E_INVALID, "malformed postfix program");
  const bool flat = !(has_and && has_or);
  const uint8_t flat_fold = has_or ? FOLD_OR : FOLD_AND;

  LoweredProgram& lp = *out;
  std::memset(&lp.p, 0, sizeof(lp.p));
  lp.in_lits.clear();
  EvalParams& P = lp.p;
  P.slot0_init = flat ? (flat_fold == FOLD_AND ? ~0ull : 0ull) : 0ull;

  std::vector<Chain> chains;
  // General mode: every leaf owns a slot for the whole evaluation, because ALL passes run
  // before ANY combine op (passes are regrouped by column, not kept in postfix order).
  std::vector<int> stack;       // slot ids (general mode)
  int next_slot = 0;
  auto alloc_slot = [&]() { return next_slot < kMaxSlots ? next_slot++ : -1; };
  int max_slot = 0;
  for (int i = 0; i < nnodes; ++i) {
    const fle_pred_node& nd = nodes[i];
    if (nd.op == FLE_OP_AND || nd.op == FLE_OP_OR) {
      if (flat) continue;
      const int b = stack.back(); stack.pop_back();
      const int a = stack.back();
      if (P.ncombs >= kMaxCombs) return fail(FLE_E_UNSUPPORTED, "predicate tree too large");
      P.combs[P.ncombs++] = CombDesc{(uint8_t)(nd.op == FLE_OP_AND ? FOLD_AND : FOLD_OR),
                                     (uint8_t)a, (uint8_t)a, (uint8_t)b};
      continue;
    }
    int slot = 0;
    uint8_t fold = flat_fold;
    if (!flat) {
      slot = alloc_slot();
      if (slot < 0) return fail(FLE_E_UNSUPPORTED, "too many leaves in a mixed AND/OR tree");
      stack.push_back(slot);
      fold = FOLD_SET;
      max_slot = std::max(max_slot, slot);
    }
    Chain ch{};
    ch.slot = (uint8_t)slot;
    ch.fold = fold;
    ch.order = (int)chains.size();
    if (nd.op == FLE_OP_ALL || nd.op == FLE_OP_NONE || (nd.op == FLE_OP_IN && nd.ncodes == 0)) {
      ch.type = PASS_CONST;
      ch.init = nd.op == FLE_OP_ALL ? 1 : 0;  // empty IN list: nothing matches (:8282-8290)
      chains.push_back(ch);
      continue;
    }
    const int bw = cols[nd.col].bit_width;
    const uint64_t mask = bw >= 32 ? 0xffffffffull : ((1ull << bw) - 1ull);
    ch.col = (uint8_t)nd.col;
    // A literal that does not fit the column's bit width.  The reference's plane algebra simply ignores
    // the high bits (fle-encoding.h:7965-7971; its precondition is value < 2^bw), which turns `x < 2^bw`
    // into `x < 0`; that is kept under the "reference_compat" option.  By default the comparison is
    // folded to what it means for codes below 2^bw: no row equals / exceeds it, every row is below it.
    if (nd.op == FLE_OP_IN) {
      ch.type = PASS_IN;
      ch.in_off = (uint32_t)lp.in_lits.size();
      for (int j = 0; j < nd.ncodes; ++j) {
        if (!mask_literals && nd.codes[j] > mask) continue;
        lp.in_lits.push_back((uint32_t)(nd.codes[j] & mask));
      }
      ch.in_cnt = (uint32_t)(lp.in_lits.size() - ch.in_off);
      if (ch.in_cnt == 0) {  // every literal folded away: nothing matches
        ch.type = PASS_CONST;
        ch.init = 0;
      }
      chains.push_back(ch);
      continue;
    }
    if (!mask_literals && nd.code > mask) {
      ch.type = PASS_CONST;
      ch.init = (nd.op == FLE_OP_LT || nd.op == FLE_OP_LE) ? 1 : 0;
      chains.push_back(ch);
      continue;
    }
    ch.lit = (uint32_t)(nd.code & mask);
    switch (nd.op) {
      case FLE_OP_EQ: ch.type = PASS_EQ; break;
      case FLE_OP_LT: ch.type = PASS_LT; ch.init = 0; ch.neg = 0; break;
      case FLE_OP_LE: ch.type = PASS_LT; ch.init = 1; ch.neg = 0; break;
      case FLE_OP_GE: ch.type = PASS_LT; ch.init = 0; ch.neg = 1; break;
      case FLE_OP_GT: ch.type = PASS_LT; ch.init = 1; ch.neg = 1; break;
    }
    chains.push_back(ch);
  }
  P.nslots = max_slot + 1;

  // group chains into passes: same column and type share plane fetches (up to 4 per pass)
  std::stable_sort(chains.begin(), chains.end(), [](const Chain& a, const Chain& b) {
    if (a.col != b.col) return a.col < b.col;
    return a.type < b.type;
  });
  // flat programs: the first chain executed initialises slot 0 (SET), the rest fold into it
  if (flat && !chains.empty()) chains[0].fold = FOLD_SET;
  size_t i = 0;
  while (i < chains.size()) {
    if (P.npasses >= kMaxPasses) return fail(FLE_E_UNSUPPORTED, "too many predicate leaves");
    PassDesc& pd = P.passes[P.npasses++];
    const Chain& c0 = chains[i];
    pd.type = c0.type;
    pd.col = c0.col;
    if (c0.type == PASS_IN || c0.type == PASS_CONST) {
      pd.k = 1;
      pd.slot[0] = c0.slot; pd.fold[0] = c0.fold; pd.neg[0] = c0.neg; pd.init[0] = c0.init;
      pd.lit[0] = c0.in_off; pd.lit[1] = c0.in_cnt;
      pd.k_real = 1;
      ++i;
      continue;
    }
    int k = 0;
    while (i < chains.size() && k < 4 && chains[i].type == c0.type && chains[i].col == c0.col) {
      const Chain& c = chains[i];
      pd.slot[k] = c.slot; pd.fold[k] = c.fold; pd.neg[k] = c.neg; pd.init[k] = c.init;
      pd.lit[k] = c.lit;
      ++k; ++i;
    }
    // merge: every chain lands in the same slot, chains 1.. with one and the same op
    bool merge = k > 1;
    for (int j = 1; j < k && merge; ++j)
      merge = pd.slot[j] == pd.slot[0] && pd.fold[j] == pd.fold[1] && pd.fold[j] != FOLD_SET;
    pd.merge = merge ? 1 : 0;
    // only K = 1, 2, 4 (LT) and 1, 4 (EQ) bodies exist: pad by repeating the last chain, which
    // changes nothing because SET/AND/OR with the same value are idempotent
    const int want = (k == 3 || (c0.type == PASS_EQ && k == 2)) ? 4 : k;
    for (int j = k; j < want; ++j) {
      pd.slot[j] = pd.slot[k - 1]; pd.fold[j] = pd.fold[k - 1]; pd.neg[j] = pd.neg[k - 1];
      pd.init[j] = pd.init[k - 1]; pd.lit[j] = pd.lit[k - 1];
    }
    // (a repeated SET cannot clobber later folds: in flat mode the only SET is chain 0 of the
    //  first pass and padding repeats the LAST chain of a pass with k >= 2; in mixed trees every
    //  chain SETs its own slot)
    pd.k = (uint8_t)want;
    pd.k_real = (uint32_t)k;
  }
  return FLE_OK;
}

// ------------------------------------------------------------------------------------------
// host: launch
// ------------------------------------------------------------------------------------------
namespace {

inline uint32_t round_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

constexpr int kNT = 256;
enum LaunchMode { MODE_DIRECT = 0, MODE_STAGED = 1, MODE_SINGLE_PASS = 2, MODE_IN_LOOKUP = 3, MODE_JIT = 4, MODE_NARROW = 5, MODE_IN_WIDE = 6,
                  MODE_IN_MUX = 7 };

// cudaOccupancyMaxActiveBlocksPerMultiprocessor / cudaFuncSetAttribute are not free; remember
// the answers per (kernel, dynamic smem)
int blocks_per_sm(fle_ctx* ctx, const void* kern, size_t smem_bytes, int* out, int nt = kNT) {
  for (const auto& e : ctx->occ_cache) {
    if (e.kern == kern && e.smem == smem_bytes) {
      *out = e.blocks;
      return FLE_OK;
    }
  }
  // MaxDynamicSharedMemorySize is a property of the KERNEL (per device, process-wide), not of a
  // context or a launch: raise it once to the device limit, under a lock, and never lower it
  cudaError_t e = cudaSuccess;
  {
    static std::mutex mu;
    static std::set<std::pair<int, const void*> > raised;
    std::lock_guard<std::mutex> lock(mu);
    const std::pair<int, const void*> key(ctx->device, kern);
    if (raised.find(key) == raised.end()) {
      cudaFuncAttributes fa;
      e = cudaFuncGetAttributes(&fa, kern);
      if (e != cudaSuccess) return set_cuda_error(ctx, e, "cudaFuncGetAttributes");
      // static + dynamic shared memory together are bounded by the opt-in limit
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               ctx->max_smem_optin - (int)fa.sharedSizeBytes);
      if (e != cudaSuccess) return set_cuda_error(ctx, e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
      raised.insert(key);
    }
  }
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nt, smem_bytes);
  if (e != cudaSuccess) {
    // (kernels loaded at run time: older runtimes answer only for statically linked functions)
    cudaGetLastError();
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return set_cuda_error(ctx, e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    const int regs_per_cta = ((std::max(fa.numRegs, 16) + 7) / 8 * 8) * nt;
    const size_t smem_per_cta = smem_bytes + fa.sharedSizeBytes + 1024;
    per_sm = (int)std::min<size_t>(std::min<size_t>(65536 / regs_per_cta, 2048 / nt),
                                   (size_t)ctx->max_smem_per_sm / smem_per_cta);
  }
  ctx->occ_cache.push_back({kern, smem_bytes, per_sm});
  *out = per_sm;
  return FLE_OK;
}

int launch_program(fle_ctx* ctx, EvalParams& P, int mode, int sum_bw, const void* jit_kern = nullptr,
                   void* jit_args = nullptr, bool jit_staged = true) {
  const int64_t G = P.ngroups;
  const void* kern = nullptr;
  int nt = kNT;       // threads per CTA
  int in_ncw = 16, in_gpt = 1;  // IN lookup kernel: consumer warps, groups per consumer thread per tile
  int mode_reported = mode;
  if (mode == MODE_JIT) {
    kern = jit_kern;
  } else if (mode == MODE_SINGLE_PASS) {
    kern = P.plane_skew[0] ? eval1_skew_kernel_for(P.passes[0].type, P.passes[0].k, (int)P.bw[0])
                           : eval1_kernel_for(P.passes[0].type, P.passes[0].k, (int)P.bw[0]);
    if (!kern) return FLE_E_UNSUPPORTED;
  } else if (mode == MODE_IN_LOOKUP) {
    kern = eval_in_kernel_for((int)P.bw[0]);
    if (!kern) return FLE_E_UNSUPPORTED;
  } else if (mode == MODE_IN_WIDE) {
    int want_ncw = 16, want_gpt = 1;
    if (ctx->in_wide >= 2) eval_in_wide_shape((int)P.bw[0], &want_ncw, &want_gpt);
    if (ctx->in_wide_warps > 0) want_ncw = ctx->in_wide_warps;
    if (ctx->in_wide_gpt > 0) want_gpt = ctx->in_wide_gpt;
    kern = eval_in_wide_kernel_for((int)P.bw[0], ctx->in_wide, want_ncw, want_gpt);
    if (kern) {
      in_ncw = want_ncw;
      in_gpt = want_gpt;
    } else {
      kern = eval_in_wide_kernel_for((int)P.bw[0], ctx->in_wide, 16, 1);
    }
    if (!kern) return FLE_E_UNSUPPORTED;
    nt = in_ncw * 32 + 32;
  } else if (mode == MODE_NARROW) {
    kern = eval1_narrow_kernel_for(P.passes[0].type, P.passes[0].k, (int)P.bw[0]);
    if (!kern) return FLE_E_UNSUPPORTED;
  } else if (mode == MODE_IN_MUX) {
    kern = P.plane_skew[0] ? eval_in_mux_skew_kernel_for((int)P.bw[0]) : eval_in_mux_kernel_for((int)P.bw[0]);
    if (!kern) return FLE_E_UNSUPPORTED;
    mode = MODE_SINGLE_PASS;  // same tile pipeline and launch shape as the single-pass compare kernels
    mode_reported = MODE_IN_MUX;
  } else {
    kern = mode == MODE_STAGED ? (const void*)eval_tma_kernel<kNT> : (const void*)eval_direct_kernel<kNT>;
  }
  // (a specialised kernel also runs unstaged -- planes straight from global memory -- for
  //  unaligned columns and for programs whose 256-group tile does not fit shared memory twice)
  const bool staged = mode != MODE_DIRECT && mode != MODE_NARROW && !(mode == MODE_JIT && !jit_staged);

  // tile sizing: groups per thread per tile so that a stage is a few tens of KB
  int gpt = 1;
  // single-pass kernels (tools/eval1_tune.py, profiles/r01_eval1_tune.txt): ONE CTA per SM with 2-3
  // stages of 40-64 KB beats 2-3 CTAs with the same bytes in flight by 2-8 % (fewer concurrent
  // streams for the memory system); narrow columns, whose stages stay small, take two CTAs
  int sp_ctas = 0, sp_stages = 0;
  if (staged) {
    gpt = ctx->eval_gpt;
    const uint32_t bytes_per_gpt = (uint32_t)nt * sum_bw * 8u;
    // (measured, tools/sweep.py tune: the interpreting kernel likes ~32 KB stages and 3 CTAs/SM,
    //  the multi-column specialised kernels ~90 KB with 1-2 CTAs/SM)
    const bool sp_narrow = mode == MODE_SINGLE_PASS && sum_bw <= 5;
    const uint32_t target = mode == MODE_STAGED ? 32u * 1024u : (mode == MODE_JIT ? 88u * 1024u : (sp_narrow ? 24u * 1024u : 48u * 1024u));
    if (gpt <= 0) {
      gpt = (int)std::max<uint32_t>(1, std::min<uint32_t>(8, (target + bytes_per_gpt / 2) / bytes_per_gpt));
      // single-pass tiles of 33-44 KB (bw 17-21 at one group per thread) are too short for the
      // dynamic tile scheduler (0.86-0.98 of roofline against 1.05 with two groups per thread)
      if (mode == MODE_SINGLE_PASS && gpt == 1 && bytes_per_gpt > 33u * 1024u && bytes_per_gpt < 44u * 1024u) gpt = 2;
    }
    gpt = std::min(gpt, 8);
    if (mode == MODE_IN_WIDE) gpt = in_gpt;  // compiled into the kernel
    if (mode == MODE_JIT && P.agg_plain) gpt = kJitAggGpt;  // compiled into the kernel (fle_eval_jit.cuh)
    if (mode_reported == MODE_IN_MUX) {
      // compute-heavy (2^bw instructions per 32 rows): two CTAs per SM so that one computes while the other waits
      // at its tile barrier, and tiles as large as two stages per CTA allow (bw 6, 600 M rows: 1 CTA x 3 stages x
      // 2 groups per thread 0.64 of roofline, 2 CTAs x 2 stages x 4 groups 0.79)
      if (ctx->eval_gpt <= 0) gpt = (int)std::max<uint32_t>(1, std::min<uint32_t>(8, (48u * 1024u) / bytes_per_gpt));
      sp_ctas = 2;
      sp_stages = 2;
    } else if (mode == MODE_SINGLE_PASS) {
      const uint32_t stage = (uint32_t)gpt * bytes_per_gpt;
      if (sp_narrow) {
        sp_ctas = stage < 20u * 1024u ? 3 : 2;
        sp_stages = 3;
      } else {
        sp_ctas = 1;
        sp_stages = (int)std::max<uint32_t>(2, std::min<uint32_t>(4, 150u * 1024u / stage));
      }
    }
  }
  // shared memory carve-up: [mbarriers][IN literals][literal masks][slots][stages]
  uint32_t off = 0;
  P.smem_bar_off = off;
  off += 8 * 8;  // up to 8 stages
  P.smem_lits_off = off;
  if (mode == MODE_DIRECT || mode == MODE_STAGED || mode == MODE_JIT) off = round_up(off + (uint32_t)P.n_in_lits * 4u, 16);
  P.smem_masks_off = off;
  if (mode == MODE_STAGED) off += (uint32_t)P.npasses * 512u;
  if (mode == MODE_IN_LOOKUP) {  // membership table, 2^bw bits (+ 32 lane-private copies for bw <= 13)
    const uint32_t tbytes = std::max(4u, (1u << P.bw[0]) / 8u);
    off += P.bw[0] <= 13 ? tbytes * 33u : tbytes;
  }
  if (mode == MODE_IN_WIDE) {  // lane-private entries; the kernel addresses the table at a compile-time offset
    if (off != 64u) return set_error(ctx, FLE_E_INVALID, "IN lookup table offset");
    off += eval_in_wide_table_bytes((int)P.bw[0], ctx->in_wide);
  }
  P.smem_slots_off = off;
  if (mode == MODE_DIRECT || mode == MODE_STAGED) off += (uint32_t)P.nslots * kNT * 8u;
  off = round_up(off, 128);
  P.smem_stage_off = off;

  int grid = 0;
  size_t smem_bytes = off;
  int rc;
  if (staged) {
    const int T = (mode == MODE_IN_WIDE ? in_ncw * 32 : nt) * gpt;
    uint32_t coff = 0;
    for (int c = 0; c < P.ncols; ++c) {
      P.col_off[c] = coff;
      // (+ 16: the bank skew between segments.  A skewed column stages 16 bytes more per segment and is read in
      //  whole 16-byte chunks: + 48 keeps the eight segments 3 chunks apart modulo 8, i.e. conflict-free)
      P.seg_stride[c] = (uint32_t)(T / 8) * P.bw[c] * 8u + (P.plane_skew[c] ? 48u : 16u);
      coff += 8u * P.seg_stride[c];
    }
    P.stage_bytes = round_up(coff, 128);
    // stages x resident CTAs: as many bytes in flight per SM as shared memory allows, with at
    // least 2 stages per CTA and preferably 3 CTAs (24 warps) per SM
    const uint32_t sm_budget = (uint32_t)ctx->max_smem_per_sm;
    int nst = ctx->eval_stages, want_ctas = ctx->eval_ctas_per_sm;
    if (nst <= 0 && want_ctas <= 0 && sp_stages > 0 &&
        (uint64_t)off + 1024u + (uint64_t)sp_stages * P.stage_bytes <= sm_budget / (uint32_t)sp_ctas) {
      nst = sp_stages;
      want_ctas = sp_ctas;
    }
    if (mode == MODE_IN_WIDE) {  // one CTA per SM with as many stages as fit beside the table
      const uint32_t room = (uint32_t)ctx->max_smem_optin - 512u /* static */ - off;
      nst = std::min(std::min(nst > 0 ? nst : 4, 4), (int)(room / P.stage_bytes));  // (full + empty mbarriers: at most 4 stages)
      if (nst < 2) return FLE_E_UNSUPPORTED;
      want_ctas = 1;
    }
    if (nst <= 0) {
      nst = 0;
      for (int ctas = (want_ctas > 0 ? want_ctas : 3); ctas >= 1 && nst < 2; --ctas) {
        const uint32_t per_cta = sm_budget / ctas;
        if (per_cta < off + 1024u + 2u * P.stage_bytes) continue;
        nst = (int)std::min<uint32_t>(4, (per_cta - off - 1024u) / P.stage_bytes);
      }
      if (nst < 2) return FLE_E_UNSUPPORTED;  // caller falls back to the direct kernel
    }
    nst = std::min(nst, 8);
    if ((uint64_t)off + (uint64_t)nst * P.stage_bytes > (uint64_t)ctx->max_smem_optin) return FLE_E_UNSUPPORTED;
    P.gpt = gpt;
    P.nstages = nst;
    smem_bytes = off + (size_t)nst * P.stage_bytes;
    int per_sm = 0;
    rc = blocks_per_sm(ctx, kern, smem_bytes, &per_sm, nt);
    if (rc != FLE_OK) return rc;
    if (per_sm < 1) return FLE_E_UNSUPPORTED;
    if (want_ctas > 0) per_sm = std::min(per_sm, want_ctas);
    const int64_t ntiles = G / T;
    const int64_t rest_blocks = ((G - ntiles * T) + nt - 1) / nt;
    grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * per_sm,
                                                        std::max<int64_t>(ntiles, rest_blocks)));
  } else {
    P.gpt = 1;
    P.nstages = 0;
    int per_sm = 0;
    rc = blocks_per_sm(ctx, kern, smem_bytes, &per_sm);
    if (rc != FLE_OK) return rc;
    per_sm = std::max(per_sm, 1);
    int64_t blocks = (G + kNT - 1) / kNT;
    if (mode == MODE_NARROW) {  // tiles of kNT * kNarrowUnits units (fle_eval1.cu), then one group per thread for the rest
      const int64_t per_tile = (int64_t)kNT * kNarrowUnits * (P.bw[0] == 1 ? 2 : 1);
      blocks = std::max<int64_t>(G / per_tile, (G % per_tile + kNT - 1) / kNT);
    }
    grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * per_sm, blocks));
  }
  rc = ensure_scratch(ctx, P.agg_plain ? 2 * grid : grid, P.n_in_lits);
  if (rc != FLE_OK) return rc;
  P.partial = ctx->d_partial;
  P.ticket = ctx->d_ticket;
  P.tile_counter = ctx->d_ticket + 1;
  group_fill(ctx, &P);
  void* args[] = {(void*)&P, jit_args};
  cudaError_t e = cudaLaunchKernel(kern, dim3(grid), dim3(nt), args, smem_bytes, ctx->stream);
  if (e != cudaSuccess) {
    cudaGetLastError();  // clear the non-sticky error
    char what[160];
    snprintf(what, sizeof(what), "cudaLaunchKernel(eval mode=%d grid=%d smem=%zu gpt=%d stages=%d bw0=%u k=%d type=%d)",
             mode, grid, smem_bytes, P.gpt, P.nstages, P.bw[0], (int)P.passes[0].k, (int)P.passes[0].type);
    return set_cuda_error(ctx, e, what);
  }
  ctx->launches += 1;
  ctx->last_mode = mode_reported;
  if (mode == MODE_JIT) ctx->jit_launches += 1;
  return FLE_OK;
}

}  // namespace

int prepare_eval(fle_ctx* ctx, const fle_col_desc* cols, int ncols, const fle_pred_node* nodes,
                 int nnodes, LoweredProgram* lp) {
  std::string err;
  int rc = lower_program(cols, ncols, nodes, nnodes, lp, &err, ctx->reference_compat);
  if (rc != FLE_OK) return set_error(ctx, rc, err);
  lp->lit_slot = -1;
  lp->lit_gen = 0;
  return FLE_OK;
}

// planes[c] rounded down to 16 bytes + plane_skew[c]; the staged range ends two groups early without slack
static void apply_skew(EvalParams& P, int ncols, bool slack) {
  for (int c = 0; c < ncols; ++c) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(P.planes[c]);
    P.plane_skew[c] = (uint32_t)(a & 15);
    P.planes[c] = reinterpret_cast<const uint64_t*>(a & ~(uintptr_t)15);
  }
  P.staged_groups = slack ? P.ngroups : std::max<int64_t>(0, P.ngroups - 2);
}

static int launch_prepared_skewed(fle_ctx* ctx, LoweredProgram* lp, const EvalParams& P0, const fle_col_desc* cols, int ncols,
                                  int64_t first_row, int64_t n, uint64_t* bitmap_dev, uint64_t* count_dev, const AggSpec* agg,
                                  int sum_bw, bool slack) {
  (void)cols; (void)first_row; (void)n; (void)bitmap_dev; (void)count_dev;
  if (agg) return FLE_E_UNSUPPORTED;
  const PassDesc& p0 = P0.passes[0];
  const bool single_pass = P0.npasses == 1 && P0.ncombs == 0 && (p0.type == PASS_LT || p0.type == PASS_EQ) &&
                           (p0.k == 1 || p0.merge) && p0.slot[0] == 0;
  const bool in_single = P0.npasses == 1 && P0.ncombs == 0 && p0.type == PASS_IN && p0.slot[0] == 0;
  if (single_pass || (in_single && (int)P0.bw[p0.col] <= ctx->in_mux_bw && (int)p0.lit[1] >= std::min(ctx->in_lookup_min, ctx->in_mux_min))) {
    EvalParams P1 = P0;
    const int c = p0.col;
    P1.ncols = 1;
    P1.planes[0] = P0.planes[c];
    P1.bw[0] = P0.bw[c];
    P1.passes[0].col = 0;
    apply_skew(P1, 1, slack);
    int mode = MODE_SINGLE_PASS;
    if (in_single) {
      mode = MODE_IN_MUX;
      std::memset(P1.in_masks, 0, sizeof(P1.in_masks));
      for (uint32_t j = 0; j < p0.lit[1]; ++j) {
        const uint32_t code = lp->in_lits[p0.lit[0] + j] & 127u;
        P1.in_masks[code] = in_mux_term_on_fma((int)P1.bw[0], (int)(code & 7u)) ? 1u : ~0u;
      }
    }
    const int rc = launch_program(ctx, P1, mode, (int)P1.bw[0]);
    if (rc != FLE_E_UNSUPPORTED) return rc;
  }
  if (P0.comb_op) return FLE_E_UNSUPPORTED;
  if (ctx->eval_jit == 0) return FLE_E_UNSUPPORTED;
  // specialised multi-column kernel over the images in place (the shape records that its columns are skewed)
  bool same = lp->jit_state != 0 && lp->jit_agg_col == -2;
  for (int c = 0; c < ncols && same; ++c) same = lp->jit_bw[c] == P0.bw[c];
  if (!same) {
    lp->jit_state = jit_plan(P0, lp->in_lits, &lp->jit_plan, -1, /*skew=*/true) ? 1 : -1;
    lp->jit_agg_col = -2;  // (marks a skewed plan)
    for (int c = 0; c < ncols; ++c) lp->jit_bw[c] = P0.bw[c];
    lp->jit_entry.reset();
  }
  if (lp->jit_state != 1) return FLE_E_UNSUPPORTED;
  if (!lp->jit_entry) lp->jit_entry = jit_lookup(lp->jit_plan.text, n, ctx->eval_jit_min_rows, ctx->eval_jit == 2, ctx->eval_jit == 3, ctx->device);
  if (!lp->jit_entry || lp->jit_entry->rc != FLE_OK) return FLE_E_UNSUPPORTED;
  EvalParams P1 = P0;
  apply_skew(P1, ncols, slack);
  jit_fill_args(lp->jit_plan, P1, lp->in_lits, &lp->jit_args);
  int rc = launch_program(ctx, P1, MODE_JIT, sum_bw, lp->jit_entry->kern, lp->jit_args.data(), true);
  if (rc == FLE_E_UNSUPPORTED) rc = launch_program(ctx, P1, MODE_JIT, sum_bw, lp->jit_entry->kern, lp->jit_args.data(), false);
  return rc;
}

int launch_prepared(fle_ctx* ctx, LoweredProgram* lp, const fle_col_desc* cols, int ncols,
                    int64_t first_row, int64_t n, uint64_t* bitmap_dev, uint64_t* count_dev, const AggSpec* agg);

// last resort for columns that are not even 8-byte aligned: an aligned copy of every such column in the context's
// scratch (funnel-shift kernel), then the ordinary paths
static int launch_prepared_realigned(fle_ctx* ctx, LoweredProgram* lp, const fle_col_desc* cols, int ncols, int64_t first_row,
                                     int64_t n, uint64_t* bitmap_dev, uint64_t* count_dev, const AggSpec* agg) {
  const int64_t G = (n + 63) / 64, g0 = first_row / 64;
  std::vector<fle_col_desc> dc(cols, cols + ncols);
  size_t total = 0;
  std::vector<size_t> off(ncols, 0);
  for (int c = 0; c < ncols; ++c) {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(cols[c].planes) + (uint64_t)g0 * cols[c].bit_width * 8u;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) continue;
    off[c] = total;
    total += ((size_t)G * cols[c].bit_width * 8u + 255) / 256 * 256;
  }
  if (total > ctx->realign_cap) {
    FLE_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->d_realign) cudaFree(ctx->d_realign);
    ctx->d_realign = nullptr;
    ctx->realign_cap = 0;
    FLE_CUDA_TRY(ctx, cudaMalloc(&ctx->d_realign, total));
    ctx->realign_cap = total;
  }
  for (int c = 0; c < ncols; ++c) {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(cols[c].planes) + (uint64_t)g0 * cols[c].bit_width * 8u;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      dc[c].planes = p;
      dc[c].bytes = cols[c].bytes - (uint64_t)g0 * cols[c].bit_width * 8u;
      continue;
    }
    const size_t nbytes = (size_t)G * cols[c].bit_width * 8u;
    uint8_t* dst = static_cast<uint8_t*>(ctx->d_realign) + off[c];
    const int rc = launch_realign(ctx, p, nbytes, dst);
    if (rc != FLE_OK) return rc;
    dc[c].planes = dst;
    dc[c].bytes = (nbytes + 15) / 16 * 16;
  }
  return launch_prepared(ctx, lp, dc.data(), ncols, 0, n, bitmap_dev, count_dev, agg);
}

int launch_prepared(fle_ctx* ctx, LoweredProgram* lp, const fle_col_desc* cols, int ncols,
                    int64_t first_row, int64_t n, uint64_t* bitmap_dev, uint64_t* count_dev, const AggSpec* agg) {
  if (n < 0 || first_row < 0 || (!bitmap_dev && !agg)) return set_error(ctx, FLE_E_INVALID, "bad row range or null bitmap");
  if (agg && (agg->col < 0 || agg->col >= ncols || !agg->plain || !agg->dict || !agg->sum_out))
    return set_error(ctx, FLE_E_INVALID, "bad aggregate arguments");
  if (first_row & 63) return set_error(ctx, FLE_E_ALIGN, "first_row must be a multiple of 64");
  EvalParams P = lp->p;  // per-launch copy
  const int64_t G = (n + 63) / 64;
  bool aligned = true, byte_skewed = false, slack = true;
  int sum_bw = 0;
  P.ncols = ncols;
  for (int c = 0; c < ncols; ++c) {
    if (cols[c].max_def_level != 0)
      return set_error(ctx, FLE_E_UNSUPPORTED, "nullable column reached the fused kernel");
    if (!cols[c].planes) return set_error(ctx, FLE_E_INVALID, "null column pointer");
    const int bw = cols[c].bit_width;
    const uint64_t need = (uint64_t)(first_row / 64 + G) * bw * 8u;
    if (cols[c].bytes < need) return set_error(ctx, FLE_E_BUFFER, "column shorter than the rows requested");
    const uint8_t* p = reinterpret_cast<const uint8_t*>(cols[c].planes) + (uint64_t)(first_row / 64) * bw * 8u;
    P.planes[c] = reinterpret_cast<const uint64_t*>(p);
    P.bw[c] = (uint32_t)bw;
    if (reinterpret_cast<uintptr_t>(p) & 15) aligned = false;
    if (reinterpret_cast<uintptr_t>(p) & 7) byte_skewed = true;
    // staged tiles of a skewed column read 16 bytes past their last group: keep the last groups out of the
    // staged range unless the buffer is known to extend that far
    if (cols[c].bytes < need + 16) slack = false;
    sum_bw += bw;
  }
  P.ngroups = G;
  P.nrows = n;
  P.bitmap = bitmap_dev;
  P.count_out = count_dev;
  P.n_in_lits = (int)lp->in_lits.size();
  P.in_lits = nullptr;
  if (P.n_in_lits > 0) {
    // the program's literal pool: uploaded once, again only if its slot has been reused since
    if (lp->lit_slot < 0 || ctx->lit_slots[lp->lit_slot].gen != lp->lit_gen) {
      const int rc = upload_literals(ctx, lp->in_lits.data(), P.n_in_lits, &lp->lit_slot, &lp->lit_gen, nullptr);
      if (rc != FLE_OK) return rc;
    }
    P.in_lits = ctx->lit_slots[lp->lit_slot].d;
  }
  if (agg) {
    P.agg_plain = agg->plain;
    P.agg_dict = agg->dict;
    P.agg_ndict = agg->ndict;
    P.agg_sum_out = agg->sum_out;
  }
  if (G == 0) {
    if (count_dev) FLE_CUDA_TRY(ctx, cudaMemsetAsync(count_dev, 0, 8, ctx->stream));
    if (agg) FLE_CUDA_TRY(ctx, cudaMemsetAsync(agg->sum_out, 0, 8, ctx->stream));
    return FLE_OK;
  }
  P.staged_groups = G;
  // Columns that are not 16-byte aligned (page payloads inside a file image): the skew-capable kernels read them
  // in place -- single-pass programs (fle_eval1_skew.cu) and specialised multi-column programs (FLE_JIT_SKEW);
  // everything else takes an aligned copy first (context scratch), so no caller ever has to realign.
  if (!aligned && ctx->eval_variant != 0) {
    const int rc = launch_prepared_skewed(ctx, lp, P, cols, ncols, first_row, n, bitmap_dev, count_dev, agg, sum_bw, slack);
    if (rc != FLE_E_UNSUPPORTED) return rc;
  }
  if (byte_skewed) return launch_prepared_realigned(ctx, lp, cols, ncols, first_row, n, bitmap_dev, count_dev, agg);
  const bool staged_ok = ctx->eval_variant != 0 && aligned;
  // one pass of compare chains folding into one word: the per-bit-width kernels of fle_eval1.cu
  const PassDesc& p0 = P.passes[0];
  const bool single_pass = P.npasses == 1 && P.ncombs == 0 && (p0.type == PASS_LT || p0.type == PASS_EQ) &&
                           (p0.k == 1 || p0.merge) && p0.slot[0] == 0;
  // one IN list on one dictionary-coded column: decode + membership lookup (fle_eval_in.cu), whose
  // cost does not depend on the list length.  Lists shorter than in_lookup_min are cheaper as
  // unrolled equality chains on the planes (specialised kernel below: 0.79-0.89 of the HBM roofline
  // up to 8 literals against 0.56 for the lookup); if that kernel is not available (not compiled
  // yet, no NVRTC) lists of at least kInLookupFloor literals come back to the lookup kernel.
  constexpr int kInLookupFloor = 6;
  const bool in_single = P.npasses == 1 && P.ncombs == 0 && p0.type == PASS_IN && p0.slot[0] == 0 &&
                         P.bw[p0.col] <= 16 && staged_ok && !agg;
  const auto launch_in_lookup = [&]() {
    EvalParams P1 = P;
    const int c = p0.col;
    P1.ncols = 1;
    P1.planes[0] = P.planes[c];
    P1.bw[0] = P.bw[c];
    P1.passes[0].col = 0;
    // narrow columns: the membership function as a multiplexer tree over the planes (fle_eval1.cu, InMuxRegs)
    if ((int)P.bw[c] <= ctx->in_mux_bw) {
      std::memset(P1.in_masks, 0, sizeof(P1.in_masks));
      for (uint32_t j = 0; j < p0.lit[1]; ++j) {
        const uint32_t code = lp->in_lits[p0.lit[0] + j] & 127u;
        P1.in_masks[code] = in_mux_term_on_fma((int)P.bw[c], (int)(code & 7u)) ? 1u : ~0u;
      }
      const int rc = launch_program(ctx, P1, MODE_IN_MUX, (int)P.bw[c]);
      if (rc != FLE_E_UNSUPPORTED) return rc;
    }
    // one CTA per SM, producer warp + 16 consumer warps: 16-bit table entries (one ALU instruction per
    // lookup) for bw <= 12, replicated bit table beyond (fle_eval_in.cu)
    if (ctx->in_wide != 0) {
      const int rc = launch_program(ctx, P1, MODE_IN_WIDE, (int)P.bw[c]);
      if (rc != FLE_E_UNSUPPORTED) return rc;
    }
    return launch_program(ctx, P1, MODE_IN_LOOKUP, (int)P.bw[c]);
  };
  // (the multiplexer tree of a narrow column costs 2^bw terms whatever the list holds: it beats unrolled equality
  //  chains -- 2 * bw LOP3 per literal -- from a handful of literals on: bw 5, 8 literals 0.73 -> 0.88 of roofline)
  const bool in_mux_short = in_single && (int)P.bw[p0.col] <= ctx->in_mux_bw && (int)p0.lit[1] >= ctx->in_mux_min;
  if (in_single && ((int)p0.lit[1] >= ctx->in_lookup_min || in_mux_short)) {
    const int rc = launch_in_lookup();
    if (rc != FLE_E_UNSUPPORTED) return rc;
  }
  if (P.comb_op) return FLE_E_UNSUPPORTED;  // only the single-IN-list kernels combine on the way out
  // narrow columns: consecutive groups are already 128-bit coalesced in global memory
  if (staged_ok && single_pass && ctx->eval_variant == 1 && !agg && (int)P.bw[p0.col] <= ctx->eval_narrow_bw &&
      (reinterpret_cast<uintptr_t>(bitmap_dev) & 15) == 0) {
    EvalParams P1 = P;
    const int c = p0.col;
    P1.ncols = 1;
    P1.planes[0] = P.planes[c];
    P1.bw[0] = P.bw[c];
    P1.passes[0].col = 0;
    const int rc = launch_program(ctx, P1, MODE_NARROW, (int)P.bw[c]);
    if (rc != FLE_E_UNSUPPORTED) return rc;
  }
  if (staged_ok && single_pass && ctx->eval_variant != 2 && !agg) {
    EvalParams P1 = P;  // only the pass's column is staged
    const int c = p0.col;
    P1.ncols = 1;
    P1.planes[0] = P.planes[c];
    P1.bw[0] = P.bw[c];
    P1.passes[0].col = 0;
    const int rc = launch_program(ctx, P1, MODE_SINGLE_PASS, (int)P.bw[c]);
    if (rc != FLE_E_UNSUPPORTED) return rc;
  }
  // Several passes, mixed AND/OR trees, IN lists inside larger programs: a kernel specialised
  // for the program's shape at run time (fle_jit.cu) when the call is large enough to pay for
  // one compilation, or the shape has been compiled before
  if ((ctx->eval_variant == 1 || agg) && ctx->eval_jit != 0) {
    const int agg_col = agg ? agg->col : -1;
    bool same_bw = lp->jit_state != 0 && lp->jit_agg_col == agg_col;
    for (int c = 0; c < ncols && same_bw; ++c) same_bw = lp->jit_bw[c] == P.bw[c];
    if (!same_bw) {
      lp->jit_state = jit_plan(P, lp->in_lits, &lp->jit_plan, agg_col) ? 1 : -1;
      lp->jit_agg_col = agg_col;
      for (int c = 0; c < ncols; ++c) lp->jit_bw[c] = P.bw[c];
      lp->jit_entry.reset();
    }
    if (lp->jit_state == 1) {
      if (!lp->jit_entry)
        lp->jit_entry = jit_lookup(lp->jit_plan.text, n, ctx->eval_jit_min_rows, ctx->eval_jit == 2 || agg, ctx->eval_jit == 3, ctx->device);
      if (lp->jit_entry && lp->jit_entry->rc == FLE_OK) {
        jit_fill_args(lp->jit_plan, P, lp->in_lits, &lp->jit_args);
        int rc = FLE_E_UNSUPPORTED;
        // With a fused aggregate the survivors' random loads and the warp-wide survivor walk want many
        // resident warps: 64 warps/SM on global loads beat 8 warps/SM behind the TMA stages (Q6, 600 M
        // rows: 0.41 vs 0.48 ms; DRAM traffic bound 0.37 ms), unless the caller forces a stage count.
        const bool want_staged = staged_ok && (!agg || ctx->eval_stages > 0);
        if (want_staged) rc = launch_program(ctx, P, MODE_JIT, sum_bw, lp->jit_entry->kern, lp->jit_args.data(), true);
        if (rc == FLE_E_UNSUPPORTED)
          rc = launch_program(ctx, P, MODE_JIT, sum_bw, lp->jit_entry->kern, lp->jit_args.data(), false);
        if (rc != FLE_E_UNSUPPORTED) return rc;
      } else if (lp->jit_entry && ctx->eval_jit == 2) {
        return set_error(ctx, FLE_E_UNSUPPORTED, "eval_jit = 2: " + lp->jit_entry->log);
      }
    } else if (ctx->eval_jit == 2) {
      return set_error(ctx, FLE_E_UNSUPPORTED, "eval_jit = 2: program too large for a specialised kernel");
    }
  }
  if (agg) return FLE_E_UNSUPPORTED;  // only specialised kernels fuse the aggregate
  if (in_single && ctx->in_lookup_min < (1 << 30) /* 2^30 and up: lookup kernel disabled */ &&
      (int)p0.lit[1] >= std::min(kInLookupFloor, ctx->in_lookup_min)) {
    const int rc = launch_in_lookup();
    if (rc != FLE_E_UNSUPPORTED) return rc;
  }
  // Everything else: the TMA-staged generic kernel; the direct-load kernel remains for unaligned inputs, tiles that do
  // not fit in shared memory, and as eval_variant = 0.
  if (staged_ok) {
    const int rc = launch_program(ctx, P, MODE_STAGED, sum_bw);
    if (rc != FLE_E_UNSUPPORTED) return rc;
  }
  return launch_program(ctx, P, MODE_DIRECT, sum_bw);
}

// Long IN lists inside a larger program.  The fused kernels walk an IN leaf's literal pool per
// group (O(L * bw) plane operations, like the reference, fle-encoding.h:8279-8290), so a program
// with a 256-literal list runs an order of magnitude below the roofline although the lookup
// kernels answer the same list at a cost independent of L.  AND / OR are associative and
// commutative on bitmaps, so every maximal chain of one operator that contains such a leaf is
// regrouped: each long list goes through the lookup kernel into a temporary bitmap, ALL other
// operands of the chain are joined into ONE fused sub-program (each of its columns still read
// once), and the bitmaps are combined word by word (n / 8 bytes each: 8 % of a bw-12 column).
// FLE_E_UNSUPPORTED = nothing to regroup, the caller runs the program as a whole.
constexpr int kInSplitMinLits = 16;

static int launch_eval_in_split(fle_ctx* ctx, const fle_col_desc* cols, int ncols, const fle_pred_node* nodes,
                                int nnodes, int64_t first_row, int64_t n, uint64_t* bitmap_dev,
                                uint64_t* count_dev) {
  if (nnodes < 2 || nnodes > FLE_MAX_NODES || !bitmap_dev) return FLE_E_UNSUPPORTED;
  const int min_lits = std::max(kInSplitMinLits, ctx->in_lookup_min);
  int lhs[FLE_MAX_NODES], rhs[FLE_MAX_NODES], start[FLE_MAX_NODES];
  bool has_big[FLE_MAX_NODES];
  {
    int stack[FLE_MAX_NODES], sp = 0;
    bool any = false;
    for (int i = 0; i < nnodes; ++i) {
      const fle_pred_node& nd = nodes[i];
      if (nd.op == FLE_OP_AND || nd.op == FLE_OP_OR) {
        if (sp < 2) return FLE_E_UNSUPPORTED;  // malformed: the normal path reports it
        rhs[i] = stack[--sp];
        lhs[i] = stack[--sp];
        start[i] = start[lhs[i]];
        has_big[i] = has_big[lhs[i]] || has_big[rhs[i]];
      } else {
        lhs[i] = rhs[i] = -1;
        start[i] = i;
        const bool leaf_col = nd.op != FLE_OP_ALL && nd.op != FLE_OP_NONE;
        if (leaf_col && (nd.col < 0 || nd.col >= ncols || cols[nd.col].max_def_level != 0)) return FLE_E_UNSUPPORTED;
        has_big[i] = nd.op == FLE_OP_IN && nd.ncodes >= min_lits && cols[nd.col].bit_width <= 16;
        any = any || has_big[i];
      }
      stack[sp++] = i;
    }
    if (sp != 1 || !any) return FLE_E_UNSUPPORTED;
  }
  // The common shape -- ONE chain of AND (or OR) at the root whose operands are long lists and subtrees without
  // one -- needs no temporary bitmap and no combine pass: the other operands run as one fused program straight
  // into the caller's bitmap, then every long list goes through its lookup kernel, which folds the bitmap in on
  // the way out (EvalParams::comb_*), in place; the last kernel also counts.
  {
    const int root = nnodes - 1, op = nodes[root].op;
    if (op == FLE_OP_AND || op == FLE_OP_OR) {
      std::vector<int> operands, todo(1, root), big;
      std::vector<fle_pred_node> joined;
      int njoined = 0;
      bool flat_ok = true;
      while (!todo.empty()) {
        const int k = todo.back();
        todo.pop_back();
        if (nodes[k].op == op && has_big[k]) {
          todo.push_back(rhs[k]);
          todo.push_back(lhs[k]);
        } else {
          operands.push_back(k);
        }
      }
      for (int k : operands) {
        if (!has_big[k]) {
          joined.insert(joined.end(), nodes + start[k], nodes + k + 1);
          if (++njoined > 1) joined.push_back(fle_pred_node{op, 0, 0, nullptr, 0, 0});
        } else if (nodes[k].op == FLE_OP_IN) {
          big.push_back(k);
        } else {
          flat_ok = false;  // a long list below another operator: the general regrouping below
        }
      }
      if (flat_ok && !big.empty()) {
        int rc = FLE_OK;
        bool have = false;
        if (njoined > 0) {
          std::vector<int> map(ncols, -1);
          std::vector<fle_col_desc> dc;
          bool consts_only = true;
          for (auto& nd : joined) {
            if (nd.op == FLE_OP_AND || nd.op == FLE_OP_OR || nd.op == FLE_OP_ALL || nd.op == FLE_OP_NONE) continue;
            consts_only = false;
            if (map[nd.col] < 0) { map[nd.col] = (int)dc.size(); dc.push_back(cols[nd.col]); }
            nd.col = map[nd.col];
          }
          if (consts_only) dc.push_back(cols[0]);  // constants need a column to launch on
          LoweredProgram lpj;
          rc = prepare_eval(ctx, dc.data(), (int)dc.size(), joined.data(), (int)joined.size(), &lpj);
          if (rc == FLE_OK) rc = launch_prepared(ctx, &lpj, dc.data(), (int)dc.size(), first_row, n, bitmap_dev, nullptr);
          if (rc != FLE_OK) return rc;
          have = true;
        }
        for (size_t b = 0; b < big.size() && rc == FLE_OK; ++b) {
          fle_pred_node leaf = nodes[big[b]];
          const fle_col_desc dc = cols[leaf.col];
          leaf.col = 0;
          LoweredProgram lpb;
          rc = prepare_eval(ctx, &dc, 1, &leaf, 1, &lpb);
          if (rc != FLE_OK) return rc;
          if (have) {
            lpb.p.comb_bm = bitmap_dev;
            lpb.p.comb_op = op == FLE_OP_AND ? FOLD_AND : FOLD_OR;
          }
          rc = launch_prepared(ctx, &lpb, &dc, 1, first_row, n, bitmap_dev, b + 1 == big.size() ? count_dev : nullptr);
          if (rc == FLE_E_UNSUPPORTED && have) {
            // (no lookup kernel for this list, e.g. the lookup kernels are switched off: materialise and combine word by word)
            flat_ok = false;
            break;
          }
          if (rc != FLE_OK) return rc;
          have = true;
        }
        if (flat_ok) return rc;
      }
    }
  }
  const int64_t nwords = (n + 63) / 64;
  // temporary bitmaps: one per long list + at most one per operator chain, carved out of a grow-only
  // buffer of the context (stream-ordered allocations would go back to the driver at every
  // synchronisation: 75 MB bitmaps, several hundred microseconds per call)
  int nbig = 0, nops = 0;
  for (int i = 0; i < nnodes; ++i) {
    if (nodes[i].op == FLE_OP_AND || nodes[i].op == FLE_OP_OR) nops += has_big[i] ? 1 : 0;
    else nbig += has_big[i] ? 1 : 0;
  }
  const size_t bm_bytes = ((size_t)std::max<int64_t>(nwords, 1) * 8 + 255) / 256 * 256;
  const size_t need = bm_bytes * (size_t)(nbig + nops);
  if (need > ctx->split_cap) {
    FLE_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->d_split) cudaFree(ctx->d_split);
    ctx->d_split = nullptr;
    ctx->split_cap = 0;
    FLE_CUDA_TRY(ctx, cudaMalloc(&ctx->d_split, need));
    ctx->split_cap = need;
  }
  std::vector<const uint64_t*> leaf_bm;
  int rc = FLE_OK;
  // evaluates a column-space postfix fragment into a fresh bitmap, returns its index in leaf_bm
  const auto materialise = [&](std::vector<fle_pred_node>& frag) -> int {
    uint64_t* lb = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(ctx->d_split) + bm_bytes * leaf_bm.size());
    std::vector<int> map(ncols, -1);
    std::vector<fle_col_desc> dc;
    for (auto& nd : frag) {
      if (nd.op == FLE_OP_AND || nd.op == FLE_OP_OR || nd.op == FLE_OP_ALL || nd.op == FLE_OP_NONE) continue;
      if (map[nd.col] < 0) { map[nd.col] = (int)dc.size(); dc.push_back(cols[nd.col]); }
      nd.col = map[nd.col];
    }
    LoweredProgram lp;
    rc = prepare_eval(ctx, dc.data(), (int)dc.size(), frag.data(), (int)frag.size(), &lp);
    if (rc == FLE_OK) rc = launch_prepared(ctx, &lp, dc.data(), (int)dc.size(), first_row, n, lb, nullptr);
    if (rc != FLE_OK) return -1;
    leaf_bm.push_back(lb);
    return (int)leaf_bm.size() - 1;
  };
  // node -> column-space fragment (no long list below) or bitmap-space postfix over leaf_bm
  struct Res { bool comb; std::vector<fle_pred_node> prog; };
  std::function<Res(int)> emit = [&](int i) -> Res {
    Res r;
    if (!has_big[i]) {
      r.comb = false;
      r.prog.assign(nodes + start[i], nodes + i + 1);
      return r;
    }
    r.comb = true;
    const int op = nodes[i].op;
    if (op != FLE_OP_AND && op != FLE_OP_OR) {  // the long list itself
      std::vector<fle_pred_node> leaf(1, nodes[i]);
      const int b = materialise(leaf);
      r.prog.push_back(fle_pred_node{FLE_OP_EQ, (int32_t)b, 0, nullptr, 0, 0});  // the combine kernel only looks at .col
      return r;
    }
    // operands of the maximal chain of `op` through subtrees that hold a long list
    std::vector<int> operands, todo(1, i);
    while (!todo.empty()) {
      const int k = todo.back();
      todo.pop_back();
      if (nodes[k].op == op && has_big[k]) {
        todo.push_back(rhs[k]);
        todo.push_back(lhs[k]);
      } else {
        operands.push_back(k);
      }
    }
    std::vector<fle_pred_node> joined;  // all the other operands as ONE fused program
    int njoined = 0, ncomb = 0;
    for (int k : operands) {
      if (rc != FLE_OK) break;
      Res sub = emit(k);
      if (!sub.comb) {  // constants only: straight into the combine program, there is no column to launch on
        bool consts = true;
        for (const auto& nd : sub.prog)
          consts = consts && (nd.op == FLE_OP_ALL || nd.op == FLE_OP_NONE || nd.op == FLE_OP_AND || nd.op == FLE_OP_OR);
        sub.comb = consts;
      }
      if (sub.comb) {
        r.prog.insert(r.prog.end(), sub.prog.begin(), sub.prog.end());
        if (++ncomb > 1) r.prog.push_back(fle_pred_node{op, 0, 0, nullptr, 0, 0});
      } else {
        joined.insert(joined.end(), sub.prog.begin(), sub.prog.end());
        if (++njoined > 1) joined.push_back(fle_pred_node{op, 0, 0, nullptr, 0, 0});
      }
    }
    if (rc == FLE_OK && njoined > 0) {
      const int b = materialise(joined);
      r.prog.push_back(fle_pred_node{FLE_OP_EQ, (int32_t)b, 0, nullptr, 0, 0});
      if (ncomb > 0) r.prog.push_back(fle_pred_node{op, 0, 0, nullptr, 0, 0});
    }
    return r;
  };
  Res root = emit(nnodes - 1);
  if (rc == FLE_OK)
    rc = launch_bitmap_combine(ctx, leaf_bm.data(), (int)leaf_bm.size(), root.prog.data(), (int)root.prog.size(), n,
                               bitmap_dev, count_dev);
  return rc;
}

int launch_eval(fle_ctx* ctx, const fle_col_desc* cols, int ncols, const fle_pred_node* nodes,
                int nnodes, int64_t first_row, int64_t n, uint64_t* bitmap_dev,
                uint64_t* count_dev) {
  if (nnodes > 1 && ctx->in_split_min_rows >= 0 && n >= ctx->in_split_min_rows && ctx->eval_variant != 0 &&
      ctx->in_lookup_min < (1 << 30)) {
    const int rc = launch_eval_in_split(ctx, cols, ncols, nodes, nnodes, first_row, n, bitmap_dev, count_dev);
    if (rc != FLE_E_UNSUPPORTED) return rc;
  }
  LoweredProgram lp;
  int rc = prepare_eval(ctx, cols, ncols, nodes, nnodes, &lp);
  if (rc != FLE_OK) return rc;
  return launch_prepared(ctx, &lp, cols, ncols, first_row, n, bitmap_dev, count_dev);
}

// ------------------------------------------------------------------------------------------
// word-wise combine of already materialised leaf bitmaps (nullable-column path: the
// reference evaluates every leaf into its own bitset and `&=` / `|=`s them,
// simple-predicates.h:145-163)
// ------------------------------------------------------------------------------------------
struct CombineParams {
  const uint64_t* leaf[FLE_MAX_NODES];
  int8_t op[FLE_MAX_NODES];   // >= 0: push leaf[op]; -1 AND; -2 OR; -3 ALL; -4 NONE
  int32_t nnodes;
  int64_t nwords, nbits;
  uint64_t* out;
  uint64_t* partial;
  unsigned int* ticket;
  uint64_t* count_out;
};

__global__ void __launch_bounds__(256) bitmap_combine_kernel(const __grid_constant__ CombineParams C) {
  uint64_t cnt = 0;
  for (int64_t w = (int64_t)blockIdx.x * 256 + threadIdx.x; w < C.nwords; w += (int64_t)gridDim.x * 256) {
    uint64_t st[16];
    int sp = 0;
    for (int i = 0; i < C.nnodes; ++i) {
      const int o = C.op[i];
      if (o >= 0) st[sp++] = C.leaf[o][w];
      else if (o == -3) st[sp++] = ~0ull;
      else if (o == -4) st[sp++] = 0ull;
      else {
        const uint64_t b = st[--sp];
        st[sp - 1] = (o == -1) ? (st[sp - 1] & b) : (st[sp - 1] | b);
      }
    }
    uint64_t r = st[0];
    const int64_t left = C.nbits - w * 64;
    if (left < 64) r &= (1ull << left) - 1ull;
    C.out[w] = r;
    cnt += __popcll(r);
  }
  // same "last block sums" publication protocol as the eval kernels
  __shared__ uint64_t warp_cnt[8];
  __shared__ bool is_last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t t = 0;
    for (int w = 0; w < 8; ++w) t += warp_cnt[w];
    C.partial[blockIdx.x] = t;
    __threadfence();
    is_last = (atomicAdd(C.ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {  // all threads of the last CTA (a serial loop over ~1200 partials costs 20-30 us)
    __threadfence();
    uint64_t t = 0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += 256) t += *((volatile uint64_t*)&C.partial[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint64_t tot = 0;
      for (int w = 0; w < 8; ++w) tot += warp_cnt[w];
      if (C.count_out) *C.count_out = tot;
      *C.ticket = 0;
    }
  }
}

int launch_bitmap_combine(fle_ctx* ctx, const uint64_t* const* leaf_bitmaps_dev, int nleaf,
                          const fle_pred_node* nodes, int nnodes, int64_t nbits, uint64_t* out_dev,
                          uint64_t* count_dev) {
  if (nnodes < 1 || nnodes > FLE_MAX_NODES || nleaf > FLE_MAX_NODES)
    return set_error(ctx, FLE_E_UNSUPPORTED, "combine program too large");
  CombineParams C;
  std::memset(&C, 0, sizeof(C));
  int depth = 0, max_depth = 0;
  for (int i = 0; i < nnodes; ++i) {
    const int op = nodes[i].op;
    if (op == FLE_OP_AND || op == FLE_OP_OR) {
      if (depth < 2) return set_error(ctx, FLE_E_INVALID, "malformed postfix program");
      C.op[i] = op == FLE_OP_AND ? -1 : -2;
      --depth;
    } else if (op == FLE_OP_ALL) {
      C.op[i] = -3; ++depth;
    } else if (op == FLE_OP_NONE) {
      C.op[i] = -4; ++depth;
    } else {
      if (nodes[i].col < 0 || nodes[i].col >= nleaf) return set_error(ctx, FLE_E_INVALID, "bad leaf index");
      C.op[i] = (int8_t)nodes[i].col;
      ++depth;
    }
    max_depth = std::max(max_depth, depth);
  }
  if (depth != 1) return set_error(ctx, FLE_E_INVALID, "malformed postfix program");
  if (max_depth > 16) return set_error(ctx, FLE_E_UNSUPPORTED, "predicate tree too deep");
  for (int i = 0; i < nleaf; ++i) C.leaf[i] = leaf_bitmaps_dev[i];
  C.nnodes = nnodes;
  C.nbits = nbits;
  C.nwords = (nbits + 63) / 64;
  C.out = out_dev;
  C.count_out = count_dev;
  if (C.nwords == 0) {
    if (count_dev) FLE_CUDA_TRY(ctx, cudaMemsetAsync(count_dev, 0, 8, ctx->stream));
    return FLE_OK;
  }
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * 8, (C.nwords + 255) / 256));
  int rc = ensure_scratch(ctx, grid, 0);
  if (rc != FLE_OK) return rc;
  C.partial = ctx->d_partial;
  C.ticket = ctx->d_ticket;
  bitmap_combine_kernel<<<grid, 256, 0, ctx->stream>>>(C);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "bitmap_combine launch");
  ctx->launches += 1;
  return FLE_OK;
}

}  // namespace fle
