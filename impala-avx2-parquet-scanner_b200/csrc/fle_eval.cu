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
// Kernel 1 (eval_tma_kernel, default): persistent CTAs; every tile of T = blockDim*gpt groups
// is staged global -> shared by TMA bulk copies (cp.async.bulk, one per column per segment)
// into an nstages-deep mbarrier ring, so ~50-150 KB per SM are in flight with no register
// cost.  The tile of each column is split into 8 segments skewed by 16 B so that the
// per-thread (strided by bw*8 bytes) 64/128-bit shared loads are bank-conflict free for
// every bit width.
// Kernel 0 (eval_direct_kernel): same algebra, planes read straight from global memory;
// used for unaligned inputs, the ragged last tile's groups and as an on-device cross-check.
#include <algorithm>
#include <cstring>

#include "fle_device.cuh"
#include "fle_internal.h"

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

// plane order -> row order (fle-encoding.h:7989-7998 does this with six swap steps), then
// clear the bits of rows past the end (the reference pushes only num_rows bits, :8000-8006).
__device__ __forceinline__ uint64_t finish_word(uint64_t planes_order, int64_t g, int64_t nrows) {
  uint64_t r = __brevll(planes_order);
  const int64_t rows_left = nrows - g * 64;
  if (rows_left < 64) r &= (rows_left <= 0) ? 0ull : ((1ull << rows_left) - 1ull);
  return r;
}

// block-wide popcount reduction + "last block sums" so the launch needs no memset/atomics
// on the result slot.
template <int NT>
__device__ __forceinline__ void publish_count(const EvalParams& P, uint64_t cnt) {
  __shared__ uint64_t warp_cnt[NT / 32];
  __shared__ bool is_last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t t = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) t += warp_cnt[w];
    P.partial[blockIdx.x] = t;
    __threadfence();
    const unsigned int prev = atomicAdd(P.ticket, 1u);
    is_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    uint64_t t = 0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += NT)
      t += *((volatile uint64_t*)&P.partial[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint64_t tot = 0;
#pragma unroll
      for (int w = 0; w < NT / 32; ++w) tot += warp_cnt[w];
      if (P.count_out) *P.count_out = tot;
      *P.ticket = 0;  // self-reset for the next launch on this stream
    }
  }
}

__device__ __forceinline__ void load_in_lits(const EvalParams& P, uint32_t* lits_smem) {
  for (int i = threadIdx.x; i < P.n_in_lits; i += blockDim.x) lits_smem[i] = P.in_lits[i];
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
// Kernel 1: TMA-staged persistent pipeline
// ------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT) eval_tma_kernel(const __grid_constant__ EvalParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint32_t* lits_smem = reinterpret_cast<uint32_t*>(smem + P.smem_lits_off);
  const uint32_t bar0 = smem_u32(smem + P.smem_bar_off);
  const uint32_t stage0 = smem_u32(smem + P.smem_stage_off);
  const int T = NT * P.gpt;          // groups per tile
  const int seg_groups = T >> 3;     // groups per skewed segment
  const int64_t ntiles = P.ngroups / T;  // full tiles; the ragged rest goes the direct way
  const int nst = P.nstages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) mbar_init(bar0 + s * 8, 1);
    mbar_fence_init();
  }
  load_in_lits(P, lits_smem);
  __syncthreads();

  // producer side (thread 0): stage `stage` <- tile `tile`
  auto issue = [&](int stage, int64_t tile) {
    const uint32_t bar = bar0 + stage * 8;
    uint32_t total = 0;
    for (int c = 0; c < P.ncols; ++c) total += (uint32_t)T * P.bw[c] * 8u;
    mbar_arrive_expect_tx(bar, total);
    for (int c = 0; c < P.ncols; ++c) {
      const uint32_t seg_bytes = (uint32_t)seg_groups * P.bw[c] * 8u;
      const uint8_t* src = reinterpret_cast<const uint8_t*>(P.planes[c]) +
                           (uint64_t)tile * T * P.bw[c] * 8u;
      const uint32_t dst = stage0 + stage * P.stage_bytes + P.col_off[c];
#pragma unroll
      for (int s = 0; s < 8; ++s)
        tma_load_1d(dst + s * P.seg_stride[c], src + (uint64_t)s * seg_bytes, seg_bytes, bar);
    }
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < nst - 1; ++s) {
      const int64_t tile = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
      if (tile < ntiles) issue(s, tile);
    }
  }

  Slots sl;
  sl.smem = smem_u32(smem + P.smem_slots_off) + threadIdx.x * 8;
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
    for (int q = 0; q < P.gpt; ++q) {
      const int idx = q * NT + threadIdx.x;
      const int seg = idx & 7, j = idx >> 3;
      const int64_t g = tile * T + (int64_t)seg * seg_groups + j;
      auto loader_of = [&](int col) {
        return SmemPlanes{sbase + P.col_off[col] + seg * P.seg_stride[col] +
                          (uint32_t)j * P.bw[col] * 8u};
      };
      const uint64_t r = finish_word(eval_group<SmemPlanes>(P, lits_smem, sl, loader_of), g, P.nrows);
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
                  LoweredProgram* out, std::string* err) {
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
    if (nd.op == FLE_OP_IN) {
      ch.type = PASS_IN;
      ch.in_off = (uint32_t)lp.in_lits.size();
      ch.in_cnt = (uint32_t)nd.ncodes;
      for (int j = 0; j < nd.ncodes; ++j) lp.in_lits.push_back((uint32_t)(nd.codes[j] & mask));
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
    pd.k = (uint8_t)k;
  }
  return FLE_OK;
}

// ------------------------------------------------------------------------------------------
// host: launch
// ------------------------------------------------------------------------------------------
namespace {

inline uint32_t round_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

template <int NT>
int launch_with_block(fle_ctx* ctx, EvalParams& P, bool use_tma, int sum_bw) {
  // shared memory carve-up: [mbarriers][IN literals][slots][stages]
  uint32_t off = 0;
  P.smem_bar_off = off;
  off += 8 * 8;  // up to 8 stages
  P.smem_lits_off = off;
  off = round_up(off + (uint32_t)P.n_in_lits * 4u, 16);
  P.smem_slots_off = off;
  off += (uint32_t)std::max(0, P.nslots - 1) * NT * 8u;
  off = round_up(off, 128);
  P.smem_stage_off = off;

  const int64_t G = P.ngroups;
  int grid = 0;
  size_t smem_bytes = off;
  const void* kern = nullptr;
  if (use_tma) {
    // tile sizing: aim at ~32 KB per stage; narrow programs take several groups per thread
    int gpt = ctx->eval_gpt;
    const uint32_t bytes_per_gpt = (uint32_t)NT * sum_bw * 8u;
    if (gpt <= 0) gpt = (int)std::max<uint32_t>(1, std::min<uint32_t>(8, (32u * 1024u) / bytes_per_gpt));
    const int T = NT * gpt;
    uint32_t coff = 0;
    for (int c = 0; c < P.ncols; ++c) {
      P.col_off[c] = coff;
      P.seg_stride[c] = (uint32_t)(T / 8) * P.bw[c] * 8u + 16u;
      coff += 8u * P.seg_stride[c];
    }
    P.stage_bytes = round_up(coff, 128);
    const uint32_t avail = (uint32_t)ctx->max_smem_optin - off - 1024u;
    int nst = ctx->eval_stages;
    if (nst <= 0) {
      // ~100 KB per CTA keeps two CTAs per SM resident with >= 2 tiles each in flight
      nst = (int)std::min<uint32_t>(4, std::max<uint32_t>(2, (96u * 1024u) / P.stage_bytes));
    }
    nst = std::min(nst, 8);
    if (nst < 2 || (uint64_t)nst * P.stage_bytes > avail) {
      nst = (int)std::min<uint32_t>(8, avail / P.stage_bytes);
      if (nst < 2) return FLE_E_UNSUPPORTED;  // caller falls back to the direct kernel
    }
    P.gpt = gpt;
    P.nstages = nst;
    smem_bytes = off + (size_t)nst * P.stage_bytes;
    kern = (const void*)eval_tma_kernel<NT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return set_cuda_error(ctx, e, "cudaFuncSetAttribute(eval_tma)");
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem_bytes);
    if (e != cudaSuccess) return set_cuda_error(ctx, e, "occupancy(eval_tma)");
    if (per_sm < 1) return FLE_E_UNSUPPORTED;
    if (ctx->eval_ctas_per_sm > 0) per_sm = std::min(per_sm, ctx->eval_ctas_per_sm);
    const int64_t ntiles = G / T;
    const int64_t rest_blocks = ((G - ntiles * T) + NT - 1) / NT;
    grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * per_sm,
                                                        std::max<int64_t>(ntiles, rest_blocks)));
  } else {
    P.gpt = 1;
    P.nstages = 0;
    kern = (const void*)eval_direct_kernel<NT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return set_cuda_error(ctx, e, "cudaFuncSetAttribute(eval_direct)");
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem_bytes);
    if (e != cudaSuccess) return set_cuda_error(ctx, e, "occupancy(eval_direct)");
    per_sm = std::max(per_sm, 1);
    grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * per_sm, (G + NT - 1) / NT));
  }
  int rc = ensure_scratch(ctx, grid, P.n_in_lits);
  if (rc != FLE_OK) return rc;
  P.partial = ctx->d_partial;
  P.ticket = ctx->d_ticket;
  P.in_lits = ctx->d_in_lits;
  void* args[] = {(void*)&P};
  cudaError_t e = cudaLaunchKernel(kern, dim3(grid), dim3(NT), args, smem_bytes, ctx->stream);
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "cudaLaunchKernel(eval)");
  ctx->launches += 1;
  return FLE_OK;
}

}  // namespace

int prepare_eval(fle_ctx* ctx, const fle_col_desc* cols, int ncols, const fle_pred_node* nodes,
                 int nnodes, LoweredProgram* lp) {
  std::string err;
  int rc = lower_program(cols, ncols, nodes, nnodes, lp, &err);
  if (rc != FLE_OK) return set_error(ctx, rc, err);
  lp->lits_uploaded = false;
  const int nl = (int)lp->in_lits.size();
  if (nl > 0) {
    rc = ensure_scratch(ctx, 1, nl);
    if (rc != FLE_OK) return rc;
    // the pinned staging buffer is reused per program: let the previous copy drain first
    FLE_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    std::memcpy(ctx->h_in_lits, lp->in_lits.data(), (size_t)nl * 4);
    FLE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_in_lits, ctx->h_in_lits, (size_t)nl * 4,
                                      cudaMemcpyHostToDevice, ctx->stream));
    lp->lits_uploaded = true;
  }
  return FLE_OK;
}

int launch_prepared(fle_ctx* ctx, LoweredProgram* lp, const fle_col_desc* cols, int ncols,
                    int64_t first_row, int64_t n, uint64_t* bitmap_dev, uint64_t* count_dev) {
  if (n < 0 || first_row < 0 || !bitmap_dev) return set_error(ctx, FLE_E_INVALID, "bad row range or null bitmap");
  if (first_row & 63) return set_error(ctx, FLE_E_ALIGN, "first_row must be a multiple of 64");
  EvalParams P = lp->p;  // per-launch copy
  const int64_t G = (n + 63) / 64;
  bool aligned = true;
  int sum_bw = 0;
  P.ncols = ncols;
  for (int c = 0; c < ncols; ++c) {
    if (cols[c].max_def_level != 0)
      return set_error(ctx, FLE_E_UNSUPPORTED, "nullable column reached the fused kernel");
    if (!cols[c].planes) return set_error(ctx, FLE_E_INVALID, "null column pointer");
    const int bw = cols[c].bit_width;
    const uint64_t need = (uint64_t)(first_row / 64 + G) * bw * 8u;
    if (cols[c].bytes < need) return set_error(ctx, FLE_E_BUFFER, "column shorter than the rows requested");
    if (reinterpret_cast<uintptr_t>(cols[c].planes) & 7)
      return set_error(ctx, FLE_E_ALIGN, "device plane pointers must be 8-byte aligned");
    const uint8_t* p = reinterpret_cast<const uint8_t*>(cols[c].planes) + (uint64_t)(first_row / 64) * bw * 8u;
    P.planes[c] = reinterpret_cast<const uint64_t*>(p);
    P.bw[c] = (uint32_t)bw;
    if (reinterpret_cast<uintptr_t>(p) & 15) aligned = false;
    sum_bw += bw;
  }
  P.ngroups = G;
  P.nrows = n;
  P.bitmap = bitmap_dev;
  P.count_out = count_dev;
  P.n_in_lits = (int)lp->in_lits.size();
  if (G == 0) {
    if (count_dev) FLE_CUDA_TRY(ctx, cudaMemsetAsync(count_dev, 0, 8, ctx->stream));
    return FLE_OK;
  }
  const bool use_tma = ctx->eval_variant != 0 && aligned;
  const int NT = ctx->eval_block == 128 ? 128 : 256;
  int rc;
  if (use_tma) {
    rc = NT == 128 ? launch_with_block<128>(ctx, P, true, sum_bw) : launch_with_block<256>(ctx, P, true, sum_bw);
    if (rc == FLE_E_UNSUPPORTED && NT == 256) rc = launch_with_block<128>(ctx, P, true, sum_bw);
    if (rc != FLE_E_UNSUPPORTED) return rc;
  }
  return NT == 128 ? launch_with_block<128>(ctx, P, false, sum_bw) : launch_with_block<256>(ctx, P, false, sum_bw);
}

int launch_eval(fle_ctx* ctx, const fle_col_desc* cols, int ncols, const fle_pred_node* nodes,
                int nnodes, int64_t first_row, int64_t n, uint64_t* bitmap_dev,
                uint64_t* count_dev) {
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
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    uint64_t tot = 0;
    for (unsigned int i = 0; i < gridDim.x; ++i) tot += *((volatile uint64_t*)&C.partial[i]);
    if (C.count_out) *C.count_out = tot;
    *C.ticket = 0;
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
