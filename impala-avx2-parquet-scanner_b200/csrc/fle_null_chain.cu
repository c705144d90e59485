// Nullable columns with NULLs present: a whole conjunct (or disjunct) chain in ONE kernel.
//
// The reference evaluates every leaf of a nullable column in the data-ordinal domain (the data stream holds only the
// non-null values) and deposits the leaf's bits through the column's not-null bitmap, then ANDs the conjuncts
// (ColumnReader<T>::Eq..In + IntersectBitset, /root/reference/hdfs-parquet-scanner.cc:326-345; conjunct `&=`
// :1857-1862).  The per-column kernels of fle_eval1_deposit.cu fold one column per launch into the output bitmap
// (block counts + prefix + evaluate-and-deposit each: Q6 over three columns = 10 launches, the output read and
// written once per column, every def-level stream read three times).  Here the chain is one pass over the rows:
//   launch 1  def_tile_counts_kernel   non-null rows per column (what the host needs to pick the path at all) AND
//                                      per 256-word tile, all columns at once (blockIdx.y = column)
//   launch 2  nc_prefix_kernel         tile counts -> exclusive prefixes, one CTA per column
//   launch 3  null_chain_kernel        persistent CTAs, one 16384-row tile per iteration: the tile's rows map to the
//                                      <= 258 consecutive data groups of every column starting at the tile's prefix,
//                                      which the copy engine stages (cp.async.bulk, mbarrier, two stages: the next
//                                      tile lands while this one is evaluated); every staged group is evaluated once
//                                      with the borrow-chain plane algebra (bit width and literals read at run time:
//                                      one kernel for every program shape), every thread deposits the result bits at
//                                      its row word's rank into the set positions of its not-null word and folds the
//                                      columns in registers; the output word is written once.
// Bytes per 64 rows: def levels twice (counts, deposit), the data once, the output once -- Q6 with 2 % NULLs:
// 228 against ~284 + three temporaries' worth of launches.  A required column in the chain is a column whose rows
// are all present (rank = row).  Covers chains of up to kNcMaxCols single-column subtrees that each lower to one pass
// (a compare, BETWEEN, <= 4 compares folded by one operator, equalities); anything else takes the per-column path.
#include <algorithm>
#include <mutex>
#include <set>
#include <utility>

#include "fle_bitmap.cuh"
#include "fle_eval_common.cuh"
#include "fle_internal.h"

namespace fle {

constexpr int kNcThreads = 256;  // one row word per thread and tile
constexpr int kNcWords = 256;    // row words per tile (16384 rows)
constexpr int kNcSeg = 32;       // staged groups per segment: thread (seg, j) evaluates group seg * 32 + j; a tile whose rows are
constexpr int kNcOver = 4;       // (nearly) all present spans up to 258 groups: the last ones go to a small overflow segment
constexpr int kNcGroups = 8 * kNcSeg + kNcOver + 12;  // result words per column (+ slack for the funnel's second word)
constexpr int kNcStages = 2;

struct NcCol {
  const uint64_t* planes;      // data stream (the non-null values; every row for a required column)
  const uint64_t* def_planes;  // 1-bit def levels; nullptr: required column
  const uint64_t* pref;        // exclusive prefixes of the tiles' non-null counts, [ntiles + 1]
  int64_t max_groups;          // groups the column buffer holds
  uint32_t bw;
  uint32_t col_off;     // byte offset of the column's region inside a stage
  uint32_t seg_stride;  // byte distance between the segments' images (== 16 mod 128: conflict-free group reads)
  uint32_t res_off;     // byte offset of the column's result words (after the stages)
  PassDesc pd;
};

struct NcParams {
  NcCol col[kNcMaxCols];
  int32_t ncols, join_op;  // FOLD_AND / FOLD_OR
  int64_t n_root, nwords, ntiles;
  uint32_t stage_bytes, pad;
  uint64_t* out;
  unsigned long long* count;
};

// ------------------------------------------------------------------------------------------
// launch 1: non-null rows per column and per tile.  One warp per tile (8 coalesced 256-byte loads in flight per
// lane group), grid-stride; the column totals feed the host's decision (any NULL at all?), the tile counts the
// prefix kernel.
// ------------------------------------------------------------------------------------------
struct DefTileParams {
  const uint64_t* planes[FLE_MAX_COLS];
  uint32_t* tile_sums[FLE_MAX_COLS];
  int64_t nwords, nrows, ntiles;
  unsigned long long* counts;
};

__global__ void __launch_bounds__(256) def_tile_counts_kernel(const __grid_constant__ DefTileParams P) {
  __shared__ uint64_t warp_buf[8];
  const uint64_t* pl = P.planes[blockIdx.y];
  if (!pl) return;
  uint32_t* sums = P.tile_sums[blockIdx.y];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t tot = 0;
  for (int64_t t = (int64_t)blockIdx.x * 8 + warp; t < P.ntiles; t += (int64_t)gridDim.x * 8) {
    const int64_t w0 = t * kNcWords + lane;
    uint64_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (w0 + j * 32 < P.nwords) ? ldg_stream64(pl + w0 + j * 32) : 0ull;
    uint32_t c = 0;
    if ((t + 1) * kNcWords * 64 <= P.nrows) {
#pragma unroll
      for (int j = 0; j < 8; ++j) c += __popcll(v[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int64_t left = P.nrows - (w0 + j * 32) * 64;  // plane order: row k of the group at bit 63 - k
        if (left < 64) v[j] = left <= 0 ? 0ull : (v[j] & (~0ull << (64 - left)));
        c += __popcll(v[j]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) {
      sums[t] = c;
      tot += c;
    }
  }
  const uint64_t b = block_sum<256>(tot, warp_buf);
  if (threadIdx.x == 0 && b) atomicAdd(&P.counts[blockIdx.y], (unsigned long long)b);
}

// launch 2: tile counts -> exclusive prefixes, the total at [n].  grid = (chunks of 4096 counts, columns); a CTA sums
// the counts in front of its chunk itself (coalesced reads of at most a few hundred KB that sit in L2: no dependency
// between CTAs, no second launch), then scans its own 4096 -- one CTA per column walking all 36 K counts of a
// 600 M-row column took 20 us.
struct NcPrefixParams {
  const uint32_t* sums[FLE_MAX_COLS];
  uint64_t* pref[FLE_MAX_COLS];
  int n;
};
constexpr int kPfPerThread = 4, kPfChunk = 1024 * kPfPerThread;

__global__ void __launch_bounds__(1024) nc_prefix_kernel(const __grid_constant__ NcPrefixParams P) {
  __shared__ uint64_t warp_tot[32];
  __shared__ uint64_t pre_s;
  const uint32_t* in = P.sums[blockIdx.y];
  uint64_t* out = P.pref[blockIdx.y];
  if (!in) return;
  const int n = P.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int base = (int)blockIdx.x * kPfChunk;
  uint64_t pre = 0;
  for (int i = tid; i < base; i += 1024 * 4) {  // (base is a multiple of 4096: four loads in flight per pass)
    const uint32_t a = in[i], b = in[i + 1024], c = in[i + 2048], d = in[i + 3072];
    pre += (uint64_t)a + b + c + d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, o);
  if (lane == 0) warp_tot[warp] = pre;
  __syncthreads();
  if (tid == 0) {
    uint64_t t = 0;
    for (int w = 0; w < 32; ++w) t += warp_tot[w];
    pre_s = t;
  }
  __syncthreads();
  const uint64_t before = pre_s;
  uint32_t v[kPfPerThread];
  uint32_t local = 0;
#pragma unroll
  for (int j = 0; j < kPfPerThread; ++j) {
    const int i = base + tid * kPfPerThread + j;
    v[j] = i < n ? in[i] : 0u;
    local += v[j];
  }
  uint32_t incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  __syncthreads();  // (warp_tot is reused)
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  uint64_t run = before + (incl - local);
  uint64_t total = before;
  for (int w = 0; w < 32; ++w) {
    const uint64_t t = warp_tot[w];
    if (w < warp) run += t;
    total += t;
  }
#pragma unroll
  for (int j = 0; j < kPfPerThread; ++j) {
    const int i = base + tid * kPfPerThread + j;
    if (i < n) out[i] = run;
    run += v[j];
  }
  if (tid == 0 && base + kPfChunk >= n) out[n] = total;  // the last chunk
}

// ------------------------------------------------------------------------------------------
// launch 3
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void lds64u(uint32_t addr, uint32_t& a, uint32_t& b) {
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr));
}

// one pass (K chains of one type on the column's planes) over a staged group; plane-order result word.  The
// literal masks (all-ones / all-zeros per plane and chain) come from a shared-memory table, [plane][chain], read
// with broadcast vector loads: two planes' worth per load for K <= 2.
template <bool IS_LT, int K>
__device__ __forceinline__ uint64_t nc_chains(const PassDesc& pd, uint32_t bw, uint32_t addr, uint32_t maddr) {
  uint32_t al[K], ah[K];
#pragma unroll
  for (int k = 0; k < K; ++k) al[k] = ah[k] = IS_LT ? (pd.init[k] ? ~0u : 0u) : ~0u;
  auto step = [&](uint64_t w, const uint32_t* mk) {
    const uint32_t wl = (uint32_t)w, wh = (uint32_t)(w >> 32);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (IS_LT) {
        al[k] = (~wl & mk[k]) | (~(wl ^ mk[k]) & al[k]);
        ah[k] = (~wh & mk[k]) | (~(wh ^ mk[k]) & ah[k]);
      } else {
        al[k] = al[k] & ~(wl ^ mk[k]);
        ah[k] = ah[k] & ~(wh ^ mk[k]);
      }
    }
  };
  auto masks2 = [&](uint32_t i, uint32_t (&mk)[2 * K]) {  // the masks of planes i and i + 1
    const uint32_t a = maddr + i * (K * 4u);
    if constexpr (K == 1) lds64u(a, mk[0], mk[1]);
    else if constexpr (K == 2) lds128u(a, mk[0], mk[1], mk[2], mk[3]);
    else {
      lds128u(a, mk[0], mk[1], mk[2], mk[3]);
      lds128u(a + 16, mk[4], mk[5], mk[6], mk[7]);
    }
  };
  uint32_t mk[2 * K];
  if (bw & 1u) {  // odd widths: 64-bit loads are conflict-free with the segment skew; even widths need 128-bit ones
    uint32_t i = 0;
#pragma unroll 2
    for (; i + 1 < bw; i += 2) {
      const uint64_t w0 = lds64(addr + i * 8u), w1 = lds64(addr + i * 8u + 8u);
      masks2(i, mk);
      step(w0, mk);
      step(w1, mk + K);
    }
    const uint64_t w0 = lds64(addr + i * 8u);
    masks2(i, mk);  // (the table holds 32 planes: the second half of the pair reads zeros)
    step(w0, mk);
  } else {
#pragma unroll 2
    for (uint32_t i = 0; i < bw; i += 2) {
      uint64_t w0, w1;
      lds128(addr + i * 8u, w0, w1);
      masks2(i, mk);
      step(w0, mk);
      step(w1, mk + K);
    }
  }
  uint32_t rl = al[0] ^ (pd.neg[0] ? ~0u : 0u), rh = ah[0] ^ (pd.neg[0] ? ~0u : 0u);
  const int mop = pd.fold[K - 1];
#pragma unroll
  for (int k = 1; k < K; ++k) {
    const uint32_t nm = pd.neg[k] ? ~0u : 0u;
    if (mop == FOLD_AND) {
      rl &= al[k] ^ nm;
      rh &= ah[k] ^ nm;
    } else {
      rl |= al[k] ^ nm;
      rh |= ah[k] ^ nm;
    }
  }
  return ((uint64_t)rh << 32) | rl;
}

// COMPARES: every pass of the chain is one or two compare chains (a compare, BETWEEN / NOT BETWEEN) -- the kernel then
// carries two of the six variants per column (a third of the code: the instruction cache holds the tile loop)
template <bool COMPARES>
__device__ __forceinline__ uint64_t nc_run_pass(const PassDesc& pd, uint32_t bw, uint32_t addr, uint32_t maddr) {
  if constexpr (COMPARES) {
    if (pd.k == 1) return nc_chains<true, 1>(pd, bw, addr, maddr);
    return nc_chains<true, 2>(pd, bw, addr, maddr);
  }
  if (pd.type == PASS_LT) {
    if (pd.k == 1) return nc_chains<true, 1>(pd, bw, addr, maddr);
    if (pd.k == 2) return nc_chains<true, 2>(pd, bw, addr, maddr);
    return nc_chains<true, 4>(pd, bw, addr, maddr);
  }
  if (pd.k == 1) return nc_chains<false, 1>(pd, bw, addr, maddr);
  if (pd.k == 2) return nc_chains<false, 2>(pd, bw, addr, maddr);
  return nc_chains<false, 4>(pd, bw, addr, maddr);
}

// 32-bit bit deposit: the i-th set bit of mask (from the LSB) receives bit i of src.  Few holes (the usual case: a
// couple of NULLs per 64 rows): a zero is opened at every clear position of mask, lowest first -- moving the bits
// from the hole upwards left by one is r += (r & -hole).  Few ones: bit by bit.
__device__ __forceinline__ uint32_t dep32_holes(uint32_t src, uint32_t mask) {
  uint32_t r = src, h = ~mask;
  while (h) {
    const uint32_t b = h & (0u - h);
    r += r & (0u - b);
    h ^= b;
  }
  return r & mask;
}
__device__ __forceinline__ uint32_t dep32_ones(uint32_t src, uint32_t mask) {
  uint32_t r = 0, m = mask;
  while (m) {
    const uint32_t b = m & (0u - m);
    r |= b & (0u - (src & 1u));
    src >>= 1;
    m ^= b;
  }
  return r;
}
// both halves of a row word; the usual case (few holes in both) runs them in one loop, two independent chains
__device__ __forceinline__ uint64_t dep2(uint32_t s0, uint32_t m0, uint32_t s1, uint32_t m1) {
  uint32_t r0, r1;
  if (__popc(m0) >= 16 && __popc(m1) >= 16) {
    uint32_t h0 = ~m0, h1 = ~m1;
    r0 = s0;
    r1 = s1;
    while (h0 | h1) {
      const uint32_t b0 = h0 & (0u - h0), b1 = h1 & (0u - h1);  // (no hole left on one side: b = 0, nothing moves)
      r0 += r0 & (0u - b0);
      r1 += r1 & (0u - b1);
      h0 ^= b0;
      h1 ^= b1;
    }
    r0 &= m0;
    r1 &= m1;
  } else {
    r0 = __popc(m0) >= 16 ? dep32_holes(s0, m0) : dep32_ones(s0, m0);
    r1 = __popc(m1) >= 16 ? dep32_holes(s1, m1) : dep32_ones(s1, m1);
  }
  return ((uint64_t)r1 << 32) | r0;
}

struct NcMeta {
  int64_t gfirst;  // first staged data group (even)
  int32_t lr0;     // rank of the tile's first row among the column's non-null rows, relative to gfirst * 64 (0..127)
  int32_t ng;      // staged groups that hold values of the tile
  int32_t left;    // 1: the last of them does not travel with the bulk copies (odd tail at the end of the buffer)
  int32_t pad;
};

// 8 consumer warps (one row word per thread and tile) + 1 producer warp (lane c owns column c: tile prefixes -> what
// the tile stages -> bulk copies, two tiles ahead of the consumers, so that no consumer ever waits for a prefix)
template <int NC, bool COMPARES>
__global__ void __launch_bounds__(kNcThreads + 32) null_chain_kernel(const __grid_constant__ NcParams P) {
  extern __shared__ __align__(128) uint8_t nc_smem[];
  __shared__ NcMeta meta[4][NC];
  __shared__ __align__(16) uint32_t warp_tot[2][NC][kNcThreads / 32];
  __shared__ __align__(16) uint32_t lit_masks[NC][32 * 4];  // [plane][chain], chains = the pass's k
  __shared__ uint64_t warp_buf[kNcThreads / 32];
  __shared__ __align__(8) uint64_t bars[kNcStages];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_producer_warp = warp == kNcThreads / 32;
  const uint32_t smem0 = smem_u32(nc_smem);
  const uint32_t bar0 = smem_u32(bars);
  const int64_t ntiles = P.ntiles;
  auto tile_of = [&](int64_t k) { return (int64_t)blockIdx.x + k * (int64_t)gridDim.x; };

  if (tid == 0) {
    for (int s = 0; s < kNcStages; ++s) mbar_init(bar0 + s * 8, (uint32_t)P.ncols);
    mbar_fence_init();
  }
  for (int idx = tid; idx < NC * 128; idx += kNcThreads + 32) {
    const int c = idx >> 7, e = idx & 127;
    uint32_t v = 0;
    if (c < P.ncols) {
      const int K = P.col[c].pd.k, i = e / K, k = e - i * K;
      if (i < 32) v = 0u - ((P.col[c].pd.lit[k] >> i) & 1u);
    }
    lit_masks[c][e] = v;
  }
  __syncthreads();

  if (is_producer_warp) {
    const bool producer = lane < P.ncols;
    const int c = producer ? lane : 0;
    const NcCol& C = P.col[c];
    const int64_t nsub = C.pref ? (int64_t)C.pref[ntiles] : P.n_root;
    const int64_t groups = min((nsub + 63) >> 6, C.max_groups);
    auto make_meta = [&](int64_t tile, uint64_t p0, uint64_t p1) {
      int64_t boff, btot;
      if (C.pref) {
        boff = (int64_t)p0;
        btot = (int64_t)(p1 - p0);
      } else {
        boff = tile * (int64_t)(kNcWords * 64);
        btot = min((int64_t)(kNcWords * 64), P.n_root - boff);
      }
      NcMeta M;
      M.gfirst = (boff >> 6) & ~1ll;
      M.lr0 = (int32_t)(boff - M.gfirst * 64);
      const int64_t gend = min(groups, (boff + btot + 63) >> 6);
      M.ng = btot > 0 ? (int)max((int64_t)0, gend - M.gfirst) : 0;
      // bulk copies move 16-byte multiples: an odd number of groups of an odd width takes one more group along when
      // the buffer holds it, else the last group travels by hand
      M.left = ((M.ng & 1) && (C.bw & 1u) && M.gfirst + M.ng + 1 > C.max_groups) ? 1 : 0;
      M.pad = 0;
      return M;
    };
    auto issue = [&](const NcMeta& M, int stage) {
      const uint32_t bar = bar0 + stage * 8;
      int nb = M.ng;
      if ((nb & 1) && (C.bw & 1u)) nb += M.left ? -1 : 1;
      if (nb <= 0) {
        mbar_arrive(bar);
        return;
      }
      const uint32_t gbytes = C.bw * 8u;
      mbar_arrive_expect_tx(bar, (uint32_t)nb * gbytes);
      const uint8_t* src = reinterpret_cast<const uint8_t*>(C.planes) + (uint64_t)M.gfirst * gbytes;
      const uint32_t dst = smem0 + stage * P.stage_bytes + C.col_off;
      for (int s = 0; s < 9; ++s) {  // (segment 8 = the overflow segment, right behind the eighth)
        const int g0 = s * kNcSeg;
        if (g0 >= nb) break;
        const int cnt = min(kNcSeg, nb - g0);
        tma_load_1d(dst + s * C.seg_stride, src + (uint64_t)g0 * gbytes, (uint32_t)cnt * gbytes, bar);
      }
    };
    auto load_pref = [&](int64_t tile, uint64_t& p0, uint64_t& p1) {
      p0 = p1 = 0;
      if (producer && tile < ntiles && C.pref) {
        p0 = C.pref[tile];
        p1 = C.pref[tile + 1];
      }
    };
    uint64_t pa, pb;  // prefixes of tile k + 2, loaded one iteration before they are used
    if (producer) {
      uint64_t q0, q1;
      load_pref(tile_of(0), q0, q1);
      meta[0][c] = make_meta(tile_of(0), q0, q1);
      load_pref(tile_of(1), q0, q1);
      meta[1][c] = make_meta(tile_of(1), q0, q1);
      issue(meta[0][c], 0);
    }
    load_pref(tile_of(2), pa, pb);
    __syncthreads();  // (prologue barrier of the consumers)
    for (int64_t k = 0; tile_of(k) < ntiles; ++k) {
      // the next tile into the other stage: its readers passed the barrier after phase C of the previous iteration
      if (producer) {
        if (tile_of(k + 1) < ntiles) issue(meta[(k + 1) & 3][c], (int)((k + 1) & 1));
        meta[(k + 2) & 3][c] = make_meta(tile_of(k + 2), pa, pb);
      }
      load_pref(tile_of(k + 3), pa, pb);
      __syncthreads();
      __syncthreads();
    }
    return;
  }

  // ---- consumers ----
  const uint32_t seg = tid & 7, jrow = tid >> 3;
  const uint32_t lm0 = smem_u32(&lit_masks[0][0]);
  auto load_raw = [&](int64_t tile, uint64_t (&raw)[NC]) {  // plane words of the def levels, as they lie in memory
    const int64_t w = tile * kNcWords + tid;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      raw[c] = ~0ull;
      if (c < P.ncols && P.col[c].def_planes) raw[c] = (tile < ntiles && w < P.nwords) ? P.col[c].def_planes[w] : 0ull;
    }
  };
  auto to_masks = [&](int64_t tile, const uint64_t (&raw)[NC], uint64_t (&m)[NC]) {  // -> not-null words, row order
    uint64_t keep = ~0ull;
    if (tile >= ntiles - 1) keep = tile < ntiles ? tail_mask_word(tile * kNcWords + tid, P.n_root) : 0ull;
#pragma unroll
    for (int c = 0; c < NC; ++c) m[c] = c < P.ncols ? (__brevll(raw[c]) & keep) : 0ull;
  };
  uint64_t m[NC], raw[NC];
  load_raw(tile_of(0), raw);
  to_masks(tile_of(0), raw, m);
  __syncthreads();

  uint64_t cnt = 0;
  for (int64_t k = 0; tile_of(k) < ntiles; ++k) {
    const int cur = (int)(k & 1);
    const int64_t tile = tile_of(k), tile1 = tile_of(k + 1);
    // A. the next tile's def words on their way (nothing looks at them before the end of the iteration)
    load_raw(tile1, raw);
    if (warp == 1) {  // odd tail group of a column that ends with the buffer (at most once per column and launch)
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        if (c < P.ncols && meta[k & 3][c].left) {
          const NcCol& C = P.col[c];
          const NcMeta& M = meta[k & 3][c];
          const int q = M.ng - 1;
          if ((uint32_t)lane < C.bw)
            sts64(smem0 + cur * P.stage_bytes + C.col_off + (q >> 5) * C.seg_stride + (q & 31) * C.bw * 8u + lane * 8u,
                  C.planes[(M.gfirst + q) * C.bw + lane]);
        }
      }
    }
    // B. warp scans of the not-null popcounts, the warps' totals to shared memory
    uint32_t excl[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const uint32_t mine = __popcll(m[c]);
      uint32_t incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      if (lane == 31) warp_tot[cur][c][warp] = incl;
      excl[c] = incl - mine;
    }
    mbar_wait(bar0 + cur * 8, (uint32_t)(k >> 1) & 1u);
    __syncthreads();
    // C. every staged group is evaluated once: thread (seg, j) takes group seg * kNcSeg + j of every column
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (c < P.ncols) {
        const NcCol& C = P.col[c];
        const int ng = meta[k & 3][c].ng;
        const uint32_t gbytes = C.bw * 8u;
        const uint32_t base = smem0 + cur * P.stage_bytes + C.col_off + seg * C.seg_stride;
        const uint32_t res = smem0 + C.res_off;
        const uint32_t maddr = lm0 + c * (32 * 4 * 4);
        {
          const int q = (int)(seg * kNcSeg + jrow);
          if (q < ng) sts64(res + q * 8u, __brevll(nc_run_pass<COMPARES>(C.pd, C.bw, base + jrow * gbytes, maddr)));
        }
        if (ng > 8 * kNcSeg && tid < ng - 8 * kNcSeg)  // overflow segment: a tile with (nearly) no NULL at all
          sts64(res + (8 * kNcSeg + tid) * 8u,
                __brevll(nc_run_pass<COMPARES>(C.pd, C.bw, smem0 + cur * P.stage_bytes + C.col_off + 8 * C.seg_stride + tid * gbytes, maddr)));
        // the warps' totals -> exclusive prefixes, in place (warp c for column c; phase D reads one word per thread)
        if (warp == c) {
          const uint32_t wa = smem_u32(&warp_tot[cur][c][0]) + (lane & 7) * 4u;
          const uint32_t v = lds32(wa);
          uint32_t incl = v;
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if ((lane & 7) >= o) incl += y;
          }
          if (lane < 8) asm volatile("st.shared.b32 [%0], %1;" ::"r"(wa), "r"(incl - v) : "memory");
        }
      }
    }
    __syncthreads();
    // D. deposit at the row word's rank (32-bit halves of the result stream), fold the columns, write the word once
    {
      const int64_t w = tile * kNcWords + tid;
      uint64_t r = P.join_op == FOLD_AND ? ~0ull : 0ull;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        if (c < P.ncols) {
          const uint32_t lr = (uint32_t)meta[k & 3][c].lr0 + excl[c] + warp_tot[cur][c][warp];
          const uint32_t res = smem0 + P.col[c].res_off;
          const uint32_t m0 = (uint32_t)m[c], m1 = (uint32_t)(m[c] >> 32);
          const uint32_t lr1 = lr + __popc(m0);
          const uint32_t a0 = res + (lr >> 5) * 4u, a1 = res + (lr1 >> 5) * 4u;
          const uint32_t s0 = __funnelshift_r(lds32(a0), lds32(a0 + 4), lr & 31u);
          const uint32_t s1 = __funnelshift_r(lds32(a1), lds32(a1 + 4), lr1 & 31u);
          const uint64_t dep = m[c] == ~0ull ? (((uint64_t)s1 << 32) | s0) : dep2(s0, m0, s1, m1);
          r = P.join_op == FOLD_AND ? (r & dep) : (r | dep);
        }
      }
      if (w < P.nwords) {
        P.out[w] = r;
        cnt += __popcll(r);
      }
    }
    to_masks(tile1, raw, m);
  }
  if (P.count) {
    // (the producer warp has left: a named barrier over the consumer warps)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) warp_buf[warp] = cnt;
    asm volatile("bar.sync 1, %0;" ::"n"(kNcThreads) : "memory");
    if (tid == 0) {
      uint64_t t = 0;
#pragma unroll
      for (int x = 0; x < kNcThreads / 32; ++x) t += warp_buf[x];
      if (t) atomicAdd(P.count, (unsigned long long)t);
    }
  }
}

static const void* null_chain_kernel_for(int nc, bool compares) {
  switch (nc) {
    case 1: return compares ? (const void*)null_chain_kernel<1, true> : (const void*)null_chain_kernel<1, false>;
    case 2: return compares ? (const void*)null_chain_kernel<2, true> : (const void*)null_chain_kernel<2, false>;
    case 3: return compares ? (const void*)null_chain_kernel<3, true> : (const void*)null_chain_kernel<3, false>;
    case 4: return compares ? (const void*)null_chain_kernel<4, true> : (const void*)null_chain_kernel<4, false>;
  }
  return nullptr;
}

int64_t null_chain_tiles(int64_t n) { return ((n + 63) / 64 + kNcWords - 1) / kNcWords; }

// dynamic shared memory of the chain kernel for these bit widths (two stages of every column + the result words)
static size_t null_chain_smem(const int* bw, int ncols) {
  size_t stage = 0;
  for (int i = 0; i < ncols; ++i) {
    const size_t seg_stride = ((size_t)kNcSeg * bw[i] * 8 + 127) / 128 * 128 + 16;
    stage += (8 * seg_stride + (size_t)kNcOver * bw[i] * 8 + 127) / 128 * 128;
  }
  return (size_t)kNcStages * stage + (size_t)ncols * kNcGroups * 8;
}

// false: the columns are too wide for the kernel's two stages (the caller keeps to the per-column kernels)
bool null_chain_fits(fle_ctx* ctx, const int* bw, int ncols) {
  return ncols >= 1 && ncols <= kNcMaxCols && null_chain_smem(bw, ncols) + 4096 <= (size_t)ctx->max_smem_optin;
}

// the sub-program [nodes, nodes + nnodes) on column `col` as ONE pass, or FLE_E_UNSUPPORTED
int null_chain_pass(fle_ctx* ctx, const fle_col_desc& col, const fle_pred_node* nodes, int nnodes, PassDesc* out) {
  if (col.bit_width < 1 || col.bit_width > 32) return FLE_E_UNSUPPORTED;
  std::vector<fle_pred_node> sub(nodes, nodes + nnodes);
  for (auto& nd : sub)
    if (nd.op != FLE_OP_AND && nd.op != FLE_OP_OR) nd.col = 0;
  fle_col_desc dc{col.planes, col.bytes, col.bit_width, 0, nullptr, 0};
  LoweredProgram lp;
  std::string err;
  if (lower_program(&dc, 1, sub.data(), nnodes, &lp, &err, ctx->reference_compat) != FLE_OK) return FLE_E_UNSUPPORTED;
  const PassDesc& p0 = lp.p.passes[0];
  const bool single_pass = lp.p.npasses == 1 && lp.p.ncombs == 0 && (p0.type == PASS_LT || p0.type == PASS_EQ) &&
                           (p0.k == 1 || p0.merge) && p0.slot[0] == 0 && (p0.k == 1 || p0.k == 2 || p0.k == 4);
  if (!single_pass) return FLE_E_UNSUPPORTED;
  *out = p0;
  return FLE_OK;
}

int launch_def_tile_counts(fle_ctx* ctx, const fle_col_desc* cols, int ncols, int64_t n, uint64_t* counts_dev,
                           uint32_t* tile_sums_dev, size_t sums_stride) {
  if (ncols < 1 || ncols > FLE_MAX_COLS) return set_error(ctx, FLE_E_INVALID, "bad column count");
  DefTileParams P;
  P.nwords = (n + 63) / 64;
  P.nrows = n;
  P.ntiles = null_chain_tiles(n);
  P.counts = reinterpret_cast<unsigned long long*>(counts_dev);
  for (int c = 0; c < FLE_MAX_COLS; ++c) {
    P.planes[c] = nullptr;
    P.tile_sums[c] = nullptr;
  }
  for (int c = 0; c < ncols; ++c) {
    if (cols[c].max_def_level == 0) continue;
    if (!cols[c].def_planes) return set_error(ctx, FLE_E_INVALID, "nullable column without def_planes");
    if (reinterpret_cast<uintptr_t>(cols[c].def_planes) & 7) return set_error(ctx, FLE_E_ALIGN, "def_planes must be 8-byte aligned");
    if (cols[c].def_bytes < (uint64_t)P.nwords * 8u) return set_error(ctx, FLE_E_BUFFER, "def-level stream shorter than the rows requested");
    P.planes[c] = reinterpret_cast<const uint64_t*>(cols[c].def_planes);
    P.tile_sums[c] = tile_sums_dev + (size_t)c * sums_stride;
  }
  FLE_CUDA_TRY(ctx, cudaMemsetAsync(counts_dev, 0, (size_t)ncols * 8, ctx->stream));
  if (P.ntiles == 0) return FLE_OK;
  const int gx = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * 4, (P.ntiles + 7) / 8));
  def_tile_counts_kernel<<<dim3(gx, ncols), 256, 0, ctx->stream>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "def_tile_counts launch");
  ctx->launches += 1;
  return FLE_OK;
}

// prefix + chain kernel; tile_sums_dev as written by launch_def_tile_counts (indexed by the column's position in cols)
int launch_null_chain(fle_ctx* ctx, const NullChainPlan& plan, const fle_col_desc* cols, int64_t n, const uint32_t* tile_sums_dev,
                      uint64_t* pref_dev, size_t sums_stride, uint64_t* out_dev, uint64_t* count_dev) {
  bool compares = true;
  for (int i = 0; i < plan.ncols; ++i) compares = compares && plan.pd[i].type == PASS_LT && (plan.pd[i].k == 1 || plan.pd[i].k == 2);
  const void* kern = null_chain_kernel_for(plan.ncols, compares);
  if (!kern || n <= 0) return FLE_E_UNSUPPORTED;
  NcParams P;
  memset(&P, 0, sizeof(P));
  P.ncols = plan.ncols;
  P.join_op = plan.join_op;
  P.n_root = n;
  P.nwords = (n + 63) / 64;
  P.ntiles = null_chain_tiles(n);
  P.out = out_dev;
  P.count = reinterpret_cast<unsigned long long*>(count_dev);
  NcPrefixParams Q;
  memset(&Q, 0, sizeof(Q));
  Q.n = (int)P.ntiles;
  uint32_t off = 0;
  bool any_pref = false;
  for (int i = 0; i < plan.ncols; ++i) {
    const fle_col_desc& col = cols[plan.col[i]];
    NcCol& C = P.col[i];
    C.planes = reinterpret_cast<const uint64_t*>(col.planes);
    C.bw = (uint32_t)col.bit_width;
    C.max_groups = (int64_t)(col.bytes / ((uint64_t)col.bit_width * 8u));
    C.pd = plan.pd[i];
    if (col.max_def_level != 0) {
      C.def_planes = reinterpret_cast<const uint64_t*>(col.def_planes);
      C.pref = pref_dev + (size_t)plan.col[i] * sums_stride;
      Q.sums[i] = tile_sums_dev + (size_t)plan.col[i] * sums_stride;
      Q.pref[i] = pref_dev + (size_t)plan.col[i] * sums_stride;
      any_pref = true;
    }
    C.col_off = off;
    C.seg_stride = ((uint32_t)kNcSeg * C.bw * 8u + 127u) / 128u * 128u + 16u;
    off += (8u * C.seg_stride + (uint32_t)kNcOver * C.bw * 8u + 127u) / 128u * 128u;
  }
  P.stage_bytes = off;
  uint32_t roff = (uint32_t)kNcStages * P.stage_bytes;
  for (int i = 0; i < plan.ncols; ++i) {
    P.col[i].res_off = roff;
    roff += (uint32_t)kNcGroups * 8u;
  }
  const size_t smem = roff;
  if (smem > (size_t)ctx->max_smem_optin - 4096) return FLE_E_UNSUPPORTED;
  if (smem > 48 * 1024) {  // opt in once per kernel and device (process-wide)
    static std::mutex mu;
    static std::set<std::pair<int, const void*> > raised;
    std::lock_guard<std::mutex> lock(mu);
    const std::pair<int, const void*> key(ctx->device, kern);
    if (raised.find(key) == raised.end()) {
      cudaError_t ea = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin - 4096);
      if (ea != cudaSuccess) return set_cuda_error(ctx, ea, "cudaFuncSetAttribute(null_chain)");
      raised.insert(key);
    }
  }
  if (count_dev) FLE_CUDA_TRY(ctx, cudaMemsetAsync(count_dev, 0, 8, ctx->stream));
  if (any_pref) {
    nc_prefix_kernel<<<dim3((unsigned)((Q.n + kPfChunk - 1) / kPfChunk), (unsigned)plan.ncols), 1024, 0, ctx->stream>>>(Q);
    ctx->launches += 1;
  }
  const int per_sm = std::max(1, std::min(2, (int)((size_t)ctx->max_smem_per_sm / (smem + 2048))));
  int grid = (int)std::min<int64_t>(P.ntiles, (int64_t)ctx->num_sms * per_sm);
  if (ctx->null_chain_ctas > 0) grid = std::min(grid, ctx->null_chain_ctas);
  void* args[] = {(void*)&P};
  cudaError_t e = cudaLaunchKernel(kern, dim3(grid), dim3(kNcThreads + 32), args, smem, ctx->stream);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_cuda_error(ctx, e, "null_chain launch");
  }
  ctx->launches += 1;
  return FLE_OK;
}

}  // namespace fle
