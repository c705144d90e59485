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
//   2. transposes each group's planes into its 64 codes in registers (16 x 16-block butterflies,
//      two codes per 32-bit word),
//   3. looks every code up in the table and assembles the 64-bit selection word directly in
//      row order,
// which costs the same for any L.  Same TMA/mbarrier tile pipeline and skewed shared layout as
// fle_eval1.cu.  Results are bit-identical to the OR-of-equalities by construction (set
// membership of the same bw-bit codes).
#include <algorithm>

#include "fle_eval_common.cuh"

namespace fle {

// The kernel is bound by the ALU pipe (LOP3 / SHF / PRMT: 64 lanes per clock per SM), not by HBM:
// ncu on the first version showed 88 % ALU-pipe utilisation (574 ALU instructions per group) with
// the FMA pipe (IMAD) at 20 %.  The group function is therefore written for few ALU instructions
// (407 now, 600 instructions in all against 903):
//   * the two 32-row halves are transposed as 16 x 16 blocks (stages 8, 4, 2, 1 only: a
//     transposed word then holds the codes of rows k and 16 + k instead of k and 32 + k, which
//     the result assembly absorbs) -- the 16-bit stage (32 PRMT) disappears;
//   * a butterfly stage is two selects,  a' = (a & ~M) | ((b >> J) & M),  b' = (b & M) | ((a << J) & ~M),
//     each ONE LOP3 (inline PTX: written in C the compiler spends a LOP3 per constant), the left
//     shift an IMAD.SHL on the FMA pipe;
//   * plane i is fed to the butterfly at the bit position the lookup wants it: code bits 0..4
//     stay at 0..4 (SHF uses them as the shift amount without masking), bits 5.. land at 7..
//     for the lane-private table, whose word w sits at byte w * 128 + lane * 4, so the word
//     address is one LOP3 (half & mask) plus an IMAD add of the lane's base;
//   * table words are stored bit-reversed: `word << (code & 31)` brings the wanted bit to bit 31
//     and the funnel shift  acc = (acc : x) << 1  appends it.
// Tried and dropped (profiles/r01_in_lookup_tuning.txt): right shifts as IMAD.HI (x >> J ==
// umulhi(x, 2^(32-J))) to unload the ALU pipe -- IMAD.HI issues at a quarter of the LOP3 rate
// (tools/micro/int_pipes.cu) and shows up as dispatch stalls, every mix was slower; 2 stages /
// 3-4 CTAs per SM instead of 4 stages / 2 CTAs -- no change, the pipe is the limit.
// REP (the one-CTA-per-SM kernel at the end of this file, bw 13..16): the table is replicated
// 32 / 16 / 8 times for bw <= 14 / 15 / 16 (64 KB at most), copy c of word w at byte
// (w * copies + c) * 4, and lane l reads copy l mod copies: conflict-free up to bw 14, two or four
// lanes per copy beyond (1.5 / 2.1 wavefronts per lookup instead of ~3.5 with one shared copy).
// The word index then sits at bits kShift.. of the half, which is where pos() puts the code bits.
template <int BW, bool REP = false>
struct InTable {
  static constexpr bool kPerLane = REP || BW <= 13;  // 2^13 bits x 32 lanes = 32 KB
  static constexpr int kWords = BW <= 5 ? 1 : (1 << (BW - 5));
  static constexpr int kCopies = !REP ? (BW <= 13 ? 32 : 1) : (BW <= 14 ? 32 : (BW == 15 ? 16 : 8));
  // log2 of the byte stride between consecutive table words
  static constexpr int kShift = kCopies == 32 ? 7 : (kCopies == 16 ? 6 : (kCopies == 8 ? 5 : 2));
  static constexpr uint32_t kAddrMask = (uint32_t)(kWords - 1) << kShift;
  static constexpr uint32_t kTableBytes = (uint32_t)kWords * 4u * (uint32_t)kCopies;
  // bit position of code bit i inside its 16-bit half of a transposed word
  FLE_HD static constexpr int pos(int i) { return (kPerLane && i >= 5) ? i + (kShift - 5) : i; }
};

// (a & ~M) | (b & M) as ONE LOP3 (written in C the compiler spends a LOP3 per constant, M and ~M)
template <uint32_t M>
__device__ __forceinline__ uint32_t bitsel(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(r) : "r"(a), "r"(b), "n"(M));
  return r;
}

template <int J, uint32_t M>
__device__ __forceinline__ void in_stage(uint32_t (&S)[16]) {
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    if ((k & J) == 0) {
      const uint32_t a = S[k], b = S[k + J];
      if (J == 8) {
        S[k] = __byte_perm(a, b, 0x3715);      // (a & 0xFF00FF00) | ((b >> 8) & 0x00FF00FF)
        S[k + J] = __byte_perm(a, b, 0x2604);  // (b & 0x00FF00FF) | ((a & 0x00FF00FF) << 8)
      } else {
        S[k] = bitsel<M>(a, b >> J);
        S[k + J] = bitsel<M>(a << J, b);
      }
    }
  }
}

// 16 x 16 bit transposes of the low and of the high halves of S[0..15] at once:
// S'[k] bit i == S[15 - i] bit (15 - k),   S'[k] bit 16 + i == S[15 - i] bit (31 - k)
__device__ __forceinline__ void in_transpose(uint32_t (&S)[16]) {
  in_stage<8, 0x00FF00FFu>(S);
  in_stage<4, 0x0F0F0F0Fu>(S);
  in_stage<2, 0x33333333u>(S);
  in_stage<1, 0x55555555u>(S);
}

template <int BW, bool REP = false>
struct InDecode {
  using IT = InTable<BW, REP>;
  uint32_t table;  // shared address of this lane's table (of the shared one for bw > 13)

  // the group's planes, in registers: bit planes of rows 0..31 / 32..63, plane i at word 15 - pos(i)
  struct Planes {
    uint32_t S0[16], S1[16];
  };

  // membership of the code in bits 0..15 of h (higher bits ignored) -> bit 31 of the result
  __device__ __forceinline__ uint32_t probe(uint32_t h) const {
    const uint32_t off = IT::kPerLane ? (h & IT::kAddrMask) : ((h >> 3) & IT::kAddrMask);
    return lds32(table + off) << (h & 31u);
  }

  template <class L>
  __device__ __forceinline__ uint64_t run(const L& ld) const {
    Planes pl;
    load(ld, pl);
    return compute(pl);
  }

  template <class L>
  __device__ __forceinline__ void load(const L& ld, Planes& pl) const {
    uint32_t (&S0)[16] = pl.S0;
    uint32_t (&S1)[16] = pl.S1;
#pragma unroll
    for (int j = 0; j < 16; ++j) S0[j] = S1[j] = 0u;
#pragma unroll
    for (int i = 0; i < BW; i += 2) {
      uint64_t a, b;
      ld.template load_pair<BW>(i, a, b);
      S0[15 - IT::pos(i)] = (uint32_t)(a >> 32);
      S1[15 - IT::pos(i)] = (uint32_t)a;
      if (i + 1 < BW) {
        S0[15 - IT::pos(i + 1 < BW ? i + 1 : i)] = (uint32_t)(b >> 32);
        S1[15 - IT::pos(i + 1 < BW ? i + 1 : i)] = (uint32_t)b;
      }
    }
  }

  __device__ __forceinline__ uint64_t compute(Planes& pl) const {
    uint32_t (&S0)[16] = pl.S0;
    uint32_t (&S1)[16] = pl.S1;
    in_transpose(S0);  // S0[k] = half(row k) << 16 | half(row 16 + k)
    in_transpose(S1);  // S1[k] = half(row 32 + k) << 16 | half(row 48 + k)
    uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0;  // rows 0..15, 16..31, 32..47, 48..63 (bit k = row base + k)
#pragma unroll
    for (int k = 15; k >= 0; --k) {
      const uint32_t a = S0[k], b = S1[k];
      r1 = __funnelshift_l(probe(a), r1, 1);
      r0 = __funnelshift_l(probe(a >> 16), r0, 1);
      r3 = __funnelshift_l(probe(b), r3, 1);
      r2 = __funnelshift_l(probe(b >> 16), r2, 1);
    }
    return ((uint64_t)(r2 | (r3 << 16)) << 32) | (uint64_t)(r0 | (r1 << 16));
  }
};

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
  InDecode<BW> dec;
  dec.table = IT::kPerLane ? smem_u32(rep_ptr) + (threadIdx.x & 31u) * 4u : smem_u32(table_ptr);
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
    atomicOr(&table_ptr[c >> 5], 0x80000000u >> (c & 31));  // bit-reversed words, see probe()
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
      const uint64_t other = comb_load(P, g);
      const uint64_t r = comb_apply(P, finish_row_order(dec.run(InSmemGroup{a}), negmask, g, P.nrows), other);
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
    const uint64_t other = comb_load(P, g);
    const uint64_t r = comb_apply(P, finish_row_order(dec.run(InGlobalGroup{P.planes[0] + g * BW}), negmask, g, P.nrows), other);
    P.bitmap[g] = r;
    cnt_bits += __popcll(r);
  }
  publish_count<NT>(P, cnt_bits);
}

// ------------------------------------------------------------------------------------------
// Entry-table kernel (bw <= 12): one ALU instruction per lookup instead of 3.5.
//
// The bit table above costs, per row, a mask for the word address, a variable shift to bring the
// row's bit to bit 31 and a funnel shift to append it -- all on the ALU pipe that already carries
// the transposes, and that pipe is what bounds the kernel.  Here every code owns a 16-bit table
// ENTRY (lane-private: entry e of lane l at byte (e >> 1) * 128 + l * 4 + (e & 1) * 2, 2^11 entries
// = 128 KB), so a lookup is
//     ALU: h = word >> 16  or  word & mask      (the half IS the entry address / 2)
//     FMA: addr = h * 2 + lane_base             (IMAD)
//     LSU: e = ld.shared.u16 [addr]
//     FMA: acc = acc * 2 + e                    (IMAD appends the row's bit)
// A 16-bit half holds 11 address bits (bit 0, then bits 6..15: the five in between are the lane),
// so at most 11 code bits are transposed.  For bw 12 the top plane is NOT transposed: the entry of
// the low 11 bits carries both answers, bit 0 = member(code) and bit 8 = member(code + 2^11); an
// accumulator takes 8 rows (answers in bytes 0 and 1), two PRMTs gather the bytes of four
// accumulators into two 32-row words M0 / M1 and the untouched plane selects between them
// bitwise:  r = (M0 & ~X) | (M1 & X)  -- one LOP3 per 32 rows.
// One CTA per SM (the table leaves 99 KB for two 48 KB stages); set membership of the same codes, so
// bit-identical to the OR of equalities by construction.
template <int BW>
struct InWide {
  static constexpr int kTB = BW < 11 ? BW : 11;  // transposed code bits
  static constexpr int kNX = BW - kTB;           // planes kept bit-sliced (0 or 1)
  static constexpr int kEntries = 1 << kTB;
  static constexpr uint32_t kTableBytes = (uint32_t)(kEntries < 2 ? 2 : kEntries) * 64u;
  FLE_HD static constexpr int pos(int i) { return i == 0 ? 0 : i + 5; }
  static constexpr uint32_t kHalfMask = 1u | ((((1u << (kTB - 1)) - 1u) << 6) & 0xFFFFu);
};

__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
  uint32_t v;
  asm volatile("{ .reg .u16 t; ld.shared.u16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(v) : "r"(addr));
  return v;
}

template <int BW>
struct InWideDecode {
  using IW = InWide<BW>;
  uint32_t table;  // shared address of this lane's column of the table

  // (inline PTX: from C the compiler folds the * 2 into the half's extraction -- shift by 15, then
  //  a mask to clear bit 0 -- and spends a second ALU instruction per lookup)
  __device__ __forceinline__ uint32_t entry(uint32_t h) const {
    uint32_t addr;
    asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(addr) : "r"(h), "r"(table));
    return lds16(addr);
  }

  // the group's planes, in registers: bit planes of rows 0..31 / 32..63 (plane i at word
  // 15 - pos(i)) and the plane that stays bit-sliced (bw 12)
  struct Planes {
    uint32_t S0[16], S1[16];
    uint64_t X;
  };

  template <class L>
  __device__ __forceinline__ uint64_t run(const L& ld) const {
    Planes pl;
    load(ld, pl);
    return compute(pl);
  }

  template <class L>
  __device__ __forceinline__ void load(const L& ld, Planes& pl) const {
    uint32_t (&S0)[16] = pl.S0;
    uint32_t (&S1)[16] = pl.S1;
    uint64_t& X = pl.X;
#pragma unroll
    for (int j = 0; j < 16; ++j) S0[j] = S1[j] = 0u;
    X = 0;
#pragma unroll
    for (int i = 0; i < BW; i += 2) {
      uint64_t a, b;
      ld.template load_pair<BW>(i, a, b);
      if (i < IW::kTB) {
        S0[15 - IW::pos(i < IW::kTB ? i : 0)] = (uint32_t)(a >> 32);
        S1[15 - IW::pos(i < IW::kTB ? i : 0)] = (uint32_t)a;
      } else {
        X = a;
      }
      if (i + 1 < BW) {
        if (i + 1 < IW::kTB) {
          S0[15 - IW::pos(i + 1 < IW::kTB ? i + 1 : 0)] = (uint32_t)(b >> 32);
          S1[15 - IW::pos(i + 1 < IW::kTB ? i + 1 : 0)] = (uint32_t)b;
        } else {
          X = b;
        }
      }
    }
  }

  __device__ __forceinline__ uint64_t compute(Planes& pl) const {
    uint32_t (&S0)[16] = pl.S0;
    uint32_t (&S1)[16] = pl.S1;
    const uint64_t X = pl.X;
    in_transpose(S0);  // S0[k] = half(row k) << 16 | half(row 16 + k)
    in_transpose(S1);  // S1[k] = half(row 32 + k) << 16 | half(row 48 + k)
    uint32_t lo, hi;
    if constexpr (IW::kNX == 0) {
      uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0;  // rows 0..15, 16..31, 32..47, 48..63
#pragma unroll
      for (int k = 15; k >= 0; --k) {
        const uint32_t a = S0[k], b = S1[k];
        r0 = r0 * 2u + entry(a >> 16);
        r1 = r1 * 2u + entry(a & IW::kHalfMask);
        r2 = r2 * 2u + entry(b >> 16);
        r3 = r3 * 2u + entry(b & IW::kHalfMask);
      }
      lo = r1 * 65536u + r0;
      hi = r3 * 65536u + r2;
    } else {
      // eight rows per accumulator: byte 0 = answers for X = 0, byte 1 = answers for X = 1
      uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
#pragma unroll
      for (int k = 7; k >= 0; --k) {
        const uint32_t u = S0[k], v = S0[k + 8], w = S1[k], z = S1[k + 8];
        a0 = a0 * 2u + entry(u >> 16);            // rows 0..7
        a1 = a1 * 2u + entry(v >> 16);            // rows 8..15
        a2 = a2 * 2u + entry(u & IW::kHalfMask);  // rows 16..23
        a3 = a3 * 2u + entry(v & IW::kHalfMask);  // rows 24..31
        b0 = b0 * 2u + entry(w >> 16);            // rows 32..39
        b1 = b1 * 2u + entry(z >> 16);
        b2 = b2 * 2u + entry(w & IW::kHalfMask);
        b3 = b3 * 2u + entry(z & IW::kHalfMask);
      }
      const uint32_t ca = a1 * 65536u + a0, cb = a3 * 65536u + a2;  // [M1 r8-15][M0 r8-15][M1 r0-7][M0 r0-7], ...
      const uint32_t cc = b1 * 65536u + b0, cd = b3 * 65536u + b2;
      const uint64_t xr = __brevll(X);  // plane order (row k at bit 63 - k) -> row order
      const uint32_t xlo = (uint32_t)xr, xhi = (uint32_t)(xr >> 32);
      lo = (__byte_perm(ca, cb, 0x6420) & ~xlo) | (__byte_perm(ca, cb, 0x7531) & xlo);
      hi = (__byte_perm(cc, cd, 0x6420) & ~xhi) | (__byte_perm(cc, cd, 0x7531) & xhi);
    }
    return ((uint64_t)hi << 32) | (uint64_t)lo;
  }
};

// One CTA of 512 threads per SM.  Tried and dropped (tools/in_wide_probe.py, 600 M rows, bw 12):
// consumer groups of 128 / 256 threads with their own stage rings and named barriers, so that warps
// drift apart (264 / 245 us against 228 us for one ring); half a group per thread in CTAs of 1024
// threads at 64 registers (285 us).  ncu: issue slots 71 % busy, ALU pipe 64 %; a group costs 592
// issued instructions (268 ALU, 195 FMA, 70 LSU), so the kernel is bound by instruction issue as
// much as by the ALU pipe.
// The two table forms behind one interface for the kernel below.
template <int BW>
struct WideEntryPolicy {  // bw <= 12: 16-bit entries, lane-private
  using Decode = InWideDecode<BW>;
  static constexpr uint32_t kTableBytes = InWide<BW>::kTableBytes;
  static constexpr uint32_t kCopies = 32;
  static constexpr uint32_t kItemsPerLit = 32;
  // item i of the table build: literal i / 32, lane i % 32
  __device__ static __forceinline__ void put(uint32_t* table, uint32_t c, uint32_t lane) {
    using IW = InWide<BW>;
    const uint32_t e = c & (uint32_t)(IW::kEntries - 1), x = c >> IW::kTB;
    atomicOr(&table[(e >> 1) * 32u + lane], 1u << ((e & 1u) * 16u + x * 8u));
  }
};

template <int BW>
struct WideBitsPolicy {  // bw 13..16: bit table, 32 / 16 / 8 copies
  using Decode = InDecode<BW, true>;
  using IT = InTable<BW, true>;
  static constexpr uint32_t kTableBytes = IT::kTableBytes;
  static constexpr uint32_t kCopies = IT::kCopies;
  static constexpr uint32_t kItemsPerLit = IT::kCopies;
  __device__ static __forceinline__ void put(uint32_t* table, uint32_t c, uint32_t copy) {
    atomicOr(&table[(c >> 5) * kCopies + copy], 0x80000000u >> (c & 31));  // bit-reversed words, see probe()
  }
};

// 16 consumer warps + one producer warp.  There is no CTA-wide barrier in the tile loop: a consumer
// warp waits for the stage's "full" mbarrier, copies its groups' planes into registers, arrives on
// the stage's "empty" mbarrier and only then evaluates; the producer refills a stage as soon as all
// 16 warps have released it.  Warps therefore drift apart by up to a tile, which keeps some of them
// in the ALU-only transposes while others are in the FMA / LSU-heavy lookups (bw 12, 600 M rows:
// 229 -> 202 us; the same change made the two-CTA bit-table kernel above 7 % slower).
constexpr int kInWideConsumers = 512;
constexpr int kInWideThreads = kInWideConsumers + 32;

template <int BW, class POL>
__global__ void __launch_bounds__(kInWideThreads, 1) eval_in_wide_kernel(const __grid_constant__ EvalParams P) {
  constexpr int NT = kInWideThreads, NC = kInWideConsumers;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t full0 = smem_u32(smem + P.smem_bar_off);  // full[s] at + 8 s, empty[s] at + 32 + 8 s (nstages <= 4)
  const uint32_t empty0 = full0 + 32u;
  const uint32_t stage0 = smem_u32(smem + P.smem_stage_off);
  uint32_t* table_ptr = reinterpret_cast<uint32_t*>(smem + P.smem_masks_off);
  typename POL::Decode dec;
  dec.table = smem_u32(table_ptr) + (threadIdx.x & (POL::kCopies - 1u)) * 4u;
  const int T = NC;  // one group per consumer thread per tile
  const int seg_groups = T >> 3;
  const int64_t ntiles = P.ngroups / T;
  const int nst = P.nstages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) {
      mbar_init(full0 + s * 8, 1);
      mbar_init(empty0 + s * 8, NC / 32);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < (int)(POL::kTableBytes / 4u); i += NT) table_ptr[i] = 0;
  __syncthreads();
  const PassDesc& pd = P.passes[0];
  const int off = (int)pd.lit[0], cnt = (int)pd.lit[1];
  for (int i = threadIdx.x; i < cnt * (int)POL::kItemsPerLit; i += NT)  // literals already masked to BW bits by the host
    POL::put(table_ptr, P.in_lits[off + i / (int)POL::kItemsPerLit], (uint32_t)i % POL::kItemsPerLit);
  __syncthreads();
  uint64_t cnt_bits = 0;

  if (threadIdx.x >= NC) {
    // ---- producer warp: one lane feeds the ring
    if (threadIdx.x == NC) {
      int stage = 0;
      uint32_t parity = 1;  // (a fresh mbarrier passes a wait on the phase before its first)
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        mbar_wait(empty0 + stage * 8, parity);
        issue_tile(P, full0, stage0, stage, tile, T);
        if (++stage == nst) {
          stage = 0;
          parity ^= 1u;
        }
      }
    }
  } else {
    // ---- consumer warps
    const uint64_t negmask = pd.neg[0] ? ~0ull : 0ull;
    const uint32_t seg = threadIdx.x & 7, jrow = threadIdx.x >> 3;
    const uint32_t thread_off = P.col_off[0] + seg * P.seg_stride[0] + jrow * (uint32_t)(BW * 8);
    const int64_t thread_g = (int64_t)seg * seg_groups + jrow;
    int stage = 0;
    uint32_t parity = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      mbar_wait(full0 + stage * 8, parity);
      typename POL::Decode::Planes pl;
      dec.load(InSmemGroup{stage0 + stage * P.stage_bytes + thread_off}, pl);
      __syncwarp();
      if ((threadIdx.x & 31u) == 0) mbar_arrive(empty0 + stage * 8);  // this warp is done with the stage
      const int64_t g = tile * T + thread_g;
      const uint64_t other = comb_load(P, g);  // (issued before the lookups: its latency hides behind them)
      const uint64_t r = comb_apply(P, finish_row_order(dec.compute(pl), negmask, g, P.nrows), other);
      P.bitmap[g] = r;
      cnt_bits += __popcll(r);
      if (++stage == nst) {
        stage = 0;
        parity ^= 1u;
      }
    }
    for (int64_t g = ntiles * T + (int64_t)blockIdx.x * NC + threadIdx.x; g < P.ngroups;
         g += (int64_t)gridDim.x * NC) {
      const uint64_t other = comb_load(P, g);
      const uint64_t r = comb_apply(P, finish_row_order(dec.run(InGlobalGroup{P.planes[0] + g * BW}), negmask, g, P.nrows), other);
      P.bitmap[g] = r;
      cnt_bits += __popcll(r);
    }
  }
  publish_count<NT>(P, cnt_bits);
}

const void* eval_in_wide_kernel_for(int bw) {
  switch (bw) {
#define X(B) case B: return (const void*)eval_in_wide_kernel<B, WideEntryPolicy<B> >;
    X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12)
#undef X
#define X(B) case B: return (const void*)eval_in_wide_kernel<B, WideBitsPolicy<B> >;
    X(13) X(14) X(15) X(16)
#undef X
  }
  return nullptr;
}

uint32_t eval_in_wide_table_bytes(int bw) {
  if (bw > 12) return (4u << (bw - 5)) * (bw <= 14 ? 32u : (bw == 15 ? 16u : 8u));  // InTable<bw, true>::kTableBytes
  const int tb = bw < 11 ? bw : 11;
  return (uint32_t)std::max(2, 1 << tb) * 64u;
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
