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
  __device__ __forceinline__ void init(uint32_t table_addr, uint32_t lane) { table = table_addr + (lane & (uint32_t)(IT::kCopies - 1)) * 4u; }

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
  __device__ __forceinline__ void init(uint32_t table_addr, uint32_t lane) { table = table_addr + lane * 4u; }

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

// ------------------------------------------------------------------------------------------
// Third form (default, in_wide = 2): THREE instructions per lookup, and a transpose no larger than the code.
//
// ncu on the entry-table kernel above (profiles/r02_in_wide_entry_ncu.txt): 594 issued instructions per
// group whatever the bit width (the butterflies are inline PTX, so zero planes do not fold away; bw 8 costs
// what bw 12 costs), issue slots 82 % busy -- the kernel is bound by instruction issue, and every lookup pays
// four of them: extract the half, IMAD the lane's table base onto it, load, IMAD-append.  Here the word that
// comes out of the transpose IS the shared-memory offset of the row's table entry, lane bits included, so a
// lookup is   ALU: extract (PRMT / LOP3 / SHF)   LSU: ld.shared [offset + constant table base]   FMA: IMAD
// that drops the entry at the row's bit position of the result word (acc += entry * 2^row): the selection
// word is assembled directly in row order, no epilogue.
//
//   bw <= 11  (InWordDecode): the low 8 planes go through 8 x 8 SWAR transposes (3 stages on 8 words per 32
//             rows: 96 instructions per group instead of 224); a transposed byte is the row's low code byte,
//             and ONE PRMT builds  offset = code << 8 | lane << 2  from it and the lane constant.  Entries are
//             32-bit words; planes 8, 9 and 10 stay bit-sliced: an entry carries the answers for all their values
//             (2 x 16 bits or 4 x 8 bits apart, or 8 bits side by side), the IMADs sort them into per-case row
//             words and the untouched planes select among those bitwise (1 / 3 LOP3 per 32 rows; for three planes
//             a second 8 x 8 SWAR transpose sorts the answers, select8).
//   bw 12     (InByteDecode): the 16 x 16 butterflies above, with the five LANE bits fed in as constant planes
//             (all-ones or all-zeros words per thread) at the positions the table layout wants them, so a
//             transposed half is  code[10:2] << 7 | lane << 2 | code[1:0]  = the offset of a BYTE entry in a
//             lane-private 64 KB table; plane 11 stays bit-sliced (answers in bits 0 and 4 of the byte,
//             rows dropped four to a byte, nibble select).
//   bw 13..16 (InByte3Decode): the same, three planes bit-sliced (eight answers per entry byte, select8).
// Set membership of the same codes, so bit-identical to the OR of equalities by construction.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kInTableOff = 64;  // the table's offset in dynamic shared memory (launch_program: right behind the mbarriers)

__device__ __forceinline__ uint32_t bitsel_r(uint32_t a, uint32_t b, uint32_t m) {  // (a & ~m) | (b & m)
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(r) : "r"(a), "r"(b), "r"(m));
  return r;
}
// acc + e * MUL as one IMAD (FMA pipe; written in C a power of two becomes a shift-add on the busier ALU pipe)
template <uint32_t MUL>
__device__ __forceinline__ uint32_t drop_at(uint32_t acc, uint32_t e) {
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(e), "n"(MUL), "r"(acc));
  return r;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// (the 8 x 8 SWAR transpose, swar8_transpose, lives in fle_device.cuh: the narrow unpack kernels use it too)
__device__ __forceinline__ void in_transpose8(uint32_t (&S)[8]) { swar8_transpose(S); }

// Eight answers per entry (three planes kept bit-sliced): bit c of a row's entry byte answers for value c of those
// planes.  The entries of a 32-row half are dropped byte-wise into eight words, E[k] byte b = entry of row
// 8 b + 7 - k, which is exactly the input the 8 x 8 SWAR transpose wants: afterwards E[7 - c] is the row-order
// word of answers for value c, and the three untouched planes pick among the eight words bitwise (7 LOP3).
__device__ __forceinline__ uint32_t select8(uint32_t (&E)[8], uint32_t x0, uint32_t x1, uint32_t x2) {
  swar8_transpose(E);
  const uint32_t a = bitsel_r(bitsel_r(E[7], E[6], x0), bitsel_r(E[5], E[4], x0), x1);
  const uint32_t b = bitsel_r(bitsel_r(E[3], E[2], x0), bitsel_r(E[1], E[0], x0), x1);
  return bitsel_r(a, b, x2);
}

template <int BW>
struct InWord {
  static constexpr int kTB = BW < 8 ? BW : 8;  // transposed code bits
  static constexpr int kNX = BW - kTB;         // planes kept bit-sliced (0 .. 3)
  static constexpr uint32_t kTableBytes = (1u << kTB) * 256u;  // entry of code e, lane l at e * 256 + l * 4
  static constexpr int kCaseBits = kNX == 0 ? 32 : (kNX == 1 ? 16 : (kNX == 2 ? 8 : 1));  // distance between the answers inside an entry
};

template <int BW>
struct InWordDecode {
  using IW = InWord<BW>;
  static_assert(IW::kNX <= 3, "bw <= 11");
  uint32_t table;  // shared address of the table
  uint32_t lane4;  // lane * 4: byte 0 of every entry offset

  __device__ __forceinline__ void init(uint32_t table_addr, uint32_t lane) {
    table = table_addr;
    lane4 = lane * 4u;
  }

  struct Planes {
    uint32_t S0[8], S1[8];
    uint64_t X[IW::kNX > 0 ? IW::kNX : 1];
  };

  template <class L>
  __device__ __forceinline__ uint64_t run(const L& ld) const {
    Planes pl;
    load(ld, pl);
    return compute(pl);
  }

  template <class L>
  __device__ __forceinline__ void load(const L& ld, Planes& pl) const {
#pragma unroll
    for (int j = 0; j < 8; ++j) pl.S0[j] = pl.S1[j] = 0u;
#pragma unroll
    for (int i = 0; i < BW; i += 2) {
      uint64_t a, b;
      ld.template load_pair<BW>(i, a, b);
      if (i < IW::kTB) {
        pl.S0[7 - (i < IW::kTB ? i : 0)] = (uint32_t)(a >> 32);
        pl.S1[7 - (i < IW::kTB ? i : 0)] = (uint32_t)a;
      } else {
        pl.X[i >= IW::kTB ? i - IW::kTB : 0] = a;
      }
      if (i + 1 < BW) {
        if (i + 1 < IW::kTB) {
          pl.S0[7 - (i + 1 < IW::kTB ? i + 1 : 0)] = (uint32_t)(b >> 32);
          pl.S1[7 - (i + 1 < IW::kTB ? i + 1 : 0)] = (uint32_t)b;
        } else {
          pl.X[i + 1 >= IW::kTB ? i + 1 - IW::kTB : 0] = b;
        }
      }
    }
  }

  // entry of the half's row R (0..31): its code byte is byte 3 - R / 8 of word R % 8
  template <int R>
  __device__ __forceinline__ uint32_t entry(const uint32_t (&S)[8]) const {
    constexpr uint32_t sel = 0x5504u | ((uint32_t)(3 - R / 8) << 4);  // [0][0][code byte][lane * 4]
    return lds32(table + __byte_perm(S[R % 8], lane4, sel));
  }

  // rows R0 .. R0 + N - 1 of a half dropped at bits 0 .. N - 1 of an accumulator
  // (two independent IMAD chains per accumulator: a single chain of 32 dependent IMADs shows up as 'wait' stalls)
  template <int R0, int N>
  __device__ __forceinline__ uint32_t rows(const uint32_t (&S)[8]) const {
    uint32_t acc0 = 0, acc1 = 0;
    rows_step<R0, N, 0>(S, acc0);
    rows_step<R0, N, 1>(S, acc1);
    return acc0 + acc1;
  }
  template <int R0, int N, int I>
  __device__ __forceinline__ void rows_step(const uint32_t (&S)[8], uint32_t& acc) const {
    if constexpr (I < N) {
      acc = drop_at<(1u << I)>(acc, entry<R0 + I>(S));
      rows_step<R0, N, I + 2>(S, acc);
    }
  }

  // 32 rows of one half -> selection word in row order
  // E[k] byte b = entry of row 8 b + 7 - k (select8)
  template <int K>
  __device__ __forceinline__ void drop_bytes(const uint32_t (&S)[8], uint32_t (&E)[8]) const {
    if constexpr (K < 8) {
      uint32_t acc = entry<7 - K>(S);
      acc = drop_at<(1u << 8)>(acc, entry<8 + 7 - K>(S));
      acc = drop_at<(1u << 16)>(acc, entry<16 + 7 - K>(S));
      E[K] = drop_at<(1u << 24)>(acc, entry<24 + 7 - K>(S));
      drop_bytes<K + 1>(S, E);
    }
  }

  __device__ __forceinline__ uint32_t half(const uint32_t (&S)[8], const uint32_t (&x)[3]) const {
    if constexpr (IW::kNX == 3) {
      uint32_t E[8];
      drop_bytes<0>(S, E);
      return select8(E, x[0], x[1], x[2]);
    } else if constexpr (IW::kNX == 0) {
      return rows<0, 32>(S);
    } else if constexpr (IW::kNX == 1) {
      // an accumulator takes 16 rows: answers for plane 8 = 0 in its low half, for plane 8 = 1 in its high half
      const uint32_t a = rows<0, 16>(S), b = rows<16, 16>(S);
      return bitsel_r(__byte_perm(a, b, 0x5410), __byte_perm(a, b, 0x7632), x[0]);
    } else {
      // an accumulator takes 8 rows: byte c = answers for (plane 9, plane 8) = c
      const uint32_t a0 = rows<0, 8>(S), a1 = rows<8, 8>(S), a2 = rows<16, 8>(S), a3 = rows<24, 8>(S);
      uint32_t m[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t sel = ((uint32_t)(4 + c) << 4) | (uint32_t)c;  // [.][.][second.byte c][first.byte c]
        m[c] = __byte_perm(__byte_perm(a0, a1, sel), __byte_perm(a2, a3, sel), 0x5410);
      }
      return bitsel_r(bitsel_r(m[0], m[1], x[0]), bitsel_r(m[2], m[3], x[0]), x[1]);
    }
  }

  __device__ __forceinline__ uint64_t compute(Planes& pl) const {
    in_transpose8(pl.S0);
    in_transpose8(pl.S1);
    uint32_t xlo[3] = {0u, 0u, 0u}, xhi[3] = {0u, 0u, 0u};
#pragma unroll
    for (int i = 0; i < IW::kNX; ++i) {
      const uint64_t xr = __brevll(pl.X[i]);  // plane order (row k at bit 63 - k) -> row order
      xlo[i] = (uint32_t)xr;
      xhi[i] = (uint32_t)(xr >> 32);
    }
    const uint32_t lo = half(pl.S0, xlo), hi = half(pl.S1, xhi);
    return ((uint64_t)hi << 32) | (uint64_t)lo;
  }
};

template <int BW>
struct InByte {
  static constexpr int kTB = BW < 11 ? BW : 11;
  static constexpr int kNX = BW - kTB;  // 0 or 1
  static constexpr uint32_t kTableBytes = kTB <= 2 ? 128u : (1u << (kTB - 2)) * 128u;  // entry e, lane l at (e >> 2) * 128 + l * 4 + (e & 3)
  // bit position of code bit i / of lane bit j inside a transposed 16-bit half (= the entry's byte offset)
  FLE_HD static constexpr int pos(int i) { return i < 2 ? i : i + 5; }
  FLE_HD static constexpr int lane_pos(int j) { return 2 + j; }
};

template <int BW>
struct InByteDecode {
  using IB = InByte<BW>;
  static_assert(IB::kNX <= 1, "bw <= 12");
  uint32_t table;
  uint32_t lane_plane[5];  // lane bit j as a constant plane: all ones or all zeros

  __device__ __forceinline__ void init(uint32_t table_addr, uint32_t lane) {
    table = table_addr;
#pragma unroll
    for (int j = 0; j < 5; ++j) lane_plane[j] = 0u - ((lane >> j) & 1u);
  }

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
#pragma unroll
    for (int j = 0; j < 16; ++j) pl.S0[j] = pl.S1[j] = 0u;
    pl.X = 0;
#pragma unroll
    for (int j = 0; j < 5; ++j) pl.S0[15 - IB::lane_pos(j)] = pl.S1[15 - IB::lane_pos(j)] = lane_plane[j];
#pragma unroll
    for (int i = 0; i < BW; i += 2) {
      uint64_t a, b;
      ld.template load_pair<BW>(i, a, b);
      if (i < IB::kTB) {
        pl.S0[15 - IB::pos(i < IB::kTB ? i : 0)] = (uint32_t)(a >> 32);
        pl.S1[15 - IB::pos(i < IB::kTB ? i : 0)] = (uint32_t)a;
      } else {
        pl.X = a;
      }
      if (i + 1 < BW) {
        if (i + 1 < IB::kTB) {
          pl.S0[15 - IB::pos(i + 1 < IB::kTB ? i + 1 : 0)] = (uint32_t)(b >> 32);
          pl.S1[15 - IB::pos(i + 1 < IB::kTB ? i + 1 : 0)] = (uint32_t)b;
        } else {
          pl.X = b;
        }
      }
    }
  }

  // after in_transpose: S[k] = offset(row k) << 16 | offset(row 16 + k)
  // (measured and dropped, tools/in_sweep.py: taking the halves apart on the idle FMA pipe -- hi = IMAD.HI(w, 2^16),
  //  lo = IMAD(hi, -2^16, w) -- +1 %; the butterflies' right shifts as IMAD.HI: 3-9 % slower, as in round 1)
  template <int R>
  __device__ __forceinline__ uint32_t entry(const uint32_t (&S)[16]) const {
    const uint32_t w = S[R % 16];
    return lds_u8(table + (R < 16 ? (w >> 16) : (w & 0xFFFFu)));
  }

  // bit position of row I (0..15) of a nibble-packed accumulator: four rows to a byte, in its low nibble
  FLE_HD static constexpr uint32_t nib_pos(int i) { return (uint32_t)(8 * (i / 4) + (i % 4)); }

  template <int R0, int N, int I, bool NIB>
  __device__ __forceinline__ void rows_step(const uint32_t (&S)[16], uint32_t& acc) const {
    if constexpr (I < N) {
      acc = drop_at<(1u << (NIB ? nib_pos(I) : (uint32_t)I))>(acc, entry<R0 + I>(S));
      rows_step<R0, N, I + 2, NIB>(S, acc);
    }
  }
  template <int R0, int N, bool NIB>
  __device__ __forceinline__ uint32_t rows(const uint32_t (&S)[16]) const {  // two independent IMAD chains
    uint32_t acc0 = 0, acc1 = 0;
    rows_step<R0, N, 0, NIB>(S, acc0);
    rows_step<R0, N, 1, NIB>(S, acc1);
    return acc0 + acc1;
  }

  // 16 rows whose entries carry two answers (bit 0: plane 11 = 0, bit 4: plane 11 = 1), chosen by the 16 bits
  // of that plane in byte pair BP (0: bytes 0, 1; 1: bytes 2, 3) of xr -> 16 result bits in bytes 0 and 2
  template <int R0, int BP>
  __device__ __forceinline__ uint32_t rows16_sel(const uint32_t (&S)[16], uint32_t xr) const {
    const uint32_t acc = rows<R0, 16, true>(S);
    const uint32_t y1 = __byte_perm(xr, 0u, BP == 0 ? 0x4140 : 0x4342);  // [0][x byte 1][0][x byte 0] of the pair
    const uint32_t y = (y1 | (y1 << 4)) & 0x0F0F0F0Fu;                   // nibble n of the 16 bits -> low nibble of byte n
    const uint32_t t = bitsel_r(acc, acc >> 4, y);                       // low nibbles: the selected answers (high: don't care)
    return bitsel<0xF0F0F0F0u>(t, t >> 4);                               // bytes 0 and 2: [nibble 1 | nibble 0], [nibble 3 | nibble 2]
  }

  __device__ __forceinline__ uint32_t half(const uint32_t (&S)[16], uint32_t xr) const {
    if constexpr (IB::kNX == 0) {
      return rows<0, 32, false>(S);
    } else {
      return __byte_perm(rows16_sel<0, 0>(S, xr), rows16_sel<16, 1>(S, xr), 0x6420);
    }
  }

  __device__ __forceinline__ uint64_t compute(Planes& pl) const {
    in_transpose(pl.S0);
    in_transpose(pl.S1);
    const uint64_t xr = __brevll(pl.X);
    const uint32_t lo = half(pl.S0, (uint32_t)xr), hi = half(pl.S1, (uint32_t)(xr >> 32));
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

// bw 13..16: byte entries with EIGHT answers (the top three planes stay bit-sliced, select8), the other bw - 3 code bits
// and the copy number go through the 16 x 16 butterflies as in InByteDecode.  A half has room for 16 - (bw - 3) copy
// bits: 32 lane-private copies up to bw 14, 16 / 8 copies for bw 15 / 16 (two / four lanes share a copy: 1.5 / 2.1
// wavefronts per lookup); 64 KB at most.  Against the replicated bit table (InDecode<BW, true>: mask, add, load,
// variable shift, funnel append = 5.5 instructions per row, 650 per group, ALU pipe 440) a row costs extract, load,
// IMAD, plus 110 per group for the select: 575 per group, ALU pipe 365.
template <int BW>
struct InByte3 {
  static constexpr int kTB = BW - 3;
  static constexpr int kCB = (16 - kTB) < 5 ? (16 - kTB) : 5;  // copy bits
  static constexpr uint32_t kCopies = 1u << kCB;
  static constexpr uint32_t kTableBytes = (1u << kTB) * kCopies;  // entry e, copy c at ((e >> 2) * copies + c) * 4 + (e & 3)
  FLE_HD static constexpr int pos(int i) { return i < 2 ? i : i + kCB; }
  FLE_HD static constexpr int copy_pos(int j) { return 2 + j; }
};

template <int BW>
struct InByte3Decode {
  using IB = InByte3<BW>;
  static_assert(BW >= 13 && BW <= 16, "bw 13..16");
  uint32_t table;
  uint32_t copy_plane[IB::kCB];  // copy-number bit j (= lane bit j) as a constant plane

  __device__ __forceinline__ void init(uint32_t table_addr, uint32_t lane) {
    table = table_addr;
#pragma unroll
    for (int j = 0; j < IB::kCB; ++j) copy_plane[j] = 0u - ((lane >> j) & 1u);
  }

  struct Planes {
    uint32_t S0[16], S1[16];
    uint64_t X[3];
  };

  template <class L>
  __device__ __forceinline__ uint64_t run(const L& ld) const {
    Planes pl;
    load(ld, pl);
    return compute(pl);
  }

  template <class L>
  __device__ __forceinline__ void load(const L& ld, Planes& pl) const {
#pragma unroll
    for (int j = 0; j < 16; ++j) pl.S0[j] = pl.S1[j] = 0u;
#pragma unroll
    for (int j = 0; j < IB::kCB; ++j) pl.S0[15 - IB::copy_pos(j)] = pl.S1[15 - IB::copy_pos(j)] = copy_plane[j];
#pragma unroll
    for (int i = 0; i < BW; i += 2) {
      uint64_t a, b;
      ld.template load_pair<BW>(i, a, b);
      if (i < IB::kTB) {
        pl.S0[15 - IB::pos(i < IB::kTB ? i : 0)] = (uint32_t)(a >> 32);
        pl.S1[15 - IB::pos(i < IB::kTB ? i : 0)] = (uint32_t)a;
      } else {
        pl.X[i >= IB::kTB ? i - IB::kTB : 0] = a;
      }
      if (i + 1 < BW) {
        if (i + 1 < IB::kTB) {
          pl.S0[15 - IB::pos(i + 1 < IB::kTB ? i + 1 : 0)] = (uint32_t)(b >> 32);
          pl.S1[15 - IB::pos(i + 1 < IB::kTB ? i + 1 : 0)] = (uint32_t)b;
        } else {
          pl.X[i + 1 >= IB::kTB ? i + 1 - IB::kTB : 0] = b;
        }
      }
    }
  }

  // after in_transpose: S[k] = offset(row k) << 16 | offset(row 16 + k)
  template <int R>
  __device__ __forceinline__ uint32_t entry(const uint32_t (&S)[16]) const {
    const uint32_t w = S[R % 16];
    return lds_u8(table + (R < 16 ? (w >> 16) : (w & 0xFFFFu)));
  }
  template <int K>
  __device__ __forceinline__ void drop_bytes(const uint32_t (&S)[16], uint32_t (&E)[8]) const {
    if constexpr (K < 8) {
      uint32_t acc = entry<7 - K>(S);
      acc = drop_at<(1u << 8)>(acc, entry<8 + 7 - K>(S));
      acc = drop_at<(1u << 16)>(acc, entry<16 + 7 - K>(S));
      E[K] = drop_at<(1u << 24)>(acc, entry<24 + 7 - K>(S));
      drop_bytes<K + 1>(S, E);
    }
  }

  __device__ __forceinline__ uint64_t compute(Planes& pl) const {
    in_transpose(pl.S0);
    in_transpose(pl.S1);
    uint32_t xlo[3], xhi[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const uint64_t xr = __brevll(pl.X[i]);  // plane order (row k at bit 63 - k) -> row order
      xlo[i] = (uint32_t)xr;
      xhi[i] = (uint32_t)(xr >> 32);
    }
    uint32_t E[8];
    drop_bytes<0>(pl.S0, E);
    const uint32_t lo = select8(E, xlo[0], xlo[1], xlo[2]);
    drop_bytes<0>(pl.S1, E);
    const uint32_t hi = select8(E, xhi[0], xhi[1], xhi[2]);
    return ((uint64_t)hi << 32) | (uint64_t)lo;
  }
};

template <int BW>
struct WideByte3Policy {  // in_wide = 2, bw 13..16
  using Decode = InByte3Decode<BW>;
  using IB = InByte3<BW>;
  static constexpr uint32_t kTableBytes = IB::kTableBytes;
  static constexpr uint32_t kCopies = 1;  // (the decoder adds the copy number itself)
  static constexpr uint32_t kItemsPerLit = IB::kCopies;
  __device__ static __forceinline__ void put(uint32_t* table, uint32_t c, uint32_t copy) {
    const uint32_t e = c & ((1u << IB::kTB) - 1u), x = c >> IB::kTB;
    atomicOr(&table[(e >> 2) * IB::kCopies + copy], 1u << ((e & 3u) * 8u + x));
  }
};

template <int BW>
struct WideWordPolicy {  // in_wide = 2, bw <= 10: 32-bit entries addressed by one PRMT
  using Decode = InWordDecode<BW>;
  using IW = InWord<BW>;
  static constexpr uint32_t kTableBytes = IW::kTableBytes;
  static constexpr uint32_t kCopies = 1;  // (the decoder adds the lane itself)
  static constexpr uint32_t kItemsPerLit = 32;
  __device__ static __forceinline__ void put(uint32_t* table, uint32_t c, uint32_t lane) {
    const uint32_t e = c & ((1u << IW::kTB) - 1u), x = c >> IW::kTB;
    atomicOr(&table[e * 64u + lane], 1u << (x * (uint32_t)(IW::kCaseBits & 31)));
  }
};

template <int BW>
struct WideBytePolicy {  // in_wide = 2, bw 11 / 12: byte entries, lane bits transposed in as constant planes
  using Decode = InByteDecode<BW>;
  using IB = InByte<BW>;
  static constexpr uint32_t kTableBytes = IB::kTableBytes;
  static constexpr uint32_t kCopies = 1;
  static constexpr uint32_t kItemsPerLit = 32;
  __device__ static __forceinline__ void put(uint32_t* table, uint32_t c, uint32_t lane) {
    const uint32_t e = c & ((1u << IB::kTB) - 1u), x = c >> IB::kTB;
    atomicOr(&table[(e >> 2) * 32u + lane], 1u << ((e & 3u) * 8u + x * 4u));
  }
};

// 16 consumer warps + one producer warp.  There is no CTA-wide barrier in the tile loop: a consumer
// warp waits for the stage's "full" mbarrier, copies its groups' planes into registers, arrives on
// the stage's "empty" mbarrier and only then evaluates; the producer refills a stage as soon as all
// 16 warps have released it.  Warps therefore drift apart by up to a tile, which keeps some of them
// in the ALU-only transposes while others are in the FMA / LSU-heavy lookups (bw 12, 600 M rows:
// 229 -> 202 us; the same change made the two-CTA bit-table kernel above 7 % slower).
// NCW consumer warps, GPT groups per consumer thread per tile (tile = NCW * 32 * GPT groups).
template <int BW, class POL, int NCW = 16, int GPT = 1>
__global__ void __launch_bounds__(NCW * 32 + 32, 1) eval_in_wide_kernel(const __grid_constant__ EvalParams P) {
  constexpr int NC = NCW * 32, NT = NC + 32;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t full0 = smem_u32(smem + P.smem_bar_off);  // full[s] at + 8 s, empty[s] at + 32 + 8 s (nstages <= 4)
  const uint32_t empty0 = full0 + 32u;
  const uint32_t stage0 = smem_u32(smem + P.smem_stage_off);
  uint32_t* table_ptr = reinterpret_cast<uint32_t*>(smem + kInTableOff);  // (a compile-time address: the lookups' immediate)
  typename POL::Decode dec;
  dec.init(smem_u32(table_ptr), threadIdx.x & 31u);
  constexpr int T = NC * GPT;
  constexpr int seg_groups = T >> 3;
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
    constexpr uint32_t kGroupStride = (uint32_t)(NC / 8) * BW * 8u;  // between a thread's groups inside its segment
    const int64_t last_g = P.ngroups - 1;
    const uint64_t tail_mask = (P.nrows & 63) ? ((1ull << (P.nrows & 63)) - 1ull) : ~0ull;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      mbar_wait(full0 + stage * 8, parity);
      typename POL::Decode::Planes pl[GPT];
#pragma unroll
      for (int q = 0; q < GPT; ++q) dec.load(InSmemGroup{stage0 + stage * P.stage_bytes + thread_off + (uint32_t)q * kGroupStride}, pl[q]);
      __syncwarp();
      if ((threadIdx.x & 31u) == 0) mbar_arrive(empty0 + stage * 8);  // this warp is done with the stage
#pragma unroll
      for (int q = 0; q < GPT; ++q) {
        const int64_t g = tile * T + thread_g + q * (NC / 8);
        const uint64_t other = comb_load(P, g);  // (issued before the lookups: its latency hides behind them)
        uint64_t r = dec.compute(pl[q]) ^ negmask;
        if (g == last_g) r &= tail_mask;  // (only the table's last group can hold rows past the end)
        r = comb_apply(P, r, other);
        P.bitmap[g] = r;
        cnt_bits += __popcll(r);
      }
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

// Launch shape of the form-2 kernels (consumer warps, groups per consumer thread per tile), measured at 600 M rows with
// L = 64 / 256 (tools/in_sweep.py --opt in_wide_warps=.. --opt in_wide_gpt=..; profiles/r02_in_sweep_variants.txt):
// two groups per thread amortise the tile loop's fixed instructions where the registers allow it (bw 8-10: +1-2 %),
// 24 warps hide more latency where they do not (bw 7, 12); 31 warps: unstable between runs (+3 % / -8 %).
void eval_in_wide_shape(int bw, int* ncw, int* gpt) {
  *ncw = (bw == 7 || bw == 11 || bw == 12) ? 24 : 16;  // (bw 13..16: the 64 KB table leaves room for 16-warp tiles only)
  *gpt = (bw >= 8 && bw <= 10) ? 2 : 1;
}

const void* eval_in_wide_kernel_for(int bw, int form, int ncw, int gpt) {
  if (form >= 2) {
#define V(B, POL, W, Q) if (bw == B && ncw == W && gpt == Q) return (const void*)eval_in_wide_kernel<B, POL<B>, W, Q>;
#define VS(B, POL) V(B, POL, 16, 1) V(B, POL, 16, 2) V(B, POL, 24, 1)
    VS(7, WideWordPolicy) VS(8, WideWordPolicy) VS(9, WideWordPolicy) VS(10, WideWordPolicy) VS(11, WideWordPolicy)
    VS(12, WideBytePolicy)
    VS(13, WideByte3Policy) VS(14, WideByte3Policy) VS(15, WideByte3Policy) VS(16, WideByte3Policy)
#undef VS
#undef V
    if (ncw != 16 || gpt != 1) return nullptr;
    switch (bw) {
#define X(B) case B: return (const void*)eval_in_wide_kernel<B, WideWordPolicy<B> >;
      X(1) X(2) X(3) X(4) X(5) X(6)
#undef X
    }
  }
  if (ncw != 16 || gpt != 1) return nullptr;
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

uint32_t eval_in_wide_table_bytes(int bw, int form) {
  if (form >= 2 && bw <= 11) return (1u << std::min(bw, 8)) * 256u;          // InWord<bw>::kTableBytes
  if (form >= 2 && bw == 12) return (1u << 9) * 128u;                         // InByte<12>::kTableBytes
  if (form >= 2 && bw <= 16) return (1u << (bw - 3)) * (1u << std::min(16 - (bw - 3), 5));  // InByte3<bw>::kTableBytes
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
