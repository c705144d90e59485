// Plain-data description of a lowered predicate program: the kernel parameter block of every
// fused-eval kernel.  Pure C-style structs so that the SAME text compiles under nvcc (library
// build) and under NVRTC (programs specialised at run time, fle_jit.cu embeds this file).
#pragma once
#ifdef __CUDACC_RTC__
#ifndef FLE_RTC_TYPES
#define FLE_RTC_TYPES
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef int int32_t;
typedef unsigned long long uint64_t;
typedef long long int64_t;
#endif
#else
#include <cstddef>
#include <cstdint>
#endif

#ifndef FLE_MAX_COLS
#define FLE_MAX_COLS 8 /* == include/fle_b200.h */
#endif
#ifndef FLE_MAX_GROUP
#define FLE_MAX_GROUP 8 /* == include/fle_b200.h: GPUs of one box */
#endif

namespace fle {

// ------------------------------------------------------------------------------------------
// Lowered form of a predicate program (built on the host by lower_program, consumed by the
// eval kernels as a __grid_constant__ parameter).
//
// Every comparison leaf becomes one "borrow chain" over the planes, LSB -> MSB:
//     acc' = (~W & C) | (~(W ^ C) & acc)          (one LOP3 per 32 rows per plane)
// which is the borrow-out of x - c (init 0 => x < c) or of x - c - 1 (init ~0 => x <= c).
//     Lt = chain(init 0)   Le = chain(init 1)   Ge = ~Lt   Gt = ~Le
// These are exact boolean functions of the row's bits, hence bit-identical to the
// reference's MSB -> LSB Mlt/Mgt/Meq recurrences (fle-encoding.h:8039-8042, :8151-8154).
// Eq / In use the reference's own XNOR-AND chain (fle-encoding.h:7986-7988, :8283-8290).
//
// Chains are grouped into passes of up to 4 chains on one column, so each plane word is
// fetched from shared memory once per pass.  Results fold into "slots": slot 0 lives in a
// register, slots >= 1 in per-thread shared-memory words; a short list of combine ops finishes
// the tree.  A program that is ONE pass (single compare, BETWEEN, Eq, <= 4 compares on one
// column) runs on the per-bit-width kernels of fle_eval1.cu, a single IN list on a column of
// <= 16 bits on the decode + lookup kernel of fle_eval_in.cu, everything else on the generic
// kernels of fle_eval.cu.
// ------------------------------------------------------------------------------------------
enum PassType : uint8_t { PASS_LT = 0, PASS_EQ = 1, PASS_IN = 2, PASS_CONST = 3 };
enum FoldOp : uint8_t { FOLD_SET = 0, FOLD_AND = 1, FOLD_OR = 2 };

constexpr int kMaxPasses = 40;
// Cross-GPU reduction fused into the tail of an eval kernel (fle_group_*, fle_group.cu): every rank owns
// a mailbox of kGroupSlots x FLE_MAX_GROUP entries of 4 x u64 {sequence flag, value 0, value 1, -};
// exchange number `seq` of a rank posts its totals into entry [seq % kGroupSlots][rank] of EVERY
// rank's mailbox with peer stores over NVLink (release) and reads the entries of all ranks in
// its own mailbox (acquire) -- protocol in fle_eval_common.cuh.
constexpr int kGroupSlots = 8;
constexpr int kGroupEntryWords = 4;
constexpr int kGroupMailWords = kGroupSlots * FLE_MAX_GROUP * kGroupEntryWords;
constexpr int kMaxCombs = 64;
constexpr int kMaxSlots = 32;

struct PassDesc {
  uint8_t type;     // PassType
  uint8_t col;      // column index
  uint8_t k;        // chains in this pass: 1, 2 or 4 (3 is padded by repeating the last chain,
                    // which is harmless because SET/AND/OR of the same value are idempotent);
                    // PASS_IN / PASS_CONST: 1
  uint8_t merge;    // 1: all chains fold into slot[0] with the same op -> combine in registers
  uint8_t slot[4];  // destination slot per chain
  uint8_t fold[4];  // FoldOp per chain
  uint8_t neg[4];   // 1: complement the chain result before folding
  uint8_t init[4];  // PASS_LT: chain init (0 => '<', 1 => '<='); PASS_CONST: the constant
  uint32_t lit[4];  // literal per chain (masked to bw bits); PASS_IN: lit[0]=offset, lit[1]=count
  uint32_t k_real;  // chains before padding (what a program specialised at run time unrolls)
};
#ifndef __CUDACC_RTC__
static_assert(sizeof(PassDesc) == 40 && offsetof(PassDesc, slot) == 4 && offsetof(PassDesc, lit) == 20,
              "the staged kernel reads PassDesc as 32-bit words");
#endif

struct CombDesc {
  uint8_t op;  // FOLD_AND / FOLD_OR
  uint8_t dst; // slot written (== a)
  uint8_t a;
  uint8_t b;
};

struct EvalParams {
  const uint64_t* planes[FLE_MAX_COLS];  // already offset to first_row's group
  uint32_t bw[FLE_MAX_COLS];
  uint32_t col_off[FLE_MAX_COLS];     // byte offset of the column's region inside a stage
  uint32_t seg_stride[FLE_MAX_COLS];  // byte distance between the 8 skewed segments
  int32_t ncols, npasses, ncombs, nslots;
  int32_t pad0;
  uint32_t smem_masks_off;                 // per-pass literal masks [npasses][4][32] u32 (staged generic kernel);
                                           // membership table (IN lookup kernel)
  int32_t gpt;      // groups per thread per tile
  int32_t nstages;  // TMA pipeline depth
  uint32_t stage_bytes;
  uint32_t smem_bar_off, smem_lits_off, smem_slots_off, smem_stage_off;
  int64_t ngroups;
  int64_t nrows;
  uint64_t slot0_init;
  uint64_t* bitmap;
  uint64_t* partial;    // per-CTA popcounts (gridDim entries)
  unsigned int* ticket; // self-resetting completion counter
  unsigned int* tile_counter;  // self-resetting dynamic tile scheduler (single-pass kernels), may be null
  uint64_t* count_out;  // may be null
  const uint32_t* in_lits;
  int32_t n_in_lits;
  int32_t agg_ndict;
  // optional aggregate fused into a specialised kernel (fle_eval_sum_product_dev):
  // *agg_sum_out = SUM over selected rows r of agg_plain[r] * agg_dict[code of row r in column agg_col]
  const int64_t* agg_plain;
  const int64_t* agg_dict;
  int64_t* agg_sum_out;
  // optional: the totals (count, aggregate) are summed over the GPUs of a group before they are
  // written (grp_world > 1); see kGroupSlots above
  uint64_t* grp_mail[FLE_MAX_GROUP];  // mailbox of every rank, the own one included
  unsigned long long* grp_lost;       // peers that timed out so far
  int32_t grp_world, grp_rank;
  int32_t grp_deferred;
  uint32_t grp_seq;                   // this exchange's sequence number (host-side counter, never 0)
  uint32_t grp_pend_seq, grp_pad;     // deferred exchange to collect first (0: none) and where its totals go
  uint64_t* grp_pend_count;
  uint64_t* grp_pend_sum;
  uint64_t grp_timeout_ns;
  PassDesc passes[kMaxPasses];
  CombDesc combs[kMaxCombs];
  // Columns whose first group does not sit on a 16-byte boundary (page payloads inside a file image start one
  // byte after the width header, dict-encoding.h:185-192): planes[c] is then rounded DOWN to 16 bytes and
  // plane_skew[c] = the bytes to skip (1..15).  The TMA producer stages the 16-byte-aligned superset of every
  // tile segment and the consumers rebuild the plane words from 32-bit loads with a funnel shift
  // (lds_pair_unaligned, fle_eval_common.cuh) -- the column is read from HBM exactly once, no realigned copy.
  // Only kernels built for it look at these fields (eval1_skew kernels, specialised kernels with FLE_JIT_SKEW).
  uint32_t plane_skew[FLE_MAX_COLS];
  int64_t staged_groups;  // groups that may go through staged tiles (the staged superset of the last groups must
                          // not run past the end of the column buffer); == ngroups for aligned columns
  // single-IN-list kernels only: the list's bitmap is combined with an earlier bitmap on the way out,
  // out[g] = IN(g) AND / OR comb_bm[g] (comb_op = FOLD_AND / FOLD_OR, 0 = none; comb_bm may be the output itself)
  const uint64_t* comb_bm;
  int32_t comb_op, comb_pad;
  // IN list on a narrow column (bw <= 7) evaluated in the plane domain (fle_eval1.cu, InMuxRegs):
  // in_masks[c] says whether code c is a member of the list -- as ~0 / 0 where the term is ANDed in on the
  // ALU pipe, as 1 / 0 where it is multiplied in on the FMA pipe (in_mux_term_on_fma)
  uint32_t in_masks[128];
};

// which of the 2^min(bw, 3) minterm terms of a leaf go to the FMA pipe (IMAD) instead of the ALU pipe (LOP3)
inline
#ifdef __CUDACC__
__host__ __device__
#endif
constexpr bool in_mux_term_on_fma(int bw, int j) { return bw >= 3 && ((j & 1) != 0 || (bw == 7 && j == 6)); }

}  // namespace fle
