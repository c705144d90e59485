// Internal declarations shared by the translation units of libfle_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fle_b200.h"
#include "fle_params.h"
#include "fle_jit.h"

namespace fle {

struct LoweredProgram {
  EvalParams p;                  // passes / combs / nslots / slot0_init filled in
  std::vector<uint32_t> in_lits; // IN literal pool
  int lit_slot = -1;             // literal slot of the context holding the pool (fle_ctx::lit_slots) ...
  uint64_t lit_gen = 0;          // ... and the slot's generation when it was written: a slot that has been
                                 // handed to another program since is uploaded again before a launch
  // run-time specialisation (fle_jit.cu), filled in lazily by launch_prepared
  JitPlan jit_plan;
  uint32_t jit_bw[FLE_MAX_COLS] = {0};  // bit widths jit_plan was made for
  int jit_state = 0;                     // 0 not planned, 1 planned, -1 shape not supported
  int jit_agg_col = -1;                  // aggregate column jit_plan was made for (-1: none)
  std::shared_ptr<JitEntry> jit_entry;
  std::vector<uint32_t> jit_args;
};

// Returns FLE_OK or FLE_E_*; err gets a human-readable reason.
// mask_literals: keep only the low bit_width bits of a literal (the reference's plane algebra); otherwise
// comparisons with literals >= 2^bit_width are folded to constants
int lower_program(const fle_col_desc* cols, int ncols, const fle_pred_node* nodes, int nnodes,
                  LoweredProgram* out, std::string* err, bool mask_literals = false);

}  // namespace fle

// ------------------------------------------------------------------------------------------
// Context
// ------------------------------------------------------------------------------------------
struct fle_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int num_sms = 0;
  int max_smem_optin = 0;   // per block
  int max_smem_per_sm = 0;
  struct OccEntry { const void* kern; size_t smem; int blocks; };
  std::vector<OccEntry> occ_cache;
  int64_t launches = 0;
  int last_mode = -1;        // launch mode of the last fused-eval kernel (tuning scripts)
  int64_t jit_launches = 0;  // ... of which kernels specialised at run time
  std::string last_error;

  // options
  int eval_variant = 1;      // 0 direct loads, 1 staged (single-pass kernels when possible), 2 staged generic only
  int eval_stages = 0;       // 0 = auto
  int eval_block = 256;
  int eval_ctas_per_sm = 0;  // 0 = auto (occupancy)
  int eval_gpt = 0;          // 0 = auto
  int eval_narrow_bw = 2;    // single-pass programs on columns up to this width skip shared-memory staging (bw 1, 2, 4)
  int in_lookup_min = 9;     // IN lists at least this long use the decode+lookup kernel (bw <= 16); shorter ones the plane form
  int in_wide = 2;           // IN lookup kernel: two-CTA bit table (0), one-CTA producer/consumer kernel with 16-bit entries (1),
                             // ... with word / byte entries addressed straight from the transposed word, bw <= 12 (2)
  int in_wide_warps = 0;     // consumer warps per CTA of the one-CTA IN lookup kernel (form 2: 16 or 24; 0 = measured default)
  int in_wide_gpt = 0;       // groups per consumer thread per tile (form 2: 1, or 2 with 16 warps; 0 = measured default)
  int in_mux_min = 4;        // lists of at least this many literals on columns up to in_mux_bw use the multiplexer tree even below in_lookup_min
  int in_mux_bw = 6;         // IN lists of in_lookup_min+ literals on columns up to this width run in the plane domain (0 = never)
  int64_t in_split_min_rows = 1ll << 20;  // calls of at least this many rows evaluate long IN lists (>= 16 literals) inside larger programs with the lookup kernels; -1 = never
  int codec_variant = 1;
  int codec_ctas_per_sm = 0;  // 0 = as many as fit
  int eval_jit = 1;                      // 0 never, 1 when n >= eval_jit_min_rows or already compiled, 2 always
  int64_t eval_jit_min_rows = 1ll << 24;
  int64_t h2d_chunk_rows = 0;  // 0 = auto
  bool reference_compat = false;  // reproduce the reference where this library deliberately differs (literal masking)

  // scratch owned by the context (device)
  uint64_t* d_partial = nullptr;  // per-CTA counts
  int partial_cap = 0;
  unsigned int* d_ticket = nullptr;
  bool counters_dirty = false;    // a CUDA error was reported: clear d_ticket before its next use
  // IN literal pools: a ring of {pinned host, device} slots, one per prepared program.  Uploads and
  // kernels are ordered by the context's one stream, so a device slot can be rewritten as soon as the
  // copy is enqueued; only the PINNED side needs care (the host must not overwrite it before the
  // previous copy has read it): one event per slot, waited for only when the ring wraps around
  // onto an upload that is still in flight.  No stream synchronisation on the enqueue path.
  struct LitSlot {
    uint32_t* h = nullptr;
    uint32_t* d = nullptr;
    int cap = 0;  // 32-bit words
    cudaEvent_t ev = nullptr;
    bool ev_pending = false;
    uint64_t gen = 0;
  };
  static constexpr int kLitSlots = 16;
  LitSlot lit_slots[kLitSlots];
  int lit_next = 0;
  uint64_t lit_gen_counter = 0;
  uint64_t* d_count = nullptr;    // 8-byte result slot for host-facing calls
  void* d_dict = nullptr;         // dictionary build scratch (hash table, lists, codes), grow-only
  size_t dict_cap = 0;
  void* d_realign = nullptr;      // aligned copies of columns no kernel can read in place, grow-only
  size_t realign_cap = 0;
  uint64_t* d_split = nullptr;    // temporary bitmaps of regrouped programs (long IN lists), grow-only
  size_t split_cap = 0;           // bytes
  uint64_t* h_count = nullptr;    // pinned
  int null_fused = 2;                  // NULL path (option "null_fused"): 0 data-domain bitmap, then deposit; 1 single-pass subtrees evaluate +
                                       // deposit in one kernel per column; 2 a chain of such subtrees in ONE kernel (fle_null_chain.cu)
  int null_chain_ctas = 0;             // CTAs of the chain kernel (0 = fill the device; tests pin 1-3 to walk the tile pipeline)
  bool nulls_seen = false;             // the previous nullable evaluation found NULLs (eval_dev_nullable)
  void* d_null = nullptr;              // temporaries of the NULL path (data-domain bitmaps, leaf bitmaps), grow-only
  size_t null_cap = 0;
  uint64_t* d_chunk_counts = nullptr;  // per-chunk counts of the host-buffer pipeline (fle_eval), grow-only
  size_t chunk_counts_cap = 0;         // 64-bit words

  // host-facing pipelines: staging buffers
  void* d_stage[2] = {nullptr, nullptr};
  size_t d_stage_cap[2] = {0, 0};
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copy[2] = {nullptr, nullptr};
  cudaEvent_t ev_done[2] = {nullptr, nullptr};

  // single-pass bitmap compaction (fle_rowgroup.cu): tile status words, tagged with a launch generation
  uint64_t* d_tile_state = nullptr;
  size_t tile_state_cap = 0;  // bytes
  uint32_t tile_gen = 0;
  cudaEvent_t ev_prefetch = nullptr;  // last fle_h2d_prefetch on the copy stream
  fle_ctx* copy_owner = nullptr;      // fle_ctx_share_copy_stream: prefetches and marks go through this context's copy stream
  static constexpr int kPrefetchMarks = 8;
  cudaEvent_t ev_mark[kPrefetchMarks] = {nullptr};  // fle_prefetch_mark / fle_prefetch_wait

  // group of GPUs whose totals are summed inside the kernels (fle_group.cu)
  int grp_world = 0, grp_rank = 0;
  bool grp_on = false;                  // set for the duration of a fle_*_group_dev call
  bool grp_deferred = false;            // option "group_deferred"
  bool grp_ipc = false;                 // peers' mailboxes were opened from IPC handles
  uint64_t* d_grp_mail = nullptr;       // own mailbox (cudaMalloc: exportable)
  unsigned long long* d_grp_state = nullptr;  // [0] peers lost to time-outs
  uint32_t grp_seq = 0;                 // exchanges issued (host-side: every member issues the same sequence)
  uint32_t grp_pend_seq = 0;            // deferred exchange not collected yet (0: none) and where its totals go
  uint64_t* grp_pend_count = nullptr;
  uint64_t* grp_pend_sum = nullptr;
  uint64_t* grp_peer_mail[FLE_MAX_GROUP] = {nullptr};
  int64_t grp_timeout_ms = 10000;
};

namespace fle {

int set_cuda_error(fle_ctx* ctx, cudaError_t e, const char* what);
int set_error(fle_ctx* ctx, int code, const std::string& msg);
int ensure_scratch(fle_ctx* ctx, int nblocks, int n_in_lits);
// allocates what a context's first launches would allocate lazily (members of a group must not allocate inside a group call)
int prime_context(fle_ctx* ctx);
// copies nwords 32-bit words into the next literal slot (pinned -> device, async on the context's stream);
// *slot / *gen identify it, *dev is the device address the kernels read
int upload_literals(fle_ctx* ctx, const void* words, int nwords, int* slot, uint64_t* gen, const uint32_t** dev);
// fle_group.cu: fills the group fields of a kernel parameter block when a group call is in progress
void group_fill(fle_ctx* ctx, EvalParams* P);
// sums {*a_dev, *b_dev (optional)} over the group in place with a one-warp kernel (paths that end in a
// kernel without the fused exchange)
int launch_group_reduce(fle_ctx* ctx, uint64_t* a_dev, uint64_t* b_dev);
void group_release(fle_ctx* ctx);

#define FLE_CUDA_TRY(ctx, expr)                                   \
  do {                                                            \
    cudaError_t _e = (expr);                                      \
    if (_e != cudaSuccess) return fle::set_cuda_error(ctx, _e, #expr); \
  } while (0)

// kernels' host launchers (fle_eval.cu, fle_codec.cu, fle_bitmap.cu)
// lower + upload the IN literal pool (once per program) ...
int prepare_eval(fle_ctx* ctx, const fle_col_desc* cols, int ncols, const fle_pred_node* nodes,
                 int nnodes, LoweredProgram* lp);
// ... then launch it any number of times on device-resident columns
// agg (optional): SUM(plain[r] * dict[code of row r in column col]) over the selected rows, fused into
// a specialised kernel; FLE_E_UNSUPPORTED when no such kernel can be had (the caller then runs
// eval + launch_sum_product).  bitmap_dev may be null only with agg.
struct AggSpec {
  const int64_t* plain;
  int col;
  const int64_t* dict;
  int ndict;
  int64_t* sum_out;
};
int launch_prepared(fle_ctx* ctx, LoweredProgram* lp, const fle_col_desc* cols_dev, int ncols,
                    int64_t first_row, int64_t n, uint64_t* bitmap_dev, uint64_t* count_dev,
                    const AggSpec* agg = nullptr);
int launch_eval(fle_ctx* ctx, const fle_col_desc* cols_dev, int ncols, const fle_pred_node* nodes,
                int nnodes, int64_t first_row, int64_t n, uint64_t* bitmap_dev,
                uint64_t* count_dev);
// fle_eval1.cu: the single-pass kernel for (PassType, chains, bit width), or nullptr
const void* eval1_kernel_for(int type, int k, int bw);
// fle_eval1.cu: unstaged kernels for bw 1, 2, 4 (coalesced 128-bit global loads), or nullptr
const void* eval1_narrow_kernel_for(int type, int k, int bw);
constexpr int kNarrowUnits = 8;  // 16/32-byte units per thread per tile of the narrow kernels
// fle_eval1_skew.cu: the same kernels for a column at an arbitrary byte address (bw <= 16; LT chains 1-2, EQ 1), or nullptr
const void* eval1_skew_kernel_for(int type, int k, int bw);
const void* eval_in_mux_skew_kernel_for(int bw);
// fle_eval1.cu: IN list on a narrow column as a multiplexer tree over the planes (bw <= 7), or nullptr
const void* eval_in_mux_kernel_for(int bw);
// fle_eval_in.cu: IN list by decode + membership-table lookup (bw <= 16), or nullptr
const void* eval_in_kernel_for(int bw);
// fle_eval_in.cu: IN list by decode + 16-bit entry table (bw <= 12, CTAs of 512 threads), or nullptr
// form 1: 16-bit entries (bw <= 12) / replicated bit table (bw 13..16); form 2: word / byte entries addressed straight
// from the transposed word (bw <= 12), same as form 1 beyond
// ncw consumer warps (+ one producer warp) per CTA, gpt groups per consumer thread per tile
const void* eval_in_wide_kernel_for(int bw, int form, int ncw, int gpt);
void eval_in_wide_shape(int bw, int* ncw, int* gpt);  // the measured default of (ncw, gpt) per bit width
uint32_t eval_in_wide_table_bytes(int bw, int form);
// out = postfix program over K device bitmaps (leaf.col = bitmap index), word by word
int launch_bitmap_combine(fle_ctx* ctx, const uint64_t* const* leaf_bitmaps_dev, int nleaf,
                          const fle_pred_node* nodes, int nnodes, int64_t nbits, uint64_t* out_dev,
                          uint64_t* count_dev);
int launch_pack_u32(fle_ctx* ctx, const uint32_t* vals_dev, int64_t n, int bw, uint64_t* enc_dev);
int launch_unpack_u32(fle_ctx* ctx, const uint64_t* enc_dev, int bw, int64_t n, uint32_t* out_dev);
int launch_unpack_narrow(fle_ctx* ctx, const uint64_t* enc_dev, int bw, int64_t n, int elem_size, void* out_dev);
// fle_bitmap.cu: non-null rows per nullable column (counts_dev[c], 0 for required columns), one launch
int launch_def_counts(fle_ctx* ctx, const fle_col_desc* cols, int ncols, int64_t n, uint64_t* counts_dev,
                      uint64_t* nulls_total_dev = nullptr);
int launch_bitmap_count(fle_ctx* ctx, const uint64_t* bitmap_dev, int64_t nbits, uint64_t* count_dev);
int launch_intersect(fle_ctx* ctx, uint64_t* root_dev, int64_t n_root, const uint64_t* sub_dev,
                     int64_t n_sub);
// out = out (comb_op: FOLD_SET / FOLD_AND / FOLD_OR) deposit(not-null bitmap read straight from 1-bit def planes, sub);
// count_dev (optional) receives popcount(out)
int launch_deposit(fle_ctx* ctx, const uint64_t* def_planes_dev, int64_t n_root, const uint64_t* sub_dev, int64_t n_sub,
                   uint64_t* out_dev, int comb_op, uint64_t* count_dev);
// fle_eval1_deposit.cu: the same with the sub-program (leaves on column 0 = col, n_sub values in its data stream)
// evaluated inside the deposit kernel -- no data-domain bitmap.  FLE_E_UNSUPPORTED: not a single supported pass.
int launch_eval_deposit(fle_ctx* ctx, const fle_col_desc& col, const fle_pred_node* nodes, int nnodes, int64_t n_sub,
                        const uint64_t* def_planes_dev, int64_t n_root, uint64_t* out_dev, int comb_op, uint64_t* count_dev);
// fle_null_chain.cu: a chain (one operator) of single-column single-pass subtrees over nullable / required columns
// evaluated, deposited and folded in one kernel
constexpr int kNcMaxCols = 4;
struct NullChainPlan {
  int ncols = 0;
  int join_op = FOLD_AND;
  int col[kNcMaxCols];      // index into the caller's column array
  PassDesc pd[kNcMaxCols];  // the column's subtree as one pass
};
int64_t null_chain_tiles(int64_t n);
bool null_chain_fits(fle_ctx* ctx, const int* bw, int ncols);
int null_chain_pass(fle_ctx* ctx, const fle_col_desc& col, const fle_pred_node* nodes, int nnodes, PassDesc* out);
// non-null rows per nullable column (counts_dev[c]) and per tile of 16384 rows (tile_sums_dev[c * stride + tile]), one launch
int launch_def_tile_counts(fle_ctx* ctx, const fle_col_desc* cols, int ncols, int64_t n, uint64_t* counts_dev,
                           uint32_t* tile_sums_dev, size_t sums_stride);
// tile counts -> prefixes (pref_dev[c * stride + tile]), then the chain kernel
int launch_null_chain(fle_ctx* ctx, const NullChainPlan& plan, const fle_col_desc* cols, int64_t n, const uint32_t* tile_sums_dev,
                      uint64_t* pref_dev, size_t sums_stride, uint64_t* out_dev, uint64_t* count_dev);
// fle_plain.cu
int launch_realign(fle_ctx* ctx, const uint8_t* src_dev, size_t nbytes, uint8_t* dst_dev);
int launch_plain_eval(fle_ctx* ctx, const void* values_dev, int elem_type, int64_t n, int op, const void* lits_host,
                      int nlits, int reference_order, uint64_t* bitmap_dev, uint64_t* count_dev);
// fle_dict.cu
int launch_dict_encode(fle_ctx* ctx, const void* values_dev, int elem_type, int64_t n, int max_entries,
                       int64_t* dict_out_dev, int* num_entries, int* bit_width, uint8_t* enc_dev, size_t enc_cap,
                       size_t* enc_len);
// fle_scan.cu
int launch_select(fle_ctx* ctx, const uint64_t* bitmap_dev, int64_t nbits, uint32_t* ids_dev, int64_t cap,
                  uint64_t* count_dev);
int launch_gather(fle_ctx* ctx, const uint64_t* planes_dev, int bw, const uint32_t* ids_dev,
                  const uint64_t* count_dev, int64_t cap, const void* dict_dev, int ndict, int elem_size,
                  void* out_dev);
// fle_rowgroup.cu
int launch_pages_gather(fle_ctx* ctx, const fle_page_ref* refs, int nrefs, uint64_t* null_rows_dev);
int launch_assemble(fle_ctx* ctx, const uint64_t* bitmap_dev, int64_t nrows, const fle_slot_desc* cols, int ncols,
                    int tuple_size, uint8_t* tuples_dev, int64_t cap_rows, uint64_t* count_dev, uint64_t* bad_codes_dev);
int launch_sum_product(fle_ctx* ctx, const uint64_t* bitmap_dev, int64_t nrows, const int64_t* plain_dev,
                       const uint64_t* planes_dev, int bw, const int64_t* dict_dev, int ndict, int64_t* sum_dev,
                       uint64_t* count_dev);

}  // namespace fle
