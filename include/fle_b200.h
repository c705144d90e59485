/* fle_b200.h -- C ABI of libfle_b200.so: the B200 (sm_100a) implementation of the
 * fixed-length-encoding (FLE) hot path of zuowang/Impala-avx2-parquet-scanner.
 *
 * This is the drop-in boundary.  Everything above it (FleEncoder/FleDecoder,
 * DictDecoder<T>, the SimplePredicate tree, the column-reader batch interface) stays host
 * C++ (impala-avx2-parquet-scanner_b200/host/) and calls only these entry points.  Plain
 * pointers and sizes; no C++/torch types.  Each function cites the reference interface it
 * replaces (file:line in the reference tree).
 *
 * Data layout (identical to the reference, fle-encoding.h:8338-8340, :426-430):
 *   64 values = 1 group = bit_width consecutive little-endian uint64 words; word i holds
 *   bit i of all 64 values; value k of the group is bit (63 - k) of every word.
 * Selection bitmaps follow boost::dynamic_bitset<unsigned long> as the reference fills it
 *   (fle-encoding.h:7989-8008): row r -> bit (r & 63) of uint64 word (r >> 6); a set bit
 *   means the row PASSES (hdfs-parquet-scanner.cc:1139-1140).  Bits past the last row are 0.
 *
 * Conventions: every function returns FLE_OK (0) or a negative FLE_E_* code and never
 * throws or aborts (the reference's predicate functions return void / DCHECK; its setup
 * paths return Status -- the host shims map codes to Status / false).  A context is
 * single-threaded (one scanner thread, one CUDA stream), many contexts may be used
 * concurrently; there is no hidden global state.  The caller owns every buffer it passes.
 * "_dev" functions take DEVICE pointers and only enqueue work on the context's stream;
 * the others take HOST pointers and return after the result is in host memory.
 * There is no CPU fallback: without a usable CUDA device every call fails with
 * FLE_E_NODEVICE / FLE_E_CUDA.
 */
#ifndef FLE_B200_H
#define FLE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define FLE_B200_VERSION 100

enum {
  FLE_OK = 0,
  FLE_E_INVALID = -1,     /* bad argument (NULL, bit width outside 1..32, malformed program) */
  FLE_E_NOMEM = -2,       /* host or device allocation failed */
  FLE_E_CUDA = -3,        /* CUDA runtime error; fle_last_error() has the text */
  FLE_E_BUFFER = -4,      /* output buffer too small / input shorter than the rows asked for */
  FLE_E_ALIGN = -5,       /* pointer or row offset not aligned as documented */
  FLE_E_UNSUPPORTED = -6, /* valid request this build cannot run (e.g. > FLE_MAX_COLS columns) */
  FLE_E_NODEVICE = -7,    /* no CUDA device: the product has no CPU path */
  FLE_E_DICT_FULL = -8    /* more distinct values than dictionary entries (DictEncoder::Put == -1) */
};

/* Predicate opcodes.  EQ..IN are leaves = FleDecoder::Eq/Lt/Le/Gt/Ge/In
 * (fle-encoding.h:7962-8313); AND/OR = AndOperate/OrOperate (simple-predicates.h:145-163);
 * ALL/NONE = the DictDecoder short-circuits `resize(num_rows, true/false)`
 * (dict-encoding.h:466,476-478,489-491,502-504,514-516,537). */
enum {
  FLE_OP_EQ = 0, FLE_OP_LT = 1, FLE_OP_LE = 2, FLE_OP_GT = 3, FLE_OP_GE = 4, FLE_OP_IN = 5,
  FLE_OP_AND = 6, FLE_OP_OR = 7, FLE_OP_ALL = 8, FLE_OP_NONE = 9
};

#define FLE_MAX_COLS 8      /* columns per fused evaluation */
#define FLE_MAX_NODES 64    /* postfix nodes per fused evaluation */
#define FLE_MAX_IN_CODES 4096
#define FLE_MAX_GROUP 8     /* GPUs of one box whose totals one evaluation can sum */
#define FLE_GROUP_HANDLE_BYTES 64

/* One encoded column (chunk / page payload).  `planes` points at the first FLE group, i.e.
 * AFTER the 1-byte bit-width header DictDecoderBase::SetData strips (dict-encoding.h:185-192).
 * Nullable columns (max_def_level > 0, hdfs-parquet-scanner.cc:342-345): def_planes is the
 * 1-bit FLE definition-level stream covering the rows; `planes` then holds only the
 * non-NULL values, indexed by non-null ordinal.  NULL rows fail every leaf. */
typedef struct fle_col_desc {
  const void* planes;
  uint64_t bytes;         /* byte length readable at planes */
  int32_t bit_width;      /* 1..32 */
  int32_t max_def_level;  /* 0 = required column, 1 = nullable */
  const void* def_planes; /* NULL when max_def_level == 0 */
  uint64_t def_bytes;
} fle_col_desc;

/* One node of a postfix (reverse-Polish) predicate program.  Leaves name a column and carry
 * dictionary CODES pre-encoded on the host (DictDecoder<T>::Eq..In, dict-encoding.h:461-541).
 * Codes are expected to fit the column's bit width (the reference's precondition); one that does
 * not is folded to the constant the comparison means ("reference_compat" = 1: its high bits are
 * ignored, as in the reference's plane algebra). */
typedef struct fle_pred_node {
  int32_t op;            /* FLE_OP_* */
  int32_t col;           /* leaf: index into cols[] */
  uint64_t code;         /* EQ/LT/LE/GT/GE literal */
  const uint64_t* codes; /* IN: host array of ncodes literals (always a HOST pointer) */
  int32_t ncodes;
  int32_t reserved;
} fle_pred_node;

typedef struct fle_ctx fle_ctx;

/* ---- context --------------------------------------------------------------------------- */
/* cuda_stream: a cudaStream_t to enqueue on (e.g. torch's current stream), or NULL to let the
 * context create its own non-blocking stream. */
int fle_ctx_create(int device, void* cuda_stream, fle_ctx** out);
int fle_ctx_destroy(fle_ctx* ctx);
int fle_ctx_set_stream(fle_ctx* ctx, void* cuda_stream);
int fle_ctx_sync(fle_ctx* ctx);
/* Tuning knobs, all optional: "eval_variant" (0 = direct-load generic kernel for everything;
 * 1 [default] = TMA-staged kernels: single-pass kernels specialised per bit width where the
 * program is one pass, the staged generic kernel otherwise; 2 = staged generic kernel for
 * everything),
 * "eval_jit" (programs that are not one pass -- several columns, mixed AND/OR trees -- can run
 * on a kernel specialised for the program's SHAPE at run time, compiled once per shape with
 * NVRTC and cached process-wide; literals are not part of the shape.  0 = never, 1 [default] =
 * once the rows evaluated with the shape in this process (this call included) reach
 * "eval_jit_min_rows" (default 2^24; the compilation happens inside that call), 2 = always, 3 = like 1 but the
 * compilation runs on a thread of the library's own and the caller keeps to the interpreting kernels until the
 * kernel is there -- no call ever waits for the compiler; while one thread compiles a shape, other threads'
 * cache hits are not held up in any mode; without a usable libnvrtc.so.12 modes 0/1/3 use the interpreting
 * kernels and mode 2 fails with FLE_E_UNSUPPORTED),
 * "eval_stages", "eval_ctas_per_sm", "eval_gpt" (0 = auto: launch shape of the staged kernels),
 * "eval_narrow_bw" (single-pass programs on columns up to this width -- default 2; 1, 2 or 4
 * exist -- read the planes with coalesced 128-bit loads instead of staging them),
 * "in_lookup_min" (IN lists of at least this many literals on columns of <= 16 bits use the
 * decode + membership-lookup kernels, shorter ones equality chains on the planes; default 9,
 * 2^30 disables the lookup kernels),
 * "in_split_min_rows" (device-resident evaluations of at least this many rows -- default 2^20 --
 * evaluate IN lists of >= 16 literals that sit INSIDE a larger program with the lookup kernels and
 * combine bitmaps, instead of walking the literal pool in the fused kernel; -1 = never),
 * "in_wide" (2 [default]: the lookup runs on the one-CTA-per-SM producer/consumer kernel with table
 * entries addressed straight from the transposed word -- 32-bit entries for bw <= 10, byte entries
 * for bw 11 / 12, three instructions per lookup -- and a replicated bit table for bw 13..16;
 * 1: the same kernel with 16-bit entries for bw <= 12 (round 1's form); 0: the two-CTA bit-table
 * kernel; the older forms stay as on-device cross-checks),
 * "in_wide_warps", "in_wide_gpt" (0 = measured default: consumer warps per CTA and groups per
 * consumer thread per tile of the form-2 lookup kernel),
 * "in_mux_bw", "in_mux_min" (IN lists on columns up to in_mux_bw bits -- default 6, at most 7; 0 = never --
 * are evaluated in the plane domain as a multiplexer tree from in_mux_min literals on, default 4),
 * "codec_variant" (0 = direct loads/stores; 1 [default] = per-bit-width choice among the staged
 * unpack kernels; 2 = per-row bulk stores, 3 = warp-private copy-out, 4 = CTA-wide copy-out),
 * "codec_ctas_per_sm" (0 = auto), "h2d_chunk_rows",
 * "null_fused" (nullable columns that DO hold NULLs, ColumnReader<T>::Eq..In + IntersectBitset,
 * hdfs-parquet-scanner.cc:326-345: 2 [default] = a conjunct / disjunct chain of up to four column subtrees that
 * are one pass each -- compares, BETWEENs, equalities; required columns may take part -- is evaluated, deposited
 * through the def levels and folded in ONE kernel; 1 = one evaluate-and-deposit kernel per column subtree;
 * 0 = data-domain bitmap, then deposit; programs a mode does not cover take the next one),
 * "null_chain_ctas" (0 = fill the device; a test knob that pins the chain kernel's grid),
 * "reference_compat" (default 0; 1 reproduces the reference where this library deliberately differs:
 * a literal >= 2^bit_width keeps only its low bit_width bits, as in the reference's plane algebra,
 * instead of being folded to the constant the comparison means -- `x < 2^bw` is true for every code;
 * the host-side drop-ins have their own switch, FleGpu::SetReferenceCompat). */
int fle_ctx_set_option(fle_ctx* ctx, const char* name, int64_t value);
int64_t fle_ctx_get_option(fle_ctx* ctx, const char* name);
/* Number of kernels of this library launched through ctx so far (bench `gpu_launches`). */
int64_t fle_ctx_launch_count(fle_ctx* ctx);
const char* fle_last_error(fle_ctx* ctx);
const char* fle_strerror(int code);
int fle_version(void);
int fle_device_count(void);

/* ---- memory helpers (so host C++ needs no CUDA headers) ---------------------------------- */
int fle_host_alloc(size_t bytes, void** out);   /* pinned host memory */
int fle_host_free(void* p);
int fle_dev_alloc(fle_ctx* ctx, size_t bytes, void** out); /* 256-byte aligned device memory */
int fle_dev_free(fle_ctx* ctx, void* p);
int fle_h2d(fle_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes); /* async on ctx stream */
int fle_d2h(fle_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes); /* async on ctx stream */

/* ---- codec ----------------------------------------------------------------------------- */
/* Byte length FleEncoder::Flush() reports for n values (fle-encoding.h:9806-9811):
 * max(1, ceil(n / 64)) * bit_width * 8. */
size_t fle_encoded_bytes(int64_t n, int bit_width);

/* Replaces FleEncoder::Put x n + Flush (fle-encoding.h:8315-8365, :9806; Pack_1..16 and
 * Pack_16_32 :8367-9803).  vals[i] must be < 2^bit_width (the reference DCHECKs it; here the
 * high bits are ignored).  Unfilled rows of the last group are written as 0 (the reference
 * leaves stale staging data there; those rows are never decoded).  FLE_E_BUFFER when
 * enc_cap < fle_encoded_bytes(n, bw), mirroring Put() == false (fle-encoding.h:8318). */
int fle_pack_u32(fle_ctx* ctx, const uint32_t* vals, int64_t n, int bit_width, uint8_t* enc,
                 size_t enc_cap, size_t* enc_len);
int fle_pack_u32_dev(fle_ctx* ctx, const uint32_t* vals_dev, int64_t n, int bit_width,
                     uint8_t* enc_dev, size_t enc_cap, size_t* enc_len);

/* Replaces FleDecoder::Get<T>(T*) x n after first_row values were consumed
 * (fle-encoding.h:404-567; Unpack_1..32 :569-7329) and the whole-group jump of
 * Get(T*, skip) / Skip (:344-402).  FLE_E_BUFFER when the rows are not inside enc_len,
 * mirroring Get() == false at the buffer guard (:407).  The _dev variant requires
 * first_row % 64 == 0 and 16-byte aligned enc_dev (FLE_E_ALIGN otherwise). */
int fle_unpack_u32(fle_ctx* ctx, const uint8_t* enc, size_t enc_len, int bit_width,
                   int64_t first_row, int64_t n, uint32_t* out);
int fle_unpack_u32_dev(fle_ctx* ctx, const uint8_t* enc_dev, size_t enc_len, int bit_width,
                       int64_t first_row, int64_t n, uint32_t* out_dev);

/* Narrow decode targets: FleDecoder::Get<T> fills 1-byte values for bit_width <= 8 and 2-byte values for
 * bit_width <= 16 (the u8 / u16 staging of fle-encoding.h:236-238, :8354-8361, used by int8 / int16 columns,
 * dict-test.cc:64-157).  Same contract as fle_unpack_u32_dev with out_dev = n values of 1 / 2 bytes
 * (16-byte aligned); FLE_E_INVALID when bit_width exceeds the element. */
int fle_unpack_u8_dev(fle_ctx* ctx, const uint8_t* enc_dev, size_t enc_len, int bit_width,
                      int64_t first_row, int64_t n, uint8_t* out_dev);
int fle_unpack_u16_dev(fle_ctx* ctx, const uint8_t* enc_dev, size_t enc_len, int bit_width,
                       int64_t first_row, int64_t n, uint16_t* out_dev);

/* ---- fused predicate evaluation --------------------------------------------------------- */
/* Replaces, in ONE pass that reads every encoded byte once:
 *   FleDecoder::Eq/Lt/Le/Gt/Ge/In            fle-encoding.h:7962-8313   (leaves)
 *   ColumnReader<T>::Eq..In + IntersectBitset hdfs-parquet-scanner.cc:326-445 (nullable cols)
 *   And/OrOperate::GetBitset                  simple-predicates.h:145-163 (inner nodes)
 *   the conjunct `&=` loop                    hdfs-parquet-scanner.cc:1857-1862
 * for rows [first_row, first_row + n) of every column (first_row % 64 == 0).
 * bitmap_out: ceil(n / 64) uint64 words.  match_count (optional): popcount of the bitmap.
 * fle_eval takes host column buffers (copied to the device inside the call, overlapped
 * with evaluation in row chunks); fle_eval_dev takes device-resident columns, bitmap and
 * count pointers and only enqueues work. */
int fle_eval(fle_ctx* ctx, const fle_col_desc* cols, int ncols, const fle_pred_node* postfix,
             int nnodes, int64_t first_row, int64_t n, uint64_t* bitmap_out,
             uint64_t* match_count);
int fle_eval_dev(fle_ctx* ctx, const fle_col_desc* cols_dev, int ncols,
                 const fle_pred_node* postfix, int nnodes, int64_t first_row, int64_t n,
                 uint64_t* bitmap_out_dev, uint64_t* match_count_dev);

/* fle_eval_dev for OPTIONAL columns that are expected to hold no NULL (every column of an Impala-written file is
 * OPTIONAL, hdfs-parquet-table-writer.cc:103-105; TPC-H data has no NULL): checks the definition levels and
 * evaluates the program as if the columns were required, WITHOUT waiting for the check -- nothing is read back,
 * the call only enqueues work.  *null_rows_dev = number of NULL rows found over all nullable columns; when it is 0
 * bitmap and count are exactly those of fle_eval_dev, otherwise they are meaningless (the data streams are
 * indexed by non-null ordinal) and the caller re-issues fle_eval_dev, which handles NULLs
 * (hdfs-parquet-scanner.cc:326-345). */
int fle_eval_no_nulls_dev(fle_ctx* ctx, const fle_col_desc* cols_dev, int ncols, const fle_pred_node* postfix,
                          int nnodes, int64_t first_row, int64_t n, uint64_t* bitmap_out_dev,
                          uint64_t* match_count_dev, uint64_t* null_rows_dev);

/* The run-time specialisation behind "eval_jit", without a device: writes the CUDA C++ generated
 * for the program's shape (compile == 0) or compiles it for sm_100a with NVRTC and writes the
 * compiler log, empty on success (compile == 1), NUL-terminated, into out[0..cap); *len (optional)
 * receives the untruncated length.  compile == 2 writes the sm_100a cubin instead (binary; for
 * cuobjdump -sass).  Bits 8..15 of `compile`: 1 + code_col to generate the fused aggregate of
 * fle_eval_sum_product_dev as well, 0 for none; bit 16: the variant for columns at arbitrary byte addresses
 * (page images evaluated in place).  Only cols[i].bit_width is read.  FLE_E_UNSUPPORTED: the
 * program has no specialised form (it is a single pass, or too large) or NVRTC is missing.
 * This is the step Impala's LLVM codegen performs for the expression trees the reference
 * interprets (simple-predicates.h:113-163). */
int fle_eval_codegen(const fle_col_desc* cols, int ncols, const fle_pred_node* postfix, int nnodes,
                     int compile, char* out, size_t cap, size_t* len);

/* ---- bitmap utilities ------------------------------------------------------------------- */
/* dynamic_bitset::count() over the first nbits bits. */
int fle_bitmap_count_dev(fle_ctx* ctx, const uint64_t* bitmap_dev, int64_t nbits,
                         uint64_t* count_dev);
/* Replaces ColumnReader<T>::IntersectBitset (hdfs-parquet-scanner.cc:326-331): the i-th set
 * bit of root (n_root bits) takes bit i of sub; root is updated in place. */
int fle_intersect_bitset(fle_ctx* ctx, uint64_t* root, int64_t n_root, const uint64_t* sub,
                         int64_t n_sub);
int fle_intersect_bitset_dev(fle_ctx* ctx, uint64_t* root_dev, int64_t n_root,
                             const uint64_t* sub_dev, int64_t n_sub);

/* ---- survivor materialisation (device-resident) ----------------------------------------- */
/* Replaces the bitmap -> skip-distance list of AssembleRows (hdfs-parquet-scanner.cc:1134-1148):
 * writes the row indices of the set bits, ascending, to row_ids_dev[0..min(count, cap)) and the
 * number of set bits to *count_dev.  nbits < 2^32. */
int fle_bitmap_select_dev(fle_ctx* ctx, const uint64_t* bitmap_dev, int64_t nbits,
                          uint32_t* row_ids_dev, int64_t cap, uint64_t* count_dev);
/* Replaces ReadValue(skip) -> DictDecoder<T>::GetValue(T*, skip) -> FleDecoder::Get(T*, skip)
 * per survivor (hdfs-parquet-scanner.cc:1151-1168, dict-encoding.h:321-330,
 * fle-encoding.h:344-379): out[j] = code of row row_ids_dev[j] for j < min(*count_dev, cap).
 * With dict_dev != NULL (elem_size 4 or 8 bytes, ndict entries) the dictionary value is written
 * instead of the code (a code >= ndict yields 0, where GetValue returns false). */
int fle_gather_dev(fle_ctx* ctx, const uint8_t* enc_dev, size_t enc_len, int bit_width,
                   const uint32_t* row_ids_dev, const uint64_t* count_dev, int64_t cap,
                   const void* dict_dev, int ndict, int elem_size, void* out_dev);
/* Downstream aggregate of a TPC-H Q6-shaped scan, fused with the materialisation:
 * *sum_dev = SUM over rows r with bitmap bit r set of plain_i64_dev[r] * dict_i64_dev[code(r)],
 * exact in int64 (e.g. l_extendedprice in cents (PLAIN) x l_discount (dictionary, cents));
 * *count_dev = number of such rows.  bitmap row r <-> plain_i64_dev[r] <-> row r of enc_dev. */
int fle_sum_product_dev(fle_ctx* ctx, const uint64_t* bitmap_dev, int64_t nrows,
                        const int64_t* plain_i64_dev, const uint8_t* enc_dev, size_t enc_len,
                        int bit_width, const int64_t* dict_i64_dev, int ndict, int64_t* sum_dev,
                        uint64_t* count_dev);

/* ---- one row group of the scanner as a batch (AssembleRows, hdfs-parquet-scanner.cc:1078-1275) -- */
/* Uploads on the context's COPY stream (a second stream), so that the column chunks of row group
 * k+1 cross PCIe while row group k is evaluated and materialised on the context's main stream;
 * fle_prefetch_fence makes the main stream wait (on the device, not the host) for every upload
 * enqueued so far.  src_host should be pinned (fle_host_alloc / fle_host_register): a pageable
 * source is staged by the driver at a fraction of the PCIe rate. */
int fle_h2d_prefetch(fle_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int fle_prefetch_fence(fle_ctx* ctx);
/* The same, per row group: fle_prefetch_mark records a point (slot 0..7) on the copy stream behind the uploads
 * enqueued so far, fle_prefetch_wait makes the main stream wait for THAT point only.  With row group k + 1's upload
 * enqueued before row group k's kernels -- as a scanner does to keep PCIe busy -- the fence above would make k's
 * kernels wait for k + 1's bytes as well; a mark per row group lets them start as soon as k's own column chunks
 * have arrived (hdfs-parquet-scanner.cc:1078-1275 reads one row group at a time; this is the overlap around it). */
int fle_prefetch_mark(fle_ctx* ctx, int slot);
int fle_prefetch_wait(fle_ctx* ctx, int slot);
/* Makes ctx upload through owner's copy stream (and use owner's marks): a scanner thread that keeps two row groups in
 * flight on two contexts -- so that neither's kernels queue behind the other's (hdfs-parquet-scanner.cc:1078-1275
 * handles one row group at a time; this is the overlap around it) -- still wants its uploads to cross PCIe ONE AFTER
 * THE OTHER: two copy streams share the link, and the row group that is needed first arrives last.  owner = NULL (or
 * ctx itself) undoes it.  owner must outlive ctx's use of it; same device. */
int fle_ctx_share_copy_stream(fle_ctx* ctx, fle_ctx* owner);
/* Pins / unpins an existing host range (e.g. the I/O buffers the pages are read into) in place. */
int fle_host_register(void* p, size_t bytes);
int fle_host_unregister(void* p);

/* One data page's FLE section inside a column-chunk image that already sits in device memory.
 * The section starts at an arbitrary byte offset (after the page header, the definition levels and
 * the 1-byte width header, dict-encoding.h:185-192) and was written with the page's own bit width
 * (the reference's writer chooses it per page, dict-encoding.h:438-444). */
typedef struct fle_page_ref {
  const void* src_dev;      /* first FLE group of the page, any alignment */
  uint64_t src_bytes;
  int32_t src_bit_width;    /* width the page was packed with */
  int32_t dst_bit_width;    /* width of the destination planes, >= src (missing high planes become zero planes) */
  void* dst_dev;            /* 8-byte aligned; groups * dst_bit_width * 8 bytes */
  int64_t groups;           /* 64-row groups of the page: ceil(rows / 64) */
  const void* def_src_dev;  /* optional: the page's definition levels (FLE, bit width 1), any alignment */
  uint64_t def_bytes;
  int64_t rows;             /* rows of the page (0 with neither def levels nor expand: groups * 64) */
  int32_t expand;           /* 1 (needs def_src_dev): the page's FLE section holds only the NON-NULL values
                             * (hdfs-parquet-scanner.cc:886-921: data values are indexed by non-null ordinal); the
                             * value of every defined row goes to the row's position, NULL rows get code 0, so that
                             * the page joins a row-aligned plane array the fused kernels scan like a required column
                             * (the definition levels, gathered as a 1-bit column of their own, say which rows are
                             * NULL); src_bytes may then be shorter than groups * src_bit_width * 8, even 0 for an
                             * all-NULL page */
  int32_t dst_bit;          /* 0..63: position of the page's FIRST ROW inside the group at dst_dev.  Pages need not
                             * hold whole groups (the reference's writer closes a dictionary page after 524225 rows,
                             * hdfs-parquet-table-writer.cc:387): page k + 1 continues in the group page k ended in --
                             * dst_dev = column base + (first row / 64) * dst_bit_width * 8, dst_bit = first row % 64;
                             * the page then writes ceil((dst_bit + rows) / 64) groups, sharing the first / last one
                             * with its neighbours (all pages of a launch are placed after those groups were cleared) */
  int32_t plain_size;       /* 0: an FLE page.  1 / 2 / 4 / 8: a PLAIN page (a column that overflowed its dictionary,
                             * hdfs-parquet-table-writer.cc:276-279) -- src_dev = its values back to back, plain_size
                             * bytes each, any alignment (the non-null ones only with expand = 1); dst_dev = the slot of
                             * the page's first row in a row-aligned VALUE array (column base + first row * plain_size,
                             * base 8-byte aligned); NULL rows get 0.  Bit widths and dst_bit are ignored. */
  int32_t ndict;
  const void* dict_dev;     /* with plain_size != 0: the page IS a dictionary page (src_bit_width, FLE groups) of a
                             * column whose later pages are PLAIN -- the writer starts every row group with a
                             * dictionary and falls back to PLAIN after 40000 distinct values
                             * (hdfs-parquet-table-writer.cc:276-279) -- and its rows are decoded into the same value
                             * array: value = dict_dev[code] (ndict entries of plain_size bytes; a code >= ndict gives 0) */
} fle_page_ref;
/* Re-bases the FLE sections of all pages (of all columns) of a row group onto aligned plane arrays
 * in ONE launch (two when a page starts or ends inside a group) -- consecutive pages of a column go to
 * consecutive row ranges of dst, so the column becomes one plane array the fused kernels can scan -- and counts
 * the rows whose definition level is 0 into *null_rows_dev (optional).  refs is a HOST array. */
int fle_pages_gather_dev(fle_ctx* ctx, const fle_page_ref* refs, int nrefs, uint64_t* null_rows_dev);

/* One materialised column of the survivors' tuples. */
typedef struct fle_slot_desc {
  const void* planes_dev;  /* FLE groups, 8-byte aligned, rows aligned with the bitmap */
  uint64_t bytes;
  int32_t bit_width;
  int32_t slot_size;       /* 1, 2, 4 or 8 bytes: sizeof(T) of ColumnReader<T> */
  const void* dict_dev;    /* ndict dictionary values of slot_size bytes (DictDecoder<T>::dict_) */
  int32_t ndict;
  int32_t tuple_offset;    /* byte offset of the slot inside a tuple */
  const void* def_planes_dev; /* optional: the column's definition levels as 1-bit FLE planes aligned with the rows
                               * (8-byte aligned); a selected row whose level is 0 gets tuple[null_offset] = 1 and its
                               * slot stays zero (ReadSlot's tuple->SetNull, hdfs-parquet-scanner.cc:515-531) */
  int32_t null_offset;     /* byte offset of the column's NULL indicator inside a tuple */
  int32_t reserved;
  const void* plain_values_dev; /* optional: a PLAIN column -- row-aligned values of slot_size bytes (fle_pages_gather_dev,
                                 * plain_size), copied into the slot as they are (ParquetPlainEncoder::Decode);
                                 * planes_dev / dict_dev are not used then */
} fle_slot_desc;
/* Replaces the materialisation loop of AssembleRows (hdfs-parquet-scanner.cc:1134-1180: skip-distance
 * list, then ReadValue(skip) -> DictDecoder<T>::GetValue -> FleDecoder::Get(T*, skip) per survivor and
 * column) for rows [0, nrows) of a bitmap: tuple j (tuple_size bytes, zero-filled, then one slot per
 * column holding the dictionary value of the j-th selected row) is written to tuples_dev for
 * j < min(selected, cap_rows), in ascending row order; *count_dev = selected rows; *bad_codes_dev =
 * codes >= ndict met on the way (GetValue() == false; their slots stay zero).  One launch. */
int fle_assemble_rows_dev(fle_ctx* ctx, const uint64_t* bitmap_dev, int64_t nrows, const fle_slot_desc* cols,
                          int ncols, int tuple_size, uint8_t* tuples_dev, int64_t cap_rows, uint64_t* count_dev,
                          uint64_t* bad_codes_dev);

/* fle_eval_dev + fle_sum_product_dev in ONE pass (the whole of TPC-H Q6 below the final
 * all-reduce): evaluates the program over rows [first_row, first_row + n) and accumulates
 *   *sum_dev = SUM over selected rows r of plain_i64_dev[r - first_row] * dict_i64_dev[code(r)]
 * where code(r) is the row's code in cols_dev[code_col] (exact int64; a code >= ndict counts 0),
 * *count_dev (optional) = number of selected rows, bitmap_out_dev (optional, may be NULL) as
 * fle_eval_dev.  Runs on a kernel specialised for the program shape ("eval_jit"): the survivors'
 * codes are picked out of the planes the predicate has just read from shared memory, so the
 * aggregate costs one 8-byte load per survivor instead of a second pass.  Without a usable NVRTC
 * (or with eval_jit = 0) the same result comes from the two separate kernels.  Columns must be
 * non-nullable. */
int fle_eval_sum_product_dev(fle_ctx* ctx, const fle_col_desc* cols_dev, int ncols,
                             const fle_pred_node* postfix, int nnodes, int64_t first_row, int64_t n,
                             const int64_t* plain_i64_dev, int code_col, const int64_t* dict_i64_dev,
                             int ndict, uint64_t* bitmap_out_dev, int64_t* sum_dev,
                             uint64_t* count_dev);

/* ---- row groups partitioned across the GPUs of one box -------------------------- */
/* The reference runs one HdfsParquetScanner per scanner thread over its own scan ranges
 * (hdfs-parquet-scanner.cc:1040) and its row groups are independent (:1056-1060); here one
 * context per GPU scans a contiguous range of row groups and the per-query totals (match count,
 * one exact int64 partial aggregate) are summed over the GPUs INSIDE the scan kernel: its last
 * thread block posts the GPU's totals into every peer's mailbox with peer stores over
 * NVLink/NVSwitch and adds up the peers' entries -- no separate collective, no host in the loop.
 * Every context of a group must issue the same sequence of fle_*_group_dev calls (like any
 * collective); a peer that does not arrive within "group_timeout_ms" (context option, default
 * 10000) is reported by fle_group_status and its contribution taken as 0.
 *
 * One process, several GPUs (or several contexts on one GPU -- then run every member's kernels once before the
 * first group call: with "group_deferred" = 0 a member's kernel waits for its peers' kernels, and the CUDA runtime
 * loads a kernel lazily at its first launch, which may need the device idle; members on different devices are not
 * coupled that way): */
int fle_group_create_local(fle_ctx* const* ctxs, int n);
/* One process per GPU: every rank exports a 64-byte handle of its mailbox (CUDA IPC), the caller
 * all-gathers the handles with its own transport (MPI, torch.distributed, a socket) and every rank
 * opens the group with all of them (handles = world x FLE_GROUP_HANDLE_BYTES bytes, rank order). */
int fle_group_export(fle_ctx* ctx, void* handle_out);
int fle_group_open(fle_ctx* ctx, int rank, int world, const void* handles);
int fle_group_close(fle_ctx* ctx);
/* Context option "group_deferred" (default 0).  0: a group call's totals are in place when its kernel
 * ends (one NVLink round trip inside the kernel; the GPUs of the group run in lockstep).  1: the
 * kernel posts its totals and returns; they are collected -- and written to the pointers the call was
 * given -- by the context's NEXT group call or by fle_group_flush (enqueues a one-warp kernel), so a
 * stream of queries pays nothing for the exchange and the GPUs may drift apart by one query.  Read
 * the totals of call k only after call k+1 or fle_group_flush has been enqueued on the stream. */
int fle_group_flush(fle_ctx* ctx);
/* Synchronises the context's stream; any output may be NULL. */
int fle_group_status(fle_ctx* ctx, int* rank, int* world, uint64_t* reductions_done, uint64_t* peers_timed_out);
/* fle_eval_dev over this GPU's rows; *total_count_dev = matches of ALL GPUs of the group.
 * The bitmap stays local (it covers this GPU's rows). */
int fle_eval_group_dev(fle_ctx* ctx, const fle_col_desc* cols_dev, int ncols, const fle_pred_node* postfix,
                       int nnodes, int64_t first_row, int64_t n, uint64_t* bitmap_out_dev,
                       uint64_t* total_count_dev);
/* fle_eval_sum_product_dev over this GPU's rows; *total_sum_dev / *total_count_dev = the exact
 * int64 aggregate / match count over ALL GPUs of the group (TPC-H Q6 end to end in one kernel
 * per GPU). */
int fle_eval_sum_product_group_dev(fle_ctx* ctx, const fle_col_desc* cols_dev, int ncols,
                                   const fle_pred_node* postfix, int nnodes, int64_t first_row, int64_t n,
                                   const int64_t* plain_i64_dev, int code_col, const int64_t* dict_i64_dev,
                                   int ndict, uint64_t* bitmap_out_dev, int64_t* total_sum_dev,
                                   uint64_t* total_count_dev);

/* ---- page framing and PLAIN-encoded columns ("next" rows) ------------- */
/* Byte layout of one uncompressed data-page body (hdfs-parquet-table-writer.cc:567-575,
 * hdfs-parquet-scanner.cc:882-919, dict-encoding.h:185-192):
 *   max_def_level > 0 : [int32 num_def_bytes][def levels, FLE bw 1][u8 bit_width][FLE groups]
 *   max_def_level == 0:                                          [u8 bit_width][FLE groups]   */
typedef struct fle_page_layout {
  uint32_t def_off, def_bytes;   /* 0, 0 for a required column */
  uint32_t data_off, data_bytes; /* FLE groups (after the bit-width byte) */
  int32_t bit_width;             /* the header byte */
  int32_t reserved;
} fle_page_layout;
/* Host-side parse of the two header fields (reads 5 bytes); FLE_E_BUFFER if the page is too
 * short for what its header promises. */
int fle_page_parse(const uint8_t* page_host, size_t size, int max_def_level, fle_page_layout* out);
/* Copies nbytes from an arbitrarily aligned DEVICE address to a 16-byte aligned device buffer
 * with a funnel-shift kernel (the FLE groups of a page start at an odd offset of the file
 * image); whole 16-byte chunks are written, bytes past nbytes as 0, so dst needs
 * (nbytes + 15) / 16 * 16 bytes.  The source is read in aligned 16-byte chunks: up to 15 bytes
 * before src_dev and after src_dev + nbytes are touched (inside any cudaMalloc'ed buffer). */
int fle_realign_dev(fle_ctx* ctx, const uint8_t* src_dev, size_t nbytes, uint8_t* dst_dev);

/* Element types of PLAIN fixed-width columns (parquet-common.h:39-76). */
enum { FLE_T_INT32 = 0, FLE_T_INT64 = 1, FLE_T_FLOAT = 2, FLE_T_DOUBLE = 3 };
/* Replaces ParquetPlainEncoder::Eq/Lt/Le/Gt/Ge/In<T> (parquet-common.h:197-255) for n values
 * stored back to back at values_dev (naturally aligned).  literals: HOST array of nlits
 * elements of the column type (1 unless op == FLE_OP_IN, at most 1024).  The predicate is
 * `value OP literal`; reference_operand_order != 0 evaluates the reference's generic template
 * literally, `literal OP value` (parquet-common.h:214,225,236,247).  Output as fle_eval_dev. */
int fle_plain_eval_dev(fle_ctx* ctx, const void* values_dev, int elem_type, int64_t n, int op,
                       const void* literals, int nlits, int reference_operand_order,
                       uint64_t* bitmap_out_dev, uint64_t* match_count_dev);

/* ---- dictionary build on the device ("next" row) ----------------------------- */
/* Replaces DictEncoder<T>::Put x n + WriteDict + WriteData (dict-encoding.h:234-274, 393-447)
 * for a whole INT32/INT64 column chunk resident on the device: deduplicates the values, SORTS
 * the dictionary (order-preserving codes), remaps and FLE-packs the codes with
 * bit_width = DictEncoderBase::bit_width() (dict-encoding.h:76-80).
 * dict_out_dev: room for max_entries int64 values (INT32 columns are widened);
 * *num_entries / *bit_width / *enc_len are HOST outputs (the call synchronises).
 * FLE_E_DICT_FULL when the column has more than max_entries (<= 40000, dict-encoding.h:157)
 * distinct values -- the writer then falls back to PLAIN (hdfs-parquet-table-writer.cc:276-279). */
int fle_dict_encode_dev(fle_ctx* ctx, const void* values_dev, int elem_type, int64_t n,
                        int max_entries, int64_t* dict_out_dev, int* num_entries, int* bit_width,
                        uint8_t* enc_dev, size_t enc_cap, size_t* enc_len);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* FLE_B200_H */
