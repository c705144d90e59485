/* TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, scalar) of the reference's FLE hot path, used by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg as the CHECKER for the CUDA
 * path.  Nothing under impala-avx2-parquet-scanner_b200/ may include, link or call it.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle_vs_ref.py,
 * tests/test_golden.py) against (1) the reference's own code, compiled unmodified from
 * /root/reference into oracle/_ref/libfle_ref.so, and (2) the golden vectors generated
 * from it and committed under tests/golden/, which include the known answers the
 * reference's fle-test.cc holds (encoded lengths :219,:224,:238; full-word byte images
 * :217-223,:235-236).
 *
 * Layout restated (fle-encoding.h:8338-8340, :426-430; fle-benchmark.cc:395-411,458-474):
 *   64 values form a group; a group of bit width bw is bw consecutive little-endian
 *   uint64 words; word i holds bit i of all 64 values; value k of the group sits at bit
 *   position 63-k of every word.
 * Bitmaps follow boost::dynamic_bitset<unsigned long>: row r -> bit (r & 63) of word r>>6.
 */
#ifndef FLE_ORACLE_H
#define FLE_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_EQ = 0, ORC_LT = 1, ORC_LE = 2, ORC_GT = 3, ORC_GE = 4, ORC_IN = 5,
       ORC_AND = 6, ORC_OR = 7, ORC_ALL = 8, ORC_NONE = 9 };

/* FleEncoder::Flush byte length: Put packs a group lazily, Flush always packs one more
 * (fle-encoding.h:8317-8327, :9806-9811) => max(1, ceil(n/64)) groups. */
int64_t orc_encoded_bytes(int64_t n, int bw);

/* FleEncoder::Put x n + Flush (fle-encoding.h:8315-8365, Pack_* :8367-9803).
 * Unfilled slots of the last group are written as 0 here; the reference leaves stale
 * staging data there, so parity on a ragged tail is defined on the first
 * n rows only.  Returns bytes written. */
int64_t orc_pack(const uint32_t* vals, int64_t n, int bw, uint64_t* out_words);

/* FleDecoder::Get x n starting at row first_row (fle-encoding.h:404-567, Unpack_* :569-7329). */
void orc_unpack(const uint64_t* words, int bw, int64_t first_row, int64_t n, uint32_t* out);

/* FleDecoder::Eq/Lt/Le/Gt/Ge (fle-encoding.h:7962-8234) and In (:8236-8313) for rows
 * [first_row, first_row + num_rows).  The reference splits a call into a scalar head over
 * the already-unpacked group and a plane-algebra body; both compute the same unsigned
 * compare when the literal is < 2^bw (its DCHECK'd precondition), which is what is
 * restated: body semantics, i.e. only the low bw bits of the literal are used.
 * Output: num_rows bits, LSB-first words, bits beyond num_rows zero.  Returns popcount. */
int64_t orc_pred(int op, const uint64_t* words, int bw, int64_t first_row, int64_t num_rows,
                 const uint64_t* lits, int nlits, uint64_t* out_words);

/* Same predicate computed the way the reference body does (plane algebra MSB->LSB with
 * Meq/Mlt/Mgt masks and the 6-step bit reversal, fle-encoding.h:8033-8065) for whole
 * groups; first_row must be a multiple of 64.  Used to cross-check orc_pred. */
int64_t orc_pred_planes(int op, const uint64_t* words, int bw, int64_t first_row,
                        int64_t num_rows, const uint64_t* lits, int nlits, uint64_t* out_words);

/* ColumnReader::IntersectBitset (hdfs-parquet-scanner.cc:326-331): the i-th set bit of
 * root takes sub[i].  root has n_root bits; modified in place.  Returns popcount. */
int64_t orc_intersect_bitset(uint64_t* root_words, int64_t n_root, const uint64_t* sub_words);

/* Cursor-faithful decoder model (fle-encoding.h:344-402): count_ in [0,64], buffer_end_
 * as a group index.  Used to prove the window/Skip loop equals whole-column slicing. */
typedef struct {
  const uint64_t* words;
  int64_t ngroups_guard; /* groups before buffer_end_guard_ */
  int bw;
  int count;             /* count_ : 64 = nothing unpacked */
  int64_t next_group;    /* buffer_end_ as a group index */
  int64_t cur_group;     /* group currently unpacked into current_value_ (-1 none) */
} orc_decoder;
void orc_decoder_init(orc_decoder* d, const uint64_t* words, int64_t nbytes, int bw);
int orc_decoder_get(orc_decoder* d, uint32_t* val);
int orc_decoder_get_skip(orc_decoder* d, uint32_t* val, int skip_rows);
int orc_decoder_skip(orc_decoder* d, int skip_rows);
/* predicate at the cursor without advancing it (head + body, fle-encoding.h:7973-8009);
 * appends num_rows bits at bit offset *nbits of out_words. */
void orc_decoder_pred(const orc_decoder* d, int op, int64_t num_rows, const uint64_t* lits,
                      int nlits, uint64_t* out_words, int64_t* nbits);

/* DictDecoder<T>::Eq/Gt/Lt/Ge/Le literal -> dictionary-code translation against a sorted
 * dictionary (dict-encoding.h:461-521).  Returns the FLE op to run (ORC_EQ/LT/GE) with
 * *code set, or ORC_ALL / ORC_NONE for the short-circuits. */
int orc_dict_translate(int op, const int64_t* dict, int64_t ndict, int64_t literal, uint64_t* code);
/* DictDecoder<T>::In (dict-encoding.h:523-541): keeps present literals' codes, in input
 * order; returns how many (0 => all-false). */
int64_t orc_dict_translate_in(const int64_t* dict, int64_t ndict, const int64_t* lits,
                              int64_t nlits, uint64_t* codes);
/* DictEncoderBase::bit_width (dict-encoding.h:76-80) with BitUtil::Log2 (bit-util.h:132-143). */
int orc_dict_bit_width(int64_t num_entries);

#ifdef __cplusplus
}
#endif
#endif
