// Host-side mirror of the column-reader batch interface of the reference's Parquet scanner
// (exec/hdfs-parquet-scanner.{h,cc}) -- only the part that sits on the FLE hot path:
//
//   BaseColumnReader / ColumnReader<T>      hdfs-parquet-scanner.cc:142-302, :304-560
//     ReadDataPage / InitDataPage           :730-924, :448-472   (page body framing only)
//     Eq/Lt/Le/Gt/Ge/In + IntersectBitset   :326-445
//     ReadValue(skip) / SkipValue / ReadSlot(skip) / SkipSlot   :1006-1035, :515-547
//     (required columns: the reference's ReadDefinitionLevel returns 1 and then compares it
//      with max_def_level() == 0, which would turn every value into NULL; here every row of
//      a required column is defined.  Impala's own writer only produces nullable columns.)
//   HdfsParquetScanner::Eq<T>..In<T>        hdfs-parquet-scanner.h:91-102, .cc:1867-1907
//   EvalSimplePredicates                    .cc:1837-1865
//   AssembleRows, simple-predicate branch   .cc:1101-1181
//
// Everything the reference gets from the rest of Impala (thrift page headers, DiskIoMgr
// streams, MemPool, Tuple/RowBatch, RuntimeState) is out of scope: pages are handed in as
// in-memory uncompressed page bodies and tuples are rows of a flat byte buffer.
//
// Page body layout (hdfs-parquet-table-writer.cc:567-575, hdfs-parquet-scanner.cc:882-919,
// dict-encoding.h:185-192):
//     nullable column:  [int32 num_def_bytes][def levels: FLE, bit width 1][u8 bw][FLE groups]
//     required column:  [u8 bw][FLE groups]
//
// Three evaluation paths, same results:
//   * row-group batch (default when it applies): ALL data pages of every column chunk are uploaded --
//     one copy per column chunk when its pages lie back to back in (pinned) host memory, on the
//     context's copy stream, so row group k+1 can cross PCIe while k is evaluated
//     (StartRowGroupUpload) -- then fle_pages_gather_dev places the pages' FLE sections into one ROW-ALIGNED
//     plane array per column: pages of any row count, cut differently per column, are placed at their row
//     offset (dst_bit); pages of nullable columns hold only their non-null values and are expanded through
//     their definition levels (code 0 at NULL rows), which become a 1-bit column of their own; NULLs are
//     counted on the way.  ONE fused fle_eval_dev covers the row group -- the plain program, or, when the
//     row group holds NULLs, the program with `AND def_c = 1` per nullable column / leaf -- and
//     fle_assemble_rows_dev writes the survivors' tuples, NULL indicators included; two host synchronisations
//     per row group however many pages it holds.  Applies to fresh readers whose predicate columns are
//     dictionary-encoded; a column with PLAIN pages (it overflowed its dictionary: dictionary pages first, PLAIN
//     pages from there on) rides along when it is only materialised: its values -- decoded through the dictionary
//     or copied -- are placed into a row-aligned value array the tuples are copied from;
//   * per leaf (the reference's structure): every leaf call goes through ColumnReader<T> ->
//     DictDecoder<T> -> FleDecoder, which evaluates the WHOLE page on the GPU on first use and
//     then serves the <=1024-row windows from the cached bitmap;
//   * fused (default): the whole SimplePredicate forest is lowered to one postfix program and
//     ONE fle_eval call covers every column for the span of rows during which no reader
//     changes page; windows are bit slices of that result.  Used whenever the span starts on
//     a 64-row group boundary of every page involved (always true for page-aligned row
//     groups); otherwise the per-leaf path takes over for that window.
#ifndef IMPALA_EXEC_HDFS_PARQUET_SCANNER_H
#define IMPALA_EXEC_HDFS_PARQUET_SCANNER_H

#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <memory>
#include <string>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <boost/dynamic_bitset.hpp>

#include "exprs/simple-predicates.h"
#include "util/dict-encoding.h"
#include "util/fle-encoding.h"

namespace impala {

class Status {
 public:
  Status() : ok_(true) {}
  explicit Status(const std::string& msg) : ok_(false), msg_(msg) {}
  bool ok() const { return ok_; }
  const std::string& GetErrorMsg() const { return msg_; }
  static Status OK() { return Status(); }
 private:
  bool ok_;
  std::string msg_;
};

// parquet.thrift:206-224 (the reference adds the two non-standard FLE encodings)
namespace parquet_encoding {
enum type { PLAIN = 0, PLAIN_DICTIONARY = 2, RLE = 3, BIT_PACKED = 4, FLE = 9, FLE_DICTIONARY = 10 };
}

class HdfsParquetScanner {
 public:
  struct DataPage {
    const uint8_t* data;  // uncompressed page body
    int size;
    int num_values;       // data_page_header.num_values (rows, NULLs included)
    int encoding;         // parquet_encoding::FLE_DICTIONARY
    DataPage() : data(NULL), size(0), num_values(0), encoding(parquet_encoding::FLE_DICTIONARY) {}
    DataPage(const uint8_t* d, int s, int n) : data(d), size(s), num_values(n),
        encoding(parquet_encoding::FLE_DICTIONARY) {}
  };

  class BaseColumnReader;
  template<typename T> class ColumnReader;

  HdfsParquetScanner() : tuple_size_(0), fused_enabled_(true), fused_pos_(0), fused_rows_(0),
      fused_evals_(0), leaf_windows_(0), device_materialise_(true), device_spans_(0),
      row_group_batch_(true), batch_state_(0), batch_bank_(0), batch_rows_(0), row_group_batches_(0),
      batch_eval_cols_(0), batch_ncols_nulls_(0), batch_nulls_seen_(false), row_group_null_batches_(0) {}
  ~HdfsParquetScanner();

  // CreateColumnReaders/CreateReader (:1410-1451, :670-720) + InitColumns (:1453-1531) for
  // one column chunk: its sorted PLAIN dictionary page and its data pages.  Returns the
  // column-reader index used by the predicate leaves (slot_idx, scalar-fn-call.cc:940-942).
  template<typename T>
  int AddColumn(int max_def_level, const uint8_t* dict_page, int dict_len,
                const std::vector<DataPage>& pages);

  // CreateSimplePredicates (:1825-1835): one root per conjunct; the scanner ANDs them.
  void AddSimplePredicate(SimplePredicate* root) { simple_predicates_.push_back(root); }

  // typed trampolines (hdfs-parquet-scanner.h:91-102, .cc:1867-1907)
  template<typename T> void Eq(int idx, int64_t num_rows, boost::dynamic_bitset<>& skip_bitset, T& val);
  template<typename T> void Lt(int idx, int64_t num_rows, boost::dynamic_bitset<>& skip_bitset, T& val);
  template<typename T> void Le(int idx, int64_t num_rows, boost::dynamic_bitset<>& skip_bitset, T& val);
  template<typename T> void Gt(int idx, int64_t num_rows, boost::dynamic_bitset<>& skip_bitset, T& val);
  template<typename T> void Ge(int idx, int64_t num_rows, boost::dynamic_bitset<>& skip_bitset, T& val);
  template<typename T> void In(int idx, int64_t num_rows, boost::dynamic_bitset<>& skip_bitset, std::vector<T>& vals);

  // :1837-1865.  Appends min(1024, rows left in every reader's page) bits to skip_bitset; a
  // set bit means the row PASSES (:1139-1140).  false = no more data or parse_status_ not ok.
  bool EvalSimplePredicates(boost::dynamic_bitset<>& skip_bitset);

  // Simple-predicate branch of AssembleRows (:1101-1181): evaluates window after window and
  // materialises the surviving rows, column by column, with ReadValue(skip).  Tuples are
  // appended to `rows` (tuple_size() bytes each: one null-indicator byte per column, then the
  // slots in AddColumn order).  Returns the number of rows that passed.
  Status AssembleRows(int64_t expected_rows_in_group, std::vector<uint8_t>* rows, int64_t* num_passed);
  // The same into a tuple buffer of the caller's (the reference materialises into the row batch's tuple memory,
  // hdfs-parquet-scanner.cc:1151-1168): when that buffer is pinned (fle_host_alloc / fle_host_register) the
  // survivors' tuples of a row-group batch are copied device -> buffer directly -- no staging copy, which is the
  // largest single item of a scanner thread's time per row group.  *bytes = tuple bytes written (num_passed *
  // tuple_size()); a buffer that is too small makes the call fail.
  Status AssembleRowsInto(int64_t expected_rows_in_group, uint8_t* tuple_buf, size_t capacity, int64_t* num_passed, size_t* bytes);

  int tuple_size() const { return tuple_size_; }
  int slot_offset(int idx) const;
  int null_offset(int idx) const { return idx; }
  int num_columns() const { return static_cast<int>(column_readers_.size()); }
  const Status& parse_status() const { return parse_status_; }

  // false: force the reference's one-pass-per-leaf structure (still on the GPU).
  void set_fused(bool on) { fused_enabled_ = on; }
  // false: keep the survivors' materialisation on the host (ReadValue(skip) per survivor from a
  // decoded copy of the page) even where the device-resident page-span path applies.
  void set_device_materialise(bool on) { device_materialise_ = on; }
  int64_t fused_evals() const { return fused_evals_; }
  int64_t leaf_windows() const { return leaf_windows_; }
  int64_t device_spans() const { return device_spans_; }
  // false: never take the whole-row-group batch path.
  void set_row_group_batch(bool on) { row_group_batch_ = on; }
  int64_t row_group_batches() const { return row_group_batches_; }
  // Extension: begins the upload of this row group's column chunks (asynchronously, on the calling
  // thread's copy stream) so that it overlaps the AssembleRows of the previous row group.  Optional:
  // AssembleRows calls it itself.  Call after the last AddColumn/AddSimplePredicate.  Returns false when the
  // row group does not qualify for the batch path (nothing is uploaded then).
  bool StartRowGroupUpload();

  BaseColumnReader* column_reader(int idx) { return column_readers_[idx]; }

 private:
  bool FusedWindow(int64_t limit_rows, const std::vector<int64_t>& page_pos,
                   boost::dynamic_bitset<>& skip_bitset);
  bool BuildFusedSpan(const std::vector<int64_t>& page_pos);
  bool LowerPredicates(FusedProgram* prog);
  // Whole-page span on the device: upload the pages' FLE sections, fused evaluation, survivor
  // row ids (fle_bitmap_select_dev) and their codes per column (fle_gather_dev); only the
  // survivors' codes come back.  Returns false (nothing consumed) when the span does not qualify.
  bool DevicePageSpan(std::vector<uint8_t>* rows, int64_t* num_passed, int64_t* span_rows);
  void* DeviceBuffer(size_t slot, size_t bytes);
  bool DeviceRowGroup(std::vector<uint8_t>* rows, int64_t* num_passed, int64_t* span_rows);

  std::vector<BaseColumnReader*> column_readers_;
  std::vector<SimplePredicate*> simple_predicates_;
  Status parse_status_;
  int tuple_size_;

  bool fused_enabled_;
  std::vector<uint64_t> fused_bitmap_;  // result of the last fused evaluation
  std::vector<int64_t> fused_start_;    // every reader's row_position() when it was built
  int64_t fused_pos_, fused_rows_;      // (unused) / rows covered
  int64_t fused_evals_, leaf_windows_;
  bool device_materialise_;
  int64_t device_spans_;
  std::vector<std::pair<void*, size_t> > dev_buffers_;  // grow-only device scratch, freed with the scanner
  std::vector<uint32_t> host_codes_;

  // row-group batch: what StartRowGroupUpload prepared
  struct BatchColumn {
    void* image_dev;          // the uploaded column-chunk image(s)
    void* planes_dev;         // groups * bit_width * 8 bytes, filled by fle_pages_gather_dev
    void* dict_dev;
    void* def_dev;            // nullable columns: the pages' definition levels as one row-aligned 1-bit plane array
    int bit_width;            // widest page
    int ndict;
    bool plain;               // every page PLAIN (the column overflowed its dictionary): planes_dev is a row-aligned value array
    int eval_col;             // the column's index among the columns the program is evaluated over (-1: a PLAIN column)
  };
  bool row_group_batch_;
  int batch_state_;           // 0 not tried, 1 uploaded, -1 does not qualify
  int batch_bank_;
  int64_t batch_rows_;
  int64_t row_group_batches_;
  std::vector<BatchColumn> batch_cols_;
  std::vector<fle_page_ref> batch_refs_;
  FusedProgram batch_prog_;
  // the same program for a row group that DOES hold NULLs: every leaf on a nullable column is ANDed with the
  // column's definition levels (a 1-bit column of its own, index batch_def_col_[c]) -- NULL fails every leaf,
  // simple-predicates.h / hdfs-parquet-scanner.cc:342-345
  std::vector<fle_pred_node> batch_prog_nulls_;
  std::vector<int> batch_def_col_;   // per column: index of its definition-level column in the NULL program, -1 = none
  int batch_eval_cols_;              // dictionary columns = columns of the plain program
  int batch_ncols_nulls_;            // columns of the NULL program (0: it does not fit, NULL row groups take the page paths)
  // what EnqueueRowGroup set up: the row group's kernels are enqueued by StartRowGroupUpload already (on the bank's own
  // context), so that they run as soon as its column chunks have arrived -- while the host still fetches the previous
  // row group's tuples
  std::vector<fle_col_desc> batch_eval_desc_;
  std::vector<fle_slot_desc> batch_slots_;
  uint64_t* batch_bitmap_dev_ = NULL;
  uint64_t* batch_hdr_dev_ = NULL;   // [0] NULL rows, [1] survivors, [2] bad codes
  uint64_t* batch_hdr_ = NULL;       // pinned copy
  uint8_t* batch_tuples_dev_ = NULL;
  bool batch_enqueued_ = false, batch_enqueued_nulls_ = false;
  bool batch_placed_ = false;        // the pages were placed and the descriptors built (a later evaluation reuses them)
  bool EnqueueRowGroup(bool with_nulls, bool first);
  uint8_t* direct_out_ = NULL;       // AssembleRowsInto: the caller's tuple buffer
  size_t direct_cap_ = 0, direct_used_ = 0;
  bool batch_nulls_seen_;            // an earlier row group of this scanner held NULLs: start with the NULL program
  int64_t row_group_null_batches_;
 public:
  int64_t row_group_null_batches() const { return row_group_null_batches_; }
  // What one row group taught about the file travels to the scanner of the next one: a scanner that expects NULLs
  // evaluates the NULL program at once instead of finding out from the counts of the plain one.
  bool saw_nulls() const { return batch_nulls_seen_; }
  void set_expect_nulls(bool v) { batch_nulls_seen_ = v; }
};

// ---------------------------------------------------------------------------------------
class HdfsParquetScanner::BaseColumnReader {
 public:
  BaseColumnReader(HdfsParquetScanner* parent, int max_def_level, int slot_size)
    : parent_(parent), max_def_level_(max_def_level), slot_size_(slot_size), tuple_offset_(0),
      null_offset_(0), next_page_(0), num_buffered_values_(0), page_num_values_(0),
      data_(NULL), data_size_(0), def_data_(NULL), def_bytes_(0), dict_decoder_base_(NULL) {}
  virtual ~BaseColumnReader() {}

  int max_def_level() const { return max_def_level_; }
  int64_t num_buffered_values() const { return num_buffered_values_; }
  void num_buffered_values(int64_t fetched_rows) { num_buffered_values_ -= fetched_rows; }  // :179-181
  int64_t page_num_values() const { return page_num_values_; }
  // Rows of the current page already consumed by ReadValue/SkipValue, i.e. the cursor the
  // per-leaf predicates would start from (predicates never advance it).
  int64_t row_position() const {
    if (max_def_level() > 0) return fle_def_levels_.get() ? fle_def_levels_->position() : 0;
    return (dict_decoder_base_ && dict_decoder_base_->data_decoder())
        ? dict_decoder_base_->data_decoder()->position() : 0;
  }

  // :730-924 reduced to the page-body framing (see file comment).
  Status ReadDataPage();

  // :1006-1035
  bool ReadValue(uint8_t* tuple, int skip_rows);
  bool SkipValue(int skip_rows);

  // fused-path accessors: the current page's encoded streams
  const uint8_t* fle_data() const { return data_ + 1; }     // after the bit-width byte
  int fle_data_len() const { return data_size_ - 1; }
  int fle_bit_width() const { return data_size_ > 0 ? data_[0] : 0; }
  const uint8_t* def_data() const { return def_data_; }
  int def_bytes() const { return def_bytes_; }
  bool page_is_dictionary() const { return page_encoding_ == parquet_encoding::FLE_DICTIONARY ||
                                           page_encoding_ == parquet_encoding::PLAIN_DICTIONARY; }
  // row-group batch accessors
  const std::vector<DataPage>& pages() const { return pages_; }
  bool fresh() const { return next_page_ == 0 && num_buffered_values_ == 0; }
  void ConsumeAllPages() { next_page_ = pages_.size(); num_buffered_values_ = 0; }
  virtual const void* dict_values(int* n, int* elem_size) const = 0;

 protected:
  friend class HdfsParquetScanner;
  virtual Status InitDataPage(uint8_t* data, int size) = 0;
  virtual bool ReadSlot(void* slot, int skip_rows) = 0;
  virtual bool SkipSlot(int skip_rows) = 0;
  virtual bool SlotFromCode(void* slot, uint32_t code) = 0;  // dictionary lookup of an already decoded code
  int ReadDefinitionLevel();

  HdfsParquetScanner* parent_;
  int max_def_level_;
  int slot_size_;
  int tuple_offset_, null_offset_;
  std::vector<DataPage> pages_;
  size_t next_page_;
  int64_t num_buffered_values_;
  int64_t page_num_values_;
  int page_encoding_;
  uint8_t* data_;
  int data_size_;
  const uint8_t* def_data_;
  int def_bytes_;
  std::unique_ptr<FleDecoder> fle_def_levels_;
  DictDecoderBase* dict_decoder_base_;
};

template<typename T>
class HdfsParquetScanner::ColumnReader : public HdfsParquetScanner::BaseColumnReader {
 public:
  ColumnReader(HdfsParquetScanner* parent, int max_def_level)
    : BaseColumnReader(parent, max_def_level, sizeof(T)) {}

  void CreateDictionaryDecoder(uint8_t* values, int size) {  // :448-451
    dict_decoder_.reset(new DictDecoder<T>(values, size, sizeof(T)));
    dict_decoder_base_ = dict_decoder_.get();
  }
  DictDecoder<T>* dict_decoder() { return dict_decoder_.get(); }

  // :326-331 -- on the device: prefix popcount + bit deposit (fle_intersect_bitset)
  void IntersectBitset(boost::dynamic_bitset<>& root_bitset, boost::dynamic_bitset<>& sub_bitset) {
    const size_t n_root = root_bitset.size(), n_sub = sub_bitset.size();
    if (n_root == 0) return;
    std::vector<uint64_t> root((n_root + 63) / 64), sub((n_sub + 63) / 64 + 1);
    boost::to_block_range(root_bitset, root.begin());
    boost::to_block_range(sub_bitset, sub.begin());
    fle_ctx* ctx = FleGpu::ctx();
    if (ctx == NULL) return;
    if (FleGpu::Check(fle_intersect_bitset(ctx, &root[0], n_root, &sub[0], n_sub), "fle_intersect_bitset") != FLE_OK) return;
    root_bitset.clear();
    root_bitset.append(root.begin(), root.end());
    root_bitset.resize(n_root);
  }

  // PLAIN data pages (columns that overflowed the dictionary, hdfs-parquet-table-writer.cc:276-279):
  // ParquetPlainEncoder::Eq..In<T> (parquet-common.h:197-255) compare num_rows values from the read
  // cursor on.  Here the page's values are uploaded once, the whole page is evaluated per (op, literal)
  // by fle_plain_eval_dev and windows are bit slices of that result.  Default: `value OP literal`, In()
  // implemented, NULLs handled like the dictionary path.  FleGpu::reference_compat(): `literal OP value`
  // (the reference's generic templates), In() appends nothing (its empty stub), no NULL handling.
  void PlainOp(int op, int64_t num_rows, boost::dynamic_bitset<>& skip_bitset, const T* lits, int nlits) {
    const bool compat = FleGpu::reference_compat();
    if (compat && op == FLE_OP_IN) return;
    const int et = PlainElemType();
    fle_ctx* ctx = FleGpu::ctx();
    if (et < 0 || ctx == NULL) {
      if (et < 0) parent_->parse_status_ = Status("PLAIN predicates: unsupported slot type");
      skip_bitset.resize(skip_bitset.size() + num_rows, false);
      return;
    }
    boost::dynamic_bitset<> data_bitset;
    boost::dynamic_bitset<>* out = &skip_bitset;
    int64_t data_rows = num_rows;
    if (!compat && max_def_level() > 0) {
      fle_def_levels_->Eq(num_rows, skip_bitset, max_def_level());
      data_rows = static_cast<int64_t>(skip_bitset.count());
      out = &data_bitset;
    }
    const std::vector<uint64_t>* bm = PlainBitmap(ctx, et, op, lits, nlits, compat);
    // the read cursor: values of this page already consumed by ReadSlot/SkipSlot
    int64_t pos = plain_pos_;
    for (int64_t i = 0; i < data_rows; ++i, ++pos) {
      const bool bit = bm != NULL && pos < plain_values_ && (((*bm)[pos >> 6] >> (pos & 63)) & 1);
      out->push_back(bit);
    }
    if (out == &data_bitset) IntersectBitset(skip_bitset, data_bitset);
  }

#define FLE_COLUMN_READER_OP(Name, ArgT)                                                   \
  void Name(int64_t num_rows, boost::dynamic_bitset<>& skip_bitset, ArgT val) {            \
    if (!page_is_dictionary()) { PlainCall##Name(num_rows, skip_bitset, val); return; }      \
    if (max_def_level() == 0) {                                                            \
      dict_decoder_->Name(num_rows, skip_bitset, val);                                     \
      return;                                                                              \
    }                                                                                      \
    fle_def_levels_->Eq(num_rows, skip_bitset, max_def_level());                           \
    boost::dynamic_bitset<> data_bitset;                                                   \
    dict_decoder_->Name(skip_bitset.count(), data_bitset, val);                            \
    IntersectBitset(skip_bitset, data_bitset);                                             \
  }
  FLE_COLUMN_READER_OP(Eq, T&)   // :333-350
  FLE_COLUMN_READER_OP(Lt, T&)   // :352-369
  FLE_COLUMN_READER_OP(Le, T&)   // :371-388
  FLE_COLUMN_READER_OP(Gt, T&)   // :390-407
  FLE_COLUMN_READER_OP(Ge, T&)   // :409-426
  FLE_COLUMN_READER_OP(In, std::vector<T>&)  // :428-445
#undef FLE_COLUMN_READER_OP

 protected:
  virtual Status InitDataPage(uint8_t* data, int size) {  // :453-472
    if (!page_is_dictionary()) {  // PLAIN: values back to back
      plain_data_ = data;
      plain_values_ = size / static_cast<int>(sizeof(T));
      plain_pos_ = 0;
      plain_cache_.clear();
      plain_dev_ = NULL;
      return Status::OK();
    }
    if (dict_decoder_.get() == NULL) return Status("File corrupt. Missing dictionary page.");
    dict_decoder_->SetData(data, size);
    return Status::OK();
  }
  virtual bool ReadSlot(void* slot, int skip_rows) {  // :515-531
    if (!page_is_dictionary()) {  // ParquetPlainEncoder::Decode<T>(data_, ..., skip_rows)
      plain_pos_ += skip_rows;
      if (plain_pos_ >= plain_values_) return false;
      memcpy(slot, plain_data_ + plain_pos_ * sizeof(T), sizeof(T));
      ++plain_pos_;
      return true;
    }
    T val;
    bool result = dict_decoder_->GetValue(&val, skip_rows);
    memcpy(slot, &val, sizeof(T));
    return result;
  }
  virtual bool SkipSlot(int skip_rows) {  // :533-547
    if (skip_rows <= 0) return true;
    if (!page_is_dictionary()) {
      plain_pos_ += skip_rows;
      return true;
    }
    return dict_decoder_->SkipValue(skip_rows);
  }
  virtual const void* dict_values(int* n, int* elem_size) const {
    *n = dict_decoder_.get() ? static_cast<int>(dict_decoder_->dict().size()) : 0;
    *elem_size = sizeof(T);
    return *n > 0 ? static_cast<const void*>(&dict_decoder_->dict()[0]) : NULL;
  }
  virtual bool SlotFromCode(void* slot, uint32_t code) {  // what GetValue does after Get(&index), dict-encoding.h:321-330
    const std::vector<T>& d = dict_decoder_->dict();
    if (code >= d.size()) return false;
    memcpy(slot, &d[code], sizeof(T));
    return true;
  }

 private:
  static int PlainElemType();
#define FLE_PLAIN_CALL(Name, OP)                                                                       \
  void PlainCall##Name(int64_t num_rows, boost::dynamic_bitset<>& bs, T& val) { PlainOp(OP, num_rows, bs, &val, 1); }
  FLE_PLAIN_CALL(Eq, FLE_OP_EQ)
  FLE_PLAIN_CALL(Lt, FLE_OP_LT)
  FLE_PLAIN_CALL(Le, FLE_OP_LE)
  FLE_PLAIN_CALL(Gt, FLE_OP_GT)
  FLE_PLAIN_CALL(Ge, FLE_OP_GE)
#undef FLE_PLAIN_CALL
  void PlainCallIn(int64_t num_rows, boost::dynamic_bitset<>& bs, std::vector<T>& vals) {
    PlainOp(FLE_OP_IN, num_rows, bs, vals.empty() ? NULL : &vals[0], static_cast<int>(vals.size()));
  }
  struct PlainCached {
    int op;
    bool swapped;
    std::vector<T> lits;
    std::vector<uint64_t> bitmap;
  };
  // whole-page bitmap for (op, literals): one upload per page, one kernel per distinct leaf
  const std::vector<uint64_t>* PlainBitmap(fle_ctx* ctx, int et, int op, const T* lits, int nlits, bool swapped) {
    for (size_t i = 0; i < plain_cache_.size(); ++i) {
      const PlainCached& c = plain_cache_[i];
      if (c.op == op && c.swapped == swapped && static_cast<int>(c.lits.size()) == nlits &&
          (nlits == 0 || memcmp(&c.lits[0], lits, sizeof(T) * nlits) == 0)) return &c.bitmap;
    }
    if (plain_values_ <= 0) return NULL;
    const size_t bytes = static_cast<size_t>(plain_values_) * sizeof(T);
    const size_t words = (static_cast<size_t>(plain_values_) + 63) / 64;
    if (plain_dev_ == NULL) {
      plain_dev_ = FleGpu::DeviceScratch(0, 60, bytes + 16);
      if (plain_dev_ == NULL) return NULL;
      if (FleGpu::Check(fle_h2d(ctx, plain_dev_, plain_data_, bytes), "fle_h2d(PLAIN page)") != FLE_OK) return NULL;
    }
    void* bm_dev = FleGpu::DeviceScratch(0, 61, words * 8 + 16);
    if (bm_dev == NULL) return NULL;
    int rc = fle_plain_eval_dev(ctx, plain_dev_, et, plain_values_, op, lits, nlits, swapped ? 1 : 0,
                                static_cast<uint64_t*>(bm_dev), NULL);
    if (FleGpu::Check(rc, "fle_plain_eval_dev") != FLE_OK) return NULL;
    plain_cache_.push_back(PlainCached());
    PlainCached& c = plain_cache_.back();
    c.op = op;
    c.swapped = swapped;
    c.lits.assign(lits, lits + nlits);
    c.bitmap.assign(words, 0);
    if (FleGpu::Check(fle_d2h(ctx, &c.bitmap[0], bm_dev, words * 8), "fle_d2h(PLAIN bitmap)") != FLE_OK ||
        FleGpu::Check(fle_ctx_sync(ctx), "fle_ctx_sync") != FLE_OK) {
      plain_cache_.pop_back();
      return NULL;
    }
    return &plain_cache_.back().bitmap;
  }

  std::unique_ptr<DictDecoder<T> > dict_decoder_;
  std::vector<uint8_t> dict_copy_;
  const uint8_t* plain_data_ = NULL;
  int64_t plain_values_ = 0, plain_pos_ = 0;
  void* plain_dev_ = NULL;
  std::vector<PlainCached> plain_cache_;
  friend class HdfsParquetScanner;
};

template<typename T> inline int HdfsParquetScanner::ColumnReader<T>::PlainElemType() { return -1; }
template<> inline int HdfsParquetScanner::ColumnReader<int32_t>::PlainElemType() { return FLE_T_INT32; }
template<> inline int HdfsParquetScanner::ColumnReader<int64_t>::PlainElemType() { return FLE_T_INT64; }
template<> inline int HdfsParquetScanner::ColumnReader<float>::PlainElemType() { return FLE_T_FLOAT; }
template<> inline int HdfsParquetScanner::ColumnReader<double>::PlainElemType() { return FLE_T_DOUBLE; }

// ---------------------------------------------------------------------------------------
inline HdfsParquetScanner::~HdfsParquetScanner() {
  for (size_t i = 0; i < column_readers_.size(); ++i) delete column_readers_[i];
  fle_ctx* ctx = dev_buffers_.empty() ? NULL : FleGpu::ctx();
  for (size_t i = 0; i < dev_buffers_.size(); ++i)
    if (ctx && dev_buffers_[i].first) fle_dev_free(ctx, dev_buffers_[i].first);
}

inline void* HdfsParquetScanner::DeviceBuffer(size_t slot, size_t bytes) {
  if (dev_buffers_.size() <= slot) dev_buffers_.resize(slot + 1, std::make_pair((void*)NULL, (size_t)0));
  std::pair<void*, size_t>& b = dev_buffers_[slot];
  if (b.second >= bytes && b.first) return b.first;
  fle_ctx* ctx = FleGpu::ctx();
  if (ctx == NULL) return NULL;
  if (b.first) fle_dev_free(ctx, b.first);
  b.first = NULL;
  b.second = 0;
  const size_t want = std::max<size_t>(bytes, 256);
  if (FleGpu::Check(fle_dev_alloc(ctx, want, &b.first), "fle_dev_alloc") != FLE_OK) return NULL;
  b.second = want;
  return b.first;
}

inline int HdfsParquetScanner::slot_offset(int idx) const { return column_readers_[idx]->tuple_offset_; }

template<typename T>
inline int HdfsParquetScanner::AddColumn(int max_def_level, const uint8_t* dict_page, int dict_len,
                                         const std::vector<DataPage>& pages) {
  ColumnReader<T>* r = new ColumnReader<T>(this, max_def_level);
  // the dictionary page is copied out of the I/O buffer as in :829-842
  r->dict_copy_.assign(dict_page, dict_page + dict_len);
  r->CreateDictionaryDecoder(r->dict_copy_.empty() ? NULL : &r->dict_copy_[0], dict_len);
  r->pages_ = pages;
  column_readers_.push_back(r);
  // tuple layout: [null byte per column][slots]; recomputed as columns are added
  int off = static_cast<int>(column_readers_.size());
  for (size_t i = 0; i < column_readers_.size(); ++i) {
    BaseColumnReader* c = column_readers_[i];
    off = (off + c->slot_size_ - 1) / c->slot_size_ * c->slot_size_;
    c->tuple_offset_ = off;
    c->null_offset_ = static_cast<int>(i);
    off += c->slot_size_;
  }
  tuple_size_ = (off + 7) / 8 * 8;
  return static_cast<int>(column_readers_.size()) - 1;
}

inline Status HdfsParquetScanner::BaseColumnReader::ReadDataPage() {
  num_buffered_values_ = 0;
  if (next_page_ >= pages_.size()) return Status::OK();  // end of column chunk
  const DataPage& pg = pages_[next_page_++];
  data_ = const_cast<uint8_t*>(pg.data);
  data_size_ = pg.size;
  page_encoding_ = pg.encoding;
  num_buffered_values_ = pg.num_values;
  page_num_values_ = pg.num_values;
  def_data_ = NULL;
  def_bytes_ = 0;
  if (max_def_level() > 0) {  // :882-919, definition_level_encoding == FLE
    if (data_size_ < 4) return Status("page too small for definition levels");
    int32_t num_definition_bytes = 0;
    memcpy(&num_definition_bytes, data_, 4);
    data_ += 4;
    data_size_ -= 4;
    if (num_definition_bytes <= 0 || num_definition_bytes > data_size_) return Status("bad definition level length");
    const int bit_width = FleLog2Ceil(max_def_level() + 1);
    fle_def_levels_.reset(new FleDecoder(data_, num_definition_bytes, bit_width));
    def_data_ = data_;
    def_bytes_ = num_definition_bytes;
    data_ += num_definition_bytes;
    data_size_ -= num_definition_bytes;
  }
  if (!page_is_dictionary() && page_encoding_ != parquet_encoding::PLAIN)
    return Status("unsupported data page encoding");
  // Data can be empty if the column contains all NULLs (:921)
  if (data_size_ != 0) return InitDataPage(data_, data_size_);
  return Status::OK();
}

inline int HdfsParquetScanner::BaseColumnReader::ReadDefinitionLevel() {  // :924-951
  if (max_def_level() == 0) return 1;
  uint8_t definition_level;
  if (!fle_def_levels_->Get(&definition_level)) return -1;
  return definition_level;
}

inline bool HdfsParquetScanner::BaseColumnReader::ReadValue(uint8_t* tuple, int skip_rows) {  // :1006-1026
  // (the reference compares the level with max_def_level() unconditionally; for a REQUIRED column
  //  ReadDefinitionLevel() returns 1 != 0, so every value reads as NULL -- kept under reference_compat)
  const bool compat = FleGpu::reference_compat();
  int tmp_skip_rows = 0;
  for (int i = 0; i < skip_rows; ++i) {
    int definition_level = ReadDefinitionLevel();
    if (definition_level < 0) return false;
    if ((max_def_level() > 0 || compat) && definition_level != max_def_level()) continue;
    ++tmp_skip_rows;
  }
  int definition_level = ReadDefinitionLevel();
  if (definition_level < 0) return false;
  if ((max_def_level() > 0 || compat) && definition_level != max_def_level()) {
    tuple[null_offset_] = 1;  // tuple->SetNull(...)
    if (!SkipSlot(tmp_skip_rows)) return false;
    return true;
  }
  return ReadSlot(tuple + tuple_offset_, tmp_skip_rows);
}

inline bool HdfsParquetScanner::BaseColumnReader::SkipValue(int skip_rows) {  // :1028-1037
  const bool compat = FleGpu::reference_compat();
  int tmp_skip_rows = 0;
  for (int i = 0; i < skip_rows; ++i) {
    int definition_level = ReadDefinitionLevel();
    if (definition_level < 0) return false;
    if ((max_def_level() > 0 || compat) && definition_level != max_def_level()) continue;
    ++tmp_skip_rows;
  }
  return SkipSlot(tmp_skip_rows);
}

#define FLE_SCANNER_TRAMPOLINE(Name, ArgT)                                                        \
  template<typename T>                                                                            \
  inline void HdfsParquetScanner::Name(int idx, int64_t num_rows,                                 \
      boost::dynamic_bitset<>& skip_bitset, ArgT val) {                                           \
    reinterpret_cast<HdfsParquetScanner::ColumnReader<T>*>(column_readers_[idx])->Name(           \
        num_rows, skip_bitset, val);                                                              \
  }
FLE_SCANNER_TRAMPOLINE(Eq, T&)
FLE_SCANNER_TRAMPOLINE(Lt, T&)
FLE_SCANNER_TRAMPOLINE(Le, T&)
FLE_SCANNER_TRAMPOLINE(Gt, T&)
FLE_SCANNER_TRAMPOLINE(Ge, T&)
FLE_SCANNER_TRAMPOLINE(In, std::vector<T>&)
#undef FLE_SCANNER_TRAMPOLINE

inline bool HdfsParquetScanner::EvalSimplePredicates(boost::dynamic_bitset<>& skip_bitset) {
  int64_t limit_rows = 1024;
  for (size_t c = 0; c < column_readers_.size(); ++c) {
    if (column_readers_[c]->num_buffered_values() == 0) {
      parse_status_ = column_readers_[c]->ReadDataPage();
      fused_rows_ = 0;  // a reader moved to a new page: the fused span ends here
      if (column_readers_[c]->num_buffered_values() == 0 || !parse_status_.ok()) return false;
    }
    if (column_readers_[c]->num_buffered_values() < limit_rows) {
      limit_rows = column_readers_[c]->num_buffered_values();
    }
  }
  std::vector<int64_t> page_pos(column_readers_.size());
  for (size_t c = 0; c < column_readers_.size(); ++c) {
    page_pos[c] = column_readers_[c]->row_position();
    column_readers_[c]->num_buffered_values(limit_rows);
  }
  if (simple_predicates_.empty()) return false;

  if (fused_enabled_ && FusedWindow(limit_rows, page_pos, skip_bitset)) return FleGpu::ok();

  ++leaf_windows_;
  simple_predicates_[0]->GetBitset(this, limit_rows, skip_bitset);
  for (size_t i = 1; i < simple_predicates_.size(); ++i) {
    boost::dynamic_bitset<> tmp_bitset;
    simple_predicates_[i]->GetBitset(this, limit_rows, tmp_bitset);
    skip_bitset &= tmp_bitset;
  }
  if (!FleGpu::ok()) {
    parse_status_ = Status(FleGpu::message());
    return false;
  }
  return true;
}

inline bool HdfsParquetScanner::LowerPredicates(FusedProgram* prog) {
  for (size_t i = 0; i < simple_predicates_.size(); ++i) {
    if (!simple_predicates_[i]->Lower(this, prog)) return false;
    if (i > 0) prog->Push(FLE_OP_AND);  // the conjunct `&=` loop, :1857-1862
  }
  return !prog->nodes.empty() && prog->nodes.size() <= FLE_MAX_NODES;
}

// AssembleRows for a span of whole pages, on the device.  The reference walks the window bitmap
// bit by bit and calls ReadValue(skip) per survivor and column (:1134-1168); the host mirror of
// that loop costs ~20 ns per ROW (4 Mi rows: 80 ms) although the predicate kernel needs
// microseconds.  Here the pages' FLE sections are uploaded once, the fused kernel writes the
// bitmap, fle_bitmap_select_dev turns it into ascending row ids, fle_gather_dev reads the
// survivors' codes of every column, and only those codes come back for the dictionary lookup.
// Qualifies: every reader at the start of a dictionary-encoded page, equal page row counts, no
// NULL in the span (def levels all present -- Impala writes them for every column), widths 1..32.
inline bool HdfsParquetScanner::DevicePageSpan(std::vector<uint8_t>* rows, int64_t* num_passed, int64_t* span_rows) {
  const int ncols = static_cast<int>(column_readers_.size());
  if (ncols == 0 || ncols > FLE_MAX_COLS || simple_predicates_.empty()) return false;
  for (int c = 0; c < ncols; ++c) {  // page advance exactly as EvalSimplePredicates does it
    BaseColumnReader* r = column_readers_[c];
    if (r->num_buffered_values() == 0) {
      parse_status_ = r->ReadDataPage();
      fused_rows_ = 0;
      if (r->num_buffered_values() == 0 || !parse_status_.ok()) return false;
    }
  }
  const int64_t span = column_readers_[0]->page_num_values();
  if (span <= 0 || span >= (int64_t(1) << 31)) return false;
  const int64_t groups = (span + 63) / 64;
  for (int c = 0; c < ncols; ++c) {
    BaseColumnReader* r = column_readers_[c];
    if (!r->page_is_dictionary() || r->page_num_values() != span || r->num_buffered_values() != span ||
        r->row_position() != 0) return false;
    const int bw = r->fle_bit_width();
    if (bw < 1 || bw > 32 || r->fle_data_len() < groups * bw * 8) return false;  // (a shorter stream means NULLs)
    if (r->max_def_level() > 1) return false;
    if (r->max_def_level() == 1) {  // every def level present?  1-bit FLE: row k of a group at bit 63 - k
      if (r->def_bytes() < groups * 8) return false;
      const uint8_t* d = r->def_data();
      for (int64_t g = 0; g < groups; ++g) {
        uint64_t w;
        memcpy(&w, d + g * 8, 8);
        const int64_t left = span - g * 64;
        const uint64_t want = left >= 64 ? ~uint64_t(0) : (~uint64_t(0) << (64 - left));
        if ((w & want) != want) return false;
      }
    }
  }
  FusedProgram prog;
  if (!LowerPredicates(&prog)) return false;
  fle_ctx* ctx = FleGpu::ctx();
  if (ctx == NULL) return false;
  std::vector<fle_col_desc> cols(ncols);
  for (int c = 0; c < ncols; ++c) {
    BaseColumnReader* r = column_readers_[c];
    const size_t bytes = static_cast<size_t>(groups) * r->fle_bit_width() * 8;
    void* dev = DeviceBuffer(c, bytes);
    if (dev == NULL) return false;
    if (FleGpu::Check(fle_h2d(ctx, dev, r->fle_data(), bytes), "fle_h2d(page)") != FLE_OK) return false;
    memset(&cols[c], 0, sizeof(fle_col_desc));
    cols[c].planes = dev;
    cols[c].bytes = bytes;
    cols[c].bit_width = r->fle_bit_width();
  }
  uint64_t* bitmap_dev = static_cast<uint64_t*>(DeviceBuffer(FLE_MAX_COLS, groups * 8));
  uint64_t* count_dev = static_cast<uint64_t*>(DeviceBuffer(FLE_MAX_COLS + 1, 8));
  uint32_t* ids_dev = static_cast<uint32_t*>(DeviceBuffer(FLE_MAX_COLS + 2, static_cast<size_t>(span) * 4));
  if (!bitmap_dev || !count_dev || !ids_dev) return false;
  int rc = fle_eval_dev(ctx, &cols[0], ncols, &prog.nodes[0], static_cast<int>(prog.nodes.size()), 0, span, bitmap_dev, NULL);
  if (rc == FLE_E_UNSUPPORTED) return false;
  if (FleGpu::Check(rc, "fle_eval_dev(page span)") != FLE_OK) return false;
  if (FleGpu::Check(fle_bitmap_select_dev(ctx, bitmap_dev, span, ids_dev, span, count_dev), "fle_bitmap_select_dev") != FLE_OK) return false;
  uint64_t passed = 0;
  if (FleGpu::Check(fle_d2h(ctx, &passed, count_dev, 8), "fle_d2h(count)") != FLE_OK) return false;
  if (FleGpu::Check(fle_ctx_sync(ctx), "fle_ctx_sync") != FLE_OK) return false;
  const size_t base = rows->size();
  rows->resize(base + static_cast<size_t>(passed) * tuple_size_, 0);
  if (passed > 0) {
    uint32_t* codes_dev = static_cast<uint32_t*>(DeviceBuffer(FLE_MAX_COLS + 3, static_cast<size_t>(passed) * 4 * ncols));
    if (!codes_dev) return false;
    for (int c = 0; c < ncols; ++c) {
      rc = fle_gather_dev(ctx, static_cast<const uint8_t*>(cols[c].planes), cols[c].bytes, cols[c].bit_width, ids_dev, count_dev,
                          static_cast<int64_t>(passed), NULL, 0, 0, codes_dev + static_cast<size_t>(c) * passed);
      if (FleGpu::Check(rc, "fle_gather_dev") != FLE_OK) return false;
    }
    host_codes_.resize(static_cast<size_t>(passed) * ncols);
    if (FleGpu::Check(fle_d2h(ctx, &host_codes_[0], codes_dev, host_codes_.size() * 4), "fle_d2h(codes)") != FLE_OK) return false;
    if (FleGpu::Check(fle_ctx_sync(ctx), "fle_ctx_sync") != FLE_OK) return false;
    for (int c = 0; c < ncols; ++c) {  // :1151-1168, column by column
      BaseColumnReader* r = column_readers_[c];
      uint8_t* tuple = &(*rows)[base];
      const uint32_t* codes = &host_codes_[static_cast<size_t>(c) * passed];
      for (uint64_t j = 0; j < passed; ++j, tuple += tuple_size_) {
        if (!r->SlotFromCode(tuple + r->tuple_offset_, codes[j])) {
          parse_status_ = Status("dictionary index out of range");
          return false;
        }
      }
    }
  }
  for (int c = 0; c < ncols; ++c) column_readers_[c]->num_buffered_values(span);  // the pages are consumed
  fused_rows_ = 0;
  ++fused_evals_;
  ++device_spans_;
  *num_passed = static_cast<int64_t>(passed);
  *span_rows = span;
  return true;
}

// ---------------------------------------------------------------------------------------
// Row-group batch (see the file comment).  Page body framing as in ReadDataPage above.
// ---------------------------------------------------------------------------------------
inline bool HdfsParquetScanner::StartRowGroupUpload() {
  if (batch_state_ != 0) return batch_state_ > 0;
  batch_state_ = -1;
  const int ncols = static_cast<int>(column_readers_.size());
  if (!row_group_batch_ || !fused_enabled_ || !device_materialise_ || FleGpu::reference_compat()) return false;
  if (ncols == 0 || ncols > FLE_MAX_COLS || simple_predicates_.empty()) return false;
  if (column_readers_[0]->pages().empty()) return false;
  // every column: fresh reader, dictionary pages only, the same number of rows in the row group.  Pages may hold any
  // number of rows (the reference's writer closes a dictionary page after 524225 rows,
  // hdfs-parquet-table-writer.cc:387) and the columns may cut their pages differently: every page is placed at its
  // row offset of the column's ONE row-aligned plane array (fle_page_ref.dst_bit).
  struct PageView { const uint8_t* fle; int64_t fle_bytes; int bw; const uint8_t* def; int64_t def_bytes; int64_t rows; bool plain; };
  std::vector<std::vector<PageView> > views(ncols);
  std::vector<char> col_is_plain(ncols, 0);
  int64_t total_rows = 0;
  for (int c = 0; c < ncols; ++c) {
    BaseColumnReader* r = column_readers_[c];
    const size_t npages = r->pages().size();
    if (!r->fresh() || npages == 0 || r->max_def_level() > 1) return false;
    views[c].resize(npages);
    int64_t col_rows = 0;
    // a column is dictionary-encoded throughout, or it overflowed its dictionary and continues PLAIN (the writer
    // starts every row group with a dictionary and gives up after 40000 distinct values,
    // hdfs-parquet-table-writer.cc:276-279): such a column can be materialised -- dictionary pages decoded, PLAIN pages
    // copied into one row-aligned value array -- but carries no predicate here
    bool any_plain = false;
    for (size_t p = 0; p < npages; ++p) any_plain = any_plain || r->pages()[p].encoding == parquet_encoding::PLAIN;
    col_is_plain[c] = any_plain;
    for (size_t p = 0; p < npages; ++p) {
      const DataPage& pg = r->pages()[p];
      const bool col_plain = pg.encoding == parquet_encoding::PLAIN;  // (this page)
      if (!col_plain && pg.encoding != parquet_encoding::FLE_DICTIONARY && pg.encoding != parquet_encoding::PLAIN_DICTIONARY) return false;
      if (pg.num_values <= 0) return false;
      col_rows += pg.num_values;
      PageView& v = views[c][p];
      const uint8_t* d = pg.data;
      int64_t left = pg.size;
      v.def = NULL;
      v.def_bytes = 0;
      v.rows = pg.num_values;
      if (r->max_def_level() > 0) {
        if (left < 4) return false;
        int32_t nd = 0;
        memcpy(&nd, d, 4);
        if (nd <= 0 || nd > left - 4 || nd < (v.rows + 63) / 64 * 8) return false;
        v.def = d + 4;
        v.def_bytes = nd;
        d += 4 + nd;
        left -= 4 + nd;
      }
      v.plain = col_plain;
      if (col_plain) {  // values back to back, r->slot_size_ bytes each (only the non-null ones)
        v.bw = 1;
        v.fle = d;
        v.fle_bytes = left;
        if (v.def == NULL && left < v.rows * r->slot_size_) return false;
      } else if (left < 1) {  // all-NULL page: no width byte, no data (:921)
        if (v.def == NULL) return false;
        v.bw = 1;
        v.fle = d;
        v.fle_bytes = 0;
      } else {
        v.bw = d[0];
        v.fle = d + 1;
        v.fle_bytes = left - 1;
      }
      if (v.bw < 1 || v.bw > 32) return false;
      // (a nullable page's FLE section holds only the non-null values: it may be shorter than the rows)
      if (!col_plain && v.def == NULL && v.fle_bytes < (v.rows + 63) / 64 * v.bw * 8) return false;
    }
    if (c == 0) total_rows = col_rows;
    else if (col_rows != total_rows) return false;
  }
  if (total_rows >= (int64_t(1) << 31)) return false;
  batch_prog_ = FusedProgram();
  if (!LowerPredicates(&batch_prog_)) return false;
  batch_cols_.assign(ncols, BatchColumn());
  batch_eval_cols_ = 0;
  for (int c = 0; c < ncols; ++c) {
    batch_cols_[c].plain = col_is_plain[c] != 0;
    batch_cols_[c].eval_col = col_is_plain[c] ? -1 : batch_eval_cols_++;
  }
  std::vector<int> node_col(batch_prog_.nodes.size(), -1);  // the leaves' columns as the scanner numbers them
  for (size_t i = 0; i < batch_prog_.nodes.size(); ++i) {
    fle_pred_node& nd = batch_prog_.nodes[i];
    if (nd.op == FLE_OP_AND || nd.op == FLE_OP_OR) continue;
    if (nd.col < 0 || nd.col >= ncols || col_is_plain[nd.col]) return false;  // (a predicate on a PLAIN column: the page paths)
    node_col[i] = nd.col;
    nd.col = batch_cols_[nd.col].eval_col;
  }
  if (batch_eval_cols_ == 0) return false;
  if (FleGpu::ctx() == NULL) return false;
  batch_bank_ = FleGpu::NextBank();
  fle_ctx* ctx = FleGpu::bank_ctx(batch_bank_);
  if (ctx == NULL) return false;
  batch_refs_.clear();
  const int64_t total_groups = (total_rows + 63) / 64;
  for (int c = 0; c < ncols; ++c) {
    BaseColumnReader* r = column_readers_[c];
    BatchColumn& bc = batch_cols_[c];
    const size_t npages = r->pages().size();
    bc.bit_width = 0;
    for (size_t p = 0; p < npages; ++p) bc.bit_width = std::max(bc.bit_width, views[c][p].bw);
    // one upload per column chunk when its pages lie back to back in host memory (as they do in the I/O buffer
    // a column chunk was read into), one per page otherwise
    const uint8_t* lo = r->pages()[0].data;
    const uint8_t* hi = lo;
    int64_t sum = 0;
    bool ascending = true;
    for (size_t p = 0; p < npages; ++p) {
      const DataPage& pg = r->pages()[p];
      if (pg.data < hi) ascending = false;
      lo = std::min(lo, pg.data);
      hi = std::max(hi, pg.data + pg.size);
      sum += pg.size;
    }
    const bool one_copy = ascending && (hi - lo) <= sum + sum / 4 + 4096;
    const size_t image_bytes = one_copy ? static_cast<size_t>(hi - lo) : static_cast<size_t>(sum + 16 * npages);
    bc.image_dev = FleGpu::DeviceScratch(batch_bank_, 3 * c, image_bytes + 64);
    bc.planes_dev = FleGpu::DeviceScratch(batch_bank_, 3 * c + 1,
                                          bc.plain ? static_cast<size_t>(total_groups) * 64 * r->slot_size_ + 64
                                                   : static_cast<size_t>(total_groups) * bc.bit_width * 8 + 64);
    bc.def_dev = NULL;
    if (r->max_def_level() > 0) {
      bc.def_dev = FleGpu::DeviceScratch(batch_bank_, 24 + c, static_cast<size_t>(total_groups) * 8 + 64);
      if (!bc.def_dev) return false;
    }
    int nd = 0, es = 0;
    const void* dv = r->dict_values(&nd, &es);
    bc.ndict = nd;
    bc.dict_dev = FleGpu::DeviceScratch(batch_bank_, 3 * c + 2, static_cast<size_t>(std::max(nd, 1)) * es);
    if (!bc.image_dev || !bc.planes_dev || !bc.dict_dev) return false;
    if (nd > 0) {
      // through pinned staging: an asynchronous copy from pageable memory makes the host wait for the copy
      // stream, which is busy with the previous column chunk
      void* stage = FleGpu::PinnedScratch(batch_bank_, 2 + c, static_cast<size_t>(nd) * es);
      if (!stage) return false;
      memcpy(stage, dv, static_cast<size_t>(nd) * es);
      if (FleGpu::Check(fle_h2d_prefetch(ctx, bc.dict_dev, stage, static_cast<size_t>(nd) * es), "fle_h2d_prefetch(dict)") != FLE_OK) return false;
    }
    if (one_copy && FleGpu::Check(fle_h2d_prefetch(ctx, bc.image_dev, lo, image_bytes), "fle_h2d_prefetch(chunk)") != FLE_OK) return false;
    int64_t row0 = 0;  // first row of the page inside the row group
    size_t off = 0;
    for (size_t p = 0; p < npages; ++p) {
      const DataPage& pg = r->pages()[p];
      const PageView& v = views[c][p];
      size_t base;  // offset of the page body inside the device image
      if (one_copy) {
        base = static_cast<size_t>(pg.data - lo);
      } else {
        base = off;
        if (FleGpu::Check(fle_h2d_prefetch(ctx, static_cast<uint8_t*>(bc.image_dev) + off, pg.data, pg.size), "fle_h2d_prefetch(page)") != FLE_OK) return false;
        off += (static_cast<size_t>(pg.size) + 15) / 16 * 16;
      }
      fle_page_ref ref;
      memset(&ref, 0, sizeof(ref));
      const uint8_t* img = static_cast<const uint8_t*>(bc.image_dev) + base;
      ref.src_dev = img + (v.fle - pg.data);
      ref.src_bytes = static_cast<uint64_t>(v.fle_bytes);
      ref.src_bit_width = v.bw;
      ref.dst_bit_width = bc.bit_width;
      const int64_t g0 = row0 >> 6;
      ref.dst_dev = static_cast<uint8_t*>(bc.planes_dev) + static_cast<size_t>(g0) * bc.bit_width * 8;
      ref.dst_bit = static_cast<int32_t>(row0 & 63);
      ref.groups = (v.rows + 63) / 64;
      if (bc.plain) {  // value slots instead of plane groups
        ref.plain_size = r->slot_size_;
        ref.dst_dev = static_cast<uint8_t*>(bc.planes_dev) + static_cast<size_t>(row0) * r->slot_size_;
        ref.dst_bit = 0;
        if (!v.plain) {  // a dictionary page in front of the PLAIN ones: decoded through the dictionary
          ref.dict_dev = bc.dict_dev;
          ref.ndict = bc.ndict;
        }
      }
      ref.rows = v.rows;
      if (v.def != NULL) {
        // the page's values are indexed by non-null ordinal: expanded into row-aligned planes (code 0 at NULL rows;
        // with no NULL in the page that is a plain copy), its definition levels counted on the way ...
        ref.def_src_dev = img + (v.def - pg.data);
        ref.def_bytes = static_cast<uint64_t>(v.def_bytes);
        ref.expand = 1;
        batch_refs_.push_back(ref);
        // ... and gathered as a 1-bit column of their own
        fle_page_ref dref;
        memset(&dref, 0, sizeof(dref));
        dref.src_dev = ref.def_src_dev;
        dref.src_bytes = ref.def_bytes;
        dref.src_bit_width = dref.dst_bit_width = 1;
        dref.dst_dev = static_cast<uint8_t*>(bc.def_dev) + static_cast<size_t>(g0) * 8;
        dref.dst_bit = static_cast<int32_t>(row0 & 63);
        dref.groups = ref.groups;
        dref.rows = v.rows;
        batch_refs_.push_back(dref);
      } else {
        batch_refs_.push_back(ref);
      }
      row0 += v.rows;
    }
  }
  // the program for a row group with NULLs: leaves on nullable columns ANDed with the column's definition levels
  // (once per column behind a pure AND tree, per leaf otherwise)
  batch_prog_nulls_.clear();
  batch_def_col_.assign(ncols, -1);
  batch_ncols_nulls_ = batch_eval_cols_;
  {
    bool and_only = true;
    for (size_t i = 0; i < batch_prog_.nodes.size(); ++i) and_only = and_only && batch_prog_.nodes[i].op != FLE_OP_OR;
    std::vector<int> used;
    for (size_t i = 0; i < batch_prog_.nodes.size(); ++i) {
      const fle_pred_node& nd = batch_prog_.nodes[i];
      batch_prog_nulls_.push_back(nd);
      if (nd.op == FLE_OP_AND || nd.op == FLE_OP_OR) continue;
      const int oc = node_col[i];
      if (oc < 0 || column_readers_[oc]->max_def_level() == 0) continue;
      if (batch_def_col_[oc] < 0) {
        batch_def_col_[oc] = batch_ncols_nulls_++;
        used.push_back(oc);
      }
      if (!and_only) {
        fle_pred_node eq;
        memset(&eq, 0, sizeof(eq));
        eq.op = FLE_OP_EQ; eq.col = batch_def_col_[oc]; eq.code = 1;
        batch_prog_nulls_.push_back(eq);
        fle_pred_node an;
        memset(&an, 0, sizeof(an));
        an.op = FLE_OP_AND;
        batch_prog_nulls_.push_back(an);
      }
    }
    if (and_only) {
      for (size_t k = 0; k < used.size(); ++k) {
        fle_pred_node eq;
        memset(&eq, 0, sizeof(eq));
        eq.op = FLE_OP_EQ; eq.col = batch_def_col_[used[k]]; eq.code = 1;
        batch_prog_nulls_.push_back(eq);
        fle_pred_node an;
        memset(&an, 0, sizeof(an));
        an.op = FLE_OP_AND;
        batch_prog_nulls_.push_back(an);
      }
    }
    if (batch_ncols_nulls_ > FLE_MAX_COLS || batch_prog_nulls_.size() > FLE_MAX_NODES) batch_ncols_nulls_ = 0;
  }
  if (FleGpu::Check(fle_prefetch_mark(ctx, batch_bank_), "fle_prefetch_mark") != FLE_OK) return false;
  batch_rows_ = total_rows;
  batch_state_ = 1;
  // the row group's kernels right behind its upload (nothing here waits; a failure is looked at by DeviceRowGroup)
  batch_enqueued_nulls_ = batch_nulls_seen_ && batch_ncols_nulls_ > 0;
  batch_enqueued_ = EnqueueRowGroup(batch_enqueued_nulls_, true);
  return true;
}

// gather + evaluation + tuple assembly + the counts on their way back, all enqueued on the bank's context; first: also
// the wait for the row group's upload and the page placement (a second evaluation, with the NULL program, reuses them)
inline bool HdfsParquetScanner::EnqueueRowGroup(bool with_nulls, bool first) {
  fle_ctx* ctx = FleGpu::bank_ctx(batch_bank_);
  if (ctx == NULL) return false;
  const int ncols = static_cast<int>(column_readers_.size());
  const int64_t n = batch_rows_, groups = (n + 63) / 64;
  const int bank = batch_bank_;
  if (first) {
    batch_bitmap_dev_ = static_cast<uint64_t*>(FleGpu::DeviceScratch(bank, 40, static_cast<size_t>(groups) * 8 + 64));
    batch_hdr_dev_ = static_cast<uint64_t*>(FleGpu::DeviceScratch(bank, 41, 64));
    batch_hdr_ = static_cast<uint64_t*>(FleGpu::PinnedScratch(bank, 0, 64));
    // tuples: room for every row (grow-only scratch; only the survivors' bytes come back)
    batch_tuples_dev_ = static_cast<uint8_t*>(FleGpu::DeviceScratch(bank, 42, static_cast<size_t>(n) * tuple_size_ + 64));
    if (!batch_bitmap_dev_ || !batch_hdr_dev_ || !batch_hdr_ || !batch_tuples_dev_) return false;
    // (this row group's uploads only: the next row group's keep crossing PCIe meanwhile)
    if (FleGpu::Check(fle_prefetch_wait(ctx, bank), "fle_prefetch_wait") != FLE_OK) return false;
    if (FleGpu::Check(fle_pages_gather_dev(ctx, &batch_refs_[0], static_cast<int>(batch_refs_.size()), batch_hdr_dev_), "fle_pages_gather_dev") != FLE_OK) return false;
    std::vector<fle_col_desc>& cols = batch_eval_desc_;
    std::vector<fle_slot_desc>& slots = batch_slots_;
    cols.assign(std::max(batch_eval_cols_, batch_ncols_nulls_), fle_col_desc());
    slots.assign(ncols, fle_slot_desc());
    for (int c = 0; c < ncols; ++c) {
      BaseColumnReader* r = column_readers_[c];
      memset(&slots[c], 0, sizeof(fle_slot_desc));
      if (batch_cols_[c].plain) {
        slots[c].plain_values_dev = batch_cols_[c].planes_dev;
        slots[c].bytes = static_cast<uint64_t>(groups) * 64 * r->slot_size_;
      } else {
        fle_col_desc& e = cols[batch_cols_[c].eval_col];
        memset(&e, 0, sizeof(fle_col_desc));
        e.planes = batch_cols_[c].planes_dev;
        e.bytes = static_cast<uint64_t>(groups) * batch_cols_[c].bit_width * 8;
        e.bit_width = batch_cols_[c].bit_width;
        slots[c].planes_dev = batch_cols_[c].planes_dev;
        slots[c].bytes = e.bytes;
      }
      slots[c].bit_width = batch_cols_[c].bit_width;
      slots[c].slot_size = r->slot_size_;
      slots[c].dict_dev = batch_cols_[c].dict_dev;
      slots[c].ndict = batch_cols_[c].ndict;
      slots[c].tuple_offset = r->tuple_offset_;
      slots[c].def_planes_dev = batch_cols_[c].def_dev;  // a survivor that is NULL here gets its indicator set (ReadSlot)
      slots[c].null_offset = r->null_offset_;
      if (batch_ncols_nulls_ > 0 && batch_def_col_[c] >= 0) {
        fle_col_desc& d = cols[batch_def_col_[c]];
        memset(&d, 0, sizeof(fle_col_desc));
        d.planes = batch_cols_[c].def_dev;
        d.bytes = static_cast<uint64_t>(groups) * 8;
        d.bit_width = 1;
      }
    }
    batch_placed_ = true;
  }
  int rc = with_nulls ? fle_eval_dev(ctx, &batch_eval_desc_[0], batch_ncols_nulls_, &batch_prog_nulls_[0], static_cast<int>(batch_prog_nulls_.size()), 0, n, batch_bitmap_dev_, NULL)
                      : fle_eval_dev(ctx, &batch_eval_desc_[0], batch_eval_cols_, &batch_prog_.nodes[0], static_cast<int>(batch_prog_.nodes.size()), 0, n, batch_bitmap_dev_, NULL);
  if (rc == FLE_E_UNSUPPORTED) return false;
  if (FleGpu::Check(rc, "fle_eval_dev(row group)") != FLE_OK) return false;
  rc = fle_assemble_rows_dev(ctx, batch_bitmap_dev_, n, &batch_slots_[0], ncols, tuple_size_, batch_tuples_dev_, n, batch_hdr_dev_ + 1, batch_hdr_dev_ + 2);
  if (FleGpu::Check(rc, "fle_assemble_rows_dev") != FLE_OK) return false;
  if (FleGpu::Check(fle_d2h(ctx, batch_hdr_, batch_hdr_dev_, 24), "fle_d2h(counts)") != FLE_OK) return false;
  return true;
}

inline double TraceNow() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

inline bool HdfsParquetScanner::DeviceRowGroup(std::vector<uint8_t>* rows, int64_t* num_passed, int64_t* span_rows) {
  if (!StartRowGroupUpload()) return false;
  batch_state_ = -1;  // one shot: whatever happens below, the batch is not retried
  const int bank = batch_bank_;
  fle_ctx* ctx = FleGpu::bank_ctx(bank);
  if (ctx == NULL) return false;
  const int ncols = static_cast<int>(column_readers_.size());
  const int64_t n = batch_rows_;
  // A row group without NULLs (every Impala-written TPC-H file) is the plain program over the row-aligned planes.
  // NULLs are only known once the counts are back: the evaluation is then repeated with the NULL program, and the
  // scanner starts with that one from its next row group on.
  bool with_nulls = batch_enqueued_nulls_;
  if (!batch_enqueued_) {
    with_nulls = batch_nulls_seen_ && batch_ncols_nulls_ > 0;
    if (!EnqueueRowGroup(with_nulls, !batch_placed_)) return false;
  }
  uint64_t* hdr = batch_hdr_;
  uint8_t* tuples_dev = batch_tuples_dev_;
  uint64_t passed = 0;
  static const bool trace = std::getenv("FLE_SCANNER_TRACE") != NULL;
  const double t_sync0 = trace ? TraceNow() : 0;
  for (;;) {
    if (FleGpu::Check(fle_ctx_sync(ctx), "fle_ctx_sync") != FLE_OK) return false;
    if (hdr[0] != 0 && !with_nulls) {
      batch_nulls_seen_ = true;
      if (batch_ncols_nulls_ == 0) return false;  // the NULL program does not fit: the page-span / window paths (nothing was consumed)
      with_nulls = true;
      if (!EnqueueRowGroup(true, false)) return false;
      continue;
    }
    break;
  }
  const double t_sync1 = trace ? TraceNow() : 0;
  if (hdr[2] != 0) {
    parse_status_ = Status("dictionary index out of range");
    return false;
  }
  passed = hdr[1];
  if (passed > 0) {
    const size_t bytes = static_cast<size_t>(passed) * tuple_size_;
    if (direct_out_ != NULL && direct_used_ + bytes <= direct_cap_) {
      // AssembleRowsInto: the tuples land in the caller's (pinned) tuple buffer, no staging copy
      if (FleGpu::Check(fle_d2h(ctx, direct_out_ + direct_used_, tuples_dev, bytes), "fle_d2h(tuples)") != FLE_OK) return false;
      if (FleGpu::Check(fle_ctx_sync(ctx), "fle_ctx_sync") != FLE_OK) return false;
      direct_used_ += bytes;
    } else {
      uint8_t* stage = static_cast<uint8_t*>(FleGpu::PinnedScratch(bank, 1, bytes));
      if (!stage) return false;
      if (FleGpu::Check(fle_d2h(ctx, stage, tuples_dev, bytes), "fle_d2h(tuples)") != FLE_OK) return false;
      if (FleGpu::Check(fle_ctx_sync(ctx), "fle_ctx_sync") != FLE_OK) return false;
      rows->insert(rows->end(), stage, stage + bytes);
    }
  }
  if (trace) {
    const double t_end = TraceNow();
    std::fprintf(stderr, "scanner trace: wait for the row group %.1f us, tuples back %.1f us (%llu survivors)\n", (t_sync1 - t_sync0) * 1e6,
                 (t_end - t_sync1) * 1e6, (unsigned long long)passed);
  }
  if (hdr[0] != 0) ++row_group_null_batches_;
  for (int c = 0; c < ncols; ++c) column_readers_[c]->ConsumeAllPages();
  fused_rows_ = 0;
  ++fused_evals_;
  ++row_group_batches_;
  *num_passed = static_cast<int64_t>(passed);
  *span_rows = n;
  return true;
}

inline bool HdfsParquetScanner::BuildFusedSpan(const std::vector<int64_t>& page_pos) {
  const int ncols = static_cast<int>(column_readers_.size());
  if (ncols > FLE_MAX_COLS) return false;
  int64_t span = -1;
  std::vector<fle_col_desc> cols(ncols);
  for (int c = 0; c < ncols; ++c) {
    BaseColumnReader* r = column_readers_[c];
    if (!r->page_is_dictionary()) return false;
    const int64_t left = r->page_num_values() - page_pos[c];
    span = span < 0 ? left : std::min(span, left);
    memset(&cols[c], 0, sizeof(fle_col_desc));
    const int bw = r->fle_bit_width();
    if (r->max_def_level() == 0) {
      if (page_pos[c] & 63) return false;  // FLE groups hold 64 rows
      if (bw < 1 || bw > 32) return false;
      const int64_t off = (page_pos[c] / 64) * bw * 8;
      cols[c].planes = r->fle_data() + off;
      cols[c].bytes = r->fle_data_len() - off;
      cols[c].bit_width = bw;
    } else {
      if (page_pos[c] != 0) return false;  // data values are indexed by non-null ordinal
      if (r->max_def_level() != 1) return false;
      cols[c].planes = r->fle_data_len() > 0 ? r->fle_data() : NULL;
      cols[c].bytes = std::max(0, r->fle_data_len());
      cols[c].bit_width = (bw >= 1 && bw <= 32) ? bw : 1;
      cols[c].max_def_level = 1;
      cols[c].def_planes = r->def_data();
      cols[c].def_bytes = r->def_bytes();
    }
  }
  if (span <= 0) return false;
  FusedProgram prog;
  if (!LowerPredicates(&prog)) return false;
  fle_ctx* ctx = FleGpu::ctx();
  if (ctx == NULL) return false;
  fused_bitmap_.assign((span + 63) / 64, 0);
  int rc = fle_eval(ctx, &cols[0], ncols, &prog.nodes[0], static_cast<int>(prog.nodes.size()), 0, span,
                    &fused_bitmap_[0], NULL);
  if (rc == FLE_E_UNSUPPORTED) return false;  // e.g. too many leaves: per-leaf path instead
  if (FleGpu::Check(rc, "fle_eval(fused)") != FLE_OK) return false;
  ++fused_evals_;
  fused_start_ = page_pos;
  fused_rows_ = span;
  return true;
}

inline bool HdfsParquetScanner::FusedWindow(int64_t limit_rows, const std::vector<int64_t>& page_pos,
                                            boost::dynamic_bitset<>& skip_bitset) {
  // the cached span is usable if every reader sits at the same offset inside it
  bool usable = fused_rows_ > 0 && fused_start_.size() == page_pos.size();
  int64_t offset = usable ? page_pos[0] - fused_start_[0] : 0;
  for (size_t c = 0; usable && c < page_pos.size(); ++c) {
    if (page_pos[c] - fused_start_[c] != offset) usable = false;
  }
  if (!usable || offset < 0 || offset + limit_rows > fused_rows_) {
    if (!BuildFusedSpan(page_pos)) return false;
    offset = 0;
    if (limit_rows > fused_rows_) return false;
  }
  fused_pos_ = offset;
  // bit slice [fused_pos_, fused_pos_ + limit_rows) of the cached result
  int64_t pos = fused_pos_;
  int64_t left = limit_rows;
  while (left > 0) {
    const int sh = static_cast<int>(pos & 63);
    uint64_t w = fused_bitmap_[pos >> 6] >> sh;
    if (sh != 0 && static_cast<size_t>((pos >> 6) + 1) < fused_bitmap_.size()) {
      w |= fused_bitmap_[(pos >> 6) + 1] << (64 - sh);
    }
    if (left >= 64) {
      skip_bitset.append(w);
      pos += 64;
      left -= 64;
    } else {
      for (int i = 0; i < left; ++i) skip_bitset.push_back((w >> i) & 1);
      pos += left;
      left = 0;
    }
  }
  fused_pos_ = pos;
  return true;
}

inline Status HdfsParquetScanner::AssembleRowsInto(int64_t expected_rows_in_group, uint8_t* tuple_buf, size_t capacity,
                                                   int64_t* num_passed, size_t* bytes) {
  direct_out_ = tuple_buf;
  direct_cap_ = capacity;
  direct_used_ = 0;
  std::vector<uint8_t> rest;  // what the page / window paths produce (they append to a vector)
  Status st = AssembleRows(expected_rows_in_group, &rest, num_passed);
  direct_out_ = NULL;
  if (!st.ok()) return st;
  // (the row-group batch comes first and covers the whole row group when it applies, so `rest` follows the direct bytes)
  if (direct_used_ + rest.size() > capacity) return Status("AssembleRowsInto: tuple buffer too small");
  if (!rest.empty()) memcpy(tuple_buf + direct_used_, &rest[0], rest.size());
  *bytes = direct_used_ + rest.size();
  return st;
}

inline Status HdfsParquetScanner::AssembleRows(int64_t expected_rows_in_group, std::vector<uint8_t>* rows,
                                               int64_t* num_passed) {
  int64_t rows_read = 0;
  const int num_column_readers = static_cast<int>(column_readers_.size());
  boost::dynamic_bitset<> skip_bitset;
  std::vector<int> skip_rows;
  *num_passed = 0;
  const bool compat = FleGpu::reference_compat();
  if (rows_read == 0) {  // the whole row group as one batch
    int64_t passed = 0, span = 0;
    if (DeviceRowGroup(rows, &passed, &span)) {
      *num_passed += passed;
      rows_read += span;
    } else if (!parse_status_.ok() || !FleGpu::ok()) {
      return FleGpu::ok() ? parse_status_ : Status(FleGpu::message());
    }
  }
  while (rows_read < expected_rows_in_group) {
    if (fused_enabled_ && device_materialise_ && !compat) {
      int64_t passed = 0, span = 0;
      if (DevicePageSpan(rows, &passed, &span)) {
        *num_passed += passed;
        rows_read += span;
        continue;
      }
      if (!parse_status_.ok() || !FleGpu::ok()) return FleGpu::ok() ? parse_status_ : Status(FleGpu::message());
    }
    skip_bitset.clear();
    if (!EvalSimplePredicates(skip_bitset)) return parse_status_;
    if (skip_bitset.size() == 0) break;
    const int bitset_size = static_cast<int>(skip_bitset.size());
    if (skip_bitset.count() == 0) {  // :1125-1131 (the reference does not look at the result here)
      for (int c = 0; c < num_column_readers; ++c) {
        if (!column_readers_[c]->SkipValue(bitset_size) && !compat) return Status("SkipValue failed");
      }
      rows_read += bitset_size;
      continue;
    }
    skip_rows.clear();  // :1134-1148
    int start = 0;
    for (int i = 0; i < bitset_size; ++i) {
      if (skip_bitset[i]) {
        skip_rows.push_back(i - start);
        start = i + 1;
      }
    }
    const int last_skip_rows = bitset_size - start;
    const size_t base = rows->size();
    rows->resize(base + skip_rows.size() * tuple_size_, 0);
    for (int c = 0; c < num_column_readers; ++c) {  // :1151-1168, column by column
      uint8_t* tuple = &(*rows)[base];
      for (size_t i = 0; i < skip_rows.size(); ++i) {
        if (!column_readers_[c]->ReadValue(tuple, skip_rows[i])) return Status("ReadValue failed");
        tuple += tuple_size_;
      }
    }
    if (last_skip_rows != 0) {  // :1107-1113
      bool result = true;
      for (int c = 0; c < num_column_readers; ++c) result = column_readers_[c]->SkipValue(last_skip_rows);
      if (!result) {
        // the reference: `if (!result) return parse_status_;` -- a Skip() that lands on the end of a page
        // reports false there and the rest of the row group is dropped without an error
        *num_passed += static_cast<int64_t>(skip_rows.size());
        return compat ? parse_status_ : Status("SkipValue failed");
      }
    }
    *num_passed += static_cast<int64_t>(skip_rows.size());
    rows_read += bitset_size;
  }
  return FleGpu::ok() ? Status::OK() : Status(FleGpu::message());
}

// ---------------------------------------------------------------------------------------
// predicate leaves (simple-predicates.h:165-205) + their lowering to dictionary codes
// ---------------------------------------------------------------------------------------
namespace detail {
inline bool PushTranslated(FusedProgram* prog, int idx, const DictCode& c) {
  switch (c.op) {
    case FLE_OP_IN: prog->PushIn(idx, c.codes); return true;
    case FLE_OP_ALL: case FLE_OP_NONE: prog->Push(c.op, idx); return true;  // idx: NULL rows still fail
    default: prog->Push(c.op, idx, c.code); return true;
  }
}
}  // namespace detail

#define FLE_DEFINE_LEAF(Name)                                                                       \
  template <typename T>                                                                             \
  inline void Name##Operate<T>::GetBitset(HdfsParquetScanner* scanner, int64_t num_rows,            \
      boost::dynamic_bitset<>& skip_bitset) {                                                       \
    scanner->Name(idx_, num_rows, skip_bitset, val_);                                               \
  }                                                                                                 \
  template <typename T>                                                                             \
  inline bool Name##Operate<T>::Lower(HdfsParquetScanner* scanner, FusedProgram* prog) {            \
    HdfsParquetScanner::ColumnReader<T>* r =                                                        \
        reinterpret_cast<HdfsParquetScanner::ColumnReader<T>*>(scanner->column_reader(idx_));       \
    if (r->dict_decoder() == NULL) return false;                                                    \
    return detail::PushTranslated(prog, idx_, r->dict_decoder()->Translate##Name(val_));            \
  }
FLE_DEFINE_LEAF(Eq)
FLE_DEFINE_LEAF(Lt)
FLE_DEFINE_LEAF(Le)
FLE_DEFINE_LEAF(Gt)
FLE_DEFINE_LEAF(Ge)
#undef FLE_DEFINE_LEAF

template <typename T>
inline void InOperate<T>::GetBitset(HdfsParquetScanner* scanner, int64_t num_rows,
    boost::dynamic_bitset<>& skip_bitset) {
  scanner->In(idx_, num_rows, skip_bitset, vals_);
}
template <typename T>
inline bool InOperate<T>::Lower(HdfsParquetScanner* scanner, FusedProgram* prog) {
  HdfsParquetScanner::ColumnReader<T>* r =
      reinterpret_cast<HdfsParquetScanner::ColumnReader<T>*>(scanner->column_reader(idx_));
  if (r->dict_decoder() == NULL) return false;
  return detail::PushTranslated(prog, idx_, r->dict_decoder()->TranslateIn(vals_));
}

}  // namespace impala

#endif
