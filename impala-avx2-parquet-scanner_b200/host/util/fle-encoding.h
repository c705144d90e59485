// Drop-in replacement for the reference's util/fle-encoding.h.
//
// Same classes, same public signatures, same observable behaviour as
// /root/reference/fle-encoding.h (FleDecoder :35-240, FleEncoder :242-342) -- but the AVX2
// bodies are gone: packing, unpacking and the encoded-domain predicates run on a B200 through
// the C ABI of libfle_b200.so (include/fle_b200.h).
//
// Granularity adaptation.  The reference is driven value by value (Put/Get) and in <=1024-row
// windows (Eq..In), which is useless as a kernel launch.  The drop-ins therefore work on the
// whole buffer they were constructed on:
//   * the first Get()/Get(skip) decodes a large window of groups in ONE fle_unpack_u32 call
//     and later calls are served from that cache with the reference's cursor arithmetic
//     (count_, buffer_end_, the guard check, including its quirks, fle-encoding.h:344-402);
//   * the first Eq/Lt/Le/Gt/Ge/In with a given literal evaluates the WHOLE buffer in ONE
//     fle_eval call and caches the bitmap; every later window with that literal is a bit
//     slice of the cache.  As in the reference the cursor is not advanced by predicates, the
//     rows left in the currently unpacked group are compared as scalars against the full
//     literal (:7973-7981) and the remaining rows come from whole groups starting at
//     buffer_end_ (:7983-8009), where only the low bit_width bits of the literal take part;
//   * Put() stages values and Flush() packs everything staged since the previous Flush in
//     ONE fle_pack_u32 call, reproducing Put()==false at the buffer guard (:251, :8318) and
//     Flush()'s "always one more group" length (:9806-9811).
// Differences, all outside what the reference defines: unfilled rows of the last group are
// written as 0 (the reference leaves stale staging bytes there, which are never decoded);
// a group that does not fit entirely inside the buffer is treated as the end of data (the
// reference would read past the buffer); predicate rows past the end of the buffer read
// as "no match"; Skip() that lands exactly on the end of the data returns true (FleGpu::SetReferenceCompat(true)
// restores the reference's `false`).
#ifndef IMPALA_FLE_ENCODING_H
#define IMPALA_FLE_ENCODING_H

#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include <boost/dynamic_bitset.hpp>

#include "fle_b200.h"
#include "util/fle-gpu-context.h"

namespace impala {

class FleDecoder {
 public:
  // buffer: FLE groups (fle-encoding.h:37).  Not owned; must outlive the decoder.
  FleDecoder(uint8_t* buffer, int buffer_len, int bit_width)
    : buffer_(buffer), buffer_len_(buffer_len), bit_width_(bit_width), count_(64),
      next_group_(0), cur_group_(-1), win_first_(0), win_groups_(0) {
    num_groups_ = (bit_width_ >= 1 && bit_width_ <= 32 && buffer_len_ > 0)
        ? buffer_len_ / (static_cast<int64_t>(bit_width_) * 8) : 0;
  }

  FleDecoder() : buffer_(NULL), buffer_len_(0), bit_width_(0), count_(64), next_group_(0),
      cur_group_(-1), num_groups_(0), win_first_(0), win_groups_(0) {}

  template<typename T>
  bool Get(T* val);

  template<typename T>
  bool Get(T* val, int skip_rows);

  bool Skip(int skip_rows);

  void Eq(int64_t num_rows, boost::dynamic_bitset<>& skip_bitset, uint64_t value) {
    Predicate(FLE_OP_EQ, num_rows, skip_bitset, &value, 1);
  }
  void Lt(int64_t num_rows, boost::dynamic_bitset<>& skip_bitset, uint64_t value) {
    Predicate(FLE_OP_LT, num_rows, skip_bitset, &value, 1);
  }
  void Le(int64_t num_rows, boost::dynamic_bitset<>& skip_bitset, uint64_t value) {
    Predicate(FLE_OP_LE, num_rows, skip_bitset, &value, 1);
  }
  void Gt(int64_t num_rows, boost::dynamic_bitset<>& skip_bitset, uint64_t value) {
    Predicate(FLE_OP_GT, num_rows, skip_bitset, &value, 1);
  }
  void Ge(int64_t num_rows, boost::dynamic_bitset<>& skip_bitset, uint64_t value) {
    Predicate(FLE_OP_GE, num_rows, skip_bitset, &value, 1);
  }
  void In(int64_t num_rows, boost::dynamic_bitset<>& skip_bitset, std::vector<uint64_t>& values) {
    Predicate(FLE_OP_IN, num_rows, skip_bitset, values.empty() ? NULL : &values[0],
              static_cast<int>(values.size()));
  }
  void Clear() {}

  int bit_width() { return bit_width_; }

  // ---- extensions used by the fused scanner path (not part of the reference class) ----
  const uint8_t* buffer() const { return buffer_; }
  int64_t buffer_len() const { return buffer_len_; }
  // Row index (from the start of the buffer) of the next value Get() would return.
  int64_t position() const {
    return count_ == 64 ? next_group_ * 64 : cur_group_ * 64 + count_;
  }
  // Hands the decoder an already evaluated whole-buffer bitmap for (op, literals), e.g. one
  // produced by a fused multi-column fle_eval; later windows are sliced from it.
  void PrimePredicate(int op, const uint64_t* lits, int nlits, const uint64_t* bitmap_words,
                      int64_t nwords) {
    CachedPred c;
    c.op = op;
    c.lits.assign(lits, lits + nlits);
    c.bitmap.assign(bitmap_words, bitmap_words + nwords);
    pred_cache_.push_back(c);
  }

 private:
  struct CachedPred {
    int op;
    std::vector<uint64_t> lits;
    std::vector<uint64_t> bitmap;  // one bit per row of the whole buffer, LSB first
  };

  // buffer_end_ >= buffer_end_guard_ (fle-encoding.h:350,407) in group units; additionally a
  // group must lie entirely inside the buffer (see header comment).
  bool AtGuard() const {
    return next_group_ * bit_width_ * 8 >= buffer_len_ || next_group_ >= num_groups_;
  }
  void UnpackNext() {  // unpack_function_(); buffer_end_ += bit_width_
    cur_group_ = next_group_;
    ++next_group_;
  }
  bool Advance(int skip_rows);
  bool ValueAt(int64_t group, int idx, uint32_t* out);
  const std::vector<uint64_t>* WholeBufferBitmap(int op, const uint64_t* lits, int nlits);
  void Predicate(int op, int64_t num_rows, boost::dynamic_bitset<>& bs, const uint64_t* lits,
                 int nlits);
  static bool Compare(int op, uint64_t v, const uint64_t* lits, int nlits);

  static const int64_t kWindowGroups = 1 << 18;  // 16 Mi values per decode call

  uint8_t* buffer_;
  int64_t buffer_len_;
  int bit_width_;
  int count_;            // index inside the unpacked group; 64 = nothing unpacked
  int64_t next_group_;   // buffer_end_, as a group index
  int64_t cur_group_;    // group whose values current_value_ would hold
  int64_t num_groups_;   // whole groups inside the buffer
  std::vector<uint32_t> window_;  // decoded values of groups [win_first_, win_first_+win_groups_)
  int64_t win_first_, win_groups_;
  std::vector<CachedPred> pred_cache_;
};

class FleEncoder {
 public:
  // buffer_len is in bytes; the encoder needs room for one spare group (fle-encoding.h:251).
  FleEncoder(uint8_t* buffer, int buffer_len, int bit_width)
    : buffer_(buffer), buffer_words_(buffer_len / 8), bit_width_(bit_width), count_(0),
      groups_done_(0), groups_flushed_(0) {}

  bool Put(uint64_t value) {
    if (count_ == 64) {
      // buffer_end_ >= buffer_end_guard_ (fle-encoding.h:8318), guard = buffer + len/8 - bw
      if (groups_done_ * bit_width_ >= buffer_words_ - bit_width_) return false;
      ++groups_done_;
      count_ = 0;
    }
    staged_.push_back(static_cast<uint32_t>(value));
    ++count_;
    return true;
  }

  // Packs what was staged since the previous Flush (one GPU call) and returns the total
  // encoded byte length so far (fle-encoding.h:9806-9811).
  int Flush() {
    const int64_t first_group = groups_flushed_;
    ++groups_done_;  // Flush always packs the current group, even an empty one
    count_ = 0;
    const int64_t ngroups = groups_done_ - first_group;
    uint8_t* dst = buffer_ + first_group * bit_width_ * 8;
    const size_t cap = static_cast<size_t>(ngroups) * bit_width_ * 8;
    fle_ctx* ctx = FleGpu::ctx();
    if (ctx != NULL) {
      size_t len = 0;
      FleGpu::Check(fle_pack_u32(ctx, staged_.empty() ? NULL : &staged_[0],
                                 static_cast<int64_t>(staged_.size()), bit_width_, dst, cap, &len),
                    "fle_pack_u32");
    }
    staged_.clear();
    groups_flushed_ = groups_done_;
    return static_cast<int>(groups_done_ * bit_width_ * 8);
  }

  void Clear() {
    count_ = 0;
    groups_done_ = 0;
    groups_flushed_ = 0;
    staged_.clear();
  }

  uint8_t* buffer() { return buffer_; }
  int len() { return static_cast<int>(groups_done_ * bit_width_ * 8); }

 private:
  uint8_t* buffer_;
  int64_t buffer_words_;
  int bit_width_;
  int count_;
  int64_t groups_done_;     // (buffer_end_ - buffer_) / bit_width_
  int64_t groups_flushed_;  // groups already written to buffer_
  std::vector<uint32_t> staged_;
};

// ---------------------------------------------------------------------------------------
// FleDecoder implementation
// ---------------------------------------------------------------------------------------
inline bool FleDecoder::Advance(int skip_rows) {  // fle-encoding.h:345-363 == :382-399
  if (count_ == 64) {
    count_ = skip_rows & (64 - 1);
    int skip_64 = (skip_rows & ~(64 - 1)) >> 6;
    next_group_ += skip_64;
    if (AtGuard()) return false;
    UnpackNext();
  } else {
    skip_rows += count_;
    count_ = skip_rows & (64 - 1);
    if (skip_rows >= 64) {
      int skip_64 = ((skip_rows & ~(64 - 1)) >> 6) - 1;
      next_group_ += skip_64;
      if (AtGuard()) return false;
      UnpackNext();
    }
  }
  return true;
}

inline bool FleDecoder::ValueAt(int64_t group, int idx, uint32_t* out) {
  if (group < 0 || group >= num_groups_) return false;
  if (group < win_first_ || group >= win_first_ + win_groups_) {
    fle_ctx* ctx = FleGpu::ctx();
    if (ctx == NULL) return false;
    const int64_t ng = std::min<int64_t>(kWindowGroups, num_groups_ - group);
    window_.resize(static_cast<size_t>(ng) * 64);
    int rc = fle_unpack_u32(ctx, buffer_, static_cast<size_t>(num_groups_) * bit_width_ * 8, bit_width_,
                            group * 64, ng * 64, &window_[0]);
    if (FleGpu::Check(rc, "fle_unpack_u32") != FLE_OK) {
      win_groups_ = 0;
      return false;
    }
    win_first_ = group;
    win_groups_ = ng;
  }
  *out = window_[static_cast<size_t>(group - win_first_) * 64 + idx];
  return true;
}

template<typename T>
inline bool FleDecoder::Get(T* val) {  // fle-encoding.h:404-567
  if (count_ == 64) {
    if (AtGuard()) return false;
    count_ = 0;
    UnpackNext();
  }
  uint32_t v;
  if (!ValueAt(cur_group_, count_, &v)) return false;
  *val = static_cast<T>(v);
  ++count_;
  return true;
}

template<typename T>
inline bool FleDecoder::Get(T* val, int skip_rows) {  // fle-encoding.h:344-379
  if (!Advance(skip_rows)) return false;
  uint32_t v;
  if (!ValueAt(cur_group_, count_, &v)) return false;
  *val = static_cast<T>(v);
  ++count_;
  return true;
}

inline bool FleDecoder::Skip(int skip_rows) {  // :381-402
  if (Advance(skip_rows)) return true;
  if (FleGpu::reference_compat()) return false;  // exactly the reference: false at the end of the data
  // Landing exactly on the end of the data is a successful skip here.  (The reference returns
  // false in that case, and its AssembleRows then silently drops the rest of the row group
  // whenever a page holds a multiple of 64 values, hdfs-parquet-scanner.cc:1107-1113.)
  if (count_ == 0 && next_group_ == num_groups_) {
    count_ = 64;
    return true;
  }
  return false;
}

inline bool FleDecoder::Compare(int op, uint64_t v, const uint64_t* lits, int nlits) {
  switch (op) {
    case FLE_OP_EQ: return v == lits[0];
    case FLE_OP_LT: return v < lits[0];
    case FLE_OP_LE: return v <= lits[0];
    case FLE_OP_GT: return v > lits[0];
    case FLE_OP_GE: return v >= lits[0];
    case FLE_OP_IN: {
      bool flag = false;
      for (int j = 0; j < nlits; ++j) flag = flag || v == lits[j];
      return flag;
    }
  }
  return false;
}

inline const std::vector<uint64_t>* FleDecoder::WholeBufferBitmap(int op, const uint64_t* lits,
                                                                  int nlits) {
  for (size_t i = 0; i < pred_cache_.size(); ++i) {
    const CachedPred& c = pred_cache_[i];
    if (c.op == op && static_cast<int>(c.lits.size()) == nlits &&
        (nlits == 0 || memcmp(&c.lits[0], lits, sizeof(uint64_t) * nlits) == 0)) {
      return &c.bitmap;
    }
  }
  fle_ctx* ctx = FleGpu::ctx();
  if (ctx == NULL || num_groups_ == 0) return NULL;
  if (pred_cache_.size() >= 16) pred_cache_.erase(pred_cache_.begin());
  pred_cache_.push_back(CachedPred());
  CachedPred& c = pred_cache_.back();
  c.op = op;
  c.lits.assign(lits, lits + nlits);
  c.bitmap.assign(static_cast<size_t>(num_groups_), 0);
  fle_col_desc col;
  memset(&col, 0, sizeof(col));
  col.planes = buffer_;
  col.bytes = static_cast<uint64_t>(num_groups_) * bit_width_ * 8;
  col.bit_width = bit_width_;
  fle_pred_node node;
  memset(&node, 0, sizeof(node));
  node.op = op;
  node.col = 0;
  // only the low bit_width bits of a literal take part in the plane algebra (fle-encoding.h:7965-7971)
  const uint64_t lit_mask = bit_width_ >= 64 ? ~uint64_t(0) : ((uint64_t(1) << bit_width_) - 1);
  std::vector<uint64_t> masked(c.lits);
  for (size_t i = 0; i < masked.size(); ++i) masked[i] &= lit_mask;
  node.code = nlits > 0 ? masked[0] : 0;
  node.codes = masked.empty() ? NULL : &masked[0];
  node.ncodes = nlits;
  int rc = fle_eval(ctx, &col, 1, &node, 1, 0, num_groups_ * 64, &c.bitmap[0], NULL);
  if (FleGpu::Check(rc, "fle_eval") != FLE_OK) {
    pred_cache_.pop_back();
    return NULL;
  }
  return &pred_cache_.back().bitmap;
}

inline void FleDecoder::Predicate(int op, int64_t num_rows, boost::dynamic_bitset<>& bs,
                                  const uint64_t* lits, int nlits) {
  // head: rows left in the already unpacked group, scalar compare (fle-encoding.h:7973-7981)
  for (int i = count_; i != 64 && num_rows > 0; ++i, --num_rows) {
    uint32_t v = 0;
    ValueAt(cur_group_, i, &v);
    bs.push_back(Compare(op, v, lits, nlits));
  }
  if (num_rows <= 0) return;
  // body: whole groups from buffer_end_ on (fle-encoding.h:7983-8009)
  const std::vector<uint64_t>* bm = WholeBufferBitmap(op, lits, nlits);
  int64_t g = next_group_;
  const int64_t have = bm == NULL ? 0 : static_cast<int64_t>(bm->size());
  const int64_t full = num_rows / 64;
  const int64_t in_range = std::max<int64_t>(0, std::min(full, have - g));
  if (in_range > 0) bs.append(bm->begin() + g, bm->begin() + g + in_range);
  for (int64_t k = in_range; k < full; ++k) bs.append(0UL);  // past the end of the buffer
  g += full;
  const int tail = static_cast<int>(num_rows - full * 64);
  if (tail > 0) {
    const uint64_t w = (g < have) ? (*bm)[g] : 0;
    for (int i = 0; i < tail; ++i) bs.push_back((w >> i) & 1);
  }
}

}  // namespace impala

#endif
