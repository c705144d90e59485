// Drop-in for the reference's util/dict-encoding.h, restricted to what the FLE hot path uses.
//
//   DictEncoder<T>   /root/reference/dict-encoding.h:113-177, 234-308, 393-447
//       hash-deduplicating dictionary, capped at 40000 entries (:157); WriteDict() SORTS the
//       dictionary by value and remaps the buffered indices (:394-406, :408-411), which is
//       what makes dictionary codes order-preserving; WriteData() = [1 byte bit width][FLE]
//       (:408-447).  The FLE packing itself runs on the GPU (FleEncoder -> fle_pack_u32).
//   DictDecoder<T>   /root/reference/dict-encoding.h:182-232, 310-336, 449-541
//       SetData() strips the bit-width byte (:185-192); GetValue/SkipValue walk the codes;
//       Eq/Gt/Lt/Ge/Le/In translate a typed literal into a dictionary CODE with
//       lower_/upper_bound on the sorted dictionary and hand it to the FLE predicates
//       (:461-541) -- "literals are pre-encoded on the host".
//
// The Impala runtime types the reference drags in (MemPool, HashUtil, StringValue,
// TimestampValue, DecimalNValue) are not part of the hot path; T is any trivially copyable,
// totally ordered value type (int8..int64, float, double, ...).  PLAIN dictionary-page
// encoding of such a T is its sizeof(T) little-endian bytes (parquet-common.h:169-195).
#ifndef IMPALA_UTIL_DICT_ENCODING_H
#define IMPALA_UTIL_DICT_ENCODING_H

#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <memory>
#include <vector>

#include <boost/dynamic_bitset.hpp>

#include "util/fle-encoding.h"

namespace impala {

// BitUtil::Log2 = ceil(log2(x)) (bit-util.h:132-143)
inline int FleLog2Ceil(uint64_t x) {
  if (x <= 1) return 0;
  --x;
  int result = 1;
  while (x >>= 1) ++result;
  return result;
}

class DictEncoderBase {
 public:
  virtual ~DictEncoderBase() {}
  virtual void WriteDict(uint8_t* buffer) = 0;
  virtual int num_entries() const = 0;
  void ClearIndices() { buffered_indices_.clear(); }

  // dict-encoding.h:76-80
  int bit_width() const {
    if (num_entries() == 0) return 0;
    if (num_entries() == 1) return 1;
    return FleLog2Ceil(num_entries());
  }

  // Bytes WriteData() needs for the buffered indices: 1 + one spare group (fle-encoding.h:251).
  int EstimatedDataEncodedSize() const {
    const int64_t groups = (static_cast<int64_t>(buffered_indices_.size()) + 63) / 64 + 2;
    return static_cast<int>(1 + groups * bit_width() * 8);
  }

  // dict-encoding.h:408-423.  Returns bytes written, -1 if the buffer is too small.
  int WriteData(uint8_t* buffer, int buffer_len) {
    for (size_t i = 0; i < buffered_indices_.size(); ++i) {
      buffered_indices_[i] = to_sorted_indice_[buffered_indices_[i]];
    }
    *buffer = static_cast<uint8_t>(bit_width());
    ++buffer;
    --buffer_len;
    FleEncoder encoder(buffer, buffer_len, bit_width());
    for (size_t i = 0; i < buffered_indices_.size(); ++i) {
      if (!encoder.Put(buffered_indices_[i])) return -1;
    }
    encoder.Flush();
    return 1 + encoder.len();
  }

  // dict-encoding.h:425-447 (per-page index lists kept by the table writer).
  int WriteData(uint8_t* buffer, int buffer_len, std::vector<int>& node_indices) {
    --buffer_len;
    int max_indice = -1;
    FleEncoder encoder(buffer + 1, buffer_len, bit_width());
    for (size_t i = 0; i < node_indices.size(); ++i) {
      const int index = node_indices[i];
      if (!encoder.Put(to_sorted_indice_[index])) return -1;
      if (index > max_indice) max_indice = index;
    }
    // Deviation: the reference stores Log2(max_indice + 1) in the header byte
    // (:438-444) although the data is packed with bit_width() (:432); a page that has not
    // seen the newest dictionary entries would then be decoded with the wrong width.  The
    // header written here is the width the data was actually packed with.
    if (FleGpu::reference_compat()) {  // byte for byte what the reference writes (:438-444)
      *buffer = max_indice == -1 ? 0 : (max_indice == 0 ? 1 : static_cast<uint8_t>(FleLog2Ceil(max_indice + 1)));
    } else {
      *buffer = static_cast<uint8_t>(bit_width());
    }
    encoder.Flush();
    return 1 + encoder.len();
  }

  int dict_encoded_size() { return dict_encoded_size_; }

 protected:
  DictEncoderBase() : dict_encoded_size_(0) {}
  std::vector<int> buffered_indices_;
  std::vector<int> to_sorted_indice_;
  int dict_encoded_size_;
};

template<typename T>
class DictEncoder : public DictEncoderBase {
 public:
  enum { MAX_ENTRIES = 40000 };  // Node::INVALID_INDEX, dict-encoding.h:157

  explicit DictEncoder(int encoded_value_size = sizeof(T)) : encoded_value_size_(encoded_value_size) {}

  // Returns bytes added to the dictionary page (0 if already present), -1 when full
  // (dict-encoding.h:255-274).
  int Put(const T& value) {
    int idx;
    int added = Lookup(value, &idx);
    if (added < 0) return -1;
    buffered_indices_.push_back(idx);
    return added;
  }
  int Put(const T& value, std::vector<int>& node_indices) {  // :234-253
    int idx;
    int added = Lookup(value, &idx);
    if (added < 0) return -1;
    node_indices.push_back(idx);
    return added;
  }

  // Sorts the dictionary, fills to_sorted_indice_, writes PLAIN values (:393-406).
  virtual void WriteDict(uint8_t* buffer) {
    const int n = static_cast<int>(values_.size());
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [this](int a, int b) { return values_[a] < values_[b]; });
    to_sorted_indice_.resize(n);
    for (int i = 0; i < n; ++i) to_sorted_indice_[order[i]] = i;
    for (int i = 0; i < n; ++i) {
      memcpy(buffer, &values_[order[i]], sizeof(T));
      buffer += encoded_value_size_;
    }
  }

  virtual int num_entries() const { return static_cast<int>(values_.size()); }

 private:
  int Lookup(const T& value, int* idx) {
    typename std::map<T, int>::iterator it = index_.find(value);
    if (it != index_.end()) {
      *idx = it->second;
      return 0;
    }
    if (values_.size() >= MAX_ENTRIES) return -1;
    *idx = static_cast<int>(values_.size());
    values_.push_back(value);
    index_[value] = *idx;
    dict_encoded_size_ += encoded_value_size_;
    return encoded_value_size_;
  }

  std::vector<T> values_;    // insertion order == unsorted dictionary index
  std::map<T, int> index_;
  int encoded_value_size_;
};

class DictDecoderBase {
 public:
  // dict-encoding.h:185-192: first byte = bit width, the rest = FLE groups.
  void SetData(uint8_t* buffer, int buffer_len) {
    uint8_t bit_width = *buffer;
    ++buffer;
    --buffer_len;
    data_decoder_.reset(new FleDecoder(buffer, buffer_len, bit_width));
  }
  virtual ~DictDecoderBase() {}
  virtual int num_entries() const = 0;
  FleDecoder* data_decoder() { return data_decoder_.get(); }

 protected:
  std::unique_ptr<FleDecoder> data_decoder_;
};

// Result of translating a typed literal: which FLE op to run with which code, or a constant.
struct DictCode {
  int op;         // FLE_OP_EQ / LT / GE / IN, or FLE_OP_ALL / FLE_OP_NONE
  uint64_t code;
  std::vector<uint64_t> codes;  // IN
};

template<typename T>
class DictDecoder : public DictDecoderBase {
 public:
  // dict_buffer: PLAIN-encoded, sorted dictionary page (dict-encoding.h:449-459).
  DictDecoder(uint8_t* dict_buffer, int dict_len, int fixed_len_size = sizeof(T)) {
    const int step = fixed_len_size > 0 ? fixed_len_size : static_cast<int>(sizeof(T));
    for (int off = 0; off + static_cast<int>(sizeof(T)) <= dict_len; off += step) {
      T value;
      memcpy(&value, dict_buffer + off, sizeof(T));
      dict_.push_back(value);
    }
  }

  virtual int num_entries() const { return static_cast<int>(dict_.size()); }
  const std::vector<T>& dict() const { return dict_; }

  bool GetValue(T* value) {  // :310-319
    int index;
    if (!data_decoder_->Get(&index)) return false;
    if (index >= static_cast<int>(dict_.size())) return false;
    *value = dict_[index];
    return true;
  }
  bool GetValue(T* value, int skip_rows) {  // :321-330
    int index;
    if (!data_decoder_->Get(&index, skip_rows)) return false;
    if (index >= static_cast<int>(dict_.size())) return false;
    *value = dict_[index];
    return true;
  }
  bool SkipValue(int skip_rows) { return data_decoder_->Skip(skip_rows); }  // :332-336

  // ---- literal -> code translation, shared by the per-leaf and the fused paths ----
  DictCode TranslateEq(const T& val) const {  // :461-470
    DictCode r = {FLE_OP_NONE, 0, {}};
    typename std::vector<T>::const_iterator it = std::lower_bound(dict_.begin(), dict_.end(), val);
    if (it == dict_.end() || val < *it) return r;
    r.op = FLE_OP_EQ;
    r.code = std::distance(dict_.begin(), it);
    return r;
  }
  DictCode TranslateGt(const T& val) const {  // :472-483
    DictCode r = {FLE_OP_NONE, 0, {}};
    if (dict_.empty() || dict_.back() <= val) return r;
    if (dict_[0] > val) { r.op = FLE_OP_ALL; return r; }
    r.op = FLE_OP_GE;
    r.code = std::distance(dict_.begin(), std::upper_bound(dict_.begin(), dict_.end(), val));
    return r;
  }
  DictCode TranslateLt(const T& val) const {  // :485-496
    DictCode r = {FLE_OP_NONE, 0, {}};
    if (dict_.empty() || dict_[0] >= val) return r;
    if (dict_.back() < val) { r.op = FLE_OP_ALL; return r; }
    r.op = FLE_OP_LT;
    r.code = std::distance(dict_.begin(), std::lower_bound(dict_.begin(), dict_.end(), val));
    return r;
  }
  DictCode TranslateGe(const T& val) const {  // :498-509
    DictCode r = {FLE_OP_NONE, 0, {}};
    if (dict_.empty() || dict_.back() < val) return r;
    if (dict_[0] >= val) { r.op = FLE_OP_ALL; return r; }
    r.op = FLE_OP_GE;
    r.code = std::distance(dict_.begin(), std::lower_bound(dict_.begin(), dict_.end(), val));
    return r;
  }
  DictCode TranslateLe(const T& val) const {  // :511-521
    DictCode r = {FLE_OP_NONE, 0, {}};
    if (dict_.empty() || dict_[0] > val) return r;
    if (dict_.back() <= val) { r.op = FLE_OP_ALL; return r; }
    r.op = FLE_OP_LT;
    r.code = std::distance(dict_.begin(), std::upper_bound(dict_.begin(), dict_.end(), val));
    return r;
  }
  DictCode TranslateIn(const std::vector<T>& vals) const {  // :523-541
    DictCode r = {FLE_OP_NONE, 0, {}};
    for (size_t i = 0; i < vals.size(); ++i) {
      typename std::vector<T>::const_iterator it = std::lower_bound(dict_.begin(), dict_.end(), vals[i]);
      if (it == dict_.end() || vals[i] < *it) continue;
      r.codes.push_back(std::distance(dict_.begin(), it));
    }
    if (!r.codes.empty()) r.op = FLE_OP_IN;
    return r;
  }

  // ---- the reference's per-leaf interface (:224-229) ----
  void Eq(int64_t num_rows, boost::dynamic_bitset<>& skip_bitset, T& val) { Run(TranslateEq(val), num_rows, skip_bitset); }
  void Gt(int64_t num_rows, boost::dynamic_bitset<>& skip_bitset, T& val) { Run(TranslateGt(val), num_rows, skip_bitset); }
  void Lt(int64_t num_rows, boost::dynamic_bitset<>& skip_bitset, T& val) { Run(TranslateLt(val), num_rows, skip_bitset); }
  void Ge(int64_t num_rows, boost::dynamic_bitset<>& skip_bitset, T& val) { Run(TranslateGe(val), num_rows, skip_bitset); }
  void Le(int64_t num_rows, boost::dynamic_bitset<>& skip_bitset, T& val) { Run(TranslateLe(val), num_rows, skip_bitset); }
  void In(int64_t num_rows, boost::dynamic_bitset<>& skip_bitset, std::vector<T>& vals) {
    Run(TranslateIn(vals), num_rows, skip_bitset);
  }

 private:
  void Run(const DictCode& c, int64_t num_rows, boost::dynamic_bitset<>& bs) {
    switch (c.op) {
      case FLE_OP_NONE: bs.resize(num_rows, false); break;  // e.g. :466
      case FLE_OP_ALL: bs.resize(num_rows, true); break;    // e.g. :478
      case FLE_OP_EQ: data_decoder_->Eq(num_rows, bs, c.code); break;
      case FLE_OP_LT: data_decoder_->Lt(num_rows, bs, c.code); break;
      case FLE_OP_GE: data_decoder_->Ge(num_rows, bs, c.code); break;
      case FLE_OP_IN: {
        std::vector<uint64_t> codes(c.codes);
        data_decoder_->In(num_rows, bs, codes);
        break;
      }
    }
  }

  std::vector<T> dict_;
};

}  // namespace impala
#endif
