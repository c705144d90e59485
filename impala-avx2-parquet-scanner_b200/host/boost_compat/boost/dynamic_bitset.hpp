// Stand-in for <boost/dynamic_bitset.hpp>, used ONLY when the real boost is not installed
// (this image has none).  Inside an Impala tree the real header is found first and this
// directory is simply not put on the include path.
//
// It implements the part of boost::dynamic_bitset<unsigned long> that the FLE hot path
// touches (fle-encoding.h:7962-8313, simple-predicates.h:145-163,
// hdfs-parquet-scanner.cc:326-331,1125-1148) with boost's documented semantics: position
// 64*b + i is bit i of block b; append(block) adds 64 positions LSB first;
// append(first, last) adds a range of blocks; resize(n, v) keeps existing bits.
#ifndef FLE_B200_BOOST_DYNAMIC_BITSET_COMPAT_HPP
#define FLE_B200_BOOST_DYNAMIC_BITSET_COMPAT_HPP
#include <cstddef>
#include <cstdint>
#include <vector>

namespace boost {

template <typename Block = unsigned long, typename Allocator = std::allocator<Block> >
class dynamic_bitset {
 public:
  typedef Block block_type;
  typedef std::size_t size_type;
  static const size_type bits_per_block = sizeof(Block) * 8;

  class reference {
   public:
    reference(Block& b, size_type pos) : b_(b), mask_(Block(1) << pos) {}
    operator bool() const { return (b_ & mask_) != 0; }
    bool operator~() const { return (b_ & mask_) == 0; }
    reference& operator=(bool x) {
      if (x) b_ |= mask_; else b_ &= ~mask_;
      return *this;
    }
    reference& operator=(const reference& rhs) { return *this = bool(rhs); }
    reference& operator|=(bool x) { if (x) b_ |= mask_; return *this; }
    reference& operator&=(bool x) { if (!x) b_ &= ~mask_; return *this; }
   private:
    Block& b_;
    const Block mask_;
  };

  dynamic_bitset() : num_bits_(0) {}
  explicit dynamic_bitset(size_type n, unsigned long value = 0) : num_bits_(0) {
    resize(n, false);
    if (n > 0 && value) bits_[0] = static_cast<Block>(value);
    zero_unused_bits();
  }

  size_type size() const { return num_bits_; }
  size_type num_blocks() const { return bits_.size(); }
  bool empty() const { return num_bits_ == 0; }
  void clear() { bits_.clear(); num_bits_ = 0; }

  void resize(size_type n, bool value = false) {
    const size_type old_blocks = bits_.size();
    const size_type need = (n + bits_per_block - 1) / bits_per_block;
    const Block fill = value ? ~Block(0) : Block(0);
    if (need != old_blocks) bits_.resize(need, fill);
    if (value && n > num_bits_) {
      const size_type extra = num_bits_ % bits_per_block;
      if (extra) bits_[old_blocks - 1] |= (fill << extra);
    }
    num_bits_ = n;
    zero_unused_bits();
  }

  void push_back(bool bit) {
    const size_type sz = num_bits_;
    resize(sz + 1);
    if (bit) bits_[sz / bits_per_block] |= Block(1) << (sz % bits_per_block);
  }

  void append(Block value) {
    const size_type r = num_bits_ % bits_per_block;
    if (r == 0) {
      bits_.push_back(value);
    } else {
      bits_.push_back(value >> (bits_per_block - r));
      bits_[bits_.size() - 2] |= (value << r);
    }
    num_bits_ += bits_per_block;
  }

  template <typename BlockInputIterator>
  void append(BlockInputIterator first, BlockInputIterator last) {
    for (; first != last; ++first) append(static_cast<Block>(*first));
  }

  reference operator[](size_type pos) { return reference(bits_[pos / bits_per_block], pos % bits_per_block); }
  bool operator[](size_type pos) const { return test(pos); }
  bool test(size_type pos) const { return (bits_[pos / bits_per_block] >> (pos % bits_per_block)) & 1; }
  dynamic_bitset& set(size_type pos, bool val = true) { (*this)[pos] = val; return *this; }

  size_type count() const {
    size_type c = 0;
    for (size_type i = 0; i < bits_.size(); ++i) c += __builtin_popcountll(bits_[i]);
    return c;
  }
  bool any() const {
    for (size_type i = 0; i < bits_.size(); ++i) if (bits_[i]) return true;
    return false;
  }
  bool none() const { return !any(); }

  dynamic_bitset& operator&=(const dynamic_bitset& rhs) {
    for (size_type i = 0; i < bits_.size(); ++i) bits_[i] &= rhs.bits_[i];
    return *this;
  }
  dynamic_bitset& operator|=(const dynamic_bitset& rhs) {
    for (size_type i = 0; i < bits_.size(); ++i) bits_[i] |= rhs.bits_[i];
    return *this;
  }
  bool operator==(const dynamic_bitset& rhs) const { return num_bits_ == rhs.num_bits_ && bits_ == rhs.bits_; }

  template <typename B, typename A, typename It>
  friend void to_block_range(const dynamic_bitset<B, A>& b, It result);

 private:
  void zero_unused_bits() {
    const size_type extra = num_bits_ % bits_per_block;
    if (extra) bits_.back() &= ~(~Block(0) << extra);
  }
  std::vector<Block, Allocator> bits_;
  size_type num_bits_;
};

template <typename B, typename A, typename It>
inline void to_block_range(const dynamic_bitset<B, A>& b, It result) {
  for (std::size_t i = 0; i < b.bits_.size(); ++i) *result++ = b.bits_[i];
}

}  // namespace boost
#endif
