// Build shim (test infrastructure): the subset of boost::dynamic_bitset<unsigned long>
// the reference's hot path relies on.  The one semantic that matters for parity is the
// block order boost documents: bit i of block b is position 64*b + i, and
// append(block) adds 64 positions, least-significant bit first.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>
namespace boost {
template <class Block = unsigned long>
class dynamic_bitset {
 public:
  class reference {
   public:
    reference(Block* b, unsigned s) : b_(b), s_(s) {}
    operator bool() const { return (*b_ >> s_) & 1; }
    reference& operator=(bool v) {
      if (v) *b_ |= (Block(1) << s_); else *b_ &= ~(Block(1) << s_);
      return *this;
    }
    reference& operator=(const reference& o) { return *this = bool(o); }
   private:
    Block* b_;
    unsigned s_;
  };
  dynamic_bitset() : n_(0) {}
  std::size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }
  void clear() { w_.clear(); n_ = 0; }
  void push_back(bool v) {
    if ((n_ & 63) == 0) w_.push_back(0);
    if (v) w_[n_ >> 6] |= Block(1) << (n_ & 63);
    ++n_;
  }
  void append(Block blk) {
    unsigned r = n_ & 63;
    if (r == 0) {
      w_.push_back(blk);
    } else {
      w_.back() |= blk << r;
      w_.push_back(blk >> (64 - r));
    }
    n_ += 64;
  }
  void resize(std::size_t n, bool v = false) {
    std::size_t old = n_;
    w_.resize((n + 63) / 64, v ? ~Block(0) : Block(0));
    n_ = n;
    if (n > old && (old & 63)) {
      Block m = ~Block(0) << (old & 63);
      if (v) w_[old >> 6] |= m; else w_[old >> 6] &= ~m;
    }
    trim();
  }
  std::size_t count() const {
    std::size_t c = 0;
    for (Block b : w_) c += __builtin_popcountll(b);
    return c;
  }
  bool operator[](std::size_t i) const { return (w_[i >> 6] >> (i & 63)) & 1; }
  reference operator[](std::size_t i) { return reference(&w_[i >> 6], i & 63); }
  dynamic_bitset& operator&=(const dynamic_bitset& o) {
    for (std::size_t i = 0; i < w_.size(); ++i) w_[i] &= o.w_[i];
    return *this;
  }
  dynamic_bitset& operator|=(const dynamic_bitset& o) {
    for (std::size_t i = 0; i < w_.size(); ++i) w_[i] |= o.w_[i];
    return *this;
  }
  std::size_t num_blocks() const { return w_.size(); }
  template <class B, class It>
  friend void to_block_range(const dynamic_bitset<B>& b, It result);
 private:
  void trim() {
    if (n_ & 63) w_.back() &= ~(~Block(0) << (n_ & 63));
  }
  std::vector<Block> w_;
  std::size_t n_;
};
template <class B, class It>
inline void to_block_range(const dynamic_bitset<B>& b, It result) {
  for (std::size_t i = 0; i < b.w_.size(); ++i) *result++ = b.w_[i];
}
}  // namespace boost
