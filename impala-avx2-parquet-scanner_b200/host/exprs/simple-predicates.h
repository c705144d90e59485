// Drop-in for the reference's exprs/simple-predicates.h (:30-205): the tree the planner's
// conjuncts are lowered to by Expr::CreateSimplePredicates (expr.h:234,
// scalar-fn-call.cc:733-965, compound-predicates.cc:36-67).
//
// GetBitset() keeps the reference signature and semantics: children are evaluated eagerly
// into fresh bitsets and combined with `&=` / `|=` (:145-163); leaves forward to the scanner's
// typed trampolines (:165-205).  Each leaf call is served by the GPU through the column
// reader (whole page evaluated once, windows sliced).
//
// New: Lower() flattens the same tree into the postfix program of the C ABI so that the
// scanner can evaluate ALL conjuncts over ALL columns in one fused kernel pass
// (fle_eval, include/fle_b200.h) instead of one pass per leaf.
#ifndef IMPALA_EXPRS_SIMPLE_PREDICATES_H_
#define IMPALA_EXPRS_SIMPLE_PREDICATES_H_

#include <stdint.h>

#include <list>
#include <vector>

#include <boost/dynamic_bitset.hpp>

#include "fle_b200.h"

namespace impala {

class HdfsParquetScanner;

// Postfix program under construction + the storage its IN lists point into.
struct FusedProgram {
  std::vector<fle_pred_node> nodes;
  std::list<std::vector<uint64_t> > in_codes;  // stable addresses
  void Push(int op, int col = 0, uint64_t code = 0) {
    fle_pred_node n;
    n.op = op; n.col = col; n.code = code; n.codes = NULL; n.ncodes = 0; n.reserved = 0;
    nodes.push_back(n);
  }
  void PushIn(int col, const std::vector<uint64_t>& codes) {
    in_codes.push_back(codes);
    fle_pred_node n;
    n.op = FLE_OP_IN; n.col = col; n.code = 0;
    n.codes = in_codes.back().empty() ? NULL : &in_codes.back()[0];
    n.ncodes = static_cast<int32_t>(in_codes.back().size());
    n.reserved = 0;
    nodes.push_back(n);
  }
};

class SimplePredicate {
 public:
  virtual void GetBitset(HdfsParquetScanner* scanner, int64_t num_rows,
      boost::dynamic_bitset<>& skip_bitset) = 0;
  // Appends this subtree to prog in postfix order.  Returns false if some leaf cannot be
  // expressed on encoded data (the scanner then keeps the per-leaf path).
  virtual bool Lower(HdfsParquetScanner* scanner, FusedProgram* prog) = 0;
  virtual ~SimplePredicate() { }
};

class AndOperate : public SimplePredicate {
 public:
  virtual void GetBitset(HdfsParquetScanner* scanner, int64_t num_rows,
      boost::dynamic_bitset<>& skip_bitset) {
    child0_->GetBitset(scanner, num_rows, skip_bitset);
    boost::dynamic_bitset<> tmp_bitset;
    child1_->GetBitset(scanner, num_rows, tmp_bitset);
    skip_bitset &= tmp_bitset;
  }
  virtual bool Lower(HdfsParquetScanner* scanner, FusedProgram* prog) {
    if (!child0_->Lower(scanner, prog) || !child1_->Lower(scanner, prog)) return false;
    prog->Push(FLE_OP_AND);
    return true;
  }
  AndOperate(SimplePredicate* child0, SimplePredicate* child1) : child0_(child0), child1_(child1) { }
 private:
  SimplePredicate* child0_;
  SimplePredicate* child1_;
};

class OrOperate : public SimplePredicate {
 public:
  virtual void GetBitset(HdfsParquetScanner* scanner, int64_t num_rows,
      boost::dynamic_bitset<>& skip_bitset) {
    child0_->GetBitset(scanner, num_rows, skip_bitset);
    boost::dynamic_bitset<> tmp_bitset;
    child1_->GetBitset(scanner, num_rows, tmp_bitset);
    skip_bitset |= tmp_bitset;
  }
  virtual bool Lower(HdfsParquetScanner* scanner, FusedProgram* prog) {
    if (!child0_->Lower(scanner, prog) || !child1_->Lower(scanner, prog)) return false;
    prog->Push(FLE_OP_OR);
    return true;
  }
  OrOperate(SimplePredicate* child0, SimplePredicate* child1) : child0_(child0), child1_(child1) { }
 private:
  SimplePredicate* child0_;
  SimplePredicate* child1_;
};

#define FLE_DECLARE_LEAF(Name)                                                            \
  template <typename T>                                                                   \
  class Name##Operate : public SimplePredicate {                                          \
   public:                                                                                \
    virtual void GetBitset(HdfsParquetScanner* scanner, int64_t num_rows,                 \
        boost::dynamic_bitset<>& skip_bitset);                                            \
    virtual bool Lower(HdfsParquetScanner* scanner, FusedProgram* prog);                  \
    Name##Operate(int idx, T val) : idx_(idx), val_(val) { }                              \
   private:                                                                               \
    int idx_;                                                                             \
    T val_;                                                                               \
  };
FLE_DECLARE_LEAF(Eq)
FLE_DECLARE_LEAF(Lt)
FLE_DECLARE_LEAF(Le)
FLE_DECLARE_LEAF(Gt)
FLE_DECLARE_LEAF(Ge)
#undef FLE_DECLARE_LEAF

template <typename T>
class InOperate : public SimplePredicate {
 public:
  virtual void GetBitset(HdfsParquetScanner* scanner, int64_t num_rows,
      boost::dynamic_bitset<>& skip_bitset);
  virtual bool Lower(HdfsParquetScanner* scanner, FusedProgram* prog);
  InOperate(int idx, std::vector<T> vals) : idx_(idx), vals_(vals.begin(), vals.end()) { }
 private:
  int idx_;
  std::vector<T> vals_;
};

}  // namespace impala

// The leaves' bodies need the scanner's typed trampolines.
#include "exec/hdfs-parquet-scanner.h"

#endif
