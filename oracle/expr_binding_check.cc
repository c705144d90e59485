// TEST INFRASTRUCTURE ONLY.  Compile-and-run check of the Expr-side binding of the drop-in:
// the reference lowers a conjunct to the SimplePredicate tree in
//   ScalarFnCall::CreateSimplePredicates / CreateOperate<T>   scalar-fn-call.cc:733-965
//   AndPredicate / OrPredicate::CreateSimplePredicates        compound-predicates.cc:36-44, :57-65
// Those function bodies are NOT copied into this repository: oracle/Makefile extracts them from where they
// lie under /root/reference into oracle/_ref/expr_binding_ref.inc (git-ignored) and this file compiles them,
// UNMODIFIED, against minimal stand-ins for the Impala classes they touch (Expr, SlotRef, Literal,
// HdfsScanNode, RuntimeState, ObjectPool, descriptors, the *Val structs) and against the DROP-IN
// host/exprs/simple-predicates.h -- i.e. the literal-type switch, the function names ("eq" "gt" "lt"
// "ge" "le" "in_set_lookup"), the Eq..InOperate<T>(slot_idx, value) constructors and And/OrOperate(child0,
// child1) all bind to the drop-in as they are.  main() then builds the expression trees of three
// conjuncts with the stand-ins, runs the reference's lowering on them and checks that the drop-in's own
// Lower() turns the resulting tree into the postfix program of the C ABI that the conjuncts mean.
// No GPU is touched (Lower() only translates literals against the dictionaries on the host).
#include <stdint.h>
#include <string.h>

#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "exprs/simple-predicates.h"  // the drop-in (pulls in exec/hdfs-parquet-scanner.h)

using std::string;
using std::stringstream;
using std::vector;

// ---- glog-flavoured checks ---------------------------------------------------------------------------
struct NullStream {
  template <typename T> NullStream& operator<<(const T&) { return *this; }
};
#define DCHECK(x) if (!(x)) std::abort(); else NullStream()
#define DCHECK_EQ(a, b) DCHECK((a) == (b))
#define DCHECK_LT(a, b) DCHECK((a) < (b))

namespace impala {

enum PrimitiveType { INVALID_TYPE = 0, TYPE_NULL, TYPE_BOOLEAN, TYPE_TINYINT, TYPE_SMALLINT, TYPE_INT, TYPE_BIGINT, TYPE_FLOAT,
                     TYPE_DOUBLE, TYPE_TIMESTAMP, TYPE_STRING, TYPE_DATE, TYPE_DATETIME, TYPE_BINARY, TYPE_DECIMAL, TYPE_CHAR,
                     TYPE_VARCHAR };
struct ColumnType {
  PrimitiveType type;
  int byte_size;
  ColumnType(PrimitiveType t = INVALID_TYPE, int bytes = 0) : type(t), byte_size(bytes) {}
  int GetByteSize() const { return byte_size; }
  string DebugString() const { return "type"; }
};

// udf.h value structs, reduced to what the call sites read
struct AnyVal { bool is_null; AnyVal() : is_null(false) {} };
struct BooleanVal : AnyVal { bool val; BooleanVal(bool v = false) : val(v) {} static BooleanVal null() { BooleanVal b; b.is_null = true; return b; } };
struct TinyIntVal : AnyVal { int8_t val; TinyIntVal(int8_t v = 0) : val(v) {} };
struct SmallIntVal : AnyVal { int16_t val; SmallIntVal(int16_t v = 0) : val(v) {} };
struct IntVal : AnyVal { int32_t val; IntVal(int32_t v = 0) : val(v) {} };
struct BigIntVal : AnyVal { int64_t val; BigIntVal(int64_t v = 0) : val(v) {} };
struct FloatVal : AnyVal { float val; FloatVal(float v = 0) : val(v) {} };
struct DoubleVal : AnyVal { double val; DoubleVal(double v = 0) : val(v) {} };
struct StringVal : AnyVal { const char* ptr; int len; StringVal() : ptr(NULL), len(0) {} };
struct TimestampVal : AnyVal { int32_t date; int64_t time_of_day; TimestampVal() : date(0), time_of_day(0) {} };
struct DecimalVal : AnyVal { int64_t val8; DecimalVal() : val8(0) {} };

// slot types the literal switch can produce: plain PODs with the comparisons DictDecoder<T> needs
template <int N>
struct FixedValue {
  int64_t v[N];
  bool operator<(const FixedValue& o) const { return memcmp(v, o.v, sizeof(v)) < 0; }
  bool operator==(const FixedValue& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
  bool operator<=(const FixedValue& o) const { return !(o < *this); }
  bool operator>(const FixedValue& o) const { return o < *this; }
  bool operator>=(const FixedValue& o) const { return !(*this < o); }
  bool operator!=(const FixedValue& o) const { return !(*this == o); }
};
struct Decimal4Value : FixedValue<1> { static Decimal4Value FromDouble(const ColumnType&, double d, bool*) { Decimal4Value r; r.v[0] = (int64_t)d; return r; } };
struct Decimal8Value : FixedValue<1> { static Decimal8Value FromDouble(const ColumnType&, double d, bool*) { Decimal8Value r; r.v[0] = (int64_t)d; return r; } };
struct Decimal16Value : FixedValue<2> { static Decimal16Value FromDouble(const ColumnType&, double d, bool*) { Decimal16Value r; r.v[0] = (int64_t)d; r.v[1] = 0; return r; } };
struct StringValue : FixedValue<2> { static StringValue FromStringVal(const StringVal& s) { StringValue r; r.v[0] = (int64_t)(intptr_t)s.ptr; r.v[1] = s.len; return r; } };
struct TimestampValue : FixedValue<2> { static TimestampValue FromTimestampVal(const TimestampVal& t) { TimestampValue r; r.v[0] = t.date; r.v[1] = t.time_of_day; return r; } };
inline Decimal4Value GetDecimal4Value(const DecimalVal& v, const ColumnType&, bool*) { Decimal4Value r; r.v[0] = v.val8; return r; }
inline Decimal8Value GetDecimal8Value(const DecimalVal& v, const ColumnType&, bool*) { Decimal8Value r; r.v[0] = v.val8; return r; }
inline Decimal16Value GetDecimal16Value(const DecimalVal& v, const ColumnType&, bool*) { Decimal16Value r; r.v[0] = v.val8; r.v[1] = 0; return r; }

// runtime / descriptors
typedef int SlotId;
struct SlotDescriptor {
  vector<int> path;
  const vector<int>& col_path() const { return path; }
};
struct DescriptorTbl {
  vector<SlotDescriptor> slots;
  const SlotDescriptor* GetSlotDescriptor(SlotId id) const { return &slots[id]; }
};
struct ObjectPool {
  vector<SimplePredicate*> owned;
  template <typename T> T* Add(T* t) { owned.push_back(t); return t; }
  ~ObjectPool() { for (size_t i = 0; i < owned.size(); ++i) delete owned[i]; }
};
struct RuntimeState {
  DescriptorTbl tbl;
  ObjectPool pool;
  const DescriptorTbl& desc_tbl() const { return tbl; }
  ObjectPool* obj_pool() { return &pool; }
};
struct HdfsScanNode {
  RuntimeState state;
  RuntimeState* runtime_state() { return &state; }
  // the materialized-slot index of a column path = the column-reader index of the scanner (scalar-fn-call.cc:940-942)
  int GetMaterializedSlotIdx(const vector<int>& path) const { return path[0]; }
};

class ExprContext;
class TupleRow;
struct TFunctionName { string function_name; };
struct TScalarFunction { string symbol; };
struct TFunction { TFunctionName name; int binary_type; string hdfs_location; TScalarFunction scalar_fn; };

class Expr {
 public:
  Expr(const ColumnType& t, bool slotref = false) : type_(t), is_slotref_(slotref) {}
  virtual ~Expr() {}
  virtual SimplePredicate* CreateSimplePredicates(HdfsScanNode*) { return NULL; }  // expr.h:234
  bool is_slotref() const { return is_slotref_; }
  const ColumnType& type() const { return type_; }
  string DebugString() const { return "expr"; }
  void AddChild(Expr* e) { children_.push_back(e); }
  vector<Expr*> children_;
 protected:
  ColumnType type_;
  bool is_slotref_;
};
class SlotRef : public Expr {
 public:
  SlotRef(const ColumnType& t, SlotId id) : Expr(t, true), slot_id_(id) {}
  SlotId slot_id() const { return slot_id_; }
 private:
  SlotId slot_id_;
};
class Literal : public Expr {
 public:
  Literal(PrimitiveType t, int64_t i, double d = 0) : Expr(ColumnType(t)), i_(i), d_(d) {}
  TinyIntVal GetTinyIntVal(ExprContext*, TupleRow*) { return TinyIntVal((int8_t)i_); }
  SmallIntVal GetSmallIntVal(ExprContext*, TupleRow*) { return SmallIntVal((int16_t)i_); }
  IntVal GetIntVal(ExprContext*, TupleRow*) { return IntVal((int32_t)i_); }
  BigIntVal GetBigIntVal(ExprContext*, TupleRow*) { return BigIntVal(i_); }
  FloatVal GetFloatVal(ExprContext*, TupleRow*) { return FloatVal((float)d_); }
  DoubleVal GetDoubleVal(ExprContext*, TupleRow*) { return DoubleVal(d_); }
  StringVal GetStringVal(ExprContext*, TupleRow*) { return StringVal(); }
  TimestampVal GetTimestampVal(ExprContext*, TupleRow*) { return TimestampVal(); }
  DecimalVal GetDecimalVal(ExprContext*, TupleRow*) { DecimalVal v; v.val8 = i_; return v; }
 private:
  int64_t i_;
  double d_;
};

// the three classes whose CreateSimplePredicates bodies come from the reference (declarations as in
// scalar-fn-call.h:81,130-131 and compound-predicates.h:48,73)
class ScalarFnCall : public Expr {
 public:
  ScalarFnCall(const string& fn_name) : Expr(ColumnType(TYPE_BOOLEAN)) { fn_.name.function_name = fn_name; fn_.binary_type = 0; }
  virtual SimplePredicate* CreateSimplePredicates(HdfsScanNode* scan_node);
  string DebugString() const;
 private:
  template <typename T>
  SimplePredicate* CreateOperate(HdfsScanNode* scan_node, SlotRef* slotref, vector<T>& val);
  TFunction fn_;
};
class AndPredicate : public Expr {
 public:
  AndPredicate() : Expr(ColumnType(TYPE_BOOLEAN)) {}
  virtual SimplePredicate* CreateSimplePredicates(HdfsScanNode* scan_node);
};
class OrPredicate : public Expr {
 public:
  OrPredicate() : Expr(ColumnType(TYPE_BOOLEAN)) {}
  virtual SimplePredicate* CreateSimplePredicates(HdfsScanNode* scan_node);
};

// ---- the reference's own function bodies, extracted at build time, unmodified -------------------------------
#include "_ref/expr_binding_ref.inc"

}  // namespace impala

using namespace impala;

static int failures = 0;
#define CHECK_EQ(a, b) do { if (!((a) == (b))) { ++failures; std::fprintf(stderr, "%s:%d: %s != %s (%lld vs %lld)\n", __FILE__, __LINE__, #a, #b, (long long)(a), (long long)(b)); } } while (0)

static Expr* Compare(const char* fn, PrimitiveType col_type, SlotId slot, PrimitiveType lit_type, int64_t lit, bool cast = false) {
  ScalarFnCall* call = new ScalarFnCall(fn);
  Expr* ref = new SlotRef(ColumnType(col_type, 4), slot);
  if (cast) {  // cast(SlotRef): a one-child expr over the slot (scalar-fn-call.cc:736-739)
    Expr* c = new Expr(ColumnType(lit_type));
    c->AddChild(ref);
    ref = c;
  }
  call->AddChild(ref);
  call->AddChild(new Literal(lit_type, lit));
  return call;
}

int main() {
  HdfsScanNode node;
  for (int c = 0; c < 3; ++c) {
    SlotDescriptor d;
    d.path.push_back(c);
    node.state.tbl.slots.push_back(d);
  }
  // three dictionary columns: shipdate (days 8000..10525), discount (0..10), quantity (1..50), one page each
  HdfsParquetScanner scanner;
  struct Col { vector<uint8_t> dict; vector<uint8_t> body; };
  Col cols[3];
  const int lo[3] = {8000, 0, 1}, n[3] = {2526, 11, 50};
  for (int c = 0; c < 3; ++c) {
    DictEncoder<int32_t> enc(4);
    vector<int> idx;
    for (int v = 0; v < n[c]; ++v) enc.Put(lo[c] + v, idx);
    cols[c].dict.resize(enc.dict_encoded_size());
    enc.WriteDict(&cols[c].dict[0]);
    cols[c].body.assign(1, 1);
    vector<HdfsParquetScanner::DataPage> pages;
    scanner.AddColumn<int32_t>(0, &cols[c].dict[0], cols[c].dict.size(), pages);
  }

  // conjunct 1: l_shipdate >= 8730 AND l_shipdate <= 9094       (BETWEEN arrives as And(Ge, Le))
  AndPredicate* between = new AndPredicate();
  between->AddChild(Compare("ge", TYPE_INT, 0, TYPE_INT, 8730));
  between->AddChild(Compare("le", TYPE_INT, 0, TYPE_INT, 9094));
  // conjunct 2: (l_discount IN (5, 6, 7, 99) OR cast(l_quantity) < 24)
  ScalarFnCall* in = new ScalarFnCall("in_set_lookup");
  in->AddChild(new SlotRef(ColumnType(TYPE_INT, 4), 1));
  const int in_vals[4] = {5, 6, 7, 99};
  for (int i = 0; i < 4; ++i) in->AddChild(new Literal(TYPE_INT, in_vals[i]));
  OrPredicate* either = new OrPredicate();
  either->AddChild(in);
  either->AddChild(Compare("lt", TYPE_INT, 2, TYPE_INT, 24, /*cast=*/true));
  // not expressible: column-vs-column, BOOLEAN literal, != -- the reference answers NULL and the scan falls back
  ScalarFnCall* colcol = new ScalarFnCall("lt");
  colcol->AddChild(new SlotRef(ColumnType(TYPE_INT, 4), 0));
  colcol->AddChild(new SlotRef(ColumnType(TYPE_INT, 4), 1));
  Expr* boolean_lit = Compare("eq", TYPE_BOOLEAN, 0, TYPE_BOOLEAN, 1);
  Expr* ne = Compare("ne", TYPE_INT, 0, TYPE_INT, 8730);

  SimplePredicate* p1 = between->CreateSimplePredicates(&node);
  SimplePredicate* p2 = either->CreateSimplePredicates(&node);
  CHECK_EQ(p1 != NULL, true);
  CHECK_EQ(p2 != NULL, true);
  CHECK_EQ(colcol->CreateSimplePredicates(&node) == NULL, true);
  CHECK_EQ(boolean_lit->CreateSimplePredicates(&node) == NULL, true);
  CHECK_EQ(ne->CreateSimplePredicates(&node) == NULL, true);
  // every literal type of the switch instantiates the drop-in's leaf templates
  const PrimitiveType lit_types[] = {TYPE_TINYINT, TYPE_SMALLINT, TYPE_INT, TYPE_BIGINT, TYPE_FLOAT, TYPE_DOUBLE, TYPE_STRING,
                                     TYPE_TIMESTAMP, TYPE_DECIMAL};
  for (size_t i = 0; i < sizeof(lit_types) / sizeof(lit_types[0]); ++i) {
    Expr* e = Compare("eq", lit_types[i] == TYPE_DECIMAL ? TYPE_DECIMAL : TYPE_INT, 0, lit_types[i], 3);
    CHECK_EQ(e->CreateSimplePredicates(&node) != NULL, true);
  }

  if (p1 && p2) {
    scanner.AddSimplePredicate(p1);
    scanner.AddSimplePredicate(p2);
    FusedProgram prog;
    CHECK_EQ(p1->Lower(&scanner, &prog), true);
    CHECK_EQ(p2->Lower(&scanner, &prog), true);
    // postfix: GE(col0, code 730) LE(col0, ...) AND IN(col1, {5,6,7}) LT(col2, code 23) OR
    CHECK_EQ(prog.nodes.size(), 6u);
    if (prog.nodes.size() == 6) {
      CHECK_EQ(prog.nodes[0].op, FLE_OP_GE); CHECK_EQ(prog.nodes[0].col, 0); CHECK_EQ(prog.nodes[0].code, 730u);
      CHECK_EQ(prog.nodes[1].col, 0); CHECK_EQ(prog.nodes[1].op == FLE_OP_LT || prog.nodes[1].op == FLE_OP_LE, true);
      CHECK_EQ(prog.nodes[1].op == FLE_OP_LT ? prog.nodes[1].code : prog.nodes[1].code + 1, 1095u);  // x <= 9094 as code < 1095 (dict-encoding.h:513-520)
      CHECK_EQ(prog.nodes[2].op, FLE_OP_AND);
      CHECK_EQ(prog.nodes[3].op, FLE_OP_IN); CHECK_EQ(prog.nodes[3].col, 1); CHECK_EQ(prog.nodes[3].ncodes, 3);  // 99 is not in the dictionary
      CHECK_EQ(prog.nodes[4].op, FLE_OP_LT); CHECK_EQ(prog.nodes[4].col, 2); CHECK_EQ(prog.nodes[4].code, 23u);
      CHECK_EQ(prog.nodes[5].op, FLE_OP_OR);
    }
  }
  std::printf("expr binding check: %d failure(s)\n", failures);
  return failures == 0 ? 0 : 1;
}
