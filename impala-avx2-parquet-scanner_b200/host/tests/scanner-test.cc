// End-to-end test of the column-reader batch interface (exec/hdfs-parquet-scanner.h):
// builds FLE_DICTIONARY column chunks the way the reference's table writer does
// (hdfs-parquet-table-writer.cc:378-432, :450-464, :547-575: one shared sorted dictionary,
// per-page definition levels + per-page index lists), attaches a SimplePredicate forest as
// ScalarFnCall/AndPredicate::CreateSimplePredicates would (scalar-fn-call.cc:938-965,
// compound-predicates.cc:36-67), and checks EvalSimplePredicates + AssembleRows against a
// row-at-a-time evaluation of the same predicate on the original values -- for the fused path
// and for the reference-shaped per-leaf path, which must agree bit for bit.
#include <stdint.h>
#include <stdlib.h>

#include <vector>

#include "exec/hdfs-parquet-scanner.h"
#include "exprs/simple-predicates.h"
#include "host-test-util.h"

using namespace impala;
using std::vector;

struct ColumnChunk {
  vector<uint8_t> dict_page;
  vector<vector<uint8_t> > page_bodies;
  vector<HdfsParquetScanner::DataPage> pages;
};

// values[i] valid iff notnull[i]
static ColumnChunk WriteColumn(const vector<int32_t>& values, const vector<bool>& notnull, int rows_per_page,
                               bool nullable) {
  ColumnChunk cc;
  DictEncoder<int32_t> dict(4);
  const int n = values.size();
  vector<vector<int> > page_indices;
  vector<vector<bool> > page_defs;
  for (int r = 0; r < n; ++r) {
    if (r % rows_per_page == 0) {
      page_indices.push_back(vector<int>());
      page_defs.push_back(vector<bool>());
    }
    const bool nn = !nullable || notnull[r];
    page_defs.back().push_back(nn);
    if (nn) EXPECT_TRUE(dict.Put(values[r], page_indices.back()) >= 0);
  }
  cc.dict_page.resize(dict.dict_encoded_size());
  dict.WriteDict(&cc.dict_page[0]);
  for (size_t p = 0; p < page_indices.size(); ++p) {
    vector<uint8_t> body;
    if (nullable) {
      vector<uint8_t> defbuf((page_defs[p].size() / 64 + 3) * 8);
      FleEncoder def_levels(&defbuf[0], defbuf.size(), 1);
      for (size_t i = 0; i < page_defs[p].size(); ++i) EXPECT_TRUE(def_levels.Put(page_defs[p][i]));
      def_levels.Flush();
      int32_t num_def_bytes = def_levels.len();
      body.resize(4 + num_def_bytes);
      memcpy(&body[0], &num_def_bytes, 4);
      memcpy(&body[4], def_levels.buffer(), num_def_bytes);
    }
    vector<uint8_t> data(1 + (page_indices[p].size() / 64 + 3) * dict.bit_width() * 8);
    int data_len = dict.WriteData(&data[0], data.size(), page_indices[p]);
    EXPECT_TRUE(data_len > 0);
    body.insert(body.end(), data.begin(), data.begin() + data_len);
    cc.page_bodies.push_back(body);
  }
  for (size_t p = 0; p < cc.page_bodies.size(); ++p) {
    cc.pages.push_back(HdfsParquetScanner::DataPage(&cc.page_bodies[p][0], cc.page_bodies[p].size(),
                                                    page_defs[p].size()));
  }
  return cc;
}

struct Table {
  int n;
  vector<int32_t> shipdate, discount, quantity;
  vector<bool> nn_shipdate, nn_discount, nn_quantity;
};

static Table MakeTable(int n, double null_frac, unsigned seed) {
  srand(seed);
  Table t;
  t.n = n;
  for (int i = 0; i < n; ++i) {
    t.shipdate.push_back(8000 + rand() % 2526);   // days
    t.discount.push_back(rand() % 11);            // cents
    t.quantity.push_back(1 + rand() % 50);
    t.nn_shipdate.push_back((rand() % 10000) >= null_frac * 10000);
    t.nn_discount.push_back((rand() % 10000) >= null_frac * 10000);
    t.nn_quantity.push_back((rand() % 10000) >= null_frac * 5000);
  }
  return t;
}

static bool RowPasses(const Table& t, int r) {
  // l_shipdate BETWEEN 8730 AND 9094 AND l_discount BETWEEN 5 AND 7 AND l_quantity < 24
  // (NULL fails every leaf, simple-predicates.h / hdfs-parquet-scanner.cc:342-345)
  return t.nn_shipdate[r] && t.shipdate[r] >= 8730 && t.shipdate[r] <= 9094 &&
         t.nn_discount[r] && t.discount[r] >= 5 && t.discount[r] <= 7 &&
         t.nn_quantity[r] && t.quantity[r] < 24;
}

static void RunScan(const Table& t, bool nullable, int pp0, int pp1, int pp2, bool fused, vector<uint8_t>* rows,
                    int64_t* passed, int64_t* fused_evals, int64_t* leaf_windows, int* tuple_size, int offs[3],
                    bool device = true, int64_t* device_spans = NULL) {
  ColumnChunk c0 = WriteColumn(t.shipdate, t.nn_shipdate, pp0, nullable);
  ColumnChunk c1 = WriteColumn(t.discount, t.nn_discount, pp1, nullable);
  ColumnChunk c2 = WriteColumn(t.quantity, t.nn_quantity, pp2, nullable);
  HdfsParquetScanner scanner;
  const int mdl = nullable ? 1 : 0;
  int i0 = scanner.AddColumn<int32_t>(mdl, &c0.dict_page[0], c0.dict_page.size(), c0.pages);
  int i1 = scanner.AddColumn<int32_t>(mdl, &c1.dict_page[0], c1.dict_page.size(), c1.pages);
  int i2 = scanner.AddColumn<int32_t>(mdl, &c2.dict_page[0], c2.dict_page.size(), c2.pages);
  // conjunct 1: And(Ge, Le) on shipdate -- BETWEEN arrives as And(Ge, Le)
  GeOperate<int32_t> ge0(i0, 8730);
  LeOperate<int32_t> le0(i0, 9094);
  AndOperate between0(&ge0, &le0);
  // conjuncts 2, 3: discount bounds as separate top-level conjuncts; 4: quantity
  GeOperate<int32_t> ge1(i1, 5);
  LeOperate<int32_t> le1(i1, 7);
  LtOperate<int32_t> lt2(i2, 24);
  scanner.AddSimplePredicate(&between0);
  scanner.AddSimplePredicate(&ge1);
  scanner.AddSimplePredicate(&le1);
  scanner.AddSimplePredicate(&lt2);
  scanner.set_fused(fused);
  scanner.set_device_materialise(device);
  Status st = scanner.AssembleRows(t.n, rows, passed);
  EXPECT_TRUE(st.ok());
  if (device_spans) *device_spans = scanner.device_spans();
  *fused_evals = scanner.fused_evals();
  *leaf_windows = scanner.leaf_windows();
  *tuple_size = scanner.tuple_size();
  offs[0] = scanner.slot_offset(i0); offs[1] = scanner.slot_offset(i1); offs[2] = scanner.slot_offset(i2);
}

static void CheckScan(const Table& t, bool nullable, int pp0, int pp1, int pp2, bool expect_fused) {
  Table tt = t;
  if (!nullable) {
    tt.nn_shipdate.assign(t.n, true); tt.nn_discount.assign(t.n, true); tt.nn_quantity.assign(t.n, true);
  }
  vector<uint8_t> rows_f, rows_l;
  int64_t passed_f = 0, passed_l = 0, fe = 0, lw = 0, fe2 = 0, lw2 = 0;
  int ts = 0, ts2 = 0, offs[3], offs2[3];
  int64_t spans = 0;
  RunScan(tt, nullable, pp0, pp1, pp2, true, &rows_f, &passed_f, &fe, &lw, &ts, offs, true, &spans);
  RunScan(tt, nullable, pp0, pp1, pp2, false, &rows_l, &passed_l, &fe2, &lw2, &ts2, offs2);
  EXPECT_EQ(fe2, 0);
  {  // fused evaluation with the survivors materialised on the host (ReadValue(skip) per survivor):
     // must give the same bytes as the device-resident page-span path
    vector<uint8_t> rows_h;
    int64_t passed_h = 0, fe3 = 0, lw3 = 0, spans_h = 0;
    int ts3 = 0, offs3[3];
    RunScan(tt, nullable, pp0, pp1, pp2, true, &rows_h, &passed_h, &fe3, &lw3, &ts3, offs3, false, &spans_h);
    EXPECT_EQ(spans_h, 0);
    EXPECT_EQ(passed_h, passed_f);
    EXPECT_TRUE(rows_h == rows_f);
  }
  if (expect_fused) { EXPECT_TRUE(fe > 0); EXPECT_EQ(lw, 0); }
  EXPECT_EQ(passed_f, passed_l);
  EXPECT_TRUE(rows_f == rows_l);  // fused and per-leaf paths agree byte for byte
  int64_t want = 0, bad = 0;
  for (int r = 0; r < tt.n; ++r) {
    if (!RowPasses(tt, r)) continue;
    if (want < passed_f) {
      const uint8_t* tuple = &rows_f[want * ts];
      int32_t a, b, c;
      memcpy(&a, tuple + offs[0], 4); memcpy(&b, tuple + offs[1], 4); memcpy(&c, tuple + offs[2], 4);
      bad += !(a == tt.shipdate[r] && b == tt.discount[r] && c == tt.quantity[r]);
      bad += tuple[0] || tuple[1] || tuple[2];  // survivors are non-NULL in all three columns
    }
    ++want;
  }
  EXPECT_EQ(passed_f, want);
  EXPECT_EQ(bad, 0);
  std::printf("    rows=%d nullable=%d pages=%d/%d/%d survivors=%lld fused_evals=%lld leaf_windows=%lld device_spans=%lld\n", tt.n,
              (int)nullable, pp0, pp1, pp2, (long long)passed_f, (long long)fe, (long long)lw, (long long)spans);
}

TEST(Scanner, EqualPagesDeviceSpans) {  // several whole-page spans, the last one ragged
  CheckScan(MakeTable(64 * 3000 + 17, 0, 11), false, 64 * 700, 64 * 700, 64 * 700, true);
  CheckScan(MakeTable(64 * 3000 + 17, 0, 12), true, 50001, 50001, 50001, true);
}
TEST(Scanner, RequiredColumnsOnePage) { CheckScan(MakeTable(64 * 3000 + 17, 0, 1), false, 1 << 30, 1 << 30, 1 << 30, true); }
TEST(Scanner, RequiredColumnsAlignedPages) { CheckScan(MakeTable(64 * 3000 + 17, 0, 2), false, 64 * 500, 64 * 1000, 64 * 250, true); }
TEST(Scanner, RequiredColumnsRaggedPages) { CheckScan(MakeTable(64 * 1000 + 5, 0, 3), false, 10007, 64 * 300, 33333, false); }
TEST(Scanner, NullableNoNulls) { CheckScan(MakeTable(64 * 2000 + 9, 0, 4), true, 50000, 50000, 50000, true); }
TEST(Scanner, NullableWithNulls) { CheckScan(MakeTable(64 * 2000 + 9, 0.03, 5), true, 50000, 50000, 50000, true); }
TEST(Scanner, NullableManyNullsRaggedPages) { CheckScan(MakeTable(64 * 700 + 1, 0.4, 6), true, 9999, 20000, 12345, false); }
TEST(Scanner, TinyRowGroups) {
  for (int n = 1; n <= 130; n += 43) CheckScan(MakeTable(n, 0.1, 7 + n), true, 1 << 30, 1 << 30, 1 << 30, true);
}

TEST(Scanner, OrAndInTree) {
  // (shipmode IN (..) OR instruct = x) AND quantity < q -- the "OR variant" of SURVEY 8d cfg3
  Table t = MakeTable(64 * 1500 + 3, 0.02, 99);
  vector<int32_t> shipmode, instruct;
  for (int i = 0; i < t.n; ++i) { shipmode.push_back(100 + rand() % 7); instruct.push_back(rand() % 4); }
  for (int fused = 0; fused < 2; ++fused) {
    ColumnChunk c0 = WriteColumn(shipmode, t.nn_shipdate, 40000, true);
    ColumnChunk c1 = WriteColumn(instruct, t.nn_discount, 40000, true);
    ColumnChunk c2 = WriteColumn(t.quantity, t.nn_quantity, 40000, true);
    HdfsParquetScanner scanner;
    int i0 = scanner.AddColumn<int32_t>(1, &c0.dict_page[0], c0.dict_page.size(), c0.pages);
    int i1 = scanner.AddColumn<int32_t>(1, &c1.dict_page[0], c1.dict_page.size(), c1.pages);
    int i2 = scanner.AddColumn<int32_t>(1, &c2.dict_page[0], c2.dict_page.size(), c2.pages);
    vector<int32_t> modes;
    modes.push_back(101); modes.push_back(104); modes.push_back(999);  // 999 is not in the dictionary
    InOperate<int32_t> in0(i0, modes);
    EqOperate<int32_t> eq1(i1, 2);
    OrOperate either(&in0, &eq1);
    LtOperate<int32_t> lt2(i2, 24);
    GtOperate<int32_t> gt_all(i2, -5);      // dictionary short-circuit: all non-NULL rows
    AndOperate root(&either, &lt2);
    scanner.AddSimplePredicate(&root);
    scanner.AddSimplePredicate(&gt_all);
    scanner.set_fused(fused != 0);
    vector<uint8_t> rows;
    int64_t passed = 0;
    EXPECT_TRUE(scanner.AssembleRows(t.n, &rows, &passed).ok());
    int64_t want = 0;
    for (int r = 0; r < t.n; ++r) {
      bool a = t.nn_shipdate[r] && (shipmode[r] == 101 || shipmode[r] == 104);
      bool b = t.nn_discount[r] && instruct[r] == 2;
      bool c = t.nn_quantity[r] && t.quantity[r] < 24;
      want += (a || b) && c;
    }
    EXPECT_EQ(passed, want);
    if (fused) EXPECT_TRUE(scanner.fused_evals() > 0);
  }
}

int main() { return RunAllTests(); }
