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
                    bool device = true, int64_t* device_spans = NULL, bool batch = true, int64_t* batches = NULL) {
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
  scanner.set_row_group_batch(batch);
  Status st = scanner.AssembleRows(t.n, rows, passed);
  EXPECT_TRUE(st.ok());
  if (device_spans) *device_spans = scanner.device_spans();
  if (batches) *batches = scanner.row_group_batches();
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
  {  // ... and as the page-span path with the whole-row-group batch switched off
    vector<uint8_t> rows_s;
    int64_t passed_s = 0, fe4 = 0, lw4 = 0, spans_s = 0, batches_s = 0;
    int ts4 = 0, offs4[3];
    RunScan(tt, nullable, pp0, pp1, pp2, true, &rows_s, &passed_s, &fe4, &lw4, &ts4, offs4, true, &spans_s, false, &batches_s);
    EXPECT_EQ(batches_s, 0);
    EXPECT_EQ(passed_s, passed_f);
    EXPECT_TRUE(rows_s == rows_f);
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

// the whole row group as ONE batch: equal pages across the columns, no NULLs (def levels present or not)
TEST(Scanner, RowGroupBatch) {
  for (int nullable = 0; nullable < 2; ++nullable) {
    Table t = MakeTable(64 * 5000 + 29, 0, 41 + nullable);
    vector<uint8_t> rows;
    int64_t passed = 0, fe = 0, lw = 0, spans = 0, batches = 0;
    int ts = 0, offs[3];
    RunScan(t, nullable != 0, 64 * 300, 64 * 300, 64 * 300, true, &rows, &passed, &fe, &lw, &ts, offs, true, &spans, true, &batches);
    EXPECT_EQ(batches, 1);
    EXPECT_EQ(spans, 0);
    EXPECT_EQ(lw, 0);
    int64_t want = 0;
    for (int r = 0; r < t.n; ++r) want += RowPasses(t, r);
    EXPECT_EQ(passed, want);
  }
  // NULLs present: the pages' values are expanded into row-aligned planes, the definition levels become 1-bit columns
  // ANDed to the leaves, the row group stays ONE batch; tuples identical to the page paths and to the truth
  for (int k = 0; k < 3; ++k) {
    const double fracs[3] = {0.01, 0.4, 0.97};
    Table t = MakeTable(64 * 5000 + 29, fracs[k], 43 + k);
    vector<uint8_t> rows, rows_p;
    int64_t passed = 0, passed_p = 0, fe = 0, lw = 0, spans = 0, batches = 0, b2 = 0;
    int ts = 0, offs[3];
    RunScan(t, true, 64 * 300, 64 * 300, 64 * 300, true, &rows, &passed, &fe, &lw, &ts, offs, true, &spans, true, &batches);
    EXPECT_EQ(batches, 1);
    EXPECT_EQ(spans, 0);
    EXPECT_EQ(lw, 0);
    RunScan(t, true, 64 * 300, 64 * 300, 64 * 300, true, &rows_p, &passed_p, &fe, &lw, &ts, offs, true, &spans, false, &b2);
    EXPECT_EQ(b2, 0);
    EXPECT_EQ(passed, passed_p);
    EXPECT_TRUE(rows == rows_p);
    int64_t want = 0, bad = 0;
    for (int r = 0; r < t.n; ++r) {
      if (!RowPasses(t, r)) continue;
      if (want < passed) {
        const uint8_t* tuple = &rows[want * ts];
        int32_t a, b, c;
        memcpy(&a, tuple + offs[0], 4); memcpy(&b, tuple + offs[1], 4); memcpy(&c, tuple + offs[2], 4);
        bad += !(a == t.shipdate[r] && b == t.discount[r] && c == t.quantity[r]) || tuple[0] || tuple[1] || tuple[2];
      }
      ++want;
    }
    EXPECT_EQ(passed, want);
    EXPECT_EQ(bad, 0);
  }
}

// an OR tree over nullable columns as one batch: survivors that are NULL in a materialised column carry the column's
// NULL indicator and a zero slot, exactly as the page paths (ReadValue -> ReadSlot) write them
TEST(Scanner, RowGroupBatchOrTreeNullIndicators) {
  Table t = MakeTable(64 * 1500, 0.2, 77);
  vector<int32_t> shipmode, instruct;
  for (int i = 0; i < t.n; ++i) { shipmode.push_back(100 + rand() % 7); instruct.push_back(rand() % 4); }
  vector<uint8_t> rows[2];
  int64_t passed[2] = {0, 0};
  for (int batch = 0; batch < 2; ++batch) {
    ColumnChunk c0 = WriteColumn(shipmode, t.nn_shipdate, 64 * 500, true);
    ColumnChunk c1 = WriteColumn(instruct, t.nn_discount, 64 * 500, true);
    ColumnChunk c2 = WriteColumn(t.quantity, t.nn_quantity, 64 * 500, true);
    HdfsParquetScanner scanner;
    int i0 = scanner.AddColumn<int32_t>(1, &c0.dict_page[0], c0.dict_page.size(), c0.pages);
    int i1 = scanner.AddColumn<int32_t>(1, &c1.dict_page[0], c1.dict_page.size(), c1.pages);
    int i2 = scanner.AddColumn<int32_t>(1, &c2.dict_page[0], c2.dict_page.size(), c2.pages);
    vector<int32_t> modes;
    modes.push_back(101); modes.push_back(104); modes.push_back(999);
    InOperate<int32_t> in0(i0, modes);
    EqOperate<int32_t> eq1(i1, 2);
    OrOperate either(&in0, &eq1);
    GtOperate<int32_t> gt_all(i2, -5);  // dictionary short-circuit: all non-NULL rows
    OrOperate root(&either, &gt_all);
    scanner.AddSimplePredicate(&root);
    scanner.set_fused(true);
    scanner.set_row_group_batch(batch != 0);
    EXPECT_TRUE(scanner.AssembleRows(t.n, &rows[batch], &passed[batch]).ok());
    EXPECT_EQ(scanner.row_group_batches(), batch);
    EXPECT_EQ(scanner.row_group_null_batches(), batch);
    if (batch) {
      const int ts = scanner.tuple_size();
      int64_t want = 0, bad = 0;
      for (int r = 0; r < t.n; ++r) {
        const bool a = t.nn_shipdate[r] && (shipmode[r] == 101 || shipmode[r] == 104);
        const bool b = t.nn_discount[r] && instruct[r] == 2;
        const bool c = t.nn_quantity[r];
        if (!(a || b || c)) continue;
        if (want < passed[1]) {
          const uint8_t* tuple = &rows[1][want * ts];
          int32_t v0, v1, v2;
          memcpy(&v0, tuple + scanner.slot_offset(i0), 4); memcpy(&v1, tuple + scanner.slot_offset(i1), 4);
          memcpy(&v2, tuple + scanner.slot_offset(i2), 4);
          bad += tuple[0] != !t.nn_shipdate[r] || tuple[1] != !t.nn_discount[r] || tuple[2] != !t.nn_quantity[r];
          bad += (t.nn_shipdate[r] ? v0 != shipmode[r] : v0 != 0) || (t.nn_discount[r] ? v1 != instruct[r] : v1 != 0) ||
                 (t.nn_quantity[r] ? v2 != t.quantity[r] : v2 != 0);
        }
        ++want;
      }
      EXPECT_EQ(passed[1], want);
      EXPECT_EQ(bad, 0);
    }
  }
  EXPECT_EQ(passed[0], passed[1]);
  EXPECT_TRUE(rows[0] == rows[1]);
}

// PLAIN data pages (a column that overflowed the dictionary, hdfs-parquet-table-writer.cc:276-279) next to a
// dictionary column: ParquetPlainEncoder::Eq..In (parquet-common.h:197-255)
TEST(Scanner, PlainPages) {
  const int n = 64 * 900 + 11, per_page = 20000;
  srand(77);
  vector<int32_t> price(n), qty(n);
  vector<bool> nn(n, true);
  for (int i = 0; i < n; ++i) { price[i] = rand() % 1000000; qty[i] = 1 + rand() % 50; }
  for (int compat = 0; compat < 2; ++compat) {
    FleGpu::SetReferenceCompat(compat != 0);
    ColumnChunk cq = WriteColumn(qty, nn, per_page, true);
    ColumnChunk cp;  // PLAIN: [int32 num_def_bytes][def levels][values back to back], no dictionary page
    for (int r0 = 0; r0 < n; r0 += per_page) {
      const int m = std::min(per_page, n - r0);
      vector<uint8_t> defbuf((m / 64 + 3) * 8);
      FleEncoder def_levels(&defbuf[0], defbuf.size(), 1);
      for (int i = 0; i < m; ++i) def_levels.Put(1);
      def_levels.Flush();
      int32_t nd = def_levels.len();
      vector<uint8_t> body(4 + nd);
      memcpy(&body[0], &nd, 4);
      memcpy(&body[4], def_levels.buffer(), nd);
      body.insert(body.end(), reinterpret_cast<const uint8_t*>(&price[r0]), reinterpret_cast<const uint8_t*>(&price[r0]) + 4 * m);
      cp.page_bodies.push_back(body);
    }
    for (size_t p = 0; p < cp.page_bodies.size(); ++p) {
      HdfsParquetScanner::DataPage pg(&cp.page_bodies[p][0], cp.page_bodies[p].size(),
                                      std::min(per_page, n - (int)p * per_page));
      pg.encoding = parquet_encoding::PLAIN;
      cp.pages.push_back(pg);
    }
    HdfsParquetScanner scanner;
    int ip = scanner.AddColumn<int32_t>(1, NULL, 0, cp.pages);
    int iq = scanner.AddColumn<int32_t>(1, &cq.dict_page[0], cq.dict_page.size(), cq.pages);
    LtOperate<int32_t> lt_price(ip, 30000);
    LtOperate<int32_t> lt_qty(iq, 24);
    scanner.AddSimplePredicate(&lt_price);
    scanner.AddSimplePredicate(&lt_qty);
    vector<uint8_t> rows;
    int64_t passed = 0;
    Status st = scanner.AssembleRows(n, &rows, &passed);
    EXPECT_TRUE(st.ok());
    // default: `price < 30000`; reference_compat: the reference's generic template evaluates `30000 < price`
    // (parquet-common.h:214)
    int64_t want = 0, bad = 0;
    for (int r = 0; r < n; ++r) {
      const bool pass = (compat ? 30000 < price[r] : price[r] < 30000) && qty[r] < 24;
      if (!pass) continue;
      if (want < passed) {
        const uint8_t* tuple = &rows[want * scanner.tuple_size()];
        int32_t a, b;
        memcpy(&a, tuple + scanner.slot_offset(ip), 4);
        memcpy(&b, tuple + scanner.slot_offset(iq), 4);
        bad += !(a == price[r] && b == qty[r] && tuple[0] == 0 && tuple[1] == 0);
      }
      ++want;
    }
    EXPECT_EQ(passed, want);
    EXPECT_EQ(bad, 0);
    std::printf("    compat=%d plain pages: %lld survivors\n", compat, (long long)passed);
  }
  FleGpu::SetReferenceCompat(false);
}

// A PLAIN column that is only materialised (l_extendedprice next to the Q6 predicate columns: too many distinct values
// for a dictionary) rides along in the row-group batch: its values are placed into a row-aligned value array, NULLs
// included, and the survivors' slots are copied from there.  Ragged pages, cut differently per column.
static ColumnChunk WritePlainColumn(const vector<int64_t>& values, const vector<bool>& notnull, int rows_per_page) {
  ColumnChunk cp;
  const int n = values.size();
  for (int r0 = 0; r0 < n; r0 += rows_per_page) {
    const int m = std::min(rows_per_page, n - r0);
    vector<uint8_t> defbuf((m / 64 + 3) * 8);
    FleEncoder def_levels(&defbuf[0], defbuf.size(), 1);
    for (int i = 0; i < m; ++i) def_levels.Put(notnull[r0 + i]);
    def_levels.Flush();
    int32_t nd = def_levels.len();
    vector<uint8_t> body(4 + nd);
    memcpy(&body[0], &nd, 4);
    memcpy(&body[4], def_levels.buffer(), nd);
    for (int i = 0; i < m; ++i) {
      if (!notnull[r0 + i]) continue;
      const uint8_t* v = reinterpret_cast<const uint8_t*>(&values[r0 + i]);
      body.insert(body.end(), v, v + 8);
    }
    cp.page_bodies.push_back(body);
  }
  for (size_t p = 0; p < cp.page_bodies.size(); ++p) {
    HdfsParquetScanner::DataPage pg(&cp.page_bodies[p][0], cp.page_bodies[p].size(), std::min(rows_per_page, n - (int)p * rows_per_page));
    pg.encoding = parquet_encoding::PLAIN;
    cp.pages.push_back(pg);
  }
  return cp;
}

// the chunk as the writer leaves it: dictionary pages until the dictionary is full, PLAIN pages from there on
static ColumnChunk WriteMixedColumn(const vector<int64_t>& values, const vector<bool>& notnull, int rows_per_page, int dict_pages) {
  ColumnChunk cc;
  DictEncoder<int64_t> dict(8);
  const int n = values.size();
  vector<vector<int> > page_indices;
  for (int p = 0; p < dict_pages && p * rows_per_page < n; ++p) {
    page_indices.push_back(vector<int>());
    for (int r = p * rows_per_page; r < std::min(n, (p + 1) * rows_per_page); ++r)
      if (notnull[r]) EXPECT_TRUE(dict.Put(values[r], page_indices.back()) >= 0);
  }
  cc.dict_page.resize(dict.dict_encoded_size());
  dict.WriteDict(&cc.dict_page[0]);
  for (int p = 0; p * rows_per_page < n; ++p) {
    const int r0 = p * rows_per_page, m = std::min(rows_per_page, n - r0);
    vector<uint8_t> defbuf((m / 64 + 3) * 8);
    FleEncoder def_levels(&defbuf[0], defbuf.size(), 1);
    for (int i = 0; i < m; ++i) def_levels.Put(notnull[r0 + i]);
    def_levels.Flush();
    int32_t nd = def_levels.len();
    vector<uint8_t> body(4 + nd);
    memcpy(&body[0], &nd, 4);
    memcpy(&body[4], def_levels.buffer(), nd);
    if (p < dict_pages) {
      vector<uint8_t> data(1 + (page_indices[p].size() / 64 + 3) * dict.bit_width() * 8);
      const int data_len = dict.WriteData(&data[0], data.size(), page_indices[p]);
      EXPECT_TRUE(data_len > 0);
      body.insert(body.end(), data.begin(), data.begin() + data_len);
    } else {
      for (int i = 0; i < m; ++i) {
        if (!notnull[r0 + i]) continue;
        const uint8_t* v = reinterpret_cast<const uint8_t*>(&values[r0 + i]);
        body.insert(body.end(), v, v + 8);
      }
    }
    cc.page_bodies.push_back(body);
  }
  for (size_t p = 0; p < cc.page_bodies.size(); ++p) {
    HdfsParquetScanner::DataPage pg(&cc.page_bodies[p][0], cc.page_bodies[p].size(), std::min(rows_per_page, n - (int)p * rows_per_page));
    if ((int)p >= dict_pages) pg.encoding = parquet_encoding::PLAIN;
    cc.pages.push_back(pg);
  }
  return cc;
}

TEST(Scanner, RowGroupBatchPlainMaterialisedColumn) {
  for (int k = 0; k < 4; ++k) {  // k & 1: NULLs; k & 2: the chunk starts with dictionary pages
    Table t = MakeTable(64 * 3000 + 41, (k & 1) ? 0.05 : 0.0, 61 + k);
    vector<int64_t> price(t.n);
    vector<bool> nn_price(t.n);
    for (int i = 0; i < t.n; ++i) { price[i] = 90000 + (int64_t)rand() * 5 % 10400000; nn_price[i] = (k & 1) == 0 || rand() % 10 != 0; }
    vector<uint8_t> rows[2];
    int64_t passed[2] = {0, 0};
    int ts = 0, off_p = 0, off_d = 0, ip = 0;
    for (int batch = 1; batch >= 0; --batch) {
      ColumnChunk c0 = WriteColumn(t.shipdate, t.nn_shipdate, 50001, true);
      ColumnChunk c1 = WriteColumn(t.discount, t.nn_discount, 64 * 700, true);
      ColumnChunk c2 = WriteColumn(t.quantity, t.nn_quantity, 33333, true);
      ColumnChunk cp = (k & 2) ? WriteMixedColumn(price, nn_price, 20011, 2) : WritePlainColumn(price, nn_price, 20011);
      HdfsParquetScanner scanner;
      int i0 = scanner.AddColumn<int32_t>(1, &c0.dict_page[0], c0.dict_page.size(), c0.pages);
      int i1 = scanner.AddColumn<int32_t>(1, &c1.dict_page[0], c1.dict_page.size(), c1.pages);
      ip = (k & 2) ? scanner.AddColumn<int64_t>(1, &cp.dict_page[0], cp.dict_page.size(), cp.pages) : scanner.AddColumn<int64_t>(1, NULL, 0, cp.pages);
      int i2 = scanner.AddColumn<int32_t>(1, &c2.dict_page[0], c2.dict_page.size(), c2.pages);
      GeOperate<int32_t> ge0(i0, 8730);
      LeOperate<int32_t> le0(i0, 9094);
      AndOperate between0(&ge0, &le0);
      GeOperate<int32_t> ge1(i1, 5);
      LeOperate<int32_t> le1(i1, 7);
      LtOperate<int32_t> lt2(i2, 24);
      scanner.AddSimplePredicate(&between0);
      scanner.AddSimplePredicate(&ge1);
      scanner.AddSimplePredicate(&le1);
      scanner.AddSimplePredicate(&lt2);
      scanner.set_row_group_batch(batch != 0);
      EXPECT_TRUE(scanner.AssembleRows(t.n, &rows[batch], &passed[batch]).ok());
      EXPECT_EQ(scanner.row_group_batches(), batch);
      ts = scanner.tuple_size();
      off_p = scanner.slot_offset(ip);
      off_d = scanner.slot_offset(i1);
    }
    EXPECT_EQ(passed[0], passed[1]);
    EXPECT_TRUE(rows[0] == rows[1]);
    int64_t want = 0, bad = 0;
    for (int r = 0; r < t.n; ++r) {
      if (!RowPasses(t, r)) continue;
      if (want < passed[1]) {
        const uint8_t* tuple = &rows[1][want * ts];
        int64_t pv;
        int32_t dv;
        memcpy(&pv, tuple + off_p, 8);
        memcpy(&dv, tuple + off_d, 4);
        bad += tuple[ip] != !nn_price[r] || (nn_price[r] ? pv != price[r] : pv != 0) || dv != t.discount[r];
      }
      ++want;
    }
    EXPECT_EQ(passed[1], want);
    EXPECT_EQ(bad, 0);
    std::printf("    plain materialised column, nulls=%d dictionary pages first=%d: %lld survivors\n", k & 1, (k >> 1) & 1, (long long)passed[1]);
  }
}

// Differential fuzz: random row counts, page cuts per column, nullability, NULL fractions and predicate shapes; the
// row-group batch must return the page paths' tuples byte for byte (those are checked against the truth above).
TEST(Scanner, RowGroupBatchFuzz) {
  srand(2024);
  int batches = 0;
  for (int it = 0; it < 24; ++it) {
    const int n = 1 + rand() % 90000;
    const double nf = (it % 4 == 0) ? 0.0 : (rand() % 1000) / 1000.0 * ((it % 3) ? 0.1 : 1.0);
    Table t = MakeTable(n, nf, 1000 + it);
    bool nullable[3];
    int pp[3];
    for (int c = 0; c < 3; ++c) {
      nullable[c] = rand() % 3 != 0;
      pp[c] = (rand() % 4 == 0) ? (1 << 30) : 1 + rand() % (n + 10);
      if (pp[c] < n / 40) pp[c] = n / 40 + 1;  // (at most ~40 pages per column)
    }
    const int shape = rand() % 3;
    vector<uint8_t> rows[2];
    int64_t passed[2] = {0, 0};
    for (int batch = 1; batch >= 0; --batch) {
      ColumnChunk c0 = WriteColumn(t.shipdate, t.nn_shipdate, pp[0], nullable[0]);
      ColumnChunk c1 = WriteColumn(t.discount, t.nn_discount, pp[1], nullable[1]);
      ColumnChunk c2 = WriteColumn(t.quantity, t.nn_quantity, pp[2], nullable[2]);
      HdfsParquetScanner scanner;
      int i0 = scanner.AddColumn<int32_t>(nullable[0], &c0.dict_page[0], c0.dict_page.size(), c0.pages);
      int i1 = scanner.AddColumn<int32_t>(nullable[1], &c1.dict_page[0], c1.dict_page.size(), c1.pages);
      int i2 = scanner.AddColumn<int32_t>(nullable[2], &c2.dict_page[0], c2.dict_page.size(), c2.pages);
      GeOperate<int32_t> ge0(i0, 8730);
      LeOperate<int32_t> le0(i0, 9094);
      AndOperate between0(&ge0, &le0);
      GeOperate<int32_t> ge1(i1, 5);
      LtOperate<int32_t> lt2(i2, 24);
      vector<int32_t> qs;
      qs.push_back(3); qs.push_back(17); qs.push_back(49); qs.push_back(1000);
      InOperate<int32_t> in2(i2, qs);
      OrOperate or01(&between0, &ge1);
      OrOperate or012(&or01, &in2);
      if (shape == 0) {  // conjunct chain
        scanner.AddSimplePredicate(&between0);
        scanner.AddSimplePredicate(&ge1);
        scanner.AddSimplePredicate(&lt2);
      } else if (shape == 1) {  // (a OR b) AND c
        scanner.AddSimplePredicate(&or01);
        scanner.AddSimplePredicate(&lt2);
      } else {  // a OR b OR c: survivors with NULLs in other columns
        scanner.AddSimplePredicate(&or012);
      }
      scanner.set_row_group_batch(batch != 0);
      EXPECT_TRUE(scanner.AssembleRows(t.n, &rows[batch], &passed[batch]).ok());
      if (batch) batches += scanner.row_group_batches();
    }
    EXPECT_EQ(passed[0], passed[1]);
    EXPECT_TRUE(rows[0] == rows[1]);
    if (rows[0] != rows[1])
      std::printf("    fuzz case %d differs: n=%d null_frac=%.3f pages=%d/%d/%d nullable=%d%d%d shape=%d\n", it, n, nf, pp[0], pp[1], pp[2],
                  (int)nullable[0], (int)nullable[1], (int)nullable[2], shape);
  }
  EXPECT_EQ(batches, 24);
}

// Four scanner threads at once (one library context each, hdfs-parquet-scanner.cc:1040), every one scanning its own
// tables with NULLs, ragged pages and a PLAIN column as one batch per row group; results against the truth.
#include <atomic>
#include <thread>
TEST(Scanner, ConcurrentScannersWithNullsAndRaggedPages) {
  std::atomic<int> bad(0), scans(0);
  vector<std::thread> th;
  for (int tix = 0; tix < 4; ++tix) {
    th.emplace_back([&, tix]() {
      for (int it = 0; it < 6; ++it) {
        Table t = MakeTable(20000 + 7919 * ((tix * 6 + it) % 11), (it % 3) * 0.04, 500 + tix * 17 + it);
        vector<int64_t> price(t.n);
        vector<bool> nn_price(t.n);
        unsigned st = 99 + tix * 31 + it;
        for (int i = 0; i < t.n; ++i) {
          st = st * 1664525u + 1013904223u;
          price[i] = 90000 + (st >> 8) % 10400000;
          nn_price[i] = (it % 2 == 0) || ((st >> 3) % 9 != 0);
        }
        ColumnChunk c0 = WriteColumn(t.shipdate, t.nn_shipdate, 5000 + 777 * tix, true);
        ColumnChunk c1 = WriteColumn(t.discount, t.nn_discount, 64 * 100, true);
        ColumnChunk c2 = WriteColumn(t.quantity, t.nn_quantity, 1 << 30, true);
        ColumnChunk cp = WriteMixedColumn(price, nn_price, 3001, 2);
        HdfsParquetScanner scanner;
        int i0 = scanner.AddColumn<int32_t>(1, &c0.dict_page[0], c0.dict_page.size(), c0.pages);
        int i1 = scanner.AddColumn<int32_t>(1, &c1.dict_page[0], c1.dict_page.size(), c1.pages);
        int ip = scanner.AddColumn<int64_t>(1, &cp.dict_page[0], cp.dict_page.size(), cp.pages);
        int i2 = scanner.AddColumn<int32_t>(1, &c2.dict_page[0], c2.dict_page.size(), c2.pages);
        GeOperate<int32_t> ge0(i0, 8730);
        LeOperate<int32_t> le0(i0, 9094);
        AndOperate between0(&ge0, &le0);
        GeOperate<int32_t> ge1(i1, 5);
        LeOperate<int32_t> le1(i1, 7);
        LtOperate<int32_t> lt2(i2, 24);
        scanner.AddSimplePredicate(&between0);
        scanner.AddSimplePredicate(&ge1);
        scanner.AddSimplePredicate(&le1);
        scanner.AddSimplePredicate(&lt2);
        vector<uint8_t> rows;
        int64_t passed = 0;
        if (!scanner.AssembleRows(t.n, &rows, &passed).ok() || scanner.row_group_batches() != 1) { ++bad; continue; }
        const int ts = scanner.tuple_size(), off_p = scanner.slot_offset(ip), off_s = scanner.slot_offset(i0);
        int64_t want = 0;
        for (int r = 0; r < t.n; ++r) {
          if (!RowPasses(t, r)) continue;
          if (want < passed) {
            const uint8_t* tuple = &rows[want * ts];
            int64_t pv;
            int32_t sv;
            memcpy(&pv, tuple + off_p, 8);
            memcpy(&sv, tuple + off_s, 4);
            if (tuple[ip] != !nn_price[r] || (nn_price[r] ? pv != price[r] : pv != 0) || sv != t.shipdate[r]) ++bad;
          }
          ++want;
        }
        if (want != passed) ++bad;
        ++scans;
      }
    });
  }
  for (auto& x : th) x.join();
  EXPECT_EQ(bad.load(), 0);
  EXPECT_EQ(scans.load(), 24);
}

// AssembleRowsInto: the same tuples, delivered into the caller's tuple buffer (batch path: device -> buffer directly;
// page paths: through the vector)
TEST(Scanner, AssembleRowsIntoCallerBuffer) {
  for (int batch = 0; batch < 2; ++batch) {
    Table t = MakeTable(64 * 900 + 5, 0.03, 321 + batch);
    vector<uint8_t> rows;
    vector<uint8_t> buf(size_t(t.n) * 32);
    int64_t passed = 0, passed2 = 0;
    size_t bytes = 0;
    for (int into = 0; into < 2; ++into) {
      ColumnChunk c0 = WriteColumn(t.shipdate, t.nn_shipdate, 20001, true);
      ColumnChunk c1 = WriteColumn(t.discount, t.nn_discount, 64 * 50, true);
      ColumnChunk c2 = WriteColumn(t.quantity, t.nn_quantity, 1 << 30, true);
      HdfsParquetScanner scanner;
      int i0 = scanner.AddColumn<int32_t>(1, &c0.dict_page[0], c0.dict_page.size(), c0.pages);
      int i1 = scanner.AddColumn<int32_t>(1, &c1.dict_page[0], c1.dict_page.size(), c1.pages);
      int i2 = scanner.AddColumn<int32_t>(1, &c2.dict_page[0], c2.dict_page.size(), c2.pages);
      GeOperate<int32_t> ge0(i0, 8730);
      LeOperate<int32_t> le1(i1, 7);
      LtOperate<int32_t> lt2(i2, 24);
      scanner.AddSimplePredicate(&ge0);
      scanner.AddSimplePredicate(&le1);
      scanner.AddSimplePredicate(&lt2);
      scanner.set_row_group_batch(batch != 0);
      if (into) {
        EXPECT_TRUE(scanner.AssembleRowsInto(t.n, &buf[0], buf.size(), &passed2, &bytes).ok());
        EXPECT_EQ((int64_t)bytes, passed2 * scanner.tuple_size());
      } else {
        EXPECT_TRUE(scanner.AssembleRows(t.n, &rows, &passed).ok());
      }
      EXPECT_EQ(scanner.row_group_batches(), batch);
    }
    EXPECT_EQ((int64_t)bytes, (int64_t)rows.size());
    EXPECT_TRUE(bytes == rows.size() && memcmp(&buf[0], &rows[0], bytes) == 0);
  }
}

// The deviations DESIGN.md lists, pinned in both modes (the reference lines are cited in util/fle-gpu-context.h)
TEST(Scanner, ReferenceCompatSwitch) {
  // 1. Skip() landing exactly on the end of the data (fle-encoding.h:381-402)
  const int bw = 7, n = 64 * 5;
  vector<uint8_t> buf((n / 64 + 2) * bw * 8);
  FleEncoder enc(&buf[0], buf.size(), bw);
  for (int i = 0; i < n; ++i) EXPECT_TRUE(enc.Put(i % 100));
  enc.Flush();
  for (int compat = 0; compat < 2; ++compat) {
    FleGpu::SetReferenceCompat(compat != 0);
    FleDecoder dec(&buf[0], (n / 64) * bw * 8, bw);
    EXPECT_TRUE(dec.Skip(n - 64));
    EXPECT_EQ((int)dec.Skip(64), compat ? 0 : 1);
  }
  // 2. a row group whose pages hold a multiple of 64 values: the reference's AssembleRows stops at the first page
  //    whose trailing Skip lands on the page end (hdfs-parquet-scanner.cc:1107-1113) and reports success
  Table t = MakeTable(64 * 64, 0, 5);
  t.quantity.assign(t.n, 1);  // every row passes `quantity < 24` ...
  for (int r = 0; r < t.n; ++r) if (r % 1024 >= 1000) t.quantity[r] = 40;  // ... except the tail of every window
  vector<bool> all(t.n, true);
  for (int compat = 0; compat < 2; ++compat) {
    FleGpu::SetReferenceCompat(compat != 0);
    ColumnChunk c = WriteColumn(t.quantity, all, 64 * 32, true);  // two pages of 2048 rows
    HdfsParquetScanner scanner;
    int i0 = scanner.AddColumn<int32_t>(1, &c.dict_page[0], c.dict_page.size(), c.pages);
    LtOperate<int32_t> lt(i0, 24);
    scanner.AddSimplePredicate(&lt);
    vector<uint8_t> rows;
    int64_t passed = 0;
    EXPECT_TRUE(scanner.AssembleRows(t.n, &rows, &passed).ok());
    int64_t want_all = 0, want_first_page = 0;
    for (int r = 0; r < t.n; ++r) { want_all += t.quantity[r] < 24; if (r < 64 * 32) want_first_page += t.quantity[r] < 24; }
    EXPECT_EQ(passed, compat ? want_first_page : want_all);
  }
  // 3. the data page's width byte (dict-encoding.h:438-444): Log2(max index + 1) vs the width the data is packed with
  for (int compat = 0; compat < 2; ++compat) {
    FleGpu::SetReferenceCompat(compat != 0);
    DictEncoder<int32_t> dict(4);
    vector<int> first_page, second_page;
    for (int v = 0; v < 40; ++v) dict.Put(1000 - v, v < 3 ? first_page : second_page);  // page 1 sees 3 of 40 entries
    vector<uint8_t> dpage(dict.dict_encoded_size());
    dict.WriteDict(&dpage[0]);
    vector<uint8_t> data(1 + 3 * dict.bit_width() * 8);
    EXPECT_TRUE(dict.WriteData(&data[0], data.size(), first_page) > 0);
    EXPECT_EQ((int)data[0], compat ? 2 : dict.bit_width());  // Log2(2 + 1) = 2; the data is packed with 6 bits either way
  }
  FleGpu::SetReferenceCompat(false);
}

int main() { return RunAllTests(); }
