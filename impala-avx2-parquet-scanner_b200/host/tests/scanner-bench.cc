// Throughput of the drop-in column-reader batch interface (exec/hdfs-parquet-scanner.h) on one
// Impala-shaped row group: three OPTIONAL FLE_DICTIONARY columns (def levels present, no NULLs),
// host-resident page bodies, the TPC-H Q6 conjuncts attached as a SimplePredicate forest,
// AssembleRows = EvalSimplePredicates window by window + materialisation of the survivors
// (hdfs-parquet-scanner.cc:1101-1181).  Reports rows/s for the fused path (one GPU evaluation per
// page span; survivors materialised on the device, or by ReadValue(skip) on the host from windows
// sliced out of the cached bitmap) and for the reference-shaped per-leaf path
// (one call per leaf and 1024-row window, still on the GPU), and checks both against a
// row-at-a-time evaluation.
//   scanner-bench [rows (default 4194304)] [rows per page (default 65536)] [repetitions (default 5)]
#include <stdint.h>
#include <stdlib.h>

#include <chrono>
#include <cstdio>
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

static ColumnChunk WriteColumn(const vector<int32_t>& values, int rows_per_page) {
  ColumnChunk cc;
  DictEncoder<int32_t> dict(4);
  const int n = values.size();
  vector<vector<int> > page_indices;
  for (int r = 0; r < n; ++r) {
    if (r % rows_per_page == 0) page_indices.push_back(vector<int>());
    if (dict.Put(values[r], page_indices.back()) < 0) { std::fprintf(stderr, "dictionary full\n"); exit(2); }
  }
  cc.dict_page.resize(dict.dict_encoded_size());
  dict.WriteDict(&cc.dict_page[0]);
  for (size_t p = 0; p < page_indices.size(); ++p) {
    const size_t rows = page_indices[p].size();
    vector<uint8_t> defbuf((rows / 64 + 3) * 8);
    FleEncoder def_levels(&defbuf[0], defbuf.size(), 1);
    for (size_t i = 0; i < rows; ++i) def_levels.Put(1);
    def_levels.Flush();
    int32_t num_def_bytes = def_levels.len();
    vector<uint8_t> body(4 + num_def_bytes);
    memcpy(&body[0], &num_def_bytes, 4);
    memcpy(&body[4], def_levels.buffer(), num_def_bytes);
    vector<uint8_t> data(1 + (rows / 64 + 3) * dict.bit_width() * 8);
    const int data_len = dict.WriteData(&data[0], data.size(), page_indices[p]);
    if (data_len <= 0) { std::fprintf(stderr, "WriteData failed\n"); exit(2); }
    body.insert(body.end(), data.begin(), data.begin() + data_len);
    cc.page_bodies.push_back(body);
  }
  size_t r0 = 0;
  for (size_t p = 0; p < cc.page_bodies.size(); ++p) {
    cc.pages.push_back(HdfsParquetScanner::DataPage(&cc.page_bodies[p][0], cc.page_bodies[p].size(), page_indices[p].size()));
    r0 += page_indices[p].size();
  }
  return cc;
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 4194304;
  const int rows_per_page = argc > 2 ? atoi(argv[2]) : 65536;
  const int reps = argc > 3 ? atoi(argv[3]) : 5;
  srand(7);
  vector<int32_t> shipdate(n), discount(n), quantity(n);
  int64_t want = 0;
  for (int i = 0; i < n; ++i) {
    shipdate[i] = 8000 + rand() % 2526;
    discount[i] = rand() % 11;
    quantity[i] = 1 + rand() % 50;
    want += shipdate[i] >= 8730 && shipdate[i] <= 9094 && discount[i] >= 5 && discount[i] <= 7 && quantity[i] < 24;
  }
  ColumnChunk c0 = WriteColumn(shipdate, rows_per_page), c1 = WriteColumn(discount, rows_per_page), c2 = WriteColumn(quantity, rows_per_page);
  size_t bytes = 0;
  for (const ColumnChunk* c : {&c0, &c1, &c2})
    for (const auto& b : c->page_bodies) bytes += b.size();
  std::printf("{\"rows\": %d, \"rows_per_page\": %d, \"page_bytes_total\": %zu, \"survivors\": %lld", n, rows_per_page, bytes, (long long)want);
  for (int mode = 2; mode >= 0; --mode) {  // 2: fused + device-resident page spans, 1: fused + host materialisation, 0: per leaf
    const int fused = mode >= 1;
    double best = 1e30;
    int64_t passed = 0, evals = 0, windows = 0;
    for (int r = 0; r < reps + 1; ++r) {  // first repetition = warm-up (context creation, kernel load)
      HdfsParquetScanner scanner;
      int i0 = scanner.AddColumn<int32_t>(1, &c0.dict_page[0], c0.dict_page.size(), c0.pages);
      int i1 = scanner.AddColumn<int32_t>(1, &c1.dict_page[0], c1.dict_page.size(), c1.pages);
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
      scanner.set_fused(fused != 0);
      scanner.set_device_materialise(mode == 2);
      vector<uint8_t> rows;
      const auto t0 = std::chrono::steady_clock::now();
      Status st = scanner.AssembleRows(n, &rows, &passed);
      const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (!st.ok() || passed != want) { std::fprintf(stderr, "scan failed: passed %lld want %lld\n", (long long)passed, (long long)want); return 1; }
      if (r > 0 && s < best) best = s;
      evals = scanner.fused_evals();
      windows = scanner.leaf_windows();
    }
    std::printf(", \"%s\": {\"ms\": %.3f, \"rows_per_s\": %.4g, \"fused_evals\": %lld, \"leaf_windows\": %lld}",
                mode == 2 ? "device_page_spans" : (mode == 1 ? "fused_host_materialise" : "per_leaf"),
                best * 1e3, n / best, (long long)evals, (long long)windows);
  }
  std::printf("}\n");
  return 0;
}
