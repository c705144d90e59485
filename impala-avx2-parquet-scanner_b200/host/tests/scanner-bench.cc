// Throughput of the drop-in column-reader batch interface (exec/hdfs-parquet-scanner.h) on
// Impala-shaped row groups: three OPTIONAL FLE_DICTIONARY columns (def levels present, no NULLs), the
// TPC-H Q6 conjuncts attached as a SimplePredicate forest, AssembleRows = predicate evaluation +
// materialisation of the survivors as tuples (hdfs-parquet-scanner.cc:1078-1275).
//
// The page bodies of a column chunk lie back to back (each behind a pseudo page header, so the FLE
// sections start at odd offsets, as in a file) in one PINNED buffer -- what an integration gets by
// allocating its I/O buffers with fle_host_alloc or pinning them with fle_host_register.
//
//   scanner-bench [--json] [--rows N] [--rows-per-page P] [--row-groups R] [--threads a,b,c] [--reps K]
//
// Reports, per thread count T: T scanner threads (one library context each, as the reference runs one
// scanner per thread, hdfs-parquet-scanner.cc:1040), each scanning R row groups, row group k+1's upload
// started before AssembleRows(k) (StartRowGroupUpload) -- rows/s and page bytes/s of the whole job; next
// to it the same bytes copied pinned -> device and nothing else (the PCIe roofline of that job on this
// box).  Single-threaded per-path numbers (row-group batch / page spans / host materialisation / per
// leaf) and a byte-for-byte comparison of the tuples of all paths come first.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <string>
#include <thread>
#include <vector>

#include "exec/hdfs-parquet-scanner.h"
#include "exprs/simple-predicates.h"
#include "host-test-util.h"

using namespace impala;
using std::vector;

struct ColumnChunk {
  vector<uint8_t> dict_page;
  uint8_t* image = NULL;  // pinned: [pseudo header][page body] ...
  size_t image_bytes = 0;
  vector<HdfsParquetScanner::DataPage> pages;
};

static const int kPseudoHeaderBytes = 23;  // a thrift PageHeader of a data page is 20-odd bytes

// notnull (optional): values[i] is NULL where notnull[i] is false -- the page's FLE section then holds only the
// non-null values (hdfs-parquet-scanner.cc:886-921)
static ColumnChunk WriteColumn(const vector<int32_t>& values, int rows_per_page, const vector<bool>* notnull = NULL) {
  ColumnChunk cc;
  DictEncoder<int32_t> dict(4);
  const int n = values.size();
  vector<vector<int> > page_indices;
  for (int r = 0; r < n; ++r) {
    if (r % rows_per_page == 0) page_indices.push_back(vector<int>());
    if (notnull && !(*notnull)[r]) continue;
    if (dict.Put(values[r], page_indices.back()) < 0) { std::fprintf(stderr, "dictionary full\n"); exit(2); }
  }
  cc.dict_page.resize(dict.dict_encoded_size());
  dict.WriteDict(&cc.dict_page[0]);
  vector<vector<uint8_t> > bodies;
  size_t total = 0;
  for (size_t p = 0; p < page_indices.size(); ++p) {
    const size_t rows = std::min<size_t>(rows_per_page, n - p * (size_t)rows_per_page);
    vector<uint8_t> defbuf((rows / 64 + 3) * 8);
    FleEncoder def_levels(&defbuf[0], defbuf.size(), 1);
    for (size_t i = 0; i < rows; ++i) def_levels.Put(notnull ? (*notnull)[p * (size_t)rows_per_page + i] : 1);
    def_levels.Flush();
    int32_t num_def_bytes = def_levels.len();
    vector<uint8_t> body(4 + num_def_bytes);
    memcpy(&body[0], &num_def_bytes, 4);
    memcpy(&body[4], def_levels.buffer(), num_def_bytes);
    vector<uint8_t> data(1 + (rows / 64 + 3) * dict.bit_width() * 8);
    const int data_len = dict.WriteData(&data[0], data.size(), page_indices[p]);
    if (data_len <= 0) { std::fprintf(stderr, "WriteData failed\n"); exit(2); }
    body.insert(body.end(), data.begin(), data.begin() + data_len);
    total += kPseudoHeaderBytes + body.size();
    bodies.push_back(body);
  }
  void* pinned = NULL;
  if (fle_host_alloc(total + 64, &pinned) != FLE_OK) { std::fprintf(stderr, "fle_host_alloc failed\n"); exit(2); }
  cc.image = static_cast<uint8_t*>(pinned);
  cc.image_bytes = total;
  size_t off = 0;
  for (size_t p = 0; p < bodies.size(); ++p) {
    memset(cc.image + off, 0x15, kPseudoHeaderBytes);
    off += kPseudoHeaderBytes;
    memcpy(cc.image + off, &bodies[p][0], bodies[p].size());
    cc.pages.push_back(HdfsParquetScanner::DataPage(cc.image + off, bodies[p].size(),
                                                    std::min<size_t>(rows_per_page, n - p * (size_t)rows_per_page)));
    off += bodies[p].size();
  }
  return cc;
}

// l_extendedprice as the writer leaves it: int64, dictionary pages until the 40000-entry dictionary is full, PLAIN pages
// (values back to back) from there on; OPTIONAL, no NULLs; 64 KB pages
static ColumnChunk WriteMixedInt64Column(const vector<int64_t>& values, int rows_per_page, int dict_pages) {
  ColumnChunk cc;
  DictEncoder<int64_t> dict(8);
  const int n = values.size();
  vector<vector<int> > page_indices;
  for (int p = 0; p < dict_pages && p * rows_per_page < n; ++p) {
    page_indices.push_back(vector<int>());
    for (int r = p * rows_per_page; r < std::min(n, (p + 1) * rows_per_page); ++r)
      if (dict.Put(values[r], page_indices.back()) < 0) { std::fprintf(stderr, "dictionary full\n"); exit(2); }
  }
  cc.dict_page.resize(dict.dict_encoded_size());
  dict.WriteDict(&cc.dict_page[0]);
  vector<vector<uint8_t> > bodies;
  vector<int> page_rows;
  size_t total = 0;
  for (int p = 0; p * rows_per_page < n; ++p) {
    const int r0 = p * rows_per_page, m = std::min(rows_per_page, n - r0);
    vector<uint8_t> defbuf((m / 64 + 3) * 8);
    FleEncoder def_levels(&defbuf[0], defbuf.size(), 1);
    for (int i = 0; i < m; ++i) def_levels.Put(1);
    def_levels.Flush();
    int32_t nd = def_levels.len();
    vector<uint8_t> body(4 + nd);
    memcpy(&body[0], &nd, 4);
    memcpy(&body[4], def_levels.buffer(), nd);
    if (p < dict_pages) {
      vector<uint8_t> data(1 + (m / 64 + 3) * dict.bit_width() * 8);
      const int data_len = dict.WriteData(&data[0], data.size(), page_indices[p]);
      if (data_len <= 0) { std::fprintf(stderr, "WriteData failed\n"); exit(2); }
      body.insert(body.end(), data.begin(), data.begin() + data_len);
    } else {
      const uint8_t* v = reinterpret_cast<const uint8_t*>(&values[r0]);
      body.insert(body.end(), v, v + (size_t)m * 8);
    }
    total += kPseudoHeaderBytes + body.size();
    bodies.push_back(body);
    page_rows.push_back(m);
  }
  void* pinned = NULL;
  if (fle_host_alloc(total + 64, &pinned) != FLE_OK) { std::fprintf(stderr, "fle_host_alloc failed\n"); exit(2); }
  cc.image = static_cast<uint8_t*>(pinned);
  cc.image_bytes = total;
  size_t off = 0;
  for (size_t p = 0; p < bodies.size(); ++p) {
    memset(cc.image + off, 0x15, kPseudoHeaderBytes);
    off += kPseudoHeaderBytes;
    memcpy(cc.image + off, &bodies[p][0], bodies[p].size());
    HdfsParquetScanner::DataPage pg(cc.image + off, bodies[p].size(), page_rows[p]);
    if ((int)p >= dict_pages) pg.encoding = parquet_encoding::PLAIN;
    cc.pages.push_back(pg);
    off += bodies[p].size();
  }
  return cc;
}

struct Workload {
  ColumnChunk c[3];
  GeOperate<int32_t> ge0{0, 8730};
  LeOperate<int32_t> le0{0, 9094};
  AndOperate between0{&ge0, &le0};
  GeOperate<int32_t> ge1{1, 5};
  LeOperate<int32_t> le1{1, 7};
  LtOperate<int32_t> lt2{2, 24};
  int n = 0;
  int64_t want = 0;
  size_t bytes = 0;
  ColumnChunk* extra_i64 = NULL;  // a fourth, materialise-only int64 column (l_extendedprice)

  void Attach(HdfsParquetScanner* s) {
    for (int k = 0; k < 3; ++k) s->AddColumn<int32_t>(1, &c[k].dict_page[0], c[k].dict_page.size(), c[k].pages);
    if (extra_i64) s->AddColumn<int64_t>(1, &extra_i64->dict_page[0], extra_i64->dict_page.size(), extra_i64->pages);
    s->AddSimplePredicate(&between0);
    s->AddSimplePredicate(&ge1);
    s->AddSimplePredicate(&le1);
    s->AddSimplePredicate(&lt2);
  }
};

static double Now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// one scanner thread: R row groups, the next one's upload in flight while the current one is assembled
// pinned_rows: the survivors' tuples go straight into a pinned tuple buffer of the thread's (AssembleRowsInto)
static bool ScanRowGroups(Workload* w, int row_groups, int64_t* passed_total, int64_t* batches, bool pinned_rows = false) {
  static thread_local uint8_t* tuple_buf = NULL;
  static const size_t kTupleBuf = 64u << 20;
  if (pinned_rows && tuple_buf == NULL) {
    void* p = NULL;
    if (fle_host_alloc(kTupleBuf, &p) != FLE_OK) return false;
    tuple_buf = static_cast<uint8_t*>(p);  // (lives as long as the thread's benchmark runs; freed with the process)
  }
  std::unique_ptr<HdfsParquetScanner> cur(new HdfsParquetScanner()), next;
  w->Attach(cur.get());
  cur->StartRowGroupUpload();
  vector<uint8_t> rows;
  *passed_total = 0;
  *batches = 0;
  const bool trace = getenv("FLE_BENCH_TRACE") != NULL;
  double t_start = 0, t_asm = 0, t_new = 0;
  for (int k = 0; k < row_groups; ++k) {
    if (k + 1 < row_groups) {
      const double a = Now();
      next.reset(new HdfsParquetScanner());
      w->Attach(next.get());
      const double b = Now();
      next->StartRowGroupUpload();
      t_new += b - a;
      t_start += Now() - b;
    }
    rows.clear();
    int64_t passed = 0;
    const double a2 = Now();
    size_t tuple_bytes = 0;
    Status st = pinned_rows ? cur->AssembleRowsInto(w->n, tuple_buf, kTupleBuf, &passed, &tuple_bytes) : cur->AssembleRows(w->n, &rows, &passed);
    t_asm += Now() - a2;
    if (trace && k + 1 == row_groups)
      std::fprintf(stderr, "trace: %d row groups: new+attach %.3f ms, StartRowGroupUpload %.3f ms, AssembleRows %.3f ms\n", row_groups,
                   t_new * 1e3, t_start * 1e3, t_asm * 1e3);
    if (!st.ok() || passed != w->want) {
      std::fprintf(stderr, "scan failed: %s passed %lld want %lld\n", st.GetErrorMsg().c_str(), (long long)passed, (long long)w->want);
      return false;
    }
    *passed_total += passed;
    *batches += cur->row_group_batches();
    cur.swap(next);
  }
  return true;
}

// the same bytes, pinned -> device, nothing else
static bool CopyRowGroups(Workload* w, int row_groups) {
  fle_ctx* ctx = FleGpu::ctx();
  if (!ctx) return false;
  void* dev[4];
  for (int k = 0; k < 3; ++k) dev[k] = FleGpu::DeviceScratch(0, 50 + k, w->c[k].image_bytes + 64);
  dev[3] = w->extra_i64 ? FleGpu::DeviceScratch(0, 53, w->extra_i64->image_bytes + 64) : NULL;
  for (int r = 0; r < row_groups; ++r) {
    for (int k = 0; k < 3; ++k)
      if (fle_h2d_prefetch(ctx, dev[k], w->c[k].image, w->c[k].image_bytes) != FLE_OK) return false;
    if (w->extra_i64 && fle_h2d_prefetch(ctx, dev[3], w->extra_i64->image, w->extra_i64->image_bytes) != FLE_OK) return false;
    if (fle_prefetch_fence(ctx) != FLE_OK || fle_ctx_sync(ctx) != FLE_OK) return false;
  }
  return true;
}

// fn(t, false) once per thread before the clock starts (context creation, scratch growth, kernel load),
// then fn(t, true) timed
template <typename Fn>
static double RunThreads(int threads, Fn fn, bool* ok) {
  std::atomic<int> ready(0), go(0);
  std::atomic<bool> all_ok(true);
  vector<std::thread> th;
  vector<double> done_at(threads, 0.0);  // (thread exit frees the thread's context and scratch: not part of the job)
  for (int t = 0; t < threads; ++t) {
    th.emplace_back([&, t]() {
      if (!fn(t, false)) all_ok.store(false);
      ready.fetch_add(1);
      while (!go.load()) std::this_thread::yield();
      if (!fn(t, true)) all_ok.store(false);
      done_at[t] = Now();
    });
  }
  while (ready.load() < threads) std::this_thread::yield();
  const double t0 = Now();
  go.store(1);
  for (auto& x : th) x.join();
  double s = 0;
  for (int t = 0; t < threads; ++t) s = std::max(s, done_at[t] - t0);
  *ok = all_ok.load();
  return s;
}

int main(int argc, char** argv) {
  int n = 4194304, rows_per_page = 65536, reps = 5, row_groups = 8;
  vector<int> thread_counts = {1, 2, 4};
  bool json = false, extra = true;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--json") json = true;
    else if (a == "--no-extra") extra = false;  // skip the NULL / writer-shaped row groups
    else if (a == "--rows" && i + 1 < argc) n = atoi(argv[++i]);
    else if (a == "--rows-per-page" && i + 1 < argc) rows_per_page = atoi(argv[++i]);
    else if (a == "--reps" && i + 1 < argc) reps = atoi(argv[++i]);
    else if (a == "--row-groups" && i + 1 < argc) row_groups = atoi(argv[++i]);
    else if (a == "--threads" && i + 1 < argc) {
      thread_counts.clear();
      for (char* tok = strtok(argv[++i], ","); tok; tok = strtok(NULL, ",")) thread_counts.push_back(atoi(tok));
    } else if (i == 1 && atoi(argv[1]) > 0) {  // old positional form: rows [rows per page [reps]]
      n = atoi(argv[1]);
      if (argc > 2) rows_per_page = atoi(argv[2]);
      if (argc > 3) reps = atoi(argv[3]);
      break;
    }
  }
  (void)json;
  Workload w;
  w.n = n;
  srand(7);
  vector<int32_t> shipdate(n), discount(n), quantity(n);
  for (int i = 0; i < n; ++i) {
    shipdate[i] = 8000 + rand() % 2526;
    discount[i] = rand() % 11;
    quantity[i] = 1 + rand() % 50;
    w.want += shipdate[i] >= 8730 && shipdate[i] <= 9094 && discount[i] >= 5 && discount[i] <= 7 && quantity[i] < 24;
  }
  w.c[0] = WriteColumn(shipdate, rows_per_page);
  w.c[1] = WriteColumn(discount, rows_per_page);
  w.c[2] = WriteColumn(quantity, rows_per_page);
  for (int k = 0; k < 3; ++k) w.bytes += w.c[k].image_bytes;
  std::printf("{\"rows\": %d, \"rows_per_page\": %d, \"pages_per_column\": %zu, \"page_bytes_total\": %zu, \"survivors\": %lld, "
              "\"page_memory\": \"pinned (fle_host_alloc), pages of a column chunk back to back behind %d-byte pseudo headers\"",
              n, rows_per_page, w.c[0].pages.size(), w.bytes, (long long)w.want, kPseudoHeaderBytes);

  // ---- one thread, one row group, path by path; the tuples of every path must be identical ---------------
  vector<uint8_t> reference_rows;
  for (int mode = 3; mode >= 0; --mode) {  // 3 row-group batch, 2 page spans on the device, 1 fused + host materialisation, 0 per leaf
    double best = 1e30;
    int64_t passed = 0, evals = 0, windows = 0;
    const int mode_reps = mode == 0 ? 1 : reps;
    for (int r = 0; r < mode_reps + 1; ++r) {  // first repetition = warm-up (context creation, kernel load)
      HdfsParquetScanner scanner;
      w.Attach(&scanner);
      scanner.set_fused(mode >= 1);
      scanner.set_device_materialise(mode >= 2);
      scanner.set_row_group_batch(mode == 3);
      vector<uint8_t> rows;
      const double t0 = Now();
      Status st = scanner.AssembleRows(n, &rows, &passed);
      const double s = Now() - t0;
      if (!st.ok() || passed != w.want) { std::fprintf(stderr, "scan failed: passed %lld want %lld\n", (long long)passed, (long long)w.want); return 1; }
      if (mode == 3 && scanner.row_group_batches() != 1) { std::fprintf(stderr, "row-group batch path did not run\n"); return 1; }
      if (r == 0) {
        if (reference_rows.empty()) reference_rows = rows;
        else if (rows != reference_rows) { std::fprintf(stderr, "tuples of path %d differ from the row-group batch path\n", mode); return 1; }
      }
      if (r > 0 && s < best) best = s;
      evals = scanner.fused_evals();
      windows = scanner.leaf_windows();
    }
    std::printf(", \"%s\": {\"ms\": %.3f, \"rows_per_s\": %.4g, \"page_GBps\": %.3f, \"fused_evals\": %lld, \"leaf_windows\": %lld}",
                mode == 3 ? "row_group_batch" : mode == 2 ? "device_page_spans" : (mode == 1 ? "fused_host_materialise" : "per_leaf"),
                best * 1e3, n / best, w.bytes / best / 1e9, (long long)evals, (long long)windows);
  }
  std::printf(", \"tuples_identical_across_paths\": true");

  // ---- the same row group with NULLs in every column (one thread): still ONE device batch -----------------------
  if (extra) {
    const double null_frac = 0.02;
    Workload wn;
    wn.n = n;
    vector<bool> nn[3];
    for (int k = 0; k < 3; ++k) nn[k].resize(n);
    for (int i = 0; i < n; ++i) {
      for (int k = 0; k < 3; ++k) nn[k][i] = (rand() % 10000) >= null_frac * 10000;
      wn.want += nn[0][i] && nn[1][i] && nn[2][i] && shipdate[i] >= 8730 && shipdate[i] <= 9094 && discount[i] >= 5 &&
                 discount[i] <= 7 && quantity[i] < 24;
    }
    wn.c[0] = WriteColumn(shipdate, rows_per_page, &nn[0]);
    wn.c[1] = WriteColumn(discount, rows_per_page, &nn[1]);
    wn.c[2] = WriteColumn(quantity, rows_per_page, &nn[2]);
    for (int k = 0; k < 3; ++k) wn.bytes += wn.c[k].image_bytes;
    std::printf(", \"with_nulls\": {\"null_fraction\": %.2f, \"page_bytes_total\": %zu, \"survivors\": %lld", null_frac, wn.bytes, (long long)wn.want);
    vector<uint8_t> ref_rows;
    // 4: row-group batch by a scanner that was told to expect NULLs (set_expect_nulls: what the previous row group of
    // the file taught), 3: row-group batch finding out by itself (plain program first), 2: page spans
    for (int mode = 4; mode >= 2; --mode) {
      double best = 1e30;
      int64_t passed = 0;
      for (int r = 0; r < reps + 1; ++r) {
        HdfsParquetScanner scanner;
        wn.Attach(&scanner);
        scanner.set_row_group_batch(mode >= 3);
        scanner.set_expect_nulls(mode == 4);
        vector<uint8_t> rows;
        const double t0 = Now();
        Status st = scanner.AssembleRows(n, &rows, &passed);
        const double s = Now() - t0;
        if (!st.ok() || passed != wn.want) { std::fprintf(stderr, "NULL scan failed: passed %lld want %lld\n", (long long)passed, (long long)wn.want); return 1; }
        if (mode >= 3 && scanner.row_group_null_batches() != 1) { std::fprintf(stderr, "NULL row-group batch did not run\n"); return 1; }
        if (r == 0) {
          if (ref_rows.empty()) ref_rows = rows;
          else if (rows != ref_rows) { std::fprintf(stderr, "NULL tuples of path %d differ\n", mode); return 1; }
        }
        if (r > 0 && s < best) best = s;
      }
      std::printf(", \"%s\": {\"ms\": %.3f, \"rows_per_s\": %.4g, \"page_GBps\": %.3f}",
                  mode == 4 ? "row_group_batch_expecting_nulls" : mode == 3 ? "row_group_batch" : "device_page_spans", best * 1e3, n / best,
                  wn.bytes / best / 1e9);
    }
    std::printf(", \"tuples_identical_across_paths\": true}");
    for (int k = 0; k < 3; ++k) fle_host_free(wn.c[k].image);
  }

  // ---- pages as the reference's writer cuts them: 524225 rows per dictionary page (hdfs-parquet-table-writer.cc:387),
  //      i.e. no page boundary on a 64-row group, and another column cut differently ---------------------------------
  if (extra) {
    Workload wr;
    wr.n = n;
    wr.want = w.want;
    wr.c[0] = WriteColumn(shipdate, 524225);
    wr.c[1] = WriteColumn(discount, 524225);
    wr.c[2] = WriteColumn(quantity, 300007);
    for (int k = 0; k < 3; ++k) wr.bytes += wr.c[k].image_bytes;
    std::printf(", \"writer_shaped_pages\": {\"rows_per_page\": [524225, 524225, 300007], \"page_bytes_total\": %zu", wr.bytes);
    vector<uint8_t> ref_rows;
    for (int mode = 3; mode >= 2; --mode) {
      double best = 1e30;
      int64_t passed = 0;
      for (int r = 0; r < reps + 1; ++r) {
        HdfsParquetScanner scanner;
        wr.Attach(&scanner);
        scanner.set_row_group_batch(mode == 3);
        vector<uint8_t> rows;
        const double t0 = Now();
        Status st = scanner.AssembleRows(n, &rows, &passed);
        const double s = Now() - t0;
        if (!st.ok() || passed != wr.want) { std::fprintf(stderr, "ragged scan failed: passed %lld want %lld\n", (long long)passed, (long long)wr.want); return 1; }
        if (mode == 3 && scanner.row_group_batches() != 1) { std::fprintf(stderr, "ragged row-group batch did not run\n"); return 1; }
        if (r == 0) {
          if (ref_rows.empty()) ref_rows = rows;
          else if (rows != ref_rows) { std::fprintf(stderr, "ragged tuples of path %d differ\n", mode); return 1; }
        }
        if (r > 0 && s < best) best = s;
      }
      std::printf(", \"%s\": {\"ms\": %.3f, \"rows_per_s\": %.4g, \"page_GBps\": %.3f}", mode == 3 ? "row_group_batch" : "page_paths", best * 1e3,
                  n / best, wr.bytes / best / 1e9);
    }
    if (ref_rows != reference_rows) { std::fprintf(stderr, "ragged tuples differ from the aligned pages' tuples\n"); return 1; }
    std::printf(", \"tuples_identical_across_paths\": true}");
    for (int k = 0; k < 3; ++k) fle_host_free(wr.c[k].image);
  }

  // ---- Q6 as the scan node runs it: l_extendedprice materialised next to the predicate columns -- int64, too many
  //      distinct values for a dictionary: dictionary pages first, PLAIN pages from there on (8192 rows per 64 KB page) ----
  if (extra) {
    Workload wq;
    wq.n = n;
    wq.want = w.want;
    vector<int64_t> price(n);
    for (int i = 0; i < n; ++i) price[i] = 90000 + ((int64_t)rand() * 977 + rand()) % 10400000;
    wq.c[0] = WriteColumn(shipdate, 524225);
    wq.c[1] = WriteColumn(discount, 524225);
    wq.c[2] = WriteColumn(quantity, 524225);
    ColumnChunk cp = WriteMixedInt64Column(price, 8192, 4);
    for (int k = 0; k < 3; ++k) wq.bytes += wq.c[k].image_bytes;
    wq.bytes += cp.image_bytes;
    std::printf(", \"q6_with_extendedprice\": {\"price_pages\": %zu, \"price_dictionary_pages\": 4, \"page_bytes_total\": %zu", cp.pages.size(), wq.bytes);
    vector<uint8_t> ref_rows;
    for (int mode = 3; mode >= 2; --mode) {
      double best = 1e30;
      int64_t passed = 0;
      const int mode_reps = mode == 3 ? reps : 1;
      for (int r = 0; r < mode_reps + (mode == 3 ? 1 : 0); ++r) {
        HdfsParquetScanner scanner;
        wq.Attach(&scanner);
        scanner.AddColumn<int64_t>(1, &cp.dict_page[0], cp.dict_page.size(), cp.pages);
        scanner.set_row_group_batch(mode == 3);
        vector<uint8_t> rows;
        const double t0 = Now();
        Status st = scanner.AssembleRows(n, &rows, &passed);
        const double s = Now() - t0;
        if (!st.ok() || passed != wq.want) { std::fprintf(stderr, "price scan failed: passed %lld want %lld\n", (long long)passed, (long long)wq.want); return 1; }
        if (mode == 3 && scanner.row_group_batches() != 1) { std::fprintf(stderr, "price row-group batch did not run\n"); return 1; }
        if (ref_rows.empty()) ref_rows = rows;
        else if (r == 0 && rows != ref_rows) { std::fprintf(stderr, "price tuples of path %d differ\n", mode); return 1; }
        if ((mode != 3 || r > 0) && s < best) best = s;
      }
      std::printf(", \"%s\": {\"ms\": %.3f, \"rows_per_s\": %.4g, \"page_GBps\": %.3f}", mode == 3 ? "row_group_batch" : "page_paths", best * 1e3,
                  n / best, wq.bytes / best / 1e9);
    }
    std::printf(", \"tuples_identical_across_paths\": true");
    // the same row group through the pipelined scan loop (row group k + 1 uploading while k is assembled)
    wq.extra_i64 = &cp;
    std::printf(", \"threads\": [");
    for (int T = 1; T <= 2; ++T) {
      bool ok = true;
      double scan_s = 1e30, copy_s = 1e30;
      for (int r = 0; r < 4; ++r) {
        bool ok1 = true;
        const double s1 = RunThreads(T, [&](int, bool timed) {
          int64_t p, b;
          const int rg = timed ? row_groups : 2;
          return ScanRowGroups(&wq, rg, &p, &b) && b == rg;
        }, &ok1);
        ok = ok && ok1;
        if (r > 0) scan_s = std::min(scan_s, s1);
        const double c1 = RunThreads(T, [&](int, bool timed) { return CopyRowGroups(&wq, timed ? row_groups : 1); }, &ok1);
        ok = ok && ok1;
        if (r > 0) copy_s = std::min(copy_s, c1);
      }
      if (!ok) { std::fprintf(stderr, "threaded price scan failed\n"); return 1; }
      const double total_rows = (double)n * row_groups * T, total_bytes = (double)wq.bytes * row_groups * T;
      std::printf("%s{\"threads\": %d, \"ms\": %.3f, \"rows_per_s\": %.4g, \"page_GBps\": %.3f, \"h2d_only_GBps\": %.3f, \"frac_of_h2d\": %.3f}",
                  T > 1 ? ", " : "", T, scan_s * 1e3, total_rows / scan_s, total_bytes / scan_s / 1e9, total_bytes / copy_s / 1e9, copy_s / scan_s);
    }
    std::printf("]}");
    for (int k = 0; k < 3; ++k) fle_host_free(wq.c[k].image);
    fle_host_free(cp.image);
  }

  // ---- T scanner threads x R row groups ---------------------------------------------------------------------
  std::printf(", \"row_groups_per_thread\": %d, \"threads\": [", row_groups);
  double best_frac = 0, best_rows_per_s = 0, best_gbps = 0, best_peak = 0;
  int best_threads = 0;
  for (size_t ti = 0; ti < thread_counts.size(); ++ti) {
    const int T = thread_counts[ti];
    bool ok = true;
    double scan_s = 1e30, copy_s = 1e30;
    for (int r = 0; r < 6; ++r) {  // (best of five: several host threads sharing one copy engine scatter by +-10 %)
      bool ok1 = true;
      const double s = RunThreads(T, [&](int, bool timed) {
        int64_t p, b;
        const int rg = timed ? row_groups : 2;
        return ScanRowGroups(&w, rg, &p, &b) && b == rg;
      }, &ok1);
      ok = ok && ok1;
      if (r > 0) scan_s = std::min(scan_s, s);
      const double c = RunThreads(T, [&](int, bool timed) { return CopyRowGroups(&w, timed ? row_groups : 1); }, &ok1);
      ok = ok && ok1;
      if (r > 0) copy_s = std::min(copy_s, c);
    }
    if (!ok) { std::fprintf(stderr, "threaded scan failed\n"); return 1; }
    const double total_rows = (double)n * row_groups * T, total_bytes = (double)w.bytes * row_groups * T;
    // ... and with the tuples delivered into a pinned tuple buffer (AssembleRowsInto: no staging copy)
    double pin_s = 1e30;
    for (int r = 0; r < 4; ++r) {
      bool ok1 = true;
      const double s = RunThreads(T, [&](int, bool timed) {
        int64_t p, b;
        const int rg = timed ? row_groups : 2;
        return ScanRowGroups(&w, rg, &p, &b, true) && b == rg;
      }, &ok1);
      if (!ok1) { std::fprintf(stderr, "threaded scan (pinned rows) failed\n"); return 1; }
      if (r > 0) pin_s = std::min(pin_s, s);
    }
    std::printf("%s{\"threads\": %d, \"ms\": %.3f, \"rows_per_s\": %.4g, \"page_GBps\": %.3f, \"h2d_only_GBps\": %.3f, \"frac_of_h2d\": %.3f, "
                "\"pinned_rows\": {\"ms\": %.3f, \"rows_per_s\": %.4g, \"frac_of_h2d\": %.3f}}",
                ti ? ", " : "", T, scan_s * 1e3, total_rows / scan_s, total_bytes / scan_s / 1e9, total_bytes / copy_s / 1e9, copy_s / scan_s,
                pin_s * 1e3, total_rows / pin_s, copy_s / pin_s);
    if (copy_s / scan_s > best_frac) {
      best_frac = copy_s / scan_s;
      best_rows_per_s = total_rows / scan_s;
      best_gbps = total_bytes / scan_s / 1e9;
      best_peak = total_bytes / copy_s / 1e9;
      best_threads = T;
    }
  }
  std::printf("], \"best\": {\"threads\": %d, \"rows_per_s\": %.4g, \"page_GBps\": %.3f, \"h2d_only_GBps\": %.3f, \"frac_of_h2d\": %.3f}}\n",
              best_threads, best_rows_per_s, best_gbps, best_peak, best_frac);
  for (int k = 0; k < 3; ++k) fle_host_free(w.c[k].image);
  return 0;
}
