// A C++ host that scans one table with several GPUs through nothing but include/fle_b200.h -- the way a scanner
// process of the reference would (one scanner per thread over disjoint row groups, hdfs-parquet-scanner.cc:1040,
// :1056-1060): no torch, no NCCL, no CUDA headers.
//
//   * the table is cut into row groups, every member takes a contiguous range of them (the arithmetic of
//     partition.rank_rows, what bench.py uses for BASELINE cfg4);
//   * one fle_ctx per member -- one per device when the box has several B200s, several contexts on device 0
//     otherwise (the exchange code is the same peer-store / poll code either way);
//   * fle_group_create_local joins them; every member thread issues the same sequence of
//     fle_eval_sum_product_group_dev / fle_eval_group_dev calls and ends up holding the SAME totals, which must equal
//     a row-at-a-time evaluation of the whole table on the host.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <thread>
#include <vector>

#include "fle_b200.h"
#include "host-test-util.h"

namespace {

struct Table {
  int64_t n;
  std::vector<uint32_t> shipdate, discount, quantity;  // dictionary codes: 2526 / 11 / 50 distinct values
  std::vector<int64_t> price;                          // l_extendedprice in cents (PLAIN)
};

uint64_t Next(uint64_t* s) {  // xorshift64*
  *s ^= *s >> 12;
  *s ^= *s << 25;
  *s ^= *s >> 27;
  return *s * 0x2545F4914F6CDD1DULL;
}

Table MakeTable(int64_t n) {
  Table t;
  t.n = n;
  uint64_t s = 0xF1E00400ULL;
  t.shipdate.resize(n); t.discount.resize(n); t.quantity.resize(n); t.price.resize(n);
  for (int64_t i = 0; i < n; ++i) {
    t.shipdate[i] = Next(&s) % 2526;
    t.discount[i] = Next(&s) % 11;
    t.quantity[i] = Next(&s) % 50;
    t.price[i] = 90000 + (int64_t)(Next(&s) % 10410000);
  }
  return t;
}

// [first_row, end_row) of `rank`: row groups dealt out contiguously (partition.rank_rows)
void RankRows(int64_t total, int world, int rank, int64_t rg_rows, int64_t* first, int64_t* end) {
  const int64_t nrg = (total + rg_rows - 1) / rg_rows;
  const int64_t lo = nrg * rank / world, hi = nrg * (rank + 1) / world;
  *first = lo * rg_rows;
  *end = hi * rg_rows < total ? hi * rg_rows : total;
}

struct Member {
  fle_ctx* ctx = NULL;
  int64_t first = 0, n = 0;
  void* planes[3] = {NULL, NULL, NULL};
  size_t plane_bytes[3] = {0, 0, 0};
  void* price = NULL;
  void* dict = NULL;
  void* bitmap = NULL;
  void* totals = NULL;  // [0] sum, [1] count
  int64_t got_sum[4] = {0, 0, 0, 0};
  uint64_t got_cnt[4] = {0, 0, 0, 0};
  int rc = FLE_OK;
};

#define CHECK_RC(m, expr) do { int _rc = (expr); if (_rc != FLE_OK && (m)->rc == FLE_OK) { (m)->rc = _rc; std::fprintf(stderr, "%s -> %d (%s)\n", #expr, _rc, fle_last_error((m)->ctx)); } } while (0)

void Upload(Member* m, const Table& t) {
  const int bws[3] = {12, 4, 6};
  const std::vector<uint32_t>* src[3] = {&t.shipdate, &t.discount, &t.quantity};
  for (int c = 0; c < 3; ++c) {
    const size_t cap = fle_encoded_bytes(m->n, bws[c]);
    std::vector<uint8_t> enc(cap + 16);
    size_t len = 0;
    CHECK_RC(m, fle_pack_u32(m->ctx, src[c]->data() + m->first, m->n, bws[c], enc.data(), cap, &len));
    m->plane_bytes[c] = len;
    CHECK_RC(m, fle_dev_alloc(m->ctx, len + 64, &m->planes[c]));
    CHECK_RC(m, fle_h2d(m->ctx, m->planes[c], enc.data(), len));
    CHECK_RC(m, fle_ctx_sync(m->ctx));  // (enc goes out of scope)
  }
  int64_t dict[11];
  for (int i = 0; i < 11; ++i) dict[i] = i;  // discount in cents
  CHECK_RC(m, fle_dev_alloc(m->ctx, (size_t)(m->n > 0 ? m->n : 1) * 8, &m->price));
  CHECK_RC(m, fle_dev_alloc(m->ctx, sizeof(dict), &m->dict));
  CHECK_RC(m, fle_dev_alloc(m->ctx, (size_t)((m->n + 63) / 64 + 1) * 8, &m->bitmap));
  CHECK_RC(m, fle_dev_alloc(m->ctx, 64, &m->totals));
  if (m->n > 0) CHECK_RC(m, fle_h2d(m->ctx, m->price, t.price.data() + m->first, (size_t)m->n * 8));
  CHECK_RC(m, fle_h2d(m->ctx, m->dict, dict, sizeof(dict)));
  CHECK_RC(m, fle_ctx_sync(m->ctx));
}

fle_pred_node Leaf(int op, int col, uint64_t code) {
  fle_pred_node nd;
  std::memset(&nd, 0, sizeof(nd));
  nd.op = op; nd.col = col; nd.code = code;
  return nd;
}

// Members that share ONE device (this test on a one-GPU box) must have their kernels loaded before the first group
// call: in immediate mode a member's kernel waits for its peers' kernels, and the CUDA runtime loads a kernel lazily at
// its first launch, which may need the device idle -- a peer whose kernel is not loaded yet would wait for the
// spinning one until the group time-out.  (One member per device, the real configuration, has no such coupling.)
void WarmUp(Member* m) {
  fle_col_desc cols[3];
  const int bws[3] = {12, 4, 6};
  for (int c = 0; c < 3; ++c) {
    std::memset(&cols[c], 0, sizeof(cols[c]));
    cols[c].planes = m->planes[c];
    cols[c].bytes = m->plane_bytes[c];
    cols[c].bit_width = bws[c];
  }
  const fle_pred_node prog[9] = {Leaf(FLE_OP_GE, 0, 1), Leaf(FLE_OP_LE, 0, 2), Leaf(FLE_OP_AND, 0, 0),
                                 Leaf(FLE_OP_GE, 1, 1), Leaf(FLE_OP_LE, 1, 2), Leaf(FLE_OP_AND, 0, 0),
                                 Leaf(FLE_OP_AND, 0, 0), Leaf(FLE_OP_LT, 2, 1), Leaf(FLE_OP_AND, 0, 0)};
  const fle_pred_node one[1] = {Leaf(FLE_OP_LT, 2, 1)};
  fle_ctx* solo[1] = {m->ctx};
  CHECK_RC(m, fle_group_create_local(solo, 1));  // a group of one: the same kernels, nobody to wait for
  CHECK_RC(m, fle_eval_sum_product_group_dev(m->ctx, cols, 3, prog, 9, 0, m->n, static_cast<const int64_t*>(m->price), 1,
                                              static_cast<const int64_t*>(m->dict), 11, static_cast<uint64_t*>(m->bitmap),
                                              static_cast<int64_t*>(m->totals), static_cast<uint64_t*>(m->totals) + 1));
  CHECK_RC(m, fle_eval_group_dev(m->ctx, cols, 3, one, 1, 0, m->n, static_cast<uint64_t*>(m->bitmap), static_cast<uint64_t*>(m->totals) + 1));
  CHECK_RC(m, fle_group_flush(m->ctx));
  CHECK_RC(m, fle_ctx_sync(m->ctx));
  CHECK_RC(m, fle_group_close(m->ctx));
}

// three queries per member, the same sequence on every member (like any collective)
void Scan(Member* m) {
  fle_col_desc cols[3];
  const int bws[3] = {12, 4, 6};
  for (int c = 0; c < 3; ++c) {
    std::memset(&cols[c], 0, sizeof(cols[c]));
    cols[c].planes = m->planes[c];
    cols[c].bytes = m->plane_bytes[c];
    cols[c].bit_width = bws[c];
  }
  const uint64_t his[3] = {1094, 1459, 2525};  // three date ranges: Q6 for 1994, for 1994-95, from 1994 on
  for (int q = 0; q < 3; ++q) {
    const fle_pred_node prog[9] = {Leaf(FLE_OP_GE, 0, 730), Leaf(FLE_OP_LE, 0, his[q]), Leaf(FLE_OP_AND, 0, 0),
                                   Leaf(FLE_OP_GE, 1, 5),   Leaf(FLE_OP_LE, 1, 7),      Leaf(FLE_OP_AND, 0, 0),
                                   Leaf(FLE_OP_AND, 0, 0),  Leaf(FLE_OP_LT, 2, 24),     Leaf(FLE_OP_AND, 0, 0)};
    int64_t* sum_dev = static_cast<int64_t*>(m->totals) + 2 * q;
    uint64_t* cnt_dev = static_cast<uint64_t*>(m->totals) + 2 * q + 1;
    CHECK_RC(m, fle_eval_sum_product_group_dev(m->ctx, cols, 3, prog, 9, 0, m->n, static_cast<const int64_t*>(m->price), 1,
                                                static_cast<const int64_t*>(m->dict), 11, static_cast<uint64_t*>(m->bitmap),
                                                sum_dev, cnt_dev));
  }
  // count-only form: quantity < 24 over the whole table
  const fle_pred_node one[1] = {Leaf(FLE_OP_LT, 2, 24)};
  CHECK_RC(m, fle_eval_group_dev(m->ctx, cols, 3, one, 1, 0, m->n, static_cast<uint64_t*>(m->bitmap),
                                 static_cast<uint64_t*>(m->totals) + 7));
  CHECK_RC(m, fle_group_flush(m->ctx));  // (deferred mode: the last call's totals are collected here)
  uint64_t host[8];
  CHECK_RC(m, fle_d2h(m->ctx, host, m->totals, sizeof(host)));
  CHECK_RC(m, fle_ctx_sync(m->ctx));
  for (int q = 0; q < 3; ++q) {
    m->got_sum[q] = (int64_t)host[2 * q];
    m->got_cnt[q] = host[2 * q + 1];
  }
  m->got_cnt[3] = host[7];
}

void RunGroup(int world, bool deferred, const Table& t) {
  const int ndev = fle_device_count();
  ASSERT_TRUE(ndev >= 1);
  std::vector<Member> ms(world);
  std::vector<fle_ctx*> ctxs(world);
  for (int r = 0; r < world; ++r) {
    ASSERT_TRUE(fle_ctx_create(r % ndev, NULL, &ms[r].ctx) == FLE_OK);
    ctxs[r] = ms[r].ctx;
    int64_t end = 0;
    RankRows(t.n, world, r, 64 * 500, &ms[r].first, &end);
    ms[r].n = end - ms[r].first;
    fle_ctx_set_option(ms[r].ctx, "group_deferred", deferred ? 1 : 0);
    fle_ctx_set_option(ms[r].ctx, "eval_jit", r % 2 == 0 ? 2 : 0);  // fused single kernel / two kernels + reduce kernel
    Upload(&ms[r], t);
    WarmUp(&ms[r]);
    EXPECT_EQ(ms[r].rc, FLE_OK);
  }
  ASSERT_TRUE(fle_group_create_local(ctxs.data(), world) == FLE_OK);
  std::vector<std::thread> th;
  for (int r = 0; r < world; ++r) th.emplace_back(Scan, &ms[r]);
  for (auto& x : th) x.join();
  // the whole table, row at a time
  const uint64_t his[3] = {1094, 1459, 2525};
  for (int q = 0; q < 3; ++q) {
    int64_t sum = 0;
    uint64_t cnt = 0;
    for (int64_t i = 0; i < t.n; ++i) {
      if (t.shipdate[i] >= 730 && t.shipdate[i] <= his[q] && t.discount[i] >= 5 && t.discount[i] <= 7 && t.quantity[i] < 24) {
        sum += t.price[i] * (int64_t)t.discount[i];
        ++cnt;
      }
    }
    for (int r = 0; r < world; ++r) {
      EXPECT_EQ(ms[r].rc, FLE_OK);
      EXPECT_EQ(ms[r].got_sum[q], sum);
      EXPECT_EQ((int64_t)ms[r].got_cnt[q], (int64_t)cnt);
    }
  }
  uint64_t cnt = 0;
  for (int64_t i = 0; i < t.n; ++i) cnt += t.quantity[i] < 24;
  for (int r = 0; r < world; ++r) EXPECT_EQ((int64_t)ms[r].got_cnt[3], (int64_t)cnt);
  for (int r = 0; r < world; ++r) {
    int rank = -1, w = -1;
    uint64_t done = 0, lost = 0;
    EXPECT_EQ(fle_group_status(ms[r].ctx, &rank, &w, &done, &lost), FLE_OK);
    EXPECT_EQ(rank, r);
    EXPECT_EQ(w, world);
    EXPECT_EQ((int64_t)lost, 0);
  }
  for (int r = 0; r < world; ++r) fle_group_close(ms[r].ctx);
  for (int r = 0; r < world; ++r) {
    for (int c = 0; c < 3; ++c) fle_dev_free(ms[r].ctx, ms[r].planes[c]);
    fle_dev_free(ms[r].ctx, ms[r].price);
    fle_dev_free(ms[r].ctx, ms[r].dict);
    fle_dev_free(ms[r].ctx, ms[r].bitmap);
    fle_dev_free(ms[r].ctx, ms[r].totals);
    fle_ctx_destroy(ms[r].ctx);
  }
}

const Table& TheTable() {
  static Table t = MakeTable(64 * 9000 + 37);
  return t;
}

}  // namespace

TEST(Group, TwoMembersImmediate) { RunGroup(2, false, TheTable()); }
TEST(Group, TwoMembersDeferred) { RunGroup(2, true, TheTable()); }
TEST(Group, FourMembersDeferred) { RunGroup(4, true, TheTable()); }
TEST(Group, ThreeMembersImmediate) { RunGroup(3, false, TheTable()); }
TEST(Group, EightMembersDeferredSomeEmpty) {  // more members than row groups' worth for some: an empty range still takes part
  static Table small = MakeTable(64 * 500 * 5 + 11);
  RunGroup(8, true, small);
}

int main() {
  // Several members share ONE device here.  A kernel that meets the driver for the first time is loaded lazily, which
  // can wait for the device to drain -- while the other members' kernels sit in the exchange waiting for this very
  // member: they would only get on by timing out (and count a lost peer).  One GPU per member, the library's real
  // deployment, has no such coupling; on a shared device load everything up front.
  setenv("CUDA_MODULE_LOADING", "EAGER", 0);  // (an explicit setting of the caller's wins)
  return RunAllTests();
}
