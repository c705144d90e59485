// extern "C" surface of libfle_b200.so (see include/fle_b200.h): contexts, memory helpers,
// the host-buffer entry points (H2D / D2H inside the call, chunked and overlapped with the
// kernels) and the device-pointer entry points.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>

#include "fle_internal.h"

namespace fle {

int set_error(fle_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->last_error = msg;
  return code;
}

int set_cuda_error(fle_ctx* ctx, cudaError_t e, const char* what) {
  if (ctx) {
    ctx->last_error = std::string(what) + ": " + cudaGetErrorString(e);
    // a kernel that did not run to its end leaves the self-resetting ticket / tile counters non-zero: the next
    // launch that uses them clears them first (reset_counters_if_dirty), or every later count would be wrong
    ctx->counters_dirty = true;
  }
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return FLE_E_NODEVICE;
  if (e == cudaErrorMemoryAllocation) return FLE_E_NOMEM;
  return FLE_E_CUDA;
}

struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
      switched = cudaSetDevice(dev) == cudaSuccess;
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

int ensure_scratch(fle_ctx* ctx, int nblocks, int n_in_lits) {
  if (ctx->counters_dirty) {  // (every launcher that hands d_ticket to a kernel comes through here first)
    ctx->counters_dirty = false;
    FLE_CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_ticket, 0, 8, ctx->stream));
  }
  if (nblocks > ctx->partial_cap) {
    const int cap = std::max(nblocks, 4096);
    uint64_t* p = nullptr;
    FLE_CUDA_TRY(ctx, cudaMalloc(&p, (size_t)cap * 8));
    if (ctx->d_partial) {
      FLE_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
      cudaFree(ctx->d_partial);
    }
    ctx->d_partial = p;
    ctx->partial_cap = cap;
  }
  (void)n_in_lits;
  return FLE_OK;
}

// Everything a first launch would allocate lazily -- the per-CTA partials, the literal slots with their pinned twins and
// events -- allocated now.  cudaMalloc / cudaMallocHost may wait for the device to drain: inside a group call that is
// a wait for the peers' kernels, which sit in the exchange waiting for THIS member's kernel (they only get on by timing
// out and count a lost peer).  A member that joins a group therefore allocates up front (fle_group.cu).
int prime_context(fle_ctx* ctx) {
  int rc = ensure_scratch(ctx, 4096, 0);
  if (rc != FLE_OK) return rc;
  for (int i = 0; i < fle_ctx::kLitSlots; ++i) {
    fle_ctx::LitSlot& S = ctx->lit_slots[i];
    if (S.cap == 0) {
      FLE_CUDA_TRY(ctx, cudaMalloc(&S.d, (size_t)4096 * 4));
      FLE_CUDA_TRY(ctx, cudaMallocHost(&S.h, (size_t)4096 * 4));
      S.cap = 4096;
    }
    if (!S.ev) FLE_CUDA_TRY(ctx, cudaEventCreateWithFlags(&S.ev, cudaEventDisableTiming));
  }
  return FLE_OK;
}

int upload_literals(fle_ctx* ctx, const void* words, int nwords, int* slot, uint64_t* gen, const uint32_t** dev) {
  const int si = ctx->lit_next;
  ctx->lit_next = (ctx->lit_next + 1) % fle_ctx::kLitSlots;
  fle_ctx::LitSlot& S = ctx->lit_slots[si];
  if (S.ev_pending) {  // the ring wrapped around: this slot's previous copy must have read the pinned side
    FLE_CUDA_TRY(ctx, cudaEventSynchronize(S.ev));
    S.ev_pending = false;
  }
  if (nwords > S.cap) {
    const int cap = std::max(nwords, 4096);
    if (S.d) {
      // the device side may still be read by kernels in flight: free it in stream order
      FLE_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
      cudaFree(S.d);
      cudaFreeHost(S.h);
      S.d = nullptr;
      S.h = nullptr;
      S.cap = 0;
    }
    FLE_CUDA_TRY(ctx, cudaMalloc(&S.d, (size_t)cap * 4));
    FLE_CUDA_TRY(ctx, cudaMallocHost(&S.h, (size_t)cap * 4));
    S.cap = cap;
  }
  if (!S.ev) FLE_CUDA_TRY(ctx, cudaEventCreateWithFlags(&S.ev, cudaEventDisableTiming));
  std::memcpy(S.h, words, (size_t)nwords * 4);
  FLE_CUDA_TRY(ctx, cudaMemcpyAsync(S.d, S.h, (size_t)nwords * 4, cudaMemcpyHostToDevice, ctx->stream));
  cudaStreamCaptureStatus cap_st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(ctx->stream, &cap_st) != cudaSuccess) cudaGetLastError();
  if (cap_st == cudaStreamCaptureStatusNone) {  // (an event recorded during capture cannot be waited for by the host)
    FLE_CUDA_TRY(ctx, cudaEventRecord(S.ev, ctx->stream));
    S.ev_pending = true;
  }
  S.gen = ++ctx->lit_gen_counter;
  if (slot) *slot = si;
  if (gen) *gen = S.gen;
  if (dev) *dev = S.d;
  return FLE_OK;
}

static int ensure_stage(fle_ctx* ctx, int idx, size_t bytes) {
  if (bytes <= ctx->d_stage_cap[idx]) return FLE_OK;
  if (ctx->d_stage[idx]) {
    FLE_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_stage[idx]);
    ctx->d_stage[idx] = nullptr;
    ctx->d_stage_cap[idx] = 0;
  }
  FLE_CUDA_TRY(ctx, cudaMalloc(&ctx->d_stage[idx], bytes));
  ctx->d_stage_cap[idx] = bytes;
  return FLE_OK;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------
// nullable columns (hdfs-parquet-scanner.cc:333-445): per leaf
//   notnull = def_levels.Eq(n, max_def_level)
//   data    = dict.Op(popcount(notnull), literal)        (indexed by non-null ordinal)
//   leaf    = IntersectBitset(notnull, data)
// When no column actually contains a NULL the deposit is the identity and the columns are
// evaluated by the fused kernel as if they were required.
// ------------------------------------------------------------------------------------------
// grow-only scratch of the NULL path (no stream-ordered allocations per call); returns the base or nullptr
static uint8_t* null_scratch(fle_ctx* ctx, size_t bytes) {
  if (ctx->null_cap < bytes) {
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return nullptr;
    if (ctx->d_null) cudaFree(ctx->d_null);
    ctx->d_null = nullptr;
    ctx->null_cap = 0;
    if (cudaMalloc(&ctx->d_null, bytes) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    ctx->null_cap = bytes;
  }
  return static_cast<uint8_t*>(ctx->d_null);
}

static int eval_dev_nullable(fle_ctx* ctx, const fle_col_desc* cols, int ncols,
                             const fle_pred_node* nodes, int nnodes, int64_t first_row, int64_t n,
                             uint64_t* bitmap_dev, uint64_t* count_dev) {
  if (first_row != 0)
    return set_error(ctx, FLE_E_UNSUPPORTED, "nullable columns are evaluated from row 0 of the page");
  const int64_t nwords = (n + 63) / 64;
  if (nwords == 0) {
    if (count_dev) FLE_CUDA_TRY(ctx, cudaMemsetAsync(count_dev, 0, 8, ctx->stream));
    return FLE_OK;
  }
  for (int c = 0; c < ncols; ++c)
    if (cols[c].max_def_level > 1) return set_error(ctx, FLE_E_UNSUPPORTED, "max_def_level > 1 (nested types)");
  // 0. (host only) A NULL row yields 0 in every leaf, and the deposit through the not-null
  // bitmap (IntersectBitset) is a bit-wise injection, so it distributes over AND and OR:
  // deposit(a) op deposit(b) == deposit(a op b).  Every maximal subtree whose leaves sit on ONE
  // nullable column is therefore evaluated as one fused program in the data-ordinal domain and
  // deposited once -- straight from the column's def-level planes (no not-null bitmap is
  // built); every maximal subtree over required columns only is one fused program over the rows.
  constexpr int kReq = -1, kMixed = -2;
  std::vector<int> node_start(nnodes, 0), node_cls(nnodes, kMixed);
  {
    std::vector<int> stack;  // node index of the subtree roots
    for (int i = 0; i < nnodes; ++i) {
      const fle_pred_node& nd = nodes[i];
      if (nd.op == FLE_OP_AND || nd.op == FLE_OP_OR) {
        if (stack.size() < 2) return set_error(ctx, FLE_E_INVALID, "malformed postfix program");
        const int rb = stack.back(); stack.pop_back();
        const int ra = stack.back(); stack.pop_back();
        node_start[i] = node_start[ra];
        node_cls[i] = (node_cls[ra] == node_cls[rb] && node_cls[ra] != kMixed) ? node_cls[ra] : kMixed;
      } else {
        const bool is_const = nd.op == FLE_OP_ALL || nd.op == FLE_OP_NONE;
        const bool in_range = nd.col >= 0 && nd.col < ncols;
        if (!is_const && !in_range) return set_error(ctx, FLE_E_INVALID, "leaf column out of range");
        const bool nullable = in_range && cols[nd.col].max_def_level != 0;
        node_start[i] = i;
        node_cls[i] = nullable ? nd.col : (is_const ? kMixed : kReq);  // a constant without a nullable column stays a constant
      }
      stack.push_back(i);
    }
    if (stack.size() != 1) return set_error(ctx, FLE_E_INVALID, "malformed postfix program");
  }
  // maximal single-class subtrees, found from the root downwards ([node_start, i] is contiguous in postfix)
  std::vector<std::pair<int, int> > subs;  // [start, end]
  for (int i = nnodes - 1; i >= 0;) {
    if (node_cls[i] != kMixed) {
      subs.push_back(std::make_pair(node_start[i], i));
      i = node_start[i] - 1;
    } else {
      --i;
    }
  }
  std::sort(subs.begin(), subs.end());
  // what joins the subtrees: ONE operator throughout (a conjunct chain: the usual case) lets every subtree fold
  // into the output bitmap as it is produced; mixed operators (or constants) take leaf bitmaps + a combine pass
  int join_op = 0, nreq = 0;
  bool uniform = true;
  {
    size_t si = 0;
    for (int i = 0; i < nnodes; ++i) {
      if (si < subs.size() && i >= subs[si].first) {
        if (node_cls[subs[si].second] == kReq) ++nreq;
        i = subs[si].second;
        ++si;
        continue;
      }
      const int op = nodes[i].op;
      if (op != FLE_OP_AND && op != FLE_OP_OR) { uniform = false; continue; }  // a constant leaf
      if (join_op == 0) join_op = op;
      else if (join_op != op) uniform = false;
    }
  }
  // the one-kernel NULL path (fle_null_chain.cu): one operator throughout, every subtree one pass on one column
  NullChainPlan plan;
  bool chain = false;
  if (ctx->null_fused >= 2 && uniform && bitmap_dev && !subs.empty() && (int)subs.size() <= kNcMaxCols) {
    chain = true;
    plan.ncols = (int)subs.size();
    plan.join_op = join_op == FLE_OP_OR ? FOLD_OR : FOLD_AND;
    for (size_t k = 0; k < subs.size() && chain; ++k) {
      const int s0 = subs[k].first, s1 = subs[k].second;
      int col = node_cls[s1];
      if (col == kReq) {  // a required subtree: exactly one column
        for (int i = s0; i <= s1; ++i) {
          if (nodes[i].op == FLE_OP_AND || nodes[i].op == FLE_OP_OR) continue;
          if (col == kReq) col = nodes[i].col;
          else if (col != nodes[i].col) chain = false;
        }
      }
      if (!chain || col < 0) { chain = false; break; }
      if ((reinterpret_cast<uintptr_t>(cols[col].planes) & 15) || cols[col].bit_width < 1) { chain = false; break; }
      if (cols[col].max_def_level == 0 && cols[col].bytes < (uint64_t)nwords * (uint64_t)cols[col].bit_width * 8u) { chain = false; break; }
      plan.col[k] = col;
      if (null_chain_pass(ctx, cols[col], nodes + s0, s1 - s0 + 1, &plan.pd[k]) != FLE_OK) chain = false;
    }
    if (chain) {  // (columns too wide for the kernel's two stages: the per-column kernels)
      int bws[kNcMaxCols];
      for (int k = 0; k < plan.ncols; ++k) bws[k] = cols[plan.col[k]].bit_width;
      chain = null_chain_fits(ctx, bws, plan.ncols);
    }
  }
  // 1. non-null rows per column (one launch, nothing written but the counts -- and, when the program is a chain the
  // one-kernel NULL path can take, the per-tile counts that path starts from) ...
  constexpr size_t kCountsBytes = 256;
  const int64_t ntiles = null_chain_tiles(n);
  const size_t sums_stride = ((size_t)ntiles + 1 + 31) / 32 * 32;  // entries per column
  uint8_t* scratch = null_scratch(ctx, kCountsBytes + (chain ? (size_t)ncols * sums_stride * 12 : 0));
  if (!scratch) return set_error(ctx, FLE_E_NOMEM, "scratch of the NULL path");
  uint64_t* d_counts = reinterpret_cast<uint64_t*>(scratch);
  uint64_t* d_pref = reinterpret_cast<uint64_t*>(scratch + kCountsBytes);  // [ncols][sums_stride] prefixes, then the counts
  uint32_t* d_sums = reinterpret_cast<uint32_t*>(d_pref + (size_t)ncols * sums_stride);
  int rc = chain ? launch_def_tile_counts(ctx, cols, ncols, n, d_counts, d_sums, sums_stride)
                 : launch_def_counts(ctx, cols, ncols, n, d_counts);
  if (rc != FLE_OK) return rc;
  // 2a. ... and, without waiting for them, the fused kernel over the columns as if they were
  // required: with no NULL anywhere (TPC-H; Impala's writer marks every column OPTIONAL) the data
  // ordinal IS the row index and this is the answer.  (Skipped when a data stream is too short to
  // hold n values -- then there are NULLs for sure.)
  std::vector<fle_col_desc> req(cols, cols + ncols);
  for (auto& d : req) { d.max_def_level = 0; d.def_planes = nullptr; d.def_bytes = 0; }
  bool optimistic = true;
  for (int c = 0; c < ncols; ++c)
    if (cols[c].bytes < (uint64_t)nwords * (uint64_t)cols[c].bit_width * 8u) optimistic = false;
  // (a context whose previous nullable call DID find NULLs -- a file that has them usually has them in every row
  //  group -- launches the NULL path right away when it is the one-kernel chain, which is correct with or without
  //  NULLs and needs nothing from the host; else it waits for the counts instead of launching a scan it would throw away)
  const bool launch_first = optimistic && !ctx->nulls_seen;
  const bool chain_first = chain && !launch_first;
  if (launch_first) {
    rc = launch_eval(ctx, req.data(), ncols, nodes, nnodes, 0, n, bitmap_dev, count_dev);
    if (rc != FLE_OK) return rc;
  } else if (chain_first) {
    rc = launch_null_chain(ctx, plan, cols, n, d_sums, d_pref, sums_stride, bitmap_dev, count_dev);
    if (rc != FLE_OK) return rc;
  }
  std::vector<uint64_t> h_counts(ncols, 0);
  FLE_CUDA_TRY(ctx, cudaMemcpyAsync(h_counts.data(), d_counts, (size_t)ncols * 8, cudaMemcpyDeviceToHost, ctx->stream));
  FLE_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  bool any_null = false;
  for (int c = 0; c < ncols; ++c)
    if (cols[c].max_def_level != 0 && (int64_t)h_counts[c] != n) any_null = true;
  ctx->nulls_seen = any_null;
  if (chain && (chain_first || any_null)) {
    // (the kernel never reads past a column buffer; a stream too short for its non-null values is reported here)
    for (int i = 0; i < plan.ncols; ++i) {
      const fle_col_desc& col = cols[plan.col[i]];
      const uint64_t vals = col.max_def_level != 0 ? h_counts[plan.col[i]] : (uint64_t)n;
      if (col.bytes < (vals + 63) / 64 * (uint64_t)col.bit_width * 8u)
        return set_error(ctx, FLE_E_BUFFER, "column shorter than its non-null values");
    }
    if (chain_first) return FLE_OK;
    return launch_null_chain(ctx, plan, cols, n, d_sums, d_pref, sums_stride, bitmap_dev, count_dev);
  }
  if (!any_null) {
    if (!optimistic) return set_error(ctx, FLE_E_BUFFER, "column shorter than the rows requested");
    if (!launch_first) return launch_eval(ctx, req.data(), ncols, nodes, nnodes, 0, n, bitmap_dev, count_dev);
    return FLE_OK;
  }

  // 2b. NULLs present, per column subtree (see the classification above)
  const bool fold = uniform && nreq <= 1 && bitmap_dev != nullptr;
  const size_t word_bytes = ((size_t)nwords * 8 + 255) / 256 * 256;
  const size_t need = kCountsBytes + word_bytes * (fold ? 1 : subs.size() + 1);
  scratch = null_scratch(ctx, need);
  if (!scratch) return set_error(ctx, FLE_E_NOMEM, "scratch of the NULL path");
  uint64_t* db = reinterpret_cast<uint64_t*>(scratch + kCountsBytes);  // data-domain bitmap of the subtree in flight (stream-ordered reuse)
  size_t bump = kCountsBytes + word_bytes;

  // the data-domain bitmap of a subtree on nullable column cls: false = the subtree is empty (no value, or NONE)
  auto data_bitmap = [&](std::vector<fle_pred_node>& sub, int cls, int64_t* m_out) -> int {
    const fle_col_desc& col = cols[cls];
    const int64_t m = (int64_t)h_counts[cls];  // non-null values = rows of the data stream
    *m_out = m;
    const bool single = sub.size() == 1;
    if (m == 0 || (single && sub[0].op == FLE_OP_NONE)) {
      *m_out = 0;
      return FLE_OK;
    }
    if (single && sub[0].op == FLE_OP_ALL) {
      cudaError_t e = cudaMemsetAsync(db, 0xFF, (size_t)((m + 63) / 64) * 8, ctx->stream);
      return e == cudaSuccess ? (int)FLE_OK : set_cuda_error(ctx, e, "cudaMemsetAsync(data bitmap)");
    }
    fle_col_desc dc{col.planes, col.bytes, col.bit_width, 0, nullptr, 0};
    for (auto& nd : sub)
      if (nd.op != FLE_OP_AND && nd.op != FLE_OP_OR) nd.col = 0;
    return launch_eval(ctx, &dc, 1, sub.data(), (int)sub.size(), 0, m, db, nullptr);
  };
  // one nullable subtree -> out (comb: FOLD_SET / FOLD_AND / FOLD_OR with what out holds): evaluate + deposit in one
  // kernel when the subtree is a single pass (fle_eval1_deposit.cu), else data-domain bitmap + deposit
  auto deposit_subtree = [&](std::vector<fle_pred_node>& sub, int cls, uint64_t* out, int comb, uint64_t* cnt) -> int {
    const uint64_t* def = reinterpret_cast<const uint64_t*>(cols[cls].def_planes);
    if (ctx->null_fused) {
      std::vector<fle_pred_node> s0(sub);
      for (auto& nd : s0)
        if (nd.op != FLE_OP_AND && nd.op != FLE_OP_OR) nd.col = 0;
      const int r = launch_eval_deposit(ctx, cols[cls], s0.data(), (int)s0.size(), (int64_t)h_counts[cls], def, n, out, comb, cnt);
      if (r != FLE_E_UNSUPPORTED) return r;
    }
    int64_t m = 0;
    int r = data_bitmap(sub, cls, &m);
    if (r != FLE_OK) return r;
    return launch_deposit(ctx, def, n, db, m, out, comb, cnt);  // (an empty subtree deposits nothing: m = 0 makes every gathered bit 0)
  };
  auto eval_required = [&](std::vector<fle_pred_node>& sub, uint64_t* out, uint64_t* cnt) -> int {
    std::vector<int> map(ncols, -1);  // remap the columns the subtree uses to 0..k-1
    std::vector<fle_col_desc> dc;
    for (auto& nd : sub) {
      if (nd.op == FLE_OP_AND || nd.op == FLE_OP_OR) continue;
      if (map[nd.col] < 0) { map[nd.col] = (int)dc.size(); dc.push_back(cols[nd.col]); }
      nd.col = map[nd.col];
    }
    return launch_eval(ctx, dc.data(), (int)dc.size(), sub.data(), (int)sub.size(), 0, n, out, cnt);
  };

  if (fold) {
    // the required subtree (if any) sets the output, every nullable subtree folds its deposit into it
    std::vector<size_t> order;
    for (size_t k = 0; k < subs.size(); ++k)
      if (node_cls[subs[k].second] == kReq) order.push_back(k);
    for (size_t k = 0; k < subs.size(); ++k)
      if (node_cls[subs[k].second] != kReq) order.push_back(k);
    for (size_t oi = 0; oi < order.size(); ++oi) {
      const int s0 = subs[order[oi]].first, s1 = subs[order[oi]].second, cls = node_cls[s1];
      const bool first = oi == 0, last = oi + 1 == order.size();
      std::vector<fle_pred_node> sub(nodes + s0, nodes + s1 + 1);
      if (cls == kReq) {
        rc = eval_required(sub, bitmap_dev, last ? count_dev : nullptr);
      } else {
        const int comb = first ? FOLD_SET : (join_op == FLE_OP_AND ? FOLD_AND : FOLD_OR);
        rc = deposit_subtree(sub, cls, bitmap_dev, comb, last ? count_dev : nullptr);
      }
      if (rc != FLE_OK) return rc;
    }
    return FLE_OK;
  }

  std::vector<const uint64_t*> leaf_bm;
  std::vector<fle_pred_node> prog;  // the combine program: mixed nodes + one leaf per subtree
  size_t si = 0;
  for (int i = 0; i < nnodes; ++i) {
    if (si >= subs.size() || i < subs[si].first) {  // mixed operator or plain constant: copied
      prog.push_back(nodes[i]);
      continue;
    }
    const int s0 = subs[si].first, s1 = subs[si].second, cls = node_cls[s1];
    ++si;
    i = s1;
    uint64_t* lb = reinterpret_cast<uint64_t*>(scratch + bump);
    bump += word_bytes;
    std::vector<fle_pred_node> sub(nodes + s0, nodes + s1 + 1);
    if (cls == kReq) {
      rc = eval_required(sub, lb, nullptr);
    } else {
      rc = deposit_subtree(sub, cls, lb, FOLD_SET, nullptr);
    }
    if (rc != FLE_OK) return rc;
    fle_pred_node ref{FLE_OP_EQ, (int32_t)leaf_bm.size(), 0, nullptr, 0, 0};  // any leaf opcode: the combine kernel only looks at .col
    prog.push_back(ref);
    leaf_bm.push_back(lb);
  }
  return launch_bitmap_combine(ctx, leaf_bm.data(), (int)leaf_bm.size(), prog.data(), (int)prog.size(), n, bitmap_dev, count_dev);
}

static bool any_nullable(const fle_col_desc* cols, int ncols) {
  for (int c = 0; c < ncols; ++c)
    if (cols[c].max_def_level != 0) return true;
  return false;
}

}  // namespace fle

using namespace fle;

#define FLE_CHECK_CTX(ctx) \
  if (!(ctx)) return FLE_E_INVALID; \
  fle::DeviceGuard _guard((ctx)->device)

extern "C" {

int fle_version(void) { return FLE_B200_VERSION; }

int fle_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

const char* fle_strerror(int code) {
  switch (code) {
    case FLE_OK: return "ok";
    case FLE_E_INVALID: return "invalid argument";
    case FLE_E_NOMEM: return "out of memory";
    case FLE_E_CUDA: return "CUDA error";
    case FLE_E_BUFFER: return "buffer too small";
    case FLE_E_ALIGN: return "misaligned pointer or row offset";
    case FLE_E_UNSUPPORTED: return "unsupported request";
    case FLE_E_NODEVICE: return "no CUDA device (this library has no CPU path)";
    case FLE_E_DICT_FULL: return "dictionary full";
  }
  return "unknown error";
}

const char* fle_last_error(fle_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

int fle_ctx_create(int device, void* cuda_stream, fle_ctx** out) {
  if (!out) return FLE_E_INVALID;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return FLE_E_NODEVICE;
  }
  if (device < 0 || device >= ndev) return FLE_E_INVALID;
  fle_ctx* ctx = new (std::nothrow) fle_ctx();
  if (!ctx) return FLE_E_NOMEM;
  ctx->device = device;
  fle::DeviceGuard guard(device);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    delete ctx;
    return FLE_E_CUDA;
  }
  if (prop.major < 10) {  // kernels are built for sm_100a only
    delete ctx;
    return FLE_E_NODEVICE;
  }
  ctx->num_sms = prop.multiProcessorCount;
  ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  ctx->max_smem_per_sm = (int)prop.sharedMemPerMultiprocessor;
  if (cuda_stream) {
    ctx->stream = (cudaStream_t)cuda_stream;
  } else {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete ctx;
      return FLE_E_CUDA;
    }
    ctx->own_stream = true;
  }
  bool ok = cudaMalloc(&ctx->d_ticket, 8) == cudaSuccess &&  // [0] completion ticket, [1] tile counter
           
            cudaMemset(ctx->d_ticket, 0, 8) == cudaSuccess &&
            cudaMalloc(&ctx->d_count, 8) == cudaSuccess &&
            cudaMallocHost(&ctx->h_count, 8) == cudaSuccess &&
            cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < 2 && ok; ++i) {
    ok = cudaEventCreateWithFlags(&ctx->ev_copy[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming) == cudaSuccess;
  }
  if (!ok || ensure_scratch(ctx, 4096, 0) != FLE_OK) {
    fle_ctx_destroy(ctx);
    return FLE_E_CUDA;
  }
  *out = ctx;
  return FLE_OK;
}

int fle_ctx_destroy(fle_ctx* ctx) {
  if (!ctx) return FLE_OK;
  fle::DeviceGuard guard(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
  }
  for (int i = 0; i < 2; ++i) {
    if (ctx->ev_copy[i]) cudaEventDestroy(ctx->ev_copy[i]);
    if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
    if (ctx->d_stage[i]) cudaFree(ctx->d_stage[i]);
  }
  if (ctx->d_partial) cudaFree(ctx->d_partial);
  if (ctx->d_ticket) cudaFree(ctx->d_ticket);
  for (auto& S : ctx->lit_slots) {
    if (S.d) cudaFree(S.d);
    if (S.h) cudaFreeHost(S.h);
    if (S.ev) cudaEventDestroy(S.ev);
  }
  if (ctx->d_count) cudaFree(ctx->d_count);
  if (ctx->d_split) cudaFree(ctx->d_split);
  if (ctx->d_chunk_counts) cudaFree(ctx->d_chunk_counts);
  if (ctx->d_null) cudaFree(ctx->d_null);
  if (ctx->d_realign) cudaFree(ctx->d_realign);
  if (ctx->d_dict) cudaFree(ctx->d_dict);
  if (ctx->d_tile_state) cudaFree(ctx->d_tile_state);
  if (ctx->ev_prefetch) cudaEventDestroy(ctx->ev_prefetch);
  for (int i = 0; i < fle_ctx::kPrefetchMarks; ++i)
    if (ctx->ev_mark[i]) cudaEventDestroy(ctx->ev_mark[i]);
  fle::group_release(ctx);
  if (ctx->h_count) cudaFreeHost(ctx->h_count);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return FLE_OK;
}

int fle_ctx_set_stream(fle_ctx* ctx, void* cuda_stream) {
  FLE_CHECK_CTX(ctx);
  FLE_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream) {
    cudaStreamDestroy(ctx->stream);
    ctx->own_stream = false;
  }
  if (cuda_stream) {
    ctx->stream = (cudaStream_t)cuda_stream;
  } else {
    FLE_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  return FLE_OK;
}

int fle_ctx_sync(fle_ctx* ctx) {
  FLE_CHECK_CTX(ctx);
  FLE_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return FLE_OK;
}

int fle_ctx_set_option(fle_ctx* ctx, const char* name, int64_t value) {
  if (!ctx || !name) return FLE_E_INVALID;
  const std::string k(name);
  if (k == "eval_variant") ctx->eval_variant = (int)value;
  else if (k == "eval_stages") ctx->eval_stages = (int)value;
  else if (k == "eval_block") ctx->eval_block = (int)value;
  else if (k == "eval_ctas_per_sm") ctx->eval_ctas_per_sm = (int)value;
  else if (k == "eval_gpt") ctx->eval_gpt = (int)value;
  else if (k == "in_lookup_min") ctx->in_lookup_min = (int)value;
  else if (k == "null_fused") ctx->null_fused = (int)std::max<int64_t>(0, std::min<int64_t>(value, 2));
  else if (k == "null_chain_ctas") ctx->null_chain_ctas = (int)std::max<int64_t>(0, value);
  else if (k == "in_mux_min") ctx->in_mux_min = (int)std::max<int64_t>(1, std::min<int64_t>(value, 1 << 30));
  else if (k == "in_wide") ctx->in_wide = (int)value;
  else if (k == "in_wide_warps") ctx->in_wide_warps = (int)value;
  else if (k == "in_wide_gpt") ctx->in_wide_gpt = (int)value;
  else if (k == "in_mux_bw") ctx->in_mux_bw = (int)std::min<int64_t>(7, std::max<int64_t>(0, value));
  else if (k == "in_split_min_rows") ctx->in_split_min_rows = value;
  else if (k == "eval_narrow_bw") ctx->eval_narrow_bw = (int)value;
  else if (k == "codec_variant") ctx->codec_variant = (int)value;
  else if (k == "codec_ctas_per_sm") ctx->codec_ctas_per_sm = (int)value;
  else if (k == "eval_jit") ctx->eval_jit = (int)value;
  else if (k == "eval_jit_min_rows") ctx->eval_jit_min_rows = value;
  else if (k == "h2d_chunk_rows") ctx->h2d_chunk_rows = value;
  else if (k == "reference_compat") ctx->reference_compat = value != 0;
  else if (k == "group_timeout_ms") ctx->grp_timeout_ms = value;
  else if (k == "group_deferred") ctx->grp_deferred = value != 0;
  else return set_error(ctx, FLE_E_INVALID, "unknown option " + k);
  return FLE_OK;
}

int64_t fle_ctx_get_option(fle_ctx* ctx, const char* name) {
  if (!ctx || !name) return -1;
  const std::string k(name);
  if (k == "eval_variant") return ctx->eval_variant;
  if (k == "eval_stages") return ctx->eval_stages;
  if (k == "eval_block") return ctx->eval_block;
  if (k == "eval_ctas_per_sm") return ctx->eval_ctas_per_sm;
  if (k == "eval_gpt") return ctx->eval_gpt;
  if (k == "in_lookup_min") return ctx->in_lookup_min;
  if (k == "null_fused") return ctx->null_fused;
  if (k == "null_chain_ctas") return ctx->null_chain_ctas;
  if (k == "in_mux_min") return ctx->in_mux_min;
  if (k == "in_wide") return ctx->in_wide;
  if (k == "in_wide_warps") return ctx->in_wide_warps;
  if (k == "in_wide_gpt") return ctx->in_wide_gpt;
  if (k == "in_mux_bw") return ctx->in_mux_bw;
  if (k == "in_split_min_rows") return ctx->in_split_min_rows;
  if (k == "eval_narrow_bw") return ctx->eval_narrow_bw;
  if (k == "codec_variant") return ctx->codec_variant;
  if (k == "codec_ctas_per_sm") return ctx->codec_ctas_per_sm;
  if (k == "eval_jit") return ctx->eval_jit;
  if (k == "eval_jit_min_rows") return ctx->eval_jit_min_rows;
  if (k == "h2d_chunk_rows") return ctx->h2d_chunk_rows;
  if (k == "reference_compat") return ctx->reference_compat ? 1 : 0;
  if (k == "group_timeout_ms") return ctx->grp_timeout_ms;
  if (k == "group_deferred") return ctx->grp_deferred ? 1 : 0;
  if (k == "jit_launches") return ctx->jit_launches;
  if (k == "last_mode") return ctx->last_mode;
  if (k == "num_sms") return ctx->num_sms;
  if (k == "max_smem_optin") return ctx->max_smem_optin;
  return -1;
}

int64_t fle_ctx_launch_count(fle_ctx* ctx) { return ctx ? ctx->launches : -1; }

// ---- memory helpers ----------------------------------------------------------------------
int fle_host_alloc(size_t bytes, void** out) {
  if (!out) return FLE_E_INVALID;
  cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? FLE_E_NOMEM : FLE_E_CUDA;
  }
  return FLE_OK;
}
int fle_host_free(void* p) {
  if (p && cudaFreeHost(p) != cudaSuccess) return FLE_E_CUDA;
  return FLE_OK;
}
int fle_dev_alloc(fle_ctx* ctx, size_t bytes, void** out) {
  FLE_CHECK_CTX(ctx);
  if (!out) return FLE_E_INVALID;
  FLE_CUDA_TRY(ctx, cudaMalloc(out, bytes ? bytes : 1));
  return FLE_OK;
}
int fle_dev_free(fle_ctx* ctx, void* p) {
  FLE_CHECK_CTX(ctx);
  if (p) {
    FLE_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    FLE_CUDA_TRY(ctx, cudaFree(p));
  }
  return FLE_OK;
}
int fle_h2d(fle_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
  FLE_CHECK_CTX(ctx);
  FLE_CUDA_TRY(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return FLE_OK;
}
int fle_d2h(fle_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
  FLE_CHECK_CTX(ctx);
  FLE_CUDA_TRY(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return FLE_OK;
}

// ---- codec -------------------------------------------------------------------------------
size_t fle_encoded_bytes(int64_t n, int bit_width) {
  int64_t g = (n + 63) / 64;
  if (g < 1) g = 1;
  return (size_t)g * (size_t)bit_width * 8;
}

int fle_pack_u32_dev(fle_ctx* ctx, const uint32_t* vals_dev, int64_t n, int bw, uint8_t* enc_dev,
                     size_t enc_cap, size_t* enc_len) {
  FLE_CHECK_CTX(ctx);
  if (n < 0 || !enc_dev || (n > 0 && !vals_dev)) return set_error(ctx, FLE_E_INVALID, "bad pack arguments");
  if (bw < 1 || bw > 32) return set_error(ctx, FLE_E_INVALID, "bit width outside 1..32");
  const size_t need = fle_encoded_bytes(n, bw);
  if (enc_len) *enc_len = need;
  if (enc_cap < need) return set_error(ctx, FLE_E_BUFFER, "encode buffer too small");
  if ((reinterpret_cast<uintptr_t>(enc_dev) & 7) || (reinterpret_cast<uintptr_t>(vals_dev) & 15))
    return set_error(ctx, FLE_E_ALIGN, "pack: enc_dev must be 8-byte and vals_dev 16-byte aligned");
  return launch_pack_u32(ctx, vals_dev, n, bw, reinterpret_cast<uint64_t*>(enc_dev));
}

int fle_unpack_u32_dev(fle_ctx* ctx, const uint8_t* enc_dev, size_t enc_len, int bw,
                       int64_t first_row, int64_t n, uint32_t* out_dev) {
  FLE_CHECK_CTX(ctx);
  if (n < 0 || first_row < 0 || !enc_dev || (n > 0 && !out_dev))
    return set_error(ctx, FLE_E_INVALID, "bad unpack arguments");
  if (bw < 1 || bw > 32) return set_error(ctx, FLE_E_INVALID, "bit width outside 1..32");
  if (first_row & 63) return set_error(ctx, FLE_E_ALIGN, "unpack_dev: first_row must be a multiple of 64");
  const int64_t g0 = first_row / 64, G = (n + 63) / 64;
  if ((uint64_t)(g0 + G) * bw * 8u > enc_len) return set_error(ctx, FLE_E_BUFFER, "rows beyond the encoded buffer");
  const uint8_t* p = enc_dev + (uint64_t)g0 * bw * 8u;
  if ((reinterpret_cast<uintptr_t>(p) & 7) || (reinterpret_cast<uintptr_t>(out_dev) & 15))
    return set_error(ctx, FLE_E_ALIGN, "unpack_dev: enc must be 8-byte and out 16-byte aligned");
  return launch_unpack_u32(ctx, reinterpret_cast<const uint64_t*>(p), bw, n, out_dev);
}

static int unpack_narrow_dev(fle_ctx* ctx, const uint8_t* enc_dev, size_t enc_len, int bw, int64_t first_row, int64_t n,
                             int elem_size, void* out_dev) {
  if (n < 0 || first_row < 0 || !enc_dev || (n > 0 && !out_dev)) return set_error(ctx, FLE_E_INVALID, "bad unpack arguments");
  if (bw < 1 || bw > 8 * elem_size) return set_error(ctx, FLE_E_INVALID, "bit width does not fit the output element");
  if (first_row & 63) return set_error(ctx, FLE_E_ALIGN, "unpack_dev: first_row must be a multiple of 64");
  const int64_t g0 = first_row / 64, G = (n + 63) / 64;
  if ((uint64_t)(g0 + G) * bw * 8u > enc_len) return set_error(ctx, FLE_E_BUFFER, "rows beyond the encoded buffer");
  const uint8_t* p = enc_dev + (uint64_t)g0 * bw * 8u;
  if ((reinterpret_cast<uintptr_t>(p) & 15) || (reinterpret_cast<uintptr_t>(out_dev) & 15))
    return set_error(ctx, FLE_E_ALIGN, "narrow unpack: enc and out must be 16-byte aligned");
  return launch_unpack_narrow(ctx, reinterpret_cast<const uint64_t*>(p), bw, n, elem_size, out_dev);
}

int fle_unpack_u8_dev(fle_ctx* ctx, const uint8_t* enc_dev, size_t enc_len, int bw, int64_t first_row, int64_t n,
                      uint8_t* out_dev) {
  FLE_CHECK_CTX(ctx);
  return unpack_narrow_dev(ctx, enc_dev, enc_len, bw, first_row, n, 1, out_dev);
}

int fle_unpack_u16_dev(fle_ctx* ctx, const uint8_t* enc_dev, size_t enc_len, int bw, int64_t first_row, int64_t n,
                       uint16_t* out_dev) {
  FLE_CHECK_CTX(ctx);
  return unpack_narrow_dev(ctx, enc_dev, enc_len, bw, first_row, n, 2, out_dev);
}

// Host variants: values/planes cross PCIe in chunks through the context's two staging buffers.
static const int64_t kCodecChunkValues = 32ll << 20;  // 32 Mi values (128 MiB of u32) per chunk

int fle_pack_u32(fle_ctx* ctx, const uint32_t* vals, int64_t n, int bw, uint8_t* enc, size_t enc_cap,
                 size_t* enc_len) {
  FLE_CHECK_CTX(ctx);
  if (n < 0 || !enc || (n > 0 && !vals)) return set_error(ctx, FLE_E_INVALID, "bad pack arguments");
  if (bw < 1 || bw > 32) return set_error(ctx, FLE_E_INVALID, "bit width outside 1..32");
  const size_t need = fle_encoded_bytes(n, bw);
  if (enc_len) *enc_len = need;
  if (enc_cap < need) return set_error(ctx, FLE_E_BUFFER, "encode buffer too small");
  if (n == 0) {
    std::memset(enc, 0, need);
    return FLE_OK;
  }
  const int64_t chunk = std::min<int64_t>(n, kCodecChunkValues);
  int rc = ensure_stage(ctx, 0, (size_t)chunk * 4);
  if (rc != FLE_OK) return rc;
  rc = ensure_stage(ctx, 1, fle_encoded_bytes(chunk, bw));
  if (rc != FLE_OK) return rc;
  for (int64_t off = 0; off < n; off += chunk) {
    const int64_t m = std::min(chunk, n - off);
    const size_t eb = fle_encoded_bytes(m, bw);
    FLE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_stage[0], vals + off, (size_t)m * 4, cudaMemcpyHostToDevice, ctx->stream));
    rc = launch_pack_u32(ctx, (const uint32_t*)ctx->d_stage[0], m, bw, (uint64_t*)ctx->d_stage[1]);
    if (rc != FLE_OK) return rc;
    FLE_CUDA_TRY(ctx, cudaMemcpyAsync(enc + (size_t)(off / 64) * bw * 8, ctx->d_stage[1], eb, cudaMemcpyDeviceToHost, ctx->stream));
    FLE_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return FLE_OK;
}

int fle_unpack_u32(fle_ctx* ctx, const uint8_t* enc, size_t enc_len, int bw, int64_t first_row,
                   int64_t n, uint32_t* out) {
  FLE_CHECK_CTX(ctx);
  if (n < 0 || first_row < 0 || !enc || (n > 0 && !out)) return set_error(ctx, FLE_E_INVALID, "bad unpack arguments");
  if (bw < 1 || bw > 32) return set_error(ctx, FLE_E_INVALID, "bit width outside 1..32");
  if (n == 0) return FLE_OK;
  const int64_t g_first = first_row / 64;
  const int64_t g_end = (first_row + n + 63) / 64;
  if ((uint64_t)g_end * bw * 8u > enc_len) return set_error(ctx, FLE_E_BUFFER, "rows beyond the encoded buffer");
  const int64_t chunk_groups = kCodecChunkValues / 64;
  int rc = ensure_stage(ctx, 0, (size_t)std::min(chunk_groups, g_end - g_first) * bw * 8);
  if (rc != FLE_OK) return rc;
  rc = ensure_stage(ctx, 1, (size_t)std::min(chunk_groups, g_end - g_first) * 256);
  if (rc != FLE_OK) return rc;
  for (int64_t g = g_first; g < g_end; g += chunk_groups) {
    const int64_t ng = std::min(chunk_groups, g_end - g);
    FLE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_stage[0], enc + (size_t)g * bw * 8, (size_t)ng * bw * 8, cudaMemcpyHostToDevice, ctx->stream));
    rc = launch_unpack_u32(ctx, (const uint64_t*)ctx->d_stage[0], bw, ng * 64, (uint32_t*)ctx->d_stage[1]);
    if (rc != FLE_OK) return rc;
    // rows of this chunk that the caller asked for
    const int64_t r0 = std::max(first_row, g * 64), r1 = std::min(first_row + n, (g + ng) * 64);
    FLE_CUDA_TRY(ctx, cudaMemcpyAsync(out + (r0 - first_row), (const uint32_t*)ctx->d_stage[1] + (r0 - g * 64),
                                      (size_t)(r1 - r0) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    FLE_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return FLE_OK;
}

// ---- fused predicate evaluation ------------------------------------------------------------
int fle_eval_dev(fle_ctx* ctx, const fle_col_desc* cols, int ncols, const fle_pred_node* nodes,
                 int nnodes, int64_t first_row, int64_t n, uint64_t* bitmap_dev, uint64_t* count_dev) {
  FLE_CHECK_CTX(ctx);
  if (!cols || !nodes || ncols < 1 || nnodes < 1) return set_error(ctx, FLE_E_INVALID, "empty program");
  if (ncols > FLE_MAX_COLS) return set_error(ctx, FLE_E_UNSUPPORTED, "too many columns");
  if (any_nullable(cols, ncols))
    return eval_dev_nullable(ctx, cols, ncols, nodes, nnodes, first_row, n, bitmap_dev, count_dev);
  return launch_eval(ctx, cols, ncols, nodes, nnodes, first_row, n, bitmap_dev, count_dev);
}

// OPTIONAL columns without a host synchronisation: every Impala-written file marks its columns OPTIONAL
// (hdfs-parquet-table-writer.cc:103-105) and TPC-H holds no NULL, so the def levels only have to be CHECKED.
int fle_eval_no_nulls_dev(fle_ctx* ctx, const fle_col_desc* cols, int ncols, const fle_pred_node* nodes, int nnodes,
                          int64_t first_row, int64_t n, uint64_t* bitmap_dev, uint64_t* count_dev, uint64_t* null_rows_dev) {
  FLE_CHECK_CTX(ctx);
  if (!cols || !nodes || ncols < 1 || nnodes < 1 || !null_rows_dev) return set_error(ctx, FLE_E_INVALID, "bad arguments");
  if (ncols > FLE_MAX_COLS) return set_error(ctx, FLE_E_UNSUPPORTED, "too many columns");
  if (first_row != 0 && any_nullable(cols, ncols))
    return set_error(ctx, FLE_E_UNSUPPORTED, "nullable columns are evaluated from row 0 of the page");
  std::vector<fle_col_desc> req(cols, cols + ncols);
  for (auto& d : req) {
    if (d.max_def_level > 1) return set_error(ctx, FLE_E_UNSUPPORTED, "max_def_level > 1 (nested types)");
    d.max_def_level = 0;
    d.def_planes = nullptr;
    d.def_bytes = 0;
  }
  int rc = FLE_OK;
  if (any_nullable(cols, ncols)) {
    rc = launch_def_counts(ctx, cols, ncols, n, nullptr, null_rows_dev);
  } else {
    FLE_CUDA_TRY(ctx, cudaMemsetAsync(null_rows_dev, 0, 8, ctx->stream));
  }
  if (rc != FLE_OK) return rc;
  return launch_eval(ctx, req.data(), ncols, nodes, nnodes, first_row, n, bitmap_dev, count_dev);
}

int fle_eval_codegen(const fle_col_desc* cols, int ncols, const fle_pred_node* nodes, int nnodes, int compile,
                     char* out, size_t cap, size_t* len) {
  if (len) *len = 0;
  if (out && cap) out[0] = '\0';
  LoweredProgram lp;
  std::string text;
  int rc = lower_program(cols, ncols, nodes, nnodes, &lp, &text);
  if (rc == FLE_OK) {
    lp.p.ncols = ncols;
    for (int c = 0; c < ncols; ++c) lp.p.bw[c] = (uint32_t)cols[c].bit_width;
    text.clear();
    const int agg_col = ((compile >> 8) & 0xff) - 1;
    const bool skew = ((compile >> 16) & 1) != 0;
    compile &= 0xff;
    if (!jit_plan(lp.p, lp.in_lits, &lp.jit_plan, agg_col, skew)) {
      rc = FLE_E_UNSUPPORTED;
      text = "program too large for a specialised kernel";
    } else if (compile) {
      std::vector<char> cubin;
      rc = jit_compile_text(lp.jit_plan.text, &cubin, &text);
      if (rc == FLE_OK && compile == 2) text.assign(cubin.begin(), cubin.end());
    } else {
      text = lp.jit_plan.text;
    }
  }
  if (len) *len = text.size();
  if (out && cap) {
    const size_t m = std::min(cap - 1, text.size());
    std::memcpy(out, text.data(), m);
    out[m] = '\0';
  }
  return rc;
}

int fle_eval(fle_ctx* ctx, const fle_col_desc* cols, int ncols, const fle_pred_node* nodes, int nnodes,
             int64_t first_row, int64_t n, uint64_t* bitmap_out, uint64_t* match_count) {
  FLE_CHECK_CTX(ctx);
  if (!cols || !nodes || ncols < 1 || nnodes < 1 || n < 0 || first_row < 0 || (n > 0 && !bitmap_out))
    return set_error(ctx, FLE_E_INVALID, "bad eval arguments");
  if (ncols > FLE_MAX_COLS) return set_error(ctx, FLE_E_UNSUPPORTED, "too many columns");
  if (first_row & 63) return set_error(ctx, FLE_E_ALIGN, "first_row must be a multiple of 64");
  if (match_count) *match_count = 0;
  if (n == 0) {
    LoweredProgram lp;
    std::string err;
    int rc0 = lower_program(cols, ncols, nodes, nnodes, &lp, &err, ctx->reference_compat);  // still validate the program
    return rc0 == FLE_OK ? FLE_OK : set_error(ctx, rc0, err);
  }
  const int64_t G = (n + 63) / 64, g0 = first_row / 64;
  int rc;

  if (any_nullable(cols, ncols)) {
    // upload everything, then run the def-level path on the device copies
    if (first_row != 0) return set_error(ctx, FLE_E_UNSUPPORTED, "nullable columns are evaluated from row 0 of the page");
    size_t total = 0;
    std::vector<size_t> off_p(ncols), off_d(ncols);
    for (int c = 0; c < ncols; ++c) {
      if (!cols[c].planes && cols[c].bytes) return set_error(ctx, FLE_E_INVALID, "null column pointer");
      off_p[c] = total; total = align_up(total + cols[c].bytes, 256);
      off_d[c] = total; total = align_up(total + (cols[c].max_def_level ? cols[c].def_bytes : 0), 256);
    }
    const size_t off_bm = total; total += (size_t)G * 8;
    rc = ensure_stage(ctx, 0, total + 256);
    if (rc != FLE_OK) return rc;
    uint8_t* base = (uint8_t*)ctx->d_stage[0];
    std::vector<fle_col_desc> dc(cols, cols + ncols);
    for (int c = 0; c < ncols; ++c) {
      if (cols[c].bytes)
        FLE_CUDA_TRY(ctx, cudaMemcpyAsync(base + off_p[c], cols[c].planes, cols[c].bytes, cudaMemcpyHostToDevice, ctx->stream));
      dc[c].planes = base + off_p[c];
      if (cols[c].max_def_level) {
        if (!cols[c].def_planes) return set_error(ctx, FLE_E_INVALID, "nullable column without def_planes");
        FLE_CUDA_TRY(ctx, cudaMemcpyAsync(base + off_d[c], cols[c].def_planes, cols[c].def_bytes, cudaMemcpyHostToDevice, ctx->stream));
        dc[c].def_planes = base + off_d[c];
      }
    }
    rc = eval_dev_nullable(ctx, dc.data(), ncols, nodes, nnodes, 0, n, (uint64_t*)(base + off_bm), ctx->d_count);
    if (rc != FLE_OK) return rc;
    FLE_CUDA_TRY(ctx, cudaMemcpyAsync(bitmap_out, base + off_bm, (size_t)G * 8, cudaMemcpyDeviceToHost, ctx->stream));
    FLE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_count, ctx->d_count, 8, cudaMemcpyDeviceToHost, ctx->stream));
    FLE_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (match_count) *match_count = *ctx->h_count;
    return FLE_OK;
  }

  // required columns: chunked, double-buffered H2D -> kernel -> D2H
  int sum_bw = 0;
  for (int c = 0; c < ncols; ++c) {
    const int bw = cols[c].bit_width;
    if (bw < 1 || bw > 32) return set_error(ctx, FLE_E_INVALID, "bit width outside 1..32");
    if (!cols[c].planes) return set_error(ctx, FLE_E_INVALID, "null column pointer");
    if (cols[c].bytes < (uint64_t)(g0 + G) * bw * 8u) return set_error(ctx, FLE_E_BUFFER, "column shorter than the rows requested");
    sum_bw += bw;
  }
  LoweredProgram lp;
  rc = prepare_eval(ctx, cols, ncols, nodes, nnodes, &lp);
  if (rc != FLE_OK) return rc;

  int64_t chunk_groups = ctx->h2d_chunk_rows > 0 ? ctx->h2d_chunk_rows / 64
                                                 : (int64_t)(48ull << 20) / ((int64_t)sum_bw * 8);
  chunk_groups = std::max<int64_t>(2048, chunk_groups / 2048 * 2048);
  chunk_groups = std::min<int64_t>(chunk_groups, (G + 2047) / 2048 * 2048);
  const int64_t nchunks = (G + chunk_groups - 1) / chunk_groups;
  std::vector<size_t> coff(ncols);
  size_t total = 0;
  for (int c = 0; c < ncols; ++c) {
    coff[c] = total;
    total = align_up(total + (size_t)chunk_groups * cols[c].bit_width * 8, 256);
  }
  const size_t off_bm = total;
  total = align_up(total + (size_t)chunk_groups * 8, 256);
  for (int b = 0; b < 2; ++b) {
    if (b == 1 && nchunks == 1) break;
    rc = ensure_stage(ctx, b, total);
    if (rc != FLE_OK) return rc;
  }
  // per-chunk counts: grow-only context scratch (no cudaMalloc / cudaFree -- a device-wide sync -- per call)
  if (ctx->chunk_counts_cap < (size_t)nchunks) {
    FLE_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->d_chunk_counts) cudaFree(ctx->d_chunk_counts);
    ctx->d_chunk_counts = nullptr;
    ctx->chunk_counts_cap = 0;
    const size_t cap = std::max<size_t>(256, (size_t)nchunks * 2);
    FLE_CUDA_TRY(ctx, cudaMalloc(&ctx->d_chunk_counts, cap * 8));
    ctx->chunk_counts_cap = cap;
  }
  uint64_t* d_counts = ctx->d_chunk_counts;
  std::vector<uint64_t> h_counts(nchunks, 0);
  rc = FLE_OK;
  // the staging buffers may still be read by work queued earlier on the context's stream (the aggregate
  // fallback keeps its bitmap in d_stage[0]): the copy stream starts behind it
  cudaError_t e = cudaEventRecord(ctx->ev_done[0], ctx->stream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[0], 0);
  for (int64_t k = 0; k < nchunks && rc == FLE_OK && e == cudaSuccess; ++k) {
    const int b = (int)(k & 1);
    const int64_t cg0 = k * chunk_groups;
    const int64_t cg = std::min(chunk_groups, G - cg0);
    const int64_t rows = std::min<int64_t>(cg * 64, n - cg0 * 64);
    uint8_t* base = (uint8_t*)ctx->d_stage[b];
    if (k >= 2) e = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[b], 0);
    std::vector<fle_col_desc> dc(cols, cols + ncols);
    for (int c = 0; c < ncols && e == cudaSuccess; ++c) {
      const int bw = cols[c].bit_width;
      const uint8_t* src = (const uint8_t*)cols[c].planes + (size_t)(g0 + cg0) * bw * 8;
      e = cudaMemcpyAsync(base + coff[c], src, (size_t)cg * bw * 8, cudaMemcpyHostToDevice, ctx->copy_stream);
      dc[c].planes = base + coff[c];
      dc[c].bytes = (uint64_t)cg * bw * 8;
    }
    if (e != cudaSuccess) break;
    e = cudaEventRecord(ctx->ev_copy[b], ctx->copy_stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[b], 0);
    if (e != cudaSuccess) break;
    rc = launch_prepared(ctx, &lp, dc.data(), ncols, 0, rows, (uint64_t*)(base + off_bm), d_counts + k);
    if (rc != FLE_OK) break;
    e = cudaMemcpyAsync(bitmap_out + cg0, base + off_bm, (size_t)cg * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_done[b], ctx->stream);
  }
  if (rc == FLE_OK && e == cudaSuccess)
    e = cudaMemcpyAsync(h_counts.data(), d_counts, (size_t)nchunks * 8, cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e2 = cudaStreamSynchronize(ctx->copy_stream);
  cudaError_t e3 = cudaStreamSynchronize(ctx->stream);
  if (rc != FLE_OK) return rc;
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "fle_eval pipeline");
  if (e2 != cudaSuccess) return set_cuda_error(ctx, e2, "fle_eval copy stream");
  if (e3 != cudaSuccess) return set_cuda_error(ctx, e3, "fle_eval stream");
  if (match_count) {
    uint64_t t = 0;
    for (uint64_t c : h_counts) t += c;
    *match_count = t;
  }
  return FLE_OK;
}

// ---- bitmap utilities ----------------------------------------------------------------------
int fle_bitmap_count_dev(fle_ctx* ctx, const uint64_t* bitmap_dev, int64_t nbits, uint64_t* count_dev) {
  FLE_CHECK_CTX(ctx);
  if (!bitmap_dev && nbits > 0) return set_error(ctx, FLE_E_INVALID, "null bitmap");
  return launch_bitmap_count(ctx, bitmap_dev, nbits, count_dev);
}

int fle_intersect_bitset_dev(fle_ctx* ctx, uint64_t* root_dev, int64_t n_root, const uint64_t* sub_dev,
                             int64_t n_sub) {
  FLE_CHECK_CTX(ctx);
  return launch_intersect(ctx, root_dev, n_root, sub_dev, n_sub);
}

int fle_intersect_bitset(fle_ctx* ctx, uint64_t* root, int64_t n_root, const uint64_t* sub, int64_t n_sub) {
  FLE_CHECK_CTX(ctx);
  if (n_root < 0 || n_sub < 0 || (n_root > 0 && !root) || (n_sub > 0 && !sub))
    return set_error(ctx, FLE_E_INVALID, "bad intersect arguments");
  if (n_root == 0) return FLE_OK;
  const size_t rb = (size_t)((n_root + 63) / 64) * 8, sb = (size_t)((n_sub + 63) / 64) * 8;
  int rc = ensure_stage(ctx, 0, rb);
  if (rc != FLE_OK) return rc;
  rc = ensure_stage(ctx, 1, std::max<size_t>(sb, 8));
  if (rc != FLE_OK) return rc;
  FLE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_stage[0], root, rb, cudaMemcpyHostToDevice, ctx->stream));
  if (sb) FLE_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_stage[1], sub, sb, cudaMemcpyHostToDevice, ctx->stream));
  rc = launch_intersect(ctx, (uint64_t*)ctx->d_stage[0], n_root, (const uint64_t*)ctx->d_stage[1], n_sub);
  if (rc != FLE_OK) return rc;
  FLE_CUDA_TRY(ctx, cudaMemcpyAsync(root, ctx->d_stage[0], rb, cudaMemcpyDeviceToHost, ctx->stream));
  FLE_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return FLE_OK;
}

// ---- survivor materialisation --------------------------------------------------------------
int fle_bitmap_select_dev(fle_ctx* ctx, const uint64_t* bitmap_dev, int64_t nbits, uint32_t* row_ids_dev,
                          int64_t cap, uint64_t* count_dev) {
  FLE_CHECK_CTX(ctx);
  return launch_select(ctx, bitmap_dev, nbits, row_ids_dev, cap, count_dev);
}

int fle_gather_dev(fle_ctx* ctx, const uint8_t* enc_dev, size_t enc_len, int bit_width,
                   const uint32_t* row_ids_dev, const uint64_t* count_dev, int64_t cap, const void* dict_dev,
                   int ndict, int elem_size, void* out_dev) {
  FLE_CHECK_CTX(ctx);
  (void)enc_len;
  if (reinterpret_cast<uintptr_t>(enc_dev) & 7) return set_error(ctx, FLE_E_ALIGN, "enc_dev must be 8-byte aligned");
  return launch_gather(ctx, reinterpret_cast<const uint64_t*>(enc_dev), bit_width, row_ids_dev, count_dev, cap,
                       dict_dev, ndict, elem_size, out_dev);
}

int fle_sum_product_dev(fle_ctx* ctx, const uint64_t* bitmap_dev, int64_t nrows, const int64_t* plain_i64_dev,
                        const uint8_t* enc_dev, size_t enc_len, int bit_width, const int64_t* dict_i64_dev,
                        int ndict, int64_t* sum_dev, uint64_t* count_dev) {
  FLE_CHECK_CTX(ctx);
  if (bit_width < 1 || bit_width > 32) return set_error(ctx, FLE_E_INVALID, "bit width outside 1..32");
  if ((uint64_t)((nrows + 63) / 64) * bit_width * 8u > enc_len) return set_error(ctx, FLE_E_BUFFER, "rows beyond the encoded buffer");
  if (reinterpret_cast<uintptr_t>(enc_dev) & 7) return set_error(ctx, FLE_E_ALIGN, "enc_dev must be 8-byte aligned");
  return launch_sum_product(ctx, bitmap_dev, nrows, plain_i64_dev, reinterpret_cast<const uint64_t*>(enc_dev),
                            bit_width, dict_i64_dev, ndict, sum_dev, count_dev);
}

int fle_eval_sum_product_dev(fle_ctx* ctx, const fle_col_desc* cols, int ncols, const fle_pred_node* nodes, int nnodes,
                             int64_t first_row, int64_t n, const int64_t* plain_i64_dev, int code_col,
                             const int64_t* dict_i64_dev, int ndict, uint64_t* bitmap_dev, int64_t* sum_dev,
                             uint64_t* count_dev) {
  FLE_CHECK_CTX(ctx);
  if (!cols || !nodes || ncols < 1 || nnodes < 1) return set_error(ctx, FLE_E_INVALID, "empty program");
  if (ncols > FLE_MAX_COLS) return set_error(ctx, FLE_E_UNSUPPORTED, "too many columns");
  if (code_col < 0 || code_col >= ncols || !plain_i64_dev || !dict_i64_dev || !sum_dev || ndict < 0)
    return set_error(ctx, FLE_E_INVALID, "bad aggregate arguments");
  if (any_nullable(cols, ncols)) return set_error(ctx, FLE_E_UNSUPPORTED, "fused aggregate over nullable columns");
  LoweredProgram lp;
  int rc = prepare_eval(ctx, cols, ncols, nodes, nnodes, &lp);
  if (rc != FLE_OK) return rc;
  const AggSpec agg{plain_i64_dev, code_col, dict_i64_dev, ndict, sum_dev};
  rc = launch_prepared(ctx, &lp, cols, ncols, first_row, n, bitmap_dev, count_dev, &agg);
  if (rc != FLE_E_UNSUPPORTED || ctx->eval_jit == 2) return rc;
  // no specialised kernel to be had: predicate pass, then the aggregate over its bitmap
  const int64_t G = (n + 63) / 64;
  if (!bitmap_dev) {
    rc = ensure_stage(ctx, 0, (size_t)std::max<int64_t>(G, 1) * 8);
    if (rc != FLE_OK) return rc;
    bitmap_dev = static_cast<uint64_t*>(ctx->d_stage[0]);
  }
  rc = launch_prepared(ctx, &lp, cols, ncols, first_row, n, bitmap_dev, nullptr);
  if (rc != FLE_OK) return rc;
  const int bw = cols[code_col].bit_width;
  const uint8_t* planes = reinterpret_cast<const uint8_t*>(cols[code_col].planes) + (uint64_t)(first_row / 64) * bw * 8u;
  return launch_sum_product(ctx, bitmap_dev, n, plain_i64_dev, reinterpret_cast<const uint64_t*>(planes), bw,
                            dict_i64_dev, ndict, sum_dev, count_dev);
}

// ---- row-group batches ---------------------------------------------------------------------------
int fle_h2d_prefetch(fle_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
  FLE_CHECK_CTX(ctx);
  if (bytes == 0) return FLE_OK;
  if (!dst_dev || !src_host) return set_error(ctx, FLE_E_INVALID, "null pointer");
  fle_ctx* c = ctx->copy_owner ? ctx->copy_owner : ctx;
  FLE_CUDA_TRY(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, c->copy_stream));
  return FLE_OK;
}

int fle_ctx_share_copy_stream(fle_ctx* ctx, fle_ctx* owner) {
  FLE_CHECK_CTX(ctx);
  if (owner == ctx) owner = nullptr;
  if (owner && (owner->device != ctx->device || owner->copy_owner)) return set_error(ctx, FLE_E_INVALID, "the copy stream's owner must be a context of the same device with a copy stream of its own");
  ctx->copy_owner = owner;
  return FLE_OK;
}

int fle_prefetch_fence(fle_ctx* ctx) {
  FLE_CHECK_CTX(ctx);
  if (!ctx->ev_prefetch) FLE_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_prefetch, cudaEventDisableTiming));
  FLE_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_prefetch, (ctx->copy_owner ? ctx->copy_owner : ctx)->copy_stream));
  FLE_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_prefetch, 0));
  return FLE_OK;
}

int fle_prefetch_mark(fle_ctx* ctx, int slot) {
  FLE_CHECK_CTX(ctx);
  if (slot < 0 || slot >= fle_ctx::kPrefetchMarks) return set_error(ctx, FLE_E_INVALID, "prefetch mark outside 0..7");
  fle_ctx* c = ctx->copy_owner ? ctx->copy_owner : ctx;
  if (!c->ev_mark[slot]) FLE_CUDA_TRY(ctx, cudaEventCreateWithFlags(&c->ev_mark[slot], cudaEventDisableTiming));
  FLE_CUDA_TRY(ctx, cudaEventRecord(c->ev_mark[slot], c->copy_stream));
  return FLE_OK;
}

int fle_prefetch_wait(fle_ctx* ctx, int slot) {
  FLE_CHECK_CTX(ctx);
  if (slot < 0 || slot >= fle_ctx::kPrefetchMarks) return set_error(ctx, FLE_E_INVALID, "prefetch mark outside 0..7");
  fle_ctx* c = ctx->copy_owner ? ctx->copy_owner : ctx;
  if (!c->ev_mark[slot]) return set_error(ctx, FLE_E_INVALID, "fle_prefetch_wait before fle_prefetch_mark");
  FLE_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, c->ev_mark[slot], 0));
  return FLE_OK;
}

int fle_host_register(void* p, size_t bytes) {
  if (!p || bytes == 0) return FLE_E_INVALID;
  cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? FLE_E_NOMEM : FLE_E_CUDA;
  }
  return FLE_OK;
}

int fle_host_unregister(void* p) {
  if (p && cudaHostUnregister(p) != cudaSuccess) {
    cudaGetLastError();
    return FLE_E_CUDA;
  }
  return FLE_OK;
}

int fle_pages_gather_dev(fle_ctx* ctx, const fle_page_ref* refs, int nrefs, uint64_t* null_rows_dev) {
  FLE_CHECK_CTX(ctx);
  return launch_pages_gather(ctx, refs, nrefs, null_rows_dev);
}

int fle_assemble_rows_dev(fle_ctx* ctx, const uint64_t* bitmap_dev, int64_t nrows, const fle_slot_desc* cols, int ncols,
                          int tuple_size, uint8_t* tuples_dev, int64_t cap_rows, uint64_t* count_dev,
                          uint64_t* bad_codes_dev) {
  FLE_CHECK_CTX(ctx);
  return launch_assemble(ctx, bitmap_dev, nrows, cols, ncols, tuple_size, tuples_dev, cap_rows, count_dev, bad_codes_dev);
}

// ---- groups of GPUs: totals summed inside the scan kernel ---------------------------------------
// The fused exchange attaches to the ONE kernel that publishes the call's totals (group_fill looks at
// grp_on and at the count pointer); paths that end in another kernel (nullable columns, long IN
// lists regrouped around the lookup kernel, the two-kernel aggregate, empty inputs) are followed by
// the one-warp reduce kernel instead.  Either way a call consumes exactly one sequence number.
int fle_eval_group_dev(fle_ctx* ctx, const fle_col_desc* cols, int ncols, const fle_pred_node* nodes, int nnodes,
                       int64_t first_row, int64_t n, uint64_t* bitmap_dev, uint64_t* total_count_dev) {
  FLE_CHECK_CTX(ctx);
  if (!total_count_dev) return set_error(ctx, FLE_E_INVALID, "group evaluation needs a count slot");
  if (ctx->grp_world < 1) return set_error(ctx, FLE_E_INVALID, "context does not belong to a group");
  if (!cols || !nodes || ncols < 1 || nnodes < 1) return set_error(ctx, FLE_E_INVALID, "empty program");
  if (ncols > FLE_MAX_COLS) return set_error(ctx, FLE_E_UNSUPPORTED, "too many columns");
  if (any_nullable(cols, ncols)) {
    int rc = eval_dev_nullable(ctx, cols, ncols, nodes, nnodes, first_row, n, bitmap_dev, total_count_dev);
    if (rc != FLE_OK) return rc;
    return launch_group_reduce(ctx, total_count_dev, nullptr);
  }
  // single-kernel path: lower + launch with the exchange fused; everything else: trailing reduce
  const int64_t launches0 = ctx->launches;
  LoweredProgram lp;
  int rc = FLE_E_UNSUPPORTED;
  bool fused = false;
  if (n > 0) {
    // (long IN lists inside larger programs keep their regrouped multi-kernel form)
    std::string err;
    bool long_in = false;
    for (int i = 0; i < nnodes && nnodes > 1; ++i)
      long_in = long_in || (nodes[i].op == FLE_OP_IN && nodes[i].ncodes >= 16);
    if (!long_in) {
      rc = prepare_eval(ctx, cols, ncols, nodes, nnodes, &lp);
      if (rc != FLE_OK) return rc;
      ctx->grp_on = true;
      rc = launch_prepared(ctx, &lp, cols, ncols, first_row, n, bitmap_dev, total_count_dev);
      ctx->grp_on = false;
      if (rc != FLE_OK) return rc;
      fused = ctx->launches == launches0 + 1;
    }
  }
  if (fused) return FLE_OK;
  if (rc == FLE_E_UNSUPPORTED) {
    rc = launch_eval(ctx, cols, ncols, nodes, nnodes, first_row, n, bitmap_dev, total_count_dev);
    if (rc != FLE_OK) return rc;
  }
  return launch_group_reduce(ctx, total_count_dev, nullptr);
}

int fle_eval_sum_product_group_dev(fle_ctx* ctx, const fle_col_desc* cols, int ncols, const fle_pred_node* nodes,
                                   int nnodes, int64_t first_row, int64_t n, const int64_t* plain_i64_dev, int code_col,
                                   const int64_t* dict_i64_dev, int ndict, uint64_t* bitmap_dev, int64_t* total_sum_dev,
                                   uint64_t* total_count_dev) {
  FLE_CHECK_CTX(ctx);
  if (!total_count_dev || !total_sum_dev) return set_error(ctx, FLE_E_INVALID, "group aggregate needs sum and count slots");
  if (ctx->grp_world < 1) return set_error(ctx, FLE_E_INVALID, "context does not belong to a group");
  const int64_t launches0 = ctx->launches;
  ctx->grp_on = n > 0;
  int rc = fle_eval_sum_product_dev(ctx, cols, ncols, nodes, nnodes, first_row, n, plain_i64_dev, code_col, dict_i64_dev,
                                    ndict, bitmap_dev, total_sum_dev, total_count_dev);
  const bool was_on = ctx->grp_on;
  ctx->grp_on = false;
  if (rc != FLE_OK) return rc;
  // one launch with the group attached = the fused kernel did the exchange; the two-kernel fallback
  // (predicate kernel with no count slot, then the aggregate kernel) did not
  if (was_on && ctx->launches == launches0 + 1) return FLE_OK;
  return launch_group_reduce(ctx, total_count_dev, reinterpret_cast<uint64_t*>(total_sum_dev));
}

// ---- page framing / PLAIN columns ------------------------------------------------------------
int fle_page_parse(const uint8_t* page, size_t size, int max_def_level, fle_page_layout* out) {
  if (!page || !out || max_def_level < 0) return FLE_E_INVALID;
  std::memset(out, 0, sizeof(*out));
  size_t pos = 0;
  if (max_def_level > 0) {  // hdfs-parquet-scanner.cc:882-919
    if (size < 4) return FLE_E_BUFFER;
    int32_t nd = 0;
    std::memcpy(&nd, page, 4);
    if (nd <= 0 || (size_t)nd > size - 4) return FLE_E_BUFFER;
    out->def_off = 4;
    out->def_bytes = (uint32_t)nd;
    pos = 4 + (size_t)nd;
  }
  if (pos == size) return FLE_OK;  // all-NULL page: no value section (:921)
  out->bit_width = page[pos];      // dict-encoding.h:185-192
  out->data_off = (uint32_t)(pos + 1);
  out->data_bytes = (uint32_t)(size - pos - 1);
  return FLE_OK;
}

int fle_realign_dev(fle_ctx* ctx, const uint8_t* src_dev, size_t nbytes, uint8_t* dst_dev) {
  FLE_CHECK_CTX(ctx);
  return launch_realign(ctx, src_dev, nbytes, dst_dev);
}

int fle_plain_eval_dev(fle_ctx* ctx, const void* values_dev, int elem_type, int64_t n, int op, const void* literals,
                       int nlits, int reference_operand_order, uint64_t* bitmap_out_dev, uint64_t* match_count_dev) {
  FLE_CHECK_CTX(ctx);
  return launch_plain_eval(ctx, values_dev, elem_type, n, op, literals, nlits, reference_operand_order,
                           bitmap_out_dev, match_count_dev);
}

int fle_dict_encode_dev(fle_ctx* ctx, const void* values_dev, int elem_type, int64_t n, int max_entries,
                        int64_t* dict_out_dev, int* num_entries, int* bit_width, uint8_t* enc_dev, size_t enc_cap,
                        size_t* enc_len) {
  FLE_CHECK_CTX(ctx);
  return launch_dict_encode(ctx, values_dev, elem_type, n, max_entries, dict_out_dev, num_entries, bit_width, enc_dev,
                            enc_cap, enc_len);
}

}  // extern "C"
