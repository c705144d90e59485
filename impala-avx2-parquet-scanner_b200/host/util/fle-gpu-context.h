// Per-thread handle on libfle_b200.so for the host-side drop-ins.
//
// The reference runs one HdfsParquetScanner (with its column readers and FLE decoders) per
// scanner thread (hdfs-parquet-scanner.cc:1040, expr-context.h:39-42); accordingly every
// thread that touches an FleDecoder/FleEncoder gets its own fle_ctx (one CUDA stream), created
// on first use and destroyed at thread exit.  Device selection: FLE_B200_DEVICE (default 0).
//
// Error convention: the reference's predicate functions return void and cannot fail; here a
// failed C-ABI call is recorded in a thread-local status that the scanner facade turns into
// parse_status_ (and Get/Put/Skip turn into `false`).  There is no CPU fallback.
#ifndef FLE_B200_HOST_GPU_CONTEXT_H
#define FLE_B200_HOST_GPU_CONTEXT_H

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "fle_b200.h"

namespace impala {

class FleGpu {
 public:
  // The calling thread's context, or NULL when it cannot be created (status() says why).
  static fle_ctx* ctx() {
    Holder& h = holder();
    if (!h.tried) {
      h.tried = true;
      const char* dev = std::getenv("FLE_B200_DEVICE");
      int rc = fle_ctx_create(dev ? std::atoi(dev) : 0, NULL, &h.ctx);
      if (rc != FLE_OK) Fail(rc, "fle_ctx_create");
      // a scanner thread never waits for the compiler: program shapes are specialised on a thread of the library's
      // own and the interpreting kernels serve the row groups until the kernel is there (FLE_B200_EVAL_JIT overrides)
      if (h.ctx) {
        const char* jit = std::getenv("FLE_B200_EVAL_JIT");
        fle_ctx_set_option(h.ctx, "eval_jit", jit ? std::atoi(jit) : 3);
      }
    }
    const int want = reference_compat() ? 1 : 0;
    if (h.ctx && h.compat_applied != want) {
      fle_ctx_set_option(h.ctx, "reference_compat", want);
      h.compat_applied = want;
    }
    return h.ctx;
  }

  // A scanner thread keeps two row groups in flight (one uploading and being evaluated while the other one's tuples
  // are fetched): each of the two scratch banks has a context -- a stream pair -- of its own, so that the work of
  // one row group never queues behind the other's.  Bank 0 is the thread's main context.
  static fle_ctx* bank_ctx(int bank) {
    if ((bank & 1) == 0) return ctx();
    Holder& h = holder();
    if (ctx() == NULL) return NULL;
    if (!h.tried2) {
      h.tried2 = true;
      const char* dev = std::getenv("FLE_B200_DEVICE");
      int rc = fle_ctx_create(dev ? std::atoi(dev) : 0, NULL, &h.ctx2);
      if (rc != FLE_OK) { h.ctx2 = NULL; Fail(rc, "fle_ctx_create (second bank)"); }
      if (h.ctx2) {
        const char* jit = std::getenv("FLE_B200_EVAL_JIT");
        fle_ctx_set_option(h.ctx2, "eval_jit", jit ? std::atoi(jit) : 3);
        // one copy stream for both banks: the uploads cross PCIe one after the other, in the order they are needed
        fle_ctx_share_copy_stream(h.ctx2, h.ctx);
      }
    }
    return h.ctx2;
  }

  // Process-wide switch (also FLE_B200_REFERENCE_COMPAT=1): reproduce the reference EXACTLY where the
  // drop-ins deliberately differ from it --
  //   FleDecoder::Skip() landing on the end of the data returns false (fle-encoding.h:381-402), and
  //     AssembleRows then drops the rest of the row group (hdfs-parquet-scanner.cc:1107-1113);
  //   DictEncoder::WriteData(buf, len, indices) stores Log2(max_indice + 1) in the page's width byte
  //     (dict-encoding.h:438-444) although the data is packed with bit_width();
  //   ReadValue()/SkipValue() compare the definition level with max_def_level() even for REQUIRED columns,
  //     whose ReadDefinitionLevel() returns 1: every value becomes NULL (hdfs-parquet-scanner.cc:1006-1037);
  //   PLAIN pages: `literal OP value` operand order, In() a no-op, no NULL handling (parquet-common.h:197-255,
  //     hdfs-parquet-scanner.cc:346-349);
  //   predicate literals >= 2^bit_width keep only their low bits (library option "reference_compat").
  static bool reference_compat() {
    std::atomic<int>& f = compat_flag();
    int v = f.load(std::memory_order_relaxed);
    if (v < 0) {
      const char* e = std::getenv("FLE_B200_REFERENCE_COMPAT");
      v = (e && std::atoi(e) != 0) ? 1 : 0;
      f.store(v, std::memory_order_relaxed);
    }
    return v != 0;
  }
  static void SetReferenceCompat(bool on) { compat_flag().store(on ? 1 : 0, std::memory_order_relaxed); }

  // Grow-only scratch of the calling thread, recycled between scanners (a scanner lives for one row
  // group; allocating per scanner costs more than scanning).  Two banks, so that row group k+1 can be
  // uploaded while row group k is still being evaluated.  Freed at thread exit.
  static void* DeviceScratch(int bank, size_t slot, size_t bytes) { return holder().Scratch(bank, slot, bytes, false); }
  static void* PinnedScratch(int bank, size_t slot, size_t bytes) { return holder().Scratch(bank, slot, bytes, true); }
  static int NextBank() { return holder().next_bank++ & 1; }

  // Records and returns rc; rc != FLE_OK marks the thread's status as failed.
  static int Check(int rc, const char* what) {
    if (rc != FLE_OK) Fail(rc, what);
    return rc;
  }

  static bool ok() { return holder().status == FLE_OK; }
  static int status() { return holder().status; }
  static const std::string& message() { return holder().message; }
  static void ClearStatus() {
    holder().status = FLE_OK;
    holder().message.clear();
  }

 private:
  struct Buf {
    void* p = NULL;
    size_t bytes = 0;
  };
  struct Holder {
    fle_ctx* ctx = NULL;
    fle_ctx* ctx2 = NULL;  // the second scratch bank's context (bank_ctx)
    bool tried = false, tried2 = false;
    int status = FLE_OK;
    int compat_applied = 0;
    int next_bank = 0;
    std::string message;
    std::vector<Buf> dev[2], pinned[2];
    void* Scratch(int bank, size_t slot, size_t bytes, bool host) {
      std::vector<Buf>& v = host ? pinned[bank & 1] : dev[bank & 1];
      if (v.size() <= slot) v.resize(slot + 1);
      Buf& b = v[slot];
      if (b.p && b.bytes >= bytes) return b.p;
      fle_ctx* c = FleGpu::ctx();
      if (c == NULL) return NULL;
      if (b.p) {
        if (host) fle_host_free(b.p); else fle_dev_free(c, b.p);
        b.p = NULL;
        b.bytes = 0;
      }
      const size_t want = bytes + bytes / 4 + 4096;
      const int rc = host ? fle_host_alloc(want, &b.p) : fle_dev_alloc(c, want, &b.p);
      if (FleGpu::Check(rc, host ? "fle_host_alloc" : "fle_dev_alloc") != FLE_OK) return NULL;
      b.bytes = want;
      return b.p;
    }
    ~Holder() {
      for (int k = 0; k < 2; ++k) {
        for (size_t i = 0; i < dev[k].size(); ++i)
          if (dev[k][i].p && ctx) fle_dev_free(ctx, dev[k][i].p);
        for (size_t i = 0; i < pinned[k].size(); ++i)
          if (pinned[k][i].p) fle_host_free(pinned[k][i].p);
      }
      if (ctx2) fle_ctx_destroy(ctx2);
      if (ctx) fle_ctx_destroy(ctx);
    }
  };
  static std::atomic<int>& compat_flag() {
    static std::atomic<int> f(-1);
    return f;
  }
  static Holder& holder() {
    static thread_local Holder h;
    return h;
  }
  static void Fail(int rc, const char* what) {
    Holder& h = holder();
    h.status = rc;
    h.message = std::string(what) + ": " + fle_strerror(rc);
    if (h.ctx) h.message += std::string(" (") + fle_last_error(h.ctx) + ")";
    std::fprintf(stderr, "[fle_b200] %s\n", h.message.c_str());
  }
};

}  // namespace impala
#endif
