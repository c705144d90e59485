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

#include <cstdio>
#include <cstdlib>
#include <string>

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
    }
    return h.ctx;
  }

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
  struct Holder {
    fle_ctx* ctx = NULL;
    bool tried = false;
    int status = FLE_OK;
    std::string message;
    ~Holder() {
      if (ctx) fle_ctx_destroy(ctx);
    }
  };
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
