// Row-group partitioning across the GPUs of one box behind the C ABI (include/fle_b200.h,
// "groups"): every GPU scans its own contiguous range of row groups (row groups are independent,
// hdfs-parquet-scanner.cc:1056-1060) and the only exchange is the SUM of the match count and of
// one exact int64 partial aggregate per query.  That sum is not a separate collective: the last
// CTA of the scan kernel, which owns the GPU's totals, posts them into every peer's mailbox with
// peer stores over NVLink/NVSwitch and sums the peers' entries from its own mailbox
// (group_exchange_warp, fle_eval_common.cuh) -- scan + all-reduce are ONE kernel, with no host
// and no second launch in the loop.
//
// Two ways to form a group:
//   * one process, several devices (or several contexts on one device): fle_group_create_local
//     (peer access enabled between the devices, mailbox pointers exchanged directly);
//   * one process per GPU (torch.distributed / MPI style): fle_group_export -> the caller
//     all-gathers the 64-byte handles by whatever transport it has -> fle_group_open
//     (cudaIpcOpenMemHandle).
#include <cstring>

#include "fle_eval_common.cuh"
#include "fle_internal.h"

namespace fle {

struct GroupReduceParams {
  uint64_t* mail[FLE_MAX_GROUP];
  unsigned long long* lost;
  int32_t world, rank;
  int32_t deferred, flush_only;
  uint32_t seq, pend_seq;
  uint64_t* pend_count;
  uint64_t* pend_sum;
  uint64_t timeout_ns;
  uint64_t* a;
  uint64_t* b;
};

// one warp: the exchange for paths whose last kernel does not carry it (a, b hold the local totals and
// receive the group's), or just the collection of a pending deferred exchange (flush_only)
__global__ void __launch_bounds__(32) group_reduce_kernel(const __grid_constant__ GroupReduceParams G) {
  const GroupView gv{G.mail, G.lost, G.world, G.rank, G.timeout_ns};
  uint64_t a = 0, b = 0;
  if (!G.flush_only) {
    a = *G.a;
    b = G.b ? *G.b : 0ull;
    group_post(gv, G.seq, a, b);
  }
  group_collect_pending(gv, G.pend_seq, G.pend_count, G.pend_sum);
  if (G.flush_only || G.deferred) return;
  group_collect(gv, G.seq, a, b);
  if (threadIdx.x == 0) {
    *G.a = a;
    if (G.b) *G.b = b;
  }
}

// the host keeps the sequence numbers: every member of a group issues the same sequence of exchanges
struct Exchange {
  uint32_t seq, pend_seq;
  uint64_t* pend_count;
  uint64_t* pend_sum;
};
static Exchange next_exchange(fle_ctx* ctx, uint64_t* count_out, uint64_t* sum_out, bool flush_only) {
  Exchange x{0, ctx->grp_pend_seq, ctx->grp_pend_count, ctx->grp_pend_sum};
  ctx->grp_pend_seq = 0;
  ctx->grp_pend_count = ctx->grp_pend_sum = nullptr;
  if (flush_only) return x;
  if (++ctx->grp_seq == 0) ++ctx->grp_seq;  // 0 tags an empty mailbox entry
  x.seq = ctx->grp_seq;
  if (ctx->grp_deferred) {
    ctx->grp_pend_seq = x.seq;
    ctx->grp_pend_count = count_out;
    ctx->grp_pend_sum = sum_out;
  }
  return x;
}

void group_fill(fle_ctx* ctx, EvalParams* P) {
  P->grp_world = 0;
  if (!ctx->grp_on || ctx->grp_world <= 1 || !P->count_out) return;
  for (int r = 0; r < ctx->grp_world; ++r) P->grp_mail[r] = ctx->grp_peer_mail[r];
  P->grp_lost = ctx->d_grp_state;
  P->grp_world = ctx->grp_world;
  P->grp_rank = ctx->grp_rank;
  P->grp_deferred = ctx->grp_deferred ? 1 : 0;
  P->grp_timeout_ns = (uint64_t)ctx->grp_timeout_ms * 1000000ull;
  const Exchange x = next_exchange(ctx, P->count_out, reinterpret_cast<uint64_t*>(P->agg_sum_out), false);
  P->grp_seq = x.seq;
  P->grp_pend_seq = x.pend_seq;
  P->grp_pend_count = x.pend_count;
  P->grp_pend_sum = x.pend_sum;
}

int launch_group_reduce(fle_ctx* ctx, uint64_t* a_dev, uint64_t* b_dev) {
  if (ctx->grp_world <= 1) return FLE_OK;
  const bool flush_only = a_dev == nullptr;
  if (flush_only && ctx->grp_pend_seq == 0) return FLE_OK;  // nothing pending
  GroupReduceParams G;
  std::memset(&G, 0, sizeof(G));
  G.deferred = ctx->grp_deferred ? 1 : 0;
  G.flush_only = flush_only ? 1 : 0;
  for (int r = 0; r < ctx->grp_world; ++r) G.mail[r] = ctx->grp_peer_mail[r];
  G.lost = ctx->d_grp_state;
  G.world = ctx->grp_world;
  G.rank = ctx->grp_rank;
  G.timeout_ns = (uint64_t)ctx->grp_timeout_ms * 1000000ull;
  G.a = a_dev;
  G.b = b_dev;
  const Exchange x = next_exchange(ctx, a_dev, b_dev, flush_only);
  G.seq = x.seq;
  G.pend_seq = x.pend_seq;
  G.pend_count = x.pend_count;
  G.pend_sum = x.pend_sum;
  group_reduce_kernel<<<1, 32, 0, ctx->stream>>>(G);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "group_reduce launch");
  ctx->launches += 1;
  return FLE_OK;
}

static int group_alloc_own(fle_ctx* ctx) {
  if (ctx->d_grp_mail) return FLE_OK;
  // cudaMalloc, not a stream-ordered pool: only such allocations can be exported over CUDA IPC
  FLE_CUDA_TRY(ctx, cudaMalloc(&ctx->d_grp_mail, (size_t)kGroupMailWords * 8));
  FLE_CUDA_TRY(ctx, cudaMemset(ctx->d_grp_mail, 0, (size_t)kGroupMailWords * 8));
  FLE_CUDA_TRY(ctx, cudaMalloc(&ctx->d_grp_state, 8));  // peers lost to time-outs
  FLE_CUDA_TRY(ctx, cudaMemset(ctx->d_grp_state, 0, 8));
  FLE_CUDA_TRY(ctx, cudaDeviceSynchronize());
  // load the reduce kernel now: with lazy module loading the first launch of a function may wait for
  // running kernels, and a member's kernel that is waiting for its peers must never be waited for
  cudaFuncAttributes fa;
  FLE_CUDA_TRY(ctx, cudaFuncGetAttributes(&fa, group_reduce_kernel));
  // ... and allocate now what the first launches would allocate lazily: an allocation can wait for the device as well
  // (a member with an empty row-group range has launched nothing before its first group call)
  return prime_context(ctx);
}

void group_release(fle_ctx* ctx) {
  if (ctx->grp_ipc) {
    for (int r = 0; r < ctx->grp_world; ++r)
      if (r != ctx->grp_rank && ctx->grp_peer_mail[r]) cudaIpcCloseMemHandle(ctx->grp_peer_mail[r]);
  }
  for (int r = 0; r < FLE_MAX_GROUP; ++r) ctx->grp_peer_mail[r] = nullptr;
  ctx->grp_world = 0;
  ctx->grp_rank = 0;
  ctx->grp_ipc = false;
  ctx->grp_seq = 0;
  ctx->grp_pend_seq = 0;
  ctx->grp_pend_count = ctx->grp_pend_sum = nullptr;
  if (ctx->d_grp_mail) cudaFree(ctx->d_grp_mail);
  if (ctx->d_grp_state) cudaFree(ctx->d_grp_state);
  ctx->d_grp_mail = nullptr;
  ctx->d_grp_state = nullptr;
}

struct DevGuard {
  int prev = -1;
  explicit DevGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DevGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

}  // namespace fle

using namespace fle;

extern "C" {

int fle_group_export(fle_ctx* ctx, void* handle_out) {
  if (!ctx || !handle_out) return FLE_E_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == FLE_GROUP_HANDLE_BYTES, "handle size");
  DevGuard g(ctx->device);
  if (ctx->grp_world > 0) return set_error(ctx, FLE_E_INVALID, "context already belongs to a group");
  int rc = group_alloc_own(ctx);
  if (rc != FLE_OK) return rc;
  cudaIpcMemHandle_t h;
  FLE_CUDA_TRY(ctx, cudaIpcGetMemHandle(&h, ctx->d_grp_mail));
  std::memcpy(handle_out, &h, sizeof(h));
  return FLE_OK;
}

int fle_group_open(fle_ctx* ctx, int rank, int world, const void* handles) {
  if (!ctx || world < 1 || world > FLE_MAX_GROUP || rank < 0 || rank >= world || (world > 1 && !handles))
    return ctx ? set_error(ctx, FLE_E_INVALID, "bad group arguments") : FLE_E_INVALID;
  DevGuard g(ctx->device);
  if (ctx->grp_world > 0) return set_error(ctx, FLE_E_INVALID, "context already belongs to a group");
  int rc = group_alloc_own(ctx);
  if (rc != FLE_OK) return rc;
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      ctx->grp_peer_mail[r] = ctx->d_grp_mail;
      continue;
    }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const uint8_t*>(handles) + (size_t)r * FLE_GROUP_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      for (int q = 0; q < r; ++q)
        if (q != rank && ctx->grp_peer_mail[q]) cudaIpcCloseMemHandle(ctx->grp_peer_mail[q]);
      for (int q = 0; q < FLE_MAX_GROUP; ++q) ctx->grp_peer_mail[q] = nullptr;
      return set_cuda_error(ctx, e, "cudaIpcOpenMemHandle(peer mailbox)");
    }
    ctx->grp_peer_mail[r] = static_cast<uint64_t*>(p);
  }
  ctx->grp_world = world;
  ctx->grp_rank = rank;
  ctx->grp_ipc = true;
  return FLE_OK;
}

int fle_group_create_local(fle_ctx* const* ctxs, int n) {
  if (!ctxs || n < 1 || n > FLE_MAX_GROUP) return FLE_E_INVALID;
  for (int i = 0; i < n; ++i) {
    if (!ctxs[i]) return FLE_E_INVALID;
    if (ctxs[i]->grp_world > 0) return set_error(ctxs[i], FLE_E_INVALID, "context already belongs to a group");
  }
  for (int i = 0; i < n; ++i) {
    DevGuard g(ctxs[i]->device);
    int rc = group_alloc_own(ctxs[i]);
    if (rc != FLE_OK) return rc;
    for (int j = 0; j < n; ++j) {
      if (ctxs[j]->device == ctxs[i]->device) continue;
      int can = 0;
      cudaError_t e = cudaDeviceCanAccessPeer(&can, ctxs[i]->device, ctxs[j]->device);
      if (e != cudaSuccess || !can) {
        cudaGetLastError();
        return set_error(ctxs[i], FLE_E_UNSUPPORTED, "no peer access between the devices of the group");
      }
      e = cudaDeviceEnablePeerAccess(ctxs[j]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return set_cuda_error(ctxs[i], e, "cudaDeviceEnablePeerAccess");
      cudaGetLastError();
    }
  }
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) ctxs[i]->grp_peer_mail[j] = ctxs[j]->d_grp_mail;
    ctxs[i]->grp_world = n;
    ctxs[i]->grp_rank = i;
    ctxs[i]->grp_ipc = false;
  }
  return FLE_OK;
}

int fle_group_flush(fle_ctx* ctx) {
  if (!ctx) return FLE_E_INVALID;
  DevGuard g(ctx->device);
  if (ctx->grp_world <= 1) return FLE_OK;
  return launch_group_reduce(ctx, nullptr, nullptr);
}

int fle_group_close(fle_ctx* ctx) {
  if (!ctx) return FLE_E_INVALID;
  DevGuard g(ctx->device);
  if (ctx->grp_world > 1) launch_group_reduce(ctx, nullptr, nullptr);
  cudaStreamSynchronize(ctx->stream);
  group_release(ctx);
  return FLE_OK;
}

int fle_group_status(fle_ctx* ctx, int* rank, int* world, uint64_t* reductions_done, uint64_t* peers_timed_out) {
  if (!ctx) return FLE_E_INVALID;
  DevGuard g(ctx->device);
  if (rank) *rank = ctx->grp_rank;
  if (world) *world = ctx->grp_world;
  unsigned long long st[2] = {ctx->grp_seq, 0};
  if (ctx->d_grp_state) {
    if (ctx->grp_world > 1) {
      int rc = launch_group_reduce(ctx, nullptr, nullptr);
      if (rc != FLE_OK) return rc;
    }
    FLE_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    FLE_CUDA_TRY(ctx, cudaMemcpy(st + 1, ctx->d_grp_state, 8, cudaMemcpyDeviceToHost));
  }
  if (reductions_done) *reductions_done = st[0];
  if (peers_timed_out) *peers_timed_out = st[1];
  return FLE_OK;
}

}  // extern "C"
