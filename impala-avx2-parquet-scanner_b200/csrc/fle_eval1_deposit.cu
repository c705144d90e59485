// Nullable columns with NULLs present: evaluate + deposit in ONE kernel.
//
// The reference evaluates a leaf on a nullable column in the data-ordinal domain (the data stream holds only the
// non-null values) and deposits the result through the column's not-null bitmap
// (ColumnReader<T>::Eq..In + IntersectBitset, /root/reference/hdfs-parquet-scanner.cc:326-345).  The two-step form
// (fused evaluation into a data-domain bitmap, then deposit_kernel, fle_bitmap.cu) writes that bitmap and reads it
// back.  Here every thread owns eight consecutive ROW words: it knows the rank of its first row among the non-null
// rows (block prefix + in-block scan of the not-null popcounts, read straight from the def-level planes), so the
// data values its rows map to are the ~8 consecutive data groups starting at rank / 64 -- it evaluates exactly those
// groups with the single-pass plane algebra of fle_eval1.cuh (compare chains on the planes, straight from global
// memory: neighbouring threads read neighbouring groups), funnels the result bits to the rank's bit offset and
// deposits them into the set positions of each not-null word.  Nothing but the def planes, the data planes and the
// output bitmap moves; the output is combined with what it already holds (conjunct chains fold as they go).
// Covers sub-programs that lower to one pass (a compare, BETWEEN, an equality) on bw <= 16; everything else takes
// the two-step form.
#include <algorithm>
#include <mutex>
#include <set>
#include <utility>

#include "fle_bitmap.cuh"
#include "fle_eval1.cuh"
#include "fle_internal.h"

namespace fle {

struct EvalDepositParams {
  PassDesc pd;
  const uint64_t* planes;   // the column's data stream (non-null values only)
  int64_t n_sub;            // values in it
  const uint64_t* def_planes;
  int64_t n_root, nwords;   // rows, row words
  const uint64_t* block_sums;
  uint64_t* out;
  int32_t comb_op, pad;
  unsigned long long* count;
};

// A CTA covers kEdWords row words (32768 rows): its rows map to at most kEdWords + 1 consecutive data groups, which are
// staged in shared memory with coalesced loads (a thread-per-group walk over global memory issues one sector request
// per 8-byte plane word: 0.94 ms for Q6 with NULLs against 0.77 ms for the two-step form).  Each staged group takes
// BW + 1 words: the odd stride spreads the per-thread group reads over the banks.
constexpr int kEdThreads = 256;
constexpr int kEdWordsPerThread = 2;
constexpr int kEdWords = kEdThreads * kEdWordsPerThread;

__global__ void __launch_bounds__(kEdThreads)
ed_block_counts_kernel(const uint64_t* __restrict__ def_planes, int64_t n_root, int64_t nwords, uint64_t* __restrict__ block_sums) {
  __shared__ uint64_t warp_buf[kEdThreads / 32];
  const int64_t w0 = (int64_t)blockIdx.x * kEdWords;
  uint64_t c = 0;
#pragma unroll
  for (int k = 0; k < kEdWordsPerThread; ++k) c += __popcll(def_word(def_planes, w0 + k * kEdThreads + threadIdx.x, nwords, n_root));
  const uint64_t t = block_sum<kEdThreads>(c, warp_buf);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = t;
}

// block sums -> exclusive prefixes, in place, plus the total at [n] (one CTA: n is a few tens of thousands).  With
// every CTA of the deposit kernel summing its predecessors itself, 18 K CTAs read 1.3 GB of block sums per launch.
__global__ void __launch_bounds__(1024) ed_prefix_kernel(uint64_t* __restrict__ sums, int n) {
  __shared__ uint64_t warp_tot[32];
  __shared__ uint64_t carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const uint64_t v = i < n ? sums[i] : 0ull;
    uint64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint64_t woff = 0;
    for (int w = 0; w < warp; ++w) woff += warp_tot[w];
    const uint64_t carry = carry_s;
    if (i < n) sums[i] = carry + woff + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + woff + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[n] = carry_s;
}

struct PaddedSmemGroup {  // a staged group: BW plane words at consecutive 8-byte slots
  uint32_t base;
  template <int BW>
  __device__ __forceinline__ void load_pair(int i, uint64_t& w0, uint64_t& w1) const {
    w0 = lds64(base + i * 8);
    w1 = (i + 1 < BW) ? lds64(base + (i + 1) * 8) : 0ull;
  }
};

template <bool IS_LT, int K, int BW>
__global__ void __launch_bounds__(kEdThreads) eval1_deposit_kernel(const __grid_constant__ EvalDepositParams D) {
  // [staged planes: (kEdWords + 2) groups x (BW + 1) words][result words: kEdWords + 2]
  extern __shared__ __align__(16) uint8_t ed_smem[];
  __shared__ uint64_t warp_buf[kEdThreads / 32];
  __shared__ uint32_t warp_scan[kEdThreads / 32];
  constexpr int kSlot = BW + 1;
  constexpr int kMaxGroups = kEdWords + 2;
  uint64_t* staged = reinterpret_cast<uint64_t*>(ed_smem);
  uint64_t* res = staged + kMaxGroups * kSlot;
  PassRegs<IS_LT, K, BW> pr;
  pr.init(D.pd);
  const int64_t block_off = (int64_t)D.block_sums[blockIdx.x];  // exclusive prefixes (ed_prefix_kernel)
  const int64_t block_total = (int64_t)D.block_sums[blockIdx.x + 1] - block_off;
  // thread owns kEdWordsPerThread CONSECUTIVE words so that its ranks are contiguous
  const int64_t w0 = (int64_t)blockIdx.x * kEdWords + (int64_t)threadIdx.x * kEdWordsPerThread;
  uint64_t m[kEdWordsPerThread];
  uint32_t mine = 0;
#pragma unroll
  for (int k = 0; k < kEdWordsPerThread; ++k) {
    m[k] = def_word(D.def_planes, w0 + k, D.nwords, D.n_root);
    mine += __popcll(m[k]);
  }
  // 1. stage the data groups this CTA's rows map to, [g_first, g_end), with coalesced loads
  const int64_t ngroups_sub = (D.n_sub + 63) >> 6;
  const int64_t g_first = block_off >> 6;
  const int64_t g_end = min(ngroups_sub, ((block_off + block_total + 63) >> 6) + 1);
  const int ng = block_total > 0 ? (int)max((int64_t)0, g_end - g_first) : 0;
  {
    // asynchronous 8-byte copies (LDGSTS): a thread has all its ~2 * BW copies in flight at once; through registers
    // the loop waits for every batch of four loads and the kernel is latency-bound (ncu: 15 % of the stall samples
    // on the first store of the batch, bw 12 290 us against 157 us of HBM time)
    const uint64_t* src = D.planes + g_first * BW;
    const uint32_t dst0 = smem_u32(staged);
    for (int i = threadIdx.x; i < ng * BW; i += kEdThreads) {
      const int g = i / BW, p = i - g * BW;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst0 + (uint32_t)(g * kSlot + p) * 8u), "l"(src + i) : "memory");
    }
    cp_async_commit();
  }
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += y;
  }
  if ((threadIdx.x & 31) == 31) warp_scan[threadIdx.x >> 5] = incl;
  cp_async_wait<0>();
  __syncthreads();
  // 2. every staged group is evaluated ONCE, by the thread whose index it has (neighbouring threads read neighbouring
  //    slots: the odd slot stride keeps the 64-bit shared loads conflict-free); data-ordinal order, values past the
  //    end of the stream read as 0
  const uint32_t stage = smem_u32(staged);
  for (int q = threadIdx.x; q < ng; q += kEdThreads)
    res[q] = finish_word(pr.run(PaddedSmemGroup{stage + (uint32_t)q * (uint32_t)(kSlot * 8)}), g_first + q, D.n_sub);
  __syncthreads();
  // 3. deposit: the bits of a row word start at its rank
  uint32_t warp_off = 0;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) warp_off += warp_scan[w];
  int64_t rank = block_off + warp_off + (incl - mine);
  uint64_t cnt = 0;
#pragma unroll
  for (int k = 0; k < kEdWordsPerThread; ++k) {
    const int64_t w = w0 + k;
    if (w < D.nwords) {
      const int q = (int)((rank >> 6) - g_first);
      const int sh = (int)(rank & 63);
      const uint64_t lo = (m[k] && q < ng) ? res[q] : 0ull;
      const uint64_t hi = (m[k] && sh && q + 1 < ng) ? res[q + 1] : 0ull;
      const uint64_t bits = sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
      uint64_t r = deposit_bits(bits, m[k]);
      if (D.comb_op == FOLD_AND) r &= D.out[w];
      else if (D.comb_op == FOLD_OR) r |= D.out[w] & tail_mask_word(w, D.n_root);
      D.out[w] = r;
      cnt += __popcll(r);
      rank += __popcll(m[k]);
    }
  }
  if (D.count) {
    __syncthreads();
    const uint64_t t = block_sum<kEdThreads>(cnt, warp_buf);
    if (threadIdx.x == 0 && t) atomicAdd(D.count, (unsigned long long)t);
  }
}

#define FLE_BW16(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)

template <bool IS_LT, int K>
static const void* deposit_for_bw(int bw) {
  switch (bw) {
#define X(B) case B: return (const void*)eval1_deposit_kernel<IS_LT, K, B>;
    FLE_BW16(X)
#undef X
  }
  return nullptr;
}

static const void* eval1_deposit_kernel_for(int type, int k, int bw) {
  if (type == PASS_LT) {
    if (k == 1) return deposit_for_bw<true, 1>(bw);
    if (k == 2) return deposit_for_bw<true, 2>(bw);
  } else if (type == PASS_EQ) {
    if (k == 1) return deposit_for_bw<false, 1>(bw);
  }
  return nullptr;
}

// FLE_E_UNSUPPORTED: the sub-program is not one supported pass (the caller evaluates and deposits in two steps)
int launch_eval_deposit(fle_ctx* ctx, const fle_col_desc& col, const fle_pred_node* nodes, int nnodes, int64_t n_sub,
                        const uint64_t* def_planes_dev, int64_t n_root, uint64_t* out_dev, int comb_op, uint64_t* count_dev) {
  if (col.bit_width < 1 || col.bit_width > 16 || n_root <= 0 || n_sub <= 0) return FLE_E_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(col.planes) & 7) return FLE_E_UNSUPPORTED;
  if (col.bytes < (uint64_t)((n_sub + 63) / 64) * (uint64_t)col.bit_width * 8u) return FLE_E_UNSUPPORTED;  // (the two-step form reports it)
  fle_col_desc dc{col.planes, col.bytes, col.bit_width, 0, nullptr, 0};
  LoweredProgram lp;
  std::string err;
  if (lower_program(&dc, 1, nodes, nnodes, &lp, &err, ctx->reference_compat) != FLE_OK) return FLE_E_UNSUPPORTED;
  const PassDesc& p0 = lp.p.passes[0];
  // (the condition launch_prepared sends a program to the single-pass kernels of fle_eval1.cu under)
  const bool single_pass = lp.p.npasses == 1 && lp.p.ncombs == 0 && (p0.type == PASS_LT || p0.type == PASS_EQ) &&
                           (p0.k == 1 || p0.merge) && p0.slot[0] == 0;
  if (!single_pass) return FLE_E_UNSUPPORTED;
  const void* kern = eval1_deposit_kernel_for(p0.type, p0.k, col.bit_width);
  if (!kern) return FLE_E_UNSUPPORTED;
  const int64_t nwords = (n_root + 63) / 64;
  const int nblocks = (int)((nwords + kEdWords - 1) / kEdWords);
  int rc = ensure_scratch(ctx, nblocks + 1, 0);
  if (rc != FLE_OK) return rc;
  const size_t smem = (size_t)(kEdWords + 2) * (size_t)(col.bit_width + 2) * 8;  // staged planes + result words
  if (smem > 48 * 1024) {  // opt in once per kernel and device (process-wide)
    static std::mutex mu;
    static std::set<std::pair<int, const void*> > raised;
    std::lock_guard<std::mutex> lock(mu);
    const std::pair<int, const void*> key(ctx->device, kern);
    if (raised.find(key) == raised.end()) {
      cudaError_t ea = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (ea != cudaSuccess) return set_cuda_error(ctx, ea, "cudaFuncSetAttribute(eval1_deposit)");
      raised.insert(key);
    }
  }
  if (count_dev) FLE_CUDA_TRY(ctx, cudaMemsetAsync(count_dev, 0, 8, ctx->stream));
  ed_block_counts_kernel<<<nblocks, kEdThreads, 0, ctx->stream>>>(def_planes_dev, n_root, nwords, ctx->d_partial);
  ed_prefix_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->d_partial, nblocks);
  EvalDepositParams D;
  D.pd = p0;
  D.planes = reinterpret_cast<const uint64_t*>(col.planes);
  D.n_sub = n_sub;
  D.def_planes = def_planes_dev;
  D.n_root = n_root;
  D.nwords = nwords;
  D.block_sums = ctx->d_partial;
  D.out = out_dev;
  D.comb_op = comb_op;
  D.pad = 0;
  D.count = reinterpret_cast<unsigned long long*>(count_dev);
  void* args[] = {(void*)&D};
  cudaError_t e = cudaLaunchKernel(kern, dim3(nblocks), dim3(kEdThreads), args, smem, ctx->stream);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_cuda_error(ctx, e, "eval1_deposit launch");
  }
  ctx->launches += 3;
  return FLE_OK;
}

}  // namespace fle
