// "Next" rows of the scope table: what sits either side of the FLE planes
// inside a data page.
//
//  * realign_kernel -- page bodies are byte streams: [int32 def_len][def levels][u8 bw][groups]
//    (/root/reference/hdfs-parquet-table-writer.cc:567-575, dict-encoding.h:185-192), so the FLE
//    groups of a page start at an ODD byte offset.  A chunk that is kept in HBM as the raw file
//    image is re-based onto 16-byte aligned plane buffers with a funnel-shift copy
//    (__funnelshift_r over adjacent 32-bit words), one coalesced pass.
//  * plain_eval_kernel -- predicates on PLAIN-encoded fixed-width columns (those that overflow
//    the 40000-entry dictionary, e.g. l_extendedprice; hdfs-parquet-table-writer.cc:276-279):
//    /root/reference/parquet-common.h:197-250 compares row by row and push_back()s one bit at
//    a time.  Here a warp compares 32 rows per instruction and __ballot_sync emits the 32-row
//    selection word directly; __popc accumulates the match count.
//    NB the reference's generic templates compare `val OP data[i]` (literal on the LEFT,
//    parquet-common.h:214,225,236,247) while its Decimal specialisations and the whole
//    dictionary path mean `data[i] OP val`; the default here is the SQL meaning
//    (`column OP literal`), the literal reference order is available as a flag.  The reference's
//    PLAIN In() is an empty stub (:252-255); here it is implemented.
#include "fle_device.cuh"
#include "fle_internal.h"

namespace fle {

// ------------------------------------------------------------------------------------------
// funnel-shift realignment
// ------------------------------------------------------------------------------------------
// One thread writes 16 output bytes from the two 16-byte aligned source chunks that hold them
// (the second chunk is the neighbour's first: an L1 hit), so both sides move as 128-bit accesses.
// `src16` = source address rounded down to 16 bytes, `skew` = bytes dropped from its front (0..15).
__global__ void __launch_bounds__(256)
realign_kernel(const uint4* __restrict__ src16, uint32_t skew, int64_t nchunks_out, int64_t nbytes,
               int64_t src_chunks, uint4* __restrict__ dst) {
  const uint32_t wsel = skew >> 2, sh = (skew & 3u) * 8u;
  for (int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x; c < nchunks_out; c += (int64_t)gridDim.x * 256) {
    const uint4 a = __ldg(src16 + c);
    // the upper chunk is only touched when it holds bytes of the stream
    const uint4 b = (skew != 0 && c + 1 < src_chunks) ? __ldg(src16 + c + 1) : make_uint4(0, 0, 0, 0);
    const uint32_t w[9] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, 0u};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t lo = 0, hi = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {  // wsel is uniform: a 4-way select, not a local-memory index
        if (wsel == (uint32_t)k) {
          lo = w[j + k];
          hi = w[j + k + 1];
        }
      }
      o[j] = __funnelshift_r(lo, hi, sh);
      const int64_t left = nbytes - (c * 16 + j * 4);
      if (left < 4) o[j] = left <= 0 ? 0u : (o[j] & ((1u << (8 * (int)left)) - 1u));  // zero the bytes past the end
    }
    dst[c] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

int launch_realign(fle_ctx* ctx, const uint8_t* src_dev, size_t nbytes, uint8_t* dst_dev) {
  if (!src_dev || !dst_dev) return set_error(ctx, FLE_E_INVALID, "null pointer");
  if (reinterpret_cast<uintptr_t>(dst_dev) & 15) return set_error(ctx, FLE_E_ALIGN, "realign destination must be 16-byte aligned");
  if (nbytes == 0) return FLE_OK;
  const uintptr_t a = reinterpret_cast<uintptr_t>(src_dev);
  const uint32_t skew = (uint32_t)(a & 15);
  const uint4* src16 = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
  // whole 16-byte chunks are written: the destination holds (nbytes + 15) / 16 * 16 bytes
  const int64_t nchunks = (int64_t)((nbytes + 15) / 16);
  const int64_t src_chunks = (int64_t)((skew + nbytes + 15) / 16);  // aligned chunks that hold stream bytes
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * 16, (nchunks + 255) / 256));
  realign_kernel<<<grid, 256, 0, ctx->stream>>>(src16, skew, nchunks, (int64_t)nbytes, src_chunks, reinterpret_cast<uint4*>(dst_dev));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "realign launch");
  ctx->launches += 1;
  return FLE_OK;
}

// ------------------------------------------------------------------------------------------
// PLAIN predicates
// ------------------------------------------------------------------------------------------
constexpr int kPlainMaxLits = 1024;

struct PlainParams {
  const void* values;
  int64_t n;
  int op;            // FLE_OP_EQ..FLE_OP_IN
  int nlits;
  int swap;          // 1: reference operand order `literal OP value`
  uint32_t* bitmap;  // 32-row words (same memory image as the 64-bit LSB-first bitmap)
  int64_t nwords32;  // words to write (even: covers whole 64-bit words)
  uint64_t* partial;
  unsigned int* ticket;
  uint64_t* count_out;
  const void* lits;  // device copy of the literals (lists longer than kPlainInlineLits)
  uint64_t lit_inline[8];  // ... shorter lists travel in the parameter block: no staging copy, no host sync
};
constexpr int kPlainInlineLits = 8;

// value OP literal, OP fixed at compile time (the reference's operand order `literal OP value` is
// the mirrored operator, chosen on the host)
template <int OP, typename T>
__device__ __forceinline__ bool plain_test(T v, T lit0, const T* lits_s, int nlits) {
  if (OP == FLE_OP_EQ) return v == lit0;
  if (OP == FLE_OP_LT) return v < lit0;
  if (OP == FLE_OP_LE) return v <= lit0;
  if (OP == FLE_OP_GT) return v > lit0;
  if (OP == FLE_OP_GE) return v >= lit0;
  bool p = false;  // FLE_OP_IN
  for (int j = 0; j < nlits; ++j) p = p || (v == lits_s[j]);
  return p;
}

// bits i of x (i < 8) -> bit 4 i
__device__ __forceinline__ uint32_t spread8_by4(uint32_t x) {
  x = (x | (x << 12)) & 0x000F000Fu;
  x = (x | (x << 6)) & 0x03030303u;
  x = (x | (x << 3)) & 0x11111111u;
  return x;
}

// VEC = 4 (4-byte types on 16-byte aligned columns): a lane loads 4 consecutive values with one
// 128-bit load, 4 x 16 bytes in flight per lane (one 4-byte load per lane and 8 in flight left the
// kernel latency-bound at 0.83 of roofline; 8-byte types reach 0.93-0.96 that way).  Ballot j then
// holds rows 4 l + j of a 128-row block, and the four 32-row words of the block are byte k of the
// four ballots interleaved bit by bit.
template <typename T, int OP, int VEC>
__global__ void __launch_bounds__(256) plain_eval_kernel(const PlainParams P) {
  __shared__ T lits_s[OP == FLE_OP_IN ? kPlainMaxLits : 1];
  __shared__ uint64_t warp_cnt[8];
  __shared__ bool is_last;
  if (OP == FLE_OP_IN) {
    for (int i = threadIdx.x; i < P.nlits; i += 256)
      lits_s[i] = P.nlits <= kPlainInlineLits ? reinterpret_cast<const T*>(P.lit_inline)[i]
                                              : reinterpret_cast<const T*>(P.lits)[i];
    __syncthreads();
  }
  const T* vals = reinterpret_cast<const T*>(P.values);
  const T lit0 = reinterpret_cast<const T*>(P.lit_inline)[0];
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * 256) >> 5;
  constexpr int U = VEC == 4 ? 16 : 8;  // 32-row words per warp iteration (VEC = 1: 8 coalesced loads in flight per lane)
  uint64_t cnt = 0;
  for (int64_t w0 = warp * U; w0 < P.nwords32; w0 += nwarps * U) {
    uint32_t mine = 0;
    if (VEC == 4 && sizeof(T) == 4 && (w0 + U) * 32 <= P.n) {
      const uint4* p4 = reinterpret_cast<const uint4*>(vals + w0 * 32) + lane;
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ldg_stream128(p4 + u * 32);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t raw[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
        uint32_t b[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          T x;
          memcpy(&x, &raw[j], 4);
          b[j] = __ballot_sync(0xffffffffu, plain_test<OP, T>(x, lit0, lits_s, P.nlits));
        }
        if ((lane >> 2) == u) {  // lane 4 u + k assembles word k of block u
          const int sh = (lane & 3) * 8;
          mine = spread8_by4((b[0] >> sh) & 0xFFu) | (spread8_by4((b[1] >> sh) & 0xFFu) << 1) |
                 (spread8_by4((b[2] >> sh) & 0xFFu) << 2) | (spread8_by4((b[3] >> sh) & 0xFFu) << 3);
        }
      }
    } else if (VEC != 4 && (w0 + U) * 32 <= P.n) {  // whole iteration inside the column: no bounds checks
      const T* p = vals + w0 * 32 + lane;
      T v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = p[u * 32];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t word = __ballot_sync(0xffffffffu, plain_test<OP, T>(v[u], lit0, lits_s, P.nlits));
        if (lane == u) mine = word;  // rows 32(w0+u) .. +31, LSB first
      }
    } else {
#pragma unroll 1
      for (int u = 0; u < U; ++u) {
        const int64_t r = (w0 + u) * 32 + lane;
        const bool pr = r < P.n && plain_test<OP, T>(vals[r < P.n ? r : 0], lit0, lits_s, P.nlits);
        const uint32_t word = __ballot_sync(0xffffffffu, pr);
        if (lane == u) mine = word;
      }
    }
    if (lane < U && w0 + lane < P.nwords32) {
      P.bitmap[w0 + lane] = mine;
      cnt += __popc(mine);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) warp_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t t = 0;
    for (int i = 0; i < 8; ++i) t += warp_cnt[i];
    P.partial[blockIdx.x] = t;
    __threadfence();
    is_last = (atomicAdd(P.ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {  // the last CTA sums the per-CTA counts, all threads (a serial loop over ~1200 partials cost 20-30 us)
    __threadfence();
    uint64_t t = 0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += 256) t += *((volatile uint64_t*)&P.partial[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    __syncthreads();
    if (lane == 0) warp_cnt[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint64_t tot = 0;
      for (int i = 0; i < 8; ++i) tot += warp_cnt[i];
      if (P.count_out) *P.count_out = tot;
      *P.ticket = 0;
    }
  }
}

template <typename T, int VEC>
void launch_plain_typed(const PlainParams& P, int op, int grid, cudaStream_t st) {
  switch (op) {
    case FLE_OP_EQ: plain_eval_kernel<T, FLE_OP_EQ, VEC><<<grid, 256, 0, st>>>(P); break;
    case FLE_OP_LT: plain_eval_kernel<T, FLE_OP_LT, VEC><<<grid, 256, 0, st>>>(P); break;
    case FLE_OP_LE: plain_eval_kernel<T, FLE_OP_LE, VEC><<<grid, 256, 0, st>>>(P); break;
    case FLE_OP_GT: plain_eval_kernel<T, FLE_OP_GT, VEC><<<grid, 256, 0, st>>>(P); break;
    case FLE_OP_GE: plain_eval_kernel<T, FLE_OP_GE, VEC><<<grid, 256, 0, st>>>(P); break;
    default: plain_eval_kernel<T, FLE_OP_IN, VEC><<<grid, 256, 0, st>>>(P); break;
  }
}

int launch_plain_eval(fle_ctx* ctx, const void* values_dev, int elem_type, int64_t n, int op, const void* lits_host,
                      int nlits, int reference_order, uint64_t* bitmap_dev, uint64_t* count_dev) {
  static const int kElemSize[4] = {4, 8, 4, 8};
  if (elem_type < 0 || elem_type > 3) return set_error(ctx, FLE_E_INVALID, "unknown element type");
  if (op < FLE_OP_EQ || op > FLE_OP_IN) return set_error(ctx, FLE_E_INVALID, "PLAIN predicates are EQ..IN leaves");
  if (n < 0 || nlits < (op == FLE_OP_IN ? 0 : 1) || nlits > kPlainMaxLits || (nlits > 0 && !lits_host))
    return set_error(ctx, FLE_E_INVALID, "bad literal list");
  if (op != FLE_OP_IN && nlits != 1) return set_error(ctx, FLE_E_INVALID, "a comparison takes exactly one literal");
  if (n > 0 && (!values_dev || !bitmap_dev)) return set_error(ctx, FLE_E_INVALID, "null pointer");
  const int es = kElemSize[elem_type];
  if (reinterpret_cast<uintptr_t>(values_dev) & (es - 1))
    return set_error(ctx, FLE_E_ALIGN, "PLAIN values must be naturally aligned (use fle_realign_dev)");
  if (n == 0) {
    if (count_dev) FLE_CUDA_TRY(ctx, cudaMemsetAsync(count_dev, 0, 8, ctx->stream));
    return FLE_OK;
  }
  PlainParams P;
  std::memset(&P, 0, sizeof(P));
  P.values = values_dev;
  P.n = n;
  P.op = op;
  P.nlits = nlits;
  P.swap = reference_order ? 1 : 0;
  P.bitmap = reinterpret_cast<uint32_t*>(bitmap_dev);
  P.nwords32 = ((n + 63) / 64) * 2;
  P.count_out = count_dev;
  const bool vec4 = es == 4 && (reinterpret_cast<uintptr_t>(values_dev) & 15) == 0;
  const int64_t warps_needed = (P.nwords32 + (vec4 ? 15 : 7)) / (vec4 ? 16 : 8);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * (es == 8 ? 6 : 8), (warps_needed + 7) / 8));
  int rc = ensure_scratch(ctx, grid, std::max(2 * nlits, 2));
  if (rc != FLE_OK) return rc;
  if (nlits <= kPlainInlineLits) {
    if (nlits > 0) std::memcpy(P.lit_inline, lits_host, (size_t)nlits * es);
  } else {
    // longer lists travel through one of the context's literal slots (no stream synchronisation)
    std::memcpy(P.lit_inline, lits_host, es);  // literal 0 is always inline
    const uint32_t* d = nullptr;
    rc = upload_literals(ctx, lits_host, (int)(((size_t)nlits * es + 3) / 4), nullptr, nullptr, &d);
    if (rc != FLE_OK) return rc;
    P.lits = d;
  }
  P.partial = ctx->d_partial;
  P.ticket = ctx->d_ticket;
  // `literal OP value` (the reference's generic template) is `value OP' literal` with OP' mirrored
  int kop = op;
  if (reference_order) {
    if (op == FLE_OP_LT) kop = FLE_OP_GT;
    else if (op == FLE_OP_GT) kop = FLE_OP_LT;
    else if (op == FLE_OP_LE) kop = FLE_OP_GE;
    else if (op == FLE_OP_GE) kop = FLE_OP_LE;
  }
  switch (elem_type) {
    case 0: vec4 ? launch_plain_typed<int32_t, 4>(P, kop, grid, ctx->stream) : launch_plain_typed<int32_t, 1>(P, kop, grid, ctx->stream); break;
    case 1: launch_plain_typed<int64_t, 1>(P, kop, grid, ctx->stream); break;
    case 2: vec4 ? launch_plain_typed<float, 4>(P, kop, grid, ctx->stream) : launch_plain_typed<float, 1>(P, kop, grid, ctx->stream); break;
    case 3: launch_plain_typed<double, 1>(P, kop, grid, ctx->stream); break;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "plain_eval launch");
  ctx->launches += 1;
  return FLE_OK;
}

}  // namespace fle
