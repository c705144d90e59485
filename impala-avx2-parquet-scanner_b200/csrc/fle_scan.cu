// Survivor materialisation: what AssembleRows does AFTER the selection bitmap
// exists -- /root/reference/hdfs-parquet-scanner.cc:1134-1168 turns the bitmap into skip
// distances and then calls ReadValue(skip) -> DictDecoder::GetValue(T*, skip) ->
// FleDecoder::Get(T*, skip) (fle-encoding.h:344-379) once per survivor and column.
// On the device that is
//   select   : bitmap -> ascending row ids            (ordered stream compaction; fle_rowgroup.cu)
//   gather   : row id -> code = bit (63 - r%64) of the bw plane words of group r/64
//   lookup   : code -> dictionary value               (dense output column)
// plus the fused aggregate a TPC-H Q6 scan needs downstream of the bitmap:
//   sum_product : SUM over survivors of plain_i64[r] * dict[code(r)]   (exact int64)
#include "fle_device.cuh"
#include "fle_internal.h"

namespace fle {

__device__ __forceinline__ uint64_t sel_tail_mask(int64_t w, int64_t nbits) {
  const int64_t left = nbits - w * 64;
  if (left >= 64) return ~0ull;
  if (left <= 0) return 0ull;
  return (1ull << left) - 1ull;
}

// code of row r: bit (63 - r%64) of each of the bw plane words of group r/64
__device__ __forceinline__ uint32_t code_at(const uint64_t* __restrict__ planes, int bw, uint32_t r) {
  const uint64_t* g = planes + (uint64_t)(r >> 6) * bw;
  const int sh = 63 - (int)(r & 63);
  uint32_t v = 0;
  for (int i = 0; i < bw; ++i) v |= (uint32_t)((ldg_gather64(g + i) >> sh) & 1ull) << i;
  return v;
}

// out[j] = code(ids[j])            (dict == nullptr)
// out[j] = dict[code(ids[j])]      (elem_size 4 or 8; codes >= ndict give 0, as GetValue()==false)
__global__ void __launch_bounds__(256)
gather_kernel(const uint64_t* __restrict__ planes, int bw, const uint32_t* __restrict__ ids,
              const uint64_t* __restrict__ count_dev, int64_t cap, const void* __restrict__ dict, int ndict,
              int elem_size, void* __restrict__ out) {
  const int64_t n = min((int64_t)*count_dev, cap);
  for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j < n; j += (int64_t)gridDim.x * 256) {
    const uint32_t c = code_at(planes, bw, ids[j]);
    if (dict == nullptr) {
      reinterpret_cast<uint32_t*>(out)[j] = c;
    } else if (elem_size == 4) {
      reinterpret_cast<uint32_t*>(out)[j] = c < (uint32_t)ndict ? reinterpret_cast<const uint32_t*>(dict)[c] : 0u;
    } else {
      reinterpret_cast<uint64_t*>(out)[j] = c < (uint32_t)ndict ? reinterpret_cast<const uint64_t*>(dict)[c] : 0ull;
    }
  }
}

// Fused Q6-style aggregate.  A warp takes 32 consecutive groups (one bitmap word per lane), then
// walks the survivors of those 2048 rows in a FLATTENED order -- survivor t of the warp goes to
// lane t % 32 -- so that every lane has one independent random load of plain[] in flight per
// step no matter how the survivors are distributed over the groups (a thread-per-group loop is
// latency-bound on that load and diverges with the popcount).  Exact int64 arithmetic;
// per-block partials are added with 64-bit atomics.
__device__ __forceinline__ int nth_set_bit64(uint64_t x, int n) {  // position of the n-th (0-based) set bit
  const uint32_t lo = (uint32_t)x;
  const int clo = __popc(lo);
  if (n < clo) return (int)__fns(lo, 0, n + 1);
  return 32 + (int)__fns((uint32_t)(x >> 32), 0, n - clo + 1);
}

__global__ void __launch_bounds__(256)
sum_product_kernel(const uint64_t* __restrict__ bm, int64_t nrows, const int64_t* __restrict__ plain,
                   const uint64_t* __restrict__ planes, int bw, const int64_t* __restrict__ dict, int ndict,
                   unsigned long long* __restrict__ sum_out, unsigned long long* __restrict__ count_out) {
  __shared__ long long warp_sum[8];
  __shared__ unsigned long long warp_cnt[8];
  __shared__ uint32_t prefix_s[8][33];
  __shared__ uint64_t word_s[8][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t nwords = (nrows + 63) >> 6;
  const int64_t nwarps = ((int64_t)gridDim.x * 256) >> 5;
  long long s = 0;
  unsigned long long c = 0;
  for (int64_t g0 = ((((int64_t)blockIdx.x * 256) >> 5) + wid) * 32; g0 < nwords; g0 += nwarps * 32) {
    const int64_t g = g0 + lane;
    const uint64_t x = g < nwords ? (bm[g] & sel_tail_mask(g, nrows)) : 0ull;
    const uint32_t mine = __popcll(x);
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) continue;
    __syncwarp();
    prefix_s[wid][lane] = incl - mine;  // exclusive
    if (lane == 31) prefix_s[wid][32] = total;
    word_s[wid][lane] = x;
    __syncwarp();
    c += mine;
    for (uint32_t t = lane; t < total; t += 32) {
      // owner lane j: prefix[j] <= t < prefix[j+1]
      int j = 0;
#pragma unroll
      for (int step = 16; step > 0; step >>= 1)
        if (prefix_s[wid][j + step] <= t) j += step;
      const int b = nth_set_bit64(word_s[wid][j], (int)(t - prefix_s[wid][j]));  // row b of group g0+j
      const int64_t gg = g0 + j;
      const long long pv = (long long)ldg_gather64(plain + gg * 64 + b);
      const uint64_t* gp = planes + (uint64_t)gg * bw;
      const int sh = 63 - b;
      uint32_t code = 0;
#pragma unroll 4
      for (int i = 0; i < bw; ++i) code |= (uint32_t)((ldg_gather64(gp + i) >> sh) & 1ull) << i;
      const long long d = code < (uint32_t)ndict ? __ldg(dict + code) : 0;
      s += pv * d;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if (lane == 0) {
    warp_sum[wid] = s;
    warp_cnt[wid] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long ts = 0;
    unsigned long long tc = 0;
    for (int i = 0; i < 8; ++i) { ts += warp_sum[i]; tc += warp_cnt[i]; }
    if (tc) {
      atomicAdd(sum_out, (unsigned long long)ts);  // two's complement: exact for signed sums too
      if (count_out) atomicAdd(count_out, tc);
    }
  }
}

int launch_gather(fle_ctx* ctx, const uint64_t* planes_dev, int bw, const uint32_t* ids_dev,
                  const uint64_t* count_dev, int64_t cap, const void* dict_dev, int ndict, int elem_size,
                  void* out_dev) {
  if (bw < 1 || bw > 32) return set_error(ctx, FLE_E_INVALID, "bit width outside 1..32");
  if (!planes_dev || !ids_dev || !count_dev || !out_dev || cap < 0)
    return set_error(ctx, FLE_E_INVALID, "bad gather arguments");
  if (dict_dev && elem_size != 4 && elem_size != 8) return set_error(ctx, FLE_E_INVALID, "dictionary elements must be 4 or 8 bytes");
  if (cap == 0) return FLE_OK;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * 8, (cap + 255) / 256));
  gather_kernel<<<grid, 256, 0, ctx->stream>>>(planes_dev, bw, ids_dev, count_dev, cap, dict_dev, ndict, elem_size, out_dev);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "gather launch");
  ctx->launches += 1;
  return FLE_OK;
}

int launch_sum_product(fle_ctx* ctx, const uint64_t* bitmap_dev, int64_t nrows, const int64_t* plain_dev,
                       const uint64_t* planes_dev, int bw, const int64_t* dict_dev, int ndict, int64_t* sum_dev,
                       uint64_t* count_dev) {
  if (bw < 1 || bw > 32) return set_error(ctx, FLE_E_INVALID, "bit width outside 1..32");
  if (nrows < 0 || !sum_dev || (nrows > 0 && (!bitmap_dev || !plain_dev || !planes_dev || !dict_dev)))
    return set_error(ctx, FLE_E_INVALID, "bad sum_product arguments");
  FLE_CUDA_TRY(ctx, cudaMemsetAsync(sum_dev, 0, 8, ctx->stream));
  if (count_dev) FLE_CUDA_TRY(ctx, cudaMemsetAsync(count_dev, 0, 8, ctx->stream));
  if (nrows == 0) return FLE_OK;
  const int64_t nwords = (nrows + 63) / 64;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * 8, (nwords + 255) / 256));
  sum_product_kernel<<<grid, 256, 0, ctx->stream>>>(bitmap_dev, nrows, plain_dev, planes_dev, bw, dict_dev, ndict,
                                                   reinterpret_cast<unsigned long long*>(sum_dev),
                                                   reinterpret_cast<unsigned long long*>(count_dev));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "sum_product launch");
  ctx->launches += 1;
  return FLE_OK;
}

}  // namespace fle
