// Selection-bitmap utilities.
//   bitmap_count     = boost::dynamic_bitset::count()
//   intersect_bitset = ColumnReader<T>::IntersectBitset
//                      /root/reference/hdfs-parquet-scanner.cc:326-331
// IntersectBitset walks root bit by bit and hands the i-th SET bit the i-th bit of sub; on
// the device that is a prefix popcount over root's words followed, per word, by a bit
// gather from sub (64-bit funnel shift across two words) and a bit deposit into the set
// positions of the root word.
#include "fle_bitmap.cuh"
#include "fle_internal.h"

namespace fle {

// Non-null rows per nullable column = set bits of its def-level stream (1-bit FLE planes: row k of a
// group at bit 63 - k; def level 1 == present, hdfs-parquet-scanner.cc:890-899), all columns in one
// launch (blockIdx.y = column).  Impala's writer marks every column OPTIONAL, so this runs on every
// real scan and usually finds no NULL at all: it only counts, nothing is written.
struct DefCountParams {
  const uint64_t* planes[FLE_MAX_COLS];
  int64_t ngroups, nrows;
  unsigned long long* counts;
  unsigned long long* nulls_total;  // optional: rows whose level is 0, summed over all columns
};

__global__ void __launch_bounds__(kBmThreads) def_count_kernel(const __grid_constant__ DefCountParams P) {
  __shared__ uint64_t warp_buf[kBmThreads / 32];
  const uint64_t* pl = P.planes[blockIdx.y];
  if (!pl) return;
  uint64_t c = 0, rows = 0;
  for (int64_t g = (int64_t)blockIdx.x * kBmThreads + threadIdx.x; g < P.ngroups; g += (int64_t)gridDim.x * kBmThreads) {
    uint64_t w = ldg_stream64(pl + g);
    const int64_t left = P.nrows - g * 64;
    if (left < 64) w &= ~0ull << (64 - left);  // (left >= 1 for every g < ngroups)
    c += __popcll(w);
    rows += left < 64 ? left : 64;
  }
  const uint64_t t = block_sum<kBmThreads>(c, warp_buf);
  if (P.nulls_total) {
    __syncthreads();
    const uint64_t r = block_sum<kBmThreads>(rows, warp_buf);
    if (threadIdx.x == 0 && r != t) atomicAdd(P.nulls_total, (unsigned long long)(r - t));
    return;
  }
  if (threadIdx.x == 0 && t) atomicAdd(&P.counts[blockIdx.y], (unsigned long long)t);
}

int launch_def_counts(fle_ctx* ctx, const fle_col_desc* cols, int ncols, int64_t n, uint64_t* counts_dev, uint64_t* nulls_total_dev) {
  if (ncols < 1 || ncols > FLE_MAX_COLS) return set_error(ctx, FLE_E_INVALID, "bad column count");
  DefCountParams P;
  P.ngroups = (n + 63) / 64;
  P.nrows = n;
  P.counts = reinterpret_cast<unsigned long long*>(counts_dev);
  P.nulls_total = reinterpret_cast<unsigned long long*>(nulls_total_dev);
  for (int c = 0; c < FLE_MAX_COLS; ++c) P.planes[c] = nullptr;
  for (int c = 0; c < ncols; ++c) {
    if (cols[c].max_def_level == 0) continue;
    if (!cols[c].def_planes) return set_error(ctx, FLE_E_INVALID, "nullable column without def_planes");
    if (reinterpret_cast<uintptr_t>(cols[c].def_planes) & 7) return set_error(ctx, FLE_E_ALIGN, "def_planes must be 8-byte aligned");
    if (cols[c].def_bytes < (uint64_t)P.ngroups * 8u) return set_error(ctx, FLE_E_BUFFER, "def-level stream shorter than the rows requested");
    P.planes[c] = reinterpret_cast<const uint64_t*>(cols[c].def_planes);
  }
  if (nulls_total_dev) FLE_CUDA_TRY(ctx, cudaMemsetAsync(nulls_total_dev, 0, 8, ctx->stream));
  else FLE_CUDA_TRY(ctx, cudaMemsetAsync(counts_dev, 0, (size_t)ncols * 8, ctx->stream));
  if (P.ngroups == 0) return FLE_OK;
  const int gx = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * 4, (P.ngroups + kBmThreads - 1) / kBmThreads));
  def_count_kernel<<<dim3(gx, ncols), kBmThreads, 0, ctx->stream>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "def_count launch");
  ctx->launches += 1;
  return FLE_OK;
}

// per-block popcounts (block b covers words [b*kBmWordsPerBlock, ...))
__global__ void __launch_bounds__(kBmThreads)
bitmap_block_counts_kernel(const uint64_t* __restrict__ bm, int64_t nbits, int64_t nwords,
                           uint64_t* __restrict__ block_sums) {
  __shared__ uint64_t warp_buf[kBmThreads / 32];
  const int64_t w0 = (int64_t)blockIdx.x * kBmWordsPerBlock;
  uint64_t c = 0;
#pragma unroll
  for (int k = 0; k < kBmWordsPerThread; ++k) {
    const int64_t w = w0 + k * kBmThreads + threadIdx.x;
    if (w < nwords) c += __popcll(bm[w] & tail_mask_word(w, nbits));
  }
  const uint64_t t = block_sum<kBmThreads>(c, warp_buf);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = t;
}

// 64 bits of sub starting at bit offset `pos` (bits past n_sub read as 0)
__device__ __forceinline__ uint64_t gather_bits(const uint64_t* __restrict__ sub, int64_t n_sub,
                                                int64_t pos) {
  const int64_t nw = (n_sub + 63) >> 6;
  const int64_t w = pos >> 6;
  const int sh = (int)(pos & 63);
  uint64_t lo = w < nw ? (sub[w] & tail_mask_word(w, n_sub)) : 0ull;
  if (sh == 0) return lo;
  uint64_t hi = (w + 1) < nw ? (sub[w + 1] & tail_mask_word(w + 1, n_sub)) : 0ull;
  return (lo >> sh) | (hi << (64 - sh));
}

__global__ void __launch_bounds__(kBmThreads)
intersect_kernel(uint64_t* __restrict__ root, int64_t n_root, int64_t nwords,
                 const uint64_t* __restrict__ sub, int64_t n_sub,
                 const uint64_t* __restrict__ block_sums) {
  __shared__ uint64_t warp_buf[kBmThreads / 32];
  __shared__ uint32_t warp_scan[kBmThreads / 32];
  // exclusive offset of this block = sum of the popcounts of all earlier blocks
  uint64_t pre = 0;
  for (int i = threadIdx.x; i < (int)blockIdx.x; i += kBmThreads) pre += block_sums[i];
  const uint64_t block_off = block_sum<kBmThreads>(pre, warp_buf);

  // thread owns kBmWordsPerThread CONSECUTIVE words so that its ranks are contiguous
  const int64_t w0 = (int64_t)blockIdx.x * kBmWordsPerBlock + (int64_t)threadIdx.x * kBmWordsPerThread;
  uint64_t m[kBmWordsPerThread];
  uint32_t mine = 0;
#pragma unroll
  for (int k = 0; k < kBmWordsPerThread; ++k) {
    const int64_t w = w0 + k;
    m[k] = w < nwords ? (root[w] & tail_mask_word(w, n_root)) : 0ull;
    mine += __popcll(m[k]);
  }
  // block-wide exclusive scan of `mine`
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += y;
  }
  if ((threadIdx.x & 31) == 31) warp_scan[threadIdx.x >> 5] = incl;
  __syncthreads();
  uint32_t warp_off = 0;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) warp_off += warp_scan[w];
  int64_t rank = (int64_t)block_off + warp_off + (incl - mine);
#pragma unroll
  for (int k = 0; k < kBmWordsPerThread; ++k) {
    const int64_t w = w0 + k;
    if (w < nwords) {
      const uint64_t bits = gather_bits(sub, n_sub, rank);
      root[w] = deposit_bits(bits, m[k]);
      rank += __popcll(m[k]);
    }
  }
}

// Deposit straight from a column's definition levels (max_def_level == 1: the 1-bit FLE plane word of a group,
// bit-reversed, IS the group's not-null word), out of place and combined with what the output already holds:
//     out[w] = out[w] (op) deposit(not-null word w, the next popcount(not-null word) bits of sub)
// -- one nullable column subtree of a conjunct chain per launch pair, with no not-null bitmap, no copy of it and no
// separate combine pass (hdfs-parquet-scanner.cc:326-345 per leaf, :1857-1862 for the chain).
__global__ void __launch_bounds__(kBmThreads)
def_block_counts_kernel(const uint64_t* __restrict__ def_planes, int64_t n_root, int64_t nwords, uint64_t* __restrict__ block_sums) {
  __shared__ uint64_t warp_buf[kBmThreads / 32];
  const int64_t w0 = (int64_t)blockIdx.x * kBmWordsPerBlock;
  uint64_t c = 0;
#pragma unroll
  for (int k = 0; k < kBmWordsPerThread; ++k) c += __popcll(def_word(def_planes, w0 + k * kBmThreads + threadIdx.x, nwords, n_root));
  const uint64_t t = block_sum<kBmThreads>(c, warp_buf);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = t;
}

__global__ void __launch_bounds__(kBmThreads)
deposit_kernel(const uint64_t* __restrict__ def_planes, int64_t n_root, int64_t nwords, const uint64_t* __restrict__ sub,
               int64_t n_sub, const uint64_t* __restrict__ block_sums, uint64_t* __restrict__ out, int comb_op,
               unsigned long long* __restrict__ count) {
  __shared__ uint64_t warp_buf[kBmThreads / 32];
  __shared__ uint32_t warp_scan[kBmThreads / 32];
  uint64_t pre = 0;
  for (int i = threadIdx.x; i < (int)blockIdx.x; i += kBmThreads) pre += block_sums[i];
  const uint64_t block_off = block_sum<kBmThreads>(pre, warp_buf);
  const int64_t w0 = (int64_t)blockIdx.x * kBmWordsPerBlock + (int64_t)threadIdx.x * kBmWordsPerThread;
  uint64_t m[kBmWordsPerThread];
  uint32_t mine = 0;
#pragma unroll
  for (int k = 0; k < kBmWordsPerThread; ++k) {
    m[k] = def_word(def_planes, w0 + k, nwords, n_root);
    mine += __popcll(m[k]);
  }
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += y;
  }
  if ((threadIdx.x & 31) == 31) warp_scan[threadIdx.x >> 5] = incl;
  __syncthreads();
  uint32_t warp_off = 0;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) warp_off += warp_scan[w];
  int64_t rank = (int64_t)block_off + warp_off + (incl - mine);
  uint64_t cnt = 0;
#pragma unroll
  for (int k = 0; k < kBmWordsPerThread; ++k) {
    const int64_t w = w0 + k;
    if (w < nwords) {
      uint64_t r = deposit_bits(gather_bits(sub, n_sub, rank), m[k]);
      if (comb_op == FOLD_AND) r &= out[w];
      else if (comb_op == FOLD_OR) r |= out[w] & tail_mask_word(w, n_root);
      out[w] = r;
      cnt += __popcll(r);
      rank += __popcll(m[k]);
    }
  }
  if (count) {
    __syncthreads();
    const uint64_t t = block_sum<kBmThreads>(cnt, warp_buf);
    if (threadIdx.x == 0 && t) atomicAdd(count, (unsigned long long)t);
  }
}

int launch_deposit(fle_ctx* ctx, const uint64_t* def_planes_dev, int64_t n_root, const uint64_t* sub_dev, int64_t n_sub,
                   uint64_t* out_dev, int comb_op, uint64_t* count_dev) {
  if (n_root < 0 || n_sub < 0 || !def_planes_dev || !out_dev || (n_sub > 0 && !sub_dev))
    return set_error(ctx, FLE_E_INVALID, "bad deposit arguments");
  if (count_dev) FLE_CUDA_TRY(ctx, cudaMemsetAsync(count_dev, 0, 8, ctx->stream));
  if (n_root == 0) return FLE_OK;
  const int64_t nwords = (n_root + 63) / 64;
  const int nblocks = (int)((nwords + kBmWordsPerBlock - 1) / kBmWordsPerBlock);
  int rc = ensure_scratch(ctx, nblocks, 0);
  if (rc != FLE_OK) return rc;
  def_block_counts_kernel<<<nblocks, kBmThreads, 0, ctx->stream>>>(def_planes_dev, n_root, nwords, ctx->d_partial);
  deposit_kernel<<<nblocks, kBmThreads, 0, ctx->stream>>>(def_planes_dev, n_root, nwords, sub_dev, n_sub, ctx->d_partial, out_dev,
                                                         comb_op, reinterpret_cast<unsigned long long*>(count_dev));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "deposit launch");
  ctx->launches += 2;
  return FLE_OK;
}

// popcount of a bitmap in ONE launch: grid-stride, eight independent 8-byte loads in flight per thread, one atomic per CTA
// (two launches -- block counts, then their sum -- cost 19.5 us for the 32 MB bitmap of 2^28 rows, most of it launch gaps)
__global__ void __launch_bounds__(kBmThreads)
bitmap_count_kernel(const uint64_t* __restrict__ bm, int64_t nbits, int64_t nwords, unsigned long long* __restrict__ count) {
  __shared__ uint64_t warp_buf[kBmThreads / 32];
  const int64_t stride = (int64_t)gridDim.x * kBmThreads;
  uint64_t c = 0;
  // every pass has its eight loads in flight at once, the ragged last pass included (a scalar remainder loop is a
  // chain of dependent DRAM round trips: 5 us of a 16 us kernel)
  for (int64_t w = (int64_t)blockIdx.x * kBmThreads + threadIdx.x; w < nwords; w += 8 * stride) {
    uint64_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = ldg_stream64(bm + min(w + j * stride, nwords - 1));  // (unconditional: the loads batch)
#pragma unroll
    for (int j = 0; j < 8; ++j) c += (w + j * stride < nwords) ? __popcll(v[j]) : 0;
  }
  // bits past the end of the last word do not count
  if (blockIdx.x == 0 && threadIdx.x == 0) c -= __popcll(bm[nwords - 1] & ~tail_mask_word(nwords - 1, nbits));
  const uint64_t t = block_sum<kBmThreads>(c, warp_buf);
  if (threadIdx.x == 0 && t) atomicAdd(count, (unsigned long long)t);
}

int launch_bitmap_count(fle_ctx* ctx, const uint64_t* bitmap_dev, int64_t nbits, uint64_t* count_dev) {
  if (nbits < 0 || !count_dev) return set_error(ctx, FLE_E_INVALID, "bad bitmap_count arguments");
  FLE_CUDA_TRY(ctx, cudaMemsetAsync(count_dev, 0, 8, ctx->stream));
  if (nbits == 0) return FLE_OK;
  const int64_t nwords = (nbits + 63) / 64;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * 8, (nwords + kBmThreads * 8 - 1) / (kBmThreads * 8)));
  bitmap_count_kernel<<<grid, kBmThreads, 0, ctx->stream>>>(bitmap_dev, nbits, nwords, reinterpret_cast<unsigned long long*>(count_dev));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "bitmap_count launch");
  ctx->launches += 1;
  return FLE_OK;
}

int launch_intersect(fle_ctx* ctx, uint64_t* root_dev, int64_t n_root, const uint64_t* sub_dev,
                     int64_t n_sub) {
  if (n_root < 0 || n_sub < 0 || !root_dev || (n_sub > 0 && !sub_dev))
    return set_error(ctx, FLE_E_INVALID, "bad intersect arguments");
  if (n_root == 0) return FLE_OK;
  const int64_t nwords = (n_root + 63) / 64;
  const int nblocks = (int)((nwords + kBmWordsPerBlock - 1) / kBmWordsPerBlock);
  int rc = ensure_scratch(ctx, nblocks, 0);
  if (rc != FLE_OK) return rc;
  // NB: block_counts assigns words to threads interleaved, intersect consecutively; both
  // cover the same words per BLOCK, which is all the block sums need.
  bitmap_block_counts_kernel<<<nblocks, kBmThreads, 0, ctx->stream>>>(root_dev, n_root, nwords, ctx->d_partial);
  intersect_kernel<<<nblocks, kBmThreads, 0, ctx->stream>>>(root_dev, n_root, nwords, sub_dev, n_sub, ctx->d_partial);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "intersect launch");
  ctx->launches += 2;
  return FLE_OK;
}

}  // namespace fle
