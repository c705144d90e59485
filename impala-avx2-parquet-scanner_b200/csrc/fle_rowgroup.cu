// Row-group batches for the drop-in scanner (host/exec/hdfs-parquet-scanner.h): what AssembleRows
// (/root/reference/hdfs-parquet-scanner.cc:1078-1275) does for ONE row group, as three device steps
// that together need a handful of launches however many data pages the column chunks hold:
//
//   pages_gather_kernel  every data page's FLE section -- at whatever byte offset it sits inside the
//                        uploaded column-chunk image (page headers, def levels and the width byte
//                        precede it: dict-encoding.h:185-192, hdfs-parquet-scanner.cc:882-919) and
//                        with whatever bit width the page was written with (the reference's writer
//                        picks the width per page, dict-encoding.h:438-444) -- is funnel-shifted into
//                        ONE aligned plane array per column at the column's widest width (missing high
//                        planes are zero planes, which leaves every code unchanged); the pages'
//                        definition levels are counted on the way (rows with def level 0 = NULLs).
//   select1_kernel       bitmap -> ascending row ids in ONE launch: decoupled look-back over 65536-row
//                        tiles (the skip-distance list of :1134-1148).
//   assemble_kernel      the same scan, but instead of row ids the survivors' TUPLES are written:
//                        code of every column (FleDecoder::Get(T*, skip), fle-encoding.h:344-379) ->
//                        dictionary value (DictDecoder::GetValue, dict-encoding.h:321-330) -> slot at its
//                        tuple offset (ReadSlot, hdfs-parquet-scanner.cc:515-531), column by column as
//                        :1151-1168 does -- select + gather + lookup + tuple layout in one pass.
#include "fle_bitmap.cuh"
#include "fle_device.cuh"
#include "fle_internal.h"

namespace fle {

// ------------------------------------------------------------------------------------------
// pages -> planes
// ------------------------------------------------------------------------------------------
struct PageRefDev {
  const uint8_t* src;      // FLE groups of the page, any byte alignment
  uint64_t* dst;           // 8-byte aligned: the destination group that receives the page's first row
  const uint8_t* def_src;  // 1-bit FLE definition levels of the page, or null
  int64_t groups;          // 64-row groups of the page
  int64_t rows;            // rows of the page
  int32_t src_bw, dst_bw;
  int64_t data_groups;     // groups the page's data stream holds (expand: the non-null values only)
  int64_t dst_groups;      // destination groups the page touches: ceil((dst_bit + rows) / 64)
  int32_t general;         // 0: aligned copy (no NULL expansion, first row on a group boundary, whole groups)
  int32_t dst_bit;         // position of the page's first row inside its first destination group
  int32_t expand;
  int32_t plain;           // 0: FLE page; 1 / 2 / 4 / 8: PLAIN page of that many bytes per value (dst = value slots)
  const uint8_t* dict;     // plain != 0 and set: the page is dictionary-encoded FLE after all, its codes are looked up here
  int32_t ndict, pad2;
};

constexpr int kGatherGroupsPerCta = 256;

// 64 bits at byte address p (any alignment) out of the two aligned words around it (up to 7 bytes
// after the value are touched: column-chunk images are uploaded with 16 bytes of slack)
__device__ __forceinline__ uint64_t load_u64_unaligned(const uint8_t* p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint64_t* q = reinterpret_cast<const uint64_t*>(a & ~(uintptr_t)7);
  const uint32_t sh = (uint32_t)(a & 7) * 8u;
  const uint64_t lo = __ldg(q);
  if (sh == 0) return lo;
  const uint64_t hi = __ldg(q + 1);
  return (lo >> sh) | (hi << (64u - sh));
}

// Destination groups that a page shares with its neighbours (its first group when the page starts inside a group,
// its last one when it ends inside a group) are cleared first: the pages OR their rows into them.
__global__ void __launch_bounds__(64) pages_edges_kernel(const PageRefDev* __restrict__ refs) {
  const PageRefDev R = refs[blockIdx.x];
  if (!R.general || R.plain || R.dst_groups <= 0) return;
  const int t = threadIdx.x & 31;
  if (t >= R.dst_bw) return;
  if (threadIdx.x < 32) {
    if (R.dst_bit != 0) R.dst[t] = 0ull;
  } else {
    if (((int64_t)R.dst_bit + R.rows) & 63) R.dst[(uint64_t)(R.dst_groups - 1) * R.dst_bw + t] = 0ull;
  }
}

// grid = (tiles of the largest page, pages).
// Aligned pages (whole groups, first row on a group boundary, no NULL expansion): a CTA copies 256 groups of one
// page; the page's plane words are consecutive in the source, so word w of the tile goes to thread w % 256
// (coalesced 8-byte reads); the destination word of source word (g, i) is g * dst_bw + i.
// Every other page is PLACED row by row: one thread per destination group.  The rows of the page that fall into
// the group, their definition bits (all ones for a page without definition levels) and the rank of the first of
// them among the page's defined rows (prefix popcount inside the CTA, the words in front of its 256 counted on the
// way) say which values of the data stream -- indexed by non-null ordinal, hdfs-parquet-scanner.cc:886-921 -- go
// where: per plane a funnel shift takes them out of the stream and a bit deposit spreads them over the defined
// rows; NULL rows read as code 0.  Pages of any row count, starting at any row of the column, with or without
// NULLs, thereby join ONE row-aligned plane array per column.
__global__ void __launch_bounds__(256)
pages_gather_kernel(const PageRefDev* __restrict__ refs, unsigned long long* __restrict__ null_rows) {
  const PageRefDev R = refs[blockIdx.y];
  const int64_t g0 = (int64_t)blockIdx.x * kGatherGroupsPerCta;
  // CTAs of this page that have groups to write (at least one: it also counts the definition levels)
  const int64_t span = R.general ? R.dst_groups : R.groups;
  const int64_t nalive = max((int64_t)1, (span + kGatherGroupsPerCta - 1) / kGatherGroupsPerCta);
  if ((int64_t)blockIdx.x >= nalive) return;
  const int64_t ng = min((int64_t)kGatherGroupsPerCta, R.groups - g0);
  if (R.general) {
    __shared__ uint64_t warp_buf[8];
    __shared__ uint32_t warp_scan[8];
    const int s = R.dst_bit;
    auto def_row_order = [&](int64_t j) -> uint64_t {  // definition bits of page rows [64 j, 64 j + 64), row r at bit r & 63
      if (j < 0 || j >= R.groups) return 0ull;
      uint64_t d = R.expand ? __brevll(load_u64_unaligned(R.def_src + j * 8)) : ~0ull;
      const int64_t left = R.rows - j * 64;
      if (left < 64) d &= (1ull << left) - 1ull;  // (left >= 1 for j < groups)
      return d;
    };
    auto def_bits = [&](int64_t r0) -> uint64_t {  // bit k = defined(page row r0 + k), 0 outside the page
      if (r0 < 0) return def_row_order(0) << (int)(-r0);  // (only the page's first destination group: -r0 = dst_bit)
      const int64_t j = r0 >> 6;
      const int sh = (int)(r0 & 63);
      uint64_t v = def_row_order(j) >> sh;
      if (sh) v |= def_row_order(j + 1) << (64 - sh);
      return v;
    };
    // defined rows of the page in front of this CTA's first destination group
    const int64_t first_r = max((int64_t)0, g0 * 64 - s);
    uint64_t rank0;
    if (R.expand) {
      uint64_t pre = 0;
      for (int64_t j = threadIdx.x; j < (first_r >> 6); j += 256) pre += __popcll(def_row_order(j));
      if (threadIdx.x == 0 && (first_r & 63)) pre += __popcll(def_row_order(first_r >> 6) & ((1ull << (first_r & 63)) - 1ull));
      rank0 = block_sum<256>(pre, warp_buf);
    } else {
      rank0 = (uint64_t)first_r;
    }
    const int64_t wl = g0 + threadIdx.x;  // destination group of the page
    const int64_t base_r = wl * 64 - s;   // page row of the group's first bit (negative: the group starts before the page)
    const uint64_t mask_r = wl < R.dst_groups ? def_bits(base_r) : 0ull;
    const uint32_t mine = __popcll(mask_r);
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += y;
    }
    if ((threadIdx.x & 31) == 31) warp_scan[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int x = 0; x < (int)(threadIdx.x >> 5); ++x) woff += warp_scan[x];
    if (R.plain) {
      // a PLAIN page: the defined rows of the group take the next popcount(mask) values of the page (plain bytes
      // each, at any byte address), NULL rows 0 -- into the column's row-aligned value array
      if (wl < R.dst_groups) {
        int64_t ord = (int64_t)rank0 + woff + (incl - mine);
        const int size = R.plain;
        uint8_t* out = reinterpret_cast<uint8_t*>(R.dst) + (uint64_t)wl * 64u * (uint32_t)size;
        const int nrows = (int)min((int64_t)64, R.rows - wl * 64);
        const int64_t nvals = (int64_t)(R.data_groups);  // (PLAIN: the number of VALUES the page holds)
        // a dictionary page of a column whose later pages are PLAIN (the column overflowed its dictionary inside this
        // row group): the codes of the group's rows sit in at most two groups of the page's stream; value = dict[code]
        uint64_t ga[32], gb[32];
        const int sbw = R.src_bw;
        if (R.dict) {
          const int64_t dg = ord >> 6;
          for (int i = 0; i < sbw; ++i) {
            ga[i] = (mine && dg < R.data_groups) ? load_u64_unaligned(R.src + ((uint64_t)dg * sbw + i) * 8u) : 0ull;
            gb[i] = (mine && dg + 1 < R.data_groups) ? load_u64_unaligned(R.src + ((uint64_t)(dg + 1) * sbw + i) * 8u) : 0ull;
          }
        }
        const int64_t ord_group0 = ord >> 6;
        for (int k = 0; k < nrows; ++k) {
          uint64_t v = 0;
          if ((mask_r >> k) & 1ull) {
            if (R.dict) {
              const uint64_t* g = (ord >> 6) == ord_group0 ? ga : gb;
              const int sh = 63 - (int)(ord & 63);
              uint32_t code = 0;
              for (int i = 0; i < sbw; ++i) code |= (uint32_t)((g[i] >> sh) & 1ull) << i;
              if (code < (uint32_t)R.ndict) {
                const uint8_t* e = R.dict + (uint64_t)code * (uint32_t)size;
                v = size == 8 ? *reinterpret_cast<const uint64_t*>(e)
                              : (size == 4 ? *reinterpret_cast<const uint32_t*>(e)
                                           : (size == 2 ? *reinterpret_cast<const uint16_t*>(e) : *e));
              }
            } else if (ord < nvals) {
              v = load_u64_unaligned(R.src + (uint64_t)ord * (uint32_t)size);
            }
            ++ord;
          }
          switch (size) {
            case 1: out[k] = (uint8_t)v; break;
            case 2: reinterpret_cast<uint16_t*>(out)[k] = (uint16_t)v; break;
            case 4: reinterpret_cast<uint32_t*>(out)[k] = (uint32_t)v; break;
            default: reinterpret_cast<uint64_t*>(out)[k] = v; break;
          }
        }
      }
    } else if (wl < R.dst_groups) {
      const int64_t rank = (int64_t)rank0 + woff + (incl - mine);
      const int64_t dg = rank >> 6;
      const int sh = (int)(rank & 63);
      const int sbw = R.src_bw, dbw = R.dst_bw;
      const bool interior = base_r >= 0 && base_r + 64 <= R.rows;  // every row of the group belongs to this page
      uint64_t* out = R.dst + (uint64_t)wl * dbw;
      const bool whole = mask_r == ~0ull && sh == 0;  // 64 defined rows that are one group of the stream: a copy
      for (int i = 0; i < sbw; ++i) {
        const uint64_t a = (mine && dg < R.data_groups) ? load_u64_unaligned(R.src + ((uint64_t)dg * sbw + i) * 8u) : 0ull;
        uint64_t v = a;
        if (!whole) {
          const uint64_t b = (mine && sh && dg + 1 < R.data_groups) ? load_u64_unaligned(R.src + ((uint64_t)(dg + 1) * sbw + i) * 8u) : 0ull;
          const uint64_t top = sh ? ((a << sh) | (b >> (64 - sh))) : a;  // ordinal rank + q at bit 63 - q
          v = __brevll(deposit_bits(__brevll(top), mask_r));
        }
        if (interior) out[i] = v;
        else if (v) atomicOr(reinterpret_cast<unsigned long long*>(out + i), (unsigned long long)v);  // (cleared by pages_edges_kernel)
      }
      if (interior)
        for (int i = sbw; i < dbw; ++i) out[i] = 0ull;
    }
  } else if (ng > 0) {
    const int sbw = R.src_bw, dbw = R.dst_bw;
    const uint8_t* s = R.src + (uint64_t)g0 * sbw * 8u;
    uint64_t* d = R.dst + (uint64_t)g0 * dbw;
    const int64_t nw = ng * sbw;
    for (int64_t w = threadIdx.x; w < nw; w += 256) {
      const int64_t g = w / sbw;
      const int i = (int)(w - g * sbw);
      d[g * dbw + i] = load_u64_unaligned(s + w * 8);
    }
    if (dbw > sbw) {
      const int pad = dbw - sbw;
      for (int64_t w = threadIdx.x; w < ng * pad; w += 256) {
        const int64_t g = w / pad;
        d[g * dbw + sbw + (int)(w - g * pad)] = 0ull;
      }
    }
  }
  // definition levels (bit width 1: one word per 64 rows, row k at bit 63 - k): count the rows that are NOT defined
  if (R.def_src && null_rows) {
    const int64_t dgroups = (R.rows + 63) / 64;
    unsigned long long nulls = 0;
    for (int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x; g < dgroups; g += nalive * 256) {
      const uint64_t w = load_u64_unaligned(R.def_src + g * 8);
      const int64_t left = R.rows - g * 64;
      const uint64_t want = left >= 64 ? ~0ull : (~0ull << (64 - left));
      nulls += __popcll(~w & want);
    }
    if (__syncthreads_or(nulls != 0)) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) nulls += __shfl_xor_sync(0xffffffffu, nulls, o);
      if ((threadIdx.x & 31) == 0 && nulls) atomicAdd(null_rows, nulls);
    }
  }
}

int launch_pages_gather(fle_ctx* ctx, const fle_page_ref* refs, int nrefs, uint64_t* null_rows_dev) {
  if (nrefs < 0 || (nrefs > 0 && !refs)) return set_error(ctx, FLE_E_INVALID, "bad page list");
  if (null_rows_dev) FLE_CUDA_TRY(ctx, cudaMemsetAsync(null_rows_dev, 0, 8, ctx->stream));
  if (nrefs == 0) return FLE_OK;
  std::vector<PageRefDev> h((size_t)nrefs);
  int64_t max_groups = 1;
  bool any_general = false;
  for (int i = 0; i < nrefs; ++i) {
    const fle_page_ref& r = refs[i];
    const bool plain = r.plain_size != 0;
    if (plain && r.plain_size != 1 && r.plain_size != 2 && r.plain_size != 4 && r.plain_size != 8)
      return set_error(ctx, FLE_E_INVALID, "PLAIN values are 1, 2, 4 or 8 bytes");
    const bool coded = plain && r.dict_dev != nullptr;  // a dictionary page decoded into a value array
    if (coded && (r.src_bit_width < 1 || r.src_bit_width > 32 || r.ndict < 0 || (reinterpret_cast<uintptr_t>(r.dict_dev) & (uintptr_t)(r.plain_size - 1))))
      return set_error(ctx, FLE_E_INVALID, "bad dictionary or bit width of a dictionary page that is decoded into values");
    if (!plain && (r.src_bit_width < 1 || r.src_bit_width > 32 || r.dst_bit_width < r.src_bit_width || r.dst_bit_width > 32))
      return set_error(ctx, FLE_E_INVALID, "page bit widths outside 1..32 or destination narrower than the page");
    if (r.groups < 0 || r.rows < 0 || (r.groups > 0 && ((!r.src_dev && !(r.expand && r.src_bytes == 0)) || !r.dst_dev)))
      return set_error(ctx, FLE_E_INVALID, "bad page reference");
    if (!plain && (r.dst_bit < 0 || r.dst_bit > 63)) return set_error(ctx, FLE_E_INVALID, "dst_bit outside 0..63");
    if (r.expand && !r.def_src_dev) return set_error(ctx, FLE_E_INVALID, "a page is expanded through its definition levels");
    const int64_t rows = (r.rows == 0 && !r.expand && !r.def_src_dev) ? r.groups * 64 : r.rows;  // (callers that only copy groups)
    if (r.groups != (rows + 63) / 64) return set_error(ctx, FLE_E_INVALID, "a page has one group per 64 rows");
    if (!plain && !r.expand && (uint64_t)r.groups * (uint64_t)r.src_bit_width * 8u > r.src_bytes)
      return set_error(ctx, FLE_E_BUFFER, "page shorter than its groups");
    if (plain && !coded && !r.expand && (uint64_t)rows * (uint64_t)r.plain_size > r.src_bytes)
      return set_error(ctx, FLE_E_BUFFER, "page shorter than its values");
    if (coded && !r.expand && (uint64_t)r.groups * (uint64_t)r.src_bit_width * 8u > r.src_bytes)
      return set_error(ctx, FLE_E_BUFFER, "page shorter than its groups");
    if (r.def_src_dev && (uint64_t)((rows + 63) / 64) * 8u > r.def_bytes)
      return set_error(ctx, FLE_E_BUFFER, "definition levels shorter than the page's rows");
    if (reinterpret_cast<uintptr_t>(r.dst_dev) & (plain ? (uintptr_t)(r.plain_size - 1) : (uintptr_t)7))
      return set_error(ctx, FLE_E_ALIGN, "plane destination must be 8-byte aligned (value slots: naturally aligned)");
    h[i].src = static_cast<const uint8_t*>(r.src_dev);
    h[i].dst = static_cast<uint64_t*>(r.dst_dev);
    h[i].def_src = static_cast<const uint8_t*>(r.def_src_dev);
    h[i].groups = r.groups;
    h[i].rows = rows;
    h[i].dict = coded ? static_cast<const uint8_t*>(r.dict_dev) : nullptr;
    h[i].ndict = coded ? r.ndict : 0;
    h[i].pad2 = 0;
    h[i].src_bw = (plain && !coded) ? 1 : r.src_bit_width;
    h[i].dst_bw = plain ? 1 : r.dst_bit_width;
    h[i].expand = r.expand ? 1 : 0;
    h[i].plain = r.plain_size;
    h[i].dst_bit = plain ? 0 : r.dst_bit;
    h[i].dst_groups = (h[i].dst_bit + rows + 63) / 64;
    // a page that ends inside a group is placed as well (the next page ORs its rows into that group) -- unless it is
    // copied to a group boundary and nothing was said about rows at all
    h[i].general = (plain || r.expand || r.dst_bit != 0 || (rows & 63) != 0) ? 1 : 0;
    // (without expansion the definition levels are only counted: the data stream then holds every row)
    h[i].data_groups = (plain && !coded) ? (int64_t)(r.src_bytes / (uint64_t)r.plain_size)  // (values, not groups)
                             : (r.expand ? (int64_t)(r.src_bytes / ((uint64_t)r.src_bit_width * 8u)) : r.groups);
    any_general = any_general || (h[i].general && !plain && (r.dst_bit != 0 || ((r.dst_bit + rows) & 63) != 0));  // groups shared with a neighbour
    max_groups = std::max(max_groups, h[i].general ? h[i].dst_groups : r.groups);
  }
  const uint32_t* d = nullptr;
  int rc = upload_literals(ctx, h.data(), (int)(h.size() * sizeof(PageRefDev) / 4), nullptr, nullptr, &d);
  if (rc != FLE_OK) return rc;
  if (any_general) {
    pages_edges_kernel<<<nrefs, 64, 0, ctx->stream>>>(reinterpret_cast<const PageRefDev*>(d));
    ctx->launches += 1;
  }
  const dim3 grid((unsigned)((max_groups + kGatherGroupsPerCta - 1) / kGatherGroupsPerCta), (unsigned)nrefs);
  pages_gather_kernel<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<const PageRefDev*>(d),
                                                    reinterpret_cast<unsigned long long*>(null_rows_dev));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "pages_gather launch");
  ctx->launches += 1;
  return FLE_OK;
}

// ------------------------------------------------------------------------------------------
// single-pass ordered compaction of a bitmap (decoupled look-back)
// ------------------------------------------------------------------------------------------
constexpr int kTileThreads = 256;
// words per thread: 4 (65536 rows per tile) for long bitmaps; 1 (16384 rows per tile) where that would leave most SMs
// without a tile -- a 4 Mi-row row group is 64 tiles of 65536 rows on 148 SMs, and a thread then walks ~5 survivors with
// a dozen dependent gathers each (assemble: 80 us); with 256 tiles the same work is spread four times wider

// tile status words: [63:40] launch generation, [39:38] 1 = aggregate / 2 = inclusive prefix, [37:0] value.
// The generation makes entries of earlier launches read as "not there yet": no reset between launches.
constexpr uint64_t kTileAgg = 1ull << 38, kTilePrefix = 2ull << 38, kTileValMask = (1ull << 38) - 1ull;

struct SlotDev {  // one materialised column
  const uint64_t* planes;
  const void* dict;
  int32_t bw, ndict, slot_size, tuple_off;
  const uint64_t* def;  // optional 1-bit definition levels, row-aligned
  int32_t null_off, pad;
  const uint8_t* plain;  // optional: row-aligned PLAIN values (slot_size bytes each) instead of planes + dictionary
};

struct TileScanParams {
  const uint64_t* bm;
  int64_t nbits, nwords;
  int32_t ntiles;
  uint32_t gen;           // 24 bits, never 0
  uint64_t* tile_state;   // ntiles words
  unsigned int* ticket;   // [0] completion ticket, [1] tile counter (both self-resetting)
  uint64_t* count_out;
  // select: row ids
  uint32_t* ids;
  int64_t cap;
  // assemble: tuples
  uint8_t* tuples;
  int32_t tuple_size, ncols;
  unsigned long long* bad_codes;
  SlotDev cols[FLE_MAX_COLS];
};

__device__ __forceinline__ uint64_t tile_tail_mask(int64_t w, int64_t nbits) {
  const int64_t left = nbits - w * 64;
  if (left >= 64) return ~0ull;
  if (left <= 0) return 0ull;
  return (1ull << left) - 1ull;
}

__device__ __forceinline__ uint32_t code_of_row(const uint64_t* __restrict__ planes, int bw, uint32_t r) {
  const uint64_t* g = planes + (uint64_t)(r >> 6) * bw;
  const int sh = 63 - (int)(r & 63);
  uint32_t v = 0;
  for (int i = 0; i < bw; ++i) v |= (uint32_t)((ldg_gather64(g + i) >> sh) & 1ull) << i;
  return v;
}

template <bool ASSEMBLE, int kTileWordsPerThread>
__global__ void __launch_bounds__(kTileThreads) tile_scan_kernel(const __grid_constant__ TileScanParams P) {
  constexpr int kTileWords = kTileThreads * kTileWordsPerThread;
  __shared__ uint32_t warp_scan[kTileThreads / 32];
  __shared__ int tile_s;
  __shared__ uint64_t base_s;
  __shared__ bool last_s;
  const uint64_t gen = (uint64_t)P.gen << 40;
  for (;;) {
    // tiles are claimed in launch order, so every predecessor of a tile is running or done: the
    // look-back below cannot wait for a CTA that has not been scheduled
    if (threadIdx.x == 0) tile_s = (int)atomicAdd(P.ticket + 1, 1u);
    __syncthreads();
    const int tile = tile_s;
    if (tile >= P.ntiles) break;
    const int64_t w0 = (int64_t)tile * kTileWords + (int64_t)threadIdx.x * kTileWordsPerThread;
    uint64_t m[kTileWordsPerThread];
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < kTileWordsPerThread; ++k) {
      const int64_t w = w0 + k;
      m[k] = w < P.nwords ? (P.bm[w] & tile_tail_mask(w, P.nbits)) : 0ull;
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
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kTileThreads / 32; ++w) {
      if (w < (int)(threadIdx.x >> 5)) woff += warp_scan[w];
      total += warp_scan[w];
    }
    if (threadIdx.x < 32) {  // first warp: publish, look back, publish the inclusive prefix
      const int lane = threadIdx.x;
      volatile uint64_t* st = P.tile_state;
      if (lane == 0) st[tile] = gen | (tile == 0 ? kTilePrefix : kTileAgg) | (uint64_t)total;
      uint64_t excl = 0;
      int idx = tile - 1;
      // look-back over windows of 32 predecessors, nearest first.  (Measured and dropped: eight windows fetched per
      // round trip -- 75 us instead of 47 us at 2^28 rows, the extra polling of the status words slows the
      // publishing CTAs down -- and tiles of 262144 rows, 64 us.  The kernel is latency-bound at that size: 33 MB
      // of bitmap are 5 us of HBM time; at the scanner's 4 Mi-row row groups it is a few microseconds.)
      while (idx >= 0) {
        const int j = idx - lane;
        uint64_t s = gen | kTilePrefix;  // lanes past the front of the array: a prefix of 0
        if (j >= 0) {
          do {
            s = st[j];
          } while ((s >> 40) != (gen >> 40) || (s & (kTileAgg | kTilePrefix)) == 0);
        }
        const unsigned has_prefix = __ballot_sync(0xffffffffu, (s & kTilePrefix) != 0);
        // the nearest predecessor that already knows its inclusive prefix closes the look-back; without
        // one all 32 aggregates are added and the window moves on
        const int first = has_prefix ? __ffs(has_prefix) - 1 : 31;
        uint64_t v = lane <= first ? (s & kTileValMask) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl += v;
        if (has_prefix) break;
        idx -= 32;
      }
      if (lane == 0) {
        if (tile != 0) st[tile] = gen | kTilePrefix | (excl + total);
        base_s = excl;
        if (tile == P.ntiles - 1 && P.count_out) *P.count_out = excl + total;
      }
    }
    __syncthreads();
    int64_t pos = (int64_t)base_s + woff + (incl - mine);
#pragma unroll
    for (int k = 0; k < kTileWordsPerThread; ++k) {
      uint64_t x = m[k];
      const uint32_t rbase = (uint32_t)((w0 + k) * 64);
      while (x) {
        const uint32_t r = rbase + (uint32_t)(__ffsll((long long)x) - 1);
        x &= x - 1;
        if (pos < P.cap) {
          if (!ASSEMBLE) {
            P.ids[pos] = r;
          } else {
            uint8_t* t = P.tuples + (uint64_t)pos * P.tuple_size;
            for (int q = 0; q < P.tuple_size; q += 8) *reinterpret_cast<uint64_t*>(t + q) = 0ull;  // null indicators + padding
            for (int c = 0; c < P.ncols; ++c) {
              const SlotDev& S = P.cols[c];
              if (S.def && !((ldg_gather64(S.def + (r >> 6)) >> (63 - (r & 63))) & 1ull)) {  // NULL: tuple->SetNull
                t[S.null_off] = 1;
                continue;
              }
              if (S.plain) {  // ParquetPlainEncoder::Decode: the value as it lies in the page
                uint8_t* slot = t + S.tuple_off;
                switch (S.slot_size) {
                  case 1: *slot = S.plain[r]; break;
                  case 2: *reinterpret_cast<uint16_t*>(slot) = reinterpret_cast<const uint16_t*>(S.plain)[r]; break;
                  case 4: *reinterpret_cast<uint32_t*>(slot) = reinterpret_cast<const uint32_t*>(S.plain)[r]; break;
                  default: *reinterpret_cast<uint64_t*>(slot) = reinterpret_cast<const uint64_t*>(S.plain)[r]; break;
                }
                continue;
              }
              const uint32_t code = code_of_row(S.planes, S.bw, r);
              if (code >= (uint32_t)S.ndict) {  // DictDecoder::GetValue() == false (dict-encoding.h:325)
                atomicAdd(P.bad_codes, 1ull);
                continue;
              }
              uint8_t* slot = t + S.tuple_off;
              switch (S.slot_size) {
                case 1: *slot = static_cast<const uint8_t*>(S.dict)[code]; break;
                case 2: *reinterpret_cast<uint16_t*>(slot) = static_cast<const uint16_t*>(S.dict)[code]; break;
                case 4: *reinterpret_cast<uint32_t*>(slot) = static_cast<const uint32_t*>(S.dict)[code]; break;
                default: *reinterpret_cast<uint64_t*>(slot) = static_cast<const uint64_t*>(S.dict)[code]; break;
              }
            }
          }
        }
        ++pos;
      }
    }
    __syncthreads();
  }
  // the last CTA to leave resets the two counters for the next launch on this stream
  if (threadIdx.x == 0) {
    __threadfence();
    last_s = atomicAdd(P.ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last_s && threadIdx.x == 0) {
    P.ticket[0] = 0;
    P.ticket[1] = 0;
  }
}

static int launch_tile_scan(fle_ctx* ctx, TileScanParams& P, bool assemble) {
  const int rc0 = ensure_scratch(ctx, 1, 0);  // (clears the ticket / tile counters after a reported CUDA error)
  if (rc0 != FLE_OK) return rc0;
  const int64_t nwords = (P.nbits + 63) / 64;
  P.nwords = nwords;
  int wpt = 4;
  if ((nwords + kTileThreads * 4 - 1) / (kTileThreads * 4) < 2 * (int64_t)ctx->num_sms) wpt = 1;
  const int kTileWords = kTileThreads * wpt;
  P.ntiles = (int)((nwords + kTileWords - 1) / kTileWords);
  if ((size_t)P.ntiles * 8 > ctx->tile_state_cap) {
    const size_t cap = std::max<size_t>((size_t)P.ntiles * 8, 64 << 10);
    if (ctx->d_tile_state) {
      FLE_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
      cudaFree(ctx->d_tile_state);
      ctx->d_tile_state = nullptr;
      ctx->tile_state_cap = 0;
    }
    FLE_CUDA_TRY(ctx, cudaMalloc(&ctx->d_tile_state, cap));
    FLE_CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_tile_state, 0, cap, ctx->stream));
    ctx->tile_state_cap = cap;
  }
  ctx->tile_gen = (ctx->tile_gen + 1) & 0xffffffu;
  if (ctx->tile_gen == 0) {  // generation wrap: entries of 2^24 launches ago must not be taken for this launch's
    FLE_CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_tile_state, 0, ctx->tile_state_cap, ctx->stream));
    ctx->tile_gen = 1;
  }
  P.gen = ctx->tile_gen;
  P.tile_state = ctx->d_tile_state;
  P.ticket = ctx->d_ticket;
  const int grid = std::max(1, std::min(P.ntiles, ctx->num_sms * 8));
  if (assemble) {
    if (wpt == 4) tile_scan_kernel<true, 4><<<grid, kTileThreads, 0, ctx->stream>>>(P);
    else tile_scan_kernel<true, 1><<<grid, kTileThreads, 0, ctx->stream>>>(P);
  } else {
    if (wpt == 4) tile_scan_kernel<false, 4><<<grid, kTileThreads, 0, ctx->stream>>>(P);
    else tile_scan_kernel<false, 1><<<grid, kTileThreads, 0, ctx->stream>>>(P);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "tile_scan launch");
  ctx->launches += 1;
  return FLE_OK;
}

int launch_select(fle_ctx* ctx, const uint64_t* bitmap_dev, int64_t nbits, uint32_t* ids_dev, int64_t cap,
                  uint64_t* count_dev) {
  if (nbits < 0 || cap < 0 || !count_dev || (nbits > 0 && !bitmap_dev) || (cap > 0 && !ids_dev))
    return set_error(ctx, FLE_E_INVALID, "bad select arguments");
  if (nbits > (int64_t)0xFFFFFFFFll) return set_error(ctx, FLE_E_UNSUPPORTED, "row ids are 32-bit: at most 2^32-1 rows per call");
  if (nbits == 0) {
    FLE_CUDA_TRY(ctx, cudaMemsetAsync(count_dev, 0, 8, ctx->stream));
    return FLE_OK;
  }
  TileScanParams P;
  std::memset(&P, 0, sizeof(P));
  P.bm = bitmap_dev;
  P.nbits = nbits;
  P.count_out = count_dev;
  P.ids = ids_dev;
  P.cap = cap;
  return launch_tile_scan(ctx, P, false);
}

int launch_assemble(fle_ctx* ctx, const uint64_t* bitmap_dev, int64_t nrows, const fle_slot_desc* cols, int ncols,
                    int tuple_size, uint8_t* tuples_dev, int64_t cap_rows, uint64_t* count_dev, uint64_t* bad_codes_dev) {
  if (nrows < 0 || cap_rows < 0 || !count_dev || !bad_codes_dev || ncols < 1 || ncols > FLE_MAX_COLS || !cols ||
      (nrows > 0 && !bitmap_dev) || (cap_rows > 0 && !tuples_dev))
    return set_error(ctx, FLE_E_INVALID, "bad assemble arguments");
  if (nrows > (int64_t)0xFFFFFFFFll) return set_error(ctx, FLE_E_UNSUPPORTED, "at most 2^32-1 rows per call");
  if (tuple_size < 8 || (tuple_size & 7) || (reinterpret_cast<uintptr_t>(tuples_dev) & 7))
    return set_error(ctx, FLE_E_ALIGN, "tuples are multiples of 8 bytes at an 8-byte aligned address");
  TileScanParams P;
  std::memset(&P, 0, sizeof(P));
  for (int c = 0; c < ncols; ++c) {
    const fle_slot_desc& s = cols[c];
    const bool plain = s.plain_values_dev != nullptr;
    if (!plain && (s.bit_width < 1 || s.bit_width > 32)) return set_error(ctx, FLE_E_INVALID, "bit width outside 1..32");
    if (s.slot_size != 1 && s.slot_size != 2 && s.slot_size != 4 && s.slot_size != 8)
      return set_error(ctx, FLE_E_INVALID, "slot sizes are 1, 2, 4 or 8 bytes");
    if (s.tuple_offset < 0 || s.tuple_offset + s.slot_size > tuple_size || (s.tuple_offset % s.slot_size))
      return set_error(ctx, FLE_E_INVALID, "slot outside the tuple or misaligned");
    if (plain) {
      if (reinterpret_cast<uintptr_t>(s.plain_values_dev) & 7) return set_error(ctx, FLE_E_ALIGN, "value arrays must be 8-byte aligned");
      if ((uint64_t)nrows * (uint64_t)s.slot_size > s.bytes) return set_error(ctx, FLE_E_BUFFER, "rows beyond the value array");
      P.cols[c].plain = static_cast<const uint8_t*>(s.plain_values_dev);
    } else {
      if (!s.planes_dev || !s.dict_dev || s.ndict < 0) return set_error(ctx, FLE_E_INVALID, "null column or dictionary");
      if ((uint64_t)((nrows + 63) / 64) * s.bit_width * 8u > s.bytes) return set_error(ctx, FLE_E_BUFFER, "rows beyond the encoded buffer");
      if (reinterpret_cast<uintptr_t>(s.planes_dev) & 7) return set_error(ctx, FLE_E_ALIGN, "planes must be 8-byte aligned");
    }
    P.cols[c].planes = static_cast<const uint64_t*>(s.planes_dev);
    P.cols[c].dict = s.dict_dev;
    P.cols[c].bw = s.bit_width;
    P.cols[c].ndict = s.ndict;
    P.cols[c].slot_size = s.slot_size;
    P.cols[c].tuple_off = s.tuple_offset;
    if (s.def_planes_dev) {
      if (reinterpret_cast<uintptr_t>(s.def_planes_dev) & 7) return set_error(ctx, FLE_E_ALIGN, "def_planes must be 8-byte aligned");
      if (s.null_offset < 0 || s.null_offset >= tuple_size) return set_error(ctx, FLE_E_INVALID, "NULL indicator outside the tuple");
      P.cols[c].def = static_cast<const uint64_t*>(s.def_planes_dev);
      P.cols[c].null_off = s.null_offset;
    }
  }
  FLE_CUDA_TRY(ctx, cudaMemsetAsync(bad_codes_dev, 0, 8, ctx->stream));
  if (nrows == 0) {
    FLE_CUDA_TRY(ctx, cudaMemsetAsync(count_dev, 0, 8, ctx->stream));
    return FLE_OK;
  }
  P.bm = bitmap_dev;
  P.nbits = nrows;
  P.count_out = count_dev;
  P.cap = cap_rows;
  P.tuples = tuples_dev;
  P.tuple_size = tuple_size;
  P.ncols = ncols;
  P.bad_codes = reinterpret_cast<unsigned long long*>(bad_codes_dev);
  return launch_tile_scan(ctx, P, true);
}

}  // namespace fle
