// FLE pack / unpack kernels, one specialisation per bit width 1..32.
//
// Replaces
//   FleEncoder::Put / Pack_1..16 / Pack_16_32 / Flush   /root/reference/fle-encoding.h:8315-9812
//   FleDecoder::Get / Unpack_1..32                       /root/reference/fle-encoding.h:404-7329
//
// In the reference's layout a group is a 64 x bw bit matrix stored plane-major, so decode
// and encode are bit-matrix transposes.  One thread owns one group and transposes it in
// registers with a fully unrolled 32x32 butterfly (fle_device.cuh); for bw <= 16 the two
// 32-row halves of a group share ONE butterfly (planes of the upper half in the high 16
// lanes), and planes >= bw are compile-time zeros that fold out of the network.
// Values move between registers and global memory through a swizzled shared-memory tile so
// that every global access is a fully coalesced 128-bit transaction.
#include <mutex>
#include <set>
#include <utility>

#include "fle_device.cuh"
#include "fle_internal.h"

namespace fle {

constexpr int kCodecThreads = 128;               // groups per tile
constexpr int kTileValueBytes = kCodecThreads * 256;  // 64 u32 values per group

// value tile in shared memory: row = group (256 B), 16-byte chunks rotated by row + row/16, so
// that 8 lanes are conflict-free both when they own consecutive rows (direct kernels) and when
// they own rows 16 apart (staged kernels: lane t works on row 16*(t&7) + (t>>3))
__device__ __forceinline__ uint32_t tile_chunk_addr(uint32_t tile, int row, int chunk) {
  return tile + row * 256 + (((chunk + row + (row >> 4)) & 15) << 4);
}

// ------------------------------------------------------------------------------------------
// unpack
// ------------------------------------------------------------------------------------------
template <int BW>
__device__ __forceinline__ void transpose_planes_to_values(const uint64_t (&w)[BW], uint32_t (&v)[64]) {
  if constexpr (BW <= 16) {
    uint32_t A[32];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      A[31 - i] = i < BW ? (uint32_t)(w[i < BW ? i : 0] >> 32) : 0u;  // rows 0..31  (bits 63..32)
      A[15 - i] = i < BW ? (uint32_t)(w[i < BW ? i : 0]) : 0u;        // rows 32..63 (bits 31..0)
    }
    transpose32(A);
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      v[k] = A[k] & 0xFFFFu;
      v[32 + k] = A[k] >> 16;
    }
  } else {
    uint32_t A[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) A[31 - i] = i < BW ? (uint32_t)(w[i < BW ? i : 0] >> 32) : 0u;
    transpose32(A);
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = A[k];
#pragma unroll
    for (int i = 0; i < 32; ++i) A[31 - i] = i < BW ? (uint32_t)(w[i < BW ? i : 0]) : 0u;
    transpose32(A);
#pragma unroll
    for (int k = 0; k < 32; ++k) v[32 + k] = A[k];
  }
}

template <int BW>
__global__ void __launch_bounds__(kCodecThreads)
unpack_kernel(const uint64_t* __restrict__ enc, int64_t ngroups, int64_t nvals, uint32_t* __restrict__ out) {
  __shared__ __align__(128) uint8_t tile_mem[kTileValueBytes];
  const uint32_t tile = smem_u32(tile_mem);
  const int t = threadIdx.x;
  const int64_t ntiles = (ngroups + kCodecThreads - 1) / kCodecThreads;
  for (int64_t tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
    const int64_t g = tl * kCodecThreads + t;
    if (g < ngroups) {
      uint64_t w[BW];
      const uint64_t* src = enc + g * BW;
#pragma unroll
      for (int i = 0; i < BW; ++i) w[i] = __ldg(src + i);
      uint32_t v[64];
      transpose_planes_to_values<BW>(w, v);
#pragma unroll
      for (int c = 0; c < 16; ++c)
        sts128u(tile_chunk_addr(tile, t, c), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    }
    __syncthreads();
    // coalesced copy-out: the tile is kTileValueBytes contiguous bytes of `out`
    const int64_t v0 = tl * kCodecThreads * 64;  // first value index of the tile
    uint32_t* dst = out + v0;
#pragma unroll 4
    for (int it = 0; it < 16; ++it) {
      const int L = it * kCodecThreads + t;  // 16-byte chunk index inside the tile
      const int64_t vi = v0 + (int64_t)L * 4;
      if (vi >= nvals) continue;
      uint32_t a, b, c, d;
      lds128u(tile_chunk_addr(tile, L >> 4, L & 15), a, b, c, d);
      if (vi + 4 <= nvals) {
        stg_stream128(dst + L * 4, make_uint4(a, b, c, d));
      } else {
        dst[L * 4] = a;
        if (vi + 1 < nvals) dst[L * 4 + 1] = b;
        if (vi + 2 < nvals) dst[L * 4 + 2] = c;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// pack
// ------------------------------------------------------------------------------------------
template <int BW>
__device__ __forceinline__ void transpose_values_to_planes(const uint32_t (&v)[64], uint64_t (&w)[BW]) {
  constexpr uint32_t kMask = BW >= 32 ? 0xFFFFFFFFu : ((1u << BW) - 1u);
  if constexpr (BW <= 16) {
    uint32_t A[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) A[k] = (v[k] & kMask) | ((v[32 + k] & kMask) << 16);
    transpose32(A);
#pragma unroll
    for (int i = 0; i < BW; ++i) w[i] = ((uint64_t)A[31 - i] << 32) | A[15 - i];
  } else {
    uint32_t A[32], B[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      A[k] = v[k] & kMask;
      B[k] = v[32 + k] & kMask;
    }
    transpose32(A);
    transpose32(B);
#pragma unroll
    for (int i = 0; i < BW; ++i) w[i] = ((uint64_t)A[31 - i] << 32) | B[31 - i];
  }
}

template <int BW>
__global__ void __launch_bounds__(kCodecThreads)
pack_kernel(const uint32_t* __restrict__ vals, int64_t nvals, int64_t ngroups, uint64_t* __restrict__ enc) {
  __shared__ __align__(128) uint8_t tile_mem[kTileValueBytes];
  const uint32_t tile = smem_u32(tile_mem);
  const int t = threadIdx.x;
  const int64_t ntiles = (ngroups + kCodecThreads - 1) / kCodecThreads;
  for (int64_t tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
    const int64_t v0 = tl * kCodecThreads * 64;
    const uint32_t* src = vals + v0;
    // coalesced copy-in; rows past the end are zero (the reference leaves stale data there)
#pragma unroll 4
    for (int it = 0; it < 16; ++it) {
      const int L = it * kCodecThreads + t;
      const int64_t vi = v0 + (int64_t)L * 4;
      uint4 x = make_uint4(0, 0, 0, 0);
      if (vi + 4 <= nvals) {
        x = ldg_stream128(src + L * 4);
      } else if (vi < nvals) {
        x.x = src[L * 4];
        if (vi + 1 < nvals) x.y = src[L * 4 + 1];
        if (vi + 2 < nvals) x.z = src[L * 4 + 2];
      }
      sts128u(tile_chunk_addr(tile, L >> 4, L & 15), x.x, x.y, x.z, x.w);
    }
    __syncthreads();
    const int64_t g = tl * kCodecThreads + t;
    if (g < ngroups) {
      uint32_t v[64];
#pragma unroll
      for (int c = 0; c < 16; ++c)
        lds128u(tile_chunk_addr(tile, t, c), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      uint64_t w[BW];
      transpose_values_to_planes<BW>(v, w);
      uint64_t* dst = enc + g * BW;
#pragma unroll
      for (int i = 0; i < BW; ++i) dst[i] = w[i];
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// TMA-staged variants (codec_variant = 4; 1 = auto picks per bit width, see unpack_choice).
//
// The direct variants above fetch / store each thread's bw plane words with 64-bit accesses that
// are bw*8 bytes apart between neighbouring lanes: one warp instruction touches 32 different
// 128-byte lines, and for wide columns the L1 wavefront rate, not HBM, bounds the kernel
// (profiles/: bw 16/24/32 at 0.47-0.58 of roofline).  Here the plane side of a 128-group tile
// moves as 8 TMA bulk copies (cp.async.bulk) between global memory and a shared-memory image
// whose 8 segments are skewed by 16 bytes, so that the per-thread 64/128-bit shared accesses are
// bank-conflict free for every bit width (same layout as the eval kernels).
// ------------------------------------------------------------------------------------------
template <int BW>
struct PlaneTile {
  static constexpr int kGroups = kCodecThreads;             // 128 groups per tile
  static constexpr int kSegGroups = kGroups / 8;            // 16
  static constexpr uint32_t kSegBytes = kSegGroups * BW * 8;
  static constexpr uint32_t kSegStride = kSegBytes + 16;
  static constexpr uint32_t kBytes = 8 * kSegStride;        // shared-memory image of one tile
  // thread t works on segment t & 7, group (t >> 3) of that segment
  __device__ static __forceinline__ uint32_t thread_off(int t) {
    return (uint32_t)(t & 7) * kSegStride + (uint32_t)(t >> 3) * (BW * 8);
  }
  __device__ static __forceinline__ int thread_group(int t) { return (t & 7) * kSegGroups + (t >> 3); }
};

template <int BW>
__global__ void __launch_bounds__(kCodecThreads)
unpack_tma_kernel(const uint64_t* __restrict__ enc, int64_t ngroups, int64_t nvals, uint32_t* __restrict__ out) {
  using PT = PlaneTile<BW>;
  constexpr int kStages = 2;
  extern __shared__ __align__(128) uint8_t smem[];
  // [value tile 32 KB][stage 0][stage 1][2 mbarriers]
  const uint32_t tile = smem_u32(smem);
  const uint32_t stage0 = tile + kTileValueBytes;
  const uint32_t bar0 = stage0 + kStages * ((PT::kBytes + 127) / 128 * 128);
  constexpr uint32_t kStageStride = (PT::kBytes + 127) / 128 * 128;
  const int t = threadIdx.x;
  const int64_t ntiles_full = ngroups / PT::kGroups;  // full tiles go through TMA

  if (t == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(bar0 + s * 8, 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](int stage, int64_t tl) {
    const uint32_t bar = bar0 + stage * 8;
    mbar_arrive_expect_tx(bar, 8 * PT::kSegBytes);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(enc) + (uint64_t)tl * PT::kGroups * BW * 8;
#pragma unroll
    for (int s = 0; s < 8; ++s)
      tma_load_1d(stage0 + stage * kStageStride + s * PT::kSegStride, src + (uint64_t)s * PT::kSegBytes, PT::kSegBytes, bar);
  };
  if (t == 0 && (int64_t)blockIdx.x < ntiles_full) issue(0, blockIdx.x);

  const int64_t ntiles = (ngroups + PT::kGroups - 1) / PT::kGroups;
  int stage = 0;
  uint32_t parity = 0;
  for (int64_t tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
    const bool full = tl < ntiles_full;
    const int row = full ? PT::thread_group(t) : t;  // tile row (group inside the tile) of this thread
    const int64_t g = tl * PT::kGroups + row;
    uint64_t w[BW];
    if (full) {
      // prefetch the next tile into the other stage (its previous contents were consumed in
      // the previous iteration, which ended with __syncthreads)
      if (t == 0) {
        const int64_t nxt = tl + gridDim.x;
        if (nxt < ntiles_full) issue(stage ^ 1, nxt);
      }
      mbar_wait(bar0 + stage * 8, parity);
      const uint32_t a = stage0 + stage * kStageStride + PT::thread_off(t);
      if constexpr (BW % 2 == 0) {
#pragma unroll
        for (int i = 0; i < BW; i += 2) lds128(a + i * 8, w[i], w[i + 1 < BW ? i + 1 : i]);
      } else {
#pragma unroll
        for (int i = 0; i < BW; ++i) w[i] = lds64(a + i * 8);
      }
      stage ^= 1;
      if (stage == 0) parity ^= 1u;
    } else if (g < ngroups) {
      const uint64_t* src = enc + g * BW;
#pragma unroll
      for (int i = 0; i < BW; ++i) w[i] = __ldg(src + i);
    }
    if (g < ngroups) {
      uint32_t v[64];
      transpose_planes_to_values<BW>(w, v);
#pragma unroll
      for (int c = 0; c < 16; ++c)
        sts128u(tile_chunk_addr(tile, row, c), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    }
    __syncthreads();
    const int64_t v0 = tl * PT::kGroups * 64;
    uint32_t* dst = out + v0;
#pragma unroll 4
    for (int it = 0; it < 16; ++it) {
      const int L = it * kCodecThreads + t;
      const int64_t vi = v0 + (int64_t)L * 4;
      if (vi >= nvals) continue;
      uint32_t a, b, c, d;
      lds128u(tile_chunk_addr(tile, L >> 4, L & 15), a, b, c, d);
      if (vi + 4 <= nvals) {
        stg_stream128(dst + L * 4, make_uint4(a, b, c, d));
      } else {
        dst[L * 4] = a;
        if (vi + 1 < nvals) dst[L * 4 + 1] = b;
        if (vi + 2 < nvals) dst[L * 4 + 2] = c;
      }
    }
    __syncthreads();
  }
}

// Narrow decode targets.  FleDecoder::Get<T> serves int8 / int16 columns from u8 / u16 staging for
// bw <= 8 / <= 16 (fle-encoding.h:236-238, :8354-8361; dict-test.cc:64-157): writing 4 bytes per value
// where 1-2 do is 2.5-3.5x the necessary HBM write traffic at bw <= 8.  Same plane-side pipeline as
// unpack_tma_kernel; the transposed words already hold two 16-bit values each (rows k and 32 + k),
// byte permutes put neighbouring rows next to each other, and the value tile is 64 * ES bytes per
// group (16-byte chunks rotated by the row, copied out with coalesced 128-bit stores).
// Algorithmic bytes per group: bw * 8 + 64 * ES.
template <int BW, int ES>
__global__ void __launch_bounds__(kCodecThreads)
unpack_narrow_kernel(const uint64_t* __restrict__ enc, int64_t ngroups, int64_t nvals, uint8_t* __restrict__ out) {
  static_assert((ES == 1 && BW <= 8) || (ES == 2 && BW <= 16), "u8 for bw <= 8, u16 for bw <= 16");
  using PT = PlaneTile<BW>;
  constexpr int kStages = 2;
  constexpr int kRowBytes = 64 * ES, kChunks = kRowBytes / 16;  // 4 or 8 chunks per group
  constexpr int kTileBytes = kCodecThreads * kRowBytes;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t tile = smem_u32(smem);
  const uint32_t stage0 = tile + kTileBytes;
  constexpr uint32_t kStageStride = (PT::kBytes + 127) / 128 * 128;
  const uint32_t bar0 = stage0 + kStages * kStageStride;
  const int t = threadIdx.x;
  const int64_t ntiles_full = ngroups / PT::kGroups;
  if (t == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(bar0 + s * 8, 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](int stage, int64_t tl) {
    const uint32_t bar = bar0 + stage * 8;
    mbar_arrive_expect_tx(bar, 8 * PT::kSegBytes);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(enc) + (uint64_t)tl * PT::kGroups * BW * 8;
#pragma unroll
    for (int s = 0; s < 8; ++s)
      tma_load_1d(stage0 + stage * kStageStride + s * PT::kSegStride, src + (uint64_t)s * PT::kSegBytes, PT::kSegBytes, bar);
  };
  if (t == 0 && (int64_t)blockIdx.x < ntiles_full) issue(0, blockIdx.x);
  auto chunk_addr = [&](int row, int chunk) { return tile + row * kRowBytes + (((chunk + row + (row >> 4)) & (kChunks - 1)) << 4); };

  const int64_t ntiles = (ngroups + PT::kGroups - 1) / PT::kGroups;
  int stage = 0;
  uint32_t parity = 0;
  for (int64_t tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
    const bool full = tl < ntiles_full;
    const int row = full ? PT::thread_group(t) : t;
    const int64_t g = tl * PT::kGroups + row;
    uint64_t w[BW];
    if (full) {
      if (t == 0) {
        const int64_t nxt = tl + gridDim.x;
        if (nxt < ntiles_full) issue(stage ^ 1, nxt);
      }
      mbar_wait(bar0 + stage * 8, parity);
      const uint32_t a = stage0 + stage * kStageStride + PT::thread_off(t);
      if constexpr (BW % 2 == 0) {
#pragma unroll
        for (int i = 0; i < BW; i += 2) lds128(a + i * 8, w[i], w[i + 1 < BW ? i + 1 : i]);
      } else {
#pragma unroll
        for (int i = 0; i < BW; ++i) w[i] = lds64(a + i * 8);
      }
      stage ^= 1;
      if (stage == 0) parity ^= 1u;
    } else if (g < ngroups) {
      const uint64_t* src = enc + g * BW;
#pragma unroll
      for (int i = 0; i < BW; ++i) w[i] = __ldg(src + i);
    }
    if (g < ngroups) {
      uint32_t A[32];  // after the transpose A[k] = value(row k) | value(row 32 + k) << 16
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        A[31 - i] = i < BW ? (uint32_t)(w[i < BW ? i : 0] >> 32) : 0u;
        A[15 - i] = i < BW ? (uint32_t)(w[i < BW ? i : 0]) : 0u;
      }
      transpose32(A);
      if constexpr (ES == 2) {
        // chunk c (8 values): rows 8c..8c+7; rows 0..31 come from the low halves, rows 32..63 from the high halves
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = (c & 3) * 8 + 2 * j;
            o[j] = __byte_perm(A[k], A[k + 1], c < 4 ? 0x5410 : 0x7632);
          }
          sts128u(chunk_addr(row, c), o[0], o[1], o[2], o[3]);
        }
      } else {
        // chunk c (16 values): rows 16c..16c+15; a value is byte 0 (rows 0..31) or byte 2 (rows 32..63) of its word
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = (c & 1) * 16 + 4 * j;
            const uint32_t sel = c < 2 ? 0x0040 : 0x0062;
            const uint32_t lo = __byte_perm(A[k], A[k + 1], sel), hi = __byte_perm(A[k + 2], A[k + 3], sel);
            o[j] = __byte_perm(lo, hi, 0x5410);
          }
          sts128u(chunk_addr(row, c), o[0], o[1], o[2], o[3]);
        }
      }
    }
    __syncthreads();
    const int64_t b0 = tl * (int64_t)kTileBytes;  // first output byte of the tile
    const int64_t nbytes = nvals * ES;
#pragma unroll
    for (int it = 0; it < kChunks; ++it) {
      const int L = it * kCodecThreads + t;  // 16-byte chunk index inside the tile
      const int64_t bi = b0 + (int64_t)L * 16;
      if (bi >= nbytes) continue;
      uint32_t a, b, c, d;
      lds128u(chunk_addr(L / kChunks, L % kChunks), a, b, c, d);
      if (bi + 16 <= nbytes) {
        stg_stream128(out + bi, make_uint4(a, b, c, d));
      } else {
        const uint32_t wds[4] = {a, b, c, d};
        for (int q = 0; q < (int)(nbytes - bi); ++q) out[bi + q] = (uint8_t)(wds[q >> 2] >> ((q & 3) * 8));
      }
    }
    __syncthreads();
  }
}

// Narrow decode targets, second form (default): the pipeline of unpack_warp_kernel below -- three plane stages handed
// back through "empty" mbarriers, warp-private copy-out, NO CTA barrier in the tile loop -- because a narrow tile is
// short (bw 8 -> u8: 8 KB in, 8 KB out) and the two CTA barriers per tile of the first form left it at 0.40-0.59 of
// its roofline (2^26 values).  u8 rows come out of two 8 x 8 SWAR transposes (fle_device.cuh: 96 instructions per
// group instead of the 32 x 32 butterfly's 240) and 4 x 4 byte transposes (16 PRMT) that put neighbouring rows next
// to each other; u16 rows out of the shared 32 x 32 butterfly as before.
// Value tile: row = group (64 * ES bytes); 16-byte chunk c of row r sits at slot (c + r / 16) mod (4 ES) of the
// row, and for u8 the two rows of a 128-byte line swap places in the upper half of the tile -- the eight lanes of
// a store phase own rows 16 apart and the eight lanes of a copy-out phase read two neighbouring rows: both hit
// eight distinct 16-byte slots.
template <int ES>
__device__ __forceinline__ uint32_t narrow_chunk_addr(uint32_t tile, int row, int chunk) {
  constexpr int kChunks = 4 * ES;
  const int prow = ES == 1 ? (row ^ ((row >> 6) & 1)) : row;
  return tile + (uint32_t)prow * (64u * ES) + ((uint32_t)((chunk + (row >> 4)) & (kChunks - 1)) << 4);
}

// the 64 values of a group as 16 * ES words in row order
template <int BW, int ES>
__device__ __forceinline__ void transpose_planes_to_narrow(const uint64_t (&w)[BW], uint32_t (&o)[16 * ES]) {
  if constexpr (ES == 1) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // rows 0..31 (bits 63..32 of the plane words), rows 32..63
      uint32_t S[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) S[7 - i] = i < BW ? (uint32_t)(w[i < BW ? i : 0] >> (h == 0 ? 32 : 0)) : 0u;
      swar8_transpose(S);  // S[x] byte b = value of the half's row 24 - 8 b + x
#pragma unroll
      for (int x0 = 0; x0 < 8; x0 += 4) {  // 4 x 4 byte transpose: Y[b] = rows 24 - 8 b + x0 .. + 3
        const uint32_t t0 = __byte_perm(S[x0], S[x0 + 1], 0x5140), t1 = __byte_perm(S[x0], S[x0 + 1], 0x7362);
        const uint32_t t2 = __byte_perm(S[x0 + 2], S[x0 + 3], 0x5140), t3 = __byte_perm(S[x0 + 2], S[x0 + 3], 0x7362);
        // output word m of the half (rows 4 m .. 4 m + 3): b = 3 - m / 2, x0 = 4 (m % 2)
        o[h * 8 + 6 + x0 / 4] = __byte_perm(t0, t2, 0x5410);  // b = 0: rows 24 + x0 ..
        o[h * 8 + 4 + x0 / 4] = __byte_perm(t0, t2, 0x7632);  // b = 1: rows 16 + x0 ..
        o[h * 8 + 2 + x0 / 4] = __byte_perm(t1, t3, 0x5410);  // b = 2: rows 8 + x0 ..
        o[h * 8 + 0 + x0 / 4] = __byte_perm(t1, t3, 0x7632);  // b = 3: rows x0 ..
      }
    }
  } else {
    uint32_t A[32];  // after the transpose A[k] = value(row k) | value(row 32 + k) << 16
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      A[31 - i] = i < BW ? (uint32_t)(w[i < BW ? i : 0] >> 32) : 0u;
      A[15 - i] = i < BW ? (uint32_t)(w[i < BW ? i : 0]) : 0u;
    }
    transpose32(A);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      o[j] = __byte_perm(A[2 * j], A[2 * j + 1], 0x5410);       // rows 2 j, 2 j + 1
      o[16 + j] = __byte_perm(A[2 * j], A[2 * j + 1], 0x7632);  // rows 32 + 2 j, 33 + 2 j
    }
  }
}

template <int BW, int ES>
__global__ void __launch_bounds__(kCodecThreads)
unpack_narrow_warp_kernel(const uint64_t* __restrict__ enc, int64_t ngroups, int64_t nvals, uint8_t* __restrict__ out) {
  static_assert((ES == 1 && BW <= 8) || (ES == 2 && BW <= 16), "u8 for bw <= 8, u16 for bw <= 16");
  using PT = PlaneTile<BW>;
  constexpr int kStages = 3;
  constexpr int kRowBytes = 64 * ES, kChunks = 4 * ES;
  constexpr int kTileBytes = kCodecThreads * kRowBytes;
  constexpr uint32_t kStageStride = (PT::kBytes + 127) / 128 * 128;
  extern __shared__ __align__(128) uint8_t smem[];
  // [value tile][stages][full mbarriers][empty mbarriers]
  const uint32_t tile = smem_u32(smem);
  const uint32_t stage0 = tile + kTileBytes;
  const uint32_t full0 = stage0 + kStages * kStageStride;
  const uint32_t empty0 = full0 + kStages * 8;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int64_t ntiles_full = nvals / (PT::kGroups * 64);

  if (t == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full0 + s * 8, 1);
      mbar_init(empty0 + s * 8, kCodecThreads / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](int stage, int64_t tl) {
    const uint32_t bar = full0 + stage * 8;
    mbar_arrive_expect_tx(bar, 8 * PT::kSegBytes);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(enc) + (uint64_t)tl * PT::kGroups * BW * 8;
#pragma unroll
    for (int s = 0; s < 8; ++s)
      tma_load_1d(stage0 + stage * kStageStride + s * PT::kSegStride, src + (uint64_t)s * PT::kSegBytes, PT::kSegBytes, bar);
  };
  if (t == 0) {
    for (int s = 0; s < kStages - 1; ++s) {
      const int64_t tl = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
      if (tl < ntiles_full) issue(s, tl);
    }
  }
  const int row = PT::thread_group(t);  // (t & 7) * 16 + (t >> 3)
  int stage = 0;
  uint32_t use = 0;
  for (int64_t tl = blockIdx.x; tl < ntiles_full; tl += gridDim.x) {
    if (t == 0) {  // refill the stage consumed in the previous iteration, two tiles ahead (protocol: unpack_warp_kernel)
      const int64_t nxt = tl + (int64_t)(kStages - 1) * gridDim.x;
      int st2 = stage + kStages - 1;
      uint32_t use2 = use;
      if (st2 >= kStages) { st2 -= kStages; use2 += 1; }
      if (nxt < ntiles_full) {
        if (use2 >= 1) mbar_wait(empty0 + st2 * 8, (use2 - 1) & 1u);
        issue(st2, nxt);
      }
    }
    mbar_wait(full0 + stage * 8, use & 1u);
    uint64_t w[BW];
    const uint32_t a = stage0 + stage * kStageStride + PT::thread_off(t);
    if constexpr (BW % 2 == 0) {
#pragma unroll
      for (int i = 0; i < BW; i += 2) lds128(a + i * 8, w[i], w[i + 1 < BW ? i + 1 : i]);
    } else {
#pragma unroll
      for (int i = 0; i < BW; ++i) w[i] = lds64(a + i * 8);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + stage * 8);
    uint32_t o[16 * ES];
    transpose_planes_to_narrow<BW, ES>(w, o);
#pragma unroll
    for (int c = 0; c < kChunks; ++c)
      sts128u(narrow_chunk_addr<ES>(tile, row, c), o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
    __syncwarp();
    uint8_t* dst = out + tl * (int64_t)kTileBytes;
#pragma unroll
    for (int it = 0; it < kChunks; ++it) {
      const int q = it * 32 + lane;   // 16-byte chunk among this warp's 32 rows
      const int rr = q / kChunks;     // this warp's rr-th row: segment rr >> 2, row (rr & 3) of the warp's four
      const int ch = q % kChunks;
      const int r = (rr >> 2) * 16 + warp * 4 + (rr & 3);
      uint32_t x, y, z, u;
      lds128u(narrow_chunk_addr<ES>(tile, r, ch), x, y, z, u);
      stg_stream128(dst + r * kRowBytes + ch * 16, make_uint4(x, y, z, u));
    }
    __syncwarp();
    if (++stage == kStages) {
      stage = 0;
      use += 1;
    }
  }
  // ragged remainder (< 128 * 64 values): plain loads and guarded stores
  for (int64_t g = ntiles_full * PT::kGroups + (int64_t)blockIdx.x * kCodecThreads + t; g < ngroups;
       g += (int64_t)gridDim.x * kCodecThreads) {
    uint64_t w[BW];
    const uint64_t* src = enc + g * BW;
#pragma unroll
    for (int i = 0; i < BW; ++i) w[i] = __ldg(src + i);
    uint32_t o[16 * ES];
    transpose_planes_to_narrow<BW, ES>(w, o);
    const int64_t left = nvals - g * 64;
    uint8_t* dst = out + g * kRowBytes;
    if (left >= 64) {
#pragma unroll
      for (int c = 0; c < kChunks; ++c) stg_stream128(dst + 16 * c, make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]));
    } else {
#pragma unroll
      for (int k = 0; k < 64; ++k) {
        if (k < left) {
          if constexpr (ES == 1) dst[k] = (uint8_t)(o[k >> 2] >> ((k & 3) * 8));
          else reinterpret_cast<uint16_t*>(dst)[k] = (uint16_t)(o[k >> 1] >> ((k & 1) * 16));
        }
      }
    }
  }
}

// Unpack with warp-private copy-out (codec_variant = 3): same data path as unpack_tma_kernel, but a
// warp copies out the 32 rows its own lanes wrote (two adjacent rows = 512 contiguous bytes per
// store instruction), so the two CTA barriers around the copy-out become __syncwarp, and the plane
// stages are handed back to the producer through per-stage "empty" mbarriers (one arrival per warp)
// instead of a CTA barrier: the four warps of a CTA drift apart and one warp's stores overlap
// another's transposes.
template <int BW>
__global__ void __launch_bounds__(kCodecThreads)
unpack_warp_kernel(const uint64_t* __restrict__ enc, int64_t ngroups, int64_t nvals, uint32_t* __restrict__ out) {
  using PT = PlaneTile<BW>;
  constexpr int kStages = 3;
  constexpr uint32_t kStageStride = (PT::kBytes + 127) / 128 * 128;
  extern __shared__ __align__(128) uint8_t smem[];
  // [value tile 32 KB][stages][full mbarriers][empty mbarriers]
  const uint32_t tile = smem_u32(smem);
  const uint32_t stage0 = tile + kTileValueBytes;
  const uint32_t full0 = stage0 + kStages * kStageStride;
  const uint32_t empty0 = full0 + kStages * 8;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int64_t ntiles_full = nvals / (PT::kGroups * 64);

  if (t == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full0 + s * 8, 1);
      mbar_init(empty0 + s * 8, kCodecThreads / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](int stage, int64_t tl) {
    const uint32_t bar = full0 + stage * 8;
    mbar_arrive_expect_tx(bar, 8 * PT::kSegBytes);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(enc) + (uint64_t)tl * PT::kGroups * BW * 8;
#pragma unroll
    for (int s = 0; s < 8; ++s)
      tma_load_1d(stage0 + stage * kStageStride + s * PT::kSegStride, src + (uint64_t)s * PT::kSegBytes, PT::kSegBytes, bar);
  };
  if (t == 0) {
    for (int s = 0; s < kStages - 1; ++s) {
      const int64_t tl = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
      if (tl < ntiles_full) issue(s, tl);
    }
  }
  const int row = PT::thread_group(t);  // (t & 7) * 16 + (t >> 3)
  int stage = 0;
  uint32_t use = 0;
  for (int64_t tl = blockIdx.x; tl < ntiles_full; tl += gridDim.x) {
    if (t == 0) {
      // refill the stage consumed in the previous iteration, two tiles ahead, once every warp has
      // read it: iteration i of this CTA uses stage i % 3 for the (i / 3)-th time (== use); the tile
      // of iteration i + 2 is fill number use2 of stage st2 and must wait for consumption use2 - 1
      const int64_t nxt = tl + (int64_t)(kStages - 1) * gridDim.x;
      int st2 = stage + kStages - 1;
      uint32_t use2 = use;
      if (st2 >= kStages) { st2 -= kStages; use2 += 1; }
      if (nxt < ntiles_full) {
        if (use2 >= 1) mbar_wait(empty0 + st2 * 8, (use2 - 1) & 1u);
        issue(st2, nxt);
      }
    }
    mbar_wait(full0 + stage * 8, use & 1u);
    uint64_t w[BW];
    const uint32_t a = stage0 + stage * kStageStride + PT::thread_off(t);
    if constexpr (BW % 2 == 0) {
#pragma unroll
      for (int i = 0; i < BW; i += 2) lds128(a + i * 8, w[i], w[i + 1 < BW ? i + 1 : i]);
    } else {
#pragma unroll
      for (int i = 0; i < BW; ++i) w[i] = lds64(a + i * 8);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + stage * 8);
    uint32_t v[64];
    transpose_planes_to_values<BW>(w, v);
#pragma unroll
    for (int c = 0; c < 16; ++c)
      sts128u(tile_chunk_addr(tile, row, c), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    __syncwarp();
    uint32_t* dst = out + tl * PT::kGroups * 64;
#pragma unroll 4
    for (int it = 0; it < 16; ++it) {
      const int q = it * 32 + lane;     // 16-byte chunk among this warp's 32 rows
      const int rr = q >> 4;            // this warp's rr-th row: segment rr >> 2, row (rr & 3) of the warp's four
      const int r = (rr >> 2) * 16 + warp * 4 + (rr & 3);
      uint32_t x, y, z, u;
      lds128u(tile_chunk_addr(tile, r, q & 15), x, y, z, u);
      stg_stream128(dst + r * 64 + (q & 15) * 4, make_uint4(x, y, z, u));
    }
    __syncwarp();
    if (++stage == kStages) {
      stage = 0;
      use += 1;
    }
  }
  // ragged remainder (< 128 * 64 values): plain loads and guarded stores
  for (int64_t g = ntiles_full * PT::kGroups + (int64_t)blockIdx.x * kCodecThreads + t; g < ngroups;
       g += (int64_t)gridDim.x * kCodecThreads) {
    uint64_t w[BW];
    const uint64_t* src = enc + g * BW;
#pragma unroll
    for (int i = 0; i < BW; ++i) w[i] = __ldg(src + i);
    uint32_t v[64];
    transpose_planes_to_values<BW>(w, v);
    const int64_t left = nvals - g * 64;
    uint32_t* dst = out + g * 64;
    if (left >= 64) {
#pragma unroll
      for (int c = 0; c < 16; ++c) stg_stream128(dst + 4 * c, make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
    } else {
#pragma unroll
      for (int k = 0; k < 64; ++k)
        if (k < left) dst[k] = v[k];
    }
  }
}

// Unpack with bulk stores (codec_variant = 2): the copy-out phase of the kernel above (16 x LDS.128
// + 16 x STG.128 per thread between two CTA barriers) is replaced by ONE 256-byte bulk store per
// thread.  A thread writes the 64 values of its group to its own shared-memory row (row stride
// 272 bytes: the 16-byte STS of neighbouring lanes land in different banks), fences the async
// proxy and hands the row to the copy engine; the next iteration only waits until the engine has
// READ the row (cp.async.bulk.wait_group.read) before overwriting it.  One CTA barrier per tile
// remains, after the plane words are in registers, so that the stage can be refilled two tiles ahead.
constexpr uint32_t kRowStride = 256 + 16;
__host__ __device__ constexpr int unpack_bulk_stages(int bw) { return bw <= 16 ? 3 : 2; }  // 3 CTAs per SM up to bw 16, 2 beyond

template <int BW>
__global__ void __launch_bounds__(kCodecThreads)
unpack_bulk_kernel(const uint64_t* __restrict__ enc, int64_t ngroups, int64_t nvals, uint32_t* __restrict__ out) {
  using PT = PlaneTile<BW>;
  constexpr int kStages = unpack_bulk_stages(BW);
  constexpr uint32_t kStageStride = (PT::kBytes + 127) / 128 * 128;
  extern __shared__ __align__(128) uint8_t smem[];
  // [value rows 128 x 272 B][stages][mbarriers]
  const uint32_t rows = smem_u32(smem);
  const uint32_t stage0 = rows + kCodecThreads * kRowStride;
  const uint32_t bar0 = stage0 + kStages * kStageStride;
  const int t = threadIdx.x;
  const uint32_t my_row = rows + t * kRowStride;
  // tiles whose 128 groups are complete rows of 64 values go through TMA both ways
  const int64_t ntiles_full = nvals / (PT::kGroups * 64);

  if (t == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(bar0 + s * 8, 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](int stage, int64_t tl) {
    const uint32_t bar = bar0 + stage * 8;
    mbar_arrive_expect_tx(bar, 8 * PT::kSegBytes);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(enc) + (uint64_t)tl * PT::kGroups * BW * 8;
#pragma unroll
    for (int s = 0; s < 8; ++s)
      tma_load_1d(stage0 + stage * kStageStride + s * PT::kSegStride, src + (uint64_t)s * PT::kSegBytes, PT::kSegBytes, bar);
  };
  if (t == 0) {
    for (int s = 0; s < kStages - 1; ++s) {
      const int64_t tl = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
      if (tl < ntiles_full) issue(s, tl);
    }
  }
  int stage = 0;
  uint32_t parity = 0;
  const int row = PT::thread_group(t);
  for (int64_t tl = blockIdx.x; tl < ntiles_full; tl += gridDim.x) {
    uint64_t w[BW];
    mbar_wait(bar0 + stage * 8, parity);
    const uint32_t a = stage0 + stage * kStageStride + PT::thread_off(t);
    if constexpr (BW % 2 == 0) {
#pragma unroll
      for (int i = 0; i < BW; i += 2) lds128(a + i * 8, w[i], w[i + 1 < BW ? i + 1 : i]);
    } else {
#pragma unroll
      for (int i = 0; i < BW; ++i) w[i] = lds64(a + i * 8);
    }
    __syncthreads();  // every thread holds its planes: this stage and the one before it are free
    if (t == 0) {
      const int64_t nxt = tl + (int64_t)(kStages - 1) * gridDim.x;
      int st2 = stage + kStages - 1;
      if (st2 >= kStages) st2 -= kStages;
      if (nxt < ntiles_full) issue(st2, nxt);
    }
    uint32_t v[64];
    transpose_planes_to_values<BW>(w, v);
    tma_store_wait_read<0>();  // the copy engine is done reading this thread's previous row
#pragma unroll
    for (int c = 0; c < 16; ++c) sts128u(my_row + c * 16, v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    fence_proxy_async_smem();
    tma_store_1d(out + (tl * PT::kGroups + row) * 64, my_row, 256);
    tma_store_commit();
    if (++stage == kStages) {
      stage = 0;
      parity ^= 1u;
    }
  }
  // ragged remainder (< 128 * 64 values): plain loads and guarded stores
  for (int64_t g = ntiles_full * PT::kGroups + (int64_t)blockIdx.x * kCodecThreads + t; g < ngroups;
       g += (int64_t)gridDim.x * kCodecThreads) {
    uint64_t w[BW];
    const uint64_t* src = enc + g * BW;
#pragma unroll
    for (int i = 0; i < BW; ++i) w[i] = __ldg(src + i);
    uint32_t v[64];
    transpose_planes_to_values<BW>(w, v);
    const int64_t left = nvals - g * 64;
    uint32_t* dst = out + g * 64;
    if (left >= 64) {
#pragma unroll
      for (int c = 0; c < 16; ++c) stg_stream128(dst + 4 * c, make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
    } else {
#pragma unroll
      for (int k = 0; k < 64; ++k)
        if (k < left) dst[k] = v[k];
    }
  }
  tma_store_wait<0>();  // bulk stores complete before the CTA (and its shared memory) goes away
}

// Pack: the value side of a tile (32 KB, contiguous in global memory) is fetched with
// asynchronous 16-byte copies (cp.async, SASS LDGSTS) straight into the swizzled tile, double
// buffered: the next tile is in flight while this one is transposed, with no staging registers.
// (The first version loaded through registers, 4 x 16 bytes per thread at a time: ncu showed
// 70 % of the stall samples on those loads and 0.60-0.77 of roofline.)  The plane image of a
// tile is written over the value buffer it was transposed from -- every thread holds its row in
// registers by then -- so a CTA needs 2 x 32 KB for any bit width and three CTAs fit an SM.
constexpr uint32_t kPackBufBytes = kTileValueBytes + 128;  // >= PlaneTile<32>::kBytes

template <int BW>
__global__ void __launch_bounds__(kCodecThreads)
pack_tma_kernel(const uint32_t* __restrict__ vals, int64_t nvals, int64_t ngroups, uint64_t* __restrict__ enc) {
  using PT = PlaneTile<BW>;
  static_assert(PT::kBytes <= kPackBufBytes, "plane image must fit the value buffer it replaces");
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t tile0 = smem_u32(smem);  // two buffers: value tile, then the plane image of the same tile
  const int t = threadIdx.x;
  const int64_t ntiles = (ngroups + PT::kGroups - 1) / PT::kGroups;
  const int64_t ntiles_full = ngroups / PT::kGroups;

  // coalesced copy-in; rows past the end are zero (the reference leaves stale data there)
  auto fetch = [&](int buf, int64_t tl) {
    const int64_t v0 = tl * PT::kGroups * 64;
    const uint32_t dst = tile0 + buf * kPackBufBytes;
    if (v0 + PT::kGroups * 64 <= nvals) {  // whole tile inside the input (CTA-uniform)
      const uint32_t* src = vals + v0 + t * 4;
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int L = it * kCodecThreads + t;  // 16-byte chunk index inside the tile
        cp_async16(tile_chunk_addr(dst, L >> 4, L & 15), src + it * kCodecThreads * 4, 16u);
      }
    } else {
#pragma unroll 1
      for (int it = 0; it < 16; ++it) {
        const int L = it * kCodecThreads + t;
        const int64_t left = nvals - (v0 + (int64_t)L * 4);
        const uint32_t nb = left >= 4 ? 16u : (left > 0 ? (uint32_t)left * 4u : 0u);
        cp_async16(tile_chunk_addr(dst, L >> 4, L & 15), nb ? vals + v0 + L * 4 : vals, nb);
      }
    }
    cp_async_commit();
  };

  int buf = 0;
  if ((int64_t)blockIdx.x < ntiles) fetch(0, blockIdx.x);
  for (int64_t tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
    cp_async_wait<0>();  // this tile's values have landed (this thread's copies; the barrier covers the rest)
    // the other buffer holds the previous tile's plane image: its bulk stores must have finished
    // READING it before the next tile's values overwrite it (only thread 0 issued them)
    if (t == 0) tma_store_wait_read<0>();
    __syncthreads();
    const int64_t nxt = tl + gridDim.x;
    if (nxt < ntiles) fetch(buf ^ 1, nxt);
    const uint32_t tile = tile0 + buf * kPackBufBytes;
    const bool full = tl < ntiles_full;
    const int row = full ? PT::thread_group(t) : t;
    const int64_t g = tl * PT::kGroups + row;
    uint32_t v[64];
#pragma unroll
    for (int c = 0; c < 16; ++c)
      lds128u(tile_chunk_addr(tile, row, c), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    __syncthreads();  // every row is in registers: the buffer may now take the plane image
    uint64_t w[BW];
    transpose_values_to_planes<BW>(v, w);
    if (full) {
      const uint32_t a = tile + PT::thread_off(t);
#pragma unroll
      for (int i = 0; i < BW; ++i) sts64(a + i * 8, w[i]);
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the bulk-copy engine
      __syncthreads();
      if (t == 0) {
        uint8_t* dst = reinterpret_cast<uint8_t*>(enc) + (uint64_t)tl * PT::kGroups * BW * 8;
#pragma unroll
        for (int s = 0; s < 8; ++s)
          tma_store_1d(dst + (uint64_t)s * PT::kSegBytes, tile + s * PT::kSegStride, PT::kSegBytes);
        tma_store_commit();
      }
    } else if (g < ngroups) {
      uint64_t* dst = enc + g * BW;
#pragma unroll
      for (int i = 0; i < BW; ++i) dst[i] = w[i];
    }
    buf ^= 1;
  }
  if (t == 0) tma_store_wait<0>();  // all bulk stores complete before the CTA exits
}

// ------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------
namespace {

typedef void (*UnpackFn)(const uint64_t*, int64_t, int64_t, uint32_t*);
typedef void (*PackFn)(const uint32_t*, int64_t, int64_t, uint64_t*);

template <int... I>
struct Seq {};
template <int N, int... I>
struct MakeSeq : MakeSeq<N - 1, N, I...> {};
template <int... I>
struct MakeSeq<0, I...> { typedef Seq<I...> type; };

template <int... I>
UnpackFn unpack_table_entry(int bw, Seq<I...>) {
  static const UnpackFn tab[] = {unpack_kernel<I>...};
  return tab[bw - 1];
}
template <int... I>
PackFn pack_table_entry(int bw, Seq<I...>) {
  static const PackFn tab[] = {pack_kernel<I>...};
  return tab[bw - 1];
}
template <int... I>
UnpackFn unpack_tma_table_entry(int bw, Seq<I...>) {
  static const UnpackFn tab[] = {unpack_tma_kernel<I>...};
  return tab[bw - 1];
}
template <int... I>
UnpackFn unpack_warp_table_entry(int bw, Seq<I...>) {
  static const UnpackFn tab[] = {unpack_warp_kernel<I>...};
  return tab[bw - 1];
}
template <int... I>
UnpackFn unpack_bulk_table_entry(int bw, Seq<I...>) {
  static const UnpackFn tab[] = {unpack_bulk_kernel<I>...};
  return tab[bw - 1];
}
typedef void (*UnpackNarrowFn)(const uint64_t*, int64_t, int64_t, uint8_t*);
template <int... I>
UnpackNarrowFn unpack_u8_warp_table_entry(int bw, Seq<I...>) {
  static const UnpackNarrowFn tab[] = {unpack_narrow_warp_kernel<I, 1>...};
  return tab[bw - 1];
}
template <int... I>
UnpackNarrowFn unpack_u16_warp_table_entry(int bw, Seq<I...>) {
  static const UnpackNarrowFn tab[] = {unpack_narrow_warp_kernel<I, 2>...};
  return tab[bw - 1];
}
template <int... I>
UnpackNarrowFn unpack_u8_table_entry(int bw, Seq<I...>) {
  static const UnpackNarrowFn tab[] = {unpack_narrow_kernel<I, 1>...};
  return tab[bw - 1];
}
template <int... I>
UnpackNarrowFn unpack_u16_table_entry(int bw, Seq<I...>) {
  static const UnpackNarrowFn tab[] = {unpack_narrow_kernel<I, 2>...};
  return tab[bw - 1];
}
template <int... I>
PackFn pack_tma_table_entry(int bw, Seq<I...>) {
  static const PackFn tab[] = {pack_tma_kernel<I>...};
  return tab[bw - 1];
}
template <int... I>
uint32_t plane_tile_bytes(int bw, Seq<I...>) {
  static const uint32_t tab[] = {PlaneTile<I>::kBytes...};
  return tab[bw - 1];
}

// opt in to > 48 KB of dynamic shared memory once per kernel and device (process-wide)
int raise_smem_limit(fle_ctx* ctx, const void* kern, size_t smem_bytes) {
  if (smem_bytes <= 48 * 1024) return FLE_OK;
  static std::mutex mu;
  static std::set<std::pair<int, const void*> > raised;
  std::lock_guard<std::mutex> lock(mu);
  const std::pair<int, const void*> key(ctx->device, kern);
  if (raised.find(key) != raised.end()) return FLE_OK;
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, kern);
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "cudaFuncGetAttributes");
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin - (int)fa.sharedSizeBytes);
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
  raised.insert(key);
  return FLE_OK;
}

int codec_ctas(fle_ctx* ctx, size_t smem) {
  const int fit = std::max<int>(1, std::min<int>(8, (int)((size_t)ctx->max_smem_per_sm / (smem + 1024))));
  return ctx->codec_ctas_per_sm > 0 ? std::min(fit, ctx->codec_ctas_per_sm) : fit;
}

int codec_grid(fle_ctx* ctx, int64_t ngroups) {
  const int64_t ntiles = (ngroups + kCodecThreads - 1) / kCodecThreads;
  // 32 KB static shared memory per CTA -> up to 7 CTAs per SM
  return (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t)ctx->num_sms * 6));
}

}  // namespace

// Which unpack kernel and how many CTAs per SM (codec_variant = 1, "auto"): measured per bit
// width at 2^26 values (profiles/r01_codec_variants.txt).  More resident CTAs are NOT better here:
// with three or more CTAs per SM the staged kernels lose 5-8 % against two -- more concurrent read
// and write streams for the memory system, nothing left to hide -- except for the narrowest
// widths, whose tiles are nearly all output.
enum UnpackKind { UNPACK_STAGED = 4, UNPACK_BULK = 2, UNPACK_WARP = 3 };
static void unpack_choice(int bw, int64_t ngroups, int* kind, int* ctas) {
  // Short runs (the reference benchmark's own 16 Mi values: a 10-20 us kernel, 7 tiles per CTA) are about ramp-up
  // and the tile loop's latency, not steady-state bandwidth: the warp-private pipeline -- three stages, no CTA
  // barrier -- wins by 8-15 % from bw 4 to 28 (tools/codec_cold_probe.py 16777216: bw 8 0.63 -> 0.78, bw 16 0.75 -> 0.83)
  if (ngroups <= (1 << 19) && bw >= 3 && bw <= 28) { *kind = UNPACK_WARP; *ctas = 3; return; }
  if (bw <= 3) { *kind = UNPACK_WARP; *ctas = 2; }
  else if (bw <= 5) { *kind = UNPACK_STAGED; *ctas = 3; }
  else if (bw <= 8) { *kind = UNPACK_BULK; *ctas = 2; }
  else if (bw <= 21) { *kind = UNPACK_STAGED; *ctas = 2; }
  else { *kind = UNPACK_BULK; *ctas = 2; }  // (bw 29..32: warp-private copy-out measured equal warm, 9 % behind with a cold L2,
                                            //  tools/codec_cold_probe.py: bw 32 0.889 vs 0.975)
}

int launch_unpack_u32(fle_ctx* ctx, const uint64_t* enc_dev, int bw, int64_t n, uint32_t* out_dev) {
  if (bw < 1 || bw > 32) return set_error(ctx, FLE_E_INVALID, "bit width outside 1..32");
  if (n == 0) return FLE_OK;
  const int64_t G = (n + 63) / 64;
  const bool enc_ok = (reinterpret_cast<uintptr_t>(enc_dev) & 15) == 0;
  const bool out_ok = (reinterpret_cast<uintptr_t>(out_dev) & 15) == 0;
  if (ctx->codec_variant != 0 && enc_ok) {
    // codec_variant: 1 auto, 2 bulk stores, 3 warp-private copy-out, 4 staged copy-out (0: direct loads)
    int kind = ctx->codec_variant, ctas = 0;
    if (kind == 1) unpack_choice(bw, G, &kind, &ctas);
    if (!out_ok || (kind != UNPACK_BULK && kind != UNPACK_WARP)) kind = UNPACK_STAGED;
    if (ctx->codec_ctas_per_sm > 0) ctas = ctx->codec_ctas_per_sm;
    const size_t stage = (plane_tile_bytes(bw, MakeSeq<32>::type()) + 127) / 128 * 128;
    UnpackFn fn;
    size_t smem;
    if (kind == UNPACK_BULK) {
      fn = unpack_bulk_table_entry(bw, MakeSeq<32>::type());
      smem = kCodecThreads * kRowStride + unpack_bulk_stages(bw) * stage + 64;
    } else if (kind == UNPACK_WARP) {
      fn = unpack_warp_table_entry(bw, MakeSeq<32>::type());
      smem = kTileValueBytes + 3 * stage + 64;
    } else {
      fn = unpack_tma_table_entry(bw, MakeSeq<32>::type());
      smem = kTileValueBytes + 2 * stage + 64;
    }
    int rc = raise_smem_limit(ctx, (const void*)fn, smem);
    if (rc != FLE_OK) return rc;
    int per_sm = std::max<int>(1, std::min<int>(8, (int)((size_t)ctx->max_smem_per_sm / (smem + 1024))));
    if (ctas > 0) per_sm = std::min(per_sm, ctas);
    const int64_t ntiles = (G + kCodecThreads - 1) / kCodecThreads;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t)ctx->num_sms * per_sm));
    fn<<<grid, kCodecThreads, smem, ctx->stream>>>(enc_dev, G, n, out_dev);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(ctx, e, "unpack kernel launch");
    ctx->launches += 1;
    return FLE_OK;
  }
  UnpackFn fn = unpack_table_entry(bw, MakeSeq<32>::type());
  fn<<<codec_grid(ctx, G), kCodecThreads, 0, ctx->stream>>>(enc_dev, G, n, out_dev);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "unpack_kernel launch");
  ctx->launches += 1;
  return FLE_OK;
}

int launch_unpack_narrow(fle_ctx* ctx, const uint64_t* enc_dev, int bw, int64_t n, int elem_size, void* out_dev) {
  if (elem_size != 1 && elem_size != 2) return set_error(ctx, FLE_E_INVALID, "narrow unpack writes 1- or 2-byte values");
  if (bw < 1 || bw > 8 * elem_size) return set_error(ctx, FLE_E_INVALID, "bit width does not fit the output element");
  if (n == 0) return FLE_OK;
  const int64_t G = (n + 63) / 64;
  // codec_variant 4: the first form (CTA-wide copy-out between two barriers); anything else: warp-private copy-out.
  // Both want 16-byte aligned pointers (TMA source, 128-bit stores).
  if (((reinterpret_cast<uintptr_t>(enc_dev) | reinterpret_cast<uintptr_t>(out_dev)) & 15) != 0)
    return set_error(ctx, FLE_E_ALIGN, "narrow unpack wants 16-byte aligned buffers");
  const bool warp_form = ctx->codec_variant != 4;
  UnpackNarrowFn fn;
  if (warp_form) fn = elem_size == 1 ? unpack_u8_warp_table_entry(bw, MakeSeq<8>::type()) : unpack_u16_warp_table_entry(bw, MakeSeq<16>::type());
  else fn = elem_size == 1 ? unpack_u8_table_entry(bw, MakeSeq<8>::type()) : unpack_u16_table_entry(bw, MakeSeq<16>::type());
  const size_t stage = (plane_tile_bytes(bw, MakeSeq<32>::type()) + 127) / 128 * 128;
  const size_t smem = (size_t)kCodecThreads * 64 * elem_size + (warp_form ? 3 : 2) * stage + 64;
  int rc = raise_smem_limit(ctx, (const void*)fn, smem);
  if (rc != FLE_OK) return rc;
  int per_sm = std::max<int>(1, std::min<int>(8, (int)((size_t)ctx->max_smem_per_sm / (smem + 1024))));
  per_sm = std::min(per_sm, ctx->codec_ctas_per_sm > 0 ? ctx->codec_ctas_per_sm : (warp_form ? 6 : 4));
  const int64_t ntiles = (G + kCodecThreads - 1) / kCodecThreads;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t)ctx->num_sms * per_sm));
  fn<<<grid, kCodecThreads, smem, ctx->stream>>>(enc_dev, G, n, static_cast<uint8_t*>(out_dev));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "unpack_narrow kernel launch");
  ctx->launches += 1;
  return FLE_OK;
}

int launch_pack_u32(fle_ctx* ctx, const uint32_t* vals_dev, int64_t n, int bw, uint64_t* enc_dev) {
  if (bw < 1 || bw > 32) return set_error(ctx, FLE_E_INVALID, "bit width outside 1..32");
  const int64_t G = (n + 63) / 64;
  if (G == 0) {
    // Flush() packs one group even when nothing was Put (fle-encoding.h:9806-9811)
    FLE_CUDA_TRY(ctx, cudaMemsetAsync(enc_dev, 0, (size_t)bw * 8, ctx->stream));
    return FLE_OK;
  }
  if (ctx->codec_variant != 0 && ((reinterpret_cast<uintptr_t>(enc_dev) | reinterpret_cast<uintptr_t>(vals_dev)) & 15) == 0) {
    PackFn fn = pack_tma_table_entry(bw, MakeSeq<32>::type());
    const size_t smem = 2 * kPackBufBytes;
    int rc = raise_smem_limit(ctx, (const void*)fn, smem);
    if (rc != FLE_OK) return rc;
    // three CTAs fit, two are faster at every bit width (0.88-0.95 against 0.83-0.89 of roofline)
    const int per_sm = ctx->codec_ctas_per_sm > 0 ? codec_ctas(ctx, smem) : std::min(2, codec_ctas(ctx, smem));
    const int64_t ntiles = (G + kCodecThreads - 1) / kCodecThreads;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t)ctx->num_sms * per_sm));
    fn<<<grid, kCodecThreads, smem, ctx->stream>>>(vals_dev, n, G, enc_dev);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(ctx, e, "pack_tma_kernel launch");
    ctx->launches += 1;
    return FLE_OK;
  }
  PackFn fn = pack_table_entry(bw, MakeSeq<32>::type());
  fn<<<codec_grid(ctx, G), kCodecThreads, 0, ctx->stream>>>(vals_dev, n, G, enc_dev);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "pack_kernel launch");
  ctx->launches += 1;
  return FLE_OK;
}

}  // namespace fle
