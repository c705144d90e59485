// Dictionary build + encode of a whole column chunk on the device.
//
// Replaces, for fixed-width integer columns already resident in HBM,
//   DictEncoder<T>::Put x n          /root/reference/dict-encoding.h:234-274  (hash table, <= 40000 entries)
//   DictEncoder<T>::WriteDict        :393-406  (SORTS the dictionary, builds to_sorted_indice_)
//   DictEncoderBase::WriteData       :408-447  (remap indices, [bit width] + FLE groups)
// so that dictionary codes are order-preserving, which is what lets the scanner translate
// `col < literal` into `code < lower_bound(literal)` (dict-encoding.h:485-496).
//   1. insert   : open-addressing hash table (2^17 slots, linear probing, 64-bit atomicCAS);
//                 more than max_entries distinct values => FLE_E_DICT_FULL, the caller keeps the
//                 column PLAIN exactly like the writer does when Put() returns -1
//                 (hdfs-parquet-table-writer.cc:276-279)
//   2. compact  : occupied slots -> (key, slot) list
//   3. rank     : rank_i = #{j : key_j < key_i} by tiled all-pairs counting (<= 40000^2 compares,
//                 exact because the keys are distinct) -> sorted dictionary + slot -> code map
//   4. encode   : value -> slot (probe) -> code (u32)
//   5. pack     : pack_kernel<bw> with bw = DictEncoderBase::bit_width() (dict-encoding.h:76-80)
#include <mutex>
#include <set>

#include "fle_device.cuh"
#include "fle_internal.h"

namespace fle {

constexpr int kDictSlots = 1 << 17;
constexpr unsigned long long kEmptyKey = 0x8000000000000000ull;  // INT64_MIN is handled out of band

__device__ __forceinline__ uint32_t dict_hash(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return (uint32_t)k & (kDictSlots - 1);
}

template <typename T>
__device__ __forceinline__ unsigned long long widen(T v) { return (unsigned long long)(long long)v; }

struct DictState {
  unsigned int count;     // distinct keys inserted (without INT64_MIN)
  unsigned int overflow;  // set when count would exceed max_entries
  unsigned int has_min;   // INT64_MIN occurs in the column
  unsigned int nlist;     // compaction cursor
};

__global__ void __launch_bounds__(256) dict_clear_kernel(unsigned long long* __restrict__ keys) {
  keys[blockIdx.x * 256 + threadIdx.x] = kEmptyKey;
}

template <typename T>
__global__ void __launch_bounds__(256)
dict_insert_kernel(const T* __restrict__ vals, int64_t n, unsigned long long* __restrict__ keys, DictState* st,
                   unsigned int max_entries) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    if (*((volatile unsigned int*)&st->overflow)) return;
    const unsigned long long k = widen(vals[i]);
    if (k == kEmptyKey) { st->has_min = 1; continue; }
    uint32_t h = dict_hash(k);
    for (int probe = 0; probe < kDictSlots; ++probe) {
      const unsigned long long cur = keys[h];
      if (cur == k) break;
      if (cur == kEmptyKey) {
        const unsigned long long old = atomicCAS(&keys[h], kEmptyKey, k);
        if (old == kEmptyKey) {
          if (atomicAdd(&st->count, 1u) + 1u > max_entries) st->overflow = 1;
          break;
        }
        if (old == k) break;
      }
      h = (h + 1) & (kDictSlots - 1);
    }
  }
}

// The same insert behind a per-CTA front cache.  A column chunk has few distinct values (that is why it is dictionary-
// encoded), so nearly every probe of the global table finds a key that is already there -- 2^28 random 8-byte reads of
// a 1 MB table are 8.6 GB of L2 sector traffic, which is what the kernel above spends its time on.  Each persistent
// CTA keeps the keys it has met in a shared-memory set (2^14 slots, multiply-shift hash, at most kFrontMax keys: a
// crowded set only costs probes); only a value the CTA has not met before goes to the global table.
constexpr int kFrontUnroll = 8;  // independent 8-byte loads per thread: 64 KB in flight per SM
constexpr int kFrontBits = 14, kFrontSlots = 1 << kFrontBits, kFrontMax = kFrontSlots * 3 / 4, kFrontProbes = 8;

__device__ __forceinline__ uint32_t front_hash(unsigned long long k) {
  return (uint32_t)((k * 0x9E3779B97F4A7C15ull) >> (64 - kFrontBits));
}

__device__ __forceinline__ void dict_insert_global(unsigned long long k, unsigned long long* __restrict__ keys, DictState* st,
                                                   unsigned int max_entries) {
  uint32_t h = dict_hash(k);
  for (int probe = 0; probe < kDictSlots; ++probe) {
    const unsigned long long cur = keys[h];
    if (cur == k) return;
    if (cur == kEmptyKey) {
      const unsigned long long old = atomicCAS(&keys[h], kEmptyKey, k);
      if (old == kEmptyKey) {
        if (atomicAdd(&st->count, 1u) + 1u > max_entries) st->overflow = 1;
        return;
      }
      if (old == k) return;
    }
    h = (h + 1) & (kDictSlots - 1);
  }
}

template <typename T>
__global__ void __launch_bounds__(1024)
dict_insert_front_kernel(const T* __restrict__ vals, int64_t n, unsigned long long* __restrict__ keys, DictState* st,
                         unsigned int max_entries) {
  extern __shared__ unsigned long long front[];  // kFrontSlots keys
  __shared__ unsigned int front_count;
  for (int i = threadIdx.x; i < kFrontSlots; i += 1024) front[i] = kEmptyKey;
  if (threadIdx.x == 0) front_count = 0;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * 1024;
  for (int64_t i0 = (int64_t)blockIdx.x * 1024 + threadIdx.x; i0 < n; i0 += kFrontUnroll * stride) {
    if (*((volatile unsigned int*)&st->overflow)) return;
    unsigned long long k4[kFrontUnroll];
#pragma unroll
    for (int j = 0; j < kFrontUnroll; ++j) k4[j] = widen(vals[min(i0 + j * stride, n - 1)]);  // (unconditional: all loads in flight)
#pragma unroll
    for (int j = 0; j < kFrontUnroll; ++j) {
      if (i0 + j * stride >= n) break;
      const unsigned long long k = k4[j];
      if (k == kEmptyKey) { st->has_min = 1; continue; }
      uint32_t h = front_hash(k);
      bool known = false, placed = false;
      for (int probe = 0; probe < kFrontProbes; ++probe) {
        const unsigned long long cur = front[h];
        if (cur == k) { known = true; break; }
        if (cur == kEmptyKey) {
          if (*((volatile unsigned int*)&front_count) >= (unsigned)kFrontMax) break;
          const unsigned long long old = atomicCAS(&front[h], kEmptyKey, k);
          if (old == kEmptyKey) { placed = true; atomicAdd(&front_count, 1u); break; }
          if (old == k) { known = true; break; }
        }
        h = (h + 1) & (kFrontSlots - 1);
      }
      (void)placed;
      if (!known) dict_insert_global(k, keys, st, max_entries);
    }
  }
}

// value -> code through a shared-memory map built from the sorted dictionary (nd <= kFrontMax): no global probe at all
template <typename T>
__global__ void __launch_bounds__(1024)
dict_encode_front_kernel(const T* __restrict__ vals, int64_t n, const long long* __restrict__ dict_sorted, int nd,
                         uint32_t* __restrict__ codes) {
  extern __shared__ unsigned long long front[];  // kFrontSlots keys, then kFrontSlots 16-bit codes
  unsigned short* fcode = reinterpret_cast<unsigned short*>(front + kFrontSlots);
  __shared__ unsigned int min_code;
  for (int i = threadIdx.x; i < kFrontSlots; i += 1024) front[i] = kEmptyKey;
  if (threadIdx.x == 0) min_code = 0xFFFFFFFFu;
  __syncthreads();
  for (int c = threadIdx.x; c < nd; c += 1024) {
    const unsigned long long k = (unsigned long long)dict_sorted[c];
    if (k == kEmptyKey) { min_code = (unsigned)c; continue; }  // INT64_MIN (at most one entry): out of band
    uint32_t h = front_hash(k);
    for (;;) {  // (distinct keys, load <= 3/4: terminates)
      if (atomicCAS(&front[h], kEmptyKey, k) == kEmptyKey) { fcode[h] = (unsigned short)c; break; }
      h = (h + 1) & (kFrontSlots - 1);
    }
  }
  __syncthreads();
  const unsigned int mc = min_code;
  const int64_t stride = (int64_t)gridDim.x * 1024;
  for (int64_t i0 = (int64_t)blockIdx.x * 1024 + threadIdx.x; i0 < n; i0 += kFrontUnroll * stride) {
    unsigned long long k4[kFrontUnroll];
#pragma unroll
    for (int j = 0; j < kFrontUnroll; ++j) k4[j] = widen(vals[min(i0 + j * stride, n - 1)]);
#pragma unroll
    for (int j = 0; j < kFrontUnroll; ++j) {
      if (i0 + j * stride >= n) break;
      const unsigned long long k = k4[j];
      uint32_t code = mc;
      if (k != kEmptyKey) {
        uint32_t h = front_hash(k);
        while (front[h] != k) h = (h + 1) & (kFrontSlots - 1);  // (every value of the column is in the dictionary)
        code = fcode[h];
      }
      codes[i0 + j * stride] = code;
    }
  }
}

__global__ void __launch_bounds__(256)
dict_compact_kernel(const unsigned long long* __restrict__ keys, DictState* st, long long* __restrict__ list_keys,
                    uint32_t* __restrict__ list_slots) {
  const int s = blockIdx.x * 256 + threadIdx.x;
  if (s == 0 && st->has_min) {  // INT64_MIN: pseudo slot kDictSlots
    const unsigned int p = atomicAdd(&st->nlist, 1u);
    list_keys[p] = (long long)kEmptyKey;
    list_slots[p] = kDictSlots;
  }
  if (s < kDictSlots && keys[s] != kEmptyKey) {
    const unsigned int p = atomicAdd(&st->nlist, 1u);
    list_keys[p] = (long long)keys[s];
    list_slots[p] = (uint32_t)s;
  }
}

__global__ void __launch_bounds__(256)
dict_rank_kernel(const long long* __restrict__ list_keys, const uint32_t* __restrict__ list_slots, int nd,
                 long long* __restrict__ dict_sorted, uint32_t* __restrict__ slot_to_code) {
  __shared__ long long tile[1024];
  const int i = blockIdx.x * 256 + threadIdx.x;
  const long long mine = i < nd ? list_keys[i] : 0;
  uint32_t rank = 0;
  for (int base = 0; base < nd; base += 1024) {
    __syncthreads();
    for (int j = threadIdx.x; j < 1024; j += 256) tile[j] = base + j < nd ? list_keys[base + j] : 0x7fffffffffffffffll;
    __syncthreads();
    const int m = min(1024, nd - base);
#pragma unroll 8
    for (int j = 0; j < m; ++j) rank += tile[j] < mine;
  }
  if (i < nd) {
    dict_sorted[rank] = mine;
    slot_to_code[list_slots[i]] = rank;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
dict_encode_kernel(const T* __restrict__ vals, int64_t n, const unsigned long long* __restrict__ keys,
                   const uint32_t* __restrict__ slot_to_code, uint32_t* __restrict__ codes) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const unsigned long long k = widen(vals[i]);
    uint32_t h = kDictSlots;
    if (k != kEmptyKey) {
      h = dict_hash(k);
      while (keys[h] != k) h = (h + 1) & (kDictSlots - 1);
    }
    codes[i] = slot_to_code[h];
  }
}

int launch_dict_encode(fle_ctx* ctx, const void* values_dev, int elem_type, int64_t n, int max_entries,
                       int64_t* dict_out_dev, int* num_entries, int* bit_width, uint8_t* enc_dev, size_t enc_cap,
                       size_t* enc_len) {
  if (elem_type != FLE_T_INT32 && elem_type != FLE_T_INT64) return set_error(ctx, FLE_E_UNSUPPORTED, "dictionary build handles INT32/INT64 columns");
  if (n < 0 || (n > 0 && !values_dev) || !dict_out_dev || !num_entries || !bit_width || !enc_dev || max_entries < 1 ||
      max_entries > 40000)
    return set_error(ctx, FLE_E_INVALID, "bad dict_encode arguments");
  const int es = elem_type == FLE_T_INT32 ? 4 : 8;
  if ((reinterpret_cast<uintptr_t>(values_dev) & (es - 1)) || (reinterpret_cast<uintptr_t>(enc_dev) & 15))
    return set_error(ctx, FLE_E_ALIGN, "values must be naturally aligned, enc_dev 16-byte aligned");
  cudaStream_t st = ctx->stream;
  // one grow-only scratch block per context (stream-ordered allocations go back to the driver at every
  // synchronisation, and this call synchronises: allocating cost more than the kernels)
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t o_keys = 0, o_state = o_keys + up((size_t)kDictSlots * 8), o_lkeys = o_state + up(sizeof(DictState)),
               o_lslots = o_lkeys + up((size_t)(max_entries + 2) * 8), o_s2c = o_lslots + up((size_t)(max_entries + 2) * 4),
               o_codes = o_s2c + up((size_t)(kDictSlots + 1) * 4), total = o_codes + up((size_t)std::max<int64_t>(n, 1) * 4);
  if (total > ctx->dict_cap) {
    FLE_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (ctx->d_dict) cudaFree(ctx->d_dict);
    ctx->d_dict = nullptr;
    ctx->dict_cap = 0;
    FLE_CUDA_TRY(ctx, cudaMalloc(&ctx->d_dict, total));
    ctx->dict_cap = total;
  }
  uint8_t* base = static_cast<uint8_t*>(ctx->d_dict);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(base + o_keys);
  DictState* state = reinterpret_cast<DictState*>(base + o_state);
  long long* list_keys = reinterpret_cast<long long*>(base + o_lkeys);
  uint32_t* list_slots = reinterpret_cast<uint32_t*>(base + o_lslots);
  uint32_t* slot_to_code = reinterpret_cast<uint32_t*>(base + o_s2c);
  uint32_t* codes = reinterpret_cast<uint32_t*>(base + o_codes);
  auto cleanup = [&]() {};
#define DICT_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return set_cuda_error(ctx, _e, #expr); } } while (0)
  // (no byte pattern gives kEmptyKey, so cudaMemset cannot fill the table)
  dict_clear_kernel<<<kDictSlots / 256, 256, 0, st>>>(keys);
  DICT_TRY(cudaMemsetAsync(state, 0, sizeof(DictState), st));
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * 8, (n + 255) / 256));
  // the kernels with a shared-memory front cache: one CTA of 1024 threads per SM; worth it from a few CTAs' worth of values on
  const size_t front_smem = (size_t)kFrontSlots * 8, front_map_smem = (size_t)kFrontSlots * 10;
  const bool front = n >= (1 << 20) && (size_t)ctx->max_smem_optin >= front_map_smem + 1024;
  const int front_grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms, (n + 4095) / 4096));
  if (front) {
    static std::mutex mu;
    static std::set<int> raised;
    std::lock_guard<std::mutex> lock(mu);
    if (raised.find(ctx->device) == raised.end()) {
      DICT_TRY(cudaFuncSetAttribute(dict_insert_front_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)front_smem));
      DICT_TRY(cudaFuncSetAttribute(dict_insert_front_kernel<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)front_smem));
      DICT_TRY(cudaFuncSetAttribute(dict_encode_front_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)front_map_smem));
      DICT_TRY(cudaFuncSetAttribute(dict_encode_front_kernel<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)front_map_smem));
      raised.insert(ctx->device);
    }
  }
  if (n > 0) {
    // The front cache pays while a CTA's set holds (nearly) every key of the column; a column with tens of thousands
    // of distinct values overflows it and is better off with twice the warps probing the global table.  The first
    // 4 Mi values tell: their distinct count picks the kernel for the rest.
    auto insert_range = [&](int64_t first, int64_t count, bool use_front) {
      const uint8_t* base = static_cast<const uint8_t*>(values_dev) + (size_t)first * es;
      const int fg = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms, (count + 4095) / 4096));
      const int gg = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->num_sms * 8, (count + 255) / 256));
      if (use_front) {
        if (es == 4) dict_insert_front_kernel<int32_t><<<fg, 1024, front_smem, st>>>((const int32_t*)base, count, keys, state, (unsigned)max_entries);
        else dict_insert_front_kernel<int64_t><<<fg, 1024, front_smem, st>>>((const int64_t*)base, count, keys, state, (unsigned)max_entries);
      } else if (es == 4) {
        dict_insert_kernel<int32_t><<<gg, 256, 0, st>>>((const int32_t*)base, count, keys, state, (unsigned)max_entries);
      } else {
        dict_insert_kernel<int64_t><<<gg, 256, 0, st>>>((const int64_t*)base, count, keys, state, (unsigned)max_entries);
      }
      ctx->launches += 1;
    };
    constexpr int64_t kSample = 1 << 22;
    if (!front) {
      insert_range(0, n, false);
    } else if (n <= 2 * kSample) {
      insert_range(0, n, true);
    } else {
      insert_range(0, kSample, true);
      DictState hs;
      DICT_TRY(cudaMemcpyAsync(&hs, state, sizeof(hs), cudaMemcpyDeviceToHost, st));
      DICT_TRY(cudaStreamSynchronize(st));
      insert_range(kSample, n - kSample, !hs.overflow && hs.count <= (unsigned)kFrontMax / 2);
    }
  }
  DictState h;
  DICT_TRY(cudaMemcpyAsync(&h, state, sizeof(h), cudaMemcpyDeviceToHost, st));
  DICT_TRY(cudaStreamSynchronize(st));
  const int nd = (int)h.count + (h.has_min ? 1 : 0);
  if (h.overflow || nd > max_entries) {
    cleanup();
    *num_entries = nd;
    return set_error(ctx, FLE_E_DICT_FULL, "more distinct values than dictionary entries: keep the column PLAIN");
  }
  *num_entries = nd;
  int bw = 0;  // DictEncoderBase::bit_width(), dict-encoding.h:76-80
  if (nd == 1) bw = 1;
  else if (nd > 1) { bw = 1; unsigned x = (unsigned)nd - 1; while (x >>= 1) ++bw; }
  *bit_width = bw;
  const size_t need = bw == 0 ? 0 : fle_encoded_bytes(n, bw);
  if (enc_len) *enc_len = need;
  if (enc_cap < need) { cleanup(); return set_error(ctx, FLE_E_BUFFER, "encode buffer too small"); }
  if (nd > 0) {
    dict_compact_kernel<<<kDictSlots / 256, 256, 0, st>>>(keys, state, list_keys, list_slots);
    dict_rank_kernel<<<(nd + 255) / 256, 256, 0, st>>>(list_keys, list_slots, nd, (long long*)dict_out_dev, slot_to_code);
    if (front && nd <= kFrontMax) {
      if (es == 4) dict_encode_front_kernel<int32_t><<<front_grid, 1024, front_map_smem, st>>>((const int32_t*)values_dev, n, (const long long*)dict_out_dev, nd, codes);
      else dict_encode_front_kernel<int64_t><<<front_grid, 1024, front_map_smem, st>>>((const int64_t*)values_dev, n, (const long long*)dict_out_dev, nd, codes);
    } else if (es == 4) {
      dict_encode_kernel<int32_t><<<grid, 256, 0, st>>>((const int32_t*)values_dev, n, keys, slot_to_code, codes);
    } else {
      dict_encode_kernel<int64_t><<<grid, 256, 0, st>>>((const int64_t*)values_dev, n, keys, slot_to_code, codes);
    }
    ctx->launches += 3;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { cleanup(); return set_cuda_error(ctx, e, "dict kernels"); }
    const int rc = launch_pack_u32(ctx, codes, n, bw, reinterpret_cast<uint64_t*>(enc_dev));
    if (rc != FLE_OK) { cleanup(); return rc; }
  }
  cleanup();
#undef DICT_TRY
  return FLE_OK;
}

}  // namespace fle
