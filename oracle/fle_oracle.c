/* TEST INFRASTRUCTURE ONLY -- see fle_oracle.h for scope, citations and parity status. */
#include "fle_oracle.h"

#include <stdlib.h>
#include <string.h>

static inline uint64_t low_mask(int bw) { return bw >= 64 ? ~0ULL : ((1ULL << bw) - 1); }

int64_t orc_encoded_bytes(int64_t n, int bw) {
  int64_t g = (n + 63) / 64;
  if (g < 1) g = 1; /* Flush packs a group even when nothing is staged, fle-encoding.h:9806 */
  return g * bw * 8;
}

/* fle-encoding.h:8338-8340: buffer_end_[l] |= ((value >> l) & 1) << (63 - count_) */
int64_t orc_pack(const uint32_t* vals, int64_t n, int bw, uint64_t* out) {
  int64_t bytes = orc_encoded_bytes(n, bw);
  memset(out, 0, (size_t)bytes);
  for (int64_t r = 0; r < n; ++r) {
    uint64_t v = vals[r];
    uint64_t* grp = out + (r >> 6) * bw;
    int pos = 63 - (int)(r & 63);
    for (int l = 0; l < bw; ++l) grp[l] |= ((v >> l) & 1ULL) << pos;
  }
  return bytes;
}

/* fle-encoding.h:426-430: current_value_[63 - l] |= ((buffer_end_[i] >> l) & 1) << i */
static inline uint32_t value_at(const uint64_t* words, int bw, int64_t row) {
  const uint64_t* grp = words + (row >> 6) * bw;
  int pos = 63 - (int)(row & 63);
  uint32_t v = 0;
  for (int i = 0; i < bw; ++i) v |= (uint32_t)((grp[i] >> pos) & 1ULL) << i;
  return v;
}

void orc_unpack(const uint64_t* words, int bw, int64_t first_row, int64_t n, uint32_t* out) {
  for (int64_t r = 0; r < n; ++r) out[r] = value_at(words, bw, first_row + r);
}

static inline int cmp_scalar(int op, uint64_t v, const uint64_t* lits, int nlits) {
  switch (op) {
    case ORC_EQ: return v == lits[0];
    case ORC_LT: return v < lits[0];
    case ORC_LE: return v <= lits[0];
    case ORC_GT: return v > lits[0];
    case ORC_GE: return v >= lits[0];
    case ORC_IN: {
      int f = 0;
      for (int j = 0; j < nlits; ++j) f = f || (v == lits[j]);
      return f;
    }
  }
  return 0;
}

int64_t orc_pred(int op, const uint64_t* words, int bw, int64_t first_row, int64_t num_rows,
                 const uint64_t* lits, int nlits, uint64_t* out) {
  int64_t nw = (num_rows + 63) / 64, cnt = 0;
  uint64_t m = low_mask(bw);
  uint64_t* ml = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(nlits > 0 ? nlits : 1));
  for (int j = 0; j < nlits; ++j) ml[j] = lits[j] & m; /* body uses bw literal bits only */
  memset(out, 0, (size_t)nw * 8);
  for (int64_t r = 0; r < num_rows; ++r) {
    uint64_t v = value_at(words, bw, first_row + r);
    if (cmp_scalar(op, v, ml, nlits)) {
      out[r >> 6] |= 1ULL << (r & 63);
      ++cnt;
    }
  }
  free(ml);
  return cnt;
}

/* fle-encoding.h:7989-7998 -- six swap steps: bit k <-> bit 63-k */
static inline uint64_t reverse64(uint64_t x) {
  x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
  x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
  x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
  x = ((x >> 8) & 0x00FF00FF00FF00FFULL) | ((x & 0x00FF00FF00FF00FFULL) << 8);
  x = ((x >> 16) & 0x0000FFFF0000FFFFULL) | ((x & 0x0000FFFF0000FFFFULL) << 16);
  return (x >> 32) | (x << 32);
}

static uint64_t group_mask(int op, const uint64_t* grp, int bw, const uint64_t* lits, int nlits) {
  uint64_t C[64];
  if (op == ORC_IN) {
    /* fle-encoding.h:8282-8290 */
    uint64_t many = 0;
    for (int v = 0; v < nlits; ++v) {
      uint64_t meq = ~0ULL;
      for (int i = 0; i < bw; ++i) {
        uint64_t c = ((lits[v] >> i) & 1ULL) ? ~0ULL : 0ULL;
        meq &= ~(grp[i] ^ c);
      }
      many |= meq;
    }
    return many;
  }
  for (int i = 0; i < bw; ++i) C[i] = ((lits[0] >> i) & 1ULL) ? ~0ULL : 0ULL; /* :7965-7971 */
  if (op == ORC_EQ) { /* :7985-7988 */
    uint64_t meq = ~0ULL;
    for (int i = 0; i < bw; ++i) meq &= ~(grp[i] ^ C[i]);
    return meq;
  }
  uint64_t mlt = 0, mgt = 0, meq = ~0ULL;
  for (int i = bw - 1; i >= 0; --i) { /* :8039-8042, :8151-8154 */
    mlt |= meq & C[i] & ~grp[i];
    mgt |= meq & ~C[i] & grp[i];
    meq &= ~(grp[i] ^ C[i]);
  }
  switch (op) {
    case ORC_LT: return mlt;
    case ORC_LE: return mlt | meq; /* :8100 */
    case ORC_GT: return mgt;
    case ORC_GE: return mgt | meq; /* :8212 */
  }
  return 0;
}

int64_t orc_pred_planes(int op, const uint64_t* words, int bw, int64_t first_row,
                        int64_t num_rows, const uint64_t* lits, int nlits, uint64_t* out) {
  int64_t g0 = first_row >> 6, cnt = 0, w = 0;
  while (num_rows > 0) {
    uint64_t m = reverse64(group_mask(op, words + (g0 + w) * bw, bw, lits, nlits));
    if (num_rows < 64) m &= (1ULL << num_rows) - 1; /* tail: only num_rows bits are pushed, :8000-8006 */
    out[w] = m;
    cnt += __builtin_popcountll(m);
    ++w;
    num_rows -= 64;
  }
  return cnt;
}

int64_t orc_intersect_bitset(uint64_t* root, int64_t n_root, const uint64_t* sub) {
  int64_t j = -1, cnt = 0;
  for (int64_t i = 0; i < n_root; ++i) {
    if ((root[i >> 6] >> (i & 63)) & 1ULL) {
      ++j;
      if (!((sub[j >> 6] >> (j & 63)) & 1ULL)) root[i >> 6] &= ~(1ULL << (i & 63));
      else ++cnt;
    }
  }
  return cnt;
}

/* ---- cursor-faithful decoder model ------------------------------------------------ */

void orc_decoder_init(orc_decoder* d, const uint64_t* words, int64_t nbytes, int bw) {
  d->words = words;
  d->ngroups_guard = nbytes; /* kept in BYTES: guard = buffer + buffer_len, fle-encoding.h:43 */
  d->bw = bw;
  d->count = 64;
  d->next_group = 0;
  d->cur_group = -1;
}

static inline int at_guard(const orc_decoder* d) {
  return d->next_group * d->bw * 8 >= d->ngroups_guard; /* buffer_end_ >= buffer_end_guard_ */
}

static inline void unpack_next(orc_decoder* d) { /* unpack_function_(); buffer_end_ += bw */
  d->cur_group = d->next_group;
  d->next_group += 1;
}

int orc_decoder_get(orc_decoder* d, uint32_t* val) { /* fle-encoding.h:404-567 */
  if (d->count == 64) {
    if (at_guard(d)) return 0;
    d->count = 0;
    unpack_next(d);
  }
  *val = value_at(d->words, d->bw, d->cur_group * 64 + d->count);
  d->count += 1;
  return 1;
}

static int advance(orc_decoder* d, int skip_rows) { /* shared by :344-363 and :381-399 */
  if (d->count == 64) {
    d->count = skip_rows & 63;
    int skip_64 = (skip_rows & ~63) >> 6;
    d->next_group += skip_64;
    if (at_guard(d)) return 0;
    unpack_next(d);
  } else {
    skip_rows += d->count;
    d->count = skip_rows & 63;
    if (skip_rows >= 64) {
      int skip_64 = ((skip_rows & ~63) >> 6) - 1;
      d->next_group += skip_64;
      if (at_guard(d)) return 0;
      unpack_next(d);
    }
  }
  return 1;
}

int orc_decoder_get_skip(orc_decoder* d, uint32_t* val, int skip_rows) {
  if (!advance(d, skip_rows)) return 0;
  *val = value_at(d->words, d->bw, d->cur_group * 64 + d->count);
  d->count += 1;
  return 1;
}

int orc_decoder_skip(orc_decoder* d, int skip_rows) { return advance(d, skip_rows); }

static inline void push_bit(uint64_t* out, int64_t* nbits, int b) {
  int64_t i = *nbits;
  if ((i & 63) == 0) out[i >> 6] = 0;
  if (b) out[i >> 6] |= 1ULL << (i & 63);
  *nbits = i + 1;
}

void orc_decoder_pred(const orc_decoder* d, int op, int64_t num_rows, const uint64_t* lits,
                      int nlits, uint64_t* out, int64_t* nbits) {
  /* head: rows left in the already-unpacked group, scalar compare against the FULL literal
   * (fle-encoding.h:7973-7981) */
  for (int i = d->count; i != 64 && num_rows > 0; ++i, --num_rows) {
    uint64_t v = value_at(d->words, d->bw, d->cur_group * 64 + i);
    push_bit(out, nbits, cmp_scalar(op, v, lits, nlits));
  }
  /* body: whole groups from buffer_end_ on, plane algebra, bw literal bits (:7983-8009) */
  int64_t g = d->next_group;
  while (num_rows > 0) {
    uint64_t m = reverse64(group_mask(op, d->words + g * d->bw, d->bw, lits, nlits));
    int64_t take = num_rows < 64 ? num_rows : 64;
    for (int64_t i = 0; i < take; ++i) push_bit(out, nbits, (int)((m >> i) & 1ULL));
    num_rows -= take;
    ++g;
  }
}

/* ---- dictionary literal translation ------------------------------------------------- */

static int64_t lower_bound_i64(const int64_t* a, int64_t n, int64_t v) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) / 2;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

static int64_t upper_bound_i64(const int64_t* a, int64_t n, int64_t v) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) / 2;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

int orc_dict_translate(int op, const int64_t* dict, int64_t n, int64_t val, uint64_t* code) {
  *code = 0;
  switch (op) {
    case ORC_EQ: { /* dict-encoding.h:461-470 */
      int64_t it = lower_bound_i64(dict, n, val);
      if (it == n || val < dict[it]) return ORC_NONE;
      *code = (uint64_t)it;
      return ORC_EQ;
    }
    case ORC_GT: /* :472-483 */
      if (dict[n - 1] <= val) return ORC_NONE;
      if (dict[0] > val) return ORC_ALL;
      *code = (uint64_t)upper_bound_i64(dict, n, val);
      return ORC_GE;
    case ORC_LT: /* :485-496 */
      if (dict[0] >= val) return ORC_NONE;
      if (dict[n - 1] < val) return ORC_ALL;
      *code = (uint64_t)lower_bound_i64(dict, n, val);
      return ORC_LT;
    case ORC_GE: /* :498-509 */
      if (dict[n - 1] < val) return ORC_NONE;
      if (dict[0] >= val) return ORC_ALL;
      *code = (uint64_t)lower_bound_i64(dict, n, val);
      return ORC_GE;
    case ORC_LE: /* :511-521 */
      if (dict[0] > val) return ORC_NONE;
      if (dict[n - 1] <= val) return ORC_ALL;
      *code = (uint64_t)upper_bound_i64(dict, n, val);
      return ORC_LT;
  }
  return ORC_NONE;
}

int64_t orc_dict_translate_in(const int64_t* dict, int64_t n, const int64_t* lits, int64_t nlits,
                              uint64_t* codes) {
  int64_t k = 0;
  for (int64_t i = 0; i < nlits; ++i) { /* dict-encoding.h:528-535 */
    int64_t it = lower_bound_i64(dict, n, lits[i]);
    if (it == n || lits[i] < dict[it]) continue;
    codes[k++] = (uint64_t)it;
  }
  return k;
}

int orc_dict_bit_width(int64_t num_entries) {
  if (num_entries == 0) return 0; /* dict-encoding.h:77 */
  if (num_entries == 1) return 1; /* :78 */
  uint64_t x = (uint64_t)num_entries - 1; /* BitUtil::Log2 = ceil(log2 x), bit-util.h:132-143 */
  int result = 1;
  while (x >>= 1) ++result;
  return result;
}
