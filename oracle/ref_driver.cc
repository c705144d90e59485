// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// Thin extern "C" driver around the UNMODIFIED reference header
// /root/reference/fle-encoding.h (included from where it lies with -I/root/reference;
// no reference source is copied into this repo).  It is compiled by oracle/Makefile
// into oracle/_ref/libfle_ref.so and is used
//   * by tests / oracle/gen_golden.py to pin the C restatement (oracle/fle_oracle.c)
//     and to generate the golden vectors under tests/golden/, and
//   * by bench.py as the "reference" CPU baseline (kind = "reference").
//
// Everything here only *calls* reference entry points:
//   FleEncoder::Put / Flush            fle-encoding.h:8315 / :9806
//   FleDecoder::Get(T*)                fle-encoding.h:404
//   FleDecoder::Get(T*, int) / Skip    fle-encoding.h:344 / :381
//   FleDecoder::Eq/Lt/Le/Gt/Ge/In      fle-encoding.h:7962..8313
// plus two loops that restate how the scanner drives them
//   EvalSimplePredicates window loop   hdfs-parquet-scanner.cc:1837-1865
//   AssembleRows survivor loop         hdfs-parquet-scanner.cc:1101-1181
//   And/OrOperate                      simple-predicates.h:145-163
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "fle-encoding.h"  // the reference, unmodified

using impala::FleDecoder;
using impala::FleEncoder;
typedef boost::dynamic_bitset<> Bitset;

namespace {

enum { OP_EQ = 0, OP_LT = 1, OP_LE = 2, OP_GT = 3, OP_GE = 4, OP_IN = 5 };

// The reference's Unpack_* start with a 32-byte load even for an 8-byte group
// (fle-encoding.h:570), so decoders are always run on a padded private copy.
struct PaddedCopy {
  std::vector<uint8_t> buf;
  PaddedCopy(const uint8_t* src, int64_t len) : buf(len + 320, 0) { memcpy(buf.data(), src, len); }
  uint8_t* data() { return buf.data(); }
};

void RunPred(FleDecoder& d, int op, int64_t num_rows, Bitset& bs, const uint64_t* lits, int nlits) {
  switch (op) {
    case OP_EQ: d.Eq(num_rows, bs, lits[0]); break;
    case OP_LT: d.Lt(num_rows, bs, lits[0]); break;
    case OP_LE: d.Le(num_rows, bs, lits[0]); break;
    case OP_GT: d.Gt(num_rows, bs, lits[0]); break;
    case OP_GE: d.Ge(num_rows, bs, lits[0]); break;
    case OP_IN: {
      std::vector<uint64_t> v(lits, lits + nlits);
      d.In(num_rows, bs, v);
      break;
    }
  }
}

void CopyBits(const Bitset& bs, uint64_t* out_words) {
  boost::to_block_range(bs, out_words);  // real boost API, also provided by both stand-ins
}

}  // namespace

extern "C" {

int ref_has_avx2() { return __builtin_cpu_supports("avx2"); }

// FleEncoder::Put x n, Flush.  Returns the byte length Flush reports, or -1 if Put
// returned false (buffer too small: the encoder needs one spare group, fle-encoding.h:251).
int64_t ref_pack(const uint32_t* vals, int64_t n, int bw, uint8_t* out, int64_t out_cap) {
  FleEncoder enc(out, (int)out_cap, bw);
  for (int64_t i = 0; i < n; ++i) {
    if (!enc.Put(vals[i])) return -1;
  }
  return enc.Flush();
}

// FleDecoder::Get x n.  Returns number of values decoded before Get returned false.
int64_t ref_unpack(const uint8_t* enc, int64_t enc_len, int bw, int64_t n, uint32_t* out) {
  PaddedCopy pc(enc, enc_len);
  FleDecoder dec(pc.data(), (int)enc_len, bw);
  for (int64_t i = 0; i < n; ++i) {
    uint32_t v = 0;
    if (!dec.Get(&v)) return i;
    out[i] = v;
  }
  return n;
}

// Sequence of Get(val, skip[i]) calls (fle-encoding.h:344).  Returns count of successes.
int64_t ref_get_skip(const uint8_t* enc, int64_t enc_len, int bw, const int32_t* skips,
                     int64_t nskips, uint32_t* out) {
  PaddedCopy pc(enc, enc_len);
  FleDecoder dec(pc.data(), (int)enc_len, bw);
  for (int64_t i = 0; i < nskips; ++i) {
    uint32_t v = 0;
    if (!dec.Get(&v, skips[i])) return i;
    out[i] = v;
  }
  return nskips;
}

// Sequence of Skip(skips[i]) calls (fle-encoding.h:381): results[i] = what each returned; *final_get =
// what a Get() after the last one returns.  Pins the end-of-data behaviour of Skip().
int64_t ref_skip_seq(const uint8_t* enc, int64_t enc_len, int bw, const int32_t* skips, int64_t nskips,
                     int32_t* results, int32_t* final_get) {
  PaddedCopy pc(enc, enc_len);
  FleDecoder dec(pc.data(), (int)enc_len, bw);
  for (int64_t i = 0; i < nskips; ++i) results[i] = dec.Skip(skips[i]) ? 1 : 0;
  uint32_t v = 0;
  *final_get = dec.Get(&v) ? 1 : 0;
  return nskips;
}

// Built against the drop-in headers this switches them to reproduce the reference exactly
// (FleGpu::SetReferenceCompat); built against the reference itself it is a no-op.
void ref_set_compat(int on) {
#ifdef FLE_B200_HOST_GPU_CONTEXT_H
  impala::FleGpu::SetReferenceCompat(on != 0);
#else
  (void)on;
#endif
}

// Advance the cursor with `pre_get` plain Get() calls, then run ONE predicate call for
// num_rows rows.  Output: LSB-first 64-bit blocks; returns the bitset size.
int64_t ref_pred(int op, const uint8_t* enc, int64_t enc_len, int bw, int64_t pre_get,
                 int64_t num_rows, const uint64_t* lits, int nlits, uint64_t* out_words) {
  PaddedCopy pc(enc, enc_len);
  FleDecoder dec(pc.data(), (int)enc_len, bw);
  for (int64_t i = 0; i < pre_get; ++i) {
    uint32_t v;
    dec.Get(&v);
  }
  Bitset bs;
  RunPred(dec, op, num_rows, bs, lits, nlits);
  CopyBits(bs, out_words);
  return (int64_t)bs.size();
}

// Scanner-faithful window loop for one column and one leaf: predicate on <= `window`
// rows, then Skip(window) (hdfs-parquet-scanner.cc:1837-1865 + :1125-1131).
int64_t ref_pred_windowed(int op, const uint8_t* enc, int64_t enc_len, int bw, int64_t n,
                          int window, const uint64_t* lits, int nlits, uint64_t* out_words) {
  PaddedCopy pc(enc, enc_len);
  FleDecoder dec(pc.data(), (int)enc_len, bw);
  Bitset all;
  int64_t done = 0;
  while (done < n) {
    int64_t w = std::min<int64_t>(window, n - done);
    Bitset bs;
    RunPred(dec, op, w, bs, lits, nlits);
    for (int64_t i = 0; i < w; ++i) all.push_back(bs[i]);
    dec.Skip((int)w);
    done += w;
  }
  CopyBits(all, out_words);
  return (int64_t)all.size();
}

// A compound tree over several FLE columns, evaluated exactly as the reference's
// SimplePredicate tree does (simple-predicates.h:145-163): one full pass per leaf into a
// fresh bitset, `&=` / `|=` at inner nodes.  Postfix program: node kinds 0..5 = leaf
// (op, col, literal range), 6 = AND, 7 = OR.
struct RefNode {
  int32_t kind;      // 0..5 leaf op, 6 AND, 7 OR
  int32_t col;       // leaf: column index
  int32_t lit_off;   // leaf: offset into lits[]
  int32_t nlits;     // leaf: literal count (1 unless IN)
};

int64_t ref_eval_tree(const uint8_t* const* encs, const int64_t* enc_lens, const int32_t* bws,
                      int ncols, const RefNode* nodes, int nnodes, const uint64_t* lits,
                      int64_t n, uint64_t* out_words) {
  std::vector<PaddedCopy*> copies;
  for (int c = 0; c < ncols; ++c) copies.push_back(new PaddedCopy(encs[c], enc_lens[c]));
  std::vector<Bitset> stack;
  for (int i = 0; i < nnodes; ++i) {
    const RefNode& nd = nodes[i];
    if (nd.kind <= OP_IN) {
      FleDecoder dec(copies[nd.col]->data(), (int)enc_lens[nd.col], bws[nd.col]);
      Bitset bs;
      RunPred(dec, nd.kind, n, bs, lits + nd.lit_off, nd.nlits);
      stack.push_back(bs);
    } else {
      Bitset b = stack.back();
      stack.pop_back();
      if (nd.kind == 6) stack.back() &= b; else stack.back() |= b;
    }
  }
  CopyBits(stack.back(), out_words);
  int64_t sz = (int64_t)stack.back().size();
  for (auto* p : copies) delete p;
  return sz;
}

// AssembleRows survivor loop for ONE column (hdfs-parquet-scanner.cc:1101-1181):
// per window: predicate -> bitset -> skip distances -> Get(val, skip) per survivor ->
// trailing Skip(last_skip_rows); an all-zero window is skipped wholesale.
// Writes the surviving codes; returns how many.
int64_t ref_scan_materialize(int op, const uint8_t* enc, int64_t enc_len, int bw, int64_t n,
                             int window, const uint64_t* lits, int nlits, uint32_t* out_vals) {
  PaddedCopy pc(enc, enc_len);
  FleDecoder dec(pc.data(), (int)enc_len, bw);
  int64_t done = 0, nout = 0;
  while (done < n) {
    int64_t w = std::min<int64_t>(window, n - done);
    Bitset bs;
    RunPred(dec, op, w, bs, lits, nlits);
    if (bs.count() == 0) {
      dec.Skip((int)w);
      done += w;
      continue;
    }
    int start = 0;
    std::vector<int> skip_rows;
    for (int i = 0; i < (int)w; ++i) {
      if (bs[i]) {
        skip_rows.push_back(i - start);
        start = i + 1;
      }
    }
    int last_skip_rows = (int)w - start;
    for (int s : skip_rows) {
      uint32_t v = 0;
      if (!dec.Get(&v, s)) return -(nout + 1);
      out_vals[nout++] = v;
    }
    if (last_skip_rows != 0) dec.Skip(last_skip_rows);
    done += w;
  }
  return nout;
}

// The same loop with the return values honoured the way AssembleRows honours them: when the trailing
// Skip(last_skip_rows) reports false the scan RETURNS (`if (!result) return parse_status_;`,
// hdfs-parquet-scanner.cc:1107-1113) -- which is what happens when a page holds a multiple of 64 values
// and the skip lands exactly on its end.  Returns the number of codes written before that.
int64_t ref_scan_materialize_checked(int op, const uint8_t* enc, int64_t enc_len, int bw, int64_t n,
                                     int window, const uint64_t* lits, int nlits, uint32_t* out_vals) {
  PaddedCopy pc(enc, enc_len);
  FleDecoder dec(pc.data(), (int)enc_len, bw);
  int64_t done = 0, nout = 0;
  while (done < n) {
    int64_t w = std::min<int64_t>(window, n - done);
    Bitset bs;
    RunPred(dec, op, w, bs, lits, nlits);
    if (bs.count() == 0) {
      dec.Skip((int)w);  // (result not looked at here, :1125-1131)
      done += w;
      continue;
    }
    int start = 0;
    std::vector<int> skip_rows;
    for (int i = 0; i < (int)w; ++i) {
      if (bs[i]) {
        skip_rows.push_back(i - start);
        start = i + 1;
      }
    }
    int last_skip_rows = (int)w - start;
    for (int s : skip_rows) {
      uint32_t v = 0;
      if (!dec.Get(&v, s)) return nout;
      out_vals[nout++] = v;
    }
    if (last_skip_rows != 0 && !dec.Skip(last_skip_rows)) return nout;
    done += w;
  }
  return nout;
}

// ---------------------------------------------------------------------------------
// CPU baseline timing (bench.py `cpu_baseline` / `--impl reference`).  Rows are split
// into contiguous, group-aligned ranges, one FleDecoder per thread, as the reference's
// one-scanner-per-thread model would (hdfs-parquet-scanner.cc:1056-1060).
// mode 0: whole-range predicate calls; mode 1: 1024-row windows + Skip (scanner loop).
// The leaf list is ANDed (BETWEEN = Ge & Le: two full passes + `&=`).
// Returns seconds for ONE pass (best of `reps`); *match_count gets the AND popcount.
// ---------------------------------------------------------------------------------
struct RefLeaf {
  int32_t op;
  int32_t col;
  uint64_t lit;
};

static void RangeAndCount(const uint8_t* const* encs, const int32_t* bws, const RefLeaf* leaves,
                          int nleaves, int64_t g0, int64_t g1, int64_t n, int mode,
                          uint64_t* count_out) {
  int64_t row0 = g0 * 64, row1 = std::min<int64_t>(g1 * 64, n);
  int64_t rows = row1 - row0;
  uint64_t cnt = 0;
  if (rows <= 0) { *count_out = 0; return; }
  std::vector<FleDecoder*> decs;
  for (int l = 0; l < nleaves; ++l) {
    int c = leaves[l].col;
    uint8_t* p = const_cast<uint8_t*>(encs[c]) + g0 * bws[c] * 8;
    decs.push_back(new FleDecoder(p, (int)std::min<int64_t>((g1 - g0) * bws[c] * 8, 0x7fffffff), bws[c]));
  }
  if (mode == 0) {
    // chunks of 2^20 rows keep the reference's int-sized arithmetic in range
    const int64_t CH = 1 << 20;
    for (int64_t off = 0; off < rows; off += CH) {
      int64_t w = std::min(CH, rows - off);
      Bitset acc;
      for (int l = 0; l < nleaves; ++l) {
        int c = leaves[l].col;
        uint8_t* p = const_cast<uint8_t*>(encs[c]) + (g0 + off / 64) * bws[c] * 8;
        FleDecoder d(p, (int)(((w + 63) / 64) * bws[c] * 8), bws[c]);
        if (l == 0) {
          RunPred(d, leaves[l].op, w, acc, &leaves[l].lit, 1);
        } else {
          Bitset tmp;
          RunPred(d, leaves[l].op, w, tmp, &leaves[l].lit, 1);
          acc &= tmp;
        }
      }
      cnt += acc.count();
    }
  } else {
    for (int64_t off = 0; off < rows; off += 1024) {
      int64_t w = std::min<int64_t>(1024, rows - off);
      Bitset acc;
      for (int l = 0; l < nleaves; ++l) {
        if (l == 0) {
          RunPred(*decs[l], leaves[l].op, w, acc, &leaves[l].lit, 1);
        } else {
          Bitset tmp;
          RunPred(*decs[l], leaves[l].op, w, tmp, &leaves[l].lit, 1);
          acc &= tmp;
        }
      }
      cnt += acc.count();
      for (int l = 0; l < nleaves; ++l) decs[l]->Skip((int)w);
    }
  }
  for (auto* d : decs) delete d;
  *count_out = cnt;
}

double ref_time_and_leaves(const uint8_t* const* encs, const int32_t* bws, int ncols,
                           const RefLeaf* leaves, int nleaves, int64_t n, int nthreads,
                           int mode, int reps, uint64_t* match_count) {
  (void)ncols;
  int64_t G = (n + 63) / 64;
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    std::vector<uint64_t> counts(nthreads, 0);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
      int64_t g0 = G * t / nthreads, g1 = G * (t + 1) / nthreads;
      th.emplace_back(RangeAndCount, encs, bws, leaves, nleaves, g0, g1, n, mode, &counts[t]);
    }
    for (auto& x : th) x.join();
    auto t1 = std::chrono::steady_clock::now();
    double s = std::chrono::duration<double>(t1 - t0).count();
    if (s < best) best = s;
    uint64_t tot = 0;
    for (uint64_t c : counts) tot += c;
    *match_count = tot;
  }
  return best;
}

// Encode (Put+Flush) and decode (Get) timing, one encoder/decoder per thread over
// disjoint value ranges (multiples of 64).  Returns seconds (best of reps).
double ref_time_pack(const uint32_t* vals, int64_t n, int bw, int nthreads, int reps) {
  double best = 1e30;
  int64_t G = (n + 63) / 64;
  std::vector<std::vector<uint8_t>> outs(nthreads);
  for (int t = 0; t < nthreads; ++t) {
    int64_t g0 = G * t / nthreads, g1 = G * (t + 1) / nthreads;
    outs[t].resize((g1 - g0 + 2) * bw * 8 + 64);
  }
  for (int r = 0; r < reps; ++r) {
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
      int64_t g0 = G * t / nthreads, g1 = G * (t + 1) / nthreads;
      int64_t r0 = g0 * 64, r1 = std::min<int64_t>(g1 * 64, n);
      th.emplace_back([=, &outs]() {
        FleEncoder enc(outs[t].data(), (int)outs[t].size(), bw);
        for (int64_t i = r0; i < r1; ++i) enc.Put(vals[i]);
        enc.Flush();
      });
    }
    for (auto& x : th) x.join();
    auto t1 = std::chrono::steady_clock::now();
    best = std::min(best, std::chrono::duration<double>(t1 - t0).count());
  }
  return best;
}

double ref_time_unpack(const uint8_t* enc, int64_t n, int bw, int nthreads, int reps,
                       uint32_t* out) {
  double best = 1e30;
  int64_t G = (n + 63) / 64;
  for (int r = 0; r < reps; ++r) {
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
      int64_t g0 = G * t / nthreads, g1 = G * (t + 1) / nthreads;
      int64_t r0 = g0 * 64, r1 = std::min<int64_t>(g1 * 64, n);
      th.emplace_back([=]() {
        // last thread's range ends at the buffer end: the caller pads the buffer by 64 B
        FleDecoder dec(const_cast<uint8_t*>(enc) + g0 * bw * 8, (int)((g1 - g0) * bw * 8), bw);
        for (int64_t i = r0; i < r1; ++i) {
          uint32_t v;
          dec.Get(&v);
          out[i] = v;
        }
      });
    }
    for (auto& x : th) x.join();
    auto t1 = std::chrono::steady_clock::now();
    best = std::min(best, std::chrono::duration<double>(t1 - t0).count());
  }
  return best;
}

}  // extern "C"
