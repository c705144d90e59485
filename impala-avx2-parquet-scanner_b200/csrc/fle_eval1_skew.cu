// Skew-capable twins of the single-pass kernels (fle_eval1.cu) for columns that do not start on a 16-byte
// boundary -- the FLE section of a data page inside a column-chunk image starts one byte after the page's width
// header (dict-encoding.h:185-192, hdfs-parquet-scanner.cc:882-919).  The TMA producer stages the 16-byte
// aligned superset of every tile segment and the consumers rebuild the plane words with 32-bit shared loads and a
// funnel shift, so a page image is evaluated in place: every encoded byte is read from HBM once and nothing is
// written but the bitmap (before: fle_realign_dev copied the column first, 3x the traffic).
// Dictionary pages only reach 16 bits (dict-encoding.h:157), so only those widths are instantiated.
#include <algorithm>

#include "fle_eval1.cuh"
#include "fle_internal.h"

namespace fle {

template <bool IS_LT, int K, int BW>
__global__ void __launch_bounds__(256) eval1_skew_kernel(const __grid_constant__ EvalParams P) {
  eval1_body<PassRegs<IS_LT, K, BW>, BW, true>(P);
}

template <int BW>
__global__ void __launch_bounds__(256) eval_in_mux_skew_kernel(const __grid_constant__ EvalParams P) {
  eval1_body<InMuxRegs<BW>, BW, true>(P);
}

#define FLE_BW16(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)

template <bool IS_LT, int K>
static const void* skew_for_bw(int bw) {
  switch (bw) {
#define X(B) case B: return (const void*)eval1_skew_kernel<IS_LT, K, B>;
    FLE_BW16(X)
#undef X
  }
  return nullptr;
}

const void* eval1_skew_kernel_for(int type, int k, int bw) {
  if (type == PASS_LT) {
    if (k == 1) return skew_for_bw<true, 1>(bw);
    if (k == 2) return skew_for_bw<true, 2>(bw);
  } else if (type == PASS_EQ) {
    if (k == 1) return skew_for_bw<false, 1>(bw);
  }
  return nullptr;
}

const void* eval_in_mux_skew_kernel_for(int bw) {
  switch (bw) {
#define X(B) case B: return (const void*)eval_in_mux_skew_kernel<B>;
    X(1) X(2) X(3) X(4) X(5) X(6) X(7)
#undef X
  }
  return nullptr;
}

}  // namespace fle
