// Run-time program specialisation (fle_jit.cu): plan, generated text, kernel cache.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "fle_params.h"

namespace fle {

// groups per thread per tile that kernels with a fused aggregate are generated (and launched) with
constexpr int kJitAggGpt = 2;

struct JitChain {
  bool lt;       // borrow chain ('<' family) or equality chain
  int bw;
  int pass, k;   // literal / init / negate come from P.passes[pass] chain k ...
  int in_idx;    // ... or (>= 0) the literal is in_lits[in_idx] of an unrolled IN list
  int mask_off;  // first of bw mask words in JitArgs::mask
};

struct JitPlan {
  std::string text;  // "fle_jit_program.h" -- also the cache key
  std::vector<JitChain> chains;
  int nmask = 0;
};

struct JitEntry {
  int rc = 0;                  // FLE_OK or why there is no kernel
  std::string log;
  void* lib = nullptr;         // cudaLibrary_t
  const void* kern = nullptr;  // cudaKernel_t, usable with cudaLaunchKernel / cudaFuncSetAttribute
};

// P: lowered program with ncols / bw[] filled in.  agg_col >= 0: also generate the fused
// SUM(plain * dict[code of column agg_col]) over the selected rows.  false: shape not supported
bool jit_plan(const EvalParams& P, const std::vector<uint32_t>& in_lits, JitPlan* plan, int agg_col = -1, bool skew = false);
// the JitArgs parameter block for the current literals
void jit_fill_args(const JitPlan& plan, const EvalParams& P, const std::vector<uint32_t>& in_lits,
                   std::vector<uint32_t>* args);
// NVRTC only (no device needed): FLE_OK and the cubin, or an error and the compiler log
int jit_compile_text(const std::string& program_text, std::vector<char>* cubin, std::string* log);
// process-wide kernel cache.  A shape that has no kernel yet is compiled (synchronously) once
// `force` is set or the rows evaluated with it so far, `rows` included, reach min_rows; until then
// the answer is null and the caller interprets.
// background: the compilation runs on its own thread (device: the CUDA device to load the kernel on), the answer
// stays null until it is done.
std::shared_ptr<JitEntry> jit_lookup(const std::string& program_text, int64_t rows, int64_t min_rows, bool force,
                                     bool background = false, int device = -1);

}  // namespace fle
