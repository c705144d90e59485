// Run-time specialisation of the fused-eval kernel per predicate-program SHAPE.
//
// Host side of fle_eval_jit.cuh: walks a lowered program (fle_params.h), writes the per-group
// evaluation function as straight-line CUDA C++ ("fle_jit_program.h"), compiles it together with
// the embedded kernel skeleton through NVRTC (libnvrtc is dlopen'ed: the library itself links
// against the CUDA runtime only), loads the cubin with cudaLibraryLoadData and caches the kernel
// process-wide under the generated text.  Literals are NOT part of the shape: they travel as a
// second kernel parameter (JitArgs: per chain init/negate words and one 0/~0 mask word per
// plane), which the generated LOP3s read as constant-bank operands.
//
// This is what Impala does with LLVM for the expression trees the reference interprets
// (simple-predicates.h:113-163); programs below `eval_jit_min_rows` keep using the interpreting
// kernels of fle_eval.cu, like Impala's disable_codegen_rows_threshold.
#include <dlfcn.h>
#include <nvrtc.h>

#include <map>
#include <memory>
#include <condition_variable>
#include <mutex>
#include <set>
#include <thread>

#include "fle_internal.h"
#include "fle_jit.h"

namespace fle {
namespace {

#include "_obj/fle_jit_embed.inc"  // kJitHeaderNames[], kJitHeaderTexts[], kJitNumHeaders

constexpr int kShortIn = 8;  // IN lists up to this long are unrolled into equality chains

void emit_fold(std::string& t, int slot, int fold, const std::string& v) {
  const char* op = fold == FOLD_SET ? "=" : (fold == FOLD_AND ? "&=" : "|=");
  t += "    s" + std::to_string(slot) + " " + op + " " + v + ";\n";
}

}  // namespace

bool jit_plan(const EvalParams& P, const std::vector<uint32_t>& in_lits, JitPlan* plan, int agg_col, bool skew) {
  JitPlan& J = *plan;
  J.text.clear();
  J.chains.clear();
  J.nmask = 0;
  std::string body;
  char b[256];
  // Flat AND / OR programs fold everything into slot 0, which starts as the operator's identity
  // (slot0_init): their one SET is emitted as that operator, so the order of the folds is free.
  int flat_op = FOLD_SET;
  if (P.ncombs == 0 && P.nslots <= 1)
    for (int r = 0; r < P.npasses && flat_op == FOLD_SET; ++r)
      for (int k = 0; k < 4 && flat_op == FOLD_SET; ++k)
        if (k < (int)std::max<uint32_t>(1, P.passes[r].k_real) && P.passes[r].fold[k] != FOLD_SET)
          flat_op = P.passes[r].fold[k];
  auto fold_of = [&](int f) { return f == FOLD_SET ? flat_op : f; };
  int p = 0;
  while (p < P.npasses) {
    const PassDesc& pd0 = P.passes[p];
    if (pd0.type == PASS_CONST) {
      emit_fold(body, pd0.slot[0], fold_of(pd0.fold[0]), pd0.init[0] ? "~0ull" : "0ull");
      ++p;
      continue;
    }
    int q = p;
    while (q < P.npasses && P.passes[q].type != PASS_CONST && P.passes[q].col == pd0.col) ++q;
    const int col = pd0.col, bw = (int)P.bw[col];
    snprintf(b, sizeof(b), "  {  // column %d, %d bits\n", col, bw);
    body += b;
    // chains swept plane by plane
    struct Sweep { int id; bool lt; };
    std::vector<Sweep> sweep;
    std::string folds;
    for (int r = p; r < q; ++r) {
      const PassDesc& pd = P.passes[r];
      if (pd.type == PASS_LT || pd.type == PASS_EQ) {
        for (int k = 0; k < (int)pd.k_real; ++k) {
          JitChain ch{};
          ch.lt = pd.type == PASS_LT;
          ch.bw = bw;
          ch.pass = r;
          ch.k = k;
          ch.in_idx = -1;
          ch.mask_off = J.nmask;
          J.nmask += bw;
          const int id = (int)J.chains.size();
          J.chains.push_back(ch);
          sweep.push_back({id, ch.lt});
          snprintf(b, sizeof(b), "jit_res(c%dl, c%dh, A.neg[%d])", id, id, id);
          emit_fold(folds, pd.slot[k], fold_of(pd.fold[k]), b);
        }
      } else if (pd.type == PASS_IN && (int)pd.lit[1] <= kShortIn) {
        std::string v;
        for (int j = 0; j < (int)pd.lit[1]; ++j) {
          JitChain ch{};
          ch.lt = false;
          ch.bw = bw;
          ch.pass = r;
          ch.k = 0;
          ch.in_idx = (int)pd.lit[0] + j;
          if (ch.in_idx >= (int)in_lits.size()) return false;
          ch.mask_off = J.nmask;
          J.nmask += bw;
          const int id = (int)J.chains.size();
          J.chains.push_back(ch);
          sweep.push_back({id, false});
          snprintf(b, sizeof(b), "%sjit_res(c%dl, c%dh, 0u)", j ? " | " : "", id, id);
          v += b;
        }
        emit_fold(folds, pd.slot[0], fold_of(pd.fold[0]), std::string(pd.neg[0] ? "~(" : "(") + v + ")");
      }
    }
    if (!sweep.empty()) {
      for (const Sweep& s : sweep) {
        snprintf(b, sizeof(b), "    uint32_t c%dl = A.ini[%d], c%dh = A.ini[%d];\n", s.id, s.id, s.id, s.id);
        body += b;
      }
      body += "    uint64_t w0, w1;\n";
      if (skew) {  // a skewed column: all plane words at once (whole-chunk loads + funnel shift), then the sweep
        snprintf(b, sizeof(b), "    uint64_t wq[%d];\n    ld.template load_all<%d, %d>(wq);\n", bw, col, bw);
        body += b;
      }
      for (int i = 0; i < bw; i += 2) {
        if (skew) {
          if (i + 1 < bw) snprintf(b, sizeof(b), "    w0 = wq[%d]; w1 = wq[%d];\n", i, i + 1);
          else snprintf(b, sizeof(b), "    w0 = wq[%d]; w1 = 0;\n", i);
        } else {
          snprintf(b, sizeof(b), "    ld.template load_pair<%d, %d>(%d, w0, w1);\n", col, bw, i);
        }
        body += b;
        for (int h = 0; h < 2 && i + h < bw; ++h) {
          for (const Sweep& s : sweep) {
            snprintf(b, sizeof(b), "    %s(c%dl, c%dh, w%d, A.mask[%d]);\n", s.lt ? "jit_lt" : "jit_eq", s.id, s.id, h,
                     J.chains[s.id].mask_off + i + h);
            body += b;
          }
        }
      }
      body += folds;
    }
    // IN lists of run-time length: planes in registers, literals from shared memory
    for (int r = p; r < q; ++r) {
      const PassDesc& pd = P.passes[r];
      if (pd.type != PASS_IN || (int)pd.lit[1] <= kShortIn) continue;
      snprintf(b, sizeof(b), "    {\n      uint64_t w[%d];\n", bw);
      body += b;
      for (int i = 0; i < bw; i += 2) {
        if (i + 1 < bw)
          snprintf(b, sizeof(b), "      ld.template load_pair<%d, %d>(%d, w[%d], w[%d]);\n", col, bw, i, i, i + 1);
        else
          snprintf(b, sizeof(b), "      { uint64_t t_; ld.template load_pair<%d, %d>(%d, w[%d], t_); }\n", col, bw, i, i);
        body += b;
      }
      snprintf(b, sizeof(b),
               "      const uint64_t v = jit_in_list<%d>(w, lits + P.passes[%d].lit[0], (int)P.passes[%d].lit[1]);\n", bw,
               r, r);
      body += b;
      std::string f;
      emit_fold(f, pd.slot[0], fold_of(pd.fold[0]), pd.neg[0] ? "~v" : "v");
      body += "  " + f + "    }\n";
    }
    body += "  }\n";
    p = q;
  }
  if (J.chains.size() > 96 || J.nmask > 2048) return false;  // keep the parameter block < 10 KB

  std::string& t = J.text;
  t = "// generated by fle_jit.cu -- one predicate-program shape\n";
  // (the marker makes jit_compile_text define FLE_JIT_SKEW ahead of the kernel skeleton, whose loaders depend on it)
  if (skew) t += "// FLE_JIT_SKEW: columns at arbitrary byte addresses (EvalParams::plane_skew)\n";
  t += "namespace fle {\n";
  snprintf(b, sizeof(b), "constexpr int kJitNCols = %d;\n", P.ncols);
  t += b;
  snprintf(b, sizeof(b), "struct JitArgs {\n  uint32_t ini[%d];\n  uint32_t neg[%d];\n  uint32_t mask[%d];\n};\n",
           std::max<int>(1, (int)J.chains.size()), std::max<int>(1, (int)J.chains.size()), std::max(1, J.nmask));
  t += b;
  t += "__device__ __forceinline__ void jit_stage_addr(const EvalParams& P, uint32_t base, uint32_t seg, uint32_t jrow,\n"
       "                                               uint32_t (&a)[kJitNCols]) {\n";
  for (int c = 0; c < P.ncols; ++c) {
    if (skew)
      snprintf(b, sizeof(b), "  a[%d] = base + P.col_off[%d] + seg * P.seg_stride[%d] + jrow * %uu + P.plane_skew[%d];\n", c, c, c,
               P.bw[c] * 8u, c);
    else
      snprintf(b, sizeof(b), "  a[%d] = base + P.col_off[%d] + seg * P.seg_stride[%d] + jrow * %uu;\n", c, c, c, P.bw[c] * 8u);
    t += b;
  }
  t += "}\n__device__ __forceinline__ void jit_advance(uint32_t (&a)[kJitNCols]) {\n";
  for (int c = 0; c < P.ncols; ++c) {
    snprintf(b, sizeof(b), "  a[%d] += %uu;\n", c, 32u * P.bw[c] * 8u);
    t += b;
  }
  t += "}\n__device__ __forceinline__ void jit_global_addr(const EvalParams& P, int64_t g, const uint64_t* (&p)[kJitNCols]) {\n";
  for (int c = 0; c < P.ncols; ++c) {
    if (skew)
      snprintf(b, sizeof(b), "  p[%d] = reinterpret_cast<const uint64_t*>(reinterpret_cast<const uint8_t*>(P.planes[%d]) + P.plane_skew[%d] + g * %u);\n",
               c, c, c, P.bw[c] * 8u);
    else
      snprintf(b, sizeof(b), "  p[%d] = P.planes[%d] + g * %u;\n", c, c, P.bw[c]);
    t += b;
  }
  t += "}\ntemplate <class L>\n"
       "__device__ __forceinline__ uint64_t jit_eval_group(const EvalParams& P, const JitArgs& A, const L& ld,\n"
       "                                                   const uint32_t* lits) {\n";
  t += "  uint64_t s0 = P.slot0_init;\n";
  for (int s = 1; s < std::max(1, P.nslots); ++s) {
    snprintf(b, sizeof(b), "  uint64_t s%d = 0;\n", s);
    t += b;
  }
  t += body;
  for (int c = 0; c < P.ncombs; ++c) {
    const CombDesc& cd = P.combs[c];
    snprintf(b, sizeof(b), "  s%d = s%d %s s%d;\n", cd.dst, cd.a, cd.op == FOLD_AND ? "&" : "|", cd.b);
    t += b;
  }
  t += "  return s0;\n}\n";
  {
    // two groups of a tile interleaved when a group is short (narrow columns: the per-group
    // dependency chains, not bandwidth, bound such programs); longer groups are better left alone
    int sum_bw = 0;
    for (int c = 0; c < P.ncols; ++c) sum_bw += (int)P.bw[c];
    snprintf(b, sizeof(b), "#define FLE_JIT_UNROLL %d\n", sum_bw <= 16 ? 2 : 1);
    t += b;
  }
  if (agg_col >= 0) {
    if (agg_col >= P.ncols || skew) return false;
    const int bw = (int)P.bw[agg_col];
    snprintf(b, sizeof(b),
             "#define FLE_JIT_AGG 1\n#define FLE_JIT_GPT %d\nconstexpr int kJitAggCol = %d;\nconstexpr int kJitAggBw = %d;\n",
             kJitAggGpt, agg_col, bw);
    t += b;
    // code of the row at plane bit `sh` of one group of the aggregate's column, planes in shared / global memory
    t += "__device__ __forceinline__ uint32_t jit_agg_code_smem(uint32_t a, int sh) {\n  uint32_t code = 0;\n";
    if (bw % 2 == 0) {
      t += "  uint64_t w0, w1;\n";
      for (int i = 0; i < bw; i += 2) {
        snprintf(b, sizeof(b),
                 "  lds128(a + %d, w0, w1);\n  code |= (uint32_t)((w0 >> sh) & 1ull) << %d;\n"
                 "  code |= (uint32_t)((w1 >> sh) & 1ull) << %d;\n", i * 8, i, i + 1);
        t += b;
      }
    } else {
      for (int i = 0; i < bw; ++i) {
        snprintf(b, sizeof(b), "  code |= (uint32_t)((lds64(a + %d) >> sh) & 1ull) << %d;\n", i * 8, i);
        t += b;
      }
    }
    t += "  return code;\n}\n"
         "__device__ __forceinline__ uint32_t jit_agg_code_global(const uint64_t* p, int sh) {\n  uint32_t code = 0;\n";
    for (int i = 0; i < bw; ++i) {
      snprintf(b, sizeof(b), "  code |= (uint32_t)((ldg_gather64(p + %d) >> sh) & 1ull) << %d;\n", i, i);
      t += b;
    }
    t += "  return code;\n}\n";
  }
  t += "}  // namespace fle\n";
  return true;
}

void jit_fill_args(const JitPlan& J, const EvalParams& P, const std::vector<uint32_t>& in_lits,
                   std::vector<uint32_t>* args) {
  const int n = std::max<int>(1, (int)J.chains.size());
  args->assign((size_t)2 * n + std::max(1, J.nmask), 0u);
  uint32_t* ini = args->data();
  uint32_t* neg = ini + n;
  uint32_t* mask = neg + n;
  for (size_t c = 0; c < J.chains.size(); ++c) {
    const JitChain& ch = J.chains[c];
    const PassDesc& pd = P.passes[ch.pass];
    uint32_t lit;
    if (ch.in_idx >= 0) {
      lit = in_lits[ch.in_idx];
      ini[c] = ~0u;
      neg[c] = 0u;
    } else {
      lit = pd.lit[ch.k];
      ini[c] = ch.lt ? (pd.init[ch.k] ? ~0u : 0u) : ~0u;
      neg[c] = pd.neg[ch.k] ? ~0u : 0u;
    }
    for (int i = 0; i < ch.bw; ++i) mask[ch.mask_off + i] = 0u - ((lit >> i) & 1u);
  }
}

// ------------------------------------------------------------------------------------------
// NVRTC (dlopen) + kernel cache
// ------------------------------------------------------------------------------------------
namespace {

struct Nvrtc {
  void* handle = nullptr;
  std::string why;
  decltype(&nvrtcCreateProgram) CreateProgram = nullptr;
  decltype(&nvrtcCompileProgram) CompileProgram = nullptr;
  decltype(&nvrtcDestroyProgram) DestroyProgram = nullptr;
  decltype(&nvrtcGetCUBINSize) GetCUBINSize = nullptr;
  decltype(&nvrtcGetCUBIN) GetCUBIN = nullptr;
  decltype(&nvrtcGetProgramLogSize) GetProgramLogSize = nullptr;
  decltype(&nvrtcGetProgramLog) GetProgramLog = nullptr;
  decltype(&nvrtcGetErrorString) GetErrorString = nullptr;
};

Nvrtc* nvrtc() {
  static Nvrtc n;
  static std::once_flag once;
  std::call_once(once, [] {
    std::vector<std::string> names;
    if (const char* e = getenv("FLE_NVRTC_LIB")) names.push_back(e);
    for (const char* s : {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so",
                          "/usr/local/cuda/lib64/libnvrtc.so"})
      names.push_back(s);
    for (const std::string& s : names) {
      n.handle = dlopen(s.c_str(), RTLD_NOW | RTLD_LOCAL);
      if (n.handle) break;
    }
    if (!n.handle) {
      n.why = "libnvrtc.so.12 not found (set FLE_NVRTC_LIB)";
      return;
    }
#define FLE_SYM(name)                                                             \
  n.name = reinterpret_cast<decltype(n.name)>(dlsym(n.handle, "nvrtc" #name)); \
  if (!n.name) n.why = "libnvrtc lacks nvrtc" #name;
    FLE_SYM(CreateProgram) FLE_SYM(CompileProgram) FLE_SYM(DestroyProgram) FLE_SYM(GetCUBINSize) FLE_SYM(GetCUBIN)
    FLE_SYM(GetProgramLogSize) FLE_SYM(GetProgramLog) FLE_SYM(GetErrorString)
#undef FLE_SYM
  });
  return &n;
}

std::mutex g_mu;
std::condition_variable g_compiled_cv;
std::map<std::string, std::shared_ptr<JitEntry> > g_cache;
std::set<std::string> g_compiling;  // shapes some thread is compiling right now
std::map<std::string, int64_t> g_pending_rows;
constexpr size_t kJitCacheMax = 1024;  // compiled shapes kept per process
struct Compilers {  // background compilations (eval_jit = 3), joined before the cache is destroyed
  std::vector<std::thread> threads;
  ~Compilers() {
    for (auto& t : threads)
      if (t.joinable()) t.join();
  }
};
Compilers g_compilers;  // (declared after the maps: destroyed before them)

}  // namespace

int jit_compile_text(const std::string& program_text, std::vector<char>* cubin, std::string* log) {
  Nvrtc* n = nvrtc();
  if (!n->handle || !n->why.empty()) {
    if (log) *log = n->why;
    return FLE_E_UNSUPPORTED;
  }
  std::vector<const char*> names, texts;
  for (int i = 0; i < kJitNumHeaders; ++i) {
    names.push_back(kJitHeaderNames[i]);
    texts.push_back(kJitHeaderTexts[i]);
  }
  names.push_back("fle_jit_program.h");
  texts.push_back(program_text.c_str());
  nvrtcProgram prog = nullptr;
  const bool skew = program_text.find("// FLE_JIT_SKEW") != std::string::npos;
  nvrtcResult r = n->CreateProgram(&prog, skew ? "#define FLE_JIT_SKEW 1\n#include \"fle_eval_jit.cuh\"\n" : "#include \"fle_eval_jit.cuh\"\n",
                                   "fle_jit_main.cu", (int)names.size(), texts.data(), names.data());
  if (r != NVRTC_SUCCESS) {
    if (log) *log = std::string("nvrtcCreateProgram: ") + n->GetErrorString(r);
    return FLE_E_CUDA;
  }
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
  r = n->CompileProgram(prog, 3, opts);
  if (r != NVRTC_SUCCESS) {
    size_t ls = 0;
    n->GetProgramLogSize(prog, &ls);
    std::string l(ls, '\0');
    if (ls) n->GetProgramLog(prog, &l[0]);
    if (log) *log = std::string("nvrtcCompileProgram: ") + n->GetErrorString(r) + "\n" + l;
    n->DestroyProgram(&prog);
    return FLE_E_CUDA;
  }
  size_t cs = 0;
  n->GetCUBINSize(prog, &cs);
  cubin->resize(cs);
  r = n->GetCUBIN(prog, cubin->data());
  n->DestroyProgram(&prog);
  if (r != NVRTC_SUCCESS || cs == 0) {
    if (log) *log = std::string("nvrtcGetCUBIN: ") + n->GetErrorString(r);
    return FLE_E_CUDA;
  }
  return FLE_OK;
}

// compile + load one shape (no lock held) and publish the entry
static std::shared_ptr<JitEntry> compile_and_publish(const std::string& program_text, int device) {
  std::shared_ptr<JitEntry> e = std::make_shared<JitEntry>();
  std::vector<char> cubin;
  int rc = jit_compile_text(program_text, &cubin, &e->log);
  if (rc == FLE_OK) {
    cudaLibrary_t lib = nullptr;
    cudaError_t ce = device >= 0 ? cudaSetDevice(device) : cudaSuccess;  // (a background thread starts on device 0)
    if (ce == cudaSuccess) ce = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (ce == cudaSuccess) {
      cudaKernel_t k = nullptr;
      ce = cudaLibraryGetKernel(&k, lib, "fle_jit_kernel");
      if (ce == cudaSuccess) {
        e->lib = lib;
        e->kern = reinterpret_cast<const void*>(k);
      }
    }
    if (ce != cudaSuccess) {
      cudaGetLastError();
      e->log = std::string("loading the specialised kernel: ") + cudaGetErrorString(ce);
      rc = FLE_E_CUDA;
    }
  }
  e->rc = rc;
  std::lock_guard<std::mutex> lock(g_mu);
  g_cache[program_text] = e;
  g_compiling.erase(program_text);
  g_compiled_cv.notify_all();
  return e;
}

std::shared_ptr<JitEntry> jit_lookup(const std::string& program_text, int64_t rows, int64_t min_rows, bool force,
                                     bool background, int device) {
  // The lock covers the maps only, never a compilation: while one thread compiles a shape (0.3-0.8 s of NVRTC),
  // cache hits of every other scanner thread go through, and callers that meet a shape IN compilation interpret
  // meanwhile (force: they wait for that one entry).
  std::unique_lock<std::mutex> lock(g_mu);
  for (;;) {
    auto it = g_cache.find(program_text);
    if (it != g_cache.end()) return it->second;
    if (g_compiling.find(program_text) == g_compiling.end()) break;
    if (!force) return nullptr;
    g_compiled_cv.wait(lock);
  }
  // not compiled yet: the shape pays for its compilation once the rows evaluated with it (this
  // call included, all contexts of the process together) reach min_rows
  if (g_pending_rows.size() > 1024) g_pending_rows.clear();  // bounded bookkeeping for one-off shapes
  int64_t& seen = g_pending_rows[program_text];
  seen += rows;
  if (!force && seen < min_rows) return nullptr;
  if (g_cache.size() >= kJitCacheMax) return nullptr;  // bounded cache: further shapes are interpreted
  g_pending_rows.erase(program_text);
  g_compiling.insert(program_text);
  if (background && !force) {
    // eval_jit = 3: the compilation runs on its own thread and the caller interprets until the kernel is there
    // (a scanner thread never stalls for NVRTC); the threads are joined when the library is unloaded
    g_compilers.threads.emplace_back(compile_and_publish, program_text, device);
    return nullptr;
  }
  lock.unlock();
  return compile_and_publish(program_text, device);
}

}  // namespace fle
