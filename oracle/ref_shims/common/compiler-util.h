// Build shim (test infrastructure) for Impala's common/compiler-util.h + glog DCHECKs.
#pragma once
#include <cassert>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#define LIKELY(expr) __builtin_expect(!!(expr), 1)
#define UNLIKELY(expr) __builtin_expect(!!(expr), 0)
#define DCHECK(x) assert(x)
#define DCHECK_GE(a, b) assert((a) >= (b))
#define DCHECK_LE(a, b) assert((a) <= (b))
#define DCHECK_GT(a, b) assert((a) > (b))
#define DCHECK_LT(a, b) assert((a) < (b))
#define DCHECK_EQ(a, b) assert((a) == (b))
#define DCHECK_NE(a, b) assert((a) != (b))
