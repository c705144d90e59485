// Build shim (test infrastructure): lets the reference's fle-encoding.h, which is
// written against boost, compile in an image that has no boost.  Only the two names
// the header uses (fle-encoding.h:51 `bind(mem_fn(&FleDecoder::Unpack_1), this)`).
#pragma once
#include <functional>
namespace boost {
using std::bind;
using std::mem_fn;
}  // namespace boost
