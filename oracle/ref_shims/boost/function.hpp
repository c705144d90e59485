// Build shim (test infrastructure): boost::function -> std::function
// (fle-encoding.h:222 `typedef boost::function<void()> UnpackFunction`).
#pragma once
#include <functional>
namespace boost {
template <class Sig>
using function = std::function<Sig>;
}  // namespace boost
