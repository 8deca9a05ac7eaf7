// Shim: range-v3 names used by eflib map 1:1 onto std::ranges.
#pragma once
#include <algorithm>
#include <ranges>
namespace ranges {
using std::ranges::for_each;
using std::ranges::subrange;
}  // namespace ranges
