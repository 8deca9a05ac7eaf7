// Shim for the one Boost.StringAlgo call the reference's OBJ loader makes (mesh_io_obj.cpp): boost::replace_all(str, from, to) -
// every non-overlapping occurrence, left to right, in place.
#pragma once
#include <string>
namespace boost {
template <class S, class A, class B>
inline void replace_all(S& s, A const& from_, B const& to_) {
  const std::string from(from_), to(to_);
  if (from.empty()) return;
  for (size_t pos = 0; (pos = s.find(from, pos)) != std::string::npos; pos += to.size()) s.replace(pos, from.size(), to);
}
namespace algorithm { using boost::replace_all; }
}  // namespace boost
