// Shim for boost::shared_array used by salvia/resource/resource_data.h.
#pragma once
#include <cstddef>
#include <memory>
namespace boost {
template <typename T>
class shared_array {
  std::shared_ptr<T[]> p_;
public:
  shared_array() = default;
  explicit shared_array(T* p) : p_(p) {}
  void reset(T* p = nullptr) { p_.reset(p); }
  T* get() const { return p_.get(); }
  T& operator[](std::ptrdiff_t i) const { return p_[i]; }
  explicit operator bool() const { return static_cast<bool>(p_); }
};
}  // namespace boost
