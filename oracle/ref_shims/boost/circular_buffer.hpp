// Shim for boost::circular_buffer used by eflib/memory/bounded_buffer.h.
#pragma once
#include <cstddef>
#include <deque>
namespace boost {
template <typename T>
class circular_buffer {
  std::deque<T> d_;
  std::size_t cap_;
public:
  typedef std::size_t size_type;
  typedef T value_type;
  typedef T const& param_type;
  explicit circular_buffer(size_type cap) : cap_(cap) {}
  void push_front(T const& v) {
    if (d_.size() == cap_) d_.pop_back();
    d_.push_front(v);
  }
  T& operator[](size_type i) { return d_[i]; }
  T const& operator[](size_type i) const { return d_[i]; }
  size_type capacity() const { return cap_; }
  size_type size() const { return d_.size(); }
};
}  // namespace boost
