// Test infrastructure only: stand-in header so the UNMODIFIED reference sources compile
// without Boost/autoconf (see oracle/README.md). Not part of the product.
#pragma once
#include <boost/graph/properties.hpp>
#include <deque>
namespace boost {
template <class T> class queue {
  std::deque<T> q_;
 public:
  void push(const T& t) { q_.push_back(t); }
  void pop() { q_.pop_front(); }
  T& top() { return q_.front(); }
  bool empty() const { return q_.empty(); }
};
}
