// Test infrastructure only: stand-in header so the UNMODIFIED reference sources compile
// without Boost/autoconf (see oracle/README.md). Not part of the product.
#pragma once
#include <memory>
namespace boost {
template <class T> class shared_array {
  std::shared_ptr<T> p_;
 public:
  shared_array() {}
  explicit shared_array(T* p) : p_(p, std::default_delete<T[]>()) {}
  T* get() const { return p_.get(); }
};
}
