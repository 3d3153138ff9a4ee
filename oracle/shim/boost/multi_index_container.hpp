// Test infrastructure only: stand-in header so the UNMODIFIED reference sources compile
// without Boost/autoconf (see oracle/README.md). Not part of the product.
// Just enough of multi_index_container for Common/InsOrderedMap.h: a vector in insertion order.
#pragma once
#include <vector>
#include <cstddef>
namespace boost { namespace multi_index {
template <class...> struct indexed_by {};
template <class...> struct random_access {};
template <class...> struct ordered_unique {};
template <class C, class T, T C::*P> struct member {};
template <class V, class I> struct multi_index_container {
  typedef std::vector<V> vec;
  vec v;
  struct index0 { const vec* p; typedef typename vec::const_iterator iterator;
    iterator begin() const { return p->begin(); } iterator end() const { return p->end(); } };
  template <int N> struct nth_index { typedef index0 type; };
  template <int N> index0 get() const { index0 i; i.p = &v; return i; }
  void push_back(const V& x) { v.push_back(x); }
  void erase(typename vec::const_iterator it) { v.erase(v.begin() + (it - v.cbegin())); }
  void clear() { v.clear(); } bool empty() const { return v.empty(); } std::size_t size() const { return v.size(); }
};
} }
