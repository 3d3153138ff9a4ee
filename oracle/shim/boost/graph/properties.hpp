// Test infrastructure only: stand-in header so the UNMODIFIED reference sources compile
// without Boost/autoconf (see oracle/README.md). Not part of the product.
#pragma once
#include <boost/graph/graph_traits.hpp>
namespace boost {
struct no_property {};
enum default_color_type { white_color, gray_color, green_color, red_color, black_color };
template <class C> struct color_traits {
  static default_color_type white() { return white_color; }
  static default_color_type gray() { return gray_color; }
  static default_color_type black() { return black_color; }
};
struct read_write_property_map_tag {};
template <class PM> struct property_traits;
enum vertex_bundle_t { vertex_bundle }; enum edge_bundle_t { edge_bundle };
enum vertex_name_t { vertex_name }; enum edge_name_t { edge_name };
enum vertex_index_t { vertex_index }; enum edge_weight_t { edge_weight };
/* the graph's own nested type when it has one (Boost does the same), no_property otherwise */
#define ABB_SHIM_NESTED(TRAIT, MEMBER) \
  template <class G> struct TRAIT { \
    template <class T> static typename T::MEMBER pick(int); template <class T> static no_property pick(...); \
    typedef decltype(pick<G>(0)) type; };
ABB_SHIM_NESTED(edge_bundle_type, edge_bundled)
ABB_SHIM_NESTED(vertex_bundle_type, vertex_bundled)
ABB_SHIM_NESTED(edge_property, edge_property_type)
ABB_SHIM_NESTED(vertex_property, vertex_property_type)
#undef ABB_SHIM_NESTED
template <class T> void function_requires() {}
}
#define BOOST_INSTALL_PROPERTY(KIND, NAME)
namespace boost {
template <class R, class PM> struct put_get_helper {};
template <class PM, class R, class K> inline R get(const put_get_helper<R, PM>& pm, const K& k) { return static_cast<const PM&>(pm)[k]; }
template <class G, class Tag> struct property_map;
struct readable_property_map_tag {};
}
