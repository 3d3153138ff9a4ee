// Test infrastructure only: stand-in header so the UNMODIFIED reference sources compile
// without Boost/autoconf (see oracle/README.md). Not part of the product.
#pragma once
#include <tuple>
#include <vector>
#include <deque>
#include <utility>
namespace boost {
/* primary template: the nested typedefs of G, as in Boost (specialisations in the reference override it) */
template <class G> struct graph_traits {
  typedef typename G::vertex_descriptor vertex_descriptor; typedef typename G::edge_descriptor edge_descriptor;
  typedef typename G::adjacency_iterator adjacency_iterator; typedef typename G::out_edge_iterator out_edge_iterator;
  typedef typename G::in_edge_iterator in_edge_iterator; typedef typename G::vertex_iterator vertex_iterator;
  typedef typename G::edge_iterator edge_iterator; typedef typename G::directed_category directed_category;
  typedef typename G::edge_parallel_category edge_parallel_category; typedef typename G::traversal_category traversal_category;
  typedef typename G::vertices_size_type vertices_size_type; typedef typename G::edges_size_type edges_size_type;
  typedef typename G::degree_size_type degree_size_type;
  static vertex_descriptor null_vertex() { return G::null_vertex(); }
};
struct directed_tag {}; struct undirected_tag {}; struct bidirectional_tag : directed_tag {};
struct adjacency_graph_tag {}; struct incidence_graph_tag {};
struct bidirectional_graph_tag : incidence_graph_tag {};
struct vertex_list_graph_tag {}; struct edge_list_graph_tag {};
struct allow_parallel_edge_tag {}; struct disallow_parallel_edge_tag {};
namespace detail { inline bool is_directed(directed_tag) { return true; } inline bool is_directed(undirected_tag) { return false; } }
using std::tie;
namespace tuples { using std::ignore; }
}
/* edge = std::pair<V,V>: found by ADL only in std, so define globally + in std-less unqualified lookup */
template <class V, class G> inline V source(std::pair<V, V> e, const G&) { return e.first; }
template <class V, class G> inline V target(std::pair<V, V> e, const G&) { return e.second; }
