// Test infrastructure only: stand-in header so the UNMODIFIED reference sources compile
// without Boost/autoconf (see oracle/README.md). Not part of the product.
#pragma once
#include <tuple>
#include <vector>
#include <deque>
#include <utility>
namespace boost {
template <class G> struct graph_traits;
struct directed_tag {}; struct undirected_tag {}; struct bidirectional_tag : directed_tag {};
struct adjacency_graph_tag {}; struct incidence_graph_tag {};
struct bidirectional_graph_tag : incidence_graph_tag {};
struct vertex_list_graph_tag {}; struct edge_list_graph_tag {};
struct allow_parallel_edge_tag {}; struct disallow_parallel_edge_tag {};
using std::tie;
namespace tuples { using std::ignore; }
}
/* edge = std::pair<V,V>: found by ADL only in std, so define globally + in std-less unqualified lookup */
template <class V, class G> inline V source(std::pair<V, V> e, const G&) { return e.first; }
template <class V, class G> inline V target(std::pair<V, V> e, const G&) { return e.second; }
