// Test infrastructure only: stand-in header so the UNMODIFIED reference sources compile
// without Boost/autoconf (see oracle/README.md). Not part of the product.
#pragma once
#include <boost/graph/graph_traits.hpp>
#include <vector>
#include <cstddef>
namespace boost {
struct vecS {}; struct bidirectionalS {};
struct al_edge { std::size_t u, v; al_edge(std::size_t a=0, std::size_t b=0):u(a),v(b){}
  bool operator==(const al_edge&o)const{return u==o.u&&v==o.v;} bool operator!=(const al_edge&o)const{return !(*this==o);} };
template <class A, class B, class D> struct adjacency_list {
  std::vector<std::vector<al_edge> > out, in;
  void grow(std::size_t n){ if(out.size()<n){out.resize(n);in.resize(n);} }
};
template <class A,class B,class D> struct graph_traits<adjacency_list<A,B,D> > {
  typedef std::size_t vertex_descriptor; typedef al_edge edge_descriptor;
  typedef std::vector<al_edge>::const_iterator out_edge_iterator; typedef out_edge_iterator in_edge_iterator;
  typedef unsigned degree_size_type;
};
template <class A,class B,class D> std::pair<al_edge,bool> add_edge(std::size_t u,std::size_t v,adjacency_list<A,B,D>&g){
  g.grow((u>v?u:v)+1); g.out[u].push_back(al_edge(u,v)); g.in[v].push_back(al_edge(u,v)); return std::make_pair(al_edge(u,v),true);}
static const std::vector<al_edge> al_empty;
template <class A,class B,class D> std::pair<std::vector<al_edge>::const_iterator,std::vector<al_edge>::const_iterator>
out_edges(std::size_t u,const adjacency_list<A,B,D>&g){ const std::vector<al_edge>&v=u<g.out.size()?g.out[u]:al_empty; return std::make_pair(v.begin(),v.end());}
template <class A,class B,class D> std::pair<std::vector<al_edge>::const_iterator,std::vector<al_edge>::const_iterator>
in_edges(std::size_t u,const adjacency_list<A,B,D>&g){ const std::vector<al_edge>&v=u<g.in.size()?g.in[u]:al_empty; return std::make_pair(v.begin(),v.end());}
template <class A,class B,class D> std::pair<al_edge,bool> edge(std::size_t u,std::size_t v,const adjacency_list<A,B,D>&g){
  bool f=false; if(u<g.out.size()) for(std::size_t i=0;i<g.out[u].size();++i) if(g.out[u][i].v==v) f=true; return std::make_pair(al_edge(u,v),f);}
template <class G> std::size_t source(const al_edge&e,const G&){return e.u;}
template <class G> std::size_t target(const al_edge&e,const G&){return e.v;}
}
