// tests/host_overlap/host_overlap.cpp -- TEST INFRASTRUCTURE ONLY (never linked into libabyssb200).
//
// Single-thread emulation of abb_overlap_build: the SAME per-item functions the CUDA kernels call
// (abyss_b200/csrc/abb_overlap.cuh) driven by plain loops, std::stable_sort in place of the device radix sorts,
// behind the product's AdjList command line and writers (abyss_b200/host/adjlist_main.h).  It lets the join logic
// and the output formats be compared byte for byte with the unmodified AdjList on a machine without a GPU.
#include "../../abyss_b200/host/adjlist_main.h"
#include "../../abyss_b200/csrc/abb_overlap.cuh"
#include <numeric>

using namespace abb;

static std::vector<abb_overlap_edge> g_edges;

template <typename K, typename V>
static void sort_pairs(std::vector<K>& key, std::vector<V>& val)
{
	std::vector<size_t> idx(key.size());
	std::iota(idx.begin(), idx.end(), 0);
	std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return key[a] < key[b]; });
	std::vector<K> k2(key.size());
	std::vector<V> v2(val.size());
	for (size_t i = 0; i < idx.size(); ++i) {
		k2[i] = key[idx[i]];
		v2[i] = val[idx[i]];
	}
	key.swap(k2);
	val.swap(v2);
}

static void build(const char* bases, const uint64_t* offsets, uint64_t n, unsigned k, unsigned min_overlap, int ss, int,
                  const abb_overlap_edge** edges, uint64_t* n_edges)
{
	g_edges.clear();
	*edges = nullptr;
	*n_edges = 0;
	if (n == 0)
		return;
	const uint64_t n2 = 2 * n;
	OvlSeqs s = { (const uint8_t*)bases, offsets, n, k - 1 };
	unsigned bad = 0;
	std::vector<uint64_t> key_p(n2), key_s(n2);
	std::vector<uint32_t> val_p(n2);
	for (uint64_t j = 0; j < n2; ++j)
		ovl_keys_item(s, (uint32_t)j, key_p.data(), val_p.data(), key_s.data(), &bad);
	if (bad) {
		std::cerr << "AdjList: a contig end holds a character that is not a nucleotide\n";
		exit(EXIT_FAILURE);
	}
	sort_pairs(key_p, val_p);
	std::vector<uint64_t> ekey;
	std::vector<int> edist;
	std::vector<uint64_t> off(n2 + 1, 0);
	for (uint64_t x = 0; x < n2; ++x)
		off[x + 1] = off[x] + ovl_join_item(s, ss, (uint32_t)x, key_s.data(), key_p.data(), val_p.data(), n2, nullptr, nullptr);
	ekey.resize(off[n2]);
	edist.resize(off[n2]);
	for (uint64_t x = 0; x < n2; ++x)
		ovl_join_item(s, ss, (uint32_t)x, key_s.data(), key_p.data(), val_p.data(), n2, ekey.data() + off[x], edist.data() + off[x]);
	const unsigned n_q = min_overlap < k - 1 ? k - 1 - min_overlap : 0;
	if (n_q) {
		std::vector<uint32_t> blunt;
		for (uint64_t x = 0; x < n2; ++x)
			if (off[x + 1] == off[x])
				blunt.push_back((uint32_t)x);
		const uint64_t nb = blunt.size(), n_rec = nb * n_q;
		std::vector<uint64_t> sk(n_rec), sv(n_rec);
		for (uint64_t i = 0; i < n_rec; ++i)
			ovl_sub_keys_item(s, blunt.data(), i / n_q, (unsigned)(i % n_q), n_q, sk.data(), sv.data());
		sort_pairs(sk, sv);
		std::vector<uint64_t> soff(nb + 1, 0);
		for (uint64_t b = 0; b < nb; ++b)
			soff[b + 1] = soff[b] + ovl_sub_join_item(s, ss, blunt.data(), b, n_q, sk.data(), sv.data(), n_rec, nullptr, nullptr);
		const uint64_t e1 = ekey.size();
		ekey.resize(e1 + soff[nb]);
		edist.resize(e1 + soff[nb]);
		for (uint64_t b = 0; b < nb; ++b)
			ovl_sub_join_item(s, ss, blunt.data(), b, n_q, sk.data(), sv.data(), n_rec, ekey.data() + e1 + soff[b], edist.data() + e1 + soff[b]);
	}
	sort_pairs(ekey, edist);
	g_edges.resize(ekey.size());
	for (size_t i = 0; i < ekey.size(); ++i) {
		g_edges[i].u = (uint32_t)(ekey[i] >> 32);
		g_edges[i].v = (uint32_t)(ekey[i] & 0xffffffffu) ^ 1u;
		g_edges[i].distance = edist[i];
	}
	*edges = g_edges.data();
	*n_edges = g_edges.size();
}

int main(int argc, char** argv) { return adjlist::run(argc, argv, build); }
