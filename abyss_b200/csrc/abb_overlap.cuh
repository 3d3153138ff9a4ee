// abb_overlap.cuh -- the stage that consumes the unitig FASTA: the contig overlap graph of AdjList
// (reference: AdjList/AdjList.cpp:140-291).  Per-item functions shared by the CUDA
// kernels (abb_overlap.cu) and by the single-thread test harness (tests/host_overlap), so that the
// join logic can be checked against the unmodified AdjList on a machine without a GPU.
//
// What AdjList computes.  Every contig i gives two vertices, node 2i ("i+") and 2i+1 ("i-", its
// reverse complement; ContigNode, Common/ContigNode.h).  With k1 = k-1:
//  (1) exact overlaps (buildOverlapGraph :247-268): edge x -> t whenever the last k1 bases of x equal
//      the first k1 bases of t, distance -(k-1).  The reference walks a hash table of suffixes; the
//      order of the out-edges of x that results is "ascending t^1" (see ovl_join_item).
//  (2) shorter overlaps (addOverlapsSA :140-201, only when min_overlap < k-1): among the vertices
//      left WITHOUT an out-edge by (1) ("blunt"), edge u -> v for every blunt u and every v whose
//      complement is blunt such that a suffix of the last k1 bases of u equals a prefix of the first
//      k1 bases of v, for the LONGEST such length q in [min_overlap, k-2]; distance -q.  The
//      reference sorts all suffixes (SuffixArray, Common/SuffixArray.h); out-edges of u end up in
//      ascending v^1 as well.
// With --SS an edge needs both vertices on the same strand (:159,262).
//
// Here both steps are hash joins on a 64-bit polynomial hash of the end strings followed by an exact
// comparison of the bases (the hash only selects candidates: the result does not depend on it), and
// one radix sort of the edges by (u, v^1) at the end.
#pragma once
#include "abb_device.cuh"

namespace abb {

struct OvlSeqs {
	const uint8_t* bases;  // concatenated contig characters (upper case)
	const uint64_t* offs;  // n + 1 offsets
	uint64_t n;            // contigs
	unsigned k1;           // k - 1
};

/** flattenAmbiguityCodes (Common/Sequence.h:50-72): an ambiguity code stands for its smallest base; N stays invalid */
ABB_HD unsigned ovl_code(unsigned char c)
{
	switch (c & 0xDFu) {
	case 'A': case 'M': case 'R': case 'W': case 'V': case 'H': case 'D': return 0;
	case 'C': case 'S': case 'Y': case 'B': return 1;
	case 'G': case 'K': return 2;
	case 'T': return 3;
	default: return 4;
	}
}

/** j-th base (2-bit code, 4 = not a base) of the first (end = 0) or last (end = 1) k1 bases of node t, read in the
 *  orientation of t */
ABB_HD unsigned ovl_base(const OvlSeqs& s, uint32_t t, int end, unsigned j)
{
	const uint64_t beg = s.offs[t >> 1];
	const uint64_t L = s.offs[(t >> 1) + 1] - beg;
	const uint64_t pos = end ? L - s.k1 + j : j; // in the string of the node
	if (t & 1) {
		const unsigned c = ovl_code(s.bases[beg + (L - 1 - pos)]);
		return c < 4 ? 3 - c : 4;
	}
	return ovl_code(s.bases[beg + pos]);
}

constexpr uint64_t kOvlMul = 0x9E3779B97F4A7C15ULL;

/** hash of bases [from, from + len) of an end string, mixed with the length */
ABB_HD uint64_t ovl_hash(const OvlSeqs& s, uint32_t t, int end, unsigned from, unsigned len, unsigned* bad)
{
	uint64_t h = 0;
	for (unsigned j = 0; j < len; ++j) {
		const unsigned c = ovl_base(s, t, end, from + j);
		if (c > 3 && bad)
			*bad = 1;
		h = h * kOvlMul + (c + 1);
	}
	h ^= (uint64_t)len * 0xD6E8FEB86659FD93ULL;
	h ^= h >> 32;
	return h * 0xD6E8FEB86659FD93ULL;
}

/** last `len` bases of u == first `len` bases of v ? */
ABB_HD bool ovl_match(const OvlSeqs& s, uint32_t u, uint32_t v, unsigned len)
{
	for (unsigned j = 0; j < len; ++j)
		if (ovl_base(s, u, 1, s.k1 - len + j) != ovl_base(s, v, 0, j))
			return false;
	return true;
}

/** first index in sorted[0, n) whose key is >= key */
ABB_HD uint64_t ovl_lower_bound(const uint64_t* sorted, uint64_t n, uint64_t key)
{
	uint64_t lo = 0, hi = n;
	while (lo < hi) {
		const uint64_t mid = lo + (hi - lo) / 2;
		if (sorted[mid] < key)
			lo = mid + 1;
		else
			hi = mid;
	}
	return lo;
}

ABB_HD uint64_t ovl_edge_key(uint32_t u, uint32_t v) { return ((uint64_t)u << 32) | (uint64_t)(v ^ 1u); }

/** step 0, item j of 2n: the table side of join (1) lists node t = j ^ 1 under the hash of its PREFIX (a stable sort by
 *  key then leaves equal keys in ascending t ^ 1, the reference's order); key_s[t] is the hash of the suffix of t.
 *  *bad is set when an end window holds something that is not a base (the reference's Kmer constructor aborts). */
ABB_HD void ovl_keys_item(const OvlSeqs& s, uint32_t j, uint64_t* key_p, uint32_t* val_p, uint64_t* key_s, unsigned* bad)
{
	const uint32_t t = j ^ 1u;
	key_p[j] = ovl_hash(s, t, 0, 0, s.k1, bad);
	val_p[j] = t;
	key_s[j] = ovl_hash(s, j, 1, 0, s.k1, bad);
}

/** join (1), source vertex x: its out-edges are the vertices t whose prefix equals the suffix of x.  Reference order:
 *  AdjList.cpp:252-264 appends, for v = x ^ 1, the edge (x, u ^ 1) for every u of suffixMap[prefix(v)] in insertion
 *  order = ascending u (readContigs :229-233); with t = u ^ 1 that is ascending t ^ 1.
 *  Returns the number of edges; writes them when ekey != nullptr. */
ABB_HD unsigned ovl_join_item(const OvlSeqs& s, int ss, uint32_t x, const uint64_t* key_s, const uint64_t* sorted_p,
                              const uint32_t* sorted_t, uint64_t n2, uint64_t* ekey, int* edist)
{
	unsigned c = 0;
	const uint64_t key = key_s[x];
	for (uint64_t i = ovl_lower_bound(sorted_p, n2, key); i < n2 && sorted_p[i] == key; ++i) {
		const uint32_t t = sorted_t[i];
		if (ss && ((t ^ x) & 1u))
			continue;
		if (!ovl_match(s, x, t, s.k1))
			continue;
		if (ekey) {
			ekey[c] = ovl_edge_key(x, t);
			edist[c] = -(int)s.k1;
		}
		++c;
	}
	return c;
}

/** join (2), table side: record (blunt vertex number b, length index qi) under the hash of the last q = q_max - qi bases
 *  of that vertex; q_max = k1 - 1, lengths down to min_overlap */
ABB_HD void ovl_sub_keys_item(const OvlSeqs& s, const uint32_t* blunt, uint64_t b, unsigned qi, unsigned n_q, uint64_t* key, uint64_t* val)
{
	const unsigned q = s.k1 - 1 - qi;
	key[b * n_q + qi] = ovl_hash(s, blunt[b], 1, s.k1 - q, q, nullptr);
	val[b * n_q + qi] = (b << 8) | qi;
}

/** join (2), query side: v = complement of the b-th blunt vertex.  addOverlapsSA(g, sa, v, vseq) (:140-164) shortens the
 *  query from k1 - 1 bases down to min_overlap and keeps, per source vertex, only the first (longest) hit. */
ABB_HD unsigned ovl_sub_join_item(const OvlSeqs& s, int ss, const uint32_t* blunt, uint64_t b, unsigned n_q, const uint64_t* sorted_key,
                                  const uint64_t* sorted_val, uint64_t n_rec, uint64_t* ekey, int* edist)
{
	const uint32_t v = blunt[b] ^ 1u;
	unsigned c = 0;
	for (unsigned qi = 0; qi < n_q; ++qi) {
		const unsigned q = s.k1 - 1 - qi;
		const uint64_t key = ovl_hash(s, v, 0, 0, q, nullptr);
		for (uint64_t i = ovl_lower_bound(sorted_key, n_rec, key); i < n_rec && sorted_key[i] == key; ++i) {
			if ((unsigned)(sorted_val[i] & 0xff) != qi)
				continue;
			const uint32_t u = blunt[sorted_val[i] >> 8];
			if (ss && ((u ^ v) & 1u))
				continue;
			if (!ovl_match(s, u, v, q))
				continue;
			bool longer = false; // `seen` (:145,160): u was already joined to v by a longer overlap
			for (unsigned q2 = q + 1; q2 < s.k1 && !longer; ++q2)
				longer = ovl_match(s, u, v, q2);
			if (longer)
				continue;
			if (ekey) {
				ekey[c] = ovl_edge_key(u, v);
				edist[c] = -(int)q;
			}
			++c;
		}
	}
	return c;
}

} // namespace abb
