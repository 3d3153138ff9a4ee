// abb_graph.cuh -- out-edges of de Bruijn graph vertices for the GraphViz dump `-g` (outputGraph, BloomDBG/bloom-dbg.h:1171-1242).
// Shared by the CUDA kernel k_successors (abb_insert.cuh) and the CPU test harness (tests/host_graph).
#pragma once
#include "../../include/abyss_b200.h"
#include "abb_device.cuh"

namespace abb {

/** RollingBloomDBG::out_edge_iterator (BloomDBG/RollingBloomDBG.h:300-360): shift the k-mer left, try A, C, G, T as the last base,
 *  keep those the filter contains (`probe(canonical hash)`).  While a vertex has exactly one out-edge the walk moves on to that
 *  successor, up to max_chain (<= 128) vertices: unbranched paths -- nearly all of a genome's graph -- cost one query per
 *  max_chain vertices instead of one per vertex.  out[s] = out-edges of the s-th vertex of the chain (s = 0: the k-mer itself):
 *  mask bit b = the successor with base b exists, hash[b] = its canonical hash (vertex identity: RollingBloomDBGVertex compares
 *  canonical k-mers).  *self = canonical hash of the start k-mer.  Returns the number of vertices expanded. */
template <typename Probe>
ABB_HD unsigned successors_chain(const uint8_t* km, unsigned k, unsigned max_chain, const Probe& probe, abb_succ_info* out, uint64_t* self)
{
	const RollTab rt = make_rolltab(k);
	HashPair h = { 0, 0 }; // NTC64 from scratch (nthash.hpp:268-272)
	for (unsigned t = 0; t < k; ++t) {
		const unsigned c = base_code(km[t]) & 3u;
		h.fh = srol1(h.fh) ^ seed_of(c);
		h.rh ^= srol_n(seed_of(3 - c), t);
	}
	*self = h.canonical();
	unsigned appended[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }; // bases appended so far, 2 bits each
	unsigned s = 0;
	while (s < max_chain) {
		// the base that leaves on this shift: from the start k-mer while it lasts, then from the appended bases
		const unsigned first = s < k ? base_code(km[s]) & 3u : (appended[(s - k) >> 4] >> (2 * ((s - k) & 15))) & 3u;
		unsigned mask = 0;
		for (unsigned b = 0; b < 4; ++b) {
			const uint64_t h0 = roll_right(h, rt, first, b).canonical();
			out[s].hash[b] = h0;
			if (probe(h0))
				mask |= 1u << b;
		}
		out[s].mask = (uint8_t)mask;
		++s;
		if ((mask & (mask - 1)) != 0 || mask == 0 || s == max_chain)
			break;
		unsigned nb = 0;
		while (!((mask >> nb) & 1))
			++nb;
		h = roll_right(h, rt, first, nb);
		appended[(s - 1) >> 4] |= nb << (2 * ((s - 1) & 15));
	}
	return s;
}

} // namespace abb
