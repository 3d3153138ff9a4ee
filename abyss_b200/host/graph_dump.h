// graph_dump.h -- the GraphViz dump of the Bloom filter de Bruijn graph (`abyss-bloom-dbg -g FILE`): the reference's traversal
// order on the host, the Bloom lookups behind two functors.  abyss_bloom_dbg.cc passes the C ABI (abb_contains_reads,
// abb_successors: CUDA), the CPU test harness (tests/host_graph) an emulation built on the same device functions.
#pragma once
#include "../../include/abyss_b200.h"
#include "reads.h"
#include <cstring>
#include <iostream>
#include <unordered_map>
#include <unordered_set>

namespace host {

/** outputGraph (bloom-dbg.h:1171-1242): every read is trimmed to its longest run of solid k-mers (trimSeq, :399-447) and a
 *  breadth-first search over the out-edges starts at its first k-mer and at the first k-mer of its reverse complement; one
 *  colour map for the whole run, vertices equal up to reverse complement, vertex and edge lines written as they are discovered
 *  (GraphvizBFSVisitor, :1096-1160).  The order is inherently sequential; what is batched on the GPU are the Bloom lookups: the
 *  solid flags of all k-mers of a batch of reads (abb_contains_reads) and the out-edges of a whole BFS level, speculatively
 *  continued along unbranched paths (abb_successors), cached until the search pops the vertex. */
template <typename BatchesFn, typename ContainsFn, typename SuccFn>
void output_graph(unsigned k, int verbose, BatchesFn for_each_batch, ContainsFn contains_reads, SuccFn successors, std::ostream& out)
{
	const unsigned kChain = 64;
	static const char BASES[] = "ACGT";
	uint64_t nodes = 0, edges = 0, readsProcessed = 0;
	std::unordered_set<uint64_t> visited; // the colour map: white = absent (every search runs to completion)
	struct Info {
		uint64_t hash[4];
		uint8_t mask;
	};
	std::unordered_map<std::string, Info> cache;
	std::vector<abb_succ_info> info;
	std::vector<unsigned> len;
	std::vector<uint64_t> self;
	std::string kbuf;
	out << "digraph g {\n";
	if (verbose)
		std::cerr << "Generating GraphViz output...\n";
	auto expand = [&](const std::vector<std::string>& frontier) { // out-edges of the frontier vertices the cache does not hold
		kbuf.clear();
		std::vector<const std::string*> miss;
		for (const std::string& u : frontier)
			if (!cache.count(u)) {
				miss.push_back(&u);
				kbuf += u;
			}
		if (miss.empty())
			return;
		info.resize(miss.size() * kChain);
		len.resize(miss.size());
		self.resize(miss.size());
		successors(kbuf.data(), miss.size(), kChain, info.data(), len.data(), self.data());
		for (size_t i = 0; i < miss.size(); ++i) {
			std::string cur = *miss[i];
			for (unsigned s = 0; s < len[i]; ++s) {
				const abb_succ_info& a = info[i * kChain + s];
				Info v;
				memcpy(v.hash, a.hash, sizeof v.hash);
				v.mask = a.mask;
				cache.emplace(cur, v);
				if (s + 1 < len[i]) { // the chain went on through the only out-edge
					unsigned b = 0;
					while (!((a.mask >> b) & 1))
						++b;
					cur.erase(0, 1);
					cur += BASES[b];
				}
			}
		}
	};
	auto bfs = [&](const std::string& start, uint64_t startHash) { // breadthFirstSearchImpl (Graph/BreadthFirstSearch.h:100-181), directed
		if (!visited.insert(startHash).second)
			return; // black: explored by an earlier search
		++nodes;
		out << '\t' << start << ";\n";
		std::vector<std::string> frontier{ start }, next;
		while (!frontier.empty()) {
			expand(frontier);
			next.clear();
			for (const std::string& u : frontier) {
				auto it = cache.find(u);
				const Info v = it->second;
				cache.erase(it);
				for (unsigned b = 0; b < 4; ++b) {
					if (!((v.mask >> b) & 1))
						continue;
					std::string t = u.substr(1) + BASES[b];
					++edges;
					out << '\t' << u << " -> " << t << ";\n"; // examine_edge
					if (visited.insert(v.hash[b]).second) {     // white: discover_vertex
						++nodes;
						out << '\t' << t << ";\n";
						next.push_back(std::move(t));
					}
				}
			}
			frontier.swap(next);
		}
	};
	std::vector<uint8_t> flag, valid;
	for_each_batch([&](const host::ReadBatch& b) {
		uint64_t slots = 0;
		for (size_t i = 0; i < b.size(); ++i) {
			const uint64_t L = b.offsets[i + 1] - b.offsets[i];
			slots += L >= k ? L - k + 1 : 0;
		}
		flag.resize(slots + 1);
		valid.resize(slots + 1);
		contains_reads(b.bases.data(), b.offsets.data(), b.size(), flag.data(), valid.data(), slots + 1);
		// trimSeq (:399-447) for every read, then the two start k-mers of each surviving read
		std::vector<std::string> starts;
		std::vector<size_t> startOf(b.size(), SIZE_MAX);
		uint64_t s0 = 0;
		for (size_t i = 0; i < b.size(); ++i) {
			const uint64_t L = b.offsets[i + 1] - b.offsets[i];
			const uint64_t w = L >= k ? L - k + 1 : 0;
			const uint64_t UNSET = UINT64_MAX;
			uint64_t prevPos = UNSET, matchStart = UNSET, matchLen = 0, maxStart = UNSET, maxLen = 0;
			for (uint64_t p = 0; p < w; ++p) {
				if (!valid[s0 + p])
					continue;
				const bool c = flag[s0 + p];
				if (!c || (prevPos != UNSET && p - prevPos > 1)) {
					if (matchStart != UNSET && matchLen > maxLen) {
						maxLen = matchLen;
						maxStart = matchStart;
					}
					matchStart = UNSET;
					matchLen = 0;
				}
				if (c) {
					if (matchStart == UNSET)
						matchStart = p;
					matchLen++;
				}
				prevPos = p;
			}
			if (matchStart != UNSET && matchLen > maxLen) {
				maxLen = matchLen;
				maxStart = matchStart;
			}
			s0 += w;
			if (maxLen == 0)
				continue;
			const char* seq = b.bases.data() + b.offsets[i] + maxStart;
			const uint64_t tl = maxLen + k - 1;
			startOf[i] = starts.size();
			starts.emplace_back(seq, k);
			std::string rc(k, 'N'); // first k-mer of the reverse complement = complement of the last k bases, reversed
			for (unsigned j = 0; j < k; ++j) {
				const char ch = seq[tl - 1 - j];
				rc[j] = ch == 'A' ? 'T' : ch == 'C' ? 'G' : ch == 'G' ? 'C' : 'A';
			}
			starts.push_back(std::move(rc));
		}
		std::vector<uint64_t> startHash(starts.size());
		if (!starts.empty()) {
			kbuf.clear();
			for (const std::string& st : starts)
				kbuf += st;
			info.resize(starts.size());
			len.resize(starts.size());
			successors(kbuf.data(), starts.size(), 1u, info.data(), len.data(), startHash.data());
		}
		for (size_t i = 0; i < b.size(); ++i) {
			if (startOf[i] != SIZE_MAX) {
				bfs(starts[startOf[i]], startHash[startOf[i]]);
				bfs(starts[startOf[i] + 1], startHash[startOf[i] + 1]);
			}
			if (++readsProcessed % 1000 == 0 && verbose)
				std::cerr << "processed " << readsProcessed << " (k-mers visited: " << nodes << ", edges visited: " << edges << ")\n";
		}
	});
	out << "}\n";
	if (verbose)
		std::cerr << "processed " << readsProcessed << " reads (k-mers visited: " << nodes << ", edges visited: " << edges << ")\nGraphViz generation complete\n";
}

} // namespace host
