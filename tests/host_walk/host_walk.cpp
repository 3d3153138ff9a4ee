// tests/host_walk/host_walk.cpp -- TEST INFRASTRUCTURE ONLY (never linked into libabyssb200).
//
// Single-thread emulation harness for the pass-2 graph logic: instantiates the SAME templates
// the CUDA kernels use (abyss_b200/csrc/abb_walk.cuh) over a trivial one-lane context, drives
// them sequentially exactly like BloomDBG::assemble at -j1 (bloom-dbg.h:783-882,972-1089), and
// prints the unitig FASTA.  It lets the traversal logic be debugged against the reference's
// golden FASTA on a machine without a GPU.  The counting filter is built with the C oracle.
//
//   host_walk K KC H COUNTERS TRIM reads.fq [readlog.tsv] > out.fa
#include "../../abyss_b200/csrc/abb_walk.cuh"
extern "C" {
#include "../../oracle/abyss_oracle.h"
}
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <unordered_set>
#include <vector>

using namespace abb;

struct HostCtx {
	unsigned k, trim, H, threshold;
	RollTab rt;
	HashCfg cfg;
	const uint8_t* counters;
	Frame* frames;
	uint64_t* look;
	std::vector<std::unique_ptr<uint8_t[]>> allocs;
	bool fail_ = false;
	unsigned long long probes = 0;

	bool contains(uint64_t h0) const
	{
		for (unsigned i = 0; i < H; ++i)
			if (counters[nth_pos(h0, cfg, i)] < threshold)
				return false;
		return true;
	}
	template <int KW>
	unsigned neighbors(const Vtx<KW>& v)
	{
		++probes;
		unsigned m = 0;
		for (unsigned n = 0; n < 8; ++n) {
			HashPair h = n < 4 ? roll_right(v.h, rt, kmer_first(v.km, k), n) : roll_left(v.h, rt, kmer_last(v.km), n - 4);
			if (contains(h.canonical()))
				m |= 1u << n;
		}
		return m;
	}
	struct Probe {
		unsigned mask;
	};
	template <int KW>
	Probe neighbors_issue(const Vtx<KW>& v) { return Probe{ neighbors(v) }; }
	unsigned neighbors_finish(const Probe& p) { return p.mask; }
	uint64_t rd64(const uint64_t* p) { return *p; }
	void wr64(uint64_t* p, uint64_t v) { *p = v; }
	uint8_t rd8(const uint8_t* p) { return *p; }
	void wr8(uint8_t* p, uint8_t v) { *p = v; }
	void sync() {}
	bool find64(const uint64_t* a, unsigned n, uint64_t key, unsigned stride)
	{
		for (unsigned i = 0; i < n; ++i)
			if (a[(size_t)i * stride] == key)
				return true;
		return false;
	}
	uint8_t* alloc(uint64_t bytes, bool zero)
	{
		allocs.emplace_back(new uint8_t[bytes + 8]);
		if (zero)
			memset(allocs.back().get(), 0, bytes);
		return allocs.back().get();
	}
	void fail(unsigned why) { fail_ = true; fprintf(stderr, "host_walk: scratch overflow %u\n", why); }
	bool failed() const { return fail_; }
	void copy8(uint8_t* d, const uint8_t* s, unsigned n) { if (n) memcpy(d, s, n); }
	void copy8_rev(uint8_t* d, const uint8_t* s, unsigned n) { for (unsigned i = 0; i < n; ++i) d[i] = s[n - 1 - i]; }
	void rehash(const uint64_t* o, unsigned ocap, uint64_t* n, unsigned ncap)
	{
		for (unsigned s = 0; s < ocap; ++s)
			if (o[s]) {
				uint64_t t = pathset_slot(o[s], ncap);
				while (n[t])
					t = (t + 1) & (ncap - 1);
				n[t] = o[s];
			}
	}
	void mark_covered(const PathSet& ps, const uint64_t* rh, uint8_t* cov, unsigned nk, const ContigOut& o)
	{
		for (unsigned j = 0; j < nk; ++j) {
			if (cov[j] || !pathset_contains(*this, ps, rh[j]))
				continue;
			if (o.popped_front && rh[j] == o.front_h)
				continue;
			if (o.popped_back && rh[j] == o.back_h)
				continue;
			cov[j] = 1;
		}
	}
};

struct Assembly {
	HostCtx* c;
	std::vector<uint8_t> assembled; // bit filter, size() bits = #counters (bloom-dbg.h:910-911)
	uint64_t mbits;
	std::unordered_set<uint64_t> contigEnd;
	size_t contigID = 0;
	const std::string* readID = nullptr;

	bool inAssembled(uint64_t h0) const
	{
		for (unsigned i = 0; i < c->H; ++i) {
			uint64_t p = nth_pos(h0, c->cfg, i);
			if (!(assembled[p >> 3] & (1u << (p & 7))))
				return false;
		}
		return true;
	}
	void addAssembled(uint64_t h0)
	{
		for (unsigned i = 0; i < c->H; ++i) {
			uint64_t p = nth_pos(h0, c->cfg, i);
			assembled[p >> 3] |= (uint8_t)(1u << (p & 7));
		}
	}
	unsigned minCount(uint64_t h0) const
	{
		unsigned mn = 255;
		for (unsigned i = 0; i < c->H; ++i) {
			unsigned v = c->counters[nth_pos(h0, c->cfg, i)];
			if (v < mn)
				mn = v;
		}
		return mn;
	}
	// outputContig (bloom-dbg.h:538-620)
	void operator()(HostCtx&, unsigned, const ContigOut& o)
	{
		const unsigned k = c->k;
		std::string seq(o.len, 'N');
		for (unsigned i = 0; i < o.len; ++i)
			seq[i] = "ACGT"[o.seq[i]];
		std::vector<uint64_t> hs(o.len - k + 1);
		std::vector<uint64_t> tmp(hs.size() * c->H);
		size_t n = abo_hash_seq(seq.data(), seq.size(), k, c->H, NULL, tmp.data(), NULL);
		if (n != hs.size()) { fprintf(stderr, "host_walk: contig hashing mismatch\n"); exit(3); }
		for (size_t i = 0; i < n; ++i)
			hs[i] = tmp[i * c->H];
		bool redundant = false;
		if (o.len < k + kFpTrim - 1) {
			if (contigEnd.count(hs.front()) && contigEnd.count(hs.back()))
				redundant = true;
			else {
				contigEnd.insert(hs.front());
				contigEnd.insert(hs.back());
			}
		} else {
			redundant = true;
			for (uint64_t h : hs)
				if (!inAssembled(h)) {
					redundant = false;
					break;
				}
		}
		if (redundant)
			return;
		for (uint64_t h : hs)
			addAssembled(h);
		unsigned cov = 0;
		for (uint64_t h : hs)
			cov += minCount(h);
		printf(">%zu %u %u read:%s\n%s\n", contigID, o.len, cov, readID->c_str(), seq.c_str());
		++contigID;
	}
};

template <int KW>
static int run(unsigned k, unsigned kc, unsigned H, uint64_t m, unsigned trim, const char* path, const char* logpath)
{
	std::vector<std::string> ids, seqs;
	{
		std::ifstream in(path);
		std::string l1, l2, l3, l4;
		while (std::getline(in, l1) && std::getline(in, l2)) {
			if (l1[0] == '@') {
				std::getline(in, l3);
				std::getline(in, l4);
			}
			std::string id = l1.substr(1, l1.find_first_of(" \t") == std::string::npos ? std::string::npos : l1.find_first_of(" \t") - 1);
			ids.push_back(id);
			for (auto& ch : l2)
				ch = (char)toupper(ch);
			seqs.push_back(l2);
		}
	}
	std::vector<uint8_t> counters(m, 0);
	for (auto& s : seqs)
		abo_cbf_load_seq(counters.data(), m, s.data(), s.size(), k, H, NULL);

	HostCtx c;
	c.k = k; c.trim = trim; c.H = H; c.threshold = kc;
	c.rt = make_rolltab(k);
	c.cfg.H = H; c.cfg.k = k; c.cfg.mod = make_fastmod(m);
	for (unsigned i = 0; i < kMaxHashes; ++i)
		c.cfg.mult[i] = (uint64_t)i ^ ((uint64_t)k * kMultiSeed);
	c.counters = counters.data();
	std::vector<Frame> frames(kFrameCap);
	std::vector<uint64_t> look(kLookCap);
	c.frames = frames.data();
	c.look = look.data();

	Assembly as;
	as.c = &c;
	as.mbits = m;
	as.assembled.assign(m / 8, 0);
	FILE* log = logpath ? fopen(logpath, "w") : nullptr;
	if (log)
		fprintf(log, "read_id\tresult\n");
	static const char* names[] = { "SHORTER_THAN_K", "NON_ACGT", "BLUNT_END", "NOT_SOLID", "ALL_KMERS_VISITED", "GENERATED_CONTIGS" };
	for (size_t r = 0; r < seqs.size(); ++r) {
		const std::string& s = seqs[r];
		int code;
		std::vector<uint64_t> hs;
		if (s.size() < k)
			code = RC_SHORTER_THAN_K;
		else if (s.find_first_not_of("ACGT") != std::string::npos)
			code = RC_NON_ACGT;
		else {
			// hasBluntEnd (bloom-dbg.h:494-532)
			Vtx<KW> first = vtx_from_codes<KW>((const uint8_t*)s.data(), k, true);
			Vtx<KW> last = vtx_from_codes<KW>((const uint8_t*)s.data() + s.size() - k, k, true);
			bool blunt = !look_ahead(c, first, REV, kFpTrim) || !look_ahead(c, vtx_revcomp(last, k), REV, kFpTrim);
			if (blunt)
				code = RC_BLUNT_END;
			else {
				std::vector<uint64_t> tmp((s.size() - k + 1) * H);
				size_t n = abo_hash_seq(s.data(), s.size(), k, H, NULL, tmp.data(), NULL);
				hs.resize(n);
				for (size_t i = 0; i < n; ++i)
					hs[i] = tmp[i * H];
				bool solid = true, visited = true;
				for (uint64_t h : hs)
					if (!c.contains(h)) {
						solid = false;
						break;
					}
				if (!solid)
					code = RC_NOT_SOLID;
				else {
					for (uint64_t h : hs)
						if (!as.inAssembled(h)) {
							visited = false;
							break;
						}
					if (visited)
						code = RC_ALL_KMERS_VISITED;
					else {
						code = RC_GENERATED_CONTIGS;
						as.readID = &ids[r];
						if (!walk_read<KW>(c, (const uint8_t*)s.data(), (unsigned)s.size(), as)) {
							fprintf(stderr, "host_walk: walk failed on read %zu\n", r);
							return 4;
						}
						c.allocs.clear();
					}
				}
			}
		}
		if (log)
			fprintf(log, "%s\t%s\n", ids[r].c_str(), names[code]);
	}
	if (log)
		fclose(log);
	fprintf(stderr, "host_walk: %zu reads, %zu contigs, %llu neighbour probes\n", seqs.size(), as.contigID, c.probes);
	return 0;
}

int main(int argc, char** argv)
{
	if (argc < 7) {
		fprintf(stderr, "usage: host_walk K KC H COUNTERS TRIM reads.fq [readlog]\n");
		return 2;
	}
	unsigned k = atoi(argv[1]), kc = atoi(argv[2]), H = atoi(argv[3]);
	uint64_t m = strtoull(argv[4], 0, 10);
	unsigned trim = atoi(argv[5]);
	const char* log = argc > 7 ? argv[7] : nullptr;
	const unsigned kw = (2 * k + 63) / 64;
	switch (kw) {
	case 1: return run<1>(k, kc, H, m, trim, argv[6], log);
	case 2: return run<2>(k, kc, H, m, trim, argv[6], log);
	case 3: return run<3>(k, kc, H, m, trim, argv[6], log);
	case 4: return run<4>(k, kc, H, m, trim, argv[6], log);
	default: return run<6>(k, kc, H, m, trim, argv[6], log);
	}
}
